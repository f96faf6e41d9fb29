"""Locate deep-zoom view centres with high iteration counts (fixture generator for the bench/test
views). Fixed-point big-int complex arithmetic (fast), Newton's method for minibrot nuclei."""
import sys, math, time
from fractions import Fraction

def to_fix(s, bits):
    # decimal string -> fixed point int with `bits` fractional bits
    f = Fraction(s)
    return (f.numerator << bits) // f.denominator

def fix_to_str(x, bits, digits):
    neg = x < 0
    x = abs(x)
    ip = x >> bits
    frac = x - (ip << bits)
    s = str((frac * 10**digits) >> bits).rjust(digits, '0')
    return ('-' if neg else '') + str(ip) + '.' + s

def orbit_min(cr, ci, bits, nmax):
    """iterate z->z^2+c from z=c (X[0]=c convention index 0); report successive minima of |z_n|"""
    zr, zi = cr, ci
    best = None; mins = []
    four = 4 << bits
    for n in range(1, nmax):
        zr2 = (zr*zr) >> bits; zi2 = (zi*zi) >> bits
        if zr2 + zi2 > (1 << (bits+20)): return mins, n
        zi = ((zr*zi) >> (bits-1)) + ci
        zr = zr2 - zi2 + cr
        m = zr*zr + zi*zi
        if best is None or m < best:
            best = m; mins.append((n+1, m))   # period = n+1 in standard (z_0=0) convention
    return mins, nmax

def newton_nucleus(cr, ci, p, bits, iters=60):
    """Newton for z_p(c)=0 (standard convention z_0 = 0)."""
    one = 1 << bits
    for it in range(iters):
        zr, zi = 0, 0
        dr, di = 0, 0
        for k in range(p):
            # d = 2 z d + 1
            ndr = ((zr*dr - zi*di) >> (bits-1)) + one
            ndi = ((zr*di + zi*dr) >> (bits-1))
            dr, di = ndr, ndi
            nzr = ((zr*zr - zi*zi) >> bits) + cr
            nzi = ((zr*zi) >> (bits-1)) + ci
            zr, zi = nzr, nzi
        # c -= z/d
        den = (dr*dr + di*di) >> bits
        if den == 0: return None
        qr = ((zr*dr + zi*di) << 0) // den
        qi = ((zi*dr - zr*di) << 0) // den
        cr -= qr; ci -= qi
        mag = max(abs(qr), abs(qi))
        if mag <= 4: break
    return cr, ci, mag

def nucleus_size(cr, ci, p, bits):
    """Munafo size estimate 1/(beta*lambda^2); returns log10|size| """
    import mpmath
    mpmath.mp.prec = bits
    c = mpmath.mpc(mpmath.mpf(cr) / mpmath.mpf(2)**bits, mpmath.mpf(ci) / mpmath.mpf(2)**bits)
    z = c; l = mpmath.mpc(1); b = mpmath.mpc(1)
    for k in range(1, p):
        l = l * 2 * z
        b = b + 1 / l
        z = z*z + c
    size = 1 / (b * l * l)
    return size

if __name__ == '__main__':
    bits = 600
    c0 = ('-0.743643887037158704752191506114774', '0.131825904205311970493132056385139')
    cr, ci = to_fix(c0[0], bits), to_fix(c0[1], bits)
    t = time.time()
    mins, n = orbit_min(cr, ci, bits, 40000)
    print('orbit ran', n, 'in', time.time()-t)
    for p, m in mins[-12:]:
        print('  period cand', p, 'log10|z|', 0.5*math.log10(m) - bits*math.log10(2))
    for p in (8007,):
        t = time.time()
        r = newton_nucleus(cr, ci, p, bits)
        print('newton', p, 'resid', r[2], 'time', time.time()-t)
        ncr, nci = r[0], r[1]
        print(' nucleus re', fix_to_str(ncr, bits, 120))
        print(' nucleus im', fix_to_str(nci, bits, 120))
        dist = math.hypot(float(ncr-cr), float(nci-ci)) / 2.0**bits
        print(' |nucleus - c0| =', dist)
        t = time.time()
        s = nucleus_size(ncr, nci, p, bits)
        import mpmath
        print(' size', mpmath.nstr(abs(s), 5), 'arg', mpmath.nstr(mpmath.arg(s), 5), 'time', time.time()-t)
