"""Named views (BASELINE.json `configs`, SURVEY.md §8d parametrisation). A view is fully
deterministic: there is no RNG anywhere on this path; the "seed" is the centre digit string.

depth D := scale on the default view: sz.re = 4*D/nc_grid, sz.im = 3*D/nr_grid (mandelbrot.cpp:14).
"""
from fractions import Fraction

# Seahorse-valley filament between the period-998 (size 6e-16) and period-8007 (size 9e-33)
# minibrots, followed down to 1e-50 by tools/find_view.py (greedy high-count auto-zoom). At 1e-50:
# reference orbit M ~ 6.2e4 < N, series skip L ~ 4.7e4, escape counts 5.2e4..6.2e4, no interior
# samples, series coefficients finite (max |C| ~ 4e180) -> inside the reference's working range.
CFG2_CENTER = ("-0.74364388703715870430384830821904672021612495145852386428093",
               "0.13182590420531197076319081924762009190133695173845852602774")


# The same filament followed on to 1e-100 (tools/zoom_view.py, 50 more decades of greedy high-count
# zoom): M ~ 1.34e5, escape counts 1.33e5..1.39e5, no interior samples. Pixel pitch 2.6e-104 — beyond
# the reference's depth limit (SURVEY finding 3): floatexp series, double deltas.
CFG3_CENTER = ("-0.74364388703715870430384830821904672021612495145852387494331251251878512463051109529950754825484219375",
               "0.13182590420531197076319081924762009190133695173845845225394611001630987887560439346518997163860480250")


# ... and on to 1e-400 (150 more levels of the same greedy zoom, two decades per level, 12 x 16 samples):
# the deep end of cfg4 (floatexp series + scaled deltas) and the path the cfg5 zoom video follows.
CFG4_CENTER = ("-0.743643887037158704303848308219046720216124951458523874943312512518785124630511095299507548254842193623512699502325002477007376750175249951002576257399759948750098509924740199245100005050232499982475747373769898997349737375520023512450270198510126242474500000000174249926745100487675767400270025017602019899499848482398492450487625994949499825000001020024499926265025249999762649265075232601500126",
               "0.1318259042053119707631908192476200919013369517384584522539461100163098788756043934651899716386048025742425739925754975745124002473752524245075240125254999999999745099000026262525740000492574744899499848752624745025739923985100507575997573737575747449997499495100252499754974750124240075509899994848487424485000755024992575755098985050257449762600252600759999497625500025007399004975512551239999997575")
# at 1e-400: M ~ 5.07e5, escape counts 5.05e5..5.09e5, ~9e3 delta updates per sample, no interior samples


# Winners of the reference's EXHAUSTIVE findProbe (mandelbrot.cpp:73-95: first probe in scan order with
# the longest arbitrary-precision orbit), found once with the product's own find_probe(0), for
# comparison with the GPU-assisted search the product uses by default. {(nr, nc): (row, col)}.
CFG2_PROBE = {(2160, 3840): (1080, 1916)}    # M = 61 513; 6 840 candidates, 13 s on 8 host cores
# At 1e-100 the criterion itself is ill-conditioned: the winner's orbit is 144 219 long in mpf at the
# reference's precision but 135 735 by FP64 perturbation (a sample on a filament, where the count changes
# at sub-pixel scale), while the candidates perturbation ranks highest (141 613 ...) are 138 429 ... in mpf.
# The GPU-assisted search therefore settles on another probe here (probe search).
CFG3_PROBE = {(8640, 15360): (6480, 5760)}   # M = 144 219; 27 360 candidates, 212 s on 8 host cores


def _dec(fr, digits=40):
    """Fraction -> scientific decimal string (exact to `digits` significant digits)."""
    if fr == 0:
        return "0"
    e = 0
    f = Fraction(fr)
    while f >= 10:
        f /= 10; e += 1
    while f < 1:
        f *= 10; e -= 1
    m = (f.numerator * 10 ** digits) // f.denominator
    s = str(m)
    return f"{s[0]}.{s[1:]}e{e}"


VIDEO_FRAMES = 600      # cfg5: key frames of the zoom video
VIDEO_DEPTH = 150       # ... from depth 1 down to 1e-150
VIDEO_N = 1 << 18       # iteration limit of every frame (the reference holds N fixed during a zoom,
                        # viewer.cpp:283; counts on this zoom path reach ~2e5 at 1e-150)


def video_frame(k, scale=1, frames=VIDEO_FRAMES):
    """Key frame k of cfg5 (BASELINE.json configs[4]): 1280x720, depth D_k = 10^(-150 k / (frames-1)),
    centred on the deep end of the zoom path (CFG4_CENTER), no multisampling."""
    from decimal import Decimal, getcontext
    getcontext().prec = 50
    nr, nc = 720 // scale, 1280 // scale
    d = Decimal(10) ** (Decimal(-VIDEO_DEPTH * k) / Decimal(frames - 1))
    sz = (format(Decimal(4) * d / nc, ".40e"), format(Decimal(3) * d / nr, ".40e"))
    return dict(nr=nr, nc=nc, N=VIDEO_N, sz=sz, center=CFG4_CENTER, tol=1e-10, sc=1, frame=k,
                label=f"cfg5: zoom video key frame {k}/{frames}, 1280x720, depth 1e-{VIDEO_DEPTH * k / (frames - 1):.1f}")


def _probe_for(table, nr, nc, y_mult):
    p = table.get((nr, nc))
    return None if p is None else (p[0] * y_mult + y_mult - 1, p[1])


def config(name, scale=1, y_mult=1):
    """Return dict(nr, nc, N, sz, center, tol, sc, label).
    scale: divide both grid dimensions by `scale` keeping the view extent (test-size variants).
    y_mult: multiply the number of grid rows keeping the extent (weak-scaling workload: N GPUs render
            a y_mult = N times vertically super-sampled raster, rows interleaved across ranks)."""
    if name == "cfg1":
        nr, nc, N, sc = 768 // scale, 1024 // scale, 1024, 1
        return dict(nr=nr * y_mult, nc=nc, N=N, sz=None if y_mult == 1 else (_dec(Fraction(4, nc)), _dec(Fraction(3, nr * y_mult))),
                    center=None, tol=1e-10, sc=sc,
                    label="cfg1: 1024x768 default full view, N=1024, plain double")
    if name == "cfg2":
        nr, nc, N, sc = 2160 // scale, 3840 // scale, 65536, 2
        d = Fraction(1, 10 ** 50)
        sz = (_dec(4 * d / nc), _dec(3 * d / (nr * y_mult)))
        return dict(nr=nr * y_mult, nc=nc, N=N, sz=sz, center=CFG2_CENTER, tol=1e-10, sc=sc,
                    probe=_probe_for(CFG2_PROBE, nr, nc, y_mult),
                    label="cfg2: 1920x1080 zoom at 1e-50, N=65536, series+perturbation tol 1e-10, 2x multisampling")
    if name == "cfg3":
        nr, nc, N, sc = 8640 // scale, 15360 // scale, 1 << 20, 4
        d = Fraction(1, 10 ** 100)
        sz = (_dec(4 * d / nc), _dec(3 * d / (nr * y_mult)))
        return dict(nr=nr * y_mult, nc=nc, N=N, sz=sz, center=CFG3_CENTER, tol=1e-10, sc=sc,
                    # weak-scaling variants (y_mult x the rows): the same reference POINT, i.e. row r*y + y - 1
                    probe=_probe_for(CFG3_PROBE, nr, nc, y_mult),
                    label="cfg3: 3840x2160 beauty render at 1e-100, N=2^20, 4x multisampling (floatexp series)")
    if name == "cfg4":
        nr, nc, N, sc = 4320 // scale, 7680 // scale, 1 << 22, 1
        d = Fraction(1, 10 ** 400)
        sz = (_dec(4 * d / nc), _dec(3 * d / (nr * y_mult)))
        return dict(nr=nr * y_mult, nc=nc, N=N, sz=sz, center=CFG4_CENTER, tol=1e-10, sc=sc,
                    label="cfg4: 7680x4320 view at 1e-400, N=4M (floatexp series + scaled floatexp deltas)")
    if name == "cfg4g":
        # "glitch-heavy region with secondary reference orbits" (BASELINE.json configs[3]): the same view with the glitch
        # rule at 1e-3 — the tolerance perturbation renderers commonly run at; the product's default is 1e-6 — and up to
        # four secondary references before the rebasing pass: tens of thousands of flagged samples per round instead of
        # a few thousand, every reference built by the host in 1 344-bit arithmetic
        c = config("cfg4", scale, y_mult)
        c.update(glitch_tol=1e-3, max_secondary=4,
                 label="cfg4g: cfg4's view with glitch tolerance 1e-3 and up to 4 secondary references (glitch-heavy)")
        return c
    raise KeyError(name)
