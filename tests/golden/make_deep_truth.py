"""Brute-force truth for views beyond the reference's depth limit (SURVEY.md finding 3: the compiled
reference dies with SIGFPE below a pixel pitch of ~1e-97, so Oracle-R cannot pin these).

For each view, sampled pixels are iterated directly (z <- z^2 + c from z = c, count = index of the first
iterate with |z|^2 > 2^20, the deep path's convention, mandelbrot.cpp:216-217) in mpmath at a precision
far above what the view needs. Writes tests/golden/deep_truth.json. Needs only mpmath (no reference)."""
import json
import os

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

VIEWS = {
    # name: nr, nc, N, sz (both axes), centre, precision bits for the brute force, samples
    "DEEP-100": dict(nr=96, nc=128, N=2000, sz="7.8125e-103", center=("0", "1"), prec=700, n=48),
    "DEEP-350": dict(nr=96, nc=128, N=4000, sz="7.8125e-353", center=("0", "1"), prec=1600, n=48),
}


def truth(v, seed=1):
    mp.mp.prec = v["prec"]
    nr, nc, N = v["nr"], v["nc"], v["N"]
    sz = mp.mpf(v["sz"])
    c0r, c0i = mp.mpf(v["center"][0]), mp.mpf(v["center"][1])
    rng = np.random.default_rng(seed)
    pts = [(int(rng.integers(nr)), int(rng.integers(nc))) for _ in range(v["n"])]
    # corners + a near-centre pixel. NOT the centre pixel (nr/2-1, nc/2) itself: it is c = i exactly, a
    # pre-periodic point whose exact orbit never escapes but is unstable, so any finite-precision delta
    # (ours, at ~295 iterations) leaves it; a measure-zero artefact of the view, not a parity case.
    pts += [(0, 0), (nr - 1, nc - 1), (nr // 2, nc // 2 - 1)]
    out = []
    for r, c in pts:
        cre = c0r + (c - nc // 2) * sz           # mandelbrot.cpp:271
        cim = c0i + (nr // 2 - r - 1) * sz        # mandelbrot.cpp:275
        zr, zi, it = cre, cim, N
        for i in range(1, N):
            zr, zi = zr * zr - zi * zi + cre, 2 * zr * zi + cim
            if zr * zr + zi * zi > 2 ** 20:
                it = i
                break
        out.append([r, c, it])
    return out


if __name__ == "__main__":
    res = {k: dict(view={a: b for a, b in v.items() if a not in ("prec", "n")}, samples=truth(v)) for k, v in VIEWS.items()}
    with open(os.path.join(HERE, "deep_truth.json"), "w") as f:
        json.dump(res, f, indent=1)
    print({k: len(v["samples"]) for k, v in res.items()})
