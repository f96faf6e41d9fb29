"""-m gpu: BASELINE.json's full-size configurations through size-independent properties.

The oracles cannot render 8.3 M (cfg2) or 132.7 M (cfg3) samples in test time, so at full size the
checks are (a) a strided sample of the frame against Oracle-P run on exactly those samples — bit for
bit wherever Oracle-P does not flag a glitch against the primary reference (flagged samples are
re-rendered against a secondary reference by the frame and are checked to be resolved), (b) the frame
rendered twice is byte-identical (the work queues, atomics and re-dealing make the ORDER of work
non-deterministic; the raster must not depend on it), (c) every sample resolved and within [0, N],
(d) the executed-iteration count of the primary round equals the oracle's on the sample when both are
restricted to it (pixel-list frame), so the frame and the list path agree with each other too."""
import hashlib
import os

import numpy as np
import pytest

import newman_b200
import oracles
from newman_b200 import _lib as L
from newman_b200 import pipeline, workloads

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def config_of(name):
    """'cfg2' ... or 'cfg5:k' = key frame k of the zoom video"""
    if name.startswith("cfg5:"):
        return workloads.video_frame(int(name.split(":")[1]))
    return workloads.config(name)


@pytest.mark.parametrize("name,n_sample", [("cfg2", 1536), ("cfg3", 384), ("cfg4", 1536), ("cfg5:100", 1536), ("cfg5:300", 1536),
                                           ("cfg5:599", 1536)])
def test_full_size_frame_properties(name, n_sample):
    cfg = config_of(name)
    nr, nc, N = cfg["nr"], cfg["nc"], cfg["N"]
    view = newman_b200.Mandelbrot(nr, nc, N=N, sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    pr = view.find_probe(1)
    print(name, "GPU-assisted probe", pr, "exhaustive", cfg.get("probe"))
    if name == "cfg2":       # the exhaustive search's winner for this grid is pinned in workloads.py
        assert tuple(cfg["probe"]) == pr[:2]
    # (cfg3: the exhaustive criterion is ill-conditioned at 1e-100, see workloads.py; any probe is a valid
    # reference point and the checks below hold for whichever is used)

    def mk(d):
        return pipeline.TableSet(d, N, cfg["tol"], 1e-6, pipeline.floatexp_level(d))
    h = view.host_tables(pr[0], pr[1])
    primary = mk(h)
    assert primary.fe == {"cfg2": 0, "cfg3": 1, "cfg4": 2, "cfg5:100": 0, "cfg5:300": 0, "cfg5:599": 2}[name]
    chain = []

    def discover(gp):
        chain.append(mk(view.host_tables(gp // nc, gp % nc)))
        return chain[-1]

    dev = newman_b200.Device(0)
    try:
        rows = np.arange(nr)
        res = pipeline.render_rounds(dev, primary, discover, nc, rows)
        a = dev.read_rows()
        it = iter(list(chain))
        res2 = pipeline.render_rounds(dev, primary, lambda gp: next(it), nc, rows)
        b = dev.read_rows()
        # (b) order-independence
        assert res2["refs"] == res["refs"]
        assert digest(a) == digest(b)
        for s1, s2 in zip(res["stats"], res2["stats"]):
            assert s1["executed_iters"] == s2["executed_iters"] and s1["glitched"] == s2["glitched"]
        # (c) every sample resolved
        assert a["iterations"].min() >= 0 and a["iterations"].max() <= N
        # (a) strided sample vs Oracle-P on the same samples, primary reference
        total = nr * nc
        pix = (np.arange(n_sample, dtype=np.int64) * (total // n_sample) + (total // n_sample) // 3).astype(np.int32)
        ex = (h["a_e"], h["b_e"], h["c_e"]) if primary.fe else None
        abc = (h["a_m"], h["b_m"], h["c_m"]) if primary.fe else (h["a"], h["b"], h["c"])
        scaled = primary.fe == 2          # floatexp eps + scaled delta states (cfg4, the deep end of cfg5)
        t = oracles.Tables(h["x_hi"], h["x_lo"], *abc, N, cfg["tol"], exps=ex,
                           eps_exps=(h["eps_re_e"], h["eps_im_e"]) if scaled else None)
        o_er, o_ei = (h["eps_re_m"], h["eps_im_m"]) if scaled else (h["eps_re"], h["eps_im"])
        exp, rq_pix, rq_it, st = oracles.p_render_deep(t, o_er, o_ei, pix_list=pix)
        e = exp.reshape(-1)[pix]
        g = a.reshape(-1)[pix]
        ok = e["iterations"] >= 0
        assert ok.mean() > 0.99
        assert np.array_equal(g["iterations"][ok], e["iterations"][ok])
        assert np.array_equal(bits(g["smoothing"][ok]), bits(e["smoothing"][ok]))
        # (d) the same samples as a pixel-list frame: identical records, glitch flags and iteration count
        if name == "cfg2":
            # (e) the adjudication fixture (tests/golden/make_k3_truth.py): the reference's own continuation at 4x its
            # precision on the 6 144 samples bench.py compares. The reference equals it everywhere; the CUDA path may miss
            # only the chaotic-tail samples FP64 perturbation cannot resolve: pinned at >= 99 %.
            z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "k3_truth_cfg2.npz"))
            assert tuple(int(x) for x in z["probe"]) == pr[:2]
            g6 = a.reshape(-1)[z["pix"]]["iterations"]
            truth = z["t1b"]["iterations"]
            bad = np.nonzero(g6 != truth)[0]
            print("cfg2 vs converged continuation: %d of %d samples differ, max |d| %d; the compiled reference differs on %d" %
                  (len(bad), len(truth), int(np.abs(g6 - truth).max()), int((z["ref"]["iterations"] != truth).sum())))
            print("   (sample id, CUDA, truth):", [(int(z["pix"][i]), int(g6[i]), int(truth[i])) for i in bad[:40]])
            assert (g6 == truth).mean() >= 0.99
        dev.frame_deep(primary.tables(), primary.arr["eps_re"], primary.arr["eps_im"], pix_list=pix)
        dev.launch()
        got = dev.read_pixels(pix)
        gpix, git = dev.requeue()
        assert np.array_equal(got["iterations"], e["iterations"])
        assert np.array_equal(bits(got["smoothing"]), bits(e["smoothing"]))
        assert sorted(gpix.tolist()) == sorted(rq_pix.tolist())
        assert dev.stats()["executed_iters"] == st["executed_iters"]
    finally:
        dev.close()
