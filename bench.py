#!/usr/bin/env python
"""bench.py — newman hot path on B200: pixel-iterations/s (Giter/s) on BASELINE.json configs[1]
("cfg2": 1920x1080 zoom at 1e-50, N = 65536, series + perturbation at tolerance 1e-10, 2x
multisampling => 2160 x 3840 grid samples).

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one process per GPU)
  python bench.py --impl reference [...]                      the reference's own CPU code (oracle/_ref)

A "step" = one full frame: series skip (K2) + perturbation (K3) for every sample, including the
glitch re-queue rounds against secondary reference orbits and the float32 smoothing fix-ups.
  value    : executed pixel-iterations (K3 delta updates) / device time, tables resident in HBM
  e2e      : same frames through the device-level C-ABI with HOST (pinned) buffers: tables H2D, raster D2H, wall clock;
             the host arbitrary-precision work (probe search, orbit, series) is outside it (host_precompute_s)
  e2e_view : the same frames through the drop-in entry point itself — nmv_render (N = 1) / nmm_render (N > 1): host
             precompute INSIDE the timed region, raster returned to host memory
The same JSON line carries, measured in the same run (fixed 3 warm-up + 3 timed steps each, stated per block):
  configs  : (N = 1) the other BASELINE configs — cfg1, cfg3, cfg4, cfg5 — value / ms_per_step / frac / e2e
  strong   : (N > 1) ONE cfg3 frame (8640 x 15360 samples, the north-star beauty render) split over the N ranks inside
             libnewman_b200.so (nmm_render: bands of 4 grid rows dealt round-robin, tables by ncclBroadcast, next-reference
             MIN by ncclAllReduce, bands back to rank 0), next to rank 0's own 1-GPU render of the same frame and
             whether the two rasters are byte-identical

N > 1 main line (weak scaling): the same view rendered with N x as many grid rows (vertical super-sampling
x N); rank r renders the interleaved rows r, r+N, ...; tables are computed on rank 0 and broadcast
(NCCL), the raster bands are gathered to rank 0. No collective sits on the per-pixel data path.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# k3_fast runs the delta recurrence alone: 2 DADD + 4 DFMA = 6 FP64 instructions, 10 flops per iteration; the
# glitch / escape decisions ride on the integer pipe (k3_filter.cuh). (The simple kernel k3_level, --k3-group 0,
# forms z and |z|^2 as well: 10 instructions, 17 flops.)
K3_INST_PER_ITER = 6
K3_FLOPS_PER_ITER = 10
K3_INST_PER_ITER_SIMPLE = 10
K3_FLOPS_PER_ITER_SIMPLE = 17
# DRAM bytes per state and full-level launch of the dominant kernel, from this round's ncu --set full capture
# profiles/r02h_k3_fast_cfg3_full_level_metrics.txt (1.063922 GB read + 1.017441 GB written for 33 177 600 states; round 1:
# 1.063988 + 1.018290): the 32-byte state of every sample in and out once = the algorithmic traffic (64 B), nothing
# re-read. It comes from a profiler run, so it is a constant of the kernel generation, not of this bench run; the number
# that moves with the run is roofline.achieved.
K3_DRAM_B_PER_STATE = (1.063922e9 + 1.017441e9) / 33177600
K2_INST_PER_EVAL = 19   # 12 DMUL + 6 DADD + ... (k2_series.cuh phase-1 body incl. the compare)


def mix_ceiling(hw, simple, executed, k_ms, peak_dadd, peak_dfma3):
    """What the measured instruction rates allow for k3_fast's instruction mix. A warp-wide FP64 instruction reads one
    64-bit register operand per cycle unless the operand-reuse cache of its slot already holds it, so a DFMA with three
    fresh operands issues at 2/3 of the pipe rate (tools/issue_probe.py, nm_fp64_peak kinds 12/13). In the order
    tools/sass_resched.py gives the quiet block — t1, t2, DADD, DADD, ndr, ndi — one DFMA per iteration reads three
    fresh operands and three read two: per iteration 2 DADD + 1 DFMA(3 fresh) + 3 DFMA(2 fresh). (Round 1 / ptxas's own
    order: 4 DFMA with three fresh operands, ceiling 1 / (2/dadd + 4/dfma3) = 2 320 Giter/s.) Iterations/s as a
    fraction of that ceiling says how close the kernel is to what the SM can deliver for THIS arithmetic; `frac`
    (against the plain DADD rate) is the stricter, instruction-mix-agnostic number."""
    peak_dfma2 = None
    if isinstance(peak_dfma3, tuple):
        peak_dfma3, peak_dfma2 = peak_dfma3
    if hw or simple or not (k_ms > 0 and peak_dadd and peak_dfma3):
        return {}
    rate = executed / (k_ms * 1e-3)
    out = {"peak_dfma_3_distinct_operands_ginst": peak_dfma3 / 1e9}
    if peak_dfma2:
        ceiling = 1.0 / (2.0 / peak_dadd + 1.0 / peak_dfma3 + 3.0 / peak_dfma2)       # iterations/s
        out["peak_dfma_2_distinct_operands_ginst"] = peak_dfma2 / 1e9
        out["mix_ceiling_ptxas_order_giter_s"] = 1.0 / (2.0 / peak_dadd + 4.0 / peak_dfma3) / 1e9
    else:
        ceiling = 1.0 / (2.0 / peak_dadd + 4.0 / peak_dfma3)
    out.update({"mix_ceiling_giter_s": ceiling / 1e9, "frac_of_mix_ceiling": rate / ceiling})
    return out


def resched_note():
    """what tools/sass_resched.py did to the library that is loaded (written next to it at build time), or None"""
    try:
        from newman_b200 import _lib
        rows = [l.strip() for l in open(_lib.LIB_PATH + ".resched.txt") if "k3_fastILi4" in l]
        return [re.sub(r"^_ZN2nm7k3_fastILi(\d)ELb(\d)EEEvNS_8K3ParamsEPNS_8PixStateE", lambda m: "k3_fast<%s,%s>" % (m.group(1), "scaled" if m.group(2) == "1" else "plain"), l) for l in rows]
    except (IOError, OSError):
        return None


METRIC = "executed pixel-iterations/sec"
UNIT = "Giter/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=int, default=1, help="divide the grid by this (debug only; invalid as a bench number)")
    ap.add_argument("--cpu-sample", type=int, default=6144, help="pixels in the timed CPU-reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ymult", type=int, default=0, help="debug: force the weak-scaling row multiplier")
    ap.add_argument("--gather", default="shm", choices=["shm", "nccl"],
                    help="N>1 e2e: how bands reach the host raster: per-rank D2H into a shared pinned raster, or NCCL gather to rank 0")
    ap.add_argument("--frames", type=int, default=24, help="cfg5: key frames rendered per step (evenly spaced over the 600)")
    ap.add_argument("--floatexp", type=int, default=0, help="force the floatexp level (1 series, 2 + scaled deltas)")
    ap.add_argument("--k3-group", type=int, default=-1, help="pixels per lane in k3_fast (4, 2; 0 = simple kernel)")
    ap.add_argument("--no-extras", action="store_true", help="main line only: no configs / strong / e2e_view blocks")
    ap.add_argument("--extra-steps", type=int, default=3, help="timed steps of each extra block (3 warm-up steps each)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Spawn the sampler (nvidia-smi takes a few 100 ms to deliver its first line: call this well
        before the timed region and bracket the region itself with mark_begin()/mark_end())."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.03]   # a line reports the ~20 ms before it arrives
        if not inside and self.rows:                                    # region shorter than a sampling period
            inside = [min(self.rows, key=lambda tr: abs(tr[0] - t1))[1]]
        sm, mx, reasons, power = [], [], set(), []
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def _have_ref():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles
    return oracles.have_ref()


def load_workloads():
    """newman_b200/workloads.py loaded by path: the view definitions (strings and sizes) without importing the package,
    so that the reference arm never touches the product (its .so is not even mapped)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("nm_workloads", os.path.join(ROOT, "newman_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def config_dict(cfg, world=1):
    """`config` of the JSON line — the same keys and values on both arms (the reference arm renders world = 1)."""
    return {"workload": cfg["label"] + (f", rows x{world} (weak scaling)" if world > 1 else ""),
            "grid": [cfg["nr"], cfg["nc"]], "N": cfg["N"], "tol": cfg["tol"], "glitch_tol": cfg.get("glitch_tol", 1e-6),
            "max_secondary": cfg.get("max_secondary", 1),
            "parallelism": f"row-interleaved bands x{world}",
            "l2": "working set per step (state queues + raster) exceeds L2; tables are meant to be L2/SMEM resident"}


def view_for(cfg):
    import newman_b200
    return newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])


def bench_sample(nr, nc, n_pix):
    """the strided sample both CPU legs time and tests/golden/make_k3_truth.py adjudicates"""
    total = nr * nc
    return (np.arange(n_pix, dtype=np.int64) * (total // n_pix) + (total // n_pix) // 3).astype(np.int32)


# ---------------------------------------------------------------------------------------------
_PORT = {}


def _port_worker(pix):
    """Oracle-P (the CPU restatement, oracle/oracle_p.c) on a pixel list: the CPU baseline for views the
    compiled reference cannot render (pixel pitch < ~1e-97: SIGFPE, SURVEY.md finding 3)."""
    import oracles
    t, er, ei = _PORT["args"]
    t0 = time.perf_counter()
    out, rq_pix, rq_it, st = oracles.p_render_deep(t, er, ei, pix_list=np.ascontiguousarray(pix, dtype=np.int32), mode=1)
    secs = time.perf_counter() - t0
    return out.reshape(-1)[pix], st["executed_iters"], secs


def port_inputs_from_host_tables(cfg, h, fe):
    """Oracle-P inputs from the product's host tables (our arm's cpu_baseline leg when the view is beyond Oracle-R)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles
    exps = (h["a_e"], h["b_e"], h["c_e"]) if fe >= 1 else None
    a, b, c = (h["a_m"], h["b_m"], h["c_m"]) if fe >= 1 else (h["a"], h["b"], h["c"])
    t = oracles.Tables(h["x_hi"], h["x_lo"], a, b, c, cfg["N"], cfg["tol"], exps=exps,
                       eps_exps=(h["eps_re_e"], h["eps_im_e"]) if fe == 2 else None)
    return t, (h["eps_re_m"] if fe == 2 else h["eps_re"]), (h["eps_im_m"] if fe == 2 else h["eps_im"])


def port_inputs_from_reference(cfg, probe):
    """Oracle-P inputs from the COMPILED REFERENCE's own orbit and series (Oracle-R computes them at any depth; only
    its per-pixel code stops below ~1e-97): what the reference arm uses, so that it never loads the product."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles
    v = oracles.RefView(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    v.precompute_at(probe[0], probe[1])
    t = v.tables()
    finite = all(np.isfinite(x).all() for x in (t.a, t.b, t.c))
    er, ei = v.eps()
    nz = er[er != 0]
    pitch_small = len(nz) >= 2 and (np.log2(np.abs(nz).max()) - np.log2(max(cfg["nc"] // 2, 1)) < -380 or not np.isfinite(nz).all())
    fe = 2 if (pitch_small or np.abs(er).max() == 0.0) else (0 if finite else 1)
    if fe == 0:
        return t, er, ei, fe
    if fe == 1:
        return v.tables_fe(), er, ei, fe
    tf, (mre, mim) = v.tables_fe(scaled=True)
    return tf, mre, mim, fe


def cpu_port_sample(cfg, inputs, n_pix, procs):
    """cpu_baseline kind "port": Oracle-P, one process per core, strided sample, single rebasing pass."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles
    oracles.oraclep()
    pix = bench_sample(cfg["nr"], cfg["nc"], n_pix)
    _PORT["args"] = inputs
    chunks = [pix[i::procs] for i in range(procs)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_port_worker, chunks)
    wall = time.perf_counter() - t0
    it = np.zeros(n_pix, dtype=np.int32)
    for i, (o, _, _) in enumerate(res):
        it[i::procs] = o["iterations"]
    return dict(pix=pix, it=it, executed=int(sum(r[1] for r in res)), busy=max(r[2] for r in res), wall=wall, procs=procs,
                kind="port")


def cpu_reference_sample(cfg, probe, n_pix, procs):
    """Time the reference's own per-pixel code (Oracle-R: /root/reference/mandelbrot.cpp compiled
    unmodified) on a strided sample of the workload, one process per core (the reference is
    single-threaded and its GMP precision is process-global). Returns dict for the JSON line."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles
    if not oracles.have_ref():
        return None
    pix = bench_sample(cfg["nr"], cfg["nc"], n_pix)
    chunks = [pix[i::procs] for i in range(procs)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, [(cfg, probe, c) for c in chunks])
    wall = time.perf_counter() - t0
    it = np.zeros(n_pix, dtype=np.int32)
    Ls = np.zeros(n_pix, dtype=np.int32)
    sm = np.zeros(n_pix, dtype=np.float32)
    busy = 0.0
    setup = 0.0
    for i, (o, L, secs, pre, _hw) in enumerate(res):
        it[i::procs] = o["iterations"]; sm[i::procs] = o["smoothing"]; Ls[i::procs] = L
        busy = max(busy, secs); setup = max(setup, pre)
    N = cfg["N"]
    if res[0][4]:   # plain-double path: it+1 loop trips when escaped, N otherwise (cardioid skips not separable here)
        executed = int(np.where(it < N, it + 1, N).sum(dtype=np.int64))
    else:           # loop 'for (i = d.size(); i < N; i++)' (mandelbrot.cpp:212): it-L+1 trips if escaped, N-L if not
        executed = int(np.where(it >= Ls, np.where(it < N, it - Ls + 1, N - Ls), 0).sum(dtype=np.int64))
    return dict(pix=pix, it=it, sm=sm, L=Ls, executed=executed, effective=int(np.minimum(it, N).sum()), wall=wall,
                busy=busy, setup=setup, procs=procs, kind="reference")


def _ref_worker(a):
    cfg, probe, pix = a
    import oracles
    t0 = time.perf_counter()
    v = oracles.RefView(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    hw = v.use_hardware()
    L = np.zeros(len(pix), dtype=np.int32)
    if not hw:
        v.precompute_at(probe[0], probe[1])  # same reference point findProbe selects (tests pin that)
        t = v.tables()
        er, ei = v.eps()
        ot = t.op()
        import ctypes as C
        P = oracles.oraclep()
        for k, p in enumerate(pix):
            r, c = divmod(int(p), cfg["nc"])
            L[k] = P.oraclep_series_L(C.byref(ot), er[c], ei[r], None, None)
    pre = time.perf_counter() - t0
    out, secs = v.compute_pixels(pix)
    return out, L, secs, pre, hw


def truth_on_sample(cfg, pix, gpu_it, ref_it):
    """parity_on_sample beyond the equal-count fraction: who is right where the CUDA path and the compiled reference
    differ. tests/golden/k3_truth_cfg2.npz (generator: tests/golden/make_k3_truth.py) holds, for exactly this sample of
    cfg2, the reference's own continuation run at 2x and 4x its precision (equal everywhere: converged). Other
    workloads have no such fixture: only the differences are described."""
    d = {}
    if ref_it is not None:
        diff = np.abs(gpu_it.astype(np.int64) - ref_it.astype(np.int64))
        d.update(max_abs_diff=int(diff.max()), n_diff=int((diff != 0).sum()),
                 median_abs_diff_where_different=float(np.median(diff[diff != 0])) if (diff != 0).any() else 0.0)
    fn = os.path.join(ROOT, "tests", "golden", "k3_truth_cfg2.npz")
    if cfg["label"].startswith("cfg2") and os.path.exists(fn):
        z = np.load(fn)
        if len(z["pix"]) == len(pix) and np.array_equal(z["pix"], pix):
            t = z["t1b"]["iterations"]
            d.update(truth="the reference's phase-3 continuation at 4x the view's precision (converged: equals the 2x run on "
                           "every sample); tests/golden/make_k3_truth.py",
                     truth_agree_gpu=float((gpu_it == t).mean()), max_abs_diff_gpu_vs_truth=int(np.abs(gpu_it - t).max()))
            if ref_it is not None:
                d.update(truth_agree_ref=float((ref_it == t).mean()), max_abs_diff_ref_vs_truth=int(np.abs(ref_it - t).max()),
                         ref_matches_fixture=bool(np.array_equal(ref_it, z["ref"]["iterations"])))
            d["explanation"] = ("where the two differ the reference is right: FP64 perturbation carries a relative error of ~5e-15 "
                                "in delta after ~7000 iterations, and samples whose last few hundred iterations are chaotic "
                                "amplify that by > 1e10 (DESIGN.md section 6)")
    return d


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, all host cores, bounded sample/step. Nothing of the
    product is imported or loaded: the views come from workloads.py loaded by path, every table from Oracle-R."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workloads = load_workloads()
    if args.workload == "cfg5":   # one representative key frame of the video (depth 1e-75), bounded sample
        cfg = workloads.video_frame(workloads.VIDEO_FRAMES // 2 - 1, scale=args.scale)
    else:
        cfg = workloads.config(args.workload, scale=args.scale, y_mult=1)
    procs = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles
    if not oracles.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (the compiled reference) was not built on this box"}))
        return 0
    t0 = time.perf_counter()
    hw = cfg["sz"] is None
    probe, inputs, fe = (0, 0), None, 0
    if not hw:
        # the reference point: the exhaustive findProbe winner pinned in workloads.py where there is one (cfg2, cfg3),
        # else the centre sample (the reference's own findProbe would take hours here and is excluded on both arms)
        probe = tuple(cfg.get("probe") or (cfg["nr"] // 2, cfg["nc"] // 2))
        t, er, ei, fe = port_inputs_from_reference(cfg, probe)
        inputs = (t, er, ei)
    pre_s = time.perf_counter() - t0
    use_port = fe >= 1   # below ~1e-97 the compiled reference's per-pixel code raises SIGFPE: its CPU port instead
    n_pix = max(procs, args.cpu_sample // 2)
    vals, walls = [], []
    for s in range(args.warmup + args.steps):
        r = cpu_port_sample(cfg, inputs, n_pix, procs) if use_port else cpu_reference_sample(cfg, probe, n_pix, procs)
        if s >= args.warmup:
            vals.append(r["executed"] / r["busy"] / 1e9)
            walls.append(r["busy"])
    value = statistics.mean(vals)
    frac = n_pix / (cfg["nr"] * cfg["nc"])
    kind = "port" if use_port else "reference"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(walls), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "mpf+f64", "data": "synthetic",
        "config": config_dict(cfg, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{n_pix} strided samples of the {cfg['nr']}x{cfg['nc']} raster per step "
                                   f"({frac:.2e} of a frame), " +
                                   ("Oracle-P (series scan + FP64 perturbation; the compiled reference cannot render this view)"
                                    if use_port else "per-pixel getIterations") + "; probe search excluded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frame_s_extrapolated": statistics.mean(walls) / frac, "host_precompute_s": pre_s, "gpu_launches": 0,
        "probe": [int(probe[0]), int(probe[1])],
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
class Env:
    """One rank of the bench: its GPU context, stream, process group and the measured FP64 issue peaks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import newman_b200
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.devt = torch.device("cuda", self.local)
        self.dev = newman_b200.Device(self.local)
        self.stream = torch.cuda.Stream(device=self.devt)
        torch.cuda.set_stream(self.stream)      # torch ops and NCCL order against the stream our kernels run on
        self.dev.set_stream(self.stream.cuda_stream)
        self.peaks = None
        self.group = None

    def measure_peaks(self):
        """FP64 issue peaks measured live on this box (MEASURED_PEAKS.json has no FP64 entry), once per process."""
        if self.peaks is None:
            d = self.dev
            self.peaks = (d.fp64_peak(0, 1 << 15)[0], d.fp64_peak(1, 1 << 15)[0],
                          (d.fp64_peak(12, 1 << 15)[0], d.fp64_peak(13, 1 << 15)[0]))
            d.sync()
        return self.peaks

    def render_group(self):
        from newman_b200 import multigpu
        if self.group is None:
            self.group = multigpu.RenderGroup(self.local, self.rank, self.world)
        return self.group

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allred(self, x, op=None):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.devt)
        self.dist.all_reduce(t, op=op or self.dist.ReduceOp.SUM)
        return float(t.item())


def measure_frames(env, args, workload, steps, warmup, cpu_baseline, y_mult):
    """The device-level measurement of one workload on env.world ranks (rows x y_mult, interleaved over the ranks):
    `value` with tables resident in HBM and device timers, `e2e` with host buffers and wall clock. Returns the JSON
    line (rank 0; None elsewhere)."""
    import newman_b200
    from newman_b200 import multigpu, pipeline, workloads
    torch, dist = env.torch, env.dist
    rank, world, local, devt, dev, stream = env.rank, env.world, env.local, env.devt, env.dev, env.stream

    cfg = workloads.config(workload, scale=args.scale, y_mult=y_mult)
    if args.floatexp:
        cfg["floatexp"] = args.floatexp
    if args.k3_group >= 0:
        dev.set_option(newman_b200._lib.OPT_K3_GROUP, args.k3_group)
    nr, nc, N = cfg["nr"], cfg["nc"], cfg["N"]
    gtol, max_sec = cfg.get("glitch_tol", 1e-6), cfg.get("max_secondary", 1)
    hw = cfg["sz"] is None
    rows = pipeline.local_rows(nr, rank, world)
    rows_t = torch.as_tensor(rows, device=devt)
    view = view_for(cfg) if rank == 0 else None
    host_pre = [0.0]

    def bcast_tables(row, col):
        """rank 0: host arbitrary-precision work (C++/GMP, threaded) -> device tensors -> NCCL broadcast."""
        if rank == 0:
            t0 = time.perf_counter()
            h = view.host_tables(row, col)
            host_pre[0] += time.perf_counter() - t0
            fe = pipeline.floatexp_level(h, cfg.get("floatexp", 0))
            hs = pipeline.TableSet(h, N, cfg["tol"], gtol, fe)   # picks the double or mantissa/exponent arrays
            meta = torch.tensor([h["M"], h["has_escape"], h["probe"][0], h["probe"][1], fe], dtype=torch.int64, device=devt)
        else:
            hs = None
            meta = torch.zeros(5, dtype=torch.int64, device=devt)
        if world > 1:
            dist.broadcast(meta, 0)
        M, he, pr, pc, fe = [int(x) for x in meta.tolist()]
        sizes = {"x_hi": 2 * (M + he), "eps_re": nc, "eps_im": nr, "eps_re_e": nc, "eps_im_e": nr}
        d = {"M": M, "has_escape": he, "probe": (pr, pc)}
        for k in pipeline.TableSet.keys(fe):
            dt = torch.int32 if k.endswith("_e") else torch.float64
            t = (torch.from_numpy(np.ascontiguousarray(hs.arr[k])).to(devt) if rank == 0
                 else torch.empty(sizes.get(k, 2 * M), dtype=dt, device=devt))
            if world > 1:
                dist.broadcast(t, 0)
            d[k] = t
        torch.cuda.synchronize()       # broadcasts have landed before the C-ABI copies from these buffers
        ts = pipeline.TableSet.__new__(pipeline.TableSet)
        ts.M, ts.has_escape, ts.fe, ts.N, ts.tol, ts.glitch_tol, ts.probe = M, he, fe, N, cfg["tol"], gtol, (pr, pc)
        ts.arr = {k: d[k] for k in pipeline.TableSet.keys(fe)}
        return ts

    reduce_pick = multigpu.make_reduce_pick(world, devt)

    eps_cache = {}

    def eps_rows(ts, key):
        k = (id(ts), key)
        if k not in eps_cache:
            e = ts.arr[key]
            eps_cache[k] = (e[rows_t].contiguous() if isinstance(e, torch.Tensor) and e.is_cuda else
                            e[torch.as_tensor(rows)].contiguous().pin_memory() if isinstance(e, torch.Tensor) else
                            np.ascontiguousarray(e[rows]))
            torch.cuda.synchronize()
        return eps_cache[k]

    stats_total = {}

    def add_stats(res):
        for st in res["stats"]:
            for k, v in st.items():
                stats_total[k] = stats_total.get(k, 0) + v

    # ---- plain-double workload (cfg1) ------------------------------------------------------------
    if hw:
        if rank == 0:
            cre, cim = view.host_coords()
            coords = [torch.from_numpy(cre).to(devt), torch.from_numpy(cim).to(devt)]
        else:
            coords = [torch.empty(nc, dtype=torch.float64, device=devt), torch.empty(nr, dtype=torch.float64, device=devt)]
        if world > 1:
            for t in coords:
                dist.broadcast(t, 0)
        cim_loc = coords[1][rows_t].contiguous()
        coords_h = [coords[0].cpu().pin_memory(), cim_loc.cpu().pin_memory()]

        def frame(device_resident=True):
            c = coords[0] if device_resident else coords_h[0]
            ci = cim_loc if device_resident else coords_h[1]
            dev.frame_hw(c, ci, N)
            dev.launch()
            amb = dev.ambiguous()
            for p in amb:  # rare; the mpf verdict needs the host view (rank 0 holds it for world == 1)
                if view is not None and world == 1 and view.host_in_cardioid(int(p) // nc, int(p) % nc):
                    dev.poke(int(p), N, 0.0)
            st = dev.stats()
            for k, v in st.items():
                stats_total[k] = stats_total.get(k, 0) + v
        chain_dev = chain_host = None
        primary = None
    else:
        # ---- deep workload: tables (rank 0 -> broadcast), discover the secondary-reference chain ----
        # findProbe (mandelbrot.cpp:73-95), GPU-assisted: rank 0 renders the candidates and checks the
        # short-list in arbitrary precision; where workloads.py pins the exhaustive search's winner for
        # this grid the two are compared (probe_matches_exhaustive)
        probe_info = {}
        if rank == 0:
            t0 = time.perf_counter()
            pr = view.find_probe(1)
            host_pre[0] += time.perf_counter() - t0
            probe_info = {"probe": [pr[0], pr[1]], "orbit_len": pr[2], "probe_exact_checks": pr[3],
                          "probe_search_s": time.perf_counter() - t0}
            if cfg.get("probe") and y_mult == 1:
                probe_info["probe_matches_exhaustive"] = tuple(cfg["probe"]) == (pr[0], pr[1])
            pr = (pr[0], pr[1])
        else:
            pr = (0, 0)                       # ignored: rank 0's tables are broadcast
        primary = bcast_tables(pr[0], pr[1])
        chain_dev = []

        def discover(gp):
            ts = bcast_tables(gp // nc, gp % nc)
            chain_dev.append(ts)
            return ts

        res0 = pipeline.render_rounds(dev, primary, discover, nc, rows, max_secondary=max_sec, reduce_pick=reduce_pick,
                                      eps_rows=eps_rows)
        refs = res0["refs"]
        to_host = lambda ts: ts.map(lambda t: t.cpu().pin_memory())
        primary_h = to_host(primary)
        chain_host = [to_host(ts) for ts in chain_dev]

        def frame(device_resident=True):
            chain = chain_dev if device_resident else chain_host
            it = iter(chain)
            res = pipeline.render_rounds(dev, primary if device_resident else primary_h, lambda gp: next(it), nc, rows,
                                         max_secondary=max_sec, reduce_pick=reduce_pick, eps_rows=eps_rows)
            assert res["refs"] == refs, f"secondary reference chain changed between frames: {res['refs']} vs {refs}; " \
                f"glitched {[st['glitched'] for st in res['stats']]}"
            add_stats(res)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(warmup, 3)):
        frame(True)

    # FP64 issue peaks measured live on this box: DFMA, DADD, DFMA reading 3 registers no neighbour shares
    peak_dfma, peak_dadd, peak_dfma3 = env.measure_peaks()

    # ---- timed: device-resident ---------------------------------------------------------------------
    stats_total.clear()
    barrier()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        frame(True)
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    st_dev = dict(stats_total)

    # ---- timed: end to end with host buffers --------------------------------------------------------
    shared = None
    if world > 1:
        lens = [len(pipeline.local_rows(nr, r, world)) for r in range(world)]
        assert len(set(lens)) == 1, "grid rows must divide evenly across ranks"
    if world > 1 and args.gather == "shm":
        shared = multigpu.SharedHostRaster.create(dev, nr, nc, rank, world, devt)   # None: /dev/shm too small
    if shared is not None:
        out_host = torch.from_numpy(shared.array) if rank == 0 else None
    else:
        out_dev = torch.empty((len(rows), nc, 2), dtype=torch.int32, device=devt)
        out_host = torch.empty((nr, nc, 2), dtype=torch.int32).pin_memory() if rank == 0 else None

    overlap = world == 1 or shared is not None   # copy-out of frame k overlaps frame k+1 (nm_read_rows_pitched_async)

    dbg = {"frame": 0.0, "read": 0.0, "drain": 0.0} if os.environ.get("NM_BENCH_DEBUG") else None

    def frame_e2e():
        t_a = time.perf_counter()
        frame(bool(os.environ.get("NM_BENCH_E2E_RESIDENT")))   # (debug: resident tables isolate the upload's share)
        t_b = time.perf_counter()
        if dbg is not None:
            dbg["frame"] += t_b - t_a
        _copy_out()
        if dbg is not None:
            dbg["read"] += time.perf_counter() - t_b

    def _copy_out():
        if world == 1:
            # D2H straight into pinned host memory, started on a second stream from a device-side snapshot: the next
            # frame's table upload and kernels do not wait for it
            dev.read_rows_pitched_async(0, len(rows), out_host.data_ptr(), nc * 8)
        elif shared is not None:
            p0, pitch = shared.band_ptr()                  # every rank: its interleaved rows, own PCIe link,
            dev.read_rows_pitched_async(0, len(rows), p0, pitch)  # straight into the shared pinned host raster
        else:
            dev.read_rows(0, len(rows), out_dev)           # band stays on the device ...
            full = multigpu.gather_bands(out_dev, nr, rank, world)   # ... NCCL gather to the host-facing rank
            if rank == 0:
                out_host.copy_(full, non_blocking=True)    # ... one D2H of the assembled raster
            torch.cuda.synchronize()

    def drain():
        t_a = time.perf_counter()
        if overlap:
            dev.read_wait()                                # every started copy has landed in host memory
        torch.cuda.synchronize()
        if dbg is not None:
            dbg["drain"] += time.perf_counter() - t_a

    frame_e2e()
    drain()
    stats_total.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        frame_e2e()
    drain()
    barrier()                                              # rank 0 holds the last assembled raster after this
    e2e_s = time.perf_counter() - t0
    if dbg is not None:
        print(f"e2e debug rank {rank}: wall {1e3 * e2e_s / steps:.2f} ms/step; host time in frame() {1e3 * dbg['frame'] / (steps + 1):.2f}, "
              f"in copy-out call {1e3 * dbg['read'] / (steps + 1):.2f}, final drain {1e3 * dbg['drain'] / 2:.2f}", file=sys.stderr, flush=True)
    st_e2e = dict(stats_total)

    # ---- reductions over ranks ------------------------------------------------------------------------
    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=devt)
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=devt)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = allmax(ms)
    e2e_s = allmax(e2e_s)
    executed = allsum(st_dev["executed_iters"])          # over all steps, all ranks
    executed_e2e = allsum(st_e2e["executed_iters"])
    evals = allsum(st_dev.get("series_evals", 0))
    launches = allsum(st_dev["kernel_launches"])
    k3_ms = allmax(st_dev.get("ms_k3", 0.0) + st_dev.get("ms_k1", 0.0))
    k2_ms = allmax(st_dev.get("ms_k2", 0.0))

    if rank == 0:
        iters = out_host.numpy()[..., 0]        # a view: {int32 iterations, float32 smoothing} records
        effective = int(np.clip(iters, 0, N).sum(dtype=np.int64))
        value = executed / (ms * 1e-3) / 1e9
        e2e_value = executed_e2e / e2e_s / 1e9
        per_step_tables = 0 if hw else (primary.nbytes() + sum(t.nbytes() for t in chain_dev))
        h2d = (8 * (nc + len(rows))) * world if hw else per_step_tables * world
        d2h = nr * nc * 8
        # roofline of the dominant kernel, FP64 pipe (no tensor cores, not HBM bound)
        simple = args.k3_group == 0
        k3_inst_iter = K3_INST_PER_ITER_SIMPLE if simple else K3_INST_PER_ITER
        k3_flops_iter = K3_FLOPS_PER_ITER_SIMPLE if simple else K3_FLOPS_PER_ITER
        k_inst = executed / world * (k3_inst_iter if not hw else 8)
        k_name = ("k3_level" if simple else "k3_fast") + " (FP64 perturbation)" if not hw else "k1_escape (plain double)"
        k_ms = k3_ms
        if k2_ms > k3_ms:
            k_name, k_inst, k_ms = "k2_series (series scan)", evals / world * K2_INST_PER_EVAL, k2_ms
        achieved = k_inst / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / steps, "timer": "wall clock, barrier+synchronize both sides",
                    "gather": "single D2H" if world == 1 else
                              ("per-rank pitched D2H into a shared pinned host raster" if shared is not None else
                               "NCCL gather to rank 0 + one D2H"),
                    "overlap": "each frame's raster leaves on a second stream from a device-side snapshot while the next frame's "
                               "tables upload and kernels run; all copies have landed before the timer stops" if overlap else "none"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp64_pipe", "kernel": k_name, "achieved": achieved, "peak": peak_dadd / 1e9,
                         "unit": "Ginst/s", "frac": achieved / (peak_dadd / 1e9) if peak_dadd else None,
                         "bound_note": "neither hbm nor tensor: FP64 instructions per 0.06 B of DRAM traffic, no contraction "
                                       "(SURVEY.md 8d); achieved = executed iterations x FP64 instructions per iteration "
                                       "(inst_per_iter) / K3 device time of whole frames (all levels, sparse ones included)",
                         "inst_per_iter": (k3_inst_iter if not hw else 8),
                         **mix_ceiling(hw, simple, executed / world, k_ms, peak_dadd, peak_dfma3),
                         "traffic": None if hw else K3_DRAM_B_PER_STATE * len(rows) * nc,
                         "traffic_note": None if hw else "bytes per full-level launch of k3_fast for this rank's states, from the "
                                         "ncu capture profiles/r02h_k3_fast_cfg3_full_level_metrics.txt (62.7 B per state; algorithmic 64 B): a constant of the kernel "
                                         "generation (a profiler run), not a measurement of this run",
                         "peak_source": "measured live: nm_fp64_peak DADD issue rate (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_dfma_ginst": peak_dfma / 1e9,
                         "fp64_tflops": executed / world * k3_flops_iter / (k3_ms * 1e-3) / 1e12 if (k3_ms > 0 and not hw) else None,
                         "fp64_tflops_peak_dfma": 2 * peak_dfma / 1e12,
                         "kernel_ms_per_step": k_ms / steps, "k2_ms_per_step": k2_ms / steps,
                         "k3_ms_per_step": k3_ms / steps,
                         "sass_resched": None if hw else resched_note()},
            "executed_iters_per_step": executed / steps, "series_evals_per_step": evals / steps,
            "effective_giter_s": effective / (ms / steps * 1e-3) / 1e9,
            "secondary_references": 0 if hw else len(refs), "glitched_per_step": st_dev.get("glitched", 0) / steps,
            "rebased_per_step": st_dev.get("rebased", 0) / steps, "fixups_per_step": st_dev.get("fixups", 0) / steps,
            "host_precompute_s": host_pre[0],
        }
        if not hw:
            line.update(probe_info)
            line["floatexp_level"] = primary.fe
            if args.k3_group >= 0:
                line["config"]["k3_group"] = args.k3_group
        if world == 1 and cpu_baseline:
            procs = os.cpu_count() or 1
            probe = (0, 0) if hw else primary.probe
            if not hw and (primary.fe >= 1 or not _have_ref()):
                r = cpu_port_sample(cfg, port_inputs_from_host_tables(cfg, view.host_tables(probe[0], probe[1]), primary.fe),
                                    max(procs, args.cpu_sample), procs)
            else:
                r = cpu_reference_sample(cfg, probe, max(procs, args.cpu_sample), procs)
            if r is not None:
                pix = r["pix"]
                eq = float((iters.reshape(-1)[pix] == r["it"]).mean())
                line["cpu_baseline"] = {
                    "value": r["executed"] / r["busy"] / 1e9, "unit": UNIT, "cores": r["procs"], "kind": r["kind"],
                    "sample": f"{len(pix)} strided samples of the {nr}x{nc} raster ({len(pix) / (nr * nc):.2e} of a frame), " +
                              ("reference getIterations per pixel" if r["kind"] == "reference" else
                               "Oracle-P per pixel (the compiled reference cannot render this view: SIGFPE below ~1e-97)") +
                              f", {r['busy']:.1f}s busy on the slowest core; probe search excluded",
                    "frame_s_extrapolated": r["busy"] * (nr * nc) / len(pix),
                    "parity_on_sample": {"equal_count_frac": eq, "n": int(len(pix)),
                                         **truth_on_sample(cfg, pix, iters.reshape(-1)[pix], r["it"] if r["kind"] == "reference" else None)}}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not built on this box"}
        if shared is not None:
            shared.close()
        return line
    if shared is not None:
        shared.close()
    return None


# ---------------------------------------------------------------------------------------------
def measure_video(env, args, steps, warmup, n_frames):
    """cfg5: key frames of the zoom video (BASELINE.json configs[4]; viewer.cpp:441-453, 271-285 render one key frame
    per zoom step, recolour it and hand it to VideoZoom::nextFrame, video.cpp:14-34, which makes the 45 in-between
    frames). Frames are independent views, so they are sharded over the ranks with NO collective on the data path
    ("scaling": "strong": the same set of frames whatever N) — longest first by an a-priori cost (deeper frames iterate
    longer), each to the least loaded rank. A step = every rank renders all of its key frames once, resolves each to RGB
    (K4) and produces the in-between frames towards it (K5, left in HBM: encoding stays with the caller). The host
    arbitrary-precision work per frame (GPU-assisted probe search, orbit, series) happens once outside the timed region
    and is reported."""
    import newman_b200
    from newman_b200 import pipeline, workloads
    from newman_b200 import palette as PAL
    torch, dist = env.torch, env.dist
    rank, world, local, devt, dev, stream = env.rank, env.world, env.local, env.devt, env.dev, env.stream

    all_rows = lambda ts, key: ts.arr[key]   # a rank renders whole frames: no row restriction
    total_frames = workloads.VIDEO_FRAMES
    n_sel = min(n_frames, total_frames)
    sel = sorted(set(int(round(i * (total_frames - 1) / max(n_sel - 1, 1))) for i in range(n_sel)))
    # longest-processing-time-first: cost ~ iterations per sample ~ depth (frame index), plain-double frames are cheap
    load = [0.0] * world
    mine = []
    for k in sorted(sel, reverse=True):
        r = min(range(world), key=lambda i: load[i])
        load[r] += 1.0 + 20.0 * k / total_frames
        if r == rank:
            mine.append(k)
    mine.sort()
    pal_rgb = PAL.MultiWaveGenerator(os.path.join(ROOT, "newman_b200", "default.pal")).cache(workloads.VIDEO_N)
    pal_d = torch.from_numpy(np.ascontiguousarray(pal_rgb)).to(devt)
    rgb_prev = [None]
    tween = [None]
    jobs = []
    host_pre = 0.0
    for k in mine:
        cfg = workloads.video_frame(k, scale=args.scale)
        t0 = time.perf_counter()
        view = view_for(cfg)
        view.set_options(device=local)
        nr, nc, N = cfg["nr"], cfg["nc"], cfg["N"]
        job = {"k": k, "cfg": cfg, "view": view, "hw": view.useHardware()}
        if job["hw"]:
            cre, cim = view.host_coords()
            job["coords_h"] = [torch.from_numpy(cre).pin_memory(), torch.from_numpy(cim).pin_memory()]
            job["coords_d"] = [t.to(devt) for t in job["coords_h"]]
        else:
            pr = view.find_probe(1)
            mk = lambda d: pipeline.TableSet(d, N, cfg["tol"], 1e-6, pipeline.floatexp_level(d))
            primary = mk(view.host_tables(pr[0], pr[1]))
            chain = []

            def discover(gp, view=view, mk=mk, chain=chain, nc=nc):
                chain.append(mk(view.host_tables(gp // nc, gp % nc)))
                return chain[-1]
            res0 = pipeline.render_rounds(dev, primary, discover, nc, np.arange(nr), eps_rows=all_rows)
            to_pin = lambda ts: ts.map(lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory())
            job.update(primary_h=to_pin(primary), chain_h=[to_pin(t) for t in chain], refs=res0["refs"], orbit_len=pr[2])
            job["primary_d"] = job["primary_h"].map(lambda t: t.to(devt))
            job["chain_d"] = [t.map(lambda a: a.to(devt)) for t in job["chain_h"]]
        job["out_h"] = torch.empty((nr, nc, 2), dtype=torch.int32).pin_memory()
        job["host_s"] = time.perf_counter() - t0
        host_pre += job["host_s"]
        jobs.append(job)
    torch.cuda.synchronize()

    tot = {}

    def add(st):
        for k, v in st.items():
            tot[k] = tot.get(k, 0) + v

    def render(job, resident, read):
        cfg = job["cfg"]
        nr, nc, N = cfg["nr"], cfg["nc"], cfg["N"]
        if job["hw"]:
            c = job["coords_d"] if resident else job["coords_h"]
            dev.frame_hw(c[0], c[1], N)
            dev.launch()
            for p in dev.ambiguous():
                if job["view"].host_in_cardioid(int(p) // nc, int(p) % nc):
                    dev.poke(int(p), N, 0.0)
            add(dev.stats())
        else:
            it = iter(job["chain_d"] if resident else job["chain_h"])
            res = pipeline.render_rounds(dev, job["primary_d"] if resident else job["primary_h"], lambda gp: next(it), nc,
                                         np.arange(nr), eps_rows=all_rows)
            assert res["refs"] == job["refs"]
            for st in res["stats"]:
                add(st)
            if os.environ.get("NM_BENCH_DEBUG") and not job.get("dbg"):
                job["dbg"] = True
                print("frame", job["k"], "M", job["orbit_len"], "fe", job["primary_h"].fe, "refs", len(job["refs"]),
                      [(round(st["ms_k2"], 2), round(st["ms_k3"], 2), st["series_evals"], st["executed_iters"], st["pixels"])
                       for st in res["stats"]], file=sys.stderr)
        # recolour (viewer.cpp:271-274) and in-between towards this key frame (video.cpp:14-34), all on the device
        rgb = torch.empty((nr, nc, 3), dtype=torch.uint8, device=devt)
        dev.resolve(pal_d, N, sc=1, smooth=True, out=rgb)
        if rgb_prev[0] is not None and rgb_prev[0].shape == rgb.shape:
            if tween[0] is None or tween[0].shape[1:3] != (nr, nc):
                tween[0] = torch.empty((45, nr, nc, 3), dtype=torch.uint8, device=devt)
            dev.video_inbetween(rgb_prev[0], rgb, nr, nc, rate=45, out=tween[0])
            tot["tween_frames"] = tot.get("tween_frames", 0) + 45
        rgb_prev[0] = rgb
        if read:
            dev.read_rows(0, nr, job["out_h"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(warmup, 3)):
        for job in jobs:
            render(job, True, False)
    peak_dfma, peak_dadd, peak_dfma3 = env.measure_peaks()

    tot.clear()
    barrier()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        for job in jobs:
            render(job, True, False)
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    st_dev = dict(tot)

    for job in jobs:
        render(job, False, True)
    tot.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        for job in jobs:
            render(job, False, True)
    barrier()
    e2e_s = time.perf_counter() - t0
    st_e2e = dict(tot)

    def allred(x, op=None):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=devt)
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return float(t.item())

    ms = allred(ms, dist.ReduceOp.MAX)
    e2e_s = allred(e2e_s, dist.ReduceOp.MAX)
    executed = allred(st_dev.get("executed_iters", 0))
    executed_e2e = allred(st_e2e.get("executed_iters", 0))
    launches = allred(st_dev.get("kernel_launches", 0))
    k_ms = allred(st_dev.get("ms_k3", 0.0) + st_dev.get("ms_k1", 0.0), dist.ReduceOp.MAX)
    k2_ms = allred(st_dev.get("ms_k2", 0.0), dist.ReduceOp.MAX)
    host_max = allred(host_pre, dist.ReduceOp.MAX)
    h2d = allred(sum((j["primary_h"].nbytes() + sum(t.nbytes() for t in j["chain_h"])) if not j["hw"] else
                     8 * (j["cfg"]["nr"] + j["cfg"]["nc"]) for j in jobs))
    d2h = allred(sum(j["cfg"]["nr"] * j["cfg"]["nc"] * 8 for j in jobs))
    n_hw = allred(sum(1 for j in jobs if j["hw"]))
    allred_tw = allred(st_dev.get("tween_frames", 0)) / max(steps, 1)
    if rank == 0:
        # mixed frames: plain-double frames execute 8, perturbation frames 6 FP64 instructions per iteration
        inst = executed / world * K3_INST_PER_ITER
        achieved = inst / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        cfg0 = workloads.video_frame(0, scale=args.scale)
        line = {
            "metric": METRIC, "value": executed / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg5: zoom video, {len(sel)} of {total_frames} key frames {cfg0['nc']}x{cfg0['nr']}, depth 1 -> "
                                   f"1e-{workloads.VIDEO_DEPTH}, N={workloads.VIDEO_N}, frames sharded longest-first over {world} GPU(s); each key frame resolved to RGB (K4) and in-betweened x45 (K5) inside the timed region",
                       "frames": sel, "plain_double_frames": int(n_hw), "parallelism": f"frames x{world}",
                       "l2": "each frame's state queues + raster exceed L2 from ~1e-20 on; tables L2/SMEM resident"},
            "e2e": {"value": executed_e2e / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / steps, "timer": "wall clock, barrier+synchronize both sides"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "fp64_pipe", "kernel": "k3_fast/k3_level (FP64 perturbation) + k1_escape on the shallow frames",
                         "achieved": achieved, "peak": peak_dadd / 1e9, "unit": "Ginst/s",
                         "frac": achieved / (peak_dadd / 1e9) if peak_dadd else None, "traffic": None,
                         "inst_per_iter": K3_INST_PER_ITER,
                         **mix_ceiling(False, False, executed / world, k_ms, peak_dadd, peak_dfma3),
                         "peak_source": "measured live: nm_fp64_peak DADD issue rate", "peak_dfma_ginst": peak_dfma / 1e9,
                         "kernel_ms_per_step": k_ms / steps, "k2_ms_per_step": k2_ms / steps},
            "executed_iters_per_step": executed / steps, "frames_per_step": len(sel),
            "frames_per_s_device": len(sel) / (ms / steps * 1e-3), "frames_per_s_e2e": len(sel) / (e2e_s / steps),
            "tween_frames_per_step": allred_tw,
            "host_precompute_s": host_max,
            "host_precompute_note": "slowest rank: GPU-assisted probe search + orbit + series + secondary references of its frames, once",
        }
        return line
    return None


def measure_view(env, args, workload, steps, y_mult=1, band=None):
    """e2e_view: the frames through the drop-in entry point itself — Mandelbrot::precompute() + every row, i.e. nmv_render
    on one GPU / nmm_render on a render group — wall clock, host arbitrary-precision work INSIDE the timed region
    (probe search, orbit, series of every reference), raster returned to host memory. One untimed call first (context
    and buffer creation, NCCL rendezvous)."""
    import newman_b200
    from newman_b200 import workloads
    cfg = workloads.config(workload, scale=args.scale, y_mult=y_mult)
    torch = env.torch
    view = view_for(cfg)
    view.set_options(device=env.local, glitch_tol=cfg.get("glitch_tol", -1.0), max_secondary=cfg.get("max_secondary", -1))
    out = torch.empty((cfg["nr"], cfg["nc"], 2), dtype=torch.int32).pin_memory() if env.rank == 0 else None
    band = band or max(cfg["sc"], 1) * (2 if cfg["sc"] < 4 else 1)
    infos = []

    def once():
        if env.world == 1:
            view.render(out)
            return view.frame_info()
        return env.render_group().render(view, band, out, 1)

    once()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        infos.append(once())
    env.barrier()
    secs = env.allred(time.perf_counter() - t0, env.dist.ReduceOp.MAX if env.world > 1 else None)
    if env.rank != 0:
        return None
    ex = sum(i["executed_iters"] for i in infos)
    return {"value": ex / secs / 1e9, "unit": UNIT, "ms_per_step": 1e3 * secs / steps, "steps": steps, "warmup": 1,
            "host_precompute_s_per_step": sum(i["host_precompute_s"] for i in infos) / steps,
            "device_ms_per_step": sum(i["device_ms"] for i in infos) / steps,
            "references": infos[-1]["references"], "d2h_bytes_per_step": cfg["nr"] * cfg["nc"] * 8,
            "entry_point": "nmv_render (Mandelbrot::precompute + computeRow)" if env.world == 1 else
                           f"nmm_render (render group of {env.world} ranks, bands of {band} rows, bands returned to rank 0 over NCCL)",
            "timer": "wall clock around the call, barrier + synchronize on both sides; includes the host arbitrary-precision work"}


def measure_exact_mode(env, args):
    """The same frame through the drop-in class with Mandelbrot::exact = 1 (untimed region; its own frame time is reported):
    after the frame, the samples whose escape count depends on how the orbit table was rounded are repeated in double-double
    arithmetic (csrc/k3_dd.cuh). Compared on the adjudication fixture's sample with the reference's counts (which equal the
    converged continuation there) — tests/golden/k3_truth_cfg2.npz, the sample bench.py's CPU legs use."""
    from newman_b200 import workloads
    fn = os.path.join(ROOT, "tests", "golden", "k3_truth_cfg2.npz")
    if not os.path.exists(fn):
        return None
    z = np.load(fn)
    cfg = workloads.config("cfg2")
    v = view_for(cfg)
    v.set_options(device=env.local)
    v.set_exact(True)
    v.render()                       # buffers of the refinement pass exist
    t0 = time.perf_counter()
    out = v.render()
    secs = time.perf_counter() - t0
    i = v.frame_info()
    it = out.reshape(-1)[z["pix"]]["iterations"]
    ref, truth = z["ref"]["iterations"], z["t1b"]["iterations"]
    d = np.abs(it.astype(np.int64) - ref.astype(np.int64))
    return {"equal_count_frac": float((it == ref).mean()), "n": int(len(it)), "n_diff": int((d != 0).sum()), "max_abs_diff": int(d.max()),
            "truth_agree_gpu": float((it == truth).mean()), "refined_samples": int(i["refined"]),
            "refined_frac_of_frame": i["refined"] / float(cfg["nr"] * cfg["nc"]),
            "device_ms": i["device_ms"], "refine_device_ms": i["refine_ms"], "frame_s": secs,
            "probe": [int(i["probe_row"]), int(i["probe_col"])], "probe_matches_fixture": [int(i["probe_row"]), int(i["probe_col"])] == [int(x) for x in z["probe"]],
            "note": "Mandelbrot::exact = 1: the frame, a second rendering against the truncated orbit (the sensitivity probe), and the "
                    "double-double pass over the samples whose count differed; reference counts from the fixture (== the compiled "
                    "reference on this sample: ref_matches_fixture above)"}


def measure_strong(env, args, steps):
    """N > 1: ONE frame of the north-star configuration (cfg3: 3840x2160, 4x multisampling => 8640 x 15360 samples, 1e-100,
    N = 2^20; reference viewer.cpp:186-253) split over the ranks inside libnewman_b200.so — strong scaling. Device time =
    the slowest rank's K2 + K3 time of the frame; the same frame rendered by rank 0 alone stands beside it, and the two
    rasters are compared byte for byte."""
    import hashlib
    from newman_b200 import workloads
    torch = env.torch
    cfg = workloads.config("cfg3", scale=args.scale)
    band = cfg["sc"]
    view = view_for(cfg)
    view.set_options(device=env.local)
    out = torch.empty((cfg["nr"], cfg["nc"], 2), dtype=torch.int32).pin_memory() if env.rank == 0 else None
    grp = env.render_group()
    grp.render(view, band, out, 1)      # untimed: buffers, first-touch
    ex0 = grp.exchange_ms()
    env.barrier()
    infos, walls = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        infos.append(grp.render(view, band, out, 1))
        env.barrier()
        walls.append(time.perf_counter() - t0)
    exch = env.allred((grp.exchange_ms() - ex0) / steps, env.dist.ReduceOp.MAX)
    peak_dfma, peak_dadd, peak_dfma3 = env.measure_peaks()
    res = None
    if env.rank == 0:
        sha_n = hashlib.sha256(out.numpy().tobytes()).hexdigest()
        single = view_for(cfg)
        single.set_options(device=env.local)
        single.render(out)               # untimed first call
        t0 = time.perf_counter()
        single.render(out)
        wall1 = time.perf_counter() - t0
        i1 = single.frame_info()
        sha_1 = hashlib.sha256(out.numpy().tobytes()).hexdigest()
        dev_ms = statistics.mean(i["device_ms"] for i in infos)
        ex = infos[-1]["executed_iters"]
        agg = ex / (dev_ms * 1e-3) / 1e9
        res = {
            "workload": cfg["label"], "grid": [cfg["nr"], cfg["nc"]], "N": cfg["N"], "band_rows": band, "n_gpus": env.world,
            "steps": steps, "warmup": 1, "scaling": "strong",
            "device_ms_per_frame": dev_ms, "value": agg, "unit": UNIT,
            "frac": agg * 1e9 * K3_INST_PER_ITER / (env.world * peak_dadd),
            "frac_note": "executed iterations x 6 FP64 instructions / (slowest rank's device time x N x measured DADD issue peak)",
            "frame_s": statistics.mean(walls), "host_precompute_s": statistics.mean(i["host_precompute_s"] for i in infos),
            "exchange_ms_per_frame_slowest_rank": exch, "references": infos[-1]["references"],
            "executed_iters": ex,
            "n1": {"device_ms_per_frame": i1["device_ms"], "value": i1["executed_iters"] / (i1["device_ms"] * 1e-3) / 1e9,
                   "frame_s": wall1, "host_precompute_s": i1["host_precompute_s"], "executed_iters": i1["executed_iters"],
                   "frac": i1["executed_iters"] * K3_INST_PER_ITER / (i1["device_ms"] * 1e-3) / peak_dadd},
            "strong_efficiency_device": i1["device_ms"] / (env.world * dev_ms),
            "raster_sha_matches_n1": sha_n == sha_1, "raster_sha256": sha_n[:16],
            "limit_note": "frame_s is bound by rank 0's host arbitrary-precision work (probe search + orbit + series of each "
                          "reference), which does not shard; device time does",
        }
    env.barrier()
    return res


def compact(line):
    """What the `configs` block keeps of a full line."""
    if line is None:
        return None
    r = line.get("roofline", {})
    d = {"workload": line["config"]["workload"], "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"],
         "steps": line["steps"], "warmup": line["warmup"], "frac": r.get("frac"), "frac_of_mix_ceiling": r.get("frac_of_mix_ceiling"),
         "kernel": r.get("kernel"), "e2e": line.get("e2e"), "gpu_launches": line.get("gpu_launches"),
         "host_precompute_s": line.get("host_precompute_s"), "executed_iters_per_step": line.get("executed_iters_per_step")}
    for k in ("secondary_references", "glitched_per_step", "floatexp_level", "orbit_len", "frames_per_step", "frames_per_s_device",
              "frames_per_s_e2e", "tween_frames_per_step", "e2e_view"):
        if k in line:
            d[k] = line[k]
    return d


def run_ours(args):
    env = Env(args)
    y_mult = args.ymult or env.world
    main_is_video = args.workload == "cfg5"
    if main_is_video:
        line = measure_video(env, args, args.steps, args.warmup, args.frames)
    else:
        line = measure_frames(env, args, args.workload, args.steps, args.warmup, not args.no_cpu_baseline, y_mult)
    if not args.no_extras and args.scale == 1:
        xs = args.extra_steps
        if env.rank == 0 and env.world == 1 and args.workload == "cfg2" and line is not None and "cpu_baseline" in line:
            line["cpu_baseline"]["parity_on_sample"]["exact_mode"] = measure_exact_mode(env, args)
        if not main_is_video:
            ev = measure_view(env, args, args.workload, xs, y_mult=y_mult)
            if line is not None:
                line["e2e_view"] = ev
        if env.world > 1:
            st = measure_strong(env, args, xs)
            if line is not None:
                line["strong"] = st
        elif args.workload == "cfg2":
            cfgs = {}
            for w in ("cfg1", "cfg3", "cfg4", "cfg4g"):
                c = compact(measure_frames(env, args, w, xs, 3, False, 1))
                if w != "cfg4g":
                    c["e2e_view"] = measure_view(env, args, w, {"cfg3": 2, "cfg4": 1}.get(w, xs))   # (cfg4: ~4 s of host work per call)
                cfgs[w] = c
            cfgs["cfg5"] = compact(measure_video(env, args, xs, 3, args.frames))
            line["configs"] = cfgs
            line["configs_note"] = ("the other BASELINE.json configs measured in this run with 3 warm-up steps and the steps given "
                                    "per block; cfg5 = %d of its 600 key frames" % args.frames)
    if env.rank == 0 and line is not None:
        print(json.dumps(line))
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        sys.exit(run_reference(a))
    sys.exit(run_ours(a))
