"""Frame orchestration above the device C-ABI: primary reference, glitch re-queue rounds with
secondary references, final rebasing pass — the Python twin of Mandelbrot::renderFrame
(newman_b200/csrc/mandelbrot_host.cpp) with the table source and the "which glitched pixel becomes
the next reference" reduction injectable, so the same loop serves one GPU, N ranks (row-interleaved
bands, torch.distributed) and the CPU-side gloo tests. Plumbing only: no arithmetic on pixels."""
import numpy as np

from . import _lib as L


class TableSet:
    """Descended tables of one reference orbit + eps arrays (numpy float64 or torch tensors)."""

    def __init__(self, d, N, tol, glitch_tol):
        self.M, self.has_escape = int(d["M"]), int(d["has_escape"])
        self.arr = {k: d[k] for k in ("x_hi", "x_lo", "a", "b", "c", "eps_re", "eps_im")}
        self.N, self.tol, self.glitch_tol = N, tol, glitch_tol
        self.probe = d.get("probe")

    def nbytes(self):
        return sum(int(np.prod(v.shape)) * 8 for v in self.arr.values())

    def tables(self):
        a = self.arr
        t = L.DeepTables(M=self.M, N=self.N, has_escape=self.has_escape, reserved=0, tol=self.tol,
                         glitch_tol=self.glitch_tol, x_hi=L.ptr(a["x_hi"]).value, x_lo=L.ptr(a["x_lo"]).value,
                         a=L.ptr(a["a"]).value, b=L.ptr(a["b"]).value, c=L.ptr(a["c"]).value)
        t._keep = a
        return t

    def map(self, fn):
        out = TableSet.__new__(TableSet)
        out.__dict__.update(self.__dict__)
        out.arr = {k: fn(v) for k, v in self.arr.items()}
        return out


def pick_reference(pix, it):
    """Index of the glitched sample flagged earliest, lowest pixel id on ties; (None) if empty."""
    if len(pix) == 0:
        return None
    order = np.lexsort((pix, it))
    return int(order[0])


def local_rows(nr, rank, world):
    """Row-interleaved band of `rank`: rows rank, rank+world, ... (work per row varies by orders of
    magnitude across a frame; interleaving balances it)."""
    return np.arange(rank, nr, world, dtype=np.int64)


def render_rounds(dev, primary, secondary_tables, nc, rows, max_secondary=1, cardioid_mode=L.CARDIOID_NONE,
                  mask=None, reduce_pick=None, eps_rows=None):
    """Run primary + secondary rounds on `dev` for the grid rows `rows` (global row indices).

    primary: TableSet whose eps_im covers ALL grid rows (indexed by global row).
    secondary_tables(global_pix) -> TableSet for the reference at that pixel (same convention).
    reduce_pick(best_iter, best_global_pix, n_local) -> (global_pix or None): the cross-rank reduction
        (identity for one GPU). Called every round by every rank.
    eps_rows(ts) -> eps_im restricted to `rows` (device or host array)
    Returns dict(rounds=[global pix of each reference], stats=[per-round nm_stats]).
    """
    rows = np.asarray(rows, dtype=np.int64)
    sel = eps_rows or (lambda ts: np.ascontiguousarray(ts.arr["eps_im"][rows]))
    ts = primary
    pix_list = None
    refs, stats = [], []
    rnd = 0
    gpix = git = np.zeros(0, dtype=np.int32)
    while True:
        mode = L.MODE_REBASE if rnd >= max_secondary else L.MODE_REQUEUE
        if rnd == 0 or len(pix_list):
            dev.frame_deep(ts.tables(), ts.arr["eps_re"], sel(ts), cardioid_mode if rnd == 0 else L.CARDIOID_NONE,
                           mask if rnd == 0 else None, pix_list, mode)
            dev.launch()
            stats.append(dev.stats())
            gpix, git = dev.requeue()
        else:  # nothing of ours is glitched, but other ranks may still need the collective pick
            gpix = git = np.zeros(0, dtype=np.int32)
        k = pick_reference(gpix, git)
        if k is None:
            cand = (np.iinfo(np.int64).max, np.iinfo(np.int64).max)
        else:
            r_loc, c = divmod(int(gpix[k]), nc)
            cand = (int(git[k]), int(rows[r_loc]) * nc + c)
        chosen = reduce_pick(cand[0], cand[1], len(gpix)) if reduce_pick else (cand[1] if k is not None else None)
        if chosen is None:
            break
        refs.append(chosen)
        ts = secondary_tables(chosen)
        pix_list = np.ascontiguousarray(gpix, dtype=np.int32)
        rnd += 1
    return dict(refs=refs, stats=stats)
