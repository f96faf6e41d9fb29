"""Frame orchestration above the device C-ABI: primary reference, glitch re-queue rounds with
secondary references, final rebasing pass — the Python twin of Mandelbrot::renderFrame
(newman_b200/csrc/mandelbrot_host.cpp) with the table source and the "which glitched pixel becomes
the next reference" reduction injectable, so the same loop serves one GPU, N ranks (row-interleaved
bands, torch.distributed) and the CPU-side gloo tests. Plumbing only: no arithmetic on pixels."""
import numpy as np

from . import _lib as L


def floatexp_level(d, force=0):
    """Which form of the tables a view needs (the rule of Mandelbrot::renderFrame): 0 doubles; 1 floatexp
    series (a descended coefficient left double range); 2 also floatexp eps + scaled delta states (pixel
    pitch below 2^-380). `d` = Mandelbrot.host_tables()."""
    fe = 0 if d.get("finite", True) else 1
    e = np.asarray(d["eps_re_e"])
    m = np.asarray(d["eps_re_m"])
    nz = m != 0
    if nz.sum() >= 2:  # pitch = spacing of neighbouring columns ~ largest |eps| / (nc/2)
        if int(e[nz].max()) - int(np.log2(max(len(e) // 2, 1))) < -380:
            fe = 2
    return max(fe, min(int(force), 2))


class TableSet:
    """Descended tables of one reference orbit + eps arrays (numpy or torch tensors).
    fe = 0: doubles. fe >= 1: a/b/c hold mantissas, a_e/b_e/c_e (int32) their exponents.
    fe == 2: eps_re/eps_im hold mantissas, eps_re_e/eps_im_e their exponents."""
    F64 = ("x_hi", "x_lo", "a", "b", "c", "eps_re", "eps_im")

    def __init__(self, d, N, tol, glitch_tol, fe=0):
        self.M, self.has_escape = int(d["M"]), int(d["has_escape"])
        self.fe = int(fe)
        src = {k: k for k in self.F64}
        if self.fe >= 1:
            src.update(a="a_m", b="b_m", c="c_m", a_e="a_e", b_e="b_e", c_e="c_e")
        if self.fe == 2:
            src.update(eps_re="eps_re_m", eps_im="eps_im_m", eps_re_e="eps_re_e", eps_im_e="eps_im_e")
        self.arr = {k: d[s] if s in d else d[k] for k, s in src.items()}
        self.N, self.tol, self.glitch_tol = N, tol, glitch_tol
        self.probe = d.get("probe")

    @staticmethod
    def keys(fe):
        k = list(TableSet.F64)
        if fe >= 1:
            k += ["a_e", "b_e", "c_e"]
        if fe == 2:
            k += ["eps_re_e", "eps_im_e"]
        return k

    def nbytes(self):
        return sum(int(np.prod(v.shape)) * (4 if k.endswith("_e") else 8) for k, v in self.arr.items())

    def tables(self, eps_im_e=None):
        """nm_deep_tables over these arrays; eps_im_e: the row-restricted exponent array (fe == 2)."""
        a = self.arr
        pv = lambda k: L.ptr(a[k]).value if k in a else None
        t = L.DeepTables(M=self.M, N=self.N, has_escape=self.has_escape, flags=0, tol=self.tol,
                         glitch_tol=self.glitch_tol, x_hi=pv("x_hi"), x_lo=pv("x_lo"), a=pv("a"), b=pv("b"), c=pv("c"),
                         a_exp=pv("a_e"), b_exp=pv("b_e"), c_exp=pv("c_e"), eps_re_exp=pv("eps_re_e"),
                         eps_im_exp=None if self.fe != 2 else L.ptr(a["eps_im_e"] if eps_im_e is None else eps_im_e).value)
        t._keep = (a, eps_im_e)
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
    eps_rows(ts, key) -> ts.arr[key] ("eps_im", or "eps_im_e" for scaled frames) restricted to `rows`
        (device or host array)
    Returns dict(rounds=[global pix of each reference], stats=[per-round nm_stats]).
    """
    rows = np.asarray(rows, dtype=np.int64)
    sel = eps_rows or (lambda ts, key: np.ascontiguousarray(ts.arr[key][rows]))
    ts = primary
    pix_list = None
    refs, stats = [], []
    rnd = 0
    gpix = git = np.zeros(0, dtype=np.int32)
    while True:
        mode = L.MODE_REBASE if rnd >= max_secondary else L.MODE_REQUEUE
        if rnd == 0 or len(pix_list):
            eps_im_e = sel(ts, "eps_im_e") if ts.fe == 2 else None
            dev.frame_deep(ts.tables(eps_im_e), ts.arr["eps_re"], sel(ts, "eps_im"),
                           cardioid_mode if rnd == 0 else L.CARDIOID_NONE, mask if rnd == 0 else None, pix_list, mode)
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
