"""Multi-rank plumbing (one process per GPU, torch.distributed): the path shards by row-interleaved
bands with NO collective on the per-pixel data path. What is exchanged:
  * tables of a reference orbit: computed on rank 0 (host GMP), broadcast            (broadcast_arrays)
  * which glitched sample becomes the next secondary reference: one MIN all-reduce     (make_reduce_pick)
  * finished raster bands: gathered to rank 0, the host-facing rank                   (gather_bands)
Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

_BIG = np.iinfo(np.int64).max


def make_reduce_pick(world, device):
    """reduce_pick(best_iter, best_global_pix, n_local) -> global pixel of the next reference or None.
    Rule (same as Mandelbrot::renderFrame): earliest flagged iteration, lowest pixel id on ties."""
    def reduce_pick(best_iter, best_pix, n_local):
        key = _BIG if best_iter == _BIG else (int(best_iter) << 40) | int(best_pix)
        if world > 1:
            t = torch.tensor([key], dtype=torch.int64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            key = int(t.item())
        return None if key == _BIG else key & ((1 << 40) - 1)
    return reduce_pick


def broadcast_arrays(arrays, meta, rank, world, device, sizes_from_meta):
    """rank 0 passes dict name->np.float64 array and an int64 meta list; others pass None.
    Returns (dict name->torch tensor on `device`, meta list)."""
    m = torch.tensor(meta if rank == 0 else [0] * len(meta), dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(m, 0)
    meta = [int(x) for x in m.tolist()]
    out = {}
    for name, n in sizes_from_meta(meta):
        t = (torch.from_numpy(np.ascontiguousarray(arrays[name])).to(device) if rank == 0
             else torch.empty(n, dtype=torch.float64, device=device))
        if world > 1:
            dist.broadcast(t, 0)
        out[name] = t
    return out, meta


def gather_bands(band, nr, rank, world):
    """band: (rows_local, nc, 2) int32 tensor of this rank's interleaved rows. Returns the assembled
    (nr, nc, 2) tensor on rank 0 (None elsewhere)."""
    if world == 1:
        return band
    bufs = [torch.empty_like(band) for _ in range(world)] if rank == 0 else None
    dist.gather(band, bufs, dst=0)
    if rank != 0:
        return None
    full = torch.empty((nr,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
    for r in range(world):
        full[r::world] = bufs[r]
    return full
