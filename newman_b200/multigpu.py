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


class SharedHostRaster:
    """The host-facing raster of an N-rank frame as ONE pinned buffer every rank can DMA into: a POSIX
    shared-memory file mapped by all ranks of the node and page-locked in each (cudaHostRegister), so
    rank r copies its interleaved rows r, r+N, ... device -> host over its own PCIe link
    (Device.read_rows_pitched) instead of funnelling the whole frame through rank 0's. Rank 0 reads the
    assembled raster after the barrier. (The NCCL alternative is gather_bands.)"""

    def __init__(self, nr, nc, rank, world, device, tag="raster"):
        import mmap
        import os
        self.nr, self.nc, self.rank, self.world = nr, nc, rank, world
        self.nbytes = nr * nc * 8
        name = [f"/dev/shm/newman_b200_{os.getpid()}_{tag}" if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        if world > 1:
            dist.barrier()
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), self.nbytes)
        self.array = np.frombuffer(self._mm, dtype=np.int32).reshape(nr, nc, 2)
        self.ptr = self.array.ctypes.data
        self._registered = False
        rt = torch.cuda.cudart()
        err = rt.cudaHostRegister(self.ptr, self.nbytes, 0)
        self._registered = int(err) == 0
        if world > 1:
            dist.barrier()
        if rank == 0:
            try:
                os.unlink(self.path)      # the mappings keep it alive; nothing is left behind in /dev/shm
            except OSError:
                pass

    def band_ptr(self):
        """Address of this rank's first row; rows are self.world * nc * 8 bytes apart."""
        return self.ptr + self.rank * self.nc * 8, self.world * self.nc * 8

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.ptr)
            self._registered = False
        self.array = None
        try:
            self._mm.close()
        except BufferError:
            pass
        self._f.close()
