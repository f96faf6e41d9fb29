"""Worker of tests/test_gpu_multi.py::test_process_group_equals_single_gpu — launched as
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... tests/multirank_nccl_worker.py
One process per GPU. Every rank builds the same views; the frame is split over the ranks inside libnewman_b200.so
(nmm_render: NCCL broadcast of the tables, 8-byte MIN all-reduce for the next reference, band return to rank 0); rank 0
compares the assembled raster and the resolved RGB with its own single-GPU render of the same view, byte for byte."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import newman_b200  # noqa: E402
from newman_b200 import multigpu, workloads  # noqa: E402
from oracles import KATS  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grp = multigpu.RenderGroup(local, rank, world)
    views = []
    for name in ("KAT-D30", "KAT-S", "KAT-1c"):
        k = dict(KATS[name])
        views.append((name, k, 2))
    c = workloads.config("cfg2", scale=8)
    views.append(("cfg2/8", dict(nr=c["nr"], nc=c["nc"], N=c["N"], sz=c["sz"], center=c["center"], tol=c["tol"]), 2))
    c = workloads.config("cfg4", scale=32)
    views.append(("cfg4/32", dict(nr=c["nr"], nc=c["nc"], N=c["N"], sz=c["sz"], center=c["center"], tol=c["tol"]), 1))
    ok = True
    pal = (np.arange(3 * 65536) * 7 % 256).astype(np.uint8).reshape(-1, 3)
    for name, k, band in views:
        for mode in (multigpu.RenderGroup.RETURN_ROOT,):
            v = newman_b200.Mandelbrot(**k)
            v.set_options(device=local)
            out = np.zeros((k["nr"], k["nc"]), dtype=newman_b200.ESCAPE_DTYPE) if rank == 0 else None
            info = grp.render(v, band, out, mode)
            rgb = np.zeros((k["nr"] // band, k["nc"] // band, 3), dtype=np.uint8) if rank == 0 else None
            grp.resolve(pal[: k["N"]], band, True, rgb, mode)
            if rank == 0:
                single = newman_b200.Mandelbrot(**k)
                single.set_options(device=local)
                want = single.render()
                same = np.array_equal(out.view(np.uint8), want.view(np.uint8))
                want_rgb = single.resolve(pal[: k["N"]], sc=band, smooth=True)
                same_rgb = np.array_equal(rgb, want_rgb)
                si = single.frame_info()
                same_n = info["executed_iters"] == si["executed_iters"] and info["references"] == si["references"]
                print(f"{name}: raster {'identical' if same else 'DIFFERS'}, rgb {'identical' if same_rgb else 'DIFFERS'}, "
                      f"executed {info['executed_iters']} vs {si['executed_iters']}, refs {info['references']} vs {si['references']}",
                      flush=True)
                ok = ok and same and same_rgb and same_n
    flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local))
    dist.broadcast(flag, 0)
    grp.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI-RANK OK" if ok else "MULTI-RANK MISMATCH", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
