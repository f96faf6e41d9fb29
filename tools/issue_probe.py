"""Issue-port probe (run on the GPU box): DFMA rate with 0/8/16/24 integer-pipe operations interleaved per
8 DFMA, at 64 and at 16 warps per SM (nm_fp64_peak kinds 4-11), next to the plain DFMA/DADD/DMUL rates."""
import json
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import newman_b200  # noqa: E402

dev = newman_b200.Device(0)
out = {}
for name, kind in (("dfma", 0), ("dadd", 1), ("dmul", 2), ("dfma+0int@64w", 4), ("dfma+8int@64w", 5),
                   ("dfma+16int@64w", 6), ("dfma+24int@64w", 7), ("dfma+0int@16w", 8), ("dfma+8int@16w", 9),
                   ("dfma+16int@16w", 10), ("dfma+24int@16w", 11), ("dfma 3 distinct operands", 12),
                   ("dfma 2 distinct operands", 13), ("dfma 1 distinct operand", 14), ("dadd 2 distinct operands", 15)):
    ips, ms = dev.fp64_peak(kind, 1 << 15)
    out[name] = {"fp64_ginst_s": ips * 1e-9, "ms": ms}
    print(f"{name:26s} {ips * 1e-9:10.1f} G FP64 inst/s  ({ms:.3f} ms)", file=sys.stderr)
print(json.dumps(out))
dev.close()
