"""Measure the FP64-pipe instruction-issue peaks (DFMA/DADD/DMUL) on the box: the roofline
denominator for K1/K3 (MEASURED_PEAKS.json has no FP64 entry). Writes gpurun_out/fp64_peak.json."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import newman_b200

d = newman_b200.Device(0)
info = d.info()
res = {"device": info}
for kind, name in ((0, "dfma"), (1, "dadd"), (2, "dmul"), (3, "k3mix")):
    d.fp64_peak(kind, 1 << 12)
    best = 0
    for it in (1 << 15, 1 << 17):
        ips, ms = d.fp64_peak(kind, it)
        best = max(best, ips)
        print(name, it, "inst/s %.4e  ms %.3f  lanes/clk/SM@maxclk %.2f" % (ips, ms, ips / (info["sm_count"] * info["sm_clock_khz"] * 1e3)))
    res[name + "_inst_per_s"] = best
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    res["nvidia_smi_after"] = q
except Exception as e:
    res["nvidia_smi_after"] = str(e)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(res))
