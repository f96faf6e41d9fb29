"""Print selected keys of the bench JSON line found on stdin (skips non-JSON lines such as NCCL banners)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print(tag, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3),
              "e2e", round(d["e2e"]["value"], 1), "e2e_ms", round(d["e2e"]["ms_per_step"], 2), "launches", d["gpu_launches"],
              "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("samples"), "eff_sum", round(d.get("effective_giter_s", 0) * d["ms_per_step"], 3),
              d.get("frames_per_s_device", ""))
