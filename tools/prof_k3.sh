#!/bin/bash
# ncu --set full capture of the dominant kernel (k3_fast) on a FULL 1024-iteration level. Run under gpurun:
#   tools/prof_k3.sh <tag> [bench.py options...]        e.g.  tools/prof_k3.sh r01k --workload cfg3 --scale 2
# Pass 1 lists every k3_fast launch of the command with its duration (cheap), pass 2 captures the longest one of
# that list (a full, unsplit level) and its successor with the full metric set. Summaries for profiles/:
#   python tools/ncu_summary.py metrics gpurun_out/<tag>_k3fast.ncu-rep > profiles/<tag>_k3_fast_metrics.txt
tag=$1; shift
cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline $*"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k3_fast --csv --log-file gpurun_out/${tag}_k3fast_list.csv $cmd > gpurun_out/${tag}_prof1.log 2>&1
skip=$(python - "$tag" <<'PY'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/{sys.argv[1]}_k3fast_list.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = []
for r in rows[hi + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", "")); u = r[ui]
        d.append(v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0))
best = max(range(len(d) - 1), key=lambda i: d[i])
print(best)
PY
)
echo "capturing k3_fast launches $skip, $((skip+1))"
ncu --set full --clock-control none --import-source on -k regex:k3_fast -s $skip -c 2 -o gpurun_out/${tag}_k3fast -f $cmd > gpurun_out/${tag}_prof2.log 2>&1
tail -n 2 gpurun_out/${tag}_prof2.log | cut -c1-200
