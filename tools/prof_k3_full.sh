#!/bin/bash
# ncu --set full capture of k3_fast on a FULL (quiet) 1024-iteration level: like tools/prof_k3.sh, but instead of the longest
# launch of the list (an escape level: loud) it takes the launch of median duration among those above half the maximum.
tag=$1; shift
cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras $*"
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
big = sorted([i for i in range(len(d) - 1) if d[i] > 0.5 * max(d)], key=lambda i: d[i])
print(big[len(big) // 2])
PY
)
echo "capturing k3_fast launch $skip (full level)"
ncu --set full --clock-control none --import-source on -k regex:k3_fast -s $skip -c 1 -o gpurun_out/${tag}_k3fast_full -f $cmd > gpurun_out/${tag}_prof2.log 2>&1
tail -n 2 gpurun_out/${tag}_prof2.log | cut -c1-200
