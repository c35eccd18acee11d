#!/bin/bash
# Round 2, GPU call Y: smoke on HEAD and the kernels it launches (what the driver's GPUTEST record will list).
TAG=${1:-r02y}
OUT=$PWD/gpurun_out
mkdir -p $OUT
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_smoke_launches.csv python -c 'import __graft_entry__ as g; g.smoke()' > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02y_smoke_launches.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name")
names = {}
for r in rows[1:]:
    names[r[k]] = names.get(r[k], 0) + 1
for n, c in names.items():
    print(c, n[:110])
PY
