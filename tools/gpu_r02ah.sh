#!/bin/bash
# Round 2, GPU call AH: what a C4 period launches with the loaded-period planner rule (ncu launch list), and an ncu capture of the
# one-tile warp-specialised kernel at a hand-over chunk's size (16384 streams).
OUT=$PWD/gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/r02ah_c4_launches.csv \
    python bench.py --workload c4 --steps 3 --warmup 3 --e2e-steps 3 --no-literal --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02ah_c4_launches.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    a = agg.setdefault(r[k], [0, 0.0]); a[0] += 1; a[1] += float(r[v].replace(",", ""))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%3d launches, mean %8.1f us  %s" % (c, t / c / 1000.0, n[:100]))
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc3 -s 3 -c 1 -o $OUT/r02ah_tc3_16_one_tile -f \
    python tools/prof_one.py 16384 16 3 2 > /dev/null 2>&1; echo "ncu exit $?"
ncu -i $OUT/r02ah_tc3_16_one_tile.ncu-rep --page details > $OUT/r02ah_ncu_tc3_16_one_tile_details.txt 2>&1
ncu -i $OUT/r02ah_tc3_16_one_tile.ncu-rep --page raw --csv > $OUT/r02ah_ncu_tc3_16_one_tile_raw.csv 2>&1
rm -f $OUT/r02ah_tc3_16_one_tile.ncu-rep
