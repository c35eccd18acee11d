#!/bin/bash
# Round 2, GPU call L: evidence on the warp-specialised default (ts 16): ncu full captures (H 16, H 40), launch list of the bench
# command, bench lines (C2 capacity, C3, C4, C5), the new self-test and generation tests.
TAG=${1:-r02l}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== new tests"
timeout 600 python -m pytest tests/test_tc_generations_gpu.py tests/test_parity_holes_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== bench (ours)"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
for w in c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> /dev/null; echo "$w exit $?"
done
python - <<'PY'
import json, glob
for name in sorted(glob.glob("gpurun_out/r02l_bench*.json")):
    try:
        d = json.load(open(name)); e = d["e2e"]
        print(name, "value %.4e" % d["value"], "dev ms", d["ms_per_step"], "| e2e %.4e" % e["value"], "n", e.get("streams_per_gpu"), "us", e.get("us_per_buffer"), "p95", e.get("us_p95"), "meets", e.get("meets_deadline"), d["roofline"]["bound"], d["roofline"]["frac"])
    except Exception as ex:
        print(name, "unreadable", ex)
PY
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 5 --e2e-streams 163840 --no-literal --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l $OUT/${TAG}_launches.csv
echo "== ncu --set full: chain_kernel_tc3<LSTM,16> at the bench's 378880 streams"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc3 -s 3 -c 1 -o $OUT/${TAG}_tc3_16_full -f \
    python tools/prof_one.py 378880 16 3 2 > $OUT/${TAG}_ncu_tc3_16.log 2>&1; echo "exit $?"
ncu -i $OUT/${TAG}_tc3_16_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc3_16_details.txt 2>&1
ncu -i $OUT/${TAG}_tc3_16_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc3_16_raw.csv 2>&1
echo "== ncu --set full: chain_kernel_tc3<LSTM,40,split 2> at C5's 262144 streams"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc3 -s 2 -c 1 -o $OUT/${TAG}_tc3_40_full -f \
    python tools/prof_one.py 262144 40 2 2 > $OUT/${TAG}_ncu_tc3_40.log 2>&1; echo "exit $?"
ncu -i $OUT/${TAG}_tc3_40_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc3_40_details.txt 2>&1
ncu -i $OUT/${TAG}_tc3_40_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc3_40_raw.csv 2>&1
rm -f $OUT/${TAG}_tc3_40_full.ncu-rep
ls -la $OUT | grep ${TAG} | head -30
