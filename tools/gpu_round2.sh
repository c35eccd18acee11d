#!/bin/bash
# Round 2: the round's evidence on HEAD (gpurun -- 'bash tools/gpu_round2.sh TAG') (warp-specialised kernel as shipped): GPU suite, smoke, both bench arms,
# C3/C4/C5, golden-vector margins, ncu launch list + full captures.
TAG=${1:-r02x}
OUT=$PWD/gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > $OUT/${TAG}_gpu_box.txt 2>&1
lscpu | head -20 >> $OUT/${TAG}_gpu_box.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q -rP > $OUT/${TAG}_pytest_gpu_full.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/${TAG}_pytest_gpu_full.log
grep -E "golden vectors, tile_mode|passed|failed" $OUT/${TAG}_pytest_gpu_full.log > $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.txt
echo "== bench (reference arm)"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2>/dev/null; echo "ref exit $?"
echo "== bench (ours)"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
for w in c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> /dev/null; echo "$w exit $?"
done
TAG=$TAG python - <<'PY'
import json, glob, os
for name in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ["TAG"])):
    try:
        d = json.load(open(name)); e = d["e2e"]
        print(name, "value %.4e" % d["value"], "dev ms", d["ms_per_step"], "| e2e %.4e" % e["value"], "n", e.get("streams_per_gpu"), "us", e.get("us_per_buffer"), "p95", e.get("us_p95"), "meets", e.get("meets_deadline"), d.get("roofline", {}).get("bound"), d.get("roofline", {}).get("frac"))
    except Exception as ex:
        print(name, "unreadable", ex)
PY
echo "== golden-vector margins per shape"
timeout 300 python tools/golden_err.py > $OUT/${TAG}_golden_margins.txt 2>&1; cat $OUT/${TAG}_golden_margins.txt
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 5 --e2e-streams 163840 --no-literal --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l $OUT/${TAG}_launches.csv
echo "== ncu --set full: chain_kernel_tc3<LSTM,16> at the bench's 378880 streams"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc3 -s 3 -c 1 -o $OUT/${TAG}_tc3_16_full -f \
    python tools/prof_one.py 378880 16 3 2 > $OUT/${TAG}_ncu_tc3_16.log 2>&1; echo "exit $?"
ncu -i $OUT/${TAG}_tc3_16_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc3_16_details.txt 2>&1
ncu -i $OUT/${TAG}_tc3_16_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc3_16_raw.csv 2>&1
ncu -i $OUT/${TAG}_tc3_16_full.ncu-rep --page source --csv > $OUT/${TAG}_ncu_tc3_16_src.csv 2>&1
echo "== ncu --set full: chain_kernel_tc3<LSTM,40,split 2> at C5's 262144 streams"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc3 -s 2 -c 1 -o $OUT/${TAG}_tc3_40_full -f \
    python tools/prof_one.py 262144 40 2 2 > $OUT/${TAG}_ncu_tc3_40.log 2>&1; echo "exit $?"
ncu -i $OUT/${TAG}_tc3_40_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc3_40_details.txt 2>&1
ncu -i $OUT/${TAG}_tc3_40_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc3_40_raw.csv 2>&1
rm -f $OUT/${TAG}_tc3_40_full.ncu-rep $OUT/${TAG}_tc3_16_full.ncu-rep
ls -la $OUT | grep ${TAG} | head -30
