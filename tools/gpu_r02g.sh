#!/bin/bash
# Round 2, GPU call G: full GPU suite + smoke on HEAD, both bench arms, C4 (per-family copy-out on the kernel streams), ncu evidence.
TAG=${1:-r02g}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
echo "== bench (ours)"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
echo "== bench (reference arm)"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2>/dev/null; echo "ref exit $?"
for t in 1 2; do
  echo "== bench c4 (try $t)"
  timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c4_$t.json 2> $OUT/${TAG}_bench_c4_$t.err; echo "exit $?"
done
echo "== bench c3"
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c3.json 2> /dev/null; echo "exit $?"
python - <<'PY'
import json, glob
for name in sorted(glob.glob("gpurun_out/r02g_bench*.json")):
    try:
        d = json.load(open(name)); e = d["e2e"]
        print(name, "value %.4e" % d["value"], "dev ms", d["ms_per_step"], "| e2e %.4e" % e["value"], "n", e.get("streams_per_gpu"), "us", e.get("us_per_buffer"), "p95", e.get("us_p95"), e.get("phase_us"), "meets", e.get("meets_deadline"))
    except Exception as ex:
        print(name, "unreadable", ex)
PY
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 5 --e2e-streams 163840 --no-literal --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l $OUT/${TAG}_launches.csv
echo "== ncu --set full: chain_kernel_tc<LSTM,16> at the bench's 378880 streams"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc -s 3 -c 1 -o $OUT/${TAG}_tc16_full -f \
    python tools/prof_one.py 378880 16 3 2 > $OUT/${TAG}_ncu_tc16.log 2>&1; echo "exit $?"
ncu -i $OUT/${TAG}_tc16_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc16_details.txt 2>&1
ncu -i $OUT/${TAG}_tc16_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc16_raw.csv 2>&1
echo "== ncu --set full: chain_kernel_tc2<LSTM,40,split 2,TMEM A> at C5's 262144 streams"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc2 -s 2 -c 1 -o $OUT/${TAG}_tc40_full -f \
    python tools/prof_one.py 262144 40 2 2 > $OUT/${TAG}_ncu_tc40.log 2>&1; echo "exit $?"
ncu -i $OUT/${TAG}_tc40_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc40_details.txt 2>&1
ncu -i $OUT/${TAG}_tc40_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc40_raw.csv 2>&1
rm -f $OUT/${TAG}_tc16_full.ncu-rep $OUT/${TAG}_tc40_full.ncu-rep.tmp
ls -la $OUT | grep ${TAG} | head -30
