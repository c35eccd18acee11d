#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, the ncu launch list of the bench command and one
# `--set full` capture of the dominant kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | head -20 >> $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
echo "== bench (ours)"
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; head -c 3000 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench (reference arm)"
timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2>/dev/null; echo "ref exit $?"; head -c 600 $OUT/${TAG}_bench_reference.json
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 5 --no-literal --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l $OUT/${TAG}_launches.csv
echo "== ncu --set full of the dominant kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc -s 3 -c 1 -o $OUT/${TAG}_tc16_full -f \
    python tools/prof_one.py 378880 16 3 5 > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i $OUT/${TAG}_tc16_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc16_details.txt 2>&1
ncu -i $OUT/${TAG}_tc16_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc16_raw.csv 2>&1
python - "$OUT/${TAG}_ncu_tc16_raw.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
    if name in hdr:
        print(name, [r[hdr.index(name)] for r in rows[1:3]])
PY
echo "== per-shape tables"
timeout 200 python tools/quick_bench_tc.py > $OUT/${TAG}_tc_shapes.txt 2>&1; tail -12 $OUT/${TAG}_tc_shapes.txt
timeout 200 python tools/quick_bench_lat.py > $OUT/${TAG}_latency_shapes.txt 2>&1; tail -20 $OUT/${TAG}_latency_shapes.txt
echo "== ncu --set full of the warp-per-stream latency kernel (1024 streams)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_w -s 3 -c 1 -o $OUT/${TAG}_w16_full -f \
    python tools/prof_one.py 1024 16 3 2 1 > $OUT/${TAG}_ncu_w_full.log 2>&1
ncu -i $OUT/${TAG}_w16_full.ncu-rep --page details > $OUT/${TAG}_ncu_w16_details.txt 2>&1
