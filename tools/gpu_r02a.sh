#!/bin/bash
# Round 2, GPU call A: parity suite, fault-hunt arms, pipe peaks, bench with the bulk plugin path, C4/C5 lines, ncu of the H=40 kernel.
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
{ nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current,pci.bus_id --format=csv; nvidia-smi topo -m; lscpu | head -30; numactl -H 2>/dev/null;
  for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; nproc; free -g | head -2; } > $OUT/${TAG}_box.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
echo "== fault hunt"
timeout 900 bash tools/race_probe.sh run 65536 50 > $OUT/${TAG}_race.txt 2>&1; cat $OUT/${TAG}_race.txt
echo "== pipe peaks"
timeout 120 python -c "
import sys; sys.path.insert(0,'.')
from aidadsp_lv2_b200 import engine
import json
print(json.dumps(engine.measure_pipe_peaks(0, 1400)))
print(json.dumps(engine.measure_pipe_peaks(0, 1400)))" 2>&1 | tee $OUT/${TAG}_pipe_peaks.txt
echo "== bench (ours)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; head -c 6000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== bench e2e without the bulk entry (A/B)"
AIDAX_HOST_NO_BULK=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-literal --no-cpu-baseline --e2e-streams 196608 > $OUT/${TAG}_bench_nobulk.json 2> $OUT/${TAG}_bench_nobulk.err; echo "exit $?"
echo "== bench (reference arm)"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2>/dev/null; echo "ref exit $?"; head -c 900 $OUT/${TAG}_bench_reference.json
for w in c3 c4 c5; do
  echo "== bench $w"
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; echo "exit $?"; head -c 1200 $OUT/${TAG}_bench_$w.json; tail -2 $OUT/${TAG}_bench_$w.err
done
echo "== c4 e2e without the family pipelining (A/B)"
AIDAX_PIPELINE_FAMILIES=0 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_c4_nopipe.json 2> $OUT/${TAG}_bench_c4_nopipe.err; echo "exit $?"
echo "== ncu --set full of the LSTM-40 tensor-core kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc -s 2 -c 1 -o $OUT/${TAG}_tc40_full -f \
    python tools/prof_one.py 262144 40 2 2 > $OUT/${TAG}_ncu_tc40.log 2>&1
echo "ncu exit $?"
ncu -i $OUT/${TAG}_tc40_full.ncu-rep --page details > $OUT/${TAG}_ncu_tc40_details.txt 2>&1
ncu -i $OUT/${TAG}_tc40_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_tc40_raw.csv 2>&1
echo "== per-shape tables"
timeout 300 python tools/quick_bench_tc.py > $OUT/${TAG}_tc_shapes.txt 2>&1; tail -12 $OUT/${TAG}_tc_shapes.txt
