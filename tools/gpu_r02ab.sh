#!/bin/bash
# Round 2, GPU call AB: launch planner for mixed-model periods that fill the GPU: C4's per-GPU shards at 2 / 4 / 8 GPUs emulated on
# one GPU (--streams 32768 / 16384 / 8192 over the same 64 models), old thresholds (AIDAX_TC_MIN_STREAMS=4096) against the new rule.
TAG=${1:-r02ab}
OUT=$PWD/gpurun_out
mkdir -p $OUT
for n in 65536 32768 16384 8192; do
  for mode in old new; do
    if [ $mode = old ]; then export AIDAX_TC_MIN_STREAMS=4096; else unset AIDAX_TC_MIN_STREAMS; fi
    timeout 400 python bench.py --workload c4 --streams $n --steps 20 --warmup 5 --e2e-steps 15 --no-cpu-baseline --no-literal > $OUT/${TAG}_c4_${n}_$mode.json 2>/dev/null
    python - <<PY
import json
d = json.load(open("$OUT/${TAG}_c4_${n}_$mode.json")); e = d["e2e"]
print("c4 streams $n $mode: device %.3f ms | e2e %.0f us p95 %.0f launches/step %s" % (d["ms_per_step"], e["us_per_buffer"], e["us_p95"], d["roofline"].get("launches_per_step")))
PY
  done
done | tee $OUT/${TAG}_planner.txt
unset AIDAX_TC_MIN_STREAMS
echo "== tests that exercise auto mode"
timeout 900 python -m pytest tests/test_workloads_gpu.py tests/test_plugin_gpu.py -m gpu -x -q 2>&1 | tail -3
