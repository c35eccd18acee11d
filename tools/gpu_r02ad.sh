#!/bin/bash
# Round 2, GPU call AD: C4's 8-GPU shard (8192 streams over 64 models) -- FP32 shapes vs tensor-core shape from 512 streams.
OUT=$PWD/gpurun_out
for n in 8192 4096; do
for mode in default tc512; do
  if [ $mode = tc512 ]; then export AIDAX_TC_MIN_STREAMS=512; else unset AIDAX_TC_MIN_STREAMS; fi
  timeout 400 python bench.py --workload c4 --streams $n --steps 20 --warmup 5 --e2e-steps 15 --no-cpu-baseline --no-literal > $OUT/r02ad_c4_${n}_$mode.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("$OUT/r02ad_c4_${n}_$mode.json")); e = d["e2e"]
print("c4 streams $n $mode: device %.3f ms | e2e %.0f us p95 %.0f" % (d["ms_per_step"], e["us_per_buffer"], e["us_p95"]))
PY
done
done | tee $OUT/r02ad_planner_small.txt
