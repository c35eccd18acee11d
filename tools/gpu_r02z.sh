#!/bin/bash
# Round 2, GPU call Z: hand-over chunk size of the plugin path with the one-tile chunk kernels (fixed 184320 / 196608 instances).
TAG=${1:-r02z}
OUT=$PWD/gpurun_out
mkdir -p $OUT
for n in 184320 196608; do
for c in 16384 12288 8192 6144; do
  AIDAX_BATCH_CHUNK=$c timeout 300 python bench.py --steps 5 --warmup 3 --e2e-steps 30 --e2e-streams $n --e2e-instances $n --no-literal --no-cpu-baseline > $OUT/${TAG}_c$c.json 2>/dev/null
  python - <<PY
import json
e = json.load(open("$OUT/${TAG}_c$c.json"))["e2e"]
print("n $n chunk $c: mean %.0f us p95 %.0f us max %.0f phase %s launches %s" % (e["us_per_buffer"], e["us_p95"], e["us_max"], e["phase_us"], e["gpu_launches"]))
PY
done
done | tee $OUT/${TAG}_chunks.txt
