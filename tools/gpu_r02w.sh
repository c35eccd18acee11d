#!/bin/bash
# Round 2, GPU call W: one-tile warp-specialised CTAs for launches that underfill the GPU.
TAG=${1:-r02w}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== generation tests"
timeout 900 python -m pytest tests/test_tc_generations_gpu.py -m gpu -x -q 2>&1 | tail -3
one() { # H N cell insz extra-env
  local H=$1 N=$2 C=$3 I=$4; shift 4
  env "$@" timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run($N,$H,cell=$C,input_size=$I,tile_mode=3,iters=20)" 2>&1 | tail -1 | sed 's/.*N=/N=/'
}
{
for N in 4096 8192 16384 32768 65536; do
  echo "H=16 tiles1: $(one 16 $N 0 1 AIDAX_TC3_TILES=1)"
  echo "H=16 tiles2: $(one 16 $N 0 1 AIDAX_TC3_TILES=2)"
done
for N in 4608 16384 32768; do
  echo "H=40 tiles1: $(one 40 $N 0 1 AIDAX_TC3_TILES=1)"
  echo "H=40 tiles2: $(one 40 $N 0 1 AIDAX_TC3_TILES=2)"
  echo "H=24 tiles1: $(one 24 $N 0 1 AIDAX_TC3_TILES=1)"
  echo "H=24 tiles3: $(one 24 $N 0 1 AIDAX_TC3_TILES=2)"
done
echo "GRU16 in3 8192 tiles1: $(one 16 8192 1 3 AIDAX_TC3_TILES=1)"
echo "GRU16 in3 8192 tiles2: $(one 16 8192 1 3 AIDAX_TC3_TILES=2)"
} | tee $OUT/${TAG}_tiles.txt
for w in c3 c4; do
  for t in 1 2; do
  AIDAX_TC3_TILES=$t timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_${w}_tiles$t.json 2> /dev/null; echo "$w tiles$t exit $?"
  done
done
python - <<'PY'
import json, glob
for name in sorted(glob.glob("gpurun_out/r02w_bench*.json")):
    d = json.load(open(name)); e = d["e2e"]
    print(name, "value %.4e" % d["value"], "dev ms", d["ms_per_step"], "| e2e %.4e" % e["value"], "n", e.get("streams_per_gpu"), "us", e.get("us_per_buffer"), "p95", e.get("us_p95"), "meets", e.get("meets_deadline"))
PY
