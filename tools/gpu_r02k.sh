#!/bin/bash
# Round 2, GPU call K: generation 3 as the engine default: full GPU suite, sub-block length per hidden size, bench.
TAG=${1:-r02k}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
one() { # H N extra-env...
  local H=$1 N=$2; shift 2
  env "$@" timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run($N,$H,tile_mode=3,iters=10)" 2>&1 | tail -1
}
{
for H in 8 12 16 20; do for ts in 0 8 16; do echo "H=$H ts_max=$ts: $(one $H 378880 AIDAX_TS_MAX=$ts)"; done; done
for H in 32 40; do for ts in 0 8 16; do echo "H=$H ts_max=$ts: $(one $H 262144 AIDAX_TS_MAX=$ts)"; done; done
echo "GRU16 in3 gen1: $(env AIDAX_TC_GEN=1 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,16,cell=1,input_size=3,tile_mode=3,iters=10)" 2>&1 | tail -1)"
echo "GRU16 in3 gen3: $(timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,16,cell=1,input_size=3,tile_mode=3,iters=10)" 2>&1 | tail -1)"
} | tee $OUT/${TAG}_ts.txt
echo "== bench (ours)"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_bench.json")); e = d["e2e"]
print("value %.4e" % d["value"], "dev ms", d["ms_per_step"], "| e2e %.4e" % e["value"], "n", e.get("streams_per_gpu"), "us", e.get("us_per_buffer"), "p95", e.get("us_p95"), "meets", e.get("meets_deadline"), "roofline", d["roofline"]["bound"], d["roofline"]["frac"])
PY
