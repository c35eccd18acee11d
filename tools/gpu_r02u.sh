#!/bin/bash
# Round 2, GPU call U: full GPU suite with every hidden size <= 40 on the warp-specialised kernel (H = 24: three tiles), C4 bench.
TAG=${1:-r02u}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
echo "== sanitizers on the three-tile shape"
cat > /tmp/san24.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np
from tc2_check import render
y, l = render(384 + 70, 24, 1, 3, {"AIDAX_TC_GEN": "3"}, 3, 0, [128, 100, 28, 128])
print("san H=24: %d launches finite %s" % (l, bool(np.isfinite(y).all())))
PY
for tool in memcheck racecheck synccheck; do timeout 300 compute-sanitizer --tool $tool --print-limit 4 python /tmp/san24.py 2>&1 | grep -E "^san|SUMMARY" ; done | tee $OUT/${TAG}_san_h24.txt
for t in 1 2; do
timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c4_$t.json 2> /dev/null; echo "c4 exit $?"
done
python - <<'PY'
import json, glob
for name in sorted(glob.glob("gpurun_out/r02u_bench*.json")):
    d = json.load(open(name)); e = d["e2e"]
    print(name, "value %.4e" % d["value"], "dev ms", d["ms_per_step"], "| e2e %.4e" % e["value"], "n", e.get("streams_per_gpu"), "us", e.get("us_per_buffer"), "p95", e.get("us_p95"), "meets", e.get("meets_deadline"))
PY
