#!/bin/bash
# Round 2, GPU call AC: full GPU suite and the C3 / C4 / C5 lines after the planner change.
TAG=${1:-r02ac}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
for w in c4 c3 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> /dev/null; echo "$w exit $?"
done
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> /dev/null; echo "c2 exit $?"
python - <<'PY'
import json, glob
for name in sorted(glob.glob("gpurun_out/r02ac_bench*.json")):
    d = json.load(open(name)); e = d["e2e"]
    print(name, "value %.4e" % d["value"], "dev ms", d["ms_per_step"], "| e2e %.4e" % e["value"], "n", e.get("streams_per_gpu"), "us", e.get("us_per_buffer"), "p95", e.get("us_p95"), "meets", e.get("meets_deadline"), d["roofline"]["bound"], d["roofline"]["frac"])
PY
