#!/bin/bash
# Round 2, GPU call AE: planner test, bench default with the bisecting e2e search.
OUT=$PWD/gpurun_out
timeout 600 python -m pytest tests/test_planner_gpu.py tests/test_tc_generations_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02ae_bench.json 2>/dev/null; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02ae_bench.json")); e = d["e2e"]
print("value %.4e" % d["value"], "| e2e n", e["streams_per_gpu"], "us", e["us_per_buffer"], "p95", e["us_p95"], "pcie", e["pcie"]["duplex_gbs_per_direction"], e["selection"][-220:])
PY
