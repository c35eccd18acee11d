#!/bin/bash
# Round 2, GPU call P: compute-sanitizer on the warp-specialised kernel, then the full GPU suite on HEAD.
TAG=${1:-r02p}
OUT=$PWD/gpurun_out
mkdir -p $OUT
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 500 compute-sanitizer --tool $tool --print-limit 8 python tools/tc3_check.py san > $OUT/${TAG}_san_$tool.txt 2>&1; echo "exit $?"
  grep -E "^san|SUMMARY|Error|error" $OUT/${TAG}_san_$tool.txt | head -12
done
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
