#!/bin/bash
# Round 2, GPU call Q: sanitizers on the warp-specialised kernel after the rendezvous / all-lane arrival change; check; timing.
TAG=${1:-r02q}
OUT=$PWD/gpurun_out
mkdir -p $OUT
echo "== check"
TC3_HS="8 16 20 32 40" timeout 900 python tools/tc3_check.py check 2>&1 | tail -12
echo "== timing"
timeout 300 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from quick_bench import run
run(378880,16,tile_mode=3,iters=10); run(378880,16,tile_mode=3,iters=10); run(262144,40,tile_mode=3,iters=10)" 2>&1 | tail -3
for tool in synccheck racecheck memcheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 500 compute-sanitizer --tool $tool --print-limit 8 python tools/tc3_check.py san > $OUT/${TAG}_san_$tool.txt 2>&1; echo "exit $?"
  grep -E "^san|SUMMARY" $OUT/${TAG}_san_$tool.txt | head -12
done
