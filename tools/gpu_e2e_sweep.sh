#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_plugin_gpu.py -x -q 2>&1 | tail -3
for cfg in "16384 4 16" "16384 4 12" "16384 4 8" "8192 4 16" "32768 4 16"; do
  set -- $cfg
  AIDAX_BATCH_CHUNK=$1 AIDAX_KERNEL_STREAMS=$2 timeout 300 python tools/e2e_one.py 378880 30 $3 2>&1 | tail -1
done
AIDAX_TRACE=1 AIDAX_BATCH_CHUNK=16384 timeout 300 python tools/e2e_one.py 378880 10 16 2>&1 | tail -27
} 2>&1 | tee gpurun_out/e2e_sweep.log
