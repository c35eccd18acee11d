#!/bin/bash
# Short GPU call while iterating on the kernels: engine parity tests, then the tensor-core throughput probe.
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh TAG'
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest engine+workloads (gpu)"
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_workloads_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log
echo "== tc probe"
timeout 300 python tools/quick_bench_tc.py 2>&1 | tee $OUT/${TAG}_quick_tc.log | tail -14
echo "== golden err"
timeout 120 python tools/golden_err.py 2>&1 | tail -8
