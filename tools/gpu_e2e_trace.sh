#!/bin/bash
mkdir -p gpurun_out
{
AIDAX_TRACE=1 AIDAX_BATCH_CHUNK=16384 timeout 300 python tools/e2e_one.py 378880 10 16 2>&1 | tail -80
python tools/prof_one.py 16384 16 3 20
python tools/prof_one.py 16384 16 0 20
python tools/prof_one.py 32768 16 3 20
python tools/prof_one.py 75776 16 3 20
} 2>&1 | tee gpurun_out/e2e_trace.log
