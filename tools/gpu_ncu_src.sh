#!/bin/bash
# One gpurun call: `ncu --set full --import-source on` of the tensor-core kernel for H=16 and H=40 (the .ncu-rep comes
# back in gpurun_out/ and is read here with `ncu -i ... --page source --csv`).
OUT=gpurun_out; mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc -s 3 -c 1 -o $OUT/src_tc16 -f \
    python tools/prof_one.py 378880 16 3 2 > $OUT/src_tc16.log 2>&1; echo "ncu h16 exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:chain_kernel_tc -s 3 -c 1 -o $OUT/src_tc40 -f \
    python tools/prof_one.py 262144 40 3 2 > $OUT/src_tc40.log 2>&1; echo "ncu h40 exit $?"
ls -la $OUT/*.ncu-rep
