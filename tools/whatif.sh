#!/bin/bash
# Scratch: build libaidax.so variants whose H=16 kernels have parts removed (AIDAX_WHATIF bits, kernels_tc.cuh) into
# tools/bin/whatif_<bits>/ -- timing experiments only, the results of those libraries are wrong by construction.
#   bash tools/whatif.sh 1 2 3 8      (here, no GPU needed);  on the GPU: python tools/whatif_run.py
cd "$(dirname "$0")/../aidadsp-lv2_b200" || exit 1
NVF="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -diag-suppress 177"
for w in "$@"; do
  ( d=../tools/bin/whatif_$w; mkdir -p $d
    nvcc $NVF -DAIDAX_VARIANT_H=16 -DAIDAX_WHATIF=$w -c csrc/variants.cu -o $d/variants_h16.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libaidax.so build/engine.o $(ls build/variants_h*.o | grep -v h16.o) $d/variants_h16.o -cudart shared &&
    rm $d/variants_h16.o && echo "built $d" ) &
done
wait
