#!/bin/bash
# Scratch: build libaidax.so variants with extra -D flags into tools/bin/<name>/ (timing / what-if experiments).
#   bash tools/variant.sh name1 "-DAIDAX_POLY_EX2=4" name2 "-DAIDAX_WHATIF=1" ...     (no GPU needed)
# Only the H=16 and H=40 variant objects are rebuilt; the rest is linked from build/.  On the GPU: python tools/variant_run.py
cd "$(dirname "$0")/../aidadsp-lv2_b200" || exit 1
NVF="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -diag-suppress 177"
HS=${VARIANT_HS:-"16 40"}
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( d=../tools/bin/$name; mkdir -p $d; objs=""; skip=""
    for h in $HS; do
      nvcc $NVF -DAIDAX_VARIANT_H=$h $flags -c csrc/variants.cu -o $d/variants_h$h.o || exit 1
      objs="$objs $d/variants_h$h.o"; skip="$skip -e h$h.o"
    done
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libaidax.so build/engine.o build/probes.o $(ls build/variants_h*.o | grep -v $skip) $objs -cudart shared &&
    cp lib/rt-neural-generic.so lib/libaidax_host.so $d/ &&   # the plugin finds libaidax.so next to itself ($ORIGIN): AIDAX_LIB_DIR=$d runs everything on the variant
    rm -f $d/variants_h*.o && echo "built $d ($flags)" ) &
done
wait
