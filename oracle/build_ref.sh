#!/usr/bin/env bash
# TEST INFRASTRUCTURE. Builds the UNMODIFIED reference from the sources where they lie under
# $AIDAX_REFERENCE (default /root/reference) into oracle/_ref/ (git-ignored, travels with gpurun).
# Mirrors rt-neural-generic/CMakeLists.txt:8-46 and modules/RTNeural/cmake/SIMDExtensions.cmake:3
# (xsimd backend, SSE2 default, alignment 16) without running the reference's own build system.
# LV2 headers are absent from the image: the plugin is compiled against a forwarding shim that
# includes include/aidax_lv2_abi.h (this repo's layout-exact restatement of the LV2 ABI).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/.." && pwd)"
REF="${AIDAX_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/rt-neural-generic/src" ]; then
    echo "build_ref: reference not present at $REF; keeping prebuilt oracle/_ref as is" >&2
    exit 0
fi
mkdir -p "$OUT/lv2shim/lv2"
for h in atom/forge.h atom/util.h atom/atom.h core/lv2.h log/log.h log/logger.h midi/midi.h patch/patch.h \
         state/state.h urid/urid.h worker/worker.h parameters/parameters.h; do
    mkdir -p "$OUT/lv2shim/lv2/$(dirname $h)"
    echo '#include "aidax_lv2_abi.h"' > "$OUT/lv2shim/lv2/$h"
done
RTN="$REF/modules/RTNeural"
COMMON=(-std=c++17 -O3 -DNDEBUG -fPIC -shared -w
        -DAIDADSP_COMMERCIAL=0 -DAIDADSP_MODEL_LOADER=1
        -I"$OUT/lv2shim" -I"$ROOT/include" -I"$REF/rt-neural-generic/src" -I"$REF/common"
        -I"$RTN" -I"$RTN/modules/xsimd/include" -I"$RTN/modules/json")
XSIMD=(-DRTNEURAL_USE_XSIMD=1 -DRTNEURAL_DEFAULT_ALIGNMENT=16)
build() { # out, extra flags..., sources are after "--"
    local out="$1"; shift
    if [ -f "$out" ] && [ "$out" -nt "$REF/rt-neural-generic/src/rt-neural-generic.cpp" ] \
       && [ "$out" -nt "$HERE/ref_model_harness.cpp" ] && [ "$out" -nt "$ROOT/include/aidax_lv2_abi.h" ] \
       && [ "$out" -nt "$HERE/build_ref.sh" ]; then return 0; fi
    echo "build_ref: $out" >&2
    g++ "$@" -o "$out"
}
# tier A: the full plugin, reference default flavour (xsimd, SSE2)
build "$OUT/rt-neural-generic.so" "${COMMON[@]}" "${XSIMD[@]}" \
      "$REF/rt-neural-generic/src/rt-neural-generic.cpp" "$REF/common/Biquad.cpp" &
# tier A': AVX2+FMA flavour (what -DRTNEURAL_USE_AVX=ON gives) -- noise-floor reference only
build "$OUT/rt-neural-generic-avx2.so" "${COMMON[@]}" -DRTNEURAL_USE_XSIMD=1 -DRTNEURAL_DEFAULT_ALIGNMENT=32 -mavx2 -mfma \
      "$REF/rt-neural-generic/src/rt-neural-generic.cpp" "$REF/common/Biquad.cpp" &
# tier B: model-only harness, xsimd SSE2
build "$OUT/libref_model.so" "${COMMON[@]}" "${XSIMD[@]}" "$HERE/ref_model_harness.cpp" &
# tier C: model-only harness, RTNeural STL backend (scalar spec; ground truth for in_size 2/3)
build "$OUT/libref_model_stl.so" "${COMMON[@]}" -DRTNEURAL_USE_STL=1 "$HERE/ref_model_harness.cpp" &
wait
ls -la "$OUT"/*.so >&2
