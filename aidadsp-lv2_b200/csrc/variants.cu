// variants.cu -- explicit instantiations of chain_kernel for one hidden size (AIDAX_VARIANT_H).
// Tile shapes (L lanes per stream group, S streams per group) per hidden size: the first entry is
// the low-latency shape (one stream per group, many lanes), the last the throughput shape
// (register tile of S streams x 4U gate rows per lane, >= 10 FFMA per 128-bit shared load).
#include "kernels.cuh"
#include "kernels_tc.cuh"
#include "kernels_tc2.cuh"
#include "kernels_tc3.cuh"
#include "kernels_w.cuh"
#include "variants.h"

#ifndef AIDAX_VARIANT_H
#error "compile with -DAIDAX_VARIANT_H=<hidden size>"
#endif

namespace aidax {

template <int H, int L, int S>
static int wsf(int ts, int np) { return Tile<0, H, L, S>::warp_smem_floats(ts, np); }

#define V(CELL, Hh, Ll, Ss, NWw)                                                                         \
    { CELL, Hh, Ll, Ss, NWw, (32 / Ll) * Ss,                                                             \
      { (const void*)&chain_kernel<CELL, Hh, Ll, Ss, NWw, 0>, (const void*)&chain_kernel<CELL, Hh, Ll, Ss, NWw, 1> }, \
      &wsf<Hh, Ll, Ss>, 0, CELL == 2 ? 0 : (Hh + 3) * Hh * 4 + Hh * 4 + Hh + 4, 2 }
#define BOTH(Hh, Ll, Ss, NWw) V(0, Hh, Ll, Ss, NWw), V(1, Hh, Ll, Ss, NWw)

// tensor-core variants: NT tiles of 128 streams per CTA, RES CTAs per SM
template <int H>
static int wsf_tc(int ts, int np) { return TcShape<H>::warp_floats(ts, np); }
#define VT(CELL, Hh, NTt, RES)                                                                           \
    { CELL, Hh, 0, 0, NTt * 4, 32,                                                                       \
      { (const void*)&chain_kernel_tc<CELL, Hh, NTt, RES, 0>, (const void*)&chain_kernel_tc<CELL, Hh, NTt, RES, 1> }, \
      &wsf_tc<Hh>, 1, (TcShape<Hh>::fixed_floats(NTt) + 3) & ~3, RES }
#define BOTH_TC(Hh, NTt, RES) VT(0, Hh, NTt, RES), VT(1, Hh, NTt, RES)

// ping-pong tensor-core variants (kernels_tc2.cuh): two tiles of 128 streams per CTA, SPLIT threads per stream in the
// recurrence, A operand in shared (ATM 0) or tensor (ATM 1) memory, RES CTAs per SM.  L carries SPLIT, S carries ATM.
// ATM bit 0: A operand in tensor memory; bit 1: one tile per thread group instead of ping-pong (PP = false).
template <int H, int SPLIT, bool ATM, bool PP>
static int wsf_tc2(int ts, int np) { return Tc2Shape<H, SPLIT, ATM, PP>::warp_floats(ts, np); }
#define TC2_SHAPE(Hh, SPL, ATM) Tc2Shape<Hh, SPL, ((ATM) & 1) != 0, ((ATM) & 2) == 0>
#define VT2(CELL, Hh, SPL, ATM, RES)                                                                     \
    { CELL, Hh, SPL, ATM, TC2_SHAPE(Hh, SPL, ATM)::NW, 256 / TC2_SHAPE(Hh, SPL, ATM)::NW,                \
      { (const void*)&chain_kernel_tc2<CELL, Hh, SPL, ((ATM) & 1) != 0, ((ATM) & 2) == 0, RES, 0>,       \
        (const void*)&chain_kernel_tc2<CELL, Hh, SPL, ((ATM) & 1) != 0, ((ATM) & 2) == 0, RES, 1> },     \
      &wsf_tc2<Hh, SPL, ((ATM) & 1) != 0, ((ATM) & 2) == 0>, 3, (TC2_SHAPE(Hh, SPL, ATM)::fixed_floats() + 3) & ~3, RES, 256 }
#define BOTH_TC2(Hh, SPL, ATM, RES) VT2(0, Hh, SPL, ATM, RES), VT2(1, Hh, SPL, ATM, RES)

// warp-specialised tensor-core variants (kernels_tc3.cuh): two tiles of 128 streams per CTA, SPLIT recurrence warps per lane
// quadrant + one filter warpgroup; RR / FR = registers per thread of the two roles after setmaxnreg.  L carries SPLIT,
// S carries 5 (bit 0: A operand in tensor memory, bit 2: warp-specialised).
#ifndef AIDAX_TC3_NFW
#define AIDAX_TC3_NFW 8 // filter warps per CTA (4: two streams per lane, 8: one)
#endif
template <int H, int SPLIT, int NFWv, int NT>
static int wsf_tc3(int ts, int np) { return Tc3Shape<H, SPLIT, NFWv, NT>::warp_floats(ts, np); }
#define TC3_SHAPE(Hh, SPL, NFWv, NT) Tc3Shape<Hh, SPL, NFWv, NT>
#define VT3(CELL, Hh, SPL, NFWv, RR, FR, RES, NT)                                                        \
    { CELL, Hh, SPL, 5, TC3_SHAPE(Hh, SPL, NFWv, NT)::NW, TC3_SHAPE(Hh, SPL, NFWv, NT)::CS / TC3_SHAPE(Hh, SPL, NFWv, NT)::NW, \
      { (const void*)&chain_kernel_tc3<CELL, Hh, SPL, NFWv, RR, FR, RES, 0, NT>,                         \
        (const void*)&chain_kernel_tc3<CELL, Hh, SPL, NFWv, RR, FR, RES, 1, NT> },                       \
      &wsf_tc3<Hh, SPL, NFWv, NT>, 3, (TC3_SHAPE(Hh, SPL, NFWv, NT)::fixed_floats() + 3) & ~3, RES, TC3_SHAPE(Hh, SPL, NFWv, NT)::CS }
#define BOTH_TC3(Hh, SPL, RR, FR, RES) VT3(0, Hh, SPL, AIDAX_TC3_NFW, RR, FR, RES, 2), VT3(1, Hh, SPL, AIDAX_TC3_NFW, RR, FR, RES, 2)
// one tile per CTA (the shape for launches that underfill the GPU): 4 SPLIT recurrence warps + 4 filter warps
#define BOTH_TC3_1T(Hh, SPL, RR, FR, RES) VT3(0, Hh, SPL, 4, RR, FR, RES, 1), VT3(1, Hh, SPL, 4, RR, FR, RES, 1)
// three tiles per CTA, one CTA per SM, one filter warp per warp-tile (12): 768 threads launch with 80 registers
#define BOTH_TC3_3T(Hh, RR, FR) VT3(0, Hh, 1, 12, RR, FR, 1, 3), VT3(1, Hh, 1, 12, RR, FR, 1, 3)

#ifndef AIDAX_TC_NT32
#define AIDAX_TC_NT32 3
#endif
#ifndef AIDAX_TC_NT40
#define AIDAX_TC_NT40 2
#endif
#ifndef AIDAX_TC_RES16
#define AIDAX_TC_RES16 4
#endif
// warp-per-stream latency variants (kernels_w.cuh): one stream per warp, hidden size <= 16
static int wsf_w(int ts, int np) { return WShape::warp_floats(ts, np); }
#define VW(CELL, Hh, NWw)                                                                                \
    { CELL, Hh, 32, 1, NWw, 1,                                                                           \
      { (const void*)&chain_kernel_w<CELL, Hh, NWw, 0>, (const void*)&chain_kernel_w<CELL, Hh, NWw, 1> }, \
      &wsf_w, 2, 0, 2 }
#define BOTH_W(Hh, NWw) VW(0, Hh, NWw), VW(1, Hh, NWw)

// Two-tile shapes (kernels_tc2.cuh) compiled per hidden size: (H, SPLIT, ATM, CTAs per SM); ATM bit 0 = A operand in tensor
// memory, bit 1 = one tile per thread group (no ping-pong).  Only shapes that BEAT kernels_tc.cuh on B200 are in the
// default build (profiles/r02c_tc2_time.txt: LSTM-40 3.18 vs 3.31 ms per 262144-stream period with the hidden units of a
// row split over two warps and A in tensor memory; every ping-pong shape and every other hidden size measured slower); the
// engine takes a listed shape in preference to the first-generation kernel.  tools/variant.sh overrides a list with
// -DAIDAX_TC2_LIST_<H>="..." to measure others (tools/tc2_check.py checks and times them).
#ifndef AIDAX_TC2_LIST_8
#define AIDAX_TC2_LIST_8
#endif
#ifndef AIDAX_TC2_LIST_12
#define AIDAX_TC2_LIST_12
#endif
#ifndef AIDAX_TC2_LIST_16
#define AIDAX_TC2_LIST_16
#endif
#ifndef AIDAX_TC2_LIST_20
#define AIDAX_TC2_LIST_20
#endif
#ifndef AIDAX_TC2_LIST_24
#define AIDAX_TC2_LIST_24
#endif
#ifndef AIDAX_TC2_LIST_32
#define AIDAX_TC2_LIST_32
#endif
#ifndef AIDAX_TC2_LIST_40
#define AIDAX_TC2_LIST_40 BOTH_TC2(40, 2, 3, 1),
#endif

// Warp-specialised shapes per hidden size (see kernels_tc3.cuh): (H, SPLIT, R registers, F registers, CTAs per SM)
#ifndef AIDAX_TC3_RREG
#define AIDAX_TC3_RREG (AIDAX_TC3_NFW == 8 ? 96 : 104)
#endif
#ifndef AIDAX_TC3_FREG
#define AIDAX_TC3_FREG 32
#endif
// SPLIT 2 (one CTA of 640 threads per SM: 96 registers per thread at launch)
#ifndef AIDAX_TC3_RREG2
#define AIDAX_TC3_RREG2 104
#endif
#ifndef AIDAX_TC3_FREG2
#define AIDAX_TC3_FREG2 (AIDAX_TC3_NFW == 8 ? 32 : 64)
#endif
#ifndef AIDAX_TC3_LIST_8
#define AIDAX_TC3_LIST_8 BOTH_TC3(8, 1, AIDAX_TC3_RREG, AIDAX_TC3_FREG, 2), BOTH_TC3_1T(8, 1, 96, 32, 4),
#endif
#ifndef AIDAX_TC3_LIST_12
#define AIDAX_TC3_LIST_12 BOTH_TC3(12, 1, AIDAX_TC3_RREG, AIDAX_TC3_FREG, 2), BOTH_TC3_1T(12, 1, 96, 32, 4),
#endif
#ifndef AIDAX_TC3_LIST_16
#define AIDAX_TC3_LIST_16 BOTH_TC3(16, 1, AIDAX_TC3_RREG, AIDAX_TC3_FREG, 2), BOTH_TC3_1T(16, 1, 96, 32, 4),
#endif
#ifndef AIDAX_TC3_LIST_20
#define AIDAX_TC3_LIST_20 BOTH_TC3(20, 1, AIDAX_TC3_RREG, AIDAX_TC3_FREG, 2), BOTH_TC3_1T(20, 1, 96, 32, 4),
#endif
#ifndef AIDAX_TC3_H24_THREE_TILES
#define AIDAX_TC3_H24_THREE_TILES 1 // H = 24: three tiles per CTA (tensor memory holds 3 x 160 columns), one CTA per SM, 12 + 12 warps at 128 / 32 registers: 1446 vs 1716 us (generation 1) per 227328-stream period, bit-identical
#endif
#ifndef AIDAX_TC3_LIST_24
#if AIDAX_TC3_H24_THREE_TILES
#define AIDAX_TC3_LIST_24 BOTH_TC3_3T(24, 128, 32), BOTH_TC3_1T(24, 1, 128, 32, 2),
#else
#define AIDAX_TC3_LIST_24 // (two tiles with SPLIT 2 measured slower than generation 1: 2040 vs 1961 us per 262144-stream period)
#endif
#endif
#ifndef AIDAX_TC3_LIST_32
#define AIDAX_TC3_LIST_32 BOTH_TC3(32, 2, AIDAX_TC3_RREG2, AIDAX_TC3_FREG2, 1), BOTH_TC3_1T(32, 2, 104, 32, 2),
#endif
#ifndef AIDAX_TC3_LIST_40
#define AIDAX_TC3_LIST_40 BOTH_TC3(40, 2, AIDAX_TC3_RREG2, AIDAX_TC3_FREG2, 1), BOTH_TC3_1T(40, 2, 104, 32, 2),
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

const KernelVariant CAT(kVariantsH, AIDAX_VARIANT_H)[] = {
#if AIDAX_VARIANT_H == 0
    V(2, 8, 8, 8, 4),
#elif AIDAX_VARIANT_H == 8
    BOTH_W(8, 4), BOTH(8, 8, 1, 4), BOTH(8, 8, 4, 4), BOTH(8, 8, 8, 4), BOTH_TC(8, 1, 4), AIDAX_TC2_LIST_8 AIDAX_TC3_LIST_8
#elif AIDAX_VARIANT_H == 12
    BOTH_W(12, 4), BOTH(12, 4, 1, 4), BOTH(12, 4, 2, 4), BOTH(12, 4, 4, 4), BOTH_TC(12, 1, 4), AIDAX_TC2_LIST_12 AIDAX_TC3_LIST_12
#elif AIDAX_VARIANT_H == 16
    BOTH_W(16, 4), BOTH(16, 16, 1, 4), BOTH(16, 16, 4, 4), BOTH(16, 8, 8, 4), BOTH_TC(16, 1, AIDAX_TC_RES16), AIDAX_TC2_LIST_16 AIDAX_TC3_LIST_16
#elif AIDAX_VARIANT_H == 20
    BOTH(20, 4, 1, 4), BOTH(20, 4, 2, 4), BOTH(20, 4, 4, 4), BOTH_TC(20, 1, 3), AIDAX_TC2_LIST_20 AIDAX_TC3_LIST_20
#elif AIDAX_VARIANT_H == 24
    BOTH(24, 8, 1, 4), BOTH(24, 8, 4, 4), BOTH(24, 8, 8, 4), BOTH_TC(24, 1, 3), AIDAX_TC2_LIST_24 AIDAX_TC3_LIST_24
#elif AIDAX_VARIANT_H == 32
    BOTH(32, 32, 1, 4), BOTH(32, 16, 4, 4), BOTH(32, 8, 4, 4), BOTH_TC(32, AIDAX_TC_NT32, 1), AIDAX_TC2_LIST_32 AIDAX_TC3_LIST_32
#elif AIDAX_VARIANT_H == 40
    BOTH(40, 8, 1, 4), BOTH(40, 8, 2, 4), BOTH(40, 8, 4, 4), BOTH_TC(40, AIDAX_TC_NT40, 1), AIDAX_TC2_LIST_40 AIDAX_TC3_LIST_40
#elif AIDAX_VARIANT_H == 64
    BOTH(64, 32, 1, 4), BOTH(64, 32, 4, 4), BOTH(64, 16, 4, 4),
#elif AIDAX_VARIANT_H == 80
    BOTH(80, 16, 1, 4), BOTH(80, 16, 2, 4), BOTH(80, 16, 4, 4),
#else
#error "unsupported hidden size"
#endif
};
const int CAT(kNumVariantsH, AIDAX_VARIANT_H) = (int)(sizeof(CAT(kVariantsH, AIDAX_VARIANT_H)) / sizeof(KernelVariant));

} // namespace aidax
