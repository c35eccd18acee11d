// kernels_tc.cuh -- tensor-core (tcgen05) variant of the fused chain kernel, for large groups of streams that
// share one model.
//
// The per-time-step gate pre-activations of 128 streams are ONE dense product
//      D[128 streams x 4H] = A[128 x K] . B[K x 4H],   A row = [h(t-1) | x(t) p1(t) p2(t) 1],  B = [U ; W ; bias]
// (RTN/lstm/lstm_xsimd.h:193-213 and gru_xsimd.h:219-239 compute the same thing one stream at a time).  It runs on
// the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 128, N = 4H, operands in shared memory, the fp32
// accumulator tile in TMEM.  fp32 accuracy is kept with the 3xTF32 operand split (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi;
// measured error ~2e-6, single TF32/BF16 products were rejected by measurement).
//
// Work decomposition: a CTA owns NT tiles of 128 streams of one model; thread r of a tile owns stream r (= TMEM
// lane r): it stages that stream's audio, runs the filter/gain stages (shared with kernels.cuh), and per time
// step reads its 4H accumulator columns with tcgen05.ld, applies the gate nonlinearities and the cell update
// thread-locally (c stays in registers for the whole buffer), reduces the Dense layer thread-locally, and writes
// h(t) (split hi/lo) straight into the A operand tile for the next step.  One elected thread per tile issues
// the MMAs and commits them to an mbarrier; the tiles of a CTA (and the CTAs of an SM) run decoupled so one
// tile's epilogue hides another's MMA + barrier latency.
//
// Shared memory operand layout (no swizzle, K-major "core matrices" of 8 rows x 16 bytes):
//      element (row, k) of A  ->  float index (k/4)*(128*4) + row*4 + k%4       (a thread writes its row with STS.128,
//      element (n,   k) of B  ->  float index (k/4)*(N*4)   + n*4   + k%4        32 threads -> 512 contiguous bytes)
#pragma once

#include "kernels.cuh"
#include "tc_prims.cuh"

// Timing experiments only (tools/whatif.sh builds scratch libraries with -DAIDAX_WHATIF=<bits>; results are wrong by
// construction): 1 = no MMA issue / completion wait, 2 = MUFU replaced by FMUL, 4 = no exponent clamps,
// 8 = no filter/gain stages.
#ifndef AIDAX_WHATIF
#define AIDAX_WHATIF 0
#endif
// Build-time choices of the tensor-core kernel, each settled by measurement on B200 (tools/variant.sh builds scratch
// libraries with other values, tools/variant_run.py / tools/acc_run.py time and check them; DESIGN.md 3.2 has the table).
#ifndef AIDAX_TC_GROUP
#define AIDAX_TC_GROUP 2 // 16-column accumulator chunks (4 hidden units each) evaluated per straight-line block
#endif
#ifndef AIDAX_SCALED_GATES
#define AIDAX_SCALED_GATES 1
#endif
#ifndef AIDAX_PREFETCH_L2
#define AIDAX_PREFETCH_L2 1
#endif
#ifndef AIDAX_B_RN
#define AIDAX_B_RN 1 // weights: hi part of the tf32 split rounded to nearest (free: B is built once per CTA)
#endif
#ifndef AIDAX_A_RN
#define AIDAX_A_RN 0 // recurrent vector: same (one more integer add per element and step)
#endif
#ifndef AIDAX_EARLY_HALF
#define AIDAX_EARLY_HALF 0
#endif
#ifndef AIDAX_DESC64
#define AIDAX_DESC64 0 // 1: carry full 64-bit operand descriptors instead of their low words
#endif
// Which exponentials of the gate epilogue are evaluated on the FMA pipe instead of the XU pipe (bit per site, see
// TcCell::ope2): measured slower for H = 16 (+2 % with the tanh(c') site alone, +30 % with all), so none.
// 1: all filter/gain stages on one side of the network as ONE six-stage software pipeline (kernels.cuh: pass_fused)
#ifndef AIDAX_FUSED_PASSES
#define AIDAX_FUSED_PASSES 0
#endif
#ifndef AIDAX_POLY_EX2
#define AIDAX_POLY_EX2 0
#endif
#ifndef AIDAX_POLY_PAIRS
#define AIDAX_POLY_PAIRS 0
#endif

namespace aidax {

#if AIDAX_DESC64
typedef uint64_t tc_desc_t;
#else
typedef uint32_t tc_desc_t;
#endif

template <int H>
struct TcShape {
    static constexpr int N = 4 * H;              // accumulator columns per tile: 4 per hidden unit (see layout.h lanes)
    static constexpr int KX = H;                 // first input row: x, p1, p2, then the constant 1 that carries the bias
    static constexpr int KP = (H + 4 + 7) / 8 * 8; // K padded to the tf32 MMA depth (8)
    static constexpr int KSTEPS = KP / 8;
    // The accumulator is produced in NSPLIT column halves with their own completion barrier (all MMAs of half 0 are
    // issued and committed first): the epilogue of the first hidden units starts while the tensor core still works
    // on the second half.  Needs halves that are legal MMA widths (multiple of 16) and more than one epilogue group.
    static constexpr int NSPLIT = (AIDAX_EARLY_HALF && (N / 2) % 16 == 0 && H >= 16) ? 2 : 1;
    static constexpr int NMMA = N / NSPLIT;
    // first epilogue group (pair of 16-column chunks) that needs columns of the second half
    static constexpr int QWAIT = NSPLIT == 1 ? 0 : ((NMMA / 16) & ~1);
    static constexpr int A_FLOATS = 128 * KP;    // one of (hi, lo)
    static constexpr int B_FLOATS = N * KP;
    static constexpr int XP = 33;                // pitch of the per-warp audio tile (odd: conflict free both ways)
    static_assert(H % 4 == 0 && NMMA % 16 == 0 && NMMA <= 256, "unsupported hidden size for the tcgen05 path");

    __host__ __device__ static constexpr int tmem_cols(int nt)
    {
        int c = 32;
        while (c < nt * N) c *= 2;
        return c;
    }
    // floats of shared memory that do not depend on the sub-block length
    __host__ __device__ static constexpr int fixed_floats(int nt)
    {
        return 2 * B_FLOATS + nt * 2 * A_FLOATS + (H + 4) /* dense w, b */ + 4 * nt /* mbarriers: 2 per tile */ + 4 /* tmem base + pad */;
    }
    __host__ __device__ static constexpr int warp_floats(int ts, int n_param_tiles) { return 32 + ts * XP * (1 + n_param_tiles); }
};

// everything a thread needs to drive its tile's recurrence
struct TcTile {
    float* a_hi;         // this thread's row in the A operand: a_hi + chunk*512 is a float4 slot
    float* a_lo;
    tc_desc_t da_hi;     // (low words of) the shared-memory matrix descriptors of the tile's A (hi, lo) and the CTA's B (hi, lo)
    tc_desc_t da_lo;     // at K step 0 (issue_mmas adds the constant high word)
    tc_desc_t db_hi;
    tc_desc_t db_lo;
    uint32_t bar_addr;   // the tile's mbarrier
    uint32_t phase;      // its current parity
    uint32_t tmem_d;     // TMEM address of the tile's accumulator (lane 0)
    uint32_t tmem_ld;    // same, with this warp's lane quadrant
    uint32_t bar_id;     // named barrier of the tile's 128 threads
    const float* dw;     // dense weights [H] + bias
    bool issuer;         // warp-uniform: this warp issues the tile's MMAs (one elected lane)
};

template <int CELL, int H, int MATH>
struct TcCell {
    using S = TcShape<H>;

    __device__ __forceinline__ static void store_chunk(const TcTile& tt, int chunk, float2 v01, float2 v23)
    {
        // 3xTF32 split (tc::split_tf32) of four values, the subtraction packed
#if AIDAX_A_RN
        constexpr uint32_t RB = 0x1000u; // hi rounded to nearest (see tc::split_tf32_rn)
#else
        constexpr uint32_t RB = 0u;
#endif
        const float2 h01 = make_float2(__uint_as_float((__float_as_uint(v01.x) + RB) & 0xFFFFE000u), __uint_as_float((__float_as_uint(v01.y) + RB) & 0xFFFFE000u));
        const float2 h23 = make_float2(__uint_as_float((__float_as_uint(v23.x) + RB) & 0xFFFFE000u), __uint_as_float((__float_as_uint(v23.y) + RB) & 0xFFFFE000u));
        const float2 d01 = tc::sub2(v01, h01), d23 = tc::sub2(v23, h23);
        *reinterpret_cast<float4*>(tt.a_hi + chunk * 512) = make_float4(h01.x, h01.y, h23.x, h23.y);
        *reinterpret_cast<float4*>(tt.a_lo + chunk * 512) =
            make_float4(__uint_as_float(__float_as_uint(d01.x) + 0x1000u), __uint_as_float(__float_as_uint(d01.y) + 0x1000u),
                        __uint_as_float(__float_as_uint(d23.x) + 0x1000u), __uint_as_float(__float_as_uint(d23.y) + 0x1000u));
    }
    __device__ __forceinline__ static void store_chunk(const TcTile& tt, int chunk, float v0, float v1, float v2, float v3)
    {
        store_chunk(tt, chunk, make_float2(v0, v1), make_float2(v2, v3));
    }

    // Issued by ONE elected thread of the tile's first warp.  Only the LOW words of the four base descriptors (start
    // address and leading byte offset) are kept, as opaque loop-invariant values; the high word (stride byte offset +
    // descriptor version) is a compile-time constant and a K step adds a constant to the low word's address field
    // (shared memory is < 256 KB, so the 14-bit field cannot carry).  This keeps the per-step uniform-datapath preamble
    // in front of the first UTCHMMA -- which sits on every stream's critical path -- to a handful of instructions.
    __device__ __forceinline__ static void issue_mmas(const TcTile& tt)
    {
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, S::NMMA);
        constexpr uint32_t LBO_A = 128 * 16, LBO_B = S::N * 16;
        constexpr uint64_t HI = ((uint64_t)((128u >> 4) & 0x3FFFu) << 32) | (1ull << 46); // SBO = 128 bytes, version 1
#if AIDAX_DESC64
        constexpr uint64_t HIW = 0;
#else
        constexpr uint64_t HIW = HI;
#endif
        const tc_desc_t a_hi = tt.da_hi, a_lo = tt.da_lo, b_hi = tt.db_hi, b_lo = tt.db_lo;
#pragma unroll
        for (int nh = 0; nh < S::NSPLIT; ++nh) {
            const uint32_t d = tt.tmem_d + nh * S::NMMA;
#pragma unroll
            for (int ks = 0; ks < S::KSTEPS; ++ks) {
                const uint32_t ao = (ks * 2 * LBO_A) >> 4, bo = (ks * 2 * LBO_B) >> 4;
                const uint32_t bn = bo + ((nh * S::NMMA * 16) >> 4);
                // smallest terms first
                tc::mma_tf32(d, HIW | (a_lo + ao), HIW | (b_hi + bn), idesc, ks > 0);
                tc::mma_tf32(d, HIW | (a_hi + ao), HIW | (b_lo + bn), idesc, 1);
                tc::mma_tf32(d, HIW | (a_hi + ao), HIW | (b_hi + bn), idesc, 1);
            }
            tc::mma_commit(tt.bar_addr + 8 * nh); // fires when every MMA issued so far has completed
        }
    }

    // accumulator column of (hidden unit u, gate lane g).  Within a 16-column chunk of 4 units the order is
    // [g0:u0 u1 | g1:u0 u1 | g2:u0 u1 | g3:u0 u1 | g0:u2 u3 | ... ] so registers (2j, 2j+1) of one tcgen05.ld hold the SAME
    // gate of two neighbouring units: an aligned register pair is directly one packed-fp32 (f32x2) operand.
    __host__ __device__ static constexpr int column_of(int u, int g) { return 16 * (u / 4) + 8 * ((u % 4) / 2) + 2 * g + (u % 2); }
    __host__ __device__ static constexpr int unit_of_column(int n) { return 4 * (n / 16) + 2 * ((n % 16) / 8) + (n % 2); }
    __host__ __device__ static constexpr int gate_of_column(int n) { return (n % 8) / 2; }

    __device__ __forceinline__ static float2 pair(const uint32_t (&raw)[16], int j)
    {
        return make_float2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
    }
#if AIDAX_WHATIF & 2
    __device__ __forceinline__ static float2 rcp2(float2 a) { return tc::mul2(a, tc::splat2(0.999f)); }
#else
    __device__ __forceinline__ static float2 rcp2(float2 a) { return make_float2(rcp_approx(a.x), rcp_approx(a.y)); }
#endif
    // 1/a, 1/b, 1/c, 1/d from ONE MUFU.RCP per lane of the pair (a..d in [1, 2^24+1]: no overflow, ~3 ulp): the kernel
    // is bound by the XU (MUFU) pipe and by issue slots, so reciprocals are batched four at a time and the extra
    // multiplies run as packed FMUL2 (two hidden units per instruction) on the otherwise idle FMA pipe.
    __device__ __forceinline__ static void inv4(float2 a, float2 b, float2 c, float2 d, float2& ia, float2& ib, float2& ic, float2& id)
    {
        const float2 ab = tc::mul2(a, b), cd = tc::mul2(c, d);
        const float2 r = rcp2(tc::mul2(ab, cd));
        const float2 rab = tc::mul2(r, cd), rcd = tc::mul2(r, ab);
        ia = tc::mul2(rab, b); ib = tc::mul2(rab, a); ic = tc::mul2(rcd, d); id = tc::mul2(rcd, c);
    }
    __device__ __forceinline__ static void inv2(float2 a, float2 b, float2& ia, float2& ib)
    {
        const float2 r = rcp2(tc::mul2(a, b));
        ia = tc::mul2(r, b); ib = tc::mul2(r, a);
    }
    // 1 + 2^e, exponent clamped so that four of them can be multiplied without overflow
    // (sigmoid(-16.6) = 6e-8, 1 - tanh(8.3) = 1.2e-7: the clamp is below fp32 resolution of the results)
    __device__ __forceinline__ static float2 one_plus_exp2(float2 e)
    {
#if AIDAX_WHATIF & 2
        return tc::add2(tc::mul2(make_float2(fminf(e.x, 24.0f), fminf(e.y, 24.0f)), tc::splat2(0.001f)), tc::splat2(1.0f));
#elif AIDAX_WHATIF & 4
        return tc::add2(make_float2(ex2_approx(e.x), ex2_approx(e.y)), tc::splat2(1.0f));
#else
        return tc::add2(make_float2(ex2_approx(fminf(e.x, 24.0f)), ex2_approx(fminf(e.y, 24.0f))), tc::splat2(1.0f));
#endif
    }
    // 2^-24 * (1 + 2^e), saturated at 1 (i.e. 2^e clamped at 2^24 - 1; e = +inf from the MUFU saturates as well)
    __device__ __forceinline__ static float2 scaled_one_plus_exp2(float2 e)
    {
        constexpr float s = 5.9604644775390625e-08f;
#if AIDAX_WHATIF & 2
        return make_float2(__saturatef(fmaf(e.x, s, s)), __saturatef(fmaf(e.y, s, s)));
#else
        return make_float2(__saturatef(fmaf(ex2_approx(e.x), s, s)), __saturatef(fmaf(ex2_approx(e.y), s, s)));
#endif
    }
    // Same value with 2^e evaluated on the FMA pipe instead of the XU pipe: e = n + f, n = round(e), |f| <= 0.5,
    // 2^f by a degree-6 minimax polynomial (max relative error 0.66 ulp in fp32 Horner form, better than MUFU.EX2's
    // 2 ulp), 2^n by adding n to the exponent field.  All packed (two hidden units per FADD2/FFMA2).  The kernel is
    // bound by the XU pipe (8 cycles per warp MUFU) while the FMA pipe idles, so a FRACTION of the exponentials --
    // chosen by measurement, AIDAX_POLY_EX2 -- is moved over: +11 issue slots, -16 XU cycles per pair.
    __device__ __forceinline__ static float2 one_plus_exp2_fma(float2 e)
    {
        const float2 ec = make_float2(fminf(fmaxf(e.x, -125.0f), 24.0f), fminf(fmaxf(e.y, -125.0f), 24.0f));
        const float2 magic = tc::splat2(12582912.0f); // 1.5 * 2^23: the sum's low mantissa bits are round(e)
        const float2 t = tc::add2(ec, magic);
        const float2 f = tc::sub2(ec, tc::sub2(t, magic));
        float2 p = tc::splat2(1.534581097e-04f);
        p = tc::fma2(p, f, tc::splat2(1.339993090e-03f));
        p = tc::fma2(p, f, tc::splat2(9.618489072e-03f));
        p = tc::fma2(p, f, tc::splat2(5.550328642e-02f));
        p = tc::fma2(p, f, tc::splat2(2.402264625e-01f));
        p = tc::fma2(p, f, tc::splat2(6.931471825e-01f));
        p = tc::fma2(p, f, tc::splat2(1.0f));
        // p in [0.707, 1.415], n in [-125, 24]: the exponent field stays normal
        const float2 r = make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23)),
                                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23)));
        return tc::add2(r, tc::splat2(1.0f));
    }
    // 2^-24 * (1 + 2^e) saturated at 1, like scaled_one_plus_exp2, with 2^e on the FMA pipe: the polynomial's coefficients carry
    // the 2^-24 (exact), e is clamped to [-100, 24] so that the exponent arithmetic stays in the normal range
    // (2^-100 vanishes next to 1; at 24 the sum saturates).
    __device__ __forceinline__ static float2 scaled_one_plus_exp2_fma(float2 e)
    {
        constexpr float s = 5.9604644775390625e-08f;
        const float2 ec = make_float2(fminf(fmaxf(e.x, -100.0f), 24.0f), fminf(fmaxf(e.y, -100.0f), 24.0f));
        const float2 magic = tc::splat2(12582912.0f);
        const float2 t = tc::add2(ec, magic);
        const float2 f = tc::sub2(ec, tc::sub2(t, magic));
        float2 p = tc::splat2(1.534581097e-04f * s);
        p = tc::fma2(p, f, tc::splat2(1.339993090e-03f * s));
        p = tc::fma2(p, f, tc::splat2(9.618489072e-03f * s));
        p = tc::fma2(p, f, tc::splat2(5.550328642e-02f * s));
        p = tc::fma2(p, f, tc::splat2(2.402264625e-01f * s));
        p = tc::fma2(p, f, tc::splat2(6.931471825e-01f * s));
        p = tc::fma2(p, f, tc::splat2(s));
        const float rx = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
        const float ry = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
        return make_float2(__saturatef(rx + s), __saturatef(ry + s));
    }
    // site: 0 = input/forget/output (LSTM) or update/reset (GRU) gates, 1 = the LSTM candidate gate g,
    // 2 = tanh of the new cell state (LSTM) / of the candidate (GRU)
    template <int SITE>
    __device__ __forceinline__ static float2 ope2(float2 e)
    {
#if AIDAX_WHATIF & 2
        return one_plus_exp2(e);
#else
        if constexpr (((AIDAX_POLY_EX2) >> SITE) & 1) return one_plus_exp2_fma(e);
        else return one_plus_exp2(e);
#endif
    }
    // In fast-math mode the B operand's columns are pre-scaled so that the accumulator IS that exponent:
    // sigmoid(x) = 1/(1 + 2^(-x log2 e)), tanh(x) = 1 - 2/(1 + 2^(2x log2 e)).  Gate lanes per unit: LSTM i,f,g,o;
    // GRU z, r, Uh.h+bh1, Wh.x+bh0 (the last two are summed as r*a2 + a3 before the tanh, so both carry its scale).
    __host__ __device__ static constexpr float column_scale(int g)
    {
        constexpr float L2E = 1.4426950408889634f;
        return CELL == 0 ? (g == 2 ? 2.f * L2E : -L2E) : (g < 2 ? -L2E : 2.f * L2E);
    }

    // state: cst (LSTM: c, GRU: h) in registers for the whole render call, h(t-1) split hi/lo in the A operand tile
    __device__ __forceinline__ static void load_state(const TcTile& tt, float (&cst)[H], const StreamArrays& sa, int slot, bool load)
    {
        const size_t cap = (size_t)sa.cap;
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
            float hv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = 4 * q + i;
                float h = 0.f, c = 0.f;
                if (load) {
                    h = sa.rnn[(size_t)u * cap + slot];
                    if (CELL == 0) c = sa.rnn[(size_t)(kMaxHidden + u) * cap + slot];
                }
                cst[u] = CELL == 0 ? c : h;
                hv[i] = h;
            }
            store_chunk(tt, q, hv[0], hv[1], hv[2], hv[3]);
        }
    }
    // LSTM: h is read back from the A operand tile as hi + (lo as the tensor core sees it): exactly the value the
    // next step would have used, so a render call split into two periods gives the same samples as one long call.
    __device__ __forceinline__ static void store_state(const TcTile& tt, const float (&cst)[H], const StreamArrays& sa, int slot)
    {
        const size_t cap = (size_t)sa.cap;
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
            float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
            if (CELL == 0) {
                hi = *reinterpret_cast<const float4*>(tt.a_hi + q * 512);
                lo = *reinterpret_cast<const float4*>(tt.a_lo + q * 512);
            }
            const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = 4 * q + i;
                if (CELL == 0) {
                    sa.rnn[(size_t)(kMaxHidden + u) * cap + slot] = cst[u];
                    sa.rnn[(size_t)u * cap + slot] = hv[i] + __uint_as_float(__float_as_uint(lv[i]) & 0xFFFFE000u);
                } else {
                    sa.rnn[(size_t)u * cap + slot] = cst[u];
                }
            }
        }
    }

    // Gate nonlinearities + cell update of the 4 (TWO: 8) hidden units whose accumulator columns are in raw0 (and raw1),
    // two units per packed instruction.  hn[2c], hn[2c+1] = new h of the units (4(q+c), +1) and (4(q+c)+2, +3).
    // (NC: length of the caller's state array -- all H units here, the units of one thread part in kernels_tc2.cuh)
    template <bool TWO, int NC>
    __device__ __forceinline__ static void eval_group(const uint32_t (&raw0)[16], const uint32_t (&raw1)[16], int q, float (&cst)[NC],
                                                      float2 (&hn)[4])
    {
        constexpr float L2E = 1.4426950408889634f;
        constexpr int NP = TWO ? 4 : 2; // unit pairs in the group
        const float2 one = tc::splat2(1.0f), m2 = tc::splat2(-2.0f);
        if (MATH == 0 && CELL == 0 && AIDAX_SCALED_GATES) {
            // LSTM as below, restated to spend fewer issue slots (the kernel runs at a fixed ~0.46 IPC per scheduler, so
            // time follows the instruction count).  With a = 1+2^-i, b = 1+2^-f, g = 1+2^(2g), d = 1+2^-o:
            //   i*tanh(g) = (g-2)/(a g),  f = d/(b d),  o = b/(b d)     -- ONE reciprocal of a*g*b*d, 10 packed ops;
            // a, g, d and the 1+2^(2c') terms are kept scaled by s = 2^-24 (exact): s*(1+2^e) = sat(fma(2^e, s, s)) -- the
            // saturation IS the overflow clamp the batched reciprocal needs (one FFMA.SAT instead of FMNMX + FADD), and
            // the leftover powers of s are folded into constants of FMAs that exist anyway.
            constexpr float s = 5.9604644775390625e-08f; // 2^-24
            const float2 s2 = tc::splat2(s), ms2 = tc::splat2(-2.0f * s * s);
            float2 ogx[NP], ec[NP], t[NP];
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) {
                const uint32_t(&raw)[16] = pr < 2 ? raw0 : raw1;
                const int j0 = 4 * (pr & 1), u = 4 * q + 2 * pr;
                const float2 A = scaled_one_plus_exp2(pair(raw, j0)), B = one_plus_exp2(pair(raw, j0 + 1));
                const float2 G = scaled_one_plus_exp2(pair(raw, j0 + 2)), D = scaled_one_plus_exp2(pair(raw, j0 + 3));
                const float2 ag = tc::mul2(A, G), bd = tc::mul2(B, D);
                const float2 r = rcp2(tc::mul2(ag, bd));
                const float2 rag = tc::mul2(r, bd), rbd = tc::mul2(r, ag); // 1/(s^2 a g), 1/(s b d)
                const float2 itg = tc::mul2(rag, tc::fma2(s2, G, ms2));    // (g-2)/(a g)
                const float2 fi = tc::mul2(rbd, D);                         // 1/b
                ogx[pr] = tc::mul2(rbd, B);                                 // 1/(s d)
                const float2 c = tc::fma2(fi, make_float2(cst[u], cst[u + 1]), itg);
                cst[u] = c.x;
                cst[u + 1] = c.y;
                // AIDAX_POLY_PAIRS: how many of the group's unit pairs take the tanh(c') exponential on the FMA pipe
                if (pr < AIDAX_POLY_PAIRS) ec[pr] = scaled_one_plus_exp2_fma(tc::mul2(c, tc::splat2(2.f * L2E)));
                else ec[pr] = scaled_one_plus_exp2(tc::mul2(c, tc::splat2(2.f * L2E)));
            }
            if constexpr (TWO) inv4(ec[0], ec[1], ec[2], ec[3], t[0], t[1], t[2], t[3]);
            else inv2(ec[0], ec[1], t[0], t[1]);
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) hn[pr] = tc::mul2(ogx[pr], tc::fma2(ms2, t[pr], s2)); // (1/d) * (1 - 2/(1+2^(2c')))
        } else if (MATH == 0 && CELL == 0) {
            // LSTM, RTN/lstm/lstm_xsimd.h:193-213, 271-283: c' = i*tanh(g) + f*c ; h' = o*tanh(c')
            float2 og[NP], ec[NP], t[NP];
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) {
                const uint32_t(&raw)[16] = pr < 2 ? raw0 : raw1;
                const int j0 = 4 * (pr & 1), u = 4 * q + 2 * pr;
                float2 ii, fi, gi;
                inv4(ope2<0>(pair(raw, j0)), ope2<0>(pair(raw, j0 + 1)), ope2<1>(pair(raw, j0 + 2)), ope2<0>(pair(raw, j0 + 3)), ii, fi, gi,
                     og[pr]);
                const float2 c = tc::fma2(ii, tc::fma2(m2, gi, one), tc::mul2(fi, make_float2(cst[u], cst[u + 1])));
                cst[u] = c.x;
                cst[u + 1] = c.y;
                ec[pr] = ope2<2>(tc::mul2(c, tc::splat2(2.f * L2E)));
            }
            if constexpr (TWO) inv4(ec[0], ec[1], ec[2], ec[3], t[0], t[1], t[2], t[3]);
            else inv2(ec[0], ec[1], t[0], t[1]);
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) hn[pr] = tc::mul2(og[pr], tc::fma2(m2, t[pr], one));
        } else if (MATH == 0) {
            // GRU, RTN/gru/gru_xsimd.h:219-239, 265-271: h' = (1-z)*tanh(r*(Uh.h+bh1) + Wh.x+bh0) + z*h
            float2 z[NP], e[NP], t[NP];
#pragma unroll
            for (int c4 = 0; c4 < NP / 2; ++c4) {
                const uint32_t(&raw)[16] = c4 == 0 ? raw0 : raw1;
                float2 r0, r1;
                inv4(ope2<0>(pair(raw, 0)), ope2<0>(pair(raw, 1)), ope2<0>(pair(raw, 4)), ope2<0>(pair(raw, 5)), z[2 * c4], r0, z[2 * c4 + 1], r1);
                if constexpr (AIDAX_SCALED_GATES) { // the candidate's 1+2^e scaled by 2^-24, see the LSTM branch
                    e[2 * c4] = scaled_one_plus_exp2(tc::fma2(r0, pair(raw, 2), pair(raw, 3)));
                    e[2 * c4 + 1] = scaled_one_plus_exp2(tc::fma2(r1, pair(raw, 6), pair(raw, 7)));
                } else {
                    e[2 * c4] = ope2<2>(tc::fma2(r0, pair(raw, 2), pair(raw, 3)));
                    e[2 * c4 + 1] = ope2<2>(tc::fma2(r1, pair(raw, 6), pair(raw, 7)));
                }
            }
            if constexpr (TWO) inv4(e[0], e[1], e[2], e[3], t[0], t[1], t[2], t[3]);
            else inv2(e[0], e[1], t[0], t[1]);
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) {
                const int u = 4 * q + 2 * pr;
                const float2 hh = tc::fma2(AIDAX_SCALED_GATES ? tc::splat2(-2.0f * 5.9604644775390625e-08f) : m2, t[pr], one);
                hn[pr] = tc::fma2(tc::sub2(one, z[pr]), hh, tc::mul2(z[pr], make_float2(cst[u], cst[u + 1])));
                cst[u] = hn[pr].x;
                cst[u + 1] = hn[pr].y;
            }
        } else {
            // precise math (expf / tanhf / IEEE division), scalar
#pragma unroll
            for (int i = 0; i < 2 * NP; ++i) {
                const uint32_t(&raw)[16] = i < 4 ? raw0 : raw1;
                const int u = 4 * q + i, cb = 8 * ((i & 3) / 2) + (i & 1); // column of gate g: cb + 2g
                const float a0 = __uint_as_float(raw[cb]), a1 = __uint_as_float(raw[cb + 2]);
                const float a2 = __uint_as_float(raw[cb + 4]), a3 = __uint_as_float(raw[cb + 6]);
                float h;
                if (CELL == 0) {
                    const float ig = sigmoid_f<1>(a0), fg = sigmoid_f<1>(a1), gg = tanh_f<1>(a2), og = sigmoid_f<1>(a3);
                    const float c = fmaf(ig, gg, fg * cst[u]);
                    cst[u] = c;
                    h = og * tanh_f<1>(c);
                } else {
                    const float zg = sigmoid_f<1>(a0), rg = sigmoid_f<1>(a1);
                    const float hh = tanh_f<1>(fmaf(rg, a2, a3));
                    h = fmaf(1.0f - zg, hh, zg * cst[u]);
                    cst[u] = h;
                }
                if (i & 1) hn[i / 2].y = h;
                else hn[i / 2].x = h;
            }
        }
    }

    // n steps of the cell for this thread's stream.  xcol: the stream's column in the audio tile (nullptr = zero input,
    // pre-warm); p1col/p2col: param columns or nullptr (then p1c/p2c).  Every thread of the tile must call this
    // with the same n.  commit: write y into the audio tile.
    __device__ static void run(TcTile& tt, float (&cst)[H], float* xcol, const float* p1col, const float* p2col, float p1c, float p2c, int n,
                               float in_gain, float out_gain, bool in_skip, bool commit)
    {
        constexpr int XP = S::XP;
        const float db = tt.dw[H];

        for (int t = 0; t < n; ++t) {
            const float xg = xcol ? xcol[t * XP] * in_gain : 0.f;
            const float v1 = p1col ? p1col[t * XP] : p1c;
            const float v2 = p2col ? p2col[t * XP] : p2c;
            store_chunk(tt, H / 4, xg, v1, v2, 1.0f);

            // operands written through the generic proxy -> visible to the tensor core; everyone done with D
            tc::fence_proxy_async();
            tc::tc_fence_before();
            tc::bar_sync(tt.bar_id, 128);
#if !(AIDAX_WHATIF & 1)
            if (tt.issuer) {
                if (tc::elect_one()) {
                    tc::tc_fence_after();
                    issue_mmas(tt);
                }
                __syncwarp();
            }
            tc::mbar_wait(tt.bar_addr, tt.phase);
            tt.phase ^= 1u;
#endif
            tc::tc_fence_after();

            // epilogue: groups of two 16-column chunks (8 hidden units) are loaded together and evaluated as one
            // straight-line block, so the 8 + 2 batched reciprocal chains of a group overlap each other's MUFU latency
            float2 y2 = make_float2(db, 0.f);
            // h(t) of the groups evaluated before the second accumulator half is complete: the MMAs still in flight read
            // h(t-1) from the A tile, so these stores wait for them
            float2 pend[S::QWAIT > 0 ? S::QWAIT * 2 : 1];
#pragma unroll
            for (int q = 0; q < H / 4; q += AIDAX_TC_GROUP) {
                constexpr int NQ = H / 4;
                const bool two = AIDAX_TC_GROUP == 2 && q + 1 < NQ;
                if (S::NSPLIT == 2 && q == S::QWAIT) {
#if !(AIDAX_WHATIF & 1)
                    tc::mbar_wait(tt.bar_addr + 8, tt.phase ^ 1u); // tt.phase was flipped after the first half's wait
                    tc::tc_fence_after();
#endif
#pragma unroll
                    for (int c = 0; c < S::QWAIT; ++c) store_chunk(tt, c, pend[2 * c], pend[2 * c + 1]);
                }
                uint32_t raw0[16], raw1[16];
                tc::tmem_ld16(tt.tmem_ld + 16 * q, raw0);
                if (two) tc::tmem_ld16(tt.tmem_ld + 16 * (q + 1), raw1);
                tc::tmem_ld_wait16(raw0);
                float2 hn[4];
                if (two) {
                    tc::tmem_ld_wait16(raw1); // already complete (one wait covers all outstanding loads); keeps the asm ordering explicit
                    eval_group<true>(raw0, raw1, q, cst, hn);
                } else {
                    eval_group<false>(raw0, raw0, q, cst, hn);
                }
                const float2* dw2 = reinterpret_cast<const float2*>(tt.dw + 4 * q);
                const bool hold = S::NSPLIT == 2 && q < S::QWAIT;
                y2 = tc::fma2(dw2[0], hn[0], y2);
                y2 = tc::fma2(dw2[1], hn[1], y2);
                if (hold) { pend[2 * q] = hn[0]; pend[2 * q + 1] = hn[1]; }
                else store_chunk(tt, q, hn[0], hn[1]);
                if (two) {
                    y2 = tc::fma2(dw2[2], hn[2], y2);
                    y2 = tc::fma2(dw2[3], hn[3], y2);
                    if (hold) { pend[2 * q + 2] = hn[2]; pend[2 * q + 3] = hn[3]; }
                    else store_chunk(tt, q + 1, hn[2], hn[3]);
                }
            }
            // dense + skip + output gain (applyModel, rt-neural-generic.cpp:170-182)
            if (xcol && commit) {
                const float y = y2.x + y2.y;
                xcol[t * XP] = (in_skip ? xg + y : y) * out_gain;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// the fused chain kernel on the tensor-core recurrence.  CELL: 0 LSTM, 1 GRU
// ------------------------------------------------------------------------------------------
template <int CELL, int H, int NT, int RES, int MATH>
__global__ void __launch_bounds__(NT * 128, RES) chain_kernel_tc(const __grid_constant__ RenderParams p)
{
    using S = TcShape<H>;
    using Cell = TcCell<CELL, H, MATH>;
    constexpr int XP = S::XP;
    constexpr int NTHREADS = NT * 128;
    constexpr int CS = NT * 128;
    extern __shared__ __align__(128) float smem_tc[];
    float* smem = smem_tc;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0); // broadcast: the compiler may treat it as warp-uniform
    const int tile = warp >> 2, r = tid & 127;
    const int blk = blockIdx.x;
    const StreamArrays& sa = p.sa;

    float* Bh = smem;
    float* Bl = Bh + S::B_FLOATS;
    float* Abase = Bl + S::B_FLOATS;
    float* dw = Abase + NT * 2 * S::A_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(dw + H + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NT);

    const ModelDev md = p.models[p.blk_model[blk]];
    const int input_size = md.input_size;
    const int n_ptiles = input_size - 1;


    // ---- B operand (hi, lo) from the packed weight block (layout.h): B[n = column_of(unit, gate)][k]
    {
        const float* wk = md.weights;
        const float* bias = wk + (H + 3) * H * 4;
        for (int idx = tid; idx < S::N * S::KP; idx += NTHREADS) {
            const int n = idx % S::N, k = idx / S::N;
            const int g = Cell::gate_of_column(n), src = 4 * Cell::unit_of_column(n) + g; // packed block: column 4*unit + gate
            float w = 0.f;
            if (k < H + 3) w = __ldg(wk + (size_t)k * S::N + src);
            else if (k == H + 3) w = __ldg(bias + src);
            if (MATH == 0) w *= Cell::column_scale(g); // the accumulator is directly the exponent of 2 the gate needs
            float hi, lo;
#if AIDAX_B_RN
            tc::split_tf32_rn(w, hi, lo);
#else
            tc::split_tf32(w, hi, lo);
#endif
            const int off = (k >> 2) * (S::N * 4) + n * 4 + (k & 3);
            Bh[off] = hi;
            Bl[off] = lo;
        }
        for (int i = tid; i < H + 1; i += NTHREADS) dw[i] = __ldg(bias + H * 4 + i);
        // the K padding of A must be finite zeros (B's padding rows are zero as well)
        for (int i = tid; i < NT * 2 * S::A_FLOATS; i += NTHREADS) Abase[i] = 0.f;
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 2 * NT; ++i) tc::mbar_init(&bars[i], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, S::tmem_cols(NT));
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    TcTile tt;
    tt.a_hi = Abase + tile * 2 * S::A_FLOATS + r * 4;
    tt.a_lo = tt.a_hi + S::A_FLOATS;
    {
        constexpr uint32_t LBO_A = 128 * 16, LBO_B = S::N * 16, SBO = 128;
        const uint32_t a_addr = tc::smem_u32(Abase + tile * 2 * S::A_FLOATS);
        tt.da_hi = (tc_desc_t)tc::make_smem_desc(a_addr, LBO_A, SBO);
        tt.da_lo = (tc_desc_t)tc::make_smem_desc(a_addr + S::A_FLOATS * 4, LBO_A, SBO);
        tt.db_hi = (tc_desc_t)tc::make_smem_desc(tc::smem_u32(Bh), LBO_B, SBO);
        tt.db_lo = (tc_desc_t)tc::make_smem_desc(tc::smem_u32(Bl), LBO_B, SBO);
    }
    tt.bar_addr = tc::smem_u32(&bars[2 * tile]); // + 8: the second accumulator half's barrier
    tt.phase = 0;
    tt.tmem_d = tmem_base + tile * S::N;
    tt.tmem_ld = tt.tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    tt.bar_id = 1 + tile;
    tt.dw = dw;
    tt.issuer = (warp & 3) == 0;

    float* wbase = smem + ((S::fixed_floats(NT) + 3) & ~3) + (size_t)warp * S::warp_floats(p.ts, n_ptiles);
    int* wslots = reinterpret_cast<int*>(wbase);
    float* xbuf = wbase + 32;
    float* p1buf = n_ptiles >= 1 ? xbuf + p.ts * XP : nullptr;
    float* p2buf = n_ptiles >= 2 ? p1buf + p.ts * XP : nullptr;

    wslots[lane] = p.blk_slots[(size_t)blk * CS + warp * 32 + lane];
    __syncwarp();
    const int cslot = wslots[lane];
    const int slot = cslot < 0 ? 0 : cslot;

    // ---- commands (open / activate / model swap) and the pre-warm they may request
    uint32_t cmd = 0;
    if (cslot >= 0) cmd = exec_commands(sa, cslot);
    const bool want_prewarm = (cmd & C_SET_MODEL) != 0;
    if (__syncthreads_or(want_prewarm)) {
        // loadModelFromPath: reset() then applyModel over 2048 zero samples with the param smoothers parked at the
        // old targets (rt-neural-generic.cpp:1035, 1052-1079)
        const float p1c = sa.p_mem[slot], p2c = sa.p_mem[(size_t)sa.cap + slot];
        float cst[H];
        Cell::load_state(tt, cst, sa, slot, false);
        Cell::run(tt, cst, nullptr, nullptr, nullptr, p1c, p2c, kPrewarmSteps, md.in_gain, md.out_gain, md.in_skip != 0, false);
        if (want_prewarm) Cell::store_state(tt, cst, sa, slot);
    }

    uint32_t flags = 0, dfl = 0;
    if (cslot >= 0) { flags = sa.flags[cslot]; dfl = sa.dflags[cslot]; }
    const bool enabled = (flags & F_ENABLED) != 0;
    const bool net_on = enabled && (flags & F_NET);
    const bool any_net = __syncthreads_or(net_on) != 0;
    bool first_run = (dfl & D_PARAM_FIRST_RUN) != 0;

    float cst[H]; // recurrent state of this thread's stream, in registers across the whole period
    if (any_net) Cell::load_state(tt, cst, sa, slot, cslot >= 0);

    for (int t0 = 0; t0 < p.n_samples; t0 += p.ts) {
        const int ts = min(p.ts, p.n_samples - t0);
        load_tile<32, XP>(xbuf, p, wslots, t0, ts, lane);
#if AIDAX_PREFETCH_L2
        // the NEXT sub-block's rows go to L2 while this one is filtered and run through the network, so that its tile
        // load does not wait on HBM with nothing else to issue (ncu source view: 7 % of all warp samples sat on those
        // loads' scoreboard).  One sub-block ahead only: prefetching the whole period up front cost 25 % more DRAM reads
        // (lines evicted before use) for no extra gain.
        if (t0 + ts < p.n_samples) {
            const int tn = t0 + ts, nn = min(p.ts, p.n_samples - tn);
            if (p.layout == 0) { // planar [stream][n]: each lane owns its stream's row
                if (cslot >= 0) {
                    const float* src = p.in + (size_t)(cslot - p.first_stream) * p.stride + tn;
                    for (int t = 0; t < nn; t += 32) tc::prefetch_l2(src + t);
                }
            } else { // interleaved [n][stream]: one 128-byte segment per sample row for the warp; lane l takes rows l, l+32, ...
                const int s0 = wslots[0];
                if (s0 >= 0) {
                    const float* src = p.in + (size_t)tn * p.stride + (s0 - p.first_stream);
                    for (int t = lane; t < nn; t += 32) tc::prefetch_l2(src + (size_t)t * p.stride);
                }
            }
        }
#endif
        __syncwarp();

        // ---- stages before the network (run(), rt-neural-generic.cpp:622-644)
        if (cslot >= 0 && enabled && !(AIDAX_WHATIF & 8)) {
            float* xcol = xbuf + lane;
#if AIDAX_FUSED_PASSES
            pass_fused<XP, false>(xcol, ts, sa, cslot, fused_stages_pre(flags));
#else
            pass_lpf_pregain<XP>(xcol, ts, sa, cslot, (flags & F_LPF) != 0);
            if (flags & F_EQ_PRE) pass_eq<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0);
#endif
            if (net_on) {
                pass_param<XP>(p1buf ? p1buf + lane : nullptr, ts, sa, 0, cslot, first_run, input_size >= 2);
                pass_param<XP>(p2buf ? p2buf + lane : nullptr, ts, sa, 1, cslot, first_run, input_size >= 3);
                if (first_run) { first_run = false; sa.dflags[cslot] = dfl & ~D_PARAM_FIRST_RUN; }
            }
        } else if (n_ptiles > 0) {
            // the param columns of idle rows still feed the (discarded) product: keep them finite
            for (int t = 0; t < ts; ++t) {
                p1buf[t * XP + lane] = 0.f;
                if (p2buf) p2buf[t * XP + lane] = 0.f;
            }
        }
        __syncwarp();

        // ---- the network
        if (any_net)
            Cell::run(tt, cst, xbuf + lane, p1buf ? p1buf + lane : nullptr, p2buf ? p2buf + lane : nullptr, 0.f, 0.f, ts, md.in_gain,
                      md.out_gain, md.in_skip != 0, net_on);

        // ---- stages after the network (:645-655)
        if (cslot >= 0 && enabled && !(AIDAX_WHATIF & 8)) {
            float* xcol = xbuf + lane;
#if AIDAX_FUSED_PASSES
            pass_fused<XP, true>(xcol, ts, sa, cslot, fused_stages_post(flags));
#else
            const bool dc_on = (flags & F_DC) != 0;
            if (flags & F_EQ_POST) {
                if (dc_on) pass_biquad<XP>(xcol, ts, sa, BQ_DC, cslot);
                pass_eq<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0);
                pass_dc_master<XP>(xcol, ts, sa, cslot, false, true);
            } else {
                pass_dc_master<XP>(xcol, ts, sa, cslot, dc_on, true);
            }
#endif
        }
        __syncwarp();
        store_tile<32, XP>(xbuf, p, wslots, t0, ts, lane);
        __syncwarp();
    }

    if (net_on) Cell::store_state(tt, cst, sa, slot);

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, S::tmem_cols(NT));
}

} // namespace aidax
