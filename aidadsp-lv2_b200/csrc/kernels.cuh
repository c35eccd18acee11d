// kernels.cuh -- sm_100a kernels of the fused rt-neural-generic chain.
//
// One launch renders one audio period (n_samples) for every stream of one model family
// (cell type x hidden size).  The audio of a warp's streams makes ONE HBM round trip: it is staged
// into shared memory, every stage of RtNeuralGeneric::run (rt-neural-generic.cpp:621-655) is applied
// in place there, and the result is written back.
//
// Work decomposition (template parameters):
//   a warp is split into G = 32/L groups of L lanes; a group owns S streams of one model;
//   lane l of a group owns the U = H/L hidden units {u*L + l}, for all S streams of its group.
//   Per time step the gate pre-activations are a register-tiled [S x 4U] x K contraction:
//   the recurrent vector h (S streams) and the weight rows (U units x 4 gates) are read from
//   shared memory with 128-bit loads (broadcast inside / across groups, conflict free), each
//   loaded value feeds 4U resp. S FFMAs.  The cell update is thread-local (a lane holds all four
//   gates of its units and their c), the new h is exchanged through a double-buffered
//   shared-memory tile guarded by one __syncwarp per step -- no block barrier in the time loop.
//   Small S / large L = low latency per buffer, large S = high FFMA-per-load throughput.
//
// The filter/gain stages run "thread per stream" (lane w <-> stream w of the warp), stage-major
// over the staged tile exactly like the reference runs them buffer-major, with double-precision
// biquad state and non-contracted arithmetic (bit-identical to the SSE2 reference build).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"

namespace aidax {

// ------------------------------------------------------------------------------------------
// math helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MATH>
__device__ __forceinline__ float sigmoid_f(float x)
{
    if constexpr (MATH == 0) {
        // 1/(1+exp(-x)), RTN/lstm/lstm_xsimd.h:347-350; exp via MUFU.EX2, division via MUFU.RCP
        return rcp_approx(1.0f + ex2_approx(x * -1.4426950408889634f));
    } else {
        return 1.0f / (1.0f + expf(-x));
    }
}
template <int MATH>
__device__ __forceinline__ float tanh_f(float x)
{
    if constexpr (MATH == 0) {
        // tanh(x) = 1 - 2/(1+exp(2x)); saturates correctly at +-inf of the exponential
        return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(x * 2.8853900817779268f)), 1.0f);
    } else {
        return tanhf(x);
    }
}

// ------------------------------------------------------------------------------------------
// filter / gain stage primitives (thread per stream, column `col` of a [ts][XP] smem tile)
// ------------------------------------------------------------------------------------------

// Biquad::process, common/Biquad.h:53-58.  Double state/coefficients, float in/out, mul and add
// kept separate (the reference's default build has no FMA contraction).
struct BiquadReg {
    double a0, a1, a2, b1, b2, z1, z2;
    __device__ __forceinline__ void load(const StreamArrays& sa, int which, int slot)
    {
        const size_t cap = (size_t)sa.cap;
        const double* c = sa.bq_coef + (size_t)which * 5 * cap + slot;
        a0 = c[0]; a1 = c[cap]; a2 = c[2 * cap]; b1 = c[3 * cap]; b2 = c[4 * cap];
        const double* z = sa.bq_z + (size_t)which * 2 * cap + slot;
        z1 = z[0]; z2 = z[cap];
    }
    __device__ __forceinline__ void store(const StreamArrays& sa, int which, int slot) const
    {
        const size_t cap = (size_t)sa.cap;
        double* z = sa.bq_z + (size_t)which * 2 * cap + slot;
        z[0] = z1; z[cap] = z2;
    }
    // float -> double is exact; WI = 1 builds the double with integer instructions instead of F2F.F64.F32, which runs on
    // the XU pipe next to the gates' MUFU.EX2/RCP (the pipe that bounds the tensor-core kernels).  Zero, subnormal, infinite
    // and NaN inputs (exponent field 0 or 255) take the conversion instruction, so every input maps to the same double.
    template <int WI>
    __device__ __forceinline__ static double widen(float in)
    {
        if constexpr (WI == 0) {
            return (double)in;
        } else {
            const uint32_t b = __float_as_uint(in);
            if ((((b >> 23) + 1u) & 0xFEu) == 0u) return (double)in;
            const uint32_t hi = (b & 0x80000000u) | (((b & 0x7FFFFFFFu) >> 3) + 0x38000000u);
            return __hiloint2double((int)hi, (int)(b << 29));
        }
    }
    template <int WI = 0>
    __device__ __forceinline__ float process(float in)
    {
        const double x = widen<WI>(in);
        const double o = __dadd_rn(__dmul_rn(x, a0), z1);
        z1 = __dsub_rn(__dadd_rn(__dmul_rn(x, a1), z2), __dmul_rn(b1, o));
        z2 = __dsub_rn(__dmul_rn(x, a2), __dmul_rn(b2, o));
        return __double2float_rn(o);
    }
};

// ExponentialValueSmoother::next, common/ValueSmoother.hpp:142-145 (no contraction)
__device__ __forceinline__ float exp_smooth_next(float& mem, float coef, float target_times_one_minus_coef)
{
    mem = __fadd_rn(__fmul_rn(mem, coef), target_times_one_minus_coef);
    return mem;
}

template <int XP, int WI = 0>
__device__ __forceinline__ void pass_biquad(float* xcol, int ts, const StreamArrays& sa, int which, int slot)
{
    BiquadReg f;
    f.load(sa, which, slot);
#pragma unroll 4
    for (int t = 0; t < ts; ++t) xcol[t * XP] = f.template process<WI>(xcol[t * XP]);
    f.store(sa, which, slot);
}

// applyToneControls cascade, rt-neural-generic.cpp:129-139: depth -> bass -> mid -> treble -> presence,
// every stage narrowed to float; band-pass mode runs `mid` alone.
template <int XP>
__device__ __forceinline__ void pass_eq(float* xcol, int ts, const StreamArrays& sa, int slot, bool bandpass)
{
    if (bandpass) {
        pass_biquad<XP>(xcol, ts, sa, BQ_MID, slot);
        return;
    }
    BiquadReg f0, f1, f2, f3, f4;
    f0.load(sa, BQ_DEPTH, slot);
    f1.load(sa, BQ_BASS, slot);
    f2.load(sa, BQ_MID, slot);
    f3.load(sa, BQ_TREBLE, slot);
    f4.load(sa, BQ_PRESENCE, slot);
    if (ts >= 5) {
        // Software-pipelined over time: in one iteration stage k works on sample t-k, so the five double-precision
        // recurrences (and their float<->double conversions) are independent instructions streams instead of one
        // 5-deep dependent chain per sample.  Every filter still sees exactly its own input sequence: bit-identical.
        float l0 = f0.process(xcol[0]);
        float l1 = f1.process(l0); l0 = f0.process(xcol[XP]);
        float l2 = f2.process(l1); l1 = f1.process(l0); l0 = f0.process(xcol[2 * XP]);
        float l3 = f3.process(l2); l2 = f2.process(l1); l1 = f1.process(l0); l0 = f0.process(xcol[3 * XP]);
#pragma unroll 2
        for (int t = 4; t < ts; ++t) {
            const float o = f4.process(l3);
            l3 = f3.process(l2);
            l2 = f2.process(l1);
            l1 = f1.process(l0);
            l0 = f0.process(xcol[t * XP]);
            xcol[(t - 4) * XP] = o;
        }
        xcol[(ts - 4) * XP] = f4.process(l3); l3 = f3.process(l2); l2 = f2.process(l1); l1 = f1.process(l0);
        xcol[(ts - 3) * XP] = f4.process(l3); l3 = f3.process(l2); l2 = f2.process(l1);
        xcol[(ts - 2) * XP] = f4.process(l3); l3 = f3.process(l2);
        xcol[(ts - 1) * XP] = f4.process(l3);
    } else {
        for (int t = 0; t < ts; ++t) {
            float x = xcol[t * XP];
            x = f0.process(x);
            x = f1.process(x);
            x = f2.process(x);
            x = f3.process(x);
            x = f4.process(x);
            xcol[t * XP] = x;
        }
    }
    f0.store(sa, BQ_DEPTH, slot);
    f1.store(sa, BQ_BASS, slot);
    f2.store(sa, BQ_MID, slot);
    f3.store(sa, BQ_TREBLE, slot);
    f4.store(sa, BQ_PRESENCE, slot);
}

// Stage 1+2 of run(): optional in_lpf then the pre-gain ramp (:622-627)
template <int XP, int WI = 0>
__device__ __forceinline__ void pass_lpf_pregain(float* xcol, int ts, const StreamArrays& sa, int slot, bool lpf_on)
{
    const float coef = sa.exp_coef[slot];
    const float tgt = __fmul_rn(sa.pre_target[slot], __fsub_rn(1.f, coef));
    float mem = sa.pre_mem[slot];
    if (lpf_on) {
        BiquadReg f;
        f.load(sa, BQ_INLPF, slot);
#pragma unroll 4
        for (int t = 0; t < ts; ++t) {
            const float x = f.template process<WI>(xcol[t * XP]);
            xcol[t * XP] = __fmul_rn(x, exp_smooth_next(mem, coef, tgt));
        }
        f.store(sa, BQ_INLPF, slot);
    } else {
#pragma unroll 4
        for (int t = 0; t < ts; ++t) xcol[t * XP] = __fmul_rn(xcol[t * XP], exp_smooth_next(mem, coef, tgt));
    }
    sa.pre_mem[slot] = mem;
}

// Stage 7 (master ramp, :654-655), optionally preceded by the dc blocker (:645-650) when no
// post-EQ sits between them.
template <int XP, int WI = 0>
__device__ __forceinline__ void pass_dc_master(float* xcol, int ts, const StreamArrays& sa, int slot, bool dc_on, bool master_on)
{
    const float coef = sa.exp_coef[slot];
    const float tgt = __fmul_rn(sa.master_target[slot], __fsub_rn(1.f, coef));
    float mem = sa.master_mem[slot];
    BiquadReg f;
    if (dc_on) f.load(sa, BQ_DC, slot);
#pragma unroll 4
    for (int t = 0; t < ts; ++t) {
        float x = xcol[t * XP];
        if (dc_on) x = f.template process<WI>(x);
        if (master_on) x = __fmul_rn(x, exp_smooth_next(mem, coef, tgt));
        xcol[t * XP] = x;
    }
    if (dc_on) f.store(sa, BQ_DC, slot);
    if (master_on) sa.master_mem[slot] = mem;
}

// All filter/gain stages on one side of the network as ONE software-pipelined pass: up to six biquads and the side's
// gain ramp.  PRE : in_lpf -> [pre-gain ramp] -> depth -> bass -> mid -> treble -> presence   (run(), :622-630)
//           POST: dc blocker -> depth -> bass -> mid -> treble -> presence -> [master ramp]   (:645-655)
// `on` bit k: stage k's biquad runs (an EQ that is bypassed, on the other side of the network or in band-pass mode
// switches its stages off).  In iteration t stage k works on sample t-k, so the six double-precision recurrences
// (serial chains of three dependent DMUL/DADD per sample, plus the float<->double conversions on the XU pipe) are
// independent instruction streams; with the stages as separate passes the single-biquad ones (in_lpf + gain,
// dc + master) ran one dependent chain per thread.  Every filter still sees exactly its own input sequence with the
// reference's operation order and the float narrowing between stages: bit-identical to the separate passes.
template <int XP, bool POST>
__device__ __forceinline__ void pass_fused(float* xcol, int ts, const StreamArrays& sa, int slot, uint32_t on)
{
    constexpr int W[6] = {POST ? BQ_DC : BQ_INLPF, BQ_DEPTH, BQ_BASS, BQ_MID, BQ_TREBLE, BQ_PRESENCE};
    constexpr int GA = POST ? 5 : 0; // the gain ramp multiplies the value that leaves this stage
    const float coef = sa.exp_coef[slot];
    const float tgt = __fmul_rn(POST ? sa.master_target[slot] : sa.pre_target[slot], __fsub_rn(1.f, coef));
    float mem = POST ? sa.master_mem[slot] : sa.pre_mem[slot];
    BiquadReg f[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        f[k].a0 = f[k].a1 = f[k].a2 = f[k].b1 = f[k].b2 = f[k].z1 = f[k].z2 = 0.0;
        if ((on >> k) & 1u) f[k].load(sa, W[k], slot);
    }
    auto stage = [&](int k, float v) -> float {
        if ((on >> k) & 1u) v = f[k].process(v);
        if (k == GA) v = __fmul_rn(v, exp_smooth_next(mem, coef, tgt));
        return v;
    };
    float l[5];
    if (ts >= 6) {
#pragma unroll
        for (int t = 0; t < 5; ++t) {  // fill: stage k <= t takes sample t - k
#pragma unroll
            for (int k = 4; k >= 0; --k)
                if (k <= t) l[k] = stage(k, k == 0 ? xcol[t * XP] : l[k - 1]);
        }
#pragma unroll 2
        for (int t = 5; t < ts; ++t) {
            const float o = stage(5, l[4]);
            l[4] = stage(4, l[3]);
            l[3] = stage(3, l[2]);
            l[2] = stage(2, l[1]);
            l[1] = stage(1, l[0]);
            l[0] = stage(0, xcol[t * XP]);
            xcol[(t - 5) * XP] = o;
        }
#pragma unroll
        for (int d = 0; d < 5; ++d) {  // drain: stage k > d takes sample ts + d - k
            xcol[(ts + d - 5) * XP] = stage(5, l[4]);
#pragma unroll
            for (int k = 4; k > d; --k) l[k] = stage(k, l[k - 1]);
        }
    } else {
        for (int t = 0; t < ts; ++t) {
            float v = xcol[t * XP];
#pragma unroll
            for (int k = 0; k < 6; ++k) v = stage(k, v);
            xcol[t * XP] = v;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if ((on >> k) & 1u) f[k].store(sa, W[k], slot);
    if (POST) sa.master_mem[slot] = mem;
    else sa.pre_mem[slot] = mem;
}
// which stages of pass_fused run for a stream's flag word
__device__ __forceinline__ uint32_t fused_stages_pre(uint32_t flags)
{
    return ((flags & F_LPF) ? 1u : 0u) | ((flags & F_EQ_PRE) ? ((flags & F_EQ_BANDPASS) ? 1u << 3 : 0x3Eu) : 0u);
}
__device__ __forceinline__ uint32_t fused_stages_post(uint32_t flags)
{
    return ((flags & F_DC) ? 1u : 0u) | ((flags & F_EQ_POST) ? ((flags & F_EQ_BANDPASS) ? 1u << 3 : 0x3Eu) : 0u);
}

// LinearValueSmoother for PARAM1/2 (common/ValueSmoother.hpp:209-240) as driven by run()
// (rt-neural-generic.cpp:633-641) and applyModel (:194, 218-219): target change detection with
// FLT_EPSILON, step recomputed only on change, first-run snap, one next() per sample.
template <int XP>
__device__ __forceinline__ void pass_param(float* pcol, int ts, const StreamArrays& sa, int which, int slot, bool first_run, bool advance)
{
    const size_t o = (size_t)which * sa.cap + slot;
    float target = sa.p_target[o], mem = sa.p_mem[o], step = sa.p_step[o];
    const float port = sa.p_port[o];
    if (fabsf(target - port) >= 1.1920929e-07f) {
        target = port;
        step = __fdiv_rn(__fsub_rn(target, mem), sa.lin_denom[slot]);
    }
    if (first_run) mem = target;
    if (advance) {
        const float astep = fabsf(step);
        for (int t = 0; t < ts; ++t) {
            const float dy = __fsub_rn(target, mem);
            mem = __fadd_rn(mem, copysignf(fminf(fabsf(dy), astep), dy));
            pcol[t * XP] = mem;
        }
    }
    sa.p_target[o] = target; sa.p_mem[o] = mem; sa.p_step[o] = step;
}

// ------------------------------------------------------------------------------------------
// audio tile staging: global <-> shared, both layouts, coalesced on the global side
// ------------------------------------------------------------------------------------------
template <int WS, int XP>
__device__ __forceinline__ void load_tile(float* xbuf, const RenderParams& p, const int* wslots, int t0, int ts, int lane)
{
    if (p.layout == 0) { // planar [stream][n]
        for (int j = 0; j < WS; ++j) {
            const int slot = wslots[j];
            if (slot < 0) { for (int t = lane; t < ts; t += 32) xbuf[t * XP + j] = 0.f; continue; }
            const float* src = p.in + (size_t)(slot - p.first_stream) * p.stride + t0;
            for (int t = lane; t < ts; t += 32) xbuf[t * XP + j] = __ldg(src + t);
        }
    } else if constexpr (WS == 32) { // interleaved [n][stream], one stream per lane: 8 rows of loads in flight per round
        const int slot = wslots[lane];
        const float* src = p.in + (size_t)t0 * p.stride + (slot - p.first_stream);
        for (int t = 0; t < ts; t += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (slot >= 0 && t + i < ts) ? __ldg(src + (size_t)(t + i) * p.stride) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (t + i < ts) xbuf[(t + i) * XP + lane] = v[i];
        }
    } else if constexpr (WS == 1) { // interleaved [n][stream], one stream per warp: a lane's loads all in flight before the first store
        const int slot = wslots[0];
        const float* src = p.in + (size_t)t0 * p.stride + (slot - p.first_stream);
        for (int t = lane; t < ts; t += 128) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (slot >= 0 && t + 32 * i < ts) ? __ldg(src + (size_t)(t + 32 * i) * p.stride) : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (t + 32 * i < ts) xbuf[(t + 32 * i) * XP] = v[i];
        }
    } else { // interleaved [n][stream]
        for (int idx = lane; idx < ts * WS; idx += 32) {
            const int t = idx / WS, j = idx - t * WS;
            const int slot = wslots[j];
            xbuf[t * XP + j] = slot < 0 ? 0.f : __ldg(p.in + (size_t)(t0 + t) * p.stride + (slot - p.first_stream));
        }
    }
}
template <int WS, int XP>
__device__ __forceinline__ void store_tile(const float* xbuf, const RenderParams& p, const int* wslots, int t0, int ts, int lane)
{
    if (p.layout == 0) {
        for (int j = 0; j < WS; ++j) {
            const int slot = wslots[j];
            if (slot < 0) continue;
            float* dst = p.out + (size_t)(slot - p.first_stream) * p.stride + t0;
            for (int t = lane; t < ts; t += 32) dst[t] = xbuf[t * XP + j];
        }
    } else {
        for (int idx = lane; idx < ts * WS; idx += 32) {
            const int t = idx / WS, j = idx - t * WS;
            const int slot = wslots[j];
            if (slot >= 0) p.out[(size_t)(t0 + t) * p.stride + (slot - p.first_stream)] = xbuf[t * XP + j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// per-stream command execution (open / activate / model swap), see layout.h C_*
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t exec_commands(const StreamArrays& sa, int slot)
{
    const uint32_t c = sa.cmd[slot];
    if (!c) return 0;
    const size_t cap = (size_t)sa.cap;
    if (c & C_SET_GAIN_MEMS) {
        sa.pre_mem[slot] = sa.cmd_pre_mem[slot];
        sa.master_mem[slot] = sa.cmd_master_mem[slot];
    }
    if (c & C_ZERO_FILTERS) {
        for (int i = 0; i < kNumBiquads * 2; ++i) sa.bq_z[(size_t)i * cap + slot] = 0.0;
    }
    uint32_t d = sa.dflags[slot];
    if (c & C_SET_MODEL) {
        // LinearValueSmoother after setSampleRate/setTimeConstant/setTargetValue(old)/clearToTargetValue
        // (rt-neural-generic.cpp:1053-1060): target moves only if it differs from 0 by >= FLT_EPSILON.
        for (int i = 0; i < 2; ++i) {
            const float old = sa.p_init[(size_t)i * cap + slot];
            float target = 0.f, step = 0.f;
            if (fabsf(old) >= 1.1920929e-07f) { target = old; step = __fdiv_rn(old, sa.lin_denom[slot]); }
            sa.p_target[(size_t)i * cap + slot] = target;
            sa.p_mem[(size_t)i * cap + slot] = target;
            sa.p_step[(size_t)i * cap + slot] = step;
        }
        d |= D_PARAM_FIRST_RUN;
    }
    if (c & C_FIRST_RUN) d |= D_PARAM_FIRST_RUN;
    sa.dflags[slot] = d;
    sa.cmd[slot] = 0;
    return c;
}

// ------------------------------------------------------------------------------------------
// the recurrent tile
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void lds_vec(float (&dst)[N], const float* src)
{
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
        }
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float2 v = *reinterpret_cast<const float2*>(src + 2 * i);
            dst[2 * i] = v.x; dst[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = src[i];
    }
}
template <int N>
__device__ __forceinline__ void sts_vec(float* dst, const float (&src)[N])
{
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i)
            *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
    } else if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) *reinterpret_cast<float2*>(dst + 2 * i) = make_float2(src[2 * i], src[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = src[i];
    }
}

template <int CELL, int H, int L, int S>
struct Tile {
    static constexpr int U = H / L;
    static constexpr int G = 32 / L;
    static constexpr int WS = G * S;      // streams per warp
    static constexpr int XP = WS + 4;     // pitch of the audio tile (keeps 16 B alignment of S-vectors)
    static constexpr int HP = WS + 4;     // pitch of the h exchange tile (rows 4 banks apart)
    static_assert(H % L == 0 && 32 % L == 0, "lanes per group must divide H and 32");
    static_assert(S <= L, "a group needs at least one lane per stream for the dense/skip step");

    // shared memory floats needed by one warp for ts samples and input size in (param tiles only if in > 1)
    __host__ __device__ static constexpr int warp_smem_floats(int ts, int n_param_tiles)
    {
        return ts * XP * (1 + n_param_tiles) + 2 * H * HP + 32 /* slot table */;
    }

    // Runs n steps of the cell for the S streams of this lane's group.
    //   xsrc  : audio tile column base (already offset to this group), nullptr => zero input (pre-warm)
    //   p1src/p2src : param tiles or nullptr; pconst used when the tile is nullptr
    //   commit_mask : bit s set => write y into the audio tile and store the state of stream s
    template <int MATH>
    __device__ static void run(const float* __restrict__ wsm, float* __restrict__ hbuf, float* xsrc, const float* p1src,
                               const float* p2src, const float (&p1c)[S], const float (&p2c)[S], int n, int input_size,
                               float in_gain, float out_gain, bool in_skip, const StreamArrays& sa, const int (&slot)[S],
                               uint32_t load_mask, uint32_t commit_mask, int l, int g)
    {
        const float* wk = wsm;                              // [(H+3)][H][4]
        const float* bias = wsm + (H + 3) * H * 4;          // [H][4]
        const float* dw = bias + H * 4;                     // [H]
        const float db = dw[H];
        const size_t cap = (size_t)sa.cap;

        // ---- state: own units' h (exchange + GRU blend) and c (LSTM)
        float cst[S][U];   // LSTM: c ; GRU: own h
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int unit = u * L + l;
                float hv = 0.f, cv = 0.f;
                if ((load_mask >> s) & 1u) {
                    hv = sa.rnn[(size_t)unit * cap + slot[s]];
                    if (CELL == 0) cv = sa.rnn[(size_t)(kMaxHidden + unit) * cap + slot[s]];
                }
                cst[s][u] = CELL == 0 ? cv : hv;
                hbuf[unit * HP + g * S + s] = hv;
            }
        }
        __syncwarp();

        int cur = 0;
        for (int t = 0; t < n; ++t) {
            const float* hc = hbuf + cur * (H * HP) + g * S;
            float* hn = hbuf + (cur ^ 1) * (H * HP) + g * S;

            float acc[S][U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4 b = *reinterpret_cast<const float4*>(bias + (u * L + l) * 4);
#pragma unroll
                for (int s = 0; s < S; ++s) { acc[s][u][0] = b.x; acc[s][u][1] = b.y; acc[s][u][2] = b.z; acc[s][u][3] = b.w; }
            }

            // ---- input rows (k = H .. H+in-1)
            float xin[S];
            if (xsrc) {
                lds_vec<S>(xin, xsrc + t * XP);
#pragma unroll
                for (int s = 0; s < S; ++s) xin[s] *= in_gain;
            } else {
#pragma unroll
                for (int s = 0; s < S; ++s) xin[s] = 0.f;
            }
            for (int i = 0; i < input_size; ++i) {
                float v[S];
                if (i == 0) {
#pragma unroll
                    for (int s = 0; s < S; ++s) v[s] = xin[s];
                } else {
                    const float* ps = (i == 1) ? p1src : p2src;
                    if (ps) lds_vec<S>(v, ps + t * XP);
                    else {
#pragma unroll
                        for (int s = 0; s < S; ++s) v[s] = (i == 1) ? p1c[s] : p2c[s];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4 w = *reinterpret_cast<const float4*>(wk + ((H + i) * H + u * L + l) * 4);
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        acc[s][u][0] = fmaf(w.x, v[s], acc[s][u][0]);
                        acc[s][u][1] = fmaf(w.y, v[s], acc[s][u][1]);
                        if (CELL == 0) acc[s][u][2] = fmaf(w.z, v[s], acc[s][u][2]);
                        acc[s][u][3] = fmaf(w.w, v[s], acc[s][u][3]);
                    }
                }
            }

            // ---- recurrent rows (k = 0 .. H-1): [S] x [U x 4] outer products
#pragma unroll 4
            for (int k = 0; k < H; ++k) {
                float hv[S];
                lds_vec<S>(hv, hc + k * HP);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4 w = *reinterpret_cast<const float4*>(wk + (k * H + u * L + l) * 4);
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        acc[s][u][0] = fmaf(w.x, hv[s], acc[s][u][0]);
                        acc[s][u][1] = fmaf(w.y, hv[s], acc[s][u][1]);
                        acc[s][u][2] = fmaf(w.z, hv[s], acc[s][u][2]);
                        if (CELL == 0) acc[s][u][3] = fmaf(w.w, hv[s], acc[s][u][3]);
                    }
                }
            }

            // ---- cell update (thread local) and h exchange
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float hnew[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    if (CELL == 0) {
                        // LSTM, RTN/lstm/lstm_xsimd.h:193-213, 271-283: c' = i*tanh(g) + f*c ; h' = o*tanh(c')
                        const float ig = sigmoid_f<MATH>(acc[s][u][0]);
                        const float fg = sigmoid_f<MATH>(acc[s][u][1]);
                        const float gg = tanh_f<MATH>(acc[s][u][2]);
                        const float og = sigmoid_f<MATH>(acc[s][u][3]);
                        const float c = fmaf(ig, gg, fg * cst[s][u]);
                        cst[s][u] = c;
                        hnew[s] = og * tanh_f<MATH>(c);
                    } else {
                        // GRU, RTN/gru/gru_xsimd.h:219-239, 265-271: h' = (1-z)*tanh(r*(Uh.h+bh1) + Wh.x+bh0) + z*h
                        const float zg = sigmoid_f<MATH>(acc[s][u][0]);
                        const float rg = sigmoid_f<MATH>(acc[s][u][1]);
                        const float hh = tanh_f<MATH>(fmaf(rg, acc[s][u][2], acc[s][u][3]));
                        const float hv = fmaf(1.0f - zg, hh, zg * cst[s][u]);
                        cst[s][u] = hv;
                        hnew[s] = hv;
                    }
                }
                sts_vec<S>(hn + (u * L + l) * HP, hnew);
            }
            __syncwarp();

            // ---- dense + skip + output gain (applyModel, rt-neural-generic.cpp:170-182), lane l <-> stream l
            if (xsrc && l < S) {
                float y0 = db, y1 = 0.f;
#pragma unroll 8
                for (int k = 0; k < H; k += 2) {
                    y0 = fmaf(dw[k], hn[k * HP + l], y0);
                    y1 = fmaf(dw[k + 1], hn[(k + 1) * HP + l], y1);
                }
                const float y = y0 + y1;
                if ((commit_mask >> l) & 1u) {
                    const float xg = xsrc[t * XP + l] * in_gain;
                    xsrc[t * XP + l] = (in_skip ? xg + y : y) * out_gain;
                }
            }
            cur ^= 1;
        }
        __syncwarp();

        // ---- write the state back
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (!((commit_mask >> s) & 1u)) continue;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int unit = u * L + l;
                if (CELL == 0) {
                    sa.rnn[(size_t)(kMaxHidden + unit) * cap + slot[s]] = cst[s][u];
                    sa.rnn[(size_t)unit * cap + slot[s]] = hbuf[cur * (H * HP) + unit * HP + g * S + s];
                } else {
                    sa.rnn[(size_t)unit * cap + slot[s]] = cst[s][u];
                }
            }
        }
        __syncwarp();
    }
};

// ------------------------------------------------------------------------------------------
// the fused chain kernel.  CELL: 0 LSTM, 1 GRU, 2 no model (filter/gain chain only)
// ------------------------------------------------------------------------------------------
template <int CELL, int H, int L, int S, int NW, int MATH>
__global__ void __launch_bounds__(NW * 32) chain_kernel(const __grid_constant__ RenderParams p)
{
    using T = Tile<CELL == 2 ? 0 : CELL, H, L, S>;
    constexpr int WS = T::WS, XP = T::XP, HP = T::HP;
    constexpr int CS = NW * WS;
    extern __shared__ __align__(16) float smem[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x;
    const StreamArrays& sa = p.sa;

    ModelDev md{};
    int input_size = 1;
    constexpr int WFLOATS = CELL == 2 ? 0 : ((H + 3) * H * 4 + H * 4 + H + 4);
    float* wsm = smem;
    if constexpr (CELL != 2) {
        md = p.models[p.blk_model[blk]];
        input_size = md.input_size;
        const float4* src = reinterpret_cast<const float4*>(md.weights);
        float4* dst = reinterpret_cast<float4*>(wsm);
        for (int i = threadIdx.x; i < WFLOATS / 4; i += NW * 32) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const int n_ptiles = input_size - 1;
    float* wbase = smem + WFLOATS + (size_t)warp * T::warp_smem_floats(p.ts, n_ptiles);
    int* wslots = reinterpret_cast<int*>(wbase);
    float* xbuf = wbase + 32;
    float* p1buf = n_ptiles >= 1 ? xbuf + p.ts * XP : nullptr;
    float* p2buf = n_ptiles >= 2 ? p1buf + p.ts * XP : nullptr;
    float* hbuf = xbuf + (size_t)p.ts * XP * (1 + n_ptiles);

    if (lane < WS) wslots[lane] = p.blk_slots[(size_t)blk * CS + warp * WS + lane];
    __syncwarp();

    // chain role: lane <-> stream `lane` of the warp
    const int cslot = lane < WS ? wslots[lane] : -1;
    // tile role
    const int l = lane % L, g = lane / L;
    int tslot[S];
    uint32_t valid_mask = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        tslot[s] = wslots[g * S + s];
        if (tslot[s] >= 0) valid_mask |= 1u << s;
        else tslot[s] = 0;
    }
    if (__all_sync(0xffffffffu, cslot < 0)) return; // whole warp empty (no slot in the warp's table)

    // ---- commands (open / activate / model swap) and the pre-warm they may request
    uint32_t cmd = 0;
    if (cslot >= 0) cmd = exec_commands(sa, cslot);
    if constexpr (CELL != 2) {
        const uint32_t prewarm_ballot = __ballot_sync(0xffffffffu, (cmd & C_SET_MODEL) != 0);
        __syncwarp();
        if (prewarm_ballot) {
            // loadModelFromPath: reset() then applyModel over 2048 zero samples with the param
            // smoothers parked at the old targets (rt-neural-generic.cpp:1035, 1052-1079)
            const uint32_t gm = (prewarm_ballot >> (g * S)) & ((1u << S) - 1u) & valid_mask;
            float p1c[S], p2c[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                p1c[s] = sa.p_mem[tslot[s]];
                p2c[s] = sa.p_mem[(size_t)sa.cap + tslot[s]];
            }
            T::template run<MATH>(wsm, hbuf, nullptr, nullptr, nullptr, p1c, p2c, kPrewarmSteps, input_size, md.in_gain,
                                  md.out_gain, md.in_skip != 0, sa, tslot, 0u, gm, l, g);
        }
    }

    uint32_t flags = 0, dfl = 0;
    if (cslot >= 0) { flags = sa.flags[cslot]; dfl = sa.dflags[cslot]; }
    const bool enabled = (flags & F_ENABLED) != 0;
    const bool net_on = CELL != 2 && enabled && (flags & F_NET);
    const uint32_t net_ballot = __ballot_sync(0xffffffffu, net_on);
    const uint32_t gnet = (net_ballot >> (g * S)) & ((1u << S) - 1u) & valid_mask;
    bool first_run = (dfl & D_PARAM_FIRST_RUN) != 0;

    for (int t0 = 0; t0 < p.n_samples; t0 += p.ts) {
        const int ts = min(p.ts, p.n_samples - t0);
        load_tile<WS, XP>(xbuf, p, wslots, t0, ts, lane);
        __syncwarp();

        // ---- stages before the network
        if (cslot >= 0 && enabled) {
            float* xcol = xbuf + lane;
            pass_lpf_pregain<XP>(xcol, ts, sa, cslot, (flags & F_LPF) != 0);
            if (flags & F_EQ_PRE) pass_eq<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0);
            if (net_on) {
                // both smoothers get setTargetValue every run(); only the ones the model consumes advance
                pass_param<XP>(p1buf ? p1buf + lane : nullptr, ts, sa, 0, cslot, first_run, input_size >= 2);
                pass_param<XP>(p2buf ? p2buf + lane : nullptr, ts, sa, 1, cslot, first_run, input_size >= 3);
                if (first_run) { first_run = false; sa.dflags[cslot] = dfl & ~D_PARAM_FIRST_RUN; }
            }
        }
        __syncwarp();

        // ---- the network
        if constexpr (CELL != 2) {
            if (net_ballot) {
                float pc[S];
#pragma unroll
                for (int s = 0; s < S; ++s) pc[s] = 0.f;
                T::template run<MATH>(wsm, hbuf, xbuf + g * S, p1buf ? p1buf + g * S : nullptr, p2buf ? p2buf + g * S : nullptr,
                                      pc, pc, ts, input_size, md.in_gain, md.out_gain, md.in_skip != 0, sa, tslot,
                                      valid_mask, gnet, l, g);
            }
        }

        // ---- stages after the network
        if (cslot >= 0 && enabled) {
            float* xcol = xbuf + lane;
            const bool dc_on = (flags & F_DC) != 0;
            if (flags & F_EQ_POST) {
                if (dc_on) pass_biquad<XP>(xcol, ts, sa, BQ_DC, cslot);
                pass_eq<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0);
                pass_dc_master<XP>(xcol, ts, sa, cslot, false, true);
            } else {
                pass_dc_master<XP>(xcol, ts, sa, cslot, dc_on, true);
            }
        }
        __syncwarp();
        store_tile<WS, XP>(xbuf, p, wslots, t0, ts, lane);
        __syncwarp();
    }
}

} // namespace aidax
