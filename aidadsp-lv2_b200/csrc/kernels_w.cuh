// kernels_w.cuh -- warp-per-stream latency shape of the fused chain kernel, for hidden sizes up to 16 and small
// batches (configs[1]'s literal 1024 streams, a lone plugin instance): ONE stream per warp, so a period's latency is
// one warp's dependent instruction chain and nothing else.  Both serial parts of the chain are spread over the lanes:
//
//   * the recurrence: hidden size padded to 16, the 64 gate rows are dealt unit-major two per lane -- lanes (2u, 2u+1)
//     own unit u: LSTM (i, f | g, o), GRU (z, r | Uh.h+bh1, Wh.x+bh0).  A lane keeps its two weight rows in REGISTERS
//     for the whole period (2 x 20 floats), reads h(t-1) as four broadcast 128-bit shared loads, runs two short FFMA
//     chains, evaluates its two gate nonlinearities, and the two lanes of a unit finish the cell update with two warp
//     shuffles.  The Dense layer is a shuffle butterfly over the 16 owner lanes.  h goes through a double-buffered
//     16-float shared tile, one __syncwarp per step.
//   * the filter/gain stages: a biquad cascade is a chain of serial recurrences, but the stages form a pipeline, so
//     lane j runs stage j on sample t-j and hands its output to lane j+1 with a shuffle (systolic array over the
//     lanes): the pre-network stages (in_lpf + pre-gain | 5 EQ bands) are ONE 6-deep pass, the post-network stages
//     (dc blocker | 5 EQ bands | master gain) ONE 7-deep pass -- one biquad's worth of instructions per sample instead
//     of seven.  Every filter still sees exactly its own input sequence with the reference's operation order, so the
//     stages stay bit-identical to kernels.cuh (and to the reference's SSE2 build).
//
// Reference lines: run() rt-neural-generic.cpp:621-655, applyModel :148-240, LSTM RTN/lstm/lstm_xsimd.h:193-283,
// GRU RTN/gru/gru_xsimd.h:219-271, Dense RTN/dense/dense_xsimd.h:253-260.
#pragma once

#include "kernels.cuh"

namespace aidax {

struct WShape {
    static constexpr int HP = 16;                 // padded hidden size
    static constexpr int KW = HP + kMaxInput + 1; // weights per gate row kept in registers: U row, W rows, bias
    __host__ __device__ static constexpr int warp_floats(int ts, int n_param_tiles)
    {
        return 32 /* slot table + pad */ + 2 * HP /* h exchange */ + ((ts + 3) & ~3) * (1 + n_param_tiles);
    }
};

// One pass of a systolic biquad cascade over x[0..ts): lane j < NST runs stage j on sample i - 2j in iteration i.
//   f / bq_on : this lane's biquad (already loaded) and whether it is active (inactive = pass-through)
//   gain_on   : this lane multiplies its output by the exponential smoother (mem, coef, tgt) -- the pre-gain ramp of
//               stage 0 in the pre pass, the master ramp of the last stage in the post pass
// Neighbouring stages are skewed by TWO iterations: a stage's output is shuffled up in the iteration after it was
// produced and consumed in the one after that, so the shuffle and the float<->double conversions of one sample overlap
// the arithmetic of the next and an iteration costs about one biquad state recurrence, not the whole
// shuffle -> convert -> multiply-add -> convert chain.  The bulk of the pass (every stage in range) runs without
// range checks; the pipeline fill and drain use the checked body.
template <bool CHECKED, int NST>
__device__ __forceinline__ void systolic_step(float* x, int ts, int i, int& t, unsigned tsu, int lane, BiquadReg& f, bool bq_on, bool gain_on,
                                              float& mem, float coef, float tgt, float& out, float& recv)
{
    const float tmp = __shfl_up_sync(0xffffffffu, out, 1);
    const float vin = lane == 0 ? x[CHECKED ? min(i, ts - 1) : i] : recv;
    // Biquad::process (common/Biquad.h:53-58), as BiquadReg::process
    const double xd = (double)vin;
    const double o = __dadd_rn(__dmul_rn(xd, f.a0), f.z1);
    const double z1n = __dsub_rn(__dadd_rn(__dmul_rn(xd, f.a1), f.z2), __dmul_rn(f.b1, o));
    const double z2n = __dsub_rn(__dmul_rn(xd, f.a2), __dmul_rn(f.b2, o));
    float v = bq_on ? __double2float_rn(o) : vin;
    // ExponentialValueSmoother::next + applyGainRamp (rt-neural-generic.cpp:34-38)
    const float g = __fadd_rn(__fmul_rn(mem, coef), tgt);
    v = gain_on ? __fmul_rn(v, g) : v;
    if (CHECKED) {
        const bool act = (unsigned)t < tsu;
        f.z1 = act ? z1n : f.z1;
        f.z2 = act ? z2n : f.z2;
        mem = (act && gain_on) ? g : mem;
        out = act ? v : out;
        if (act && lane == NST - 1) x[t] = v;
    } else {
        // stages that are off (or lanes beyond the cascade) carry state nobody stores
        f.z1 = z1n;
        f.z2 = z2n;
        mem = gain_on ? g : mem;
        out = v;
        if (lane == NST - 1) x[t] = v;
    }
    recv = tmp;
    ++t;
}

template <int NST>
__device__ __forceinline__ void systolic_pass(float* x, int ts, int lane, BiquadReg& f, bool bq_on, bool gain_on, float& mem, float coef,
                                              float tgt)
{
    constexpr int LAG = 2 * (NST - 1);
    const unsigned tsu = lane < NST ? (unsigned)ts : 0u;
    float out = 0.f, recv = 0.f;
    int t = -2 * lane;
    int i = 0;
    for (; i < min(LAG, ts + LAG); ++i) systolic_step<true, NST>(x, ts, i, t, tsu, lane, f, bq_on, gain_on, mem, coef, tgt, out, recv);
#pragma unroll 2
    for (; i < ts; ++i) systolic_step<false, NST>(x, ts, i, t, tsu, lane, f, bq_on, gain_on, mem, coef, tgt, out, recv);
    for (; i < ts + LAG; ++i) systolic_step<true, NST>(x, ts, i, t, tsu, lane, f, bq_on, gain_on, mem, coef, tgt, out, recv);
}

template <int CELL, int H, int MATH>
struct TileW {
    static constexpr int HP = WShape::HP, KW = WShape::KW;
    static_assert(H <= HP && H % 4 == 0, "warp-per-stream shape: hidden size up to 16");

    // this lane's two gate rows from the packed weight block (layout.h): rows 2*half, 2*half+1 of unit u
    __device__ __forceinline__ static void load_weights(const float* __restrict__ wpk, int lane, float (&w0)[KW], float (&w1)[KW], float& dwu,
                                                        float& db)
    {
        const int u = lane >> 1, g0 = 2 * (lane & 1);
        const float* bias = wpk + (H + 3) * H * 4;
#pragma unroll
        for (int k = 0; k < KW; ++k) { w0[k] = 0.f; w1[k] = 0.f; }
        dwu = 0.f;
        db = __ldg(bias + H * 4 + H);
        if (u < H) {
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(wpk + ((size_t)k * H + u) * 4 + g0));
                w0[k] = w.x; w1[k] = w.y;
            }
#pragma unroll
            for (int i = 0; i < kMaxInput; ++i) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(wpk + ((size_t)(H + i) * H + u) * 4 + g0));
                w0[HP + i] = w.x; w1[HP + i] = w.y;
            }
            const float2 b = __ldg(reinterpret_cast<const float2*>(bias + u * 4 + g0));
            w0[HP + kMaxInput] = b.x; w1[HP + kMaxInput] = b.y;
            dwu = (lane & 1) ? 0.f : __ldg(bias + H * 4 + u);
        }
    }

    // n steps of the cell.  x: the stream's audio tile (nullptr = zero input, pre-warm); p1/p2: param tiles or nullptr
    // (then p1c/p2c); load: read h/c from the state arrays (else start from zero); commit: write y and the state back.
    __device__ static void run(const float (&w0)[KW], const float (&w1)[KW], float dwu, float db, float* hbuf, float* x, const float* p1,
                               const float* p2, float p1c, float p2c, int n, int input_size, float in_gain, float out_gain, bool in_skip,
                               const StreamArrays& sa, int slot, bool load, bool commit, int lane)
    {
        const int u = lane >> 1;
        const bool half = (lane & 1) != 0, owner = !half && u < H;
        const size_t cap = (size_t)sa.cap;
        float cst = 0.f; // LSTM: c, GRU: h of unit u (owner lanes)
        {
            float hv = 0.f, cv = 0.f;
            if (owner && load) {
                hv = sa.rnn[(size_t)u * cap + slot];
                if (CELL == 0) cv = sa.rnn[(size_t)(kMaxHidden + u) * cap + slot];
            }
            cst = CELL == 0 ? cv : hv;
            if (!half) hbuf[u] = hv;
        }
        __syncwarp();

        int cur = 0;
        float hlast = 0.f, yprev = 0.f, xgprev = 0.f;
        const float* xs = x ? x : hbuf; const int xstep = x ? 1 : 0;
        const float* p1s = p1 ? p1 : hbuf; const int p1step = p1 ? 1 : 0;
        const float* p2s = p2 ? p2 : hbuf; const int p2step = p2 ? 1 : 0;
        for (int t = 0; t < n; ++t) {
            const float* hc = hbuf + cur * HP;
            float hv[HP];
            lds_vec<HP>(hv, hc);
            // branch-free operand fetch (the time loop is one basic block): absent tiles read a valid dummy address
            const float xr = xs[t * xstep], p1r = p1s[t * p1step], p2r = p2s[t * p2step];
            const float xg = x ? xr * in_gain : 0.f;
            const float v1 = p1 ? p1r : p1c, v2 = p2 ? p2r : p2c;

            // two gate rows, two partial sums each (four independent FFMA chains); the weight rows of inputs the
            // model does not have are zero in the packed block
            float a0 = w0[HP + kMaxInput], a1 = w1[HP + kMaxInput];
            float b0 = w0[HP] * xg, b1 = w1[HP] * xg;
            a0 = fmaf(w0[HP + 1], v1, a0); a1 = fmaf(w1[HP + 1], v1, a1);
            b0 = fmaf(w0[HP + 2], v2, b0); b1 = fmaf(w1[HP + 2], v2, b1);
#pragma unroll
            for (int k = 0; k < H; k += 2) {
                a0 = fmaf(w0[k], hv[k], a0);
                a1 = fmaf(w1[k], hv[k], a1);
                b0 = fmaf(w0[k + 1], hv[k + 1], b0);
                b1 = fmaf(w1[k + 1], hv[k + 1], b1);
            }
            a0 += b0;
            a1 += b1;

            // Dense of the PREVIOUS step: a shuffle butterfly over the owner lanes (odd lanes and padding units carry
            // 0).  Deferred by one step and kept in the same straight-line block as the FFMA chains above, so its four
            // dependent shuffles are scheduled into their shadow instead of sitting on the recurrence's critical path.
            {
                float y = yprev;
                y += __shfl_xor_sync(0xffffffffu, y, 2);
                y += __shfl_xor_sync(0xffffffffu, y, 4);
                y += __shfl_xor_sync(0xffffffffu, y, 8);
                y += __shfl_xor_sync(0xffffffffu, y, 16);
                if (x && commit && lane == 0 && t > 0) x[t - 1] = (in_skip ? xgprev + (y + db) : (y + db)) * out_gain;
            }

            float hnew;
            if (CELL == 0) {
                // LSTM: even lane (i, f), odd lane (g, o);  c' = i*tanh(g) + f*c ; h' = o*tanh(c')
                const float e0 = half ? tanh_f<MATH>(a0) : sigmoid_f<MATH>(a0);
                const float e1 = sigmoid_f<MATH>(a1);
                const float tg = __shfl_xor_sync(0xffffffffu, e0, 1), og = __shfl_xor_sync(0xffffffffu, e1, 1);
                const float c = fmaf(e0, tg, e1 * cst);
                cst = c;
                hnew = og * tanh_f<MATH>(c);
            } else {
                // GRU: even lane (z, r), odd lane (Uh.h+bh1, Wh.x+bh0);  h' = (1-z)*tanh(r*a2 + a3) + z*h
                const float zg = sigmoid_f<MATH>(a0), rg = sigmoid_f<MATH>(a1);
                const float rx = __shfl_xor_sync(0xffffffffu, rg, 1);
                const float hh = tanh_f<MATH>(fmaf(rx, a0, a1));
                const float hx = __shfl_xor_sync(0xffffffffu, hh, 1);
                hnew = fmaf(1.0f - zg, hx, zg * cst);
                cst = hnew;
            }
            if (!owner) hnew = 0.f;
            hlast = hnew;
            if (!half) hbuf[(cur ^ 1) * HP + u] = hnew;
            yprev = dwu * hnew;
            xgprev = xg;
            __syncwarp();
            cur ^= 1;
        }
        if (n > 0) { // Dense of the last step
            float y = yprev;
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            y += __shfl_xor_sync(0xffffffffu, y, 4);
            y += __shfl_xor_sync(0xffffffffu, y, 8);
            y += __shfl_xor_sync(0xffffffffu, y, 16);
            if (x && commit && lane == 0) x[n - 1] = (in_skip ? xgprev + (y + db) : (y + db)) * out_gain;
        }
        __syncwarp();

        if (commit && owner) {
            if (CELL == 0) {
                sa.rnn[(size_t)(kMaxHidden + u) * cap + slot] = cst;
                sa.rnn[(size_t)u * cap + slot] = n > 0 ? hlast : hbuf[cur * HP + u];
            } else {
                sa.rnn[(size_t)u * cap + slot] = cst;
            }
        }
        __syncwarp();
    }
};

// ------------------------------------------------------------------------------------------
// the fused chain kernel, one stream per warp.  CELL: 0 LSTM, 1 GRU
// ------------------------------------------------------------------------------------------
template <int CELL, int H, int NW, int MATH>
__global__ void __launch_bounds__(NW * 32) chain_kernel_w(const __grid_constant__ RenderParams p)
{
    using T = TileW<CELL, H, MATH>;
    constexpr int HP = WShape::HP, KW = WShape::KW;
    extern __shared__ __align__(16) float smem_w[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x;
    const StreamArrays& sa = p.sa;
    const size_t cap = (size_t)sa.cap;

    const int cslot = p.blk_slots[(size_t)blk * NW + warp];
    if (cslot < 0) return; // warp-uniform: no stream for this warp

    const ModelDev md = p.models[p.blk_model[blk]];
    const int input_size = md.input_size;
    const int n_ptiles = input_size - 1;
    const int tsp = (p.ts + 3) & ~3;

    float* wbase = smem_w + (size_t)warp * WShape::warp_floats(p.ts, n_ptiles);
    int* wslots = reinterpret_cast<int*>(wbase);
    float* hbuf = wbase + 32;
    float* xbuf = hbuf + 2 * HP;
    float* p1buf = n_ptiles >= 1 ? xbuf + tsp : nullptr;
    float* p2buf = n_ptiles >= 2 ? p1buf + tsp : nullptr;
    if (lane == 0) wslots[0] = cslot;

    float w0[KW], w1[KW], dwu, db;
    T::load_weights(md.weights, lane, w0, w1, dwu, db);

    // ---- commands (open / activate / model swap) and the pre-warm they may request
    uint32_t cmd = 0;
    if (lane == 0) cmd = exec_commands(sa, cslot);
    cmd = __shfl_sync(0xffffffffu, cmd, 0);
    __syncwarp();
    if (cmd & C_SET_MODEL) {
        // loadModelFromPath: reset() then applyModel over 2048 zero samples with the param smoothers parked at the
        // old targets (rt-neural-generic.cpp:1035, 1052-1079)
        const float p1c = sa.p_mem[cslot], p2c = sa.p_mem[cap + cslot];
        T::run(w0, w1, dwu, db, hbuf, nullptr, nullptr, nullptr, p1c, p2c, kPrewarmSteps, input_size, md.in_gain, md.out_gain, md.in_skip != 0,
               sa, cslot, false, true, lane);
    }

    const uint32_t flags = sa.flags[cslot], dfl = sa.dflags[cslot];
    const bool enabled = (flags & F_ENABLED) != 0;
    const bool net_on = enabled && (flags & F_NET);
    bool first_run = (dfl & D_PARAM_FIRST_RUN) != 0;
    const bool eq_pre = (flags & F_EQ_PRE) != 0, eq_post = (flags & F_EQ_POST) != 0, bandpass = (flags & F_EQ_BANDPASS) != 0;
    const float coef = sa.exp_coef[cslot];
    const float one_minus = __fsub_rn(1.f, coef);

    for (int t0 = 0; t0 < p.n_samples; t0 += p.ts) {
        const int ts = min(p.ts, p.n_samples - t0);
        load_tile<1, 1>(xbuf, p, wslots, t0, ts, lane);
        __syncwarp();

        if (enabled) {
            // ---- stages before the network (run(), :622-630): lane 0 in_lpf + pre-gain ramp, lanes 1..5 the EQ bands
            {
                const bool is_eq = lane >= 1 && lane <= 5;
                const bool bq_on = lane == 0 ? (flags & F_LPF) != 0 : (is_eq && eq_pre && (!bandpass || lane == BQ_MID));
                BiquadReg f{};
                if (bq_on) f.load(sa, lane, cslot); // stage j <-> biquad index j (BQ_INLPF, BQ_DEPTH .. BQ_PRESENCE)
                float mem = 0.f, tgt = 0.f;
                if (lane == 0) { mem = sa.pre_mem[cslot]; tgt = __fmul_rn(sa.pre_target[cslot], one_minus); }
                systolic_pass<6>(xbuf, ts, lane, f, bq_on, lane == 0, mem, coef, tgt);
                if (bq_on) f.store(sa, lane, cslot);
                if (lane == 0) sa.pre_mem[cslot] = mem;
            }
            if (net_on) {
                // both smoothers get setTargetValue every run(); only the ones the model consumes advance (:633-641)
                if (lane < 2) pass_param<1>(lane == 0 ? p1buf : p2buf, ts, sa, lane, cslot, first_run, input_size >= 2 + lane);
                if (first_run) {
                    first_run = false;
                    if (lane == 0) sa.dflags[cslot] = dfl & ~D_PARAM_FIRST_RUN;
                }
            }
        }
        __syncwarp();

        // ---- the network
        if (net_on) T::run(w0, w1, dwu, db, hbuf, xbuf, p1buf, p2buf, 0.f, 0.f, ts, input_size, md.in_gain, md.out_gain, md.in_skip != 0, sa, cslot, true, true, lane);

        // ---- stages after the network (:645-655): lane 0 dc blocker, lanes 1..5 the EQ bands, lane 6 the master ramp
        if (enabled) {
            const bool is_eq = lane >= 1 && lane <= 5;
            const bool bq_on = lane == 0 ? (flags & F_DC) != 0 : (is_eq && eq_post && (!bandpass || lane == BQ_MID));
            BiquadReg f{};
            if (bq_on) f.load(sa, lane == 0 ? BQ_DC : lane, cslot);
            float mem = 0.f, tgt = 0.f;
            if (lane == 6) { mem = sa.master_mem[cslot]; tgt = __fmul_rn(sa.master_target[cslot], one_minus); }
            systolic_pass<7>(xbuf, ts, lane, f, bq_on, lane == 6, mem, coef, tgt);
            if (bq_on) f.store(sa, lane == 0 ? BQ_DC : lane, cslot);
            if (lane == 6) sa.master_mem[cslot] = mem;
        }
        __syncwarp();
        store_tile<1, 1>(xbuf, p, wslots, t0, ts, lane);
        __syncwarp();
    }
}

} // namespace aidax
