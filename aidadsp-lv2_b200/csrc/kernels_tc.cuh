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

namespace aidax {

template <int H>
struct TcShape {
    static constexpr int N = 4 * H;              // accumulator columns per tile: 4 per hidden unit (see layout.h lanes)
    static constexpr int KX = H;                 // first input row: x, p1, p2, then the constant 1 that carries the bias
    static constexpr int KP = (H + 4 + 7) / 8 * 8; // K padded to the tf32 MMA depth (8)
    static constexpr int KSTEPS = KP / 8;
    static constexpr int NSPLIT = N > 256 ? 2 : 1; // one MMA covers at most N = 256
    static constexpr int NMMA = N / NSPLIT;
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
        return 2 * B_FLOATS + nt * 2 * A_FLOATS + (H + 4) /* dense w, b */ + 2 * nt /* mbarriers */ + 4 /* tmem base + pad */;
    }
    __host__ __device__ static constexpr int warp_floats(int ts, int n_param_tiles) { return 32 + ts * XP * (1 + n_param_tiles); }
};

// everything a thread needs to drive its tile's recurrence
struct TcTile {
    float* a_hi;         // this thread's row in the A operand: a_hi + chunk*512 is a float4 slot
    float* a_lo;
    uint32_t a_hi_addr;  // shared-space byte addresses of the tile's A (hi, lo) and the CTA's B (hi, lo)
    uint32_t a_lo_addr;
    uint32_t b_hi_addr;
    uint32_t b_lo_addr;
    uint32_t bar_addr;   // the tile's mbarrier
    uint32_t phase;      // its current parity
    uint32_t tmem_d;     // TMEM address of the tile's accumulator (lane 0)
    uint32_t tmem_ld;    // same, with this warp's lane quadrant
    uint32_t bar_id;     // named barrier of the tile's 128 threads
    const float* dw;     // dense weights [H] + bias
    bool leader;
};

template <int CELL, int H, int MATH>
struct TcCell {
    using S = TcShape<H>;

    __device__ __forceinline__ static void store_chunk(const TcTile& tt, int chunk, float v0, float v1, float v2, float v3)
    {
        float h0, h1, h2, h3, l0, l1, l2, l3;
        tc::split_tf32(v0, h0, l0);
        tc::split_tf32(v1, h1, l1);
        tc::split_tf32(v2, h2, l2);
        tc::split_tf32(v3, h3, l3);
        *reinterpret_cast<float4*>(tt.a_hi + chunk * 512) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4*>(tt.a_lo + chunk * 512) = make_float4(l0, l1, l2, l3);
    }

    __device__ __forceinline__ static void issue_mmas(const TcTile& tt)
    {
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, S::NMMA);
        constexpr uint32_t LBO_A = 128 * 16, LBO_B = S::N * 16, SBO = 128;
#pragma unroll
        for (int ks = 0; ks < S::KSTEPS; ++ks) {
            const uint32_t ao = ks * 2 * LBO_A, bo = ks * 2 * LBO_B;
#pragma unroll
            for (int nh = 0; nh < S::NSPLIT; ++nh) {
                const uint32_t d = tt.tmem_d + nh * S::NMMA, bn = bo + nh * S::NMMA * 16;
                // smallest terms first
                tc::mma_tf32(d, tc::make_smem_desc(tt.a_lo_addr + ao, LBO_A, SBO), tc::make_smem_desc(tt.b_hi_addr + bn, LBO_B, SBO), idesc, ks > 0);
                tc::mma_tf32(d, tc::make_smem_desc(tt.a_hi_addr + ao, LBO_A, SBO), tc::make_smem_desc(tt.b_lo_addr + bn, LBO_B, SBO), idesc, 1);
                tc::mma_tf32(d, tc::make_smem_desc(tt.a_hi_addr + ao, LBO_A, SBO), tc::make_smem_desc(tt.b_hi_addr + bn, LBO_B, SBO), idesc, 1);
            }
        }
        tc::mma_commit(tt.bar_addr);
    }

    // 1/a, 1/b, 1/c, 1/d from ONE MUFU.RCP (a..d in [1, 2^24+1]: no overflow, ~3 ulp): the kernel is bound by the
    // XU (MUFU) pipe, so reciprocals are batched four at a time and the extra multiplies go to the idle FMA pipe.
    __device__ __forceinline__ static void inv4(float a, float b, float c, float d, float& ia, float& ib, float& ic, float& id)
    {
        const float ab = a * b, cd = c * d;
        const float r = rcp_approx(ab * cd);
        const float rab = r * cd, rcd = r * ab;
        ia = rab * b; ib = rab * a; ic = rcd * d; id = rcd * c;
    }
    // 1 + 2^e, exponent clamped so that four of them can be multiplied without overflow
    // (sigmoid(-16.6) = 6e-8, 1 - tanh(8.3) = 1.2e-7: the clamp is below fp32 resolution of the results)
    __device__ __forceinline__ static float one_plus_exp2(float e) { return 1.0f + ex2_approx(fminf(e, 24.0f)); }
    // In fast-math mode the B operand's columns are pre-scaled so that the accumulator IS that exponent:
    // sigmoid(x) = 1/(1 + 2^(-x log2 e)), tanh(x) = 1 - 2/(1 + 2^(2x log2 e)).  Column order per unit: LSTM i,f,g,o;
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

    // gate nonlinearities + cell update of the 4 hidden units whose 16 accumulator columns are in `raw`
    __device__ __forceinline__ static void eval_chunk(const uint32_t (&raw)[16], int q, float (&cst)[H], float (&hn)[4])
    {
        constexpr float L2E = 1.4426950408889634f;
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __uint_as_float(raw[i]);
        if (MATH == 0 && CELL == 0) {
            // LSTM, RTN/lstm/lstm_xsimd.h:193-213, 271-283: c' = i*tanh(g) + f*c ; h' = o*tanh(c')
            float og[4], e4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = 4 * q + i;
                float ii, fi, gi;
                inv4(one_plus_exp2(a[4 * i]), one_plus_exp2(a[4 * i + 1]), one_plus_exp2(a[4 * i + 2]), one_plus_exp2(a[4 * i + 3]),
                     ii, fi, gi, og[i]);
                const float c = fmaf(ii, fmaf(-2.f, gi, 1.f), fi * cst[u]);
                cst[u] = c;
                e4[i] = one_plus_exp2(c * (2.f * L2E));
            }
            float t0, t1, t2, t3;
            inv4(e4[0], e4[1], e4[2], e4[3], t0, t1, t2, t3);
            hn[0] = og[0] * fmaf(-2.f, t0, 1.f);
            hn[1] = og[1] * fmaf(-2.f, t1, 1.f);
            hn[2] = og[2] * fmaf(-2.f, t2, 1.f);
            hn[3] = og[3] * fmaf(-2.f, t3, 1.f);
        } else if (MATH == 0) {
            // GRU, RTN/gru/gru_xsimd.h:219-239, 265-271: h' = (1-z)*tanh(r*(Uh.h+bh1) + Wh.x+bh0) + z*h
            float z[4], e4[4];
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
                float r0, r1;
                inv4(one_plus_exp2(a[4 * i]), one_plus_exp2(a[4 * i + 1]), one_plus_exp2(a[4 * i + 4]), one_plus_exp2(a[4 * i + 5]),
                     z[i], r0, z[i + 1], r1);
                e4[i] = one_plus_exp2(fmaf(r0, a[4 * i + 2], a[4 * i + 3]));
                e4[i + 1] = one_plus_exp2(fmaf(r1, a[4 * i + 6], a[4 * i + 7]));
            }
            float t4[4];
            inv4(e4[0], e4[1], e4[2], e4[3], t4[0], t4[1], t4[2], t4[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = 4 * q + i;
                const float hh = fmaf(-2.f, t4[i], 1.f);
                hn[i] = fmaf(1.0f - z[i], hh, z[i] * cst[u]);
                cst[u] = hn[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = 4 * q + i;
                if (CELL == 0) {
                    const float ig = sigmoid_f<1>(a[4 * i]), fg = sigmoid_f<1>(a[4 * i + 1]);
                    const float gg = tanh_f<1>(a[4 * i + 2]), og = sigmoid_f<1>(a[4 * i + 3]);
                    const float c = fmaf(ig, gg, fg * cst[u]);
                    cst[u] = c;
                    hn[i] = og * tanh_f<1>(c);
                } else {
                    const float zg = sigmoid_f<1>(a[4 * i]), rg = sigmoid_f<1>(a[4 * i + 1]);
                    const float hh = tanh_f<1>(fmaf(rg, a[4 * i + 2], a[4 * i + 3]));
                    hn[i] = fmaf(1.0f - zg, hh, zg * cst[u]);
                    cst[u] = hn[i];
                }
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
            if (tt.leader) {
                tc::tc_fence_after();
                issue_mmas(tt);
            }
            tc::mbar_wait(tt.bar_addr, tt.phase);
            tt.phase ^= 1u;
            tc::tc_fence_after();

            // epilogue: groups of two 16-column chunks (8 hidden units) are loaded together and evaluated as one
            // straight-line block, so the 8 + 2 batched reciprocal chains of a group overlap each other's MUFU latency
            float y0 = db, y1 = 0.f;
#pragma unroll
            for (int q = 0; q < H / 4; q += 2) {
                constexpr int NQ = H / 4;
                const bool two = q + 1 < NQ;
                uint32_t raw0[16], raw1[16];
                tc::tmem_ld16(tt.tmem_ld + 16 * q, raw0);
                if (two) tc::tmem_ld16(tt.tmem_ld + 16 * (q + 1), raw1);
                tc::tmem_ld_wait16(raw0);
                float hn0[4], hn1[4];
                eval_chunk(raw0, q, cst, hn0);
                if (two) {
                    tc::tmem_ld_wait16(raw1); // already complete (one wait covers all outstanding loads); keeps the asm ordering explicit
                    eval_chunk(raw1, q + 1, cst, hn1);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    y0 = fmaf(tt.dw[4 * q + i], hn0[i], y0);
                    if (two) y1 = fmaf(tt.dw[4 * q + 4 + i], hn1[i], y1);
                }
                store_chunk(tt, q, hn0[0], hn0[1], hn0[2], hn0[3]);
                if (two) store_chunk(tt, q + 1, hn1[0], hn1[1], hn1[2], hn1[3]);
            }
            // dense + skip + output gain (applyModel, rt-neural-generic.cpp:170-182)
            if (xcol && commit) {
                const float y = y0 + y1;
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

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = warp >> 2, r = tid & 127;
    const int blk = blockIdx.x;
    const StreamArrays& sa = p.sa;

    float* Bh = smem;
    float* Bl = Bh + S::B_FLOATS;
    float* Abase = Bl + S::B_FLOATS;
    float* dw = Abase + NT * 2 * S::A_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(dw + H + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NT);

    const ModelDev md = p.models[p.blk_model[blk]];
    const int input_size = md.input_size;
    const int n_ptiles = input_size - 1;

    // ---- B operand (hi, lo) from the packed weight block (layout.h): B[n = 4*unit + lane][k]
    {
        const float* wk = md.weights;
        const float* bias = wk + (H + 3) * H * 4;
        for (int idx = tid; idx < S::N * S::KP; idx += NTHREADS) {
            const int n = idx % S::N, k = idx / S::N;
            float w = 0.f;
            if (k < H + 3) w = __ldg(wk + (size_t)k * S::N + n);
            else if (k == H + 3) w = __ldg(bias + n);
            if (MATH == 0) w *= Cell::column_scale(n & 3); // the accumulator is directly the exponent of 2 the gate needs
            float hi, lo;
            tc::split_tf32(w, hi, lo);
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
        for (int i = 0; i < NT; ++i) tc::mbar_init(&bars[i], 1);
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
    tt.a_hi_addr = tc::smem_u32(Abase + tile * 2 * S::A_FLOATS);
    tt.a_lo_addr = tt.a_hi_addr + S::A_FLOATS * 4;
    tt.b_hi_addr = tc::smem_u32(Bh);
    tt.b_lo_addr = tc::smem_u32(Bl);
    tt.bar_addr = tc::smem_u32(&bars[tile]);
    tt.phase = 0;
    tt.tmem_d = tmem_base + tile * S::N;
    tt.tmem_ld = tt.tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    tt.bar_id = 1 + tile;
    tt.dw = dw;
    tt.leader = r == 0;

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
        __syncwarp();

        // ---- stages before the network (run(), rt-neural-generic.cpp:622-644)
        if (cslot >= 0 && enabled) {
            float* xcol = xbuf + lane;
            pass_lpf_pregain<XP>(xcol, ts, sa, cslot, (flags & F_LPF) != 0);
            if (flags & F_EQ_PRE) pass_eq<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0);
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
        store_tile<32, XP>(xbuf, p, wslots, t0, ts, lane);
        __syncwarp();
    }

    if (net_on) Cell::store_state(tt, cst, sa, slot);

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, S::tmem_cols(NT));
}

} // namespace aidax
