// kernels_tc3.cuh -- third-generation tensor-core shape of the fused chain kernel: WARP-SPECIALISED.
//
// What generations 1 and 2 measured: inside the recurrence the XU pipe (MUFU ex2/rcp) is ~90 % busy, but
// 38 % of every warp's time goes to the filter/gain stages of run() -- serial double-precision recurrences at low ILP,
// global-memory round trips for tiles, coefficients and filter state -- during which the XU pipe idles.  With every warp
// alternating between the two phases (4 warps per scheduler, 128 registers each) the pipe that bounds the kernel is busy
// 61 % of the time.  Here the two roles get their own warps and their own register budgets:
//
//   * recurrence warps (R): NTILES tiles of 128 streams per CTA (1, 2 or 3), SPLIT warps per TMEM lane quadrant.  They do nothing but
//     Tc2Cell::run -- wait for the MMA, gate epilogue, write h(t) and x(t+1) into the A operand in TENSOR MEMORY, hand
//     the tile back to the tensor core -- for the whole period, sub-block after sub-block.
//   * filter warps (F): NFW warps per CTA in whole warpgroups -- 8 in the shipped shapes, one per warp-tile of 32 streams
//     (lane l owns stream 32 wt + l; with NFW = 4 a lane owns two streams and runs them one after the other).  An F warp
//     stages its streams' audio, runs the stages in front of the network for sub-block k+1 and the stages behind it for
//     sub-block k-1 while the R warps work on sub-block k.  Sub-blocks travel through a two-deep ring of shared-memory
//     tiles guarded by mbarriers (`full`: F -> R, `done`: R -> F), one pair per warp-tile and ring slot, so that no
//     block-wide barrier sits in the period and the R warps wait for a filter stage only while the pipeline fills.
//     (Measured on B200: with 4 F warps -- one per scheduler -- the R warps spent 13 % of their time waiting on `full`,
//     a lone F warp cannot hide its own dependency latency; 8 F warps cover for each other.)
//   * `setmaxnreg` moves registers from the F warpgroup (FREG) to the R warpgroups (RREG): the launch carries
//     65536 / (threads x CTAs per SM) registers per thread and the roles re-split them.
//
// Arithmetic is unchanged: the R role IS Tc2Cell (one tile per thread group, A operand in tensor memory), the F role
// runs the pass_* functions of kernels.cuh (the five EQ bands as five passes instead of the software pipeline: every
// filter still sees its own input sequence, float between stages), so the output is bit-identical to generation 1/2.
#pragma once

#include "kernels_tc2.cuh"

namespace aidax {

namespace tc {
__device__ __forceinline__ void mbar_arrive(uint32_t bar_addr)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// Wait of a warp that is NOT on the critical path (the filter warps waiting for a sub-block to come back from the recurrence
// warps): poll and sleep.  mbarrier.try_wait's suspended form (NANOSLEEP.SYNCS) wakes on every mbarrier event of the CTA --
// here two MMA completions per time step -- and each wake-up re-issues the check: ncu counted 26 % of the kernel's executed
// instructions in those retry loops, issue slots taken from the recurrence warps.
#ifndef AIDAX_TC3_POLL_NS
#define AIDAX_TC3_POLL_NS 1000
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar_addr, uint32_t parity)
{
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
        if (ok) break;
#if AIDAX_TC3_POLL_NS > 0
        __nanosleep(AIDAX_TC3_POLL_NS);
#endif
    }
}
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
} // namespace tc

template <int H, int SPLIT, int NFW = 8, int NTILES = 2>
struct Tc3Shape {
    using S = TcShape<H>;
    static constexpr int N = S::N, KP = S::KP, XP = S::XP;
    static constexpr int UH = H / SPLIT;
    static constexpr int GROUP = 128 * SPLIT;  // recurrence threads of one tile (meet at its named barrier)
    static constexpr int NR = NTILES * GROUP;  // recurrence threads
    static constexpr int NRW = NR / 32;
    static constexpr int NTH = NR + 32 * NFW;  // + the filter warps: whole warpgroups
    static constexpr int NWT = 4 * NTILES;     // warp-tiles of 32 streams
    static constexpr int FS = NWT / NFW;       // warp-tiles (streams per lane) of a filter warp
    static_assert(NFW % 4 == 0 && NWT % NFW == 0, "filter warps come in whole warpgroups and share the warp-tiles evenly");
    static constexpr int NW = NTH / 32;
    static constexpr int CS = 128 * NTILES;    // streams per CTA: NTILES tiles of 128
    static constexpr int NBUF = 2;             // ring depth of the audio sub-block tiles
    static constexpr int YB = SPLIT > 1 ? SPLIT : 0;
    static constexpr int TILE_COLS = N + 2 * KP; // TMEM columns of one tile: accumulator | A hi | A lo
    static constexpr int NBARS = NTILES + 2 * NWT * NBUF + 3; // MMA completion per tile, full / done per (warp-tile, ring slot), 3 CTA-wide rendezvous
    static_assert(H % SPLIT == 0 && UH % 4 == 0, "the units of a thread must be whole 16-column accumulator chunks");
    static_assert(NTILES * TILE_COLS <= 512, "the tiles must fit tensor memory");

    __host__ __device__ static constexpr int tmem_cols()
    {
        int c = 32;
        while (c < NTILES * TILE_COLS) c *= 2;
        return c;
    }
    __host__ __device__ static constexpr int fixed_floats()
    {
        return 2 * S::B_FLOATS + (H + 4) + 2 * NBARS + 8 /* tmem base, CTA flag word, pad */ + 2 * CS /* slot table, per-stream flags */;
    }
    // the ring of audio tiles (NWT warp-tiles of 32 streams per slot), expressed per launch warp as the engine sizes it
    __host__ __device__ static constexpr int warp_floats(int ts, int n_param_tiles) { return (NBUF * NWT * ts * XP * (1 + n_param_tiles + YB) + NW - 1) / NW; }
};

// Sub-block schedule of a period: the FIRST sub-block is short (AIDAX_TC3_TS0 samples) so that the recurrence warps of a new
// CTA start after a short fill of the filter pipeline; the others have p.ts samples.
#ifndef AIDAX_TC3_TS0
#define AIDAX_TC3_TS0 8 // measured on B200 (LSTM-16, 378880 streams): 0: 1721, 2: 1715, 4: 1707, 8: 1695 us per period
#endif
struct SubBlocks {
    int n, ts, ts0, count;
    __device__ __forceinline__ SubBlocks(int n_samples, int ts_)
        : n(n_samples), ts(ts_), ts0((AIDAX_TC3_TS0 > 0 && AIDAX_TC3_TS0 < ts_ && n_samples > AIDAX_TC3_TS0) ? AIDAX_TC3_TS0 : ts_)
    {
        count = n <= ts0 ? 1 : 1 + (n - ts0 + ts - 1) / ts;
    }
    __device__ __forceinline__ int start(int k) const { return k == 0 ? 0 : ts0 + (k - 1) * ts; }
    __device__ __forceinline__ int len(int k) const { return k == 0 ? min(ts0, n) : min(ts, n - start(k)); }
};

enum : int { CF_PREWARM = 1, CF_NET = 2 }; // CTA-wide flag word written by the filter warps before the first barrier

// the five EQ bands as five passes over the tile (filter warps run with a small register budget; pass_eq's software
// pipeline keeps 5 x 7 doubles live).  Same values: stage k's input is the float stage k-1 left in the tile.
#ifndef AIDAX_TC3_WIDEN_INT
#define AIDAX_TC3_WIDEN_INT 0 // 1: the filter warps widen float -> double with integer instructions (BiquadReg::widen): 7 of the 114 XU lane-ops per stream-sample
#endif
#ifndef AIDAX_TC3_PREFETCH
#define AIDAX_TC3_PREFETCH 0 // 1: a filter pass asks for the NEXT pass's coefficients and state (L1) before it starts its own loop
#endif
__device__ __forceinline__ void prefetch_bq(const StreamArrays& sa, int which, int slot)
{
#if AIDAX_TC3_PREFETCH
    const size_t cap = (size_t)sa.cap;
    const double* c = sa.bq_coef + (size_t)which * 5 * cap + slot;
#pragma unroll
    for (int i = 0; i < 5; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + i * cap));
    const double* z = sa.bq_z + (size_t)which * 2 * cap + slot;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(z));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(z + cap));
#endif
}
template <int XP>
__device__ __forceinline__ void pass_eq_serial(float* xcol, int ts, const StreamArrays& sa, int slot, bool bandpass, int next_which)
{
    // (no calls in the filter role: a callee's registers are not bounded by setmaxnreg)
    const int k0 = bandpass ? BQ_MID : BQ_DEPTH, k1 = bandpass ? BQ_MID : BQ_PRESENCE;
#pragma unroll 1
    for (int k = k0; k <= k1; ++k) {
        prefetch_bq(sa, k < k1 ? k + 1 : next_which, slot);
        pass_biquad<XP, AIDAX_TC3_WIDEN_INT>(xcol, ts, sa, k, slot);
    }
}

template <int CELL, int H, int SPLIT, int NFW, int RREG, int FREG, int RES, int MATH, int NTILES = 2>
__global__ void __launch_bounds__(Tc3Shape<H, SPLIT, NFW, NTILES>::NTH, RES) chain_kernel_tc3(const __grid_constant__ RenderParams p)
{
    using Sh = Tc3Shape<H, SPLIT, NFW, NTILES>;
    using S = TcShape<H>;
    using Cell = Tc2Cell<CELL, H, SPLIT, true, false, MATH>;
    using Base = TcCell<CELL, H, MATH>;
    constexpr int XP = Sh::XP, NTH = Sh::NTH, CS = Sh::CS, UH = Sh::UH, NRW = Sh::NRW, NBUF = Sh::NBUF, NWT = Sh::NWT;
    // The roles meet from different code paths, so the CTA-wide rendezvous are mbarriers (every thread arrives, then waits for
    // phase 0), not bar.sync: each is passed once per launch.  Hand-overs through `full` / `done` are arrivals of EVERY lane of
    // the producing warp (each thread releases its own shared-memory writes; compute-sanitizer's racecheck follows that).
    extern __shared__ __align__(128) float smem_tc3[];
    float* smem = smem_tc3;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int blk = blockIdx.x;
    const StreamArrays& sa = p.sa;

    float* Bh = smem;
    float* Bl = Bh + S::B_FLOATS;
    float* dw = Bl + S::B_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(dw + H + 4); // [NTILES] MMA completion; then full[NWT][NBUF], done[NWT][NBUF], 3 rendezvous
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Sh::NBARS);
    int* cta_flags = reinterpret_cast<int*>(tmem_slot + 1);
    int* wslots = reinterpret_cast<int*>(tmem_slot + 8);
    int* sflags = wslots + CS;
    float* audio = smem + ((Sh::fixed_floats() + 3) & ~3);

    const ModelDev md = p.models[p.blk_model[blk]];
    const int input_size = md.input_size;
    const int n_ptiles = input_size - 1;

    // ---- B operand (hi, lo) from the packed weight block (layout.h): B[n = column_of(unit, gate)][k]
    {
        const float* wk = md.weights;
        const float* bias = wk + (H + 3) * H * 4;
        for (int idx = tid; idx < S::N * S::KP; idx += NTH) {
            const int n = idx % S::N, k = idx / S::N;
            const int g = Base::gate_of_column(n), src = 4 * Base::unit_of_column(n) + g;
            float w = 0.f;
            if (k < H + 3) w = __ldg(wk + (size_t)k * S::N + src);
            else if (k == H + 3) w = __ldg(bias + src);
            if (MATH == 0) w *= Base::column_scale(g);
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
        for (int i = tid; i < H + 1; i += NTH) dw[i] = __ldg(bias + H * 4 + i);
        for (int i = tid; i < CS; i += NTH) {
            wslots[i] = p.blk_slots[(size_t)blk * CS + i];
            sflags[i] = 0;
        }
    }
    if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < NTILES; ++i) tc::mbar_init(&bars[i], 1);
#pragma unroll 1
        for (int i = 0; i < NWT * NBUF; ++i) {
            tc::mbar_init(&bars[NTILES + i], 32);                       // full: the lanes of the warp-tile's filter warp
            tc::mbar_init(&bars[NTILES + NWT * NBUF + i], 32 * SPLIT);  // done: the lanes of every recurrence warp of the warp-tile
        }
#pragma unroll 1
        for (int i = 0; i < 3; ++i) tc::mbar_init(&bars[NTILES + 2 * NWT * NBUF + i], NTH);
        *cta_flags = 0;
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, Sh::tmem_cols());
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tile_floats = p.ts * XP * (1 + n_ptiles + Sh::YB);
    const uint32_t full0 = tc::smem_u32(&bars[NTILES]), done0 = tc::smem_u32(&bars[NTILES + NWT * NBUF]);
    const uint32_t meet0 = tc::smem_u32(&bars[NTILES + 2 * NWT * NBUF]);
    auto rendezvous = [&](int which) {
        tc::mbar_arrive(meet0 + 8 * which);
        tc::mbar_wait(meet0 + 8 * which, 0u);
    };
    const SubBlocks sb(p.n_samples, p.ts);
    const int nsb = sb.count;

    if (warp >= NRW) {
        // =====================================================================================================
        // filter warpgroup
        // =====================================================================================================
        tc::reg_dec<FREG>();
        const int fq = warp - NRW;
        // ---- commands (open / activate / model swap), per-stream flags for the recurrence warps
        {
            int cf = 0;
#pragma unroll 1
            for (int f = 0; f < Sh::FS; ++f) {
                const int si = (NFW * f + fq) * 32 + lane; // index in the CTA's slot table = tile * 128 + row
                const int cslot = wslots[si];
                int sf = 0;
                if (cslot >= 0) {
                    sf = SF_VALID;
                    const uint32_t cmd = exec_commands(sa, cslot);
                    if (cmd & C_SET_MODEL) { sf |= SF_PREWARM; cf |= CF_PREWARM; }
                    const uint32_t flags = sa.flags[cslot];
                    if ((flags & F_ENABLED) && (flags & F_NET)) { sf |= SF_NET; cf |= CF_NET; }
                }
                sflags[si] = sf;
            }
            cf = __reduce_or_sync(0xffffffffu, cf);
            if (lane == 0 && cf) atomicOr(cta_flags, cf);
        }
        __threadfence_block();
        rendezvous(0);                                    // (A) slot flags, CTA flags and the commands' effects are visible
        const int cf = *reinterpret_cast<volatile int*>(cta_flags);
        if (cf & CF_PREWARM) rendezvous(1);               // (A2) the recurrence warps have read the parked param values
        const bool any_net = (cf & CF_NET) != 0;

#pragma unroll 1
        for (int k = 0; k < nsb + NBUF; ++k) {
#pragma unroll 1
            for (int f = 0; f < Sh::FS; ++f) {
                const int wt = NFW * f + fq;
                const int* ws = wslots + wt * 32;
                const int cslot = ws[lane];
                const int b = k % NBUF;
                float* xb = audio + (size_t)(b * NWT + wt) * tile_floats;
                float* xcol = xb + lane;
                uint32_t flags = 0;
                if (cslot >= 0) flags = sa.flags[cslot];
                const bool live = cslot >= 0 && (flags & F_ENABLED) && !(AIDAX_WHATIF & 8);

                // ---- sub-block k - NBUF: stages after the network (run(), rt-neural-generic.cpp:645-655), write back
                if (k >= NBUF) {
                    const int j = k - NBUF, t0 = sb.start(j), ts = sb.len(j);
                    if (any_net) tc::mbar_wait_relaxed(done0 + 8 * (wt * NBUF + b), (uint32_t)(j / NBUF) & 1u);
                    if (live) {
                        if constexpr (SPLIT > 1) {
                            // Dense partial sums of the parts + skip + output gain (applyModel, :170-182)
                            if (any_net && (flags & F_NET)) {
                                const size_t ystep = (size_t)p.ts * XP;
                                const float* y0 = xb + (size_t)(1 + n_ptiles) * ystep + lane;
                                const bool skip = md.in_skip != 0;
                                for (int t = 0; t < ts; ++t) {
                                    float y = y0[t * XP];
#pragma unroll
                                    for (int h2 = 1; h2 < SPLIT; ++h2) y += y0[h2 * ystep + t * XP];
                                    const float xgv = xcol[t * XP] * md.in_gain;
                                    xcol[t * XP] = (skip ? xgv + y : y) * md.out_gain;
                                }
                            }
                        }
                        const bool dc_on = (flags & F_DC) != 0;
                        if (flags & F_EQ_POST) {
                            prefetch_bq(sa, (flags & F_EQ_BANDPASS) ? BQ_MID : BQ_DEPTH, cslot);
                            if (dc_on) pass_biquad<XP, AIDAX_TC3_WIDEN_INT>(xcol, ts, sa, BQ_DC, cslot);
                            pass_eq_serial<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0, BQ_INLPF);
                            pass_dc_master<XP, AIDAX_TC3_WIDEN_INT>(xcol, ts, sa, cslot, false, true);
                        } else {
                            prefetch_bq(sa, BQ_INLPF, cslot);
                            pass_dc_master<XP, AIDAX_TC3_WIDEN_INT>(xcol, ts, sa, cslot, dc_on, true);
                        }
                    }
                    __syncwarp();
                    store_tile<32, XP>(xb, p, ws, t0, ts, lane);
                    __syncwarp();
                }
                // ---- sub-block k: stage in, stages before the network (:622-644), hand over
                if (k < nsb) {
                    const int t0 = sb.start(k), ts = sb.len(k);
                    load_tile<32, XP>(xb, p, ws, t0, ts, lane);
                    __syncwarp();
                    float* p1b = n_ptiles >= 1 ? xb + p.ts * XP : nullptr;
                    float* p2b = n_ptiles >= 2 ? xb + 2 * p.ts * XP : nullptr;
                    if (live) {
                        if (flags & F_EQ_PRE) prefetch_bq(sa, (flags & F_EQ_BANDPASS) ? BQ_MID : BQ_DEPTH, cslot);
                        pass_lpf_pregain<XP, AIDAX_TC3_WIDEN_INT>(xcol, ts, sa, cslot, (flags & F_LPF) != 0);
                        if (flags & F_EQ_PRE) pass_eq_serial<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0, BQ_DC);
                        if (flags & F_NET) {
                            const uint32_t dfl = sa.dflags[cslot];
                            const bool first = (dfl & D_PARAM_FIRST_RUN) != 0;
                            pass_param<XP>(p1b ? p1b + lane : nullptr, ts, sa, 0, cslot, first, input_size >= 2);
                            pass_param<XP>(p2b ? p2b + lane : nullptr, ts, sa, 1, cslot, first, input_size >= 3);
                            if (first) sa.dflags[cslot] = dfl & ~D_PARAM_FIRST_RUN;
                        }
                    } else if (n_ptiles > 0) {
                        // the param columns of idle rows still feed the (discarded) product: keep them finite
                        for (int t = 0; t < ts; ++t) {
                            p1b[t * XP + lane] = 0.f;
                            if (p2b) p2b[t * XP + lane] = 0.f;
                        }
                    }
                    __syncwarp();
                    tc::mbar_arrive(full0 + 8 * (wt * NBUF + b));
                }
            }
        }
    } else {
        // =====================================================================================================
        // recurrence warps: tile g = warp / (4 SPLIT), lane quadrant q, part hs of the hidden units
        // =====================================================================================================
        tc::reg_inc<RREG>();
        const int g = warp / (4 * SPLIT), wq = warp % (4 * SPLIT);
        const int q = wq & 3, hs = wq >> 2, row = q * 32 + lane;
        const int wt = 4 * g + q;

        Tc2Tile tt[1];
        tc_desc_t db_hi, db_lo;
        {
            constexpr uint32_t LBO_B = S::N * 16, SBO = 128;
            db_hi = (tc_desc_t)tc::make_smem_desc(tc::smem_u32(Bh), LBO_B, SBO);
            db_lo = (tc_desc_t)tc::make_smem_desc(tc::smem_u32(Bl), LBO_B, SBO);
            const uint32_t tile_base = tmem_base + g * Sh::TILE_COLS;
            const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
            tt[0].tmem_d = tile_base;
            tt[0].tmem_ld = tile_base + lane_sel;
            tt[0].tma_hi = tile_base + S::N;
            tt[0].tma_lo = tile_base + S::N + S::KP;
            tt[0].ta_hi = tt[0].tma_hi + lane_sel;
            tt[0].ta_lo = tt[0].tma_lo + lane_sel;
            tt[0].a_hi = nullptr;
            tt[0].a_lo = nullptr;
            tt[0].da_hi = tt[0].da_lo = 0;
            tt[0].bar_addr = tc::smem_u32(&bars[g]);
            tt[0].phase = 0;
            tt[0].bar_id = 1 + g;
        }
        const bool issuer = wq == 0; // the first warp of a tile's thread group issues that tile's MMAs
        // tensor memory is not initialised: the K padding of this thread's A rows must be finite zeros
        if (hs == 0) {
#pragma unroll
            for (int c = H / 4; c < S::KP / 4; ++c) Cell::store_chunk(tt[0], c, 0.f, 0.f, 0.f, 0.f);
        }

        rendezvous(0);                                    // (A)
        const int cf = *reinterpret_cast<volatile int*>(cta_flags);
        const int sf = sflags[g * 128 + row];
        const int rslot = max(wslots[g * 128 + row], 0);
        const bool any_net = (cf & CF_NET) != 0;
        const bool rnet = (sf & SF_NET) != 0;

        if (cf & CF_PREWARM) {
            // loadModelFromPath: reset() then applyModel over 2048 zero samples with the param smoothers parked at the
            // old targets (rt-neural-generic.cpp:1035, 1052-1079)
            typename Cell::Lane ln[1];
            ln[0].xcol = nullptr; ln[0].p1col = nullptr; ln[0].p2col = nullptr; ln[0].ycol = nullptr;
            ln[0].p1c = sa.p_mem[rslot];
            ln[0].p2c = sa.p_mem[(size_t)sa.cap + rslot];
            ln[0].commit = false;
            rendezvous(1);                                // (A2) from here on the filter warps may advance the param smoothers
            float cst[1][UH];
            Cell::load_state(tt[0], cst[0], sa, rslot, false, hs * UH);
            Cell::run(tt, cst, ln, dw, kPrewarmSteps, md.in_gain, md.out_gain, md.in_skip != 0, hs, issuer, db_hi, db_lo);
            Cell::store_state(tt[0], cst[0], sa, rslot, hs * UH, (sf & SF_PREWARM) != 0);
            tc::bar_sync(8 + g, Sh::GROUP);               // every part's state is in global memory before anyone reloads it
        }

        float cst[1][UH]; // recurrent state of this thread's units, in registers across the whole period
        if (any_net) {
            Cell::load_state(tt[0], cst[0], sa, rslot, (sf & SF_VALID) != 0, hs * UH);
#pragma unroll 1
            for (int k = 0; k < nsb; ++k) {
                const int b = k % NBUF, ts = sb.len(k);
                float* xb = audio + (size_t)(b * NWT + wt) * tile_floats;
                tc::mbar_wait(full0 + 8 * (wt * NBUF + b), (uint32_t)(k / NBUF) & 1u);
                typename Cell::Lane ln[1];
                ln[0].xcol = xb + lane;
                ln[0].p1col = n_ptiles >= 1 ? xb + p.ts * XP + lane : nullptr;
                ln[0].p2col = n_ptiles >= 2 ? xb + 2 * p.ts * XP + lane : nullptr;
                ln[0].ycol = SPLIT > 1 ? xb + (size_t)(1 + n_ptiles + hs) * p.ts * XP + lane : nullptr;
                ln[0].p1c = 0.f; ln[0].p2c = 0.f;
                ln[0].commit = rnet;
                Cell::run(tt, cst, ln, dw, ts, md.in_gain, md.out_gain, md.in_skip != 0, hs, issuer, db_hi, db_lo);
                __syncwarp();
                tc::mbar_arrive(done0 + 8 * (wt * NBUF + b));
            }
            Cell::store_state(tt[0], cst[0], sa, rslot, hs * UH, rnet);
        }
    }

    tc::tc_fence_before();
    rendezvous(2);
    if (warp == 0) tc::tmem_dealloc(tmem_base, Sh::tmem_cols());
}

} // namespace aidax
