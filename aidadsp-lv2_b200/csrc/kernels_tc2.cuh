// kernels_tc2.cuh -- second-generation tensor-core shape of the fused chain kernel: TWO tiles per thread group in
// ping-pong, optionally the hidden units of a stream split over several threads, optionally the A operand in TMEM.
//
// kernels_tc.cuh (one tile per 128 threads, several CTAs per SM) is bound by latency, not by a pipe: ncu shows the XU
// pipe 65 % busy, issue slots 44 %, 0.7 eligible warps per scheduler.  Per time step a tile goes
//     operands written -> named barrier (all 4 warps) -> MMA issue -> tensor core -> mbarrier wake -> tcgen05.ld -> epilogue
// and nothing in its own warps can run between the barrier and the wake-up; the tiles of an SM drift into lockstep
// (MMA phases with an idle XU pipe, then epilogue phases with a saturated one).  Here a thread owns row r of TWO tiles
// A and B of the same CTA and alternates:
//     wait(A) . epilogue A(t) . publish h_A(t), x_A(t+1) . [arrive | issuer: sync + issue MMA A(t+1)] .
//     wait(B) . epilogue B(t) . publish h_B(t), x_B(t+1) . [arrive | issuer: sync + issue MMA B(t+1)] . ...
// so the tensor-core round trip of one tile is always covered by the other tile's epilogue in the SAME warp, and only
// the issuing warp ever waits at the named barrier (the other warps signal with bar.arrive and go on).
//
// SPLIT: threads per stream in the recurrence.  SPLIT = 1: thread r evaluates all H hidden units of its row (128
// threads per CTA, two streams per thread).  SPLIT = k: warps w, w+4, ... share TMEM lane quadrant w%4; each evaluates
// H/k units of row r in both tiles (128 k threads; the first 256 of them run the filter stages, one stream each) --
// a k-th of the recurrent state per thread, so H = 32/40 keep 8..20 warps per SM busy in the epilogue instead of 2 x 4
// that wait on each other.  The Dense(H->1) partial sums of the parts are combined by the post-network filter stage.
//
// ATMEM: the A operand [h(t-1) | x p1 p2 1] (hi, lo) lives in TENSOR MEMORY next to the accumulator (tcgen05.mma "TS"
// form; rows written with tcgen05.st by their owner threads) instead of shared memory: no per-step STS +
// fence.proxy.async, and the 24..49 KB of shared memory per tile go to a longer audio sub-block.
// (tools/tc_ts_probe.cu checks the TS form stand-alone.)
#pragma once

#include "kernels_tc.cuh"

namespace aidax {

namespace tc {
// Wait for a tile's MMA completion.  AIDAX_MMA_POLL_NS = 0: mbarrier.try_wait with a suspend-time hint (the warp parks, but
// NANOSLEEP.SYNCS wakes on every mbarrier event of the CTA and re-issues the check: ~15 wake-ups of 5 instructions per
// wait in the warp-specialised kernel); > 0: test_wait + a plain nanosleep of that length.
#ifndef AIDAX_TC2_EARLY_X
#define AIDAX_TC2_EARLY_X 0
#endif
#ifndef AIDAX_MMA_POLL_NS
#define AIDAX_MMA_POLL_NS 0
#endif
__device__ __forceinline__ void mbar_wait_mma(uint32_t bar_addr, uint32_t parity)
{
#if AIDAX_MMA_POLL_NS > 0
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
        __nanosleep(AIDAX_MMA_POLL_NS);
    }
#else
    mbar_wait(bar_addr, parity);
#endif
}
} // namespace tc

template <int H, int SPLIT, bool ATMEM, bool PP = true>
struct Tc2Shape {
    using S = TcShape<H>;
    static constexpr int N = S::N, KP = S::KP, KSTEPS = S::KSTEPS, XP = S::XP;
    static constexpr int UH = H / SPLIT;       // hidden units per thread
    static constexpr int NCH = UH / 4;         // 16-column accumulator chunks per thread
    static constexpr int TPT = PP ? 2 : 1;     // tiles per thread: both tiles in ping-pong, or one tile per thread group
    static constexpr int GROUP = 128 * SPLIT;  // threads that own one tile's rows (and meet at its named barrier)
    static constexpr int NTH = PP ? GROUP : 2 * GROUP;
    static constexpr int NW = NTH / 32;
    static constexpr int CS = 256;             // streams per CTA: two tiles of 128
    static constexpr int YB = SPLIT > 1 ? SPLIT : 0; // partial Dense sums kept per half until the post-network stage
    static constexpr int TILE_COLS = N + (ATMEM ? 2 * KP : 0); // TMEM columns of one tile: accumulator [| A hi | A lo]
    static_assert(H % SPLIT == 0 && UH % 4 == 0, "the units of a thread must be whole 16-column accumulator chunks");
    static_assert(2 * TILE_COLS <= 512, "two tiles must fit tensor memory");

    __host__ __device__ static constexpr int tmem_cols()
    {
        int c = 32;
        while (c < 2 * TILE_COLS) c *= 2;
        return c;
    }
    // floats of shared memory that do not depend on the sub-block length
    __host__ __device__ static constexpr int fixed_floats()
    {
        return 2 * S::B_FLOATS + (ATMEM ? 0 : 4 * S::A_FLOATS) + (H + 4) /* dense w, b */ + 4 /* 2 mbarriers */ + 4 /* tmem base + pad */ +
               2 * CS /* slot table, per-stream flags */;
    }
    // audio tiles of the CTA's 8 warp-tiles (32 streams each), expressed per launch warp as the engine sizes it
    // (rounded up: with more than 8 warps a launch warp's share is a fraction of a tile)
    __host__ __device__ static constexpr int warp_floats(int ts, int n_param_tiles) { return (8 * ts * XP * (1 + n_param_tiles + YB) + NW - 1) / NW; }
};

struct Tc2Tile {
    float* a_hi;         // SS: this thread's row in the tile's A operand (a_hi + chunk*512 is a float4 slot)
    float* a_lo;
    uint32_t ta_hi;      // TS: TMEM address (with this warp's lane quadrant) of column 0 of the thread's A row, hi / lo
    uint32_t ta_lo;
    tc_desc_t da_hi;     // SS issue: low words of the A descriptors at K step 0
    tc_desc_t da_lo;
    uint32_t tma_hi;     // TS issue: TMEM address of A (lane 0)
    uint32_t tma_lo;
    uint32_t bar_addr;   // the tile's mbarrier (MMA completion)
    uint32_t phase;
    uint32_t tmem_d;     // accumulator, lane 0
    uint32_t tmem_ld;    // accumulator, this warp's lane quadrant
    uint32_t bar_id;     // named barrier "operands of this tile are written"
};

enum : int { SF_VALID = 1, SF_NET = 2, SF_PREWARM = 4 }; // per-stream flags the filter-stage owner leaves for the recurrence threads

template <int CELL, int H, int SPLIT, bool ATMEM, bool PP, int MATH>
struct Tc2Cell {
    using Sh = Tc2Shape<H, SPLIT, ATMEM, PP>;
    using S = TcShape<H>;
    using Base = TcCell<CELL, H, MATH>;
    static constexpr int UH = Sh::UH, NCH = Sh::NCH, XP = Sh::XP, TPT = Sh::TPT;

    // 3xTF32 split of four values of this thread's A row (tc::split_tf32 + lo rounded to nearest, as TcCell::store_chunk)
    __device__ __forceinline__ static void store_chunk(const Tc2Tile& tt, int chunk, float2 v01, float2 v23)
    {
#if AIDAX_A_RN
        constexpr uint32_t RB = 0x1000u;
#else
        constexpr uint32_t RB = 0u;
#endif
        const float2 h01 = make_float2(__uint_as_float((__float_as_uint(v01.x) + RB) & 0xFFFFE000u), __uint_as_float((__float_as_uint(v01.y) + RB) & 0xFFFFE000u));
        const float2 h23 = make_float2(__uint_as_float((__float_as_uint(v23.x) + RB) & 0xFFFFE000u), __uint_as_float((__float_as_uint(v23.y) + RB) & 0xFFFFE000u));
        const float2 d01 = tc::sub2(v01, h01), d23 = tc::sub2(v23, h23);
        const uint32_t l0 = __float_as_uint(d01.x) + 0x1000u, l1 = __float_as_uint(d01.y) + 0x1000u;
        const uint32_t l2 = __float_as_uint(d23.x) + 0x1000u, l3 = __float_as_uint(d23.y) + 0x1000u;
        if constexpr (ATMEM) {
            tc::tmem_st4(tt.ta_hi + 4 * chunk, __float_as_uint(h01.x), __float_as_uint(h01.y), __float_as_uint(h23.x), __float_as_uint(h23.y));
            tc::tmem_st4(tt.ta_lo + 4 * chunk, l0, l1, l2, l3);
        } else {
            *reinterpret_cast<float4*>(tt.a_hi + chunk * 512) = make_float4(h01.x, h01.y, h23.x, h23.y);
            *reinterpret_cast<float4*>(tt.a_lo + chunk * 512) = make_float4(__uint_as_float(l0), __uint_as_float(l1), __uint_as_float(l2), __uint_as_float(l3));
        }
    }
    __device__ __forceinline__ static void store_chunk(const Tc2Tile& tt, int chunk, float v0, float v1, float v2, float v3)
    {
        store_chunk(tt, chunk, make_float2(v0, v1), make_float2(v2, v3));
    }
    // this thread's operand rows are complete: make them visible to the tensor core (after the barrier that follows)
    __device__ __forceinline__ static void publish()
    {
        if constexpr (ATMEM) tc::tmem_st_wait();
        else tc::fence_proxy_async();
        tc::tc_fence_before();
    }

    // issued by ONE elected thread (see TcCell::issue_mmas for the descriptor economy)
    __device__ __forceinline__ static void issue_mmas(const Tc2Tile& tt, tc_desc_t b_hi, tc_desc_t b_lo)
    {
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, S::N);
        constexpr uint32_t LBO_A = 128 * 16, LBO_B = S::N * 16;
        constexpr uint64_t HI = ((uint64_t)((128u >> 4) & 0x3FFFu) << 32) | (1ull << 46); // SBO = 128 bytes, version 1
#if AIDAX_DESC64
        constexpr uint64_t HIW = 0;
#else
        constexpr uint64_t HIW = HI;
#endif
#pragma unroll
        for (int ks = 0; ks < S::KSTEPS; ++ks) {
            const uint32_t bo = (ks * 2 * LBO_B) >> 4;
            // smallest terms first
            if constexpr (ATMEM) {
                tc::mma_tf32_ts(tt.tmem_d, tt.tma_lo + 8 * ks, HIW | (b_hi + bo), idesc, ks > 0);
                tc::mma_tf32_ts(tt.tmem_d, tt.tma_hi + 8 * ks, HIW | (b_lo + bo), idesc, 1);
                tc::mma_tf32_ts(tt.tmem_d, tt.tma_hi + 8 * ks, HIW | (b_hi + bo), idesc, 1);
            } else {
                const uint32_t ao = (ks * 2 * LBO_A) >> 4;
                tc::mma_tf32(tt.tmem_d, HIW | (tt.da_lo + ao), HIW | (b_hi + bo), idesc, ks > 0);
                tc::mma_tf32(tt.tmem_d, HIW | (tt.da_hi + ao), HIW | (b_lo + bo), idesc, 1);
                tc::mma_tf32(tt.tmem_d, HIW | (tt.da_hi + ao), HIW | (b_hi + bo), idesc, 1);
            }
        }
        tc::mma_commit(tt.bar_addr);
    }

    // recurrent state of this thread's units [u0, u0 + UH) of one stream: cst (LSTM: c, GRU: h) in registers, h split in the A row
    __device__ __forceinline__ static void load_state(const Tc2Tile& tt, float (&cst)[UH], const StreamArrays& sa, int slot, bool load, int u0)
    {
        const size_t cap = (size_t)sa.cap;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            float hv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = u0 + 4 * c + i;
                float h = 0.f, cc = 0.f;
                if (load) {
                    h = sa.rnn[(size_t)u * cap + slot];
                    if (CELL == 0) cc = sa.rnn[(size_t)(kMaxHidden + u) * cap + slot];
                }
                cst[4 * c + i] = CELL == 0 ? cc : h;
                hv[i] = h;
            }
            store_chunk(tt, u0 / 4 + c, hv[0], hv[1], hv[2], hv[3]);
        }
    }
    // LSTM: h is read back from the A row as hi + (lo as the tensor core sees it), see TcCell::store_state.
    // `write`: per-lane (streams of one warp differ: network bypassed, not bound yet ...).  The read-back from tensor memory
    // is a warp-collective tcgen05.ld, so EVERY lane executes it and only the global stores are predicated.
    __device__ __forceinline__ static void store_state(const Tc2Tile& tt, const float (&cst)[UH], const StreamArrays& sa, int slot, int u0, bool write)
    {
        const size_t cap = (size_t)sa.cap;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            float hv[4] = {0.f, 0.f, 0.f, 0.f}, lv[4] = {0.f, 0.f, 0.f, 0.f};
            if (CELL == 0) {
                if constexpr (ATMEM) {
                    uint32_t a[4], b[4];
                    tc::tmem_ld4(tt.ta_hi + 4 * (u0 / 4 + c), a);
                    tc::tmem_ld4(tt.ta_lo + 4 * (u0 / 4 + c), b);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { hv[i] = __uint_as_float(a[i]); lv[i] = __uint_as_float(b[i]); }
                } else {
                    const float4 hi = *reinterpret_cast<const float4*>(tt.a_hi + (u0 / 4 + c) * 512);
                    const float4 lo = *reinterpret_cast<const float4*>(tt.a_lo + (u0 / 4 + c) * 512);
                    hv[0] = hi.x; hv[1] = hi.y; hv[2] = hi.z; hv[3] = hi.w;
                    lv[0] = lo.x; lv[1] = lo.y; lv[2] = lo.z; lv[3] = lo.w;
                }
            }
            if (!write) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = u0 + 4 * c + i;
                if (CELL == 0) {
                    sa.rnn[(size_t)(kMaxHidden + u) * cap + slot] = cst[4 * c + i];
                    sa.rnn[(size_t)u * cap + slot] = hv[i] + __uint_as_float(__float_as_uint(lv[i]) & 0xFFFFE000u);
                } else {
                    sa.rnn[(size_t)u * cap + slot] = cst[4 * c + i];
                }
            }
        }
    }

    // per-tile arguments of run()
    struct Lane {
        float* xcol;        // the stream's column in the audio tile (nullptr = zero input: pre-warm)
        const float* p1col; // param columns or nullptr (then p1c / p2c)
        const float* p2col;
        float* ycol;        // SPLIT > 1: this half's partial Dense sums; SPLIT == 1: unused (y goes to xcol)
        float p1c, p2c;
        bool commit;        // SPLIT == 1: write y into the audio tile
    };

    __device__ __forceinline__ static void sync_and_issue(const Tc2Tile& tt, bool issuer, tc_desc_t b_hi, tc_desc_t b_lo)
    {
#if !(AIDAX_WHATIF & 1)
        if (issuer) {
            tc::bar_sync(tt.bar_id, Sh::GROUP); // every thread that owns rows of this tile has written them
            if (tc::elect_one()) {
                tc::tc_fence_after();
                issue_mmas(tt, b_hi, b_lo);
            }
            __syncwarp();
        } else {
            tc::bar_arrive(tt.bar_id, Sh::GROUP);
        }
#endif
    }

    // n steps of the cell for this thread's units of row r in its TPT tiles (both tiles of the CTA in ping-pong, or the one
    // tile of its thread group).  Every thread of the CTA must call this with the same n.  hs: which part of the hidden
    // units (0 .. SPLIT-1); the hs == 0 thread also feeds the input columns.
    __device__ __forceinline__ static void run(Tc2Tile (&tt)[TPT], float (&cst)[TPT][UH], const Lane (&ln)[TPT], const float* dw, int n, float in_gain, float out_gain, bool in_skip,
                               int hs, bool issuer, tc_desc_t b_hi, tc_desc_t b_lo)
    {
        const int u0 = hs * UH, c0 = u0 / 4;
        const float db = hs == 0 ? dw[H] : 0.f;
        float xg[TPT];
#pragma unroll
        for (int g = 0; g < TPT; ++g) xg[g] = 0.f;
        if (n <= 0) return;

        // ---- prologue: input columns of step 0, first MMAs
#pragma unroll
        for (int g = 0; g < TPT; ++g) {
            if (hs == 0) {
                xg[g] = ln[g].xcol ? ln[g].xcol[0] * in_gain : 0.f;
                const float v1 = ln[g].p1col ? ln[g].p1col[0] : ln[g].p1c;
                const float v2 = ln[g].p2col ? ln[g].p2col[0] : ln[g].p2c;
                store_chunk(tt[g], H / 4, xg[g], v1, v2, 1.0f);
            }
            publish();
            sync_and_issue(tt[g], issuer, b_hi, b_lo);
        }

        for (int t = 0; t < n; ++t) {
#pragma unroll
            for (int g = 0; g < TPT; ++g) {
#if !(AIDAX_WHATIF & 1)
                tc::mbar_wait_mma(tt[g].bar_addr, tt[g].phase);
                tt[g].phase ^= 1u;
#endif
                tc::tc_fence_after();

#if AIDAX_TC2_EARLY_X
                // The next step's input columns do not depend on this step's epilogue, and the MMAs that read x(t) are
                // complete: x(t+1) goes into the A row first, so that its shared-memory load, split and tcgen05.st are off
                // the tail of the step (only the h chunks and the store wait remain in front of the hand-over).
                const float xg_now = xg[g];
                if (t + 1 < n && hs == 0) {
                    xg[g] = ln[g].xcol ? ln[g].xcol[(t + 1) * XP] * in_gain : 0.f;
                    const float v1 = ln[g].p1col ? ln[g].p1col[(t + 1) * XP] : ln[g].p1c;
                    const float v2 = ln[g].p2col ? ln[g].p2col[(t + 1) * XP] : ln[g].p2c;
                    store_chunk(tt[g], H / 4, xg[g], v1, v2, 1.0f);
                }
#else
                const float xg_now = xg[g];
#endif

                // epilogue of this thread's units: pairs of 16-column chunks (8 hidden units) as one straight-line block
                float2 y2 = make_float2(db, 0.f);
#pragma unroll
                for (int c = 0; c < NCH; c += AIDAX_TC_GROUP) {
                    const bool two = AIDAX_TC_GROUP == 2 && c + 1 < NCH;
                    uint32_t raw0[16], raw1[16];
                    tc::tmem_ld16(tt[g].tmem_ld + 16 * (c0 + c), raw0);
                    if (two) tc::tmem_ld16(tt[g].tmem_ld + 16 * (c0 + c + 1), raw1);
                    tc::tmem_ld_wait16(raw0);
                    float2 hn[4];
                    if (two) {
                        tc::tmem_ld_wait16(raw1);
                        Base::template eval_group<true>(raw0, raw1, c, cst[g], hn);
                    } else {
                        Base::template eval_group<false>(raw0, raw0, c, cst[g], hn);
                    }
                    const float2* dw2 = reinterpret_cast<const float2*>(dw + 4 * (c0 + c));
                    y2 = tc::fma2(dw2[0], hn[0], y2);
                    y2 = tc::fma2(dw2[1], hn[1], y2);
                    store_chunk(tt[g], c0 + c, hn[0], hn[1]);
                    if (two) {
                        y2 = tc::fma2(dw2[2], hn[2], y2);
                        y2 = tc::fma2(dw2[3], hn[3], y2);
                        store_chunk(tt[g], c0 + c + 1, hn[2], hn[3]);
                    }
                }
                const float y = y2.x + y2.y;
                // dense + skip + output gain (applyModel, rt-neural-generic.cpp:170-182); with the units split over several
                // threads the halves' partial sums are combined after the sub-block (combine_halves)
                if constexpr (SPLIT == 1) {
                    if (ln[g].xcol && ln[g].commit) ln[g].xcol[t * XP] = (in_skip ? xg_now + y : y) * out_gain;
                } else {
                    if (ln[g].ycol) ln[g].ycol[t * XP] = y;
                }
                // input columns of the next step, then hand the tile back to the tensor core
                if (t + 1 < n) {
                    if (hs == 0 && !AIDAX_TC2_EARLY_X) {
                        xg[g] = ln[g].xcol ? ln[g].xcol[(t + 1) * XP] * in_gain : 0.f;
                        const float v1 = ln[g].p1col ? ln[g].p1col[(t + 1) * XP] : ln[g].p1c;
                        const float v2 = ln[g].p2col ? ln[g].p2col[(t + 1) * XP] : ln[g].p2c;
                        store_chunk(tt[g], H / 4, xg[g], v1, v2, 1.0f);
                    }
                    publish();
                    sync_and_issue(tt[g], issuer, b_hi, b_lo);
                } else {
                    publish(); // h(t) must be in place before the next sub-block's first barrier
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// the fused chain kernel on the two-tile tensor-core recurrence.  CELL: 0 LSTM, 1 GRU
//   PP = true : every thread owns row r of BOTH tiles and alternates their epilogues (128 SPLIT threads)
//   PP = false: thread group g owns tile g (2 x 128 SPLIT threads); the tiles overlap through the warp scheduler, as in
//               kernels_tc.cuh, but a row's hidden units are spread over SPLIT warps of the same lane quadrant
// ------------------------------------------------------------------------------------------
template <int CELL, int H, int SPLIT, bool ATMEM, bool PP, int RES, int MATH>
__global__ void __launch_bounds__(Tc2Shape<H, SPLIT, ATMEM, PP>::NTH, RES) chain_kernel_tc2(const __grid_constant__ RenderParams p)
{
    using Sh = Tc2Shape<H, SPLIT, ATMEM, PP>;
    using S = TcShape<H>;
    using Cell = Tc2Cell<CELL, H, SPLIT, ATMEM, PP, MATH>;
    using Base = TcCell<CELL, H, MATH>;
    constexpr int XP = Sh::XP, NTH = Sh::NTH, CS = Sh::CS, UH = Sh::UH, TPT = Sh::TPT;
    constexpr int FS = NTH >= 256 ? 1 : 2;            // streams per thread in the filter stages: the CTA's 256 streams over its first 256 threads
    constexpr bool WARP_LOCAL = PP && SPLIT == 1;      // a warp's recurrence reads only audio tiles the same warp staged
    extern __shared__ __align__(128) float smem_tc2[];
    float* smem = smem_tc2;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int gsel = PP ? 0 : warp / (4 * SPLIT);      // thread group = tile (PP: every thread has both)
    const int wq = PP ? warp : warp % (4 * SPLIT);
    const int q = wq & 3, hs = wq >> 2, row = q * 32 + lane;
    const bool factive = warp < 8; // filter-stage role: warp w < 8 owns warp-tile w (4 warps per CTA: warp w owns w and 4 + w)
    const int blk = blockIdx.x;
    const StreamArrays& sa = p.sa;

    float* Bh = smem;
    float* Bl = Bh + S::B_FLOATS;
    float* Abase = Bl + S::B_FLOATS;
    float* dw = Abase + (ATMEM ? 0 : 4 * S::A_FLOATS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(dw + H + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    int* wslots = reinterpret_cast<int*>(tmem_slot + 4);
    int* sflags = wslots + CS;

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
        if constexpr (!ATMEM)
            for (int i = tid; i < 4 * S::A_FLOATS; i += NTH) Abase[i] = 0.f; // the K padding of A must be finite zeros
        for (int i = tid; i < CS; i += NTH) {
            wslots[i] = p.blk_slots[(size_t)blk * CS + i];
            sflags[i] = 0;
        }
    }
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, Sh::tmem_cols());
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- this thread's tiles: PP: k = 0, 1 -> tiles 0, 1; otherwise k = 0 -> tile gsel
    Tc2Tile tt[TPT];
    int tile_of[TPT];
    tc_desc_t db_hi, db_lo;
    {
        constexpr uint32_t LBO_A = 128 * 16, LBO_B = S::N * 16, SBO = 128;
        db_hi = (tc_desc_t)tc::make_smem_desc(tc::smem_u32(Bh), LBO_B, SBO);
        db_lo = (tc_desc_t)tc::make_smem_desc(tc::smem_u32(Bl), LBO_B, SBO);
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            const int g = PP ? k : gsel;
            tile_of[k] = g;
            const uint32_t tile_base = tmem_base + g * Sh::TILE_COLS;
            const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
            tt[k].tmem_d = tile_base;
            tt[k].tmem_ld = tile_base + lane_sel;
            tt[k].tma_hi = tile_base + S::N;
            tt[k].tma_lo = tile_base + S::N + S::KP;
            tt[k].ta_hi = tt[k].tma_hi + lane_sel;
            tt[k].ta_lo = tt[k].tma_lo + lane_sel;
            tt[k].a_hi = Abase + g * 2 * S::A_FLOATS + row * 4;
            tt[k].a_lo = tt[k].a_hi + S::A_FLOATS;
            tt[k].da_hi = tt[k].da_lo = 0;
            if constexpr (!ATMEM) {
                const uint32_t a_addr = tc::smem_u32(Abase + g * 2 * S::A_FLOATS);
                tt[k].da_hi = (tc_desc_t)tc::make_smem_desc(a_addr, LBO_A, SBO);
                tt[k].da_lo = (tc_desc_t)tc::make_smem_desc(a_addr + S::A_FLOATS * 4, LBO_A, SBO);
            }
            tt[k].bar_addr = tc::smem_u32(&bars[g]);
            tt[k].phase = 0;
            tt[k].bar_id = 1 + g;
        }
    }
    const bool issuer = wq == 0; // the first warp of a tile's thread group issues that tile's MMAs
    if constexpr (ATMEM) {
        // tensor memory is not initialised: the K padding of this thread's A rows must be finite zeros
        if (hs == 0) {
#pragma unroll
            for (int k = 0; k < TPT; ++k)
#pragma unroll
                for (int c = H / 4; c < S::KP / 4; ++c) Cell::store_chunk(tt[k], c, 0.f, 0.f, 0.f, 0.f);
        }
    }

    // audio tiles: 8 warp-tiles of 32 streams (tile g, quadrant q -> wt = 4 g + q)
    const int tile_floats = p.ts * XP * (1 + n_ptiles + Sh::YB);
    float* audio = smem + ((Sh::fixed_floats() + 3) & ~3);
    auto xbuf_of = [&](int wt) { return audio + (size_t)wt * tile_floats; };
    auto p1buf_of = [&](int wt) { return n_ptiles >= 1 ? xbuf_of(wt) + p.ts * XP : nullptr; };
    auto p2buf_of = [&](int wt) { return n_ptiles >= 2 ? xbuf_of(wt) + 2 * p.ts * XP : nullptr; };
    auto ybuf_of = [&](int wt, int half) { return xbuf_of(wt) + (size_t)(1 + n_ptiles + half) * p.ts * XP; };

    // ---- filter-stage role: warp w < 8 owns the 32 streams of warp-tile w (with 4 warps: w and 4 + w), one per lane
    int fslot[FS], fwt[FS];
    uint32_t fflags[FS], fdfl[FS];
    bool ffirst[FS];
    bool want_prewarm = false;
#pragma unroll
    for (int f = 0; f < FS; ++f) {
        fwt[f] = FS == 2 ? 4 * f + warp : (warp & 7);
        const int si = fwt[f] * 32 + lane;           // index in the CTA's slot table = tile * 128 + row
        fslot[f] = factive ? wslots[si] : -1;
        uint32_t cmd = 0;
        if (fslot[f] >= 0) cmd = exec_commands(sa, fslot[f]);
        int sf = fslot[f] >= 0 ? SF_VALID : 0;
        if (cmd & C_SET_MODEL) { sf |= SF_PREWARM; want_prewarm = true; }
        if (factive) sflags[si] = sf;
    }
    __syncthreads(); // (the slot table and the flags are read across warps below)
    // ---- recurrence role: units [hs*UH, +UH) of row `row` in this thread's tiles
    int rslot[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) rslot[k] = max(wslots[tile_of[k] * 128 + row], 0);

    if (__syncthreads_or(want_prewarm)) {
        // loadModelFromPath: reset() then applyModel over 2048 zero samples with the param smoothers parked at the
        // old targets (rt-neural-generic.cpp:1035, 1052-1079)
        float cst[TPT][UH];
        typename Cell::Lane ln[TPT];
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            Cell::load_state(tt[k], cst[k], sa, rslot[k], false, hs * UH);
            ln[k].xcol = nullptr; ln[k].p1col = nullptr; ln[k].p2col = nullptr; ln[k].ycol = nullptr;
            ln[k].p1c = sa.p_mem[rslot[k]];
            ln[k].p2c = sa.p_mem[(size_t)sa.cap + rslot[k]];
            ln[k].commit = false;
        }
        Cell::run(tt, cst, ln, dw, kPrewarmSteps, md.in_gain, md.out_gain, md.in_skip != 0, hs, issuer, db_hi, db_lo);
#pragma unroll
        for (int k = 0; k < TPT; ++k)
            Cell::store_state(tt[k], cst[k], sa, rslot[k], hs * UH, (sflags[tile_of[k] * 128 + row] & SF_PREWARM) != 0);
        __syncthreads(); // every part's state is in global memory before anyone reloads it
    }

    bool any_net_local = false;
#pragma unroll
    for (int f = 0; f < FS; ++f) {
        fflags[f] = 0; fdfl[f] = 0;
        if (fslot[f] >= 0) { fflags[f] = sa.flags[fslot[f]]; fdfl[f] = sa.dflags[fslot[f]]; }
        ffirst[f] = (fdfl[f] & D_PARAM_FIRST_RUN) != 0;
        const bool net_on = (fflags[f] & F_ENABLED) && (fflags[f] & F_NET);
        if (net_on) { sflags[fwt[f] * 32 + lane] |= SF_NET; any_net_local = true; }
    }
    const bool any_net = __syncthreads_or(any_net_local) != 0;

    float cst[TPT][UH]; // recurrent state of this thread's units, in registers across the whole period
    bool rnet[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        const int sf = sflags[tile_of[k] * 128 + row];
        rnet[k] = (sf & SF_NET) != 0;
        if (any_net) Cell::load_state(tt[k], cst[k], sa, rslot[k], (sf & SF_VALID) != 0, hs * UH);
    }

    for (int t0 = 0; t0 < p.n_samples; t0 += p.ts) {
        const int ts = min(p.ts, p.n_samples - t0);
        // ---- stage in this warp's tiles (the warp-tiles it filters)
#pragma unroll
        for (int f = 0; f < FS; ++f) {
            if (!factive) break;
            const int wt = fwt[f];
            const int* ws = wslots + wt * 32;
            load_tile<32, XP>(xbuf_of(wt), p, ws, t0, ts, lane);
#if AIDAX_PREFETCH_L2
            if (t0 + ts < p.n_samples) { // the NEXT sub-block's rows go to L2 while this one is processed
                const int tn = t0 + ts, nn = min(p.ts, p.n_samples - tn);
                if (p.layout == 0) {
                    if (fslot[f] >= 0) {
                        const float* src = p.in + (size_t)(fslot[f] - p.first_stream) * p.stride + tn;
                        for (int t = 0; t < nn; t += 32) tc::prefetch_l2(src + t);
                    }
                } else {
                    const int s0 = ws[0];
                    if (s0 >= 0) {
                        const float* src = p.in + (size_t)tn * p.stride + (s0 - p.first_stream);
                        for (int t = lane; t < nn; t += 32) tc::prefetch_l2(src + (size_t)t * p.stride);
                    }
                }
            }
#endif
        }
        __syncwarp();

        // ---- stages before the network (run(), rt-neural-generic.cpp:622-644)
#pragma unroll 1
        for (int f = 0; f < FS; ++f) {
            const int wt = fwt[f], cslot = fslot[f];
            const uint32_t flags = fflags[f];
            float* xcol = xbuf_of(wt) + lane;
            float* p1b = p1buf_of(wt);
            float* p2b = p2buf_of(wt);
            if (cslot >= 0 && (flags & F_ENABLED) && !(AIDAX_WHATIF & 8)) {
#if AIDAX_FUSED_PASSES
                pass_fused<XP, false>(xcol, ts, sa, cslot, fused_stages_pre(flags));
#else
                pass_lpf_pregain<XP>(xcol, ts, sa, cslot, (flags & F_LPF) != 0);
                if (flags & F_EQ_PRE) pass_eq<XP>(xcol, ts, sa, cslot, (flags & F_EQ_BANDPASS) != 0);
#endif
                if (flags & F_NET) {
                    pass_param<XP>(p1b ? p1b + lane : nullptr, ts, sa, 0, cslot, ffirst[f], input_size >= 2);
                    pass_param<XP>(p2b ? p2b + lane : nullptr, ts, sa, 1, cslot, ffirst[f], input_size >= 3);
                    if (ffirst[f]) { ffirst[f] = false; sa.dflags[cslot] = fdfl[f] & ~D_PARAM_FIRST_RUN; }
                }
            } else if (n_ptiles > 0 && factive) {
                // the param columns of idle rows still feed the (discarded) product: keep them finite
                for (int t = 0; t < ts; ++t) {
                    p1b[t * XP + lane] = 0.f;
                    if (p2b) p2b[t * XP + lane] = 0.f;
                }
            }
        }
        if constexpr (WARP_LOCAL) __syncwarp();
        else __syncthreads();                   // other warps staged (part of) the columns this thread's recurrence reads

        // ---- the network
        if (any_net) {
            typename Cell::Lane ln[TPT];
#pragma unroll
            for (int k = 0; k < TPT; ++k) {
                const int wt = 4 * tile_of[k] + q;
                ln[k].xcol = xbuf_of(wt) + lane;
                ln[k].p1col = n_ptiles >= 1 ? p1buf_of(wt) + lane : nullptr;
                ln[k].p2col = n_ptiles >= 2 ? p2buf_of(wt) + lane : nullptr;
                ln[k].ycol = SPLIT > 1 ? ybuf_of(wt, hs) + lane : nullptr;
                ln[k].p1c = 0.f; ln[k].p2c = 0.f;
                ln[k].commit = rnet[k];
            }
            Cell::run(tt, cst, ln, dw, ts, md.in_gain, md.out_gain, md.in_skip != 0, hs, issuer, db_hi, db_lo);
        }
        if constexpr (WARP_LOCAL) __syncwarp();
        else __syncthreads();

        // ---- stages after the network (:645-655)
#pragma unroll 1
        for (int f = 0; f < FS; ++f) {
            const int wt = fwt[f], cslot = fslot[f];
            const uint32_t flags = fflags[f];
            float* xcol = xbuf_of(wt) + lane;
            if (cslot >= 0 && (flags & F_ENABLED) && !(AIDAX_WHATIF & 8)) {
                if constexpr (SPLIT > 1) {
                    // Dense partial sums of the parts + skip + output gain (applyModel, rt-neural-generic.cpp:170-182)
                    if (any_net && (flags & F_NET)) {
                        const float* y0 = ybuf_of(wt, 0) + lane;
                        const size_t ystep = (size_t)p.ts * XP;
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
        }
        __syncwarp();
#pragma unroll
        for (int f = 0; f < FS; ++f)
            if (factive) store_tile<32, XP>(xbuf_of(fwt[f]), p, wslots + fwt[f] * 32, t0, ts, lane);
        if constexpr (WARP_LOCAL) __syncwarp();
        else __syncthreads();                   // the next sub-block's staging overwrites tiles other warps may still read
    }

#pragma unroll
    for (int k = 0; k < TPT; ++k)
        Cell::store_state(tt[k], cst[k], sa, rslot[k], hs * UH, rnet[k]);

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, Sh::tmem_cols());
}

} // namespace aidax
