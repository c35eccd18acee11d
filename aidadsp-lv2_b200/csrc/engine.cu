// engine.cu -- host side of the C ABI in include/aidax.h: stream/model bookkeeping, control
// decode, dirty-range uploads, block tables and kernel launches.  CUDA runtime only (no torch).
#include <cuda_runtime.h>
#include <sys/stat.h>

#include <algorithm>
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/aidax.h"
#include "dsp_host.h"
#include "layout.h"
#include "model_host.h"
#include "variants.h"

namespace aidax {

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t err__ = (expr);                                                             \
        if (err__ != cudaSuccess)                                                               \
            return fail(AIDAX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

template <typename T>
struct DevArray {
    T* d = nullptr;
    size_t n = 0;
    // The zero fill runs on the legacy default stream and is NOT ordered against the engine's non-blocking streams:
    // callers that launch work right after an allocation follow it with sync_legacy_stream() (engine_create does once).
    cudaError_t alloc(size_t count, bool zero = true)
    {
        n = count;
        cudaError_t e = cudaMalloc((void**)&d, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess && zero) e = cudaMemset(d, 0, std::max<size_t>(count, 1) * sizeof(T));
        return e;
    }
    void release() { if (d) cudaFree(d); d = nullptr; n = 0; }
};

struct ModelEntry {
    HostModel host;
    float* d_weights = nullptr;
    std::string cache_key; // path + mtime + size of the file it was parsed from ("" = registered from arrays)
    int holds = 0;         // ids handed out by aidax_model_load_json / aidax_model_register and not yet released
    int refs = 0;          // streams bound to it
};

struct Family {       // one kernel family = (cell, hidden); cell 2 = streams without a network
    int cell, H;
    std::vector<const KernelVariant*> variants; // FP32 variants, sorted by streams per warp ascending
    const KernelVariant* tc = nullptr;          // tensor-core variant (kind 1), if compiled for this hidden size
    std::vector<const KernelVariant*> tc2;      // ping-pong tensor-core variants (kind 3), preferred one first
};

struct LaunchPlan {
    const KernelVariant* kv = nullptr;
    int n_blocks = 0;
    size_t tbl_offset_model = 0; // offsets into the uploaded table buffers (ints)
    size_t tbl_offset_slots = 0;
    int max_param_tiles = 0;
    std::vector<std::pair<int, int>> runs; // the launch's streams as maximal runs [a, b) of consecutive ids
};

// hooks around the launches of one range (flush_deferred: per-family input waits and output copies)
struct LaunchHooks {
    std::function<int(const LaunchPlan&, cudaStream_t)> before, after;
};

// the launches that render one stream range [first, first + n): block -> (model, slots) tables on the device
struct RangePlan {
    uint64_t epoch = 0;
    int tile_mode = -1;
    std::vector<LaunchPlan> launches;
    DevArray<int> d_tables;
};

} // namespace aidax

using namespace aidax;

struct aidax_engine {
    int device = 0;
    int cap = 0;
    int max_block = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    int tile_mode = 0, math_mode = 0;
    int tc_min_streams = 0;    // AIDAX_TC_MIN_STREAMS: > 0 overrides the family size from which auto mode takes the tensor-core shape (tc_threshold)
    int tc3_tiles = 0;         // AIDAX_TC3_TILES: 0 auto, 1 one-tile CTAs, 2 multi-tile CTAs (see choose_variant)
    int ts_max = 0;            // AIDAX_TS_MAX: cap of the sub-block length (measurement knob; results do not depend on it)
    int tc_gen = 3, tc2_split = -1, tc2_atmem = -1; // which tensor-core kernel generation / shape (see choose_variant)
    int sm_count = 148;
    size_t smem_optin = 0, smem_per_sm = 0;
    uint64_t launches = 0;

    // host shadows
    std::vector<StreamHost> streams;
    int n_open_hi = 0; // one past the highest stream id ever opened
    std::vector<uint8_t> dirty_ctrl, dirty_coef, dirty_cmd;
    // streams whose PARAM1/2 port value differs from the smoother target the reference would hold: the only
    // ones the per-render host bookkeeping has to visit (keeps a render call O(changed streams), not O(streams))
    std::vector<int> track_list;
    std::vector<uint8_t> track_flag;
    std::vector<uint32_t> h_flags, h_cmd;
    std::vector<float> h_pre_target, h_master_target, h_p_port, h_p_init, h_cmd_pre_mem, h_cmd_master_mem, h_exp_coef, h_lin_denom;
    std::vector<double> h_bq_coef;
    bool any_dirty = false;
    // Uploads that kernels on OTHER streams consume (weights, the model table, block tables) go through one dedicated
    // stream and are waited for before the call returns (upload_blocking): a pageable-memory cudaMemcpy on the legacy
    // stream may return while its DMA is still in flight, and nothing orders that DMA against the engine's non-blocking
    // kernel streams.
    std::mutex mu_up;
    cudaStream_t stream_up = nullptr;

    // device arrays
    DevArray<uint32_t> d_flags, d_cmd, d_dflags;
    DevArray<float> d_pre_target, d_master_target, d_p_port, d_p_init, d_cmd_pre_mem, d_cmd_master_mem, d_exp_coef, d_lin_denom;
    DevArray<float> d_pre_mem, d_master_mem, d_p_mem, d_p_step, d_p_target, d_rnn;
    DevArray<double> d_bq_coef, d_bq_z;
    DevArray<float> d_io_in, d_io_out;
    DevArray<ModelDev> d_models;
    std::vector<int> h_tables;

    std::vector<ModelEntry> models;
    std::map<std::string, int> model_cache; // cache_key -> id, only while the model is held or bound
    bool models_dirty = false;

    std::vector<Family> families;
    // plan cache
    uint64_t membership_epoch = 1;
    std::map<std::pair<int, int>, RangePlan> plans; // keyed by (first, n); the pipelined host path renders in chunks
    std::vector<const void*> attr_done;
    // pipelined host rendering: copy-in / copy-out streams, a few kernel streams (so that the kernels of
    // small ranges share the GPU) and per-range events; `pending` = ranges submitted since the last wait
    static constexpr int kMaxKernelStreams = 16;
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    cudaStream_t kstream[kMaxKernelStreams] = {};
    int n_kstreams = 16; // a heterogeneous period has up to 14 kernel families of a few thousand streams each: all of them fit the GPU at once
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_fan_fork = nullptr, ev_fan_join[kMaxKernelStreams] = {}; // family fan-out inside one range (launch_all)
    int fan_next = 0;
    std::vector<cudaEvent_t> ev_in, ev_done, ev_out;
    bool pipeline_families = true; // AIDAX_PIPELINE_FAMILIES=0: deferred ranges wait for every copy-in and copy out after every kernel
    int grade_unit = 4096, grade_mid = 16384; // piece sizes of aidax_render_host_submit_graded (AIDAX_GRADE_UNIT / AIDAX_GRADE_MID)
    bool defer_hetero = true; // AIDAX_DEFER_HETERO=0: render heterogeneous hand-over ranges one by one (submit_range)
    bool debug_sync = false; // AIDAX_DEBUG_SYNC=1: synchronise after every kernel launch and report which one faulted
    bool trace = false;      // AIDAX_TRACE=1: per-range timeline of the host-fed pipeline on stderr
    int trace_left = 3, trace_skip = 0; // AIDAX_TRACE_SKIP=k: leave out the first k multi-range periods (model binding, fade-in)
    struct PendingRange {
        int first, n, n_samples;
        bool deferred;      // kernels not launched yet: the range waits for the period's other heterogeneous ranges (wait_ranges)
        float* out;         // deferred copy-out target
        size_t stride;
        int layout;
    };
    std::vector<PendingRange> pending;

    StreamArrays sa() const
    {
        StreamArrays a{};
        a.flags = d_flags.d; a.cmd = d_cmd.d; a.pre_target = d_pre_target.d; a.master_target = d_master_target.d;
        a.p_port = d_p_port.d; a.p_init = d_p_init.d; a.cmd_pre_mem = d_cmd_pre_mem.d; a.cmd_master_mem = d_cmd_master_mem.d;
        a.exp_coef = d_exp_coef.d; a.lin_denom = d_lin_denom.d; a.bq_coef = d_bq_coef.d;
        a.dflags = d_dflags.d; a.pre_mem = d_pre_mem.d; a.master_mem = d_master_mem.d; a.p_mem = d_p_mem.d;
        a.p_step = d_p_step.d; a.p_target = d_p_target.d; a.bq_z = d_bq_z.d; a.rnn = d_rnn.d; a.cap = cap;
        return a;
    }
};

namespace aidax {

static void collect_families(aidax_engine* e)
{
    struct Src { const KernelVariant* v; int n; };
    const Src srcs[] = {
        {kVariantsH0, kNumVariantsH0},   {kVariantsH8, kNumVariantsH8},   {kVariantsH12, kNumVariantsH12},
        {kVariantsH16, kNumVariantsH16}, {kVariantsH20, kNumVariantsH20}, {kVariantsH24, kNumVariantsH24},
        {kVariantsH32, kNumVariantsH32}, {kVariantsH40, kNumVariantsH40}, {kVariantsH64, kNumVariantsH64},
        {kVariantsH80, kNumVariantsH80},
    };
    for (const Src& s : srcs)
        for (int i = 0; i < s.n; ++i) {
            const KernelVariant* kv = &s.v[i];
            Family* fam = nullptr;
            for (Family& f : e->families)
                if (f.cell == kv->cell && (kv->cell == 2 || f.H == kv->H)) fam = &f;
            if (!fam) { e->families.push_back(Family{kv->cell, kv->H, {}}); fam = &e->families.back(); }
            if (kv->kind == 1) fam->tc = kv;
            else if (kv->kind == 3) fam->tc2.push_back(kv);
            else fam->variants.push_back(kv);
        }
    for (Family& f : e->families)
        std::stable_sort(f.variants.begin(), f.variants.end(), [](const KernelVariant* a, const KernelVariant* b) { return a->ws < b->ws; });
}

static void mark_ctrl(aidax_engine* e, int s) { e->dirty_ctrl[s] = 1; e->any_dirty = true; }

static void note_param_tracking(aidax_engine* e, int s)
{
    const StreamHost& sh = e->streams[s];
    if (e->track_flag[s]) return;
    if (sh.master_seen != sh.master_target || float_not_equal(sh.param_target[0], sh.ctrl[AIDAX_CTRL_PARAM1]) ||
        float_not_equal(sh.param_target[1], sh.ctrl[AIDAX_CTRL_PARAM2])) {
        e->track_flag[s] = 1;
        e->track_list.push_back(s);
    }
}

static void write_coefs(aidax_engine* e, int s, uint32_t mask)
{
    const StreamHost& sh = e->streams[s];
    for (int b = 0; b < kNumBiquads; ++b) {
        if (!((mask >> b) & 1u)) continue;
        const BiquadCoef& c = sh.bq[b];
        const double v[5] = {c.a0, c.a1, c.a2, c.b1, c.b2};
        for (int i = 0; i < 5; ++i) e->h_bq_coef[(size_t)(b * 5 + i) * e->cap + s] = v[i];
    }
    e->dirty_coef[s] = 1;
    e->any_dirty = true;
}

static void sync_stream_shadow(aidax_engine* e, int s)
{
    const StreamHost& sh = e->streams[s];
    e->h_flags[s] = sh.flags;
    e->h_pre_target[s] = sh.pre_target;
    e->h_master_target[s] = sh.master_target;
    e->h_p_port[s] = sh.ctrl[AIDAX_CTRL_PARAM1];
    e->h_p_port[(size_t)e->cap + s] = sh.ctrl[AIDAX_CTRL_PARAM2];
    mark_ctrl(e, s);
    note_param_tracking(e, s);
}

// upload [lo, hi) of a per-stream array with `rows` rows of pitch cap
template <typename T>
static cudaError_t upload_rows(aidax_engine* e, cudaStream_t st, T* dev, const T* host, int rows, int lo, int hi)
{
    if (hi <= lo) return cudaSuccess;
    if (rows == 1) return cudaMemcpyAsync(dev + lo, host + lo, (size_t)(hi - lo) * sizeof(T), cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dev + lo, (size_t)e->cap * sizeof(T), host + lo, (size_t)e->cap * sizeof(T), (size_t)(hi - lo) * sizeof(T),
                             rows, cudaMemcpyHostToDevice, st);
}

static cudaError_t sync_kernel_streams(aidax_engine* e);

#ifdef AIDAX_LEGACY_UPLOADS
constexpr bool kTableZeroFill = true;  // round 1: cudaMalloc + cudaMemset (a fill KERNEL on the legacy stream) + pageable cudaMemcpy
#else
constexpr bool kTableZeroFill = false;
#endif

// Host -> device copy that has LANDED when the call returns, whatever stream consumes the data afterwards.
// (AIDAX_LEGACY_UPLOADS restores the round-1 behaviour for the fault-reproduction build only, tools/race_probe.sh.)
enum { UP_TABLES = 1, UP_MODELS = 2, UP_WEIGHTS = 4 };
static cudaError_t upload_blocking(aidax_engine* e, void* dst, const void* src, size_t bytes, int what)
{
    if (!bytes) return cudaSuccess;
#ifdef AIDAX_LEGACY_UPLOADS
    // fault-reproduction build: AIDAX_LEGACY_MASK says which uploads go the round-1 way (default: all of them)
    static const int mask = std::getenv("AIDAX_LEGACY_MASK") ? std::atoi(std::getenv("AIDAX_LEGACY_MASK")) : 7;
    if (mask & what) return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
#else
    (void)what;
#endif
    {
        std::lock_guard<std::mutex> lk(e->mu_up);
        cudaError_t err = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream_up);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream_up);
        return err;
    }
}

static int upload_dirty(aidax_engine* e, int first, int n, cudaStream_t st)
{
    if (e->models_dirty) {
        std::vector<ModelDev> md(e->models.size());
        for (size_t i = 0; i < e->models.size(); ++i) {
            const ModelEntry& m = e->models[i];
            md[i] = ModelDev{m.d_weights, m.host.desc.in_gain, m.host.desc.out_gain, m.host.desc.in_skip, m.host.desc.input_size};
        }
        if (e->d_models.n < md.size()) {
            CUDA_TRY(sync_kernel_streams(e)); // launches in flight still read the old table
            e->d_models.release();
            CUDA_TRY(e->d_models.alloc(std::max<size_t>(md.size() * 2, 64), kTableZeroFill));
        }
        CUDA_TRY(upload_blocking(e, e->d_models.d, md.data(), md.size() * sizeof(ModelDev), UP_MODELS));
        e->models_dirty = false;
    }
    if (!e->any_dirty) return AIDAX_OK;
    const int last = std::min(first + n, e->n_open_hi);
    int clo = last, chi = first, flo = last, fhi = first, mlo = last, mhi = first;
    bool remaining = false;
    for (int s = first; s < last; ++s) {
        if (e->dirty_ctrl[s]) { clo = std::min(clo, s); chi = s + 1; e->dirty_ctrl[s] = 0; }
        if (e->dirty_coef[s]) { flo = std::min(flo, s); fhi = s + 1; e->dirty_coef[s] = 0; }
        if (e->dirty_cmd[s]) { mlo = std::min(mlo, s); mhi = s + 1; e->dirty_cmd[s] = 0; }
    }
    if (first > 0 || last < e->n_open_hi) {
        for (int s = 0; s < e->n_open_hi && !remaining; ++s)
            if ((s < first || s >= last) && (e->dirty_ctrl[s] || e->dirty_coef[s] || e->dirty_cmd[s])) remaining = true;
    }
    e->any_dirty = remaining;
#ifdef AIDAX_LEGACY_UPLOADS
    if (std::getenv("AIDAX_UPLOAD_ORDER") == nullptr || std::atoi(std::getenv("AIDAX_UPLOAD_ORDER")) != 0)
        if (!e->pending.empty() && (chi > clo || fhi > flo || mhi > mlo)) CUDA_TRY(sync_kernel_streams(e));
#endif
    CUDA_TRY(upload_rows(e, st, e->d_flags.d, e->h_flags.data(), 1, clo, chi));
    CUDA_TRY(upload_rows(e, st, e->d_pre_target.d, e->h_pre_target.data(), 1, clo, chi));
    CUDA_TRY(upload_rows(e, st, e->d_master_target.d, e->h_master_target.data(), 1, clo, chi));
    CUDA_TRY(upload_rows(e, st, e->d_p_port.d, e->h_p_port.data(), 2, clo, chi));
    CUDA_TRY(upload_rows(e, st, e->d_bq_coef.d, e->h_bq_coef.data(), kNumBiquads * 5, flo, fhi));
    if (mhi > mlo) {
        // Commands for streams outside [mlo, mhi) of this upload keep their device copy: the device
        // clears a command when it executes it, and the host clears its shadow after uploading it.
        CUDA_TRY(upload_rows(e, st, e->d_cmd.d, e->h_cmd.data(), 1, mlo, mhi));
        CUDA_TRY(upload_rows(e, st, e->d_p_init.d, e->h_p_init.data(), 2, mlo, mhi));
        CUDA_TRY(upload_rows(e, st, e->d_cmd_pre_mem.d, e->h_cmd_pre_mem.data(), 1, mlo, mhi));
        CUDA_TRY(upload_rows(e, st, e->d_cmd_master_mem.d, e->h_cmd_master_mem.data(), 1, mlo, mhi));
        CUDA_TRY(upload_rows(e, st, e->d_exp_coef.d, e->h_exp_coef.data(), 1, mlo, mhi));
        CUDA_TRY(upload_rows(e, st, e->d_lin_denom.d, e->h_lin_denom.data(), 1, mlo, mhi));
        // pageable host memory: the async copies above are staged before returning, so the shadow can be cleared
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int s = mlo; s < mhi; ++s) e->h_cmd[s] = 0;
    }
    return AIDAX_OK;
}

// Family size from which auto mode takes the tensor-core shape.  Measured on B200 with the one-tile warp-specialised kernel
// against the FP32 shapes, one family alone on the GPU (profiles/r02aa_small_families.txt, us per 128-sample period at
// 512 / 1024 / 2048 / 4096 streams): H 20: 200 201 206 208 (FP32) vs 188 189 189 190 -> 512;  H 24: 154 154 154 236 vs 210 ->
// 3072;  H 32: 121 158 283 472 vs 261 -> 2048;  H 40: 340 340 341 541 vs 340 -> 2048;  H <= 16: the warp-per-stream shape wins up
// to 40 streams per SM (76 / 129 us at 2048 / 4096 vs 152), checked first in choose_variant.
// `range_streams` = all streams of the range being planned: when a mixed-model period is large enough to fill the GPU
// anyway (C4 sharded over 2 / 4 GPUs: 14 families of 2340 / 1170 streams), the families compete for throughput, not for
// latency, and the tensor-core shape renders a stream with a fraction of the FP32 shapes' issue slots: every family of at
// least 512 streams takes it then.
static int tc_threshold(const aidax_engine* e, const Family& fam, int range_streams)
{
    if (e->tc_min_streams > 0) return e->tc_min_streams;
    const bool loaded = range_streams >= e->sm_count * 64;
    if (loaded) return 512;
    switch (fam.H) {
    case 20: return 512;
    case 24: return 3072;
    case 32: case 40: return 2048;
    default: return 4096;
    }
}

static const KernelVariant* choose_variant(const aidax_engine* e, const Family& fam, int n_streams, int range_streams)
{
    const auto& v = fam.variants;
    const int tc_min = tc_threshold(e, fam, range_streams);
    const bool loaded = e->tc_min_streams <= 0 && range_streams >= e->sm_count * 64 && range_streams > n_streams;
    // auto: small batches take the warp-per-stream shape while it is the fastest (measured on B200, LSTM-16: 76 vs 113 us
    // per period at 2048 streams, 129 vs 207 us at 4096 -- register-tiled and tensor-core shape alike --, 234 vs 207 us
    // at 8192) ...
    // (in a loaded mixed-model period only below the tensor-core threshold: one warp per stream is the most expensive way to
    // render a stream once the GPU is full)
    if (e->tile_mode == 0 && !v.empty() && v.front()->kind == 2 && n_streams <= e->sm_count * 40 && !(loaded && n_streams >= tc_min)) return v.front();
    if ((fam.tc || !fam.tc2.empty()) && (e->tile_mode == 3 || (e->tile_mode == 0 && n_streams >= tc_min))) {
        // AIDAX_TC_GEN: 3 (default) = the warp-specialised shapes of kernels_tc3.cuh (S bit 2), 2 = the two-tile shapes of
        // kernels_tc2.cuh where one is compiled for this hidden size (AIDAX_TC2_SPLIT / AIDAX_TC2_ATMEM pick among several in
        // measurement builds), 1 = kernels_tc.cuh
        if (e->tc_gen >= 2 && !fam.tc2.empty()) {
            if (e->tc_gen == 3) {
                // Warp-specialised shapes come with one tile of 128 streams per CTA and with two (three) tiles.  A family that
                // cannot give every SM a multi-tile CTA pair runs one-tile CTAs: twice as many CTAs, half the recurrence warps
                // per scheduler, so a period's 128 dependent time steps finish sooner (hand-over chunks of the host-fed path,
                // the families of a mixed-model period).  AIDAX_TC3_TILES = 1 / 2 pins the one-tile / multi-tile shape.
                const KernelVariant *one = nullptr, *multi = nullptr;
                for (const KernelVariant* kv : fam.tc2)
                    if (kv->S & 4) {
                        if (kv->cs == 128) { if (!one) one = kv; }
                        else if (!multi) multi = kv;
                    }
                // measured on B200 (profiles/r02w_tc3_one_tile.txt): LSTM-16 152 vs 214 us per period up to 16384 streams, level at
                // 32768, multi-tile ahead beyond; LSTM-40 340 vs 418 us up to 16384, behind at 32768; LSTM-24 (three-tile CTAs,
                // one per SM) 211 vs 354 us up to 16384 and still 300 vs 355 us at 32768
                const int one_tile_max = e->sm_count * ((multi && multi->cs > 256) ? 256 : 128);
                const bool want_one = e->tc3_tiles == 1 || (e->tc3_tiles == 0 && n_streams <= one_tile_max);
                if (one && (want_one || !multi)) return one;
                if (multi) return multi;
            }
            for (int gen = std::min(e->tc_gen, 2); gen >= 2; --gen) {
                for (const KernelVariant* kv : fam.tc2)
                    if (((kv->S & 4) != 0) == (gen == 3) && (e->tc2_split < 0 || kv->L == e->tc2_split) &&
                        (e->tc2_atmem < 0 || (kv->S & 3) == (e->tc2_atmem & 3)))
                        return kv;
                for (const KernelVariant* kv : fam.tc2)
                    if (((kv->S & 4) != 0) == (gen == 3)) return kv;
            }
        }
        if (fam.tc) return fam.tc;
    }
    if (v.size() == 1) return v[0];
    if (e->tile_mode == 1) return v.front();
    if (e->tile_mode == 2) return v.back();
    // ... otherwise the widest tile that still gives every SM at least 4 warps; failing that trade FFMA
    // efficiency for parallelism (fewer streams per warp).
    const int want_warps = e->sm_count * 4;
    for (int i = (int)v.size() - 1; i > 0; --i)
        if ((n_streams + v[i]->ws - 1) / v[i]->ws >= want_warps) return v[i];
    return v.front();
}

// every stream a render kernel may be running on (block tables are replaced only when all of them are idle)
static cudaError_t sync_kernel_streams(aidax_engine* e)
{
    cudaError_t err = cudaStreamSynchronize(e->stream);
    for (int k = 0; k < e->n_kstreams && err == cudaSuccess; ++k)
        if (e->kstream[k] != e->stream) err = cudaStreamSynchronize(e->kstream[k]);
    return err;
}

static int build_plan(aidax_engine* e, int first, int n, RangePlan** out)
{
    if (e->plans.size() > 256) { // ranges come and go (chunk sizes change): drop stale tables
        CUDA_TRY(sync_kernel_streams(e));
        for (auto& kv : e->plans) kv.second.d_tables.release();
        e->plans.clear();
    }
    RangePlan& rp = e->plans[std::make_pair(first, n)];
    *out = &rp;
    if (rp.epoch == e->membership_epoch && rp.tile_mode == e->tile_mode) return AIDAX_OK;
    rp.launches.clear();
    e->h_tables.clear();
    const int last = std::min(first + n, e->n_open_hi);
    int range_streams = 0; // streams of the range that run a network (the launch planner's view of how full the GPU will be)
    for (int s = first; s < last; ++s) range_streams += (e->streams[s].open && e->streams[s].model >= 0) ? 1 : 0;
    for (const Family& fam : e->families) {
        // gather (model, slot) pairs of this family, grouped by model
        std::map<int, std::vector<int>> by_model;
        int total = 0;
        for (int s = first; s < last; ++s) {
            const StreamHost& sh = e->streams[s];
            if (!sh.open) continue;
            if (fam.cell == 2) {
                if (sh.model < 0) { by_model[-1].push_back(s); ++total; }
            } else if (sh.model >= 0) {
                const aidax_model_desc& d = e->models[sh.model].host.desc;
                if (d.cell == fam.cell && d.hidden == fam.H) { by_model[sh.model].push_back(s); ++total; }
            }
        }
        if (!total) continue;
        LaunchPlan lp;
        lp.kv = choose_variant(e, fam, total, range_streams);
        const int CS = lp.kv->cs ? lp.kv->cs : lp.kv->ws * lp.kv->NW;
        std::vector<int> blk_model, blk_slots, all_slots;
        for (auto& kv : by_model) {
            const std::vector<int>& slots = kv.second;
            all_slots.insert(all_slots.end(), slots.begin(), slots.end());
            if (kv.first >= 0) lp.max_param_tiles = std::max(lp.max_param_tiles, e->models[kv.first].host.desc.input_size - 1);
            for (size_t off = 0; off < slots.size(); off += CS) {
                blk_model.push_back(std::max(kv.first, 0));
                for (int j = 0; j < CS; ++j) blk_slots.push_back(off + j < slots.size() ? slots[off + j] : -1);
            }
        }
        std::sort(all_slots.begin(), all_slots.end());
        for (size_t i = 0; i < all_slots.size();) {
            size_t j = i + 1;
            while (j < all_slots.size() && all_slots[j] == all_slots[j - 1] + 1) ++j;
            lp.runs.emplace_back(all_slots[i], all_slots[j - 1] + 1);
            i = j;
        }
        lp.n_blocks = (int)blk_model.size();
        lp.tbl_offset_model = e->h_tables.size();
        e->h_tables.insert(e->h_tables.end(), blk_model.begin(), blk_model.end());
        lp.tbl_offset_slots = e->h_tables.size();
        e->h_tables.insert(e->h_tables.end(), blk_slots.begin(), blk_slots.end());
        rp.launches.push_back(lp);
    }
    // Launch order: the families with the largest shared-memory / tensor-memory footprint first.  The kernels of a
    // heterogeneous period run concurrently (launch_all fans them out); an LSTM-40 CTA needs a whole SM's tensor memory, so
    // launched last it lands on SMs that already hold small-model CTAs and sits in tcgen05.alloc until those retire.
    std::stable_sort(rp.launches.begin(), rp.launches.end(), [](const LaunchPlan& a, const LaunchPlan& b) {
        return a.kv->fixed_smem_floats > b.kv->fixed_smem_floats || (a.kv->fixed_smem_floats == b.kv->fixed_smem_floats && a.kv->H > b.kv->H);
    });
    if (rp.d_tables.n < e->h_tables.size()) {
        CUDA_TRY(sync_kernel_streams(e));
        rp.d_tables.release();
        CUDA_TRY(rp.d_tables.alloc(e->h_tables.size() * 2 + 1024, kTableZeroFill));
    }
    if (!e->h_tables.empty()) {
        CUDA_TRY(sync_kernel_streams(e)); // a previous launch may still read the old table
        // ... and the launch that follows (on a non-blocking kernel stream) must see the new one: see upload_blocking
        CUDA_TRY(upload_blocking(e, rp.d_tables.d, e->h_tables.data(), e->h_tables.size() * sizeof(int), UP_TABLES));
    }
    rp.epoch = e->membership_epoch;
    rp.tile_mode = e->tile_mode;
    return AIDAX_OK;
}

static int launch_all(aidax_engine* e, cudaStream_t st, int first, int n, int n_samples, const float* d_in, float* d_out, int layout, size_t stride,
                      const LaunchHooks* hooks = nullptr)
{
    int rc = upload_dirty(e, first, n, st);
    if (rc) return rc;
    RangePlan* rp = nullptr;
    rc = build_plan(e, first, n, &rp);
    if (rc) return rc;
    // A range with several kernel families (heterogeneous models, C4) fans its launches out over the kernel streams:
    // the families own disjoint streams, each launch alone underfills the GPU (a few thousand streams), and run one
    // after the other they cost the sum of their latencies.  Fork from / join back into `st` with events, so callers
    // keep seeing one ordered stream.
    const bool fan = rp->launches.size() > 1 && e->n_kstreams > 1 && !e->debug_sync;
    if (fan) CUDA_TRY(cudaEventRecord(e->ev_fan_fork, st));
    unsigned used = 0;
    int li = 0;
    for (const LaunchPlan& lp : rp->launches) {
        const KernelVariant* kv = lp.kv;
        cudaStream_t ls = st;
        if (fan) {
            const int k = (e->fan_next + li++) % e->n_kstreams;
            if (e->kstream[k] != st) {
                ls = e->kstream[k];
                if (!(used & (1u << k))) CUDA_TRY(cudaStreamWaitEvent(ls, e->ev_fan_fork, 0));
                used |= 1u << k;
            }
        }
        // samples per shared-memory sub-block: as many as fit with the variant's target residency, capped at the period
        // the warp-specialised shapes pipeline sub-blocks between their filter and recurrence warps: shorter sub-blocks
        // shorten the pipeline's fill and drain (measured on B200: 16 beats 32 and 8 for every hidden size up to 20,
        // 8..16 are level for H 32/40)
        const int ts_cap = e->ts_max > 0 ? e->ts_max : ((kv->kind == 3 && (kv->S & 4)) ? 16 : 0);
        int ts = ts_cap > 0 ? std::min(n_samples, ts_cap) : n_samples;
        const size_t budget = std::min(e->smem_optin, e->smem_per_sm / (size_t)kv->ctas_per_sm - 1024);
        auto smem_for = [&](int t) { return (size_t)(kv->fixed_smem_floats + kv->NW * kv->warp_smem_floats(t, lp.max_param_tiles)) * sizeof(float); };
        // halving steps with the 3/4 points in between (128, 96, 64, 48, 32, 24, 16, 12, 8): results do not depend on ts
        while (ts > 8 && smem_for(ts) > budget) {
            const int three_quarters = (ts * 3 / 4 + 3) & ~3;
            ts = (three_quarters < ts && three_quarters > (ts + 1) / 2 && ts % 3 != 0) ? three_quarters : (ts % 3 == 0 ? ts * 2 / 3 : (ts + 1) / 2);
        }
        if (smem_for(ts) > e->smem_optin) return fail(AIDAX_ERR_CAPACITY, "kernel variant needs %zu B shared memory", smem_for(ts));
        const size_t smem = smem_for(ts);
        const void* fn = kv->fn[e->math_mode ? 1 : 0];
        if (std::find(e->attr_done.begin(), e->attr_done.end(), fn) == e->attr_done.end()) {
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_optin));
            e->attr_done.push_back(fn);
        }
        RenderParams p{};
        p.in = d_in; p.out = d_out; p.layout = layout; p.stride = (long long)stride; p.first_stream = first;
        p.n_samples = n_samples; p.ts = ts; p.math_mode = e->math_mode;
        p.blk_model = rp->d_tables.d + lp.tbl_offset_model;
        p.blk_slots = rp->d_tables.d + lp.tbl_offset_slots;
        p.models = e->d_models.d;
        p.sa = e->sa();
        void* args[] = {&p};
        if (hooks && hooks->before) { const int rh = hooks->before(lp, ls); if (rh) return rh; }
        CUDA_TRY(cudaLaunchKernel(fn, dim3(lp.n_blocks), dim3(kv->NW * 32), args, smem, ls));
        e->launches++;
        if (hooks && hooks->after) { const int rh = hooks->after(lp, ls); if (rh) return rh; }
        if (e->debug_sync) { // AIDAX_DEBUG_SYNC=1: attribute an asynchronous fault to the launch that caused it
            const cudaError_t err = cudaStreamSynchronize(ls);
            if (err != cudaSuccess)
                return fail(AIDAX_ERR_CUDA, "kernel fault: %s -- cell %d H %d L %d S %d kind %d, range [%d +%d), %d blocks x %d threads, ts %d, smem %zu, layout %d",
                            cudaGetErrorString(err), kv->cell, kv->H, kv->L, kv->S, kv->kind, first, n, lp.n_blocks, kv->NW * 32, ts, smem, layout);
        }
    }
    if (fan) {
        e->fan_next = (e->fan_next + li) % e->n_kstreams;
        for (int k = 0; k < e->n_kstreams; ++k)
            if (used & (1u << k)) {
                CUDA_TRY(cudaEventRecord(e->ev_fan_join[k], e->kstream[k]));
                CUDA_TRY(cudaStreamWaitEvent(st, e->ev_fan_join[k], 0));
            }
    }
    return AIDAX_OK;
}

// call with e->mu held (n_open_hi and the open flags change under it)
static int check_stream(aidax_engine* e, int s)
{
    if (s < 0 || s >= e->n_open_hi || !e->streams[s].open) return fail(AIDAX_ERR_ARG, "unknown stream %d", s);
    return AIDAX_OK;
}

} // namespace aidax

// ------------------------------------------------------------------ FP32 pipe peak probe
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters)
{
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 1.0f + 1e-3f * (float)(threadIdx.x + j);
    const float x = 0.9999f, y = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == 123.456f) out[0] = s; // never true; keeps the loop alive
}

// ================================================================================ C ABI
extern "C" {

const char* aidax_last_error(void) { return g_last_error.c_str(); }
int aidax_abi_version(void) { return AIDAX_ABI_VERSION; }

int aidax_engine_create(int device, int max_streams, int max_block, aidax_engine** out)
{
    if (!out || max_streams <= 0 || max_block <= 0) return fail(AIDAX_ERR_ARG, "bad engine_create arguments");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(AIDAX_ERR_CUDA, "no CUDA device available (the engine has no CPU fallback)");
    if (device < 0 || device >= n_dev) return fail(AIDAX_ERR_ARG, "device %d out of range (%d devices)", device, n_dev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(AIDAX_ERR_CUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
    std::unique_ptr<aidax_engine> e(new aidax_engine());
    e->device = device;
    e->cap = (max_streams + 31) & ~31;
    e->max_block = max_block;
    e->sm_count = prop.multiProcessorCount;
    e->smem_optin = prop.sharedMemPerBlockOptin;
    e->smem_per_sm = prop.sharedMemPerMultiprocessor;
    if (const char* v = std::getenv("AIDAX_TC_MIN_STREAMS")) e->tc_min_streams = std::atoi(v);
    if (const char* v = std::getenv("AIDAX_TC_GEN")) e->tc_gen = std::min(3, std::max(1, std::atoi(v)));
    if (const char* v = std::getenv("AIDAX_TS_MAX")) e->ts_max = std::atoi(v);
    if (const char* v = std::getenv("AIDAX_TC3_TILES")) e->tc3_tiles = std::atoi(v);
    if (const char* v = std::getenv("AIDAX_TC2_SPLIT")) e->tc2_split = std::atoi(v);
    if (const char* v = std::getenv("AIDAX_TC2_ATMEM")) e->tc2_atmem = std::atoi(v);
    if (const char* v = std::getenv("AIDAX_TILE_MODE")) e->tile_mode = std::max(0, std::min(std::atoi(v), 3)); // initial kernel-shape choice (aidax_engine_configure overrides)
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream_out, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream_up, cudaStreamNonBlocking));
    if (const char* v = std::getenv("AIDAX_KERNEL_STREAMS")) e->n_kstreams = std::max(1, std::min(std::atoi(v), (int)aidax_engine::kMaxKernelStreams));
    e->kstream[0] = e->stream;
    for (int k = 1; k < e->n_kstreams; ++k) CUDA_TRY(cudaStreamCreateWithFlags(&e->kstream[k], cudaStreamNonBlocking));
    e->trace = std::getenv("AIDAX_TRACE") != nullptr;
    if (const char* v = std::getenv("AIDAX_TRACE_SKIP")) e->trace_skip = std::atoi(v);
    e->debug_sync = std::getenv("AIDAX_DEBUG_SYNC") != nullptr;
    if (const char* v = std::getenv("AIDAX_DEFER_HETERO")) e->defer_hetero = std::atoi(v) != 0;
    if (const char* v = std::getenv("AIDAX_PIPELINE_FAMILIES")) e->pipeline_families = std::atoi(v) != 0;
    if (const char* v = std::getenv("AIDAX_GRADE_UNIT")) e->grade_unit = std::max(128, std::atoi(v));
    if (const char* v = std::getenv("AIDAX_GRADE_MID")) e->grade_mid = std::max(e->grade_unit, std::atoi(v));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, e->trace ? cudaEventDefault : cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fan_fork, cudaEventDisableTiming));
    for (int k = 0; k < e->n_kstreams; ++k) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fan_join[k], cudaEventDisableTiming));
    const size_t cap = (size_t)e->cap;
    e->streams.resize(cap);
    e->dirty_ctrl.assign(cap, 0); e->dirty_coef.assign(cap, 0); e->dirty_cmd.assign(cap, 0);
    e->track_flag.assign(cap, 0);
    e->h_flags.assign(cap, 0); e->h_cmd.assign(cap, 0);
    e->h_pre_target.assign(cap, 0.f); e->h_master_target.assign(cap, 0.f);
    e->h_p_port.assign(2 * cap, 0.f); e->h_p_init.assign(2 * cap, 0.f);
    e->h_cmd_pre_mem.assign(cap, 0.f); e->h_cmd_master_mem.assign(cap, 0.f);
    e->h_exp_coef.assign(cap, 0.f); e->h_lin_denom.assign(cap, 4800.f);
    e->h_bq_coef.assign((size_t)kNumBiquads * 5 * cap, 0.0);
    CUDA_TRY(e->d_flags.alloc(cap)); CUDA_TRY(e->d_cmd.alloc(cap)); CUDA_TRY(e->d_dflags.alloc(cap));
    CUDA_TRY(e->d_pre_target.alloc(cap)); CUDA_TRY(e->d_master_target.alloc(cap));
    CUDA_TRY(e->d_p_port.alloc(2 * cap)); CUDA_TRY(e->d_p_init.alloc(2 * cap));
    CUDA_TRY(e->d_cmd_pre_mem.alloc(cap)); CUDA_TRY(e->d_cmd_master_mem.alloc(cap));
    CUDA_TRY(e->d_exp_coef.alloc(cap)); CUDA_TRY(e->d_lin_denom.alloc(cap));
    CUDA_TRY(e->d_pre_mem.alloc(cap)); CUDA_TRY(e->d_master_mem.alloc(cap));
    CUDA_TRY(e->d_p_mem.alloc(2 * cap)); CUDA_TRY(e->d_p_step.alloc(2 * cap)); CUDA_TRY(e->d_p_target.alloc(2 * cap));
    CUDA_TRY(e->d_rnn.alloc((size_t)2 * kMaxHidden * cap));
    CUDA_TRY(e->d_bq_coef.alloc((size_t)kNumBiquads * 5 * cap)); CUDA_TRY(e->d_bq_z.alloc((size_t)kNumBiquads * 2 * cap));
    CUDA_TRY(e->d_io_in.alloc(cap * (size_t)max_block)); CUDA_TRY(e->d_io_out.alloc(cap * (size_t)max_block));
    CUDA_TRY(cudaDeviceSynchronize()); // the zero fills above ran on the legacy stream; everything else uses non-blocking streams
    collect_families(e.get());
    *out = e.release();
    return AIDAX_OK;
}

int aidax_engine_destroy(aidax_engine* e)
{
    if (!e) return AIDAX_OK;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (ModelEntry& m : e->models) if (m.d_weights) cudaFree(m.d_weights);
    e->d_flags.release(); e->d_cmd.release(); e->d_dflags.release(); e->d_pre_target.release(); e->d_master_target.release();
    e->d_p_port.release(); e->d_p_init.release(); e->d_cmd_pre_mem.release(); e->d_cmd_master_mem.release();
    e->d_exp_coef.release(); e->d_lin_denom.release(); e->d_pre_mem.release(); e->d_master_mem.release();
    e->d_p_mem.release(); e->d_p_step.release(); e->d_p_target.release(); e->d_rnn.release();
    e->d_bq_coef.release(); e->d_bq_z.release(); e->d_io_in.release(); e->d_io_out.release();
    e->d_models.release();
    for (auto& kv : e->plans) kv.second.d_tables.release();
    for (cudaEvent_t ev : e->ev_in) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_done) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_out) cudaEventDestroy(ev);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_fan_fork) cudaEventDestroy(e->ev_fan_fork);
    for (cudaEvent_t ev : e->ev_fan_join) if (ev) cudaEventDestroy(ev);
    for (int k = 1; k < e->n_kstreams; ++k) if (e->kstream[k]) cudaStreamDestroy(e->kstream[k]);
    if (e->stream_in) cudaStreamDestroy(e->stream_in);
    if (e->stream_out) cudaStreamDestroy(e->stream_out);
    if (e->stream_up) cudaStreamDestroy(e->stream_up);
    cudaStreamDestroy(e->stream);
    delete e;
    return AIDAX_OK;
}

void* aidax_engine_cuda_stream(aidax_engine* e) { return e ? (void*)e->stream : nullptr; }

int aidax_engine_configure(aidax_engine* e, int tile_mode, int math_mode)
{
    if (!e || tile_mode < 0 || tile_mode > 3 || math_mode < 0 || math_mode > 1) return fail(AIDAX_ERR_ARG, "bad configure arguments");
    std::lock_guard<std::mutex> lk(e->mu);
    e->tile_mode = tile_mode;
    e->math_mode = math_mode;
    return AIDAX_OK;
}

// Packs the weights and puts them on the device WITHOUT the engine lock (parsing, packing, cudaMalloc and the copy may
// take milliseconds and run on the LV2 worker thread; every instance's audio-thread run() takes e->mu), then takes
// the lock only to append the entry.  The returned id is held (see aidax_model_release).
static int register_host_model(aidax_engine* e, HostModel&& hm, const std::string& cache_key, int* id_out)
{
    const std::vector<float> packed = pack_model(hm);
    ModelEntry me;
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaMalloc((void**)&me.d_weights, packed.size() * sizeof(float)));
    const cudaError_t err = upload_blocking(e, me.d_weights, packed.data(), packed.size() * sizeof(float), UP_WEIGHTS);
    if (err != cudaSuccess) {
        cudaFree(me.d_weights);
        return fail(AIDAX_ERR_CUDA, "weight upload failed: %s", cudaGetErrorString(err));
    }
    me.host = std::move(hm);
    me.cache_key = cache_key;
    me.holds = 1;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!cache_key.empty()) {
        // another thread may have loaded the same file meanwhile: keep one copy
        auto it = e->model_cache.find(cache_key);
        if (it != e->model_cache.end()) {
            cudaFree(me.d_weights);
            e->models[it->second].holds++;
            if (id_out) *id_out = it->second;
            return AIDAX_OK;
        }
    }
    e->models.push_back(std::move(me));
    e->models_dirty = true;
    const int id = (int)e->models.size() - 1;
    if (!cache_key.empty()) e->model_cache[cache_key] = id;
    if (id_out) *id_out = id;
    return AIDAX_OK;
}

// Drops the device weights of a model nobody holds or is bound to (call with e->mu held).  The pointer is returned so
// that the caller frees it after unlocking: cudaFree waits for the device, and in-flight kernels may still read it.
static float* retire_model_locked(aidax_engine* e, int id)
{
    ModelEntry& m = e->models[id];
    if (m.holds > 0 || m.refs > 0 || !m.d_weights) return nullptr;
    float* w = m.d_weights;
    m.d_weights = nullptr;
    if (!m.cache_key.empty()) {
        auto it = e->model_cache.find(m.cache_key);
        if (it != e->model_cache.end() && it->second == id) e->model_cache.erase(it);
    }
    m.host.W.clear(); m.host.U.clear(); m.host.b.clear(); m.host.dense_w.clear();
    m.host.W.shrink_to_fit(); m.host.U.shrink_to_fit();
    e->models_dirty = true;
    return w;
}

int aidax_model_register(aidax_engine* e, const aidax_model_desc* desc, const float* W, const float* U, const float* b,
                         const float* dense_w, float dense_b, int* model_id_out)
{
    if (!e || !desc || !W || !U || !b || !dense_w) return fail(AIDAX_ERR_ARG, "null argument");
    if (desc->cell != AIDAX_CELL_LSTM && desc->cell != AIDAX_CELL_GRU) return fail(AIDAX_ERR_MODEL, "unknown cell type %d", desc->cell);
    if (!hidden_size_supported(desc->hidden)) return fail(AIDAX_ERR_MODEL, "hidden size %d is not in the variant table", desc->hidden);
    if (desc->input_size < 1 || desc->input_size > kMaxInput) return fail(AIDAX_ERR_MODEL, "input_size %d not supported", desc->input_size);
    const int H = desc->hidden, IN = desc->input_size, G = desc->cell == AIDAX_CELL_LSTM ? 4 : 3;
    HostModel hm;
    hm.desc = *desc;
    hm.desc.in_skip = desc->in_skip != 0;
    hm.W.assign(W, W + (size_t)IN * G * H);
    hm.U.assign(U, U + (size_t)H * G * H);
    hm.b.assign(b, b + (size_t)(desc->cell == AIDAX_CELL_LSTM ? 4 * H : 6 * H));
    hm.dense_w.assign(dense_w, dense_w + H);
    hm.dense_b = dense_b;
    return register_host_model(e, std::move(hm), std::string(), model_id_out);
}

// The reference re-reads the file on every patch:Set (loadModelFromPath, rt-neural-generic.cpp:963-1013).  Here a
// file is parsed once per (path, modification time, size) while some instance still holds or runs the result, so
// that N instances loading one capture share one weight block -- and a capture re-exported to the same path is read
// again.
static std::string model_cache_key(const char* path)
{
    struct stat st;
    if (stat(path, &st) != 0) return std::string();
    char buf[96];
    snprintf(buf, sizeof(buf), "|%lld.%09ld|%lld", (long long)st.st_mtim.tv_sec, (long)st.st_mtim.tv_nsec, (long long)st.st_size);
    return std::string(path) + buf;
}

int aidax_model_load_json(aidax_engine* e, const char* path, int* model_id_out, int* input_size_out)
{
    if (!e || !path) return fail(AIDAX_ERR_ARG, "null argument");
    const std::string key = model_cache_key(path);
    int id = -1;
    if (!key.empty()) {
        std::lock_guard<std::mutex> lk(e->mu);
        auto it = e->model_cache.find(key);
        if (it != e->model_cache.end()) {
            id = it->second;
            e->models[id].holds++;
        }
    }
    if (id < 0) {
        HostModel hm;
        try {
            hm = load_model_file(path);
        } catch (const std::exception& ex) {
            return fail(AIDAX_ERR_MODEL, "Unable to load json file: %s Error: %s", path, ex.what());
        }
        const int rc = register_host_model(e, std::move(hm), key, &id);
        if (rc) return rc;
    }
    std::lock_guard<std::mutex> lk(e->mu);
    if (model_id_out) *model_id_out = id;
    if (input_size_out) *input_size_out = e->models[id].host.desc.input_size;
    return AIDAX_OK;
}

int aidax_model_release(aidax_engine* e, int model_id)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    float* dead = nullptr;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (model_id < 0 || model_id >= (int)e->models.size()) return fail(AIDAX_ERR_ARG, "unknown model %d", model_id);
        ModelEntry& m = e->models[model_id];
        if (m.holds <= 0) return fail(AIDAX_ERR_STATE, "model %d is not held", model_id);
        m.holds--;
        dead = retire_model_locked(e, model_id);
    }
    if (dead) {
        cudaSetDevice(e->device);
        cudaFree(dead);
    }
    return AIDAX_OK;
}

int aidax_model_describe(aidax_engine* e, int model_id, aidax_model_desc* out)
{
    if (!e || !out) return fail(AIDAX_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(e->mu);
    if (model_id < 0 || model_id >= (int)e->models.size()) return fail(AIDAX_ERR_ARG, "unknown model %d", model_id);
    *out = e->models[model_id].host.desc;
    return AIDAX_OK;
}

int aidax_model_check_json(const char* path, aidax_model_desc* out)
{
    if (!path) return fail(AIDAX_ERR_ARG, "null path");
    try {
        const HostModel hm = load_model_file(path);
        (void)pack_model(hm);
        if (out) *out = hm.desc;
    } catch (const std::exception& ex) {
        return fail(AIDAX_ERR_MODEL, "Unable to load json file: %s Error: %s", path, ex.what());
    }
    return AIDAX_OK;
}

// The reference's load-time sanity check (RtNeuralGeneric::testModel, rt-neural-generic.cpp:900-955; compiled into DEBUG
// builds at :1066-1073): the network alone over the file's `input_batch`, gains forced to 1, both parameters at 0, from
// reset(), compared sample by sample with `output_batch` at TEST_MODEL_THR = 1e-5 (rt-neural-generic.h:182).  Here it runs
// on the GPU through the ordinary render path (a scratch stream bound to a gain-less copy of the model), so it checks what
// the kernels of THIS engine configuration compute for this capture.
int aidax_model_selftest(aidax_engine* e, int model_id, float* max_err_out, int* n_errs_out, int* n_samples_out)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    HostModel hm;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (model_id < 0 || model_id >= (int)e->models.size() || !e->models[model_id].d_weights) return fail(AIDAX_ERR_ARG, "unknown model %d", model_id);
        hm = e->models[model_id].host;
    }
    if (hm.test_in.empty()) return fail(AIDAX_ERR_STATE, "model %d carries no input_batch / output_batch", model_id);
    const std::vector<float> x = hm.test_in, want = hm.test_out;
    hm.desc.in_gain = 1.0f;  // "Gain correction inject unwanted errors" (:903-907)
    hm.desc.out_gain = 1.0f;
    hm.test_in.clear(); hm.test_out.clear();
    int tmp_model = -1, s = -1;
    int rc = register_host_model(e, std::move(hm), std::string(), &tmp_model);
    if (rc) return rc;
    std::vector<float> y(x.size(), 0.f);
    rc = aidax_stream_open(e, 48000.0, &s);
    if (rc == AIDAX_OK) {
        float c[AIDAX_N_CONTROLS] = {};
        c[AIDAX_CTRL_EQBYPASS] = 1.0f;   // the network alone: no in_lpf, no EQ, no dc blocker, unity gains
        c[AIDAX_CTRL_ENABLED] = 1.0f;
        c[AIDAX_CTRL_BFREQ] = 305.0f; c[AIDAX_CTRL_MFREQ] = 750.0f; c[AIDAX_CTRL_MIDQ] = 0.707f; c[AIDAX_CTRL_TFREQ] = 2000.0f;
        rc = aidax_stream_set_model(e, s, tmp_model, 0.f, 0.f);  // "Conditioned models tested with all params at 0" (:908-916)
        if (rc == AIDAX_OK) rc = aidax_stream_set_controls(e, s, c, 0);
        const int blk = std::min(e->max_block, 128);
        std::vector<float> zero((size_t)blk, 0.f), sink((size_t)blk, 0.f);
        // one silent period executes the stream's commands (open, model binding incl. its pre-warm) ...
        if (rc == AIDAX_OK) rc = aidax_render_host(e, s, 1, blk, zero.data(), sink.data(), AIDAX_LAYOUT_PLANAR, (size_t)blk);
        // ... then the test starts from reset() with the gain ramps at their targets
        const std::vector<float> z((size_t)kMaxHidden, 0.f);
        if (rc == AIDAX_OK) rc = aidax_stream_set_state(e, s, z.data(), z.data(), nullptr);
        if (rc == AIDAX_OK) rc = aidax_stream_activate(e, s);
        for (size_t off = 0; off < x.size() && rc == AIDAX_OK; off += (size_t)e->max_block) {
            const int n = (int)std::min<size_t>((size_t)e->max_block, x.size() - off);
            rc = aidax_render_host(e, s, 1, n, x.data() + off, y.data() + off, AIDAX_LAYOUT_PLANAR, (size_t)n);
        }
        aidax_stream_close(e, s);
    }
    aidax_model_release(e, tmp_model);
    if (rc) return rc;
    float max_err = 0.f;
    int n_errs = 0;
    for (size_t i = 0; i < x.size(); ++i) {
        const float err = std::fabs(y[i] - want[i]);
        max_err = std::max(max_err, err);
        if ((double)err > 1.0e-5) ++n_errs; // TEST_MODEL_THR
    }
    if (max_err_out) *max_err_out = max_err;
    if (n_errs_out) *n_errs_out = n_errs;
    if (n_samples_out) *n_samples_out = (int)x.size();
    return AIDAX_OK;
}

int aidax_design_biquad(int type, double fc, double q, double gain_db, double coef[5])
{
    if (!coef || type < 0 || type > 6) return fail(AIDAX_ERR_ARG, "bad biquad arguments");
    const BiquadCoef c = design_biquad(type, fc, q, gain_db);
    coef[0] = c.a0; coef[1] = c.a1; coef[2] = c.a2; coef[3] = c.b1; coef[4] = c.b2;
    return AIDAX_OK;
}

int aidax_decode_controls(double sample_rate, const float controls[AIDAX_N_CONTROLS], int loading, int model_bound,
                          uint32_t* flags, float* pre_target, float* master_target, double coefs[35])
{
    if (!controls) return fail(AIDAX_ERR_ARG, "null controls");
    StreamHost sh;
    init_stream_host(sh, sample_rate);
    decode_controls(sh, controls, loading != 0, model_bound != 0);
    if (flags) *flags = sh.flags;
    if (pre_target) *pre_target = sh.pre_target;
    if (master_target) *master_target = sh.master_target;
    if (coefs)
        for (int b = 0; b < kNumBiquads; ++b) {
            const BiquadCoef& c = sh.bq[b];
            coefs[b * 5 + 0] = c.a0; coefs[b * 5 + 1] = c.a1; coefs[b * 5 + 2] = c.a2; coefs[b * 5 + 3] = c.b1; coefs[b * 5 + 4] = c.b2;
        }
    return AIDAX_OK;
}

int aidax_stream_open(aidax_engine* e, double sample_rate, int* stream_id_out)
{
    if (!e || !stream_id_out) return fail(AIDAX_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(e->mu);
    int s = -1;
    if (e->n_open_hi < e->cap) s = e->n_open_hi++;
    else
        for (int i = 0; i < e->cap; ++i)
            if (!e->streams[i].open) { s = i; break; }
    if (s < 0) return fail(AIDAX_ERR_CAPACITY, "all %d stream slots are in use", e->cap);
    StreamHost& sh = e->streams[s];
    init_stream_host(sh, sample_rate);
    decode_controls(sh, sh.ctrl, /*loading=*/true, /*model_bound=*/false);
    sh.have_ctrl = false;
    sh.pre_target = 1.f; // instantiate: preGain target 1, master target 0 until the first run()
    sh.master_target = 0.f;
    sync_stream_shadow(e, s);
    write_coefs(e, s, (1u << kNumBiquads) - 1u);
    e->h_exp_coef[s] = exp_smoother_coef(sample_rate);
    e->h_lin_denom[s] = 0.1f * 48000.0f;
    e->h_cmd[s] = C_SET_GAIN_MEMS | C_ZERO_FILTERS;
    e->h_cmd_pre_mem[s] = 1.f;
    e->h_cmd_master_mem[s] = 0.f;
    e->dirty_cmd[s] = 1;
    e->any_dirty = true;
    e->membership_epoch++;
    *stream_id_out = s;
    return AIDAX_OK;
}

int aidax_stream_close(aidax_engine* e, int stream)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    float* dead = nullptr;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        int rc = check_stream(e, stream);
        if (rc) return rc;
        StreamHost& sh = e->streams[stream];
        if (sh.model >= 0 && --e->models[sh.model].refs == 0) dead = retire_model_locked(e, sh.model);
        sh.open = false;
        sh.model = -1;
        e->membership_epoch++;
    }
    if (dead) {
        cudaSetDevice(e->device);
        cudaFree(dead); // waits for the device: launches in flight may still read the block
    }
    return AIDAX_OK;
}

int aidax_stream_activate(aidax_engine* e, int stream)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = check_stream(e, stream);
    if (rc) return rc;
    StreamHost& sh = e->streams[stream];
    e->h_cmd[stream] |= C_SET_GAIN_MEMS | (sh.model >= 0 ? C_FIRST_RUN : 0u);
    e->h_cmd_pre_mem[stream] = sh.pre_target;
    e->h_cmd_master_mem[stream] = sh.master_seen;
    e->dirty_cmd[stream] = 1;
    e->any_dirty = true;
    return AIDAX_OK;
}

int aidax_stream_set_model(aidax_engine* e, int stream, int model_id, float old_param1, float old_param2)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    float* dead = nullptr;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        int rc = check_stream(e, stream);
        if (rc) return rc;
        if (model_id >= (int)e->models.size()) return fail(AIDAX_ERR_ARG, "unknown model %d", model_id);
        if (model_id >= 0 && !e->models[model_id].d_weights) return fail(AIDAX_ERR_STATE, "model %d was released", model_id);
        StreamHost& sh = e->streams[stream];
        const int old_model = sh.model;
        sh.model = model_id < 0 ? -1 : model_id;
        if (sh.model >= 0) e->models[sh.model].refs++;
        if (old_model >= 0 && --e->models[old_model].refs == 0) dead = retire_model_locked(e, old_model);
        if (model_id >= 0) {
            const aidax_model_desc& d = e->models[model_id].host.desc;
            e->h_lin_denom[stream] = 0.1f * d.samplerate;
            e->h_p_init[stream] = old_param1;
            e->h_p_init[(size_t)e->cap + stream] = old_param2;
            e->h_cmd[stream] |= C_SET_MODEL;
            e->dirty_cmd[stream] = 1;
            // the new model's smoother targets, as setTargetValue would leave them (ValueSmoother.hpp:209-216)
            sh.param_target[0] = float_not_equal(0.f, old_param1) ? old_param1 : 0.f;
            sh.param_target[1] = float_not_equal(0.f, old_param2) ? old_param2 : 0.f;
        }
        // F_NET depends on the binding
        if (sh.have_ctrl) decode_controls(sh, sh.ctrl, sh.loading, sh.model >= 0);
        else {
            float c[AIDAX_N_CONTROLS];
            std::memcpy(c, sh.ctrl, sizeof(c));
            const float pt = sh.pre_target, mt = sh.master_target;
            decode_controls(sh, c, sh.loading, sh.model >= 0);
            sh.have_ctrl = false; sh.pre_target = pt; sh.master_target = mt;
        }
        sync_stream_shadow(e, stream);
        e->any_dirty = true;
        e->membership_epoch++;
    }
    if (dead) {
        cudaSetDevice(e->device);
        cudaFree(dead);
    }
    return AIDAX_OK;
}

int aidax_stream_set_controls(aidax_engine* e, int stream, const float controls[AIDAX_N_CONTROLS], int loading)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    if (!controls) return fail(AIDAX_ERR_ARG, "null controls");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = check_stream(e, stream);
    if (rc) return rc;
    StreamHost& sh = e->streams[stream];
    const uint32_t redesigned = decode_controls(sh, controls, loading != 0, sh.model >= 0);
    if (redesigned) write_coefs(e, stream, redesigned);
    sync_stream_shadow(e, stream);
    return AIDAX_OK;
}

int aidax_streams_set_params(aidax_engine* e, int first_stream, int n, const float* p1, const float* p2)
{
    if (!e || n < 0 || first_stream < 0 || first_stream + n > e->n_open_hi) return fail(AIDAX_ERR_ARG, "bad stream range");
    std::lock_guard<std::mutex> lk(e->mu);
    for (int i = 0; i < n; ++i) {
        const int s = first_stream + i;
        StreamHost& sh = e->streams[s];
        if (p1) { sh.ctrl[AIDAX_CTRL_PARAM1] = p1[i]; e->h_p_port[s] = p1[i]; }
        if (p2) { sh.ctrl[AIDAX_CTRL_PARAM2] = p2[i]; e->h_p_port[(size_t)e->cap + s] = p2[i]; }
        e->dirty_ctrl[s] = 1;
        note_param_tracking(e, s);
    }
    e->any_dirty = true;
    return AIDAX_OK;
}

int aidax_stream_get_params(aidax_engine* e, int stream, float* target1, float* target2)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = check_stream(e, stream);
    if (rc) return rc;
    const StreamHost& sh = e->streams[stream];
    if (target1) *target1 = sh.model >= 0 ? sh.param_target[0] : 0.f;
    if (target2) *target2 = sh.model >= 0 ? sh.param_target[1] : 0.f;
    return AIDAX_OK;
}

int aidax_stream_get_state(aidax_engine* e, int stream, float* h, float* c, double* biquad_z)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = check_stream(e, stream);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    const StreamHost& sh = e->streams[stream];
    const size_t cap = (size_t)e->cap;
    if (sh.model >= 0) {
        const int H = e->models[sh.model].host.desc.hidden;
        if (h) CUDA_TRY(cudaMemcpy2D(h, sizeof(float), e->d_rnn.d + stream, cap * sizeof(float), sizeof(float), H, cudaMemcpyDeviceToHost));
        if (c) CUDA_TRY(cudaMemcpy2D(c, sizeof(float), e->d_rnn.d + (size_t)kMaxHidden * cap + stream, cap * sizeof(float), sizeof(float), H, cudaMemcpyDeviceToHost));
    }
    if (biquad_z)
        CUDA_TRY(cudaMemcpy2D(biquad_z, sizeof(double), e->d_bq_z.d + stream, cap * sizeof(double), sizeof(double), kNumBiquads * 2, cudaMemcpyDeviceToHost));
    return AIDAX_OK;
}

int aidax_stream_set_state(aidax_engine* e, int stream, const float* h, const float* c, const double* biquad_z)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = check_stream(e, stream);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    const StreamHost& sh = e->streams[stream];
    const size_t cap = (size_t)e->cap;
    if (sh.model >= 0) {
        const int H = e->models[sh.model].host.desc.hidden;
        if (h) CUDA_TRY(cudaMemcpy2D(e->d_rnn.d + stream, cap * sizeof(float), h, sizeof(float), sizeof(float), H, cudaMemcpyHostToDevice));
        if (c) CUDA_TRY(cudaMemcpy2D(e->d_rnn.d + (size_t)kMaxHidden * cap + stream, cap * sizeof(float), c, sizeof(float), sizeof(float), H, cudaMemcpyHostToDevice));
    }
    if (biquad_z)
        CUDA_TRY(cudaMemcpy2D(e->d_bq_z.d + stream, cap * sizeof(double), biquad_z, sizeof(double), sizeof(double), kNumBiquads * 2, cudaMemcpyHostToDevice));
    return AIDAX_OK;
}

static int render_common_checks(aidax_engine* e, int first, int n, int n_samples, const void* in, void* out, int layout)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    if (first < 0 || n < 0 || first + n > e->cap) return fail(AIDAX_ERR_ARG, "bad stream range [%d, %d)", first, first + n);
    if (n_samples < 0 || n_samples > e->max_block) return fail(AIDAX_ERR_CAPACITY, "n_samples %d exceeds max_block %d", n_samples, e->max_block);
    if (n_samples > 0 && n > 0 && (!in || !out)) return fail(AIDAX_ERR_ARG, "null audio buffer");
    if (layout != AIDAX_LAYOUT_PLANAR && layout != AIDAX_LAYOUT_INTERLEAVED) return fail(AIDAX_ERR_ARG, "unknown layout %d", layout);
    return AIDAX_OK;
}

static void after_render(aidax_engine* e, int first, int n, int n_samples)
{
    const int last = std::min(first + n, e->n_open_hi);
    size_t keep = 0;
    for (size_t i = 0; i < e->track_list.size(); ++i) {
        const int s = e->track_list[i];
        if (s < first || s >= last) { e->track_list[keep++] = s; continue; }
        if (e->streams[s].open) track_run_targets(e->streams[s], n_samples);
        e->track_flag[s] = 0; // any later change of the ports, flags or model re-lists the stream
    }
    e->track_list.resize(keep);
}

int aidax_render_device(aidax_engine* e, int first_stream, int n_streams, int n_samples, const float* d_in, float* d_out,
                        int layout, size_t stride)
{
    int rc = render_common_checks(e, first_stream, n_streams, n_samples, d_in, d_out, layout);
    if (rc) return rc;
    if (n_samples == 0 || n_streams == 0) return AIDAX_OK; // run() with 0 samples is a legal no-op for audio (:606-609)
    std::lock_guard<std::mutex> lk(e->mu);
    CUDA_TRY(cudaSetDevice(e->device));
    rc = launch_all(e, e->stream, first_stream, n_streams, n_samples, d_in, d_out, layout, stride);
    if (rc) return rc;
    after_render(e, first_stream, n_streams, n_samples);
    return AIDAX_OK;
}

// One stream range through the copy-in stream, a kernel stream and the copy-out stream; returns without waiting.
// (`serial`: a small lone range runs copy, kernel and copy back to back on the engine stream -- no cross-stream events.)
static int submit_range(aidax_engine* e, int first_stream, int n_streams, int n_samples, const float* in, float* out, int layout, size_t stride,
                        bool serial = false)
{
    const size_t idx = e->pending.size();
    while (e->ev_in.size() <= idx) {
        cudaEvent_t a, b, c;
        const unsigned fl = e->trace ? cudaEventDefault : cudaEventDisableTiming;
        CUDA_TRY(cudaEventCreateWithFlags(&a, fl));
        CUDA_TRY(cudaEventCreateWithFlags(&b, fl));
        CUDA_TRY(cudaEventCreateWithFlags(&c, fl));
        e->ev_in.push_back(a);
        e->ev_done.push_back(b);
        e->ev_out.push_back(c);
    }
    cudaStream_t ks = serial ? e->stream : e->kstream[idx % (size_t)e->n_kstreams];
    cudaStream_t s_in = serial ? e->stream : e->stream_in, s_out = serial ? e->stream : e->stream_out;
    if (idx == 0 && !serial) {
        // work enqueued on the engine stream by aidax_render_device stays ordered before this period
        CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));
        for (int k = 0; k < e->n_kstreams; ++k)
            if (e->kstream[k] != e->stream) CUDA_TRY(cudaStreamWaitEvent(e->kstream[k], e->ev_fork, 0));
    }
    const size_t rows = layout == AIDAX_LAYOUT_PLANAR ? (size_t)n_streams : (size_t)n_samples;
    const size_t cols = layout == AIDAX_LAYOUT_PLANAR ? (size_t)n_samples : (size_t)n_streams;
    const size_t doff = (size_t)first_stream * (size_t)e->max_block; // every stream owns max_block samples of the device arena
    if (stride == cols) {
        CUDA_TRY(cudaMemcpyAsync(e->d_io_in.d + doff, in, rows * cols * sizeof(float), cudaMemcpyHostToDevice, s_in));
    } else {
        CUDA_TRY(cudaMemcpy2DAsync(e->d_io_in.d + doff, cols * sizeof(float), in, stride * sizeof(float), cols * sizeof(float), rows,
                                   cudaMemcpyHostToDevice, s_in));
    }
    // A range that holds several kernel families (heterogeneous models) and is one of many hand-over ranges of a period
    // is NOT rendered on its own: cutting the period into ranges cuts every family into pieces too small for the
    // throughput shapes (C4 through the plugins: 14 families of 1-2 k streams per range ran the FP32 latency shapes,
    // 5.5 ms per period against 1.8 ms device-resident).  Its input and controls go to the device now, its kernels are
    // launched per family over all such ranges at once when the period is complete (wait_ranges).
    bool defer = false;
    if (!serial && e->defer_hetero && layout == AIDAX_LAYOUT_PLANAR && n_samples == e->max_block && stride == cols) {
        RangePlan* rp = nullptr;
        int rcp = build_plan(e, first_stream, n_streams, &rp);
        if (rcp) return rcp;
        defer = rp->launches.size() > 1;
    }
    if (defer) {
        int rcu = upload_dirty(e, first_stream, n_streams, s_in);
        if (rcu) return rcu;
        CUDA_TRY(cudaEventRecord(e->ev_in[idx], s_in));
        e->pending.push_back(aidax_engine::PendingRange{first_stream, n_streams, n_samples, true, out, stride, layout});
        return AIDAX_OK;
    }
    if (!serial) {
        CUDA_TRY(cudaEventRecord(e->ev_in[idx], s_in));
        CUDA_TRY(cudaStreamWaitEvent(ks, e->ev_in[idx], 0));
    }
    const int rc = launch_all(e, ks, first_stream, n_streams, n_samples, e->d_io_in.d + doff, e->d_io_out.d + doff, layout, cols);
    if (rc) return rc;
    if (!serial) {
        CUDA_TRY(cudaEventRecord(e->ev_done[idx], ks));
        CUDA_TRY(cudaStreamWaitEvent(s_out, e->ev_done[idx], 0));
    }
    if (stride == cols) {
        CUDA_TRY(cudaMemcpyAsync(out, e->d_io_out.d + doff, rows * cols * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    } else {
        CUDA_TRY(cudaMemcpy2DAsync(out, stride * sizeof(float), e->d_io_out.d + doff, cols * sizeof(float), cols * sizeof(float), rows,
                                   cudaMemcpyDeviceToHost, s_out));
    }
    if (e->trace && !serial) CUDA_TRY(cudaEventRecord(e->ev_out[idx], s_out));
    e->pending.push_back(aidax_engine::PendingRange{first_stream, n_streams, n_samples, false, nullptr, 0, layout});
    return AIDAX_OK;
}

// Kernels and copy-out of the ranges submit_range deferred: runs of adjacent ranges are rendered as one range each, one
// launch per kernel family.  The families are pipelined against the copies on both sides: a family's kernel waits only
// for the copy-in of the ranges that hold its streams (the ranges of a period arrive one after another while the host
// is still handing over later ones), and its streams are copied out as soon as that kernel is done, while the other
// families are still running -- instead of copy-in of everything -> all kernels -> copy-out of everything.
static int flush_deferred(aidax_engine* e)
{
    std::vector<size_t> order;
    for (size_t i = 0; i < e->pending.size(); ++i) if (e->pending[i].deferred) order.push_back(i);
    if (order.empty()) return AIDAX_OK;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return e->pending[a].first < e->pending[b].first; });
    for (size_t a = 0; a < order.size();) {
        size_t b = a + 1;
        const aidax_engine::PendingRange& r0 = e->pending[order[a]];
        int end = r0.first + r0.n;
        while (b < order.size() && e->pending[order[b]].first == end && e->pending[order[b]].n_samples == r0.n_samples) {
            end += e->pending[order[b]].n;
            ++b;
        }
        const size_t doff = (size_t)r0.first * (size_t)e->max_block;
        const size_t row = (size_t)r0.n_samples;
        bool per_family = e->pipeline_families;
        LaunchHooks hooks;
        hooks.before = [&](const LaunchPlan& lp, cudaStream_t ls) -> int {
            for (size_t k = a; k < b; ++k) {
                const aidax_engine::PendingRange& r = e->pending[order[k]];
                bool touches = !per_family;
                for (size_t i = 0; i < lp.runs.size() && !touches; ++i) touches = lp.runs[i].first < r.first + r.n && lp.runs[i].second > r.first;
                if (touches) CUDA_TRY(cudaStreamWaitEvent(ls, e->ev_in[order[k]], 0));
            }
            return AIDAX_OK;
        };
        hooks.after = [&](const LaunchPlan& lp, cudaStream_t ls) -> int {
            if (!per_family) return AIDAX_OK;
            // The family's streams go home on the family's OWN kernel stream, right behind its kernel: on one shared
            // copy-out stream the copies would queue in launch order, and the launches go out largest-footprint first
            // (build_plan), i.e. slowest first -- every small family's output would wait for the LSTM-40 kernel.
            for (const auto& run : lp.runs)
                for (size_t k = a; k < b; ++k) { // a run of streams may cross hand-over ranges: their host rows need not be adjacent
                    const aidax_engine::PendingRange& r = e->pending[order[k]];
                    const int lo = std::max(run.first, r.first), hi = std::min(run.second, r.first + r.n);
                    if (hi <= lo) continue;
                    CUDA_TRY(cudaMemcpyAsync(r.out + (size_t)(lo - r.first) * row, e->d_io_out.d + (size_t)lo * (size_t)e->max_block, (size_t)(hi - lo) * row * sizeof(float),
                                             cudaMemcpyDeviceToHost, ls));
                }
            return AIDAX_OK;
        };
        {
            // a period cut into very many stream runs (models interleaved stream by stream) is cheaper to copy out whole
            RangePlan* rp = nullptr;
            const int rcp = build_plan(e, r0.first, end - r0.first, &rp);
            if (rcp) return rcp;
            size_t n_runs = 0;
            for (const LaunchPlan& lp : rp->launches) n_runs += lp.runs.size();
            if (n_runs > 512) per_family = false;
        }
        const int rc = launch_all(e, e->stream, r0.first, end - r0.first, r0.n_samples, e->d_io_in.d + doff, e->d_io_out.d + doff, r0.layout,
                                  (size_t)r0.n_samples, &hooks);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(e->ev_done[order[a]], e->stream));
        if (!per_family) CUDA_TRY(cudaStreamWaitEvent(e->stream_out, e->ev_done[order[a]], 0));
        for (size_t k = a; k < b; ++k) {
            const aidax_engine::PendingRange& r = e->pending[order[k]];
            const size_t off = (size_t)r.first * (size_t)e->max_block;
            if (!per_family)
                CUDA_TRY(cudaMemcpyAsync(r.out, e->d_io_out.d + off, (size_t)r.n * (size_t)r.n_samples * sizeof(float), cudaMemcpyDeviceToHost,
                                         e->stream_out));
            if (k != a) CUDA_TRY(cudaEventRecord(e->ev_done[order[k]], e->stream));
            if (e->trace) CUDA_TRY(cudaEventRecord(e->ev_out[order[k]], e->stream_out));
        }
        a = b;
    }
    return AIDAX_OK;
}

static int wait_ranges(aidax_engine* e)
{
    if (e->pending.empty()) return AIDAX_OK;
    const int rc_flush = flush_deferred(e);
    // every D2H copy is on stream_out and waits for its kernel, so this covers the kernel streams too;
    // the engine stream is re-joined so that later aidax_render_device calls stay ordered after this period
    cudaError_t err = sync_kernel_streams(e);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream_out);
    if (e->trace && err == cudaSuccess && e->pending.size() > 1 && e->trace_skip > 0) --e->trace_skip;
    else if (e->trace && err == cudaSuccess && e->pending.size() > 1 && e->trace_left > 0) {
        --e->trace_left;
        fprintf(stderr, "aidax trace: %zu ranges (ms after the first submit: copy-in done, kernels done, copy-out done)\n", e->pending.size());
        for (size_t i = 0; i < e->pending.size(); ++i) {
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, e->ev_fork, e->ev_in[i]);
            cudaEventElapsedTime(&b, e->ev_fork, e->ev_done[i]);
            cudaEventElapsedTime(&c, e->ev_fork, e->ev_out[i]);
            fprintf(stderr, "  range %2zu [%7d +%6d]  in %7.3f  kernel %7.3f  out %7.3f\n", i, e->pending[i].first, e->pending[i].n, a, b, c);
        }
    }
    for (const aidax_engine::PendingRange& r : e->pending) after_render(e, r.first, r.n, r.n_samples);
    e->pending.clear();
    CUDA_TRY(err);
    return rc_flush;
}

static int render_host_checks(aidax_engine* e, int first_stream, int n_streams, int n_samples, const float* in, float* out, int layout, size_t stride)
{
    int rc = render_common_checks(e, first_stream, n_streams, n_samples, in, out, layout);
    if (rc) return rc;
    // Staged through the engine's device arena with a dense pitch (planar: n_samples, interleaved: n_streams)
    if (layout == AIDAX_LAYOUT_PLANAR ? stride < (size_t)n_samples : stride < (size_t)n_streams)
        return fail(AIDAX_ERR_ARG, "stride %zu smaller than a row", stride);
    return AIDAX_OK;
}

int aidax_render_host(aidax_engine* e, int first_stream, int n_streams, int n_samples, const float* in, float* out,
                      int layout, size_t stride)
{
    int rc = render_host_checks(e, first_stream, n_streams, n_samples, in, out, layout, stride);
    if (rc) return rc;
    if (n_samples == 0 || n_streams == 0) return AIDAX_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    CUDA_TRY(cudaSetDevice(e->device));
    // Large periods are cut into stream chunks that flow through the copy-in, kernel and copy-out streams, so that on a
    // PCIe-attached GPU the two copy directions and the kernels overlap; a small period is one chunk.
    const size_t bytes = (size_t)n_streams * (size_t)n_samples * sizeof(float);
    int n_chunks = (int)std::min<size_t>(8, std::max<size_t>(1, bytes / ((size_t)8 << 20)));
    int chunk = (n_streams + n_chunks - 1) / n_chunks;
    chunk = (chunk + 1023) & ~1023; // whole tiles of every kernel shape
    n_chunks = (n_streams + chunk - 1) / chunk;
    const bool serial = n_chunks == 1 && e->pending.empty();
    for (int c = 0; c < n_chunks && rc == AIDAX_OK; ++c) {
        const int c0 = c * chunk, cn = std::min(chunk, n_streams - c0);
        const size_t hoff = layout == AIDAX_LAYOUT_PLANAR ? (size_t)c0 * stride : (size_t)c0; // rows resp. columns of the caller's block
        rc = submit_range(e, first_stream + c0, cn, n_samples, in + hoff, out + hoff, layout, stride, serial);
    }
    const int rc_wait = wait_ranges(e);
    return rc ? rc : rc_wait;
}

int aidax_render_host_submit(aidax_engine* e, int first_stream, int n_streams, int n_samples, const float* in, float* out,
                             int layout, size_t stride)
{
    int rc = render_host_checks(e, first_stream, n_streams, n_samples, in, out, layout, stride);
    if (rc) return rc;
    if (n_samples == 0 || n_streams == 0) return AIDAX_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    CUDA_TRY(cudaSetDevice(e->device));
    return submit_range(e, first_stream, n_streams, n_samples, in, out, layout, stride);
}

int aidax_render_host_submit_graded(aidax_engine* e, int first_stream, int n_streams, int n_samples, const float* in, float* out, size_t stride)
{
    int rc = render_host_checks(e, first_stream, n_streams, n_samples, in, out, AIDAX_LAYOUT_PLANAR, stride);
    if (rc) return rc;
    if (n_samples == 0 || n_streams == 0) return AIDAX_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    CUDA_TRY(cudaSetDevice(e->device));
    // A period that is complete when it is handed over is bound by the copies: its duration is the copy-in of the first
    // piece (nothing else can run yet) + the remaining copy-in + kernel and copy-out of the LAST piece (nothing overlaps
    // them).  So the pieces are small at both ends and large in the middle, where a piece only has to keep both copy
    // engines and the SMs busy: 4096, 8192, 16384 ... 16384, 8192, 4096 streams (4096 = the smallest batch that still takes
    // the tensor-core shape).
    const int unit = e->grade_unit, mid = e->grade_mid;
    int done = 0;
    std::vector<int> tail;
    int left = n_streams;
    std::vector<int> pieces;
    if (left > 2 * unit) { pieces.push_back(unit); left -= unit; }
    if (left > 6 * unit) { pieces.push_back(2 * unit); left -= 2 * unit; }
    if (left > 2 * unit) { tail.push_back(unit); left -= unit; }
    if (left > 6 * unit) { tail.push_back(2 * unit); left -= 2 * unit; }
    while (left > 0) {
        const int take = left >= 2 * mid ? mid : left; // the last middle piece takes the remainder (< 2 mid)
        pieces.push_back(take);
        left -= take;
    }
    for (size_t i = tail.size(); i-- > 0;) pieces.push_back(tail[i]);
    for (int take : pieces) {
        rc = submit_range(e, first_stream + done, take, n_samples, in + (size_t)done * stride, out + (size_t)done * stride, AIDAX_LAYOUT_PLANAR, stride);
        if (rc) return rc;
        done += take;
    }
    return AIDAX_OK;
}

int aidax_render_host_wait(aidax_engine* e)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    CUDA_TRY(cudaSetDevice(e->device));
    return wait_ranges(e);
}

int aidax_engine_synchronize(aidax_engine* e)
{
    if (!e) return fail(AIDAX_ERR_ARG, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return AIDAX_OK;
}

void* aidax_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        fail(AIDAX_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void aidax_host_free(void* p) { if (p) cudaFreeHost(p); }

int aidax_measure_fp32_peak(int device, int millis, double* tflops_out)
{
    if (!tflops_out) return fail(AIDAX_ERR_ARG, "null output");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    float* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, sizeof(float)));
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    const int grid = prop.multiProcessorCount * 8, iters = 8192;
    const double flop_per_launch = (double)grid * 256.0 * 16.0 * 2.0 * iters;
    fp32_peak_kernel<<<grid, 256>>>(d, iters); // warm-up
    CUDA_TRY(cudaDeviceSynchronize());
    double best = 0.0, spent = 0.0;
    while (spent < (double)millis) {
        CUDA_TRY(cudaEventRecord(a));
        fp32_peak_kernel<<<grid, 256>>>(d, iters);
        CUDA_TRY(cudaEventRecord(b));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        spent += ms;
        best = std::max(best, flop_per_launch / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops_out = best;
    return AIDAX_OK;
}

int aidax_measure_pcie(int device, size_t bytes, double gbs_out[3])
{
    if (!gbs_out || bytes < ((size_t)1 << 20)) return fail(AIDAX_ERR_ARG, "bad measure_pcie arguments");
    CUDA_TRY(cudaSetDevice(device));
    float *h_a = nullptr, *h_b = nullptr, *d_a = nullptr, *d_b = nullptr;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t a = nullptr, b = nullptr, c = nullptr;
    int rc = AIDAX_OK;
    auto body = [&]() -> int {
        CUDA_TRY(cudaHostAlloc((void**)&h_a, bytes, cudaHostAllocDefault));
        CUDA_TRY(cudaHostAlloc((void**)&h_b, bytes, cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc((void**)&d_a, bytes));
        CUDA_TRY(cudaMalloc((void**)&d_b, bytes));
        std::memset(h_a, 0, bytes);
        std::memset(h_b, 0, bytes);
        CUDA_TRY(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&a));
        CUDA_TRY(cudaEventCreate(&b));
        CUDA_TRY(cudaEventCreate(&c));
        double best[3] = {0.0, 0.0, 0.0};
        for (int it = 0; it < 4; ++it) {
            float ms = 0.f;
            CUDA_TRY(cudaEventRecord(a, s0));
            CUDA_TRY(cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s0));
            CUDA_TRY(cudaEventRecord(b, s0));
            CUDA_TRY(cudaEventSynchronize(b));
            CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
            best[0] = std::max(best[0], (double)bytes / (ms * 1e-3) / 1e9);
            CUDA_TRY(cudaEventRecord(a, s0));
            CUDA_TRY(cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s0));
            CUDA_TRY(cudaEventRecord(b, s0));
            CUDA_TRY(cudaEventSynchronize(b));
            CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
            best[1] = std::max(best[1], (double)bytes / (ms * 1e-3) / 1e9);
            // both directions at once: the slower direction's rate while the other one is busy
            CUDA_TRY(cudaEventRecord(a, s0));
            CUDA_TRY(cudaStreamWaitEvent(s1, a, 0));
            CUDA_TRY(cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s0));
            CUDA_TRY(cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s1));
            CUDA_TRY(cudaEventRecord(c, s1));
            CUDA_TRY(cudaStreamWaitEvent(s0, c, 0));
            CUDA_TRY(cudaEventRecord(b, s0));
            CUDA_TRY(cudaEventSynchronize(b));
            CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
            best[2] = std::max(best[2], (double)bytes / (ms * 1e-3) / 1e9);
        }
        for (int i = 0; i < 3; ++i) gbs_out[i] = best[i];
        return AIDAX_OK;
    };
    rc = body();
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    if (c) cudaEventDestroy(c);
    if (s0) cudaStreamDestroy(s0);
    if (s1) cudaStreamDestroy(s1);
    if (d_a) cudaFree(d_a);
    if (d_b) cudaFree(d_b);
    if (h_a) cudaFreeHost(h_a);
    if (h_b) cudaFreeHost(h_b);
    return rc;
}

uint64_t aidax_engine_launch_count(aidax_engine* e) { return e ? e->launches : 0; }

int aidax_engine_describe_plan(aidax_engine* e, int first_stream, int n_streams, char* out, int cap)
{
    if (!e || !out || cap < 1) return fail(AIDAX_ERR_ARG, "bad describe_plan arguments");
    std::lock_guard<std::mutex> lk(e->mu);
    if (first_stream < 0 || n_streams < 0 || first_stream + n_streams > e->cap) return fail(AIDAX_ERR_ARG, "bad stream range");
    CUDA_TRY(cudaSetDevice(e->device));
    RangePlan* rp = nullptr;
    const int rc = build_plan(e, first_stream, n_streams, &rp);
    if (rc) return rc;
    static const char* const kKind[] = {"fp32", "tc1", "warp", "tc2"};
    std::string text;
    for (const LaunchPlan& lp : rp->launches) {
        const KernelVariant* kv = lp.kv;
        int streams = 0;
        for (const auto& r : lp.runs) streams += r.second - r.first;
        const int cs = kv->cs ? kv->cs : kv->ws * kv->NW;
        char line[160];
        snprintf(line, sizeof(line), "%s H=%d shape=%s streams=%d ctas=%d streams_per_cta=%d threads=%d\n",
                 kv->cell == 0 ? "lstm" : kv->cell == 1 ? "gru" : "none", kv->cell == 2 ? 0 : kv->H,
                 (kv->kind == 3 && (kv->S & 4)) ? "tc3" : kKind[kv->kind & 3], streams, lp.n_blocks, cs, kv->NW * 32);
        text += line;
    }
    snprintf(out, (size_t)cap, "%s", text.c_str());
    return AIDAX_OK;
}

int aidax_engine_pci_bus_id(aidax_engine* e, char* out, int cap)
{
    if (!e || !out || cap < 13) return fail(AIDAX_ERR_ARG, "bad pci_bus_id arguments");
    CUDA_TRY(cudaDeviceGetPCIBusId(out, cap, e->device));
    for (char* c = out; *c; ++c) *c = (char)tolower(*c); // sysfs spells bus ids in lower case
    return AIDAX_OK;
}

} // extern "C"
