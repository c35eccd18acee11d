// plugin.cpp -- rt-neural-generic.so: the reference's LV2 surface on top of the CUDA engine.
//
// Same URI, same 25-port table, same worker/state extension data, same patch:Set model loading
// and notify message as AidaDSP/aidadsp-lv2 rt-neural-generic (rt-neural-generic.cpp:11-29,
// 244-333, 376-480, 484-660, 681-893; uris.h:26-94; ttl/rt-neural-generic.ttl) -- but an instance
// holds no DSP: it is a handle on one stream of a process-wide aidax engine, and run() either
// renders that single stream synchronously (plain LV2 hosts) or, inside a batch opened through the
// #batch extension (include/aidax_batch.h), defers to one launch for all instances.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/aidax.h"
#include "../../include/aidax_batch.h"
#include "../../include/aidax_lv2_abi.h"

namespace {

// ---------------------------------------------------------------------------- port table
// ports_t, rt-neural-generic.h:84-112 (AIDADSP_MODEL_LOADER=1, AIDADSP_PARAMS=2, optional dc blocker)
enum Port : uint32_t {
    P_IN = 0, P_OUT_1, P_CONTROL, P_NOTIFY, P_IN_LPF, P_PREGAIN, P_NET_BYPASS, P_PARAM1, P_PARAM2, P_EQ_BYPASS, P_EQ_POS,
    P_BASS, P_BFREQ, P_MID, P_MFREQ, P_MIDQ, P_MTYPE, P_TREBLE, P_TFREQ, P_DEPTH, P_PRESENCE, P_DCBLOCKER, P_MASTER,
    P_INPUT_SIZE, P_ENABLED, P_COUNT
};
static_assert((int)P_IN_LPF == (int)AIDAX_CTRL_PORT_BASE && (int)P_COUNT == (int)AIDAX_CTRL_PORT_BASE + (int)AIDAX_N_CONTROLS, "port table drifted");

#define PLUGIN__json AIDAX_PLUGIN_URI "#json"
#define PLUGIN__applyJson AIDAX_PLUGIN_URI "#applyJson"

struct Uris { // uris.h:33-67
    LV2_URID atom_Float, atom_Path, atom_Resource, atom_Sequence, atom_URID, atom_eventTransfer, applyJson, json,
        midi_Event, param_gain, patch_Get, patch_Set, patch_property, patch_value;
};

// worker messages, rt-neural-generic.h:132-157
enum WorkerMessageType { kWorkerLoad, kWorkerApply, kWorkerFree };
struct WorkerMessage { WorkerMessageType type; };
struct WorkerLoadMessage { WorkerMessageType type; char path[1024]; };
struct LoadedModel { int model_id; int input_size; char* path; float old_param1, old_param2; };
struct WorkerApplyMessage { WorkerMessageType type; LoadedModel* model; };

// ---------------------------------------------------------------------------- process-wide engine
struct EngineHub {
    std::mutex mu;
    aidax_engine* engine = nullptr;
    int max_instances = 0, max_block = 0, device = -1;
    int refs = 0;
    // batch state: run() parks its ports in the slot of its own stream (no lock: every instance owns one
    // slot, begin()/end() are called by the host while no run() is in flight), end() walks the slots in order
    std::atomic<bool> batching{false};
    struct Pending { int stream; const float* in; float* out; uint32_t n; };
    struct Slot { const float* in; float* out; uint32_t n; uint32_t epoch; };
    std::vector<Slot> slots;
    uint32_t epoch = 0;
    // Streaming hand-over: stream ids are cut into fixed chunks; the run() that parks the last live instance of a
    // chunk submits that chunk's copy-in / kernels / copy-out at once (aidax_render_host_submit), so the transfers
    // of early chunks overlap the run() calls of later ones.  end() renders what is left and waits.
    // Counting is two-level so that run() calls on different host threads do not fight over one cache line: a run()
    // bumps the counter of its 64-stream block (a host thread works through neighbouring instances, so a block's
    // line stays in its cache); the run() that fills a block bumps the chunk's counter.
    static constexpr int kBlockStreams = 64;
    int chunk_streams = 16384;
    struct alignas(64) Block {
        std::atomic<int> parked{0};   // instances of this block that ran in the current batch
        std::atomic<int> live{0};     // instances whose stream id lies in this block
    };
    struct alignas(64) Chunk {
        std::atomic<int> blocks_done{0};  // blocks of this chunk whose live instances have all run in the current batch
        std::atomic<int> blocks_live{0};  // blocks with at least one live instance
        std::atomic<uint32_t> submitted_epoch{0};
    };
    std::unique_ptr<Block[]> blocks;
    std::unique_ptr<Chunk[]> chunks;
    int n_blocks = 0, n_chunks = 0;
    std::atomic<int> submit_error{0};
};
EngineHub g_hub;

int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Creates the engine on first use.  Sizes: configure() > environment > defaults.
aidax_engine* hub_engine()
{
    std::lock_guard<std::mutex> lk(g_hub.mu);
    if (g_hub.engine) return g_hub.engine;
    if (g_hub.max_instances <= 0) g_hub.max_instances = env_int("AIDAX_MAX_STREAMS", 1024);
    if (g_hub.max_block <= 0) g_hub.max_block = env_int("AIDAX_MAX_BLOCK", 2048);
    if (g_hub.device < 0) g_hub.device = env_int("AIDAX_DEVICE", 0);
    if (aidax_engine_create(g_hub.device, g_hub.max_instances, g_hub.max_block, &g_hub.engine) != AIDAX_OK) {
        fprintf(stderr, "rt-neural-generic (aidax): cannot create the CUDA engine: %s\n", aidax_last_error());
        g_hub.engine = nullptr;
    } else {
        g_hub.slots.assign((size_t)g_hub.max_instances + 32, EngineHub::Slot{nullptr, nullptr, 0u, 0u});
        g_hub.chunk_streams = std::max(128, env_int("AIDAX_BATCH_CHUNK", 16384)) & ~127; // whole tiles of every kernel shape
        g_hub.n_chunks = (int)((g_hub.slots.size() + g_hub.chunk_streams - 1) / g_hub.chunk_streams);
        g_hub.n_blocks = g_hub.n_chunks * (g_hub.chunk_streams / EngineHub::kBlockStreams);
        g_hub.chunks.reset(new EngineHub::Chunk[g_hub.n_chunks]);
        g_hub.blocks.reset(new EngineHub::Block[g_hub.n_blocks]);
    }
    return g_hub.engine;
}

int batch_configure(int max_instances, int max_block, int device)
{
    std::lock_guard<std::mutex> lk(g_hub.mu);
    if (g_hub.engine)
        return (max_instances <= g_hub.max_instances && max_block <= g_hub.max_block && (device < 0 || device == g_hub.device)) ? 0 : AIDAX_ERR_STATE;
    g_hub.max_instances = max_instances;
    g_hub.max_block = max_block;
    if (device >= 0) g_hub.device = device;
    return 0;
}

int batch_begin()
{
    std::lock_guard<std::mutex> lk(g_hub.mu);
    g_hub.epoch++;
    g_hub.submit_error.store(0, std::memory_order_relaxed);
    for (int c = 0; c < g_hub.n_chunks; ++c) g_hub.chunks[c].blocks_done.store(0, std::memory_order_relaxed);
    for (int b = 0; b < g_hub.n_blocks; ++b) g_hub.blocks[b].parked.store(0, std::memory_order_relaxed);
    g_hub.batching.store(true, std::memory_order_release);
    return 0;
}

// Renders a set of deferred run() calls with as few engine calls as possible: runs of consecutive
// stream ids whose port buffers are rows of one arena (constant stride) go straight from the
// caller's memory; anything else is gathered through pinned staging rows.
int render_pending(std::vector<EngineHub::Pending>& items)
{
    aidax_engine* e = g_hub.engine;
    if (!e || items.empty()) return 0;
    // items arrive sorted by stream id (batch_end walks the slot table in order)
    size_t i = 0;
    while (i < items.size()) {
        const EngineHub::Pending& a = items[i];
        size_t j = i + 1;
        ptrdiff_t stride = 0;
        bool regular = true;
        if (j < items.size() && items[j].stream == a.stream + 1 && items[j].n == a.n) {
            stride = items[j].in - a.in;
            regular = stride >= (ptrdiff_t)a.n && (items[j].out - a.out) == stride;
            if (regular)
                while (j < items.size() && items[j].stream == a.stream + (int)(j - i) && items[j].n == a.n &&
                       items[j].in == a.in + (ptrdiff_t)(j - i) * stride && items[j].out == a.out + (ptrdiff_t)(j - i) * stride)
                    ++j;
            else
                j = i + 1;
        }
        const int count = (int)(j - i);
        int rc;
        if (count > 1) {
            rc = aidax_render_host_submit(e, a.stream, count, (int)a.n, a.in, a.out, AIDAX_LAYOUT_PLANAR, (size_t)stride);
        } else {
            rc = aidax_render_host(e, a.stream, 1, (int)a.n, a.in, a.out, AIDAX_LAYOUT_PLANAR, a.n);
        }
        if (rc) return rc;
        i = j;
    }
    return 0;
}

// Called by the run() that completed chunk c (all of its live instances have parked).  If the parked ports are the
// rows of one arena (the batch host's layout) the chunk goes to the GPU now; anything irregular is left to end().
void submit_chunk(int c)
{
    const int lo = c * g_hub.chunk_streams, hi = std::min<int>(lo + g_hub.chunk_streams, (int)g_hub.slots.size());
    const uint32_t epoch = g_hub.epoch;
    int first = lo;
    while (first < hi && g_hub.slots[first].epoch != epoch) ++first; // closed streams at the edges are fine
    int last = hi;
    while (last > first && g_hub.slots[last - 1].epoch != epoch) --last;
    if (last - first < 2) return;
    const EngineHub::Slot& a = g_hub.slots[first];
    const ptrdiff_t stride = g_hub.slots[first + 1].in - a.in;
    if (stride < (ptrdiff_t)a.n) return;
    for (int s = first; s < last; ++s) {
        const EngineHub::Slot& sl = g_hub.slots[s];
        if (sl.epoch != epoch || sl.n != a.n || sl.in != a.in + (ptrdiff_t)(s - first) * stride || sl.out != a.out + (ptrdiff_t)(s - first) * stride)
            return;
    }
    const int rc = aidax_render_host_submit(g_hub.engine, first, last - first, (int)a.n, a.in, a.out, AIDAX_LAYOUT_PLANAR, (size_t)stride);
    if (rc != AIDAX_OK) {
        g_hub.submit_error.store(rc, std::memory_order_relaxed);
        return;
    }
    g_hub.chunks[c].submitted_epoch.store(epoch, std::memory_order_release);
}

int batch_end()
{
    static std::vector<EngineHub::Pending> items; // reused across periods (the host calls end() from one thread)
    items.clear();
    {
        std::lock_guard<std::mutex> lk(g_hub.mu);
        g_hub.batching.store(false, std::memory_order_release);
        const int n_slots = (int)g_hub.slots.size(), per_chunk = g_hub.chunk_streams / EngineHub::kBlockStreams;
        for (int c = 0; c < g_hub.n_chunks; ++c) {
            if (g_hub.chunks[c].submitted_epoch.load(std::memory_order_acquire) == g_hub.epoch) continue; // already on the GPU
            for (int b = c * per_chunk; b < (c + 1) * per_chunk; ++b) {
                if (g_hub.blocks[b].parked.load(std::memory_order_acquire) == 0) continue;
                const int s1 = std::min(n_slots, (b + 1) * EngineHub::kBlockStreams);
                for (int s = b * EngineHub::kBlockStreams; s < s1; ++s) {
                    const EngineHub::Slot& sl = g_hub.slots[s];
                    if (sl.epoch == g_hub.epoch) items.push_back(EngineHub::Pending{s, sl.in, sl.out, sl.n});
                }
            }
        }
    }
    int rc = render_pending(items);
    const int rc_wait = g_hub.engine ? aidax_render_host_wait(g_hub.engine) : 0;
    if (!rc) rc = rc_wait;
    if (!rc) rc = g_hub.submit_error.load(std::memory_order_relaxed);
    return rc;
}

uint64_t batch_launch_count() { return g_hub.engine ? aidax_engine_launch_count(g_hub.engine) : 0; }

const AidaxBatchInterface g_batch_iface = {
    AIDAX_BATCH_ABI_VERSION, batch_configure, aidax_host_alloc, aidax_host_free, batch_begin, batch_end, batch_launch_count, aidax_last_error,
};

// ---------------------------------------------------------------------------- one plugin instance
// The fields run() touches on an ordinary period (no atom events, no control change) sit in the first three
// cache lines: with hundreds of thousands of instances per period run() is bound by host memory traffic.
struct alignas(64) Plugin {
    // ---- hot
    const float* in = nullptr;
    float* out_1 = nullptr;
    const LV2_Atom_Sequence* control_port = nullptr;
    LV2_Atom_Sequence* notify_port = nullptr;
    float* ctl_row = nullptr;        // set when control ports 4..24 are one contiguous float row (what a batch host connects)
    aidax_engine* engine = nullptr;
    int stream = -1;
    int last_input_size = 0;         // rt-neural-generic.cpp:316
    uint32_t notify_capacity = 0;    // atom.size of the notify port at the start of the last run() (:528-534)
    LV2_URID urid_sequence = 0;
    bool loading = true;             // :318
    bool latched_valid = false;
    bool latched_loading = true;
    bool notify_armed = false;       // a run() has seen both atom ports: work_response may report (the reference's forge.buf != NULL)
    bool forge_ready = false;        // the forge already points at this period's notify sequence
    float latched[AIDAX_N_CONTROLS];
    // ---- cold
    float* ctl[P_COUNT] = {};
    LV2_URID_Map* map = nullptr;
    LV2_Worker_Schedule* schedule = nullptr;
    LV2_Log_Log* log = nullptr;
    LV2_Log_Logger logger{};
    LV2_Atom_Forge forge{};
    LV2_Atom_Forge_Frame notify_frame{};
    Uris uris{};
    double samplerate = 48000.0;
    LoadedModel* model = nullptr;
};

// Points the forge at the notify port and opens the sequence, as run() does at :528-534.
void open_notify_sequence(Plugin* self)
{
    lv2_atom_forge_set_buffer(&self->forge, (uint8_t*)self->notify_port, self->notify_capacity);
    lv2_atom_forge_sequence_head(&self->forge, &self->notify_frame, 0);
    self->forge_ready = true;
}

void map_uris(LV2_URID_Map* map, Uris* u)
{
    u->atom_Float = map->map(map->handle, LV2_ATOM__Float);
    u->atom_Path = map->map(map->handle, LV2_ATOM__Path);
    u->atom_Resource = map->map(map->handle, LV2_ATOM__Resource);
    u->atom_Sequence = map->map(map->handle, LV2_ATOM__Sequence);
    u->atom_URID = map->map(map->handle, LV2_ATOM__URID);
    u->atom_eventTransfer = map->map(map->handle, LV2_ATOM__eventTransfer);
    u->applyJson = map->map(map->handle, PLUGIN__applyJson);
    u->json = map->map(map->handle, PLUGIN__json);
    u->midi_Event = map->map(map->handle, LV2_MIDI__MidiEvent);
    u->param_gain = map->map(map->handle, LV2_PARAMETERS__gain);
    u->patch_Get = map->map(map->handle, LV2_PATCH__Get);
    u->patch_Set = map->map(map->handle, LV2_PATCH__Set);
    u->patch_property = map->map(map->handle, LV2_PATCH__property);
    u->patch_value = map->map(map->handle, LV2_PATCH__value);
}

void free_loaded(LoadedModel* m)
{
    if (!m) return;
    free(m->path);
    delete m;
}

// ---------------------------------------------------------------------------- LV2 callbacks

// rt-neural-generic.cpp:244-333
LV2_Handle instantiate(const LV2_Descriptor*, double samplerate, const char*, const LV2_Feature* const* features)
{
    Plugin* self = new Plugin();
    self->samplerate = samplerate;
    for (int i = 0; features && features[i]; ++i) {
        if (!strcmp(features[i]->URI, LV2_URID__map)) self->map = (LV2_URID_Map*)features[i]->data;
        else if (!strcmp(features[i]->URI, LV2_WORKER__schedule)) self->schedule = (LV2_Worker_Schedule*)features[i]->data;
        else if (!strcmp(features[i]->URI, LV2_LOG__log)) self->log = (LV2_Log_Log*)features[i]->data;
    }
    if (!self->map) {
        fprintf(stderr, "Error! Missing feature urid:map %s %d\n", __func__, __LINE__);
        delete self;
        return nullptr;
    }
    if (!self->schedule) {
        fprintf(stderr, "Error! Missing feature work:schedule %s %d\n", __func__, __LINE__);
        delete self;
        return nullptr;
    }
    map_uris(self->map, &self->uris);
    lv2_log_logger_init(&self->logger, self->map, self->log);
    lv2_atom_forge_init(&self->forge, self->map);
    self->urid_sequence = self->uris.atom_Sequence;

    self->engine = hub_engine();
    if (!self->engine || aidax_stream_open(self->engine, samplerate, &self->stream) != AIDAX_OK) {
        // No CPU fallback: without the CUDA engine the plugin cannot exist.
        lv2_log_error(&self->logger, "aidax engine unavailable: %s\n", aidax_last_error());
        delete self;
        return nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(g_hub.mu);
        g_hub.refs++;
        if (g_hub.blocks[self->stream / EngineHub::kBlockStreams].live.fetch_add(1) == 0)
            g_hub.chunks[self->stream / g_hub.chunk_streams].blocks_live.fetch_add(1);
    }
    self->loading = true;
    self->last_input_size = 0;
    self->model = nullptr;
    return (LV2_Handle)self;
}

// rt-neural-generic.cpp:376-480
void connect_port(LV2_Handle instance, uint32_t port, void* data)
{
    Plugin* self = (Plugin*)instance;
    switch (port) {
    case P_IN: self->in = (const float*)data; break;
    case P_OUT_1: self->out_1 = (float*)data; break;
    case P_CONTROL: self->control_port = (const LV2_Atom_Sequence*)data; break;
    case P_NOTIFY: self->notify_port = (LV2_Atom_Sequence*)data; break;
    default:
        if (port < P_COUNT) {
            self->ctl[port] = (float*)data;
            bool row = self->ctl[P_IN_LPF] != nullptr;
            for (uint32_t q = P_IN_LPF; row && q < P_COUNT; ++q) row = self->ctl[q] == self->ctl[P_IN_LPF] + (q - P_IN_LPF);
            self->ctl_row = row ? self->ctl[P_IN_LPF] : nullptr;
        }
        break;
    }
}

// rt-neural-generic.cpp:337-365
void activate(LV2_Handle instance)
{
    Plugin* self = (Plugin*)instance;
    aidax_stream_activate(self->engine, self->stream);
}

void deactivate(LV2_Handle) {}

// rt-neural-generic.cpp:484-660
void run(LV2_Handle instance, uint32_t n_samples)
{
    Plugin* self = (Plugin*)instance;
    const Uris* uris = &self->uris; // (only read when the control port carries events)

    // control ports -> engine (only when something changed; the decode itself mirrors :489-517)
    float c[AIDAX_N_CONTROLS];
    if (self->ctl_row) {
        memcpy(c, self->ctl_row, sizeof(c));
        c[AIDAX_CTRL_MODELINSIZE] = 0.f;
        // :518; stores that would not change memory are skipped throughout the ordinary-period path: a written line
        // costs a write-back, and with ~400k instances per period run() is bound by host memory traffic
        if (self->ctl_row[AIDAX_CTRL_MODELINSIZE] != (float)self->last_input_size) self->ctl_row[AIDAX_CTRL_MODELINSIZE] = (float)self->last_input_size;
    } else {
        for (int i = 0; i < AIDAX_N_CONTROLS; ++i) {
            const float* p = self->ctl[AIDAX_CTRL_PORT_BASE + i];
            c[i] = (p && i != AIDAX_CTRL_MODELINSIZE) ? *p : 0.f;
        }
        if (self->ctl[P_INPUT_SIZE]) *self->ctl[P_INPUT_SIZE] = (float)self->last_input_size; // :518
    }

    // atom messages (:526-586)
    if (self->notify_port && self->control_port) {
        const uint32_t capacity = self->notify_port->atom.size;
        if (self->notify_capacity != capacity) self->notify_capacity = capacity;
        if (!self->notify_armed) self->notify_armed = true;
        if (self->control_port->atom.size <= sizeof(LV2_Atom_Sequence_Body)) {
            // no events this period: leave what lv2_atom_forge_sequence_head leaves in the notify port (an empty
            // sequence, if it fits) without touching the forge; work_response opens it on demand
            if (self->notify_capacity >= sizeof(LV2_Atom_Sequence)) {
                LV2_Atom_Sequence* ns = self->notify_port;
                ns->atom.size = (uint32_t)sizeof(LV2_Atom_Sequence_Body);
                ns->atom.type = self->urid_sequence;
                ns->body.unit = 0;
                ns->body.pad = 0;
            }
            if (self->forge_ready) self->forge_ready = false;
        } else {
            open_notify_sequence(self);
            LV2_ATOM_SEQUENCE_FOREACH(self->control_port, ev)
            {
                if (!lv2_atom_forge_is_object_type(&self->forge, ev->body.type)) {
                    lv2_log_trace(&self->logger, "Unknown event type %d\n", ev->body.type);
                    continue;
                }
                const LV2_Atom_Object* obj = (const LV2_Atom_Object*)&ev->body;
                if (obj->body.otype != uris->patch_Set) {
                    lv2_log_trace(&self->logger, "Unknown object type %d\n", obj->body.otype);
                    continue;
                }
                const LV2_Atom* property = nullptr;
                const LV2_Atom* value = nullptr;
                lv2_atom_object_get(obj, uris->patch_property, &property, uris->patch_value, &value, 0);
                if (!property) { lv2_log_trace(&self->logger, "patch:Set message with no property\n"); continue; }
                if (property->type != uris->atom_URID) { lv2_log_trace(&self->logger, "patch:Set property is not a URID\n"); continue; }
                if (((const LV2_Atom_URID*)property)->body != uris->json) { lv2_log_trace(&self->logger, "patch:Set property body is not json\n"); continue; }
                if (!value) { lv2_log_trace(&self->logger, "patch:Set message with no value\n"); continue; }
                if (value->type != uris->atom_Path) { lv2_log_trace(&self->logger, "patch:Set value is not a Path\n"); continue; }
                lv2_log_trace(&self->logger, "Queueing set message\n");
                WorkerLoadMessage msg;
                memset(&msg, 0, sizeof(msg));
                msg.type = kWorkerLoad;
                memcpy(msg.path, value + 1, std::min(value->size, (uint32_t)(sizeof(msg.path) - 1u)));
                self->schedule->schedule_work(self->schedule->handle, sizeof(msg), &msg);
                self->loading = true;
            }
        }
    }

    if (!self->latched_valid || self->latched_loading != self->loading || memcmp(c, self->latched, sizeof(c)) != 0) {
        aidax_stream_set_controls(self->engine, self->stream, c, self->loading ? 1 : 0);
        memcpy(self->latched, c, sizeof(c));
        self->latched_valid = true;
        self->latched_loading = self->loading;
    }

    if (n_samples == 0) return; // pre-run (:606-609)
    if (!self->in || !self->out_1) return;

    bool deferred = false;
    if (g_hub.batching.load(std::memory_order_acquire) && self->stream >= 0 && (size_t)self->stream < g_hub.slots.size()) {
        g_hub.slots[self->stream] = EngineHub::Slot{self->in, self->out_1, n_samples, g_hub.epoch};
        deferred = true;
        EngineHub::Block& bl = g_hub.blocks[self->stream / EngineHub::kBlockStreams];
        if (bl.parked.fetch_add(1, std::memory_order_acq_rel) + 1 == bl.live.load(std::memory_order_relaxed)) {
            const int c = self->stream / g_hub.chunk_streams;
            EngineHub::Chunk& ch = g_hub.chunks[c];
            if (ch.blocks_done.fetch_add(1, std::memory_order_acq_rel) + 1 == ch.blocks_live.load(std::memory_order_relaxed)) submit_chunk(c);
        }
    }
    if (!deferred) {
        // lone instance: render this stream now and wait, so run() has LV2 semantics
        const int rc = aidax_render_host(self->engine, self->stream, 1, (int)n_samples, self->in, self->out_1, AIDAX_LAYOUT_PLANAR, n_samples);
        if (rc != AIDAX_OK) {
            lv2_log_error(&self->logger, "aidax render failed: %s\n", aidax_last_error());
            if (self->out_1 != self->in) memcpy(self->out_1, self->in, sizeof(float) * n_samples);
        }
    }
}

// rt-neural-generic.cpp:664-677
void cleanup(LV2_Handle instance)
{
    Plugin* self = (Plugin*)instance;
    aidax_stream_close(self->engine, self->stream);
    free_loaded(self->model);
    {
        std::lock_guard<std::mutex> lk(g_hub.mu);
        g_hub.refs--;
        if (g_hub.blocks[self->stream / EngineHub::kBlockStreams].live.fetch_sub(1) == 1)
            g_hub.chunks[self->stream / g_hub.chunk_streams].blocks_live.fetch_sub(1);
    }
    delete self;
}

// rt-neural-generic.cpp:707-759
LV2_State_Status restore(LV2_Handle instance, LV2_State_Retrieve_Function retrieve, LV2_State_Handle handle, uint32_t,
                         const LV2_Feature* const* features)
{
    Plugin* self = (Plugin*)instance;
    size_t size = 0;
    uint32_t type = 0, valflags = 0;
    const void* value = retrieve(handle, self->uris.json, &size, &type, &valflags);
    if (value) {
        lv2_log_note(&self->logger, "Restoring file %s\n", (const char*)value);
        WorkerLoadMessage msg;
        memset(&msg, 0, sizeof(msg));
        msg.type = kWorkerLoad;
        LV2_State_Map_Path* map_path = nullptr;
        LV2_State_Free_Path* free_path = nullptr;
        for (int i = 0; features && features[i]; ++i) {
            if (!strcmp(features[i]->URI, LV2_STATE__mapPath)) map_path = (LV2_State_Map_Path*)features[i]->data;
            else if (!strcmp(features[i]->URI, LV2_STATE__freePath)) free_path = (LV2_State_Free_Path*)features[i]->data;
        }
        if (map_path) {
            char* apath = map_path->absolute_path(map_path->handle, (const char*)value);
            memcpy(msg.path, apath, std::min(strlen(apath), sizeof(msg.path) - 1u));
            if (free_path) free_path->free_path(free_path->handle, apath);
            else free(apath);
        } else {
            memcpy(msg.path, value, std::min(size, sizeof(msg.path) - 1u));
        }
        self->schedule->schedule_work(self->schedule->handle, sizeof(msg), &msg);
    }
    return LV2_STATE_SUCCESS;
}

// rt-neural-generic.cpp:763-796
LV2_State_Status save(LV2_Handle instance, LV2_State_Store_Function store, LV2_State_Handle handle, uint32_t,
                      const LV2_Feature* const* features)
{
    Plugin* self = (Plugin*)instance;
    if (!self->model) return LV2_STATE_SUCCESS; // nothing loaded yet
    LV2_State_Map_Path* map_path = nullptr;
    for (int i = 0; features && features[i]; ++i)
        if (!strcmp(features[i]->URI, LV2_STATE__mapPath)) map_path = (LV2_State_Map_Path*)features[i]->data;
    if (!map_path) return LV2_STATE_ERR_NO_FEATURE;
    char* apath = map_path->abstract_path(map_path->handle, self->model->path);
    store(handle, self->uris.json, apath, strlen(apath) + 1, self->uris.atom_Path, LV2_STATE_IS_POD | LV2_STATE_IS_PORTABLE);
    free(apath);
    return LV2_STATE_SUCCESS;
}

// rt-neural-generic.cpp:807-848 (+ loadModelFromPath :963-1085: parse, variant selection, weights -> engine)
LV2_Worker_Status work(LV2_Handle instance, LV2_Worker_Respond_Function respond, LV2_Worker_Respond_Handle handle, uint32_t,
                       const void* data)
{
    Plugin* self = (Plugin*)instance;
    const WorkerMessage* msg = (const WorkerMessage*)data;
    switch (msg->type) {
    case kWorkerLoad: {
        const char* path = ((const WorkerLoadMessage*)data)->path;
        float p1 = 0.f, p2 = 0.f;
        aidax_stream_get_params(self->engine, self->stream, &p1, &p2); // :822-825
        int model_id = -1, input_size = 0;
        if (aidax_model_load_json(self->engine, path, &model_id, &input_size) == AIDAX_OK) {
            lv2_log_note(&self->logger, "Successfully loaded json file: %s\n", path);
            LoadedModel* m = new LoadedModel{model_id, input_size, strdup(path), p1, p2};
            // "cache input size for later" (:1082): published before the swap, as in the reference
            self->last_input_size = input_size;
            WorkerApplyMessage reply = {kWorkerApply, m};
            respond(handle, sizeof(reply), &reply);
        } else {
            // failed load: old model keeps running, `loading` stays set and the output stays muted (:1017-1020, 1041-1044)
            lv2_log_error(&self->logger, "%s\n", aidax_last_error());
        }
        return LV2_WORKER_SUCCESS;
    }
    case kWorkerFree:
        free_loaded(((const WorkerApplyMessage*)data)->model);
        return LV2_WORKER_SUCCESS;
    case kWorkerApply:
        break; // should not happen
    }
    return LV2_WORKER_ERR_UNKNOWN;
}

// rt-neural-generic.cpp:859-893
LV2_Worker_Status work_response(LV2_Handle instance, uint32_t, const void* data)
{
    Plugin* self = (Plugin*)instance;
    const WorkerMessage* msg = (const WorkerMessage*)data;
    if (msg->type != kWorkerApply) return LV2_WORKER_ERR_UNKNOWN;

    WorkerApplyMessage reply = {kWorkerFree, self->model};
    self->model = ((const WorkerApplyMessage*)data)->model;
    // swap: bind the new model; the engine resets and pre-warms the recurrent state on the GPU
    aidax_stream_set_model(self->engine, self->stream, self->model->model_id, self->model->old_param1, self->model->old_param2);
    self->schedule->schedule_work(self->schedule->handle, sizeof(reply), &reply);
    lv2_log_trace(&self->logger, "New model in use\n");

    // report the change to host/ui (write_set_file, uris.h:76-94)
    if (self->notify_armed) {
        if (!self->forge_ready) open_notify_sequence(self);
        lv2_atom_forge_frame_time(&self->forge, 0);
        LV2_Atom_Forge_Frame frame;
        lv2_atom_forge_object(&self->forge, &frame, 0, self->uris.patch_Set);
        lv2_atom_forge_key(&self->forge, self->uris.patch_property);
        lv2_atom_forge_urid(&self->forge, self->uris.json);
        lv2_atom_forge_key(&self->forge, self->uris.patch_value);
        lv2_atom_forge_path(&self->forge, self->model->path, (uint32_t)strlen(self->model->path) + 1);
        lv2_atom_forge_pop(&self->forge, &frame);
    }
    self->loading = false;
    lv2_log_trace(&self->logger, "loading = false\n");
    return LV2_WORKER_SUCCESS;
}

// rt-neural-generic.cpp:681-698
const void* extension_data(const char* uri)
{
    static const LV2_State_Interface state = {save, restore};
    static const LV2_Worker_Interface worker = {work, work_response, nullptr};
    if (!strcmp(uri, LV2_STATE__interface)) return &state;
    if (!strcmp(uri, LV2_WORKER__interface)) return &worker;
    if (!strcmp(uri, AIDAX_BATCH__interface)) return &g_batch_iface;
    return nullptr;
}

const LV2_Descriptor g_descriptor = {AIDAX_PLUGIN_URI, instantiate, connect_port, activate, run, deactivate, cleanup, extension_data};

} // namespace

// rt-neural-generic.cpp:24-29
LV2_SYMBOL_EXPORT const LV2_Descriptor* lv2_descriptor(uint32_t index) { return index == 0 ? &g_descriptor : nullptr; }
