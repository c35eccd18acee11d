// batch_host.cpp -- libaidax_host.so: a minimal, multi-instance LV2 host for rt-neural-generic
// binaries (see include/aidax_host.h).  No CUDA here: the GPU is reached only through the plugin.
#include <dlfcn.h>
#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/aidax_batch.h"
#include "../../include/aidax_host.h"
#include "../../include/aidax_lv2_abi.h"

namespace {

thread_local std::string g_err;
void set_err(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

constexpr int kPortCount = 25;  // ttl/rt-neural-generic.ttl: ports 0..24
constexpr int kFirstControl = 4;
constexpr int kCtlPitch = 32;   // floats per instance in the control arena: ports 4..24 at the start of a 128-byte aligned row
constexpr int kPortIn = 0, kPortOut = 1, kPortControl = 2, kPortNotify = 3;

// ------------------------------------------------------------------ urid:map
struct UridTable {
    std::mutex mu;
    std::map<std::string, LV2_URID> ids;
    std::vector<std::string> uris;
    LV2_URID map(const char* uri)
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = ids.find(uri);
        if (it != ids.end()) return it->second;
        uris.push_back(uri);
        const LV2_URID id = (LV2_URID)uris.size();
        ids[uri] = id;
        return id;
    }
};
LV2_URID urid_map_cb(LV2_URID_Map_Handle h, const char* uri) { return ((UridTable*)h)->map(uri); }

// ------------------------------------------------------------------ log:log
int log_vprintf_cb(LV2_Log_Handle h, LV2_URID, const char* fmt, va_list ap)
{
    if (!*(bool*)h) return 0;
    return vfprintf(stderr, fmt, ap);
}
int log_printf_cb(LV2_Log_Handle h, LV2_URID type, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    const int r = log_vprintf_cb(h, type, fmt, ap);
    va_end(ap);
    return r;
}

// ------------------------------------------------------------------ state:mapPath (identity)
char* map_abstract(LV2_State_Map_Path_Handle, const char* p) { return strdup(p); }
char* map_absolute(LV2_State_Map_Path_Handle, const char* p) { return strdup(p); }

// ------------------------------------------------------------------ a tiny thread pool
// pins the calling thread to one cpu (best effort)
void pin_self(int cpu)
{
    if (cpu < 0) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(cpu, &set);
    sched_setaffinity(0, sizeof(set), &set);
}

class Pool {
public:
    // cpus: one entry per worker (worker 0 is the thread that calls run()), empty = leave the scheduler alone
    Pool(int n, const std::vector<int>& cpus) : n_(std::max(1, n)), cpus_(cpus)
    {
        for (int i = 1; i < n_; ++i) threads_.emplace_back([this, i] {
            if (!cpus_.empty()) pin_self(cpus_[(size_t)i % cpus_.size()]);
            loop(i);
        });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    // runs fn(worker_index) on every worker (the caller is worker 0) and waits
    void run(const std::function<void(int)>& fn)
    {
        if (n_ == 1) { fn(0); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            pending_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void loop(int idx)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(idx);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    int n_;
    std::vector<int> cpus_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct Instance;
struct WorkItem { std::vector<uint8_t> data; };

struct Instance {
    LV2_Handle handle = nullptr;
    aidax_host* host = nullptr;
    LV2_Worker_Schedule schedule{};
    LV2_Feature f_map{}, f_sched{}, f_log{};
    const LV2_Feature* features[4] = {};
    std::deque<WorkItem> scheduled;  // written by run()/work_response (audio context)
    std::deque<WorkItem> responses;  // written by work() through respond()
    std::vector<uint8_t> control_buf; // only while a patch:Set is pending
    bool control_pending = false;
    uint8_t* notify_buf = nullptr;
    int notify_count = 0;
    std::string last_notify_path;
};

} // namespace

struct aidax_host {
    void* dl = nullptr;
    const LV2_Descriptor* desc = nullptr;
    const LV2_Worker_Interface* worker = nullptr;
    const LV2_State_Interface* state = nullptr;
    const AidaxBatchInterface* batch = nullptr;
    double sample_rate = 48000.0;
    int n = 0, max_block = 0;
    size_t stride = 0;
    float* in = nullptr;
    float* out = nullptr;
    bool arena_from_plugin = false;
    bool inplace = false;
    float* controls = nullptr;   // [n][kCtlPitch], 64-byte aligned; port p of instance i at i*kCtlPitch + p - kFirstControl
    float* ctl(int i, int port) { return controls + (size_t)i * kCtlPitch + (port - kFirstControl); }
    const float* ctl(int i, int port) const { return controls + (size_t)i * kCtlPitch + (port - kFirstControl); }
    // what the per-period loop needs of an instance, packed (the full records below are touched on events only)
    struct Hot { LV2_Handle handle; uint8_t* notify_buf; bool control_pending; };
    std::vector<Hot> hot;
    // bulk path (AidaxBatchInterface::run_block): the handles as one array and, per instance, whether this period it
    // must go through run() itself -- a patch:Set is pending on its CONTROL port, or its NOTIFY port holds something
    // other than the empty sequence its last run() left (first period; a notification the host has just consumed)
    bool bulk = false;
    bool ports_hint = true;      // this host writes control ports only through aidax_host_set_control (which marks the instance), so it can vouch for the others
    std::vector<void*> handles;
    std::vector<uint8_t> slow;
    int n_active = 0;            // instances [0, n_active) take part in a period (aidax_host_set_active)
    std::vector<int> cpus;       // cpus the pool is pinned to (empty = not pinned)
    std::vector<Instance> inst;
    std::vector<uint8_t> notify_storage;
    uint32_t notify_cap = 0;
    alignas(8) uint8_t empty_seq[sizeof(LV2_Atom_Sequence)];
    UridTable urids;
    LV2_URID_Map map{};
    LV2_Log_Log log{};
    bool verbose = false;
    LV2_URID u_sequence = 0, u_object = 0, u_urid = 0, u_path = 0, u_patch_set = 0, u_patch_property = 0, u_patch_value = 0, u_json = 0;
    std::unique_ptr<Pool> pool;
    std::atomic<bool> any_scheduled{false};
    double t_run = 0.0, t_end = 0.0, n_periods = 0.0;
};

namespace {

LV2_Worker_Status schedule_cb(LV2_Worker_Schedule_Handle h, uint32_t size, const void* data)
{
    Instance* in = (Instance*)h;
    WorkItem w;
    w.data.assign((const uint8_t*)data, (const uint8_t*)data + size);
    in->scheduled.push_back(std::move(w));
    in->host->any_scheduled.store(true, std::memory_order_relaxed);
    return LV2_WORKER_SUCCESS;
}
LV2_Worker_Status respond_cb(LV2_Worker_Respond_Handle h, uint32_t size, const void* data)
{
    Instance* in = (Instance*)h;
    WorkItem w;
    w.data.assign((const uint8_t*)data, (const uint8_t*)data + size);
    in->responses.push_back(std::move(w));
    return LV2_WORKER_SUCCESS;
}

void init_sequence(uint8_t* buf, uint32_t body_size, LV2_URID seq_type)
{
    LV2_Atom_Sequence* s = (LV2_Atom_Sequence*)buf;
    s->atom.size = body_size;
    s->atom.type = seq_type;
    s->body.unit = 0;
    s->body.pad = 0;
}

// forge a sequence holding one patch:Set{property=#json, value=Path}
void build_load_message(aidax_host* h, Instance& in, const char* path)
{
    const uint32_t cap = 256 + (uint32_t)strlen(path);
    in.control_buf.assign(cap, 0);
    LV2_Atom_Forge forge;
    lv2_atom_forge_init(&forge, &h->map);
    lv2_atom_forge_set_buffer(&forge, in.control_buf.data(), cap);
    LV2_Atom_Forge_Frame seq_frame, obj_frame;
    lv2_atom_forge_sequence_head(&forge, &seq_frame, 0);
    lv2_atom_forge_frame_time(&forge, 0);
    lv2_atom_forge_object(&forge, &obj_frame, 0, h->u_patch_set);
    lv2_atom_forge_key(&forge, h->u_patch_property);
    lv2_atom_forge_urid(&forge, h->u_json);
    lv2_atom_forge_key(&forge, h->u_patch_value);
    lv2_atom_forge_path(&forge, path, (uint32_t)strlen(path));
    lv2_atom_forge_pop(&forge, &obj_frame);
    lv2_atom_forge_pop(&forge, &seq_frame);
    in.control_pending = true;
    const size_t idx = (size_t)(&in - h->inst.data());
    if (idx < h->hot.size()) h->hot[idx].control_pending = true;
    if (idx < h->slow.size()) h->slow[idx] = 1;
}

// count patch:Set notifications in an instance's NOTIFY buffer
void scan_notify(aidax_host* h, Instance& in)
{
    const LV2_Atom_Sequence* seq = (const LV2_Atom_Sequence*)in.notify_buf;
    if (seq->atom.type != h->u_sequence || seq->atom.size > h->notify_cap) return; // untouched (still "capacity")
    LV2_ATOM_SEQUENCE_FOREACH(seq, ev)
    {
        if (ev->body.type != h->u_object) continue;
        const LV2_Atom_Object* obj = (const LV2_Atom_Object*)&ev->body;
        if (obj->body.otype != h->u_patch_set) continue;
        const LV2_Atom* value = nullptr;
        lv2_atom_object_get(obj, h->u_patch_value, &value, 0);
        in.notify_count++;
        if (value && value->type == h->u_path) in.last_notify_path = (const char*)(value + 1);
    }
}

// run() of one instance the way an LV2 host calls it
inline void run_one(aidax_host* h, int i, uint32_t n_samples, uint32_t notify_room)
{
    aidax_host::Hot& hot = h->hot[i];
    // an LV2 host presents an output atom port as an empty chunk whose size is the capacity
    LV2_Atom_Sequence* ns = (LV2_Atom_Sequence*)hot.notify_buf;
    ns->atom.size = notify_room;
    ns->atom.type = 0;
    if (hot.control_pending) h->desc->connect_port(hot.handle, kPortControl, h->inst[i].control_buf.data());
    h->desc->run(hot.handle, n_samples);
    if (hot.control_pending) {
        h->desc->connect_port(hot.handle, kPortControl, h->empty_seq);
        hot.control_pending = false;
        h->inst[i].control_pending = false;
        h->inst[i].control_buf.clear();
    }
}

void run_instances(aidax_host* h, int lo, int hi, uint32_t n_samples)
{
    const uint32_t notify_room = h->notify_cap - (uint32_t)sizeof(LV2_Atom);
    if (h->bulk) {
        // instances with nothing pending go to the plugin in runs (a few ns each); the others one by one
        uint8_t* slow = h->slow.data();
        int i = lo;
        while (i < hi) {
            if (slow[i]) {
                run_one(h, i, n_samples, notify_room);
                slow[i] = 0;
                ++i;
                continue;
            }
            int j = i + 1;
            while (j < hi && !slow[j]) ++j;
            h->batch->run_block(h->handles.data() + i, (uint32_t)(j - i), n_samples, h->ports_hint ? AIDAX_RUN_PORTS_UNCHANGED : 0u);
            i = j;
        }
        return;
    }
    constexpr int kAhead = 4;
    for (int i = lo; i < hi; ++i) {
        // run() of a small-state plugin is bound by cache misses on per-instance data that lives in several places
        // (plugin object, control row, notify buffer): start fetching a few instances ahead
        if (i + kAhead < h->n) {
            const aidax_host::Hot& nx = h->hot[i + kAhead];
            const char* ph = (const char*)nx.handle;
            __builtin_prefetch(ph);
            __builtin_prefetch(ph + 64);
            __builtin_prefetch(ph + 128);
            __builtin_prefetch(nx.notify_buf, 1);
            __builtin_prefetch(h->ctl(i + kAhead, kFirstControl), 1);
            __builtin_prefetch(h->ctl(i + kAhead, kFirstControl + 16));
        }
        run_one(h, i, n_samples, notify_room);
    }
}

// The worker round trip of one instance: work() for everything scheduled, then work_response()
// for every reply (which may schedule follow-up work, e.g. freeing the previous model).
void drain_worker(aidax_host* h, Instance& in)
{
    for (int round = 0; round < 4 && (!in.scheduled.empty() || !in.responses.empty()); ++round) {
        while (!in.scheduled.empty()) {
            WorkItem w = std::move(in.scheduled.front());
            in.scheduled.pop_front();
            if (h->worker && h->worker->work) h->worker->work(in.handle, respond_cb, &in, (uint32_t)w.data.size(), w.data.data());
        }
        while (!in.responses.empty()) {
            WorkItem w = std::move(in.responses.front());
            in.responses.pop_front();
            if (h->worker && h->worker->work_response) h->worker->work_response(in.handle, (uint32_t)w.data.size(), w.data.data());
        }
    }
}

// "0-3,8,10-11" -> {0,1,2,3,8,10,11}
std::vector<int> parse_cpulist(const std::string& text)
{
    std::vector<int> out;
    size_t i = 0;
    while (i < text.size()) {
        while (i < text.size() && !isdigit((unsigned char)text[i])) ++i;
        if (i >= text.size()) break;
        int a = 0;
        while (i < text.size() && isdigit((unsigned char)text[i])) a = a * 10 + (text[i++] - '0');
        int b = a;
        if (i < text.size() && text[i] == '-') {
            ++i;
            b = 0;
            while (i < text.size() && isdigit((unsigned char)text[i])) b = b * 10 + (text[i++] - '0');
        }
        for (int c = a; c <= b && out.size() < 4096; ++c) out.push_back(c);
    }
    return out;
}

// The cpus this process should run its period threads on: those local to the plugin's GPU (sysfs local_cpulist of
// its PCI device) that the process is allowed to use, rotated by the rank of the process among the local ranks
// (torchrun's LOCAL_RANK) so that processes driving neighbouring GPUs do not sit on each other's cores.
cpu_set_t g_initial_affinity;   // of the thread that opened the first host, before any pinning
bool g_have_initial_affinity = false;

std::vector<int> gpu_local_cpus(const AidaxBatchInterface* batch, int n_threads)
{
    std::vector<int> allowed;
    if (!g_have_initial_affinity) {
        CPU_ZERO(&g_initial_affinity);
        if (sched_getaffinity(0, sizeof(g_initial_affinity), &g_initial_affinity) != 0) return {};
        g_have_initial_affinity = true;
    }
    const cpu_set_t& set = g_initial_affinity;
    for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &set)) allowed.push_back(c);
    std::vector<int> local;
    char bus[64] = {0};
    if (batch && batch->abi_version >= 2 && batch->device_pci_bus_id && batch->device_pci_bus_id(bus, (int)sizeof(bus)) == 0) {
        const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
        if (FILE* f = fopen(path.c_str(), "r")) {
            char line[4096] = {0};
            if (fgets(line, sizeof(line), f)) {
                for (int c : parse_cpulist(line))
                    if (std::find(allowed.begin(), allowed.end(), c) != allowed.end()) local.push_back(c);
            }
            fclose(f);
        }
    }
    if (local.empty()) local = allowed; // no NUMA information (or a VM without it): all allowed cpus
    if (local.empty()) return {};
    const char* lr = getenv("LOCAL_RANK");
    const int rank = lr && *lr ? atoi(lr) : 0;
    std::vector<int> out;
    for (int t = 0; t < std::max(1, n_threads); ++t) out.push_back(local[(size_t)(rank * std::max(1, n_threads) + t) % local.size()]);
    return out;
}

struct StateStore { std::string value; LV2_URID key = 0; };
LV2_State_Status store_cb(LV2_State_Handle hd, uint32_t key, const void* value, size_t size, uint32_t, uint32_t)
{
    StateStore* s = (StateStore*)hd;
    s->key = key;
    s->value.assign((const char*)value, size ? size - 1 : 0);
    return LV2_STATE_SUCCESS;
}
const void* retrieve_cb(LV2_State_Handle hd, uint32_t key, size_t* size, uint32_t* type, uint32_t* flags)
{
    StateStore* s = (StateStore*)hd;
    if (key != s->key) return nullptr;
    *size = s->value.size() + 1;
    *type = 0;
    *flags = LV2_STATE_IS_POD | LV2_STATE_IS_PORTABLE;
    return s->value.c_str();
}

} // namespace

extern "C" {

const char* aidax_host_error(void) { return g_err.c_str(); }

aidax_host* aidax_host_open(const char* plugin_so, double sample_rate, int n_instances, int max_block, int n_threads, int device)
{
    if (!plugin_so || n_instances <= 0 || max_block <= 0) { set_err("bad arguments"); return nullptr; }
    std::unique_ptr<aidax_host> h(new aidax_host());
    h->dl = dlopen(plugin_so, RTLD_NOW | RTLD_LOCAL);
    if (!h->dl) { set_err("dlopen(%s): %s", plugin_so, dlerror()); return nullptr; }
    auto fn = (LV2_Descriptor_Function)dlsym(h->dl, "lv2_descriptor");
    if (!fn) { set_err("%s does not export lv2_descriptor", plugin_so); return nullptr; }
    h->desc = fn(0);
    if (!h->desc) { set_err("lv2_descriptor(0) returned NULL"); return nullptr; }
    h->worker = (const LV2_Worker_Interface*)h->desc->extension_data(LV2_WORKER__interface);
    h->state = (const LV2_State_Interface*)h->desc->extension_data(LV2_STATE__interface);
    h->batch = (const AidaxBatchInterface*)h->desc->extension_data(AIDAX_BATCH__interface);
    if (h->batch && h->batch->abi_version != AIDAX_BATCH_ABI_VERSION) h->batch = nullptr;
    // AIDAX_HOST_NO_BATCH=1: behave like a stock LV2 host (every run() renders on its own)
    const bool no_batch = getenv("AIDAX_HOST_NO_BATCH") != nullptr;
    const AidaxBatchInterface* batch_for_setup = h->batch;
    h->sample_rate = sample_rate;
    h->n = n_instances;
    h->max_block = max_block;
    h->stride = (size_t)((max_block + 31) & ~31);
    h->verbose = getenv("AIDAX_HOST_VERBOSE") != nullptr;
    h->map.handle = &h->urids;
    h->map.map = urid_map_cb;
    h->log.handle = &h->verbose;
    h->log.printf = log_printf_cb;
    h->log.vprintf = log_vprintf_cb;
    h->u_sequence = h->urids.map(LV2_ATOM__Sequence);
    h->u_object = h->urids.map(LV2_ATOM__Object);
    h->u_urid = h->urids.map(LV2_ATOM__URID);
    h->u_path = h->urids.map(LV2_ATOM__Path);
    h->u_patch_set = h->urids.map(LV2_PATCH__Set);
    h->u_patch_property = h->urids.map(LV2_PATCH__property);
    h->u_patch_value = h->urids.map(LV2_PATCH__value);
    h->u_json = h->urids.map((std::string(h->desc->URI) + "#json").c_str());
    init_sequence(h->empty_seq, sizeof(LV2_Atom_Sequence_Body), h->u_sequence);

    const size_t arena_bytes = (size_t)n_instances * h->stride * sizeof(float);
    if (batch_for_setup && batch_for_setup->configure(n_instances, max_block, device) != 0) {
        set_err("plugin engine already sized differently");
        return nullptr;
    }
    if (no_batch) h->batch = nullptr;
    h->bulk = h->batch && h->batch->abi_version >= 2 && h->batch->run_block && getenv("AIDAX_HOST_NO_BULK") == nullptr;
    h->ports_hint = getenv("AIDAX_HOST_NO_PORTS_HINT") == nullptr; // (=1: let the plugin compare every control row itself)

    h->notify_cap = n_instances <= 8192 ? 2048u : 384u;
    if (const char* v = getenv("AIDAX_HOST_NOTIFY_CAP")) h->notify_cap = (uint32_t)std::max(16, atoi(v)) & ~7u; // tests: a NOTIFY port too small for a message
    h->inst.resize(n_instances);
    auto instantiate_one = [&](int i) -> bool {
        Instance& in = h->inst[i];
        in.host = h.get();
        in.schedule.handle = &in;
        in.schedule.schedule_work = schedule_cb;
        in.f_map = LV2_Feature{LV2_URID__map, &h->map};
        in.f_sched = LV2_Feature{LV2_WORKER__schedule, &in.schedule};
        in.f_log = LV2_Feature{LV2_LOG__log, &h->log};
        in.features[0] = &in.f_map; in.features[1] = &in.f_sched; in.features[2] = &in.f_log; in.features[3] = nullptr;
        in.handle = h->desc->instantiate(h->desc, sample_rate, "", in.features);
        if (!in.handle) {
            set_err("instantiate failed for instance %d", i);
            for (int k = 0; k < i; ++k) h->desc->cleanup(h->inst[k].handle);
            return false;
        }
        return true;
    };
    // The first instance brings the plugin's engine (and with it the GPU) into existence; the period threads and the
    // audio arenas are then placed next to that GPU: one thread per cpu, cpus local to the GPU's PCIe root, pinned
    // memory first touched from there.  AIDAX_HOST_PIN=0 leaves everything to the scheduler.
    if (!instantiate_one(0)) return nullptr;
    const char* pin_env = getenv("AIDAX_HOST_PIN");
    if (!(pin_env && *pin_env && atoi(pin_env) == 0)) {
        h->cpus = gpu_local_cpus(h->batch, n_threads);
        if (!h->cpus.empty()) pin_self(h->cpus[0]);
    }

    if (h->batch) {
        h->in = (float*)h->batch->arena_alloc(arena_bytes);
        h->out = (float*)h->batch->arena_alloc(arena_bytes);
        h->arena_from_plugin = true;
        if (!h->in || !h->out) { set_err("pinned arena allocation failed: %s", h->batch->last_error()); return nullptr; }
    } else {
        h->in = (float*)aligned_alloc(64, arena_bytes);
        h->out = (float*)aligned_alloc(64, arena_bytes);
        if (!h->in || !h->out) { set_err("arena allocation failed"); return nullptr; }
    }
    memset(h->in, 0, arena_bytes);
    memset(h->out, 0, arena_bytes);
    if (getenv("AIDAX_HOST_INPLACE")) { // tests: OUT connected to the IN buffer (rt-neural-generic.cpp:613 allows in == out_1)
        if (h->arena_from_plugin) h->batch->arena_free(h->out); else free(h->out);
        h->out = h->in;
        h->inplace = true;
    }

    // ttl defaults (ttl/rt-neural-generic.ttl:92-313)
    if (posix_memalign((void**)&h->controls, 64, (size_t)n_instances * kCtlPitch * sizeof(float)) != 0) { set_err("control arena allocation failed"); return nullptr; }
    memset(h->controls, 0, (size_t)n_instances * kCtlPitch * sizeof(float));
    for (int i = 0; i < n_instances; ++i) {
        *h->ctl(i, 4) = 66.216f; *h->ctl(i, 12) = 305.f; *h->ctl(i, 14) = 750.f; *h->ctl(i, 15) = 0.707f;
        *h->ctl(i, 18) = 2000.f; *h->ctl(i, 21) = 1.f; *h->ctl(i, 24) = 1.f;
    }
    h->notify_storage.assign((size_t)n_instances * h->notify_cap, 0);
    for (int i = 0; i < n_instances; ++i) {
        if (i > 0 && !instantiate_one(i)) return nullptr;
        Instance& in = h->inst[i];
        in.notify_buf = h->notify_storage.data() + (size_t)i * h->notify_cap;
        h->desc->connect_port(in.handle, kPortIn, h->in + (size_t)i * h->stride);
        h->desc->connect_port(in.handle, kPortOut, h->out + (size_t)i * h->stride);
        h->desc->connect_port(in.handle, kPortControl, h->empty_seq);
        h->desc->connect_port(in.handle, kPortNotify, in.notify_buf);
        for (int p = kFirstControl; p < kPortCount; ++p) h->desc->connect_port(in.handle, p, h->ctl(i, p));
    }
    h->hot.resize(n_instances);
    h->handles.resize(n_instances);
    h->slow.assign(n_instances, 1); // first period: every instance reads its ports through run()
    for (int i = 0; i < n_instances; ++i) {
        h->hot[i] = aidax_host::Hot{h->inst[i].handle, h->inst[i].notify_buf, false};
        h->handles[i] = h->inst[i].handle;
    }
    h->n_active = n_instances;
    h->pool.reset(new Pool(n_threads, h->cpus));
    return h.release();
}

void aidax_host_close(aidax_host* h)
{
    if (!h) return;
    h->pool.reset();
    for (Instance& in : h->inst)
        if (in.handle) {
            drain_worker(h, in);
            h->desc->cleanup(in.handle);
        }
    if (h->arena_from_plugin) { h->batch->arena_free(h->in); if (!h->inplace) h->batch->arena_free(h->out); }
    else { free(h->in); if (!h->inplace) free(h->out); }
    free(h->controls);
    // the plugin binary stays loaded: it may own process-wide state (CUDA context, engine)
    if (!h->cpus.empty() && g_have_initial_affinity) sched_setaffinity(0, sizeof(g_initial_affinity), &g_initial_affinity);
    delete h;
}

int aidax_host_n_instances(const aidax_host* h) { return h ? h->n : 0; }
int aidax_host_is_batched(const aidax_host* h) { return h && h->batch ? 1 : 0; }
uint64_t aidax_host_launch_count(const aidax_host* h) { return h && h->batch ? h->batch->launch_count() : 0; }
void aidax_host_timing(const aidax_host* h, double out[3])
{
    if (!h || !out) return;
    out[0] = h->t_run; out[1] = h->t_end; out[2] = h->n_periods;
}
float* aidax_host_input(aidax_host* h) { return h ? h->in : nullptr; }
float* aidax_host_output(aidax_host* h) { return h ? h->out : nullptr; }
size_t aidax_host_stride(const aidax_host* h) { return h ? h->stride : 0; }

int aidax_host_set_control(aidax_host* h, int instance, int port, float value)
{
    if (!h || port < 4 || port >= kPortCount || instance >= h->n) { set_err("bad control address"); return 1; }
    // the instance reads its ports through run() in the next period (see run_instances)
    if (instance >= 0) { *h->ctl(instance, port) = value; h->slow[instance] = 1; }
    else for (int i = 0; i < h->n; ++i) { *h->ctl(i, port) = value; h->slow[i] = 1; }
    return 0;
}

float aidax_host_get_control(const aidax_host* h, int instance, int port)
{
    if (!h || port < 4 || port >= kPortCount || instance < 0 || instance >= h->n) return 0.f;
    return *h->ctl(instance, port);
}

int aidax_host_load_model(aidax_host* h, int instance, const char* json_path)
{
    if (!h || !json_path || instance >= h->n) { set_err("bad arguments"); return 1; }
    if (instance >= 0) build_load_message(h, h->inst[instance], json_path);
    else for (Instance& in : h->inst) build_load_message(h, in, json_path);
    return 0;
}

int aidax_host_activate(aidax_host* h)
{
    if (!h) return 1;
    if (h->desc->activate) for (Instance& in : h->inst) h->desc->activate(in.handle);
    return 0;
}
int aidax_host_deactivate(aidax_host* h)
{
    if (!h) return 1;
    if (h->desc->deactivate) for (Instance& in : h->inst) h->desc->deactivate(in.handle);
    return 0;
}

int aidax_host_run(aidax_host* h, int n_samples)
{
    if (!h || n_samples < 0 || n_samples > h->max_block) { set_err("bad n_samples"); return 1; }
    const int nt = h->pool->size();
    const int n = h->n_active;
    int rc = 0;
    const auto t0 = std::chrono::steady_clock::now();
    if (h->batch) h->batch->begin();
    // Instances are handed out in ascending blocks, so all threads work on the same neighbourhood of the arena and the
    // plugin's batch extension sees its chunks complete one after another (it ships a chunk to the GPU as soon as the
    // last instance of the chunk has run); it also balances threads when run() itself computes (the reference).
    // Blocks are multiples of 64 instances: the plugin counts arrivals per 64-stream block, one cache line per block.
    std::atomic<int> next{0};
    const int grain = h->bulk ? (n >= 1024 * nt ? 1024 : 256) : (n >= 64 * nt ? 64 : std::max(1, n / (4 * nt)));
    h->pool->run([&](int) {
        for (;;) {
            const int lo = next.fetch_add(grain, std::memory_order_relaxed);
            if (lo >= n) break;
            run_instances(h, lo, std::min(n, lo + grain), (uint32_t)n_samples);
        }
    });
    const auto t1 = std::chrono::steady_clock::now();
    if (h->batch) {
        rc = h->batch->end();
        if (rc) set_err("batch end failed: %s", h->batch->last_error());
    }
    const auto t2 = std::chrono::steady_clock::now();
    h->t_run += std::chrono::duration<double>(t1 - t0).count();
    h->t_end += std::chrono::duration<double>(t2 - t1).count();
    h->n_periods += 1.0;
    if (h->any_scheduled.load(std::memory_order_relaxed)) {
        h->any_scheduled.store(false, std::memory_order_relaxed);
        const int n_all = h->n;
        h->pool->run([&](int w) {
            const int lo = (int)((int64_t)n_all * w / nt), hi = (int)((int64_t)n_all * (w + 1) / nt);
            for (int i = lo; i < hi; ++i) {
                Instance& in = h->inst[i];
                if (in.scheduled.empty() && in.responses.empty()) continue;
                drain_worker(h, in);
                scan_notify(h, in);
                h->slow[i] = 1; // its NOTIFY port may hold a notification now: the next period resets it through run()
            }
        });
    }
    return rc;
}

int aidax_host_set_active(aidax_host* h, int n_active)
{
    if (!h || n_active < 0 || n_active > h->n) { set_err("bad active count"); return 1; }
    h->n_active = n_active;
    return 0;
}

int aidax_host_is_bulk(const aidax_host* h) { return h && h->bulk ? 1 : 0; }

int aidax_host_cpus(const aidax_host* h, int* out, int cap)
{
    if (!h) return 0;
    const int n = (int)h->cpus.size();
    for (int i = 0; i < n && i < cap && out; ++i) out[i] = h->cpus[i];
    return n;
}

int aidax_host_notify_count(const aidax_host* h, int instance)
{
    return (h && instance >= 0 && instance < h->n) ? h->inst[instance].notify_count : 0;
}
const char* aidax_host_last_notify_path(const aidax_host* h, int instance)
{
    return (h && instance >= 0 && instance < h->n) ? h->inst[instance].last_notify_path.c_str() : "";
}

int aidax_host_save_state(aidax_host* h, int instance, char* path_out, int path_cap)
{
    if (!h || !h->state || instance < 0 || instance >= h->n) { set_err("no state interface"); return -1; }
    LV2_State_Map_Path mp{nullptr, map_abstract, map_absolute};
    LV2_Feature f_mp{LV2_STATE__mapPath, &mp};
    const LV2_Feature* feats[] = {&f_mp, nullptr};
    StateStore st;
    const int rc = (int)h->state->save(h->inst[instance].handle, store_cb, &st, 0, feats);
    if (path_out && path_cap > 0) snprintf(path_out, (size_t)path_cap, "%s", st.value.c_str());
    return rc;
}

int aidax_host_restore_state(aidax_host* h, int instance, const char* path)
{
    if (!h || !h->state || !path || instance < 0 || instance >= h->n) { set_err("no state interface"); return -1; }
    LV2_State_Map_Path mp{nullptr, map_abstract, map_absolute};
    LV2_Feature f_mp{LV2_STATE__mapPath, &mp};
    const LV2_Feature* feats[] = {&f_mp, nullptr};
    StateStore st;
    st.key = h->u_json;
    st.value = path;
    const int rc = (int)h->state->restore(h->inst[instance].handle, retrieve_cb, &st, 0, feats);
    h->any_scheduled.store(true, std::memory_order_relaxed); // restore schedules a load; run it with the next period
    return rc;
}

} // extern "C"
