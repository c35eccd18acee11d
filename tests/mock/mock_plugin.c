/* mock_plugin.c -- a CPU-only stand-in for rt-neural-generic.so that offers the #batch extension (ABI 2), so that the
 * batch host's period logic (include/aidax_host.h) can be tested without a GPU: which instances go through run(), which
 * through run_block(), with which flags, how the worker round trip and set_active interact.  Audio: out = 2 * in.
 * Built on the fly by tests/test_host_bulk.py.  Test infrastructure only. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aidax_batch.h"
#include "../../include/aidax_lv2_abi.h"

typedef struct {
    const float* in; float* out; const LV2_Atom_Sequence* control; LV2_Atom_Sequence* notify;
    LV2_Worker_Schedule* schedule; int index; int runs; int block_runs; int events; int worked;
} Mock;

static int g_next = 0, g_begin = 0, g_end = 0, g_in_batch = 0;
static void* g_all[65536];
static uint32_t g_last_flags = 0;
static int g_block_calls = 0;

static LV2_Handle instantiate(const LV2_Descriptor* d, double sr, const char* p, const LV2_Feature* const* f)
{
    (void)d; (void)sr; (void)p;
    Mock* m = (Mock*)calloc(1, sizeof(Mock));
    for (int i = 0; f && f[i]; ++i)
        if (!strcmp(f[i]->URI, LV2_WORKER__schedule)) m->schedule = (LV2_Worker_Schedule*)f[i]->data;
    m->index = g_next++;
    if (m->index < 65536) g_all[m->index] = m;
    return m;
}
static void connect_port(LV2_Handle h, uint32_t port, void* data)
{
    Mock* m = (Mock*)h;
    if (port == 0) m->in = (const float*)data;
    else if (port == 1) m->out = (float*)data;
    else if (port == 2) m->control = (const LV2_Atom_Sequence*)data;
    else if (port == 3) m->notify = (LV2_Atom_Sequence*)data;
}
static void render(Mock* m, uint32_t n) { for (uint32_t i = 0; i < n; ++i) m->out[i] = 2.f * m->in[i]; }
static void run(LV2_Handle h, uint32_t n)
{
    Mock* m = (Mock*)h;
    m->runs++;
    if (m->control && m->control->atom.size > sizeof(LV2_Atom_Sequence_Body)) {
        m->events++;
        int msg[2] = {0, m->index};
        m->schedule->schedule_work(m->schedule->handle, sizeof(msg), msg);
    }
    if (m->notify) { /* what the real plugin leaves: an empty sequence */
        m->notify->atom.size = (uint32_t)sizeof(LV2_Atom_Sequence_Body);
    }
    render(m, n);
}
static void activate(LV2_Handle h) { (void)h; }
static void cleanup(LV2_Handle h) { free(h); }
static LV2_Worker_Status work(LV2_Handle h, LV2_Worker_Respond_Function respond, LV2_Worker_Respond_Handle rh, uint32_t size, const void* data)
{
    (void)size;
    int reply[2] = {1, ((const int*)data)[1]};
    ((Mock*)h)->worked++;
    respond(rh, sizeof(reply), reply);
    return LV2_WORKER_SUCCESS;
}
static LV2_Worker_Status work_response(LV2_Handle h, uint32_t size, const void* data)
{
    (void)size; (void)data;
    ((Mock*)h)->worked += 100;
    return LV2_WORKER_SUCCESS;
}
static int b_configure(int a, int b, int c) { (void)a; (void)b; (void)c; return 0; }
static void* b_alloc(size_t n) { return aligned_alloc(64, (n + 63) & ~(size_t)63); }
static void b_free(void* p) { free(p); }
static int b_begin(void) { g_begin++; g_in_batch = 1; return 0; }
static int b_end(void) { g_end++; g_in_batch = 0; return 0; }
static uint64_t b_launches(void) { return 0; }
static const char* b_err(void) { return ""; }
static int b_run_block(void* const* handles, uint32_t count, uint32_t n, uint32_t flags)
{
    __sync_fetch_and_add(&g_block_calls, 1);
    g_last_flags = flags;
    for (uint32_t i = 0; i < count; ++i) {
        Mock* m = (Mock*)handles[i];
        m->block_runs++;
        render(m, n);
    }
    return 0;
}
static int b_bus(char* out, int cap) { (void)out; (void)cap; return 1; }
static const AidaxBatchInterface g_batch = {AIDAX_BATCH_ABI_VERSION, b_configure, b_alloc, b_free, b_begin, b_end, b_launches, b_err, b_run_block, b_bus};
static const void* extension_data(const char* uri)
{
    static const LV2_Worker_Interface worker = {work, work_response, NULL};
    if (!strcmp(uri, LV2_WORKER__interface)) return &worker;
    if (!strcmp(uri, AIDAX_BATCH__interface)) return &g_batch;
    return NULL;
}
static const LV2_Descriptor g_desc = {AIDAX_PLUGIN_URI, instantiate, connect_port, activate, run, NULL, cleanup, extension_data};
LV2_SYMBOL_EXPORT const LV2_Descriptor* lv2_descriptor(uint32_t i) { return i == 0 ? &g_desc : NULL; }

/* inspection hooks for the test */
LV2_SYMBOL_EXPORT int mock_counter(int which) { return which == 0 ? g_begin : which == 1 ? g_end : which == 2 ? g_block_calls : (int)g_last_flags; }
LV2_SYMBOL_EXPORT int mock_instance(int index, int which)
{
    Mock* m = (Mock*)g_all[index];
    return which == 0 ? m->runs : which == 1 ? m->block_runs : which == 2 ? m->events : m->worked;
}
