/*
 * aidax_batch.h -- the one addition to the reference's LV2 surface: a batch extension.
 *
 * LV2's run() is synchronous per instance (rt-neural-generic.cpp:484), but the GPU wants every
 * stream of an audio period in ONE launch.  A batch-render host obtains this interface through the
 * plugin's own extension_data() (rt-neural-generic.cpp:681-698 is where the reference serves the
 * worker/state interfaces) and brackets the N run() calls of a period with begin()/end():
 *
 *     const AidaxBatchInterface* b = desc->extension_data(AIDAX_BATCH__interface);
 *     b->configure(n_instances, max_block, device);            // once, before instantiate
 *     float* arena = b->arena_alloc(bytes);                    // pinned host memory for the audio ports
 *     ... instantiate N instances, connect_port() rows of the arena ...
 *     per period:  b->begin();  for each instance: desc->run(h, n);  b->end();
 *
 * Between begin() and end(), run() does everything the reference's run() does at control rate
 * (control decode, atom messages, worker scheduling) but defers the audio; end() renders all
 * deferred instances with one engine call and returns when the output buffers are filled.
 * A host that knows nothing about this extension (jalv, mod-host) never calls begin(), and run()
 * then renders its single stream synchronously -- plain LV2 semantics.
 */
#ifndef AIDAX_BATCH_H
#define AIDAX_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the C ABI is the only thing a library built with -fvisibility=hidden exports */
#endif

#define AIDAX_PLUGIN_URI "http://aidadsp.cc/plugins/aidadsp-bundle/rt-neural-generic"
#define AIDAX_BATCH__interface AIDAX_PLUGIN_URI "#batch"

typedef struct AidaxBatchInterface {
    uint32_t abi_version; /* AIDAX_BATCH_ABI_VERSION */
    /* Size the process-wide engine before the first instantiate(); device < 0 = $AIDAX_DEVICE or 0.
     * Returns 0 on success, nonzero if an engine with a different size already exists. */
    int (*configure)(int max_instances, int max_block, int device);
    /* Pinned host memory for audio port buffers (cudaHostAlloc); NULL on failure. */
    void* (*arena_alloc)(size_t bytes);
    void (*arena_free)(void* p);
    /* Start deferring run() calls of the calling process. */
    int (*begin)(void);
    /* Render every deferred instance, block until the outputs are in the port buffers.
     * Returns 0, or an AIDAX_ERR_* code (the outputs are then undefined). */
    int (*end)(void);
    /* Kernel launches issued by the engine so far (bench accounting). */
    uint64_t (*launch_count)(void);
    /* Text of the last engine error on this thread. */
    const char* (*last_error)(void);

    /* ---- since AIDAX_BATCH_ABI_VERSION 2 (a host checks abi_version >= 2 before using them) ---- */

    /* Bulk form of run(): has exactly the effect of
     *     for (i = 0; i < count; ++i) descriptor->run(handles[i], n_samples);
     * (rt-neural-generic.cpp:484-660) but costs a few nanoseconds per instance on an ordinary period.  An instance
     * whose control ports hold the values latched by its last run(), whose CONTROL atom port carries no event and
     * whose model is not being swapped needs no control-rate work at all (the reference would decode the same dB
     * values into the same gains, find every EQ parameter unchanged and write the same ModelInSize); for those the
     * plugin only notes that the instance takes part in the period.  Every other instance goes through run() itself.
     * Differences a host must know about: on the short path the NOTIFY port is not touched (it still holds what the
     * instance's last run() left there -- an empty sequence unless a model-change notification was written), so the
     * host must not reset it to a "capacity" chunk beforehand; call run() for an instance whose notification it has
     * just consumed.  May be called from several threads at once on disjoint sets of instances, between begin() and
     * end().  Outside a batch it simply loops over run().  Returns 0.
     * flags: AIDAX_RUN_PORTS_UNCHANGED = the host asserts that since the last run() / run_block() of each of these
     * instances it has written none of their control ports and queued no event on their CONTROL ports (a host that
     * owns the control buffers knows when it writes them).  The plugin then skips the comparison of the 21 control
     * values per instance as well; an instance that still needs control-rate work for reasons of its own (first period,
     * a model swap in flight) goes through run() regardless. */
    int (*run_block)(void* const* handles, uint32_t count, uint32_t n_samples, uint32_t flags);

    /* PCI bus id ("0000:1b:00.0") of the engine's GPU, so that a host can pin its threads and arenas to the GPU's NUMA
     * node (/sys/bus/pci/devices/<id>/local_cpulist).  Returns 0 and fills `out`, or nonzero if there is no engine. */
    int (*device_pci_bus_id)(char* out, int cap);
} AidaxBatchInterface;

#define AIDAX_BATCH_ABI_VERSION 2
#define AIDAX_RUN_PORTS_UNCHANGED 1u

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
