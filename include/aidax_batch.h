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
} AidaxBatchInterface;

#define AIDAX_BATCH_ABI_VERSION 1

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
