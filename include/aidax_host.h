/*
 * aidax_host.h -- C ABI of the batch-render LV2 host (libaidax_host.so / aidax-batch-host).
 *
 * The reference ships no host: an LV2 host (mod-host, jalv, a DAW) dlopens rt-neural-generic.so and
 * calls lv2_descriptor(0)->instantiate / connect_port / run (rt-neural-generic.cpp:11-29).  This
 * host does exactly that for N instances of ANY binary exporting that descriptor -- the CUDA-backed
 * plugin of this repository or the unmodified reference build -- providing urid:map, a synchronous
 * work:schedule (scheduled work is executed and answered right after the period's run() calls),
 * log:log and state:mapPath, a planar [N][max_block] audio arena the ports are connected to, and
 * patch:Set model-load messages on the CONTROL atom port (rt-neural-generic.cpp:537-576).
 * If the plugin offers the #batch extension (include/aidax_batch.h) the N run() calls of a period
 * are bracketed by begin()/end() and rendered by one GPU launch sequence; otherwise instances are
 * run one by one on `n_threads` host threads (how the CPU baseline is timed).
 */
#ifndef AIDAX_HOST_H
#define AIDAX_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the C ABI is the only thing a library built with -fvisibility=hidden exports */
#endif

typedef struct aidax_host aidax_host;

const char* aidax_host_error(void);

/* device: CUDA device for a batched plugin (< 0 = plugin default); ignored by CPU plugins. */
aidax_host* aidax_host_open(const char* plugin_so, double sample_rate, int n_instances, int max_block, int n_threads, int device);
void aidax_host_close(aidax_host* h);

int aidax_host_n_instances(const aidax_host* h);
int aidax_host_is_batched(const aidax_host* h);
uint64_t aidax_host_launch_count(const aidax_host* h);
/* Wall-clock seconds spent since open() in the two phases of aidax_host_run: out[0] = the N run() calls (control decode,
 * atom handling, port parking), out[1] = the batch end() (copies + kernels + synchronise), out[2] = periods rendered. */
void aidax_host_timing(const aidax_host* h, double out[3]);

/* Audio arena: row i (stride aidax_host_stride() floats) is connected to instance i's IN / OUT port. */
float* aidax_host_input(aidax_host* h);
float* aidax_host_output(aidax_host* h);
size_t aidax_host_stride(const aidax_host* h);

/* Control ports by LV2 port index (4..24, ttl/rt-neural-generic.ttl).  instance < 0 = all instances. */
int aidax_host_set_control(aidax_host* h, int instance, int port, float value);
float aidax_host_get_control(const aidax_host* h, int instance, int port);

/* Queue a patch:Set{property=<plugin>#json, value=Path} for the next run() of the instance(s). */
int aidax_host_load_model(aidax_host* h, int instance, const char* json_path);

int aidax_host_activate(aidax_host* h);
int aidax_host_deactivate(aidax_host* h);

/* One audio period: run(n_samples) on every active instance, then the worker round trip.  With a plugin that offers
 * AidaxBatchInterface::run_block (include/aidax_batch.h) instances with nothing pending are handed over in runs. */
int aidax_host_run(aidax_host* h, int n_samples);

/* Only instances [0, n_active) take part in the following periods (all instances after open()).  An LV2 host is free
 * not to run an instance; the bench uses this to find the largest instance count whose period meets the deadline. */
int aidax_host_set_active(aidax_host* h, int n_active);

/* 1 if periods go through the plugin's bulk entry (run_block), 0 if every instance is run() one by one
 * (plugins without the extension, AIDAX_HOST_NO_BULK=1). */
int aidax_host_is_bulk(const aidax_host* h);

/* The cpus the period threads are pinned to (one per thread; thread 0 is the caller of aidax_host_run): cpus local to
 * the plugin's GPU when the plugin reports its PCI device, rotated by $LOCAL_RANK, else the allowed cpus in order.
 * AIDAX_HOST_PIN=0 disables pinning.  Returns the count and fills up to `cap` entries. */
int aidax_host_cpus(const aidax_host* h, int* out, int cap);

/* Number of patch:Set notifications an instance has emitted on its NOTIFY port, and the last path. */
int aidax_host_notify_count(const aidax_host* h, int instance);
const char* aidax_host_last_notify_path(const aidax_host* h, int instance);

/* state:interface save -> abstract path stored under #json (empty string if nothing loaded). Returns LV2_State_Status. */
int aidax_host_save_state(aidax_host* h, int instance, char* path_out, int path_cap);
/* state:interface restore with the given path. */
int aidax_host_restore_state(aidax_host* h, int instance, const char* path);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
