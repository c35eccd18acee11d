/*
 * aidax.h -- C ABI of the B200-native batch engine for the rt-neural-generic hot path.
 *
 * This is the drop-in boundary for the per-sample DSP chain of AidaDSP/aidadsp-lv2
 * (RtNeuralGeneric::run, rt-neural-generic/src/rt-neural-generic.cpp:484-660): plain pointers
 * and sizes, `extern "C"`, int status codes, no exceptions and no torch types.  One engine owns
 * one GPU; a "stream" is what the reference calls a plugin instance (one mono audio stream with
 * its own filters, smoothers, model binding and recurrent state).  All streams of an engine are
 * rendered by ONE aidax_render_* call per audio period.
 *
 * Each entry point cites the reference interface it replaces.  The LV2 plugin built on top of
 * this ABI (csrc/plugin.cpp -> rt-neural-generic.so) keeps the reference's lv2_descriptor /
 * instantiate / connect_port / run / worker / state surface unchanged.
 *
 * There is NO CPU fallback: every call that needs the device fails with AIDAX_ERR_CUDA when no
 * sm_100 GPU is usable.
 */
#ifndef AIDAX_H
#define AIDAX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the C ABI is the only thing a library built with -fvisibility=hidden exports */
#endif

#define AIDAX_ABI_VERSION 2

/* status codes (all entry points returning int) */
enum {
    AIDAX_OK = 0,
    AIDAX_ERR_ARG = 1,      /* bad argument / unknown id */
    AIDAX_ERR_CUDA = 2,     /* CUDA runtime failure or no usable device */
    AIDAX_ERR_MODEL = 3,    /* model file unreadable / unsupported architecture (reference: loadModelFromPath returns nullptr, rt-neural-generic.cpp:1017-1044) */
    AIDAX_ERR_CAPACITY = 4, /* max_streams / max_block exceeded */
    AIDAX_ERR_STATE = 5     /* call not valid in the current state */
};

/* cell family, reference: variant/generate_variant_hpp.py:4 */
enum { AIDAX_CELL_LSTM = 0, AIDAX_CELL_GRU = 1 };

/* audio buffer layouts accepted by the render calls */
enum {
    AIDAX_LAYOUT_PLANAR = 0,      /* [stream][n_samples]  (LV2 port layout, row stride given in samples) */
    AIDAX_LAYOUT_INTERLEAVED = 1  /* [n_samples][stream]  (stream-interleaved, row stride given in streams) */
};

/*
 * Control values of one stream, in the reference's port order (ports_t,
 * rt-neural-generic/src/rt-neural-generic.h:84-112; ttl/rt-neural-generic.ttl:92-313).
 * Index = LV2 port index - AIDAX_CTRL_PORT_BASE.  Raw port values (dB, %, Hz ...), decoded by
 * the engine exactly as run() does (rt-neural-generic.cpp:489-517 and 50-127).
 */
enum {
    AIDAX_CTRL_PORT_BASE = 4,
    AIDAX_CTRL_ANTIALIASING = 0, /* port 4  IN_LPF %   (0 = off)                         */
    AIDAX_CTRL_PREGAIN = 1,      /* port 5  dB                                           */
    AIDAX_CTRL_NETBYPASS = 2,    /* port 6  > 0.5 bypasses the network                   */
    AIDAX_CTRL_PARAM1 = 3,       /* port 7                                               */
    AIDAX_CTRL_PARAM2 = 4,       /* port 8                                               */
    AIDAX_CTRL_EQBYPASS = 5,     /* port 9                                               */
    AIDAX_CTRL_EQPOS = 6,        /* port 10 0 post, 1 pre                                */
    AIDAX_CTRL_BASS = 7,         /* port 11 dB                                           */
    AIDAX_CTRL_BFREQ = 8,        /* port 12 Hz                                           */
    AIDAX_CTRL_MID = 9,          /* port 13 dB                                           */
    AIDAX_CTRL_MFREQ = 10,       /* port 14 Hz                                           */
    AIDAX_CTRL_MIDQ = 11,        /* port 15                                              */
    AIDAX_CTRL_MTYPE = 12,       /* port 16 0 peak, 1 bandpass                           */
    AIDAX_CTRL_TREBLE = 13,      /* port 17 dB                                           */
    AIDAX_CTRL_TFREQ = 14,       /* port 18 Hz                                           */
    AIDAX_CTRL_DEPTH = 15,       /* port 19 dB                                           */
    AIDAX_CTRL_PRESENCE = 16,    /* port 20 dB                                           */
    AIDAX_CTRL_DCBLOCKER = 17,   /* port 21 == 1 enables the 35 Hz high-pass             */
    AIDAX_CTRL_MASTER = 18,      /* port 22 dB                                           */
    AIDAX_CTRL_MODELINSIZE = 19, /* port 23 OUTPUT port, ignored on input                */
    AIDAX_CTRL_ENABLED = 20,     /* port 24 > 0.5 = process, else straight copy          */
    AIDAX_N_CONTROLS = 21
};

typedef struct aidax_engine aidax_engine;

/* Description of one neural model, reference: DynamicModel (rt-neural-generic.h:115-129) +
 * the JSON fields read by loadModelFromPath (rt-neural-generic.cpp:972-1013). */
typedef struct aidax_model_desc {
    int32_t cell;        /* AIDAX_CELL_*                                    */
    int32_t hidden;      /* 8,12,16,20,24,32,40,64,80 (model_variant.hpp)   */
    int32_t input_size;  /* 1..3 (MAX_INPUT_SIZE)                           */
    int32_t in_skip;     /* 0/1                                             */
    float in_gain;       /* linear (already DB_CO'd)                        */
    float out_gain;      /* linear                                          */
    float samplerate;    /* model sample rate (param smoother rate)         */
} aidax_model_desc;

/* ------------------------------------------------------------------ engine ---- */

/* Library-wide: last error text of the calling thread ("" if none). */
const char* aidax_last_error(void);
int aidax_abi_version(void);

/* Create an engine on CUDA device `device` for up to `max_streams` streams and render calls of
 * up to `max_block` samples.  Replaces nothing in the reference (there one instance == one CPU
 * thread); it is the container the batch host creates once per GPU. */
int aidax_engine_create(int device, int max_streams, int max_block, aidax_engine** out);
int aidax_engine_destroy(aidax_engine* e);

/* cudaStream_t (as void*) all engine work is enqueued on; for external event timing. */
void* aidax_engine_cuda_stream(aidax_engine* e);

/* Tuning: which kernel variant renders a model family.  mode 0 = auto (by stream count: FP32 latency shape ->
 * FP32 throughput shape -> tensor-core shape for families of >= 4096 streams, env AIDAX_TC_MIN_STREAMS),
 * 1 = FP32 latency shape (many lanes per stream), 2 = FP32 throughput shape (register-tiled, many streams per
 * warp), 3 = tcgen05 tensor-core shape (128 streams per tile, 3xTF32 gate contraction; hidden <= 40, falls back
 * to mode 0's FP32 choice for hidden 64/80).
 * math: 0 = fast (MUFU ex2/rcp), 1 = precise (expf + IEEE division). */
int aidax_engine_configure(aidax_engine* e, int tile_mode, int math_mode);

/* ------------------------------------------------------------------ models ---- */

/* Register a model from weights in the reference JSON layout (Appendix B of SURVEY.md;
 * RTN/model_loader.h:149-196 GRU, :227-263 LSTM, :21-43 Dense):
 *   W [input_size][G*hidden], U [hidden][G*hidden], row-major, G = 4 (LSTM: i,f,c,o) or 3 (GRU: z,r,h)
 *   b LSTM: [4*hidden]; GRU: [2][3*hidden]
 *   dense_w [hidden], dense_b scalar.
 * Replaces custom_model_creator + ModelT::parseJson (model_variant.hpp:656-875, RTN/ModelT.h:318-379). */
int aidax_model_register(aidax_engine* e, const aidax_model_desc* desc, const float* W, const float* U,
                         const float* b, const float* dense_w, float dense_b, int* model_id_out);

/* Parse a reference-schema JSON file and register it.  Same acceptance rules as
 * loadModelFromPath (rt-neural-generic.cpp:963-1085): in_shape[-1] <= 3, in_skip <= 1, gains in
 * dB, samplerate from metadata.samplerate|samplerate if numeric else 48000, layers[0].type in
 * {lstm,gru}, hidden in the variant table.  The reference re-parses the file on every load (:963-1013); here a
 * file is parsed once per (path, modification time, size) while the result is still held or bound, so N instances
 * loading one capture share one weight block and a capture re-exported to the same path is read again.
 * Parsing, packing and the weight upload run WITHOUT the engine lock (the LV2 worker thread must not stall run()).
 * The returned id is HELD by the caller until aidax_model_release.
 * Returns AIDAX_ERR_MODEL on any failure (the reference returns nullptr). */
int aidax_model_load_json(aidax_engine* e, const char* path, int* model_id_out, int* input_size_out);

/* Give back an id obtained from aidax_model_load_json / aidax_model_register (the reference frees a DynamicModel in
 * freeModel on the worker thread, rt-neural-generic.cpp:1093-1101).  Streams bound to the model keep it alive; its
 * device weights are freed when the last hold and the last binding are gone.  Optional: an id never released stays
 * valid for the life of the engine. */
int aidax_model_release(aidax_engine* e, int model_id);

/* Opt-in load-time check: the reference's testModel (rt-neural-generic.cpp:900-955, called from loadModelFromPath at
 * :1066-1073 in DEBUG builds).  Renders the `input_batch` the model file carries through the network alone (gains 1,
 * PARAM1/2 = 0, state reset) on the GPU, in the engine's current kernel configuration, and compares it with the file's
 * `output_batch`: *n_errs_out = samples off by more than TEST_MODEL_THR = 1e-5 (rt-neural-generic.h:182), *max_err_out
 * the largest difference.  Needs one free stream slot.  AIDAX_ERR_STATE when the file held no vectors (or the model
 * was registered from arrays). */
int aidax_model_selftest(aidax_engine* e, int model_id, float* max_err_out, int* n_errs_out, int* n_samples_out);

int aidax_model_describe(aidax_engine* e, int model_id, aidax_model_desc* out);

/* Host-only (no GPU needed): parse + validate a model file with the rules above and report what
 * would be registered.  AIDAX_OK or AIDAX_ERR_MODEL (aidax_last_error() has the reason). */
int aidax_model_check_json(const char* path, aidax_model_desc* out);

/* Host-only: the control-rate arithmetic, exposed for verification against the reference.
 * aidax_design_biquad: Biquad::calcBiquad (common/Biquad.cpp:67-165), type 0..6 = lowpass, highpass,
 *   bandpass, notch, peak, lowshelf, highshelf; fc normalised; coef = {a0,a1,a2,b1,b2}.
 * aidax_decode_controls: what one run() derives from its control ports for a fresh instance
 *   (rt-neural-generic.cpp:489-517, 50-127): flag word (layout.h F_*), gain targets and the 7x5 filter
 *   coefficients in the order in_lpf, depth, bass, mid, treble, presence, dc. */
int aidax_design_biquad(int type, double fc, double q, double gain_db, double coef[5]);
int aidax_decode_controls(double sample_rate, const float controls[AIDAX_N_CONTROLS], int loading, int model_bound,
                          uint32_t* flags, float* pre_target, float* master_target, double coefs[35]);

/* ----------------------------------------------------------------- streams ---- */

/* Open a stream: mirrors RtNeuralGeneric::instantiate (rt-neural-generic.cpp:244-333):
 * pre-gain smoother at 1, master smoother at 0 (both T60 0.1 s), dc blocker 35 Hz HPF, in_lpf at
 * 66.216 %, EQ filters at their instantiate-time coefficients, no model, loading = true. */
int aidax_stream_open(aidax_engine* e, double sample_rate, int* stream_id_out);
int aidax_stream_close(aidax_engine* e, int stream);

/* Mirrors RtNeuralGeneric::activate (:337-365): both gain smoothers jump to their targets and
 * the model's paramFirstRun flag is set.  Recurrent state is NOT reset (reset is #if 0 there). */
int aidax_stream_activate(aidax_engine* e, int stream);

/* Bind a model (or -1 for none).  Mirrors work() + work_response() (:807-893) and the pre-warm of
 * loadModelFromPath (:1052-1079): recurrent state = reset() followed by 2048 zero-input steps with
 * the param smoothers parked at (old_param1, old_param2); paramFirstRun = true.  The pre-warm runs
 * on the GPU at the start of the next render call. */
int aidax_stream_set_model(aidax_engine* e, int stream, int model_id, float old_param1, float old_param2);

/* Latch the control-port values for the next render call; `loading` mutes the master gain target
 * (rt-neural-generic.cpp:654).  Only needs to be called when something changed. */
int aidax_stream_set_controls(aidax_engine* e, int stream, const float controls[AIDAX_N_CONTROLS], int loading);

/* Bulk variant of set_controls for PARAM1/PARAM2 automation: p1/p2 are [n] arrays for streams
 * first..first+n-1 (stream ids are dense small integers in open order). */
int aidax_streams_set_params(aidax_engine* e, int first_stream, int n, const float* p1, const float* p2);

/* Current targets of the PARAM1/PARAM2 smoothers of the bound model (0,0 if none): what the reference
 * reads with param1Coeff.getTargetValue() in work() (rt-neural-generic.cpp:822-825) to seed the next model. */
int aidax_stream_get_params(aidax_engine* e, int stream, float* target1, float* target2);

/* Test/debug observability: copy the recurrent state (h[hidden], c[hidden]; c only for LSTM) and
 * the 7x2 biquad states (order: in_lpf, depth, bass, mid, treble, presence, dc) of a stream. */
int aidax_stream_get_state(aidax_engine* e, int stream, float* h, float* c, double* biquad_z);
/* Inverse of get_state (NULL = leave as is): checkpoint/restore of the DSP state, which the reference
 * never persists (LV2 state only stores the model path, rt-neural-generic.cpp:763-796).  Synchronises. */
int aidax_stream_set_state(aidax_engine* e, int stream, const float* h, const float* c, const double* biquad_z);

/* ------------------------------------------------------------------ render ---- */

/* Render one audio period for streams [first_stream, first_stream + n_streams) (all OPEN streams in
 * that range; closed ids are skipped).  Replaces N calls of RtNeuralGeneric::run
 * (rt-neural-generic.cpp:484-660), stages :621-655 fused into one kernel per model family.
 *
 * aidax_render_host:   `in`/`out` are HOST buffers (pinned preferred, see aidax_host_alloc);
 *                      H2D copy, kernels, D2H copy and a final synchronise are inside the call.
 * aidax_render_device: `in`/`out` are DEVICE pointers on the engine's GPU; the call only
 *                      enqueues work on the engine's CUDA stream (no synchronise).
 * Row r of the buffers belongs to stream first_stream + r.  in may equal out (in place,
 * rt-neural-generic.cpp:613).  stride: samples per row (planar) or streams per row (interleaved).
 */
int aidax_render_host(aidax_engine* e, int first_stream, int n_streams, int n_samples,
                      const float* in, float* out, int layout, size_t stride);
int aidax_render_device(aidax_engine* e, int first_stream, int n_streams, int n_samples,
                        const float* d_in, float* d_out, int layout, size_t stride);
int aidax_engine_synchronize(aidax_engine* e);

/* The two halves of aidax_render_host, for a host that hands a large period over in pieces while its
 * instances are still running (the #batch extension does, include/aidax_batch.h): submit() enqueues the
 * H2D copy, the kernels and the D2H copy of one stream range on the engine's copy/kernel streams and
 * returns at once; wait() blocks until every range submitted since the last wait() is in its output
 * buffer.  Ranges of one period must not overlap; the buffers must stay valid until wait() returns.
 * Both may be called from any thread. */
int aidax_render_host_submit(aidax_engine* e, int first_stream, int n_streams, int n_samples,
                             const float* in, float* out, int layout, size_t stride);
int aidax_render_host_wait(aidax_engine* e);
/* submit() for a LARGE planar range that is complete at hand-over: the engine cuts it into pieces that are small at both
 * ends of the period (the first copy-in and the last kernel + copy-out overlap nothing) and large in between, and submits
 * them in order.  What the #batch extension's end() uses when the whole period arrived through run_block(). */
int aidax_render_host_submit_graded(aidax_engine* e, int first_stream, int n_streams, int n_samples,
                                    const float* in, float* out, size_t stride);

/* Pinned host memory for audio arenas (cudaHostAlloc / cudaFreeHost). */
void* aidax_host_alloc(size_t bytes);
void aidax_host_free(void* p);

/* Roofline denominator for the recurrence kernels: sustained FP32 FFMA throughput (TFLOP/s, 2 flop per
 * FFMA) of `device`, measured with a register-resident FFMA loop over all SMs for about `millis` ms. */
int aidax_measure_fp32_peak(int device, int millis, double* tflops_out);

/* Bench support (kernel K6): synthetic input generated on the device, so that a measurement's audio never
 * crosses PCIe.  x[s][n] = 0.5 sin(2 pi (20 t + 999 t^2) + 0.37 s) + 0.05 u(s, n), t = n / 48000 (the 20 Hz -> 20 kHz sweep of
 * SURVEY 8(d) plus counter-based uniform noise; aidadsp_lv2_b200.signals.sweep_hash_block is the bit-exact numpy mirror of
 * the noise and agrees with the sweep to one float rounding).  Rows/columns as in aidax_render_device; enqueued on
 * `cuda_stream` (e.g. aidax_engine_cuda_stream) of `device`. */
int aidax_generate_input(int device, void* cuda_stream, float* d_out, int first_stream, int n_streams, long long start_sample, int n_samples,
                         int layout, size_t stride, uint32_t seed);

/* Measured denominators for every pipe the chain kernels load (csrc/probes.cu), each from a register/shared-memory
 * resident saturating loop run for about millis/7 ms:
 *   out[0] FP32 FFMA TFLOP/s   out[1] XU MUFU.EX2 G lane-ops/s   out[2] FP64 DFMA TFLOP/s
 *   out[3] XU F2F (f32<->f64 conversion) G lane-ops/s
 *   out[4..6] tcgen05.mma kind::tf32 M=128, N=256 / 160 / 64 (K=8 per instruction, operands in shared memory) dense TFLOP/s
 *   out[7] SM clock in MHz while the tensor probe ran. */
int aidax_measure_pipe_peaks(int device, int millis, double out[8]);

/* Ceiling of the host-fed path (aidax_render_host): pinned-memory copy rates of `device` over a `bytes`-sized
 * buffer, GB/s: [0] host->device alone, [1] device->host alone, [2] per direction with both running at once. */
int aidax_measure_pcie(int device, size_t bytes, double gbs_out[3]);

/* PCI bus id of the engine's GPU ("0000:1b:00.0", cudaDeviceGetPCIBusId): lets a host place its threads and pinned
 * arenas on the GPU's NUMA node (/sys/bus/pci/devices/<id>/local_cpulist). */
int aidax_engine_pci_bus_id(aidax_engine* e, char* out, int cap);

/* Diagnostic: the launches the planner would issue for the streams [first_stream, first_stream + n_streams) in the
 * engine's current configuration, one text line per kernel family, largest footprint first:
 *   "<lstm|gru|none> H=<hidden> shape=<warp|fp32|tc1|tc2|tc3> streams=<n> ctas=<n> streams_per_cta=<n> threads=<n>"
 * (warp = warp-per-stream latency shape, fp32 = register-tiled, tc1..tc3 = tensor-core generations; tc3 with 128
 * streams per CTA is the one-tile shape).  Truncated to `cap` bytes. */
int aidax_engine_describe_plan(aidax_engine* e, int first_stream, int n_streams, char* out, int cap);

/* Number of kernel launches issued by the engine since creation (for bench accounting). */
uint64_t aidax_engine_launch_count(aidax_engine* e);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* AIDAX_H */
