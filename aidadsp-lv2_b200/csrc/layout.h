// layout.h -- data layout in HBM shared by the host engine and the sm_100a kernels.
//
// Every per-stream quantity is a structure-of-arrays indexed by the stream's slot
// (slot == stream id, dense small integers), so that the 32 streams a warp owns are read and
// written with coalesced 128-byte transactions.  Host-written arrays ("controls") are uploaded
// as dirty ranges before a render call; device-owned arrays ("state") are only touched by kernels.
#pragma once
#include <stdint.h>

namespace aidax {

constexpr int kMaxHidden = 80;   // largest variant (variant/generate_variant_hpp.py:6)
constexpr int kMaxInput = 3;     // MAX_INPUT_SIZE (model_variant.hpp:4)
constexpr int kNumBiquads = 7;   // rt-neural-generic.h:311-317
constexpr int kPrewarmSteps = 2048; // rt-neural-generic.cpp:1077-1078

// biquad order inside the coefficient/state arrays
enum { BQ_INLPF = 0, BQ_DEPTH = 1, BQ_BASS = 2, BQ_MID = 3, BQ_TREBLE = 4, BQ_PRESENCE = 5, BQ_DC = 6 };

// per-stream flag word (host-written)
enum : uint32_t {
    F_ENABLED = 1u << 0,   // *enabled > 0.5f                         (rt-neural-generic.cpp:495, 612)
    F_LPF = 1u << 1,       // in_lpf_pc != 0.0f                       (:622)
    F_EQ_PRE = 1u << 2,    // eq_position == 1 && eq_bypass == 0      (:628)
    F_EQ_POST = 1u << 3,   // eq_position == 0 && eq_bypass == 0      (:651)
    F_EQ_BANDPASS = 1u << 4, // mid_type == BANDPASS: only `mid` runs (:130-132)
    F_DC = 1u << 5,        // *dc_blocker_param == 1.0f               (:646)
    F_NET = 1u << 6,       // model != nullptr && !net_bypass         (:631-632)
};

// per-stream command word (host-written, executed and cleared by the next render kernel)
enum : uint32_t {
    C_SET_GAIN_MEMS = 1u << 0, // pre_mem/master_mem := cmd_pre_mem/cmd_master_mem (instantiate :283-290, activate :341-342)
    C_ZERO_FILTERS = 1u << 1,  // all biquad z1,z2 := 0 (Biquad ctor, Biquad.cpp:32-35)
    C_SET_MODEL = 1u << 2,     // new model bound: h=c=0, param smoothers parked at p_init, 2048-step pre-warm (:1052-1079)
    C_FIRST_RUN = 1u << 3,     // paramFirstRun := true (activate :349-351, load :1061)
};

// device-side per-stream flags (device-owned)
enum : uint32_t { D_PARAM_FIRST_RUN = 1u << 0 };

struct StreamArrays {
    // ---- host-written controls
    uint32_t* flags;      // [cap]
    uint32_t* cmd;        // [cap]
    float* pre_target;    // [cap]  DB_CO(PREGAIN)
    float* master_target; // [cap]  loading ? 0 : DB_CO(MASTER)
    float* p_port;        // [2][cap] raw PARAM1/PARAM2 port values
    float* p_init;        // [2][cap] old_param1/2 handed to a new model (C_SET_MODEL)
    float* cmd_pre_mem;   // [cap]
    float* cmd_master_mem;// [cap]
    float* exp_coef;      // [cap]  exp(-1/(tau*fs)), ValueSmoother.hpp:148-151
    float* lin_denom;     // [cap]  tau*model_samplerate, ValueSmoother.hpp:237-240
    double* bq_coef;      // [7*5][cap]  a0,a1,a2,b1,b2 per biquad
    // ---- device-owned state
    uint32_t* dflags;     // [cap]
    float* pre_mem;       // [cap]
    float* master_mem;    // [cap]
    float* p_mem;         // [2][cap]
    float* p_step;        // [2][cap]
    float* p_target;      // [2][cap]
    double* bq_z;         // [7*2][cap]  z1,z2 per biquad
    float* rnn;           // [2*kMaxHidden][cap]: rows 0..H-1 = h, rows kMaxHidden.. = c (LSTM)
    int cap;
};

// one registered model as the kernels see it
struct ModelDev {
    const float* weights; // packed, see pack_model() in model_host.h
    float in_gain;
    float out_gain;
    int in_skip;
    int input_size;
};

// packed weight block (floats):  wk[(H+3)][H][4] | bias[H][4] | dense_w[H] | dense_b | pad
//   rows 0..H-1 of wk are the recurrent rows (U), rows H..H+2 the input rows (W, zero if unused)
//   LSTM lanes: (i, f, g, o);  GRU lanes: (z, r, Uh.h [+bh1], Wh.x [+bh0])
inline int packed_weight_floats(int H) { return (H + 3) * H * 4 + H * 4 + H + 4; }

struct RenderParams {
    const float* in;
    float* out;
    int layout;        // AIDAX_LAYOUT_*
    long long stride;
    int first_stream;  // buffer row r <-> slot first_stream + r
    int n_samples;
    int ts;            // samples per shared-memory sub-block
    int math_mode;     // 0 fast, 1 precise
    const int* blk_model;  // [nblk]
    const int* blk_slots;  // [nblk][CS], -1 = empty
    const ModelDev* models;
    StreamArrays sa;
};

} // namespace aidax
