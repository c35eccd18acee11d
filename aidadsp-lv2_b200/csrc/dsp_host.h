// dsp_host.h -- control-rate (per-buffer, per-stream) host arithmetic of the hot path.
//
// The reference evaluates these on the audio thread at the top of every run()
// (rt-neural-generic.cpp:489-517) and inside applyToneControls (:50-127); they run at control
// rate, use libm in double precision, and their results (5 coefficients per filter, two gain
// targets, a flag word) are what the GPU kernels consume.  Keeping them on the host with the same
// libm makes the filter/gain chain bit-identical to the reference build.
// Compile with -ffp-contract=off so no FMA contraction changes the rounding.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/aidax.h"
#include "layout.h"

namespace aidax {

enum BiquadType { BQT_LOWPASS = 0, BQT_HIGHPASS, BQT_BANDPASS, BQT_NOTCH, BQT_PEAK, BQT_LOWSHELF, BQT_HIGHSHELF };

struct BiquadCoef { double a0, a1, a2, b1, b2; };

// Nigel Redmon / EarLevel biquad design, restating Biquad::calcBiquad (common/Biquad.cpp:67-165).
// fc is normalised (f / fs).  Expression order follows the reference so results match bit for bit.
inline BiquadCoef design_biquad(int type, double fc, double q, double gain_db)
{
    BiquadCoef c{1.0, 0.0, 0.0, 0.0, 0.0};
    const double V = std::pow(10, std::fabs(gain_db) / 20.0);
    const double K = std::tan(M_PI * fc);
    const double KK = K * K;
    double norm;
    switch (type) {
    case BQT_LOWPASS:
        norm = 1 / (1 + K / q + KK);
        c.a0 = KK * norm; c.a1 = 2 * c.a0; c.a2 = c.a0;
        c.b1 = 2 * (KK - 1) * norm; c.b2 = (1 - K / q + KK) * norm;
        break;
    case BQT_HIGHPASS:
        norm = 1 / (1 + K / q + KK);
        c.a0 = 1 * norm; c.a1 = -2 * c.a0; c.a2 = c.a0;
        c.b1 = 2 * (KK - 1) * norm; c.b2 = (1 - K / q + KK) * norm;
        break;
    case BQT_BANDPASS:
        norm = 1 / (1 + K / q + KK);
        c.a0 = K / q * norm; c.a1 = 0; c.a2 = -c.a0;
        c.b1 = 2 * (KK - 1) * norm; c.b2 = (1 - K / q + KK) * norm;
        break;
    case BQT_NOTCH:
        norm = 1 / (1 + K / q + KK);
        c.a0 = (1 + KK) * norm; c.a1 = 2 * (KK - 1) * norm; c.a2 = c.a0;
        c.b1 = c.a1; c.b2 = (1 - K / q + KK) * norm;
        break;
    case BQT_PEAK:
        if (gain_db >= 0) {
            norm = 1 / (1 + 1 / q * K + KK);
            c.a0 = (1 + V / q * K + KK) * norm; c.a1 = 2 * (KK - 1) * norm; c.a2 = (1 - V / q * K + KK) * norm;
            c.b1 = c.a1; c.b2 = (1 - 1 / q * K + KK) * norm;
        } else {
            norm = 1 / (1 + V / q * K + KK);
            c.a0 = (1 + 1 / q * K + KK) * norm; c.a1 = 2 * (KK - 1) * norm; c.a2 = (1 - 1 / q * K + KK) * norm;
            c.b1 = c.a1; c.b2 = (1 - V / q * K + KK) * norm;
        }
        break;
    case BQT_LOWSHELF:
        if (gain_db >= 0) {
            norm = 1 / (1 + std::sqrt(2) * K + KK);
            c.a0 = (1 + std::sqrt(2 * V) * K + V * K * K) * norm; c.a1 = 2 * (V * K * K - 1) * norm;
            c.a2 = (1 - std::sqrt(2 * V) * K + V * K * K) * norm;
            c.b1 = 2 * (KK - 1) * norm; c.b2 = (1 - std::sqrt(2) * K + KK) * norm;
        } else {
            norm = 1 / (1 + std::sqrt(2 * V) * K + V * K * K);
            c.a0 = (1 + std::sqrt(2) * K + KK) * norm; c.a1 = 2 * (KK - 1) * norm; c.a2 = (1 - std::sqrt(2) * K + KK) * norm;
            c.b1 = 2 * (V * K * K - 1) * norm; c.b2 = (1 - std::sqrt(2 * V) * K + V * K * K) * norm;
        }
        break;
    case BQT_HIGHSHELF:
        if (gain_db >= 0) {
            norm = 1 / (1 + std::sqrt(2) * K + KK);
            c.a0 = (V + std::sqrt(2 * V) * K + KK) * norm; c.a1 = 2 * (KK - V) * norm;
            c.a2 = (V - std::sqrt(2 * V) * K + KK) * norm;
            c.b1 = 2 * (KK - 1) * norm; c.b2 = (1 - std::sqrt(2) * K + KK) * norm;
        } else {
            norm = 1 / (V + std::sqrt(2 * V) * K + KK);
            c.a0 = (1 + std::sqrt(2) * K + KK) * norm; c.a1 = 2 * (KK - 1) * norm; c.a2 = (1 - std::sqrt(2) * K + KK) * norm;
            c.b1 = 2 * (KK - V) * norm; c.b2 = (V - std::sqrt(2 * V) * K + KK) * norm;
        }
        break;
    default:
        break;
    }
    return c;
}

// DB_CO, rt-neural-generic.h:160
inline float db_to_coef(float g) { return g > -90.0f ? powf(10.0f, g * 0.05f) : 0.0f; }

// MAP(pc, 0, 100, INLPF_MAX_CO, INLPF_MIN_CO), rt-neural-generic.h:167, 177-179 (all float arithmetic)
inline float lpf_percent_to_fc(float pc)
{
    const float out_min = 0.99f * 0.5f, out_max = 0.25f * 0.5f;
    return (((pc - 0.0f) * (out_max - out_min) / (100.0f - 0.0f)) + out_min);
}

// ExponentialValueSmoother::updateCoef with T60 = 0.1 s, ValueSmoother.hpp:106-115, 148-151
inline float exp_smoother_coef(double host_samplerate)
{
    const float sr = (float)host_samplerate;
    const float tau = 0.1f * (float)(1.0 / 6.91);
    return std::exp(-1.f / (tau * sr));
}

// d_isNotEqual<float>, ValueSmoother.hpp:41-46
inline bool float_not_equal(float a, float b) { return std::abs(a - b) >= FLT_EPSILON; }

// Host shadow of one stream: everything run() keeps in RtNeuralGeneric members that is decided
// at control rate (rt-neural-generic.h:198-239).
struct StreamHost {
    bool open = false;
    double samplerate = 48000.0;
    int model = -1;          // engine model id or -1
    bool loading = true;     // :318
    float ctrl[AIDAX_N_CONTROLS];
    bool have_ctrl = false;
    // *_old change-detection members
    float in_lpf_pc_old = 66.216f;
    float bass_db_old = 0.f, bass_freq_old = 250.f;
    float mid_db_old = 0.f, mid_freq_old = 600.f, mid_q_old = 0.707f, mid_type_old = 0.f;
    float treble_db_old = 0.f, treble_freq_old = 1500.f;
    float depth_db_old = 0.f, presence_db_old = 0.f;
    // decoded values (mirrors of what is uploaded)
    uint32_t flags = 0;
    float pre_target = 1.f;    // preGain.setTargetValue(1.f) at instantiate
    float master_target = 0.f; // masterGain.setTargetValue(0.f) at instantiate; decoded eagerly (what the next enabled run() will set)
    float master_seen = 0.f;   // the target the reference's masterGain really holds: only an ENABLED run() with n_samples > 0 sets it (:654)
    float param_target[2] = {0.f, 0.f}; // LinearValueSmoother targets as the reference would hold them
    BiquadCoef bq[kNumBiquads];
};

// Default control values = ttl defaults (ttl/rt-neural-generic.ttl:92-313)
inline void default_controls(float c[AIDAX_N_CONTROLS])
{
    std::memset(c, 0, sizeof(float) * AIDAX_N_CONTROLS);
    c[AIDAX_CTRL_ANTIALIASING] = 66.216f;
    c[AIDAX_CTRL_BFREQ] = 305.f;
    c[AIDAX_CTRL_MFREQ] = 750.f;
    c[AIDAX_CTRL_MIDQ] = 0.707f;
    c[AIDAX_CTRL_TFREQ] = 2000.f;
    c[AIDAX_CTRL_DCBLOCKER] = 1.f;
    c[AIDAX_CTRL_ENABLED] = 1.f;
}

// instantiate-time filters (rt-neural-generic.cpp:292-314)
inline void init_stream_host(StreamHost& s, double samplerate)
{
    s = StreamHost();
    s.open = true;
    s.samplerate = samplerate;
    default_controls(s.ctrl);
    s.bq[BQ_DC] = design_biquad(BQT_HIGHPASS, 35.0f / samplerate, 0.707f, 0.0f);
    s.bq[BQ_INLPF] = design_biquad(BQT_LOWPASS, lpf_percent_to_fc(s.in_lpf_pc_old), 0.707f, 0.0f);
    s.bq[BQ_BASS] = design_biquad(BQT_LOWSHELF, s.bass_freq_old / samplerate, 0.707f, s.bass_db_old);
    s.bq[BQ_MID] = design_biquad(BQT_PEAK, s.mid_freq_old / samplerate, s.mid_q_old, s.mid_db_old);
    s.bq[BQ_TREBLE] = design_biquad(BQT_HIGHSHELF, s.treble_freq_old / samplerate, 0.707f, s.treble_db_old);
    s.bq[BQ_DEPTH] = design_biquad(BQT_PEAK, 75.0f / samplerate, 0.707f, s.depth_db_old);
    s.bq[BQ_PRESENCE] = design_biquad(BQT_HIGHSHELF, 900.0f / samplerate, 0.707f, s.presence_db_old);
}

// Decode one run()'s worth of control ports.  Returns a bit mask of which biquads were redesigned.
// Mirrors rt-neural-generic.cpp:489-517 (top of run) and :53-127 (applyToneControls change
// detection).  The EQ refresh is done eagerly: the coefficients are a pure function of the current
// port values, so refreshing them while the EQ is bypassed is unobservable.
inline uint32_t decode_controls(StreamHost& s, const float c[AIDAX_N_CONTROLS], bool loading, bool model_bound)
{
    uint32_t redesigned = 0;
    if (c != s.ctrl) std::memcpy(s.ctrl, c, sizeof(float) * AIDAX_N_CONTROLS); // (callers re-decode a stream's own latched row)
    s.have_ctrl = true;
    s.loading = loading;

    const float pregain = db_to_coef(c[AIDAX_CTRL_PREGAIN]);
    const float master = db_to_coef(c[AIDAX_CTRL_MASTER]);
    const bool net_bypass = c[AIDAX_CTRL_NETBYPASS] > 0.5f;
    const float in_lpf_pc = c[AIDAX_CTRL_ANTIALIASING];
    const float eq_position = c[AIDAX_CTRL_EQPOS];
    const float eq_bypass = c[AIDAX_CTRL_EQBYPASS];
    const bool enabled = c[AIDAX_CTRL_ENABLED] > 0.5f;

    s.pre_target = pregain;
    if (in_lpf_pc != s.in_lpf_pc_old) {
        s.bq[BQ_INLPF] = design_biquad(BQT_LOWPASS, lpf_percent_to_fc(in_lpf_pc), 0.707f, 0.0f);
        s.in_lpf_pc_old = in_lpf_pc;
        redesigned |= 1u << BQ_INLPF;
    }

    const double sr = s.samplerate;
    const float bass_db = c[AIDAX_CTRL_BASS], bass_freq = c[AIDAX_CTRL_BFREQ];
    if (bass_db != s.bass_db_old || bass_freq != s.bass_freq_old) {
        s.bass_db_old = bass_db; s.bass_freq_old = bass_freq;
        s.bq[BQ_BASS] = design_biquad(BQT_LOWSHELF, bass_freq / sr, 0.707f, bass_db);
        redesigned |= 1u << BQ_BASS;
    }
    const float mid_db = c[AIDAX_CTRL_MID], mid_freq = c[AIDAX_CTRL_MFREQ], mid_q = c[AIDAX_CTRL_MIDQ], mid_type = c[AIDAX_CTRL_MTYPE];
    if (mid_db != s.mid_db_old || mid_freq != s.mid_freq_old || mid_q != s.mid_q_old || mid_type != s.mid_type_old) {
        s.mid_db_old = mid_db; s.mid_freq_old = mid_freq; s.mid_q_old = mid_q; s.mid_type_old = mid_type;
        s.bq[BQ_MID] = design_biquad(mid_type == 1.0f ? BQT_BANDPASS : BQT_PEAK, mid_freq / sr, mid_q, mid_db);
        redesigned |= 1u << BQ_MID;
    }
    const float treble_db = c[AIDAX_CTRL_TREBLE], treble_freq = c[AIDAX_CTRL_TFREQ];
    if (treble_db != s.treble_db_old || treble_freq != s.treble_freq_old) {
        s.treble_db_old = treble_db; s.treble_freq_old = treble_freq;
        s.bq[BQ_TREBLE] = design_biquad(BQT_HIGHSHELF, treble_freq / sr, 0.707f, treble_db);
        redesigned |= 1u << BQ_TREBLE;
    }
    const float depth_db = c[AIDAX_CTRL_DEPTH];
    if (depth_db != s.depth_db_old) {
        s.depth_db_old = depth_db;
        s.bq[BQ_DEPTH] = design_biquad(BQT_PEAK, 75.0f / sr, 0.707f, depth_db);
        redesigned |= 1u << BQ_DEPTH;
    }
    const float presence_db = c[AIDAX_CTRL_PRESENCE];
    if (presence_db != s.presence_db_old) {
        s.presence_db_old = presence_db;
        s.bq[BQ_PRESENCE] = design_biquad(BQT_HIGHSHELF, 900.0f / sr, 0.707f, presence_db);
        redesigned |= 1u << BQ_PRESENCE;
    }

    uint32_t f = 0;
    if (enabled) f |= F_ENABLED;
    if (in_lpf_pc != 0.0f) f |= F_LPF;
    if (eq_position == 1.0f && eq_bypass == 0.0f) f |= F_EQ_PRE;
    if (eq_position == 0.0f && eq_bypass == 0.0f) f |= F_EQ_POST;
    if (mid_type == 1.0f) f |= F_EQ_BANDPASS;
    if (c[AIDAX_CTRL_DCBLOCKER] == 1.0f) f |= F_DC;
    if (model_bound && !net_bypass) f |= F_NET;
    s.flags = f;
    // masterGain.setTargetValue only happens on the enabled path with n_samples > 0 (:606-619, 654).  The kernels read
    // master_target only for such runs, so it is decoded eagerly; activate() -- which snaps the smoother to the target
    // the reference holds (:341-342) -- uses master_seen, maintained by track_run_targets after each render.
    s.master_target = loading ? 0.f : master;
    return redesigned;
}

// What the reference's LinearValueSmoother targets would be after a processed buffer
// (rt-neural-generic.cpp:634-635 + ValueSmoother.hpp:209-216); used for old_param hand-over
// to the next model (work(), :822-825).
inline void track_run_targets(StreamHost& s, int n_samples)
{
    if (n_samples <= 0 || !(s.flags & F_ENABLED)) return;
    s.master_seen = s.master_target;
    if (!(s.flags & F_NET)) return;
    for (int i = 0; i < 2; ++i) {
        const float v = s.ctrl[AIDAX_CTRL_PARAM1 + i];
        if (float_not_equal(s.param_target[i], v)) s.param_target[i] = v;
    }
}

} // namespace aidax
