/*
 * aidax_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-stream CPU restatement of the rt-neural-generic hot path of
 * AidaDSP/aidadsp-lv2 @ f86e21d, used only as the checker in tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.  Nothing under aidadsp-lv2_b200/ links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against (1) the six bundled
 * models' input_batch/output_batch vectors (models/deer ink studios/*.json, compared at 1e-5 by
 * the reference itself, rt-neural-generic.cpp:900-955, rt-neural-generic.h:182), (2) RTNeural's
 * lstm_1d / gru_1d Keras vectors (modules/RTNeural/test_data/*.csv, test_configs.hpp:15-30), and
 * (3) outputs of the unmodified reference compiled into oracle/_ref (tiers A/B/C) for the stages
 * no reference vector covers (filters, smoothers, conditioned models, run() end to end).
 *
 * Arithmetic follows the reference's default build (xsimd, SSE2, no FMA contraction):
 * compile with -ffp-contract=off.  exp/tanh come from libm instead of xsimd's polynomial
 * kernels (xsimd_generic_math.hpp:663-686, xsimd_generic_trigo.hpp:867-952); both are ~1 ulp.
 *
 * Each function cites the reference lines it follows.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_H 80
#define ORACLE_N_CONTROLS 21

/* control indices: LV2 port index - 4 (ports_t, rt-neural-generic.h:84-112) */
enum {
    C_ANTIALIASING = 0, C_PREGAIN, C_NETBYPASS, C_PARAM1, C_PARAM2, C_EQBYPASS, C_EQPOS, C_BASS, C_BFREQ, C_MID,
    C_MFREQ, C_MIDQ, C_MTYPE, C_TREBLE, C_TFREQ, C_DEPTH, C_PRESENCE, C_DCBLOCKER, C_MASTER, C_MODELINSIZE, C_ENABLED
};

/* ------------------------------------------------------------------ Biquad (common/Biquad.{h,cpp}) */
enum { BQ_LOWPASS = 0, BQ_HIGHPASS, BQ_BANDPASS, BQ_NOTCH, BQ_PEAK, BQ_LOWSHELF, BQ_HIGHSHELF };

typedef struct {
    double a0, a1, a2, b1, b2; /* coefficients */
    double z1, z2;             /* transposed direct form II state */
} OracleBiquad;

/* Biquad::calcBiquad, common/Biquad.cpp:67-165.  fc normalised to the sample rate. */
void oracle_biquad_design(OracleBiquad* f, int type, double fc, double q, double gain_db)
{
    double norm;
    const double V = pow(10, fabs(gain_db) / 20.0);
    const double K = tan(M_PI * fc);
    switch (type) {
    case BQ_LOWPASS:
        norm = 1 / (1 + K / q + K * K);
        f->a0 = K * K * norm; f->a1 = 2 * f->a0; f->a2 = f->a0;
        f->b1 = 2 * (K * K - 1) * norm; f->b2 = (1 - K / q + K * K) * norm;
        break;
    case BQ_HIGHPASS:
        norm = 1 / (1 + K / q + K * K);
        f->a0 = 1 * norm; f->a1 = -2 * f->a0; f->a2 = f->a0;
        f->b1 = 2 * (K * K - 1) * norm; f->b2 = (1 - K / q + K * K) * norm;
        break;
    case BQ_BANDPASS:
        norm = 1 / (1 + K / q + K * K);
        f->a0 = K / q * norm; f->a1 = 0; f->a2 = -f->a0;
        f->b1 = 2 * (K * K - 1) * norm; f->b2 = (1 - K / q + K * K) * norm;
        break;
    case BQ_NOTCH:
        norm = 1 / (1 + K / q + K * K);
        f->a0 = (1 + K * K) * norm; f->a1 = 2 * (K * K - 1) * norm; f->a2 = f->a0;
        f->b1 = f->a1; f->b2 = (1 - K / q + K * K) * norm;
        break;
    case BQ_PEAK:
        if (gain_db >= 0) {
            norm = 1 / (1 + 1 / q * K + K * K);
            f->a0 = (1 + V / q * K + K * K) * norm; f->a1 = 2 * (K * K - 1) * norm; f->a2 = (1 - V / q * K + K * K) * norm;
            f->b1 = f->a1; f->b2 = (1 - 1 / q * K + K * K) * norm;
        } else {
            norm = 1 / (1 + V / q * K + K * K);
            f->a0 = (1 + 1 / q * K + K * K) * norm; f->a1 = 2 * (K * K - 1) * norm; f->a2 = (1 - 1 / q * K + K * K) * norm;
            f->b1 = f->a1; f->b2 = (1 - V / q * K + K * K) * norm;
        }
        break;
    case BQ_LOWSHELF:
        if (gain_db >= 0) {
            norm = 1 / (1 + sqrt(2) * K + K * K);
            f->a0 = (1 + sqrt(2 * V) * K + V * K * K) * norm; f->a1 = 2 * (V * K * K - 1) * norm;
            f->a2 = (1 - sqrt(2 * V) * K + V * K * K) * norm;
            f->b1 = 2 * (K * K - 1) * norm; f->b2 = (1 - sqrt(2) * K + K * K) * norm;
        } else {
            norm = 1 / (1 + sqrt(2 * V) * K + V * K * K);
            f->a0 = (1 + sqrt(2) * K + K * K) * norm; f->a1 = 2 * (K * K - 1) * norm; f->a2 = (1 - sqrt(2) * K + K * K) * norm;
            f->b1 = 2 * (V * K * K - 1) * norm; f->b2 = (1 - sqrt(2 * V) * K + V * K * K) * norm;
        }
        break;
    case BQ_HIGHSHELF:
        if (gain_db >= 0) {
            norm = 1 / (1 + sqrt(2) * K + K * K);
            f->a0 = (V + sqrt(2 * V) * K + K * K) * norm; f->a1 = 2 * (K * K - V) * norm;
            f->a2 = (V - sqrt(2 * V) * K + K * K) * norm;
            f->b1 = 2 * (K * K - 1) * norm; f->b2 = (1 - sqrt(2) * K + K * K) * norm;
        } else {
            norm = 1 / (V + sqrt(2 * V) * K + K * K);
            f->a0 = (1 + sqrt(2) * K + K * K) * norm; f->a1 = 2 * (K * K - 1) * norm; f->a2 = (1 - sqrt(2) * K + K * K) * norm;
            f->b1 = 2 * (K * K - V) * norm; f->b2 = (V - sqrt(2 * V) * K + K * K) * norm;
        }
        break;
    }
}

/* test hook: coefficients only */
void oracle_biquad_coefs(int type, double fc, double q, double gain_db, double out[5])
{
    OracleBiquad f;
    memset(&f, 0, sizeof(f));
    oracle_biquad_design(&f, type, fc, q, gain_db);
    out[0] = f.a0; out[1] = f.a1; out[2] = f.a2; out[3] = f.b1; out[4] = f.b2;
}

/* Biquad::process, common/Biquad.h:53-58 */
static inline float biquad_process(OracleBiquad* f, float in)
{
    const double out = in * f->a0 + f->z1;
    f->z1 = in * f->a1 + f->z2 - f->b1 * out;
    f->z2 = in * f->a2 - f->b2 * out;
    return (float)out;
}

/* applyBiquadFilter, rt-neural-generic.cpp:42-46 */
static void apply_biquad(float* out, const float* in, OracleBiquad* f, uint32_t n)
{
    for (uint32_t i = 0; i < n; ++i) out[i] = biquad_process(f, in[i]);
}

/* ------------------------------------------------------- smoothers (common/ValueSmoother.hpp) */
typedef struct { float coef, target, mem, tau, sr; } ExpSmoother;   /* :82-152 */
typedef struct { float step, target, mem, tau, sr; } LinSmoother;   /* :166-241 */

static int f_not_equal(float a, float b) { return fabsf(a - b) >= FLT_EPSILON; } /* :41-46 */

static void exp_update(ExpSmoother* s) { s->coef = expf(-1.f / (s->tau * s->sr)); }                         /* :148-151 */
static void exp_set_sr(ExpSmoother* s, float sr) { if (f_not_equal(s->sr, sr)) { s->sr = sr; exp_update(s); } } /* :98-105 */
static void exp_set_t60(ExpSmoother* s, float t60)                                                            /* :107-116 */
{
    const float tau = t60 * (float)(1.0 / 6.91);
    if (f_not_equal(s->tau, tau)) { s->tau = tau; exp_update(s); }
}
static inline float exp_next(ExpSmoother* s) { return (s->mem = s->mem * s->coef + s->target * (1.f - s->coef)); } /* :142-145 */

static void lin_update(LinSmoother* s) { s->step = (s->target - s->mem) / (s->tau * s->sr); }                /* :237-240 */
static void lin_set_sr(LinSmoother* s, float sr) { if (f_not_equal(s->sr, sr)) { s->sr = sr; lin_update(s); } }
static void lin_set_tau(LinSmoother* s, float tau) { if (f_not_equal(s->tau, tau)) { s->tau = tau; lin_update(s); } }
static void lin_set_target(LinSmoother* s, float t) { if (f_not_equal(s->target, t)) { s->target = t; lin_update(s); } } /* :209-216 */
static inline float lin_next(LinSmoother* s)                                                                  /* :229-234 */
{
    const float y0 = s->mem;
    const float dy = s->target - y0;
    return (s->mem = y0 + copysignf(fminf(fabsf(dy), fabsf(s->step)), dy));
}

/* ------------------------------------------------------------------------ the recurrent model */
typedef struct OracleModel {
    int cell;      /* 0 lstm, 1 gru */
    int H, in;
    int use_double;
    /* weights in the reference JSON layout */
    float* W;      /* [in][G*H] */
    float* U;      /* [H][G*H] */
    float* b;      /* lstm [4H] ; gru [2][3H] */
    float* dw;     /* [H] */
    float db;
    /* state */
    float h[ORACLE_MAX_H], c[ORACLE_MAX_H];
    double hd[ORACLE_MAX_H], cd[ORACLE_MAX_H];
} OracleModel;

OracleModel* oracle_model_create(int cell, int H, int in, const float* W, const float* U, const float* b, const float* dw,
                                 float db, int use_double)
{
    if (H < 1 || H > ORACLE_MAX_H || in < 1 || in > 3 || cell < 0 || cell > 1) return NULL;
    const int G = cell == 0 ? 4 : 3;
    OracleModel* m = (OracleModel*)calloc(1, sizeof(OracleModel));
    m->cell = cell; m->H = H; m->in = in; m->use_double = use_double; m->db = db;
    m->W = (float*)malloc(sizeof(float) * in * G * H);  memcpy(m->W, W, sizeof(float) * in * G * H);
    m->U = (float*)malloc(sizeof(float) * H * G * H);   memcpy(m->U, U, sizeof(float) * H * G * H);
    const int nb = cell == 0 ? 4 * H : 6 * H;
    m->b = (float*)malloc(sizeof(float) * nb);          memcpy(m->b, b, sizeof(float) * nb);
    m->dw = (float*)malloc(sizeof(float) * H);          memcpy(m->dw, dw, sizeof(float) * H);
    return m;
}

void oracle_model_destroy(OracleModel* m)
{
    if (!m) return;
    free(m->W); free(m->U); free(m->b); free(m->dw); free(m);
}

/* LSTMLayerT::reset / GRULayerT::reset, lstm_xsimd.tpp:189-207, gru_xsimd.tpp:273-285 */
void oracle_model_reset(OracleModel* m)
{
    memset(m->h, 0, sizeof(m->h)); memset(m->c, 0, sizeof(m->c));
    memset(m->hd, 0, sizeof(m->hd)); memset(m->cd, 0, sizeof(m->cd));
}

void oracle_model_get_state(const OracleModel* m, float* h, float* c)
{
    for (int j = 0; j < m->H; ++j) {
        h[j] = m->use_double ? (float)m->hd[j] : m->h[j];
        c[j] = m->use_double ? (float)m->cd[j] : m->c[j];
    }
}

static inline float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }    /* lstm_xsimd.h:347-350 */
static inline double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

/* One sample of ModelT<float,in,1,LSTMLayerT,DenseT>::forward.
 * Gate order i,f,c,o along 4H (lstm_xsimd.tpp:216-219).  Accumulation order of the reference:
 * recurrent_mat_mul starts from 0 and adds h[k]*U[k][j] for k ascending (lstm_xsimd.h:313-328);
 * in == 1: gate = W*x + (rec + b)  (xsimd::fma, :200)   in > 1: gate = rec + b + kernel  (:176),
 * kernel_mat_mul likewise from 0, i ascending (:330-345). */
static float lstm_step_f(OracleModel* m, const float* x)
{
    const int H = m->H, in = m->in, G4 = 4 * H;
    float pre[4][ORACLE_MAX_H];
    for (int g = 0; g < 4; ++g)
        for (int j = 0; j < H; ++j) {
            float rec = 0.f;
            for (int k = 0; k < H; ++k) rec += m->h[k] * m->U[k * G4 + g * H + j];
            if (in == 1) {
                pre[g][j] = m->W[g * H + j] * x[0] + (rec + m->b[g * H + j]);
            } else {
                float ker = 0.f;
                for (int i = 0; i < in; ++i) ker += x[i] * m->W[i * G4 + g * H + j];
                pre[g][j] = rec + m->b[g * H + j] + ker;
            }
        }
    float y = 0.f;
    for (int j = 0; j < H; ++j) {
        const float it = sigmoid_f(pre[0][j]), ft = sigmoid_f(pre[1][j]), ot = sigmoid_f(pre[3][j]);
        const float ct = it * tanhf(pre[2][j]) + ft * m->c[j];      /* computeOutputsInternal :271-283 */
        m->c[j] = ct;
        m->h[j] = ot * tanhf(ct);
    }
    for (int j = 0; j < H; ++j) y += m->h[j] * m->dw[j];            /* DenseT<T,H,1>::forward, dense_xsimd.h:253-260 */
    return y + m->db;
}

static float lstm_step_d(OracleModel* m, const float* x)
{
    const int H = m->H, in = m->in, G4 = 4 * H;
    double pre[4][ORACLE_MAX_H];
    for (int g = 0; g < 4; ++g)
        for (int j = 0; j < H; ++j) {
            double a = m->b[g * H + j];
            for (int k = 0; k < H; ++k) a += m->hd[k] * (double)m->U[k * G4 + g * H + j];
            for (int i = 0; i < in; ++i) a += (double)x[i] * (double)m->W[i * G4 + g * H + j];
            pre[g][j] = a;
        }
    double y = m->db;
    for (int j = 0; j < H; ++j) {
        const double it = sigmoid_d(pre[0][j]), ft = sigmoid_d(pre[1][j]), ot = sigmoid_d(pre[3][j]);
        const double ct = it * tanh(pre[2][j]) + ft * m->cd[j];
        m->cd[j] = ct;
        m->hd[j] = ot * tanh(ct);
    }
    for (int j = 0; j < H; ++j) y += m->hd[j] * (double)m->dw[j];
    return (float)y;
}

/* One sample of ModelT<float,in,1,GRULayerT,DenseT>::forward.  Gate order z,r,h (gru_xsimd.tpp:295-297);
 * bz, br = sum of both bias rows, bh0 = b[0][2H..], bh1 = b[1][2H..] (tpp:330-333);
 * h_hat = tanh(r*(Uh.h + bh1) + (Wh.x + bh0)) (gru_xsimd.h:213, 236); h' = (1-z)*h_hat + z*h (:270). */
static float gru_step_f(OracleModel* m, const float* x)
{
    const int H = m->H, in = m->in, G3 = 3 * H;
    const float* b0 = m->b;
    const float* b1 = m->b + G3;
    float hn[ORACLE_MAX_H];
    for (int j = 0; j < H; ++j) {
        float rz = 0.f, rr = 0.f, rh = 0.f;
        for (int k = 0; k < H; ++k) {
            rz += m->h[k] * m->U[k * G3 + j];
            rr += m->h[k] * m->U[k * G3 + H + j];
            rh += m->h[k] * m->U[k * G3 + 2 * H + j];
        }
        const float bz = b0[j] + b1[j], br = b0[H + j] + b1[H + j];
        float z, r, hh;
        if (in == 1) {
            z = sigmoid_f(m->W[j] * x[0] + (rz + bz));
            r = sigmoid_f(m->W[H + j] * x[0] + (rr + br));
            hh = tanhf(r * (rh + b1[2 * H + j]) + (m->W[2 * H + j] * x[0] + b0[2 * H + j]));
        } else {
            float kz = 0.f, kr = 0.f, kh = 0.f;
            for (int i = 0; i < in; ++i) {
                kz += x[i] * m->W[i * G3 + j];
                kr += x[i] * m->W[i * G3 + H + j];
                kh += x[i] * m->W[i * G3 + 2 * H + j];
            }
            z = sigmoid_f(rz + bz + kz);
            r = sigmoid_f(rr + br + kr);
            hh = tanhf(r * (rh + b1[2 * H + j]) + (b0[2 * H + j] + kh));
        }
        hn[j] = (1.0f - z) * hh + z * m->h[j];
    }
    float y = 0.f;
    for (int j = 0; j < H; ++j) { m->h[j] = hn[j]; y += hn[j] * m->dw[j]; }
    return y + m->db;
}

static float gru_step_d(OracleModel* m, const float* x)
{
    const int H = m->H, in = m->in, G3 = 3 * H;
    const float* b0 = m->b;
    const float* b1 = m->b + G3;
    double hn[ORACLE_MAX_H];
    for (int j = 0; j < H; ++j) {
        double az = (double)b0[j] + (double)b1[j], ar = (double)b0[H + j] + (double)b1[H + j];
        double ahu = b1[2 * H + j], ahw = b0[2 * H + j];
        for (int k = 0; k < H; ++k) {
            az += m->hd[k] * (double)m->U[k * G3 + j];
            ar += m->hd[k] * (double)m->U[k * G3 + H + j];
            ahu += m->hd[k] * (double)m->U[k * G3 + 2 * H + j];
        }
        for (int i = 0; i < in; ++i) {
            az += (double)x[i] * (double)m->W[i * G3 + j];
            ar += (double)x[i] * (double)m->W[i * G3 + H + j];
            ahw += (double)x[i] * (double)m->W[i * G3 + 2 * H + j];
        }
        const double z = sigmoid_d(az), r = sigmoid_d(ar);
        hn[j] = (1.0 - z) * tanh(r * ahu + ahw) + z * m->hd[j];
    }
    double y = m->db;
    for (int j = 0; j < H; ++j) { m->hd[j] = hn[j]; y += hn[j] * (double)m->dw[j]; }
    return (float)y;
}

/* ModelT::forward (RTN/ModelT.h:258-309); x has `in` floats */
float oracle_model_step(OracleModel* m, const float* x)
{
    if (m->cell == 0) return m->use_double ? lstm_step_d(m, x) : lstm_step_f(m, x);
    return m->use_double ? gru_step_d(m, x) : gru_step_f(m, x);
}

/* raw forward over n samples; in: [n][in_size] */
void oracle_model_forward(OracleModel* m, const float* in, float* out, int n)
{
    for (int i = 0; i < n; ++i) out[i] = oracle_model_step(m, in + (size_t)i * m->in);
}

/* ------------------------------------------------------------------------------ one stream */
typedef struct OracleStream {
    double samplerate;
    ExpSmoother pre_gain, master_gain;
    OracleBiquad dc_blocker, in_lpf, bass, mid, treble, depth, presence;
    float in_lpf_pc_old;
    float bass_db_old, bass_freq_old, mid_db_old, mid_freq_old, mid_q_old, mid_type_old;
    float treble_db_old, treble_freq_old, depth_db_old, presence_db_old;
    int loading;
    int last_input_size;
    /* DynamicModel, rt-neural-generic.h:115-129 */
    OracleModel* model;
    int in_skip;
    float in_gain, out_gain, model_sr;
    LinSmoother p1, p2;
    int param_first_run;
} OracleStream;

static float db_co(float g) { return g > -90.0f ? powf(10.0f, g * 0.05f) : 0.0f; }  /* DB_CO, rt-neural-generic.h:160 */

/* MAP(x,0,100,INLPF_MAX_CO,INLPF_MIN_CO), rt-neural-generic.h:167,177-179 */
static float lpf_fc(float pc)
{
    const float omin = 0.99f * 0.5f, omax = 0.25f * 0.5f;
    return (((pc - 0.0f) * (omax - omin) / (100.0f - 0.0f)) + omin);
}

/* RtNeuralGeneric::instantiate, rt-neural-generic.cpp:244-333 (DSP members only) */
OracleStream* oracle_stream_create(double samplerate)
{
    OracleStream* s = (OracleStream*)calloc(1, sizeof(OracleStream));
    s->samplerate = samplerate;
    exp_set_sr(&s->pre_gain, (float)samplerate); exp_set_t60(&s->pre_gain, 0.1f);
    s->pre_gain.target = 1.f; s->pre_gain.mem = s->pre_gain.target;
    exp_set_sr(&s->master_gain, (float)samplerate); exp_set_t60(&s->master_gain, 0.1f);
    s->master_gain.target = 0.f; s->master_gain.mem = s->master_gain.target;
    oracle_biquad_design(&s->dc_blocker, BQ_HIGHPASS, 35.0f / samplerate, 0.707f, 0.0f);
    s->in_lpf_pc_old = 66.216f;
    oracle_biquad_design(&s->in_lpf, BQ_LOWPASS, lpf_fc(s->in_lpf_pc_old), 0.707f, 0.0f);
    s->bass_db_old = 0.f; s->bass_freq_old = 250.f;
    oracle_biquad_design(&s->bass, BQ_LOWSHELF, s->bass_freq_old / samplerate, 0.707f, s->bass_db_old);
    s->mid_db_old = 0.f; s->mid_freq_old = 600.f; s->mid_q_old = 0.707f; s->mid_type_old = 0.f;
    oracle_biquad_design(&s->mid, BQ_PEAK, s->mid_freq_old / samplerate, s->mid_q_old, s->mid_db_old);
    s->treble_db_old = 0.f; s->treble_freq_old = 1500.f;
    oracle_biquad_design(&s->treble, BQ_HIGHSHELF, s->treble_freq_old / samplerate, 0.707f, s->treble_db_old);
    s->depth_db_old = 0.f;
    oracle_biquad_design(&s->depth, BQ_PEAK, 75.0f / samplerate, 0.707f, s->depth_db_old);
    s->presence_db_old = 0.f;
    oracle_biquad_design(&s->presence, BQ_HIGHSHELF, 900.0f / samplerate, 0.707f, s->presence_db_old);
    s->last_input_size = 0;
    s->loading = 1;
    s->model = NULL;
    return s;
}

void oracle_stream_destroy(OracleStream* s)
{
    if (!s) return;
    oracle_model_destroy(s->model);
    free(s);
}

/* applyModel, rt-neural-generic.cpp:148-240 */
static void apply_model(OracleStream* s, float* out, uint32_t n)
{
    OracleModel* m = s->model;
    float v[3] = {0.f, 0.f, 0.f};
    for (uint32_t i = 0; i < n; ++i) {
        out[i] *= s->in_gain;
        v[0] = out[i];
        if (m->in >= 2) v[1] = lin_next(&s->p1);
        if (m->in >= 3) v[2] = lin_next(&s->p2);
        const float y = oracle_model_step(m, v);
        if (s->in_skip) out[i] += y; else out[i] = y;
        out[i] *= s->out_gain;
    }
}

/* work() + loadModelFromPath + work_response (rt-neural-generic.cpp:807-893, 1022-1084): takes ownership of m.
 * The old params are read from the CURRENT model's smoothers (:822-825) by the caller-visible getter below. */
void oracle_stream_set_model(OracleStream* s, OracleModel* m, int in_skip, float in_gain, float out_gain, float model_sr)
{
    float old1 = 0.f, old2 = 0.f;
    if (s->model) { old1 = s->p1.target; old2 = s->p2.target; }
    oracle_model_destroy(s->model);
    s->model = m;
    if (!m) { s->last_input_size = 0; return; }
    oracle_model_reset(m);
    s->in_skip = in_skip != 0; s->in_gain = in_gain; s->out_gain = out_gain; s->model_sr = model_sr;
    memset(&s->p1, 0, sizeof(s->p1)); memset(&s->p2, 0, sizeof(s->p2));
    lin_set_sr(&s->p1, model_sr); lin_set_tau(&s->p1, 0.1f); lin_set_target(&s->p1, old1); s->p1.mem = s->p1.target;
    lin_set_sr(&s->p2, model_sr); lin_set_tau(&s->p2, 0.1f); lin_set_target(&s->p2, old2); s->p2.mem = s->p2.target;
    s->param_first_run = 1;
    float warm[2048];
    memset(warm, 0, sizeof(warm));
    apply_model(s, warm, 2048);                     /* pre-warm, :1077-1078 */
    s->last_input_size = m->in;
    s->loading = 0;                                 /* work_response :889 */
}

void oracle_stream_set_loading(OracleStream* s, int loading) { s->loading = loading; }
void oracle_stream_get_params(const OracleStream* s, float* t1, float* t2) { *t1 = s->p1.target; *t2 = s->p2.target; }

/* RtNeuralGeneric::activate, rt-neural-generic.cpp:337-365 */
void oracle_stream_activate(OracleStream* s)
{
    s->pre_gain.mem = s->pre_gain.target;
    s->master_gain.mem = s->master_gain.target;
    if (s->model) s->param_first_run = 1;
}

/* applyToneControls, rt-neural-generic.cpp:50-140 */
static void apply_tone_controls(OracleStream* s, const float* c, float* buf, uint32_t n)
{
    const double sr = s->samplerate;
    if (c[C_BASS] != s->bass_db_old || c[C_BFREQ] != s->bass_freq_old) {
        s->bass_db_old = c[C_BASS]; s->bass_freq_old = c[C_BFREQ];
        oracle_biquad_design(&s->bass, BQ_LOWSHELF, c[C_BFREQ] / sr, 0.707f, c[C_BASS]);
    }
    if (c[C_MID] != s->mid_db_old || c[C_MFREQ] != s->mid_freq_old || c[C_MIDQ] != s->mid_q_old || c[C_MTYPE] != s->mid_type_old) {
        s->mid_db_old = c[C_MID]; s->mid_freq_old = c[C_MFREQ]; s->mid_q_old = c[C_MIDQ]; s->mid_type_old = c[C_MTYPE];
        oracle_biquad_design(&s->mid, c[C_MTYPE] == 1.0f ? BQ_BANDPASS : BQ_PEAK, c[C_MFREQ] / sr, c[C_MIDQ], c[C_MID]);
    }
    if (c[C_TREBLE] != s->treble_db_old || c[C_TFREQ] != s->treble_freq_old) {
        s->treble_db_old = c[C_TREBLE]; s->treble_freq_old = c[C_TFREQ];
        oracle_biquad_design(&s->treble, BQ_HIGHSHELF, c[C_TFREQ] / sr, 0.707f, c[C_TREBLE]);
    }
    if (c[C_DEPTH] != s->depth_db_old) {
        s->depth_db_old = c[C_DEPTH];
        oracle_biquad_design(&s->depth, BQ_PEAK, 75.0f / sr, 0.707f, c[C_DEPTH]);
    }
    if (c[C_PRESENCE] != s->presence_db_old) {
        s->presence_db_old = c[C_PRESENCE];
        oracle_biquad_design(&s->presence, BQ_HIGHSHELF, 900.0f / sr, 0.707f, c[C_PRESENCE]);
    }
    if (c[C_MTYPE] == 1.0f) {
        apply_biquad(buf, buf, &s->mid, n);
    } else {
        apply_biquad(buf, buf, &s->depth, n);
        apply_biquad(buf, buf, &s->bass, n);
        apply_biquad(buf, buf, &s->mid, n);
        apply_biquad(buf, buf, &s->treble, n);
        apply_biquad(buf, buf, &s->presence, n);
    }
}

/* RtNeuralGeneric::run, rt-neural-generic.cpp:484-660 (audio part; atoms are control plane).
 * c: the 21 control-port values in port order starting at port 4.  Returns ModelInSize. */
int oracle_stream_run(OracleStream* s, const float* c, const float* in, float* out, uint32_t n)
{
    const float pregain = db_co(c[C_PREGAIN]);
    const float master = db_co(c[C_MASTER]);
    const int net_bypass = c[C_NETBYPASS] > 0.5f;
    const float in_lpf_pc = c[C_ANTIALIASING];
    const float eq_position = c[C_EQPOS], eq_bypass = c[C_EQBYPASS];
    const int enabled = c[C_ENABLED] > 0.5f;

    s->pre_gain.target = pregain;
    if (in_lpf_pc != s->in_lpf_pc_old) {
        oracle_biquad_design(&s->in_lpf, BQ_LOWPASS, lpf_fc(in_lpf_pc), 0.707f, 0.0f);
        s->in_lpf_pc_old = in_lpf_pc;
    }
    if (n == 0) return s->last_input_size;
    if (!enabled) {
        if (out != in) memcpy(out, in, sizeof(float) * n);
        return s->last_input_size;
    }
    if (in_lpf_pc != 0.0f) apply_biquad(out, in, &s->in_lpf, n);
    else if (out != in) memcpy(out, in, sizeof(float) * n);
    for (uint32_t i = 0; i < n; ++i) out[i] = out[i] * exp_next(&s->pre_gain);          /* applyGainRamp :34-38 */
    if (eq_position == 1.0f && eq_bypass == 0.0f) apply_tone_controls(s, c, out, n);
    if (s->model && !net_bypass) {
        lin_set_target(&s->p1, c[C_PARAM1]);
        lin_set_target(&s->p2, c[C_PARAM2]);
        if (s->param_first_run) {
            s->param_first_run = 0;
            s->p1.mem = s->p1.target;
            s->p2.mem = s->p2.target;
        }
        apply_model(s, out, n);
    }
    if (c[C_DCBLOCKER] == 1.0f) apply_biquad(out, out, &s->dc_blocker, n);
    if (eq_position == 0.0f && eq_bypass == 0.0f) apply_tone_controls(s, c, out, n);
    s->master_gain.target = s->loading ? 0.f : master;
    for (uint32_t i = 0; i < n; ++i) out[i] = out[i] * exp_next(&s->master_gain);
    return s->last_input_size;
}

void oracle_stream_get_state(const OracleStream* s, float* h, float* c, double* z14)
{
    if (s->model) oracle_model_get_state(s->model, h, c);
    const OracleBiquad* f[7] = {&s->in_lpf, &s->depth, &s->bass, &s->mid, &s->treble, &s->presence, &s->dc_blocker};
    for (int i = 0; i < 7; ++i) { z14[2 * i] = f[i]->z1; z14[2 * i + 1] = f[i]->z2; }
}
