"""The algebra of the tensor-core kernel's gate epilogue (kernels_tc.cuh, eval_group with AIDAX_SCALED_GATES), restated in
numpy float32 and checked against the LSTM / GRU cell in float64 -- including saturated gates, where the batched
reciprocal would overflow without the 2^-24 scaling + saturation.  No GPU: this pins the formulas, the GPU tests pin
the kernel."""
import numpy as np

f32 = np.float32
S = f32(2.0 ** -24)
L2E = f32(1.4426950408889634)


def _sat(x):
    return np.clip(x, f32(0), f32(1)).astype(np.float32)


def _scaled_one_plus_exp2(e):
    """2^-24 * (1 + 2^e) saturated at 1: FFMA.SAT(2^e, s, s); 2^e = +inf saturates as well."""
    with np.errstate(over="ignore"):
        p = np.exp2(e.astype(np.float64)).astype(np.float32)
    return _sat(p * S + S)


def _one_plus_exp2(e):
    return (np.exp2(np.minimum(e, f32(24))).astype(np.float32) + f32(1)).astype(np.float32)


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x.astype(np.float64)))


def _lstm_scaled(ai, af, ag, ao, c0):
    """accumulators already scaled as the B columns are: i,f,o by -log2 e, g by 2 log2 e"""
    A = _scaled_one_plus_exp2(ai); B = _one_plus_exp2(af); G = _scaled_one_plus_exp2(ag); D = _scaled_one_plus_exp2(ao)
    ag_, bd = (A * G).astype(f32), (B * D).astype(f32)
    r = (f32(1) / (ag_ * bd)).astype(f32)
    rag, rbd = (r * bd).astype(f32), (r * ag_).astype(f32)
    itg = (rag * (S * G + f32(-2) * S * S)).astype(f32)       # i * tanh(g) = (g - 2) / (a g)
    fi, ogx = (rbd * D).astype(f32), (rbd * B).astype(f32)     # f = 1/b ; 1/(s d)
    c = (fi * c0 + itg).astype(f32)
    ec = _scaled_one_plus_exp2((c * (f32(2) * L2E)).astype(f32))
    t = (f32(1) / ec).astype(f32)
    h = (ogx * (f32(-2) * S * S * t + S)).astype(f32)          # o * tanh(c')
    return c, h


def test_scaled_gate_lstm_matches_the_cell_including_saturation():
    rng = np.random.default_rng(0)
    for scale in (1.0, 5.0, 20.0, 80.0):
        i, f, g, o = [(rng.standard_normal(200000) * scale).astype(f32) for _ in range(4)]
        c0 = (rng.standard_normal(200000) * 2).astype(f32)
        ct = _sig(i) * np.tanh(g.astype(np.float64)) + _sig(f) * c0
        ht = _sig(o) * np.tanh(ct)
        c, h = _lstm_scaled((i * -L2E).astype(f32), (f * -L2E).astype(f32), (g * (2 * L2E)).astype(f32), (o * -L2E).astype(f32), c0)
        assert np.isfinite(c).all() and np.isfinite(h).all(), scale
        # fp32 resolution of c (|c| up to ~10) and of h in [-1, 1]
        assert np.abs(c - ct).max() <= 4e-6, (scale, np.abs(c - ct).max())
        assert np.abs(h - ht).max() <= 2e-6, (scale, np.abs(h - ht).max())


def test_scaled_candidate_gru_matches_the_cell():
    rng = np.random.default_rng(1)
    for scale in (1.0, 10.0, 60.0):
        z, r, a2, a3 = [(rng.standard_normal(200000) * scale).astype(f32) for _ in range(4)]
        h0 = (rng.uniform(-1, 1, 200000)).astype(f32)
        zt, rt = _sig(z), _sig(r)
        ht = (1 - zt) * np.tanh(rt * a2.astype(np.float64) + a3) + zt * h0
        # kernel: z, r from one reciprocal of the product of the two unscaled terms; candidate term scaled by 2^-24
        Z, R = _one_plus_exp2((z * -L2E).astype(f32)), _one_plus_exp2((r * -L2E).astype(f32))
        rr = (f32(1) / (Z * R)).astype(f32)
        zg, rg = (rr * R).astype(f32), (rr * Z).astype(f32)
        # the accumulator columns of the candidate carry 2 log2 e (column_scale)
        e = _scaled_one_plus_exp2((rg * (a2 * (2 * L2E)).astype(f32) + (a3 * (2 * L2E)).astype(f32)).astype(f32))
        t = (f32(1) / e).astype(f32)
        hh = (f32(-2) * S * t + f32(1)).astype(f32)
        h = ((f32(1) - zg) * hh + zg * h0).astype(f32)
        assert np.isfinite(h).all()
        assert np.abs(h - ht).max() <= 2e-5 * max(1.0, scale / 10), (scale, np.abs(h - ht).max())


def test_batched_reciprocal_never_overflows():
    # four saturated terms: the scaled form multiplies to >= 2^-96 and <= 2^24 + 1; the unscaled clamped form to <= 2^97
    e = np.array([200.0, 200.0, 200.0, 200.0, -200.0, 0.0, 24.0, 23.999], dtype=f32)
    w = _scaled_one_plus_exp2(e)
    assert (w <= 1).all() and (w >= S).all()
    assert np.isfinite(f32(1) / (w[0] * w[1] * w[2] * w[3]))
    u = _one_plus_exp2(e)
    assert np.isfinite(f32(1) / (u[0] * u[1] * u[2] * u[3]))
