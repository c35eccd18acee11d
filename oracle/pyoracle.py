"""ctypes wrapper of oracle/aidax_oracle.c (TEST INFRASTRUCTURE -- the checker, never the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "_build/liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(HERE, "_build", "liboracle.so")


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "_build", "liboracle.so")
        src = os.path.join(HERE, "aidax_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        fp, vp, dp = C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_double)
        L.oracle_biquad_coefs.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, dp]
        L.oracle_model_create.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, C.c_float, C.c_int]
        L.oracle_model_create.restype = vp
        L.oracle_model_destroy.argtypes = [vp]
        L.oracle_model_reset.argtypes = [vp]
        L.oracle_model_forward.argtypes = [vp, fp, fp, C.c_int]
        L.oracle_model_get_state.argtypes = [vp, fp, fp]
        L.oracle_stream_create.argtypes = [C.c_double]
        L.oracle_stream_create.restype = vp
        L.oracle_stream_destroy.argtypes = [vp]
        L.oracle_stream_set_model.argtypes = [vp, vp, C.c_int, C.c_float, C.c_float, C.c_float]
        L.oracle_stream_set_loading.argtypes = [vp, C.c_int]
        L.oracle_stream_get_params.argtypes = [vp, fp, fp]
        L.oracle_stream_activate.argtypes = [vp]
        L.oracle_stream_run.argtypes = [vp, fp, fp, fp, C.c_uint32]
        L.oracle_stream_run.restype = C.c_int
        L.oracle_stream_get_state.argtypes = [vp, fp, fp, dp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def biquad_coefs(type_, fc, q, gain_db):
    out = np.zeros(5, dtype=np.float64)
    lib().oracle_biquad_coefs(type_, fc, q, gain_db, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _create_model(m, use_double):
    W = np.ascontiguousarray(m.W, dtype=np.float32)
    U = np.ascontiguousarray(m.U, dtype=np.float32)
    b = np.ascontiguousarray(m.b, dtype=np.float32)
    dw = np.ascontiguousarray(m.dense_w, dtype=np.float32)
    h = lib().oracle_model_create(m.cell, m.hidden, m.input_size, _fp(W), _fp(U), _fp(b), _fp(dw), float(m.dense_b), int(use_double))
    if not h:
        raise ValueError("oracle_model_create rejected the model")
    return h


class OracleModel:
    """Raw ModelT::forward restatement (no gains, no skip)."""

    def __init__(self, m, use_double=False):
        self.m = m
        self.h = _create_model(m, use_double)
        lib().oracle_model_reset(self.h)

    def reset(self):
        lib().oracle_model_reset(self.h)

    def forward(self, x):
        """x: [n] (input_size 1) or [n][input_size]."""
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.m.input_size)
        out = np.empty(x.shape[0], dtype=np.float32)
        lib().oracle_model_forward(self.h, _fp(x), _fp(out), x.shape[0])
        return out

    def state(self):
        h = np.zeros(self.m.hidden, dtype=np.float32)
        c = np.zeros(self.m.hidden, dtype=np.float32)
        lib().oracle_model_get_state(self.h, _fp(h), _fp(c))
        return h, c

    def __del__(self):
        try:
            lib().oracle_model_destroy(self.h)
        except Exception:
            pass


class OracleStream:
    """One plugin instance worth of run() (rt-neural-generic.cpp:484-660)."""

    def __init__(self, samplerate=48000.0):
        self.h = lib().oracle_stream_create(float(samplerate))
        self.model = None

    def set_model(self, m, use_double=False):
        """Mirrors work()+loadModelFromPath+work_response: includes the 2048-sample pre-warm; clears `loading`."""
        self.model = m
        if m is None:
            lib().oracle_stream_set_model(self.h, None, 0, 1.0, 1.0, 48000.0)
            return
        mh = _create_model(m, use_double)  # ownership passes to the stream
        lib().oracle_stream_set_model(self.h, mh, int(m.in_skip), float(m.in_gain), float(m.out_gain), float(m.samplerate))

    def set_loading(self, v):
        lib().oracle_stream_set_loading(self.h, int(bool(v)))

    def activate(self):
        lib().oracle_stream_activate(self.h)

    def params(self):
        a, b = C.c_float(0), C.c_float(0)
        lib().oracle_stream_get_params(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def run(self, controls, x):
        c = np.ascontiguousarray(controls, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        lib().oracle_stream_run(self.h, _fp(c), _fp(x), _fp(out), x.size)
        return out

    def state(self):
        H = self.model.hidden if self.model is not None else 1
        h = np.zeros(max(H, 1), dtype=np.float32)
        c = np.zeros(max(H, 1), dtype=np.float32)
        z = np.zeros(14, dtype=np.float64)
        lib().oracle_stream_get_state(self.h, _fp(h), _fp(c), z.ctypes.data_as(C.POINTER(C.c_double)))
        return h, c, z

    def __del__(self):
        try:
            lib().oracle_stream_destroy(self.h)
        except Exception:
            pass
