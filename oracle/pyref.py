"""ctypes wrappers around the UNMODIFIED reference compiled into oracle/_ref by oracle/build_ref.sh
(TEST INFRASTRUCTURE).  Tier B/C: model-only harness (libref_model*.so).  Tier A (the full plugin
through its LV2 surface) is driven by the repo's own batch-render host, see tests/lv2host.py."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available(name="libref_model.so"):
    return os.path.exists(os.path.join(REF_DIR, name))


class RefModel:
    """ModelT::forward of the reference (xsimd SSE2 by default; flavour='stl' for the scalar backend)."""

    def __init__(self, json_path, flavour="xsimd"):
        name = "libref_model.so" if flavour == "xsimd" else "libref_model_stl.so"
        self.lib = C.CDLL(os.path.join(REF_DIR, name))
        self.lib.ref_model_create.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        self.lib.ref_model_create.restype = C.c_void_p
        self.lib.ref_model_forward.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
        self.lib.ref_model_reset.argtypes = [C.c_void_p]
        self.lib.ref_model_destroy.argtypes = [C.c_void_p]
        self.lib.ref_model_flavour.restype = C.c_char_p
        n = C.c_int(0)
        self.h = self.lib.ref_model_create(os.fsencode(json_path), C.byref(n))
        if not self.h:
            raise ValueError("reference refused the model file %s" % json_path)
        self.input_size = n.value

    @property
    def flavour(self):
        return self.lib.ref_model_flavour().decode()

    def reset(self):
        self.lib.ref_model_reset(self.h)

    def forward(self, x):
        """x: [n] or [n][input_size]; padded to 4 lanes with zeros (SURVEY.md hazard H1)."""
        x = np.asarray(x, dtype=np.float32).reshape(-1, self.input_size)
        xp = np.zeros((x.shape[0], 4), dtype=np.float32)
        xp[:, : self.input_size] = x
        out = np.empty(x.shape[0], dtype=np.float32)
        self.lib.ref_model_forward(self.h, xp.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)), x.shape[0])
        return out

    def __del__(self):
        try:
            self.lib.ref_model_destroy(self.h)
        except Exception:
            pass
