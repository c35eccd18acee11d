"""ctypes binding of include/aidax_host.h: the batch-render LV2 host.  Works with any
rt-neural-generic binary (this repo's CUDA plugin or the unmodified reference build)."""
import ctypes as C
import os

import numpy as np

from .build import lib_path

PORT = dict(IN=0, OUT=1, CONTROL=2, NOTIFY=3, ANTIALIASING=4, PREGAIN=5, NETBYPASS=6, PARAM1=7, PARAM2=8, EQBYPASS=9,
            EQPOS=10, BASS=11, BFREQ=12, MID=13, MFREQ=14, MIDQ=15, MTYPE=16, TREBLE=17, TFREQ=18, DEPTH=19, PRESENCE=20,
            DCBLOCKER=21, MASTER=22, MODELINSIZE=23, ENABLED=24)

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = lib_path("libaidax_host.so")
        if not os.path.exists(path):
            raise ImportError("%s is missing: build first (__graft_entry__.build())" % path)
        L = C.CDLL(path)
        vp = C.c_void_p
        L.aidax_host_error.restype = C.c_char_p
        L.aidax_host_open.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
        L.aidax_host_open.restype = vp
        L.aidax_host_close.argtypes = [vp]
        L.aidax_host_is_batched.argtypes = [vp]
        L.aidax_host_launch_count.argtypes = [vp]
        L.aidax_host_launch_count.restype = C.c_uint64
        L.aidax_host_timing.argtypes = [vp, C.POINTER(C.c_double)]
        L.aidax_host_timing.restype = None
        L.aidax_host_input.argtypes = [vp]
        L.aidax_host_input.restype = vp
        L.aidax_host_output.argtypes = [vp]
        L.aidax_host_output.restype = vp
        L.aidax_host_stride.argtypes = [vp]
        L.aidax_host_stride.restype = C.c_size_t
        L.aidax_host_set_control.argtypes = [vp, C.c_int, C.c_int, C.c_float]
        L.aidax_host_get_control.argtypes = [vp, C.c_int, C.c_int]
        L.aidax_host_get_control.restype = C.c_float
        L.aidax_host_load_model.argtypes = [vp, C.c_int, C.c_char_p]
        L.aidax_host_activate.argtypes = [vp]
        L.aidax_host_deactivate.argtypes = [vp]
        L.aidax_host_run.argtypes = [vp, C.c_int]
        L.aidax_host_notify_count.argtypes = [vp, C.c_int]
        L.aidax_host_last_notify_path.argtypes = [vp, C.c_int]
        L.aidax_host_last_notify_path.restype = C.c_char_p
        L.aidax_host_save_state.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
        L.aidax_host_restore_state.argtypes = [vp, C.c_int, C.c_char_p]
        L.aidax_host_set_active.argtypes = [vp, C.c_int]
        L.aidax_host_is_bulk.argtypes = [vp]
        L.aidax_host_cpus.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
        _lib = L
    return _lib


def plugin_path():
    """The CUDA-backed rt-neural-generic.so of this repository."""
    return lib_path("rt-neural-generic.so")


class BatchHost:
    def __init__(self, plugin_so=None, n_instances=1, max_block=128, samplerate=48000.0, n_threads=1, device=-1):
        self.lib = _load()
        plugin_so = plugin_so or plugin_path()
        self.h = self.lib.aidax_host_open(os.fsencode(plugin_so), float(samplerate), n_instances, max_block, n_threads, device)
        if not self.h:
            raise RuntimeError("aidax_host_open failed: %s" % self.lib.aidax_host_error().decode())
        self.n, self.max_block = n_instances, max_block
        self.stride = int(self.lib.aidax_host_stride(self.h))
        n_floats = self.n * self.stride
        self.input = np.frombuffer((C.c_float * n_floats).from_address(self.lib.aidax_host_input(self.h)), dtype=np.float32).reshape(self.n, self.stride)
        self.output = np.frombuffer((C.c_float * n_floats).from_address(self.lib.aidax_host_output(self.h)), dtype=np.float32).reshape(self.n, self.stride)

    @property
    def batched(self):
        return bool(self.lib.aidax_host_is_batched(self.h))

    @property
    def bulk(self):
        """True if periods go through the plugin's run_block entry (include/aidax_batch.h)."""
        return bool(self.lib.aidax_host_is_bulk(self.h))

    @property
    def cpus(self):
        """cpus the period threads are pinned to ([] = not pinned)."""
        buf = (C.c_int * 1024)()
        n = self.lib.aidax_host_cpus(self.h, buf, 1024)
        return [int(buf[i]) for i in range(min(n, 1024))]

    def set_active(self, n_active):
        """Only instances [0, n_active) take part in the following periods."""
        if self.lib.aidax_host_set_active(self.h, int(n_active)):
            raise ValueError(self.lib.aidax_host_error().decode())

    @property
    def launch_count(self):
        return int(self.lib.aidax_host_launch_count(self.h))

    def set_control(self, port, value, instance=-1):
        port = PORT[port] if isinstance(port, str) else port
        if self.lib.aidax_host_set_control(self.h, instance, port, float(value)):
            raise ValueError(self.lib.aidax_host_error().decode())

    def get_control(self, port, instance=0):
        port = PORT[port] if isinstance(port, str) else port
        return float(self.lib.aidax_host_get_control(self.h, instance, port))

    def load_model(self, path, instance=-1):
        if self.lib.aidax_host_load_model(self.h, instance, os.fsencode(path)):
            raise ValueError(self.lib.aidax_host_error().decode())

    def activate(self):
        self.lib.aidax_host_activate(self.h)

    def run(self, n_samples):
        if self.lib.aidax_host_run(self.h, n_samples):
            raise RuntimeError(self.lib.aidax_host_error().decode())

    def process(self, x):
        """x: [n_instances][n_samples]; returns a copy of the outputs."""
        n = x.shape[1]
        self.input[:, :n] = x
        self.run(n)
        return self.output[:, :n].copy()

    def timing(self):
        """(seconds in the N run() calls, seconds in batch end(), periods) since open()."""
        out = (C.c_double * 3)()
        self.lib.aidax_host_timing(self.h, out)
        return float(out[0]), float(out[1]), float(out[2])

    def notify_count(self, instance=0):
        return int(self.lib.aidax_host_notify_count(self.h, instance))

    def last_notify_path(self, instance=0):
        return self.lib.aidax_host_last_notify_path(self.h, instance).decode()

    def save_state(self, instance=0):
        buf = C.create_string_buffer(2048)
        rc = self.lib.aidax_host_save_state(self.h, instance, buf, 2048)
        return rc, buf.value.decode()

    def restore_state(self, path, instance=0):
        return self.lib.aidax_host_restore_state(self.h, instance, os.fsencode(path))

    def close(self):
        if self.h:
            self.input = self.output = None
            self.lib.aidax_host_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
