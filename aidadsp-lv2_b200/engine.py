"""ctypes binding of include/aidax.h.  Thin: every method is one C-ABI call."""
import ctypes as C
import os

import numpy as np

from .build import lib_path

N_CONTROLS = 21
LAYOUT_PLANAR = 0
LAYOUT_INTERLEAVED = 1

# control indices = LV2 port index - 4 (include/aidax.h, reference ports_t rt-neural-generic.h:84-112)
CTRL = dict(ANTIALIASING=0, PREGAIN=1, NETBYPASS=2, PARAM1=3, PARAM2=4, EQBYPASS=5, EQPOS=6, BASS=7, BFREQ=8,
            MID=9, MFREQ=10, MIDQ=11, MTYPE=12, TREBLE=13, TFREQ=14, DEPTH=15, PRESENCE=16, DCBLOCKER=17,
            MASTER=18, MODELINSIZE=19, ENABLED=20)


def default_controls(**overrides):
    """ttl defaults (ttl/rt-neural-generic.ttl:92-313) with overrides by symbol name."""
    c = np.zeros(N_CONTROLS, dtype=np.float32)
    c[CTRL["ANTIALIASING"]] = 66.216
    c[CTRL["BFREQ"]] = 305.0
    c[CTRL["MFREQ"]] = 750.0
    c[CTRL["MIDQ"]] = 0.707
    c[CTRL["TFREQ"]] = 2000.0
    c[CTRL["DCBLOCKER"]] = 1.0
    c[CTRL["ENABLED"]] = 1.0
    for k, v in overrides.items():
        c[CTRL[k]] = v
    return c


class AidaxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("aidax error %d: %s" % (code, msg))
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [("cell", C.c_int32), ("hidden", C.c_int32), ("input_size", C.c_int32), ("in_skip", C.c_int32),
                ("in_gain", C.c_float), ("out_gain", C.c_float), ("samplerate", C.c_float)]


_lib = None


def load_library():
    """Load lib/libaidax.so.  Fails loudly if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path("libaidax.so")
    if not os.path.exists(path):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first" % path)
    lib = C.CDLL(path)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
    lib.aidax_last_error.restype = C.c_char_p
    lib.aidax_engine_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.aidax_engine_destroy.argtypes = [vp]
    lib.aidax_engine_cuda_stream.argtypes = [vp]
    lib.aidax_engine_cuda_stream.restype = vp
    lib.aidax_engine_configure.argtypes = [vp, C.c_int, C.c_int]
    lib.aidax_model_register.argtypes = [vp, C.POINTER(ModelDesc), fp, fp, fp, fp, C.c_float, ip]
    lib.aidax_model_load_json.argtypes = [vp, C.c_char_p, ip, ip]
    lib.aidax_model_describe.argtypes = [vp, C.c_int, C.POINTER(ModelDesc)]
    lib.aidax_model_release.argtypes = [vp, C.c_int]
    lib.aidax_model_selftest.argtypes = [vp, C.c_int, fp, ip, ip]
    lib.aidax_engine_pci_bus_id.argtypes = [vp, C.c_char_p, C.c_int]
    lib.aidax_engine_describe_plan.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.aidax_stream_open.argtypes = [vp, C.c_double, ip]
    lib.aidax_stream_close.argtypes = [vp, C.c_int]
    lib.aidax_stream_activate.argtypes = [vp, C.c_int]
    lib.aidax_stream_set_model.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float]
    lib.aidax_stream_set_controls.argtypes = [vp, C.c_int, fp, C.c_int]
    lib.aidax_streams_set_params.argtypes = [vp, C.c_int, C.c_int, fp, fp]
    lib.aidax_stream_get_state.argtypes = [vp, C.c_int, fp, fp, C.POINTER(C.c_double)]
    lib.aidax_stream_set_state.argtypes = [vp, C.c_int, fp, fp, C.POINTER(C.c_double)]
    lib.aidax_render_host.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_size_t]
    lib.aidax_render_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_size_t]
    lib.aidax_render_host_submit.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_size_t]
    lib.aidax_render_host_wait.argtypes = [vp]
    lib.aidax_engine_synchronize.argtypes = [vp]
    lib.aidax_host_alloc.argtypes = [C.c_size_t]
    lib.aidax_host_alloc.restype = vp
    lib.aidax_host_free.argtypes = [vp]
    lib.aidax_engine_launch_count.argtypes = [vp]
    lib.aidax_engine_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class PinnedArray:
    """float32 numpy view over cudaHostAlloc'ed memory (aidax_host_alloc)."""

    def __init__(self, shape):
        self.lib = load_library()
        n = int(np.prod(shape))
        self.ptr = self.lib.aidax_host_alloc(n * 4)
        if not self.ptr:
            raise AidaxError(2, self.lib.aidax_last_error().decode())
        buf = (C.c_float * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.aidax_host_free(self.ptr)
            self.ptr = None


class Engine:
    """One engine == one GPU.  Streams are dense integer ids in open order."""

    def __init__(self, device=0, max_streams=1024, max_block=128):
        self.lib = load_library()
        self.h = C.c_void_p()
        self._chk(self.lib.aidax_engine_create(device, max_streams, max_block, C.byref(self.h)))
        self.max_streams, self.max_block = max_streams, max_block

    def _chk(self, rc):
        if rc != 0:
            raise AidaxError(rc, self.lib.aidax_last_error().decode())

    def close(self):
        if self.h:
            self.lib.aidax_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def configure(self, tile_mode=0, math_mode=0):
        self._chk(self.lib.aidax_engine_configure(self.h, tile_mode, math_mode))

    @property
    def cuda_stream(self):
        return self.lib.aidax_engine_cuda_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.aidax_engine_launch_count(self.h))

    # -- models
    def register_model(self, m):
        """m: models.Model (reference JSON layout arrays)."""
        d = ModelDesc(m.cell, m.hidden, m.input_size, int(m.in_skip), float(m.in_gain), float(m.out_gain), float(m.samplerate))
        W = np.ascontiguousarray(m.W, dtype=np.float32)
        U = np.ascontiguousarray(m.U, dtype=np.float32)
        b = np.ascontiguousarray(m.b, dtype=np.float32)
        dw = np.ascontiguousarray(m.dense_w, dtype=np.float32).reshape(-1)
        mid = C.c_int(-1)
        self._chk(self.lib.aidax_model_register(self.h, C.byref(d), _fptr(W), _fptr(U), _fptr(b), _fptr(dw), float(m.dense_b), C.byref(mid)))
        return mid.value

    def load_json(self, path):
        mid, insz = C.c_int(-1), C.c_int(0)
        self._chk(self.lib.aidax_model_load_json(self.h, os.fsencode(path), C.byref(mid), C.byref(insz)))
        return mid.value, insz.value

    def release_model(self, mid):
        """Give back an id from load_json / register_model; the weights go when no stream is bound to it any more."""
        self._chk(self.lib.aidax_model_release(self.h, mid))

    @property
    def pci_bus_id(self):
        buf = C.create_string_buffer(64)
        self._chk(self.lib.aidax_engine_pci_bus_id(self.h, buf, 64))
        return buf.value.decode()

    def describe_plan(self, first=0, n=None):
        """The planner's launches for a stream range: list of dicts (cell, H, shape, streams, ctas, streams_per_cta, threads)."""
        buf = C.create_string_buffer(1 << 14)
        self._chk(self.lib.aidax_engine_describe_plan(self.h, first, self.max_streams if n is None else n, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            f = line.split()
            d = {"cell": f[0]}
            for kv in f[1:]:
                k, v = kv.split("=")
                d[k] = v if k == "shape" else int(v)
            out.append(d)
        return out

    def selftest_model(self, mid):
        """The reference's load-time testModel on the GPU: (max abs error, samples above 1e-5, samples compared)."""
        err, bad, n = C.c_float(0), C.c_int(0), C.c_int(0)
        self._chk(self.lib.aidax_model_selftest(self.h, mid, C.byref(err), C.byref(bad), C.byref(n)))
        return err.value, bad.value, n.value

    def describe_model(self, mid):
        d = ModelDesc()
        self._chk(self.lib.aidax_model_describe(self.h, mid, C.byref(d)))
        return d

    # -- streams
    def open_stream(self, samplerate=48000.0):
        sid = C.c_int(-1)
        self._chk(self.lib.aidax_stream_open(self.h, float(samplerate), C.byref(sid)))
        return sid.value

    def close_stream(self, s):
        self._chk(self.lib.aidax_stream_close(self.h, s))

    def activate(self, s):
        self._chk(self.lib.aidax_stream_activate(self.h, s))

    def set_model(self, s, model_id, old_param1=0.0, old_param2=0.0):
        self._chk(self.lib.aidax_stream_set_model(self.h, s, model_id, old_param1, old_param2))

    def set_controls(self, s, controls, loading=False):
        c = np.ascontiguousarray(controls, dtype=np.float32)
        assert c.size == N_CONTROLS
        self._chk(self.lib.aidax_stream_set_controls(self.h, s, _fptr(c), int(bool(loading))))

    def set_params(self, first, p1=None, p2=None):
        n = len(p1) if p1 is not None else len(p2)
        a1 = np.ascontiguousarray(p1, dtype=np.float32) if p1 is not None else None
        a2 = np.ascontiguousarray(p2, dtype=np.float32) if p2 is not None else None
        self._chk(self.lib.aidax_streams_set_params(self.h, first, n, _fptr(a1) if a1 is not None else None,
                                                    _fptr(a2) if a2 is not None else None))

    def get_state(self, s, hidden):
        h = np.zeros(hidden, dtype=np.float32)
        c = np.zeros(hidden, dtype=np.float32)
        z = np.zeros(14, dtype=np.float64)
        self._chk(self.lib.aidax_stream_get_state(self.h, s, _fptr(h), _fptr(c), z.ctypes.data_as(C.POINTER(C.c_double))))
        return h, c, z

    def set_state(self, s, h=None, c=None, z=None):
        h = np.ascontiguousarray(h, dtype=np.float32) if h is not None else None
        c = np.ascontiguousarray(c, dtype=np.float32) if c is not None else None
        z = np.ascontiguousarray(z, dtype=np.float64) if z is not None else None
        self._chk(self.lib.aidax_stream_set_state(self.h, s, _fptr(h) if h is not None else None, _fptr(c) if c is not None else None,
                                                  z.ctypes.data_as(C.POINTER(C.c_double)) if z is not None else None))

    # -- render
    def render_host(self, x, out=None, first=0, layout=LAYOUT_PLANAR):
        """x: float32 [n_streams][n_samples] (planar) or [n_samples][n_streams] (interleaved) host array."""
        x = np.ascontiguousarray(x, dtype=np.float32) if not (isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags.c_contiguous) else x
        if out is None:
            out = np.empty_like(x)
        n_streams, n_samples = (x.shape if layout == LAYOUT_PLANAR else x.shape[::-1])
        self._chk(self.lib.aidax_render_host(self.h, first, n_streams, n_samples, x.ctypes.data, out.ctypes.data, layout, x.shape[1]))
        return out

    def render_host_submit(self, x, out, first=0):
        """Asynchronous half of render_host for a planar [n_streams][n_samples] block (x/out must outlive render_host_wait)."""
        assert x.dtype == np.float32 and x.flags.c_contiguous and out.dtype == np.float32 and out.flags.c_contiguous
        self._chk(self.lib.aidax_render_host_submit(self.h, first, x.shape[0], x.shape[1], x.ctypes.data, out.ctypes.data, LAYOUT_PLANAR, x.shape[1]))

    def render_host_wait(self):
        self._chk(self.lib.aidax_render_host_wait(self.h))

    def render_device(self, d_in, d_out, n_streams, n_samples, first=0, layout=LAYOUT_INTERLEAVED, stride=None):
        """d_in/d_out: device pointers (ints), e.g. torch_tensor.data_ptr().  Asynchronous."""
        if stride is None:
            stride = n_streams if layout == LAYOUT_INTERLEAVED else n_samples
        self._chk(self.lib.aidax_render_device(self.h, first, n_streams, n_samples, d_in, d_out, layout, stride))

    def synchronize(self):
        self._chk(self.lib.aidax_engine_synchronize(self.h))

    def generate_input(self, d_out, n_streams, n_samples, start=0, first=0, layout=LAYOUT_INTERLEAVED, stride=None, seed=0, device=0):
        """Device-side synthetic input (bench support, kernel K6) into the device pointer d_out, on the engine's stream."""
        if stride is None:
            stride = n_streams if layout == LAYOUT_INTERLEAVED else n_samples
        self.lib.aidax_generate_input.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_size_t, C.c_uint32]
        self._chk(self.lib.aidax_generate_input(device, self.cuda_stream, d_out, first, n_streams, start, n_samples, layout, stride, seed))


def measure_pcie(device=0, nbytes=256 << 20):
    """Pinned-memory copy rates in GB/s: (H2D alone, D2H alone, per direction with both at once)."""
    lib = load_library()
    lib.aidax_measure_pcie.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_double)]
    v = (C.c_double * 3)()
    rc = lib.aidax_measure_pcie(device, nbytes, v)
    if rc != 0:
        raise RuntimeError("aidax_measure_pcie: %s" % lib.aidax_last_error().decode())
    return float(v[0]), float(v[1]), float(v[2])


def measure_fp32_peak(device=0, millis=200):
    """Sustained FFMA throughput of the device in TFLOP/s (roofline denominator of the recurrence kernels)."""
    lib = load_library()
    lib.aidax_measure_fp32_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    v = C.c_double(0.0)
    rc = lib.aidax_measure_fp32_peak(device, millis, C.byref(v))
    if rc != 0:
        raise AidaxError(rc, lib.aidax_last_error().decode())
    return v.value


def measure_pipe_peaks(device=0, millis=700):
    """Measured pipe peaks (csrc/probes.cu): dict with fp32_tflops, mufu_gops, fp64_tflops, f2f_gops,
    tf32_tflops_n256 / _n160 / _n64 (tcgen05 kind::tf32, M=128) and the SM clock the probes ran at."""
    lib = load_library()
    lib.aidax_measure_pipe_peaks.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    v = (C.c_double * 8)()
    rc = lib.aidax_measure_pipe_peaks(device, millis, v)
    if rc != 0:
        raise AidaxError(rc, lib.aidax_last_error().decode())
    keys = ("fp32_tflops", "mufu_gops", "fp64_tflops", "f2f_gops", "tf32_tflops_n256", "tf32_tflops_n160", "tf32_tflops_n64", "sm_mhz")
    return {k: float(v[i]) for i, k in enumerate(keys)}


# ---- host-only helpers (no GPU needed)
def check_json(path):
    lib = load_library()
    lib.aidax_model_check_json.argtypes = [C.c_char_p, C.POINTER(ModelDesc)]
    d = ModelDesc()
    rc = lib.aidax_model_check_json(os.fsencode(path), C.byref(d))
    if rc != 0:
        raise AidaxError(rc, lib.aidax_last_error().decode())
    return d


def design_biquad(type_, fc, q, gain_db):
    lib = load_library()
    lib.aidax_design_biquad.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
    out = np.zeros(5, dtype=np.float64)
    rc = lib.aidax_design_biquad(type_, fc, q, gain_db, out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise AidaxError(rc, lib.aidax_last_error().decode())
    return out


def decode_controls(controls, samplerate=48000.0, loading=False, model_bound=True):
    lib = load_library()
    lib.aidax_decode_controls.argtypes = [C.c_double, C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    c = np.ascontiguousarray(controls, dtype=np.float32)
    flags, pre, master = C.c_uint32(0), C.c_float(0), C.c_float(0)
    coefs = np.zeros(35, dtype=np.float64)
    rc = lib.aidax_decode_controls(samplerate, _fptr(c), int(loading), int(model_bound), C.byref(flags), C.byref(pre), C.byref(master),
                                   coefs.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise AidaxError(rc, lib.aidax_last_error().decode())
    return flags.value, pre.value, master.value, coefs.reshape(7, 5)
