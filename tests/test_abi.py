"""The C-ABI libraries load and export every symbol their headers declare (no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(aidax_[a-z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if not n.endswith("_t")))


@pytest.mark.parametrize("header,lib", [("aidax.h", "libaidax.so"), ("aidax_host.h", "libaidax_host.so")])
def test_library_exports_every_declared_symbol(native_libs, header, lib):
    L = C.CDLL(ax.lib_path(lib))
    names = _declared_functions(header)
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "%s does not export %s" % (lib, n)


def test_plugin_descriptor_surface(native_libs):
    """lv2_descriptor(0) has the reference URI and all seven callbacks; index 1 is NULL (rt-neural-generic.cpp:24-29)."""
    L = C.CDLL(ax.plugin_path())

    class Desc(C.Structure):
        _fields_ = [("URI", C.c_char_p)] + [(n, C.c_void_p) for n in
                    ("instantiate", "connect_port", "activate", "run", "deactivate", "cleanup", "extension_data")]
    L.lv2_descriptor.restype = C.POINTER(Desc)
    L.lv2_descriptor.argtypes = [C.c_uint32]
    d = L.lv2_descriptor(0)
    assert d.contents.URI == b"http://aidadsp.cc/plugins/aidadsp-bundle/rt-neural-generic"
    for n in ("instantiate", "connect_port", "activate", "run", "deactivate", "cleanup", "extension_data"):
        assert getattr(d.contents, n)
    assert not L.lv2_descriptor(1)
    ext = C.CFUNCTYPE(C.c_void_p, C.c_char_p)(d.contents.extension_data)
    assert ext(b"http://lv2plug.in/ns/ext/state#interface")
    assert ext(b"http://lv2plug.in/ns/ext/worker#interface")
    assert ext(b"http://aidadsp.cc/plugins/aidadsp-bundle/rt-neural-generic#batch")
    assert not ext(b"http://example.org/unknown")
    # instantiate without urid:map / work:schedule must return NULL (rt-neural-generic.cpp:265-273)
    inst = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_double, C.c_char_p, C.c_void_p)(d.contents.instantiate)
    empty = (C.c_void_p * 1)(None)
    assert not inst(None, 48000.0, b"", C.cast(empty, C.c_void_p))


def test_abi_version_and_error_string(native_libs):
    lib = engine.load_library()
    assert lib.aidax_abi_version() == 2
    h = C.c_void_p()
    rc = lib.aidax_engine_create(0, 0, 128, C.byref(h))  # invalid size: argument error before any CUDA call
    assert rc == 1 and b"bad" in lib.aidax_last_error()


def test_no_cpu_fallback_without_gpu(native_libs):
    """On a box without a GPU the engine refuses to exist (there is no CPU path to fall back to)."""
    lib = engine.load_library()
    h = C.c_void_p()
    rc = lib.aidax_engine_create(0, 64, 128, C.byref(h))
    if rc == 0:  # GPU box
        lib.aidax_engine_destroy(h)
    else:
        assert rc == 2 and b"no CPU fallback" in lib.aidax_last_error()
