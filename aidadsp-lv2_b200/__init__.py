"""aidax-b200: B200-native batch engine for the rt-neural-generic hot path of AidaDSP/aidadsp-lv2.

The product is the C-ABI shared library ``lib/libaidax.so`` (CUDA kernels + engine), the LV2 plugin
``lib/rt-neural-generic.so`` built on it and the batch-render host ``lib/libaidax_host.so``.
This Python package is plumbing for tests and benchmarks: ctypes bindings, model-file utilities and
synthetic signals.  It never computes audio itself and has no CPU fallback.
"""
from .build import build, lib_path, LIB_DIR, PKG_DIR  # noqa: F401
from .engine import Engine, AidaxError, CTRL, N_CONTROLS, default_controls, LAYOUT_PLANAR, LAYOUT_INTERLEAVED  # noqa: F401
from . import models, signals, sharding  # noqa: F401
from .host import BatchHost, plugin_path, PORT  # noqa: F401
