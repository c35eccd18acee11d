"""In-tree build of the native libraries (make + nvcc for sm_100a; works without a GPU)."""
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# AIDAX_LIB_DIR: measurement runs point the bindings at a library built with other compile-time choices (tools/variant.sh)
LIB_DIR = os.environ.get("AIDAX_LIB_DIR") or os.path.join(PKG_DIR, "lib")


def lib_path(name):
    return os.path.join(LIB_DIR, name)


def build(jobs=None, verbose=False):
    """Compile every CUDA/C++ target for sm_100a.  Idempotent (make)."""
    jobs = jobs or max(2, (os.cpu_count() or 4))
    cmd = ["make", "-C", PKG_DIR, "-j%d" % jobs, "all"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("native build failed (see output above)")
    return LIB_DIR
