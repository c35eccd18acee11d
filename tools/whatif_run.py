"""Scratch: time the H=16 tensor-core kernel of every tools/bin/whatif_*/libaidax.so (see tools/whatif.sh)."""
import glob, os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, "tools"); sys.path.insert(0, ".")
    import aidadsp_lv2_b200.engine as e
    lib = sys.argv[1]
    if lib != "base":
        e.lib_path = lambda name: lib
    from quick_bench import run
    run(262144, 16, tile_mode=3)
    run(131072, 16, cell=1, input_size=3, tile_mode=3)
else:
    for lib in ["base"] + sorted(glob.glob("tools/bin/whatif_*/libaidax.so")):
        print("==", lib, flush=True)
        subprocess.run([sys.executable, "tools/whatif_run.py", lib])
