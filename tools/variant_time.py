"""Scratch (round 2): time the headline shapes on the in-tree library and on every tools/bin/<name>/ variant library
(tools/variant.sh), one process per library (AIDAX_LIB_DIR).   python tools/variant_time.py [name ...]"""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, "tools"); sys.path.insert(0, ".")
    from quick_bench import run
    run(378880, 16, tile_mode=3, iters=10)
    run(262144, 40, tile_mode=3, iters=10)
    run(131072, 16, cell=1, input_size=3, tile_mode=3, iters=10)
    run(262144, 32, tile_mode=3, iters=10)
else:
    names = sys.argv[1:] or sorted(d for d in os.listdir("tools/bin") if os.path.exists("tools/bin/%s/rt-neural-generic.so" % d))
    for name in ["base"] + names:
        env = dict(os.environ)
        if name != "base":
            env["AIDAX_LIB_DIR"] = os.path.abspath("tools/bin/" + name)
        print("==", name, flush=True)
        subprocess.run([sys.executable, "tools/variant_time.py", "--one"], env=env, timeout=600)
