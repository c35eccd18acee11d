"""Scratch: time (and sanity-check against the bundled golden vectors) every tools/bin/*/libaidax.so built by
tools/variant.sh, next to the in-tree library ("base")."""
import glob, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] != "--only":
    sys.path.insert(0, "tools"); sys.path.insert(0, ".")
    import numpy as np
    import aidadsp_lv2_b200.engine as e
    lib = sys.argv[1]
    if lib != "base":
        e.lib_path = lambda name: lib
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import models
    from quick_bench import run
    run(262144, 16, tile_mode=3)
    run(262144, 16, tile_mode=3, iters=40)
    run(262144, 40, tile_mode=3, iters=10)
    run(262144, 32, tile_mode=3, iters=10)
    run(131072, 16, cell=1, input_size=3, tile_mode=3)
    # golden vectors of the bundled LSTM-12 captures through the tensor-core shape (H=12 objects come from build/ unless rebuilt)
    paths = sorted(glob.glob("tests/golden/models/*.json"))
    ms = [models.load_json(p) for p in paths]
    for hsel in (16, 32, 40):
        # synthetic model, tc shape vs FP32 latency shape with precise math: isolates the epilogue approximations
        outs = []
        for tm, math in ((1, 1), (3, 0)):
            eng = ax.Engine(max_streams=128, max_block=2048)
            eng.configure(tile_mode=tm, math_mode=math)
            mid = eng.register_model(models.synthetic(0, hsel, 1, seed=3))
            sids = []
            for i in range(128):
                s = eng.open_stream(); eng.set_model(s, mid)
                eng.set_controls(s, ax.default_controls(EQPOS=1.0), loading=False); sids.append(s)
            rng = np.random.default_rng(5)
            x = (rng.random((128, 2048), dtype=np.float32) - 0.5)
            y = None
            for _ in range(4):
                y = eng.render_host(x)
            outs.append(y); eng.close()
        print("  H=%d tc(fast) vs fp32(precise) after 8192 samples: max abs %.3e" % (hsel, np.abs(outs[0] - outs[1]).max()), flush=True)
else:
    only = sys.argv[2:] if len(sys.argv) > 2 else None
    libs = ["base"] + sorted(glob.glob("tools/bin/*/libaidax.so"))
    for lib in libs:
        if only and lib != "base" and not any(o in lib for o in only):
            continue
        print("==", lib, flush=True)
        subprocess.run([sys.executable, "tools/variant_run.py", lib])
