"""Scratch: accuracy of every tools/bin/*/libaidax.so (tools/variant.sh) next to the in-tree library: the bundled golden
vectors (2048 samples) and the 10-second full-chain run of tests/test_engine_gpu.py::test_ten_seconds_per_kernel_shape,
tensor-core shape, fast gates.  The oracle's 10-second output is computed once and cached in /tmp."""
import glob, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] != "--only":
    sys.path.insert(0, "tools"); sys.path.insert(0, ".")
    import numpy as np
    import aidadsp_lv2_b200.engine as e
    lib = sys.argv[1]
    if lib != "base":
        e.lib_path = lambda name: lib
    import aidadsp_lv2_b200 as ax
    from aidadsp_lv2_b200 import models, signals, workloads
    from oracle import pyoracle
    T = 128
    paths = sorted(glob.glob("tests/golden/models/*.json"))
    ms = [models.load_json(p) for p in paths]
    # golden vectors
    eng = ax.Engine(max_streams=64, max_block=2048)
    eng.configure(tile_mode=3, math_mode=0)
    sids = []
    for p in paths:
        mid, _ = eng.load_json(p)
        s = eng.open_stream(); eng.set_model(s, mid)
        eng.set_controls(s, ax.default_controls(ANTIALIASING=0.0, DCBLOCKER=0.0, EQBYPASS=1.0), loading=False)
        sids.append(s)
    eng.render_host(np.zeros((len(sids), 128), dtype=np.float32))
    for s, m in zip(sids, ms):
        z = np.zeros(m.hidden, dtype=np.float32)
        eng.set_state(s, h=z, c=z); eng.activate(s)
    y = eng.render_host(np.stack([m.input_batch for m in ms]))
    print("  golden  ", " ".join("%.2e" % np.abs(y[i] - m.output_batch).max() for i, m in enumerate(ms)), flush=True)
    eng.close()
    # ten seconds, as the test does
    ms2 = ms + [workloads.get(n).models[0] for n in ("c2", "c3", "c5")]
    ctl = ax.default_controls(EQPOS=0.0, BASS=3.0, MID=-2.0, TREBLE=4.0, DEPTH=1.0, PRESENCE=2.0, PARAM1=0.4, PARAM2=0.6)
    n, NB = len(ms2), 3750
    x = signals.sweep_noise_block(n, T * NB)
    cache = "/tmp/oracle_10s.npy"
    if os.path.exists(cache):
        yo = np.load(cache)
    else:
        orcs = []
        for m in ms2:
            o = pyoracle.OracleStream(); o.set_model(m); orcs.append(o)
        yo = np.stack([o.run(ctl, x[i]) for i, o in enumerate(orcs)])
        np.save(cache, yo)
    eng = ax.Engine(max_streams=64, max_block=T)
    eng.configure(tile_mode=3, math_mode=0)
    for m in ms2:
        s = eng.open_stream(); eng.set_model(s, eng.register_model(m)); eng.set_controls(s, ctl, loading=False)
    ye = np.empty_like(x)
    for b in range(NB):
        ye[:, b * T:(b + 1) * T] = eng.render_host(np.ascontiguousarray(x[:, b * T:(b + 1) * T]))
    err = np.abs(ye - yo)
    print("  10 s    ", " ".join("%.2e" % v for v in err.max(axis=1)), flush=True)
    eng.close()
else:
    only = sys.argv[2:] if len(sys.argv) > 2 else None
    for lib in ["base"] + sorted(glob.glob("tools/bin/*/libaidax.so")):
        if only and lib != "base" and not any(o in lib for o in only):
            continue
        print("==", lib, flush=True)
        subprocess.run([sys.executable, "tools/acc_run.py", lib])
