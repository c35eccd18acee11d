"""Scratch: max abs error of each kernel shape on the bundled golden vectors (reference threshold 1e-5)."""
import glob, sys
sys.path.insert(0, ".")
import numpy as np
import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import models
paths = sorted(glob.glob("tests/golden/models/*.json"))
ms = [models.load_json(p) for p in paths]
for tm in (1, 2, 3):
    for math in (0, 1):
        eng = ax.Engine(max_streams=64, max_block=2048)
        eng.configure(tile_mode=tm, math_mode=math)
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
        print("tile", tm, "math", math, " ".join("%.2e" % np.abs(y[i] - m.output_batch).max() for i, m in enumerate(ms)), flush=True)
        eng.close()
