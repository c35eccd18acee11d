"""Scratch: render a few periods of one shape so ncu can capture the dominant kernel.
   python tools/prof_one.py [N] [H] [warm] [iters] [tile_mode] [cell] [in]"""
import sys
sys.path.insert(0, ".")
import torch
import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import models

a = [int(v) for v in sys.argv[1:]]
n, hidden, warm, iters, tile_mode, cell, insz = (a + [378880, 16, 3, 5, 0, 0, 1][len(a):])[:7]
eng = ax.Engine(max_streams=n, max_block=128)
eng.configure(tile_mode=tile_mode, math_mode=0)
mid = eng.register_model(models.synthetic(cell, hidden, insz, seed=1))
c = ax.default_controls(EQPOS=1.0)
for i in range(n):
    s = eng.open_stream(); eng.set_model(s, mid); eng.set_controls(s, c, loading=False)
x = torch.rand(128, n, device="cuda") - 0.5
y = torch.empty_like(x)
for _ in range(warm + iters):
    eng.render_device(x.data_ptr(), y.data_ptr(), n, 128, layout=ax.LAYOUT_INTERLEAVED)
eng.synchronize()
print("rendered %d periods of %d streams, %d launches" % (warm + iters, n, eng.launch_count))
eng.close()
