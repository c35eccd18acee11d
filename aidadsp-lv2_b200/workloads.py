"""The BASELINE.json configurations as data: which models, how many streams, which controls.

C1  single stream, bundled LSTM-12 JSON through run(), EQ on (post), 10 s sweep+noise  (CPU reference case)
C2  1024 streams of one LSTM-16 model, fused pre-EQ, stream-interleaved buffers         (bench headline)
C3  8192 streams of a conditioned GRU-16 (input_size 3: audio + 2 automated params)
C4  65536 streams over 64 heterogeneous models (LSTM/GRU, hidden 8-40, input 1-3)
C5  262144 streams of one LSTM-40 model
Synthetic models are re-created from the committed seed table tests/golden/synthetic_models.json
(chosen by tests/golden/make_synthetic_models.py)."""
import json
import os

from . import models
from .engine import default_controls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED_TABLE = os.path.join(ROOT, "tests", "golden", "synthetic_models.json")
BUFFER = 128
FS = 48000

_FALLBACK = {"c2": dict(cell=0, hidden=16, input_size=1, seed=1000), "c3": dict(cell=1, hidden=16, input_size=3, seed=1001),
             "c5": dict(cell=0, hidden=40, input_size=1, seed=1002)}


def _table():
    if os.path.exists(SEED_TABLE):
        with open(SEED_TABLE) as f:
            return json.load(f)
    t = dict(_FALLBACK)
    hs = (8, 12, 16, 20, 24, 32, 40)
    t["c4"] = [dict(cell=i % 2, hidden=hs[(i // 2) % 7], input_size=1 + (i // 14) % 3, seed=1100 + i) for i in range(64)]
    return t


def _mk(spec):
    return models.synthetic(spec["cell"], spec["hidden"], spec["input_size"], spec["seed"])


def model_flops(m):
    """Algorithmic flop per stream-sample of the network."""
    H, i = m.hidden, m.input_size
    return (8 * H * (H + i) + 10 * H + 1) if m.cell == models.CELL_LSTM else (6 * H * (H + i) + 12 * H + 1)


def chain_flops(controls):
    """Filter/gain stages: 9 flop per active biquad, 4 per gain smoother, 3 for the model gains/skip."""
    c = controls
    n_bq = (1 if c[0] != 0.0 else 0) + (1 if c[17] == 1.0 else 0)
    if c[5] == 0.0 and c[6] in (0.0, 1.0):
        n_bq += 1 if c[12] == 1.0 else 5
    return 9 * n_bq + 2 * 4 + 3


class Workload:
    def __init__(self, name, description, model_list, n_streams, controls, automate_params=False):
        self.name, self.description = name, description
        self.models = model_list          # list of models.Model
        self.n_streams = n_streams
        self.controls = controls          # one control vector for all streams
        self.automate_params = automate_params

    def model_of_stream(self, s, n_total=None):
        """Streams are assigned to models in contiguous runs (sorted by model)."""
        n_total = n_total or self.n_streams
        return min(len(self.models) - 1, (s * len(self.models)) // n_total)

    def flops_per_sample(self, s=0):
        return model_flops(self.models[self.model_of_stream(s)]) + chain_flops(self.controls)

    def mean_flops_per_sample(self):
        return sum(model_flops(m) for m in self.models) / len(self.models) + chain_flops(self.controls)


def get(name, n_streams=None):
    t = _table()
    name = name.lower()
    if name == "c2":
        return Workload("c2", "1024 streams x one LSTM-16 model, fused pre-EQ, stream-interleaved 128-sample buffers",
                        [_mk(t["c2"])], n_streams or 1024, default_controls(EQPOS=1.0))
    if name == "c3":
        return Workload("c3", "8192 streams x conditioned GRU-16 (input_size 3), PARAM1/2 automated per buffer",
                        [_mk(t["c3"])], n_streams or 8192, default_controls(PARAM1=0.5, PARAM2=0.5), automate_params=True)
    if name == "c4":
        return Workload("c4", "65536 streams over 64 heterogeneous models (LSTM/GRU, hidden 8-40, input 1-3)",
                        [_mk(s) for s in t["c4"]], n_streams or 65536, default_controls(PARAM1=0.3, PARAM2=0.7))
    if name == "c5":
        return Workload("c5", "262144 streams x one LSTM-40 model (shared weights)", [_mk(t["c5"])], n_streams or 262144, default_controls())
    if name == "capacity":
        return Workload("capacity", "C2's model and chain (LSTM-16, pre-EQ) at the stream count that fills the GPU",
                        [_mk(t["c2"])], n_streams or 131072, default_controls(EQPOS=1.0))
    raise ValueError("unknown workload %r" % name)
