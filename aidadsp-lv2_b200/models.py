"""Model files in the reference's JSON schema (exporter field names from
modules/RTNeural/python/model_utils.py:66-101): read, write and synthesize.

Weight layout (Keras order, what RTN/model_loader.h:149-263 reads):
  layers[0].weights = [ W[in][G*H], U[H][G*H], b ]   G = 4 (lstm: i,f,c,o; b[4H]) or 3 (gru: z,r,h; b[2][3H])
  layers[1].weights = [ w[H][1], b[1] ]
"""
import json
import math
from dataclasses import dataclass, field

import numpy as np

CELL_LSTM, CELL_GRU = 0, 1
HIDDEN_SIZES = (8, 12, 16, 20, 24, 32, 40, 64, 80)  # variant/generate_variant_hpp.py:6


def db_co(g):
    """DB_CO, rt-neural-generic.h:160 (float arithmetic)."""
    g = np.float32(g)
    return float(np.float32(10.0) ** (g * np.float32(0.05))) if g > -90.0 else 0.0


@dataclass
class Model:
    cell: int
    hidden: int
    input_size: int
    W: np.ndarray
    U: np.ndarray
    b: np.ndarray
    dense_w: np.ndarray
    dense_b: float
    in_skip: int = 0
    in_gain: float = 1.0   # linear
    out_gain: float = 1.0  # linear
    samplerate: float = 48000.0
    in_gain_db: float = None
    out_gain_db: float = None
    input_batch: np.ndarray = None
    output_batch: np.ndarray = None
    meta: dict = field(default_factory=dict)

    @property
    def cell_name(self):
        return "lstm" if self.cell == CELL_LSTM else "gru"

    @property
    def n_gates(self):
        return 4 if self.cell == CELL_LSTM else 3


def from_json_dict(d):
    """Same acceptance rules as loadModelFromPath (rt-neural-generic.cpp:972-1013)."""
    input_size = int(d["in_shape"][-1])
    if input_size > 3:
        raise ValueError("Value for input_size not supported")
    in_skip = int(d["in_skip"]) if isinstance(d.get("in_skip"), (int, float)) else 0
    if in_skip > 1:
        raise ValueError("Values for in_skip > 1 are not supported")
    ig_db = d.get("in_gain") if isinstance(d.get("in_gain"), (int, float)) else None
    og_db = d.get("out_gain") if isinstance(d.get("out_gain"), (int, float)) else None
    sr = 48000.0
    md = d.get("metadata") or {}
    if isinstance(md.get("samplerate"), (int, float)) and not isinstance(md.get("samplerate"), bool):
        sr = float(md["samplerate"])
    elif isinstance(d.get("samplerate"), (int, float)):
        sr = float(d["samplerate"])
    rnn, dense = d["layers"][0], d["layers"][1]
    cell = {"lstm": CELL_LSTM, "gru": CELL_GRU}[rnn["type"]]
    H = int(rnn["shape"][-1])
    if H not in HIDDEN_SIZES:
        raise ValueError("Unable to identify a known model architecture!")
    G = 4 if cell == CELL_LSTM else 3
    W = np.asarray(rnn["weights"][0], dtype=np.float32).reshape(input_size, G * H)
    U = np.asarray(rnn["weights"][1], dtype=np.float32).reshape(H, G * H)
    b = np.asarray(rnn["weights"][2], dtype=np.float32)
    b = b.reshape(4 * H) if cell == CELL_LSTM else b.reshape(2, 3 * H)
    dw = np.asarray(dense["weights"][0], dtype=np.float32).reshape(H)
    db = float(np.float32(dense["weights"][1][0]))
    m = Model(cell, H, input_size, W, U, b, dw, db, in_skip,
              db_co(ig_db) if ig_db is not None else 1.0, db_co(og_db) if og_db is not None else 1.0, sr, ig_db, og_db)
    if isinstance(d.get("input_batch"), list):
        m.input_batch = np.asarray(d["input_batch"], dtype=np.float32)
        m.output_batch = np.asarray(d["output_batch"], dtype=np.float32)
    m.meta = md
    return m


def load_json(path):
    with open(path, "r") as f:
        return from_json_dict(json.load(f))


def to_json_dict(m):
    H = m.hidden
    d = {
        "in_shape": [None, None, m.input_size],
        "layers": [
            {"type": m.cell_name, "activation": "", "shape": [None, None, H],
             "weights": [np.asarray(m.W, dtype=np.float64).reshape(m.input_size, -1).tolist(),
                         np.asarray(m.U, dtype=np.float64).reshape(H, -1).tolist(),
                         np.asarray(m.b, dtype=np.float64).tolist()]},
            {"type": "dense", "activation": "", "shape": [None, None, 1],
             "weights": [np.asarray(m.dense_w, dtype=np.float64).reshape(H, 1).tolist(), [float(m.dense_b)]]},
        ],
        "metadata": dict(m.meta) if m.meta else {"name": "synthetic", "samplerate": str(int(m.samplerate))},
    }
    if m.in_skip:
        d["in_skip"] = int(m.in_skip)
    if m.in_gain_db is not None:
        d["in_gain"] = float(m.in_gain_db)
    if m.out_gain_db is not None:
        d["out_gain"] = float(m.out_gain_db)
    if m.samplerate != 48000.0:
        d["samplerate"] = float(m.samplerate)
    return d


def save_json(m, path):
    # float32 values survive the float64 repr round trip exactly
    with open(path, "w") as f:
        json.dump(to_json_dict(m), f)


def synthetic(cell, hidden, input_size, seed, in_skip=0, in_gain_db=None, out_gain_db=None, scale=1.0):
    """Synthetic model as SURVEY.md 8(d) prescribes: weights U(-1/sqrt(H), 1/sqrt(H)) (PyTorch RNN
    init, what the AIDA-X trainer starts from), dense bias 0.1*U(-1,1); deterministic in `seed`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    G = 4 if cell == CELL_LSTM else 3
    k = scale / math.sqrt(hidden)
    W = rng.uniform(-k, k, size=(input_size, G * hidden)).astype(np.float32)
    U = rng.uniform(-k, k, size=(hidden, G * hidden)).astype(np.float32)
    if cell == CELL_LSTM:
        b = rng.uniform(-k, k, size=(4 * hidden,)).astype(np.float32)
    else:
        b = rng.uniform(-k, k, size=(2, 3 * hidden)).astype(np.float32)
    dw = rng.uniform(-k, k, size=(hidden,)).astype(np.float32)
    db = float(np.float32(0.1 * rng.uniform(-1, 1)))
    return Model(cell, hidden, input_size, W, U, b, dw, db, in_skip,
                 db_co(in_gain_db) if in_gain_db is not None else 1.0,
                 db_co(out_gain_db) if out_gain_db is not None else 1.0, 48000.0, in_gain_db, out_gain_db,
                 meta={"name": "synthetic-%s-%d-%d-seed%d" % ("lstm" if cell == CELL_LSTM else "gru", hidden, input_size, seed),
                       "samplerate": "48000"})
