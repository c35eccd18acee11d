"""Host-side logic that needs no GPU: model-file acceptance rules, control decode, sharding (gloo, world 2)."""
import json
import os

import numpy as np
import pytest

import aidadsp_lv2_b200 as ax
from aidadsp_lv2_b200 import engine, models, sharding
from oracle import pyoracle
from conftest import GOLDEN, bundled_model_paths


def test_check_json_accepts_bundled_models(native_libs):
    for p in bundled_model_paths():
        d = engine.check_json(p)
        assert (d.cell, d.hidden, d.input_size, d.in_skip) == (0, 12, 1, 0)
        assert d.in_gain == 1.0 and d.out_gain == 1.0
        assert d.samplerate == 48000.0  # metadata.samplerate is the STRING "48000" -> ignored -> default


def test_check_json_rules(native_libs, tmp_path):
    """loadModelFromPath acceptance rules (rt-neural-generic.cpp:972-1044)."""
    base = models.to_json_dict(models.synthetic(models.CELL_GRU, 16, 3, seed=3, in_skip=1, in_gain_db=-6.0, out_gain_db=3.0))

    def check(mut):
        d = json.loads(json.dumps(base))
        mut(d)
        p = str(tmp_path / "x.json")
        json.dump(d, open(p, "w"))
        return engine.check_json(p)

    d = check(lambda d: None)
    assert (d.cell, d.hidden, d.input_size, d.in_skip) == (1, 16, 3, 1)
    assert abs(d.in_gain - 10 ** (-6 / 20)) < 1e-6 and abs(d.out_gain - 10 ** (3 / 20)) < 1e-6
    assert check(lambda d: d.update(samplerate=44100)).samplerate == 44100.0
    assert check(lambda d: d["metadata"].update(samplerate=96000)).samplerate == 96000.0
    assert check(lambda d: d.update(in_gain=-100.0)).in_gain == 0.0  # DB_CO: <= -90 dB is silence
    for bad in (lambda d: d.update(in_shape=[None, None, 4]),          # input_size > MAX_INPUT_SIZE
                lambda d: d.update(in_skip=2),                          # in_skip > 1
                lambda d: d["layers"][0].update(type="conv1d"),         # unknown architecture
                lambda d: d["layers"][0].update(shape=[None, None, 17]),  # hidden size not in the variant table
                lambda d: d["layers"][0]["weights"][1].append([0.0] * 48),  # more recurrent rows than hidden
                lambda d: d.pop("layers"),
                lambda d: d.pop("in_shape")):
        with pytest.raises(engine.AidaxError) as ei:
            check(bad)
        assert ei.value.code == 3
    with pytest.raises(engine.AidaxError):
        engine.check_json(os.path.join(GOLDEN, "rtneural", "bad_lstm.json"))
    with pytest.raises(engine.AidaxError):
        engine.check_json(str(tmp_path / "missing.json"))
    p = str(tmp_path / "garbage.json")
    open(p, "w").write("{ not json")
    with pytest.raises(engine.AidaxError):
        engine.check_json(p)


def test_decode_controls_flags_and_targets(native_libs):
    """Exact float tests of run() (SURVEY.md 8b 'Control decoding')."""
    F = dict(ENABLED=1, LPF=2, EQ_PRE=4, EQ_POST=8, BANDPASS=16, DC=32, NET=64)
    flags, pre, master, coefs = engine.decode_controls(ax.default_controls())
    assert flags == F["ENABLED"] | F["LPF"] | F["EQ_POST"] | F["DC"] | F["NET"]
    assert pre == 1.0 and master == 1.0
    flags, _, master, _ = engine.decode_controls(ax.default_controls(), loading=True)
    assert master == 0.0
    flags, _, _, _ = engine.decode_controls(ax.default_controls(EQPOS=1.0, ANTIALIASING=0.0, DCBLOCKER=0.5, NETBYPASS=0.51, MTYPE=1.0))
    assert flags == F["ENABLED"] | F["EQ_PRE"] | F["BANDPASS"]
    flags, _, _, _ = engine.decode_controls(ax.default_controls(EQPOS=0.5))  # any other EQPOS value disables the EQ
    assert flags & (F["EQ_PRE"] | F["EQ_POST"]) == 0
    flags, _, _, _ = engine.decode_controls(ax.default_controls(EQBYPASS=1.0, ENABLED=0.5))
    assert flags & (F["EQ_PRE"] | F["EQ_POST"] | F["ENABLED"]) == 0
    flags, _, _, _ = engine.decode_controls(ax.default_controls(), model_bound=False)
    assert flags & F["NET"] == 0
    _, pre, master, _ = engine.decode_controls(ax.default_controls(PREGAIN=-90.0, MASTER=15.0))
    assert pre == 0.0 and abs(master - 10 ** 0.75) < 1e-5


def test_decode_controls_coefficients_match_oracle(native_libs):
    c = ax.default_controls(BASS=3.0, BFREQ=120.0, MID=-2.0, MFREQ=900.0, MIDQ=1.3, TREBLE=4.0, TFREQ=3000.0, DEPTH=1.0, PRESENCE=-2.0, ANTIALIASING=30.0)
    _, _, _, coefs = engine.decode_controls(c, samplerate=44100.0)
    sr = 44100.0
    f32 = np.float32
    lpf_fc = float(((f32(30.0) - f32(0)) * (f32(0.25) * f32(0.5) - f32(0.99) * f32(0.5)) / (f32(100.0) - f32(0))) + f32(0.99) * f32(0.5))
    q = float(f32(0.707))
    expect = [
        pyoracle.biquad_coefs(0, lpf_fc, q, 0.0),
        pyoracle.biquad_coefs(4, float(f32(75.0)) / sr, q, 1.0),
        pyoracle.biquad_coefs(5, float(f32(120.0)) / sr, q, 3.0),
        pyoracle.biquad_coefs(4, float(f32(900.0)) / sr, float(f32(1.3)), -2.0),
        pyoracle.biquad_coefs(6, float(f32(3000.0)) / sr, q, 4.0),
        pyoracle.biquad_coefs(6, float(f32(900.0)) / sr, q, -2.0),
        pyoracle.biquad_coefs(1, float(f32(35.0)) / sr, q, 0.0),
    ]
    for i in range(7):
        assert np.array_equal(coefs[i], expect[i]), i


def test_shard_range_partitions():
    for n, w in [(1024, 1), (1024, 2), (65536, 8), (1000, 8), (7, 4), (262144, 8)]:
        ranges = [sharding.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
        for s in (0, n // 3, n - 1):
            lo, hi = ranges[sharding.shard_of(s, n, w)]
            assert lo <= s < hi


def _gloo_worker(rank, world, port, n_streams, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_streams, rank, world)
    # every rank renders only its own streams; the only cross-rank step of the bench is the timing reduction
    owned = torch.zeros(n_streams, dtype=torch.int32)
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, int(owned.min()), int(owned.max()), float(t)))
    dist.destroy_process_group()


def test_sharding_world_size_2_gloo():
    """N>1 path on CPU: two ranks partition the streams disjointly and completely; max-over-ranks timing reduce."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, mn, mx, tmax in res:
        assert mn == 1 and mx == 1  # every stream owned exactly once
        assert tmax == 2.0
