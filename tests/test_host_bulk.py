"""The batch host's period logic with a plugin that offers the bulk entry (AidaxBatchInterface::run_block, ABI 2), on
the CPU with a mock plugin (tests/mock/mock_plugin.c): which instances go through run(), which through run_block(),
the ports-unchanged hint, the worker round trip, set_active and thread pinning.  The real plugin's side of the same
protocol is covered on the GPU (tests/test_plugin_gpu.py runs it against the reference binary)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import aidadsp_lv2_b200 as ax

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mock(tmp_path_factory, native_libs):
    so = str(tmp_path_factory.mktemp("mock") / "mock_plugin.so")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "mock", "mock_plugin.c")], check=True)
    lib = C.CDLL(so)
    return so, lib


def _counts(lib, n):
    runs = np.array([lib.mock_instance(i, 0) for i in range(n)])
    blocks = np.array([lib.mock_instance(i, 1) for i in range(n)])
    return runs, blocks


def test_bulk_periods_route_instances(mock, tmp_path, monkeypatch):
    so, lib = mock
    base = lib.mock_instance  # noqa: F841  (instances are numbered in creation order across hosts: use one host per test module)
    n, T = 300, 64
    host = ax.BatchHost(so, n_instances=n, max_block=T, n_threads=3)
    assert host.batched and host.bulk
    assert len(host.cpus) == 3          # one cpu per period thread (pinning on by default)
    x = np.random.default_rng(0).random((n, T), dtype=np.float32)
    # period 1: every instance reads its ports through run()
    y = host.process(x)
    assert np.array_equal(y, 2 * x)
    runs, blocks = _counts(lib, n)
    assert (runs == 1).all() and (blocks == 0).all()
    assert lib.mock_counter(0) == 1 and lib.mock_counter(1) == 1
    # period 2: nothing pending -> every instance through run_block, with the host vouching for its control ports
    y = host.process(x)
    assert np.array_equal(y, 2 * x)
    runs, blocks = _counts(lib, n)
    assert (runs == 1).all() and (blocks == 1).all()
    assert lib.mock_counter(3) == 1     # AIDAX_RUN_PORTS_UNCHANGED
    # a control write sends exactly that instance through run() once
    host.set_control("MASTER", -3.0, instance=17)
    host.process(x)
    runs, blocks = _counts(lib, n)
    assert runs[17] == 2 and blocks[17] == 1
    assert (np.delete(runs, 17) == 1).all() and (np.delete(blocks, 17) == 2).all()
    host.process(x)
    runs, blocks = _counts(lib, n)
    assert runs[17] == 2 and blocks[17] == 2
    # a patch:Set goes through run() (the event is on the CONTROL port), is worked and answered after the period, and
    # the instance takes one more run() so that its NOTIFY port is reset after the host has read it
    p = str(tmp_path / "m.json")
    open(p, "w").write("{}")
    host.load_model(p, instance=5)
    host.process(x)
    assert lib.mock_instance(5, 2) == 1 and lib.mock_instance(5, 3) == 101   # one event seen, work + work_response
    r5 = lib.mock_instance(5, 0)
    host.process(x)
    assert lib.mock_instance(5, 0) == r5 + 1
    host.process(x)
    assert lib.mock_instance(5, 0) == r5 + 1                                 # back on the bulk path
    # only the first n_active instances take part
    before = _counts(lib, n)
    host.set_active(100)
    host.input[:, :T] = x
    host.output[:] = 0
    host.run(T)
    after = _counts(lib, n)
    assert ((after[0] + after[1]) - (before[0] + before[1]) == (np.arange(n) < 100)).all()
    assert np.array_equal(host.output[:100, :T], 2 * x[:100]) and not host.output[100:].any()
    host.set_active(n)
    host.close()


def test_host_without_bulk_or_hint(mock, monkeypatch):
    so, lib = mock
    first = 300  # instances created by the test above
    monkeypatch.setenv("AIDAX_HOST_NO_PORTS_HINT", "1")
    monkeypatch.setenv("AIDAX_HOST_PIN", "0")
    host = ax.BatchHost(so, n_instances=40, max_block=32, n_threads=2)
    assert host.bulk and host.cpus == []
    x = np.ones((40, 32), dtype=np.float32)
    host.process(x)
    host.process(x)
    assert lib.mock_counter(3) == 0     # no hint: the plugin compares the control rows itself
    host.close()
    monkeypatch.setenv("AIDAX_HOST_NO_BULK", "1")
    host = ax.BatchHost(so, n_instances=40, max_block=32, n_threads=2)
    assert host.batched and not host.bulk
    for _ in range(3):
        assert np.array_equal(host.process(x), 2 * x)
    assert all(lib.mock_instance(first + 40 + i, 0) == 3 and lib.mock_instance(first + 40 + i, 1) == 0 for i in range(40))
    host.close()
