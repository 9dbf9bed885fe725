import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def builtin_bytes():
    with open(os.path.join(ROOT, "nnnoiseless_b200", "data", "weights.rnn"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def sh_bytes():
    with open(os.path.join(GOLDEN, "sh.rnn"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def testing_raw():
    """test_data/testing.raw as f32 (src/lib.rs:168-174), first 100 whole frames."""
    x = np.fromfile(os.path.join(GOLDEN, "testing.raw"), dtype="<i2").astype(np.float32)
    return x[: (len(x) // 480) * 480].reshape(-1, 480)


@pytest.fixture(scope="session")
def reference_output():
    return np.fromfile(os.path.join(GOLDEN, "reference_output.raw"), dtype="<i2")


def golden_metric(out_frames, reference_output):
    """src/lib.rs:184-194: `as i16` (truncate, saturate), sum (ref-out)^2 / sum out^2."""
    o = np.concatenate(list(out_frames))
    oi = np.clip(np.trunc(o), -32768, 32767).astype(np.int16)
    assert len(oi) == len(reference_output)
    xx = (oi.astype(np.float64) ** 2).sum()
    d = ((reference_output.astype(np.float64) - oi.astype(np.float64)) ** 2).sum()
    return d / xx, int(np.abs(reference_output.astype(np.int64) - oi).max())


def synth_streams(n_streams, n_frames, seed=1234, start_stream=0):
    """Synthetic white+sine PCM-valued streams, [B][T*480] float32: per stream s a
    Philox(key=seed, counter=s) generator draws f in [100,4000] Hz (log-uniform), A in [1000,12000],
    sigma in [100,3000], phase in [0,2pi); x = clamp(round(A sin(2 pi f n/48000 + phi) + sigma N(0,1)))."""
    n = n_frames * 480
    out = np.empty((n_streams, n), np.float32)
    t = np.arange(n, dtype=np.float64)
    for i in range(n_streams):
        g = np.random.Generator(np.random.Philox(key=seed, counter=start_stream + i))
        f = 100.0 * (40.0 ** g.random())
        a = 1000.0 + 11000.0 * g.random()
        sg = 100.0 + 2900.0 * g.random()
        ph = 2 * np.pi * g.random()
        x = a * np.sin(2 * np.pi * f * t / 48000.0 + ph) + sg * g.standard_normal(n)
        out[i] = np.clip(np.rint(x), -32768, 32767).astype(np.float32)
    return out
