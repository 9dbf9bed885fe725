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


from nnnoiseless_b200.synth import synth_streams  # noqa: E402,F401
