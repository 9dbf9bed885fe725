"""CPU tests: the oracle against the reference's own golden vector and fixtures."""
import numpy as np
import pytest

import oracle
from conftest import golden_metric


def test_golden_vector(builtin_bytes, testing_raw, reference_output):
    """src/lib.rs:196-213 compare_to_reference, restated: metric < 1e-4 (we get ~1.7e-6, <= 1 LSB)."""
    st = oracle.State(oracle.Model(builtin_bytes))
    outs = []
    for f in range(100):
        o, _ = st.process_frame(testing_raw[f])
        if f > 0:
            outs.append(o)
    metric, maxdiff = golden_metric(outs, reference_output)
    assert metric < 1e-4
    assert metric < 1e-5 and maxdiff <= 1


def test_frozen_intermediates(builtin_bytes, testing_raw):
    """Pitch track / VAD of testing.raw frozen from the survey's independent numpy probe."""
    st = oracle.State(oracle.Model(builtin_bytes))
    pitch, vad = [], []
    for f in range(20):
        _, v = st.process_frame(testing_raw[f])
        pitch.append(st.taps().pitch)
        vad.append(v)
    assert pitch == [203, 185, 60, 208, 212, 762, 288, 423, 379, 437, 406, 410, 409, 420, 416, 420, 414, 427, 320, 292]
    assert np.allclose(vad[:10], [.247, .201, .076, .043, .059, .032, .697, .935, .983, .995], atol=6e-4)


def test_in_place_alias(builtin_bytes, testing_raw):
    """rnnoise_demo.c:52 calls process_frame with out == in."""
    m = oracle.Model(builtin_bytes)
    a, b = oracle.State(m), oracle.State(m)
    import ctypes as C
    for f in range(5):
        o1, v1 = a.process_frame(testing_raw[f])
        buf = testing_raw[f].copy()
        v2 = oracle.lib().nno_process_frame(b._h, buf.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))
        assert np.array_equal(o1, buf) and v1 == v2


def test_fft_against_f64_dft():
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = rng.standard_normal(960).astype(np.float32) * 1000
        X = oracle.rfft960(x)
        Xn = np.fft.rfft(x.astype(np.float64))
        assert np.abs(X - Xn).max() / np.abs(Xn).max() < 1e-6
        assert X[0].imag == 0 and X[480].imag == 0
        y = oracle.irfft960(Xn.astype(np.complex64))
        assert np.abs(y - 960.0 * x).max() / (960.0 * np.abs(x).max()) < 1e-6


def test_activations():
    xs = np.linspace(-9, 9, 2001, dtype=np.float32)
    t = np.array([oracle.lib().nno_tansig(float(v)) for v in xs])
    assert np.abs(t - np.tanh(xs)).max() < 2e-4
    s = np.array([oracle.lib().nno_sigmoid(float(v)) for v in xs])
    assert np.abs(s - 1 / (1 + np.exp(-xs.astype(np.float64)))).max() < 2e-4
    assert oracle.lib().nno_tansig(float("nan")) == 1.0  # reversed tests catch NaN (src/util.rs:30-33)


def test_model_geometry(builtin_bytes, sh_bytes):
    assert oracle.Model(builtin_bytes).describe() == [(42, 24, 0), (24, 24, 2), (90, 48, 2), (114, 96, 2), (96, 22, 1), (24, 1, 1)]
    assert oracle.Model(sh_bytes).describe() == [(42, 24, 0), (24, 24, 0), (90, 48, 2), (114, 96, 0), (96, 22, 1), (24, 1, 1)]


def test_silence_keeps_state(builtin_bytes, testing_raw):
    """E < 0.04 -> zero features, vad 0, RNN/ceps state untouched (src/features.rs:160-166)."""
    st = oracle.State(oracle.Model(builtin_bytes))
    for f in range(10):
        st.process_frame(testing_raw[f])
    for _ in range(40):  # let the high-pass memory decay and flush the 1728-sample history
        o, v = st.process_frame(np.zeros(480, np.float32))
    t = st.taps()
    assert t.silence == 1 and v == 0.0 and not np.any(np.array(t.features))
    assert t.pitch >= 60


def test_batch_driver_matches_single(builtin_bytes, testing_raw):
    m = oracle.Model(builtin_bytes)
    x = np.stack([testing_raw[:20], testing_raw[20:40]])  # [2][20][480]
    r = oracle.run_batch(m, x, n_threads=2)
    for s in range(2):
        st = oracle.State(m)
        for f in range(20):
            o, v = st.process_frame(x[s, f])
            assert np.array_equal(o, r["out"][s, f]) and v == r["vad"][s, f] and st.taps().pitch == r["pitch"][s, f]
