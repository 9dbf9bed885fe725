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


def test_post_silence_is_ill_conditioned_for_any_f32_fft(builtin_bytes):
    """SURVEY H6, demonstrated instead of asserted: signal -> 45 frames of digital zeros -> signal.  Right after the cut
    the analysis window holds only the smooth tail of the high-pass filter, the upper bands sit at the rounding floor
    of the FFT and the pitch-correlation features divide that floor by itself (src/features.rs:136-137); the GRUs
    remember it.  Three CORRECT FFTs -- the pinned f32 Stockham (mode 0), an f64 DFT rounded once (mode 1) and the same
    f32 Stockham with another radix order (mode 2) -- agree to ~5e-7 before the cut and differ by 1e-4 (whole batch) to
    ~1e-3 (single streams) afterwards, VAD by a few 1e-4; the pitch period stays identical.  This is the yardstick for the
    tolerance of tests/test_gpu_parity.py::test_silence_path_and_recovery (a third f32 FFT, on the GPU)."""
    from conftest import synth_streams
    B = 64
    sig = synth_streams(B, 8, seed=11).reshape(B, 8, 480)
    x = np.concatenate([sig, np.zeros((B, 45, 480), np.float32), sig], axis=1)
    m = oracle.Model(builtin_bytes)
    res = {}
    try:
        for mode in (0, 1, 2):
            oracle.set_fft_mode(mode)
            res[mode] = oracle.run_batch(m, x, n_threads=0)
    finally:
        oracle.set_fft_mode(0)

    def rr(a, b, ax=None):
        a = a.astype(np.float64); b = b.astype(np.float64)
        return np.sqrt(((a - b) ** 2).sum(axis=ax) / np.maximum((b ** 2).sum(axis=ax), 1e-30))

    for a, b in ((0, 1), (2, 1), (0, 2)):
        assert np.array_equal(res[a]["pitch"], res[b]["pitch"])
        assert rr(res[a]["out"][:, :8], res[b]["out"][:, :8]) <= 2e-6                   # before the cut: tight
        after = rr(res[a]["out"][:, 53:], res[b]["out"][:, 53:])
        per = rr(res[a]["out"][:, 53:], res[b]["out"][:, 53:], ax=(1, 2))
        dv = np.abs(res[a]["vad"] - res[b]["vad"]).max()
        assert 3e-5 <= after <= 1e-3, (a, b, after)                                       # >= 50x amplification, bounded
        assert 2e-4 <= per.max() <= 5e-3, (a, b, per.max())                              # single streams reach ~1e-3
        assert 5e-5 <= dv <= 2e-3, (a, b, dv)
    for mode in (0, 1, 2):
        assert not res[mode]["vad"][:, 30:53].any()                                       # silent frames: vad exactly 0
