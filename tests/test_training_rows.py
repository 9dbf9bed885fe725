"""training-data rows (src/training.rs).  CPU: the oracle's restatement against hand-checked
properties of the reference's loop; GPU: rnnoise_train_* against the oracle, lane by lane."""
import numpy as np
import pytest

import oracle
from nnnoiseless_b200.synth import synth_streams

EBAND_5MS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40, 48, 60, 78, 100]  # src/lib.rs:55-58


def _inputs(L, T, seed):
    rng = np.random.default_rng(seed)
    sig = synth_streams(L, T, seed=seed).reshape(L, T, 480)       # tone + noise, i16-valued
    noise = np.round(rng.normal(0.0, 1.0, (L, T, 480)) * rng.uniform(50, 3000, (L, 1, 1))).astype(np.float32)
    # make the frame energy sweep the VAD thresholds 1e7 / 1e8 / 1e9 (src/training.rs:380-397)
    scale = rng.choice([0.01, 0.05, 0.3, 1.0, 2.5], size=(L, T, 1)).astype(np.float32)
    sig = np.clip(np.round(sig * scale), -32768, 32767).astype(np.float32)
    return sig, noise


def _oracle_rows(sig, noise, params_by_frame):
    L, T, _ = sig.shape
    rows = np.empty((T, L, 87), np.float32)
    for l in range(L):
        tr = oracle.Trainer()
        for t in range(T):
            if t in params_by_frame:
                tr.set_params(params_by_frame[t][l:l + 1])
            rows[t, l] = tr.frame(sig[l, t], noise[l, t])
    return rows


def test_band_lp_matches_reference_expression():
    for lowpass in list(range(0, 500, 7)) + [59, 60, 399, 400, 401, 480]:
        exp = next((i for i, x in enumerate(EBAND_5MS) if (x << 2) > lowpass), 21)
        assert oracle.train_band_lp(lowpass) == exp


def test_oracle_row_layout_and_masks():
    T = 24
    sig, noise = _inputs(3, T, seed=5)
    from nnnoiseless_b200.training import SIM_PARAMS_DTYPE
    p = np.zeros(3, SIM_PARAMS_DTYPE)
    p["signal_gain"], p["noise_gain"], p["band_lp"] = [1.0, 1.0, 0.0], [1.0, 0.5, 0.0], [21, 9, 21]
    rows = _oracle_rows(sig, noise, {0: p})
    assert rows.shape == (T, 3, 87)
    assert set(np.unique(rows[:, :, 86])) <= {0.0, 0.5, 1.0}           # vad levels
    g = rows[:, :, 42:64]
    assert np.all((g == -1.0) | ((g >= 0.0) & (g <= 1.0)))
    assert np.all(g[:, 1, 10:] == -1.0)                                 # band_lp = 9 -> bands >= 10 masked
    assert np.all(g[-1, 0, :] >= 0.0)                                   # band_lp = 21 -> cutoff 22, nothing masked
    # both gains zero: digital silence in all three extractors -> zero features, all gains masked, log10(0.01) noise
    assert np.all(rows[:, 2, :42] == 0.0) and np.all(rows[:, 2, 42:64] == -1.0)
    assert np.allclose(rows[:, 2, 64:86], -2.0, atol=1e-6)
    # a gain is sqrt(clean / combined) <= 1: with noise_gain > 0 most low bands are strictly inside (0, 1)
    assert np.any((g[:, 0, :8] > 0.0) & (g[:, 0, :8] < 1.0))


@pytest.mark.gpu
def test_training_rows_match_oracle():
    from nnnoiseless_b200 import training as tr
    L, T = 70, 90                                                       # 70 lanes: crosses the 64-lane block boundary
    sig, noise = _inputs(L, T, seed=11)
    rng = np.random.default_rng(3)
    p0 = tr.default_params(L)
    p1 = tr.randomize(L, rng)
    p1["signal_gain"][0], p1["noise_gain"][0] = 0.0, 0.0               # silence lane
    p1["noise_gain"][1] = 0.0                                           # vad == 0 && noise_gain == 0 -> cutoff 0
    p2 = tr.randomize(L, rng)
    p2["signal_gain"][0], p2["noise_gain"][0] = 0.0, 0.0               # stays silent: the high-pass tail needs ~40 frames to die
    sched = {0: p0, 12: p1, 24: p2}
    want = _oracle_rows(sig, noise, sched)

    tb = tr.TrainingBatch(L)
    got = np.empty_like(want)
    for t0, t1 in ((0, 12), (12, 24), (24, T)):                         # randomize() happens at frame boundaries
        tb.set_params(sched[t0])
        got[t0:t1] = tb.process_host(np.ascontiguousarray(sig[:, t0:t1].transpose(1, 0, 2)),
                                     np.ascontiguousarray(noise[:, t0:t1].transpose(1, 0, 2)))
    # vad and the mask pattern come from order-exact arithmetic (frame energy, f64 biquads): exact
    assert np.array_equal(got[:, :, 86], want[:, :, 86])
    gm, wm = got[:, :, 42:64] == -1.0, want[:, :, 42:64] == -1.0
    assert (gm != wm).mean() < 1e-3                                      # only the 5e-2 energy threshold can differ
    ok = ~(gm | wm)
    assert np.max(np.abs(got[:, :, 42:64][ok] - want[:, :, 42:64][ok])) < 1e-4
    assert np.max(np.abs(got[:, :, 64:86] - want[:, :, 64:86])) < 1e-4   # log10 band energies of the noise
    # features: same f32 tolerance class as the denoise path (different FFT butterfly order)
    df = np.abs(got[:, :, :42] - want[:, :, :42])
    assert np.sqrt(np.mean(df[:, 2:] ** 2)) < 1e-4 and df[:, 2:].max() < 5e-3
    wz, gz = (want[:, :, :42] == 0).all(-1), (got[:, :, :42] == 0).all(-1)   # silent frames: zero features (:160-166)
    assert np.array_equal(wz, gz) and wz[:, 0].sum() > 10 and not wz[:, 1:].any()
    assert np.array_equal(got[:, :, 40], want[:, :, 40])                 # pitch feature = 0.01 (T - 300): exact period


@pytest.mark.gpu
def test_training_device_api_and_chunking_bitwise():
    import torch
    from nnnoiseless_b200 import training as tr
    L, T = 33, 10
    sig, noise = _inputs(L, T, seed=2)
    a = tr.TrainingBatch(L)
    ref = a.process_host(np.ascontiguousarray(sig.transpose(1, 0, 2)), np.ascontiguousarray(noise.transpose(1, 0, 2)))
    # device API, lane-major layouts, on a user stream
    ds, dn = torch.from_numpy(sig).cuda(), torch.from_numpy(noise).cuda()     # [L][T][480]
    rows = torch.zeros((L, T, 87), device="cuda")
    st = torch.cuda.Stream()
    b = tr.TrainingBatch(L)
    with torch.cuda.stream(st):
        b.process_device(rows.data_ptr(), ds.data_ptr(), dn.data_ptr(), 4, T * 480, 480, T * 87, 87, st.cuda_stream)
        b.process_device(rows.data_ptr() + 4 * 87 * 4, ds.data_ptr() + 4 * 480 * 4, dn.data_ptr() + 4 * 480 * 4, T - 4, T * 480, 480,
                         T * 87, 87, st.cuda_stream)
    st.synchronize()
    assert np.array_equal(rows.cpu().numpy().transpose(1, 0, 2), ref)
