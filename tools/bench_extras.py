"""Throughput of the SURVEY §8(f) rows built around the hot path (N2 file front-end, N4 training rows) on one GPU.
Not the headline metric (bench.py is): one JSON line for profiles/."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nnnoiseless_b200 as nb  # noqa: E402
from nnnoiseless_b200 import files, training as tr  # noqa: E402
from nnnoiseless_b200.synth import synth_streams  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    out = {}
    # ---- N4: training rows, device-resident inputs, L lanes x T frames -------------------------------------
    L, T = 16384, 20
    rng = np.random.default_rng(0)
    base = synth_streams(64, T, seed=1).reshape(64, T, 480)
    sig = torch.from_numpy(np.ascontiguousarray(base[rng.integers(0, 64, L)].transpose(1, 0, 2))).cuda()   # [T][L][480]
    noi = (torch.randn((T, L, 480), device="cuda") * 300).round()
    rows = torch.empty((T, L, 87), device="cuda")
    tb = tr.TrainingBatch(L)
    tb.set_params(tr.randomize(L, rng))
    st = torch.cuda.current_stream().cuda_stream
    sec = timed(lambda: tb.process_device(rows.data_ptr(), sig.data_ptr(), noi.data_ptr(), T, 480, L * 480, 87, L * 87, st))
    out["training_rows"] = {"lanes": L, "frames": T, "rows_per_s": L * T / sec, "ms_per_frame_step": 1e3 * sec / T,
                            "note": "3 feature extractors per lane: 3L = %d streams through hp/pitch/analysis" % (3 * L)}
    del tb, sig, noi, rows
    # ---- N2: resampler alone (host buffers in and out) --------------------------------------------------------
    ch, secs_audio = 64, 30
    x = np.ascontiguousarray(synth_streams(ch, 44100 * secs_audio // 480 + 1, seed=3)[:, :44100 * secs_audio].T)
    sec = timed(lambda: files.resample(x, 44100 / 48000), reps=2)
    out["resample_host"] = {"channels": ch, "seconds_of_audio_each": secs_audio, "out_samples_per_s": ch * 48000 * secs_audio / sec,
                            "x_realtime_per_channel_sum": ch * secs_audio / sec}
    # ---- N2: whole files (decode + resample + denoise + encode), 64 mono 44.1 kHz WAVs of 20 s ------------------------
    d = tempfile.mkdtemp()
    n_files, secs_audio = 64, 20
    pcm = synth_streams(n_files, 44100 * secs_audio // 480 + 1, seed=5)[:, :44100 * secs_audio].astype(np.int16)
    pairs = []
    import wave
    for i in range(n_files):
        p = os.path.join(d, "in%03d.wav" % i)
        with wave.open(p, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(44100)
            w.writeframes(pcm[i].tobytes())
        pairs.append((p, os.path.join(d, "out%03d.wav" % i)))
    sec = timed(lambda: files.denoise_files(pairs), reps=2)
    out["denoise_files"] = {"files": n_files, "seconds_of_audio_each": secs_audio, "rate_in": 44100, "wall_s": sec,
                            "x_realtime_sum": n_files * secs_audio / sec, "frames_per_s": n_files * secs_audio * 100 / sec}
    out["kernel_launches"] = nb.kernel_launches()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
