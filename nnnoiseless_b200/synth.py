"""Synthetic workload generator shared by tests and bench."""
import numpy as np


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
