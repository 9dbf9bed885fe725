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


def synth_mixed(n_streams, n_frames, seed=4321, speech=None, block=256):
    """Harder material for the pitch-exactness sweeps, [B][T*480] float32, four families by stream index mod 4:
    0: white+sine as synth_streams; 1: harmonic stack (2..11 partials, 1/h roll-off) with vibrato and a little noise;
    2: `speech` (1-D int16 array, e.g. the reference's test_data/testing.raw) looped from a random offset with a random
    gain, with runs of digital silence; 3: nearly pure tones (the near-tie stress case for find_best_pitch)."""
    n = n_frames * 480
    out = np.empty((n_streams, n), np.float32)
    t = np.arange(n, dtype=np.float64)
    fr = (np.arange(n) // 480)
    for b0 in range(0, n_streams, block):
        b1 = min(n_streams, b0 + block)
        m = b1 - b0
        g = np.random.Generator(np.random.Philox(key=seed, counter=b0))
        f0 = 100.0 * (40.0 ** g.random(m))[:, None]
        a = (1000.0 + 11000.0 * g.random(m))[:, None]
        sg = (100.0 + 2900.0 * g.random(m))[:, None]
        ph = (2 * np.pi * g.random(m))[:, None]
        vib = (0.002 + 0.02 * g.random(m))[:, None]
        vrate = (3.0 + 5.0 * g.random(m))[:, None]
        nh = 2 + (g.random(m) * 10).astype(np.int64)
        gain = (0.05 + 1.5 * g.random(m))[:, None]
        off = (g.random(m) * (len(speech) if speech is not None else 1)).astype(np.int64)
        noise = g.standard_normal((m, n)).astype(np.float32)
        fam = (np.arange(b0, b1) % 4)
        x = np.empty((m, n), np.float64)
        for i in range(m):
            if fam[i] == 0:
                x[i] = a[i] * np.sin(2 * np.pi * f0[i] * t / 48000.0 + ph[i]) + sg[i] * noise[i]
            elif fam[i] == 1:
                fi = np.maximum(f0[i] * 0.25 * (1.0 + vib[i] * np.sin(2 * np.pi * vrate[i] * t / 48000.0)), 60.0)
                phase = ph[i] + 2 * np.pi * np.cumsum(fi) / 48000.0
                v = np.zeros(n)
                for h in range(1, int(nh[i]) + 1):
                    v += a[i] / h * np.sin(h * phase)
                x[i] = v + 0.3 * sg[i] * noise[i]
            elif fam[i] == 2 and speech is not None:
                idx = (off[i] + np.arange(n)) % len(speech)
                v = gain[i] * speech[idx].astype(np.float64) + 0.02 * sg[i] * noise[i]
                v[((fr // 13) % 5) == 4] = 0.0
                x[i] = v
            else:
                s = (1.0 + 20.0 * g.random()) if (b0 + i) % 8 == 3 else 0.1 * sg[i]
                x[i] = a[i] * np.sin(2 * np.pi * f0[i] * t / 48000.0 + ph[i]) + s * noise[i]
        out[b0:b1] = np.clip(np.rint(x), -32768, 32767).astype(np.float32)
    return out
