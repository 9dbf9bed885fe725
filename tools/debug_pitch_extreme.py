"""Debug aid: which stream-frames of the extreme-input scenario differ from the oracle, per NNB_PITCH_EXACT mode."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nnnoiseless_b200 as nb
import oracle
from nnnoiseless_b200.synth import synth_mixed
speech = np.fromfile(os.path.join(ROOT, "tests", "golden", "testing.raw"), dtype="<i2")
B, T = 96, 24
x = synth_mixed(B, T, seed=77, speech=speech).reshape(B, T, 480)
scale = np.ones(B, np.float32)
scale[0::12] = 1e-9; scale[1::12] = 3e-5; scale[2::12] = 1e4; scale[3::12] = 1.0 / 32768.0
x = x * scale[:, None, None]
bb = open(nb.BUILTIN_WEIGHTS_PATH, "rb").read()
ref = oracle.run_batch(oracle.Model(bb), x, n_threads=0)
xt = np.ascontiguousarray(x.transpose(1, 0, 2))
for mode in ("0", "1", "2", "3"):
    for rep in range(3):
        os.environ["NNB_PITCH_EXACT"] = mode
        b = nb.DenoiseBatch(B)
        bad = []
        for t in range(T):
            b.process_host(xt[t:t + 1])
            p = b.taps()["pitch"]
            for s in np.argwhere(p != ref["pitch"][:, t])[:, 0]:
                bad.append((t, int(s), float(scale[s]), int(p[s]), int(ref["pitch"][s, t])))
        print("mode", mode, "rep", rep, "mismatches", len(bad), bad[:6], b.pitch_stats())
