"""Which stage differs between identical streams at different batch positions? (run on a GPU box)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nnnoiseless_b200 as nb
from nnnoiseless_b200.synth import synth_streams
T = 12
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x1 = synth_streams(1, T, seed=5).reshape(1, T, 480)
a = nb.DenoiseBatch(1)
b = nb.DenoiseBatch(B)
for t in range(T):
    o1, v1 = a.process_host(np.ascontiguousarray(x1[:, t][None]))
    ta = a.taps()
    xb = np.ascontiguousarray(np.broadcast_to(x1[:, t][None], (1, B, 480)))
    ob, vb = b.process_host(xb)
    tb = b.taps()
    def nd(name, arr, ref):
        bad = np.nonzero(np.any(arr.reshape(B, -1) != ref.reshape(1, -1), axis=1))[0]
        if len(bad):
            i = bad[0]
            d = np.abs(arr.reshape(B, -1)[i].astype(np.float64) - ref.reshape(-1).astype(np.float64)).max()
            print("  frame", t, name, "differs in", len(bad), "streams; first", bad[:8], "maxabs", d)
            if name in ("features", "gains"):
                row = arr.reshape(B, -1)[i]; r0 = ref.reshape(-1)
                idx = np.nonzero(row != r0)[0]
                print("    idx", idx.tolist()); print("    got", row[idx].tolist()); print("    ref", r0[idx].tolist())
    nd("pitch", tb["pitch"], ta["pitch"]); nd("silence", tb["silence"], ta["silence"])
    nd("features", tb["features"], ta["features"]); nd("gains", tb["gains"], ta["gains"])
    nd("out", ob[0], o1[0]); nd("vad", vb[0], v1[0])
print("done")
