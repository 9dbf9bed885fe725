import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import nnnoiseless_b200 as nb
from nnnoiseless_b200.synth import synth_streams
B = 5
sig = synth_streams(B, 8, seed=11).reshape(B, 8, 480)
x = np.concatenate([sig, np.zeros((B, 45, 480), np.float32), sig], axis=1)
mb = open(nb.BUILTIN_WEIGHTS_PATH, "rb").read()
om = oracle.Model(mb)
sts = [oracle.State(om) for _ in range(B)]
b = nb.DenoiseBatch(B)
for t in range(x.shape[1]):
    go, gv = b.process_host(np.ascontiguousarray(x[:, t][None]))
    tp = b.taps()
    line = []
    for s in range(B):
        oo, ov = sts[s].process_frame(x[s, t])
        tt = sts[s].taps()
        num = np.sqrt(((go[0, s].astype(np.float64) - oo) ** 2).sum()); den = np.sqrt((oo.astype(np.float64) ** 2).sum()) + 1e-30
        df = np.abs(np.array(tt.features) - tp["features"][s]).max()
        line.append("%d/%d p%d/%d r%.1e f%.1e v%.1e" % (tt.silence, tp["silence"][s], tt.pitch, tp["pitch"][s], num / den, df, abs(ov - gv[0, s])))
    print(t, " | ".join(line[:3]))
