import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import nnnoiseless_b200 as nb
from nnnoiseless_b200.synth import synth_streams
B = 3
sig = synth_streams(B, 8, seed=11).reshape(B, 8, 480)
x = np.concatenate([sig, np.zeros((B, 4, 480), np.float32)], axis=1)
om = oracle.Model(open(nb.BUILTIN_WEIGHTS_PATH, "rb").read())
sts = [oracle.State(om) for _ in range(B)]
b = nb.DenoiseBatch(B)
np.set_printoptions(linewidth=200, precision=5, suppress=True)
for t in range(x.shape[1]):
    go, gv = b.process_host(np.ascontiguousarray(x[:, t][None]))
    tp = b.taps()
    for s in range(B):
        sts[s].process_frame(x[s, t])
    if t in (8, 9):
        tt = sts[0].taps()
        f0 = np.array(tt.features); f1 = tp["features"][0]
        print("frame", t, "oracle", f0); print("gpu   ", f1); print("diff  ", f1 - f0)
        print("ex", np.array(tt.ex)); print("ep", np.array(tt.ep)); print("exp", np.array(tt.exp))
