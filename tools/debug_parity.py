"""Stage-by-stage comparison of the CUDA path with the oracle (run on a GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import nnnoiseless_b200 as nb
from conftest import golden_metric

mb = open(nb.BUILTIN_WEIGHTS_PATH, "rb").read()
om = oracle.Model(mb)
x = np.fromfile(os.path.join(ROOT, "tests/golden/testing.raw"), dtype="<i2").astype(np.float32)[:48000].reshape(100, 480)
ref = np.fromfile(os.path.join(ROOT, "tests/golden/reference_output.raw"), dtype="<i2")

st = oracle.State(om)
b = nb.DenoiseBatch(1)
outs = []
worst = dict(feat=0, gain=0, out=0, vad=0)
pm = 0
for f in range(100):
    oo, ov = st.process_frame(x[f])
    t = st.taps()
    go, gv = b.process_host(x[f][None, None, :])
    tp = b.taps()
    of = np.array(t.features); og = np.array(t.gains)
    df = np.abs(of - tp["features"][0]).max(); dg = np.abs(og - tp["gains"][0]).max() if not t.silence else 0
    do = np.abs(oo - go[0, 0]).max(); dv = abs(ov - gv[0, 0])
    if tp["pitch"][0] != t.pitch or tp["silence"][0] != t.silence: pm += 1
    if f < 5 or tp["pitch"][0] != t.pitch:
        print(f, "pitch", t.pitch, tp["pitch"][0], "sil", t.silence, tp["silence"][0], "dfeat %.3g dgain %.3g dout %.3g dvad %.3g" % (df, dg, do, dv))
    worst["feat"] = max(worst["feat"], df); worst["gain"] = max(worst["gain"], dg)
    worst["out"] = max(worst["out"], do); worst["vad"] = max(worst["vad"], dv)
    if f > 0: outs.append(go[0, 0])
print("pitch/silence mismatches:", pm, "worst", worst)
print("golden metric (gpu):", golden_metric(outs, ref))
