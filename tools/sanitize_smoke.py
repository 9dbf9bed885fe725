"""Small end-to-end run used under compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nnnoiseless_b200 as nb
from nnnoiseless_b200.synth import synth_streams
B, T = 37, 6
x = np.ascontiguousarray(synth_streams(B, T, seed=5).reshape(B, T, 480).transpose(1, 0, 2))
b = nb.DenoiseBatch(B)
o, v = b.process_host(x)
o16, _ = nb.DenoiseBatch(B).process_pcm16_host(x.astype(np.int16))
z, _ = nb.DenoiseBatch(3).process_host(np.zeros((3, 3, 480), np.float32))
print("ok", float(np.abs(o).mean()), int(np.abs(o16).max()), float(np.abs(z).max()))
