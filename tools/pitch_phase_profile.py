"""Per-phase cycle counts of pitch_kernel.  Build the profiling variant here (NNB_VARIANT=prof python -m nnnoiseless_b200.build),
run on a GPU box with NNB_LIB=nnnoiseless_b200/lib/libnnnoiseless_b200_prof.so python tools/pitch_phase_profile.py [B]."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nnnoiseless_b200 as nb
from nnnoiseless_b200.synth import synth_streams
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = synth_streams(64, 6, seed=3).reshape(64, 6, 480)
xt = np.ascontiguousarray(np.tile(x.transpose(1, 0, 2), (1, B // 64, 1)))
b = nb.DenoiseBatch(B)
b.process_host(xt[:2])
L = nb.lib()
buf = (C.c_ulonglong * 16)()
L.nnb_pitch_prof_read(buf, 1)
b.process_host(xt[2:])
L.nnb_pitch_prof_read(buf, 0)
nblk = (B + 15) // 16 * 4  # 4 profiled frames
names = ["downsample", "autocorr", "lpc", "fir", "yn chain+energies+coarse fma", "certified select", "exact recompute", "fine windows",
         "fine select+lags", "rd fma", "ladder", "exact rd+final"]
tot = sum(buf)
for i in range(12):
    if buf[i]:
        print("%-14s %9.0f cycles/block  %5.1f%%" % (names[i], buf[i] / nblk, 100.0 * buf[i] / tot))
print("total %.0f cycles/block" % (tot / nblk))
