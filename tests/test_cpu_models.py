"""CPU checks of the two pieces of device logic that can be proven without a GPU (they run in the `-m "not gpu"` suite):

* `nnnoiseless_b200/csrc/fft480.cuh` is host+device code: `tools/fft480_model.cpp` emulates the warp lane by lane (index
  maps, Good-Thomas DFT15, twiddles, in-register FFT32 with bit reversal, real-FFT split, inverse pre-twist) and
  compares with double-precision DFT sums.
* the certificate that lets `pitch_kernel` use FMA sums: `tools/pitch_fast_model.c` runs the same
  decision logic on the CPU against the oracle's order-exact pitch path -- any accepted (unflagged) frame whose period or
  gain differs from the oracle's is a hole in the certificate.
Both are test infrastructure (the second #includes oracle/nno_oracle.c).
"""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path, name, cmd):
    exe = str(tmp_path / name)
    r = subprocess.run(cmd + ["-o", exe], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def test_fft480_decomposition_against_f64_dft(tmp_path):
    exe = _build(tmp_path, "fft480_model", ["g++", "-O2", "tools/fft480_model.cpp"])
    r = subprocess.run([exe], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    errs = [float(x) for x in re.findall(r"rel rms error ([0-9.eE+-]+)", r.stdout)]
    assert len(errs) == 3, r.stdout
    assert max(errs) < 5e-7, r.stdout  # an f32 FFT of this size sits at ~1e-7


def test_pitch_certificate_model_has_no_mismatch(tmp_path):
    exe = _build(tmp_path, "pitch_fast_model",
                 ["gcc", "-O2", "-march=native", "-ffp-contract=off", "-fopenmp", "tools/pitch_fast_model.c", "-lm"])
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    # 512 streams x 60 frames: all four signal families and all eight amplitude regimes (1e-9 ... 1e4 x int16 range)
    r = subprocess.run([exe, "512", "60"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"frames (\d+) flagged (\d+)", r.stdout)
    mm = re.search(r"MISMATCHES among unflagged: period (\d+) gain (\d+)", r.stdout)
    assert m and mm, r.stdout
    frames, flagged = int(m.group(1)), int(m.group(2))
    assert frames == 512 * 60
    assert int(mm.group(1)) == 0 and int(mm.group(2)) == 0, r.stdout + r.stderr
    # the exact path is an exception, not the rule -- except in the amplitude regimes that are routed there on purpose
    assert flagged < 0.5 * frames, r.stdout
