#!/usr/bin/env python3
"""Tiny melange power-amplifier batch for `compute-sanitizer --tool racecheck --kernel-name kns=pa_melange`: 3 rows (one full warp of two
tiles + a warp with a lone tile) x 10 samples, one of them hot enough for the backward-Euler retry; checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import openwurli_b200 as ow
import oracle_lib as O
n = 10
t = np.arange(n) / 44100.0
x = np.ascontiguousarray(np.stack([0.02 * np.sin(2 * np.pi * 440 * t), 0.6 * np.sin(2 * np.pi * 2000 * t), np.zeros(n)]))
y, rails, cnt = ow.power_amp_batch(x, 44100.0, want_state=True)
err = 0.0
for i in range(3):
    yo, r = np.zeros(n), np.zeros(2 * n)
    O.lib().owo_power_amp_melange(44100.0, 1, O.dptr(np.ascontiguousarray(x[i])), n, O.dptr(yo), O.dptr(r), -1)
    err = max(err, float(np.abs(y[i] - yo).max()))
print("racecheck workload ok: max_abs", err, "be retries", cnt[:, 1].tolist())
