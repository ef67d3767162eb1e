#!/usr/bin/env python3
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import openwurli_b200 as ow
import oracle_lib as O
fs = 44100.0
n = 3000
t = np.arange(n) / fs
x = np.stack([1e-6 * np.sin(2 * np.pi * 220 * t), 1e-3 * np.sin(2 * np.pi * 220 * t), 0.05 * np.sin(2 * np.pi * 220 * t), np.zeros(n)])
for depth, r in ((0.0, 1e6), (0.0, 50000.0), (0.5, 0.0)):
    got = ow.preamp_batch(x, fs, oversample=True, tremolo_depth=depth, r_ldr=r, preamp_model=ow.LEGACY8)
    ref = np.zeros_like(x)
    O.lib().owo_preamp_batch_model(O.dptr(x), n, x.shape[0], n, fs, 1, depth, r, O.dptr(ref), n, 1, O.LEGACY8)
    for i in range(x.shape[0]):
        e = got[i] - ref[i]
        nz = np.nonzero(e)[0]
        print(f"depth {depth} r {r:g} row {i}: peak {np.abs(ref[i]).max():.3e} max_abs {np.abs(e).max():.3e} first nonzero err at {nz[0] if len(nz) else -1} "
              f"err[:8]={np.array2string(e[:8], precision=2)}")
