#!/usr/bin/env python3
"""Profiling workload of the melange power amplifier kernel: one wave (2368 rows) x n samples of chain-level signals, launched 3 times.
ncu --set full --import-source on --clock-control none -k regex:pa_melange --launch-skip 1 -c 1 -o gpurun_out/<tag>_pa python tools/pa_prof.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import openwurli_b200 as ow
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 128
sr = 44100.0
t = (2000 + np.arange(ns)) / sr
amps = np.linspace(0.005, 0.15, rows)[:, None]
freqs = 110.0 * 2 ** (np.arange(rows) % 48 / 12.0)[:, None]
x = torch.from_numpy(np.ascontiguousarray(amps * np.sin(2 * np.pi * freqs * t))).cuda()
y = torch.empty_like(x)
for _ in range(3):
    ow.power_amp_batch(x, sr, out=y)
torch.cuda.synchronize()
_, _, cnt = ow.power_amp_batch(x[:64].cpu().numpy(), sr, want_state=True)
print("ok", float(y.abs().max()), "resets/row", float(cnt[:, 0].mean()), "be/row", float(cnt[:, 1].mean()))
