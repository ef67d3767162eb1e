#!/usr/bin/env python3
"""Timing probe of the melange power amplifier on the device: owg_power_amp_batch on n rows x seconds of a chain-level signal, and chain B
(C3-style grid slice) with power_amp_model = PA_MELANGE next to the behavioural amplifier; the CPU oracle on one row beside it.
Usage: pa_time.py [n_rows] [seconds]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import openwurli_b200 as ow
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8128
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
sr = 44100.0
ns = int(secs * sr)
t = np.arange(ns) / sr
amps = np.linspace(0.005, 0.15, n)[:, None]
freqs = 110.0 * 2 ** (np.arange(n) % 48 / 12.0)[:, None]
x = torch.from_numpy(np.ascontiguousarray(amps * np.sin(2 * np.pi * freqs * t) * np.exp(-2.0 * t))).cuda()
y = torch.empty_like(x)
ow.power_amp_batch(x[:64], sr, out=y[:64])  # settled-state cache, module load
res = {}
for tag, rows in (("full", n), ("one_wave", min(n, 2368))):
    torch.cuda.synchronize(); t0 = time.time()
    ow.power_amp_batch(x[:rows], sr, out=y[:rows]); torch.cuda.synchronize()
    dt = time.time() - t0
    res[tag] = dict(rows=rows, seconds=dt, audio_s_per_s=rows * secs / dt, us_per_sample_per_wave=dt / ns * 1e6)
_, rails, cnt = ow.power_amp_batch(x[:256].cpu().numpy(), sr, want_state=True)
res["resets_per_row_mean"] = float(cnt[:, 0].mean()); res["be_retries_per_row_mean"] = float(cnt[:, 1].mean())
try:
    import oracle_lib as O
    L = O.lib()
    xr = np.ascontiguousarray(x[n // 2].cpu().numpy()); yo = np.zeros(ns); r = np.zeros((ns, 2))
    t0 = time.time(); resets = L.owo_power_amp_melange(sr, 1, O.dptr(xr), ns, O.dptr(yo), O.dptr(r), -1); dt = time.time() - t0
    res["cpu_oracle_one_row"] = dict(seconds=dt, audio_s_per_s=secs / dt, resets=int(resets), max_abs_diff_vs_gpu=float(np.abs(yo - y[n // 2].cpu().numpy()).max()))
except Exception as e:
    res["cpu_oracle_one_row"] = str(e)
# chain B slice with both amplifiers
jobs = [ow.bench_job(note=33 + (i % 64), velocity=1 + (i * 7) % 127, duration=min(secs, 1.0), tremolo_depth=0.5) for i in range(min(n, 2032))]
for tag, pam in (("chain_behavioral", ow.PA_BEHAVIORAL), ("chain_melange_pa", ow.PA_MELANGE)):
    out = torch.empty((len(jobs), int(min(secs, 1.0) * sr)), dtype=torch.float64, device="cuda")
    ow.render_bench(jobs[:8], out=out[:8], power_amp_model=pam)
    torch.cuda.synchronize(); t0 = time.time()
    ow.render_bench(jobs, out=out, power_amp_model=pam); torch.cuda.synchronize()
    dt = time.time() - t0
    res[tag] = dict(renders=len(jobs), seconds=dt, audio_s_per_s=len(jobs) * min(secs, 1.0) / dt, peak=float(out.abs().max()))
print(json.dumps(res))
