#!/usr/bin/env python3
"""Tiny workload for `compute-sanitizer --tool memcheck`: ragged chain-B batch with odd lengths and an odd row stride (paired 16-byte
stores, zero-filled tails), voices, chain B with the melange power amplifier; checked against the oracle like smoke()."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import openwurli_b200 as ow
import oracle_lib as O
durs = (0.0101, 0.0033, 0.0075)
jobs = [ow.bench_job(note=60, velocity=100, duration=durs[0]), ow.bench_job(note=45, velocity=127, duration=durs[1], tremolo_depth=0.5),
        ow.bench_job(note=72, velocity=30, duration=durs[2], tremolo_depth=0.5)]
out = np.full((3, 447), 9.0)   # odd stride: rows start at odd 8-byte offsets
g = ow.render_bench(jobs, out=out)
c = O.render_bench([O.bench_job(60, 100, dur=durs[0]), O.bench_job(45, 127, dur=durs[1], depth=0.5), O.bench_job(72, 30, dur=durs[2], depth=0.5)], threads=2)
err = 0.0
for i, d in enumerate(durs):
    n = int(d * 44100.0)
    err = max(err, float(np.abs(g[i, :n] - c[i, :n]).max()))
    assert np.all(g[i, n:int(durs[0] * 44100.0)] == 0.0) and np.all(g[i, int(durs[0] * 44100.0):] == 9.0)
v = ow.render_voices([ow.voice_job(60, 100, 44100.0, 0.0071)], out=np.zeros((1, 313)))
p = ow.render_bench(jobs[:1], power_amp_model=ow.PA_MELANGE)
print("memcheck workload ok: chain B max_abs", err, "voice peak", float(np.abs(v).max()), "melange chain peak", float(np.abs(p).max()))
