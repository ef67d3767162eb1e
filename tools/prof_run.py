#!/usr/bin/env python3
"""Small fixed workload for ncu: one execution of a grid slice (chain B).
Usage: prof_run.py [depth] [stride] [duration] [reps] [ldr]   (ldr=1e5 exact nominal -> no R-step transient: steady state)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow
depth = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dur = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ldr = float(sys.argv[5]) if len(sys.argv) > 5 else 1e6
if ldr == 1e5:
    ldr = 9.99999999999999854e4
jobs = [ow.bench_job(note=33 + k // 127, velocity=1 + k % 127, duration=dur, tremolo_depth=depth, ldr=ldr) for k in range(0, 8128, stride)]
pl = ow.Plan.bench(jobs)
out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
for _ in range(reps):
    pl.execute(out)
torch.cuda.synchronize()
print("done", len(jobs), pl.last_timing(), pl.kernel_launches)
