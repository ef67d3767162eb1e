#!/usr/bin/env python3
"""One tremolo plan + execute (64 jobs, 1 s) for timing the oscillator kernels under `ncu --metrics gpu__time_duration.sum`.
Usage: OWG_TREM_KERNEL=thread|tile trem_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow
os.environ["OWG_TREM_CTOR_CACHE"] = "0"
jobs = [ow.bench_job(note=40 + k % 40, velocity=20 + k, duration=1.0, tremolo_depth=0.5) for k in range(64)]
pl = ow.Plan.bench(jobs)
out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
pl.execute(out); torch.cuda.synchronize()
print("done", os.environ.get("OWG_TREM_KERNEL"))
