#!/usr/bin/env python3
"""Time of the Twin-T oscillator constructor (Tremolo::new: 50 + 2*sr steps on one device thread) = plan creation of one tremolo job."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow
job = [ow.bench_job(note=60, velocity=100, duration=0.01, tremolo_depth=0.5)]
ow.render_bench(job)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pl = ow.Plan.bench(job)
    torch.cuda.synchronize()   # device-wide: includes the constructor launched on the plan's oscillator stream
    dt = time.perf_counter() - t0
    pl.close()
    print(f"oscillator constructor: {dt*1e3:.1f} ms for {50 + 176400} steps -> {dt*1e6/(50+176400):.3f} us/step", flush=True)
