#!/usr/bin/env python3
"""Kernel timing probe (CUDA events inside libowgpu): chain/total ms for the grid at a few sizes; FP64 probes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow
from openwurli_b200.api import fp64_probe

if "--probes" in sys.argv:
    for name, mode in [("DADD ns/dep-op", 10), ("DMUL ns/dep-op", 11), ("DFMA ns/dep-op", 12), ("DDIV ns/dep-op", 13)]:
        print(name, round(fp64_probe(mode), 3), flush=True)
    for lanes in (32, 16, 8, 4):
        print(f"unfused warp-instr rate with {lanes} active lanes: {fp64_probe(100 + lanes):.3f} T warp-instr/s", flush=True)

MODEL = int(os.environ.get("GT_MODEL", "0"))  # 0 melange12, 1 legacy8


def run(depth, dur, stride=1, reps=2):
    jobs = [ow.bench_job(note=33 + k // 127, velocity=1 + k % 127, duration=dur, tremolo_depth=depth) for k in range(0, 8128, stride)]
    pl = ow.Plan.bench(jobs, preamp_model=MODEL)
    out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
    best = None
    for _ in range(reps):
        pl.execute(out); torch.cuda.synchronize()
        t = pl.last_timing()
        best = t if best is None or t[1] < best[1] else best
    print(f"depth={depth} dur={dur} n={len(jobs)}: chain {best[0]:.1f} ms, total {best[1]:.1f} ms -> {len(jobs)*dur/(best[1]*1e-3):.0f} audio-s/s", flush=True)
    pl.close()

if "--scale" in sys.argv:
    k = int(sys.argv[sys.argv.index("--scale") + 1])
    def run_big(depth, dur, k):
        jobs = [ow.bench_job(note=33 + (i // 127) % 64, velocity=1 + i % 127, duration=dur, tremolo_depth=depth, volume=0.3 + 0.5 * (i // 8128) / max(k, 1))
                for i in range(8128 * k)]
        pl = ow.Plan.bench(jobs, preamp_model=MODEL)
        out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
        for _ in range(2):
            pl.execute(out); torch.cuda.synchronize()
        t = pl.last_timing()
        print(f"[throughput regime] depth={depth} dur={dur} n={len(jobs)}: chain {t[0]:.1f} ms, total {t[1]:.1f} ms -> {len(jobs)*dur/(t[1]*1e-3):.0f} audio-s/s", flush=True)
        pl.close()
    for depth in (0.0, 0.5):
        run_big(depth, 0.5, k)
else:
    for depth in (0.0, 0.5):
        for dur in (0.5, 1.5):
            run(depth, dur)
