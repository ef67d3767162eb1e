#!/usr/bin/env python3
"""Chain-kernel A/B timing probe (GPU): device time of the chain launches (plan.last_timing()[0]) for several kernels / tile
geometries on slices of the C3 grid.  Usage: tile_probe.py [quick]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow


def run(stride, depth, dur, env, reps=2):
    for k in ("OWG_CHAIN_KERNEL", "OWG_TILE_IPW", "OWG_LANES_PER_WARP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    jobs = [ow.bench_job(note=33 + k // 127, velocity=1 + k % 127, duration=dur, tremolo_depth=depth) for k in range(0, 8128, stride)]
    pl = ow.Plan.bench(jobs)
    out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
    best = None
    for _ in range(reps):
        pl.execute(out)
        torch.cuda.synchronize()
        t = pl.last_timing()
        best = t if best is None or t[0] < best[0] else best
    n_samp = int(dur * 44100)
    res = {"n": len(jobs), "depth": depth, "dur": dur, "env": env, "chain_ms": round(best[0], 2), "total_ms": round(best[1], 2),
           "us_per_base_sample": round(best[0] * 1e3 / n_samp, 3), "checksum": float(out.double().abs().sum().item())}
    pl.close()
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    dur = 0.5
    for depth in (0.0, 0.5):
        for stride in (1, 2, 8):
            run(stride, depth, dur, {"OWG_CHAIN_KERNEL": "split"})
            run(stride, depth, dur, {"OWG_CHAIN_KERNEL": "tile"})
            if stride >= 2:
                run(stride, depth, dur, {"OWG_CHAIN_KERNEL": "tile", "OWG_TILE_IPW": "4"})
            if stride >= 8:
                run(stride, depth, dur, {"OWG_CHAIN_KERNEL": "tile", "OWG_TILE_IPW": "1"})
