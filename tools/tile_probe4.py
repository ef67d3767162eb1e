#!/usr/bin/env python3
"""Quick check of the tile kernel: BIT-exact against the split kernel on a slice, then the C3 chain time (GPU)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow


def run(stride, depth, dur, env):
    for k in ("OWG_CHAIN_KERNEL", "OWG_TILE_IPW"):
        os.environ.pop(k, None)
    os.environ.update(env)
    jobs = [ow.bench_job(note=33 + k // 127, velocity=1 + k % 127, duration=dur, tremolo_depth=depth) for k in range(0, 8128, stride)]
    pl = ow.Plan.bench(jobs)
    out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
    pl.execute(out); torch.cuda.synchronize()
    pl.execute(out); torch.cuda.synchronize()
    t = pl.last_timing()
    pl.close()
    print(json.dumps({"n": len(jobs), "depth": depth, "dur": dur, "env": env, "chain_ms": round(t[0], 2)}), flush=True)
    return out


if __name__ == "__main__":
    for depth, dur in ((0.5, 0.6), (0.0, 0.6)):
        a = run(8, depth, dur, {"OWG_CHAIN_KERNEL": "split"})
        b = run(8, depth, dur, {"OWG_CHAIN_KERNEL": "tile"})
        print(json.dumps({"bit_identical_tile_vs_split": bool(torch.equal(a, b)), "max_abs_diff": float((a - b).abs().max().item())}), flush=True)
        del a, b
    run(1, 0.5, 3.0, {"OWG_CHAIN_KERNEL": "tile"})
    run(1, 0.0, 3.0, {"OWG_CHAIN_KERNEL": "tile"})
