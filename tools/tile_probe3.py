#!/usr/bin/env python3
"""Tile vs split: timing and BIT-exact output comparison on a grid slice (GPU)."""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow


def counters():
    a = (C.c_uint64 * 9)()
    ow.lib().owg_debug_counters(a, 9, 1)
    return list(a)


def run(stride, depth, dur, env, diag=False):
    for k in ("OWG_CHAIN_KERNEL", "OWG_TILE_IPW"):
        os.environ.pop(k, None)
    os.environ.update(env)
    jobs = [ow.bench_job(note=33 + k // 127, velocity=1 + k % 127, duration=dur, tremolo_depth=depth) for k in range(0, 8128, stride)]
    pl = ow.Plan.bench(jobs, collect_diag=diag)
    out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
    pl.execute(out); torch.cuda.synchronize()
    counters()
    pl.execute(out); torch.cuda.synchronize()
    t = pl.last_timing()
    c = counters()
    res = {"n": len(jobs), "depth": depth, "dur": dur, "env": env, "chain_ms": round(t[0], 2), "us_per_base_sample": round(t[0] * 1e3 / int(dur * 44100), 3)}
    if diag and c[6]:
        res.update({"dk_wait_frac": round(c[0] / max(c[1], 1), 4), "trips_per_warp_step": round(c[4] / c[6], 3), "dk_cycles_per_step": round((c[1] - c[0]) / c[6], 1), "rare_lane_iters": c[8]})
    pl.close()
    print(json.dumps(res), flush=True)
    return out


if __name__ == "__main__":
    for depth, dur in ((0.5, 0.5), (0.5, 2.0), (0.0, 1.0)):
        a = run(4, depth, dur, {"OWG_CHAIN_KERNEL": "split"})
        b = run(4, depth, dur, {"OWG_CHAIN_KERNEL": "tile"})
        print(json.dumps({"bit_identical_tile_vs_split": bool(torch.equal(a, b)), "max_abs_diff": float((a - b).abs().max().item())}), flush=True)
        run(4, depth, dur, {"OWG_CHAIN_KERNEL": "tile"}, diag=True)
        del a, b
    run(1, 0.5, 3.0, {"OWG_CHAIN_KERNEL": "split"})
    run(1, 0.5, 3.0, {"OWG_CHAIN_KERNEL": "tile"})
