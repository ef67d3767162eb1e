#!/usr/bin/env python3
"""Tile-kernel diagnostics (GPU): wait/total cycles of DK and I/O warps and Newton trip statistics (owg_debug_counters) per config."""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow


def counters(reset=True):
    a = (C.c_uint64 * 9)()
    ow.lib().owg_debug_counters(a, 9, 1 if reset else 0)
    return list(a)


def run(stride, depth, dur, env, diag, no_pa=False):
    for k in ("OWG_CHAIN_KERNEL", "OWG_TILE_IPW", "OWG_LANES_PER_WARP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    jobs = [ow.bench_job(note=33 + k // 127, velocity=1 + k % 127, duration=dur, tremolo_depth=depth, no_poweramp=no_pa) for k in range(0, 8128, stride)]
    pl = ow.Plan.bench(jobs, collect_diag=diag)
    out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
    pl.execute(out); torch.cuda.synchronize()
    counters()
    pl.execute(out); torch.cuda.synchronize()
    t = pl.last_timing()
    c = counters()
    n_samp = int(dur * 44100)
    res = {"n": len(jobs), "depth": depth, "dur": dur, "env": env, "diag": diag, "no_pa": no_pa, "chain_ms": round(t[0], 2), "us_per_base_sample": round(t[0] * 1e3 / n_samp, 3)}
    if diag and c[6]:
        res.update({"dk_wait_frac": round(c[0] / max(c[1], 1), 4), "io_wait_frac": round(c[2] / max(c[3], 1), 4), "trips_per_warp_step": round(c[4] / c[6], 3),
                    "iters_per_inst_step": round(c[5] / max(c[7], 1), 3), "dk_cycles_per_step": round((c[1] - c[0]) / c[6], 1), "rare_lane_iters": c[8]})
    pl.close()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    for dur in (0.5, 2.0):
        for stride in (1, 8):
            run(stride, 0.5, dur, {"OWG_CHAIN_KERNEL": "tile"}, True)
            run(stride, 0.5, dur, {"OWG_CHAIN_KERNEL": "tile"}, False)
            run(stride, 0.5, dur, {"OWG_CHAIN_KERNEL": "tile"}, False, no_pa=True)
            run(stride, 0.5, dur, {"OWG_CHAIN_KERNEL": "split"}, False)
        run(8, 0.5, dur, {"OWG_CHAIN_KERNEL": "tile", "OWG_TILE_IPW": "4"}, True)
        run(8, 0.5, dur, {"OWG_CHAIN_KERNEL": "tile", "OWG_TILE_IPW": "2"}, True)
        run(8, 0.5, dur, {"OWG_CHAIN_KERNEL": "tile", "OWG_TILE_IPW": "1"}, True)
