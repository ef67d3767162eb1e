#!/usr/bin/env python3
"""One-thread vs 8-lane Twin-T oscillator: BIT-exact output comparison and timing on a small tremolo batch (GPU).
The batch is small enough that the render is bound by the serial oscillator, so the wall time of an execute is its time."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow

os.environ["OWG_TREM_CTOR_CACHE"] = "0"   # exercise Tremolo::new in both kernels


def run(mode, dur, rates=(44100.0,), diag=False):
    os.environ["OWG_TREM_KERNEL"] = mode
    jobs = [ow.bench_job(note=40 + k % 40, velocity=20 + k, duration=dur, tremolo_depth=(0.5 if k % 2 else 1.0), sample_rate=rates[k % len(rates)])
            for k in range(64)]
    t0 = time.time()
    pl = ow.Plan.bench(jobs, collect_diag=diag)
    out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
    pl.execute(out); torch.cuda.synchronize()
    t1 = time.time()
    pl.execute(out); torch.cuda.synchronize()
    t2 = time.time()
    res = {"mode": mode, "dur": dur, "rates": list(rates), "plan_plus_first_execute_s": round(t1 - t0, 3), "execute_s": round(t2 - t1, 3),
           "us_per_osc_step": round((t2 - t1) * 1e6 / (dur * 88200.0), 3)}
    import ctypes as C
    cnt = (C.c_uint64 * 18)()
    ow.lib().owg_debug_counters(cnt, 18, 1)
    res["generic_iters"] = int(cnt[17])
    if diag:
        d = ow.last_diag()
        res["trm_hist"] = list(d.tremolo_nr_iter_hist)[:8]; res["trm_be"] = int(d.tremolo_be_fallback)
    pl.close()
    print(json.dumps(res), flush=True)
    return out


if __name__ == "__main__":
    for dur, rates in ((1.0, (44100.0,)), (0.5, (44100.0, 48000.0, 96000.0))):
        a = run("thread", dur, rates)
        b = run("tile", dur, rates)
        print(json.dumps({"bit_identical_tile_vs_thread": bool(torch.equal(a, b)), "max_abs_diff": float((a - b).abs().max().item()),
                          "peak": float(a.abs().max().item())}), flush=True)
    run("tile", 1.0, diag=True)
    run("thread", 1.0, diag=True)
