#!/usr/bin/env python3
"""Run-to-run determinism of the full C3 grid (device output): executes a plan several times and reports differing samples."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow
depth = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
dur = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
jobs = [ow.bench_job(note=33 + k, velocity=v, duration=dur, tremolo_depth=depth) for k in range(64) for v in range(1, 128)]
pl = ow.Plan.bench(jobs)
out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
if "--zero-first" in sys.argv:
    out.zero_()
if "--nan-first" in sys.argv:
    out.fill_(float("nan"))
torch.cuda.synchronize()
pl.execute(out)
first = out.clone()
prev = None
for rep in range(3):
    out.zero_()
    pl.execute(out)
    same = torch.equal(out, first)
    s1, s2 = first.sum(dim=1), out.sum(dim=1)
    print(f"rep {rep}: equal={same} row-sum equal={torch.equal(s1, s2)} equal-to-previous-rep={None if prev is None else torch.equal(out, prev)}", flush=True)
    prev = out.clone()
    if not same:
        d = (out != first)
        rows = d.any(dim=1).nonzero().flatten()
        print("  differing rows:", rows.numel(), rows[:10].tolist())
        r = int(rows[0])
        cols = d[r].nonzero().flatten()
        print(f"  row {r}: {cols.numel()} differing samples, first at {int(cols[0])}, max abs diff {(out[r]-first[r]).abs().max().item():.3e}")
        c0 = int(cols[0])
        print("  first:", first[r, c0:c0 + 4].tolist(), "\n  now:  ", out[r, c0:c0 + 4].tolist())
