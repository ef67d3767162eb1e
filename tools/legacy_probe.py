#!/usr/bin/env python3
"""Parity probe for the legacy 8-node preamp path: per-job max-abs / rel-L2 vs the oracle and Newton histograms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import openwurli_b200 as ow
import oracle_lib as O
from test_gpu_parity import to_oracle_b

jobs = [ow.bench_job(note=m, velocity=v, duration=0.3, ldr=r, tremolo_depth=d) for m, v, r, d in
        [(60, 100, 1e6, 0), (33, 127, 1e6, 0), (96, 127, 19000.0, 0), (48, 1, 50000.0, 0), (48, 1, 1e6, 0), (84, 64, 5.0, 0), (57, 90, 2.5e6, 0),
         (60, 100, 1e6, 0.5), (40, 127, 1e6, 1.0), (48, 1, 1e6, 0.5), (72, 30, 1e6, 0.25)]]
got = ow.render_bench(jobs, collect_diag=True, preamp_model=ow.LEGACY8)
dg = ow.last_diag()
ref = O.render_bench([to_oracle_b(j) for j in jobs], threads=4, preamp_model=O.LEGACY8)
dc = O.last_diag()
for i, j in enumerate(jobs):
    e = got[i] - ref[i]
    k = int(np.argmax(np.abs(e)))
    print(f"job {i}: midi {j.v.midi} vel {j.v.velocity*127:.0f} ldr {j.r_ldr:g} depth {j.tremolo_depth}: peak {np.abs(ref[i]).max():.3e} "
          f"max_abs {np.abs(e).max():.3e} at {k} rel_l2 {np.linalg.norm(e)/max(np.linalg.norm(ref[i]),1e-300):.3e} first>1e-12 at {int(np.argmax(np.abs(e)>1e-12))}")
print("gpu hist", list(dg.nr_iter_hist)[:8]); print("cpu hist", list(dc.nr_iter_hist)[:8])
