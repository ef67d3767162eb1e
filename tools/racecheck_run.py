#!/usr/bin/env python3
"""Tiny chain-B batch for `compute-sanitizer --tool racecheck --kernel-name kns=chain_tile`: 0.01 s notes, static LDR and tremolo,
checked against the oracle like smoke() (racecheck slows the instrumented kernel by orders of magnitude)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import openwurli_b200 as ow
import oracle_lib as O
jobs = [ow.bench_job(note=60, velocity=100, duration=0.01), ow.bench_job(note=45, velocity=127, duration=0.01, tremolo_depth=0.5),
        ow.bench_job(note=72, velocity=30, duration=0.01, tremolo_depth=0.5)]
g = ow.render_bench(jobs)
c = O.render_bench([O.bench_job(60, 100, dur=0.01), O.bench_job(45, 127, dur=0.01, depth=0.5), O.bench_job(72, 30, dur=0.01, depth=0.5)], threads=2)
print("racecheck workload ok: max_abs", float(np.abs(g - c).max()))
