#!/usr/bin/env python3
"""Generates tests/golden/oracle_v1.npz from the CPU oracle (oracle/).  The reference itself cannot be run
here (Rust, no toolchain), so these vectors pin the ORACLE's behaviour (regression guard) and give the GPU
tests a committed, box-independent comparison target.  Regenerate with:  python tests/golden/make_golden.py"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import oracle_lib as O

CASES_V = [(60, 100, 44100.0, 0.05), (33, 1, 44100.0, 0.05), (96, 127, 48000.0, 0.03), (72, 64, 96000.0, 0.02)]
CASES_B = [dict(midi=60, vel=100, dur=0.05), dict(midi=40, vel=127, dur=0.05, depth=0.5),
           dict(midi=84, vel=30, dur=0.03, sr=96000.0, volume=0.9, speaker=0.0),
           dict(midi=52, vel=90, dur=0.03, sr=48000.0, r_ldr=19000.0, speaker=0.4)]


def main():
    out = {}
    for i, (m, v, sr, d) in enumerate(CASES_V):
        out[f"voice_{i}"] = O.render_voices([O.voice_job(m, v, sr=sr, dur=d)])[0]
    for i, kw in enumerate(CASES_B):
        out[f"bench_{i}"] = O.render_bench([O.bench_job(**kw)])[0]
    np.savez_compressed(os.path.join(HERE, "oracle_v1.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
