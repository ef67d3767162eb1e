"""The oracle against its committed golden vectors (tests/golden/oracle_v1.npz, made by make_golden.py).
Setup arithmetic goes through glibc libm, whose last-bit results may differ between CPU models, so the
comparison allows 1e-12 relative; on the box that generated them it is bit-exact."""
import os

import numpy as np

import oracle_lib as O
from golden.make_golden import CASES_B, CASES_V

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz"))


def _close(a, b):
    return np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1e-30) + 1e-18


def test_oracle_voice_golden():
    for i, (m, v, sr, d) in enumerate(CASES_V):
        assert _close(O.render_voices([O.voice_job(m, v, sr=sr, dur=d)])[0], G[f"voice_{i}"]), i


def test_oracle_bench_golden():
    for i, kw in enumerate(CASES_B):
        got = O.render_bench([O.bench_job(**kw)])[0]
        ref = G[f"bench_{i}"]
        assert np.abs(got - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-30), i


def test_oracle_threads_do_not_change_results():
    jobs = [O.bench_job(60 + k, 80, dur=0.02) for k in range(4)]
    assert np.array_equal(O.render_bench(jobs, threads=1), O.render_bench(jobs, threads=4))
