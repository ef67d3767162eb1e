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


# ---- v2 vectors: legacy preamp, engine stream, calibrate rows, render-midi ---------------------------------------------------------
from golden.make_golden import CAL_CFG, CAL_NOTES, CAL_VELS, CASES_L, ENGINE_EVENTS, MIDI_EVENTS  # noqa: E402

G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v2.npz"))


def _midi_oracle(model):
    import ctypes as C
    n = int((MIDI_EVENTS[-1][0] + 0.05) * 44100.0)
    arr = (O.MidiEvent * len(MIDI_EVENTS))(*[O.MidiEvent(t, k, a if k != 2 else 0, b if k == 0 else (b if k == 2 else 0), 0, 0)
                                              for t, k, a, b in MIDI_EVENTS])
    y = np.zeros(n)
    assert O.lib().owo_render_midi(C.cast(arr, C.c_void_p), len(MIDI_EVENTS), n, 0.6, 1.0, 0, model, O.dptr(y), None) == 0
    return y


def test_oracle_v2_golden():
    for i, kw in enumerate(CASES_L):
        got = O.render_bench([O.bench_job(**kw)], preamp_model=O.LEGACY8)[0]
        assert np.abs(got - G2[f"legacy_bench_{i}"]).max() <= 1e-8, i      # the legacy solver's dead zone makes it libm-sensitive at 1e-9
    for model in (0, 1):
        tol = 1e-9 if model == 0 else 1e-7
        e = O.render_engines([O.engine_job(ENGINE_EVENTS, sr=44100.0, dur=0.1, depth=0.5, speaker=0.5, warm_up=False)], preamp_model=model)[0]
        assert np.abs(e.astype(np.float64) - G2[f"engine_{model}"].astype(np.float64)).max() <= max(tol, 6e-8), model
        c = O.calibrate_rows(CAL_NOTES, CAL_VELS, CAL_CFG, preamp_model=model)
        assert np.abs(c - G2[f"calibrate_{model}"]).max() <= (1e-6 if model == 0 else 1e-3), model
        assert np.abs(_midi_oracle(model) - G2[f"midi_{model}"]).max() <= tol, model
