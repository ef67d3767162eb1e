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


# v2: paths added later in round 1 -- legacy 8-node preamp, engine stream, calibrate rows, render-midi
CASES_L = [dict(midi=60, vel=100, dur=0.05), dict(midi=45, vel=120, dur=0.05, depth=0.7), dict(midi=88, vel=60, dur=0.03, r_ldr=19000.0)]
ENGINE_EVENTS = [(0, 0, 60, 0.8), (512, 0, 64, 0.6), (1024, 2, 1, 0.0), (1536, 1, 60, 0.0), (2560, 2, 0, 0.0), (3072, 0, 67, 0.9)]
CAL_NOTES, CAL_VELS, CAL_CFG = [48, 72], [40, 127], (0.75, 0.75, 0.02, 0.82, -35.0, -0.04)
MIDI_EVENTS = [(0.0, 0, 60, 100), (0.01, 0, 64, 80), (0.03, 2, 0, 127), (0.04, 1, 60, 0), (0.06, 2, 0, 0), (0.07, 1, 64, 0)]


def main_v2():
    import ctypes as C
    out = {}
    for i, kw in enumerate(CASES_L):
        out[f"legacy_bench_{i}"] = O.render_bench([O.bench_job(**kw)], preamp_model=O.LEGACY8)[0]
    for model in (0, 1):
        out[f"engine_{model}"] = O.render_engines([O.engine_job(ENGINE_EVENTS, sr=44100.0, dur=0.1, depth=0.5, speaker=0.5, warm_up=False)],
                                                  preamp_model=model)[0]
        out[f"calibrate_{model}"] = O.calibrate_rows(CAL_NOTES, CAL_VELS, CAL_CFG, preamp_model=model)
        n = int((MIDI_EVENTS[-1][0] + 0.05) * 44100.0)
        arr = (O.MidiEvent * len(MIDI_EVENTS))(*[O.MidiEvent(t, k, a if k != 2 else 0, b if k == 0 else (b if k == 2 else 0), 0, 0)
                                                  for t, k, a, b in MIDI_EVENTS])
        y = np.zeros(n)
        assert O.lib().owo_render_midi(C.cast(arr, C.c_void_p), len(MIDI_EVENTS), n, 0.6, 1.0, 0, model, O.dptr(y), None) == 0
        out[f"midi_{model}"] = y
    np.savez_compressed(os.path.join(HERE, "oracle_v2.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


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
    main_v2()
