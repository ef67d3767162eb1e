#!/usr/bin/env python3
"""Compares the reference dumps written by tools/rust_dump (cargo run --release -- DIR) with the CPU oracle and, when a CUDA
device is present, with libowgpu.  Exit code 0 iff every buffer is within the north-star tolerance (max-abs <= 1e-6 full scale,
relative L2 <= 1e-7; f32 engine stream: 2e-7 / 1e-6).  Usage: compare.py DIR"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import openwurli_b200 as ow

d = sys.argv[1]
EV = [(0, 0, 60, 0.8), (4096, 0, 64, 0.6), (8192, 2, 1, 0.0), (12288, 1, 60, 0.0), (16384, 0, 60, 0.9), (24576, 2, 0, 0.0)]
cases = {
    "voice_60_100.f64": ("voice", dict(midi=60, vel=100, dur=2.0)),
    "voice_33_127.f64": ("voice", dict(midi=33, vel=127, dur=1.0)),
    "voice_96_1.f64": ("voice", dict(midi=96, vel=1, dur=1.0)),
    "bench_60_100_static.f64": ("bench", dict(midi=60, vel=100, dur=2.0)),
    "bench_60_100_trem05.f64": ("bench", dict(midi=60, vel=100, dur=2.0, depth=0.5)),
    "bench_40_127_trem10_48k.f64": ("bench", dict(midi=40, vel=127, dur=1.0, sr=48000.0, depth=1.0, volume=0.80, speaker=0.4)),
    "bench_84_64_ldr19k_96k.f64": ("bench", dict(midi=84, vel=64, dur=1.0, sr=96000.0, r_ldr=19000.0, speaker=0.0)),
}
gpu = ow.device_count() > 0
bad = 0


def report(name, who, got, ref, tol_abs, tol_l2):
    global bad
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    l2 = float(np.sqrt((err ** 2).sum() / max((ref.astype(np.float64) ** 2).sum(), 1e-300)))
    ok = err.max() <= tol_abs and l2 <= tol_l2
    bad += 0 if ok else 1
    print(f"{name:32s} {who:7s} max_abs {err.max():.3e}  rel_l2 {l2:.3e}  bit_identical {bool(np.array_equal(got, ref))}  {'ok' if ok else 'FAIL'}")


for name, (kind, kw) in cases.items():
    path = os.path.join(d, name)
    if not os.path.exists(path):
        print("missing", path); bad += 1; continue
    ref = np.fromfile(path, dtype="<f8")
    if kind == "voice":
        report(name, "oracle", O.render_voices([O.voice_job(**kw)])[0][:len(ref)], ref, 1e-6, 1e-7)
        if gpu:
            report(name, "gpu", ow.render_voices([ow.voice_job(kw["midi"], kw["vel"], kw.get("sr", 44100.0), kw["dur"])])[0][:len(ref)], ref, 1e-6, 1e-7)
    else:
        report(name, "oracle", O.render_bench([O.bench_job(**kw)])[0][:len(ref)], ref, 1e-6, 1e-7)
        if gpu:
            j = ow.bench_job(note=kw["midi"], velocity=kw["vel"], duration=kw["dur"], sample_rate=kw.get("sr", 44100.0), ldr=kw.get("r_ldr", 1e6),
                             tremolo_depth=kw.get("depth", 0.0), volume=kw.get("volume", 0.60), speaker=kw.get("speaker", 1.0))
            report(name, "gpu", ow.render_bench([j])[0][:len(ref)], ref, 1e-6, 1e-7)
path = os.path.join(d, "engine_stream.f32")
if os.path.exists(path):
    ref = np.fromfile(path, dtype="<f4")
    report("engine_stream.f32", "oracle", O.render_engines([O.engine_job(EV, sr=44100.0, dur=1.0, volume=0.5, depth=0.5, speaker=0.5)])[0][:len(ref)], ref, 2e-7, 1e-6)
    if gpu:
        report("engine_stream.f32", "gpu", ow.render_engines([ow.engine_job(EV, sample_rate=44100.0, duration=1.0, volume=0.5, tremolo_depth=0.5, speaker_character=0.5)])[0][:len(ref)], ref, 2e-7, 1e-6)
else:
    print("missing", path); bad += 1
sys.exit(1 if bad else 0)
