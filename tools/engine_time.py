#!/usr/bin/env python3
"""Timing + parity probe for the WurliEngine stream path (BASELINE config 5 slices).
Usage: engine_time.py [n_engines] [seconds] [sample_rate] [notes_per_second] [--no-check]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import openwurli_b200 as ow

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n_eng = int(args[0]) if len(args) > 0 else 2048
dur = float(args[1]) if len(args) > 1 else 1.0
sr = float(args[2]) if len(args) > 2 else 96000.0
rate = float(args[3]) if len(args) > 3 else 40.0


def stream_events(seed):
    s = seed * 2654435761 % (2 ** 32) or 1
    def rnd():
        nonlocal s
        s ^= (s << 13) & 0xFFFFFFFF; s ^= s >> 17; s ^= (s << 5) & 0xFFFFFFFF
        return s / 4294967296.0
    ev, t_s, pedal, next_pedal = [], 0.0, False, 3.0
    while True:
        t_s += -np.log(max(rnd(), 1e-12)) / rate
        if t_s >= dur:
            break
        note = 33 + int(rnd() * 64)
        vel = float(np.float32(0.2 + 0.8 * rnd()))
        ev.append((int(t_s * sr), ow.NOTE_ON, note, vel))
        off = t_s + 0.05 + 1.95 * rnd()
        if off < dur:
            ev.append((int(off * sr), ow.NOTE_OFF, note, 0.0))
        if t_s > next_pedal:
            pedal = not pedal
            ev.append((int(t_s * sr), ow.SUSTAIN, 1 if pedal else 0, 0.0))
            next_pedal += 3.0
    ev.sort(key=lambda e: e[0])
    return ev

evs = [stream_events(k + 1) for k in range(n_eng)]
ej = [ow.engine_job(e, sample_rate=sr, duration=dur) for e in evs]
out = torch.empty((n_eng, int(dur * sr)), dtype=torch.float32, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ow.render_engines(ej, out=out)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    d = ow.last_diag()
    print(f"engines={n_eng} dur={dur} sr={sr:.0f}: {t:.3f} s -> {n_eng*dur/t:.0f} audio-s/s; note_ons={d.nr_iter_hist[0]} steals={d.nr_iter_hist[1]} "
          f"freed={d.nr_iter_hist[2]} max_active={d.nr_iter_hist[3]} launches={d.kernels_launched}", flush=True)
if "--no-check" not in sys.argv:
    import oracle_lib as O
    ns = 4
    ref = O.render_engines([O.engine_job(evs[k], sr=sr, dur=dur) for k in range(ns)], threads=4)
    got = out[:ns].cpu().numpy().astype(np.float64)
    print("max abs err vs oracle on", ns, "streams:", float(np.abs(got - ref.astype(np.float64)).max()), flush=True)
