#!/usr/bin/env python3
"""Throughput of the other BASELINE.json configs (C1, C2, C4 slice, C5 slice) on one GPU, with the CPU oracle timed on
a bounded sample of the same jobs.  Prints one JSON object; results are copied into BASELINE.md section 3."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import openwurli_b200 as ow
import oracle_lib as O

THREADS = O.lib().owo_hardware_threads() or os.cpu_count() or 1
res = {"host_threads": THREADS}


def sync():
    torch.cuda.synchronize()


def timed(fn, reps=2):
    fn(); sync()
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter(); fn(); sync(); best = min(best, time.perf_counter() - t)
    return best


# C1a / C1b: single note (latency of one render through the public API)
t = timed(lambda: ow.Voice.render_note(60, 100 / 127.0, 2.0, 44100.0))
tc = time.perf_counter(); O.render_voices([O.voice_job(60, 100, dur=2.0)]); tc = time.perf_counter() - tc
res["C1a_voice_single_note_2s"] = {"gpu_s": t, "gpu_audio_s_per_s": 2.0 / t, "cpu_1thread_audio_s_per_s": 2.0 / tc}
for depth in (0.0, 0.5):
    t = timed(lambda: ow.render_bench([ow.bench_job(note=60, velocity=100, duration=2.0, tremolo_depth=depth)]), reps=1)
    tc = time.perf_counter(); O.render_bench([O.bench_job(60, 100, dur=2.0, depth=depth)]); tc = time.perf_counter() - tc
    res[f"C1b_chainB_single_note_2s_depth{depth:g}"] = {"gpu_s": t, "gpu_audio_s_per_s": 2.0 / t, "cpu_1thread_audio_s_per_s": 2.0 / tc}

# C2: preamp-only batch, 4096 instances, 48 kHz base (preamp at 96 kHz), tremolo depth 0.5, 2 s
n_inst, fs, dur = 4096, 48000.0, 2.0
n_samp = int(fs * dur)
i = np.arange(n_inst)
f = 55.0 * (2093.0 / 55.0) ** ((i % 64) / 63.0)
amp = np.array([0.001, 0.002, 0.005, 0.010])[(i // 64) % 4]
ph = 2 * np.pi * ((i // 256) % 16) / 16.0
x = torch.from_numpy(amp[:, None] * np.sin(2 * np.pi * f[:, None] * np.arange(n_samp)[None, :] / fs + ph[:, None])).cuda()
y = torch.empty_like(x)
t = timed(lambda: ow.preamp_batch(x, fs, oversample=True, tremolo_depth=0.5, out=y), reps=1)
ns = 2 * THREADS
xs = x[:: n_inst // ns][:ns].cpu().numpy().copy()
ys = np.zeros_like(xs)
tc = time.perf_counter()
O.lib().owo_preamp_batch(O.dptr(xs), n_samp, ns, n_samp, fs, 1, 0.5, 0.0, O.dptr(ys), n_samp, THREADS)
tc = time.perf_counter() - tc
err = float(np.abs(y[:: n_inst // ns][:ns].cpu().numpy() - ys).max())
res["C2_preamp_batch_4096x2s_48k_depth0.5"] = {"gpu_s": t, "gpu_audio_s_per_s": n_inst * dur / t, "cpu_audio_s_per_s": ns * dur / tc, "cpu_threads": THREADS,
                                               "cpu_sample_instances": ns, "max_abs_err_vs_oracle_on_sample": err}
del x, y

# C4 slice: volume x depth x speaker (8 x 8 x 8) over 64 keys, velocity 100, 0.5 s, metrics mode
jobs = [ow.calibrate_job(33 + k, 100, volume=(a + 1) / 8.0, tremolo_depth=b / 7.0, speaker=c / 7.0)
        for k in range(64) for a in range(8) for b in range(8) for c in range(8)]
t = timed(lambda: ow.render_bench_metrics(jobs), reps=1)
res["C4_slice_8x8x8x64keys_0.5s_metrics"] = {"renders": len(jobs), "gpu_s": t, "gpu_audio_s_per_s": len(jobs) * 0.5 / t,
                                              "full_C4_2097152_renders_estimated_s_per_gpu": t * 2097152 / len(jobs)}

# C4 as BASELINE.json states it (volume x tremolo depth x speaker character, 32 x 32 x 32, over 64 keys), a 64-key slice of the depth axis:
# 4 depths x 32 volumes x 32 speaker settings x 64 keys = 262 144 renders (1/8 of the config); voice + preamp run once per (key, depth)
jobs = [ow.calibrate_job(33 + k, 100, volume=0.05 + 0.95 * a / 31.0, speaker=c / 31.0, tremolo_depth=b / 3.0)
        for k in range(64) for b in range(4) for a in range(32) for c in range(32)]
t = timed(lambda: ow.render_bench_metrics(jobs), reps=1)
res["C4_eighth_4x32x32x64keys_0.5s_metrics_shared_prefix"] = {"renders": len(jobs), "distinct_prefixes": 64 * 4, "gpu_s": t,
                                                              "gpu_audio_s_per_s": len(jobs) * 0.5 / t,
                                                              "full_C4_2097152_renders_estimated_s_per_gpu": t * 2097152 / len(jobs)}

# C5 slice: polyphonic engine streams at 96 kHz with stealing, block 512, warm-up on
def stream_events(seed, dur, sr):
    s = seed * 2654435761 % (2 ** 32) or 1
    def rnd():
        nonlocal s
        s ^= (s << 13) & 0xFFFFFFFF; s ^= s >> 17; s ^= (s << 5) & 0xFFFFFFFF
        return s / 4294967296.0
    ev, t_s, pedal = [], 0.0, False
    next_pedal = 3.0
    while True:
        t_s += -np.log(max(rnd(), 1e-12)) / 40.0  # Poisson, 40 note-ons per second
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

n_eng, dur, sr = 2048, 1.0, 96000.0
evs = [stream_events(k + 1, dur, sr) for k in range(n_eng)]
ej = [ow.engine_job(e, sample_rate=sr, duration=dur) for e in evs]
out = torch.empty((n_eng, int(dur * sr)), dtype=torch.float32, device="cuda")
t = timed(lambda: ow.render_engines(ej, out=out), reps=1)
d = ow.last_diag()
ns = max(THREADS, 4)
tc = time.perf_counter()
ref = O.render_engines([O.engine_job(evs[k], sr=sr, dur=dur) for k in range(ns)], threads=THREADS)
tc = time.perf_counter() - tc
got = out[:ns].cpu().numpy()
res["C5_slice_engines_96k"] = {"streams": n_eng, "seconds_each": dur, "gpu_s": t, "gpu_audio_s_per_s": n_eng * dur / t,
                               "cpu_audio_s_per_s": ns * dur / tc, "cpu_threads": THREADS, "note_ons": int(d.nr_iter_hist[0]), "steals": int(d.nr_iter_hist[1]),
                               "max_active_voices": int(d.nr_iter_hist[3]),
                               "max_abs_err_vs_oracle_on_sample": float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())}
print(json.dumps(res, indent=1))
