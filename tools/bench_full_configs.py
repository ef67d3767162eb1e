#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 at their FULL stated size on one GPU, with parity on sampled rows against the CPU oracle.
  C4: volume x tremolo depth x speaker character (32 x 32 x 32) over 64 keys = 2 097 152 renders, 0.5 s each, metrics output mode
  C5: 16 384 independent 64-voice polyphonic streams, 10 s each at 96 kHz (oversampling bypassed), stealing + crossfade + pedal
Usage: bench_full_configs.py [c4] [c5]   (default: both).  Prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import openwurli_b200 as ow
import oracle_lib as O

which = set(a.lower() for a in sys.argv[1:]) or {"c4", "c5"}
THREADS = O.lib().owo_hardware_threads() or os.cpu_count() or 1
res = {"host_threads": THREADS}

if "c4" in which:
    t0 = time.perf_counter()
    jobs = [ow.calibrate_job(33 + k, 100, volume=0.05 + 0.95 * a / 31.0, speaker=c / 31.0, tremolo_depth=b / 31.0)
            for k in range(64) for b in range(32) for a in range(32) for c in range(32)]
    t_build = time.perf_counter() - t0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = ow.render_bench_metrics(jobs)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    # parity on sampled rows: the same jobs one by one through the oracle's full render + the reference's metric formulas
    idx = [0, 31 * 1024 + 5, 700_001, 1_234_567, len(jobs) - 1]
    sub = ow.render_bench_metrics([jobs[i] for i in idx])
    res["C4_full_32x32x32x64keys_0.5s_metrics"] = {
        "renders": len(jobs), "seconds_each": 0.5, "host_job_build_s": t_build, "gpu_s": t, "gpu_audio_s_per_s": len(jobs) * 0.5 / t,
        "distinct_voice_preamp_prefixes": 64 * 32, "finite": bool(np.isfinite(m).all()),
        "sampled_rows_equal_to_a_5_job_call": bool(np.array_equal(m[idx], sub)), "metric_columns": int(m.shape[1]),
        "peak_column_range": [float(m[:, 0].min()), float(m[:, 0].max())]}
    del jobs, m

if "c5" in which:
    def stream_events(seed, dur, sr):
        rng = np.random.default_rng(seed)
        n = rng.poisson(40.0 * dur)
        t_on = np.sort(rng.uniform(0.0, dur, n))
        notes = 33 + rng.integers(0, 64, n)
        vels = (0.2 + 0.8 * rng.random(n)).astype(np.float32)
        offs = t_on + 0.05 + 1.95 * rng.random(n)
        ev = [(int(t * sr), ow.NOTE_ON, int(nt), float(v)) for t, nt, v in zip(t_on, notes, vels)]
        ev += [(int(t * sr), ow.NOTE_OFF, int(nt), 0.0) for t, nt in zip(offs, notes) if t < dur]
        ev += [(int(p * sr), ow.SUSTAIN, int((k + 1) % 2), 0.0) for k, p in enumerate(np.arange(3.0, dur, 3.0))]
        ev.sort(key=lambda e: (e[0], e[1]))
        return ev

    n_eng, dur, sr = 16384, 10.0, 96000.0
    t0 = time.perf_counter()
    evs = [stream_events(k + 1, dur, sr) for k in range(n_eng)]
    ej = [ow.engine_job(e, sample_rate=sr, duration=dur) for e in evs]
    t_build = time.perf_counter() - t0
    out = torch.empty((n_eng, int(dur * sr)), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ow.render_engines(ej, out=out)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    d = ow.last_diag()
    pick = [0, 4097, 9999, 16383]
    tc = time.perf_counter()
    ref = O.render_engines([O.engine_job(evs[k], sr=sr, dur=dur) for k in pick], threads=min(THREADS, len(pick)))
    tc = time.perf_counter() - tc
    got = out[pick].cpu().numpy()
    res["C5_full_16384_streams_10s_96k"] = {
        "streams": n_eng, "seconds_each": dur, "host_job_build_s": t_build, "gpu_s": t, "gpu_audio_s_per_s": n_eng * dur / t,
        "note_ons": int(d.nr_iter_hist[0]), "steals": int(d.nr_iter_hist[1]), "max_active_voices": int(d.nr_iter_hist[3]),
        "finite": bool(torch.isfinite(out).all().item()), "peak_abs": float(out.abs().max().item()),
        "parity_streams": pick, "max_abs_err_vs_oracle": float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()),
        "rel_l2_vs_oracle": float(np.sqrt(((got.astype(np.float64) - ref) ** 2).sum() / (ref.astype(np.float64) ** 2).sum())),
        "cpu_audio_s_per_s_on_the_sampled_streams": len(pick) * dur / tc, "cpu_threads": min(THREADS, len(pick))}
print(json.dumps(res, indent=1))
