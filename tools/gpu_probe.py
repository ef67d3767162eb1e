#!/usr/bin/env python3
"""Ad-hoc GPU bring-up probe: parity of libowgpu against the CPU oracle + first timings. Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import openwurli_b200 as ow
import oracle_lib as O


def cmp(name, a, b):
    d = np.abs(a - b)
    nz = int((a != b).sum())
    l2 = np.sqrt((d ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-300)
    first = int(np.argmax(a != b)) if nz else -1
    print(f"{name}: max_abs={d.max():.3e} rel_l2={l2:.3e} mismatching={nz}/{a.size} first_mismatch={first} peak={np.abs(b).max():.4f}", flush=True)


print("devices", ow.device_count(), flush=True)
t = time.time(); print("fp64 peak fma", ow.fp64_peak(fma=True), "T instr/s; nofma", ow.fp64_peak(fma=False), time.time() - t, flush=True)
# chain V
jobs = [ow.voice_job(m, v, duration=1.0) for m, v in [(60, 100), (33, 1), (96, 127), (72, 64), (45, 30)]]
t = time.time(); g = ow.render_voices(jobs); print("gpu voices", time.time() - t)
c = O.render_voices([O.voice_job(j.midi, 0, dur=1.0, vel_norm=j.velocity) for j in jobs])
for i in range(len(jobs)):
    cmp(f"voice[{jobs[i].midi},{jobs[i].velocity:.3f}]", g[i], c[i])
# chain B depth 0
bj = [ow.bench_job(note=m, velocity=v, duration=0.5) for m, v in [(60, 100), (40, 127), (84, 20)]]
t = time.time(); g = ow.render_bench(bj, collect_diag=True); print("gpu bench d0", time.time() - t)
d = ow.last_diag(); print("gpu main hist", list(d.nr_iter_hist), "be", d.be_fallback, "launches", d.kernels_launched)
c = O.render_bench([O.bench_job(j.v.midi, 0, dur=0.5) for j in bj]) if False else None
oj = []
for j in bj:
    o = O.bench_job(j.v.midi, 100, dur=0.5); o.v.velocity = j.v.velocity; oj.append(o)
c = O.render_bench(oj)
d = O.last_diag(); print("cpu main hist", list(d.nr_iter_hist), "be", d.be_fallback)
for i in range(len(bj)):
    cmp(f"benchB d0 [{bj[i].v.midi}]", g[i], c[i])
# chain B depth 0.5
bj = [ow.bench_job(note=m, velocity=v, duration=0.25, tremolo_depth=0.5) for m, v in [(60, 100), (48, 90)]]
t = time.time(); g = ow.render_bench(bj, collect_diag=True); print("gpu bench d0.5", time.time() - t)
d = ow.last_diag(); print("gpu main hist", list(d.nr_iter_hist), "trem", list(d.tremolo_nr_iter_hist))
oj = []
for j in bj:
    o = O.bench_job(j.v.midi, 100, dur=0.25, depth=0.5); o.v.velocity = j.v.velocity; oj.append(o)
c = O.render_bench(oj)
d = O.last_diag(); print("cpu main hist", list(d.nr_iter_hist), "trem", list(d.tremolo_nr_iter_hist))
for i in range(len(bj)):
    cmp(f"benchB d0.5 [{bj[i].v.midi}]", g[i], c[i])
# timing: grid slice
import torch
for nkeys, nvel, dur in [(64, 127, 0.25)]:
    jobs = [ow.bench_job(note=33 + k, velocity=1 + v, duration=dur) for k in range(nkeys) for v in range(nvel)]
    t = time.time(); pl = ow.Plan.bench(jobs); print("plan", len(jobs), time.time() - t, flush=True)
    out = torch.empty((len(jobs), pl.max_samples), dtype=torch.float64, device="cuda")
    for rep in range(2):
        t = time.time(); pl.execute(out); torch.cuda.synchronize(); dt = time.time() - t
        print(f"grid {len(jobs)} x {dur}s: {dt:.3f}s -> {len(jobs)*dur/dt:.0f} audio-s/s; kernel ms {pl.last_timing()}", flush=True)
