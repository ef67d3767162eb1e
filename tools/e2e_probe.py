#!/usr/bin/env python3
"""End-to-end breakdown of one-shot chain-B renders of the C3 grid on the current device: plan (host parameterisation + uploads +
constructor launch) / execute (device work + overlapped copy-back).  Usage: e2e_probe.py [depth] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openwurli_b200 as ow
depth = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tag = os.environ.get("CUDA_VISIBLE_DEVICES", "?")
dist = None
if "RANK" in os.environ:  # under torchrun: one rank per GPU, reps aligned by a barrier like bench.py's e2e leg
    import torch.distributed as dist
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    tag = f"rank{lr}"
def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
jobs = [ow.bench_job(note=33 + k // 127, velocity=1 + k % 127, duration=3.0, tremolo_depth=depth) for k in range(8128)]
t0 = time.perf_counter()
host = torch.empty((len(jobs), 132300), dtype=torch.float64).pin_memory()
print(f"[gpu {tag}] pin 8.6 GB: {time.perf_counter()-t0:.2f} s", flush=True)
ow.render_bench(jobs[:64], out=host[:64])
for r in range(reps):
    barrier(); t0 = time.perf_counter()
    pl = ow.Plan.bench(jobs)
    t1 = time.perf_counter()
    pl.execute(host)
    t2 = time.perf_counter()
    dev_ms = pl.last_timing()
    pl.close()
    t3 = time.perf_counter()
    tm = None
    print(f"[gpu {tag}] rep {r}: plan {t1-t0:.3f} s, execute(host out) {t2-t1:.3f} s, close {t3-t2:.3f} s, total {t3-t0:.3f} -> {8128*3/(t3-t0):.0f} audio-s/s; device chain/total ms {dev_ms}", flush=True)
    barrier(); t0 = time.perf_counter()
    ow.render_bench(jobs, out=host)
    print(f"[gpu {tag}] rep {r}: one-shot render_bench {time.perf_counter()-t0:.3f} s", flush=True)
