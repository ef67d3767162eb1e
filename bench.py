#!/usr/bin/env python3
"""bench.py -- audio-seconds rendered per wall-second on the OpenWurli calibration grid (BASELINE.json C3).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--tremolo-depth D] [--duration S]

One "step" = one complete pass of the hot path over the 64-key x 127-velocity grid (8128 `preamp-bench render`
jobs, chain B, MLP on, 3 s per note, 44.1 kHz).  The default workload has tremolo depth 0.5, i.e. the full
north_star chain including the Twin-T/LDR coupling; the static-LDR CLI default is measured
in the same run and reported under "variants".  Under torchrun (N>1) two modes are measured in the same run:
  weak   : every rank renders its own full grid (seeds offset per rank) -- the machine-filling regime; top-level `value`
  strong : ONE grid, cut over the ranks with openwurli_b200.shard.shard_indices (the north-star split "8128 renders on 8 GPUs"),
           each rank copying its rows into ONE host buffer in shared memory (host-side gather, no collective) -- `strong_scaling`
--scaling strong makes the strong-scaled number the top-level `value`.  Renders are independent: no data-path collective.

`value`  : device-timed (CUDA events, max over ranks), init records resident in HBM, output left in HBM.
`e2e`    : the public API ow.render_bench(jobs, out=pinned_host): host note-on setup + H2D + kernels + D2H of
           every rendered sample, timed with the host clock around the call.
--impl reference times the CPU oracle (the reference's algorithm; the Rust reference cannot be built here) with
all host threads on a bounded sample of the same grid.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KEYS, VELS, FS = 64, 127, 44100.0
METRIC = "audio_seconds_rendered_per_second"
# ncu --set full summary of one chunk launch of the dominant kernel (tools/ncu_summary.py): roofline.traffic is read from it
TRAFFIC_ARTEFACT = "profiles/chain_kernel_traffic.json"
UNIT = "audio-s/s"


def grid_jobs(ow, duration, depth, seed_offset=0, stride=1):
    jobs = []
    k = 0
    for key in range(KEYS):
        for vel in range(1, VELS + 1):
            if k % stride == 0:
                midi = 33 + key
                jobs.append(ow.bench_job(note=midi, velocity=vel, duration=duration, tremolo_depth=depth,
                                         seed=(ow.default_noise_seed(midi) + seed_offset) & 0xFFFFFFFF))
            k += 1
    return jobs


def workload_name(duration, depth, model="melange12"):
    trem = f"tremolo_depth {depth:g}" if depth > 0 else "static LDR 1 MOhm (CLI default)"
    pre = "melange 12-node preamp" if model == "melange12" else "legacy 8-node preamp"
    return (f"C3 grid {KEYS} keys x {VELS} velocities = {KEYS * VELS} renders x {duration:g} s, chain B "
            f"(`preamp-bench render`), MLP on, 44.1 kHz 2x-oversampled {pre}, {trem}")


# ---- algorithmic FLOPs (DESIGN.md "Roofline") ---------------------------------------------------------
def algorithmic_flops(n_inst, n_samp, mean_nr_iters, depth, model=0, n_groups=1):
    """De-duplicated algorithmic FP64 operations of one pass (+,-,*,/,sqrt = 1 each; fast_exp = 18; libm = 20)."""
    if model == 1:
        # legacy 8-node step (dk_preamp_legacy.rs:447-554): 2 dense 8x8 mat-vec (256) + rhs/SM/K corrections (~60) + node update (56)
        # + final currents (2 exp, ~46) + per loop pass: 2 exp + residual/Jacobian/2x2 solve (~75); mean_nr_iters = loop passes
        dk_step = 256 + 60 + 56 + 46 + 75.0 * mean_nr_iters
        per_sample = 129 + 50 + 2 * dk_step + 200 + 62
        shared = n_groups * n_samp * 2 * dk_step
        if depth > 0:
            shared += n_groups * n_samp * 2 * 1060               # Twin-T/LDR step only: no matrix rebuild in this model
        return n_inst * n_samp * per_sample + shared
    dk_step = 85 + 288 + 72 + 70 + 220.0 * mean_nr_iters       # rhs + S*rhs + S_NI*i + checks + NR iterations
    per_sample = 129 + 50 + 2 * dk_step + 200 + 62               # voice + oversampler + 2 DK steps + power amp + speaker
    inst = n_inst * n_samp * per_sample
    shared = n_groups * n_samp * 2 * dk_step                     # the shadow solve, once per group
    if depth > 0:
        shared += n_groups * n_samp * 2 * (1060 + 5650)          # Twin-T/LDR step + 12x12 LU rebuild per preamp sample
    return inst + shared


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); smax.append(float(f[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_sample_jobs(O, duration, depth, n_jobs):
    """Bounded, evenly spread sample of the grid for the CPU arms."""
    total = KEYS * VELS
    jobs = []
    for s in range(n_jobs):
        k = (s * total) // n_jobs
        key, vel = k // VELS, k % VELS + 1
        jobs.append(O.bench_job(33 + key, vel, dur=duration, depth=depth))
    return jobs


def run_cpu(O, duration, depth, n_jobs, threads, model=0):
    import numpy as np
    jobs = cpu_sample_jobs(O, duration, depth, n_jobs)
    t0 = time.perf_counter()
    out = O.render_bench(jobs, threads=threads, preamp_model=model)
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(out))
    return n_jobs * duration / dt, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--duration", type=float, default=3.0)
    ap.add_argument("--tremolo-depth", type=float, default=0.5)
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grid-stride", type=int, default=1, help="debug: take every k-th grid job")
    ap.add_argument("--scaling", default="both", choices=["weak", "strong", "both"],
                    help="N>1: weak = a full grid per rank; strong = one grid split over the ranks; both = weak as `value`, strong under `strong_scaling`")
    ap.add_argument("--preamp-model", default="melange12", choices=["melange12", "legacy8"],
                    help="melange12 = the north-star 12-node DK preamp (cargo feature melange-preamp); legacy8 = the reference's default build")
    args = ap.parse_args()
    model = 1 if args.preamp_model == "legacy8" else 0

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = {"workload": workload_name(args.duration, args.tremolo_depth, args.preamp_model), "renders_per_gpu": KEYS * VELS // args.grid_stride,
           "sample_rate": FS, "duration_s": args.duration, "tremolo_depth": args.tremolo_depth, "preamp_model": args.preamp_model,
           "cache_policy": "outputs (8.6 GB per pass) and streamed matrices far exceed the 126 MB L2; no flush needed",
           "parallelism": f"instances sharded over {world} GPU(s), no collective"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        import oracle_lib as O
        threads = O.lib().owo_hardware_threads() or os.cpu_count() or 1
        # bounded sample: ~2 jobs per thread, full-length renders
        n_jobs = max(2 * threads, 8)
        O.lib().owo_preamp_settled((O.C.c_double * 19)())  # process-wide settled-state cache, like the reference's OnceLock
        for _ in range(args.warmup):
            run_cpu(O, min(args.duration, 0.1), args.tremolo_depth, threads, threads, model)
        t0 = time.perf_counter()
        vals = []
        for _ in range(args.steps):
            v, dt = run_cpu(O, args.duration, args.tremolo_depth, n_jobs, threads, model)
            vals.append(v)
        wall = time.perf_counter() - t0
        value = args.steps * n_jobs * args.duration / wall
        sample = f"{n_jobs} of {KEYS * VELS} grid renders per step (evenly spread keys/velocities), full {args.duration:g} s each"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": wall / max(args.steps, 1) * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "cpu_model": O.cpu_model(),
                                 "build": "g++ -O3 -march=x86-64-v3 -ffp-contract=off", "stage_ns_single_thread": O.stage_timers(40000)},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "CPU oracle = line-by-line C++ restatement of the reference (Rust toolchain absent; see DESIGN.md)"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------ our arm
    import numpy as np
    import torch
    import openwurli_b200 as ow

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.cuda.current_device()
    tstream = torch.cuda.Stream(device=dev)   # the plan launches on this stream and the timing events are recorded on it
    stream = tstream.cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_passes(depth, steps, warm, model=model):
        jobs = grid_jobs(ow, args.duration, depth, seed_offset=rank, stride=args.grid_stride)
        plan = ow.Plan.bench(jobs, device=dev, stream=stream, preamp_model=model)
        out = torch.empty((len(jobs), plan.max_samples), dtype=torch.float64, device="cuda")
        for _ in range(warm):
            plan.execute(out)
        barrier()
        sampler = ClockSampler(dev)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tstream)
        main_ms, launches = 0.0, 0
        for _ in range(steps):
            plan.execute(out)
            main_ms += plan.last_timing()[0]
            launches += plan.kernel_launches
        e1.record(tstream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        finite = bool(torch.isfinite(out[:: max(len(jobs) // 64, 1)]).all().item())
        peak = float(out.abs().max().item())
        n_inst, n_samp = len(jobs), plan.max_samples
        plan.close()
        del out
        return dict(ms=ms, main_ms=main_ms, launches=launches, clocks=clocks, finite=finite, peak=peak, n_inst=n_inst, n_samp=n_samp)

    def e2e_passes(depth, steps, model=model):
        jobs = grid_jobs(ow, args.duration, depth, seed_offset=rank, stride=args.grid_stride)
        n_samp = int(args.duration * FS)
        host = torch.empty((len(jobs), n_samp), dtype=torch.float64).pin_memory()
        ow.render_bench(jobs, out=host, device=dev, preamp_model=model)  # warm-up of the path (allocations, settled-state cache)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ts = time.perf_counter()
            ow.render_bench(jobs, out=host, device=dev, preamp_model=model)
            if os.environ.get("OWG_BENCH_DEBUG"):
                print(f"[rank {rank}] e2e step {time.perf_counter() - ts:.3f} s", file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        checksum = float(host[:, ::997].abs().sum().item())
        probe = ow.Plan.bench(jobs[:64], device=dev, preamp_model=model)  # init-record bytes per job, as the library counts its uploads
        h2d = int(probe.h2d_bytes / 64 * len(jobs))
        probe.close()
        d2h = len(jobs) * n_samp * 8
        del host
        return dict(s=t.item(), h2d=h2d, d2h=d2h, checksum=checksum, n_inst=len(jobs))

    def mean_nr_iterations(depth):
        """Mean Newton iterations per DK step on a 1-in-64 sample of the grid (device counters), for the FLOP model."""
        jobs = grid_jobs(ow, args.duration, depth, stride=64)
        out = torch.empty((len(jobs), int(args.duration * FS)), dtype=torch.float64, device="cuda")
        pl = ow.Plan.bench(jobs, device=dev, stream=stream, collect_diag=True, preamp_model=model)
        pl.execute(out)
        d = ow.last_diag()
        pl.close()
        h = np.array(list(d.nr_iter_hist), dtype=np.float64)
        if model == 1:  # legacy: bucket b = b Newton updates = b + 1 device evaluations in the loop
            return float((h * (np.arange(16) + 1)).sum() / max(h.sum(), 1.0))
        # bucket b = last_nr_iterations b -> b+1 iterations ran; bucket 15 (>=15) counted as 16 (lower bound)
        return float((h * (np.arange(16) + 1)).sum() / max(h.sum(), 1.0))

    def strong_passes(depth, steps, warm):
        """ONE grid split over the ranks (shard.shard_indices): device-timed max over ranks, then the same through the public API with
        every rank writing its rows into one host buffer in POSIX shared memory (the host-side gather)."""
        from openwurli_b200 import shard
        jobs_all = grid_jobs(ow, args.duration, depth, stride=args.grid_stride)     # identical on every rank
        idx = shard.shard_indices(jobs_all, world, rank)
        mine = [jobs_all[i] for i in idx]
        plan = ow.Plan.bench(mine, device=dev, stream=stream, preamp_model=model)
        out = torch.empty((max(len(mine), 1), plan.max_samples), dtype=torch.float64, device="cuda")
        for _ in range(warm):
            plan.execute(out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tstream)
        for _ in range(steps):
            plan.execute(out)
        e1.record(tstream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        plan.close()
        del out
        # ---- end to end: one result buffer on the host, every rank fills its slab
        n_samp = int(args.duration * FS)
        n_all = len(jobs_all)
        gather = "one POSIX shared-memory buffer, each rank's D2H copy lands in its row range"
        shm = None
        host = None
        contiguous = len(idx) > 0 and idx == list(range(idx[0], idx[0] + len(idx)))
        try:
            from multiprocessing import shared_memory
            if not contiguous:
                raise RuntimeError("shard is not a contiguous row range")
            name = "owg_bench_" + os.environ.get("MASTER_PORT", "0")
            nbytes = n_all * n_samp * 8
            if rank == 0:
                free = os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize
                ok = torch.tensor([1 if free > nbytes + (1 << 28) else 0], device="cuda")
                if ok.item():
                    shm = shared_memory.SharedMemory(name=name, create=True, size=nbytes)
            else:
                ok = torch.tensor([0], device="cuda")
            dist.broadcast(ok, 0)
            if not ok.item():
                raise RuntimeError("/dev/shm too small")
            dist.barrier()
            if rank != 0:
                shm = shared_memory.SharedMemory(name=name)
            full = np.ndarray((n_all, n_samp), dtype=np.float64, buffer=shm.buf)
            host = torch.from_numpy(full[idx[0]:idx[0] + len(idx)])
            rc = torch.cuda.cudart().cudaHostRegister(host.data_ptr(), host.numel() * 8, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed ({rc})")
        except Exception as ex:  # fall back to private pinned slabs (still one box, no inter-process copy either way)
            gather = f"per-rank pinned slabs ({ex})"
            host = torch.empty((max(len(mine), 1), n_samp), dtype=torch.float64).pin_memory()
        ow.render_bench(mine, out=host, device=dev, preamp_model=model)
        barrier()
        t0 = time.perf_counter()
        e_steps = max(1, min(steps, 2))
        for _ in range(e_steps):
            ow.render_bench(mine, out=host, device=dev, preamp_model=model)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        checksum = torch.tensor([float(host[:, ::997].abs().sum().item())], dtype=torch.float64, device="cuda")
        dist.all_reduce(checksum, op=dist.ReduceOp.SUM)
        barrier()
        if shm is not None:
            try:
                torch.cuda.cudart().cudaHostUnregister(host.data_ptr())
            except Exception:
                pass
            del host, full
            shm.close()
            if rank == 0:
                shm.unlink()
        audio = n_all * args.duration
        return {"scaling": "strong", "value": audio * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "renders_total": n_all,
                "renders_per_gpu": [len(idx)], "e2e": {"value": audio * e_steps / t.item(), "unit": UNIT, "d2h_bytes_per_step": n_all * n_samp * 8,
                                                        "gather": gather, "checksum": checksum.item()},
                "note": "one C3 grid split over the ranks by instance (shard.shard_indices); a render is a sample-serial recurrence, so a grid "
                        "that already under-fills one GPU gains little from more GPUs: the per-step latency of a render, not throughput, bounds it"}

    main_run = timed_passes(args.tremolo_depth, args.steps, warmup)
    audio_s = main_run["n_inst"] * args.duration * world
    value = audio_s * args.steps / (main_run["ms"] * 1e-3)
    e2e = e2e_passes(args.tremolo_depth, max(1, min(args.steps, 2)))
    e2e_value = e2e["n_inst"] * args.duration * world * max(1, min(args.steps, 2)) / e2e["s"]

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": main_run["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"]},
            "gpu_launches": main_run["launches"], "clocks": main_run["clocks"],
            "checks": {"finite": main_run["finite"], "peak_abs": main_run["peak"], "e2e_checksum": e2e["checksum"]}}

    if world > 1 and args.scaling in ("strong", "both"):
        st = strong_passes(args.tremolo_depth, args.steps, warmup)
        if args.scaling == "strong":
            line.update(value=st["value"], ms_per_step=st["ms_per_step"], scaling="strong", e2e={"value": st["e2e"]["value"], "unit": UNIT,
                        "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": st["e2e"]["d2h_bytes_per_step"]})
            line["weak_scaling"] = {"value": value, "unit": UNIT, "ms_per_step": main_run["ms"] / args.steps, "e2e": e2e_value}
            line["config"]["renders_per_gpu"] = KEYS * VELS // args.grid_stride // world
        line["strong_scaling"] = st
    if rank == 0:
        # roofline of the dominant kernel (chain_kernel): algorithmic FLOPs / its CUDA-event duration inside the timed region
        fma_peak = ow.fp64_peak(device=dev, fma=True) * 2.0     # TFLOP/s, DFMA = 2 flop
        unfused_peak = ow.fp64_peak(device=dev, fma=False)      # T instr/s = TFLOP/s for uncontracted code
        iters = mean_nr_iterations(args.tremolo_depth)
        flops = algorithmic_flops(main_run["n_inst"], main_run["n_samp"], iters, args.tremolo_depth, model)
        kernel_s = main_run["main_ms"] * 1e-3 / args.steps
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, TRAFFIC_ARTEFACT)))
        except (OSError, ValueError):
            pass
        traffic_applies = bool(traffic.get("dram_bytes")) and model == 0 and main_run["n_inst"] == 8128 and args.tremolo_depth > 0
        achieved = flops / kernel_s / 1e12
        line["roofline"] = {"bound": "fp64", "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak,
                            "traffic": traffic["dram_bytes"] if traffic_applies else None,
                            "traffic_note": (f"dram__bytes_read.sum + dram__bytes_write.sum of one chunk launch of the dominant kernel, read from "
                                             f"{TRAFFIC_ARTEFACT} (ncu --set full capture: {traffic.get('note', '')}); algorithmic bytes of that launch: "
                                             f"{traffic.get('algorithmic_bytes')}") if traffic_applies else "no ncu capture for this configuration",
                            "peak_source": "owg_fp64_peak DFMA micro-benchmark measured in this run "
                            "(MEASURED_PEAKS.json has no FP64 entry; B200 nominal 37 TFLOP/s)",
                            "peak_unfused_tflops": unfused_peak, "frac_of_unfused": achieved / unfused_peak,
                            "kernel": "owgd::chain_legacy_kernel" if model == 1 else
                            ("owgd::chain_tile_kernel" if main_run["n_inst"] <= 2 * 2 * 148 * 4 * 7 else "owgd::chain_kernel"), "kernel_ms_per_step": kernel_s * 1e3, "mean_nr_iterations": iters,
                            "algorithmic_gflop_per_step": flops / 1e9}
        if not args.no_variants:
            other = 0.0 if args.tremolo_depth > 0 else 0.5
            v = timed_passes(other, max(1, args.steps - 1), 1) if world == 1 else None
            if v:
                line["variants"] = {workload_name(args.duration, other, args.preamp_model): {
                    "value": v["n_inst"] * args.duration * max(1, args.steps - 1) / (v["ms"] * 1e-3), "unit": UNIT,
                    "ms_per_step": v["ms"] / max(1, args.steps - 1)}}
                # machine-filling batch (throughput regime): 4 grids at different volumes, 0.5 s each (the C4 sweep's render length)
                big = []
                for g in range(4):
                    for j in grid_jobs(ow, 0.5, args.tremolo_depth, seed_offset=g):
                        j.volume = 0.3 + 0.15 * g
                        big.append(j)
                plan = ow.Plan.bench(big, device=dev, stream=stream)
                outb = torch.empty((len(big), plan.max_samples), dtype=torch.float64, device="cuda")
                plan.execute(outb)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(tstream); plan.execute(outb); e1.record(tstream); torch.cuda.synchronize()
                msb = e0.elapsed_time(e1)
                line["variants"][f"throughput regime: {len(big)} renders x 0.5 s (4 grids), tremolo_depth {args.tremolo_depth:g}"] = {
                    "value": len(big) * 0.5 / (msb * 1e-3), "unit": UNIT, "ms_per_step": msb}
                plan.close()
                del outb
                if model == 0:  # the reference's default-build preamp (legacy 8-node) on the same grid
                    lv = timed_passes(args.tremolo_depth, 1, 1, model=1)
                    le = e2e_passes(args.tremolo_depth, 1, model=1)
                    entry = {"value": lv["n_inst"] * args.duration / (lv["ms"] * 1e-3), "unit": UNIT, "ms_per_step": lv["ms"],
                             "e2e": le["n_inst"] * args.duration / le["s"]}
                    if not args.no_cpu_baseline:
                        import oracle_lib as O
                        threads = O.lib().owo_hardware_threads() or os.cpu_count() or 1
                        vc, dtc = run_cpu(O, args.duration, args.tremolo_depth, max(2 * threads, 8), threads, 1)
                        entry["cpu_baseline"] = {"value": vc, "unit": UNIT, "cores": threads, "kind": "port"}
                    line["variants"][f"reference default build: {workload_name(args.duration, args.tremolo_depth, 'legacy8')}"] = entry
        if not args.no_variants and world == 1 and model == 0:
            # chain B with the melange 7-BJT power amplifier (`--no-default-features`): every 4th grid job x 0.5 s, a bounded
            # slice (the amplifier costs far more per sample than the rest of the chain: pa_melange_kernel)
            pj = grid_jobs(ow, 0.5, args.tremolo_depth)[::4]
            outp = torch.empty((len(pj), int(0.5 * 44100.0)), dtype=torch.float64, device="cuda")
            ow.render_bench(pj[:16], out=outp[:16], power_amp_model=ow.PA_MELANGE)   # settled amplifier state, module load
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ow.render_bench(pj, out=outp, power_amp_model=ow.PA_MELANGE)
            torch.cuda.synchronize()
            dtp = time.perf_counter() - t0
            line.setdefault("variants", {})[f"melange power amplifier: {len(pj)} grid renders x 0.5 s, chain B with power_amp_model = melange (rail sag on)"] = {
                "value": len(pj) * 0.5 / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3, "finite": bool(torch.isfinite(outp).all().item()),
                "timing": "host clock around one call with device-resident output (includes the plan's H2D of the job records)"}
            del outp
        if not args.no_cpu_baseline:
            import oracle_lib as O
            threads = O.lib().owo_hardware_threads() or os.cpu_count() or 1
            n_jobs = max(2 * threads, 8)
            O.lib().owo_preamp_settled((O.C.c_double * 19)())
            v, dt = run_cpu(O, args.duration, args.tremolo_depth, n_jobs, threads, model)
            v1, dt1 = run_cpu(O, args.duration, args.tremolo_depth, 1, 1, model)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_jobs} evenly spread grid renders x {args.duration:g} s, {threads} threads "
                                              f"({dt:.1f} s); single-thread: {v1:.2f} audio-s/s",
                                    "single_thread": v1, "cpu_model": O.cpu_model(), "build": "g++ -O3 -march=x86-64-v3 -ffp-contract=off",
                                    "stage_ns_single_thread": O.stage_timers(40000),
                                    "reference_published_single_thread": {"melange_tremolo": 4.2, "melange_static": 38.0, "legacy_tremolo": 29.0,
                                                                          "source": "CHANGELOG.md:105-111, 30 s renders incl. one-time setup, hardware unstated"}}
        print(json.dumps(line))
    elif not args.no_variants and world == 1:
        pass
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
