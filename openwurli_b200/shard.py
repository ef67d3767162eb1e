"""Multi-GPU sharding of a render job list: renders are independent, so ranks take
disjoint slices and there is no data-path collective -- only a host-side gather of results/metrics.

Jobs are ordered so that instances sharing a preamp group (sample rate, LDR trajectory) stay
contiguous, then cut into `world_size` contiguous ranges balanced by rendered samples."""


def group_key(job):
    """Same key libowgpu uses to share DK matrices and the shadow solve (owgpu.cu, owg_plan_bench)."""
    v = job.v if hasattr(job, "v") else job
    trem = getattr(job, "tremolo_depth", 0.0)
    if trem > 0.0:
        return (v.sample_rate, 1, trem)
    return (v.sample_rate, 0, getattr(job, "r_ldr", 0.0))


def job_cost(job):
    v = job.v if hasattr(job, "v") else job
    x = v.duration_s * v.sample_rate
    return int(x) if x > 0 else 0


def shard_indices(jobs, world_size, rank):
    """Indices (into `jobs`) this rank renders. Deterministic on every rank; disjoint; covers all jobs."""
    assert 0 <= rank < world_size
    order = sorted(range(len(jobs)), key=lambda i: (group_key(jobs[i]), i))
    total = sum(job_cost(jobs[i]) for i in order)
    if total == 0:
        return [i for k, i in enumerate(order) if k % world_size == rank]
    bounds, acc, cut = [0], 0, 1
    for pos, i in enumerate(order):
        acc += job_cost(jobs[i])
        while cut < world_size and acc >= total * cut / world_size:
            bounds.append(pos + 1)
            cut += 1
    while len(bounds) < world_size + 1:
        bounds.append(len(order))
    bounds[-1] = len(order)
    return order[bounds[rank]:bounds[rank + 1]]
