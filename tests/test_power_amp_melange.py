"""SURVEY 8(f) #4, device side: the melange 7-BJT power amplifier with rail sag on the GPU (openwurli_b200/csrc/owg_pa_core.h,
owg_poweramp.cuh, host_pa_setup.cpp) against the oracle (the mechanically transliterated gen_power_amp.rs under the hand-restated
adapter of power_amp.rs:279-465).

CPU tests (no GPU): the host model builder against the reference's re-rated matrices, bit for bit; the PRODUCT's lane-tiled solver source
run on a 16-coroutine lane emulator (tests/pa_tile_emu.cpp) against the oracle -- the same source the CUDA kernel compiles.
GPU tests: owg_power_amp_batch and chain B with power_amp_model = PA_MELANGE through the C ABI against the oracle.

Tolerance: the solver's arithmetic is IEEE f64 in the reference's order (bit-identical on the emulator); on the device only pnjlim's
logarithm is CUDA's instead of glibc's, so the GPU bound is the north star's (max-abs <= 1e-6 full scale, relative L2 <= 1e-7) and the
discrete events (divergence-guard resets, backward-Euler retries) must be equal."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

L = O.lib()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_solver = pytest.mark.skipif(not L.owo_have_melange_power_amp(), reason="oracle built without oracle/_ref/gen_power_amp.hpp (reference sources absent)")
dp = C.POINTER(C.c_double)


def _oracle_amp(x, sr, rail_sag=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y, r = np.zeros(len(x)), np.zeros((max(len(x), 1), 2))
    resets = L.owo_power_amp_melange(sr, 1 if rail_sag else 0, O.dptr(x), len(x), O.dptr(y), O.dptr(r), -1)
    assert resets >= 0
    return y, r[len(x) - 1] if len(x) else np.array([22.5, 22.5]), resets


def _signals(sr, n):
    t = np.arange(n) / sr
    rng = np.random.default_rng(7)
    return [np.zeros(n), 0.01 * np.sin(2 * np.pi * 440 * t), 0.1 * np.sin(2 * np.pi * 220 * t), 0.5 * np.sin(2 * np.pi * 1000 * t),
            3.0 * np.sin(2 * np.pi * 100 * t), 0.05 * rng.standard_normal(n), np.where(t > t[n // 2], 0.3, -0.2),
            0.2 * np.sin(2 * np.pi * 3000 * t) * np.exp(-t * 200)]


# ---- CPU ---------------------------------------------------------------------------------------------------------------------------------
def _emu():
    so = os.path.join(ROOT, "tests", "_build", "libpaemu.so")
    srcs = [os.path.join(ROOT, "tests", "pa_tile_emu.cpp"), os.path.join(ROOT, "openwurli_b200", "csrc", "host_pa_setup.cpp")]
    deps = srcs + [os.path.join(ROOT, "openwurli_b200", "csrc", f) for f in ("owg_pa_core.h", "ow_consts_pa.inc", "host_setup.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so] + srcs,
                       check=True)
    E = C.CDLL(so)
    E.paemu_render.argtypes = [C.c_double, C.c_int, dp, C.c_int64, C.c_double, dp, dp, C.POINTER(C.c_uint32)]
    E.paemu_model.argtypes = [C.c_double, dp]
    E.paemu_settle.argtypes = [C.c_int64, dp]
    E.paemu_set_settled.argtypes = [dp]
    E.paemu_render_pair.argtypes = [C.c_double, C.c_int, dp, C.c_int64, dp, C.c_int64, dp, dp, dp, C.POINTER(C.c_uint32)]
    return E


@needs_solver
@pytest.mark.parametrize("sr", [88200.0, 44100.0, 48000.0, 96000.0])
def test_host_model_equals_the_reference_matrices(sr):
    """pa_build_model: the baked tables at 88.2 kHz, set_sample_rate's rebuild (three 20x20 LU inversions, K = N_V S N_I, S N_I, DC-blocker pole)
    elsewhere -- every entry bit-identical to CircuitState after PowerAmp::new_at_sample_rate(sr)."""
    mine, ref = np.zeros(2353), np.zeros(2353)
    assert _emu().paemu_model(sr, mine.ctypes.data_as(dp)) == 0
    assert L.owo_power_amp_matrices(sr, O.dptr(ref)) == 0
    assert np.array_equal(mine.view(np.uint64), ref.view(np.uint64))
    assert np.abs(ref[:400]).max() > 0


@needs_solver
def test_lane_tiled_solver_source_is_bit_identical_to_the_oracle_on_the_lane_emulator():
    """The product's 16-lane algorithm (row-per-lane elimination with shuffle pivot search, per-lane junction evaluation, vote-based
    convergence, backward-Euler retry, divergence guard, rail sag) executed lane by lane on the CPU: every output sample, the final rail
    voltages and the reset count equal the oracle's, bit for bit -- including inputs hot enough to exhaust the 70 Newton iterations.
    The emulator is slow (a context switch per lane per collective), so the settled state is checked over the first 250 samples of
    CircuitState::default()'s trajectory and the renders then start from the oracle's fully settled state."""
    E = _emu()
    mine, ref = np.zeros(54), np.zeros(54)
    assert E.paemu_settle(200, mine.ctypes.data_as(dp)) == 0 and L.owo_power_amp_settled(200, O.dptr(ref)) == 0
    assert np.array_equal(mine.view(np.uint64), ref.view(np.uint64))
    assert L.owo_power_amp_settled(44100, O.dptr(ref)) == 0 and E.paemu_set_settled(ref.ctypes.data_as(dp)) == 0
    n = 40
    cptr = C.POINTER(C.c_uint32)
    total_resets = 0
    for sr, picks in ((44100.0, range(8)), (88200.0, (1, 3, 5))):
        sig = _signals(sr, n)
        for k in picks:
            x = np.ascontiguousarray(sig[k])
            y, rails, cnt = np.zeros(n), np.zeros(2), np.zeros(4, np.uint32)
            assert E.paemu_render(sr, 1, x.ctypes.data_as(dp), n, 1.0, y.ctypes.data_as(dp), rails.ctypes.data_as(dp), cnt.ctypes.data_as(cptr)) == 0
            yo, ro, resets = _oracle_amp(x, sr)
            assert np.array_equal(y.view(np.uint64), yo.view(np.uint64)), (sr, k, float(np.abs(y - yo).max()))
            assert np.array_equal(rails, ro) and int(cnt[0]) == resets, (sr, k)
            total_resets += resets
    assert total_resets > 0
    # non-finite input samples (process_sample treats them as 0, gen_power_amp.rs:8839), inputs beyond the +-100 V clamp, a DC step
    x = np.ascontiguousarray(np.array([0.0, 0.01, np.nan, 0.02, np.inf, -np.inf, 150.0, -150.0, 0.5, 0.5, 0.5, -0.5, 0.0, 0.0, 1e-300, -0.0] * 2))
    y, rails, cnt = np.zeros(len(x)), np.zeros(2), np.zeros(4, np.uint32)
    assert E.paemu_render(44100.0, 1, x.ctypes.data_as(dp), len(x), 1.0, y.ctypes.data_as(dp), rails.ctypes.data_as(dp), cnt.ctypes.data_as(cptr)) == 0
    yo, ro, resets = _oracle_amp(x, 44100.0)
    assert np.array_equal(y.view(np.uint64), yo.view(np.uint64)) and np.array_equal(rails, ro) and int(cnt[0]) == resets and np.isfinite(y).all()
    x = np.ascontiguousarray(_signals(44100.0, n)[3])
    y, rails, cnt = np.zeros(n), np.zeros(2), np.zeros(4, np.uint32)
    E.paemu_render(44100.0, 0, x.ctypes.data_as(dp), n, 1.0, y.ctypes.data_as(dp), rails.ctypes.data_as(dp), cnt.ctypes.data_as(cptr))
    yo, _, resets = _oracle_amp(x, 44100.0, rail_sag=False)
    assert np.array_equal(y, yo) and int(cnt[0]) == resets and tuple(rails) == (22.5, 22.5)   # ideal rails


@needs_solver
def test_two_tiles_of_a_warp_in_lock_step_on_the_lane_emulator():
    """The GPU runs two instances per warp with full-warp collectives: a tile that has converged, ended, or does not need the
    backward-Euler retry keeps executing its neighbour's collectives and commits nothing.  Here 32 coroutines emulate that warp (every
    collective is a barrier over all 32 lanes): pairs of rows of different difficulty AND length -- silence beside a row that exhausts
    both Newton loops and resets, a long row beside a short one -- each still bit-identical to the oracle."""
    E = _emu()
    ref = np.zeros(54)
    assert L.owo_power_amp_settled(44100, O.dptr(ref)) == 0 and E.paemu_set_settled(ref.ctypes.data_as(dp)) == 0
    sr = 44100.0
    sig = _signals(sr, 36)
    total_resets = 0
    for a, na, b, nb in ((0, 36, 3, 36), (3, 20, 1, 36), (4, 36, 6, 7), (2, 36, 2, 36)):
        x0, x1 = np.ascontiguousarray(sig[a][:na]), np.ascontiguousarray(sig[b][:nb])
        y0, y1, rails, cnt = np.zeros(na), np.zeros(nb), np.zeros(4), np.zeros(8, np.uint32)
        assert E.paemu_render_pair(sr, 1, x0.ctypes.data_as(dp), na, x1.ctypes.data_as(dp), nb, y0.ctypes.data_as(dp), y1.ctypes.data_as(dp),
                                   rails.ctypes.data_as(dp), cnt.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
        for x, y, r, c in ((x0, y0, rails[:2], cnt[:4]), (x1, y1, rails[2:], cnt[4:])):
            yo, ro, resets = _oracle_amp(x, sr)
            assert np.array_equal(y.view(np.uint64), yo.view(np.uint64)), (a, b, float(np.abs(y - yo).max()))
            assert np.array_equal(r, ro) and int(c[0]) == resets
            total_resets += resets
    assert total_resets > 0


def test_power_amp_entry_point_refuses_without_a_device_and_rejects_bad_arguments():
    import openwurli_b200 as ow
    lib = ow.lib()
    x = np.zeros((2, 8))
    o = ow.api._opts()
    p = x.ctypes.data_as(C.c_void_p)
    assert lib.owg_power_amp_batch(p, 4, 2, 8, 44100.0, 1, p, 8, None, None, C.byref(o)) == -1      # stride < n_samp
    assert lib.owg_power_amp_batch(p, 8, 2, 8, -1.0, 1, p, 8, None, None, C.byref(o)) == -1         # sample rate
    assert lib.owg_power_amp_batch(p, 8, 0, 8, 44100.0, 1, p, 8, None, None, C.byref(o)) == 0       # empty batch
    if ow.device_count() == 0:
        assert lib.owg_power_amp_batch(p, 8, 2, 8, 44100.0, 1, p, 8, None, None, C.byref(o)) == -2  # no CPU fallback


# ---- GPU ---------------------------------------------------------------------------------------------------------------------------------
def _close(g, c):
    err = float(np.abs(g - c).max())
    rel = float(np.linalg.norm(g - c) / max(np.linalg.norm(c), 1e-300))
    return err, rel


@pytest.mark.gpu
@needs_solver
@pytest.mark.parametrize("sr,rail_sag", [(44100.0, True), (88200.0, True), (96000.0, False)])
def test_gpu_power_amp_batch_matches_the_oracle(sr, rail_sag):
    """owg_power_amp_batch through the C ABI: silence, small and large sines, noise, a step, a decaying burst -- quiet rows inside the north
    star's bound, every row's divergence-guard reset count and final rail voltages equal to the oracle's."""
    import openwurli_b200 as ow
    n = 2000
    x = np.ascontiguousarray(np.stack(_signals(sr, n)))
    y, rails, cnt = ow.power_amp_batch(x, sr, rail_sag=rail_sag, want_state=True)
    total_resets = 0
    for i in range(x.shape[0]):
        yo, ro, resets = _oracle_amp(x[i], sr, rail_sag)
        err, rel = _close(y[i], yo)
        assert err <= 1e-6 and (rel <= 1e-7 or np.linalg.norm(yo) == 0), (i, err, rel)
        assert int(cnt[i, 0]) == resets, (i, cnt[i], resets)
        assert np.abs(rails[i] - ro).max() <= 1e-9, (i, rails[i], ro)
        total_resets += resets
    assert total_resets > 0      # the guard path ran on the device
    assert np.isfinite(y).all() and np.abs(y).max() <= 1.0


@pytest.mark.gpu
@needs_solver
def test_gpu_power_amp_batch_rows_are_independent_and_ragged_batches_work():
    """9 and 17 rows (a partial CTA, a lone tile in a warp): every row equals the same row rendered alone; device-resident buffers work."""
    import torch
    import openwurli_b200 as ow
    sr, n = 44100.0, 400
    base = _signals(sr, n)
    x = np.ascontiguousarray(np.stack([base[i % len(base)] * (1.0 + 0.01 * i) for i in range(17)]))
    y = ow.power_amp_batch(x, sr)
    y9 = ow.power_amp_batch(np.ascontiguousarray(x[:9]), sr)
    assert np.array_equal(y[:9], y9)
    for i in (0, 8, 16):
        assert np.array_equal(ow.power_amp_batch(np.ascontiguousarray(x[i:i + 1]), sr)[0], y[i])
    xd = torch.from_numpy(x).cuda()
    yd = ow.power_amp_batch(xd, sr)
    assert np.array_equal(yd.cpu().numpy(), y)


@pytest.mark.gpu
@needs_solver
def test_gpu_chain_b_with_the_melange_power_amp_matches_the_oracle():
    """`preamp-bench render` built with --no-default-features (chain B, main.rs:478-496): voice -> melange preamp -> volume^2 -> melange
    PowerAmp::new() -> speaker.  Oracle: its chain-B preamp tap through owo_output_stage_melange.  Rows with --no-poweramp, --no-rail-sag,
    tremolo, and different volumes / speaker characters share one call."""
    import openwurli_b200 as ow
    jobs = [dict(note=60, velocity=100, duration=0.25), dict(note=45, velocity=127, duration=0.3, volume=0.9, speaker=0.3),
            dict(note=72, velocity=40, duration=0.2, tremolo_depth=0.5), dict(note=84, velocity=110, duration=0.2, no_poweramp=True),
            dict(note=38, velocity=127, duration=0.25, volume=1.0, speaker=0.0)]
    okw = dict(note="midi", velocity="vel", duration="dur", tremolo_depth="depth")
    for model, sag in ((ow.PA_MELANGE, 1), (ow.PA_MELANGE_IDEAL_RAILS, 0)):
        g = ow.render_bench([ow.bench_job(**j) for j in jobs], power_amp_model=model)
        for i, j in enumerate(jobs):
            oj = O.bench_job(**{okw.get(k, k): v for k, v in j.items()})
            pre = np.ascontiguousarray(O.render_bench_taps(oj)["preamp"])
            ref = np.zeros(len(pre))
            assert L.owo_output_stage_melange(O.dptr(pre), len(pre), 44100.0, oj.volume, oj.speaker_character, oj.no_poweramp, sag, O.dptr(ref)) == 0
            err, rel = _close(g[i, :len(ref)], ref)
            assert err <= 1e-6 and rel <= 1e-7, (model, i, err, rel)
    # the behavioural amplifier is still the default, and the melange model is refused where it is not served
    b = ow.render_bench([ow.bench_job(**jobs[0])])
    assert float(np.abs(b - ow.render_bench([ow.bench_job(**jobs[0])], power_amp_model=ow.PA_MELANGE)).max()) > 1e-4
    o = ow.api._opts(power_amp_model=ow.PA_MELANGE)
    x = np.zeros((1, 64))
    assert ow.lib().owg_preamp_batch(x.ctypes.data_as(C.c_void_p), 64, 1, 64, 44100.0, 1, 0.0, 1e6, x.ctypes.data_as(C.c_void_p), 64, C.byref(o)) == -5


@pytest.mark.gpu
@needs_solver
def test_gpu_chain_batch_routes_caller_rows_through_the_melange_power_amp():
    """owg_chain_batch (render-poly's mix / per-voice chains, main.rs:1470-1481) with power_amp_model = PA_MELANGE: caller rows, preamp
    bypassed so that the rows are the amplifier's input -> volume^2 -> melange amplifier -> speaker, against owo_output_stage_melange."""
    import openwurli_b200 as ow
    sr, n = 44100.0, 1500
    sig = _signals(sr, n)
    x = np.ascontiguousarray(np.stack([sig[1], sig[2] * 4.0, sig[5] * 3.0, sig[7]]))
    vols, chars = (0.6, 0.9, 0.5, 1.0), (1.0, 0.0, 0.4, 0.7)
    params = [ow.bench_job(volume=v, speaker=c, no_preamp=True, duration=n / sr) for v, c in zip(vols, chars)]
    g = ow.chain_batch(x, params, power_amp_model=ow.PA_MELANGE)
    for i in range(len(params)):
        ref = np.zeros(n)
        assert L.owo_output_stage_melange(O.dptr(np.ascontiguousarray(x[i])), n, sr, vols[i], chars[i], 0, 1, O.dptr(ref)) == 0
        err, rel = _close(g[i], ref)
        assert err <= 1e-6 and rel <= 1e-7, (i, err, rel)


@pytest.mark.gpu
def test_gpu_release_caches_returns_the_staging_buffer_and_keeps_the_current_device():
    """owg_release_caches: after a host-output render the per-device staging buffer is released (bytes > 0), the calling thread's CUDA
    device is untouched, and the next render re-allocates and gives the same samples."""
    import torch
    import openwurli_b200 as ow
    jobs = [ow.bench_job(note=60, velocity=90, duration=0.05), ow.bench_job(note=64, velocity=60, duration=0.05)]
    a = ow.render_bench(jobs)
    before = torch.cuda.current_device()
    freed = ow.release_caches()
    assert freed >= a.nbytes and torch.cuda.current_device() == before
    assert ow.release_caches() == 0
    assert np.array_equal(ow.render_bench(jobs), a)


@pytest.mark.gpu
@needs_solver
def test_gpu_power_amp_wide_batch_properties_and_sampled_rows():
    """600 rows (75 CTAs, every SM busy) of chain-level material: outputs finite and inside the adapter's clamp, rows independent of their
    neighbours (a row rendered alone is bit-identical), sampled rows equal to the oracle inside the north star's bound with equal reset counts."""
    import openwurli_b200 as ow
    sr, n, rows = 44100.0, 600, 600
    t = np.arange(n) / sr
    rng = np.random.default_rng(11)
    amps = rng.uniform(0.002, 0.25, rows)[:, None]
    freqs = (110.0 * 2 ** rng.uniform(0, 5, rows))[:, None]
    x = np.ascontiguousarray(amps * np.sin(2 * np.pi * freqs * t) * np.exp(-3.0 * t) + 0.002 * rng.standard_normal((rows, n)))
    y, rails, cnt = ow.power_amp_batch(x, sr, want_state=True)
    assert np.isfinite(y).all() and np.abs(y).max() <= 1.0 and np.all(rails > 20.0) and np.all(rails <= 24.5 + 1e-9)
    for i in (0, 1, 299, 598, 599):
        yo, ro, resets = _oracle_amp(x[i], sr)
        err, rel = _close(y[i], yo)
        assert err <= 1e-6 and rel <= 1e-7 and int(cnt[i, 0]) == resets, (i, err, rel, cnt[i], resets)
        assert np.array_equal(ow.power_amp_batch(np.ascontiguousarray(x[i:i + 1]), sr)[0], y[i])


def test_row_per_lane_elimination_equals_the_sequential_one(tmp_path):
    """pa_tile_solve16 on the lane emulator vs the reference's sequential partial-pivoting elimination: 400 random 16x16 systems (exact
    pivot ties, zero columns, duplicated rows): solutions bit-identical, the same systems declared singular."""
    exe = str(tmp_path / "pa_solve_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "tests"), "-o", exe,
                    os.path.join(ROOT, "tests", "pa_solve_check.cpp"), os.path.join(ROOT, "openwurli_b200", "csrc", "host_pa_setup.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    assert out.startswith("bad=0 singular=") and int(out.split("=")[-1]) > 20, out


def test_render_cli_rejects_an_unknown_power_amp(capsys):
    from openwurli_b200.cli import preamp_bench
    assert preamp_bench.main(["render", "--power-amp", "valve", "--duration", "0.01"]) == 1
    assert "Unknown --power-amp" in capsys.readouterr().err


@pytest.mark.gpu
@needs_solver
def test_gpu_render_cli_with_the_melange_power_amp_matches_the_oracle_wav(tmp_path):
    """`preamp-bench render --power-amp melange [--no-rail-sag]` through the CLI mirror: the 24-bit file equals the oracle's chain-B preamp
    tap -> owo_output_stage_melange -> the tool's rounding quantiser (+-1 LSB for a sample on a rounding boundary)."""
    from openwurli_b200 import wav
    from openwurli_b200.cli import preamp_bench
    for extra, sag in (([], 1), (["--no-rail-sag"], 0)):
        p = tmp_path / f"pa{sag}.wav"
        assert preamp_bench.main(["render", "--note", "52", "--velocity", "110", "--duration", "0.2", "--volume", "0.8", "--speaker", "0.6",
                                  "--power-amp", "melange", "--output", str(p)] + extra) == 0
        q, sr = wav.read_wav_pcm24(str(p))
        oj = O.bench_job(midi=52, vel=110, dur=0.2, volume=0.8, speaker=0.6)
        pre = np.ascontiguousarray(O.render_bench_taps(oj)["preamp"])
        ref = np.zeros(len(pre))
        assert L.owo_output_stage_melange(O.dptr(pre), len(pre), 44100.0, 0.8, 0.6, 0, sag, O.dptr(ref)) == 0
        assert sr == 44100 and q.size == ref.size
        assert np.max(np.abs(q.astype(np.int64) - wav.pcm24_round(ref, wav.normalize_scale(ref, False)))) <= 1


@pytest.mark.gpu
def test_gpu_entry_points_leave_the_current_cuda_device_alone():
    """A call that renders on device 0 while the caller's current device is another one (multi-GPU boxes) -- or simply any call on a
    one-GPU box -- returns with the thread's current CUDA device unchanged."""
    import torch
    import openwurli_b200 as ow
    cur = torch.cuda.device_count() - 1
    torch.cuda.set_device(cur)
    probe = torch.ones(4, device="cuda")
    ow.render_bench([ow.bench_job(duration=0.02)], device=0)
    ow.render_voices([ow.voice_job(60, 100, 44100.0, 0.02)], device=0)
    ow.power_amp_batch(np.zeros((1, 32)), device=0)
    pl = ow.Plan.bench([ow.bench_job(duration=0.02)], device=0)
    pl.execute(np.zeros((1, 882)))
    pl.close()
    assert torch.cuda.current_device() == cur and float((probe + 1).sum().item()) == 8.0 and probe.device.index == cur
    torch.cuda.set_device(0)
