"""CPU-only tests: the C-ABI library loads and exports every symbol include/owgpu.h declares, argument
validation and the no-CPU-fallback rule, the product's host-side note-on setup against the oracle
(bit-exact), and the multi-GPU sharding logic (world_size-2 gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import openwurli_b200 as ow
from openwurli_b200 import _abi, shard
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = ow.device_count() > 0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "owgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(owg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 17
    L = ow.lib()
    for name in declared:
        assert hasattr(L, name), f"libowgpu.so does not export {name}"
    assert set(declared) == set(_abi.EXPORTS)
    assert L.owg_abi_version() == 1


def test_struct_layouts_match_header(tmp_path):
    """sizeof of every ABI struct as the C compiler sees include/owgpu.h == the ctypes mirrors."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "owgpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(owg_voice_job), '
                   'sizeof(owg_bench_job), sizeof(owg_event), sizeof(owg_engine_job), sizeof(owg_opts), sizeof(owg_diag));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(t) for t in (_abi.VoiceJob, _abi.BenchJob, _abi.Event, _abi.EngineJob, _abi.Opts, _abi.Diag)]
    assert sizes[:2] == [40, 80]
    assert [C.sizeof(t) for t in (O.VoiceJob, O.BenchJob, O.Diag)] == [sizes[0], sizes[1], sizes[5]]


def test_bad_arguments_are_rejected():
    L = ow.lib()
    h = C.c_void_p()
    assert L.owg_plan_bench(None, 3, None, C.byref(h)) == _abi.OWG_E_BAD_ARG
    bad = ow.bench_job(sample_rate=0.0)
    arr = (_abi.BenchJob * 1)(bad)
    assert L.owg_plan_bench(arr, 1, None, C.byref(h)) == _abi.OWG_E_BAD_ARG
    bad = ow.voice_job(duration=-1.0)
    arr = (_abi.VoiceJob * 1)(bad)
    assert L.owg_plan_voices(arr, 1, None, C.byref(h)) == _abi.OWG_E_BAD_ARG
    assert b"invalid" in L.owg_last_error()


@pytest.mark.skipif(HAS_GPU, reason="checks the behaviour on a machine without a CUDA device")
def test_no_cpu_fallback_without_device():
    with pytest.raises(ow.OwgError) as e:
        ow.Voice.render_note(60, 0.8, 0.1, 44100.0)
    assert e.value.code == _abi.OWG_E_NO_DEVICE
    with pytest.raises(ow.OwgError) as e:
        ow.render_bench([ow.bench_job(duration=0.01)])
    assert e.value.code == _abi.OWG_E_NO_DEVICE


@pytest.mark.skipif(HAS_GPU, reason="checks the behaviour on a machine without a CUDA device")
def test_every_render_entry_point_refuses_without_device():
    """No entry point has a CPU path: engines, preamp batch, chain batch, metrics, calibrate and render-midi all return NO_DEVICE."""
    from openwurli_b200 import smf
    x = np.zeros((2, 64))
    calls = [lambda: ow.render_engines([ow.engine_job([(0, ow.NOTE_ON, 60, 0.5)], duration=0.01)]),
             lambda: ow.preamp_batch(x, 48000.0),
             lambda: ow.chain_batch(x, [ow.bench_job(), ow.bench_job()]),
             lambda: ow.render_bench_metrics([ow.calibrate_job(60, 100)]),
             lambda: ow.render_calibrate([60], [100]),
             lambda: ow.render_midi([[(0.0, smf.NOTE_ON, 60, 100)]], tail=0.01),
             lambda: ow.render_voices([ow.voice_job(duration=0.01)])]
    for f in calls:
        with pytest.raises(ow.OwgError) as e:
            f()
        assert e.value.code == _abi.OWG_E_NO_DEVICE


def test_bad_arguments_are_rejected_before_any_device_work():
    L = ow.lib()
    assert L.owg_render_engines(None, 1, None, 0, None) == _abi.OWG_E_BAD_ARG
    assert L.owg_render_midi(None, 1, None, 0, None) == _abi.OWG_E_BAD_ARG
    assert L.owg_chain_batch(None, 0, 1, 1, None, 7, None, 0, None) == _abi.OWG_E_BAD_ARG
    assert L.owg_render_calibrate(None, 1, None, 0.1, 0.4, None, None) == _abi.OWG_E_BAD_ARG
    assert L.owg_render_calibrate(None, 0, None, 0.4, 0.1, None, None) == _abi.OWG_E_BAD_ARG     # empty window
    assert L.owg_render_engines(None, 0, None, 0, None) == 0 and L.owg_render_midi(None, 0, None, 0, None) == 0   # empty batches are fine
    assert b"owg_" in L.owg_last_error() or len(L.owg_last_error()) > 0


def _host_init(job):
    out = np.zeros(61)
    assert ow.lib().owg_host_voice_init(C.byref(job), O.dptr(out)) == 0
    return out


def _oracle_init(job):
    oj = O.VoiceJob(job.midi, job.mlp_enabled, job.attack_noise, 0, job.noise_seed, job.velocity, job.sample_rate,
                    job.duration_s, job.ds_override)
    out = np.zeros(61)
    O.lib().owo_voice_init(C.byref(oj), O.dptr(out))
    return out


def test_host_note_on_setup_is_bit_identical_to_oracle():
    """Every (key, velocity) of the 64 x 127 grid, MLP on and off: the product's host-side parameterisation
    (openwurli_b200/csrc/host_setup.cpp) must equal the oracle's Voice::note_on restatement bit for bit."""
    bad = 0
    for midi in range(33, 97):
        for vel in range(1, 128, 3):
            for mlp in (False, True):
                j = ow.voice_job(midi, vel, 44100.0, 0.25, mlp=mlp)
                a, b = _host_init(j), _oracle_init(j)
                if not np.array_equal(a, b):
                    bad += 1
                    if bad < 5:
                        print(midi, vel, mlp, np.nonzero(a != b)[0], (a - b)[a != b])
    assert bad == 0


@pytest.mark.parametrize("sr", [44100.0, 48000.0, 96000.0, 22050.0])
def test_host_setup_other_rates_and_overrides(sr):
    for midi, vel, seed, ds, noise in [(33, 1, 0, None, True), (96, 127, 1, 0.3, True), (60, 64, 0xFFFFFFFF, None, False),
                                       (20, 100, 7, None, True), (110, 50, 9, 0.5, True)]:
        j = ow.voice_job(midi, vel, sr, 0.1, mlp=True, attack_noise=noise, seed=seed, displacement_scale=ds)
        assert np.array_equal(_host_init(j), _oracle_init(j)), (midi, vel, seed)
    j = ow.voice_job(60, 0, sr, 0.1, velocity_norm=0.0)
    assert np.array_equal(_host_init(j), _oracle_init(j))


def test_host_chain_setup_matches_oracle():
    for sr in (44100.0, 48000.0, 96000.0):
        for ch in (0.0, 0.0005, 0.001, 0.3, 0.5, 0.9985, 0.999, 1.0, 1.7, -0.2):
            j = ow.bench_job(speaker=ch, volume=0.37, sample_rate=sr)
            a, b = np.zeros(18), np.zeros(18)
            assert ow.lib().owg_host_chain_init(C.byref(j), O.dptr(a)) == 0
            oj = O.bench_job(sr=sr, speaker=ch, volume=0.37)
            O.lib().owo_chain_init(C.byref(oj), O.dptr(b))
            assert np.array_equal(a, b), (sr, ch, a - b)


def test_job_defaults_mirror_the_cli():
    j = ow.bench_job()  # preamp-bench render defaults, main.rs:372-392
    assert (j.v.midi, j.v.duration_s, j.r_ldr, j.volume, j.speaker_character, j.tremolo_depth, j.v.sample_rate) == \
        (60, 2.0, 1e6, 0.60, 1.0, 0.0, 44100.0)
    assert j.v.velocity == 100 / 127.0 and j.v.mlp_enabled == 1 and j.v.attack_noise == 1
    assert j.v.noise_seed == (60 * 2654435761) % 2 ** 32  # main.rs:405
    v = ow.voice_job()  # Voice::render_note: MLP off (voice.rs:209)
    assert v.mlp_enabled == 0


def test_shard_partition_properties():
    jobs = [ow.bench_job(note=33 + k % 64, velocity=1 + k % 127, duration=0.1 + (k % 7) * 0.05,
                         tremolo_depth=(k % 3) * 0.25) for k in range(500)]
    for ws in (1, 2, 3, 8):
        parts = [shard.shard_indices(jobs, ws, r) for r in range(ws)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(500))
        costs = [sum(shard.job_cost(jobs[i]) for i in p) for p in parts]
        assert max(costs) - min(costs) <= 2 * max(shard.job_cost(j) for j in jobs) + 1
    assert shard.shard_indices([], 4, 1) == []


def test_shard_gloo_world2(tmp_path):
    """world_size-2 gloo run of the same partition + host-side gather bench.py uses."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import openwurli_b200 as ow\nfrom openwurli_b200 import shard\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "jobs = [ow.bench_job(note=33 + k % 64, velocity=1 + k % 127, duration=0.25) for k in range(101)]\n"
        "mine = shard.shard_indices(jobs, w, r)\n"
        "cnt = torch.tensor([len(mine), sum(mine)], dtype=torch.int64)\n"
        "dist.all_reduce(cnt)\n"
        "t = torch.tensor([1.0 + r]); dist.all_reduce(t, op=dist.ReduceOp.MAX)\n"
        "assert cnt[0].item() == 101 and cnt[1].item() == sum(range(101)), cnt\n"
        "assert t.item() == float(w)\n"
        "dist.barrier(); dist.destroy_process_group()\n"
        "print('ok', r)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_legacy_preamp_plan_constants_match_oracle_bitwise():
    """make_legacy_group (host_setup.cpp) vs the oracle's DkPreamp::new / reset / set_ldr_resistance: same arithmetic, same bits."""
    import ctypes as C
    for sr in (88200.0, 96000.0, 44100.0, 48000.0, 192000.0):
        for r in (float("nan"), 1e6, 1e6 + 0.005, 19000.0, 500.0, 5e6, float("inf")):
            a, b = np.zeros(188), np.zeros(188)
            assert ow.lib().owg_host_legacy_group(sr, r, O.dptr(a)) == 0
            assert O.lib().owo_legacy_group(sr, r, O.dptr(b)) == 0
            assert a.tobytes() == b.tobytes(), (sr, r, np.nonzero(a != b)[0][:8])
    # S_base really is the inverse of A_base = 2C/T + G_base  (A_base = A_neg_base + 2 G_base, with G from the netlist)
    a = np.zeros(188)
    ow.lib().owg_host_legacy_group(88200.0, 1e6, O.dptr(a))
    S, An = a[:64].reshape(8, 8), a[64:128].reshape(8, 8)
    assert np.allclose(S, S.T, rtol=1e-9, atol=1e-18)           # reciprocal network
    assert abs(a[186] - 1e-6) < 1e-20 and a[187] == 1.0 / 1e6


def test_tile_term_tables_reproduce_build_rhs_bit_for_bit(tmp_path):
    """owg_tile_tables.h (the lane-tiled kernel's build_rhs slots: 4 lanes x 18 padded terms) against the straightforward rows."""
    exe = tmp_path / "tile_tables_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "tests", "tile_tables_check.cpp")])
    assert subprocess.check_output([str(exe)]).decode().strip().endswith("OK")


def test_constant_tables_of_product_and_oracle_are_the_same_extraction():
    """The product and the oracle each carry their own copy of the generated-solver constants (neither tree includes from the
    other); the two copies must be byte-identical, and identical to a fresh extraction when the reference is present."""
    a = open(os.path.join(ROOT, "openwurli_b200", "csrc", "ow_consts.inc"), "rb").read()
    b = open(os.path.join(ROOT, "oracle", "ow_consts.inc"), "rb").read()
    assert a == b


def test_caller_buffers_are_checked_before_the_c_call():
    """api.py: an undersized `out` (rows or columns) is refused in Python -- the C side sees only the stride -- and the two ow_consts.inc
    copies (product / oracle, written from one extraction) are byte-identical."""
    import openwurli_b200 as ow
    jobs = [ow.bench_job(duration=0.01), ow.bench_job(duration=0.02)]
    with pytest.raises(ValueError):
        ow.render_bench(jobs, out=np.zeros((1, 882)))
    with pytest.raises(ValueError):
        ow.render_bench(jobs, out=np.zeros((2, 441)))
    with pytest.raises(ValueError):
        ow.render_voices([j.v for j in jobs], out=np.zeros((2, 100)))
    with pytest.raises(ValueError):
        ow.preamp_batch(np.zeros((2, 64)), 44100.0, out=np.zeros((2, 32)))
    with pytest.raises(ValueError):
        ow.power_amp_batch(np.zeros((2, 64)), out=np.zeros((1, 64)))
    assert ow.release_caches() == 0      # nothing cached without a device
    a = open(os.path.join(ROOT, "openwurli_b200", "csrc", "ow_consts.inc"), "rb").read()
    b = open(os.path.join(ROOT, "oracle", "ow_consts.inc"), "rb").read()
    assert a == b


def test_rust_sys_crate_declares_exactly_the_header_functions():
    """bindings/rust/owgpu-sys (uncompiled here: no Rust toolchain) must at least declare every function of include/owgpu.h and nothing
    else, with the same number of parameters, and mirror the size-critical structs field for field in count."""
    hdr = open(os.path.join(ROOT, "include", "owgpu.h")).read()
    rs = open(os.path.join(ROOT, "bindings", "rust", "owgpu-sys", "src", "lib.rs")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    c_fns = {m.group(1): m.group(2) for m in re.finditer(r"\b(owg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr_nc)}
    r_fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (owg_[a-z0-9_]+)\s*\(([^)]*)\)", rs)}
    assert set(c_fns) == set(r_fns), (sorted(set(c_fns) - set(r_fns)), sorted(set(r_fns) - set(c_fns)))
    for name, args in c_fns.items():
        n_c = 0 if args.strip() in ("", "void") else args.count(",") + 1
        n_r = len([a for a in r_fns[name].split(",") if a.strip()])
        assert n_c == n_r, (name, n_c, n_r)
    # owg_opts: same number of fields (the layout test against ctypes sizes is test_struct_sizes...)
    c_opts = re.search(r"typedef struct owg_opts \{(.*?)\} owg_opts;", hdr_nc, re.S).group(1)
    r_opts = re.search(r"pub struct owg_opts \{(.*?)\n\}", rs, re.S).group(1)
    assert c_opts.count(";") == len(re.findall(r"pub \w+:", r_opts))
