"""ctypes binding of the CPU oracle (oracle/_build/liboworacle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "liboworacle.so")


class VoiceJob(C.Structure):
    _fields_ = [("midi", C.c_uint8), ("mlp_enabled", C.c_uint8), ("attack_noise", C.c_uint8), ("flags", C.c_uint8),
                ("noise_seed", C.c_uint32), ("velocity", C.c_double), ("sample_rate", C.c_double),
                ("duration_s", C.c_double), ("ds_override", C.c_double)]


class BenchJob(C.Structure):
    _fields_ = [("v", VoiceJob), ("r_ldr", C.c_double), ("tremolo_depth", C.c_double), ("volume", C.c_double),
                ("speaker_character", C.c_double), ("no_preamp", C.c_int32), ("no_poweramp", C.c_int32)]


class Event(C.Structure):
    _fields_ = [("sample", C.c_int64), ("kind", C.c_uint8), ("note", C.c_uint8), ("_pad0", C.c_uint16),
                ("velocity", C.c_float)]


class EngineJob(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("duration_s", C.c_double), ("volume", C.c_double),
                ("tremolo_depth", C.c_double), ("speaker_character", C.c_double), ("mlp_enabled", C.c_int32),
                ("block_size", C.c_int32), ("warm_up", C.c_int32), ("_pad0", C.c_int32),
                ("ev", C.POINTER(Event)), ("n_ev", C.c_int64)]


class Diag(C.Structure):
    _fields_ = [("nr_iter_hist", C.c_uint64 * 16), ("nr_max_iter", C.c_uint64), ("be_fallback", C.c_uint64),
                ("voltage_damp", C.c_uint64), ("nan_reset", C.c_uint64), ("shadow_nr_iter_hist", C.c_uint64 * 16),
                ("shadow_be_fallback", C.c_uint64), ("shadow_nan_reset", C.c_uint64),
                ("poweramp_iter_hist", C.c_uint64 * 9), ("tremolo_nr_iter_hist", C.c_uint64 * 16),
                ("tremolo_be_fallback", C.c_uint64), ("kernels_launched", C.c_uint64)]


def default_seed(midi):
    return (midi * 2654435761) & 0xFFFFFFFF


def voice_job(midi=60, vel=100, sr=44100.0, dur=2.0, mlp=False, noise=True, seed=None, ds=float("nan"), vel_norm=None):
    return VoiceJob(midi, 1 if mlp else 0, 1 if noise else 0, 0, default_seed(midi) if seed is None else seed,
                    (vel / 127.0) if vel_norm is None else vel_norm, sr, dur, ds)


def bench_job(midi=60, vel=100, sr=44100.0, dur=2.0, mlp=True, noise=True, r_ldr=1e6, depth=0.0, volume=0.60,
              speaker=1.0, no_preamp=False, no_poweramp=False, seed=None, ds=float("nan")):
    return BenchJob(voice_job(midi, vel, sr, dur, mlp, noise, seed, ds), r_ldr, depth, volume, speaker,
                    1 if no_preamp else 0, 1 if no_poweramp else 0)


def n_samples(dur, sr):
    return int(dur * sr)


class MidiEvent(C.Structure):  # include/owgpu.h owg_midi_event
    _fields_ = [("time_s", C.c_double), ("kind", C.c_uint8), ("note", C.c_uint8), ("velocity", C.c_uint8), ("_pad0", C.c_uint8),
                ("_pad1", C.c_int32)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        dp = C.POINTER(C.c_double)
        L.owo_render_voices.argtypes = [C.POINTER(VoiceJob), C.c_int64, dp, C.c_int64, C.c_int]
        L.owo_render_bench.argtypes = [C.POINTER(BenchJob), C.c_int64, dp, C.c_int64, C.c_int]
        L.owo_render_bench_taps.argtypes = [C.POINTER(BenchJob), dp, dp, dp, dp, dp]
        L.owo_preamp_batch.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_double,
                                       C.c_double, dp, C.c_int64, C.c_int]
        L.owo_render_engines.argtypes = [C.POINTER(EngineJob), C.c_int64, C.POINTER(C.c_float), C.c_int64, C.c_int]
        L.owo_render_engines_model.argtypes = [C.POINTER(EngineJob), C.c_int64, C.POINTER(C.c_float), C.c_int64, C.c_int, C.c_int]
        L.owo_alias_stimulus.argtypes = [C.c_uint8, C.c_uint8, C.c_double, C.c_double, C.c_double, dp]
        L.owo_last_diag.argtypes = [C.POINTER(Diag)]
        L.owo_engine_script.argtypes = [C.c_double, C.c_int, dp, C.c_int64, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int32), C.c_int64, dp]
        L.owo_engine_script.restype = C.c_int64
        L.owo_stage_timers.argtypes = [C.c_int, dp]
        for name, args in [("owo_midi_to_freq", [C.c_int]), ("owo_tip_mass_ratio", [C.c_int]),
                           ("owo_reed_length_mm", [C.c_int]), ("owo_reed_compliance", [C.c_int]),
                           ("owo_pickup_displacement_scale", [C.c_int]), ("owo_fundamental_decay_rate", [C.c_int]),
                           ("owo_output_scale", [C.c_int, C.c_double]), ("owo_velocity_exponent", [C.c_int]),
                           ("owo_velocity_scurve", [C.c_double]), ("owo_register_trim_db", [C.c_int]),
                           ("owo_pickup_rms_proxy", [C.c_double] * 3), ("owo_freq_detune", [C.c_int]),
                           ("owo_dwell_time", [C.c_double] * 2), ("owo_onset_ramp_time", [C.c_double] * 2),
                           ("owo_pickup_soft_saturate", [C.c_double]), ("owo_fast_exp", [C.c_double]),
                           ("owo_poweramp", [C.c_double]),
                           ("owo_biquad_bp_gain", [C.c_double] * 4 + [C.c_int])]:
            f = getattr(L, name)
            f.argtypes = args
            f.restype = C.c_double
        L.owo_mode_ratios.argtypes = [C.c_double, dp]
        L.owo_spatial_coupling.argtypes = [C.c_double, C.c_double, dp]
        L.owo_mode_amplitude_offsets.argtypes = [C.c_int, dp]
        L.owo_dwell_attenuation.argtypes = [C.c_double, C.c_double, dp, dp]
        L.owo_mlp_infer.argtypes = [C.c_int, C.c_double, dp]
        L.owo_note_params.argtypes = [C.c_int, dp]
        L.owo_preamp_matrices.argtypes = [C.c_double, C.c_double, dp, dp, dp, dp]
        L.owo_preamp_settled.argtypes = [dp]
        L.owo_preamp_run.argtypes = [C.c_double, C.c_double, dp, C.c_int64, dp]
        L.owo_tremolo_run.argtypes = [C.c_double, C.c_double, C.c_int64, dp]
        L.owo_tremolo_osc.argtypes = [C.c_double, C.c_int64, C.c_int64, dp, dp]
        L.owo_speaker_run.argtypes = [C.c_double, C.c_double, dp, C.c_int64, dp]
        L.owo_oversampler_roundtrip.argtypes = [dp, C.c_int64, dp]
        L.owo_render_bench_model.argtypes = [C.POINTER(BenchJob), C.c_int64, dp, C.c_int64, C.c_int, C.c_int]
        L.owo_rail_dynamics.argtypes = [C.c_double, dp, C.c_int64, dp]
        L.owo_power_amp_melange.argtypes = [C.c_double, C.c_int, dp, C.c_int64, dp, dp, C.c_int64]
        L.owo_power_amp_melange.restype = C.c_int64
        L.owo_output_stage_melange.argtypes = [dp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, dp]
        L.owo_power_amp_matrices.argtypes = [C.c_double, dp]
        L.owo_power_amp_settled.argtypes = [C.c_int64, dp]
        L.owo_alias_analyze.argtypes = [dp, C.c_int64, C.c_double, C.c_double, C.c_double, dp]
        L.owo_preamp_batch_diag.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double, dp, C.c_int64, C.c_int]
        L.owo_preamp_batch_model.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_double,
                                             C.c_double, dp, C.c_int64, C.c_int, C.c_int]
        L.owo_chain_batch.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.POINTER(BenchJob), C.c_int, dp, C.c_int64, C.c_int, C.c_int]
        L.owo_render_midi.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.POINTER(C.c_uint64)]
        L.owo_calibrate_rows.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, dp, C.c_int, C.c_double, C.c_double,
                                         C.c_int, dp, C.c_int]
        L.owo_legacy_dc.argtypes = [C.c_double, dp]
        L.owo_legacy_group.argtypes = [C.c_double, C.c_double, dp]
        L.owo_legacy_run.argtypes = [C.c_double, C.c_double, C.c_int, dp, C.c_int64, dp, dp]
        L.owo_legacy_idle_pump.argtypes = [C.c_double, C.c_double, C.c_int64, dp, dp]
        L.owo_voice_init.argtypes = [C.POINTER(VoiceJob), dp]
        L.owo_chain_init.argtypes = [C.POINTER(BenchJob), dp]
        _lib = L
    return _lib


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def vec(fn, n, *args):
    out = np.zeros(n)
    fn(*args, dptr(out))
    return out


def render_voices(jobs, threads=1):
    n = len(jobs)
    ns = [n_samples(j.duration_s, j.sample_rate) for j in jobs]
    stride = max(ns) if ns else 0
    out = np.zeros((n, stride))
    arr = (VoiceJob * n)(*jobs)
    assert lib().owo_render_voices(arr, n, dptr(out), stride, threads) == 0
    return out


MELANGE12, LEGACY8 = 0, 1


def render_bench(jobs, threads=1, preamp_model=MELANGE12):
    n = len(jobs)
    ns = [n_samples(j.v.duration_s, j.v.sample_rate) for j in jobs]
    stride = max(ns) if ns else 0
    out = np.zeros((n, stride))
    arr = (BenchJob * n)(*jobs)
    assert lib().owo_render_bench_model(arr, n, dptr(out), stride, threads, preamp_model) == 0
    return out


def calibrate_rows(notes, velocities, cfg6=(0.85, 0.75, 0.02, 0.95, -35.0, -0.04), zero_trim=False, volume=0.40, speaker=1.0,
                   preamp_model=0, threads=4):
    nn, nv = len(notes), len(velocities)
    rows = np.zeros((nn * nv, 18))
    a = (C.c_uint8 * nn)(*notes)
    b = (C.c_uint8 * nv)(*velocities)
    c = np.array(cfg6, dtype=np.float64)
    assert lib().owo_calibrate_rows(a, nn, b, nv, dptr(c), 1 if zero_trim else 0, volume, speaker, preamp_model, dptr(rows), threads) == 0
    return rows


def render_bench_taps(job):
    n = n_samples(job.v.duration_s, job.v.sample_rate)
    nos = 2 * n if job.v.sample_rate < 88200.0 else n
    fin, voice, pre = np.zeros(n), np.zeros(n), np.zeros(n)
    r, sh = np.zeros(nos), np.zeros(nos)
    lib().owo_render_bench_taps(C.byref(job), dptr(fin), dptr(voice), dptr(pre), dptr(r), dptr(sh))
    return dict(final=fin, voice=voice, preamp=pre, r_ldr=r, shadow=sh)


def last_diag():
    d = Diag()
    lib().owo_last_diag(C.byref(d))
    return d


def engine_job(events, sr=44100.0, dur=1.0, volume=0.5, depth=0.5, speaker=0.0, mlp=True, block=512, warm_up=True):
    arr = (Event * max(len(events), 1))()
    for i, (smp, kind, note, vel) in enumerate(events):
        arr[i] = Event(int(smp), int(kind), int(note), 0, float(vel))
    j = EngineJob(sr, dur, volume, depth, speaker, 1 if mlp else 0, block, 1 if warm_up else 0, 0, arr, len(events))
    j._keepalive = arr
    return j


def render_engines(jobs, threads=1, preamp_model=0):
    n = len(jobs)
    stride = max([n_samples(j.duration_s, j.sample_rate) for j in jobs], default=0)
    out = np.zeros((n, stride), dtype=np.float32)
    arr = (EngineJob * n)(*jobs)
    assert lib().owo_render_engines_model(arr, n, out.ctypes.data_as(C.POINTER(C.c_float)), stride, threads, preamp_model) == 0
    return out


# ---- engine session scripts (owo_engine_script): the vocabulary of the reference's engine tests (engine.rs:682-1178) ----
NOTE_ON, NOTE_OFF, SUSTAIN, RENDER, QUERY, SET_VOLUME, SET_DEPTH, SET_CHARACTER, RESET, SET_SAMPLE_RATE, SET_MLP = range(11)


def engine_script(ops, sr=44100.0, model=0):
    """Runs a script of WurliEngine calls on the oracle.  ops: list of (kind, a, b).  Returns (captured float32 samples, list of
    QUERY results as dicts, final smoother values)."""
    a = np.array([[float(k), float(x), float(y)] for (k, x, y) in ops], dtype=np.float64)
    n_cap = int(sum(x for (k, x, y) in ops if k == RENDER and y))
    n_q = sum(1 for (k, x, y) in ops if k == QUERY)
    out = np.zeros(max(n_cap, 1), dtype=np.float32)
    counts = np.zeros(max(n_q, 1) * 8, dtype=np.int32)
    sm = np.zeros(3)
    n = lib().owo_engine_script(sr, model, a.ctypes.data_as(C.POINTER(C.c_double)), len(ops), out.ctypes.data_as(C.POINTER(C.c_float)), len(out),
                                counts.ctypes.data_as(C.POINTER(C.c_int32)), len(counts), sm.ctypes.data_as(C.POINTER(C.c_double)))
    assert n == n_cap, n
    keys = ("active", "held", "sustained", "releasing", "has_steal", "note_held", "note_sustained", "sustain_flag")
    qs = [dict(zip(keys, counts[8 * i:8 * i + 8].tolist())) for i in range(n_q)]
    return out[:n_cap], qs, sm


def stage_timers(n=100000):
    """ns per preamp-rate sample of the restatement, single thread: (preamp step with matrix rebuild, preamp step static, oscillator step)."""
    out = (C.c_double * 3)()
    lib().owo_stage_timers(int(n), out)
    return {"preamp_step_with_rebuild_ns": out[0], "preamp_step_static_ns": out[1], "tremolo_oscillator_step_ns": out[2]}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"
