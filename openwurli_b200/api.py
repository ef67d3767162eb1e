"""Python host layer over libowgpu: job construction with the reference's defaults, plans, output buffers."""
import ctypes as C
import math

import numpy as np

from . import _abi
from ._abi import BenchJob, Diag, Opts, VoiceJob, OWG_OUT_DEVICE, OWG_OUT_HOST, check, lib

NAN = float("nan")


def default_noise_seed(midi):
    """(midi as u32).wrapping_mul(2654435761) -- voice.rs:208, preamp-bench main.rs:405."""
    return (int(midi) * 2654435761) & 0xFFFFFFFF


VOICE_NO_ONSET = 1


def voice_job(midi=60, velocity=100, sample_rate=44100.0, duration=2.0, mlp=False, attack_noise=True, seed=None,
              displacement_scale=None, velocity_norm=None, no_onset=False):
    """One Voice::note_on + render job.  `velocity` is MIDI 0..127 (normalised as vel/127.0 in f64 like the CLIs,
    reed-renderer main.rs:83); pass velocity_norm to give the 0..1 value directly."""
    v = (velocity / 127.0) if velocity_norm is None else float(velocity_norm)
    return VoiceJob(int(midi), 1 if mlp else 0, 1 if attack_noise else 0, VOICE_NO_ONSET if no_onset else 0,
                    default_noise_seed(midi) if seed is None else int(seed) & 0xFFFFFFFF, v, float(sample_rate),
                    float(duration), NAN if displacement_scale is None else float(displacement_scale))


def bench_job(note=60, velocity=100, duration=2.0, ldr=1_000_000.0, volume=0.60, speaker=1.0, tremolo_depth=0.0,
              sample_rate=44100.0, no_poweramp=False, no_preamp=False, no_attack_noise=False, no_mlp=False,
              displacement_scale=None, seed=None, no_onset=False):
    """One `preamp-bench render` job; keyword names and defaults are the CLI flags (main.rs:372-392).
    no_onset=True builds the reed with onset_time 0 as `run_calibrate` does (main.rs:1165-1174)."""
    return BenchJob(voice_job(note, velocity, sample_rate, duration, not no_mlp, not no_attack_noise, seed,
                              displacement_scale, no_onset=no_onset), float(ldr), float(tremolo_depth), float(volume), float(speaker),
                    1 if no_preamp else 0, 1 if no_poweramp else 0)


def _samples(duration, sample_rate):
    x = duration * sample_rate  # `(duration * sample_rate) as usize`: truncation
    return int(x) if x > 0 and math.isfinite(x) else 0


MELANGE12, LEGACY8 = 0, 1  # owg_opts.preamp_model: gen_preamp.rs 12-node (north star) | dk_preamp_legacy.rs 8-node (reference default build)


PA_BEHAVIORAL, PA_MELANGE, PA_MELANGE_IDEAL_RAILS = 0, 1, 2   # owg_opts.power_amp_model


def _opts(device=-1, out_location=OWG_OUT_HOST, stream=None, collect_diag=False, preamp_model=MELANGE12, devices=None, power_amp_model=PA_BEHAVIORAL):
    o = Opts()
    lib().owg_default_opts(C.byref(o))
    o.preamp_model = int(preamp_model)
    o.power_amp_model = int(power_amp_model)
    o.device = device
    o.out_location = out_location
    o.stream = stream
    o.collect_diag = 1 if collect_diag else 0
    if devices is not None:  # in-call multi-GPU fan-out (owg_opts.device_mask)
        mask = 0
        for d in devices:
            mask |= 1 << int(d)
        o.device_mask = mask
    return o


def release_caches(device=-1):
    """Free the library's grow-only device staging buffers (host-output renders keep one as large as the largest render so far);
    returns the bytes released."""
    return int(lib().owg_release_caches(int(device)))


def device_count():
    return lib().owg_device_count()


def last_diag():
    d = Diag()
    check(lib().owg_last_diag(C.byref(d)))
    return d


def fp64_peak(device=-1, fma=True, ms_target=50.0):
    """Measured FP64-pipe instruction rate (1e12 instr/s); a DFMA is one instruction = 2 flop."""
    r = C.c_double(0.0)
    check(lib().owg_fp64_peak(device, 1 if fma else 0, ms_target, C.byref(r)))
    return r.value


def fp64_probe(mode, device=-1):
    """Raw owg_fp64_peak modes: 10..13 = ns per dependent DADD/DMUL/DFMA/DDIV; 100+k = warp-instr/s (1e12) with k active lanes."""
    r = C.c_double(0.0)
    check(lib().owg_fp64_peak(device, int(mode), 20.0, C.byref(r)))
    return r.value


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _out_ptr(out):
    """(pointer, stride, location) of a caller buffer: numpy float64 [n,stride] or torch tensor (cpu / cuda)."""
    if isinstance(out, np.ndarray):
        assert out.dtype == np.float64 and out.ndim == 2 and out.flags.c_contiguous
        return out.ctypes.data, out.shape[1], OWG_OUT_HOST
    import torch
    assert isinstance(out, torch.Tensor) and out.dtype == torch.float64 and out.dim() == 2 and out.is_contiguous()
    if out.is_cuda:
        # The library launches on its own (non-blocking) stream unless one is passed in: work the caller has queued on torch's
        # current stream for this buffer (fills, copies) must have landed before our kernels touch it.
        torch.cuda.current_stream(out.device).synchronize()
    return out.data_ptr(), out.shape[1], (OWG_OUT_DEVICE if out.is_cuda else OWG_OUT_HOST)


class Plan:
    """A planned batch: note-on parameterisation done once on the host, init records resident in HBM.
    execute() may be called repeatedly; each call is a complete independent render."""

    def __init__(self, handle, n, kind):
        self._h = handle
        self.n = n
        self.kind = kind

    @classmethod
    def bench(cls, jobs, device=-1, stream=None, collect_diag=False, preamp_model=MELANGE12):
        arr = (BenchJob * len(jobs))(*jobs)
        h = C.c_void_p()
        o = _opts(device, OWG_OUT_HOST, stream, collect_diag, preamp_model)
        check(lib().owg_plan_bench(arr, len(jobs), C.byref(o), C.byref(h)))
        return cls(h, len(jobs), "bench")

    @classmethod
    def voices(cls, jobs, device=-1, stream=None, collect_diag=False):
        arr = (VoiceJob * len(jobs))(*jobs)
        h = C.c_void_p()
        o = _opts(device, OWG_OUT_HOST, stream, collect_diag)
        check(lib().owg_plan_voices(arr, len(jobs), C.byref(o), C.byref(h)))
        return cls(h, len(jobs), "voices")

    @property
    def max_samples(self):
        return lib().owg_plan_samples(self._h, -1)

    def samples(self, i):
        return lib().owg_plan_samples(self._h, i)

    def execute(self, out):
        ptr, stride, loc = _out_ptr(out)
        assert out.shape[0] >= self.n
        check(lib().owg_plan_execute(self._h, ptr, stride, loc))
        return out

    @property
    def h2d_bytes(self):
        return lib().owg_plan_h2d_bytes(self._h)

    @property
    def kernel_launches(self):
        return lib().owg_plan_kernel_launches(self._h)

    def last_timing(self):
        a, b = C.c_float(0), C.c_float(0)
        check(lib().owg_plan_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if self._h:
            lib().owg_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _alloc_out(n, stride, out):
    """The caller's buffer (checked: at least n rows of at least `stride` samples; the C side can only check the stride) or a fresh one."""
    if out is not None:
        if tuple(out.shape)[0] < n or (n > 0 and tuple(out.shape)[1] < stride):
            raise ValueError(f"out has shape {tuple(out.shape)}, need at least ({n}, {stride})")
        return out
    return np.zeros((n, stride), dtype=np.float64)


def _device_for(out, device):
    """A CUDA tensor is rendered on ITS device: opts.device = -1 would mean the current device, which need not be the tensor's."""
    if not isinstance(out, np.ndarray) and getattr(out, "is_cuda", False):
        idx = out.device.index
        if device not in (-1, idx):
            raise ValueError(f"out lives on cuda:{idx} but device={device} was requested")
        return idx
    return device


def render_voices(jobs, out=None, device=-1, collect_diag=False, devices=None):
    """Batch of Voice::note_on + render (chain V). Returns [n, max_samples] float64.  devices=[0, 1, ...]: one call fans the job
    list out over those GPUs (host output only)."""
    stride = max([_samples(j.duration_s, j.sample_rate) for j in jobs], default=0)
    out = _alloc_out(len(jobs), stride, out)
    if len(jobs) == 0 or stride == 0:
        return out
    ptr, st, loc = _out_ptr(out)
    arr = (VoiceJob * len(jobs))(*jobs)
    o = _opts(_device_for(out, device), loc, None, collect_diag, devices=devices)
    check(lib().owg_render_voices(arr, len(jobs), ptr, st, C.byref(o)))
    return out


def render_bench(jobs, out=None, device=-1, collect_diag=False, preamp_model=MELANGE12, devices=None, power_amp_model=PA_BEHAVIORAL):
    """Batch of `preamp-bench render` (chain B). Returns [n, max_samples] float64 (pre-WAV samples).  devices=[0, 1, ...]: one call
    fans the job list out over those GPUs (contiguous ranges balanced by rendered samples; host output only).
    preamp_model: MELANGE12 (`--features melange-preamp`, the north-star path) or LEGACY8 (the reference's default build).
    power_amp_model: PA_BEHAVIORAL (`legacy-power-amp`, the default build), PA_MELANGE (the melange 7-BJT solver with rail sag,
    `--no-default-features`) or PA_MELANGE_IDEAL_RAILS (`--no-rail-sag`)."""
    stride = max([_samples(j.v.duration_s, j.v.sample_rate) for j in jobs], default=0)
    out = _alloc_out(len(jobs), stride, out)
    if len(jobs) == 0 or stride == 0:
        return out
    ptr, st, loc = _out_ptr(out)
    arr = (BenchJob * len(jobs))(*jobs)
    o = _opts(_device_for(out, device), loc, None, collect_diag, preamp_model, devices=devices, power_amp_model=power_amp_model)
    check(lib().owg_render_bench(arr, len(jobs), ptr, st, C.byref(o)))
    return out


def preamp_batch(x, fs_base, oversample=True, tremolo_depth=0.0, r_ldr=1_000_000.0, out=None, device=-1, preamp_model=MELANGE12,
                 collect_diag=False):
    """Preamp-only batch (BASELINE config 2): rows of `x` [n_inst, n_samp] (numpy float64, or a torch CUDA tensor)
    through upsample_2x -> DkPreamp::process_sample x2 -> downsample_2x (`process_oversampled`, preamp-bench main.rs:961-974),
    LDR driven by Tremolo::new(tremolo_depth, fs_preamp) when tremolo_depth > 0 (main.rs:432-461), else static r_ldr."""
    if out is None:
        out = np.zeros_like(x) if isinstance(x, np.ndarray) else x.new_zeros(x.shape)
    pin, sin, lin = _out_ptr(x)
    pout, sout, lout = _out_ptr(out)
    assert lin == lout, "input and output must both be host or both be device buffers"
    if tuple(out.shape) != tuple(x.shape):
        raise ValueError(f"out has shape {tuple(out.shape)}, x has {tuple(x.shape)}")
    o = _opts(_device_for(out, device), lout, preamp_model=preamp_model, collect_diag=collect_diag)
    check(lib().owg_preamp_batch(pin, sin, x.shape[0], x.shape[1], float(fs_base), 1 if oversample else 0,
                                 float(tremolo_depth), float(r_ldr), pout, sout, C.byref(o)))
    return out


def power_amp_batch(x, sample_rate=44100.0, rail_sag=True, out=None, device=-1, want_state=False):
    """The melange 7-BJT power amplifier alone (power_amp.rs melange_adapter + gen_power_amp.rs): row i of `x` [n_inst, n_samp] through
    PowerAmp::new_at_sample_rate(sample_rate), set_rail_sag(rail_sag), process().  want_state: also return (rails [n, 2] = rail_voltages()
    after the last sample, counters [n, 4] = divergence-guard resets, backward-Euler retries, NaN resets, last_nr_iterations)."""
    if out is None:
        out = np.zeros_like(x) if isinstance(x, np.ndarray) else x.new_zeros(x.shape)
    pin, sin, lin = _out_ptr(x)
    pout, sout, lout = _out_ptr(out)
    assert lin == lout, "input and output must both be host or both be device buffers"
    if tuple(out.shape) != tuple(x.shape):
        raise ValueError(f"out has shape {tuple(out.shape)}, x has {tuple(x.shape)}")
    o = _opts(_device_for(out, device), lout)
    rails = np.zeros((x.shape[0], 2), dtype=np.float64)
    counters = np.zeros((x.shape[0], 4), dtype=np.uint32)
    check(lib().owg_power_amp_batch(pin, sin, x.shape[0], x.shape[1], float(sample_rate), 1 if rail_sag else 0, pout, sout,
                                    rails.ctypes.data_as(C.POINTER(C.c_double)) if want_state else None,
                                    counters.ctypes.data_as(C.POINTER(C.c_uint32)) if want_state else None, C.byref(o)))
    return (out, rails, counters) if want_state else out


RESET_THEN_SET, SET_THEN_RESET = 0, 1


def chain_batch(x, params, init_order=RESET_THEN_SET, out=None, device=-1, preamp_model=MELANGE12, power_amp_model=PA_BEHAVIORAL):
    """Rows of `x` [n_inst, n_samp] (any mono signal, e.g. a sum of voices) through the full chain B of params[i] (bench_job(...): sample
    rate, ldr, tremolo depth, volume, speaker, bypass flags).  init_order = SET_THEN_RESET is how `render-poly` / `render-midi` construct
    the static preamp (main.rs:1463-1464)."""
    if out is None:
        out = np.zeros_like(x) if isinstance(x, np.ndarray) else x.new_zeros(x.shape)
    pin, sin, lin = _out_ptr(x)
    pout, sout, lout = _out_ptr(out)
    assert lin == lout and len(params) == x.shape[0]
    if tuple(out.shape) != tuple(x.shape):
        raise ValueError(f"out has shape {tuple(out.shape)}, x has {tuple(x.shape)}")
    arr = (BenchJob * len(params))(*params)
    o = _opts(_device_for(out, device), lout, preamp_model=preamp_model, power_amp_model=power_amp_model)
    check(lib().owg_chain_batch(pin, sin, x.shape[0], x.shape[1], arr, int(init_order), pout, sout, C.byref(o)))
    return out


MIDI_ON, MIDI_OFF, MIDI_PEDAL = 0, 1, 2


def render_midi(streams, volume=0.60, speaker=1.0, no_poweramp=False, tail=2.0, device=-1, preamp_model=MELANGE12, power_amp_model=PA_BEHAVIORAL):
    """`preamp-bench render-midi` for a batch of event lists (smf.timed_events output: [(time_s, kind, note, velocity)]): the tool's own
    voice manager and chain.  Returns (list of float64 arrays, one per stream)."""
    from . import smf
    jobs, keep, ns = [], [], []
    for ev in streams:
        arr = (_abi.MidiEvent * max(len(ev), 1))()
        for k, (t, kind, a, b) in enumerate(ev):
            code = MIDI_ON if kind == smf.NOTE_ON else (MIDI_OFF if kind == smf.NOTE_OFF else MIDI_PEDAL)
            arr[k] = _abi.MidiEvent(float(t), code, int(a) if code != MIDI_PEDAL else 0, int(b) if code == MIDI_ON else (int(a) if code == MIDI_PEDAL else 0), 0, 0)
        n = smf.total_samples(ev, tail, 44100.0)
        keep.append(arr)
        ns.append(n)
        jobs.append(_abi.MidiJob(arr, len(ev), n, float(volume), float(speaker), 1 if no_poweramp else 0, 0))
    stride = max(ns, default=0)
    out = np.zeros((len(jobs), max(stride, 1)), dtype=np.float64)
    if jobs and stride:
        ja = (_abi.MidiJob * len(jobs))(*jobs)
        o = _opts(device, OWG_OUT_HOST, preamp_model=preamp_model, power_amp_model=power_amp_model)
        check(lib().owg_render_midi(ja, len(jobs), out.ctypes.data, out.shape[1], C.byref(o)))
    return [out[i, :ns[i]] for i in range(len(jobs))]


METRIC_COLUMNS = ("peak_db", "rms_db", "h2_h1_db", "peak", "mean_sq", "h1", "h2")


def render_bench_metrics(jobs, window=(0.100, 0.400), device=-1, preamp_model=MELANGE12):
    """Chain B with the `run_calibrate` T5 analysis reduced on the device (output mode "metrics", BASELINE config 4):
    returns [n, 7] float64 = METRIC_COLUMNS over the window [start_s, end_s) (default 100-400 ms, main.rs:1139-1141)."""
    out = np.zeros((len(jobs), 7), dtype=np.float64)
    if len(jobs) == 0:
        return out
    arr = (BenchJob * len(jobs))(*jobs)
    o = _opts(device, OWG_OUT_HOST, preamp_model=preamp_model)
    check(lib().owg_render_bench_metrics(arr, len(jobs), float(window[0]), float(window[1]),
                                         out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(o)))
    return out


def calibrate_job(note, velocity, volume=0.60, speaker=1.0, tremolo_depth=0.0):
    """The render `preamp-bench calibrate` measures (run_calibrate, main.rs:1128-1263): 0.5 s at 44.1 kHz, reed built
    with onset 0, no attack noise, no MLP, default CalibrationConfig, preamp at a static 1 MOhm."""
    return bench_job(note=note, velocity=velocity, duration=0.5, volume=volume, speaker=speaker, tremolo_depth=tremolo_depth,
                     no_mlp=True, no_attack_noise=True, no_onset=True)


CALIBRATE_COLUMNS = ("ds_at_c4", "ds_actual", "y_peak", "t2_peak_db", "t2_rms_db", "t2_h2_h1_db", "t3_peak_db", "t3_rms_db", "t4_peak_db",
                     "t4_rms_db", "t4_h2_h1_db", "t5_peak_db", "t5_rms_db", "t5_h2_h1_db", "proxy_db", "trim_db", "proxy_error_db",
                     "tanh_compression_db")


def calib_cfg(ds_at_c4=None, ds_clamp_max=None, zero_trim=False, **kw):
    """CalibrationConfig (tables.rs:256-277); defaults = CalibrationConfig::default().  `preamp-bench calibrate` itself defaults to
    ds_at_c4=0.75, ds_clamp_max=0.82 (main.rs:1072-1078)."""
    c = _abi.CalibCfg()
    lib().owg_default_calib_cfg(C.byref(c))
    if ds_at_c4 is not None:
        c.ds_at_c4 = float(ds_at_c4)
    if ds_clamp_max is not None:
        c.ds_clamp_hi = float(ds_clamp_max)
    c.zero_trim = 1 if zero_trim else 0
    for k, v in kw.items():
        setattr(c, k, float(v))
    return c


def render_calibrate(notes, velocities, cfg=None, volume=0.40, speaker=1.0, window=(0.100, 0.400), device=-1, preamp_model=MELANGE12):
    """`preamp-bench calibrate` (run_calibrate, main.rs:1127-1260) for notes x velocities: all five taps reduced on the device.
    Returns [len(notes) * len(velocities), 18] float64 in CALIBRATE_COLUMNS order (note-major like the CSV)."""
    jobs = [calibrate_job(n, v, volume=volume, speaker=speaker) for n in notes for v in velocities]
    out = np.zeros((len(jobs), len(CALIBRATE_COLUMNS)), dtype=np.float64)
    if not jobs:
        return out
    arr = (BenchJob * len(jobs))(*jobs)
    o = _opts(device, OWG_OUT_HOST, preamp_model=preamp_model)
    check(lib().owg_render_calibrate(arr, len(jobs), C.byref(cfg) if cfg is not None else None, float(window[0]), float(window[1]),
                                     out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(o)))
    return out


NOTE_ON, NOTE_OFF, SUSTAIN = 0, 1, 2
# parameter automation events (engine.rs:378-388): the value travels in the event's velocity field, `note` is ignored
SET_VOLUME, SET_TREMOLO_DEPTH, SET_SPEAKER_CHARACTER = 3, 4, 5


def engine_job(events, sample_rate=44100.0, duration=1.0, volume=0.5, tremolo_depth=0.5, speaker_character=0.0,
               mlp=True, block_size=512, warm_up=True):
    """One WurliEngine stream (engine.rs:194-462).  `events` = [(sample, kind, note, velocity)], sorted by sample;
    kind in NOTE_ON / NOTE_OFF / SUSTAIN (note != 0 = pedal down); velocity is rounded to f32 like note_on(u8, f32);
    SET_VOLUME / SET_TREMOLO_DEPTH / SET_SPEAKER_CHARACTER carry the new smoother target in `velocity` (applied at the start of their block).
    Defaults are the engine's own (volume 0.5, depth 0.5, character 0.0, MLP on; engine.rs:224-226); warm_up=True
    constructs through set_sample_rate() as the plugin does (0.6 s warm-up)."""
    arr = (_abi.Event * max(len(events), 1))()
    for i, (smp, kind, note, vel) in enumerate(events):
        arr[i] = _abi.Event(int(smp), int(kind), int(note), 0, float(vel))
    j = _abi.EngineJob(float(sample_rate), float(duration), float(volume), float(tremolo_depth), float(speaker_character),
                       1 if mlp else 0, int(block_size), 1 if warm_up else 0, 0, arr, len(events))
    j._keepalive = arr
    return j


def render_engines(jobs, out=None, device=-1, preamp_model=MELANGE12):
    """Batch of WurliEngine streams (chain E). Returns [n, max_samples] float32 (WurliEngine::render writes f32).
    preamp_model=LEGACY8 renders what the reference's default plugin build (legacy 8-node preamp) produces."""
    stride = max([_samples(j.duration_s, j.sample_rate) for j in jobs], default=0)
    if out is None:
        out = np.zeros((len(jobs), stride), dtype=np.float32)
    if len(jobs) == 0 or stride == 0:
        return out
    if isinstance(out, np.ndarray):
        assert out.dtype == np.float32 and out.flags.c_contiguous
        ptr, st, loc = out.ctypes.data, out.shape[1], OWG_OUT_HOST
    else:
        import torch
        assert out.dtype == torch.float32 and out.is_contiguous()
        if out.is_cuda:
            torch.cuda.current_stream(out.device).synchronize()
        ptr, st, loc = out.data_ptr(), out.shape[1], (OWG_OUT_DEVICE if out.is_cuda else OWG_OUT_HOST)
    arr = (_abi.EngineJob * len(jobs))(*jobs)
    o = _opts(device, loc, preamp_model=preamp_model)
    check(lib().owg_render_engines(arr, len(jobs), ptr, st, C.byref(o)))
    return out


ALIAS_COLUMNS = 29  # f0_hz, h1_dbfs, harmonic_db[12], harmonic_dbc[12], max_step_up_db, max_step_up_from_harmonic, hf_band_dbc


class AliasAuditResult:
    """alias_audit::AliasAuditResult (alias_audit.rs:60-80) from one row of owg_alias_analyze."""

    def __init__(self, row):
        self.f0_hz, self.h1_dbfs = float(row[0]), float(row[1])
        self.harmonic_db, self.harmonic_dbc = np.array(row[2:14]), np.array(row[14:26])
        self.max_step_up_db, self.max_step_up_from_harmonic, self.hf_band_dbc = float(row[26]), int(row[27]), float(row[28])


def note_hz(note):
    """alias_audit::midi_note_hz (alias_audit.rs:279-282)."""
    return 440.0 * 2.0 ** ((note - 69.0) / 12.0)


def alias_analyze(rows, sample_rate, nominal_f0, analyze_seconds=0.5, device=-1):
    """alias_audit::analyze (alias_audit.rs:163-204) for every row of `rows` ([n, samples] float32 / float64, numpy or torch CUDA):
    the reduction runs on the device; returns [n, ALIAS_COLUMNS] float64."""
    n = rows.shape[0]
    res = np.zeros((n, ALIAS_COLUMNS))
    f0 = np.ascontiguousarray(np.asarray(nominal_f0, dtype=np.float64))
    assert f0.shape == (n,)
    if isinstance(rows, np.ndarray):
        assert rows.flags.c_contiguous and rows.dtype in (np.float32, np.float64)
        ptr, st, loc, f32 = rows.ctypes.data, rows.shape[1], OWG_OUT_HOST, rows.dtype == np.float32
    else:
        import torch
        assert rows.is_contiguous() and rows.dtype in (torch.float32, torch.float64)
        if rows.is_cuda:
            torch.cuda.current_stream(rows.device).synchronize()
        ptr, st, loc, f32 = rows.data_ptr(), rows.shape[1], (OWG_OUT_DEVICE if rows.is_cuda else OWG_OUT_HOST), rows.dtype == torch.float32
    o = _opts(device, loc)
    dp = C.POINTER(C.c_double)
    check(lib().owg_alias_analyze(ptr, 1 if f32 else 0, st, n, rows.shape[1], float(sample_rate), float(analyze_seconds),
                                  f0.ctypes.data_as(dp), res.ctypes.data_as(dp), C.byref(o)))
    return res


def render_engines_alias(jobs, nominal_f0, analyze_seconds=0.5, device=-1, preamp_model=MELANGE12):
    """Engine streams rendered AND analysed on the device (alias_audit::run_with_note batched): only [n, ALIAS_COLUMNS] numbers come back."""
    n = len(jobs)
    res = np.zeros((n, ALIAS_COLUMNS))
    f0 = np.ascontiguousarray(np.asarray(nominal_f0, dtype=np.float64))
    assert f0.shape == (n,)
    arr = (_abi.EngineJob * max(n, 1))(*jobs)
    o = _opts(device, OWG_OUT_HOST, preamp_model=preamp_model)
    dp = C.POINTER(C.c_double)
    check(lib().owg_render_engines_alias(arr, n, float(analyze_seconds), f0.ctypes.data_as(dp), res.ctypes.data_as(dp), C.byref(o)))
    return res


class Voice:
    """Mirror of openwurli_dsp::voice::Voice's one-shot constructor (voice.rs:191-221)."""

    @staticmethod
    def render_note(midi_note, velocity, duration_secs, sample_rate):
        """Voice::render_note: velocity is 0..1; MLP off; seed = midi*2654435761. Returns f64 samples."""
        return Voice.render_note_with_scale(midi_note, velocity, duration_secs, sample_rate, None)

    @staticmethod
    def render_note_with_scale(midi_note, velocity, duration_secs, sample_rate, displacement_scale):
        j = voice_job(midi_note, 0, sample_rate, duration_secs, mlp=False, attack_noise=True,
                      displacement_scale=displacement_scale, velocity_norm=velocity)
        return render_voices([j])[0]


def reed_renderer(note=60, velocity=100, duration=1.0):
    """`reed-renderer -n NOTE -v VEL -d DUR` (tools/reed-renderer/src/main.rs:70-107): 44.1 kHz, voice only.
    Returns the f64 buffer the CLI would quantise to 24-bit PCM."""
    return Voice.render_note(note, velocity / 127.0, duration, 44100.0)


def preamp_bench_render(**flags):
    """`preamp-bench render` with its CLI flag names as keywords (see bench_job). Returns f64 samples."""
    return render_bench([bench_job(**flags)])[0]


from .wav import pcm24_round, pcm24_truncate  # noqa: E402,F401  (WAV quantisations of the two reference tools)
