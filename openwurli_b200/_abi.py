"""ctypes view of include/owgpu.h (libowgpu.so).  No compute happens in Python and there is no
CPU fallback: if the CUDA library is missing or no device is usable, calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libowgpu.so")

OWG_OK = 0
OWG_E_BAD_ARG, OWG_E_NO_DEVICE, OWG_E_CUDA, OWG_E_OOM, OWG_E_UNSUPPORTED = -1, -2, -3, -4, -5
OWG_OUT_HOST, OWG_OUT_DEVICE = 0, 1


class VoiceJob(C.Structure):
    """owg_voice_job == arguments of Voice::note_on (voice.rs:28-34) + render length."""
    _fields_ = [("midi", C.c_uint8), ("mlp_enabled", C.c_uint8), ("attack_noise", C.c_uint8), ("flags", C.c_uint8),
                ("noise_seed", C.c_uint32), ("velocity", C.c_double), ("sample_rate", C.c_double),
                ("duration_s", C.c_double), ("ds_override", C.c_double)]


class BenchJob(C.Structure):
    """owg_bench_job == flags of `preamp-bench render` (tools/preamp-bench/src/main.rs:372-392)."""
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


class Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("out_location", C.c_int32), ("precision", C.c_int32),
                ("preamp_model", C.c_int32), ("stream", C.c_void_p), ("collect_diag", C.c_int32),
                ("device_mask", C.c_uint32), ("power_amp_model", C.c_int32), ("_reserved", C.c_int32 * 5)]


class MidiEvent(C.Structure):
    _fields_ = [("time_s", C.c_double), ("kind", C.c_uint8), ("note", C.c_uint8), ("velocity", C.c_uint8), ("_pad0", C.c_uint8),
                ("_pad1", C.c_int32)]


class MidiJob(C.Structure):
    _fields_ = [("ev", C.POINTER(MidiEvent)), ("n_ev", C.c_int64), ("n_samples", C.c_int64), ("volume", C.c_double),
                ("speaker_character", C.c_double), ("no_poweramp", C.c_int32), ("_pad0", C.c_int32)]


class CalibCfg(C.Structure):
    _fields_ = [("ds_at_c4", C.c_double), ("ds_exponent", C.c_double), ("ds_clamp_lo", C.c_double), ("ds_clamp_hi", C.c_double),
                ("target_db", C.c_double), ("voicing_slope", C.c_double), ("zero_trim", C.c_int32), ("_pad0", C.c_int32)]


class Diag(C.Structure):
    _fields_ = [("nr_iter_hist", C.c_uint64 * 16), ("nr_max_iter", C.c_uint64), ("be_fallback", C.c_uint64),
                ("voltage_damp", C.c_uint64), ("nan_reset", C.c_uint64), ("shadow_nr_iter_hist", C.c_uint64 * 16),
                ("shadow_be_fallback", C.c_uint64), ("shadow_nan_reset", C.c_uint64),
                ("poweramp_iter_hist", C.c_uint64 * 9), ("tremolo_nr_iter_hist", C.c_uint64 * 16),
                ("tremolo_be_fallback", C.c_uint64), ("kernels_launched", C.c_uint64)]


EXPORTS = ["owg_abi_version", "owg_device_count", "owg_last_error", "owg_default_opts", "owg_render_voices",
           "owg_render_bench", "owg_render_engines", "owg_preamp_batch", "owg_plan_bench", "owg_plan_voices",
           "owg_plan_execute", "owg_plan_samples", "owg_plan_h2d_bytes", "owg_plan_kernel_launches", "owg_plan_last_timing",
           "owg_plan_destroy", "owg_last_diag", "owg_fp64_peak", "owg_host_voice_init", "owg_host_chain_init", "owg_host_legacy_group", "owg_render_calibrate", "owg_default_calib_cfg", "owg_chain_batch", "owg_render_midi", "owg_selftest_division", "owg_render_bench_metrics", "owg_debug_counters", "owg_alias_analyze", "owg_render_engines_alias", "owg_power_amp_batch", "owg_release_caches"]

_lib = None


class OwgError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libowgpu error {code}: {text}")
        self.code = code


def lib():
    """Load libowgpu.so (fails loudly when the CUDA extension has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `make -C openwurli_b200/csrc` (or __graft_entry__.build()); "
                              "openwurli_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.owg_abi_version.restype = C.c_int
        L.owg_device_count.restype = C.c_int
        L.owg_last_error.restype = C.c_char_p
        L.owg_default_opts.argtypes = [C.POINTER(Opts)]
        L.owg_render_voices.argtypes = [C.POINTER(VoiceJob), C.c_int64, C.c_void_p, C.c_int64, C.POINTER(Opts)]
        L.owg_render_bench.argtypes = [C.POINTER(BenchJob), C.c_int64, C.c_void_p, C.c_int64, C.POINTER(Opts)]
        L.owg_render_engines.argtypes = [C.POINTER(EngineJob), C.c_int64, C.c_void_p, C.c_int64, C.POINTER(Opts)]
        L.owg_preamp_batch.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_double,
                                       C.c_double, C.c_void_p, C.c_int64, C.POINTER(Opts)]
        L.owg_plan_bench.argtypes = [C.POINTER(BenchJob), C.c_int64, C.POINTER(Opts), C.POINTER(C.c_void_p)]
        L.owg_plan_voices.argtypes = [C.POINTER(VoiceJob), C.c_int64, C.POINTER(Opts), C.POINTER(C.c_void_p)]
        L.owg_plan_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        L.owg_plan_samples.argtypes = [C.c_void_p, C.c_int64]
        L.owg_plan_samples.restype = C.c_int64
        L.owg_plan_h2d_bytes.argtypes = [C.c_void_p]
        L.owg_plan_h2d_bytes.restype = C.c_int64
        L.owg_plan_kernel_launches.argtypes = [C.c_void_p]
        L.owg_plan_kernel_launches.restype = C.c_int64
        L.owg_plan_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.owg_plan_destroy.argtypes = [C.c_void_p]
        L.owg_plan_destroy.restype = None
        L.owg_last_diag.argtypes = [C.POINTER(Diag)]
        L.owg_fp64_peak.argtypes = [C.c_int32, C.c_int32, C.c_float, dp]
        L.owg_host_voice_init.argtypes = [C.POINTER(VoiceJob), dp]
        L.owg_host_chain_init.argtypes = [C.POINTER(BenchJob), dp]
        L.owg_host_legacy_group.argtypes = [C.c_double, C.c_double, dp]
        L.owg_release_caches.argtypes = [C.c_int32]
        L.owg_release_caches.restype = C.c_int64
        L.owg_power_amp_batch.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32, C.c_void_p, C.c_int64,
                                          C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(Opts)]
        L.owg_chain_batch.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(BenchJob), C.c_int32, C.c_void_p, C.c_int64,
                                      C.POINTER(Opts)]
        L.owg_render_midi.argtypes = [C.POINTER(MidiJob), C.c_int64, C.c_void_p, C.c_int64, C.POINTER(Opts)]
        L.owg_default_calib_cfg.argtypes = [C.POINTER(CalibCfg)]
        L.owg_default_calib_cfg.restype = None
        L.owg_render_calibrate.argtypes = [C.POINTER(BenchJob), C.c_int64, C.POINTER(CalibCfg), C.c_double, C.c_double, dp, C.POINTER(Opts)]
        L.owg_render_bench_metrics.argtypes = [C.POINTER(BenchJob), C.c_int64, C.c_double, C.c_double, dp, C.POINTER(Opts)]
        L.owg_debug_counters.argtypes = [C.POINTER(C.c_uint64), C.c_int32, C.c_int32]
        L.owg_alias_analyze.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double, C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(Opts)]
        L.owg_render_engines_alias.argtypes = [C.POINTER(EngineJob), C.c_int64, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(Opts)]
        L.owg_selftest_division.argtypes = [C.c_int64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def check(rc):
    if rc != OWG_OK:
        raise OwgError(rc, lib().owg_last_error().decode("utf-8", "replace"))
