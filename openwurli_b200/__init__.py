"""openwurli_b200 -- batched OpenWurli (Wurlitzer 200A) rendering on NVIDIA B200.

Host-side mirror of the reference's interfaces for the offline-render hot path:

  Voice.render_note(midi, velocity, duration_secs, sample_rate)      <- openwurli-dsp voice.rs:191-198
  render_voices(jobs)            batch of Voice::note_on + render    <- voice.rs:28-179 (chain V)
  render_bench(jobs)             batch of `preamp-bench render`      <- tools/preamp-bench/src/main.rs:371-496 (chain B)
  reed_renderer(...) / preamp_bench_render(...)   the two CLI entry points with their flags

All arithmetic runs in hand-written sm_100a CUDA behind the C ABI in include/owgpu.h
(openwurli_b200/lib/libowgpu.so).  There is no CPU fallback.
"""
from ._abi import (BenchJob, Diag, OwgError, VoiceJob, OWG_OUT_DEVICE, OWG_OUT_HOST, lib)
from .api import (release_caches, PA_BEHAVIORAL, PA_MELANGE, PA_MELANGE_IDEAL_RAILS, power_amp_batch, SET_VOLUME, SET_TREMOLO_DEPTH, SET_SPEAKER_CHARACTER, ALIAS_COLUMNS, AliasAuditResult, alias_analyze, render_engines_alias, note_hz, RESET_THEN_SET, SET_THEN_RESET, chain_batch, render_midi, CALIBRATE_COLUMNS, LEGACY8, MELANGE12, METRIC_COLUMNS, calib_cfg, render_calibrate, NOTE_OFF, NOTE_ON, SUSTAIN, calibrate_job, render_bench_metrics, Plan, Voice, bench_job, engine_job, render_engines, default_noise_seed, device_count, fp64_peak, last_diag, preamp_batch, preamp_bench_render,
                  reed_renderer, render_bench, render_voices, voice_job)

__all__ = ["release_caches", "PA_BEHAVIORAL", "PA_MELANGE", "PA_MELANGE_IDEAL_RAILS", "power_amp_batch", "SET_VOLUME", "SET_TREMOLO_DEPTH", "SET_SPEAKER_CHARACTER", "ALIAS_COLUMNS", "AliasAuditResult", "alias_analyze", "render_engines_alias", "note_hz", "RESET_THEN_SET", "SET_THEN_RESET", "chain_batch", "render_midi", "CALIBRATE_COLUMNS", "calib_cfg", "render_calibrate", "LEGACY8", "MELANGE12", "METRIC_COLUMNS", "calibrate_job", "render_bench_metrics", "NOTE_OFF", "NOTE_ON", "SUSTAIN", "engine_job", "render_engines", "BenchJob", "Diag", "OwgError", "VoiceJob", "OWG_OUT_DEVICE", "OWG_OUT_HOST", "lib", "Plan", "Voice",
           "bench_job", "default_noise_seed", "device_count", "fp64_peak", "last_diag", "preamp_batch", "preamp_bench_render",
           "reed_renderer", "render_bench", "render_voices", "voice_job"]
