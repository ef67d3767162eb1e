//! Raw bindings to `include/owgpu.h` (ABI version 1) plus the thin safe helpers INTEGRATION.md describes.
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Rust toolchain.  The struct layouts are checked
//! against the C header by `tests/test_host_logic.py` (sizeof through gcc) on the C side; keep the two in step.
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_void};

pub const OWG_OK: i32 = 0;
pub const OWG_E_BAD_ARG: i32 = -1;
pub const OWG_E_NO_DEVICE: i32 = -2;
pub const OWG_E_CUDA: i32 = -3;
pub const OWG_E_OOM: i32 = -4;
pub const OWG_E_UNSUPPORTED: i32 = -5;

pub const OWG_OUT_HOST: i32 = 0;
pub const OWG_OUT_DEVICE: i32 = 1;
pub const OWG_PREAMP_MELANGE12: i32 = 0; // cfg(feature = "melange-preamp")
pub const OWG_PREAMP_LEGACY8: i32 = 1; // the reference's default build
pub const OWG_VOICE_NO_ONSET: u8 = 1;
pub const OWG_EV_NOTE_ON: u8 = 0;
pub const OWG_EV_NOTE_OFF: u8 = 1;
pub const OWG_EV_SUSTAIN: u8 = 2;
/// Parameter automation (engine.rs:378-388): the new smoother target travels in `owg_event::velocity`.
pub const OWG_EV_SET_VOLUME: u8 = 3;
pub const OWG_EV_SET_TREMOLO_DEPTH: u8 = 4;
pub const OWG_EV_SET_SPEAKER_CHARACTER: u8 = 5;
/// Columns of one `owg_alias_analyze` result row (alias_audit::AliasAuditResult).
pub const OWG_ALIAS_COLUMNS: usize = 29;
pub const OWG_ROWS_F64: i32 = 0;
pub const OWG_ROWS_F32: i32 = 1;
pub const OWG_INIT_RESET_THEN_SET: i32 = 0;
pub const OWG_INIT_SET_THEN_RESET: i32 = 1;
pub const OWG_METRIC_COLUMNS: usize = 7;
pub const OWG_CALIBRATE_COLUMNS: usize = 18;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_voice_job {
    pub midi: u8,
    pub mlp_enabled: u8,
    pub attack_noise: u8,
    pub flags: u8,
    pub noise_seed: u32,
    pub velocity: f64,
    pub sample_rate: f64,
    pub duration_s: f64,
    pub ds_override: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_bench_job {
    pub v: owg_voice_job,
    pub r_ldr: f64,
    pub tremolo_depth: f64,
    pub volume: f64,
    pub speaker_character: f64,
    pub no_preamp: i32,
    pub no_poweramp: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_event {
    pub sample: i64,
    pub kind: u8,
    pub note: u8,
    pub _pad0: u16,
    pub velocity: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_engine_job {
    pub sample_rate: f64,
    pub duration_s: f64,
    pub volume: f64,
    pub tremolo_depth: f64,
    pub speaker_character: f64,
    pub mlp_enabled: i32,
    pub block_size: i32,
    pub warm_up: i32,
    pub _pad0: i32,
    pub ev: *const owg_event,
    pub n_ev: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_midi_event {
    pub time_s: f64,
    pub kind: u8,
    pub note: u8,
    pub velocity: u8,
    pub _pad0: u8,
    pub _pad1: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_midi_job {
    pub ev: *const owg_midi_event,
    pub n_ev: i64,
    pub n_samples: i64,
    pub volume: f64,
    pub speaker_character: f64,
    pub no_poweramp: i32,
    pub _pad0: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_calib_cfg {
    pub ds_at_c4: f64,
    pub ds_exponent: f64,
    pub ds_clamp_lo: f64,
    pub ds_clamp_hi: f64,
    pub target_db: f64,
    pub voicing_slope: f64,
    pub zero_trim: i32,
    pub _pad0: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_opts {
    pub device: i32,
    pub out_location: i32,
    pub precision: i32,
    pub preamp_model: i32,
    pub stream: *mut c_void,
    pub collect_diag: i32,
    /// bit d = use CUDA device d; two or more bits: the call fans out over those GPUs (host output)
    pub device_mask: u32,
    /// OWG_POWER_AMP_*: 0 behavioural (`legacy-power-amp`, default build), 1 melange 7-BJT solver with rail sag, 2 the same with ideal rails
    pub power_amp_model: i32,
    pub _reserved: [i32; 5],
}

pub const OWG_POWER_AMP_BEHAVIORAL: i32 = 0;
pub const OWG_POWER_AMP_MELANGE: i32 = 1;
pub const OWG_POWER_AMP_MELANGE_IDEAL_RAILS: i32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct owg_diag {
    pub nr_iter_hist: [u64; 16],
    pub nr_max_iter: u64,
    pub be_fallback: u64,
    pub voltage_damp: u64,
    pub nan_reset: u64,
    pub shadow_nr_iter_hist: [u64; 16],
    pub shadow_be_fallback: u64,
    pub shadow_nan_reset: u64,
    pub poweramp_iter_hist: [u64; 9],
    pub tremolo_nr_iter_hist: [u64; 16],
    pub tremolo_be_fallback: u64,
    pub kernels_launched: u64,
}

#[repr(C)]
pub struct owg_plan {
    _private: [u8; 0],
}

extern "C" {
    pub fn owg_abi_version() -> i32;
    pub fn owg_device_count() -> i32;
    pub fn owg_last_error() -> *const c_char;
    pub fn owg_default_opts(o: *mut owg_opts);
    pub fn owg_default_calib_cfg(c: *mut owg_calib_cfg);
    pub fn owg_render_voices(jobs: *const owg_voice_job, n: i64, out: *mut f64, stride: i64, opts: *const owg_opts) -> i32;
    pub fn owg_render_bench(jobs: *const owg_bench_job, n: i64, out: *mut f64, stride: i64, opts: *const owg_opts) -> i32;
    pub fn owg_render_bench_metrics(jobs: *const owg_bench_job, n: i64, window_start_s: f64, window_end_s: f64, metrics: *mut f64,
                                    opts: *const owg_opts) -> i32;
    pub fn owg_render_calibrate(jobs: *const owg_bench_job, n: i64, cfg: *const owg_calib_cfg, window_start_s: f64, window_end_s: f64,
                                rows: *mut f64, opts: *const owg_opts) -> i32;
    pub fn owg_preamp_batch(input: *const f64, in_stride: i64, n_inst: i64, n_samp: i64, fs_base: f64, oversample: i32, tremolo_depth: f64,
                            r_ldr_static: f64, out: *mut f64, out_stride: i64, opts: *const owg_opts) -> i32;
    /// PowerAmp::new_at_sample_rate + set_rail_sag + process per row (power_amp.rs:279-465 over gen_power_amp.rs); `rails` `[n][2]` and
    /// `counters` `[n][4]` (resets, backward-Euler retries, NaN resets, last_nr_iterations) may be null.
    pub fn owg_power_amp_batch(input: *const f64, in_stride: i64, n_inst: i64, n_samp: i64, sample_rate: f64, rail_sag: i32, out: *mut f64,
                               out_stride: i64, rails: *mut f64, counters: *mut u32, opts: *const owg_opts) -> i32;
    // ---- probes and self-tests (what the test-suite pins the host logic and the device arithmetic with) ----
    /// Voice::note_on's init record (61 doubles) as the host computes it
    pub fn owg_host_voice_init(job: *const owg_voice_job, out61: *mut f64) -> i32;
    /// chain-B output-stage parameters of one job (18 doubles)
    pub fn owg_host_chain_init(job: *const owg_bench_job, out18: *mut f64) -> i32;
    /// the legacy 8-node preamp's plan-time record (188 doubles)
    pub fn owg_host_legacy_group(preamp_sr: f64, r_static: f64, out188: *mut f64) -> i32;
    /// FP64 pipe micro-benchmark (the roofline denominator): tera-instructions per second of DFMA (fma = 1) or unfused DADD/DMUL (fma = 0)
    pub fn owg_fp64_peak(device: i32, fma: i32, ms_target: f32, tera_instr_per_s: *mut f64) -> i32;
    /// profiling counters of the lane-tiled kernels (DIAG builds of a call)
    pub fn owg_debug_counters(out: *mut u64, n: i32, reset: i32) -> i32;
    /// shared-reciprocal division against the compiler's `/` on random operands, on the device
    pub fn owg_selftest_division(n_per_thread: i64, seed: u64, mismatches: *mut u64, tested: *mut u64) -> i32;
    /// frees the grow-only device staging buffers no call is using (device = -1: every device); returns the bytes released
    pub fn owg_release_caches(device: i32) -> i64;
    pub fn owg_chain_batch(input: *const f64, in_stride: i64, n_inst: i64, n_samp: i64, params: *const owg_bench_job, init_order: i32,
                           out: *mut f64, out_stride: i64, opts: *const owg_opts) -> i32;
    pub fn owg_render_engines(jobs: *const owg_engine_job, n: i64, out: *mut f32, stride: i64, opts: *const owg_opts) -> i32;
    pub fn owg_render_midi(jobs: *const owg_midi_job, n: i64, out: *mut f64, stride: i64, opts: *const owg_opts) -> i32;
    /// alias_audit::analyze (alias_audit.rs:163-282) on the device; `results` is host memory `[n_rows][OWG_ALIAS_COLUMNS]`.
    pub fn owg_alias_analyze(rows: *const core::ffi::c_void, row_dtype: i32, stride: i64, n_rows: i64, n_samples: i64, sample_rate: f64,
                             analyze_seconds: f64, nominal_f0: *const f64, results: *mut f64, opts: *const owg_opts) -> i32;
    /// Engine streams rendered and analysed on the device; all jobs share sample_rate and duration.
    pub fn owg_render_engines_alias(jobs: *const owg_engine_job, n: i64, analyze_seconds: f64, nominal_f0: *const f64, results: *mut f64,
                                    opts: *const owg_opts) -> i32;
    pub fn owg_plan_bench(jobs: *const owg_bench_job, n: i64, opts: *const owg_opts, plan: *mut *mut owg_plan) -> i32;
    pub fn owg_plan_voices(jobs: *const owg_voice_job, n: i64, opts: *const owg_opts, plan: *mut *mut owg_plan) -> i32;
    pub fn owg_plan_execute(plan: *mut owg_plan, out: *mut f64, stride: i64, out_location: i32) -> i32;
    pub fn owg_plan_samples(plan: *const owg_plan, i: i64) -> i64;
    pub fn owg_plan_h2d_bytes(plan: *const owg_plan) -> i64;
    pub fn owg_plan_kernel_launches(plan: *const owg_plan) -> i64;
    pub fn owg_plan_last_timing(plan: *const owg_plan, main_kernel_ms: *mut f32, total_ms: *mut f32) -> i32;
    pub fn owg_plan_destroy(plan: *mut owg_plan);
    pub fn owg_last_diag(out: *mut owg_diag) -> i32;
}

/// The preamp the reference build would have compiled in (`dk_preamp/mod.rs:14-20`).
pub fn default_opts(melange_preamp: bool) -> owg_opts {
    let mut o = core::mem::MaybeUninit::<owg_opts>::uninit();
    // SAFETY: owg_default_opts fully initialises the struct.
    let mut o = unsafe {
        owg_default_opts(o.as_mut_ptr());
        o.assume_init()
    };
    o.preamp_model = if melange_preamp { OWG_PREAMP_MELANGE12 } else { OWG_PREAMP_LEGACY8 };
    o
}

fn check(rc: i32) -> Result<(), String> {
    if rc == OWG_OK {
        return Ok(());
    }
    // SAFETY: owg_last_error returns a NUL-terminated thread-local string.
    let msg = unsafe { std::ffi::CStr::from_ptr(owg_last_error()) }.to_string_lossy().into_owned();
    Err(format!("libowgpu error {rc}: {msg}"))
}

/// Batch form of `Voice::render_note` (voice.rs:191-221): MLP off, default seed, attack noise on.
pub fn render_notes(notes: &[(u8, f64)], duration_secs: f64, sample_rate: f64) -> Result<Vec<Vec<f64>>, String> {
    let n_samp = (duration_secs * sample_rate) as usize; // voice.rs:214 truncation
    if notes.is_empty() || n_samp == 0 {
        return Ok(vec![Vec::new(); notes.len()]);
    }
    let jobs: Vec<owg_voice_job> = notes
        .iter()
        .map(|&(m, v)| owg_voice_job {
            midi: m,
            mlp_enabled: 0,
            attack_noise: 1,
            flags: 0,
            noise_seed: (m as u32).wrapping_mul(2654435761),
            velocity: v,
            sample_rate,
            duration_s: duration_secs,
            ds_override: f64::NAN,
        })
        .collect();
    let mut out = vec![0.0f64; jobs.len() * n_samp];
    // SAFETY: pointers and sizes describe the vectors above; opts = NULL selects the defaults.
    check(unsafe { owg_render_voices(jobs.as_ptr(), jobs.len() as i64, out.as_mut_ptr(), n_samp as i64, core::ptr::null()) })?;
    Ok(out.chunks(n_samp).map(|c| c.to_vec()).collect())
}
