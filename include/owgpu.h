/* owgpu.h -- C ABI of libowgpu: batched OpenWurli rendering on NVIDIA B200 (sm_100a).
 *
 * The reference (hal0zer0/openwurli v0.6.0) is a single-threaded Rust DSP library with no FFI
 * boundary on this path; the entry points below are batch-shaped versions of the seams its own
 * callers use.  Each one cites the reference interface it replaces.  Plain pointers and sizes
 * only; the library never retains caller pointers after a call returns (plans copy what they
 * need).  There is NO CPU fallback: every render call fails with OWG_E_NO_DEVICE when no CUDA
 * device is usable.
 *
 * Every entry point that selects a CUDA device (opts->device / device_mask, a plan's device) restores the calling thread's current
 * device before it returns.
 *
 * Numeric trouble is not an error: the reference's guards (NaN -> reset / zeros, BE fallback,
 * voltage damping) are reproduced on the device and reported through owg_last_diag().
 */
#ifndef OWGPU_H
#define OWGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OWG_ABI_VERSION 1

/* ---- error codes ---------------------------------------------------------------------------- */
#define OWG_OK 0
#define OWG_E_BAD_ARG (-1)    /* null pointer, n<0, stride too small, sample_rate<=0, ... */
#define OWG_E_NO_DEVICE (-2)  /* no usable CUDA device (there is no CPU fallback) */
#define OWG_E_CUDA (-3)       /* CUDA runtime error; text via owg_last_error() */
#define OWG_E_OOM (-4)        /* host or device allocation failed */
#define OWG_E_UNSUPPORTED (-5)/* option combination not implemented by this build */

/* ---- job descriptors ------------------------------------------------------------------------ */

/* = the arguments of Voice::note_on (crates/openwurli-dsp/src/voice.rs:28-34) plus the two
 * post-construction overrides the reference's harnesses use (set_displacement_scale voice.rs:145,
 * disable_attack_noise voice.rs:150) and the render length of Voice::render_note (voice.rs:191-221:
 * n = (duration_s * sample_rate) truncated). */
typedef struct owg_voice_job {
    uint8_t midi;          /* MIDI note 33..96 */
    uint8_t mlp_enabled;   /* 0/1: MLP v2 per-note corrections (voice.rs:62-66) */
    uint8_t attack_noise;  /* 0 = disable_attack_noise() */
    uint8_t flags;         /* OWG_VOICE_* bits */
    uint32_t noise_seed;   /* seed of the jitter and attack-noise LCGs (voice.rs:208: midi*2654435761) */
    double velocity;       /* 0..1 (callers pass vel/127.0) */
    double sample_rate;    /* Hz */
    double duration_s;     /* seconds */
    double ds_override;    /* pickup displacement-scale override; NaN = none */
} owg_voice_job;

#define OWG_VOICE_NO_ONSET 1 /* onset_time = 0.0: the reed is built as run_calibrate does (main.rs:1165-1174) */

/* = the flags of `preamp-bench render` (tools/preamp-bench/src/main.rs:372-392), chain B. */
typedef struct owg_bench_job {
    owg_voice_job v;
    double r_ldr;             /* --ldr, static LDR path resistance when tremolo_depth <= 0 (1e6) */
    double tremolo_depth;     /* --tremolo-depth (0.0 = static r_ldr, main.rs:432-440) */
    double volume;            /* --volume (0.60); applied as volume^2 before the power amp */
    double speaker_character; /* --speaker (1.0) */
    int32_t no_preamp;        /* --no-preamp */
    int32_t no_poweramp;      /* --no-poweramp */
} owg_bench_job;

/* MIDI-ish event for WurliEngine streams (engine.rs:299-374). */
#define OWG_EV_NOTE_ON 0
#define OWG_EV_NOTE_OFF 1
#define OWG_EV_SUSTAIN 2 /* note != 0 -> pedal down */
/* Parameter automation (engine.rs:378-388; the plugin calls these at the top of every process() block, openwurli-plugin/src/lib.rs):
 * the new target is `velocity` (f32 like the plugin's parameter values), the smoothers ramp to it over 5 ms. */
#define OWG_EV_SET_VOLUME 3
#define OWG_EV_SET_TREMOLO_DEPTH 4
#define OWG_EV_SET_SPEAKER_CHARACTER 5
typedef struct owg_event {
    int64_t sample;  /* base-rate sample index at which the event applies (block-quantised like the plugin host) */
    uint8_t kind;
    uint8_t note;
    uint16_t _pad0;
    float velocity;  /* f32, as WurliEngine::note_on takes it (engine.rs:299) */
} owg_event;

/* = WurliEngine::{new, set_sample_rate, set_volume, set_tremolo_depth, set_speaker_character,
 *   set_mlp_enabled, render} (engine.rs:194-462), chain E. */
typedef struct owg_engine_job {
    double sample_rate, duration_s, volume, tremolo_depth, speaker_character;
    int32_t mlp_enabled;
    int32_t block_size; /* render() block length; voice freeing happens at block boundaries (engine.rs:461) */
    int32_t warm_up;    /* 1 = construct through set_sample_rate (0.6 s warm-up, engine.rs:261-286) */
    int32_t _pad0;
    const owg_event* ev; /* sorted by sample */
    int64_t n_ev;
} owg_engine_job;

/* ---- options -------------------------------------------------------------------------------- */
#define OWG_OUT_HOST 0   /* `out` is host memory (pinned recommended); D2H copy inside the call */
#define OWG_OUT_DEVICE 1 /* `out` is device memory on `device`; no copy.  The library works on opts->stream (or its own
                          * non-blocking stream): any work the caller queued on OTHER streams for that buffer must be complete. */

#define OWG_PRECISION_F64_EXACT 0 /* IEEE f64, no FMA contraction: op-for-op the reference's arithmetic */

#define OWG_PREAMP_MELANGE12 0 /* gen_preamp.rs 12-node DK solver (--features melange-preamp; the north-star path) */
#define OWG_PREAMP_LEGACY8 1   /* dk_preamp_legacy.rs hand-written 8-node solver: the reference's DEFAULT build
                                * (openwurli-dsp/Cargo.toml:10-19); chain B, preamp batch and metrics entry points */

#define OWG_POWER_AMP_BEHAVIORAL 0          /* power_amp.rs:167-275 (`legacy-power-amp`, the DEFAULT build): closed-loop NR approximation */
#define OWG_POWER_AMP_MELANGE 1             /* power_amp.rs:279-465 + gen_power_amp.rs (`--no-default-features`): melange 7-BJT class-AB
                                             * nodal solver with rail sag and the divergence guard */
#define OWG_POWER_AMP_MELANGE_IDEAL_RAILS 2 /* the same with set_rail_sag(false) (`preamp-bench render --no-rail-sag`, main.rs:481-483) */

typedef struct owg_opts {
    int32_t device;       /* CUDA device ordinal; -1 = current device */
    int32_t out_location; /* OWG_OUT_HOST | OWG_OUT_DEVICE */
    int32_t precision;    /* OWG_PRECISION_* */
    int32_t preamp_model; /* OWG_PREAMP_* */
    void* stream;         /* cudaStream_t to launch on; NULL = library-owned stream */
    int32_t collect_diag; /* 1 = keep per-call solver counters for owg_last_diag() */
    uint32_t device_mask; /* bit d = use CUDA device d.  0 = single device (`device`).  With two or more bits set, owg_render_voices and
                           * owg_render_bench (host output) fan the job list out over the selected GPUs inside the call: contiguous job ranges
                           * balanced by rendered samples, one worker thread, stream and staging buffer per GPU, every GPU copying its rows
                           * straight into the caller's buffer.  Renders are independent: no collective, each GPU recomputes the (tiny) shared
                           * sequences of the preamp groups it touches.  Order the jobs by preamp group to keep groups on one GPU. */
    int32_t power_amp_model; /* OWG_POWER_AMP_*; the melange models are served by owg_render_bench, owg_chain_batch, owg_render_midi and
                              * owg_power_amp_batch (chain B's output stage: PowerAmp::new() at 44.1 kHz, main.rs:480); every other entry
                              * point answers OWG_E_UNSUPPORTED for them */
    int32_t _reserved[5];
} owg_opts;

/* Solver counters of the last call on this thread that had collect_diag=1 (sums over all jobs).
 * Mirrors the reference's per-solver diag fields (gen_preamp.rs:1613-1632,1663). */
typedef struct owg_diag {
    uint64_t nr_iter_hist[16]; /* main preamp: histogram of last_nr_iterations (bucket 15 = >=15) */
    uint64_t nr_max_iter, be_fallback, voltage_damp, nan_reset;
    uint64_t shadow_nr_iter_hist[16];
    uint64_t shadow_be_fallback, shadow_nan_reset;
    uint64_t poweramp_iter_hist[9];
    uint64_t tremolo_nr_iter_hist[16];
    uint64_t tremolo_be_fallback;
    uint64_t kernels_launched; /* CUDA kernels this call launched */
} owg_diag;

/* ---- introspection -------------------------------------------------------------------------- */
int owg_abi_version(void);
int owg_device_count(void);          /* usable CUDA devices (0 when none: every render then fails) */
const char* owg_last_error(void);    /* thread-local text of the last failure */
void owg_default_opts(owg_opts* o);  /* device=-1, host output, f64 exact, melange12, no diag */

/* ---- one-shot batch renders ----------------------------------------------------------------- */

/* Batch of Voice::render_note_with_scale (voice.rs:201-221) / Voice::note_on + render (chain V:
 * reed + attack noise + pickup + post-pickup gain).  out is [n][stride] f64; job i writes
 * (uint64)(duration_s*sample_rate) samples at out + i*stride; in a ragged batch the shorter rows are zero-filled up to the
 * longest render of the call (columns beyond that are not touched).  The same holds for owg_render_bench. */
int owg_render_voices(const owg_voice_job* jobs, int64_t n, double* out, int64_t stride, const owg_opts* opts);

/* Batch of `preamp-bench render` (main.rs:371-496), chain B: voice -> [2x oversampler] ->
 * melange preamp (+ Twin-T/LDR tremolo) -> volume^2 -> power amp -> speaker -> POST_SPEAKER_GAIN.
 * f64 samples before WAV quantisation. */
int owg_render_bench(const owg_bench_job* jobs, int64_t n, double* out, int64_t stride, const owg_opts* opts);

/* Batch of WurliEngine streams (engine.rs), chain E; f32 out like WurliEngine::render. */
int owg_render_engines(const owg_engine_job* jobs, int64_t n, float* out, int64_t stride, const owg_opts* opts);

/* Preamp-only batch: n_inst input signals [n_inst][n_samp] (row stride in_stride) through
 * upsample_2x -> DkPreamp::process_sample x2 -> downsample_2x (`process_oversampled`,
 * main.rs:961-974), with Tremolo::new(depth, fs_preamp) driving set_ldr_resistance before every
 * preamp sample when tremolo_depth > 0 (main.rs:432-461), else reset()+set_ldr_resistance(r_ldr_static).
 * `in`/`out` follow opts->out_location (both host or both device). */
int owg_preamp_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, double fs_base,
                     int oversample, double tremolo_depth, double r_ldr_static, double* out, int64_t out_stride,
                     const owg_opts* opts);

/* The melange power amplifier alone: row i = PowerAmp::new_at_sample_rate(sample_rate) (power_amp.rs:326-349: the settled
 * CircuitState of gen_power_amp.rs re-rated with set_sample_rate), set_rail_sag(rail_sag != 0), y[t] = process(x[t]) (power_amp.rs:373-436:
 * gen_power_amp::process_sample (gen_power_amp.rs:8838) / HEADROOM, divergence guard with last-good hold, clamp to +-1, RailDynamics::step).
 * One 16-lane tile per row.  rails (optional) [n_inst][2] = rail_voltages() after the last sample; counters (optional) [n_inst][4] =
 * {divergence-guard resets, backward-Euler retries, NaN resets, last_nr_iterations}.  in / out follow opts->out_location. */
int owg_power_amp_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, double sample_rate, int32_t rail_sag, double* out,
                        int64_t out_stride, double* rails, uint32_t* counters, const owg_opts* opts);

/* Generic form of the above: rows of `in` (any mono signal at the instance's base rate, e.g. a sum of voices as in `preamp-bench
 * render-poly`, main.rs:1380-1560) through the full chain B of params[i] (sample_rate, r_ldr, tremolo_depth, volume,
 * speaker_character, no_preamp, no_poweramp; the voice fields are ignored).  init_order selects how the caller constructs the
 * static preamp: OWG_INIT_RESET_THEN_SET = `reset(); set_ldr_resistance(r)` (cmd_render, main.rs:438-439), OWG_INIT_SET_THEN_RESET =
 * `set_ldr_resistance(r); reset()` (render-poly / render-midi, main.rs:1463-1464, 1753-1754: the melange preamp falls back to its
 * settled 100 kOhm state, the legacy preamp re-solves its DC point at r). */
#define OWG_INIT_RESET_THEN_SET 0
#define OWG_INIT_SET_THEN_RESET 1
int owg_chain_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, const owg_bench_job* params, int32_t init_order,
                    double* out, int64_t out_stride, const owg_opts* opts);

/* `preamp-bench render-midi` (main.rs:1603-1895): the tool's own voice manager (64 slots, first free slot else the oldest, no
 * crossfade; note-off to the oldest sounding voice of that key; sustain pedal defers note-offs; silent voices are dropped before each
 * 64-sample chunk), voices summed in slot order, then chain B with the static preamp in SET_THEN_RESET order, f64 output.
 * Events carry absolute time in seconds exactly as the tool's SMF walk produces them and must be sorted by time (stable). */
#define OWG_MIDI_NOTE_ON 0
#define OWG_MIDI_NOTE_OFF 1
#define OWG_MIDI_PEDAL 2 /* velocity != 0 -> pedal down */
typedef struct owg_midi_event {
    double time_s;
    uint8_t kind, note, velocity, _pad0;
    int32_t _pad1;
} owg_midi_event;
typedef struct owg_midi_job {
    const owg_midi_event* ev;
    int64_t n_ev;
    int64_t n_samples;       /* total_samples = ((last_event_time + tail) * 44100) as usize */
    double volume;           /* --volume (0.60) */
    double speaker_character;/* --speaker (1.0) */
    int32_t no_poweramp;     /* --no-poweramp */
    int32_t _pad0;
} owg_midi_job;
int owg_render_midi(const owg_midi_job* jobs, int64_t n, double* out, int64_t stride, const owg_opts* opts);

/* Output mode "metrics" (BASELINE config 4): chain B with the `run_calibrate` analysis reduced on the device
 * instead of returning samples (tools/preamp-bench/src/main.rs:1139-1141, 1215-1223, 893-938): over the window
 * [window_start_s, window_end_s) of the final (T5) signal, per job:
 *   metrics[7*i + 0] peak_db   = 20 log10(max|x|)            (-120 below 1e-15)
 *   metrics[7*i + 1] rms_db    = 10 log10(mean x^2)          (-120 when 0)
 *   metrics[7*i + 2] h2_h1_db  = 20 log10(|X(2 f0)| / |X(f0)|), f0 = midi_to_freq(midi), single-bin DFT
 *   metrics[7*i + 3..6]        = raw peak, mean square, |X(f0)|, |X(2 f0)|
 * All jobs of one call must share sample_rate and cover the window.  `metrics` is host memory [n][7]. */
#define OWG_METRIC_COLUMNS 7
int owg_render_bench_metrics(const owg_bench_job* jobs, int64_t n, double window_start_s, double window_end_s, double* metrics,
                             const owg_opts* opts);

/* `preamp-bench calibrate` (tools/preamp-bench/src/main.rs:1069-1300): every column of its CSV, reduced on the device at the five tap
 * points T1 reed, T2 pickup, T3 voice (x output_scale), T4 preamp, T5 final, over [window_start_s, window_end_s).  `cfg` is the
 * CalibrationConfig the CLI builds (tables.rs:256-277; NULL = CalibrationConfig::default()); the jobs are built the way run_calibrate
 * builds its reed (OWG_VOICE_NO_ONSET, mlp_enabled 0, attack_noise 0, r_ldr 1e6, tremolo_depth 0).  rows is host memory
 * [n][OWG_CALIBRATE_COLUMNS]:
 *   0 ds_at_c4  1 ds_actual  2 y_peak  3 t2_peak_db  4 t2_rms_db  5 t2_h2_h1_db  6 t3_peak_db  7 t3_rms_db  8 t4_peak_db  9 t4_rms_db
 *   10 t4_h2_h1_db  11 t5_peak_db  12 t5_rms_db  13 t5_h2_h1_db  14 proxy_db  15 trim_db  16 proxy_error_db  17 tanh_compression_db */
typedef struct owg_calib_cfg {
    double ds_at_c4, ds_exponent, ds_clamp_lo, ds_clamp_hi, target_db, voicing_slope;
    int32_t zero_trim, _pad0;
} owg_calib_cfg;
void owg_default_calib_cfg(owg_calib_cfg* cfg); /* CalibrationConfig::default(): 0.85, 0.75, (0.02, 0.95), -35, -0.04, false */
#define OWG_CALIBRATE_COLUMNS 18
int owg_render_calibrate(const owg_bench_job* jobs, int64_t n, const owg_calib_cfg* cfg, double window_start_s, double window_end_s,
                         double* rows, const owg_opts* opts);

/* ---- planned renders: note-on parameterisation done once, inputs resident in HBM ---------- */
typedef struct owg_plan owg_plan;

/* Runs the host-side note-on setup (tables.rs / variation.rs / mlp_correction.rs / hammer.rs /
 * voice.rs:28-142 / reed.rs:108-182) for every job and uploads the per-voice init records. */
int owg_plan_bench(const owg_bench_job* jobs, int64_t n, const owg_opts* opts, owg_plan** plan);
int owg_plan_voices(const owg_voice_job* jobs, int64_t n, const owg_opts* opts, owg_plan** plan);
/* Executes the plan (any number of times; each execution is a full, independent render). */
int owg_plan_execute(owg_plan* plan, double* out, int64_t stride, int32_t out_location);
/* Samples job i produces; max over jobs when i < 0. */
int64_t owg_plan_samples(const owg_plan* plan, int64_t i);
/* Bytes of init records / tables the plan uploaded to the device (host->device traffic of the setup). */
int64_t owg_plan_h2d_bytes(const owg_plan* plan);
/* CUDA kernels one execution launches. */
int64_t owg_plan_kernel_launches(const owg_plan* plan);
/* Device time (ms) of the dominant per-instance kernel in the last execution (CUDA events on the
 * plan's stream), and of the whole execution. */
int owg_plan_last_timing(const owg_plan* plan, float* main_kernel_ms, float* total_ms);
void owg_plan_destroy(owg_plan* plan);

int owg_last_diag(owg_diag* out);

/* The library keeps, per device and for the life of the process, the settled solver states (the reference's OnceLock caches:
 * melange_adapter.rs:12, power_amp.rs:289) and a grow-only device staging buffer for host-output renders (as large as the largest
 * one-shot render so far: 8.6 GB for the C3 grid).  This call frees the staging buffers that no call is using (device = -1: on every
 * device) and returns the bytes released, so a co-resident framework gets the memory back; the next host-output render re-allocates.
 * The current CUDA device of the calling thread is left as it was. */
int64_t owg_release_caches(int32_t device);

/* ---- host-logic probes (no device needed) -------------------------------------------------- */
/* Flattened per-voice init record the kernels start from (what Voice::note_on leaves behind,
 * voice.rs:28-142 / reed.rs:108-182): out[0..41] = 7x{cos_inc, sin_inc, phase_inc, amplitude,
 * decay_mult, jitter_drift}; then jitter_revert, jitter_diffusion, onset_ramp_inc, onset_shape_exp,
 * pickup_beta, pickup_ds, post_pickup_gain, noise_amp, noise_decay, bq_b0, bq_b1, bq_b2, bq_a1,
 * bq_a2, onset_ramp_samples, n_samples, jitter_state, noise_rng, noise_remaining (61 doubles). */
#define OWG_VOICE_INIT_DOUBLES 61
int owg_host_voice_init(const owg_voice_job* job, double* out61);
/* Speaker / volume parameters of one bench job: volume, a2, a3, norm, thermal_coeff, thermal_alpha,
 * hpf b0 b1 b2 a1 a2, lpf b0 b1 b2 a1 a2, tanh flag, oversample flag (18 doubles). */
#define OWG_CHAIN_INIT_DOUBLES 18
int owg_host_chain_init(const owg_bench_job* job, double* out18);
/* Plan-time constants of the legacy 8-node preamp (dk_preamp_legacy.rs:269-412) at `preamp_sr`: S_base[64], A_neg_base[64],
 * 2w[8], S[:,FB][8], S[:,E1]-S[:,C1][8], S[:,E2]-S[:,C2][8], K[4], N_v S_fb[2], S_fb N_i[2], s_fb_fb, g_cin, gc_1pc, c_cin,
 * DC state at 1 MOhm (v[8], i_nl[2], v_nl[2], j_cin, cin_rhs_prev), g_ldr after set_ldr_resistance(r_static), 1/1e6. */
#define OWG_LEGACY_GROUP_DOUBLES 188
int owg_host_legacy_group(double preamp_sr, double r_static, double* out188);

/* ---- FP64 pipe micro-benchmark (roofline denominator; MEASURED_PEAKS.json has no FP64 entry) - */
/* Runs a register-resident stream of dependent-free DFMA (fma=1) or DADD+DMUL pairs (fma=0) on
 * every SM for about `ms_target` ms and returns the achieved rate in 1e12 FP64 pipe instr/s
 * (one DFMA = one instruction = 2 flop). */
int owg_fp64_peak(int32_t device, int32_t fma, float ms_target, double* tera_instr_per_s);

/* Device self-test: the library's shared-reciprocal division (recip_prepare/div_by, owg_device.cuh) against the
 * compiler's IEEE-754 f64 division on 303104*n_per_thread pseudo-random operand pairs; *mismatches must be 0. */
/* Diagnostic counters of the lane-tiled chain kernel, accumulated by calls made with collect_diag = 1 on the current device:
 * [0] DK-warp cycles waiting for input  [1] DK-warp cycles total  [2] I/O-warp cycles waiting for the preamp  [3] I/O-warp cycles
 * total  [4] Newton loop trips over DK warp-steps  [5] Newton iterations over instance-steps  [6] DK warp-steps  [7] instance-steps
 * [8] Newton iterations repeated by the generic (reference-order) code, counted per lane.
 * [9..16] DK-warp cycles by section (head, build_rhs, S*rhs, Newton junction row, Newton elimination, Newton step/votes, S_NI + guard
 * vote, adapter + state shift)  [17] Newton iterations of the 8-lane Twin-T oscillator that took the generic code.
 * reset != 0 clears them after reading. */
int owg_debug_counters(uint64_t* out, int32_t n, int32_t reset);

/* ---- alias_audit::analyze on the device (crates/openwurli-dsp/src/alias_audit.rs:163-282) ------------------------------------------
 * The click-band metrics of `preamp-bench alias-audit` / tests/alias_audit_regression.rs reduced where the stream was rendered:
 * steady-state tail of analyze_seconds, f0 refined on the reference's 0.1 Hz grid (+-5 Hz around nominal_f0), 12 harmonic single-bin
 * DFTs, plateau metric over H6..H11, 5-18 kHz band RMS through four RBJ biquads.  results is host memory [n_rows][OWG_ALIAS_COLUMNS]:
 *   0 f0_hz  1 h1_dbfs  2..13 harmonic_db  14..25 harmonic_dbc  26 max_step_up_db  27 max_step_up_from_harmonic  28 hf_band_dbc */
#define OWG_ALIAS_COLUMNS 29
#define OWG_ROWS_F64 0
#define OWG_ROWS_F32 1
/* rows: [n_rows][stride] samples of type row_dtype, in host or device memory as opts->out_location says; the last
 * floor(sample_rate * analyze_seconds) samples of the first n_samples of every row are analysed. */
int owg_alias_analyze(const void* rows, int32_t row_dtype, int64_t stride, int64_t n_rows, int64_t n_samples, double sample_rate,
                      double analyze_seconds, const double* nominal_f0, double* results, const owg_opts* opts);
/* owg_render_engines into a device buffer of the library, then owg_alias_analyze on samples [0, n_samples_i) of every stream, where
 * n_samples_i = floor(duration_s * sample_rate) (alias_audit::render_stimulus + analyze, alias_audit.rs:131-204, batched; the streams
 * never leave the GPU).  All jobs must share sample_rate and duration. */
int owg_render_engines_alias(const owg_engine_job* jobs, int64_t n, double analyze_seconds, const double* nominal_f0, double* results,
                             const owg_opts* opts);

int owg_selftest_division(int64_t n_per_thread, uint64_t seed, uint64_t* mismatches, uint64_t* tested);

#ifdef __cplusplus
}
#endif
#endif /* OWGPU_H */
