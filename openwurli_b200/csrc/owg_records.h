// Plain-data records handed from the host-side note-on setup to the CUDA kernels.
// One VoiceInit per voice = what the reference's Voice::note_on (voice.rs:28-142) leaves in
// ModalReed / AttackNoise / Pickup / Voice after construction (SURVEY.md 8(a) "per-voice init record").
#pragma once
#include <stdint.h>

#define OWG_NUM_MODES 7

struct OwgVoiceInit {
    // ModalReed (reed.rs:44-85) -- per-mode constants and initial OU drift
    double cos_inc[OWG_NUM_MODES];
    double sin_inc[OWG_NUM_MODES];
    double phase_inc[OWG_NUM_MODES];
    double amplitude[OWG_NUM_MODES];
    double decay_mult[OWG_NUM_MODES];
    double jitter_drift[OWG_NUM_MODES];
    double jitter_revert, jitter_diffusion;
    double onset_ramp_inc, onset_shape_exp;
    // Pickup (pickup.rs:88-94) and Voice (voice.rs:16)
    double pickup_beta, pickup_ds, post_pickup_gain;
    // AttackNoise (hammer.rs:108-117)
    double noise_amp, noise_decay;
    double bq_b0, bq_b1, bq_b2, bq_a1, bq_a2;
    double sample_rate;
    uint64_t onset_ramp_samples;
    uint64_t n_samples;      // (duration_s * sample_rate) truncated
    uint32_t jitter_state;   // LCG state after the 14 Box-Muller draws
    uint32_t noise_rng;      // AttackNoise LCG seed
    uint32_t noise_remaining;
    uint8_t midi;
    uint8_t _pad[3];
};

// Per-instance parameters of the shared mono chain in `preamp-bench render` order (chain B,
// main.rs:478-496): volume^2 -> PowerAmp -> Speaker -> POST_SPEAKER_GAIN.
struct OwgChainInit {
    double volume;          // job.volume (applied as x*volume*volume)
    // Speaker (speaker.rs:50-61) after set_character(): coefficients and biquads
    double spk_a2, spk_a3, spk_norm;  // norm = 1 + a2 + a3
    double spk_thermal_coeff, spk_thermal_alpha;
    double hpf_b0, hpf_b1, hpf_b2, hpf_a1, hpf_a2;
    double lpf_b0, lpf_b1, lpf_b2, lpf_a1, lpf_a2;
    int32_t spk_tanh;       // character >= 0.001
    int32_t group;          // preamp matrix group index
    int32_t no_preamp, no_poweramp;
    int32_t oversample;     // sample_rate < 88200
    int32_t pre_only;       // preamp-only harness (owg_preamp_batch): the row holds the input, the preamp output replaces it
};

// One preamp "group" = instances that share (preamp sample rate, LDR trajectory): they share the
// DK matrices S, K, S*N_i, a_neg and the zero-input shadow ("pump") sequence.
struct OwgPreampGroup {
    double preamp_sr;
    double r_static;         // clamped static R (tremolo_depth <= 0)
    double tremolo_depth;    // > 0: Twin-T/LDR trajectory
    int32_t use_defaults;    // 48 kHz and never dirtied: baked *_DEFAULT tables apply (gen_preamp.rs:1941-1955)
    int32_t dc_at_r;         // legacy model, `set_ldr_resistance(r); reset()`: both instances start from the DC point solved at r
    int64_t n_os;            // preamp-rate samples to produce (max over the group's instances)
};

// Matrix record layout (doubles) consumed by the DK step.
#define OWG_MAT_S 0        // S[12][12]
#define OWG_MAT_SNI 144    // S_NI[12][3]
#define OWG_MAT_K 180      // K[3][3]
#define OWG_MAT_AN66 189   // a_neg[6][6] (the only R-dependent a_neg entry)
#define OWG_MAT_STRIDE 190 // per preamp-rate sample in tremolo mode
#define OWG_AN_SPARSE 38   // structural non-zeros of a_neg used by build_rhs, row-major order of appearance

// ---- legacy 8-node preamp (dk_preamp_legacy.rs, the reference's default build; owg_opts.preamp_model = OWG_PREAMP_LEGACY8) ----
// One record per preamp group (doubles), computed on the host at plan time: every matrix is R_ldr-independent.
#define OWG_LG_S 0        // S_base[8][8] = (2C/T + G_base)^-1
#define OWG_LG_AN 64      // A_neg_base[8][8] = 2C/T - G_base
#define OWG_LG_W2 128     // 2 w
#define OWG_LG_SFB 136    // S_base[:,FB]
#define OWG_LG_D0 144     // S_base[:,EMIT1] - S_base[:,COLL1]
#define OWG_LG_D1 152     // S_base[:,EMIT2] - S_base[:,COLL2]
#define OWG_LG_K 160      // K[2][2]
#define OWG_LG_NVSFB 164  // N_v * S_base[:,FB]  (2)
#define OWG_LG_SFBNI 166  // S_base[FB,:] * N_i  (2)
#define OWG_LG_SFBFB 168
#define OWG_LG_GCIN 169
#define OWG_LG_GC1PC 170
#define OWG_LG_CCIN 171
#define OWG_LG_V0 172     // DkState::at_dc at R_ldr = 1 MOhm: v[8], i_nl[2], v_nl[2], j_cin, cin_rhs_prev
#define OWG_LG_INL0 180
#define OWG_LG_VNL0 182
#define OWG_LG_JCIN0 184
#define OWG_LG_CINPREV0 185
#define OWG_LG_GSTATIC 186  // g_ldr after `reset(); set_ldr_resistance(r)` (static groups)
#define OWG_LG_GINIT 187    // 1 / 1e6: g_ldr_prev seen by the first sample
#define OWG_LG_STRIDE 188

// ---- chain E (WurliEngine streams) -----------------------------------------------------------------------------------
struct DamperRow {  // ModalReed::start_damper (reed.rs:191-216) per MIDI key, computed on the host with glibc
    double rate[7], mult[7];
    double ramp_samples;
    int32_t enabled, _pad;
};

struct SpkUpdate {  // one entry of an engine's schedule
    // _pad == 0: one Speaker::update_coefficients (speaker.rs:89-101) on the host-simulated character ramp
    // _pad == 1: WurliEngine::set_volume(a2) (engine.rs:378-380) at the start of a block
    int64_t at;   // base-rate sample index (counted from the first render() sample incl. warm-up) at which it takes effect
    double a2, a3, norm, thermal_coeff;
    double hpf_b0, hpf_b1, hpf_b2, hpf_a1, hpf_a2, lpf_b0, lpf_b1, lpf_b2, lpf_a1, lpf_a2;
    int32_t tanh_on, _pad;
};

struct EngineDesc {  // one WurliEngine stream
    double sample_rate, volume_target;
    int64_t n_samples;       // rendered base-rate samples (after the warm-up)
    int64_t n_warm;          // warm-up base-rate samples (0 or floor(0.6*sr))
    int32_t block_size, oversample, group, spk_sched, n_spk_updates, ramp_samples;
    int64_t ev_begin, ev_end; // range in the event array
};

struct EngineEvent {  // host-prepared: NOTE_ON carries the index of its precomputed OwgVoiceInit
    int64_t sample;
    int32_t kind, note;
    int64_t vinit;
};

struct MidiEvent {  // render-midi: host-prepared; NOTE_ON carries the index of its OwgVoiceInit
    int64_t chunk;       // first 64-sample chunk whose start time is >= the event's time (main.rs:1790-1795)
    int32_t kind, note;
    int64_t vinit;
};

struct EngineGroup {
    double sample_rate;       // base rate
    double preamp_sr;
    double depth_target;      // set_tremolo_depth() target applied right after the warm-up
    int64_t n_warm_os, n_os; // preamp-rate samples: warm-up, then rendered
    int32_t oversample, ramp_samples, use_defaults, _pad;
    int32_t dep_ev_begin, dep_ev_end;  // this group's set_tremolo_depth events in the launch's DepthEv array
};

