// Host-side note-on parameterisation (see host_setup.cpp).
#pragma once
#include "../../include/owgpu.h"
#include "owg_records.h"

struct PaModel;  // owg_pa_core.h

namespace owg {
// Melange power amplifier (gen_power_amp.rs) at `sample_rate`: baked 88.2 kHz tables or set_sample_rate's rebuilt ones, device constants,
// rail-sag coefficients (host_pa_setup.cpp).
void pa_build_model(double sample_rate, PaModel* m);
// Voice::note_on(+overrides) -> flat init record (voice.rs:28-142).
void make_voice_init(const owg_voice_job& job, OwgVoiceInit* out);
// Speaker / volume parameters of one `preamp-bench render` job (main.rs:478-496).
void make_chain_init(const owg_bench_job& job, int group, OwgChainInit* out);
// Attack-noise raised-cosine fade-in table (hammer.rs:161-168), 16 entries.
void noise_fade_table(double* t16);
// ModalReed::start_damper (reed.rs:191-216) for every MIDI key at `sample_rate`: 128 rows.
void make_damper_rows(double sample_rate, DamperRow* rows128);
// Host simulation of the engine's speaker-character smoother + Speaker::set_character threshold logic
// (engine.rs:67-130, 436-439; speaker.rs:63-101): every coefficient update with the render() sample index it takes
// effect at (entry 0 = the constructor state, at = -1).  The target is applied after `n_warm` samples.
int make_speaker_schedule(double sample_rate, double character_target, int64_t n_warm, int64_t n_total, uint32_t ramp_samples,
                          SpkUpdate* out, int max_out);
// The same with automation: chr_at / chr_target = set_speaker_character calls (render() sample index incl. warm-up, sorted; the first
// is the construction-time target at n_warm), vol_at / vol_target = set_volume calls, merged into one schedule in time order.
struct AutoEvent { int64_t at; double target; };
int make_engine_schedule(double sample_rate, const AutoEvent* chr, int n_chr, const AutoEvent* vol, int n_vol, int64_t n_total, uint32_t ramp_samples,
                         SpkUpdate* out, int max_out);
// Legacy 8-node preamp (dk_preamp_legacy.rs:269-412): R_ldr-independent matrices, Sherman-Morrison vectors, Cin-R1 companion
// constants and the DC operating point at 1 MOhm for `preamp_sr`; rec = OWG_LG_STRIDE doubles (owg_records.h).
// `r_static`: the static LDR resistance handed to set_ldr_resistance after reset() (NaN = tremolo group).
// dc_at_r: `set_ldr_resistance(r_static); reset()` -- the DC operating point, g_ldr and g_ldr_prev are those of r_static.
void make_legacy_group(double preamp_sr, double r_static, double* rec, bool dc_at_r = false);
// CalibrationConfig-parametrised table functions (tables.rs:256-288, 578-616, 465-503) for `preamp-bench calibrate`, whose CLI
// defaults (ds_at_c4 0.75, upper clamp 0.82) differ from CalibrationConfig::default().
double calib_displacement_scale(int midi, const owg_calib_cfg& cfg);
double calib_output_scale(int midi, double velocity, const owg_calib_cfg& cfg);
double register_trim_db(int midi);
// 10^(-80/20): the is_silent threshold (reed.rs:310), through glibc pow like the reference.
double silent_threshold();
// tables::midi_to_freq (tables.rs:34-36)
double note_frequency(int midi);
// melange-primitives Biquad::lowpass / highpass (RBJ cookbook, normalised by a0): out5 = b0, b1, b2, a1, a2.  kind 0 = LP, 1 = HP.
void rbj_coefficients(int kind, double fc, double q, double fs, double* out5);
}  // namespace owg
