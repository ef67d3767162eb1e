// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity checker and CPU baseline), never shipped and
// never on the product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// C API over the restatement headers (ow_voice.hpp, ow_preamp.hpp, ow_tremolo.hpp, ow_chain.hpp,
// ow_engine.hpp).  Signatures mirror include/owgpu.h with the prefix owo_.
//
// PARITY PINNING: the reference (Rust) cannot be compiled here; this oracle is pinned by the
// reference's own known-answer tests and baked-constant identities only (tests/test_oracle_*.py).
#include "ow_engine.hpp"
#include "ow_power_amp_melange.hpp"
#include "../include/owgpu.h"
#include <thread>
#include <atomic>
#include <cstdio>
#include <chrono>

using namespace ow;

static BenchJob to_bench(const owg_bench_job& j) {
    BenchJob b;
    b.midi = j.v.midi; b.velocity = j.v.velocity; b.sample_rate = j.v.sample_rate; b.duration_s = j.v.duration_s;
    b.noise_seed = j.v.noise_seed; b.mlp_enabled = j.v.mlp_enabled != 0; b.ds_override = j.v.ds_override;
    b.attack_noise = j.v.attack_noise != 0; b.zero_onset = (j.v.flags & OWG_VOICE_NO_ONSET) != 0; b.r_ldr = j.r_ldr; b.tremolo_depth = j.tremolo_depth; b.volume = j.volume;
    b.speaker_character = j.speaker_character; b.no_preamp = j.no_preamp != 0; b.no_poweramp = j.no_poweramp != 0;
    return b;
}

template <class F>
static void parallel_for(int64_t n, int threads, F f) {
    if (threads <= 1 || n <= 1) { for (int64_t i = 0; i < n; i++) f(i); return; }
    std::atomic<int64_t> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&] { for (;;) { int64_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& t : th) t.join();
}

static thread_local owg_diag g_diag;

extern "C" {

int owo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---- chain V ------------------------------------------------------------------------------
int owo_render_voices(const owg_voice_job* jobs, int64_t n, double* out, int64_t stride, int threads) {
    if (!jobs || !out || n < 0) return OWG_E_BAD_ARG;
    parallel_for(n, threads, [&](int64_t i) {
        const owg_voice_job& j = jobs[i];
        Voice v;
        v.note_on(j.midi, j.velocity, j.sample_rate, j.noise_seed, j.mlp_enabled != 0, (j.flags & OWG_VOICE_NO_ONSET) != 0);
        if (j.ds_override == j.ds_override) v.pickup.displacement_scale = j.ds_override;
        if (!j.attack_noise) v.noise.disable();
        const size_t ns = (size_t)f64_as_u64(j.duration_s * j.sample_rate);
        double* o = out + i * stride;
        for (size_t off = 0; off < ns; off += 1024) v.render(o + off, std::min<size_t>(1024, ns - off));
    });
    return OWG_OK;
}

// ---- chain B ------------------------------------------------------------------------------
int owo_render_bench(const owg_bench_job* jobs, int64_t n, double* out, int64_t stride, int threads) {
    if (!jobs || !out || n < 0) return OWG_E_BAD_ARG;
    pre::settled_state();
    std::vector<ChainDiag> dgs((size_t)n);
    parallel_for(n, threads, [&](int64_t i) {
        std::vector<double> r = render_bench(to_bench(jobs[i]), nullptr, &dgs[i]);
        std::memcpy(out + i * stride, r.data(), r.size() * sizeof(double));
    });
    std::memset(&g_diag, 0, sizeof(g_diag));
    for (auto& d : dgs) {
        for (int b = 0; b < 16; b++) { g_diag.nr_iter_hist[b] += d.main.nr_iter_hist[b]; g_diag.shadow_nr_iter_hist[b] += d.shadow.nr_iter_hist[b]; g_diag.tremolo_nr_iter_hist[b] += d.trem_nr_hist[b]; }
        for (int b = 0; b < 9; b++) g_diag.poweramp_iter_hist[b] += d.pa_iter_hist[b];
        g_diag.nr_max_iter += d.main.nr_max_iter; g_diag.be_fallback += d.main.be_fallback; g_diag.voltage_damp += d.main.voltage_damp;
        g_diag.nan_reset += d.main.nan_reset; g_diag.shadow_be_fallback += d.shadow.be_fallback; g_diag.shadow_nan_reset += d.shadow.nan_reset;
        g_diag.tremolo_be_fallback += d.trem_be;
    }
    return OWG_OK;
}

// chain B with taps for one job: any of the tap pointers may be null.
// voice/preamp/final: n samples; r_ldr/shadow: n_os samples (2n when oversampled).
int owo_render_bench_taps(const owg_bench_job* job, double* fin, double* voice, double* preamp, double* r_ldr, double* shadow) {
    std::vector<double> tv, tp, tr, ts;
    Taps t; t.voice = &tv; t.preamp = &tp; t.r_ldr = &tr; t.shadow = &ts;
    std::vector<double> r = render_bench(to_bench(*job), &t, nullptr);
    if (fin) std::memcpy(fin, r.data(), r.size() * 8);
    if (voice) std::memcpy(voice, tv.data(), tv.size() * 8);
    if (preamp) std::memcpy(preamp, tp.data(), tp.size() * 8);
    if (r_ldr) std::memcpy(r_ldr, tr.data(), tr.size() * 8);
    if (shadow) std::memcpy(shadow, ts.data(), ts.size() * 8);
    return OWG_OK;
}

// chain B with the preamp implementation chosen as owg_opts.preamp_model does for the product (0 melange-12, 1 legacy-8)
int owo_render_bench_model(const owg_bench_job* jobs, int64_t n, double* out, int64_t stride, int threads, int preamp_model) {
    if (!jobs || !out || n < 0) return OWG_E_BAD_ARG;
    if (preamp_model == 0) return owo_render_bench(jobs, n, out, stride, threads);
    std::vector<ChainDiag> dgs((size_t)n);
    parallel_for(n, threads, [&](int64_t i) {
        BenchJob b = to_bench(jobs[i]);
        b.preamp_model = preamp_model;
        std::vector<double> r = render_bench(b, nullptr, &dgs[i]);
        std::memcpy(out + i * stride, r.data(), r.size() * sizeof(double));
    });
    std::memset(&g_diag, 0, sizeof(g_diag));
    for (auto& d : dgs) {
        for (int b = 0; b < 16; b++) { g_diag.nr_iter_hist[b] += d.main.nr_iter_hist[b]; g_diag.tremolo_nr_iter_hist[b] += d.trem_nr_hist[b]; }
        for (int b = 0; b < 9; b++) g_diag.poweramp_iter_hist[b] += d.pa_iter_hist[b];
        g_diag.nan_reset += d.main.nan_reset;
        g_diag.tremolo_be_fallback += d.trem_be;
    }
    return OWG_OK;
}

int owo_preamp_batch_model(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, double fs_base, int oversample,
                           double tremolo_depth, double r_ldr_static, double* out, int64_t out_stride, int threads, int preamp_model) {
    if (!in || !out || n_inst < 0 || n_samp < 0) return OWG_E_BAD_ARG;
    if (preamp_model == 0) pre::settled_state();
    parallel_for(n_inst, threads, [&](int64_t i) {
        preamp_batch_one(in + i * in_stride, (size_t)n_samp, fs_base, oversample != 0, tremolo_depth, r_ldr_static, out + i * out_stride, preamp_model);
    });
    return OWG_OK;
}

// ---- legacy 8-node preamp probes (dk_preamp_legacy.rs tests) -----------------------------------
// out[0..7] = v_dc, out[8..9] = v_nl, out[10] = max |S_base * A_base - I| is not available (A is not kept): instead
// out[10] = s_fb_fb, out[11] = g_cin, out[12] = c_cin, out[13..16] = K
int owo_legacy_dc(double sample_rate, double* out) {
    leg::DkPreamp p(sample_rate);
    for (int i = 0; i < 8; i++) out[i] = p.v_dc[i];
    out[8] = p.main_st.v_nl[0]; out[9] = p.main_st.v_nl[1];
    out[10] = p.s_fb_fb; out[11] = p.g_cin; out[12] = p.c_cin;
    out[13] = p.k[0][0]; out[14] = p.k[0][1]; out[15] = p.k[1][0]; out[16] = p.k[1][1];
    return OWG_OK;
}
// Same layout as owg_host_legacy_group (include/owgpu.h): the product's plan-time constants, from the oracle's DkPreamp.
int owo_legacy_group(double sample_rate, double r_static, double* o) {
    leg::DkPreamp p(sample_rate);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) { o[i * 8 + j] = p.s_base[i][j]; o[64 + i * 8 + j] = p.a_neg_base[i][j]; }
    for (int i = 0; i < 8; i++) {
        o[128 + i] = p.two_w[i]; o[136 + i] = p.s_fb_col[i];
        o[144 + i] = p.s_base[i][leg::EMIT1] - p.s_base[i][leg::COLL1];
        o[152 + i] = p.s_base[i][leg::EMIT2] - p.s_base[i][leg::COLL2];
        o[172 + i] = p.main_st.v[i];
    }
    o[160] = p.k[0][0]; o[161] = p.k[0][1]; o[162] = p.k[1][0]; o[163] = p.k[1][1];
    o[164] = p.nv_sfb[0]; o[165] = p.nv_sfb[1]; o[166] = p.sfb_ni[0]; o[167] = p.sfb_ni[1];
    o[168] = p.s_fb_fb; o[169] = p.g_cin; o[170] = p.gc_1pc; o[171] = p.c_cin;
    o[180] = p.main_st.i_nl[0]; o[181] = p.main_st.i_nl[1]; o[182] = p.main_st.v_nl[0]; o[183] = p.main_st.v_nl[1];
    o[184] = p.main_st.j_cin; o[185] = p.main_st.cin_rhs_prev;
    o[187] = p.g_ldr;
    p.reset();
    if (r_static == r_static) p.set_ldr_resistance(r_static);
    o[186] = p.g_ldr;
    return OWG_OK;
}
// Raw preamp run (no oversampler): x[n] -> y[n] at `sample_rate`, R_ldr set once before the first sample;
// do_reset != 0 calls reset() first (measure_gain, :878-899); pump (shadow output) optional.
int owo_legacy_run(double sample_rate, double r_ldr, int do_reset, const double* x, int64_t n, double* y, double* pump) {
    leg::DkPreamp p(sample_rate);
    p.set_ldr_resistance(r_ldr);
    if (do_reset) p.reset();
    for (int64_t i = 0; i < n; i++) { double pu = 0.0; y[i] = p.process_sample(x[i], &pu); if (pump) pump[i] = pu; }
    return OWG_OK;
}
// test_idle_pump_level (:1966-2025): Tremolo::new(depth, sr) drives set_ldr_resistance, zero input; y = main - shadow.
int owo_legacy_idle_pump(double sample_rate, double depth, int64_t n, double* y, double* pump) {
    leg::DkPreamp p(sample_rate);
    Tremolo t(depth, sample_rate);
    for (int64_t i = 0; i < n; i++) { p.set_ldr_resistance(t.process()); double pu = 0.0; y[i] = p.process_sample(0.0, &pu); if (pump) pump[i] = pu; }
    return OWG_OK;
}

// chain B over caller-supplied rows (owg_chain_batch)
int owo_chain_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, const owg_bench_job* params, int init_order,
                    double* out, int64_t out_stride, int threads, int preamp_model) {
    if (!in || !out || !params) return OWG_E_BAD_ARG;
    if (preamp_model == 0) pre::settled_state();
    parallel_for(n_inst, threads, [&](int64_t i) {
        BenchJob b = to_bench(params[i]);
        b.preamp_model = preamp_model;
        std::vector<double> r = chain_rows(in + i * in_stride, (size_t)n_samp, b, init_order == 1);
        std::memcpy(out + i * out_stride, r.data(), r.size() * sizeof(double));
    });
    return OWG_OK;
}

// render-midi (owg_render_midi): one job
int owo_render_midi(const owg_midi_event* ev, int64_t n_ev, int64_t n_samples, double volume, double speaker, int no_poweramp, int preamp_model,
                    double* out, uint64_t* counters2) {
    std::vector<MidiEvt> e((size_t)n_ev);
    for (int64_t k = 0; k < n_ev; k++) { e[k].time_s = ev[k].time_s; e[k].kind = ev[k].kind; e[k].note = ev[k].note; e[k].velocity = ev[k].velocity; }
    if (preamp_model == 0) pre::settled_state();
    uint64_t a = 0, b = 0;
    std::vector<double> r = render_midi(e, (size_t)n_samples, volume, speaker, no_poweramp != 0, preamp_model, &a, &b);
    std::memcpy(out, r.data(), r.size() * sizeof(double));
    if (counters2) { counters2[0] = a; counters2[1] = b; }
    return OWG_OK;
}

// `preamp-bench calibrate`: rows[n_notes * n_vels][18] (include/owgpu.h owg_render_calibrate columns), cfg6 = ds_at_c4, ds_exponent,
// ds_clamp_lo, ds_clamp_hi, target_db, voicing_slope
int owo_calibrate_rows(const uint8_t* notes, int n_notes, const uint8_t* vels, int n_vels, const double* cfg6, int zero_trim, double volume,
                       double speaker, int preamp_model, double* rows, int threads) {
    CalibrationConfig cfg;
    cfg.ds_at_c4 = cfg6[0]; cfg.ds_exponent = cfg6[1]; cfg.ds_clamp_lo = cfg6[2]; cfg.ds_clamp_hi = cfg6[3]; cfg.target_db = cfg6[4];
    cfg.voicing_slope = cfg6[5]; cfg.zero_trim = zero_trim != 0;
    if (preamp_model == 0) pre::settled_state();
    parallel_for((int64_t)n_notes * n_vels, threads, [&](int64_t k) {
        const CalibrateRow r = calibrate_row(notes[k / n_vels], vels[k % n_vels], cfg, volume, speaker, preamp_model);
        std::memcpy(rows + k * 18, r.v, sizeof(r.v));
    });
    return OWG_OK;
}

int owo_last_diag(owg_diag* out) { if (!out) return OWG_E_BAD_ARG; *out = g_diag; return OWG_OK; }

// ---- preamp-only batch (C2) ------------------------------------------------------------------
int owo_preamp_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, double fs_base, int oversample,
                     double tremolo_depth, double r_ldr_static, double* out, int64_t out_stride, int threads) {
    if (!in || !out || n_inst < 0 || n_samp < 0) return OWG_E_BAD_ARG;
    pre::settled_state();
    parallel_for(n_inst, threads, [&](int64_t i) {
        preamp_batch_one(in + i * in_stride, (size_t)n_samp, fs_base, oversample != 0, tremolo_depth, r_ldr_static, out + i * out_stride);
    });
    return OWG_OK;
}

// Preamp-only batch that also gathers the melange preamp's guard counters (Newton histogram, max-iteration exits, BE fallbacks,
// voltage damping, NaN resets) into owo_last_diag: the parity tests drive those guards with out-of-range inputs.
int owo_preamp_batch_diag(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, double fs_base, int oversample,
                          double tremolo_depth, double r_ldr_static, double* out, int64_t out_stride, int threads) {
    if (!in || !out || n_inst < 0 || n_samp < 0) return OWG_E_BAD_ARG;
    pre::settled_state();
    std::vector<ChainDiag> dgs((size_t)n_inst);
    parallel_for(n_inst, threads, [&](int64_t i) {
        preamp_batch_one(in + i * in_stride, (size_t)n_samp, fs_base, oversample != 0, tremolo_depth, r_ldr_static, out + i * out_stride, 0, &dgs[i]);
    });
    std::memset(&g_diag, 0, sizeof(g_diag));
    for (auto& d : dgs) {
        for (int b = 0; b < 16; b++) { g_diag.nr_iter_hist[b] += d.main.nr_iter_hist[b]; g_diag.shadow_nr_iter_hist[b] += d.shadow.nr_iter_hist[b]; }
        g_diag.nr_max_iter += d.main.nr_max_iter; g_diag.be_fallback += d.main.be_fallback; g_diag.voltage_damp += d.main.voltage_damp;
        g_diag.nan_reset += d.main.nan_reset; g_diag.shadow_be_fallback += d.shadow.be_fallback; g_diag.shadow_nan_reset += d.shadow.nan_reset;
    }
    return OWG_OK;
}

// ---- chain E ----------------------------------------------------------------------------------
static void run_engine(const owg_engine_job& j, float* out, int preamp_model = 0) {
    WurliEngine eng(j.sample_rate, preamp_model);
    // The plugin constructs at a nominal rate and then calls set_sample_rate (lib.rs:96-97), which warms up.
    if (j.warm_up) eng.set_sample_rate(j.sample_rate);
    eng.volume.set_target(j.volume);
    eng.tremolo_depth.set_target(j.tremolo_depth);
    eng.speaker_character.set_target(j.speaker_character);
    eng.mlp_enabled = j.mlp_enabled != 0;
    const int64_t total = (int64_t)f64_as_u64(j.sample_rate * j.duration_s);
    const int64_t bs = j.block_size > 0 ? j.block_size : 512;
    int64_t e = 0;
    for (int64_t pos = 0; pos < total; pos += bs) {
        const int64_t len = std::min<int64_t>(bs, total - pos);
        while (e < j.n_ev && j.ev[e].sample < pos + len) {  // events are applied at the start of the block they fall in
            const owg_event& ev = j.ev[e++];
            if (ev.kind == OWG_EV_NOTE_ON) eng.note_on(ev.note, ev.velocity);
            else if (ev.kind == OWG_EV_NOTE_OFF) eng.note_off(ev.note);
            else if (ev.kind == OWG_EV_SUSTAIN) eng.set_sustain(ev.note != 0);
            else if (ev.kind == OWG_EV_SET_VOLUME) eng.volume.set_target((double)ev.velocity);                        // engine.rs:378-380
            else if (ev.kind == OWG_EV_SET_TREMOLO_DEPTH) eng.tremolo_depth.set_target((double)ev.velocity);          // :382-384
            else if (ev.kind == OWG_EV_SET_SPEAKER_CHARACTER) eng.speaker_character.set_target((double)ev.velocity);  // :386-388
        }
        eng.render(out + pos, (size_t)len);
    }
}
int owo_render_engines(const owg_engine_job* jobs, int64_t n, float* out, int64_t stride, int threads) {
    if (!jobs || !out || n < 0) return OWG_E_BAD_ARG;
    pre::settled_state();
    parallel_for(n, threads, [&](int64_t i) { run_engine(jobs[i], out + i * stride); });
    return OWG_OK;
}

int owo_render_engines_model(const owg_engine_job* jobs, int64_t n, float* out, int64_t stride, int threads, int preamp_model) {
    if (!jobs || !out || n < 0) return OWG_E_BAD_ARG;
    if (preamp_model == 0) pre::settled_state();
    parallel_for(n, threads, [&](int64_t i) { run_engine(jobs[i], out + i * stride, preamp_model); });
    return OWG_OK;
}

// alias_audit::render_stimulus (alias_audit.rs:131-161): engine without warm_up, 6x1024 settle blocks, then note.
int owo_alias_stimulus(uint8_t note, uint8_t velocity, double sr, double seconds, double volume, double* out) {
    WurliEngine eng(sr);
    eng.ensure_buffer_capacity(1024);
    eng.volume.set_target(volume);
    eng.tremolo_depth.set_target(0.0);
    eng.speaker_character.set_target(0.0);
    eng.mlp_enabled = true;
    float buf[1024];
    for (int i = 0; i < 6; i++) eng.render(buf, 1024);
    eng.note_on(note, (float)velocity / 127.0f);
    const size_t total = (size_t)f64_as_u64(sr * seconds);
    for (size_t pos = 0; pos < total; pos += 1024) {
        const size_t len = std::min<size_t>(1024, total - pos);
        eng.render(buf, len);
        for (size_t k = 0; k < len; k++) out[pos + k] = (double)buf[k];
    }
    return OWG_OK;
}

// alias_audit::analyze (alias_audit.rs:163-282) on a rendered signal: steady-state tail of `analyze_seconds`, f0 refinement on a
// 0.1 Hz grid, 12 harmonic single-bin DFTs, plateau metric over H6..H11, 5-18 kHz band RMS through four RBJ biquads.
// out[29] = f0_hz, h1_dbfs, harmonic_db[12], harmonic_dbc[12], max_step_up_db, max_step_up_from_harmonic, hf_band_dbc
int owo_alias_analyze(const double* signal, int64_t n, double sr, double analyze_seconds, double nominal_f0, double* out) {
    const int64_t analyze_n = (int64_t)f64_as_u64(sr * analyze_seconds);
    if (!signal || !out || n < analyze_n || analyze_n <= 0) return OWG_E_BAD_ARG;
    const double* tail = signal + (n - analyze_n);
    auto dft_magnitude = [&](double freq) {  // :225-236
        const double nn = (double)analyze_n;
        double re = 0.0, im = 0.0;
        const double omega = 2.0 * PI * freq / sr;
        for (int64_t i = 0; i < analyze_n; i++) {
            const double phase = omega * (double)i;
            re += tail[i] * std::cos(phase);
            im -= tail[i] * std::sin(phase);
        }
        return 2.0 * std::sqrt((re / nn) * (re / nn) + (im / nn) * (im / nn));
    };
    auto mag_to_db = [](double mag) { return mag > 0.0 ? 20.0 * std::log10(mag) : -200.0; };
    double best_f = nominal_f0, best_mag = dft_magnitude(nominal_f0);  // refine_f0 :248-261
    for (double f = nominal_f0 - 5.0; f <= nominal_f0 + 5.0; f += 0.1) {
        const double mag = dft_magnitude(f);
        if (mag > best_mag) { best_mag = mag; best_f = f; }
    }
    const double f0 = best_f;
    const double h1 = dft_magnitude(f0);
    double* hdb = out + 2;
    double* hdbc = out + 14;
    for (int k = 0; k < 12; k++) {
        const double mag = dft_magnitude((double)(k + 1) * f0);
        hdb[k] = mag_to_db(mag);
        hdbc[k] = h1 > 0.0 ? 20.0 * std::log10(mag / h1) : -200.0;
    }
    hdbc[0] = 0.0;
    double worst = -INFINITY;
    int worst_from = 6;
    for (int i = 5; i < 10; i++) {  // plateau_metric :206-222
        const double delta = hdbc[i + 1] - hdbc[i];
        if (delta > worst) { worst = delta; worst_from = i + 1; }
    }
    Biquad hp1, hp2, lp1, lp2;  // bandpass_rms :266-277
    hp1.set(Biquad::HP, 5000.0, 0.70710678118654752440, sr); hp2 = hp1;
    lp1.set(Biquad::LP, 18000.0, 0.70710678118654752440, sr); lp2 = lp1;
    double sum_sq = 0.0;
    for (int64_t i = 0; i < analyze_n; i++) {
        const double y = lp2.process(lp1.process(hp2.process(hp1.process(tail[i]))));
        sum_sq += y * y;
    }
    const double hf_rms = std::sqrt(sum_sq / (double)analyze_n);
    out[0] = f0; out[1] = mag_to_db(h1);
    out[26] = worst; out[27] = (double)worst_from;
    out[28] = h1 > 0.0 ? 20.0 * std::log10(hf_rms / h1) : -200.0;
    return OWG_OK;
}

// ---- melange power amplifier + rail sag (power_amp.rs; SURVEY 8(f) #4: oracle side only) --------------------------------------------
int owo_have_melange_power_amp() {
#ifdef OW_HAVE_MELANGE_POWER_AMP
    return 1;
#else
    return 0;
#endif
}
// RailDynamics alone: v_out[n] -> (v_rail_pos, v_rail_neg) after each step; rails[2n]
int owo_rail_dynamics(double sample_rate, const double* v_out, int64_t n, double* rails) {
    ow::pam::RailDynamics r(sample_rate);
    for (int64_t i = 0; i < n; i++) { r.step(v_out[i]); rails[2 * i] = r.v_rail_pos; rails[2 * i + 1] = r.v_rail_neg; }
    return OWG_OK;
}
// PowerAmp::new_at_sample_rate(sr); set_rail_sag(rail_sag); y[i] = process(x[i]); rails (optional) = rail_voltages() after each sample;
// toggle_off_at >= 0: set_rail_sag(false) before that sample.  Returns the number of divergence-guard resets, or < 0.
int64_t owo_power_amp_melange(double sample_rate, int rail_sag, const double* x, int64_t n, double* y, double* rails, int64_t toggle_off_at) {
#ifdef OW_HAVE_MELANGE_POWER_AMP
    ow::pam::PowerAmp pa(sample_rate);
    pa.set_rail_sag(rail_sag != 0);
    for (int64_t i = 0; i < n; i++) {
        if (i == toggle_off_at) pa.set_rail_sag(false);
        y[i] = pa.process(x[i]);
        if (rails) pa.rail_voltages(rails[2 * i], rails[2 * i + 1]);
    }
    return (int64_t)pa.resets;
#else
    (void)sample_rate; (void)rail_sag; (void)x; (void)n; (void)y; (void)rails; (void)toggle_off_at;
    return -1;
#endif
}

// chain B's output stage with the melange amplifier (tools/preamp-bench/src/main.rs:478-496 built with --no-default-features):
// attenuated = pre * volume * volume; PowerAmp::new() (44.1 kHz) with set_rail_sag(!no_rail_sag); Speaker::new(sr), set_character; * POST_SPEAKER_GAIN
int owo_output_stage_melange(const double* pre, int64_t n, double sample_rate, double volume, double speaker_character, int no_poweramp, int rail_sag,
                             double* out) {
#ifdef OW_HAVE_MELANGE_POWER_AMP
    ow::pam::PowerAmp pa;
    pa.set_rail_sag(rail_sag != 0);
    Speaker spk(sample_rate);
    spk.set_character(speaker_character);
    for (int64_t i = 0; i < n; i++) {
        const double att = pre[i] * volume * volume;
        const double amp = no_poweramp ? att : pa.process(att);
        out[i] = spk.process(amp) * POST_SPEAKER_GAIN;
    }
    return OWG_OK;
#else
    (void)pre; (void)n; (void)sample_rate; (void)volume; (void)speaker_character; (void)no_poweramp; (void)rail_sag; (void)out;
    return -1;
#endif
}
// CircuitState::default() + n_extra x process_sample(0.0) (n_extra = 44100: compute_settled_state, power_amp.rs:291-296):
// out54 = v_prev[20] i_nl_prev[16] i_nl_prev_prev[16] dc_block_x_prev dc_block_y_prev
int owo_power_amp_settled(int64_t n_extra, double* out54) {
#ifdef OW_HAVE_MELANGE_POWER_AMP
    gen_power_amp::CircuitState st = gen_power_amp::CircuitState::default_();
    for (int64_t i = 0; i < n_extra; i++) gen_power_amp::process_sample(0.0, st);
    double* o = out54;
    for (double v : st.v_prev) *o++ = v;
    for (double v : st.i_nl_prev) *o++ = v;
    for (double v : st.i_nl_prev_prev) *o++ = v;
    *o++ = st.dc_block_x_prev[0]; *o++ = st.dc_block_y_prev[0];
    return OWG_OK;
#else
    (void)n_extra; (void)out54;
    return -1;
#endif
}
// CircuitState matrices after PowerAmp::new_at_sample_rate(sr): out = s[400] k[256] s_ni[320] s_be[400] k_be[256] s_ni_be[320] a_neg_be[400] dc_block_r
int owo_power_amp_matrices(double sample_rate, double* out) {
#ifdef OW_HAVE_MELANGE_POWER_AMP
    ow::pam::PowerAmp pa(sample_rate);
    double* o = out;
    for (auto& r : pa.state.s) for (double v : r) *o++ = v;
    for (auto& r : pa.state.k) for (double v : r) *o++ = v;
    for (auto& r : pa.state.s_ni) for (double v : r) *o++ = v;
    for (auto& r : pa.state.s_be) for (double v : r) *o++ = v;
    for (auto& r : pa.state.k_be) for (double v : r) *o++ = v;
    for (auto& r : pa.state.s_ni_be) for (double v : r) *o++ = v;
    for (auto& r : pa.state.a_neg_be) for (double v : r) *o++ = v;
    *o = pa.state.dc_block_r;
    return OWG_OK;
#else
    (void)sample_rate; (void)out;
    return -1;
#endif
}

// ---- known-answer probes (host-side setup functions) --------------------------------------------
double owo_midi_to_freq(int midi) { return midi_to_freq((uint8_t)midi); }
double owo_tip_mass_ratio(int midi) { return tip_mass_ratio((uint8_t)midi); }
void owo_mode_ratios(double mu, double* out7) { mode_ratios(mu, out7); }
double owo_reed_length_mm(int midi) { return reed_length_mm((uint8_t)midi); }
double owo_reed_compliance(int midi) { return reed_compliance((uint8_t)midi); }
double owo_pickup_displacement_scale(int midi) { return pickup_displacement_scale((uint8_t)midi); }
void owo_spatial_coupling(double mu, double len_mm, double* out7) { spatial_coupling_coefficients(mu, len_mm, out7); }
double owo_fundamental_decay_rate(int midi) { return fundamental_decay_rate((uint8_t)midi); }
double owo_output_scale(int midi, double vel) { return output_scale((uint8_t)midi, vel); }
double owo_velocity_exponent(int midi) { return velocity_exponent((uint8_t)midi); }
double owo_velocity_scurve(double v) { return velocity_scurve(v); }
double owo_register_trim_db(int midi) { return register_trim_db((uint8_t)midi); }
double owo_pickup_rms_proxy(double ds, double f0, double fc) { return pickup_rms_proxy(ds, f0, fc); }
double owo_freq_detune(int midi) { return freq_detune((uint8_t)midi); }
void owo_mode_amplitude_offsets(int midi, double* out7) { mode_amplitude_offsets((uint8_t)midi, out7); }
double owo_dwell_time(double v, double f0) { return dwell_time(v, f0); }
double owo_onset_ramp_time(double v, double f0) { return onset_ramp_time(v, f0); }
void owo_dwell_attenuation(double v, double f0, const double* ratios7, double* out7) { dwell_attenuation(v, f0, ratios7, out7); }
void owo_mlp_infer(int midi, double vel, double* out11) {
    MlpCorrections c = mlp_infer((uint8_t)midi, vel);
    for (int i = 0; i < 5; i++) { out11[i] = c.freq_offsets_cents[i]; out11[5 + i] = c.decay_offsets[i]; }
    out11[10] = c.ds_correction;
}
double owo_pickup_soft_saturate(double y) { return Pickup::soft_saturate(y); }
double owo_fast_exp(double x) { return pre::fast_exp(x); }
void owo_note_params(int midi, double* out22) {
    NoteParams p = note_params((uint8_t)midi);
    out22[0] = p.fundamental_hz;
    for (int i = 0; i < 7; i++) { out22[1 + i] = p.mode_ratios[i]; out22[8 + i] = p.mode_amplitudes[i]; out22[15 + i] = p.mode_decay_rates[i]; }
}
// biquad magnitude probe: run a sine through a band-pass and return peak of the tail (filters.rs:66-99)
double owo_biquad_bp_gain(double fc, double q, double fs, double f_test, int n) {
    Biquad b; b.set(Biquad::BP, fc, q, fs);
    double peak = 0.0;
    for (int i = 0; i < n; i++) {
        double y = b.process(std::sin(TAU * f_test * (double)i / fs));
        if (i > n / 2 && std::fabs(y) > peak) peak = std::fabs(y);
    }
    return peak;
}

// ---- preamp / tremolo probes ----------------------------------------------------------------------
// Rebuild matrices at (sample_rate, r) and export S (144), K (9), S_NI (36), a_neg (144).
void owo_preamp_matrices(double sample_rate, double r, double* s144, double* k9, double* sni36, double* aneg144) {
    pre::CircuitState st; st.set_default();
    st.current_sample_rate = sample_rate; st.pot_0_resistance = r;
    st.rebuild_matrices();
    std::memcpy(s144, st.s, sizeof(st.s)); std::memcpy(k9, st.k, sizeof(st.k));
    std::memcpy(sni36, st.s_ni, sizeof(st.s_ni)); std::memcpy(aneg144, st.a_neg, sizeof(st.a_neg));
}
// Settled preamp state (melange_adapter.rs:14-20): v_prev(12) i_nl_prev(3) i_nl_prev_prev(3) input_prev(1)
void owo_preamp_settled(double* out19) {
    const pre::CircuitState& s = pre::settled_state();
    std::memcpy(out19, s.v_prev, 96); std::memcpy(out19 + 12, s.i_nl_prev, 24); std::memcpy(out19 + 15, s.i_nl_prev_prev, 24); out19[18] = s.input_prev;
}
// DkPreamp driven by an input signal at preamp rate with static R (reset(); set_ldr_resistance(r)) -- gain tests.
void owo_preamp_run(double preamp_sr, double r_ldr, const double* in, int64_t n, double* out) {
    pre::DkPreamp p(preamp_sr);
    p.reset(); p.set_ldr_resistance(r_ldr);
    for (int64_t i = 0; i < n; i++) out[i] = p.process_sample(in[i]);
}
// Tremolo::new(depth, sr) then n process() calls -> shunt R; also raw oscillator volts when v_out != null.
void owo_tremolo_run(double depth, double sr, int64_t n, double* r_out) {
    Tremolo t(depth, sr);
    for (int64_t i = 0; i < n; i++) r_out[i] = t.process();
}
void owo_tremolo_osc(double sr, int64_t n_settle, int64_t n, double* v_out, double* state_out18) {
    trm::CircuitState s; s.set_default();
    if (std::fabs(sr - trm::SAMPLE_RATE) > 0.5) s.set_sample_rate(sr);
    for (int64_t i = 0; i < n_settle; i++) trm::process_sample(0.0, s);
    for (int64_t i = 0; i < n; i++) v_out[i] = trm::process_sample(0.0, s);
    if (state_out18) { std::memcpy(state_out18, s.v_prev, 56); std::memcpy(state_out18 + 7, s.i_nl_prev, 32); std::memcpy(state_out18 + 11, s.i_nl_prev_prev, 32); }
}
double owo_poweramp(double x) { PowerAmp p; return p.process(x); }
void owo_speaker_run(double sr, double character, const double* in, int64_t n, double* out) {
    Speaker s(sr); s.set_character(character);
    for (int64_t i = 0; i < n; i++) out[i] = s.process(in[i]);
}
void owo_oversampler_roundtrip(const double* in, int64_t n, double* out) {
    Oversampler os;
    for (int64_t i = 0; i < n; i++) { double a, b; os.up1(in[i], a, b); out[i] = os.down1(a, b); }
}
// Flattened Voice::note_on state, same layout as owg_host_voice_init (include/owgpu.h).
void owo_voice_init(const owg_voice_job* j, double* o) {
    Voice v;
    v.note_on(j->midi, j->velocity, j->sample_rate, j->noise_seed, j->mlp_enabled != 0, (j->flags & OWG_VOICE_NO_ONSET) != 0);
    if (j->ds_override == j->ds_override) v.pickup.displacement_scale = j->ds_override;
    if (!j->attack_noise) v.noise.disable();
    int k = 0;
    for (int m = 0; m < 7; m++) { const Mode& q = v.reed.modes[m]; o[k++] = q.cos_inc; o[k++] = q.sin_inc; o[k++] = q.phase_inc; o[k++] = q.amplitude; o[k++] = q.decay_mult; o[k++] = q.jitter_drift; }
    o[k++] = v.reed.jitter_revert; o[k++] = v.reed.jitter_diffusion; o[k++] = v.reed.onset_ramp_inc; o[k++] = v.reed.onset_shape_exp;
    o[k++] = v.pickup.beta; o[k++] = v.pickup.displacement_scale; o[k++] = v.post_pickup_gain; o[k++] = v.noise.amplitude; o[k++] = v.noise.decay_per_sample;
    o[k++] = v.noise.bpf.b0; o[k++] = v.noise.bpf.b1; o[k++] = v.noise.bpf.b2; o[k++] = v.noise.bpf.a1; o[k++] = v.noise.bpf.a2;
    o[k++] = (double)v.reed.onset_ramp_samples; o[k++] = (double)f64_as_u64(j->duration_s * j->sample_rate); o[k++] = (double)v.reed.jitter_state;
    o[k++] = (double)v.noise.rng_state; o[k++] = (double)v.noise.remaining;
}
void owo_chain_init(const owg_bench_job* j, double* o) {
    Speaker s(j->v.sample_rate);
    s.set_character(j->speaker_character);
    const double v[18] = {j->volume, s.a2, s.a3, 1.0 + s.a2 + s.a3, s.thermal_coeff, s.thermal_alpha, s.hpf.b0, s.hpf.b1, s.hpf.b2, s.hpf.a1, s.hpf.a2,
                          s.lpf.b0, s.lpf.b1, s.lpf.b2, s.lpf.a1, s.lpf.a2, s.character < 0.001 ? 0.0 : 1.0, j->v.sample_rate < 88200.0 ? 1.0 : 0.0};
    for (int i = 0; i < 18; i++) o[i] = v[i];
}

// Per-stage single-thread timers of the restatement (ns per preamp-rate sample), for the CPU-baseline sanity gate:
// out[0] gen_preamp step with the per-sample matrix rebuild (the tremolo case: set_runtime_R + process_sample), out[1] the same with
// a static R_ldr, out[2] one Twin-T oscillator step (gen_tremolo process_sample), all at 88.2 kHz.
int owo_stage_timers(int n, double* out3) {
    if (!out3 || n <= 0) return OWG_E_BAD_ARG;
    using clk = std::chrono::steady_clock;
    volatile double sink = 0.0;
    pre::CircuitState a;
    a.set_default();
    a.set_sample_rate(88200.0);
    auto t0 = clk::now();
    for (int i = 0; i < n; i++) { a.set_runtime_R_r_ldr(50000.0 + (i % 1000)); sink = sink + pre::process_sample(1e-3 * std::sin(i * 0.03), a); }
    auto t1 = clk::now();
    for (int i = 0; i < n; i++) sink = sink + pre::process_sample(1e-3 * std::sin(i * 0.03), a);
    auto t2 = clk::now();
    trm::CircuitState t;
    t.set_default();
    t.set_sample_rate(88200.0);
    auto t3 = clk::now();
    for (int i = 0; i < n; i++) sink = sink + trm::process_sample(0.0, t);
    auto t4 = clk::now();
    out3[0] = std::chrono::duration<double, std::nano>(t1 - t0).count() / n;
    out3[1] = std::chrono::duration<double, std::nano>(t2 - t1).count() / n;
    out3[2] = std::chrono::duration<double, std::nano>(t4 - t3).count() / n;
    return OWG_OK;
}

// Engine session probe for the restated reference tests (engine.rs:682-1178): runs a script of WurliEngine calls and records the
// voice-state counts at every QUERY.  ops[i] = {kind, a, b}:
//   0 note_on(a, b)  1 note_off(a)  2 set_sustain(a != 0)  3 render(a samples; b != 0: append to `out`)  4 QUERY(note a)
//   5 set_volume(a)  6 set_tremolo_depth(a)  7 set_speaker_character(a)  8 reset()  9 set_sample_rate(a)  10 set_mlp_enabled(a != 0)
// Each QUERY appends 8 ints to `counts`: active, held, sustained, releasing, has_steal_voice_for(a), voices with note a Held,
// voices with note a Sustained, sustain flag.  Returns the number of samples written to `out` (or -1 on overflow).
int64_t owo_engine_script(double sample_rate, int preamp_model, const double* ops, int64_t n_ops, float* out, int64_t out_cap, int32_t* counts, int64_t counts_cap,
                          double* smoothers /* [3] volume / depth / character .current at the end, may be null */) {
    WurliEngine e(sample_rate, preamp_model);
    int64_t n_out = 0, n_q = 0;
    std::vector<float> buf;
    for (int64_t i = 0; i < n_ops; i++) {
        const int kind = (int)ops[3 * i];
        const double a = ops[3 * i + 1], b = ops[3 * i + 2];
        switch (kind) {
            case 0: e.note_on((uint8_t)a, (float)b); break;
            case 1: e.note_off((uint8_t)a); break;
            case 2: e.set_sustain(a != 0.0); break;
            case 3: {
                const size_t len = (size_t)a;
                buf.assign(len, 0.0f);
                e.render(buf.data(), len);
                if (b != 0.0) {
                    if (n_out + (int64_t)len > out_cap) return -1;
                    std::memcpy(out + n_out, buf.data(), len * sizeof(float));
                    n_out += (int64_t)len;
                }
                break;
            }
            case 4: {
                if ((n_q + 1) * 8 > counts_cap) return -1;
                int32_t* c = counts + n_q * 8;
                const uint8_t note = (uint8_t)a;  // the reference's test helpers compare the stored (clamped) note with the raw argument
                c[0] = e.active_voice_count(); c[1] = e.count_state(VState::Held); c[2] = e.count_state(VState::Sustained);
                c[3] = e.count_state(VState::Releasing);
                c[4] = c[5] = c[6] = 0;
                for (auto& sl : e.voices) {
                    if (sl.midi_note == note && sl.steal_voice) c[4] = 1;  // has_steal_voice_for, engine.rs:631-635
                    if (sl.midi_note == note && sl.state == VState::Held) c[5]++;
                    if (sl.midi_note == note && sl.state == VState::Sustained) c[6]++;
                }
                c[7] = e.sustain_held ? 1 : 0;
                n_q++;
                break;
            }
            case 5: e.volume.set_target(a); break;
            case 6: e.tremolo_depth.set_target(a); break;
            case 7: e.speaker_character.set_target(a); break;
            case 8: e.reset(); break;
            case 9: e.set_sample_rate(a); break;
            case 10: e.mlp_enabled = a != 0.0; break;
            default: return -2;
        }
    }
    if (smoothers) { smoothers[0] = e.volume.current; smoothers[1] = e.tremolo_depth.current; smoothers[2] = e.speaker_character.current; }
    return n_out;
}

}  // extern "C"
