// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity checker), never on the product path.
//
// Restatement of the shared mono chain stages (oversampler.rs, power_amp.rs
// mod behavioral, speaker.rs), of the `preamp-bench render` harness ("chain B",
// tools/preamp-bench/src/main.rs:371-496) and of WurliEngine ("chain E",
// engine.rs).
#pragma once
#include "ow_tremolo.hpp"
#include "ow_preamp_legacy.hpp"
#include <memory>

namespace ow {

// ---- oversampler.rs ---------------------------------------------------------------------
struct Oversampler {  // oversampler.rs:17-147
    static constexpr double A[3] = {0.036681502163648, 0.248030921580110, 0.643184620136480};
    static constexpr double B[3] = {0.110377634768680, 0.420399304190880, 0.854640112701920};
    double ua[3] = {0, 0, 0}, ub[3] = {0, 0, 0}, da[3] = {0, 0, 0}, db[3] = {0, 0, 0};
    double down_delay = 0.0;
    static inline double branch(const double* coef, double* st, double x) {
        double y = x;
        for (int i = 0; i < 3; i++) {
            const double yy = coef[i] * y + st[i];
            st[i] = y - coef[i] * yy;
            y = yy;
        }
        return y;
    }
    inline void up1(double x, double& even, double& odd) {
        even = branch(A, ua, x);
        odd = branch(B, ub, x);
    }
    inline double down1(double in_even, double in_odd) {
        const double a = branch(A, da, in_even);
        const double b = branch(B, db, in_odd);
        const double out = (a + down_delay) * 0.5;
        down_delay = b;
        return out;
    }
    void upsample_2x(const double* in, size_t n, double* out) {
        for (size_t i = 0; i < n; i++) up1(in[i], out[2 * i], out[2 * i + 1]);
    }
    void downsample_2x(const double* in, double* out, size_t n_out) {
        for (size_t i = 0; i < n_out; i++) out[i] = down1(in[2 * i], in[2 * i + 1]);
    }
    void reset() { *this = Oversampler(); }
};

// ---- power_amp.rs (mod behavioral) -------------------------------------------------------
struct PowerAmp {  // power_amp.rs:167-240
    static constexpr double OPEN_LOOP_GAIN = 19000.0;
    static constexpr double FEEDBACK_BETA = 220.0 / (220.0 + 15000.0);
    static constexpr double HEADROOM = 22.0, CROSSOVER_VT = 0.013, QUIESCENT_GAIN = 0.1, NR_TOL = 1e-6;
    double closed_loop_gain = OPEN_LOOP_GAIN / (1.0 + OPEN_LOOP_GAIN * FEEDBACK_BETA);
    uint64_t iter_hist[9] = {0};
    inline double process(double input) {
        double y = rclamp(input * closed_loop_gain, -HEADROOM + NR_TOL, HEADROOM - NR_TOL);
        int it = 0;
        for (; it < 8; it++) {
            const double error = input - FEEDBACK_BETA * y;
            const double v = OPEN_LOOP_GAIN * error;
            // forward_path, power_amp.rs:227-240
            const double v_sq = v * v;
            const double vt_sq = CROSSOVER_VT * CROSSOVER_VT;
            const double exp_term = std::exp(-v_sq / vt_sq);
            const double q = QUIESCENT_GAIN;
            const double cross_gain = q + (1.0 - q) * (1.0 - exp_term);
            const double v_cross = v * cross_gain;
            const double dcross_dv = cross_gain + v * (1.0 - q) * (2.0 * v / vt_sq) * exp_term;
            const double tanh_arg = v_cross / HEADROOM;
            const double tanh_val = std::tanh(tanh_arg);
            const double f_val = HEADROOM * tanh_val;
            const double f_deriv = (1.0 - tanh_val * tanh_val) * dcross_dv;
            const double residual = y - f_val;
            const double jacobian = 1.0 + OPEN_LOOP_GAIN * FEEDBACK_BETA * f_deriv;
            const double delta = residual / jacobian;
            y -= delta;
            if (std::fabs(delta) < NR_TOL) { it++; break; }
        }
        iter_hist[std::min(it, 8)]++;
        return y / HEADROOM;
    }
};

// ---- speaker.rs ---------------------------------------------------------------------------
struct Speaker {  // speaker.rs:33-138
    Biquad hpf, lpf;
    double character, sample_rate, a2, a3, thermal_coeff, thermal_alpha, thermal_state;
    explicit Speaker(double sr) {
        hpf.set(Biquad::HP, 30.0, 0.75, sr);
        lpf.set(Biquad::LP, 5500.0, 0.707, sr);
        character = 1.0;
        sample_rate = sr;
        a2 = a3 = thermal_coeff = 0.0;
        thermal_alpha = 1.0 / (5.0 * sr);
        thermal_state = 0.0;
        update_coefficients();
    }
    void set_character(double ch) {
        const double c = rclamp(ch, 0.0, 1.0);
        if (std::fabs(c - character) > 0.002) { character = c; update_coefficients(); }
    }
    void update_coefficients() {
        const double c = character;
        const double hpf_hz = 20.0 * std::pow(30.0 / 20.0, c);
        const double lpf_hz = 20000.0 * std::pow(5500.0 / 20000.0, c);
        hpf.set(Biquad::HP, hpf_hz, 0.75, sample_rate);   // set_type keeps state
        lpf.set(Biquad::LP, lpf_hz, 0.707, sample_rate);
        a2 = 0.2 * c;
        a3 = 0.6 * c;
        thermal_coeff = 2.0 * c;
    }
    inline double process(double input) {
        const double x2 = input * input;
        const double x3 = x2 * input;
        const double shaped = (input + a2 * x2 + a3 * x3) / (1.0 + a2 + a3);
        const double limited = character < 0.001 ? shaped : std::tanh(shaped);
        const double power = x2;
        thermal_state += (power - thermal_state) * thermal_alpha;
        const double thermal_gain = 1.0 / (1.0 + thermal_coeff * std::sqrt(thermal_state));
        const double filtered = hpf.process(limited * thermal_gain);
        return lpf.process(filtered);
    }
    void reset() { hpf.reset(); lpf.reset(); thermal_state = 0.0; }
};

// ---- chain B: `preamp-bench render` (main.rs:371-496) --------------------------------------
struct BenchJob {
    uint8_t midi = 60;
    double velocity = 100.0 / 127.0;  // already normalised (vel/127)
    double sample_rate = 44100.0;
    double duration_s = 2.0;
    uint32_t noise_seed = 0;
    bool mlp_enabled = true;
    double ds_override = NAN;
    bool attack_noise = true;
    bool zero_onset = false;
    double r_ldr = 1.0e6;
    double tremolo_depth = 0.0;
    double volume = 0.60;
    double speaker_character = 1.0;
    bool no_preamp = false, no_poweramp = false;
    int preamp_model = 0;  // 0 = melange 12-node (dk_preamp/melange_adapter.rs), 1 = legacy 8-node (dk_preamp_legacy.rs, the default build)
};

// The two DkPreamp implementations behind the reference's `PreampModel` seam (preamp.rs; dk_preamp/mod.rs:14-20).
struct AnyPreamp {
    std::unique_ptr<pre::DkPreamp> mel;
    std::unique_ptr<leg::DkPreamp> lg;
    AnyPreamp(int model, double sr) {
        if (model == 1) lg.reset(new leg::DkPreamp(sr));
        else mel.reset(new pre::DkPreamp(sr));
    }
    void reset() { if (lg) lg->reset(); else mel->reset(); }
    void set_ldr_resistance(double r) { if (lg) lg->set_ldr_resistance(r); else mel->set_ldr_resistance(r); }
    double process_sample(double x, double* pump = nullptr) { return lg ? lg->process_sample(x, pump) : mel->process_sample(x, nullptr, pump); }
};

struct Taps {  // optional per-stage taps (T3 voice out, T4 preamp out, T5 final)
    std::vector<double>* voice = nullptr;
    std::vector<double>* preamp = nullptr;
    std::vector<double>* r_ldr = nullptr;     // shunt R fed to the preamp per OS sample
    std::vector<double>* shadow = nullptr;    // pump (shadow) output per OS sample
};

struct ChainDiag {
    pre::Diag main, shadow;
    uint64_t pa_iter_hist[9] = {0};
    uint64_t trem_nr_hist[16] = {0};
    uint64_t trem_be = 0;
};

static inline std::vector<double> render_bench(const BenchJob& j, Taps* taps = nullptr, ChainDiag* dg = nullptr) {
    const double sr = j.sample_rate;
    const bool do_oversample = sr < 88200.0;
    const double preamp_sr = do_oversample ? sr * 2.0 : sr;
    const size_t n = (size_t)f64_as_u64(j.duration_s * sr);
    std::vector<double> reed(n, 0.0);
    {
        Voice v;
        v.note_on(j.midi, j.velocity, sr, j.noise_seed, j.mlp_enabled, j.zero_onset);
        if (j.ds_override == j.ds_override) v.pickup.displacement_scale = j.ds_override;
        if (!j.attack_noise) v.noise.disable();
        for (size_t off = 0; off < n; off += 1024) v.render(reed.data() + off, std::min<size_t>(1024, n - off));
    }
    if (taps && taps->voice) *taps->voice = reed;
    std::vector<double> pout(n, 0.0);
    if (j.no_preamp) pout = reed;
    else {
        AnyPreamp preamp(j.preamp_model, preamp_sr);
        Tremolo* trem = nullptr;
        if (j.tremolo_depth > 0.0) trem = new Tremolo(j.tremolo_depth, preamp_sr);
        else { preamp.reset(); preamp.set_ldr_resistance(j.r_ldr); }
        auto step = [&](double x) {
            if (trem) {
                const double r = trem->process();
                if (taps && taps->r_ldr) taps->r_ldr->push_back(r);
                preamp.set_ldr_resistance(r);
            }
            double pump = 0.0;
            const double y = preamp.process_sample(x, &pump);
            if (taps && taps->shadow) taps->shadow->push_back(pump);
            return y;
        };
        if (do_oversample) {
            Oversampler os;
            for (size_t i = 0; i < n; i++) {
                double u0, u1;
                os.up1(reed[i], u0, u1);
                const double p0 = step(u0);
                const double p1 = step(u1);
                pout[i] = os.down1(p0, p1);
            }
        } else {
            for (size_t i = 0; i < n; i++) pout[i] = step(reed[i]);
        }
        if (dg) {
            if (preamp.mel) { dg->main = preamp.mel->diag_main; dg->shadow = preamp.mel->diag_shadow; }
            else { for (int b = 0; b < 8; b++) dg->main.nr_iter_hist[b] = preamp.lg->nr_iter_hist[b]; dg->main.nan_reset = preamp.lg->nan_reset; }
            if (trem) { std::memcpy(dg->trem_nr_hist, trem->osc.nr_iter_hist, sizeof(dg->trem_nr_hist)); dg->trem_be = trem->osc.diag_be_fallback; }
        }
        delete trem;
    }
    if (taps && taps->preamp) *taps->preamp = pout;
    PowerAmp pa;
    Speaker spk(sr);
    spk.set_character(j.speaker_character);
    std::vector<double> fin(n, 0.0);
    for (size_t i = 0; i < n; i++) {
        const double att = pout[i] * j.volume * j.volume;
        const double amp = j.no_poweramp ? att : pa.process(att);
        fin[i] = spk.process(amp) * POST_SPEAKER_GAIN;
    }
    if (dg) std::memcpy(dg->pa_iter_hist, pa.iter_hist, sizeof(dg->pa_iter_hist));
    return fin;
}

// ---- preamp-only harness (C2): process_oversampled pattern, main.rs:961-974 + tremolo as in cmd_render :432-461
static inline void preamp_batch_one(const double* in, size_t n, double fs_base, bool oversample,
                                    double tremolo_depth_or_neg, double r_ldr_static, double* out, int preamp_model = 0, ChainDiag* dg = nullptr) {
    const double preamp_sr = oversample ? fs_base * 2.0 : fs_base;
    AnyPreamp preamp(preamp_model, preamp_sr);
    Tremolo* trem = nullptr;
    if (tremolo_depth_or_neg > 0.0) trem = new Tremolo(tremolo_depth_or_neg, preamp_sr);
    else { preamp.reset(); preamp.set_ldr_resistance(r_ldr_static); }
    auto step = [&](double x) {
        if (trem) preamp.set_ldr_resistance(trem->process());
        return preamp.process_sample(x);
    };
    if (oversample) {
        Oversampler os;
        for (size_t i = 0; i < n; i++) {
            double u0, u1;
            os.up1(in[i], u0, u1);
            const double p0 = step(u0);
            const double p1 = step(u1);
            out[i] = os.down1(p0, p1);
        }
    } else {
        for (size_t i = 0; i < n; i++) out[i] = step(in[i]);
    }
    if (dg && preamp.mel) { dg->main = preamp.mel->diag_main; dg->shadow = preamp.mel->diag_shadow; }
    delete trem;
}

// ---- rows through chain B with either construction order of the static preamp (cmd_render vs render-poly / render-midi) ----------
static inline std::vector<double> chain_rows(const double* in, size_t n, const BenchJob& j, bool set_then_reset) {
    const double sr = j.sample_rate;
    const bool do_oversample = sr < 88200.0;
    const double preamp_sr = do_oversample ? sr * 2.0 : sr;
    std::vector<double> pout(in, in + n);
    if (!j.no_preamp) {
        AnyPreamp preamp(j.preamp_model, preamp_sr);
        Tremolo* trem = nullptr;
        if (j.tremolo_depth > 0.0) trem = new Tremolo(j.tremolo_depth, preamp_sr);
        else if (set_then_reset) { preamp.set_ldr_resistance(j.r_ldr); preamp.reset(); }   // main.rs:1463-1464, 1753-1754
        else { preamp.reset(); preamp.set_ldr_resistance(j.r_ldr); }                       // main.rs:438-439
        auto step = [&](double x) {
            if (trem) preamp.set_ldr_resistance(trem->process());
            return preamp.process_sample(x);
        };
        if (do_oversample) {
            Oversampler os;
            for (size_t i = 0; i < n; i++) {
                double u0, u1;
                os.up1(in[i], u0, u1);
                const double p0 = step(u0);
                const double p1 = step(u1);
                pout[i] = os.down1(p0, p1);
            }
        } else {
            for (size_t i = 0; i < n; i++) pout[i] = step(in[i]);
        }
        delete trem;
    }
    PowerAmp pa;
    Speaker spk(sr);
    spk.set_character(j.speaker_character);
    std::vector<double> fin(n, 0.0);
    for (size_t i = 0; i < n; i++) {
        const double att = pout[i] * j.volume * j.volume;
        fin[i] = spk.process(j.no_poweramp ? att : pa.process(att)) * POST_SPEAKER_GAIN;
    }
    return fin;
}

// ---- `preamp-bench render-midi` (main.rs:1726-1880): the tool's own voice manager, 64-sample chunks, chain B ---------------------
struct MidiEvt { double time_s; int kind; uint8_t note, velocity; };  // kind 0 on, 1 off, 2 pedal (velocity != 0 = down)
static inline std::vector<double> render_midi(const std::vector<MidiEvt>& events, size_t total_samples, double volume, double speaker_char,
                                              bool no_poweramp, int preamp_model, uint64_t* note_ons = nullptr, uint64_t* peak_poly = nullptr) {
    const double BASE_SR = 44100.0;
    struct Slot { std::unique_ptr<Voice> voice; bool active = false; uint8_t midi_note = 0; uint64_t age = 0; };
    std::vector<Slot> voices(64);
    uint64_t age_counter = 0, n_on = 0, peak = 0;
    std::vector<double> sum(total_samples, 0.0);
    std::vector<double> voice_buf(64);
    size_t event_idx = 0, sample_pos = 0;
    bool pedal_down = false;
    std::vector<uint8_t> pedal_held;
    auto note_off_oldest = [&](uint8_t note) {
        Slot* best = nullptr;
        for (auto& sl : voices) if (sl.active && sl.midi_note == note && (!best || sl.age < best->age)) best = &sl;
        if (best && best->voice) best->voice->note_off();
    };
    while (sample_pos < total_samples) {
        const size_t chunk_end = std::min(sample_pos + 64, total_samples);
        const size_t len = chunk_end - sample_pos;
        const double chunk_time = (double)sample_pos / BASE_SR;
        while (event_idx < events.size() && events[event_idx].time_s <= chunk_time) {
            const MidiEvt& e = events[event_idx];
            if (e.kind == 0) {
                const uint8_t note = (uint8_t)rclamp((double)e.note, 33.0, 96.0);
                const double vel = (double)e.velocity / 127.0;
                age_counter += 1; n_on += 1;
                size_t slot = 64;
                for (size_t i = 0; i < 64; i++) if (!voices[i].active) { slot = i; break; }
                if (slot == 64) { slot = 0; for (size_t i = 1; i < 64; i++) if (voices[i].age < voices[slot].age) slot = i; }
                const uint32_t seed = (uint32_t)note * 2654435761u + (uint32_t)age_counter;
                voices[slot].voice.reset(new Voice());
                voices[slot].voice->note_on(note, vel, BASE_SR, seed, true);
                voices[slot].active = true; voices[slot].midi_note = note; voices[slot].age = age_counter;
                uint64_t act = 0;
                for (auto& sl : voices) act += sl.active ? 1 : 0;
                peak = std::max(peak, act);
            } else if (e.kind == 1) {
                const uint8_t note = (uint8_t)rclamp((double)e.note, 33.0, 96.0);
                if (pedal_down) pedal_held.push_back(note); else note_off_oldest(note);
            } else {
                pedal_down = e.velocity != 0;
                if (!pedal_down) { for (uint8_t h : pedal_held) note_off_oldest(h); pedal_held.clear(); }
            }
            event_idx++;
        }
        for (auto& sl : voices) if (sl.active && sl.voice && sl.voice->is_silent()) { sl.active = false; sl.voice.reset(); }
        for (auto& sl : voices) {
            if (!sl.active || !sl.voice) continue;
            std::fill(voice_buf.begin(), voice_buf.begin() + len, 0.0);
            sl.voice->render(voice_buf.data(), len);
            for (size_t i = 0; i < len; i++) sum[sample_pos + i] += voice_buf[i];
        }
        sample_pos = chunk_end;
    }
    if (note_ons) *note_ons = n_on;
    if (peak_poly) *peak_poly = peak;
    BenchJob j;
    j.sample_rate = BASE_SR; j.r_ldr = 1000000.0; j.tremolo_depth = 0.0; j.volume = volume; j.speaker_character = speaker_char;
    j.no_poweramp = no_poweramp; j.preamp_model = preamp_model;
    return chain_rows(sum.data(), total_samples, j, true);
}

// ---- `preamp-bench calibrate` (main.rs:1069-1260): one CalibrateRow, restated literally (own reed, own pickup, T1..T5) ----------
struct CalibrateRow { double v[18]; };
static inline double dft_magnitude(const double* x, size_t n, double freq, double sr) {  // main.rs:893-903
    double re = 0.0, im = 0.0;
    for (size_t i = 0; i < n; i++) {
        const double phase = 2.0 * PI * freq * (double)i / sr;
        re += x[i] * std::cos(phase);
        im -= x[i] * std::sin(phase);
    }
    const double nn = (double)n;
    return 2.0 * std::sqrt((re / nn) * (re / nn) + (im / nn) * (im / nn));
}
static inline double peak_abs(const double* x, size_t n) { double p = 0.0; for (size_t i = 0; i < n; i++) p = rmax(p, std::fabs(x[i])); return p; }
static inline double to_dbfs(double v) { return v > 1e-15 ? 20.0 * std::log10(v) : -120.0; }  // main.rs:2241-2247
static inline double rms_db(const double* x, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += x[i] * x[i];
    const double m = s / (double)n;
    return m > 0.0 ? 10.0 * std::log10(m) : -120.0;
}
static inline double h2_h1_ratio_db(const double* x, size_t n, double f0, double sr) {
    const double h1 = dft_magnitude(x, n, f0, sr), h2 = dft_magnitude(x, n, 2.0 * f0, sr);
    return h1 > 1e-15 ? 20.0 * std::log10(h2 / h1) : -120.0;
}
static inline CalibrateRow calibrate_row(uint8_t note, uint8_t vel_byte, const CalibrationConfig& cfg, double volume, double speaker_char,
                                         int preamp_model) {
    const double BASE_SR = 44100.0, duration = 0.5;
    const size_t measure_start = (size_t)f64_as_u64(0.100 * BASE_SR), measure_end = (size_t)f64_as_u64(0.400 * BASE_SR);
    const NoteParams params = note_params(note);
    const double freq = params.fundamental_hz;
    const double ds_actual = pickup_displacement_scale_with_config(note, cfg);
    const double velocity = (double)vel_byte / 127.0;
    // T1: raw reed (onset_time 0, no MLP, no attack noise)
    const double detuned = params.fundamental_hz * freq_detune(note);
    double dwell[NUM_MODES], amp_offsets[NUM_MODES], amplitudes[NUM_MODES];
    dwell_attenuation(velocity, detuned, params.mode_ratios, dwell);
    mode_amplitude_offsets(note, amp_offsets);
    const double vel_scale = std::pow(velocity_scurve(velocity), velocity_exponent(note));
    for (int i = 0; i < NUM_MODES; i++) amplitudes[i] = params.mode_amplitudes[i] * dwell[i] * amp_offsets[i] * vel_scale;
    ModalReed reed;
    reed.init(detuned, params.mode_ratios, amplitudes, params.mode_decay_rates, 0.0, velocity, BASE_SR, (uint32_t)note * 2654435761u);
    const size_t n = (size_t)f64_as_u64(duration * BASE_SR);
    std::vector<double> reed_buf(n, 0.0);
    reed.render(reed_buf.data(), n);
    const size_t wn = measure_end - measure_start;
    const double reed_peak = peak_abs(reed_buf.data() + measure_start, wn);
    // T2: pickup
    Pickup pickup;
    pickup.init(BASE_SR, ds_actual);
    std::vector<double> t2 = reed_buf;
    pickup.process(t2.data(), n);
    // T3: output scale
    const double out_scale = output_scale_with_config(note, velocity, cfg);
    std::vector<double> t3(n);
    for (size_t i = 0; i < n; i++) t3[i] = t2[i] * out_scale;
    // T4: preamp, 2x oversampled, R_ldr = 1 MOhm on a fresh preamp (no reset)
    AnyPreamp preamp(preamp_model, BASE_SR * 2.0);
    preamp.set_ldr_resistance(1000000.0);
    std::vector<double> t4(n);
    {
        Oversampler os;
        for (size_t i = 0; i < n; i++) {
            double u0, u1;
            os.up1(t3[i], u0, u1);
            const double p0 = preamp.process_sample(u0);
            const double p1 = preamp.process_sample(u1);
            t4[i] = os.down1(p0, p1);
        }
    }
    // T5: volume^2 -> power amp -> speaker
    PowerAmp pa;
    Speaker spk(BASE_SR);
    spk.set_character(speaker_char);
    std::vector<double> t5(n);
    for (size_t i = 0; i < n; i++) t5[i] = spk.process(pa.process(t4[i] * volume * volume)) * POST_SPEAKER_GAIN;
    CalibrateRow r;
    const double* w2 = t2.data() + measure_start; const double* w3 = t3.data() + measure_start;
    const double* w4 = t4.data() + measure_start; const double* w5 = t5.data() + measure_start;
    r.v[0] = cfg.ds_at_c4; r.v[1] = ds_actual; r.v[2] = reed_peak * ds_actual;
    r.v[3] = to_dbfs(peak_abs(w2, wn)); r.v[4] = rms_db(w2, wn); r.v[5] = h2_h1_ratio_db(w2, wn, freq, BASE_SR);
    r.v[6] = to_dbfs(peak_abs(w3, wn)); r.v[7] = rms_db(w3, wn);
    r.v[8] = to_dbfs(peak_abs(w4, wn)); r.v[9] = rms_db(w4, wn); r.v[10] = h2_h1_ratio_db(w4, wn, freq, BASE_SR);
    r.v[11] = to_dbfs(peak_abs(w5, wn)); r.v[12] = rms_db(w5, wn); r.v[13] = h2_h1_ratio_db(w5, wn, freq, BASE_SR);
    r.v[14] = 20.0 * std::log10(out_scale);
    r.v[15] = cfg.zero_trim ? 0.0 : register_trim_db(note);
    r.v[16] = r.v[7] - cfg.target_db;
    r.v[17] = r.v[8] - r.v[11];
    return r;
}

}  // namespace ow
