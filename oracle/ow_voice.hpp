// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity checker), never shipped, never on the product path.
//
// Line-by-line f64 restatement of the reference's per-voice synthesis
// (hal0zer0/openwurli v0.6.0, crates/openwurli-dsp/src/). Same operation
// order, no FMA contraction (compile with -ffp-contract=off), glibc libm for
// every transcendental -- the libm Rust's std calls on Linux.
//
// Parity pinning: the reference cannot be built in this environment (no Rust
// toolchain), so this restatement is pinned only by the reference's own
// known-answer unit tests and fixtures (see tests/test_oracle_known_answers.py);
// there is no sample-level golden vector upstream.  The Biquad arithmetic
// (melange-primitives @ de9dc81, un-vendored) is restated from the RBJ Audio
// EQ Cookbook + DF-II-transposed description in filters.rs:1-11 and is
// UNPINNED at bit level.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

namespace ow {

#define OWC_TABLE(name) static const double name
#define OWC_SCALAR(name) static const double name
#include "ow_consts.inc"
#undef OWC_TABLE
#undef OWC_SCALAR

constexpr int NUM_MODES = 7;                       // tables.rs:6
constexpr double PI = 3.14159265358979323846;      // std::f64::consts::PI
constexpr double TAU = 6.28318530717958647692;     // std::f64::consts::TAU

// ---- Rust numeric semantics -------------------------------------------------
static inline double rclamp(double x, double lo, double hi) {  // f64::clamp
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
static inline double rmax(double a, double b) { return std::fmax(a, b); }  // f64::max
static inline double rmin(double a, double b) { return std::fmin(a, b); }  // f64::min
static inline uint32_t f64_as_u32(double x) {  // `as u32`: saturating, NaN -> 0
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}
static inline uint64_t f64_as_u64(double x) {  // `as u64` / `as usize`
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)x;
}
static inline uint8_t f64_as_u8(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 255.0) return 255;
    return (uint8_t)x;
}

// ---- tables.rs ----------------------------------------------------------------
static const double BASE_MODE_AMPLITUDES[NUM_MODES] = {1.0, 0.005, 0.0035, 0.0018, 0.0011, 0.0007, 0.0005};  // tables.rs:30-31

static inline double midi_to_freq(uint8_t midi) {  // tables.rs:34-36
    return 440.0 * std::pow(2.0, ((double)midi - 69.0) / 12.0);
}

static inline double tip_mass_ratio(uint8_t midi) {  // tables.rs:52-80
    const double m = (double)midi;
    static const double ax[5] = {33.0, 52.0, 62.0, 74.0, 96.0};
    static const double ay[5] = {0.10, 0.00, 0.00, 0.02, 0.01};
    if (m <= ax[0]) return ay[0];
    if (m >= ax[4]) return ay[4];
    for (int i = 0; i < 4; i++) {
        if (m <= ax[i + 1]) {
            double t = (m - ax[i]) / (ax[i + 1] - ax[i]);
            return ay[i] + t * (ay[i + 1] - ay[i]);
        }
    }
    return 0.0;
}

static inline void eigenvalues(double mu, double out[NUM_MODES]) {  // tables.rs:88-150
    static const double tmu[8] = {0.00, 0.01, 0.05, 0.10, 0.15, 0.20, 0.30, 0.50};
    static const double tb[8][NUM_MODES] = {
        {1.8751, 4.6941, 7.8548, 10.9955, 14.1372, 17.2788, 20.4204},
        {1.8584, 4.6849, 7.8504, 10.9930, 14.1356, 17.2776, 20.4195},
        {1.7920, 4.6477, 7.8316, 10.9830, 14.1288, 17.2726, 20.4158},
        {1.7227, 4.6024, 7.8077, 10.9700, 14.1198, 17.2660, 20.4110},
        {1.6625, 4.5618, 7.7859, 10.9580, 14.1114, 17.2598, 20.4065},
        {1.6097, 4.5254, 7.7659, 10.9470, 14.1036, 17.2540, 20.4023},
        {1.5201, 4.4620, 7.7310, 10.9280, 14.0894, 17.2434, 20.3946},
        {1.3853, 4.3601, 7.6745, 10.8970, 14.0650, 17.2252, 20.3814},
    };
    const double muc = rclamp(mu, 0.0, 0.50);
    int lo = 0;
    for (int i = 7; i >= 0; i--) {  // rposition(row.mu <= mu_clamped)
        if (tmu[i] <= muc) { lo = i; break; }
    }
    const int hi = std::min(lo + 1, 7);
    const double t = (tmu[hi] > tmu[lo]) ? (muc - tmu[lo]) / (tmu[hi] - tmu[lo]) : 0.0;
    for (int i = 0; i < NUM_MODES; i++) out[i] = tb[lo][i] + t * (tb[hi][i] - tb[lo][i]);
}

static inline void mode_ratios(double mu, double out[NUM_MODES]) {  // tables.rs:153-157
    double b[NUM_MODES];
    eigenvalues(mu, b);
    const double b1sq = b[0] * b[0];
    for (int i = 0; i < NUM_MODES; i++) out[i] = (b[i] * b[i]) / b1sq;
}

static inline double reed_length_mm(uint8_t midi) {  // tables.rs:167-175
    const double n = rclamp((double)midi - 32.0, 1.0, 64.0);
    const double inches = (n <= 20.0) ? 3.0 - n / 20.0 : 2.0 - (n - 20.0) / 44.0;
    return inches * 25.4;
}

static inline void reed_blank_dims(uint8_t midi, double* w_mm, double* t_mm) {  // tables.rs:186-212
    int reed = std::min(std::max((int)midi - 32, 1), 64);
    double width_inch;
    if (reed <= 14) width_inch = 0.151;
    else if (reed <= 20) width_inch = 0.127;
    else if (reed <= 42) width_inch = 0.121;
    else if (reed <= 50) width_inch = 0.111;
    else width_inch = 0.098;
    double thickness_inch;
    if (reed <= 16) thickness_inch = 0.026;
    else if (reed <= 26) {
        double t = ((double)reed - 16.0) / 10.0;
        thickness_inch = 0.026 + t * (0.034 - 0.026);
    } else thickness_inch = 0.034;
    *w_mm = width_inch * 25.4;
    *t_mm = thickness_inch * 25.4;
}

static inline double reed_compliance(uint8_t midi) {  // tables.rs:218-222
    const double l = reed_length_mm(midi);
    double w, t;
    reed_blank_dims(midi, &w, &t);
    return (l * l * l) / (w * t * t * t);
}

struct CalibrationConfig {  // tables.rs:256-277
    double ds_at_c4 = 0.85;
    double ds_exponent = 0.75;
    double ds_clamp_lo = 0.02, ds_clamp_hi = 0.95;
    double target_db = -35.0;
    double voicing_slope = -0.04;
    bool zero_trim = false;
};

static inline double pickup_displacement_scale_with_config(uint8_t midi, const CalibrationConfig& cfg) {  // tables.rs:283-288
    const double c = reed_compliance(midi);
    const double c_ref = reed_compliance(60);
    const double ds = cfg.ds_at_c4 * std::pow(c / c_ref, cfg.ds_exponent);
    return rclamp(ds, cfg.ds_clamp_lo, cfg.ds_clamp_hi);
}
static inline double pickup_displacement_scale(uint8_t midi) {
    return pickup_displacement_scale_with_config(midi, CalibrationConfig());
}

static inline double mode_shape(double beta, double xi) {  // tables.rs:296-300
    const double sigma = (std::cosh(beta) + std::cos(beta)) / (std::sinh(beta) + std::sin(beta));
    const double bx = beta * xi;
    return std::cosh(bx) - std::cos(bx) - sigma * (std::sinh(bx) - std::sin(bx));
}

static inline void spatial_coupling_coefficients(double mu, double reed_len_mm, double out[NUM_MODES]) {  // tables.rs:324-370
    double betas[NUM_MODES];
    eigenvalues(mu, betas);
    const double ell_over_l = rclamp(6.0 / reed_len_mm, 0.0, 1.0);
    double kappa_raw[NUM_MODES] = {0};
    const int N_SIMPSON = 32;
    const double xi_start = 1.0 - ell_over_l;
    for (int mode = 0; mode < NUM_MODES; mode++) {
        const double beta = betas[mode];
        const double tip_val = mode_shape(beta, 1.0);
        if (std::fabs(tip_val) < 1e-30 || ell_over_l < 1e-12) {
            kappa_raw[mode] = 1.0;
            continue;
        }
        const double h = ell_over_l / (double)N_SIMPSON;
        double sum = mode_shape(beta, xi_start) + mode_shape(beta, 1.0);
        for (int j = 1; j < N_SIMPSON; j++) {
            const double xi = xi_start + (double)j * h;
            const double coeff = (j % 2 == 1) ? 4.0 : 2.0;
            sum += coeff * mode_shape(beta, xi);
        }
        const double integral = sum * h / 3.0;
        const double k = std::fabs(integral / (ell_over_l * tip_val));
        kappa_raw[mode] = rclamp(k, 0.0, 1.0);
    }
    const double k1 = kappa_raw[0];
    if (k1 > 1e-30) {
        for (int i = 0; i < NUM_MODES; i++) out[i] = rclamp(kappa_raw[i] / k1, 0.0, 1.0);
    } else {
        for (int i = 0; i < NUM_MODES; i++) out[i] = 1.0;
    }
}

static inline double fundamental_decay_rate(uint8_t midi) {  // tables.rs:384-387
    const double f = midi_to_freq(midi);
    return rmax(0.005 * std::pow(f, 1.22), 3.0);
}

static inline void mode_decay_rates(uint8_t midi, const double ratios[NUM_MODES], double out[NUM_MODES]) {  // tables.rs:407-410
    const double base = fundamental_decay_rate(midi);
    for (int i = 0; i < NUM_MODES; i++) out[i] = base * ratios[i] * ratios[i];
}

static inline double pickup_rms_proxy(double ds, double f0, double fc) {  // tables.rs:438-454
    if (ds < 1e-10) return 0.0;
    const double r = (1.0 - std::sqrt(1.0 - ds * ds)) / ds;
    const double inv_sqrt = 1.0 / std::sqrt(1.0 - ds * ds);
    double sum_sq = 0.0;
    double r_n = r;
    for (uint32_t n = 1; n <= 8; n++) {
        const double cn = 2.0 * r_n * inv_sqrt;
        const double nf = (double)n * f0;
        const double hpf_n = nf / std::sqrt(nf * nf + fc * fc);
        sum_sq += (cn * hpf_n) * (cn * hpf_n);
        r_n *= r;
    }
    return std::sqrt(sum_sq);
}

static inline double register_trim_db(uint8_t midi) {  // tables.rs:465-503
    static const double ax[13] = {36, 40, 44, 48, 52, 56, 60, 64, 68, 72, 76, 80, 84};
    static const double ay[13] = {-1.3, 0.0, -1.3, 0.7, 0.2, -1.0, 0.0, 0.9, 1.2, 0.0, 1.8, 2.4, 3.6};
    const double m = (double)midi;
    if (m <= ax[0]) return ay[0];
    if (m >= ax[12]) return ay[12];
    for (int i = 0; i < 12; i++) {
        if (m <= ax[i + 1]) {
            const double t = (m - ax[i]) / (ax[i + 1] - ax[i]);
            return ay[i] + t * (ay[i + 1] - ay[i]);
        }
    }
    return 0.0;
}

constexpr double POST_SPEAKER_GAIN = 7.498942093324558;  // tables.rs:536
constexpr double FIXED_CIRCUIT_DRIVE = 0.25;             // tables.rs:557

static inline double velocity_exponent(uint8_t midi) {  // tables.rs:632-651
    const double m = (double)midi;
    const double center = 62.0, sigma = 15.0, max_exp = 1.7, treble_min = 1.3, bass_min = 0.55;
    const double z = (m - center) / sigma;
    const double t = std::exp(-0.5 * (z * z));
    const double min_exp = (m < center) ? bass_min : treble_min;
    return min_exp + t * (max_exp - min_exp);
}

static inline double velocity_scurve(double velocity) {  // tables.rs:659-665
    const double k = 1.5;
    const double s = 1.0 / (1.0 + std::exp(-k * (velocity - 0.5)));
    const double s0 = 1.0 / (1.0 + std::exp(k * 0.5));
    const double s1 = 1.0 / (1.0 + std::exp(-k * 0.5));
    return (s - s0) / (s1 - s0);
}

static inline double output_scale_with_config(uint8_t midi, double velocity_norm, const CalibrationConfig& cfg) {  // tables.rs:578-616
    const double HPF_FC = 2312.0;
    const double ds = pickup_displacement_scale_with_config(midi, cfg);
    const double f0 = midi_to_freq(midi);
    const double scurve_v = velocity_scurve(velocity_norm);
    const double vel_scale = std::pow(scurve_v, velocity_exponent(midi));
    const double vel_scale_c4 = std::pow(scurve_v, velocity_exponent(60));
    const double effective_ds = rmax(ds * vel_scale, 1e-6);
    const double effective_ds_ref = rmax(cfg.ds_at_c4 * vel_scale_c4, 1e-6);
    const double rms = pickup_rms_proxy(effective_ds, f0, HPF_FC);
    const double rms_ref = pickup_rms_proxy(effective_ds_ref, midi_to_freq(60), HPF_FC);
    const double flat_db = -20.0 * std::log10(rms / rms_ref);
    const double voicing_db = cfg.voicing_slope * rmax((double)midi - 60.0, 0.0);
    const double trim = cfg.zero_trim ? 0.0 : register_trim_db(midi);
    const double vel_blend = std::pow(velocity_norm, 1.3);
    const double effective_trim = trim * vel_blend;
    return std::pow(10.0, (cfg.target_db + flat_db + voicing_db + effective_trim) / 20.0);
}
static inline double output_scale(uint8_t midi, double v) { return output_scale_with_config(midi, v, CalibrationConfig()); }

struct NoteParams {  // tables.rs:667-672, 804-830
    double fundamental_hz;
    double mode_ratios[NUM_MODES];
    double mode_amplitudes[NUM_MODES];
    double mode_decay_rates[NUM_MODES];
};
static inline NoteParams note_params(uint8_t midi) {
    NoteParams p;
    p.fundamental_hz = midi_to_freq(midi);
    const double mu = tip_mass_ratio(midi);
    mode_ratios(mu, p.mode_ratios);
    mode_decay_rates(midi, p.mode_ratios, p.mode_decay_rates);
    double coupling[NUM_MODES];
    spatial_coupling_coefficients(mu, reed_length_mm(midi), coupling);
    for (int i = 0; i < NUM_MODES; i++) p.mode_amplitudes[i] = BASE_MODE_AMPLITUDES[i] * coupling[i];
    return p;
}

// ---- variation.rs -------------------------------------------------------------
static inline double hash_f64(uint8_t midi, uint32_t seed) {  // variation.rs:10-19
    uint32_t h = 2166136261u;
    h ^= (uint32_t)midi;
    h *= 16777619u;
    h ^= seed;
    h *= 16777619u;
    h ^= h >> 16;
    h *= 2654435769u;
    return (double)(h & 0x00FFFFFFu) / 16777216.0;
}
static inline double freq_detune(uint8_t midi) {  // variation.rs:26-29
    const double r = hash_f64(midi, 0xDEAD) * 2.0 - 1.0;
    return 1.0 + r * 0.00173;
}
static inline void mode_amplitude_offsets(uint8_t midi, double out[NUM_MODES]) {  // variation.rs:33-38
    for (int i = 0; i < NUM_MODES; i++) {
        const double r = hash_f64(midi, 0xBEEFu + (uint32_t)i) * 2.0 - 1.0;
        out[i] = 1.0 + r * 0.08;
    }
}

// ---- hammer.rs (closed forms) ---------------------------------------------------
static inline double dwell_time(double velocity, double f0) {  // hammer.rs:26-29
    const double cycles = 0.75 + 0.25 * (1.0 - velocity);
    return rclamp(cycles / f0, 0.0003, 0.020);
}
static inline double onset_ramp_time(double velocity, double f0) {  // hammer.rs:53-57
    const double period_s = 1.0 / f0;
    const double periods = 1.0 + 1.0 * (1.0 - velocity);
    return rmax(periods * period_s, 0.002);
}
static inline void dwell_attenuation(double velocity, double f0, const double ratios[NUM_MODES], double atten[NUM_MODES]) {  // hammer.rs:69-90
    const double t_dwell = dwell_time(velocity, f0);
    const double sigma_sq = 8.0 * 8.0;
    for (int i = 0; i < NUM_MODES; i++) {
        const double ft = f0 * ratios[i] * t_dwell;
        atten[i] = std::exp(-ft * ft / (2.0 * sigma_sq));
    }
    const double a0 = atten[0];
    if (a0 > 1e-30) {
        for (int i = 0; i < NUM_MODES; i++) atten[i] /= a0;
    }
}

// ---- mlp_correction.rs ----------------------------------------------------------
struct MlpCorrections {
    double freq_offsets_cents[5];
    double decay_offsets[5];
    double ds_correction;
};
static inline MlpCorrections mlp_identity() {  // mlp_correction.rs:49-55
    MlpCorrections c;
    for (int i = 0; i < 5; i++) { c.freq_offsets_cents[i] = 0.0; c.decay_offsets[i] = 1.0; }
    c.ds_correction = 1.0;
    return c;
}
static inline MlpCorrections mlp_infer(uint8_t midi_note, double velocity) {  // mlp_correction.rs:61-140
    const double midi = (double)midi_note;
    double fade;
    if (midi < 65.0) fade = rclamp((midi - (65.0 - 12.0)) / 12.0, 0.0, 1.0);
    else if (midi > 97.0) fade = rclamp(((97.0 + 12.0) - midi) / 12.0, 0.0, 1.0);
    else fade = 1.0;
    if (fade <= 0.0) return mlp_identity();
    const double midi_norm = rclamp((midi - 21.0) / (108.0 - 21.0), 0.0, 1.0);
    const double vel_norm = rclamp(velocity, 0.0, 1.0);
    const double input[2] = {midi_norm, vel_norm};
    double h1[16], h2[16], raw[11];
    for (int i = 0; i < 16; i++) {
        double sum = MLP_B1[i];
        for (int j = 0; j < 2; j++) sum += MLP_W1[i][j] * input[j];
        h1[i] = sum > 0.0 ? sum : 0.0;
    }
    for (int i = 0; i < 16; i++) {
        double sum = MLP_B2[i];
        for (int j = 0; j < 16; j++) sum += MLP_W2[i][j] * h1[j];
        h2[i] = sum > 0.0 ? sum : 0.0;
    }
    for (int i = 0; i < 11; i++) {
        double sum = MLP_B3[i];
        for (int j = 0; j < 16; j++) sum += MLP_W3[i][j] * h2[j];
        raw[i] = sum * MLP_TARGET_STDS[i] + MLP_TARGET_MEANS[i];
    }
    MlpCorrections c;
    for (int h = 0; h < 5; h++) c.freq_offsets_cents[h] = rclamp(raw[h] * fade, -100.0, 100.0);
    for (int h = 0; h < 5; h++) {
        const double raw_decay = rclamp(raw[5 + h], 0.3, 3.0);
        c.decay_offsets[h] = 1.0 + (raw_decay - 1.0) * fade;
    }
    const double raw_ds = rclamp(raw[10], 0.7, 1.2);
    c.ds_correction = 1.0 + (raw_ds - 1.0) * fade;
    return c;
}

// ---- filters.rs : Biquad (melange-primitives restated; UNPINNED at bit level) -----
struct Biquad {
    double b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0;
    double s1 = 0, s2 = 0;
    enum Kind { LP, HP, BP };
    void set(Kind kind, double fc, double q, double fs) {  // RBJ Audio EQ Cookbook
        const double w0 = 2.0 * PI * fc / fs;
        const double cw = std::cos(w0), sw = std::sin(w0);
        const double alpha = sw / (2.0 * q);
        double nb0, nb1, nb2;
        if (kind == LP) { nb0 = (1.0 - cw) / 2.0; nb1 = 1.0 - cw; nb2 = (1.0 - cw) / 2.0; }
        else if (kind == HP) { nb0 = (1.0 + cw) / 2.0; nb1 = -(1.0 + cw); nb2 = (1.0 + cw) / 2.0; }
        else { nb0 = sw / 2.0; nb1 = 0.0; nb2 = -sw / 2.0; }  // constant skirt gain
        const double a0 = 1.0 + alpha;
        b0 = nb0 / a0; b1 = nb1 / a0; b2 = nb2 / a0;
        a1 = (-2.0 * cw) / a0; a2 = (1.0 - alpha) / a0;
    }
    inline double process(double x) {  // DF-II transposed
        const double y = b0 * x + s1;
        s1 = b1 * x - a1 * y + s2;
        s2 = b2 * x - a2 * y;
        return y;
    }
    void reset() { s1 = 0; s2 = 0; }
};

// ---- reed.rs ----------------------------------------------------------------------
struct Mode {  // reed.rs:44-67
    double s, c, cos_inc, sin_inc, phase_inc, amplitude, decay_mult, envelope, jitter_drift, damper_rate, damper_mult;
};

static inline double lcg_uniform_scaled(uint32_t& state) {  // reed.rs:90-94
    state = state * 1664525u + 1013904223u;
    const double u = (double)(state >> 1) / (4294967295.0 / 2.0);
    return (u * 2.0 - 1.0) * 1.7320508080;
}

struct ModalReed {
    Mode modes[NUM_MODES];
    uint64_t sample;
    uint64_t onset_ramp_samples;
    double onset_ramp_inc, onset_shape_exp;
    bool damper_active;
    double damper_ramp_samples, damper_release_count;
    bool damper_ramp_done;
    uint32_t jitter_state;
    double jitter_revert, jitter_diffusion;

    void init(double fundamental_hz, const double ratios[NUM_MODES], const double amplitudes[NUM_MODES],
              const double decay_rates_db[NUM_MODES], double onset_time_s, double velocity, double sample_rate,
              uint32_t jitter_seed) {  // reed.rs:108-182
        const double dt = 1.0 / sample_rate;
        jitter_revert = std::exp(-dt / 0.020);
        jitter_diffusion = 0.0004 * std::sqrt(1.0 - jitter_revert * jitter_revert);
        uint32_t js = std::max(jitter_seed, 1u);
        double initial_drifts[NUM_MODES];
        for (int i = 0; i < NUM_MODES; i++) {
            js = js * 1664525u + 1013904223u;
            const double u1 = (double)(js >> 1) / (4294967295.0 / 2.0);
            js = js * 1664525u + 1013904223u;
            const double u2 = (double)(js >> 1) / (4294967295.0 / 2.0);
            const double r = std::sqrt(-2.0 * std::log(rmax(u1, 1e-30)));
            initial_drifts[i] = 0.0004 * r * std::cos(TAU * u2);
        }
        for (int i = 0; i < NUM_MODES; i++) {
            const double freq = fundamental_hz * ratios[i];
            const double phase_inc = TAU * freq / sample_rate;
            const double alpha_nepers = decay_rates_db[i] / 8.686;
            const double decay_per_sample = alpha_nepers / sample_rate;
            Mode& m = modes[i];
            m.s = 0.0; m.c = 1.0;
            m.cos_inc = std::cos(phase_inc);
            m.sin_inc = std::sin(phase_inc);
            m.phase_inc = phase_inc;
            m.amplitude = amplitudes[i];
            m.decay_mult = std::exp(-decay_per_sample);
            m.envelope = 1.0;
            m.jitter_drift = initial_drifts[i];
            m.damper_rate = 0.0;
            m.damper_mult = 1.0;
        }
        const uint64_t ramp_samps = f64_as_u64(std::round(onset_time_s * sample_rate));
        onset_ramp_samples = ramp_samps;
        onset_ramp_inc = ramp_samps > 0 ? PI / (double)ramp_samps : 0.0;
        onset_shape_exp = 1.0 + (1.0 - velocity);
        sample = 0;
        damper_active = false;
        damper_ramp_samples = 0.0;
        damper_release_count = 0.0;
        damper_ramp_done = false;
        jitter_state = js;
    }

    void start_damper(uint8_t midi_note, double sample_rate) {  // reed.rs:191-216
        if (midi_note >= 92) return;
        const double base_rate = rmax(55.0 * std::pow(2.0, ((double)midi_note - 60.0) / 24.0), 0.5);
        double p3 = 1.0;  // 3.0.powi(m): exact small integers
        for (int m = 0; m < NUM_MODES; m++) {
            const double factor = rmin(base_rate * p3, 2000.0);
            modes[m].damper_rate = factor / sample_rate;
            modes[m].damper_mult = std::exp(-modes[m].damper_rate);
            p3 *= 3.0;
        }
        const double ramp_time = midi_note < 48 ? 0.050 : (midi_note < 72 ? 0.025 : 0.008);
        damper_ramp_samples = ramp_time * sample_rate;
        damper_active = true;
        damper_release_count = 0.0;
        damper_ramp_done = false;
    }

    void render(double* output, size_t n_out) {  // reed.rs:219-306 (additive)
        const double revert = jitter_revert, diffusion = jitter_diffusion;
        for (size_t k = 0; k < n_out; k++) {
            double sum = 0.0;
            if (damper_active) {
                damper_release_count += 1.0;
                const double t = damper_release_count;
                const double ramp = damper_ramp_samples;
                if (!damper_ramp_done) {
                    if (t > ramp) damper_ramp_done = true;
                    else {
                        for (auto& m : modes) {
                            const double inst_rate = m.damper_rate * t / ramp;
                            m.envelope *= std::exp(-inst_rate);
                        }
                    }
                }
                if (damper_ramp_done) {
                    for (auto& m : modes) m.envelope *= m.damper_mult;
                }
            }
            double onset;
            if (sample < onset_ramp_samples) {
                const double n = (double)sample;
                const double cosine = 0.5 * (1.0 - std::cos(n * onset_ramp_inc));
                if (onset_shape_exp <= 1.001) onset = cosine;
                else if (onset_shape_exp >= 1.999) onset = cosine * cosine;
                else onset = std::pow(cosine, onset_shape_exp);
            } else onset = 1.0;
            if ((sample & 15) == 0) {
                for (auto& m : modes) {
                    const double noise = lcg_uniform_scaled(jitter_state);
                    m.jitter_drift = revert * m.jitter_drift + diffusion * noise;
                }
            }
            for (auto& m : modes) {
                sum += m.amplitude * m.s * onset * m.envelope;
                const double delta_phase = m.jitter_drift * m.phase_inc;
                const double ci = m.cos_inc - delta_phase * m.sin_inc;
                const double si = m.sin_inc + delta_phase * m.cos_inc;
                const double s_new = m.s * ci + m.c * si;
                const double c_new = m.c * ci - m.s * si;
                m.s = s_new;
                m.c = c_new;
                m.envelope *= m.decay_mult;
            }
            if ((sample & 1023) == 0 && sample > 0) {
                for (auto& m : modes) {
                    const double r_sq = m.s * m.s + m.c * m.c;
                    const double r_inv = 1.0 / std::sqrt(r_sq);
                    m.s *= r_inv;
                    m.c *= r_inv;
                }
            }
            output[k] += sum;
            sample += 1;
        }
    }

    bool is_silent(double threshold_db) const {  // reed.rs:309-314
        const double thr = std::pow(10.0, threshold_db / 20.0);
        for (auto& m : modes)
            if (!(std::fabs(m.amplitude * m.envelope) <= thr)) return false;
        return true;
    }
    double release_seconds(double sample_rate) const { return damper_active ? damper_release_count / sample_rate : 0.0; }
};

// ---- hammer.rs : AttackNoise ---------------------------------------------------------
struct AttackNoise {  // hammer.rs:108-198
    double amplitude, decay_per_sample;
    uint32_t remaining, fade_in_remaining;
    Biquad bpf;
    uint32_t rng_state;
    void init(double velocity, double f0, double sample_rate, uint32_t seed) {
        amplitude = 0.025 * velocity * velocity;
        const double tau = 0.003;
        decay_per_sample = std::exp(-1.0 / (tau * sample_rate));
        remaining = f64_as_u32(0.015 * sample_rate);
        fade_in_remaining = 16;
        const double center = rclamp(f0 * 5.0, 200.0, 2000.0);
        bpf = Biquad();
        bpf.set(Biquad::BP, center, 0.7, sample_rate);
        rng_state = seed;
    }
    size_t render(double* output, size_t n_out) {
        const size_t count = std::min((size_t)remaining, n_out);
        double amp = amplitude;
        uint32_t fade_in = fade_in_remaining;
        for (size_t k = 0; k < count; k++) {
            double env;
            if (fade_in > 0) {
                const uint32_t pos = 16 - fade_in;
                const double t = (double)pos / 16.0;
                fade_in -= 1;
                env = 0.5 * (1.0 - std::cos(PI * t));
            } else env = 1.0;
            rng_state = rng_state * 1664525u + 1013904223u;
            const double noise = (double)(int32_t)rng_state / 2147483647.0;
            const double filtered = bpf.process(noise);
            output[k] += amp * env * filtered;
            amp *= decay_per_sample;
        }
        amplitude = amp;
        fade_in_remaining = fade_in;
        remaining -= (uint32_t)count;
        return count;
    }
    bool is_done() const { return remaining == 0; }
    void disable() { remaining = 0; }
};

// ---- pickup.rs -------------------------------------------------------------------------
struct Pickup {  // pickup.rs:33-149
    double q, beta, displacement_scale;
    void init(double sample_rate, double scale) {
        const double TAU_RC = 287.0e3 * 240.0e-12;
        const double dt = 1.0 / sample_rate;
        beta = dt / (2.0 * TAU_RC);
        q = 1.0;
        displacement_scale = scale;
    }
    static inline double soft_saturate(double y) {  // pickup.rs:72-80
        const double abs_y = std::fabs(y);
        if (abs_y < 0.94) return y;
        const double range = 0.98 - 0.94;
        const double saturated = 0.94 + range * std::tanh((abs_y - 0.94) / range);
        return std::copysign(saturated, y);
    }
    void process(double* buf, size_t n) {
        const double scale = displacement_scale, b = beta;
        for (size_t i = 0; i < n; i++) {
            const double y = soft_saturate(buf[i] * scale);
            const double one_minus_y = 1.0 - y;
            const double alpha = b * one_minus_y;
            const double q_next = (q * (1.0 - alpha) + 2.0 * b) / (1.0 + alpha);
            q = q_next;
            buf[i] = (q_next * one_minus_y - 1.0) * 1.8375;
        }
    }
};

// ---- voice.rs -----------------------------------------------------------------------------
struct Voice {
    ModalReed reed;
    Pickup pickup;
    AttackNoise noise;
    double post_pickup_gain;
    double sample_rate;
    uint8_t midi_note;

    // zero_onset: onset_time_s = 0.0, as run_calibrate builds its reed directly (preamp-bench main.rs:1165-1174)
    void note_on(uint8_t midi, double velocity, double sr, uint32_t noise_seed, bool mlp_enabled, bool zero_onset = false) {  // voice.rs:28-142
        const NoteParams params = note_params(midi);
        const double detuned = params.fundamental_hz * freq_detune(midi);
        double dwell[NUM_MODES];
        dwell_attenuation(velocity, detuned, params.mode_ratios, dwell);
        const double onset_time = zero_onset ? 0.0 : onset_ramp_time(velocity, detuned);
        double amp_offsets[NUM_MODES];
        mode_amplitude_offsets(midi, amp_offsets);
        double amplitudes[NUM_MODES];
        for (int i = 0; i < NUM_MODES; i++) amplitudes[i] = params.mode_amplitudes[i] * dwell[i] * amp_offsets[i];
        const double vel_exp = velocity_exponent(midi);
        const double vel_scale = std::pow(velocity_scurve(velocity), vel_exp);
        for (int i = 0; i < NUM_MODES; i++) amplitudes[i] *= vel_scale;
        const MlpCorrections corr = mlp_enabled ? mlp_infer(midi, velocity) : mlp_identity();
        double corrected_ratios[NUM_MODES], corrected_decay[NUM_MODES];
        for (int i = 0; i < NUM_MODES; i++) { corrected_ratios[i] = params.mode_ratios[i]; corrected_decay[i] = params.mode_decay_rates[i]; }
        for (int i = 1; i < 6; i++) corrected_ratios[i] *= std::pow(2.0, corr.freq_offsets_cents[i - 1] / 1200.0);
        for (int i = 1; i < 6; i++) corrected_decay[i] /= corr.decay_offsets[i - 1];
        const double corrected_ds = pickup_displacement_scale(midi) * corr.ds_correction;
        reed.init(detuned, corrected_ratios, amplitudes, corrected_decay, onset_time, velocity, sr, noise_seed);
        pickup.init(sr, 0.85);
        pickup.displacement_scale = corrected_ds;
        noise.init(velocity, detuned, sr, noise_seed);
        const double base_output_scale = output_scale(midi, velocity);
        const double base_ds = pickup_displacement_scale(midi);
        double comp = 1.0;
        if (std::fabs(corr.ds_correction - 1.0) > 1e-6) {
            const double f0 = midi_to_freq(midi);
            const double proxy_base = pickup_rms_proxy(base_ds, f0, 2312.0);
            const double proxy_corr = pickup_rms_proxy(corrected_ds, f0, 2312.0);
            comp = (proxy_corr > 1e-10) ? std::sqrt(proxy_base / proxy_corr) : 1.0;
        }
        post_pickup_gain = base_output_scale * comp;
        sample_rate = sr;
        midi_note = midi;
    }
    void note_off() { reed.start_damper(midi_note, sample_rate); }  // voice.rs:156-158
    void render(double* out, size_t n) {  // voice.rs:162-179
        for (size_t i = 0; i < n; i++) out[i] = 0.0;
        reed.render(out, n);
        if (!noise.is_done()) noise.render(out, n);
        pickup.process(out, n);
        const double g = post_pickup_gain;
        for (size_t i = 0; i < n; i++) out[i] *= g;
    }
    bool is_silent() const {  // voice.rs:183-188
        if (reed.damper_active && reed.release_seconds(sample_rate) > 10.0) return true;
        return reed.is_silent(-80.0);
    }
};

// Voice::render_note_with_scale, voice.rs:201-221 (ds_override: NaN = none)
static inline std::vector<double> render_note(uint8_t midi, double velocity, double duration_s, double sr, double ds_override) {
    const uint32_t seed = (uint32_t)midi * 2654435761u;
    Voice v;
    v.note_on(midi, velocity, sr, seed, false);
    if (ds_override == ds_override) v.pickup.displacement_scale = ds_override;
    const size_t n = (size_t)f64_as_u64(duration_s * sr);
    std::vector<double> out(n, 0.0);
    for (size_t off = 0; off < n; off += 1024) v.render(out.data() + off, std::min<size_t>(1024, n - off));
    return out;
}

}  // namespace ow
