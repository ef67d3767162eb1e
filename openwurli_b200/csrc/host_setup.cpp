// Host-side note-on parameterisation for libowgpu.
//
// Everything a Wurlitzer voice needs that is computed ONCE per note lives here and runs on the
// host with glibc libm -- the same libm Rust's std calls on Linux -- so the per-voice init records
// the kernels start from are bit-identical to what the reference's Voice::note_on produces.
// References (hal0zer0/openwurli v0.6.0, crates/openwurli-dsp/src/):
//   tables.rs:34-830   per-key physics (frequency, tip mass, beam eigenvalues, reed geometry,
//                      spatial pickup coupling, decay law, velocity curves, output scale)
//   variation.rs:10-38 per-key detune / mode-amplitude hash
//   hammer.rs:26-146   dwell filter, onset ramp time, attack-noise parameters
//   mlp_correction.rs:61-140 + mlp_weights.rs   2->16->16->11 correction MLP
//   voice.rs:28-142, reed.rs:108-182, pickup.rs:111-118   assembly into reed/pickup/noise state
//   speaker.rs:63-101, filters.rs (RBJ cookbook biquads)   speaker coefficients
// Layout differs from the reference on purpose: per-key quantities are tabulated once per
// process (KeyRow), velocity-dependent ones are evaluated per job, output is a flat POD record.
#include "host_setup.h"

#include <cmath>
#include <cstring>
#include <mutex>

namespace owg {
namespace {

#define OWC_TABLE(name) const double name
#define OWC_SCALAR(name) const double name
#include "ow_consts.inc"
#undef OWC_TABLE
#undef OWC_SCALAR

constexpr int NM = OWG_NUM_MODES;
constexpr double kPi = 3.14159265358979323846;
constexpr double kTau = 6.28318530717958647692;
constexpr double kPickupFc = 2312.0;

inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// piecewise-linear lookup with flat ends (tables.rs:52-80 and :465-503 share this shape)
double pwl(const double* xs, const double* ys, int n, double m) {
    if (m <= xs[0]) return ys[0];
    if (m >= xs[n - 1]) return ys[n - 1];
    for (int i = 0; i + 1 < n; i++) {
        if (m <= xs[i + 1]) {
            const double t = (m - xs[i]) / (xs[i + 1] - xs[i]);
            return ys[i] + t * (ys[i + 1] - ys[i]);
        }
    }
    return 0.0;
}

// tables.rs:88-150 -- beta_n(mu) by linear interpolation in a fixed table
const double kMu[8] = {0.00, 0.01, 0.05, 0.10, 0.15, 0.20, 0.30, 0.50};
const double kBeta[8][NM] = {
    {1.8751, 4.6941, 7.8548, 10.9955, 14.1372, 17.2788, 20.4204},
    {1.8584, 4.6849, 7.8504, 10.9930, 14.1356, 17.2776, 20.4195},
    {1.7920, 4.6477, 7.8316, 10.9830, 14.1288, 17.2726, 20.4158},
    {1.7227, 4.6024, 7.8077, 10.9700, 14.1198, 17.2660, 20.4110},
    {1.6625, 4.5618, 7.7859, 10.9580, 14.1114, 17.2598, 20.4065},
    {1.6097, 4.5254, 7.7659, 10.9470, 14.1036, 17.2540, 20.4023},
    {1.5201, 4.4620, 7.7310, 10.9280, 14.0894, 17.2434, 20.3946},
    {1.3853, 4.3601, 7.6745, 10.8970, 14.0650, 17.2252, 20.3814},
};
void beam_betas(double mu, double* b) {
    const double m = clampd(mu, 0.0, 0.50);
    int lo = 0;
    for (int i = 0; i < 8; i++)
        if (kMu[i] <= m) lo = i;  // last row with mu <= m
    const int hi = lo + 1 < 8 ? lo + 1 : 7;
    const double t = kMu[hi] > kMu[lo] ? (m - kMu[lo]) / (kMu[hi] - kMu[lo]) : 0.0;
    for (int i = 0; i < NM; i++) b[i] = kBeta[lo][i] + t * (kBeta[hi][i] - kBeta[lo][i]);
}

double freq_of_key(int midi) { return 440.0 * std::pow(2.0, ((double)midi - 69.0) / 12.0); }  // tables.rs:34-36

double reed_len_mm(int midi) {  // tables.rs:167-175
    const double n = clampd((double)midi - 32.0, 1.0, 64.0);
    const double inches = n <= 20.0 ? 3.0 - n / 20.0 : 2.0 - (n - 20.0) / 44.0;
    return inches * 25.4;
}

double compliance(int midi) {  // tables.rs:186-222
    int reed = midi - 32;
    reed = reed < 1 ? 1 : (reed > 64 ? 64 : reed);
    const double w_in = reed <= 14 ? 0.151 : reed <= 20 ? 0.127 : reed <= 42 ? 0.121 : reed <= 50 ? 0.111 : 0.098;
    double t_in;
    if (reed <= 16) t_in = 0.026;
    else if (reed <= 26) {
        const double t = ((double)reed - 16.0) / 10.0;
        t_in = 0.026 + t * (0.034 - 0.026);
    } else t_in = 0.034;
    const double l = reed_len_mm(midi), w = w_in * 25.4, t = t_in * 25.4;
    return (l * l * l) / (w * t * t * t);
}

double displacement_scale(int midi) {  // tables.rs:283-288 with CalibrationConfig::default()
    const double ds = 0.85 * std::pow(compliance(midi) / compliance(60), 0.75);
    return clampd(ds, 0.02, 0.95);
}

double shape(double beta, double xi) {  // tables.rs:296-300
    const double sigma = (std::cosh(beta) + std::cos(beta)) / (std::sinh(beta) + std::sin(beta));
    const double bx = beta * xi;
    return std::cosh(bx) - std::cos(bx) - sigma * (std::sinh(bx) - std::sin(bx));
}

void pickup_coupling(const double* betas, double reed_mm, double* kappa) {  // tables.rs:324-370
    const double span = clampd(6.0 / reed_mm, 0.0, 1.0);
    const double xi0 = 1.0 - span;
    double raw[NM];
    for (int m = 0; m < NM; m++) {
        const double tip = shape(betas[m], 1.0);
        if (std::fabs(tip) < 1e-30 || span < 1e-12) { raw[m] = 1.0; continue; }
        const double h = span / 32.0;
        double acc = shape(betas[m], xi0) + shape(betas[m], 1.0);
        for (int j = 1; j < 32; j++) acc += ((j & 1) ? 4.0 : 2.0) * shape(betas[m], xi0 + (double)j * h);
        const double integral = acc * h / 3.0;
        raw[m] = clampd(std::fabs(integral / (span * tip)), 0.0, 1.0);
    }
    if (raw[0] > 1e-30) for (int m = 0; m < NM; m++) kappa[m] = clampd(raw[m] / raw[0], 0.0, 1.0);
    else for (int m = 0; m < NM; m++) kappa[m] = 1.0;
}

double rms_proxy(double ds, double f0, double fc) {  // tables.rs:438-454
    if (ds < 1e-10) return 0.0;
    const double r = (1.0 - std::sqrt(1.0 - ds * ds)) / ds;
    const double inv = 1.0 / std::sqrt(1.0 - ds * ds);
    double acc = 0.0, rn = r;
    for (int n = 1; n <= 8; n++) {
        const double cn = 2.0 * rn * inv;
        const double nf = (double)n * f0;
        const double h = nf / std::sqrt(nf * nf + fc * fc);
        acc += (cn * h) * (cn * h);
        rn *= r;
    }
    return std::sqrt(acc);
}

double vel_exponent(int midi) {  // tables.rs:632-651
    const double m = (double)midi;
    const double z = (m - 62.0) / 15.0;
    const double bell = std::exp(-0.5 * (z * z));
    const double floor_exp = m < 62.0 ? 0.55 : 1.3;
    return floor_exp + bell * (1.7 - floor_exp);
}

double vel_scurve(double v) {  // tables.rs:659-665
    const double k = 1.5;
    const double s = 1.0 / (1.0 + std::exp(-k * (v - 0.5)));
    const double s0 = 1.0 / (1.0 + std::exp(k * 0.5));
    const double s1 = 1.0 / (1.0 + std::exp(-k * 0.5));
    return (s - s0) / (s1 - s0);
}

uint32_t key_hash(int midi, uint32_t salt) {  // variation.rs:10-18
    uint32_t h = 2166136261u;
    h ^= (uint32_t)(uint8_t)midi;
    h *= 16777619u;
    h ^= salt;
    h *= 16777619u;
    h ^= h >> 16;
    h *= 2654435769u;
    return h & 0x00FFFFFFu;
}

// Everything that depends on the key only.
struct KeyRow {
    bool ready = false;
    double f_nominal, f_detuned;
    double ratio[NM], amp[NM], decay_db[NM], amp_jitter[NM];
    double ds, vel_exp, trim_db, voicing_db;
};
KeyRow g_keys[256];
std::mutex g_keys_mu;
double g_vel_exp_c4 = 0.0, g_f_c4 = 0.0;

const KeyRow& key_row(int midi) {
    KeyRow& k = g_keys[midi & 255];
    std::lock_guard<std::mutex> lock(g_keys_mu);
    if (k.ready) return k;
    if (g_f_c4 == 0.0) { g_vel_exp_c4 = vel_exponent(60); g_f_c4 = freq_of_key(60); }
    k.f_nominal = freq_of_key(midi);
    {  // variation.rs:26-29
        const double r = ((double)key_hash(midi, 0xDEAD) / 16777216.0) * 2.0 - 1.0;
        k.f_detuned = k.f_nominal * (1.0 + r * 0.00173);
    }
    static const double mx[5] = {33.0, 52.0, 62.0, 74.0, 96.0}, my[5] = {0.10, 0.00, 0.00, 0.02, 0.01};
    const double mu = pwl(mx, my, 5, (double)midi);
    double betas[NM];
    beam_betas(mu, betas);
    const double b1sq = betas[0] * betas[0];
    for (int i = 0; i < NM; i++) k.ratio[i] = (betas[i] * betas[i]) / b1sq;
    const double base_decay = std::fmax(0.005 * std::pow(k.f_nominal, 1.22), 3.0);  // tables.rs:384-387
    for (int i = 0; i < NM; i++) k.decay_db[i] = base_decay * k.ratio[i] * k.ratio[i];
    static const double base_amp[NM] = {1.0, 0.005, 0.0035, 0.0018, 0.0011, 0.0007, 0.0005};
    double kappa[NM];
    pickup_coupling(betas, reed_len_mm(midi), kappa);
    for (int i = 0; i < NM; i++) k.amp[i] = base_amp[i] * kappa[i];
    for (int i = 0; i < NM; i++) {  // variation.rs:33-38
        const double r = ((double)key_hash(midi, 0xBEEFu + (uint32_t)i) / 16777216.0) * 2.0 - 1.0;
        k.amp_jitter[i] = 1.0 + r * 0.08;
    }
    k.ds = displacement_scale(midi);
    k.vel_exp = vel_exponent(midi);
    static const double tx[13] = {36, 40, 44, 48, 52, 56, 60, 64, 68, 72, 76, 80, 84};
    static const double ty[13] = {-1.3, 0.0, -1.3, 0.7, 0.2, -1.0, 0.0, 0.9, 1.2, 0.0, 1.8, 2.4, 3.6};
    k.trim_db = pwl(tx, ty, 13, (double)midi);
    k.voicing_db = -0.04 * std::fmax((double)midi - 60.0, 0.0);
    k.ready = true;
    return k;
}

// mlp_correction.rs:61-140
struct Corrections { double cents[5], decay[5], ds; };
Corrections corrections(int midi, double velocity, bool enabled) {
    Corrections c;
    for (int i = 0; i < 5; i++) { c.cents[i] = 0.0; c.decay[i] = 1.0; }
    c.ds = 1.0;
    if (!enabled) return c;
    const double m = (double)midi;
    double fade = 1.0;
    if (m < 65.0) fade = clampd((m - (65.0 - 12.0)) / 12.0, 0.0, 1.0);
    else if (m > 97.0) fade = clampd(((97.0 + 12.0) - m) / 12.0, 0.0, 1.0);
    if (fade <= 0.0) return c;
    const double x[2] = {clampd((m - 21.0) / (108.0 - 21.0), 0.0, 1.0), clampd(velocity, 0.0, 1.0)};
    double a[16], b[16], y[11];
    for (int i = 0; i < 16; i++) {
        double s = MLP_B1[i];
        for (int j = 0; j < 2; j++) s += MLP_W1[i][j] * x[j];
        a[i] = s > 0.0 ? s : 0.0;
    }
    for (int i = 0; i < 16; i++) {
        double s = MLP_B2[i];
        for (int j = 0; j < 16; j++) s += MLP_W2[i][j] * a[j];
        b[i] = s > 0.0 ? s : 0.0;
    }
    for (int i = 0; i < 11; i++) {
        double s = MLP_B3[i];
        for (int j = 0; j < 16; j++) s += MLP_W3[i][j] * b[j];
        y[i] = s * MLP_TARGET_STDS[i] + MLP_TARGET_MEANS[i];
    }
    for (int h = 0; h < 5; h++) c.cents[h] = clampd(y[h] * fade, -100.0, 100.0);
    for (int h = 0; h < 5; h++) c.decay[h] = 1.0 + (clampd(y[5 + h], 0.3, 3.0) - 1.0) * fade;
    c.ds = 1.0 + (clampd(y[10], 0.7, 1.2) - 1.0) * fade;
    return c;
}

double output_gain(const KeyRow& k, int midi, double velocity) {  // tables.rs:578-616, default CalibrationConfig
    const double sv = vel_scurve(velocity);
    const double vs = std::pow(sv, k.vel_exp);
    const double vs_c4 = std::pow(sv, g_vel_exp_c4);
    const double eds = std::fmax(k.ds * vs, 1e-6);
    const double eds_ref = std::fmax(0.85 * vs_c4, 1e-6);
    const double rms = rms_proxy(eds, k.f_nominal, kPickupFc);
    const double rms_ref = rms_proxy(eds_ref, g_f_c4, kPickupFc);
    const double flat_db = -20.0 * std::log10(rms / rms_ref);
    const double eff_trim = k.trim_db * std::pow(velocity, 1.3);
    (void)midi;
    return std::pow(10.0, (-35.0 + flat_db + k.voicing_db + eff_trim) / 20.0);
}

struct Rbj { double b0, b1, b2, a1, a2; };
// melange-primitives Biquad (un-vendored; restated from the RBJ Audio EQ Cookbook per filters.rs:1-11)
Rbj rbj(int kind /*0 LP 1 HP 2 BP-skirt*/, double fc, double q, double fs) {
    const double w0 = 2.0 * kPi * fc / fs;
    const double cw = std::cos(w0), sw = std::sin(w0);
    const double alpha = sw / (2.0 * q);
    double n0, n1, n2;
    if (kind == 0) { n0 = (1.0 - cw) / 2.0; n1 = 1.0 - cw; n2 = (1.0 - cw) / 2.0; }
    else if (kind == 1) { n0 = (1.0 + cw) / 2.0; n1 = -(1.0 + cw); n2 = (1.0 + cw) / 2.0; }
    else { n0 = sw / 2.0; n1 = 0.0; n2 = -sw / 2.0; }
    const double a0 = 1.0 + alpha;
    return Rbj{n0 / a0, n1 / a0, n2 / a0, (-2.0 * cw) / a0, (1.0 - alpha) / a0};
}

inline uint32_t lcg(uint32_t s) { return s * 1664525u + 1013904223u; }
inline uint64_t trunc_u64(double x) { return !(x == x) || x <= 0.0 ? 0 : (x >= 18446744073709551615.0 ? UINT64_MAX : (uint64_t)x); }
inline uint32_t trunc_u32(double x) { return !(x == x) || x <= 0.0 ? 0 : (x >= 4294967295.0 ? 4294967295u : (uint32_t)x); }

}  // namespace

void make_voice_init(const owg_voice_job& job, OwgVoiceInit* out) {
    const int midi = job.midi;
    const double vel = job.velocity, fs = job.sample_rate;
    const KeyRow& k = key_row(midi);
    std::memset(out, 0, sizeof(*out));

    // hammer.rs:69-90 dwell filter on the detuned fundamental; hammer.rs:53-57 onset time
    const double t_dwell = clampd((0.75 + 0.25 * (1.0 - vel)) / k.f_detuned, 0.0003, 0.020);
    double dwell[NM];
    for (int i = 0; i < NM; i++) {
        const double ft = k.f_detuned * k.ratio[i] * t_dwell;
        dwell[i] = std::exp(-ft * ft / (2.0 * (8.0 * 8.0)));
    }
    const double d0 = dwell[0];
    if (d0 > 1e-30) for (int i = 0; i < NM; i++) dwell[i] /= d0;
    const double onset_s = (job.flags & OWG_VOICE_NO_ONSET) ? 0.0 : std::fmax((1.0 + 1.0 * (1.0 - vel)) * (1.0 / k.f_detuned), 0.002);

    // voice.rs:44-57 amplitudes
    const double vel_scale = std::pow(vel_scurve(vel), k.vel_exp);
    double amp[NM];
    for (int i = 0; i < NM; i++) amp[i] = k.amp[i] * dwell[i] * k.amp_jitter[i];
    for (int i = 0; i < NM; i++) amp[i] *= vel_scale;

    // voice.rs:62-88 MLP corrections on modes 1..5 and the displacement scale
    const Corrections c = corrections(midi, vel, job.mlp_enabled != 0);
    double ratio[NM], decay_db[NM];
    for (int i = 0; i < NM; i++) { ratio[i] = k.ratio[i]; decay_db[i] = k.decay_db[i]; }
    for (int i = 1; i < 6; i++) ratio[i] *= std::pow(2.0, c.cents[i - 1] / 1200.0);
    for (int i = 1; i < 6; i++) decay_db[i] /= c.decay[i - 1];
    const double ds_corr = k.ds * c.ds;

    // reed.rs:118-134 OU coefficients and Box-Muller initial drifts
    const double dt = 1.0 / fs;
    out->jitter_revert = std::exp(-dt / 0.020);
    out->jitter_diffusion = 0.0004 * std::sqrt(1.0 - out->jitter_revert * out->jitter_revert);
    uint32_t js = job.noise_seed > 1u ? job.noise_seed : 1u;
    for (int i = 0; i < NM; i++) {
        js = lcg(js);
        const double u1 = (double)(js >> 1) / (4294967295.0 / 2.0);
        js = lcg(js);
        const double u2 = (double)(js >> 1) / (4294967295.0 / 2.0);
        const double r = std::sqrt(-2.0 * std::log(std::fmax(u1, 1e-30)));
        out->jitter_drift[i] = 0.0004 * r * std::cos(kTau * u2);
    }
    out->jitter_state = js;
    // reed.rs:137-156 per-mode rotation and decay
    for (int i = 0; i < NM; i++) {
        const double f = k.f_detuned * ratio[i];
        const double inc = kTau * f / fs;
        out->phase_inc[i] = inc;
        out->cos_inc[i] = std::cos(inc);
        out->sin_inc[i] = std::sin(inc);
        out->amplitude[i] = amp[i];
        out->decay_mult[i] = std::exp(-((decay_db[i] / 8.686) / fs));
    }
    // reed.rs:159-166 onset ramp
    out->onset_ramp_samples = trunc_u64(std::round(onset_s * fs));
    out->onset_ramp_inc = out->onset_ramp_samples > 0 ? kPi / (double)out->onset_ramp_samples : 0.0;
    out->onset_shape_exp = 1.0 + (1.0 - vel);

    // pickup.rs:111-118, voice.rs:100-101
    out->pickup_beta = dt / (2.0 * (287.0e3 * 240.0e-12));
    out->pickup_ds = ds_corr;
    if (job.ds_override == job.ds_override) out->pickup_ds = job.ds_override;

    // hammer.rs:126-146
    out->noise_amp = 0.025 * vel * vel;
    out->noise_decay = std::exp(-1.0 / (0.003 * fs));
    out->noise_remaining = job.attack_noise ? trunc_u32(0.015 * fs) : 0u;
    out->noise_rng = job.noise_seed;
    const Rbj bp = rbj(2, clampd(k.f_detuned * 5.0, 200.0, 2000.0), 0.7, fs);
    out->bq_b0 = bp.b0; out->bq_b1 = bp.b1; out->bq_b2 = bp.b2; out->bq_a1 = bp.a1; out->bq_a2 = bp.a2;

    // voice.rs:106-132 post-pickup gain with MLP level compensation
    double comp = 1.0;
    if (std::fabs(c.ds - 1.0) > 1e-6) {
        const double pb = rms_proxy(k.ds, k.f_nominal, kPickupFc);
        const double pc = rms_proxy(ds_corr, k.f_nominal, kPickupFc);
        comp = pc > 1e-10 ? std::sqrt(pb / pc) : 1.0;
    }
    out->post_pickup_gain = output_gain(k, midi, vel) * comp;

    out->sample_rate = fs;
    out->n_samples = trunc_u64(job.duration_s * fs);
    out->midi = job.midi;
}

void make_chain_init(const owg_bench_job& job, int group, OwgChainInit* out) {
    std::memset(out, 0, sizeof(*out));
    const double fs = job.v.sample_rate;
    out->volume = job.volume;
    // Speaker::new then set_character (speaker.rs:63-101): constructed at character 1.0; a change
    // within +-0.002 of 1.0 is ignored.
    double ch = 1.0;
    const double want = clampd(job.speaker_character, 0.0, 1.0);
    if (std::fabs(want - ch) > 0.002) ch = want;
    const double hpf_hz = 20.0 * std::pow(30.0 / 20.0, ch);
    const double lpf_hz = 20000.0 * std::pow(5500.0 / 20000.0, ch);
    const Rbj hp = rbj(1, hpf_hz, 0.75, fs), lp = rbj(0, lpf_hz, 0.707, fs);
    out->hpf_b0 = hp.b0; out->hpf_b1 = hp.b1; out->hpf_b2 = hp.b2; out->hpf_a1 = hp.a1; out->hpf_a2 = hp.a2;
    out->lpf_b0 = lp.b0; out->lpf_b1 = lp.b1; out->lpf_b2 = lp.b2; out->lpf_a1 = lp.a1; out->lpf_a2 = lp.a2;
    out->spk_a2 = 0.2 * ch;
    out->spk_a3 = 0.6 * ch;
    out->spk_norm = 1.0 + out->spk_a2 + out->spk_a3;
    out->spk_thermal_coeff = 2.0 * ch;
    out->spk_thermal_alpha = 1.0 / (5.0 * fs);
    out->spk_tanh = ch < 0.001 ? 0 : 1;
    out->group = group;
    out->no_preamp = job.no_preamp;
    out->no_poweramp = job.no_poweramp;
    out->oversample = fs < 88200.0 ? 1 : 0;
    out->pre_only = 0;
}

void make_damper_rows(double sample_rate, DamperRow* rows) {
    for (int midi = 0; midi < 128; midi++) {
        DamperRow& r = rows[midi];
        std::memset(&r, 0, sizeof(r));
        if (midi >= 92) { r.enabled = 0; continue; }  // top keys: no damper (reed.rs:193-195)
        r.enabled = 1;
        const double base_rate = std::fmax(55.0 * std::pow(2.0, ((double)midi - 60.0) / 24.0), 0.5);
        double p3 = 1.0;  // 3.0.powi(m): exact
        for (int m = 0; m < NM; m++) {
            const double factor = std::fmin(base_rate * p3, 2000.0);
            r.rate[m] = factor / sample_rate;
            r.mult[m] = std::exp(-r.rate[m]);
            p3 *= 3.0;
        }
        const double ramp_time = midi < 48 ? 0.050 : (midi < 72 ? 0.025 : 0.008);
        r.ramp_samples = ramp_time * sample_rate;
    }
}

namespace {
void fill_spk(SpkUpdate* u, int64_t at, double ch, double fs) {
    const double hpf_hz = 20.0 * std::pow(30.0 / 20.0, ch);
    const double lpf_hz = 20000.0 * std::pow(5500.0 / 20000.0, ch);
    const Rbj hp = rbj(1, hpf_hz, 0.75, fs), lp = rbj(0, lpf_hz, 0.707, fs);
    std::memset(u, 0, sizeof(*u));
    u->at = at;
    u->a2 = 0.2 * ch; u->a3 = 0.6 * ch; u->norm = 1.0 + u->a2 + u->a3; u->thermal_coeff = 2.0 * ch;
    u->hpf_b0 = hp.b0; u->hpf_b1 = hp.b1; u->hpf_b2 = hp.b2; u->hpf_a1 = hp.a1; u->hpf_a2 = hp.a2;
    u->lpf_b0 = lp.b0; u->lpf_b1 = lp.b1; u->lpf_b2 = lp.b2; u->lpf_a1 = lp.a1; u->lpf_a2 = lp.a2;
    u->tanh_on = ch < 0.001 ? 0 : 1;
}
}  // namespace

int make_speaker_schedule(double fs, double target, int64_t n_warm, int64_t n_total, uint32_t ramp, SpkUpdate* out, int max_out) {
    int n = 0;
    double character = 1.0;  // Speaker::new (speaker.rs:63-78)
    if (n < max_out) fill_spk(&out[n++], -1, character, fs);
    // LinearSmoother::new(0.0, ramp) (engine.rs:226)
    double cur = 0.0, tgt = 0.0, step = 0.0;
    uint32_t remaining = 0;
    // the smoother is flat except for `ramp` samples after the target is set; simulate sample 0, then the ramp window
    for (int64_t t = 0; t < n_total; t++) {
        if (t == n_warm) {  // set_speaker_character(target) -> LinearSmoother::set_target (engine.rs:85-98)
            if (!(std::fabs(target - tgt) < 1e-9)) {
                tgt = target;
                const double delta = tgt - cur;
                if (ramp == 0) { cur = tgt; remaining = 0; }
                else { step = delta / (double)ramp; remaining = ramp; }
            }
        }
        if (remaining > 0) {
            cur += step;
            remaining -= 1;
            if (remaining == 0) cur = tgt;
        }
        const double c = clampd(cur, 0.0, 1.0);  // Speaker::set_character (speaker.rs:81-87)
        if (std::fabs(c - character) > 0.002) {
            character = c;
            if (n < max_out) fill_spk(&out[n++], t, character, fs);
            else return -1;
        }
        if (remaining == 0 && t > n_warm) break;  // nothing can change any more
        if (remaining == 0 && t < n_warm) t = n_warm - 1;  // skip the flat part of the warm-up
    }
    return n;
}

int make_engine_schedule(double fs, const AutoEvent* chr, int n_chr, const AutoEvent* vol, int n_vol, int64_t n_total, uint32_t ramp, SpkUpdate* out,
                         int max_out) {
    int n = 0;
    double character = 1.0;  // Speaker::new (speaker.rs:63-78)
    if (n < max_out) fill_spk(&out[n++], -1, character, fs);
    double cur = 0.0, tgt = 0.0, step = 0.0;  // LinearSmoother::new(0.0, ramp) (engine.rs:226)
    uint32_t remaining = 0;
    int ci = 0, vi = 0;
    int64_t t = 0;
    while (t < n_total) {
        while (vi < n_vol && vol[vi].at <= t) {  // set_volume at the start of the block: its own schedule entry
            if (n >= max_out) return -1;
            std::memset(&out[n], 0, sizeof(SpkUpdate));
            out[n].at = vol[vi].at; out[n].a2 = vol[vi].target; out[n]._pad = 1;
            n++; vi++;
        }
        while (ci < n_chr && chr[ci].at <= t) {  // set_speaker_character -> LinearSmoother::set_target (engine.rs:85-98)
            const double target = chr[ci++].target;
            if (!(std::fabs(target - tgt) < 1e-9)) {
                tgt = target;
                const double delta = tgt - cur;
                if (ramp == 0) { cur = tgt; remaining = 0; }
                else { step = delta / (double)ramp; remaining = ramp; }
            }
        }
        if (remaining > 0) {
            cur += step;
            remaining -= 1;
            if (remaining == 0) cur = tgt;
        }
        const double c = clampd(cur, 0.0, 1.0);  // Speaker::set_character (speaker.rs:81-87)
        if (std::fabs(c - character) > 0.002) {
            character = c;
            if (n < max_out) fill_spk(&out[n++], t, character, fs);
            else return -1;
        }
        if (remaining == 0) {  // flat until the next event: jump there
            int64_t nxt = n_total;
            if (ci < n_chr) nxt = std::min<int64_t>(nxt, chr[ci].at);
            if (vi < n_vol) nxt = std::min<int64_t>(nxt, vol[vi].at);
            t = std::max<int64_t>(t + 1, nxt);
        } else t++;
    }
    return n;
}

// ---- legacy 8-node preamp: plan-time constants ----------------------------------------------------------------------
namespace {
enum LgNode { N_BASE1 = 0, N_EMIT1, N_COLL1, N_EMIT2, N_EMIT2B, N_COLL2, N_OUT, N_FB, LGN };
struct LgBranch { int a, b; double value; };  // two-terminal element between nodes a and b (b < 0: to a supply / ground)
const double LG_VCC = 15.0, LG_IS = 3.03e-14, LG_VT = 0.026, LG_VBE_MAX = 0.85;

// Gauss-Jordan with partial pivoting on [M | I] (dk_preamp_legacy.rs:122-168): the operation order fixes the rounding.
void lg_invert(const double (*m)[LGN], double (*inv)[LGN]) {
    double w[LGN][2 * LGN];
    for (int r = 0; r < LGN; r++)
        for (int c = 0; c < LGN; c++) { w[r][c] = m[r][c]; w[r][LGN + c] = r == c ? 1.0 : 0.0; }
    for (int col = 0; col < LGN; col++) {
        int piv = col;
        double best = std::fabs(w[col][col]);
        for (int r = col + 1; r < LGN; r++)
            if (std::fabs(w[r][col]) > best) { best = std::fabs(w[r][col]); piv = r; }
        if (piv != col)
            for (int c = 0; c < 2 * LGN; c++) { const double t = w[col][c]; w[col][c] = w[piv][c]; w[piv][c] = t; }
        const double p = w[col][col];
        for (int c = 0; c < 2 * LGN; c++) w[col][c] /= p;
        for (int r = 0; r < LGN; r++) {
            if (r == col) continue;
            const double f = w[r][col];
            for (int c = 0; c < 2 * LGN; c++) w[r][c] -= f * w[col][c];
        }
    }
    for (int r = 0; r < LGN; r++)
        for (int c = 0; c < LGN; c++) inv[r][c] = w[r][LGN + c];
}
void lg_matvec(const double (*m)[LGN], const double* x, double* y) {
    for (int r = 0; r < LGN; r++) {
        double acc = 0.0;
        for (int c = 0; c < LGN; c++) acc += m[r][c] * x[c];
        y[r] = acc;
    }
}
// K = N_v S N_i with N_v rows (BASE1-EMIT1, COLL1-EMIT2), N_i columns (EMIT1-COLL1, EMIT2-COLL2)  (:424-435)
void lg_kernel(const double (*s)[LGN], double* k4) {
    k4[0] = s[N_BASE1][N_EMIT1] - s[N_BASE1][N_COLL1] - s[N_EMIT1][N_EMIT1] + s[N_EMIT1][N_COLL1];
    k4[1] = s[N_BASE1][N_EMIT2] - s[N_BASE1][N_COLL2] - s[N_EMIT1][N_EMIT2] + s[N_EMIT1][N_COLL2];
    k4[2] = s[N_COLL1][N_EMIT1] - s[N_COLL1][N_COLL1] - s[N_EMIT2][N_EMIT1] + s[N_EMIT2][N_COLL1];
    k4[3] = s[N_COLL1][N_EMIT2] - s[N_COLL1][N_COLL2] - s[N_EMIT2][N_EMIT2] + s[N_EMIT2][N_COLL2];
}
double lg_clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
double lg_ic(double vbe) { return LG_IS * (std::exp(lg_clampd(vbe, -1.0, LG_VBE_MAX) / LG_VT) - 1.0); }  // :668-671
}  // namespace

void make_legacy_group(double fs, double r_static, double* rec, bool dc_at_r) {
    // circuit (dk_preamp_legacy.rs:24-41, stamps :281-310), in the reference's stamping order (the sums into a diagonal
    // entry are order-sensitive in the last bit)
    const LgBranch resistors[] = {
        {N_BASE1, -2, 2000000.0},  // R2 to Vcc
        {N_BASE1, -1, 470000.0},   // R3 to ground
        {N_EMIT1, -1, 33000.0},    // Re1
        {N_COLL1, -2, 150000.0},   // Rc1 to Vcc
        {N_EMIT2, N_EMIT2B, 270.0},
        {N_EMIT2B, -1, 820.0},
        {N_COLL2, -2, 1800.0},     // Rc2 to Vcc
        {N_COLL2, N_OUT, 6800.0},  // R9
        {N_OUT, N_FB, 56000.0},    // R10
    };
    const LgBranch caps[] = {{N_COLL1, N_BASE1, 100.0e-12}, {N_COLL2, N_COLL1, 100.0e-12}, {N_EMIT1, N_FB, 4.7e-6}, {N_EMIT2, N_EMIT2B, 22.0e-6}};
    const double r1 = 22000.0, cin = 0.022e-6;
    double g[LGN][LGN] = {{0}}, c[LGN][LGN] = {{0}}, w[LGN] = {0};
    for (const LgBranch& e : resistors) {
        const double y = 1.0 / e.value;
        g[e.a][e.a] += y;
        if (e.b >= 0) { g[e.b][e.b] += y; g[e.a][e.b] -= y; g[e.b][e.a] -= y; }
        else if (e.b == -2) w[e.a] += LG_VCC / e.value;
    }
    for (const LgBranch& e : caps) { c[e.a][e.a] += e.value; c[e.b][e.b] += e.value; c[e.a][e.b] -= e.value; c[e.b][e.a] -= e.value; }
    double g_dc[LGN][LGN];
    std::memcpy(g_dc, g, sizeof(g));
    // Cin in series with R1 as a bilinear companion (:273-277)
    const double t_s = 1.0 / fs;
    const double two_over_t = 2.0 / t_s;
    const double alpha = 2.0 * r1 * cin * fs;
    const double g_cin = (2.0 * cin * fs) / (1.0 + alpha);
    const double c_cin = (1.0 - alpha) / (1.0 + alpha);
    g[N_BASE1][N_BASE1] += g_cin;
    double a_pos[LGN][LGN], s_base[LGN][LGN];
    for (int r = 0; r < LGN; r++)
        for (int k = 0; k < LGN; k++) {
            const double ct = two_over_t * c[r][k];
            a_pos[r][k] = ct + g[r][k];
            rec[OWG_LG_AN + r * LGN + k] = ct - g[r][k];
        }
    lg_invert(a_pos, s_base);
    for (int r = 0; r < LGN; r++) {
        for (int k = 0; k < LGN; k++) rec[OWG_LG_S + r * LGN + k] = s_base[r][k];
        rec[OWG_LG_W2 + r] = 2.0 * w[r];
        rec[OWG_LG_SFB + r] = s_base[r][N_FB];
        rec[OWG_LG_D0 + r] = s_base[r][N_EMIT1] - s_base[r][N_COLL1];
        rec[OWG_LG_D1 + r] = s_base[r][N_EMIT2] - s_base[r][N_COLL2];
    }
    lg_kernel(s_base, rec + OWG_LG_K);
    rec[OWG_LG_NVSFB + 0] = s_base[N_BASE1][N_FB] - s_base[N_EMIT1][N_FB];
    rec[OWG_LG_NVSFB + 1] = s_base[N_COLL1][N_FB] - s_base[N_EMIT2][N_FB];
    rec[OWG_LG_SFBNI + 0] = s_base[N_FB][N_EMIT1] - s_base[N_FB][N_COLL1];
    rec[OWG_LG_SFBNI + 1] = s_base[N_FB][N_EMIT2] - s_base[N_FB][N_COLL2];
    rec[OWG_LG_SFBFB] = s_base[N_FB][N_FB];
    rec[OWG_LG_GCIN] = g_cin;
    rec[OWG_LG_GC1PC] = g_cin * (1.0 + c_cin);
    rec[OWG_LG_CCIN] = c_cin;
    // DC operating point at R_ldr = 1 MOhm (full_dc_solve, :370-412): Newton on the 2x2 kernel of the resistive network;
    // with dc_at_r (`set_ldr_resistance(r); reset()`, :620-642) at the resistance set_ldr_resistance left in r_ldr
    double r_init = 1000000.0;
    if (dc_at_r && r_static == r_static) {
        const double nr = r_static > 1000.0 ? r_static : 1000.0;
        if (std::fabs(nr - r_init) > 0.01) r_init = nr;
    }
    double g_full[LGN][LGN], s_dc[LGN][LGN], k_dc[4], sv[LGN];
    std::memcpy(g_full, g_dc, sizeof(g_dc));
    g_full[N_FB][N_FB] += 1.0 / r_init;
    lg_invert(g_full, s_dc);
    lg_kernel(s_dc, k_dc);
    lg_matvec(s_dc, w, sv);
    const double p0 = sv[N_BASE1] - sv[N_EMIT1], p1 = sv[N_COLL1] - sv[N_EMIT2];
    double vbe0 = 0.56, vbe1 = 0.66;
    for (int it = 0; it < 100; it++) {
        const double e0 = std::exp(lg_clampd(vbe0, -1.0, LG_VBE_MAX) / LG_VT), e1 = std::exp(lg_clampd(vbe1, -1.0, LG_VBE_MAX) / LG_VT);
        const double ic0 = LG_IS * (e0 - 1.0), gm0 = (LG_IS / LG_VT) * e0, ic1 = LG_IS * (e1 - 1.0), gm1 = (LG_IS / LG_VT) * e1;
        const double f0 = vbe0 - p0 - k_dc[0] * ic0 - k_dc[1] * ic1;
        const double f1 = vbe1 - p1 - k_dc[2] * ic0 - k_dc[3] * ic1;
        if (std::fabs(f0) < 1e-12 && std::fabs(f1) < 1e-12) break;
        const double j00 = 1.0 - k_dc[0] * gm0, j01 = -k_dc[1] * gm1, j10 = -k_dc[2] * gm0, j11 = 1.0 - k_dc[3] * gm1;
        const double inv_det = 1.0 / (j00 * j11 - j01 * j10);
        const double d0 = inv_det * (j11 * f0 - j01 * f1), d1 = inv_det * (j00 * f1 - j10 * f0);
        const double lim = 2.0 * LG_VT;
        vbe0 -= lg_clampd(d0, -lim, lim);
        vbe1 -= lg_clampd(d1, -lim, lim);
    }
    const double ic0 = lg_ic(vbe0), ic1 = lg_ic(vbe1);
    double rhs[LGN], v_dc[LGN];
    for (int r = 0; r < LGN; r++) rhs[r] = w[r];
    rhs[N_EMIT1] += ic0; rhs[N_COLL1] -= ic0; rhs[N_EMIT2] += ic1; rhs[N_COLL2] -= ic1;
    lg_matvec(s_dc, rhs, v_dc);
    for (int r = 0; r < LGN; r++) rec[OWG_LG_V0 + r] = v_dc[r];
    rec[OWG_LG_INL0] = ic0; rec[OWG_LG_INL0 + 1] = ic1;
    rec[OWG_LG_VNL0] = vbe0; rec[OWG_LG_VNL0 + 1] = vbe1;
    rec[OWG_LG_JCIN0] = g_cin * v_dc[N_BASE1];
    rec[OWG_LG_CINPREV0] = g_cin * v_dc[N_BASE1];
    // `reset(); set_ldr_resistance(r)` (main.rs:438-439; :620-626): f64::max(r, 1000) and the 0.01 Ohm change threshold
    rec[OWG_LG_GINIT] = 1.0 / r_init;
    double g_static = 1.0 / r_init;
    if (!dc_at_r && r_static == r_static) {
        const double nr = r_static > 1000.0 ? r_static : 1000.0;
        if (std::fabs(nr - r_init) > 0.01) g_static = 1.0 / nr;
    }
    rec[OWG_LG_GSTATIC] = g_static;
}

double register_trim_db(int midi) { return key_row(midi).trim_db; }

double calib_displacement_scale(int midi, const owg_calib_cfg& cfg) {  // tables.rs:283-288
    const double ds = cfg.ds_at_c4 * std::pow(compliance(midi) / compliance(60), cfg.ds_exponent);
    return clampd(ds, cfg.ds_clamp_lo, cfg.ds_clamp_hi);
}

double calib_output_scale(int midi, double velocity, const owg_calib_cfg& cfg) {  // tables.rs:578-616
    const KeyRow& k = key_row(midi);
    const double ds = calib_displacement_scale(midi, cfg);
    const double sv = vel_scurve(velocity);
    const double vs = std::pow(sv, k.vel_exp);
    const double vs_c4 = std::pow(sv, g_vel_exp_c4);
    const double eds = std::fmax(ds * vs, 1e-6);
    const double eds_ref = std::fmax(cfg.ds_at_c4 * vs_c4, 1e-6);
    const double rms = rms_proxy(eds, k.f_nominal, kPickupFc);
    const double rms_ref = rms_proxy(eds_ref, g_f_c4, kPickupFc);
    const double flat_db = -20.0 * std::log10(rms / rms_ref);
    const double voicing_db = cfg.voicing_slope * std::fmax((double)midi - 60.0, 0.0);
    const double trim = cfg.zero_trim ? 0.0 : k.trim_db;
    const double eff_trim = trim * std::pow(velocity, 1.3);
    return std::pow(10.0, (cfg.target_db + flat_db + voicing_db + eff_trim) / 20.0);
}

double note_frequency(int midi) { return freq_of_key(midi); }

double silent_threshold() { return std::pow(10.0, -80.0 / 20.0); }

void noise_fade_table(double* t16) {  // hammer.rs:161-168: 0.5*(1-cos(pi*pos/16)), pos=0..15
    for (int pos = 0; pos < 16; pos++) t16[pos] = 0.5 * (1.0 - std::cos(kPi * ((double)pos / 16.0)));
}

void rbj_coefficients(int kind, double fc, double q, double fs, double* out5) {
    const Rbj r = rbj(kind, fc, q, fs);
    out5[0] = r.b0; out5[1] = r.b1; out5[2] = r.b2; out5[3] = r.a1; out5[4] = r.a2;
}

}  // namespace owg
