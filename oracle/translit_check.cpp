// TEST INFRASTRUCTURE: pins the hand-written oracle to the reference's generated solvers.
//
// oracle/_ref/gen_preamp.hpp and gen_tremolo.hpp are MECHANICAL transliterations of the reference's gen_preamp.rs / gen_tremolo.rs
// (tools/transliterate_gen.py: a Rust-subset parser printing C++, no hand-restated line).  This program drives them and the
// hand restatement (ow_preamp.hpp, ow_tremolo.hpp -- what every GPU parity test is checked against) with identical inputs and
// demands BIT-identical outputs and next states:
//   1. random states: sample rate, LDR resistance, node voltages / junction currents around (and far from) the DC point, inputs
//      from microvolts to beyond the clamp, non-finite inputs, active BE cooldown -- one step plus three follow-up steps each;
//   2. trajectories: long runs with swept sine amplitude (to 60 V), per-sample R_ldr modulation, NaN / inf samples.
// It also reports how often the reference's guard paths fired (BE fallback, damping, NaN reset, max iterations), so the
// comparison is known to have exercised them.  Exit code 0 = all identical.  Usage: translit_check [trials] [trajectory_samples]
#include "ow_tremolo.hpp"
#include "ow_preamp.hpp"
#include "_ref/gen_preamp.hpp"
#include "_ref/gen_tremolo.hpp"
#include <cstdio>
#include <cstdlib>

static uint64_t g_rng = 0x243F6A8885A308D3ull;
static uint64_t rnd64() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return g_rng; }
static double rnd01() { return (double)(rnd64() >> 11) / 9007199254740992.0; }
static double rnd_sym() { return 2.0 * rnd01() - 1.0; }
static double rnd_log(double lo, double hi) { return std::exp(std::log(lo) + rnd01() * (std::log(hi) - std::log(lo))); }
static bool same(double a, double b) { uint64_t x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8); return x == y || (a != a && b != b); }

static long g_bad = 0;
static void expect(bool ok, const char* what, long trial) {
    if (!ok) { if (g_bad < 20) std::printf("MISMATCH %s (trial %ld)\n", what, trial); g_bad++; }
}

// ---- preamp ----------------------------------------------------------------------------------------------------------------------
static void compare_pre(const ow::pre::CircuitState& a, const gen_preamp::CircuitState& b, long trial) {
    using namespace ow::pre;
    for (int i = 0; i < N; i++) expect(same(a.v_prev[i], b.v_prev[i]), "pre v_prev", trial);
    for (int i = 0; i < M; i++) { expect(same(a.i_nl_prev[i], b.i_nl_prev[i]), "pre i_nl_prev", trial); expect(same(a.i_nl_prev_prev[i], b.i_nl_prev_prev[i]), "pre i_nl_prev_prev", trial); }
    expect(same(a.input_prev, b.input_prev), "pre input_prev", trial);
    expect(a.be_cooldown == b.be_cooldown, "pre be_cooldown", trial);
    expect(a.last_nr_iterations == b.last_nr_iterations, "pre last_nr_iterations", trial);
    expect(same(a.pot_0_resistance, b.pot_0_resistance), "pre pot_0_resistance", trial);
    expect(a.matrices_dirty == b.matrices_dirty, "pre matrices_dirty", trial);
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { expect(same(a.s[i][j], b.s[i][j]), "pre S", trial); expect(same(a.a_neg[i][j], b.a_neg[i][j]), "pre A_neg", trial); }
    for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) expect(same(a.k[i][j], b.k[i][j]), "pre K", trial);
    for (int i = 0; i < N; i++) for (int j = 0; j < M; j++) expect(same(a.s_ni[i][j], b.s_ni[i][j]), "pre S_NI", trial);
}

static void poke_pre(ow::pre::CircuitState& a, gen_preamp::CircuitState& b) {
    using namespace ow::pre;
    const double scale = rnd01() < 0.15 ? rnd_log(1.0, 80.0) : rnd_log(1e-7, 1.0);
    for (int i = 0; i < N; i++) { const double v = ow::PRE_DC_OP[i] + scale * rnd_sym(); a.v_prev[i] = v; b.v_prev[i] = v; }
    const double iscale = rnd01() < 0.1 ? rnd_log(1e-3, 1.0) : rnd_log(1e-9, 1e-3);
    for (int i = 0; i < M; i++) {
        const double c = ow::PRE_DC_NL_I[i] * (1.0 + 0.1 * rnd_sym()) + iscale * rnd_sym() * 1e-3;
        const double d = c + iscale * rnd_sym() * 1e-4;
        a.i_nl_prev[i] = c; b.i_nl_prev[i] = c; a.i_nl_prev_prev[i] = d; b.i_nl_prev_prev[i] = d;
    }
    const double xp = rnd_log(1e-6, 10.0) * rnd_sym();
    a.input_prev = xp; b.input_prev = xp;
    const uint32_t cd = rnd01() < 0.05 ? (uint32_t)(rnd64() % 65) : 0u;
    a.be_cooldown = cd; b.be_cooldown = cd;
    if (rnd01() < 0.01) {  // a non-finite state entry: the only way into the NaN reset (gen_preamp.rs:3615-3636)
        const int k = (int)(rnd64() % N);
        const double bad = (rnd64() & 1) ? std::numeric_limits<double>::quiet_NaN() : std::numeric_limits<double>::infinity();
        if (rnd64() & 1) { a.v_prev[k] = bad; b.v_prev[k] = bad; } else { a.i_nl_prev[k % M] = bad; b.i_nl_prev[k % M] = bad; }
    }
}

static double rnd_input(long t) {
    const double u = rnd01();
    if (u < 0.002) return std::numeric_limits<double>::quiet_NaN();
    if (u < 0.004) return (t & 1) ? std::numeric_limits<double>::infinity() : -std::numeric_limits<double>::infinity();
    if (u < 0.05) return rnd_log(1.0, 500.0) * rnd_sym();
    return rnd_log(1e-6, 1.0) * rnd_sym();
}

static void preamp_random(long trials, ow::pre::Diag& dg, gen_preamp::CircuitState& counters) {
    const double rates[5] = {48000.0, 88200.0, 96000.0, 44100.0, 192000.0};
    for (long t = 0; t < trials; t++) {
        ow::pre::CircuitState a;
        a.set_default();
        a.diag = &dg;
        gen_preamp::CircuitState b = gen_preamp::CircuitState::default_();
        const double sr = rnd01() < 0.1 ? rnd_log(8000.0, 400000.0) : rates[rnd64() % 5];
        a.set_sample_rate(sr); b.set_sample_rate(sr);
        if (rnd01() < 0.9) { const double r = rnd01() < 0.05 ? rnd_log(10.0, 1e8) : rnd_log(1e3, 1e6); a.set_runtime_R_r_ldr(r); b.set_runtime_R_r_ldr(r); }
        poke_pre(a, b);
        for (int step = 0; step < 4; step++) {
            if (step && rnd01() < 0.5) { const double r = rnd_log(1e3, 1e6); a.set_runtime_R_r_ldr(r); b.set_runtime_R_r_ldr(r); }
            const double x = rnd_input(t);
            const double ya = ow::pre::process_sample(x, a);
            const double yb = gen_preamp::process_sample(x, b)[0];
            expect(same(ya, yb), "pre output", t);
            compare_pre(a, b, t);
        }
        counters.diag_be_fallback_count += b.diag_be_fallback_count;
        counters.diag_voltage_damp_count += b.diag_voltage_damp_count;
        counters.diag_nan_reset_count += b.diag_nan_reset_count;
        counters.diag_nr_max_iter_count += b.diag_nr_max_iter_count;
    }
}

static void preamp_trajectory(long n, double sr, ow::pre::Diag& dg, gen_preamp::CircuitState& counters) {
    ow::pre::CircuitState a;
    a.set_default();
    a.diag = &dg;
    gen_preamp::CircuitState b = gen_preamp::CircuitState::default_();
    a.set_sample_rate(sr); b.set_sample_rate(sr);
    double phase = 0.0, lfo = 0.0;
    for (long t = 0; t < n; t++) {
        const double prog = (double)t / (double)n;
        const double amp = 1e-3 * std::pow(60000.0, prog);  // 1 mV .. 60 V
        const double f = 220.0 * (1.0 + 7.0 * prog);
        phase += 2.0 * 3.14159265358979323846 * f / sr;
        lfo += 2.0 * 3.14159265358979323846 * 5.6 / sr;
        double x = amp * std::sin(phase);
        if (t % 50021 == 50000) x = std::numeric_limits<double>::quiet_NaN();
        if (t % 70001 == 70000) x = std::numeric_limits<double>::infinity();
        if (t % 3 != 2 || prog > 0.5) {  // LDR modulation (every sample in the second half, as under tremolo)
            const double r = std::exp(std::log(2.0e4) + std::log(40.0) * 0.5 * (1.0 + std::sin(lfo)));
            a.set_runtime_R_r_ldr(r); b.set_runtime_R_r_ldr(r);
        }
        const double ya = ow::pre::process_sample(x, a);
        const double yb = gen_preamp::process_sample(x, b)[0];
        expect(same(ya, yb), "pre trajectory output", t);
        if ((t & 1023) == 0) compare_pre(a, b, t);
    }
    compare_pre(a, b, n);
    counters.diag_be_fallback_count += b.diag_be_fallback_count;
    counters.diag_voltage_damp_count += b.diag_voltage_damp_count;
    counters.diag_nan_reset_count += b.diag_nan_reset_count;
    counters.diag_nr_max_iter_count += b.diag_nr_max_iter_count;
}

// ---- tremolo oscillator ---------------------------------------------------------------------------------------------------------------
static void compare_trm(const ow::trm::CircuitState& a, const gen_tremolo::CircuitState& b, long trial) {
    using namespace ow::trm;
    for (int i = 0; i < N; i++) expect(same(a.v_prev[i], b.v_prev[i]), "trm v_prev", trial);
    for (int i = 0; i < M; i++) { expect(same(a.i_nl_prev[i], b.i_nl_prev[i]), "trm i_nl_prev", trial); expect(same(a.i_nl_prev_prev[i], b.i_nl_prev_prev[i]), "trm i_nl_prev_prev", trial); }
    expect(same(a.input_prev, b.input_prev), "trm input_prev", trial);
    expect(a.last_nr_iterations == b.last_nr_iterations, "trm last_nr_iterations", trial);
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { expect(same(a.s[i][j], b.s[i][j]), "trm S", trial); expect(same(a.a_neg[i][j], b.a_neg[i][j]), "trm A_neg", trial); }
    for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) expect(same(a.k[i][j], b.k[i][j]), "trm K", trial);
    for (int i = 0; i < N; i++) for (int j = 0; j < M; j++) expect(same(a.s_ni[i][j], b.s_ni[i][j]), "trm S_NI", trial);
}

static void tremolo_trajectory(long n, double sr, gen_tremolo::CircuitState& counters) {
    ow::trm::CircuitState a;
    a.set_default();  // CircuitState::default() incl. its 50 warm-up samples
    gen_tremolo::CircuitState b = gen_tremolo::CircuitState::default_();
    compare_trm(a, b, -1);
    if (std::fabs(sr - 48000.0) > 0.5) { a.set_sample_rate(sr); b.set_sample_rate(sr); }
    compare_trm(a, b, -2);
    for (long t = 0; t < n; t++) {
        const double ya = ow::trm::process_sample(0.0, a);
        const double yb = gen_tremolo::process_sample(0.0, b)[0];
        expect(same(ya, yb), "trm trajectory output", t);
        if ((t & 4095) == 0) compare_trm(a, b, t);
    }
    compare_trm(a, b, n);
    counters.diag_be_fallback_count += b.diag_be_fallback_count;
    counters.diag_nan_reset_count += b.diag_nan_reset_count;
    counters.diag_nr_max_iter_count += b.diag_nr_max_iter_count;
}

static void tremolo_random(long trials, gen_tremolo::CircuitState& counters) {
    const double rates[3] = {48000.0, 88200.0, 96000.0};
    ow::trm::CircuitState a0;
    a0.set_default();
    const gen_tremolo::CircuitState b0 = gen_tremolo::CircuitState::default_();
    for (long t = 0; t < trials; t++) {
        ow::trm::CircuitState a = a0;
        gen_tremolo::CircuitState b = b0;
        if ((t & 7) == 0) { const double sr = rnd01() < 0.3 ? rnd_log(20000.0, 200000.0) : rates[rnd64() % 3]; a.set_sample_rate(sr); b.set_sample_rate(sr); }
        const double scale = rnd01() < 0.15 ? rnd_log(1.0, 40.0) : rnd_log(1e-6, 1.0);
        for (int i = 0; i < ow::trm::N; i++) { const double v = a.v_prev[i] + scale * rnd_sym(); a.v_prev[i] = v; b.v_prev[i] = v; }
        for (int i = 0; i < ow::trm::M; i++) {
            const double c = a.i_nl_prev[i] * (1.0 + 0.2 * rnd_sym()), d = c * (1.0 + 0.01 * rnd_sym());
            a.i_nl_prev[i] = c; b.i_nl_prev[i] = c; a.i_nl_prev_prev[i] = d; b.i_nl_prev_prev[i] = d;
        }
        for (int step = 0; step < 4; step++) {
            const double x = rnd01() < 0.02 ? rnd_input(t) : 0.0;
            const double ya = ow::trm::process_sample(x, a);
            const double yb = gen_tremolo::process_sample(x, b)[0];
            expect(same(ya, yb), "trm output", t);
            compare_trm(a, b, t);
        }
        counters.diag_be_fallback_count += b.diag_be_fallback_count - b0.diag_be_fallback_count;
        counters.diag_nan_reset_count += b.diag_nan_reset_count - b0.diag_nan_reset_count;
        counters.diag_nr_max_iter_count += b.diag_nr_max_iter_count - b0.diag_nr_max_iter_count;
    }
}

int main(int argc, char** argv) {
    const long trials = argc > 1 ? std::atol(argv[1]) : 100000;
    const long traj = argc > 2 ? std::atol(argv[2]) : 300000;
    ow::pre::Diag dg;
    gen_preamp::CircuitState pc = gen_preamp::CircuitState::default_();
    preamp_random(trials, dg, pc);
    const long bad_pre_random = g_bad;
    for (double sr : {48000.0, 88200.0, 96000.0}) preamp_trajectory(traj, sr, dg, pc);
    const long bad_pre = g_bad;
    gen_tremolo::CircuitState tc = gen_tremolo::CircuitState::default_();
    tc.diag_be_fallback_count = tc.diag_nan_reset_count = tc.diag_nr_max_iter_count = 0;
    tremolo_random(trials / 4, tc);
    for (double sr : {48000.0, 88200.0, 96000.0}) tremolo_trajectory(traj, sr, tc);
    std::printf("{\"trials\": %ld, \"trajectory_samples\": %ld, \"preamp_random_mismatches\": %ld, \"preamp_mismatches\": %ld, \"tremolo_mismatches\": %ld, "
                "\"preamp_be_fallback\": %llu, \"preamp_voltage_damp\": %llu, \"preamp_nan_reset\": %llu, \"preamp_nr_max_iter\": %llu, "
                "\"preamp_pnjlim_ln\": %llu, \"preamp_rebuilds\": %llu, "
                "\"tremolo_be_fallback\": %llu, \"tremolo_nan_reset\": %llu, \"tremolo_nr_max_iter\": %llu}\n",
                trials, traj, bad_pre_random, bad_pre, g_bad - bad_pre, (unsigned long long)pc.diag_be_fallback_count, (unsigned long long)pc.diag_voltage_damp_count,
                (unsigned long long)pc.diag_nan_reset_count, (unsigned long long)pc.diag_nr_max_iter_count, (unsigned long long)dg.pnjlim_ln,
                (unsigned long long)dg.rebuilds, (unsigned long long)tc.diag_be_fallback_count, (unsigned long long)tc.diag_nan_reset_count,
                (unsigned long long)tc.diag_nr_max_iter_count);
    return g_bad ? 1 : 0;
}
