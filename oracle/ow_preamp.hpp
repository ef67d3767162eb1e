// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity checker), never on the product path.
//
// Restatement of the reference's melange-generated 12-node DK preamp solver
// (crates/openwurli-dsp/src/gen_preamp.rs) and its adapter
// (dk_preamp/melange_adapter.rs), thermal noise omitted (default OFF; the
// reference seeds it from the system clock, gen_preamp.rs:1511-1520).
// Same operation order; no FMA except the explicit mul_add at gen_preamp.rs:3592.
#pragma once
#include "ow_voice.hpp"

namespace ow {
namespace pre {

constexpr int N = 12, M = 3;
constexpr double SAMPLE_RATE = 48000.0;  // gen_preamp.rs:38

struct Diag {
    uint64_t nr_iter_hist[16] = {0};  // bucket min(iter,15) of last_nr_iterations per trap solve
    uint64_t nr_max_iter = 0, be_fallback = 0, voltage_damp = 0, nan_reset = 0, singular = 0, rebuilds = 0;
    uint64_t pnjlim_ln = 0;
};

static inline double fast_exp(double x) {  // gen_preamp.rs:2277-2302
    x = rclamp(x, -40.0, 40.0);
    const double LN2_INV = 1.4426950408889634;  // std::f64::consts::LOG2_E
    const double LN2_HI = 0.6931471803691238;
    const double LN2_LO = 1.9082149292705877e-10;
    const double SHIFT = 6755399441055744.0;
    const double z = x * LN2_INV + SHIFT;
    int64_t zb, sb;
    std::memcpy(&zb, &z, 8);
    std::memcpy(&sb, &SHIFT, 8);
    const int64_t n_i64 = zb - sb;
    const double n = (double)n_i64;
    const double f = (x - n * LN2_HI) - n * LN2_LO;
    const double p = 1.0 + f * (1.0 + f * (0.5 + f * (0.16666666666666607 + f * (0.04166666666665876 + f * 0.008333333333492337))));
    const uint64_t pb = ((uint64_t)(1023 + n_i64)) << 52;
    double pow2n;
    std::memcpy(&pow2n, &pb, 8);
    return p * pow2n;
}

static inline double pnjlim(double vnew, double vold, double vt, double vcrit, Diag* dg) {  // gen_preamp.rs:2340-2355
    if (vnew > vcrit && std::fabs(vnew - vold) > vt + vt) {
        if (dg) dg->pnjlim_ln++;
        if (vold >= 0.0) {
            const double arg = 1.0 + (vnew - vold) / vt;
            if (arg > 0.0) return vold + vt * std::log(arg);
            return vcrit;
        }
        return vt * std::log(vnew / vt);
    }
    return vnew;
}

static inline double diode_current(double v_d, double is, double n_vt) {  // gen_preamp.rs:2416-2419
    const double vc = rclamp(v_d, -40.0 * n_vt, 40.0 * n_vt);
    return is * (fast_exp(vc / n_vt) - 1.0);
}
static inline double diode_conductance(double v_d, double is, double n_vt) {  // gen_preamp.rs:2423-2426
    const double vc = rclamp(v_d, -40.0 * n_vt, 40.0 * n_vt);
    return (is / n_vt) * fast_exp(vc / n_vt);
}

// invert_n, gen_preamp.rs:2117-2219. Returns singular flag.
static inline bool invert12(const double a[N][N], double result[N][N]) {
    double lu[N][N];
    int perm[N];
    std::memcpy(lu, a, sizeof(lu));
    for (int i = 0; i < N; i++) perm[i] = i;
    auto identity = [&]() {
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) result[i][j] = (i == j) ? 1.0 : 0.0;
    };
    for (int k = 0; k < N; k++) {
        int max_row = k;
        double max_val = std::fabs(lu[k][k]);
        for (int i = k + 1; i < N; i++) {
            const double v = std::fabs(lu[i][k]);
            if (v > max_val) { max_val = v; max_row = i; }
        }
        if (max_val < 1e-30) { identity(); return true; }
        if (max_row != k) {
            for (int j = 0; j < N; j++) std::swap(lu[k][j], lu[max_row][j]);
            std::swap(perm[k], perm[max_row]);
        }
        const double pivot = lu[k][k];
        for (int i = k + 1; i < N; i++) {
            const double m = lu[i][k] / pivot;
            lu[i][k] = m;
            for (int j = k + 1; j < N; j++) lu[i][j] -= m * lu[k][j];
        }
    }
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) result[i][j] = 0.0;
    for (int col = 0; col < N; col++) {
        double b[N] = {0};
        int start = N;
        for (int i = 0; i < N; i++) {
            if (perm[i] == col) { b[i] = 1.0; start = i; break; }
        }
        for (int i = start + 1; i < N; i++) {
            double sum = b[i];
            for (int j = start; j < i; j++) sum -= lu[i][j] * b[j];
            b[i] = sum;
        }
        for (int i = N - 1; i >= 0; i--) {
            double sum = b[i];
            for (int j = i + 1; j < N; j++) sum -= lu[i][j] * b[j];
            const double pivot = lu[i][i];
            if (std::fabs(pivot) < 1e-30) { identity(); return true; }
            b[i] = sum / pivot;
        }
        for (int i = 0; i < N; i++) result[i][col] = b[i];
    }
    return false;
}

struct CircuitState {  // gen_preamp.rs:1596-1746 (noise fields omitted)
    double v_prev[N], i_nl_prev[M], i_nl_prev_prev[M], dc_operating_point[N];
    double input_prev;
    uint32_t last_nr_iterations;
    double s[N][N], a_neg[N][N], k[M][M], s_ni[N][M];
    double s_be[N][N], k_be[M][M], s_ni_be[N][M], a_neg_be[N][N];
    uint32_t be_cooldown;
    double pot_0_resistance;
    double current_sample_rate;
    bool matrices_dirty;
    Diag* diag = nullptr;

    void set_default() {  // gen_preamp.rs:1748-1821
        std::memcpy(v_prev, PRE_DC_OP, sizeof(v_prev));
        std::memcpy(i_nl_prev, PRE_DC_NL_I, sizeof(i_nl_prev));
        std::memcpy(i_nl_prev_prev, PRE_DC_NL_I, sizeof(i_nl_prev_prev));
        std::memcpy(dc_operating_point, PRE_DC_OP, sizeof(dc_operating_point));
        input_prev = 0.0;
        last_nr_iterations = 0;
        std::memcpy(s, PRE_S_DEFAULT, sizeof(s));
        std::memcpy(a_neg, PRE_A_NEG_DEFAULT, sizeof(a_neg));
        std::memcpy(k, PRE_K_DEFAULT, sizeof(k));
        std::memcpy(s_ni, PRE_S_NI_DEFAULT, sizeof(s_ni));
        std::memcpy(s_be, PRE_S_BE_DEFAULT, sizeof(s_be));
        std::memcpy(k_be, PRE_K_BE_DEFAULT, sizeof(k_be));
        std::memcpy(s_ni_be, PRE_S_NI_BE_DEFAULT, sizeof(s_ni_be));
        std::memcpy(a_neg_be, PRE_A_NEG_BE_DEFAULT, sizeof(a_neg_be));
        be_cooldown = 0;
        pot_0_resistance = 9.99999999999999854e4;
        current_sample_rate = SAMPLE_RATE;
        matrices_dirty = false;
    }

    void rebuild_matrices() {  // gen_preamp.rs:1990-2063
        if (diag) diag->rebuilds++;
        const double internal_rate = current_sample_rate * 1.0;
        const double alpha = 2.0 * internal_rate;
        double g_eff[N][N];
        std::memcpy(g_eff, PRE_G, sizeof(g_eff));
        {
            const double delta_g = 1.0 / pot_0_resistance - PRE_POT_0_G_NOM;
            g_eff[6][6] += delta_g;
        }
        double a[N][N], an[N][N];
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) a[i][j] = g_eff[i][j] + alpha * PRE_C[i][j];
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) an[i][j] = alpha * PRE_C[i][j] - g_eff[i][j];
        for (int j = 0; j < N; j++) an[11][j] = 0.0;
        double sn[N][N];
        const bool singular = invert12(a, sn);
        if (singular && diag) diag->singular++;
        double sni[N][M];
        for (int i = 0; i < N; i++)
            for (int j = 0; j < M; j++) {
                double sum = 0.0;
                for (int kk = 0; kk < N; kk++) sum += sn[i][kk] * PRE_N_I[j][kk];
                sni[i][j] = sum;
            }
        double kn[M][M];
        for (int i = 0; i < M; i++)
            for (int j = 0; j < M; j++) {
                double sum = 0.0;
                for (int n_idx = 0; n_idx < N; n_idx++) sum += PRE_N_V[i][n_idx] * sni[n_idx][j];
                kn[i][j] = sum;
            }
        std::memcpy(s, sn, sizeof(s));
        std::memcpy(a_neg, an, sizeof(a_neg));
        std::memcpy(k, kn, sizeof(k));
        std::memcpy(s_ni, sni, sizeof(s_ni));
    }

    void set_sample_rate(double sr) {  // gen_preamp.rs:1930-1958
        if (!(sr > 0.0 && std::isfinite(sr))) return;
        current_sample_rate = sr;
        if (std::fabs(sr - SAMPLE_RATE) < 0.5) {
            std::memcpy(s, PRE_S_DEFAULT, sizeof(s));
            std::memcpy(a_neg, PRE_A_NEG_DEFAULT, sizeof(a_neg));
            std::memcpy(k, PRE_K_DEFAULT, sizeof(k));
            std::memcpy(s_ni, PRE_S_NI_DEFAULT, sizeof(s_ni));
            std::memcpy(s_be, PRE_S_BE_DEFAULT, sizeof(s_be));
            std::memcpy(k_be, PRE_K_BE_DEFAULT, sizeof(k_be));
            std::memcpy(s_ni_be, PRE_S_NI_BE_DEFAULT, sizeof(s_ni_be));
            std::memcpy(a_neg_be, PRE_A_NEG_BE_DEFAULT, sizeof(a_neg_be));
            return;
        }
        rebuild_matrices();
    }

    void set_runtime_R_r_ldr(double resistance) {  // gen_preamp.rs:1973-1984
        if (!std::isfinite(resistance)) return;
        const double r = rclamp(resistance, 1.0e3, 1.0e6);
        if (std::fabs(r - pot_0_resistance) < 1e-12) return;
        pot_0_resistance = r;
        matrices_dirty = true;
    }
};

// solve_nonlinear, gen_preamp.rs:3122-3357. `kk` is the kernel in effect (k or, during BE, k_be).
static inline void solve_nonlinear(const double p[M], CircuitState& st, const double kk[M][M], double i_nl[M]) {
    const int MAX_ITER = 265;
    const double SING = 1e-15;
    const double d0_is = PRE_DEVICE_0_IS, d0_nvt = PRE_DEVICE_0_N_VT;
    const double d1_is = PRE_DEVICE_1_IS, d1_vt = PRE_DEVICE_1_VT;
    const double d2_is = PRE_DEVICE_2_IS, d2_vt = PRE_DEVICE_2_VT;
    for (int i = 0; i < M; i++) i_nl[i] = 2.0 * st.i_nl_prev[i] - st.i_nl_prev_prev[i];
    for (int iter = 0; iter < MAX_ITER; iter++) {
        const double v_d0 = p[0] + kk[0][0] * i_nl[0] + kk[0][1] * i_nl[1] + kk[0][2] * i_nl[2];
        const double v_d1 = p[1] + kk[1][0] * i_nl[0] + kk[1][1] * i_nl[1] + kk[1][2] * i_nl[2];
        const double v_d2 = p[2] + kk[2][0] * i_nl[0] + kk[2][1] * i_nl[1] + kk[2][2] * i_nl[2];
        const double i_dev0 = diode_current(v_d0, d0_is, d0_nvt);
        const double jdev_0_0 = diode_conductance(v_d0, d0_is, d0_nvt);
        const double vbe_1 = v_d1 * 1.0;
        const double exp_be_1 = fast_exp(vbe_1 / (PRE_DEVICE_1_NF * d1_vt));
        const double i_dev1 = d1_is * (exp_be_1 - 1.0) * 1.0;
        const double jdev_1_1 = d1_is / (PRE_DEVICE_1_NF * d1_vt) * exp_be_1;
        const double vbe_2 = v_d2 * 1.0;
        const double exp_be_2 = fast_exp(vbe_2 / (PRE_DEVICE_2_NF * d2_vt));
        const double i_dev2 = d2_is * (exp_be_2 - 1.0) * 1.0;
        const double jdev_2_2 = d2_is / (PRE_DEVICE_2_NF * d2_vt) * exp_be_2;
        const double f0 = i_nl[0] - i_dev0, f1 = i_nl[1] - i_dev1, f2 = i_nl[2] - i_dev2;
        double a[3][3] = {
            {1.0 - jdev_0_0 * kk[0][0], 0.0 - jdev_0_0 * kk[0][1], 0.0 - jdev_0_0 * kk[0][2]},
            {0.0 - jdev_1_1 * kk[1][0], 1.0 - jdev_1_1 * kk[1][1], 0.0 - jdev_1_1 * kk[1][2]},
            {0.0 - jdev_2_2 * kk[2][0], 0.0 - jdev_2_2 * kk[2][1], 1.0 - jdev_2_2 * kk[2][2]}};
        double b[3] = {f0, f1, f2};
        bool singular = false;
        for (int col = 0; col < 3; col++) {
            int max_row = col;
            double max_val = std::fabs(a[col][col]);
            for (int row = col + 1; row < 3; row++) {
                if (std::fabs(a[row][col]) > max_val) { max_val = std::fabs(a[row][col]); max_row = row; }
            }
            if (max_val < SING) { singular = true; break; }
            if (max_row != col) {
                for (int j = 0; j < 3; j++) std::swap(a[col][j], a[max_row][j]);
                std::swap(b[col], b[max_row]);
            }
            const double pivot = a[col][col];
            for (int row = col + 1; row < 3; row++) {
                const double factor = a[row][col] / pivot;
                for (int j = col + 1; j < 3; j++) a[row][j] -= factor * a[col][j];
                b[row] -= factor * b[col];
            }
        }
        if (!singular) {
            for (int i = 2; i >= 0; i--) {
                double sum = b[i];
                for (int j = i + 1; j < 3; j++) sum -= a[i][j] * b[j];
                if (std::fabs(a[i][i]) < SING) { singular = true; break; }
                b[i] = sum / a[i][i];
            }
        }
        if (!singular) {
            const double delta0 = b[0], delta1 = b[1], delta2 = b[2];
            const double dv0 = -(kk[0][0] * delta0 + kk[0][1] * delta1 + kk[0][2] * delta2);
            const double dv1 = -(kk[1][0] * delta0 + kk[1][1] * delta1 + kk[1][2] * delta2);
            const double dv2 = -(kk[2][0] * delta0 + kk[2][1] * delta1 + kk[2][2] * delta2);
            double alpha[3] = {1.0, 1.0, 1.0};
            bool any_limited = false;
            if (std::fabs(dv0) > 1e-4) {
                const double v_lim = pnjlim(v_d0 + dv0, v_d0, d0_nvt, PRE_DEVICE_0_VCRIT, st.diag);
                const double ratio = rmax((v_lim - v_d0) / dv0, 0.01);
                if (ratio < alpha[0]) { alpha[0] = ratio; if (ratio < 1.0) any_limited = true; }
            }
            if (std::fabs(dv1) > 1e-4) {
                const double v_lim = pnjlim(v_d1 + dv1, v_d1, d1_vt, PRE_DEVICE_1_VCRIT, st.diag);
                const double ratio = rmax((v_lim - v_d1) / dv1, 0.01);
                if (ratio < alpha[1]) { alpha[1] = ratio; if (ratio < 1.0) any_limited = true; }
            }
            if (std::fabs(dv2) > 1e-4) {
                const double v_lim = pnjlim(v_d2 + dv2, v_d2, d2_vt, PRE_DEVICE_2_VCRIT, st.diag);
                const double ratio = rmax((v_lim - v_d2) / dv2, 0.01);
                if (ratio < alpha[2]) { alpha[2] = ratio; if (ratio < 1.0) any_limited = true; }
            }
            double alpha_scalar = rmin(alpha[0], rmin(alpha[1], alpha[2]));
            if (alpha_scalar < 1.0) any_limited = true;
            const double max_di = rmax(rmax(std::fabs(delta0), std::fabs(delta1)), std::fabs(delta2));
            if (max_di * alpha_scalar > 0.1) alpha_scalar = rmin(rmax(0.1 / max_di, 0.01), alpha_scalar);
            i_nl[0] -= alpha_scalar * delta0;
            i_nl[1] -= alpha_scalar * delta1;
            i_nl[2] -= alpha_scalar * delta2;
            bool conv = true;
            if (!any_limited) {
                { const double step = dv0 * alpha_scalar; const double v_new = v_d0 + step;
                  const double thr = 1e-3 * rmax(std::fabs(v_d0), std::fabs(v_new)) + 1e-6; if (std::fabs(step) > thr) conv = false; }
                { const double step = dv1 * alpha_scalar; const double v_new = v_d1 + step;
                  const double thr = 1e-3 * rmax(std::fabs(v_d1), std::fabs(v_new)) + 1e-6; if (std::fabs(step) > thr) conv = false; }
                { const double step = dv2 * alpha_scalar; const double v_new = v_d2 + step;
                  const double thr = 1e-3 * rmax(std::fabs(v_d2), std::fabs(v_new)) + 1e-6; if (std::fabs(step) > thr) conv = false; }
            }
            { const double thr = 1e-3 * rmax(rmax(std::fabs(i_nl[0]), std::fabs(i_dev0)), 1e-9) + 1e-12; if (std::fabs(f0) > thr) conv = false; }
            { const double thr = 1e-3 * rmax(rmax(std::fabs(i_nl[1]), std::fabs(i_dev1)), 1e-9) + 1e-12; if (std::fabs(f1) > thr) conv = false; }
            { const double thr = 1e-3 * rmax(rmax(std::fabs(i_nl[2]), std::fabs(i_dev2)), 1e-9) + 1e-12; if (std::fabs(f2) > thr) conv = false; }
            if (conv) { st.last_nr_iterations = (uint32_t)iter; return; }
        } else {
            { const double c = rmax(std::fabs(i_nl[0]) * 0.1, 0.01); i_nl[0] -= rclamp(f0 * 0.5, -c, c); }
            { const double c = rmax(std::fabs(i_nl[1]) * 0.1, 0.01); i_nl[1] -= rclamp(f1 * 0.5, -c, c); }
            { const double c = rmax(std::fabs(i_nl[2]) * 0.1, 0.01); i_nl[2] -= rclamp(f2 * 0.5, -c, c); }
        }
    }
    st.last_nr_iterations = (uint32_t)MAX_ITER;
    for (int i = 0; i < M; i++) if (!std::isfinite(i_nl[i])) i_nl[i] = st.i_nl_prev[i];
}

// process_sample, gen_preamp.rs:3399-3663. Returns output[0].
static inline double process_sample(double input, CircuitState& st) {
    input = std::isfinite(input) ? rclamp(input, -100.0, 100.0) : 0.0;
    if (st.matrices_dirty) { st.rebuild_matrices(); st.matrices_dirty = false; }
    for (int i = 0; i < N; i++) st.v_prev[i] = st.v_prev[i] + 1e-25 - 1e-25;
    for (int i = 0; i < M; i++) st.i_nl_prev[i] = st.i_nl_prev[i] + 1e-25 - 1e-25;
    const bool force_be = st.be_cooldown > 0;
    if (st.be_cooldown > 0) st.be_cooldown -= 1;

    // build_rhs, gen_preamp.rs:3041-3095
    double rhs[N];
    std::memcpy(rhs, PRE_RHS_CONST, sizeof(rhs));
    const double (*an)[N] = st.a_neg;
    const double* vp = st.v_prev;
    rhs[0] += an[0][0] * vp[0] + an[0][1] * vp[1];
    rhs[1] += an[1][0] * vp[0] + an[1][1] * vp[1] + an[1][2] * vp[2];
    rhs[2] += an[2][1] * vp[1] + an[2][2] * vp[2] + an[2][3] * vp[3] + an[2][4] * vp[4] + an[2][5] * vp[5];
    rhs[3] += an[3][2] * vp[2] + an[3][3] * vp[3] + an[3][4] * vp[4] + an[3][7] * vp[7] + an[3][11] * vp[11];
    rhs[4] += an[4][2] * vp[2] + an[4][3] * vp[3] + an[4][4] * vp[4] + an[4][7] * vp[7] + an[4][8] * vp[8];
    rhs[5] += an[5][2] * vp[2] + an[5][5] * vp[5] + an[5][6] * vp[6];
    rhs[6] += an[6][5] * vp[5] + an[6][6] * vp[6] + an[6][10] * vp[10];
    rhs[7] += an[7][3] * vp[3] + an[7][4] * vp[4] + an[7][7] * vp[7] + an[7][10] * vp[10];
    rhs[8] += an[8][4] * vp[4] + an[8][8] * vp[8] + an[8][9] * vp[9];
    rhs[9] += an[9][8] * vp[8] + an[9][9] * vp[9];
    rhs[10] += an[10][6] * vp[6] + an[10][7] * vp[7] + an[10][10] * vp[10];
    rhs[2] += PRE_N_I[0][2] * st.i_nl_prev[0];
    rhs[2] += PRE_N_I[1][2] * st.i_nl_prev[1];
    rhs[4] += PRE_N_I[1][4] * st.i_nl_prev[1];
    rhs[4] += PRE_N_I[2][4] * st.i_nl_prev[2];
    rhs[5] += PRE_N_I[1][5] * st.i_nl_prev[1];
    rhs[7] += PRE_N_I[2][7] * st.i_nl_prev[2];
    rhs[8] += PRE_N_I[2][8] * st.i_nl_prev[2];
    rhs[0] += (input + st.input_prev) / 1.0;

    double v_pred[N];
    for (int i = 0; i < N; i++) {  // mat_vec_mul_s, :3099-3109
        double sum = 0.0;
        for (int j = 0; j < N; j++) sum += st.s[i][j] * rhs[j];
        v_pred[i] = sum;
    }
    const double p[M] = {-v_pred[2], v_pred[2] - v_pred[5], v_pred[4] - v_pred[8]};  // :3113-3115
    double i_nl[M];
    solve_nonlinear(p, st, st.k, i_nl);
    if (st.diag) st.diag->nr_iter_hist[std::min<uint32_t>(st.last_nr_iterations, 15)]++;
    double v[N];
    for (int i = 0; i < N; i++) {  // compute_final_voltages, :3367-3375
        double acc = v_pred[i];
        for (int j = 0; j < M; j++) acc += st.s_ni[i][j] * i_nl[j];
        v[i] = acc;
    }

    // Step 6b: BE fallback, :3478-3572
    const bool nr_failed = st.last_nr_iterations >= 265u;
    bool ringing = false;
    for (int i = 0; i < 11; i++) if (std::fabs(v[i]) > 55.0) { ringing = true; break; }
    const bool need_be = nr_failed || ringing || force_be;
    if (need_be) {
        if (nr_failed && st.diag) st.diag->nr_max_iter++;
        if (ringing || nr_failed) st.be_cooldown = 64;
        if (st.diag) st.diag->be_fallback++;
        double rhs_be[N];
        for (int i = 0; i < N; i++) {
            double sum = PRE_RHS_CONST_BE[i];
            for (int j = 0; j < N; j++) sum += st.a_neg_be[i][j] * st.v_prev[j];
            for (int j = 0; j < M; j++) sum += PRE_N_I[j][i] * st.i_nl_prev[j];
            rhs_be[i] = sum;
        }
        rhs_be[0] += input / 1.0;
        double v_pred_be[N];
        for (int i = 0; i < N; i++) {
            double sum = 0.0;
            for (int j = 0; j < N; j++) sum += st.s_be[i][j] * rhs_be[j];
            v_pred_be[i] = sum;
        }
        double p_be[M];
        for (int i = 0; i < M; i++) {
            double sum = 0.0;
            for (int j = 0; j < N; j++) sum += PRE_N_V[i][j] * v_pred_be[j];
            p_be[i] = sum;
        }
        double i_nl_be[M];
        solve_nonlinear(p_be, st, st.k_be, i_nl_be);
        for (int i = 0; i < N; i++) {
            double acc = v_pred_be[i];
            for (int j = 0; j < M; j++) acc += st.s_ni_be[i][j] * i_nl_be[j];
            v[i] = acc;
        }
        for (int j = 0; j < M; j++) i_nl[j] = i_nl_be[j];
    }

    // Step 6c: voltage damping, :3574-3613
    {
        double max_delta = 0.0;
        for (int i = 0; i < 11; i++) {
            const double d = std::fabs(v[i] - st.v_prev[i]);
            if (d > max_delta) max_delta = d;
        }
        double max_dc = 0.0;
        for (int i = 0; i < 11; i++) {
            const double a = std::fabs(st.dc_operating_point[i]);
            if (a > max_dc) max_dc = a;
        }
        const double damp_thresh = std::fma(max_dc, 0.05, 2.0);
        if (max_delta > damp_thresh) {
            if (st.diag) st.diag->voltage_damp++;
            const double damp = rmax(damp_thresh / max_delta, 0.01);
            for (int i = 0; i < N; i++) v[i] = st.v_prev[i] + damp * (v[i] - st.v_prev[i]);
            for (int i = 0; i < M; i++) i_nl[i] = st.i_nl_prev[i] + damp * (i_nl[i] - st.i_nl_prev[i]);
        }
    }

    // Step 7: NaN reset, :3615-3636
    bool finite = true;
    for (int i = 0; i < N; i++) if (!std::isfinite(v[i])) finite = false;
    if (!finite) {
        std::memcpy(st.v_prev, st.dc_operating_point, sizeof(st.v_prev));
        std::memcpy(st.i_nl_prev, PRE_DC_NL_I, sizeof(st.i_nl_prev));
        std::memcpy(st.i_nl_prev_prev, PRE_DC_NL_I, sizeof(st.i_nl_prev_prev));
        st.input_prev = 0.0;
        st.pot_0_resistance = 9.99999999999999854e4;
        st.be_cooldown = 0;
        if (st.diag) st.diag->nan_reset++;
        return rclamp(st.dc_operating_point[10] * 1.0, -10.0, 10.0);
    }
    std::memcpy(st.v_prev, v, sizeof(st.v_prev));
    std::memcpy(st.i_nl_prev_prev, st.i_nl_prev, sizeof(st.i_nl_prev_prev));
    std::memcpy(st.i_nl_prev, i_nl, sizeof(st.i_nl_prev));
    st.input_prev = input;
    if (st.last_nr_iterations >= 265u && st.diag) st.diag->nr_max_iter++;
    const double raw = std::isfinite(v[10]) ? v[10] : 0.0;
    const double scaled = raw * 1.0;
    return std::isfinite(scaled) ? scaled : 0.0;
}

// ---- dk_preamp/melange_adapter.rs ---------------------------------------------------------
static inline const CircuitState& settled_state() {  // melange_adapter.rs:12-20 (OnceLock)
    static CircuitState cached = [] {
        CircuitState s;
        s.set_default();
        for (int i = 0; i < 176400; i++) process_sample(0.0, s);
        return s;
    }();
    return cached;
}
static inline CircuitState init_state(double sample_rate) {  // melange_adapter.rs:22-29
    CircuitState st = settled_state();
    if (std::fabs(sample_rate - SAMPLE_RATE) > 0.5) st.set_sample_rate(sample_rate);
    return st;
}

struct DkPreamp {  // melange_adapter.rs:31-94
    CircuitState main, shadow;
    double sample_rate;
    Diag diag_main, diag_shadow;
    explicit DkPreamp(double sr) : sample_rate(sr) {
        main = init_state(sr);
        shadow = init_state(sr);
        main.diag = &diag_main;
        shadow.diag = &diag_shadow;
    }
    DkPreamp(const DkPreamp&) = delete;
    void reset() {
        main = init_state(sample_rate);
        shadow = init_state(sample_rate);
        main.diag = &diag_main;
        shadow.diag = &diag_shadow;
    }
    void set_ldr_resistance(double r) {
        main.set_runtime_R_r_ldr(r);
        shadow.set_runtime_R_r_ldr(r);
    }
    double process_sample(double input, double* main_out_tap = nullptr, double* pump_tap = nullptr) {
        const double main_out = pre::process_sample(input, main);
        const double pump = pre::process_sample(0.0, shadow);
        if (main_out_tap) *main_out_tap = main_out;
        if (pump_tap) *pump_tap = pump;
        const double result = main_out - pump;
        if (!std::isfinite(result)) { reset(); return 0.0; }
        return result;
    }
};

}  // namespace pre
}  // namespace ow
