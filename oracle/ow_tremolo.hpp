// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity checker), never on the product path.
//
// Restatement of the reference's melange-generated Twin-T tremolo oscillator
// (crates/openwurli-dsp/src/gen_tremolo.rs: N=7, M=4, two Ebers-Moll BJTs,
// Schur-complement NR, BE fallback) and of tremolo.rs (LED drive -> CdS
// envelope -> power-law R_ldr -> vibrato-pot shunt impedance).
#pragma once
#include "ow_preamp.hpp"

namespace ow {
namespace trm {

constexpr int N = 7, M = 4;
constexpr int MAX_ITER = 50;
constexpr double SAMPLE_RATE = 48000.0;

using pre::fast_exp;

static inline double pnjlim(double vnew, double vold, double vt, double vcrit) {  // gen_tremolo.rs:1203-1218
    if (vnew > vcrit && std::fabs(vnew - vold) > vt + vt) {
        if (vold >= 0.0) {
            const double arg = 1.0 + (vnew - vold) / vt;
            if (arg > 0.0) return vold + vt * std::log(arg);
            return vcrit;
        }
        return vt * std::log(vnew / vt);
    }
    return vnew;
}

// bjt_evaluate, gen_tremolo.rs:1546-1636, specialised to what the call sites pass:
// use_gp=false, ISE=ISC=0, sign=+1 (Ebers-Moll branch only).
struct BjtOut { double ic, ib, jac[4]; };
static inline BjtOut bjt_evaluate_em(double vbe, double vbc, double is, double vt, double nf, double nr,
                                     double beta_f, double beta_r) {
    const double sign = 1.0;
    const double vbe_eff = sign * vbe;
    const double vbc_eff = sign * vbc;
    const double nf_vt = nf * vt;
    const double nr_vt = nr * vt;
    const double exp_be = fast_exp(vbe_eff / nf_vt);
    const double exp_bc = fast_exp(vbc_eff / nr_vt);
    const double i_cc = is * (exp_be - exp_bc);
    const double ib_fwd = is / beta_f * (exp_be - 1.0);
    const double ib_rev = is / beta_r * (exp_bc - 1.0);
    const double ib_leak_be = 0.0, ib_leak_bc = 0.0;
    const double dib_fwd_dvbe = (is / (beta_f * nf_vt)) * exp_be;
    const double dib_rev_dvbc = (is / (beta_r * nr_vt)) * exp_bc;
    const double dib_leak_dvbe = 0.0, dib_leak_dvbc = 0.0;
    BjtOut o;
    o.ic = sign * (i_cc - is / beta_r * (exp_bc - 1.0));
    o.ib = sign * (ib_fwd + ib_rev + ib_leak_be + ib_leak_bc);
    o.jac[0] = is / nf_vt * exp_be;
    o.jac[1] = -(is / nr_vt) * exp_bc - (is / (beta_r * nr_vt)) * exp_bc;
    o.jac[2] = dib_fwd_dvbe + dib_leak_dvbe;
    o.jac[3] = dib_rev_dvbc + dib_leak_dvbc;
    return o;
}

// invert_n, gen_tremolo.rs:2273-2342. Returns false if singular (result untouched).
static inline bool invert7(const double a[N][N], double result[N][N]) {
    double lu[N][N];
    int perm[N];
    std::memcpy(lu, a, sizeof(lu));
    for (int i = 0; i < N; i++) perm[i] = i;
    for (int k = 0; k < N; k++) {
        int max_row = k;
        double max_val = std::fabs(lu[k][k]);
        for (int i = k + 1; i < N; i++) {
            const double v = std::fabs(lu[i][k]);
            if (v > max_val) { max_val = v; max_row = i; }
        }
        if (max_val < 1e-30) return false;
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
    double res[N][N] = {{0}};
    for (int col = 0; col < N; col++) {
        double b[N] = {0};
        int start = N;
        for (int i = 0; i < N; i++) if (perm[i] == col) { b[i] = 1.0; start = i; break; }
        for (int i = start + 1; i < N; i++) {
            double sum = b[i];
            for (int j = start; j < i; j++) sum -= lu[i][j] * b[j];
            b[i] = sum;
        }
        for (int i = N - 1; i >= 0; i--) {
            double sum = b[i];
            for (int j = i + 1; j < N; j++) sum -= lu[i][j] * b[j];
            const double pivot = lu[i][i];
            if (std::fabs(pivot) < 1e-30) return false;
            b[i] = sum / pivot;
        }
        for (int i = 0; i < N; i++) res[i][col] = b[i];
    }
    std::memcpy(result, res, sizeof(res));
    return true;
}

struct CircuitState {  // gen_tremolo.rs:1853-1963 (only the fields process_sample reads/writes)
    double v_prev[N], i_nl_prev[M], i_nl_prev_prev[M], dc_operating_point[N];
    double input_prev;
    uint32_t last_nr_iterations;
    uint64_t diag_nr_max_iter = 0, diag_be_fallback = 0, diag_nan_reset = 0;
    uint64_t nr_iter_hist[16] = {0};
    double a_neg[N][N], a_neg_be[N][N];
    double s[N][N], k[M][M], s_ni[N][M];
    double s_be[N][N], k_be[M][M], s_ni_be[N][M];

    void set_default();  // incl. the 50-sample warm-up, gen_tremolo.rs:1965-2024, 2071-2075
    void rebuild_matrices(double internal_rate);
    void set_sample_rate(double sr) {  // gen_tremolo.rs:2111-2133
        if (!(sr > 0.0 && std::isfinite(sr))) return;
        if (std::fabs(sr - SAMPLE_RATE) < 0.5) {
            std::memcpy(a_neg, TRM_A_NEG_DEFAULT, sizeof(a_neg));
            std::memcpy(a_neg_be, TRM_A_NEG_BE_DEFAULT, sizeof(a_neg_be));
            std::memcpy(s, TRM_S_DEFAULT, sizeof(s));
            std::memcpy(s_be, TRM_S_BE_DEFAULT, sizeof(s_be));
            std::memcpy(k, TRM_K_DEFAULT, sizeof(k));
            std::memcpy(s_ni, TRM_S_NI_DEFAULT, sizeof(s_ni));
            std::memcpy(k_be, TRM_K_BE_DEFAULT, sizeof(k_be));
            std::memcpy(s_ni_be, TRM_S_NI_BE_DEFAULT, sizeof(s_ni_be));
            return;
        }
        rebuild_matrices(sr * 1.0);
    }
};

static inline void kernel_products(const double s[N][N], double k[M][M], double s_ni[N][M]) {  // gen_tremolo.rs:2172-2195
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            double sum = 0.0;
            for (int a = 0; a < N; a++) {
                double s_ni_aj = 0.0;
                for (int b = 0; b < N; b++) s_ni_aj += s[a][b] * TRM_N_I[b][j];
                sum += TRM_N_V[i][a] * s_ni_aj;
            }
            k[i][j] = sum;
        }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < M; j++) {
            double sum = 0.0;
            for (int a = 0; a < N; a++) sum += s[i][a] * TRM_N_I[a][j];
            s_ni[i][j] = sum;
        }
}

inline void CircuitState::rebuild_matrices(double internal_rate) {  // gen_tremolo.rs:2139-2258 (s_sub family unused by process_sample)
    const double alpha = 2.0 * internal_rate;
    const double alpha_be = internal_rate;
    double a[N][N], a_be[N][N];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            a[i][j] = TRM_G[i][j] + alpha * TRM_C[i][j];
            a_neg[i][j] = alpha * TRM_C[i][j] - TRM_G[i][j];
            a_be[i][j] = TRM_G[i][j] + alpha_be * TRM_C[i][j];
            a_neg_be[i][j] = alpha_be * TRM_C[i][j];
        }
    for (int j = 0; j < N; j++) { a_neg[6][j] = 0.0; a_neg_be[6][j] = 0.0; }
    if (invert7(a, s)) kernel_products(s, k, s_ni);
    if (invert7(a_be, s_be)) kernel_products(s_be, k_be, s_ni_be);
}

// One NR attempt shared by the trapezoidal and BE passes differs in its limiting
// logic (the reference has two hand-unrolled variants), so both are restated.
static inline void solve4(double a[4][4], double b[4], bool& singular) {  // gen_tremolo.rs:2515-2561
    singular = false;
    for (int col = 0; col < 4; col++) {
        int max_row = col;
        double max_val = std::fabs(a[col][col]);
        for (int row = col + 1; row < 4; row++)
            if (std::fabs(a[row][col]) > max_val) { max_val = std::fabs(a[row][col]); max_row = row; }
        if (max_val < 1e-15) { singular = true; break; }
        if (max_row != col) {
            for (int j = 0; j < 4; j++) std::swap(a[col][j], a[max_row][j]);
            std::swap(b[col], b[max_row]);
        }
        const double pivot = a[col][col];
        for (int row = col + 1; row < 4; row++) {
            const double factor = a[row][col] / pivot;
            for (int j = col + 1; j < 4; j++) a[row][j] -= factor * a[col][j];
            b[row] -= factor * b[col];
        }
    }
    if (!singular) {
        for (int i = 3; i >= 0; i--) {
            double sum = b[i];
            for (int j = i + 1; j < 4; j++) sum -= a[i][j] * b[j];
            if (std::fabs(a[i][i]) < 1e-15) { singular = true; break; }
            b[i] = sum / a[i][i];
        }
    }
}

static inline void jacobian4(const BjtOut& q0, const BjtOut& q1, const double k[M][M], double a[4][4]) {  // :2497-2512
    const double j00 = q0.jac[0], j01 = q0.jac[1], j10 = q0.jac[2], j11 = q0.jac[3];
    const double j22 = q1.jac[0], j23 = q1.jac[1], j32 = q1.jac[2], j33 = q1.jac[3];
    a[0][0] = 1.0 - j00 * k[0][0] - j01 * k[1][0];
    a[0][1] = 0.0 - j00 * k[0][1] - j01 * k[1][1];
    a[0][2] = 0.0 - j00 * k[0][2] - j01 * k[1][2];
    a[0][3] = 0.0 - j00 * k[0][3] - j01 * k[1][3];
    a[1][0] = 0.0 - j10 * k[0][0] - j11 * k[1][0];
    a[1][1] = 1.0 - j10 * k[0][1] - j11 * k[1][1];
    a[1][2] = 0.0 - j10 * k[0][2] - j11 * k[1][2];
    a[1][3] = 0.0 - j10 * k[0][3] - j11 * k[1][3];
    a[2][0] = 0.0 - j22 * k[2][0] - j23 * k[3][0];
    a[2][1] = 0.0 - j22 * k[2][1] - j23 * k[3][1];
    a[2][2] = 1.0 - j22 * k[2][2] - j23 * k[3][2];
    a[2][3] = 0.0 - j22 * k[2][3] - j23 * k[3][3];
    a[3][0] = 0.0 - j32 * k[2][0] - j33 * k[3][0];
    a[3][1] = 0.0 - j32 * k[2][1] - j33 * k[3][1];
    a[3][2] = 0.0 - j32 * k[2][2] - j33 * k[3][2];
    a[3][3] = 1.0 - j32 * k[2][3] - j33 * k[3][3];
}

// process_sample, gen_tremolo.rs:2353-3116. Returns output[0] = v[0].
static inline double process_sample(double input, CircuitState& st) {
    input = std::isfinite(input) ? rclamp(input, -100.0, 100.0) : 0.0;
    for (int i = 0; i < N; i++) st.v_prev[i] = st.v_prev[i] + 1e-25 - 1e-25;
    for (int i = 0; i < M; i++) st.i_nl_prev[i] = st.i_nl_prev[i] + 1e-25 - 1e-25;

    double rhs[N];
    std::memcpy(rhs, TRM_RHS_CONST, sizeof(rhs));
    const double (*an)[N] = st.a_neg;
    const double* vp = st.v_prev;
    rhs[0] += an[0][0] * vp[0];
    rhs[0] += an[0][1] * vp[1];
    rhs[0] += an[0][3] * vp[3];
    rhs[0] += an[0][5] * vp[5];
    rhs[1] += an[1][0] * vp[0];
    rhs[1] += an[1][1] * vp[1];
    rhs[1] += an[1][2] * vp[2];
    rhs[2] += an[2][1] * vp[1];
    rhs[2] += an[2][2] * vp[2];
    rhs[2] += an[2][3] * vp[3];
    rhs[3] += an[3][0] * vp[0];
    rhs[3] += an[3][2] * vp[2];
    rhs[3] += an[3][3] * vp[3];
    rhs[4] += an[4][4] * vp[4];
    rhs[5] += an[5][0] * vp[0];
    rhs[5] += an[5][5] * vp[5];
    rhs[5] += an[5][6] * vp[6];
    rhs[0] += TRM_N_I[0][0] * st.i_nl_prev[0];
    rhs[0] += TRM_N_I[0][2] * st.i_nl_prev[2];
    rhs[2] += TRM_N_I[2][1] * st.i_nl_prev[1];
    rhs[4] += TRM_N_I[4][0] * st.i_nl_prev[0];
    rhs[4] += TRM_N_I[4][1] * st.i_nl_prev[1];
    rhs[4] += TRM_N_I[4][3] * st.i_nl_prev[3];
    const double input_conductance = 1.0 / TRM_INPUT_RESISTANCE;
    rhs[0] += (input + st.input_prev) * input_conductance;
    st.input_prev = input;

    double v_pred[N];
    for (int i = 0; i < N; i++) {
        double sum = 0.0;
        for (int j = 0; j < N; j++) sum += st.s[i][j] * rhs[j];
        v_pred[i] = sum;
    }
    double p[M];
    p[0] = TRM_N_V[0][2] * v_pred[2] + TRM_N_V[0][4] * v_pred[4];
    p[1] = TRM_N_V[1][0] * v_pred[0] + TRM_N_V[1][2] * v_pred[2];
    p[2] = TRM_N_V[2][4] * v_pred[4];
    p[3] = TRM_N_V[3][0] * v_pred[0] + TRM_N_V[3][4] * v_pred[4];

    double i_nl[M];
    for (int i = 0; i < M; i++) i_nl[i] = 2.0 * st.i_nl_prev[i] - st.i_nl_prev_prev[i];
    st.last_nr_iterations = MAX_ITER;
    const double vt0 = TRM_DEVICE_0_VT, vt1 = TRM_DEVICE_1_VT;
    const double (*k)[M] = st.k;

    for (int iter = 0; iter < MAX_ITER; iter++) {
        const double v_d0 = p[0] + k[0][0] * i_nl[0] + k[0][1] * i_nl[1] + k[0][2] * i_nl[2] + k[0][3] * i_nl[3];
        const double v_d1 = p[1] + k[1][0] * i_nl[0] + k[1][1] * i_nl[1] + k[1][2] * i_nl[2];
        const double v_d2 = p[2] + k[2][0] * i_nl[0] + k[2][1] * i_nl[1] + k[2][3] * i_nl[3];
        const double v_d3 = p[3] + k[3][0] * i_nl[0] + k[3][1] * i_nl[1] + k[3][2] * i_nl[2] + k[3][3] * i_nl[3];
        const BjtOut q0 = bjt_evaluate_em(v_d0, v_d1, TRM_DEVICE_0_IS, vt0, TRM_DEVICE_0_NF, TRM_DEVICE_0_NR, TRM_DEVICE_0_BETA_F, TRM_DEVICE_0_BETA_R);
        const BjtOut q1 = bjt_evaluate_em(v_d2, v_d3, TRM_DEVICE_1_IS, vt1, TRM_DEVICE_1_NF, TRM_DEVICE_1_NR, TRM_DEVICE_1_BETA_F, TRM_DEVICE_1_BETA_R);
        const double f0 = i_nl[0] - q0.ic, f1 = i_nl[1] - q0.ib, f2 = i_nl[2] - q1.ic, f3 = i_nl[3] - q1.ib;
        double a[4][4];
        jacobian4(q0, q1, k, a);
        double b[4] = {f0, f1, f2, f3};
        bool singular;
        solve4(a, b, singular);
        if (!singular) {
            const double delta0 = b[0], delta1 = b[1], delta2 = b[2], delta3 = b[3];
            const double it0 = i_nl[0] - delta0, it1 = i_nl[1] - delta1, it2 = i_nl[2] - delta2, it3 = i_nl[3] - delta3;
            const double v_trial0 = p[0] + k[0][0] * it0 + k[0][1] * it1 + k[0][2] * it2 + k[0][3] * it3;
            const double v_trial1 = p[1] + k[1][0] * it0 + k[1][1] * it1 + k[1][2] * it2 + k[1][3] * it3;
            const double v_trial2 = p[2] + k[2][0] * it0 + k[2][1] * it1 + k[2][2] * it2 + k[2][3] * it3;
            const double v_trial3 = p[3] + k[3][0] * it0 + k[3][1] * it1 + k[3][2] * it2 + k[3][3] * it3;
            bool any_limited = false;
            const double dvt0 = v_trial0 - v_d0;
            const double v_lim0 = std::fabs(dvt0) > 1e-4 ? pnjlim(v_trial0, v_d0, vt0, TRM_DEVICE_0_VCRIT) : v_trial0;
            const double dvt1 = v_trial1 - v_d1;
            const double v_lim1 = std::fabs(dvt1) > 1e-4 ? pnjlim(v_trial1, v_d1, vt0, TRM_DEVICE_0_VCRIT) : v_trial1;
            const double dvt2 = v_trial2 - v_d2;
            const double v_lim2 = std::fabs(dvt2) > 1e-4 ? pnjlim(v_trial2, v_d2, vt1, TRM_DEVICE_1_VCRIT) : v_trial2;
            const double dvt3 = v_trial3 - v_d3;
            const double v_lim3 = std::fabs(dvt3) > 1e-4 ? pnjlim(v_trial3, v_d3, vt1, TRM_DEVICE_1_VCRIT) : v_trial3;
            double global_alpha = 1.0;
            auto lim = [&](double v_lim, double v_d, double dv_trial) {
                const double dv_lim = v_lim - v_d;
                if (std::fabs(dv_trial) > 1e-15) {
                    const double r = (dv_trial * dv_lim < 0.0) ? 0.0 : rclamp(dv_lim / dv_trial, 0.0, 1.0);
                    if (r < global_alpha) { global_alpha = r; any_limited = true; }
                }
            };
            lim(v_lim0, v_d0, dvt0);
            lim(v_lim1, v_d1, dvt1);
            lim(v_lim2, v_d2, dvt2);
            lim(v_lim3, v_d3, dvt3);
            {
                const double max_dv = rmax(rmax(rmax(std::fabs(dvt0 * global_alpha), std::fabs(dvt1 * global_alpha)),
                                               std::fabs(dvt2 * global_alpha)), std::fabs(dvt3 * global_alpha));
                if (max_dv > 3.5) { global_alpha *= rmax(3.5 / max_dv, 0.1); any_limited = true; }
            }
            i_nl[0] -= global_alpha * delta0;
            i_nl[1] -= global_alpha * delta1;
            i_nl[2] -= global_alpha * delta2;
            i_nl[3] -= global_alpha * delta3;
            if (!any_limited) {
                bool conv = true;
                auto chk = [&](double dv_trial, double v_d) {
                    const double dv = dv_trial * global_alpha;
                    const double thr = 1e-3 * rmax(std::fabs(v_d), std::fabs(v_d + dv)) + 1e-6;
                    if (std::fabs(dv) > thr) conv = false;
                };
                chk(dvt0, v_d0); chk(dvt1, v_d1); chk(dvt2, v_d2); chk(dvt3, v_d3);
                if (conv) { st.last_nr_iterations = (uint32_t)iter; break; }
            }
        } else {
            { const double c = rmax(std::fabs(i_nl[0]) * 0.1, 0.01); i_nl[0] -= rclamp(f0 * 0.5, -c, c); }
            { const double c = rmax(std::fabs(i_nl[1]) * 0.1, 0.01); i_nl[1] -= rclamp(f1 * 0.5, -c, c); }
            { const double c = rmax(std::fabs(i_nl[2]) * 0.1, 0.01); i_nl[2] -= rclamp(f2 * 0.5, -c, c); }
            { const double c = rmax(std::fabs(i_nl[3]) * 0.1, 0.01); i_nl[3] -= rclamp(f3 * 0.5, -c, c); }
        }
    }
    st.nr_iter_hist[std::min<uint32_t>(st.last_nr_iterations, 15)]++;

    double v[N];
    for (int i = 0; i < N; i++) {
        double acc = v_pred[i];
        for (int j = 0; j < M; j++) acc += st.s_ni[i][j] * i_nl[j];
        v[i] = acc;
    }
    const bool converged = st.last_nr_iterations < (uint32_t)MAX_ITER;

    if (!converged) {  // BE fallback, gen_tremolo.rs:2757-3083
        st.diag_nr_max_iter++;
        st.diag_be_fallback++;
        double rhs_be[N];
        for (int i = 0; i < N; i++) {
            double sum = TRM_RHS_CONST_BE[i];
            for (int j = 0; j < N; j++) sum += st.a_neg_be[i][j] * st.v_prev[j];
            for (int j = 0; j < M; j++) sum += TRM_N_I[i][j] * st.i_nl_prev[j];
            rhs_be[i] = sum;
        }
        rhs_be[0] += input * input_conductance;
        double v_pred_be[N];
        for (int i = 0; i < N; i++) {
            double sum = 0.0;
            for (int j = 0; j < N; j++) sum += st.s_be[i][j] * rhs_be[j];
            v_pred_be[i] = sum;
        }
        double p_be[M];
        for (int i = 0; i < M; i++) {
            double sum = 0.0;
            for (int j = 0; j < N; j++) sum += TRM_N_V[i][j] * v_pred_be[j];
            p_be[i] = sum;
        }
        for (int i = 0; i < M; i++) i_nl[i] = 2.0 * st.i_nl_prev[i] - st.i_nl_prev_prev[i];
        const double (*kb)[M] = st.k_be;
        for (int iter = 0; iter < MAX_ITER; iter++) {
            const double v_d0 = p_be[0] + kb[0][0] * i_nl[0] + kb[0][1] * i_nl[1] + kb[0][2] * i_nl[2] + kb[0][3] * i_nl[3];
            const double v_d1 = p_be[1] + kb[1][0] * i_nl[0] + kb[1][1] * i_nl[1] + kb[1][2] * i_nl[2] + kb[1][3] * i_nl[3];
            const double v_d2 = p_be[2] + kb[2][0] * i_nl[0] + kb[2][1] * i_nl[1] + kb[2][2] * i_nl[2] + kb[2][3] * i_nl[3];
            const double v_d3 = p_be[3] + kb[3][0] * i_nl[0] + kb[3][1] * i_nl[1] + kb[3][2] * i_nl[2] + kb[3][3] * i_nl[3];
            const BjtOut q0 = bjt_evaluate_em(v_d0, v_d1, TRM_DEVICE_0_IS, vt0, TRM_DEVICE_0_NF, TRM_DEVICE_0_NR, TRM_DEVICE_0_BETA_F, TRM_DEVICE_0_BETA_R);
            const BjtOut q1 = bjt_evaluate_em(v_d2, v_d3, TRM_DEVICE_1_IS, vt1, TRM_DEVICE_1_NF, TRM_DEVICE_1_NR, TRM_DEVICE_1_BETA_F, TRM_DEVICE_1_BETA_R);
            const double f0 = i_nl[0] - q0.ic, f1 = i_nl[1] - q0.ib, f2 = i_nl[2] - q1.ic, f3 = i_nl[3] - q1.ib;
            double a[4][4];
            jacobian4(q0, q1, kb, a);
            double b[4] = {f0, f1, f2, f3};
            bool singular;
            solve4(a, b, singular);
            if (!singular) {
                const double delta0 = b[0], delta1 = b[1], delta2 = b[2], delta3 = b[3];
                const double dv0 = -(kb[0][0] * delta0 + kb[0][1] * delta1 + kb[0][2] * delta2 + kb[0][3] * delta3);
                const double dv1 = -(kb[1][0] * delta0 + kb[1][1] * delta1 + kb[1][2] * delta2 + kb[1][3] * delta3);
                const double dv2 = -(kb[2][0] * delta0 + kb[2][1] * delta1 + kb[2][2] * delta2 + kb[2][3] * delta3);
                const double dv3 = -(kb[3][0] * delta0 + kb[3][1] * delta1 + kb[3][2] * delta2 + kb[3][3] * delta3);
                double alpha[4] = {1.0, 1.0, 1.0, 1.0};
                bool any_limited = false;
                auto lim = [&](int idx, double dv, double v_d, double vt, double vcrit) {
                    if (std::fabs(dv) > 1e-4) {
                        const double v_lim = pnjlim(v_d + dv, v_d, vt, vcrit);
                        const double ratio = rmax((v_lim - v_d) / dv, 0.01);
                        if (ratio < alpha[idx]) { alpha[idx] = ratio; if (ratio < 1.0) any_limited = true; }
                    }
                };
                lim(0, dv0, v_d0, vt0, TRM_DEVICE_0_VCRIT);
                lim(1, dv1, v_d1, vt0, TRM_DEVICE_0_VCRIT);
                lim(2, dv2, v_d2, vt1, TRM_DEVICE_1_VCRIT);
                lim(3, dv3, v_d3, vt1, TRM_DEVICE_1_VCRIT);
                { const double d = rmin(alpha[0], alpha[1]); alpha[0] = d; alpha[1] = d; }
                { const double d = rmin(alpha[2], alpha[3]); alpha[2] = d; alpha[3] = d; }
                const double max_dv = rmax(rmax(rmax(std::fabs(dv0 * alpha[0]), std::fabs(dv1 * alpha[1])),
                                               std::fabs(dv2 * alpha[2])), std::fabs(dv3 * alpha[3]));
                if (max_dv > 3.5) {
                    const double factor = rmax(3.5 / max_dv, 0.1);
                    for (int i = 0; i < 4; i++) alpha[i] *= factor;
                }
                i_nl[0] -= alpha[0] * delta0;
                i_nl[1] -= alpha[1] * delta1;
                i_nl[2] -= alpha[2] * delta2;
                i_nl[3] -= alpha[3] * delta3;
                if (!any_limited) {
                    bool conv = true;
                    auto chk = [&](double dv, double al, double v_d) {
                        const double step = dv * al;
                        const double v_new = v_d + step;
                        const double thr = 1e-3 * rmax(std::fabs(v_d), std::fabs(v_new)) + 1e-6;
                        if (std::fabs(step) > thr) conv = false;
                    };
                    chk(dv0, alpha[0], v_d0); chk(dv1, alpha[1], v_d1); chk(dv2, alpha[2], v_d2); chk(dv3, alpha[3], v_d3);
                    if (conv) { st.last_nr_iterations = (uint32_t)iter; break; }
                }
            } else {
                i_nl[0] -= rclamp(f0 * 0.5, -0.01, 0.01);
                i_nl[1] -= rclamp(f1 * 0.5, -0.01, 0.01);
                i_nl[2] -= rclamp(f2 * 0.5, -0.01, 0.01);
                i_nl[3] -= rclamp(f3 * 0.5, -0.01, 0.01);
            }
        }
        for (int i = 0; i < N; i++) {
            double acc = v_pred_be[i];
            for (int j = 0; j < M; j++) acc += st.s_ni_be[i][j] * i_nl[j];
            v[i] = acc;
        }
    }

    bool finite = true;
    for (int i = 0; i < N; i++) if (!std::isfinite(v[i])) finite = false;
    if (!finite) {  // gen_tremolo.rs:3087-3096
        std::memcpy(st.v_prev, st.dc_operating_point, sizeof(st.v_prev));
        std::memcpy(st.i_nl_prev, TRM_DC_NL_I, sizeof(st.i_nl_prev));
        std::memcpy(st.i_nl_prev_prev, TRM_DC_NL_I, sizeof(st.i_nl_prev_prev));
        st.input_prev = 0.0;
        st.diag_nan_reset++;
        return 4.26480458363572357e0;
    }
    std::memcpy(st.v_prev, v, sizeof(st.v_prev));
    std::memcpy(st.i_nl_prev_prev, st.i_nl_prev, sizeof(st.i_nl_prev_prev));
    std::memcpy(st.i_nl_prev, i_nl, sizeof(st.i_nl_prev));
    const double raw = std::isfinite(v[0]) ? v[0] : 0.0;
    const double scaled = raw * 1.0;
    return std::isfinite(scaled) ? scaled : 0.0;
}

inline void CircuitState::set_default() {
    std::memcpy(v_prev, TRM_DC_OP, sizeof(v_prev));
    std::memcpy(i_nl_prev, TRM_DC_NL_I, sizeof(i_nl_prev));
    std::memcpy(i_nl_prev_prev, TRM_DC_NL_I, sizeof(i_nl_prev_prev));
    std::memcpy(dc_operating_point, TRM_DC_OP, sizeof(dc_operating_point));
    input_prev = 0.0;
    last_nr_iterations = 0;
    std::memcpy(a_neg, TRM_A_NEG_DEFAULT, sizeof(a_neg));
    std::memcpy(a_neg_be, TRM_A_NEG_BE_DEFAULT, sizeof(a_neg_be));
    std::memcpy(s, TRM_S_DEFAULT, sizeof(s));
    std::memcpy(k, TRM_K_DEFAULT, sizeof(k));
    std::memcpy(s_ni, TRM_S_NI_DEFAULT, sizeof(s_ni));
    std::memcpy(s_be, TRM_S_BE_DEFAULT, sizeof(s_be));
    std::memcpy(k_be, TRM_K_BE_DEFAULT, sizeof(k_be));
    std::memcpy(s_ni_be, TRM_S_NI_BE_DEFAULT, sizeof(s_ni_be));
    for (int i = 0; i < 50; i++) process_sample(0.0, *this);  // warmup()
}

}  // namespace trm

// ---- tremolo.rs (default feature set: circuit oscillator) --------------------------------
struct Tremolo {
    trm::CircuitState osc;
    double sample_rate, depth, r_ldr, ldr_envelope, ldr_attack, ldr_release;
    double ln_r_max, ln_min_minus_max;
    static constexpr double R_LDR_MIN = 9000.0, R_LDR_MAX = 1000000.0, GAMMA = 0.9;

    void settle_osc() {  // tremolo.rs:92-102 / 202-212
        osc.set_default();
        if (std::fabs(sample_rate - trm::SAMPLE_RATE) > 0.5) osc.set_sample_rate(sample_rate);
        const size_t n = (size_t)f64_as_u64(sample_rate * 2.0);
        for (size_t i = 0; i < n; i++) trm::process_sample(0.0, osc);
    }
    Tremolo(double depth_, double sr) {  // tremolo.rs:84-115 (depth stored unclamped)
        sample_rate = sr;
        settle_osc();
        depth = depth_;
        r_ldr = R_LDR_MAX;
        ldr_envelope = 0.0;
        ldr_attack = std::exp(-1.0 / (0.0025 * sr));
        ldr_release = std::exp(-1.0 / (0.035 * sr));
        ln_r_max = std::log(R_LDR_MAX);
        ln_min_minus_max = std::log(R_LDR_MIN) - std::log(R_LDR_MAX);
    }
    void set_depth(double d) { depth = rclamp(d, 0.0, 1.0); }  // tremolo.rs:117-119
    double shunt_impedance() const {  // tremolo.rs:152-167
        const double r_upper = 50000.0 * (1.0 - depth);
        const double r_lower = 50000.0 * depth;
        const double top = r_upper > 0.0 ? r_upper * 18000.0 / (r_upper + 18000.0) : 0.0;
        const double branch = 680.0 + r_ldr;
        const double low = r_lower > 0.0 ? r_lower * branch / (r_lower + branch) : 0.0;
        return top + low;
    }
    double process() {  // tremolo.rs:121-146
        const double v_out = trm::process_sample(0.0, osc);
        const double led_drive = rclamp((10.95 - v_out) / (10.95 - 0.70), 0.0, 1.0);
        const double coeff = led_drive > ldr_envelope ? ldr_attack : ldr_release;
        ldr_envelope = led_drive + coeff * (ldr_envelope - led_drive);
        const double drive = rclamp(ldr_envelope, 0.0, 1.0);
        if (drive < 1e-6) r_ldr = R_LDR_MAX;
        else {
            const double log_r = ln_r_max + ln_min_minus_max * std::pow(drive, GAMMA);
            r_ldr = std::exp(log_r);
        }
        return shunt_impedance();
    }
    void reset() {  // tremolo.rs:191-216
        settle_osc();
        ldr_envelope = 0.0;
        r_ldr = R_LDR_MAX;
    }
};

}  // namespace ow
