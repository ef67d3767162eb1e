// TEST INFRASTRUCTURE ONLY -- CPU oracle (parity checker), never on the product path.
//
// Restatement of the hand-written 8-node legacy DK preamp, crates/openwurli-dsp/src/dk_preamp_legacy.rs -- the preamp of
// the reference's DEFAULT build (openwurli-dsp/Cargo.toml:10-19: `melange-preamp` is opt-in).  Fixed S_base = (2C/T + G)^-1
// without R_ldr, Sherman-Morrison correction for the LDR on node FB, Newton-Raphson (<= 6 iterations) on the 2x2 BJT
// kernel with libm exp, Cin-R1 bilinear companion, zero-input shadow instance for pump cancellation.
// Pinned by the reference's own tests (tests/test_oracle_known_answers.py): SPICE DC operating point (:901-947), ~6 dB gain
// at 1 MOhm and higher at 19 kOhm (:949-981), melange-vs-legacy gate within 2 dB (dk_preamp/mod.rs:100-117), shadow pump
// cancellation < -100 dB (:1966-2025), S*A = I.  Sample-level parity: unpinned (no Rust toolchain here).
#pragma once
#include "ow_voice.hpp"

namespace ow {
namespace leg {

static constexpr int N = 8;
static constexpr double VCC = 15.0;
static constexpr double R1 = 22000.0, R2 = 2000000.0, R3 = 470000.0, RE1 = 33000.0, RC1 = 150000.0, RE2A = 270.0, RE2B = 820.0,
                        RC2 = 1800.0, R9 = 6800.0, R10 = 56000.0;                                    // :24-34
static constexpr double CIN = 0.022e-6, C3 = 100.0e-12, C4 = 100.0e-12, CE1 = 4.7e-6, CE2 = 22.0e-6;  // :37-41
static constexpr double IS = 3.03e-14, VT = 0.026, IS_OVER_VT = IS / VT, VBE_MAX = 0.85;              // :44-50
enum { BASE1 = 0, EMIT1, COLL1, EMIT2, EMIT2B, COLL2, OUT, FB };                                      // :53-60

typedef double Mat8[N][N];
typedef double Vec8[N];

static inline void mat_vec_mul(const Mat8 a, const Vec8 x, Vec8 y) {  // :76-86
    for (int i = 0; i < N; i++) {
        double sum = 0.0;
        for (int j = 0; j < N; j++) sum += a[i][j] * x[j];
        y[i] = sum;
    }
}

static inline void mat_inverse(const Mat8 m, Mat8 inv) {  // Gauss-Jordan with partial pivoting, :122-168
    double aug[N][2 * N];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) { aug[i][j] = m[i][j]; aug[i][N + j] = i == j ? 1.0 : 0.0; }
    for (int col = 0; col < N; col++) {
        double max_val = std::fabs(aug[col][col]);
        int max_row = col;
        for (int row = col + 1; row < N; row++)
            if (std::fabs(aug[row][col]) > max_val) { max_val = std::fabs(aug[row][col]); max_row = row; }
        if (max_row != col)
            for (int j = 0; j < 2 * N; j++) std::swap(aug[col][j], aug[max_row][j]);
        const double pivot = aug[col][col];
        for (int j = 0; j < 2 * N; j++) aug[col][j] /= pivot;
        for (int row = 0; row < N; row++) {
            if (row != col) {
                const double factor = aug[row][col];
                for (int j = 0; j < 2 * N; j++) aug[row][j] -= factor * aug[col][j];
            }
        }
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) inv[i][j] = aug[i][N + j];
}

static inline void stamp_resistor(Mat8 g, int i, int j, double r) {  // :646-652
    const double cond = 1.0 / r;
    g[i][i] += cond; g[j][j] += cond; g[i][j] -= cond; g[j][i] -= cond;
}
static inline void stamp_capacitor(Mat8 c, int i, int j, double cap) {  // :654-659
    c[i][i] += cap; c[j][j] += cap; c[i][j] -= cap; c[j][i] -= cap;
}
static inline double bjt_ic(double vbe) {  // :668-671
    const double v = rclamp(vbe, -1.0, VBE_MAX);
    return IS * (std::exp(v / VT) - 1.0);
}
static inline void bjt_ic_gm(double vbe, double& ic, double& gm) {  // :686-690
    const double v = rclamp(vbe, -1.0, VBE_MAX);
    const double e = std::exp(v / VT);
    ic = IS * (e - 1.0);
    gm = IS_OVER_VT * e;
}
static inline void compute_k(const Mat8 s, double k[2][2]) {  // :424-435
    k[0][0] = s[BASE1][EMIT1] - s[BASE1][COLL1] - s[EMIT1][EMIT1] + s[EMIT1][COLL1];
    k[0][1] = s[BASE1][EMIT2] - s[BASE1][COLL2] - s[EMIT1][EMIT2] + s[EMIT1][COLL2];
    k[1][0] = s[COLL1][EMIT1] - s[COLL1][COLL1] - s[EMIT2][EMIT1] + s[EMIT2][COLL1];
    k[1][1] = s[COLL1][EMIT2] - s[COLL1][COLL2] - s[EMIT2][EMIT2] + s[EMIT2][COLL2];
}

struct DkState {  // :237-256
    double j_cin, cin_rhs_prev;
    Vec8 v;
    double i_nl[2], v_nl[2];
    void at_dc(double g_cin, const double v_nl_dc[2], const Vec8 v_dc) {
        j_cin = g_cin * v_dc[BASE1];
        cin_rhs_prev = g_cin * v_dc[BASE1];
        for (int i = 0; i < N; i++) v[i] = v_dc[i];
        i_nl[0] = bjt_ic(v_nl_dc[0]); i_nl[1] = bjt_ic(v_nl_dc[1]);
        v_nl[0] = v_nl_dc[0]; v_nl[1] = v_nl_dc[1];
    }
};

struct DkPreamp {
    Mat8 s_base, a_neg_base, g_dc_base;
    double k[2][2];
    Vec8 two_w, s_fb_col, s_fb_row, v_dc;
    double s_fb_fb, nv_sfb[2], sfb_ni[2], g_cin, c_cin, gc_1pc;
    DkState main_st, shadow_st;
    double r_ldr, g_ldr, g_ldr_prev;
    uint64_t nr_iter_hist[8] = {0};  // diagnostics of the MAIN instance (not in the reference)
    uint64_t nan_reset = 0;

    static void full_dc_solve(const Mat8 g_dc_base, const Vec8 w, double r_ldr, double v_nl[2], Vec8 v_dc) {  // :370-412
        Mat8 g_full, s_dc;
        std::memcpy(g_full, g_dc_base, sizeof(Mat8));
        g_full[FB][FB] += 1.0 / r_ldr;
        mat_inverse(g_full, s_dc);
        double k_dc[2][2];
        compute_k(s_dc, k_dc);
        Vec8 sv;
        mat_vec_mul(s_dc, w, sv);
        const double p_dc[2] = {sv[BASE1] - sv[EMIT1], sv[COLL1] - sv[EMIT2]};
        v_nl[0] = 0.56; v_nl[1] = 0.66;
        for (int iter = 0; iter < 100; iter++) {
            double ic0, gm0, ic1, gm1;
            bjt_ic_gm(v_nl[0], ic0, gm0);
            bjt_ic_gm(v_nl[1], ic1, gm1);
            const double f0 = v_nl[0] - p_dc[0] - k_dc[0][0] * ic0 - k_dc[0][1] * ic1;
            const double f1 = v_nl[1] - p_dc[1] - k_dc[1][0] * ic0 - k_dc[1][1] * ic1;
            if (std::fabs(f0) < 1e-12 && std::fabs(f1) < 1e-12) break;
            const double j00 = 1.0 - k_dc[0][0] * gm0;
            const double j01 = -k_dc[0][1] * gm1;
            const double j10 = -k_dc[1][0] * gm0;
            const double j11 = 1.0 - k_dc[1][1] * gm1;
            const double det = j00 * j11 - j01 * j10;
            const double inv_det = 1.0 / det;
            const double dv0 = inv_det * (j11 * f0 - j01 * f1);
            const double dv1 = inv_det * (j00 * f1 - j10 * f0);
            const double max_step = 2.0 * VT;
            v_nl[0] -= rclamp(dv0, -max_step, max_step);
            v_nl[1] -= rclamp(dv1, -max_step, max_step);
        }
        const double ic[2] = {bjt_ic(v_nl[0]), bjt_ic(v_nl[1])};
        Vec8 dc_rhs;
        for (int i = 0; i < N; i++) dc_rhs[i] = w[i];
        dc_rhs[EMIT1] += ic[0]; dc_rhs[COLL1] -= ic[0]; dc_rhs[EMIT2] += ic[1]; dc_rhs[COLL2] -= ic[1];
        mat_vec_mul(s_dc, dc_rhs, v_dc);
    }

    explicit DkPreamp(double sample_rate) {  // :269-366
        const double t = 1.0 / sample_rate;
        const double two_over_t = 2.0 / t;
        const double alpha_cin = 2.0 * R1 * CIN * sample_rate;
        g_cin = (2.0 * CIN * sample_rate) / (1.0 + alpha_cin);
        c_cin = (1.0 - alpha_cin) / (1.0 + alpha_cin);
        gc_1pc = g_cin * (1.0 + c_cin);
        Mat8 g_base;
        Vec8 w;
        std::memset(g_base, 0, sizeof(g_base));
        std::memset(w, 0, sizeof(w));
        g_base[BASE1][BASE1] += 1.0 / R2;
        w[BASE1] += VCC / R2;
        g_base[BASE1][BASE1] += 1.0 / R3;
        g_base[EMIT1][EMIT1] += 1.0 / RE1;
        g_base[COLL1][COLL1] += 1.0 / RC1;
        w[COLL1] += VCC / RC1;
        stamp_resistor(g_base, EMIT2, EMIT2B, RE2A);
        g_base[EMIT2B][EMIT2B] += 1.0 / RE2B;
        g_base[COLL2][COLL2] += 1.0 / RC2;
        w[COLL2] += VCC / RC2;
        stamp_resistor(g_base, COLL2, OUT, R9);
        stamp_resistor(g_base, OUT, FB, R10);
        std::memcpy(g_dc_base, g_base, sizeof(Mat8));
        g_base[BASE1][BASE1] += g_cin;
        Mat8 c;
        std::memset(c, 0, sizeof(c));
        stamp_capacitor(c, COLL1, BASE1, C3);
        stamp_capacitor(c, COLL2, COLL1, C4);
        stamp_capacitor(c, EMIT1, FB, CE1);
        stamp_capacitor(c, EMIT2, EMIT2B, CE2);
        Mat8 a_base;
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                const double tc = two_over_t * c[i][j];
                a_base[i][j] = tc + g_base[i][j];
                a_neg_base[i][j] = tc - g_base[i][j];
            }
        for (int i = 0; i < N; i++) two_w[i] = 2.0 * w[i];
        mat_inverse(a_base, s_base);
        compute_k(s_base, k);
        for (int i = 0; i < N; i++) { s_fb_col[i] = s_base[i][FB]; s_fb_row[i] = s_base[FB][i]; }
        s_fb_fb = s_base[FB][FB];
        nv_sfb[0] = s_fb_col[BASE1] - s_fb_col[EMIT1];
        nv_sfb[1] = s_fb_col[COLL1] - s_fb_col[EMIT2];
        sfb_ni[0] = s_fb_row[EMIT1] - s_fb_row[COLL1];
        sfb_ni[1] = s_fb_row[EMIT2] - s_fb_row[COLL2];
        const double r_ldr_init = 1000000.0;
        double v_nl_dc[2];
        full_dc_solve(g_dc_base, w, r_ldr_init, v_nl_dc, v_dc);
        main_st.at_dc(g_cin, v_nl_dc, v_dc);
        shadow_st = main_st;
        r_ldr = r_ldr_init;
        g_ldr = 1.0 / r_ldr_init;
        g_ldr_prev = 1.0 / r_ldr_init;
    }

    double dk_step(DkState& st, double input, uint64_t* hist) const {  // :447-554
        Vec8 rhs;
        mat_vec_mul(a_neg_base, st.v, rhs);
        rhs[FB] -= g_ldr_prev * st.v[FB];
        const double cin_rhs_now = g_cin * input + st.j_cin;
        rhs[BASE1] += cin_rhs_now + st.cin_rhs_prev;
        rhs[EMIT1] += st.i_nl[0];
        rhs[COLL1] -= st.i_nl[0];
        rhs[EMIT2] += st.i_nl[1];
        rhs[COLL2] -= st.i_nl[1];
        for (int i = 0; i < N; i++) rhs[i] += two_w[i];
        Vec8 v_pred_base;
        mat_vec_mul(s_base, rhs, v_pred_base);
        const double sm_k = g_ldr / (1.0 + s_fb_fb * g_ldr);
        const double sm_vpred = sm_k * v_pred_base[FB];
        Vec8 v_pred;
        for (int i = 0; i < N; i++) v_pred[i] = v_pred_base[i] - sm_vpred * s_fb_col[i];
        const double p[2] = {v_pred[BASE1] - v_pred[EMIT1], v_pred[COLL1] - v_pred[EMIT2]};
        const double k00 = k[0][0] - sm_k * nv_sfb[0] * sfb_ni[0];
        const double k01 = k[0][1] - sm_k * nv_sfb[0] * sfb_ni[1];
        const double k10 = k[1][0] - sm_k * nv_sfb[1] * sfb_ni[0];
        const double k11 = k[1][1] - sm_k * nv_sfb[1] * sfb_ni[1];
        double v_nl[2] = {st.v_nl[0], st.v_nl[1]};
        int iters = 0;
        for (int iter = 0; iter < 6; iter++) {
            double ic0, gm0, ic1, gm1;
            bjt_ic_gm(v_nl[0], ic0, gm0);
            bjt_ic_gm(v_nl[1], ic1, gm1);
            const double f0 = v_nl[0] - p[0] - k00 * ic0 - k01 * ic1;
            const double f1 = v_nl[1] - p[1] - k10 * ic0 - k11 * ic1;
            if (std::fabs(f0) < 1e-9 && std::fabs(f1) < 1e-9) break;
            const double j00 = 1.0 - k00 * gm0;
            const double j01 = -k01 * gm1;
            const double j10 = -k10 * gm0;
            const double j11 = 1.0 - k11 * gm1;
            const double det = j00 * j11 - j01 * j10;
            if (std::fabs(det) < 1e-30) break;
            const double inv_det = 1.0 / det;
            v_nl[0] -= inv_det * (j11 * f0 - j01 * f1);
            v_nl[1] -= inv_det * (j00 * f1 - j10 * f0);
            iters++;
        }
        if (hist) hist[iters]++;
        const double ic_new[2] = {bjt_ic(v_nl[0]), bjt_ic(v_nl[1])};
        const double sfb_ni_dot_ic = sfb_ni[0] * ic_new[0] + sfb_ni[1] * ic_new[1];
        for (int i = 0; i < N; i++) {
            const double s_ni_i = ic_new[0] * (s_base[i][EMIT1] - s_base[i][COLL1]) + ic_new[1] * (s_base[i][EMIT2] - s_base[i][COLL2]);
            st.v[i] = v_pred[i] + s_ni_i - sm_k * sfb_ni_dot_ic * s_fb_col[i];
        }
        st.cin_rhs_prev = cin_rhs_now;
        const double dv_cin = input - st.v[BASE1];
        st.j_cin = -gc_1pc * dv_cin - c_cin * st.j_cin;
        st.i_nl[0] = ic_new[0]; st.i_nl[1] = ic_new[1];
        st.v_nl[0] = v_nl[0]; st.v_nl[1] = v_nl[1];
        return st.v[OUT];
    }

    double process_sample(double input, double* pump_out = nullptr) {  // :557-618
        const double main_out = dk_step(main_st, input, nr_iter_hist);
        const double pump = dk_step(shadow_st, 0.0, nullptr);
        if (pump_out) *pump_out = pump;
        g_ldr_prev = g_ldr;
        const double result = main_out - pump;
        if (!std::isfinite(result)) { nan_reset++; reset(); return 0.0; }
        return result;
    }
    void set_ldr_resistance(double r_ldr_path) {  // :620-626 (f64::max: a NaN argument yields the other operand)
        const double new_r = r_ldr_path != r_ldr_path ? 1000.0 : (r_ldr_path > 1000.0 ? r_ldr_path : 1000.0);
        if (std::fabs(new_r - r_ldr) > 0.01) { r_ldr = new_r; g_ldr = 1.0 / new_r; }
    }
    void reset() {  // :628-642
        Vec8 w;
        for (int i = 0; i < N; i++) w[i] = two_w[i] * 0.5;
        double v_nl_dc[2];
        full_dc_solve(g_dc_base, w, r_ldr, v_nl_dc, v_dc);
        g_ldr = 1.0 / r_ldr;
        g_ldr_prev = g_ldr;
        main_st.at_dc(g_cin, v_nl_dc, v_dc);
        shadow_st = main_st;
    }
};

}  // namespace leg
}  // namespace ow
