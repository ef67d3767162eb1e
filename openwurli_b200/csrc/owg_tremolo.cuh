// Twin-T tremolo oscillator (gen_tremolo.rs: N=7 nodes, M=4 = two Ebers-Moll BJTs, Schur-complement
// NR with pnjlim, BE fallback) and the LDR law of tremolo.rs, device side.  Input-independent: one
// sequence per preamp sample rate, shared bit-exactly by every instance of a launch.
#pragma once
#include "owg_device.cuh"

namespace owgd {

constexpr int TN = 7, TM = 4, T_MAX_ITER = 50;

struct TrmMats {  // rate-dependent matrices (gen_tremolo.rs:1917-1936; s_sub family is never read by process_sample)
    double a_neg[TN][TN], a_neg_be[TN][TN];
    double s[TN][TN], k[TM][TM], s_ni[TN][TM];
    double s_be[TN][TN], k_be[TM][TM], s_ni_be[TN][TM];
};
struct TrmState {
    double v[TN], il[TM], ilpp[TM], xin_prev;
};
struct TrmDiag { uint32_t hist[16]; uint32_t be_fallback, nan_reset; };
struct BjtK;


// gen_tremolo.rs:2273-2342; returns false if singular (out untouched)
__device__ inline bool trm_invert(const double a[TN][TN], double out[TN][TN]) {
    double lu[TN][TN];
    int perm[TN];
    for (int i = 0; i < TN; i++) { perm[i] = i; for (int j = 0; j < TN; j++) lu[i][j] = a[i][j]; }
    for (int k = 0; k < TN; k++) {
        int max_row = k;
        double max_val = fabs(lu[k][k]);
        for (int i = k + 1; i < TN; i++) { const double v = fabs(lu[i][k]); if (v > max_val) { max_val = v; max_row = i; } }
        if (max_val < 1e-30) return false;
        if (max_row != k) {
            for (int j = 0; j < TN; j++) { const double t = lu[k][j]; lu[k][j] = lu[max_row][j]; lu[max_row][j] = t; }
            const int t = perm[k]; perm[k] = perm[max_row]; perm[max_row] = t;
        }
        const double pivot = lu[k][k];
        for (int i = k + 1; i < TN; i++) {
            const double m = lu[i][k] / pivot;
            lu[i][k] = m;
            for (int j = k + 1; j < TN; j++) lu[i][j] -= m * lu[k][j];
        }
    }
    double res[TN][TN];
    for (int col = 0; col < TN; col++) {
        double b[TN];
        for (int i = 0; i < TN; i++) b[i] = 0.0;
        int start = TN;
        for (int i = 0; i < TN; i++) if (perm[i] == col) { b[i] = 1.0; start = i; break; }
        for (int i = start + 1; i < TN; i++) {
            double sum = b[i];
            for (int j = start; j < i; j++) sum -= lu[i][j] * b[j];
            b[i] = sum;
        }
        for (int i = TN - 1; i >= 0; i--) {
            double sum = b[i];
            for (int j = i + 1; j < TN; j++) sum -= lu[i][j] * b[j];
            const double pivot = lu[i][i];
            if (fabs(pivot) < 1e-30) return false;
            b[i] = sum / pivot;
        }
        for (int i = 0; i < TN; i++) res[i][col] = b[i];
    }
    for (int i = 0; i < TN; i++) for (int j = 0; j < TN; j++) out[i][j] = res[i][j];
    return true;
}

__device__ inline void trm_products(const double s[TN][TN], double k[TM][TM], double s_ni[TN][TM]) {  // gen_tremolo.rs:2172-2195
    for (int i = 0; i < TM; i++)
        for (int j = 0; j < TM; j++) {
            double sum = 0.0;
            for (int a = 0; a < TN; a++) {
                double t = 0.0;
                for (int b = 0; b < TN; b++) t += s[a][b] * TRM_N_I[b][j];
                sum += TRM_N_V[i][a] * t;
            }
            k[i][j] = sum;
        }
    for (int i = 0; i < TN; i++)
        for (int j = 0; j < TM; j++) {
            double sum = 0.0;
            for (int a = 0; a < TN; a++) sum += s[i][a] * TRM_N_I[a][j];
            s_ni[i][j] = sum;
        }
}

__device__ inline void trm_defaults(TrmMats& m) {  // gen_tremolo.rs:1998-2011
    for (int i = 0; i < TN; i++)
        for (int j = 0; j < TN; j++) {
            m.a_neg[i][j] = TRM_A_NEG_DEFAULT[i][j]; m.a_neg_be[i][j] = TRM_A_NEG_BE_DEFAULT[i][j];
            m.s[i][j] = TRM_S_DEFAULT[i][j]; m.s_be[i][j] = TRM_S_BE_DEFAULT[i][j];
        }
    for (int i = 0; i < TM; i++) for (int j = 0; j < TM; j++) { m.k[i][j] = TRM_K_DEFAULT[i][j]; m.k_be[i][j] = TRM_K_BE_DEFAULT[i][j]; }
    for (int i = 0; i < TN; i++) for (int j = 0; j < TM; j++) { m.s_ni[i][j] = TRM_S_NI_DEFAULT[i][j]; m.s_ni_be[i][j] = TRM_S_NI_BE_DEFAULT[i][j]; }
}

__device__ __noinline__ void trm_rebuild(TrmMats& m, double rate) {  // gen_tremolo.rs:2139-2222
    const double alpha = 2.0 * rate, alpha_be = rate;
    double a[TN][TN], a_be[TN][TN];
    for (int i = 0; i < TN; i++)
        for (int j = 0; j < TN; j++) {
            a[i][j] = TRM_G[i][j] + alpha * TRM_C[i][j];
            m.a_neg[i][j] = alpha * TRM_C[i][j] - TRM_G[i][j];
            a_be[i][j] = TRM_G[i][j] + alpha_be * TRM_C[i][j];
            m.a_neg_be[i][j] = alpha_be * TRM_C[i][j];
        }
    for (int j = 0; j < TN; j++) { m.a_neg[6][j] = 0.0; m.a_neg_be[6][j] = 0.0; }
    if (trm_invert(a, m.s)) trm_products(m.s, m.k, m.s_ni);
    if (trm_invert(a_be, m.s_be)) trm_products(m.s_be, m.k_be, m.s_ni_be);
}

struct Bjt { double ic, ib, j0, j1, j2, j3; };
// Loop-invariant sub-expressions of bjt_evaluate (same operations on the same constants the reference
// re-evaluates on every call, gen_tremolo.rs:1566-1636), computed once.
struct BjtK { double is, nf_vt, nr_vt, is_bf, is_br, is_nfvt, is_nrvt, is_br_nrvt, is_bf_nfvt; Recip r_nf, r_nr; };
__device__ __forceinline__ BjtK bjt_consts(double is, double vt, double nf, double nr, double bf, double br) {
    BjtK k;
    k.is = is; k.nf_vt = nf * vt; k.nr_vt = nr * vt;
    k.is_bf = is / bf; k.is_br = is / br;
    k.is_nfvt = is / k.nf_vt; k.is_nrvt = is / k.nr_vt;
    k.is_br_nrvt = is / (br * k.nr_vt); k.is_bf_nfvt = is / (bf * k.nf_vt);
    k.r_nf = recip_prepare(k.nf_vt); k.r_nr = recip_prepare(k.nr_vt);
    return k;
}
// bjt_evaluate Ebers-Moll branch (gen_tremolo.rs:1566-1636) for use_gp=false, ISE=ISC=0, sign=+1.
template <bool EXACT>
__device__ __forceinline__ Bjt bjt_em(double vbe, double vbc, const BjtK& k, DivPolicy<EXACT>& D) {
    const double exp_be = fast_exp(EXACT ? vbe / k.nf_vt : D.div(vbe, k.r_nf));
    const double exp_bc = fast_exp(EXACT ? vbc / k.nr_vt : D.div(vbc, k.r_nr));
    const double i_cc = k.is * (exp_be - exp_bc);
    const double ib_fwd = k.is_bf * (exp_be - 1.0);
    const double ib_rev = k.is_br * (exp_bc - 1.0);
    Bjt o;
    o.ic = i_cc - k.is_br * (exp_bc - 1.0);
    o.ib = ib_fwd + ib_rev;
    o.j0 = k.is_nfvt * exp_be;
    o.j1 = -k.is_nrvt * exp_bc - k.is_br_nrvt * exp_bc;
    o.j2 = k.is_bf_nfvt * exp_be;
    o.j3 = k.is_br_nrvt * exp_bc;
    return o;
}

struct TrmK { BjtK q0, q1; double input_conductance; };
__device__ __forceinline__ TrmK trm_consts() {
    TrmK k;
    k.q0 = bjt_consts(TRM_DEVICE_0_IS, TRM_DEVICE_0_VT, TRM_DEVICE_0_NF, TRM_DEVICE_0_NR, TRM_DEVICE_0_BETA_F, TRM_DEVICE_0_BETA_R);
    k.q1 = bjt_consts(TRM_DEVICE_1_IS, TRM_DEVICE_1_VT, TRM_DEVICE_1_NF, TRM_DEVICE_1_NR, TRM_DEVICE_1_BETA_F, TRM_DEVICE_1_BETA_R);
    k.input_conductance = 1.0 / TRM_INPUT_RESISTANCE;
    return k;
}
__device__ __forceinline__ double trm_pnjlim(double vnew, double vold, double vt, double vcrit) { return pnjlim(vnew, vold, vt, vcrit); }

// gen_tremolo.rs:2515-2561; fully unrolled so that every index is a compile-time constant (registers, no local memory)
template <bool EXACT>
__device__ __forceinline__ void trm_solve4(double a[4][4], double b[4], bool& singular, DivPolicy<EXACT>& D) {
    singular = false;
    Recip rp[4];
#pragma unroll
    for (int col = 0; col < 4; col++) { rp[col].r = 0.0; rp[col].nb = 0.0; rp[col].b = 1.0; }
#pragma unroll
    for (int col = 0; col < 4; col++) {
        if (!singular) {
            int max_row = col;
            double max_val = fabs(a[col][col]);
#pragma unroll
            for (int row = col + 1; row < 4; row++) if (fabs(a[row][col]) > max_val) { max_val = fabs(a[row][col]); max_row = row; }
            if (max_val < KC(14)) singular = true;
            else {
#pragma unroll
                for (int row = col + 1; row < 4; row++) {
                    if (max_row == row) {
#pragma unroll
                        for (int j = 0; j < 4; j++) { const double t = a[col][j]; a[col][j] = a[row][j]; a[row][j] = t; }
                        const double t = b[col]; b[col] = b[row]; b[row] = t;
                    }
                }
                rp[col] = D.prep(a[col][col]);  // one reciprocal per pivot, shared by the column's factors and by
                                                       // the back-substitution (bit-identical to dividing each time)
#pragma unroll
                for (int row = col + 1; row < 4; row++) {
                    const double factor = D.div(a[row][col], rp[col]);
#pragma unroll
                    for (int j = col + 1; j < 4; j++) a[row][j] -= factor * a[col][j];
                    b[row] -= factor * b[col];
                }
            }
        }
    }
    if (!singular) {
#pragma unroll
        for (int i = 3; i >= 0; i--) {
            if (!singular) {
                double sum = b[i];
#pragma unroll
                for (int j = i + 1; j < 4; j++) sum -= a[i][j] * b[j];
                if (fabs(a[i][i]) < KC(14)) singular = true;
                else b[i] = D.div(sum, rp[i]);
            }
        }
    }
}

__device__ __forceinline__ void trm_jac(const Bjt& q0, const Bjt& q1, const double k[TM][TM], double a[4][4]) {  // :2497-2512
    a[0][0] = 1.0 - q0.j0 * k[0][0] - q0.j1 * k[1][0];
    a[0][1] = 0.0 - q0.j0 * k[0][1] - q0.j1 * k[1][1];
    a[0][2] = 0.0 - q0.j0 * k[0][2] - q0.j1 * k[1][2];
    a[0][3] = 0.0 - q0.j0 * k[0][3] - q0.j1 * k[1][3];
    a[1][0] = 0.0 - q0.j2 * k[0][0] - q0.j3 * k[1][0];
    a[1][1] = 1.0 - q0.j2 * k[0][1] - q0.j3 * k[1][1];
    a[1][2] = 0.0 - q0.j2 * k[0][2] - q0.j3 * k[1][2];
    a[1][3] = 0.0 - q0.j2 * k[0][3] - q0.j3 * k[1][3];
    a[2][0] = 0.0 - q1.j0 * k[2][0] - q1.j1 * k[3][0];
    a[2][1] = 0.0 - q1.j0 * k[2][1] - q1.j1 * k[3][1];
    a[2][2] = 1.0 - q1.j0 * k[2][2] - q1.j1 * k[3][2];
    a[2][3] = 0.0 - q1.j0 * k[2][3] - q1.j1 * k[3][3];
    a[3][0] = 0.0 - q1.j2 * k[2][0] - q1.j3 * k[3][0];
    a[3][1] = 0.0 - q1.j2 * k[2][1] - q1.j3 * k[3][1];
    a[3][2] = 0.0 - q1.j2 * k[2][2] - q1.j3 * k[3][2];
    a[3][3] = 1.0 - q1.j2 * k[2][3] - q1.j3 * k[3][3];
}

// BE fallback of the oscillator (gen_tremolo.rs:2757-3083); v receives the BE voltages.
// Cold path: state and results travel through a scratch buffer (shared memory) so the hot path's register arrays are
// never address-taken.  sc: [0..6] v_prev, [7..10] il, [11..14] ilpp  ->  [15..21] v, [22..25] il.
#define OWG_TRM_SCRATCH 34
__device__ __noinline__ uint32_t trm_be(double input, double* sc, const TrmMats& m, const TrmK& kq) {
    TrmState st;
    for (int i = 0; i < TN; i++) st.v[i] = sc[i];
    for (int i = 0; i < TM; i++) { st.il[i] = sc[7 + i]; st.ilpp[i] = sc[11 + i]; }
    double v[TN], il[TM];
    const double input_conductance = kq.input_conductance;
    const double vt0 = TRM_DEVICE_0_VT, vt1 = TRM_DEVICE_1_VT;
    double rhs_be[TN], v_pred_be[TN], p_be[TM];
    for (int i = 0; i < TN; i++) {
        double sum = TRM_RHS_CONST_BE[i];
        for (int j = 0; j < TN; j++) sum += m.a_neg_be[i][j] * st.v[j];
        for (int j = 0; j < TM; j++) sum += TRM_N_I[i][j] * st.il[j];
        rhs_be[i] = sum;
    }
    rhs_be[0] += input * input_conductance;
    for (int i = 0; i < TN; i++) {
        double sum = 0.0;
        for (int j = 0; j < TN; j++) sum += m.s_be[i][j] * rhs_be[j];
        v_pred_be[i] = sum;
    }
    for (int i = 0; i < TM; i++) {
        double sum = 0.0;
        for (int j = 0; j < TN; j++) sum += TRM_N_V[i][j] * v_pred_be[j];
        p_be[i] = sum;
    }
    for (int i = 0; i < TM; i++) il[i] = 2.0 * st.il[i] - st.ilpp[i];
    uint32_t result = T_MAX_ITER;
    const double (*kb)[TM] = m.k_be;
    for (int iter = 0; iter < T_MAX_ITER; iter++) {
        const double v_d0 = p_be[0] + kb[0][0] * il[0] + kb[0][1] * il[1] + kb[0][2] * il[2] + kb[0][3] * il[3];
        const double v_d1 = p_be[1] + kb[1][0] * il[0] + kb[1][1] * il[1] + kb[1][2] * il[2] + kb[1][3] * il[3];
        const double v_d2 = p_be[2] + kb[2][0] * il[0] + kb[2][1] * il[1] + kb[2][2] * il[2] + kb[2][3] * il[3];
        const double v_d3 = p_be[3] + kb[3][0] * il[0] + kb[3][1] * il[1] + kb[3][2] * il[2] + kb[3][3] * il[3];
        DivPolicy<true> D;
        const Bjt q0 = bjt_em(v_d0, v_d1, kq.q0, D);
        const Bjt q1 = bjt_em(v_d2, v_d3, kq.q1, D);
        const double f0 = il[0] - q0.ic, f1 = il[1] - q0.ib, f2 = il[2] - q1.ic, f3 = il[3] - q1.ib;
        double a[4][4];
        trm_jac(q0, q1, kb, a);
        double b[4] = {f0, f1, f2, f3};
        bool singular;
        trm_solve4(a, b, singular, D);
        if (!singular) {
            const double d0 = b[0], d1 = b[1], d2 = b[2], d3 = b[3];
            const double dv0 = -(kb[0][0] * d0 + kb[0][1] * d1 + kb[0][2] * d2 + kb[0][3] * d3);
            const double dv1 = -(kb[1][0] * d0 + kb[1][1] * d1 + kb[1][2] * d2 + kb[1][3] * d3);
            const double dv2 = -(kb[2][0] * d0 + kb[2][1] * d1 + kb[2][2] * d2 + kb[2][3] * d3);
            const double dv3 = -(kb[3][0] * d0 + kb[3][1] * d1 + kb[3][2] * d2 + kb[3][3] * d3);
            double al[4] = {1.0, 1.0, 1.0, 1.0};
            bool any_limited = false;
            const double dvs[4] = {dv0, dv1, dv2, dv3};
            const double vds[4] = {v_d0, v_d1, v_d2, v_d3};
            for (int q = 0; q < 4; q++) {
                if (fabs(dvs[q]) > 1e-4) {
                    const double v_lim = trm_pnjlim(vds[q] + dvs[q], vds[q], q < 2 ? vt0 : vt1, q < 2 ? TRM_DEVICE_0_VCRIT : TRM_DEVICE_1_VCRIT);
                    const double ratio = fmax((v_lim - vds[q]) / dvs[q], 0.01);
                    if (ratio < al[q]) { al[q] = ratio; if (ratio < 1.0) any_limited = true; }
                }
            }
            { const double d = fmin(al[0], al[1]); al[0] = d; al[1] = d; }
            { const double d = fmin(al[2], al[3]); al[2] = d; al[3] = d; }
            const double max_dv = fmax(fmax(fmax(fabs(dv0 * al[0]), fabs(dv1 * al[1])), fabs(dv2 * al[2])), fabs(dv3 * al[3]));
            if (max_dv > 3.5) {
                const double factor = fmax(3.5 / max_dv, 0.1);
                for (int q = 0; q < 4; q++) al[q] *= factor;
            }
            il[0] -= al[0] * d0; il[1] -= al[1] * d1; il[2] -= al[2] * d2; il[3] -= al[3] * d3;
            if (!any_limited) {
                bool conv = true;
                for (int q = 0; q < 4; q++) {
                    const double step = dvs[q] * al[q];
                    const double v_new = vds[q] + step;
                    const double thr = 1e-3 * fmax(fabs(vds[q]), fabs(v_new)) + 1e-6;
                    if (fabs(step) > thr) conv = false;
                }
                if (conv) { result = (uint32_t)iter; break; }
            }
        } else {
            il[0] -= rclamp(f0 * 0.5, -0.01, 0.01);
            il[1] -= rclamp(f1 * 0.5, -0.01, 0.01);
            il[2] -= rclamp(f2 * 0.5, -0.01, 0.01);
            il[3] -= rclamp(f3 * 0.5, -0.01, 0.01);
        }
    }
    for (int i = 0; i < TN; i++) {
        double acc = v_pred_be[i];
        for (int j = 0; j < TM; j++) acc += m.s_ni_be[i][j] * il[j];
        v[i] = acc;
    }
    for (int i = 0; i < TN; i++) sc[15 + i] = v[i];
    for (int i = 0; i < TM; i++) sc[22 + i] = il[i];
    return result;
}

// One iteration of the trapezoidal Newton loop (gen_tremolo.rs:2423-2745). Returns true when converged.
template <bool EXACT>
__device__ __forceinline__ bool trm_nr_iter(const double p[TM], const double (*k)[TM], const TrmK& kq, double il[TM], unsigned& bad_out) {
    DivPolicy<EXACT> D;
    const double vt0 = TRM_DEVICE_0_VT, vt1 = TRM_DEVICE_1_VT;
    bool converged = false;
    {
        const double v_d0 = p[0] + k[0][0] * il[0] + k[0][1] * il[1] + k[0][2] * il[2] + k[0][3] * il[3];
        const double v_d1 = p[1] + k[1][0] * il[0] + k[1][1] * il[1] + k[1][2] * il[2];
        const double v_d2 = p[2] + k[2][0] * il[0] + k[2][1] * il[1] + k[2][3] * il[3];
        const double v_d3 = p[3] + k[3][0] * il[0] + k[3][1] * il[1] + k[3][2] * il[2] + k[3][3] * il[3];
        const Bjt q0 = bjt_em(v_d0, v_d1, kq.q0, D);
        const Bjt q1 = bjt_em(v_d2, v_d3, kq.q1, D);
        const double f0 = il[0] - q0.ic, f1 = il[1] - q0.ib, f2 = il[2] - q1.ic, f3 = il[3] - q1.ib;
        double a[4][4];
        trm_jac(q0, q1, k, a);
        double b[4] = {f0, f1, f2, f3};
        bool singular;
        trm_solve4(a, b, singular, D);
        if (!singular) {
            const double d0 = b[0], d1 = b[1], d2 = b[2], d3 = b[3];
            const double it0 = il[0] - d0, it1 = il[1] - d1, it2 = il[2] - d2, it3 = il[3] - d3;
            const double vt_0 = p[0] + k[0][0] * it0 + k[0][1] * it1 + k[0][2] * it2 + k[0][3] * it3;
            const double vt_1 = p[1] + k[1][0] * it0 + k[1][1] * it1 + k[1][2] * it2 + k[1][3] * it3;
            const double vt_2 = p[2] + k[2][0] * it0 + k[2][1] * it1 + k[2][2] * it2 + k[2][3] * it3;
            const double vt_3 = p[3] + k[3][0] * it0 + k[3][1] * it1 + k[3][2] * it2 + k[3][3] * it3;
            bool any_limited = false;
            const double vds[4] = {v_d0, v_d1, v_d2, v_d3};
            const double vts[4] = {vt_0, vt_1, vt_2, vt_3};
            double dvt[4], vlim[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                dvt[q] = vts[q] - vds[q];
                vlim[q] = fabs(dvt[q]) > KC(13) ? trm_pnjlim(vts[q], vds[q], q < 2 ? vt0 : vt1, q < 2 ? TRM_DEVICE_0_VCRIT : TRM_DEVICE_1_VCRIT) : vts[q];
            }
            double ga = 1.0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double dv_lim = vlim[q] - vds[q];
                if (fabs(dvt[q]) > KC(14)) {
                    // vlim == vtrial (no limiting) makes dv_lim the very same subtraction as dvt: the ratio is exactly 1
                    const double r = (vlim[q] == vts[q]) ? 1.0 : ((dvt[q] * dv_lim < 0.0) ? 0.0 : rclamp(dv_lim / dvt[q], 0.0, 1.0));
                    if (r < ga) { ga = r; any_limited = true; }
                }
            }
            {
                const double max_dv = fmax(fmax(fmax(fabs(dvt[0] * ga), fabs(dvt[1] * ga)), fabs(dvt[2] * ga)), fabs(dvt[3] * ga));
                if (max_dv > 3.5) { ga *= fmax(3.5 / max_dv, 0.1); any_limited = true; }
            }
            il[0] -= ga * d0; il[1] -= ga * d1; il[2] -= ga * d2; il[3] -= ga * d3;
            if (!any_limited) {
                bool conv = true;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double dv = dvt[q] * ga;
                    const double thr = KC(9) * fmax(fabs(vds[q]), fabs(vds[q] + dv)) + KC(10);
                    if (fabs(dv) > thr) conv = false;
                }
                converged = conv;
            }
        } else {
            { const double c = fmax(fabs(il[0]) * KC(16), KC(15)); il[0] -= rclamp(f0 * KC(7), -c, c); }
            { const double c = fmax(fabs(il[1]) * KC(16), KC(15)); il[1] -= rclamp(f1 * KC(7), -c, c); }
            { const double c = fmax(fabs(il[2]) * KC(16), KC(15)); il[2] -= rclamp(f2 * KC(7), -c, c); }
            { const double c = fmax(fabs(il[3]) * KC(16), KC(15)); il[3] -= rclamp(f3 * KC(7), -c, c); }
        }
    }
    bad_out = D.bad;
    return converged;
}
// Rare path: the same iteration with plain IEEE divisions; il in sc[26..29] (in/out), p in sc[30..33].
__device__ __noinline__ bool trm_nr_iter_exact(double* sc, const TrmMats& m, const TrmK& kq) {
    double il[TM], p[TM];
    for (int q = 0; q < TM; q++) { il[q] = sc[26 + q]; p[q] = sc[30 + q]; }
    unsigned bad;
    const bool conv = trm_nr_iter<true>(p, m.k, kq, il, bad);
    for (int q = 0; q < TM; q++) sc[26 + q] = il[q];
    return conv;
}

// process_sample with input 0 (the only way the reference drives it, tremolo.rs:184), gen_tremolo.rs:2353-3116.
__device__ __forceinline__ double trm_step(TrmState& st, const TrmMats& m, const TrmK& kq, TrmDiag* dg, double* sc) {
    const double input = 0.0;
#pragma unroll
    for (int i = 0; i < TN; i++) st.v[i] = st.v[i] + KC(8) - KC(8);
#pragma unroll
    for (int i = 0; i < TM; i++) st.il[i] = st.il[i] + KC(8) - KC(8);
    double rhs[TN];
    for (int i = 0; i < TN; i++) rhs[i] = TRM_RHS_CONST[i];
    const double (*an)[TN] = m.a_neg;
    const double* vp = st.v;
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
    rhs[0] += TRM_N_I[0][0] * st.il[0];
    rhs[0] += TRM_N_I[0][2] * st.il[2];
    rhs[2] += TRM_N_I[2][1] * st.il[1];
    rhs[4] += TRM_N_I[4][0] * st.il[0];
    rhs[4] += TRM_N_I[4][1] * st.il[1];
    rhs[4] += TRM_N_I[4][3] * st.il[3];
    const double input_conductance = kq.input_conductance;
    rhs[0] += (input + st.xin_prev) * input_conductance;
    st.xin_prev = input;

    double v_pred[TN];
#pragma unroll
    for (int i = 0; i < TN; i++) {
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < TN; j++) sum += m.s[i][j] * rhs[j];
        v_pred[i] = sum;
    }
    double p[TM];
    p[0] = TRM_N_V[0][2] * v_pred[2] + TRM_N_V[0][4] * v_pred[4];
    p[1] = TRM_N_V[1][0] * v_pred[0] + TRM_N_V[1][2] * v_pred[2];
    p[2] = TRM_N_V[2][4] * v_pred[4];
    p[3] = TRM_N_V[3][0] * v_pred[0] + TRM_N_V[3][4] * v_pred[4];
    double il[TM];
    for (int i = 0; i < TM; i++) il[i] = 2.0 * st.il[i] - st.ilpp[i];
    uint32_t last = T_MAX_ITER;
    const double (*k)[TM] = m.k;
    for (int iter = 0; iter < T_MAX_ITER; iter++) {
        double nl[TM];
#pragma unroll
        for (int q = 0; q < TM; q++) nl[q] = il[q];
        unsigned bad;
        bool conv = trm_nr_iter<false>(p, k, kq, nl, bad);
        if (bad) {  // an operand left the fast division's validated range: redo the iteration with plain IEEE divisions
#pragma unroll
            for (int q = 0; q < TM; q++) { sc[26 + q] = il[q]; sc[30 + q] = p[q]; }
            conv = trm_nr_iter_exact(sc, m, kq);
#pragma unroll
            for (int q = 0; q < TM; q++) nl[q] = sc[26 + q];
        }
#pragma unroll
        for (int q = 0; q < TM; q++) il[q] = nl[q];
        if (conv) { last = (uint32_t)iter; break; }
    }
    if (dg) dg->hist[last < 15u ? last : 15u]++;
    double v[TN];
#pragma unroll
    for (int i = 0; i < TN; i++) {
        double acc = v_pred[i];
#pragma unroll
        for (int j = 0; j < TM; j++) acc += m.s_ni[i][j] * il[j];
        v[i] = acc;
    }
    if (!(last < (uint32_t)T_MAX_ITER)) {
        if (dg) dg->be_fallback++;
#pragma unroll
        for (int i = 0; i < TN; i++) sc[i] = st.v[i];
#pragma unroll
        for (int i = 0; i < TM; i++) { sc[7 + i] = st.il[i]; sc[11 + i] = st.ilpp[i]; }
        trm_be(input, sc, m, kq);
#pragma unroll
        for (int i = 0; i < TN; i++) v[i] = sc[15 + i];
#pragma unroll
        for (int i = 0; i < TM; i++) il[i] = sc[22 + i];
    }
    bool fin = true;
    for (int i = 0; i < TN; i++) fin = fin && finite64(v[i]);
    if (!fin) {
        for (int i = 0; i < TN; i++) st.v[i] = TRM_DC_OP[i];
        for (int i = 0; i < TM; i++) { st.il[i] = TRM_DC_NL_I[i]; st.ilpp[i] = TRM_DC_NL_I[i]; }
        st.xin_prev = 0.0;
        if (dg) dg->nan_reset++;
        return 4.26480458363572357e0;
    }
    for (int i = 0; i < TN; i++) st.v[i] = v[i];
    for (int i = 0; i < TM; i++) { st.ilpp[i] = st.il[i]; st.il[i] = il[i]; }
    return v[0];
}

}  // namespace owgd
