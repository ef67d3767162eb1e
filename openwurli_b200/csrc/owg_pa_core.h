// The melange-generated 7-BJT class-AB power amplifier with rail sag on a 16-lane tile per instance.
//   solver   crates/openwurli-dsp/src/gen_power_amp.rs:8838 (process_sample: N = 20 nodes, M = 16 junction dimensions of eight Gummel-Poon
//            devices with parasitic-resistance inner solves, dense 16x16 partial-pivoting Newton step, <= 70 iterations, backward-Euler
//            retry with per-device damping), :7880-8150 (device model), :8421 (default state), :12440 (set_sample_rate / rebuild_matrices)
//   adapter  crates/openwurli-dsp/src/power_amp.rs:279-465 (settled state, divergence guard, last-good hold), :65-165 (RailDynamics)
//
// Layout of one instance on its tile (lane l = 0..15):
//   * lane l owns junction dimension l (device l >> 1: even lane = collector current / row, odd lane = base current / row), Jacobian row l
//     of the Newton system, node row l of the 20-node linear part and, for l < 4, node row 16 + l as well;
//   * the 16x16 elimination is row-per-lane: pivot search by butterfly shuffle (largest magnitude, lowest row on ties: the reference's
//     scan order), pivot row broadcast through a double-buffered shared-memory line, every lane updating its own row -- each row sees
//     exactly the reference's operations in the reference's order, so the result is bit-identical to the sequential elimination;
//   * back substitution is inherently serial (row i needs x[i+1..15] in ascending order); lane i forms x[i] and broadcasts it;
//   * vectors that every lane needs in full (v_prev, rhs, v_pred, i_nl) meet in shared memory.
// The file is portable C++: `T` supplies the tile collectives (CUDA: sub-warp shuffles and __syncwarp on the tile's mask; tests: 16
// coroutines in lock-step, tests/pa_tile_emu.cpp), so the same source is checked against the oracle on the CPU and runs on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PA_HD __host__ __device__ __forceinline__
#else
#define PA_HD inline
#endif

#define PA_N 20
#define PA_M 16
#define PA_MAX_ITER 70
#define PA_NDEV 8
#define PA_KS 17  // row stride of the 16-column matrices in shared memory: lane l reads row l, (34 l + 2 j) mod 32 distinct over l
#define PA_SS 21  // row stride of S: (42 l + 2 j) mod 32 distinct over l = 0..15
#define PA_HEADROOM 22.0          // power_amp.rs:287
#define PA_RAIL_V_OPEN 24.5       // power_amp.rs:11-48
#define PA_RAIL_DC_BIAS 22.5
#define PA_RAIL_R_EFF 3.5
#define PA_SPEAKER_LOAD_OHMS 8.0
#define PA_SETTLE_SAMPLES (50 + 44100)  // CircuitState::default() warm-up (gen_power_amp.rs:8490) + compute_settled_state (power_amp.rs:291-296)

// raw device parameters (ow_consts_pa.inc: PA_DEV[d]) ...
enum { PA_IS, PA_VT, PA_BF, PA_BR, PA_NF, PA_NR, PA_ISE, PA_NE, PA_ISC, PA_NC, PA_SIGN, PA_VAF, PA_VAR, PA_IKF, PA_IKR, PA_VCRIT, PA_RB, PA_RC, PA_RE, PA_DEVF };
// ... and the constant sub-expressions of bjt_evaluate, formed once on the host with the same IEEE operations the reference performs per call
enum { PD_IS, PD_SIGN, PD_NF_VT, PD_NR_VT, PD_NE_VT, PD_NC_VT, PD_ISE, PD_ISC, PD_IS_BF, PD_IS_BR, PD_G_FWD, PD_G_REV, PD_G_LBE, PD_G_LBC,
       PD_VAF, PD_VAR, PD_IKF, PD_IKR, PD_Q2F, PD_Q2R, PD_GCC_F, PD_GCC_R, PD_RB, PD_RC, PD_RE, PD_MAX_STEP, PD_VT, PD_VCRIT,
       // device only: reciprocals of the eight model-constant divisors of bjt_evaluate, produced on the GPU by the first half of its own
       // IEEE division sequence (owg_device.cuh recip_prepare) when the model is staged; see PA_DIV_CONST
       PD_R_NF_VT, PD_R_NR_VT, PD_R_NE_VT, PD_R_NC_VT, PD_R_VAR, PD_R_VAF, PD_R_IKF, PD_R_IKR, PD_N };

// a / d[IB] for a model-constant divisor.  Host: the plain IEEE division.  Device (owg_poweramp.cuh): the second half of the compiler's own
// division sequence on the staged reciprocal d[IR] -- the same instructions `a / b` expands to, so the same bits, at a third of the cost.
#if !defined(__CUDA_ARCH__)
#define PA_DIV_CONST(a, d, IB, IR) ((a) / (d)[IB])
#elif !defined(PA_DIV_CONST)
#error "owg_poweramp.cuh defines PA_DIV_CONST for the device pass"
#endif

struct PaModel {  // one per sample rate; built by owg::pa_build_model (host_setup.cpp), read-only on the device
    double s[PA_N][PA_N], k[PA_M][PA_M], s_ni[PA_N][PA_M];
    double s_be[PA_N][PA_N], k_be[PA_M][PA_M], s_ni_be[PA_N][PA_M];
    double a_neg_be[PA_N][PA_N];
    double n_i[PA_N][PA_M], n_v[PA_M][PA_N];
    double rhs_const[PA_N], rhs_const_be[PA_N], dc_op[PA_N], dc_nl_i[PA_M];
    double dev[PA_NDEV][PD_N];
    double a_val[PA_N][5];  // structural non-zeros of a_neg (= alpha C), row-major, the terms gen_power_amp.rs:8842-8895 spells out
    double nv_val[PA_M][2]; // the two non-zeros of each N_V row (:8908-8923)
    int32_t a_col[PA_N][5], a_cnt[PA_N], nv_col[PA_M][2];
    double input_conductance, dc_block_r, nan_out, out_node;
    double alpha_attack, alpha_release, alpha_i_avg;  // RailDynamics::set_sample_rate (power_amp.rs:103-108)
    double sample_rate;
    int32_t rerated, _pad;  // set_sample_rate ran (|sr - 88200| >= 0.5): matrices rebuilt, DC-blocker history cleared (:12440-12470)
};
struct PaSettled {  // the cached settled CircuitState (power_amp.rs:289-297): what `reset()` and the constructor clone
    double v_prev[PA_N], i_nl_prev[PA_M], i_nl_pp[PA_M], dc_x, dc_y;
};
struct PaShared {  // the hot part of the model, padded for conflict-free row-per-lane reads
    double s[PA_N * PA_SS], k[PA_M * PA_KS], s_ni[PA_N * PA_KS], k_be[PA_M * PA_KS];
    double rhs_const[PA_N], a_val[PA_N][5], nv_val[PA_M][2], dev[PA_NDEV][PD_N];
    int32_t a_col[PA_N][5], a_cnt[PA_N], nv_col[PA_M][2];
};
struct PaScratch {  // per tile
    double vp[PA_N];       // flushed v_prev of this sample (kept for the backward-Euler retry)
    double w[PA_N];        // rhs, then v_pred
    double ci[PA_M];       // i_nl / i_trial gather
    double ip[PA_M];       // flushed i_nl_prev
    double a[PA_M][PA_M + 1];  // the Newton system, one row per lane (column 16 = right-hand side); odd stride: conflict-free rows
    double x[PA_M];            // its solution
    int32_t sing, _pad;        // back substitution met a zero pivot
};
struct PaLane {  // per-lane registers carried from sample to sample
    double v0, v1;          // v_prev[l], v_prev[16 + l] (l < 4)
    double ip, ipp;         // i_nl_prev[l], i_nl_prev_prev[l]
    double dc_x, dc_y;      // DC blocker (replicated in every lane)
    double rail_pos, rail_neg, iavg_pos, iavg_neg, last_good;  // adapter (replicated)
    uint32_t last_iters, resets, be_fallbacks, nan_resets;
};

PA_HD bool pa_finite(double x) { return fabs(x) <= 1.7976931348623157e308; }
PA_HD double pa_clamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }  // f64::clamp (NaN stays NaN)
PA_HD int64_t pa_bits(double x) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PA_HD double pa_from_bits(int64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
// gen_power_amp.rs fast_exp (:7745-7770): range-reduced degree-5 polynomial, clamped to +-40
PA_HD double pa_exp(double x) {
    x = pa_clamp(x, -40.0, 40.0);
    const double SHIFT = 6755399441055744.0;
    const double z = x * 1.4426950408889634 + SHIFT;
    const int64_t n_i64 = pa_bits(z) - pa_bits(SHIFT);
    const double n = (double)n_i64;
    const double f = (x - n * 0.6931471803691238) - n * 1.9082149292705877e-10;
    const double p = 1.0 + f * (1.0 + f * (0.5 + f * (0.16666666666666607 + f * (0.04166666666665876 + f * 0.008333333333492337))));
    return p * pa_from_bits((int64_t)((uint64_t)(1023 + n_i64) << 52));
}
// pnjlim (:7790-7805); the logarithm is libm's (glibc on the reference side, CUDA's on the device: <= 1 ulp apart, inside the limiter only)
PA_HD double pa_pnjlim(double vnew, double vold, double vt, double vcrit) {
    if (vnew > vcrit && fabs(vnew - vold) > vt + vt) {
        if (vold >= 0.0) {
            const double arg = 1.0 + (vnew - vold) / vt;
            return arg > 0.0 ? vold + vt * log(arg) : vcrit;
        }
        return vt * log(vnew / vt);
    }
    return vnew;
}

// constant sub-expressions of one device (host, once per model)
inline void pa_dev_derive(const double* r, double* d) {
    const double is = r[PA_IS], vt = r[PA_VT], nf_vt = r[PA_NF] * vt, nr_vt = r[PA_NR] * vt;
    d[PD_IS] = is; d[PD_SIGN] = r[PA_SIGN]; d[PD_NF_VT] = nf_vt; d[PD_NR_VT] = nr_vt;
    d[PD_NE_VT] = r[PA_NE] * vt; d[PD_NC_VT] = r[PA_NC] * vt; d[PD_ISE] = r[PA_ISE]; d[PD_ISC] = r[PA_ISC];
    d[PD_IS_BF] = is / r[PA_BF]; d[PD_IS_BR] = is / r[PA_BR];
    d[PD_G_FWD] = is / (r[PA_BF] * nf_vt); d[PD_G_REV] = is / (r[PA_BR] * nr_vt);
    d[PD_G_LBE] = r[PA_ISE] / (r[PA_NE] * vt); d[PD_G_LBC] = r[PA_ISC] / (r[PA_NC] * vt);
    d[PD_VAF] = r[PA_VAF]; d[PD_VAR] = r[PA_VAR]; d[PD_IKF] = r[PA_IKF]; d[PD_IKR] = r[PA_IKR];
    d[PD_Q2F] = is / (nf_vt * r[PA_IKF]); d[PD_Q2R] = is / (nr_vt * r[PA_IKR]);
    d[PD_GCC_F] = is / nf_vt; d[PD_GCC_R] = (-is) / nr_vt;
    d[PD_RB] = r[PA_RB]; d[PD_RC] = r[PA_RC]; d[PD_RE] = r[PA_RE]; d[PD_MAX_STEP] = 4.0 * vt; d[PD_VT] = vt; d[PD_VCRIT] = r[PA_VCRIT];
    for (int k = PD_R_NF_VT; k < PD_N; k++) d[k] = 0.0;  // filled on the device when the model is staged
}

// bjt_evaluate, Gummel-Poon branch (gen_power_amp.rs:7990-8060; every device of this circuit has USE_GP = true)
PA_HD void pa_bjt_evaluate(double vbe, double vbc, const double* d, double& ic, double& ib, double* jac) {
    const double sign = d[PD_SIGN], is = d[PD_IS];
    const double vbe_eff = sign * vbe, vbc_eff = sign * vbc;
    const double exp_be = pa_exp(PA_DIV_CONST(vbe_eff, d, PD_NF_VT, PD_R_NF_VT)), exp_bc = pa_exp(PA_DIV_CONST(vbc_eff, d, PD_NR_VT, PD_R_NR_VT));
    const bool lbe = d[PD_ISE] > 0.0, lbc = d[PD_ISC] > 0.0;
    const double exp_be_leak = lbe ? pa_exp(PA_DIV_CONST(vbe_eff, d, PD_NE_VT, PD_R_NE_VT)) : 0.0;
    const double exp_bc_leak = lbc ? pa_exp(PA_DIV_CONST(vbc_eff, d, PD_NC_VT, PD_R_NC_VT)) : 0.0;
    const double i_cc = is * (exp_be - exp_bc);
    const double ib_fwd = d[PD_IS_BF] * (exp_be - 1.0), ib_rev = d[PD_IS_BR] * (exp_bc - 1.0);
    const double ib_leak_be = lbe ? d[PD_ISE] * (exp_be_leak - 1.0) : 0.0;
    const double ib_leak_bc = lbc ? d[PD_ISC] * (exp_bc_leak - 1.0) : 0.0;
    const double dib_fwd_dvbe = d[PD_G_FWD] * exp_be, dib_rev_dvbc = d[PD_G_REV] * exp_bc;
    const double dib_leak_dvbe = lbe ? d[PD_G_LBE] * exp_be_leak : 0.0;
    const double dib_leak_dvbc = lbc ? d[PD_G_LBC] * exp_bc_leak : 0.0;
    const double q1_denom = (1.0 - PA_DIV_CONST(vbe_eff, d, PD_VAR, PD_R_VAR)) - PA_DIV_CONST(vbc_eff, d, PD_VAF, PD_R_VAF);
    double q1 = 1.0, dq1_dvbe = 0.0, dq1_dvbc = 0.0;
    if (!(q1_denom <= 0.0 || fabs(q1_denom) < 1e-30)) {
        q1 = 1.0 / q1_denom;
        dq1_dvbe = PA_DIV_CONST(q1 * q1, d, PD_VAR, PD_R_VAR);
        dq1_dvbc = PA_DIV_CONST(q1 * q1, d, PD_VAF, PD_R_VAF);
    }
    const double cbe = is * (exp_be - 1.0), cbc = is * (exp_bc - 1.0);
    const double q2 = PA_DIV_CONST(cbe, d, PD_IKF, PD_R_IKF) + PA_DIV_CONST(cbc, d, PD_IKR, PD_R_IKR);
    const double dq2_dvbe = d[PD_Q2F] * exp_be, dq2_dvbc = d[PD_Q2R] * exp_bc;
    const double disc = fmax(1.0 + 4.0 * q2, 0.0);
    const double dd = sqrt(disc);
    const double dd_dvbe = dd > 1e-15 ? (2.0 * dq2_dvbe) / dd : 0.0;
    const double dd_dvbc = dd > 1e-15 ? (2.0 * dq2_dvbc) / dd : 0.0;
    const double qb = (q1 * (1.0 + dd)) / 2.0;
    const double dqb_dvbe = (dq1_dvbe * (1.0 + dd)) / 2.0 + (q1 * dd_dvbe) / 2.0;
    const double dqb_dvbc = (dq1_dvbc * (1.0 + dd)) / 2.0 + (q1 * dd_dvbc) / 2.0;
    ic = sign * (i_cc / qb - d[PD_IS_BR] * (exp_bc - 1.0));
    ib = sign * (((ib_fwd + ib_rev) + ib_leak_be) + ib_leak_bc);
    const double dicc_dvbe = d[PD_GCC_F] * exp_be, dicc_dvbc = d[PD_GCC_R] * exp_bc;
    const double qb2 = fmax(qb * qb, 1e-30);
    const double quotient_dvbe = (dicc_dvbe * qb - i_cc * dqb_dvbe) / qb2;
    const double quotient_dvbc = (dicc_dvbc * qb - i_cc * dqb_dvbc) / qb2;
    const double d_bc_term_dvbc = d[PD_G_REV] * exp_bc;
    jac[0] = quotient_dvbe; jac[1] = quotient_dvbc - d_bc_term_dvbc;
    jac[2] = dib_fwd_dvbe + dib_leak_dvbe; jac[3] = dib_rev_dvbc + dib_leak_dvbc;
}

// bjt_with_parasitics (gen_power_amp.rs:8062-8150): inner 2x2 Newton on the internal junction voltages behind RB / RC / RE (<= 15 updates),
// then the terminal-voltage Jacobian through the inverse of the inner Jacobian.  The reference re-evaluates the device after leaving the
// loop; when it left by convergence or a vanishing determinant that evaluation repeats the last one bit for bit, so one call site serves
// both (the 16th pass is the evaluation after 15 updates).
PA_HD void pa_bjt_with_parasitics(double vbe_ext, double vbc_ext, const double* d, double& ic, double& ib, double* jac) {
    const double rb = d[PD_RB], rc = d[PD_RC], re = d[PD_RE], max_step = d[PD_MAX_STEP];
    double vbe_int = vbe_ext, vbc_int = vbc_ext;
    double j11, j12, j21, j22, det;
    for (int it = 0;; it++) {
        pa_bjt_evaluate(vbe_int, vbc_int, d, ic, ib, jac);
        j11 = (1.0 + jac[2] * rb) + (jac[0] + jac[2]) * re;
        j12 = jac[3] * rb + (jac[1] + jac[3]) * re;
        j21 = jac[2] * rb - jac[0] * rc;
        j22 = (1.0 + jac[3] * rb) - jac[1] * rc;
        det = j11 * j22 - j12 * j21;
        if (it == 15) break;
        const double f1 = ((vbe_int - vbe_ext) + ib * rb) + (ic + ib) * re;
        const double f2 = ((vbc_int - vbc_ext) + ib * rb) - ic * rc;
        if (fabs(f1) < 1e-10 && fabs(f2) < 1e-10) break;
        if (fabs(det) < 1e-30) break;
        const double inv_det = 1.0 / det;
        const double dvbe = (j22 * f1 - j12 * f2) * inv_det;
        const double dvbc = (j11 * f2 - j21 * f1) * inv_det;
        vbe_int -= pa_clamp(dvbe, -max_step, max_step);
        vbc_int -= pa_clamp(dvbc, -max_step, max_step);
    }
    if (fabs(det) < 1e-30) return;
    const double inv_det = 1.0 / det;
    const double fi11 = j22 * inv_det, fi12 = (-j12) * inv_det, fi21 = (-j21) * inv_det, fi22 = j11 * inv_det;
    const double e0 = jac[0] * fi11 + jac[1] * fi21, e1 = jac[0] * fi12 + jac[1] * fi22;
    const double e2 = jac[2] * fi11 + jac[3] * fi21, e3 = jac[2] * fi12 + jac[3] * fi22;
    jac[0] = e0; jac[1] = e1; jac[2] = e2; jac[3] = e3;
}

// stage the hot part of the model into the padded shared layout (all threads of the CTA; caller synchronises)
PA_HD void pa_stage_shared(const PaModel& m, PaShared& sh, int tid, int nthreads) {
    for (int i = tid; i < PA_N * PA_N; i += nthreads) sh.s[(i / PA_N) * PA_SS + i % PA_N] = m.s[i / PA_N][i % PA_N];
    for (int i = tid; i < PA_M * PA_M; i += nthreads) {
        sh.k[(i / PA_M) * PA_KS + i % PA_M] = m.k[i / PA_M][i % PA_M];
        sh.k_be[(i / PA_M) * PA_KS + i % PA_M] = m.k_be[i / PA_M][i % PA_M];
    }
    for (int i = tid; i < PA_N * PA_M; i += nthreads) sh.s_ni[(i / PA_M) * PA_KS + i % PA_M] = m.s_ni[i / PA_M][i % PA_M];
    for (int i = tid; i < PA_N; i += nthreads) {
        sh.rhs_const[i] = m.rhs_const[i]; sh.a_cnt[i] = m.a_cnt[i];
        for (int c = 0; c < 5; c++) { sh.a_val[i][c] = m.a_val[i][c]; sh.a_col[i][c] = m.a_col[i][c]; }
    }
    for (int i = tid; i < PA_M; i += nthreads)
        for (int c = 0; c < 2; c++) { sh.nv_val[i][c] = m.nv_val[i][c]; sh.nv_col[i][c] = m.nv_col[i][c]; }
    for (int i = tid; i < PA_NDEV * PD_N; i += nthreads) sh.dev[i / PD_N][i % PD_N] = m.dev[i / PD_N][i % PD_N];
}

// ---- the 16x16 Newton linear solve, row per lane (gen_power_amp.rs:9300-9345 / :10870-10915) -------------------------------------------
// sc.a[l][0..15 | 16]: lane l's row and right-hand side, written by the caller.  Returns false when the reference's elimination or back
// substitution reports a singular system; otherwise sc.x[0..15] holds the solution.  Rolled loops over rows in shared memory (a register
// copy of the row needs every index static, i.e. the whole elimination unrolled: 50 k instructions, instruction-fetch bound).  A row swap
// exchanges the two lanes' row indices, not the rows.  The two tiles of a warp run this in lock-step: there is no early exit (a tile that
// has found its system singular keeps executing the collectives on values nobody reads).
template <class T>
PA_HD bool pa_tile_solve16(const T& t, PaScratch& sc) {
    const int l = t.lane;
    int myrow = l;  // the storage row that currently holds logical row l
    bool singular = false;
    if (l == 0) sc.sing = 0;
#pragma unroll 1
    for (int col = 0; col < PA_M; col++) {
        // pivot: largest |a[row][col]| over rows >= col, the lowest row on ties (the reference scans upwards with a strict `>`;
        // a NaN in a lower row is never selected, a NaN on the diagonal stays the pivot: it is entered as +inf and recognised below)
        // The magnitudes are non-negative, so their bit patterns order like their values: a 64-bit maximum in two 32-bit reductions
        // (redux.sync on the device), then the lowest lane that holds it.
        const double mine = fabs(sc.a[myrow][col]);
        const bool eligible = l >= col && (mine == mine || l == col);
        const int64_t key = !eligible ? 0 : (mine == mine ? pa_bits(mine) : pa_bits(INFINITY));
        const uint32_t hi = (uint32_t)((uint64_t)key >> 32), lo = (uint32_t)(uint64_t)key;
        const uint32_t m_hi = t.max_u32(hi);
        const bool cand = eligible && hi == m_hi;
        const uint32_t m_lo = t.max_u32(cand ? lo : 0u);
        const int mi = t.first_lane(cand && lo == m_lo);
        const double mv = pa_from_bits((int64_t)(((uint64_t)m_hi << 32) | (uint64_t)m_lo));
        const double diag_abs = t.shfl(mine, col);
        const double max_val = mi == col ? diag_abs : mv;  // the reference's max_val (NaN when the diagonal is NaN and nothing beats it)
        singular = singular || max_val < 1e-15;
        const bool sw = !singular && mi != col;
        const int other = t.shfl_i(myrow, l == col ? mi : (l == mi ? col : l));
        if (sw) myrow = other;  // rows col and mi change hands
        const int prow = t.shfl_i(myrow, col);
        t.sync();  // every row is up to date (the caller's assembly, the previous column's updates)
        if (l > col) {
            const double* pr = sc.a[prow];
            double* ar = sc.a[myrow];
            const double factor = ar[col] / pr[col];
#pragma unroll 4
            for (int j = col + 1; j <= PA_M; j++) ar[j] -= factor * pr[j];
        }
    }
#pragma unroll 1
    for (int i = PA_M - 1; i >= 0; i--) {
        t.sync();  // x[i + 1 ..] (and, the first time, the last column's updates) are visible
        if (l == i) {
            const double* ar = sc.a[myrow];
            double sum = ar[PA_M];
#pragma unroll 4
            for (int j = i + 1; j < PA_M; j++) sum -= ar[j] * sc.x[j];
            if (fabs(ar[i]) < 1e-15) sc.sing = 1;
            sc.x[i] = sum / ar[i];
        }
    }
    t.sync();
    return !(singular || sc.sing != 0);
}

// One Newton loop of process_sample on the tile.  be = false: the first loop (:8930-10560, one global step scale); be = true: the
// backward-Euler retry (:10600-12260, one scale per device).  p = this lane's p[l]; i_nl = this lane's iterate (in/out).  Returns the
// iteration index at which the loop converged, or PA_MAX_ITER.  One body serves both loops and both tiles of a warp: every collective is
// executed by every lane in every iteration (a tile that is done, or not concerned, is `active == false` and commits nothing), so the
// CUDA build uses full-warp shuffles / votes and the kernel holds a single copy of the junction model and the elimination.
template <class T>
PA_HD uint32_t pa_tile_newton(const T& t, const PaShared& sh, PaScratch& sc, const bool be, bool active, const double p, double& i_nl) {
    const int l = t.lane;
    const double* dv = sh.dev[l >> 1];
    const double* kmat = be ? sh.k_be : sh.k;
    const double* krow = kmat + l * PA_KS;
    const double* kA = kmat + (l & ~1) * PA_KS;  // K rows 2d and 2d + 1 of this lane's device
    const double* kB = kA + PA_KS;
    const bool odd = (l & 1) != 0;
    uint32_t result = PA_MAX_ITER;
#pragma unroll 1
    for (uint32_t iter = 0; iter < PA_MAX_ITER; iter++) {
        if (!t.warp_any(active)) break;
        t.sync();  // the previous iteration's readers of ci / x are done
        sc.ci[l] = i_nl;
        t.sync();
        double v_d = p;
#pragma unroll 4
        for (int j = 0; j < PA_M; j++) v_d += krow[j] * sc.ci[j];
        const double vbe = t.shfl(v_d, l & ~1), vbc = t.shfl(v_d, l | 1);
        double ic, ib, jac[4];
        pa_bjt_with_parasitics(vbe, vbc, dv, ic, ib, jac);
        const double i_dev = odd ? ib : ic, jA = odd ? jac[2] : jac[0], jB = odd ? jac[3] : jac[1];
        const double f = i_nl - i_dev;
        {
            double* ar = sc.a[l];
#pragma unroll 4
            for (int c = 0; c < PA_M; c++) ar[c] = ((c == l ? 1.0 : 0.0) - jA * kA[c]) - jB * kB[c];
            ar[PA_M] = f;
        }
        const bool ok = pa_tile_solve16(t, sc);
        const double delta = sc.x[l];
        sc.ci[l] = i_nl - delta;  // i_trial (first loop); the readers of ci above are past the solve's barriers
        t.sync();
        // the junction-voltage step this lane limits and tests, and its limiter ratio
        double dvj, r = INFINITY, alpha = 1.0;
        bool limited = false;
        if (!be) {
            double v_trial = p;
#pragma unroll 4
            for (int j = 0; j < PA_M; j++) v_trial += krow[j] * sc.ci[j];
            dvj = v_trial - v_d;
            const double v_lim = fabs(dvj) > 1e-4 ? pa_pnjlim(v_trial, v_d, dv[PD_VT], dv[PD_VCRIT]) : v_trial;
            const double dv_lim = v_lim - v_d;
            if (fabs(dvj) > 1e-15) r = dvj * dv_lim < 0.0 ? 0.0 : pa_clamp(dv_lim / dvj, 0.0, 1.0);
            if (!(r == r)) r = INFINITY;  // a NaN ratio never lowers global_alpha (`r < global_alpha` is false)
        } else {
            double acc = krow[0] * sc.x[0];  // dv = -(K_be[l][0] delta0 + ... + K_be[l][15] delta15)
#pragma unroll 4
            for (int j = 1; j < PA_M; j++) acc += krow[j] * sc.x[j];
            dvj = -acc;
            if (fabs(dvj) > 1e-4) {
                const double v_lim = pa_pnjlim(v_d + dvj, v_d, dv[PD_VT], dv[PD_VCRIT]);
                const double ratio = fmax((v_lim - v_d) / dvj, 0.01);
                if (ratio < alpha) { alpha = ratio; limited = ratio < 1.0; }
            }
        }
        double rmin = r;
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
            const double o = t.shfl_xor(rmin, off);
            rmin = o < rmin ? o : rmin;
        }
        const bool any_lim_be = t.any(limited);
        const double pair = t.shfl_xor(alpha, 1);
        bool any_limited;
        double scale;  // what multiplies this lane's delta
        if (!be) {
            any_limited = rmin < 1.0;
            scale = any_limited ? rmin : 1.0;  // global_alpha
        } else {
            any_limited = any_lim_be;
            scale = fmin(alpha, pair);  // one factor per device: alpha[2d] = alpha[2d + 1] = min of the pair
        }
        double mx = fabs(dvj * scale);  // the largest junction step AFTER the limiter's scaling
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) mx = fmax(mx, t.shfl_xor(mx, off));
        if (mx > 3.5) {
            scale *= fmax(3.5 / mx, 0.1);
            if (!be) any_limited = true;
        }
        const double step = dvj * scale;
        const double thr = 1e-3 * fmax(fabs(v_d), fabs(v_d + step)) + 1e-6;
        const bool not_conv = t.any(fabs(step) > thr);
        if (active) {
            if (ok) {
                i_nl -= scale * delta;
                if (!any_limited && !not_conv) { result = iter; active = false; }
            } else if (be) {  // singular Jacobian: damped fixed-point step on the residual
                i_nl -= pa_clamp(f * 0.5, -0.01, 0.01);
            } else {
                const double cl = fmax(fabs(i_nl) * 0.1, 0.01);
                i_nl -= pa_clamp(f * 0.5, -cl, cl);
            }
        }
    }
    return result;
}

PA_HD double pa_rhs_row(const PaShared& sh, const PaScratch& sc, int i) {
    double sum = sh.rhs_const[i];
    for (int c = 0; c < sh.a_cnt[i]; c++) sum += sh.a_val[i][c] * sc.vp[sh.a_col[i][c]];
    return sum;
}

// gen_power_amp::process_sample (:8838-12345) for the tile's instance.  Returns output[0] (every lane holds it).
template <class T>
PA_HD double pa_tile_process_sample(const T& t, const PaModel& m, const PaShared& sh, PaScratch& sc, PaLane& s, double input, double off_pos,
                                    double off_neg) {
    const int l = t.lane;
    const bool two = l < 4;  // this lane also owns node row 16 + l
    const double in = pa_finite(input) ? pa_clamp(input, -100.0, 100.0) : 0.0;
    s.v0 = (s.v0 + 1e-25) - 1e-25;
    if (two) s.v1 = (s.v1 + 1e-25) - 1e-25;
    s.ip = (s.ip + 1e-25) - 1e-25;
    t.sync();  // the previous sample's readers of the scratch vectors are done
    sc.vp[l] = s.v0;
    if (two) sc.vp[16 + l] = s.v1;
    sc.ip[l] = s.ip;
    t.sync();
    double r0 = pa_rhs_row(sh, sc, l), r1 = 0.0;
    if (l == 0) r0 += in * m.input_conductance;
    if (two) {
        r1 = pa_rhs_row(sh, sc, 16 + l);
        if (l == 2) r1 += off_pos;
        if (l == 3) r1 += off_neg;
    }
    sc.w[l] = r0;
    if (two) sc.w[16 + l] = r1;
    t.sync();
    double vp0 = 0.0, vp1 = 0.0;
    {
        const double* srow = sh.s + l * PA_SS;
#pragma unroll
        for (int j = 0; j < PA_N; j++) vp0 += srow[j] * sc.w[j];
        if (two) {
            const double* srow1 = sh.s + (16 + l) * PA_SS;
#pragma unroll
            for (int j = 0; j < PA_N; j++) vp1 += srow1[j] * sc.w[j];
        }
    }
    t.sync();  // every lane has read rhs
    sc.w[l] = vp0;
    if (two) sc.w[16 + l] = vp1;
    t.sync();
    double p = sh.nv_val[l][0] * sc.w[sh.nv_col[l][0]] + sh.nv_val[l][1] * sc.w[sh.nv_col[l][1]];
    double i_nl = 2.0 * s.ip - s.ipp;
    uint32_t iters = PA_MAX_ITER;
    bool be = false, act = true;  // this tile's loop flavour and whether it takes part
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {  // one call site: the kernel holds one copy of the Newton body
        const uint32_t it = pa_tile_newton(t, sh, sc, be, act, p, i_nl);
        if (act) iters = it;
        if (phase == 1) break;
        const bool need_be = iters >= PA_MAX_ITER;
        if (!t.warp_any(need_be)) break;
        // ---- backward-Euler retry (:10565-12275): dense products in the reference's order; cold tables come from global memory.
        // The warp walks through it together; only a tile that needs it computes and commits ----
        double rb0 = 0.0, rb1 = 0.0;
        if (need_be) {
            s.be_fallbacks++;
            rb0 = m.rhs_const_be[l];
            for (int j = 0; j < PA_N; j++) rb0 += m.a_neg_be[l][j] * sc.vp[j];
            for (int j = 0; j < PA_M; j++) rb0 += m.n_i[l][j] * sc.ip[j];
            if (l == 0) rb0 += in * m.input_conductance;
            if (two) {
                rb1 = m.rhs_const_be[16 + l];
                for (int j = 0; j < PA_N; j++) rb1 += m.a_neg_be[16 + l][j] * sc.vp[j];
                for (int j = 0; j < PA_M; j++) rb1 += m.n_i[16 + l][j] * sc.ip[j];
            }
        }
        t.sync();
        if (need_be) {
            sc.w[l] = rb0;
            if (two) sc.w[16 + l] = rb1;
        }
        t.sync();
        if (need_be) {
            vp0 = 0.0; vp1 = 0.0;
            for (int j = 0; j < PA_N; j++) vp0 += m.s_be[l][j] * sc.w[j];
            if (two) for (int j = 0; j < PA_N; j++) vp1 += m.s_be[16 + l][j] * sc.w[j];
        }
        t.sync();
        if (need_be) {
            sc.w[l] = vp0;
            if (two) sc.w[16 + l] = vp1;
        }
        t.sync();
        if (need_be) {
            p = 0.0;
            for (int j = 0; j < PA_N; j++) p += m.n_v[l][j] * sc.w[j];
            i_nl = 2.0 * s.ip - s.ipp;
        }
        be = need_be; act = need_be;
    }
    const bool need_be = be;
    s.last_iters = iters;
    t.sync();
    sc.ci[l] = i_nl;
    t.sync();
    double v0 = vp0, v1 = vp1;
    if (!need_be) {
        const double* sni = sh.s_ni;
#pragma unroll
        for (int j = 0; j < PA_M; j++) v0 += sni[l * PA_KS + j] * sc.ci[j];
        if (two) {
#pragma unroll
            for (int j = 0; j < PA_M; j++) v1 += sni[(16 + l) * PA_KS + j] * sc.ci[j];
        }
    } else {
        for (int j = 0; j < PA_M; j++) v0 += m.s_ni_be[l][j] * sc.ci[j];
        if (two) for (int j = 0; j < PA_M; j++) v1 += m.s_ni_be[16 + l][j] * sc.ci[j];
    }
    const bool bad = t.any(!pa_finite(v0) || (two && !pa_finite(v1)));  // :12296-12310
    const double raw_out = t.shfl(v0, 8);  // OUTPUT_NODES = [8], OUTPUT_SCALES = [1.0]
    if (bad) {
        s.v0 = m.dc_op[l];
        if (two) s.v1 = m.dc_op[16 + l];
        s.ip = m.dc_nl_i[l]; s.ipp = m.dc_nl_i[l];
        s.dc_x = 0.0; s.dc_y = 0.0;
        s.nan_resets++;
        return m.nan_out;
    }
    s.v0 = v0;
    if (two) s.v1 = v1;
    s.ipp = s.ip;
    s.ip = i_nl;
    const double dc_blocked = (raw_out - s.dc_x) + m.dc_block_r * s.dc_y;
    s.dc_x = raw_out;
    s.dc_y = dc_blocked;
    return pa_clamp(dc_blocked, -30.0, 30.0);
}

// init_state (power_amp.rs:299-306): clone the settled state; set_sample_rate (when it runs) clears the DC blocker's history
PA_HD void pa_lane_load_settled(const PaModel& m, const PaSettled& st, int l, PaLane& s) {
    s.v0 = st.v_prev[l];
    s.v1 = l < 4 ? st.v_prev[16 + l] : 0.0;
    s.ip = st.i_nl_prev[l]; s.ipp = st.i_nl_pp[l];
    s.dc_x = m.rerated ? 0.0 : st.dc_x; s.dc_y = m.rerated ? 0.0 : st.dc_y;
}

struct PaNoPost { PA_HD double operator()(double v) { return v; } };
// PowerAmp::process (power_amp.rs:373-436) over a row, 16 samples per block (lane l loads x[t0 + l] and stores y[t0 + l]: coalesced).
// vol: the row is attenuated as (x * vol) * vol before the amplifier (chain B's audio-taper volume, main.rs:489; 1.0 = plain process()).
// bypass: `--no-poweramp` (the attenuated sample goes straight to `post`).  post: per-sample output stage after the amplifier (chain B: the
// speaker), run redundantly by every lane on the replicated amplifier output.
// settle != nullptr: the raw solver on silence from CircuitState::default()'s initial state for n samples, final state -> *settle
// (compute_settled_state); no adapter.
// The two tiles of a warp walk through the samples together (n_steps = the longer of their rows; `valid` = false for a tile that only keeps
// its warp company): one call site of process_sample, every collective executed by every lane.
template <class T, class Post>
PA_HD void pa_tile_render(const T& t, const PaModel& m, const PaShared& sh, PaScratch& sc, const PaSettled* settled, const double* x, double* y,
                          int64_t n, int64_t n_steps, bool valid, double vol, bool rail_sag, bool bypass, PaSettled* settle, double* rails_out,
                          uint32_t* counters_out, Post& post) {
    const int l = t.lane;
    PaLane s;
    s.rail_pos = PA_RAIL_DC_BIAS; s.rail_neg = PA_RAIL_DC_BIAS; s.iavg_pos = 0.0; s.iavg_neg = 0.0; s.last_good = 0.0;
    s.last_iters = 0; s.resets = 0; s.be_fallbacks = 0; s.nan_resets = 0;
    if (settle) {  // gen_power_amp.rs:8421-8490
        s.v0 = m.dc_op[l]; s.v1 = l < 4 ? m.dc_op[16 + l] : 0.0;
        s.ip = m.dc_nl_i[l]; s.ipp = m.dc_nl_i[l];
        s.dc_x = m.nan_out; s.dc_y = 0.0;
    } else pa_lane_load_settled(m, *settled, l, s);
    const bool adapter = settle == nullptr;
    // what the caller is told about the row's end (a tile that keeps its longer neighbour company runs past its own last sample)
    double end_rail_pos = s.rail_pos, end_rail_neg = s.rail_neg;
    uint32_t end_resets = 0, end_be = 0, end_nan = 0, end_iters = 0;
    for (int64_t t0 = 0; t0 < n_steps; t0 += 16) {
        const int nb = n_steps - t0 < 16 ? (int)(n_steps - t0) : 16;
        const bool mine = valid && t0 + l < n;
        double xin = (adapter && mine) ? x[t0 + l] : 0.0, yout = 0.0;
        for (int k = 0; k < nb; k++) {
            const double input = adapter ? (t.shfl(xin, k) * vol) * vol : 0.0;
            const double off_pos = (adapter && rail_sag) ? s.rail_pos - PA_RAIL_DC_BIAS : 0.0;
            const double off_neg = (adapter && rail_sag) ? s.rail_neg - PA_RAIL_DC_BIAS : 0.0;
            const double raw = pa_tile_process_sample(t, m, sh, sc, s, input, off_pos, off_neg);
            const bool insane = t.any(!pa_finite(s.v0) || fabs(s.v0) > 100.0 || (l < 4 && (!pa_finite(s.v1) || fabs(s.v1) > 100.0)));
            double out = raw;
            if (adapter) {
                const double result = raw / PA_HEADROOM;
                const bool nr_failed = s.last_iters >= PA_MAX_ITER - 1;
                if (!pa_finite(result) || nr_failed || insane) {  // divergence guard: re-clone the settled state, hold the last good output
                    pa_lane_load_settled(m, *settled, l, s);
                    s.rail_pos = PA_RAIL_DC_BIAS; s.rail_neg = PA_RAIL_DC_BIAS; s.iavg_pos = 0.0; s.iavg_neg = 0.0;
                    s.resets++;
                    out = s.last_good;
                } else {
                    out = pa_clamp(result, -1.0, 1.0);
                    s.last_good = out;
                    if (rail_sag) {  // RailDynamics::step (power_amp.rs:128-155)
                        const double i_pos = fmax(raw / PA_SPEAKER_LOAD_OHMS, 0.0), i_neg = fmax(-raw / PA_SPEAKER_LOAD_OHMS, 0.0);
                        s.iavg_pos += m.alpha_i_avg * (i_pos - s.iavg_pos);
                        s.iavg_neg += m.alpha_i_avg * (i_neg - s.iavg_neg);
                        const double target_pos = PA_RAIL_V_OPEN - s.iavg_pos * PA_RAIL_R_EFF;
                        const double target_neg = PA_RAIL_V_OPEN - s.iavg_neg * PA_RAIL_R_EFF;
                        const double ap = target_pos < s.rail_pos ? m.alpha_attack : m.alpha_release;
                        const double an = target_neg < s.rail_neg ? m.alpha_attack : m.alpha_release;
                        s.rail_pos += ap * (target_pos - s.rail_pos);
                        s.rail_neg += an * (target_neg - s.rail_neg);
                    }
                }
                if (bypass) out = input;
                out = post(out);
            }
            if (l == k) yout = out;
            if (t0 + k + 1 == n) {
                end_rail_pos = s.rail_pos; end_rail_neg = s.rail_neg;
                end_resets = s.resets; end_be = s.be_fallbacks; end_nan = s.nan_resets; end_iters = s.last_iters;
            }
        }
        if (adapter && y && mine) y[t0 + l] = yout;
    }
    if (settle) {
        settle->v_prev[l] = s.v0;
        if (l < 4) settle->v_prev[16 + l] = s.v1;
        settle->i_nl_prev[l] = s.ip; settle->i_nl_pp[l] = s.ipp;
        if (l == 0) { settle->dc_x = s.dc_x; settle->dc_y = s.dc_y; }
    }
    if (l == 0 && valid) {
        if (rails_out) { rails_out[0] = rail_sag ? end_rail_pos : PA_RAIL_DC_BIAS; rails_out[1] = rail_sag ? end_rail_neg : PA_RAIL_DC_BIAS; }
        if (counters_out) { counters_out[0] = end_resets; counters_out[1] = end_be; counters_out[2] = end_nan; counters_out[3] = end_iters; }
    }
}
