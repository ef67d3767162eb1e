// Lane-tiled Twin-T oscillator step (sm_100a): one oscillator = 8 lanes of one warp.
//
// The oscillator (gen_tremolo.rs:2353-3116, N = 7 nodes, M = 4 junctions of two Ebers-Moll BJTs) is input-independent and
// sample-serial: one sequence per preamp rate, shared by every instance of a launch.  As one thread it is a dependency chain of
// ~2400 instructions per step (4.1 us), and since nothing else of a tremolo render can start before its chunk of the
// sequence exists, that chain bounds the whole render once the chain kernel is faster than ~1 s per 3 s of audio.
// Here the step is spread over 8 lanes exactly like the preamp's (owg_tile.cuh):
//   * lane r < 7 owns node row r of build_rhs / S * rhs / S_NI * i_nl (lane 7 mirrors row 6): every row keeps the reference's
//     left-to-right summation order, rows padded with (-0.0 x 1.0) products to a common length;
//   * lane j = lane & 3 owns junction j of the Newton iteration: its junction voltage, ONE exponential (the partner junction of the
//     same transistor sits in lane j ^ 1, one shuffle away), its row of the Jacobian and its residual; the rows meet in shared memory
//     and every lane runs the 4x4 partial-pivoting elimination on identical data (trm_solve4, the generic code);
//   * the fast path covers "no junction step above 0.1 mV" (no pnjlim, gamma = 1, no 3.5 V cap) -- the oscillator's normal life;
//     anything else (limiter, singular pivot, a quotient outside the fast division's range, max iterations, non-finite) is handed
//     to lane 0, which runs the generic reference-order code (trm_nr_iter_exact / trm_be) on the gathered state.
// Every arithmetic operation is the reference's in the reference's order: the sequence is bit-identical to trm_step's.
#pragma once
#include "owg_tremolo.cuh"

namespace owgd {

#define OWG_TT_MASK 0xFFu
#define OWG_TT_TERMS 6
// home buffer (doubles): [0..6] flushed v_prev, [7] 1.0, [8..11] flushed i_nl_prev, [12..15] i_nl_prev_prev (cold path only)
#define OWG_TT_ONE 7
#define OWG_TT_IL 8
#define OWG_TT_PP 12

struct TrmRhsTerm { unsigned char kind, i, j, x; };  // kind 0: a_neg[i][j], 1: N_I[i][j], 2: pad (-0.0 x 1.0); x = home index of the operand
#define OWG_TT_PAD {2, 0, 0, OWG_TT_ONE}
// build_rhs rows (gen_tremolo.rs:2371-2410): structural non-zeros of a_neg in column order, then the N_i terms in device order
__constant__ TrmRhsTerm c_trm_rows[8][OWG_TT_TERMS] = {
    /*0*/ {{0, 0, 0, 0}, {0, 0, 1, 1}, {0, 0, 3, 3}, {0, 0, 5, 5}, {1, 0, 0, OWG_TT_IL + 0}, {1, 0, 2, OWG_TT_IL + 2}},
    /*1*/ {{0, 1, 0, 0}, {0, 1, 1, 1}, {0, 1, 2, 2}, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD},
    /*2*/ {{0, 2, 1, 1}, {0, 2, 2, 2}, {0, 2, 3, 3}, {1, 2, 1, OWG_TT_IL + 1}, OWG_TT_PAD, OWG_TT_PAD},
    /*3*/ {{0, 3, 0, 0}, {0, 3, 2, 2}, {0, 3, 3, 3}, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD},
    /*4*/ {{0, 4, 4, 4}, {1, 4, 0, OWG_TT_IL + 0}, {1, 4, 1, OWG_TT_IL + 1}, {1, 4, 3, OWG_TT_IL + 3}, OWG_TT_PAD, OWG_TT_PAD},
    /*5*/ {{0, 5, 0, 0}, {0, 5, 5, 5}, {0, 5, 6, 6}, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD},
    /*6*/ {OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD},
    /*7 = mirror of 6*/ {OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD, OWG_TT_PAD}};

struct TrmTileSm {
    double xs[16];
    double rs[8];
    __align__(16) double2 ex[2][3][4];  // Newton rows: [iteration parity][(a0,a1) | (a2,a3) | (f, -)][junction]
};

// per-lane constants, reloaded whenever the matrices change (set_sample_rate at step 50)
struct TrmLaneK {
    double coef[OWG_TT_TERMS];
    uint32_t xa[OWG_TT_TERMS];
    double rhs_const;
    double s_row[TN], sni_row[TM];
    double kr[TM];           // K row of this lane's junction
    double ka[TM], kb[TM];   // K rows 2b, 2b+1 of this lane's transistor b = junction >> 1
    double kd2, kd3;         // coefficients of the 3rd and 4th term of v_d (the generated code skips K[1][3] and K[2][2])
    double nva, nvb;         // N_v coefficients of p
    double jd[TM];           // row `junction` of the identity
};

__device__ __forceinline__ void trm_lane_consts(TrmLaneK& c, const TrmMats& m, const double* xs, const int r, const int jq) {
#pragma unroll
    for (int s = 0; s < OWG_TT_TERMS; s++) {
        const TrmRhsTerm t = c_trm_rows[r][s];
        c.coef[s] = t.kind == 0 ? m.a_neg[t.i][t.j] : (t.kind == 1 ? TRM_N_I[t.i][t.j] : -0.0);
        c.xa[s] = (uint32_t)__cvta_generic_to_shared(xs + t.x);
    }
    c.rhs_const = TRM_RHS_CONST[r];
#pragma unroll
    for (int j = 0; j < TN; j++) c.s_row[j] = m.s[r][j];
#pragma unroll
    for (int j = 0; j < TM; j++) { c.sni_row[j] = m.s_ni[r][j]; c.kr[j] = m.k[jq][j]; c.ka[j] = m.k[jq & 2][j]; c.kb[j] = m.k[(jq & 2) + 1][j]; c.jd[j] = j == jq ? 1.0 : 0.0; }
    c.kd2 = jq == 2 ? m.k[2][3] : m.k[jq][2];
    c.kd3 = (jq == 0 || jq == 3) ? m.k[jq][3] : -0.0;
    // p = N_v * v_pred (gen_tremolo.rs:2412-2420): p0 = NV[0][2] v2 + NV[0][4] v4, p1 = NV[1][0] v0 + NV[1][2] v2, p2 = NV[2][4] v4,
    // p3 = NV[3][0] v0 + NV[3][4] v4
    c.nva = jq == 0 ? TRM_N_V[0][2] : (jq == 1 ? TRM_N_V[1][0] : (jq == 2 ? TRM_N_V[2][4] : TRM_N_V[3][0]));
    c.nvb = jq == 0 ? TRM_N_V[0][4] : (jq == 1 ? TRM_N_V[1][2] : TRM_N_V[3][4]);
}

struct TrmTileDiag { uint32_t hist[16]; uint32_t be_fallback, nan_reset, generic_iters; };

// One process_sample(0.0) of the oscillator on 8 lanes.  State: home buffer `sm.xs` (flushed v_prev, i_nl_prev, i_nl_prev_prev) plus
// the registers (pi: Newton predictor of this step; raw_v / raw_il: the unflushed state the step leaves, kept for the launch's
// final TrmRun record).  Returns v[0] (identical in every lane).
__device__ __forceinline__ double trm_step_tile(TrmTileSm& sm, const TrmLaneK& c, const TrmMats& m, const TrmK& kq, double* sc, double (&pi)[TM],
                                                double& raw_v, double (&raw_il)[TM], const int lane, TrmTileDiag* dg) {
    const int r = lane < TN ? lane : TN - 1, jq = lane & 3;
    const bool even = (jq & 1) == 0;
    // ---- build_rhs row r (gen_tremolo.rs:2371-2410); input = input_prev = 0, so rhs[0] += (0 + 0) * g_in = +0.0 ----
    double acc = c.rhs_const;
#pragma unroll
    for (int s = 0; s < OWG_TT_TERMS; s++) acc += c.coef[s] * owg_lds64(c.xa[s]);
    acc += r == 0 ? (0.0 + 0.0) * kq.input_conductance : -0.0;
    sm.rs[lane] = acc;
    __syncwarp(OWG_TT_MASK);
    // ---- v_pred row r = S[r][:] . rhs, from 0.0 (gen_tremolo.rs:2412) ----
    double vp = 0.0;
#pragma unroll
    for (int j = 0; j < TN; j++) vp += c.s_row[j] * sm.rs[j];
    // ---- p of this lane's junction ----
    const double va = __shfl_sync(OWG_TT_MASK, vp, jq == 0 ? 2 : (jq == 2 ? 4 : 0), 8);
    const double vb = __shfl_sync(OWG_TT_MASK, vp, jq == 1 ? 2 : 4, 8);
    const double p = jq == 2 ? c.nva * va : c.nva * va + c.nvb * vb;
    // ---- Newton loop (gen_tremolo.rs:2423-2745), start = first-order predictor ----
    double il0 = pi[0], il1 = pi[1], il2 = pi[2], il3 = pi[3];
    const BjtK& bk = (jq & 2) ? kq.q1 : kq.q0;
    const Recip rc = even ? bk.r_nf : bk.r_nr;
    uint32_t last = T_MAX_ITER;
    for (int iter = 0; iter < T_MAX_ITER; iter++) {
        DivPolicy<false> D;
        const double x2 = jq == 2 ? il3 : il2;
        const double x3 = (jq == 0 || jq == 3) ? il3 : 1.0;
        const double v_d = p + c.kr[0] * il0 + c.kr[1] * il1 + c.kd2 * x2 + c.kd3 * x3;
        const double e_own = fast_exp(D.div(v_d, rc));
        const double e_par = __shfl_xor_sync(OWG_TT_MASK, e_own, 1, 8);
        const double exp_be = even ? e_own : e_par, exp_bc = even ? e_par : e_own;
        // bjt_evaluate, Ebers-Moll branch (gen_tremolo.rs:1566-1636): even lanes need (ic, j0, j1), odd lanes (ib, j2, j3)
        const double ib_rev = bk.is_br * (exp_bc - 1.0);
        const double ic = bk.is * (exp_be - exp_bc) - ib_rev;
        const double ib = bk.is_bf * (exp_be - 1.0) + ib_rev;
        const double ja = even ? bk.is_nfvt * exp_be : bk.is_bf_nfvt * exp_be;
        const double jb = even ? -bk.is_nrvt * exp_bc - bk.is_br_nrvt * exp_bc : bk.is_br_nrvt * exp_bc;
        const double il_own = jq == 0 ? il0 : (jq == 1 ? il1 : (jq == 2 ? il2 : il3));
        const double f = il_own - (even ? ic : ib);
        // row jq of J = I - J_dev K (gen_tremolo.rs:2497-2512)
        double2* e0 = &sm.ex[iter & 1][0][0];
        if (lane < 4) {
            e0[jq] = make_double2(c.jd[0] - ja * c.ka[0] - jb * c.kb[0], c.jd[1] - ja * c.ka[1] - jb * c.kb[1]);
            e0[4 + jq] = make_double2(c.jd[2] - ja * c.ka[2] - jb * c.kb[2], c.jd[3] - ja * c.ka[3] - jb * c.kb[3]);
            e0[8 + jq] = make_double2(f, 0.0);
        }
        __syncwarp(OWG_TT_MASK);
        double a[4][4], b[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double2 t0 = e0[q], t1 = e0[4 + q], t2 = e0[8 + q];
            a[q][0] = t0.x; a[q][1] = t0.y; a[q][2] = t1.x; a[q][3] = t1.y; b[q] = t2.x;
        }
        bool singular;
        trm_solve4(a, b, singular, D);
        const double d0 = b[0], d1 = b[1], d2 = b[2], d3 = b[3];
        const double it0 = il0 - d0, it1 = il1 - d1, it2 = il2 - d2, it3 = il3 - d3;
        const double vt = p + c.kr[0] * it0 + c.kr[1] * it1 + c.kr[2] * it2 + c.kr[3] * it3;
        const double dvt = vt - v_d;
        // fast path <=> no junction moves by more than 0.1 mV: then pnjlim returns its argument, every ratio is exactly 1, gamma = 1,
        // the 3.5 V cap is out of reach and il -= 1.0 * d is the trial point itself.  `wild` keeps everything below finite.
        const bool wild = !(fabs(d0) <= KC(38)) || !(fabs(d1) <= KC(38)) || !(fabs(d2) <= KC(38)) || !(fabs(d3) <= KC(38));
        const bool generic = singular || D.bad != 0u || wild || !(fabs(dvt) <= KC(13));
        const double thr = KC(9) * fmax(fabs(v_d), fabs(v_d + dvt)) + KC(10);
        const bool fail = fabs(dvt) > thr;
        const unsigned gbal = __ballot_sync(OWG_TT_MASK, generic);
        bool conv;
        if (gbal != 0u) {  // generic reference-order iteration on lane 0 from the same iterate
            if (lane < 4) sc[30 + jq] = p;
            if (lane == 0) { sc[26] = il0; sc[27] = il1; sc[28] = il2; sc[29] = il3; }
            __syncwarp(OWG_TT_MASK);
            if (lane == 0) { sc[25] = trm_nr_iter_exact(sc, m, kq) ? 1.0 : 0.0; if (dg) dg->generic_iters++; }
            __syncwarp(OWG_TT_MASK);
            il0 = sc[26]; il1 = sc[27]; il2 = sc[28]; il3 = sc[29];
            conv = sc[25] != 0.0;
            __syncwarp(OWG_TT_MASK);
        } else {
            il0 = it0; il1 = it1; il2 = it2; il3 = it3;
            conv = (__ballot_sync(OWG_TT_MASK, fail) & 0xFu) == 0u;
        }
        if (conv) { last = (uint32_t)iter; break; }
    }
    if (dg && lane == 0) dg->hist[last < 15u ? last : 15u]++;
    // ---- v = v_pred + S_NI * i_nl (gen_tremolo.rs:2747-2755) ----
    double v = vp;
    v += c.sni_row[0] * il0;
    v += c.sni_row[1] * il1;
    v += c.sni_row[2] * il2;
    v += c.sni_row[3] * il3;
    const double pl0 = sm.xs[OWG_TT_IL], pl1 = sm.xs[OWG_TT_IL + 1], pl2 = sm.xs[OWG_TT_IL + 2], pl3 = sm.xs[OWG_TT_IL + 3];  // flushed i_nl_prev
    if (!(last < (uint32_t)T_MAX_ITER)) {  // max iterations: backward-Euler fallback (gen_tremolo.rs:2757-3083), generic code on lane 0
        __syncwarp(OWG_TT_MASK);
        if (lane == 0) {
            if (dg) dg->be_fallback++;
            for (int i = 0; i < TN; i++) sc[i] = sm.xs[i];
            for (int i = 0; i < TM; i++) { sc[7 + i] = sm.xs[OWG_TT_IL + i]; sc[11 + i] = sm.xs[OWG_TT_PP + i]; }
            trm_be(0.0, sc, m, kq);
        }
        __syncwarp(OWG_TT_MASK);
        v = sc[15 + r];
        il0 = sc[22]; il1 = sc[23]; il2 = sc[24]; il3 = sc[25];
        __syncwarp(OWG_TT_MASK);
    }
    const bool fin = (__ballot_sync(OWG_TT_MASK, !finite64(v)) & 0x7Fu) == 0u;
    __syncwarp(OWG_TT_MASK);  // every lane has read the home buffer of this step before anyone rewrites it
    double out;
    if (!fin) {  // gen_tremolo.rs:3085-3105
        if (dg && lane == 0) dg->nan_reset++;
        raw_v = TRM_DC_OP[r];
#pragma unroll
        for (int i = 0; i < TM; i++) raw_il[i] = TRM_DC_NL_I[i];
        sm.xs[lane < TN ? lane : 6] = raw_v + KC(8) - KC(8);
        double nl[TM];
#pragma unroll
        for (int i = 0; i < TM; i++) nl[i] = raw_il[i] + KC(8) - KC(8);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < TM; i++) { sm.xs[OWG_TT_IL + i] = nl[i]; sm.xs[OWG_TT_PP + i] = TRM_DC_NL_I[i]; }
        }
#pragma unroll
        for (int i = 0; i < TM; i++) pi[i] = 2.0 * nl[i] - TRM_DC_NL_I[i];
        out = 4.26480458363572357e0;
    } else {
        // state shift (gen_tremolo.rs:3107-3114) into the home buffer, flushed for the next step; next Newton start
        raw_v = v;
        raw_il[0] = il0; raw_il[1] = il1; raw_il[2] = il2; raw_il[3] = il3;
        if (lane < TN) sm.xs[lane] = v + KC(8) - KC(8);
        const double nl0 = il0 + KC(8) - KC(8), nl1 = il1 + KC(8) - KC(8), nl2 = il2 + KC(8) - KC(8), nl3 = il3 + KC(8) - KC(8);
        if (lane == 0) {
            sm.xs[OWG_TT_IL] = nl0; sm.xs[OWG_TT_IL + 1] = nl1; sm.xs[OWG_TT_IL + 2] = nl2; sm.xs[OWG_TT_IL + 3] = nl3;
            sm.xs[OWG_TT_PP] = pl0; sm.xs[OWG_TT_PP + 1] = pl1; sm.xs[OWG_TT_PP + 2] = pl2; sm.xs[OWG_TT_PP + 3] = pl3;
        }
        pi[0] = 2.0 * nl0 - pl0; pi[1] = 2.0 * nl1 - pl1; pi[2] = 2.0 * nl2 - pl2; pi[3] = 2.0 * nl3 - pl3;
        out = __shfl_sync(OWG_TT_MASK, v, 0, 8);
    }
    __syncwarp(OWG_TT_MASK);
    return out;
}

// Loads a TrmState (raw, as the one-thread kernels keep it) into the tile's home buffer and registers.
__device__ __forceinline__ void trm_tile_load(TrmTileSm& sm, const TrmState& st, double (&pi)[TM], double& raw_v, double (&raw_il)[TM], const int lane) {
    const int r = lane < TN ? lane : TN - 1;
    raw_v = st.v[r];
    double nl[TM];
#pragma unroll
    for (int i = 0; i < TM; i++) { raw_il[i] = st.il[i]; nl[i] = st.il[i] + KC(8) - KC(8); pi[i] = 2.0 * nl[i] - st.ilpp[i]; }
    if (lane < TN) sm.xs[lane] = raw_v + KC(8) - KC(8);
    if (lane == 0) {
        sm.xs[OWG_TT_ONE] = 1.0;
#pragma unroll
        for (int i = 0; i < TM; i++) { sm.xs[OWG_TT_IL + i] = nl[i]; sm.xs[OWG_TT_PP + i] = st.ilpp[i]; }
    }
    __syncwarp(OWG_TT_MASK);
}
// ... and back: v / il raw, ilpp = the flushed i_nl_prev of the last step = what the home buffer holds as i_nl_prev_prev's successor
__device__ __forceinline__ void trm_tile_store(const TrmTileSm& sm, TrmState& st, const double raw_v, const double (&raw_il)[TM], const int lane) {
    __syncwarp(OWG_TT_MASK);
    if (lane < TN) st.v[lane] = raw_v;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < TM; i++) { st.il[i] = raw_il[i]; st.ilpp[i] = sm.xs[OWG_TT_PP + i]; }
        st.xin_prev = 0.0;
    }
}

}  // namespace owgd
