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

__device__ __forceinline__ double owg_lds64(uint32_t addr) {  // shared-memory load from a 32-bit shared-window address
    double x;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
    return x;
}

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
    __align__(16) double2 ex[2][3][8];  // Newton rows: [iteration parity][(a0,a1) | (a2,a3) | (f, 1/a0)][lane]; lanes 4..7 are mirror copies
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

struct TrmTileDiag { uint32_t hist[16]; uint32_t be_fallback, nan_reset; };
__device__ unsigned long long g_trm_generic;  // Newton iterations of the tiled oscillator that took the generic code (debug counter)

// One process_sample(0.0) of the oscillator on 8 lanes.  State: home buffer `sm.xs` (flushed v_prev, i_nl_prev, i_nl_prev_prev) plus
// the registers (pi: Newton predictor of this step; raw_v / raw_il: the unflushed state the step leaves, kept for the launch's
// final TrmRun record).  Returns v[0] (identical in every lane).
__device__ __forceinline__ double trm_step_tile(TrmTileSm& sm, const TrmLaneK& c, const TrmMats& m, const TrmK& kq, double* sc, double (&pi)[TM],
                                                double& raw_v, double (&raw_il)[TM], const int lane, TrmTileDiag* dg, uint32_t& generic_count) {
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
    const double cja = even ? bk.is_nfvt : bk.is_bf_nfvt;
    const bool lane2 = jq == 2, lane03 = (jq == 0) | (jq == 3), lane1 = jq == 1;
    uint32_t last = T_MAX_ITER;
#pragma unroll 1
    for (int iter = 0; iter < T_MAX_ITER; iter++) {
        unsigned bad = 0;
        const double x2 = lane2 ? il3 : il2;
        const double x3 = lane03 ? il3 : 1.0;
        const double v_d = p + c.kr[0] * il0 + c.kr[1] * il1 + c.kd2 * x2 + c.kd3 * x3;
        const double e_own = fast_exp_sl(div_sl(v_d, rc, bad, true));
        const double e_par = __shfl_xor_sync(OWG_TT_MASK, e_own, 1, 8);
        const double exp_be = even ? e_own : e_par, exp_bc = even ? e_par : e_own;
        // bjt_evaluate, Ebers-Moll branch (gen_tremolo.rs:1566-1636): even lanes need (ic, j0, j1), odd lanes (ib, j2, j3)
        const double ib_rev = bk.is_br * (exp_bc - 1.0);
        const double ic = bk.is * (exp_be - exp_bc) - ib_rev;
        const double ib = bk.is_bf * (exp_be - 1.0) + ib_rev;
        const double ja = cja * exp_be;                 // j0 = is/(nf vt) exp_be | j2 = is/(bf nf vt) exp_be
        const double j3 = bk.is_br_nrvt * exp_bc;
        const double j1 = -bk.is_nrvt * exp_bc - j3;    // the reference's second product is j3's, bit for bit
        const double jb = even ? j1 : j3;
        const double il_own = lane2 ? il2 : (lane1 ? il1 : (jq == 0 ? il0 : il3));
        const double f = il_own - (even ? ic : ib);
        // row jq of J = I - J_dev K (gen_tremolo.rs:2497-2512), its right-hand side, and -- speculatively, while the other rows are still
        // being written -- the reciprocal of its first element: whichever row wins the first pivot search brings its reciprocal along
        const double a_own0 = c.jd[0] - ja * c.ka[0] - jb * c.kb[0];
        const Recip r_own = recip_prepare(a_own0);
        // (lanes 4..7 mirror lanes 0..3 bit for bit and store into their own slots, which nobody reads: no branch around the stores)
        double2* e0 = &sm.ex[iter & 1][0][0];
        e0[lane] = make_double2(a_own0, c.jd[1] - ja * c.ka[1] - jb * c.kb[1]);
        e0[8 + lane] = make_double2(c.jd[2] - ja * c.ka[2] - jb * c.kb[2], c.jd[3] - ja * c.ka[3] - jb * c.kb[3]);
        e0[16 + lane] = make_double2(f, r_own.r);
        __syncwarp(OWG_TT_MASK);
        // ---- 4x4 elimination with partial pivoting (gen_tremolo.rs:2515-2561), straight-line, identically in every lane ----
        // column 0: the sequential "first strict maximum" search is a set of pairwise comparisons; swap(0, max_row) by ADDRESS
        bool singular;
        double d0, d1, d2, d3;
        {
            const double m0 = fabs(e0[0].x), m1 = fabs(e0[1].x), m2 = fabs(e0[2].x), m3 = fabs(e0[3].x);
            // (bitwise & | on purpose: no short-circuit branches in the loop body)
            const bool w3 = (m3 > m0) & (m3 > m1) & (m3 > m2);
            const bool w2 = !w3 & (m2 > m0) & (m2 > m1);
            const bool w1 = !w3 & !w2 & (m1 > m0);
            const int mr = (w3 ? 3 : 0) | (w2 ? 2 : 0) | (w1 ? 1 : 0);
            const int i1 = w1 ? 0 : 1, i2 = w2 ? 0 : 2, i3 = w3 ? 0 : 3;
            const double2 Pa = e0[mr], Pb = e0[8 + mr], Pc = e0[16 + mr];
            const double2 Aa = e0[i1], Ab = e0[8 + i1], Ac = e0[16 + i1];
            const double2 Ba = e0[i2], Bb = e0[8 + i2], Bc = e0[16 + i2];
            const double2 Ca = e0[i3], Cb = e0[8 + i3], Cc = e0[16 + i3];
            const double P0 = Pa.x, P1 = Pa.y, P2 = Pb.x, P3 = Pb.y, PB = Pc.x;
            singular = fabs(P0) < KC(14);
            Recip rp; rp.r = Pc.y; rp.nb = -P0; rp.b = P0;
            double A1 = Aa.y, A2 = Ab.x, A3 = Ab.y, AB = Ac.x;
            double B1 = Ba.y, B2 = Bb.x, B3 = Bb.y, BB = Bc.x;
            double C1 = Ca.y, C2 = Cb.x, C3 = Cb.y, CB = Cc.x;
            const double fa = div_sl(Aa.x, rp, bad, true), fb = div_sl(Ba.x, rp, bad, true), fc = div_sl(Ca.x, rp, bad, true);
            A1 -= fa * P1; A2 -= fa * P2; A3 -= fa * P3; AB -= fa * PB;
            B1 -= fb * P1; B2 -= fb * P2; B3 -= fb * P3; BB -= fb * PB;
            C1 -= fc * P1; C2 -= fc * P2; C3 -= fc * P3; CB -= fc * PB;
            // column 1 among rows A, B, C
            const double n1 = fabs(A1), n2 = fabs(B1), n3 = fabs(C1);
            const bool y3 = (n3 > n1) & (n3 > n2);
            const bool y2 = !y3 & (n2 > n1);
            const double S1 = y3 ? C1 : (y2 ? B1 : A1), S2 = y3 ? C2 : (y2 ? B2 : A2), S3 = y3 ? C3 : (y2 ? B3 : A3), SB = y3 ? CB : (y2 ? BB : AB);
            const double T1 = y2 ? A1 : B1, TB0 = y2 ? AB : BB;
            double T2 = y2 ? A2 : B2, T3 = y2 ? A3 : B3;
            const double U1 = y3 ? A1 : C1, UB0 = y3 ? AB : CB;
            double U2 = y3 ? A2 : C2, U3 = y3 ? A3 : C3;
            singular = singular | (fabs(S1) < KC(14));
            const Recip rs = recip_prepare(S1);
            const double ft = div_sl(T1, rs, bad, true), fu = div_sl(U1, rs, bad, true);
            T2 -= ft * S2; T3 -= ft * S3;
            const double TB = TB0 - ft * SB;
            U2 -= fu * S2; U3 -= fu * S3;
            const double UB = UB0 - fu * SB;
            // column 2 among rows T, U
            const bool z = fabs(U2) > fabs(T2);
            const double V2 = z ? U2 : T2, V3 = z ? U3 : T3, VB = z ? UB : TB;
            const double W2 = z ? T2 : U2, WB0 = z ? TB : UB;
            double W3 = z ? T3 : U3;
            singular = singular | (fabs(V2) < KC(14));
            const Recip rv = recip_prepare(V2);
            const double fw = div_sl(W2, rv, bad, true);
            W3 -= fw * V3;
            const double WB = WB0 - fw * VB;
            // column 3: pivot test only; back substitution, sums in ascending column order
            singular = singular | (fabs(W3) < KC(14));
            const Recip rw = recip_prepare(W3);
            d3 = div_sl(WB, rw, bad, true);
            d2 = div_sl(VB - V3 * d3, rv, bad, true);
            double s1 = SB - S2 * d2;
            s1 -= S3 * d3;
            d1 = div_sl(s1, rs, bad, true);
            double s0 = PB - P1 * d1;
            s0 -= P2 * d2;
            s0 -= P3 * d3;
            d0 = div_sl(s0, rp, bad, true);
        }
        const double it0 = il0 - d0, it1 = il1 - d1, it2 = il2 - d2, it3 = il3 - d3;
        const double vt = p + c.kr[0] * it0 + c.kr[1] * it1 + c.kr[2] * it2 + c.kr[3] * it3;
        const double dvt = vt - v_d;
        // fast path <=> pnjlim returns its argument at every junction (it limits only when vtrial > vcrit AND |vtrial - v_d| > 2 vt,
        // gen_tremolo.rs:1181-1196) and no junction moves by more than 3.5 V: then every ratio is exactly 1, gamma = 1, the cap does
        // not act and il -= 1.0 * d is the trial point itself (gen_tremolo.rs:2640-2700).  `wild` keeps everything below finite.
        const bool wild = !(fabs(d0) <= KC(38)) | !(fabs(d1) <= KC(38)) | !(fabs(d2) <= KC(38)) | !(fabs(d3) <= KC(38));
        const double jvt = (jq & 2) ? TRM_DEVICE_1_VT : TRM_DEVICE_0_VT, jvcrit = (jq & 2) ? TRM_DEVICE_1_VCRIT : TRM_DEVICE_0_VCRIT;
        const bool generic = singular | (bad != 0u) | wild | !(fabs(dvt) <= 3.5) | ((vt > jvcrit) & (fabs(dvt) > jvt + jvt));
        const double thr = KC(9) * fmax(fabs(v_d), fabs(v_d + dvt)) + KC(10);
        const bool fail = fabs(dvt) > thr;
        const unsigned gbal = __ballot_sync(OWG_TT_MASK, generic);
        bool conv;
        if (gbal != 0u) {  // generic reference-order iteration on lane 0 from the same iterate
            if (lane < 4) sc[30 + jq] = p;
            if (lane == 0) { sc[26] = il0; sc[27] = il1; sc[28] = il2; sc[29] = il3; }
            __syncwarp(OWG_TT_MASK);
            if (lane == 0) { sc[25] = trm_nr_iter_exact(sc, m, kq) ? 1.0 : 0.0; generic_count++; }
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
