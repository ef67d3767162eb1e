// Legacy 8-node DK preamp on the device (crates/openwurli-dsp/src/dk_preamp_legacy.rs: the preamp of the reference's
// default build), selected with owg_opts.preamp_model = OWG_PREAMP_LEGACY8.
//
// All matrices are R_ldr-independent and computed on the host at plan time (host_setup.cpp make_legacy_group); the LDR enters
// each sample through two scalars (g_ldr, g_ldr_prev) and a Sherman-Morrison correction, so tremolo groups need no per-sample
// matrix records -- only the g_ldr sequence produced by tremolo_group_kernel.  One warp = up to 31 instances of one group
// plus the group's zero-input shadow instance in lane 31 (out = main - shadow, dk_preamp_legacy.rs:557-607).
#pragma once
#include "owg_kernels.cuh"

namespace owgd {

struct LgState { double v[8], i_nl[2], v_nl[2], j_cin, cin_prev; };

#define LG_BASE1 0
#define LG_EMIT1 1
#define LG_COLL1 2
#define LG_EMIT2 3
#define LG_EMIT2B 4
#define LG_COLL2 5
#define LG_OUT 6
#define LG_FB 7

// bjt_ic_gm (dk_preamp_legacy.rs:686-690): one libm exp per junction per Newton iteration
__device__ __forceinline__ void lg_ic_gm(double vbe, double& ic, double& gm) {
    const double e = exp(rclamp(vbe, -1.0, 0.85) / 0.026);
    ic = 3.03e-14 * (e - 1.0);
    gm = (3.03e-14 / 0.026) * e;
}
__device__ __forceinline__ double lg_ic(double vbe) { return 3.03e-14 * (exp(rclamp(vbe, -1.0, 0.85) / 0.026) - 1.0); }

// dk_step (dk_preamp_legacy.rs:447-554).  m = the group's record in shared memory.  Returns v[OUT]; *iters = Newton updates.
__device__ __forceinline__ double lg_step(LgState& st, const double* __restrict__ m, double input, double g_ldr, double g_prev, int* iters) {
    double rhs[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) sum += m[OWG_LG_AN + i * 8 + j] * st.v[j];
        rhs[i] = sum;
    }
    rhs[LG_FB] -= g_prev * st.v[LG_FB];
    const double cin_now = m[OWG_LG_GCIN] * input + st.j_cin;
    rhs[LG_BASE1] += cin_now + st.cin_prev;
    rhs[LG_EMIT1] += st.i_nl[0];
    rhs[LG_COLL1] -= st.i_nl[0];
    rhs[LG_EMIT2] += st.i_nl[1];
    rhs[LG_COLL2] -= st.i_nl[1];
#pragma unroll
    for (int i = 0; i < 8; i++) rhs[i] += m[OWG_LG_W2 + i];
    double vpb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) sum += m[OWG_LG_S + i * 8 + j] * rhs[j];
        vpb[i] = sum;
    }
    const double sm_k = g_ldr / (1.0 + m[OWG_LG_SFBFB] * g_ldr);
    const double sm_vpred = sm_k * vpb[LG_FB];
    double vp[8];
#pragma unroll
    for (int i = 0; i < 8; i++) vp[i] = vpb[i] - sm_vpred * m[OWG_LG_SFB + i];
    const double p0 = vp[LG_BASE1] - vp[LG_EMIT1], p1 = vp[LG_COLL1] - vp[LG_EMIT2];
    const double k00 = m[OWG_LG_K + 0] - sm_k * m[OWG_LG_NVSFB + 0] * m[OWG_LG_SFBNI + 0];
    const double k01 = m[OWG_LG_K + 1] - sm_k * m[OWG_LG_NVSFB + 0] * m[OWG_LG_SFBNI + 1];
    const double k10 = m[OWG_LG_K + 2] - sm_k * m[OWG_LG_NVSFB + 1] * m[OWG_LG_SFBNI + 0];
    const double k11 = m[OWG_LG_K + 3] - sm_k * m[OWG_LG_NVSFB + 1] * m[OWG_LG_SFBNI + 1];
    double v0 = st.v_nl[0], v1 = st.v_nl[1];
    int n_it = 0;
#pragma unroll 1
    for (int it = 0; it < 6; it++) {
        double ic0, gm0, ic1, gm1;
        lg_ic_gm(v0, ic0, gm0);
        lg_ic_gm(v1, ic1, gm1);
        const double f0 = v0 - p0 - k00 * ic0 - k01 * ic1;
        const double f1 = v1 - p1 - k10 * ic0 - k11 * ic1;
        if (fabs(f0) < 1e-9 && fabs(f1) < 1e-9) break;
        const double j00 = 1.0 - k00 * gm0;
        const double j01 = -k01 * gm1;
        const double j10 = -k10 * gm0;
        const double j11 = 1.0 - k11 * gm1;
        const double det = j00 * j11 - j01 * j10;
        if (fabs(det) < 1e-30) break;
        const double inv_det = 1.0 / det;
        v0 -= inv_det * (j11 * f0 - j01 * f1);
        v1 -= inv_det * (j00 * f1 - j10 * f0);
        n_it++;
    }
    *iters = n_it;
    const double ic0 = lg_ic(v0), ic1 = lg_ic(v1);
    const double dot = m[OWG_LG_SFBNI + 0] * ic0 + m[OWG_LG_SFBNI + 1] * ic1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double s_ni = ic0 * m[OWG_LG_D0 + i] + ic1 * m[OWG_LG_D1 + i];
        st.v[i] = vp[i] + s_ni - sm_k * dot * m[OWG_LG_SFB + i];
    }
    st.cin_prev = cin_now;
    const double dv_cin = input - st.v[LG_BASE1];
    st.j_cin = -m[OWG_LG_GC1PC] * dv_cin - m[OWG_LG_CCIN] * st.j_cin;
    st.i_nl[0] = ic0; st.i_nl[1] = ic1;
    st.v_nl[0] = v0; st.v_nl[1] = v1;
    return st.v[LG_OUT];
}

__device__ __forceinline__ void lg_init(LgState& st, const double* __restrict__ m) {
#pragma unroll
    for (int i = 0; i < 8; i++) st.v[i] = m[OWG_LG_V0 + i];
    st.i_nl[0] = m[OWG_LG_INL0]; st.i_nl[1] = m[OWG_LG_INL0 + 1];
    st.v_nl[0] = m[OWG_LG_VNL0]; st.v_nl[1] = m[OWG_LG_VNL0 + 1];
    st.j_cin = m[OWG_LG_JCIN0]; st.cin_prev = m[OWG_LG_CINPREV0];
}

// chain B / preamp-only batch with the legacy preamp; same contract as chain_kernel (in-place on `out`, chunked with carried
// state, optional metrics).  TREM: g_seq[group_rec_index[group]][preamp-rate sample] holds g_ldr after each set_ldr_resistance.
#define OWG_LG_CARRY 34  // 14 preamp + 13 oversampler + 5 speaker + g_ldr_prev + spare
template <bool TREM>
__global__ void __launch_bounds__(32) chain_legacy_kernel(const WarpEntry* __restrict__ warps, const int32_t* __restrict__ order,
                                                          const OwgChainInit* __restrict__ cinits, const unsigned long long* __restrict__ n_samples,
                                                          const double* __restrict__ grecs /*[group][OWG_LG_STRIDE]*/, const double* __restrict__ g_seq,
                                                          const int32_t* __restrict__ group_rec_index, int64_t g_stride,
                                                          double* __restrict__ out, int64_t stride, DevDiag* diag,
                                                          int64_t t_begin, int64_t t_end, double* __restrict__ carry /*[warp][OWG_CARRY][32]*/,
                                                          double* __restrict__ metrics, const double* __restrict__ f0s, int64_t w_begin, int64_t w_end, int taps) {
    __shared__ double s_m[OWG_LG_STRIDE];
    __shared__ OwgChainInit s_ci[32];
    const int lane = threadIdx.x;
    const WarpEntry we = warps[blockIdx.x];
    const bool is_shadow = lane == 31;
    const bool is_main = lane < we.count;
    const int32_t job = is_main ? order[we.first + lane] : -1;
    for (int e = lane; e < OWG_LG_STRIDE; e += 32) s_m[e] = grecs[(size_t)we.group * OWG_LG_STRIDE + e];
    if (is_main) s_ci[lane] = cinits[job];
    else {
        OwgChainInit z;
        z.volume = 0.0; z.spk_a2 = 0.0; z.spk_a3 = 0.0; z.spk_norm = 1.0; z.spk_thermal_coeff = 0.0; z.spk_thermal_alpha = 0.0;
        z.hpf_b0 = z.hpf_b1 = z.hpf_b2 = z.hpf_a1 = z.hpf_a2 = 0.0; z.lpf_b0 = z.lpf_b1 = z.lpf_b2 = z.lpf_a1 = z.lpf_a2 = 0.0;
        z.spk_tanh = 0; z.group = we.group; z.no_preamp = 0; z.no_poweramp = 1; z.oversample = 0; z.pre_only = 0;
        s_ci[lane] = z;
    }
    __syncwarp();
    const OwgChainInit& ci = s_ci[lane];
    const int oversample = __shfl_sync(0xffffffffu, ci.oversample, 0);
    const unsigned long long ns = is_main ? n_samples[job] : 0ull;
    double* o = is_main ? out + (size_t)job * stride : nullptr;
    const double* gs = TREM ? g_seq + (size_t)group_rec_index[we.group] * g_stride : nullptr;

    LgState st;
    lg_init(st, s_m);  // DkPreamp::new (+ reset() in static mode): both instances at the 1 MOhm DC point
    double g_prev = s_m[OWG_LG_GINIT];
    double ua[3] = {0, 0, 0}, ub[3] = {0, 0, 0}, da[3] = {0, 0, 0}, db[3] = {0, 0, 0};
    double down_delay = 0.0;
    SpkState spk = {0.0, 0.0, 0.0, 0.0, 0.0};
    const double vol = ci.volume;
    const bool bypass_preamp = ci.no_preamp != 0;
    uint32_t hist[7] = {0, 0, 0, 0, 0, 0, 0};
    uint32_t pa_hist[9];
    for (int i = 0; i < 9; i++) pa_hist[i] = 0;
    uint32_t nan_resets = 0;
    const int n_sub = oversample ? 2 : 1;
    double* cw = carry ? carry + (size_t)blockIdx.x * OWG_CARRY * 32 + lane : nullptr;
    if (cw && t_begin > 0) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) st.v[i] = cw[(k++) * 32];
        st.i_nl[0] = cw[(k++) * 32]; st.i_nl[1] = cw[(k++) * 32]; st.v_nl[0] = cw[(k++) * 32]; st.v_nl[1] = cw[(k++) * 32];
        st.j_cin = cw[(k++) * 32]; st.cin_prev = cw[(k++) * 32];
#pragma unroll
        for (int i = 0; i < 3; i++) { ua[i] = cw[(k++) * 32]; ub[i] = cw[(k++) * 32]; da[i] = cw[(k++) * 32]; db[i] = cw[(k++) * 32]; }
        down_delay = cw[(k++) * 32];
        spk.thermal = cw[(k++) * 32]; spk.h1 = cw[(k++) * 32]; spk.h2 = cw[(k++) * 32]; spk.l1 = cw[(k++) * 32]; spk.l2 = cw[(k++) * 32];
        g_prev = cw[(k++) * 32];
    }
    double m_peak = 0.0, m_sq = 0.0, m_re1 = 0.0, m_im1 = 0.0, m_re2 = 0.0, m_im2 = 0.0, m_f0 = 0.0, m_sr = 1.0;
    double q_peak = 0.0, q_sq = 0.0, q_re1 = 0.0, q_im1 = 0.0, q_re2 = 0.0, q_im2 = 0.0;  // T4 (preamp output), calibrate taps only
    if (metrics && is_main) {
        const double* mj = metrics + (size_t)job * OWG_METRICS;
        m_peak = mj[0]; m_sq = mj[1]; m_re1 = mj[2]; m_im1 = mj[3]; m_re2 = mj[4]; m_im2 = mj[5];
        if (taps) { q_peak = mj[OWG_MET_T4]; q_sq = mj[OWG_MET_T4 + 1]; q_re1 = mj[OWG_MET_T4 + 2]; q_im1 = mj[OWG_MET_T4 + 3]; q_re2 = mj[OWG_MET_T4 + 4]; q_im2 = mj[OWG_MET_T4 + 5]; }
        m_f0 = f0s[2 * job]; m_sr = f0s[2 * job + 1];
    }
    const int64_t t_stop = t_end < we.n_max ? t_end : we.n_max;
    int64_t tos = t_begin * n_sub;
    const double g_static = s_m[OWG_LG_GSTATIC];
    double x_next = (is_main && (unsigned long long)t_begin < ns) ? o[t_begin] : 0.0;
    double g_next = TREM ? (tos < g_stride ? gs[tos] : g_prev) : g_static;
    for (int64_t t = t_begin; t < t_stop; t++) {
        const bool live = is_main && (unsigned long long)t < ns;
        const double x = x_next;
        x_next = (is_main && (unsigned long long)(t + 1) < ns) ? o[t + 1] : 0.0;
        double u0 = x, u1 = 0.0;
        if (oversample) {
            u0 = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, ua, x);
            u1 = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, ub, x);
        }
        if (is_shadow) { u0 = 0.0; u1 = 0.0; }
        double p0 = 0.0, p1 = 0.0;
#pragma unroll 1
        for (int j = 0; j < n_sub; j++) {
            const double g_ldr = g_next;
            if (TREM) g_next = tos + 1 < g_stride ? gs[tos + 1] : g_ldr;
            int iters;
            const double main_out = lg_step(st, s_m, j == 0 ? u0 : u1, g_ldr, g_prev, &iters);
            g_prev = g_ldr;  // shared R_ldr tracking, updated after both instances stepped (:595)
            if (diag && live && !bypass_preamp) {
#pragma unroll
                for (int k = 0; k < 7; k++) hist[k] += (iters == k) ? 1u : 0u;
            }
            const double pump = __shfl_sync(0xffffffffu, main_out, 31);
            double res = main_out - pump;
            if (!finite64(res)) {
                // reference: reset() = DC solve at the CURRENT R_ldr for both instances (:608-615, 628-642).  Deviation: this
                // instance alone restarts from the plan-time 1 MOhm DC point (the shadow is shared); unreachable with finite input.
                nan_resets++;
                lg_init(st, s_m);
                res = 0.0;
            }
            if (j == 0) p0 = res; else p1 = res;
            tos += 1;
        }
        double pre_out;
        if (oversample) {
            const double a = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, da, p0);
            const double b = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, db, p1);
            pre_out = (a + down_delay) * 0.5;
            down_delay = b;
        } else pre_out = p0;
        if (bypass_preamp) pre_out = x;
        if (live && ci.pre_only) o[t] = pre_out;
        else if (live) {
            const double att = pre_out * vol * vol;
            const double amped = ci.no_poweramp ? att : poweramp(att, diag ? pa_hist : nullptr);
            const double y_final = speaker(amped, spk, ci) * 7.498942093324558;
            if (metrics) {
                if (t >= w_begin && t < w_end) {
                    const double ii = (double)(t - w_begin);
                    m_peak = fmax(m_peak, fabs(y_final));
                    m_sq += y_final * y_final;
                    const double ph1 = 2.0 * 3.14159265358979323846 * m_f0 * ii / m_sr;
                    const double ph2 = 2.0 * 3.14159265358979323846 * (2.0 * m_f0) * ii / m_sr;
                    const double c1 = cos(ph1), s1 = sin(ph1), c2 = cos(ph2), s2 = sin(ph2);
                    m_re1 += y_final * c1; m_im1 -= y_final * s1;
                    m_re2 += y_final * c2; m_im2 -= y_final * s2;
                    if (taps) {
                        q_peak = fmax(q_peak, fabs(pre_out));
                        q_sq += pre_out * pre_out;
                        q_re1 += pre_out * c1; q_im1 -= pre_out * s1;
                        q_re2 += pre_out * c2; q_im2 -= pre_out * s2;
                    }
                }
            } else o[t] = y_final;
        }
    }
    if (metrics && is_main) {
        double* mj = metrics + (size_t)job * OWG_METRICS;
        mj[0] = m_peak; mj[1] = m_sq; mj[2] = m_re1; mj[3] = m_im1; mj[4] = m_re2; mj[5] = m_im2;
        if (taps) { mj[OWG_MET_T4] = q_peak; mj[OWG_MET_T4 + 1] = q_sq; mj[OWG_MET_T4 + 2] = q_re1; mj[OWG_MET_T4 + 3] = q_im1; mj[OWG_MET_T4 + 4] = q_re2; mj[OWG_MET_T4 + 5] = q_im2; }
    }
    if (cw && t_stop < we.n_max) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) cw[(k++) * 32] = st.v[i];
        cw[(k++) * 32] = st.i_nl[0]; cw[(k++) * 32] = st.i_nl[1]; cw[(k++) * 32] = st.v_nl[0]; cw[(k++) * 32] = st.v_nl[1];
        cw[(k++) * 32] = st.j_cin; cw[(k++) * 32] = st.cin_prev;
#pragma unroll
        for (int i = 0; i < 3; i++) { cw[(k++) * 32] = ua[i]; cw[(k++) * 32] = ub[i]; cw[(k++) * 32] = da[i]; cw[(k++) * 32] = db[i]; }
        cw[(k++) * 32] = down_delay;
        cw[(k++) * 32] = spk.thermal; cw[(k++) * 32] = spk.h1; cw[(k++) * 32] = spk.h2; cw[(k++) * 32] = spk.l1; cw[(k++) * 32] = spk.l2;
        cw[(k++) * 32] = g_prev;
    }
    if (diag && is_main) {
        for (int i = 0; i < 7; i++) if (hist[i]) atomicAdd(&diag->main_hist[i], (unsigned long long)hist[i]);
        for (int i = 0; i < 9; i++) if (pa_hist[i]) atomicAdd(&diag->pa_hist[i], (unsigned long long)pa_hist[i]);
        if (nan_resets) atomicAdd(&diag->main_nan, (unsigned long long)nan_resets);
    }
}

}  // namespace owgd
