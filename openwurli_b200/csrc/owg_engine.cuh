// Chain E on the device: WurliEngine streams (crates/openwurli-dsp/src/engine.rs) -- 64 voice slots per engine with
// stealing + 5 ms linear crossfade, sustain pedal, per-block voice clean-up, parameter smoothers, shared mono chain
// with the power amp inside the (optionally 2x oversampled) preamp loop, f32 output.
//
// Mapping: one thread per engine stream (streams are independent; config 5 has 16 384 of them).  Everything that
// does not depend on the stream's notes -- Twin-T/LDR trajectory incl. the depth smoother, per-sample DK matrices,
// the zero-input shadow solve, the warm-up -- is computed once per group by the shared kernels and read by all
// engine threads.  Voice run-time state lives in a per-engine pool in global memory and is loaded into registers
// for one render() block at a time.
#pragma once
#include "owg_kernels.cuh"

namespace owgd {

struct VoiceRT {  // run-time state of one Voice (voice.rs:16-23: reed + pickup + noise + gain)
    double s[7], c[7], env[7], drift[7];
    double cos_inc[7], sin_inc[7], phase_inc[7], amp[7], decay[7];
    double damper_rate[7], damper_mult[7];
    double revert, diffusion, onset_inc, onset_exp;
    double n_amp, n_decay, b0, b1, b2, a1, a2, z1, z2;
    double q, beta, ds, gain;
    double damper_ramp_samples, damper_release_count;
    unsigned long long sample, onset_n;
    uint32_t jit, n_rng, n_left, n_total;
    uint8_t midi, damper_active, damper_ramp_done, _pad[5];
};

__device__ __forceinline__ void voice_from_init(VoiceRT& v, const OwgVoiceInit& vi) {
#pragma unroll
    for (int m = 0; m < 7; m++) {
        v.s[m] = 0.0; v.c[m] = 1.0; v.env[m] = 1.0; v.drift[m] = vi.jitter_drift[m];
        v.cos_inc[m] = vi.cos_inc[m]; v.sin_inc[m] = vi.sin_inc[m]; v.phase_inc[m] = vi.phase_inc[m];
        v.amp[m] = vi.amplitude[m]; v.decay[m] = vi.decay_mult[m];
        v.damper_rate[m] = 0.0; v.damper_mult[m] = 1.0;
    }
    v.revert = vi.jitter_revert; v.diffusion = vi.jitter_diffusion; v.onset_inc = vi.onset_ramp_inc; v.onset_exp = vi.onset_shape_exp;
    v.n_amp = vi.noise_amp; v.n_decay = vi.noise_decay; v.b0 = vi.bq_b0; v.b1 = vi.bq_b1; v.b2 = vi.bq_b2; v.a1 = vi.bq_a1; v.a2 = vi.bq_a2;
    v.z1 = 0.0; v.z2 = 0.0; v.q = 1.0; v.beta = vi.pickup_beta; v.ds = vi.pickup_ds; v.gain = vi.post_pickup_gain;
    v.damper_ramp_samples = 0.0; v.damper_release_count = 0.0;
    v.sample = 0ull; v.onset_n = vi.onset_ramp_samples;
    v.jit = vi.jitter_state; v.n_rng = vi.noise_rng; v.n_left = vi.noise_remaining; v.n_total = vi.noise_remaining;
    v.midi = vi.midi; v.damper_active = 0; v.damper_ramp_done = 0;
}

// Voice::render for `len` samples (voice.rs:162-179 with reed.rs:219-306 incl. the 3-phase damper), accumulated into
// acc[t*stride] in the engine's summation order: acc += sample (main voice) or acc += sample * fade(t) (steal voice,
// engine.rs:480-488).  fade_start/fade_len < 0 selects the plain sum.  Returns false if a non-finite sample was produced.
__device__ __noinline__ bool voice_render_block(VoiceRT* __restrict__ vp, double* __restrict__ acc, long long stride, int len,
                                                long long fade_start, double fade_len) {
    double s[7], c[7], env[7], drift[7];
#pragma unroll
    for (int m = 0; m < 7; m++) { s[m] = vp->s[m]; c[m] = vp->c[m]; env[m] = vp->env[m]; drift[m] = vp->drift[m]; }
    const double revert = vp->revert, diffusion = vp->diffusion, onset_inc = vp->onset_inc, onset_exp = vp->onset_exp;
    const int onset_mode = onset_exp <= 1.001 ? 0 : (onset_exp >= 1.999 ? 1 : 2);
    const unsigned long long onset_n = vp->onset_n;
    unsigned long long smp = vp->sample;
    uint32_t jit = vp->jit, n_rng = vp->n_rng, n_left = vp->n_left;
    const uint32_t n_total = vp->n_total;
    double n_amp = vp->n_amp, z1 = vp->z1, z2 = vp->z2, q = vp->q;
    const double n_decay = vp->n_decay, b0 = vp->b0, b1 = vp->b1, b2 = vp->b2, a1 = vp->a1, a2 = vp->a2;
    const double beta = vp->beta, ds = vp->ds, gain = vp->gain;
    bool damper_active = vp->damper_active != 0, ramp_done = vp->damper_ramp_done != 0;
    double release_count = vp->damper_release_count;
    const double ramp = vp->damper_ramp_samples;
    bool all_finite = true;
    for (int t = 0; t < len; t++) {
        if (damper_active) {  // reed.rs:227-247
            release_count += 1.0;
            if (!ramp_done) {
                if (release_count > ramp) ramp_done = true;
                else {
#pragma unroll
                    for (int m = 0; m < 7; m++) env[m] *= exp(-(vp->damper_rate[m] * release_count / ramp));
                }
            }
            if (ramp_done) {
#pragma unroll
                for (int m = 0; m < 7; m++) env[m] *= vp->damper_mult[m];
            }
        }
        double onset = 1.0;
        if (smp < onset_n) {
            const double cosine = 0.5 * (1.0 - cos((double)smp * onset_inc));
            onset = onset_mode == 0 ? cosine : (onset_mode == 1 ? cosine * cosine : pow(cosine, onset_exp));
        }
        if ((smp & 15ull) == 0ull) {
#pragma unroll
            for (int m = 0; m < 7; m++) {
                jit = jit * 1664525u + 1013904223u;
                const double u = (double)(jit >> 1) / (4294967295.0 / 2.0);
                drift[m] = revert * drift[m] + diffusion * ((u * 2.0 - 1.0) * 1.7320508080);
            }
        }
        double sum = 0.0;
#pragma unroll
        for (int m = 0; m < 7; m++) {
            sum += vp->amp[m] * s[m] * onset * env[m];
            const double dp = drift[m] * vp->phase_inc[m];
            const double ci = vp->cos_inc[m] - dp * vp->sin_inc[m];
            const double si = vp->sin_inc[m] + dp * vp->cos_inc[m];
            const double s_new = s[m] * ci + c[m] * si;
            const double c_new = c[m] * ci - s[m] * si;
            s[m] = s_new; c[m] = c_new;
            env[m] *= vp->decay[m];
        }
        if ((smp & 1023ull) == 0ull && smp > 0ull) {
#pragma unroll
            for (int m = 0; m < 7; m++) {
                const double r_inv = 1.0 / sqrt(s[m] * s[m] + c[m] * c[m]);
                s[m] *= r_inv; c[m] *= r_inv;
            }
        }
        double x = 0.0 + sum;
        if (n_left > 0u) {
            const uint32_t played = n_total - n_left;
            const double envn = played < 16u ? c_noise_fade[played] : 1.0;
            n_rng = n_rng * 1664525u + 1013904223u;
            const double white = (double)(int32_t)n_rng / 2147483647.0;
            const double y = b0 * white + z1;
            z1 = b1 * white - a1 * y + z2;
            z2 = b2 * white - a2 * y;
            x += n_amp * envn * y;
            n_amp *= n_decay;
            n_left -= 1u;
        }
        double yy = x * ds;
        {
            const double ay = fabs(yy);
            if (!(ay < 0.94)) yy = copysign(0.94 + (0.98 - 0.94) * tanh((ay - 0.94) / (0.98 - 0.94)), yy);
        }
        const double omy = 1.0 - yy;
        const double alpha = beta * omy;
        q = (q * (1.0 - alpha) + 2.0 * beta) / (1.0 + alpha);
        const double out = ((q * omy - 1.0) * 1.8375) * gain;
        all_finite = all_finite && finite64(out);
        if (fade_start >= 0) {  // steal voice: gain = saturating_sub(steal_fade, i) / fade_len  (engine.rs:483-486)
            const long long rem = fade_start - (long long)t;
            acc[(long long)t * stride] += out * ((double)(rem > 0 ? rem : 0) / fade_len);
        } else acc[(long long)t * stride] += out;
        smp += 1ull;
    }
#pragma unroll
    for (int m = 0; m < 7; m++) { vp->s[m] = s[m]; vp->c[m] = c[m]; vp->env[m] = env[m]; vp->drift[m] = drift[m]; }
    vp->sample = smp; vp->jit = jit; vp->n_rng = n_rng; vp->n_left = n_left; vp->n_amp = n_amp; vp->z1 = z1; vp->z2 = z2; vp->q = q;
    vp->damper_active = damper_active ? 1 : 0; vp->damper_ramp_done = ramp_done ? 1 : 0; vp->damper_release_count = release_count;
    return all_finite;
}

// Voice::is_silent (voice.rs:183-188, reed.rs:309-314); thr = 10^(-80/20) evaluated on the host.
__device__ __forceinline__ bool voice_is_silent(const VoiceRT& v, double sample_rate, double thr) {
    if (v.damper_active && v.damper_release_count / sample_rate > 10.0) return true;
#pragma unroll
    for (int m = 0; m < 7; m++) if (!(fabs(v.amp[m] * v.env[m]) <= thr)) return false;
    return true;
}

__device__ __forceinline__ void voice_note_off(VoiceRT& v, const DamperRow* __restrict__ rows) {  // voice.rs:156-158 -> reed.rs:191-216
    const DamperRow& r = rows[v.midi];
    if (!r.enabled) return;
#pragma unroll
    for (int m = 0; m < 7; m++) { v.damper_rate[m] = r.rate[m]; v.damper_mult[m] = r.mult[m]; }
    v.damper_ramp_samples = r.ramp_samples;
    v.damper_active = 1;
    v.damper_release_count = 0.0;
    v.damper_ramp_done = 0;
}

// ---- shared per-group sequences for engines -------------------------------------------------------------------------
// Tremolo::new(0.5, os_sr) [2 s settle] then one process() per preamp-rate sample with the depth trajectory of the engine's
// LinearSmoother (engine.rs:67-130, 532-547): 0.5 during the warm-up, then a ramp_samples-long linear ramp to the target.
// Output: pot_0_resistance in effect per preamp-rate sample (warm-up first).
__global__ void engine_tremolo_kernel(const EngineGroup* groups, int n_groups, double* pot_seq, long long pot_stride) {
    const int gi = blockIdx.x;
    if (gi >= n_groups || threadIdx.x != 0) return;
    const EngineGroup gr = groups[gi];
    const double sr = gr.preamp_sr;
    __shared__ TrmMats m;
    __shared__ TrmK kq;
    __shared__ double trm_sc[OWG_TRM_SCRATCH];
    kq = trm_consts();
    trm_defaults(m);
    TrmState st;
    for (int i = 0; i < TN; i++) st.v[i] = TRM_DC_OP[i];
    for (int i = 0; i < TM; i++) { st.il[i] = TRM_DC_NL_I[i]; st.ilpp[i] = TRM_DC_NL_I[i]; }
    st.xin_prev = 0.0;
    const double ldr_attack = exp(-1.0 / (0.0025 * sr));
    const double ldr_release = exp(-1.0 / (0.035 * sr));
    const double ln_r_max = log(1000000.0);
    const double ln_min_minus_max = log(9000.0) - log(1000000.0);
    double env = 0.0;
    double pot = 9.99999999999999854e4;
    // depth smoother (LinearSmoother): current = target = 0.5 until set_tremolo_depth(target) after the warm-up
    double sm_current = 0.5, sm_target = 0.5, sm_step = 0.0;
    uint32_t sm_remaining = 0;
    double depth = 0.5;  // Tremolo::new(0.5, ..) stores 0.5; set_depth() clamps to [0,1]
    double* o = pot_seq + (size_t)gi * pot_stride;
    const double tot = sr * 2.0;
    const long long n_settle = !(tot == tot) || tot <= 0.0 ? 0ll : (long long)tot;
    const long long n_pre = 50 + n_settle;
    const long long n_live = gr.n_warm_os + gr.n_os;
    const int sub = gr.oversample ? 2 : 1;
    for (long long n = 0; n < n_pre + n_live; n++) {
        if (n == 50 && fabs(sr - 48000.0) > 0.5) trm_rebuild(m, sr * 1.0);
        const bool live = n >= n_pre;
        if (live) {
            const long long tl = n - n_pre;
            if (tl == gr.n_warm_os) {  // set_tremolo_depth(target): LinearSmoother::set_target (engine.rs:85-98)
                if (!(fabs(gr.depth_target - sm_target) < 1e-9)) {
                    sm_target = gr.depth_target;
                    const double delta = sm_target - sm_current;
                    if (gr.ramp_samples == 0) { sm_current = sm_target; sm_remaining = 0; }
                    else { sm_step = delta / (double)gr.ramp_samples; sm_remaining = (uint32_t)gr.ramp_samples; }
                }
            }
            if ((tl % sub) == 0) {  // once per base-rate sample: depth = smoother.next(); tremolo.set_depth(depth)
                if (sm_remaining > 0) {
                    sm_current += sm_step;
                    sm_remaining -= 1;
                    if (sm_remaining == 0) sm_current = sm_target;
                }
                depth = rclamp(sm_current, 0.0, 1.0);
            }
        }
        const double v_out = trm_step(st, m, kq, nullptr, trm_sc);
        if (live) {
            const double led = rclamp((10.95 - v_out) / (10.95 - 0.70), 0.0, 1.0);
            const double coeff = led > env ? ldr_attack : ldr_release;
            env = led + coeff * (env - led);
            const double drive = rclamp(env, 0.0, 1.0);
            double r_ldr;
            if (drive < 1e-6) r_ldr = 1000000.0;
            else r_ldr = exp(ln_r_max + ln_min_minus_max * pow(drive, 0.9));
            const double r_upper = 50000.0 * (1.0 - depth);
            const double r_lower = 50000.0 * depth;
            const double top = r_upper > 0.0 ? r_upper * 18000.0 / (r_upper + 18000.0) : 0.0;
            const double branch = 680.0 + r_ldr;
            const double low = r_lower > 0.0 ? r_lower * branch / (r_lower + branch) : 0.0;
            const double z = top + low;
            if (finite64(z)) {
                const double r = rclamp(z, 1.0e3, 1.0e6);
                if (!(fabs(r - pot) < 1e-12)) pot = r;
            }
            o[n - n_pre] = pot;
        }
    }
}

// One thread per (group, preamp-rate sample): rebuild_matrices for that sample's pot value (all samples are dirty: the
// first set_ldr_resistance moves the pot off the settled 100 kOhm).
__global__ void engine_matrix_kernel(const EngineGroup* groups, int n_groups, const double* pot_seq, long long pot_stride,
                                     double* recs, long long rec_stride_t, double* ans) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int gi = blockIdx.y;
    if (gi >= n_groups) return;
    const EngineGroup gr = groups[gi];
    if (t == 0) {
        double rec[OWG_MAT_STRIDE];
        if (gr.use_defaults) dk_default_record(rec, ans + (size_t)gi * OWG_AN_SPARSE);
        else dk_rebuild(gr.preamp_sr, 9.99999999999999854e4, rec, ans + (size_t)gi * OWG_AN_SPARSE);
    }
    if (t >= gr.n_warm_os + gr.n_os) return;
    const double pot = pot_seq[(size_t)gi * pot_stride + t];
    double* rec = recs + ((size_t)gi * rec_stride_t + t) * OWG_MAT_STRIDE;
    if (gr.use_defaults && pot == 9.99999999999999854e4) {
        bool all_same = true;
        for (long long u = 0; u <= t && all_same; u++) all_same = pot_seq[(size_t)gi * pot_stride + u] == pot;
        if (all_same) { dk_default_record(rec, nullptr); return; }
    }
    dk_rebuild(gr.preamp_sr, pot, rec, nullptr);
}

// Zero-input solve per group over warm-up + render: during the warm-up main and shadow are the same computation, so
// this also yields every engine's main state at the end of the warm-up (post_warm).  pump[t] covers the rendered part.
__global__ void engine_shadow_kernel(const EngineGroup* groups, int n_groups, const DkState* settled, const double* recs,
                                     long long rec_stride_t, const double* ans, double* pump, long long pump_stride, DkState* post_warm) {
    const int gi = blockIdx.x;
    if (gi >= n_groups || threadIdx.x != 0) return;
    __shared__ double s_an[OWG_AN_SPARSE];
    __shared__ double cold[OWG_COLD_SCRATCH];
    const EngineGroup gr = groups[gi];
    for (int e = 0; e < OWG_AN_SPARSE; e++) s_an[e] = ans[(size_t)gi * OWG_AN_SPARSE + e];
    DkState st = *settled;
    const DkDev dv = dk_dev();
    const double* grec = recs + (size_t)gi * rec_stride_t * OWG_MAT_STRIDE;
    const long long n_live = gr.n_warm_os + gr.n_os;
    if (gr.n_warm_os == 0) post_warm[gi] = st;
    for (long long t = 0; t < n_live; t++) {
        const double* m = grec + (size_t)t * OWG_MAT_STRIDE;
        const double y = dk_step<false>(0.0, st, m, s_an, m[OWG_MAT_AN66], dv, nullptr, cold, 1);
        if (t >= gr.n_warm_os) pump[(size_t)gi * pump_stride + (t - gr.n_warm_os)] = y;
        if (t + 1 == gr.n_warm_os) post_warm[gi] = st;
    }
}

// ---- the engine kernel: one thread per WurliEngine stream ----------------------------------------------------------------
struct EngineDiag { unsigned long long nan_guard, out_nan, steals, note_ons, voices_freed, max_active; };

#define OWG_SLOT_FREE 0
#define OWG_SLOT_HELD 1
#define OWG_SLOT_SUSTAINED 2
#define OWG_SLOT_RELEASING 3

__global__ void __launch_bounds__(32) engine_kernel(const EngineDesc* __restrict__ engines, int n_engines, const EngineEvent* __restrict__ events,
                                                    const OwgVoiceInit* __restrict__ vinits, const DamperRow* __restrict__ dampers /*[sched][128]*/,
                                                    const int32_t* __restrict__ damper_sched, const SpkUpdate* __restrict__ spk_updates,
                                                    const long long* __restrict__ spk_offsets, const EngineGroup* __restrict__ groups,
                                                    const DkState* __restrict__ post_warm, const double* __restrict__ recs, long long rec_stride_t,
                                                    const double* __restrict__ ans, const double* __restrict__ pump, long long pump_stride,
                                                    VoiceRT* __restrict__ pool /*[engine][128]*/, double* __restrict__ scratch /*[max_block][n_engines]*/,
                                                    double silent_thr, float* __restrict__ out, long long out_stride, long long max_samples, EngineDiag* diag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_engines) return;
    __shared__ double s_cold[OWG_COLD_SCRATCH * 32];
    const EngineDesc ed = engines[e];
    const EngineGroup gr = groups[ed.group];
    const double sr = ed.sample_rate;
    VoiceRT* mypool = pool + (size_t)e * 128;
    double* acc = scratch + e;                 // acc[t * n_engines]: lane-contiguous
    const long long acc_stride = n_engines;
    const DamperRow* drows = dampers + (size_t)damper_sched[e] * 128;
    const double* grec = recs + (size_t)ed.group * rec_stride_t * OWG_MAT_STRIDE;
    const double* gan = ans + (size_t)ed.group * OWG_AN_SPARSE;
    const double* gpump = pump + (size_t)ed.group * pump_stride;
    const SpkUpdate* sched = spk_updates + spk_offsets[ed.spk_sched];

    // slot table (engine.rs:36-61)
    uint8_t st_state[64], st_note[64], st_cur[64], st_has_voice[64], st_has_steal[64];
    unsigned long long st_age[64];
    uint32_t st_fade[64], st_fade_len[64];
    for (int i = 0; i < 64; i++) { st_state[i] = OWG_SLOT_FREE; st_note[i] = 0; st_cur[i] = 0; st_has_voice[i] = 0; st_has_steal[i] = 0; st_age[i] = 0; st_fade[i] = 0; st_fade_len[i] = 0; }
    unsigned long long age_counter = 0;
    bool sustain_held = false;

    // shared mono chain state
    DkState dk = post_warm[ed.group];
    const DkDev dv = dk_dev();
    double ua[3] = {0, 0, 0}, ub[3] = {0, 0, 0}, da[3] = {0, 0, 0}, db[3] = {0, 0, 0}, down_delay = 0.0;
    SpkState spk = {0.0, 0.0, 0.0, 0.0, 0.0};
    OwgChainInit sc;  // speaker coefficients in effect
    sc.spk_a2 = 0.0; sc.spk_a3 = 0.0; sc.spk_norm = 1.0; sc.spk_thermal_coeff = 0.0; sc.spk_thermal_alpha = 1.0 / (5.0 * sr);
    sc.spk_tanh = 0;
    int spk_next = 0;
    long long spk_clock = ed.n_warm;  // the schedule counts from the first render() sample incl. the warm-up
    while (spk_next < ed.n_spk_updates && sched[spk_next].at < spk_clock) {  // updates that happened during the warm-up
        const SpkUpdate& u = sched[spk_next++];
        sc.spk_a2 = u.a2; sc.spk_a3 = u.a3; sc.spk_norm = u.norm; sc.spk_thermal_coeff = u.thermal_coeff; sc.spk_tanh = u.tanh_on;
        sc.hpf_b0 = u.hpf_b0; sc.hpf_b1 = u.hpf_b1; sc.hpf_b2 = u.hpf_b2; sc.hpf_a1 = u.hpf_a1; sc.hpf_a2 = u.hpf_a2;
        sc.lpf_b0 = u.lpf_b0; sc.lpf_b1 = u.lpf_b1; sc.lpf_b2 = u.lpf_b2; sc.lpf_a1 = u.lpf_a1; sc.lpf_a2 = u.lpf_a2;
    }
    // volume smoother: current = target = 0.5, then set_volume(target) before the first rendered block
    double vol_current = 0.5, vol_target = 0.5, vol_step = 0.0;
    uint32_t vol_remaining = 0;
    if (!(fabs(ed.volume_target - vol_target) < 1e-9)) {
        vol_target = ed.volume_target;
        const double delta = vol_target - vol_current;
        if (ed.ramp_samples == 0) vol_current = vol_target;
        else { vol_step = delta / (double)ed.ramp_samples; vol_remaining = (uint32_t)ed.ramp_samples; }
    }
    unsigned long long d_nan_guard = 0, d_out_nan = 0, d_steals = 0, d_note_ons = 0, d_freed = 0, d_max_active = 0;
    const uint32_t fade_samples = (uint32_t)fmin(fmax(sr * 0.005, 0.0), 4294967295.0);  // (sample_rate * 0.005) as u32
    long long ev = ed.ev_begin;
    long long tos = 0;
    float* o = out + (size_t)e * out_stride;

    for (long long pos = 0; pos < ed.n_samples; pos += ed.block_size) {
        const int len = (int)((ed.n_samples - pos) < (long long)ed.block_size ? (ed.n_samples - pos) : (long long)ed.block_size);
        // ---- events that fall in this block (applied at its start) ----
        while (ev < ed.ev_end && events[ev].sample < pos + len) {
            const EngineEvent evv = events[ev++];
            if (evv.kind == OWG_EV_NOTE_ON) {  // engine.rs:299-338
                const uint8_t note = (uint8_t)evv.note;
                for (int i = 0; i < 64; i++) {
                    if (st_state[i] == OWG_SLOT_SUSTAINED && st_note[i] == note) {
                        st_state[i] = OWG_SLOT_RELEASING;
                        if (st_has_voice[i]) voice_note_off(mypool[2 * i + st_cur[i]], drows);
                    }
                }
                int best = 0;
                {  // allocate_voice, engine.rs:569-590
                    unsigned long long best_pr = 0xFFFFFFFFFFFFFFFFull;
                    bool found_free = false;
                    for (int i = 0; i < 64 && !found_free; i++) {
                        unsigned long long pr;
                        if (st_state[i] == OWG_SLOT_FREE) { best = i; found_free = true; break; }
                        else if (st_state[i] == OWG_SLOT_RELEASING) pr = st_age[i];
                        else if (st_state[i] == OWG_SLOT_SUSTAINED) pr = st_age[i] + 0xFFFFFFFFFFFFFFFFull / 4;
                        else pr = st_age[i] + 0xFFFFFFFFFFFFFFFFull / 2;
                        if (pr < best_pr) { best_pr = pr; best = i; }
                    }
                }
                if (st_state[best] != OWG_SLOT_FREE) {  // steal: the running voice becomes the fading steal voice
                    st_has_steal[best] = st_has_voice[best];
                    st_cur[best] ^= 1;                   // new voice goes to the other pool entry (drops any older steal voice)
                    st_fade[best] = fade_samples;
                    st_fade_len[best] = fade_samples;
                    d_steals++;
                } else if (st_has_steal[best]) {
                    // Free slot that still fades a stolen voice: the new voice must not overwrite it
                    // (slot.voice is None here, so the steal voice keeps living in the other entry)
                }
                age_counter += 1;
                // pick the pool entry: st_cur points at the entry of `voice`; the steal voice (if any) lives in the other one
                voice_from_init(mypool[2 * best + st_cur[best]], vinits[evv.vinit]);
                st_has_voice[best] = 1;
                st_state[best] = OWG_SLOT_HELD;
                st_note[best] = note;
                st_age[best] = age_counter;
                d_note_ons++;
            } else if (evv.kind == OWG_EV_NOTE_OFF) {  // engine.rs:340-359
                const uint8_t note = (uint8_t)evv.note;
                int oldest = -1;
                for (int i = 0; i < 64; i++)
                    if (st_state[i] == OWG_SLOT_HELD && st_note[i] == note && (oldest < 0 || st_age[i] < st_age[oldest])) oldest = i;
                if (oldest >= 0) {
                    if (sustain_held) st_state[oldest] = OWG_SLOT_SUSTAINED;
                    else {
                        st_state[oldest] = OWG_SLOT_RELEASING;
                        if (st_has_voice[oldest]) voice_note_off(mypool[2 * oldest + st_cur[oldest]], drows);
                    }
                }
            } else if (evv.kind == OWG_EV_SUSTAIN) {  // engine.rs:361-374
                const bool held = evv.note != 0;
                if (sustain_held && !held) {
                    for (int i = 0; i < 64; i++) {
                        if (st_state[i] == OWG_SLOT_SUSTAINED) {
                            st_state[i] = OWG_SLOT_RELEASING;
                            if (st_has_voice[i]) voice_note_off(mypool[2 * i + st_cur[i]], drows);
                        }
                    }
                }
                sustain_held = held;
            }
        }
        // ---- render_voices_to_preamp_out (engine.rs:466-521) ----
        for (int t = 0; t < len; t++) acc[(long long)t * acc_stride] = 0.0;
        bool all_finite = true;
        unsigned long long active = 0;
        for (int i = 0; i < 64; i++) {
            if (st_state[i] == OWG_SLOT_FREE && !st_has_steal[i]) continue;
            if (st_has_voice[i]) {
                all_finite = voice_render_block(&mypool[2 * i + st_cur[i]], acc, acc_stride, len, -1, 1.0) && all_finite;
                active++;
            }
            if (st_has_steal[i]) {
                all_finite = voice_render_block(&mypool[2 * i + (st_cur[i] ^ 1)], acc, acc_stride, len, (long long)st_fade[i], (double)st_fade_len[i]) && all_finite;
                st_fade[i] = st_fade[i] > (uint32_t)len ? st_fade[i] - (uint32_t)len : 0u;
                if (st_fade[i] == 0) st_has_steal[i] = 0;
            }
        }
        if (active > d_max_active) d_max_active = active;
        {
            bool sum_finite = true;
            for (int t = 0; t < len; t++) sum_finite = sum_finite && finite64(acc[(long long)t * acc_stride]);
            if (!sum_finite) {  // NaN guard, engine.rs:499-521 (re-renders every voice to find the culprit)
                d_nan_guard++;
                for (int i = 0; i < 64; i++) {
                    if (st_state[i] == OWG_SLOT_FREE && !st_has_steal[i]) continue;
                    if (st_has_voice[i]) {
                        for (int t = 0; t < len; t++) acc[(long long)t * acc_stride] = 0.0;
                        if (!voice_render_block(&mypool[2 * i + st_cur[i]], acc, acc_stride, len, -1, 1.0)) { st_state[i] = OWG_SLOT_FREE; st_has_voice[i] = 0; }
                    }
                    if (st_has_steal[i]) {
                        for (int t = 0; t < len; t++) acc[(long long)t * acc_stride] = 0.0;
                        if (!voice_render_block(&mypool[2 * i + (st_cur[i] ^ 1)], acc, acc_stride, len, -1, 1.0)) { st_has_steal[i] = 0; st_fade[i] = 0; }
                    }
                }
                for (int t = 0; t < len; t++) acc[(long long)t * acc_stride] = 0.0;
            }
        }
        (void)all_finite;
        // ---- shared chain (engine.rs:523-566) then speaker / volume / f32 (engine.rs:436-459) ----
        for (int t = 0; t < len; t++) {
            const double x = acc[(long long)t * acc_stride];
            double stage_out;
            if (ed.oversample) {
                const double u0 = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, ua, x);
                const double u1 = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, ub, x);
                double p0 = 0.0, p1 = 0.0;
#pragma unroll 1
                for (int j = 0; j < 2; j++) {
                    const double* m = grec + (size_t)(gr.n_warm_os + tos) * OWG_MAT_STRIDE;
                    const double main_out = dk_step<false>(j == 0 ? u0 : u1, dk, m, gan, m[OWG_MAT_AN66], dv, nullptr, s_cold + threadIdx.x, 32);
                    double res = main_out - gpump[tos];
                    if (!finite64(res)) { dk = post_warm[ed.group]; res = 0.0; }
                    const double pa = poweramp(res * 0.25, nullptr);
                    if (j == 0) p0 = pa; else p1 = pa;
                    tos += 1;
                }
                const double a = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, da, p0);
                const double b = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, db, p1);
                stage_out = (a + down_delay) * 0.5;
                down_delay = b;
            } else {
                const double* m = grec + (size_t)(gr.n_warm_os + tos) * OWG_MAT_STRIDE;
                const double main_out = dk_step<false>(x, dk, m, gan, m[OWG_MAT_AN66], dv, nullptr, s_cold + threadIdx.x, 32);
                double res = main_out - gpump[tos];
                if (!finite64(res)) { dk = post_warm[ed.group]; res = 0.0; }
                stage_out = poweramp(res * 0.25, nullptr);
                tos += 1;
            }
            while (spk_next < ed.n_spk_updates && sched[spk_next].at <= spk_clock) {  // set_character() -> update_coefficients()
                const SpkUpdate& u = sched[spk_next++];
                sc.spk_a2 = u.a2; sc.spk_a3 = u.a3; sc.spk_norm = u.norm; sc.spk_thermal_coeff = u.thermal_coeff; sc.spk_tanh = u.tanh_on;
                sc.hpf_b0 = u.hpf_b0; sc.hpf_b1 = u.hpf_b1; sc.hpf_b2 = u.hpf_b2; sc.hpf_a1 = u.hpf_a1; sc.hpf_a2 = u.hpf_a2;
                sc.lpf_b0 = u.lpf_b0; sc.lpf_b1 = u.lpf_b1; sc.lpf_b2 = u.lpf_b2; sc.lpf_a1 = u.lpf_a1; sc.lpf_a2 = u.lpf_a2;
            }
            spk_clock += 1;
            const double shaped = speaker(stage_out, spk, sc);
            if (vol_remaining > 0) {
                vol_current += vol_step;
                vol_remaining -= 1;
                if (vol_remaining == 0) vol_current = vol_target;
            }
            const float smp = (float)(shaped * 7.498942093324558 * vol_current);
            if (isfinite(smp)) o[pos + t] = smp;
            else {  // engine.rs:449-458: reset chain, emit 0 (the shared shadow cannot be reset per engine; counted in diag)
                d_out_nan++;
                dk = post_warm[ed.group];
                for (int k = 0; k < 3; k++) { ua[k] = ub[k] = da[k] = db[k] = 0.0; }
                down_delay = 0.0;
                spk.thermal = spk.h1 = spk.h2 = spk.l1 = spk.l2 = 0.0;
                o[pos + t] = 0.0f;
            }
        }
        // ---- cleanup_voices (engine.rs:592-602) ----
        for (int i = 0; i < 64; i++) {
            if (st_state[i] != OWG_SLOT_FREE && st_has_voice[i] && voice_is_silent(mypool[2 * i + st_cur[i]], sr, silent_thr)) {
                st_state[i] = OWG_SLOT_FREE;
                st_has_voice[i] = 0;
                d_freed++;
            }
        }
    }
    for (long long t = ed.n_samples; t < max_samples; t++) o[t] = 0.0f;  // ragged batch: rows end in silence
    if (diag) {
        atomicAdd(&diag->nan_guard, d_nan_guard); atomicAdd(&diag->out_nan, d_out_nan); atomicAdd(&diag->steals, d_steals);
        atomicAdd(&diag->note_ons, d_note_ons); atomicAdd(&diag->voices_freed, d_freed); atomicMax(&diag->max_active, d_max_active);
    }
}

}  // namespace owgd
