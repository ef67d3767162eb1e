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
#include "owg_legacy.cuh"

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
                    const Recip rr = recip_prepare(ramp);  // seven quotients, one divisor: bit-identical to `/` (owg_device.cuh)
#pragma unroll
                    for (int m = 0; m < 7; m++) env[m] *= exp(-div_by(vp->damper_rate[m] * release_count, rr));
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
// One entry of an engine's schedule (host-prepared, owg_records.h): kind 0 = Speaker::update_coefficients (the host simulated the
// character smoother and set_character's 0.002 dead-band), kind 1 = WurliEngine::set_volume -> LinearSmoother::set_target
// (engine.rs:85-98, 378-380) at the start of the block the event falls in.
__device__ __forceinline__ void sched_apply(const SpkUpdate& u, OwgChainInit& sc, double& vol_current, double& vol_target, double& vol_step,
                                            uint32_t& vol_remaining, const int ramp_samples) {
    if (u._pad == 1) {
        const double target = u.a2;
        if (!(fabs(target - vol_target) < 1e-9)) {
            vol_target = target;
            const double delta = target - vol_current;
            if (ramp_samples == 0) { vol_current = target; vol_remaining = 0; }
            else { vol_step = delta / (double)ramp_samples; vol_remaining = (uint32_t)ramp_samples; }
        }
        return;
    }
    sc.spk_a2 = u.a2; sc.spk_a3 = u.a3; sc.spk_norm = u.norm; sc.spk_thermal_coeff = u.thermal_coeff; sc.spk_tanh = u.tanh_on;
    sc.hpf_b0 = u.hpf_b0; sc.hpf_b1 = u.hpf_b1; sc.hpf_b2 = u.hpf_b2; sc.hpf_a1 = u.hpf_a1; sc.hpf_a2 = u.hpf_a2;
    sc.lpf_b0 = u.lpf_b0; sc.lpf_b1 = u.lpf_b1; sc.lpf_b2 = u.lpf_b2; sc.lpf_a1 = u.lpf_a1; sc.lpf_a2 = u.lpf_a2;
}

// LinearSmoother (engine.rs:67-130, 532-547): 0.5 during the warm-up, then a ramp_samples-long linear ramp to the target.
// Output: pot_0_resistance in effect per preamp-rate sample (warm-up first).
struct EngTrmRun {  // oscillator state between chunk launches
    TrmState st;
    long long n_done;  // steps done so far (50 warm-up + 2*sr settle + live)
};
struct EngLdrRun {  // LDR envelope, resistance tracking and depth-smoother state between chunk launches
    double env, pot, sm_current, sm_target, sm_step, depth;
    uint32_t sm_remaining, _pad;
    long long n_done;  // live samples done
};

// live_end_base: produce oscillator output for live (warm-up + rendered) preamp-rate samples below n_warm_os + live_end_base*sub.
__global__ void engine_tremolo_kernel(const EngineGroup* groups, int n_groups, double* pot_seq, long long pot_stride, EngTrmRun* run,
                                      long long live_end_base) {
    const int gi = blockIdx.x;
    if (gi >= n_groups || threadIdx.x != 0) return;
    const EngineGroup gr = groups[gi];
    const double sr = gr.preamp_sr;
    __shared__ TrmMats m;
    __shared__ TrmK kq;
    __shared__ double trm_sc[OWG_TRM_SCRATCH];
    kq = trm_consts();
    trm_defaults(m);
    const double tot = sr * 2.0;
    const long long n_settle = !(tot == tot) || tot <= 0.0 ? 0ll : (long long)tot;
    const long long n_pre = 50 + n_settle;
    const long long n_live = gr.n_warm_os + gr.n_os;
    const int sub = gr.oversample ? 2 : 1;
    long long live_end = gr.n_warm_os + live_end_base * sub;
    if (live_end > n_live) live_end = n_live;
    const long long n_begin = run[gi].n_done, n_end = n_pre + live_end;
    if (n_begin >= n_end) return;
    TrmState st;
    if (n_begin == 0) {
        for (int i = 0; i < TN; i++) st.v[i] = TRM_DC_OP[i];
        for (int i = 0; i < TM; i++) { st.il[i] = TRM_DC_NL_I[i]; st.ilpp[i] = TRM_DC_NL_I[i]; }
        st.xin_prev = 0.0;
    } else {
        st = run[gi].st;
        if (n_begin > 50 && fabs(sr - 48000.0) > 0.5) trm_rebuild(m, sr * 1.0);  // same deterministic rebuild as at step 50
    }
    double* o = pot_seq + (size_t)gi * pot_stride;
    for (long long n = n_begin; n < n_end; n++) {
        if (n == 50 && fabs(sr - 48000.0) > 0.5) trm_rebuild(m, sr * 1.0);
        const double v_out = trm_step(st, m, kq, nullptr, trm_sc);
        if (n >= n_pre) o[n - n_pre] = v_out;
    }
    run[gi].st = st;
    run[gi].n_done = n_end;
}

// Tremolo::process after the oscillator + the engine's depth smoother (engine.rs:67-130, 532-547: 0.5 during the warm-up, then a
// ramp_samples-long linear ramp to the target) + the preamp's resistance tracking; in place on seq (volts in, pot_0_resistance
// out), `dseq` is scratch for the depth trajectory.  One CTA per group; serial recurrences on thread 0, maps on all threads.
struct DepthEv { long long at_os; double target; };  // WurliEngine::set_tremolo_depth at the start of a block (preamp-rate index incl. warm-up)
__global__ void __launch_bounds__(256) engine_ldr_kernel(const EngineGroup* groups, int n_groups, double* seq, double* dseq, long long seq_stride,
                                                         EngLdrRun* run, long long live_end_base, int legacy, const DepthEv* __restrict__ dep_ev) {
    // legacy != 0: the consumer is the 8-node legacy preamp (set_ldr_resistance: max(R, 1000), 0.01 Ohm threshold); output = g_ldr
    const int gi = blockIdx.x;
    if (gi >= n_groups) return;
    const EngineGroup gr = groups[gi];
    const double sr = gr.preamp_sr;
    const int sub = gr.oversample ? 2 : 1;
    const long long n_live = gr.n_warm_os + gr.n_os;
    long long t1 = gr.n_warm_os + live_end_base * sub;
    if (t1 > n_live) t1 = n_live;
    const long long t0 = run[gi].n_done;
    if (t0 >= t1) return;
    double* o = seq + (size_t)gi * seq_stride;
    double* dd = dseq + (size_t)gi * seq_stride;
    const int tid = threadIdx.x, nth = blockDim.x;
    for (long long t = t0 + tid; t < t1; t += nth) o[t] = rclamp((10.95 - o[t]) / (10.95 - 0.70), 0.0, 1.0);
    __syncthreads();
    if (tid == 0) {
        const double ldr_attack = exp(-1.0 / (0.0025 * sr));
        const double ldr_release = exp(-1.0 / (0.035 * sr));
        EngLdrRun R = run[gi];
        if (t0 == 0) {
            R.env = 0.0;
            // depth smoother (LinearSmoother): current = target = 0.5 until set_tremolo_depth(target) after the warm-up
            R.sm_current = 0.5; R.sm_target = 0.5; R.sm_step = 0.0; R.sm_remaining = 0;
            R.depth = 0.5;  // Tremolo::new(0.5, ..) stores 0.5; set_depth() clamps to [0,1]
        }
        double env = R.env, sm_current = R.sm_current, sm_target = R.sm_target, sm_step = R.sm_step, depth = R.depth;
        uint32_t sm_remaining = R.sm_remaining;
        int dep_next = gr.dep_ev_begin;
        while (dep_next < gr.dep_ev_end && dep_ev[dep_next].at_os < t0) dep_next++;  // events of earlier chunks are already applied
        for (long long tl = t0; tl < t1; tl++) {
            // set_tremolo_depth(target): LinearSmoother::set_target (engine.rs:85-98, 382-384) -- the construction-time target right after
            // the warm-up, then the stream's automation events at the start of their blocks
            auto set_target = [&](const double target) {
                if (!(fabs(target - sm_target) < 1e-9)) {
                    sm_target = target;
                    const double delta = sm_target - sm_current;
                    if (gr.ramp_samples == 0) { sm_current = sm_target; sm_remaining = 0; }
                    else { sm_step = delta / (double)gr.ramp_samples; sm_remaining = (uint32_t)gr.ramp_samples; }
                }
            };
            if (tl == gr.n_warm_os) set_target(gr.depth_target);
            while (dep_next < gr.dep_ev_end && dep_ev[dep_next].at_os == tl) set_target(dep_ev[dep_next++].target);
            if ((tl % sub) == 0) {  // once per base-rate sample: depth = smoother.next(); tremolo.set_depth(depth)
                if (sm_remaining > 0) {
                    sm_current += sm_step;
                    sm_remaining -= 1;
                    if (sm_remaining == 0) sm_current = sm_target;
                }
                depth = rclamp(sm_current, 0.0, 1.0);
            }
            const double led = o[tl];
            const double coeff = led > env ? ldr_attack : ldr_release;
            env = led + coeff * (env - led);
            o[tl] = env;
            dd[tl] = depth;
        }
        R.env = env; R.sm_current = sm_current; R.sm_target = sm_target; R.sm_step = sm_step; R.depth = depth; R.sm_remaining = sm_remaining;
        run[gi].env = R.env; run[gi].sm_current = R.sm_current; run[gi].sm_target = R.sm_target; run[gi].sm_step = R.sm_step;
        run[gi].depth = R.depth; run[gi].sm_remaining = R.sm_remaining;
    }
    __syncthreads();
    {
        const double ln_r_max = log(1000000.0);
        const double ln_min_minus_max = log(9000.0) - log(1000000.0);
        for (long long t = t0 + tid; t < t1; t += nth) {
            const double drive = rclamp(o[t], 0.0, 1.0);
            const double depth = dd[t];
            double r_ldr;
            if (drive < 1e-6) r_ldr = 1000000.0;
            else r_ldr = exp(ln_r_max + ln_min_minus_max * pow(drive, 0.9));
            const double r_upper = 50000.0 * (1.0 - depth);
            const double r_lower = 50000.0 * depth;
            const double top = r_upper > 0.0 ? r_upper * 18000.0 / (r_upper + 18000.0) : 0.0;
            const double branch = 680.0 + r_ldr;
            const double low = r_lower > 0.0 ? r_lower * branch / (r_lower + branch) : 0.0;
            o[t] = top + low;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double pot = t0 == 0 ? (legacy ? 1000000.0 : 9.99999999999999854e4) : run[gi].pot;
        for (long long t = t0; t < t1; t++) {
            const double z = o[t];
            if (legacy) {
                const double r = z > 1000.0 ? z : 1000.0;
                if (fabs(r - pot) > 0.01) pot = r;
                o[t] = 1.0 / pot;
            } else {
                if (finite64(z)) {
                    const double r = rclamp(z, 1.0e3, 1.0e6);
                    if (!(fabs(r - pot) < 1e-12)) pot = r;
                }
                o[t] = pot;
            }
        }
        run[gi].pot = pot;
        run[gi].n_done = t1;
    }
}

// One thread per (group, preamp-rate sample): rebuild_matrices for that sample's pot value (all samples are dirty: the
// first set_ldr_resistance moves the pot off the settled 100 kOhm).
__global__ void engine_matrix_kernel(const EngineGroup* groups, int n_groups, const double* pot_seq, long long pot_stride,
                                     double* recs, long long rec_stride_t, double* ans, long long live_begin_base, long long live_end_base) {
    const int gi = blockIdx.y;
    if (gi >= n_groups) return;
    const EngineGroup gr = groups[gi];
    const int sub = gr.oversample ? 2 : 1;
    // chunk = live samples [begin, end): the first chunk also covers the warm-up
    const long long begin = live_begin_base <= 0 ? 0 : gr.n_warm_os + live_begin_base * sub;
    const long long end = gr.n_warm_os + live_end_base * sub;
    const long long t = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= end) return;
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

// Zero-input solve per group over the warm-up: main and shadow are the same computation there, so this yields both
// states at the end of the warm-up (post_warm).  The rendered part of the shadow runs in lane 31 of the chain warps.
__global__ void engine_shadow_kernel(const EngineGroup* groups, int n_groups, const DkState* settled, const double* recs,
                                     long long rec_stride_t, const double* ans, DkState* post_warm) {
    const int gi = blockIdx.x;
    if (gi >= n_groups || threadIdx.x != 0) return;
    __shared__ double s_an[OWG_AN_SPARSE];
    __shared__ double cold[OWG_COLD_SCRATCH];
    const EngineGroup gr = groups[gi];
    for (int e = 0; e < OWG_AN_SPARSE; e++) s_an[e] = ans[(size_t)gi * OWG_AN_SPARSE + e];
    DkState st = *settled;
    const DkDev dv = dk_dev();
    const double* grec = recs + (size_t)gi * rec_stride_t * OWG_MAT_STRIDE;
    for (long long t = 0; t < gr.n_warm_os; t++) {
        const double* m = grec + (size_t)t * OWG_MAT_STRIDE;
        (void)dk_step<false>(0.0, st, m, s_an, m[OWG_MAT_AN66], dv, nullptr, cold, 1);
    }
    post_warm[gi] = st;
}

// ---- WurliEngine streams: voice-parallel block pipeline ---------------------------------------------------------------
// The slot state machine (allocation, stealing, sustain, clean-up) and the NaN guard depend only on the voices' own state,
// never on the shared chain, so a stream is rendered as
//   per render() block ("round"):  engine_events_kernel   thread / engine : events of the block -> slot table -> render list
//                                  engine_voice_mix_kernel  CTA / engine, thread / voice: Voice::render into a shared-memory
//                                                           tile, ordered (slot-order) sum -> mix[engine][t]  (f64)
//                                  engine_post_kernel     thread / engine : NaN guard (rare), cleanup_voices
//   per segment of rounds:         engine_chain_kernel    lane / engine  : oversampler, DK preamp - pump, power amp, speaker,
//                                                           volume smoother, f32 -- on its own stream, one segment behind.
struct EngineDiag { unsigned long long nan_guard, out_nan, steals, note_ons, voices_freed, max_active; };

#define OWG_SLOT_FREE 0
#define OWG_SLOT_HELD 1
#define OWG_SLOT_SUSTAINED 2
#define OWG_SLOT_RELEASING 3
#define OWG_ENGINE_TILE 16   // samples per shared-memory tile of the voice/mix kernel
#define OWG_ENGINE_ITEMS 64  // threads (= render-list items per pass) of the voice/mix kernel

struct EngineState {  // slot table (engine.rs:36-61) + the current round's render list
    unsigned long long st_age[64];
    uint32_t st_fade[64], st_fade_len[64];
    uint8_t st_state[64], st_note[64], st_cur[64], st_has_voice[64], st_has_steal[64];
    unsigned long long age_counter;
    long long ev;               // next event of this engine
    int32_t sustain_held, n_items, nan_flag, _pad;
    int32_t item_fade[128];     // steal_fade at the start of the block for steal voices, -1 for slot voices
    uint32_t item_fade_len[128];
    uint8_t item_pool[128];     // pool entry of item k (items are in the engine's summation order)
    unsigned long long d_steals, d_note_ons, d_freed, d_max_active, d_nan_guard;
};

struct EngineChainState {  // shared mono chain of one engine between segments
    DkState dk;   // melange 12-node preamp
    LgState lg;   // legacy 8-node preamp
    double g_prev;
    double ua[3], ub[3], da[3], db[3], down_delay;
    SpkState spk;
    OwgChainInit sc;
    double vol_current, vol_target, vol_step;
    uint32_t vol_remaining;
    int32_t spk_next;
    long long spk_clock;
    unsigned long long d_out_nan;
};

struct EngineWarp { int32_t group, first, count, block_size; long long n_max; };  // engines of one (group, block size) in lanes 0..count-1

__global__ void engine_init_kernel(const EngineDesc* __restrict__ engines, int n_engines, const EngineGroup* __restrict__ groups,
                                   const DkState* __restrict__ post_warm, EngineState* __restrict__ states, EngineChainState* __restrict__ chains,
                                   const LgState* __restrict__ post_warm_lg, const double* __restrict__ g_warm_last) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_engines) return;
    const EngineDesc ed = engines[e];
    if (states) {
    EngineState& S = states[e];
    for (int i = 0; i < 64; i++) { S.st_state[i] = OWG_SLOT_FREE; S.st_note[i] = 0; S.st_cur[i] = 0; S.st_has_voice[i] = 0; S.st_has_steal[i] = 0; S.st_age[i] = 0; S.st_fade[i] = 0; S.st_fade_len[i] = 0; }
    S.age_counter = 0; S.ev = ed.ev_begin; S.sustain_held = 0; S.n_items = 0; S.nan_flag = 0;
    S.d_steals = S.d_note_ons = S.d_freed = S.d_max_active = S.d_nan_guard = 0;
    }
    if (chains) {
        EngineChainState& C = chains[e];
        if (post_warm_lg) { C.lg = post_warm_lg[ed.group]; C.g_prev = g_warm_last[ed.group]; }
        else C.dk = post_warm[ed.group];
        for (int k = 0; k < 3; k++) { C.ua[k] = C.ub[k] = C.da[k] = C.db[k] = 0.0; }
        C.down_delay = 0.0;
        C.spk.thermal = C.spk.h1 = C.spk.h2 = C.spk.l1 = C.spk.l2 = 0.0;
        C.sc.spk_a2 = 0.0; C.sc.spk_a3 = 0.0; C.sc.spk_norm = 1.0; C.sc.spk_thermal_coeff = 0.0; C.sc.spk_thermal_alpha = 1.0 / (5.0 * ed.sample_rate);
        C.sc.spk_tanh = 0;
        C.sc.hpf_b0 = C.sc.hpf_b1 = C.sc.hpf_b2 = C.sc.hpf_a1 = C.sc.hpf_a2 = 0.0;
        C.sc.lpf_b0 = C.sc.lpf_b1 = C.sc.lpf_b2 = C.sc.lpf_a1 = C.sc.lpf_a2 = 0.0;
        C.spk_next = 0;
        C.spk_clock = ed.n_warm;  // the schedule counts from the first render() sample incl. the warm-up
        // volume smoother: current = target = 0.5, then set_volume(target) before the first rendered block
        C.vol_current = 0.5; C.vol_target = 0.5; C.vol_step = 0.0; C.vol_remaining = 0;
        if (!(fabs(ed.volume_target - C.vol_target) < 1e-9)) {
            C.vol_target = ed.volume_target;
            const double delta = C.vol_target - C.vol_current;
            if (ed.ramp_samples == 0) C.vol_current = C.vol_target;
            else { C.vol_step = delta / (double)ed.ramp_samples; C.vol_remaining = (uint32_t)ed.ramp_samples; }
        }
        C.d_out_nan = 0;
    }
}

// Events that fall in round r's block (applied at its start, like the host loop around WurliEngine::render), then the
// block's render list in the summation order of render_voices_to_preamp_out (engine.rs:466-493).
__global__ void engine_events_kernel(const EngineDesc* __restrict__ engines, int n_engines, long long round, const EngineEvent* __restrict__ events,
                                     const OwgVoiceInit* __restrict__ vinits, const DamperRow* __restrict__ dampers /*[sched][128]*/,
                                     const int32_t* __restrict__ damper_sched, VoiceRT* __restrict__ pool /*[engine][128]*/,
                                     EngineState* __restrict__ states) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_engines) return;
    const EngineDesc ed = engines[e];
    const long long pos = round * (long long)ed.block_size;
    if (pos >= ed.n_samples) return;
    const int len = (int)((ed.n_samples - pos) < (long long)ed.block_size ? (ed.n_samples - pos) : (long long)ed.block_size);
    EngineState& S = states[e];
    VoiceRT* mypool = pool + (size_t)e * 128;
    const DamperRow* drows = dampers + (size_t)damper_sched[e] * 128;
    const uint32_t fade_samples = (uint32_t)fmin(fmax(ed.sample_rate * 0.005, 0.0), 4294967295.0);  // (sample_rate * 0.005) as u32
    long long ev = S.ev;
    while (ev < ed.ev_end && events[ev].sample < pos + len) {
        const EngineEvent evv = events[ev++];
        if (evv.kind == OWG_EV_NOTE_ON) {  // engine.rs:299-338
            const uint8_t note = (uint8_t)evv.note;
            for (int i = 0; i < 64; i++) {
                if (S.st_state[i] == OWG_SLOT_SUSTAINED && S.st_note[i] == note) {
                    S.st_state[i] = OWG_SLOT_RELEASING;
                    if (S.st_has_voice[i]) voice_note_off(mypool[2 * i + S.st_cur[i]], drows);
                }
            }
            int best = 0;
            {  // allocate_voice, engine.rs:569-590
                unsigned long long best_pr = 0xFFFFFFFFFFFFFFFFull;
                for (int i = 0; i < 64; i++) {
                    unsigned long long pr;
                    const uint8_t stt = S.st_state[i];
                    if (stt == OWG_SLOT_FREE) { best = i; break; }
                    else if (stt == OWG_SLOT_RELEASING) pr = S.st_age[i];
                    else if (stt == OWG_SLOT_SUSTAINED) pr = S.st_age[i] + 0xFFFFFFFFFFFFFFFFull / 4;
                    else pr = S.st_age[i] + 0xFFFFFFFFFFFFFFFFull / 2;
                    if (pr < best_pr) { best_pr = pr; best = i; }
                }
            }
            if (S.st_state[best] != OWG_SLOT_FREE) {  // steal: the running voice becomes the fading steal voice
                S.st_has_steal[best] = S.st_has_voice[best];
                S.st_cur[best] ^= 1;                   // new voice goes to the other pool entry (drops any older steal voice)
                S.st_fade[best] = fade_samples;
                S.st_fade_len[best] = fade_samples;
                S.d_steals++;
            }
            // a Free slot may still fade a stolen voice: it lives in the other pool entry and is left alone
            S.age_counter += 1;
            voice_from_init(mypool[2 * best + S.st_cur[best]], vinits[evv.vinit]);
            S.st_has_voice[best] = 1;
            S.st_state[best] = OWG_SLOT_HELD;
            S.st_note[best] = note;
            S.st_age[best] = S.age_counter;
            S.d_note_ons++;
        } else if (evv.kind == OWG_EV_NOTE_OFF) {  // engine.rs:340-359
            const uint8_t note = (uint8_t)evv.note;
            int oldest = -1;
            for (int i = 0; i < 64; i++)
                if (S.st_state[i] == OWG_SLOT_HELD && S.st_note[i] == note && (oldest < 0 || S.st_age[i] < S.st_age[oldest])) oldest = i;
            if (oldest >= 0) {
                if (S.sustain_held) S.st_state[oldest] = OWG_SLOT_SUSTAINED;
                else {
                    S.st_state[oldest] = OWG_SLOT_RELEASING;
                    if (S.st_has_voice[oldest]) voice_note_off(mypool[2 * oldest + S.st_cur[oldest]], drows);
                }
            }
        } else if (evv.kind == OWG_EV_SUSTAIN) {  // engine.rs:361-374
            const bool held = evv.note != 0;
            if (S.sustain_held && !held) {
                for (int i = 0; i < 64; i++) {
                    if (S.st_state[i] == OWG_SLOT_SUSTAINED) {
                        S.st_state[i] = OWG_SLOT_RELEASING;
                        if (S.st_has_voice[i]) voice_note_off(mypool[2 * i + S.st_cur[i]], drows);
                    }
                }
            }
            S.sustain_held = held ? 1 : 0;
        }
    }
    S.ev = ev;
    // render list + the fade bookkeeping that follows each steal voice's render (engine.rs:480-492)
    int n_items = 0;
    unsigned long long active = 0;
    for (int i = 0; i < 64; i++) {
        if (S.st_state[i] == OWG_SLOT_FREE && !S.st_has_steal[i]) continue;
        if (S.st_has_voice[i]) {
            S.item_pool[n_items] = (uint8_t)(2 * i + S.st_cur[i]); S.item_fade[n_items] = -1; S.item_fade_len[n_items] = 1u;
            n_items++; active++;
        }
        if (S.st_has_steal[i]) {
            S.item_pool[n_items] = (uint8_t)(2 * i + (S.st_cur[i] ^ 1)); S.item_fade[n_items] = (int32_t)S.st_fade[i]; S.item_fade_len[n_items] = S.st_fade_len[i];
            n_items++;
            S.st_fade[i] = S.st_fade[i] > (uint32_t)len ? S.st_fade[i] - (uint32_t)len : 0u;
            if (S.st_fade[i] == 0) S.st_has_steal[i] = 0;
        }
    }
    S.n_items = n_items;
    if (active > S.d_max_active) S.d_max_active = active;
}

// Register-resident Voice (reed + attack noise + pickup + gain); one sample per call, same arithmetic as voice_kernel /
// voice_render_block (voice.rs:162-179, reed.rs:219-306, hammer.rs:150-179, pickup.rs:130-149).
// The 35 per-mode constants (cos_inc, sin_inc, phase_inc, amplitude, decay) live in shared memory, one column per thread
// (kb[(a * 7 + m) * OWG_ENGINE_ITEMS], a = 0..4), which halves the register footprint and doubles the resident warps: the kernel is
// bound by the latency of one voice sample, so throughput scales with the warps per scheduler.
#define OWG_VK(a, m) kb[((a) * 7 + (m)) * OWG_ENGINE_ITEMS]
struct VoiceRegs {
    double s[7], c[7], env[7], drift[7];
    double revert, diffusion, onset_inc, onset_exp, n_amp, n_decay, b0, b1, b2, a1, a2, z1, z2, q, beta, ds, gain, ramp, release_count;
    unsigned long long smp, onset_n;
    uint32_t jit, n_rng, n_left, n_total;
    int onset_mode;
    bool damper_active, ramp_done;
};

__device__ __forceinline__ void voice_regs_load(VoiceRegs& r, const VoiceRT* __restrict__ vp, double* __restrict__ kb) {
#pragma unroll
    for (int m = 0; m < 7; m++) {
        r.s[m] = vp->s[m]; r.c[m] = vp->c[m]; r.env[m] = vp->env[m]; r.drift[m] = vp->drift[m];
        OWG_VK(0, m) = vp->cos_inc[m]; OWG_VK(1, m) = vp->sin_inc[m]; OWG_VK(2, m) = vp->phase_inc[m]; OWG_VK(3, m) = vp->amp[m]; OWG_VK(4, m) = vp->decay[m];
    }
    r.revert = vp->revert; r.diffusion = vp->diffusion; r.onset_inc = vp->onset_inc; r.onset_exp = vp->onset_exp;
    r.onset_mode = r.onset_exp <= 1.001 ? 0 : (r.onset_exp >= 1.999 ? 1 : 2);
    r.onset_n = vp->onset_n; r.smp = vp->sample; r.jit = vp->jit; r.n_rng = vp->n_rng; r.n_left = vp->n_left; r.n_total = vp->n_total;
    r.n_amp = vp->n_amp; r.z1 = vp->z1; r.z2 = vp->z2; r.q = vp->q;
    r.n_decay = vp->n_decay; r.b0 = vp->b0; r.b1 = vp->b1; r.b2 = vp->b2; r.a1 = vp->a1; r.a2 = vp->a2;
    r.beta = vp->beta; r.ds = vp->ds; r.gain = vp->gain;
    r.damper_active = vp->damper_active != 0; r.ramp_done = vp->damper_ramp_done != 0;
    r.release_count = vp->damper_release_count; r.ramp = vp->damper_ramp_samples;
}

__device__ __forceinline__ void voice_regs_store(const VoiceRegs& r, VoiceRT* __restrict__ vp) {
#pragma unroll
    for (int m = 0; m < 7; m++) { vp->s[m] = r.s[m]; vp->c[m] = r.c[m]; vp->env[m] = r.env[m]; vp->drift[m] = r.drift[m]; }
    vp->sample = r.smp; vp->jit = r.jit; vp->n_rng = r.n_rng; vp->n_left = r.n_left; vp->n_amp = r.n_amp; vp->z1 = r.z1; vp->z2 = r.z2; vp->q = r.q;
    vp->damper_active = r.damper_active ? 1 : 0; vp->damper_ramp_done = r.ramp_done ? 1 : 0; vp->damper_release_count = r.release_count;
}

__device__ __forceinline__ double voice_sample(VoiceRegs& r, const VoiceRT* __restrict__ vp, const double* __restrict__ kb) {
    if (r.damper_active) {  // reed.rs:227-247
        r.release_count += 1.0;
        if (!r.ramp_done) {
            if (r.release_count > r.ramp) r.ramp_done = true;
            else {
                const Recip rr = recip_prepare(r.ramp);  // seven quotients, one divisor: bit-identical to `/` (owg_device.cuh)
#pragma unroll
                for (int m = 0; m < 7; m++) r.env[m] *= exp(-div_by(vp->damper_rate[m] * r.release_count, rr));
            }
        }
        if (r.ramp_done) {
#pragma unroll
            for (int m = 0; m < 7; m++) r.env[m] *= vp->damper_mult[m];
        }
    }
    double onset = 1.0;
    if (r.smp < r.onset_n) {
        const double cosine = 0.5 * (1.0 - cos((double)r.smp * r.onset_inc));
        onset = r.onset_mode == 0 ? cosine : (r.onset_mode == 1 ? cosine * cosine : pow(cosine, r.onset_exp));
    }
    if ((r.smp & 15ull) == 0ull) {
#pragma unroll
        for (int m = 0; m < 7; m++) {
            r.jit = r.jit * 1664525u + 1013904223u;
            const double u = (double)(r.jit >> 1) / (4294967295.0 / 2.0);
            r.drift[m] = r.revert * r.drift[m] + r.diffusion * ((u * 2.0 - 1.0) * 1.7320508080);
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int m = 0; m < 7; m++) {
        const double k_cos = OWG_VK(0, m), k_sin = OWG_VK(1, m);
        sum += OWG_VK(3, m) * r.s[m] * onset * r.env[m];
        const double dp = r.drift[m] * OWG_VK(2, m);
        const double ci = k_cos - dp * k_sin;
        const double si = k_sin + dp * k_cos;
        const double s_new = r.s[m] * ci + r.c[m] * si;
        const double c_new = r.c[m] * ci - r.s[m] * si;
        r.s[m] = s_new; r.c[m] = c_new;
        r.env[m] *= OWG_VK(4, m);
    }
    if ((r.smp & 1023ull) == 0ull && r.smp > 0ull) {
#pragma unroll
        for (int m = 0; m < 7; m++) {
            const double r_inv = 1.0 / sqrt(r.s[m] * r.s[m] + r.c[m] * r.c[m]);
            r.s[m] *= r_inv; r.c[m] *= r_inv;
        }
    }
    double x = 0.0 + sum;
    if (r.n_left > 0u) {
        const uint32_t played = r.n_total - r.n_left;
        const double envn = played < 16u ? c_noise_fade[played] : 1.0;
        r.n_rng = r.n_rng * 1664525u + 1013904223u;
        const double white = (double)(int32_t)r.n_rng / 2147483647.0;
        const double y = r.b0 * white + r.z1;
        r.z1 = r.b1 * white - r.a1 * y + r.z2;
        r.z2 = r.b2 * white - r.a2 * y;
        x += r.n_amp * envn * y;
        r.n_amp *= r.n_decay;
        r.n_left -= 1u;
    }
    double yy = x * r.ds;
    {
        const double ay = fabs(yy);
        if (!(ay < 0.94)) yy = copysign(0.94 + (0.98 - 0.94) * tanh((ay - 0.94) / (0.98 - 0.94)), yy);
    }
    const double omy = 1.0 - yy;
    const double alpha = r.beta * omy;
    r.q = (r.q * (1.0 - alpha) + 2.0 * r.beta) / (1.0 + alpha);
    r.smp += 1ull;
    return ((r.q * omy - 1.0) * 1.8375) * r.gain;
}

// One CTA (64 threads) per engine, one thread per render-list item: each voice renders a 16-sample tile into shared
// memory, then thread j sums column j over the items IN LIST ORDER (sum_buf[i] += voice_buf[i] [* gain], engine.rs:476-488)
// and writes mix[engine][t].  Lists longer than 64 (steal voices) take further passes that continue the running sums.
__global__ void __launch_bounds__(OWG_ENGINE_ITEMS, 8) engine_voice_mix_kernel(const EngineDesc* __restrict__ engines, long long round,
                                                                               VoiceRT* __restrict__ pool, EngineState* __restrict__ states,
                                                                               double* __restrict__ mix, long long mix_stride, long long seg_round0) {
    __shared__ double tile[OWG_ENGINE_ITEMS * (OWG_ENGINE_TILE + 1)];
    __shared__ double s_k[35 * OWG_ENGINE_ITEMS];
    __shared__ int32_t s_fade[OWG_ENGINE_ITEMS];
    __shared__ double s_fade_len[OWG_ENGINE_ITEMS];
    const int e = blockIdx.x;
    const int tid = threadIdx.x;
    const EngineDesc ed = engines[e];
    const long long pos = round * (long long)ed.block_size;
    if (pos >= ed.n_samples) return;
    const int len = (int)((ed.n_samples - pos) < (long long)ed.block_size ? (ed.n_samples - pos) : (long long)ed.block_size);
    EngineState& S = states[e];
    const int n_items = S.n_items;
    double* mrow = mix + (size_t)e * mix_stride + (round - seg_round0) * (long long)ed.block_size;
    if (n_items == 0) {
        for (int t = tid; t < len; t += OWG_ENGINE_ITEMS) mrow[t] = 0.0;
        return;
    }
    double* kb = s_k + tid;
    bool bad = false;
    for (int base = 0; base < n_items; base += OWG_ENGINE_ITEMS) {
        const int k = base + tid;
        const int nk = n_items - base < OWG_ENGINE_ITEMS ? n_items - base : OWG_ENGINE_ITEMS;
        const bool act = k < n_items;
        const bool last_pass = base + OWG_ENGINE_ITEMS >= n_items;
        VoiceRT* vp = pool + (size_t)e * 128 + (act ? S.item_pool[k] : 0);
        VoiceRegs r;
        if (act) { voice_regs_load(r, vp, kb); s_fade[tid] = S.item_fade[k]; s_fade_len[tid] = (double)S.item_fade_len[k]; }
        for (int t0 = 0; t0 < len; t0 += OWG_ENGINE_TILE) {
            const int tl = len - t0 < OWG_ENGINE_TILE ? len - t0 : OWG_ENGINE_TILE;
            if (act) {
                double* row = tile + tid * (OWG_ENGINE_TILE + 1);
                for (int t = 0; t < tl; t++) row[t] = voice_sample(r, vp, kb);
            }
            __syncthreads();
            if (tid < tl) {
                double a = base == 0 ? 0.0 : mrow[t0 + tid];
                const int tb = t0 + tid;  // sample index inside the block
                for (int kk = 0; kk < nk; kk++) {
                    double v = tile[kk * (OWG_ENGINE_TILE + 1) + tid];
                    const int32_t f = s_fade[kk];
                    if (f >= 0) {  // gain = saturating_sub(steal_fade, i) / fade_len (engine.rs:483-486)
                        const int32_t rem = f - tb;
                        v = v * ((double)(rem > 0 ? rem : 0) / s_fade_len[kk]);
                    }
                    a += v;
                }
                mrow[t0 + tid] = a;
                if (last_pass && !finite64(a)) bad = true;
            }
            __syncthreads();
        }
        if (act) voice_regs_store(r, vp);
        __syncthreads();
    }
    if (bad) S.nan_flag = 1;
}

// After the block's voices: the NaN guard (engine.rs:495-521; re-renders every voice to find the culprit) and
// cleanup_voices (engine.rs:592-602).  The chain of this block runs later but reads neither.
__global__ void engine_post_kernel(const EngineDesc* __restrict__ engines, int n_engines, long long round, VoiceRT* __restrict__ pool,
                                   EngineState* __restrict__ states, double* __restrict__ mix, long long mix_stride, long long seg_round0,
                                   double silent_thr) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_engines) return;
    const EngineDesc ed = engines[e];
    const long long pos = round * (long long)ed.block_size;
    if (pos >= ed.n_samples) return;
    const int len = (int)((ed.n_samples - pos) < (long long)ed.block_size ? (ed.n_samples - pos) : (long long)ed.block_size);
    EngineState& S = states[e];
    VoiceRT* mypool = pool + (size_t)e * 128;
    if (S.nan_flag) {
        S.nan_flag = 0;
        S.d_nan_guard++;
        double* acc = mix + (size_t)e * mix_stride + (round - seg_round0) * (long long)ed.block_size;
        for (int i = 0; i < 64; i++) {
            if (S.st_state[i] == OWG_SLOT_FREE && !S.st_has_steal[i]) continue;
            if (S.st_has_voice[i]) {
                for (int t = 0; t < len; t++) acc[t] = 0.0;
                if (!voice_render_block(&mypool[2 * i + S.st_cur[i]], acc, 1, len, -1, 1.0)) { S.st_state[i] = OWG_SLOT_FREE; S.st_has_voice[i] = 0; }
            }
            if (S.st_has_steal[i]) {
                for (int t = 0; t < len; t++) acc[t] = 0.0;
                if (!voice_render_block(&mypool[2 * i + (S.st_cur[i] ^ 1)], acc, 1, len, -1, 1.0)) { S.st_has_steal[i] = 0; S.st_fade[i] = 0; }
            }
        }
        for (int t = 0; t < len; t++) acc[t] = 0.0;
    }
    for (int i = 0; i < 64; i++) {
        if (S.st_state[i] != OWG_SLOT_FREE && S.st_has_voice[i] && voice_is_silent(mypool[2 * i + S.st_cur[i]], ed.sample_rate, silent_thr)) {
            S.st_state[i] = OWG_SLOT_FREE;
            S.st_has_voice[i] = 0;
            S.d_freed++;
        }
    }
}

// Shared mono chain for rounds [round0, round1) (engine.rs:523-566 then 436-459).  One warp = up to 31 engines of one
// (group, block size) plus the group's zero-input shadow solve in lane 31 (melange_adapter.rs:72-81); the per-sample DK
// records are staged in shared memory with cp.async one step ahead, as in chain_kernel<TREM>.
__global__ void __launch_bounds__(32) engine_chain_kernel(const EngineWarp* __restrict__ warps, const int32_t* __restrict__ order,
                                                          const EngineDesc* __restrict__ engines, long long round0, long long round1,
                                                          const SpkUpdate* __restrict__ spk_updates, const long long* __restrict__ spk_offsets,
                                                          const EngineGroup* __restrict__ groups, const DkState* __restrict__ post_warm,
                                                          const double* __restrict__ recs, long long rec_stride_t, const double* __restrict__ ans,
                                                          EngineChainState* __restrict__ chains, DkState* __restrict__ shadow_states /*[warp]*/,
                                                          const double* __restrict__ mix, long long mix_stride, float* __restrict__ out,
                                                          long long out_stride, long long max_samples, int last_segment) {
    __shared__ __align__(16) double s_rec[2 * OWG_MAT_STRIDE];
    __shared__ double s_an[OWG_AN_SPARSE];
    __shared__ OwgChainInit s_ci[32];
    __shared__ double s_cold[OWG_COLD_SCRATCH * 32];
    const int lane = threadIdx.x;
    const EngineWarp we = warps[blockIdx.x];
    const bool is_shadow = lane == 31;
    const bool is_main = lane < we.count;
    const int e = is_main ? order[we.first + lane] : order[we.first];
    const EngineDesc ed = engines[e];
    const EngineGroup gr = groups[we.group];
    for (int k = lane; k < OWG_AN_SPARSE; k += 32) s_an[k] = ans[(size_t)we.group * OWG_AN_SPARSE + k];
    const double* grec = recs + (size_t)we.group * rec_stride_t * OWG_MAT_STRIDE;
    const long long n_rec = gr.n_warm_os + gr.n_os;
    const long long T0 = round0 * (long long)we.block_size;
    long long T1 = round1 * (long long)we.block_size;
    if (T1 > we.n_max) T1 = we.n_max;
    const long long ns = is_main ? ed.n_samples : 0;
    float* o = out + (size_t)e * out_stride;
    if (T0 < T1) {
        EngineChainState& C = chains[e];
        const SpkUpdate* sched = spk_updates + spk_offsets[ed.spk_sched];
        const double* mrow = mix + (size_t)e * mix_stride;
        DkState dk;
        if (is_shadow) dk = T0 == 0 ? post_warm[we.group] : shadow_states[blockIdx.x];
        else dk = C.dk;
        const DkDev dv = dk_dev();
        double ua[3], ub[3], da[3], db[3];
        for (int k = 0; k < 3; k++) { ua[k] = C.ua[k]; ub[k] = C.ub[k]; da[k] = C.da[k]; db[k] = C.db[k]; }
        double down_delay = C.down_delay;
        SpkState spk = C.spk;
        s_ci[lane] = C.sc;
        __syncwarp();
        OwgChainInit& sc = s_ci[lane];
        int spk_next = C.spk_next;
        long long spk_clock = C.spk_clock;
        double vol_current = C.vol_current;
        double vol_target = C.vol_target, vol_step = C.vol_step;
        uint32_t vol_remaining = C.vol_remaining;
        unsigned long long d_out_nan = 0;
        if (T0 == 0 && is_main) {
            while (spk_next < ed.n_spk_updates && sched[spk_next].at < spk_clock) {  // updates that happened during the warm-up
                sched_apply(sched[spk_next++], sc, vol_current, vol_target, vol_step, vol_remaining, ed.ramp_samples);
            }
        }
        const int n_sub = gr.oversample ? 2 : 1;
        long long tos = gr.n_warm_os + T0 * n_sub;  // record index (records include the warm-up)
        if (tos < n_rec) {
            const double* src = grec + (size_t)tos * OWG_MAT_STRIDE;
            double* dst = s_rec + (tos & 1) * OWG_MAT_STRIDE;
            for (int c = lane; c < OWG_MAT_STRIDE / 2; c += 32) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 2 * c) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        double x_next = (is_main && T0 < ns) ? mrow[0] : 0.0;
        for (long long t = T0; t < T1; t++) {
            const bool live = is_main && t < ns;
            const double x = x_next;
            x_next = (is_main && t + 1 < ns && t + 1 < T1) ? mrow[t + 1 - T0] : 0.0;
            double u0 = x, u1 = 0.0;
            if (n_sub == 2) {
                u0 = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, ua, x);
                u1 = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, ub, x);
            }
            if (!is_main) { u0 = 0.0; u1 = 0.0; }
            double pp0 = 0.0, pp1 = 0.0;
#pragma unroll 1
            for (int j = 0; j < n_sub; j++) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                if (tos + 1 < n_rec) {
                    const double* src = grec + (size_t)(tos + 1) * OWG_MAT_STRIDE;
                    double* dst = s_rec + ((tos + 1) & 1) * OWG_MAT_STRIDE;
                    for (int c = lane; c < OWG_MAT_STRIDE / 2; c += 32) {
                        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 2 * c) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                const double* m = s_rec + (tos & 1) * OWG_MAT_STRIDE;
                const double main_out = dk_step<false>(j == 0 ? u0 : u1, dk, m, s_an, m[OWG_MAT_AN66], dv, nullptr, s_cold + lane, 32);
                const double pump = __shfl_sync(0xffffffffu, main_out, 31);
                double res = main_out - pump;
                if (!finite64(res)) { if (!is_shadow) dk = post_warm[we.group]; res = 0.0; }
                const double pa = poweramp(res * 0.25, nullptr);
                if (j == 0) pp0 = pa; else pp1 = pa;
                tos += 1;
            }
            double stage_out;
            if (n_sub == 2) {
                const double a = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, da, pp0);
                const double b = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, db, pp1);
                stage_out = (a + down_delay) * 0.5;
                down_delay = b;
            } else stage_out = pp0;
            if (live) {
                while (spk_next < ed.n_spk_updates && sched[spk_next].at <= spk_clock) {  // set_character() -> update_coefficients()
                    sched_apply(sched[spk_next++], sc, vol_current, vol_target, vol_step, vol_remaining, ed.ramp_samples);
                }
                spk_clock += 1;
                const double shaped = speaker(stage_out, spk, sc);
                if (vol_remaining > 0) {
                    vol_current += vol_step;
                    vol_remaining -= 1;
                    if (vol_remaining == 0) vol_current = vol_target;
                }
                const float smp = (float)(shaped * 7.498942093324558 * vol_current);
                if (isfinite(smp)) o[t] = smp;
                else {  // engine.rs:449-458: reset chain, emit 0 (the shared shadow cannot be reset per engine; counted in diag)
                    d_out_nan++;
                    dk = post_warm[we.group];
                    for (int k = 0; k < 3; k++) { ua[k] = ub[k] = da[k] = db[k] = 0.0; }
                    down_delay = 0.0;
                    spk.thermal = spk.h1 = spk.h2 = spk.l1 = spk.l2 = 0.0;
                    o[t] = 0.0f;
                }
            }
        }
        if (is_shadow) shadow_states[blockIdx.x] = dk;
        if (is_main) {
            C.dk = dk;
            for (int k = 0; k < 3; k++) { C.ua[k] = ua[k]; C.ub[k] = ub[k]; C.da[k] = da[k]; C.db[k] = db[k]; }
            C.down_delay = down_delay; C.spk = spk; C.sc = sc; C.spk_next = spk_next; C.spk_clock = spk_clock;
            C.vol_current = vol_current; C.vol_remaining = vol_remaining; C.vol_target = vol_target; C.vol_step = vol_step; C.d_out_nan += d_out_nan;
        }
    }
    if (last_segment && is_main)
        for (long long t = ed.n_samples; t < max_samples; t++) o[t] = 0.0f;  // ragged batch: rows end in silence
}

// Warp-specialised form of engine_chain_kernel (same split as chain_split_kernel): warp A = record staging + DK step + main - shadow,
// warp B = mix sample -> upsampler ... power amp (per sub-step) -> downsampler -> speaker schedule / speaker -> volume smoother -> f32.
// At 96 kHz (one DK step per sample) the input/output stages are half of the per-sample latency, so this nearly doubles the chain rate
// of small and medium stream counts.  Bit-identical to engine_chain_kernel except in the (unreachable with finite input) non-finite
// OUTPUT sample guard: the reset of the preamp state reaches warp A up to two samples late (counted in out_nan either way).
__global__ void __launch_bounds__(64) engine_chain_split_kernel(const EngineWarp* __restrict__ warps, const int32_t* __restrict__ order,
                                                                const EngineDesc* __restrict__ engines, long long round0, long long round1,
                                                                const SpkUpdate* __restrict__ spk_updates, const long long* __restrict__ spk_offsets,
                                                                const EngineGroup* __restrict__ groups, const DkState* __restrict__ post_warm,
                                                                const double* __restrict__ recs, long long rec_stride_t, const double* __restrict__ ans,
                                                                EngineChainState* __restrict__ chains, DkState* __restrict__ shadow_states /*[warp]*/,
                                                                const double* __restrict__ mix, long long mix_stride, float* __restrict__ out,
                                                                long long out_stride, long long max_samples, int last_segment) {
    __shared__ __align__(16) double s_rec[2 * OWG_MAT_STRIDE];
    __shared__ double s_an[OWG_AN_SPARSE];
    __shared__ OwgChainInit s_ci[32];
    __shared__ double s_cold[OWG_COLD_SCRATCH * 32];
    __shared__ double s_u[2][2][32], s_p[2][2][32];
    __shared__ int s_swap;
    __shared__ volatile int s_reset[32];  // warp B -> warp A: non-finite output sample, reset the preamp state of this lane
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        s_swap = (int)((wid >> 2) & 1u);
    }
    if (threadIdx.x < 32) s_reset[threadIdx.x] = 0;
    __syncthreads();
    const bool warp_a = (threadIdx.x < 32) != (s_swap != 0);
    const EngineWarp we = warps[blockIdx.x];
    const bool is_shadow = lane == 31;
    const bool is_main = lane < we.count;
    const int e = is_main ? order[we.first + lane] : order[we.first];
    const EngineDesc ed = engines[e];
    const EngineGroup gr = groups[we.group];
    const long long n_rec = gr.n_warm_os + gr.n_os;
    const long long T0 = round0 * (long long)we.block_size;
    long long T1 = round1 * (long long)we.block_size;
    if (T1 > we.n_max) T1 = we.n_max;
    const long long ns = is_main ? ed.n_samples : 0;
    float* o = out + (size_t)e * out_stride;
    const int n_sub = gr.oversample ? 2 : 1;
    if (T0 < T1) {
        EngineChainState& C = chains[e];
        if (warp_a) {
            // ================================ warp A: the DK preamp ================================
            for (int k = lane; k < OWG_AN_SPARSE; k += 32) s_an[k] = ans[(size_t)we.group * OWG_AN_SPARSE + k];
            const double* grec = recs + (size_t)we.group * rec_stride_t * OWG_MAT_STRIDE;
            DkState dk;
            if (is_shadow) dk = T0 == 0 ? post_warm[we.group] : shadow_states[blockIdx.x];
            else dk = C.dk;
            const DkDev dv = dk_dev();
            long long tos = gr.n_warm_os + T0 * n_sub;
            if (tos < n_rec) {
                const double* src = grec + (size_t)tos * OWG_MAT_STRIDE;
                double* dst = s_rec + (tos & 1) * OWG_MAT_STRIDE;
                for (int c = lane; c < OWG_MAT_STRIDE / 2; c += 32) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 2 * c) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            for (long long t = T0; t < T1; t++) {
                const int slot = (int)(t & 1);
                owg_bar_sync(OWG_BAR_UFULL + slot);
                double u0 = s_u[slot][0][lane], u1 = s_u[slot][1][lane];
                if (!is_main) { u0 = 0.0; u1 = 0.0; }
                if (s_reset[lane]) { s_reset[lane] = 0; if (!is_shadow) dk = post_warm[we.group]; }
                double r0 = 0.0, r1 = 0.0;
#pragma unroll 1
                for (int j = 0; j < n_sub; j++) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    if (tos + 1 < n_rec) {
                        const double* src = grec + (size_t)(tos + 1) * OWG_MAT_STRIDE;
                        double* dst = s_rec + ((tos + 1) & 1) * OWG_MAT_STRIDE;
                        for (int c = lane; c < OWG_MAT_STRIDE / 2; c += 32) {
                            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * c);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 2 * c) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    const double* m = s_rec + (tos & 1) * OWG_MAT_STRIDE;
                    const double main_out = dk_step<false>(j == 0 ? u0 : u1, dk, m, s_an, m[OWG_MAT_AN66], dv, nullptr, s_cold + lane, 32);
                    const double pump = __shfl_sync(0xffffffffu, main_out, 31);
                    double res = main_out - pump;
                    if (!finite64(res)) { if (!is_shadow) dk = post_warm[we.group]; res = 0.0; }
                    if (j == 0) r0 = res; else r1 = res;
                    tos += 1;
                }
                s_p[slot][0][lane] = r0;
                s_p[slot][1][lane] = r1;
                owg_bar_arrive(OWG_BAR_PFULL + slot);
            }
            if (is_shadow) shadow_states[blockIdx.x] = dk;
            if (is_main) C.dk = dk;
        } else {
            // ================================ warp B: input and output stages ================================
            const SpkUpdate* sched = spk_updates + spk_offsets[ed.spk_sched];
            const double* mrow = mix + (size_t)e * mix_stride;
            double ua[3], ub[3], da[3], db[3];
            for (int k = 0; k < 3; k++) { ua[k] = C.ua[k]; ub[k] = C.ub[k]; da[k] = C.da[k]; db[k] = C.db[k]; }
            double down_delay = C.down_delay;
            SpkState spk = C.spk;
            s_ci[lane] = C.sc;
            __syncwarp();
            OwgChainInit& sc = s_ci[lane];
            int spk_next = C.spk_next;
            long long spk_clock = C.spk_clock;
            double vol_current = C.vol_current;
            double vol_target = C.vol_target, vol_step = C.vol_step;
            uint32_t vol_remaining = C.vol_remaining;
            unsigned long long d_out_nan = 0;
            if (T0 == 0 && is_main) {
                while (spk_next < ed.n_spk_updates && sched[spk_next].at < spk_clock) {  // updates that happened during the warm-up
                    sched_apply(sched[spk_next++], sc, vol_current, vol_target, vol_step, vol_remaining, ed.ramp_samples);
                }
            }
            auto produce = [&](long long t) {
                const double x = (is_main && t < ns) ? mrow[t - T0] : 0.0;
                double u0 = x, u1 = 0.0;
                if (n_sub == 2) {
                    u0 = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, ua, x);
                    u1 = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, ub, x);
                }
                const int slot = (int)(t & 1);
                s_u[slot][0][lane] = u0;
                s_u[slot][1][lane] = u1;
                owg_bar_arrive(OWG_BAR_UFULL + slot);
            };
            produce(T0);
            if (T0 + 1 < T1) produce(T0 + 1);
            for (long long t = T0; t < T1; t++) {
                const int slot = (int)(t & 1);
                owg_bar_sync(OWG_BAR_PFULL + slot);
                const double r0 = s_p[slot][0][lane], r1 = s_p[slot][1][lane];
                if (t + 2 < T1) produce(t + 2);
                const bool live = is_main && t < ns;
                double stage_out;
                const double pp0 = poweramp(r0 * 0.25, nullptr);
                if (n_sub == 2) {
                    const double pp1 = poweramp(r1 * 0.25, nullptr);
                    const double a = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, da, pp0);
                    const double b = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, db, pp1);
                    stage_out = (a + down_delay) * 0.5;
                    down_delay = b;
                } else stage_out = pp0;
                if (live) {
                    while (spk_next < ed.n_spk_updates && sched[spk_next].at <= spk_clock) {  // set_character() -> update_coefficients()
                        sched_apply(sched[spk_next++], sc, vol_current, vol_target, vol_step, vol_remaining, ed.ramp_samples);
                    }
                    spk_clock += 1;
                    const double shaped = speaker(stage_out, spk, sc);
                    if (vol_remaining > 0) {
                        vol_current += vol_step;
                        vol_remaining -= 1;
                        if (vol_remaining == 0) vol_current = vol_target;
                    }
                    const float smp = (float)(shaped * 7.498942093324558 * vol_current);
                    if (isfinite(smp)) o[t] = smp;
                    else {  // engine.rs:449-458: reset chain, emit 0
                        d_out_nan++;
                        s_reset[lane] = 1;
                        for (int k = 0; k < 3; k++) { ua[k] = ub[k] = da[k] = db[k] = 0.0; }
                        down_delay = 0.0;
                        spk.thermal = spk.h1 = spk.h2 = spk.l1 = spk.l2 = 0.0;
                        o[t] = 0.0f;
                    }
                }
            }
            if (is_main) {
                for (int k = 0; k < 3; k++) { C.ua[k] = ua[k]; C.ub[k] = ub[k]; C.da[k] = da[k]; C.db[k] = db[k]; }
                C.down_delay = down_delay; C.spk = spk; C.sc = sc; C.spk_next = spk_next; C.spk_clock = spk_clock;
                C.vol_current = vol_current; C.vol_remaining = vol_remaining; C.vol_target = vol_target; C.vol_step = vol_step; C.d_out_nan += d_out_nan;
            }
        }
    }
    if (last_segment && is_main && !warp_a)
        for (long long t = ed.n_samples; t < max_samples; t++) o[t] = 0.0f;  // ragged batch: rows end in silence
}

// ---- legacy 8-node preamp variants (owg_opts.preamp_model = OWG_PREAMP_LEGACY8) ---------------------------------------------------
// Warm-up: DkPreamp::new (DC point at 1 MOhm) then n_warm_os zero-input steps under the tremolo's g_ldr sequence; main and shadow
// are the same computation there.  g_last = g_ldr in effect at the last warm-up step (the first rendered step's g_ldr_prev).
__global__ void engine_shadow_legacy_kernel(const EngineGroup* groups, int n_groups, const double* lgrecs, const double* g_seq, long long g_stride,
                                            LgState* post_warm, double* g_last) {
    const int gi = blockIdx.x;
    if (gi >= n_groups || threadIdx.x != 0) return;
    __shared__ double s_m[OWG_LG_STRIDE];
    const EngineGroup gr = groups[gi];
    for (int e = 0; e < OWG_LG_STRIDE; e++) s_m[e] = lgrecs[(size_t)gi * OWG_LG_STRIDE + e];
    LgState st;
    lg_init(st, s_m);
    double g_prev = s_m[OWG_LG_GINIT];
    const double* gs = g_seq + (size_t)gi * g_stride;
    for (long long t = 0; t < gr.n_warm_os; t++) {
        int iters;
        const double g = gs[t];
        (void)lg_step(st, s_m, 0.0, g, g_prev, &iters);
        g_prev = g;
    }
    post_warm[gi] = st;
    g_last[gi] = g_prev;
}

__global__ void __launch_bounds__(32) engine_chain_legacy_kernel(const EngineWarp* __restrict__ warps, const int32_t* __restrict__ order,
                                                                 const EngineDesc* __restrict__ engines, long long round0, long long round1,
                                                                 const SpkUpdate* __restrict__ spk_updates, const long long* __restrict__ spk_offsets,
                                                                 const EngineGroup* __restrict__ groups, const LgState* __restrict__ post_warm,
                                                                 const double* __restrict__ g_warm_last, const double* __restrict__ lgrecs,
                                                                 const double* __restrict__ g_seq, long long g_stride,
                                                                 EngineChainState* __restrict__ chains, LgState* __restrict__ shadow_states /*[warp]*/,
                                                                 const double* __restrict__ mix, long long mix_stride, float* __restrict__ out,
                                                                 long long out_stride, long long max_samples, int last_segment) {
    __shared__ double s_m[OWG_LG_STRIDE];
    __shared__ OwgChainInit s_ci[32];
    const int lane = threadIdx.x;
    const EngineWarp we = warps[blockIdx.x];
    const bool is_shadow = lane == 31;
    const bool is_main = lane < we.count;
    const int e = is_main ? order[we.first + lane] : order[we.first];
    const EngineDesc ed = engines[e];
    const EngineGroup gr = groups[we.group];
    for (int k = lane; k < OWG_LG_STRIDE; k += 32) s_m[k] = lgrecs[(size_t)we.group * OWG_LG_STRIDE + k];
    const double* gs = g_seq + (size_t)we.group * g_stride;
    const long long n_rec = gr.n_warm_os + gr.n_os;
    const long long T0 = round0 * (long long)we.block_size;
    long long T1 = round1 * (long long)we.block_size;
    if (T1 > we.n_max) T1 = we.n_max;
    const long long ns = is_main ? ed.n_samples : 0;
    float* o = out + (size_t)e * out_stride;
    __syncwarp();
    if (T0 < T1) {
        EngineChainState& C = chains[e];
        const SpkUpdate* sched = spk_updates + spk_offsets[ed.spk_sched];
        const double* mrow = mix + (size_t)e * mix_stride;
        LgState st;
        if (is_shadow) st = T0 == 0 ? post_warm[we.group] : shadow_states[blockIdx.x];
        else st = C.lg;
        double g_prev = T0 == 0 ? g_warm_last[we.group] : C.g_prev;  // warp-uniform (all lanes stepped with the same g sequence)
        double ua[3], ub[3], da[3], db[3];
        for (int k = 0; k < 3; k++) { ua[k] = C.ua[k]; ub[k] = C.ub[k]; da[k] = C.da[k]; db[k] = C.db[k]; }
        double down_delay = C.down_delay;
        SpkState spk = C.spk;
        s_ci[lane] = C.sc;
        __syncwarp();
        OwgChainInit& sc = s_ci[lane];
        int spk_next = C.spk_next;
        long long spk_clock = C.spk_clock;
        double vol_current = C.vol_current;
        double vol_target = C.vol_target, vol_step = C.vol_step;
        uint32_t vol_remaining = C.vol_remaining;
        unsigned long long d_out_nan = 0;
        if (T0 == 0 && is_main) {
            while (spk_next < ed.n_spk_updates && sched[spk_next].at < spk_clock) {  // updates that happened during the warm-up
                sched_apply(sched[spk_next++], sc, vol_current, vol_target, vol_step, vol_remaining, ed.ramp_samples);
            }
        }
        const int n_sub = gr.oversample ? 2 : 1;
        long long tos = gr.n_warm_os + T0 * n_sub;
        double x_next = (is_main && T0 < ns) ? mrow[0] : 0.0;
        double g_next = tos < n_rec ? gs[tos] : g_prev;
        for (long long t = T0; t < T1; t++) {
            const bool live = is_main && t < ns;
            const double x = x_next;
            x_next = (is_main && t + 1 < ns && t + 1 < T1) ? mrow[t + 1 - T0] : 0.0;
            double u0 = x, u1 = 0.0;
            if (n_sub == 2) {
                u0 = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, ua, x);
                u1 = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, ub, x);
            }
            if (!is_main) { u0 = 0.0; u1 = 0.0; }
            double pp0 = 0.0, pp1 = 0.0;
#pragma unroll 1
            for (int j = 0; j < n_sub; j++) {
                const double g_ldr = g_next;
                g_next = tos + 1 < n_rec ? gs[tos + 1] : g_ldr;
                int iters;
                const double main_out = lg_step(st, s_m, j == 0 ? u0 : u1, g_ldr, g_prev, &iters);
                g_prev = g_ldr;
                const double pump = __shfl_sync(0xffffffffu, main_out, 31);
                double res = main_out - pump;
                if (!finite64(res)) { if (!is_shadow) st = post_warm[we.group]; res = 0.0; }  // see chain_legacy_kernel: deviation from reset()
                const double pa = poweramp(res * 0.25, nullptr);
                if (j == 0) pp0 = pa; else pp1 = pa;
                tos += 1;
            }
            double stage_out;
            if (n_sub == 2) {
                const double a = allpass3(OWG_OS_A0, OWG_OS_A1, OWG_OS_A2, da, pp0);
                const double b = allpass3(OWG_OS_B0, OWG_OS_B1, OWG_OS_B2, db, pp1);
                stage_out = (a + down_delay) * 0.5;
                down_delay = b;
            } else stage_out = pp0;
            if (live) {
                while (spk_next < ed.n_spk_updates && sched[spk_next].at <= spk_clock) {
                    sched_apply(sched[spk_next++], sc, vol_current, vol_target, vol_step, vol_remaining, ed.ramp_samples);
                }
                spk_clock += 1;
                const double shaped = speaker(stage_out, spk, sc);
                if (vol_remaining > 0) {
                    vol_current += vol_step;
                    vol_remaining -= 1;
                    if (vol_remaining == 0) vol_current = vol_target;
                }
                const float smp = (float)(shaped * 7.498942093324558 * vol_current);
                if (isfinite(smp)) o[t] = smp;
                else {
                    d_out_nan++;
                    st = post_warm[we.group];
                    for (int k = 0; k < 3; k++) { ua[k] = ub[k] = da[k] = db[k] = 0.0; }
                    down_delay = 0.0;
                    spk.thermal = spk.h1 = spk.h2 = spk.l1 = spk.l2 = 0.0;
                    o[t] = 0.0f;
                }
            }
        }
        if (is_shadow) shadow_states[blockIdx.x] = st;
        if (is_main) {
            C.lg = st; C.g_prev = g_prev;
            for (int k = 0; k < 3; k++) { C.ua[k] = ua[k]; C.ub[k] = ub[k]; C.da[k] = da[k]; C.db[k] = db[k]; }
            C.down_delay = down_delay; C.spk = spk; C.sc = sc; C.spk_next = spk_next; C.spk_clock = spk_clock;
            C.vol_current = vol_current; C.vol_remaining = vol_remaining; C.vol_target = vol_target; C.vol_step = vol_step; C.d_out_nan += d_out_nan;
        }
    }
    if (last_segment && is_main)
        for (long long t = ed.n_samples; t < max_samples; t++) o[t] = 0.0f;
}

// ---- `preamp-bench render-midi` voice manager (main.rs:1735-1850), one thread per MIDI stream, one call per 64-sample chunk --------
// Reuses EngineState (st_state = active flag, st_note, st_age; pool entry 2*slot) and the render list consumed by
// engine_voice_mix_kernel.  held[] = the tool's `pedal_held` vector (note-offs deferred while the pedal is down).
__global__ void midi_events_kernel(const EngineDesc* __restrict__ engines, int n_streams, long long round, const MidiEvent* __restrict__ events,
                                   const OwgVoiceInit* __restrict__ vinits, const DamperRow* __restrict__ dampers /*[128]*/,
                                   VoiceRT* __restrict__ pool /*[stream][128]*/, EngineState* __restrict__ states, uint8_t* __restrict__ held,
                                   int32_t* __restrict__ held_count, const long long* __restrict__ held_offset, double silent_thr) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_streams) return;
    const EngineDesc ed = engines[e];
    const long long pos = round * (long long)ed.block_size;
    if (pos >= ed.n_samples) return;
    EngineState& S = states[e];
    VoiceRT* mypool = pool + (size_t)e * 128;
    uint8_t* myheld = held + held_offset[e];
    int nheld = held_count[e];
    long long ev = S.ev;
    auto note_off_oldest = [&](uint8_t note) {  // the oldest ACTIVE slot of that key (min_by_key(age)); its voice gets note_off()
        int best = -1;
        for (int i = 0; i < 64; i++)
            if (S.st_state[i] && S.st_note[i] == note && (best < 0 || S.st_age[i] < S.st_age[best])) best = i;
        if (best >= 0) voice_note_off(mypool[2 * best], dampers);
    };
    while (ev < ed.ev_end && events[ev].chunk <= round) {
        const MidiEvent evv = events[ev++];
        if (evv.kind == OWG_MIDI_NOTE_ON) {
            S.age_counter += 1;
            int slot = -1;
            for (int i = 0; i < 64; i++) if (!S.st_state[i]) { slot = i; break; }       // position(|s| !s.active)
            if (slot < 0) {                                                              // else min_by_key(age) over all slots
                slot = 0;
                for (int i = 1; i < 64; i++) if (S.st_age[i] < S.st_age[slot]) slot = i;
            }
            voice_from_init(mypool[2 * slot], vinits[evv.vinit]);
            S.st_state[slot] = 1; S.st_note[slot] = (uint8_t)evv.note; S.st_age[slot] = S.age_counter;
            S.d_note_ons++;
            unsigned long long active = 0;
            for (int i = 0; i < 64; i++) active += S.st_state[i] ? 1 : 0;
            if (active > S.d_max_active) S.d_max_active = active;
        } else if (evv.kind == OWG_MIDI_NOTE_OFF) {
            if (S.sustain_held) myheld[nheld++] = (uint8_t)evv.note;
            else note_off_oldest((uint8_t)evv.note);
        } else {  // pedal
            S.sustain_held = evv.note != 0 ? 1 : 0;
            if (!S.sustain_held) {
                for (int k = 0; k < nheld; k++) note_off_oldest(myheld[k]);
                nheld = 0;
            }
        }
    }
    S.ev = ev;
    held_count[e] = nheld;
    // clean up silent voices, then the render list in slot order (main.rs:1852-1876)
    int n_items = 0;
    for (int i = 0; i < 64; i++) {
        if (!S.st_state[i]) continue;
        if (voice_is_silent(mypool[2 * i], ed.sample_rate, silent_thr)) { S.st_state[i] = 0; S.d_freed++; continue; }
        S.item_pool[n_items] = (uint8_t)(2 * i); S.item_fade[n_items] = -1; S.item_fade_len[n_items] = 1u;
        n_items++;
    }
    S.n_items = n_items;
}

__global__ void engine_diag_kernel(const EngineState* __restrict__ states, const EngineChainState* __restrict__ chains, int n_engines, EngineDiag* diag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_engines) return;
    atomicAdd(&diag->nan_guard, states[e].d_nan_guard); atomicAdd(&diag->out_nan, chains[e].d_out_nan); atomicAdd(&diag->steals, states[e].d_steals);
    atomicAdd(&diag->note_ons, states[e].d_note_ons); atomicAdd(&diag->voices_freed, states[e].d_freed); atomicMax(&diag->max_active, states[e].d_max_active);
}

}  // namespace owgd
