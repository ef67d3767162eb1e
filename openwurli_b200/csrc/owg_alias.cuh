// alias_audit::analyze (alias_audit.rs:163-282) as a device-side reduction: the rendered streams never leave the GPU, only 29 numbers per
// stream do.  One CTA per stream:
//   phase A  refine_f0: thread c evaluates the single-bin DFT of candidate c (nominal f0, then nominal - 5 Hz ... + 5 Hz in accumulated
//            0.1 Hz steps, the reference's own `f += 0.1` sequence); thread 0 then takes the first strict maximum in candidate order;
//   phase B  threads 0..11: the harmonic DFTs at (k + 1) f0;  thread 12: the 5-18 kHz band RMS through four RBJ biquads (serial);
//   phase C  thread 0: dB conversions, plateau metric.
// Every DFT is accumulated sample by sample in the reference's order by ONE thread, so the sums are the reference's sums up to the
// <= 1-2 ulp difference between CUDA's and glibc's sin / cos (the same budget as every other libm call).
#pragma once
#include "owg_device.cuh"

namespace owgd {

#define OWG_ALIAS_MAX_CAND 120
#define OWG_ALIAS_OUT 29

struct AliasBq { double b0, b1, b2, a1, a2; };

template <typename T>
__device__ __forceinline__ double alias_dft_mag(const T* __restrict__ tail, const long long n, const double freq, const double sr) {
    const double nn = (double)n;
    double re = 0.0, im = 0.0;
    const double omega = 2.0 * 3.14159265358979323846 * freq / sr;
    for (long long i = 0; i < n; i++) {
        const double phase = omega * (double)i;
        double sn, cs;
        sincos(phase, &sn, &cs);
        const double s = (double)tail[i];
        re += s * cs;
        im -= s * sn;
    }
    return 2.0 * sqrt((re / nn) * (re / nn) + (im / nn) * (im / nn));
}

template <typename T>
__global__ void __launch_bounds__(128) alias_analyze_kernel(const T* __restrict__ rows, const long long stride, const long long n_samples, const long long analyze_n,
                                                            const double sr, const double* __restrict__ nominal_f0, const AliasBq hp, const AliasBq lp,
                                                            double* __restrict__ out /*[row][29]*/) {
    __shared__ double s_cand[OWG_ALIAS_MAX_CAND], s_mag[OWG_ALIAS_MAX_CAND];
    __shared__ int s_ncand;
    __shared__ double s_f0, s_hmag[12], s_hf;
    const int row = blockIdx.x, tid = threadIdx.x;
    const T* tail = rows + (size_t)row * stride + (n_samples - analyze_n);
    const double nominal = nominal_f0[row];
    if (tid == 0) {
        int n = 0;
        s_cand[n++] = nominal;
        for (double f = nominal - 5.0; f <= nominal + 5.0 && n < OWG_ALIAS_MAX_CAND; f += 0.1) s_cand[n++] = f;
        s_ncand = n;
    }
    __syncthreads();
    if (tid < s_ncand) s_mag[tid] = alias_dft_mag(tail, analyze_n, s_cand[tid], sr);
    __syncthreads();
    if (tid == 0) {
        double best_f = s_cand[0], best_mag = s_mag[0];
        for (int c = 1; c < s_ncand; c++)
            if (s_mag[c] > best_mag) { best_mag = s_mag[c]; best_f = s_cand[c]; }
        s_f0 = best_f;
    }
    __syncthreads();
    const double f0 = s_f0;
    if (tid < 12) s_hmag[tid] = alias_dft_mag(tail, analyze_n, (double)(tid + 1) * f0, sr);
    else if (tid == 12) {
        double h1a = 0, h1b = 0, h2a = 0, h2b = 0, l1a = 0, l1b = 0, l2a = 0, l2b = 0, sum_sq = 0.0;
        for (long long i = 0; i < analyze_n; i++) {
            const double x = (double)tail[i];
            const double y1 = hp.b0 * x + h1a;  h1a = hp.b1 * x - hp.a1 * y1 + h1b;  h1b = hp.b2 * x - hp.a2 * y1;
            const double y2 = hp.b0 * y1 + h2a; h2a = hp.b1 * y1 - hp.a1 * y2 + h2b; h2b = hp.b2 * y1 - hp.a2 * y2;
            const double y3 = lp.b0 * y2 + l1a; l1a = lp.b1 * y2 - lp.a1 * y3 + l1b; l1b = lp.b2 * y2 - lp.a2 * y3;
            const double y4 = lp.b0 * y3 + l2a; l2a = lp.b1 * y3 - lp.a1 * y4 + l2b; l2b = lp.b2 * y3 - lp.a2 * y4;
            sum_sq += y4 * y4;
        }
        s_hf = sqrt(sum_sq / (double)analyze_n);
    }
    __syncthreads();
    if (tid == 0) {
        double* o = out + (size_t)row * OWG_ALIAS_OUT;
        const double h1 = s_hmag[0];  // dft_magnitude(tail, f0, sr): the same sum as harmonic 1
        o[0] = f0;
        o[1] = h1 > 0.0 ? 20.0 * log10(h1) : -200.0;
        for (int k = 0; k < 12; k++) {
            const double mag = s_hmag[k];
            o[2 + k] = mag > 0.0 ? 20.0 * log10(mag) : -200.0;
            o[14 + k] = h1 > 0.0 ? 20.0 * log10(mag / h1) : -200.0;
        }
        o[14] = 0.0;
        double worst = -INFINITY;
        int worst_from = 6;
        for (int i = 5; i < 10; i++) {
            const double delta = o[14 + i + 1] - o[14 + i];
            if (delta > worst) { worst = delta; worst_from = i + 1; }
        }
        o[26] = worst; o[27] = (double)worst_from;
        o[28] = h1 > 0.0 ? 20.0 * log10(s_hf / h1) : -200.0;
    }
}

}  // namespace owgd
