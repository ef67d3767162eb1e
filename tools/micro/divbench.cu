// Micro-benchmarks for FP64 division latency on sm_100a. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <cmath>

__device__ __forceinline__ double rcp64h(double b) {
    double r;
    asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo,hi}, %1; rcp.approx.ftz.f64 %0, %1; }" : "=d"(r) : "d"(b));
    return r;
}
// custom division: fp32 seed -> 3 Newton steps -> Markstein correction
__device__ __forceinline__ double div_custom(double a, double b) {
    float bf = (float)b;
    float rf = __frcp_rn(bf);           // ~24 bits
    double r = (double)rf;
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);                   // ~48 bits
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);                   // ~96 bits -> limited by rounding: ~1 ulp
    double q = a * r;
    double rem = fma(-b, q, a);
    return fma(rem, r, q);
}
__global__ void lat(double* sink, int iters, double a, int mode) {
    double x = 1.0 + threadIdx.x * 1e-3;
    for (int i = 0; i < iters; i++) {
        if (mode == 0) { x = x / a; x = x / a; x = x / a; x = x / a; }                       // dividend-dependent chain
        else if (mode == 1) { x = a / x; x = a / x; x = a / x; x = a / x; }                  // divisor-dependent chain
        else if (mode == 2) { x = rcp64h(x) + 1.0; x = rcp64h(x) + 1.0; x = rcp64h(x) + 1.0; x = rcp64h(x) + 1.0; }
        else if (mode == 3) { x = div_custom(a, x); x = div_custom(a, x); x = div_custom(a, x); x = div_custom(a, x); }
        else if (mode == 4) { x = __drcp_rn(x) + 1.0; x = __drcp_rn(x) + 1.0; x = __drcp_rn(x) + 1.0; x = __drcp_rn(x) + 1.0; }
        else if (mode == 5) { x = sqrt(x) + 1.0; x = sqrt(x) + 1.0; x = sqrt(x) + 1.0; x = sqrt(x) + 1.0; }
        else if (mode == 6) { x = (double)__frcp_rn((float)x) + 1.0; x = (double)__frcp_rn((float)x) + 1.0; x = (double)__frcp_rn((float)x) + 1.0; x = (double)__frcp_rn((float)x) + 1.0; }
    }
    sink[threadIdx.x] = x;
}
// exactness: compare custom division with '/' on random inputs
__global__ void check(uint64_t seed, int n, unsigned long long* bad, double lo_exp, double hi_exp) {
    uint64_t s = seed + (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    unsigned long long local = 0;
    for (int i = 0; i < n; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        uint64_t ma = (s & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        uint64_t mb = (s & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
        double a = __longlong_as_double(ma), b = __longlong_as_double(mb);
        int ea = (int)((s >> 52) & 63) - 32, eb = (int)((s >> 58) & 31) - 16;
        a = ldexp(a, ea); b = ldexp(b, eb);
        if (s & 1) a = -a;
        if (s & 2) b = -b;
        if (div_custom(a, b) != a / b) local++;
    }
    atomicAdd(bad, local);
}
int main() {
    double* sink; cudaMalloc(&sink, 32 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"x/a (dividend chain)", "a/x (divisor chain)", "MUFU.RCP64H+DADD", "custom div a/x", "__drcp_rn+DADD", "sqrt+DADD", "frcp32+cvt+DADD"};
    for (int mode = 0; mode < 7; mode++) {
        int it = 1 << 15;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); lat<<<1, 32>>>(sink, it, 1.37, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-24s %.1f ns per op\n", names[mode], best * 1e6 / (4.0 * it));
    }
    unsigned long long* bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    check<<<1184, 256>>>(12345, 4096, bad, 0, 0);
    unsigned long long h; cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
    printf("custom div mismatches: %llu of %llu\n", h, 1184ull * 256 * 4096);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
