// CUDA side of the melange power amplifier: the tile collectives of owg_pa_core.h on sub-warp shuffles, the batch kernel
// (one 16-lane tile per instance, two instances per warp, 8 per CTA) and the settled-state kernel.
//   reference: crates/openwurli-dsp/src/gen_power_amp.rs:8838 (process_sample), power_amp.rs:279-465 (adapter), :65-165 (rail sag);
//   chain B's output stage: tools/preamp-bench/src/main.rs:478-496 (volume^2 -> PowerAmp::new() -> Speaker -> POST_SPEAKER_GAIN).
#pragma once
#include "owg_device.cuh"

// a / b for a model-constant divisor b whose reciprocal r was staged by recip_prepare(): div_by() is the second half of the division
// sequence the compiler emits for `a / b` (same instructions, same bits; operands outside its validated range take the full division)
namespace owgd {
__device__ __forceinline__ double pa_div_const(double a, double b, double r) {
    Recip rc;
    rc.r = r; rc.nb = -b; rc.b = b;
    return div_by(a, rc);
}
}  // namespace owgd
#if defined(__CUDA_ARCH__)
#define PA_DIV_CONST(a, d, IB, IR) owgd::pa_div_const((a), (d)[IB], (d)[IR])
#endif
#include "owg_pa_core.h"

namespace owgd {

// 16 lanes of a warp.  The two tiles of a warp execute in lock-step (owg_pa_core.h keeps every collective on warp-uniform control flow),
// so shuffles and votes name the full warp -- a sub-warp mask compiles to a WARPSYNC / collective / ENDCOLLECTIVE sequence around every one
struct PaCudaTile {
    int lane;
    unsigned half;  // bit mask of this tile's lanes inside the warp
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ double shfl(double x, int src) const { return __shfl_sync(0xffffffffu, x, src, 16); }
    __device__ __forceinline__ double shfl_xor(double x, int m) const { return __shfl_xor_sync(0xffffffffu, x, m, 16); }
    __device__ __forceinline__ int shfl_i(int x, int src) const { return __shfl_sync(0xffffffffu, x, src, 16); }
    __device__ __forceinline__ int shfl_xor_i(int x, int m) const { return __shfl_xor_sync(0xffffffffu, x, m, 16); }
    __device__ __forceinline__ bool any(bool p) const { return (__ballot_sync(0xffffffffu, p) & half) != 0u; }   // over the tile
    __device__ __forceinline__ bool warp_any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }           // over both tiles
    __device__ __forceinline__ uint32_t max_u32(uint32_t x) const { return __reduce_max_sync(half, x); }          // redux.sync over the tile
    __device__ __forceinline__ int first_lane(bool p) const { return (__ffs((int)(__ballot_sync(0xffffffffu, p) & half)) - 1) & 15; }
};

// reciprocals of the model-constant divisors (PA_DIV_CONST), one device per thread; the caller synchronises the CTA around it
__device__ __forceinline__ void pa_stage_reciprocals(PaShared& sh) {
    if (threadIdx.x < PA_NDEV) {
        double* d = sh.dev[threadIdx.x];
        d[PD_R_NF_VT] = recip_prepare(d[PD_NF_VT]).r; d[PD_R_NR_VT] = recip_prepare(d[PD_NR_VT]).r;
        d[PD_R_NE_VT] = recip_prepare(d[PD_NE_VT]).r; d[PD_R_NC_VT] = recip_prepare(d[PD_NC_VT]).r;
        d[PD_R_VAR] = recip_prepare(d[PD_VAR]).r; d[PD_R_VAF] = recip_prepare(d[PD_VAF]).r;
        d[PD_R_IKF] = recip_prepare(d[PD_IKF]).r; d[PD_R_IKR] = recip_prepare(d[PD_IKR]).r;
    }
}

#define OWG_PA_TILES_PER_CTA 8
#define OWG_PA_THREADS (OWG_PA_TILES_PER_CTA * 16)
#define OWG_PA_CTAS_PER_SM 4  // 128 registers per thread: 16 warps = 32 instances per SM (measured: 3 -> 4 CTAs +11 %, 5 no better and slower per wave)

struct PaSpeakerPost {  // chain B's tail behind the amplifier: Speaker::process * POST_SPEAKER_GAIN (main.rs:495); off for the plain adapter
    SpkState spk;
    const OwgChainInit* ci;
    __device__ __forceinline__ double operator()(double v) { return ci ? speaker(v, spk, *ci) * 7.498942093324558 : v; }
};

// rows: [n_inst][stride] in place (input -> output).  index: optional list of the instances this launch covers (one sample-rate model).
// cinits != nullptr: chain B output stage (volume^2, --no-poweramp, speaker); otherwise the plain adapter (owg_power_amp_batch).
__global__ void __launch_bounds__(OWG_PA_THREADS, OWG_PA_CTAS_PER_SM) pa_melange_kernel(const PaModel* __restrict__ model, const PaSettled* __restrict__ settled, double* __restrict__ rows,
                                                                    int64_t stride, const int32_t* __restrict__ index, int64_t n_tiles,
                                                                    const unsigned long long* __restrict__ n_samples, int64_t n_samp_all, int rail_sag,
                                                                    const OwgChainInit* __restrict__ cinits, double* __restrict__ rails,
                                                                    uint32_t* __restrict__ counters) {
    __shared__ PaShared sh;
    __shared__ PaScratch sc[OWG_PA_TILES_PER_CTA];
    pa_stage_shared(*model, sh, threadIdx.x, blockDim.x);
    __syncthreads();
    pa_stage_reciprocals(sh);
    __syncthreads();
    const int tile_in_cta = threadIdx.x >> 4;
    const int64_t tile = (int64_t)blockIdx.x * OWG_PA_TILES_PER_CTA + tile_in_cta;
    const bool valid = tile < n_tiles;  // an odd tile count leaves one half-warp without a row: it keeps its neighbour company
    const int64_t tsafe = valid ? tile : n_tiles - 1;
    const int64_t inst = index ? index[tsafe] : tsafe;
    PaCudaTile t;
    t.lane = threadIdx.x & 15;
    t.half = 0xffffu << (threadIdx.x & 16);
    double* row = rows + (size_t)inst * stride;
    const int64_t n = n_samples ? (int64_t)n_samples[inst] : n_samp_all;
    const int64_t n_other = __shfl_xor_sync(0xffffffffu, n, 16);
    const int64_t n_steps = n > n_other ? n : n_other;
    double* r_out = rails ? rails + 2 * inst : nullptr;
    uint32_t* c_out = counters ? counters + 4 * inst : nullptr;
    PaSpeakerPost post;
    post.spk = SpkState{0.0, 0.0, 0.0, 0.0, 0.0};
    post.ci = cinits ? &cinits[inst] : nullptr;
    const double vol = cinits ? cinits[inst].volume : 1.0;
    const bool bypass = cinits ? cinits[inst].no_poweramp != 0 : false;
    pa_tile_render(t, *model, sh, sc[tile_in_cta], settled, row, row, n, n_steps, valid, vol, rail_sag != 0, bypass, nullptr, r_out, c_out, post);
}

// compute_settled_state (power_amp.rs:291-296) behind CircuitState::default() (gen_power_amp.rs:8421-8490): one tile, 50 + 44 100 silent
// samples of the raw solver with the baked 88.2 kHz tables; once per device and process (the reference's OnceLock).  The second half-warp
// runs the same trajectory redundantly (lock-step collectives) and writes nothing.
__global__ void __launch_bounds__(32) pa_settle_kernel(const PaModel* __restrict__ model, PaSettled* __restrict__ out, PaSettled* __restrict__ scratch_out) {
    __shared__ PaShared sh;
    __shared__ PaScratch sc[2];
    pa_stage_shared(*model, sh, threadIdx.x, blockDim.x);
    __syncthreads();
    pa_stage_reciprocals(sh);
    __syncthreads();
    PaCudaTile t;
    t.lane = threadIdx.x & 15;
    t.half = 0xffffu << (threadIdx.x & 16);
    PaSpeakerPost post;
    post.ci = nullptr;
    pa_tile_render(t, *model, sh, sc[threadIdx.x >> 4], nullptr, nullptr, nullptr, PA_SETTLE_SAMPLES, PA_SETTLE_SAMPLES, true, 1.0, false, false,
                   threadIdx.x < 16 ? out : scratch_out, nullptr, nullptr, post);
}

}  // namespace owgd
