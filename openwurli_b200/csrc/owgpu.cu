// libowgpu C ABI (include/owgpu.h): plans, launches, copies.  No CPU fallback: every entry point
// that renders fails with OWG_E_NO_DEVICE when no CUDA device is usable.
#include "../../include/owgpu.h"
#include "host_setup.h"
#include "owg_kernels.cuh"
#include "owg_tile.cuh"
#include "owg_legacy.cuh"
#include "owg_engine.cuh"
#include "owg_alias.cuh"
#include "owg_poweramp.cuh"

#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

using namespace owgd;

namespace {

thread_local std::string g_err;
thread_local owg_diag g_last_diag;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

// Every entry point that selects a device leaves the calling thread's current CUDA device as it found it (a co-resident framework
// such as torch assumes its own notion of "current device" survives a library call).
struct DeviceRestore {
    int prev = -1;
    DeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            const int _code = (_e == cudaErrorMemoryAllocation) ? OWG_E_OOM : OWG_E_CUDA;         \
            return fail(_code, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
        }                                                                                          \
    } while (0)

inline int popcount32(uint32_t x) { int c = 0; while (x) { c += (int)(x & 1u); x >>= 1; } return c; }

int usable_devices() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// Per-device process-wide cache (the reference's OnceLock, melange_adapter.rs:12): settled preamp state.
struct DeviceCache {
    DkState* d_settled = nullptr;
    bool fade_uploaded = false;
    bool dkdev_uploaded = false;
    // grow-only device staging buffer for host-output calls (avoids an 8.6 GB cudaMalloc/cudaFree per one-shot render)
    void* stage = nullptr;
    size_t stage_bytes = 0;
    bool stage_in_use = false;
    // Tremolo::new per preamp rate: the oscillator state after the 50 warm-up + 2*sr settle samples (tremolo.rs:92-102).  It depends on
    // the rate only (depth enters shunt_impedance, tremolo.rs:152-167), so it is computed once per device and rate -- like the
    // reference would if it cached its constructor the way it caches the preamp's settled state.
    struct TrmCtor { TrmRun* state = nullptr; cudaEvent_t ready = nullptr; };
    std::map<uint64_t, TrmCtor> trm_ctor;
    // melange power amplifier: the settled CircuitState (power_amp.rs:289-297, the reference's OnceLock) and one model per sample rate
    PaSettled* d_pa_settled = nullptr;
    std::map<uint64_t, PaModel*> d_pa_models;
};
std::mutex g_cache_mu;
std::map<int, DeviceCache> g_cache;

int ensure_device_cache(int device, cudaStream_t stream, DeviceCache** out, int64_t* launches) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    DeviceCache& c = g_cache[device];
    if (!c.fade_uploaded) {
        double fade[16];
        owg::noise_fade_table(fade);
        CK(cudaMemcpyToSymbol(c_noise_fade, fade, sizeof(fade)));
        c.fade_uploaded = true;
    }
    if (!c.dkdev_uploaded) {  // junction constants of the tiled kernel's Newton loop, computed on this device (owg_tile.cuh)
        DkDev* tmp = nullptr;
        CK(cudaMalloc(&tmp, sizeof(DkDev)));
        dkdev_init_kernel<<<1, 32, 0, stream>>>(tmp);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        CK(cudaMemcpyToSymbol(c_dkdev, tmp, sizeof(DkDev), 0, cudaMemcpyDeviceToDevice));
        CK(cudaFree(tmp));
        c.dkdev_uploaded = true;
        if (launches) *launches += 1;
    }
    if (!c.d_settled) {
        CK(cudaMalloc(&c.d_settled, sizeof(DkState)));
        settle_kernel<<<1, 32, 0, stream>>>(c.d_settled);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        if (launches) *launches += 1;
    }
    *out = &c;
    return OWG_OK;
}

thread_local int64_t g_h2d_bytes = 0;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t count) {
        if (count <= n) return OWG_OK;
        if (p) { cudaFree(p); p = nullptr; n = 0; }
        if (count == 0) return OWG_OK;
        CK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
        return OWG_OK;
    }
    int upload(const std::vector<T>& h, cudaStream_t s) {
        g_h2d_bytes += h.size() * sizeof(T);
        if (int rc = alloc(h.size())) return rc;
        if (!h.empty()) CK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
        return OWG_OK;
    }
};

}  // namespace

struct owg_plan {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int kind = 0;  // 0 = voices (chain V), 1 = bench (chain B), 2 = preamp-only batch (input rows supplied by the caller)
    double* metrics_ptr = nullptr;   // device [n][OWG_METRICS] accumulators when the output mode is "metrics"
    int64_t w_begin = 0, w_end = 0;  // analysis window, base-rate samples
    DevBuf<double> d_f0s;            // per job: nominal fundamental, sample rate
    const double* in_ptr = nullptr;  // kind 2: caller's input [n][in_stride] (host or device, like `out`)
    int in_location = -1;            // -1 = same as the output's location
    int64_t in_stride = 0;
    bool collect_diag = false;
    bool legacy = false;             // owg_opts.preamp_model == OWG_PREAMP_LEGACY8
    bool use_split = false;          // chain_split_kernel (decided at plan time from the batch size)
    bool use_tile = false;           // chain_tile_kernel: 4 lanes per instance (the default for melange chain-B batches)
    int tile_ipw = 0;                // instance tiles per DK warp (1..7)
    bool taps = false;               // calibrate taps T1..T4 in addition to T5 (owg_render_calibrate)
    DevBuf<double> d_legacy_recs;    // [group][OWG_LG_STRIDE]
    int64_t n = 0;
    std::vector<unsigned long long> n_samples;
    unsigned long long max_samples = 0;
    // host copies
    std::vector<OwgPreampGroup> groups;
    std::vector<int32_t> group_rec_index;
    std::vector<int> trem_group_ids;
    std::vector<WarpEntry> warps_static, warps_trem;
    int64_t trem_n_os_max = 0;
    // device
    DevBuf<OwgVoiceInit> d_vinit;
    DevBuf<OwgChainInit> d_cinit;
    DevBuf<unsigned long long> d_nsamp;
    DevBuf<int32_t> d_order;
    DevBuf<WarpEntry> d_warps_static, d_warps_trem;
    DevBuf<OwgPreampGroup> d_groups;
    DevBuf<int32_t> d_group_rec_index;
    DevBuf<int> d_trem_ids;
    DevBuf<double> d_static_recs, d_ans, d_pot_seq, d_trem_recs, d_carry;
    DevBuf<TrmRun> d_trm_run, d_trm_ctor;       // oscillator state: running / as constructed (Tremolo::new, computed at plan time)
    DevBuf<LdrRun> d_ldr_run;                   // LDR envelope + resistance-tracking state between chunks
    cudaStream_t stream_copy = nullptr;          // device->host copy-back of finished chunks
    cudaStream_t stream_voice = nullptr;         // the tail of the voice render runs here, beside the first chain chunks
    cudaEvent_t ev_voice1 = nullptr, ev_voice2 = nullptr;
    DevBuf<double> d_vcarry;                     // voice recurrence state between the two voice launches
    cudaStream_t stream_trem = nullptr;          // the serial Twin-T oscillator runs here, one chunk ahead of its consumers
    std::vector<cudaEvent_t> chunk_events;       // oscillator chunk c finished
    std::vector<cudaEvent_t> chain_ev;           // pairs around every chain launch (device time of the dominant kernel)
    DevBuf<double> d_stage;  // device-side output when the caller's buffer is host memory
    DevBuf<DevDiag> d_diag;
    DeviceCache* cache = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    int64_t launches_last = 0;
    int64_t h2d_bytes = 0;
    float main_ms = 0.f, total_ms = 0.f;
    ~owg_plan() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (evk0) cudaEventDestroy(evk0);
        if (evk1) cudaEventDestroy(evk1);
        for (auto e : chunk_events) cudaEventDestroy(e);
        for (auto e : chain_ev) cudaEventDestroy(e);
        if (stream_trem) cudaStreamDestroy(stream_trem);
        if (stream_copy) cudaStreamDestroy(stream_copy);
        if (stream_voice) cudaStreamDestroy(stream_voice);
        if (ev_voice1) cudaEventDestroy(ev_voice1);
        if (ev_voice2) cudaEventDestroy(ev_voice2);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace {

int plan_common(owg_plan* pl, const owg_opts* opts) {
    if (usable_devices() <= 0) return fail(OWG_E_NO_DEVICE, "no usable CUDA device (libowgpu has no CPU fallback)");
    owg_opts o;
    if (opts) o = *opts; else owg_default_opts(&o);
    if (o.precision != OWG_PRECISION_F64_EXACT) return fail(OWG_E_UNSUPPORTED, "only OWG_PRECISION_F64_EXACT is implemented");
    if (o.preamp_model != OWG_PREAMP_MELANGE12 && o.preamp_model != OWG_PREAMP_LEGACY8) return fail(OWG_E_UNSUPPORTED, "unknown preamp_model");
    if (o.power_amp_model != OWG_POWER_AMP_BEHAVIORAL)
        return fail(OWG_E_UNSUPPORTED, "power_amp_model: the melange power amplifier is served by owg_render_bench, owg_chain_batch, owg_render_midi and owg_power_amp_batch");
    pl->legacy = o.preamp_model == OWG_PREAMP_LEGACY8;
    int dev = o.device;
    if (popcount32(o.device_mask) == 1) { dev = 0; while (!(o.device_mask & (1u << dev))) dev++; }  // a one-bit mask names the device
    if (dev < 0) CK(cudaGetDevice(&dev));
    CK(cudaSetDevice(dev));
    pl->device = dev;
    if (o.stream) { pl->stream = (cudaStream_t)o.stream; pl->own_stream = false; }
    else { CK(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking)); pl->own_stream = true; }
    pl->collect_diag = o.collect_diag != 0;
    CK(cudaEventCreate(&pl->ev0)); CK(cudaEventCreate(&pl->ev1)); CK(cudaEventCreate(&pl->evk0)); CK(cudaEventCreate(&pl->evk1));
    int64_t l = 0;
    if (int rc = ensure_device_cache(dev, pl->stream, &pl->cache, &l)) return rc;
    return OWG_OK;
}

// One instance of the shared mono chain as the planner sees it.
struct InstSpec {
    int init_order = 0;   // OWG_INIT_RESET_THEN_SET (cmd_render) | OWG_INIT_SET_THEN_RESET (render-poly / render-midi)
    double fs;            // base sample rate
    int oversample;       // 2x oversampled preamp
    unsigned long long n_samples;
    double depth, r_ldr;  // tremolo depth (>0) or static LDR resistance
    OwgChainInit ci;
};

// Group instances by (base rate, oversampling, static R | tremolo depth) -- they share DK matrices and the shadow solve --
// and pack each group into warps of 31 instances + 1 shadow lane, longest renders first.
void build_groups_and_warps(owg_plan* pl, std::vector<InstSpec>& specs, std::vector<int32_t>* order_out) {
    typedef std::tuple<double, int, int, double, int> Key;  // (fs, oversample, is_trem, r_or_depth, legacy DC at r)
    std::map<Key, int> key_to_group;
    std::vector<std::vector<int32_t>> members;
    const double R0 = 9.99999999999999854e4;
    const size_t n = specs.size();
    pl->n_samples.resize(n);
    for (size_t i = 0; i < n; i++) {
        InstSpec& sp = specs[i];
        const double psr = sp.oversample ? sp.fs * 2.0 : sp.fs;
        const bool trem = sp.depth > 0.0;
        double r_eff = R0;
        bool dirty = false;
        int dc_at_r = 0;
        if (pl->legacy) {  // legacy: r = max(r_ldr, 1000) with a 0.01 Ohm change threshold against the constructor's 1 MOhm (dk_preamp_legacy.rs:620-626)
            r_eff = 1.0e6;
            if (!trem) {
                const double r = sp.r_ldr > 1000.0 ? sp.r_ldr : 1000.0;
                if (std::fabs(r - 1.0e6) > 0.01) r_eff = r;
                dc_at_r = sp.init_order == OWG_INIT_SET_THEN_RESET ? 1 : 0;  // reset() after the set: DC point re-solved at r (:628-642)
            }
        } else if (!trem && sp.init_order == OWG_INIT_SET_THEN_RESET) {
            // melange: reset() re-clones the settled 100 kOhm state after the set (melange_adapter.rs:22-29): r_eff stays R0
        } else if (!trem && std::isfinite(sp.r_ldr)) {  // reset(); set_ldr_resistance(r_ldr)  (main.rs:438-439, gen_preamp.rs:1973-1984)
            const double r = sp.r_ldr < 1.0e3 ? 1.0e3 : (sp.r_ldr > 1.0e6 ? 1.0e6 : sp.r_ldr);
            if (!(std::fabs(r - R0) < 1e-12)) { r_eff = r; dirty = true; }
        }
        const Key key(sp.fs, sp.oversample, trem ? 1 : 0, trem ? sp.depth : r_eff, dc_at_r);
        auto it = key_to_group.find(key);
        int g;
        if (it == key_to_group.end()) {
            g = (int)pl->groups.size();
            key_to_group[key] = g;
            OwgPreampGroup gr;
            std::memset(&gr, 0, sizeof(gr));
            gr.preamp_sr = psr;
            gr.r_static = r_eff;
            gr.tremolo_depth = trem ? sp.depth : 0.0;
            gr.use_defaults = (std::fabs(psr - 48000.0) <= 0.5 && !dirty) ? 1 : 0;
            gr.dc_at_r = dc_at_r;
            gr.n_os = 0;
            pl->groups.push_back(gr);
            members.emplace_back();
        } else g = it->second;
        sp.ci.group = g;
        sp.ci.oversample = sp.oversample;
        pl->n_samples[i] = sp.n_samples;
        pl->max_samples = std::max<unsigned long long>(pl->max_samples, sp.n_samples);
        const int64_t nos = (int64_t)sp.n_samples * (sp.oversample ? 2 : 1);
        pl->groups[g].n_os = std::max<int64_t>(pl->groups[g].n_os, nos);
        members[g].push_back((int32_t)i);
    }
    // record indices: static groups index d_static_recs by group id, tremolo groups index d_trem_recs
    pl->group_rec_index.assign(pl->groups.size(), 0);
    for (size_t g = 0; g < pl->groups.size(); g++) {
        if (pl->groups[g].tremolo_depth > 0.0) {
            pl->group_rec_index[g] = (int32_t)pl->trem_group_ids.size();
            pl->trem_group_ids.push_back((int)g);
            pl->trem_n_os_max = std::max(pl->trem_n_os_max, pl->groups[g].n_os);
        } else pl->group_rec_index[g] = (int32_t)g;
    }
    // Lanes per warp.  A warp's Newton loop runs max-over-lanes iterations, and at small batch sizes the kernel is
    // latency-bound with idle schedulers, so fewer instances per warp (more warps) shortens every warp's critical path.
    // Aim for ~1 warp per SM sub-partition (148 SMs x 4; measured optimum), capped at 31 instances + the shadow lane.
    int lpw = (int)((n + 592 - 1) / 592);
    lpw = lpw < 1 ? 1 : (lpw > 31 ? 31 : lpw);
    // Warp-specialised chain (chain_split_kernel: 2 warps per instance group) when the whole batch still fits in one wave of
    // 3 CTAs per SM -- the fourth quarter of each register file stays free for the oscillator / matrix / voice kernels that
    // run concurrently on the other streams.  Larger batches are throughput-bound and keep the one-warp kernel.
    {
        const char* split_env = getenv("OWG_CHAIN_SPLIT");
        const bool allow = !(split_env && split_env[0] == '0') && !pl->collect_diag && !pl->legacy;
        const int lpw_split = (int)((n + 444 - 1) / 444);
        pl->use_split = allow && lpw_split <= 31;
        if (pl->use_split) lpw = lpw_split < 1 ? 1 : lpw_split;
    }
    if (const char* e = getenv("OWG_LANES_PER_WARP")) { const int v = atoi(e); if (v >= 1 && v <= 31) lpw = v; }
    // Lane-tiled chain (chain_tile_kernel: 4 lanes per instance, 4 DK warps + 1 I/O warp per CTA, 2 CTAs per SM): the default
    // while the batch fits two waves of CTAs; beyond that the batch is throughput-bound and the one-thread-per-instance kernels
    // spend fewer issue slots per instance.  OWG_CHAIN_KERNEL = tile | split | warp forces one of the three.
    {
        const char* ck = getenv("OWG_CHAIN_KERNEL");
        const bool force_tile = ck && ck[0] == 't';
        const bool forbid_tile = ck && (ck[0] == 's' || ck[0] == 'w');
        if (ck && ck[0] == 'w') pl->use_split = false;
        const int64_t cta_slots = 2 * 148;
        int ipw = (int)((n + cta_slots * OWG_TILE_AW - 1) / (cta_slots * OWG_TILE_AW));
        ipw = ipw < 1 ? 1 : (ipw > OWG_TILE_IPW ? OWG_TILE_IPW : ipw);
        if (const char* e = getenv("OWG_TILE_IPW")) { const int v = atoi(e); if (v >= 1 && v <= OWG_TILE_IPW) ipw = v; }
        const bool fits = (int64_t)n <= 2 * cta_slots * OWG_TILE_AW * OWG_TILE_IPW;
        pl->use_tile = !pl->legacy && !forbid_tile && (force_tile || fits);
        if (pl->use_tile) { pl->tile_ipw = ipw; pl->use_split = false; lpw = OWG_TILE_AW * ipw; }
    }
    std::vector<int32_t>& order = *order_out;
    order.reserve(n);
    for (size_t g = 0; g < pl->groups.size(); g++) {
        std::vector<int32_t>& mem = members[g];
        std::stable_sort(mem.begin(), mem.end(), [&](int32_t a, int32_t b) { return pl->n_samples[a] > pl->n_samples[b]; });
        for (size_t off = 0; off < mem.size(); off += (size_t)lpw) {
            WarpEntry we;
            we.group = (int32_t)g;
            we.first = (int32_t)order.size();
            we.count = (int32_t)std::min<size_t>((size_t)lpw, mem.size() - off);
            we._pad = 0;
            we.n_max = 0;
            for (int k = 0; k < we.count; k++) {
                order.push_back(mem[off + k]);
                we.n_max = std::max<int64_t>(we.n_max, (int64_t)pl->n_samples[mem[off + k]]);
            }
            (pl->groups[g].tremolo_depth > 0.0 ? pl->warps_trem : pl->warps_static).push_back(we);
        }
    }
}

// The Twin-T oscillator runs on 8 lanes per group (tremolo_group_tile_kernel) unless OWG_TREM_KERNEL=thread asks for the one-thread
// kernel (kept for A/B checks: the two produce bit-identical sequences).
bool trem_tile_enabled() {
    const char* e = getenv("OWG_TREM_KERNEL");
    return !(e && e[0] == 't' && e[1] == 'h');
}

// Tremolo::new for every tremolo group, launched as early as possible on the oscillator stream (it is ~0.6 s of serial
// device work per 88.2 kHz group and nothing else of the plan depends on it).
int launch_tremolo_ctor(owg_plan* pl) {
    if (pl->trem_group_ids.empty()) return OWG_OK;
    const int nt = (int)pl->trem_group_ids.size();
    int rc = pl->d_groups.upload(pl->groups, pl->stream);
    if (!rc) rc = pl->d_trem_ids.upload(pl->trem_group_ids, pl->stream);
    if (!rc) rc = pl->d_trm_run.alloc((size_t)nt);
    if (!rc) rc = pl->d_trm_ctor.alloc((size_t)nt);
    if (!rc) rc = pl->d_ldr_run.alloc((size_t)nt);
    if (!rc) rc = pl->d_pot_seq.alloc((size_t)nt * (size_t)pl->trem_n_os_max);
    if (!rc && cudaStreamSynchronize(pl->stream) != cudaSuccess) rc = fail(OWG_E_CUDA, "plan upload failed");
    if (!rc && !pl->stream_trem && cudaStreamCreateWithFlags(&pl->stream_trem, cudaStreamNonBlocking) != cudaSuccess)
        rc = fail(OWG_E_CUDA, "stream creation failed");
    if (!rc) {
        // Constructors run at plan time, per-sample processing at execute time: Tremolo::new (50 warm-up + 2*sr settle
        // samples of the Twin-T oscillator, tremolo.rs:84-115) is evaluated once per device and preamp rate (it does not depend on
        // the depth) and cached, like Voice::note_on on the host and DkPreamp::new's cached settled state.  Asynchronous: every
        // later oscillator launch goes to the same in-order stream, so nothing has to wait here.
        auto bits = [](double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; };
        std::lock_guard<std::mutex> lock(g_cache_mu);
        DeviceCache& c = *pl->cache;
        bool all_cached = getenv("OWG_TREM_CTOR_CACHE") == nullptr || getenv("OWG_TREM_CTOR_CACHE")[0] != '0';
        for (int gi = 0; gi < nt && all_cached; gi++) all_cached = c.trm_ctor.count(bits(pl->groups[pl->trem_group_ids[gi]].preamp_sr)) != 0;
        if (all_cached) {
            for (int gi = 0; gi < nt && !rc; gi++) {
                const DeviceCache::TrmCtor& tc = c.trm_ctor[bits(pl->groups[pl->trem_group_ids[gi]].preamp_sr)];
                if (cudaStreamWaitEvent(pl->stream_trem, tc.ready, 0) != cudaSuccess ||
                    cudaMemcpyAsync(pl->d_trm_ctor.p + gi, tc.state, sizeof(TrmRun), cudaMemcpyDeviceToDevice, pl->stream_trem) != cudaSuccess)
                    rc = fail(OWG_E_CUDA, "tremolo constructor cache copy failed");
            }
        } else {
            if (trem_tile_enabled())
                tremolo_group_tile_kernel<<<nt, 32, 0, pl->stream_trem>>>(pl->d_groups.p, pl->d_trem_ids.p, nt, pl->d_pot_seq.p, pl->trem_n_os_max,
                                                                          pl->d_trm_ctor.p, -1, -1, nullptr);
            else
                tremolo_group_kernel<<<nt, 32, 0, pl->stream_trem>>>(pl->d_groups.p, pl->d_trem_ids.p, nt, pl->d_pot_seq.p, pl->trem_n_os_max,
                                                                     pl->d_trm_ctor.p, -1, -1, nullptr);
            if (cudaGetLastError() != cudaSuccess) rc = fail(OWG_E_CUDA, "tremolo constructor kernel launch failed");
            for (int gi = 0; gi < nt && !rc; gi++) {
                const uint64_t key = bits(pl->groups[pl->trem_group_ids[gi]].preamp_sr);
                if (c.trm_ctor.count(key)) continue;
                DeviceCache::TrmCtor tc;
                if (cudaMalloc(&tc.state, sizeof(TrmRun)) != cudaSuccess || cudaEventCreateWithFlags(&tc.ready, cudaEventDisableTiming) != cudaSuccess ||
                    cudaMemcpyAsync(tc.state, pl->d_trm_ctor.p + gi, sizeof(TrmRun), cudaMemcpyDeviceToDevice, pl->stream_trem) != cudaSuccess ||
                    cudaEventRecord(tc.ready, pl->stream_trem) != cudaSuccess) { rc = fail(OWG_E_CUDA, "tremolo constructor cache fill failed"); break; }
                c.trm_ctor[key] = tc;
            }
        }
    }
    return rc;
}

int upload_chain_plan(owg_plan* pl, const std::vector<OwgChainInit>& ci, const std::vector<int32_t>& order) {
    int rc = pl->d_cinit.upload(ci, pl->stream);
    if (!rc) rc = pl->d_nsamp.upload(pl->n_samples, pl->stream);
    if (!rc) rc = pl->d_order.upload(order, pl->stream);
    if (!rc) rc = pl->d_warps_static.upload(pl->warps_static, pl->stream);
    if (!rc) rc = pl->d_warps_trem.upload(pl->warps_trem, pl->stream);
    if (!rc && pl->trem_group_ids.empty()) rc = pl->d_groups.upload(pl->groups, pl->stream);  // else uploaded by launch_tremolo_ctor
    if (!rc) rc = pl->d_group_rec_index.upload(pl->group_rec_index, pl->stream);
    if (pl->legacy) {
        std::vector<double> recs(pl->groups.size() * OWG_LG_STRIDE);
        for (size_t g = 0; g < pl->groups.size(); g++)
            owg::make_legacy_group(pl->groups[g].preamp_sr, pl->groups[g].tremolo_depth > 0.0 ? NAN : pl->groups[g].r_static, &recs[g * OWG_LG_STRIDE],
                                   pl->groups[g].dc_at_r != 0);
        if (!rc) rc = pl->d_legacy_recs.upload(recs, pl->stream);
    } else {
        if (!rc) rc = pl->d_static_recs.alloc(pl->groups.size() * OWG_MAT_STRIDE);
        if (!rc) rc = pl->d_ans.alloc(pl->groups.size() * OWG_AN_SPARSE);
        if (!rc && !pl->trem_group_ids.empty())
            rc = pl->d_trem_recs.alloc(pl->trem_group_ids.size() * (size_t)pl->trem_n_os_max * OWG_MAT_STRIDE);
    }
    if (!rc && pl->collect_diag) rc = pl->d_diag.alloc(1);
    if (!rc && cudaStreamSynchronize(pl->stream) != cudaSuccess) rc = fail(OWG_E_CUDA, "plan upload failed");
    return rc;
}

bool bad_voice_job(const owg_voice_job& j) {
    return !(j.sample_rate > 0.0) || !std::isfinite(j.sample_rate) || !(j.duration_s >= 0.0) || !std::isfinite(j.duration_s) ||
           !std::isfinite(j.velocity);
}

}  // namespace

extern "C" {

int owg_abi_version(void) { return OWG_ABI_VERSION; }
int owg_device_count(void) { return usable_devices(); }
const char* owg_last_error(void) { return g_err.c_str(); }
void owg_default_opts(owg_opts* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->device = -1;
    o->out_location = OWG_OUT_HOST;
    o->precision = OWG_PRECISION_F64_EXACT;
    o->preamp_model = OWG_PREAMP_MELANGE12;
}

int owg_plan_voices(const owg_voice_job* jobs, int64_t n, const owg_opts* opts, owg_plan** plan) {
    DeviceRestore restore_device_;
    if (!plan || n < 0 || (n > 0 && !jobs)) return fail(OWG_E_BAD_ARG, "owg_plan_voices: bad argument");
    for (int64_t i = 0; i < n; i++) if (bad_voice_job(jobs[i])) return fail(OWG_E_BAD_ARG, "owg_plan_voices: job with invalid sample_rate/duration/velocity");
    owg_plan* pl = new owg_plan();
    g_h2d_bytes = 0;
    pl->kind = 0;
    pl->n = n;
    if (int rc = plan_common(pl, opts)) { delete pl; return rc; }
    std::vector<OwgVoiceInit> vi((size_t)n);
    pl->n_samples.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        owg::make_voice_init(jobs[i], &vi[i]);
        pl->n_samples[i] = vi[i].n_samples;
        pl->max_samples = std::max<unsigned long long>(pl->max_samples, vi[i].n_samples);
    }
    int rc = pl->d_vinit.upload(vi, pl->stream);
    if (!rc) rc = pl->d_nsamp.upload(pl->n_samples, pl->stream);
    if (!rc && cudaStreamSynchronize(pl->stream) != cudaSuccess) rc = fail(OWG_E_CUDA, "plan upload failed");
    if (rc) { delete pl; return rc; }
    pl->h2d_bytes = g_h2d_bytes;
    *plan = pl;
    return OWG_OK;
}

// post_gain: optional per-job override of Voice::post_pickup_gain (calibrate's output_scale under a non-default CalibrationConfig)
static int plan_bench_impl(const owg_bench_job* jobs, int64_t n, const owg_opts* opts, owg_plan** plan, const double* post_gain,
                           bool force_pre_only = false) {
    if (!plan || n < 0 || (n > 0 && !jobs)) return fail(OWG_E_BAD_ARG, "owg_plan_bench: bad argument");
    for (int64_t i = 0; i < n; i++) {
        if (bad_voice_job(jobs[i].v)) return fail(OWG_E_BAD_ARG, "owg_plan_bench: job with invalid sample_rate/duration/velocity");
        if (!std::isfinite(jobs[i].volume) || !std::isfinite(jobs[i].speaker_character) || std::isnan(jobs[i].tremolo_depth))
            return fail(OWG_E_BAD_ARG, "owg_plan_bench: non-finite volume/speaker/tremolo_depth");
    }
    owg_plan* pl = new owg_plan();
    g_h2d_bytes = 0;
    pl->kind = 1;
    pl->n = n;
    if (int rc = plan_common(pl, opts)) { delete pl; return rc; }

    std::vector<InstSpec> specs((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        const owg_bench_job& j = jobs[i];
        InstSpec& sp = specs[i];
        sp.fs = j.v.sample_rate;
        sp.oversample = j.v.sample_rate < 88200.0 ? 1 : 0;
        const double ns = j.v.duration_s * j.v.sample_rate;  // (duration * sample_rate) as usize
        sp.n_samples = !(ns == ns) || ns <= 0.0 ? 0ull : (unsigned long long)ns;
        sp.depth = j.tremolo_depth;
        sp.r_ldr = j.r_ldr;
        owg::make_chain_init(j, 0, &sp.ci);
        if (force_pre_only) sp.ci.pre_only = 1;
    }
    std::vector<int32_t> order;
    build_groups_and_warps(pl, specs, &order);
    int rc = launch_tremolo_ctor(pl);  // the device settles the oscillators while the host parameterises the voices
    std::vector<OwgVoiceInit> vi((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        owg::make_voice_init(jobs[i].v, &vi[i]);
        if (post_gain) vi[i].post_pickup_gain = post_gain[i];
    }
    std::vector<OwgChainInit> ci((size_t)n);
    for (int64_t i = 0; i < n; i++) ci[i] = specs[i].ci;
    if (!rc) rc = pl->d_vinit.upload(vi, pl->stream);
    if (!rc) rc = upload_chain_plan(pl, ci, order);
    if (rc) { delete pl; return rc; }
    pl->h2d_bytes = g_h2d_bytes;
    *plan = pl;
    return OWG_OK;
}

int owg_plan_bench(const owg_bench_job* jobs, int64_t n, const owg_opts* opts, owg_plan** plan) {
    DeviceRestore restore_device_;
    return plan_bench_impl(jobs, n, opts, plan, nullptr);
}

int64_t owg_plan_samples(const owg_plan* pl, int64_t i) {
    if (!pl) return -1;
    if (i < 0) return (int64_t)pl->max_samples;
    if (i >= pl->n) return -1;
    return (int64_t)pl->n_samples[(size_t)i];
}

int64_t owg_plan_h2d_bytes(const owg_plan* pl) { return pl ? pl->h2d_bytes : -1; }

int64_t owg_plan_kernel_launches(const owg_plan* pl) { return pl ? pl->launches_last : -1; }

int owg_plan_last_timing(const owg_plan* pl, float* main_kernel_ms, float* total_ms) {
    if (!pl) return fail(OWG_E_BAD_ARG, "null plan");
    if (main_kernel_ms) *main_kernel_ms = pl->main_ms;
    if (total_ms) *total_ms = pl->total_ms;
    return OWG_OK;
}

void owg_plan_destroy(owg_plan* pl) {
    if (!pl) return;
    DeviceRestore restore_device_;
    cudaSetDevice(pl->device);
    delete pl;
}

int owg_plan_execute(owg_plan* pl, double* out, int64_t stride, int32_t out_location) {
    DeviceRestore restore_device_;
    if (!pl) return fail(OWG_E_BAD_ARG, "null plan");
    if (pl->n == 0) return OWG_OK;
    if (!out) return fail(OWG_E_BAD_ARG, "null output");
    if (stride < (int64_t)pl->max_samples) return fail(OWG_E_BAD_ARG, "stride smaller than the longest render");
    CK(cudaSetDevice(pl->device));
    cudaStream_t s = pl->stream;
    double* dout = out;
    bool borrowed_stage = false;
    if (out_location == OWG_OUT_HOST) {
        const size_t need = (size_t)pl->n * (size_t)stride * sizeof(double);
        {
            std::lock_guard<std::mutex> lock(g_cache_mu);
            DeviceCache& c = *pl->cache;
            if (!c.stage_in_use) {
                if (c.stage_bytes < need) {
                    if (c.stage) cudaFree(c.stage);
                    c.stage = nullptr; c.stage_bytes = 0;
                    if (cudaMalloc(&c.stage, need) == cudaSuccess) c.stage_bytes = need; else cudaGetLastError();
                }
                if (c.stage_bytes >= need) { c.stage_in_use = true; borrowed_stage = true; dout = (double*)c.stage; }
            }
        }
        if (!borrowed_stage) {
            if (int rc = pl->d_stage.alloc((size_t)pl->n * (size_t)stride)) return rc;
            dout = pl->d_stage.p;
        }
    } else if (out_location != OWG_OUT_DEVICE) return fail(OWG_E_BAD_ARG, "bad out_location");
    struct StageReturn {  // give the borrowed staging buffer back on every exit path
        DeviceCache* c; bool on;
        ~StageReturn() { if (on) { std::lock_guard<std::mutex> lock(g_cache_mu); c->stage_in_use = false; } }
    } stage_return{pl->cache, borrowed_stage};
    int64_t launches = 0;
    CK(cudaEventRecord(pl->ev0, s));
    auto zero_tails = [&]() -> int {  // ragged batch: rows shorter than the longest end in silence (defined output, whatever the staging buffer held)
        bool ragged = false;
        for (unsigned long long v : pl->n_samples) if (v != pl->max_samples) { ragged = true; break; }
        if (ragged && pl->d_nsamp.p) {
            zero_tails_kernel<<<(unsigned)pl->n, 128, 0, s>>>(dout, stride, pl->d_nsamp.p, pl->n, (int64_t)pl->max_samples);
            CK(cudaGetLastError());
            launches += 1;
        }
        return OWG_OK;
    };
    if (pl->kind != 2) { if (int rc = zero_tails()) return rc; }
    if (pl->collect_diag) CK(cudaMemsetAsync(pl->d_diag.p, 0, sizeof(DevDiag), s));
    // Tremolo groups: the Twin-T oscillator is one serial thread per group, so it is pipelined: it runs on its own stream in
    // chunks, and chunk c's LDR law + matrices + chain run on the main stream while the oscillator already produces chunk c+1.
    // It depends on nothing else in this call, so it is enqueued first (before the voice kernel); the first chunk is short so
    // that the chain can start early.
    const int64_t CH_BASE = 8192, CH_FIRST = 2048;  // base-rate samples per chunk
    int64_t voice_split_at = -1;                    // > 0: the voice render is cut at this sample into two launches
    auto chunk_lo = [&](int64_t c) -> int64_t { return c <= 0 ? 0 : CH_FIRST + (c - 1) * CH_BASE; };
    int64_t n_chunks = 0;
    const int nt = pl->kind >= 1 ? (int)pl->trem_group_ids.size() : 0;
    if (nt > 0) {
        if (!pl->stream_trem) CK(cudaStreamCreateWithFlags(&pl->stream_trem, cudaStreamNonBlocking));
        n_chunks = (int64_t)pl->max_samples <= CH_FIRST ? 1 : 1 + ((int64_t)pl->max_samples - CH_FIRST + CH_BASE - 1) / CH_BASE;
        while ((int64_t)pl->chunk_events.size() < n_chunks + 1) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            pl->chunk_events.push_back(e);
        }
        if (int rc = pl->d_carry.alloc(pl->warps_trem.size() * (size_t)OWG_CARRY * 32)) return rc;
        CK(cudaEventRecord(pl->chunk_events[n_chunks], s));               // orders the oscillator stream after earlier work on `s`
        CK(cudaStreamWaitEvent(pl->stream_trem, pl->chunk_events[n_chunks], 0));
        CK(cudaMemcpyAsync(pl->d_trm_run.p, pl->d_trm_ctor.p, (size_t)nt * sizeof(TrmRun), cudaMemcpyDeviceToDevice, pl->stream_trem));
        for (int64_t c = 0; c < n_chunks; c++) {
            const int64_t os0 = chunk_lo(c) * 2, os1 = chunk_lo(c + 1) * 2;  // covers 2x-oversampled groups; native-rate groups use half
            if (trem_tile_enabled())
                tremolo_group_tile_kernel<<<nt, 32, 0, pl->stream_trem>>>(pl->d_groups.p, pl->d_trem_ids.p, nt, pl->d_pot_seq.p, pl->trem_n_os_max,
                                                                          pl->d_trm_run.p, os0, os1, pl->collect_diag ? pl->d_diag.p : nullptr);
            else
                tremolo_group_kernel<<<nt, 32, 0, pl->stream_trem>>>(pl->d_groups.p, pl->d_trem_ids.p, nt, pl->d_pot_seq.p, pl->trem_n_os_max,
                                                                     pl->d_trm_run.p, os0, os1, pl->collect_diag ? pl->d_diag.p : nullptr);
            CK(cudaGetLastError());
            CK(cudaEventRecord(pl->chunk_events[c], pl->stream_trem));
            launches++;
        }
    }
    if (pl->kind == 2) {
        // preamp-only batch: the rows start as the caller's input signals
        CK(cudaMemcpy2DAsync(dout, (size_t)stride * sizeof(double), pl->in_ptr, (size_t)pl->in_stride * sizeof(double),
                             (size_t)pl->max_samples * sizeof(double), (size_t)pl->n,
                             (pl->in_location < 0 ? out_location : pl->in_location) == OWG_OUT_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
        if (int rc = zero_tails()) return rc;  // behind the input copy: what the caller's rows held beyond their length is not output
    } else {  // chain V for every job
        const int threads = 32;
        const int blocks = (int)((pl->n + threads - 1) / threads);
        // Tremolo-only plans consume the voice rows chunk by chunk: render the first two chunks' worth now and the rest on another
        // stream, beside the first chain chunks (the voice kernel is latency-bound at one warp per 32 renders: ~46 ms for 3 s).
        voice_split_at = (pl->kind == 1 && nt > 0 && pl->warps_static.empty() && (int64_t)pl->max_samples > 2 * (CH_FIRST + CH_BASE))
                             ? CH_FIRST + CH_BASE : -1;
        if (voice_split_at > 0) {
            if (int rc = pl->d_vcarry.alloc((size_t)OWG_VOICE_CARRY * (size_t)pl->n)) return rc;
            if (!pl->stream_voice) CK(cudaStreamCreateWithFlags(&pl->stream_voice, cudaStreamNonBlocking));
            if (!pl->ev_voice1) { CK(cudaEventCreateWithFlags(&pl->ev_voice1, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&pl->ev_voice2, cudaEventDisableTiming)); }
        }
        const bool taps = pl->taps && pl->metrics_ptr;
        for (int part = 0; part < (voice_split_at > 0 ? 2 : 1); part++) {
            const int64_t tb = part == 0 ? 0 : voice_split_at, te = (voice_split_at > 0 && part == 0) ? voice_split_at : -1;
            cudaStream_t vs = part == 0 ? s : pl->stream_voice;
            if (part == 1) { CK(cudaEventRecord(pl->ev_voice1, s)); CK(cudaStreamWaitEvent(vs, pl->ev_voice1, 0)); }
            double* vc = voice_split_at > 0 ? pl->d_vcarry.p : nullptr;
            if (taps)
                voice_kernel<true><<<blocks, threads, 0, vs>>>(pl->d_vinit.p, pl->n, dout, stride, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end, tb, te, vc);
            else
                voice_kernel<false><<<blocks, threads, 0, vs>>>(pl->d_vinit.p, pl->n, dout, stride, nullptr, nullptr, 0, 0, tb, te, vc);
            CK(cudaGetLastError());
            if (part == 1) CK(cudaEventRecord(pl->ev_voice2, vs));
            if (part == 1) launches++;
        }
        launches++;
    }
    bool copied_by_chunks = false;
    size_t chain_ev_used = 0;
    auto chain_event = [&]() -> cudaEvent_t {
        if (chain_ev_used == pl->chain_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); pl->chain_ev.push_back(e); }
        return pl->chain_ev[chain_ev_used++];
    };
    if (pl->kind >= 1) {
        // warp-specialised chain (DK preamp | input + output stages) unless counters are collected or it is switched off
        const char* split_env = getenv("OWG_CHAIN_SPLIT");
        const bool split_chain = pl->use_split;
        const int swap_roles = (split_env && split_env[0] == '2') ? 0 : 1;  // OWG_CHAIN_SPLIT=2: split without the role placement (A/B test)
        const int ng = (int)pl->groups.size();
        if (!pl->legacy) {
            static_matrix_kernel<<<(ng + 31) / 32, 32, 0, s>>>(pl->d_groups.p, ng, pl->d_static_recs.p, pl->d_ans.p);
            CK(cudaGetLastError());
            launches++;
        }
        if (nt > 0 && !pl->legacy) {
            tremolo_an_kernel<<<(ng + 31) / 32, 32, 0, s>>>(pl->d_groups.p, ng, pl->d_ans.p);
            CK(cudaGetLastError());
            launches++;
        }
        CK(cudaEventRecord(pl->evk0, s));
        // Host output: the rows of a finished chunk are copied back on a third stream while the next chunk computes
        // (only when all rows belong to one family of groups, so that a column range is final for every row at once).
        const bool overlap_d2h = out_location == OWG_OUT_HOST && !pl->metrics_ptr && (pl->warps_static.empty() != pl->warps_trem.empty());
        if (overlap_d2h && !pl->stream_copy) CK(cudaStreamCreateWithFlags(&pl->stream_copy, cudaStreamNonBlocking));
        auto copy_columns = [&](int64_t b0, int64_t b1, cudaEvent_t after) -> int {
            const int64_t hi = std::min<int64_t>(b1, (int64_t)pl->max_samples);
            if (hi <= b0) return OWG_OK;
            CK(cudaStreamWaitEvent(pl->stream_copy, after, 0));
            CK(cudaMemcpy2DAsync(out + b0, (size_t)stride * sizeof(double), dout + b0, (size_t)stride * sizeof(double),
                                 (size_t)(hi - b0) * sizeof(double), (size_t)pl->n, cudaMemcpyDeviceToHost, pl->stream_copy));
            return OWG_OK;
        };
        if (!pl->warps_static.empty()) {  // static groups do not depend on the oscillator: they run while it settles
            const int nb = (int)pl->warps_static.size();
            const int64_t SCH = overlap_d2h ? 4 * CH_BASE : INT64_MAX / 2;  // chunked only to overlap the copy-back
            if (overlap_d2h) { if (int rc = pl->d_carry.alloc((size_t)nb * OWG_CARRY * 32)) return rc; }
            for (int64_t b0 = 0; b0 < (int64_t)pl->max_samples; b0 += SCH) {
                const int64_t b1 = b0 + SCH;
                cudaEvent_t e0 = chain_event(), e1 = chain_event();
                CK(cudaEventRecord(e0, s));
                if (pl->legacy)
                    chain_legacy_kernel<false><<<nb, 32, 0, s>>>(pl->d_warps_static.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->d_legacy_recs.p, nullptr,
                                                                 pl->d_group_rec_index.p, 0, dout, stride, pl->collect_diag ? pl->d_diag.p : nullptr, b0, b1,
                                                                 overlap_d2h ? pl->d_carry.p : nullptr, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end,
                                                                 pl->taps ? 1 : 0);
                else if (pl->use_tile && pl->collect_diag)
                    chain_tile_kernel<false, true><<<nb, OWG_TILE_THREADS, 0, s>>>(pl->d_warps_static.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                                  pl->d_static_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, 0, dout, stride, pl->d_diag.p,
                                                                                  b0, b1, overlap_d2h ? pl->d_carry.p : nullptr, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin,
                                                                                  pl->w_end, pl->taps ? 1 : 0, pl->tile_ipw);
                else if (pl->use_tile)
                    chain_tile_kernel<false, false><<<nb, OWG_TILE_THREADS, 0, s>>>(pl->d_warps_static.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                                   pl->d_static_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, 0, dout, stride, nullptr,
                                                                                   b0, b1, overlap_d2h ? pl->d_carry.p : nullptr, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin,
                                                                                   pl->w_end, pl->taps ? 1 : 0, pl->tile_ipw);
                else if (!pl->collect_diag && split_chain)
                    chain_split_kernel<false><<<nb, 64, 0, s>>>(pl->d_warps_static.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                pl->d_static_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, 0, dout, stride,
                                                                b0, b1, overlap_d2h ? pl->d_carry.p : nullptr, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin,
                                                                pl->w_end, swap_roles, pl->taps ? 1 : 0);
                else if (pl->collect_diag)
                    chain_kernel<false, true><<<nb, 32, 0, s>>>(pl->d_warps_static.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                 pl->d_static_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, 0, dout, stride, pl->d_diag.p,
                                                                 b0, b1, overlap_d2h ? pl->d_carry.p : nullptr, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin,
                                                                 pl->w_end);
                else
                    chain_kernel<false, false><<<nb, 32, 0, s>>>(pl->d_warps_static.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                  pl->d_static_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, 0, dout, stride, nullptr,
                                                                  b0, b1, overlap_d2h ? pl->d_carry.p : nullptr, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin,
                                                                  pl->w_end);
                CK(cudaGetLastError());
                CK(cudaEventRecord(e1, s));
                launches++;
                if (overlap_d2h) { if (int rc = copy_columns(b0, b1, e1)) return rc; }
            }
        }
        if (!pl->warps_trem.empty()) {
            const int nb = (int)pl->warps_trem.size();
            for (int64_t c = 0; c < n_chunks; c++) {
                const int64_t b0 = chunk_lo(c), b1 = chunk_lo(c + 1);
                CK(cudaStreamWaitEvent(s, pl->chunk_events[c], 0));
                if (voice_split_at > 0 && b0 == voice_split_at) CK(cudaStreamWaitEvent(s, pl->ev_voice2, 0));  // the tail of the voice rows
                // oscillator volts -> LDR law -> the preamp's resistance tracking, in place (off the oscillator's serial thread)
                tremolo_ldr_kernel<<<nt, 256, 0, s>>>(pl->d_groups.p, pl->d_trem_ids.p, nt, pl->d_pot_seq.p, pl->trem_n_os_max, pl->d_ldr_run.p,
                                                      2 * b0, 2 * b1, pl->legacy ? 1 : 0);
                CK(cudaGetLastError());
                launches++;
                if (!pl->legacy) {
                    dim3 grid((unsigned)((2 * CH_BASE + 63) / 64), (unsigned)nt);
                    tremolo_matrix_kernel<<<grid, 64, 0, s>>>(pl->d_groups.p, pl->d_trem_ids.p, nt, pl->d_pot_seq.p, pl->trem_n_os_max, pl->d_trem_recs.p,
                                                              pl->trem_n_os_max, 2 * b0, 2 * b1);
                    CK(cudaGetLastError());
                }
                cudaEvent_t e0 = chain_event(), e1 = chain_event();
                CK(cudaEventRecord(e0, s));
                if (pl->legacy)
                    chain_legacy_kernel<true><<<nb, 32, 0, s>>>(pl->d_warps_trem.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->d_legacy_recs.p, pl->d_pot_seq.p,
                                                                pl->d_group_rec_index.p, pl->trem_n_os_max, dout, stride, pl->collect_diag ? pl->d_diag.p : nullptr,
                                                                b0, b1, pl->d_carry.p, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end, pl->taps ? 1 : 0);
                else if (pl->use_tile && pl->collect_diag)
                    chain_tile_kernel<true, true><<<nb, OWG_TILE_THREADS, 0, s>>>(pl->d_warps_trem.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                                 pl->d_trem_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, pl->trem_n_os_max, dout, stride,
                                                                                 pl->d_diag.p, b0, b1, pl->d_carry.p, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end,
                                                                                 pl->taps ? 1 : 0, pl->tile_ipw);
                else if (pl->use_tile)
                    chain_tile_kernel<true, false><<<nb, OWG_TILE_THREADS, 0, s>>>(pl->d_warps_trem.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                                  pl->d_trem_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, pl->trem_n_os_max, dout, stride,
                                                                                  nullptr, b0, b1, pl->d_carry.p, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end,
                                                                                  pl->taps ? 1 : 0, pl->tile_ipw);
                else if (!pl->collect_diag && split_chain)
                    chain_split_kernel<true><<<nb, 64, 0, s>>>(pl->d_warps_trem.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                               pl->d_trem_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, pl->trem_n_os_max, dout, stride,
                                                               b0, b1, pl->d_carry.p, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end, swap_roles,
                                                               pl->taps ? 1 : 0);
                else if (pl->collect_diag)
                    chain_kernel<true, true><<<nb, 32, 0, s>>>(pl->d_warps_trem.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                pl->d_trem_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, pl->trem_n_os_max, dout, stride,
                                                                pl->d_diag.p, b0, b1, pl->d_carry.p, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end);
                else
                    chain_kernel<true, false><<<nb, 32, 0, s>>>(pl->d_warps_trem.p, pl->d_order.p, pl->d_cinit.p, pl->d_nsamp.p, pl->cache->d_settled,
                                                                 pl->d_trem_recs.p, pl->d_ans.p, pl->d_group_rec_index.p, pl->trem_n_os_max, dout, stride,
                                                                 nullptr, b0, b1, pl->d_carry.p, pl->metrics_ptr, pl->d_f0s.p, pl->w_begin, pl->w_end);
                CK(cudaGetLastError());
                CK(cudaEventRecord(e1, s));
                launches += pl->legacy ? 1 : 2;
                if (overlap_d2h) { if (int rc = copy_columns(b0, b1, e1)) return rc; }
            }
        }
        if (overlap_d2h) {  // join the copy stream into the main stream
            cudaEvent_t ej = chain_event();
            CK(cudaEventRecord(ej, pl->stream_copy));
            CK(cudaStreamWaitEvent(s, ej, 0));
            copied_by_chunks = true;
        }
        CK(cudaEventRecord(pl->evk1, s));
    } else {
        CK(cudaEventRecord(pl->evk0, s));
        CK(cudaEventRecord(pl->evk1, s));
    }
    if (out_location == OWG_OUT_HOST && !copied_by_chunks) {
        // one 2-D copy: rows of max_samples doubles, device pitch == host pitch == stride
        CK(cudaMemcpy2DAsync(out, (size_t)stride * sizeof(double), dout, (size_t)stride * sizeof(double),
                             (size_t)pl->max_samples * sizeof(double), (size_t)pl->n, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(pl->ev1, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventElapsedTime(&pl->total_ms, pl->ev0, pl->ev1));
    pl->main_ms = 0.f;  // device time of the chain kernel launches (busy time, not the wait for the oscillator)
    for (size_t k = 0; k + 1 < (chain_ev_used & ~(size_t)1); k += 2) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, pl->chain_ev[k], pl->chain_ev[k + 1]));
        pl->main_ms += ms;
    }
    (void)pl->evk0;
    pl->launches_last = launches;
    if (pl->collect_diag) {
        DevDiag h;
        CK(cudaMemcpy(&h, pl->d_diag.p, sizeof(h), cudaMemcpyDeviceToHost));
        owg_diag& d = g_last_diag;
        std::memset(&d, 0, sizeof(d));
        for (int i = 0; i < 16; i++) { d.nr_iter_hist[i] = h.main_hist[i]; d.shadow_nr_iter_hist[i] = h.sh_hist[i]; d.tremolo_nr_iter_hist[i] = h.trm_hist[i]; }
        for (int i = 0; i < 9; i++) d.poweramp_iter_hist[i] = h.pa_hist[i];
        d.nr_max_iter = h.main_nr_max; d.be_fallback = h.main_be; d.voltage_damp = h.main_damp; d.nan_reset = h.main_nan + h.adapter_nan;
        d.shadow_be_fallback = h.sh_be; d.shadow_nan_reset = h.sh_nan; d.tremolo_be_fallback = h.trm_be;
        d.kernels_launched = (uint64_t)launches;
    }
    return OWG_OK;
}

extern "C++" {
namespace {
// In-call multi-GPU fan-out (owg_opts.device_mask): contiguous job ranges balanced by rendered samples, one worker thread per
// selected GPU, each rendering its range through the single-device entry point straight into the caller's host rows.
template <class Job, class SampleCount, class RenderOne>
int fan_out_devices(const Job* jobs, int64_t n, double* out, int64_t stride, const owg_opts& o, SampleCount n_samples_of, RenderOne render_one) {
    const int ndev_avail = usable_devices();
    std::vector<int> devs;
    for (int d = 0; d < 32 && d < ndev_avail; d++) if (o.device_mask & (1u << d)) devs.push_back(d);
    if (devs.empty()) return fail(OWG_E_BAD_ARG, "device_mask selects no usable CUDA device");
    if (o.out_location != OWG_OUT_HOST) return fail(OWG_E_BAD_ARG, "device_mask fan-out needs host output (OWG_OUT_HOST)");
    if (o.stream) return fail(OWG_E_BAD_ARG, "device_mask fan-out uses library-owned streams (opts->stream must be NULL)");
    const int nd = (int)devs.size();
    double total = 0.0;
    for (int64_t i = 0; i < n; i++) total += (double)n_samples_of(jobs[i]);
    std::vector<int64_t> bounds(nd + 1, n);
    bounds[0] = 0;
    if (total <= 0.0) { for (int k = 1; k < nd; k++) bounds[k] = n * k / nd; }
    else {
        double acc = 0.0;
        int cut = 1;
        for (int64_t i = 0; i < n && cut < nd; i++) {
            acc += (double)n_samples_of(jobs[i]);
            while (cut < nd && acc >= total * cut / nd) bounds[cut++] = i + 1;
        }
    }
    std::vector<int> rcs(nd, OWG_OK);
    std::vector<std::string> errs(nd);
    std::vector<owg_diag> diags(nd);
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++) {
        th.emplace_back([&, k] {
            const int64_t b0 = bounds[k], b1 = bounds[k + 1];
            if (b1 <= b0) return;
            owg_opts ok = o;
            ok.device = devs[k];
            ok.device_mask = 0;
            rcs[k] = render_one(jobs + b0, b1 - b0, out + (size_t)b0 * (size_t)stride, stride, &ok);
            if (rcs[k] != OWG_OK) errs[k] = g_err;
            else if (o.collect_diag) diags[k] = g_last_diag;
        });
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < nd; k++) if (rcs[k] != OWG_OK) return fail(rcs[k], "device " + std::to_string(devs[k]) + ": " + errs[k]);
    {   // every GPU has written its rows up to ITS longest render; the call's contract is silence up to the longest render of the CALL
        double gmax = 0.0;
        for (int64_t i = 0; i < n; i++) gmax = std::max(gmax, std::floor((double)n_samples_of(jobs[i])));
        for (int k = 0; k < nd; k++) {
            double kmax = 0.0;
            for (int64_t i = bounds[k]; i < bounds[k + 1]; i++) kmax = std::max(kmax, std::floor((double)n_samples_of(jobs[i])));
            if (kmax < gmax)
                for (int64_t i = bounds[k]; i < bounds[k + 1]; i++)
                    std::memset(out + (size_t)i * (size_t)stride + (size_t)kmax, 0, (size_t)(gmax - kmax) * sizeof(double));
        }
    }
    if (o.collect_diag) {  // counters of a fanned-out call = sums over the GPUs
        owg_diag& d = g_last_diag;
        std::memset(&d, 0, sizeof(d));
        for (int k = 0; k < nd; k++) {
            const uint64_t* src = reinterpret_cast<const uint64_t*>(&diags[k]);
            uint64_t* dst = reinterpret_cast<uint64_t*>(&d);
            for (size_t w = 0; w < sizeof(owg_diag) / sizeof(uint64_t); w++) dst[w] += src[w];
        }
    }
    return OWG_OK;
}
}  // namespace
}  // extern "C++"

int owg_render_voices(const owg_voice_job* jobs, int64_t n, double* out, int64_t stride, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (opts && popcount32(opts->device_mask) >= 2 && n > 0 && jobs && out)
        return fan_out_devices(jobs, n, out, stride, *opts, [](const owg_voice_job& j) { const double x = j.duration_s * j.sample_rate; return x > 0.0 ? x : 0.0; },
                               [](const owg_voice_job* j, int64_t m, double* o, int64_t st, const owg_opts* op) { return owg_render_voices(j, m, o, st, op); });
    owg_plan* pl = nullptr;
    if (int rc = owg_plan_voices(jobs, n, opts, &pl)) return rc;
    const int rc = owg_plan_execute(pl, out, stride, opts ? opts->out_location : OWG_OUT_HOST);
    owg_plan_destroy(pl);
    return rc;
}

// ---- melange power amplifier (gen_power_amp.rs + power_amp.rs melange_adapter) ----------------------------------------
namespace {
// model of `sample_rate` and the settled state on `device` (cached for the life of the process)
int ensure_pa(int device, cudaStream_t stream, double sample_rate, const PaModel** d_model, const PaSettled** d_settled, int64_t* launches) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    DeviceCache& c = g_cache[device];
    auto model_for = [&](double sr, PaModel** out) -> int {
        uint64_t key; std::memcpy(&key, &sr, 8);
        auto it = c.d_pa_models.find(key);
        if (it == c.d_pa_models.end()) {
            std::vector<PaModel> h(1);
            owg::pa_build_model(sr, &h[0]);
            PaModel* d = nullptr;
            CK(cudaMalloc(&d, sizeof(PaModel)));
            CK(cudaMemcpy(d, h.data(), sizeof(PaModel), cudaMemcpyHostToDevice));
            g_h2d_bytes += sizeof(PaModel);
            it = c.d_pa_models.emplace(key, d).first;
        }
        *out = it->second;
        return OWG_OK;
    };
    if (!c.d_pa_settled) {
        PaModel* d_default = nullptr;
        if (int rc = model_for(88200.0, &d_default)) return rc;
        PaSettled* d = nullptr;
        CK(cudaMalloc(&d, 2 * sizeof(PaSettled)));  // [1]: the redundant second half-warp's copy
        pa_settle_kernel<<<1, 32, 0, stream>>>(d_default, d, d + 1);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        c.d_pa_settled = d;
        if (launches) *launches += 1;
    }
    PaModel* dm = nullptr;
    if (int rc = model_for(sample_rate, &dm)) return rc;
    *d_model = dm;
    *d_settled = c.d_pa_settled;
    return OWG_OK;
}

// the amplifier over device rows, in place; d_index / n_tiles select the instances of this launch
int launch_pa_rows(int device, cudaStream_t st, double sample_rate, double* d_rows, int64_t stride, const int32_t* d_index, int64_t n_tiles,
                   const unsigned long long* d_ns, int64_t n_samp_all, int rail_sag, const OwgChainInit* d_ci, double* d_rails, uint32_t* d_counters,
                   int64_t* launches) {
    if (n_tiles <= 0) return OWG_OK;
    const PaModel* dm = nullptr;
    const PaSettled* ds = nullptr;
    if (int rc = ensure_pa(device, st, sample_rate, &dm, &ds, launches)) return rc;
    const unsigned grid = (unsigned)((n_tiles + OWG_PA_TILES_PER_CTA - 1) / OWG_PA_TILES_PER_CTA);
    pa_melange_kernel<<<grid, OWG_PA_THREADS, 0, st>>>(dm, ds, d_rows, stride, d_index, n_tiles, d_ns, n_samp_all, rail_sag, d_ci, d_rails, d_counters);
    CK(cudaGetLastError());
    if (launches) *launches += 1;
    return OWG_OK;
}

int resolve_device(const owg_opts& o, int* dev_out) {
    if (usable_devices() <= 0) return fail(OWG_E_NO_DEVICE, "no usable CUDA device (libowgpu has no CPU fallback)");
    int dev = o.device;
    if (popcount32(o.device_mask) == 1) { dev = 0; while (!(o.device_mask & (1u << dev))) dev++; }
    if (dev < 0) CK(cudaGetDevice(&dev));
    CK(cudaSetDevice(dev));
    *dev_out = dev;
    return OWG_OK;
}

// chain B with the melange amplifier: voice + preamp through the usual plan (pre-amplifier rows on the device), then volume^2 -> melange
// PowerAmp::new() (44.1 kHz, main.rs:480) -> speaker in pa_melange_kernel
int render_bench_melange_pa(const owg_bench_job* jobs, int64_t n, double* out, int64_t stride, const owg_opts& opts) {
    if (n < 0 || (n > 0 && (!jobs || !out))) return fail(OWG_E_BAD_ARG, "owg_render_bench: bad argument");
    if (n == 0) return OWG_OK;
    owg_opts o = opts;
    const int rail_sag = opts.power_amp_model == OWG_POWER_AMP_MELANGE ? 1 : 0;
    o.power_amp_model = OWG_POWER_AMP_BEHAVIORAL;
    o.out_location = OWG_OUT_DEVICE;
    owg_plan* pl = nullptr;
    if (int rc = plan_bench_impl(jobs, n, &o, &pl, nullptr, true)) return rc;
    const int64_t dstride = std::max<int64_t>((int64_t)pl->max_samples, 1);
    int rc = stride < (int64_t)pl->max_samples ? fail(OWG_E_BAD_ARG, "stride smaller than the longest render") : OWG_OK;
    DevBuf<double> rows;
    DevBuf<OwgChainInit> d_ci;
    DevBuf<unsigned long long> d_ns;
    if (!rc) rc = rows.alloc((size_t)n * (size_t)dstride);
    if (!rc) rc = owg_plan_execute(pl, rows.p, dstride, OWG_OUT_DEVICE);
    cudaStream_t st = pl->stream;
    std::vector<OwgChainInit> ci((size_t)n);
    std::vector<unsigned long long> ns((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        owg::make_chain_init(jobs[i], 0, &ci[i]);
        const double nsd = jobs[i].v.duration_s * jobs[i].v.sample_rate;
        ns[i] = !(nsd == nsd) || nsd <= 0.0 ? 0ull : (unsigned long long)nsd;
    }
    if (!rc) rc = d_ci.upload(ci, st);
    if (!rc) rc = d_ns.upload(ns, st);
    int64_t launches = 0;
    if (!rc) rc = launch_pa_rows(pl->device, st, 44100.0, rows.p, dstride, nullptr, n, d_ns.p, 0, rail_sag, d_ci.p, nullptr, nullptr, &launches);
    if (!rc) {
        const cudaMemcpyKind kind = opts.out_location == OWG_OUT_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        if (cudaMemcpy2DAsync(out, (size_t)stride * sizeof(double), rows.p, (size_t)dstride * sizeof(double), (size_t)pl->max_samples * sizeof(double), (size_t)n,
                              kind, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            rc = fail(OWG_E_CUDA, std::string("melange power amplifier stage failed: ") + cudaGetErrorString(cudaGetLastError()));
    }
    g_last_diag.kernels_launched += (uint64_t)launches;
    owg_plan_destroy(pl);
    return rc;
}
}  // namespace

int64_t owg_release_caches(int32_t device) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    int64_t freed = 0;
    int prev = -1;
    const bool have_prev = cudaGetDevice(&prev) == cudaSuccess;
    for (auto& kv : g_cache) {
        if (device >= 0 && kv.first != device) continue;
        DeviceCache& c = kv.second;
        if (c.stage && !c.stage_in_use && cudaSetDevice(kv.first) == cudaSuccess) {
            if (cudaFree(c.stage) == cudaSuccess) freed += (int64_t)c.stage_bytes; else cudaGetLastError();
            c.stage = nullptr; c.stage_bytes = 0;
        }
    }
    if (have_prev) cudaSetDevice(prev);
    return freed;
}

int owg_power_amp_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, double sample_rate, int32_t rail_sag, double* out,
                        int64_t out_stride, double* rails, uint32_t* counters, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (n_inst < 0 || n_samp < 0 || !(sample_rate > 0.0) || !std::isfinite(sample_rate)) return fail(OWG_E_BAD_ARG, "owg_power_amp_batch: bad argument");
    if (n_inst == 0 || n_samp == 0) return OWG_OK;
    if (!in || !out || in_stride < n_samp || out_stride < n_samp) return fail(OWG_E_BAD_ARG, "owg_power_amp_batch: null buffer or stride < n_samp");
    owg_opts o;
    if (opts) o = *opts; else owg_default_opts(&o);
    int dev = 0;
    if (int rc = resolve_device(o, &dev)) return rc;
    cudaStream_t st = (cudaStream_t)o.stream;
    bool own = false;
    if (!st) { CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); own = true; }
    const bool on_device = o.out_location == OWG_OUT_DEVICE;
    DevBuf<double> rows, d_rails;
    DevBuf<uint32_t> d_cnt;
    int rc = rows.alloc((size_t)n_inst * (size_t)n_samp);
    if (!rc && rails) rc = d_rails.alloc((size_t)n_inst * 2);
    if (!rc && counters) rc = d_cnt.alloc((size_t)n_inst * 4);
    auto cuda_ok = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && !rc) rc = fail(OWG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); };
    const size_t w = (size_t)n_samp * sizeof(double);
    if (!rc) cuda_ok(cudaMemcpy2DAsync(rows.p, w, in, (size_t)in_stride * sizeof(double), w, (size_t)n_inst, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st), "input copy");
    if (!on_device) g_h2d_bytes += (int64_t)(w * (size_t)n_inst);
    int64_t launches = 0;
    if (!rc) rc = launch_pa_rows(dev, st, sample_rate, rows.p, n_samp, nullptr, n_inst, nullptr, n_samp, rail_sag ? 1 : 0, nullptr, rails ? d_rails.p : nullptr,
                                 counters ? d_cnt.p : nullptr, &launches);
    if (!rc) cuda_ok(cudaMemcpy2DAsync(out, (size_t)out_stride * sizeof(double), rows.p, w, w, (size_t)n_inst, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st), "output copy");
    if (!rc) cuda_ok(cudaStreamSynchronize(st), "melange power amplifier kernel");
    if (!rc && rails) cuda_ok(cudaMemcpy(rails, d_rails.p, (size_t)n_inst * 2 * sizeof(double), cudaMemcpyDeviceToHost), "rails copy");
    if (!rc && counters) cuda_ok(cudaMemcpy(counters, d_cnt.p, (size_t)n_inst * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost), "counters copy");
    if (own) cudaStreamDestroy(st);
    g_last_diag.kernels_launched = (uint64_t)launches;
    return rc;
}

int owg_render_bench(const owg_bench_job* jobs, int64_t n, double* out, int64_t stride, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (opts && popcount32(opts->device_mask) >= 2 && n > 0 && jobs && out)
        return fan_out_devices(jobs, n, out, stride, *opts, [](const owg_bench_job& j) { const double x = j.v.duration_s * j.v.sample_rate; return x > 0.0 ? x : 0.0; },
                               [](const owg_bench_job* j, int64_t m, double* o, int64_t st, const owg_opts* op) { return owg_render_bench(j, m, o, st, op); });
    if (opts && opts->power_amp_model != OWG_POWER_AMP_BEHAVIORAL) {
        if (opts->power_amp_model != OWG_POWER_AMP_MELANGE && opts->power_amp_model != OWG_POWER_AMP_MELANGE_IDEAL_RAILS)
            return fail(OWG_E_UNSUPPORTED, "unknown power_amp_model");
        return render_bench_melange_pa(jobs, n, out, stride, *opts);
    }
    owg_plan* pl = nullptr;
    if (int rc = owg_plan_bench(jobs, n, opts, &pl)) return rc;
    const int rc = owg_plan_execute(pl, out, stride, opts ? opts->out_location : OWG_OUT_HOST);
    owg_plan_destroy(pl);
    return rc;
}

int owg_render_bench_metrics(const owg_bench_job* jobs, int64_t n, double window_start_s, double window_end_s, double* metrics,
                             const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (n < 0 || (n > 0 && (!jobs || !metrics)) || !(window_end_s > window_start_s) || !(window_start_s >= 0.0))
        return fail(OWG_E_BAD_ARG, "owg_render_bench_metrics: bad argument");
    if (n == 0) return OWG_OK;
    const double sr = jobs[0].v.sample_rate;
    if (!(sr > 0.0) || !std::isfinite(sr)) return fail(OWG_E_BAD_ARG, "owg_render_bench_metrics: invalid sample_rate");
    const int64_t w0 = (int64_t)(window_start_s * sr), w1 = (int64_t)(window_end_s * sr);  // `(0.100 * BASE_SR) as usize`
    for (int64_t i = 0; i < n; i++) {
        if (jobs[i].v.sample_rate != sr) return fail(OWG_E_BAD_ARG, "owg_render_bench_metrics: all jobs must share one sample_rate");
        if (!(jobs[i].v.duration_s * sr >= (double)w1)) return fail(OWG_E_BAD_ARG, "owg_render_bench_metrics: analysis window exceeds a job's duration");
    }
    owg_opts o;
    if (opts) o = *opts; else owg_default_opts(&o);
    o.out_location = OWG_OUT_DEVICE;
    const int64_t BATCH = 65536;  // bounds the device scratch (voice samples) independently of n
    DevBuf<double> scratch, d_metrics;
    std::vector<double> raw;
    // Shared prefixes: jobs that differ only in the output stage (volume, speaker character, power-amp bypass) have the same voice and
    // the same preamp output.  When a sweep repeats prefixes, voice + preamp run once per distinct prefix and every job gets its own
    // output stage from post_stage_metrics_kernel (BASELINE config 4: 32 x 32 output-stage settings per key and tremolo depth).
    {
        typedef std::tuple<uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t> PKey;
        auto bits = [](double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; };
        std::map<PKey, int32_t> seen;
        std::vector<int32_t> prefix_of((size_t)n);
        std::vector<int64_t> first_job;
        for (int64_t i = 0; i < n; i++) {
            const owg_bench_job& j = jobs[i];
            const uint64_t small = (uint64_t)j.v.midi | ((uint64_t)j.v.mlp_enabled << 8) | ((uint64_t)j.v.attack_noise << 16) | ((uint64_t)j.v.flags << 24) |
                                   ((uint64_t)j.v.noise_seed << 32);
            const PKey k(small, bits(j.v.velocity), bits(j.v.duration_s), bits(j.v.ds_override), bits(j.r_ldr), bits(j.tremolo_depth),
                         (uint64_t)(j.no_preamp != 0));
            auto it = seen.find(k);
            if (it == seen.end()) { it = seen.emplace(k, (int32_t)first_job.size()).first; first_job.push_back(i); }
            prefix_of[i] = it->second;
        }
        const int64_t np = (int64_t)first_job.size();
        const char* dedup_env = getenv("OWG_SWEEP_DEDUP");
        if (np * 2 <= n && !(dedup_env && dedup_env[0] == '0')) {
            const double nwin = (double)(w1 - w0);
            for (int64_t p0 = 0; p0 < np; p0 += BATCH) {
                const int64_t npb = std::min<int64_t>(BATCH, np - p0);
                std::vector<owg_bench_job> pj((size_t)npb);
                for (int64_t k = 0; k < npb; k++) pj[k] = jobs[first_job[p0 + k]];
                owg_plan* pl = nullptr;
                if (int rc = plan_bench_impl(pj.data(), npb, &o, &pl, nullptr, true)) return rc;
                const int64_t stride = (int64_t)pl->max_samples;
                int rc = scratch.alloc((size_t)npb * (size_t)stride);
                if (!rc) rc = owg_plan_execute(pl, scratch.p, stride, OWG_OUT_DEVICE);
                cudaStream_t st = pl->stream;
                // the jobs of these prefixes, in prefix order so that a warp reads one prefix row
                std::vector<int64_t> members;
                for (int64_t i = 0; i < n; i++) if (prefix_of[i] >= p0 && prefix_of[i] < p0 + npb) members.push_back(i);
                std::stable_sort(members.begin(), members.end(), [&](int64_t a, int64_t b) { return prefix_of[a] < prefix_of[b]; });
                const int64_t nm = (int64_t)members.size();
                std::vector<OwgChainInit> ci((size_t)nm);
                std::vector<int32_t> pref((size_t)nm);
                std::vector<unsigned long long> ns((size_t)nm);
                std::vector<double> f0s((size_t)nm * 2);
                for (int64_t k = 0; k < nm; k++) {
                    const owg_bench_job& j = jobs[members[k]];
                    owg::make_chain_init(j, 0, &ci[k]);
                    pref[k] = prefix_of[members[k]] - (int32_t)p0;
                    const double nsd = j.v.duration_s * j.v.sample_rate;
                    ns[k] = !(nsd == nsd) || nsd <= 0.0 ? 0ull : (unsigned long long)nsd;
                    f0s[2 * k] = owg::note_frequency(j.v.midi); f0s[2 * k + 1] = sr;
                }
                DevBuf<OwgChainInit> d_ci; DevBuf<int32_t> d_pref; DevBuf<unsigned long long> d_ns; DevBuf<double> d_f0;
                if (!rc) rc = d_ci.upload(ci, st);
                if (!rc) rc = d_pref.upload(pref, st);
                if (!rc) rc = d_ns.upload(ns, st);
                if (!rc) rc = d_f0.upload(f0s, st);
                if (!rc) rc = d_metrics.alloc((size_t)nm * OWG_METRICS);
                if (!rc) {
                    post_stage_metrics_kernel<<<(unsigned)((nm + 127) / 128), 128, 0, st>>>(scratch.p, stride, d_pref.p, d_ci.p, d_ns.p, nm, d_metrics.p, d_f0.p, w0, w1);
                    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) rc = fail(OWG_E_CUDA, "post-stage kernel failed");
                }
                if (!rc) {
                    raw.resize((size_t)nm * OWG_METRICS);
                    if (cudaMemcpy(raw.data(), d_metrics.p, raw.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(OWG_E_CUDA, "metrics copy failed");
                }
                owg_plan_destroy(pl);
                if (rc) return rc;
                for (int64_t k = 0; k < nm; k++) {
                    const double* r = &raw[(size_t)k * OWG_METRICS];
                    double* m = metrics + (size_t)members[k] * OWG_METRIC_COLUMNS;
                    const double peak = r[0], mean_sq = r[1] / nwin;
                    const double h1 = 2.0 * std::sqrt((r[2] / nwin) * (r[2] / nwin) + (r[3] / nwin) * (r[3] / nwin));
                    const double h2 = 2.0 * std::sqrt((r[4] / nwin) * (r[4] / nwin) + (r[5] / nwin) * (r[5] / nwin));
                    m[0] = peak > 1e-15 ? 20.0 * std::log10(peak) : -120.0;
                    m[1] = mean_sq > 0.0 ? 10.0 * std::log10(mean_sq) : -120.0;
                    m[2] = h1 > 1e-15 ? 20.0 * std::log10(h2 / h1) : -120.0;
                    m[3] = peak; m[4] = mean_sq; m[5] = h1; m[6] = h2;
                }
            }
            return OWG_OK;
        }
    }
    for (int64_t b0 = 0; b0 < n; b0 += BATCH) {
        const int64_t nb = std::min<int64_t>(BATCH, n - b0);
        owg_plan* pl = nullptr;
        if (int rc = owg_plan_bench(jobs + b0, nb, &o, &pl)) return rc;
        std::vector<double> f0s((size_t)nb * 2);
        for (int64_t i = 0; i < nb; i++) { f0s[2 * i] = owg::note_frequency(jobs[b0 + i].v.midi); f0s[2 * i + 1] = sr; }
        int rc = pl->d_f0s.upload(f0s, pl->stream);
        const int64_t stride = (int64_t)pl->max_samples;
        if (!rc) rc = scratch.alloc((size_t)nb * (size_t)stride);
        if (!rc) rc = d_metrics.alloc((size_t)nb * OWG_METRICS);
        if (!rc && cudaMemsetAsync(d_metrics.p, 0, (size_t)nb * OWG_METRICS * sizeof(double), pl->stream) != cudaSuccess) rc = fail(OWG_E_CUDA, "memset failed");
        if (!rc) {
            pl->metrics_ptr = d_metrics.p;
            pl->w_begin = w0;
            pl->w_end = w1;
            rc = owg_plan_execute(pl, scratch.p, stride, OWG_OUT_DEVICE);
        }
        if (!rc) {
            raw.resize((size_t)nb * OWG_METRICS);
            if (cudaMemcpy(raw.data(), d_metrics.p, raw.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(OWG_E_CUDA, "metrics copy failed");
        }
        owg_plan_destroy(pl);
        if (rc) return rc;
        const double nwin = (double)(w1 - w0);
        for (int64_t i = 0; i < nb; i++) {
            const double* r = &raw[(size_t)i * OWG_METRICS];
            double* m = metrics + (size_t)(b0 + i) * OWG_METRIC_COLUMNS;
            const double peak = r[0], mean_sq = r[1] / nwin;
            const double h1 = 2.0 * std::sqrt((r[2] / nwin) * (r[2] / nwin) + (r[3] / nwin) * (r[3] / nwin));
            const double h2 = 2.0 * std::sqrt((r[4] / nwin) * (r[4] / nwin) + (r[5] / nwin) * (r[5] / nwin));
            m[0] = peak > 1e-15 ? 20.0 * std::log10(peak) : -120.0;
            m[1] = mean_sq > 0.0 ? 10.0 * std::log10(mean_sq) : -120.0;
            m[2] = h1 > 1e-15 ? 20.0 * std::log10(h2 / h1) : -120.0;
            m[3] = peak; m[4] = mean_sq; m[5] = h1; m[6] = h2;
        }
    }
    return OWG_OK;
}

void owg_default_calib_cfg(owg_calib_cfg* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->ds_at_c4 = 0.85; c->ds_exponent = 0.75; c->ds_clamp_lo = 0.02; c->ds_clamp_hi = 0.95; c->target_db = -35.0; c->voicing_slope = -0.04;
    c->zero_trim = 0;
}

int owg_render_calibrate(const owg_bench_job* jobs, int64_t n, const owg_calib_cfg* cfg_in, double window_start_s, double window_end_s,
                         double* rows, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (n < 0 || (n > 0 && (!jobs || !rows)) || !(window_end_s > window_start_s) || !(window_start_s >= 0.0))
        return fail(OWG_E_BAD_ARG, "owg_render_calibrate: bad argument");
    if (n == 0) return OWG_OK;
    owg_calib_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else owg_default_calib_cfg(&cfg);
    const double sr = jobs[0].v.sample_rate;
    if (!(sr > 0.0) || !std::isfinite(sr)) return fail(OWG_E_BAD_ARG, "owg_render_calibrate: invalid sample_rate");
    const int64_t w0 = (int64_t)(window_start_s * sr), w1 = (int64_t)(window_end_s * sr);
    for (int64_t i = 0; i < n; i++) {
        if (jobs[i].v.sample_rate != sr) return fail(OWG_E_BAD_ARG, "owg_render_calibrate: all jobs must share one sample_rate");
        if (!(jobs[i].v.duration_s * sr >= (double)w1)) return fail(OWG_E_BAD_ARG, "owg_render_calibrate: analysis window exceeds a job's duration");
    }
    owg_opts o;
    if (opts) o = *opts; else owg_default_opts(&o);
    o.out_location = OWG_OUT_DEVICE;
    o.collect_diag = 0;
    const int64_t BATCH = 444 * 31;  // one wave of the warp-specialised chain kernel (the taps live in its I/O warp)
    DevBuf<double> scratch, d_metrics;
    std::vector<double> raw;
    const double nwin = (double)(w1 - w0);
    auto db20 = [](double v) { return v > 1e-15 ? 20.0 * std::log10(v) : -120.0; };                    // to_dbfs, main.rs:2241-2247
    auto rms_db = [&](double sum_sq) { const double m = sum_sq / nwin; return m > 0.0 ? 10.0 * std::log10(m) : -120.0; };
    auto h2h1 = [&](const double* r) {  // h2_h1_ratio_db over dft_magnitude, main.rs:893-903, 928-936
        const double h1 = 2.0 * std::sqrt((r[2] / nwin) * (r[2] / nwin) + (r[3] / nwin) * (r[3] / nwin));
        const double h2 = 2.0 * std::sqrt((r[4] / nwin) * (r[4] / nwin) + (r[5] / nwin) * (r[5] / nwin));
        return h1 > 1e-15 ? 20.0 * std::log10(h2 / h1) : -120.0;
    };
    for (int64_t b0 = 0; b0 < n; b0 += BATCH) {
        const int64_t nb = std::min<int64_t>(BATCH, n - b0);
        // run_calibrate builds pickup and output gain from ITS CalibrationConfig (main.rs:1146, 1191): per-job overrides
        std::vector<owg_bench_job> jb(jobs + b0, jobs + b0 + nb);
        std::vector<double> gain((size_t)nb), f0s((size_t)nb * 2);
        for (int64_t i = 0; i < nb; i++) {
            jb[i].v.ds_override = owg::calib_displacement_scale(jb[i].v.midi, cfg);
            gain[i] = owg::calib_output_scale(jb[i].v.midi, jb[i].v.velocity, cfg);
            f0s[2 * i] = owg::note_frequency(jb[i].v.midi);
            f0s[2 * i + 1] = sr;
        }
        owg_plan* pl = nullptr;
        if (int rc = plan_bench_impl(jb.data(), nb, &o, &pl, gain.data())) return rc;
        int rc = pl->d_f0s.upload(f0s, pl->stream);
        const int64_t stride = (int64_t)pl->max_samples;
        if (!rc) rc = scratch.alloc((size_t)nb * (size_t)stride);
        if (!rc) rc = d_metrics.alloc((size_t)nb * OWG_METRICS);
        if (!rc && cudaMemsetAsync(d_metrics.p, 0, (size_t)nb * OWG_METRICS * sizeof(double), pl->stream) != cudaSuccess) rc = fail(OWG_E_CUDA, "memset failed");
        if (!rc && !(pl->use_tile || pl->use_split || pl->legacy)) rc = fail(OWG_E_UNSUPPORTED, "owg_render_calibrate: taps need the lane-tiled, the warp-specialised or the legacy chain kernel");
        if (!rc) {
            pl->metrics_ptr = d_metrics.p;
            pl->taps = true;
            pl->w_begin = w0;
            pl->w_end = w1;
            rc = owg_plan_execute(pl, scratch.p, stride, OWG_OUT_DEVICE);
        }
        if (!rc) {
            raw.resize((size_t)nb * OWG_METRICS);
            if (cudaMemcpy(raw.data(), d_metrics.p, raw.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(OWG_E_CUDA, "metrics copy failed");
        }
        owg_plan_destroy(pl);
        if (rc) return rc;
        for (int64_t i = 0; i < nb; i++) {
            const double* r = &raw[(size_t)i * OWG_METRICS];
            double* m = rows + (size_t)(b0 + i) * OWG_CALIBRATE_COLUMNS;
            const int midi = jb[i].v.midi;
            m[0] = cfg.ds_at_c4;
            m[1] = jb[i].v.ds_override;
            m[2] = r[OWG_MET_T1] * m[1];                                   // y_peak = reed_peak * ds_actual
            m[3] = db20(r[OWG_MET_T2]); m[4] = rms_db(r[OWG_MET_T2 + 1]); m[5] = h2h1(r + OWG_MET_T2);
            m[6] = db20(r[OWG_MET_T3]); m[7] = rms_db(r[OWG_MET_T3 + 1]);
            m[8] = db20(r[OWG_MET_T4]); m[9] = rms_db(r[OWG_MET_T4 + 1]); m[10] = h2h1(r + OWG_MET_T4);
            m[11] = db20(r[0]); m[12] = rms_db(r[1]); m[13] = h2h1(r);
            m[14] = 20.0 * std::log10(gain[i]);                            // proxy
            m[15] = cfg.zero_trim ? 0.0 : owg::register_trim_db(midi);
            m[16] = m[7] - cfg.target_db;                                  // proxy_error = t3_rms - target
            m[17] = m[8] - m[11];                                          // tanh_compression = t4_pk - t5_pk
        }
    }
    return OWG_OK;
}

int owg_render_engines(const owg_engine_job* jobs, int64_t n, float* out, int64_t stride, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (n < 0 || (n > 0 && (!jobs || !out))) return fail(OWG_E_BAD_ARG, "owg_render_engines: bad argument");
    if (n == 0) return OWG_OK;
    for (int64_t i = 0; i < n; i++) {
        const owg_engine_job& j = jobs[i];
        if (!(j.sample_rate > 0.0) || !std::isfinite(j.sample_rate) || !(j.duration_s >= 0.0) || !std::isfinite(j.duration_s) ||
            !std::isfinite(j.volume) || !std::isfinite(j.tremolo_depth) || !std::isfinite(j.speaker_character) || j.n_ev < 0 ||
            (j.n_ev > 0 && !j.ev))
            return fail(OWG_E_BAD_ARG, "owg_render_engines: invalid job");
        for (int64_t k = 1; k < j.n_ev; k++)
            if (j.ev[k].sample < j.ev[k - 1].sample) return fail(OWG_E_BAD_ARG, "owg_render_engines: events must be sorted by sample");
    }
    owg_plan pl;  // used for device / stream / cache plumbing only
    if (int rc = plan_common(&pl, opts)) return rc;
    const bool legacy = pl.legacy;
    cudaStream_t s = pl.stream;
    const int out_location = opts ? opts->out_location : OWG_OUT_HOST;
    const bool timing = getenv("OWG_ENGINE_TIMING") != nullptr;
    const auto t_host0 = std::chrono::steady_clock::now();

    auto trunc_u64 = [](double x) -> unsigned long long { return !(x == x) || x <= 0.0 ? 0ull : (x >= 18446744073709551615.0 ? ~0ull : (unsigned long long)x); };
    auto trunc_u32 = [](double x) -> uint32_t { return !(x == x) || x <= 0.0 ? 0u : (x >= 4294967295.0 ? 4294967295u : (uint32_t)x); };

    std::vector<EngineDesc> eng((size_t)n);
    std::vector<EngineGroup> groups;
    std::vector<EngineEvent> events;
    std::vector<OwgVoiceInit> vinits;
    std::vector<DamperRow> dampers;
    std::vector<int32_t> damper_sched((size_t)n);
    std::vector<SpkUpdate> spk_updates;
    std::vector<long long> spk_offsets;
    std::map<std::tuple<double, double, int, std::vector<std::pair<long long, double>>>, int> group_key;
    std::map<double, int> damper_key;
    // schedule = (rate, warm-up length, stream length class, character events, volume events) -> (schedule id, n updates)
    std::map<std::tuple<double, long long, std::vector<std::pair<long long, double>>, std::vector<std::pair<long long, double>>>, std::pair<int, int>> spk_key;
    std::vector<DepthEv> depth_events;
    long long max_samples = 0, max_block = 1, pot_stride = 0;
    for (int64_t i = 0; i < n; i++) {
        const owg_engine_job& j = jobs[i];
        EngineDesc& e = eng[i];
        const double sr = j.sample_rate;
        e.sample_rate = sr;
        e.volume_target = j.volume;
        e.n_samples = (long long)trunc_u64(sr * j.duration_s);
        e.n_warm = j.warm_up ? (long long)trunc_u64(sr * 0.6) : 0;
        e.block_size = j.block_size > 0 ? j.block_size : 512;
        e.oversample = sr < 88200.0 ? 1 : 0;
        e.ramp_samples = (int32_t)std::max<uint32_t>(trunc_u32(sr * 0.005), 1u);
        max_samples = std::max(max_samples, (long long)e.n_samples);
        max_block = std::max<long long>(max_block, e.block_size);
        const int sub = e.oversample ? 2 : 1;
        // parameter automation: set_volume / set_tremolo_depth / set_speaker_character events act at the start of their block
        std::vector<std::pair<long long, double>> ev_vol, ev_dep, ev_chr;
        ev_chr.emplace_back((long long)e.n_warm, j.speaker_character);  // the construction-time target, right after the warm-up
        for (int64_t k = 0; k < j.n_ev; k++) {
            const owg_event& ev = j.ev[k];
            if (ev.kind != OWG_EV_SET_VOLUME && ev.kind != OWG_EV_SET_TREMOLO_DEPTH && ev.kind != OWG_EV_SET_SPEAKER_CHARACTER) continue;
            if (ev.sample < 0 || ev.sample >= e.n_samples) continue;
            const long long at = (long long)e.n_warm + (ev.sample / e.block_size) * (long long)e.block_size;
            (ev.kind == OWG_EV_SET_VOLUME ? ev_vol : (ev.kind == OWG_EV_SET_TREMOLO_DEPTH ? ev_dep : ev_chr)).emplace_back(at, (double)ev.velocity);
        }
        // shared sequences: (rate, depth target, warm-up, depth automation)
        const auto gk = std::make_tuple(sr, j.tremolo_depth, j.warm_up ? 1 : 0, ev_dep);
        auto git = group_key.find(gk);
        if (git == group_key.end()) {
            EngineGroup g;
            std::memset(&g, 0, sizeof(g));
            g.sample_rate = sr;
            g.preamp_sr = e.oversample ? sr * 2.0 : sr;
            g.depth_target = j.tremolo_depth;
            g.n_warm_os = e.n_warm * sub;
            g.n_os = 0;
            g.oversample = e.oversample;
            g.ramp_samples = e.ramp_samples;
            g.use_defaults = std::fabs(g.preamp_sr - 48000.0) <= 0.5 ? 1 : 0;
            g.dep_ev_begin = (int32_t)depth_events.size();
            for (auto& de : ev_dep) depth_events.push_back(DepthEv{de.first * sub, de.second});
            g.dep_ev_end = (int32_t)depth_events.size();
            group_key[gk] = (int)groups.size();
            e.group = (int)groups.size();
            groups.push_back(g);
        } else e.group = git->second;
        groups[e.group].n_os = std::max<int64_t>(groups[e.group].n_os, e.n_samples * sub);
        // damper table per rate
        auto dit = damper_key.find(sr);
        if (dit == damper_key.end()) {
            damper_key[sr] = (int)(dampers.size() / 128);
            damper_sched[i] = (int32_t)(dampers.size() / 128);
            dampers.resize(dampers.size() + 128);
            owg::make_damper_rows(sr, &dampers[dampers.size() - 128]);
        } else damper_sched[i] = dit->second;
        // speaker coefficient / volume schedule per (rate, warm-up length, character events, volume events)
        const auto sk = std::make_tuple(sr, (long long)e.n_warm, ev_chr, ev_vol);
        auto sit = spk_key.find(sk);
        if (sit == spk_key.end()) {
            const int cap = (int)((e.ramp_samples + 8) * ev_chr.size() + ev_vol.size() + 8);
            std::vector<SpkUpdate> tmp((size_t)cap);
            std::vector<owg::AutoEvent> ac, av;
            for (auto& x : ev_chr) ac.push_back(owg::AutoEvent{x.first, x.second});
            for (auto& x : ev_vol) av.push_back(owg::AutoEvent{x.first, x.second});
            const int nu = owg::make_engine_schedule(sr, ac.data(), (int)ac.size(), av.data(), (int)av.size(), e.n_warm + e.n_samples + e.ramp_samples + 2,
                                                     (uint32_t)e.ramp_samples, tmp.data(), cap);
            if (nu < 0) return fail(OWG_E_CUDA, "speaker schedule overflow");
            spk_key[sk] = std::make_pair((int)spk_offsets.size(), nu);
            e.spk_sched = (int)spk_offsets.size();
            e.n_spk_updates = nu;
            spk_offsets.push_back((long long)spk_updates.size());
            spk_updates.insert(spk_updates.end(), tmp.begin(), tmp.begin() + nu);
        } else { e.spk_sched = sit->second.first; e.n_spk_updates = sit->second.second; }
    }
    if (stride < max_samples) return fail(OWG_E_BAD_ARG, "owg_render_engines: stride smaller than the longest stream");
    for (auto& g : groups) pot_stride = std::max<long long>(pot_stride, g.n_warm_os + g.n_os);
    // The oscillator's constructor (50 + 2*sr settle steps) and the engines' warm-up depend on the groups only: they start on their own
    // stream now and run while the host parameterises every note-on below.
    const int ng = (int)groups.size();
    DevBuf<EngineGroup> d_groups; DevBuf<double> d_pot; DevBuf<EngTrmRun> d_trmrun;
    cudaStream_t so = nullptr;
    CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t st; ~StreamGuard() { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } } } so_guard{so};
    {
        int rc0 = d_groups.upload(groups, so);
        if (!rc0) rc0 = d_pot.alloc((size_t)ng * (size_t)pot_stride);
        if (!rc0) rc0 = d_trmrun.alloc((size_t)ng);
        if (rc0) return rc0;
        CK(cudaMemsetAsync(d_trmrun.p, 0, (size_t)ng * sizeof(EngTrmRun), so));
        CK(cudaStreamSynchronize(so));  // the group table is read by kernels on all three streams
        engine_tremolo_kernel<<<ng, 32, 0, so>>>(d_groups.p, ng, d_pot.p, pot_stride, d_trmrun.p, 0);
        CK(cudaGetLastError());
    }
    for (int64_t i = 0; i < n; i++) {
        const owg_engine_job& j = jobs[i];
        EngineDesc& e = eng[i];
        const double sr = j.sample_rate;
        // events -> voice init records (WurliEngine::note_on, engine.rs:299-338)
        e.ev_begin = (long long)events.size();
        unsigned long long age = 0;
        for (int64_t k = 0; k < j.n_ev; k++) {
            const owg_event& ev = j.ev[k];
            EngineEvent d;
            d.sample = ev.sample;
            d.kind = ev.kind;
            d.note = ev.kind == OWG_EV_SUSTAIN ? ev.note : (ev.note < 33 ? 33 : (ev.note > 96 ? 96 : ev.note));
            d.vinit = -1;
            if (ev.kind == OWG_EV_NOTE_ON) {
                age += 1;
                owg_voice_job vj;
                std::memset(&vj, 0, sizeof(vj));
                vj.midi = (uint8_t)d.note;
                vj.mlp_enabled = j.mlp_enabled ? 1 : 0;
                vj.attack_noise = 1;
                vj.noise_seed = (uint32_t)d.note * 2654435761u + (uint32_t)age;
                vj.velocity = (double)ev.velocity;  // f32 -> f64 (engine.rs:330)
                vj.sample_rate = sr;
                vj.duration_s = 0.0;
                vj.ds_override = NAN;
                d.vinit = (long long)vinits.size();
                vinits.emplace_back();
                owg::make_voice_init(vj, &vinits.back());
            } else if (ev.kind == OWG_EV_SET_VOLUME || ev.kind == OWG_EV_SET_TREMOLO_DEPTH || ev.kind == OWG_EV_SET_SPEAKER_CHARACTER) {
                if (!std::isfinite(ev.velocity)) return fail(OWG_E_BAD_ARG, "owg_render_engines: non-finite parameter value");
                continue;  // handled through the schedules above, not by the voice state machine
            } else if (ev.kind != OWG_EV_NOTE_OFF && ev.kind != OWG_EV_SUSTAIN) return fail(OWG_E_BAD_ARG, "owg_render_engines: unknown event kind");
            events.push_back(d);
        }
        e.ev_end = (long long)events.size();
    }
    if (events.empty()) events.push_back(EngineEvent{0, OWG_EV_SUSTAIN, 0, -1});
    if (vinits.empty()) vinits.emplace_back();

    // rounds (= render() blocks) and segments (= rounds per chain launch; the mix buffers form a ring, one per segment in flight)
    long long n_rounds = 0;
    for (auto& e : eng) n_rounds = std::max<long long>(n_rounds, (e.n_samples + e.block_size - 1) / e.block_size);
    long long seg_rounds = std::max<long long>(1, 8192 / max_block);
    while (seg_rounds > 1 && (double)n * (double)seg_rounds * (double)max_block * 16.0 > 6.0e9) seg_rounds /= 2;
    const long long mix_stride = seg_rounds * max_block;
    const long long n_segs = (n_rounds + seg_rounds - 1) / seg_rounds;
    const long long ring = std::max<long long>(1, std::min<long long>(std::min<long long>(n_segs, 64),
                                                                      (long long)(8.0e9 / ((double)n * (double)mix_stride * 8.0))));

    // chain warps: engines of one (group, block size) share a warp (common record index and round boundaries)
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, pl.device);
    int lanes = (int)((n + (long long)sm_count * 4 - 1) / ((long long)sm_count * 4));
    lanes = std::min(31, std::max(1, lanes));
    // warp-specialised chain (DK warp + I/O warp per CTA) while the batch fits one wave of 3 CTAs per SM; melange model only
    bool engine_split = false;
    {
        const char* env = getenv("OWG_CHAIN_SPLIT");
        const int lanes_split = (int)((n + (long long)sm_count * 3 - 1) / ((long long)sm_count * 3));
        if (!legacy && !(env && env[0] == '0') && lanes_split <= 31) { engine_split = true; lanes = std::max(1, lanes_split); }
    }
    if (const char* env = getenv("OWG_ENGINE_LANES")) lanes = std::min(31, std::max(1, atoi(env)));
    std::vector<int32_t> eorder((size_t)n);
    for (int64_t i = 0; i < n; i++) eorder[i] = (int32_t)i;
    std::stable_sort(eorder.begin(), eorder.end(), [&](int32_t a, int32_t b) {
        if (eng[a].group != eng[b].group) return eng[a].group < eng[b].group;
        if (eng[a].block_size != eng[b].block_size) return eng[a].block_size < eng[b].block_size;
        return eng[a].n_samples > eng[b].n_samples;
    });
    std::vector<EngineWarp> ewarps;
    for (int64_t i = 0; i < n;) {
        EngineWarp w;
        w.group = eng[eorder[i]].group; w.block_size = eng[eorder[i]].block_size; w.first = (int32_t)i; w.count = 0; w.n_max = 0;
        while (i < n && w.count < lanes && eng[eorder[i]].group == w.group && eng[eorder[i]].block_size == w.block_size) {
            w.n_max = std::max<long long>(w.n_max, eng[eorder[i]].n_samples);
            w.count++; i++;
        }
        ewarps.push_back(w);
    }

    DevBuf<EngineDesc> d_eng; DevBuf<EngineEvent> d_events; DevBuf<OwgVoiceInit> d_vinits;
    DevBuf<DamperRow> d_dampers; DevBuf<int32_t> d_dsched, d_eorder; DevBuf<SpkUpdate> d_spk; DevBuf<long long> d_spkoff;
    DevBuf<double> d_recs, d_ans, d_mix; DevBuf<DkState> d_post, d_shadow; DevBuf<VoiceRT> d_pool; DevBuf<float> d_out;
    DevBuf<EngineState> d_states; DevBuf<EngineChainState> d_chains; DevBuf<EngineWarp> d_ewarps; DevBuf<EngLdrRun> d_ldrrun; DevBuf<DepthEv> d_depev; DevBuf<double> d_depth, d_lgrecs, d_glast;
    DevBuf<LgState> d_post_lg, d_shadow_lg;
    DevBuf<EngineDiag> d_diag;
    int rc = d_eng.upload(eng, s);
    if (!rc) rc = d_events.upload(events, s);
    if (!rc) rc = d_vinits.upload(vinits, s);
    if (!rc) rc = d_dampers.upload(dampers, s);
    if (!rc) rc = d_dsched.upload(damper_sched, s);
    if (!rc) rc = d_spk.upload(spk_updates, s);
    if (!rc) rc = d_spkoff.upload(spk_offsets, s);
    if (!rc) rc = d_eorder.upload(eorder, s);
    if (!rc) rc = d_ewarps.upload(ewarps, s);
    if (legacy) {
        std::vector<double> lgrecs((size_t)ng * OWG_LG_STRIDE);
        for (int g = 0; g < ng; g++) owg::make_legacy_group(groups[g].preamp_sr, NAN, &lgrecs[(size_t)g * OWG_LG_STRIDE]);
        if (!rc) rc = d_lgrecs.upload(lgrecs, s);
        if (!rc) rc = d_glast.alloc((size_t)ng);
        if (!rc) rc = d_post_lg.alloc((size_t)ng);
        if (!rc) rc = d_shadow_lg.alloc(ewarps.size());
    } else {
        if (!rc) rc = d_recs.alloc((size_t)ng * (size_t)pot_stride * OWG_MAT_STRIDE);
        if (!rc) rc = d_ans.alloc((size_t)ng * OWG_AN_SPARSE);
        if (!rc) rc = d_post.alloc((size_t)ng);
        if (!rc) rc = d_shadow.alloc(ewarps.size());
    }
    if (!rc) rc = d_ldrrun.alloc((size_t)ng);
    if (depth_events.empty()) depth_events.push_back(DepthEv{-1, 0.0});
    if (!rc) rc = d_depev.upload(depth_events, s);
    if (!rc) rc = d_depth.alloc((size_t)ng * (size_t)pot_stride);
    if (!rc) rc = d_pool.alloc((size_t)n * 128);
    if (!rc) rc = d_states.alloc((size_t)n);
    if (!rc) rc = d_chains.alloc((size_t)n);
    if (!rc) rc = d_mix.alloc((size_t)ring * (size_t)n * (size_t)mix_stride);
    if (!rc) rc = d_diag.alloc(1);
    float* dout = out;
    if (!rc && out_location == OWG_OUT_HOST) { rc = d_out.alloc((size_t)n * (size_t)stride); dout = d_out.p; }
    if (rc) return rc;
    // three streams: `s` renders voices round by round; `so` runs the serial Twin-T oscillator one chunk ahead; `sc` builds the
    // chunk's DK matrices and runs the chain (with the shadow solve in lane 31) behind both
    cudaStream_t sc = nullptr;
    CK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
    StreamGuard sc_guard{sc};
    std::vector<cudaEvent_t> evs;
    struct EventGuard { std::vector<cudaEvent_t>* v; ~EventGuard() { for (auto e : *v) cudaEventDestroy(e); } } ev_guard{&evs};
    auto new_event = [&](cudaEvent_t* e) -> cudaError_t { cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming); if (r == cudaSuccess) evs.push_back(*e); return r; };
    cudaEvent_t ev_up;
    CK(new_event(&ev_up));
    cudaEvent_t tv[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const auto t_host1 = std::chrono::steady_clock::now();
    if (timing) for (auto& e : tv) { CK(cudaEventCreate(&e)); evs.push_back(e); }
    CK(cudaMemsetAsync(d_diag.p, 0, sizeof(EngineDiag), s));
    CK(cudaMemsetAsync(d_ldrrun.p, 0, (size_t)ng * sizeof(EngLdrRun), s));
    CK(cudaMemsetAsync(d_pool.p, 0, (size_t)n * 128 * sizeof(VoiceRT), s));
    CK(cudaEventRecord(ev_up, s));
    CK(cudaStreamWaitEvent(sc, ev_up, 0));
    if (timing) { CK(cudaEventRecord(tv[0], so)); CK(cudaEventRecord(tv[3], s)); CK(cudaEventRecord(tv[2], sc)); }
    const unsigned eb = (unsigned)((n + 63) / 64);
    engine_init_kernel<<<eb, 64, 0, s>>>(d_eng.p, (int)n, d_groups.p, nullptr, d_states.p, nullptr, nullptr, nullptr);
    CK(cudaGetLastError());
    const double silent_thr = owg::silent_threshold();
    int64_t launches = 1;
    const long long n_chunks = std::max<long long>(n_segs, 1);
    std::vector<cudaEvent_t> ev_voices((size_t)n_chunks), ev_chain((size_t)n_chunks), ev_osc((size_t)n_chunks);
    // oscillator chunks: chunk 0 = Tremolo::new (warm-up + settle) + the engines' warm-up + segment 0
    for (long long sg = 0; sg < n_chunks; sg++) {
        const long long r1 = std::min(n_rounds, (sg + 1) * seg_rounds);
        engine_tremolo_kernel<<<ng, 32, 0, so>>>(d_groups.p, ng, d_pot.p, pot_stride, d_trmrun.p, r1 * max_block);
        CK(cudaGetLastError());
        CK(new_event(&ev_osc[sg]));
        CK(cudaEventRecord(ev_osc[sg], so));
        launches += 1;
    }
    if (timing) CK(cudaEventRecord(tv[1], so));
    long long max_warm_os = 0;
    for (auto& g : groups) max_warm_os = std::max<long long>(max_warm_os, g.n_warm_os);
    for (long long sg = 0; sg < n_chunks; sg++) {
        const long long r0 = sg * seg_rounds, r1 = std::min(n_rounds, r0 + seg_rounds);
        double* mixbuf = d_mix.p + (size_t)(sg % ring) * (size_t)n * (size_t)mix_stride;
        if (sg >= ring) CK(cudaStreamWaitEvent(s, ev_chain[sg - ring], 0));  // the chain has consumed this mix buffer
        for (long long r = r0; r < r1; r++) {
            engine_events_kernel<<<eb, 64, 0, s>>>(d_eng.p, (int)n, r, d_events.p, d_vinits.p, d_dampers.p, d_dsched.p, d_pool.p, d_states.p);
            engine_voice_mix_kernel<<<(unsigned)n, OWG_ENGINE_ITEMS, 0, s>>>(d_eng.p, r, d_pool.p, d_states.p, mixbuf, mix_stride, r0);
            engine_post_kernel<<<eb, 64, 0, s>>>(d_eng.p, (int)n, r, d_pool.p, d_states.p, mixbuf, mix_stride, r0, silent_thr);
            launches += 3;
        }
        CK(cudaGetLastError());
        CK(new_event(&ev_voices[sg]));
        CK(cudaEventRecord(ev_voices[sg], s));
        // matrices of this chunk, then (first chunk) the warm-up solve and the chain states
        CK(cudaStreamWaitEvent(sc, ev_osc[sg], 0));
        engine_ldr_kernel<<<ng, 256, 0, sc>>>(d_groups.p, ng, d_pot.p, d_depth.p, pot_stride, d_ldrrun.p, r1 * max_block, legacy ? 1 : 0, d_depev.p);
        CK(cudaGetLastError());
        launches += 1;
        if (legacy) {
            if (sg == 0) {
                engine_shadow_legacy_kernel<<<ng, 32, 0, sc>>>(d_groups.p, ng, d_lgrecs.p, d_pot.p, pot_stride, d_post_lg.p, d_glast.p);
                CK(cudaGetLastError());
                engine_init_kernel<<<eb, 64, 0, sc>>>(d_eng.p, (int)n, d_groups.p, nullptr, nullptr, d_chains.p, d_post_lg.p, d_glast.p);
                CK(cudaGetLastError());
                launches += 2;
            }
            CK(cudaStreamWaitEvent(sc, ev_voices[sg], 0));
            engine_chain_legacy_kernel<<<(unsigned)ewarps.size(), 32, 0, sc>>>(d_ewarps.p, d_eorder.p, d_eng.p, r0, r1, d_spk.p, d_spkoff.p, d_groups.p,
                                                                              d_post_lg.p, d_glast.p, d_lgrecs.p, d_pot.p, pot_stride, d_chains.p,
                                                                              d_shadow_lg.p, mixbuf, mix_stride, dout, stride, max_samples,
                                                                              sg + 1 == n_chunks ? 1 : 0);
            CK(cudaGetLastError());
            launches += 1;
            CK(new_event(&ev_chain[sg]));
            CK(cudaEventRecord(ev_chain[sg], sc));
            continue;
        }
        {
            const long long chunk_len = (sg == 0 ? max_warm_os : 0) + (r1 - r0) * max_block * 2;
            dim3 grid((unsigned)std::max<long long>(1, (chunk_len + 63) / 64), (unsigned)ng);
            engine_matrix_kernel<<<grid, 64, 0, sc>>>(d_groups.p, ng, d_pot.p, pot_stride, d_recs.p, pot_stride, d_ans.p, r0 * max_block, r1 * max_block);
            CK(cudaGetLastError());
            launches += 1;
        }
        if (sg == 0) {
            engine_shadow_kernel<<<ng, 32, 0, sc>>>(d_groups.p, ng, pl.cache->d_settled, d_recs.p, pot_stride, d_ans.p, d_post.p);
            CK(cudaGetLastError());
            engine_init_kernel<<<eb, 64, 0, sc>>>(d_eng.p, (int)n, d_groups.p, d_post.p, nullptr, d_chains.p, nullptr, nullptr);
            CK(cudaGetLastError());
            launches += 2;
        }
        CK(cudaStreamWaitEvent(sc, ev_voices[sg], 0));
        if (engine_split)
            engine_chain_split_kernel<<<(unsigned)ewarps.size(), 64, 0, sc>>>(d_ewarps.p, d_eorder.p, d_eng.p, r0, r1, d_spk.p, d_spkoff.p, d_groups.p,
                                                                             d_post.p, d_recs.p, pot_stride, d_ans.p, d_chains.p, d_shadow.p, mixbuf,
                                                                             mix_stride, dout, stride, max_samples, sg + 1 == n_chunks ? 1 : 0);
        else
            engine_chain_kernel<<<(unsigned)ewarps.size(), 32, 0, sc>>>(d_ewarps.p, d_eorder.p, d_eng.p, r0, r1, d_spk.p, d_spkoff.p, d_groups.p, d_post.p,
                                                                       d_recs.p, pot_stride, d_ans.p, d_chains.p, d_shadow.p, mixbuf, mix_stride, dout, stride,
                                                                       max_samples, sg + 1 == n_chunks ? 1 : 0);
        CK(cudaGetLastError());
        launches += 1;
        CK(new_event(&ev_chain[sg]));
        CK(cudaEventRecord(ev_chain[sg], sc));
    }
    {
        if (timing) CK(cudaEventRecord(tv[4], s));
        cudaEvent_t ev_s_done;
        CK(new_event(&ev_s_done));
        CK(cudaEventRecord(ev_s_done, s));
        CK(cudaStreamWaitEvent(sc, ev_s_done, 0));
    }
    engine_diag_kernel<<<eb, 64, 0, sc>>>(d_states.p, d_chains.p, (int)n, d_diag.p);
    CK(cudaGetLastError());
    if (out_location == OWG_OUT_HOST)
        CK(cudaMemcpy2DAsync(out, (size_t)stride * sizeof(float), dout, (size_t)stride * sizeof(float), (size_t)max_samples * sizeof(float),
                             (size_t)n, cudaMemcpyDeviceToHost, sc));
    if (timing) CK(cudaEventRecord(tv[5], sc));
    CK(cudaStreamSynchronize(sc));
    CK(cudaStreamSynchronize(s));
    if (timing) {
        float a = 0, b = 0, c = 0, d = 0;
        cudaEventElapsedTime(&a, tv[0], tv[1]); cudaEventElapsedTime(&c, tv[3], tv[4]); cudaEventElapsedTime(&d, tv[2], tv[5]);
        (void)b;
        fprintf(stderr, "[owg engines] host prep %.1f ms (%zu note-ons), oscillator stream %.1f ms, voices stream %.1f ms, matrices+chain stream (incl. waits, copy) %.1f ms, rounds %lld, segments %lld (ring %lld), lanes %d, warps %zu\n",
                std::chrono::duration<double, std::milli>(t_host1 - t_host0).count(), vinits.size(), a, c, d, n_rounds, n_segs, ring, lanes, ewarps.size());
    }
    {
        EngineDiag h;
        CK(cudaMemcpy(&h, d_diag.p, sizeof(h), cudaMemcpyDeviceToHost));
        owg_diag& d = g_last_diag;
        std::memset(&d, 0, sizeof(d));
        d.nan_reset = h.nan_guard + h.out_nan;
        d.kernels_launched = launches;
        // engine-specific counters are reported through the generic histogram slots: [0]=note-ons, [1]=steals, [2]=voices freed, [3]=max active voices
        d.nr_iter_hist[0] = h.note_ons; d.nr_iter_hist[1] = h.steals; d.nr_iter_hist[2] = h.voices_freed; d.nr_iter_hist[3] = h.max_active;
    }
    return OWG_OK;
}

int owg_preamp_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, double fs_base, int oversample,
                     double tremolo_depth, double r_ldr_static, double* out, int64_t out_stride, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (n_inst < 0 || n_samp < 0 || !(fs_base > 0.0) || !std::isfinite(fs_base) || std::isnan(tremolo_depth))
        return fail(OWG_E_BAD_ARG, "owg_preamp_batch: bad argument");
    if (n_inst == 0 || n_samp == 0) return OWG_OK;
    if (!in || !out || in_stride < n_samp || out_stride < n_samp) return fail(OWG_E_BAD_ARG, "owg_preamp_batch: null buffer or stride < n_samp");
    owg_plan* pl = new owg_plan();
    g_h2d_bytes = 0;
    pl->kind = 2;
    pl->n = n_inst;
    if (int rc = plan_common(pl, opts)) { delete pl; return rc; }
    std::vector<InstSpec> specs((size_t)n_inst);
    for (int64_t i = 0; i < n_inst; i++) {
        InstSpec& sp = specs[i];
        sp.fs = fs_base;
        sp.oversample = oversample ? 1 : 0;
        sp.n_samples = (unsigned long long)n_samp;
        sp.depth = tremolo_depth;
        sp.r_ldr = r_ldr_static;
        std::memset(&sp.ci, 0, sizeof(sp.ci));
        sp.ci.spk_norm = 1.0;
        sp.ci.no_poweramp = 1;
        sp.ci.pre_only = 1;
    }
    std::vector<int32_t> order;
    build_groups_and_warps(pl, specs, &order);
    std::vector<OwgChainInit> ci((size_t)n_inst);
    for (int64_t i = 0; i < n_inst; i++) ci[i] = specs[i].ci;
    int rc = launch_tremolo_ctor(pl);
    if (!rc) rc = upload_chain_plan(pl, ci, order);
    if (!rc) {
        pl->in_ptr = in;
        pl->in_stride = in_stride;
        rc = owg_plan_execute(pl, out, out_stride, opts ? opts->out_location : OWG_OUT_HOST);
    }
    owg_plan_destroy(pl);
    return rc;
}

static int chain_batch_impl(const double* in, int in_location, int64_t in_stride, int64_t n_inst, const unsigned long long* n_samp_each,
                            int64_t n_samp_max, const owg_bench_job* params, int32_t init_order, double* out, int64_t out_stride,
                            const owg_opts* opts, bool force_pre_only = false) {
    if (!force_pre_only && opts && opts->power_amp_model != OWG_POWER_AMP_BEHAVIORAL) {
        // the melange amplifier (render-poly / render-midi built with --no-default-features, main.rs:1470-1481, 1756-1889): the chain up to
        // the preamp output into device rows, then volume^2 -> PowerAmp::new() -> speaker in pa_melange_kernel
        if (opts->power_amp_model != OWG_POWER_AMP_MELANGE && opts->power_amp_model != OWG_POWER_AMP_MELANGE_IDEAL_RAILS)
            return fail(OWG_E_UNSUPPORTED, "unknown power_amp_model");
        owg_opts o = *opts;
        o.power_amp_model = OWG_POWER_AMP_BEHAVIORAL;
        o.out_location = OWG_OUT_DEVICE;
        int dev = 0;
        if (int rc = resolve_device(o, &dev)) return rc;
        o.device = dev; o.device_mask = 0;
        cudaStream_t st = (cudaStream_t)o.stream;
        bool own = false;
        if (!st) { CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); own = true; o.stream = st; }
        DevBuf<double> rows;
        DevBuf<OwgChainInit> d_ci;
        DevBuf<unsigned long long> d_ns;
        int rc = rows.alloc((size_t)n_inst * (size_t)n_samp_max);
        const int in_loc = in_location >= 0 ? in_location : opts->out_location;
        if (!rc) rc = chain_batch_impl(in, in_loc, in_stride, n_inst, n_samp_each, n_samp_max, params, init_order, rows.p, n_samp_max, &o, true);
        std::vector<OwgChainInit> ci((size_t)n_inst);
        std::vector<unsigned long long> ns((size_t)n_inst);
        for (int64_t i = 0; i < n_inst; i++) {
            owg::make_chain_init(params[i], 0, &ci[i]);
            ns[i] = n_samp_each ? n_samp_each[i] : (unsigned long long)n_samp_max;
        }
        if (!rc) rc = d_ci.upload(ci, st);
        if (!rc) rc = d_ns.upload(ns, st);
        int64_t launches = 0;
        if (!rc) rc = launch_pa_rows(dev, st, 44100.0, rows.p, n_samp_max, nullptr, n_inst, d_ns.p, 0, opts->power_amp_model == OWG_POWER_AMP_MELANGE ? 1 : 0,
                                     d_ci.p, nullptr, nullptr, &launches);
        if (!rc) {
            const cudaMemcpyKind kind = opts->out_location == OWG_OUT_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            if (cudaMemcpy2DAsync(out, (size_t)out_stride * sizeof(double), rows.p, (size_t)n_samp_max * sizeof(double), (size_t)n_samp_max * sizeof(double),
                                  (size_t)n_inst, kind, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
                rc = fail(OWG_E_CUDA, std::string("melange power amplifier stage failed: ") + cudaGetErrorString(cudaGetLastError()));
        }
        if (own) cudaStreamDestroy(st);
        g_last_diag.kernels_launched += (uint64_t)launches;
        return rc;
    }
    owg_plan* pl = new owg_plan();
    g_h2d_bytes = 0;
    pl->kind = 2;
    pl->n = n_inst;
    if (int rc = plan_common(pl, opts)) { delete pl; return rc; }
    std::vector<InstSpec> specs((size_t)n_inst);
    for (int64_t i = 0; i < n_inst; i++) {
        InstSpec& sp = specs[i];
        const owg_bench_job& j = params[i];
        sp.init_order = init_order;
        sp.fs = j.v.sample_rate;
        sp.oversample = j.v.sample_rate < 88200.0 ? 1 : 0;
        sp.n_samples = n_samp_each ? n_samp_each[i] : (unsigned long long)n_samp_max;
        sp.depth = j.tremolo_depth;
        sp.r_ldr = j.r_ldr;
        owg::make_chain_init(j, 0, &sp.ci);
        if (force_pre_only) sp.ci.pre_only = 1;
    }
    std::vector<int32_t> order;
    build_groups_and_warps(pl, specs, &order);
    std::vector<OwgChainInit> ci((size_t)n_inst);
    for (int64_t i = 0; i < n_inst; i++) ci[i] = specs[i].ci;
    int rc = launch_tremolo_ctor(pl);
    if (!rc) rc = upload_chain_plan(pl, ci, order);
    if (!rc) {
        pl->in_ptr = in;
        pl->in_stride = in_stride;
        pl->in_location = in_location;
        rc = owg_plan_execute(pl, out, out_stride, opts ? opts->out_location : OWG_OUT_HOST);
    }
    owg_plan_destroy(pl);
    return rc;
}

int owg_chain_batch(const double* in, int64_t in_stride, int64_t n_inst, int64_t n_samp, const owg_bench_job* params, int32_t init_order,
                    double* out, int64_t out_stride, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (n_inst < 0 || n_samp < 0 || (init_order != OWG_INIT_RESET_THEN_SET && init_order != OWG_INIT_SET_THEN_RESET))
        return fail(OWG_E_BAD_ARG, "owg_chain_batch: bad argument");
    if (n_inst == 0 || n_samp == 0) return OWG_OK;
    if (!in || !out || !params || in_stride < n_samp || out_stride < n_samp) return fail(OWG_E_BAD_ARG, "owg_chain_batch: null buffer or stride < n_samp");
    for (int64_t i = 0; i < n_inst; i++) {
        const owg_bench_job& j = params[i];
        if (!(j.v.sample_rate > 0.0) || !std::isfinite(j.v.sample_rate) || !std::isfinite(j.volume) || !std::isfinite(j.speaker_character) ||
            std::isnan(j.tremolo_depth))
            return fail(OWG_E_BAD_ARG, "owg_chain_batch: invalid chain parameters");
    }
    return chain_batch_impl(in, -1, in_stride, n_inst, nullptr, n_samp, params, init_order, out, out_stride, opts);
}

int owg_render_midi(const owg_midi_job* jobs, int64_t n, double* out, int64_t stride, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (n < 0 || (n > 0 && (!jobs || !out))) return fail(OWG_E_BAD_ARG, "owg_render_midi: bad argument");
    if (n == 0) return OWG_OK;
    const double SR = 44100.0;  // BASE_SR: the tool renders at 44.1 kHz only
    const int CHUNK = 64;
    int64_t max_samples = 0;
    for (int64_t i = 0; i < n; i++) {
        const owg_midi_job& j = jobs[i];
        if (j.n_ev < 0 || (j.n_ev > 0 && !j.ev) || j.n_samples < 0 || !std::isfinite(j.volume) || !std::isfinite(j.speaker_character))
            return fail(OWG_E_BAD_ARG, "owg_render_midi: invalid job");
        for (int64_t k = 0; k < j.n_ev; k++) {
            if (!std::isfinite(j.ev[k].time_s) || (k > 0 && j.ev[k].time_s < j.ev[k - 1].time_s)) return fail(OWG_E_BAD_ARG, "owg_render_midi: events must be finite and sorted by time");
            if (j.ev[k].kind > OWG_MIDI_PEDAL) return fail(OWG_E_BAD_ARG, "owg_render_midi: unknown event kind");
        }
        max_samples = std::max<int64_t>(max_samples, j.n_samples);
    }
    if (max_samples == 0) return OWG_OK;
    if (stride < max_samples) return fail(OWG_E_BAD_ARG, "owg_render_midi: stride smaller than the longest stream");
    owg_plan pl;  // device / stream / cache plumbing of the voice phase
    owg_opts o_plan;
    if (opts) o_plan = *opts; else owg_default_opts(&o_plan);
    o_plan.power_amp_model = OWG_POWER_AMP_BEHAVIORAL;  // the amplifier model concerns the chain stage only (chain_batch_impl below)
    if (int rc = plan_common(&pl, &o_plan)) return rc;
    cudaStream_t s = pl.stream;
    std::vector<EngineDesc> eng((size_t)n);
    std::vector<MidiEvent> events;
    std::vector<OwgVoiceInit> vinits;
    std::vector<long long> held_offset((size_t)n);
    std::vector<owg_bench_job> params((size_t)n);
    std::vector<unsigned long long> n_each((size_t)n);
    long long held_total = 0;
    for (int64_t i = 0; i < n; i++) {
        const owg_midi_job& j = jobs[i];
        EngineDesc& e = eng[i];
        std::memset(&e, 0, sizeof(e));
        e.sample_rate = SR; e.n_samples = j.n_samples; e.block_size = CHUNK;
        e.ev_begin = (long long)events.size();
        held_offset[i] = held_total;
        uint32_t age = 0;
        long long chunk = 0;
        for (int64_t k = 0; k < j.n_ev; k++) {
            const owg_midi_event& ev = j.ev[k];
            // the event is applied in the first chunk whose start time `sample_pos as f64 / BASE_SR` is >= time_s (main.rs:1790-1795)
            while ((double)(chunk * CHUNK) / SR < ev.time_s) chunk++;
            MidiEvent d;
            d.chunk = chunk; d.kind = ev.kind; d.vinit = -1;
            d.note = ev.kind == OWG_MIDI_PEDAL ? (ev.velocity != 0 ? 1 : 0) : (ev.note < 33 ? 33 : (ev.note > 96 ? 96 : ev.note));
            if (ev.kind == OWG_MIDI_NOTE_ON) {
                age += 1;
                owg_voice_job vj;
                std::memset(&vj, 0, sizeof(vj));
                vj.midi = (uint8_t)d.note; vj.mlp_enabled = 1; vj.attack_noise = 1;
                vj.noise_seed = (uint32_t)d.note * 2654435761u + age;
                vj.velocity = (double)ev.velocity / 127.0;
                vj.sample_rate = SR; vj.duration_s = 0.0; vj.ds_override = NAN;
                d.vinit = (long long)vinits.size();
                vinits.emplace_back();
                owg::make_voice_init(vj, &vinits.back());
            } else if (ev.kind == OWG_MIDI_NOTE_OFF) held_total++;
            events.push_back(d);
        }
        e.ev_end = (long long)events.size();
        std::memset(&params[i], 0, sizeof(owg_bench_job));
        params[i].v.sample_rate = SR; params[i].r_ldr = 1000000.0; params[i].tremolo_depth = 0.0; params[i].volume = j.volume;
        params[i].speaker_character = j.speaker_character; params[i].no_poweramp = j.no_poweramp;
        n_each[i] = (unsigned long long)j.n_samples;
    }
    if (events.empty()) events.push_back(MidiEvent{0, OWG_MIDI_PEDAL, 0, -1});
    if (vinits.empty()) vinits.emplace_back();
    std::vector<DamperRow> dampers(128);
    owg::make_damper_rows(SR, dampers.data());
    DevBuf<EngineDesc> d_eng; DevBuf<MidiEvent> d_events; DevBuf<OwgVoiceInit> d_vinits; DevBuf<DamperRow> d_dampers; DevBuf<long long> d_heldoff;
    DevBuf<VoiceRT> d_pool; DevBuf<EngineState> d_states; DevBuf<uint8_t> d_held; DevBuf<int32_t> d_heldcount; DevBuf<double> d_rows;
    int rc = d_eng.upload(eng, s);
    if (!rc) rc = d_events.upload(events, s);
    if (!rc) rc = d_vinits.upload(vinits, s);
    if (!rc) rc = d_dampers.upload(dampers, s);
    if (!rc) rc = d_heldoff.upload(held_offset, s);
    if (!rc) rc = d_pool.alloc((size_t)n * 128);
    if (!rc) rc = d_states.alloc((size_t)n);
    if (!rc) rc = d_held.alloc((size_t)std::max<long long>(held_total, 1));
    if (!rc) rc = d_heldcount.alloc((size_t)n);
    if (!rc) rc = d_rows.alloc((size_t)n * (size_t)max_samples);
    if (rc) return rc;
    CK(cudaMemsetAsync(d_pool.p, 0, (size_t)n * 128 * sizeof(VoiceRT), s));
    CK(cudaMemsetAsync(d_heldcount.p, 0, (size_t)n * sizeof(int32_t), s));
    CK(cudaMemsetAsync(d_rows.p, 0, (size_t)n * (size_t)max_samples * sizeof(double), s));  // ragged batch: rows end in silence
    const unsigned eb = (unsigned)((n + 63) / 64);
    engine_init_kernel<<<eb, 64, 0, s>>>(d_eng.p, (int)n, nullptr, nullptr, d_states.p, nullptr, nullptr, nullptr);
    CK(cudaGetLastError());
    const long long n_rounds = (max_samples + CHUNK - 1) / CHUNK;
    const double silent_thr = owg::silent_threshold();
    for (long long r = 0; r < n_rounds; r++) {
        midi_events_kernel<<<eb, 64, 0, s>>>(d_eng.p, (int)n, r, d_events.p, d_vinits.p, d_dampers.p, d_pool.p, d_states.p, d_held.p, d_heldcount.p,
                                             d_heldoff.p, silent_thr);
        engine_voice_mix_kernel<<<(unsigned)n, OWG_ENGINE_ITEMS, 0, s>>>(d_eng.p, r, d_pool.p, d_states.p, d_rows.p, max_samples, 0);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    {   // counters: [0] note-ons, [2] voices dropped as silent, [3] peak polyphony
        std::vector<EngineState> hs((size_t)n);
        CK(cudaMemcpy(hs.data(), d_states.p, (size_t)n * sizeof(EngineState), cudaMemcpyDeviceToHost));
        owg_diag& d = g_last_diag;
        std::memset(&d, 0, sizeof(d));
        for (auto& st : hs) { d.nr_iter_hist[0] += st.d_note_ons; d.nr_iter_hist[2] += st.d_freed; d.nr_iter_hist[3] = std::max<uint64_t>(d.nr_iter_hist[3], st.d_max_active); }
        d.kernels_launched = (uint64_t)(2 * n_rounds + 1);
    }
    // voices summed: the rest is chain B over the rows (static preamp, `set_ldr_resistance(1e6); reset()` order)
    owg_opts o;
    if (opts) o = *opts; else owg_default_opts(&o);
    o.stream = nullptr;
    o.collect_diag = 0;
    return chain_batch_impl(d_rows.p, OWG_OUT_DEVICE, max_samples, n, n_each.data(), max_samples, params.data(), OWG_INIT_SET_THEN_RESET, out,
                            stride, &o);
}

int owg_host_voice_init(const owg_voice_job* job, double* o) {
    if (!job || !o) return fail(OWG_E_BAD_ARG, "owg_host_voice_init: null");
    OwgVoiceInit v;
    owg::make_voice_init(*job, &v);
    int k = 0;
    for (int m = 0; m < 7; m++) { o[k++] = v.cos_inc[m]; o[k++] = v.sin_inc[m]; o[k++] = v.phase_inc[m]; o[k++] = v.amplitude[m]; o[k++] = v.decay_mult[m]; o[k++] = v.jitter_drift[m]; }
    o[k++] = v.jitter_revert; o[k++] = v.jitter_diffusion; o[k++] = v.onset_ramp_inc; o[k++] = v.onset_shape_exp;
    o[k++] = v.pickup_beta; o[k++] = v.pickup_ds; o[k++] = v.post_pickup_gain; o[k++] = v.noise_amp; o[k++] = v.noise_decay;
    o[k++] = v.bq_b0; o[k++] = v.bq_b1; o[k++] = v.bq_b2; o[k++] = v.bq_a1; o[k++] = v.bq_a2;
    o[k++] = (double)v.onset_ramp_samples; o[k++] = (double)v.n_samples; o[k++] = (double)v.jitter_state; o[k++] = (double)v.noise_rng;
    o[k++] = (double)v.noise_remaining;
    return OWG_OK;
}

int owg_host_legacy_group(double preamp_sr, double r_static, double* o) {
    if (!o || !(preamp_sr > 0.0) || !std::isfinite(preamp_sr)) return fail(OWG_E_BAD_ARG, "owg_host_legacy_group: bad argument");
    static_assert(OWG_LG_STRIDE == OWG_LEGACY_GROUP_DOUBLES, "record layout");
    owg::make_legacy_group(preamp_sr, r_static, o);
    return OWG_OK;
}

int owg_host_chain_init(const owg_bench_job* job, double* o) {
    if (!job || !o) return fail(OWG_E_BAD_ARG, "owg_host_chain_init: null");
    OwgChainInit c;
    owg::make_chain_init(*job, 0, &c);
    const double v[18] = {c.volume, c.spk_a2, c.spk_a3, c.spk_norm, c.spk_thermal_coeff, c.spk_thermal_alpha, c.hpf_b0, c.hpf_b1, c.hpf_b2,
                          c.hpf_a1, c.hpf_a2, c.lpf_b0, c.lpf_b1, c.lpf_b2, c.lpf_a1, c.lpf_a2, (double)c.spk_tanh, (double)c.oversample};
    for (int i = 0; i < 18; i++) o[i] = v[i];
    return OWG_OK;
}

int owg_last_diag(owg_diag* out) {
    if (!out) return fail(OWG_E_BAD_ARG, "null diag");
    *out = g_last_diag;
    return OWG_OK;
}

int owg_selftest_division(int64_t n_per_thread, uint64_t seed, uint64_t* mismatches, uint64_t* tested) {
    DeviceRestore restore_device_;
    if (!mismatches || n_per_thread <= 0) return fail(OWG_E_BAD_ARG, "owg_selftest_division: bad argument");
    if (usable_devices() <= 0) return fail(OWG_E_NO_DEVICE, "no usable CUDA device");
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, sizeof(unsigned long long)));
    CK(cudaMemset(d, 0, sizeof(unsigned long long)));
    const int blocks = 1184, threads = 256;
    division_selftest_kernel<<<blocks, threads>>>((unsigned long long)seed, (int)n_per_thread, d);
    CK(cudaGetLastError());
    unsigned long long h = 0;
    CK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = h;
    if (tested) *tested = (uint64_t)blocks * threads * (uint64_t)n_per_thread;
    return OWG_OK;
}

int owg_debug_counters(uint64_t* out, int32_t n, int32_t reset) {
    if (!out || n < 0) return fail(OWG_E_BAD_ARG, "owg_debug_counters: bad argument");
    if (usable_devices() <= 0) return fail(OWG_E_NO_DEVICE, "no usable CUDA device");
    unsigned long long h[18] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaMemcpyFromSymbol(h, g_tile_prof, 8 * sizeof(unsigned long long)));
    CK(cudaMemcpyFromSymbol(&h[8], g_tile_rare, sizeof(unsigned long long)));
    CK(cudaMemcpyFromSymbol(&h[9], g_tile_sec, 8 * sizeof(unsigned long long)));
    CK(cudaMemcpyFromSymbol(&h[17], g_trm_generic, sizeof(unsigned long long)));
    for (int i = 0; i < n && i < 18; i++) out[i] = h[i];
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        CK(cudaMemcpyToSymbol(g_tile_prof, z, sizeof(z)));
        CK(cudaMemcpyToSymbol(g_tile_rare, z, sizeof(unsigned long long)));
        CK(cudaMemcpyToSymbol(g_tile_sec, z, sizeof(z)));
        CK(cudaMemcpyToSymbol(g_trm_generic, z, sizeof(unsigned long long)));
    }
    return OWG_OK;
}

static int alias_analyze_device(const void* d_rows, int32_t row_dtype, int64_t stride, int64_t n_rows, int64_t n_samples, double sample_rate,
                                double analyze_seconds, const double* nominal_f0, double* results, cudaStream_t s) {
    const double an = sample_rate * analyze_seconds;
    const int64_t analyze_n = !(an == an) || an <= 0.0 ? 0 : (int64_t)an;
    if (analyze_n <= 0 || n_samples < analyze_n) return fail(OWG_E_BAD_ARG, "owg_alias_analyze: the streams are shorter than the analysis window");
    DevBuf<double> d_f0, d_out;
    std::vector<double> f0v(nominal_f0, nominal_f0 + n_rows);
    int rc = d_f0.upload(f0v, s);
    if (!rc) rc = d_out.alloc((size_t)n_rows * OWG_ALIAS_OUT);
    if (rc) return rc;
    AliasBq hp, lp;
    double c5[5];
    owg::rbj_coefficients(1, 5000.0, 0.70710678118654752440, sample_rate, c5);
    hp = AliasBq{c5[0], c5[1], c5[2], c5[3], c5[4]};
    owg::rbj_coefficients(0, 18000.0, 0.70710678118654752440, sample_rate, c5);
    lp = AliasBq{c5[0], c5[1], c5[2], c5[3], c5[4]};
    if (row_dtype == OWG_ROWS_F32)
        alias_analyze_kernel<float><<<(unsigned)n_rows, 128, 0, s>>>((const float*)d_rows, stride, n_samples, analyze_n, sample_rate, d_f0.p, hp, lp, d_out.p);
    else
        alias_analyze_kernel<double><<<(unsigned)n_rows, 128, 0, s>>>((const double*)d_rows, stride, n_samples, analyze_n, sample_rate, d_f0.p, hp, lp, d_out.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(results, d_out.p, (size_t)n_rows * OWG_ALIAS_OUT * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return OWG_OK;
}

int owg_alias_analyze(const void* rows, int32_t row_dtype, int64_t stride, int64_t n_rows, int64_t n_samples, double sample_rate,
                      double analyze_seconds, const double* nominal_f0, double* results, const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (!rows || !nominal_f0 || !results || n_rows < 0 || n_samples < 0 || stride < n_samples || (row_dtype != OWG_ROWS_F64 && row_dtype != OWG_ROWS_F32) ||
        !(sample_rate > 0.0) || !std::isfinite(sample_rate) || !(analyze_seconds > 0.0))
        return fail(OWG_E_BAD_ARG, "owg_alias_analyze: bad argument");
    if (usable_devices() <= 0) return fail(OWG_E_NO_DEVICE, "no usable CUDA device");
    if (n_rows == 0) return OWG_OK;
    owg_opts o;
    if (opts) o = *opts; else owg_default_opts(&o);
    if (o.device >= 0) CK(cudaSetDevice(o.device));
    cudaStream_t s = (cudaStream_t)o.stream;
    const size_t esz = row_dtype == OWG_ROWS_F32 ? 4 : 8;
    if (o.out_location == OWG_OUT_DEVICE) return alias_analyze_device(rows, row_dtype, stride, n_rows, n_samples, sample_rate, analyze_seconds, nominal_f0, results, s);
    DevBuf<unsigned char> d_rows;
    if (int rc = d_rows.alloc((size_t)n_rows * (size_t)stride * esz)) return rc;
    CK(cudaMemcpyAsync(d_rows.p, rows, (size_t)n_rows * (size_t)stride * esz, cudaMemcpyHostToDevice, s));
    return alias_analyze_device(d_rows.p, row_dtype, stride, n_rows, n_samples, sample_rate, analyze_seconds, nominal_f0, results, s);
}

int owg_render_engines_alias(const owg_engine_job* jobs, int64_t n, double analyze_seconds, const double* nominal_f0, double* results,
                             const owg_opts* opts) {
    DeviceRestore restore_device_;
    if (!jobs || !nominal_f0 || !results || n < 0) return fail(OWG_E_BAD_ARG, "owg_render_engines_alias: bad argument");
    if (usable_devices() <= 0) return fail(OWG_E_NO_DEVICE, "no usable CUDA device");
    if (n == 0) return OWG_OK;
    for (int64_t i = 1; i < n; i++)
        if (jobs[i].sample_rate != jobs[0].sample_rate || jobs[i].duration_s != jobs[0].duration_s)
            return fail(OWG_E_BAD_ARG, "owg_render_engines_alias: all streams must share sample_rate and duration");
    const double ns = jobs[0].sample_rate * jobs[0].duration_s;
    const int64_t n_samples = !(ns == ns) || ns <= 0.0 ? 0 : (int64_t)ns;
    owg_opts o;
    if (opts) o = *opts; else owg_default_opts(&o);
    if (o.device >= 0) CK(cudaSetDevice(o.device));
    DevBuf<float> d_sig;
    if (int rc = d_sig.alloc((size_t)n * (size_t)n_samples)) return rc;
    o.out_location = OWG_OUT_DEVICE;
    if (int rc = owg_render_engines(jobs, n, d_sig.p, n_samples, &o)) return rc;
    return alias_analyze_device(d_sig.p, OWG_ROWS_F32, n_samples, n, n_samples, jobs[0].sample_rate, analyze_seconds, nominal_f0, results,
                                (cudaStream_t)o.stream);
}

int owg_fp64_peak(int32_t device, int32_t fma_mode, float ms_target, double* tera_instr_per_s) {
    DeviceRestore restore_device_;
    if (!tera_instr_per_s) return fail(OWG_E_BAD_ARG, "null result");
    if (usable_devices() <= 0) return fail(OWG_E_NO_DEVICE, "no usable CUDA device");
    if (device >= 0) CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaGetDeviceProperties(&prop, dev));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    double* sink = nullptr;
    CK(cudaMalloc(&sink, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    if (fma_mode >= 10 && fma_mode <= 13) {  // dependent-chain latency: result = ns per dependent op (DADD, DMUL, DFMA, DDIV)
        const int it = 1 << 16;
        double best_ns = 1e30;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            fp64_latency_kernel<<<1, 32>>>(sink, it, 1.0000001, 1e-9, fma_mode - 10);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best_ns = std::min(best_ns, (double)ms * 1e6 / (4.0 * it));
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
        *tera_instr_per_s = best_ns;
        return OWG_OK;
    }
    if (fma_mode >= 100) {  // partial-warp throughput: fma_mode-100 active lanes per warp; result = 1e12 warp-instr-lanes/s (active lanes only)
        const int lanes = fma_mode - 100;
        const int it = 1 << 16;
        double best_rate = 0.0;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            fp64_partial_warp_kernel<<<blocks, threads>>>(sink, it, 1.0000001, 1e-9, lanes);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best_rate = std::max(best_rate, (double)(threads / 32) * blocks * (double)it * 8.0 / (ms * 1e-3) / 1e12);  // warp-instructions/s
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
        *tera_instr_per_s = best_rate;
        return OWG_OK;
    }
    int iters = 1 << 14;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        if (fma_mode) fp64_peak_kernel<true><<<blocks, threads>>>(sink, iters, 1.0000001, 1e-9);
        else fp64_peak_kernel<false><<<blocks, threads>>>(sink, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double instr = (double)threads * blocks * (double)iters * 8.0;
        const double rate = instr / (ms * 1e-3) / 1e12;
        if (rep > 0 && rate > best) best = rate;
        if (ms < ms_target && iters < (1 << 24)) iters *= 2;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    *tera_instr_per_s = best;
    return OWG_OK;
}

}  // extern "C"
