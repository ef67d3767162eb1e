// TEST HARNESS: runs the PRODUCT's tile-collective power-amplifier code (openwurli_b200/csrc/owg_pa_core.h, the source the CUDA kernel is
// built from) on the CPU, one coroutine per lane in lock-step, so that the lane-tiled algorithm can be compared with the oracle without a
// GPU.  Collectives: a lane that reaches a collective publishes its operand and yields round-robin until all 16 lanes have arrived.
// Not part of the product (the product path is the CUDA kernel in owg_poweramp.cuh); built by tests/test_power_amp_melange.py.
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../openwurli_b200/csrc/host_setup.h"
#include "../openwurli_b200/csrc/owg_pa_core.h"

namespace {
constexpr int L = 16;      // lanes per tile
constexpr int LMAX = 32;   // one emulated warp: one or two tiles
struct Emu {
    ucontext_t main_ctx, ctx[LMAX];
    std::vector<char> stacks[LMAX];
    bool done[LMAX];
    int nl = L;            // coroutines in this run (16: one tile, 32: two tiles in lock-step, as on the GPU)
    int cur = 0, arrived = 0;
    unsigned long gen = 0;
    double slot_d[LMAX];
    int slot_i[LMAX];
    void yield_next() {
        const int from = cur;
        int nxt = (cur + 1) % nl;
        while (done[nxt] && nxt != from) nxt = (nxt + 1) % nl;
        if (nxt == from) return;
        cur = nxt;
        swapcontext(&ctx[from], &ctx[nxt]);
    }
    // Every collective is a barrier over ALL lanes of the emulated warp, like a full-mask __shfl_sync / __syncwarp on the device: a tile
    // that skipped a collective its neighbour executes would pair up different collectives and fail the comparison with the oracle.
    void barrier() {
        const unsigned long g = gen;
        if (++arrived == nl) { arrived = 0; gen++; return; }
        while (gen == g) yield_next();
    }
};
Emu* g_emu;
struct EmuTile {
    int lane, base;  // lane within the tile, first lane of the tile in the warp
    void sync() const { g_emu->barrier(); }
    double shfl(double x, int src) const { g_emu->slot_d[base + lane] = x; g_emu->barrier(); const double r = g_emu->slot_d[base + src]; g_emu->barrier(); return r; }
    double shfl_xor(double x, int m) const { return shfl(x, lane ^ m); }
    int shfl_i(int x, int src) const { g_emu->slot_i[base + lane] = x; g_emu->barrier(); const int r = g_emu->slot_i[base + src]; g_emu->barrier(); return r; }
    int shfl_xor_i(int x, int m) const { return shfl_i(x, lane ^ m); }
    bool any(bool p) const {  // over the tile
        g_emu->slot_i[base + lane] = p ? 1 : 0; g_emu->barrier();
        int r = 0; for (int i = 0; i < L; i++) r |= g_emu->slot_i[base + i];
        g_emu->barrier(); return r != 0;
    }
    bool warp_any(bool p) const {  // over both tiles
        g_emu->slot_i[base + lane] = p ? 1 : 0; g_emu->barrier();
        int r = 0; for (int i = 0; i < g_emu->nl; i++) r |= g_emu->slot_i[i];
        g_emu->barrier(); return r != 0;
    }
    uint32_t max_u32(uint32_t x) const {
        g_emu->slot_i[base + lane] = (int)x; g_emu->barrier();
        uint32_t r = 0; for (int i = 0; i < L; i++) { const uint32_t v = (uint32_t)g_emu->slot_i[base + i]; if (v > r) r = v; }
        g_emu->barrier(); return r;
    }
    int first_lane(bool p) const {
        g_emu->slot_i[base + lane] = p ? 1 : 0; g_emu->barrier();
        int r = L - 1; for (int i = L - 1; i >= 0; i--) if (g_emu->slot_i[base + i]) r = i;
        g_emu->barrier(); return r;
    }
};
struct Job {
    const PaModel* m; const PaShared* sh; PaScratch* sc; const PaSettled* settled; const double* x; double* y; int64_t n; double pre_gain;
    bool rail_sag; PaSettled* settle; double* rails; uint32_t* counters; int64_t n_steps;
};
Job g_jobs[2];
void lane_main(int idx) {
    const Job& j = g_jobs[idx / L];
    EmuTile t{idx % L, (idx / L) * L};
    PaNoPost post;
    pa_tile_render(t, *j.m, *j.sh, *j.sc, j.settled, j.x, j.y, j.n, j.n_steps, true, j.pre_gain, j.rail_sag, false, j.settle, j.rails, j.counters, post);
    g_emu->done[idx] = true;
    // hand over to a lane that is still running, or back to main when this was the last one
    for (int i = 1; i <= g_emu->nl; i++) {
        const int nxt = (idx + i) % g_emu->nl;
        if (!g_emu->done[nxt]) { g_emu->cur = nxt; setcontext(&g_emu->ctx[nxt]); }
    }
    setcontext(&g_emu->main_ctx);
}
void run_tiles(const Job* jobs, int n_tiles) {
    Emu emu;
    g_emu = &emu;
    emu.nl = n_tiles * L;
    int64_t steps = 0;
    for (int k = 0; k < n_tiles; k++) steps = jobs[k].n > steps ? jobs[k].n : steps;
    for (int k = 0; k < n_tiles; k++) { g_jobs[k] = jobs[k]; g_jobs[k].n_steps = steps; }
    for (int i = 0; i < emu.nl; i++) {
        emu.done[i] = false;
        emu.stacks[i].resize(1 << 20);
        getcontext(&emu.ctx[i]);
        emu.ctx[i].uc_stack.ss_sp = emu.stacks[i].data();
        emu.ctx[i].uc_stack.ss_size = emu.stacks[i].size();
        emu.ctx[i].uc_link = nullptr;
        makecontext(&emu.ctx[i], (void (*)())lane_main, 1, i);
    }
    emu.cur = 0;
    swapcontext(&emu.main_ctx, &emu.ctx[0]);
}
void run_tile(const Job& job) { run_tiles(&job, 1); }
PaSettled g_settled;
bool g_have_settled = false;
}  // namespace

extern "C" {
// PowerAmp::new_at_sample_rate(sample_rate), set_rail_sag(rail_sag), y[i] = process(x[i] * pre_gain); rails2 = rail_voltages() at the end;
// counters4 = {divergence-guard resets, backward-Euler retries, NaN resets, last_nr_iterations}
int paemu_render(double sample_rate, int rail_sag, const double* x, int64_t n, double pre_gain, double* y, double* rails2, uint32_t* counters4) {
    static PaModel m0, m;
    static PaShared sh;
    static PaScratch sc;
    if (!g_have_settled) {
        if (const char* cache = getenv("PAEMU_SETTLED_CACHE")) {  // debugging aid: reuse a settled state computed by an earlier run
            if (FILE* f = fopen(cache, "rb")) { g_have_settled = fread(&g_settled, sizeof(g_settled), 1, f) == 1; fclose(f); }
        }
    }
    if (!g_have_settled) {
        owg::pa_build_model(88200.0, &m0);
        pa_stage_shared(m0, sh, 0, 1);
        run_tile(Job{&m0, &sh, &sc, nullptr, nullptr, nullptr, PA_SETTLE_SAMPLES, 1.0, false, &g_settled, nullptr, nullptr, 0});
        g_have_settled = true;
        if (const char* cache = getenv("PAEMU_SETTLED_CACHE")) {
            if (FILE* f = fopen(cache, "wb")) { fwrite(&g_settled, sizeof(g_settled), 1, f); fclose(f); }
        }
    }
    owg::pa_build_model(sample_rate, &m);
    pa_stage_shared(m, sh, 0, 1);
    run_tile(Job{&m, &sh, &sc, &g_settled, x, y, n, pre_gain, rail_sag != 0, nullptr, rails2, counters4, 0});
    return 0;
}
// Two rows on the two tiles of one emulated warp, in lock-step as on the GPU (rows of different length and difficulty: one tile converges or
// ends while the other keeps iterating, retries with backward Euler, or resets).  Needs the settled state (paemu_render or paemu_set_settled).
int paemu_render_pair(double sample_rate, int rail_sag, const double* x0, int64_t n0, const double* x1, int64_t n1, double* y0, double* y1,
                      double* rails4, uint32_t* counters8) {
    static PaModel m;
    static PaShared sh;
    static PaScratch sc[2];
    if (!g_have_settled) return -1;
    owg::pa_build_model(sample_rate, &m);
    pa_stage_shared(m, sh, 0, 1);
    Job jobs[2] = {Job{&m, &sh, &sc[0], &g_settled, x0, y0, n0, 1.0, rail_sag != 0, nullptr, rails4, counters8, 0},
                   Job{&m, &sh, &sc[1], &g_settled, x1, y1, n1, 1.0, rail_sag != 0, nullptr, rails4 + 2, counters8 + 4, 0}};
    run_tiles(jobs, 2);
    return 0;
}
// CircuitState::default() followed by n_extra silent samples of the raw solver (n_extra = 44100: compute_settled_state):
// out54 = v_prev[20] i_nl_prev[16] i_nl_prev_prev[16] dc_block_x_prev dc_block_y_prev
int paemu_settle(int64_t n_extra, double* out54) {
    static PaModel m0;
    static PaShared sh;
    static PaScratch sc;
    PaSettled st;
    owg::pa_build_model(88200.0, &m0);
    pa_stage_shared(m0, sh, 0, 1);
    run_tile(Job{&m0, &sh, &sc, nullptr, nullptr, nullptr, 50 + n_extra, 1.0, false, &st, nullptr, nullptr, 0});
    memcpy(out54, &st, sizeof(st));
    return 0;
}
// use this settled state (e.g. the oracle's) instead of computing it on the emulator (44 150 emulated samples take minutes)
int paemu_set_settled(const double* in54) {
    memcpy(&g_settled, in54, sizeof(g_settled));
    g_have_settled = true;
    return 0;
}
// the model's matrices for the host-setup test: out = s[400] k[256] s_ni[320] s_be[400] k_be[256] s_ni_be[320] a_neg_be[400] dc_block_r
int paemu_model(double sample_rate, double* out) {
    static PaModel m;
    owg::pa_build_model(sample_rate, &m);
    double* o = out;
    memcpy(o, m.s, sizeof(m.s)); o += 400;
    memcpy(o, m.k, sizeof(m.k)); o += 256;
    memcpy(o, m.s_ni, sizeof(m.s_ni)); o += 320;
    memcpy(o, m.s_be, sizeof(m.s_be)); o += 400;
    memcpy(o, m.k_be, sizeof(m.k_be)); o += 256;
    memcpy(o, m.s_ni_be, sizeof(m.s_ni_be)); o += 320;
    memcpy(o, m.a_neg_be, sizeof(m.a_neg_be)); o += 400;
    *o = m.dc_block_r;
    return 0;
}
}
