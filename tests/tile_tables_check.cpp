// Host check of openwurli_b200/csrc/owg_tile_tables.h: the lane-tiled evaluation of build_rhs (4 lanes x 14 padded term slots)
// must reproduce the straightforward rows (gen_preamp.rs:3041-3095, as restated in owg_device.cuh / oracle) bit for bit,
// including the -0.0 x 1.0 padding and the per-row summation order.  Built and run by tests/test_host_logic.py.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include "../openwurli_b200/csrc/owg_tile_tables.h"

static const OwgRhsTerm ROWS[12][OWG_TILE_ROW_TERMS] = OWG_RHS_ROWS_INIT;

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static double rnd() {
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    const double u = (double)(rng >> 11) / 9007199254740992.0;
    const double mag = std::exp((u - 0.5) * 40.0);
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    return (rng & 1) ? mag : -mag;
}

int main() {
    long bad = 0;
    for (int trial = 0; trial < 20000; trial++) {
        double an[38], vp[12], il[3], ni[7], rhs11 = 15.0, input = rnd(), xin_prev = rnd();
        for (double& x : an) x = rnd();
        for (double& x : vp) x = rnd();
        for (double& x : il) x = rnd();
        for (double& x : ni) x = (rng & 2) ? 1.0 : -1.0;
        if (trial % 7 == 0) { vp[trial % 12] = 0.0; vp[(trial + 5) % 12] = -0.0; il[trial % 3] = -0.0; }
        const double an66 = rnd();
        // reference rows (same text as owg_device.cuh dk_step; entry 24 = [6][6] replaced by an66)
        double ref[12];
        ref[0] = an[0] * vp[0] + an[1] * vp[1];
        ref[1] = an[2] * vp[0] + an[3] * vp[1] + an[4] * vp[2];
        ref[2] = an[5] * vp[1] + an[6] * vp[2] + an[7] * vp[3] + an[8] * vp[4] + an[9] * vp[5];
        ref[3] = an[10] * vp[2] + an[11] * vp[3] + an[12] * vp[4] + an[13] * vp[7] + an[14] * vp[11];
        ref[4] = an[15] * vp[2] + an[16] * vp[3] + an[17] * vp[4] + an[18] * vp[7] + an[19] * vp[8];
        ref[5] = an[20] * vp[2] + an[21] * vp[5] + an[22] * vp[6];
        ref[6] = an[23] * vp[5] + an66 * vp[6] + an[25] * vp[10];
        ref[7] = an[26] * vp[3] + an[27] * vp[4] + an[28] * vp[7] + an[29] * vp[10];
        ref[8] = an[30] * vp[4] + an[31] * vp[8] + an[32] * vp[9];
        ref[9] = an[33] * vp[8] + an[34] * vp[9];
        ref[10] = an[35] * vp[6] + an[36] * vp[7] + an[37] * vp[10];
        ref[11] = rhs11;
        ref[2] += ni[0] * il[0];
        ref[2] += ni[1] * il[1];
        ref[4] += ni[2] * il[1];
        ref[4] += ni[3] * il[2];
        ref[5] += ni[4] * il[1];
        ref[7] += ni[5] * il[2];
        ref[8] += ni[6] * il[2];
        ref[0] += (input + xin_prev) / 1.0;
        // tiled evaluation: position-major gather buffer, exactly as the kernel fills and reads it
        static const int TROWS[4][3] = OWG_TILE_ROWS_INIT;
        static const int LOC[16] = OWG_TILE_LOC_INIT;
        double xs[16];
        for (int q = 0; q < 4; q++) for (int p = 0; p < 3; p++) xs[p * 4 + q] = vp[TROWS[q][p]];
        for (int i = 0; i < 3; i++) xs[12 + i] = il[i];
        xs[15] = 1.0;
        auto coef_of = [&](int c) -> double {
            if (c < 38) return an[c];
            if (c >= OWG_TC_NI02 && c <= OWG_TC_NI28) return ni[c - OWG_TC_NI02];
            if (c == OWG_TC_RHS11) return rhs11;
            return -0.0;
        };
        double got[12];
        for (int q = 0; q < 4; q++) {
            double r[3];
            for (int s = 0; s < OWG_TILE_SLOTS; s++) {
                const int pos = OWG_TILE_SLOT_POS(s), k = OWG_TILE_SLOT_K(s);
                const int row = TROWS[q][pos];
                const OwgRhsTerm tm = ROWS[row][k];
                double c = coef_of(tm.c);
                if (q == OWG_TILE_AN66_LANE && s == OWG_TILE_AN66_SLOT) {  // the kernel's per-sample a_neg[6][6]
                    if (tm.c != 24) { std::printf("an66 slot does not hold a_neg entry 24\n"); bad++; }
                    c = an66;
                } else if (tm.c == 24) { std::printf("a_neg entry 24 outside the an66 slot\n"); bad++; }
                const double term = c * xs[LOC[tm.x]];
                if (k == 0) r[pos] = term; else r[pos] += term;
            }
            r[2] += q == 0 ? (input + xin_prev) / 1.0 : -0.0;  // row 0 = lane 0, position 2
            for (int p = 0; p < 3; p++) got[TROWS[q][p]] = r[p];
        }
        for (int i = 0; i < 12; i++) {
            uint64_t a, b;
            std::memcpy(&a, &ref[i], 8); std::memcpy(&b, &got[i], 8);
            if (a != b) { if (bad < 5) std::printf("row %d trial %d: %.17g vs %.17g\n", i, trial, ref[i], got[i]); bad++; }
        }
    }
    // structural checks: every row fits the slot count of its position; rows and LOC are inverse maps; every a_neg entry used once
    static const int TROWS2[4][3] = OWG_TILE_ROWS_INIT;
    static const int LOC2[16] = OWG_TILE_LOC_INIT;
    static const int CAP[3] = {7, 4, 3};
    int used[38] = {0}, seen[12] = {0};
    for (int q = 0; q < 4; q++)
        for (int p = 0; p < 3; p++) {
            const int r = TROWS2[q][p];
            seen[r]++;
            if (LOC2[r] != p * 4 + q) { std::printf("LOC[%d] is not the inverse of ROWS\n", r); bad++; }
            for (int k = CAP[p]; k < OWG_TILE_ROW_TERMS; k++)
                if (ROWS[r][k].c != OWG_TC_PAD) { std::printf("row %d has more than %d terms\n", r, CAP[p]); bad++; }
        }
    for (int r = 0; r < 12; r++) if (seen[r] != 1) { std::printf("row %d owned %d times\n", r, seen[r]); bad++; }
    for (int i = 12; i < 16; i++) if (LOC2[i] != i) { std::printf("LOC[%d] must be the identity\n", i); bad++; }
    if (TROWS2[0][2] != 0 || TROWS2[1][0] != 2 || TROWS2[1][1] != 5 || TROWS2[2][0] != 4 || TROWS2[2][1] != 8 || TROWS2[2][2] != 10 || TROWS2[3][2] != 11) {
        std::printf("the kernel's fixed roles (input row, p rows, output row, V-source row) moved\n"); bad++;
    }
    for (int r = 0; r < 12; r++)
        for (int k = 0; k < OWG_TILE_ROW_TERMS; k++)
            if (ROWS[r][k].c < 38) used[ROWS[r][k].c]++;
    for (int e = 0; e < 38; e++) if (used[e] != 1) { std::printf("a_neg entry %d used %d times\n", e, used[e]); bad++; }
    std::printf("%s\n", bad ? "FAIL" : "OK");
    return bad ? 1 : 0;
}
