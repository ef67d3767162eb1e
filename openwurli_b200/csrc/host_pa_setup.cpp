// Host-side construction of the melange power amplifier's model for one sample rate: what CircuitState::default() holds at 88.2 kHz and
// what CircuitState::set_sample_rate / rebuild_matrices (gen_power_amp.rs:12440-12610) make of it at any other rate, plus the rail-sag
// smoothing coefficients (power_amp.rs:103-108).  Plain IEEE f64 in the reference's operation order (-ffp-contract=off), so the tables are
// bit-identical to the reference's; checked against the oracle's transliterated solver in tests/test_power_amp_melange.py.
#include <cmath>
#include <cstring>

#include "host_setup.h"
#include "owg_pa_core.h"

namespace owg {
namespace {
#define OWC_TABLE(name) const double name
#define OWC_SCALAR(name) const double name
#include "ow_consts_pa.inc"
#undef OWC_TABLE
#undef OWC_SCALAR

// invert_n (gen_power_amp.rs:8680-8760): LU with partial pivoting, one unit vector per column, forward substitution from the pivot row on
bool invert20(const double (*a)[PA_N], double (*result)[PA_N]) {
    double lu[PA_N][PA_N];
    int perm[PA_N];
    std::memcpy(lu, a, sizeof(lu));
    for (int i = 0; i < PA_N; i++) perm[i] = i;
    for (int k = 0; k < PA_N; k++) {
        int max_row = k;
        double max_val = std::fabs(lu[k][k]);
        for (int i = k + 1; i < PA_N; i++) {
            const double v = std::fabs(lu[i][k]);
            if (v > max_val) { max_val = v; max_row = i; }
        }
        if (max_val < 1e-30) return false;
        if (max_row != k) {
            for (int j = 0; j < PA_N; j++) { const double tmp = lu[k][j]; lu[k][j] = lu[max_row][j]; lu[max_row][j] = tmp; }
            const int tp = perm[k]; perm[k] = perm[max_row]; perm[max_row] = tp;
        }
        const double pivot = lu[k][k];
        for (int i = k + 1; i < PA_N; i++) {
            const double mul = lu[i][k] / pivot;
            lu[i][k] = mul;
            for (int j = k + 1; j < PA_N; j++) lu[i][j] -= mul * lu[k][j];
        }
    }
    for (int col = 0; col < PA_N; col++) {
        double b[PA_N];
        for (int i = 0; i < PA_N; i++) b[i] = 0.0;
        int start = PA_N;
        for (int i = 0; i < PA_N; i++)
            if (perm[i] == col) { b[i] = 1.0; start = i; break; }
        for (int i = start + 1; i < PA_N; i++) {
            double sum = b[i];
            for (int j = start; j < i; j++) sum -= lu[i][j] * b[j];
            b[i] = sum;
        }
        for (int i = PA_N - 1; i >= 0; i--) {
            double sum = b[i];
            for (int j = i + 1; j < PA_N; j++) sum -= lu[i][j] * b[j];
            const double pivot = lu[i][i];
            if (std::fabs(pivot) < 1e-30) return false;
            b[i] = sum / pivot;
        }
        for (int i = 0; i < PA_N; i++) result[i][col] = b[i];
    }
    return true;
}

// S -> K = N_V (S N_I) and S N_I, with the reference's loop nests (:12490-12515)
void derive_k_sni(const double (*s)[PA_N], double (*k)[PA_M], double (*s_ni)[PA_M]) {
    for (int i = 0; i < PA_M; i++)
        for (int j = 0; j < PA_M; j++) {
            double sum = 0.0;
            for (int a = 0; a < PA_N; a++) {
                double s_ni_aj = 0.0;
                for (int b = 0; b < PA_N; b++) s_ni_aj += s[a][b] * PA_N_I[b][j];
                sum += PA_N_V[i][a] * s_ni_aj;
            }
            k[i][j] = sum;
        }
    for (int i = 0; i < PA_N; i++)
        for (int j = 0; j < PA_M; j++) {
            double sum = 0.0;
            for (int a = 0; a < PA_N; a++) sum += s[i][a] * PA_N_I[a][j];
            s_ni[i][j] = sum;
        }
}
}  // namespace

void pa_build_model(double sample_rate, PaModel* m) {
    std::memset(m, 0, sizeof(*m));
    // CircuitState::default(): the baked 88.2 kHz tables
    std::memcpy(m->s, PA_S_DEFAULT, sizeof(m->s));
    std::memcpy(m->k, PA_K_DEFAULT, sizeof(m->k));
    std::memcpy(m->s_ni, PA_S_NI_DEFAULT, sizeof(m->s_ni));
    std::memcpy(m->s_be, PA_S_BE_DEFAULT, sizeof(m->s_be));
    std::memcpy(m->k_be, PA_K_BE_DEFAULT, sizeof(m->k_be));
    std::memcpy(m->s_ni_be, PA_S_NI_BE_DEFAULT, sizeof(m->s_ni_be));
    std::memcpy(m->a_neg_be, PA_A_NEG_BE_DEFAULT, sizeof(m->a_neg_be));
    double a_neg[PA_N][PA_N];
    std::memcpy(a_neg, PA_A_NEG_DEFAULT, sizeof(a_neg));
    m->dc_block_r = PA_DC_BLOCK_R;
    m->rerated = 0;
    const bool valid = sample_rate > 0.0 && std::isfinite(sample_rate);
    if (valid && !(std::fabs(sample_rate - PA_SAMPLE_RATE) < 0.5)) {
        // set_sample_rate -> rebuild_matrices(internal_rate = sample_rate * 1.0): alpha = alpha_be = internal_rate (a backward-Euler
        // companion model in both slots), a = G + alpha C, a_neg = alpha C with the two voltage-source rows cleared
        const double internal_rate = sample_rate * 1.0;
        const double alpha = internal_rate, alpha_be = internal_rate;
        double a[PA_N][PA_N], a_be[PA_N][PA_N];
        for (int i = 0; i < PA_N; i++)
            for (int j = 0; j < PA_N; j++) {
                a[i][j] = PA_G[i][j] + alpha * PA_C[i][j];
                a_neg[i][j] = alpha * PA_C[i][j];
                a_be[i][j] = PA_G[i][j] + alpha_be * PA_C[i][j];
                m->a_neg_be[i][j] = alpha_be * PA_C[i][j];
            }
        for (int i = 18; i < 20; i++)
            for (int j = 0; j < PA_N; j++) { a_neg[i][j] = 0.0; m->a_neg_be[i][j] = 0.0; }
        double inv[PA_N][PA_N];
        if (invert20(a, inv)) {
            std::memcpy(m->s, inv, sizeof(inv));
            derive_k_sni(m->s, m->k, m->s_ni);
        }
        if (invert20(a_be, inv)) {
            std::memcpy(m->s_be, inv, sizeof(inv));
            derive_k_sni(m->s_be, m->k_be, m->s_ni_be);
        }
        m->dc_block_r = 1.0 - ((2.0 * 3.14159265358979323846264338327950288) * 5.0) / internal_rate;
        m->rerated = 1;
    }
    // the structural non-zeros of a_neg (those of C, rows 0..17): the terms process_sample spells out one by one (:8842-8895)
    for (int i = 0; i < PA_N; i++) {
        int c = 0;
        if (i < 18)
            for (int j = 0; j < PA_N; j++)
                if (PA_C[i][j] != 0.0) { m->a_col[i][c] = j; m->a_val[i][c] = a_neg[i][j]; c++; }
        m->a_cnt[i] = c;
    }
    for (int i = 0; i < PA_M; i++) {
        int c = 0;
        for (int j = 0; j < PA_N; j++)
            if (PA_N_V[i][j] != 0.0 && c < 2) { m->nv_col[i][c] = j; m->nv_val[i][c] = PA_N_V[i][j]; c++; }
    }
    std::memcpy(m->n_i, PA_N_I, sizeof(m->n_i));
    std::memcpy(m->n_v, PA_N_V, sizeof(m->n_v));
    std::memcpy(m->rhs_const, PA_RHS_CONST, sizeof(m->rhs_const));
    std::memcpy(m->rhs_const_be, PA_RHS_CONST_BE, sizeof(m->rhs_const_be));
    std::memcpy(m->dc_op, PA_DC_OP, sizeof(m->dc_op));
    std::memcpy(m->dc_nl_i, PA_DC_NL_I, sizeof(m->dc_nl_i));
    for (int d = 0; d < PA_NDEV; d++) pa_dev_derive(PA_DEV[d], m->dev[d]);
    m->input_conductance = 1.0 / PA_INPUT_RESISTANCE;
    m->nan_out = PA_NAN_OUT;
    m->out_node = 8.0;
    // RailDynamics::set_sample_rate (power_amp.rs:103-108)
    const double dt = 1.0 / sample_rate;
    m->alpha_attack = 1.0 - std::exp(-dt / 0.008);
    m->alpha_release = 1.0 - std::exp(-dt / 0.015);
    m->alpha_i_avg = 1.0 - std::exp(-dt / 0.030);
    m->sample_rate = sample_rate;
}

}  // namespace owg
