// Host-compiled check of dcreg_b200/csrc/k2_fast.cuh (nvcc -O2, no GPU): the warm-started Jacobi and the adjugate
// inverse with the FullPivLU invertibility decision, on random / ill-conditioned / rank-deficient 3x3 Gram blocks.
// (The MUFU-seeded reciprocals only exist on the device; on the host the same code runs with exact 1/x, 1/sqrt(x): this
// test covers the algorithms, tests/test_gpu_configs.py the device arithmetic.)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../dcreg_b200/csrc/k2_fast.cuh"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) * (1.0 / 9007199254740992.0);
}

// Gram block sum_k s_k n_k n_k^T with `rank` independent directions and a spread of scales
static void make_block(int rank, double spread, double* A) {
    memset(A, 0, 9 * sizeof(double));
    double basis[3][3];
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) basis[k][i] = urand() * 2.0 - 1.0;
    const int terms = rank == 3 ? 40 : rank;
    for (int t = 0; t < terms; ++t) {
        double n[3] = {0, 0, 0};
        if (rank == 3) {
            for (int k = 0; k < 3; ++k) { const double c = (urand() * 2.0 - 1.0) * pow(spread, -(double)k); for (int i = 0; i < 3; ++i) n[i] += c * basis[k][i]; }
        } else {
            for (int i = 0; i < 3; ++i) n[i] = basis[t][i];
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i * 3 + j] += n[i] * n[j];
    }
}

// reference: full-pivot LU decision exactly as dla::fullpiv_inverse<3> (small_la.cuh), host loops
static bool ref_invertible(const double* Ain) {
    double A[9]; memcpy(A, Ain, sizeof(A));
    double maxpivot = 0.0;
    for (int k = 0; k < 3; ++k) {
        int br = k, bc = k; double bv = -1.0;
        for (int i = k; i < 3; ++i) for (int j = k; j < 3; ++j) { const double v = fabs(A[i * 3 + j]); if (v > bv) { bv = v; br = i; bc = j; } }
        if (bv > maxpivot) maxpivot = bv;
        if (bv == 0.0) return false;
        if (br != k) for (int j = 0; j < 3; ++j) { const double t = A[k * 3 + j]; A[k * 3 + j] = A[br * 3 + j]; A[br * 3 + j] = t; }
        if (bc != k) for (int i = 0; i < 3; ++i) { const double t = A[i * 3 + k]; A[i * 3 + k] = A[i * 3 + bc]; A[i * 3 + bc] = t; }
        const double piv = A[k * 3 + k];
        for (int i = k + 1; i < 3; ++i) { const double f = A[i * 3 + k] / piv; A[i * 3 + k] = f; for (int j = k + 1; j < 3; ++j) A[i * 3 + j] -= f * A[k * 3 + j]; }
    }
    const double thr = 2.220446049250313e-16 * 3.0 * maxpivot;
    for (int k = 0; k < 3; ++k) if (fabs(A[k * 3 + k]) <= thr) return false;
    return true;
}

int main() {
    int bad = 0, n_sys = 0, n_sing = 0, decisions_differ = 0;
    double worst_eig = 0, worst_res = 0, worst_orth = 0, worst_inv = 0, worst_warm = 0;
    long sweeps_cold = 0, sweeps_warm = 0;
    for (int it = 0; it < 200000; ++it) {
        const int rank = (it % 10 == 0) ? 1 + (it / 10) % 2 : 3;
        const double spread = pow(10.0, urand() * 1.5);                  // eigenvalue ratios up to spread^4 = 1e6
        double A[9], w[3], V[9];
        make_block(rank, spread, A);
        ++n_sys;
        const int sc = k2f::jacobi_eigh3_warm(A, nullptr, w, V);
        sweeps_cold += sc;
        const double scale = fabs(w[2]) > 0 ? fabs(w[2]) : 1.0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double r = 0, o = 0;
            for (int k = 0; k < 3; ++k) { r += A[i * 3 + k] * V[k * 3 + j]; o += V[k * 3 + i] * V[k * 3 + j]; }
            worst_res = fmax(worst_res, fabs(r - w[j] * V[i * 3 + j]) / scale);
            worst_orth = fmax(worst_orth, fabs(o - (i == j ? 1.0 : 0.0)));
        }
        if (!(w[0] <= w[1] && w[1] <= w[2])) ++bad;
        // perturb the block a little (the next ICP iteration) and restart warm from V
        double B[9], w2[3], V2[9], w3[3], V3[9], P[9];
        make_block(3, spread, P);
        for (int e = 0; e < 9; ++e) B[e] = A[e] + 1e-3 * P[e] * (A[0] + A[4] + A[8]) / (P[0] + P[4] + P[8] + 1e-300);
        const int sw = k2f::jacobi_eigh3_warm(B, V, w2, V2);
        k2f::jacobi_eigh3_warm(B, nullptr, w3, V3);
        sweeps_warm += sw;
        for (int k = 0; k < 3; ++k) worst_warm = fmax(worst_warm, fabs(w2[k] - w3[k]) / fmax(fabs(w3[2]), 1e-300));
        // relative error of every eigenvalue, in units of the condition number (the warm start forms V^T S V in floating point)
        if (rank == 3) for (int k = 0; k < 3; ++k) worst_eig = fmax(worst_eig, fabs(w2[k] - w3[k]) / fmax(fabs(w3[k]), 1e-300) / (w3[2] / fmax(w3[0], 1e-300)));
        // inverse + decision
        double inv[9];
        const bool ok = k2f::spd_inverse3(A, inv), ok_ref = ref_invertible(A);
        if (ok != ok_ref) ++decisions_differ;
        if (!ok_ref) ++n_sing;
        if (ok && rank == 3) {
            const double cond = w[2] / fmax(w[0], 1e-300);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * inv[k * 3 + j];
                worst_inv = fmax(worst_inv, fabs(s - (i == j ? 1.0 : 0.0)) / cond);
            }
        }
    }
    printf("%d systems (%d singular), decisions differing from FullPivLU: %d, order errors: %d\n", n_sys, n_sing, decisions_differ, bad);
    printf("residual %.2e orthogonality %.2e warm-vs-cold eigenvalue (abs/lmax) %.2e (rel/cond, full rank) %.2e inverse/cond %.2e\n",
           worst_res, worst_orth, worst_warm, worst_eig, worst_inv);
    printf("mean sweeps cold %.2f warm %.2f\n", (double)sweeps_cold / n_sys, (double)sweeps_warm / n_sys);
    const bool pass = bad == 0 && decisions_differ <= n_sys / 10000 && n_sing > 1000 && worst_res < 1e-14 && worst_orth < 1e-14 && worst_warm < 1e-14 &&
                      worst_eig < 1e-14 && worst_inv < 1e-14 && sweeps_warm < sweeps_cold;
    printf(pass ? "K2_FAST_OK\n" : "K2_FAST_FAIL\n");
    return pass ? 0 : 1;
}
