// Host-side check (no GPU needed): colpiv_qr_solve_reg<5,3> == colpiv_qr_solve<5,3> bit for bit on random,
// ill-conditioned and rank-deficient 5x3 systems.   nvcc -O2 -o /tmp/test_qr_reg tools/test_qr_reg.cu && /tmp/test_qr_reg
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include "../dcreg_b200/csrc/small_la.cuh"

static unsigned long long st = 88172645463325252ull;
static double rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; }

int main() {
    long bad = 0, n = 0;
    for (int trial = 0; trial < 400000; ++trial) {
        double P[5][3];
        const int kind = trial % 8;
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 3; ++j) P[i][j] = (float)((rnd() - 0.5) * (kind == 1 ? 80.0 : 2.0));
        if (kind == 2) for (int i = 0; i < 5; ++i) P[i][2] = 0.0;                         // z = 0 floor
        if (kind == 3) for (int i = 0; i < 5; ++i) P[i][1] = (float)(2.0 * P[i][0]);      // dependent columns
        if (kind == 4) { for (int j = 0; j < 3; ++j) { P[1][j] = P[0][j]; P[3][j] = P[2][j]; } }   // duplicate rows
        if (kind == 5) for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) P[i][j] = (float)(i * (j + 1) * 0.25);   // collinear lattice
        if (kind == 6) for (int i = 0; i < 5; ++i) P[i][2] = (float)(1e-6 * (rnd() - 0.5));   // thin
        if (kind == 7) for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) P[i][j] = 0.0;
        double A1[15], b1[5], x1[3], A2[5][3], b2[5], x2[3];
        for (int i = 0; i < 5; ++i) { for (int j = 0; j < 3; ++j) { A1[i * 3 + j] = P[i][j]; A2[i][j] = P[i][j]; } b1[i] = b2[i] = -1.0; }
        dla::colpiv_qr_solve<5, 3>(A1, b1, x1);
        dla::colpiv_qr_solve_reg<5, 3>(A2, b2, x2);
        ++n;
        if (memcmp(x1, x2, sizeof(x1)) != 0) {
            bool bothnan = true;
            for (int j = 0; j < 3; ++j) if (!(std::isnan(x1[j]) && std::isnan(x2[j])) && x1[j] != x2[j]) bothnan = false;
            if (!bothnan) { if (bad < 5) printf("mismatch kind %d: %.17g %.17g %.17g vs %.17g %.17g %.17g\n", kind, x1[0], x1[1], x1[2], x2[0], x2[1], x2[2]); ++bad; }
        }
    }
    printf("%ld systems, %ld mismatches\n", n, bad);
    return bad != 0;
}
