/*
 * dcreg_oracle.c - CPU oracle (C99 + OpenMP) for the DCReg hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of TestRunner::Point2PlaneICP_SO3_OpenMP and of the degeneracy engine it calls,
 * used (a) as an independent checker next to the NumPy twin (oracle/dcreg_oracle.py) and (b) as the CPU
 * baseline / `bench.py --impl reference` arm: the reference binary itself cannot be built here (Eigen, PCL/FLANN,
 * yaml-cpp, Ceres, TBB, Open3D are absent and there is no network; its "Ours" stage is a stub in the released
 * source), see DESIGN.md.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference leg may
 * load this library; the product (dcreg_b200) never does.
 *
 * Parity status: PINNED - tests/test_oracle_c.py checks it against the reference's shipped dumps
 * (tests/golden/golden.json: G1 released-code trajectories of four baseline methods to 1e-8, G2 "Ours" to 5e-7)
 * and against the NumPy twin.
 *
 * Reference lines (relative to the DCReg checkout) are cited at each function.  Third-party pieces the reference
 * takes from Eigen 3.3.7 / PCL 1.10 (FLANN) are restated from their published algorithms:
 *   exact k-NN (KdTreeFLANN::nearestKSearch, L2_Simple float distances)  -> kd-tree below
 *   ColPivHouseholderQR::solve                                            -> qr_colpiv_solve
 *   SelfAdjointEigenSolver                                                -> jacobi_sym (cyclic Jacobi)
 *   FullPivLU::inverse / isInvertible                                     -> lu_fullpiv_inverse
 *
 * Two threading modes: mode 0 "reference-faithful" = OpenMP num_threads(8) on the
 * correspondence loop only, serial Jacobian build and serial A^T A (icp_test_runner.cpp:1714, 1863-1915);
 * mode 1 "best-effort" = all host cores on the correspondence loop AND an OpenMP reduction over the 27 sums
 * (the SymmetricHessianComputer pattern, hessian_computer.h:62-123).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ params */
typedef struct {
    double search_radius;
    int max_iterations, detection, handling, use_weight_derivative;
    double conv_rot, conv_trans, cond_thresh, eig_thresh, kappa_target, pcg_tol;
    int pcg_max_iter, fixed_iterations;
    double std_reg_gamma;
    int thread_mode;   /* 0 reference-faithful (8 threads, serial H), 1 best-effort (all cores) */
    int n_threads;     /* > 0: explicit OpenMP team size (bench.py picks it from the CPUs this process may really use);
                          0: mode 0 -> 8 (icp_test_runner.cpp:1714), mode 1 -> omp_get_max_threads() */
} orc_params;

typedef struct {
    int n_eff, n_pt, is_degenerate, pcg_iterations;
    int mask[6];
    double rmse, fitness, objective;
    double H[36], g[6], dx[6], T[16];
    double eig_full[6], lam_schur_rot[3], lam_schur_trans[3];
    double cond_schur_rot, cond_schur_trans, cond_diag_rot, cond_diag_trans, cond_full;
    double P[36];
} orc_iter;

enum { DET_NONE = 0, DET_SCHUR = 1, DET_EVD = 2, DET_SUB = 3, DET_SVD = 4 };
enum { HAND_NONE = 0, HAND_TREG = 1, HAND_AREG = 2, HAND_PCG = 3, HAND_SR = 4, HAND_TSVD = 5 };

/* ------------------------------------------------------------------------------------------------ kd-tree */
typedef struct {
    int n;
    const float* pts;  /* n x 3 */
    int* idx;          /* permutation */
    int* split_dim;    /* per node */
    float* split_val;
    int* left;         /* child node ids, -1 = leaf */
    int* right;
    int* lo;           /* leaf range */
    int* hi;
    int n_nodes;
} kdtree;

#define KD_LEAF 12

static int kd_build_rec(kdtree* t, int lo, int hi) {
    const int node = t->n_nodes++;
    t->lo[node] = lo; t->hi[node] = hi; t->left[node] = t->right[node] = -1;
    if (hi - lo <= KD_LEAF) return node;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; ++i)
        for (int k = 0; k < 3; ++k) {
            const float v = t->pts[3 * t->idx[i] + k];
            if (v < mn[k]) mn[k] = v;
            if (v > mx[k]) mx[k] = v;
        }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    if (mx[dim] == mn[dim]) return node;   /* all identical: keep as leaf */
    /* median by nth_element (quickselect) */
    int l = lo, r = hi - 1;
    const int mid = (lo + hi) / 2;
    while (l < r) {
        const float pv = t->pts[3 * t->idx[(l + r) / 2] + dim];
        int i = l, j = r;
        while (i <= j) {
            while (t->pts[3 * t->idx[i] + dim] < pv) ++i;
            while (t->pts[3 * t->idx[j] + dim] > pv) --j;
            if (i <= j) { const int tmp = t->idx[i]; t->idx[i] = t->idx[j]; t->idx[j] = tmp; ++i; --j; }
        }
        if (j < mid) l = i;
        if (mid < i) r = j;
    }
    t->split_dim[node] = dim;
    t->split_val[node] = t->pts[3 * t->idx[mid] + dim];
    const int lc = kd_build_rec(t, lo, mid);
    const int rc = kd_build_rec(t, mid, hi);
    t->left[node] = lc; t->right[node] = rc;
    return node;
}

static kdtree* kd_build(const float* pts, int n) {
    kdtree* t = (kdtree*)calloc(1, sizeof(kdtree));
    t->n = n; t->pts = pts;
    t->idx = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) t->idx[i] = i;
    const int cap = 2 * (n / (KD_LEAF / 2) + 2);
    t->split_dim = (int*)malloc(sizeof(int) * (size_t)cap);
    t->split_val = (float*)malloc(sizeof(float) * (size_t)cap);
    t->left = (int*)malloc(sizeof(int) * (size_t)cap);
    t->right = (int*)malloc(sizeof(int) * (size_t)cap);
    t->lo = (int*)malloc(sizeof(int) * (size_t)cap);
    t->hi = (int*)malloc(sizeof(int) * (size_t)cap);
    kd_build_rec(t, 0, n);
    return t;
}

static void kd_free(kdtree* t) {
    if (!t) return;
    free(t->idx); free(t->split_dim); free(t->split_val); free(t->left); free(t->right); free(t->lo); free(t->hi);
    free(t);
}

typedef struct { float d2[5]; int id[5]; int cnt; } knn5;

/* FLANN L2_Simple: float differences, float accumulation; ties broken by point index */
static inline void knn_offer(knn5* k, float d2, int id) {
    if (k->cnt == 5 && (d2 > k->d2[4] || (d2 == k->d2[4] && id > k->id[4]))) return;
    int pos = k->cnt < 5 ? k->cnt : 4;
    while (pos > 0 && (k->d2[pos - 1] > d2 || (k->d2[pos - 1] == d2 && k->id[pos - 1] > id))) {
        k->d2[pos] = k->d2[pos - 1]; k->id[pos] = k->id[pos - 1]; --pos;
    }
    k->d2[pos] = d2; k->id[pos] = id;
    if (k->cnt < 5) k->cnt++;
}

static void kd_search(const kdtree* t, int node, const float q[3], knn5* k) {
    if (t->left[node] < 0) {
        for (int i = t->lo[node]; i < t->hi[node]; ++i) {
            const int id = t->idx[i];
            const float* p = t->pts + 3 * id;
            const volatile float ex = q[0] - p[0], ey = q[1] - p[1], ez = q[2] - p[2];
            volatile float d2 = ex * ex;
            d2 = d2 + ey * ey;     /* volatile: no FMA contraction, float rounding at every step */
            d2 = d2 + ez * ez;
            knn_offer(k, d2, id);
        }
        return;
    }
    const int dim = t->split_dim[node];
    const float diff = q[dim] - t->split_val[node];
    const int nearc = diff < 0 ? t->left[node] : t->right[node];
    const int farc = diff < 0 ? t->right[node] : t->left[node];
    kd_search(t, nearc, q, k);
    if (k->cnt < 5 || diff * diff <= k->d2[4]) kd_search(t, farc, q, k);
}

/* ------------------------------------------------------------------------------------------------ dense LA */
/* Householder QR with column pivoting, least-squares / square solve (Eigen ColPivHouseholderQR semantics:
   columns below the nonzero-pivot threshold get a zero solution component). A: m x n row-major (destroyed). */
static void qr_colpiv_solve(int m, int n, double* A, double* b, double* x) {
    const double eps = 2.220446049250313e-16;
    double nu[6], nd[6];
    int perm[6];
    double maxn = 0;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += A[i * n + j] * A[i * n + j];
        nu[j] = nd[j] = sqrt(s); perm[j] = j;
        if (nu[j] > maxn) maxn = nu[j];
    }
    const double thr = (maxn * eps) * (maxn * eps) / (double)m;
    int nz = n;
    for (int k = 0; k < n; ++k) {
        int big = k;
        for (int j = k + 1; j < n; ++j) if (nu[j] > nu[big]) big = j;
        if (nz == n && nu[big] * nu[big] < thr * (double)(m - k)) nz = k;
        if (big != k) {
            for (int i = 0; i < m; ++i) { double t = A[i * n + k]; A[i * n + k] = A[i * n + big]; A[i * n + big] = t; }
            double t = nu[k]; nu[k] = nu[big]; nu[big] = t;
            t = nd[k]; nd[k] = nd[big]; nd[big] = t;
            int tp = perm[k]; perm[k] = perm[big]; perm[big] = tp;
        }
        double tail = 0;
        for (int i = k + 1; i < m; ++i) tail += A[i * n + k] * A[i * n + k];
        const double c0 = A[k * n + k];
        double tau, beta;
        if (tail <= 2.2250738585072014e-308) { tau = 0; beta = c0; for (int i = k + 1; i < m; ++i) A[i * n + k] = 0; }
        else {
            beta = sqrt(c0 * c0 + tail);
            if (c0 >= 0) beta = -beta;
            for (int i = k + 1; i < m; ++i) A[i * n + k] /= (c0 - beta);
            tau = (beta - c0) / beta;
        }
        A[k * n + k] = beta;
        for (int j = k + 1; j < n; ++j) {
            double t = A[k * n + j];
            for (int i = k + 1; i < m; ++i) t += A[i * n + k] * A[i * n + j];
            A[k * n + j] -= tau * t;
            for (int i = k + 1; i < m; ++i) A[i * n + j] -= tau * A[i * n + k] * t;
        }
        if (k < nz) {
            double t = b[k];
            for (int i = k + 1; i < m; ++i) t += A[i * n + k] * b[i];
            b[k] -= tau * t;
            for (int i = k + 1; i < m; ++i) b[i] -= tau * A[i * n + k] * t;
        }
        for (int j = k + 1; j < n; ++j) {
            if (nu[j] != 0) {
                double t = fabs(A[k * n + j]) / nu[j];
                t = (1 + t) * (1 - t);
                if (t < 0) t = 0;
                const double r = nu[j] / nd[j];
                if (t * r * r <= 1.4901161193847656e-08) {
                    double s = 0;
                    for (int i = k + 1; i < m; ++i) s += A[i * n + j] * A[i * n + j];
                    nd[j] = nu[j] = sqrt(s);
                } else nu[j] *= sqrt(t);
            }
        }
    }
    double c[6] = {0};
    for (int i = nz - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < nz; ++j) s -= A[i * n + j] * c[j];
        c[i] = s / A[i * n + i];
    }
    for (int i = 0; i < n; ++i) x[i] = 0;
    for (int i = 0; i < nz; ++i) x[perm[i]] = c[i];
}

/* cyclic Jacobi for a symmetric n x n matrix (n <= 6); eigenvalues ascending, eigenvectors in columns of V */
static void jacobi_sym(int n, const double* Ain, double* w, double* V) {
    double A[36];
    memcpy(A, Ain, sizeof(double) * (size_t)(n * n));
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, dg = 0;
        for (int i = 0; i < n; ++i) { dg += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j]; }
        if (off <= 1e-34 * dg || off == 0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (fabs(apq) < 1e-300) { A[p * n + q] = A[q * n + p] = 0; continue; }
                const double th = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c, tau = s / (1 + c);
                A[p * n + p] -= t * apq; A[q * n + q] += t * apq; A[p * n + q] = A[q * n + p] = 0;
                for (int k = 0; k < n; ++k) if (k != p && k != q) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = A[p * n + k] = akp - s * (akq + tau * akp);
                    A[k * n + q] = A[q * n + k] = akq + s * (akp - tau * akq);
                }
                for (int k = 0; k < n; ++k) {
                    const double vp = V[k * n + p], vq = V[k * n + q];
                    V[k * n + p] = vp - s * (vq + tau * vp);
                    V[k * n + q] = vq + s * (vp - tau * vq);
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
    for (int i = 0; i < n - 1; ++i) {
        int m = i;
        for (int j = i + 1; j < n; ++j) if (w[j] < w[m]) m = j;
        if (m != i) {
            double t = w[i]; w[i] = w[m]; w[m] = t;
            for (int k = 0; k < n; ++k) { t = V[k * n + i]; V[k * n + i] = V[k * n + m]; V[k * n + m] = t; }
        }
    }
}

/* Gauss-Jordan inverse with full pivoting; returns 0 if singular (pivot <= eps*n*maxpivot), FullPivLU-like */
static int inv_fullpiv(int n, const double* Ain, double* Inv) {
    double A[36], B[36];
    int rp[6], cp[6];
    memcpy(A, Ain, sizeof(double) * (size_t)(n * n));
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) B[i * n + j] = (i == j);
    double maxpiv = 0, piv[6];
    for (int k = 0; k < n; ++k) {
        int br = k, bc = k; double bv = -1;
        for (int i = k; i < n; ++i) for (int j = k; j < n; ++j) if (fabs(A[i * n + j]) > bv) { bv = fabs(A[i * n + j]); br = i; bc = j; }
        if (bv == 0) return 0;
        if (bv > maxpiv) maxpiv = bv;
        rp[k] = br; cp[k] = bc;
        if (br != k) for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[br * n + j]; A[br * n + j] = t; t = B[k * n + j]; B[k * n + j] = B[br * n + j]; B[br * n + j] = t; }
        if (bc != k) for (int i = 0; i < n; ++i) { double t = A[i * n + k]; A[i * n + k] = A[i * n + bc]; A[i * n + bc] = t; }
        piv[k] = A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            const double f = A[i * n + k] / A[k * n + k];
            for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
            for (int j = 0; j < n; ++j) B[i * n + j] -= f * B[k * n + j];
        }
    }
    for (int k = 0; k < n; ++k) if (fabs(piv[k]) <= 2.220446049250313e-16 * n * maxpiv) return 0;
    /* back substitution: A is upper triangular now; solve A Y = B */
    for (int c = 0; c < n; ++c)
        for (int i = n - 1; i >= 0; --i) {
            double s = B[i * n + c];
            for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * B[j * n + c];
            B[i * n + c] = s / A[i * n + i];
        }
    /* undo column permutation: rows of Y are in permuted variable order */
    for (int k = n - 1; k >= 0; --k) if (cp[k] != k)
        for (int j = 0; j < n; ++j) { double t = B[k * n + j]; B[k * n + j] = B[cp[k] * n + j]; B[cp[k] * n + j] = t; }
    (void)rp;
    memcpy(Inv, B, sizeof(double) * (size_t)(n * n));
    return 1;
}

static void mat3mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0; for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j]; C[i * 3 + j] = s; }
}

/* ------------------------------------------------------------------------------------------------ K2 oracle */
static int pcg6(const double* H, const double* g, const double* P, int max_it, double tol, double* x) {
    double r[6], z[6], p[6], Hp[6];
    for (int i = 0; i < 6; ++i) { x[i] = 0; r[i] = g[i]; }
    double rz = 0;
    for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += P[i * 6 + j] * r[j]; z[i] = s; }
    for (int i = 0; i < 6; ++i) { p[i] = z[i]; rz += r[i] * z[i]; }
    int it;
    for (it = 1; it <= max_it; ++it) {
        double pHp = 0, rn = 0;
        for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += H[i * 6 + j] * p[j]; Hp[i] = s; pHp += p[i] * s; }
        const double al = rz / pHp;
        for (int i = 0; i < 6; ++i) { x[i] += al * p[i]; r[i] -= al * Hp[i]; rn += r[i] * r[i]; }
        if (sqrt(rn) < tol) break;
        double rz2 = 0;
        for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += P[i * 6 + j] * r[j]; z[i] = s; rz2 += r[i] * s; }
        const double be = rz2 / rz;
        for (int i = 0; i < 6; ++i) p[i] = z[i] + be * p[i];
        rz = rz2;
    }
    return it > max_it ? max_it : it;
}

static void solve_qr6(const double* H, const double* g, double* x) {
    double A[36], b[6];
    memcpy(A, H, sizeof(A)); memcpy(b, g, sizeof(b));
    qr_colpiv_solve(6, 6, A, b, x);
}

/* DCReg::analyzeDegeneracy + solveDegenerateSystem (dcreg.hpp:45-264), Schur block (icp_test_runner.cpp:2418-2469),
   Schur detection / preconditioner / PCG from the paper's Alg. 1, 3 and Eq. 18-21, 43-46 */
void orc_analyze_and_solve(const double* H, const double* g, const orc_params* p, orc_iter* o) {
    double lam[6], V[36];
    jacobi_sym(6, H, lam, V);
    memcpy(o->eig_full, lam, sizeof(lam));
    int order[6] = {0, 1, 2, 3, 4, 5};
    for (int i = 0; i < 5; ++i) { int m = i; for (int j = i + 1; j < 6; ++j) if (fabs(lam[order[j]]) > fabs(lam[order[m]])) m = j; int t = order[i]; order[i] = order[m]; order[m] = t; }
    double sv[6];
    for (int i = 0; i < 6; ++i) sv[i] = fabs(lam[order[i]]);
    o->cond_full = sv[5] > 1e-12 ? sv[0] / sv[5] : INFINITY;
    double HRR[9], Htt[9], HRt[9], HtR[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        HRR[i * 3 + j] = H[i * 6 + j]; Htt[i * 3 + j] = H[(i + 3) * 6 + j + 3];
        HRt[i * 3 + j] = H[i * 6 + j + 3]; HtR[i * 3 + j] = H[(i + 3) * 6 + j]; }
    double l3[3], V3[9];
    jacobi_sym(3, HRR, l3, V3); o->cond_diag_rot = l3[2] / fmax(l3[0], 1e-12);
    jacobi_sym(3, Htt, l3, V3); o->cond_diag_trans = l3[2] / fmax(l3[0], 1e-12);
    double Vr[9], Vt[9], HttI[9], HRRI[9];
    const int okt = inv_fullpiv(3, Htt, HttI), okr = inv_fullpiv(3, HRR, HRRI);
    const int schur_ok = okt && okr;
    for (int i = 0; i < 36; ++i) o->P[i] = (i % 7 == 0);
    memset(o->mask, 0, sizeof(o->mask)); o->is_degenerate = 0; o->pcg_iterations = 0;
    if (schur_ok) {
        double T1[9], T2[9], SR[9], St[9];
        mat3mul(HRt, HttI, T1); mat3mul(T1, HtR, T2); for (int i = 0; i < 9; ++i) SR[i] = HRR[i] - T2[i];
        mat3mul(HtR, HRRI, T1); mat3mul(T1, HRt, T2); for (int i = 0; i < 9; ++i) St[i] = Htt[i] - T2[i];
        for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) {
            double a = 0.5 * (SR[i * 3 + j] + SR[j * 3 + i]); SR[i * 3 + j] = SR[j * 3 + i] = a;
            a = 0.5 * (St[i * 3 + j] + St[j * 3 + i]); St[i * 3 + j] = St[j * 3 + i] = a; }
        jacobi_sym(3, SR, o->lam_schur_rot, Vr);
        jacobi_sym(3, St, o->lam_schur_trans, Vt);
        o->cond_schur_rot = o->lam_schur_rot[2] / fmax(o->lam_schur_rot[0], 1e-12);
        o->cond_schur_trans = o->lam_schur_trans[2] / fmax(o->lam_schur_trans[0], 1e-12);
    } else {
        for (int i = 0; i < 3; ++i) o->lam_schur_rot[i] = o->lam_schur_trans[i] = NAN;
        o->cond_schur_rot = o->cond_schur_trans = INFINITY;
    }
    switch (p->detection) {
        case DET_SCHUR:
            if (schur_ok) for (int blk = 0; blk < 2; ++blk) {
                const double* l = blk ? o->lam_schur_trans : o->lam_schur_rot;
                const double* Vb = blk ? Vt : Vr;
                double lt[3];
                for (int i = 0; i < 3; ++i) {
                    if (l[2] / fmax(l[i], 1e-12) > p->cond_thresh) { o->mask[blk * 3 + i] = 1; o->is_degenerate = 1; }
                    lt[i] = fmax(l[i], l[2] / p->kappa_target);
                }
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                    double s = 0; for (int k = 0; k < 3; ++k) s += Vb[i * 3 + k] * Vb[j * 3 + k] / lt[k];
                    o->P[(blk * 3 + i) * 6 + blk * 3 + j] = s; }
            }
            break;
        case DET_EVD:
            for (int i = 0; i < 6; ++i) if (lam[i] < p->eig_thresh) { o->mask[i] = 1; o->is_degenerate = 1; }
            break;
        case DET_SVD:
            o->is_degenerate = o->cond_full > p->cond_thresh;
            if (o->is_degenerate) for (int i = 0; i < 6; ++i) if (lam[5] / lam[i] > p->cond_thresh) o->mask[i] = 1;
            break;
        default: break;
    }
    double* dx = o->dx;
    switch (p->handling) {
        case HAND_TREG: {
            double Hr[36]; memcpy(Hr, H, sizeof(Hr));
            if (o->is_degenerate) for (int i = 0; i < 6; ++i) Hr[i * 7] += p->std_reg_gamma;
            solve_qr6(Hr, g, dx); break; }
        case HAND_PCG:
            if (o->is_degenerate) o->pcg_iterations = pcg6(H, g, o->P, p->pcg_max_iter, p->pcg_tol, dx);
            else solve_qr6(H, g, dx);
            break;
        case HAND_SR: {
            double x0[6]; solve_qr6(H, g, x0);
            if (o->is_degenerate) {
                int good = 0; for (int i = 0; i < 6; ++i) dx[i] = 0;
                for (int k = 0; k < 6; ++k) if (!o->mask[k]) {
                    ++good; double d = 0; for (int i = 0; i < 6; ++i) d += V[i * 6 + k] * x0[i];
                    for (int i = 0; i < 6; ++i) dx[i] += V[i * 6 + k] * d; }
                if (!good) for (int i = 0; i < 6; ++i) dx[i] = 0;
            } else memcpy(dx, x0, sizeof(x0));
            break; }
        case HAND_TSVD: {
            int kept = 0; for (int i = 0; i < 6; ++i) dx[i] = 0;
            for (int i = 0; i < 6; ++i) if (!o->mask[i] && sv[i] > 1e-9) {   /* mask/sigma index quirk kept (dcreg.hpp:232-237) */
                ++kept; const int e = order[i]; double d = 0;
                for (int r = 0; r < 6; ++r) d += V[r * 6 + e] * g[r];
                const double sc = (lam[e] >= 0 ? 1.0 : -1.0) * d / sv[i];
                for (int r = 0; r < 6; ++r) dx[r] += V[r * 6 + e] * sc; }
            if (!kept) for (int i = 0; i < 6; ++i) dx[i] = 0;
            break; }
        default: solve_qr6(H, g, dx); break;
    }
}

/* ------------------------------------------------------------------------------------------------ the loop */
typedef struct { const float* src; int n; const float* tgt; int m; kdtree* tree; } orc_scene;

orc_scene* orc_scene_create(const float* src, int n, const float* tgt, int m) {
    orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
    s->src = src; s->n = n; s->tgt = tgt; s->m = m;
    s->tree = kd_build(tgt, m);          /* outside the timed region, like ICPContext::setTargetCloud (utils.hpp:393-424) */
    return s;
}
void orc_scene_destroy(orc_scene* s) { if (s) { kd_free(s->tree); free(s); } }

static void so3_exp(const double* w, double* E) {      /* math_utils.hpp:20-33 */
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (th < 1e-10) { const double e[9] = {1, -w[2], w[1], w[2], 1, -w[0], -w[1], w[0], 1}; memcpy(E, e, sizeof(e)); return; }
    const double a[3] = {w[0] / th, w[1] / th, w[2] / th};
    const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    double K2[9]; mat3mul(K, K, K2);
    const double s = sin(th), c1 = 1 - cos(th);
    for (int i = 0; i < 9; ++i) E[i] = (i % 4 == 0) + s * K[i] + c1 * K2[i];
}

/* per-slot stage S1 (icp_test_runner.cpp:1714-1813): returns 1 if the slot is an effective correspondence */
static inline int slot_s1(const orc_scene* sc, int i, const double* R, const double* t, double r2max, int use_wd,
                          double* nrm, double* r_out, double* s_out, double* ds_out, int* near5) {
    const float* p = sc->src + 3 * i;
    float q[3];
    for (int k = 0; k < 3; ++k) q[k] = (float)(R[k * 3] * p[0] + R[k * 3 + 1] * p[1] + R[k * 3 + 2] * p[2] + t[k]);   /* utils.hpp:630-636 */
    knn5 kn; kn.cnt = 0;
    kd_search(sc->tree, 0, q, &kn);
    *near5 = 0;
    if (kn.cnt < 5 || !((double)kn.d2[4] < r2max)) return 0;
    *near5 = 1;
    double A[15], A0[15], b[5], x[3];
    for (int j = 0; j < 5; ++j) { for (int k = 0; k < 3; ++k) A0[j * 3 + k] = A[j * 3 + k] = sc->tgt[3 * kn.id[j] + k]; b[j] = -1; }
    qr_colpiv_solve(5, 3, A, b, x);
    const double ps = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    if (!(ps >= 1e-6)) return 0;
    const double n0 = x[0] / ps, n1 = x[1] / ps, n2 = x[2] / ps, d = 1 / ps;
    double worst = 0;
    for (int j = 0; j < 5; ++j) { double e = n0 * A0[j * 3] + n1 * A0[j * 3 + 1] + n2 * A0[j * 3 + 2] + d; e *= e; if (e > worst) worst = e; }
    if (!(worst < 0.04)) return 0;
    const double r = n0 * q[0] + n1 * q[1] + n2 * q[2] + d;
    const double s = fmax(0.0, 1 - 0.9 * fabs(r));
    double ds = 0;
    if (use_wd && s > 0 && s < 1) ds = r > 0 ? -0.9 : 0.9;
    if (!(s > 0.1)) return 0;
    nrm[0] = n0; nrm[1] = n1; nrm[2] = n2; *r_out = r; *s_out = s; *ds_out = ds;
    return 1;
}

/* Jacobian row (math_utils.hpp:102-121, icp_test_runner.cpp:1863-1907) with the float32 round trips of coeff */
static inline void jac_row(const float* p, const double* R, const double* nrm, double r, double s, double ds, double* J, double* b) {
    const double nu[3] = {(double)(float)(s * nrm[0]) / s, (double)(float)(s * nrm[1]) / s, (double)(float)(s * nrm[2]) / s};
    double a[3];
    for (int k = 0; k < 3; ++k) a[k] = nu[0] * R[k] + nu[1] * R[3 + k] + nu[2] * R[6 + k];   /* n^T R */
    const double w = s + r * ds;
    J[0] = w * (p[1] * a[2] - p[2] * a[1]);
    J[1] = w * (p[2] * a[0] - p[0] * a[2]);
    J[2] = w * (p[0] * a[1] - p[1] * a[0]);
    J[3] = w * a[0]; J[4] = w * a[1]; J[5] = w * a[2];
    *b = -(double)(float)(s * r);
}

/* TestRunner::Point2PlaneICP_SO3_OpenMP (icp_test_runner.cpp:1611-2060).
   Returns status: 0 ok, 1 not enough points, 2 non-finite update.  log may be NULL. */
int orc_icp_run(const orc_scene* sc, const orc_params* prm, const double* T_init, double* T_out, orc_iter* log, int log_cap,
                int* n_iter, int* converged) {
    double R[9], t[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r * 3 + c] = T_init[r * 4 + c]; t[r] = T_init[r * 4 + 3]; }
    const int n = sc->n;
    double* nrm = (double*)malloc(sizeof(double) * 3 * (size_t)n);
    double* rr = (double*)malloc(sizeof(double) * (size_t)n);
    double* ss = (double*)malloc(sizeof(double) * (size_t)n);
    double* dd = (double*)malloc(sizeof(double) * (size_t)n);
    unsigned char* flag = (unsigned char*)malloc((size_t)n);
    int status = 0, iters = 0; *converged = 0;
#ifdef _OPENMP
    const int nthreads = prm->n_threads > 0 ? prm->n_threads : (prm->thread_mode == 0 ? 8 : omp_get_max_threads());
#else
    const int nthreads = 1;
#endif
    const double r2max = prm->search_radius * prm->search_radius;
    for (int it = 0; it < prm->max_iterations; ++it) {
        int n_eff = 0, n_pt = 0; double sr2 = 0;
#pragma omp parallel for num_threads(nthreads) reduction(+ : n_eff, n_pt, sr2) schedule(dynamic, 256)
        for (int i = 0; i < n; ++i) {
            int near5;
            flag[i] = (unsigned char)slot_s1(sc, i, R, t, r2max, prm->use_weight_derivative, nrm + 3 * i, rr + i, ss + i, dd + i, &near5);
            n_pt += near5;
            if (flag[i]) { n_eff++; sr2 += rr[i] * rr[i]; }
        }
        if (n_eff < 10) { status = 1; iters = it + 1; break; }                       /* :1847-1854 */
        double H[36] = {0}, g[6] = {0}, sb2 = 0;
        if (prm->thread_mode == 0) {                                                  /* serial build + A^T A */
            for (int i = 0; i < n; ++i) if (flag[i]) {
                double J[6], b; jac_row(sc->src + 3 * i, R, nrm + 3 * i, rr[i], ss[i], dd[i], J, &b);
                for (int a = 0; a < 6; ++a) { for (int c = a; c < 6; ++c) H[a * 6 + c] += J[a] * J[c]; g[a] += J[a] * b; }
                sb2 += b * b;
            }
        } else {
            double acc[28] = {0};
#pragma omp parallel for num_threads(nthreads) reduction(+ : acc[:28]) schedule(static)
            for (int i = 0; i < n; ++i) if (flag[i]) {
                double J[6], b; jac_row(sc->src + 3 * i, R, nrm + 3 * i, rr[i], ss[i], dd[i], J, &b);
                int k = 0;
                for (int a = 0; a < 6; ++a) for (int c = a; c < 6; ++c) acc[k++] += J[a] * J[c];
                for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * b;
                acc[27] += b * b;
            }
            int k = 0;
            for (int a = 0; a < 6; ++a) for (int c = a; c < 6; ++c) H[a * 6 + c] = acc[k++];
            for (int a = 0; a < 6; ++a) g[a] = acc[21 + a];
            sb2 = acc[27];
        }
        for (int a = 0; a < 6; ++a) for (int c = 0; c < a; ++c) H[a * 6 + c] = H[c * 6 + a];
        orc_iter tmp; orc_iter* o = (log && it < log_cap) ? &log[it] : &tmp;
        o->n_eff = n_eff; o->n_pt = n_pt; o->rmse = sqrt(sr2 / n_eff); o->fitness = (double)n_pt / n; o->objective = 0.5 * sb2;
        memcpy(o->H, H, sizeof(H)); memcpy(o->g, g, sizeof(g));
        orc_analyze_and_solve(H, g, prm, o);
        int finite = 1; for (int a = 0; a < 6; ++a) if (!isfinite(o->dx[a])) finite = 0;
        if (!finite) { status = 2; iters = it; break; }                               /* :1942-1950 */
        double E[9], Rn[9]; so3_exp(o->dx, E); mat3mul(R, E, Rn);                     /* boxplus, math_utils.hpp:158-166 */
        for (int k = 0; k < 3; ++k) t[k] += R[k * 3] * o->dx[3] + R[k * 3 + 1] * o->dx[4] + R[k * 3 + 2] * o->dx[5];
        memcpy(R, Rn, sizeof(R));
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) o->T[r * 4 + c] = R[r * 3 + c]; o->T[r * 4 + 3] = t[r]; }
        o->T[12] = o->T[13] = o->T[14] = 0; o->T[15] = 1;
        iters = it + 1;
        const double dR = sqrt(o->dx[0] * o->dx[0] + o->dx[1] * o->dx[1] + o->dx[2] * o->dx[2]);
        const double dT = sqrt(o->dx[3] * o->dx[3] + o->dx[4] * o->dx[4] + o->dx[5] * o->dx[5]);
        if (!prm->fixed_iterations && dR < prm->conv_rot && dT < prm->conv_trans) { *converged = 1; break; }   /* :1998-2002 */
    }
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T_out[r * 4 + c] = R[r * 3 + c]; T_out[r * 4 + 3] = t[r]; }
    T_out[12] = T_out[13] = T_out[14] = 0; T_out[15] = 1;
    *n_iter = iters;
    free(nrm); free(rr); free(ss); free(dd); free(flag);
    return status;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int orc_sizeof_iter(void) { return (int)sizeof(orc_iter); }
int orc_sizeof_params(void) { return (int)sizeof(orc_params); }
