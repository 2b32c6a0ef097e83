// pcreg_math.cuh -- small FP64 linear algebra shared by every kernel (host + device).
//
// 3x3 one-sided Jacobi SVD, symmetric 3x3 Jacobi eigen-solver, Kabsch pose from the 16 weighted
// sums, row-vector 4x4 composition.  Matrices are ROW-MAJOR double[9] / double[16] in here
// (element (r,c) at [r*3+c] / [r*4+c]); the C ABI converts to MATLAB column-major at the edge.
//
// Reference arithmetic restated (not copied): estimateTransform.m:41-71 (centroids, H = m_c*d_c',
// svd, R = V*U', t = cd - R*cm, T = [R t;0 1]'), MATLAB pca/eig documentation for the eigen part.
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define PCREG_HD __host__ __device__ __forceinline__
#else
#define PCREG_HD inline
#endif

namespace pcreg {

PCREG_HD double det3(const double* A) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
           A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// C = A * B (3x3 row-major)
PCREG_HD void mul3(const double* A, const double* B, double* C) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            C[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
}

// C = A * B (4x4 row-major)
PCREG_HD void mul4(const double* A, const double* B, double* C) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[r * 4 + k] * B[k * 4 + c];
            C[r * 4 + c] = s;
        }
}

// One-sided (Hestenes) Jacobi SVD of a 3x3:  A = U * diag(S) * V^T, S descending, V a proper or
// improper orthogonal matrix built from plane rotations (+ the descending sort permutation).
// Accurate to a few ulp of the largest singular value for EVERY singular value (no A^T A squaring).
// Columns of U belonging to singular values below `tiny` are completed by cross products.
PCREG_HD void svd3(const double* A, double* U, double* S, double* V) {
    double G[9];
    for (int i = 0; i < 9; ++i) { G[i] = A[i]; V[i] = 0.0; }
    V[0] = V[4] = V[8] = 1.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            double alpha = 0.0, beta = 0.0, gamma = 0.0;
            for (int i = 0; i < 3; ++i) {
                alpha += G[i * 3 + p] * G[i * 3 + p];
                beta  += G[i * 3 + q] * G[i * 3 + q];
                gamma += G[i * 3 + p] * G[i * 3 + q];
            }
            if (gamma == 0.0 || fabs(gamma) <= 1.1e-16 * sqrt(alpha * beta)) continue;
            rotated = true;
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t);
            const double s = c * t;
            for (int i = 0; i < 3; ++i) {
                const double gp = G[i * 3 + p], gq = G[i * 3 + q];
                G[i * 3 + p] = c * gp - s * gq;
                G[i * 3 + q] = s * gp + c * gq;
                const double vp = V[i * 3 + p], vq = V[i * 3 + q];
                V[i * 3 + p] = c * vp - s * vq;
                V[i * 3 + q] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    double s[3];
    for (int k = 0; k < 3; ++k)
        s[k] = sqrt(G[0 * 3 + k] * G[0 * 3 + k] + G[1 * 3 + k] * G[1 * 3 + k] + G[2 * 3 + k] * G[2 * 3 + k]);
    // descending sort (3 elements), permuting columns of G and V
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2 - a; ++b)
            if (s[b] < s[b + 1]) {
                double tmp = s[b]; s[b] = s[b + 1]; s[b + 1] = tmp;
                for (int i = 0; i < 3; ++i) {
                    tmp = G[i * 3 + b]; G[i * 3 + b] = G[i * 3 + b + 1]; G[i * 3 + b + 1] = tmp;
                    tmp = V[i * 3 + b]; V[i * 3 + b] = V[i * 3 + b + 1]; V[i * 3 + b + 1] = tmp;
                }
            }
    const double tiny = s[0] * 1e-300 + DBL_MIN;
    const double rel  = s[0] * 4.0 * DBL_EPSILON;      // below this a direction is numerically null
    int nvalid = 0;
    for (int k = 0; k < 3; ++k) {
        S[k] = s[k];
        if (s[k] > rel && s[k] > tiny) {
            for (int i = 0; i < 3; ++i) U[i * 3 + k] = G[i * 3 + k] / s[k];
            nvalid = k + 1;
        }
    }
    if (nvalid == 3) return;
    if (nvalid == 2) {           // u2 = u0 x u1
        U[0 * 3 + 2] = U[1 * 3 + 0] * U[2 * 3 + 1] - U[2 * 3 + 0] * U[1 * 3 + 1];
        U[1 * 3 + 2] = U[2 * 3 + 0] * U[0 * 3 + 1] - U[0 * 3 + 0] * U[2 * 3 + 1];
        U[2 * 3 + 2] = U[0 * 3 + 0] * U[1 * 3 + 1] - U[1 * 3 + 0] * U[0 * 3 + 1];
        return;
    }
    if (nvalid == 0) {           // zero matrix: U = I
        for (int i = 0; i < 9; ++i) U[i] = 0.0;
        U[0] = U[4] = U[8] = 1.0;
        return;
    }
    // nvalid == 1: complete u0 to an orthonormal basis
    {
        const double ux = U[0], uy = U[3], uz = U[6];
        int m = 0;                                  // axis least aligned with u0
        if (fabs(uy) < fabs(ux)) m = 1;
        if (fabs(uz) < fabs(m == 0 ? ux : uy)) m = 2;
        double e[3] = {0.0, 0.0, 0.0};
        e[m] = 1.0;
        double bx = uy * e[2] - uz * e[1], by = uz * e[0] - ux * e[2], bz = ux * e[1] - uy * e[0];
        const double bn = sqrt(bx * bx + by * by + bz * bz);
        bx /= bn; by /= bn; bz /= bn;
        U[1] = bx; U[4] = by; U[7] = bz;
        U[2] = uy * bz - uz * by; U[5] = uz * bx - ux * bz; U[8] = ux * by - uy * bx;
    }
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (only the upper triangle of A is trusted,
// the matrix is symmetrised first).  Returns eigenvalues w (unsorted) and eigenvectors as the
// COLUMNS of V (row-major storage).
PCREG_HD void eigsym3(const double* Ain, double* w, double* V) {
    double A[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) A[r * 3 + c] = 0.5 * (Ain[r * 3 + c] + Ain[c * 3 + r]);
    for (int i = 0; i < 9; ++i) V[i] = 0.0;
    V[0] = V[4] = V[8] = 1.0;
    for (int sweep = 0; sweep < 50; ++sweep) {
        const double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        const double dia = fabs(A[0]) + fabs(A[4]) + fabs(A[8]);
        if (off == 0.0 || off <= 1e-18 * dia) break;
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            const double apq = A[p * 3 + q];
            if (apq == 0.0) continue;
            const double app = A[p * 3 + p], aqq = A[q * 3 + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
            const double c = 1.0 / sqrt(1.0 + t * t);
            const double s = t * c;
            // A <- J^T A J with J = rotation in the (p,q) plane
            for (int k = 0; k < 3; ++k) {               // columns p,q
                const double akp = A[k * 3 + p], akq = A[k * 3 + q];
                A[k * 3 + p] = c * akp - s * akq;
                A[k * 3 + q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; ++k) {               // rows p,q
                const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                A[p * 3 + k] = c * apk - s * aqk;
                A[q * 3 + k] = s * apk + c * aqk;
            }
            A[p * 3 + q] = 0.0; A[q * 3 + p] = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                V[k * 3 + p] = c * vkp - s * vkq;
                V[k * 3 + q] = s * vkp + c * vkq;
            }
        }
    }
    w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
}

// Sort eigenpairs (columns of V) ascending (dir = +1) or descending (dir = -1), stable.
PCREG_HD void eigsort3(double* w, double* V, int dir) {
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2 - a; ++b) {
            const bool sw = (dir > 0) ? (w[b] > w[b + 1]) : (w[b] < w[b + 1]);
            if (sw) {
                double tmp = w[b]; w[b] = w[b + 1]; w[b + 1] = tmp;
                for (int i = 0; i < 3; ++i) {
                    tmp = V[i * 3 + b]; V[i * 3 + b] = V[i * 3 + b + 1]; V[i * 3 + b + 1] = tmp;
                }
            }
        }
}

// The 17 running sums of one weighted correspondence set.  "q" is the moving (query / pts2 / "m" of
// estimateTransform.m:42) side, "m" the fixed (model / pts1 / "d" of :41) side, both taken
// RELATIVE TO A PIVOT to keep the one-pass cross-covariance well conditioned.
struct KabschSums {
    double sw;        // sum w
    double sq[3];     // sum w * q
    double sm[3];     // sum w * m
    double sqm[9];    // sum w * q_a * m_b   (row a, col b)
    double swd2;      // sum w * d^2  (for rmse)
};
constexpr int KABSCH_NSUMS = 17;

// dT (row-major, row-vector convention) with [q,1]*dT ~ [m,1]  (estimateTransform(pts1 = m, pts2 = q)).
// pivot_q / pivot_m: what was subtracted from the q side / the m side before summing.
PCREG_HD void kabsch_from_sums(const KabschSums& s, const double* pivot_q, const double* pivot_m, bool reflection_fix, double* dT) {
    const double inv = 1.0 / s.sw;
    double cq[3], cm[3], H[9];
    for (int a = 0; a < 3; ++a) { cq[a] = s.sq[a] * inv; cm[a] = s.sm[a] * inv; }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) H[a * 3 + b] = s.sqm[a * 3 + b] - s.sq[a] * cm[b];   // estimateTransform.m:58
    double U[9], S[3], V[9];
    svd3(H, U, S, V);                                                                      // :60
    double R[9];
    for (int pass = 0; pass < 2; ++pass) {
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)                                                   // R = V * U'  (:62)
                R[r * 3 + c] = V[r * 3 + 0] * U[c * 3 + 0] + V[r * 3 + 1] * U[c * 3 + 1] + V[r * 3 + 2] * U[c * 3 + 2];
        if (pass == 1 || !reflection_fix || det3(R) >= 0.0) break;
        for (int i = 0; i < 3; ++i) V[i * 3 + 2] = -V[i * 3 + 2];                        // flip the weakest direction
    }
    // t = cd - R*cm (:63) with cd = model centroid, cm = query centroid, pivot restored
    double t[3];
    for (int r = 0; r < 3; ++r) {
        const double Rcq = R[r * 3 + 0] * (cq[0] + pivot_q[0]) + R[r * 3 + 1] * (cq[1] + pivot_q[1]) + R[r * 3 + 2] * (cq[2] + pivot_q[2]);
        t[r] = (cm[r] + pivot_m[r]) - Rcq;
    }
    // T = [R t; 0 1]'  (:66-71)
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) dT[r * 4 + c] = R[c * 3 + r];
    dT[0 * 4 + 3] = 0.0; dT[1 * 4 + 3] = 0.0; dT[2 * 4 + 3] = 0.0;
    dT[3 * 4 + 0] = t[0]; dT[3 * 4 + 1] = t[1]; dT[3 * 4 + 2] = t[2]; dT[3 * 4 + 3] = 1.0;
}

// MATLAB eps(x) for x >= 0: spacing of doubles at x.
PCREG_HD double spacing(double x) {
    if (!(x > 0.0)) return 4.9406564584124654e-324;
    int e;
    frexp(x, &e);                       // x = f * 2^e, f in [0.5, 1)
    const double sp = ldexp(1.0, e - 53);
    return sp > 0.0 ? sp : 4.9406564584124654e-324;
}

// MATLAB rank() of an n x 3 matrix given its singular values: count(s > max(n,3) * eps(max s)).
PCREG_HD int rank_from_sv(const double* s, long long n) {
    double smax = s[0];
    if (s[1] > smax) smax = s[1];
    if (s[2] > smax) smax = s[2];
    const double tol = (double)(n > 3 ? n : 3) * spacing(smax);
    return (s[0] > tol) + (s[1] > tol) + (s[2] > tol);
}

}  // namespace pcreg
