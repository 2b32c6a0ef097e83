// kabsch_ransac.cu -- batched estimateTransform (estimateTransform.m:2-74) and the hypothesis
// scoring loop of ransac.m:40-73, one warp per hypothesis, sample triplets supplied by the host.
#include <math.h>
#include <stdio.h>
#include <float.h>
#include <algorithm>
#include <vector>
#include <thread>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"
#include "pcreg_math.cuh"

namespace pcreg {

// ---- small-N two-pass Kabsch in registers (N = 3 with the reference's synthetic 4th point, or 4) ----
// p1 = pts1 ("d" of estimateTransform.m:41), p2 = pts2 ("m" of :42); T row-major row-vector with
// [p2,1]*T = [p1,1].  Returns 0, or 1 where the reference's rank guard (:11-14) returns [].
__device__ __forceinline__ void sort3(double& a, double& b, double& c) {
    double t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
}

__device__ __noinline__ int kabsch3(const double* __restrict__ P1 /*[9] row i = point i*/, const double* __restrict__ P2,
                                    bool reflection_fix, double* __restrict__ T) {
    // rank guard on the RAW 3x3 matrices (rows = points), exact one-sided Jacobi singular values
    {
        double U[9], S[3], V[9];
        svd3(P1, U, S, V);
        if (rank_from_sv(S, 3) < 3) return 1;
        svd3(P2, U, S, V);
        if (rank_from_sv(S, 3) < 2) return 1;
    }
    double a[4][3], b[4][3];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) { a[i][k] = P1[i * 3 + k]; b[i][k] = P2[i * 3 + k]; }
    // 4th point: mean + unit normal * median edge length (estimateTransform.m:18-37)
    for (int set = 0; set < 2; ++set) {
        double (*p)[3] = set == 0 ? a : b;
        double c[3], e1[3], e2[3], n[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = (p[0][k] + p[1][k] + p[2][k]) / 3.0;
            e1[k] = p[2][k] - p[1][k];
            e2[k] = p[2][k] - p[0][k];
        }
        n[0] = e1[1] * e2[2] - e1[2] * e2[1];
        n[1] = e1[2] * e2[0] - e1[0] * e2[2];
        n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        // vecnorm(pts - circshift(pts,1,1)): rows 0-2, 1-0, 2-1
        double l0 = norm3_exact(p[0][0] - p[2][0], p[0][1] - p[2][1], p[0][2] - p[2][2]);
        double l1 = norm3_exact(p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]);
        double l2 = norm3_exact(p[2][0] - p[1][0], p[2][1] - p[1][1], p[2][2] - p[1][2]);
        sort3(l0, l1, l2);
        const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int k = 0; k < 3; ++k) p[3][k] = c[k] + (n[k] / nn) * l1;
    }
    double c1[3], c2[3];
    for (int k = 0; k < 3; ++k) {
        c1[k] = (a[0][k] + a[1][k] + a[2][k] + a[3][k]) * 0.25;
        c2[k] = (b[0][k] + b[1][k] + b[2][k] + b[3][k]) * 0.25;
    }
    KabschSums s;
    s.sw = 4.0; s.swd2 = 0.0;
    for (int k = 0; k < 3; ++k) { s.sq[k] = 0.0; s.sm[k] = 0.0; }
    for (int k = 0; k < 9; ++k) s.sqm[k] = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int r = 0; r < 3; ++r) {
            const double q = b[i][r] - c2[r];
            s.sq[r] += q;
            s.sm[r] += a[i][r] - c1[r];
            for (int cc = 0; cc < 3; ++cc) s.sqm[r * 3 + cc] += q * (a[i][cc] - c1[cc]);
        }
    kabsch_from_sums(s, c2, c1, reflection_fix, T);
    return 0;
}

// 29 running sums of an unweighted / weighted pair set: the 17 Kabsch sums + the two raw Gram matrices
// (upper triangles) that the rank guard needs.
constexpr int NS_FIT = 29;
__device__ __forceinline__ void fit_accumulate(double* s, double w, const double* p1, const double* p2,
                                               const double* piv1, const double* piv2) {
    const double q0 = p2[0] - piv2[0], q1 = p2[1] - piv2[1], q2 = p2[2] - piv2[2];
    const double m0 = p1[0] - piv1[0], m1 = p1[1] - piv1[1], m2 = p1[2] - piv1[2];
    const double wq0 = w * q0, wq1 = w * q1, wq2 = w * q2;
    s[0] += w;
    s[1] += wq0; s[2] += wq1; s[3] += wq2;
    s[4] += w * m0; s[5] += w * m1; s[6] += w * m2;
    s[7] += wq0 * m0; s[8] += wq0 * m1; s[9] += wq0 * m2;
    s[10] += wq1 * m0; s[11] += wq1 * m1; s[12] += wq1 * m2;
    s[13] += wq2 * m0; s[14] += wq2 * m1; s[15] += wq2 * m2;
    // raw Gram (rank guard uses the RAW, uncentred, unweighted points: estimateTransform.m:11)
    s[17] += p1[0] * p1[0]; s[18] += p1[0] * p1[1]; s[19] += p1[0] * p1[2];
    s[20] += p1[1] * p1[1]; s[21] += p1[1] * p1[2]; s[22] += p1[2] * p1[2];
    s[23] += p2[0] * p2[0]; s[24] += p2[0] * p2[1]; s[25] += p2[0] * p2[2];
    s[26] += p2[1] * p2[1]; s[27] += p2[1] * p2[2]; s[28] += p2[2] * p2[2];
}

// rank(pts) >= need of an n x 3 point matrix (estimateTransform.m:11: MATLAB rank = singular values above
// max(size) * eps(largest)).  The Gram matrix (summed in the same pass as the Kabsch sums) gives the singular values as
// square roots of its eigenvalues, but only down to ~1e-8 of the largest.  Directions whose estimate is below 1e-6 of the
// largest are therefore "uncertain": the caller measures them again DIRECTLY on the data -- sqrt(sum_i (p_i . v)^2) along the
// Gram eigenvector v, a second pass over the points -- which resolves them to a few eps of the largest singular value,
// the scale of MATLAB's own tolerance (tests/test_gpu_kabsch_ransac.py brackets the agreement).
struct RankProbe {
    double V[9];        // eigenvectors of the Gram matrix (columns)
    double smax, tol;   // largest singular value, MATLAB's tolerance
    int sure;           // directions certainly above the tolerance
    int uncertain[3];   // columns of V to measure again
    int nunc;
};
__device__ __forceinline__ void rank_probe(const double* g6, long long n, RankProbe& r) {
    double A[9] = {g6[0], g6[1], g6[2], g6[1], g6[3], g6[4], g6[2], g6[4], g6[5]};
    double w[3];
    eigsym3(A, w, r.V);
    double sv[3];
    for (int k = 0; k < 3; ++k) sv[k] = w[k] > 0.0 ? sqrt(w[k]) : 0.0;
    r.smax = fmax(sv[0], fmax(sv[1], sv[2]));
    r.tol = (double)(n > 3 ? n : 3) * spacing(r.smax);
    r.sure = 0; r.nunc = 0;
    for (int k = 0; k < 3; ++k) {
        if (sv[k] > 1.0e-6 * r.smax && sv[k] > r.tol) ++r.sure;
        else r.uncertain[r.nunc++] = k;
    }
}
// squared projections of one point on the uncertain directions (accumulated by the caller, then reduced)
__device__ __forceinline__ void rank_refine_accumulate(const RankProbe& r, const double* p, double* acc /*[3]*/) {
    for (int u = 0; u < r.nunc; ++u) {
        const int k = r.uncertain[u];
        const double t = p[0] * r.V[0 * 3 + k] + p[1] * r.V[1 * 3 + k] + p[2] * r.V[2 * 3 + k];
        acc[u] += t * t;
    }
}
__device__ __forceinline__ int rank_final(const RankProbe& r, const double* acc) {
    int rank = r.sure;
    for (int u = 0; u < r.nunc; ++u) rank += (sqrt(acc[u]) > r.tol) ? 1 : 0;
    return rank;
}

__device__ __forceinline__ void kabsch_from_fit_sums(const double* s, const double* piv1, const double* piv2, bool reflection_fix, double* T) {
    KabschSums ks;
    ks.sw = s[0];
    for (int k = 0; k < 3; ++k) { ks.sq[k] = s[1 + k]; ks.sm[k] = s[4 + k]; }
    for (int k = 0; k < 9; ++k) ks.sqm[k] = s[7 + k];
    ks.swd2 = 0.0;
    kabsch_from_sums(ks, piv2, piv1, reflection_fix, T);
}

// ---- pcreg_kabsch_batch: one block per problem ----------------------------------------------------
__global__ void __launch_bounds__(128) k_kabsch_batch(const double* __restrict__ p1, const double* __restrict__ p2,
                                                      const double* __restrict__ w, int64_t ld,
                                                      const int64_t* __restrict__ offsets, int reflection_fix,
                                                      double* __restrict__ T_rm, int32_t* __restrict__ status) {
    __shared__ double red[NS_FIT * 32];
    const int64_t b = blockIdx.x;
    const int64_t r0 = offsets[b], n = offsets[b + 1] - r0;
    const int tid = threadIdx.x;
    double T[16];
    int st = 1;
    if (n == 3 && !w) {
        if (tid == 0) {
            double P1[9], P2[9];
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) { P1[i * 3 + k] = p1[k * ld + r0 + i]; P2[i * 3 + k] = p2[k * ld + r0 + i]; }
            st = kabsch3(P1, P2, reflection_fix != 0, T);
        }
    } else if (n >= 1) {
        double piv1[3], piv2[3];
        for (int k = 0; k < 3; ++k) { piv1[k] = p1[k * ld + r0]; piv2[k] = p2[k * ld + r0]; }
        double s[NS_FIT];
#pragma unroll
        for (int k = 0; k < NS_FIT; ++k) s[k] = 0.0;
        for (int64_t i = tid; i < n; i += blockDim.x) {
            const double a[3] = {p1[r0 + i], p1[ld + r0 + i], p1[2 * ld + r0 + i]};
            const double c[3] = {p2[r0 + i], p2[ld + r0 + i], p2[2 * ld + r0 + i]};
            fit_accumulate(s, w ? w[r0 + i] : 1.0, a, c, piv1, piv2);
        }
        block_sum<NS_FIT>(s, red);
        // rank guard (estimateTransform.m:11-14): every thread holds the reduced sums, so the probes are block-uniform
        RankProbe r1, r2;
        rank_probe(s + 17, n, r1);
        rank_probe(s + 23, n, r2);
        int rank1 = r1.sure, rank2 = r2.sure;
        if ((r1.sure < 3 && r1.nunc > 0) || (r2.sure < 2 && r2.nunc > 0)) {       // uncertain directions decide: measure them on the data
            double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int64_t i = tid; i < n; i += blockDim.x) {
                const double a[3] = {p1[r0 + i], p1[ld + r0 + i], p1[2 * ld + r0 + i]};
                const double c[3] = {p2[r0 + i], p2[ld + r0 + i], p2[2 * ld + r0 + i]};
                rank_refine_accumulate(r1, a, acc);
                rank_refine_accumulate(r2, c, acc + 3);
            }
            block_sum<6>(acc, red);
            rank1 = rank_final(r1, acc);
            rank2 = rank_final(r2, acc + 3);
        }
        if (tid == 0) {
            st = (s[0] > 0.0 && rank1 >= 3 && rank2 >= 2) ? 0 : 1;
            if (st == 0) kabsch_from_fit_sums(s, piv1, piv2, reflection_fix != 0, T);
        }
    }
    if (tid == 0) {
        status[b] = st;
        for (int k = 0; k < 16; ++k) T_rm[b * 16 + k] = st == 0 ? T[k] : nan("");
    }
}

// ---- ransac scoring: one warp per hypothesis --------------------------------------------------------
struct RansacArgs {
    const double* p1; const double* p2; int64_t P; int64_t ld;
    const int32_t* triplets; int64_t nhyp;
    double thDist; double thInlr; int refine; int reflection_fix;
    int32_t* cnt; int32_t* cnt_ref; double* T_rm;      // [nhyp], [nhyp], [nhyp][16] (NaN where no TForm kept)
    // batch of windows (pcreg_ransac_batch): hypothesis h belongs to window h / hyp_per_win, whose pairs are rows
    // win_off[w] .. win_off[w+1]-1; triplets are relative to the window; thInlr = round(ratio * P_w) per window
    const int64_t* win_off; int64_t hyp_per_win; double ratio;
};
// the pairs one hypothesis is scored on
struct RansacView { const double* p1; const double* p2; int64_t P; int64_t ld; double thDist; double thInlr; };

__device__ __forceinline__ void load_pt(const double* __restrict__ p, int64_t ld, int64_t i, double* o) {
    o[0] = p[i]; o[1] = p[ld + i]; o[2] = p[2 * ld + i];
}

__device__ __forceinline__ int warp_count_inliers(const RansacView& a, const double* T, int lane) {
    int c = 0;
    for (int64_t i = lane; i < a.P; i += 32) {
        double x[3], y[3], qx, qy, qz;
        load_pt(a.p1, a.ld, i, x);
        load_pt(a.p2, a.ld, i, y);
        quick_tf(T, y[0], y[1], y[2], qx, qy, qz);
        c += dist2_exact(x[0], x[1], x[2], qx, qy, qz) < a.thDist ? 1 : 0;     // calcDists + ransac.m:49
    }
    return warp_sum_i(c);
}

__global__ void __launch_bounds__(256) k_ransac_score(const __grid_constant__ RansacArgs g) {
    const int lane = threadIdx.x & 31;
    const int64_t h = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (h >= g.nhyp) return;
    int64_t row0 = 0, np_win = g.P;
    double th_win = g.thInlr;
    if (g.win_off) {
        const int64_t w = h / g.hyp_per_win;
        row0 = g.win_off[w];
        np_win = g.win_off[w + 1] - row0;
        th_win = floor(g.ratio * (double)np_win + 0.5);                         // MATLAB round, ransac.m:28
    }
    // (the view is built once from values: patching the pointers of a copy of the __grid_constant__ struct in place was
    // miscompiled by nvcc 12.9 -- the pointer increments were dropped)
    const RansacView a{g.p1 + row0, g.p2 + row0, np_win, g.ld, g.thDist, th_win};
    const double nanv = nan("");
    double T1[16];
    int st = 1;
    int32_t s3[3];
    for (int k = 0; k < 3; ++k) s3[k] = g.triplets[h * 3 + k];
    const bool valid = a.P >= 3 && s3[0] >= 0 && s3[1] >= 0 && s3[2] >= 0 && s3[0] < a.P && s3[1] < a.P && s3[2] < a.P;
    if (valid) {
        double P1[9], P2[9];
        for (int i = 0; i < 3; ++i) { load_pt(a.p1, a.ld, s3[i], P1 + 3 * i); load_pt(a.p2, a.ld, s3[i], P2 + 3 * i); }
        st = kabsch3(P1, P2, g.reflection_fix != 0, T1);                        // ransac.m:45
    }
    int cnt = 0, cnt_ref = 0;
    bool keep = false;
    double Tk[16];
    if (st == 0) {
        cnt = warp_count_inliers(a, T1, lane);                                  // ransac.m:48-50
        if ((double)cnt >= a.thInlr) {                                          // :53
            if (g.refine) {
                // refit on the inliers (:55): pivots = first sample point of each set
                double piv1[3], piv2[3];
                load_pt(a.p1, a.ld, s3[0], piv1);
                load_pt(a.p2, a.ld, s3[0], piv2);
                double T2[16];
                int st2 = 1;
                if (cnt == 3) {
                    // exactly three inliers: the reference's 3-point branch applies to the refit too
                    int64_t id[3] = {0, 0, 0};
                    int found = 0;
                    for (int64_t i = 0; i < a.P && found < 3; ++i) {             // every lane scans (tiny, rare)
                        double x[3], y[3], qx, qy, qz;
                        load_pt(a.p1, a.ld, i, x);
                        load_pt(a.p2, a.ld, i, y);
                        quick_tf(T1, y[0], y[1], y[2], qx, qy, qz);
                        if (dist2_exact(x[0], x[1], x[2], qx, qy, qz) < a.thDist) id[found++] = i;
                    }
                    double P1[9], P2[9];
                    for (int i = 0; i < 3; ++i) { load_pt(a.p1, a.ld, id[i], P1 + 3 * i); load_pt(a.p2, a.ld, id[i], P2 + 3 * i); }
                    st2 = kabsch3(P1, P2, g.reflection_fix != 0, T2);
                } else {
                    double s[NS_FIT];
#pragma unroll
                    for (int k = 0; k < NS_FIT; ++k) s[k] = 0.0;
                    for (int64_t i = lane; i < a.P; i += 32) {
                        double x[3], y[3], qx, qy, qz;
                        load_pt(a.p1, a.ld, i, x);
                        load_pt(a.p2, a.ld, i, y);
                        quick_tf(T1, y[0], y[1], y[2], qx, qy, qz);
                        if (dist2_exact(x[0], x[1], x[2], qx, qy, qz) < a.thDist) fit_accumulate(s, 1.0, x, y, piv1, piv2);
                    }
#pragma unroll
                    for (int k = 0; k < NS_FIT; ++k) s[k] = warp_sum(s[k]);
                    RankProbe r1, r2;                                            // rank guard of the refit (estimateTransform.m:11-14)
                    rank_probe(s + 17, cnt, r1);
                    rank_probe(s + 23, cnt, r2);
                    int rank1 = r1.sure, rank2 = r2.sure;
                    if ((r1.sure < 3 && r1.nunc > 0) || (r2.sure < 2 && r2.nunc > 0)) {
                        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                        for (int64_t i = lane; i < a.P; i += 32) {
                            double x[3], y[3], qx, qy, qz;
                            load_pt(a.p1, a.ld, i, x);
                            load_pt(a.p2, a.ld, i, y);
                            quick_tf(T1, y[0], y[1], y[2], qx, qy, qz);
                            if (dist2_exact(x[0], x[1], x[2], qx, qy, qz) < a.thDist) { rank_refine_accumulate(r1, x, acc); rank_refine_accumulate(r2, y, acc + 3); }
                        }
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
                        rank1 = rank_final(r1, acc);
                        rank2 = rank_final(r2, acc + 3);
                    }
                    st2 = (rank1 >= 3 && rank2 >= 2) ? 0 : 1;
                    if (st2 == 0) kabsch_from_fit_sums(s, piv1, piv2, g.reflection_fix != 0, T2);
                }
                if (st2 == 0) {
                    cnt_ref = warp_count_inliers(a, T2, lane);                  // :56-58
                    if ((double)cnt_ref >= a.thInlr) {                          // :59-61
                        keep = true;
                        for (int k = 0; k < 16; ++k) Tk[k] = T2[k];
                    }
                }
            } else {
                keep = true;                                                    // :63
                for (int k = 0; k < 16; ++k) Tk[k] = T1[k];
            }
        }
    }
    if (lane == 0) {
        g.cnt[h] = cnt;
        g.cnt_ref[h] = cnt_ref;
    }
    if (lane < 16) g.T_rm[h * 16 + lane] = keep ? Tk[lane] : nanv;
}

// inlier flags of one transform (final calcDists, ransac.m:78,92)
__global__ void k_inlier_flags(const double* __restrict__ p1, const double* __restrict__ p2, int64_t P, int64_t ld,
                               const double* __restrict__ T_rm, double thDist, uint8_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double T[16];
    for (int k = 0; k < 16; ++k) T[k] = T_rm[k];
    double qx, qy, qz;
    quick_tf(T, p2[i], p2[ld + i], p2[2 * ld + i], qx, qy, qz);
    flags[i] = dist2_exact(p1[i], p1[ld + i], p1[2 * ld + i], qx, qy, qz) < thDist ? 1 : 0;
}

// ---- documented counter-based sampler for ransac.m:42-43 (randperm(ptNum) -> first 3) ------------------
// MATLAB's global RNG stream cannot be matched, so the drop-in defines its own reproducible sampler:
//   u_k = splitmix64(seed + 0x9E3779B97F4A7C15 * (3*h + k + 1)),  k = 0,1,2
//   i0 = u_0 mod P;  i1 = u_1 mod (P-1), skipping i0;  i2 = u_2 mod (P-2), skipping both
// i.e. a uniformly random ORDERED 3-subset without replacement, like the first three entries of randperm.
// oracle/primitives.py:ransac_triplets restates it in numpy for the parity tests.
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void k_gen_triplets(unsigned long long seed, int64_t nhyp, int64_t P, int32_t* __restrict__ tri) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nhyp) return;
    const unsigned long long u0 = splitmix64(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(3 * h + 1));
    const unsigned long long u1 = splitmix64(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(3 * h + 2));
    const unsigned long long u2 = splitmix64(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(3 * h + 3));
    const long long i0 = (long long)(u0 % (unsigned long long)P);
    long long i1 = (long long)(u1 % (unsigned long long)(P - 1));
    if (i1 >= i0) ++i1;
    long long i2 = (long long)(u2 % (unsigned long long)(P - 2));
    const long long lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
    if (i2 >= lo) ++i2;
    if (i2 >= hi) ++i2;
    tri[3 * h + 0] = (int32_t)i0; tri[3 * h + 1] = (int32_t)i1; tri[3 * h + 2] = (int32_t)i2;
}

// ---- batch of windows (pcreg_ransac_batch) ---------------------------------------------------------------
// The same sampler per window: window w draws from seeds[w] over its own P_w pairs (indices relative to the window).
__device__ __forceinline__ void draw_triplet(unsigned long long seed, int64_t h, int64_t P, int32_t* t) {
    const unsigned long long u0 = splitmix64(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(3 * h + 1));
    const unsigned long long u1 = splitmix64(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(3 * h + 2));
    const unsigned long long u2 = splitmix64(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(3 * h + 3));
    const long long i0 = (long long)(u0 % (unsigned long long)P);
    long long i1 = (long long)(u1 % (unsigned long long)(P - 1));
    if (i1 >= i0) ++i1;
    long long i2 = (long long)(u2 % (unsigned long long)(P - 2));
    const long long lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
    if (i2 >= lo) ++i2;
    if (i2 >= hi) ++i2;
    t[0] = (int32_t)i0; t[1] = (int32_t)i1; t[2] = (int32_t)i2;
}
__global__ void k_gen_triplets_batch(const unsigned long long* __restrict__ seeds, const int64_t* __restrict__ win_off,
                                     int64_t nwin, int64_t iter_num, int32_t* __restrict__ tri) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nwin * iter_num) return;
    const int64_t w = g / iter_num, h = g - w * iter_num, P = win_off[w + 1] - win_off[w];
    int32_t t[3] = {-1, -1, -1};                        // fewer than 3 pairs: no sample (the window reports [])
    if (P >= 3) draw_triplet(seeds[w], h, P, t);
    tri[3 * g + 0] = t[0]; tri[3 * g + 1] = t[1]; tri[3 * g + 2] = t[2];
}

// One block per window: first arg-max of the deciding counts (ransac.m:69-73), numSuccess (ransac.m:94-98), the
// winner's transform (row-major -> the ABI's column-major) and the window's status (1 = ransac.m:75-89 returns []).
__global__ void __launch_bounds__(256) k_ransac_pick(const int32_t* __restrict__ dec, const double* __restrict__ T_rm,
                                                     const int64_t* __restrict__ win_off, int64_t iter_num, double ratio,
                                                     double* __restrict__ T_best_cm, int64_t* __restrict__ best_hyp,
                                                     int64_t* __restrict__ max_inl, int64_t* __restrict__ n_succ,
                                                     int32_t* __restrict__ status) {
    __shared__ int s_cnt[256];
    __shared__ long long s_idx[256];
    __shared__ long long s_succ[256];
    const int64_t w = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t P = win_off[w + 1] - win_off[w];
    const double thInlr = floor(ratio * (double)P + 0.5);
    const int32_t* d = dec + w * iter_num;
    int bc = -1; long long bi = 0, ns = 0;
    for (int64_t h = tid; h < iter_num; h += 256) {      // ascending h per thread: '>' keeps the first maximum
        const int c = d[h];
        if (c > bc) { bc = c; bi = h; }
        ns += ((double)c >= thInlr) ? 1 : 0;
    }
    s_cnt[tid] = bc; s_idx[tid] = bi; s_succ[tid] = ns;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            const int c2 = s_cnt[tid + o]; const long long i2 = s_idx[tid + o];
            if (c2 > s_cnt[tid] || (c2 == s_cnt[tid] && c2 >= 0 && i2 < s_idx[tid])) { s_cnt[tid] = c2; s_idx[tid] = i2; }
            s_succ[tid] += s_succ[tid + o];
        }
        __syncthreads();
    }
    const long long b = s_idx[0];
    const double* Tb = T_rm + (w * iter_num + b) * 16;
    const bool ok = (Tb[0] == Tb[0]);                    // NaN: no TForm kept at the arg-max
    if (tid < 16) {
        const int r = tid >> 2, c = tid & 3;
        T_best_cm[w * 16 + c * 4 + r] = ok ? Tb[tid] : nan("");
    }
    if (tid == 0) {
        status[w] = ok ? 0 : 1;
        best_hyp[w] = ok ? b : -1;
        max_inl[w] = ok ? (int64_t)s_cnt[0] : 0;
        n_succ[w] = ok ? s_succ[0] : 0;
    }
}

// final calcDists of every window with its winner (ransac.m:78,92): one thread per pair, window by binary search
__global__ void k_inlier_flags_batch(const double* __restrict__ p1, const double* __restrict__ p2, int64_t ntotal, int64_t ld,
                                     const int64_t* __restrict__ win_off, int64_t nwin, const double* __restrict__ T_rm,
                                     const int64_t* __restrict__ best_hyp, int64_t iter_num, double thDist,
                                     uint8_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntotal) return;
    int64_t lo = 0, hi = nwin;                           // largest w with win_off[w] <= i (empty windows skipped by '<=')
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (win_off[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t b = best_hyp[lo];
    uint8_t f = 0;
    if (b >= 0) {
        const double* T = T_rm + (lo * iter_num + b) * 16;
        double qx, qy, qz;
        quick_tf(T, p2[i], p2[ld + i], p2[2 * ld + i], qx, qy, qz);
        f = dist2_exact(p1[i], p1[ld + i], p1[2 * ld + i], qx, qy, qz) < thDist ? 1 : 0;
    }
    flags[i] = f;
}

}  // namespace pcreg

using namespace pcreg;

extern "C" {

int pcreg_kabsch_batch(const double* p1, const double* p2, const double* w, int64_t ld, const int64_t* offsets,
                       int64_t nbatch, int reflection_fix, double* T16, int32_t* status) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(p1 && p2 && offsets && T16 && status, "pcreg_kabsch_batch: null pointer");
    PCREG_REQUIRE(nbatch >= 1, "pcreg_kabsch_batch: nbatch must be >= 1");
    const int64_t ntotal = offsets[nbatch];
    PCREG_REQUIRE(offsets[0] == 0 && ntotal >= 0 && ld >= ntotal, "pcreg_kabsch_batch: bad offsets / ld");
    for (int64_t b = 0; b < nbatch; ++b) PCREG_REQUIRE(offsets[b + 1] >= offsets[b], "pcreg_kabsch_batch: offsets must be non-decreasing");
    PCREG_CUDA(cudaSetDevice(ctx().device));
    cudaStream_t st = 0;
    const size_t nel = (size_t)std::max<int64_t>(ntotal, 1);
    DevBuf<double> d1(nel * 3), d2(nel * 3), dw(w ? nel : 0), dT((size_t)nbatch * 16), dTc((size_t)nbatch * 16);
    DevBuf<int64_t> doff((size_t)nbatch + 1);
    DevBuf<int32_t> dst((size_t)nbatch);
    for (int a = 0; a < 3; ++a) {
        PCREG_CUDA(cudaMemcpyAsync(d1.p + a * nel, p1 + a * ld, (size_t)ntotal * 8, cudaMemcpyHostToDevice, st));
        PCREG_CUDA(cudaMemcpyAsync(d2.p + a * nel, p2 + a * ld, (size_t)ntotal * 8, cudaMemcpyHostToDevice, st));
    }
    if (w) PCREG_CUDA(cudaMemcpyAsync(dw.p, w, (size_t)ntotal * 8, cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemcpyAsync(doff.p, offsets, ((size_t)nbatch + 1) * 8, cudaMemcpyHostToDevice, st));
    k_kabsch_batch<<<(unsigned)nbatch, 128, 0, st>>>(d1.p, d2.p, w ? dw.p : nullptr, (int64_t)nel, doff.p, reflection_fix, dT.p, dst.p);
    PCREG_LAUNCHED();
    transpose16_launch(dT.p, dTc.p, nbatch, st);
    PCREG_CUDA(cudaMemcpyAsync(T16, dTc.p, dTc.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(status, dst.p, dst.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
    return PCREG_OK;
    PCREG_API_END
}

// shared body of pcreg_ransac_score / pcreg_ransac_run: triplets already on the device
static int ransac_core(const double* p1, const double* p2, int64_t P, int64_t ld, const int32_t* d_tri, int64_t nhyp,
                       const pcreg_ransac_opts* opts, double* T16_best, int32_t* inl_idx, int64_t* n_inl, int64_t* n_succ,
                       int64_t* max_inl, int64_t* best_hyp, int32_t* inl_counts, int32_t* inl_counts_refined, double* T16_all,
                       cudaStream_t st) {
    DevBuf<double> d1((size_t)P * 3), d2((size_t)P * 3), dT((size_t)nhyp * 16);
    DevBuf<int32_t> dcnt((size_t)nhyp), dcntr((size_t)nhyp);
    for (int a = 0; a < 3; ++a) {
        PCREG_CUDA(cudaMemcpyAsync(d1.p + a * P, p1 + a * ld, (size_t)P * 8, cudaMemcpyHostToDevice, st));
        PCREG_CUDA(cudaMemcpyAsync(d2.p + a * P, p2 + a * ld, (size_t)P * 8, cudaMemcpyHostToDevice, st));
    }
    RansacArgs a{};
    a.p1 = d1.p; a.p2 = d2.p; a.P = P; a.ld = P; a.triplets = d_tri; a.nhyp = nhyp;
    a.thDist = opts->thDist;
    a.thInlr = floor(opts->thInlrRatio * (double)P + 0.5);            // MATLAB round, ransac.m:28
    a.refine = opts->refine; a.reflection_fix = opts->reflection_fix;
    a.cnt = dcnt.p; a.cnt_ref = dcntr.p; a.T_rm = dT.p;
    const int64_t blocks = (nhyp * 32 + 255) / 256;
    PCREG_REQUIRE(blocks < 2147483647LL, "ransac: too many hypotheses");
    k_ransac_score<<<(unsigned)blocks, 256, 0, st>>>(a);
    PCREG_LAUNCHED();
    std::vector<int32_t> hc((size_t)nhyp), hcr((size_t)nhyp);
    PCREG_CUDA(cudaMemcpyAsync(hc.data(), dcnt.p, dcnt.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(hcr.data(), dcntr.p, dcntr.bytes(), cudaMemcpyDeviceToHost, st));
    DevBuf<double> dTc(T16_all ? (size_t)nhyp * 16 : 0);
    if (T16_all) {
        transpose16_launch(dT.p, dTc.p, nhyp, st);
        PCREG_CUDA(cudaMemcpyAsync(T16_all, dTc.p, dTc.bytes(), cudaMemcpyDeviceToHost, st));
    }
    PCREG_CUDA(cudaStreamSynchronize(st));
    if (inl_counts) std::copy(hc.begin(), hc.end(), inl_counts);
    if (inl_counts_refined) std::copy(hcr.begin(), hcr.end(), inl_counts_refined);
    // first arg-max of the counts that decide (ransac.m:69-73), host side: a scan of nhyp int32
    const std::vector<int32_t>& dec = opts->refine ? hcr : hc;
    int64_t bi = 0;
    for (int64_t h = 1; h < nhyp; ++h) if (dec[h] > dec[bi]) bi = h;
    double Tb[16];
    PCREG_CUDA(cudaMemcpy(Tb, dT.p + bi * 16, sizeof Tb, cudaMemcpyDeviceToHost));
    *n_inl = 0; *n_succ = 0; *max_inl = 0; *best_hyp = -1;
    if (!(Tb[0] == Tb[0])) {                                           // no TForm kept there: ransac.m:75-89
        for (int k = 0; k < 16; ++k) T16_best[k] = nan("");
        return PCREG_DEGENERATE;
    }
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T16_best[c * 4 + r] = Tb[r * 4 + c];
    DevBuf<uint8_t> dflags((size_t)P);
    k_inlier_flags<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(d1.p, d2.p, P, P, dT.p + bi * 16, opts->thDist, dflags.p);
    PCREG_LAUNCHED();
    std::vector<uint8_t> hf((size_t)P);
    PCREG_CUDA(cudaMemcpy(hf.data(), dflags.p, (size_t)P, cudaMemcpyDeviceToHost));
    int64_t ni = 0;
    for (int64_t i = 0; i < P; ++i) if (hf[i]) inl_idx[ni++] = (int32_t)i;
    *n_inl = ni;
    int64_t ns = 0;
    for (int64_t h = 0; h < nhyp; ++h) ns += ((double)dec[h] >= a.thInlr) ? 1 : 0;       // ransac.m:94-98
    *n_succ = ns; *max_inl = dec[bi]; *best_hyp = bi;
    return PCREG_OK;
}

int pcreg_ransac_score(const double* p1, const double* p2, int64_t P, int64_t ld, const int32_t* triplets, int64_t nhyp,
                       const pcreg_ransac_opts* opts, double* T16_best, int32_t* inl_idx, int64_t* n_inl, int64_t* n_succ,
                       int64_t* max_inl, int64_t* best_hyp, int32_t* inl_counts, int32_t* inl_counts_refined, double* T16_all) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(p1 && p2 && triplets && opts && T16_best && inl_idx && n_inl && n_succ && max_inl && best_hyp,
                  "pcreg_ransac_score: null pointer");
    PCREG_REQUIRE(P >= 3 && ld >= P && nhyp >= 1, "pcreg_ransac_score: need P >= 3, ld >= P, nhyp >= 1");
    PCREG_CUDA(cudaSetDevice(ctx().device));
    cudaStream_t st = 0;
    DevBuf<int32_t> dtri((size_t)nhyp * 3);
    PCREG_CUDA(cudaMemcpyAsync(dtri.p, triplets, dtri.bytes(), cudaMemcpyHostToDevice, st));
    return ransac_core(p1, p2, P, ld, dtri.p, nhyp, opts, T16_best, inl_idx, n_inl, n_succ, max_inl, best_hyp, inl_counts,
                       inl_counts_refined, T16_all, st);
    PCREG_API_END
}

int pcreg_ransac_run(const double* p1, const double* p2, int64_t P, int64_t ld, int64_t iter_num, uint64_t seed,
                     const pcreg_ransac_opts* opts, double* T16_best, int32_t* inl_idx, int64_t* n_inl, int64_t* n_succ,
                     int64_t* max_inl, int64_t* best_hyp, int32_t* triplets_out) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(p1 && p2 && opts && T16_best && inl_idx && n_inl && n_succ && max_inl && best_hyp, "pcreg_ransac_run: null pointer");
    PCREG_REQUIRE(P >= 3 && ld >= P && iter_num >= 1, "pcreg_ransac_run: need P >= 3, ld >= P, iter_num >= 1");
    PCREG_CUDA(cudaSetDevice(ctx().device));
    cudaStream_t st = 0;
    DevBuf<int32_t> dtri((size_t)iter_num * 3);
    k_gen_triplets<<<(unsigned)((iter_num + 255) / 256), 256, 0, st>>>((unsigned long long)seed, iter_num, P, dtri.p);
    PCREG_LAUNCHED();
    if (triplets_out) PCREG_CUDA(cudaMemcpyAsync(triplets_out, dtri.p, dtri.bytes(), cudaMemcpyDeviceToHost, st));
    return ransac_core(p1, p2, P, ld, dtri.p, iter_num, opts, T16_best, inl_idx, n_inl, n_succ, max_inl, best_hyp, nullptr, nullptr,
                       nullptr, st);
    PCREG_API_END
}

}  // extern "C" (the slice helper below is internal)

// The windows [0, nwin) described by `offsets` (offsets[0] = 0) on the calling thread's device slot.
static void ransac_batch_slice(const double* p1, const double* p2, int64_t ld, const int64_t* offsets, int64_t nwin, int64_t iter_num,
                               const int32_t* triplets, const uint64_t* seeds, const pcreg_ransac_opts* opts, double* T16,
                               int32_t* inl_idx, int64_t* n_inl, int64_t* n_succ, int64_t* max_inl, int64_t* best_hyp, int32_t* status) {
    const int64_t ntotal = offsets[nwin];
    const int64_t nhyp = nwin * iter_num;
    PCREG_CUDA(cudaSetDevice(ctx().device));
    cudaStream_t st = 0;
    const size_t nel = (size_t)std::max<int64_t>(ntotal, 1);
    DevBuf<double> d1(nel * 3), d2(nel * 3), dT((size_t)nhyp * 16), dTb((size_t)nwin * 16);
    DevBuf<int64_t> doff((size_t)nwin + 1), dbest((size_t)nwin), dmax((size_t)nwin), dsucc((size_t)nwin);
    DevBuf<int32_t> dtri((size_t)nhyp * 3), dcnt((size_t)nhyp), dcntr((size_t)nhyp), dstat((size_t)nwin);
    DevBuf<unsigned long long> dseed(triplets ? 0 : (size_t)nwin);
    DevBuf<uint8_t> dflags(nel);
    for (int a = 0; a < 3; ++a) {
        PCREG_CUDA(cudaMemcpyAsync(d1.p + a * nel, p1 + a * ld, (size_t)ntotal * 8, cudaMemcpyHostToDevice, st));
        PCREG_CUDA(cudaMemcpyAsync(d2.p + a * nel, p2 + a * ld, (size_t)ntotal * 8, cudaMemcpyHostToDevice, st));
    }
    PCREG_CUDA(cudaMemcpyAsync(doff.p, offsets, ((size_t)nwin + 1) * 8, cudaMemcpyHostToDevice, st));
    if (triplets) {
        PCREG_CUDA(cudaMemcpyAsync(dtri.p, triplets, dtri.bytes(), cudaMemcpyHostToDevice, st));
    } else {
        PCREG_CUDA(cudaMemcpyAsync(dseed.p, seeds, (size_t)nwin * 8, cudaMemcpyHostToDevice, st));
        k_gen_triplets_batch<<<(unsigned)((nhyp + 255) / 256), 256, 0, st>>>(dseed.p, doff.p, nwin, iter_num, dtri.p);
        PCREG_LAUNCHED();
    }
    RansacArgs a{};
    a.p1 = d1.p; a.p2 = d2.p; a.P = 0; a.ld = (int64_t)nel; a.triplets = dtri.p; a.nhyp = nhyp;
    a.thDist = opts->thDist; a.thInlr = 0.0; a.refine = opts->refine; a.reflection_fix = opts->reflection_fix;
    a.cnt = dcnt.p; a.cnt_ref = dcntr.p; a.T_rm = dT.p;
    a.win_off = doff.p; a.hyp_per_win = iter_num; a.ratio = opts->thInlrRatio;
    k_ransac_score<<<(unsigned)((nhyp * 32 + 255) / 256), 256, 0, st>>>(a);
    PCREG_LAUNCHED();
    k_ransac_pick<<<(unsigned)nwin, 256, 0, st>>>(opts->refine ? dcntr.p : dcnt.p, dT.p, doff.p, iter_num, opts->thInlrRatio, dTb.p,
                                                  dbest.p, dmax.p, dsucc.p, dstat.p);
    PCREG_LAUNCHED();
    if (ntotal > 0) {
        k_inlier_flags_batch<<<(unsigned)((ntotal + 255) / 256), 256, 0, st>>>(d1.p, d2.p, ntotal, (int64_t)nel, doff.p, nwin, dT.p,
                                                                                dbest.p, iter_num, opts->thDist, dflags.p);
        PCREG_LAUNCHED();
    }
    std::vector<uint8_t> hf((size_t)ntotal);
    PCREG_CUDA(cudaMemcpyAsync(T16, dTb.p, dTb.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(best_hyp, dbest.p, dbest.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(max_inl, dmax.p, dmax.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(n_succ, dsucc.p, dsucc.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(status, dstat.p, dstat.bytes(), cudaMemcpyDeviceToHost, st));
    if (ntotal > 0) PCREG_CUDA(cudaMemcpyAsync(hf.data(), dflags.p, (size_t)ntotal, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
    for (int64_t w = 0; w < nwin; ++w) {                   // inlierIdx = find(dist < thDist) per window, relative, ascending
        int64_t ni = 0;
        if (status[w] == 0)
            for (int64_t i = offsets[w]; i < offsets[w + 1]; ++i)
                if (hf[(size_t)i]) inl_idx[offsets[w] + ni++] = (int32_t)(i - offsets[w]);
        n_inl[w] = ni;
    }
}

extern "C" {

int pcreg_ransac_batch(const double* p1, const double* p2, int64_t ld, const int64_t* offsets, int64_t nwin, int64_t iter_num,
                       const int32_t* triplets, const uint64_t* seeds, const pcreg_ransac_opts* opts, double* T16,
                       int32_t* inl_idx, int64_t* n_inl, int64_t* n_succ, int64_t* max_inl, int64_t* best_hyp, int32_t* status) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(p1 && p2 && offsets && opts && T16 && inl_idx && n_inl && n_succ && max_inl && best_hyp && status,
                  "pcreg_ransac_batch: null pointer");
    PCREG_REQUIRE(triplets || seeds, "pcreg_ransac_batch: need triplets or seeds");
    PCREG_REQUIRE(nwin >= 1 && iter_num >= 1, "pcreg_ransac_batch: need nwin >= 1 and iter_num >= 1");
    const int64_t ntotal = offsets[nwin];
    PCREG_REQUIRE(offsets[0] == 0 && ntotal >= 0 && ld >= ntotal, "pcreg_ransac_batch: bad offsets / ld");
    for (int64_t w = 0; w < nwin; ++w) PCREG_REQUIRE(offsets[w + 1] >= offsets[w], "pcreg_ransac_batch: offsets must be non-decreasing");
    const int64_t nhyp = nwin * iter_num;
    PCREG_REQUIRE(nhyp < ((int64_t)1 << 26), "pcreg_ransac_batch: more than 2^26 hypotheses in one call");
    // Windows are independent (the reference runs one ransac per window under parfor, slideMatchingWindow_v2.m:178): contiguous
    // shares of windows, one host thread per selected device; each window's result equals the single-window call, so the
    // split changes no bit.
    const int nd = (int)std::min<int64_t>(num_slots(), nwin);
    const int64_t per = (nwin + nd - 1) / nd;
    std::vector<int> rc((size_t)nd, PCREG_OK);
    std::vector<std::thread> workers;
    auto run = [&](int k) {
        const int64_t w0 = k * per, wn = std::min(per, nwin - w0);
        if (wn <= 0) return;
        std::vector<int64_t> off((size_t)wn + 1);
        for (int64_t w = 0; w <= wn; ++w) off[(size_t)w] = offsets[w0 + w] - offsets[w0];
        const int64_t r0 = offsets[w0];
        use_slot(k);
        ransac_batch_slice(p1 + r0, p2 + r0, ld, off.data(), wn, iter_num, triplets ? triplets + w0 * iter_num * 3 : nullptr,
                           seeds ? seeds + w0 : nullptr, opts, T16 + w0 * 16, inl_idx + r0, n_inl + w0, n_succ + w0, max_inl + w0,
                           best_hyp + w0, status + w0);
    };
    for (int k = 1; k < nd; ++k) workers.emplace_back([&, k]() { rc[k] = guarded([&]() { run(k); }); });
    rc[0] = guarded([&]() { run(0); });
    for (auto& w : workers) w.join();
    use_slot(0);
    for (int k = 0; k < nd; ++k) if (rc[k] != PCREG_OK) return rc[k];
    return PCREG_OK;
    PCREG_API_END
}

}  // extern "C"
