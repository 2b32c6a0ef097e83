// nn_brute.cu -- tiled brute-force nearest neighbour for sm_100a: FP32 scan on the FMA pipe,
// exact FP64 decision.
//
// What it replaces: the NN correspondence step of the ICP loop BASELINE.json's north_star names.
// The reference has no such code; semantics are MATLAB knnsearch(X,Y,'K',1):
// Euclidean, FP64, ties -> smallest index (the reference's only literal 1-NN loop is
// ColorCodeModel.m:15-18).  Queries are q = quickTF(src_i, T_h) (quickTF.m:5-7), computed on the fly,
// never materialised.
//
// Algorithm
//   scan value   d'(j) = |m_j|^2 - 2 q.m_j   (3 FFMA per pair, |m|^2 in the float4 w lane; all
//                coordinates relative to the model pivot), i.e. d^2 - |q|^2 in FP32.
//   fast path    per group of G model points a thread keeps only the group minimum of d' for each of
//                its Q register-resident queries (FMNMX3), no index tracking.
//   slow path    a group whose minimum is <= (best d' so far + 2E) is re-scanned; every point inside
//                that band is evaluated EXACTLY in FP64 with the oracle's formula and operation order
//                on the original coordinates, and competes on (d2, original index).  E bounds the FP32
//                error of d' against the true FP64 distance (derivation in DESIGN.md), so the true
//                nearest neighbour -- and every exact tie of it -- is always inside the band: the
//                returned index and d2 are the FP64 brute-force answers, bit for bit.
//   bound        the band needs a good initial "best so far": the previous ICP iteration's
//                correspondence when there is one, else a pure-min pass over a 1/8 sub-sample (the
//                scan order is a random permutation, so the head of the array is a random sample).
//   tiles        model tiles of 1024 float4 (16 KB) are staged in shared memory by the TMA engine
//                (cp.async.bulk + mbarrier, double-buffered); all 256 threads read them as broadcast
//                LDS.128.  Few queries (C2: 10 k) are spread over the SMs by splitting the MODEL across
//                blockIdx.y; partial (d2, idx) winners are merged by a small kernel.
#include <float.h>
#include <math.h>
#include <algorithm>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"

namespace pcreg {

constexpr int BRUTE_THREADS = 256;
constexpr int BRUTE_G = 16;                 // group size of the fast path
constexpr float BRUTE_ERR_COEF = 7.2e-7f;   // 12 * 2^-24: E = coef * (|q32| + max|m32|)^2

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

struct BruteArgs {
    const float4* m4; const int32_t* perm; const ModelPointD* md;
    const double* sx; const double* sy; const double* sz; int64_t ns;
    const double* T; int64_t nq;
    double px, py, pz; float max_norm;
    int tile0;                  // first tile of the scanned range
    int ntiles;                 // tiles in the scanned range
    int tiles_per_split;
    const int32_t* prev;        // [nq] or null
    const float* bound_in;      // [nbound][nq] or null
    int nbound;
    float* pmin;                // PURE_MIN output [gridDim.y][nq]
    double* pd2; int32_t* pidx; // exact output [gridDim.y][nq]
};

struct SlowState { float thr; float best32; double best64; int32_t bidx; };

// Re-scan one group for one query; exact FP64 evaluation of everything inside the band.
__device__ __noinline__ SlowState nn_slow(const float4* __restrict__ grp, int64_t jbase, float ax, float ay, float az,
                                          float twoE, SlowState s, int64_t g, const BruteArgs& a) {
    const int64_t h = g / a.ns, i = g - h * a.ns;
    double qx, qy, qz;
    quick_tf(a.T + h * 16, a.sx[i], a.sy[i], a.sz[i], qx, qy, qz);
#pragma unroll 1
    for (int j = 0; j < BRUTE_G; ++j) {
        const float4 m = grp[j];
        float d = fmaf(ax, m.x, m.w);
        d = fmaf(ay, m.y, d);
        d = fmaf(az, m.z, d);
        if (d <= s.thr) {
            const int32_t orig = a.perm[jbase + j];
            if (orig >= 0) {
                const ModelPointD p = a.md[orig];
                const double d64 = dist2_exact(p.x, p.y, p.z, qx, qy, qz);
                if (d64 < s.best64 || (d64 == s.best64 && orig < s.bidx)) { s.best64 = d64; s.bidx = orig; }
                if (d < s.best32) { s.best32 = d; s.thr = __fadd_ru(d, twoE); }
            }
        }
    }
    return s;
}

template <int Q, bool PURE_MIN>
__global__ void __launch_bounds__(BRUTE_THREADS, 2) k_nn_brute(const __grid_constant__ BruteArgs a) {
    __shared__ __align__(128) float4 tile[2][BRUTE_TILE];
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x;
    const int64_t qbase = (int64_t)blockIdx.x * (BRUTE_THREADS * Q);
    const int t_begin = a.tile0 + blockIdx.y * a.tiles_per_split;
    int t_end = t_begin + a.tiles_per_split;
    if (t_end > a.tile0 + a.ntiles) t_end = a.tile0 + a.ntiles;
    const int nt = t_end - t_begin;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    constexpr uint32_t TILE_BYTES = BRUTE_TILE * sizeof(float4);
    if (tid == 0 && nt > 0) {
        mbar_expect_tx(&full[0], TILE_BYTES);
        bulk_g2s(&tile[0][0], a.m4 + (int64_t)t_begin * BRUTE_TILE, TILE_BYTES, &full[0]);
    }

    // ---- per-thread query state ----
    float ax[Q], ay[Q], az[Q], thr[Q], best32[Q], twoE[Q];
    double best64[Q];
    int32_t bidx[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const int64_t g = qbase + (int64_t)k * BRUTE_THREADS + tid;
        ax[k] = ay[k] = az[k] = 0.f;
        thr[k] = -FLT_MAX; best32[k] = FLT_MAX; twoE[k] = 0.f;
        best64[k] = INFINITY; bidx[k] = -1;
        if (g < a.nq) {
            const int64_t h = g / a.ns, i = g - h * a.ns;
            double qx, qy, qz;
            quick_tf(a.T + h * 16, a.sx[i], a.sy[i], a.sz[i], qx, qy, qz);
            const float fx = __double2float_rn(qx - a.px), fy = __double2float_rn(qy - a.py), fz = __double2float_rn(qz - a.pz);
            ax[k] = -2.f * fx; ay[k] = -2.f * fy; az[k] = -2.f * fz;
            const float qn = __fsqrt_ru(__fmaf_ru(fx, fx, __fmaf_ru(fy, fy, __fmul_ru(fz, fz))));
            const float s = __fadd_ru(qn, a.max_norm);
            twoE[k] = __fmul_ru(__fmul_ru(2.02f * BRUTE_ERR_COEF, s), s);
            float bound = FLT_MAX;
            if (!PURE_MIN) {
                if (a.prev) {
                    const int32_t p = a.prev[g];
                    if (p >= 0) {
                        const ModelPointD mp = a.md[p];
                        const float mx = __double2float_rn(mp.x - a.px), my = __double2float_rn(mp.y - a.py), mz = __double2float_rn(mp.z - a.pz);
                        const float mw = __double2float_rn((double)mx * (double)mx + (double)my * (double)my + (double)mz * (double)mz);
                        float d = fmaf(ax[k], mx, mw);
                        d = fmaf(ay[k], my, d);
                        d = fmaf(az[k], mz, d);
                        bound = d;
                    }
                } else if (a.bound_in) {
                    for (int b = 0; b < a.nbound; ++b) bound = fminf(bound, a.bound_in[(int64_t)b * a.nq + g]);
                }
                best32[k] = bound;
                thr[k] = (bound < FLT_MAX) ? __fadd_ru(bound, twoE[k]) : FLT_MAX;
            }
        }
    }

    // ---- stream the model tiles ----
#pragma unroll 1
    for (int it = 0; it < nt; ++it) {
        const int cur = it & 1;
        if (tid == 0 && it + 1 < nt) {
            mbar_expect_tx(&full[cur ^ 1], TILE_BYTES);
            bulk_g2s(&tile[cur ^ 1][0], a.m4 + (int64_t)(t_begin + it + 1) * BRUTE_TILE, TILE_BYTES, &full[cur ^ 1]);
        }
        mbar_wait(&full[cur], (uint32_t)((it >> 1) & 1));
        const float4* __restrict__ tp = &tile[cur][0];
#pragma unroll 1
        for (int g0 = 0; g0 < BRUTE_TILE; g0 += BRUTE_G) {
            float gm[Q];
#pragma unroll
            for (int jj = 0; jj < BRUTE_G; jj += 2) {
                const float4 m0 = tp[g0 + jj], m1 = tp[g0 + jj + 1];
#pragma unroll
                for (int k = 0; k < Q; ++k) {
                    float d0 = fmaf(ax[k], m0.x, m0.w);
                    float d1 = fmaf(ax[k], m1.x, m1.w);
                    d0 = fmaf(ay[k], m0.y, d0);
                    d1 = fmaf(ay[k], m1.y, d1);
                    d0 = fmaf(az[k], m0.z, d0);
                    d1 = fmaf(az[k], m1.z, d1);
                    if (PURE_MIN)      best32[k] = fmin3(best32[k], d0, d1);
                    else if (jj == 0)  gm[k] = fminf(d0, d1);
                    else               gm[k] = fmin3(gm[k], d0, d1);
                }
            }
            if (!PURE_MIN) {
                bool hit = false;
#pragma unroll
                for (int k = 0; k < Q; ++k) hit |= (gm[k] <= thr[k]);
                if (hit) {
                    const int64_t jbase = (int64_t)(t_begin + it) * BRUTE_TILE + g0;
#pragma unroll
                    for (int k = 0; k < Q; ++k) {
                        if (gm[k] <= thr[k]) {
                            SlowState s{thr[k], best32[k], best64[k], bidx[k]};
                            s = nn_slow(tp + g0, jbase, ax[k], ay[k], az[k], twoE[k], s,
                                        qbase + (int64_t)k * BRUTE_THREADS + tid, a);
                            thr[k] = s.thr; best32[k] = s.best32; best64[k] = s.best64; bidx[k] = s.bidx;
                        }
                    }
                }
            }
        }
        __syncthreads();        // everyone is done with tile[cur] before it is refilled
    }

#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const int64_t g = qbase + (int64_t)k * BRUTE_THREADS + tid;
        if (g < a.nq) {
            const int64_t o = (int64_t)blockIdx.y * a.nq + g;
            if (PURE_MIN) a.pmin[o] = best32[k];
            else { a.pd2[o] = best64[k]; a.pidx[o] = bidx[k]; }
        }
    }
}

// merge the per-split winners on (d2, original index)
__global__ void k_nn_merge(const double* __restrict__ pd2, const int32_t* __restrict__ pidx, int nsplit, int64_t nq,
                           int32_t* __restrict__ idx, double* __restrict__ d2) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nq) return;
    double bd = INFINITY;
    int32_t bi = -1;
    for (int s = 0; s < nsplit; ++s) {
        const double d = pd2[(int64_t)s * nq + g];
        const int32_t i = pidx[(int64_t)s * nq + g];
        if (i >= 0 && (d < bd || (d == bd && (bi < 0 || i < bi)))) { bd = d; bi = i; }
    }
    idx[g] = bi;
    if (d2) d2[g] = bd;
}

template <int Q>
static void brute_run(const pcreg_model* m, BruteArgs a, const int32_t* d_prev, int32_t* d_idx, double* d_d2,
                      NNScratch& sc, cudaStream_t st) {
    const int64_t nq = a.nq;
    const int ntiles_all = (int)(m->n_pad / BRUTE_TILE);
    const int64_t qblocks = (nq + BRUTE_THREADS * Q - 1) / (BRUTE_THREADS * Q);
    const int target_blocks = ctx().sm_count * 4;
    auto plan_split = [&](int ntiles, int& nsplit, int& tps) {
        int want = (int)std::max<int64_t>(1, (target_blocks + qblocks - 1) / qblocks);
        want = std::min(want, std::max(1, ntiles / 2));
        tps = (ntiles + want - 1) / want;
        nsplit = (ntiles + tps - 1) / tps;
    };
    a.prev = d_prev;
    a.bound_in = nullptr; a.nbound = 0;
    if (!d_prev) {
        // bound pass: pure minimum of d' over the first 1/8 of the (randomly ordered) scan array
        const int nsub = std::max(1, ntiles_all / 8);
        int ns0, tps0;
        plan_split(nsub, ns0, tps0);
        if (sc.pmin.n < (size_t)ns0 * nq) sc.pmin.alloc((size_t)ns0 * nq);
        BruteArgs b = a;
        b.tile0 = 0; b.ntiles = nsub; b.tiles_per_split = tps0; b.pmin = sc.pmin.p;
        dim3 grid((unsigned)qblocks, (unsigned)ns0);
        k_nn_brute<Q, true><<<grid, BRUTE_THREADS, 0, st>>>(b);
        PCREG_LAUNCHED();
        a.bound_in = sc.pmin.p; a.nbound = ns0;
    }
    int nsplit, tps;
    plan_split(ntiles_all, nsplit, tps);
    a.tile0 = 0; a.ntiles = ntiles_all; a.tiles_per_split = tps;
    if (nsplit == 1) {
        a.pd2 = d_d2; a.pidx = d_idx;
        if (!d_d2) { if (sc.pd2.n < (size_t)nq) sc.pd2.alloc((size_t)nq); a.pd2 = sc.pd2.p; }
    } else {
        if (sc.pd2.n < (size_t)nsplit * nq) sc.pd2.alloc((size_t)nsplit * nq);
        if (sc.pidx.n < (size_t)nsplit * nq) sc.pidx.alloc((size_t)nsplit * nq);
        a.pd2 = sc.pd2.p; a.pidx = sc.pidx.p;
    }
    dim3 grid((unsigned)qblocks, (unsigned)nsplit);
    k_nn_brute<Q, false><<<grid, BRUTE_THREADS, 0, st>>>(a);
    PCREG_LAUNCHED();
    if (nsplit > 1) {
        k_nn_merge<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(sc.pd2.p, sc.pidx.p, nsplit, nq, d_idx, d_d2);
        PCREG_LAUNCHED();
    }
}

void nn_brute_launch(const pcreg_model* m, const double* d_sx, const double* d_sy, const double* d_sz, int64_t ns,
                     const double* d_T, int64_t nhyp, const int32_t* d_prev, int32_t* d_idx, double* d_d2,
                     NNScratch& sc, cudaStream_t st) {
    BruteArgs a{};
    a.m4 = m->m4.p; a.perm = m->perm.p; a.md = m->md.p;
    a.sx = d_sx; a.sy = d_sy; a.sz = d_sz; a.ns = ns; a.T = d_T; a.nq = nhyp * ns;
    a.px = m->pivot[0]; a.py = m->pivot[1]; a.pz = m->pivot[2]; a.max_norm = m->max_norm;
    PCREG_REQUIRE(a.nq > 0, "nn_brute: no queries");
    PCREG_REQUIRE((a.nq + BRUTE_THREADS * 4 - 1) / (BRUTE_THREADS * 4) < 2147483647LL, "nn_brute: too many queries in one launch");
    // few queries: fewer per thread so that more threads share the work; many: 8 per thread
    if (a.nq >= (int64_t)ctx().sm_count * BRUTE_THREADS * 8) brute_run<8>(m, a, d_prev, d_idx, d_d2, sc, st);
    else                                                    brute_run<4>(m, a, d_prev, d_idx, d_d2, sc, st);
}

}  // namespace pcreg
