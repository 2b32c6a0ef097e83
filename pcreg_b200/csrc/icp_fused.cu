// icp_fused.cu -- the whole ICP of one hypothesis in ONE thread block: every pass (exact NN -> select / weight -> 17 FP64 sums
// -> Kabsch SVD -> pose compose) runs inside a single kernel launch, the hypotheses of a batch being independent
// (the reference runs them under parfor: slideMatchingWindow_v2.m:178, completeExperiment.m:265).
//
// Why.  With the Voronoi voxel map (nn_vox.cu) an NN pass costs ~0.6 ms for 20 M queries; the per-pass kernels then spend as
// much again writing correspondences / residuals / trim keys to HBM and reading them back, and a small batch (strong scaling:
// a few hundred hypotheses per GPU) is bound by the 60-odd launches.  Here the correspondences and the trim keys of a hypothesis live in
// SHARED memory (4 + 8 bytes per source point), the pose in shared memory, and nothing but the voxel lists, the source cloud
// and the gathered model points is read from global memory; there is one launch per batch and no host synchronisation.
//
// Same arithmetic as the per-pass path (pcreg_icp.cuh, pcreg_vox.cuh, pcreg_select.cuh are shared; thread t sums the
// correspondences t, t + UPD_THREADS, ... in both), so poses, RMSE history and correspondences are bit-identical to it and to the
// brute-force path (tests/test_gpu_icp.py).  Composition as in icp.cu: quickTF.m:5-7, ransac.m:49, AlignPoints_KNN.m:20-26,
// AlignPoints_weighted.m:16-18, estimateTransform.m:41-71.
// A query whose voxel has no list (or that lies outside the padded box) is walked through the occupancy pyramid by its own
// thread (walk_search, pcreg_grid.cuh), warm-started by its previous correspondence.
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <algorithm>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"
#include "pcreg_grid.cuh"
#include "pcreg_vox.cuh"
#include "pcreg_select.cuh"
#include "pcreg_icp.cuh"

namespace pcreg {

#ifndef FUSED_SKIP_UNSELECTED
#define FUSED_SKIP_UNSELECTED 1 // trimmed mode: the sums pass does not gather the model points of unselected correspondences (C3: 35.8 -> 35.2 ms)
#endif
#ifndef FUSED_INLINE_SUMS
#define FUSED_INLINE_SUMS 0     // 1: modes without a trim accumulate the 17 sums inside the NN loop (no sums pass, no second gather).
#endif                          // Measured SLOWER (C4 polish, 16 384 poses: 592 vs 465 ms): the 34 accumulator registers push the NN loop
                                // into spills.  Kept as a switch.

__device__ __noinline__ void fused_walk(const GridArgs& a, double qx, double qy, double qz, int32_t warm, int32_t& bidx, double& best) {
    Query Q;
    Q.qx = qx; Q.qy = qy; Q.qz = qz;
    setup_query_at(a.g, a.md, Q, 0.f, warm);
    int lc = 0, ext_slot = -1;
    double thr2 = 0.0;
    unsigned long long n1 = 0, n2 = 0, n3 = 0;
    walk_search<false>(a, 0, Q, false, lc, ext_slot, thr2, 0.f, n1, n2, n3);
    bidx = Q.bidx; best = Q.best;
}
__device__ __noinline__ bool fused_pose_update(const double (&s)[KABSCH_NSUMS], long long n_used, const double* pivot, bool reflection_fix,
                                               double* Ts /*shared, row-major 4x4, updated in place*/) {
    double Tc[16], Tn[16];
    for (int k = 0; k < 16; ++k) Tc[k] = Ts[k];
    if (!icp_pose_update(s, n_used, pivot, reflection_fix, Tc, Tn)) return false;
    for (int k = 0; k < 16; ++k) Ts[k] = Tn[k];
    return true;
}

// KNN = the trimmed mode (its selection scratch -- 18 KB of shared memory -- exists only in that instantiation).
template <bool KNN>
__global__ void __launch_bounds__(UPD_THREADS, FUSED_MIN_BLOCKS) k_icp_fused(const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ double Ts[16];
    __shared__ double red[KABSCH_NSUMS * 32];
    __shared__ long long redll[32];
    __shared__ HistSelShared hsel;
    __shared__ unsigned long long red_u64[64];
    __shared__ int s_frozen;

    const int ns = a.ns;
    int32_t* __restrict__ idx_s = (int32_t*)dyn;
    unsigned long long* __restrict__ keys_s = (unsigned long long*)(dyn + (((size_t)ns * 4 + 15) / 16) * 16);    // KNN mode only
    const int64_t h = blockIdx.x;
    const int tid = threadIdx.x;
    const VoxView& V = a.g.vox;
    const GridPoint* __restrict__ pts = a.g.g.pts;
    const bool reject = a.thDist2 > 0.0;
    constexpr bool knn = KNN;
    constexpr bool two_pass = KNN || !FUSED_INLINE_SUMS;      // sums in a pass of their own (needed after a trim selection)
    const double px = a.pivot[0], py = a.pivot[1], pz = a.pivot[2];
    unsigned long long c_read = 0, c_gather = 0, c_walk = 0;
    long long t_nn = 0, t_sel = 0, t_sum = 0, t_svd = 0, t_mark = 0;       // phase clocks of thread 0 (profiling only)
    const bool timing = a.counters != nullptr && tid == 0;

    if (tid < 16) Ts[tid] = a.T[h * 16 + tid];
    if (tid == 0) s_frozen = a.frozen[h];
    __syncthreads();

    for (int it = 0; it <= a.iters; ++it) {
        const bool last = it == a.iters;
        // ---- exact nearest neighbours of the hypothesis' ns queries (two per trip: their header / entry loads overlap) ----
        long long nkept = 0;
        unsigned long long kmin = ~0ull, kmax = 0ull;
        double s[KABSCH_NSUMS];
#pragma unroll
        for (int k = 0; k < KABSCH_NSUMS; ++k) s[k] = 0.0;
        long long n_used = 0;
        unsigned long long vK = 0ull;
        bool all_eq = false;
        auto weight_of = [&](int i, int32_t j, double d) -> double {
            const bool keep = j >= 0 && (!reject || d < a.thDist2);
            double w = 0.0;
            if (knn) w = (j >= 0 && key_selected(keys_s[i], vK, all_eq)) ? 1.0 : 0.0;
            else if (a.mode == PCREG_ICP_PLAIN) w = keep ? 1.0 : 0.0;
            else if (keep) w = fmax(__dsub_rn(a.R_w, __dsqrt_rn(d)), 0.0);
            if (a.w_src) w = __dmul_rn(w, a.w_src[i]);
            return w;
        };
        if (timing) t_mark = clock64();
        for (int i0 = tid; i0 < ns; i0 += 2 * UPD_THREADS) {
            const int i1 = i0 + UPD_THREADS;
            const bool has1 = i1 < ns;
            const int i1c = has1 ? i1 : i0;
            double q0x, q0y, q0z, q1x, q1y, q1z;
            quick_tf(Ts, a.sx[i0], a.sy[i0], a.sz[i0], q0x, q0y, q0z);
            quick_tf(Ts, a.sx[i1c], a.sy[i1c], a.sz[i1c], q1x, q1y, q1z);
            uint2 hd0 = make_uint2(0u, 0u), hd1 = make_uint2(0u, 0u);
            float x0, y0, z0, x1, y1, z1;
            const bool ok0 = vox_lookup(V, q0x, q0y, q0z, hd0, x0, y0, z0);
            const bool ok1 = has1 && vox_lookup(V, q1x, q1y, q1z, hd1, x1, y1, z1);
            int32_t b0 = -1, b1 = -1;
            double d0 = INFINITY, d1 = INFINITY;
            unsigned ng = 0;
            double w0c[3], w1c[3];                               // the winners' coordinates (modes without a trim: sums right here)
            if (ok0) { vox_scan<8, true>(V, pts, hd0, x0, y0, z0, q0x, q0y, q0z, b0, d0, ng, two_pass ? nullptr : w0c); c_read += hd0.y; c_gather += ng; }
            else     { fused_walk(a.g, q0x, q0y, q0z, it > 0 ? idx_s[i0] : -1, b0, d0); ++c_walk; }
            if (has1) {
                if (ok1) { vox_scan<8, true>(V, pts, hd1, x1, y1, z1, q1x, q1y, q1z, b1, d1, ng, two_pass ? nullptr : w1c); c_read += hd1.y; c_gather += ng; }
                else     { fused_walk(a.g, q1x, q1y, q1z, it > 0 ? idx_s[i1] : -1, b1, d1); ++c_walk; }
            }
            idx_s[i0] = b0;
            if (has1) idx_s[i1] = b1;
            if (!two_pass) {
                ModelPointD m;
                if (ok0) { m.x = w0c[0]; m.y = w0c[1]; m.z = w0c[2]; m.pad = 0.0; } else m = a.g.md[b0 >= 0 ? b0 : 0];
                const double w0 = weight_of(i0, b0, d0);
                if (w0 > 0.0) ++n_used;
                icp_accumulate(s, w0, d0, q0x, q0y, q0z, m, px, py, pz);
                if (has1) {
                    if (ok1) { m.x = w1c[0]; m.y = w1c[1]; m.z = w1c[2]; m.pad = 0.0; } else m = a.g.md[b1 >= 0 ? b1 : 0];
                    const double w1 = weight_of(i1, b1, d1);
                    if (w1 > 0.0) ++n_used;
                    icp_accumulate(s, w1, d1, q1x, q1y, q1z, m, px, py, pz);
                }
            }
            if (knn) {
                const bool keep0 = b0 >= 0 && (!reject || d0 < a.thDist2);
                const unsigned long long key0 = keep0 ? dbits(__dsqrt_rn(d0)) : KEY_NOSEL;
                keys_s[i0] = key0;
                if (keep0) { ++nkept; kmin = key0 < kmin ? key0 : kmin; kmax = key0 > kmax ? key0 : kmax; }
                if (has1) {
                    const bool keep1 = b1 >= 0 && (!reject || d1 < a.thDist2);
                    const unsigned long long key1 = keep1 ? dbits(__dsqrt_rn(d1)) : KEY_NOSEL;
                    keys_s[i1] = key1;
                    if (keep1) { ++nkept; kmin = key1 < kmin ? key1 : kmin; kmax = key1 > kmax ? key1 : kmax; }
                }
            }
        }
        __syncthreads();
        if (timing) { const long long t = clock64(); t_nn += t - t_mark; t_mark = t; }
        // ---- trim: the round(k_frac * kept) smallest residuals, stable tie rule (AlignPoints_KNN.m:20-26) ----
        if (knn) {
            nkept = block_sum_ll(nkept, redll);
            const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long a0 = __shfl_xor_sync(0xffffffffu, kmin, o), a1 = __shfl_xor_sync(0xffffffffu, kmax, o);
                kmin = a0 < kmin ? a0 : kmin;
                kmax = a1 > kmax ? a1 : kmax;
            }
            if (lane == 0) { red_u64[warp] = kmin; red_u64[32 + warp] = kmax; }
            __syncthreads();
            kmin = ~0ull; kmax = 0ull;
            for (int w = 0; w < UPD_THREADS / 32; ++w) {
                kmin = red_u64[w] < kmin ? red_u64[w] : kmin;
                kmax = red_u64[32 + w] > kmax ? red_u64[32 + w] : kmax;
            }
            __syncthreads();
            long long K = (long long)floor(a.k_frac * (double)nkept + 0.5);     // MATLAB round (AlignPoints_KNN.m:21)
            if (K > nkept) K = nkept;
            block_hist_select(keys_s, ns, K, kmin, kmax, hsel, vK, all_eq, a.tie_order);
        }
        if (timing) { const long long t = clock64(); t_sel += t - t_mark; t_mark = t; }
        // ---- weights + the 17 sums (thread t: correspondences t, t + UPD_THREADS, ... in this order, as k_icp_update) ----
        if (two_pass) {
            for (int i = tid; i < ns; i += UPD_THREADS) {
                const int32_t j = idx_s[i];
#if FUSED_SKIP_UNSELECTED
                // trimmed mode: a correspondence outside the selection has weight zero and would add exact zeros to every sum --
                // its model point is not gathered at all
                if (knn && !(j >= 0 && key_selected(keys_s[i], vK, all_eq))) continue;
#endif
                double qx, qy, qz;
                quick_tf(Ts, a.sx[i], a.sy[i], a.sz[i], qx, qy, qz);
                const ModelPointD m = a.g.md[j >= 0 ? j : 0];
                const double d = j >= 0 ? dist2_exact(m.x, m.y, m.z, qx, qy, qz) : (double)INFINITY;      // the NN step's d2, bit for bit
                const double w = weight_of(i, j, d);
                if (w > 0.0) ++n_used;
                icp_accumulate(s, w, d, qx, qy, qz, m, px, py, pz);
            }
        }
        block_sum<KABSCH_NSUMS>(s, red);
        n_used = block_sum_ll(n_used, redll);
        if (timing) { const long long t = clock64(); t_sum += t - t_mark; t_mark = t; }
        if (tid == 0) {
            const double sw = s[0];
            const double rmse = (sw > 0.0) ? sqrt(s[16] / sw) : nan("");
            if (a.rmse_hist) a.rmse_hist[h * a.hist_stride + it] = rmse;
            if (last) { a.rmse[h] = rmse; a.n_used[h] = (int32_t)n_used; }
            if (!last && !s_frozen) {
                if (!fused_pose_update(s, n_used, a.pivot, a.reflection_fix != 0, Ts)) s_frozen = 1;
            }
        }
        if (timing) { const long long t = clock64(); t_svd += t - t_mark; t_mark = t; }
        __syncthreads();
    }
    if (tid < 16) a.T[h * 16 + tid] = Ts[tid];
    if (tid == 0) a.frozen[h] = s_frozen;
    if (a.idx_out) {
        int32_t* __restrict__ out = a.idx_out + h * (int64_t)ns;
        for (int r = tid; r < ns; r += UPD_THREADS) out[a.perm ? a.perm[r] : r] = idx_s[r];
    }
    if (a.counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_read += __shfl_xor_sync(0xffffffffu, c_read, o);
            c_gather += __shfl_xor_sync(0xffffffffu, c_gather, o);
            c_walk += __shfl_xor_sync(0xffffffffu, c_walk, o);
        }
        if (tid == 0) {
            atomicAdd(&a.counters[11], (unsigned long long)t_nn); atomicAdd(&a.counters[12], (unsigned long long)t_sel);
            atomicAdd(&a.counters[13], (unsigned long long)t_sum); atomicAdd(&a.counters[14], (unsigned long long)t_svd);
        }
        if ((tid & 31) == 0) {
            if (c_read) atomicAdd(&a.counters[6], c_read);
            if (c_gather) atomicAdd(&a.counters[7], c_gather);
            if (c_walk) atomicAdd(&a.counters[4], c_walk);
        }
    }
}

size_t icp_fused_smem(int64_t ns, int mode) {
    return (((size_t)ns * 4 + 15) / 16) * 16 + (mode == PCREG_ICP_KNN ? (size_t)ns * 8 : 0);
}

bool icp_fused_eligible(const pcreg_model* m, int64_t ns, int64_t nhyp, const pcreg_icp_opts& o) {
    if (!m->has_vox || o.nn != PCREG_NN_GRID) return false;
    if (const char* e = getenv("PCREG_FUSED")) { if (e[0] == '0') return false; if (e[0] == '1') nhyp = (int64_t)1 << 40; }
    // one block per hypothesis: a batch smaller than the SM count leaves SMs idle, the per-pass kernels (all SMs on every
    // pass) are faster there (C2, one hypothesis: 13.9 ms fused vs the per-pass path; PCREG_FUSED=1 forces the fused kernel)
    if (nhyp < ctx().sm_count) return false;
    if (ns >= ((int64_t)1 << 30)) return false;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k_icp_fused<true>) != cudaSuccess) { cudaGetLastError(); return false; }
    return icp_fused_smem(ns, o.mode) + fa.sharedSizeBytes + 1024 <= ctx().smem_optin;
}

void icp_fused_launch(FusedArgs& a, int64_t nhyp, cudaStream_t st) {
    const size_t dyn = icp_fused_smem(a.ns, a.mode);
    if (a.mode == PCREG_ICP_KNN) {
        PCREG_CUDA(cudaFuncSetAttribute(k_icp_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        k_icp_fused<true><<<(unsigned)nhyp, UPD_THREADS, dyn, st>>>(a);
    } else {
        PCREG_CUDA(cudaFuncSetAttribute(k_icp_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        k_icp_fused<false><<<(unsigned)nhyp, UPD_THREADS, dyn, st>>>(a);
    }
    PCREG_LAUNCHED();
}

}  // namespace pcreg
