// align.cu -- the six AlignPoints* PCA local-reference-frame functions, batched: one thread block per
// neighbourhood.  Restated from AlignPoints.m:1-29, AlignPoints_KNN.m:1-60, AlignPoints_knn.m:1-43,
// AlignPoints_weighted.m:1-49, AlignPoints_c.m:1-44, AlignPoints_KNN_c.m:1-57 (+ getLocalPoints.m:5-36)
// and MATLAB's documented pca(...,'Algorithm','eig'[, 'Centered','off']):
//   eigen-decomposition of Xc'*Xc/(n-1) (/n uncentred), eigenvalues DESCENDING, every coeff column
//   signed so that its largest-magnitude element is positive, score = Xc*coeff.
// All arithmetic is FP64 whatever the input class; pts_aligned is written back in the input class.
#include <math.h>
#include <float.h>
#include <algorithm>
#include <type_traits>
#include <vector>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"
#include "pcreg_math.cuh"
#include "pcreg_select.cuh"

namespace pcreg {

constexpr int ALIGN_THREADS = 256;

struct AlignArgs {
    int kind;
    const void* pts; int is_double; int64_t ld;
    const int64_t* offsets;
    double k_frac; int64_t k_abs; double R_w; double r_local; int64_t min_local; int C1; int C2;
    void* out; double* coeff9; double* c3; int32_t* status;
    unsigned long long* keys;       // [ntotal] scratch
};

struct PtLoader {
    const void* p; int is_double; int64_t ld; int64_t r0;
    __device__ __forceinline__ void get(int64_t i, double& x, double& y, double& z) const {
        if (is_double) {
            const double* d = (const double*)p;
            x = d[r0 + i]; y = d[ld + r0 + i]; z = d[2 * ld + r0 + i];
        } else {
            const float* f = (const float*)p;
            x = (double)f[r0 + i]; y = (double)f[ld + r0 + i]; z = (double)f[2 * ld + r0 + i];
        }
    }
};

// KNN = one of the kinds with a K-nearest-to-the-centroid selection (its scratch -- 17 KB of shared memory -- exists only in
// that instantiation)
template <bool KNN>
__global__ void __launch_bounds__(ALIGN_THREADS, 3) k_align_points(const __grid_constant__ AlignArgs a) {
    __shared__ double red[10 * 32];
    __shared__ long long redll[32];
    __shared__ typename std::conditional<KNN, HistSelShared, int>::type hsel_store;
    __shared__ unsigned long long red_u64[KNN ? 64 : 1];
    __shared__ double sh_coeff[9];      // coeff_unambig, row-major
    __shared__ double sh_pca[9];        // coeff before disambiguation

    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t r0 = a.offsets[b];
    const int64_t N = a.offsets[b + 1] - r0;
    const PtLoader L{a.pts, a.is_double, a.ld, r0};
    unsigned long long* __restrict__ keys = a.keys + r0;
    const int kind = a.kind;
    constexpr bool knn = KNN;

    if (N <= 0) {
        if (tid == 0) {
            a.status[b] = 1;
            for (int k = 0; k < 9; ++k) a.coeff9[b * 9 + k] = nan("");
            for (int k = 0; k < 3; ++k) a.c3[b * 3 + k] = nan("");
        }
        return;
    }

    // ---- 1) centroid of all points (mean(pts,1)) ----
    double c[3];
    {
        double s[3] = {0.0, 0.0, 0.0};
        for (int64_t i = tid; i < N; i += ALIGN_THREADS) {
            double x, y, z;
            L.get(i, x, y, z);
            s[0] += x; s[1] += y; s[2] += z;
        }
        block_sum<3>(s, red);
        for (int k = 0; k < 3; ++k) c[k] = s[k] / (double)N;
    }

    // ---- 2) K nearest to the centroid (KNN kinds) ----
    unsigned long long vK = 0ull;
    bool all_eq = false;
    long long K = 0;
    if (knn) {
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int64_t i = tid; i < N; i += ALIGN_THREADS) {
            double x, y, z;
            L.get(i, x, y, z);
            const unsigned long long key = dbits(norm3_exact(x - c[0], y - c[1], z - c[2]));     // AlignPoints_KNN.m:22-23
            keys[i] = key;
            kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax;
        }
        if (kind == PCREG_ALIGN_KNN_ABS) K = a.k_abs < N ? a.k_abs : N;     // AlignPoints_knn.m:12
        else                             K = (long long)floor((double)N * a.k_frac + 0.5);   // round(), AlignPoints_KNN.m:21
        if (K > N) K = N;
        if (K < 0) K = 0;
        {   // block min / max of the keys: the range of the one-pass histogram selection (pcreg_select.cuh)
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
            for (int w = 0; w < ALIGN_THREADS / 32; ++w) {
                kmin = red_u64[w] < kmin ? red_u64[w] : kmin;
                kmax = red_u64[32 + w] > kmax ? red_u64[32 + w] : kmax;
            }
            __syncthreads();
        }
        if constexpr (KNN) block_hist_select(keys, N, K, kmin, kmax, hsel_store, vK, all_eq, nullptr);
    }

    // ---- KNN_C: centroid of the selected relative points, then the r-ball around it ----
    double c2[3] = {0.0, 0.0, 0.0};
    if (kind == PCREG_ALIGN_KNN_C) {
        double s[3] = {0.0, 0.0, 0.0};
        for (int64_t i = tid; i < N; i += ALIGN_THREADS) {
            if (!key_selected(keys[i], vK, all_eq)) continue;
            double x, y, z;
            L.get(i, x, y, z);
            s[0] += x - c[0]; s[1] += y - c[1]; s[2] += z - c[2];
        }
        block_sum<3>(s, red);
        for (int k = 0; k < 3; ++k) c2[k] = K > 0 ? s[k] / (double)K : 0.0;  // AlignPoints_KNN_c.m:22
    }

    // membership of point i in the PCA set, and its PCA coordinates ("base")
    auto member = [&](int64_t i, double x, double y, double z, double* base) -> bool {
        switch (kind) {
            case PCREG_ALIGN_PLAIN:
                base[0] = x; base[1] = y; base[2] = z;
                return true;
            case PCREG_ALIGN_KNN_FRAC:
            case PCREG_ALIGN_KNN_ABS:
                base[0] = x - c[0]; base[1] = y - c[1]; base[2] = z - c[2];
                return key_selected(keys[i], vK, all_eq);
            case PCREG_ALIGN_WEIGHTED:
                base[0] = x - c[0]; base[1] = y - c[1]; base[2] = z - c[2];
                return true;
            case PCREG_ALIGN_C: {
                base[0] = x - c[0]; base[1] = y - c[1]; base[2] = z - c[2];
                return norm3_exact(base[0], base[1], base[2]) < a.r_local;   // getLocalPoints.m:23-25 (strict)
            }
            default: {   // KNN_C: relative to c, then relative to the sub-centroid (getLocalPoints.m:23)
                const double rx = x - c[0], ry = y - c[1], rz = z - c[2];
                base[0] = rx - c2[0]; base[1] = ry - c2[1]; base[2] = rz - c2[2];
                if (!key_selected(keys[i], vK, all_eq)) return false;
                return norm3_exact(base[0], base[1], base[2]) < a.r_local;
            }
        }
    };

    // ---- 3) size + mean of the PCA set ----
    double mu[3] = {0.0, 0.0, 0.0};
    long long n_set = 0;
    {
        double s[3] = {0.0, 0.0, 0.0};
        long long cnt = 0;
        for (int64_t i = tid; i < N; i += ALIGN_THREADS) {
            double x, y, z, bs[3];
            L.get(i, x, y, z);
            if (member(i, x, y, z, bs)) { s[0] += bs[0]; s[1] += bs[1]; s[2] += bs[2]; ++cnt; }
        }
        block_sum<3>(s, red);
        n_set = block_sum_ll(cnt, redll);
        const bool centred = !(kind == PCREG_ALIGN_WEIGHTED) && !(kind == PCREG_ALIGN_KNN_FRAC && a.C1);
        if (centred && n_set > 0) for (int k = 0; k < 3; ++k) mu[k] = s[k] / (double)n_set;
    }
    const bool need_local = (kind == PCREG_ALIGN_C || kind == PCREG_ALIGN_KNN_C);
    const bool empty = (need_local && n_set < a.min_local) || n_set <= 0;     // AlignPoints_c.m:16-18
    if (empty) {
        if (tid == 0) {
            a.status[b] = 1;
            for (int k = 0; k < 9; ++k) a.coeff9[b * 9 + k] = nan("");
            for (int k = 0; k < 3; ++k) a.c3[b * 3 + k] = c[k];
        }
        return;
    }

    // ---- 4) scatter matrix of the PCA set ----
    {
        double s[6] = {0, 0, 0, 0, 0, 0};
        for (int64_t i = tid; i < N; i += ALIGN_THREADS) {
            double x, y, z, bs[3];
            L.get(i, x, y, z);
            if (!member(i, x, y, z, bs)) continue;
            double w = 1.0;
            if (kind == PCREG_ALIGN_WEIGHTED) {
                const double d = norm3_exact(bs[0], bs[1], bs[2]);            // AlignPoints_weighted.m:13
                w = fmax(__dsub_rn(a.R_w, d), 0.0);                           // :16-18
            }
            const double u0 = bs[0] - mu[0], u1 = bs[1] - mu[1], u2 = bs[2] - mu[2];
            s[0] += w * u0 * u0; s[1] += w * u0 * u1; s[2] += w * u0 * u2;
            s[3] += w * u1 * u1; s[4] += w * u1 * u2; s[5] += w * u2 * u2;
        }
        block_sum<6>(s, red);
        if (tid == 0) {
            double dof = 1.0;
            if (kind != PCREG_ALIGN_WEIGHTED) {
                const bool uncentred = (kind == PCREG_ALIGN_KNN_FRAC && a.C1);
                dof = uncentred ? (double)n_set : (double)(n_set > 1 ? n_set - 1 : 1);
            }
            double A[9] = {s[0] / dof, s[1] / dof, s[2] / dof, s[1] / dof, s[3] / dof, s[4] / dof, s[2] / dof, s[4] / dof, s[5] / dof};
            double w[3], V[9];
            eigsym3(A, w, V);
            if (kind == PCREG_ALIGN_WEIGHTED) {
                eigsort3(w, V, +1);                 // eig() of a symmetric matrix: ascending (AlignPoints_weighted.m:24)
            } else {
                eigsort3(w, V, -1);                 // pca: descending
                for (int col = 0; col < 3; ++col) { // pca sign convention: largest |element| positive (first max on ties)
                    int im = 0;
                    double am = fabs(V[0 * 3 + col]);
                    for (int r = 1; r < 3; ++r) if (fabs(V[r * 3 + col]) > am) { am = fabs(V[r * 3 + col]); im = r; }
                    if (V[im * 3 + col] < 0.0) for (int r = 0; r < 3; ++r) V[r * 3 + col] = -V[r * 3 + col];
                }
            }
            for (int k = 0; k < 9; ++k) sh_pca[k] = V[k];
        }
        __syncthreads();
    }

    // ---- 5) sign votes (AlignPoints.m:10-25) ----
    {
        long long vx = 0, vz = 0;
        const bool raw_votes = (kind == PCREG_ALIGN_KNN_FRAC && a.C2);          // AlignPoints_KNN.m:39-41
        for (int64_t i = tid; i < N; i += ALIGN_THREADS) {
            double x, y, z, bs[3];
            L.get(i, x, y, z);
            double u0, u1, u2;
            if (raw_votes) { u0 = x; u1 = y; u2 = z; }
            else {
                if (!member(i, x, y, z, bs)) continue;
                u0 = bs[0] - mu[0]; u1 = bs[1] - mu[1]; u2 = bs[2] - mu[2];
            }
            const double l0 = u0 * sh_pca[0] + u1 * sh_pca[3] + u2 * sh_pca[6];
            const double l2 = u0 * sh_pca[2] + u1 * sh_pca[5] + u2 * sh_pca[8];
            vx += l0 > 0.0 ? 1 : 0;
            vz += l2 > 0.0 ? 1 : 0;
        }
        vx = block_sum_ll(vx, redll);
        vz = block_sum_ll(vz, redll);
        if (tid == 0) {
            // vote threshold: size(pts,1)/2 everywhere except AlignPoints_KNN_c (size(pts_lrf,1)/2, :34)
            const double kthr = (kind == PCREG_ALIGN_KNN_C) ? 0.5 * (double)n_set : 0.5 * (double)N;
            const double xs = ((double)vx >= kthr) ? 1.0 : -1.0;
            const double zs = ((double)vz >= kthr) ? 1.0 : -1.0;
            double M[9];
            for (int r = 0; r < 3; ++r) { M[r * 3 + 0] = sh_pca[r * 3 + 0] * xs; M[r * 3 + 1] = sh_pca[r * 3 + 1]; M[r * 3 + 2] = sh_pca[r * 3 + 2] * zs; }
            const double ys = det3(M);               // a float ~ +-1, NOT snapped (AlignPoints.m:22)
            for (int r = 0; r < 3; ++r) {
                sh_coeff[r * 3 + 0] = sh_pca[r * 3 + 0] * xs;
                sh_coeff[r * 3 + 1] = sh_pca[r * 3 + 1] * ys;
                sh_coeff[r * 3 + 2] = sh_pca[r * 3 + 2] * zs;
            }
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc) a.coeff9[b * 9 + cc * 3 + r] = sh_coeff[r * 3 + cc];    // column-major out
            for (int k = 0; k < 3; ++k) a.c3[b * 3 + k] = c[k];
            a.status[b] = 0;
        }
        __syncthreads();
    }

    // ---- 6) pts_aligned = pts * coeff_unambig (uncentred pts, AlignPoints.m:28) ----
    for (int64_t i = tid; i < N; i += ALIGN_THREADS) {
        double x, y, z;
        L.get(i, x, y, z);
        const double o0 = x * sh_coeff[0] + y * sh_coeff[3] + z * sh_coeff[6];
        const double o1 = x * sh_coeff[1] + y * sh_coeff[4] + z * sh_coeff[7];
        const double o2 = x * sh_coeff[2] + y * sh_coeff[5] + z * sh_coeff[8];
        if (a.is_double) {
            double* o = (double*)a.out;
            o[r0 + i] = o0; o[a.ld + r0 + i] = o1; o[2 * a.ld + r0 + i] = o2;
        } else {
            float* o = (float*)a.out;
            o[r0 + i] = (float)o0; o[a.ld + r0 + i] = (float)o1; o[2 * a.ld + r0 + i] = (float)o2;
        }
    }
}

}  // namespace pcreg

using namespace pcreg;

extern "C" {

void pcreg_align_opts_default(pcreg_align_opts* o) {
    if (!o) return;
    o->k_frac = 0.85; o->k_abs = 500; o->R_w = 3.5; o->r_local = 2.0; o->min_local = 25; o->C1 = 0; o->C2 = 0;
}

int pcreg_align_points(int kind, const void* pts, int is_double, int64_t ld, const int64_t* offsets, int64_t nbatch,
                       const pcreg_align_opts* opts, void* pts_aligned, double* coeff9, double* c3, int32_t* status) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(pts && offsets && pts_aligned && coeff9 && c3 && status, "pcreg_align_points: null pointer");
    PCREG_REQUIRE(kind >= PCREG_ALIGN_PLAIN && kind <= PCREG_ALIGN_KNN_C, "pcreg_align_points: bad kind");
    PCREG_REQUIRE(nbatch >= 1, "pcreg_align_points: nbatch must be >= 1");
    const int64_t ntotal = offsets[nbatch];
    PCREG_REQUIRE(offsets[0] == 0 && ntotal >= 0 && ld >= ntotal, "pcreg_align_points: bad offsets / ld");
    for (int64_t b = 0; b < nbatch; ++b) PCREG_REQUIRE(offsets[b + 1] >= offsets[b], "pcreg_align_points: offsets must be non-decreasing");
    pcreg_align_opts o;
    pcreg_align_opts_default(&o);
    if (opts) o = *opts;
    PCREG_CUDA(cudaSetDevice(ctx().device));
    cudaStream_t st = 0;
    const size_t el = is_double ? 8 : 4;
    const size_t nel = (size_t)std::max<int64_t>(ntotal, 1);
    DevBuf<unsigned char> d_in(nel * 3 * el), d_out(nel * 3 * el);
    DevBuf<unsigned long long> d_keys(nel);
    DevBuf<int64_t> d_off((size_t)nbatch + 1);
    DevBuf<double> d_coeff((size_t)nbatch * 9), d_c((size_t)nbatch * 3);
    DevBuf<int32_t> d_st((size_t)nbatch);
    for (int a = 0; a < 3; ++a)
        PCREG_CUDA(cudaMemcpyAsync(d_in.p + a * nel * el, (const unsigned char*)pts + (size_t)a * ld * el, (size_t)ntotal * el, cudaMemcpyHostToDevice, st));
    // rows of degenerate neighbourhoods stay untouched: start from the caller's output buffer
    for (int a = 0; a < 3; ++a)
        PCREG_CUDA(cudaMemcpyAsync(d_out.p + a * nel * el, (const unsigned char*)pts_aligned + (size_t)a * ld * el, (size_t)ntotal * el, cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemcpyAsync(d_off.p, offsets, ((size_t)nbatch + 1) * 8, cudaMemcpyHostToDevice, st));
    AlignArgs a{};
    a.kind = kind; a.pts = d_in.p; a.is_double = is_double; a.ld = (int64_t)nel; a.offsets = d_off.p;
    a.k_frac = o.k_frac; a.k_abs = o.k_abs; a.R_w = o.R_w; a.r_local = o.r_local; a.min_local = o.min_local; a.C1 = o.C1; a.C2 = o.C2;
    a.out = d_out.p; a.coeff9 = d_coeff.p; a.c3 = d_c.p; a.status = d_st.p; a.keys = d_keys.p;
    const bool prof = ctx().profiling;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) { e0 = pooled_event(0); e1 = pooled_event(1); PCREG_CUDA(cudaEventRecord(e0, st)); }
    if (a.kind == PCREG_ALIGN_KNN_FRAC || a.kind == PCREG_ALIGN_KNN_ABS || a.kind == PCREG_ALIGN_KNN_C)
        k_align_points<true><<<(unsigned)nbatch, ALIGN_THREADS, 0, st>>>(a);
    else
        k_align_points<false><<<(unsigned)nbatch, ALIGN_THREADS, 0, st>>>(a);
    PCREG_LAUNCHED();
    if (prof) PCREG_CUDA(cudaEventRecord(e1, st));
    for (int k = 0; k < 3; ++k)
        PCREG_CUDA(cudaMemcpyAsync((unsigned char*)pts_aligned + (size_t)k * ld * el, d_out.p + k * nel * el, (size_t)ntotal * el, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(coeff9, d_coeff.p, d_coeff.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(c3, d_c.p, d_c.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(status, d_st.p, d_st.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
    if (prof) {                                          // pcreg_last_profile: out[24] = kernel ms, out[25] = points
        float ms = 0.f;
        PCREG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        Context& c = ctx();
        for (int i = 0; i < 32; ++i) c.profile[i] = 0.0;
        c.profile[24] = ms; c.profile[25] = (double)ntotal;
    }
    return PCREG_OK;
    PCREG_API_END
}

}  // extern "C"
