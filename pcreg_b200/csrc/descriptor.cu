// descriptor.cu -- getSpacialHistogramDescriptors.m:2-183 on the GPU: for every keypoint the radius-R neighbourhood
// (getLocalPoints.m:5-36, batched, local_points.cu), the inline PCA local reference frame with variance rejection
// and majority-vote sign disambiguation (:71-144), and the trivariate spherical histogram (:147-172, histcn.m:97-131).
// One thread block per accepted keypoint; the neighbourhood points never leave the device.
//
// Reference quirks reproduced on purpose:
//   * phi = atan2(y, y) (:152) -- phi only takes the values pi/4, -3pi/4, 0, -pi;
//   * the sign votes count the K selected points against K/2 (:128-131; AlignPoints_KNN.m uses N/2 instead);
//   * histcounts semantics: left-closed bins, the last bin also holds its right edge, anything else (outside the edges,
//     NaN -- e.g. theta of a point that coincides with the keypoint) is dropped (histcn.m:125).
#include <math.h>
#include <algorithm>
#include <vector>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"
#include "pcreg_math.cuh"
#include "pcreg_select.cuh"

namespace pcreg {

constexpr int DESC_THREADS = 256;
constexpr int DESC_MAX_BINS = 4096;

struct DescArgs {
    const double* pts; int64_t ld;          // [3][ld] neighbourhood points relative to their keypoint
    const int64_t* offsets;                 // [nkey + 1]
    const int32_t* lp_status;               // [nkey] 1 = getLocalPoints returned []
    int knn; double k_frac;                 // options.k: 'all' / 1 -> knn = 0
    int need_pca, align; double th0, th1;   // options.thVar, options.ALIGN_POINTS
    const double* e_r; const double* e_t; const double* e_p; int nr, nt, np;      // bin edges (n + 1 values each)
    unsigned long long* keys;               // [ntotal] scratch
    double* desc;                           // [nkey][nr*nt*np]
    int32_t* status;                        // [nkey] 0 = descriptor valid, 1 = no neighbourhood, 2 = variance rejection
};

// third output of histcounts(x, edges): 1-based bin, 0 = not counted
__device__ __forceinline__ int hist_bin(double x, const double* __restrict__ e, int n) {
    if (!(x >= e[0]) || !(x <= e[n])) return 0;          // also catches NaN
    if (x == e[n]) return n;
    int lo = 0, hi = n;                                  // invariant: e[lo] <= x < e[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= e[mid]) lo = mid; else hi = mid;
    }
    return lo + 1;
}

__global__ void __launch_bounds__(DESC_THREADS) k_spatial_hist(const __grid_constant__ DescArgs a) {
    __shared__ double red[10 * 32];
    __shared__ long long redll[32];
    __shared__ HistSelShared hsel;
    __shared__ unsigned long long red_u64[64];
    __shared__ double sh_pca[9], sh_coeff[9];
    __shared__ int sh_reject;
    __shared__ int hist[DESC_MAX_BINS];

    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const int nbins = a.nr * a.nt * a.np;
    double* __restrict__ out = a.desc + b * nbins;
    const int64_t r0 = a.offsets[b];
    const int64_t N = a.offsets[b + 1] - r0;
    if (a.lp_status[b] != 0 || N <= 0) {
        if (tid == 0) a.status[b] = 1;
        for (int j = tid; j < nbins; j += DESC_THREADS) out[j] = nan("");
        return;
    }
    const double* __restrict__ X = a.pts + r0;
    const double* __restrict__ Y = a.pts + a.ld + r0;
    const double* __restrict__ Z = a.pts + 2 * a.ld + r0;
    unsigned long long* __restrict__ keys = a.keys + r0;

    // ---- K nearest to the centroid (:76-83; stable sort -> ties by lower index) ----
    long long K = N;
    unsigned long long vK = 0ull;
    bool all_eq = false;
    if (a.knn) {
        double s[3] = {0.0, 0.0, 0.0};
        for (int64_t i = tid; i < N; i += DESC_THREADS) { s[0] += X[i]; s[1] += Y[i]; s[2] += Z[i]; }
        block_sum<3>(s, red);
        const double c0 = s[0] / (double)N, c1 = s[1] / (double)N, c2 = s[2] / (double)N;
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int64_t i = tid; i < N; i += DESC_THREADS) {
            const unsigned long long key = dbits(norm3_exact(X[i] - c0, Y[i] - c1, Z[i] - c2));
            keys[i] = key;
            kmin = key < kmin ? key : kmin; kmax = key > kmax ? key : kmax;
        }
        K = (long long)floor((double)N * a.k_frac + 0.5);                     // round(num_points*K), :77
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
            for (int w = 0; w < DESC_THREADS / 32; ++w) {
                kmin = red_u64[w] < kmin ? red_u64[w] : kmin;
                kmax = red_u64[32 + w] > kmax ? red_u64[32 + w] : kmax;
            }
            __syncthreads();
        }
        block_hist_select(keys, N, K, kmin, kmax, hsel, vK, all_eq, nullptr);
    }
    auto member = [&](int64_t i) -> bool { return !a.knn || key_selected(keys[i], vK, all_eq); };

    if (tid == 0) sh_reject = 0;
    __syncthreads();
    if (a.need_pca) {
        // ---- pca(pts_k, 'Algorithm', 'eig') (:90): mean, covariance / (K - 1), eigenvectors by descending eigenvalue ----
        double mu[3];
        {
            double s[3] = {0.0, 0.0, 0.0};
            for (int64_t i = tid; i < N; i += DESC_THREADS) if (member(i)) { s[0] += X[i]; s[1] += Y[i]; s[2] += Z[i]; }
            block_sum<3>(s, red);
            for (int k = 0; k < 3; ++k) mu[k] = K > 0 ? s[k] / (double)K : 0.0;
        }
        {
            double s[6] = {0, 0, 0, 0, 0, 0};
            for (int64_t i = tid; i < N; i += DESC_THREADS) {
                if (!member(i)) continue;
                const double u0 = X[i] - mu[0], u1 = Y[i] - mu[1], u2 = Z[i] - mu[2];
                s[0] += u0 * u0; s[1] += u0 * u1; s[2] += u0 * u2; s[3] += u1 * u1; s[4] += u1 * u2; s[5] += u2 * u2;
            }
            block_sum<6>(s, red);
            if (tid == 0) {
                const double dof = (double)(K > 1 ? K - 1 : 1);
                double A[9] = {s[0] / dof, s[1] / dof, s[2] / dof, s[1] / dof, s[3] / dof, s[4] / dof, s[2] / dof, s[4] / dof, s[5] / dof};
                double w[3], V[9];
                eigsym3(A, w, V);
                eigsort3(w, V, -1);
                for (int col = 0; col < 3; ++col) {      // pca sign convention: largest |element| positive
                    int im = 0;
                    double am = fabs(V[0 * 3 + col]);
                    for (int r = 1; r < 3; ++r) if (fabs(V[r * 3 + col]) > am) { am = fabs(V[r * 3 + col]); im = r; }
                    if (V[im * 3 + col] < 0.0) for (int r = 0; r < 3; ++r) V[r * 3 + col] = -V[r * 3 + col];
                }
                for (int k = 0; k < 9; ++k) sh_pca[k] = V[k];
                if ((w[0] / w[1] < a.th0) || (w[1] / w[2] < a.th1)) sh_reject = 1;      // :117-120
            }
            __syncthreads();
        }
        if (sh_reject) {
            if (tid == 0) a.status[b] = 2;
            for (int j = tid; j < nbins; j += DESC_THREADS) out[j] = nan("");
            return;
        }
        if (a.align) {
            // ---- sign disambiguation over the K scores (:128-141) ----
            long long vx = 0, vz = 0;
            for (int64_t i = tid; i < N; i += DESC_THREADS) {
                if (!member(i)) continue;
                const double u0 = X[i] - mu[0], u1 = Y[i] - mu[1], u2 = Z[i] - mu[2];
                const double l0 = u0 * sh_pca[0] + u1 * sh_pca[3] + u2 * sh_pca[6];
                const double l2 = u0 * sh_pca[2] + u1 * sh_pca[5] + u2 * sh_pca[8];
                vx += l0 > 0.0 ? 1 : 0;
                vz += l2 > 0.0 ? 1 : 0;
            }
            vx = block_sum_ll(vx, redll);
            vz = block_sum_ll(vz, redll);
            if (tid == 0) {
                const double kthr = 0.5 * (double)K;
                const double xs = ((double)vx >= kthr) ? 1.0 : -1.0, zs = ((double)vz >= kthr) ? 1.0 : -1.0;
                double M[9];
                for (int r = 0; r < 3; ++r) { M[r * 3 + 0] = sh_pca[r * 3 + 0] * xs; M[r * 3 + 1] = sh_pca[r * 3 + 1]; M[r * 3 + 2] = sh_pca[r * 3 + 2] * zs; }
                const double ys = det3(M);               // not snapped to +-1 (:137)
                for (int r = 0; r < 3; ++r) {
                    sh_coeff[r * 3 + 0] = sh_pca[r * 3 + 0] * xs;
                    sh_coeff[r * 3 + 1] = sh_pca[r * 3 + 1] * ys;
                    sh_coeff[r * 3 + 2] = sh_pca[r * 3 + 2] * zs;
                }
            }
            __syncthreads();
        }
    }

    // ---- spherical histogram of ALL neighbourhood points (:147-164) ----
    for (int j = tid; j < nbins; j += DESC_THREADS) hist[j] = 0;
    __syncthreads();
    for (int64_t i = tid; i < N; i += DESC_THREADS) {
        double x = X[i], y = Y[i], z = Z[i];
        if (a.align) {                                                          // pts_local * coeff_unambig (:143)
            const double o0 = x * sh_coeff[0] + y * sh_coeff[3] + z * sh_coeff[6];
            const double o1 = x * sh_coeff[1] + y * sh_coeff[4] + z * sh_coeff[7];
            const double o2 = x * sh_coeff[2] + y * sh_coeff[5] + z * sh_coeff[8];
            x = o0; y = o1; z = o2;
        }
        const double r = norm3_exact(x, y, z);
        const double theta = acos(z / r);
        const double phi = atan2(y, y);                                         // sic (:152)
        const int br = hist_bin(r, a.e_r, a.nr), bt = hist_bin(theta, a.e_t, a.nt), bp = hist_bin(phi, a.e_p, a.np);
        if (br > 0 && bt > 0 && bp > 0) atomicAdd(&hist[(br - 1) + a.nr * ((bt - 1) + a.nt * (bp - 1))], 1);   // reshape(counts, [], 1), :164
    }
    __syncthreads();
    for (int j = tid; j < nbins; j += DESC_THREADS) out[j] = (double)hist[j];
    if (tid == 0) a.status[b] = 0;
}

}  // namespace pcreg

using namespace pcreg;

extern "C" {

void pcreg_desc_opts_default(pcreg_desc_opts* o) {
    if (!o) return;
    o->min_pts = 500; o->max_pts = 6000; o->R = 3.5; o->thVar[0] = 1.0; o->thVar[1] = 1.0; o->k_frac = 0.0; o->align_points = 1;
}

int pcreg_spatial_histogram(const pcreg_model* m, const double* keypoints, int64_t nkey, int64_t ld, const pcreg_desc_opts* opts,
                            const double* r_edges, int nr, const double* theta_edges, int nt, const double* phi_edges, int np,
                            double* desc, int32_t* status, int64_t* counts) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(m && keypoints && opts && r_edges && theta_edges && phi_edges && desc && status, "pcreg_spatial_histogram: null pointer");
    PCREG_REQUIRE(nkey >= 1 && ld >= nkey, "pcreg_spatial_histogram: bad sizes");
    PCREG_REQUIRE(nr >= 1 && nt >= 1 && np >= 1 && (int64_t)nr * nt * np <= DESC_MAX_BINS, "pcreg_spatial_histogram: at most 4096 bins");
    PCREG_REQUIRE(opts->R > 0.0, "pcreg_spatial_histogram: R must be positive");
    PCREG_CUDA(cudaSetDevice(ctx().device));
    cudaStream_t st = 0;
    const int nbins = nr * nt * np;
    LocalPointsDev lp;
    local_points_device(m, keypoints, nkey, ld, opts->R, opts->min_pts, opts->max_pts, lp, st);
    DevBuf<double> d_edges((size_t)(nr + nt + np + 3)), d_desc((size_t)nkey * nbins);
    DevBuf<int32_t> d_lpst((size_t)nkey), d_status((size_t)nkey);
    DevBuf<unsigned long long> d_keys((size_t)lp.nel);
    PCREG_CUDA(cudaMemcpyAsync(d_edges.p, r_edges, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemcpyAsync(d_edges.p + nr + 1, theta_edges, (size_t)(nt + 1) * 8, cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemcpyAsync(d_edges.p + nr + nt + 2, phi_edges, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemcpyAsync(d_lpst.p, lp.status.data(), (size_t)nkey * 4, cudaMemcpyHostToDevice, st));
    DescArgs a{};
    a.pts = lp.pts.p; a.ld = lp.nel; a.offsets = lp.d_offsets.p; a.lp_status = d_lpst.p;
    a.knn = (opts->k_frac > 0.0 && opts->k_frac != 1.0) ? 1 : 0;           // strcmp(K,'all') || K == 1 (:74)
    a.k_frac = opts->k_frac;
    a.align = opts->align_points ? 1 : 0;
    a.need_pca = (!(opts->thVar[0] == 1.0 && opts->thVar[1] == 1.0) || a.align) ? 1 : 0;      // :85
    a.th0 = opts->thVar[0]; a.th1 = opts->thVar[1];
    a.e_r = d_edges.p; a.e_t = d_edges.p + nr + 1; a.e_p = d_edges.p + nr + nt + 2; a.nr = nr; a.nt = nt; a.np = np;
    a.keys = d_keys.p; a.desc = d_desc.p; a.status = d_status.p;
    k_spatial_hist<<<(unsigned)nkey, DESC_THREADS, 0, st>>>(a);
    PCREG_LAUNCHED();
    PCREG_CUDA(cudaMemcpyAsync(desc, d_desc.p, d_desc.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(status, d_status.p, d_status.bytes(), cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
    if (counts) for (int64_t k = 0; k < nkey; ++k) counts[k] = lp.counts[(size_t)k];
    return PCREG_OK;
    PCREG_API_END
}

}  // extern "C"
