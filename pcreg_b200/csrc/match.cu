// match.cu -- getMatches.m:1-56 on the GPU: descriptor weighting (constant "un-normalisation" element, element-wise
// power) followed by an EXHAUSTIVE matchFeatures (unit-vector normalisation, SAD / SSD scores, nearest + second
// nearest per surface descriptor, MatchThreshold, MaxRatio, forward-backward uniqueness).
//
// The reference's drivers ask matchFeatures for 'Method','Approximate' (completeExperiment.m:118), a randomised
// kd-forest of the closed Computer Vision Toolbox; this is the exact search it approximates.
//
// An L1 distance is not a contraction, so the score matrix is CUDA-core FP64 work: 2 DADD per (pair, dimension) for
// SAD (t = a - b; acc += |t|), DADD + DFMA for SSD.  k_match_scores computes 64 x 64 score tiles (256 threads, 4 x 4
// scores per thread, 16-dimension operand tiles double-buffered in shared memory) and never writes the score matrix:
// the epilogue reduces each tile to (nearest, second nearest, arg) per row and (nearest, arg) per column, small merge
// kernels combine the tiles.  Every sum runs over the dimensions in ascending order, ties go to the smaller index
// (MATLAB min / partial sort take the first).
#include <float.h>
#include <math.h>
#include <limits.h>
#include <algorithm>
#include <vector>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"

namespace pcreg {

constexpr int MT = 64;          // score tile edge
constexpr int MK = 16;          // dimensions per operand tile
constexpr int MTHREADS = 256;
constexpr int64_t MATCH_CHUNK_COLS = 16384;      // model descriptors per launch (bounds the partial buffers)

struct MatchArgs {
    const double* A; int64_t n1, n1p;        // [kdim][n1p] column-major working copy of the surface descriptors
    const double* B; int64_t n2, n2p;        // [kdim][n2p] model descriptors
    int kdim;                                // padded to a multiple of MK (padding is zero in both operands)
    int64_t col0;                            // first model descriptor of this launch
    double* row_d1; double* row_d2; int32_t* row_j1;     // [column blocks of the launch][n1p]
    double* col_d; int32_t* col_i;                       // [row blocks][columns of the launch]
    int64_t ncols;                                       // columns of the launch (multiple of MT)
};

struct Top2 { double d1, d2; int j1; };
__device__ __forceinline__ void top2_merge(Top2& a, double bd1, double bd2, int bj1) {
    const bool b_wins = bd1 < a.d1 || (bd1 == a.d1 && bj1 < a.j1);
    const double loser = b_wins ? a.d1 : bd1;
    a.d2 = fmin(fmin(a.d2, bd2), loser);
    if (b_wins) { a.d1 = bd1; a.j1 = bj1; }
}

template <int METRIC>
__global__ void __launch_bounds__(MTHREADS, 3) k_match_scores(const __grid_constant__ MatchArgs a) {
    __shared__ __align__(16) double As[2][MK][MT];
    __shared__ __align__(16) double Bs[2][MK][MT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t r0 = (int64_t)blockIdx.x * MT, c0 = a.col0 + (int64_t)blockIdx.y * MT;
    // operand tile = MK x MT doubles = 512 double2; thread t moves double2 number t and t + 256
    const int lk = tid >> 5, lc = (tid & 31) * 2;
    const double* gA = a.A + (int64_t)lk * a.n1p + r0 + lc;
    const double* gB = a.B + (int64_t)lk * a.n2p + c0 + lc;
    const int64_t sA8 = 8 * a.n1p, sB8 = 8 * a.n2p;

    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    double2 pa0 = *reinterpret_cast<const double2*>(gA), pa1 = *reinterpret_cast<const double2*>(gA + sA8);
    double2 pb0 = *reinterpret_cast<const double2*>(gB), pb1 = *reinterpret_cast<const double2*>(gB + sB8);
    *reinterpret_cast<double2*>(&As[0][lk][lc]) = pa0; *reinterpret_cast<double2*>(&As[0][lk + 8][lc]) = pa1;
    *reinterpret_cast<double2*>(&Bs[0][lk][lc]) = pb0; *reinterpret_cast<double2*>(&Bs[0][lk + 8][lc]) = pb1;
    __syncthreads();
    const int nkt = a.kdim / MK;
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) {
            gA += (int64_t)MK * a.n1p; gB += (int64_t)MK * a.n2p;
            pa0 = *reinterpret_cast<const double2*>(gA); pa1 = *reinterpret_cast<const double2*>(gA + sA8);
            pb0 = *reinterpret_cast<const double2*>(gB); pb1 = *reinterpret_cast<const double2*>(gB + sB8);
        }
#pragma unroll
        for (int k = 0; k < MK; ++k) {
            const double2 a01 = *reinterpret_cast<const double2*>(&As[buf][k][ty * 4]);
            const double2 a23 = *reinterpret_cast<const double2*>(&As[buf][k][ty * 4 + 2]);
            // columns {2tx, 2tx+1, 32+2tx, 32+2tx+1}: the 16 lanes of a half warp read 256 contiguous bytes per load
            // (a 32-byte stride would put lanes t and t+4 on the same banks)
            const double2 b01 = *reinterpret_cast<const double2*>(&Bs[buf][k][tx * 2]);
            const double2 b23 = *reinterpret_cast<const double2*>(&Bs[buf][k][32 + tx * 2]);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double t = av[i] - bv[j];
                    if (METRIC == PCREG_METRIC_SAD) acc[i][j] += fabs(t);
                    else acc[i][j] = fma(t, t, acc[i][j]);
                }
        }
        if (kt + 1 < nkt) {
            *reinterpret_cast<double2*>(&As[buf ^ 1][lk][lc]) = pa0; *reinterpret_cast<double2*>(&As[buf ^ 1][lk + 8][lc]) = pa1;
            *reinterpret_cast<double2*>(&Bs[buf ^ 1][lk][lc]) = pb0; *reinterpret_cast<double2*>(&Bs[buf ^ 1][lk + 8][lc]) = pb1;
        }
        __syncthreads();
    }

    // ---- epilogue: the tile never reaches memory ----
    const double INF = INFINITY;
    int colof[4];                                            // tile column of acc[.][j], ascending in j
#pragma unroll
    for (int j = 0; j < 4; ++j) colof[j] = tx * 2 + (j & 1) + (j >> 1) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (r0 + ty * 4 + i >= a.n1 || c0 + colof[j] >= a.n2) acc[i][j] = INF;       // padding rows / columns never win
    // rows: nearest / second nearest model descriptor among this tile's 64 columns
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        Top2 t{INF, INF, INT_MAX};
#pragma unroll
        for (int j = 0; j < 4; ++j) top2_merge(t, acc[i][j], INF, (int)(c0 + colof[j]));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {                    // the 16 threads of a row group are one half warp
            const double od1 = __shfl_xor_sync(0xffffffffu, t.d1, o), od2 = __shfl_xor_sync(0xffffffffu, t.d2, o);
            const int oj = __shfl_xor_sync(0xffffffffu, t.j1, o);
            top2_merge(t, od1, od2, oj);
        }
        if (tx == 0) {
            const int64_t o = (int64_t)blockIdx.y * a.n1p + r0 + ty * 4 + i;
            a.row_d1[o] = t.d1; a.row_d2[o] = t.d2; a.row_j1[o] = t.j1;
        }
    }
    // columns: nearest surface descriptor among this tile's 64 rows (first row on ties)
    double* cs_d = &As[0][0][0];                             // [16][64] doubles = As[0]
    int* cs_i = reinterpret_cast<int*>(&Bs[0][0][0]);        // [16][64] ints
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double d = INF; int bi = INT_MAX;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (acc[i][j] < d) { d = acc[i][j]; bi = (int)(r0 + ty * 4 + i); }
        cs_d[ty * MT + colof[j]] = d;
        cs_i[ty * MT + colof[j]] = bi;
    }
    __syncthreads();
    if (tid < MT) {
        double d = INF; int bi = INT_MAX;
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            const double v = cs_d[g * MT + tid];
            if (v < d) { d = v; bi = cs_i[g * MT + tid]; }
        }
        const int64_t o = (int64_t)blockIdx.x * a.ncols + (int64_t)blockIdx.y * MT + tid;
        a.col_d[o] = d; a.col_i[o] = bi;
    }
}

// merge the column-block partials of one launch into the running (nearest, second nearest, arg) of every row
__global__ void k_match_merge_rows(const double* __restrict__ pd1, const double* __restrict__ pd2, const int32_t* __restrict__ pj1,
                                   int64_t n1, int64_t n1p, int ncb, int first, double* d1, double* d2, int32_t* j1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    Top2 t{INFINITY, INFINITY, INT_MAX};
    if (!first) { t.d1 = d1[i]; t.d2 = d2[i]; t.j1 = j1[i]; }
    for (int cb = 0; cb < ncb; ++cb) top2_merge(t, pd1[(int64_t)cb * n1p + i], pd2[(int64_t)cb * n1p + i], pj1[(int64_t)cb * n1p + i]);
    d1[i] = t.d1; d2[i] = t.d2; j1[i] = t.j1;
}
// every row block of a launch is present: the nearest surface descriptor of each model descriptor of the launch
__global__ void k_match_merge_cols(const double* __restrict__ pd, const int32_t* __restrict__ pi, int64_t ncols, int nrb,
                                   int64_t col0, int64_t n2, int32_t* back) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols || col0 + c >= n2) return;
    double d = INFINITY; int bi = INT_MAX;
    for (int rb = 0; rb < nrb; ++rb) {
        const double v = pd[(int64_t)rb * ncols + c];
        if (v < d) { d = v; bi = pi[(int64_t)rb * ncols + c]; }
    }
    back[col0 + c] = bi;
}

// matchFeatures' three filters, per surface descriptor
__global__ void k_match_select(const double* __restrict__ d1, const double* __restrict__ d2, const int32_t* __restrict__ j1,
                               const int32_t* __restrict__ back, int64_t n1, int64_t n2, double thr, double max_ratio, int unique,
                               int32_t* keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const double a = d1[i];
    const int j = j1[i];
    bool ok = (a <= thr) && j >= 0 && (int64_t)j < n2;              // weak matches (also drops NaN scores)
    if (ok && n2 > 1) {                                             // ambiguous matches
        const double b = d2[i];
        const double ratio = (b < 1e-6) ? 1.0 : a / b;
        ok = ratio <= max_ratio;
    }
    if (ok && unique) ok = back[j] == (int32_t)i;                   // forward-backward
    keep[i] = ok ? 1 : 0;
}

// ---- descriptor weighting (getMatches.m:22-41) + matchFeatures' normalisation, in place on the working copies ----
__global__ void k_match_l1(const double* __restrict__ X, int64_t n, int64_t np, int dim, double* l1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < dim; ++k) s += fabs(X[(int64_t)k * np + i]);           // vecnorm(., 1, 2): the 1-norm (:24)
    l1[i] = s;
}
__global__ void __launch_bounds__(256) k_match_mean(const double* __restrict__ l1, int64_t n, double* out) {
    __shared__ double scratch[32];
    double v[1] = {0.0};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) v[0] += l1[i];
    block_sum<1>(v, scratch);
    if (threadIdx.x == 0) out[0] = v[0] / (double)n;
}
__global__ void k_match_weight(double* __restrict__ X, int64_t n, int64_t np, int dim, int unnormalize, double norm_factor,
                               const double* __restrict__ avg, int change_metric, double metric_factor) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ss = 0.0;
    for (int k = 0; k < dim; ++k) {
        double v = X[(int64_t)k * np + i];
        if (change_metric) { v = pow(v, metric_factor); X[(int64_t)k * np + i] = v; }      // :36-37
        ss += v * v;
    }
    int dimx = dim;
    if (unnormalize) {                                                                      // :25-26
        double c = norm_factor * avg[0];
        if (change_metric) c = pow(c, metric_factor);
        X[(int64_t)dim * np + i] = c;
        ss += c * c;
        dimx = dim + 1;
    }
    const double den = sqrt(ss) + DBL_EPSILON;                                              // matchFeatures: unit vectors
    for (int k = 0; k < dimx; ++k) X[(int64_t)k * np + i] = X[(int64_t)k * np + i] / den;
}

}  // namespace pcreg

using namespace pcreg;

extern "C" {

void pcreg_match_opts_default(pcreg_match_opts* o) {
    if (!o) return;
    // completeExperiment.m:112-122
    o->unnormalize = 1; o->norm_factor = 2.0; o->change_metric = 1; o->metric_factor = 0.6;
    o->match_threshold = 10.0; o->max_ratio = 0.99; o->metric = PCREG_METRIC_SAD; o->unique = 1;
}

int pcreg_get_matches(const double* desc_surface, int64_t n1, int64_t ld1, const double* desc_model, int64_t n2, int64_t ld2,
                      int64_t dim, const pcreg_match_opts* opts, int32_t* index_pairs, double* match_metric, int64_t* n_matches) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(opts && n_matches, "pcreg_get_matches: null pointer");
    PCREG_REQUIRE(n1 >= 0 && n2 >= 0 && dim >= 1 && dim < (1 << 20), "pcreg_get_matches: bad sizes");
    PCREG_REQUIRE(opts->metric == PCREG_METRIC_SAD || opts->metric == PCREG_METRIC_SSD, "pcreg_get_matches: metric must be SAD or SSD");
    *n_matches = 0;
    if (n1 == 0 || n2 == 0) return PCREG_OK;                       // matchFeatures returns an empty list
    PCREG_REQUIRE(desc_surface && desc_model && index_pairs, "pcreg_get_matches: null pointer");
    PCREG_REQUIRE(ld1 >= n1 && ld2 >= n2, "pcreg_get_matches: leading dimension smaller than the row count");
    PCREG_REQUIRE(n1 < ((int64_t)1 << 31) - MT && n2 < ((int64_t)1 << 31) - MT, "pcreg_get_matches: too many descriptors");
    PCREG_CUDA(cudaSetDevice(ctx().device));
    cudaStream_t st = 0;
    const int dimx = (int)dim + (opts->unnormalize ? 1 : 0);
    const int kdim = (dimx + MK - 1) / MK * MK;
    const int64_t n1p = (n1 + MT - 1) / MT * MT, n2p = (n2 + MT - 1) / MT * MT;
    const int nrb = (int)(n1p / MT);
    PCREG_REQUIRE(nrb <= 2147483647 / 1, "pcreg_get_matches: too many surface descriptors");

    DevBuf<double> A((size_t)kdim * n1p), B((size_t)kdim * n2p), l1((size_t)(n1 + n2)), avg(1);
    PCREG_CUDA(cudaMemsetAsync(A.p, 0, A.bytes(), st));
    PCREG_CUDA(cudaMemsetAsync(B.p, 0, B.bytes(), st));
    h2d_columns(A.p, n1p, desc_surface, ld1, n1, dim, st);           // pageable MATLAB / numpy memory -> pinned bounce buffers -> device
    h2d_columns(B.p, n2p, desc_model, ld2, n2, dim, st);
    const unsigned g1 = (unsigned)((n1 + 127) / 128), g2 = (unsigned)((n2 + 127) / 128);
    if (opts->unnormalize) {
        k_match_l1<<<g1, 128, 0, st>>>(A.p, n1, n1p, (int)dim, l1.p);
        PCREG_LAUNCHED();
        k_match_l1<<<g2, 128, 0, st>>>(B.p, n2, n2p, (int)dim, l1.p + n1);
        PCREG_LAUNCHED();
        k_match_mean<<<1, 256, 0, st>>>(l1.p, n1 + n2, avg.p);
        PCREG_LAUNCHED();
    }
    k_match_weight<<<g1, 128, 0, st>>>(A.p, n1, n1p, (int)dim, opts->unnormalize, opts->norm_factor, avg.p, opts->change_metric, opts->metric_factor);
    PCREG_LAUNCHED();
    k_match_weight<<<g2, 128, 0, st>>>(B.p, n2, n2p, (int)dim, opts->unnormalize, opts->norm_factor, avg.p, opts->change_metric, opts->metric_factor);
    PCREG_LAUNCHED();

    const int64_t chunk = std::min<int64_t>(n2p, MATCH_CHUNK_COLS);
    const int max_ncb = (int)(chunk / MT);
    DevBuf<double> pr_d1((size_t)max_ncb * n1p), pr_d2((size_t)max_ncb * n1p), pc_d((size_t)nrb * chunk);
    DevBuf<int32_t> pr_j1((size_t)max_ncb * n1p), pc_i((size_t)nrb * chunk);
    DevBuf<double> d1((size_t)n1), d2((size_t)n1);
    DevBuf<int32_t> j1((size_t)n1), back((size_t)n2), keep((size_t)n1);

    const bool prof = ctx().profiling;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float score_ms = 0.f;
    for (int64_t col0 = 0; col0 < n2p; col0 += chunk) {
        const int64_t ncols = std::min<int64_t>(chunk, n2p - col0);
        MatchArgs a{};
        a.A = A.p; a.n1 = n1; a.n1p = n1p; a.B = B.p; a.n2 = n2; a.n2p = n2p; a.kdim = kdim; a.col0 = col0; a.ncols = ncols;
        a.row_d1 = pr_d1.p; a.row_d2 = pr_d2.p; a.row_j1 = pr_j1.p; a.col_d = pc_d.p; a.col_i = pc_i.p;
        const dim3 grid((unsigned)nrb, (unsigned)(ncols / MT));
        if (prof) { e0 = pooled_event(0); e1 = pooled_event(1); PCREG_CUDA(cudaEventRecord(e0, st)); }
        if (opts->metric == PCREG_METRIC_SAD) k_match_scores<PCREG_METRIC_SAD><<<grid, MTHREADS, 0, st>>>(a);
        else                                  k_match_scores<PCREG_METRIC_SSD><<<grid, MTHREADS, 0, st>>>(a);
        PCREG_LAUNCHED();
        if (prof) {
            PCREG_CUDA(cudaEventRecord(e1, st));
            PCREG_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            PCREG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            score_ms += ms;
        }
        k_match_merge_rows<<<g1, 128, 0, st>>>(pr_d1.p, pr_d2.p, pr_j1.p, n1, n1p, (int)(ncols / MT), col0 == 0 ? 1 : 0, d1.p, d2.p, j1.p);
        PCREG_LAUNCHED();
        k_match_merge_cols<<<(unsigned)((ncols + 127) / 128), 128, 0, st>>>(pc_d.p, pc_i.p, ncols, nrb, col0, n2, back.p);
        PCREG_LAUNCHED();
    }
    // largest possible score between two unit vectors (matchFeatures: MatchThreshold is a percentage of it)
    const double max_val = opts->metric == PCREG_METRIC_SAD ? 2.0 * sqrt((double)dimx) : 4.0;
    const double thr = opts->match_threshold * 0.01 * max_val;
    k_match_select<<<g1, 128, 0, st>>>(d1.p, d2.p, j1.p, back.p, n1, n2, thr, opts->max_ratio, opts->unique ? 1 : 0, keep.p);
    PCREG_LAUNCHED();

    std::vector<int32_t> h_keep((size_t)n1), h_j1((size_t)n1);
    std::vector<double> h_d1((size_t)n1);
    PCREG_CUDA(cudaMemcpyAsync(h_keep.data(), keep.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(h_j1.data(), j1.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaMemcpyAsync(h_d1.data(), d1.p, (size_t)n1 * 8, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
    int64_t np = 0;
    for (int64_t i = 0; i < n1; ++i) {
        if (!h_keep[(size_t)i]) continue;
        index_pairs[2 * np] = (int32_t)i; index_pairs[2 * np + 1] = h_j1[(size_t)i];
        if (match_metric) match_metric[np] = h_d1[(size_t)i];
        ++np;
    }
    *n_matches = np;
    if (prof) {
        Context& c = ctx();
        for (int i = 0; i < 32; ++i) c.profile[i] = 0.0;
        c.profile[24] = (double)score_ms;
        c.profile[25] = (double)n1 * (double)n2 * (double)dimx;      // (pair, dimension) terms of the score matrix
    }
    return PCREG_OK;
    PCREG_API_END
}

}  // extern "C"
