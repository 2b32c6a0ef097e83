// local_points.cu -- getLocalPoints.m:5-36 batched over centres (the reference's dominant descriptor-stage
// cost: one O(N_model) scan per keypoint, twice: getSpacialHistogramDescriptors.m:50,68).
//
// For each centre c_k: the model points with vecnorm(p - c_k) < R (strict, getLocalPoints.m:25), RELATIVE to
// c_k, in ORIGINAL model order (:26-28), and "[]" (status 1) when the count is outside [min_points, max_points]
// (:17-19,31-34; the cube pre-filter of :8-15 contains the sphere, so it never changes the result).
//
// Order-preserving brute-force compaction: a warp owns a chunk of 2048 consecutive model points and a tile of 32
// centres; the hit mask of every (32-point group, centre) is a ballot, running per-centre offsets live one per
// lane.  Model points are read once per 32 centres.  FP64 throughout (class double out).
#include <math.h>
#include <algorithm>
#include <vector>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"

namespace pcreg {

constexpr int LP_CHUNK = 2048;          // model points per warp
constexpr int LP_TILE = 32;             // centres per warp

struct LocalArgs {
    const ModelPointD* md; int64_t n;
    const double* cx; const double* cy; const double* cz; int64_t nc;
    double R;
    int32_t* chunk_cnt;                 // [ntiles*32][nchunks]  hits of centre k in chunk j
    // fill pass
    const int64_t* chunk_base;          // [ntiles*32][nchunks]  output row of the first hit of (k, j); < 0: skip centre
    double* out; int64_t ld_out; double* dists; int32_t* orig;
    int nchunks;
};

// vecnorm(p - c) < R decided exactly: sqrt only when d2 is within rounding of R^2
__device__ __forceinline__ bool inside(double dx, double dy, double dz, double R, double R2lo, double R2hi, double& dist) {
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (d2 > R2hi) return false;
    dist = __dsqrt_rn(d2);
    return d2 < R2lo || dist < R;
}

template <bool FILL>
__global__ void __launch_bounds__(32) k_local_points(const __grid_constant__ LocalArgs a) {
    const int lane = threadIdx.x;
    const int chunk = blockIdx.x;
    const int64_t k0 = (int64_t)blockIdx.y * LP_TILE;
    const int64_t kc = k0 + lane;                                  // this lane's centre (for the running counters)
    const bool kvalid = kc < a.nc;
    const double mcx = kvalid ? a.cx[kc] : 0.0, mcy = kvalid ? a.cy[kc] : 0.0, mcz = kvalid ? a.cz[kc] : 0.0;
    const double R2 = a.R * a.R, R2lo = R2 * (1.0 - 1e-15), R2hi = R2 * (1.0 + 1e-15);
    long long run = 0;                                             // hits of centre `lane` so far in this chunk
    long long base = 0;
    if (FILL && kvalid) base = a.chunk_base[kc * a.nchunks + chunk];
    const int ntile = (int)min((int64_t)LP_TILE, a.nc - k0);
    const int64_t p0 = (int64_t)chunk * LP_CHUNK;
    for (int64_t g = p0; g < min(a.n, p0 + LP_CHUNK); g += 32) {
        const int64_t i = g + lane;
        const bool pv = i < a.n;
        ModelPointD p;
        p.x = p.y = p.z = 0.0;
        if (pv) p = a.md[i];
        for (int c = 0; c < ntile; ++c) {
            const double ccx = __shfl_sync(0xffffffffu, mcx, c), ccy = __shfl_sync(0xffffffffu, mcy, c), ccz = __shfl_sync(0xffffffffu, mcz, c);
            const double dx = __dsub_rn(p.x, ccx), dy = __dsub_rn(p.y, ccy), dz = __dsub_rn(p.z, ccz);   // pts - c (getLocalPoints.m:23)
            double dist = 0.0;
            const bool hit = pv && inside(dx, dy, dz, a.R, R2lo, R2hi, dist);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (FILL) {
                const long long cb = __shfl_sync(0xffffffffu, base, c), cr = __shfl_sync(0xffffffffu, run, c);
                if (hit && cb >= 0) {
                    const int64_t row = cb + cr + __popc(m & ((1u << lane) - 1u));
                    a.out[row] = dx; a.out[a.ld_out + row] = dy; a.out[2 * a.ld_out + row] = dz;
                    if (a.dists) a.dists[row] = dist;
                    if (a.orig) a.orig[row] = (int32_t)i;
                }
            }
            if (lane == c) run += __popc(m);
        }
    }
    if (!FILL && kvalid) a.chunk_cnt[kc * a.nchunks + chunk] = (int32_t)run;
}

// per centre: exclusive scan of the chunk counts (+ the caller's row offset), or -1 for skipped centres
__global__ void k_local_scan(const int32_t* __restrict__ cnt, int nchunks, int64_t nc, const int64_t* __restrict__ offsets,
                             const int32_t* __restrict__ status, int64_t* __restrict__ base, int64_t* __restrict__ totals) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    long long acc = 0;
    const bool skip = status && status[k] != 0;
    const long long off = offsets ? offsets[k] : 0;
    for (int j = 0; j < nchunks; ++j) {
        if (base) base[k * nchunks + j] = skip ? -1 : off + acc;
        acc += cnt[k * nchunks + j];
    }
    if (totals) totals[k] = acc;
}

static void local_points_run(const pcreg_model* m, const double* centres, int64_t nc, int64_t ld, double R, bool fill,
                             const int64_t* offsets, const int32_t* status_in, int64_t* counts_out, double* pts_rel, int64_t ld_out,
                             double* dists, int32_t* orig) {
    PCREG_REQUIRE(m && centres && nc >= 1 && ld >= nc && R > 0.0, "local_points: bad arguments");
    cudaStream_t st = 0;
    const int nchunks = (int)((m->n + LP_CHUNK - 1) / LP_CHUNK);
    const int64_t ntiles = (nc + LP_TILE - 1) / LP_TILE;
    PCREG_REQUIRE(ntiles <= 65535, "local_points: at most 2,097,120 centres per call");
    DevBuf<double> dc((size_t)nc * 3);
    for (int a = 0; a < 3; ++a) PCREG_CUDA(cudaMemcpyAsync(dc.p + a * nc, centres + a * ld, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
    DevBuf<int32_t> cnt((size_t)ntiles * LP_TILE * nchunks);
    DevBuf<int64_t> totals((size_t)nc);
    LocalArgs a{};
    a.md = m->md.p; a.n = m->n; a.cx = dc.p; a.cy = dc.p + nc; a.cz = dc.p + 2 * nc; a.nc = nc; a.R = R;
    a.chunk_cnt = cnt.p; a.nchunks = nchunks;
    dim3 grid((unsigned)nchunks, (unsigned)ntiles);
    k_local_points<false><<<grid, 32, 0, st>>>(a);
    PCREG_LAUNCHED();
    if (!fill) {
        k_local_scan<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(cnt.p, nchunks, nc, nullptr, nullptr, nullptr, totals.p);
        PCREG_LAUNCHED();
        PCREG_CUDA(cudaMemcpyAsync(counts_out, totals.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
        PCREG_CUDA(cudaStreamSynchronize(st));
        return;
    }
    const int64_t ntotal = offsets[nc];
    PCREG_REQUIRE(ntotal >= 0 && ld_out >= ntotal, "local_points_fill: bad offsets / ld_out");
    DevBuf<int64_t> d_off((size_t)nc + 1), base((size_t)nc * nchunks);
    DevBuf<int32_t> d_st(status_in ? (size_t)nc : 0), d_orig(orig ? (size_t)std::max<int64_t>(ntotal, 1) : 0);
    const size_t nel = (size_t)std::max<int64_t>(ntotal, 1);
    DevBuf<double> d_out(nel * 3), d_dist(dists ? nel : 0);
    PCREG_CUDA(cudaMemcpyAsync(d_off.p, offsets, ((size_t)nc + 1) * 8, cudaMemcpyHostToDevice, st));
    if (status_in) PCREG_CUDA(cudaMemcpyAsync(d_st.p, status_in, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    k_local_scan<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(cnt.p, nchunks, nc, d_off.p, status_in ? d_st.p : nullptr, base.p, nullptr);
    PCREG_LAUNCHED();
    a.chunk_base = base.p; a.out = d_out.p; a.ld_out = (int64_t)nel; a.dists = dists ? d_dist.p : nullptr; a.orig = orig ? d_orig.p : nullptr;
    k_local_points<true><<<grid, 32, 0, st>>>(a);
    PCREG_LAUNCHED();
    for (int k = 0; k < 3; ++k)
        PCREG_CUDA(cudaMemcpyAsync(pts_rel + (size_t)k * ld_out, d_out.p + k * nel, (size_t)ntotal * 8, cudaMemcpyDeviceToHost, st));
    if (dists) PCREG_CUDA(cudaMemcpyAsync(dists, d_dist.p, (size_t)ntotal * 8, cudaMemcpyDeviceToHost, st));
    if (orig) PCREG_CUDA(cudaMemcpyAsync(orig, d_orig.p, (size_t)ntotal * 4, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
}

// Device-resident variant for the descriptor pipeline (descriptor.cu): count -> host prefix sum over the accepted
// centres -> fill; the neighbourhood points never leave the GPU.
void local_points_device(const pcreg_model* m, const double* centres, int64_t nc, int64_t ld, double R, int64_t min_points,
                         int64_t max_points, LocalPointsDev& out, cudaStream_t st) {
    PCREG_REQUIRE(m && centres && nc >= 1 && ld >= nc && R > 0.0, "local_points: bad arguments");
    const int nchunks = (int)((m->n + LP_CHUNK - 1) / LP_CHUNK);
    const int64_t ntiles = (nc + LP_TILE - 1) / LP_TILE;
    PCREG_REQUIRE(ntiles <= 65535, "local_points: at most 2,097,120 centres per call");
    DevBuf<double> dc((size_t)nc * 3);
    for (int a = 0; a < 3; ++a) PCREG_CUDA(cudaMemcpyAsync(dc.p + a * nc, centres + a * ld, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
    DevBuf<int32_t> cnt((size_t)ntiles * LP_TILE * nchunks);
    DevBuf<int64_t> totals((size_t)nc);
    LocalArgs a{};
    a.md = m->md.p; a.n = m->n; a.cx = dc.p; a.cy = dc.p + nc; a.cz = dc.p + 2 * nc; a.nc = nc; a.R = R;
    a.chunk_cnt = cnt.p; a.nchunks = nchunks;
    dim3 grid((unsigned)nchunks, (unsigned)ntiles);
    k_local_points<false><<<grid, 32, 0, st>>>(a);
    PCREG_LAUNCHED();
    k_local_scan<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(cnt.p, nchunks, nc, nullptr, nullptr, nullptr, totals.p);
    PCREG_LAUNCHED();
    out.counts.assign((size_t)nc, 0);
    PCREG_CUDA(cudaMemcpyAsync(out.counts.data(), totals.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
    out.status.assign((size_t)nc, 0);
    out.offsets.assign((size_t)nc + 1, 0);
    for (int64_t k = 0; k < nc; ++k) {
        const int64_t c = out.counts[k];
        out.status[k] = (c < min_points || (max_points >= 0 && c > max_points)) ? 1 : 0;          // getLocalPoints.m:31-34
        out.offsets[k + 1] = out.offsets[k] + (out.status[k] ? 0 : c);
    }
    out.ntotal = out.offsets[nc];
    out.nel = std::max<int64_t>(out.ntotal, 1);
    out.pts.alloc((size_t)out.nel * 3);
    out.d_offsets.alloc((size_t)nc + 1);
    DevBuf<int64_t> base((size_t)nc * nchunks);
    DevBuf<int32_t> d_st((size_t)nc);
    PCREG_CUDA(cudaMemcpyAsync(out.d_offsets.p, out.offsets.data(), ((size_t)nc + 1) * 8, cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemcpyAsync(d_st.p, out.status.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    k_local_scan<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(cnt.p, nchunks, nc, out.d_offsets.p, d_st.p, base.p, nullptr);
    PCREG_LAUNCHED();
    a.chunk_base = base.p; a.out = out.pts.p; a.ld_out = out.nel; a.dists = nullptr; a.orig = nullptr;
    k_local_points<true><<<grid, 32, 0, st>>>(a);
    PCREG_LAUNCHED();
    PCREG_CUDA(cudaStreamSynchronize(st));          // the per-call scratch above is released on return
}

}  // namespace pcreg

using namespace pcreg;

extern "C" {

int pcreg_local_points_count(const pcreg_model* m, const double* centres, int64_t nc, int64_t ld, double R, int64_t min_points,
                             int64_t max_points, int64_t* counts, int32_t* status) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(counts && status, "pcreg_local_points_count: null pointer");
    PCREG_CUDA(cudaSetDevice(ctx().device));
    local_points_run(m, centres, nc, ld, R, false, nullptr, nullptr, counts, nullptr, 0, nullptr, nullptr);
    for (int64_t k = 0; k < nc; ++k)
        status[k] = (counts[k] < min_points || (max_points >= 0 && counts[k] > max_points)) ? 1 : 0;      // getLocalPoints.m:31-34
    return PCREG_OK;
    PCREG_API_END
}

int pcreg_local_points_fill(const pcreg_model* m, const double* centres, int64_t nc, int64_t ld, double R, const int64_t* offsets,
                            const int32_t* status, double* pts_rel, int64_t ld_out, double* dists, int32_t* orig_idx) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(offsets && pts_rel, "pcreg_local_points_fill: null pointer");
    PCREG_CUDA(cudaSetDevice(ctx().device));
    local_points_run(m, centres, nc, ld, R, true, offsets, status, nullptr, pts_rel, ld_out, dists, orig_idx);
    return PCREG_OK;
    PCREG_API_END
}

}  // extern "C"
