// local_points.cu -- getLocalPoints.m:5-36 batched over centres (the reference's dominant descriptor-stage
// cost: one O(N_model) scan per keypoint, twice: getSpacialHistogramDescriptors.m:50,68).
//
// For each centre c_k: the model points with vecnorm(p - c_k) < R (strict, getLocalPoints.m:25), RELATIVE to
// c_k, in ORIGINAL model order (:26-28), and "[]" (status 1) when the count is outside [min_points, max_points]
// (:17-19,31-34; the cube pre-filter of :8-15 contains the sphere, so it never changes the result).
//
// Two implementations with identical results:
//  * models WITH a uniform grid (k_local_grid): one block per centre walks the (y,z) cell rows the ball can reach -- the cells
//    of a row are one contiguous run of the cell-sorted point array, exactly as the row scan of nn_grid.cu uses them --
//    tests every point of those runs exactly, and (fill pass) sorts the hits by ORIGINAL index in shared memory so that the
//    neighbourhood comes out in model order like the reference's masked indexing.  Cost ~ points inside the ball's bounding
//    rows, not N_model: 10^5 keypoints on a 16 M-point model take 11 ms (centres visited in Morton order of their cells: the
//    model is read from L2) instead of a minute.
//  * models without a grid (k_local_points): order-preserving brute-force compaction -- a warp owns a chunk of 2048
//    consecutive model points and a tile of 32 centres; the hit mask of every (32-point group, centre) is a ballot, running
//    per-centre offsets live one per lane.  Model points are read once per 32 centres.
// FP64 throughout (class double out); the membership test is the same function in both.
#include <math.h>
#include <stdlib.h>
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

// ---- grid path ---------------------------------------------------------------------------------------------------------
constexpr int LPG_THREADS = 256;
constexpr int LPG_MAX_HITS = 32768;         // hits of one centre the fill pass sorts in shared memory (the reference caps at 6000)

struct LocalGridArgs {
    GridView g;
    const ModelPointD* md;
    const double* cx; const double* cy; const double* cz; int64_t nc;
    double R;
    int64_t* totals;                    // count pass: hits per centre
    const int64_t* offsets;             // fill pass: rows of centre k = offsets[k] .. offsets[k+1]-1
    const int32_t* status;              // fill pass: centres with status != 0 are skipped (may be null)
    double* out; int64_t ld_out; double* dists; int32_t* orig;
    int sort_cap;                       // fill pass: power of two >= the largest accepted count
    const int32_t* order;               // block b works on centre order[b]: centres in Morton order of their grid cells, so that
                                        // blocks running at the same time read the same region of the model (L2 reuse)
};

// distance from v to the slab [lo, lo + w] along one axis (0 inside), minus a rounding allowance: never too large
__device__ __forceinline__ double slab_dist(double v, double lo, double w) {
    const double d = fmax(fmax(lo - v, v - (lo + w)), 0.0);
    return fmax(d - 1e-9 * (fabs(v) + fabs(lo) + w), 0.0);
}

template <bool FILL>
__global__ void __launch_bounds__(LPG_THREADS) k_local_grid(const __grid_constant__ LocalGridArgs a) {
    extern __shared__ int32_t s_hits[];                      // FILL: original indices of the hits (sorted at the end)
    __shared__ int32_t s_start[LPG_THREADS];
    __shared__ int32_t s_incl[LPG_THREADS];
    __shared__ int32_t s_warp[LPG_THREADS / 32];
    __shared__ unsigned int s_n;
    const GridView& G = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t k = a.order ? (int64_t)a.order[blockIdx.x] : (int64_t)blockIdx.x;
    int64_t row0 = 0, nk = 0;
    if (FILL) {
        if (a.status && a.status[k] != 0) return;
        row0 = a.offsets[k];
        nk = a.offsets[k + 1] - row0;
        if (nk <= 0) return;
    }
    const double cx = a.cx[k], cy = a.cy[k], cz = a.cz[k];
    const double R = a.R, R2 = R * R, R2lo = R2 * (1.0 - 1e-15), R2hi = R2 * (1.0 + 1e-15);
    const double cell = G.cell;
    // cell span of the ball's bounding cube, one cell of slack, clamped to the grid (a point may sit in the cell next to the
    // one its coordinates suggest: binning rounds)
    int lo[3], hi[3];
    const double cc[3] = {cx, cy, cz};
    bool any = true;
    for (int ax = 0; ax < 3; ++ax) {
        const double l = floor((cc[ax] - R - G.origin[ax]) * G.inv_cell) - 1.0, h = floor((cc[ax] + R - G.origin[ax]) * G.inv_cell) + 1.0;
        lo[ax] = (int)fmax(l, 0.0);
        hi[ax] = (int)fmin(h, (double)(G.dims[0][ax] - 1));
        if (!(l <= (double)(G.dims[0][ax] - 1)) || !(h >= 0.0) || lo[ax] > hi[ax]) any = false;
    }
    if (tid == 0) s_n = 0;
    __syncthreads();
    long long my_hits = 0;
    if (any) {
        const int nyr = hi[1] - lo[1] + 1;
        const int64_t nrows = (int64_t)nyr * (hi[2] - lo[2] + 1);
        const int dx0 = G.dims[0][0], dy0 = G.dims[0][1];
        for (int64_t rb = 0; rb < nrows; rb += LPG_THREADS) {
            // ---- one (y,z) cell row per thread: the run of points of the cells the ball can reach in x ----
            const int64_t r = rb + tid;
            int32_t start = 0, len = 0;
            if (r < nrows) {
                const int iz = lo[2] + (int)(r / nyr), iy = lo[1] + (int)(r % nyr);
                const double dy = slab_dist(cy, G.origin[1] + (double)iy * cell, cell), dz = slab_dist(cz, G.origin[2] + (double)iz * cell, cell);
                const double rem = R2hi - dy * dy - dz * dz;
                if (rem >= 0.0) {
                    const double rx = sqrt(rem) * (1.0 + 1e-9) + 1e-9 * (fabs(cx) + cell);
                    const int xa = max(lo[0], (int)fmax(floor((cx - rx - G.origin[0]) * G.inv_cell) - 1.0, 0.0));
                    const int xb = min(hi[0], (int)fmin(floor((cx + rx - G.origin[0]) * G.inv_cell) + 1.0, (double)(dx0 - 1)));
                    if (xa <= xb) {
                        const int64_t c0 = ((int64_t)iz * dy0 + iy) * dx0;
                        start = G.cell_start[c0 + xa];
                        len = G.cell_start[c0 + xb + 1] - start;
                    }
                }
            }
            // ---- block-wide inclusive scan of the run lengths ----
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < warp; ++w) woff += s_warp[w];
            int total = 0;
            for (int w = 0; w < LPG_THREADS / 32; ++w) total += s_warp[w];
            s_start[tid] = start;
            s_incl[tid] = incl + woff;
            __syncthreads();
            // ---- the points of those runs, flattened: a warp takes a contiguous slice, its lanes consecutive points; a lane finds
            //      the run of its first point by bisection and then only steps forward (32 points further is ~2 runs further) ----
            const int per = ((total + LPG_THREADS - 1) / LPG_THREADS) * 32;
            const int t1 = min((warp + 1) * per, total);
            int t = warp * per + lane;
            if (t < t1) {
                int r = 0, hi_r = LPG_THREADS - 1;                   // first run whose inclusive count exceeds t
                while (r < hi_r) {
                    const int mid = (r + hi_r) >> 1;
                    if (s_incl[mid] > t) hi_r = mid; else r = mid + 1;
                }
                for (; t < t1; t += 32) {
                    while (s_incl[r] <= t) ++r;                      // t < total = s_incl[last]: stops inside the array
                    const int excl = r > 0 ? s_incl[r - 1] : 0;
                    const GridPoint gp = G.pts[s_start[r] + (t - excl)];
                    const double dx = __dsub_rn(gp.x, cx), dy = __dsub_rn(gp.y, cy), dz = __dsub_rn(gp.z, cz);   // pts - c (getLocalPoints.m:23)
                    double dist = 0.0;
                    if (inside(dx, dy, dz, R, R2lo, R2hi, dist)) {
                        if (FILL) {
                            const unsigned at = atomicAdd(&s_n, 1u);
                            if (at < (unsigned)a.sort_cap) s_hits[at] = gp.orig;
                        } else {
                            ++my_hits;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    if (!FILL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_hits += __shfl_xor_sync(0xffffffffu, my_hits, o);
        __shared__ long long s_cnt[LPG_THREADS / 32];
        if (lane == 0) s_cnt[warp] = my_hits;
        __syncthreads();
        if (tid == 0) {
            long long t = 0;
            for (int w = 0; w < LPG_THREADS / 32; ++w) t += s_cnt[w];
            a.totals[k] = t;
        }
        return;
    }
    // ---- fill: sort the hits by original index (bitonic, shared memory), then write them in model order ----
    const int n = (int)min((unsigned)a.sort_cap, s_n);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = n + tid; i < np2; i += LPG_THREADS) s_hits[i] = 0x7fffffff;
    __syncthreads();
    for (int kk = 2; kk <= np2; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += LPG_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const int32_t x = s_hits[i], y = s_hits[l];
                    const bool up = (i & kk) == 0;
                    if ((x > y) == up) { s_hits[i] = y; s_hits[l] = x; }
                }
            }
            __syncthreads();
        }
    for (int j = tid; j < n && j < nk; j += LPG_THREADS) {
        const int32_t o = s_hits[j];
        const ModelPointD p = a.md[o];
        const double dx = __dsub_rn(p.x, cx), dy = __dsub_rn(p.y, cy), dz = __dsub_rn(p.z, cz);
        double dist = 0.0;
        inside(dx, dy, dz, R, R2lo, R2hi, dist);
        const int64_t row = row0 + j;
        a.out[row] = dx; a.out[a.ld_out + row] = dy; a.out[2 * a.ld_out + row] = dz;
        if (a.dists) a.dists[row] = dist;
        if (a.orig) a.orig[row] = o;
    }
}

// the grid path pays when the ball's bounding rows are a small part of the grid
static bool local_grid_worthwhile(const pcreg_model* m, double R) {
    if (!m->has_grid) return false;
    static const bool off = [] { const char* e = getenv("PCREG_LOCAL_GRID"); return e && e[0] == '0'; }();   // tuning switch, read once
    if (off) return false;
    const GridView& G = m->grid;
    const double span = 2.0 * R * G.inv_cell + 3.0;
    return span * span <= 0.25 * (double)G.dims[0][1] * (double)G.dims[0][2] || m->n > 4000000;
}
// Centres in Morton order of their (clamped) grid cells: host-side LSD radix sort of 30-bit codes, uploaded as a permutation.
static void local_grid_order(const pcreg_model* m, const double* centres, int64_t nc, int64_t ld, DevBuf<int32_t>& d_order, cudaStream_t st) {
    const GridView& G = m->grid;
    auto spread = [](uint32_t v) { v &= 0x3ffu; v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u; return v; };
    int shift = 0;
    while ((std::max({G.dims[0][0], G.dims[0][1], G.dims[0][2]}) - 1) >> shift > 1023) ++shift;
    std::vector<uint32_t> key((size_t)nc), key2((size_t)nc);
    std::vector<int32_t> idx((size_t)nc), idx2((size_t)nc);
    for (int64_t k = 0; k < nc; ++k) {
        uint32_t c[3];
        for (int ax = 0; ax < 3; ++ax) {
            const double v = (centres[ax * ld + k] - G.origin[ax]) * G.inv_cell;
            c[ax] = (uint32_t)std::fmin(std::fmax(v, 0.0), (double)(G.dims[0][ax] - 1)) >> shift;       // fmax(NaN, 0) = 0
        }
        key[(size_t)k] = spread(c[0]) | (spread(c[1]) << 1) | (spread(c[2]) << 2);
        idx[(size_t)k] = (int32_t)k;
    }
    std::vector<uint32_t> cnt(1u << 15);
    for (int pass = 0; pass < 2; ++pass) {                  // two 15-bit digits
        const int sh = 15 * pass;
        std::fill(cnt.begin(), cnt.end(), 0u);
        for (int64_t k = 0; k < nc; ++k) ++cnt[(key[(size_t)k] >> sh) & 0x7fffu];
        uint32_t run = 0;
        for (auto& v : cnt) { const uint32_t t = v; v = run; run += t; }
        for (int64_t k = 0; k < nc; ++k) { const uint32_t at = cnt[(key[(size_t)k] >> sh) & 0x7fffu]++; key2[at] = key[(size_t)k]; idx2[at] = idx[(size_t)k]; }
        key.swap(key2); idx.swap(idx2);
    }
    d_order.alloc((size_t)nc);
    PCREG_CUDA(cudaMemcpyAsync(d_order.p, idx.data(), (size_t)nc * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaStreamSynchronize(st));                  // idx is a local
}
static void local_grid_count(const pcreg_model* m, const double* d_c, int64_t nc, double R, int64_t* d_totals, const int32_t* d_order, cudaStream_t st) {
    LocalGridArgs a{};
    a.g = m->grid; a.md = m->md.p; a.cx = d_c; a.cy = d_c + nc; a.cz = d_c + 2 * nc; a.nc = nc; a.R = R; a.totals = d_totals; a.order = d_order;
    k_local_grid<false><<<(unsigned)nc, LPG_THREADS, 0, st>>>(a);
    PCREG_LAUNCHED();
}
static void local_grid_fill(const pcreg_model* m, const double* d_c, int64_t nc, double R, const int64_t* d_offsets, const int32_t* d_status,
                            int64_t max_count, double* d_out, int64_t ld_out, double* d_dists, int32_t* d_orig, const int32_t* d_order, cudaStream_t st) {
    LocalGridArgs a{};
    a.order = d_order;
    a.g = m->grid; a.md = m->md.p; a.cx = d_c; a.cy = d_c + nc; a.cz = d_c + 2 * nc; a.nc = nc; a.R = R;
    a.offsets = d_offsets; a.status = d_status; a.out = d_out; a.ld_out = ld_out; a.dists = d_dists; a.orig = d_orig;
    int cap = 32;
    while (cap < max_count) cap <<= 1;
    a.sort_cap = cap;
    const size_t dyn = (size_t)cap * sizeof(int32_t);
    PCREG_CUDA(cudaFuncSetAttribute(k_local_grid<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_local_grid<true><<<(unsigned)nc, LPG_THREADS, dyn, st>>>(a);
    PCREG_LAUNCHED();
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
    DevBuf<int64_t> totals((size_t)nc);
    // ---- models with a grid: walk the ball's cell rows instead of the whole model ----
    int64_t max_count = 0;
    if (fill) for (int64_t k = 0; k < nc; ++k) if (!status_in || status_in[k] == 0) max_count = std::max(max_count, offsets[k + 1] - offsets[k]);
    if (local_grid_worthwhile(m, R) && (!fill || max_count <= LPG_MAX_HITS)) {
        DevBuf<int32_t> d_order;
        if (nc >= 1024) local_grid_order(m, centres, nc, ld, d_order, st);
        if (!fill) {
            local_grid_count(m, dc.p, nc, R, totals.p, d_order.p, st);
            PCREG_CUDA(cudaMemcpyAsync(counts_out, totals.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
            PCREG_CUDA(cudaStreamSynchronize(st));
            return;
        }
        const int64_t ntotal = offsets[nc];
        PCREG_REQUIRE(ntotal >= 0 && ld_out >= ntotal, "local_points_fill: bad offsets / ld_out");
        const size_t nel = (size_t)std::max<int64_t>(ntotal, 1);
        DevBuf<int64_t> d_off((size_t)nc + 1);
        DevBuf<int32_t> d_st(status_in ? (size_t)nc : 0), d_orig(orig ? nel : 0);
        DevBuf<double> d_out(nel * 3), d_dist(dists ? nel : 0);
        PCREG_CUDA(cudaMemcpyAsync(d_off.p, offsets, ((size_t)nc + 1) * 8, cudaMemcpyHostToDevice, st));
        if (status_in) PCREG_CUDA(cudaMemcpyAsync(d_st.p, status_in, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
        local_grid_fill(m, dc.p, nc, R, d_off.p, status_in ? d_st.p : nullptr, max_count, d_out.p, (int64_t)nel, dists ? d_dist.p : nullptr,
                        orig ? d_orig.p : nullptr, d_order.p, st);
        for (int k = 0; k < 3; ++k)
            PCREG_CUDA(cudaMemcpyAsync(pts_rel + (size_t)k * ld_out, d_out.p + k * nel, (size_t)ntotal * 8, cudaMemcpyDeviceToHost, st));
        if (dists) PCREG_CUDA(cudaMemcpyAsync(dists, d_dist.p, (size_t)ntotal * 8, cudaMemcpyDeviceToHost, st));
        if (orig) PCREG_CUDA(cudaMemcpyAsync(orig, d_orig.p, (size_t)ntotal * 4, cudaMemcpyDeviceToHost, st));
        PCREG_CUDA(cudaStreamSynchronize(st));
        return;
    }
    DevBuf<int32_t> cnt((size_t)ntiles * LP_TILE * nchunks);
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
    DevBuf<int64_t> totals((size_t)nc);
    const bool use_grid = local_grid_worthwhile(m, R);
    DevBuf<int32_t> cnt(use_grid ? 0 : (size_t)ntiles * LP_TILE * nchunks);
    LocalArgs a{};
    a.md = m->md.p; a.n = m->n; a.cx = dc.p; a.cy = dc.p + nc; a.cz = dc.p + 2 * nc; a.nc = nc; a.R = R;
    a.chunk_cnt = cnt.p; a.nchunks = nchunks;
    dim3 grid((unsigned)nchunks, (unsigned)ntiles);
    DevBuf<int32_t> d_order;
    if (use_grid && nc >= 1024) local_grid_order(m, centres, nc, ld, d_order, st);
    if (use_grid) {
        local_grid_count(m, dc.p, nc, R, totals.p, d_order.p, st);
    } else {
        k_local_points<false><<<grid, 32, 0, st>>>(a);
        PCREG_LAUNCHED();
        k_local_scan<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(cnt.p, nchunks, nc, nullptr, nullptr, nullptr, totals.p);
        PCREG_LAUNCHED();
    }
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
    DevBuf<int32_t> d_st((size_t)nc);
    PCREG_CUDA(cudaMemcpyAsync(out.d_offsets.p, out.offsets.data(), ((size_t)nc + 1) * 8, cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemcpyAsync(d_st.p, out.status.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    int64_t max_count = 0;
    for (int64_t k = 0; k < nc; ++k) if (!out.status[k]) max_count = std::max(max_count, out.counts[k]);
    if (use_grid && max_count <= LPG_MAX_HITS) {
        local_grid_fill(m, dc.p, nc, R, out.d_offsets.p, d_st.p, max_count, out.pts.p, out.nel, nullptr, nullptr, d_order.p, st);
        PCREG_CUDA(cudaStreamSynchronize(st));
        return;
    }
    if (use_grid) {                                         // counts came from the grid pass: the brute fill needs its own chunk counts
        cnt.alloc((size_t)ntiles * LP_TILE * nchunks);
        a.chunk_cnt = cnt.p;
        k_local_points<false><<<grid, 32, 0, st>>>(a);
        PCREG_LAUNCHED();
    }
    DevBuf<int64_t> base((size_t)nc * nchunks);
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
