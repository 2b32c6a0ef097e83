// nn_grid.cu -- exact nearest neighbour on the uniform grid + occupancy pyramid (large models).
//
// Semantics identical to nn_brute.cu (knnsearch K=1, FP64, ties -> smallest original index); the
// reference's nearest analogue of a spatially pruned search is speedyDescriptors.m:44-60 (boxes with a
// halo) and getLocalPoints.m:8-15 (cube pre-filter).
//
// Two kernels per NN pass, one thread per query in both:
//   k_nn_grid_direct  warm-started queries whose ball (radius = distance to the previous iteration's
//                     correspondence) has a bounding cube of at most 4 x 4 x 4 level-0 cells: the <= 16
//                     (y,z) cell rows are contiguous point runs, fetched with independent loads and
//                     scanned.  Everything else is appended to a work list (warp-aggregated atomics).
//   k_nn_grid_walk    branch-and-bound walk of the occupancy pyramid with a small explicit stack for the
//                     work list (far queries, and every query of the first iteration).  Keeping the two
//                     populations in separate launches keeps the lanes of a warp on similar work.
// Pruning tests run in FP32 in CELL UNITS with a conservative slop (a cell is skipped only when a
// guaranteed LOWER bound of its distance exceeds a guaranteed UPPER bound of the best distance); every
// visited point is evaluated in FP64 with the oracle's formula and operation order and competes on
// (d2, original index) -- the answer equals the FP64 brute-force answer bit for bit.
#include <math.h>
#include <float.h>
#include <algorithm>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"

namespace pcreg {

struct GridArgs {
    GridView g;
    const ModelPointD* md;
    const double* sx; const double* sy; const double* sz; int64_t ns;
    const double* T; int64_t nq;
    const int32_t* prev;
    int32_t* idx; double* d2;
    float* lb2;                     // [nq] in/out or null: lower bound of the distance from the query to every model point
                                    //      OTHER than its correspondence (temporal-coherence certificate, see below)
    const float* delta;             // [nhyp] or null: upper bound of how far the last pose update moved any source point
    float gap_cells;                // the row scan is exhaustive within (best distance + gap), in cell units
    int32_t* worklist;              // [nq] query ids for the walk kernel (direct kernel appends)
    unsigned int* work_count;       // number of entries in worklist
    unsigned long long* counters;   // [0] points visited, [1] leaf cells / rows visited, [2] nodes popped (may be null)
};

constexpr int GRID_STACK = 80;
constexpr float GRID_SLOP_ABS = 1.3e-4f;   // cell units: FP32 rounding of the query (<= 1024 cells) + of the difference
constexpr float GRID_SLOP_REL = 2.5e-7f;

// stack entry: [63:34] lower bound (positive float, lowest mantissa bit dropped = rounded down),
//              [33:30] level, [29:20] z, [19:10] y, [9:0] x
__device__ __forceinline__ unsigned long long pack_entry(float lb, int level, int x, int y, int z) {
    return ((unsigned long long)(__float_as_uint(lb) >> 1) << 34) | ((unsigned long long)(unsigned)level << 30) |
           ((unsigned long long)(unsigned)z << 20) | ((unsigned long long)(unsigned)y << 10) | (unsigned long long)(unsigned)x;
}

// conservative (never too large) distance along one axis, in cell units, from q to the slab [lo, lo+edge]
__device__ __forceinline__ float axis_lb(float q, float lo, float edge) {
    const float d = fmaxf(lo - q, q - (lo + edge));
    return fmaxf(fmaf(-GRID_SLOP_REL, fabsf(d), d) - GRID_SLOP_ABS, 0.f);
}
// conservative squared distance, in cell units, from the query to the box [x, x+1] * edge (corners exact in FP32)
__device__ __forceinline__ float box_lb(float qx, float qy, float qz, int x, int y, int z, float edge) {
    const float dx = axis_lb(qx, (float)x * edge, edge), dy = axis_lb(qy, (float)y * edge, edge), dz = axis_lb(qz, (float)z * edge, edge);
    return (dx * dx + dy * dy + dz * dz) * (1.f - 6e-7f);
}
// guaranteed upper bound of best (squared distance) in squared cell units, as a float
__device__ __forceinline__ float best_ub_cells(double best, double inv_cell2) {
    const double b = best * inv_cell2;
    return (b < 3.0e38) ? __double2float_ru(b) * (1.f + 6e-7f) : FLT_MAX;
}
// same for the ball of radius (sqrt(best) + gap): everything inside it gets visited
__device__ __forceinline__ float best_ub_cells_gap(double best, double inv_cell, float gap_cells) {
    const double r = sqrt(best) * inv_cell + (double)gap_cells;
    const double b = r * r;
    return (b < 3.0e38) ? __double2float_ru(b) * (1.f + 6e-7f) : FLT_MAX;
}

struct Query {
    double qx, qy, qz;      // FP64 query (oracle order)
    float fx, fy, fz;       // in cell units, FP32 (pruning only)
    double best; int32_t bidx; float bestc;
    double second;          // smallest exact d2 among visited points other than the current best
    int ilx, ihx, ily, ihy, ilz, ihz;   // level-0 cell span of the ball's bounding cube (valid when has_span)
    bool has_span;
};

__device__ __forceinline__ void setup_query(const GridArgs& a, int64_t gq, Query& Q) {
    const GridView& G = a.g;
    const unsigned h = (unsigned)gq / (unsigned)a.ns, i = (unsigned)gq - h * (unsigned)a.ns;      // nq < 2^31 (launcher)
    quick_tf(a.T + (size_t)h * 16, a.sx[i], a.sy[i], a.sz[i], Q.qx, Q.qy, Q.qz);
    Q.fx = __double2float_rn((Q.qx - G.origin[0]) * G.inv_cell);
    Q.fy = __double2float_rn((Q.qy - G.origin[1]) * G.inv_cell);
    Q.fz = __double2float_rn((Q.qz - G.origin[2]) * G.inv_cell);
    Q.best = INFINITY;
    Q.bidx = -1;
    if (a.prev) {
        const int32_t p = a.prev[gq];
        if (p >= 0) {
            const ModelPointD mp = a.md[p];
            Q.best = dist2_exact(mp.x, mp.y, mp.z, Q.qx, Q.qy, Q.qz);
            Q.bidx = p;
        }
    }
    Q.bestc = (a.gap_cells > 0.f) ? best_ub_cells_gap(Q.best, G.inv_cell, a.gap_cells) : best_ub_cells(Q.best, G.inv_cell * G.inv_cell);
    Q.second = INFINITY;
    Q.has_span = false;
    if (Q.bidx >= 0 && Q.bestc < 1.0e12f) {
        // ball radius in cells (upper bound) -> integer cell span
        const float rc = __fsqrt_ru(Q.bestc) + 2.0f * GRID_SLOP_ABS + GRID_SLOP_REL * (fabsf(Q.fx) + fabsf(Q.fy) + fabsf(Q.fz));
        const float big = 1.0e6f;
        Q.ilx = (int)floorf(fmaxf(fminf(Q.fx - rc, big), -big)); Q.ihx = (int)floorf(fmaxf(fminf(Q.fx + rc, big), -big));
        Q.ily = (int)floorf(fmaxf(fminf(Q.fy - rc, big), -big)); Q.ihy = (int)floorf(fmaxf(fminf(Q.fy + rc, big), -big));
        Q.ilz = (int)floorf(fmaxf(fminf(Q.fz - rc, big), -big)); Q.ihz = (int)floorf(fmaxf(fminf(Q.fz + rc, big), -big));
        Q.has_span = true;
    }
}

__device__ __forceinline__ void scan_points(const GridView& G, int32_t s0, int32_t s1, Query& Q, double inv_cell2) {
    bool improved = false;
    for (int32_t p = s0; p < s1; ++p) {
        const GridPoint gp = G.pts[p];
        const double d = dist2_exact(gp.x, gp.y, gp.z, Q.qx, Q.qy, Q.qz);
        if (d < Q.best || (d == Q.best && gp.orig < Q.bidx)) { Q.best = d; Q.bidx = gp.orig; improved = true; }
    }
    if (improved) Q.bestc = best_ub_cells(Q.best, inv_cell2);
}

__device__ __forceinline__ void flush_counters(unsigned long long* counters, unsigned long long n_pts,
                                               unsigned long long n_cells, unsigned long long n_nodes) {
    if (!counters) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_pts += __shfl_xor_sync(0xffffffffu, n_pts, o);
        n_cells += __shfl_xor_sync(0xffffffffu, n_cells, o);
        n_nodes += __shfl_xor_sync(0xffffffffu, n_nodes, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_pts) atomicAdd(&counters[0], n_pts);
        if (n_cells) atomicAdd(&counters[1], n_cells);
        if (n_nodes) atomicAdd(&counters[2], n_nodes);
    }
}

// ---- kernel 1: row scan ----------------------------------------------------------------------------------
// Cells of one x-row are contiguous in memory, so for a fixed (y,z) the cells the ball can touch are ONE
// run of points [cell_start[row + xa], cell_start[row + xb + 1]).  A query whose bounding cube spans at
// most GRID_ROW_SPAN cells in y and z walks its rows.
//
// The cost per query varies by more than 10x (ball radius), so a fixed thread <-> query mapping leaves most
// lanes of a warp idle behind its slowest query.  Instead each WARP owns a contiguous range of queries and
// every lane runs a small state machine: one trip of the loop is either "process one point" or "advance one
// row"; a lane that finishes its query takes the next one of the warp's range.  Fetching (pose transform,
// warm-start bound) is batched: it runs only when >= GRID_FETCH_BATCH lanes are idle, so it executes with
// many active lanes too.
constexpr int GRID_ROW_SPAN = 8;
constexpr int GRID_FETCH_BATCH = 8;

template <bool COH>
__global__ void __launch_bounds__(128, 8) k_nn_grid_direct(const __grid_constant__ GridArgs a) {
    const GridView& G = a.g;
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per_warp = (a.nq + nwarps - 1) / nwarps;
    int64_t next = warp_id * per_warp;                       // warp-uniform cursor into this warp's range
    const int64_t end = min(a.nq, next + per_warp);
    const double inv_cell2 = G.inv_cell * G.inv_cell;
    const int dx0 = G.dims[0][0], dy0 = G.dims[0][1], dz0 = G.dims[0][2];
    unsigned long long n_pts = 0, n_cells = 0;

    Query Q;
    int64_t gq = -1;
    bool have = false;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0, z1 = 0, y = 0, z = 0;
    int32_t p = 0, e = 0;
    float dz2 = 0.f;

    while (true) {
        // ---- batched fetch ----
        const unsigned idle = __ballot_sync(0xffffffffu, !have);
        const int nidle = __popc(idle);
        if (next < end && (nidle >= GRID_FETCH_BATCH || nidle == 32 || (idle && end - next <= 0))) {
            bool defer = false;
            if (!have) {
                const int64_t cand = next + __popc(idle & ((1u << lane) - 1u));
                if (cand < end) {
                    gq = cand;
                    setup_query(a, gq, Q);
                    // temporal-coherence certificate: every other model point was at distance >= lb2 before the
                    // last pose update, which moved this query by at most delta; if the old correspondence is
                    // now strictly closer than lb2 - delta it is still the unique nearest neighbour.
                    bool certified = false;
                    if (COH && a.lb2 && Q.bidx >= 0) {
                        const float lb_new = __fsub_rd(a.lb2[gq], a.delta[(unsigned)gq / (unsigned)a.ns]);
                        const float d1 = __double2float_ru(sqrt(Q.best)) * (1.f + 1e-6f) + 1e-30f;
                        if (d1 < lb_new * (1.f - 1e-6f)) {
                            certified = true;
                            a.lb2[gq] = lb_new;
                            a.idx[gq] = Q.bidx;
                            if (a.d2) a.d2[gq] = Q.best;
                            if (a.counters) atomicAdd(&a.counters[3], 1ull);
                        }
                    }
                    if (certified) {
                        // nothing to scan
                    } else if (Q.has_span && Q.ihy - Q.ily < GRID_ROW_SPAN && Q.ihz - Q.ilz < GRID_ROW_SPAN && Q.ihx - Q.ilx < 4 * GRID_ROW_SPAN) {
                        x0 = max(Q.ilx, 0); x1 = min(Q.ihx, dx0 - 1);
                        y0 = max(Q.ily, 0); y1 = min(Q.ihy, dy0 - 1);
                        z = max(Q.ilz, 0); z1 = min(Q.ihz, dz0 - 1);
                        y = y0;
                        p = 0; e = 0;
                        if (x0 > x1 || y0 > y1 || z > z1) z = z1 + 1;       // nothing to scan: the warm start stands
                        const float t = axis_lb(Q.fz, (float)z, 1.f);
                        dz2 = t * t;
                        have = true;
                    } else {
                        defer = true;
                    }
                }
            }
            next += nidle;
            // deferred queries go to the walk kernel's work list (warp-aggregated append)
            const unsigned dm = __ballot_sync(0xffffffffu, defer);
            if (dm) {
                const int leader = __ffs(dm) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(a.work_count, (unsigned)__popc(dm));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (defer) a.worklist[base + __popc(dm & ((1u << lane) - 1u))] = (int32_t)gq;
            }
            continue;
        }
        if (nidle == 32) break;                              // nothing in flight and nothing left to fetch
        // ---- one trip: lanes out of points advance one row, then every lane with points processes one ----
        if (have && p >= e) {
            if (z > z1) {                                    // rows exhausted: done with this query
                a.idx[gq] = Q.bidx;
                if (a.d2) a.d2[gq] = Q.best;
                if (COH && a.lb2) {
                    // every point within (best distance + gap) was visited: anything else is at least that far
                    const double r1 = sqrt(Q.best) + (double)a.gap_cells * G.cell;
                    const double r2 = sqrt(Q.second);
                    a.lb2[gq] = __double2float_rd(fmin(r1, r2) * (1.0 - 1e-9));
                }
                have = false;
            } else {                                         // ROW step
                float dy2 = axis_lb(Q.fy, (float)y, 1.f);
                dy2 *= dy2;
                const float lb = (dy2 + dz2) * (1.f - 6e-7f);
                if (lb <= Q.bestc) {
                    // cells of this row the ball can reach: |x - fx| <= sqrt(bestc - lb) (+ slop)
                    const float rx = __fsqrt_ru(Q.bestc - lb) * (1.f + 1e-6f) + 2.f * GRID_SLOP_ABS + GRID_SLOP_REL * fabsf(Q.fx);
                    const int xa = max(x0, (int)floorf(Q.fx - rx)), xb = min(x1, (int)floorf(Q.fx + rx));
                    if (xa <= xb) {
                        const int64_t c0 = ((int64_t)z * dy0 + y) * dx0;
                        p = G.cell_start[c0 + xa];
                        e = G.cell_start[c0 + xb + 1];
                        if (e > p) { ++n_cells; n_pts += (unsigned long long)(e - p); }
                    }
                }
                if (++y > y1) {
                    y = y0; ++z;
                    const float t = axis_lb(Q.fz, (float)z, 1.f);
                    dz2 = t * t;
                }
            }
        }
        if (have && p < e) {                                 // POINT step
            const GridPoint gp = G.pts[p];
            ++p;
            const double d = dist2_exact(gp.x, gp.y, gp.z, Q.qx, Q.qy, Q.qz);
            if (COH) {
                if (gp.orig != Q.bidx) {
                    if (d < Q.best || (d == Q.best && gp.orig < Q.bidx)) {
                        Q.second = Q.best;                   // the old best becomes the runner-up
                        Q.best = d; Q.bidx = gp.orig;
                        Q.bestc = best_ub_cells_gap(Q.best, G.inv_cell, a.gap_cells);
                    } else if (d < Q.second) {
                        Q.second = d;
                    }
                }
            } else if (d < Q.best || (d == Q.best && gp.orig < Q.bidx)) {
                Q.best = d; Q.bidx = gp.orig;
                Q.bestc = best_ub_cells(Q.best, inv_cell2);
            }
        }
    }
    flush_counters(a.counters, n_pts, n_cells, 0);
}

// ---- kernel 2: pyramid walk over the work list (worklist == nullptr: every query) ---------------------------
__global__ void __launch_bounds__(128) k_nn_grid_walk(const __grid_constant__ GridArgs a) {
    const GridView& G = a.g;
    const int64_t count = a.worklist ? (int64_t)*a.work_count : a.nq;
    const double inv_cell2 = G.inv_cell * G.inv_cell;
    unsigned long long n_pts = 0, n_cells = 0, n_nodes = 0;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < count; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gq = a.worklist ? (int64_t)a.worklist[w] : w;
        Query Q;
        setup_query(a, gq, Q);
        const float fx = Q.fx, fy = Q.fy, fz = Q.fz;
        unsigned long long stack[GRID_STACK];
        int sp = 0;
        const int top = G.nlevels - 1;
        bool from_root = true;
        if (Q.has_span) {
            // lowest level whose <= 2 x 2 x 2 nodes cover the ball's bounding cube
            int l = 0;
            while (l < top && (((Q.ihx >> l) - (Q.ilx >> l)) > 1 || ((Q.ihy >> l) - (Q.ily >> l)) > 1 || ((Q.ihz >> l) - (Q.ilz >> l)) > 1)) ++l;
            if (((Q.ihx >> l) - (Q.ilx >> l)) <= 1 && ((Q.ihy >> l) - (Q.ily >> l)) <= 1 && ((Q.ihz >> l) - (Q.ilz >> l)) <= 1) {
                from_root = false;
                const float edge = (float)(1 << l);
                const int dxl = G.dims[l][0], dyl = G.dims[l][1], dzl = G.dims[l][2];
                for (int z = Q.ilz >> l; z <= (Q.ihz >> l); ++z) {
                    if (z < 0 || z >= dzl) continue;
                    for (int y = Q.ily >> l; y <= (Q.ihy >> l); ++y) {
                        if (y < 0 || y >= dyl) continue;
                        for (int x = Q.ilx >> l; x <= (Q.ihx >> l); ++x) {
                            if (x < 0 || x >= dxl) continue;
                            if (l > 0 && G.mask[l][((int64_t)z * dyl + y) * dxl + x] == 0) continue;
                            const float lb = box_lb(fx, fy, fz, x, y, z, edge);
                            if (lb <= Q.bestc) stack[sp++] = pack_entry(lb, l, x, y, z);
                        }
                    }
                }
            }
        }
        if (from_root) stack[sp++] = pack_entry(0.f, top, 0, 0, 0);

        while (true) {
            int32_t s0 = 0, s1 = 0;
            bool have_leaf = false;
            // phase 1: pop / expand until a leaf cell is in hand
            while (sp > 0) {
                const unsigned long long e = stack[--sp];
                const float lbf = __uint_as_float((unsigned)(e >> 34) << 1);
                if (lbf > Q.bestc) continue;
                const unsigned lo32 = (unsigned)e;
                const int level = (int)((e >> 30) & 0xF);
                const int ix = (int)(lo32 & 1023u), iy = (int)((lo32 >> 10) & 1023u), iz = (int)((lo32 >> 20) & 1023u);
                ++n_nodes;
                if (level == 0) {
                    const int64_t c = ((int64_t)iz * G.dims[0][1] + iy) * G.dims[0][0] + ix;
                    s0 = G.cell_start[c];
                    s1 = G.cell_start[c + 1];
                    have_leaf = true;
                    break;
                }
                const int64_t c = ((int64_t)iz * G.dims[level][1] + iy) * G.dims[level][0] + ix;
                unsigned m = G.mask[level][c];
                const float edge = (float)(1 << (level - 1));              // child edge in cells
                const float cx = (float)(2 * ix + 1) * edge, cy = (float)(2 * iy + 1) * edge, cz = (float)(2 * iz + 1) * edge;
                const unsigned oct = (fx >= cx ? 1u : 0u) | (fy >= cy ? 2u : 0u) | (fz >= cz ? 4u : 0u);
                // permute the mask so that bit t <-> child (t ^ oct); then high bits = far octants
                if (oct & 1u) m = ((m & 0xAAu) >> 1) | ((m & 0x55u) << 1);
                if (oct & 2u) m = ((m & 0xCCu) >> 2) | ((m & 0x33u) << 2);
                if (oct & 4u) m = ((m & 0xF0u) >> 4) | ((m & 0x0Fu) << 4);
                while (m) {
                    const int t = 31 - __clz(m);                           // far first, so the near octant is popped first
                    m &= ~(1u << t);
                    const int k = t ^ (int)oct;
                    const int x = 2 * ix + (k & 1), y = 2 * iy + ((k >> 1) & 1), z = 2 * iz + (k >> 2);
                    const float lb = box_lb(fx, fy, fz, x, y, z, edge);
                    if (lb <= Q.bestc && sp < GRID_STACK) stack[sp++] = pack_entry(lb, level - 1, x, y, z);
                }
            }
            if (!have_leaf) break;
            // phase 2: exact FP64 scan of the leaf's points
            ++n_cells;
            n_pts += (unsigned long long)(s1 - s0);
            scan_points(G, s0, s1, Q, inv_cell2);
        }
        a.idx[gq] = Q.bidx;
        if (a.d2) a.d2[gq] = Q.best;
        if (a.lb2) a.lb2[gq] = 0.f;                          // the walk gives no exhaustive-radius guarantee
    }
    flush_counters(a.counters, n_pts, n_cells, n_nodes);
}

void nn_grid_launch(const pcreg_model* m, const double* d_sx, const double* d_sy, const double* d_sz, int64_t ns,
                    const double* d_T, int64_t nhyp, const int32_t* d_prev, int32_t* d_idx, double* d_d2,
                    unsigned long long* d_counters, GridScratch& sc, float* d_lb2, const float* d_delta, cudaStream_t st) {
    PCREG_REQUIRE(m->has_grid, "grid NN requested but the model was created without build_grid");
    GridArgs a{};
    a.g = m->grid; a.md = m->md.p;
    a.sx = d_sx; a.sy = d_sy; a.sz = d_sz; a.ns = ns; a.T = d_T; a.nq = nhyp * ns;
    a.prev = d_prev; a.idx = d_idx; a.d2 = d_d2; a.counters = d_counters;
    a.lb2 = d_lb2; a.delta = d_delta;
    a.gap_cells = d_lb2 ? 0.15f : 0.f;
    PCREG_REQUIRE(a.nq > 0, "nn_grid: no queries");
    PCREG_REQUIRE(a.nq < 2147483647LL, "nn_grid: too many queries in one launch");
    const int64_t blocks = (a.nq + 127) / 128;
    const int walk_blocks = (int)std::min<int64_t>(blocks, (int64_t)ctx().sm_count * 64);
    if (d_prev) {
        if (sc.worklist.n < (size_t)a.nq) sc.worklist.alloc((size_t)a.nq);
        if (sc.count.n < 1) sc.count.alloc(1);
        a.worklist = sc.worklist.p; a.work_count = sc.count.p;
        PCREG_CUDA(cudaMemsetAsync(sc.count.p, 0, sizeof(unsigned int), st));
        const int direct_blocks = (int)std::min<int64_t>(blocks, (int64_t)ctx().sm_count * 16);
        if (d_lb2) k_nn_grid_direct<true><<<direct_blocks, 128, 0, st>>>(a);
        else       k_nn_grid_direct<false><<<direct_blocks, 128, 0, st>>>(a);
        PCREG_LAUNCHED();
        k_nn_grid_walk<<<walk_blocks, 128, 0, st>>>(a);
        PCREG_LAUNCHED();
    } else {
        a.worklist = nullptr; a.work_count = nullptr;
        k_nn_grid_walk<<<walk_blocks, 128, 0, st>>>(a);
        PCREG_LAUNCHED();
    }
}

}  // namespace pcreg
