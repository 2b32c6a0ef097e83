// nn_grid.cu -- exact nearest neighbour on the uniform grid + occupancy pyramid (large models).
//
// Semantics identical to nn_brute.cu (knnsearch K=1, FP64, ties -> smallest original index); the
// reference's nearest analogue of a spatially pruned search is speedyDescriptors.m:44-60 (boxes with a
// halo) and getLocalPoints.m:8-15 (cube pre-filter).  One thread per query walks the pyramid with a
// small explicit stack (branch and bound):
//   * pruning tests run in FP32 in CELL UNITS with a conservative slop (a cell is skipped only when a
//     guaranteed LOWER bound of its distance exceeds a guaranteed UPPER bound of the best distance);
//   * every visited point is evaluated in FP64 with the oracle's formula and operation order, and
//     competes on (d2, original index) -- the answer equals the FP64 brute-force answer bit for bit;
//   * the previous ICP iteration's correspondence seeds the bound, and the walk then starts at the
//     lowest pyramid level whose <= 2x2x2 nodes cover the ball (no descent from the root);
//   * "while-while" traversal: all lanes pop/expand until they hold a leaf, then all scan points.
#include <math.h>
#include <float.h>

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
    unsigned long long* counters;   // [0] points visited, [1] leaf cells visited, [2] nodes popped (may be null)
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

// conservative (never too large) squared distance, in cell units, from the query to the box
// [x, x+1] * 2^level  (all box corners are exact in FP32)
__device__ __forceinline__ float box_lb(float qx, float qy, float qz, int x, int y, int z, float edge) {
    const float lx = (float)x * edge, ly = (float)y * edge, lz = (float)z * edge;
    float dx = fmaxf(lx - qx, qx - (lx + edge));
    float dy = fmaxf(ly - qy, qy - (ly + edge));
    float dz = fmaxf(lz - qz, qz - (lz + edge));
    dx = fmaxf(fmaf(-GRID_SLOP_REL, fabsf(dx), dx) - GRID_SLOP_ABS, 0.f);
    dy = fmaxf(fmaf(-GRID_SLOP_REL, fabsf(dy), dy) - GRID_SLOP_ABS, 0.f);
    dz = fmaxf(fmaf(-GRID_SLOP_REL, fabsf(dz), dz) - GRID_SLOP_ABS, 0.f);
    return (dx * dx + dy * dy + dz * dz) * (1.f - 6e-7f);
}

__device__ __forceinline__ float best_ub_cells(double best, double inv_cell2) {
    // guaranteed upper bound of best (squared distance) in squared cell units, as a float
    const double b = best * inv_cell2;
    return (b < 3.0e38) ? __double2float_ru(b) * (1.f + 6e-7f) : FLT_MAX;
}

__global__ void __launch_bounds__(128) k_nn_grid(const __grid_constant__ GridArgs a) {
    const int64_t gq = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_pts = 0, n_cells = 0, n_nodes = 0;
    if (gq < a.nq) {
        const GridView& G = a.g;
        const int64_t h = gq / a.ns, i = gq - h * a.ns;
        double qx, qy, qz;
        quick_tf(a.T + h * 16, a.sx[i], a.sy[i], a.sz[i], qx, qy, qz);
        const double inv_cell2 = G.inv_cell * G.inv_cell;
        // query in cell units (FP32 for the pruning tests only)
        const float fx = __double2float_rn((qx - G.origin[0]) * G.inv_cell);
        const float fy = __double2float_rn((qy - G.origin[1]) * G.inv_cell);
        const float fz = __double2float_rn((qz - G.origin[2]) * G.inv_cell);
        double best = INFINITY;
        int32_t bidx = -1;
        if (a.prev) {
            const int32_t p = a.prev[gq];
            if (p >= 0) {
                const ModelPointD mp = a.md[p];
                best = dist2_exact(mp.x, mp.y, mp.z, qx, qy, qz);
                bidx = p;
            }
        }
        float bestc = best_ub_cells(best, inv_cell2);

        unsigned long long stack[GRID_STACK];
        int sp = 0;
        const int top = G.nlevels - 1;
        bool from_root = true;
        if (bidx >= 0 && bestc < 1.0e12f) {
            // ball radius in cells (upper bound) -> integer cell span -> lowest level with <= 2 nodes per axis
            const float rc = __fsqrt_ru(bestc) + 2.0f * GRID_SLOP_ABS + GRID_SLOP_REL * (fabsf(fx) + fabsf(fy) + fabsf(fz));
            const float big = 1.0e6f;
            const int ilx = (int)floorf(fmaxf(fminf(fx - rc, big), -big)), ihx = (int)floorf(fmaxf(fminf(fx + rc, big), -big));
            const int ily = (int)floorf(fmaxf(fminf(fy - rc, big), -big)), ihy = (int)floorf(fmaxf(fminf(fy + rc, big), -big));
            const int ilz = (int)floorf(fmaxf(fminf(fz - rc, big), -big)), ihz = (int)floorf(fmaxf(fminf(fz + rc, big), -big));
            int l = 0;
            while (l < top && (((ihx >> l) - (ilx >> l)) > 1 || ((ihy >> l) - (ily >> l)) > 1 || ((ihz >> l) - (ilz >> l)) > 1)) ++l;
            if (((ihx >> l) - (ilx >> l)) <= 1 && ((ihy >> l) - (ily >> l)) <= 1 && ((ihz >> l) - (ilz >> l)) <= 1) {
                from_root = false;
                const float edge = (float)(1 << l);
                const int dxl = G.dims[l][0], dyl = G.dims[l][1], dzl = G.dims[l][2];
                for (int z = ilz >> l; z <= (ihz >> l); ++z) {
                    if (z < 0 || z >= dzl) continue;
                    for (int y = ily >> l; y <= (ihy >> l); ++y) {
                        if (y < 0 || y >= dyl) continue;
                        for (int x = ilx >> l; x <= (ihx >> l); ++x) {
                            if (x < 0 || x >= dxl) continue;
                            if (l > 0 && G.mask[l][((int64_t)z * dyl + y) * dxl + x] == 0) continue;
                            const float lb = box_lb(fx, fy, fz, x, y, z, edge);
                            if (lb <= bestc) stack[sp++] = pack_entry(lb, l, x, y, z);
                        }
                    }
                }
            }
        }
        if (from_root) stack[sp++] = pack_entry(0.f, top, 0, 0, 0);

        while (true) {
            int32_t s0 = 0, s1 = 0;
            bool have_leaf = false;
            // ---- phase 1: pop / expand until a leaf cell is in hand ----
            while (sp > 0) {
                const unsigned long long e = stack[--sp];
                const float lbf = __uint_as_float((unsigned)(e >> 34) << 1);
                if (lbf > bestc) continue;
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
                    if (lb <= bestc && sp < GRID_STACK) stack[sp++] = pack_entry(lb, level - 1, x, y, z);
                }
            }
            if (!have_leaf) break;
            // ---- phase 2: exact FP64 scan of the leaf's points ----
            ++n_cells;
            n_pts += (unsigned long long)(s1 - s0);
            bool improved = false;
            for (int32_t p = s0; p < s1; ++p) {
                const GridPoint gp = G.pts[p];
                const double d = dist2_exact(gp.x, gp.y, gp.z, qx, qy, qz);
                if (d < best || (d == best && gp.orig < bidx)) { best = d; bidx = gp.orig; improved = true; }
            }
            if (improved) bestc = best_ub_cells(best, inv_cell2);
        }
        a.idx[gq] = bidx;
        if (a.d2) a.d2[gq] = best;
    }
    if (a.counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_pts += __shfl_xor_sync(0xffffffffu, n_pts, o);
            n_cells += __shfl_xor_sync(0xffffffffu, n_cells, o);
            n_nodes += __shfl_xor_sync(0xffffffffu, n_nodes, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&a.counters[0], n_pts);
            atomicAdd(&a.counters[1], n_cells);
            atomicAdd(&a.counters[2], n_nodes);
        }
    }
}

void nn_grid_launch(const pcreg_model* m, const double* d_sx, const double* d_sy, const double* d_sz, int64_t ns,
                    const double* d_T, int64_t nhyp, const int32_t* d_prev, int32_t* d_idx, double* d_d2,
                    unsigned long long* d_counters, cudaStream_t st) {
    PCREG_REQUIRE(m->has_grid, "grid NN requested but the model was created without build_grid");
    GridArgs a{};
    a.g = m->grid; a.md = m->md.p;
    a.sx = d_sx; a.sy = d_sy; a.sz = d_sz; a.ns = ns; a.T = d_T; a.nq = nhyp * ns;
    a.prev = d_prev; a.idx = d_idx; a.d2 = d_d2; a.counters = d_counters;
    PCREG_REQUIRE(a.nq > 0, "nn_grid: no queries");
    const int64_t blocks = (a.nq + 127) / 128;
    PCREG_REQUIRE(blocks < 2147483647LL, "nn_grid: too many queries in one launch");
    k_nn_grid<<<(unsigned)blocks, 128, 0, st>>>(a);
    PCREG_LAUNCHED();
}

}  // namespace pcreg
