// nn_grid.cu -- exact nearest neighbour on the uniform grid + occupancy pyramid (large models).
//
// Semantics identical to nn_brute.cu (knnsearch K=1, FP64, ties -> smallest original index); the
// reference's nearest analogue of a spatially pruned search is speedyDescriptors.m:44-60 (boxes with a
// halo) and getLocalPoints.m:8-15 (cube pre-filter).  One thread per query walks the pyramid with a
// small explicit stack (branch and bound): children are visited nearest-octant first, a cell is
// skipped only when a CONSERVATIVE lower bound of its distance exceeds the best exact distance so far,
// and every visited point is evaluated in FP64 with the oracle's formula -- so the answer equals the
// FP64 brute-force answer bit for bit.  The previous ICP iteration's correspondence seeds the bound.
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

// stack entry: [63:34] lower bound (positive float, lowest mantissa bit dropped = rounded down),
//              [33:30] level, [29:20] z, [19:10] y, [9:0] x
__device__ __forceinline__ unsigned long long pack_entry(float lb, int level, int x, int y, int z) {
    return ((unsigned long long)(__float_as_uint(lb) >> 1) << 34) | ((unsigned long long)(unsigned)level << 30) |
           ((unsigned long long)(unsigned)z << 20) | ((unsigned long long)(unsigned)y << 10) | (unsigned long long)(unsigned)x;
}

__global__ void __launch_bounds__(128) k_nn_grid(const __grid_constant__ GridArgs a) {
    const int64_t gq = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_pts = 0, n_cells = 0, n_nodes = 0;
    if (gq < a.nq) {
        const int64_t h = gq / a.ns, i = gq - h * a.ns;
        double qx, qy, qz;
        quick_tf(a.T + h * 16, a.sx[i], a.sy[i], a.sz[i], qx, qy, qz);
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
        const GridView& G = a.g;
        unsigned long long stack[GRID_STACK];
        int sp = 0;
        stack[sp++] = pack_entry(0.f, G.nlevels - 1, 0, 0, 0);
        const double slop = 1e-9 * G.cell;
        while (sp > 0) {
            const unsigned long long e = stack[--sp];
            const float lbf = __uint_as_float((unsigned)(e >> 34) << 1);
            if ((double)lbf > best) continue;
            const unsigned lo32 = (unsigned)e;
            const int level = (int)((e >> 30) & 0xF) ;
            const int ix = (int)(lo32 & 1023u), iy = (int)((lo32 >> 10) & 1023u), iz = (int)((lo32 >> 20) & 1023u);
            ++n_nodes;
            if (level == 0) {
                const int64_t c = ((int64_t)iz * G.dims[0][1] + iy) * G.dims[0][0] + ix;
                const int32_t s0 = G.cell_start[c], s1 = G.cell_start[c + 1];
                ++n_cells;
                n_pts += (unsigned long long)(s1 - s0);
                for (int32_t p = s0; p < s1; ++p) {
                    const GridPoint gp = G.pts[p];
                    const double d = dist2_exact(gp.x, gp.y, gp.z, qx, qy, qz);
                    if (d < best || (d == best && gp.orig < bidx)) { best = d; bidx = gp.orig; }
                }
                continue;
            }
            const int64_t c = ((int64_t)iz * G.dims[level][1] + iy) * G.dims[level][0] + ix;
            const unsigned mask = G.mask[level][c];
            const int cl = level - 1;
            const double edge = ldexp(G.cell, cl);                 // child cell edge
            // octant of q relative to the node centre (centre = lo + edge)
            const double cx0 = G.origin[0] + (double)(2 * ix + 1) * edge;
            const double cy0 = G.origin[1] + (double)(2 * iy + 1) * edge;
            const double cz0 = G.origin[2] + (double)(2 * iz + 1) * edge;
            const int oct = (qx >= cx0 ? 1 : 0) | (qy >= cy0 ? 2 : 0) | (qz >= cz0 ? 4 : 0);
#pragma unroll 1
            for (int t = 7; t >= 0; --t) {
                const int k = t ^ oct;
                if (!((mask >> k) & 1u)) continue;
                const int x = 2 * ix + (k & 1), y = 2 * iy + ((k >> 1) & 1), z = 2 * iz + (k >> 2);
                const double lx = G.origin[0] + (double)x * edge - slop, hx = lx + edge + 2.0 * slop;
                const double ly = G.origin[1] + (double)y * edge - slop, hy = ly + edge + 2.0 * slop;
                const double lz = G.origin[2] + (double)z * edge - slop, hz = lz + edge + 2.0 * slop;
                const double ddx = fmax(fmax(lx - qx, qx - hx), 0.0);
                const double ddy = fmax(fmax(ly - qy, qy - hy), 0.0);
                const double ddz = fmax(fmax(lz - qz, qz - hz), 0.0);
                const double lb = (ddx * ddx + ddy * ddy + ddz * ddz) * (1.0 - 1e-12);
                if (lb > best) continue;
                const float lbd = __double2float_rd(lb);
                if (sp < GRID_STACK) stack[sp++] = pack_entry(lbd, cl, x, y, z);
            }
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
