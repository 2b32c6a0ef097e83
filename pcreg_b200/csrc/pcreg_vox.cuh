// pcreg_vox.cuh -- device side of the Voronoi voxel map (built by nn_vox.cu): voxel addressing and the exact list scan
// of one query.  Shared by k_nn_vox (one launch per NN pass) and the fused ICP kernel (icp_fused.cu).
#pragma once
#include <float.h>
#include "pcreg_internal.h"
#include "pcreg_dev.cuh"

namespace pcreg {

// hdr arrays are stored in 4 x 4 x 4 bricks: queries of a warp are neighbours in space, so their headers share sectors
__host__ __device__ __forceinline__ int64_t brick_index(int x, int y, int z, const int32_t* tiles) {
    const int64_t tile = ((int64_t)(z >> 2) * tiles[1] + (y >> 2)) * tiles[0] + (x >> 2);
    return (tile << 6) | (int64_t)(((z & 3) << 4) | ((y & 3) << 2) | (x & 3));
}
__device__ __forceinline__ void brick_decode(int64_t slot, const int32_t* tiles, int& x, int& y, int& z) {
    const int64_t tile = slot >> 6;
    const int in = (int)(slot & 63);
    const int tx = (int)(tile % tiles[0]), ty = (int)((tile / tiles[0]) % tiles[1]), tz = (int)(tile / ((int64_t)tiles[0] * tiles[1]));
    x = tx * 4 + (in & 3); y = ty * 4 + ((in >> 2) & 3); z = tz * 4 + (in >> 4);
}
// the one formula for a voxel centre (build and query must agree to FP64 rounding)
__device__ __forceinline__ double vox_centre(double origin, int i, double s) { return __fma_rn((double)i + 0.5, s, origin); }

// Voxel of q: its list header and q relative to the voxel centre (FP32).  False when the map cannot answer (q outside the
// padded box -- a NaN pose fails the comparisons too -- or a voxel whose list was dropped): the caller walks the pyramid.
__device__ __forceinline__ bool vox_lookup(const VoxView& V, double qx, double qy, double qz, uint2& hd, float& x, float& y, float& z) {
    const double ux = (qx - V.origin[0]) * V.inv_s, uy = (qy - V.origin[1]) * V.inv_s, uz = (qz - V.origin[2]) * V.inv_s;
    if (!(ux >= 0.0 && uy >= 0.0 && uz >= 0.0 && ux < (double)V.dims[0] && uy < (double)V.dims[1] && uz < (double)V.dims[2])) return false;
    const int ix = (int)ux, iy = (int)uy, iz = (int)uz;
    hd = V.hdr[brick_index(ix, iy, iz, V.tiles)];
    x = __double2float_rn(qx - vox_centre(V.origin[0], ix, V.s));
    y = __double2float_rn(qy - vox_centre(V.origin[1], iy, V.s));
    z = __double2float_rn(qz - vox_centre(V.origin[2], iz, V.s));
    return hd.y != 0u;
}

struct __align__(32) EntPair { float4 a, b; };

// Exact nearest neighbour of q among the entries of its voxel's list: FP32 scan (BATCH entries in flight), then the entries
// inside the FP32 error band -- normally one -- decided in FP64 with the oracle's formula on (d2, original index).
// BATCH = list entries a thread has in flight per trip: 4 in k_nn_vox (more registers cost it occupancy: C5 25.0 vs 25.9 ms),
// 8 in the fused kernel (128 registers anyway: C3 37.3 vs 36.3 ms; 2: 41.4).
// PAIRED: lists start on even entries and odd ones end with a far sentinel (nn_vox.cu), so two entries can be read with one
// 256-bit load (LDG.E.256 on sm_100a): half the load instructions -- and L1 wavefronts -- of the scan.  The fused kernel uses
// it (C3 37.3 -> 35.8 ms per step, C4 481 -> 381 ms per 16 384 poses); k_nn_vox does not: six more registers cost it a block
// per SM (C5 322 -> 340 ms per step; at the same occupancy with spills 354; C3 per-pass path 18.8 -> 18.0 ms, not worth it).
// Measured and dropped: reading the last pair again past the end of a list and selecting a far distance instead of the
// predicated load -- fewer instructions but more loads: C3 35.8 -> 36.6 ms, C4 381 -> 399 ms.  The scan is bound by load
// wavefronts, not by issue slots.
template <int BATCH, bool PAIRED>
__device__ __forceinline__ void vox_scan(const VoxView& V, const GridPoint* __restrict__ pts, uint2 hd, float x, float y, float z,
                                         double qx, double qy, double qz, int32_t& bidx, double& best, unsigned& n_gather,
                                         double* win = nullptr /* optional: the winner's coordinates [3] */) {
    const float4* __restrict__ L = V.ent + hd.x;
    const uint32_t n = hd.y;
    float m1 = FLT_MAX, m2 = FLT_MAX;
    int r1 = 0;
    auto take = [&](const float4 en) {
        const float dx = en.x - x, dy = en.y - y, dz = en.z - z;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d < m1) { m2 = m1; m1 = d; r1 = __float_as_int(en.w); } else m2 = fminf(m2, d);
    };
    if (PAIRED) {
        static_assert(BATCH % 2 == 0, "vox_scan: BATCH counts entries, loaded in pairs");
        const EntPair* __restrict__ L2 = reinterpret_cast<const EntPair*>(L);
        const uint32_t np = (n + 1u) >> 1;
        EntPair far;
        far.a = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f); far.b = far.a;
        for (uint32_t k = 0; k < np; k += BATCH / 2) {
            EntPair e[BATCH / 2];
#pragma unroll
            for (int u = 0; u < BATCH / 2; ++u) e[u] = (k + u < np) ? L2[k + u] : far;      // BATCH / 2 loads in flight
#pragma unroll
            for (int u = 0; u < BATCH / 2; ++u) { take(e[u].a); take(e[u].b); }
        }
    } else {
        const float4 far = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f);
        for (uint32_t k = 0; k < n; k += BATCH) {
            float4 e[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) e[u] = (k + u < n) ? L[k + u] : far;      // BATCH loads in flight
#pragma unroll
            for (int u = 0; u < BATCH; ++u) take(e[u]);
        }
    }
    const float thr = fmaf(m1, 3e-6f, m1) + V.band_abs;
    {
        const GridPoint gp = pts[r1];
        best = dist2_exact(gp.x, gp.y, gp.z, qx, qy, qz);
        bidx = gp.orig;
        if (win) { win[0] = gp.x; win[1] = gp.y; win[2] = gp.z; }
    }
    n_gather = 1;
    if (m2 <= thr) {                                     // more than one entry inside the FP32 error band: decide in FP64
        for (uint32_t k = 0; k < n; ++k) {
            const float4 en = L[k];
            const float dx = en.x - x, dy = en.y - y, dz = en.z - z;
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const int r = __float_as_int(en.w);
            if (d <= thr && r != r1) {
                const GridPoint gp = pts[r];
                const double dd = dist2_exact(gp.x, gp.y, gp.z, qx, qy, qz);
                if (dd < best || (dd == best && gp.orig < bidx)) {
                    best = dd; bidx = gp.orig;
                    if (win) { win[0] = gp.x; win[1] = gp.y; win[2] = gp.z; }
                }
                ++n_gather;
            }
        }
    }
}

}  // namespace pcreg
