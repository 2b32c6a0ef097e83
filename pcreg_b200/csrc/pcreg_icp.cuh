// pcreg_icp.cuh -- the per-correspondence arithmetic of one ICP pass, shared by k_icp_update (icp.cu, one launch per pass)
// and the fused per-hypothesis kernel (icp_fused.cu) so that both produce the same bits:
//   rejection   squared-distance compare (ransac.m:49: dist < thDist keeps)
//   weights     w = max(R_w - r, 0) (AlignPoints_weighted.m:16-18), times the caller's w_src
//   17 sums     estimateTransform.m:41-58 as one pass (sum w, sum w q, sum w m, sum w q m', sum w d^2), relative to a pivot
//   pose update estimateTransform.m:60-71 (kabsch_from_sums) and T <- T * dT
#pragma once
#include "pcreg_internal.h"
#include "pcreg_dev.cuh"
#include "pcreg_math.cuh"

namespace pcreg {

// ICP passes sum over the correspondences of one hypothesis with this many threads (thread t takes i = t, t + N, ...);
// both kernels use the same count so that their fixed-tree reductions agree bit for bit
#ifdef UPD_THREADS_OVERRIDE
constexpr int UPD_THREADS = UPD_THREADS_OVERRIDE;
#else
constexpr int UPD_THREADS = 256;       // fused kernel, C3 step, same-session A/B (threads x blocks per SM, registers): 256 x 2 (128) 38.9 ms;
#endif                                 // 224 x 2: 40.0; 320 x 2 (96): 40.4; 192 x 2 (168): 42.0; 384 x 2 (80, spills): 44.4; 512 x 1 (128): 45.0;
                                       // 160 x 2: 47.7; 512 x 2 (64): 48-53; 256 x 3 (80): 51.5; 128 x 2 (212, no spills): 53.7
#ifndef FUSED_MIN_BLOCKS
#define FUSED_MIN_BLOCKS 2
#endif

__device__ __forceinline__ void icp_accumulate(double (&s)[KABSCH_NSUMS], double w, double d, double qx, double qy, double qz,
                                               const ModelPointD& m, double px, double py, double pz) {
    const double q0 = qx - px, q1 = qy - py, q2 = qz - pz;
    const double m0 = m.x - px, m1 = m.y - py, m2 = m.z - pz;
    const double wq0 = w * q0, wq1 = w * q1, wq2 = w * q2;
    s[0] += w;
    s[1] += wq0; s[2] += wq1; s[3] += wq2;
    s[4] += w * m0; s[5] += w * m1; s[6] += w * m2;
    s[7] += wq0 * m0; s[8] += wq0 * m1; s[9] += wq0 * m2;
    s[10] += wq1 * m0; s[11] += wq1 * m1; s[12] += wq1 * m2;
    s[13] += wq2 * m0; s[14] += wq2 * m1; s[15] += wq2 * m2;
    s[16] += w * d;
}

// Thread 0 of a hypothesis after the block reduction: rigid fit from the sums and T <- T * dT.  Returns false (pose left
// alone, hypothesis to be frozen) with fewer than 3 usable correspondences.  Tc: current pose (row-major), Tn: new pose.
__device__ __forceinline__ bool icp_pose_update(const double (&s)[KABSCH_NSUMS], long long n_used, const double* pivot, bool reflection_fix,
                                                const double* Tc, double* Tn) {
    const double sw = s[0];
    if (n_used < 3 || !(sw > 0.0)) return false;
    KabschSums ks;
    ks.sw = sw;
    for (int k = 0; k < 3; ++k) { ks.sq[k] = s[1 + k]; ks.sm[k] = s[4 + k]; }
    for (int k = 0; k < 9; ++k) ks.sqm[k] = s[7 + k];
    ks.swd2 = s[16];
    double dT[16];
    kabsch_from_sums(ks, pivot, pivot, reflection_fix, dT);
    mul4(Tc, dT, Tn);
    return true;
}

// out-of-line copy for the kernels' hot loops' sake: one thread per hypothesis runs it, its SVD must not claim their registers
__device__ __noinline__ static bool icp_pose_update_call(const double (&s)[KABSCH_NSUMS], long long n_used, const double* pivot,
                                                         bool reflection_fix, const double* Tc, double* Tn) {
    return icp_pose_update(s, n_used, pivot, reflection_fix, Tc, Tn);
}

}  // namespace pcreg
