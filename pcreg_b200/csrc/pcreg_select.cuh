// pcreg_select.cuh -- fast block-wide "K smallest" selection for the ICP trim (AlignPoints_KNN.m:20-26
// rule: keep the round(k*n) smallest residuals, MATLAB's stable sort breaks ties by lower index).
//
// Residuals are continuous, so instead of eight 8-bit radix passes (whose leading passes put every key
// into one or two bins and serialise on shared-memory atomics) the K-th value is located with ONE linear
// histogram over [min, max] (monotone binning in FP64), the few keys of the boundary bin are collected,
// and the K-th is ranked among them exactly on the full 64-bit key.  Falls back to the radix select
// (pcreg_dev.cuh) when the boundary bin is crowded (heavy duplicates).  Exact ties at the boundary are
// resolved in ORIGINAL index order, as in block_radix_select.
#pragma once
#include "pcreg_dev.cuh"

namespace pcreg {

constexpr int SEL_BINS = 2048;
constexpr int SEL_CAP = 1024;
#ifndef SEL_UNROLL
#define SEL_UNROLL 4
#endif

struct HistSelShared {
    int hist[SEL_BINS];
    unsigned long long cand[SEL_CAP];
    int ncand;
    int bin;                    // boundary bin
    long long before;           // keys in lower bins
    unsigned long long vK;
    long long need_eq;
    int count_eq;
    unsigned long long kmin, kmax;
    RadixSelShared radix;       // fallback + tie resolution scratch
};

// keys[0..n): bit patterns of non-negative doubles (KEY_NOSEL = excluded); kmin/kmax: min / max over the
// included keys (block-uniform); K >= 1 and K <= number of included keys.
__device__ __forceinline__ void block_hist_select(unsigned long long* __restrict__ keys, long long n, long long K,
                                                  unsigned long long kmin, unsigned long long kmax, HistSelShared& sh,
                                                  unsigned long long& vK, bool& all_eq, const int32_t* __restrict__ order) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (K <= 0) { vK = 0ull; all_eq = false; return; }
    const double rmin = __longlong_as_double((long long)kmin), rmax = __longlong_as_double((long long)kmax);
    bool use_radix = !(rmax > rmin) || !(rmax - rmin < 1.0e300);
    long long need_eq = 0;
    int count_eq = 0;
    if (!use_radix) {
        const double scale = (double)SEL_BINS / (rmax - rmin);
        for (int b = tid; b < SEL_BINS; b += nthr) sh.hist[b] = 0;
        if (tid == 0) sh.ncand = 0;
        __syncthreads();
        // SEL_UNROLL keys per trip, loaded before any is used: with keys in global memory (k_icp_update) a pass over them is
        // bound by the load latency of one key per thread otherwise
        for (long long i0 = tid; i0 < n; i0 += (long long)SEL_UNROLL * nthr) {
            unsigned long long kk[SEL_UNROLL];
#pragma unroll
            for (int u = 0; u < SEL_UNROLL; ++u) { const long long i = i0 + (long long)u * nthr; kk[u] = i < n ? keys[i] : KEY_NOSEL; }
#pragma unroll
            for (int u = 0; u < SEL_UNROLL; ++u) {
                const unsigned long long key = kk[u];
                if (key == KEY_NOSEL) continue;
                int b = (int)((__longlong_as_double((long long)key) - rmin) * scale);
                b = b < 0 ? 0 : (b >= SEL_BINS ? SEL_BINS - 1 : b);
                atomicAdd(&sh.hist[b], 1);
            }
        }
        __syncthreads();
        if (tid < 32) {                                   // warp 0: find the bin where the cumulative count reaches K
            constexpr int PER = SEL_BINS / 32;
            int s = 0;
            for (int k = 0; k < PER; ++k) s += sh.hist[tid * PER + k];
            int inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= o) inc += t;
            }
            const long long excl = (long long)inc - s;
            if (excl < K && K <= (long long)inc) {
                long long cum = excl;
                int b = tid * PER;
                for (;; ++b) {
                    if (cum + sh.hist[b] >= K) break;
                    cum += sh.hist[b];
                }
                sh.bin = b;
                sh.before = cum;
            }
        }
        __syncthreads();
        const int bstar = sh.bin;
        for (long long i0 = tid; i0 < n; i0 += (long long)SEL_UNROLL * nthr) {
            unsigned long long kk[SEL_UNROLL];
#pragma unroll
            for (int u = 0; u < SEL_UNROLL; ++u) { const long long i = i0 + (long long)u * nthr; kk[u] = i < n ? keys[i] : KEY_NOSEL; }
#pragma unroll
            for (int u = 0; u < SEL_UNROLL; ++u) {
                const unsigned long long key = kk[u];
                if (key == KEY_NOSEL) continue;
                int b = (int)((__longlong_as_double((long long)key) - rmin) * scale);
                b = b < 0 ? 0 : (b >= SEL_BINS ? SEL_BINS - 1 : b);
                if (b == bstar) {
                    const int pos = atomicAdd(&sh.ncand, 1);
                    if (pos < SEL_CAP) sh.cand[pos] = key;
                }
            }
        }
        __syncthreads();
        const int nc = sh.ncand;
        if (nc > SEL_CAP) {
            use_radix = true;                             // crowded boundary bin: exact radix select instead
        } else {
            const long long kk = K - sh.before;           // rank of the K-th key inside the boundary bin (1-based)
            for (int t = tid; t < nc; t += nthr) {
                const unsigned long long mine = sh.cand[t];
                int less = 0, eq = 0;
                for (int j = 0; j < nc; ++j) {
                    const unsigned long long o = sh.cand[j];
                    less += o < mine ? 1 : 0;
                    eq += o == mine ? 1 : 0;
                }
                if ((long long)less < kk && kk <= (long long)(less + eq)) {       // every holder of that key writes the same values
                    sh.vK = mine;
                    sh.need_eq = kk - less;
                    sh.count_eq = eq;
                }
            }
            __syncthreads();
            vK = sh.vK;
            need_eq = sh.need_eq;
            count_eq = sh.count_eq;
        }
    }
    if (use_radix) {                                      // block-uniform decision
        block_radix_select(keys, n, K, sh.radix, vK, all_eq, order);
        return;
    }
    all_eq = (need_eq == (long long)count_eq);
    if (!all_eq) {
        // stable tie rule: rank the keys equal to vK in ORIGINAL index order, keep the first need_eq
        RadixSelShared& rs = sh.radix;
        if (tid == 0) rs.run_eq = 0;
        __syncthreads();
        const int lane = tid & 31, warp = tid >> 5, nwarp = (nthr + 31) >> 5;
        for (long long base = 0; base < n; base += nthr) {
            const long long r = base + tid;
            const long long i = (r < n) ? (order ? (long long)order[r] : r) : n;
            const bool eq = i < n && keys[i] == vK;
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) rs.warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int off = rs.run_eq;
            for (int w = 0; w < warp; ++w) off += rs.warp_cnt[w];
            const int rank = off + __popc(bal & ((1u << lane) - 1u));
            if (eq) keys[i] = ((long long)rank < need_eq) ? KEY_SEL : KEY_NOSEL;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < nwarp; ++w) t += rs.warp_cnt[w]; rs.run_eq += t; }
            __syncthreads();
        }
    }
    __syncthreads();
}

}  // namespace pcreg
