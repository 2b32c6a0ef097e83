// pcreg_grid.cuh -- kernel argument block and warp helpers shared by the grid NN kernels (nn_grid.cu) and the
// voxel-map list scan (nn_vox.cu).
#pragma once
#include "pcreg_internal.h"

namespace pcreg {

struct GridArgs {
    GridView g;
    const ModelPointD* md;
    const double* sx; const double* sy; const double* sz; int64_t ns;
    const double* T; int64_t nq;
    const int32_t* prev;
    int32_t* idx; double* d2;
    VoxView vox;                    // Voronoi voxel map of the model (nn_vox.cu; vox.hdr == nullptr: none)
    CandView cl;                    // candidate lists of the chunk (cl.cnt == nullptr: disabled)
    const double* skip_thr;         // [nhyp] or null: list kernel skips queries whose previous residual exceeds it (lazy trimming)
    const int32_t* in_list;         // direct kernel: the queries to process (nullptr: all nq)
    const unsigned int* in_count;   //                and how many
    int32_t* worklist;              // [nq] query ids handed to the next kernel (list -> direct -> walk)
    unsigned int* work_count;       // number of entries in worklist
    unsigned long long* cursor;     // direct kernel: next unassigned position of its input (zeroed before the launch)
    int32_t* overflow;              // warp walk: queries it hands on to the per-lane walk (frontier too large, no bound)
    unsigned int* overflow_count;
    int ww_cap;                     // warp walk: frontier entries it may use (<= WW_CAP; the tests force overflows with a small one)
    int fetch_batch, chunk;         // (tuning)
    int chain;                      // first-pass walk: consecutive queries per thread
    int row_span;                   // direct kernel: widest (y,z) cell span it row-scans itself
    unsigned long long* counters;   // profiling only (may be null): [0] points / [1] rows visited by the row scan, [2] pyramid
                                    // nodes popped, [3] queries answered from their list, [4] queries walked, [5] queries
                                    // row-scanned, [6] list entries read, [7] list points gathered, [8] points / [9] leaf cells
                                    // visited by the walk
};

__device__ __forceinline__ void flush_counters(unsigned long long* counters, unsigned long long n_pts,
                                               unsigned long long n_cells, unsigned long long n_nodes, int i_pts = 0, int i_cells = 1,
                                               int i_nodes = 2) {
    if (!counters) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_pts += __shfl_xor_sync(0xffffffffu, n_pts, o);
        n_cells += __shfl_xor_sync(0xffffffffu, n_cells, o);
        n_nodes += __shfl_xor_sync(0xffffffffu, n_nodes, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_pts) atomicAdd(&counters[i_pts], n_pts);
        if (n_cells) atomicAdd(&counters[i_cells], n_cells);
        if (n_nodes) atomicAdd(&counters[i_nodes], n_nodes);
    }
}

// warp-aggregated append of query ids to the next kernel's work list
__device__ __forceinline__ void worklist_append(const GridArgs& a, bool defer, int64_t gq, int lane) {
    const unsigned dm = __ballot_sync(0xffffffffu, defer);
    if (dm) {
        const int leader = __ffs(dm) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(a.work_count, (unsigned)__popc(dm));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (defer) a.worklist[base + __popc(dm & ((1u << lane) - 1u))] = (int32_t)gq;
    }
}

}  // namespace pcreg
