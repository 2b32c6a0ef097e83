// model.cu -- library context, error state, the GPU-resident model cloud handle and the
// on-device uniform-grid build (hand-written counting sort + occupancy pyramid).
//
// Reference context: every driver loads one dense model cloud and reuses it for thousands of
// alignments (completeExperiment.m:15, slideMatchingWindow_v2.m:15, class single as written by
// upsampleMesh.m:21); the handle makes that cloud resident in HBM once.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <atomic>
#include <mutex>
#include <thread>
#include <functional>
#include <vector>
#include <algorithm>

#include "pcreg_internal.h"
#include "pcreg_dev.cuh"

namespace pcreg {

static thread_local char g_err[1024] = "";
static char g_err_global[1024] = "";
static std::atomic<long long> g_launches{0};
static Context g_ctxs[PCREG_MAX_DEVICES];
static int g_nslots = 0;
static thread_local int tl_slot = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    memcpy(g_err_global, g_err, sizeof g_err);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- caching device allocator (one pool per device; blocks are found by pointer, so a buffer may be released from
// any host thread) ----
static std::mutex g_pool_mu;
static constexpr size_t POOL_KEEP_BYTES = (size_t)24 << 30;
static void pool_trim_locked(Context& c, size_t keep_bytes) {
    size_t free_bytes = 0;
    for (auto& b : c.pool) if (!b.used) free_bytes += b.bytes;
    if (free_bytes <= keep_bytes) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c.device);
    for (size_t i = 0; i < c.pool.size() && free_bytes > keep_bytes;) {
        if (!c.pool[i].used) {
            cudaFree(c.pool[i].p);
            free_bytes -= c.pool[i].bytes;
            c.pool[i] = c.pool.back();
            c.pool.pop_back();
        } else ++i;
    }
    if (prev >= 0) cudaSetDevice(prev);
}
void* pool_alloc(size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    Context& c = ctx();
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (int i = 0; i < (int)c.pool.size(); ++i) {
            const PoolBlock& b = c.pool[i];
            if (!b.used && b.bytes >= bytes && b.bytes <= 2 * bytes + 4096 && (best < 0 || b.bytes < c.pool[best].bytes)) best = i;
        }
        if (best >= 0) { c.pool[best].used = true; return c.pool[best].p; }
    }
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_trim(0);                                   // give cached blocks back and retry once
        e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) throw CudaFail{e, "cudaMalloc", __FILE__, __LINE__};
    }
    std::lock_guard<std::mutex> lk(g_pool_mu);
    c.pool.push_back(PoolBlock{q, bytes, true});
    return q;
}
void pool_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (int k = 0; k < std::max(g_nslots, 1); ++k) {
        Context& c = g_ctxs[k];
        bool mine = false;
        size_t free_bytes = 0;
        for (auto& b : c.pool) {
            if (b.p == p) { b.used = false; mine = true; }
            if (!b.used) free_bytes += b.bytes;
        }
        if (mine) {
            if (free_bytes > POOL_KEEP_BYTES) pool_trim_locked(c, POOL_KEEP_BYTES / 2);
            return;
        }
    }
}
void pool_trim(size_t keep_bytes) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    pool_trim_locked(ctx(), keep_bytes);
}
cudaEvent_t pooled_event(size_t i) {
    Context& c = ctx();
    while (c.events.size() <= i) {
        cudaEvent_t e;
        PCREG_CUDA(cudaEventCreate(&e));
        c.events.push_back(e);
    }
    return c.events[i];
}
cudaStream_t lane_stream(int i) {
    Context& c = ctx();
    while ((int)c.streams.size() <= i) {
        cudaStream_t s;
        PCREG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        c.streams.push_back(s);
    }
    return c.streams[i];
}
Context& ctx() { return g_ctxs[tl_slot]; }
int num_slots() { return g_nslots; }
int current_slot() { return tl_slot; }
void use_slot(int slot) {
    if (slot < 0 || slot >= std::max(g_nslots, 1)) throw ArgError{"internal: bad device slot"};
    tl_slot = slot;
    if (g_ctxs[slot].initialised) PCREG_CUDA(cudaSetDevice(g_ctxs[slot].device));
}

constexpr size_t PINNED_BYTES = (size_t)16 << 20;
void h2d_columns(double* d_dst, int64_t ld_dst, const double* h_src, int64_t ld_src, int64_t rows, int64_t ncols, cudaStream_t st) {
    if (rows <= 0 || ncols <= 0) return;
    Context& c = ctx();
    for (int b = 0; b < 2; ++b) {
        if (!c.pinned[b]) {
            PCREG_CUDA(cudaHostAlloc(&c.pinned[b], PINNED_BYTES, cudaHostAllocDefault));
            PCREG_CUDA(cudaEventCreateWithFlags(&c.pinned_ev[b], cudaEventDisableTiming));
        }
    }
    const size_t col_bytes = (size_t)rows * sizeof(double);
    if (col_bytes > PINNED_BYTES) {           // a single column larger than a bounce buffer: let the runtime stage it
        PCREG_CUDA(cudaMemcpy2DAsync(d_dst, (size_t)ld_dst * 8, h_src, (size_t)ld_src * 8, col_bytes, (size_t)ncols, cudaMemcpyHostToDevice, st));
        return;
    }
    const int64_t per = (int64_t)(PINNED_BYTES / col_bytes);
    int b = 0;
    bool used[2] = {false, false};
    for (int64_t c0 = 0; c0 < ncols; c0 += per, b ^= 1) {
        const int64_t nc = std::min<int64_t>(per, ncols - c0);
        if (used[b]) PCREG_CUDA(cudaEventSynchronize(c.pinned_ev[b]));          // the DMA that read this buffer has finished
        char* stage = (char*)c.pinned[b];
        if (ld_src == rows) memcpy(stage, h_src + c0 * ld_src, (size_t)nc * col_bytes);
        else for (int64_t k = 0; k < nc; ++k) memcpy(stage + (size_t)k * col_bytes, h_src + (c0 + k) * ld_src, col_bytes);
        PCREG_CUDA(cudaMemcpy2DAsync(d_dst + c0 * ld_dst, (size_t)ld_dst * 8, stage, col_bytes, col_bytes, (size_t)nc, cudaMemcpyHostToDevice, st));
        PCREG_CUDA(cudaEventRecord(c.pinned_ev[b], st));
        used[b] = true;
    }
    // the caller may reuse / free h_src right away (it was copied), but the bounce buffers are still being read:
    // the next h2d_columns call waits on its own events only if it reuses them, so drain here to keep it simple
    for (int k = 0; k < 2; ++k) if (used[k]) PCREG_CUDA(cudaEventSynchronize(c.pinned_ev[k]));
}
void require_init() {
    if (!ctx().initialised) throw ArgError{"pcreg_init has not been called (or failed): no CUDA device, and there is no CPU fallback"};
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// scan-order float4 array: (x,y,z) relative to the pivot rounded to FP32, w = |.|^2 of the ROUNDED
// coordinates (computed in FP64, rounded once).  Padding entries are far-away sentinels.
__global__ void k_build_m4(const ModelPointD* __restrict__ md, const int32_t* __restrict__ perm, int64_t n,
                           int64_t n_pad, double px, double py, double pz, float4* __restrict__ m4) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    float4 o;
    if (j < n) {
        const ModelPointD p = md[perm[j]];
        o.x = __double2float_rn(p.x - px);
        o.y = __double2float_rn(p.y - py);
        o.z = __double2float_rn(p.z - pz);
        o.w = __double2float_rn((double)o.x * (double)o.x + (double)o.y * (double)o.y + (double)o.z * (double)o.z);
    } else {
        o.x = 1.0e18f; o.y = 1.0e18f; o.z = 1.0e18f; o.w = 3.0e36f;
    }
    m4[j] = o;
}

__device__ __forceinline__ int cell_coord(double v, double origin, double inv_cell, int dim) {
    int c = (int)floor((v - origin) * inv_cell);
    return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

__global__ void k_grid_count(const ModelPointD* __restrict__ md, int64_t n, GridView g, int32_t* __restrict__ cell_of,
                             int32_t* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ModelPointD p = md[i];
    const int cx = cell_coord(p.x, g.origin[0], g.inv_cell, g.dims[0][0]);
    const int cy = cell_coord(p.y, g.origin[1], g.inv_cell, g.dims[0][1]);
    const int cz = cell_coord(p.z, g.origin[2], g.inv_cell, g.dims[0][2]);
    const int32_t c = (cz * g.dims[0][1] + cy) * g.dims[0][0] + cx;
    cell_of[i] = c;
    atomicAdd(&count[c], 1);
}

// --- exclusive scan (three phases, hand-written) ---
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_excl_scan(int v, int* smem /*[32]*/, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = smem[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;               // exclusive warp offsets
        if (lane == 31) smem[32] = winc;     // block total
    }
    __syncthreads();
    const int res = inc - v + smem[warp];
    total = smem[32];
    __syncthreads();
    return res;
}

// phase 1: in-place exclusive scan of each SCAN_BLOCK chunk, chunk totals to block_sums
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_local(int32_t* __restrict__ data, int64_t n, int32_t* __restrict__ block_sums) {
    __shared__ int smem[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? data[base + k] : 0; s += v[k]; }
    int total;
    int off = block_excl_scan(s, smem, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) data[base + k] = off; off += v[k]; }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
// phase 2: one block scans the chunk totals (in place, exclusive); grand total to *total_out
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(int32_t* __restrict__ sums, int nb, int32_t* __restrict__ total_out) {
    __shared__ int smem[33];
    int carry = 0;
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const int v = (i < nb) ? sums[i] : 0;
        int total;
        const int off = block_excl_scan(v, smem, total);
        if (i < nb) sums[i] = off + carry;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}
// phase 3: add chunk offsets; also writes the terminating entry data[n] = grand total
__global__ void k_scan_add(int32_t* __restrict__ data, int64_t n, const int32_t* __restrict__ sums, const int32_t* __restrict__ total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += sums[i / SCAN_BLOCK];
    if (i == n) data[n] = *total;
}

__global__ void k_grid_scatter(const ModelPointD* __restrict__ md, int64_t n, const int32_t* __restrict__ cell_of,
                               const int32_t* __restrict__ cell_start, int32_t* __restrict__ cursor,
                               GridPoint* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = cell_of[i];
    const int32_t pos = cell_start[c] + atomicAdd(&cursor[c], 1);
    const ModelPointD p = md[i];
    GridPoint gp;
    gp.x = p.x; gp.y = p.y; gp.z = p.z; gp.orig = (int32_t)i; gp.pad = 0;
    out[pos] = gp;
}

// occupancy pyramid: level-l mask bit (dz<<2 | dy<<1 | dx) set iff that child at level l-1 is non-empty
__global__ void k_grid_mask(GridView g, int level, const int32_t* __restrict__ cell_start, const uint8_t* __restrict__ below,
                            uint8_t* __restrict__ out, unsigned long long* __restrict__ occupied0) {
    const int dx = g.dims[level][0], dy = g.dims[level][1], dz = g.dims[level][2];
    const int64_t total = (int64_t)dx * dy * dz;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total) return;
    const int ix = (int)(c % dx), iy = (int)((c / dx) % dy), iz = (int)(c / ((int64_t)dx * dy));
    const int bx = g.dims[level - 1][0], by = g.dims[level - 1][1], bz = g.dims[level - 1][2];
    unsigned m = 0;
    int occ = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int x = ix * 2 + (k & 1), y = iy * 2 + ((k >> 1) & 1), z = iz * 2 + (k >> 2);
        if (x >= bx || y >= by || z >= bz) continue;
        const int64_t cc = ((int64_t)z * by + y) * bx + x;
        bool ne;
        if (level == 1) ne = cell_start[cc + 1] > cell_start[cc];
        else            ne = below[cc] != 0;
        if (ne) { m |= 1u << k; ++occ; }
    }
    out[c] = (uint8_t)m;
    if (level == 1 && occ) atomicAdd(occupied0, (unsigned long long)occ);
}

// ------------------------------------------------------------------------------------------------
// grid build (host driver)
// ------------------------------------------------------------------------------------------------
void grid_build(pcreg_model* m, const pcreg_model_opts& o, cudaStream_t st) {
    const int64_t n = m->n;
    GridView g{};
    double ext[3], maxext = 0.0;
    for (int a = 0; a < 3; ++a) { ext[a] = m->bbox_hi[a] - m->bbox_lo[a]; maxext = std::max(maxext, ext[a]); }
    if (!(maxext > 0.0)) maxext = 1.0;
    double cell = o.cell_size;
    if (!(cell > 0.0)) {
        static const double cpp_env = [] { const char* e = getenv("PCREG_CPP"); return e ? atof(e) : 0.0; }();
        const double cpp = o.cells_per_point > 0.0 ? o.cells_per_point : (cpp_env > 0.0 ? cpp_env : 32.0);
        const double cap = (double)(o.max_cells > 0 ? o.max_cells : ((int64_t)1 << 27));
        const double target = std::min(cap, std::max(64.0, cpp * (double)n));
        double vol = 1.0;
        for (int a = 0; a < 3; ++a) vol *= std::max(ext[a], 1e-3 * maxext);
        cell = cbrt(vol / target);
    }
    cell = std::max(cell, maxext / 1000.0);           // at most ~1001 cells per axis (10-bit coordinates)
    const double cap_cells = (double)(o.max_cells > 0 ? o.max_cells : ((int64_t)1 << 27));
    for (;;) {                                        // honour the cap exactly
        double tot = 1.0;
        for (int a = 0; a < 3; ++a) tot *= floor(ext[a] / cell) + 1.0;
        if (tot <= cap_cells) break;
        cell *= 1.05;
    }
    g.cell = cell;
    g.inv_cell = 1.0 / cell;
    for (int a = 0; a < 3; ++a) {
        g.origin[a] = m->bbox_lo[a];
        g.dims[0][a] = (int32_t)floor(ext[a] / cell) + 1;
    }
    int L = 1;
    while (!(g.dims[L - 1][0] == 1 && g.dims[L - 1][1] == 1 && g.dims[L - 1][2] == 1)) {
        PCREG_REQUIRE(L < GRID_MAX_LEVELS, "grid: too many pyramid levels");
        for (int a = 0; a < 3; ++a) g.dims[L][a] = (g.dims[L - 1][a] + 1) / 2;
        ++L;
    }
    g.nlevels = L;
    const int64_t ncell0 = (int64_t)g.dims[0][0] * g.dims[0][1] * g.dims[0][2];
    PCREG_REQUIRE(ncell0 < ((int64_t)1 << 31) - 2 && n < ((int64_t)1 << 31) - 2, "grid: too many cells or points for int32 indexing");

    m->g_cell_start.alloc((size_t)ncell0 + 1);
    m->g_pts.alloc((size_t)n);
    DevBuf<int32_t> cell_of((size_t)n), cursor((size_t)ncell0);
    const int nb_scan = (int)((ncell0 + SCAN_BLOCK - 1) / SCAN_BLOCK);
    DevBuf<int32_t> sums((size_t)nb_scan + 1), total(1);
    DevBuf<unsigned long long> occ(1);
    PCREG_CUDA(cudaMemsetAsync(m->g_cell_start.p, 0, m->g_cell_start.bytes(), st));
    PCREG_CUDA(cudaMemsetAsync(cursor.p, 0, cursor.bytes(), st));
    PCREG_CUDA(cudaMemsetAsync(occ.p, 0, sizeof(unsigned long long), st));

    const int T = 256;
    k_grid_count<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(m->md.p, n, g, cell_of.p, m->g_cell_start.p);
    PCREG_LAUNCHED();
    k_scan_local<<<nb_scan, SCAN_THREADS, 0, st>>>(m->g_cell_start.p, ncell0, sums.p);
    PCREG_LAUNCHED();
    k_scan_sums<<<1, SCAN_THREADS, 0, st>>>(sums.p, nb_scan, total.p);
    PCREG_LAUNCHED();
    k_scan_add<<<(unsigned)((ncell0 + 1 + T - 1) / T), T, 0, st>>>(m->g_cell_start.p, ncell0, sums.p, total.p);
    PCREG_LAUNCHED();
    k_grid_scatter<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(m->md.p, n, cell_of.p, m->g_cell_start.p, cursor.p, m->g_pts.p);
    PCREG_LAUNCHED();

    g.pts = m->g_pts.p;
    g.cell_start = m->g_cell_start.p;
    m->g_masks.clear();
    m->g_masks.resize(L);
    for (int l = 0; l < GRID_MAX_LEVELS; ++l) g.mask[l] = nullptr;
    for (int l = 1; l < L; ++l) {
        const int64_t nc = (int64_t)g.dims[l][0] * g.dims[l][1] * g.dims[l][2];
        m->g_masks[l].alloc((size_t)nc);
        k_grid_mask<<<(unsigned)((nc + T - 1) / T), T, 0, st>>>(g, l, m->g_cell_start.p, l > 1 ? m->g_masks[l - 1].p : nullptr,
                                                                m->g_masks[l].p, occ.p);
        PCREG_LAUNCHED();
        g.mask[l] = m->g_masks[l].p;
    }
    unsigned long long h_occ = 0;
    PCREG_CUDA(cudaMemcpyAsync(&h_occ, occ.p, sizeof h_occ, cudaMemcpyDeviceToHost, st));
    PCREG_CUDA(cudaStreamSynchronize(st));
    m->g_occupied = (L > 1) ? (int64_t)h_occ : (n > 0 ? 1 : 0);
    m->grid = g;
    m->has_grid = true;
}

}  // namespace pcreg

using namespace pcreg;

// ------------------------------------------------------------------------------------------------
// C ABI: lifetime + model handle
// ------------------------------------------------------------------------------------------------
extern "C" {

int pcreg_abi_version(void) { return 2; }
const char* pcreg_last_error(void) { return g_err[0] ? g_err : g_err_global; }
int64_t pcreg_launch_count(void) { return (int64_t)g_launches.load(); }

int pcreg_init(const int* devices, int ndev) {
    PCREG_API_BEGIN
    PCREG_REQUIRE(ndev >= 0 && ndev <= PCREG_MAX_DEVICES, "pcreg_init: ndev out of range");
    PCREG_REQUIRE(ndev <= 1 || devices, "pcreg_init: devices is null");
    if (g_nslots > 0) pcreg_shutdown();                 // re-initialisation: drop the previous device set
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("pcreg_init: no usable CUDA device (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return PCREG_ERR_CUDA;
    }
    const int n = std::max(ndev, 1);
    for (int k = 0; k < n; ++k) {
        const int dev = (devices && ndev >= 1) ? devices[k] : 0;
        PCREG_REQUIRE(dev >= 0 && dev < count, "pcreg_init: device ordinal out of range");
        // (PCREG_ALLOW_DUP_DEVICES=1: the tests run the multi-slot path -- worker threads, per-slot pools and streams, replicas --
        // with two slots on the one GPU of a single-GPU box)
        const char* dup_e = getenv("PCREG_ALLOW_DUP_DEVICES");
        const bool allow_dup = dup_e && dup_e[0] == '1';
        for (int j = 0; j < k; ++j) PCREG_REQUIRE(allow_dup || devices[j] != dev, "pcreg_init: the same device listed twice");
        PCREG_CUDA(cudaSetDevice(dev));
        cudaDeviceProp p;
        PCREG_CUDA(cudaGetDeviceProperties(&p, dev));
        if (p.major < 10) {
            set_error("pcreg_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, p.major, p.minor);
            return PCREG_ERR_CUDA;
        }
        PCREG_CUDA(cudaFree(0));                        // create the context now, not inside the first timed call
        Context& c = g_ctxs[k];
        c = Context();
        c.slot = k;
        c.device = dev;
        c.sm_count = p.multiProcessorCount;
        c.smem_optin = p.sharedMemPerBlockOptin;
        c.total_mem = p.totalGlobalMem;
        c.initialised = true;
    }
    g_nslots = n;
    tl_slot = 0;
    PCREG_CUDA(cudaSetDevice(g_ctxs[0].device));
    g_launches.store(0);
    return PCREG_OK;
    PCREG_API_END
}

int pcreg_device_count(void) { return g_nslots; }

int pcreg_shutdown(void) {
    for (int k = 0; k < g_nslots; ++k) {
        Context& c = g_ctxs[k];
        if (!c.initialised) continue;
        cudaSetDevice(c.device);
        cudaDeviceSynchronize();
        for (cudaEvent_t e : c.events) cudaEventDestroy(e);
        c.events.clear();
        for (cudaStream_t st : c.streams) cudaStreamDestroy(st);
        c.streams.clear();
        for (int b = 0; b < 2; ++b) {
            if (c.pinned[b]) { cudaFreeHost(c.pinned[b]); c.pinned[b] = nullptr; }
            if (c.pinned_ev[b]) { cudaEventDestroy(c.pinned_ev[b]); c.pinned_ev[b] = nullptr; }
        }
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            pool_trim_locked(c, 0);
        }
        c.initialised = false;
    }
    g_nslots = 0;
    tl_slot = 0;
    return PCREG_OK;
}

int pcreg_set_profiling(int enabled) {
    for (int k = 0; k < PCREG_MAX_DEVICES; ++k) { g_ctxs[k].profiling = enabled != 0; g_ctxs[k].profiling_counters = enabled != 2; }
    return PCREG_OK;
}
int pcreg_last_profile(double out[32]) {
    if (!out) return PCREG_ERR_ARG;
    for (int i = 0; i < 32; ++i) out[i] = ctx().profile[i];
    return PCREG_OK;
}

// one device-resident copy of the model on the calling thread's slot
static pcreg_model* model_build_on_slot(const std::vector<ModelPointD>& h, const std::vector<int32_t>& perm, const double* lo,
                                        const double* hi, float max_norm, const pcreg_model_opts& o) {
    const int64_t n = (int64_t)h.size();
    std::unique_ptr<pcreg_model> m(new pcreg_model());
    m->slot = current_slot();
    m->n = n;
    m->n_pad = ((n + BRUTE_TILE - 1) / BRUTE_TILE) * BRUTE_TILE;
    for (int a = 0; a < 3; ++a) { m->bbox_lo[a] = lo[a]; m->bbox_hi[a] = hi[a]; m->pivot[a] = 0.5 * (lo[a] + hi[a]); }
    m->max_norm = max_norm;
    cudaStream_t st = 0;
    m->md.alloc((size_t)n);
    m->perm.alloc((size_t)m->n_pad);
    m->m4.alloc((size_t)m->n_pad);
    PCREG_CUDA(cudaMemcpyAsync(m->md.p, h.data(), (size_t)n * sizeof(ModelPointD), cudaMemcpyHostToDevice, st));
    PCREG_CUDA(cudaMemsetAsync(m->perm.p, 0xff, m->perm.bytes(), st));
    PCREG_CUDA(cudaMemcpyAsync(m->perm.p, perm.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    k_build_m4<<<(unsigned)((m->n_pad + 255) / 256), 256, 0, st>>>(m->md.p, m->perm.p, n, m->n_pad, m->pivot[0], m->pivot[1], m->pivot[2], m->m4.p);
    PCREG_LAUNCHED();
    if (o.build_grid) {
        grid_build(m.get(), o, st);
        vox_build(m.get(), o, st);
    }
    PCREG_CUDA(cudaStreamSynchronize(st));
    return m.release();
}

int pcreg_model_create(const void* xyz, int is_double, int64_t n, int64_t ld, const pcreg_model_opts* opts, pcreg_model** out) {
    PCREG_API_BEGIN
    require_init();
    PCREG_REQUIRE(xyz && out, "pcreg_model_create: null pointer");
    PCREG_REQUIRE(n >= 1 && ld >= n, "pcreg_model_create: need n >= 1 and ld >= n");
    PCREG_REQUIRE(n < ((int64_t)1 << 31) - BRUTE_TILE, "pcreg_model_create: model too large for int32 indices");
    pcreg_model_opts o{};
    if (opts) o = *opts;
    use_slot(0);

    // host staging: FP64 AoS in original order, bounding box, pivot, FP32 norm bound
    std::vector<ModelPointD> h((size_t)n);
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < n; ++i) {
        double v[3];
        for (int a = 0; a < 3; ++a)
            v[a] = is_double ? ((const double*)xyz)[a * ld + i] : (double)((const float*)xyz)[a * ld + i];
        PCREG_REQUIRE(std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]), "pcreg_model_create: non-finite model coordinate");
        h[i].x = v[0]; h[i].y = v[1]; h[i].z = v[2]; h[i].pad = 0.0;
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], v[a]); hi[a] = std::max(hi[a], v[a]); }
    }
    double maxn2 = 0.0, pivot[3];
    for (int a = 0; a < 3; ++a) pivot[a] = 0.5 * (lo[a] + hi[a]);
    for (int64_t i = 0; i < n; ++i) {
        const float x = (float)(h[i].x - pivot[0]), y = (float)(h[i].y - pivot[1]), z = (float)(h[i].z - pivot[2]);
        maxn2 = std::max(maxn2, (double)x * x + (double)y * y + (double)z * z);
    }
    const float max_norm = (float)(sqrt(maxn2) * (1.0 + 1e-6));
    // scan-order permutation (Fisher-Yates, splitmix64)
    std::vector<int32_t> perm((size_t)n);
    for (int64_t i = 0; i < n; ++i) perm[i] = (int32_t)i;
    uint64_t s = o.shuffle_seed + 0x9E3779B97F4A7C15ull;
    auto next = [&s]() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
    for (int64_t i = n - 1; i > 0; --i) {
        const int64_t j = (int64_t)(next() % (uint64_t)(i + 1));
        std::swap(perm[i], perm[j]);
    }
    // the model is replicated on every selected device: slots >= 1 build their copy on worker threads
    const int ns = num_slots();
    std::vector<pcreg_model*> copies((size_t)ns, nullptr);
    std::vector<int> rc((size_t)ns, PCREG_OK);
    std::vector<std::thread> workers;
    for (int k = 1; k < ns; ++k)
        workers.emplace_back([&, k]() { rc[k] = guarded([&]() { use_slot(k); copies[k] = model_build_on_slot(h, perm, lo, hi, max_norm, o); }); });
    rc[0] = guarded([&]() { copies[0] = model_build_on_slot(h, perm, lo, hi, max_norm, o); });
    for (auto& w : workers) w.join();
    use_slot(0);
    for (int k = 0; k < ns; ++k)
        if (rc[k] != PCREG_OK) {
            for (int j = 0; j < ns; ++j) if (copies[j]) { use_slot(j); cudaDeviceSynchronize(); delete copies[j]; }
            use_slot(0);
            return rc[k];
        }
    for (int k = 1; k < ns; ++k) copies[0]->replicas.push_back(copies[k]);
    *out = copies[0];
    return PCREG_OK;
    PCREG_API_END
}

int pcreg_model_destroy(pcreg_model* m) {
    PCREG_API_BEGIN
    if (m) {
        for (pcreg_model* r : m->replicas) { use_slot(r->slot); cudaDeviceSynchronize(); delete r; }
        use_slot(m->slot);
        cudaDeviceSynchronize();
        delete m;
        use_slot(0);
    }
    return PCREG_OK;
    PCREG_API_END
}

int64_t pcreg_model_size(const pcreg_model* m) { return m ? m->n : -1; }

int pcreg_model_grid_info(const pcreg_model* m, int32_t dims[3], double* cell_size, int64_t* occupied) {
    if (!m || !m->has_grid) { set_error("pcreg_model_grid_info: model has no grid"); return PCREG_ERR_STATE; }
    if (dims) for (int a = 0; a < 3; ++a) dims[a] = m->grid.dims[0][a];
    if (cell_size) *cell_size = m->grid.cell;
    if (occupied) *occupied = m->g_occupied;
    return PCREG_OK;
}

}  // extern "C"
