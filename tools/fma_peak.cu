// Microbenchmark: FP32 FMA pipe peak on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2),
// and the cost of interleaving FMNMX / 3-input FMNMX3. Decides the brute-force NN inner loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ unsigned long long pack2(float a, float b){
    unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b){
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c){
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c){
    float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int NCH = 8;      // independent chains per thread
constexpr int ITERS = 4096;

// mode 0: scalar FFMA only. flops/thread = ITERS*NCH*3*2
__global__ void k_ffma(float* out, float s){
    float a[NCH], x = s + threadIdx.x * 1e-7f, y = s * 0.5f, z = s * 0.25f;
    for (int i = 0; i < NCH; i++) a[i] = i * 0.1f;
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < NCH; i++) { a[i] = fmaf(a[i], x, y); a[i] = fmaf(a[i], y, z); a[i] = fmaf(a[i], z, x); }
    }
    float r = 0; for (int i = 0; i < NCH; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 1: packed FFMA2 only. flops/thread = ITERS*NCH*3*4
__global__ void k_ffma2(float* out, float s){
    unsigned long long a[NCH], x = pack2(s + threadIdx.x * 1e-7f, s), y = pack2(s * 0.5f, s * 0.3f), z = pack2(s * 0.25f, s*0.1f);
    for (int i = 0; i < NCH; i++) a[i] = pack2(i * 0.1f, i * 0.2f);
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < NCH; i++) { a[i] = ffma2(a[i], x, y); a[i] = ffma2(a[i], y, z); a[i] = ffma2(a[i], z, x); }
    }
    float r = 0; for (int i = 0; i < NCH; i++) { float p, q; unpack2(a[i], p, q); r += p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 2: NN-like scalar: per pair 3 FFMA (fresh chain from w) + 1 FMNMX. "useful" flops = 6/pair
__global__ void k_nn_scalar(float* out, float s){
    float best[NCH], qx[NCH], qy[NCH], qz[NCH];
    for (int i = 0; i < NCH; i++) { best[i] = 1e30f; qx[i] = s * i; qy[i] = s + i; qz[i] = s - i; }
    float mx = s + threadIdx.x * 1e-7f, my = s * 0.5f, mz = s * 0.25f, mw = 3.f;
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < NCH; i++) {
            float d = fmaf(qx[i], mx, mw); d = fmaf(qy[i], my, d); d = fmaf(qz[i], mz, d);
            best[i] = fminf(best[i], d);
        }
        mx += 1.f; my -= 1.f;   // keep the compiler from hoisting
    }
    float r = 0; for (int i = 0; i < NCH; i++) r += best[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 3: NN-like packed: per 2 pairs 3 FFMA2 + 1 FMNMX3. useful flops = 12 per group
__global__ void k_nn_packed3(float* out, float s){
    float best[NCH]; unsigned long long qx[NCH], qy[NCH], qz[NCH];
    for (int i = 0; i < NCH; i++) { best[i] = 1e30f; qx[i] = pack2(s * i, s * i); qy[i] = pack2(s + i, s + i); qz[i] = pack2(s - i, s - i); }
    float fx = s + threadIdx.x * 1e-7f, fy = s * 0.5f;
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        unsigned long long mx = pack2(fx, fx + 1.f), my = pack2(fy, fy - 1.f), mz = pack2(fx, fy), mw = pack2(3.f, fy);
        #pragma unroll
        for (int i = 0; i < NCH; i++) {
            unsigned long long d = ffma2(qx[i], mx, mw); d = ffma2(qy[i], my, d); d = ffma2(qz[i], mz, d);
            float d0, d1; unpack2(d, d0, d1);
            best[i] = fmin3(best[i], d0, d1);
        }
        fx += 1.f; fy -= 1.f;
    }
    float r = 0; for (int i = 0; i < NCH; i++) r += best[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 4: packed + two 2-input FMNMX per group
__global__ void k_nn_packed2(float* out, float s){
    float best[NCH]; unsigned long long qx[NCH], qy[NCH], qz[NCH];
    for (int i = 0; i < NCH; i++) { best[i] = 1e30f; qx[i] = pack2(s * i, s * i); qy[i] = pack2(s + i, s + i); qz[i] = pack2(s - i, s - i); }
    float fx = s + threadIdx.x * 1e-7f, fy = s * 0.5f;
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        unsigned long long mx = pack2(fx, fx + 1.f), my = pack2(fy, fy - 1.f), mz = pack2(fx, fy), mw = pack2(3.f, fy);
        #pragma unroll
        for (int i = 0; i < NCH; i++) {
            unsigned long long d = ffma2(qx[i], mx, mw); d = ffma2(qy[i], my, d); d = ffma2(qz[i], mz, d);
            float d0, d1; unpack2(d, d0, d1);
            best[i] = fminf(best[i], fminf(d0, d1));
        }
        fx += 1.f; fy -= 1.f;
    }
    float r = 0; for (int i = 0; i < NCH; i++) r += best[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 5: FP64 DFMA only
__global__ void k_dfma(float* out, float s){
    double a[NCH], x = s + threadIdx.x * 1e-7, y = s * 0.5, z = s * 0.25;
    for (int i = 0; i < NCH; i++) a[i] = i * 0.1;
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < NCH; i++) { a[i] = fma(a[i], x, y); a[i] = fma(a[i], y, z); a[i] = fma(a[i], z, x); }
    }
    double r = 0; for (int i = 0; i < NCH; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)r;
}

template <typename K> double timeit(K kern, float* out, int blocks, int threads, int reps){
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) kern<<<blocks, threads>>>(out, 1.0001f);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) kern<<<blocks, threads>>>(out, 1.0001f);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
    return ms / reps * 1e-3;
}

int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 32 * 1024));
    const int reps = 20;
    for (int threads : {128, 256, 512, 1024}) for (int bps : {1, 2, 4}) {
        if (threads * bps > 2048) continue;
        int blocks = sms * bps; double nthr = (double)blocks * threads;
        double t0 = timeit(k_ffma, out, blocks, threads, reps);
        double t1 = timeit(k_ffma2, out, blocks, threads, reps);
        double t2 = timeit(k_nn_scalar, out, blocks, threads, reps);
        double t3 = timeit(k_nn_packed3, out, blocks, threads, reps);
        double t4 = timeit(k_nn_packed2, out, blocks, threads, reps);
        double t5 = timeit(k_dfma, out, blocks, threads, reps);
        double base = nthr * ITERS * NCH;
        printf("thr %4d x %d/SM | FFMA %.2f TF | FFMA2 %.2f TF | nn_scalar(3FFMA+FMNMX) %.2f TF useful | nn_packed(3FFMA2+FMNMX3) %.2f TF | nn_packed(3FFMA2+2FMNMX) %.2f TF | DFMA %.2f TF\n",
               threads, bps, base * 6 / t0 * 1e-12, base * 12 / t1 * 1e-12, base * 6 / t2 * 1e-12, base * 12 / t3 * 1e-12, base * 12 / t4 * 1e-12, base * 6 / t5 * 1e-12);
    }
    return 0;
}
