// Inner-loop microbenchmark for the brute-force NN kernel: smem-resident model tile read as broadcast
// LDS.128, Q register-resident queries per thread, group minimum via FMNMX3.  Scalar FFMA vs packed
// FFMA2 (fma.rn.f32x2).  "useful" = 6 FLOP per (query, point) pair.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nn_microbench nn_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int TILE = 1024;
constexpr int REPS = 64;      // passes over the tile per launch
constexpr int G = 16;

__device__ __forceinline__ float fmin3(float a, float b, float c){ float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int Q>
__global__ void __launch_bounds__(256, 2) k_scalar(const float4* __restrict__ m, float* out, float thr0){
    __shared__ __align__(16) float4 t[TILE];
    for (int i = threadIdx.x; i < TILE; i += blockDim.x) t[i] = m[i];
    __syncthreads();
    float ax[Q], ay[Q], az[Q], thr[Q], acc[Q];
    for (int k = 0; k < Q; k++){ ax[k] = 0.001f*(threadIdx.x+k); ay[k] = 0.002f*k - 1.f; az[k] = 0.5f + 0.01f*k; thr[k] = thr0; acc[k] = 1e30f; }
    #pragma unroll 1
    for (int r = 0; r < REPS; r++){
        #pragma unroll 1
        for (int g0 = 0; g0 < TILE; g0 += G){
            float gm[Q];
            #pragma unroll
            for (int jj = 0; jj < G; jj += 2){
                const float4 m0 = t[g0+jj], m1 = t[g0+jj+1];
                #pragma unroll
                for (int k = 0; k < Q; k++){
                    float d0 = fmaf(ax[k], m0.x, m0.w), d1 = fmaf(ax[k], m1.x, m1.w);
                    d0 = fmaf(ay[k], m0.y, d0); d1 = fmaf(ay[k], m1.y, d1);
                    d0 = fmaf(az[k], m0.z, d0); d1 = fmaf(az[k], m1.z, d1);
                    gm[k] = (jj == 0) ? fminf(d0, d1) : fmin3(gm[k], d0, d1);
                }
            }
            bool hit = false;
            #pragma unroll
            for (int k = 0; k < Q; k++) hit |= (gm[k] <= thr[k]);
            if (hit){
                #pragma unroll
                for (int k = 0; k < Q; k++) if (gm[k] <= thr[k]) { acc[k] = fminf(acc[k], gm[k]); thr[k] = gm[k] - 1.f; }
            }
        }
    }
    float s = 0; for (int k = 0; k < Q; k++) s += acc[k] + thr[k];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

// packed: tile layout per point pair: {x0,x1,y0,y1},{z0,z1,w0,w1}
template <int Q>
__global__ void __launch_bounds__(256, 2) k_packed(const float4* __restrict__ m, float* out, float thr0){
    __shared__ __align__(16) float4 t[TILE];
    for (int i = threadIdx.x; i < TILE; i += blockDim.x) t[i] = m[i];
    __syncthreads();
    float2 ax[Q], ay[Q], az[Q]; float thr[Q], acc[Q];
    for (int k = 0; k < Q; k++){ float a = 0.001f*(threadIdx.x+k), b = 0.002f*k - 1.f, c = 0.5f + 0.01f*k; ax[k] = make_float2(a,a); ay[k] = make_float2(b,b); az[k] = make_float2(c,c); thr[k] = thr0; acc[k] = 1e30f; }
    #pragma unroll 1
    for (int r = 0; r < REPS; r++){
        #pragma unroll 1
        for (int g0 = 0; g0 < TILE; g0 += G){
            float gm[Q];
            #pragma unroll
            for (int jj = 0; jj < G; jj += 2){
                const float4 A = t[g0+jj], B = t[g0+jj+1];
                const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w), Z = make_float2(B.x, B.y), W = make_float2(B.z, B.w);
                #pragma unroll
                for (int k = 0; k < Q; k++){
                    float2 d = __ffma2_rn(ax[k], X, W);
                    d = __ffma2_rn(ay[k], Y, d);
                    d = __ffma2_rn(az[k], Z, d);
                    gm[k] = (jj == 0) ? fminf(d.x, d.y) : fmin3(gm[k], d.x, d.y);
                }
            }
            bool hit = false;
            #pragma unroll
            for (int k = 0; k < Q; k++) hit |= (gm[k] <= thr[k]);
            if (hit){
                #pragma unroll
                for (int k = 0; k < Q; k++) if (gm[k] <= thr[k]) { acc[k] = fminf(acc[k], gm[k]); thr[k] = gm[k] - 1.f; }
            }
        }
    }
    float s = 0; for (int k = 0; k < Q; k++) s += acc[k] + thr[k];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

template <typename K> double timeit(K kern, const float4* m, float* out, int blocks, int reps){
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) kern<<<blocks, 256>>>(m, out, -1e30f);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) kern<<<blocks, 256>>>(m, out, -1e30f);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
    return ms / reps * 1e-3;
}

int main(){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    float4* hm = (float4*)malloc(sizeof(float4)*TILE);
    for (int i = 0; i < TILE; i++) hm[i] = make_float4(0.01f*i, 1.f-0.02f*i, 0.5f, 3.f+i);
    float4* m; float* out; CK(cudaMalloc(&m, sizeof(float4)*TILE)); CK(cudaMalloc(&out, sizeof(float)*sms*8*256));
    CK(cudaMemcpy(m, hm, sizeof(float4)*TILE, cudaMemcpyHostToDevice));
    for (int bps : {1, 2}) {
        const int blocks = sms * bps;
        const double thr = (double)blocks * 256;
        #define RUN(name, kern, Q) { double t = timeit(kern, m, out, blocks, 10); printf("%-14s Q=%2d blocks/SM=%d : %7.2f TF useful\n", name, Q, bps, thr * Q * (double)TILE * REPS * 6 / t * 1e-12); }
        RUN("scalar", k_scalar<4>, 4); RUN("scalar", k_scalar<8>, 8);
        RUN("packed", k_packed<4>, 4); RUN("packed", k_packed<8>, 8); RUN("packed", k_packed<12>, 12);
    }
    return 0;
}
