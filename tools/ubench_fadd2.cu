// Microbenchmark: FADD vs FADD2 (add.rn.f32x2) issue/throughput on sm_100a, alone and with the
// cost-volume inner-loop mix (1 LDS.128 per 16 lane-adds).  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void add2(float &a, float &b, float x, float y)
{
    u64 A, X;
    asm("mov.b64 %0, {%1,%2};" : "=l"(A) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1,%2};" : "=l"(X) : "f"(x), "f"(y));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(A) : "l"(X));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(A));
}
template <int MODE>  // 0 FADD regs only, 1 FADD2 regs only, 2 LDS+FADD, 3 LDS+FADD2
__global__ void __launch_bounds__(256) k(float4 *out, int iters, float seed)
{
    __shared__ float4 tile[64 * 32];
    for (int i = threadIdx.x; i < 64 * 32; i += 256) tile[i] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float4 acc[4];
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0, 0, 0, 0);
    const float4 *p = tile + (threadIdx.x & 31);
    float4 v = make_float4(seed, seed * 2, seed * 3, seed * 4);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int kk = 0; kk < 64; ++kk) {
            if (MODE >= 2) v = p[kk * 32];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (MODE & 1) { add2(acc[i].x, acc[i].y, v.x, v.y); add2(acc[i].z, acc[i].w, v.z, v.w); }
                else { acc[i].x = __fadd_rn(acc[i].x, v.x); acc[i].y = __fadd_rn(acc[i].y, v.y);
                       acc[i].z = __fadd_rn(acc[i].z, v.z); acc[i].w = __fadd_rn(acc[i].w, v.w); }
            }
        }
    }
    float4 r = make_float4(0, 0, 0, 0);
    for (int i = 0; i < 4; ++i) { r.x += acc[i].x; r.y += acc[i].y; r.z += acc[i].z; r.w += acc[i].w; }
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
template <int MODE> void run(const char *name, float4 *out, int occ)
{
    int dev_sms = 148; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 2000, grid = dev_sms * occ;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 256>>>(out, 10, 1e-30f);
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(out, iters, 1e-30f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double adds = (double)grid * 256 * iters * 64 * 16;
    printf("%-14s occ=%d  %.3f ms  %.2f Tadd/s  %.1f lane-adds/clk/SM (at %d MHz nominal)\n", name, occ, ms,
           adds / ms * 1e-9, adds / (ms * 1e-3) / dev_sms / (clk * 1e3), clk / 1000);
}
int main()
{
    float4 *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float4));
    for (int occ : {2, 4, 8}) {
        run<0>("FADD", out, occ); run<1>("FADD2", out, occ); run<2>("LDS+FADD", out, occ); run<3>("LDS+FADD2", out, occ);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
