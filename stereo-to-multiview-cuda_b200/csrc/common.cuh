// common.cuh — shared device helpers and layout constants (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace s2mv {

// The reference's cost-initialisation kernels run in 160-wide blocks
// (d_ci_adcensus.cu:46); two of their shared-memory indices step one slot
// outside their half at the block edges.  Parity needs the same
// columns special-cased, whatever tiling this implementation uses.
constexpr int kRefBlockW = 160;
constexpr int kAdLutSize = 766;  // 3*255+1
constexpr int kCenLutSize = 65;  // Hamming 0..64 (d_alu.cu:7-15)

__host__ __device__ inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// packed arms: byte0 = UP, byte1 = DOWN, byte2 = LEFT, byte3 = RIGHT
__device__ __forceinline__ int arm_up(uint32_t a) { return a & 0xff; }
__device__ __forceinline__ int arm_down(uint32_t a) { return (a >> 8) & 0xff; }
__device__ __forceinline__ int arm_left(uint32_t a) { return (a >> 16) & 0xff; }
__device__ __forceinline__ int arm_right(uint32_t a) { return a >> 24; }

// Truncated Hamming distance of the reference (d_alu.cu:7-15): only the low
// 32 bits of the XOR survive `int c = a ^ b`; bits 0..30 count once, the sign
// bit is seen by 33 of the 64 arithmetic shifts.
__device__ __forceinline__ int ref_hamdist32(uint32_t a, uint32_t b)
{
    uint32_t x = a ^ b;
    return __popc(x & 0x7fffffffu) + 33 * (int)(x >> 31);
}

// 1 - ex2.approx((-c * inv) * log2e): the instruction sequence nvcc emits for
// `1.0 - __expf(-c*inv)` in ci_adcensus_kernel (d_ci_adcensus.cu:27-31).
__device__ __forceinline__ float ref_one_minus_exp(float c, float inv)
{
    float t = __fmul_rn(-c, inv);
    t = __fmul_rn(t, 1.4426950408889634f);  // 0x3FB8AA3B
    float e;
    asm("ex2.approx.f32 %0, %1;" : "=f"(e) : "f"(t));
    return __fsub_rn(1.0f, e);
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Exact-order accumulate: plain IEEE adds, never contracted.  sm_100a's packed add (FADD2: two
// independent round-to-nearest fp32 adds per lane, one issue slot) halves the issue cost of the
// aggregation loops, whose ceiling is instruction issue, not the FP32 pipe (tools/ubench_fadd2.cu).
__device__ __forceinline__ void add2_rn(float &a0, float &a1, float v0, float v1)
{
    unsigned long long A, V;
    asm("mov.b64 %0, {%1,%2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(V) : "f"(v0), "f"(v1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(A) : "l"(V));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(A));
}
// Two independent fp32 operations per lane in one issue slot (sm_100a f32x2 forms); each half rounds exactly
// like the scalar __f*_rn.  A pair lives in a 64-bit register: low word = first element.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b)
{
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b)
{
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c)
{
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ void acc4(float4 &a, const float4 v)
{
#ifdef S2MV_NO_FADD2
    a.x = __fadd_rn(a.x, v.x);
    a.y = __fadd_rn(a.y, v.y);
    a.z = __fadd_rn(a.z, v.z);
    a.w = __fadd_rn(a.w, v.w);
#else
    add2_rn(a.x, a.y, v.x, v.y);
    add2_rn(a.z, a.w, v.z, v.w);
#endif
}

}  // namespace s2mv
