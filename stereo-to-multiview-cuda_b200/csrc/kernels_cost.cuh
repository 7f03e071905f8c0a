// kernels_cost.cuh — the cost-volume stages: ADCensus cost initialisation,
// 4-pass cross aggregation (H, V, V, H) and winner-takes-all.
//
// Volume layout in HBM: [view][y][x][Dp] fp32, disparity innermost, Dp = D
// padded (to a power of two <= 128, or to a multiple of 128).  One pixel's
// disparity vector is one contiguous 16*LP-byte run (LP = Dp/4 float4 lanes), so
//   * a group of LP lanes owns one pixel: all of them share the pixel's arms,
//     the window loops have warp-uniform trip counts and no lane is wasted
//     on a neighbour's longer arm (D = 128: one warp per pixel);
//   * every global access is a full-line 128-bit vector access.
// Sums are the reference's sequential ascending fp32 adds from 0
// (d_ca_cross_sum.cu:282-290): bit-exact, never reassociated or contracted.
//
// Kernels here: what the stage entry points need besides the line kernel of kernels_line.cuh (which
// does the frame path's four passes):
//   k_hpass<MODE,...>     one image row segment per CTA: builds the AD / census / ADCensus cost tile in shared
//                         memory and stores it (s2mv_ci_ad, s2mv_ci_census, s2mv_ci_adcensus)
//   k_wta_finish          D > 128: (cost, d) keys of the chunked WTA -> disparities
//   k_planes_to_vol / k_vol_to_planes / k_wta_planes   the reference's plane tables <-> the volume layout
#pragma once
#include <float.h>

#include "common.cuh"

namespace s2mv {

constexpr int kHThreads = 256;

struct HArgs {
    // cost-initialisation inputs (FROM_CI)
    const uint32_t *pixL, *pixR, *cenL, *cenR;
    const float *lutAd, *lutCen;
    // volume in/out, per view slot
    const float4 *in[2];
    float4 *out[2];
    const uint32_t *arms[2];
    float *disp[2];
    unsigned long long *wta_key[2];  // multi-chunk WTA (D > 128)
    int H, W, D, zd;
    int LP;       // float4 lanes per pixel handled by one CTA (<= 32, power of two)
    int LPtot;    // float4 lanes per pixel in the volume (= LP * nchunks)
    int nchunks;  // disparity chunks of 4*LP
    int S;        // segment width in pixels
    int halo;     // = usd when summing, else 0
    int M;        // CI margin: max(zd, D-1-zd)
    int view_first;
};

// Reference operand emulation for the block-edge columns.  The
// reference indexes a flat shared array [left row | right row]; at tx = 0 and
// tx = 159 (and, for zero_disp = 0, for the anchor itself) the index leaves its
// half.  Returns the four operands exactly as the reference kernels read them.
struct CiOperands { uint32_t ad_own, ad_other, cen_own, cen_other; };

__device__ __forceinline__ CiOperands
ref_ci_operands(int view, int gx, int d, int D, int zd, int W, const uint32_t *__restrict__ rowPixL,
                const uint32_t *__restrict__ rowPixR, const uint32_t *__restrict__ rowCenL,
                const uint32_t *__restrict__ rowCenR)
{
    CiOperands o;
    const int tx = gx % kRefBlockW, bs = gx - tx;
    {   // ci_ad_kernel_5 with the launch parameters of d_ci_adcensus.cu:57-59,109
        const int pad = (D - zd > zd) ? (D - zd) : (zd - 1);
        const int smc = kRefBlockW + 2 * pad;
        o.ad_own = (view == 0 ? rowPixL : rowPixR)[gx];
        int f = (view == 0) ? (smc + tx + pad + (d - zd)) : (tx + pad - (d - zd));
        o.ad_other = (f < smc) ? rowPixL[clampi(bs - pad + f, 0, W - 1)]
                               : rowPixR[clampi(bs - pad + f - smc, 0, W - 1)];
    }
    {   // ci_census_kernel_6 with d_ci_adcensus.cu:117-120
        const int smc = kRefBlockW + D - 1, padl = zd - 1, padr = D - zd;
        int fo = (view == 0) ? (tx + padr) : (smc + tx + padl);
        int fx = (view == 0) ? (smc + tx + padl + (d - zd)) : (tx + padr - (d - zd));
        o.cen_own = (fo < smc) ? rowCenL[clampi(bs - padr + fo, 0, W - 1)]
                               : rowCenR[clampi(bs - padl + fo - smc, 0, W - 1)];
        o.cen_other = (fx < smc) ? rowCenL[clampi(bs - padr + fx, 0, W - 1)]
                                 : rowCenR[clampi(bs - padl + fx - smc, 0, W - 1)];
    }
    return o;
}

// order-preserving map float -> uint32 (handles any sign), for the atomicMin WTA
__device__ __forceinline__ uint32_t float_orderable(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// CI_MODE: 0 = tile loaded from the volume; 1 = ADCensus combined cost;
//          2 = AD cost only (float(sum)*0.33333334f); 3 = Hamming cost only.
// (2 and 3 exist for the stage-parity entry points s2mv_ci_ad / s2mv_ci_census.)
template <int CI_MODE, bool DO_SUM, bool STORE, bool DO_WTA>
__global__ void __launch_bounds__(kHThreads)
k_hpass(const HArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int y = blockIdx.y;
    const int vslot = blockIdx.z / a.nchunks, chunk = blockIdx.z % a.nchunks;
    const int view = a.view_first + vslot;
    const int LP = a.LP, Dc = 4 * LP, d0 = chunk * Dc;
    const int W = a.W, D = a.D, zd = a.zd;
    const int x0 = blockIdx.x * a.S;
    const int S = min(a.S, W - x0);
    const int halo = DO_SUM ? a.halo : 0;
    const int P = a.S + 2 * halo;  // tile positions; position p <-> column x0 - halo + p
    float *C = reinterpret_cast<float *>(smem_raw);
    float4 *C4 = reinterpret_cast<float4 *>(smem_raw);

    if (CI_MODE != 0) {
        const int M = a.M, NP = P + 2 * M;
        uint32_t *sPL = reinterpret_cast<uint32_t *>(C + (size_t)P * Dc);
        uint32_t *sPR = sPL + NP, *sCL = sPR + NP, *sCR = sCL + NP;
        float *sLutAd = reinterpret_cast<float *>(sCR + NP);
        float *sLutCen = sLutAd + 768;
        const int xb = x0 - halo - M;
        const size_t row = (size_t)y * W;
        for (int i = tid; i < NP; i += kHThreads) {
            size_t g = row + clampi(xb + i, 0, W - 1);
            sPL[i] = a.pixL[g];
            sPR[i] = a.pixR[g];
            sCL[i] = a.cenL[g];
            sCR[i] = a.cenR[g];
        }
        for (int i = tid; i < kAdLutSize; i += kHThreads) sLutAd[i] = a.lutAd[i];
        if (tid < kCenLutSize) sLutCen[tid] = a.lutCen[tid];
        __syncthreads();

        const int warp = tid >> 5, lane = tid & 31;
        const uint32_t *ownPix = view == 0 ? sPL : sPR, *othPix = view == 0 ? sPR : sPL;
        const uint32_t *ownCen = view == 0 ? sCL : sCR, *othCen = view == 0 ? sCR : sCL;
        const int sgn = view == 0 ? 1 : -1;
        for (int p = warp; p < P; p += kHThreads / 32) {
            const int gx = x0 - halo + p;
            if (gx < 0 || gx >= W) continue;  // never inside any window (arms stop at the border)
            const int io = p + M;
            const uint32_t op = ownPix[io], oc = ownCen[io];
            const int tx = gx % kRefBlockW;
            const bool edge = (tx == 0) || (tx == kRefBlockW - 1);
            for (int dl = lane; dl < Dc; dl += 32) {
                const int d = d0 + dl;
                float c = 0.0f;
                if (d < D) {
                    uint32_t p_own = op, c_own = oc;
                    uint32_t p_oth = othPix[io + sgn * (d - zd)], c_oth = othCen[io + sgn * (d - zd)];
                    if (edge) {  // 2 of 160 columns: replay the reference's flat indexing
                        CiOperands o = ref_ci_operands(view, gx, d, D, zd, W, a.pixL + row, a.pixR + row,
                                                       a.cenL + row, a.cenR + row);
                        p_own = o.ad_own; p_oth = o.ad_other; c_own = o.cen_own; c_oth = o.cen_other;
                    }
                    const int sad = (int)__vsadu4(p_own, p_oth);  // x byte is 0 in both
                    const int ham = ref_hamdist32(c_own, c_oth);
                    if (CI_MODE == 1) c = __fadd_rn(sLutAd[sad], sLutCen[ham]);
                    else if (CI_MODE == 2) c = __fmul_rn((float)sad, 0.33333333333f);
                    else c = (float)ham;
                }
                C[(size_t)p * Dc + dl] = c;
            }
        }
    } else {
        // tile <- volume rows, 16 B per cp.async, whole 64*LP/... byte lines per pixel
        const float4 *src = a.in[vslot] + (size_t)y * W * a.LPtot + (size_t)chunk * LP;
        for (int i = tid; i < P * LP; i += kHThreads) {
            const int p = i / LP, q = i - p * LP;
            const int gx = x0 - halo + p;
            if (gx >= 0 && gx < W) cp_async16(&C4[i], src + (size_t)gx * a.LPtot + q);
        }
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();

    const int G = kHThreads / LP;  // pixels in flight per sweep
    const int grp = tid / LP, q = tid - grp * LP;
    const uint32_t *arms = a.arms[vslot] + (size_t)y * W;
    for (int pb = 0; pb < S; pb += G) {
        const int px = pb + grp;
        const bool active = px < S;
        const int gx = x0 + px;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            if (DO_SUM) {
                const uint32_t ar = arms[gx];
                const int L = arm_left(ar), n = L + arm_right(ar);
                const float4 *c = C4 + (size_t)(px + halo - L) * LP + q;
                for (int k = 0; k < n; ++k) acc4(acc, c[(size_t)k * LP]);  // window [x-L, x+R)
            } else {
                acc = C4[(size_t)px * LP + q];
            }
            if (STORE) a.out[vslot][((size_t)y * W + gx) * a.LPtot + (size_t)chunk * LP + q] = acc;
        }
        if (DO_WTA) {
            // dc_wta_kernel (d_dc_wta.cu:9-35): strict '>' from FLT_MAX, first minimum wins
            float best = FLT_MAX;
            int bestd = 0x7fffffff;
            if (active) {
                const int d = d0 + 4 * q;
                if (d + 0 < D && best > acc.x) { best = acc.x; bestd = d; }
                if (d + 1 < D && best > acc.y) { best = acc.y; bestd = d + 1; }
                if (d + 2 < D && best > acc.z) { best = acc.z; bestd = d + 2; }
                if (d + 3 < D && best > acc.w) { best = acc.w; bestd = d + 3; }
            }
            for (int off = LP >> 1; off > 0; off >>= 1) {
                float ov = __shfl_xor_sync(0xffffffffu, best, off);
                int od = __shfl_xor_sync(0xffffffffu, bestd, off);
                if (ov < best || (ov == best && od < bestd)) { best = ov; bestd = od; }
            }
            if (active && q == 0) {
                if (bestd == 0x7fffffff) bestd = 0;
                if (a.nchunks == 1) {
                    a.disp[vslot][(size_t)y * W + gx] = (float)bestd - (float)zd;
                } else {
                    unsigned long long key = ((unsigned long long)float_orderable(best) << 32) | (uint32_t)bestd;
                    atomicMin(a.wta_key[vslot] + (size_t)y * W + gx, key);
                }
            }
        }
    }
}

// multi-chunk WTA epilogue: key -> disparity
__global__ void k_wta_finish(const unsigned long long *__restrict__ key, float *__restrict__ disp, int zd, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    disp[i] = (float)(int)(uint32_t)(key[i] & 0xffffffffull) - (float)zd;
}

// Stage-API layout converters: D contiguous planes <-> [pixel][Dp] volume.
__global__ void k_planes_to_vol(const float *__restrict__ planes, float *__restrict__ vol, int D, int Dp, size_t n)
{
    __shared__ float tile[32][33];
    const size_t p0 = (size_t)blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int d = d0 + j;
        const size_t p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (d < D && p < n) ? planes[(size_t)d * n + p] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const size_t p = p0 + j;
        const int d = d0 + threadIdx.x;
        if (p < n && d < Dp) vol[p * Dp + d] = tile[threadIdx.x][j];
    }
}
__global__ void k_vol_to_planes(const float *__restrict__ vol, float *__restrict__ planes, int D, int Dp, size_t n)
{
    __shared__ float tile[32][33];
    const size_t p0 = (size_t)blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const size_t p = p0 + j;
        const int d = d0 + threadIdx.x;
        tile[j][threadIdx.x] = (p < n && d < Dp) ? vol[p * Dp + d] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int d = d0 + j;
        const size_t p = p0 + threadIdx.x;
        if (d < D && p < n) planes[(size_t)d * n + p] = tile[threadIdx.x][j];
    }
}

// dc_wta_kernel on the reference's plane layout (stage API)
__global__ void __launch_bounds__(256)
k_wta_planes(const float *__restrict__ planes, float *__restrict__ disp, int D, int zd, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lowest = FLT_MAX, lowest_d = 0.0f;
    for (int d = 0; d < D; ++d) {
        float c = planes[(size_t)d * n + i];
        if (lowest > c) { lowest = c; lowest_d = (float)d; }
    }
    disp[i] = lowest_d - (float)zd;
}

}  // namespace s2mv
