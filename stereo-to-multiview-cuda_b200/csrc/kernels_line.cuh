// kernels_line.cuh — the hot cost-volume kernel: one "line segment" per CTA.
//
// A line is an image row (horizontal passes) or an image column (vertical
// passes).  A CTA owns S consecutive outputs of one line for one view and one
// 128-disparity chunk.  It stages the S + 2*usd input positions of that line in
// shared memory as [position][LP float4] (LP float4 lanes = 4*LP disparities,
// one 16*LP-byte run per position = one contiguous run of the volume), then a
// team of LP lanes produces FOUR consecutive outputs at a time:
//
//     for k in union of the four windows:   v = tile[k]            (one LDS.128)
//         acc_i += v   for every output i whose window holds k     (<= 16 FADD)
//
// so each staged value is read from shared memory about (n+3)/4 times instead
// of n (n = window length): the passes were shared-memory-bandwidth bound, not
// HBM bound, in the first version (profiles/r1_costvol_ncu_full_baseline.txt).
// Every accumulator still receives exactly the reference's additions, ascending
// from 0.0f (d_ca_cross_sum.cu:282-290): a skipped k is predicated off, never
// replaced by "+ 0".  The four windows overlap almost entirely, so the loop is
// split into head (some windows not started: one compare each), core (all four
// open: no compares) and tail (some closed).
//
// Modes:
//   LM_CI_H   tile <- ADCensus cost built in registers (4 pixels x 4
//             disparities per thread: 7 operand words serve 16 evaluations, all
//             shared-memory traffic is conflict-free LDS.128/STS.128), H sum,
//             store.  The initial volume never exists in HBM.
//   LM_H      tile <- volume row, H sum, store            (stage API, pass 4 w/o WTA)
//   LM_H_WTA  tile <- volume row, H sum, winner-takes-all (the final volume never exists)
//   LM_V      tile <- volume column, V sum, store
#pragma once
#include <float.h>

#include "common.cuh"
#include "kernels_cost.cuh"

namespace s2mv {

constexpr int kLineThreads = 256;

enum LineMode { LM_CI_H = 0, LM_H = 1, LM_H_WTA = 2, LM_V = 3 };

struct LineArgs {
    // cost-initialisation inputs (LM_CI_H)
    const uint32_t *pixL, *pixR, *cenL, *cenR;
    const float *lutAd, *lutCen;
    float inv_ad;
    // volume in/out, per view slot
    const float4 *in[2];
    float4 *out[2];
    const uint32_t *arms[2];
    float *disp[2];
    unsigned long long *wta_key[2];  // multi-chunk WTA (D > 128)
    int H, W, D, zd;
    int LPtot;    // float4 lanes per pixel in the volume (= LP * nchunks)
    int nchunks;  // disparity chunks of 4*LP
    int S;        // outputs per CTA along the line (multiple of 4)
    int halo;     // usd
    int view_first;
};

// shared-memory bytes of one CTA
__host__ __device__ inline size_t line_smem_bytes(int S, int halo, int LP, bool ci)
{
    const size_t P4 = ((size_t)S + 2 * halo + 3) & ~(size_t)3;
    size_t b = P4 * LP * 16 + (size_t)S * 4;
    if (ci) b += (2 * P4 + 2 * (P4 + 4 * LP)) * 4 + (68 + 768) * 4;
    return b;
}

// AD half of the combined cost for one evaluation.  Same instruction sequence as
// k_build_luts -> ref_one_minus_exp (d_ci_ad.cu:73-159 + d_ci_adcensus.cu:27-31),
// evaluated in place: the 766-entry table costs ~3.5 bank-conflicted shared
// wavefronts per warp lookup, the arithmetic 7 issue slots and one MUFU.
__device__ __forceinline__ float ad_term(int sad, float inv_ad)
{
    // (float)sad without the conversion unit: exact for sad < 2^23
    const float f = __fsub_rn(__int_as_float(0x4B000000 | sad), 8388608.0f);
    // ref_one_minus_exp with ex2.approx.ftz: the non-ftz form only differs (by its range-scaling
    // prologue/epilogue, 3 extra instructions) when e = ex2(t) is subnormal, and then 1 - e rounds to
    // 1.0f either way.  tests/test_gpu_stages.py compares all 766 values with the table.
    float t = __fmul_rn(-__fmul_rn(f, 0.33333333333f), inv_ad);
    t = __fmul_rn(t, 1.4426950408889634f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return __fsub_rn(1.0f, e);
}

template <int LP, bool PLUS, bool FULL_D>
__device__ __forceinline__ void ci_fill_tile(float4 *__restrict__ C4, const uint32_t *__restrict__ sOwnP,
                                             const uint32_t *__restrict__ sOwnC, const uint32_t *__restrict__ sOthP,
                                             const uint32_t *__restrict__ sOthC, const float *__restrict__ sLutAd,
                                             const float *__restrict__ sLutCen, float inv_ad, int ngrp, int team,
                                             int q, int dbase, int D)
{
    constexpr int TEAMS = kLineThreads / LP;
    constexpr int Dc = 4 * LP;
    bool dv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dv[j] = dbase + j < D;
    for (int g = team; g < ngrp; g += TEAMS) {
        const int p0 = 4 * g;
        const uint4 oP = *reinterpret_cast<const uint4 *>(sOwnP + p0);
        const uint4 oC = *reinterpret_cast<const uint4 *>(sOwnC + p0);
        const int bi = PLUS ? (p0 + 4 * q) : (p0 - 4 * q + Dc - 4);
        const uint4 a0 = *reinterpret_cast<const uint4 *>(sOthP + bi);
        const uint4 a1 = *reinterpret_cast<const uint4 *>(sOthP + bi + 4);
        const uint4 b0 = *reinterpret_cast<const uint4 *>(sOthC + bi);
        const uint4 b1 = *reinterpret_cast<const uint4 *>(sOthC + bi + 4);
        const uint32_t op[4] = {oP.x, oP.y, oP.z, oP.w}, oc[4] = {oC.x, oC.y, oC.z, oC.w};
        const uint32_t wp[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const uint32_t wc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // left view (PLUS): other = R[x + (d - zd)] -> word i + j  (d_ci_ad.cu:133-144);
                // right view: other = L[x - (d - zd)] -> word 4 + i - j
                const int w = PLUS ? (i + j) : (4 + i - j);
                const int sad = (int)__vsadu4(op[i], wp[w]);  // x byte is 0 in both
                const uint32_t x = oc[i] ^ wc[w];
                // ref_hamdist32 (d_alu.cu:7-15) = popc(x) + 32 * bit31(x), as a byte offset into the table
                const uint32_t off = ((uint32_t)__popc(x) << 2) + ((x >> 31) << 7);
                const float cen = *reinterpret_cast<const float *>(reinterpret_cast<const char *>(sLutCen) + off);
#ifdef S2MV_AD_LUT
                const float ad = sLutAd[sad];
#else
                const float ad = ad_term(sad, inv_ad);
#endif
                const float c = __fadd_rn(ad, cen);
                r[j] = (FULL_D || dv[j]) ? c : 0.0f;
            }
            C4[(size_t)(p0 + i) * LP + q] = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
}

// The four window sums of one output group.  se[i] = start | end << 16 in tile
// positions.  `base` already points at this lane's float4 of position 0.
template <int LP>
__device__ __forceinline__ void sum_group4(const float4 *__restrict__ base, const uint32_t se[4], float4 acc[4])
{
    int s[4], e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s[i] = (int)(se[i] & 0xffffu);
        e[i] = (int)(se[i] >> 16);
        acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int lo = min(min(s[0], s[1]), min(s[2], s[3])), cs = max(max(s[0], s[1]), max(s[2], s[3]));
    const int ce = min(min(e[0], e[1]), min(e[2], e[3])), hi = max(max(e[0], e[1]), max(e[2], e[3]));
    const float4 *p = base + (size_t)lo * LP;
    int k = lo;
    if (cs < ce) {
#pragma unroll 1
        for (; k < cs; ++k, p += LP) {  // head: every window still ends later, some have not started
            const float4 v = *p;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (k >= s[i]) acc4(acc[i], v);
        }
#pragma unroll 2
        for (; k < ce; ++k, p += LP) {  // core: all four windows open
            const float4 v = *p;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc4(acc[i], v);
        }
#pragma unroll 1
        for (; k < hi; ++k, p += LP) {  // tail: every window has started, some have ended
            const float4 v = *p;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (k < e[i]) acc4(acc[i], v);
        }
    } else {
#pragma unroll 1
        for (; k < hi; ++k, p += LP) {  // no common core (very short windows)
            const float4 v = *p;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (k >= s[i] && k < e[i]) acc4(acc[i], v);
        }
    }
}

template <int MODE, int LP>
__global__ void __launch_bounds__(kLineThreads, (LP >= 16 ? 3 : 2))
k_line(const LineArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool VERT = (MODE == LM_V);
    constexpr int Dc = 4 * LP;
    constexpr int TEAMS = kLineThreads / LP;
    const int tid = threadIdx.x, team = tid / LP, q = tid % LP;
    const int ln = VERT ? blockIdx.x : blockIdx.y;
    const int seg = VERT ? blockIdx.y : blockIdx.x;
    const int vslot = blockIdx.z / a.nchunks, chunk = blockIdx.z % a.nchunks;
    const int W = a.W;
    const int LEN = VERT ? a.H : W;
    const int t0 = seg * a.S;
    const int Sact = min(a.S, LEN - t0);
    const int halo = a.halo;
    const int P = a.S + 2 * halo, P4 = (P + 3) & ~3;
    const int d0 = chunk * Dc;

    float4 *C4 = reinterpret_cast<float4 *>(smem_raw);
    uint32_t *sSE = reinterpret_cast<uint32_t *>(C4 + (size_t)P4 * LP);

    // windows of this segment's outputs, in tile positions: [o + halo - A, o + halo + B)
    {
        const uint32_t *__restrict__ arms = a.arms[vslot];
        for (int o = tid; o < a.S; o += kLineThreads) {
            uint32_t se = 0;
            if (o < Sact) {
                const int t = t0 + o;
                const uint32_t ar = __ldg(arms + (VERT ? (size_t)t * W + ln : (size_t)ln * W + t));
                const int A = VERT ? arm_up(ar) : arm_left(ar), B = VERT ? arm_down(ar) : arm_right(ar);
                se = (uint32_t)(o + halo - A) | ((uint32_t)(o + halo + B) << 16);
            }
            sSE[o] = se;
        }
    }

    // position p of the tile <-> line coordinate t0 - halo + p
    const size_t pos_stride4 = VERT ? (size_t)W * a.LPtot : (size_t)a.LPtot;
    const size_t line_base4 = (VERT ? (size_t)ln * a.LPtot : (size_t)ln * W * a.LPtot) + (size_t)chunk * LP + q;

    if (MODE == LM_CI_H) {
        const int view = a.view_first + vslot;
        uint32_t *sOwnP = sSE + a.S, *sOwnC = sOwnP + P4;
        const int NO = P4 + Dc;
        uint32_t *sOthP = sOwnC + P4, *sOthC = sOthP + NO;
        float *sLutCen = reinterpret_cast<float *>(sOthC + NO);
        float *sLutAd = sLutCen + 68;
        const size_t row = (size_t)ln * W;
        const int xb = t0 - halo;
        // other-view words start at the column that makes every thread's 8-word window 16-byte aligned
        const int xo = (view == 0) ? (xb - a.zd + d0) : (xb + a.zd - d0 - Dc);
        const uint32_t *gOwnP = (view == 0 ? a.pixL : a.pixR) + row, *gOwnC = (view == 0 ? a.cenL : a.cenR) + row;
        const uint32_t *gOthP = (view == 0 ? a.pixR : a.pixL) + row, *gOthC = (view == 0 ? a.cenR : a.cenL) + row;
        for (int i = tid; i < P4; i += kLineThreads) {
            const int x = clampi(xb + i, 0, W - 1);
            sOwnP[i] = __ldg(gOwnP + x);
            sOwnC[i] = __ldg(gOwnC + x);
        }
        for (int i = tid; i < NO; i += kLineThreads) {
            const int x = clampi(xo + i, 0, W - 1);
            sOthP[i] = __ldg(gOthP + x);
            sOthC[i] = __ldg(gOthC + x);
        }
        if (tid < kCenLutSize) sLutCen[tid] = a.lutCen[tid];
#ifdef S2MV_AD_LUT
        for (int i = tid; i < kAdLutSize; i += kLineThreads) sLutAd[i] = a.lutAd[i];
#endif
        __syncthreads();
        const bool full_d = d0 + Dc <= a.D;  // no padded disparities in this chunk
#define S2MV_CI_FILL(PLUS, FULL) \
    ci_fill_tile<LP, PLUS, FULL>(C4, sOwnP, sOwnC, sOthP, sOthC, sLutAd, sLutCen, a.inv_ad, P4 / 4, team, q, d0 + 4 * q, a.D)
        if (view == 0) { if (full_d) S2MV_CI_FILL(true, true); else S2MV_CI_FILL(true, false); }
        else           { if (full_d) S2MV_CI_FILL(false, true); else S2MV_CI_FILL(false, false); }
#undef S2MV_CI_FILL
        __syncthreads();
        // the reference's 160-wide blocks read one slot outside their half at tx = 0 / 159.
        // Replay its flat indexing for those columns (they come in pairs 160m-1, 160m; at most three per
        // tile); every thread takes d = tid, tid + 256, ...
        {
            float *C = reinterpret_cast<float *>(smem_raw);
            const int xlo = max(xb, 0), xhi = min(xb + P, W);
            for (int e = (xb <= 0) ? 0 : ((xb + kRefBlockW - 1) / kRefBlockW) * kRefBlockW; e - 1 < xhi; e += kRefBlockW) {
                for (int gx = e - 1; gx <= e; ++gx) {
                    if (gx < xlo || gx >= xhi) continue;
                    const int p = gx - xb;
                    for (int dl = tid; dl < Dc; dl += kLineThreads) {
                        const int d = d0 + dl;
                        float c = 0.0f;
                        if (d < a.D) {
                            const CiOperands o = ref_ci_operands(view, gx, d, a.D, a.zd, W, a.pixL + row, a.pixR + row,
                                                                 a.cenL + row, a.cenR + row);
                            const int sad = (int)__vsadu4(o.ad_own, o.ad_other);
                            const int ham = ref_hamdist32(o.cen_own, o.cen_other);
#ifdef S2MV_AD_LUT
                            c = __fadd_rn(sLutAd[sad], sLutCen[ham]);
#else
                            c = __fadd_rn(ad_term(sad, a.inv_ad), sLutCen[ham]);
#endif
                        }
                        C[(size_t)p * Dc + dl] = c;
                    }
                }
            }
        }
    } else {
        // tile <- volume, 16 B per cp.async; positions outside the line are never inside a window
        const float4 *__restrict__ src = a.in[vslot] + line_base4;
        for (int p = team; p < P; p += TEAMS) {
            const int t = t0 - halo + p;
            if (t >= 0 && t < LEN) cp_async16(&C4[(size_t)p * LP + q], src + (size_t)t * pos_stride4);
        }
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();

    const int ngroups = (Sact + 3) >> 2;
    const int niter = (ngroups + TEAMS - 1) / TEAMS;
    const float4 *base = C4 + q;
    for (int it = 0; it < niter; ++it) {
        const int g = it * TEAMS + team;
        const bool active = g < ngroups;
        uint32_t se[4] = {0u, 0u, 0u, 0u};
        if (active) {
            const uint4 w = *reinterpret_cast<const uint4 *>(sSE + 4 * g);
            se[0] = w.x; se[1] = w.y; se[2] = w.z; se[3] = w.w;
#pragma unroll
            for (int i = 1; i < 4; ++i)
                if (4 * g + i >= Sact) se[i] = se[0];  // past the line end: shadow output 0, never stored
        }
        float4 acc[4];
        sum_group4<LP>(base, se, acc);

        if (MODE != LM_H_WTA) {
            if (active) {
                float4 *__restrict__ dst = a.out[vslot] + line_base4 + (size_t)(t0 + 4 * g) * pos_stride4;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * g + i < Sact) dst[(size_t)i * pos_stride4] = acc[i];
            }
        } else {
            // dc_wta_kernel (d_dc_wta.cu:9-35): strict '>' from FLT_MAX, first minimum wins.
            // Aggregated ADCensus costs are >= 0, so their bit patterns order like the floats.
            const int dq = d0 + 4 * q;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float best = FLT_MAX;
                int bestd = 0x7fffffff;
                if (active) {
                    if (dq + 0 < a.D && best > acc[i].x) { best = acc[i].x; bestd = dq; }
                    if (dq + 1 < a.D && best > acc[i].y) { best = acc[i].y; bestd = dq + 1; }
                    if (dq + 2 < a.D && best > acc[i].z) { best = acc[i].z; bestd = dq + 2; }
                    if (dq + 3 < a.D && best > acc[i].w) { best = acc[i].w; bestd = dq + 3; }
                }
                if (LP == 32) {
                    // one warp = one pixel: CREDUX.MIN over the cost bits, lowest lane holding it has the lowest d
                    const uint32_t bits = bestd == 0x7fffffff ? 0xffffffffu : __float_as_uint(best);
                    const uint32_t m = __reduce_min_sync(0xffffffffu, bits);
                    const uint32_t who = __ballot_sync(0xffffffffu, bits == m);
                    bestd = __shfl_sync(0xffffffffu, bestd, __ffs(who) - 1);
                    best = __uint_as_float(m);
                } else {
#pragma unroll
                    for (int off = LP >> 1; off > 0; off >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                        const int od = __shfl_xor_sync(0xffffffffu, bestd, off);
                        if (ov < best || (ov == best && od < bestd)) { best = ov; bestd = od; }
                    }
                }
                if (active && q == 0 && 4 * g + i < Sact) {
                    if (bestd == 0x7fffffff) bestd = 0;
                    const size_t pix = (size_t)ln * W + (t0 + 4 * g + i);
                    if (a.nchunks == 1) {
                        a.disp[vslot][pix] = (float)bestd - (float)a.zd;
                    } else {
                        const unsigned long long key =
                            ((unsigned long long)float_orderable(best) << 32) | (uint32_t)bestd;
                        atomicMin(a.wta_key[vslot] + pix, key);
                    }
                }
            }
        }
    }
}

}  // namespace s2mv
