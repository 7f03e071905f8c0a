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
    int ln_first;  // first line of this launch (column strips of the vertical passes)
    int d_first;   // disparity of chunk 0 of this launch (chunk-sequential mode: one 128-disparity chunk per launch)
    int use_keys;  // WTA through 64-bit (cost, d) atomicMin keys: the disparity range spans several chunks/launches
    // vertical passes of a row band: outputs are rows [v_begin, v_end), readable input rows [v_lo, v_hi)
    // (whole image: 0, H, 0, H).  Row indices stay image rows; the volume pointers of a band are biased so
    // that row v_lo is the first allocated one.
    int v_begin, v_end, v_lo, v_hi;
    // Row-band neighbours (peer-to-peer halo, fused into the producing pass): outputs on local rows
    // < peer_lo_end are ALSO stored into the upper neighbour's volume, rows >= peer_hi_begin into the lower
    // one's.  peer_out[side][view slot] is the neighbour's buffer (peer-mapped over NVLink, or another context
    // of this process) biased so that THIS band's local row index addresses the neighbour's halo row.
    // Null / 0 / INT_MAX: no neighbour on that side, or not a pass whose output is exchanged.
    float4 *peer_out[2][2];
    int peer_lo_end, peer_hi_begin;
};

// shared-memory bytes of one CTA
__host__ __device__ inline size_t line_smem_bytes(int S, int halo, int LP, bool ci)
{
    const size_t P4 = ((size_t)S + 2 * halo + 3) & ~(size_t)3;
    size_t b = P4 * LP * 16 + (size_t)S * 8;
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
//
// Every accumulator's additions are one contiguous ascending run, so a group splits into
//   private head [s_i, cs)  |  shared core [cs, ce)  |  private tail [ce, e_i)      cs = max s, ce = min e
// and none of the three needs a per-position predicate.  In the common case the four starts and
// the four ends are non-decreasing (neighbouring windows slide; 84-85 % of the groups of the bench
// frame): the heads then form a staircase -- [s0,s1) feeds output 0, [s1,s2) outputs 0-1, [s2,s3)
// outputs 0-2 -- and the tails its mirror image, so each tile position is still read exactly once
// (one LDS.128 per position of the union) and every issued add is a useful one.  Other groups run
// their heads and tails output by output; groups without a common core run four plain loops.
// (The first formulation predicated 16 adds per position on four compares and spent 3/4 of its issue
// slots on compares, loop control and masked-off adds: profiles/r1b_line_kernels_ncu_full.txt.)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}

// acc[i] += tile[p] for every output i in MASK, p = from, from + PB, ... < to (shared-memory byte addresses)
template <int PB, int MASK>
__device__ __forceinline__ void run_add(uint32_t &p, const uint32_t pend, float4 acc[4])
{
#pragma unroll 1
    for (; p != pend; p += PB) {
        const float4 v = lds128<0>(p);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (MASK & (1 << i)) acc4(acc[i], v);
    }
}

// s, e: window start / end of the four outputs as BYTE offsets into the tile (position * PB);
// tq: shared-memory address of this lane's float4 of position 0.
template <int LP>
__device__ __forceinline__ void sum_group4(const uint32_t tq, const uint4 ws, const uint4 we, float4 acc[4])
{
    constexpr int PB = LP * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t cs = max(max(ws.x, ws.y), max(ws.z, ws.w));
    const uint32_t ce = min(min(we.x, we.y), min(we.z, we.w));
    uint32_t p;
    if (cs < ce) {
        const uint32_t pcs = tq + cs, pce = tq + ce;
        const bool mono = ws.x <= ws.y && ws.y <= ws.z && ws.z <= ws.w && we.x <= we.y && we.y <= we.z && we.z <= we.w;
        if (mono) {
            p = tq + ws.x;
            run_add<PB, 0x1>(p, tq + ws.y, acc);
            run_add<PB, 0x3>(p, tq + ws.z, acc);
            run_add<PB, 0x7>(p, pcs, acc);
        } else {
            p = tq + ws.x; run_add<PB, 0x1>(p, pcs, acc);
            p = tq + ws.y; run_add<PB, 0x2>(p, pcs, acc);
            p = tq + ws.z; run_add<PB, 0x4>(p, pcs, acc);
            p = tq + ws.w; run_add<PB, 0x8>(p, pcs, acc);
            p = pcs;
        }
        // core: all four windows open; four positions in flight per trip
        {
            const uint32_t n4 = ((ce - cs) / (4 * PB)) * (4 * PB);
            const uint32_t p4 = pcs + n4;
#pragma unroll 1
            for (; p != p4; p += 4 * PB) {
                const float4 v0 = lds128<0>(p), v1 = lds128<PB>(p), v2 = lds128<2 * PB>(p), v3 = lds128<3 * PB>(p);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc4(acc[i], v0);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc4(acc[i], v1);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc4(acc[i], v2);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc4(acc[i], v3);
            }
            run_add<PB, 0xf>(p, pce, acc);
        }
        if (mono) {
            run_add<PB, 0xe>(p, tq + we.y, acc);
            run_add<PB, 0xc>(p, tq + we.z, acc);
            run_add<PB, 0x8>(p, tq + we.w, acc);
        } else {
            p = pce; run_add<PB, 0x1>(p, tq + we.x, acc);
            p = pce; run_add<PB, 0x2>(p, tq + we.y, acc);
            p = pce; run_add<PB, 0x4>(p, tq + we.z, acc);
            p = pce; run_add<PB, 0x8>(p, tq + we.w, acc);
        }
    } else {
        // no common core (short, scattered windows): four plain runs
        p = tq + ws.x; run_add<PB, 0x1>(p, tq + we.x, acc);
        p = tq + ws.y; run_add<PB, 0x2>(p, tq + we.y, acc);
        p = tq + ws.z; run_add<PB, 0x4>(p, tq + we.z, acc);
        p = tq + ws.w; run_add<PB, 0x8>(p, tq + we.w, acc);
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
    // segments of one line are consecutive CTAs in both orientations: the halo a segment shares with its
    // neighbour is re-read while it is still in L2
    const int ln = a.ln_first + blockIdx.y;
    const int seg = blockIdx.x;
    const int vslot = blockIdx.z / a.nchunks, chunk = blockIdx.z % a.nchunks;
    const int W = a.W;
    const int LEN = VERT ? a.v_end : W;            // end of the outputs along the line
    const int t0 = (VERT ? a.v_begin : 0) + seg * a.S;
    const int Sact = min(a.S, LEN - t0);
    const int IN_LO = VERT ? a.v_lo : 0, IN_HI = VERT ? a.v_hi : W;  // readable input positions
    const int halo = a.halo;
    const int P = a.S + 2 * halo, P4 = (P + 3) & ~3;
    const int d0 = a.d_first + chunk * Dc;

    float4 *C4 = reinterpret_cast<float4 *>(smem_raw);
    uint32_t *sS = reinterpret_cast<uint32_t *>(C4 + (size_t)P4 * LP);
    uint32_t *sE = sS + a.S;
    constexpr uint32_t PB = LP * 16;  // bytes per tile position

    // windows of this segment's outputs as byte offsets into the tile: [o + halo - A, o + halo + B) * PB.
    // Slots past the line end inside the last group of four shadow that group's first output (summed,
    // never stored), so the sum loops need no special case.
    {
        const uint32_t *__restrict__ arms = a.arms[vslot];
        for (int o = tid; o < a.S; o += kLineThreads) {
            const int oe = o < Sact ? o : (o & ~3);
            uint32_t ws = 0, we = 0;
            if (oe < Sact) {
                const int t = t0 + oe;
                const uint32_t ar = __ldg(arms + (VERT ? (size_t)t * W + ln : (size_t)ln * W + t));
                const int A = VERT ? arm_up(ar) : arm_left(ar), B = VERT ? arm_down(ar) : arm_right(ar);
                ws = (uint32_t)(oe + halo - A) * PB;
                we = (uint32_t)(oe + halo + B) * PB;
            }
            sS[o] = ws;
            sE[o] = we;
        }
    }

    // position p of the tile <-> line coordinate t0 - halo + p
    const size_t pos_stride4 = VERT ? (size_t)W * a.LPtot : (size_t)a.LPtot;
    const size_t line_base4 = (VERT ? (size_t)ln * a.LPtot : (size_t)ln * W * a.LPtot) + (size_t)chunk * LP + q;

    if (MODE == LM_CI_H) {
        const int view = a.view_first + vslot;
        uint32_t *sOwnP = sE + a.S, *sOwnC = sOwnP + P4;
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
        const int p_lo = max(0, IN_LO + halo - t0), p_hi = min(P, IN_HI - t0 + halo);
        const long long step = (long long)TEAMS * (long long)pos_stride4 * 16;
        const char *srcp = reinterpret_cast<const char *>(a.in[vslot] + line_base4) +
                           (long long)(t0 - halo + team) * (long long)pos_stride4 * 16;
        uint32_t sdst = smem_u32(C4) + (uint32_t)(team * LP + q) * 16u;
        for (int p = team; p < P; p += TEAMS, srcp += step, sdst += TEAMS * PB)
            if ((unsigned)(p - p_lo) < (unsigned)(p_hi - p_lo))
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(srcp) : "memory");
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();

    const int ngroups = (Sact + 3) >> 2;
    const bool full_d = d0 + Dc <= a.D;  // no padded disparities in this chunk
    const int niter = (ngroups + TEAMS - 1) / TEAMS;
    const uint32_t tq = smem_u32(C4) + (uint32_t)q * 16u;
    const uint32_t *pS = sS + 4 * team;
    const long long ostride = (long long)pos_stride4 * 16;  // bytes between consecutive outputs of the line
    char *dstp = nullptr;
    if (MODE != LM_H_WTA)
        dstp = reinterpret_cast<char *>(a.out[vslot] + line_base4) + (long long)(t0 + 4 * team) * ostride;
    const bool has_peer = MODE != LM_H_WTA && (a.peer_out[0][vslot] != nullptr || a.peer_out[1][vslot] != nullptr);
    for (int it = 0, g = team; it < niter; ++it, g += TEAMS, pS += 4 * TEAMS, dstp += 4 * TEAMS * ostride) {
        const bool active = g < ngroups;
        uint4 ws = make_uint4(0u, 0u, 0u, 0u), we = ws;
        if (active) {
            ws = *reinterpret_cast<const uint4 *>(pS);
            we = *reinterpret_cast<const uint4 *>(pS + a.S);
        }
        float4 acc[4];
        sum_group4<LP>(tq, ws, we, acc);

        if (MODE != LM_H_WTA) {
            if (active) {
                const int rem = Sact - 4 * g;
                if (rem >= 4) {
                    *reinterpret_cast<float4 *>(dstp) = acc[0];
                    *reinterpret_cast<float4 *>(dstp + ostride) = acc[1];
                    *reinterpret_cast<float4 *>(dstp + 2 * ostride) = acc[2];
                    *reinterpret_cast<float4 *>(dstp + 3 * ostride) = acc[3];
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < rem) *reinterpret_cast<float4 *>(dstp + i * ostride) = acc[i];
                }
                if (has_peer) {  // block-uniform: this launch feeds a neighbouring band's halo rows
                    // (everything about the neighbours is fetched here, in the rare path, not kept in registers)
                    const bool peer_up = a.peer_out[0][vslot] != nullptr, peer_dn = a.peer_out[1][vslot] != nullptr;
                    // byte distance from this band's own output to the same element in the neighbour's buffer
                    const long long d_up = reinterpret_cast<char *>(a.peer_out[0][vslot]) - reinterpret_cast<char *>(a.out[vslot]);
                    const long long d_dn = reinterpret_cast<char *>(a.peer_out[1][vslot]) - reinterpret_cast<char *>(a.out[vslot]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = VERT ? (t0 + 4 * g + i) : ln;  // local image row of output i
                        if (i < rem && r < a.peer_lo_end && peer_up) *reinterpret_cast<float4 *>(dstp + i * ostride + d_up) = acc[i];
                        if (i < rem && r >= a.peer_hi_begin && peer_dn) *reinterpret_cast<float4 *>(dstp + i * ostride + d_dn) = acc[i];
                    }
                }
            }
        } else {
            // dc_wta_kernel (d_dc_wta.cu:9-35): strict '>' from FLT_MAX, first minimum wins.
            // Aggregated ADCensus costs are >= +0, so their bit patterns order like the floats.
            const int dq = d0 + 4 * q;
            if (LP == 32 && full_d) {
                // One warp = one pixel, no padded disparities.  Lane minimum (two integer min instructions),
                // CREDUX.MIN across the warp; the lowest lane holding the minimum owns the lowest winning
                // disparity, finds which of its four it was and stores: no shuffles, no per-disparity selects.
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t bx = __float_as_uint(acc[i].x), by = __float_as_uint(acc[i].y),
                                   bz = __float_as_uint(acc[i].z), bw = __float_as_uint(acc[i].w);
                    const uint32_t lm = min(min(bx, by), min(bz, bw));
                    const uint32_t m = __reduce_min_sync(0xffffffffu, lm);
                    const uint32_t who = __ballot_sync(0xffffffffu, lm == m);
                    if (active && q == __ffs(who) - 1 && 4 * g + i < Sact) {
                        const int j = bx == m ? 0 : (by == m ? 1 : (bz == m ? 2 : 3));
                        const size_t pix = (size_t)ln * W + (t0 + 4 * g + i);
                        if (!a.use_keys) {
                            a.disp[vslot][pix] = (float)(dq + j) - (float)a.zd;
                        } else {
                            const unsigned long long key =
                                ((unsigned long long)float_orderable(__uint_as_float(m)) << 32) | (uint32_t)(dq + j);
                            atomicMin(a.wta_key[vslot] + pix, key);
                        }
                    }
                }
            } else {
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
                    if (!a.use_keys) {
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
}

// Neighbour hand-shake of the row-band mode, on the stream, no host involvement: after a pass whose output
// feeds the neighbours' halos, k_band_signal publishes everything this stream has written so far and bumps
// an epoch word in the neighbour's memory; before the pass that reads the halos, k_band_wait spins on this
// band's own epoch word until the neighbour has got that far.  (Bounded: a neighbour that never arrives sets
// *status instead of hanging the GPU.)
__global__ void k_band_signal(unsigned int *peer_flag, unsigned int epoch)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"(epoch) : "memory");
}
__global__ void k_band_wait(const unsigned int *flag, unsigned int epoch, unsigned int *status, long long max_spins)
{
    unsigned int v = 0;
    for (long long spin = 0; spin < max_spins; ++spin) {  // default ~20 s at 1 us per probe: a late neighbour, not a lost one
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - epoch) >= 0) return;
        __nanosleep(1000);
    }
    *status = 1u;  // mapped host memory: every later s2mv_band_* call on this band fails
    __threadfence_system();
}

}  // namespace s2mv
