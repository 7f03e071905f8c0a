// kernels_so.cuh — four-direction scanline optimisation (the stage the reference names dc_hslo).
//
// PARITY UNPINNED: the reference's d_dc_hslo.cu is an unfinished stub and its only call site is
// commented out (image_io.cpp:307-316), so there is no reference output.  The stage is specified in
// DESIGN.md §3.4 (Mei et al. 2011, with the stub's call shape, constants, colour measure and penalty
// tiers; the CPU statement of it lives with the tests) and these kernels match that specification bit
// for bit: fp32, one rounding per operation, in the order written there.
//
//   C_r(p, d) = C(p, d) + min(C_r(p-r, d), C_r(p-r, d-1) + P1, C_r(p-r, d+1) + P1, m + P2) - m
//
// A scanline is a strictly sequential recurrence over its pixels, parallel over disparities and
// over scanlines: one warp walks one scanline, lane l holds disparities 4l..4l+3 of the running
// C_r(p-r, .) in registers; the d-1 / d+1 neighbours across lanes come from two shuffles, m from
// one CREDUX.MIN over the cost bit patterns (all costs are >= +0, so the bits order like the
// floats; disparities beyond num_disp hold +inf and never win).  The volume is disparity-innermost,
// so every step's load is one coalesced run whatever the direction.  The next step's cost and
// pixels are fetched before the current step's recurrence is evaluated.
// The four directions run as four launches accumulating ((lr + rl) + tb) + bt in that order; the
// last one scales by 0.25 and reduces straight to the winner-takes-all disparity.
#pragma once
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace s2mv {

struct SoArgs {
    const float4 *cost[2];      // aggregated volumes, [y][x][LPtot] float4, per view slot
    float4 *acc[2];             // running sum of the directions (same layout)
    float *disp[2];             // WTA output of the last direction (may be null)
    const uint32_t *pix[2];     // packed BGRx of the left / right image
    int H, W, D, zd, LPtot;
    int view_first;             // view of slot 0 (0 = left)
    int dx, dy;                 // scan direction r
    int first, last;            // first: acc = C_r;  otherwise acc += C_r;  last: acc *= 0.25, WTA
    int store_cost;             // last direction: also store the scaled cost volume
    float T, P1[3], P2[3];
};

// mean of B, G, R as the stub computes it for the right image (d_dc_hslo.cu:67-68): float division
__device__ __forceinline__ float so_gray(uint32_t bgrx) { return __fdiv_rn((float)(int)__vsadu4(bgrx & 0x00ffffffu, 0u), 3.0f); }

__global__ void __launch_bounds__(128)
k_so_dir(const SoArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int vslot = blockIdx.y, view = a.view_first + vslot;
    const int horizontal = a.dx != 0;
    const int nlines = horizontal ? a.H : a.W, len = horizontal ? a.W : a.H;
    const int ln = blockIdx.x * 4 + warp;
    if (ln >= nlines) return;
    const int W = a.W, LPtot = a.LPtot;
    const bool have = lane < LPtot;         // this lane holds real (possibly padded) disparities
    const int dbase = 4 * lane;
    const uint32_t *__restrict__ own = a.pix[view];
    const uint32_t *__restrict__ oth = a.pix[1 - view];
    const float4 *__restrict__ cost = a.cost[vslot];
    float4 *__restrict__ acc = a.acc[vslot];
    const float INF = __int_as_float(0x7f800000);
    const int forward = (a.dx > 0 || a.dy > 0);
    // signed column shift of disparity d: +(d - zd) for the left view, -(d - zd) for the right
    int sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sh[j] = view == 0 ? (dbase + j - a.zd) : -(dbase + j - a.zd);

    auto coords = [&](int t, int &x, int &y) {
        const int k = forward ? t : len - 1 - t;
        x = horizontal ? k : ln;
        y = horizontal ? ln : k;
    };
    auto fetch = [&](int x, int y, float4 &c, float &g_own, float g_oth[4]) {
        const size_t pixi = (size_t)y * W + x;
        c = have ? __ldg(cost + pixi * LPtot + lane) : make_float4(INF, INF, INF, INF);
        g_own = so_gray(__ldg(own + pixi));
#pragma unroll
        for (int j = 0; j < 4; ++j) g_oth[j] = so_gray(__ldg(oth + (size_t)y * W + clampi(x + sh[j], 0, W - 1)));
    };

    float4 prev = make_float4(INF, INF, INF, INF);
    float gp_own = 0.f, gp_oth[4] = {0.f, 0.f, 0.f, 0.f};
    int x, y;
    coords(0, x, y);
    float4 c_n;
    float g_own_n, g_oth_n[4];
    fetch(x, y, c_n, g_own_n, g_oth_n);
    for (int t = 0; t < len; ++t) {
        const float4 c = c_n;
        const float g_own = g_own_n;
        float g_oth[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) g_oth[j] = g_oth_n[j];
        const int cx = x, cy = y;
        if (t + 1 < len) {              // next step's operands, before this step's dependent chain
            coords(t + 1, x, y);
            fetch(x, y, c_n, g_own_n, g_oth_n);
        }
        float cur[4] = {c.x, c.y, c.z, c.w};
        if (t > 0) {
            const float pv[4] = {prev.x, prev.y, prev.z, prev.w};
            const float lo = __shfl_up_sync(0xffffffffu, prev.w, 1);     // C_r(p-r, 4l - 1)
            const float hi = __shfl_down_sync(0xffffffffu, prev.x, 1);   // C_r(p-r, 4l + 4)
            const float below[4] = {lane == 0 ? INF : lo, pv[0], pv[1], pv[2]};
            const float above[4] = {pv[1], pv[2], pv[3], lane == 31 ? INF : hi};
            const float lm = fminf(fminf(pv[0], pv[1]), fminf(pv[2], pv[3]));
            const float m = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(lm)));
            const float D1 = fabsf(__fsub_rn(g_own, gp_own));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float D2 = fabsf(__fsub_rn(g_oth[j], gp_oth[j]));
                int tier;
                if (D1 < a.T && D2 < a.T) tier = 0;
                else if ((D1 < a.T && D2 > a.T) || (D1 > a.T && D2 < a.T)) tier = 1;
                else tier = 2;
                const float P1 = a.P1[tier], P2 = a.P2[tier];
                float best = pv[j];
                // the specification skips d-1 at d = 0 and d+1 at d = D-1: those neighbours hold +inf here
                const bool has_above = dbase + j < a.D - 1;
                float v = __fadd_rn(below[j], P1);
                best = v < best ? v : best;
                v = has_above ? __fadd_rn(above[j], P1) : INF;
                best = v < best ? v : best;
                v = __fadd_rn(m, P2);
                best = v < best ? v : best;
                cur[j] = __fsub_rn(__fadd_rn(cur[j], best), m);
            }
        }
        // disparities beyond num_disp never take part
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (dbase + j >= a.D) cur[j] = INF;
        prev = make_float4(cur[0], cur[1], cur[2], cur[3]);
        gp_own = g_own;
#pragma unroll
        for (int j = 0; j < 4; ++j) gp_oth[j] = g_oth[j];

        const size_t pixi = (size_t)cy * W + cx;
        float4 s = prev;
        if (!a.first && have) {
            const float4 o = acc[pixi * LPtot + lane];
            s = make_float4(__fadd_rn(o.x, s.x), __fadd_rn(o.y, s.y), __fadd_rn(o.z, s.z), __fadd_rn(o.w, s.w));
        }
        if (a.last) {
            s = make_float4(__fmul_rn(s.x, 0.25f), __fmul_rn(s.y, 0.25f), __fmul_rn(s.z, 0.25f), __fmul_rn(s.w, 0.25f));
            if (a.disp[vslot]) {
                // first minimum over d (dc_wta_kernel, d_dc_wta.cu:19-34); costs >= +0: compare the bit patterns
                const uint32_t b[4] = {__float_as_uint(s.x), __float_as_uint(s.y), __float_as_uint(s.z), __float_as_uint(s.w)};
                const uint32_t lmin = min(min(b[0], b[1]), min(b[2], b[3]));
                const uint32_t mm = __reduce_min_sync(0xffffffffu, lmin);
                const uint32_t who = __ballot_sync(0xffffffffu, lmin == mm);
                if (lane == __ffs(who) - 1) {
                    const int j = b[0] == mm ? 0 : (b[1] == mm ? 1 : (b[2] == mm ? 2 : 3));
                    a.disp[vslot][pixi] = (float)(dbase + j) - (float)a.zd;
                }
            }
        }
        if (have && (!a.last || a.store_cost)) acc[pixi * LPtot + lane] = s;
    }
}

}  // namespace s2mv
