// kernels_refine.cuh — disparity refinement: left/right cross-check with
// disocclusion labelling, iterative region voting, bilateral filter.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace s2mv {

// dr_dcc_kernel + dr_ddc_kernel (d_dr_dcc.cu:35-82).  outliers must be zeroed
// and disoccl set to 1 beforehand (d_io.cu:139-144, d_dr_dcc.cu:103-109).
__global__ void __launch_bounds__(256)
k_dcc(const float *__restrict__ dispL, const float *__restrict__ dispR, uint8_t *__restrict__ outL,
      uint8_t *__restrict__ outR, uint8_t *__restrict__ disL, uint8_t *__restrict__ disR, int H, int W)
{
    int gx = blockIdx.x * blockDim.x + threadIdx.x;
    int gy = blockIdx.y;
    if (gx >= W) return;
    const size_t row = (size_t)gy * W;
    float d = dispL[row + gx];
    int c = clampi(gx + (int)d, 0, W - 1);
    if (fabsf(__fsub_rn(d, dispR[row + c])) > 1.0f) outL[row + gx] = 1;
    disR[row + c] = 0;  // every writer stores 0: order-free
    d = dispR[row + gx];
    c = clampi(gx - (int)d, 0, W - 1);
    if (fabsf(__fsub_rn(d, dispL[row + c])) > 1.0f) outR[row + gx] = 1;
    disL[row + c] = 0;
}

// dr_merge_errors_kernel (d_dr_dcc.cu:18-33)
__global__ void __launch_bounds__(256)
k_dcc_merge(uint8_t *__restrict__ outL, uint8_t *__restrict__ outR, const uint8_t *__restrict__ disL,
            const uint8_t *__restrict__ disR, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (outL[i] == 1 && disL[i] == 1) outL[i] = 2;
    if (outR[i] == 1 && disR[i] == 1) outR[i] = 2;
}

// The same three steps (flags cleared / dis-occlusion marks set, cross-check + marks, merge) for one image ROW per
// block: a pixel's check reads and marks only its own row, so the marks live in shared memory and the whole stage is
// one launch that writes each outlier byte once -- instead of four memsets and two kernels (k_dcc, k_dcc_merge: kept
// for rows that do not fit shared memory).  sm: [2][W] bytes.
__global__ void __launch_bounds__(256)
k_dcc_row(const float *__restrict__ dispL, const float *__restrict__ dispR, uint8_t *__restrict__ outL,
          uint8_t *__restrict__ outR, int H, int W)
{
    extern __shared__ uint8_t dis_sm[];
    uint8_t *disL = dis_sm, *disR = dis_sm + W;
    const size_t row = (size_t)blockIdx.x * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) { disL[x] = 1; disR[x] = 1; }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const float dl = dispL[row + x];
        disR[clampi(x + (int)dl, 0, W - 1)] = 0;  // every writer stores 0: order-free
        const float dr = dispR[row + x];
        disL[clampi(x - (int)dr, 0, W - 1)] = 0;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        float d = dispL[row + x];
        int c = clampi(x + (int)d, 0, W - 1);
        const bool ol = fabsf(__fsub_rn(d, dispR[row + c])) > 1.0f;
        d = dispR[row + x];
        c = clampi(x - (int)d, 0, W - 1);
        const bool orr = fabsf(__fsub_rn(d, dispL[row + c])) > 1.0f;
        outL[row + x] = ol ? (disL[x] ? 2 : 1) : 0;
        outR[row + x] = orr ? (disR[x] ? 2 : 1) : 0;
    }
}

// ---- iterative region voting (d_dr_irv.cu:17-43,134-269) -----------------
// The reference gives every pixel a thread and every outlier thread a private
// 65-int histogram in local memory, five times per view.  Here the outliers
// are compacted ONCE into a list (they only ever shrink: a vote can clear an
// outlier, nothing creates one); each iteration votes on the list — one warp
// per outlier, one lane per row of its cross-shaped support so the dependent
// arms -> disparity loads of different rows overlap, histogram in shared memory
// — then applies the accepted votes and writes the survivors as the next
// iteration's list.  Votes read a snapshot (separate kernels): the race-free
// reading of Q15.
struct IrvArgs {
    float *disp[2];
    uint8_t *outliers[2];
    const uint32_t *arms[2];
    int *list[2];        // outlier pixel indices of this iteration
    int *next[2];        // survivors (k_irv_apply)
    int *vote[2];        // accepted disparity or kNoVote, per list entry
    int *count[2];       // length of list
    int *next_count[2];  // length of next
    int *ticket[2];      // next list entry to hand out (k_irv_vote_dense); reset by k_irv_apply
    int *accepted[2];    // [iteration]: votes k_irv_apply accepted in that iteration (null: not tracked)
    int it;              // iteration of this launch
    int H, W, nbins, zd, usd, thresh_s;
    float thresh_h;
    // dense path (k_irv_hseg + k_irv_vote_dense): per-pixel histograms of the horizontal arm span
    uint8_t *hseg[2];    // [pixel][nbp] counts, nbp = nbins rounded up to a multiple of 128; null: sparse path only
    int nbp;
    // incremental iterations (dense path): stamp[pixel] = 1 + the iteration whose vote changed the pixel (0: never);
    // hchg[pixel] = the pixel's horizontal span holds a pixel changed by the previous iteration
    uint8_t *stamp[2];
    uint8_t *hchg[2];
    // rows voted in this iteration, [row_lo, row_hi): the whole image, or -- in a row band -- the band's own rows plus
    // what the remaining iterations and the filters after them can still carry into the own rows (s2mv_api.cu)
    int row_lo, row_hi;
    int dense_min;       // list length from which an iteration takes the dense path
    // k_irv_sparse_all (all iterations in one cooperative launch): iterations, and the row range as launch_irv derives it
    // per iteration -- rows within post_reach + usd * (iterations - 1 - it) of [keep0, keep1) (keep1 < 0: all rows)
    int iterations, keep0, keep1, post_reach;
    int *hint;           // mapped host memory, [view]: length of the first iteration's list (read by the NEXT frame's host code)
    int col_votes;       // > 0: dense iterations vote column by column (k_irv_vote_col), this many columns (1..4) per
                         // ticket; vote[] is then indexed by PIXEL
};
constexpr int kNoVote = -0x7fffffff;

// An iteration that accepted no vote left disparities and outlier flags as they were, so every later
// iteration would recompute the same rejected votes: its kernels return at once (exact, not a heuristic).
__device__ __forceinline__ bool irv_settled(const IrvArgs &a, int v)
{
    return a.accepted[v] != nullptr && a.it > 0 && a.accepted[v][a.it - 1] == 0;
}

// append `mine` items per lane to a global list with one atomic per warp; returns this lane's base
__device__ __forceinline__ int warp_append_base(int *counter, int mine)
{
    const int lane = threadIdx.x & 31;
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - mine;
}

// 16 pixels per thread
__global__ void __launch_bounds__(256)
k_irv_compact(const IrvArgs a)
{
    const int v = blockIdx.y;
    const size_t n = (size_t)a.H * a.W;
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const uint8_t *__restrict__ outl = a.outliers[v];
    uint32_t nz = 0;  // bit j: pixel i0 + j is an outlier
    if (i0 + 16 <= n) {
        const uint4 w = *reinterpret_cast<const uint4 *>(outl + i0);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // high bit of every non-zero byte
            const uint32_t hb = (((ws[k] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | ws[k]) & 0x80808080u;
            nz |= (((hb >> 7) & 1u) | ((hb >> 14) & 2u) | ((hb >> 21) & 4u) | ((hb >> 28) & 8u)) << (4 * k);
        }
    } else {
        for (int j = 0; j < 16; ++j)
            if (i0 + j < n && outl[i0 + j] != 0) nz |= 1u << j;
    }
    int pos = warp_append_base(a.count[v], __popc(nz));
    int *__restrict__ list = a.list[v];
    while (nz) {
        const int j = __ffs(nz) - 1;
        nz &= nz - 1;
        list[pos++] = (int)(i0 + j);
    }
}

constexpr int kIrvWarps = 8;

// One outlier's vote by the sparse rule (one warp; hist: the warp's nbins counters in shared memory): the body of
// k_irv_vote and of k_irv_sparse_all.  Returns through vote_out (lane 0 writes).
__device__ __forceinline__ void irv_vote_entry(const IrvArgs &a, int v, int pix, int row_lo, int row_hi, int *hist, int lane,
                                               int *vote_out)
{
    const float *__restrict__ disp = a.disp[v];
    const uint8_t *__restrict__ outl = a.outliers[v];
    const uint32_t *__restrict__ arms = a.arms[v];
    const int W = a.W;
    {
        const int gy = pix / W, gx = pix - gy * W;
        if (gy < row_lo || gy >= row_hi) {  // not voted in this iteration (row band: cannot reach the own rows any more)
            if (lane == 0) *vote_out = kNoVote;
            return;
        }
        for (int b = lane; b < a.nbins; b += 32) hist[b] = 0;
        __syncwarp();
        const uint32_t ac = arms[pix];
        const int cu = min(arm_up(ac), a.usd), nrows = cu + arm_down(ac) + 1;  // rows [-cu, +cd] inclusive
        int cnt = 0;
        // the arms of the support rows first (lane r <-> rows r, r + 32, ...; nrows <= 2*usd+1 <= 129), then one
        // row per step with the lanes ACROSS its span: every row is one or two coalesced requests that do not
        // depend on each other, instead of one lane walking a row pixel by pixel
        uint32_t rarm[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int r = lane + 32 * j;
            rarm[j] = r < nrows ? arms[(size_t)(gy - cu + r) * W + gx] : 0u;
        }
#pragma unroll 8
        for (int r = 0; r < nrows; ++r) {
            const int rj = r >> 5;
            const uint32_t pick = rj == 0 ? rarm[0] : (rj == 1 ? rarm[1] : (rj == 2 ? rarm[2] : (rj == 3 ? rarm[3] : rarm[4])));
            const uint32_t ar = __shfl_sync(0xffffffffu, pick, r & 31);
            const int cl = arm_left(ar), span = cl + arm_right(ar) + 1;  // inclusive [-L, R]
            const size_t row = (size_t)(gy - cu + r) * W + (gx - cl);
            for (int k0 = 0; k0 < span; k0 += 32) {
                const int k = k0 + lane;
                // both loads issue together (the disparity is not waited for behind the outlier flag)
                const bool in = k < span;
                const uint8_t o = in ? outl[row + k] : (uint8_t)1;
                const float dv = in ? disp[row + k] : 0.0f;
                const bool valid = o == 0;
                const int bin = clampi((int)dv + a.zd, 0, a.nbins - 1);
                const unsigned act = __ballot_sync(0xffffffffu, valid);
                if (valid) {
                    // neighbours mostly vote for the same bin: one shared-memory add per distinct bin
                    const unsigned same = __match_any_sync(act, bin);
                    if (lane == __ffs(same) - 1) atomicAdd(&hist[bin], __popc(same));
                    ++cnt;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        // first bin holding the maximum count (strict '<' scan from bin 0)
        int best = 0, bestb = 0x7fffffff;
        for (int b = lane; b < a.nbins; b += 32) {
            int h = hist[b];
            if (best < h) { best = h; bestb = b; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            int ob = __shfl_xor_sync(0xffffffffu, best, off);
            int obb = __shfl_xor_sync(0xffffffffu, bestb, off);
            if (ob > best || (ob == best && obb < bestb)) { best = ob; bestb = obb; }
        }
        __syncwarp();
        if (lane == 0) {
            int max_d = (best > 0) ? (bestb - a.zd) : (int)disp[pix];
            // dr_irv_kernel_3: ratio test on the histogram INDEX (Q16)
            bool ok = cnt > a.thresh_s && __fdiv_rn((float)(max_d + a.zd), (float)cnt) > a.thresh_h;
            *vote_out = ok ? max_d : kNoVote;
        }
    }
}

__global__ void __launch_bounds__(kIrvWarps * 32)
k_irv_vote(const IrvArgs a)
{
    extern __shared__ int hist_all[];  // [kIrvWarps][nbins]
    const int v = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *hist = hist_all + warp * a.nbins;
    if (irv_settled(a, v)) return;
    const int count = *a.count[v];
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.next_count[v] = 0;  // consumed by k_irv_apply, which runs after
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.it == 0 && a.hint) a.hint[v] = count;  // for the next frame's host code
    if (a.hseg[v] && count >= a.dense_min) return;                  // this iteration is k_irv_vote_dense's
    for (int e = blockIdx.x * kIrvWarps + warp; e < count; e += gridDim.x * kIrvWarps)
        irv_vote_entry(a, v, a.list[v][e], a.row_lo, a.row_hi, hist, lane, a.vote[v] + e);
}

// ---- dense path ------------------------------------------------------------
// With many outliers (occlusion-heavy frames: half the pixels of the synthetic streams) the per-outlier
// gather above re-reads every support row once per outlier of its column.  The vote histogram of outlier p
// is the sum, over the pixels q of p's vertical arm, of the histogram of q's horizontal span -- and that
// inner histogram depends on q alone.  k_irv_hseg builds it once per pixel (8-bit counts: a span holds at
// most 2*usd+1 <= 129 pixels), k_irv_vote_dense adds <= 2*usd+1 of them per outlier with packed 16-bit
// adds: one coalesced 128-byte load per support row instead of a row of scattered loads and shared-memory
// atomics.  Same histogram, same count, same first-maximum rule as k_irv_vote.
constexpr int kHsegThreads = 128;
#ifndef S2MV_IRV_BATCH
#define S2MV_IRV_BATCH 4
#endif

// From the second iteration on only the pixels whose span holds a pixel the previous iteration changed are
// rebuilt (the vote of iteration t reads the state after iteration t-1; a span without a change has the histogram
// it had): the kernel first looks at the change stamps of the tile and its arm reach and, in the common case of a
// late iteration, leaves the tile alone.  hchg records per pixel whether its span changed, for k_irv_vote_dense.
template <int NW>  // 128-bin words per pixel: nbp = 128 * NW
__global__ void __launch_bounds__(kHsegThreads)
k_irv_hseg(const IrvArgs a)
{
    extern __shared__ __align__(16) uint8_t hs[];  // [kHsegThreads][nbp + 4]: the pad staggers the banks
    __shared__ uint8_t chg[kHsegThreads + 2 * 64];  // "changed by the previous iteration", tile + arm reach (usd <= 64)
    __shared__ uint8_t need[kHsegThreads];
    const int v = blockIdx.y;
    if (irv_settled(a, v) || *a.count[v] < a.dense_min) return;
    constexpr int nbp = 128 * NW, pitch = nbp + 4;
    const int W = a.W;
    const int tiles_x = (W + kHsegThreads - 1) / kHsegThreads, ntiles = tiles_x * a.H;
    const float *__restrict__ disp = a.disp[v];
    const uint8_t *__restrict__ outl = a.outliers[v];
    const uint32_t *__restrict__ arms = a.arms[v];
    const uint8_t *__restrict__ stamp = a.stamp[v];
    uint8_t *__restrict__ hseg = a.hseg[v];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const bool incremental = a.it > 0 && stamp != nullptr;
    uint32_t *hw = reinterpret_cast<uint32_t *>(hs);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int gy = tile / tiles_x, bx = (tile - gy * tiles_x) * kHsegThreads;
        if (gy < a.row_lo - a.usd || gy >= a.row_hi + a.usd) continue;  // no voted outlier reads this row (block-uniform)
        const int gx = bx + t;
        const size_t row = (size_t)gy * W;
        int mine = 1;  // this pixel's histogram has to be (re)built
        if (incremental) {
            int any = 0;
            for (int i = t; i < kHsegThreads + 2 * a.usd; i += kHsegThreads) {
                const int x = bx - a.usd + i;
                const uint8_t c = (x >= 0 && x < W && stamp[row + x] == (uint8_t)a.it) ? 1 : 0;
                chg[i] = c;
                any |= c;
            }
            if (!__syncthreads_or(any)) {  // nothing in reach changed: histograms and flags of the tile stand
                if (gx < W) a.hchg[v][row + gx] = 0;
                continue;
            }
            mine = 0;
            if (gx < W) {
                const uint32_t ar = arms[row + gx];
                const int cl = arm_left(ar), cr = arm_right(ar);
                if (cl > a.usd || cr > a.usd) mine = 1;  // arms from a caller (stage API) beyond usd: always rebuilt
                else for (int k = -cl; k <= cr; ++k) mine |= chg[t + a.usd + k];
                a.hchg[v][row + gx] = (uint8_t)mine;
            }
        }
        need[t] = (uint8_t)(mine && gx < W);
        {
            uint4 *z = reinterpret_cast<uint4 *>(hs);  // kHsegThreads * pitch bytes is a multiple of 16
            for (int i = t; i < kHsegThreads * pitch / 16; i += kHsegThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        if (gx < W && mine) {
            const uint32_t ar = arms[row + gx];
            const int cl = arm_left(ar), span = cl + arm_right(ar) + 1;  // inclusive [-L, R]
            uint8_t *h = hs + t * pitch;
            const float *__restrict__ dp = disp + row + (gx - cl);
            const uint8_t *__restrict__ op = outl + row + (gx - cl);
            for (int k = 0; k < span; ++k)
                if (op[k] == 0) h[clampi((int)dp[k] + a.zd, 0, a.nbins - 1)] += 1;
        }
        __syncthreads();
        // out: one pixel's nbp bytes per warp step, 128 bytes per store instruction, pointers walked by increments
        const int npix = min(kHsegThreads, W - bx);
        constexpr int wpp = nbp / 4, wps = pitch / 4;  // 32-bit words per pixel: global / shared
        uint32_t *__restrict__ d = reinterpret_cast<uint32_t *>(hseg + ((size_t)gy * W + bx) * nbp) + (size_t)warp * wpp + lane;
        const uint32_t *sp = hw + warp * wps + lane;
        for (int i = warp; i < npix; i += kHsegThreads / 32, d += (kHsegThreads / 32) * wpp, sp += (kHsegThreads / 32) * wps) {
            if (!need[i]) continue;
#pragma unroll
            for (int w = 0; w < NW; ++w) d[32 * w] = sp[32 * w];
        }
        __syncthreads();
    }
}

template <int NW>  // 128-bin words per lane: nbp = 128 * NW
__global__ void __launch_bounds__(kIrvWarps * 32)
k_irv_vote_dense(const IrvArgs a)
{
    const int v = blockIdx.y;
    const int lane = threadIdx.x & 31;
    if (irv_settled(a, v)) return;
    const int count = *a.count[v];
    if (count < a.dense_min) return;
    const float *__restrict__ disp = a.disp[v];
    const uint32_t *__restrict__ arms = a.arms[v];
    const uint32_t *__restrict__ hseg = reinterpret_cast<const uint32_t *>(a.hseg[v]);
    const int W = a.W;
    constexpr int WPP = 32 * NW;  // 32-bit words per pixel
    // Entries are handed out in list (= raster) order, a few per ticket (4 measured best): the warps in flight then work on
    // neighbouring image rows, whose span histograms they share through L2 (with a fixed stride per warp the
    // uneven cost per entry lets the warps drift apart)
    constexpr int kBatch = S2MV_IRV_BATCH;
    for (int e = 0, e_end = 0;; ++e) {
        if (e == e_end) {
            if (lane == 0) e = atomicAdd(a.ticket[v], kBatch);
            e = __shfl_sync(0xffffffffu, e, 0);
            e_end = min(e + kBatch, count);
            if (e >= count) break;
        }
        const int pix = a.list[v][e];
        const int gy = pix / W, gx = pix - gy * W;
        if (gy < a.row_lo || gy >= a.row_hi) {  // not voted in this iteration
            if (lane == 0) a.vote[v][e] = kNoVote;
            continue;
        }
        const uint32_t ac = arms[pix];
        const int cu = min(arm_up(ac), a.usd), nrows = cu + arm_down(ac) + 1;  // rows [-cu, +cd] inclusive
        if (a.it > 0 && a.stamp[v] != nullptr) {
            // Still on the list = its last vote was rejected.  If no pixel of its support changed since (no support
            // row's span holds a pixel the previous iteration changed), the same vote would be rejected again.
            int hit = 0;
            for (int r = lane; r < nrows; r += 32) hit |= a.hchg[v][(size_t)(gy - cu + r) * W + gx];
            if (!__any_sync(0xffffffffu, hit)) {
                if (lane == 0) a.vote[v][e] = kNoVote;
                continue;
            }
        }
        uint32_t lo[NW], hi[NW];  // 16-bit fields: bins (0, 2) and (1, 3) of each packed word
#pragma unroll
        for (int w = 0; w < NW; ++w) lo[w] = hi[w] = 0u;
        const uint32_t *__restrict__ p = hseg + ((size_t)(gy - cu) * W + gx) * WPP + lane;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r, p += (size_t)W * WPP) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t x = __ldg(p + 32 * w);
                lo[w] += x & 0x00ff00ffu;
                hi[w] += (x >> 8) & 0x00ff00ffu;
            }
        }
        // total count; first bin holding the maximum count: maximise (count, -bin)
        uint32_t cnt = 0, key = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t c[4] = {lo[w] & 0xffffu, hi[w] & 0xffffu, lo[w] >> 16, hi[w] >> 16};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cnt += c[j];
                const uint32_t bin = (uint32_t)(128 * w + 4 * lane + j);
                const uint32_t k = c[j] ? ((c[j] << 10) | (1023u - bin)) : 0u;
                key = max(key, k);
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        key = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0) {
            const int best = (int)(key >> 10), bestb = 1023 - (int)(key & 1023u);
            int max_d = (best > 0) ? (bestb - a.zd) : (int)disp[pix];
            // dr_irv_kernel_3: ratio test on the histogram INDEX (Q16)
            bool ok = (int)cnt > a.thresh_s && __fdiv_rn((float)(max_d + a.zd), (float)(int)cnt) > a.thresh_h;
            a.vote[v][e] = ok ? max_d : kNoVote;
        }
    }
}

// ---- dense path, column walk -------------------------------------------------
// k_irv_vote_dense adds up to 2*usd+1 span histograms per outlier, and vertically adjacent outliers of a column add
// almost the same ones: their vertical arms end at the same colour edge (or both at usd rows, one row apart).  Here a
// warp owns a 32 x 32 pixel tile and walks each of its columns downwards with the histogram of the current row window
// in registers; the next outlier's window is reached by adding the rows that enter and subtracting the rows that leave
// (packed 16-bit fields: a row that leaves was added before, so no field borrows), or rebuilt when that is cheaper.
// An occluded region then costs about two 128-byte row reads per outlier instead of its arm length; sums of integer
// counts, so the histogram -- and with it count, first maximum and vote -- is the one k_irv_vote_dense forms.
// Votes are stored per PIXEL (k_irv_apply reads them there when col_votes is set).
// 6 blocks per SM (40 registers): the walk is a chain of dependent loads per warp, so warps in flight are what hides
// it -- 48 per SM against 32 took 8 % off the refinement of a config-3 frame; 64 (32 registers, spills) gave it back
template <int NW>
__global__ void __launch_bounds__(kIrvWarps * 32, 6)
k_irv_vote_col(const IrvArgs a)
{
    const int v = blockIdx.y;
    const int lane = threadIdx.x & 31;
    if (irv_settled(a, v)) return;
    const int count = *a.count[v];
    if (count < a.dense_min) return;
    const float *__restrict__ disp = a.disp[v];
    const uint8_t *__restrict__ outl = a.outliers[v];
    const uint32_t *__restrict__ arms = a.arms[v];
    const uint8_t *__restrict__ hchg = a.hchg[v];
    const uint32_t *__restrict__ hseg = reinterpret_cast<const uint32_t *>(a.hseg[v]);
    int *__restrict__ vote = a.vote[v];
    const int W = a.W, H = a.H, usd = a.usd;
    constexpr int WPP = 32 * NW;  // 32-bit words per pixel
    const bool incremental = a.it > 0 && a.stamp[v] != nullptr;
    // a ticket = kColW adjacent columns of a 32-row strip; strips in raster order, so the warps in flight sit on
    // neighbouring columns of the same rows (their flag, arm and histogram reads share sectors and L2 lines)
    const int kColW = a.col_votes;
    const int tiles_x = (W + kColW - 1) / kColW, tiles_y = (a.row_hi - a.row_lo + 31) / 32, ntiles = tiles_x * tiles_y;
    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(a.ticket[v], 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int xs = tx * kColW, ys = a.row_lo + ty * 32;
        const int y_l = ys + lane;                      // lane <-> row of the strip
        const bool row_ok = y_l < a.row_hi;
        const int ncol = min(kColW, W - xs);
        // outlier flags: bit j of `orow` = pixel (y_l, xs + j)
        uint32_t orow = 0;
        if (row_ok) {
            const uint8_t *op = outl + (size_t)y_l * W + xs;
            if (ncol == 4 && ((uintptr_t)op & 3) == 0) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(op);
                const uint32_t hb = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;  // non-zero bytes
                orow = ((hb >> 7) & 1u) | ((hb >> 14) & 2u) | ((hb >> 21) & 4u) | ((hb >> 28) & 8u);
            } else {
                for (int j = 0; j < ncol; ++j) orow |= (op[j] != 0 ? 1u : 0u) << j;
            }
        }
        if (!__any_sync(0xffffffffu, orow != 0)) continue;
        for (int j = 0; j < ncol; ++j) {
            uint32_t m = __ballot_sync(0xffffffffu, (orow >> j) & 1u);  // bit i: (ys + i, gx) is an outlier
            if (m == 0) continue;
            const int gx = xs + j;
            const uint32_t arm_l = row_ok ? arms[(size_t)y_l * W + gx] : 0u;
            // rows of this column whose span changed in the previous iteration, [ys - usd, ys + 32 + usd): bit r of word k
            uint32_t chg[5] = {0u, 0u, 0u, 0u, 0u};
            if (incremental) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int r = ys - usd + 32 * k + lane;
                    const bool in = 32 * k < 32 + 2 * usd && r >= 0 && r < H;
                    chg[k] = __ballot_sync(0xffffffffu, in && hchg[(size_t)r * W + gx] != 0);
                }
            }
            uint32_t lo[NW], hi[NW];  // 16-bit fields: bins (0, 2) and (1, 3) of each packed word
#pragma unroll
            for (int w = 0; w < NW; ++w) lo[w] = hi[w] = 0u;
            int wlo = 0, whi = -1;  // rows in the window, inclusive; empty
            const size_t rstride = (size_t)W * WPP;
            const uint32_t *__restrict__ col = hseg + (size_t)gx * WPP + lane;
            // rows wlo and whi of the column: the window is walked by pointer steps, not by a 64-bit multiply per row
            const uint32_t *__restrict__ p_top = col, *__restrict__ p_bot = col;
            // lane i keeps the count and key of the outlier in row ys + i; the votes of the strip are formed afterwards
            // with the lanes side by side (the conversion, the division and the store once per strip, not per outlier)
            int my_cnt = -1;
            uint32_t my_key = 0;
            const uint32_t m_all = m;
            while (m) {
                const int i = __ffs(m) - 1;
                m &= m - 1;
                const int gy = ys + i;
                const uint32_t ac = __shfl_sync(0xffffffffu, arm_l, i);
                const int nlo = gy - min(arm_up(ac), usd), nhi = gy + arm_down(ac);
                if (incremental) {
                    // still listed = its last vote was rejected; with no changed span in its support it would be again
                    const int b0 = nlo - (ys - usd), b1 = nhi - (ys - usd);  // bit range, inclusive
                    uint32_t hit = b1 >= 32 + 2 * usd ? 1u : 0u;  // (a caller's arm beyond usd: outside the tracked rows)
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int s0 = max(b0 - 32 * k, 0), s1 = min(b1 - 32 * k, 31);
                        if (s0 <= s1) hit |= chg[k] & ((0xffffffffu >> (31 - s1)) & (0xffffffffu << s0));
                    }
                    if (!hit) continue;  // my_cnt of lane i stays -1: no vote
                }
                const int nrows = nhi - nlo + 1;
                const bool slide = whi >= wlo && abs(nlo - wlo) + abs(nhi - whi) < nrows;
                if (!slide) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) lo[w] = hi[w] = 0u;
                    wlo = nlo;
                    whi = nlo - 1;
                    p_top = col + (size_t)nlo * rstride;
                    p_bot = p_top - rstride;
                }
                // rows that leave (above the new top / below the new bottom), then rows that enter
                for (; wlo < nlo; ++wlo, p_top += rstride) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const uint32_t x = __ldg(p_top + 32 * w);
                        lo[w] -= x & 0x00ff00ffu;
                        hi[w] -= (x >> 8) & 0x00ff00ffu;
                    }
                }
                for (; whi > nhi; --whi, p_bot -= rstride) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const uint32_t x = __ldg(p_bot + 32 * w);
                        lo[w] -= x & 0x00ff00ffu;
                        hi[w] -= (x >> 8) & 0x00ff00ffu;
                    }
                }
                for (; wlo > nlo;) {
                    --wlo;
                    p_top -= rstride;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const uint32_t x = __ldg(p_top + 32 * w);
                        lo[w] += x & 0x00ff00ffu;
                        hi[w] += (x >> 8) & 0x00ff00ffu;
                    }
                }
#pragma unroll 4
                for (; whi < nhi;) {
                    ++whi;
                    p_bot += rstride;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const uint32_t x = __ldg(p_bot + 32 * w);
                        lo[w] += x & 0x00ff00ffu;
                        hi[w] += (x >> 8) & 0x00ff00ffu;
                    }
                }
                // total count; first bin holding the maximum count: maximise (count, -bin).  An all-zero histogram gives
                // a key with count 0 (best == 0 below); the four fields of a lane add up without leaving 16 bits
                uint32_t cnt = 0, key = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const uint32_t c[4] = {lo[w] & 0xffffu, hi[w] & 0xffffu, lo[w] >> 16, hi[w] >> 16};
                    const uint32_t s2 = lo[w] + hi[w];
                    cnt += (s2 & 0xffffu) + (s2 >> 16);
                    const uint32_t inv = 1023u - (uint32_t)(128 * w + 4 * lane);
#pragma unroll
                    for (int q = 0; q < 4; ++q) key = max(key, c[q] * 1024u + (inv - q));
                }
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                key = __reduce_max_sync(0xffffffffu, key);
                if (lane == i) {
                    my_cnt = (int)cnt;
                    my_key = key;
                }
            }
            if ((m_all >> lane) & 1u) {  // this lane's row holds an outlier of the column
                const size_t pix = (size_t)y_l * W + gx;
                int out = kNoVote;
                if (my_cnt >= 0) {
                    const int best = (int)(my_key >> 10), bestb = 1023 - (int)(my_key & 1023u);
                    const int max_d = (best > 0) ? (bestb - a.zd) : (int)disp[pix];
                    // dr_irv_kernel_3: ratio test on the histogram INDEX (Q16)
                    const bool ok = my_cnt > a.thresh_s && __fdiv_rn((float)(max_d + a.zd), (float)my_cnt) > a.thresh_h;
                    if (ok) out = max_d;
                }
                vote[pix] = out;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_irv_apply(const IrvArgs a)
{
    const int v = blockIdx.y;
    if (irv_settled(a, v)) return;
    const int count = *a.count[v];
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.ticket[v] = 0;  // for the next iteration's vote
    const int stride = gridDim.x * blockDim.x;
    // this iteration's votes came from k_irv_vote_col: stored per pixel, for the rows it walked
    const bool by_pixel = a.col_votes && a.hseg[v] && count >= a.dense_min;
    int taken = 0;
    for (int e0 = blockIdx.x * blockDim.x; e0 < count; e0 += stride) {  // block-uniform trip count
        const int e = e0 + threadIdx.x;
        int pix = -1;
        if (e < count) {
            pix = a.list[v][e];
            int vote = kNoVote;
            if (!by_pixel) {
                vote = a.vote[v][e];
            } else if (pix >= a.row_lo * a.W && pix < a.row_hi * a.W) {
                vote = a.vote[v][pix];
            }
            if (vote != kNoVote) {
                a.outliers[v][pix] = 0;
                a.disp[v][pix] = (float)vote;
                if (a.stamp[v]) a.stamp[v][pix] = (uint8_t)(a.it + 1);
                pix = -1;
                ++taken;
            }
        }
        const int pos = warp_append_base(a.next_count[v], pix >= 0 ? 1 : 0);
        if (pix >= 0) a.next[v][pos] = pix;
    }
    if (a.accepted[v] != nullptr) {
        taken = __reduce_add_sync(0xffffffffu, taken);
        if ((threadIdx.x & 31) == 0 && taken) atomicAdd(a.accepted[v] + a.it, taken);
    }
}

// ---- every iteration in one launch (light frames) ---------------------------
// A frame with few outliers spends region voting on launches: five iterations x (vote, apply) of a few microseconds
// each, most of them returning at once after the iteration that accepted nothing.  This kernel is the sparse path of
// all iterations in ONE cooperative launch -- vote, grid barrier, apply, grid barrier -- with the same vote per outlier
// (irv_vote_entry), the same snapshot semantics (votes of an iteration read the state the previous one left) and the
// same stop rule.  The host launches it when the PREVIOUS frame's list was short (a.hint, written here and by
// k_irv_vote); it is exact whatever the list length, only slower than the dense path on a long list.
__global__ void __launch_bounds__(kIrvWarps * 32)
k_irv_sparse_all(IrvArgs a)
{
    extern __shared__ int hist_all[];  // [kIrvWarps][nbins]
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int v = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *hist = hist_all + warp * a.nbins;
    int *list = a.list[v], *next = a.next[v], *count_p = a.count[v], *next_count_p = a.next_count[v];
    bool settled = false;
    for (int it = 0; it < a.iterations; ++it) {
        int row_lo = 0, row_hi = a.H;
        if (a.keep1 >= 0) {
            const int m = a.post_reach + a.usd * (a.iterations - 1 - it);
            row_lo = max(0, a.keep0 - m);
            row_hi = min(a.H, a.keep1 + m);
        }
        const int count = settled ? 0 : *count_p;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            *next_count_p = 0;
            if (it == 0 && a.hint) a.hint[v] = count;
        }
        for (int e = blockIdx.x * kIrvWarps + warp; e < count; e += gridDim.x * kIrvWarps)
            irv_vote_entry(a, v, list[e], row_lo, row_hi, hist, lane, a.vote[v] + e);
        grid.sync();
        // apply (k_irv_apply): accepted votes clear the outlier, the others form the next list
        int taken = 0;
        const int stride = gridDim.x * blockDim.x;
        for (int e0 = blockIdx.x * blockDim.x; e0 < count; e0 += stride) {  // block-uniform trip count
            const int e = e0 + threadIdx.x;
            int pix = -1;
            if (e < count) {
                pix = list[e];
                const int vote = a.vote[v][e];
                if (vote != kNoVote) {
                    a.outliers[v][pix] = 0;
                    a.disp[v][pix] = (float)vote;
                    pix = -1;
                    ++taken;
                }
            }
            const int pos = warp_append_base(next_count_p, pix >= 0 ? 1 : 0);
            if (pix >= 0) next[pos] = pix;
        }
        taken = __reduce_add_sync(0xffffffffu, taken);
        if (lane == 0 && taken) atomicAdd(a.accepted[v] + it, taken);
        grid.sync();
        // an iteration that accepted nothing leaves the state as it was: the later ones would repeat it
        if (!settled && a.accepted[v][it] == 0) settled = true;
        bool any_alive = false;  // the same answer in every block of the grid: they leave together
        for (int vv = 0; vv < (int)gridDim.y; ++vv) any_alive |= a.accepted[vv][it] != 0;
        if (!any_alive) break;
        int *t = list; list = next; next = t;
        t = count_p; count_p = next_count_p; next_count_p = t;
    }
}

// ---- bilateral filter on disparity (d_filter_bilateral.cu:222-304) --------
// weight = spatial * colour[(int)|a - s|]; norm += weight; res = fma(s, weight, res);
// out = res / norm (div.rn) — exactly the compiled reference's operation order.
constexpr int kBilW = 32, kBilH = 8;

__global__ void __launch_bounds__(kBilW *kBilH)
k_bilateral(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ spatial,
            const float *__restrict__ colour, int radius, int ncolour, int H, int W)
{
    extern __shared__ float bsm[];
    const int tw = kBilW + 2 * radius, th = kBilH + 2 * radius, kw = 2 * radius + 1;
    float *tile = bsm, *ssp = tile + tw * th, *scol = ssp + kw * kw;
    const int tid = threadIdx.y * kBilW + threadIdx.x, nt = kBilW * kBilH;
    const int bx = blockIdx.x * kBilW, by = blockIdx.y * kBilH;
    for (int i = tid; i < tw * th; i += nt) {
        int ty = i / tw, tx = i - ty * tw;
        tile[i] = in[(size_t)clampi(by + ty - radius, 0, H - 1) * W + clampi(bx + tx - radius, 0, W - 1)];
    }
    for (int i = tid; i < kw * kw; i += nt) ssp[i] = spatial[i];
    for (int i = tid; i < ncolour; i += nt) scol[i] = colour[i];
    __syncthreads();
    const int gx = bx + threadIdx.x, gy = by + threadIdx.y;
    if (gx >= W || gy >= H) return;
    const float va = tile[(threadIdx.y + radius) * tw + threadIdx.x + radius];
    float norm = 0.0f, res = 0.0f;
    for (int y = 0; y < kw; ++y) {
        const float *trow = tile + (threadIdx.y + y) * tw + threadIdx.x;
        const float *srow = ssp + y * kw;
        for (int x = 0; x < kw; ++x) {
            const float vs = trow[x];
            const int ci = min((int)fabsf(__fsub_rn(va, vs)), ncolour - 1);
            const float w = __fmul_rn(srow[x], scol[ci]);
            norm = __fadd_rn(norm, w);
            res = __fmaf_rn(vs, w, res);
        }
    }
    out[(size_t)gy * W + gx] = __fdiv_rn(res, norm);
}

// Register-blocked form for a compile-time radius: a thread produces 4 horizontally adjacent outputs.
// Per kernel row it loads the 4 + 2R tile values they share and the row's 2R + 1 spatial weights with
// LDS.128 (3.3 shared loads per 60 taps instead of 120); only the colour-table lookup stays per tap.
// Each output still accumulates ky-major, kx-minor, exactly like the loop above.  Both views in one
// launch (blockIdx.z).  BOUNDED: |a - s| < 2^23 and (int)|a - s| < ncolour are guaranteed by the caller
// (disparities of this pipeline), so the index is taken with a round-toward-zero add instead of the
// conversion unit and needs no clamp.
constexpr int kBil4W = 128, kBil4H = 8;

template <int R, bool BOUNDED>
__global__ void __launch_bounds__(256)
k_bilateral4(const float *__restrict__ in0, const float *__restrict__ in1, float *__restrict__ out0,
             float *__restrict__ out1, const float *__restrict__ spatial, const float *__restrict__ colour,
             int ncolour, int H, int W)
{
    constexpr int KW = 2 * R + 1, KWP = (KW + 3) & ~3;
    constexpr int TWP = (kBil4W + 2 * R + 3) & ~3, TH = kBil4H + 2 * R;
    constexpr int NV = (4 + 2 * R + 3) & ~3;  // tile values per thread per kernel row, rounded to float4s
    static_assert(kBil4W - 4 + NV <= TWP, "row reads stay inside the padded tile");
    extern __shared__ __align__(16) float bsm4[];
    float *tile = bsm4, *ssp = tile + TWP * TH, *scol = ssp + KWP * KW;
    const float *__restrict__ in = blockIdx.z ? in1 : in0;
    float *__restrict__ out = blockIdx.z ? out1 : out0;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int bx = blockIdx.x * kBil4W, by = blockIdx.y * kBil4H;
    for (int i = tid; i < TWP * TH; i += 256) {
        const int ty = i / TWP, tx = i - ty * TWP;
        tile[i] = in[(size_t)clampi(by + ty - R, 0, H - 1) * W + clampi(bx + tx - R, 0, W - 1)];
    }
    for (int i = tid; i < KWP * KW; i += 256) {
        const int ky = i / KWP, kx = i - ky * KWP;
        ssp[i] = kx < KW ? spatial[ky * KW + kx] : 0.0f;
    }
    for (int i = tid; i < ncolour; i += 256) scol[i] = colour[i];
    __syncthreads();
    const int x0 = 4 * threadIdx.x, gx = bx + x0, gy = by + threadIdx.y;
    if (gx >= W || gy >= H) return;
    float va[4], norm[4], res[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        va[o] = tile[(threadIdx.y + R) * TWP + x0 + R + o];
        norm[o] = 0.0f;
        res[o] = 0.0f;
    }
#pragma unroll 1
    for (int ky = 0; ky < KW; ++ky) {
        float v[NV], w[KWP];
        const float4 *trow = reinterpret_cast<const float4 *>(tile + (threadIdx.y + ky) * TWP + x0);
        const float4 *wrow = reinterpret_cast<const float4 *>(ssp + ky * KWP);
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
            const float4 t = trow[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < KWP / 4; ++i) {
            const float4 t = wrow[i];
            w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
#pragma unroll
            for (int kx = 0; kx < KW; ++kx) {
                const float vs = v[o + kx];
                const float ad = fabsf(__fsub_rn(va[o], vs));
                // BOUNDED: adding 2^21 toward zero leaves 4 * |a - s| (to a quarter) in the mantissa;
                // masking the two fraction bits gives the table's byte offset, 4 * floor(|a - s|)
                int cb;
                if (BOUNDED) cb = __float_as_int(__fadd_rz(ad, 2097152.0f)) & 0x7ffffc;
                else cb = 4 * min((int)ad, ncolour - 1);
                const float wt = __fmul_rn(w[kx], *reinterpret_cast<const float *>(reinterpret_cast<const char *>(scol) + cb));
                norm[o] = __fadd_rn(norm[o], wt);
                res[o] = __fmaf_rn(vs, wt, res[o]);
            }
        }
    }
    float *o4 = out + (size_t)gy * W + gx;
#pragma unroll
    for (int o = 0; o < 4; ++o)
        if (gx + o < W) o4[o] = __fdiv_rn(res[o], norm[o]);
}

// k_bilateral4<R, true> with the per-tap arithmetic issued two outputs at a time (f32x2): the subtract, the
// weight product, the norm add and the value FMA of outputs (0,1) and (2,3) pair up; only the colour-table
// lookup stays per tap — 10 issue slots per two taps instead of 14.  A pair needs its two tile values in one
// aligned 64-bit register: for even kx they are adjacent in the row as loaded (tile A), for odd kx in the same
// row shifted by one (tile B, a second copy of the tile staged next to it).  The spatial weights are the same
// for every thread: they arrive already duplicated into pairs as a kernel parameter and are read through the
// constant path, which keeps them off the shared-memory pipe (the kernel's other ceiling: 60 table lookups
// plus the row loads per 60 taps).  Every output accumulates ky-major, kx-minor with the same operations and
// roundings as k_bilateral4.
template <int R>
struct BilPairs {
    float2 w[2 * R + 1][(2 * R + 2) & ~1];   // (w, w) per tap, rows padded to an even count
};

template <int R>
__global__ void __launch_bounds__(256, 2)
k_bilateral4p(const float *__restrict__ in0, const float *__restrict__ in1, float *__restrict__ out0,
              float *__restrict__ out1, const __grid_constant__ BilPairs<R> wpairs, const float *__restrict__ colour,
              int ncolour, int H, int W)
{
    constexpr int KW = 2 * R + 1;
    constexpr int TWP = (kBil4W + 2 * R + 3) & ~3, TH = kBil4H + 2 * R;
    constexpr int NA = (4 + 2 * R + 3) / 4;                     // float4 loads of tile A per kernel row
    constexpr int NB = (2 + 2 * R + 3) / 4;                     // and of tile B
    static_assert(kBil4W - 4 + 4 * NA <= TWP, "row reads stay inside the padded tile");
    extern __shared__ __align__(16) float bsm4[];
    float *tileA = bsm4, *tileB = tileA + TWP * TH;
    float *scol = tileB + TWP * TH;
    const float *__restrict__ in = blockIdx.z ? in1 : in0;
    float *__restrict__ out = blockIdx.z ? out1 : out0;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int bx = blockIdx.x * kBil4W, by = blockIdx.y * kBil4H;
    // A tile that holds ONE disparity, halo included (flat regions; the whole frame of a pair of identical images),
    // gives every output the same 225 operations on the same operands: one thread forms the value, the block stores it.
    float first = 0.0f;
    int same = 1;
    for (int i = tid; i < TWP * TH; i += 256) {
        const int ty = i / TWP, tx = i - ty * TWP;
        const float *row = in + (size_t)clampi(by + ty - R, 0, H - 1) * W;
        const float va = row[clampi(bx + tx - R, 0, W - 1)], vb = row[clampi(bx + tx + 1 - R, 0, W - 1)];
        tileA[i] = va;
        tileB[i] = vb;
        if (i == tid) first = va;
        same &= (__float_as_uint(va) == __float_as_uint(first)) & (__float_as_uint(vb) == __float_as_uint(first));
    }
    for (int i = tid; i < ncolour; i += 256) scol[i] = colour[i];
    __syncthreads();
    const bool uniform = __syncthreads_and(same && __float_as_uint(first) == __float_as_uint(tileA[0])) != 0;
    const int x0 = 4 * threadIdx.x, gx = bx + x0, gy = by + threadIdx.y;
    const bool inside = gx < W && gy < H;
    __shared__ float uni_out;
    float r[4] = {0.f, 0.f, 0.f, 0.f}, nm[4] = {1.f, 1.f, 1.f, 1.f};
    if (uniform ? tid == 0 : inside) {
        const float *centre = tileA + (threadIdx.y + R) * TWP + x0 + R;
        const f32x2_t nva[2] = {pack2(-centre[0], -centre[1]), pack2(-centre[2], -centre[3])};
        f32x2_t norm[2] = {pack2(0.0f, 0.0f), pack2(0.0f, 0.0f)}, res[2] = {norm[0], norm[0]};
#pragma unroll 1
        for (int ky = 0; ky < KW; ++ky) {
            f32x2_t pa[2 * NA], pb[2 * NB];
            const ulonglong2 *arow = reinterpret_cast<const ulonglong2 *>(tileA + (threadIdx.y + ky) * TWP + x0);
            const ulonglong2 *brow = reinterpret_cast<const ulonglong2 *>(tileB + (threadIdx.y + ky) * TWP + x0);
            const float2 *wrow = wpairs.w[ky];
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                const ulonglong2 t = arow[i];
                pa[2 * i] = t.x; pa[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const ulonglong2 t = brow[i];
                pb[2 * i] = t.x; pb[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int op = 0; op < 2; ++op) {        // outputs (0,1), then (2,3)
#pragma unroll
                for (int kx = 0; kx < KW; ++kx) {
                    // (v[2 op + kx], v[2 op + kx + 1]) of the row
                    const f32x2_t vs = (kx & 1) ? pb[op + (kx - 1) / 2] : pa[op + kx / 2];
                    float d0, d1;
                    unpack2(add2(vs, nva[op]), d0, d1);
                    const int c0 = __float_as_int(__fadd_rz(fabsf(d0), 2097152.0f)) & 0x7ffffc;
                    const int c1 = __float_as_int(__fadd_rz(fabsf(d1), 2097152.0f)) & 0x7ffffc;
                    const float s0 = *reinterpret_cast<const float *>(reinterpret_cast<const char *>(scol) + c0);
                    const float s1 = *reinterpret_cast<const float *>(reinterpret_cast<const char *>(scol) + c1);
                    const f32x2_t wt = mul2(pack2(wrow[kx].x, wrow[kx].y), pack2(s0, s1));
                    norm[op] = add2(norm[op], wt);
                    res[op] = fma2(vs, wt, res[op]);
                }
            }
        }
        unpack2(res[0], r[0], r[1]); unpack2(res[1], r[2], r[3]);
        unpack2(norm[0], nm[0], nm[1]); unpack2(norm[1], nm[2], nm[3]);
    }
    if (uniform) {
        if (tid == 0) uni_out = __fdiv_rn(r[0], nm[0]);  // thread 0's first output is every output of the tile
        __syncthreads();
        if (!inside) return;
        const float q = uni_out;
        float *o4 = out + (size_t)gy * W + gx;
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (gx + o < W) o4[o] = q;
        return;
    }
    if (!inside) return;
    float *o4 = out + (size_t)gy * W + gx;
#pragma unroll
    for (int o = 0; o < 4; ++o)
        if (gx + o < W) o4[o] = __fdiv_rn(r[o], nm[o]);
}

}  // namespace s2mv
