// kernels_refine.cuh — disparity refinement: left/right cross-check with
// disocclusion labelling, iterative region voting, bilateral filter.
#pragma once
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

// ---- iterative region voting (d_dr_irv.cu:17-43,134-269) -----------------
// The reference gives every pixel a thread and every outlier thread a private
// 65-int histogram in local memory.  Here the outliers (typically 10-20% of
// the image, shrinking each iteration) are compacted into a list and each one
// is voted on by a whole warp: lanes sweep the rows of the cross-shaped
// support, the histogram lives in shared memory.  Votes read a snapshot and
// are applied by a separate kernel (the race-free reading of Q15).
struct IrvArgs {
    float *disp[2];
    uint8_t *outliers[2];
    const uint32_t *arms[2];
    int *list[2];    // outlier pixel indices
    int *vote[2];    // accepted disparity or kNoVote
    int *count[2];   // list length
    int H, W, nbins, zd, usd, thresh_s;
    float thresh_h;
};
constexpr int kNoVote = -0x7fffffff;

__global__ void __launch_bounds__(256)
k_irv_compact(const IrvArgs a)
{
    const int v = blockIdx.y;
    const size_t n = (size_t)a.H * a.W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (a.outliers[v][i] != 0) {
        int k = atomicAdd(a.count[v], 1);
        a.list[v][k] = (int)i;
    }
}

constexpr int kIrvWarps = 8;

__global__ void __launch_bounds__(kIrvWarps * 32)
k_irv_vote(const IrvArgs a)
{
    extern __shared__ int hist_all[];  // [kIrvWarps][nbins]
    const int v = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *hist = hist_all + warp * a.nbins;
    const int count = *a.count[v];
    const float *__restrict__ disp = a.disp[v];
    const uint8_t *__restrict__ outl = a.outliers[v];
    const uint32_t *__restrict__ arms = a.arms[v];
    const int W = a.W;
    for (int e = blockIdx.x * kIrvWarps + warp; e < count; e += gridDim.x * kIrvWarps) {
        const int pix = a.list[v][e];
        const int gy = pix / W, gx = pix - gy * W;
        for (int b = lane; b < a.nbins; b += 32) hist[b] = 0;
        __syncwarp();
        const uint32_t ac = arms[pix];
        const int cu = min(arm_up(ac), a.usd), cd = arm_down(ac);
        int cnt = 0;
        for (int y = -cu; y <= cd; ++y) {
            const size_t row = (size_t)(gy + y) * W;
            const uint32_t ar = arms[row + gx];
            const int cl = arm_left(ar), span = cl + arm_right(ar) + 1;  // inclusive [-L, R]
            for (int k = lane; k < span; k += 32) {
                const size_t s = row + (gx - cl + k);
                if (outl[s] == 0) {
                    int bin = clampi((int)disp[s] + a.zd, 0, a.nbins - 1);
                    atomicAdd(&hist[bin], 1);
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
            a.vote[v][e] = ok ? max_d : kNoVote;
        }
    }
}

__global__ void __launch_bounds__(256)
k_irv_apply(const IrvArgs a)
{
    const int v = blockIdx.y;
    const int count = *a.count[v];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const int vote = a.vote[v][e];
        if (vote != kNoVote) {
            const int pix = a.list[v][e];
            a.outliers[v][pix] = 0;
            a.disp[v][pix] = (float)vote;
        }
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

}  // namespace s2mv
