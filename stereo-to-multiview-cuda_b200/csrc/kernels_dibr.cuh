// kernels_dibr.cuh — depth-image-based rendering: coverage maps, mask
// dilation, backward warps + blend for the intermediate views, and the
// slanted-lenticular interlace.
#pragma once
#include "common.cuh"

namespace s2mv {

// dibr_find_occlusion_kernel for both directions (d_dibr_occl.cu:114-159):
// occl_r[x + trunc(disp_l)] = 1, occl_l[x + trunc(-disp_r)] = 1 ("1" = covered).
// Both maps must be zeroed first.  All writers store 1: order-free.
__global__ void __launch_bounds__(256)
k_occl(const float *__restrict__ dispL, const float *__restrict__ dispR, uint8_t *__restrict__ occlL,
       uint8_t *__restrict__ occlR, int H, int W)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const size_t row = (size_t)y * W;
    int sd = (int)__fmul_rn(dispL[row + x], 1.0f);
    occlR[row + clampi(x + sd, 0, W - 1)] = 1;
    sd = (int)__fmul_rn(dispR[row + x], -1.0f);
    occlL[row + clampi(x + sd, 0, W - 1)] = 1;
}

// filter_bleed_1_kernel (d_filter.cu:105-139) and, optionally fused,
// dibr_occl_to_mask_kernel (d_dibr_occl.cu:17-31).
__global__ void __launch_bounds__(256)
k_bleed(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, float *__restrict__ mask, int radius,
        int H, int W)
{
    int tx = blockIdx.x * blockDim.x + threadIdx.x;
    int ty = blockIdx.y;
    if (tx >= W) return;
    const int kernel_sz = (2 * radius + 1) * (2 * radius + 1);
    int count = 0;
    for (int y = -radius; y <= radius; ++y)
        for (int x = -radius; x <= radius; ++x) {
            int sx = tx + x, sy = ty + y;
            if (sx < 0) sx = -sx;
            if (sy < 0) sy = -sy;
            if (sx > W - 1) sx = W - 1 - x;
            if (sy > H - 1) sy = H - 1 - y;
            if (in[(size_t)sy * W + sx] > 0) ++count;
        }
    const uint8_t a = in[(size_t)ty * W + tx];
    const uint8_t r = ((double)count > (kernel_sz - 1) * 0.30) ? (uint8_t)1 : a;
    out[(size_t)ty * W + tx] = r;
    if (mask) mask[(size_t)ty * W + tx] = (r == 1) ? 1.0f : 0.0f;
}

// k_occl, k_bleed (radius 1) and the mask conversion of BOTH views for one image row per block: the occlusion marks
// of a row come from that row's disparities alone, so the three rows the 3x3 bleed reads are formed in shared memory
// ([2 views][3 rows][W] bytes) and only the two float masks are written -- one launch instead of two memsets and
// three kernels over four byte planes.  Row and column reflection exactly as k_bleed's.  H >= 3.
__global__ void __launch_bounds__(256)
k_occl_bleed_mask_row(const float *__restrict__ dispL, const float *__restrict__ dispR, float *__restrict__ maskL,
                      float *__restrict__ maskR, int H, int W)
{
    extern __shared__ uint8_t occ_sm[];
    const int ty = blockIdx.x;
    uint8_t *oL = occ_sm, *oR = occ_sm + 3 * (size_t)W;  // [k][x], k = 0..2 <-> dy = -1..1
    for (int i = threadIdx.x; i < 6 * W; i += blockDim.x) occ_sm[i] = 0;
    __syncthreads();
    int rows[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int y = k - 1;
        int sy = ty + y;
        if (sy < 0) sy = -sy;
        if (sy > H - 1) sy = H - 1 - y;
        rows[k] = sy;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const size_t row = (size_t)rows[k] * W;
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            int sd = (int)__fmul_rn(dispL[row + x], 1.0f);
            oR[k * W + clampi(x + sd, 0, W - 1)] = 1;  // every writer stores 1: order-free
            sd = (int)__fmul_rn(dispR[row + x], -1.0f);
            oL[k * W + clampi(x + sd, 0, W - 1)] = 1;
        }
    }
    __syncthreads();
    const size_t row = (size_t)ty * W;
    for (int tx = threadIdx.x; tx < W; tx += blockDim.x) {
        int cl = 0, cr = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int x = -1; x <= 1; ++x) {
                int sx = tx + x;
                if (sx < 0) sx = -sx;
                if (sx > W - 1) sx = W - 1 - x;
                cl += oL[k * W + sx] > 0;
                cr += oR[k * W + sx] > 0;
            }
        // (double)count > (9 - 1) * 0.30 = 2.4
        const uint8_t rl = cl > 2 ? (uint8_t)1 : oL[W + tx], rr = cr > 2 ? (uint8_t)1 : oR[W + tx];
        maskL[row + tx] = rl == 1 ? 1.0f : 0.0f;
        maskR[row + tx] = rr == 1 ? 1.0f : 0.0f;
    }
}

__global__ void k_occl_to_mask(const uint8_t *__restrict__ occl, float *__restrict__ mask, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = occl[i] == 1 ? 1.0f : 0.0f;
}

// filter_gaussian_1_kernel_1 (d_filter_gaussian.cu:9-88): out = max(v, blur(v)),
// with op_invertnormf_kernel (d_op.cu:7-16) optionally folded into the tile load.
constexpr int kGaW = 32, kGaH = 8;

__global__ void __launch_bounds__(kGaW *kGaH)
k_gauss_dilate(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ kernel,
               int radius, int invert, int H, int W)
{
    extern __shared__ float gsm[];
    const int tw = kGaW + 2 * radius, th = kGaH + 2 * radius, kw = 2 * radius + 1;
    float *tile = gsm, *sk = tile + tw * th;
    const int tid = threadIdx.y * kGaW + threadIdx.x, nt = kGaW * kGaH;
    const int bx = blockIdx.x * kGaW, by = blockIdx.y * kGaH;
    for (int i = tid; i < tw * th; i += nt) {
        int ty = i / tw, tx = i - ty * tw;
        float v = in[(size_t)clampi(by + ty - radius, 0, H - 1) * W + clampi(bx + tx - radius, 0, W - 1)];
        tile[i] = invert ? __fsub_rn(1.0f, v) : v;
    }
    for (int i = tid; i < kw * kw; i += nt) sk[i] = kernel[i];
    __syncthreads();
    const int gx = bx + threadIdx.x, gy = by + threadIdx.y;
    if (gx >= W || gy >= H) return;
    const float va = tile[(threadIdx.y + radius) * tw + threadIdx.x + radius];
    float res = 0.0f, norm = 0.0f;
    for (int y = 0; y < kw; ++y) {
        const float *trow = tile + (threadIdx.y + y) * tw + threadIdx.x;
        const float *krow = sk + y * kw;
        for (int x = 0; x < kw; ++x) {
            const float w = krow[x];
            norm = __fadd_rn(norm, w);
            res = __fmaf_rn(trow[x], w, res);
        }
    }
    const float q = __fdiv_rn(res, norm);
    out[(size_t)gy * W + gx] = (va < q) ? q : va;
}

// Register-blocked form for a compile-time radius (see k_bilateral4): 4 adjacent outputs per thread,
// tile values and kernel weights loaded by LDS.128 once per kernel row, one FFMA per tap.  `norm` — the
// sequential fp32 sum of the weights, identical for every pixel — is computed once by the host in the
// same order (host_kernel_norm) instead of 441 times per pixel.
constexpr int kGa4W = 128, kGa4H = 8;

template <int R>
__global__ void __launch_bounds__(256)
k_gauss_dilate4(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ kernel, float norm,
                int invert, int H, int W)
{
    constexpr int KW = 2 * R + 1, KWP = (KW + 3) & ~3;
    constexpr int TWP = (kGa4W + 2 * R + 3) & ~3, TH = kGa4H + 2 * R;
    constexpr int NV = (4 + 2 * R + 3) & ~3;
    static_assert(kGa4W - 4 + NV <= TWP, "row reads stay inside the padded tile");
    extern __shared__ __align__(16) float gsm4[];
    float *tile = gsm4, *sk = tile + TWP * TH;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int bx = blockIdx.x * kGa4W, by = blockIdx.y * kGa4H;
    // Masks are 0/1 planes with the occlusions as islands: most tiles, halo included, hold one value.  Every output
    // of such a tile is the same number -- the same 441 operations on the same operands -- so one thread forms it
    // (with the same loop, not a closed form) and the block stores it.
    float first = 0.0f;
    int same = 1;
    for (int i = tid; i < TWP * TH; i += 256) {
        const int ty = i / TWP, tx = i - ty * TWP;
        const float v = in[(size_t)clampi(by + ty - R, 0, H - 1) * W + clampi(bx + tx - R, 0, W - 1)];
        const float t = invert ? __fsub_rn(1.0f, v) : v;
        tile[i] = t;
        if (i == tid) first = t;
        same &= (__float_as_uint(t) == __float_as_uint(first));
    }
    for (int i = tid; i < KWP * KW; i += 256) {
        const int ky = i / KWP, kx = i - ky * KWP;
        sk[i] = kx < KW ? kernel[ky * KW + kx] : 0.0f;
    }
    __syncthreads();
    const bool uniform = __syncthreads_and(same && __float_as_uint(first) == __float_as_uint(tile[0])) != 0;
    const int x0 = 4 * threadIdx.x, gx = bx + x0, gy = by + threadIdx.y;
    const bool inside = gx < W && gy < H;
    __shared__ float uni_q;
    float res[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (uniform ? tid == 0 : inside) {
#pragma unroll 1
        for (int ky = 0; ky < KW; ++ky) {
            float v[NV], w[KWP];
            const float4 *trow = reinterpret_cast<const float4 *>(tile + (threadIdx.y + ky) * TWP + x0);
            const float4 *wrow = reinterpret_cast<const float4 *>(sk + ky * KWP);
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
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int kx = 0; kx < KW; ++kx) res[o] = __fmaf_rn(v[o + kx], w[kx], res[o]);
        }
    }
    if (uniform) {
        if (tid == 0) {  // its first output is every output of the tile
            const float va = tile[R * TWP + R];
            const float q = __fdiv_rn(res[0], norm);
            uni_q = (va < q) ? q : va;
        }
        __syncthreads();
        if (!inside) return;
        const float q = uni_q;
        float *o4 = out + (size_t)gy * W + gx;
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (gx + o < W) o4[o] = q;
        return;
    }
    if (!inside) return;
    float *o4 = out + (size_t)gy * W + gx;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        if (gx + o >= W) break;
        const float va = tile[(threadIdx.y + R) * TWP + x0 + R + o];
        const float q = __fdiv_rn(res[o], norm);
        o4[o] = (va < q) ? q : va;
    }
}

// One intermediate view per blockIdx.z: the two backward warps
// (dibr_backward_warp_kernel, d_dibr_bwarp.cu:5-22) and the blend
// (mux_merge_AB_kernel, d_mux_common.cu:23-46) of d_dibr_dbm, fused.
// `tmask` is the dilated inverse of mask_r, which does not depend on the view,
// so it is computed once per frame instead of once per view (d_dibr_bwarp.cu:60-63).
struct DbmArgs {
    const uint32_t *pixL, *pixR;
    const float *dispL, *dispR, *maskL, *maskR, *tmask;
    uint8_t *views;      // [num_views][H][W][3]
    float shift[16];     // per intermediate view
    int view_index[16];
    int H, W;
};

// The conversions of this kernel (28 per pixel and view) ran on the quarter-rate conversion pipe and bounded
// it (ncu: 82 % of that pipe).  They are all between small non-negative integers and floats, so they are done
// on the FP32 pipe with the 2^23 bias instead, bit for bit:
//   (float)n       = as_float(0x4b000000 | n) - 2^23                    for 0 <= n < 2^23
//   (uint)trunc(v) = as_uint(v + 2^23, rounded toward zero) & 0x7fffff   for 0 <= v < 2^23
constexpr float kTwo23 = 8388608.0f;
__device__ __forceinline__ float small_to_float(uint32_t n) { return __fsub_rn(__uint_as_float(0x4b000000u | n), kTwo23); }
__device__ __forceinline__ uint32_t trunc_bits(float v) { return __float_as_uint(__fadd_rz(v, kTwo23)); }

__device__ __forceinline__ uint32_t warp_fetch(const uint32_t *__restrict__ pixrow, float disp, float shift, float ftx, int W)
{
    // PTX of the reference: fma.rn(shift, disp, (float)tx); max 0; min W-1; cvt.rzi
    float fx = __fmaf_rn(shift, disp, ftx);
    fx = fminf(fmaxf(fx, 0.0f), small_to_float((uint32_t)(W - 1)));
    return pixrow[trunc_bits(fx) & 0x7fffffu];  // bilinear at integral coordinates = plain fetch (Q23)
}

__global__ void __launch_bounds__(256)
k_dbm(const DbmArgs a)
{
    int tx = blockIdx.x * blockDim.x + threadIdx.x;
    int ty = blockIdx.y;
    if (tx >= a.W) return;
    const int vi = blockIdx.z;
    const float shift = a.shift[vi];
    const size_t row = (size_t)ty * a.W, i = row + tx;
    // masks are sums of non-negative weights; a (never observed) negative or NaN product must still convert
    // to 0 as cvt.rzi.u32 does, hence the max with 0 on the mask factors
    const float mr = fmaxf(a.maskR[i], 0.0f), ml = fmaxf(a.maskL[i], 0.0f), m = fmaxf(a.tmask[i], 0.0f);
    const float ftx = small_to_float((uint32_t)tx);
    const uint32_t pl = warp_fetch(a.pixL + row, a.dispR[i], -shift, ftx, a.W);
    const uint32_t pr = warp_fetch(a.pixR + row, a.dispL[i], (float)(1.0 - (double)shift), ftx, a.W);
    const float im = fmaxf(__fsub_rn(1.0f, a.tmask[i]), 0.0f);
    uint8_t *o = a.views + ((size_t)a.view_index[vi] * a.H * a.W + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // byte c of the pixel under the 2^23 bias: one PRMT
        const float cl = __fsub_rn(__uint_as_float(__byte_perm(pl, 0x4b000000u, 0x7440 + c)), kTwo23);
        const float cr = __fsub_rn(__uint_as_float(__byte_perm(pr, 0x4b000000u, 0x7440 + c)), kTwo23);
        const float wl = small_to_float(trunc_bits(__fmul_rn(cl, mr)) & 0xffu);  // warped left, masked by mask_r
        const float wr = small_to_float(trunc_bits(__fmul_rn(cr, ml)) & 0xffu);  // warped right, masked by mask_l
        const uint32_t b = trunc_bits(__fmul_rn(im, wl));
        const uint32_t q = trunc_bits(__fmul_rn(m, wr));
        o[c] = (uint8_t)(b + q);                                                   // low bytes add mod 256
    }
}

// Four horizontally adjacent pixels per thread, for W % 4 == 0 and 16-byte aligned planes (the launcher
// checks): the masks and disparities arrive as float4 and the 12 output bytes leave as three aligned words, so
// a warp writes 384 contiguous bytes instead of three interleaved runs of single bytes.  Per pixel the
// arithmetic is k_dbm's.
__global__ void __launch_bounds__(256)
k_dbm4(const DbmArgs a)
{
    const int W4 = a.W >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= W4 * a.H) return;
    const int ty = idx / W4, x0 = (idx - ty * W4) * 4;
    const int vi = blockIdx.z;
    const float shift = a.shift[vi], shift_r = (float)(1.0 - (double)shift);
    const size_t row = (size_t)ty * a.W, i = row + x0;
    const float4 mr4 = __ldg(reinterpret_cast<const float4 *>(a.maskR + i));
    const float4 ml4 = __ldg(reinterpret_cast<const float4 *>(a.maskL + i));
    const float4 tm4 = __ldg(reinterpret_cast<const float4 *>(a.tmask + i));
    const float4 dr4 = __ldg(reinterpret_cast<const float4 *>(a.dispR + i));
    const float4 dl4 = __ldg(reinterpret_cast<const float4 *>(a.dispL + i));
    const float mrs[4] = {mr4.x, mr4.y, mr4.z, mr4.w}, mls[4] = {ml4.x, ml4.y, ml4.z, ml4.w};
    const float tms[4] = {tm4.x, tm4.y, tm4.z, tm4.w};
    const float drs[4] = {dr4.x, dr4.y, dr4.z, dr4.w}, dls[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
    uint32_t pl[4], pr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float ftx = small_to_float((uint32_t)(x0 + j));
        pl[j] = warp_fetch(a.pixL + row, drs[j], -shift, ftx, a.W);
        pr[j] = warp_fetch(a.pixR + row, dls[j], shift_r, ftx, a.W);
    }
    uint32_t bytes[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float mr = fmaxf(mrs[j], 0.0f), ml = fmaxf(mls[j], 0.0f), m = fmaxf(tms[j], 0.0f);
        const float im = fmaxf(__fsub_rn(1.0f, tms[j]), 0.0f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float cl = __fsub_rn(__uint_as_float(__byte_perm(pl[j], 0x4b000000u, 0x7440 + c)), kTwo23);
            const float cr = __fsub_rn(__uint_as_float(__byte_perm(pr[j], 0x4b000000u, 0x7440 + c)), kTwo23);
            const float wl = small_to_float(trunc_bits(__fmul_rn(cl, mr)) & 0xffu);
            const float wr = small_to_float(trunc_bits(__fmul_rn(cr, ml)) & 0xffu);
            bytes[3 * j + c] = trunc_bits(__fmul_rn(im, wl)) + trunc_bits(__fmul_rn(m, wr));  // low byte is the result
        }
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(a.views + ((size_t)a.view_index[vi] * a.H * a.W + i) * 3);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t lo = __byte_perm(bytes[4 * k], bytes[4 * k + 1], 0x0040);      // b0 | b1 << 8
        const uint32_t hi = __byte_perm(bytes[4 * k + 2], bytes[4 * k + 3], 0x0040);  // b2 | b3 << 8
        o[k] = __byte_perm(lo, hi, 0x5410);
    }
}

// mux_multiview_kernel_2 / mux_multiview_kernel (d_mux_multiview.cu:38-124)
struct MuxArgs {
    const uint8_t *views[16];
    uint8_t *out;
    int num_views, Hin, Win, Hout, Wout, variant;
    float y_interval, inv_y_interval;
    int rint_y;  // (int)roundf(y_interval)
    // row band of a taller frame: local row 0 is frame row `row0`; the resampling and the interlace phase
    // are those of the frame (Hframe_in x Hframe_out).  Whole image: 0, Hin, Hout.
    int row0, Hframe_in, Hframe_out;
};

__device__ __forceinline__ uint8_t bilinear_u8(const uint8_t *__restrict__ data, int off, int x0, int x1, int y0,
                                                int y1, float wx, float wy, int W)
{
    // fast_bilinear_interp (d_mux_multiview.cu:10-36) with the PTX's fma placement
    const float v00 = (float)data[((size_t)y0 * W + x0) * 3 + off], v01 = (float)data[((size_t)y0 * W + x1) * 3 + off];
    const float v10 = (float)data[((size_t)y1 * W + x0) * 3 + off], v11 = (float)data[((size_t)y1 * W + x1) * 3 + off];
    const float iwx = __fsub_rn(1.0f, wx), iwy = __fsub_rn(1.0f, wy);
    const float top = __fmaf_rn(iwx, v00, __fmul_rn(wx, v01));
    const float bot = __fmaf_rn(iwx, v10, __fmul_rn(wx, v11));
    return (uint8_t)__float2uint_rz(__fmaf_rn(iwy, top, __fmul_rn(wy, bot)));
}

__global__ void __launch_bounds__(256)
k_mux(const MuxArgs a)
{
    int tx = blockIdx.x * blockDim.x + threadIdx.x;
    int ty = blockIdx.y;
    if (tx >= a.Wout) return;
    float xs = __fmul_rn(__fdiv_rn((float)tx, (float)a.Wout), (float)a.Win);
    const int ty_local = ty;
    ty += a.row0;  // frame row
    float ys = __fmul_rn(__fdiv_rn((float)ty, (float)a.Hframe_out), (float)a.Hframe_in);
    xs = (float)fmin(fmax((double)xs, 0.0), (double)(float)(a.Win - 1));
    ys = (float)fmin(fmax((double)ys, 0.0), (double)(float)(a.Hframe_in - 1));
    const float xi = (float)a.num_views;
    float yv;
    if (a.variant == 2) {
        yv = __fadd_rn((float)(ty % a.rint_y), 1.0f);
        yv = __fmul_rn(yv, xi);
        yv = __fmul_rn(a.inv_y_interval, yv);
    } else {
        yv = (float)((double)(ty % a.rint_y) + 1.0);
        yv = __fdiv_rn(__fmul_rn(yv, xi), a.y_interval);
    }
    int xv = (tx * 3 + (int)yv) % ((int)xi);
    int rv = xv < 0 ? xv + a.num_views : xv;
    int gv = rv + 1, bv = rv + 2;
    if (gv >= a.num_views) gv -= a.num_views;
    if (bv >= a.num_views) bv -= a.num_views;
    const int x0 = (int)floorf(xs), yf = (int)floorf(ys);
    const int x1 = min(x0 + 1, a.Win - 1);
    const float wx = __fsub_rn(xs, (float)x0), wy = __fsub_rn(ys, (float)yf);
    // frame rows -> rows of this (sub-)image; a band's apron rows may fall outside it and are never kept
    const int y0 = min(max(yf - a.row0, 0), a.Hin - 1);
    const int y1 = min(max(min(yf + 1, a.Hframe_in - 1) - a.row0, 0), a.Hin - 1);
    uint8_t *o = a.out + ((size_t)ty_local * a.Wout + tx) * 3;
    o[0] = bilinear_u8(a.views[bv], 0, x0, x1, y0, y1, wx, wy, a.Win);
    o[1] = bilinear_u8(a.views[gv], 1, x0, x1, y0, y1, wx, wy, a.Win);
    o[2] = bilinear_u8(a.views[rv], 2, x0, x1, y0, y1, wx, wy, a.Win);
}


// ---- forward warp (d_dibr_fwarp.cu:9-25, d_dibr_dfm :27-95) ---------------------------------------------
// dibr_forward_warp_kernel scatters every source pixel to column clamp(x + trunc(disp * shift)); when several
// sources of a row land on one destination the reference's result depends on thread timing (last
// writer wins, no ordering).  Here the collision is resolved deterministically: THE LOWEST SOURCE COLUMN WINS
// (what a right-to-left scan of the row produces -- and what the reference's kernel, built for sm_100, is observed
// to produce on a B200 for ~98 % of the colliding destinations: tests/test_fwarp.py prints the figure).  Pass 1
// records the winning source per destination with an atomicMin, pass 2 gathers; destinations nothing maps to stay
// 0 (the reference memsets its output).  Where a destination has at most one source the result is the
// reference's, which the test checks against the reference's own kernel.
__global__ void __launch_bounds__(256)
k_fwarp_claim(const float *__restrict__ disp, float shift, int *__restrict__ winner, int H, int W)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t row = (size_t)y * W;
    const int sd = (int)__fmul_rn(disp[row + x], shift);
    atomicMin(winner + row + clampi(x + sd, 0, W - 1), x);
}

__global__ void __launch_bounds__(256)
k_fwarp_gather(const uint8_t *__restrict__ in, const int *__restrict__ winner, uint8_t *__restrict__ out, int H, int W)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    const int s = winner[i];
    uint8_t b = 0, g = 0, r = 0;
    if (s < W) {
        const uint8_t *p = in + ((size_t)y * W + s) * 3;
        b = p[0]; g = p[1]; r = p[2];
    }
    out[i * 3] = b; out[i * 3 + 1] = g; out[i * 3 + 2] = r;
}

// mux_merge_AB_kernel (d_mux_common.cu:23-46): b = (u8)trunc((1 - m) * b) + (u8)trunc(m * a), per channel
__global__ void __launch_bounds__(256)
k_merge_ab(uint8_t *__restrict__ img_b, const uint8_t *__restrict__ img_a, const float *__restrict__ mask_a, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = mask_a ? mask_a[i] : 0.0f, im = __fsub_rn(1.0f, m);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const unsigned a = (unsigned)__fmul_rn(m, (float)img_a[i * 3 + c]);
        const unsigned b = (unsigned)__fmul_rn(im, (float)img_b[i * 3 + c]);
        img_b[i * 3 + c] = (uint8_t)((uint8_t)b + (uint8_t)a);
    }
}

}  // namespace s2mv
