// kernels_prep.cuh — per-pixel preparation: SBS split, BGRx packing, gray,
// census transform, cross arms.  All O(W*H); none of them touches the volume.
#pragma once
#include "common.cuh"

namespace s2mv {

// demux_sbs (d_demux_common.cu:8-33) + mux_average_kernel (d_mux_common.cu:7-21)
// in one pass.  srcL/srcR are two row-pitched BGR images (for an SBS frame they
// alias one buffer, srcR = srcL + 3*W).  Writes 32-bit BGRx pixels (one aligned
// load per pixel for every later stage), gray, and optional packed BGR copies
// (the outermost views of the interlace).
__global__ void __launch_bounds__(256)
k_unpack(const uint8_t *__restrict__ srcL, const uint8_t *__restrict__ srcR, size_t pitch,
         uint32_t *__restrict__ pixL, uint32_t *__restrict__ pixR,
         uint8_t *__restrict__ grayL, uint8_t *__restrict__ grayR,
         uint8_t *__restrict__ bgrL, uint8_t *__restrict__ bgrR, int H, int W)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const float c = 0.3333333333333f;  // 0x3EAAAAAB
    size_t i = (size_t)y * W + x;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        const uint8_t *s = (v ? srcR : srcL) + (size_t)y * pitch + (size_t)x * 3;
        uint32_t b = s[0], g = s[1], r = s[2];
        (v ? pixR : pixL)[i] = b | (g << 8) | (r << 16);
        // PTX of the reference: mul(g,c); fma(b,c,.); fma(r,c,.); cvt.rzi.u32; st.u8
        float f = __fmaf_rn((float)r, c, __fmaf_rn((float)b, c, __fmul_rn((float)g, c)));
        (v ? grayR : grayL)[i] = (uint8_t)__float2uint_rz(f);
        uint8_t *o = v ? bgrR : bgrL;
        if (o) {
            o[i * 3 + 0] = (uint8_t)b;
            o[i * 3 + 1] = (uint8_t)g;
            o[i * 3 + 2] = (uint8_t)r;
        }
    }
}

// mux_average_kernel alone (stage API)
__global__ void __launch_bounds__(256)
k_gray(const uint8_t *__restrict__ img, uint8_t *__restrict__ gray, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c = 0.3333333333333f;
    float b = (float)img[i * 3], g = (float)img[i * 3 + 1], r = (float)img[i * 3 + 2];
    float f = __fmaf_rn(r, c, __fmaf_rn(b, c, __fmul_rn(g, c)));
    gray[i] = (uint8_t)__float2uint_rz(f);
}

// tx_census_9x7_kernel_3 (d_ci_census.cu:18-50).  FULL = the 48-bit string the
// reference stores; !FULL = its low 32 bits, the only ones alu_hamdist_64 ever
// sees (rows y = -1, 1, 2, 3 of the 9x7 window).
template <bool FULL, typename OutT>
__global__ void __launch_bounds__(256)
k_census(const uint8_t *__restrict__ gray, OutT *__restrict__ census, int H, int W)
{
    int gx = blockIdx.x * blockDim.x + threadIdx.x;
    int gy = blockIdx.y;
    if (gx >= W) return;
    uint8_t centre = gray[(size_t)gy * W + gx];
    OutT c = 0;
#pragma unroll
    for (int y = -3; y <= 3; ++y) {
        if (y == 0) continue;
        if (!FULL && y < -1) continue;
        int cy = clampi(gy + y, 0, H - 1);
        const uint8_t *row = gray + (size_t)cy * W;
#pragma unroll
        for (int x = -4; x <= 4; ++x) {
            if (x == 0) continue;
            int cx = clampi(gx + x, 0, W - 1);
            c = (OutT)(c << 1);
            if (row[cx] < centre) c = c + 1;
        }
    }
    census[(size_t)gy * W + gx] = c;
}

// Both views in one launch (blockIdx.z), low-32-bit form, gray staged in shared
// memory: a 64x16 output tile reads a (64+8)x(16+4) byte tile once instead of
// 32 clamped global byte loads per pixel.
constexpr int kCenW = 64, kCenH = 16;
__global__ void __launch_bounds__(kCenW *kCenH / 4)
k_census_tile(const uint8_t *__restrict__ gray0, const uint8_t *__restrict__ gray1, uint32_t *__restrict__ cen0,
              uint32_t *__restrict__ cen1, int H, int W)
{
    constexpr int TW = kCenW + 8, TH = kCenH + 4;  // rows y-1 .. y+3
    __shared__ uint8_t tile[TH][TW];
    const uint8_t *__restrict__ gray = blockIdx.z ? gray1 : gray0;
    uint32_t *__restrict__ census = blockIdx.z ? cen1 : cen0;
    const int bx = blockIdx.x * kCenW, by = blockIdx.y * kCenH;
    const int tid = threadIdx.x, nt = kCenW * kCenH / 4;
    for (int i = tid; i < TW * TH; i += nt) {
        const int ty = i / TW, tx = i - ty * TW;
        tile[ty][tx] = gray[(size_t)clampi(by + ty - 1, 0, H - 1) * W + clampi(bx + tx - 4, 0, W - 1)];
    }
    __syncthreads();
    // tile[ty][tx] = gray[clamp(by + ty - 1)][clamp(bx + tx - 4)]: the reference's own coordinate clamp.
    // Each thread: one column, 4 consecutive rows.
    const int lx = tid % kCenW, ly0 = (tid / kCenW) * 4;
    const int gx = bx + lx;
    if (gx >= W) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ly = ly0 + j, gy = by + ly;
        if (gy >= H) break;
        const uint8_t centre = tile[ly + 1][lx + 4];
        uint32_t c = 0;
#pragma unroll
        for (int y = -1; y <= 3; ++y) {
            if (y == 0) continue;
#pragma unroll
            for (int x = -4; x <= 4; ++x) {
                if (x == 0) continue;
                c <<= 1;
                if (tile[ly + 1 + y][lx + 4 + x] < centre) c += 1;
            }
        }
        census[(size_t)gy * W + gx] = c;
    }
}

// ca_cross_construction_kernel (d_ca_cross.cu:17-172): the arm is assigned
// before the colour test, so it ends on the first failing pixel (Q6).
__device__ __forceinline__ int max_abs_diff3(uint32_t a, uint32_t b)
{
    uint32_t d = __vabsdiffu4(a, b);
    return max(max((int)(d & 0xff), (int)((d >> 8) & 0xff)), (int)((d >> 16) & 0xff));
}

// Both views in one launch, pixels staged in shared memory: the walk is a chain of up to 4*usd dependent
// loads; out of L2 each step costs ~600 cycles, out of shared memory ~30.
constexpr int kArmW = 64, kArmH = 16;
// The colour differences are integers, so the reference's float tests (float)v > t are evaluated as
// v > floor(t) on integers (arm_threshold): the two int->float conversions per step ran on the quarter-rate
// conversion pipe and were what bounded this kernel (ncu: math-pipe throttle 5 of 7.7 stall slots per issue).
__host__ __device__ __forceinline__ int arm_threshold(float t)
{
    return (int)floorf(fminf(fmaxf(t, -1.0f), 1.0e9f));  // any t < 0 passes every test, any t >= 255 none
}

// "some colour byte of d exceeds n" for 0 <= n <= 127 without unpacking the bytes: a byte with its top bit set
// exceeds n; otherwise adding 127 - n to its low seven bits sets the top bit exactly when it does.  k127 holds
// 127 - n in each of the three colour bytes.
__device__ __forceinline__ uint32_t over_bits(uint32_t d, uint32_t k127) { return ((d & 0x007f7f7fu) + k127) | d; }

// SWAR: both thresholds lie in [0, 127] (the launcher checks) and the tests run on packed bytes.
template <bool SWAR>
__device__ __forceinline__ int arm_walk_tile(const uint32_t *__restrict__ t, int pitch, int x, int y, int dx, int dy,
                                             uint32_t anchor, int ucd_i, int lcd_i, int usd, int lsd, int H, int W)
{
    const int step = dy * pitch + dx;
    // steps that stay inside the image (the reference tests the coordinates every step)
    const int limit = min(usd, dx ? (dx > 0 ? W - 1 - x : x) : (dy > 0 ? H - 1 - y : y));
    const int near = min(limit, lsd);
    const uint32_t ku = 0x00010101u * (uint32_t)(127 - ucd_i), kl = 0x00010101u * (uint32_t)(127 - lcd_i);
    const uint32_t *__restrict__ p = t;
    uint32_t prev = anchor;
    int s = 1;
    // within lsd: the colour must stay close to the anchor and to the previous pixel
    for (; s <= near; ++s) {
        p += step;
        const uint32_t c = *p;
        if (SWAR) {
            if ((over_bits(__vabsdiffu4(c, anchor), kl) | over_bits(__vabsdiffu4(c, prev), kl)) & 0x00808080u) return s;
        } else {
            if (max_abs_diff3(c, anchor) > lcd_i || max_abs_diff3(c, prev) > lcd_i) return s;
        }
        prev = c;
    }
    // beyond lsd: close to the anchor under the other threshold
    for (; s <= limit; ++s) {
        p += step;
        const uint32_t c = *p;
        if (SWAR) {
            if (over_bits(__vabsdiffu4(c, anchor), ku) & 0x00808080u) return s;
        } else {
            if (max_abs_diff3(c, anchor) > ucd_i) return s;
        }
    }
    return limit;
}

template <bool SWAR>
__global__ void __launch_bounds__(kArmW *kArmH)
k_arms_tile(const uint32_t *__restrict__ pix0, const uint32_t *__restrict__ pix1, uint32_t *__restrict__ arms0,
            uint32_t *__restrict__ arms1, int ui, int li, int usd, int lsd, int H, int W)
{
    extern __shared__ uint32_t atile[];
    const uint32_t *__restrict__ pix = blockIdx.z ? pix1 : pix0;
    uint32_t *__restrict__ arms = blockIdx.z ? arms1 : arms0;
    const int TW = kArmW + 2 * usd, TH = kArmH + 2 * usd;
    const int bx = blockIdx.x * kArmW, by = blockIdx.y * kArmH;
    const int tid = threadIdx.y * kArmW + threadIdx.x;
    for (int i = tid; i < TW * TH; i += kArmW * kArmH) {
        const int ty = i / TW, tx = i - ty * TW;
        const int gx = bx + tx - usd, gy = by + ty - usd;
        atile[i] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? pix[(size_t)gy * W + gx] : 0u;
    }
    __syncthreads();
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x >= W || y >= H) return;
    const uint32_t *__restrict__ t = atile + (threadIdx.y + usd) * TW + threadIdx.x + usd;
    const uint32_t a = *t;
    const int u = arm_walk_tile<SWAR>(t, TW, x, y, 0, -1, a, ui, li, usd, lsd, H, W);
    const int d = arm_walk_tile<SWAR>(t, TW, x, y, 0, +1, a, ui, li, usd, lsd, H, W);
    const int l = arm_walk_tile<SWAR>(t, TW, x, y, -1, 0, a, ui, li, usd, lsd, H, W);
    const int r = arm_walk_tile<SWAR>(t, TW, x, y, +1, 0, a, ui, li, usd, lsd, H, W);
    arms[(size_t)y * W + x] = (uint32_t)u | ((uint32_t)d << 8) | ((uint32_t)l << 16) | ((uint32_t)r << 24);
}

// ---- half-resolution variant (adcensus_stm_2, d_io.cu:240-508) -------------
// alu_bilinear_interp / alu_bilinear_interp_f (d_alu.cu:17-71) as nvcc -O3 compiles them:
// top = fma(1-wx, v00, wx*v01), bot likewise, res = fma(1-wy, top, wy*bot).
__device__ __forceinline__ float bilerp_ref(float v00, float v01, float v10, float v11, float wx, float wy)
{
    const float iwx = __fsub_rn(1.0f, wx), iwy = __fsub_rn(1.0f, wy);
    const float top = __fmaf_rn(iwx, v00, __fmul_rn(wx, v01));
    const float bot = __fmaf_rn(iwx, v10, __fmul_rn(wx, v11));
    return __fmaf_rn(iwy, top, __fmul_rn(wy, bot));
}
// d_tx_scale.cu:19-20,41-42: fmin(fmax(((float) t / (float) out_n) * (float) in_n, 0), (float)(in_n - 1))
__device__ __forceinline__ float sample_coord_ref(int t, int out_n, int in_n)
{
    const float v = __fmul_rn(__fdiv_rn((float)t, (float)out_n), (float)in_n);
    return fminf(fmaxf(v, 0.0f), (float)(in_n - 1));
}

// tx_scale_bilinear_kernel (d_tx_scale.cu:30-52): tightly packed BGR in, tightly packed BGR out
__global__ void __launch_bounds__(256)
k_scale_bilinear_bgr(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int in_rows, int in_cols, int out_rows,
                     int out_cols)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y;
    if (gx >= out_cols || gy >= out_rows) return;
    const float xs = sample_coord_ref(gx, out_cols, in_cols), ys = sample_coord_ref(gy, out_rows, in_rows);
    const int x0 = (int)floorf(xs), y0 = (int)floorf(ys);
    const int x1 = min(x0 + 1, in_cols - 1), y1 = min(y0 + 1, in_rows - 1);
    const float wx = __fsub_rn(xs, (float)x0), wy = __fsub_rn(ys, (float)y0);
    const uint8_t *p00 = in + ((size_t)y0 * in_cols + x0) * 3, *p01 = in + ((size_t)y0 * in_cols + x1) * 3;
    const uint8_t *p10 = in + ((size_t)y1 * in_cols + x0) * 3, *p11 = in + ((size_t)y1 * in_cols + x1) * 3;
    uint8_t *o = out + ((size_t)gy * out_cols + gx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        o[c] = (uint8_t)__float2uint_rz(bilerp_ref((float)p00[c], (float)p01[c], (float)p10[c], (float)p11[c], wx, wy));
}

// tx_disp_scale_kernel (d_tx_scale.cu:8-28)
__global__ void __launch_bounds__(256)
k_disp_scale(float *__restrict__ out, const float *__restrict__ in, int out_rows, int out_cols, int in_rows, int in_cols,
             float scale)
{
    const int tx = blockIdx.x * blockDim.x + threadIdx.x, ty = blockIdx.y;
    if (tx >= out_cols || ty >= out_rows) return;
    const float xs = sample_coord_ref(tx, out_cols, in_cols), ys = sample_coord_ref(ty, out_rows, in_rows);
    const int x0 = (int)floorf(xs), y0 = (int)floorf(ys);
    const int x1 = min(x0 + 1, in_cols - 1), y1 = min(y0 + 1, in_rows - 1);
    const float wx = __fsub_rn(xs, (float)x0), wy = __fsub_rn(ys, (float)y0);
    const float v = bilerp_ref(in[(size_t)y0 * in_cols + x0], in[(size_t)y0 * in_cols + x1], in[(size_t)y1 * in_cols + x0],
                               in[(size_t)y1 * in_cols + x1], wx, wy);
    out[(size_t)ty * out_cols + tx] = __fmul_rn(v, scale);
}

// packed arms <-> the reference's four byte planes (UP, DOWN, LEFT, RIGHT)
__global__ void k_arms_unpack(const uint32_t *__restrict__ arms, uint8_t *__restrict__ planes, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = arms[i];
    planes[i] = a & 0xff;
    planes[n + i] = (a >> 8) & 0xff;
    planes[2 * n + i] = (a >> 16) & 0xff;
    planes[3 * n + i] = a >> 24;
}
__global__ void k_arms_pack(const uint8_t *__restrict__ planes, uint32_t *__restrict__ arms, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    arms[i] = (uint32_t)planes[i] | ((uint32_t)planes[n + i] << 8) | ((uint32_t)planes[2 * n + i] << 16) |
              ((uint32_t)planes[3 * n + i] << 24);
}

// Exponential tables of the combine step, built with the reference's own
// instruction sequence so the table entries are bit-identical to what
// ci_adcensus_kernel would compute per element.
__global__ void k_build_luts(float inv_ad, float inv_cen, float *__restrict__ lut_ad, float *__restrict__ lut_cen)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kAdLutSize) {
        // ci_ad_kernel_5: (float)(|dB|+|dG|+|dR|) * 0.33333333333f  (one mul.f32)
        float ad = __fmul_rn((float)i, 0.33333333333f);
        lut_ad[i] = ref_one_minus_exp(ad, inv_ad);
    }
    if (i < kCenLutSize) lut_cen[i] = ref_one_minus_exp((float)i, inv_cen);
}

}  // namespace s2mv
