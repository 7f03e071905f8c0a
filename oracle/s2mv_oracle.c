/*
 * s2mv_oracle.c — CPU restatement of the reference pipeline.  See
 * s2mv_oracle.h for scope and parity status.  TEST INFRASTRUCTURE ONLY.
 *
 * Build: gcc -O3 -mavx2 -mfma -ffp-contract=off -fopenmp -shared -fPIC
 * (-ffp-contract=off is required: every fused multiply-add below is an
 *  explicit fmaf placed where the reference's compiled PTX has fma.rn.f32,
 *  and nowhere else.)
 *
 * Undefined behaviour in the reference is resolved to the in-domain formula
 * (SURVEY §2.3): partial 160-wide blocks read clamped pixels (Q5), the
 * transposes are exact for any H,W (Q9), the IRV histogram has num_disp bins
 * (Q14) and reads a consistent snapshot (Q15), the bilateral tile is fully
 * populated (Q19), every output pixel of the interlace is written (Q29).
 */
#include "s2mv_oracle.h"

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_BLOCK_W 160 /* d_ci_adcensus.cu:46 — the CI kernels' block width */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------- demux */
void orc_demux_sbs(const uint8_t *sbs, uint8_t *img_l, uint8_t *img_r,
                   int num_rows, int num_cols_sbs, int num_cols, int elem_sz)
{
    /* d_demux_common.cu:17-31: columns < num_cols go left, the rest go right
     * at column tx - num_cols.  Columns beyond 2*num_cols would land outside
     * the right image in the reference; they are ignored here. */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < num_rows; ++y)
        for (int x = 0; x < num_cols_sbs; ++x) {
            const uint8_t *s = sbs + ((size_t)y * num_cols_sbs + x) * elem_sz;
            uint8_t *d;
            if (x < num_cols)
                d = img_l + ((size_t)y * num_cols + x) * elem_sz;
            else if (x - num_cols < num_cols)
                d = img_r + ((size_t)y * num_cols + (x - num_cols)) * elem_sz;
            else
                continue;
            d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
        }
}

/* ----------------------------------------------------------------- gray */
void orc_gray(const uint8_t *img, uint8_t *gray, int num_rows, int num_cols, int elem_sz)
{
    const float c = 0.3333333333333f; /* rounds to 0x3EAAAAAB, as all three literals do */
    size_t n = (size_t)num_rows * num_cols;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float b = (float)img[i * elem_sz], g = (float)img[i * elem_sz + 1], r = (float)img[i * elem_sz + 2];
        float v = fmaf(r, c, fmaf(b, c, g * c)); /* PTX of mux_average_kernel */
        gray[i] = (uint8_t)(unsigned)v;          /* cvt.rzi.u32.f32 + st.u8   */
    }
}

/* --------------------------------------------------------------- census */
void orc_census(const uint8_t *gray, uint64_t *census, int num_rows, int num_cols)
{
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < num_rows; ++gy)
        for (int gx = 0; gx < num_cols; ++gx) {
            uint64_t c = 0;
            uint8_t centre = gray[(size_t)gy * num_cols + gx];
            for (int y = -3; y <= 3; ++y)
                for (int x = -4; x <= 4; ++x) {
                    if (x == 0 || y == 0) continue; /* d_ci_census.cu:41 */
                    int cx = clampi(gx + x, 0, num_cols - 1);
                    int cy = clampi(gy + y, 0, num_rows - 1);
                    c <<= 1;
                    if (gray[(size_t)cy * num_cols + cx] < centre) c += 1;
                }
            census[(size_t)gy * num_cols + gx] = c;
        }
}

int orc_hamdist(uint64_t a, uint64_t b)
{
    /* d_alu.cu:9: `int c = a ^ b;` keeps the low 32 bits as a signed int, then
     * 64 arithmetic right shifts: bits 0..30 once each, the sign bit 33 times. */
    uint32_t x = (uint32_t)(a ^ b);
    return __builtin_popcount(x & 0x7FFFFFFFu) + 33 * (int)(x >> 31);
}

/* ------------------------------------------------------------- AD cost
 * The reference fills, per 160-wide block and row, a flat shared array
 * [ left row | right row ] of sm_cols pixels each starting at block_start -
 * padding, and indexes it with tx + padding +- (d - zero_disp).  When
 * padding = zero_disp - 1 the index leaves its half by one pixel at d = 0
 * (SURVEY Q4); the flat array reproduces that. */
static int ad_padding(int num_disp, int zero_disp)
{
    int positive = num_disp - zero_disp; /* d_ci_adcensus.cu:57-58 */
    return positive > zero_disp ? positive : zero_disp - 1;
}

static inline int ad_sum(const uint8_t *a, const uint8_t *b)
{
    return abs((int)a[0] - (int)b[0]) + abs((int)a[1] - (int)b[1]) + abs((int)a[2] - (int)b[2]);
}

/* Integer |dB|+|dG|+|dR| for both views at (gx,gy,d), reference indexing. */
static inline void ad_pair(const uint8_t *row_l, const uint8_t *row_r, int gx, int d,
                           int num_disp, int zero_disp, int num_cols, int elem_sz,
                           int *sum_l, int *sum_r)
{
    int pad = ad_padding(num_disp, zero_disp);
    int sm_cols = REF_BLOCK_W + 2 * pad;
    int tx = gx % REF_BLOCK_W, bs = gx - tx;
    int idx_r = tx + pad + (d - zero_disp); /* into the right half */
    int idx_l = tx + pad - (d - zero_disp); /* into the left half  */
    const uint8_t *pr, *pl;
    /* flat position = half offset + index; position p maps to image
     * (p < sm_cols ? left : right) at column clamp(bs - pad + p % sm_cols) */
    int fr = sm_cols + idx_r, fl = idx_l;
    {
        const uint8_t *row = fr < sm_cols ? row_l : row_r;
        int col = clampi(bs - pad + (fr < sm_cols ? fr : fr - sm_cols), 0, num_cols - 1);
        pr = row + (size_t)col * elem_sz;
    }
    {
        const uint8_t *row = fl < sm_cols ? row_l : row_r;
        int col = clampi(bs - pad + (fl < sm_cols ? fl : fl - sm_cols), 0, num_cols - 1);
        pl = row + (size_t)col * elem_sz;
    }
    *sum_l = ad_sum(row_l + (size_t)gx * elem_sz, pr);
    *sum_r = ad_sum(row_r + (size_t)gx * elem_sz, pl);
}

void orc_ad_cost(const uint8_t *img_l, const uint8_t *img_r, float *cost_l, float *cost_r,
                 int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz)
{
    size_t plane = (size_t)num_rows * num_cols;
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < num_rows; ++gy) {
        const uint8_t *rl = img_l + (size_t)gy * num_cols * elem_sz;
        const uint8_t *rr = img_r + (size_t)gy * num_cols * elem_sz;
        for (int d = 0; d < num_disp; ++d)
            for (int gx = 0; gx < num_cols; ++gx) {
                int sl, sr;
                ad_pair(rl, rr, gx, d, num_disp, zero_disp, num_cols, elem_sz, &sl, &sr);
                /* PTX: cvt.rn.f32.u32 of the integer sum, one mul.f32 by 0x3EAAAAAB */
                cost_l[d * plane + (size_t)gy * num_cols + gx] = (float)sl * 0.33333333333f;
                cost_r[d * plane + (size_t)gy * num_cols + gx] = (float)sr * 0.33333333333f;
            }
    }
}

/* --------------------------------------------------------- census cost */
static inline void census_pair(const uint64_t *row_l, const uint64_t *row_r, int gx, int d,
                               int num_disp, int zero_disp, int num_cols, int *ham_l, int *ham_r)
{
    int sm_cols = REF_BLOCK_W + num_disp - 1; /* d_ci_adcensus.cu:117-120 */
    int pad_l = zero_disp - 1, pad_r = num_disp - zero_disp;
    int tx = gx % REF_BLOCK_W, bs = gx - tx;
    /* flat array: [0,sm_cols) = CL(clamp(bs - pad_r + i)); [sm_cols,2sm_cols) = CR(clamp(bs - pad_l + i)) */
    int f[4] = { tx + pad_r,                            /* l1 */
                 sm_cols + tx + pad_l,                  /* r1 */
                 sm_cols + tx + pad_l + (d - zero_disp),/* r2 */
                 tx + pad_r - (d - zero_disp) };        /* l2 */
    uint64_t v[4];
    for (int k = 0; k < 4; ++k) {
        int p = f[k];
        if (p < sm_cols) v[k] = row_l[clampi(bs - pad_r + p, 0, num_cols - 1)];
        else             v[k] = row_r[clampi(bs - pad_l + (p - sm_cols), 0, num_cols - 1)];
    }
    *ham_l = orc_hamdist(v[0], v[2]);
    *ham_r = orc_hamdist(v[1], v[3]);
}

void orc_census_cost(const uint64_t *census_l, const uint64_t *census_r,
                     float *cost_l, float *cost_r,
                     int num_disp, int zero_disp, int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < num_rows; ++gy) {
        const uint64_t *rl = census_l + (size_t)gy * num_cols, *rr = census_r + (size_t)gy * num_cols;
        for (int d = 0; d < num_disp; ++d)
            for (int gx = 0; gx < num_cols; ++gx) {
                int hl, hr;
                census_pair(rl, rr, gx, d, num_disp, zero_disp, num_cols, &hl, &hr);
                cost_l[d * plane + (size_t)gy * num_cols + gx] = (float)hl;
                cost_r[d * plane + (size_t)gy * num_cols + gx] = (float)hr;
            }
    }
}

/* ------------------------------------------------------------ combine */
static float one_minus_exp(float cost, float inv_coeff)
{
    /* PTX of ci_adcensus_kernel: neg, mul by inv, mul by 0x3FB8AA3B,
     * ex2.approx.f32, 1 - x.  exp2f stands in for ex2.approx on the CPU. */
    float t = (-cost) * inv_coeff;
    t = t * 1.4426950408889634f;
    return 1.0f - exp2f(t);
}

void orc_exp_luts(float ad_coeff, float census_coeff, float *lut_ad, float *lut_cen)
{
    float inv_ad = (float)(1.0 / ad_coeff), inv_cen = (float)(1.0 / census_coeff); /* d_ci_adcensus.cu:160 */
    for (int s = 0; s < ORC_AD_LUT_SIZE; ++s)
        lut_ad[s] = one_minus_exp((float)s * 0.33333333333f, inv_ad);
    for (int h = 0; h < ORC_CEN_LUT_SIZE; ++h)
        lut_cen[h] = one_minus_exp((float)h, inv_cen);
}

void orc_ci_adcensus(const uint8_t *img_l, const uint8_t *img_r,
                     float *cost_l, float *cost_r,
                     const float *lut_ad, const float *lut_cen,
                     float ad_coeff, float census_coeff, int num_disp, int zero_disp,
                     int num_rows, int num_cols, int elem_sz)
{
    size_t plane = (size_t)num_rows * num_cols;
    float la[ORC_AD_LUT_SIZE], lc[ORC_CEN_LUT_SIZE];
    if (!lut_ad || !lut_cen) {
        orc_exp_luts(ad_coeff, census_coeff, la, lc);
        lut_ad = la; lut_cen = lc;
    }
    uint8_t *gl = (uint8_t *)malloc(plane), *gr = (uint8_t *)malloc(plane);
    uint64_t *cl = (uint64_t *)malloc(plane * 8), *cr = (uint64_t *)malloc(plane * 8);
    orc_gray(img_l, gl, num_rows, num_cols, elem_sz);
    orc_gray(img_r, gr, num_rows, num_cols, elem_sz);
    orc_census(gl, cl, num_rows, num_cols);
    orc_census(gr, cr, num_rows, num_cols);
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < num_rows; ++gy) {
        const uint8_t *rl = img_l + (size_t)gy * num_cols * elem_sz;
        const uint8_t *rr = img_r + (size_t)gy * num_cols * elem_sz;
        const uint64_t *kl = cl + (size_t)gy * num_cols, *kr = cr + (size_t)gy * num_cols;
        for (int d = 0; d < num_disp; ++d)
            for (int gx = 0; gx < num_cols; ++gx) {
                int sl, sr, hl, hr;
                ad_pair(rl, rr, gx, d, num_disp, zero_disp, num_cols, elem_sz, &sl, &sr);
                census_pair(kl, kr, gx, d, num_disp, zero_disp, num_cols, &hl, &hr);
                cost_l[d * plane + (size_t)gy * num_cols + gx] = lut_ad[sl] + lut_cen[hl];
                cost_r[d * plane + (size_t)gy * num_cols + gx] = lut_ad[sr] + lut_cen[hr];
            }
    }
    free(gl); free(gr); free(cl); free(cr);
}

/* ----------------------------------------------------------- cross arms */
static inline int mad3(const uint8_t *a, const uint8_t *b)
{
    int d0 = abs((int)a[0] - (int)b[0]), d1 = abs((int)a[1] - (int)b[1]), d2 = abs((int)a[2] - (int)b[2]);
    return imax(imax(d0, d1), d2);
}

static int arm_walk(const uint8_t *img, int x, int y, int dx, int dy, float ucd, float lcd,
                    int usd, int lsd, int num_rows, int num_cols, int elem_sz)
{
    /* d_ca_cross.cu:41-69: the arm takes the value of the step *before* the
     * colour test, so it ends on the first failing pixel (or usd / border). */
    const uint8_t *a = img + ((size_t)y * num_cols + x) * elem_sz;
    const uint8_t *p = a;
    int arm = 0;
    for (int s = 1; s <= usd; ++s) {
        int cx = x + dx * s, cy = y + dy * s;
        if (cx < 0 || cx > num_cols - 1 || cy < 0 || cy > num_rows - 1) break;
        arm = s;
        const uint8_t *c = img + ((size_t)cy * num_cols + cx) * elem_sz;
        int ac = mad3(c, a), cp = mad3(c, p);
        if (s > lsd) {
            if ((float)ac > ucd) break;
        } else {
            if ((float)ac > lcd || (float)cp > lcd) break;
        }
        p = c;
    }
    return arm;
}

void orc_cross_arms(const uint8_t *img, uint8_t *arms, float ucd, float lcd, int usd, int lsd,
                    int num_rows, int num_cols, int elem_sz)
{
    size_t plane = (size_t)num_rows * num_cols;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < num_rows; ++y)
        for (int x = 0; x < num_cols; ++x) {
            size_t i = (size_t)y * num_cols + x;
            arms[0 * plane + i] = (uint8_t)arm_walk(img, x, y, 0, -1, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
            arms[1 * plane + i] = (uint8_t)arm_walk(img, x, y, 0, +1, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
            arms[2 * plane + i] = (uint8_t)arm_walk(img, x, y, -1, 0, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
            arms[3 * plane + i] = (uint8_t)arm_walk(img, x, y, +1, 0, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
        }
}

/* ---------------------------------------------------------- aggregation */
void orc_ca_pass(const float *in, float *out, const uint8_t *arms, int dir,
                 int num_disp, int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
    const uint8_t *arm_a = arms + (dir ? 0 : 2) * plane; /* UP or LEFT   */
    const uint8_t *arm_b = arms + (dir ? 1 : 3) * plane; /* DOWN or RIGHT */
    ptrdiff_t step = dir ? num_cols : 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < num_disp; ++d)
        for (int y = 0; y < num_rows; ++y) {
            const float *src = in + d * plane;
            float *dst = out + d * plane;
            for (int x = 0; x < num_cols; ++x) {
                size_t i = (size_t)y * num_cols + x;
                /* d_ca_cross_sum.cu:282-290: window [p - a, p + b), ascending, from 0 */
                float asum = 0;
                for (int k = -(int)arm_a[i]; k < (int)arm_b[i]; ++k)
                    asum = asum + src[(ptrdiff_t)i + k * step];
                dst[i] = asum;
            }
        }
}

void orc_ca_aggregate(float *cost, const uint8_t *arms, int num_disp, int num_rows, int num_cols)
{
    size_t n = (size_t)num_disp * num_rows * num_cols;
    float *tmp = (float *)malloc(n * sizeof(float));
    /* d_ca_cross.cu:255-271: H, V, V, H; the result lands back in `cost` */
    orc_ca_pass(cost, tmp, arms, 0, num_disp, num_rows, num_cols);
    orc_ca_pass(tmp, cost, arms, 1, num_disp, num_rows, num_cols);
    orc_ca_pass(cost, tmp, arms, 1, num_disp, num_rows, num_cols);
    orc_ca_pass(tmp, cost, arms, 0, num_disp, num_rows, num_cols);
    free(tmp);
}

/* ------------------------------------------------------------------ WTA */
void orc_wta(const float *cost, float *disp, int num_disp, int zero_disp, int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < plane; ++i) {
        float lowest = FLT_MAX, lowest_d = 0;
        for (int d = 0; d < num_disp; ++d) {
            float c = cost[d * plane + i];
            if (lowest > c) { lowest = c; lowest_d = (float)d; } /* strict: first minimum */
        }
        disp[i] = lowest_d - (float)zero_disp;
    }
}

/* ------------------------------------------------- scanline optimisation */
/* PARITY UNPINNED: the reference's d_dc_hslo.cu is an unfinished stub (its kernels compute penalty maps
 * and nothing else, d_dc_hslo.cu:9-95; the only call site is commented out, image_io.cpp:307-316), so
 * there is no reference output to match.  This function is the SPECIFICATION of the stage as built
 * here: the four-direction scanline optimisation of Mei et al. 2011 ("On Building an Accurate Stereo
 * Matching System on Graphics Hardware", named in the reference's README:4) with everything the stub does
 * define taken from it: the call shape (d_dc_hslo.cu:97-101), the constants T = 15, H1 = 1.0, H2 = 3.0
 * (image_io.cpp:311-313), the colour measure (mean of B, G, R; d_dc_hslo.cu:55-70), the three penalty tiers
 * H, H/4, H/10 and their strict comparisons (d_dc_hslo.cu:73-93, 124-127).
 *
 *   C_r(p, d) = C(p, d) + min( C_r(p-r, d), C_r(p-r, d-1) + P1, C_r(p-r, d+1) + P1, m + P2 ) - m,
 *               m = min_k C_r(p-r, k);     C_r = C on the first pixel of a scanline
 *   D1 = |g_own(p) - g_own(p-r)|,  D2 = |g_other(q) - g_other(q-r)|,  g = (B + G + R) / 3,
 *               q = p shifted by +(d - zd) columns for the left view, -(d - zd) for the right (clamped)
 *   (P1, P2) = (H1, H2) if D1 < T and D2 < T;  (H1/4, H2/4) if exactly one of them is < T and the other > T;
 *              (H1/10, H2/10) otherwise
 *   C2 = (((C_lr + C_rl) + C_tb) + C_bt) * 0.25,  r = (+1,0), (-1,0), (0,+1), (0,-1) in that order
 * followed by winner-takes-all (first minimum, as orc_wta).  fp32 throughout, one rounding per operation
 * in exactly the order written.  cost: [d][y][x]; `cost_out` (may be NULL) receives C2. */
static inline float so_gray(const uint8_t *img, int x, int y, int num_cols, int elem_sz)
{
    const uint8_t *p = img + ((size_t)y * num_cols + x) * elem_sz;
    return (float)((int)p[0] + (int)p[1] + (int)p[2]) / 3.0f;
}

static void so_direction(const float *cost, float *acc, int first, const uint8_t *img_own, const uint8_t *img_other,
                         int view, int dx, int dy, float T, float H1, float H2, int num_disp, int zero_disp,
                         int num_rows, int num_cols, int elem_sz)
{
    const size_t plane = (size_t)num_rows * num_cols;
    const int nlines = dx ? num_rows : num_cols, len = dx ? num_cols : num_rows;
    const float P1t[3] = {H1, H1 / 4.0f, H1 / 10.0f}, P2t[3] = {H2, H2 / 4.0f, H2 / 10.0f};
#pragma omp parallel for schedule(static)
    for (int ln = 0; ln < nlines; ++ln) {
        float *prev = (float *)malloc(sizeof(float) * num_disp), *cur = (float *)malloc(sizeof(float) * num_disp);
        for (int t = 0; t < len; ++t) {
            const int k = (dx > 0 || dy > 0) ? t : len - 1 - t;
            const int x = dx ? k : ln, y = dx ? ln : k;
            const size_t i = (size_t)y * num_cols + x;
            if (t == 0) {
                for (int d = 0; d < num_disp; ++d) cur[d] = cost[d * plane + i];
            } else {
                const int px = x - dx, py = y - dy;
                float m = prev[0];
                for (int d = 1; d < num_disp; ++d) m = prev[d] < m ? prev[d] : m;
                const float D1 = fabsf(so_gray(img_own, x, y, num_cols, elem_sz) - so_gray(img_own, px, py, num_cols, elem_sz));
                for (int d = 0; d < num_disp; ++d) {
                    const int sh = view == 0 ? (d - zero_disp) : -(d - zero_disp);
                    const int qx = x + sh < 0 ? 0 : (x + sh > num_cols - 1 ? num_cols - 1 : x + sh);
                    const int qpx = px + sh < 0 ? 0 : (px + sh > num_cols - 1 ? num_cols - 1 : px + sh);
                    const float D2 = fabsf(so_gray(img_other, qx, y, num_cols, elem_sz) - so_gray(img_other, qpx, py, num_cols, elem_sz));
                    int tier;
                    if (D1 < T && D2 < T) tier = 0;
                    else if ((D1 < T && D2 > T) || (D1 > T && D2 < T)) tier = 1;
                    else tier = 2;
                    const float P1 = P1t[tier], P2 = P2t[tier];
                    float best = prev[d];
                    if (d > 0) { const float v = prev[d - 1] + P1; best = v < best ? v : best; }
                    if (d < num_disp - 1) { const float v = prev[d + 1] + P1; best = v < best ? v : best; }
                    { const float v = m + P2; best = v < best ? v : best; }
                    cur[d] = (cost[d * plane + i] + best) - m;
                }
            }
            for (int d = 0; d < num_disp; ++d) {
                if (first) acc[d * plane + i] = cur[d];
                else acc[d * plane + i] = acc[d * plane + i] + cur[d];
            }
            float *tmp = prev; prev = cur; cur = tmp;
        }
        free(prev); free(cur);
    }
}

void orc_so(const float *cost, float *cost_out, float *disp, const uint8_t *img_own, const uint8_t *img_other,
            int view, float T, float H1, float H2, int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz)
{
    const size_t n = (size_t)num_rows * num_cols * num_disp;
    float *acc = (float *)malloc(sizeof(float) * n);
    so_direction(cost, acc, 1, img_own, img_other, view, +1, 0, T, H1, H2, num_disp, zero_disp, num_rows, num_cols, elem_sz);
    so_direction(cost, acc, 0, img_own, img_other, view, -1, 0, T, H1, H2, num_disp, zero_disp, num_rows, num_cols, elem_sz);
    so_direction(cost, acc, 0, img_own, img_other, view, 0, +1, T, H1, H2, num_disp, zero_disp, num_rows, num_cols, elem_sz);
    so_direction(cost, acc, 0, img_own, img_other, view, 0, -1, T, H1, H2, num_disp, zero_disp, num_rows, num_cols, elem_sz);
    for (size_t i = 0; i < n; ++i) acc[i] = acc[i] * 0.25f;
    if (disp) orc_wta(acc, disp, num_disp, zero_disp, num_rows, num_cols);
    if (cost_out) memcpy(cost_out, acc, sizeof(float) * n);
    free(acc);
}

/* ------------------------------------------------------------------ DCC */
void orc_dcc(uint8_t *outliers_l, uint8_t *outliers_r, const float *disp_l, const float *disp_r,
             int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
    uint8_t *dis_l = (uint8_t *)malloc(plane), *dis_r = (uint8_t *)malloc(plane);
    memset(dis_l, 1, plane); memset(dis_r, 1, plane);
    memset(outliers_l, 0, plane); memset(outliers_r, 0, plane);
    for (int y = 0; y < num_rows; ++y)
        for (int x = 0; x < num_cols; ++x) {
            size_t i = (size_t)y * num_cols + x;
            /* dr_dcc_kernel, thresh = 1.0 */
            float d = disp_l[i];
            int c = clampi(x + (int)d, 0, num_cols - 1);
            if (fabsf(d - disp_r[(size_t)y * num_cols + c]) > 1.0f) outliers_l[i] = 1;
            d = disp_r[i];
            c = clampi(x - (int)d, 0, num_cols - 1);
            if (fabsf(d - disp_l[(size_t)y * num_cols + c]) > 1.0f) outliers_r[i] = 1;
            /* dr_ddc_kernel: every writer stores 0, order-free */
            c = clampi(x + (int)disp_l[i], 0, num_cols - 1);
            dis_r[(size_t)y * num_cols + c] = 0;
            c = clampi(x - (int)disp_r[i], 0, num_cols - 1);
            dis_l[(size_t)y * num_cols + c] = 0;
        }
    for (size_t i = 0; i < plane; ++i) { /* dr_merge_errors_kernel */
        if (outliers_l[i] == 1 && dis_l[i] == 1) outliers_l[i] = 2;
        if (outliers_r[i] == 1 && dis_r[i] == 1) outliers_r[i] = 2;
    }
    free(dis_l); free(dis_r);
}

/* ------------------------------------------------------------------ IRV */
static void irv_votes(const float *disp, const uint8_t *outliers, const uint8_t *arms,
                      int *max_disp, int *reliable, int num_rows, int num_cols,
                      int num_disp, int zero_disp, int usd)
{
    size_t plane = (size_t)num_rows * num_cols;
    int nbins = imax(num_disp, 65);
#pragma omp parallel
    {
        int *hist = (int *)malloc(sizeof(int) * nbins);
#pragma omp for schedule(dynamic, 4)
        for (int gy = 0; gy < num_rows; ++gy)
            for (int gx = 0; gx < num_cols; ++gx) {
                size_t i = (size_t)gy * num_cols + gx;
                if (outliers[i] == 0) continue;
                int cu = arms[0 * plane + i], cd = arms[1 * plane + i];
                if (cu > usd) cu = usd; /* d_dr_irv.cu:181-182 */
                int max_bin = 0, max_d = (int)disp[i], total = 0;
                memset(hist, 0, sizeof(int) * nbins);
                for (int y = -cu; y <= cd; ++y) {
                    int yy = clampi(gy + y, 0, num_rows - 1);
                    size_t r = (size_t)(gy + y) * num_cols + gx; /* arms read unclamped (:190-191) */
                    int cl = arms[2 * plane + r], cr = arms[3 * plane + r];
                    for (int x = -cl; x <= cr; ++x) {
                        size_t s = (size_t)yy * num_cols + clampi(gx + x, 0, num_cols - 1);
                        if (outliers[s] == 0) {
                            hist[(int)disp[s] + zero_disp]++;
                            total++;
                        }
                    }
                }
                for (int b = 0; b < nbins; ++b)
                    if (max_bin < hist[b]) { max_bin = hist[b]; max_d = b - zero_disp; }
                max_disp[i] = max_d;
                reliable[i] = total;
            }
        free(hist);
    }
}

static void irv_apply(float *disp, uint8_t *outliers, const int *max_disp, int *reliable,
                      int thresh_s, float thresh_h, size_t plane, int zero_disp)
{
    for (size_t i = 0; i < plane; ++i) {
        if (outliers[i] == 0) continue;
        int tr = reliable[i], md = max_disp[i];
        if (tr > thresh_s && (float)(md + zero_disp) / (float)tr > thresh_h) { /* d_dr_irv.cu:36 */
            outliers[i] = 0;
            reliable[i] += 1;
            disp[i] = (float)md;
        }
    }
}

void orc_irv(float *disp, uint8_t *outliers, const uint8_t *arms, int thresh_s, float thresh_h,
             int num_rows, int num_cols, int num_disp, int zero_disp, int usd, int iterations,
             int host_variant)
{
    size_t plane = (size_t)num_rows * num_cols;
    int *max_disp = (int *)calloc(plane, sizeof(int)), *reliable = (int *)calloc(plane, sizeof(int));
    if (host_variant) {
        irv_votes(disp, outliers, arms, max_disp, reliable, num_rows, num_cols, num_disp, zero_disp, usd);
        for (int it = 0; it < iterations; ++it)
            irv_apply(disp, outliers, max_disp, reliable, thresh_s, thresh_h, plane, zero_disp);
    } else {
        for (int it = 0; it < iterations; ++it) {
            irv_votes(disp, outliers, arms, max_disp, reliable, num_rows, num_cols, num_disp, zero_disp, usd);
            irv_apply(disp, outliers, max_disp, reliable, thresh_s, thresh_h, plane, zero_disp);
        }
    }
    free(max_disp); free(reliable);
}

/* ------------------------------------------------------- filter weights */
#define REF_PI 3.14159265359f /* d_filter_gaussian.cu:7, d_filter_bilateral.cu:8 */

void orc_gaussian_kernel(float *kernel, int radius, float sigma)
{
    /* gaussian2D, d_filter_gaussian.cu:237-242: pow() promotes to double */
    int w = radius * 2 + 1;
    for (int y = -radius; y <= radius; ++y)
        for (int x = -radius; x <= radius; ++x) {
            float variance = (float)pow((double)sigma, 2.0);
            float exponent = (float)(-(pow((double)(float)x, 2.0) + pow((double)(float)y, 2.0)) /
                                     (double)(2 * variance));
            kernel[(x + radius) + (y + radius) * w] = expf(exponent) / (2 * REF_PI * variance);
        }
}

void orc_gaussian_1d(float *kernel, int size, float sigma)
{
    /* gaussian1D_host, d_filter_bilateral.cu:26-39 (float overloads of exp/sqrt) */
    for (int i = 0; i < size; ++i) {
        float variance = (float)pow((double)sigma, 2.0);
        float power = (float)pow((double)(float)i, 2.0);
        float exponent = -power / (2 * variance);
        kernel[i] = expf(exponent) / sqrtf(2 * REF_PI * variance);
    }
}

/* ------------------------------------------------------------ bilateral */
void orc_bilateral(float *img, int radius, float sigma_color, float sigma_spatial,
                   int num_rows, int num_cols, int num_disp)
{
    size_t plane = (size_t)num_rows * num_cols;
    int w = 2 * radius + 1;
    float *spatial = (float *)malloc(sizeof(float) * w * w);
    float *color = (float *)malloc(sizeof(float) * imax(num_disp, 1));
    float *out = (float *)malloc(sizeof(float) * plane);
    orc_gaussian_kernel(spatial, radius, sigma_spatial);
    orc_gaussian_1d(color, num_disp, sigma_color);
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < num_rows; ++gy)
        for (int gx = 0; gx < num_cols; ++gx) {
            float va = img[(size_t)gy * num_cols + gx];
            float norm = 0.0f, res = 0.0f;
            for (int y = -radius; y <= radius; ++y) {
                int sy = clampi(gy + y, 0, num_rows - 1);
                for (int x = -radius; x <= radius; ++x) {
                    int sx = clampi(gx + x, 0, num_cols - 1);
                    float vs = img[(size_t)sy * num_cols + sx];
                    float weight = spatial[(x + radius) + (y + radius) * w] * color[(int)fabsf(va - vs)];
                    norm = norm + weight;
                    res = fmaf(vs, weight, res); /* PTX: fma.rn.f32 */
                }
            }
            out[(size_t)gy * num_cols + gx] = res / norm;
        }
    memcpy(img, out, sizeof(float) * plane);
    free(spatial); free(color); free(out);
}

/* ------------------------------------------------------ occlusion / mask */
void orc_occl(uint8_t *occl_l, uint8_t *occl_r, const float *disp_l, const float *disp_r,
              int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
    memset(occl_l, 0, plane); memset(occl_r, 0, plane);
    for (int y = 0; y < num_rows; ++y)
        for (int x = 0; x < num_cols; ++x) {
            size_t i = (size_t)y * num_cols + x;
            int sd = (int)(disp_l[i] * (float)1); /* d_dibr_occl.cu:124, dir = +1 */
            occl_r[(size_t)y * num_cols + clampi(x + sd, 0, num_cols - 1)] = 1;
            sd = (int)(disp_r[i] * (float)-1);
            occl_l[(size_t)y * num_cols + clampi(x + sd, 0, num_cols - 1)] = 1;
        }
}

void orc_bleed(uint8_t *img, int radius, int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
    int kernel_sz = (2 * radius + 1) * (2 * radius + 1);
    uint8_t *out = (uint8_t *)malloc(plane);
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < num_rows; ++ty)
        for (int tx = 0; tx < num_cols; ++tx) {
            int count = 0;
            for (int y = -radius; y <= radius; ++y)
                for (int x = -radius; x <= radius; ++x) {
                    int sx = tx + x, sy = ty + y; /* d_filter.cu:121-127 */
                    if (sx < 0) sx = -sx;
                    if (sy < 0) sy = -sy;
                    if (sx > num_cols - 1) sx = num_cols - 1 - x;
                    if (sy > num_rows - 1) sy = num_rows - 1 - y;
                    if (img[(size_t)sy * num_cols + sx] > 0) count++;
                }
            out[(size_t)ty * num_cols + tx] =
                ((double)count > (kernel_sz - 1) * 0.30) ? 1 : img[(size_t)ty * num_cols + tx];
        }
    memcpy(img, out, plane);
    free(out);
}

void orc_occl_to_mask(float *mask, const uint8_t *occl, int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
    for (size_t i = 0; i < plane; ++i) mask[i] = occl[i] == 1 ? 1.0f : 0.0f;
}

void orc_gaussian_dilate(float *img, int radius, float sigma, int num_rows, int num_cols)
{
    size_t plane = (size_t)num_rows * num_cols;
    int w = 2 * radius + 1;
    float *kernel = (float *)malloc(sizeof(float) * w * w);
    float *out = (float *)malloc(sizeof(float) * plane);
    orc_gaussian_kernel(kernel, radius, sigma);
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < num_rows; ++gy)
        for (int gx = 0; gx < num_cols; ++gx) {
            float va = img[(size_t)gy * num_cols + gx];
            float res = 0.0f, norm = 0.0f;
            for (int y = -radius; y <= radius; ++y) {
                int sy = clampi(gy + y, 0, num_rows - 1);
                for (int x = -radius; x <= radius; ++x) {
                    int sx = clampi(gx + x, 0, num_cols - 1);
                    float weight = kernel[(x + radius) + (y + radius) * w];
                    norm = norm + weight;
                    res = fmaf(img[(size_t)sy * num_cols + sx], weight, res); /* PTX: fma.rn.f32 */
                }
            }
            float q = res / norm;
            out[(size_t)gy * num_cols + gx] = (va < q) ? q : va; /* d_filter_gaussian.cu:84-87 */
        }
    memcpy(img, out, sizeof(float) * plane);
    free(kernel); free(out);
}

/* ----------------------------------------------------------------- DIBR */
void orc_bwarp(uint8_t *out, const uint8_t *in, const float *mask, const float *disp,
               float shift, int num_rows, int num_cols, int elem_sz)
{
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < num_rows; ++ty)
        for (int tx = 0; tx < num_cols; ++tx) {
            size_t i = (size_t)ty * num_cols + tx;
            /* PTX: fma.rn(shift, disp, (float)tx); max 0; min W-1; cvt.rzi */
            float fx = fmaf(shift, disp[i], (float)tx);
            fx = fmaxf(fx, 0.0f);
            fx = fminf(fx, (float)(num_cols - 1));
            int sx = (int)fx;
            /* bilinear at integral coordinates degenerates to a plain fetch (Q23) */
            const uint8_t *s = in + ((size_t)ty * num_cols + sx) * elem_sz;
            for (int c = 0; c < 3; ++c)
                out[i * elem_sz + c] = (uint8_t)(unsigned)((float)s[c] * mask[i]);
        }
}

void orc_merge_ab(uint8_t *img_b, const uint8_t *img_a, const float *mask_a,
                  int num_rows, int num_cols, int elem_sz)
{
    size_t plane = (size_t)num_rows * num_cols;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < plane; ++i) {
        float m = mask_a[i], im = 1.0f - m;
        for (int c = 0; c < 3; ++c) {
            unsigned a = (unsigned)(m * (float)img_a[i * elem_sz + c]);
            unsigned b = (unsigned)(im * (float)img_b[i * elem_sz + c]);
            img_b[i * elem_sz + c] = (uint8_t)((uint8_t)b + (uint8_t)a);
        }
    }
}

/* dibr_forward_warp_kernel (d_dibr_fwarp.cu:9-25).  The reference's scatter has no ordering between sources
 * that land on one destination; this statement fixes it: the row is scanned right to left, so the
 * LOWEST source column wins (the outcome the reference's kernel is observed to produce on a B200 for ~98 % of the
 * colliding destinations).  `hits` (optional) receives the number of sources per destination, so that a test
 * can tell the destinations on which the reference itself is well defined (hits <= 1).  PARITY UNPINNED on the
 * others. */
void orc_fwarp(uint8_t *out, const uint8_t *in, const float *disp, float shift, int *hits,
               int num_rows, int num_cols, int elem_sz)
{
    memset(out, 0, (size_t)num_rows * num_cols * elem_sz);
    if (hits) memset(hits, 0, (size_t)num_rows * num_cols * sizeof(int));
    for (int ty = 0; ty < num_rows; ++ty)
        for (int tx = num_cols - 1; tx >= 0; --tx) {
            size_t i = (size_t)ty * num_cols + tx;
            int sd = (int)(disp[i] * shift);
            int sx = tx + sd;
            sx = sx < 0 ? 0 : (sx > num_cols - 1 ? num_cols - 1 : sx);
            size_t o = (size_t)ty * num_cols + sx;
            for (int c = 0; c < 3; ++c) out[o * elem_sz + c] = in[i * elem_sz + c];
            if (hits) hits[o] += 1;
        }
}

/* d_dibr_dfm (d_dibr_fwarp.cu:27-95): both forward warps, then mux_merge_AB(out_l, out_r, mask) with the mask of a
 * zeroed occlusion map (0 everywhere: :55-66), i.e. the left warp survives. */
void orc_dibr_dfm(uint8_t *out, const uint8_t *img_l, const uint8_t *img_r, const float *disp_l, const float *disp_r,
                  float shift, int num_rows, int num_cols, int elem_sz)
{
    size_t plane = (size_t)num_rows * num_cols;
    uint8_t *out_r = (uint8_t *)malloc(plane * elem_sz);
    float *mask = (float *)calloc(plane, sizeof(float));
    orc_fwarp(out, img_l, disp_l, shift, NULL, num_rows, num_cols, elem_sz);
    orc_fwarp(out_r, img_r, disp_r, (float)(1.0 - shift), NULL, num_rows, num_cols, elem_sz);
    orc_merge_ab(out, out_r, mask, num_rows, num_cols, elem_sz);
    free(out_r);
    free(mask);
}

void orc_dbm(uint8_t *out, const uint8_t *img_l, const uint8_t *img_r,
             const float *disp_l, const float *disp_r, const float *mask_l, const float *mask_r,
             float shift, int gauss_radius, float gauss_sigma,
             int num_rows, int num_cols, int elem_sz)
{
    size_t plane = (size_t)num_rows * num_cols;
    uint8_t *out_r = (uint8_t *)calloc(plane * elem_sz, 1);
    float *tmask = (float *)malloc(sizeof(float) * plane);
    memset(out, 0, plane * elem_sz);
    orc_bwarp(out, img_l, mask_r, disp_r, -shift, num_rows, num_cols, elem_sz);
    orc_bwarp(out_r, img_r, mask_l, disp_l, (float)(1.0 - shift), num_rows, num_cols, elem_sz);
    for (size_t i = 0; i < plane; ++i) tmask[i] = 1.0f - mask_r[i]; /* d_op.cu:7-16 */
    orc_gaussian_dilate(tmask, gauss_radius, gauss_sigma, num_rows, num_cols);
    orc_merge_ab(out, out_r, tmask, num_rows, num_cols, elem_sz);
    free(out_r); free(tmask);
}

/* ------------------------------------------------------------ interlace */
static inline uint8_t bilinear_u8(const uint8_t *data, int elem_sz, int off, float cx, float cy,
                                  int width, int height)
{
    /* fast_bilinear_interp, d_mux_multiview.cu:10-36, with the PTX's fma placement */
    int x0 = (int)floorf(cx), y0 = (int)floorf(cy);
    int x1 = imin(x0 + 1, width - 1), y1 = imin(y0 + 1, height - 1);
    float wx = cx - (float)x0, wy = cy - (float)y0;
    float v00 = (float)data[((size_t)y0 * width + x0) * elem_sz + off];
    float v01 = (float)data[((size_t)y0 * width + x1) * elem_sz + off];
    float v10 = (float)data[((size_t)y1 * width + x0) * elem_sz + off];
    float v11 = (float)data[((size_t)y1 * width + x1) * elem_sz + off];
    float top = fmaf(1.0f - wx, v00, wx * v01);
    float bot = fmaf(1.0f - wx, v10, wx * v11);
    return (uint8_t)(unsigned)fmaf(1.0f - wy, top, wy * bot);
}

void orc_mux_multiview(const uint8_t *const *views, uint8_t *out, int num_views, float angle,
                       int num_rows_in, int num_cols_in, int num_rows_out, int num_cols_out,
                       int elem_sz, int kernel_variant)
{
#define MUX_PI 3.1415926535f /* d_mux_multiview.cu:8 */
    /* d_mux_multiview.cu:146: float*float, then double division, tan, divisions */
    float y_interval = (float)((double)(float)num_views / tan((double)(angle * MUX_PI) / 180.0) / (double)(float)elem_sz);
    float inv_y_interval = 1.0f / y_interval;
    int ri = (int)roundf(y_interval);
    float x_interval = (float)num_views;
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < num_rows_out; ++ty)
        for (int tx = 0; tx < num_cols_out; ++tx) {
            float x_samp = ((float)tx / (float)num_cols_out) * (float)num_cols_in;
            float y_samp = ((float)ty / (float)num_rows_out) * (float)num_rows_in;
            x_samp = (float)fmin(fmax((double)x_samp, 0.0), (double)(float)(num_cols_in - 1));
            y_samp = (float)fmin(fmax((double)y_samp, 0.0), (double)(float)(num_rows_in - 1));
            float y_view;
            if (kernel_variant == 2) {
                y_view = (float)(ty % ri) + 1.0f;
                y_view = y_view * x_interval;
                y_view = inv_y_interval * y_view; /* PTX: (yv*xi) then * inv */
            } else {
                /* mux_multiview_kernel :104-105: `+ 1.0` in double, then float; mul, div.rn */
                y_view = (float)((double)(ty % ri) + 1.0);
                y_view = y_view * x_interval / y_interval;
            }
            int x_view = (tx * 3 + (int)y_view) % ((int)x_interval);
            int r_view = x_view;
            if (r_view < 0) r_view += num_views;
            int g_view = r_view + 1, b_view = r_view + 2;
            if (g_view >= num_views) g_view -= num_views;
            if (b_view >= num_views) b_view -= num_views;
            size_t o = ((size_t)ty * num_cols_out + tx) * elem_sz;
            out[o + 0] = bilinear_u8(views[b_view], elem_sz, 0, x_samp, y_samp, num_cols_in, num_rows_in);
            out[o + 1] = bilinear_u8(views[g_view], elem_sz, 1, x_samp, y_samp, num_cols_in, num_rows_in);
            out[o + 2] = bilinear_u8(views[r_view], elem_sz, 2, x_samp, y_samp, num_cols_in, num_rows_in);
        }
}

/* -------------------------------------------------------- full pipeline */
void orc_costvol(const uint8_t *img_l, const uint8_t *img_r, float *disp_l, float *disp_r,
                 int num_rows, int num_cols, int elem_sz, int num_disp, int zero_disp,
                 float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                 const float *lut_ad, const float *lut_cen)
{
    size_t plane = (size_t)num_rows * num_cols;
    float *cost_l = (float *)malloc(sizeof(float) * plane * num_disp);
    float *cost_r = (float *)malloc(sizeof(float) * plane * num_disp);
    uint8_t *arms = (uint8_t *)malloc(4 * plane);
    orc_ci_adcensus(img_l, img_r, cost_l, cost_r, lut_ad, lut_cen, ad_coeff, census_coeff,
                    num_disp, zero_disp, num_rows, num_cols, elem_sz);
    orc_cross_arms(img_l, arms, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
    orc_ca_aggregate(cost_l, arms, num_disp, num_rows, num_cols);
    orc_wta(cost_l, disp_l, num_disp, zero_disp, num_rows, num_cols);
    orc_cross_arms(img_r, arms, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
    orc_ca_aggregate(cost_r, arms, num_disp, num_rows, num_cols);
    orc_wta(cost_r, disp_r, num_disp, zero_disp, num_rows, num_cols);
    free(cost_l); free(cost_r); free(arms);
}

void orc_adcensus_stm(const uint8_t *img_sbs, float *disp_l, float *disp_r, uint8_t *interlaced,
                      int num_rows, int num_cols_sbs, int num_cols,
                      int num_rows_out, int num_cols_out, int elem_sz,
                      int num_views, int angle, int num_disp, int zero_disp,
                      float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                      int thresh_s, float thresh_h,
                      const float *lut_ad, const float *lut_cen, const orc_taps_t *taps)
{
    size_t plane = (size_t)num_rows * num_cols, imgsz = plane * elem_sz;
    uint8_t *img_l = (uint8_t *)malloc(imgsz), *img_r = (uint8_t *)malloc(imgsz);
    orc_demux_sbs(img_sbs, img_l, img_r, num_rows, num_cols_sbs, num_cols, elem_sz);

    float *cost_l = (float *)malloc(sizeof(float) * plane * num_disp);
    float *cost_r = (float *)malloc(sizeof(float) * plane * num_disp);
    uint8_t *arms_l = (uint8_t *)malloc(4 * plane), *arms_r = (uint8_t *)malloc(4 * plane);
    orc_ci_adcensus(img_l, img_r, cost_l, cost_r, lut_ad, lut_cen, ad_coeff, census_coeff,
                    num_disp, zero_disp, num_rows, num_cols, elem_sz);
    orc_cross_arms(img_l, arms_l, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
    orc_ca_aggregate(cost_l, arms_l, num_disp, num_rows, num_cols);
    orc_cross_arms(img_r, arms_r, ucd, lcd, usd, lsd, num_rows, num_cols, elem_sz);
    orc_ca_aggregate(cost_r, arms_r, num_disp, num_rows, num_cols);
    orc_wta(cost_l, disp_l, num_disp, zero_disp, num_rows, num_cols);
    orc_wta(cost_r, disp_r, num_disp, zero_disp, num_rows, num_cols);
    if (taps) {
        if (taps->acost_l) memcpy(taps->acost_l, cost_l, sizeof(float) * plane * num_disp);
        if (taps->acost_r) memcpy(taps->acost_r, cost_r, sizeof(float) * plane * num_disp);
        if (taps->arms_l) memcpy(taps->arms_l, arms_l, 4 * plane);
        if (taps->arms_r) memcpy(taps->arms_r, arms_r, 4 * plane);
        if (taps->wta_l) memcpy(taps->wta_l, disp_l, sizeof(float) * plane);
        if (taps->wta_r) memcpy(taps->wta_r, disp_r, sizeof(float) * plane);
    }
    free(cost_l); free(cost_r);

    uint8_t *out_l = (uint8_t *)malloc(plane), *out_r = (uint8_t *)malloc(plane);
    orc_dcc(out_l, out_r, disp_l, disp_r, num_rows, num_cols);
    if (taps && taps->outliers_l) memcpy(taps->outliers_l, out_l, plane);
    if (taps && taps->outliers_r) memcpy(taps->outliers_r, out_r, plane);
    /* d_io.cu:147-151 */
    orc_irv(disp_l, out_l, arms_l, thresh_s, thresh_h, num_rows, num_cols, num_disp, zero_disp, usd, 5, 0);
    orc_irv(disp_r, out_r, arms_r, thresh_s, thresh_h, num_rows, num_cols, num_disp, zero_disp, usd, 5, 0);
    if (taps && taps->irv_l) memcpy(taps->irv_l, disp_l, sizeof(float) * plane);
    if (taps && taps->irv_r) memcpy(taps->irv_r, disp_r, sizeof(float) * plane);
    orc_bilateral(disp_l, 7, 5, 10, num_rows, num_cols, num_disp);
    orc_bilateral(disp_r, 7, 5, 10, num_rows, num_cols, num_disp);
    free(out_l); free(out_r); free(arms_l); free(arms_r);

    uint8_t *occl_l = (uint8_t *)malloc(plane), *occl_r = (uint8_t *)malloc(plane);
    float *mask_l = (float *)malloc(sizeof(float) * plane), *mask_r = (float *)malloc(sizeof(float) * plane);
    orc_occl(occl_l, occl_r, disp_l, disp_r, num_rows, num_cols);
    orc_bleed(occl_l, 1, num_rows, num_cols);
    orc_bleed(occl_r, 1, num_rows, num_cols);
    orc_occl_to_mask(mask_l, occl_l, num_rows, num_cols);
    orc_occl_to_mask(mask_r, occl_r, num_rows, num_cols);
    if (taps && taps->mask_l) memcpy(taps->mask_l, mask_l, sizeof(float) * plane);
    if (taps && taps->mask_r) memcpy(taps->mask_r, mask_r, sizeof(float) * plane);

    /* d_io.cu:178-191: views[0] = right, views[V-1] = left */
    uint8_t *views_mem = (uint8_t *)calloc(imgsz * (size_t)num_views, 1);
    const uint8_t **views = (const uint8_t **)malloc(sizeof(uint8_t *) * num_views);
    views[0] = img_r;
    views[num_views - 1] = img_l;
    for (int v = 1; v < num_views - 1; ++v) {
        float shift = (float)(1.0 - ((1.0 * (double)(float)v) / ((double)(float)num_views - 1.0)));
        uint8_t *dst = views_mem + (size_t)v * imgsz;
        orc_dbm(dst, img_l, img_r, disp_l, disp_r, mask_l, mask_r, shift, 10, 15.0f,
                num_rows, num_cols, elem_sz);
        views[v] = dst;
    }
    if (taps && taps->views) {
        memcpy(taps->views, img_r, imgsz);
        memcpy(taps->views + (size_t)(num_views - 1) * imgsz, img_l, imgsz);
        for (int v = 1; v < num_views - 1; ++v)
            memcpy(taps->views + (size_t)v * imgsz, views[v], imgsz);
    }
    orc_mux_multiview(views, interlaced, num_views, (float)angle, num_rows, num_cols,
                      num_rows_out, num_cols_out, elem_sz, 2);

    free(views); free(views_mem);
    free(occl_l); free(occl_r); free(mask_l); free(mask_r);
    free(img_l); free(img_r);
}

/* ------------------------------------------------ half-resolution variant */
/* alu_bilinear_interp / alu_bilinear_interp_f (d_alu.cu:17-71) with the fma placement nvcc gives them at
 * -O3 (PTX of d_alu.cu: top = fma(1-wx, v00, wx*v01), bot likewise, res = fma(1-wy, top, wy*bot)). */
static inline float orc_bilerp(float v00, float v01, float v10, float v11, float wx, float wy)
{
    float iwx = 1.0f - wx, iwy = 1.0f - wy;
    float top = fmaf(iwx, v00, wx * v01);
    float bot = fmaf(iwx, v10, wx * v11);
    return fmaf(iwy, top, wy * bot);
}

static inline float orc_sample_coord(int t, int out_n, int in_n)
{
    /* d_tx_scale.cu:19-20,41-42: fmin(fmax(((float) t / (float) out_n) * (float) in_n, 0), (float)(in_n - 1)) */
    float v = ((float)t / (float)out_n) * (float)in_n;
    v = v < 0.0f ? 0.0f : v;
    return v > (float)(in_n - 1) ? (float)(in_n - 1) : v;
}

/* tx_scale_bilinear_kernel (d_tx_scale.cu:30-52) */
void orc_scale_bilinear(const uint8_t *img_in, uint8_t *img_out, int in_rows, int in_cols, int out_rows, int out_cols,
                        int elem_sz)
{
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < out_rows; ++gy)
        for (int gx = 0; gx < out_cols; ++gx) {
            float xs = orc_sample_coord(gx, out_cols, in_cols), ys = orc_sample_coord(gy, out_rows, in_rows);
            int x0 = (int)floorf(xs), y0 = (int)floorf(ys);
            int x1 = imin(x0 + 1, in_cols - 1), y1 = imin(y0 + 1, in_rows - 1);
            float wx = xs - (float)x0, wy = ys - (float)y0;
            for (int c = 0; c < 3; ++c) {
                float v00 = img_in[((size_t)y0 * in_cols + x0) * elem_sz + c], v01 = img_in[((size_t)y0 * in_cols + x1) * elem_sz + c];
                float v10 = img_in[((size_t)y1 * in_cols + x0) * elem_sz + c], v11 = img_in[((size_t)y1 * in_cols + x1) * elem_sz + c];
                img_out[((size_t)gy * out_cols + gx) * elem_sz + c] = (uint8_t)(unsigned int)orc_bilerp(v00, v01, v10, v11, wx, wy);
            }
        }
}

/* tx_disp_scale_kernel (d_tx_scale.cu:8-28) */
void orc_disp_scale(float *disp_out, const float *disp_in, int out_rows, int out_cols, int in_rows, int in_cols, float scale)
{
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < out_rows; ++ty)
        for (int tx = 0; tx < out_cols; ++tx) {
            float xs = orc_sample_coord(tx, out_cols, in_cols), ys = orc_sample_coord(ty, out_rows, in_rows);
            int x0 = (int)floorf(xs), y0 = (int)floorf(ys);
            int x1 = imin(x0 + 1, in_cols - 1), y1 = imin(y0 + 1, in_rows - 1);
            float wx = xs - (float)x0, wy = ys - (float)y0;
            float v = orc_bilerp(disp_in[(size_t)y0 * in_cols + x0], disp_in[(size_t)y0 * in_cols + x1],
                                 disp_in[(size_t)y1 * in_cols + x0], disp_in[(size_t)y1 * in_cols + x1], wx, wy);
            disp_out[(size_t)ty * out_cols + tx] = v * scale;
        }
}

/* adcensus_stm_2 (d_io.cu:240-508): disparity estimation on bilinearly down-scaled images, disparities scaled
 * back up (x 1/disp_scale), DIBR + interlace at full resolution.  disp_l / disp_r are the FULL-resolution maps. */
void orc_adcensus_stm_2(const uint8_t *img_sbs, float *disp_l, float *disp_r, uint8_t *interlaced,
                        int num_rows, int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out,
                        int num_rows_disp, int num_cols_disp, int elem_sz, float disp_scale,
                        int num_views, int angle, int num_disp, int zero_disp,
                        float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                        int thresh_s, float thresh_h, const float *lut_ad, const float *lut_cen)
{
    size_t plane = (size_t)num_rows * num_cols, imgsz = plane * elem_sz;
    size_t dplane = (size_t)num_rows_disp * num_cols_disp;
    uint8_t *img_l = (uint8_t *)malloc(imgsz), *img_r = (uint8_t *)malloc(imgsz);
    orc_demux_sbs(img_sbs, img_l, img_r, num_rows, num_cols_sbs, num_cols, elem_sz);
    uint8_t *low_l = (uint8_t *)malloc(dplane * elem_sz), *low_r = (uint8_t *)malloc(dplane * elem_sz);
    orc_scale_bilinear(img_l, low_l, num_rows, num_cols, num_rows_disp, num_cols_disp, elem_sz);
    orc_scale_bilinear(img_r, low_r, num_rows, num_cols, num_rows_disp, num_cols_disp, elem_sz);

    float *cost_l = (float *)malloc(sizeof(float) * dplane * num_disp);
    float *cost_r = (float *)malloc(sizeof(float) * dplane * num_disp);
    uint8_t *arms_l = (uint8_t *)malloc(4 * dplane), *arms_r = (uint8_t *)malloc(4 * dplane);
    float *dl = (float *)malloc(sizeof(float) * dplane), *dr = (float *)malloc(sizeof(float) * dplane);
    orc_ci_adcensus(low_l, low_r, cost_l, cost_r, lut_ad, lut_cen, ad_coeff, census_coeff,
                    num_disp, zero_disp, num_rows_disp, num_cols_disp, elem_sz);
    orc_cross_arms(low_l, arms_l, ucd, lcd, usd, lsd, num_rows_disp, num_cols_disp, elem_sz);
    orc_ca_aggregate(cost_l, arms_l, num_disp, num_rows_disp, num_cols_disp);
    orc_cross_arms(low_r, arms_r, ucd, lcd, usd, lsd, num_rows_disp, num_cols_disp, elem_sz);
    orc_ca_aggregate(cost_r, arms_r, num_disp, num_rows_disp, num_cols_disp);
    orc_wta(cost_l, dl, num_disp, zero_disp, num_rows_disp, num_cols_disp);
    orc_wta(cost_r, dr, num_disp, zero_disp, num_rows_disp, num_cols_disp);
    free(cost_l); free(cost_r);
    uint8_t *out_l = (uint8_t *)malloc(dplane), *out_r = (uint8_t *)malloc(dplane);
    orc_dcc(out_l, out_r, dl, dr, num_rows_disp, num_cols_disp);
    orc_irv(dl, out_l, arms_l, thresh_s, thresh_h, num_rows_disp, num_cols_disp, num_disp, zero_disp, usd, 5, 0);
    orc_irv(dr, out_r, arms_r, thresh_s, thresh_h, num_rows_disp, num_cols_disp, num_disp, zero_disp, usd, 5, 0);
    orc_bilateral(dl, 7, 5, 10, num_rows_disp, num_cols_disp, num_disp);
    orc_bilateral(dr, 7, 5, 10, num_rows_disp, num_cols_disp, num_disp);
    free(out_l); free(out_r); free(arms_l); free(arms_r); free(low_l); free(low_r);
    /* d_io.cu:425-426: scale factor 1.0f / disp_scale */
    orc_disp_scale(disp_l, dl, num_rows, num_cols, num_rows_disp, num_cols_disp, 1.0f / disp_scale);
    orc_disp_scale(disp_r, dr, num_rows, num_cols, num_rows_disp, num_cols_disp, 1.0f / disp_scale);
    free(dl); free(dr);

    uint8_t *occl_l = (uint8_t *)malloc(plane), *occl_r = (uint8_t *)malloc(plane);
    float *mask_l = (float *)malloc(sizeof(float) * plane), *mask_r = (float *)malloc(sizeof(float) * plane);
    orc_occl(occl_l, occl_r, disp_l, disp_r, num_rows, num_cols);
    orc_bleed(occl_l, 1, num_rows, num_cols);
    orc_bleed(occl_r, 1, num_rows, num_cols);
    orc_occl_to_mask(mask_l, occl_l, num_rows, num_cols);
    orc_occl_to_mask(mask_r, occl_r, num_rows, num_cols);
    uint8_t *views_mem = (uint8_t *)calloc(imgsz * (size_t)num_views, 1);
    const uint8_t **views = (const uint8_t **)malloc(sizeof(uint8_t *) * num_views);
    views[0] = img_r;
    views[num_views - 1] = img_l;
    for (int v = 1; v < num_views - 1; ++v) {
        float shift = (float)(1.0 - ((1.0 * (double)(float)v) / ((double)(float)num_views - 1.0)));
        uint8_t *dst = views_mem + (size_t)v * imgsz;
        orc_dbm(dst, img_l, img_r, disp_l, disp_r, mask_l, mask_r, shift, 10, 15.0f, num_rows, num_cols, elem_sz);
        views[v] = dst;
    }
    orc_mux_multiview(views, interlaced, num_views, (float)angle, num_rows, num_cols, num_rows_out, num_cols_out, elem_sz, 2);
    free(views); free(views_mem);
    free(occl_l); free(occl_r); free(mask_l); free(mask_r);
    free(img_l); free(img_r);
}
