/*
 * ref_harness.cu — headless C entry points over the UNMODIFIED reference
 * (moddyz/stereo-to-multiview-cuda) compiled for sm_100 by build_ref.sh.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference's own host drivers (image_io.cpp,
 * video_io.cpp) need OpenCV 2.x and a display; this file replaces them with
 * plain C functions that forward to the reference's C++ entry points so that
 * tests/ can pin the CPU oracle and the product against the reference's own
 * kernels on the GPU box.  No reference source is copied: the functions
 * below only call what the reference's headers declare.
 */
#include <cstdio>
#include <map>
#include <cuda_runtime.h>

#include "d_io.h" /* reference header (OpenCV includes are satisfied by empty stubs) */

extern "C" {

int ref_device_ok(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

/* d_io.h:32-40 */
void ref_adcensus_stm(unsigned char *img_sbs, float *disp_l, float *disp_r, unsigned char *interlaced,
                      int num_rows, int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out,
                      int elem_sz, int num_views, int angle, int num_disp, int zero_disp,
                      float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                      int thresh_s, float thresh_h)
{
    adcensus_stm(img_sbs, disp_l, disp_r, interlaced, num_rows, num_cols_sbs, num_cols, num_rows_out,
                 num_cols_out, elem_sz, num_views, angle, num_disp, zero_disp, ad_coeff, census_coeff,
                 ucd, lcd, usd, lsd, thresh_s, thresh_h);
    cudaDeviceSynchronize();
}

/* Host-pointer stage wrappers, image_io.cpp:171-292 call order. */
void ref_ci_adcensus(unsigned char *img_l, unsigned char *img_r, float **cost_l, float **cost_r,
                     float ad_coeff, float census_coeff, int num_disp, int zero_disp,
                     int num_rows, int num_cols, int elem_sz)
{
    ci_adcensus(img_l, img_r, cost_l, cost_r, ad_coeff, census_coeff, num_disp, zero_disp, num_rows, num_cols, elem_sz);
}

void ref_ca_cross(unsigned char *img, unsigned char **cross, float **cost, float **acost,
                  float ucd, float lcd, int usd, int lsd, int num_disp, int num_rows, int num_cols, int elem_sz)
{
    ca_cross(img, cross, cost, acost, ucd, lcd, usd, lsd, num_disp, num_rows, num_cols, elem_sz);
}

void ref_dc_wta(float **cost, float *disp, int num_disp, int zero_disp, int num_rows, int num_cols)
{
    dc_wta(cost, disp, num_disp, zero_disp, num_rows, num_cols);
}

void ref_dr_dcc(unsigned char *outliers_l, unsigned char *outliers_r, float *disp_l, float *disp_r,
                int num_rows, int num_cols)
{
    dr_dcc(outliers_l, outliers_r, disp_l, disp_r, num_rows, num_cols);
}

void ref_dr_irv(float *disp, unsigned char *outliers, unsigned char **cross, int thresh_s, float thresh_h,
                int num_rows, int num_cols, int num_disp, int zero_disp, int usd, int iterations)
{
    dr_irv(disp, outliers, cross, thresh_s, thresh_h, num_rows, num_cols, num_disp, zero_disp, usd, iterations);
}

void ref_filter_bilateral_1(float *img, int radius, float sigma_color, float sigma_spatial,
                            int num_rows, int num_cols, int num_disp)
{
    filter_bilateral_1(img, radius, sigma_color, sigma_spatial, num_rows, num_cols, num_disp);
}

void ref_dibr_occl(unsigned char *occl_l, unsigned char *occl_r, float *disp_l, float *disp_r,
                   int num_rows, int num_cols)
{
    dibr_occl(occl_l, occl_r, disp_l, disp_r, num_rows, num_cols);
}

void ref_filter_bleed_1(unsigned char *img, int radius, int num_rows, int num_cols)
{
    filter_bleed_1(img, radius, num_rows, num_cols);
}

void ref_dibr_occl_to_mask(float *mask_l, float *mask_r, unsigned char *occl_l, unsigned char *occl_r,
                           int num_rows, int num_cols)
{
    dibr_occl_to_mask(mask_l, mask_r, occl_l, occl_r, num_rows, num_cols);
}

void ref_dibr_dbm(unsigned char *img_out, unsigned char *img_in_l, unsigned char *img_in_r,
                  float *disp_l, float *disp_r, unsigned char *occl_l, unsigned char *occl_r,
                  float *mask_l, float *mask_r, float shift, int num_rows, int num_cols, int elem_sz)
{
    dibr_dbm(img_out, img_in_l, img_in_r, disp_l, disp_r, occl_l, occl_r, mask_l, mask_r, shift,
             num_rows, num_cols, elem_sz);
}

/* d_dibr_fwarp.h:18-21 (never called by the reference's drivers; racy where sources collide).
 * dibr_dfm frees d_img_out_r twice (d_dibr_fwarp.cu:186,193): the second cudaFree leaves cudaErrorInvalidValue as the
 * runtime's last error, which any other user of the same libcudart in the process (torch) would trip over -- it is
 * read and dropped here. */
void ref_dibr_dfm(unsigned char *img_out, unsigned char *img_in_l, unsigned char *img_in_r, float *disp_l, float *disp_r,
                  float shift, int num_rows, int num_cols, int elem_sz)
{
    dibr_dfm(img_out, img_in_l, img_in_r, disp_l, disp_r, shift, num_rows, num_cols, elem_sz);
    cudaDeviceSynchronize();
    cudaGetLastError();
}

void ref_mux_multiview(unsigned char **views, unsigned char *out, int num_views, float angle,
                       int num_rows_in, int num_cols_in, int num_rows_out, int num_cols_out, int elem_sz)
{
    mux_multiview(views, out, num_views, angle, num_rows_in, num_cols_in, num_rows_out, num_cols_out, elem_sz);
}

/* Time the reference's own device pipeline (adcensus_stm, per-frame malloc/free
 * included, exactly as shipped) with CUDA events: returns ms per call, best of n. */
float ref_time_adcensus_stm(unsigned char *img_sbs, float *disp_l, float *disp_r, unsigned char *interlaced,
                            int num_rows, int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out,
                            int elem_sz, int num_views, int angle, int num_disp, int zero_disp,
                            float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                            int thresh_s, float thresh_h, int warmup, int iters)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int i = 0; i < warmup + iters; ++i) {
        cudaEventRecord(a);
        adcensus_stm(img_sbs, disp_l, disp_r, interlaced, num_rows, num_cols_sbs, num_cols, num_rows_out,
                     num_cols_out, elem_sz, num_views, angle, num_disp, zero_disp, ad_coeff, census_coeff,
                     ucd, lcd, usd, lsd, thresh_s, thresh_h);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (i >= warmup && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return best;
}


#ifdef REF_PATCHED
/* ---- libs2mv_ref_patched.so only (oracle/build_ref.sh, patches P1-P6, T, A) ------------------------------
 * T: stage taps.  The scratch copy of d_io.cu calls ref_tap(id, ...) at five points of adcensus_stm; a tap
 *    copies the two device buffers into host buffers a test registered with ref_tap_set, and is a no-op
 *    otherwise.  ids: 0 aggregated cost volumes (L, R: [D][H][W] f32), 1 WTA disparities, 2 outliers after
 *    the cross-check, 3 disparities after region voting, 4 masks, 5 outliers after region voting,
 *    6 cross arms ([4][H][W] u8 per view), 7 (ref_tap_views) the num_views view images.
 * A: every cudaMalloc / cudaFree of the reference's translation units arrives here (they are compiled with
 *    -DcudaMalloc=ref_pool_malloc -DcudaFree=ref_pool_free).  Pool off (default): forwarded unchanged.
 *    Pool on: freed blocks are kept and handed back to the next request of the same size, which removes
 *    the per-frame allocation cost from the second frame on without touching the reference's code
 *    (SURVEY §8d "with and without its per-frame malloc/free"). */
static void *g_tap_a[8], *g_tap_b[8];
static bool g_pool_on = false;
static std::multimap<size_t, void *> g_pool_free;
static std::map<void *, size_t> g_pool_live;
static long g_pool_hits = 0, g_pool_misses = 0;

void ref_tap_set(int id, void *host_a, void *host_b)
{
    if (id >= 0 && id < 8) { g_tap_a[id] = host_a; g_tap_b[id] = host_b; }
}

void ref_tap(int id, const void *da, const void *db, size_t bytes)
{
    if (id < 0 || id >= 8) return;
    if (g_tap_a[id]) cudaMemcpy(g_tap_a[id], da, bytes, cudaMemcpyDeviceToHost);
    if (g_tap_b[id]) cudaMemcpy(g_tap_b[id], db, bytes, cudaMemcpyDeviceToHost);
}

void ref_tap_views(unsigned char **views, int num_views, size_t bytes)
{
    if (!g_tap_a[7]) return;
    for (int v = 0; v < num_views; ++v)
        cudaMemcpy((unsigned char *)g_tap_a[7] + (size_t)v * bytes, views[v], bytes, cudaMemcpyDeviceToHost);
}

cudaError_t ref_pool_malloc(void **p, size_t bytes)
{
    if (g_pool_on) {
        auto it = g_pool_free.find(bytes);
        if (it != g_pool_free.end()) {
            *p = it->second;
            g_pool_free.erase(it);
            g_pool_live[*p] = bytes;
            ++g_pool_hits;
            return cudaSuccess;
        }
        ++g_pool_misses;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess && g_pool_on) g_pool_live[*p] = bytes;
    return e;
}

cudaError_t ref_pool_free(void *p)
{
    if (g_pool_on) {
        auto it = g_pool_live.find(p);
        if (it != g_pool_live.end()) {
            /* cudaFree orders itself after all device work; keep that for the next user of the block */
            cudaDeviceSynchronize();
            g_pool_free.insert({it->second, p});
            g_pool_live.erase(it);
            return cudaSuccess;
        }
    }
    return cudaFree(p);
}

void ref_pool_enable(int on)
{
    if (!on) {
        for (auto &kv : g_pool_free) cudaFree(kv.second);
        g_pool_free.clear();
        g_pool_live.clear();
    }
    g_pool_on = on != 0;
    g_pool_hits = g_pool_misses = 0;
}

void ref_pool_stats(long *hits, long *misses)
{
    *hits = g_pool_hits;
    *misses = g_pool_misses;
}
#endif /* REF_PATCHED */

} /* extern "C" */
