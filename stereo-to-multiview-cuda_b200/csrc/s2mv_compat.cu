// s2mv_compat.cu — the reference's C++ symbols as shims over the C ABI.
// See include/s2mv_compat.h.  Errors follow cuda_utils.h:12-21: message on
// stderr, exit(1).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/s2mv.h"
#include "../../include/s2mv_compat.h"

static void check(int status, const char *who)
{
    if (status != S2MV_OK) {
        fprintf(stderr, "%s: %s (%s)\n", who, s2mv_status_string(status), s2mv_last_error());
        exit(1);
    }
}

static s2mv_ctx *g_ctx = nullptr;
static s2mv_params g_prm;

void adcensus_stm(unsigned char *img_sbs, float *disp_l, float *disp_r, unsigned char *interlaced, int num_rows,
                  int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out, int elem_sz, int num_views,
                  int angle, int num_disp, int zero_disp, float ad_coeff, float census_coeff, float ucd, float lcd,
                  int usd, int lsd, int thresh_s, float thresh_h)
{
    // The reference allocates and frees everything per frame; here the context and its
    // arena persist across calls and are rebuilt only when a parameter changes.
    s2mv_params p;
    s2mv_default_params(&p);
    p.num_rows = num_rows; p.num_cols = num_cols; p.num_rows_out = num_rows_out; p.num_cols_out = num_cols_out;
    p.elem_sz = elem_sz; p.num_views = num_views; p.angle = angle; p.num_disp = num_disp; p.zero_disp = zero_disp;
    p.ad_coeff = ad_coeff; p.census_coeff = census_coeff; p.ucd = ucd; p.lcd = lcd; p.usd = usd; p.lsd = lsd;
    p.thresh_s = thresh_s; p.thresh_h = thresh_h;
    if (!g_ctx) {
        check(s2mv_create(&g_ctx, 0), "adcensus_stm");
        memset(&g_prm, 0, sizeof(g_prm));
        // The reference's video loop reuses its frame / disparity / output buffers (video_io.cpp:125-158): the
        // shim page-locks a caller buffer in place once the same pointer has arrived twice in a row, instead of
        // staging 35 MB through pinned memory every frame, and lets go of it when it stops arriving
        // (s2mv_set_host_registration mode 2, contract in s2mv.h).  S2MV_HOST_REGISTER=0 keeps the staging path.
        if (!(getenv("S2MV_HOST_REGISTER") && atoi(getenv("S2MV_HOST_REGISTER")) == 0))
            check(s2mv_set_host_registration(g_ctx, 2), "adcensus_stm");
    }
    if (memcmp(&p, &g_prm, sizeof(p)) != 0) {
        check(s2mv_configure(g_ctx, &p), "adcensus_stm");
        g_prm = p;
    }
    check(s2mv_process_sbs(g_ctx, img_sbs, num_cols_sbs, disp_l, disp_r, interlaced), "adcensus_stm");
}

static s2mv_ctx *g_ctx2 = nullptr;
static s2mv_params g_prm2;
static int g_rows_disp = 0, g_cols_disp = 0;
static float g_disp_scale = 0.f;

void adcensus_stm_2(unsigned char *img_sbs, float *disp_l, float *disp_r, unsigned char *interlaced, int num_rows,
                    int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out, int num_rows_disp,
                    int num_cols_disp, int elem_sz, float disp_scale, int num_views, int angle, int num_disp,
                    int zero_disp, float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                    int thresh_s, float thresh_h)
{
    s2mv_params p;
    s2mv_default_params(&p);
    p.num_rows = num_rows; p.num_cols = num_cols; p.num_rows_out = num_rows_out; p.num_cols_out = num_cols_out;
    p.elem_sz = elem_sz; p.num_views = num_views; p.angle = angle; p.num_disp = num_disp; p.zero_disp = zero_disp;
    p.ad_coeff = ad_coeff; p.census_coeff = census_coeff; p.ucd = ucd; p.lcd = lcd; p.usd = usd; p.lsd = lsd;
    p.thresh_s = thresh_s; p.thresh_h = thresh_h;
    if (!g_ctx2) {
        check(s2mv_create(&g_ctx2, 0), "adcensus_stm_2");
        memset(&g_prm2, 0, sizeof(g_prm2));
        if (!(getenv("S2MV_HOST_REGISTER") && atoi(getenv("S2MV_HOST_REGISTER")) == 0))
            check(s2mv_set_host_registration(g_ctx2, 2), "adcensus_stm_2");
    }
    if (memcmp(&p, &g_prm2, sizeof(p)) != 0 || num_rows_disp != g_rows_disp || num_cols_disp != g_cols_disp ||
        disp_scale != g_disp_scale) {
        check(s2mv_configure_2(g_ctx2, &p, num_rows_disp, num_cols_disp, disp_scale), "adcensus_stm_2");
        g_prm2 = p; g_rows_disp = num_rows_disp; g_cols_disp = num_cols_disp; g_disp_scale = disp_scale;
    }
    check(s2mv_process_sbs_2(g_ctx2, img_sbs, num_cols_sbs, disp_l, disp_r, interlaced), "adcensus_stm_2");
}

void ci_adcensus(unsigned char *img_l, unsigned char *img_r, float **cost_l, float **cost_r, float ad_coeff,
                 float census_coeff, int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz)
{
    check(s2mv_ci_adcensus(nullptr, img_l, img_r, cost_l, cost_r, ad_coeff, census_coeff, num_disp, zero_disp,
                           num_rows, num_cols, elem_sz), "ci_adcensus");
}

void ca_cross(unsigned char *img, unsigned char **cross, float **cost, float **acost, float ucd, float lcd, int usd,
              int lsd, int num_disp, int num_rows, int num_cols, int elem_sz)
{
    check(s2mv_ca_cross(nullptr, img, cross, cost, acost, ucd, lcd, usd, lsd, num_disp, num_rows, num_cols, elem_sz),
          "ca_cross");
}

void dc_wta(float **cost, float *disp, int num_disp, int zero_disp, int num_rows, int num_cols)
{
    check(s2mv_dc_wta(nullptr, cost, disp, num_disp, zero_disp, num_rows, num_cols), "dc_wta");
}

void dc_hslo(float **cost, float *disp, unsigned char *img_l, unsigned char *img_r, float T, float H1, float H2,
             int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz)
{
    check(s2mv_dc_so(nullptr, cost, disp, nullptr, img_l, img_r, 0, T, H1, H2, num_disp, zero_disp, num_rows, num_cols,
                     elem_sz), "dc_hslo");
}

void dr_dcc(unsigned char *outliers_l, unsigned char *outliers_r, float *disp_l, float *disp_r, int num_rows,
            int num_cols)
{
    check(s2mv_dr_dcc(nullptr, outliers_l, outliers_r, disp_l, disp_r, num_rows, num_cols), "dr_dcc");
}

void dr_irv(float *disp, unsigned char *outliers, unsigned char **cross, int thresh_s, float thresh_h, int num_rows,
            int num_cols, int num_disp, int zero_disp, int usd, int iterations)
{
    check(s2mv_dr_irv(nullptr, disp, outliers, cross, thresh_s, thresh_h, num_rows, num_cols, num_disp, zero_disp, usd,
                      iterations, /*host_variant=*/1), "dr_irv");
}

void filter_bilateral_1(float *img, int radius, float sigma_color, float sigma_spatial, int num_rows, int num_cols,
                        int num_disp)
{
    check(s2mv_filter_bilateral_1(nullptr, img, radius, sigma_color, sigma_spatial, num_rows, num_cols, num_disp),
          "filter_bilateral_1");
}

void dibr_occl(unsigned char *occl_l, unsigned char *occl_r, float *disp_l, float *disp_r, int num_rows, int num_cols)
{
    check(s2mv_dibr_occl(nullptr, occl_l, occl_r, disp_l, disp_r, num_rows, num_cols), "dibr_occl");
}

void filter_bleed_1(unsigned char *img, int radius, int num_rows, int num_cols)
{
    check(s2mv_filter_bleed_1(nullptr, img, radius, num_rows, num_cols), "filter_bleed_1");
}

void dibr_occl_to_mask(float *mask_l, float *mask_r, unsigned char *occl_l, unsigned char *occl_r, int num_rows,
                       int num_cols)
{
    check(s2mv_dibr_occl_to_mask(nullptr, mask_l, mask_r, occl_l, occl_r, num_rows, num_cols), "dibr_occl_to_mask");
}

void filter_gaussian_1(float *img, int radius, float sigma_spatial, int num_rows, int num_cols)
{
    check(s2mv_filter_gaussian_1(nullptr, img, radius, sigma_spatial, num_rows, num_cols), "filter_gaussian_1");
}

void dibr_dbm(unsigned char *img_out, unsigned char *img_in_l, unsigned char *img_in_r, float *disp_l, float *disp_r,
              unsigned char * /*occl_l*/, unsigned char * /*occl_r*/, float *mask_l, float *mask_r, float shift,
              int num_rows, int num_cols, int elem_sz)
{
    // the host wrapper blurs with radius 7, sigma 10 (d_dibr_bwarp.cu:151)
    check(s2mv_dibr_dbm(nullptr, img_out, img_in_l, img_in_r, disp_l, disp_r, mask_l, mask_r, shift, 7, 10.0f,
                        num_rows, num_cols, elem_sz), "dibr_dbm");
}

void dibr_dfm(unsigned char *img_out, unsigned char *img_in_l, unsigned char *img_in_r, float *disp_l, float *disp_r,
              float shift, int num_rows, int num_cols, int elem_sz)
{
    check(s2mv_dibr_dfm(nullptr, img_out, img_in_l, img_in_r, disp_l, disp_r, shift, num_rows, num_cols, elem_sz), "dibr_dfm");
}

void mux_multiview(unsigned char **views, unsigned char *out_data, int num_views, float angle, int in_rows,
                   int in_cols, int out_rows, int out_cols, int elem_sz)
{
    check(s2mv_mux_multiview(nullptr, views, out_data, num_views, angle, in_rows, in_cols, out_rows, out_cols, elem_sz,
                             0), "mux_multiview");
}
