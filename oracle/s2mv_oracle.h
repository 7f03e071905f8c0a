/*
 * s2mv_oracle.h — CPU restatement of the reference's stereo->multiview frame
 * pipeline (moddyz/stereo-to-multiview-cuda, `adcensus_stm`, d_io.cu:7-238).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or as the reported CPU baseline.
 *
 * Parity status: the reference ships no golden vectors or tests,
 * so this restatement is pinned against the reference's own kernels compiled
 * for sm_100 (oracle/build_ref.sh -> oracle/_ref/libs2mv_ref.so) on the GPU
 * box (tests/test_ref_parity.py).  Every stage except the `ex2.approx`
 * combine (d_ci_adcensus.cu:10-36) is bit-exact on the CPU; the combine takes
 * its two exponential tables as inputs so that a GPU-produced table makes the
 * whole chain bit-exact.
 *
 * Layouts follow the reference: images are tightly packed interleaved BGR
 * (elem_sz bytes per pixel), cost volumes are D contiguous planes of H*W
 * floats (plane d at base + d*H*W), arms are 4 planes of H*W bytes in the
 * order UP, DOWN, LEFT, RIGHT (d_ca_cross.cu:9-15).
 */
#ifndef S2MV_ORACLE_H
#define S2MV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_AD_LUT_SIZE 766   /* 3*255 + 1 distinct absolute-difference sums */
#define ORC_CEN_LUT_SIZE 65   /* Hamming distances 0..64 (d_alu.cu:7-15)     */

/* d_demux_common.cu:8-33 */
void orc_demux_sbs(const uint8_t *sbs, uint8_t *img_l, uint8_t *img_r,
                   int num_rows, int num_cols_sbs, int num_cols, int elem_sz);
/* d_mux_common.cu:7-21 (PTX: trunc(fma(r,c,fma(b,c,g*c))), c = 0x3EAAAAAB) */
void orc_gray(const uint8_t *img, uint8_t *gray, int num_rows, int num_cols, int elem_sz);
/* d_ci_census.cu:18-50 — full 48-bit string */
void orc_census(const uint8_t *gray, uint64_t *census, int num_rows, int num_cols);
/* d_alu.cu:7-15 — truncated Hamming: popc(lo31) + 33*bit31 */
int orc_hamdist(uint64_t a, uint64_t b);
/* d_ci_ad.cu:73-159 launched as in d_ci_adcensus.cu:46-59,109 (160-wide blocks) */
void orc_ad_cost(const uint8_t *img_l, const uint8_t *img_r, float *cost_l, float *cost_r,
                 int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
/* d_ci_census.cu:197-254 launched as in d_ci_adcensus.cu:117-144 */
void orc_census_cost(const uint64_t *census_l, const uint64_t *census_r,
                     float *cost_l, float *cost_r,
                     int num_disp, int zero_disp, int num_rows, int num_cols);
/* CPU stand-in for 1 - ex2.approx((-c*inv)*log2e), d_ci_adcensus.cu:27-31 */
void orc_exp_luts(float ad_coeff, float census_coeff, float *lut_ad, float *lut_cen);
/* d_ci_adcensus.cu:38-186.  lut_ad/lut_cen may be NULL (CPU tables). */
void orc_ci_adcensus(const uint8_t *img_l, const uint8_t *img_r,
                     float *cost_l, float *cost_r,
                     const float *lut_ad, const float *lut_cen,
                     float ad_coeff, float census_coeff, int num_disp, int zero_disp,
                     int num_rows, int num_cols, int elem_sz);
/* d_ca_cross.cu:17-172 */
void orc_cross_arms(const uint8_t *img, uint8_t *arms, float ucd, float lcd, int usd, int lsd,
                    int num_rows, int num_cols, int elem_sz);
/* d_ca_cross.cu:255-271 + d_ca_cross_sum.cu:148-198,243-293: H,V,V,H in place */
void orc_ca_aggregate(float *cost, const uint8_t *arms, int num_disp, int num_rows, int num_cols);
/* single pass, exposed for per-pass parity: dir 0 = horizontal, 1 = vertical */
void orc_ca_pass(const float *in, float *out, const uint8_t *arms, int dir,
                 int num_disp, int num_rows, int num_cols);
/* d_dc_wta.cu:9-35 */
void orc_wta(const float *cost, float *disp, int num_disp, int zero_disp, int num_rows, int num_cols);
/* four-direction scanline optimisation + WTA: PARITY UNPINNED (the reference's d_dc_hslo.cu is a stub); this is
 * the specification of the stage, see s2mv_oracle.c */
void orc_so(const float *cost, float *cost_out, float *disp, const uint8_t *img_own, const uint8_t *img_other,
            int view, float T, float H1, float H2, int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
/* d_dr_dcc.cu:18-128; outliers are overwritten (reference memsets them to 0 first) */
void orc_dcc(uint8_t *outliers_l, uint8_t *outliers_r, const float *disp_l, const float *disp_r,
             int num_rows, int num_cols);
/* d_dr_irv.cu:17-43,134-269.  host_variant != 0 follows dr_irv (:272-366):
 * one vote pass, then `iterations` applications of the same votes. */
void orc_irv(float *disp, uint8_t *outliers, const uint8_t *arms, int thresh_s, float thresh_h,
             int num_rows, int num_cols, int num_disp, int zero_disp, int usd, int iterations,
             int host_variant);
/* d_filter_gaussian.cu:237-255 and d_filter_bilateral.cu:26-39 (host libm) */
void orc_gaussian_kernel(float *kernel, int radius, float sigma);
void orc_gaussian_1d(float *kernel, int size, float sigma);
/* d_filter_bilateral.cu:222-304,517-568 */
void orc_bilateral(float *img, int radius, float sigma_color, float sigma_spatial,
                   int num_rows, int num_cols, int num_disp);
/* d_dibr_occl.cu:114-159 */
void orc_occl(uint8_t *occl_l, uint8_t *occl_r, const float *disp_l, const float *disp_r,
              int num_rows, int num_cols);
/* d_filter.cu:105-167 */
void orc_bleed(uint8_t *img, int radius, int num_rows, int num_cols);
/* d_dibr_occl.cu:17-31 */
void orc_occl_to_mask(float *mask, const uint8_t *occl, int num_rows, int num_cols);
/* d_filter_gaussian.cu:9-88,130-171: out = max(v, blur(v)) */
void orc_gaussian_dilate(float *img, int radius, float sigma, int num_rows, int num_cols);
/* d_dibr_bwarp.cu:5-22 */
void orc_bwarp(uint8_t *out, const uint8_t *in, const float *mask, const float *disp,
               float shift, int num_rows, int num_cols, int elem_sz);
/* d_mux_common.cu:23-46 */
void orc_fwarp(uint8_t *out, const uint8_t *in, const float *disp, float shift, int *hits,
               int num_rows, int num_cols, int elem_sz);
void orc_dibr_dfm(uint8_t *out, const uint8_t *img_l, const uint8_t *img_r, const float *disp_l, const float *disp_r,
                  float shift, int num_rows, int num_cols, int elem_sz);
void orc_merge_ab(uint8_t *img_b, const uint8_t *img_a, const float *mask_a,
                  int num_rows, int num_cols, int elem_sz);
/* d_dibr_bwarp.cu:24-70 (radius 10 sigma 15) / :75-183 (radius 7 sigma 10) */
void orc_dbm(uint8_t *out, const uint8_t *img_l, const uint8_t *img_r,
             const float *disp_l, const float *disp_r, const float *mask_l, const float *mask_r,
             float shift, int gauss_radius, float gauss_sigma,
             int num_rows, int num_cols, int elem_sz);
/* d_mux_multiview.cu:38-153; kernel_variant 2 = mux_multiview_kernel_2, 1 = mux_multiview_kernel */
void orc_mux_multiview(const uint8_t *const *views, uint8_t *out, int num_views, float angle,
                       int num_rows_in, int num_cols_in, int num_rows_out, int num_cols_out,
                       int elem_sz, int kernel_variant);

/* Optional taps on the full pipeline (any pointer may be NULL). */
typedef struct {
    float *wta_l, *wta_r;           /* H*W, disparities straight after WTA        */
    uint8_t *outliers_l, *outliers_r; /* H*W, after the cross-check (before IRV)  */
    float *irv_l, *irv_r;           /* H*W, after region voting (before bilateral) */
    uint8_t *arms_l, *arms_r;       /* 4*H*W                                       */
    float *mask_l, *mask_r;         /* H*W, after bleed + to-mask                  */
    uint8_t *views;                 /* num_views*H*W*elem_sz                       */
    float *acost_l, *acost_r;       /* D*H*W aggregated cost                       */
} orc_taps_t;

/* d_io.cu:7-238.  `angle` is an int exactly as in d_io.h:36. */
void orc_adcensus_stm(const uint8_t *img_sbs, float *disp_l, float *disp_r, uint8_t *interlaced,
                      int num_rows, int num_cols_sbs, int num_cols,
                      int num_rows_out, int num_cols_out, int elem_sz,
                      int num_views, int angle, int num_disp, int zero_disp,
                      float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                      int thresh_s, float thresh_h,
                      const float *lut_ad, const float *lut_cen, const orc_taps_t *taps);

/* Cost-volume leg only (CI + CA + WTA, both views): the MDE/s metric. */
void orc_costvol(const uint8_t *img_l, const uint8_t *img_r, float *disp_l, float *disp_r,
                 int num_rows, int num_cols, int elem_sz, int num_disp, int zero_disp,
                 float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                 const float *lut_ad, const float *lut_cen);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif

/* half-resolution variant: tx_scale_bilinear_kernel, tx_disp_scale_kernel (d_tx_scale.cu:8-52), adcensus_stm_2
 * (d_io.cu:240-508) */
void orc_scale_bilinear(const uint8_t *img_in, uint8_t *img_out, int in_rows, int in_cols, int out_rows, int out_cols,
                        int elem_sz);
void orc_disp_scale(float *disp_out, const float *disp_in, int out_rows, int out_cols, int in_rows, int in_cols, float scale);
void orc_adcensus_stm_2(const uint8_t *img_sbs, float *disp_l, float *disp_r, uint8_t *interlaced,
                        int num_rows, int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out,
                        int num_rows_disp, int num_cols_disp, int elem_sz, float disp_scale,
                        int num_views, int angle, int num_disp, int zero_disp,
                        float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                        int thresh_s, float thresh_h, const float *lut_ad, const float *lut_cen);

#endif
