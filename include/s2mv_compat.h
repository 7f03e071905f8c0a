/*
 * s2mv_compat.h — the reference's own C++ entry points for the hot path,
 * exported by libs2mv.so with the reference's exact (Itanium-mangled) symbols
 * so that video_io.cpp and image_io.cpp link unchanged against this library
 * instead of the reference's device objects.  Prototypes are identical to the
 * reference headers cited per function; each is a one-line shim over the C ABI
 * of s2mv.h.  As in the reference, all return void and a CUDA failure prints
 * to stderr and exits (cuda_utils.h:12-21).
 */
#ifndef S2MV_COMPAT_H
#define S2MV_COMPAT_H

/* d_io.h:32-40   _Z12adcensus_stmPhPfS0_S_iiiiiiiiiiffffiiif */
void adcensus_stm(unsigned char *img_sbs, float *disp_l, float *disp_r, unsigned char *interlaced,
                  int num_rows, int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out, int elem_sz,
                  int num_views, int angle, int num_disp, int zero_disp, float ad_coeff, float census_coeff,
                  float ucd, float lcd, int usd, int lsd, int thresh_s, float thresh_h);
/* d_io.h:42-53   _Z14adcensus_stm_2PhPfS0_S_iiiiiiiifiiiiffffiiif  (half-resolution variant, d_io.cu:240-508) */
void adcensus_stm_2(unsigned char *img_sbs, float *disp_l, float *disp_r, unsigned char *interlaced, int num_rows,
                    int num_cols_sbs, int num_cols, int num_rows_out, int num_cols_out, int num_rows_disp,
                    int num_cols_disp, int elem_sz, float disp_scale, int num_views, int angle, int num_disp,
                    int zero_disp, float ad_coeff, float census_coeff, float ucd, float lcd, int usd, int lsd,
                    int thresh_s, float thresh_h);
/* d_ci_adcensus.h  _Z11ci_adcensusPhS_PPfS1_ffiiiii */
void ci_adcensus(unsigned char *img_l, unsigned char *img_r, float **cost_l, float **cost_r, float ad_coeff,
                 float census_coeff, int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
/* d_ca_cross.h     _Z8ca_crossPhPS_PPfS2_ffiiiiii */
void ca_cross(unsigned char *img, unsigned char **cross, float **cost, float **acost, float ucd, float lcd,
              int usd, int lsd, int num_disp, int num_rows, int num_cols, int elem_sz);
/* d_dc_wta.h       _Z6dc_wtaPPfS_iiii */
void dc_wta(float **cost, float *disp, int num_disp, int zero_disp, int num_rows, int num_cols);
/* d_dc_hslo.h:19-23  _Z7dc_hsloPPfS_PhS1_fffiiiii — a stub in the reference (it allocates, computes penalty maps
 * and returns); here the scanline optimisation it was meant to be, for the left view (s2mv.h: s2mv_dc_so) */
void dc_hslo(float **cost, float *disp, unsigned char *img_l, unsigned char *img_r, float T, float H1, float H2,
             int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
/* d_dr_dcc.h       _Z6dr_dccPhS_PfS0_ii */
void dr_dcc(unsigned char *outliers_l, unsigned char *outliers_r, float *disp_l, float *disp_r, int num_rows,
            int num_cols);
/* d_dr_irv.h       _Z6dr_irvPfPhPS0_ifiiiiii */
void dr_irv(float *disp, unsigned char *outliers, unsigned char **cross, int thresh_s, float thresh_h,
            int num_rows, int num_cols, int num_disp, int zero_disp, int usd, int iterations);
/* d_filter_bilateral.h  _Z18filter_bilateral_1Pfiffiii */
void filter_bilateral_1(float *img, int radius, float sigma_color, float sigma_spatial, int num_rows,
                        int num_cols, int num_disp);
/* d_dibr_occl.h    _Z9dibr_occlPhS_PfS0_ii */
void dibr_occl(unsigned char *occl_l, unsigned char *occl_r, float *disp_l, float *disp_r, int num_rows,
               int num_cols);
/* d_filter.h       _Z14filter_bleed_1Phiii */
void filter_bleed_1(unsigned char *img, int radius, int num_rows, int num_cols);
/* d_dibr_occl.h    _Z17dibr_occl_to_maskPfS_PhS0_ii */
void dibr_occl_to_mask(float *mask_l, float *mask_r, unsigned char *occl_l, unsigned char *occl_r,
                       int num_rows, int num_cols);
/* d_filter_gaussian.h  _Z17filter_gaussian_1Pfifii */
void filter_gaussian_1(float *img, int radius, float sigma_spatial, int num_rows, int num_cols);
/* d_dibr_bwarp.h   _Z8dibr_dbmPhS_S_PfS0_S_S_S0_S0_fiii */
void dibr_dbm(unsigned char *img_out, unsigned char *img_in_l, unsigned char *img_in_r, float *disp_l,
              float *disp_r, unsigned char *occl_l, unsigned char *occl_r, float *mask_l, float *mask_r,
              float shift, int num_rows, int num_cols, int elem_sz);
/* d_dibr_fwarp.h:18-21  _Z8dibr_dfmPhS_S_PfS0_fiii  (never called by the reference's drivers; s2mv.h: s2mv_dibr_dfm) */
void dibr_dfm(unsigned char *img_out, unsigned char *img_in_l, unsigned char *img_in_r, float *disp_l, float *disp_r,
              float shift, int num_rows, int num_cols, int elem_sz);
/* d_mux_multiview.h  _Z13mux_multiviewPPhS_ifiiiii */
void mux_multiview(unsigned char **views, unsigned char *out_data, int num_views, float angle, int in_rows,
                   int in_cols, int out_rows, int out_cols, int elem_sz);

#endif
