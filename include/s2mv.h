/*
 * s2mv.h — C ABI of the B200-native stereo -> multiview frame pipeline.
 *
 * Drop-in boundary for the hot path of moddyz/stereo-to-multiview-cuda:
 * everything `adcensus_stm` (reference d_io.cu:7-238, d_io.h:32-40) executes,
 * plus the per-stage host-pointer entry points the reference's image driver
 * calls (image_io.cpp:171-292).  Plain pointers and sizes only; every
 * function returns an s2mv_status (0 = OK).  The reference's own C++ symbols
 * (`adcensus_stm`, `ci_adcensus`, `ca_cross`, ...) are exported by the same
 * library as one-line shims over this ABI (include/s2mv_compat.h), so
 * video_io.cpp / image_io.cpp link unchanged.
 *
 * Layout conventions are the reference's: images are tightly packed
 * interleaved BGR, `elem_sz` (= 3) bytes per pixel; cost volumes at this
 * boundary are tables of `num_disp` plane pointers, each plane
 * num_rows*num_cols floats; cross arms are tables of 4 plane pointers in the
 * order UP, DOWN, LEFT, RIGHT (d_ca_cross.cu:9-15).
 *
 * There is no CPU fallback: every entry point fails with
 * S2MV_ERR_NO_DEVICE / S2MV_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef S2MV_H
#define S2MV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    S2MV_OK = 0,
    S2MV_ERR_NO_DEVICE = 1,   /* no CUDA device / wrong architecture           */
    S2MV_ERR_CUDA = 2,        /* a CUDA runtime call or kernel launch failed    */
    S2MV_ERR_BAD_PARAM = 3,   /* sizes / parameters outside the supported range */
    S2MV_ERR_NOT_CONFIGURED = 4,
    S2MV_ERR_OOM = 5
} s2mv_status;

/* Parameters of adcensus_stm (d_io.h:32-40) plus the constants the reference
 * hard-codes at its call sites (d_io.cu:147-151, d_dibr_bwarp.cu:63).
 * s2mv_default_params() fills the latter with the video path's values. */
typedef struct {
    int num_rows, num_cols;         /* one view                                 */
    int num_rows_out, num_cols_out; /* interlaced frame                         */
    int elem_sz;                    /* bytes per pixel, must be 3               */
    int num_views, angle;           /* angle in whole degrees (d_io.h:36)       */
    int num_disp, zero_disp;
    float ad_coeff, census_coeff;
    float ucd, lcd;
    int usd, lsd;
    int thresh_s;
    float thresh_h;
    int irv_iterations;             /* 5   d_io.cu:147                          */
    int bilateral_radius;           /* 7   d_io.cu:150                          */
    float bilateral_sigma_color;    /* 5                                        */
    float bilateral_sigma_spatial;  /* 10                                       */
    int mask_blur_radius;           /* 10  d_dibr_bwarp.cu:63                   */
    float mask_blur_sigma;          /* 15                                       */
} s2mv_params;

typedef struct s2mv_ctx s2mv_ctx;

const char *s2mv_status_string(int status);
const char *s2mv_last_error(void);  /* thread-local detail of the last failure */
void s2mv_default_params(s2mv_params *p);

/* Context = one GPU, one stream, one arena sized by s2mv_configure().  */
int s2mv_create(s2mv_ctx **ctx, int device);
void s2mv_destroy(s2mv_ctx *ctx);
int s2mv_configure(s2mv_ctx *ctx, const s2mv_params *p);
size_t s2mv_arena_bytes(const s2mv_ctx *ctx);
/* num_disp > 128: the four aggregation passes are independent per disparity,
 * so the arena can hold ONE 128-disparity chunk of both volumes and run the
 * chunks one after another (winner-takes-all merges through 64-bit keys)
 * instead of holding num_disp-deep volumes: 8K D=512 needs 68 GB instead of
 * 272 GB.  mode -1 (default): chosen at s2mv_configure when the full volumes
 * do not fit the device; 0: never; 1: whenever num_disp > 128.  Call before
 * s2mv_configure.  Results are identical in both layouts. */
int s2mv_set_chunk_sequential(s2mv_ctx *ctx, int mode);
int s2mv_is_chunk_sequential(const s2mv_ctx *ctx);
int s2mv_device_sm_count(const s2mv_ctx *ctx);

/* ---- frame entry points (replace adcensus_stm, d_io.cu:7-238) ---------- */

/* Host buffers in, host buffers out; synchronous (returns after the D2H
 * copies), exactly the contract of adcensus_stm.  `img_sbs` is
 * num_rows x num_cols_sbs x elem_sz; left view = columns [0,num_cols),
 * right view = [num_cols, 2*num_cols).  Outputs may be NULL to skip them. */
int s2mv_process_sbs(s2mv_ctx *ctx, const uint8_t *img_sbs, int num_cols_sbs,
                     float *disp_l, float *disp_r, uint8_t *interlaced);

/* Pageable caller buffers (cv::Mat::data in the reference's drivers) are staged through the context's pinned
 * buffers by default: two extra host copies of 35 MB per 1080p frame.  With host registration a caller buffer
 * is page-locked in place (cudaHostRegister) and DMA'd directly from then on.  mode:
 *   0  staged (default of s2mv_create);
 *   1  always: each distinct buffer is registered the first time it is seen and stays registered until the
 *      context is destroyed or the mode changes;
 *   2  auto (what the adcensus_stm / adcensus_stm_2 shims run, unless S2MV_HOST_REGISTER=0): a buffer is
 *      registered when the SAME pointer arrives in two consecutive calls -- the reference's video loop reuses
 *      its four buffers for the whole run (video_io.cpp:125-158) -- and unregistered as soon as a call arrives
 *      without it.
 * CONTRACT (modes 1 and 2): a buffer must stay allocated while it is registered, i.e. in mode 2 until the
 * first call that no longer passes it (or the mode changes / the context is destroyed); memory freed while
 * registered and handed out again by the allocator would be written through a stale mapping.  Changing the
 * mode unregisters everything. */
int s2mv_set_host_registration(s2mv_ctx *ctx, int mode);

/* Same work on DEVICE pointers, enqueued on `stream` (a cudaStream_t; NULL =
 * the context's own stream) without synchronising. */
int s2mv_process_sbs_device(s2mv_ctx *ctx, const uint8_t *d_img_sbs, int num_cols_sbs,
                            float *d_disp_l, float *d_disp_r, uint8_t *d_interlaced, void *stream);

/* Two-resolution variant = adcensus_stm_2 (d_io.cu:240-508, d_tx_scale.cu:8-52): both views are scaled down
 * bilinearly to num_rows_disp x num_cols_disp, disparities are estimated and refined there, scaled back up
 * (bilinear, x 1/disp_scale) and drive DIBR + interlace at full resolution.  disp_l / disp_r are the
 * FULL-resolution maps.  s2mv_configure_2 replaces s2mv_configure for such a context. */
int s2mv_configure_2(s2mv_ctx *ctx, const s2mv_params *p, int num_rows_disp, int num_cols_disp, float disp_scale);
int s2mv_process_sbs_2(s2mv_ctx *ctx, const uint8_t *img_sbs, int num_cols_sbs,
                       float *disp_l, float *disp_r, uint8_t *interlaced);
int s2mv_process_sbs_2_device(s2mv_ctx *ctx, const uint8_t *d_img_sbs, int num_cols_sbs,
                              float *d_disp_l, float *d_disp_r, uint8_t *d_interlaced, void *stream);

/* Cost-volume leg only: cost initialisation + 4-pass cross aggregation + WTA,
 * both views (the MDE/s metric; BASELINE config 5).  Device pointers. */
int s2mv_costvol_device(s2mv_ctx *ctx, const uint8_t *d_img_sbs, int num_cols_sbs,
                        float *d_disp_l, float *d_disp_r, void *stream);

/* Frame path option (default off: adcensus_stm never runs it): scanline optimisation of both views between
 * cross aggregation and winner-takes-all, with image_io.cpp:311-313's T = 15, H1 = 1, H2 = 3 or the
 * caller's.  Needs num_disp <= 128 and a whole-frame context. */
int s2mv_enable_so(s2mv_ctx *ctx, int on, float T, float H1, float H2);

int s2mv_synchronize(s2mv_ctx *ctx);

/* ---- row-band mode: ONE frame over several contexts / GPUs ----------------
 * The reference has no multi-GPU path; this is the single-large-frame split
 * of SURVEY §8(e).  Each context owns the contiguous rows [band_y0, band_y1)
 * of the frame described by `frame` and works on a sub-image = those rows plus
 * `apron` rows either side (0 = default, >= 6*usd + 18).  The cheap O(W*H)
 * stages run on the sub-image as if it were a whole image (exact on the own
 * rows); the cost-volume passes run on the own rows only and the usd volume
 * rows either side that the vertical passes read are exchanged between
 * neighbouring bands by the caller (cudaMemcpyPeerAsync / NCCL send-recv):
 *
 *   s2mv_band_prepare(frame)           demux, gray, census, arms of the sub-image
 *   s2mv_band_pass(1)                  cost init + horizontal pass     -> volume A
 *   exchange s2mv_band_halo(1, ...)    up/down, both views
 *   s2mv_band_pass(2)                  vertical pass                   -> volume B
 *   exchange s2mv_band_halo(2, ...)
 *   s2mv_band_pass(3); s2mv_band_pass(4)   vertical; horizontal + WTA  -> own rows of s2mv_band_disp()
 *   fill the other rows of s2mv_band_disp() planes from the other bands
 *   s2mv_band_finish()                 refinement + DIBR + interlace, own rows out
 *
 * s2mv_configure_band_ex additionally takes the row count of the frame's SMALLEST band and fuse_vertical (the
 * same values on every band).  When every band has at least 2*usd rows (and num_disp > 64) the two vertical
 * passes can run as one launch: the bands then exchange 2*usd rows of volume A once (s2mv_band_info reports
 * the rows per exchange), s2mv_band_pass(2) does both vertical passes, s2mv_band_halo(2, ...) returns 0 bytes
 * and s2mv_band_pass(3) returns at once -- the call sequence above stays valid as it is.  fuse_vertical: 0 = never,
 * 1 = whenever possible, -1 = where it is faster (bands of 40*usd rows and more).  s2mv_configure_band = _ex
 * with (0, 0).
 *
 * Results equal the single-context frame bit for bit (tests/test_gpu_rowband.py).
 * All pointers are DEVICE pointers; every call is asynchronous on `stream`
 * (NULL = the context's stream).  Requires output size == input size. */
int s2mv_configure_band(s2mv_ctx *ctx, const s2mv_params *frame, int band_y0, int band_y1, int apron);
int s2mv_configure_band_ex(s2mv_ctx *ctx, const s2mv_params *frame, int band_y0, int band_y1, int apron, int min_band_rows,
                           int fuse_vertical);
int s2mv_band_info(const s2mv_ctx *ctx, int *local_y0, int *local_rows, int *own_first, int *own_rows, int *halo_rows);
int s2mv_band_prepare(s2mv_ctx *ctx, const uint8_t *d_img_sbs_frame, int num_cols_sbs, void *stream);
int s2mv_band_pass(s2mv_ctx *ctx, int pass, void *stream);
/* after_pass 1|2, view 0|1, side 0 (towards row 0) | 1, recv 0 (own rows the neighbour needs) | 1 (halo rows
 * to fill); *bytes == 0 at the frame's edges */
int s2mv_band_halo(s2mv_ctx *ctx, int after_pass, int view, int side, int recv, void **d_ptr, size_t *bytes);
int s2mv_band_disp(s2mv_ctx *ctx, int view, float **d_plane);
int s2mv_band_finish(s2mv_ctx *ctx, float *d_disp_l_band, float *d_disp_r_band, uint8_t *d_interlaced_band, void *stream);
/* Peer-to-peer halos: once a band knows its neighbours' volumes -- another context of this process
 * (s2mv_band_connect) or another process's GPU memory mapped through CUDA IPC (s2mv_band_ipc_export on the
 * owner, s2mv_band_ipc_connect on the neighbour; one process per GPU, handles carried by any host channel) --
 * passes 1 and 2 store the rows next to a band edge straight into the neighbour's halo rows over NVLink, from
 * inside the producing kernel, and order themselves with an epoch word per neighbour on the stream.  The
 * caller then exchanges no halos (s2mv_band_halo is not needed); every band of the frame must run the same
 * pass sequence.  side 0 = the band above, 1 = the band below.  s2mv_band_status reports a neighbour that
 * never arrived (it synchronises the stream). */
typedef struct {
    unsigned char mem[3][64];   /* cudaIpcMemHandle_t of volume A, volume B, the epoch words */
    int frame_y0, frame_y1;     /* the band's own rows in the frame */
    int local_y0, vlo, vrows;   /* sub-image origin, first volume row (local), volume rows */
    int num_cols, lptot, frame_rows, device;
    int halo_rows, fused;       /* rows of volume A exchanged per neighbour; vertical passes fused (must agree) */
} s2mv_band_ipc;
int s2mv_band_connect(s2mv_ctx *ctx, int side, s2mv_ctx *neighbour);
int s2mv_band_ipc_export(s2mv_ctx *ctx, s2mv_band_ipc *out);
int s2mv_band_ipc_connect(s2mv_ctx *ctx, int side, const s2mv_band_ipc *neighbour);
int s2mv_band_status(s2mv_ctx *ctx, void *stream);

/* ---- asynchronous frame stream (the video loop, video_io.cpp:139-160) -----
 * The same frames through the same kernels as s2mv_process_sbs, with the host
 * <-> device copies (d_io.cu:43-44,153-154,205) taken off the critical path:
 * `depth` slots of pinned host + device in/out buffers, H2D of frame i+1 and
 * D2H of frame i-1 overlap the kernels of frame i.  Frames come back in
 * submission order.  submit fails (S2MV_ERR_BAD_PARAM) when all slots are in
 * flight; collect blocks until the oldest frame's outputs are on the host.
 *   s2mv_stream_input_buffer: the pinned input buffer the NEXT submit will
 *     use, so a decoder can write into it directly; then submit(ctx, NULL).
 *   s2mv_stream_submit(ctx, img_sbs): a page-locked img_sbs (cudaHostAlloc /
 *     cudaHostRegister) is copied to the device from where it lies and must
 *     stay unchanged until that frame is collected; pageable memory is staged
 *     through the slot's pinned buffer and may be reused at once.
 *   s2mv_stream_collect: copies into disp_l/disp_r/interlaced when non-NULL
 *     and/or returns the slot's pinned output buffers (valid until `depth`
 *     further submits). */
int s2mv_stream_open(s2mv_ctx *ctx, int depth, int num_cols_sbs);
int s2mv_stream_input_buffer(s2mv_ctx *ctx, uint8_t **pinned_img_sbs);
int s2mv_stream_submit(s2mv_ctx *ctx, const uint8_t *img_sbs);
int s2mv_stream_collect(s2mv_ctx *ctx, float *disp_l, float *disp_r, uint8_t *interlaced,
                        const float **pinned_disp_l, const float **pinned_disp_r,
                        const uint8_t **pinned_interlaced);
int s2mv_stream_pending(const s2mv_ctx *ctx);
int s2mv_stream_close(s2mv_ctx *ctx);

/* Device-event timing of the last s2mv_process_sbs* call, in milliseconds:
 * [0] prepare (demux, gray, census, arms)  [1] cost volume (CI+CA+WTA)
 * [2] refinement (DCC, IRV, bilateral)     [3] DIBR + interlace
 * Valid after s2mv_synchronize(); returns S2MV_ERR_BAD_PARAM if timing was
 * not enabled with s2mv_enable_timing(ctx, 1). */
int s2mv_enable_timing(s2mv_ctx *ctx, int on);
int s2mv_last_timings(s2mv_ctx *ctx, float ms[4]);
/* the four cost-volume kernels of that call: [0] cost init + horizontal pass 1,
 * [1] vertical pass 2, [2] vertical pass 3, [3] horizontal pass 4 + WTA */
int s2mv_last_costvol_kernel_timings(s2mv_ctx *ctx, float ms[4]);
/* number of kernels the last frame call launched */
int s2mv_last_launch_count(const s2mv_ctx *ctx);

/* The two exponential tables of the combine step, as the GPU computes them
 * (1 - ex2.approx((-c/coeff)*log2e), d_ci_adcensus.cu:27-31): 766 floats
 * indexed by |dB|+|dG|+|dR| and 65 floats indexed by Hamming distance. */
int s2mv_get_exp_tables(s2mv_ctx *ctx, float ad_coeff, float census_coeff,
                        float *lut_ad_766, float *lut_cen_65);
/* The AD exponential term exactly as the fused cost-initialisation kernel
 * evaluates it in place (no table; ex2.approx.ftz) for every |dB|+|dG|+|dR| in
 * 0..765: must equal lut_ad_766 bit for bit (tests/test_gpu_stages.py). */
int s2mv_get_ad_terms(s2mv_ctx *ctx, float ad_coeff, float *ad_terms_766);

/* Taps on the last frame processed by this context (host destinations, any
 * may be NULL): WTA disparities, cross-check outliers, post-IRV disparities,
 * packed arms (4 planes each), occlusion masks, the num_views synthesised
 * views.  For parity tests; s2mv_enable_taps(ctx, 1) must precede the frame
 * (it adds a few device-to-device copies per frame, so it is off by default). */
int s2mv_enable_taps(s2mv_ctx *ctx, int on);
int s2mv_read_taps(s2mv_ctx *ctx, float *wta_l, float *wta_r,
                   uint8_t *outliers_l, uint8_t *outliers_r,
                   float *irv_l, float *irv_r,
                   uint8_t *arms_l, uint8_t *arms_r,
                   float *mask_l, float *mask_r, uint8_t *views);

/* ---- per-stage entry points, HOST pointers (image_io.cpp:171-292) ------
 * Each one uploads, runs the same kernels the frame path uses, downloads;
 * they create a transient context on device 0 when `ctx` is NULL. */

/* d_ci_adcensus.cu:188-378 */
int s2mv_ci_adcensus(s2mv_ctx *ctx, const uint8_t *img_l, const uint8_t *img_r,
                     float **cost_l, float **cost_r, float ad_coeff, float census_coeff,
                     int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
/* building blocks of the above, exposed for bit-exact integer parity:
 * gray (d_mux_common.cu:7-21), 48-bit census (d_ci_census.cu:18-50),
 * AD cost (d_ci_ad.cu:73-159), Hamming cost (d_ci_census.cu:197-254) */
int s2mv_gray(s2mv_ctx *ctx, const uint8_t *img, uint8_t *gray, int num_rows, int num_cols, int elem_sz);
int s2mv_census(s2mv_ctx *ctx, const uint8_t *gray, uint64_t *census, int num_rows, int num_cols);
int s2mv_ci_ad(s2mv_ctx *ctx, const uint8_t *img_l, const uint8_t *img_r, float **cost_l, float **cost_r,
               int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
int s2mv_ci_census(s2mv_ctx *ctx, const uint8_t *img_l, const uint8_t *img_r, float **cost_l, float **cost_r,
                   int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
/* d_ca_cross.cu:275-444: arms out (4 planes), cost in, aggregated cost out */
int s2mv_ca_cross(s2mv_ctx *ctx, const uint8_t *img, uint8_t **cross, float **cost, float **acost,
                  float ucd, float lcd, int usd, int lsd,
                  int num_disp, int num_rows, int num_cols, int elem_sz);
/* d_dc_wta.cu:61-122 */
int s2mv_dc_wta(s2mv_ctx *ctx, float **cost, float *disp, int num_disp, int zero_disp,
                int num_rows, int num_cols);
/* Scanline optimisation + WTA of ONE view's aggregated cost: the stage the reference declares as dc_hslo
 * (d_dc_hslo.cu:97-101) but never finished or called (image_io.cpp:307-316) -- PARITY UNPINNED, there is no
 * reference output; the specification is oracle/s2mv_oracle.c:orc_so (Mei et al. 2011, four directions, the
 * stub's constants, colour measure and penalty tiers).  view 0: img_own = left image, 1: img_own = right.
 * disp and cost_out (num_disp planes receiving the optimised cost) may each be NULL.  num_disp <= 128. */
int s2mv_dc_so(s2mv_ctx *ctx, float **cost, float *disp, float **cost_out, const uint8_t *img_own,
               const uint8_t *img_other, int view, float T, float H1, float H2,
               int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz);
/* d_dr_dcc.cu:130-205 */
int s2mv_dr_dcc(s2mv_ctx *ctx, uint8_t *outliers_l, uint8_t *outliers_r,
                const float *disp_l, const float *disp_r, int num_rows, int num_cols);
/* d_dr_irv.cu:272-366 when host_variant != 0 (one vote pass), d_dr_irv.cu:222-269 otherwise */
int s2mv_dr_irv(s2mv_ctx *ctx, float *disp, uint8_t *outliers, uint8_t **cross,
                int thresh_s, float thresh_h, int num_rows, int num_cols,
                int num_disp, int zero_disp, int usd, int iterations, int host_variant);
/* d_filter_bilateral.cu:570-632 */
int s2mv_filter_bilateral_1(s2mv_ctx *ctx, float *img, int radius, float sigma_color, float sigma_spatial,
                            int num_rows, int num_cols, int num_disp);
/* d_dibr_occl.cu:161-220 */
int s2mv_dibr_occl(s2mv_ctx *ctx, uint8_t *occl_l, uint8_t *occl_r,
                   const float *disp_l, const float *disp_r, int num_rows, int num_cols);
/* d_filter.cu:169-206 */
int s2mv_filter_bleed_1(s2mv_ctx *ctx, uint8_t *img, int radius, int num_rows, int num_cols);
/* d_dibr_occl.cu:33-112 */
int s2mv_dibr_occl_to_mask(s2mv_ctx *ctx, float *mask_l, float *mask_r,
                           const uint8_t *occl_l, const uint8_t *occl_r, int num_rows, int num_cols);
/* d_filter_gaussian.cu:173-235 (out = max(v, blur v)) */
int s2mv_filter_gaussian_1(s2mv_ctx *ctx, float *img, int radius, float sigma_spatial,
                           int num_rows, int num_cols);
/* d_dibr_bwarp.cu:75-183 when blur_radius/sigma = 7/10; the frame path uses 10/15 */
int s2mv_dibr_dbm(s2mv_ctx *ctx, uint8_t *img_out, const uint8_t *img_in_l, const uint8_t *img_in_r,
                  const float *disp_l, const float *disp_r, const float *mask_l, const float *mask_r,
                  float shift, int blur_radius, float blur_sigma,
                  int num_rows, int num_cols, int elem_sz);
/* d_dibr_fwarp.cu:97-197 (dibr_dfm): forward-warp alternative to dibr_dbm.  Neither reference driver calls it and its
 * scatter races when several sources land on one destination (no ordering); here the lowest source column wins.
 * PARITY UNPINNED on colliding destinations; identical to the reference everywhere else (tests/test_fwarp.py). */
int s2mv_dibr_dfm(s2mv_ctx *ctx, uint8_t *img_out, const uint8_t *img_in_l, const uint8_t *img_in_r,
                  const float *disp_l, const float *disp_r, float shift,
                  int num_rows, int num_cols, int elem_sz);
/* d_mux_multiview.cu:155-222; kernel_variant 0 = choose as the reference does
 * (kernel 2 when num_rows_out % num_views == 0, else kernel 1) */
int s2mv_mux_multiview(s2mv_ctx *ctx, uint8_t **views, uint8_t *out, int num_views, float angle,
                       int num_rows_in, int num_cols_in, int num_rows_out, int num_cols_out,
                       int elem_sz, int kernel_variant);

#ifdef __cplusplus
}
#endif
#endif /* S2MV_H */
