// s2mv_api.cu — context, arena and the C ABI of include/s2mv.h.
//
// The reference allocates and frees >30 device buffers (four of them multi-GB)
// and synchronises the device ~25 times per frame (d_io.cu:43-237).  Here one
// context owns one arena sized at configure time; a frame is a fixed sequence
// of launches on one stream with no host synchronisation inside it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/s2mv.h"
#include "common.cuh"
#include "kernels_cost.cuh"
#include "kernels_dibr.cuh"
#include "kernels_line.cuh"
#include "kernels_line2.cuh"
#include "kernels_prep.cuh"
#include "kernels_refine.cuh"
#include "kernels_so.cuh"

using namespace s2mv;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(e__ == cudaErrorMemoryAllocation ? S2MV_ERR_OOM : S2MV_ERR_CUDA, "%s:%d %s: %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e__));                              \
    } while (0)
#define KCHECK() CU(cudaGetLastError())
#define TRY(call)                     \
    do {                              \
        int s__ = (call);             \
        if (s__ != S2MV_OK) return s__; \
    } while (0)

extern "C" const char *s2mv_status_string(int s)
{
    switch (s) {
        case S2MV_OK: return "ok";
        case S2MV_ERR_NO_DEVICE: return "no usable CUDA device";
        case S2MV_ERR_CUDA: return "CUDA error";
        case S2MV_ERR_BAD_PARAM: return "bad parameter";
        case S2MV_ERR_NOT_CONFIGURED: return "context not configured";
        case S2MV_ERR_OOM: return "out of device memory";
    }
    return "unknown";
}
extern "C" const char *s2mv_last_error(void) { return g_err; }

extern "C" void s2mv_default_params(s2mv_params *p)
{
    memset(p, 0, sizeof(*p));
    p->elem_sz = 3;
    p->num_views = 8;
    p->angle = 18;
    p->num_disp = 64;
    p->zero_disp = 32;
    p->ad_coeff = 10.f;
    p->census_coeff = 30.f;
    p->ucd = 20.f;
    p->lcd = 6.f;
    p->usd = 17;
    p->lsd = 9;
    p->thresh_s = 20;
    p->thresh_h = 0.4f;
    p->irv_iterations = 5;          // d_io.cu:147
    p->bilateral_radius = 7;        // d_io.cu:150
    p->bilateral_sigma_color = 5.f;
    p->bilateral_sigma_spatial = 10.f;
    p->mask_blur_radius = 10;       // d_dibr_bwarp.cu:63
    p->mask_blur_sigma = 15.f;
}

// ------------------------------------------------------ host-side weights
#define REF_PI 3.14159265359f  // d_filter_gaussian.cu:7
// gaussian2D / generateGaussianKernel (d_filter_gaussian.cu:237-255), same libm calls
static void host_gaussian_kernel(std::vector<float> &k, int radius, float sigma)
{
    int w = 2 * radius + 1;
    k.resize((size_t)w * w);
    for (int y = -radius; y <= radius; ++y)
        for (int x = -radius; x <= radius; ++x) {
            float variance = (float)pow((double)sigma, 2.0);
            float exponent = (float)(-(pow((double)(float)x, 2.0) + pow((double)(float)y, 2.0)) / (double)(2 * variance));
            k[(x + radius) + (y + radius) * w] = expf(exponent) / (2 * REF_PI * variance);
        }
}
// gaussian1D_host / generateGaussian1D (d_filter_bilateral.cu:26-39)
static void host_gaussian_1d(std::vector<float> &k, int size, float sigma)
{
    k.resize(size > 0 ? size : 1);
    for (int i = 0; i < size; ++i) {
        float variance = (float)pow((double)sigma, 2.0);
        float power = (float)pow((double)(float)i, 2.0);
        float exponent = -power / (2 * variance);
        k[i] = expf(exponent) / sqrtf(2 * REF_PI * variance);
    }
}

// ------------------------------------------------------------ cost plan
struct CostPlan {
    int D, Dp, LP, LPtot, nchunks, usd, M, S_ci;
    bool chunk_seq;                     // volumes hold ONE 128-disparity chunk; chunks run one after another
    size_t smem_ci;                     // CI-only stage kernels (k_hpass)
    int S_h, S_h4, S_v;                 // outputs per CTA along a row (pass 1 / pass 4) / a column (k_line)
    size_t smem_line_ci, smem_line_h, smem_line_h4, smem_line_v;
    // k_line2 (persistent, pipelined; LP = 32 only): tile geometry per pass kind, 0 = not available
    bool line2;
    int l2_cfg;
    int l2_HP, l2_S_ci, l2_S_h, l2_S_v;
    size_t l2_smem_ci, l2_smem_h, l2_smem_v;
    // k_line_vv (the two vertical passes fused): rows per tile, halo rows, shared memory; vv = usable
    bool vv;
    int vv_S, vv_HP;
    size_t vv_smem;
};

// k_line2 configurations (consumer warps NW, outputs per block B, ring stages NS, outputs per tile ~ NW * B):
//   cfg 0: loaded-tile passes 16 x 6, 3 stages of ~96 outputs;  pass 1 (tiles are computed) 12 x 8, 3 stages
//   cfg 1: every pass 16 x 4, 4 stages of ~64 outputs (one more tile in flight, shorter blocks)
//   cfg 2: as cfg 0, but pass 1 sums blocks of 4 outputs, two per warp (96 outputs per tile)
//   cfg 3: as cfg 0, but the loaded-tile passes run 24 consumer warps x 4-output blocks (96 outputs per tile)
constexpr int kVVNA = 8, kVVB = 6;   // k_line_vv: 8 + 8 consumer warps, 6 rows per block, 48 rows per tile
struct L2Cfg { int NW, B, NS; };
static const L2Cfg kL2Cfg[4][2] = {{{16, 6, 3}, {12, 8, 3}}, {{16, 4, 4}, {16, 4, 4}}, {{16, 6, 3}, {12, 4, 3}}, {{24, 4, 3}, {12, 8, 3}}};  // [cfg][0 loaded, 1 computed]

// outputs per tile along a line of `len` outputs: close to NW * B, a multiple of the block size, the line cut evenly
static int line2_segment(int len, int B, int NW, int SUB = 1)
{
    const int unit = B * SUB, full = NW * unit;  // a warp works on SUB blocks at a time
    const int nseg = (len + full - 1) / full;
    int S = (len + nseg - 1) / nseg;
    S = ((S + unit - 1) / unit) * unit;
    return S < unit ? unit : S;
}

static size_t hpass_smem(int S, int halo, int Dc, int M, bool ci)
{
    size_t P = (size_t)S + 2 * halo;
    size_t b = P * Dc * sizeof(float);
    if (ci) b += 4 * (P + 2 * (size_t)M) * sizeof(uint32_t) + (768 + 68) * sizeof(float);
    return b;
}

static int pick_segment(int W, int halo, int Dc, int M, bool ci, int LP, size_t *smem_out)
{
    // widest segment whose tile leaves room for 3 CTAs per SM (<= 74 KB each),
    // else 2 (<= 110 KB), else 1; a multiple of the pixels processed per sweep
    const int G = kHThreads / LP;
    const size_t budgets[3] = {74 * 1024, 110 * 1024, 220 * 1024};
    for (int b = 0; b < 3; ++b) {
        int S = ((W + G - 1) / G) * G;
        if (S > 256) S = 256;
        for (; S >= G; S -= G) {
            size_t sm = hpass_smem(S, halo, Dc, M, ci);
            if (sm <= budgets[b] && (S >= 4 * halo || b == 2 || S >= W)) {
                *smem_out = sm;
                return S;
            }
        }
    }
    return 0;
}

// Segment length for k_line along a line of `len` outputs: the longest multiple of 4 whose tile lets
// three CTAs share an SM (two, then one, when the halo is too wide for that), then evened out so the
// last segment of the line is not a sliver.
static int pick_line_segment(int len, int halo, int LP, bool ci, size_t *smem_out, int first_budget = 1)
{
    // shared memory per CTA that lets 4 / 3 / 2 / 1 CTAs share an SM
    const size_t budgets[4] = {55 * 1024, 73 * 1024, 110 * 1024, 224 * 1024};
    for (int b = first_budget; b < 4; ++b) {
        int smax = 0;
        for (int S = 4; S <= 512; S += 4)
            if (line_smem_bytes(S, halo, LP, ci) <= budgets[b]) smax = S;
        if (smax == 0 || (smax < 2 * halo && b < 3 && smax < len)) continue;
        const int nseg = (len + smax - 1) / smax;
        int S = (((len + nseg - 1) / nseg) + 3) & ~3;
        if (S > smax) S = smax;
        *smem_out = line_smem_bytes(S, halo, LP, ci);
        return S;
    }
    return 0;
}

static int make_plan(CostPlan &pl, int H, int W, int D, int zd, int usd, int l2_cfg)
{
    if (D < 1 || zd < 0 || zd > D || usd < 0 || usd > 64) return fail(S2MV_ERR_BAD_PARAM, "num_disp/zero_disp/usd out of range");
    pl.D = D;
    pl.usd = usd;
    if (D <= 128) {
        int dp = 4;
        while (dp < D) dp <<= 1;
        pl.Dp = dp;
        pl.nchunks = 1;
        pl.LP = dp / 4;
    } else {
        pl.Dp = ((D + 127) / 128) * 128;
        pl.nchunks = pl.Dp / 128;
        pl.LP = 32;
    }
    pl.LPtot = pl.Dp / 4;
    pl.chunk_seq = false;
    pl.M = zd > D - 1 - zd ? zd : D - 1 - zd;
    if (pl.M < 0) pl.M = 0;
    pl.S_ci = pick_segment(W, usd, 4 * pl.LP, pl.M, true, pl.LP, &pl.smem_ci);
    if (!pl.S_ci) return fail(S2MV_ERR_BAD_PARAM, "tile does not fit shared memory");
    pl.S_h = pick_line_segment(W, usd, pl.LP, true, &pl.smem_line_ci);
    pl.smem_line_h = line_smem_bytes(pl.S_h, usd, pl.LP, false);
    // pass 4 (load, sum, WTA) hides its tile loads better with four smaller CTAs per SM (measured:
    // 0.727 -> 0.684 ms at 1080p D=128); pass 1 computes its tile and prefers the longer segment
    pl.S_h4 = pick_line_segment(W, usd, pl.LP, false, &pl.smem_line_h4, 0);
    pl.S_v = pick_line_segment(H, usd, pl.LP, false, &pl.smem_line_v);
    if (!pl.S_h || !pl.S_v || !pl.S_h4) return fail(S2MV_ERR_BAD_PARAM, "usd too large for the shared-memory tile");
    // persistent pipelined line kernel: one warp per pixel (128 disparities per chunk), three tiles per SM
    pl.line2 = false;
    pl.vv = false;
    if (pl.LP == 32 || pl.LP == 16) {
        // LP = 16 (num_disp <= 64): two pixels per warp, tiles twice as long (same bytes); configuration 0 only
        if (pl.LP == 16) l2_cfg = 0;
        const int SUB = 32 / pl.LP;
        pl.l2_cfg = l2_cfg;
        const L2Cfg &ld = kL2Cfg[l2_cfg][0], &ci = kL2Cfg[l2_cfg][1];
        pl.l2_HP = (usd + 1) & ~1;
        pl.l2_S_ci = l2_cfg == 2 ? line2_segment(W, 8, 12) : line2_segment(W, ci.B, ci.NW, SUB);
        pl.l2_S_h = line2_segment(W, ld.B, ld.NW, SUB);
        pl.l2_S_v = line2_segment(H, ld.B, ld.NW, SUB);
        pl.l2_smem_ci = line2_smem_bytes(pl.l2_S_ci, pl.l2_HP, ci.B, true, ci.NS, pl.LP);
        pl.l2_smem_h = line2_smem_bytes(pl.l2_S_h, pl.l2_HP, ld.B, false, ld.NS, pl.LP);
        pl.l2_smem_v = line2_smem_bytes(pl.l2_S_v, pl.l2_HP, ld.B, false, ld.NS, pl.LP);
        const size_t cap = 227 * 1024;
        pl.line2 = pl.l2_smem_ci <= cap && pl.l2_smem_h <= cap && pl.l2_smem_v <= cap &&
                   pl.l2_S_ci / ci.B <= kL2MaxBlocks && pl.l2_S_h / ld.B <= kL2MaxBlocks && pl.l2_S_v / ld.B <= kL2MaxBlocks;
        pl.vv_HP = ((usd + kVVB - 1) / kVVB) * kVVB;
        if (pl.vv_HP < kVVB) pl.vv_HP = kVVB;
        pl.vv_S = kVVNA * kVVB;
        pl.vv_smem = linevv_smem_bytes<kVVNA, kVVB>(pl.vv_S, pl.vv_HP);
        pl.vv = pl.line2 && pl.LP == 32 && pl.vv_smem <= cap;
    }
    return S2MV_OK;
}

// ------------------------------------------------------------ the context
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct s2mv_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    bool configured = false, timing = false, taps = false;
    // row-band mode (s2mv_band.inl): this context is a sub-image of a taller frame
    bool band = false;
    int band_frame_rows = 0, band_ly0 = 0, band_o0 = 0, band_o1 = 0, band_vlo = 0, band_vhi = 0;
    int band_halo1 = 0;        // rows of pass 1's output exchanged with each neighbour: usd, or 2*usd when band_fused
    bool band_fused = false;   // the two vertical passes run as one launch (k_line_vv) on the band's rows
    // peer-to-peer halos: the neighbours' volumes as seen from this process / device
    struct BandPeer {
        bool connected = false;
        float *vol[2] = {};          // neighbour's volume A / B allocation bases (peer-mapped or same process)
        unsigned int *flag = nullptr;  // the neighbour's epoch word that THIS band bumps
        long long row_bias = 0;      // neighbour allocation row = local row + row_bias
        size_t view_stride4 = 0;     // neighbour's float4 per view
        void *ipc[3] = {};           // mappings to close (cudaIpcOpenMemHandle)
    } band_peer[2];
    unsigned int *band_flags = nullptr;  // [0] bumped by the upper neighbour, [1] by the lower
    unsigned int band_epoch = 0;
    // set to 1 by a k_band_wait whose neighbour never arrived: pinned, mapped host memory, so every later
    // s2mv_band_* call sees it without synchronising (band_status_d = the device's view of the same word)
    unsigned int *band_status_h = nullptr, *band_status_d = nullptr;
    // same-process neighbours that hold pointers into THIS band's volumes (s2mv_band_connect): they are
    // disconnected before the arena they point into is freed
    std::vector<s2mv_ctx *> band_attached_by;
    s2mv_ctx *band_peer_ctx[2] = {};  // the same-process contexts this band stores into (null: IPC or none)
    // two-resolution mode (adcensus_stm_2): this context works at full resolution (no cost volumes), `lo` is a
    // whole context at the disparity resolution
    s2mv_ctx *lo = nullptr;
    bool no_volume = false;
    uint8_t *lo_bgr[2] = {};
    float disp_scale = 1.f;
    bool so_on = false;        // scanline optimisation between aggregation and WTA (s2mv_enable_so)
    float so_T = 15.f, so_H1 = 1.f, so_H2 = 3.f;
    int chunk_seq_mode = -1;  // -1 auto (when the full volumes do not fit), 0 never, 1 whenever D > 128
    s2mv_params prm;
    CostPlan plan;
    size_t arena_bytes = 0;
    std::vector<void *> allocs;

    // arena
    uint8_t *sbs = nullptr;
    size_t sbs_bytes = 0;
    uint32_t *pix[2] = {}, *cen[2] = {}, *arms[2] = {};
    uint8_t *gray[2] = {};
    float *lutAd = nullptr, *lutCen = nullptr;
    float lut_ad_coeff = -1.f, lut_cen_coeff = -1.f;
    float *vol[2] = {};  // ping-pong volumes, each [2 views][H][W][Dp]
    unsigned long long *wta_key[2] = {};
    float *disp[2] = {}, *dispF[2] = {};  // WTA/IRV disparities; bilateral output
    uint8_t *outl[2] = {}, *disoccl[2] = {};
    int *irv_list[2] = {}, *irv_list2[2] = {}, *irv_vote[2] = {}, *irv_count = nullptr;
    uint8_t *irv_hseg[2] = {};  // dense region-voting histograms, [pixel][nbp]; null when not allocated
    uint8_t *irv_stamp[2] = {}, *irv_hchg[2] = {};  // incremental dense voting: change stamps, "span changed" flags
    int irv_nbp = 0;
    float *bil_spatial = nullptr, *bil_colour = nullptr, *gauss_kernel = nullptr;
    std::vector<float> h_gauss_kernel;  // host copy of gauss_kernel (for its fp32 sum)
    std::vector<float> h_bil_spatial;   // host copy of bil_spatial (passed to k_bilateral4p as a kernel parameter)
    uint8_t *occl[2] = {}, *occlB[2] = {};
    float *mask[2] = {}, *tmask = nullptr;
    uint8_t *views = nullptr, *interlaced = nullptr;
    // taps
    float *tap_wta[2] = {}, *tap_irv[2] = {};
    uint8_t *tap_outl[2] = {};
    // pinned staging for the host entry point
    uint8_t *h_sbs = nullptr, *h_interlaced = nullptr;
    float *h_disp[2] = {};
    size_t h_sbs_bytes = 0;
    // timing
    // caller buffers page-locked in place (s2mv_set_host_registration): address -> bytes
    int host_reg = 0;  // 0 staged, 1 page-lock every pageable caller buffer, 2 auto: page-lock the buffers that recur
    std::vector<std::pair<const void *, size_t>> host_regs;
    const void *host_last[4] = {};  // the previous synchronous call's four caller buffers (auto mode)
    // test / A-B hooks read from the environment once, at create time
    int env_irv_dense_min = -1;     // S2MV_IRV_DENSE_MIN (-1: not set)
    bool env_dcc_split = false;     // S2MV_DCC_SPLIT: cross-check and occlusion marks as memsets + separate kernels over byte planes
                                    // (the forms for rows beyond shared memory) instead of the one-row-per-block kernels
    int env_irv_coop = 1;           // S2MV_IRV_COOP: 0 never / 1 when the previous frame's lists were short / 2 always
    int *irv_hint_h = nullptr, *irv_hint_d = nullptr;  // mapped host memory: [view] first-iteration list length of the last frame
    int irv_coop_bps = 0;           // co-resident blocks per SM of k_irv_sparse_all (0: not asked yet, < 0: no cooperative launch)
    size_t irv_coop_smem = 0;
    bool env_irv_colw_set = false;  // S2MV_IRV_COLW given: the column walk whatever the image size (test hook)
    int env_irv_colw = 1;           // S2MV_IRV_COLW: columns per ticket of k_irv_vote_col (1..4; 1 measured best)
    bool env_irv_list_votes = false; // S2MV_IRV_LIST_VOTES: dense iterations vote per list entry (k_irv_vote_dense), not per column
    bool env_bilateral_scalar = false;  // S2MV_BILATERAL_SCALAR
    bool env_line_v1 = false;           // S2MV_LINE_V1: the first form of the cost-volume kernel (k_line) for every plan
    int env_l2_cfg = 0;                 // S2MV_L2_CFG: k_line2 configuration (kL2Cfg)
    // S2MV_VOL_PAD: byte offset of the second volume inside its allocation (multiple of 512).  Measured at 1080p D=128 with
    // the vertical passes as two launches: 0.87 ms per pass at offsets 0 ... 1 MB and 16 MB, 0.765 ms at 3 MB and 1 GB
    // (the read and the write stream of a pass then fall on different DRAM channels at the same time).
    long long env_vol_pad = 3 << 20;
    bool env_line_bulk = false;         // S2MV_LINE_BULK: k_line2 with per-position bulk copies instead of the tensor map
    int *line_ctr = nullptr;            // k_line2 work counters, one per pass
    CUtensorMap tmap_vol[2][2];         // tensor maps of the ping-pong volumes: [buffer A/B][row tile / column tile]
    CUtensorMap tmap_vv;                // buffer A, column tile of the fused vertical passes
    int tmap_rows[2] = {1, 1}, tmap_rows_vv = 1;  // positions per tensor copy of those maps
    int env_tmap_rows = 256;            // S2MV_TMAP_ROWS: wanted positions per tensor copy
    bool env_no_vv = false;             // S2MV_NO_VV: run the vertical passes as two launches
    bool tmap_ok = false;
    long long env_band_wait_spins = 20000000;  // S2MV_BAND_WAIT_SPINS: probes (1 us apart) before a halo wait gives up
    cudaEvent_t ev_refined = nullptr;   // disparities final (before DIBR): the synchronous call starts their D2H here
    cudaStream_t st_aux = nullptr;      // second copy stream of the synchronous host call
    cudaEvent_t ev[5] = {};
    cudaEvent_t kev[5] = {};  // around the four cost-volume kernels
    int launches = 0;
    float y_interval = 0.f;
    // asynchronous frame stream (s2mv_stream.inl)
    struct StreamSlot {
        uint8_t *d_sbs = nullptr, *d_out = nullptr, *h_sbs = nullptr, *h_out = nullptr;
        float *d_disp[2] = {}, *h_disp[2] = {};
        cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
        bool busy = false;
    };
    std::vector<StreamSlot> slots;
    cudaStream_t st_in = nullptr, st_out = nullptr;
    int stream_cols_sbs = 0, slot_head = 0, slot_tail = 0, slots_pending = 0;
};
static void stream_release(s2mv_ctx *c);
static void host_unregister_all(s2mv_ctx *c);

static int dev_alloc(s2mv_ctx *c, void **p, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    CU(cudaMalloc(p, bytes));
    c->allocs.push_back(*p);
    c->arena_bytes += bytes;
    return S2MV_OK;
}
template <typename T>
static int dev_alloc_t(s2mv_ctx *c, T **p, size_t count) { return dev_alloc(c, (void **)p, count * sizeof(T)); }

static void free_arena(s2mv_ctx *c)
{
    stream_release(c);
    // neighbours of this process that store into this arena must stop before it goes away ...
    for (s2mv_ctx *n : c->band_attached_by)
        for (int side = 0; side < 2; ++side)
            if (n->band_peer_ctx[side] == c) {
                cudaSetDevice(n->device);
                cudaStreamSynchronize(n->stream);
                n->band_peer[side] = s2mv_ctx::BandPeer();
                n->band_peer_ctx[side] = nullptr;
            }
    c->band_attached_by.clear();
    cudaSetDevice(c->device);
    // ... and this band forgets the neighbours it stored into
    for (int side = 0; side < 2; ++side) {
        if (s2mv_ctx *n = c->band_peer_ctx[side]) {
            auto &v = n->band_attached_by;
            v.erase(std::remove(v.begin(), v.end(), c), v.end());
        }
        c->band_peer_ctx[side] = nullptr;
    }
    for (auto &pr : c->band_peer) {
        for (void *m : pr.ipc)
            if (m) cudaIpcCloseMemHandle(m);
        pr = s2mv_ctx::BandPeer();
    }
    c->band_flags = nullptr;
    if (c->band_status_h) cudaFreeHost(c->band_status_h);
    c->band_status_h = c->band_status_d = nullptr;
    if (c->irv_hint_h) cudaFreeHost(c->irv_hint_h);
    c->irv_hint_h = c->irv_hint_d = nullptr;
    for (void *p : c->allocs) cudaFree(p);
    c->allocs.clear();
    c->arena_bytes = 0;
    if (c->h_sbs) cudaFreeHost(c->h_sbs);
    if (c->h_interlaced) cudaFreeHost(c->h_interlaced);
    for (int v = 0; v < 2; ++v)
        if (c->h_disp[v]) cudaFreeHost(c->h_disp[v]);
    c->h_sbs = c->h_interlaced = nullptr;
    c->h_disp[0] = c->h_disp[1] = nullptr;
    c->h_sbs_bytes = 0;
    c->sbs = nullptr;
    c->sbs_bytes = 0;
    c->configured = false;
}

extern "C" int s2mv_create(s2mv_ctx **out, int device)
{
    if (!out) return fail(S2MV_ERR_BAD_PARAM, "null ctx pointer");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(S2MV_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(S2MV_ERR_BAD_PARAM, "device %d out of range (%d visible)", device, n);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(S2MV_ERR_NO_DEVICE, "device %d is sm_%d%d; this build targets sm_100a", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));
    s2mv_ctx *c = new s2mv_ctx();
    if (const char *e = getenv("S2MV_IRV_DENSE_MIN")) c->env_irv_dense_min = atoi(e) < 0 ? 0 : atoi(e);
    if (const char *e = getenv("S2MV_IRV_LIST_VOTES")) c->env_irv_list_votes = atoi(e) != 0;
    if (const char *e = getenv("S2MV_DCC_SPLIT")) c->env_dcc_split = atoi(e) != 0;
    if (const char *e = getenv("S2MV_IRV_COOP")) c->env_irv_coop = std::min(2, std::max(0, atoi(e)));
    if (const char *e = getenv("S2MV_IRV_COLW")) { c->env_irv_colw = std::min(4, std::max(1, atoi(e))); c->env_irv_colw_set = true; }
    if (const char *e = getenv("S2MV_BILATERAL_SCALAR")) c->env_bilateral_scalar = atoi(e) != 0;
    if (const char *e = getenv("S2MV_LINE_V1")) c->env_line_v1 = atoi(e) != 0;
    if (const char *e = getenv("S2MV_L2_CFG")) c->env_l2_cfg = std::min(3, std::max(0, atoi(e)));
    if (const char *e = getenv("S2MV_NO_VV")) c->env_no_vv = atoi(e) != 0;
    if (const char *e = getenv("S2MV_VOL_PAD")) c->env_vol_pad = (std::max(0LL, atoll(e)) / 512) * 512;
    if (const char *e = getenv("S2MV_TMAP_ROWS")) c->env_tmap_rows = std::max(1, atoi(e));
    if (const char *e = getenv("S2MV_LINE_BULK")) c->env_line_bulk = atoi(e) != 0;
    if (const char *e = getenv("S2MV_BAND_WAIT_SPINS")) c->env_band_wait_spins = atoll(e) > 0 ? atoll(e) : 1;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return fail(S2MV_ERR_CUDA, "cudaStreamCreate failed");
    }
    for (int i = 0; i < 5; ++i) cudaEventCreate(&c->ev[i]);
    cudaEventCreateWithFlags(&c->ev_refined, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&c->st_aux, cudaStreamNonBlocking);
    for (int i = 0; i < 5; ++i) cudaEventCreate(&c->kev[i]);
    *out = c;
    return S2MV_OK;
}

extern "C" void s2mv_destroy(s2mv_ctx *c)
{
    if (!c) return;
    if (c->lo) { s2mv_destroy(c->lo); c->lo = nullptr; }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_arena(c);
    for (int i = 0; i < 5; ++i) {
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
        if (c->kev[i]) cudaEventDestroy(c->kev[i]);
    }
    host_unregister_all(c);
    if (c->ev_refined) cudaEventDestroy(c->ev_refined);
    if (c->st_aux) cudaStreamDestroy(c->st_aux);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" size_t s2mv_arena_bytes(const s2mv_ctx *c) { return c ? c->arena_bytes : 0; }
extern "C" int s2mv_device_sm_count(const s2mv_ctx *c) { return c ? c->sm_count : 0; }
extern "C" int s2mv_last_launch_count(const s2mv_ctx *c) { return c ? c->launches : 0; }

template <typename K>
static int set_smem(K kernel, size_t bytes)
{
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return S2MV_OK;
}

template <int LP>
static int set_line_attrs()
{
    const size_t big = 227 * 1024;
    TRY(set_smem(k_line<LM_CI_H, LP>, big));
    TRY(set_smem(k_line<LM_H, LP>, big));
    TRY(set_smem(k_line<LM_H_WTA, LP>, big));
    TRY(set_smem(k_line<LM_V, LP>, big));
    return S2MV_OK;
}

template <int NW, int B, int NS>
static int set_line2_attrs()
{
    const size_t big = 227 * 1024;
    TRY((set_smem(k_line2<LM_H, NW, B, NS>, big)));
    TRY((set_smem(k_line2<LM_H_WTA, NW, B, NS>, big)));
    TRY((set_smem(k_line2<LM_V, NW, B, NS>, big)));
    return S2MV_OK;
}

static int set_kernel_attrs()
{
    const size_t big = 227 * 1024;
    TRY(set_smem(k_hpass<1, false, true, false>, big));
    TRY(set_smem(k_hpass<2, false, true, false>, big));
    TRY(set_smem(k_hpass<3, false, true, false>, big));
    TRY(set_line_attrs<1>());
    TRY(set_line_attrs<2>());
    TRY(set_line_attrs<4>());
    TRY(set_line_attrs<8>());
    TRY(set_line_attrs<16>());
    TRY(set_line_attrs<32>());
    TRY((set_smem(k_line_vv<kVVNA, kVVB>, big)));
    TRY((set_smem(k_line2<LM_CI_H, 12, 8, 3>, big)));
    TRY((set_smem(k_line2<LM_CI_H, 16, 4, 4>, big)));
    TRY((set_smem(k_line2<LM_CI_H, 12, 4, 3>, big)));
    TRY((set_line2_attrs<16, 6, 3>()));
    TRY((set_smem(k_line2<LM_CI_H, 12, 8, 3, 16>, big)));
    TRY((set_smem(k_line2<LM_H, 16, 6, 3, 16>, big)));
    TRY((set_smem(k_line2<LM_H_WTA, 16, 6, 3, 16>, big)));
    TRY((set_smem(k_line2<LM_V, 16, 6, 3, 16>, big)));
    TRY((set_line2_attrs<16, 4, 4>()));
    TRY((set_line2_attrs<24, 4, 3>()));
    TRY(set_smem(k_bilateral, 160 * 1024));
    TRY(set_smem(k_bilateral4<7, true>, 64 * 1024));
    TRY(set_smem(k_bilateral4<7, false>, 64 * 1024));
    TRY(set_smem(k_bilateral4p<7>, 96 * 1024));
    TRY(set_smem(k_gauss_dilate4<10>, 64 * 1024));
    TRY(set_smem(k_gauss_dilate, 160 * 1024));
    TRY(set_smem(k_irv_vote, 64 * 1024));
    TRY(set_smem(k_irv_hseg<1>, 100 * 1024));
    TRY(set_smem(k_irv_hseg<2>, 100 * 1024));
    TRY(set_smem(k_irv_hseg<3>, 100 * 1024));
    TRY(set_smem(k_irv_hseg<4>, 100 * 1024));
    TRY(set_smem(k_arms_tile<true>, 160 * 1024));
    TRY(set_smem(k_arms_tile<false>, 160 * 1024));
    return S2MV_OK;
}

// k_line<MODE, LP> for the plan's LP
template <int MODE>
static int launch_line(const CostPlan &pl, dim3 grid, size_t smem, cudaStream_t st, const LineArgs &a)
{
    switch (pl.LP) {
        case 1: k_line<MODE, 1><<<grid, kLineThreads, smem, st>>>(a); break;
        case 2: k_line<MODE, 2><<<grid, kLineThreads, smem, st>>>(a); break;
        case 4: k_line<MODE, 4><<<grid, kLineThreads, smem, st>>>(a); break;
        case 8: k_line<MODE, 8><<<grid, kLineThreads, smem, st>>>(a); break;
        case 16: k_line<MODE, 16><<<grid, kLineThreads, smem, st>>>(a); break;
        case 32: k_line<MODE, 32><<<grid, kLineThreads, smem, st>>>(a); break;
        default: return fail(S2MV_ERR_BAD_PARAM, "unsupported lane count %d", pl.LP);
    }
    KCHECK();
    return S2MV_OK;
}

static int build_luts(s2mv_ctx *c, float ad_coeff, float census_coeff, cudaStream_t st)
{
    if (c->lut_ad_coeff == ad_coeff && c->lut_cen_coeff == census_coeff) return S2MV_OK;
    // d_ci_adcensus.cu:160 passes 1.0/coeff (double) through float kernel parameters
    float inv_ad = (float)(1.0 / ad_coeff), inv_cen = (float)(1.0 / census_coeff);
    k_build_luts<<<(kAdLutSize + 255) / 256, 256, 0, st>>>(inv_ad, inv_cen, c->lutAd, c->lutCen);
    KCHECK();
    c->lut_ad_coeff = ad_coeff;
    c->lut_cen_coeff = census_coeff;
    return S2MV_OK;
}

struct BandSpec { int frame_rows, ly0, o0, o1, sub_rows, min_band_rows, fuse; };
static int configure_impl(s2mv_ctx *c, const s2mv_params *p, const BandSpec *band);

extern "C" int s2mv_configure(s2mv_ctx *c, const s2mv_params *p)
{
    if (c) {
        if (c->lo) { s2mv_destroy(c->lo); c->lo = nullptr; }
        c->no_volume = false;
    }
    return configure_impl(c, p, nullptr);
}

// Tensor maps for k_line2's tile loads: the volume as a 4-D tensor [view][row][column][disparity] of fp32, box =
// one tile (P positions of one line x 128 disparities).  cuTensorMapEncodeTiled is a driver entry point; it is
// resolved through the runtime, so the library does not link against libcuda.  Without it (or on an encoding
// error) the kernel falls back to one 512-byte bulk copy per tile position.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}

// positions per tensor copy: the largest divisor of P not above `want`
static int tmap_split_rows(int P, int want)
{
    for (int r = std::min(P, want); r > 1; --r)
        if (P % r == 0) return r;
    return 1;
}

static bool make_volume_tmaps(s2mv_ctx *c, size_t vol_rows)
{
    const CostPlan &pl = c->plan;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || !pl.line2 || c->no_volume) return false;
    const cuuint64_t W = (cuuint64_t)c->prm.num_cols, dfl = (cuuint64_t)pl.LPtot * 4;
    const cuuint64_t gdim[4] = {dfl, W, (cuuint64_t)vol_rows, 2};
    const cuuint64_t gstr[3] = {dfl * 4, dfl * 4 * W, dfl * 4 * W * (cuuint64_t)vol_rows};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int buf = 0; buf < 2; ++buf)
        for (int vert = 0; vert < 2; ++vert) {
            const int Pfull = (vert ? pl.l2_S_v : pl.l2_S_h) + 2 * pl.l2_HP;
            const cuuint32_t P = (cuuint32_t)tmap_split_rows(Pfull, c->env_tmap_rows);
            c->tmap_rows[vert] = (int)P;
            const cuuint32_t box[4] = {(cuuint32_t)(4 * pl.LP), vert ? 1u : P, vert ? P : 1u, 1};
            CUresult r = enc(&c->tmap_vol[buf][vert], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->vol[buf], gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return false;
        }
    if (pl.vv) {
        c->tmap_rows_vv = tmap_split_rows(pl.vv_S + 2 * pl.vv_HP, c->env_tmap_rows);
        const cuuint32_t box[4] = {128, 1, (cuuint32_t)c->tmap_rows_vv, 1};
        CUresult r = enc(&c->tmap_vv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->vol[0], gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

static int configure_impl(s2mv_ctx *c, const s2mv_params *p, const BandSpec *band)
{
    if (!c || !p) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (p->elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3 (BGR), got %d", p->elem_sz);
    if (p->num_rows < 1 || p->num_cols < 1 || p->num_rows_out < 1 || p->num_cols_out < 1)
        return fail(S2MV_ERR_BAD_PARAM, "image sizes must be positive");
    if (p->num_views < 2 || p->num_views > 16) return fail(S2MV_ERR_BAD_PARAM, "num_views must be in [2,16]");
    if ((size_t)p->num_rows * p->num_cols > 0x7fffffffull / 4) return fail(S2MV_ERR_BAD_PARAM, "image too large");
    if (p->bilateral_radius < 0 || p->bilateral_radius > 24 || p->mask_blur_radius < 0 || p->mask_blur_radius > 24)
        return fail(S2MV_ERR_BAD_PARAM, "filter radius out of range");
    if (p->ad_coeff == 0.f || p->census_coeff == 0.f) return fail(S2MV_ERR_BAD_PARAM, "coefficients must be non-zero");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CostPlan pl;
    TRY(make_plan(pl, p->num_rows, p->num_cols, p->num_disp, p->zero_disp, p->usd, c->env_l2_cfg));
    {   // interlace geometry (d_mux_multiview.cu:146): must not divide by zero later (Q27)
        float yi = (float)((double)(float)p->num_views / tan((double)((float)p->angle * 3.1415926535f) / 180.0) / (double)(float)p->elem_sz);
        if (!(fabsf(yi) < 1e9f) || (int)roundf(yi) == 0)
            return fail(S2MV_ERR_BAD_PARAM, "angle %d gives a zero interlace period", p->angle);
        c->y_interval = yi;
    }
    free_arena(c);
    TRY(set_kernel_attrs());
    c->prm = *p;
    c->plan = pl;
    c->band = band != nullptr;
    if (c->so_on && (pl.nchunks != 1 || band)) c->so_on = false;  // only for num_disp <= 128, whole frames
    if (band) {
        c->band_frame_rows = band->frame_rows; c->band_ly0 = band->ly0; c->band_o0 = band->o0; c->band_o1 = band->o1;
        // The vertical passes read usd rows beyond the rows they write.  Run one after the other, the band receives usd
        // rows of pass 1's output and then usd rows of pass 2's from each neighbour.  Fused (k_line_vv), it receives
        // 2*usd rows of pass 1's output once and forms the usd rows of pass 2's output next to its edges itself; the
        // neighbours must own that many rows, which the caller states (min_band_rows: every band of the frame must
        // reach the same decision).  Fusing pays when the 4*usd extra rows per band (pass 1 stores them twice, the fused
        // kernel walks them) are a small part of it: measured on 4K D=256, 1080-row bands gain 2 %, 540-row bands lose
        // 3 %, 270-row bands lose 3 % (profiles/r2_experiments.md); fuse < 0 takes bands of 40*usd rows and more.
        const bool can_fuse = pl.vv && encode_tiled_fn() != nullptr && !c->env_no_vv && !c->env_line_v1 && !c->env_line_bulk &&
                              band->min_band_rows >= 2 * p->usd;
        const bool fused = can_fuse && (band->fuse > 0 || (band->fuse < 0 && band->min_band_rows >= 40 * p->usd));
        c->band_fused = fused;
        c->band_halo1 = fused ? 2 * p->usd : p->usd;
        c->band_vlo = std::max(0, band->o0 - c->band_halo1);
        c->band_vhi = std::min(band->sub_rows, band->o1 + c->band_halo1);
    }
    c->lut_ad_coeff = c->lut_cen_coeff = -1.f;
    const size_t n = (size_t)p->num_rows * p->num_cols;
    if (pl.nchunks > 1) {
        // D > 128: the four passes are independent per disparity, so the volumes only ever need one
        // 128-disparity chunk when chunks run one after another (WTA merges through 64-bit keys).
        // Chosen automatically when two full ping-pong volumes would not fit the device.
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const double full = 2.0 * 2.0 * (double)n * pl.Dp * sizeof(float);
        const bool want = !band && (c->chunk_seq_mode == 1 || (c->chunk_seq_mode < 0 && full > 0.85 * (double)free_b));
        if (want) {
            pl.chunk_seq = true;
            pl.LPtot = pl.LP;  // volume rows hold one chunk
            c->plan = pl;
        }
    }
    // a row band keeps only its own rows plus usd halo rows either side of the volumes
    const size_t vol_rows = band ? (size_t)(c->band_vhi - c->band_vlo) : (size_t)p->num_rows;
    const size_t vol_elems = c->no_volume ? 4 : 2 * vol_rows * p->num_cols * (pl.chunk_seq ? 4 * pl.LP : pl.Dp);
    for (int v = 0; v < 2; ++v) {
        TRY(dev_alloc_t(c, &c->pix[v], n));
        TRY(dev_alloc_t(c, &c->cen[v], n));
        TRY(dev_alloc_t(c, &c->arms[v], n));
        TRY(dev_alloc_t(c, &c->gray[v], n));
        TRY(dev_alloc_t(c, &c->disp[v], n));
        TRY(dev_alloc_t(c, &c->dispF[v], n));
        TRY(dev_alloc_t(c, &c->outl[v], n));
        TRY(dev_alloc_t(c, &c->disoccl[v], n));
        TRY(dev_alloc_t(c, &c->irv_list[v], n + 16));
        TRY(dev_alloc_t(c, &c->irv_list2[v], n + 16));
        TRY(dev_alloc_t(c, &c->irv_vote[v], n));
        TRY(dev_alloc_t(c, &c->irv_stamp[v], n));
        TRY(dev_alloc_t(c, &c->irv_hchg[v], n));
        TRY(dev_alloc_t(c, &c->occl[v], n));
        TRY(dev_alloc_t(c, &c->occlB[v], n));
        TRY(dev_alloc_t(c, &c->mask[v], n));
        TRY(dev_alloc_t(c, &c->tap_wta[v], n));
        TRY(dev_alloc_t(c, &c->tap_irv[v], n));
        TRY(dev_alloc_t(c, &c->tap_outl[v], n));
        if (pl.nchunks > 1) TRY(dev_alloc_t(c, &c->wta_key[v], n));
        {   // the second ping-pong volume starts S2MV_VOL_PAD bytes into its allocation (the passes read one
            // volume and write the other at the same offsets; their relative placement decides how the two streams fall on
            // the DRAM channels)
            const size_t pad = (v == 1 && !band) ? (size_t)c->env_vol_pad / sizeof(float) : 0;  // bands export allocation bases
            TRY(dev_alloc_t(c, &c->vol[v], vol_elems + pad));
            c->vol[v] += pad;
        }
    }
    {   // dense region voting: one byte per (pixel, histogram bin); skipped when it would crowd the device
        const int nbins = p->num_disp > 65 ? p->num_disp : 65, nbp = ((nbins + 127) / 128) * 128;
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        c->irv_hseg[0] = c->irv_hseg[1] = nullptr;
        c->irv_nbp = 0;
        if (!c->no_volume && nbp <= 512 && 2.0 * (double)n * nbp < 0.25 * (double)free_b) {
            TRY(dev_alloc_t(c, &c->irv_hseg[0], n * nbp));
            TRY(dev_alloc_t(c, &c->irv_hseg[1], n * nbp));
            c->irv_nbp = nbp;
        }
    }
    TRY(dev_alloc_t(c, &c->irv_count, 8 + 2 * 64));
    TRY(dev_alloc_t(c, &c->line_ctr, 8));
    if (band) {
        TRY(dev_alloc_t(c, &c->band_flags, 4));
        CU(cudaMemset(c->band_flags, 0, 4 * sizeof(unsigned int)));
        c->band_epoch = 0;
        CU(cudaHostAlloc((void **)&c->band_status_h, sizeof(unsigned int), cudaHostAllocMapped));
        *c->band_status_h = 0;
        CU(cudaHostGetDevicePointer((void **)&c->band_status_d, c->band_status_h, 0));
    }
    {   // list lengths of the last frame's region voting, written by its kernels for the next frame's launch decision
        CU(cudaHostAlloc((void **)&c->irv_hint_h, 2 * sizeof(int), cudaHostAllocMapped));
        c->irv_hint_h[0] = c->irv_hint_h[1] = -1;
        CU(cudaHostGetDevicePointer((void **)&c->irv_hint_d, c->irv_hint_h, 0));
    }
    TRY(dev_alloc_t(c, &c->tmask, n));
    TRY(dev_alloc_t(c, &c->lutAd, 768));
    TRY(dev_alloc_t(c, &c->lutCen, 68));
    TRY(dev_alloc_t(c, &c->views, (size_t)p->num_views * n * 3));
    TRY(dev_alloc_t(c, &c->interlaced, (size_t)p->num_rows_out * p->num_cols_out * 3));
    {
        std::vector<float> k;
        host_gaussian_kernel(k, p->bilateral_radius, p->bilateral_sigma_spatial);
        TRY(dev_alloc_t(c, &c->bil_spatial, k.size()));
        CU(cudaMemcpy(c->bil_spatial, k.data(), k.size() * sizeof(float), cudaMemcpyHostToDevice));
        c->h_bil_spatial = k;
        host_gaussian_1d(k, p->num_disp, p->bilateral_sigma_color);
        TRY(dev_alloc_t(c, &c->bil_colour, k.size()));
        CU(cudaMemcpy(c->bil_colour, k.data(), k.size() * sizeof(float), cudaMemcpyHostToDevice));
        host_gaussian_kernel(k, p->mask_blur_radius, p->mask_blur_sigma);
        c->h_gauss_kernel = k;
        TRY(dev_alloc_t(c, &c->gauss_kernel, k.size()));
        CU(cudaMemcpy(c->gauss_kernel, k.data(), k.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    TRY(build_luts(c, p->ad_coeff, p->census_coeff, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->tmap_ok = make_volume_tmaps(c, vol_rows);
    if (band && !c->tmap_ok) c->band_fused = false;  // the wider halo stays: the separate passes only use usd rows of it
    c->configured = true;
    return S2MV_OK;
}

extern "C" int s2mv_set_chunk_sequential(s2mv_ctx *c, int mode)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (mode < -1 || mode > 1) return fail(S2MV_ERR_BAD_PARAM, "mode must be -1 (auto), 0 or 1");
    c->chunk_seq_mode = mode;
    return S2MV_OK;
}
extern "C" int s2mv_is_chunk_sequential(const s2mv_ctx *c) { return c && c->configured && c->plan.chunk_seq ? 1 : 0; }

extern "C" int s2mv_enable_so(s2mv_ctx *c, int on, float T, float H1, float H2)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (on && c->configured && (c->plan.nchunks != 1 || c->band))
        return fail(S2MV_ERR_BAD_PARAM, "scanline optimisation needs num_disp <= 128 and a whole-frame context");
    c->so_on = on != 0;
    c->so_T = T; c->so_H1 = H1; c->so_H2 = H2;
    return S2MV_OK;
}

extern "C" int s2mv_enable_timing(s2mv_ctx *c, int on)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    c->timing = on != 0;
    return S2MV_OK;
}
extern "C" int s2mv_enable_taps(s2mv_ctx *c, int on)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    c->taps = on != 0;
    return S2MV_OK;
}

extern "C" int s2mv_synchronize(s2mv_ctx *c)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return S2MV_OK;
}

extern "C" int s2mv_last_timings(s2mv_ctx *c, float ms[4])
{
    if (!c || !ms) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (!c->timing) return fail(S2MV_ERR_BAD_PARAM, "timing not enabled");
    for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
    return S2MV_OK;
}

extern "C" int s2mv_last_costvol_kernel_timings(s2mv_ctx *c, float ms[4])
{
    if (!c || !ms) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (!c->timing) return fail(S2MV_ERR_BAD_PARAM, "timing not enabled");
    for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&ms[i], c->kev[i], c->kev[i + 1]));
    return S2MV_OK;
}

// ------------------------------------------------------- stage launchers
struct Dims { int H, W; };

static int launch_census(s2mv_ctx *c, const uint8_t *g0, const uint8_t *g1, uint32_t *c0, uint32_t *c1, int nviews,
                         int H, int W, cudaStream_t st)
{
    k_census_tile<<<dim3((W + kCenW - 1) / kCenW, (H + kCenH - 1) / kCenH, nviews), kCenW * kCenH / 4, 0, st>>>(
        g0, g1, c0, c1, H, W);
    KCHECK();
    c->launches += 1;
    return S2MV_OK;
}

static int launch_arms(s2mv_ctx *c, const uint32_t *p0, const uint32_t *p1, uint32_t *a0, uint32_t *a1, int nviews,
                       float ucd, float lcd, int usd, int lsd, int H, int W, cudaStream_t st)
{
    if (usd < 0 || usd > 64) return fail(S2MV_ERR_BAD_PARAM, "usd out of range");
    const size_t smem = (size_t)(kArmW + 2 * usd) * (kArmH + 2 * usd) * sizeof(uint32_t);
    const int ui = arm_threshold(ucd), li = arm_threshold(lcd);
    const dim3 g((W + kArmW - 1) / kArmW, (H + kArmH - 1) / kArmH, nviews), b(kArmW, kArmH);
    if (ui >= 0 && ui <= 127 && li >= 0 && li <= 127)
        k_arms_tile<true><<<g, b, smem, st>>>(p0, p1, a0, a1, ui, li, usd, lsd, H, W);
    else
        k_arms_tile<false><<<g, b, smem, st>>>(p0, p1, a0, a1, ui, li, usd, lsd, H, W);
    KCHECK();
    c->launches += 1;
    return S2MV_OK;
}

static int launch_prepare(s2mv_ctx *c, const uint8_t *srcL, const uint8_t *srcR, size_t pitch, uint8_t *bgrL,
                          uint8_t *bgrR, cudaStream_t st)
{
    const s2mv_params &p = c->prm;
    const int H = p.num_rows, W = p.num_cols;
    dim3 g((W + 255) / 256, H);
    k_unpack<<<g, 256, 0, st>>>(srcL, srcR, pitch, c->pix[0], c->pix[1], c->gray[0], c->gray[1], bgrL, bgrR, H, W);
    KCHECK();
    c->launches += 1;
    TRY(launch_census(c, c->gray[0], c->gray[1], c->cen[0], c->cen[1], 2, H, W, st));
    TRY(launch_arms(c, c->pix[0], c->pix[1], c->arms[0], c->arms[1], 2, p.ucd, p.lcd, p.usd, p.lsd, H, W, st));
    return S2MV_OK;
}

static void fill_hargs(const s2mv_ctx *c, HArgs &a, int H, int W, int zd)
{
    const CostPlan &pl = c->plan;
    memset(&a, 0, sizeof(a));
    a.pixL = c->pix[0]; a.pixR = c->pix[1]; a.cenL = c->cen[0]; a.cenR = c->cen[1];
    a.lutAd = c->lutAd; a.lutCen = c->lutCen;
    a.H = H; a.W = W; a.D = pl.D; a.zd = zd;
    a.LP = pl.LP; a.LPtot = pl.LPtot; a.nchunks = pl.nchunks;
    a.halo = pl.usd; a.M = pl.M; a.view_first = 0;
}

static void fill_largs(const s2mv_ctx *c, LineArgs &a, int H, int W, int zd, float ad_coeff)
{
    const CostPlan &pl = c->plan;
    memset(&a, 0, sizeof(a));
    a.pixL = c->pix[0]; a.pixR = c->pix[1]; a.cenL = c->cen[0]; a.cenR = c->cen[1];
    a.lutAd = c->lutAd; a.lutCen = c->lutCen;
    a.inv_ad = (float)(1.0 / ad_coeff);  // d_ci_adcensus.cu:160 (see build_luts)
    a.H = H; a.W = W; a.D = pl.D; a.zd = zd;
    a.LPtot = pl.LPtot; a.nchunks = pl.chunk_seq ? 1 : pl.nchunks;
    a.use_keys = pl.nchunks > 1;
    a.halo = pl.usd; a.view_first = 0;
}

// Image rows a launch works on: outputs [own0, own1); the volume buffers hold rows [vlo, vhi) (a whole
// image: 0, H, 0, H; a row band: its own rows plus usd halo rows either side).
struct RowRange { int own0, own1, vlo, vhi; };

// One of the four aggregation passes (H, V, V, H; d_ca_cross.cu:255-271) over volume buffers A/B for
// `nviews` view slots.  pass 1: (CI ->) H : . -> A  [stage API: B -> A]; pass 2: V : A -> B; pass 3: V : B -> A;
// pass 4: H : A -> B, or A -> disparities.  from_ci: pass 1 builds its input in shared memory (A is only
// written).  to_wta: pass 4 reduces to disparities instead of storing (A is only read).
static int launch_pass(s2mv_ctx *c, LineArgs a, int pass, float4 *A, float4 *B, size_t view_stride4, int nviews,
                       bool from_ci, bool to_wta, const RowRange &rr, cudaStream_t st)
{
    const CostPlan &pl = c->plan;
    const int W = a.W, rows = rr.own1 - rr.own0;
    if (rows <= 0) return S2MV_OK;
    // bias the bases so that image row r lives at (r - vlo) of the allocation
    const size_t bias = (size_t)rr.vlo * W * pl.LPtot;
    const float4 *in_base = ((pass == 1 || pass == 3) ? B : A) - bias;
    float4 *out_base = ((pass == 1 || pass == 3) ? A : B) - bias;
    for (int v = 0; v < nviews; ++v) { a.in[v] = in_base + v * view_stride4; a.out[v] = out_base + v * view_stride4; }
    a.v_begin = rr.own0; a.v_end = rr.own1; a.v_lo = rr.vlo; a.v_hi = rr.vhi;
    if (pl.line2 && !c->env_line_v1) {
        // persistent pipelined kernel: one CTA per SM, lines handed out through a counter zeroed on the stream
        const bool vert = pass == 2 || pass == 3, ci = pass == 1 && from_ci;
        Line2Args L;
        memset(&L, 0, sizeof(L));
        a.ln_first = vert ? 0 : rr.own0;
        L.a = a;
        L.S = vert ? pl.l2_S_v : (ci ? pl.l2_S_ci : pl.l2_S_h);
        L.HP = pl.l2_HP;
        L.P = L.S + 2 * L.HP;
        const int len = vert ? rows : W;
        L.tiles_per_line = (len + L.S - 1) / L.S;
        L.claims_per_line = vert ? 1 : std::max(1, (L.tiles_per_line + 7) / 15);
        L.tiles_per_claim = (L.tiles_per_line + L.claims_per_line - 1) / L.claims_per_line;
        L.nlines = vert ? W : rows;
        L.nz = nviews * a.nchunks;
        L.nclaims = L.nz * L.nlines * L.claims_per_line;
        L.counter = c->line_ctr + pass;
        // the tile source: passes 1 (stage API) and 3 read buffer B, passes 2 and 4 buffer A
        const float4 *src = (pass == 1 || pass == 3) ? B : A;
        const int src_buf = src == reinterpret_cast<float4 *>(c->vol[1]) ? 1 : 0;
        const bool own_vol = src == reinterpret_cast<float4 *>(c->vol[src_buf]) && W == c->prm.num_cols &&
                             (nviews == 1 || view_stride4 == (size_t)(rr.vhi - rr.vlo) * W * pl.LPtot);
        L.use_tmap = c->tmap_ok && !ci && own_vol && !c->env_line_bulk;
        L.row_bias = rr.vlo;
        L.tmap_rows = c->tmap_rows[vert ? 1 : 0];
        const CUtensorMap &tm = c->tmap_vol[src_buf][vert ? 1 : 0];
        CU(cudaMemsetAsync(L.counter, 0, sizeof(int), st));
        const int grid = std::min(c->sm_count, L.nclaims);
        const int mode = ci ? LM_CI_H : (vert ? LM_V : ((pass == 4 && to_wta) ? LM_H_WTA : LM_H));
        const size_t smem = ci ? pl.l2_smem_ci : (vert ? pl.l2_smem_v : pl.l2_smem_h);
#define S2MV_L2_LAUNCH(MODE, NW, B, NS) k_line2<MODE, NW, B, NS><<<grid, (NW + kL2Producers) * 32, smem, st>>>(L, tm)
        if (pl.LP == 16) {
#define S2MV_L2_LAUNCH16(MODE, NW, B, NS) k_line2<MODE, NW, B, NS, 16><<<grid, (NW + kL2Producers) * 32, smem, st>>>(L, tm)
            if (mode == LM_CI_H) S2MV_L2_LAUNCH16(LM_CI_H, 12, 8, 3);
            else if (mode == LM_V) S2MV_L2_LAUNCH16(LM_V, 16, 6, 3);
            else if (mode == LM_H_WTA) S2MV_L2_LAUNCH16(LM_H_WTA, 16, 6, 3);
            else S2MV_L2_LAUNCH16(LM_H, 16, 6, 3);
#undef S2MV_L2_LAUNCH16
        } else if (pl.l2_cfg == 3 && mode != LM_CI_H) {
            if (mode == LM_V) S2MV_L2_LAUNCH(LM_V, 24, 4, 3);
            else if (mode == LM_H_WTA) S2MV_L2_LAUNCH(LM_H_WTA, 24, 4, 3);
            else S2MV_L2_LAUNCH(LM_H, 24, 4, 3);
        } else if (pl.l2_cfg != 1) {
            if (mode == LM_CI_H && pl.l2_cfg == 2) S2MV_L2_LAUNCH(LM_CI_H, 12, 4, 3);
            else if (mode == LM_CI_H) S2MV_L2_LAUNCH(LM_CI_H, 12, 8, 3);
            else if (mode == LM_V) S2MV_L2_LAUNCH(LM_V, 16, 6, 3);
            else if (mode == LM_H_WTA) S2MV_L2_LAUNCH(LM_H_WTA, 16, 6, 3);
            else S2MV_L2_LAUNCH(LM_H, 16, 6, 3);
        } else {
            if (mode == LM_CI_H) S2MV_L2_LAUNCH(LM_CI_H, 16, 4, 4);
            else if (mode == LM_V) S2MV_L2_LAUNCH(LM_V, 16, 4, 4);
            else if (mode == LM_H_WTA) S2MV_L2_LAUNCH(LM_H_WTA, 16, 4, 4);
            else S2MV_L2_LAUNCH(LM_H, 16, 4, 4);
        }
#undef S2MV_L2_LAUNCH
        KCHECK();
        c->launches += 1;
        return S2MV_OK;
    }
    if (pass == 1 || pass == 4) {
        const int S = pass == 1 ? pl.S_h : pl.S_h4;
        const dim3 gh((W + S - 1) / S, rows, nviews * a.nchunks);
        a.S = S;
        a.ln_first = rr.own0;
        if (pass == 1) {
            if (from_ci) TRY(launch_line<LM_CI_H>(pl, gh, pl.smem_line_ci, st, a));
            else TRY(launch_line<LM_H>(pl, gh, pl.smem_line_h, st, a));
        } else {
            if (to_wta) TRY(launch_line<LM_H_WTA>(pl, gh, pl.smem_line_h4, st, a));
            else TRY(launch_line<LM_H>(pl, gh, pl.smem_line_h4, st, a));
        }
    } else {
        const dim3 gv((rows + pl.S_v - 1) / pl.S_v, W, nviews * a.nchunks);
        a.S = pl.S_v;
        a.ln_first = 0;
        TRY(launch_line<LM_V>(pl, gv, pl.smem_line_v, st, a));
    }
    c->launches += 1;
    return S2MV_OK;
}

static int launch_aggregate(s2mv_ctx *c, const LineArgs &a, float4 *A, float4 *B, size_t view_stride4, int nviews,
                            bool from_ci, bool to_wta, cudaStream_t st)
{
    const RowRange rr = {0, a.H, 0, a.H};
    if (c->timing) CU(cudaEventRecord(c->kev[0], st));
    for (int pass = 1; pass <= 4; ++pass) {
        TRY(launch_pass(c, a, pass, A, B, view_stride4, nviews, from_ci, to_wta, rr, st));
        if (c->timing) CU(cudaEventRecord(c->kev[pass], st));
    }
    return S2MV_OK;
}

// The two vertical passes in one launch (k_line_vv): reads volume buffer A (pass 1's output) through its tensor map,
// writes buffer B.  Whole-frame contexts only (a row band exchanges pass 2's output with its neighbours).
// rows: the rows buffer A holds, as image rows [vlo, vhi) of `a`, and the rows [own0, own1) to store (a whole image: 0, H, 0, H).
static int launch_vv(s2mv_ctx *c, LineArgs a, float4 *Bout, size_t view_stride4, int nviews, const RowRange &rr, cudaStream_t st)
{
    const CostPlan &pl = c->plan;
    LineVVArgs L;
    memset(&L, 0, sizeof(L));
    for (int v = 0; v < nviews; ++v) {
        a.out[v] = Bout + v * view_stride4;                 // row 0 of the launch = image row vlo = the buffers' first row
        a.arms[v] += (size_t)rr.vlo * a.W;
    }
    a.H = rr.vhi - rr.vlo;
    L.out_r0 = rr.own0 - rr.vlo;
    L.out_r1 = rr.own1 - rr.vlo;
    L.a = a;
    L.S = pl.vv_S; L.HP = pl.vv_HP; L.P = L.S + 2 * L.HP;
    L.tiles_per_col = (a.H + L.S - 1) / L.S;
    L.ncols = a.W;
    L.nz = nviews * a.nchunks;
    L.nclaims = L.nz * L.ncols;
    L.counter = c->line_ctr + 5;
    L.tmap_rows = c->tmap_rows_vv;
    CU(cudaMemsetAsync(L.counter, 0, sizeof(int), st));
    const int grid = std::min(c->sm_count, L.nclaims);
    k_line_vv<kVVNA, kVVB><<<grid, (2 * kVVNA + kL2Producers) * 32, pl.vv_smem, st>>>(L, c->tmap_vv);
    KCHECK();
    c->launches += 1;
    return S2MV_OK;
}

// Four-direction scanline optimisation of `nviews` aggregated volumes (slot v at cost + v * view_stride4)
// into acc (same layout), then WTA into disp[] (kernels_so.cuh; specification: DESIGN.md §3.4, held to it by tests/test_so.py).
static int launch_so(s2mv_ctx *c, const float4 *cost, float4 *acc, size_t view_stride4, float *const disp[2], int nviews,
                     int view_first, float T, float H1, float H2, int D, int zd, int H, int W, bool store_cost,
                     cudaStream_t st)
{
    const CostPlan &pl = c->plan;
    if (pl.nchunks != 1) return fail(S2MV_ERR_BAD_PARAM, "scanline optimisation supports num_disp <= 128 (got %d)", D);
    SoArgs a;
    memset(&a, 0, sizeof(a));
    for (int v = 0; v < nviews; ++v) { a.cost[v] = cost + v * view_stride4; a.acc[v] = acc + v * view_stride4; a.disp[v] = disp[v]; }
    a.pix[0] = c->pix[0]; a.pix[1] = c->pix[1];
    a.H = H; a.W = W; a.D = D; a.zd = zd; a.LPtot = pl.LPtot; a.view_first = view_first;
    a.T = T;
    a.P1[0] = H1; a.P1[1] = H1 / 4.0f; a.P1[2] = H1 / 10.0f;   // d_dc_hslo.cu:124-127
    a.P2[0] = H2; a.P2[1] = H2 / 4.0f; a.P2[2] = H2 / 10.0f;
    a.store_cost = store_cost ? 1 : 0;
    const int dirs[4][2] = {{+1, 0}, {-1, 0}, {0, +1}, {0, -1}};
    for (int k = 0; k < 4; ++k) {
        a.dx = dirs[k][0]; a.dy = dirs[k][1];
        a.first = k == 0; a.last = k == 3;
        const int nlines = a.dx ? H : W;
        k_so_dir<<<dim3((nlines + 3) / 4, nviews), 128, 0, st>>>(a);
        KCHECK();
    }
    c->launches += 4;
    return S2MV_OK;
}

// CI + H, V, V, H + WTA for both views: A <- CI+H; B <- V(A); A <- V(B); disp <- WTA(H(A))
static int launch_costvol(s2mv_ctx *c, float *dispL, float *dispR, cudaStream_t st)
{
    const s2mv_params &p = c->prm;
    const CostPlan &pl = c->plan;
    const int H = p.num_rows, W = p.num_cols;
    const size_t n = (size_t)H * W, view_stride4 = n * pl.LPtot;
    float4 *A = reinterpret_cast<float4 *>(c->vol[0]), *B = reinterpret_cast<float4 *>(c->vol[1]);
    LineArgs a;
    fill_largs(c, a, H, W, p.zero_disp, p.ad_coeff);
    for (int v = 0; v < 2; ++v) {
        a.arms[v] = c->arms[v];
        a.wta_key[v] = c->wta_key[v];
    }
    a.disp[0] = dispL; a.disp[1] = dispR;
    if (pl.nchunks > 1)
        for (int v = 0; v < 2; ++v) CU(cudaMemsetAsync(c->wta_key[v], 0xff, n * sizeof(unsigned long long), st));
    if (c->so_on) {
        // aggregated volume kept (pass 4 stores into B), scanline optimisation B -> A, WTA from there
        float *disp[2] = {dispL, dispR};
        TRY(launch_aggregate(c, a, A, B, view_stride4, 2, true, false, st));
        TRY(launch_so(c, B, A, view_stride4, disp, 2, 0, c->so_T, c->so_H1, c->so_H2, p.num_disp, p.zero_disp, H, W, false, st));
        return S2MV_OK;
    }
    // fused vertical passes: CI+H1 -> A, V2∘V3: A -> B, H4+WTA from B
    const bool fused = pl.line2 && pl.vv && c->tmap_ok && !c->env_line_v1 && !c->env_line_bulk && !c->env_no_vv;
    auto aggregate = [&](const LineArgs &aa) -> int {
        if (!fused) return launch_aggregate(c, aa, A, B, view_stride4, 2, true, true, st);
        const RowRange rr = {0, aa.H, 0, aa.H};
        if (c->timing) CU(cudaEventRecord(c->kev[0], st));
        TRY(launch_pass(c, aa, 1, A, B, view_stride4, 2, true, true, rr, st));
        if (c->timing) CU(cudaEventRecord(c->kev[1], st));
        TRY(launch_vv(c, aa, B, view_stride4, 2, rr, st));
        if (c->timing) { CU(cudaEventRecord(c->kev[2], st)); CU(cudaEventRecord(c->kev[3], st)); }
        TRY(launch_pass(c, aa, 4, /*read*/ B, /*unused*/ A, view_stride4, 2, true, true, rr, st));
        if (c->timing) CU(cudaEventRecord(c->kev[4], st));
        return S2MV_OK;
    };
    if (pl.chunk_seq) {
        for (int ch = 0; ch < pl.nchunks; ++ch) {
            a.d_first = ch * 4 * pl.LP;
            TRY(aggregate(a));
        }
    } else {
        TRY(aggregate(a));
    }
    if (pl.nchunks > 1) {
        k_wta_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->wta_key[0], dispL, p.zero_disp, n);
        k_wta_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->wta_key[1], dispR, p.zero_disp, n);
        KCHECK();
        c->launches += 2;
    }
    return S2MV_OK;
}

static int launch_dcc(s2mv_ctx *c, const float *dL, const float *dR, uint8_t *oL, uint8_t *oR, int H, int W, cudaStream_t st)
{
    const size_t n = (size_t)H * W;
    if (2 * (size_t)W <= 48 * 1024 && !c->env_dcc_split) {  // one launch, the marks of a row in shared memory
        k_dcc_row<<<H, 256, 2 * (size_t)W, st>>>(dL, dR, oL, oR, H, W);
        KCHECK();
        c->launches += 1;
        return S2MV_OK;
    }
    CU(cudaMemsetAsync(oL, 0, n, st));
    CU(cudaMemsetAsync(oR, 0, n, st));
    CU(cudaMemsetAsync(c->disoccl[0], 1, n, st));
    CU(cudaMemsetAsync(c->disoccl[1], 1, n, st));
    dim3 g((W + 255) / 256, H);
    k_dcc<<<g, 256, 0, st>>>(dL, dR, oL, oR, c->disoccl[0], c->disoccl[1], H, W);
    KCHECK();
    k_dcc_merge<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(oL, oR, c->disoccl[0], c->disoccl[1], n);
    KCHECK();
    c->launches += 2;
    return S2MV_OK;
}

// nviews = 1 or 2 views voted together; arrays indexed by view slot
static int launch_irv(s2mv_ctx *c, float *const disp[2], uint8_t *const outl[2], const uint32_t *const arms[2],
                      int nviews, int H, int W, int D, int zd, int usd, int thresh_s, float thresh_h, int iterations,
                      cudaStream_t st, int keep0 = 0, int keep1 = -1, int post_reach = 0)
{
    // keep0/keep1/post_reach (row bands): only rows [keep0, keep1) of the result are used, after filters that reach
    // post_reach rows.  Iteration k of K then only has to vote the rows within post_reach + usd * (K - 1 - k) of them:
    // a vote further out cannot travel into the kept rows in the iterations that remain (each carries usd rows).
    const size_t n = (size_t)H * W;
    IrvArgs a;
    memset(&a, 0, sizeof(a));
    for (int v = 0; v < nviews; ++v) {
        a.disp[v] = disp[v]; a.outliers[v] = outl[v]; a.arms[v] = arms[v];
        a.list[v] = c->irv_list[v]; a.next[v] = c->irv_list2[v]; a.vote[v] = c->irv_vote[v];
        a.count[v] = c->irv_count + v; a.next_count[v] = c->irv_count + 2 + v;
        a.ticket[v] = c->irv_count + 4 + v;
        a.accepted[v] = iterations <= 64 ? c->irv_count + 8 + 64 * v : nullptr;
    }
    a.H = H; a.W = W; a.nbins = D > 65 ? D : 65; a.zd = zd; a.usd = usd; a.thresh_s = thresh_s; a.thresh_h = thresh_h;
    const size_t hist_bytes = (size_t)kIrvWarps * a.nbins * sizeof(int);
    if (hist_bytes > 64 * 1024) return fail(S2MV_ERR_BAD_PARAM, "num_disp too large for the voting histogram");
    if (iterations <= 0) return S2MV_OK;
    // outliers -> list, once; every iteration then votes on its list and leaves the survivors as the next one
    CU(cudaMemsetAsync(c->irv_count, 0, (8 + 2 * 64) * sizeof(int), st));
    k_irv_compact<<<dim3((unsigned)((n + 4095) / 4096), nviews), 256, 0, st>>>(a);
    KCHECK();
    c->launches += 1;
    // lists longer than 1/32 of the image take the dense path (decided on the device, per iteration and view)
    const bool dense_ok = c->irv_hseg[0] && c->irv_nbp >= a.nbins && (size_t)H * W == (size_t)c->prm.num_rows * c->prm.num_cols;
    a.nbp = c->irv_nbp;
    // measured crossover of the two paths at 128 bins: a list of about n/28; the dense path's cost grows with
    // the bin count (bytes per pixel histogram), the sparse one's does not
    a.dense_min = (int)(n / 32) * (c->irv_nbp > 128 ? c->irv_nbp / 128 : 1) + 1;
    if (c->env_irv_dense_min >= 0) a.dense_min = c->env_irv_dense_min;  // test hook (read at create): 0 = always dense, huge = never
    for (int v = 0; v < nviews; ++v) {
        a.hseg[v] = dense_ok ? c->irv_hseg[v] : nullptr;
        a.stamp[v] = dense_ok ? c->irv_stamp[v] : nullptr;
        a.hchg[v] = dense_ok ? c->irv_hchg[v] : nullptr;
    }
    if (iterations >= 255) for (int v = 0; v < nviews; ++v) a.stamp[v] = nullptr;
    a.iterations = iterations; a.keep0 = keep0; a.keep1 = keep1; a.post_reach = post_reach;
    a.hint = c->irv_hint_d;
    if (c->env_irv_coop && iterations <= 64 && c->irv_coop_bps >= 0) {
        // All iterations in one cooperative launch when the lists are expected to be short: the previous frame's were
        // (a scene cut costs one slow frame, not a wrong one), or there is no dense path to prefer on long ones.
        bool light = c->env_irv_coop == 2 || !dense_ok;
        if (!light && c->irv_hint_h) {
            light = true;
            for (int v = 0; v < nviews; ++v) {
                const int h = ((volatile int *)c->irv_hint_h)[v];
                if (h < 0 || h >= a.dense_min) light = false;
            }
        }
        if (light) {
            if (c->irv_coop_bps == 0 || c->irv_coop_smem != hist_bytes) {
                int coop = 0, bps = 0;
                CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
                if (coop) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_irv_sparse_all, kIrvWarps * 32, hist_bytes));
                c->irv_coop_bps = bps > 0 ? bps : -1;
                c->irv_coop_smem = hist_bytes;
            }
            if (c->irv_coop_bps > 0) {
                for (int v = 0; v < nviews; ++v) { a.hseg[v] = nullptr; a.stamp[v] = nullptr; a.hchg[v] = nullptr; }
                const int gx = std::max(1, std::min(c->sm_count * 4, c->sm_count * c->irv_coop_bps / nviews));
                void *args[] = {&a};
                CU(cudaLaunchCooperativeKernel((const void *)k_irv_sparse_all, dim3(gx, nviews), dim3(kIrvWarps * 32), args,
                                               hist_bytes, st));
                c->launches += 1;
                return S2MV_OK;
            }
        }
    }
    if (dense_ok && iterations > 1 && iterations < 255)
        for (int v = 0; v < nviews; ++v) CU(cudaMemsetAsync(c->irv_stamp[v], 0, n, st));
    // the column walk hands out (column, 32-row strip) tickets, each a serial chain of up to 32 votes: it needs several
    // tickets per resident warp to balance (640x384: 15 k tickets for 9.5 k warps, 0.48 ms against 0.42 ms by list entry)
    const long long col_tickets = (long long)W * ((H + 31) / 32) * nviews;
    const bool col_forced = c->env_irv_colw_set;
    a.col_votes = dense_ok && !c->env_irv_list_votes && usd <= 64 && (col_forced || col_tickets >= 4ll * c->sm_count * 64)
                      ? c->env_irv_colw : 0;
    for (int it = 0; it < iterations; ++it) {
        a.it = it;
        a.row_lo = 0;
        a.row_hi = H;
        if (keep1 >= 0) {
            const int m = post_reach + usd * (iterations - 1 - it);
            a.row_lo = std::max(0, keep0 - m);
            a.row_hi = std::min(H, keep1 + m);
        }
        if (dense_ok) {
            {
                const dim3 gh(c->sm_count * 8, nviews);
                const size_t sh = (size_t)kHsegThreads * (a.nbp + 4);
                switch (a.nbp / 128) {
                    case 1: k_irv_hseg<1><<<gh, kHsegThreads, sh, st>>>(a); break;
                    case 2: k_irv_hseg<2><<<gh, kHsegThreads, sh, st>>>(a); break;
                    case 3: k_irv_hseg<3><<<gh, kHsegThreads, sh, st>>>(a); break;
                    default: k_irv_hseg<4><<<gh, kHsegThreads, sh, st>>>(a); break;
                }
            }
            KCHECK();
            const dim3 gd(c->sm_count * 8, nviews);
            if (a.col_votes) switch (a.nbp / 128) {
                case 1: k_irv_vote_col<1><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
                case 2: k_irv_vote_col<2><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
                case 3: k_irv_vote_col<3><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
                default: k_irv_vote_col<4><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
            }
            else switch (a.nbp / 128) {
                case 1: k_irv_vote_dense<1><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
                case 2: k_irv_vote_dense<2><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
                case 3: k_irv_vote_dense<3><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
                default: k_irv_vote_dense<4><<<gd, kIrvWarps * 32, 0, st>>>(a); break;
            }
            KCHECK();
            c->launches += 2;
        }
        k_irv_vote<<<dim3(c->sm_count * 4, nviews), kIrvWarps * 32, hist_bytes, st>>>(a);
        KCHECK();
        k_irv_apply<<<dim3(c->sm_count, nviews), 256, 0, st>>>(a);
        KCHECK();
        c->launches += 2;
        for (int v = 0; v < nviews; ++v) {
            int *t = a.list[v]; a.list[v] = a.next[v]; a.next[v] = t;
            t = a.count[v]; a.count[v] = a.next_count[v]; a.next_count[v] = t;
        }
    }
    return S2MV_OK;
}

// k_dbm4 when the planes allow vector access, k_dbm otherwise
static void launch_dbm(const DbmArgs &d, int nviews, cudaStream_t st)
{
    auto al16 = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
    if (d.W % 4 == 0 && al16(d.dispL) && al16(d.dispR) && al16(d.maskL) && al16(d.maskR) && al16(d.tmask) &&
        ((uintptr_t)d.views & 3) == 0) {
        const long long threads = (long long)(d.W / 4) * d.H;
        k_dbm4<<<dim3((unsigned)((threads + 255) / 256), 1, nviews), 256, 0, st>>>(d);
    } else {
        k_dbm<<<dim3((d.W + 255) / 256, d.H, nviews), 256, 0, st>>>(d);
    }
}

// in/out per view slot (nviews = 1 or 2).  `bounded`: the inputs are this pipeline's own disparities
// (|a - s| <= num_disp - 1), see k_bilateral4.
static int launch_bilateral(s2mv_ctx *c, const float *const in[2], float *const out[2], int nviews,
                            const float *spatial, const float *colour, int radius, int ncolour, bool bounded, int H,
                            int W, cudaStream_t st, const float *h_spatial = nullptr)
{
    if (radius == 7 && (size_t)ncolour * sizeof(float) <= 32 * 1024) {
        constexpr int R = 7, KW = 15, KWP = 16, TWP = (kBil4W + 2 * R + 3) & ~3, TH = kBil4H + 2 * R;
        const size_t smem = ((size_t)TWP * TH + KWP * KW + ncolour) * sizeof(float);
        dim3 g((W + kBil4W - 1) / kBil4W, (H + kBil4H - 1) / kBil4H, nviews);
        const bool packed = !c->env_bilateral_scalar;  // test / A-B hook (read at create): the one-output-at-a-time kernel
        if (bounded && packed && h_spatial) {
            const size_t smem2 = ((size_t)2 * TWP * TH + ncolour) * sizeof(float);
            BilPairs<R> w2;
            for (int ky = 0; ky < KW; ++ky)
                for (int kx = 0; kx < KWP; ++kx) {
                    const float w = kx < KW ? h_spatial[ky * KW + kx] : 0.0f;
                    w2.w[ky][kx] = make_float2(w, w);
                }
            k_bilateral4p<R><<<g, dim3(32, 8), smem2, st>>>(in[0], in[nviews - 1], out[0], out[nviews - 1], w2, colour,
                                                            ncolour, H, W);
        } else if (bounded)
            k_bilateral4<R, true><<<g, dim3(32, 8), smem, st>>>(in[0], in[nviews - 1], out[0], out[nviews - 1], spatial,
                                                                 colour, ncolour, H, W);
        else
            k_bilateral4<R, false><<<g, dim3(32, 8), smem, st>>>(in[0], in[nviews - 1], out[0], out[nviews - 1], spatial,
                                                                  colour, ncolour, H, W);
        KCHECK();
        c->launches += 1;
        return S2MV_OK;
    }
    const int tw = kBilW + 2 * radius, th = kBilH + 2 * radius, kw = 2 * radius + 1;
    size_t smem = ((size_t)tw * th + (size_t)kw * kw + ncolour) * sizeof(float);
    if (smem > 160 * 1024) return fail(S2MV_ERR_BAD_PARAM, "bilateral tile too large");
    dim3 g((W + kBilW - 1) / kBilW, (H + kBilH - 1) / kBilH);
    for (int v = 0; v < nviews; ++v) {
        k_bilateral<<<g, dim3(kBilW, kBilH), smem, st>>>(in[v], out[v], spatial, colour, radius, ncolour, H, W);
        KCHECK();
        c->launches += 1;
    }
    return S2MV_OK;
}

// sum of the kernel weights exactly as every thread of filter_gaussian_1_kernel_1 accumulates it
// (d_filter_gaussian.cu:60-80: row-major, fp32, from 0)
static float host_kernel_norm(const float *k, int radius)
{
    const int kw = 2 * radius + 1;
    volatile float norm = 0.0f;  // volatile: one rounded fp32 add per weight, never widened or reassociated
    for (int i = 0; i < kw * kw; ++i) norm = norm + k[i];
    return norm;
}

static int launch_gauss(s2mv_ctx *c, const float *in, float *out, const float *kernel, const float *host_kernel,
                        int radius, int invert, int H, int W, cudaStream_t st)
{
    if (radius == 10 && host_kernel) {
        constexpr int R = 10, KW = 21, KWP = 24, TWP = (kGa4W + 2 * R + 3) & ~3, TH = kGa4H + 2 * R;
        const size_t smem = ((size_t)TWP * TH + KWP * KW) * sizeof(float);
        dim3 g((W + kGa4W - 1) / kGa4W, (H + kGa4H - 1) / kGa4H);
        k_gauss_dilate4<R><<<g, dim3(32, 8), smem, st>>>(in, out, kernel, host_kernel_norm(host_kernel, radius), invert,
                                                          H, W);
        KCHECK();
        c->launches += 1;
        return S2MV_OK;
    }
    const int tw = kGaW + 2 * radius, th = kGaH + 2 * radius, kw = 2 * radius + 1;
    size_t smem = ((size_t)tw * th + (size_t)kw * kw) * sizeof(float);
    dim3 g((W + kGaW - 1) / kGaW, (H + kGaH - 1) / kGaH);
    k_gauss_dilate<<<g, dim3(kGaW, kGaH), smem, st>>>(in, out, kernel, radius, invert, H, W);
    KCHECK();
    c->launches += 1;
    return S2MV_OK;
}

static int launch_mux(s2mv_ctx *c, const uint8_t *const *views, uint8_t *out, int V, float angle, int Hin, int Win,
                      int Hout, int Wout, int elem_sz, int variant, cudaStream_t st, int row0 = 0, int Hframe_in = 0,
                      int Hframe_out = 0)
{
    MuxArgs m;
    memset(&m, 0, sizeof(m));
    for (int v = 0; v < V; ++v) m.views[v] = views[v];
    m.out = out; m.num_views = V; m.Hin = Hin; m.Win = Win; m.Hout = Hout; m.Wout = Wout;
    m.row0 = row0; m.Hframe_in = Hframe_in ? Hframe_in : Hin; m.Hframe_out = Hframe_out ? Hframe_out : Hout;
    // d_mux_multiview.cu:146 (PI is the float literal 3.1415926535f)
    float yi = (float)((double)(float)V / tan((double)(angle * 3.1415926535f) / 180.0) / (double)(float)elem_sz);
    m.y_interval = yi;
    m.inv_y_interval = 1.0f / yi;
    m.rint_y = (int)roundf(yi);
    if (m.rint_y == 0) return fail(S2MV_ERR_BAD_PARAM, "interlace period rounds to zero");
    m.variant = variant ? variant : ((Hout % V == 0) ? 2 : 1);  // d_mux_multiview.cu:183-191
    dim3 g((Wout + 255) / 256, Hout);
    k_mux<<<g, 256, 0, st>>>(m);
    KCHECK();
    c->launches += 1;
    return S2MV_OK;
}

// --------------------------------------------------------- frame pipeline
// Refinement (cross-check, region voting, bilateral; d_io.cu:139-151) and DIBR + interlace
// (d_io.cu:160-236) from the WTA disparities in c->disp[].  Records timing events 3 and 4.
static int run_dibr(s2mv_ctx *c, const float *fl, const float *fr, uint8_t *d_interlaced, cudaStream_t st);
static int run_refine(s2mv_ctx *c, float *fl, float *fr, cudaStream_t st);

static int run_refine_dibr(s2mv_ctx *c, float *d_disp_l, float *d_disp_r, uint8_t *d_interlaced, cudaStream_t st)
{
    float *fl = d_disp_l ? d_disp_l : c->dispF[0], *fr = d_disp_r ? d_disp_r : c->dispF[1];
    TRY(run_refine(c, fl, fr, st));
    CU(cudaEventRecord(c->ev_refined, st));
    if (c->timing) CU(cudaEventRecord(c->ev[3], st));
    return run_dibr(c, fl, fr, d_interlaced, st);
}

// cross-check, region voting, bilateral (d_io.cu:139-151): c->disp[] -> fl / fr
static int run_refine(s2mv_ctx *c, float *fl, float *fr, cudaStream_t st)
{
    const s2mv_params &p = c->prm;
    const int H = p.num_rows, W = p.num_cols;
    const size_t n = (size_t)H * W;
    if (c->taps)
        for (int v = 0; v < 2; ++v)
            CU(cudaMemcpyAsync(c->tap_wta[v], c->disp[v], n * sizeof(float), cudaMemcpyDeviceToDevice, st));

    // refinement: cross-check, region voting, bilateral (d_io.cu:139-151)
    TRY(launch_dcc(c, c->disp[0], c->disp[1], c->outl[0], c->outl[1], H, W, st));
    if (c->taps)
        for (int v = 0; v < 2; ++v) CU(cudaMemcpyAsync(c->tap_outl[v], c->outl[v], n, cudaMemcpyDeviceToDevice, st));
    {
        float *disp[2] = {c->disp[0], c->disp[1]};
        uint8_t *outl[2] = {c->outl[0], c->outl[1]};
        const uint32_t *arms[2] = {c->arms[0], c->arms[1]};
        // a row band keeps its own rows only; what follows the voting reaches bilateral + bleed + mask blur rows
        const int post = p.bilateral_radius + 1 + p.mask_blur_radius;
        TRY(launch_irv(c, disp, outl, arms, 2, H, W, p.num_disp, p.zero_disp, p.usd, p.thresh_s, p.thresh_h,
                       p.irv_iterations, st, c->band ? c->band_o0 : 0, c->band ? c->band_o1 : -1, post));
    }
    if (c->taps)
        for (int v = 0; v < 2; ++v)
            CU(cudaMemcpyAsync(c->tap_irv[v], c->disp[v], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    {
        const float *bin[2] = {c->disp[0], c->disp[1]};
        float *bout[2] = {fl, fr};
        TRY(launch_bilateral(c, bin, bout, 2, c->bil_spatial, c->bil_colour, p.bilateral_radius, p.num_disp, true, H, W,
                             st, c->h_bil_spatial.data()));
    }
    return S2MV_OK;
}

// DIBR + interlace (d_io.cu:160-236) from the refined disparities fl / fr.  Records timing event 4.
static int run_dibr(s2mv_ctx *c, const float *fl, const float *fr, uint8_t *d_interlaced, cudaStream_t st)
{
    const s2mv_params &p = c->prm;
    const int H = p.num_rows, W = p.num_cols, V = p.num_views;
    const size_t n = (size_t)H * W;
    // DIBR (d_io.cu:160-191)
    if (H >= 3 && W >= 2 && 6 * (size_t)W <= 48 * 1024 && !c->env_dcc_split) {
        // occlusion marks, bleed and mask conversion of both views, one row per block (marks in shared memory)
        k_occl_bleed_mask_row<<<H, 256, 6 * (size_t)W, st>>>(fl, fr, c->mask[0], c->mask[1], H, W);
        KCHECK();
        c->launches += 1;
    } else {
        CU(cudaMemsetAsync(c->occl[0], 0, n, st));
        CU(cudaMemsetAsync(c->occl[1], 0, n, st));
        dim3 g((W + 255) / 256, H);
        k_occl<<<g, 256, 0, st>>>(fl, fr, c->occl[0], c->occl[1], H, W);
        KCHECK();
        for (int v = 0; v < 2; ++v) {
            k_bleed<<<g, 256, 0, st>>>(c->occl[v], c->occlB[v], c->mask[v], 1, H, W);
            KCHECK();
        }
        c->launches += 3;
    }
    TRY(launch_gauss(c, c->mask[1], c->tmask, c->gauss_kernel, c->h_gauss_kernel.data(), p.mask_blur_radius, 1, H, W,
                     st));
    if (V > 2) {
        DbmArgs d;
        memset(&d, 0, sizeof(d));
        d.pixL = c->pix[0]; d.pixR = c->pix[1]; d.dispL = fl; d.dispR = fr;
        d.maskL = c->mask[0]; d.maskR = c->mask[1]; d.tmask = c->tmask; d.views = c->views; d.H = H; d.W = W;
        for (int v = 1; v < V - 1; ++v) {
            // d_io.cu:187: float shift = 1.0 - ((1.0 * (float) v) / ((float) num_views - 1.0));
            d.shift[v - 1] = (float)(1.0 - ((1.0 * (double)(float)v) / ((double)(float)V - 1.0)));
            d.view_index[v - 1] = v;
        }
        launch_dbm(d, V - 2, st);
        KCHECK();
        c->launches += 1;
    }
    const uint8_t *vp[16];
    for (int v = 0; v < V; ++v) vp[v] = c->views + (size_t)v * n * 3;
    TRY(launch_mux(c, vp, d_interlaced ? d_interlaced : c->interlaced, V, (float)p.angle, H, W, p.num_rows_out,
                   p.num_cols_out, p.elem_sz, 2, st, c->band ? c->band_ly0 : 0, c->band ? c->band_frame_rows : 0,
                   c->band ? c->band_frame_rows : 0));
    if (c->timing) CU(cudaEventRecord(c->ev[4], st));
    return S2MV_OK;
}

static int run_frame(s2mv_ctx *c, const uint8_t *d_sbs, int num_cols_sbs, float *d_disp_l, float *d_disp_r,
                     uint8_t *d_interlaced, bool costvol_only, cudaStream_t st)
{
    const s2mv_params &p = c->prm;
    const int H = p.num_rows, W = p.num_cols, V = p.num_views;
    const size_t n = (size_t)H * W;
    if (num_cols_sbs < 2 * W) return fail(S2MV_ERR_BAD_PARAM, "num_cols_sbs (%d) < 2*num_cols (%d)", num_cols_sbs, 2 * W);
    if (c->band) return fail(S2MV_ERR_BAD_PARAM, "this context is a row band: use the s2mv_band_* sequence");
    if (c->no_volume) return fail(S2MV_ERR_BAD_PARAM, "this context is two-resolution (s2mv_configure_2): use s2mv_process_sbs_2*");
    c->launches = 0;
    if (c->timing) CU(cudaEventRecord(c->ev[0], st));
    // views[0] = right image, views[V-1] = left image (d_io.cu:181-182)
    uint8_t *view0 = c->views, *viewN = c->views + (size_t)(V - 1) * n * 3;
    TRY(launch_prepare(c, d_sbs, d_sbs + (size_t)W * 3, (size_t)num_cols_sbs * 3, costvol_only ? nullptr : viewN,
                       costvol_only ? nullptr : view0, st));
    TRY(build_luts(c, p.ad_coeff, p.census_coeff, st));
    if (c->timing) CU(cudaEventRecord(c->ev[1], st));

    float *wl = costvol_only && d_disp_l ? d_disp_l : c->disp[0];
    float *wr = costvol_only && d_disp_r ? d_disp_r : c->disp[1];
    TRY(launch_costvol(c, wl, wr, st));
    if (c->timing) CU(cudaEventRecord(c->ev[2], st));
    if (costvol_only) {
        if (c->timing) {
            CU(cudaEventRecord(c->ev[3], st));
            CU(cudaEventRecord(c->ev[4], st));
        }
        return S2MV_OK;
    }
    return run_refine_dibr(c, d_disp_l, d_disp_r, d_interlaced, st);
}

extern "C" int s2mv_process_sbs_device(s2mv_ctx *c, const uint8_t *d_img_sbs, int num_cols_sbs, float *d_disp_l,
                                       float *d_disp_r, uint8_t *d_interlaced, void *stream)
{
    if (!c || !d_img_sbs) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (!c->configured) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_configure first");
    CU(cudaSetDevice(c->device));
    return run_frame(c, d_img_sbs, num_cols_sbs, d_disp_l, d_disp_r, d_interlaced, false,
                     stream ? (cudaStream_t)stream : c->stream);
}

extern "C" int s2mv_costvol_device(s2mv_ctx *c, const uint8_t *d_img_sbs, int num_cols_sbs, float *d_disp_l,
                                   float *d_disp_r, void *stream)
{
    if (!c || !d_img_sbs) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (!c->configured) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_configure first");
    CU(cudaSetDevice(c->device));
    return run_frame(c, d_img_sbs, num_cols_sbs, d_disp_l, d_disp_r, nullptr, true,
                     stream ? (cudaStream_t)stream : c->stream);
}

static int run_frame_2(s2mv_ctx *c, const uint8_t *d_sbs, int num_cols_sbs, float *d_disp_l, float *d_disp_r,
                       uint8_t *d_interlaced, cudaStream_t st);

static void host_unregister_all(s2mv_ctx *c)
{
    for (auto &r : c->host_regs) cudaHostUnregister(const_cast<void *>(r.first));
    cudaGetLastError();
    c->host_regs.clear();
}

// Page-lock a pageable caller buffer in place so that it can be DMA'd without the staging copy.  The
// reference's video driver hands the same cv::Mat buffers to adcensus_stm for its whole run
// (video_io.cpp:125-137), so this happens once per buffer.  Returns false (-> staging) when the range cannot
// be registered; a stale overlapping registration (the caller freed and reallocated) is dropped and retried.
static bool host_register(s2mv_ctx *c, const void *ptr, size_t bytes)
{
    for (auto &r : c->host_regs)
        if (r.first == ptr && r.second >= bytes) return true;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterDefault);
        if (e == cudaSuccess) {
            if (c->host_regs.size() >= 16) {  // keep the table small: drop the oldest
                cudaHostUnregister(const_cast<void *>(c->host_regs.front().first));
                c->host_regs.erase(c->host_regs.begin());
            }
            c->host_regs.emplace_back(ptr, bytes);
            return true;
        }
        cudaGetLastError();
        if (e != cudaErrorHostMemoryAlreadyRegistered) return false;
        // overlaps something registered earlier (by this table or by the caller): forget ours that overlap, retry once
        const char *lo = (const char *)ptr, *hi = lo + bytes;
        bool dropped = false;
        for (size_t i = 0; i < c->host_regs.size();) {
            const char *a = (const char *)c->host_regs[i].first, *b = a + c->host_regs[i].second;
            if (a < hi && lo < b) {
                cudaHostUnregister(const_cast<void *>(c->host_regs[i].first));
                c->host_regs.erase(c->host_regs.begin() + i);
                dropped = true;
            } else ++i;
        }
        cudaGetLastError();
        if (!dropped) return false;
    }
    return false;
}

extern "C" int s2mv_set_host_registration(s2mv_ctx *c, int mode)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (mode < 0 || mode > 2) return fail(S2MV_ERR_BAD_PARAM, "mode must be 0 (staged), 1 (always) or 2 (auto)");
    CU(cudaSetDevice(c->device));
    if (mode != c->host_reg) {
        CU(cudaStreamSynchronize(c->stream));
        host_unregister_all(c);
        for (auto &q : c->host_last) q = nullptr;
    }
    c->host_reg = mode;
    return S2MV_OK;
}

// host buffers in / out, synchronous: s2mv_process_sbs (adcensus_stm) and s2mv_process_sbs_2 (adcensus_stm_2)
static int process_host(s2mv_ctx *c, const uint8_t *img_sbs, int num_cols_sbs, float *disp_l, float *disp_r,
                        uint8_t *interlaced, bool two_res)
{
    if (!c || !img_sbs) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (!c->configured) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_configure first");
    if (two_res && !c->lo) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_configure_2 first");
    CU(cudaSetDevice(c->device));
    const s2mv_params &p = c->prm;
    const size_t n = (size_t)p.num_rows * p.num_cols;
    const size_t sbs_bytes = (size_t)p.num_rows * num_cols_sbs * 3;
    const size_t out_bytes = (size_t)p.num_rows_out * p.num_cols_out * 3;
    if (c->sbs_bytes < sbs_bytes) {
        if (c->sbs) {  // regrown (a wider num_cols_sbs): release the old buffer now, not at the next configure
            CU(cudaStreamSynchronize(c->stream));
            c->allocs.erase(std::remove(c->allocs.begin(), c->allocs.end(), (void *)c->sbs), c->allocs.end());
            cudaFree(c->sbs);
            c->arena_bytes -= c->sbs_bytes;
            c->sbs = nullptr;
            c->sbs_bytes = 0;
        }
        void *d = nullptr;
        CU(cudaMalloc(&d, sbs_bytes));
        c->allocs.push_back(d);
        c->arena_bytes += sbs_bytes;
        c->sbs = (uint8_t *)d;
        c->sbs_bytes = sbs_bytes;
    }
    // Pageable caller buffers (cv::Mat::data in video_io.cpp:139-146) are staged through
    // pinned memory so the copies run at full PCIe/NVLink-C2C rate and overlap nothing else.
    if (c->h_sbs_bytes < sbs_bytes) {
        if (c->h_sbs) cudaFreeHost(c->h_sbs);
        CU(cudaMallocHost((void **)&c->h_sbs, sbs_bytes));
        c->h_sbs_bytes = sbs_bytes;
    }
    if (!c->h_interlaced) CU(cudaMallocHost((void **)&c->h_interlaced, out_bytes));
    for (int v = 0; v < 2; ++v)
        if (!c->h_disp[v]) CU(cudaMallocHost((void **)&c->h_disp[v], n * sizeof(float)));
    cudaStream_t st = c->stream;
    // Pinned caller buffers are DMA'd directly; pageable ones (cv::Mat::data in video_io.cpp:139-146)
    // are staged through the context's pinned buffers.
    auto is_pinned = [](const void *ptr) {
        if (!ptr) return false;
        cudaPointerAttributes attr;
        bool ok = cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        return ok;
    };
    bool pin_in = is_pinned(img_sbs), pin_dl = is_pinned(disp_l), pin_dr = is_pinned(disp_r), pin_out = is_pinned(interlaced);
    if (c->host_reg) {  // pageable buffers are page-locked in place once and DMA'd directly from then on
        // mode 1: every buffer, the first time it is seen.  mode 2 (auto, what the adcensus_stm shim runs): a
        // buffer is page-locked when the SAME pointer arrives in two consecutive calls -- the pattern of the
        // reference's video loop, which reuses its four cv::Mat buffers for the whole run (video_io.cpp:125-158)
        // -- and a registered buffer that stops arriving is unregistered at once (its owner may free it).
        const void *now[4] = {img_sbs, disp_l, disp_r, interlaced};
        const size_t nb[4] = {sbs_bytes, n * sizeof(float), n * sizeof(float), out_bytes};
        bool *pin[4] = {&pin_in, &pin_dl, &pin_dr, &pin_out};
        if (c->host_reg == 2)
            for (size_t i = 0; i < c->host_regs.size();) {
                const void *r = c->host_regs[i].first;
                if (r != now[0] && r != now[1] && r != now[2] && r != now[3]) {
                    cudaHostUnregister(const_cast<void *>(r));
                    cudaGetLastError();
                    c->host_regs.erase(c->host_regs.begin() + i);
                } else ++i;
            }
        for (int i = 0; i < 4; ++i) {
            if (!now[i] || *pin[i]) continue;
            if (c->host_reg == 1 || now[i] == c->host_last[i]) *pin[i] = host_register(c, now[i], nb[i]);
        }
        for (int i = 0; i < 4; ++i) c->host_last[i] = now[i];
    }
    if (pin_in) {
        CU(cudaMemcpyAsync(c->sbs, img_sbs, sbs_bytes, cudaMemcpyHostToDevice, st));
    } else {
        memcpy(c->h_sbs, img_sbs, sbs_bytes);
        CU(cudaMemcpyAsync(c->sbs, c->h_sbs, sbs_bytes, cudaMemcpyHostToDevice, st));
    }
    if (two_res) TRY(run_frame_2(c, c->sbs, num_cols_sbs, c->dispF[0], c->dispF[1], c->interlaced, st));
    else TRY(run_frame(c, c->sbs, num_cols_sbs, c->dispF[0], c->dispF[1], c->interlaced, false, st));
    // the disparity maps are final before DIBR starts: their D2H runs on a second stream underneath it
    CU(cudaStreamWaitEvent(c->st_aux, c->ev_refined, 0));
    if (disp_l) CU(cudaMemcpyAsync(pin_dl ? disp_l : c->h_disp[0], c->dispF[0], n * sizeof(float), cudaMemcpyDeviceToHost, c->st_aux));
    if (disp_r) CU(cudaMemcpyAsync(pin_dr ? disp_r : c->h_disp[1], c->dispF[1], n * sizeof(float), cudaMemcpyDeviceToHost, c->st_aux));
    if (interlaced) CU(cudaMemcpyAsync(pin_out ? interlaced : c->h_interlaced, c->interlaced, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(c->st_aux));
    CU(cudaStreamSynchronize(st));
    if (disp_l && !pin_dl) memcpy(disp_l, c->h_disp[0], n * sizeof(float));
    if (disp_r && !pin_dr) memcpy(disp_r, c->h_disp[1], n * sizeof(float));
    if (interlaced && !pin_out) memcpy(interlaced, c->h_interlaced, out_bytes);
    return S2MV_OK;
}

extern "C" int s2mv_process_sbs(s2mv_ctx *c, const uint8_t *img_sbs, int num_cols_sbs, float *disp_l, float *disp_r,
                                uint8_t *interlaced)
{
    return process_host(c, img_sbs, num_cols_sbs, disp_l, disp_r, interlaced, false);
}

extern "C" int s2mv_process_sbs_2(s2mv_ctx *c, const uint8_t *img_sbs, int num_cols_sbs, float *disp_l, float *disp_r,
                                  uint8_t *interlaced)
{
    return process_host(c, img_sbs, num_cols_sbs, disp_l, disp_r, interlaced, true);
}

extern "C" int s2mv_get_exp_tables(s2mv_ctx *c, float ad_coeff, float census_coeff, float *lut_ad, float *lut_cen)
{
    if (!c || !lut_ad || !lut_cen) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    CU(cudaSetDevice(c->device));
    float *d = nullptr;
    CU(cudaMalloc((void **)&d, (768 + 68) * sizeof(float)));
    float inv_ad = (float)(1.0 / ad_coeff), inv_cen = (float)(1.0 / census_coeff);
    k_build_luts<<<(kAdLutSize + 255) / 256, 256, 0, c->stream>>>(inv_ad, inv_cen, d, d + 768);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(lut_ad, d, kAdLutSize * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lut_cen, d + 768, kCenLutSize * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(S2MV_ERR_CUDA, "exp tables: %s", cudaGetErrorString(e));
    return S2MV_OK;
}

__global__ void k_ad_terms(float inv_ad, float *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kAdLutSize) out[i] = ad_term(i, inv_ad);
}

extern "C" int s2mv_get_ad_terms(s2mv_ctx *c, float ad_coeff, float *ad_terms)
{
    if (!c || !ad_terms) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (ad_coeff == 0.f) return fail(S2MV_ERR_BAD_PARAM, "ad_coeff must be non-zero");
    CU(cudaSetDevice(c->device));
    float *d = nullptr;
    CU(cudaMalloc((void **)&d, 768 * sizeof(float)));
    k_ad_terms<<<(kAdLutSize + 255) / 256, 256, 0, c->stream>>>((float)(1.0 / ad_coeff), d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(ad_terms, d, kAdLutSize * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(S2MV_ERR_CUDA, "ad terms: %s", cudaGetErrorString(e));
    return S2MV_OK;
}

extern "C" int s2mv_read_taps(s2mv_ctx *c, float *wta_l, float *wta_r, uint8_t *outliers_l, uint8_t *outliers_r,
                              float *irv_l, float *irv_r, uint8_t *arms_l, uint8_t *arms_r, float *mask_l,
                              float *mask_r, uint8_t *views)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (!c->configured) return fail(S2MV_ERR_NOT_CONFIGURED, "not configured");
    if (!c->taps) return fail(S2MV_ERR_BAD_PARAM, "taps not enabled (s2mv_enable_taps)");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const s2mv_params &p = c->prm;
    const size_t n = (size_t)p.num_rows * p.num_cols;
    float *wta[2] = {wta_l, wta_r}, *irv[2] = {irv_l, irv_r}, *mask[2] = {mask_l, mask_r};
    uint8_t *outl[2] = {outliers_l, outliers_r}, *arms[2] = {arms_l, arms_r};
    for (int v = 0; v < 2; ++v) {
        if (wta[v]) CU(cudaMemcpy(wta[v], c->tap_wta[v], n * sizeof(float), cudaMemcpyDeviceToHost));
        if (irv[v]) CU(cudaMemcpy(irv[v], c->tap_irv[v], n * sizeof(float), cudaMemcpyDeviceToHost));
        if (outl[v]) CU(cudaMemcpy(outl[v], c->tap_outl[v], n, cudaMemcpyDeviceToHost));
        if (mask[v]) CU(cudaMemcpy(mask[v], c->mask[v], n * sizeof(float), cudaMemcpyDeviceToHost));
        if (arms[v]) {
            uint8_t *d = nullptr;
            CU(cudaMalloc((void **)&d, 4 * n));
            k_arms_unpack<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->arms[v], d, n);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(arms[v], d, 4 * n, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            cudaFree(d);
            if (e != cudaSuccess) return fail(S2MV_ERR_CUDA, "arms tap: %s", cudaGetErrorString(e));
        }
    }
    if (views) CU(cudaMemcpy(views, c->views, (size_t)p.num_views * n * 3, cudaMemcpyDeviceToHost));
    return S2MV_OK;
}

#include "s2mv_stages.inl"
#include "s2mv_stream.inl"
#include "s2mv_band.inl"
#include "s2mv_halfres.inl"
