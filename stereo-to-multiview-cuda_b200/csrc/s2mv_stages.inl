// s2mv_stages.inl — host-pointer per-stage entry points (image_io.cpp:171-292
// call them in sequence).  Each uploads its inputs, runs the SAME kernels the
// frame path uses, and downloads.  Included at the end of s2mv_api.cu.

namespace {

// temp device buffer, freed on scope exit
struct Tmp {
    void *p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { CU(cudaMalloc(&p, bytes ? bytes : 16)); return S2MV_OK; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

s2mv_ctx *g_stage_ctx = nullptr;

// A context whose arena matches (H, W, D, zd, usd): the caller's if it does,
// else a process-wide one, reconfigured when the shape changes.
int acquire(s2mv_ctx *user, int H, int W, int D, int zd, int usd, s2mv_ctx **out)
{
    auto matches = [&](const s2mv_ctx *c) {
        return c && c->configured && !c->band && !c->no_volume && c->prm.num_rows == H && c->prm.num_cols == W && c->prm.num_disp == D &&
               c->prm.zero_disp == zd && c->prm.usd == usd;
    };
    if (matches(user)) { *out = user; return S2MV_OK; }
    if (!g_stage_ctx) TRY(s2mv_create(&g_stage_ctx, user ? user->device : 0));
    if (!matches(g_stage_ctx)) {
        s2mv_params p;
        s2mv_default_params(&p);
        p.num_rows = p.num_rows_out = H;
        p.num_cols = p.num_cols_out = W;
        p.num_disp = D; p.zero_disp = zd; p.usd = usd;
        TRY(s2mv_configure(g_stage_ctx, &p));
    }
    CU(cudaSetDevice(g_stage_ctx->device));
    *out = g_stage_ctx;
    return S2MV_OK;
}

// the stage entry points exchange whole cost volumes with the host: they need them resident
int need_full_volume(const s2mv_ctx *c)
{
    if (c->plan.chunk_seq)
        return fail(S2MV_ERR_OOM, "the full %d-disparity cost volumes do not fit this device (chunk-sequential arena); "
                                  "the per-stage entry points need them resident", c->plan.D);
    return S2MV_OK;
}

int upload(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return S2MV_OK;
}
int download(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    return S2MV_OK;
}

// cost initialisation of both views into vol[0] in one of the three CI modes, then planes out
template <int MODE>
int stage_ci(s2mv_ctx *user, const uint8_t *img_l, const uint8_t *img_r, float **cost_l, float **cost_r,
             float ad_coeff, float census_coeff, int D, int zd, int H, int W, int elem_sz)
{
    if (!img_l || !img_r || !cost_l || !cost_r) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3");
    s2mv_ctx *c;
    TRY(acquire(user, H, W, D, zd, user && user->configured ? user->prm.usd : 17, &c));
    TRY(need_full_volume(c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    const CostPlan &pl = c->plan;
    Tmp imgs, planes;
    TRY(imgs.alloc(2 * n * 3));
    TRY(planes.alloc((size_t)D * n * sizeof(float)));
    TRY(upload(imgs.as<uint8_t>(), img_l, n * 3, st));
    TRY(upload(imgs.as<uint8_t>() + n * 3, img_r, n * 3, st));
    dim3 g((W + 255) / 256, H);
    k_unpack<<<g, 256, 0, st>>>(imgs.as<uint8_t>(), imgs.as<uint8_t>() + n * 3, (size_t)W * 3, c->pix[0], c->pix[1],
                                c->gray[0], c->gray[1], nullptr, nullptr, H, W);
    KCHECK();
    TRY(launch_census(c, c->gray[0], c->gray[1], c->cen[0], c->cen[1], 2, H, W, st));
    TRY(build_luts(c, ad_coeff, census_coeff, st));
    HArgs a;
    fill_hargs(c, a, H, W, zd);
    float4 *A = reinterpret_cast<float4 *>(c->vol[0]);
    for (int v = 0; v < 2; ++v) {
        a.out[v] = A + (size_t)v * n * pl.LPtot;
        a.arms[v] = c->arms[v];
    }
    a.S = pl.S_ci;
    dim3 g1((W + pl.S_ci - 1) / pl.S_ci, H, 2 * pl.nchunks);
    k_hpass<MODE, false, true, false><<<g1, kHThreads, pl.smem_ci, st>>>(a);
    KCHECK();
    for (int v = 0; v < 2; ++v) {
        dim3 gt((unsigned)((n + 31) / 32), (pl.Dp + 31) / 32);
        k_vol_to_planes<<<gt, dim3(32, 8), 0, st>>>(c->vol[0] + (size_t)v * n * pl.Dp, planes.as<float>(), D, pl.Dp, n);
        KCHECK();
        float **dst = v ? cost_r : cost_l;
        for (int d = 0; d < D; ++d) TRY(download(dst[d], planes.as<float>() + (size_t)d * n, n * sizeof(float), st));
        CU(cudaStreamSynchronize(st));
    }
    return S2MV_OK;
}

}  // namespace

extern "C" int s2mv_ci_adcensus(s2mv_ctx *ctx, const uint8_t *img_l, const uint8_t *img_r, float **cost_l,
                                float **cost_r, float ad_coeff, float census_coeff, int num_disp, int zero_disp,
                                int num_rows, int num_cols, int elem_sz)
{
    return stage_ci<1>(ctx, img_l, img_r, cost_l, cost_r, ad_coeff, census_coeff, num_disp, zero_disp, num_rows,
                       num_cols, elem_sz);
}
extern "C" int s2mv_ci_ad(s2mv_ctx *ctx, const uint8_t *img_l, const uint8_t *img_r, float **cost_l, float **cost_r,
                          int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz)
{
    return stage_ci<2>(ctx, img_l, img_r, cost_l, cost_r, 10.f, 30.f, num_disp, zero_disp, num_rows, num_cols, elem_sz);
}
extern "C" int s2mv_ci_census(s2mv_ctx *ctx, const uint8_t *img_l, const uint8_t *img_r, float **cost_l,
                              float **cost_r, int num_disp, int zero_disp, int num_rows, int num_cols, int elem_sz)
{
    return stage_ci<3>(ctx, img_l, img_r, cost_l, cost_r, 10.f, 30.f, num_disp, zero_disp, num_rows, num_cols, elem_sz);
}

extern "C" int s2mv_gray(s2mv_ctx *ctx, const uint8_t *img, uint8_t *gray, int H, int W, int elem_sz)
{
    if (!img || !gray) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, 64, 32, 17, &c));
    const size_t n = (size_t)H * W;
    Tmp in, out;
    TRY(in.alloc(n * 3));
    TRY(out.alloc(n));
    TRY(upload(in.p, img, n * 3, c->stream));
    k_gray<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(in.as<uint8_t>(), out.as<uint8_t>(), n);
    KCHECK();
    TRY(download(gray, out.p, n, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return S2MV_OK;
}

extern "C" int s2mv_census(s2mv_ctx *ctx, const uint8_t *gray, uint64_t *census, int H, int W)
{
    if (!gray || !census) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, 64, 32, 17, &c));
    const size_t n = (size_t)H * W;
    Tmp in, out;
    TRY(in.alloc(n));
    TRY(out.alloc(n * 8));
    TRY(upload(in.p, gray, n, c->stream));
    k_census<true, unsigned long long><<<dim3((W + 255) / 256, H), 256, 0, c->stream>>>(
        in.as<uint8_t>(), out.as<unsigned long long>(), H, W);
    KCHECK();
    TRY(download(census, out.p, n * 8, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return S2MV_OK;
}

extern "C" int s2mv_ca_cross(s2mv_ctx *ctx, const uint8_t *img, uint8_t **cross, float **cost, float **acost,
                             float ucd, float lcd, int usd, int lsd, int D, int H, int W, int elem_sz)
{
    if (!img || !cross || !cost || !acost) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, D, D / 2, usd, &c));  // zero_disp plays no part in aggregation
    TRY(need_full_volume(c));
    cudaStream_t st = c->stream;
    const CostPlan &pl = c->plan;
    const size_t n = (size_t)H * W;
    Tmp dimg, planes, armp;
    TRY(dimg.alloc(n * 3));
    TRY(planes.alloc((size_t)D * n * sizeof(float)));
    TRY(armp.alloc(4 * n));
    TRY(upload(dimg.p, img, n * 3, st));
    for (int d = 0; d < D; ++d) TRY(upload(planes.as<float>() + (size_t)d * n, cost[d], n * sizeof(float), st));
    dim3 g((W + 255) / 256, H);
    k_unpack<<<g, 256, 0, st>>>(dimg.as<uint8_t>(), dimg.as<uint8_t>(), (size_t)W * 3, c->pix[0], c->pix[1],
                                c->gray[0], c->gray[1], nullptr, nullptr, H, W);
    KCHECK();
    TRY(launch_arms(c, c->pix[0], c->pix[0], c->arms[0], c->arms[0], 1, ucd, lcd, usd, lsd, H, W, st));
    dim3 gt((unsigned)((n + 31) / 32), (pl.Dp + 31) / 32);
    k_planes_to_vol<<<gt, dim3(32, 8), 0, st>>>(planes.as<float>(), c->vol[0], D, pl.Dp, n);
    KCHECK();
    // input planes were packed into vol[0]; the passes run vol[0] -> vol[1] -> ... and end in vol[1]
    float4 *V0 = reinterpret_cast<float4 *>(c->vol[0]), *V1 = reinterpret_cast<float4 *>(c->vol[1]);
    LineArgs a;
    fill_largs(c, a, H, W, c->prm.zero_disp, c->prm.ad_coeff);
    a.arms[0] = c->arms[0];
    // launch_aggregate reads pass 1 from its B and leaves pass 4 in its B: B = vol[0], A = vol[1]
    TRY(launch_aggregate(c, a, V1, V0, 0, 1, false, false, st));
    k_vol_to_planes<<<gt, dim3(32, 8), 0, st>>>(c->vol[0], planes.as<float>(), D, pl.Dp, n);
    KCHECK();
    for (int d = 0; d < D; ++d) TRY(download(acost[d], planes.as<float>() + (size_t)d * n, n * sizeof(float), st));
    k_arms_unpack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->arms[0], armp.as<uint8_t>(), n);
    KCHECK();
    for (int k = 0; k < 4; ++k) TRY(download(cross[k], armp.as<uint8_t>() + (size_t)k * n, n, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_dc_wta(s2mv_ctx *ctx, float **cost, float *disp, int D, int zd, int H, int W)
{
    if (!cost || !disp) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, D, zd, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    Tmp planes;
    TRY(planes.alloc((size_t)D * n * sizeof(float)));
    for (int d = 0; d < D; ++d) TRY(upload(planes.as<float>() + (size_t)d * n, cost[d], n * sizeof(float), st));
    k_wta_planes<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(planes.as<float>(), c->disp[0], D, zd, n);
    KCHECK();
    TRY(download(disp, c->disp[0], n * sizeof(float), st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

// the stage the reference declares as dc_hslo (d_dc_hslo.cu:97-101): scanline optimisation of one view's
// aggregated cost + WTA.  PARITY UNPINNED (stub in the reference); specification: DESIGN.md §3.4.
extern "C" int s2mv_dc_so(s2mv_ctx *ctx, float **cost, float *disp, float **cost_out, const uint8_t *img_own,
                          const uint8_t *img_other, int view, float T, float H1, float H2, int D, int zd, int H, int W,
                          int elem_sz)
{
    if (!cost || !img_own || !img_other || (!disp && !cost_out)) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3");
    if (view < 0 || view > 1) return fail(S2MV_ERR_BAD_PARAM, "view must be 0 (left) or 1 (right)");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, D, zd, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    TRY(need_full_volume(c));
    const CostPlan &pl = c->plan;
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    Tmp planes, d_own, d_oth;
    TRY(planes.alloc((size_t)D * n * sizeof(float)));
    TRY(d_own.alloc(n * 3));
    TRY(d_oth.alloc(n * 3));
    for (int d = 0; d < D; ++d) TRY(upload(planes.as<float>() + (size_t)d * n, cost[d], n * sizeof(float), st));
    TRY(upload(d_own.p, img_own, n * 3, st));
    TRY(upload(d_oth.p, img_other, n * 3, st));
    // pix[0] = left image, pix[1] = right image
    const uint8_t *dl = view == 0 ? d_own.as<uint8_t>() : d_oth.as<uint8_t>();
    const uint8_t *dr = view == 0 ? d_oth.as<uint8_t>() : d_own.as<uint8_t>();
    dim3 g((W + 255) / 256, H);
    k_unpack<<<g, 256, 0, st>>>(dl, dr, (size_t)W * 3, c->pix[0], c->pix[1], c->gray[0], c->gray[1], nullptr, nullptr, H, W);
    KCHECK();
    dim3 gt((unsigned)((n + 31) / 32), (pl.Dp + 31) / 32);
    k_planes_to_vol<<<gt, dim3(32, 8), 0, st>>>(planes.as<float>(), c->vol[0], D, pl.Dp, n);
    KCHECK();
    float *dv[2] = {disp ? c->disp[0] : nullptr, nullptr};
    TRY(launch_so(c, reinterpret_cast<const float4 *>(c->vol[0]), reinterpret_cast<float4 *>(c->vol[1]), 0, dv, 1, view, T,
                  H1, H2, D, zd, H, W, cost_out != nullptr, st));
    if (disp) TRY(download(disp, c->disp[0], n * sizeof(float), st));
    if (cost_out) {
        k_vol_to_planes<<<gt, dim3(32, 8), 0, st>>>(c->vol[1], planes.as<float>(), D, pl.Dp, n);
        KCHECK();
        for (int d = 0; d < D; ++d) TRY(download(cost_out[d], planes.as<float>() + (size_t)d * n, n * sizeof(float), st));
    }
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_dr_dcc(s2mv_ctx *ctx, uint8_t *outliers_l, uint8_t *outliers_r, const float *disp_l,
                           const float *disp_r, int H, int W)
{
    if (!outliers_l || !outliers_r || !disp_l || !disp_r) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    TRY(upload(c->disp[0], disp_l, n * sizeof(float), st));
    TRY(upload(c->disp[1], disp_r, n * sizeof(float), st));
    TRY(launch_dcc(c, c->disp[0], c->disp[1], c->outl[0], c->outl[1], H, W, st));
    TRY(download(outliers_l, c->outl[0], n, st));
    TRY(download(outliers_r, c->outl[1], n, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_dr_irv(s2mv_ctx *ctx, float *disp, uint8_t *outliers, uint8_t **cross, int thresh_s,
                           float thresh_h, int H, int W, int D, int zd, int usd, int iterations, int host_variant)
{
    if (!disp || !outliers || !cross) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, D, zd, usd, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    Tmp armp;
    TRY(armp.alloc(4 * n));
    for (int k = 0; k < 4; ++k) TRY(upload(armp.as<uint8_t>() + (size_t)k * n, cross[k], n, st));
    k_arms_pack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(armp.as<uint8_t>(), c->arms[0], n);
    KCHECK();
    TRY(upload(c->disp[0], disp, n * sizeof(float), st));
    TRY(upload(c->outl[0], outliers, n, st));
    // dr_irv (d_dr_irv.cu:272-366) votes once and applies the same votes `iterations`
    // times: after the first application nothing is an outlier any more (Q18)
    int iters = host_variant ? (iterations > 0 ? 1 : 0) : iterations;
    float *dv[2] = {c->disp[0], nullptr};
    uint8_t *ov[2] = {c->outl[0], nullptr};
    const uint32_t *av[2] = {c->arms[0], nullptr};
    TRY(launch_irv(c, dv, ov, av, 1, H, W, D, zd, usd, thresh_s, thresh_h, iters, st));
    TRY(download(disp, c->disp[0], n * sizeof(float), st));
    TRY(download(outliers, c->outl[0], n, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_filter_bilateral_1(s2mv_ctx *ctx, float *img, int radius, float sigma_color, float sigma_spatial,
                                       int H, int W, int num_disp)
{
    if (!img) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (radius < 0 || radius > 24 || num_disp < 1) return fail(S2MV_ERR_BAD_PARAM, "radius/num_disp out of range");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    std::vector<float> sp, col;
    host_gaussian_kernel(sp, radius, sigma_spatial);
    host_gaussian_1d(col, num_disp, sigma_color);
    Tmp dsp, dcol;
    TRY(dsp.alloc(sp.size() * sizeof(float)));
    TRY(dcol.alloc(col.size() * sizeof(float)));
    TRY(upload(dsp.p, sp.data(), sp.size() * sizeof(float), st));
    TRY(upload(dcol.p, col.data(), col.size() * sizeof(float), st));
    TRY(upload(c->disp[0], img, n * sizeof(float), st));
    {
        // arbitrary caller data: index clamped, converted with cvt.rzi like the reference (not the bounded fast path)
        const float *bin[2] = {c->disp[0], c->disp[0]};
        float *bout[2] = {c->dispF[0], c->dispF[0]};
        TRY(launch_bilateral(c, bin, bout, 1, dsp.as<float>(), dcol.as<float>(), radius, num_disp, false, H, W, st));
    }
    TRY(download(img, c->dispF[0], n * sizeof(float), st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_dibr_occl(s2mv_ctx *ctx, uint8_t *occl_l, uint8_t *occl_r, const float *disp_l,
                              const float *disp_r, int H, int W)
{
    if (!occl_l || !occl_r || !disp_l || !disp_r) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    TRY(upload(c->disp[0], disp_l, n * sizeof(float), st));
    TRY(upload(c->disp[1], disp_r, n * sizeof(float), st));
    CU(cudaMemsetAsync(c->occl[0], 0, n, st));
    CU(cudaMemsetAsync(c->occl[1], 0, n, st));
    k_occl<<<dim3((W + 255) / 256, H), 256, 0, st>>>(c->disp[0], c->disp[1], c->occl[0], c->occl[1], H, W);
    KCHECK();
    TRY(download(occl_l, c->occl[0], n, st));
    TRY(download(occl_r, c->occl[1], n, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_filter_bleed_1(s2mv_ctx *ctx, uint8_t *img, int radius, int H, int W)
{
    if (!img) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (radius < 0 || radius > 8) return fail(S2MV_ERR_BAD_PARAM, "radius out of range");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    TRY(upload(c->occl[0], img, n, st));
    k_bleed<<<dim3((W + 255) / 256, H), 256, 0, st>>>(c->occl[0], c->occlB[0], nullptr, radius, H, W);
    KCHECK();
    TRY(download(img, c->occlB[0], n, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_dibr_occl_to_mask(s2mv_ctx *ctx, float *mask_l, float *mask_r, const uint8_t *occl_l,
                                      const uint8_t *occl_r, int H, int W)
{
    if (!mask_l || !mask_r || !occl_l || !occl_r) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    TRY(upload(c->occl[0], occl_l, n, st));
    TRY(upload(c->occl[1], occl_r, n, st));
    for (int v = 0; v < 2; ++v) {
        k_occl_to_mask<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->occl[v], c->mask[v], n);
        KCHECK();
    }
    TRY(download(mask_l, c->mask[0], n * sizeof(float), st));
    TRY(download(mask_r, c->mask[1], n * sizeof(float), st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_filter_gaussian_1(s2mv_ctx *ctx, float *img, int radius, float sigma_spatial, int H, int W)
{
    if (!img) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (radius < 0 || radius > 24) return fail(S2MV_ERR_BAD_PARAM, "radius out of range");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    std::vector<float> k;
    host_gaussian_kernel(k, radius, sigma_spatial);
    Tmp dk;
    TRY(dk.alloc(k.size() * sizeof(float)));
    TRY(upload(dk.p, k.data(), k.size() * sizeof(float), st));
    TRY(upload(c->mask[0], img, n * sizeof(float), st));
    TRY(launch_gauss(c, c->mask[0], c->tmask, dk.as<float>(), k.data(), radius, 0, H, W, st));
    TRY(download(img, c->tmask, n * sizeof(float), st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_dibr_dbm(s2mv_ctx *ctx, uint8_t *img_out, const uint8_t *img_in_l, const uint8_t *img_in_r,
                             const float *disp_l, const float *disp_r, const float *mask_l, const float *mask_r,
                             float shift, int blur_radius, float blur_sigma, int H, int W, int elem_sz)
{
    if (!img_out || !img_in_l || !img_in_r || !disp_l || !disp_r || !mask_l || !mask_r)
        return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3");
    if (blur_radius < 0 || blur_radius > 24) return fail(S2MV_ERR_BAD_PARAM, "radius out of range");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    std::vector<float> k;
    host_gaussian_kernel(k, blur_radius, blur_sigma);
    Tmp dk, imgs;
    TRY(dk.alloc(k.size() * sizeof(float)));
    TRY(imgs.alloc(2 * n * 3));
    TRY(upload(dk.p, k.data(), k.size() * sizeof(float), st));
    TRY(upload(imgs.as<uint8_t>(), img_in_l, n * 3, st));
    TRY(upload(imgs.as<uint8_t>() + n * 3, img_in_r, n * 3, st));
    TRY(upload(c->dispF[0], disp_l, n * sizeof(float), st));
    TRY(upload(c->dispF[1], disp_r, n * sizeof(float), st));
    TRY(upload(c->mask[0], mask_l, n * sizeof(float), st));
    TRY(upload(c->mask[1], mask_r, n * sizeof(float), st));
    k_unpack<<<dim3((W + 255) / 256, H), 256, 0, st>>>(imgs.as<uint8_t>(), imgs.as<uint8_t>() + n * 3, (size_t)W * 3,
                                                      c->pix[0], c->pix[1], c->gray[0], c->gray[1], nullptr, nullptr, H, W);
    KCHECK();
    TRY(launch_gauss(c, c->mask[1], c->tmask, dk.as<float>(), k.data(), blur_radius, 1, H, W, st));
    DbmArgs d;
    memset(&d, 0, sizeof(d));
    d.pixL = c->pix[0]; d.pixR = c->pix[1]; d.dispL = c->dispF[0]; d.dispR = c->dispF[1];
    d.maskL = c->mask[0]; d.maskR = c->mask[1]; d.tmask = c->tmask; d.views = c->views; d.H = H; d.W = W;
    d.shift[0] = shift; d.view_index[0] = 0;
    launch_dbm(d, 1, st);
    KCHECK();
    TRY(download(img_out, c->views, n * 3, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

// d_dibr_dfm (d_dibr_fwarp.cu:27-95): forward-warp the left image by disp_l * shift and the right one by
// disp_r * (1 - shift), then mux_merge_AB(out_l, out_r, mask) with the mask the reference builds from an
// all-zero occlusion map -- i.e. 0 everywhere, so the merged image is the left warp (reproduced as written).
// Colliding sources: the lowest source column wins (kernels_dibr.cuh; the reference races).
extern "C" int s2mv_dibr_dfm(s2mv_ctx *ctx, uint8_t *img_out, const uint8_t *img_in_l, const uint8_t *img_in_r,
                             const float *disp_l, const float *disp_r, float shift, int H, int W, int elem_sz)
{
    if (!img_out || !img_in_l || !img_in_r || !disp_l || !disp_r) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3");
    s2mv_ctx *c;
    TRY(acquire(ctx, H, W, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)H * W;
    Tmp imgs, outs;
    TRY(imgs.alloc(2 * n * 3));
    TRY(outs.alloc(2 * n * 3));
    uint8_t *in_l = imgs.as<uint8_t>(), *in_r = in_l + n * 3, *out_l = outs.as<uint8_t>(), *out_r = out_l + n * 3;
    TRY(upload(in_l, img_in_l, n * 3, st));
    TRY(upload(in_r, img_in_r, n * 3, st));
    TRY(upload(c->dispF[0], disp_l, n * sizeof(float), st));
    TRY(upload(c->dispF[1], disp_r, n * sizeof(float), st));
    const dim3 g((W + 255) / 256, H);
    const float shifts[2] = {shift, (float)(1.0 - (double)shift)};  // d_dibr_fwarp.cu:72: 1.0 - shift in double
    const uint8_t *ins[2] = {in_l, in_r};
    uint8_t *os[2] = {out_l, out_r};
    for (int v = 0; v < 2; ++v) {
        int *winner = c->irv_vote[v];  // an n-int scratch plane of the arena
        CU(cudaMemsetAsync(winner, 0x7f, n * sizeof(int), st));  // 0x7f7f7f7f: above every column
        k_fwarp_claim<<<g, 256, 0, st>>>(c->dispF[v], shifts[v], winner, H, W);
        KCHECK();
        k_fwarp_gather<<<g, 256, 0, st>>>(ins[v], winner, os[v], H, W);
        KCHECK();
    }
    // dibr_occl_to_mask of a zeroed map (d_dibr_fwarp.cu:55-66): the mask is 0 everywhere
    k_merge_ab<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out_l, out_r, nullptr, n);
    KCHECK();
    TRY(download(img_out, out_l, n * 3, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}

extern "C" int s2mv_mux_multiview(s2mv_ctx *ctx, uint8_t **views, uint8_t *out, int V, float angle, int Hin, int Win,
                                  int Hout, int Wout, int elem_sz, int kernel_variant)
{
    if (!views || !out) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (elem_sz != 3) return fail(S2MV_ERR_BAD_PARAM, "elem_sz must be 3");
    if (V < 1 || V > 16) return fail(S2MV_ERR_BAD_PARAM, "num_views must be in [1,16]");
    s2mv_ctx *c;
    TRY(acquire(ctx, Hin, Win, ctx && ctx->configured ? ctx->prm.num_disp : 64,
                ctx && ctx->configured ? ctx->prm.zero_disp : 32, ctx && ctx->configured ? ctx->prm.usd : 17, &c));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)Hin * Win, no = (size_t)Hout * Wout;
    Tmp dv, dout;
    TRY(dv.alloc((size_t)V * n * 3));
    TRY(dout.alloc(no * 3));
    const uint8_t *vp[16];
    for (int v = 0; v < V; ++v) {
        TRY(upload(dv.as<uint8_t>() + (size_t)v * n * 3, views[v], n * 3, st));
        vp[v] = dv.as<uint8_t>() + (size_t)v * n * 3;
    }
    TRY(launch_mux(c, vp, dout.as<uint8_t>(), V, angle, Hin, Win, Hout, Wout, elem_sz, kernel_variant, st));
    TRY(download(out, dout.p, no * 3, st));
    CU(cudaStreamSynchronize(st));
    return S2MV_OK;
}
