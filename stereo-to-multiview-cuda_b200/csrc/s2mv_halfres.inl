// s2mv_halfres.inl — adcensus_stm_2 (d_io.cu:240-508): disparities estimated on bilinearly
// down-scaled images, scaled back up, DIBR + interlace at full resolution.  Two contexts: `c`
// at full resolution (no cost volumes), c->lo at the disparity resolution (the whole estimation
// pipeline); one stream.
extern "C" int s2mv_configure_2(s2mv_ctx *c, const s2mv_params *p, int num_rows_disp, int num_cols_disp, float disp_scale)
{
    if (!c || !p) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (num_rows_disp < 1 || num_cols_disp < 1) return fail(S2MV_ERR_BAD_PARAM, "disparity-resolution size must be positive");
    if (!(disp_scale > 0.f)) return fail(S2MV_ERR_BAD_PARAM, "disp_scale must be positive");
    if (c->lo) { s2mv_destroy(c->lo); c->lo = nullptr; }
    c->no_volume = true;
    int st = configure_impl(c, p, nullptr);
    if (st != S2MV_OK) { c->no_volume = false; return st; }
    c->disp_scale = disp_scale;
    TRY(s2mv_create(&c->lo, c->device));
    s2mv_params q = *p;
    q.num_rows = q.num_rows_out = num_rows_disp;
    q.num_cols = q.num_cols_out = num_cols_disp;
    c->lo->chunk_seq_mode = c->chunk_seq_mode;
    st = configure_impl(c->lo, &q, nullptr);
    if (st == S2MV_OK) {
        cudaSetDevice(c->device);
        for (int v = 0; v < 2 && st == S2MV_OK; ++v)
            st = dev_alloc_t(c, &c->lo_bgr[v], (size_t)num_rows_disp * num_cols_disp * 3);
    }
    if (st != S2MV_OK) {  // leave no half-built two-resolution context behind
        s2mv_destroy(c->lo);
        c->lo = nullptr;
        c->configured = false;
    }
    return st;
}

static int run_frame_2(s2mv_ctx *c, const uint8_t *d_sbs, int num_cols_sbs, float *d_disp_l, float *d_disp_r,
                       uint8_t *d_interlaced, cudaStream_t st)
{
    s2mv_ctx *lo = c->lo;
    const s2mv_params &p = c->prm, &q = lo->prm;
    const int H = p.num_rows, W = p.num_cols, V = p.num_views, Hd = q.num_rows, Wd = q.num_cols;
    const size_t n = (size_t)H * W;
    if (num_cols_sbs < 2 * W) return fail(S2MV_ERR_BAD_PARAM, "num_cols_sbs (%d) < 2*num_cols (%d)", num_cols_sbs, 2 * W);
    c->launches = 0;
    lo->launches = 0;
    if (c->timing) CU(cudaEventRecord(c->ev[0], st));
    // demux_sbs (d_io.cu:291): full-resolution packed pixels + the two outer views
    uint8_t *view0 = c->views, *viewN = c->views + (size_t)(V - 1) * n * 3;
    dim3 g((W + 255) / 256, H);
    k_unpack<<<g, 256, 0, st>>>(d_sbs, d_sbs + (size_t)W * 3, (size_t)num_cols_sbs * 3, c->pix[0], c->pix[1], c->gray[0],
                                c->gray[1], viewN, view0, H, W);
    KCHECK();
    // tx_scale_bilinear_kernel x2 (d_io.cu:301-303)
    dim3 gd((Wd + 255) / 256, Hd);
    k_scale_bilinear_bgr<<<gd, 256, 0, st>>>(viewN, c->lo_bgr[0], H, W, Hd, Wd);
    k_scale_bilinear_bgr<<<gd, 256, 0, st>>>(view0, c->lo_bgr[1], H, W, Hd, Wd);
    KCHECK();
    c->launches += 3;
    // estimation at the disparity resolution (d_io.cu:309-420)
    TRY(launch_prepare(lo, c->lo_bgr[0], c->lo_bgr[1], (size_t)Wd * 3, nullptr, nullptr, st));
    TRY(build_luts(lo, q.ad_coeff, q.census_coeff, st));
    if (c->timing) CU(cudaEventRecord(c->ev[1], st));
    TRY(launch_costvol(lo, lo->disp[0], lo->disp[1], st));
    if (c->timing) CU(cudaEventRecord(c->ev[2], st));
    TRY(run_refine(lo, lo->dispF[0], lo->dispF[1], st));
    // tx_disp_scale_kernel x2 (d_io.cu:425-426): 1.0f / disp_scale
    float *fl = d_disp_l ? d_disp_l : c->dispF[0], *fr = d_disp_r ? d_disp_r : c->dispF[1];
    const float inv = 1.0f / c->disp_scale;
    k_disp_scale<<<g, 256, 0, st>>>(fl, lo->dispF[0], H, W, Hd, Wd, inv);
    k_disp_scale<<<g, 256, 0, st>>>(fr, lo->dispF[1], H, W, Hd, Wd, inv);
    KCHECK();
    c->launches += 2;
    CU(cudaEventRecord(c->ev_refined, st));
    if (c->timing) CU(cudaEventRecord(c->ev[3], st));
    TRY(run_dibr(c, fl, fr, d_interlaced, st));
    c->launches += lo->launches;
    return S2MV_OK;
}

static int check_2(const s2mv_ctx *c)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (!c->configured || !c->lo) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_configure_2 first");
    return S2MV_OK;
}

extern "C" int s2mv_process_sbs_2_device(s2mv_ctx *c, const uint8_t *d_img_sbs, int num_cols_sbs, float *d_disp_l,
                                         float *d_disp_r, uint8_t *d_interlaced, void *stream)
{
    TRY(check_2(c));
    if (!d_img_sbs) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    CU(cudaSetDevice(c->device));
    return run_frame_2(c, d_img_sbs, num_cols_sbs, d_disp_l, d_disp_r, d_interlaced, stream ? (cudaStream_t)stream : c->stream);
}
