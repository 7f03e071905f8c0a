// s2mv_band.inl — one frame split into contiguous row bands over several contexts (one per GPU).
//
// A band context is an ordinary context over a SUB-IMAGE of the frame: its own rows plus an apron
// of `apron` rows either side (clipped at the frame's top and bottom).  Every stage of the pipeline
// is a local operator with a finite vertical reach, so running it on the sub-image as if it were a
// whole image gives the frame's exact values on every row further than that reach from an interior
// sub-image edge:
//     census 3, arms usd, region voting usd per iteration (5 iterations), bilateral 7,
//     bleed 1, mask blur 10  ->  5*usd + 18 rows of WTA disparities, 6*usd + 18 rows of pixels
// (103 / 120 rows at usd = 17; the default apron is 128).  The cheap O(W*H) stages run that way.
// The cost volume does not: its passes run on the band's OWN rows only, and the usd rows either
// side that the two vertical passes read are received from the neighbouring bands (the caller
// moves them: peer copies / NCCL send-recv over NVLink; rows of the disparity-innermost volume are
// contiguous, so a halo is one flat run per view).  The sequence per frame is
//     prepare -> pass 1 (CI + H) -> halo A -> pass 2 (V) -> halo B -> pass 3 (V) -> pass 4 (H + WTA)
//     -> gather the WTA disparities of the sub-image rows -> finish (refine + DIBR + interlace).
static int band_check(const s2mv_ctx *c)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (!c->configured || !c->band) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_configure_band first");
    return S2MV_OK;
}

// A wait that ran out leaves the halo rows of that frame stale: every later call on the band fails until it
// is reconfigured (the word is in mapped host memory: no synchronisation needed to look at it).
static int band_poisoned(const s2mv_ctx *c)
{
    if (c->band_status_h && *(volatile unsigned int *)c->band_status_h)
        return fail(S2MV_ERR_CUDA, "a neighbouring band never reached the pass this band waited for: the frame is invalid");
    return S2MV_OK;
}

static int configure_band_impl(s2mv_ctx *c, const s2mv_params *frame, int band_y0, int band_y1, int apron, int min_band_rows,
                               int fuse)
{
    if (!c || !frame) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    const int H = frame->num_rows, usd = frame->usd;
    if (band_y0 < 0 || band_y1 > H || band_y0 >= band_y1) return fail(S2MV_ERR_BAD_PARAM, "band [%d,%d) outside the %d-row frame", band_y0, band_y1, H);
    if (frame->num_rows_out != frame->num_rows || frame->num_cols_out != frame->num_cols)
        return fail(S2MV_ERR_BAD_PARAM, "row-band mode needs output size == input size");
    if ((band_y0 > 0 || band_y1 < H) && band_y1 - band_y0 < usd)
        return fail(S2MV_ERR_BAD_PARAM, "a band must have at least usd (%d) rows", usd);
    const int reach = 6 * usd + 18;  // see the header comment
    if (apron <= 0) apron = ((reach + 31) / 32) * 32;
    if (apron < reach) return fail(S2MV_ERR_BAD_PARAM, "apron %d < vertical reach %d", apron, reach);
    BandSpec b;
    b.frame_rows = H;
    b.ly0 = band_y0 - apron > 0 ? band_y0 - apron : 0;
    const int ly1 = band_y1 + apron < H ? band_y1 + apron : H;
    b.o0 = band_y0 - b.ly0;
    b.o1 = band_y1 - b.ly0;
    b.sub_rows = ly1 - b.ly0;
    b.min_band_rows = min_band_rows;
    b.fuse = fuse;
    s2mv_params p = *frame;
    p.num_rows = p.num_rows_out = ly1 - b.ly0;
    return configure_impl(c, &p, &b);   // the volume rows [vlo, vhi) follow from the plan (fused vertical passes or not)
}

extern "C" int s2mv_configure_band(s2mv_ctx *c, const s2mv_params *frame, int band_y0, int band_y1, int apron)
{
    return configure_band_impl(c, frame, band_y0, band_y1, apron, 0, 0);
}

// min_band_rows: the smallest band of the frame; fuse_vertical: 0 = two vertical passes, 1 = one launch whenever every band
// has 2*usd rows, -1 = one launch where that is faster (tall bands).  The same values on every band of the frame.  Fused,
// the bands exchange halo rows once, not twice (s2mv_band_info reports the rows per exchange; pass 3 is then part of
// pass 2 and returns at once).
extern "C" int s2mv_configure_band_ex(s2mv_ctx *c, const s2mv_params *frame, int band_y0, int band_y1, int apron,
                                      int min_band_rows, int fuse_vertical)
{
    if (frame && min_band_rows > 0 && band_y1 - band_y0 < min_band_rows)
        return fail(S2MV_ERR_BAD_PARAM, "band of %d rows is smaller than min_band_rows %d", band_y1 - band_y0, min_band_rows);
    return configure_band_impl(c, frame, band_y0, band_y1, apron, min_band_rows, fuse_vertical);
}

extern "C" int s2mv_band_info(const s2mv_ctx *c, int *local_y0, int *local_rows, int *own_first, int *own_rows,
                              int *halo_rows)
{
    TRY(band_check(c));
    if (local_y0) *local_y0 = c->band_ly0;
    if (local_rows) *local_rows = c->prm.num_rows;
    if (own_first) *own_first = c->band_o0;
    if (own_rows) *own_rows = c->band_o1 - c->band_o0;
    if (halo_rows) *halo_rows = c->band_halo1;
    return S2MV_OK;
}

static void band_geometry(const s2mv_ctx *c, RowRange &rr, size_t &row4, size_t &view_stride4)
{
    rr.own0 = c->band_o0; rr.own1 = c->band_o1; rr.vlo = c->band_vlo; rr.vhi = c->band_vhi;
    row4 = (size_t)c->prm.num_cols * c->plan.LPtot;                 // float4 per volume row
    view_stride4 = (size_t)(c->band_vhi - c->band_vlo) * row4;
}

// demux + gray + census + arms of the sub-image; `d_img_sbs_frame` is the WHOLE frame on this device
extern "C" int s2mv_band_prepare(s2mv_ctx *c, const uint8_t *d_img_sbs_frame, int num_cols_sbs, void *stream)
{
    TRY(band_check(c));
    TRY(band_poisoned(c));
    if (!d_img_sbs_frame) return fail(S2MV_ERR_BAD_PARAM, "null frame");
    const s2mv_params &p = c->prm;
    const int H = p.num_rows, W = p.num_cols, V = p.num_views;
    if (num_cols_sbs < 2 * W) return fail(S2MV_ERR_BAD_PARAM, "num_cols_sbs (%d) < 2*num_cols (%d)", num_cols_sbs, 2 * W);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t n = (size_t)H * W, pitch = (size_t)num_cols_sbs * 3;
    const uint8_t *sub = d_img_sbs_frame + (size_t)c->band_ly0 * pitch;
    c->launches = 0;
    uint8_t *view0 = c->views, *viewN = c->views + (size_t)(V - 1) * n * 3;
    TRY(launch_prepare(c, sub, sub + (size_t)W * 3, pitch, viewN, view0, st));
    TRY(build_luts(c, p.ad_coeff, p.census_coeff, st));
    if (c->plan.nchunks > 1)
        for (int v = 0; v < 2; ++v) CU(cudaMemsetAsync(c->wta_key[v], 0xff, n * sizeof(unsigned long long), st));
    return S2MV_OK;
}

// pass 1: cost initialisation + horizontal pass -> A (own rows); 2: vertical A -> B; 3: vertical B -> A;
// 4: horizontal + winner-takes-all -> the sub-image's disparity planes (own rows)
extern "C" int s2mv_band_pass(s2mv_ctx *c, int pass, void *stream)
{
    TRY(band_check(c));
    TRY(band_poisoned(c));
    if (pass < 1 || pass > 4) return fail(S2MV_ERR_BAD_PARAM, "pass must be 1..4");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const s2mv_params &p = c->prm;
    const size_t n = (size_t)p.num_rows * p.num_cols;
    RowRange rr;
    size_t row4, view_stride4;
    band_geometry(c, rr, row4, view_stride4);
    LineArgs a;
    fill_largs(c, a, p.num_rows, p.num_cols, p.zero_disp, p.ad_coeff);
    for (int v = 0; v < 2; ++v) { a.arms[v] = c->arms[v]; a.wta_key[v] = c->wta_key[v]; a.disp[v] = c->disp[v]; }
    const bool p2p = c->band_peer[0].connected || c->band_peer[1].connected;
    const bool fused = c->band_fused;
    if (fused && pass == 3) return S2MV_OK;                        // done by pass 2
    const bool waits = pass == 2 || (pass == 3 && !fused), pushes = pass == 1 || (pass == 2 && !fused);
    a.peer_lo_end = 0;
    a.peer_hi_begin = 0x7fffffff;
    if (p2p && waits) {
        // the halo rows this pass reads are written by the neighbours' previous pass: wait for their epoch
        for (int side = 0; side < 2; ++side)
            if (c->band_peer[side].connected) {
                k_band_wait<<<1, 1, 0, st>>>(c->band_flags + side, c->band_epoch, c->band_status_d, c->env_band_wait_spins);
                KCHECK();
            }
    }
    if (p2p && pushes) {
        // this pass's rows next to a band edge also go straight into the neighbour's halo rows
        const int usd = pass == 1 ? c->band_halo1 : p.usd, buf = pass == 1 ? 0 : 1;
        for (int side = 0; side < 2; ++side) {
            const s2mv_ctx::BandPeer &pr = c->band_peer[side];
            if (!pr.connected) continue;
            for (int v = 0; v < 2; ++v)
                a.peer_out[side][v] = reinterpret_cast<float4 *>(pr.vol[buf]) + (size_t)v * pr.view_stride4 +
                                      pr.row_bias * (long long)row4;
            if (side == 0) a.peer_lo_end = rr.own0 + usd; else a.peer_hi_begin = rr.own1 - usd;
        }
    }
    float4 *A = reinterpret_cast<float4 *>(c->vol[0]), *B = reinterpret_cast<float4 *>(c->vol[1]);
    if (fused && pass == 2) {
        TRY(launch_vv(c, a, B, view_stride4, 2, rr, st));          // A (own rows + 2*usd halo rows) -> B (own rows)
    } else if (fused && pass == 4) {
        TRY(launch_pass(c, a, 4, /*read*/ B, /*unused*/ A, view_stride4, 2, true, true, rr, st));
    } else {
        TRY(launch_pass(c, a, pass, A, B, view_stride4, 2, true, true, rr, st));
    }
    if (p2p && pushes) {
        c->band_epoch += 1;
        for (int side = 0; side < 2; ++side)
            if (c->band_peer[side].connected) {
                k_band_signal<<<1, 1, 0, st>>>(c->band_peer[side].flag, c->band_epoch);
                KCHECK();
            }
    }
    if (pass == 4 && c->plan.nchunks > 1) {
        k_wta_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->wta_key[0], c->disp[0], p.zero_disp, n);
        k_wta_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->wta_key[1], c->disp[1], p.zero_disp, n);
        KCHECK();
        c->launches += 2;
    }
    return S2MV_OK;
}

// Halo runs of the volume written by `after_pass` (1 -> A, 2 -> B) for one view.
//   side 0 = towards row 0, 1 = towards the last row;  recv 0: the usd own rows next to that side (what the
//   neighbour on that side needs), recv 1: the usd halo rows beyond that side (where its rows go).
// *bytes is 0 where the band touches the frame's edge.
extern "C" int s2mv_band_halo(s2mv_ctx *c, int after_pass, int view, int side, int recv, void **ptr, size_t *bytes)
{
    TRY(band_check(c));
    if ((after_pass != 1 && after_pass != 2) || view < 0 || view > 1 || side < 0 || side > 1 || !ptr || !bytes)
        return fail(S2MV_ERR_BAD_PARAM, "bad halo selector");
    RowRange rr;
    size_t row4, view_stride4;
    band_geometry(c, rr, row4, view_stride4);
    const int usd = after_pass == 1 ? c->band_halo1 : (c->band_fused ? 0 : c->prm.usd);
    const bool at_edge = side == 0 ? (c->band_ly0 + rr.own0 == 0) : (c->band_ly0 + rr.own1 == c->band_frame_rows);
    int r0, r1;
    if (side == 0) { r0 = recv ? rr.own0 - usd : rr.own0; r1 = r0 + usd; }
    else           { r0 = recv ? rr.own1 : rr.own1 - usd; r1 = r0 + usd; }
    if (at_edge || usd == 0) { *ptr = nullptr; *bytes = 0; return S2MV_OK; }
    if (r0 < rr.vlo || r1 > rr.vhi) return fail(S2MV_ERR_BAD_PARAM, "halo rows [%d,%d) outside the band's volume rows [%d,%d)", r0, r1, rr.vlo, rr.vhi);
    float4 *base = reinterpret_cast<float4 *>(c->vol[after_pass == 1 ? 0 : 1]);
    *ptr = base + (size_t)view * view_stride4 + (size_t)(r0 - rr.vlo) * row4;
    *bytes = (size_t)usd * row4 * sizeof(float4);
    return S2MV_OK;
}

// the sub-image's WTA disparity plane of one view (num_rows x num_cols floats, local rows); pass 4 fills
// the own rows, the caller fills the apron rows from the other bands before s2mv_band_finish
extern "C" int s2mv_band_disp(s2mv_ctx *c, int view, float **d_plane)
{
    TRY(band_check(c));
    if (view < 0 || view > 1 || !d_plane) return fail(S2MV_ERR_BAD_PARAM, "bad argument");
    *d_plane = c->disp[view];
    return S2MV_OK;
}

// refinement + DIBR + interlace on the sub-image, then the band's OWN rows of the three outputs
// (own_rows x num_cols floats / x3 bytes each; any may be NULL)
extern "C" int s2mv_band_finish(s2mv_ctx *c, float *d_disp_l_band, float *d_disp_r_band, uint8_t *d_interlaced_band,
                                void *stream)
{
    TRY(band_check(c));
    TRY(band_poisoned(c));
    CU(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const s2mv_params &p = c->prm;
    const size_t W = p.num_cols, own = (size_t)(c->band_o1 - c->band_o0), off = (size_t)c->band_o0 * W;
    TRY(run_refine_dibr(c, c->dispF[0], c->dispF[1], c->interlaced, st));
    if (d_disp_l_band) CU(cudaMemcpyAsync(d_disp_l_band, c->dispF[0] + off, own * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (d_disp_r_band) CU(cudaMemcpyAsync(d_disp_r_band, c->dispF[1] + off, own * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (d_interlaced_band) CU(cudaMemcpyAsync(d_interlaced_band, c->interlaced + off * 3, own * W * 3, cudaMemcpyDeviceToDevice, st));
    return S2MV_OK;
}

// ---- peer-to-peer halos -------------------------------------------------------------------------
// With the neighbours connected, passes 1 and 2 store the rows next to a band edge straight into the
// neighbour's halo rows (peer-mapped memory: NVLink stores issued by the producing kernel, tile by tile)
// and the caller exchanges nothing: s2mv_band_pass orders the bands with an epoch word per neighbour
// (k_band_signal / k_band_wait on the stream).  Every band of a frame must then run the same pass
// sequence, and all of a frame's passes 1 (and 2) must be enqueued before any band waits when several
// bands share one stream.
static int band_attach(s2mv_ctx *c, int side, float *volA, float *volB, unsigned int *peer_flags, int peer_ly0,
                       int peer_vlo, int peer_vrows)
{
    s2mv_ctx::BandPeer &pr = c->band_peer[side];
    pr.vol[0] = volA;
    pr.vol[1] = volB;
    pr.flag = peer_flags + (1 - side);  // I am the neighbour's lower (side 0: I attach my upper) / upper neighbour
    pr.row_bias = (long long)c->band_ly0 - peer_ly0 - peer_vlo;
    pr.view_stride4 = (size_t)peer_vrows * c->prm.num_cols * c->plan.LPtot;
    pr.connected = true;
    return S2MV_OK;
}

// same process: `peer` is the band context above (side 0) or below (side 1) this one
extern "C" int s2mv_band_connect(s2mv_ctx *c, int side, s2mv_ctx *peer)
{
    TRY(band_check(c));
    TRY(band_check(peer));
    if (side < 0 || side > 1) return fail(S2MV_ERR_BAD_PARAM, "side must be 0 (upper neighbour) or 1 (lower)");
    if (peer->prm.num_cols != c->prm.num_cols || peer->plan.LPtot != c->plan.LPtot || peer->band_frame_rows != c->band_frame_rows)
        return fail(S2MV_ERR_BAD_PARAM, "neighbour bands must belong to the same frame");
    const int my_edge = c->band_ly0 + (side == 0 ? c->band_o0 : c->band_o1);
    const int peer_edge = peer->band_ly0 + (side == 0 ? peer->band_o1 : peer->band_o0);
    if (my_edge != peer_edge) return fail(S2MV_ERR_BAD_PARAM, "bands are not adjacent (rows %d vs %d)", my_edge, peer_edge);
    if (peer->band_halo1 != c->band_halo1 || peer->band_fused != c->band_fused)
        return fail(S2MV_ERR_BAD_PARAM, "neighbour bands were configured with different halo exchanges (min_band_rows must agree)");
    if (peer->device != c->device) {
        CU(cudaSetDevice(c->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(S2MV_ERR_CUDA, "no peer access %d -> %d: %s", c->device, peer->device, cudaGetErrorString(e));
        cudaGetLastError();
    }
    if (c->band_peer_ctx[side] != peer) {
        if (s2mv_ctx *old = c->band_peer_ctx[side]) {
            auto &v = old->band_attached_by;
            v.erase(std::remove(v.begin(), v.end(), c), v.end());
        }
        c->band_peer_ctx[side] = peer;
        peer->band_attached_by.push_back(c);   // peer's free_arena disconnects this band first
    }
    return band_attach(c, side, peer->vol[0], peer->vol[1], peer->band_flags, peer->band_ly0, peer->band_vlo,
                       peer->band_vhi - peer->band_vlo);
}

// other process (one process per GPU): what a neighbour needs to map this band's volumes
extern "C" int s2mv_band_ipc_export(s2mv_ctx *c, s2mv_band_ipc *out)
{
    TRY(band_check(c));
    if (!out) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    CU(cudaSetDevice(c->device));
    memset(out, 0, sizeof(*out));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    void *ptrs[3] = {c->vol[0], c->vol[1], c->band_flags};
    for (int i = 0; i < 3; ++i) {
        CU(cudaIpcGetMemHandle(&h, ptrs[i]));
        memcpy(out->mem[i], &h, 64);
    }
    out->frame_y0 = c->band_ly0 + c->band_o0;
    out->frame_y1 = c->band_ly0 + c->band_o1;
    out->local_y0 = c->band_ly0;
    out->vlo = c->band_vlo;
    out->vrows = c->band_vhi - c->band_vlo;
    out->num_cols = c->prm.num_cols;
    out->lptot = c->plan.LPtot;
    out->frame_rows = c->band_frame_rows;
    out->device = c->device;
    out->halo_rows = c->band_halo1;
    out->fused = c->band_fused ? 1 : 0;
    return S2MV_OK;
}

extern "C" int s2mv_band_ipc_connect(s2mv_ctx *c, int side, const s2mv_band_ipc *peer)
{
    TRY(band_check(c));
    if (!peer || side < 0 || side > 1) return fail(S2MV_ERR_BAD_PARAM, "bad argument");
    if (peer->num_cols != c->prm.num_cols || peer->lptot != c->plan.LPtot || peer->frame_rows != c->band_frame_rows)
        return fail(S2MV_ERR_BAD_PARAM, "neighbour bands must belong to the same frame");
    const int my_edge = c->band_ly0 + (side == 0 ? c->band_o0 : c->band_o1);
    const int peer_edge = side == 0 ? peer->frame_y1 : peer->frame_y0;
    if (my_edge != peer_edge) return fail(S2MV_ERR_BAD_PARAM, "bands are not adjacent (rows %d vs %d)", my_edge, peer_edge);
    if (peer->halo_rows != c->band_halo1 || (peer->fused != 0) != c->band_fused)
        return fail(S2MV_ERR_BAD_PARAM, "neighbour bands were configured with different halo exchanges (min_band_rows must agree)");
    CU(cudaSetDevice(c->device));
    s2mv_ctx::BandPeer &pr = c->band_peer[side];
    void *m[3] = {};
    for (int i = 0; i < 3; ++i) {
        cudaIpcMemHandle_t h;
        memcpy(&h, peer->mem[i], 64);
        CU(cudaIpcOpenMemHandle(&m[i], h, cudaIpcMemLazyEnablePeerAccess));
        pr.ipc[i] = m[i];
    }
    return band_attach(c, side, (float *)m[0], (float *)m[1], (unsigned int *)m[2], peer->local_y0, peer->vlo, peer->vrows);
}

// 0 when every wait of this band met its neighbour; synchronises the stream first, so the answer covers
// everything enqueued so far.  (s2mv_band_prepare / _pass / _finish look at the same word without
// synchronising and fail once a wait has run out.)
extern "C" int s2mv_band_status(s2mv_ctx *c, void *stream)
{
    TRY(band_check(c));
    CU(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CU(cudaStreamSynchronize(st));
    return band_poisoned(c);
}
