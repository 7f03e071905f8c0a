// s2mv_stream.inl — asynchronous frame stream: the video loop of the reference
// (video_io.cpp:139-160: decode -> adcensus_stm -> display, one synchronous call per frame,
// pageable copies inside it, d_io.cu:43-44,153-154,205) with the copies taken off the critical
// path.  `depth` slots each own pinned host buffers and device in/out buffers; three streams:
//   st_in   H2D of frame i+1          |  overlaps
//   stream  the frame's kernels       |  compute of frame i
//   st_out  D2H of frame i-1          |
// The arena (volumes, planes) is shared: frames are serialised on the compute stream, which is
// the bottleneck resource anyway.  Results are identical to s2mv_process_sbs (same kernels, same
// order); only the waiting moves.
static void stream_release(s2mv_ctx *c)
{
    if (c->slots.empty() && !c->st_in) return;
    cudaSetDevice(c->device);
    if (c->st_in) cudaStreamSynchronize(c->st_in);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->st_out) cudaStreamSynchronize(c->st_out);
    for (auto &s : c->slots) {
        if (s.d_sbs) cudaFree(s.d_sbs);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_sbs) cudaFreeHost(s.h_sbs);
        if (s.h_out) cudaFreeHost(s.h_out);
        for (int v = 0; v < 2; ++v) {
            if (s.d_disp[v]) cudaFree(s.d_disp[v]);
            if (s.h_disp[v]) cudaFreeHost(s.h_disp[v]);
        }
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    c->slots.clear();
    if (c->st_in) cudaStreamDestroy(c->st_in);
    if (c->st_out) cudaStreamDestroy(c->st_out);
    c->st_in = c->st_out = nullptr;
    c->slot_head = c->slot_tail = c->slots_pending = 0;
    c->stream_cols_sbs = 0;
}

extern "C" int s2mv_stream_open(s2mv_ctx *c, int depth, int num_cols_sbs)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (!c->configured) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_configure first");
    if (depth < 1 || depth > 16) return fail(S2MV_ERR_BAD_PARAM, "depth must be in [1,16]");
    const s2mv_params &p = c->prm;
    if (num_cols_sbs < 2 * p.num_cols) return fail(S2MV_ERR_BAD_PARAM, "num_cols_sbs (%d) < 2*num_cols (%d)", num_cols_sbs, 2 * p.num_cols);
    CU(cudaSetDevice(c->device));
    stream_release(c);
    CU(cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    const size_t n = (size_t)p.num_rows * p.num_cols;
    const size_t sbs_bytes = (size_t)p.num_rows * num_cols_sbs * 3;
    const size_t out_bytes = (size_t)p.num_rows_out * p.num_cols_out * 3;
    c->slots.resize(depth);
    for (auto &s : c->slots) {
        CU(cudaMalloc((void **)&s.d_sbs, sbs_bytes));
        CU(cudaMalloc((void **)&s.d_out, out_bytes));
        CU(cudaMallocHost((void **)&s.h_sbs, sbs_bytes));
        CU(cudaMallocHost((void **)&s.h_out, out_bytes));
        for (int v = 0; v < 2; ++v) {
            CU(cudaMalloc((void **)&s.d_disp[v], n * sizeof(float)));
            CU(cudaMallocHost((void **)&s.h_disp[v], n * sizeof(float)));
        }
        CU(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    c->stream_cols_sbs = num_cols_sbs;
    return S2MV_OK;
}

extern "C" int s2mv_stream_close(s2mv_ctx *c)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    stream_release(c);
    return S2MV_OK;
}

extern "C" int s2mv_stream_pending(const s2mv_ctx *c) { return c ? c->slots_pending : 0; }

extern "C" int s2mv_stream_input_buffer(s2mv_ctx *c, uint8_t **pinned)
{
    if (!c || !pinned) return fail(S2MV_ERR_BAD_PARAM, "null argument");
    if (c->slots.empty()) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_stream_open first");
    if (c->slots_pending == (int)c->slots.size()) return fail(S2MV_ERR_BAD_PARAM, "all %d slots in flight: collect a frame first", (int)c->slots.size());
    *pinned = c->slots[c->slot_head].h_sbs;
    return S2MV_OK;
}

extern "C" int s2mv_stream_submit(s2mv_ctx *c, const uint8_t *img_sbs)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (c->slots.empty()) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_stream_open first");
    if (c->slots_pending == (int)c->slots.size())
        return fail(S2MV_ERR_BAD_PARAM, "all %d slots in flight: collect a frame first", (int)c->slots.size());
    CU(cudaSetDevice(c->device));
    const s2mv_params &p = c->prm;
    const size_t n = (size_t)p.num_rows * p.num_cols;
    const size_t sbs_bytes = (size_t)p.num_rows * c->stream_cols_sbs * 3;
    const size_t out_bytes = (size_t)p.num_rows_out * p.num_cols_out * 3;
    s2mv_ctx::StreamSlot &s = c->slots[c->slot_head];
    // NULL / the slot's own buffer: already in place.  A page-locked frame (cudaHostAlloc / cudaHostRegister'ed by the
    // caller) is read by the copy engine where it lies -- it must stay unchanged until the frame is collected; any
    // other frame is staged through the slot's pinned buffer.
    const uint8_t *src = s.h_sbs;
    if (img_sbs && img_sbs != s.h_sbs) {
        cudaPointerAttributes at;
        const bool locked = cudaPointerGetAttributes(&at, img_sbs) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (locked) src = img_sbs;
        else memcpy(s.h_sbs, img_sbs, sbs_bytes);
    }
    CU(cudaMemcpyAsync(s.d_sbs, src, sbs_bytes, cudaMemcpyHostToDevice, c->st_in));
    CU(cudaEventRecord(s.ev_in, c->st_in));
    CU(cudaStreamWaitEvent(c->stream, s.ev_in, 0));
    TRY(run_frame(c, s.d_sbs, c->stream_cols_sbs, s.d_disp[0], s.d_disp[1], s.d_out, false, c->stream));
    CU(cudaEventRecord(s.ev_done, c->stream));
    CU(cudaStreamWaitEvent(c->st_out, s.ev_done, 0));
    CU(cudaMemcpyAsync(s.h_disp[0], s.d_disp[0], n * sizeof(float), cudaMemcpyDeviceToHost, c->st_out));
    CU(cudaMemcpyAsync(s.h_disp[1], s.d_disp[1], n * sizeof(float), cudaMemcpyDeviceToHost, c->st_out));
    CU(cudaMemcpyAsync(s.h_out, s.d_out, out_bytes, cudaMemcpyDeviceToHost, c->st_out));
    CU(cudaEventRecord(s.ev_out, c->st_out));
    s.busy = true;
    c->slot_head = (c->slot_head + 1) % (int)c->slots.size();
    c->slots_pending += 1;
    return S2MV_OK;
}

extern "C" int s2mv_stream_collect(s2mv_ctx *c, float *disp_l, float *disp_r, uint8_t *interlaced,
                                   const float **pinned_disp_l, const float **pinned_disp_r,
                                   const uint8_t **pinned_interlaced)
{
    if (!c) return fail(S2MV_ERR_BAD_PARAM, "null ctx");
    if (c->slots.empty()) return fail(S2MV_ERR_NOT_CONFIGURED, "call s2mv_stream_open first");
    if (c->slots_pending == 0) return fail(S2MV_ERR_BAD_PARAM, "no frame in flight");
    CU(cudaSetDevice(c->device));
    const s2mv_params &p = c->prm;
    const size_t n = (size_t)p.num_rows * p.num_cols;
    const size_t out_bytes = (size_t)p.num_rows_out * p.num_cols_out * 3;
    s2mv_ctx::StreamSlot &s = c->slots[c->slot_tail];
    CU(cudaEventSynchronize(s.ev_out));
    if (disp_l) memcpy(disp_l, s.h_disp[0], n * sizeof(float));
    if (disp_r) memcpy(disp_r, s.h_disp[1], n * sizeof(float));
    if (interlaced) memcpy(interlaced, s.h_out, out_bytes);
    if (pinned_disp_l) *pinned_disp_l = s.h_disp[0];
    if (pinned_disp_r) *pinned_disp_r = s.h_disp[1];
    if (pinned_interlaced) *pinned_interlaced = s.h_out;
    s.busy = false;
    c->slot_tail = (c->slot_tail + 1) % (int)c->slots.size();
    c->slots_pending -= 1;
    return S2MV_OK;
}
