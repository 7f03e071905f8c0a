"""One frame over several GPUs in contiguous row bands (mode 2; BASELINE config 4).

The reference has no multi-GPU path.  Here every band is a context in row-band mode (include/s2mv.h,
`s2mv_configure_band` ...): the cost-volume passes run on the band's own rows, the `usd` volume rows either
side that the two vertical passes read come from the neighbouring bands (nearest-neighbour halo exchange,
no collective), and the cheap O(W*H) stages run on the band plus an apron after the bands' WTA disparity
rows have been gathered (two small planes).

Two transports carry the same schedule:
  * `LocalBands`  — several band contexts in ONE process (any devices, also all on one GPU): halos move
                    with Tensor.copy_ (cudaMemcpyPeerAsync across devices).  Used by the single-GPU parity
                    test and as the reference for the distributed path.
  * `DistBand`    — one process per GPU under torchrun: halos move with torch.distributed point-to-point
                    ops (NCCL send/recv over NVLink on the box, gloo in the CPU tests), disparity rows with
                    one all_gather of row-padded planes.
The host-side schedule (who sends which rows to whom) is `halo_schedule` / `gather_rows`, plain Python on
shapes only, and is what tests/test_host_logic.py exercises with gloo on CPU tensors.
"""
import ctypes as C


# ------------------------------------------------------------------ partition
def row_bands(num_rows, num_bands, min_rows=1):
    """Contiguous [y0, y1) bands covering [0, num_rows): sizes differ by at most one row."""
    if num_bands < 1 or num_rows < num_bands * max(min_rows, 1):
        raise ValueError(f"cannot split {num_rows} rows into {num_bands} bands of >= {min_rows} rows")
    base, extra = divmod(num_rows, num_bands)
    out, y = [], 0
    for b in range(num_bands):
        h = base + (1 if b < extra else 0)
        out.append((y, y + h))
        y += h
    return out


def halo_schedule(band, num_bands):
    """Point-to-point transfers of band `band` after a horizontal/vertical pass: a list of
    (peer, send_side, recv_side) with side 0 = towards row 0, 1 = towards the last row.  Band b sends its
    first own rows up to b-1 (which receives them below its last row) and its last own rows down to b+1."""
    ops = []
    if band > 0:
        ops.append((band - 1, 0, 0))          # send my top rows up; receive my upper halo from above
    if band < num_bands - 1:
        ops.append((band + 1, 1, 1))          # send my bottom rows down; receive my lower halo from below
    return ops


def gather_rows(bands, local_y0, local_rows):
    """Which frame rows a sub-image [local_y0, local_y0 + local_rows) takes from which band's own rows:
    a list of (band, frame_y0, frame_y1)."""
    out = []
    lo, hi = local_y0, local_y0 + local_rows
    for b, (y0, y1) in enumerate(bands):
        a, e = max(lo, y0), min(hi, y1)
        if a < e:
            out.append((b, a, e))
    return out


# ---------------------------------------------------------- device memory view
class _DevMem:
    """Zero-copy torch view of raw device memory (the arena belongs to the C library)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _as_tensor(ptr, shape, typestr, device):
    import torch
    return torch.as_tensor(_DevMem(ptr, shape, typestr), device=torch.device("cuda", device))


class BandIpc(C.Structure):
    """s2mv_band_ipc (include/s2mv.h): what a neighbouring process needs to map a band's volumes."""
    _fields_ = [("mem", (C.c_ubyte * 64) * 3), ("frame_y0", C.c_int), ("frame_y1", C.c_int), ("local_y0", C.c_int),
                ("vlo", C.c_int), ("vrows", C.c_int), ("num_cols", C.c_int), ("lptot", C.c_int), ("frame_rows", C.c_int),
                ("device", C.c_int), ("halo_rows", C.c_int), ("fused", C.c_int)]


class RowBand:
    """One band context: thin mirror of the s2mv_band_* C ABI with torch views of its halo runs."""

    def __init__(self, device, band_y0, band_y1, apron=0, min_band_rows=0, fuse_vertical=None, **frame_params):
        """min_band_rows: rows of the frame's smallest band (0 = unknown: separate vertical passes, usd halo rows);
        fuse_vertical: True = one launch for both vertical passes whenever every band has 2*usd rows, False = never,
        None = where it is faster (tall bands)."""
        from . import Pipeline, default_params, lib
        _declare(lib())
        self.pipe = Pipeline(device)
        self.device = device
        L = self.pipe._L
        p = default_params(**frame_params)
        _check(L.s2mv_configure_band_ex(self.pipe._ctx, C.byref(p), int(band_y0), int(band_y1), int(apron),
                                        int(min_band_rows), -1 if fuse_vertical is None else int(bool(fuse_vertical))))
        self.frame = p
        self.y0, self.y1 = band_y0, band_y1
        v = [C.c_int() for _ in range(5)]
        _check(L.s2mv_band_info(self.pipe._ctx, *[C.byref(x) for x in v]))
        self.local_y0, self.local_rows, self.own_first, self.own_rows, self.halo_rows = [x.value for x in v]
        self.W = p.num_cols

    def close(self):
        self.pipe.close()

    def _st(self, stream):
        return self.pipe._stream(stream)

    def prepare(self, d_sbs_ptr, num_cols_sbs, stream=None):
        _check(self.pipe._L.s2mv_band_prepare(self.pipe._ctx, C.c_void_p(d_sbs_ptr), int(num_cols_sbs),
                                                              self._st(stream)))

    def run_pass(self, k, stream=None):
        _check(self.pipe._L.s2mv_band_pass(self.pipe._ctx, int(k), self._st(stream)))

    def halo(self, after_pass, view, side, recv):
        """float32 torch view of one halo run, or None at the frame's edge."""
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _check(self.pipe._L.s2mv_band_halo(self.pipe._ctx, after_pass, view, side, int(recv),
                                                           C.byref(ptr), C.byref(nbytes)))
        if not nbytes.value:
            return None
        return _as_tensor(ptr.value, (nbytes.value // 4,), "<f4", self.device)

    def disp_plane(self, view):
        ptr = C.c_void_p()
        _check(self.pipe._L.s2mv_band_disp(self.pipe._ctx, view, C.byref(ptr)))
        return _as_tensor(ptr.value, (self.local_rows, self.W), "<f4", self.device)

    # peer-to-peer halos (side 0 = the band above, 1 = the band below)
    def connect(self, side, other):
        """Same process: `other` is the neighbouring RowBand."""
        _check(self.pipe._L.s2mv_band_connect(self.pipe._ctx, int(side), other.pipe._ctx))

    def ipc_export(self):
        """bytes of this band's s2mv_band_ipc, to be carried to the neighbouring processes."""
        b = BandIpc()
        _check(self.pipe._L.s2mv_band_ipc_export(self.pipe._ctx, C.byref(b)))
        return bytes(b)

    def ipc_connect(self, side, blob):
        b = BandIpc.from_buffer_copy(blob)
        _check(self.pipe._L.s2mv_band_ipc_connect(self.pipe._ctx, int(side), C.byref(b)))

    def status(self, stream=None):
        _check(self.pipe._L.s2mv_band_status(self.pipe._ctx, self._st(stream)))

    def finish(self, d_disp_l, d_disp_r, d_interlaced, stream=None):
        """torch tensors (own_rows x W [x3]) or None."""
        g = lambda t: C.c_void_p(0 if t is None else t.data_ptr())  # noqa: E731
        _check(self.pipe._L.s2mv_band_finish(self.pipe._ctx, g(d_disp_l), g(d_disp_r), g(d_interlaced),
                                                             self._st(stream)))


def _check(status):
    from . import _check as chk
    chk(status)


def _declare(L):
    L.s2mv_band_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.s2mv_band_pass.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.s2mv_band_halo.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.s2mv_band_disp.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.s2mv_band_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.s2mv_configure_band.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.s2mv_configure_band_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.s2mv_band_info.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.s2mv_band_connect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.s2mv_band_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
    L.s2mv_band_ipc_connect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.s2mv_band_status.argtypes = [C.c_void_p, C.c_void_p]


# ------------------------------------------------------------- one process
class LocalBands:
    """All bands of a frame in this process; `devices` gives each band's GPU (repeat an index to run
    several bands on one GPU).  process() returns the assembled (disp_l, disp_r, interlaced) tensors."""

    def __init__(self, devices, apron=0, p2p=True, fuse_vertical=None, **frame_params):
        """p2p: the producing passes store the halo rows straight into the neighbouring band's volume (peer stores,
        epoch words on the stream); False: halos are copied between the passes (Tensor.copy_)."""
        import torch
        from . import default_params
        self.torch = torch
        p = default_params(**frame_params)
        self.H, self.W = p.num_rows, p.num_cols
        self.bands = row_bands(self.H, len(devices), min_rows=p.usd)
        mbr = min(y1 - y0 for y0, y1 in self.bands)
        self.ctx = [RowBand(d, y0, y1, apron, mbr, fuse_vertical, **frame_params) for d, (y0, y1) in zip(devices, self.bands)]
        self.p2p = p2p
        if p2p:
            for b, c in enumerate(self.ctx):
                if b > 0:
                    c.connect(0, self.ctx[b - 1])
                if b < len(self.ctx) - 1:
                    c.connect(1, self.ctx[b + 1])

    def close(self):
        for c in self.ctx:
            c.close()

    def _sync(self):
        for d in {c.device for c in self.ctx}:
            self.torch.cuda.synchronize(d)

    def _exchange(self, after_pass):
        n = len(self.ctx)
        for b, c in enumerate(self.ctx):
            for peer, send_side, _ in halo_schedule(b, n):
                for view in (0, 1):
                    src = c.halo(after_pass, view, send_side, False)
                    dst = self.ctx[peer].halo(after_pass, view, 1 - send_side, True)
                    if src is not None and dst is not None:   # nothing after pass 2 when the vertical passes are fused
                        dst.copy_(src)
        self._sync()

    def process(self, d_sbs_by_device, num_cols_sbs):
        """d_sbs_by_device: {device: uint8 tensor of the whole SBS frame on that device}."""
        torch = self.torch
        for c in self.ctx:
            with torch.cuda.device(c.device):
                st = torch.cuda.current_stream(c.device).cuda_stream
                c.prepare(d_sbs_by_device[c.device].data_ptr(), num_cols_sbs, st)
                c.run_pass(1, st)
        if not self.p2p:
            self._sync()
            self._exchange(1)
        for c in self.ctx:
            with torch.cuda.device(c.device):
                c.run_pass(2, torch.cuda.current_stream(c.device).cuda_stream)
        if not self.p2p:
            self._sync()
            self._exchange(2)
        for c in self.ctx:
            with torch.cuda.device(c.device):
                st = torch.cuda.current_stream(c.device).cuda_stream
                c.run_pass(3, st)
                c.run_pass(4, st)
        self._sync()
        # WTA disparities: every sub-image takes the rows it does not own from the bands that do
        planes = [[c.disp_plane(v) for v in (0, 1)] for c in self.ctx]
        for b, c in enumerate(self.ctx):
            for src_b, a, e in gather_rows(self.bands, c.local_y0, c.local_rows):
                if src_b == b:
                    continue
                s = self.ctx[src_b]
                for v in (0, 1):
                    planes[b][v][a - c.local_y0:e - c.local_y0].copy_(planes[src_b][v][a - s.local_y0:e - s.local_y0])
        self._sync()
        dev0 = self.ctx[0].device
        dl = torch.empty((self.H, self.W), dtype=torch.float32, device=f"cuda:{dev0}")
        dr = torch.empty_like(dl)
        out = torch.empty((self.H, self.W, 3), dtype=torch.uint8, device=f"cuda:{dev0}")
        for c in self.ctx:
            with torch.cuda.device(c.device):
                bl = torch.empty((c.own_rows, self.W), dtype=torch.float32, device=f"cuda:{c.device}")
                br = torch.empty_like(bl)
                bo = torch.empty((c.own_rows, self.W, 3), dtype=torch.uint8, device=f"cuda:{c.device}")
                c.finish(bl, br, bo, torch.cuda.current_stream(c.device).cuda_stream)
                torch.cuda.synchronize(c.device)
                dl[c.y0:c.y1].copy_(bl)
                dr[c.y0:c.y1].copy_(br)
                out[c.y0:c.y1].copy_(bo)
        self._sync()
        if self.p2p:
            for c in self.ctx:   # a wait that ran out means stale halo rows: an error, not a frame
                c.status(self.torch.cuda.current_stream(c.device).cuda_stream)
        return dl, dr, out


# --------------------------------------------------- one process per GPU
def exchange_halos_dist(send_recv, rank, world, dist):
    """send_recv(side, recv) -> list of tensors (one per view) or None at a frame edge.  Posts every
    send/recv of this rank's halo_schedule as one batch of point-to-point ops and waits for them."""
    ops = []
    for peer, send_side, recv_side in halo_schedule(rank, world):
        for t in send_recv(send_side, False) or []:
            if t is not None:
                ops.append(dist.P2POp(dist.isend, t, peer))
        for t in send_recv(recv_side, True) or []:
            if t is not None:
                ops.append(dist.P2POp(dist.irecv, t, peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def disparity_row_plan(bands, extents, rank):
    """Which disparity rows rank `rank` sends to and receives from which rank so that every sub-image
    (extents[b] = (local_y0, local_rows) of band b) holds the WTA disparities of all its rows: two lists of
    (peer, frame_y0, frame_y1), each ordered by peer and row (both ends of a pair post their ops in the same order)."""
    send, recv = [], []
    for b, (ly0, rows) in enumerate(extents):
        for src, a, e in gather_rows(bands, ly0, rows):
            if src == b:
                continue
            if src == rank:
                send.append((b, a, e))
            if b == rank:
                recv.append((src, a, e))
    return sorted(send), sorted(recv)


def exchange_disparity_rows_dist(planes, local_y0, own_y0, send, recv, dist):
    """planes: this band's sub-image disparity planes (one per view).  Own rows go out to the sub-images that
    hold them in their aprons, the apron rows come in -- one batch of point-to-point ops (a band's apron reaches
    its nearest neighbours only, unless bands are shorter than the apron; the plan covers both)."""
    ops = []
    for plane in planes:
        for peer, a, e in send:
            ops.append(dist.P2POp(dist.isend, plane[a - local_y0:e - local_y0], peer))
        for peer, a, e in recv:
            ops.append(dist.P2POp(dist.irecv, plane[a - local_y0:e - local_y0], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def allgather_rows_dist(own, bands, dist, torch):
    """own: this rank's (own_rows x W) tensor.  Returns the (H x W) frame assembled from every rank's own rows
    (bands may differ by one row: planes are padded to the tallest band for the collective)."""
    world = len(bands)
    hmax = max(y1 - y0 for y0, y1 in bands)
    pad = torch.zeros((hmax,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    pad[:own.shape[0]].copy_(own)
    allp = torch.empty((world * hmax,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    dist.all_gather_into_tensor(allp, pad)      # rank b's rows land at [b * hmax, (b + 1) * hmax)
    return torch.cat([allp[b * hmax:b * hmax + (y1 - y0)] for b, (y0, y1) in enumerate(bands)], dim=0)


class DistBand:
    """This rank's band of a frame split over the process group (one process per GPU)."""

    def __init__(self, device, rank, world, apron=0, transport="p2p", fuse_vertical=None, **frame_params):
        """transport "p2p": neighbours' volumes mapped through CUDA IPC, halo rows stored over NVLink by the
        producing kernels; "nccl": halos sent with torch.distributed point-to-point ops between the passes."""
        import torch
        import torch.distributed as dist
        from . import default_params
        self.torch, self.dist = torch, dist
        self.transport = transport if world > 1 else "none"
        p = default_params(**frame_params)
        self.H, self.W = p.num_rows, p.num_cols
        self.rank, self.world = rank, world
        self.bands = row_bands(self.H, world, min_rows=p.usd)
        y0, y1 = self.bands[rank]
        mbr = min(b1 - b0 for b0, b1 in self.bands)
        self.ctx = RowBand(device, y0, y1, apron, mbr, fuse_vertical, **frame_params)
        c = self.ctx
        self.out_l = torch.empty((c.own_rows, self.W), dtype=torch.float32, device=f"cuda:{device}")
        self.out_r = torch.empty_like(self.out_l)
        self.out_i = torch.empty((c.own_rows, self.W, 3), dtype=torch.uint8, device=f"cuda:{device}")
        if world > 1:
            extents = [None] * world
            dist.all_gather_object(extents, (c.local_y0, c.local_rows))
            self._disp_send, self._disp_recv = disparity_row_plan(self.bands, extents, rank)
        if self.transport == "p2p":
            blobs = [None] * world
            dist.all_gather_object(blobs, c.ipc_export())
            if rank > 0:
                c.ipc_connect(0, blobs[rank - 1])
            if rank < world - 1:
                c.ipc_connect(1, blobs[rank + 1])
            dist.barrier()

    def close(self):
        self.ctx.close()

    def check(self):
        """Raise if any wait of this band ran out before its neighbour arrived (synchronises the stream)."""
        self.ctx.status(self.torch.cuda.current_stream().cuda_stream)

    def process(self, d_sbs, num_cols_sbs, phases=None, check=True):
        """d_sbs: the whole SBS frame on this rank's GPU.  Returns this band's rows of
        (disp_l, disp_r, interlaced).  `phases`: a dict that receives the device time (ms) of every phase of
        this call.  check=True (default): synchronise and raise if a halo wait of this frame ran out before the
        neighbour arrived (its rows would be stale); check=False leaves everything enqueued on torch's current
        stream -- the caller then calls check() itself (the next frame's calls fail anyway once a wait has
        run out)."""
        torch, dist, c = self.torch, self.dist, self.ctx
        st = torch.cuda.current_stream().cuda_stream
        marks = []

        def mark(name):
            if phases is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        c.prepare(d_sbs.data_ptr(), num_cols_sbs, st)
        mark("prepare")
        c.run_pass(1, st)
        mark("pass1")
        if self.transport == "nccl":
            exchange_halos_dist(lambda side, recv: [c.halo(1, v, side, recv) for v in (0, 1)], self.rank, self.world, dist)
        mark("halo1")
        c.run_pass(2, st)
        mark("pass2")
        if self.transport == "nccl":
            exchange_halos_dist(lambda side, recv: [c.halo(2, v, side, recv) for v in (0, 1)], self.rank, self.world, dist)
        mark("halo2")
        c.run_pass(3, st)
        c.run_pass(4, st)
        mark("pass3+4")
        if self.world > 1:
            # the apron rows of the WTA disparity planes come from the bands that own them (NCCL send/recv)
            exchange_disparity_rows_dist([c.disp_plane(v) for v in (0, 1)], c.local_y0, c.y0, self._disp_send,
                                         self._disp_recv, dist)
        mark("gather")
        c.finish(self.out_l, self.out_r, self.out_i, st)
        mark("finish")
        if phases is not None:
            marks[-1][1].synchronize()
            for (_, a), (name, b) in zip(marks, marks[1:]):
                phases[name] = phases.get(name, 0.0) + a.elapsed_time(b)
        if check and self.transport == "p2p":
            self.check()
        return self.out_l, self.out_r, self.out_i
