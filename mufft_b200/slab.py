"""Slab-decomposed 2D C2C transform of ONE large [Ny][Nx] array over the ranks of a
torch.distributed group (BASELINE configs[4]).

The reference has no distributed path: `mufft_execute_plan_2d` (ref fft.c:852-989) runs the row
pass, then the column pass, on one array in one address space.  Here rank g owns rows
[g*Ny/G, (g+1)*Ny/G).  The row pass is local; one all-to-all block transpose (NCCL over
NVLink/NVSwitch: rank g sends the (Ny/G x Nx/G) tile of columns [p*Nx/G, (p+1)*Nx/G) to rank p)
turns row slabs into column slabs; the column pass is local again.  Result layout
("transposed-out", the usual contract of slab FFTs): rank g holds columns
[g*Nx/G, (g+1)*Nx/G) of the spectrum as a row-major [Ny][Nx/G] array.  `natural=True` adds a
second all-to-all that brings the spectrum back to row slabs [Ny/G][Nx].

The local transforms go through the C ABI (`mufft_execute_plan_1d_batch_async` for rows,
`mufft_execute_plan_1d_columns` for columns, both hand-written CUDA); `local_rows` /
`local_cols` can be injected so the exchange logic is testable on CPU with the gloo backend.

Two exchange paths:

* ``exchange="p2p"`` (default on CUDA when the ranks share a node): the column-slab buffers are
  plain device allocations mapped into every rank through CUDA IPC, and the row kernel itself
  stores element x of every row straight into the slab of device x // (Nx/G)
  (`mufft_execute_plan_1d_batch_scatter_async`: the kernel's stores ARE the all-to-all and travel
  over NVLink).  No pack pass, no NCCL call on the data path; one tiny all-reduce orders the
  devices before the column pass.
* ``exchange="nccl"``: the row pass is cut into `chunks` blocks of rows; block c is packed into
  per-peer tiles and its all-to-all (NCCL) runs on a side stream while block c+1 is being
  transformed.  This is also the path of the gloo CPU tests.

With `natural=True` the return exchange is pure communication (NCCL) in both cases.
"""
import ctypes

import torch
import torch.distributed as dist


class _DeviceArray:
    """A raw device allocation exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<c8", "data": (int(ptr), False), "version": 3}


class SlabFFT2D:
    def __init__(self, nx, ny, direction, group=None, device=None, chunks=4, local_rows=None, local_cols=None, exchange=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        g = self.world
        if nx % g or ny % g:
            raise ValueError("Nx and Ny must be divisible by the number of ranks")
        self.nx, self.ny, self.direction = nx, ny, direction
        self.rows_local, self.cols_local = ny // g, nx // g
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.chunks = max(1, min(chunks, self.rows_local))
        while self.rows_local % self.chunks:
            self.chunks -= 1
        self._plans = None
        if local_rows is None or local_cols is None:
            import mufft_b200 as m  # the CUDA library; fails loudly when it is not built
            if self.device.type != "cuda":
                raise RuntimeError("SlabFFT2D runs its transforms on a CUDA device (there is no CPU path)")
            self._m = m
            # plans (twiddle tables, scratch) live on the device that is current when they are created: make it ours
            torch.cuda.set_device(self.device)
            if m.mufft_cuda_set_device(self.device.index if self.device.index is not None else torch.cuda.current_device()):
                raise RuntimeError(m.mufft_cuda_last_error().decode())
            px = m.mufft_create_plan_1d_c2c(nx, direction, 0)
            py = m.mufft_create_plan_1d_c2c(ny, direction, 0)
            if not px or not py:
                raise RuntimeError("plan creation failed: " + m.mufft_cuda_last_error().decode())
            self._plans = (px, py)
            local_rows, local_cols = self._cuda_rows, self._cuda_cols
        self.local_rows, self.local_cols = local_rows, local_cols
        self._side = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        if exchange is None:
            exchange = "p2p" if (self._plans is not None and g > 1) else "nccl"
        if exchange == "p2p" and (self._plans is None or g & (g - 1)):
            raise ValueError("the p2p exchange needs the CUDA library and a power-of-two number of ranks")
        self.exchange = exchange
        self._ipc = None  # (own allocation, [peer pointers], flag tensor)

    def close(self):
        if self._ipc is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # nobody may still be writing into a slab that is about to go away
            own, peers, _ = self._ipc
            for r, p in enumerate(peers):
                if r != self.rank:
                    self._m.mufft_cuda_ipc_close(p)
            self._m.mufft_cuda_device_free(own)
            self._ipc = None
        if self._plans:
            for p in self._plans:
                self._m.mufft_free_plan_1d(p)
            self._plans = None

    # ---- local passes through the C ABI --------------------------------------------------------
    def _use_current_stream(self):
        self._m.mufft_cuda_set_stream(torch.cuda.current_stream(self.device).cuda_stream or None)

    def _cuda_rows(self, out, inp):
        """out[r, :] = FFT(inp[r, :]) for a block of rows ([rows][Nx] complex64 each)."""
        self._use_current_stream()
        err = self._m.mufft_execute_plan_1d_batch_async(self._plans[0], out.data_ptr(), inp.data_ptr(), inp.shape[0], 0, 0)
        if err:
            raise RuntimeError(self._m.mufft_cuda_last_error().decode())

    def _cuda_cols(self, out, inp):
        """out[:, c] = FFT(inp[:, c]) for a [Ny][W] column slab."""
        self._use_current_stream()
        w = inp.shape[1]
        err = self._m.mufft_execute_plan_1d_columns(self._plans[1], out.data_ptr(), inp.data_ptr(), w, w * 8, w * 8)
        if err:
            raise RuntimeError(self._m.mufft_cuda_last_error().decode())

    # ---- the distributed transform ---------------------------------------------------------------
    def _p2p_setup(self):
        """Allocates this rank's column slab as a shareable device allocation and maps everybody else's."""
        m = self._m
        nbytes = self.ny * self.cols_local * 8
        own = m.mufft_cuda_device_alloc(nbytes)
        if not own:
            raise RuntimeError("cudaMalloc failed: " + m.mufft_cuda_last_error().decode())
        handle = ctypes.create_string_buffer(64)
        if m.mufft_cuda_ipc_export(handle, own):
            raise RuntimeError(m.mufft_cuda_last_error().decode())
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=self.group)
        peers = []
        for r, h in enumerate(handles):
            if r == self.rank:
                peers.append(own)
            else:
                p = m.mufft_cuda_ipc_import(ctypes.create_string_buffer(h, 64))
                if not p:
                    raise RuntimeError("cudaIpcOpenMemHandle failed: " + m.mufft_cuda_last_error().decode())
                peers.append(p)
        self._ipc = (own, peers, torch.zeros(1, dtype=torch.float32, device=self.device))
        self._peer_array = (ctypes.c_void_p * self.world)(*peers)

    def buffers(self):
        """Work buffers (allocate once, reuse across executes).  With the p2p exchange `recv` aliases the plan-owned,
        IPC-shared column slab: it must not be used after close()."""
        c64 = dict(dtype=torch.complex64, device=self.device)
        g, rl, cl = self.world, self.rows_local, self.cols_local
        if self.exchange == "p2p":
            if self._ipc is None:
                self._p2p_setup()
            recv = torch.as_tensor(_DeviceArray(self._ipc[0], (self.ny, cl)), device=self.device)
            return {"recv": recv, "send": None, "rows": None}
        return {
            "rows": torch.empty((rl, self.nx), **c64),          # row-transformed slab
            "send": torch.empty((g, rl, cl), **c64),            # per-peer tiles, contiguous
            "recv": torch.empty((self.ny, cl), **c64),          # = [G][rows_local][cols_local]: the column slab
        }

    def execute(self, out, inp, buf=None, natural=False):
        """inp: this rank's [Ny/G][Nx] complex64 row slab.  out: [Ny][Nx/G] (transposed-out) or, with
        natural=True, [Ny/G][Nx].  Returns out."""
        g, rl, cl = self.world, self.rows_local, self.cols_local
        buf = buf or self.buffers()
        if self.exchange == "p2p":
            return self._execute_p2p(out, inp, buf, natural)
        rows, send, recv = buf["rows"], buf["send"], buf["recv"]
        recv_tiles = recv.view(g, rl, cl)
        rc = rl // self.chunks
        works = []
        for c in range(self.chunks):
            lo, hi = c * rc, (c + 1) * rc
            self.local_rows(rows[lo:hi], inp[lo:hi])
            # pack: tile for peer p = columns [p*cl, (p+1)*cl) of this block of rows
            send[:, lo:hi, :].copy_(rows[lo:hi].view(rc, g, cl).permute(1, 0, 2))
            if g == 1:
                recv_tiles[:, lo:hi, :].copy_(send[:, lo:hi, :])
                continue
            ins = [send[p, lo:hi, :] for p in range(g)]
            outs = [recv_tiles[s, lo:hi, :] for s in range(g)]
            if self._side is not None:
                self._side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self._side):
                    works.append(dist.all_to_all(outs, ins, group=self.group, async_op=True))
            elif dist.get_backend(self.group) == "gloo":
                self._gloo_all_to_all(outs, ins)
            else:
                dist.all_to_all(outs, ins, group=self.group)
        for w in works:
            w.wait()
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        if not natural:
            self.local_cols(out, recv)
            return out
        self.local_cols(recv, recv)
        # second exchange: back to row slabs.  Tile for peer p = rows [p*rl, (p+1)*rl) of the column slab.
        back = send  # [G][rl][cl], reused
        if g == 1:
            back.copy_(recv_tiles)
        else:
            ins = [recv_tiles[p] for p in range(g)]
            outs = [back[s] for s in range(g)]
            if dist.get_backend(self.group) == "gloo":
                self._gloo_all_to_all(outs, ins)
            else:
                dist.all_to_all(outs, ins, group=self.group)
        out.view(rl, g, cl).copy_(back.permute(1, 0, 2))
        return out

    def _execute_p2p(self, out, inp, buf, natural):
        """Row pass with the exchange fused into its stores, device barrier, column pass."""
        m, g, rl, cl = self._m, self.world, self.rows_local, self.cols_local
        recv = buf["recv"]
        flag = self._ipc[2]
        # (1) nobody may still be reading its slab from the previous execute when we overwrite it
        dist.all_reduce(flag, group=self.group)
        self._use_current_stream()
        err = m.mufft_execute_plan_1d_batch_scatter_async(self._plans[0], self._peer_array, g, self.rank * rl, inp.data_ptr(), rl, 0)
        if err:
            raise RuntimeError(m.mufft_cuda_last_error().decode())
        # (2) every rank's row kernel (and with it its remote stores) has completed
        dist.all_reduce(flag, group=self.group)
        if not natural:
            self.local_cols(out, recv)
            return out
        self.local_cols(recv, recv)
        recv_tiles = recv.view(g, rl, cl)
        back = buf.get("back")
        if back is None:
            back = buf["back"] = torch.empty((g, rl, cl), dtype=torch.complex64, device=self.device)
        dist.all_to_all([back[s] for s in range(g)], [recv_tiles[p] for p in range(g)], group=self.group)
        out.view(rl, g, cl).copy_(back.permute(1, 0, 2))
        return out

    def _gloo_all_to_all(self, outs, ins):
        """gloo has no all_to_all: pairwise exchange (CPU tests only)."""
        g, me = self.world, self.rank
        outs[me].copy_(ins[me])
        reqs = []
        for p in range(g):
            if p != me:
                reqs.append(dist.isend(ins[p].contiguous(), p, group=self.group))
        for s in range(g):
            if s != me:
                tmp = torch.empty_like(outs[s])
                dist.recv(tmp, s, group=self.group)
                outs[s].copy_(tmp)
        for r in reqs:
            r.wait()


class SlabRFFT2D:
    """Slab-decomposed 2D REAL transform pair of one large [Ny][Nx] array (BASELINE configs[4], R2C).

    forward (R2C, ref fft.c:918-989 order: real row transforms, then complex column transforms):
      rank g holds real rows [g*Ny/G, (g+1)*Ny/G) as float32 [Ny/G][Nx].  The local row pass is the
      library's R2C kernel (Nx/2+1 bins per row), which HALVES the exchanged volume; one all-to-all
      turns the row slabs of the half spectrum into column slabs; the local column pass is C2C.
      Result, transposed-out: rank g holds columns [g*H/G, (g+1)*H/G) of the H = Nx/2 lower bins as
      [Ny][H/G] complex64; the Nyquist column (bin Nx/2) rides with the LAST rank, whose slab is
      [Ny][H/G + 1].
    inverse (C2R, ref fft.c:861-916 order: columns first, then the real row transforms): takes
      exactly that column-slab layout, runs the inverse column pass, the reverse all-to-all, and the
      library's C2R row kernel; result float32 [Ny/G][Nx], unnormalised like the reference
      (inverse(forward(x)) = Nx*Ny*x).

    The exchange is NCCL all-to-all with per-peer tiles (uneven for the last rank), block c in
    flight on a side stream while block c+1 of rows is transformed; the local passes go through
    the C ABI.  `local_*` hooks let the exchange logic run on CPU under gloo.

    ``exchange="p2p"`` (forward only, opt-in until it has run on hardware; needs Nx/2 within the
    library's single-kernel row limit and a power-of-two number of ranks): the column slabs are device
    allocations mapped into every rank through CUDA IPC and the R2C row kernel's epilogue stores bin k
    of every row straight into the slab of rank min(k // (H/G), G-1)
    (`mufft_execute_plan_1d_batch_scatter_async` on the R2C plan): no row buffer, no pack and unpack
    copies, no NCCL call on the data path.  The inverse keeps the NCCL exchange.
    """

    def __init__(self, nx, ny, group=None, device=None, chunks=4, local_r2c_rows=None, local_c2r_rows=None, local_cols=None, exchange=None):
        self.group = group
        self.world = g = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if nx % (2 * g) or ny % g or nx < 4:
            raise ValueError("Nx/2 and Ny must be divisible by the number of ranks")
        self.nx, self.ny, self.h = nx, ny, nx // 2
        self.rows_local, self.cols_even = ny // g, self.h // g
        self.widths = [self.cols_even] * (g - 1) + [self.cols_even + 1]  # the Nyquist column goes to the last rank
        self.cols_local = self.widths[self.rank]
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.chunks = max(1, min(chunks, self.rows_local))
        while self.rows_local % self.chunks:
            self.chunks -= 1
        self._plans = None
        if local_r2c_rows is None or local_c2r_rows is None or local_cols is None:
            import mufft_b200 as m
            if self.device.type != "cuda":
                raise RuntimeError("SlabRFFT2D runs its transforms on a CUDA device (there is no CPU path)")
            self._m = m
            torch.cuda.set_device(self.device)  # plans live on the device that is current when they are created
            if m.mufft_cuda_set_device(self.device.index if self.device.index is not None else torch.cuda.current_device()):
                raise RuntimeError(m.mufft_cuda_last_error().decode())
            plans = (m.mufft_create_plan_1d_r2c(nx, 0), m.mufft_create_plan_1d_c2r(nx, 0),
                     m.mufft_create_plan_1d_c2c(ny, m.MUFFT_FORWARD, 0), m.mufft_create_plan_1d_c2c(ny, m.MUFFT_INVERSE, 0))
            if not all(plans):
                raise RuntimeError("plan creation failed: " + m.mufft_cuda_last_error().decode())
            self._plans = plans
            local_r2c_rows, local_c2r_rows, local_cols = self._cuda_r2c_rows, self._cuda_c2r_rows, self._cuda_cols
        self.local_r2c_rows, self.local_c2r_rows, self.local_cols = local_r2c_rows, local_c2r_rows, local_cols
        self._side = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        exchange = exchange or "nccl"
        if exchange == "p2p" and (self._plans is None or g & (g - 1)):
            raise ValueError("the p2p exchange needs the CUDA library and a power-of-two number of ranks")
        self.exchange = exchange
        self._ipc = None  # (own slab allocation, [peer pointers], flag tensor)

    def close(self):
        if self._ipc is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # nobody may still be writing into a slab that is about to go away
            own, peers, _ = self._ipc
            for r, p in enumerate(peers):
                if r != self.rank:
                    self._m.mufft_cuda_ipc_close(p)
            self._m.mufft_cuda_device_free(own)
            self._ipc = None
        if self._plans:
            for p in self._plans:
                self._m.mufft_free_plan_1d(p)
            self._plans = None

    def _p2p_setup(self):
        """This rank's column slab [Ny][cols_local] as a shareable device allocation; maps everybody else's."""
        m = self._m
        own = m.mufft_cuda_device_alloc(self.ny * self.cols_local * 8)
        if not own:
            raise RuntimeError("cudaMalloc failed: " + m.mufft_cuda_last_error().decode())
        handle = ctypes.create_string_buffer(64)
        if m.mufft_cuda_ipc_export(handle, own):
            raise RuntimeError(m.mufft_cuda_last_error().decode())
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=self.group)
        peers = []
        for r, h in enumerate(handles):
            if r == self.rank:
                peers.append(own)
            else:
                p = m.mufft_cuda_ipc_import(ctypes.create_string_buffer(h, 64))
                if not p:
                    raise RuntimeError("cudaIpcOpenMemHandle failed: " + m.mufft_cuda_last_error().decode())
                peers.append(p)
        self._ipc = (own, peers, torch.zeros(1, dtype=torch.float32, device=self.device))
        self._peer_array = (ctypes.c_void_p * self.world)(*peers)

    def _use_current_stream(self):
        self._m.mufft_cuda_set_stream(torch.cuda.current_stream(self.device).cuda_stream or None)

    def _check(self, err):
        if err:
            raise RuntimeError(self._m.mufft_cuda_last_error().decode())

    def _cuda_r2c_rows(self, out, inp):
        """out[r, :Nx/2+1] = R2C(inp[r, :]); out rows are Nx/2+1 complex apart."""
        self._use_current_stream()
        self._check(self._m.mufft_execute_plan_1d_batch_async(self._plans[0], out.data_ptr(), inp.data_ptr(), inp.shape[0],
                                                              out.stride(0) * 8, inp.stride(0) * 4))

    def _cuda_c2r_rows(self, out, inp):
        self._use_current_stream()
        self._check(self._m.mufft_execute_plan_1d_batch_async(self._plans[1], out.data_ptr(), inp.data_ptr(), inp.shape[0],
                                                              out.stride(0) * 4, inp.stride(0) * 8))

    def _cuda_cols(self, out, inp, direction):
        self._use_current_stream()
        w = inp.shape[1]
        plan = self._plans[2] if direction < 0 else self._plans[3]
        self._check(self._m.mufft_execute_plan_1d_columns(plan, out.data_ptr(), inp.data_ptr(), w, out.stride(0) * 8, inp.stride(0) * 8))

    def buffers(self):
        c64 = dict(dtype=torch.complex64, device=self.device)
        g, rl = self.world, self.rows_local
        if self.exchange == "p2p":
            if self._ipc is None:
                self._p2p_setup()
            slab = torch.as_tensor(_DeviceArray(self._ipc[0], (self.ny, self.cols_local)), device=self.device)
        else:
            slab = torch.empty((self.ny, self.cols_local), **c64)
        return {
            "rows": torch.empty((rl, self.h + 1), **c64),                              # half spectra of the local rows
            "send": [torch.empty((rl, w), **c64) for w in self.widths],                # tile for peer p: its columns of my rows
            "recv": [torch.empty((rl, self.cols_local), **c64) for _ in range(g)],     # tile from peer s: my columns of its rows
            "slab": slab,
        }

    def _col_range(self, p):
        lo = p * self.cols_even
        return lo, lo + self.widths[p]

    def _all_to_all(self, outs, ins, works):
        if self.world == 1:
            outs[0].copy_(ins[0])
        elif dist.get_backend(self.group) == "gloo":
            SlabFFT2D._gloo_all_to_all(self, outs, ins)
        elif self._side is not None and works is not None:
            self._side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._side):
                works.append(dist.all_to_all(outs, ins, group=self.group, async_op=True))
        else:
            dist.all_to_all(outs, ins, group=self.group)

    def _drain(self, works):
        for w in works:
            w.wait()
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)

    def forward(self, out, inp, buf=None):
        """inp float32 [Ny/G][Nx] -> out complex64 [Ny][cols_local] (this rank's columns of the half spectrum)."""
        g, rl = self.world, self.rows_local
        buf = buf or self.buffers()
        if self.exchange == "p2p":
            flag = self._ipc[2]
            dist.all_reduce(flag, group=self.group)  # nobody still reads its slab from the previous call
            self._use_current_stream()
            self._check(self._m.mufft_execute_plan_1d_batch_scatter_async(self._plans[0], self._peer_array, g, self.rank * rl,
                                                                          inp.data_ptr(), rl, inp.stride(0) * 4))
            dist.all_reduce(flag, group=self.group)  # every rank's row kernel, and with it its remote stores, has completed
            self.local_cols(out, buf["slab"], -1)
            return out
        rows, send, recv = buf["rows"], buf["send"], buf["recv"]
        rc = rl // self.chunks
        works = []
        for c in range(self.chunks):
            lo, hi = c * rc, (c + 1) * rc
            self.local_r2c_rows(rows[lo:hi], inp[lo:hi])
            for p in range(g):
                a, b = self._col_range(p)
                send[p][lo:hi].copy_(rows[lo:hi, a:b])
            self._all_to_all([recv[s][lo:hi] for s in range(g)], [send[p][lo:hi] for p in range(g)], works)
        self._drain(works)
        slab = buf["slab"]
        for s in range(g):
            slab[s * rl:(s + 1) * rl].copy_(recv[s])
        self.local_cols(out, slab, -1)
        return out

    def inverse(self, out, inp, buf=None):
        """inp complex64 [Ny][cols_local] (the layout forward() produced) -> out float32 [Ny/G][Nx]."""
        g, rl = self.world, self.rows_local
        buf = buf or self.buffers()
        rows, send, recv, slab = buf["rows"], buf["send"], buf["recv"], buf["slab"]
        self.local_cols(slab, inp, +1)
        # tile for peer p = its rows of my columns; arrives as my rows of peer s's columns
        ins = [slab[p * rl:(p + 1) * rl] for p in range(g)]
        outs = [send[s] for s in range(g)]  # send[s] has exactly the shape [rl][widths[s]]
        works = []
        self._all_to_all(outs, ins, works)
        self._drain(works)
        for s in range(g):
            a, b = self._col_range(s)
            rows[:, a:b].copy_(outs[s])
        self.local_c2r_rows(out, rows)
        return out
