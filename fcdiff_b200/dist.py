"""
Edge sharding over the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch; ``gloo``
in the CPU tests).  Rank r owns the contiguous edge range
``[r*chunkC, min((r+1)*chunkC, C))`` of ``b`` / ``bt`` (rows) for the
edge-parallel kernels K2, K3, K4, and the patient range
``[r*chunkU, min((r+1)*chunkU, U))`` for the region sweep K2b, which is
sequential over regions but independent across patients.  Exchanges per EM
iteration:

* all-gather of ``lq_F`` / ``q_F`` (24*C bytes each) after K2,
* all-gather of ``lq_R`` / ``q_R`` (16*N*U bytes each) after K2b,
* all-reduce (sum) of <= 6 doubles after K3a, per K3b evaluation and after K4:
  on GPUs a one-shot exchange over NVLink peer memory whose result lands in
  mapped host memory (:class:`PeerWindow`, csrc/fcd_comm.cu) -- these exchanges
  are pure latency, and NCCL + a download per call cost more than the kernels
  they follow; NCCL / gloo otherwise.

The reference has no distributed code; the sharded fit must equal the
single-device fit up to the order of the fp64 partial sums.
"""
import torch
import torch.distributed as dist


def _ceil_div(a, b):
    return (a + b - 1) // b


class PeerWindow(object):
    """All-reduce of <= 8 doubles over NVLink peer memory, delivered to the host
    (``fcd_allreduce_small`` / ``fcd_wait_result``).  Every rank creates a window
    in its device memory; the CUDA IPC handles travel once through the process
    group; afterwards an exchange is one kernel launch and a host-side spin on
    mapped pinned memory.  ``world == 1`` (no group) publishes reduction results
    to the host the same way."""

    TIMEOUT_MS = 30000

    def __init__(self, group=None, rank=0, world=1):
        import ctypes
        import numpy as np
        from . import _lib
        self._ct = ctypes
        self.lib = _lib.load()
        self._check = _lib.check
        (self.rank, self.world) = (int(rank), int(world))
        self.max_vals = int(self.lib.fcd_comm_max_vals())
        if self.world > int(self.lib.fcd_comm_max_world()):
            raise ValueError("PeerWindow: world size %d exceeds %d" % (self.world, self.lib.fcd_comm_max_world()))
        self.seq = 0
        self._opened = []
        self._own = ctypes.c_void_p(0)
        self._stream = None           # the ONE stream every exchange of this window is issued on
        self.windows = (ctypes.c_void_p * max(self.world, 1))()
        self.ok = True
        self.why = ""
        if self.world > 1:
            # No exception may escape between two collectives (the other ranks would hang in the next
            # one): every step records its outcome, the ranks exchange it, and all of them either
            # keep the window or fall back together (EdgeShards.peer_window).
            hb = int(self.lib.fcd_comm_handle_bytes())
            handle = ctypes.create_string_buffer(hb)
            try:
                self._check(self.lib.fcd_comm_window_create(ctypes.byref(self._own)), "fcd_comm_window_create")
                self._check(self.lib.fcd_comm_window_export(self._own, handle), "fcd_comm_window_export")
            except Exception as exc:       # no IPC support, out of memory, ...
                (self.ok, self.why) = (False, str(exc))
            handles = [None] * self.world
            dist.all_gather_object(handles, (self.ok, bytes(handle.raw)), group=group)
            if all(h[0] for h in handles):
                for r in range(self.world):
                    if r == self.rank:
                        self.windows[r] = self._own
                        continue
                    peer = ctypes.c_void_p(0)
                    buf = ctypes.create_string_buffer(handles[r][1], hb)
                    try:
                        self._check(self.lib.fcd_comm_window_open(buf, ctypes.byref(peer)), "fcd_comm_window_open")
                    except Exception as exc:   # a rank on another node, no peer access between the two GPUs
                        (self.ok, self.why) = (False, str(exc))
                        break
                    self._opened.append(peer)
                    self.windows[r] = peer
            else:
                (self.ok, self.why) = (False, self.why or "a peer rank could not create its window")
            # every window is open everywhere before the first exchange -- or nobody uses them
            flags = [None] * self.world
            dist.all_gather_object(flags, self.ok, group=group)
            if not all(flags):
                (self.ok, self.why) = (False, self.why or "a peer rank could not open the windows")
        self._result = ctypes.c_void_p(0)
        self._check(self.lib.fcd_host_result_alloc(ctypes.byref(self._result)), "fcd_host_result_alloc")
        self._out = np.zeros(self.max_vals, dtype=np.float64)
        self._out_p = ctypes.c_void_p(self._out.ctypes.data)

    def bind_stream(self, stream):
        """The two-parity slot scheme (csrc/fcd_comm.cuh) is safe only if a rank issues its exchanges
        in one order on ONE stream -- those launched here and those run inside the solver's evaluation
        kernels alike: the window binds to the first stream it sees."""
        sv = getattr(stream, "value", stream)
        if self.world == 1:
            return                                 # publication only: no peers, no slots to protect
        if self._stream is None:
            self._stream = sv
        elif sv != self._stream:
            raise RuntimeError("PeerWindow: exchanges of one window must all be issued on the same stream")

    def allreduce(self, vec_dev, n, stream):
        """Sums ``vec_dev[:n]`` (device, float64) over the ranks in place and returns
        the sums as a fresh host array."""
        return self.allreduce_end(self.allreduce_begin(vec_dev, n, stream))

    def allreduce_begin(self, vec_dev, n, stream):
        """Enqueues the exchange (one kernel) and returns a ticket for ``allreduce_end``.  At most
        one exchange may be outstanding per window (there is one host result block)."""
        self.bind_stream(stream)
        self.seq += 1
        rc = self.lib.fcd_allreduce_small(self._ct.c_void_p(vec_dev.data_ptr()), n, self.windows, self.rank,
                                          self.world, self.seq, self._result, stream)
        if rc != 0:
            self._check(rc, "fcd_allreduce_small")
        return (self.seq, n)

    def allreduce_keep(self, vec_dev, n, keep0, nkeep, stream):
        """``allreduce`` of ``vec_dev[:n]``; returns the n sums followed by this rank's own
        ``vec_dev[keep0:keep0 + nkeep]`` as they were before the sum (one publication, one wait)."""
        self.bind_stream(stream)
        self.seq += 1
        rc = self.lib.fcd_allreduce_small_keep(self._ct.c_void_p(vec_dev.data_ptr()), n, keep0, nkeep, self.windows,
                                               self.rank, self.world, self.seq, self._result, stream)
        if rc != 0:
            self._check(rc, "fcd_allreduce_small_keep")
        return self.allreduce_end((self.seq, n + nkeep))

    def allreduce_end(self, ticket):
        (seq, n) = ticket
        rc = self.lib.fcd_wait_result(self._result, n, seq, self._out_p, self.TIMEOUT_MS)
        if rc != 0:
            self._check(rc, "fcd_wait_result")
        return self._out[:n].copy()

    def close(self):
        for p in self._opened:
            self.lib.fcd_comm_window_close(p)
        self._opened = []
        if self._own:
            self.lib.fcd_comm_window_destroy(self._own)
            self._own = self._ct.c_void_p(0)
        if self._result:
            self.lib.fcd_host_result_free(self._result)
            self._result = self._ct.c_void_p(0)


class EdgeShards(object):
    """Shard map + the three collectives of the sharded fit."""

    def __init__(self, group=None, rank=None, world=None):
        self.group = group
        if rank is None:
            rank = dist.get_rank(group)
        if world is None:
            world = dist.get_world_size(group)
        self.rank = int(rank)
        self.world = int(world)
        self._peer = None             # PeerWindow, created on first use (NCCL groups on CUDA only)
        self._peer_tried = False

    def peer_window(self):
        """The NVLink peer window of this group, or None (gloo / CPU groups)."""
        if not self._peer_tried:
            self._peer_tried = True
            if torch.cuda.is_available() and dist.is_initialized() and dist.get_backend(self.group) == "nccl":
                pw = PeerWindow(self.group, self.rank, self.world)
                if pw.ok:                  # all ranks agree (PeerWindow.__init__)
                    self._peer = pw
                else:                      # ranks on several nodes / no P2P: NCCL all-reduce + download instead
                    self.peer_fallback_reason = pw.why
                    pw.close()
        return self._peer

    def reduce_read(self, res, n=None, stream=None):
        """Sums the device vector ``res.dev`` (a ``_dev.SmallResult``) over the ranks in
        place and returns the sums on the host: one peer-memory exchange, or an
        all-reduce + download where no peer window exists."""
        n = res.dev.numel() if n is None else n
        pw = self.peer_window()
        if pw is not None and n <= pw.max_vals:
            from . import _dev
            return pw.allreduce(res.dev, n, _dev.stream() if stream is None else stream)
        dist.all_reduce(res.dev, op=dist.ReduceOp.SUM, group=self.group)
        return res.read(stream)

    def reduce_read_keep(self, res, n, keep0, nkeep, stream=None):
        """``reduce_read`` of ``res.dev[:n]`` that also returns this rank's own ``res.dev[keep0:keep0 + nkeep]``
        (before the sum) behind the sums, or None where no peer window exists."""
        pw = self.peer_window()
        if pw is None or n + nkeep > pw.max_vals:
            return None
        from . import _dev
        return pw.allreduce_keep(res.dev, n, keep0, nkeep, _dev.stream() if stream is None else stream)

    def reduce_begin(self, res, n=None, stream=None):
        """``reduce_read`` in two halves (peer window only): enqueue now, collect later."""
        n = res.dev.numel() if n is None else n
        pw = self.peer_window()
        if pw is None or n > pw.max_vals:
            return None
        from . import _dev
        return (pw, pw.allreduce_begin(res.dev, n, _dev.stream() if stream is None else stream))

    @staticmethod
    def reduce_end(handle):
        (pw, ticket) = handle
        return pw.allreduce_end(ticket)

    def key(self):
        return (self.rank, self.world)

    @staticmethod
    def chunk(total, world):
        return _ceil_div(total, world)

    def span(self, total, rank=None):
        """(start, length) of this rank's contiguous share of ``total`` items."""
        r = self.rank if rank is None else rank
        ch = self.chunk(total, self.world)
        start = min(r * ch, total)
        return start, min(ch, total - start)

    def ranges(self, C, U):
        (c0, Cl) = self.span(C)
        (u0, Ul) = self.span(U)
        return (c0, Cl, u0, Ul)

    # ------------------------------------------------------------------ collectives
    def allreduce_terms(self, out, idxs):
        """Sums the device vector ``out`` over ranks in place (one collective, no
        staging).  Entries in ``idxs`` are edge-local partial sums; the other
        entries are already complete on every rank, so their sum is divided by
        the world size by ``fix_replicated`` after the download."""
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def any_rank(self, flag):
        """True on every rank if ``flag`` is true on any rank (one tiny all-reduce):
        used where the ranks must agree on a code path whose collectives differ."""
        pw = self.peer_window()
        if pw is not None:
            from . import _dev
            if getattr(self, "_flag_vec", None) is None:
                self._flag_vec = _dev.zeros((1,))
            self._flag_vec.fill_(1.0 if flag else 0.0)
            return bool(pw.allreduce(self._flag_vec, 1, _dev.stream())[0] > 0)
        dev = "cuda" if (torch.cuda.is_available() and dist.get_backend(self.group) == "nccl") else "cpu"
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return bool(t.item() > 0)

    def fix_replicated(self, host_vec, idxs):
        """Undo the summation of entries that were complete on every rank."""
        keep = set(idxs)
        for i in range(len(host_vec)):
            if i not in keep:
                host_vec[i] = host_vec[i] / self.world
        return host_vec

    def edge_buffer_len(self, C):
        """Length of a flat [C*3] edge array padded so that every rank's rows are an equal-sized
        chunk: the all-gather then runs in place on the array itself."""
        return self.world * self.chunk(C, self.world) * 3

    def allgather_edges(self, lqF, qF, C):
        """lqF, qF: flat [C*3] (or padded to ``edge_buffer_len(C)``: gathered in place, no
        staging copies); this rank's edge rows are valid on entry, all rows on exit."""
        ch = self.chunk(C, self.world)
        (start, length) = self.span(C)
        if lqF.numel() >= self.world * ch * 3 and qF.numel() >= self.world * ch * 3 and lqF.is_cuda:
            for t in (lqF, qF):
                full = t[:self.world * ch * 3]
                dist.all_gather_into_tensor(full, full[self.rank * ch * 3:(self.rank + 1) * ch * 3], group=self.group)
            return
        local = lqF.new_zeros((2, ch * 3))
        local[0, :length * 3].copy_(lqF[start * 3:(start + length) * 3])
        local[1, :length * 3].copy_(qF[start * 3:(start + length) * 3])
        gathered = lqF.new_empty((self.world, 2, ch * 3))
        dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        lqF[:C * 3].copy_(gathered[:, 0].reshape(-1)[:C * 3])
        qF[:C * 3].copy_(gathered[:, 1].reshape(-1)[:C * 3])

    def allgather_edge_array(self, t, C):
        """One flat [C*3] edge array padded to ``edge_buffer_len(C)``, gathered in place (this rank's
        rows valid on entry).  Inside ``run()`` only q_F travels every iteration -- the region sweep
        reads all edges of its patients -- while every device-side reader of lq_F (K3a, K4) keeps to
        this rank's rows: lq_F is completed once, when ``run()`` ends."""
        ch = self.chunk(C, self.world)
        if t.numel() < self.world * ch * 3 or not t.is_cuda:
            raise ValueError("allgather_edge_array: the array must be a padded device buffer (edge_buffer_len)")
        full = t[:self.world * ch * 3]
        dist.all_gather_into_tensor(full, full[self.rank * ch * 3:(self.rank + 1) * ch * 3], group=self.group)

    def allgather_patients(self, lqR, qR, N, U):
        """lqR, qR: flat [N*U*2]; this rank's patient columns are valid on entry, all
        columns on exit.  One collective for both arrays; on the GPU the staging is one
        pack and one unpack kernel over reused buffers."""
        ch = self.chunk(U, self.world)
        (u0, Ul) = self.span(U)
        if lqR.is_cuda:
            from . import _dev, _lib
            lib = _lib.load()
            key = (N, ch, lqR.device)
            if getattr(self, "_stage_key", None) != key:
                self._stage = (lqR.new_empty((2 * N * ch * 2,)), lqR.new_empty((self.world * 2 * N * ch * 2,)))
                self._stage_key = key
            (local, gathered) = self._stage
            _lib.check(lib.fcd_pack_patients(_dev.ptr(lqR), _dev.ptr(qR), N, U, u0, Ul, ch, _dev.ptr(local),
                                             _dev.stream()), "fcd_pack_patients")
            dist.all_gather_into_tensor(gathered, local, group=self.group)
            _lib.check(lib.fcd_unpack_patients(_dev.ptr(gathered), self.world, N, U, ch, _dev.ptr(lqR), _dev.ptr(qR),
                                               _dev.stream()), "fcd_unpack_patients")
            return
        local = lqR.new_zeros((2, N, ch, 2))
        local[0, :, :Ul].copy_(lqR.view(N, U, 2)[:, u0:u0 + Ul])
        local[1, :, :Ul].copy_(qR.view(N, U, 2)[:, u0:u0 + Ul])
        gathered = lqR.new_empty((self.world, 2, N, ch, 2))
        dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        for (i, t) in enumerate((lqR, qR)):
            t.view(N, U, 2).copy_(gathered[:, i].permute(1, 0, 2, 3).reshape(N, self.world * ch, 2)[:, :U])

    def exchange_patient_blocks(self, bt_local, Cl, C, U):
        """Edge-sharded -> patient-sharded re-layout of the patient correlations
        (one all-to-all at set-up): rank r holds rows [c0, c0+Cl) of bt for all
        patients and receives every edge of its own patients.  Returns the
        contiguous (C, Ul) block."""
        (u0, Ul) = self.span(U)
        send, in_split, out_split = [], [], []
        for r in range(self.world):
            (ur, Ur) = self.span(U, r)
            send.append(bt_local[:, ur:ur + Ur].reshape(-1))
            in_split.append(Cl * Ur)
            (_, Cr) = self.span(C, r)
            out_split.append(Cr * Ul)
        sendbuf = torch.cat(send) if send else bt_local.new_empty((0,))
        recv = bt_local.new_empty((C * Ul,))
        dist.all_to_all_single(recv, sendbuf, output_split_sizes=out_split, input_split_sizes=in_split,
                               group=self.group)
        return recv.view(C, Ul)


def init_from_env(backend="nccl"):
    """Initialises ``torch.distributed`` from torchrun's environment, binds the
    local GPU and returns an :class:`EdgeShards` (None when WORLD_SIZE is 1)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return EdgeShards()
