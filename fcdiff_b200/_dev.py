"""
Device-buffer plumbing: PyTorch owns device memory and streams, nothing else.
Every arithmetic kernel of the hot path lives in ``libfcdiff_b200.so``.
"""
import contextlib
import ctypes
import os

import numpy as np
import torch

from . import _lib

_ws = {}


_have_cuda = None


def device():
    global _have_cuda
    if _have_cuda is None:
        _have_cuda = bool(torch.cuda.is_available())
    if not _have_cuda:
        raise _lib.FcdError(
            "fcdiff_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    """Current torch stream as a ``cudaStream_t`` for the C-ABI."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_side = {}


def side_stream():
    """Per-device side stream for host -> device uploads that overlap the main stream's kernels."""
    key = device().index
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=device())
    return _side[key]


def ptr(t):
    """Device address of a torch tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def _stream_key():
    """(device, stream): state that kernels of ONE stream use in order -- the reduction workspace, the
    publication window, the solver block -- exists once per stream, so that fits running on different
    streams of a device (replica sweeps, fcdiff_b200/sweep.py) do not share it."""
    dev = device()
    return (dev.index, int(torch.cuda.current_stream().cuda_stream))


def workspace():
    """Per-(device, stream) reduction workspace (zero-filled once; calls restore it)."""
    dev = device()
    key = _stream_key()
    if key not in _ws:
        n = _lib.load().fcd_workspace_bytes() // 8
        _ws[key] = torch.zeros(n, dtype=torch.float64, device=dev)
    return _ws[key]


def empty(shape, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype, device=device())


def upload(a, dtype=np.float64):
    """Host array -> contiguous device tensor.  Pinned host memory is copied
    asynchronously on the current stream."""
    a = np.ascontiguousarray(a, dtype=dtype)
    t = torch.from_numpy(a)
    return t.to(device(), non_blocking=t.is_pinned())


def upload_rows(a, pitch):
    """(R, K) host array -> (R, pitch) device tensor, zero padded columns."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    R, K = a.shape
    if pitch == K:
        return upload(a)
    out = zeros((R, pitch))
    out[:, :K].copy_(torch.from_numpy(a), non_blocking=False)
    return out


def download(t):
    """Device tensor -> host array.  Anything beyond a few KB goes through pinned memory from torch's
    caching host allocator (the link's full rate instead of the pageable path's ~10 GB/s); the array
    returned owns the pinned block."""
    t = t.detach()
    if t.is_cuda and t.numel() * t.element_size() >= (1 << 16):
        t = t.contiguous()
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h.numpy()
    return t.cpu().numpy()


_publisher = {}


def publisher():
    """Per-(device, stream) world-1 ``PeerWindow``: reduction results reach the host through
    mapped pinned memory (one tiny kernel + a host-side spin) instead of a
    cudaMemcpyAsync + stream synchronisation."""
    key = _stream_key()
    if key not in _publisher:
        from .dist import PeerWindow
        _publisher[key] = PeerWindow()
    return _publisher[key]


class SolverBlock(object):
    """Device state + mapped host publication block of the device-resident (eta, epsilon)
    solver (``fcd_solver_*``, csrc/fcd_solver.cuh), one per (device, stream)."""

    TIMEOUT_MS = 30000

    def __init__(self):
        lib = _lib.load()
        self.lib = lib
        self.state = torch.zeros(int(lib.fcd_solver_state_bytes()) // 8, dtype=torch.float64, device=device())
        self.pub = ctypes.c_void_p(0)
        _lib.check(lib.fcd_host_mapped_alloc(int(lib.fcd_solver_published_bytes()), ctypes.byref(self.pub)),
                   "fcd_host_mapped_alloc")
        self.seq = 0
        self.host = _lib.SolverState()

    def wait(self):
        """Blocks until the last launched evaluation has published; returns the state."""
        rc = self.lib.fcd_solver_wait(self.pub, self.seq, ctypes.byref(self.host), self.TIMEOUT_MS)
        if rc != 0:
            _lib.check(rc, "fcd_solver_wait")
        return self.host


_solver = {}


def solver_block():
    key = _stream_key()
    if key not in _solver:
        _solver[key] = SolverBlock()
    return _solver[key]


class SmallResult(object):
    """A few doubles produced by a reduction kernel: one device vector reused for
    every call; reading it costs one publication kernel and a spin on mapped
    host memory (vectors of more than 8 values: one asynchronous copy and one
    stream synchronisation, inside the library)."""

    def __init__(self, n, dtype=torch.float64):
        self.dev = torch.zeros(n, dtype=dtype, device=device())
        self.host = torch.zeros(n, dtype=dtype).pin_memory()
        self.host_np = self.host.numpy()
        self._args = (ctypes.c_void_p(self.host.data_ptr()), ctypes.c_void_p(self.dev.data_ptr()), 8 * n)
        self._download = _lib.load().fcd_download
        self._n = n
        self._pub = publisher() if (dtype == torch.float64 and n <= 8) else None

    def post(self, stream_handle=None):
        """Enqueues the publication of the vector; ``collect`` waits for it (small vectors only).  Work
        enqueued in between runs before the host looks: one wait can serve several results."""
        st = stream() if stream_handle is None else stream_handle
        return self._pub.allreduce_begin(self.dev, self._n, st)

    def collect(self, ticket):
        return self._pub.allreduce_end(ticket)

    def read(self, stream_handle=None):
        st = stream() if stream_handle is None else stream_handle
        if self._pub is not None:
            return self._pub.allreduce(self.dev, self._n, st)
        rc = self._download(self._args[0], self._args[1], self._args[2], st)
        if rc != 0:
            _lib.check(rc, "fcd_download")
        return self.host_np.copy()


def even(n):
    return n + (n & 1)


class KernelTimers(object):
    """CUDA-event brackets around individual C-ABI calls on the current stream
    (``fit.profile = KernelTimers()``): per-kernel launch durations measured
    live, for bench.py's roofline figures."""

    def __init__(self):
        self.events = {}

    @contextlib.contextmanager
    def __call__(self, name):
        """Yields a one-element list: the number of kernel executions the bracket covers (default 1;
        a batch of solver evaluations sets it once it knows how many of its launches did work)."""
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        count = [1]
        s.record()
        try:
            yield count
        finally:
            e.record()
            self.events.setdefault(name, []).append((s, e, count))

    def reset(self):
        self.events = {}

    def relabel_last(self, name, k, new_name):
        """Books the last ``k`` brackets of ``name`` under ``new_name`` (solver launches that found the
        solve finished and exited at once are not evaluations)."""
        if k <= 0 or name not in self.events:
            return
        moved = self.events[name][-k:]
        del self.events[name][-k:]
        self.events.setdefault(new_name, []).extend(moved)

    def summary(self):
        """name -> (launches, total_ms, mean_ms)."""
        torch.cuda.synchronize()
        out = {}
        for name, evs in self.events.items():
            ms = [s.elapsed_time(e) for (s, e, _) in evs]
            n = sum(int(c[0]) for (_, _, c) in evs)
            out[name] = (n, float(sum(ms)), float(sum(ms) / max(n, 1)))
        return out

    def launch_ms(self, name):
        """Durations of the individual brackets of ``name`` in launch order (a mean hides that the first
        iterations after the uniform start carry more records than the settled ones)."""
        torch.cuda.synchronize()
        return [float(s.elapsed_time(e)) for (s, e, _) in self.events.get(name, [])]


_NULL = contextlib.nullcontext()
_NVTX = bool(os.environ.get("FCD_NVTX"))      # FCD_NVTX=1: an NVTX range per kernel family (ncu --nvtx, nsys)


@contextlib.contextmanager
def _nvtx_range(profile, name):
    lib = _lib.load()
    lib.fcd_nvtx_push(name.encode())
    try:
        with (_NULL if profile is None else profile(name)) as count:
            yield count
    finally:
        lib.fcd_nvtx_pop()


def timed(profile, name):
    if _NVTX:
        return _nvtx_range(profile, name)
    return _NULL if profile is None else profile(name)
