"""
Device-buffer plumbing: PyTorch owns device memory and streams, nothing else.
Every arithmetic kernel of the hot path lives in ``libfcdiff_b200.so``.
"""
import contextlib
import ctypes

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


def workspace():
    """Per-device reduction workspace (zero-filled once; calls restore it)."""
    dev = device()
    key = dev.index
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
    """Per-device world-1 ``PeerWindow``: reduction results reach the host through
    mapped pinned memory (one tiny kernel + a host-side spin) instead of a
    cudaMemcpyAsync + stream synchronisation."""
    key = device().index
    if key not in _publisher:
        from .dist import PeerWindow
        _publisher[key] = PeerWindow()
    return _publisher[key]


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
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        try:
            yield
        finally:
            e.record()
            self.events.setdefault(name, []).append((s, e))

    def reset(self):
        self.events = {}

    def summary(self):
        """name -> (launches, total_ms, mean_ms)."""
        torch.cuda.synchronize()
        out = {}
        for name, evs in self.events.items():
            ms = [s.elapsed_time(e) for (s, e) in evs]
            out[name] = (len(ms), float(sum(ms)), float(sum(ms) / max(len(ms), 1)))
        return out


_NULL = contextlib.nullcontext()


def timed(profile, name):
    return _NULL if profile is None else profile(name)
