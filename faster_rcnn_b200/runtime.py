"""Per-device runtime: one C-ABI handle per (process, device), PyTorch for device memory/streams.

PyTorch is plumbing here (allocator, pinned staging, streams, torch.distributed); every kernel
that runs is one of ours, launched through libfrcnn_b200.so on torch's current stream.
"""
import ctypes as C
import threading

import numpy as np
import torch

from . import _lib

_contexts = {}
_lock = threading.Lock()
_tls = threading.local()          # per-thread context overrides (use_context)
_PIN_RING = 4                     # pinned staging buffers per (dtype, size): reuse waits on that buffer's own copy only


class Context:
    """Owns a frcnn_handle for one CUDA device."""

    def __init__(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError("faster_rcnn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.frcnn_create(C.byref(h), int(device))
        if rc != _lib.OK:
            raise _lib.FrcnnError(rc, "frcnn_create(device=%d) failed (needs compute capability 10.x)" % device)
        self.handle = h
        self._anchor_cache = {}
        self._pinned = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.frcnn_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- plumbing -----------------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.handle, self.stream, *args)
        if rc != _lib.OK:
            _lib.check(self.handle, rc)

    def reserve(self, nbytes):
        _lib.check(self.handle, self.lib.frcnn_reserve(self.handle, int(nbytes)))

    @property
    def launches(self):
        return int(self.lib.frcnn_launch_count(self.handle))

    def anchors(self, anchor_dims):
        """int32 host copy of an (A,2) [height,width] table as a ctypes pointer (cached)."""
        arr = np.ascontiguousarray(np.asarray(anchor_dims), dtype=np.int32)
        if arr.ndim != 2 or arr.shape[1] != 2:
            raise ValueError("anchor_dims must be (A, 2) [height, width]")
        key = arr.tobytes()
        hit = self._anchor_cache.get(key)
        if hit is None:
            hit = (arr, arr.ctypes.data_as(C.c_void_p), arr.shape[0])
            self._anchor_cache[key] = hit
        return hit[1], hit[2]

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # -- host <-> device staging through pinned memory -----------------------------------------
    def to_device(self, array, dtype=None):
        """numpy -> device tensor via a cached pinned staging buffer (async on the current stream)."""
        a = np.ascontiguousarray(array, dtype=dtype)
        t = torch.from_numpy(a)
        key = ("h2d", t.dtype, t.numel())
        ring = self._pinned.get(key)
        if ring is None:
            if len(self._pinned) > 64:
                self._pinned.clear()
            ring = self._pinned[key] = {"next": 0, "slots": []}
        if len(ring["slots"]) < _PIN_RING:
            ring["slots"].append([torch.empty(t.numel(), dtype=t.dtype).pin_memory(), None])
            slot = ring["slots"][-1]
        else:
            slot = ring["slots"][ring["next"]]
            ring["next"] = (ring["next"] + 1) % _PIN_RING
            # only THIS buffer's previous upload (whatever stream issued it) must have drained before it is rewritten;
            # with a ring of buffers that copy is several uploads old, so the wait is normally already satisfied
            if slot[1] is not None:
                slot[1].synchronize()
        pin = slot[0]
        pin.copy_(t.reshape(-1))
        out = pin.to(self.device, non_blocking=True).reshape(t.shape)
        if slot[1] is None:
            slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(self.device))
        return out

    def to_host(self, tensor):
        """device tensor -> numpy (synchronous)."""
        return tensor.cpu().numpy()


def get_context(device=None):
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    override = getattr(_tls, "override", None)
    if override and device in override:
        return override[device]
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = Context(device)
            _contexts[device] = ctx
        return ctx


class use_context:
    """`with use_context(ctx):` makes `get_context(ctx.device)` return `ctx` inside the block, FOR THE CALLING THREAD
    only (other threads keep the shared per-device context).  Used to run (and capture) a sequence of ops on a PRIVATE
    handle whose scratch arena nobody else can grow or move afterwards."""

    def __init__(self, ctx):
        self.ctx, self.prev = ctx, None

    def __enter__(self):
        if not hasattr(_tls, "override"):
            _tls.override = {}
        self.prev = _tls.override.get(self.ctx.device.index)
        _tls.override[self.ctx.device.index] = self.ctx
        return self.ctx

    def __exit__(self, *exc):
        if self.prev is None:
            _tls.override.pop(self.ctx.device.index, None)
        else:
            _tls.override[self.ctx.device.index] = self.prev
        return False


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())
