"""Context + storage-image targets (replaces sol::Context / sol::Image2d for this path:
src/context.rs:239-369, src/texture.rs:36-96, examples/5-pathtrace.rs:57-80)."""
import ctypes

import numpy as np

from . import _native as N


class Context:
    """One CUDA device + one stream.  `stream` is a raw cudaStream_t (int) — pass
    torch.cuda.current_stream().cuda_stream so torch events time the kernels and torch / NCCL work on that stream is
    ordered with libsolb's.  None (not given) creates a private non-blocking stream.  The handle 0 IS a stream — the
    legacy default stream, which is what torch.cuda.current_stream().cuda_stream returns until the caller switches
    streams — and is adopted as such (passed down as cudaStreamLegacy), not mistaken for "not given"."""

    STREAM_LEGACY = 0x1  # cudaStreamLegacy: explicit handle of the legacy default stream (driver_types.h)

    def __init__(self, device=0, stream=None):
        self._h = ctypes.c_void_p()
        self._lib = N.lib()
        if stream is None:
            handle = None
        else:
            handle = ctypes.c_void_p(int(stream) if int(stream) != 0 else self.STREAM_LEGACY)
        N.check(self._lib.solb_ctx_create(int(device), handle, ctypes.byref(self._h)))
        self.device = int(device)
        self.stream = None if stream is None else int(stream)

    @property
    def handle(self):
        return self._h

    def preload(self):
        """load every kernel + create the scratch pool now (what Pipeline::new's shader compilation is to the reference)"""
        N.check(self._lib.solb_ctx_preload(self._h), self._h)

    def synchronize(self):
        N.check(self._lib.solb_synchronize(self._h), self._h)

    def fence(self):
        """a frame-in-flight fence (AppFrameData::in_flight_fence, src/renderer.rs:8), created signalled"""
        return Fence(self)

    def host_alloc(self, shape, dtype=np.uint8):
        """page-locked host array for uploads / read-backs (solb_host_alloc); freed with the returned array"""
        return _PinnedArray.make(self, shape, dtype)

    def trim(self):
        """hand the cached build scratch back to the driver (solb_ctx_trim)"""
        N.check(self._lib.solb_ctx_trim(self._h), self._h)

    def stats(self):
        st = N.Stats()
        N.check(self._lib.solb_stats_get(self._h, ctypes.byref(st)), self._h)
        return st

    def reset_stats(self):
        N.check(self._lib.solb_stats_reset(self._h), self._h)

    def set_timing(self, enabled):
        N.check(self._lib.solb_set_timing(self._h, int(bool(enabled))), self._h)

    def set_blue_noise(self, rgba8):
        """rgba8: uint8 [h, w, 4], rows already flipped like Texture2d::new (src/texture.rs:490-493)."""
        a = np.ascontiguousarray(rgba8, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[2] == 4
        N.check(self._lib.solb_set_blue_noise(self._h, a.ctypes.data_as(ctypes.c_void_p), a.shape[1], a.shape[0]), self._h)

    def close(self):
        if self._h:
            self._lib.solb_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Fence:
    """solb_fence_*: signal() after a frame's work has been enqueued, wait() before reusing that frame slot's host buffers
    (queue_submit(.., in_flight_fence) / wait_for_and_reset_fence, src/renderer.rs:310-317, 123-131)."""

    def __init__(self, context):
        self.context = context
        self._lib = N.lib()
        self._h = ctypes.c_void_p()
        N.check(self._lib.solb_fence_create(context.handle, ctypes.byref(self._h)), context.handle)

    def signal(self):
        N.check(self._lib.solb_fence_signal(self._h), self.context.handle)

    def wait(self):
        N.check(self._lib.solb_fence_wait(self._h), self.context.handle)

    def close(self):
        if self._h:
            self._lib.solb_fence_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _PinnedArray:
    """owner of one solb_host_alloc block; numpy views keep it alive through .base"""

    def __init__(self, context, nbytes):
        self.context = context
        self._lib = N.lib()
        self.ptr = ctypes.c_void_p()
        N.check(self._lib.solb_host_alloc(context.handle, int(nbytes), ctypes.byref(self.ptr)), context.handle)
        self.nbytes = int(nbytes)
        self.__array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr.value, False), "version": 3}

    @classmethod
    def make(cls, context, shape, dtype):
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        owner = cls(context, max(n, 1))
        return np.asarray(owner)[:n].view(dt).reshape(shape)

    def __del__(self):
        try:
            if self.ptr and self.context.handle:
                self._lib.solb_host_free(self.context.handle, self.ptr)
        except Exception:
            pass


_NP = {N.FORMAT_RGBA32F: (np.float32, 4), N.FORMAT_RGBA8: (np.uint8, 4), N.FORMAT_RG32UI: (np.uint32, 2)}


class Image2d:
    """Device storage image (zero-initialised).  Formats: rgba32f accumulation, rgba8 render, rg32ui ids."""

    def __init__(self, context, width, height, fmt=N.FORMAT_RGBA32F):
        self.context = context
        self._lib = N.lib()
        self._h = ctypes.c_void_p()
        N.check(self._lib.solb_target_create(context.handle, int(width), int(height), int(fmt), ctypes.byref(self._h)), context.handle)
        self.width, self.height, self.format = int(width), int(height), int(fmt)

    @property
    def handle(self):
        return self._h

    def clear(self):
        N.check(self._lib.solb_target_clear(self._h), self.context.handle)

    def readback(self, out=None):
        """Device -> host copy of the whole image (replaces the blit to the present image)."""
        dt, nc = _NP[self.format]
        if out is None:
            out = np.empty((self.height, self.width, nc), dtype=dt)
        assert out.flags.c_contiguous and out.nbytes == self.nbytes
        N.check(self._lib.solb_target_readback(self._h, out.ctypes.data_as(ctypes.c_void_p), out.nbytes), self.context.handle)
        return out

    def readback_async(self, out):
        """the same copy enqueued without the wait (solb_target_readback_async): `out` is complete after a Fence signalled
        later has been waited for, or after Context.synchronize(); give it page-locked memory (Context.host_alloc)"""
        assert out.flags.c_contiguous and out.nbytes == self.nbytes
        N.check(self._lib.solb_target_readback_async(self._h, out.ctypes.data_as(ctypes.c_void_p), out.nbytes), self.context.handle)
        return out

    def upload(self, arr):
        dt, nc = _NP[self.format]
        a = np.ascontiguousarray(arr, dtype=dt)
        assert a.nbytes == self.nbytes
        N.check(self._lib.solb_target_upload(self._h, a.ctypes.data_as(ctypes.c_void_p), a.nbytes), self.context.handle)

    @property
    def nbytes(self):
        dt, nc = _NP[self.format]
        return self.width * self.height * nc * np.dtype(dt).itemsize

    def device_ptr(self):
        p = ctypes.c_void_p()
        N.check(self._lib.solb_target_device_ptr(self._h, ctypes.byref(p)), self.context.handle)
        return p.value

    def as_torch(self):
        """Zero-copy torch view of the device memory (for torch.distributed reduces)."""
        import torch

        dt, nc = _NP[self.format]
        typestr = {np.float32: "<f4", np.uint8: "|u1", np.uint32: "<u4"}[dt]

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (self.height, self.width, nc), "typestr": typestr,
                                      "data": (self.device_ptr(), False), "version": 2}
        h._keepalive = self
        if dt is np.uint32:  # torch has limited uint32 support: view as int32
            h.__cuda_array_interface__["typestr"] = "<i4"
        return torch.as_tensor(h, device="cuda:%d" % self.context.device)

    def close(self):
        if self._h:
            self._lib.solb_target_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
