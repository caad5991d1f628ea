"""Buffer plumbing.  The product backend is CUDA through PyTorch (device memory, streams); torch
is used for nothing else.  `EmuBackend` exists only for tests/emu (CPU emulation of the kernel
sources) and has to be injected explicitly with the emulated library's path."""
import ctypes as C

import numpy as np

from . import _lib

_NP2TORCH = {}


class CudaBackend:
    name = "cuda"

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("lddecode_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = _lib.load()
        if self.lib.ldd_device_count() < 1:
            raise RuntimeError("libldd_b200.so sees no CUDA device")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.device_index = self.device.index
        global _NP2TORCH
        _NP2TORCH = {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
                     np.dtype("uint8"): torch.uint8, np.dtype("int16"): torch.int16,
                     np.dtype("uint16"): torch.uint16, np.dtype("int32"): torch.int32,
                     np.dtype("int64"): torch.int64, np.dtype("uint32"): torch.uint32}

    def empty(self, n, dtype):
        return self.torch.empty(int(n), dtype=_NP2TORCH[np.dtype(dtype)], device=self.device)

    def zeros(self, n, dtype):
        return self.torch.zeros(int(n), dtype=_NP2TORCH[np.dtype(dtype)], device=self.device)

    def fill_zero(self, buf):
        buf.zero_()

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        t = self.torch.from_numpy(arr)
        return t.to(self.device, non_blocking=False)

    def to_host(self, buf):
        return buf.cpu().numpy()

    def dtype_of(self, np_dtype):
        return _NP2TORCH[np.dtype(np_dtype)]

    def ptr(self, buf):
        return C.c_void_p(buf.data_ptr())

    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    def is_device_buffer(self, x):
        return self.torch.is_tensor(x) and x.is_cuda

    # -- streams / events / pinned staging (whole-capture pipeline)
    def new_stream(self, high_priority=False):
        return self.torch.cuda.Stream(self.device, priority=-1 if high_priority else 0)

    def stream_ctx(self, stream):
        return self.torch.cuda.stream(stream)

    def current_stream_obj(self):
        return self.torch.cuda.current_stream(self.device)

    def wait_stream(self, waiter, other):
        waiter.wait_stream(other)

    def pinned(self, n, dtype):
        return self.torch.empty(int(n), dtype=_NP2TORCH[np.dtype(dtype)]).pin_memory()

    def copy_async(self, dst, src):
        dst.copy_(src, non_blocking=True)

    def record_event(self):
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.current_stream(self.device))
        return ev

    def wait_event(self, ev):
        ev.synchronize()

    def stream_wait_event(self, stream, ev):
        stream.wait_event(ev)

    def host_view(self, pinned_buf):
        return pinned_buf.numpy()


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class EmuBackend:
    """numpy buffers + the g++-compiled emulation library.  Tests only."""
    name = "emu"

    def __init__(self, lib_path):
        self.lib = _lib.load(lib_path)
        self.device_index = 0

    def empty(self, n, dtype):
        return np.empty(int(n), dtype=dtype)

    def zeros(self, n, dtype):
        return np.zeros(int(n), dtype=dtype)

    def fill_zero(self, buf):
        buf[...] = 0

    def to_device(self, arr):
        return np.ascontiguousarray(arr).copy()

    def to_host(self, buf):
        if hasattr(buf, "data_ptr"):
            import torch
            if buf.dtype == torch.uint16:
                return buf.view(torch.int16).numpy().view(np.uint16).copy()
            return buf.numpy().copy()
        return np.array(buf, copy=True)

    def dtype_of(self, np_dtype):
        return np.dtype(np_dtype)

    def ptr(self, buf):
        if hasattr(buf, "data_ptr"):                # a torch CPU tensor (the gloo gather's send buffers)
            return C.c_void_p(buf.data_ptr())
        return C.c_void_p(buf.ctypes.data)

    def stream(self):
        return C.c_void_p(0)

    def synchronize(self):
        pass

    def is_device_buffer(self, x):
        return False

    # -- streams / events: everything is synchronous in the emulation
    def new_stream(self, high_priority=False):
        return None

    def stream_ctx(self, stream):
        return _NullCtx()

    def current_stream_obj(self):
        return None

    def wait_stream(self, waiter, other):
        pass

    def pinned(self, n, dtype):
        return np.empty(int(n), dtype=dtype)

    def copy_async(self, dst, src):
        dst[...] = src

    def record_event(self):
        return None

    def wait_event(self, ev):
        pass

    def stream_wait_event(self, stream, ev):
        pass

    def host_view(self, pinned_buf):
        return pinned_buf
