"""Device context: one CUDA device + one stream (pb200_ctx).  By default it borrows torch's current
stream for that device so that tensors allocated/filled by torch and the library's kernels are ordered,
and torch.cuda.Event timing brackets the library's launches."""
import ctypes as C

import torch

from ._lib import check, lib

_default = {}


class Context:
    def __init__(self, device=0, use_torch_stream=True):
        h = C.c_void_p()
        check(lib().pb200_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._stream = None
        self._follow_torch = bool(use_torch_stream)
        if use_torch_stream:
            self.bind_current_stream()

    def use_stream(self, cuda_stream):
        check(lib().pb200_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))
        self._stream = cuda_stream

    def bind_current_stream(self):
        """launch on torch's CURRENT stream of this context's device (called by every wrapper before it enters the
        library, so that work issued under `with torch.cuda.stream(s):` stays ordered with torch's own kernels)"""
        if not self._follow_torch:
            return self
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            self.use_stream(s)
        return self

    def synchronize(self):
        check(lib().pb200_ctx_synchronize(self._h))

    def trim(self):
        """release the temporaries cached between calls (sort buffers, trees) back to the driver"""
        check(lib().pb200_ctx_trim(self._h))

    def set_param(self, key, value):
        check(lib().pb200_ctx_set_param(self._h, key.encode(), int(value)))

    def profile(self, on=True):
        """switch the phase timer on/off (CUDA event pairs around the phases of every library call on this context)"""
        self.set_param("profile.phases", 1 if on else 0)

    def profile_read(self):
        """[(phase name, milliseconds)] recorded since the last read, in call order; synchronises"""
        buf = C.create_string_buffer(1 << 16)
        check(lib().pb200_ctx_profile_read(self._h, buf, len(buf)))
        out = []
        self.last_profile_host_us = []
        for line in buf.value.decode().splitlines():
            f = line.split("\t")
            out.append((f[0], float(f[1])))
            self.last_profile_host_us.append((f[0], float(f[2]), float(f[3])))  # host time of scope entry / exit
        return out

    def bind_host_thread(self):
        """bind the calling thread to the CPUs NUMA-local to this device (call before allocating pinned HOST buffers on
        a multi-GPU box); returns {"numa_node": n or -1, "cpus": number of CPUs bound (0 = unchanged)}"""
        node, ncpu = C.c_int(-1), C.c_int(0)
        check(lib().pb200_ctx_bind_host_thread(self._h, C.byref(node), C.byref(ncpu)))
        return {"numa_node": node.value, "cpus": ncpu.value}

    def close(self):
        if self._h:
            lib().pb200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_context(device=None):
    """process-wide default context per device (raises PastureB200Error(-100) without a CUDA device), bound to torch's
    current stream of that device at the time of the call"""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, str):
        device = torch.device(device)
    if isinstance(device, torch.device):
        if device.type != "cuda":
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        else:
            device = device.index if device.index is not None else torch.cuda.current_device()
    if device not in _default:
        _default[device] = Context(device)
    return _default[device].bind_current_stream()


def context_for(ctx, *buffers):
    """the context a wrapper should use: the caller's (re-bound to the current stream), else the default context of the
    device the first device-resident buffer / tensor lives on, else of torch's current device.  A device buffer on another
    device than an explicitly passed context is an error (the library would dereference foreign pointers)."""
    dev = None
    for b in buffers:
        d = getattr(b, "device", None)
        if d is not None and torch.device(d).type == "cuda":
            dev = torch.device(d)
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            break
    if ctx is not None:
        if dev is not None and dev.index != ctx.device:
            raise ValueError(f"buffer lives on {dev} but the context / converter was created for cuda:{ctx.device}")
        return ctx.bind_current_stream()
    return get_context(dev)


def kernel_launch_count():
    return int(lib().pb200_kernel_launch_count())
