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
        if use_torch_stream:
            self.use_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def use_stream(self, cuda_stream):
        check(lib().pb200_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        check(lib().pb200_ctx_synchronize(self._h))

    def trim(self):
        """release the temporaries cached between calls (sort buffers, trees) back to the driver"""
        check(lib().pb200_ctx_trim(self._h))

    def set_param(self, key, value):
        check(lib().pb200_ctx_set_param(self._h, key.encode(), int(value)))

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
    """process-wide default context per device (raises PastureB200Error(-100) without a CUDA device)"""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else 0
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def kernel_launch_count():
    return int(lib().pb200_kernel_launch_count())
