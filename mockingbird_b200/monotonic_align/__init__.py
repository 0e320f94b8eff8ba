"""Drop-in for the reference's top-level ``monotonic_align`` package (monotonic_align/__init__.py:6-19): ``maximum_path(neg_cent,
mask)`` with the search running on the device (csrc/monotonic.cu) instead of a CPU round trip through Cython."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def maximum_path(neg_cent: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """neg_cent: [b, t_t, t_s]; mask: [b, t_t, t_s] -> path [b, t_t, t_s] (same device / dtype, 0/1)"""
    if neg_cent.device.type != "cuda":
        raise _lib.MbError("mockingbird_b200 maximum_path runs on CUDA only (no CPU fallback)")
    dev, dtype = neg_cent.device, neg_cent.dtype
    values = neg_cent.detach().to(torch.float32).contiguous().clone()  # the reference works on a float32 copy too
    b, t_t, t_s = values.shape
    path = torch.empty(b, t_t, t_s, dtype=torch.int32, device=dev)
    t_t_max = mask.sum(1)[:, 0].to(torch.int32).contiguous()
    t_s_max = mask.sum(2)[:, 0].to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().mb_monotonic_path(C.c_void_p(values.data_ptr()), C.c_void_p(path.data_ptr()),
                                                C.c_void_p(t_t_max.data_ptr()), C.c_void_p(t_s_max.data_ptr()), b, t_t, t_s,
                                                C.c_void_p(stream)))
    return path.to(dtype=dtype)
