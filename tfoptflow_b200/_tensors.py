"""Tensor plumbing: anything that speaks DLPack (PyTorch, TF2 eager, CuPy, JAX) is viewed
zero-copy as a torch CUDA tensor; torch supplies device memory and the current stream only."""
from __future__ import annotations

import torch

from . import _ffi

_DT = {torch.float32: _ffi.PWC_F32, torch.float16: _ffi.PWC_F16, torch.bfloat16: _ffi.PWC_BF16}


def as_cuda_tensor(x, what: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "__dlpack__"):
            x = torch.from_dlpack(x)
        else:
            raise TypeError(f"{what}: expected a torch tensor or an object exposing __dlpack__, got {type(x)}")
    if x.device.type != "cuda":
        raise RuntimeError(f"{what}: tensor must live on a CUDA device (got {x.device}); there is no CPU path")
    if x.dtype not in _DT:
        raise TypeError(f"{what}: unsupported dtype {x.dtype} (float32, float16, bfloat16)")
    return x if x.is_contiguous() else x.contiguous()


def dtype_code(t: torch.Tensor) -> int:
    return _DT[t.dtype]


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
