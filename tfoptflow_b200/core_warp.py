"""Drop-in for tfoptflow/core_warp.py: `dense_image_warp(image, flow, name='dense_image_warp')`
(reference: tfoptflow/core_warp.py:153-202), backed by sm_100a CUDA kernels.

To let the pyramid loop's `corr(c1, warp(c2, flow))` (tfoptflow/model_pwcnet.py:1561-1564) run as ONE
kernel without touching the caller, the result is a lazy `WarpedTensor`: `cost_volume` recognises it
and fuses; any other use (a torch function, a method, DLPack export) materialises it first."""
from __future__ import annotations

import torch

from . import ops
from ._tensors import as_cuda_tensor


class WarpedTensor:
    """Deferred dense_image_warp(image, flow * flow_scale).  Behaves like the tensor it stands for."""

    def __init__(self, image, flow, flow_scale: float = 1.0, name: str = "dense_image_warp"):
        self.image, self.flow, self.flow_scale, self.name = image, flow, float(flow_scale), name
        self._value = None

    # -- tensor-like surface -------------------------------------------------------------
    @property
    def shape(self):
        return self.image.shape

    @property
    def dtype(self):
        return self.image.dtype

    @property
    def device(self):
        return self.image.device

    def materialize(self) -> torch.Tensor:
        if self._value is None:
            f = self.flow if self.flow_scale == 1.0 else self.flow * self.flow_scale
            self._value = ops.dense_image_warp_eager(self.image, f)
        return self._value

    def __dlpack__(self, *a, **k):
        return self.materialize().__dlpack__(*a, **k)

    def __dlpack_device__(self):
        return self.materialize().__dlpack_device__()

    def __getattr__(self, item):  # only reached for names not defined above
        return getattr(self.materialize(), item)

    def __repr__(self):
        state = "materialised" if self._value is not None else "lazy"
        return f"WarpedTensor({self.name}, shape={list(self.shape)}, dtype={self.dtype}, {state})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def unwrap(x):
            if isinstance(x, WarpedTensor):
                return x.materialize()
            if isinstance(x, (list, tuple)):
                return type(x)(unwrap(y) for y in x)
            return x
        return func(*unwrap(args), **{k: unwrap(v) for k, v in (kwargs or {}).items()})


def _forward_op(name):
    def op(self, *args, **kwargs):
        return getattr(self.materialize(), name)(*args, **kwargs)
    op.__name__ = name
    return op


for _n in ("add", "radd", "sub", "rsub", "mul", "rmul", "truediv", "rtruediv", "neg", "pow", "matmul", "getitem",
           "eq", "ne", "lt", "le", "gt", "ge", "abs", "len", "iter", "float", "bool"):
    setattr(WarpedTensor, f"__{_n}__", _forward_op(f"__{_n}__"))
WarpedTensor.__hash__ = object.__hash__  # defining __eq__ would otherwise drop it


def dense_image_warp(image, flow, name="dense_image_warp", lazy: bool = True):
    """Image warping using per-pixel flow vectors: out[b,j,i,c] = image[b, j - flow[b,j,i,0], i - flow[b,j,i,1], c]
    (bilinear, edge-clamped; channel 0 of `flow` displaces rows).  Same contract as the reference;
    raises ValueError for wrong ranks / height or width < 2 (tfoptflow/core_warp.py:64-82,173-177)."""
    image, flow = as_cuda_tensor(image, "image"), as_cuda_tensor(flow, "flow")
    if image.dim() != 4:
        raise ValueError("Grid must be 4 dimensional. Received: " + str(list(image.shape)))
    if flow.dim() != 4 or flow.shape[-1] != 2 or tuple(flow.shape[:3]) != tuple(image.shape[:3]):
        raise ValueError("Query points must be [batch, height, width, 2]. Received: " + str(list(flow.shape)))
    if image.shape[1] < 2 or image.shape[2] < 2:
        raise ValueError("dense_image_warp needs height >= 2 and width >= 2")
    w = WarpedTensor(image, flow, 1.0, name)
    return w if lazy else w.materialize()
