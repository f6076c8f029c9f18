"""A numpy (and a torch-autograd) stand-in for the ~25 TensorFlow symbols that the reference's hot-path
files import, so that `/root/reference/tfoptflow/core_warp.py` and `core_costvol.py` can be imported and
executed UNMODIFIED in this image (TensorFlow 1.10.1 itself cannot be installed: Python 3.12, no network).

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden_from_reference.py and
tests/test_oracle_vs_reference.py).  Nothing under tfoptflow_b200/ imports it.

What this pins: the reference's own Python source decides every op, its order, its dtype casts, the gather
indices, the padding and the channel order; the stand-in only supplies the element-wise kernels, which are
IEEE arithmetic in the tensor dtype (numpy computes fp16 ops by rounding the exact fp32 result, which is what
Eigen's half kernels do).  One kernel-level assumption is ours: `reduce_mean` of an fp16 tensor accumulates in
fp32 and rounds once (numpy's behaviour and TF's GPU reduction); an all-fp16 accumulation would differ inside
the 1e-2 fp16 tolerance.

Symbols provided (reference call sites):
  tf.pad / shape / unstack / slice / reduce_mean / concat / nn.leaky_relu          core_costvol.py:27-38
  ops.name_scope / convert_to_tensor, constant_op.constant, dtypes.int32           core_warp.py:67-114
  array_ops.shape / unstack / expand_dims / reshape / gather / meshgrid / stack    core_warp.py:70-200
  math_ops.cast / floor / minimum / maximum / range                                core_warp.py:101-124,190
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("PWC_REFERENCE_DIR", "/root/reference/tfoptflow")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "core_warp.py")) and \
        os.path.isfile(os.path.join(REFERENCE_DIR, "core_costvol.py"))


# ------------------------------------------------------------------------------------------------ numpy backend
class _NumpyBackend:
    int32 = np.int32

    @staticmethod
    def tensor(x):
        return np.asarray(x)

    @staticmethod
    def constant(v, dtype=None):
        return np.asarray(v, dtype=dtype)

    @staticmethod
    def shape(x):
        return np.asarray(np.shape(x), dtype=np.int32)

    @staticmethod
    def unstack(x, axis=0):
        x = np.asarray(x)
        if x.ndim == 1 and axis == 0 and x.dtype.kind in "iu":  # a shape vector: hand back python ints
            return [int(v) for v in x]
        return [np.take(x, i, axis=axis) for i in range(x.shape[axis])]

    @staticmethod
    def expand_dims(x, axis):
        return np.expand_dims(x, axis)

    @staticmethod
    def reshape(x, shape):
        return np.reshape(x, [int(s) for s in shape])

    @staticmethod
    def gather(params, indices):
        return params[np.asarray(indices)]

    @staticmethod
    def meshgrid(*xs):
        return np.meshgrid(*xs)  # TF default indexing is 'xy', like numpy

    @staticmethod
    def stack(xs, axis=0):
        return np.stack(xs, axis=axis)

    @staticmethod
    def cast(x, dtype):
        return np.asarray(x).astype(dtype)

    @staticmethod
    def floor(x):
        return np.floor(x)

    @staticmethod
    def minimum(a, b):
        return np.minimum(a, b)

    @staticmethod
    def maximum(a, b):
        return np.maximum(a, b)

    @staticmethod
    def range(n):
        return np.arange(int(n), dtype=np.int32)

    @staticmethod
    def pad(x, paddings):
        return np.pad(x, [[int(a), int(b)] for a, b in paddings])

    @staticmethod
    def slice(x, begin, size):
        idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
        return x[idx]

    @staticmethod
    def reduce_mean(x, axis=None, keepdims=False):
        return np.mean(x, axis=axis, keepdims=keepdims).astype(x.dtype, copy=False)

    @staticmethod
    def concat(xs, axis):
        return np.concatenate(xs, axis=axis)

    @staticmethod
    def leaky_relu(x, alpha=0.2, name=None):  # TF 1.10 nn_ops.leaky_relu: maximum(alpha * features, features)
        return np.maximum(np.asarray(alpha, dtype=x.dtype) * x, x)


# ------------------------------------------------------------------------------------------------ torch backend
class _TorchBackend:
    """Same symbols over torch tensors, so that torch.autograd differentiates the reference's own graph
    (fp64).  torch's maximum/minimum split the gradient at exact ties where TF routes it to one side, so
    backward fixtures made with it use non-integer coordinates and non-zero pre-activations only."""

    def __init__(self):
        import torch
        self.t = torch
        self.int32 = torch.int64  # index arithmetic; the values are the same

    def tensor(self, x):
        return x if isinstance(x, self.t.Tensor) else self.t.as_tensor(x)

    def constant(self, v, dtype=None):
        return self.t.tensor(v, dtype=dtype)

    def shape(self, x):
        return self.t.tensor(list(x.shape), dtype=self.t.int64)

    def unstack(self, x, axis=0):
        if x.dim() == 1 and not x.dtype.is_floating_point:
            return [int(v) for v in x]
        return list(self.t.unbind(x, dim=axis))

    def expand_dims(self, x, axis):
        return x.unsqueeze(axis)

    def reshape(self, x, shape):
        return x.reshape([int(s) for s in shape])

    def gather(self, params, indices):
        return params[indices]

    def meshgrid(self, *xs):
        return self.t.meshgrid(*xs, indexing="xy")

    def stack(self, xs, axis=0):
        return self.t.stack(list(xs), dim=axis)

    def cast(self, x, dtype):
        x = self.tensor(x)
        return x.to(dtype)

    def floor(self, x):
        return self.t.floor(x)

    def minimum(self, a, b):
        return self.t.minimum(self.tensor(a), self.tensor(b))

    def maximum(self, a, b):
        return self.t.maximum(self.tensor(a), self.tensor(b))

    def range(self, n):
        return self.t.arange(int(n), dtype=self.t.int64)

    def pad(self, x, paddings):
        flat = []
        for a, b in reversed([[int(a), int(b)] for a, b in paddings]):
            flat += [a, b]
        return self.t.nn.functional.pad(x, flat)

    def slice(self, x, begin, size):
        idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
        return x[idx]

    def reduce_mean(self, x, axis=None, keepdims=False):
        return x.mean(dim=axis, keepdim=keepdims)

    def concat(self, xs, axis):
        return self.t.cat(list(xs), dim=axis)

    def leaky_relu(self, x, alpha=0.2, name=None):
        return self.t.maximum(alpha * x, x)


# ------------------------------------------------------------------------------------------------ fake modules
def _make_modules(be):
    def mod(name):
        m = types.ModuleType(name)
        m.__dict__["__pwc_shim__"] = True
        return m

    tf = mod("tensorflow")
    tf.pad, tf.shape, tf.unstack, tf.slice = be.pad, be.shape, be.unstack, be.slice
    tf.reduce_mean, tf.concat = be.reduce_mean, be.concat
    tf.nn = mod("tensorflow.nn")
    tf.nn.leaky_relu = be.leaky_relu

    python = mod("tensorflow.python")
    framework = mod("tensorflow.python.framework")
    constant_op = mod("tensorflow.python.framework.constant_op")
    constant_op.constant = be.constant
    dtypes = mod("tensorflow.python.framework.dtypes")
    dtypes.int32 = be.int32
    ops = mod("tensorflow.python.framework.ops")
    ops.name_scope = lambda *a, **k: contextlib.nullcontext()
    ops.convert_to_tensor = be.tensor
    ops_pkg = mod("tensorflow.python.ops")
    array_ops = mod("tensorflow.python.ops.array_ops")
    for n in ("shape", "unstack", "expand_dims", "reshape", "gather", "meshgrid", "stack"):
        setattr(array_ops, n, getattr(be, n))
    math_ops = mod("tensorflow.python.ops.math_ops")
    for n in ("cast", "floor", "minimum", "maximum", "range"):
        setattr(math_ops, n, getattr(be, n))

    tf.python = python
    python.framework, python.ops = framework, ops_pkg
    framework.constant_op, framework.dtypes, framework.ops = constant_op, dtypes, ops
    ops_pkg.array_ops, ops_pkg.math_ops = array_ops, math_ops
    return {m.__name__: m for m in (tf, python, framework, constant_op, dtypes, ops, ops_pkg, array_ops, math_ops)}


def _import_unmodified(filename: str, modname: str):
    path = os.path.join(REFERENCE_DIR, filename)
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)  # executes the reference's source as it lies on disk
    return m


class Reference:
    """`core_warp` / `core_costvol` of the reference, imported unmodified over the chosen backend."""

    def __init__(self, backend: str = "numpy"):
        if not reference_available():
            raise FileNotFoundError(f"reference sources not found under {REFERENCE_DIR}")
        self.be = _NumpyBackend() if backend == "numpy" else _TorchBackend()
        fake = _make_modules(self.be)
        saved = {k: sys.modules.get(k) for k in fake}
        sys.modules.update(fake)
        try:
            self.core_warp = _import_unmodified("core_warp.py", f"_pwc_ref_core_warp_{backend}")
            self.core_costvol = _import_unmodified("core_costvol.py", f"_pwc_ref_core_costvol_{backend}")
        finally:  # the reference modules keep their own bindings; do not leave a fake `tensorflow` behind
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v

    def dense_image_warp(self, image, flow):
        return self.core_warp.dense_image_warp(image, flow)

    def cost_volume(self, c1, warp, search_range):
        return self.core_costvol.cost_volume(c1, warp, search_range, "corr")

    def warp_cost_volume(self, c1, c2, flow, search_range, flow_scale=1.0):
        """The loop body's call pattern, tfoptflow/model_pwcnet.py:1552-1564: `corr(c1, warp(c2, up_flow * scaler))`,
        or `corr(c1, c2)` at the top level."""
        if flow is None:
            return self.cost_volume(c1, c2, search_range)
        if isinstance(flow, np.ndarray):
            scaled = flow * np.asarray(flow_scale).astype(flow.dtype)  # tensor * python float stays in the tensor dtype
        else:
            scaled = flow * flow_scale
        return self.cost_volume(c1, self.dense_image_warp(c2, scaled), search_range)

    def backward_fp64(self, g, c1, c2, flow, search_range, flow_scale=1.0):
        """torch.autograd through the reference's own source in float64 (torch backend only).
        Returns (out, dc1, dc2, dflow|None) as numpy float64."""
        t = self.be.t
        tc1 = t.from_numpy(np.asarray(c1, np.float64)).requires_grad_()
        tc2 = t.from_numpy(np.asarray(c2, np.float64)).requires_grad_()
        tfl = None if flow is None else t.from_numpy(np.asarray(flow, np.float64)).requires_grad_()
        out = self.warp_cost_volume(tc1, tc2, tfl, search_range, flow_scale)
        out.backward(t.from_numpy(np.asarray(g, np.float64)))
        return (out.detach().numpy(), tc1.grad.numpy(), tc2.grad.numpy(), None if tfl is None else tfl.grad.numpy())
