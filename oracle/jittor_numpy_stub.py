"""A numpy-backed stand-in for the handful of `jittor` calls the reference's HOST-side transform
code makes (TEST INFRASTRUCTURE -- never imported by the product).

Jittor cannot be installed here (no network), but jrender/renderer/transform/{look_at,look,
perspective,orthogonal}.py and jrender/structures/utils/faces_vertices.py are pure tensor algebra.
Executing those files, where they lie under /root/reference, against this stub runs the REFERENCE'S
OWN formulas (operand order, eps values, broadcasting) in float32 numpy, which pins
oracle/preraster.py and the fused CUDA stage at the source level.  What it cannot pin is Jittor's
internal reduction order inside matmul / norm (unspecified) -- hence float tolerances downstream.

jt.normalize follows jittor.misc.normalize: `x / max(||x||_2, eps)` along `dim`.
"""
import sys
import types

import numpy as np


class Var(np.ndarray):
    """ndarray with the jittor.Var methods the reference transform code uses."""

    def float32(self):
        return np.asarray(self, dtype=np.float32).view(Var)

    def unsqueeze(self, dim):
        return np.expand_dims(np.asarray(self), dim).view(Var)

    def broadcast(self, shape):
        return np.broadcast_to(np.asarray(self), tuple(shape)).copy().view(Var)

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        return np.transpose(np.asarray(self), axes).view(Var)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return np.reshape(np.asarray(self), shape).view(Var)

    @property
    def shape(self):  # jittor shapes are lists (look_at.py:17: `[batch_size] + eye.shape`)
        return list(np.asarray(self).shape)

    # -- the few extra Var methods the loss / lighting sources use
    def size(self, dim=None):
        return list(np.asarray(self).shape) if dim is None else np.asarray(self).shape[dim]

    def int(self):
        return np.asarray(self).astype(np.int32).view(Var)

    def numpy(self):
        return np.asarray(self)

    def stop_grad(self):
        return self

    def numel(self):
        return int(np.asarray(self).size)

    def pow(self, e):
        return np.power(np.asarray(self), np.float32(e)).astype(np.asarray(self).dtype).view(Var)

    def sqrt(self):
        return np.sqrt(np.asarray(self)).view(Var)

    def sum(self, dims=None, **kw):   # jittor: x.sum(dims) with a tuple of axes
        if isinstance(dims, list):
            dims = tuple(dims)
        return np.asarray(np.asarray(self).sum(axis=dims, dtype=np.asarray(self).dtype)).view(Var)

    def float64(self):
        return np.asarray(self, dtype=np.float64).view(Var)

    def float(self):
        return self.float32()

    def permute(self, *axes):
        return self.transpose(*axes)

    def sync(self):
        return self

    def clone(self):
        return np.array(np.asarray(self), copy=True).view(Var)

    def sqr(self):
        a = np.asarray(self)
        return (a * a).view(Var)

    def min(self, dim=None, **kw):   # jittor: x.min(0) reduces that axis
        return np.asarray(np.asarray(self).min(axis=dim)).view(Var)

    def max(self, dim=None, **kw):
        return np.asarray(np.asarray(self).max(axis=dim)).view(Var)

    def long(self):
        return np.asarray(self).astype(np.int64).view(Var)

    def view(self, *shape):   # jittor's Var.view == reshape (shadows ndarray.view on purpose: reference code only)
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if len(shape) == 1 and isinstance(shape[0], type):   # ndarray.view(Var) used inside this stub
            return np.ndarray.view(self, shape[0])
        return np.reshape(np.asarray(self), shape).view(Var)

    def reindex_reduce(self, op, shape, indexes, extras):
        """Only the scatter-add pattern structures/mesh.py:241-244 uses: out[e0[i0], i1] += x[i0, i1]."""
        assert op == "sum" and list(indexes) == ["@e0(i0)", "i1"] and len(extras) == 1
        out = np.zeros(tuple(shape), np.asarray(self).dtype)
        np.add.at(out, np.asarray(extras[0]).astype(np.int64), np.asarray(self))
        return out.view(Var)

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(np.asarray(i) if isinstance(i, Var) else i for i in idx)
        elif isinstance(idx, Var):
            idx = np.asarray(idx)
        return np.asarray(np.asarray(self)[idx]).view(Var)


def _v(x):
    return np.asarray(x).view(Var)


def array(x, dtype=None):
    a = np.asarray(x)
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64:
        a = a.astype(np.float32)   # jt.array of Python floats is float32
    elif a.dtype == np.int64:
        a = a.astype(np.int32)
    return a.view(Var)


def normalize(x, p=2, dim=1, eps=1e-12):
    a = np.asarray(x)
    dt = a.dtype if a.dtype == np.float64 else np.float32   # structures/mesh.py:219-221 normalises in float64
    a = a.astype(dt)
    n = np.sqrt((a * a).sum(axis=dim, keepdims=True, dtype=dt)).astype(dt)
    return (a / np.maximum(n, dt.type(eps) if hasattr(dt, "type") else dt(eps))).astype(dt).view(Var)


def cross(a, b, dim=-1):
    a, b = np.asarray(a), np.asarray(b)
    dt = np.float64 if a.dtype == np.float64 or b.dtype == np.float64 else np.float32
    return np.cross(a.astype(dt), b.astype(dt), axis=dim).astype(dt).view(Var)


def matmul(a, b):
    return np.matmul(np.asarray(a, np.float32), np.asarray(b, np.float32)).astype(np.float32).view(Var)


def tan(a):
    return np.tan(np.asarray(a, np.float32)).astype(np.float32).view(Var)


def concat(xs, dim=0):
    return np.concatenate([np.asarray(x) for x in xs], axis=dim).view(Var)


def stack(xs, dim=0):
    return np.stack([np.asarray(x) for x in xs], axis=dim).view(Var)


def install():
    """Put the stub into sys.modules as `jittor`; returns a function that restores the old state."""
    jt = types.ModuleType("jittor")
    jt.array, jt.normalize, jt.cross, jt.matmul, jt.tan, jt.stack = array, normalize, cross, matmul, tan, stack
    jt.abs = lambda a: np.abs(np.asarray(a)).view(Var)
    jt.sum = lambda a, dim=None: np.asarray(np.sum(np.asarray(a), axis=dim)).view(Var)
    jt.contrib = types.SimpleNamespace(concat=concat)
    jt.Var = Var
    jt.sqrt = lambda a: np.sqrt(np.asarray(a)).view(Var)
    jt.cos = lambda a: np.cos(np.asarray(a)).astype(np.asarray(a).dtype).view(Var)
    jt.sin = lambda a: np.sin(np.asarray(a)).astype(np.asarray(a).dtype).view(Var)
    jt.arange = lambda n: np.arange(int(n), dtype=np.int32).view(Var)
    _shape = lambda sh: (sh,) if isinstance(sh, int) else tuple(sh)
    jt.zeros = lambda shape, dtype="float32": np.zeros(_shape(shape), dtype).view(Var)
    jt.ones = lambda shape, dtype="float32": np.ones(_shape(shape), dtype).view(Var)
    jt.empty = lambda shape, dtype="float32": np.zeros(_shape(shape), dtype).view(Var)
    jt.transpose = lambda a, axes: np.transpose(np.asarray(a), axes).view(Var)
    jt.ones_like = lambda a: np.ones_like(np.asarray(a)).view(Var)
    jt.clamp = lambda a, min_v=None, max_v=None: np.clip(np.asarray(a), min_v, max_v).view(Var)
    jt.pow = lambda a, e: np.power(np.asarray(a), np.float32(e)).astype(np.asarray(a).dtype).view(Var)
    nn = types.ModuleType("jittor.nn")

    class Module(object):
        def __call__(self, *a, **kw):
            return self.execute(*a, **kw)
    nn.Module = Module

    def pool(x, kernel_size, op, stride=None):   # nn.pool(images, 2, "mean", stride=2) of dr/softras/rasterizer.py:55
        assert op == "mean" and kernel_size == 2 and stride in (2, None)
        a = np.asarray(x)
        b_, c_, h_, w_ = a.shape
        return a.reshape(b_, c_, h_ // 2, 2, w_ // 2, 2).mean(axis=(3, 5), dtype=a.dtype).view(Var)
    nn.pool = pool

    class Function(object):   # jittor.Function: calling the object runs execute()
        def __call__(self, *a, **kw):
            return self.execute(*a, **kw)
    jt.Function = Function
    nn.relu = lambda a: np.maximum(np.asarray(a), 0).view(Var)
    jt.nn = nn
    saved = {k: sys.modules.get(k) for k in ("jittor", "jittor.nn")}
    sys.modules["jittor"] = jt
    sys.modules["jittor.nn"] = nn

    def restore():
        for k, m in saved.items():
            if m is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = m
    return restore


def load_reference_function(path, name):
    """exec one reference .py (read in place, never copied) and return its function / class `name`."""
    mod = types.ModuleType("ref_host_" + name)
    exec(compile(open(path).read(), path, "exec"), mod.__dict__)
    return getattr(mod, name)
