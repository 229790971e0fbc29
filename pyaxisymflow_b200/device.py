"""Device-memory plumbing for the drop-in layer.

Every kernel entry of ``libaxisym_b200`` takes raw device pointers.  The public functions of
this package accept three kinds of arrays:

* :class:`DeviceField` -- a CUDA float64 ``torch.Tensor`` with the small NumPy surface the
  reference's example drivers use (slicing, arithmetic, ``np.fabs``/``np.amax``/``np.sum`` ...),
  so a driver keeps its fields resident on the GPU between kernel calls;
* CUDA ``torch.Tensor`` -- used as is (zero copy);
* ``numpy.ndarray`` -- *parity mode*: staged host->device, the kernel runs on the GPU, and
  outputs are copied back into the caller's array (this is what the tests against the CPU
  oracle use; it is not a CPU fallback -- the arithmetic always happens in the CUDA kernels).

PyTorch is only the allocator / stream provider here.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import AxbGrid


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.AxbError("pyaxisymflow_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def unwrap(x):
    return x.t if isinstance(x, DeviceField) else x


class Stage:
    """Per-call staging area: uploads host arrays once, copies outputs back on finish()."""

    def __init__(self):
        _require_cuda()
        self._seen = {}
        self._host_wb = []
        self._dev_wb = []

    def dev(self, x, out=False, dtype=torch.float64):
        if x is None:
            return None
        if isinstance(x, DeviceField):
            x = x.t
        if isinstance(x, torch.Tensor):
            if not x.is_cuda:
                raise TypeError("torch tensors passed to pyaxisymflow_b200 must live on the GPU")
            if x.dtype == torch.bool and dtype == torch.uint8:
                x = x.view(torch.uint8)
            if x.dtype != dtype:
                raise TypeError(f"expected a {dtype} tensor, got {x.dtype} (the reference binds with noconvert)")
            return x
        if isinstance(x, np.ndarray):
            key = id(x)
            if key in self._seen:
                t = self._seen[key]
            else:
                h = np.ascontiguousarray(x)
                if h.dtype == np.bool_ and dtype == torch.uint8:
                    h = h.view(np.uint8)
                t = torch.from_numpy(h).cuda()
                if t.dtype != dtype:
                    raise TypeError(f"expected a {dtype} array, got {x.dtype} (the reference binds with noconvert)")
                self._seen[key] = t
            if out and all(a is not x for a, _ in self._host_wb):
                self._host_wb.append((x, t))
            return t
        raise TypeError(f"unsupported array type {type(x)!r}")

    def fields(self, outs=(), ins=()):
        """Stage 2-D fields and give them one common row pitch.  Returns (outs, ins, ld)."""
        touts = [self.dev(x, out=True) for x in outs]
        tins = [self.dev(x) for x in ins]
        allt = [t for t in touts + tins if t is not None]
        shape = allt[0].shape
        for t in allt:
            if t.ndim != 2 or t.shape != shape:
                raise ValueError(f"field shapes differ: {tuple(t.shape)} vs {tuple(shape)}")
        lds = {t.stride(0) for t in allt}
        ok = len(lds) == 1 and all(t.stride(1) == 1 for t in allt)
        if not ok:
            def fix(t, is_out):
                if t is None or (t.stride(1) == 1 and t.stride(0) == shape[1]):
                    return t
                c = t.contiguous()
                if is_out:
                    self._dev_wb.append((t, c))
                return c
            # identical input tensors must stay identical after the fix (aliasing semantics)
            cache = {}
            def fix_cached(t, is_out):
                if t is None:
                    return None
                k = (t.data_ptr(), t.stride())
                if k not in cache:
                    cache[k] = fix(t, is_out)
                return cache[k]
            touts = [fix_cached(t, True) for t in touts]
            tins = [fix_cached(t, False) for t in tins]
            ld = shape[1]
        else:
            ld = lds.pop()
        return touts, tins, int(ld)

    def finish(self):
        for orig, tmp in self._dev_wb:
            orig.copy_(tmp)
        for host, t in self._host_wb:
            host[...] = t.cpu().numpy()


def make_grid(nr, nz, ld, dx, slab=None, rows=None, batch=None):
    """axb_grid_t for a full single-GPU field, for a z-slab (kz0, nz_global, ku0, ku1), or -- ``rows=(ju0, ju1)``
    -- for an r-slab block whose rows [ju0, ju1) are the owned ones (the only rows fused reductions count).
    ``batch=(members, field stride in elements, scalar stride in doubles)`` describes an ensemble served by one
    launch per operation (include/axisym_b200.h, axb_grid_t)."""
    ju0, ju1 = (0, 0) if rows is None else rows
    b, bs, ss = (0, 0, 0) if batch is None else batch
    if slab is None:
        return AxbGrid(nr, nz, ld, float(dx), 0, nz, 0, nz, ju0, ju1, b, ss, bs)
    kz0, nzg, ku0, ku1 = slab
    return AxbGrid(nr, nz, ld, float(dx), kz0, nzg, ku0, ku1, ju0, ju1, b, ss, bs)


def coord_1d(stage, X, axis, n):
    """1-D cell-centre coordinates from the reference's meshgrid arrays (R[:, 0] / Z[0, :])."""
    if isinstance(X, DeviceField):
        X = X.t
    if isinstance(X, np.ndarray):
        v = X if X.ndim == 1 else (X[:, 0] if axis == 0 else X[0, :])
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda()
    else:
        v = X if X.ndim == 1 else (X[:, 0] if axis == 0 else X[0, :])
        t = v.contiguous()
    if t.numel() != n:
        raise ValueError(f"coordinate array has {t.numel()} entries, field needs {n}")
    return t


# --------------------------------------------------------------------------------------
# DeviceField: NumPy-flavoured view of a CUDA tensor for the drivers' glue arithmetic
# --------------------------------------------------------------------------------------
_UFUNCS = {
    np.fabs: torch.abs, np.absolute: torch.abs, np.sqrt: torch.sqrt, np.sin: torch.sin,
    np.cos: torch.cos, np.exp: torch.exp, np.add: torch.add, np.subtract: torch.sub,
    np.multiply: torch.mul, np.true_divide: torch.div, np.negative: torch.neg,
    np.power: torch.pow, np.maximum: torch.maximum, np.minimum: torch.minimum,
    np.greater: torch.gt, np.less: torch.lt, np.greater_equal: torch.ge, np.less_equal: torch.le,
}


def _t(x, like=None):
    """operand -> something torch accepts next to a CUDA tensor"""
    if isinstance(x, DeviceField):
        return x.t
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).to(like.device if like is not None else "cuda")
    if isinstance(x, np.generic):
        return x.item()
    return x


class DeviceField:
    """float64 (or bool) field resident on the GPU with the NumPy surface the drivers need."""

    __array_priority__ = 1000

    def __init__(self, t):
        if isinstance(t, DeviceField):
            t = t.t
        if isinstance(t, np.ndarray):
            _require_cuda()
            t = torch.from_numpy(np.ascontiguousarray(t)).cuda()
        self.t = t

    # -- construction helpers ----------------------------------------------------------
    @staticmethod
    def zeros(shape, dtype=torch.float64):
        _require_cuda()
        return DeviceField(torch.zeros(shape, dtype=dtype, device="cuda"))

    @staticmethod
    def meshgrid(z, r):
        """``Z, R = np.meshgrid(z, r)`` on the device."""
        zt, rt = _t(DeviceField(np.asarray(z, dtype=np.float64))), _t(DeviceField(np.asarray(r, dtype=np.float64)))
        R, Z = torch.meshgrid(rt, zt, indexing="ij")
        return DeviceField(Z.contiguous()), DeviceField(R.contiguous())

    # -- ndarray-like attributes ---------------------------------------------------------
    shape = property(lambda s: tuple(s.t.shape))
    ndim = property(lambda s: s.t.ndim)
    size = property(lambda s: s.t.numel())
    dtype = property(lambda s: np.dtype(np.bool_) if s.t.dtype == torch.bool else np.dtype(str(s.t.dtype).split(".")[-1]))

    def get(self):
        return self.t.cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a if dtype is None else a.astype(dtype)

    def copy(self):
        return DeviceField(self.t.clone())

    def reshape(self, *shape):
        return DeviceField(self.t.reshape(*shape))

    def astype(self, dtype):
        return DeviceField(self.t.to(torch.from_numpy(np.zeros(0, dtype=dtype)).dtype))

    def __len__(self):
        return self.t.shape[0]

    def __float__(self):
        return float(self.t)

    def __bool__(self):
        return bool(self.t)

    def __getitem__(self, idx):
        r = self.t[unwrap(idx) if not isinstance(idx, tuple) else tuple(unwrap(i) for i in idx)]
        return DeviceField(r) if r.ndim else r.item()

    def __setitem__(self, idx, val):
        if isinstance(idx, tuple):
            idx = tuple(unwrap(i) for i in idx)
        else:
            idx = unwrap(idx)
        self.t[idx] = _t(val, self.t)

    # -- arithmetic ------------------------------------------------------------------------
    def _bin(self, other, op, rev=False):
        o = _t(other, self.t)
        return DeviceField(op(o, self.t) if rev else op(self.t, o))

    def __add__(s, o): return s._bin(o, torch.add)
    def __radd__(s, o): return s._bin(o, torch.add, True)
    def __sub__(s, o): return s._bin(o, torch.sub)
    def __rsub__(s, o): return s._bin(o, lambda a, b: torch.sub(torch.as_tensor(a, device=b.device), b) if not torch.is_tensor(a) else torch.sub(a, b), True)
    def __mul__(s, o): return s._bin(o, torch.mul)
    def __rmul__(s, o): return s._bin(o, torch.mul, True)
    def __truediv__(s, o): return s._bin(o, torch.div)
    def __rtruediv__(s, o): return s._bin(o, lambda a, b: torch.div(torch.as_tensor(a, device=b.device, dtype=b.dtype), b) if not torch.is_tensor(a) else torch.div(a, b), True)
    def __pow__(s, o): return s._bin(o, torch.pow)
    def __neg__(s): return DeviceField(-s.t)
    def __abs__(s): return DeviceField(s.t.abs())
    def __gt__(s, o): return s._bin(o, torch.gt)
    def __ge__(s, o): return s._bin(o, torch.ge)
    def __lt__(s, o): return s._bin(o, torch.lt)
    def __le__(s, o): return s._bin(o, torch.le)
    def __invert__(s): return DeviceField(~s.t)

    def _ibin(self, other, name):
        getattr(self.t, name)(_t(other, self.t))
        return self

    def __iadd__(s, o): return s._ibin(o, "add_")
    def __isub__(s, o): return s._ibin(o, "sub_")
    def __imul__(s, o): return s._ibin(o, "mul_")
    def __itruediv__(s, o): return s._ibin(o, "div_")

    # -- NumPy protocols -------------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or ufunc not in _UFUNCS or kwargs.get("out") is not None:
            return NotImplemented
        args = [_t(a, self.t) for a in inputs]
        args = [a if torch.is_tensor(a) else torch.as_tensor(a, device=self.t.device) for a in args]
        return DeviceField(_UFUNCS[ufunc](*args))

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.amax, np.max):
            return args[0].t.max().item()
        if func in (np.amin, np.min):
            return args[0].t.min().item()
        if func is np.sum:
            return args[0].t.sum().item()
        if func is np.where:
            if len(args) == 1:
                return tuple(torch.nonzero(unwrap(args[0]), as_tuple=True))
            c, a, b = (_t(x, self.t) for x in args)
            a = a if torch.is_tensor(a) else torch.as_tensor(a, device=self.t.device, dtype=torch.float64)
            b = b if torch.is_tensor(b) else torch.as_tensor(b, device=self.t.device, dtype=torch.float64)
            return DeviceField(torch.where(c, a, b))
        if func is np.flip:
            axis = kwargs.get("axis", args[1] if len(args) > 1 else None)
            dims = list(range(self.t.ndim)) if axis is None else [axis]
            return DeviceField(torch.flip(args[0].t, dims))
        if func is np.zeros_like:
            return DeviceField(torch.zeros_like(args[0].t))
        if func is np.copy:
            return args[0].copy()
        return NotImplemented
