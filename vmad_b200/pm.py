"""Device (B200) implementation of the pmesh object protocol that vmad's FastPM library calls.

This is the host-side mirror of the reference-facing interface: ``ParticleMesh``, ``RealField``,
``ComplexField``, ``Layout`` answer to the same methods, argument names and return conventions
that ``/root/reference/vmad/lib/fastpm.py`` uses on pmesh objects (SURVEY.md Appendix A lists
every call site), so that file -- or ``vmad_b200.lib.fastpm`` -- runs on top of it.  All
arithmetic happens in libvmadpm.so (hand-written sm_100a kernels + cuFFT) through the C ABI in
``include/vmadpm.h``; torch is used only to own device buffers and to name the stream.

There is no CPU path: constructing a ParticleMesh without a CUDA device raises.
"""
import ctypes
import os
import numbers
import weakref

import numpy
import torch

from . import _capi as C

_F8 = torch.float64
_C16 = torch.complex128


# ---------------------------------------------------------------------------------------------
# runtime: one context per process / device
# ---------------------------------------------------------------------------------------------
try:    # the raw handle of torch's current stream without building a Stream object (~10x cheaper;
        # this is asked before nearly every launch)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


class Runtime(object):
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise C.VpmError("vmad_b200 needs a CUDA device (B200, sm_100a); it has no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device)
        self._device_index = self.device.index
        self._stream = torch.cuda.current_stream(self.device).cuda_stream
        h = ctypes.c_void_p()
        C.check(C.lib.vpm_ctx_create(ctypes.byref(h), device, ctypes.c_void_p(self._stream)))
        self.ctx = h

    def push_stream(self, stream):
        """launch the following calls on ``stream`` (a torch stream) without draining the current
        one; only for calls that use neither cuFFT nor the workspace.  ``pop_stream`` restores."""
        C.check(C.lib.vpm_ctx_push_stream(self.ctx, ctypes.c_void_p(stream.cuda_stream)))
        self._pushed = True

    def pop_stream(self):
        C.check(C.lib.vpm_ctx_pop_stream(self.ctx))
        self._pushed = False

    def sync_stream(self):
        """Follow torch's current stream (cheap when unchanged)."""
        if getattr(self, '_pushed', False):
            return
        s = _raw_stream(self._device_index)
        if s != self._stream:
            C.check(C.lib.vpm_ctx_set_stream(self.ctx, ctypes.c_void_p(s)))
            self._stream = s

    def synchronize(self):
        C.check(C.lib.vpm_sync(self.ctx))

    def empty(self, shape, dtype=_F8):
        return torch.empty(shape, dtype=dtype, device=self.device)


_runtime = None


def runtime():
    global _runtime
    if _runtime is None:
        _runtime = Runtime()
    return _runtime


def synchronize():
    runtime().synchronize()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


_KEEP_CELL_INDEX = False
_PAINT_ALGO = 0

# bench.py sets this to a list to time every paint launch of a step in place (CUDA events on
# the launching stream); entries are (start_event, end_event, algorithmic_bytes)
PAINT_EVENTS = None


class _timed_paint(object):
    """context manager around one paint launch; a no-op unless PAINT_EVENTS is a list"""
    __slots__ = ('nbytes', 'e0', 'kernel')

    def __init__(self, nbytes, kernel='k_paint_agg'):
        self.nbytes = nbytes
        self.kernel = kernel
        self.e0 = None

    def __enter__(self):
        if PAINT_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PAINT_EVENTS.append((self.e0, e1, self.nbytes, self.kernel))
        return False


def set_paint_algorithm(name):
    """'aggregated': warp-aggregated fp64 reductions into L2 (needs only the sorted positions;
    summation order not fixed); 'gather': block gather -- no atomics, no zero fill, every cell
    written once, run-to-run bit-identical; 'tile': the shared-memory tile formulation
    (ablation).  'gather' and 'tile' need the per-cell offsets, so layouts keep them while one
    of the two is selected."""
    global _KEEP_CELL_INDEX, _PAINT_ALGO
    algo = {'aggregated': 0, 'tile': 1, 'gather': 2}[name]
    C.check(C.lib.vpm_paint_set_algorithm(algo))
    _KEEP_CELL_INDEX = (algo != 0)
    _PAINT_ALGO = algo


def _ncomp(t):
    """values per row of a [n, ...] tensor (also right for n = 0)"""
    c = 1
    for d in t.shape[1:]:
        c *= int(d)
    return c


def _is_zero_literal(x):
    # vmad's ZeroGradient is an int subclass equal to 0; the literal 0 marks "no gradient"
    return isinstance(x, numbers.Number) and not isinstance(x, bool) and x == 0


# ---------------------------------------------------------------------------------------------
# device arrays
# ---------------------------------------------------------------------------------------------
def _upload(a, dtype=None):
    rt = runtime()
    a = numpy.asarray(a)
    if dtype is None:
        dtype = numpy.complex128 if numpy.iscomplexobj(a) else numpy.float64
    a = numpy.ascontiguousarray(a, dtype=dtype)
    if not a.flags.writeable:
        a = a.copy()
    return torch.from_numpy(a).to(rt.device)


class DeviceArray(object):
    """fp64 (or complex128) array resident in HBM.

    Satisfies what vmad's stdlib needs of a value on the tape (vmad/core/stdlib/operators.py
    :14-152): ``+ - * /`` and negation with another array, a python scalar, or the literal
    ``0``; plus ``len``, ``shape`` and the numpy functions the operator library applies to
    particle arrays (``numpy.stack`` / ``numpy.take`` / ``numpy.sum``, vmad/lib/linalg.py:145-149,
    256-262) via the ``__array_function__`` protocol.  Arrays are immutable by the operator
    contract (the tape aliases them), so ``x + 0`` returns ``x`` itself.
    """
    __array_priority__ = 1000.0

    def __init__(self, t, cellindex=None):
        assert isinstance(t, torch.Tensor) and t.is_cuda
        if not t.is_contiguous():
            t = t.contiguous()
        self.t = t
        # set by Layout.exchange: the rows of this array are in cell-sorted order of `cellindex`
        self._cellindex = cellindex

    # -- construction helpers ---------------------------------------------------------
    @classmethod
    def from_host(kls, a):
        return kls(_upload(a))

    def __vmad_hyper_key__(self):
        """memo key as a hyper-argument of an autooperator (vm/ops.py:_hyper_key): only a
        constant of the graph (``_frozen``: q, uploaded once) may be keyed by identity"""
        if getattr(self, '_frozen', False):
            return ('frozen', id(self))
        raise TypeError("a mutable device array cannot key a memoised model")

    def _like(self, t):
        r = DeviceArray(t)
        return r

    def _coerce(self, other):
        """other -> torch tensor on device with this array's kind, or None if unsupported"""
        if isinstance(other, DeviceArray):
            return other.t
        if isinstance(other, numpy.ndarray) and other.ndim > 0:
            return _upload(other)
        return None

    # -- ndarray-ish surface ------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return numpy.dtype('c16') if self.t.is_complex() else numpy.dtype('f8')

    @property
    def is_complex(self):
        return self.t.is_complex()

    def __len__(self):
        return self.t.shape[0]

    def numpy(self):
        return self.t.cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def copy(self):
        return self._like(self.t.clone())

    @property
    def _ndouble(self):
        return self.t.numel() * (2 if self.t.is_complex() else 1)

    # -- arithmetic ---------------------------------------------------------------------
    def _axpby(self, a, b=0.0, y=None, c=0.0):
        rt = runtime()
        rt.sync_stream()
        out = torch.empty_like(self.t)
        C.check(C.lib.vpm_axpby(rt.ctx, self._ndouble, a, _ptr(self.t), b, _ptr(y), c, _ptr(out)))
        return self._like(out)

    def _binary_same(self, o, fn, conj=0):
        rt = runtime()
        rt.sync_stream()
        out = torch.empty_like(self.t)
        if self.t.is_complex():
            C.check(C.lib.vpm_cmul(rt.ctx, self.t.numel(), _ptr(self.t), _ptr(o), conj, _ptr(out)))
        else:
            C.check(fn(rt.ctx, self.t.numel(), _ptr(self.t), _ptr(o), _ptr(out)))
        return self._like(out)

    def _rows(self, w):
        """self[n, k] * w[n]"""
        rt = runtime()
        rt.sync_stream()
        out = torch.empty_like(self.t)
        n = self.t.shape[0]
        C.check(C.lib.vpm_mul_rows(rt.ctx, n, _ncomp(self.t), _ptr(self.t), _ptr(w), _ptr(out)))
        return self._like(out)

    @staticmethod
    def _scalar(o):
        if isinstance(o, numbers.Real) and not isinstance(o, bool):
            return float(o)
        if isinstance(o, numpy.ndarray) and o.ndim == 0 and o.dtype.kind in 'fiu':
            return float(o)
        return None

    def __add__(self, o):
        s = self._scalar(o)
        if s is not None:
            if s == 0.0:
                return self
            if self.t.is_complex():
                raise TypeError("complex array + non-zero scalar is not supported")
            return self._axpby(1.0, c=s)
        y = self._coerce(o)
        if y is None:
            return NotImplemented
        if y.shape != self.t.shape or y.is_complex() != self.t.is_complex():
            raise ValueError("shape/dtype mismatch in add: %s vs %s" % (self.shape, tuple(y.shape)))
        return self._axpby(1.0, 1.0, y)

    __radd__ = __add__

    def __sub__(self, o):
        s = self._scalar(o)
        if s is not None:
            if s == 0.0:
                return self
            if self.t.is_complex():
                raise TypeError("complex array - non-zero scalar is not supported")
            return self._axpby(1.0, c=-s)
        y = self._coerce(o)
        if y is None:
            return NotImplemented
        if y.shape != self.t.shape or y.is_complex() != self.t.is_complex():
            raise ValueError("shape/dtype mismatch in sub")
        return self._axpby(1.0, -1.0, y)

    def __rsub__(self, o):
        s = self._scalar(o)
        if s is not None:
            if s != 0.0 and self.t.is_complex():
                raise TypeError("non-zero scalar - complex array is not supported")
            return self._axpby(-1.0, c=s)
        y = self._coerce(o)
        if y is None:
            return NotImplemented
        return self._like(y)._axpby(1.0, -1.0, self.t)

    def __neg__(self):
        return self._axpby(-1.0)

    def __pos__(self):
        return self

    def __mul__(self, o):
        s = self._scalar(o)
        if s is not None:
            return self._axpby(s)
        if isinstance(o, numbers.Complex):
            raise TypeError("multiplication by a complex scalar is not supported")
        y = self._coerce(o)
        if y is None:
            return NotImplemented
        if y.shape == self.t.shape and y.is_complex() == self.t.is_complex():
            return self._binary_same(y, C.lib.vpm_mul)
        if (not y.is_complex()) and (not self.t.is_complex()):
            # row scaling [n, k] * [n] or [n, k] * [n, 1]
            if self.t.dim() == 2 and y.numel() == self.t.shape[0]:
                return self._rows(y)
            if y.dim() == 2 and self.t.numel() == y.shape[0]:
                return self._like(y)._rows(self.t)
        raise ValueError("unsupported shapes in mul: %s vs %s" % (self.shape, tuple(y.shape)))

    __rmul__ = __mul__

    def __truediv__(self, o):
        s = self._scalar(o)
        if s is not None:
            with numpy.errstate(divide='ignore'):       # array semantics: x / 0 -> inf / nan
                return self._axpby(float(numpy.float64(1.0) / numpy.float64(s)))
        y = self._coerce(o)
        if y is None:
            return NotImplemented
        if y.shape != self.t.shape or self.t.is_complex():
            raise ValueError("unsupported shapes in div")
        return self._binary_same(y, C.lib.vpm_div)

    def __rtruediv__(self, o):
        s = self._scalar(o)
        if s is None or self.t.is_complex():
            return NotImplemented
        rt = runtime()
        rt.sync_stream()
        num = torch.empty_like(self.t)
        C.check(C.lib.vpm_fill(rt.ctx, num.numel(), s, _ptr(num)))
        return self._like(num)._binary_same(self.t, C.lib.vpm_div)

    def __pow__(self, p):
        if p == 2:
            return self * self
        if p == 1:
            return self
        raise TypeError("only ** 1 and ** 2 are supported on device arrays")

    def conj(self):
        if not self.t.is_complex():
            return self
        rt = runtime()
        rt.sync_stream()
        out = torch.empty_like(self.t)
        # conj(x) = 1 * conj(x): multiply the constant 1 by conj(x)
        one = torch.empty_like(self.t)
        C.check(C.lib.vpm_fill(rt.ctx, one.numel() * 2, 0.0, _ptr(one)))
        one_r = torch.view_as_real(one)
        one_r[..., 0] = 1.0
        C.check(C.lib.vpm_cmul(rt.ctx, self.t.numel(), _ptr(one), _ptr(self.t), 1, _ptr(out)))
        return self._like(out)

    # -- reductions (synchronise: they return host scalars) -------------------------------
    def sum(self, axis=None, dtype=None, out=None, **kw):
        if axis is not None:
            raise TypeError("only full reductions are supported on device arrays")
        if self.t.is_complex():
            raise TypeError("sum of a complex device array is not supported")
        rt = runtime()
        rt.sync_stream()
        r = ctypes.c_double()
        C.check(C.lib.vpm_reduce_sum(rt.ctx, self.t.numel(), _ptr(self.t), ctypes.byref(r)))
        return r.value

    def dot(self, o):
        y = self._coerce(o)
        rt = runtime()
        rt.sync_stream()
        r = ctypes.c_double()
        C.check(C.lib.vpm_reduce_dot(rt.ctx, self._ndouble, _ptr(self.t), _ptr(y), ctypes.byref(r)))
        return r.value

    # -- stack / take ---------------------------------------------------------------------
    def take(self, i, axis=-1):
        if self.t.dim() != 2 or axis not in (-1, 1):
            raise TypeError("take is supported along the last axis of [n, k] arrays")
        rt = runtime()
        rt.sync_stream()
        n, k = self.t.shape
        out = rt.empty((n,))
        C.check(C.lib.vpm_take_col(rt.ctx, n, k, int(i), _ptr(self.t), _ptr(out)))
        return DeviceArray(out, cellindex=self._cellindex)

    # -- numpy protocol glue ----------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs.get('out') is not None:
            return NotImplemented
        if ufunc is numpy.multiply:
            a, b = inputs
            return b.__rmul__(a) if b is self and not isinstance(a, DeviceArray) else a.__mul__(b)
        if ufunc is numpy.add:
            a, b = inputs
            return b.__radd__(a) if b is self and not isinstance(a, DeviceArray) else a.__add__(b)
        if ufunc is numpy.subtract:
            a, b = inputs
            return b.__rsub__(a) if b is self and not isinstance(a, DeviceArray) else a.__sub__(b)
        if ufunc in (numpy.true_divide, numpy.divide):
            a, b = inputs
            return b.__rtruediv__(a) if b is self and not isinstance(a, DeviceArray) else a.__truediv__(b)
        if ufunc is numpy.negative:
            return -inputs[0]
        if ufunc is numpy.conjugate:
            return inputs[0].conj()
        return NotImplemented

    def __array_function__(self, func, types, args, kwargs):
        if func is numpy.stack:
            return stack(args[0], **kwargs) if len(args) == 1 else stack(*args, **kwargs)
        if func is numpy.take:
            a = args[0]
            i = args[1] if len(args) > 1 else kwargs['indices']
            axis = args[2] if len(args) > 2 else kwargs.get('axis', None)
            return a.take(i, axis=axis)
        if func is numpy.sum:
            return args[0].sum(**{k: v for k, v in kwargs.items() if k == 'axis'})
        if func is numpy.shape:
            return args[0].shape
        if func is numpy.conj or func is numpy.conjugate:
            return args[0].conj()
        if func is numpy.copy:
            return args[0].copy()
        return NotImplemented


def stack(arrays, axis=-1):
    """``numpy.stack`` of k ``[n]`` device arrays along the last axis -> ``[n, k]``."""
    arrays = list(arrays)
    if axis not in (-1, 1) or not all(isinstance(a, DeviceArray) and a.t.dim() == 1 for a in arrays):
        raise TypeError("device stack supports 1-D arrays along the last axis")
    rt = runtime()
    rt.sync_stream()
    n = arrays[0].t.shape[0]
    k = len(arrays)
    out = rt.empty((n, k))
    cols = (ctypes.c_void_p * k)(*[a.t.data_ptr() for a in arrays])
    C.check(C.lib.vpm_stack(rt.ctx, n, k, cols, _ptr(out)))
    ci = arrays[0]._cellindex
    if not all(a._cellindex is ci for a in arrays):
        ci = None
    return DeviceArray(out, cellindex=ci)


def asdevice(x):
    """numpy / DeviceArray -> DeviceArray (scalars and the literal 0 pass through)."""
    if isinstance(x, DeviceArray):
        return x
    if isinstance(x, numpy.ndarray) and x.ndim > 0:
        return DeviceArray.from_host(x)
    return x


def ashost(x):
    if isinstance(x, DeviceArray):
        return x.numpy()
    return x


# ---------------------------------------------------------------------------------------------
# communicator shim (fastpm.py:196,251,422,596; lib/mpi.py:24,30,53,57; chisquare.py:32)
# ---------------------------------------------------------------------------------------------
class SelfComm(object):
    rank = 0
    size = 1

    def allreduce(self, x, op=None):
        return x

    def barrier(self):
        pass


# ---------------------------------------------------------------------------------------------
# k-space filters recognised by the fused kernels
# ---------------------------------------------------------------------------------------------
class KList(list):
    """``k`` argument of ``Field.apply`` callbacks; ``normp`` as in pmesh (fastpm.py:141,329)."""

    def normp(self, p=2, zeromode=None):
        r = 0
        for ki in self:
            r = r + numpy.abs(ki) ** p
        r = numpy.array(r, dtype='f8')
        if zeromode is not None:
            r[r == 0] = zeromode
        return r


# ---------------------------------------------------------------------------------------------
# fields
# ---------------------------------------------------------------------------------------------
class Field(DeviceArray):
    def __init__(self, pm, t):
        DeviceArray.__init__(self, t)
        self.pm = pm

    def _like(self, t):
        return type(self)(self.pm, t)

    @property
    def value(self):
        """Host copy of the field (read-only view semantics are not provided)."""
        return self.numpy()

    def __getitem__(self, index):
        return self.numpy()[index]

    def __setitem__(self, index, v):
        if index is not Ellipsis and index != slice(None):
            raise TypeError("device fields only support whole-array assignment")
        self.t.copy_(self.pm._as_field_tensor(v, self.t.is_complex()))

    def apply(self, func, kind='wavenumber', out=None):
        """``Field.apply(func(k, v), kind=, out=)`` (fastpm.py:79,83,87; test_fastpm.py:56).

        ``func`` is evaluated ONCE on the host with numpy coordinate vectors and ``v = 1``
        to obtain the (broadcastable) multiplier table; the multiply itself is one device pass.
        This covers every ``func`` that is linear in ``v`` (all transfer functions are).
        """
        if out is not None and out is not Ellipsis:
            raise TypeError("device Field.apply supports out=None or Ellipsis")
        table = func(self._coords(kind), 1.0)
        r = self._apply_table(table)
        if out is Ellipsis:
            self.t = r.t
            return self
        return r


class RealField(Field):
    def _coords(self, kind):
        pm = self.pm
        xs = []
        for d in range(pm.ndim):
            shape = [1] * pm.ndim
            shape[d] = pm.rshape[d]
            i = numpy.arange(pm.Nmesh[d])
            if pm.P > 1 and d == 0:
                i = i[pm.rstart0:pm.rstart0 + pm.rshape[0]]
            x = i if kind == 'index' else i * (pm.BoxSize[d] / pm.Nmesh[d])
            xs.append(x.reshape(shape))
        return KList(xs)

    def _apply_table(self, table):
        table = numpy.broadcast_to(numpy.asarray(table, dtype='f8'), self.pm.rshape)
        return self * table

    def r2c(self, out=None):
        """fastpm.py:50,56: ``rfftn / Ntot``."""
        return self.pm._r2c(self, 1.0 / self.pm.Ntot)

    def c2r_vjp(self, out=None):
        """fastpm.py:67: ``Ntot * r2c`` (plain forward transform; cotangent convention as in
        the oracle, pinned by test_fastpm.py:101-116)."""
        return self.pm._r2c(self, 1.0)

    def cnorm(self):
        """fastpm.py:15."""
        return self.pm.comm.allreduce(self.dot(self))

    def cdot(self, other):
        """fastpm.py:21."""
        return self.pm.comm.allreduce(self.dot(self.pm._as_field(other, False)))

    def cdot_vjp(self, v):
        return self * v

    def readout(self, pos, layout=None, resampler=None, out=None):
        """fastpm.py:252."""
        return self.pm._readout(self, pos, layout)

    def readout_vjp(self, pos, v, layout=None, resampler=None):
        """fastpm.py:256: returns ``(_mesh, _pos)``."""
        return self.pm._readout_vjp(self, pos, v, layout)

    def readout_jvp(self, pos, v_self=None, v_pos=None, layout=None, resampler=None):
        """fastpm.py:263."""
        return self.pm._readout_jvp(self, pos, v_self, v_pos, layout)


class ComplexField(Field):
    def _coords(self, kind):
        pm = self.pm
        ks = []
        for d in range(pm.ndim):
            shape = [1] * pm.ndim
            shape[d] = pm.cshape[d]
            j = pm.freq_index(d)
            if kind == 'index':
                k = j
            elif kind == 'circular':
                k = j * (2 * numpy.pi / pm.Nmesh[d])
            elif kind == 'wavenumber':
                k = j * (2 * numpy.pi / pm.BoxSize[d])
            else:
                raise ValueError("kind must be wavenumber, circular or index")
            ks.append(k.reshape(shape))
        return KList(ks)

    def _apply_table(self, table, conj=False):
        return self.pm._transfer_table(self, numpy.asarray(table), conj)

    def apply_transfer_function(self, tf, kind='wavenumber', conj=False):
        """``self * tf(k)`` (or ``conj(tf(k))``) with the table of ``tf`` cached on the device"""
        return self.pm._transfer_callable(self, tf, kind, conj)

    def c2r(self, out=None):
        """fastpm.py:64,70: un-normalised inverse."""
        return self.pm._c2r(self, 1.0)

    def r2c_vjp(self, out=None):
        """fastpm.py:53: ``c2r / Ntot``."""
        return self.pm._c2r(self, 1.0 / self.pm.Ntot)

    def _cdot_raw(self, other):
        pm = self.pm
        rt = runtime()
        rt.sync_stream()
        o = pm._as_field(other, True)
        r = (ctypes.c_double * 2)()
        C.check(C.lib.vpm_reduce_cdot(rt.ctx, C.i3(pm._cshape3), int(pm.Nmesh[-1]), _ptr(self.t),
                                      _ptr(o.t), r))
        return complex(r[0], r[1])

    def cnorm(self):
        return self.pm.comm.allreduce(self._cdot_raw(self).real)

    def cdot(self, other):
        """fastpm.py:519,527: sum over the full spectrum of ``x conj(y)``."""
        return self.pm.comm.allreduce(self._cdot_raw(other))

    def cdot_vjp(self, v):
        """fastpm.py:522-523."""
        return self * v

    def _expand_hermitian(self, i, v):
        N = int(self.pm.Nmesh[-1])
        mask = (i[-1] != 0) & (numpy.abs(i[-1]) != N // 2)
        return v * (1 + mask)


# ---------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------
class Layout(object):
    """Result of ``pm.decompose`` (fastpm.py:272): the stable cell-key sort of the particles.

    ``exchange`` delivers rows in cell-sorted order (and tags them so ``paint`` can use the
    sorted kernel without re-sorting); ``gather`` is its adjoint (fastpm.py:281-308).
    """

    def __init__(self, pm, keys, perm, cell_start, oldlength):
        self.pm = pm
        self.keys = keys            # int32 [np]   cell key of every input particle (or None)
        self.indices = perm         # int32 [np]   slot -> input particle (stable)
        self.cell_start = cell_start  # int32 [ncell + 1] (or None: only the tile paint needs it)
        self.oldlength = oldlength
        self.newlength = int(perm.shape[0])
        self._const = None          # (weakref to a constant input array, its sorted rows)

    def exchange(self, data, scale=1.0):
        """fastpm.py:289,301,308.  ``scale``: the rows arrive multiplied by this constant."""
        data = asdevice(data)
        if not isinstance(data, DeviceArray):
            raise TypeError("exchange needs an array")
        frozen = getattr(data, '_frozen', False) and scale == 1.0
        if frozen and self._const is not None and self._const[0]() is data:
            return DeviceArray(self._const[1], cellindex=self)
        rt = runtime()
        rt.sync_stream()
        t = data.t
        n = self.newlength
        ncomp = _ncomp(t)
        out = rt.empty((n,) + tuple(t.shape[1:]))
        C.check(C.lib.vpm_take_rows_scaled(rt.ctx, _ptr(t), _ptr(self.indices), n, ncomp, float(scale), _ptr(out)))
        if frozen:      # a constant of the graph (q): its sorted copy is a constant too
            self._const = (weakref.ref(data), out)
        return DeviceArray(out, cellindex=self)

    def gather(self, data, mode='sum', init=None, scale=1.0):
        """fastpm.py:286,293,304.  ``init`` / ``scale``: returns ``init + scale * gathered``."""
        data = asdevice(data)
        if not isinstance(data, DeviceArray):
            raise TypeError("gather needs an array")
        rt = runtime()
        rt.sync_stream()
        t = data.t
        n = self.newlength
        ncomp = _ncomp(t)
        # single rank: the permutation is a bijection, every output row is written once
        out = rt.empty((self.oldlength,) + tuple(t.shape[1:]))
        C.check(C.lib.vpm_scatter_rows(rt.ctx, _ptr(t), _ptr(self.indices), n, ncomp, _ptr(out)))
        r = DeviceArray(out)
        if init is not None or scale != 1.0:
            r = r * float(scale) if init is None else asdevice(init)._axpby(1.0, float(scale), out)
        return r


class DistLayout(object):
    """``pm.decompose`` over several ranks (slab decomposition along mesh axis 0).

    Routing (done by ``vpm_route_build``, stable, deterministic): a particle goes to the rank
    owning its floor plane and, when different, to the rank owning the next plane -- the two
    planes a CIC particle touches; each rank paints / reads only its own planes and ``gather``
    sums the partial results of the copies, so ``exchange`` and ``gather`` stay mutual adjoints
    (fastpm.py:281-308) and P-rank results equal 1-rank results to rounding.  What arrives is
    cell-sorted locally (keys counted from the lower ghost plane).
    """

    def __init__(self, pm, pos, keep_index=None):
        rt = pm.rt
        rt.sync_stream()
        self.pm = pm
        comm = pm.comm
        P = pm.P
        _, x3 = pm._pos3(pos)
        n = x3.shape[0]
        self.oldlength = n
        send_idx = rt.empty((max(2 * n, 1),), torch.int32)
        start = (ctypes.c_int32 * (P + 1))()
        C.check(C.lib.vpm_route_build(rt.ctx, ctypes.byref(pm._mesh), P, _ptr(x3), n,
                                      _ptr(send_idx), send_idx.shape[0], start))
        self.send_start = [int(start[d]) for d in range(P + 1)]
        self.send_counts = [start[d + 1] - start[d] for d in range(P)]
        self.nsend = int(start[P])
        self.send_idx = send_idx[:self.nsend].clone()  # drop the over-allocation
        # counts[s][d] = rows rank s sends to rank d: every rank sees the whole matrix, so the
        # row offsets inside every peer's receive area are known without a second exchange
        counts = comm.allgather_counts(self.send_counts)
        pm._note_route_counts(counts)
        me = pm.rank
        self.recv_counts = [int(c) for c in counts[:, me]]
        self.nrecv = int(sum(self.recv_counts))
        self.recv_start = [0] + [int(v) for v in numpy.cumsum(self.recv_counts)]
        self._fwd_base = [int(counts[:me, d].sum()) for d in range(P)]     # my rows in d's area
        self._back_base = [int(counts[s, :me].sum()) for s in range(P)]    # my rows in s's send order
        self._max_rows = int(max(counts.sum(axis=0).max(), counts.sum(axis=1).max()))
        self.px = pm._peer()
        # positions travel once here; the local cell sort is built on what arrived
        xr = self._forward(x3)
        local = pm._decompose_local(DeviceArray(xr), keep_index)
        self.indices = local.indices
        self.cell_start = local.cell_start
        self.keys = local.keys
        self.newlength = self.nrecv
        # `exchange(pos)` of the very array this layout was built from (the usual next call,
        # fastpm.py:416-418) is answered from here instead of a second all-to-all
        pos = asdevice(pos)
        self._src = weakref.ref(pos) if isinstance(pos, DeviceArray) and pm.ndim == 3 else None
        if self._src is not None:
            xs = rt.empty(tuple(xr.shape))
            C.check(C.lib.vpm_take_rows(rt.ctx, _ptr(xr), _ptr(self.indices), self.nrecv, 3, _ptr(xs)))
            self._xs = xs       # the bare tensor: a DeviceArray tagged with this layout would form a
                                # reference cycle and keep ~50 MB alive until the cyclic GC runs
        else:
            self._xs = None
        self.take_idx = self.send_remote = None
        if self.px is not None and not getattr(pm, 'force_recompute', False):
            # rows that stay on this rank (nearly all of them) bypass the receive area: composed
            # indices let exchange / gather touch the home array directly (8 B/particle more on
            # the tape per layout -- skipped in the memory-lean mode of the largest runs)
            self.take_idx = rt.empty((max(self.nrecv, 1),), torch.int32)
            self.send_remote = rt.empty((max(self.nsend, 1),), torch.int32)
            nself = self.recv_start[me + 1] - self.recv_start[me]
            self.inv_remote = rt.empty((max(self.nrecv - nself, 1),), torch.int32)
            C.check(C.lib.vpm_route_compose(rt.ctx, _ptr(self.indices), self.nrecv, _ptr(self.send_idx),
                                            self.nsend, self.recv_start[me], self.recv_start[me + 1],
                                            self.send_start[me], _ptr(self.take_idx), _ptr(self.send_remote),
                                            _ptr(self.inv_remote)))

    def _peer_forward(self, t, ncomp, skip=-1):
        """rows -> the peers' receive areas (pack fused with the NVLink stores); returns the
        transfer handle (epoch, buffer) -- the caller consumes my receive area, then releases"""
        px = self.px
        px.ensure(8 * self._max_rows * ncomp)
        e, b = px.begin()
        px.put_rows(b, t, self.send_idx, self.nsend, ncomp, 1, self.send_start, self._fwd_base, skip)
        return e, b

    def _forward(self, t):
        """rows of a local [oldlength, ...] tensor -> rows received here, ordered by source"""
        rt = self.pm.rt
        ncomp = _ncomp(t)
        if self.px is not None:
            e, b = self._peer_forward(t, ncomp)
            out = rt.empty((self.nrecv,) + tuple(t.shape[1:]))
            C.check(C.lib.vpm_axpby(rt.ctx, self.nrecv * ncomp, 1.0, self.px.recv_ptr(b), 0.0, None, 0.0,
                                    _ptr(out)))
            return out
        send = rt.empty((self.nsend,) + tuple(t.shape[1:]))
        C.check(C.lib.vpm_take_rows(rt.ctx, _ptr(t), _ptr(self.send_idx), self.nsend, ncomp, _ptr(send)))
        return self.pm.comm.alltoallv_rows(send, self.send_counts, self.recv_counts)

    def exchange(self, data, scale=1.0):
        """fastpm.py:289,301,308.  ``scale``: the rows arrive multiplied by this constant."""
        data = asdevice(data)
        if scale == 1.0 and self._xs is not None and self._src is not None and self._src() is data:
            return DeviceArray(self._xs, cellindex=self)
        rt = self.pm.rt
        rt.sync_stream()
        scale = float(scale)
        if self.px is not None:
            # put -> (flags) -> the local sort gathers straight out of the receive area
            t = data.t
            ncomp = _ncomp(t)
            out = rt.empty((self.nrecv,) + tuple(t.shape[1:]))
            if self.take_idx is not None:
                e, b = self._peer_forward(t, ncomp, skip=self.pm.rank)
                C.check(C.lib.vpm_take_rows2_scaled(rt.ctx, _ptr(t), self.px.recv_ptr(b), _ptr(self.take_idx),
                                                    self.nrecv, ncomp, scale, _ptr(out)))
            else:
                e, b = self._peer_forward(t, ncomp)
                C.check(C.lib.vpm_take_rows_scaled(rt.ctx, self.px.recv_ptr(b), _ptr(self.indices), self.nrecv,
                                                   ncomp, scale, _ptr(out)))
            return DeviceArray(out, cellindex=self)
        recv = self._forward(data.t)
        ncomp = _ncomp(recv)
        out = rt.empty(tuple(recv.shape))
        C.check(C.lib.vpm_take_rows_scaled(rt.ctx, _ptr(recv), _ptr(self.indices), self.nrecv, ncomp, scale, _ptr(out)))
        return DeviceArray(out, cellindex=self)

    def home_rows(self):
        """slot -> home row on this rank (negative: the row lives elsewhere / holds no particle), or
        None when the layout cannot split a gather into a local store and a remote sum"""
        return self.take_idx if self.px is not None else None

    def gather_remote(self, t, out, scale=1.0):
        """the cross-rank half of ``gather``: the rows of ``t`` (slot order) whose home is on another
        rank travel there and are summed into ``out``; the local rows were stored by the caller
        (vpm_readout3_home / vpm_readout_grad4_home)"""
        rt = self.pm.rt
        ncomp = _ncomp(t)
        px = self.px
        px.ensure(8 * self._max_rows * ncomp)
        _, b = px.begin()
        px.put_rows(b, t, self.inv_remote, self.nrecv, ncomp, 2, self.recv_start, self._back_base,
                    skip=self.pm.rank)
        _scatter_add_remote(rt, px, b, self.send_remote, self.send_start, self.pm.rank, ncomp, out, float(scale))

    def gather(self, data, mode='sum', init=None, scale=1.0):
        """fastpm.py:286,293,304: back to the home rank and order, copies summed.
        ``init`` / ``scale``: returns ``init + scale * gathered`` (the sum starts from a copy of
        ``init`` instead of zeros and the factor rides on the scatter)"""
        data = asdevice(data)
        rt = self.pm.rt
        rt.sync_stream()
        t = data.t
        ncomp = _ncomp(t)
        scale = float(scale)
        shape = (self.oldlength,) + tuple(t.shape[1:])
        out = asdevice(init).t.clone() if init is not None else torch.zeros(shape, dtype=_F8, device=rt.device)
        if self.px is not None:
            # sorted row k sits at receive position indices[k]; it goes back to its source rank,
            # into the row it had in that rank's send order, and is summed into its home row there
            px = self.px
            px.ensure(8 * self._max_rows * ncomp)
            e, b = px.begin()
            if self.take_idx is None:
                px.put_rows(b, t, self.indices, self.nrecv, ncomp, 0, self.recv_start, self._back_base)
                C.check(C.lib.vpm_scatter_add_rows_scaled(rt.ctx, px.recv_ptr(b), _ptr(self.send_idx), self.nsend,
                                                          ncomp, scale, _ptr(out)))
                return DeviceArray(out)
            # only the rows that came from other ranks are visited (through the inverse of the sort)
            px.put_rows(b, t, self.inv_remote, self.nrecv, ncomp, 2, self.recv_start, self._back_base,
                        skip=self.pm.rank)
            # rows that never left go straight into their home rows
            C.check(C.lib.vpm_scatter_add_rows_scaled(rt.ctx, _ptr(t), _ptr(self.take_idx), self.nrecv, ncomp, scale,
                                                      _ptr(out)))
            _scatter_add_remote(rt, px, b, self.send_remote, self.send_start, self.pm.rank, ncomp, out, scale)
            return DeviceArray(out)
        unsorted = rt.empty(tuple(t.shape))
        C.check(C.lib.vpm_scatter_rows(rt.ctx, _ptr(t), _ptr(self.indices), self.nrecv, ncomp, _ptr(unsorted)))
        back = self.pm.comm.alltoallv_rows(unsorted, self.recv_counts, self.send_counts)
        C.check(C.lib.vpm_scatter_add_rows_scaled(rt.ctx, _ptr(back), _ptr(self.send_idx), self.nsend, ncomp, scale,
                                                  _ptr(out)))
        return DeviceArray(out)


import os as _os
_C2R_STRIDED = _os.environ.get('VMAD_B200_C2R_STRIDED', '1') != '0'


def _scatter_add_remote(rt, px, b, send_remote, send_start, me, ncomp, out, scale=1.0):
    """what came back from the other ranks (receive area b, in send order) is summed into the home
    rows; the positions of the self segment hold nothing, so only the ranges around it are visited"""
    for lo, hi in ((0, send_start[me]), (send_start[me + 1], send_start[-1])):
        if hi > lo:
            src = ctypes.c_void_p(px.my_bufs[b] + 8 * ncomp * lo)
            C.check(C.lib.vpm_scatter_add_rows_scaled(rt.ctx, src, _ptr(send_remote[lo:hi]), hi - lo, ncomp,
                                                      float(scale), _ptr(out)))


class RoutingOverflow(C.VpmError):
    """a fixed-capacity segment of the static particle routing was too small"""


class RoutingPlan(object):
    """Fixed capacities ``cap[s][d]`` (rows rank s may send to rank d) of the particle exchange.
    With them every array of a cross-rank layout has a shape the host knows without reading the
    per-peer counts back from the device: send order and receive order are divided into
    segments of fixed capacity (by destination / by source), the rows actually used are counted
    on the device, the rest are inert padding.  That removes the one host synchronisation of
    ``decompose`` and makes a multi-rank step capturable in a CUDA graph."""

    def __init__(self, cap, rank):
        cap = numpy.asarray(cap, dtype='i8')
        P = len(cap)
        self.cap = cap
        self.P = P
        self.sbase = [0] + [int(v) for v in numpy.cumsum(cap[rank, :])]      # my send segments
        self.rbase = [0] + [int(v) for v in numpy.cumsum(cap[:, rank])]      # my receive segments
        self.cap_send, self.cap_recv = self.sbase[-1], self.rbase[-1]
        # exchange: my rows start at row sum_{s < me} cap[s][d] of rank d's receive order
        self.fwd_base = [int(cap[:rank, d].sum()) for d in range(P)]
        # gather: my rows start at row sum_{d < me} cap[s][d] of rank s's send order
        self.back_base = [int(cap[s, :rank].sum()) for s in range(P)]
        self.max_rows = int(max(cap.sum(axis=0).max(), cap.sum(axis=1).max()))


class StaticDistLayout(object):
    """``pm.decompose`` over several ranks with the fixed-capacity routing of a ``RoutingPlan``:
    the same routing rule, per-destination order and results as ``DistLayout`` -- but no value
    ever travels to the host.  Positions without a particle (the unused tail of a segment) are
    stored as a sentinel position that lies on no plane this rank holds: the cell sort files
    them last, paint and readout skip them (no plane of theirs is local), gather masks them."""

    def __init__(self, pm, pos, plan, keep_index=None):
        rt = pm.rt
        rt.sync_stream()
        self.pm = pm
        self.plan = plan
        P, me = pm.P, pm.rank
        _, x3 = pm._pos3(pos)
        n = x3.shape[0]
        self.oldlength = n
        self.px = px = pm._peer()
        self.send_idx = rt.empty((max(plan.cap_send, 1),), torch.int32)
        self.scnt = torch.zeros((P + 1,), dtype=torch.int32, device=rt.device)   # rows per destination + overflow flag
        seg = (ctypes.c_int32 * (P + 1))(*plan.sbase)
        C.check(C.lib.vpm_route_build_static(rt.ctx, ctypes.byref(pm._mesh), P, _ptr(x3), n, seg,
                                             _ptr(self.send_idx), _ptr(self.scnt)))
        pm._overflow |= self.scnt[P:]
        pm._last_static = weakref.ref(self)
        # positions travel once here, with the per-destination counts in the mailboxes
        _, b = px.begin()
        px.put_rows(b, x3, self.send_idx, plan.cap_send, 3, 1, plan.sbase, plan.fwd_base, skip=-1,
                    mail=self.scnt, sentinel=[pm._sentinel(d) for d in range(P)])
        self.rcnt = rt.empty((P,), torch.int32)
        px.read_mail(b, self.rcnt)
        # the local cell sort reads the receive area in place (padding rows carry the sentinel)
        local = pm._decompose_local(None, keep_index, ptr=px.recv_ptr(b), n=plan.cap_recv)
        self.indices = local.indices
        self.cell_start = local.cell_start
        self.keys = local.keys
        self.newlength = plan.cap_recv
        # padding sorted last (its key is the overflow bucket): mark those slots, so that every
        # later take writes the sentinel position / 0 there and every scatter or put skips them
        C.check(C.lib.vpm_route_mark_padding(rt.ctx, _ptr(self.indices), plan.cap_recv, _ptr(self.rcnt), P))
        self._fill = (ctypes.c_double * 3)(*pm._sentinel())
        C.check(C.lib.vpm_ctx_set_fill_row(rt.ctx, self._fill))
        xs = rt.empty((plan.cap_recv, 3))
        C.check(C.lib.vpm_take_rows(rt.ctx, px.recv_ptr(b), _ptr(self.indices), plan.cap_recv, 3, _ptr(xs)))
        pos = asdevice(pos)
        self._src = weakref.ref(pos) if isinstance(pos, DeviceArray) and pm.ndim == 3 else None
        self._xs = xs
        self.take_idx = self.send_remote = None
        if not getattr(pm, 'force_recompute', False):
            # rows that stay on this rank (nearly all) bypass the receive area (see DistLayout)
            self.take_idx = rt.empty((max(plan.cap_recv, 1),), torch.int32)
            self.send_remote = rt.empty((max(plan.cap_send, 1),), torch.int32)
            self.inv_remote = rt.empty((max(plan.cap_recv - int(plan.cap[me, me]), 1),), torch.int32)
            C.check(C.lib.vpm_route_compose(rt.ctx, _ptr(self.indices), plan.cap_recv, _ptr(self.send_idx),
                                            plan.cap_send, plan.rbase[me], plan.rbase[me + 1], plan.sbase[me],
                                            _ptr(self.take_idx), _ptr(self.send_remote), _ptr(self.inv_remote)))

    def _area(self, ncomp):
        if 8 * self.plan.max_rows * ncomp > self.px.nbytes:
            self.px.ensure(8 * self.plan.max_rows * ncomp)      # collective; not during a capture

    def exchange(self, data, scale=1.0):
        """fastpm.py:289,301,308.  ``scale``: the rows arrive multiplied by this constant."""
        data = asdevice(data)
        if scale == 1.0 and self._xs is not None and self._src is not None and self._src() is data:
            return DeviceArray(self._xs, cellindex=self)
        scale = float(scale)
        pm, plan, px = self.pm, self.plan, self.px
        rt = pm.rt
        rt.sync_stream()
        t = data.t
        ncomp = _ncomp(t)
        self._area(ncomp)
        out = rt.empty((plan.cap_recv,) + tuple(t.shape[1:]))
        C.check(C.lib.vpm_ctx_set_fill_row(rt.ctx, self._fill))
        _, b = px.begin()
        if self.take_idx is not None:
            px.put_rows(b, t, self.send_idx, plan.cap_send, ncomp, 1, plan.sbase, plan.fwd_base, skip=pm.rank)
            C.check(C.lib.vpm_take_rows2_scaled(rt.ctx, _ptr(t), px.recv_ptr(b), _ptr(self.take_idx), plan.cap_recv,
                                                ncomp, scale, _ptr(out)))
        else:
            px.put_rows(b, t, self.send_idx, plan.cap_send, ncomp, 1, plan.sbase, plan.fwd_base)
            C.check(C.lib.vpm_take_rows_scaled(rt.ctx, px.recv_ptr(b), _ptr(self.indices), plan.cap_recv, ncomp,
                                               scale, _ptr(out)))
        return DeviceArray(out, cellindex=self)

    def home_rows(self):
        return self.take_idx

    def gather_remote(self, t, out, scale=1.0):
        """see DistLayout.gather_remote"""
        pm, plan, px = self.pm, self.plan, self.px
        ncomp = _ncomp(t)
        self._area(ncomp)
        _, b = px.begin()
        px.put_rows(b, t, self.inv_remote, plan.cap_recv, ncomp, 2, plan.rbase, plan.back_base, skip=pm.rank)
        _scatter_add_remote(pm.rt, px, b, self.send_remote, plan.sbase, pm.rank, ncomp, out, float(scale))

    def gather(self, data, mode='sum', init=None, scale=1.0):
        """fastpm.py:286,293,304: back to the home rank and order, copies summed.
        ``init`` / ``scale``: returns ``init + scale * gathered``"""
        data = asdevice(data)
        pm, plan, px = self.pm, self.plan, self.px
        rt = pm.rt
        rt.sync_stream()
        t = data.t
        ncomp = _ncomp(t)
        scale = float(scale)
        self._area(ncomp)
        shape = (self.oldlength,) + tuple(t.shape[1:])
        out = asdevice(init).t.clone() if init is not None else torch.zeros(shape, dtype=_F8, device=rt.device)
        _, b = px.begin()
        if self.take_idx is None:
            px.put_rows(b, t, self.indices, plan.cap_recv, ncomp, 0, plan.rbase, plan.back_base, seg_count=self.rcnt)
            C.check(C.lib.vpm_scatter_add_rows_scaled(rt.ctx, px.recv_ptr(b), _ptr(self.send_idx), plan.cap_send,
                                                      ncomp, scale, _ptr(out)))
            return DeviceArray(out)
        # only the rows that came from other ranks are visited (through the inverse of the sort;
        # receive positions no sorted slot refers to -- the padding -- carry -1 there)
        px.put_rows(b, t, self.inv_remote, plan.cap_recv, ncomp, 2, plan.rbase, plan.back_base, skip=pm.rank)
        C.check(C.lib.vpm_scatter_add_rows_scaled(rt.ctx, _ptr(t), _ptr(self.take_idx), plan.cap_recv, ncomp, scale,
                                                  _ptr(out)))
        _scatter_add_remote(rt, px, b, self.send_remote, plan.sbase, pm.rank, ncomp, out, scale)
        return DeviceArray(out)


# ---------------------------------------------------------------------------------------------
# particle mesh
# ---------------------------------------------------------------------------------------------
class ParticleMesh(object):
    """``ParticleMesh(Nmesh, BoxSize, comm, dtype='f8', resampler='cic')`` (fastpm.py:542-546)."""

    def __init__(self, Nmesh, BoxSize, comm=None, dtype='f8', resampler='cic', np=None):
        Nmesh = numpy.array(Nmesh, dtype='i8').ravel()
        ndim = len(Nmesh)
        if ndim not in (2, 3):
            raise ValueError("only 2-D and 3-D meshes")
        box = numpy.empty(ndim, dtype='f8')
        box[...] = BoxSize
        if numpy.dtype(dtype) != numpy.dtype('f8'):
            raise ValueError("vmad_b200 computes in fp64 only")
        if resampler != 'cic':
            raise ValueError("only the CIC resampler is implemented")
        self.rt = runtime()
        self.Nmesh = Nmesh
        self.BoxSize = box
        self.ndim = ndim
        self.comm = comm if comm is not None else SelfComm()
        self.dtype = numpy.dtype(dtype)
        self.resampler = resampler
        self.P = int(getattr(self.comm, 'size', 1))
        self.rank = int(getattr(self.comm, 'rank', 0))
        # global shapes
        self.grshape = tuple(int(n) for n in Nmesh)
        self.gcshape = tuple(int(n) for n in Nmesh[:-1]) + (int(Nmesh[-1]) // 2 + 1,)
        # local shapes: real mesh sharded in slabs along axis 0, complex mesh along axis 1
        # ("transposed out", so that r2c and c2r need one all-to-all each)
        if self.P > 1:
            if ndim != 3:
                raise ValueError("multi-rank meshes must be 3-D")
            if Nmesh[0] % self.P or Nmesh[1] % self.P:
                raise ValueError("Nmesh[0] and Nmesh[1] must be divisible by the number of ranks")
            n0l, n1l = int(Nmesh[0]) // self.P, int(Nmesh[1]) // self.P
            self.rstart0, self.cstart1 = self.rank * n0l, self.rank * n1l
            self.rshape = (n0l,) + self.grshape[1:]
            self.cshape = (self.gcshape[0], n1l, self.gcshape[2])
        else:
            self.rstart0 = self.cstart1 = 0
            self.rshape = self.grshape
            self.cshape = self.gcshape
        self.Ntot = int(Nmesh.prod())
        self.scale = self.Nmesh / self.BoxSize  # the one rounded fp64 scale (as in the oracle)

        # 3-D geometry handed to the C ABI (2-D meshes are [1, N1, N2])
        n3 = [1] * (3 - ndim) + [int(n) for n in Nmesh]
        s3 = [1.0] * (3 - ndim) + [float(s) for s in self.scale]
        self._n3 = n3
        self._cshape3 = [1] * (3 - ndim) + list(self.cshape)
        m = C.vpm_mesh()
        for d in range(3):
            m.n[d] = n3[d]
            m.scale[d] = s3[d]
        m.start0 = self.rstart0
        m.nloc0 = self.rshape[0] if ndim == 3 else 1
        m.stride1 = n3[2]
        m.stride0 = n3[1] * n3[2]
        m.ghost = 1 if self.P > 1 else 0
        self._mesh = m
        self._ncell = int(C.lib.vpm_layout_ncell(ctypes.byref(m)))
        self._tables = {}
        self._frozen_layouts = {}     # id(constant position array) -> (weakref, layout)
        self._tf_cache = {}
        self._route_stats = None      # element-wise max of the count matrices seen by decompose
        self._route_plan = None       # RoutingPlan: decompose never talks to the host
        self._overflow = None

    def __vmad_hyper_key__(self):
        """memo key as a hyper-argument of an autooperator: the mesh is immutable configuration;
        the switches a model build reads (lib/fastpm.py:lpt) are part of the key"""
        return ('pm', id(self), getattr(self, 'force_stencil', True), self.P)

    # -- host-side tables ---------------------------------------------------------------
    def freq_index(self, d):
        """integer frequencies of the LOCAL part of complex axis d"""
        N = int(self.Nmesh[d])
        j = numpy.arange(self.gcshape[d], dtype='i8')
        j[j >= (N + 1) // 2] -= N  # even N: Nyquist stored as -N/2; odd N: no Nyquist
        if self.P > 1 and d == 1:
            j = j[self.cstart1:self.cstart1 + self.cshape[1]]
        return j

    def _k_table(self, d3):
        """device wavenumber vector along 3-D axis d3 (zeros for the dummy axis of 2-D)"""
        key = ('k', d3)
        if key not in self._tables:
            d = d3 - (3 - self.ndim)
            if d < 0:
                k = numpy.zeros(1)
            else:
                k = self.freq_index(d) * (2 * numpy.pi / self.BoxSize[d])  # local part
            self._tables[key] = _upload(k)
        return self._tables[key]

    def _axis_table(self, key, d, values):
        """device complex table along mesh axis d (cached under key)"""
        if key not in self._tables:
            v = numpy.ascontiguousarray(values, dtype='c16')
            assert v.shape == (self.cshape[d],)
            self._tables[key] = _upload(v)
        return self._tables[key]

    # -- conversions ----------------------------------------------------------------------
    def _as_field_tensor(self, value, is_complex):
        shape = self.cshape if is_complex else self.rshape
        dtype = _C16 if is_complex else _F8
        if isinstance(value, DeviceArray):
            t = value.t
            if tuple(t.shape) != shape:
                raise ValueError("field shape mismatch: %s vs %s" % (tuple(t.shape), shape))
            return t
        if isinstance(value, numbers.Number):
            return torch.full(shape, value, dtype=dtype, device=self.rt.device)
        a = numpy.asarray(value)
        gshape = self.gcshape if is_complex else self.grshape
        if self.P > 1 and a.shape == gshape:
            # a global host array: keep the local slab
            a = a[:, self.cstart1:self.cstart1 + shape[1]] if is_complex \
                else a[self.rstart0:self.rstart0 + shape[0]]
        a = numpy.broadcast_to(a, shape)
        return _upload(a, numpy.complex128 if is_complex else numpy.float64)

    def _as_field(self, value, is_complex):
        kls = ComplexField if is_complex else RealField
        if isinstance(value, kls) and value.pm is self:
            return value
        return kls(self, self._as_field_tensor(value, is_complex))

    def create(self, type=None, value=None, mode=None):
        """fastpm.py:29,39,52,55,66,69,228,262."""
        type = type if type is not None else mode
        if type not in ('real', 'complex'):
            raise ValueError("type must be 'real' or 'complex'")
        is_complex = type == 'complex'
        if value is None:
            value = 0.0
        if isinstance(value, DeviceArray) and not (isinstance(value, Field) and value.pm is self):
            return (ComplexField if is_complex else RealField)(
                self, self._as_field_tensor(value, is_complex))
        return self._as_field(value, is_complex)

    def generate_uniform_particle_grid(self, shift=0.5, dtype='f8', return_host=False):
        """``(i + shift) * BoxSize / Nmesh``, C order, ``[Np, ndim]`` (fastpm.py:532)."""
        cell = self.BoxSize / self.Nmesh
        axes = [(numpy.arange(n) + shift) * cell[d] for d, n in enumerate(self.grshape)]
        if self.P > 1:
            axes[0] = axes[0][self.rstart0:self.rstart0 + self.rshape[0]]
        if return_host:
            grid = numpy.meshgrid(*axes, indexing='ij')
            return numpy.stack([g.ravel() for g in grid], axis=-1).astype('f8')
        # the per-axis coordinates are the host's numbers (the oracle's); only the broadcast into
        # [Np, ndim] happens on the device (3.2 GB at 512^3: not worth a host round trip)
        shape = tuple(len(a) for a in axes)
        t = torch.empty(shape + (self.ndim,), dtype=_F8, device=self.rt.device)
        for d, a in enumerate(axes):
            view = [1] * self.ndim
            view[d] = len(a)
            t[..., d] = _upload(a).view(view)
        q = DeviceArray(t.view(-1, self.ndim))
        q._frozen = True      # a constant of the graph: decompose() caches its layout
        return q

    def generate_whitenoise(self, seed, unitary=False, type='complex', mean=0.0, mode=None,
                            return_host=False):
        """Seeded white noise with unit variance per complex mode (numpy ``default_rng``; pmesh's
        own generator is not reproducible here).  Generated on the host."""
        type = mode if mode is not None else type
        rng = numpy.random.default_rng(seed)
        g = rng.standard_normal(self.grshape)
        c = numpy.fft.rfftn(g) / self.Ntot ** 0.5
        if unitary:
            a = numpy.abs(c)
            a[a == 0] = 1.0
            c = c / a
        c[(0,) * self.ndim] = mean
        if type == 'real':
            r = numpy.fft.irfftn(c, s=self.grshape) * self.Ntot
            return r if return_host else self.create('real', value=r)
        return c if return_host else self.create('complex', value=c)

    def asparticles(self, x):
        """particle array (host or device) -> device array; done once when a graph is built.
        A copy made here from host data is a constant of the graph (``_frozen``): ``decompose``
        caches the layout of such an array -- the Lagrangian positions ``q`` are decomposed twice
        per LPT evaluation and never change."""
        if isinstance(x, DeviceArray):
            return x
        d = asdevice(numpy.asarray(x) if isinstance(x, (list, tuple)) else x)
        if isinstance(d, DeviceArray):
            d._frozen = True
        return d

    # -- positions --------------------------------------------------------------------------
    def _pos3(self, pos):
        """positions as a contiguous [Np, 3] device tensor (2-D gets a zero first column)"""
        pos = asdevice(pos)
        if not isinstance(pos, DeviceArray):
            raise TypeError("positions must be an array")
        t = pos.t
        if t.dim() != 2 or t.shape[1] != self.ndim:
            raise ValueError("positions must be [Np, %d]" % self.ndim)
        if self.ndim == 2:
            t3 = torch.zeros((t.shape[0], 3), dtype=_F8, device=t.device)
            t3[:, 1:] = t
            return pos, t3
        return pos, t

    def _grad_from3(self, g3):
        if self.ndim == 2:
            return g3[:, 1:].contiguous()
        return g3

    # -- fixed-capacity routing across ranks ------------------------------------------------------
    def _note_route_counts(self, counts):
        counts = numpy.asarray(counts, dtype='i8')
        self._route_stats = counts.copy() if self._route_stats is None else numpy.maximum(self._route_stats, counts)

    def plan_static_routing(self, slack=2.0, floor=4096, cap=None):
        """Switch ``decompose`` to the fixed-capacity routing (``StaticDistLayout``).  The
        capacities come from the per-peer counts observed so far (every decompose of at least one
        representative step in the default mode): rows that stay on their rank are bounded by the
        rank's particle count; for the others ``slack`` x the largest count seen + ``floor``.
        Collective (identical on all ranks: the count matrices were all-gathered).  Returns the
        plan; ``check_routing()`` reports whether a later decompose outgrew it."""
        if self.P == 1:
            return None
        if self._peer() is None:
            raise C.VpmError("the fixed-capacity routing needs the peer-memory exchange")
        n0, n0l = int(self.Nmesh[0]), self.rshape[0]
        if n0 - n0l < 3:
            raise C.VpmError("the fixed-capacity routing needs at least 3 mesh planes outside the local slab")
        if cap is None:
            if self._route_stats is None:
                raise C.VpmError("plan_static_routing: no decompose has been observed yet")
            st = self._route_stats
            cap = numpy.ceil(st * slack).astype('i8') + floor
            for r in range(self.P):
                cap[r, r] = max(int(st[r].sum()), int(st[r, r]))     # <= the rank's own particle count
        plan = RoutingPlan(cap, self.rank)
        self._peer().ensure(24 * plan.max_rows)
        self._route_plan = plan
        self._frozen_layouts.clear()
        if self._overflow is None:
            self._overflow = torch.zeros((1,), dtype=torch.int32, device=self.rt.device)
        return plan

    def check_routing(self):
        """synchronises; raises RoutingOverflow if a decompose since the last check did not fit
        the plan (rows were dropped: results are wrong -- re-plan with more slack and repeat)"""
        if self._overflow is None:
            return
        bad = self.comm.allreduce(int(self._overflow.item()))     # collective: all ranks raise together
        if bad:
            self._overflow.zero_()
            last = getattr(self, '_last_static', lambda: None)()
            info = "" if last is None else (" (rank %d: capacities %s, last layout sent %s)" % (
                self.rank, last.plan.cap[self.rank].tolist(), last.scnt.cpu().tolist()))
            raise RoutingOverflow("a segment of the fixed-capacity particle routing overflowed on %d rank(s)%s"
                                  % (bad, info))

    def _global_count(self, n):
        """sum over ranks of a particle count; the answer for a given local count is cached (the
        home arrays of a simulation never change length), so a step has no host-side collective"""
        hit = self._tables.get(('count', n))
        if hit is None:
            hit = self._tables[('count', n)] = int(self.comm.allreduce(int(n)))
        return hit

    def _sentinel(self, rank=None):
        """a position on a mesh plane that ``rank`` (default: this rank) neither holds nor keys as
        its lower ghost: inert padding for that rank's cell sort, paint and readout"""
        n0, n0l = int(self.Nmesh[0]), self.rshape[0]
        cell = self.BoxSize / self.Nmesh
        rank = self.rank if rank is None else rank
        i0 = (rank * n0l + n0l + 1) % n0
        return ((i0 + 0.5) * cell[0], 0.25 * cell[1], 0.25 * cell[2])

    # -- decompose (fastpm.py:272) ------------------------------------------------------------
    def cell_keys(self, pos):
        rt = self.rt
        rt.sync_stream()
        _, x3 = self._pos3(pos)
        n = x3.shape[0]
        keys = rt.empty((n,), torch.int32)
        C.check(C.lib.vpm_cell_keys(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), n, _ptr(keys)))
        return keys

    def decompose(self, pos, smoothing=None, keep_index=None):
        """fastpm.py:272.  One rank: the stable cell sort.  Several ranks: route every particle
        to the slab(s) that own the two planes it touches, then cell-sort what arrived."""
        frozen = isinstance(pos, DeviceArray) and getattr(pos, '_frozen', False) and keep_index is None
        if frozen:
            hit = self._frozen_layouts.get(id(pos))
            if hit is not None and hit[0]() is pos:
                return hit[1]
        if self.P > 1:
            layout = StaticDistLayout(self, pos, self._route_plan, keep_index) if self._route_plan is not None \
                else DistLayout(self, pos, keep_index)
        else:
            layout = self._decompose_local(pos, keep_index)
        if frozen:
            cache, key = self._frozen_layouts, id(pos)
            cache[key] = (weakref.ref(pos, lambda r, cache=cache, key=key: cache.pop(key, None)), layout)
        return layout

    def _decompose_local(self, pos, keep_index=None, ptr=None, n=None):
        """``keep_index``: also keep the cell keys and cell offsets on the layout (the sorted
        permutation alone is what exchange / gather and the default paint need; the layout sits
        on the tape of every force evaluation, so the extra 8 B/particle + 4 B/cell matter).
        ``ptr, n``: positions given as a raw device pointer to ``[n, 3]`` doubles (a receive area)"""
        rt = self.rt
        rt.sync_stream()
        if ptr is None:
            _, x3 = self._pos3(pos)
            n = x3.shape[0]
            ptr = _ptr(x3)
        keys = rt.empty((n,), torch.int32)
        perm = rt.empty((n,), torch.int32)
        cell_start = rt.empty((self._ncell + 1,), torch.int32)
        C.check(C.lib.vpm_layout_build(rt.ctx, ctypes.byref(self._mesh), ptr, n, _ptr(keys),
                                       _ptr(perm), _ptr(cell_start)))
        if keep_index is None:
            keep_index = _KEEP_CELL_INDEX
        if not keep_index:
            keys = cell_start = None
        return Layout(self, keys, perm, cell_start, n)

    def _sorted(self, pos, layout, extras=()):
        """Return (layout, xs3, extras_sorted): particles in cell-sorted order.

        ``pos`` already tagged by ``Layout.exchange`` of a layout of this mesh is used as is.
        """
        pos = asdevice(pos)
        if layout is None and isinstance(pos, DeviceArray) and pos._cellindex is not None \
                and pos._cellindex.pm is self:
            lay = pos._cellindex
            _, x3 = self._pos3(pos)
            return lay, x3, [asdevice(e) for e in extras], False
        if layout is None:
            # no layout: the positions are taken as local (pmesh semantics); sorting them
            # locally is an optimisation, not an exchange
            layout = self._decompose_local(pos)
        xs = layout.exchange(pos)
        _, x3 = self._pos3(xs)
        ex = []
        for e in extras:
            e = asdevice(e)
            ex.append(layout.exchange(e) if isinstance(e, DeviceArray) else e)
        return layout, x3, ex, True

    # -- paint family (fastpm.py:224,229,238) ---------------------------------------------------
    def paint(self, pos, mass=1.0, layout=None, hold=False, out=None, resampler=None):
        rt = self.rt
        lay, x3, (mass,), _ = self._sorted(pos, layout, (mass,))
        rt.sync_stream()
        marr = mass.t if isinstance(mass, DeviceArray) else None
        m0 = 1.0 if marr is not None else float(mass)
        o = rt.empty(self.rshape)
        cs = self._cell_start_of(lay, x3)
        with _timed_paint((24.0 + (8.0 if marr is not None else 0.0)) * x3.shape[0] + 8.0 * o.numel()):
            C.check(C.lib.vpm_paint_sorted(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), _ptr(marr), m0,
                                           x3.shape[0], _ptr(cs), _ptr(o)))
        return RealField(self, o)

    def paint_cols3(self, xs, mass2d, layout=None, pad=0):
        """three paints of the columns of a sorted [n, 3] array (the mesh half of the vjp of the
        three force readouts, fastpm.py:256 x3) -> three RealFields; one pass over the particles
        with the block-gather algorithm"""
        rt = self.rt
        lay, x3, (mass2d,), _ = self._sorted(xs, layout, (mass2d,))
        rt.sync_stream()
        n = x3.shape[0]
        mt = mass2d.t
        if pad:
            # three slabs with `pad` halo planes on each side of axis 0 (the stencil adjoint reads
            # the halo of the first one); the paint fills the interiors.  Returns the padded array.
            n0l, N1, N2 = self.rshape
            m3 = rt.empty((3, n0l + 2 * pad, N1, N2))
            with _timed_paint(48.0 * n + 24.0 * n0l * N1 * N2, 'k_paint_agg3'):
                C.check(C.lib.vpm_paint_sorted_cols3(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), _ptr(mt), 3, n,
                                                     _ptr(self._cell_start_of(lay, x3)), _ptr(m3[0, pad]),
                                                     m3[0].numel()))
            return m3
        if _PAINT_ALGO in (0, 2):
            m3 = rt.empty((3,) + tuple(self.rshape))
            with _timed_paint(48.0 * n + 8.0 * m3.numel(), 'k_paint_agg3' if _PAINT_ALGO == 0 else 'k_paint_gather<3>'):
                C.check(C.lib.vpm_paint_sorted_cols3(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), _ptr(mt), 3, n,
                                                     _ptr(self._cell_start_of(lay, x3)), _ptr(m3),
                                                     m3[0].numel()))
            return [RealField(self, m3[d]) for d in range(3)]
        out = []
        for d in range(3):
            m = rt.empty(self.rshape)
            with _timed_paint(32.0 * n + 8.0 * m.numel()):
                C.check(C.lib.vpm_paint_sorted_col(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), _ptr(mt), 3, d, n,
                                                   _ptr(self._cell_start_of(lay, x3)), _ptr(m)))
            out.append(RealField(self, m))
        return out

    def _cell_start_of(self, lay, x3):
        """cell offsets for the kernels that need them (tile / row paint); rebuilt from the
        sorted positions when the layout did not keep them"""
        if lay.cell_start is not None:
            return lay.cell_start
        if not _KEEP_CELL_INDEX:
            return self._dummy_cs()  # the aggregated paint ignores it
        lay2 = self._decompose_local(DeviceArray(x3 if self.ndim == 3 else x3[:, 1:].contiguous()),
                                     keep_index=True)
        lay.cell_start = lay2.cell_start
        return lay.cell_start

    def paint_vjp(self, v, pos, layout=None, mass=1.0, resampler=None):
        """returns ``(_pos, _mass)`` (fastpm.py:229)."""
        rt = self.rt
        v = self._as_field(v, False)
        pos = asdevice(pos)
        mass = asdevice(mass)
        if layout is not None:
            xs = layout.exchange(pos)
            ms = layout.exchange(mass) if isinstance(mass, DeviceArray) else mass
        else:
            xs, ms = pos, mass
        _, x3 = self._pos3(xs)
        rt.sync_stream()
        n = x3.shape[0]
        marr = ms.t if isinstance(ms, DeviceArray) else None
        m0 = 1.0 if marr is not None else float(ms)
        val = rt.empty((n,))
        g3 = rt.empty((n, 3))
        C.check(C.lib.vpm_readout_grad(rt.ctx, ctypes.byref(self._mesh), _ptr(v.t), _ptr(x3),
                                       _ptr(marr), m0, n, _ptr(val), _ptr(g3)))
        _pos = DeviceArray(self._grad_from3(g3))
        _mass = DeviceArray(val)
        if layout is not None:
            _pos = layout.gather(_pos)
            _mass = layout.gather(_mass)
        return _pos, _mass

    def paint_jvp(self, pos, v_mass=None, mass=1.0, v_pos=None, layout=None, resampler=None):
        """fastpm.py:238."""
        rt = self.rt
        if _is_zero_literal(v_pos):
            v_pos = None
        if _is_zero_literal(v_mass):
            v_mass = None
        if v_pos is None and v_mass is None:
            return self.create('real')
        if v_pos is None:
            return self.paint(pos, mass=v_mass, layout=layout)
        extras = (mass, v_pos, v_mass)
        lay, x3, (mass, v_pos, v_mass), _ = self._sorted(pos, layout, extras)
        rt.sync_stream()
        _, vp3 = self._pos3(v_pos)
        marr = mass.t if isinstance(mass, DeviceArray) else None
        m0 = 1.0 if marr is not None else float(mass)
        if v_mass is not None and not isinstance(v_mass, DeviceArray):
            vm = torch.full((x3.shape[0],), float(v_mass), dtype=_F8, device=rt.device)
        else:
            vm = v_mass.t if v_mass is not None else None
        o = rt.empty(self.rshape)
        C.check(C.lib.vpm_paint_jvp_sorted(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), _ptr(marr),
                                           m0, _ptr(vp3), _ptr(vm), x3.shape[0],
                                           _ptr(self._cell_start_of(lay, x3)), _ptr(o)))
        return RealField(self, o)

    def _paint_any_order(self, pos, mass=1.0):
        """The aggregated-reduction kernel on particles in ANY order (it merges whatever equal
        or z-adjacent cells happen to sit in neighbouring lanes); benchmark helper."""
        rt = self.rt
        _, x3 = self._pos3(pos)
        rt.sync_stream()
        C.check(C.lib.vpm_paint_set_algorithm(0))
        o = rt.empty(self.rshape)
        C.check(C.lib.vpm_paint_sorted(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), None, float(mass),
                                       x3.shape[0], _ptr(self._dummy_cs()), _ptr(o)))
        return RealField(self, o)

    def _dummy_cs(self):
        if 'dummy_cs' not in self._tables:
            self._tables['dummy_cs'] = torch.zeros((4,), dtype=torch.int32, device=self.rt.device)
        return self._tables['dummy_cs']

    def paint_rows(self, pos, mass=1.0, layout=None):
        """Ablation: the earlier row-gather formulation of the sorted paint."""
        rt = self.rt
        lay, x3, (mass,), _ = self._sorted(pos, layout, (mass,))
        rt.sync_stream()
        marr = mass.t if isinstance(mass, DeviceArray) else None
        m0 = 1.0 if marr is not None else float(mass)
        o = rt.empty(self.rshape)
        if lay.cell_start is None:
            lay = self._decompose_local(DeviceArray(x3), keep_index=True)
        C.check(C.lib.vpm_paint_sorted_rows(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), _ptr(marr),
                                            m0, x3.shape[0], _ptr(lay.cell_start), _ptr(o)))
        return RealField(self, o)

    def paint_atomic(self, pos, mass=1.0):
        """Baseline scatter with global fp64 atomics (for the atomic-throughput comparison)."""
        rt = self.rt
        _, x3 = self._pos3(pos)
        mass = asdevice(mass)
        rt.sync_stream()
        marr = mass.t if isinstance(mass, DeviceArray) else None
        m0 = 1.0 if marr is not None else float(mass)
        o = rt.empty(self.rshape)
        C.check(C.lib.vpm_paint_atomic(rt.ctx, ctypes.byref(self._mesh), _ptr(x3), _ptr(marr), m0,
                                       x3.shape[0], _ptr(o)))
        return RealField(self, o)

    # -- readout family (fastpm.py:252,256,263) ----------------------------------------------------
    def _readout(self, field, pos, layout):
        rt = self.rt
        pos = asdevice(pos)
        if layout is not None and self.P == 1:
            # one rank: the layout only re-orders.  A gather needs no sorted input (a scatter
            # does), and exchange + gather would cost four more passes over the particles
            layout = None
        xs = layout.exchange(pos) if layout is not None else pos
        _, x3 = self._pos3(xs)
        rt.sync_stream()
        n = x3.shape[0]
        val = rt.empty((n,))
        C.check(C.lib.vpm_readout(rt.ctx, ctypes.byref(self._mesh), _ptr(field.t), _ptr(x3), n,
                                  _ptr(val)))
        r = DeviceArray(val, cellindex=xs._cellindex if layout is None else None)
        if layout is not None:
            r = layout.gather(r)
        return r

    def readout3(self, f0, f1, f2, pos, gather=None, kick=None, want_value=True):
        """Three fields read at once into ``[Np, 3]`` (readout x3 + linalg.stack fused).
        ``gather``: a one-rank Layout whose ``gather`` is fused into the store (rows land in
        home order); a cross-rank layout gathers afterwards.  ``kick=(y, fac)``: also return
        ``y + fac * value`` (in home order), computed in the same store when the gather is fused;
        ``want_value=False`` then skips the value itself.  Returns ``value`` or
        ``(value or None, y + fac * value)``."""
        rt = self.rt
        pos = asdevice(pos)
        _, x3 = self._pos3(pos)
        rt.sync_stream()
        n = x3.shape[0]
        fused = isinstance(gather, Layout) and gather.newlength == n == gather.oldlength
        if kick is not None and fused:
            y = asdevice(kick[0])
            val = rt.empty((n, 3)) if want_value else None
            yo = rt.empty((n, 3))
            C.check(C.lib.vpm_readout3(rt.ctx, ctypes.byref(self._mesh), _ptr(f0.t), _ptr(f1.t), _ptr(f2.t),
                                       _ptr(x3), n, _ptr(gather.indices), _ptr(val), _ptr(y.t), float(kick[1]),
                                       _ptr(yo)))
            return (DeviceArray(val) if want_value else None), DeviceArray(yo)
        home = gather.home_rows() if (gather is not None and not fused and hasattr(gather, 'home_rows')) else None
        if home is not None and (kick is None or not want_value):
            # several ranks: rows whose home is here are stored (kick included) by the readout itself,
            # the partial values of the others travel from the slot-order copy and are summed after
            val = rt.empty((n, 3))
            init = asdevice(kick[0]) if kick is not None else None
            scale = float(kick[1]) if kick is not None else 1.0
            out = init.t.clone() if init is not None else torch.zeros((gather.oldlength, 3), dtype=_F8, device=rt.device)
            C.check(C.lib.vpm_readout3_home(rt.ctx, ctypes.byref(self._mesh), _ptr(f0.t), _ptr(f1.t), _ptr(f2.t),
                                            _ptr(x3), n, _ptr(home), _ptr(val),
                                            _ptr(init.t) if init is not None else None, scale, _ptr(out)))
            gather.gather_remote(val, out, scale)
            return (None, DeviceArray(out)) if kick is not None else DeviceArray(out)
        val = rt.empty((n, 3))
        C.check(C.lib.vpm_readout3(rt.ctx, ctypes.byref(self._mesh), _ptr(f0.t), _ptr(f1.t),
                                   _ptr(f2.t), _ptr(x3), n, _ptr(gather.indices) if fused else None,
                                   _ptr(val), None, 0.0, None))
        if fused:
            r = DeviceArray(val)
        else:
            r = DeviceArray(val, cellindex=pos._cellindex)
            if gather is not None:
                if kick is not None and not want_value:
                    # the kick rides on the gather: the sum starts from a copy of y, factor on the scatter
                    return None, gather.gather(r, init=kick[0], scale=kick[1])
                r = gather.gather(r)
        if kick is not None:
            return r, asdevice(kick[0])._axpby(1.0, float(kick[1]), r.t)
        return r

    def _readout_vjp(self, field, pos, v, layout):
        rt = self.rt
        pos = asdevice(pos)
        v = asdevice(v)
        if layout is not None:
            xs = layout.exchange(pos)
            vs = layout.exchange(v)
        else:
            xs, vs = pos, v
        # mesh half: paint with mass = v (sorted kernel; sorts here if xs is not tagged)
        _mesh = self.paint(xs, mass=vs, layout=None)
        _, x3 = self._pos3(xs)
        rt.sync_stream()
        n = x3.shape[0]
        g3 = rt.empty((n, 3))
        C.check(C.lib.vpm_readout_grad(rt.ctx, ctypes.byref(self._mesh), _ptr(field.t), _ptr(x3),
                                       _ptr(vs.t), 1.0, n, None, _ptr(g3)))
        _pos = DeviceArray(self._grad_from3(g3))
        if layout is not None:
            _pos = layout.gather(_pos)
        return _mesh, _pos

    def _readout_jvp(self, field, pos, v_self, v_pos, layout):
        rt = self.rt
        pos = asdevice(pos)
        if _is_zero_literal(v_self):
            v_self = None
        if _is_zero_literal(v_pos):
            v_pos = None
        if v_self is None and v_pos is None:
            return 0
        if layout is not None:
            xs = layout.exchange(pos)
            v_pos = layout.exchange(v_pos) if v_pos is not None else None
        else:
            xs = pos
        _, x3 = self._pos3(xs)
        fdot = self._as_field(v_self, False).t if v_self is not None else None
        xdot = self._pos3(v_pos)[1] if v_pos is not None else None
        rt.sync_stream()
        n = x3.shape[0]
        val = rt.empty((n,))
        C.check(C.lib.vpm_readout_jvp(rt.ctx, ctypes.byref(self._mesh), _ptr(field.t), _ptr(fdot),
                                      _ptr(x3), _ptr(xdot), n, _ptr(val)))
        r = DeviceArray(val)
        if layout is not None:
            r = layout.gather(r)
        return r

    # -- FFT (fastpm.py:50,53,56,64,67,70) ------------------------------------------------------------
    def _r2c(self, real, scale):
        rt = self.rt
        rt.sync_stream()
        if self.P > 1:
            return self._r2c_slab(real, scale)
        o = rt.empty(self.cshape, _C16)
        C.check(C.lib.vpm_r2c(rt.ctx, C.i3(self._n3), _ptr(real.t), _ptr(o), float(scale)))
        return ComplexField(self, o)

    def _c2r(self, cplx, scale, consume=False):
        """``consume``: the caller owns ``cplx`` and gives it up (no input-preserving copy)"""
        rt = self.rt
        rt.sync_stream()
        if self.P > 1:
            return self._c2r_slab(cplx, scale, consume)
        o = rt.empty(self.rshape)
        fn = C.lib.vpm_c2r_inplace if consume else C.lib.vpm_c2r
        C.check(fn(rt.ctx, C.i3(self._n3), _ptr(cplx.t), _ptr(o), float(scale)))
        return RealField(self, o)

    def _peer(self):
        """peer-memory exchange for the FFT transposes (None: fall back to NCCL all-to-all)"""
        if getattr(self, '_peer_x', False) is False:
            self._peer_x = None
            import os
            if os.environ.get('VMAD_B200_TRANSPOSE', 'peer') == 'peer' and 1 < self.P <= 8 \
                    and getattr(self.comm, 'device', None) is not None and self.comm.device.type == 'cuda':
                from .dist import PeerExchange
                slab_bytes = 16 * int(numpy.prod(self.cshape))
                px, why = None, None
                try:
                    px = PeerExchange(self.comm, 3 * slab_bytes)
                except Exception as e:  # no peer access / IPC on this rank
                    why = e
                # all ranks or none: a mixed choice would deadlock
                if self.comm.allreduce(int(px is not None)) == self.P:
                    self._peer_x = px
                else:
                    import warnings
                    warnings.warn("peer-memory exchange unavailable (%r); using NCCL all-to-all" % (why,))
        return self._peer_x

    def _fft_chunks(self):
        """plane groups of the pipelined forward slab transform (1: no pipelining, the default: measured
        on 2 GPUs at 512^3 the pipelined form gains nothing -- 296.8 / 296.5 / 298.2 ms per step with 1 / 2 / 4
        groups -- the put and the 2-D transforms compete for the same HBM bandwidth)"""
        k = getattr(self, '_fft_k', None)
        if k is None:
            import os
            k = int(os.environ.get('VMAD_B200_FFT_CHUNKS', '1'))
            n0l = self.rshape[0]
            slab_bytes = 16 * int(numpy.prod(self.cshape))
            while k > 1 and (n0l % k or slab_bytes // k < (8 << 20)):
                k //= 2
            self._fft_k = k = max(k, 1)
        return k

    def _side_stream(self):
        s = getattr(self, '_side', None)
        if s is None:
            self._side = s = torch.cuda.Stream(device=self.rt.device)
        return s

    # slab-decomposed transforms: local cuFFT, one all-to-all transpose each -------------------
    def _r2c_slab(self, real, scale):
        """[n0l, N1, N2] real -> [N0, n1l, Nh] complex: 2-D D2Z over the local planes, pack,
        all-to-all (the distributed-FFT transpose), 1-D Z2Z along axis 0"""
        rt = self.rt
        P = self.P
        n0l, N1, N2 = self.rshape
        N0, n1l, Nh = self.cshape
        a = rt.empty((n0l, N1, Nh), _C16)
        px = self._peer()
        K = self._fft_chunks() if px is not None else 1
        if K > 1:
            # pipeline over groups of planes: while group c travels (pack + NVLink stores on a side
            # stream; the flag protocol spans the K launches), the 2-D transforms of group c + 1 run
            inner = n1l * Nh
            pc = n0l // K
            _, b = px.begin()
            main = torch.cuda.current_stream()
            side = self._side_stream()
            for c in range(K):
                rt.sync_stream()
                C.check(C.lib.vpm_fft_r2c_planes(rt.ctx, pc, N1, N2, _ptr(real.t[c * pc:(c + 1) * pc]),
                                                 _ptr(a[c * pc:(c + 1) * pc])))
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                rt.push_stream(side)
                try:
                    px.put(b, a[c * pc:(c + 1) * pc], pc, inner, P * inner, inner,
                           (self.rank * n0l + c * pc) * inner, first=(c == 0), last=(c == K - 1))
                finally:
                    rt.pop_stream()
            main.wait_stream(side)
            o = rt.empty(self.cshape, _C16)
            C.check(C.lib.vpm_fft_c2c_axis0_oop(rt.ctx, N0, inner, px.recv_ptr(b), _ptr(o), 0, float(scale)))
            return ComplexField(self, o)
        C.check(C.lib.vpm_fft_r2c_planes(rt.ctx, n0l, N1, N2, _ptr(real.t), _ptr(a)))
        if px is not None:
            # pack + transfer in one kernel: block d is stored straight into rank d's buffer
            inner = n1l * Nh
            e, b = px.begin()
            px.put(b, a, n0l, inner, P * inner, inner, self.rank * n0l * inner)
            o = rt.empty(self.cshape, _C16)
            C.check(C.lib.vpm_fft_c2c_axis0_oop(rt.ctx, N0, inner, px.recv_ptr(b),
                                                _ptr(o), 0, float(scale)))
            return ComplexField(self, o)
        # [n0l, P, n1l*Nh] -> [P, n0l, n1l*Nh]: block d goes to rank d
        send = rt.empty((P, n0l, n1l, Nh), _C16)
        C.check(C.lib.vpm_swap_leading_axes(rt.ctx, n0l, P, n1l * Nh, _ptr(a), _ptr(send)))
        recv = self.comm.alltoall_blocks(send)          # [P (source slab), n0l, n1l, Nh] = [N0, n1l, Nh]
        C.check(C.lib.vpm_fft_c2c_axis0(rt.ctx, N0, n1l * Nh, _ptr(recv), 0, float(scale)))
        return ComplexField(self, recv.view(N0, n1l, Nh))

    def _r2c_slab_many(self, reals, scale):
        """several real fields at once: one all-to-all for all of them"""
        rt = self.rt
        rt.sync_stream()
        P, K = self.P, len(reals)
        n0l, N1, N2 = self.rshape
        N0, n1l, Nh = self.cshape
        px = self._peer()
        if px is not None and K <= 3:
            inner = n1l * Nh
            slab = N0 * inner
            e, b = px.begin()
            for k, real in enumerate(reals):
                a = rt.empty((n0l, N1, Nh), _C16)
                C.check(C.lib.vpm_fft_r2c_planes(rt.ctx, n0l, N1, N2, _ptr(real.t), _ptr(a)))
                px.put(b, a, n0l, inner, P * inner, inner, k * slab + self.rank * n0l * inner, first=(k == 0), last=(k == K - 1))
            out = []
            for k in range(K):
                o = rt.empty(self.cshape, _C16)
                C.check(C.lib.vpm_fft_c2c_axis0_oop(rt.ctx, N0, inner, px.recv_ptr(b, k * slab),
                                                    _ptr(o), 0, float(scale)))
                out.append(ComplexField(self, o))
            return out
        send = rt.empty((P, K, n0l, n1l, Nh), _C16)
        a = rt.empty((n0l, N1, Nh), _C16)
        tmp = rt.empty((P, n0l, n1l, Nh), _C16)
        for k, real in enumerate(reals):
            C.check(C.lib.vpm_fft_r2c_planes(rt.ctx, n0l, N1, N2, _ptr(real.t), _ptr(a)))
            C.check(C.lib.vpm_swap_leading_axes(rt.ctx, n0l, P, n1l * Nh, _ptr(a), _ptr(tmp)))
            send[:, k].copy_(tmp)
        recv = self.comm.alltoall_blocks(send)            # [P, K, n0l, n1l, Nh]
        out = []
        for k in range(K):
            o = recv[:, k].contiguous().view(N0, n1l, Nh)
            C.check(C.lib.vpm_fft_c2c_axis0(rt.ctx, N0, n1l * Nh, _ptr(o), 0, float(scale)))
            out.append(ComplexField(self, o))
        return out

    def _c2r_slab_many(self, cplxs, scale):
        """several complex fields (given up by the caller) at once: one all-to-all"""
        rt = self.rt
        rt.sync_stream()
        P, K = self.P, len(cplxs)
        n0l, N1, N2 = self.rshape
        N0, n1l, Nh = self.cshape
        px = self._peer()
        if px is not None and K <= 3:
            inner = n1l * Nh
            blk = n0l * inner
            slab = P * blk
            e, b = px.begin()
            for k, c in enumerate(cplxs):
                C.check(C.lib.vpm_fft_c2c_axis0(rt.ctx, N0, inner, _ptr(c.t), 1, 1.0))
                px.put(b, c.t, n0l, inner, inner, blk, k * slab + self.rank * inner, first=(k == 0),
                       last=(k == K - 1), d_row=P * inner)
            out = []
            for k in range(K):
                o = rt.empty(self.rshape)
                C.check(C.lib.vpm_fft_c2r_planes(rt.ctx, n0l, N1, N2, px.recv_ptr(b, k * slab), _ptr(o), float(scale)))
                out.append(RealField(self, o))
            return out
        send = rt.empty((P, K, n0l, n1l, Nh), _C16)
        for k, c in enumerate(cplxs):
            C.check(C.lib.vpm_fft_c2c_axis0(rt.ctx, N0, n1l * Nh, _ptr(c.t), 1, 1.0))
            send[:, k].copy_(c.t.view(P, n0l, n1l, Nh))
        recv = self.comm.alltoall_blocks(send)            # [P (column block), K, n0l, n1l, Nh]
        out = []
        a = rt.empty((n0l, N1, Nh), _C16)
        for k in range(K):
            blk = recv[:, k].contiguous()
            C.check(C.lib.vpm_swap_leading_axes(rt.ctx, P, n0l, n1l * Nh, _ptr(blk), _ptr(a)))
            o = rt.empty(self.rshape)
            C.check(C.lib.vpm_fft_c2r_planes(rt.ctx, n0l, N1, N2, _ptr(a), _ptr(o), float(scale)))
            out.append(RealField(self, o))
        return out

    def _c2r_slab(self, cplx, scale, consume=False, out=None):
        """``out``: a contiguous [n0l, N1, N2] tensor to receive the result (the interior of a
        halo-padded array)"""
        rt = self.rt
        P = self.P
        n0l, N1, N2 = self.rshape
        N0, n1l, Nh = self.cshape
        px = self._peer()
        if px is not None:
            inner = n1l * Nh
            blk = n0l * inner
            work = rt.empty(self.cshape, _C16)
            C.check(C.lib.vpm_fft_c2c_axis0_oop(rt.ctx, N0, inner, _ptr(cplx.t), _ptr(work), 1, 1.0))
            e, b = px.begin()
            o = rt.empty(self.rshape) if out is None else out
            if _C2R_STRIDED:
                # block d = the planes of rank d; its rows land [plane, source's columns] in d's receive
                # area, i.e. in [n0l, N1, Nh] order: the 2-D inverse transforms read the area in place
                px.put(b, work, n0l, inner, inner, blk, self.rank * inner, d_row=P * inner)
                C.check(C.lib.vpm_fft_c2r_planes(rt.ctx, n0l, N1, N2, px.recv_ptr(b), _ptr(o), float(scale)))
                return RealField(self, o)
            px.put(b, work, 1, blk, 0, blk, self.rank * blk)
            a = rt.empty((n0l, N1, Nh), _C16)
            C.check(C.lib.vpm_swap_leading_axes(rt.ctx, P, n0l, inner, px.recv_ptr(b), _ptr(a)))
            C.check(C.lib.vpm_fft_c2r_planes(rt.ctx, n0l, N1, N2, _ptr(a), _ptr(o), float(scale)))
            return RealField(self, o)
        work = cplx.t if consume else cplx.t.clone()    # operators must not mutate their inputs
        C.check(C.lib.vpm_fft_c2c_axis0(rt.ctx, N0, n1l * Nh, _ptr(work), 1, 1.0))
        recv = self.comm.alltoall_blocks(work.view(P, n0l, n1l, Nh))   # [P (column block), n0l, n1l, Nh]
        a = rt.empty((n0l, N1, Nh), _C16)
        C.check(C.lib.vpm_swap_leading_axes(rt.ctx, P, n0l, n1l * Nh, _ptr(recv), _ptr(a)))
        o = rt.empty(self.rshape) if out is None else out
        C.check(C.lib.vpm_fft_c2r_planes(rt.ctx, n0l, N1, N2, _ptr(a), _ptr(o), float(scale)))
        return RealField(self, o)

    def _halo_fill(self, padded, width=2):
        """fill the ``width`` halo planes on each side of a [n0l + 2 width, N1, N2] slab array from
        the neighbouring ranks' boundary planes: ONE peer put (my lower planes -> previous rank's
        upper halo, my upper planes -> next rank's lower halo) + one copy kernel"""
        rt = self.rt
        px = self._peer()
        n0l, N1, N2 = self.rshape
        cnt = width * N1 * N2 // 2          # `width` planes of doubles, in complex elements
        plane = N1 * N2 * 8
        prev, nxt = (self.rank - 1) % self.P, (self.rank + 1) % self.P
        base = padded.data_ptr()
        px.ensure(32 * cnt)
        _, b = px.begin()
        # interior = planes [width, width + n0l): its first `width` planes go to the previous rank's
        # HI slot, its last `width` planes to the next rank's LO slot -- one launch
        px.put_blocks(b, base, cnt, [(prev, cnt, cnt), (nxt, n0l * N1 * N2 // 2, 0)])
        rt.sync_stream()
        lo = ctypes.c_void_p(base)
        hi = ctypes.c_void_p(base + (n0l + width) * plane)
        C.check(C.lib.vpm_copy2(rt.ctx, cnt, px.recv_ptr(b, 0), lo, px.recv_ptr(b, cnt), hi))

    # -- k-space (fastpm.py:79,83,87) ---------------------------------------------------------------
    def _transfer(self, cplx, scale=1.0, laplace=False, tables=(None, None, None), conj=False):
        """out = in * scale * [-1/k^2] * prod_d tables[d][i_d]; tables indexed by mesh axis"""
        rt = self.rt
        rt.sync_stream()
        o = rt.empty(self.cshape, _C16)
        off = 3 - self.ndim
        a = [None, None, None]
        for d, t in enumerate(tables[:self.ndim]):
            a[d + off] = t
        C.check(C.lib.vpm_transfer(rt.ctx, C.i3(self._cshape3), _ptr(cplx.t), _ptr(o), float(scale),
                                   1 if laplace else 0, _ptr(self._k_table(0)),
                                   _ptr(self._k_table(1)), _ptr(self._k_table(2)), _ptr(a[0]),
                                   _ptr(a[1]), _ptr(a[2]), 1 if conj else 0))
        return ComplexField(self, o)

    def _prepare_table(self, table):
        """host multiplier table (broadcastable to the complex shape) -> (device tensor,
        element strides with 0 on broadcast axes, is_complex)"""
        table = numpy.asarray(table)
        if table.ndim == 0:
            table = table.reshape((1,) * self.ndim)
        if table.ndim != self.ndim:
            raise ValueError("transfer table must be broadcastable to %s" % (self.cshape,))
        is_c = numpy.iscomplexobj(table)
        tt = _upload(table, numpy.complex128 if is_c else numpy.float64)
        strides = []
        acc = 1
        for d in reversed(range(self.ndim)):
            if table.shape[d] == 1:
                strides.append(0)
            elif table.shape[d] == self.cshape[d]:
                strides.append(acc)
                acc *= table.shape[d]
            else:
                raise ValueError("transfer table shape %s does not broadcast to %s"
                                 % (table.shape, self.cshape))
        strides = [0] * (3 - self.ndim) + list(reversed(strides))
        return tt, strides, is_c

    def _transfer_table(self, cplx, table, conj=False, prepared=None):
        """generic multiplier: a broadcastable host table (any axis may have length 1)"""
        rt = self.rt
        tt, strides, is_c = prepared if prepared is not None else self._prepare_table(table)
        rt.sync_stream()
        o = rt.empty(self.cshape, _C16)
        C.check(C.lib.vpm_transfer_table(rt.ctx, C.i3(self._cshape3), _ptr(cplx.t), _ptr(o),
                                         _ptr(tt), C.i3(strides), 1 if is_c else 0,
                                         1 if conj else 0))
        return ComplexField(self, o)

    def _transfer_callable(self, cplx, tf, kind, conj):
        """``x * tf(k)`` for an arbitrary python ``tf``: the table is evaluated on the host the
        first time this function (same object, same closure VALUES) is seen and then stays on the
        device (a model built once keeps calling the same closure, e.g. ``induce_correlation``'s)."""
        from .vm.ops import _hyper_key, _Uncacheable
        try:        # by value: the function, its closure cells and defaults (a callable whose
            key = (_hyper_key(tf), kind)    # state cannot be fingerprinted is evaluated every time)
        except _Uncacheable:
            return self._transfer_table(cplx, tf(cplx._coords(kind)), conj)
        hit = self._tf_cache.get(key)
        if hit is None:
            prepared = self._prepare_table(tf(cplx._coords(kind)))
            if len(self._tf_cache) >= 32:
                self._tf_cache.pop(next(iter(self._tf_cache)))
            hit = (tf, prepared)  # holding tf keeps the identity part of its key unique
            self._tf_cache[key] = hit
        return self._transfer_table(cplx, None, conj, prepared=hit[1])

    @property
    def digitized(self):
        """device side of ``apply_digitized`` (fastpm.py:97-213)"""
        return _Digitized(self)

    def force_kspace(self, rhok, scale, grad_tables, want_p=True):
        """fused ``gravity`` k-space pass (fastpm.py:424-431); 3-D only"""
        rt = self.rt
        rt.sync_stream()
        assert self.ndim == 3
        p = rt.empty(self.cshape, _C16) if want_p else None
        g = [rt.empty(self.cshape, _C16) for _ in range(3)]
        C.check(C.lib.vpm_force_kspace(rt.ctx, C.i3(self._cshape3), _ptr(rhok.t), float(scale),
                                       _ptr(self._k_table(0)), _ptr(self._k_table(1)),
                                       _ptr(self._k_table(2)), _ptr(grad_tables[0]),
                                       _ptr(grad_tables[1]), _ptr(grad_tables[2]), _ptr(p),
                                       _ptr(g[0]), _ptr(g[1]), _ptr(g[2])))
        return (ComplexField(self, p) if want_p else None), [ComplexField(self, t) for t in g]


class _Digitized(object):
    """binned transfer functions on the device: index arrays from the user's digitizer, the
    per-mode multiply ``x * table[bin]`` and the binned reduction of the vjp w.r.t. the table"""

    def __init__(self, pm):
        self.pm = pm

    def bins(self, ind, conj):
        """host (ind, conj) of the local complex shape -> (int32 device tensor, uint8 tensor or None)"""
        pm = self.pm
        ind = numpy.ascontiguousarray(numpy.broadcast_to(numpy.asarray(ind), pm.cshape))
        ind = numpy.clip(ind, -1, 2 ** 31 - 2).astype('i4')        # anything below 0 is "no bin"
        conj = numpy.broadcast_to(numpy.asarray(conj, dtype='?'), pm.cshape)
        ind_d = torch.from_numpy(ind).to(pm.rt.device)
        conj_d = torch.from_numpy(numpy.ascontiguousarray(conj).astype('u1')).to(pm.rt.device) if conj.any() else None
        return ind_d, conj_d

    def mul(self, x, ind, conj, table, dflt, conj_table=False):
        pm, rt = self.pm, self.pm.rt
        x = pm._as_field(x, True)
        t = _upload(numpy.asarray(table).ravel(), numpy.complex128)
        rt.sync_stream()
        o = rt.empty(pm.cshape, _C16)
        C.check(C.lib.vpm_digitized_mul(rt.ctx, x.t.numel(), _ptr(x.t), _ptr(ind), _ptr(conj), t.numel(), _ptr(t),
                                        float(dflt), 1 if conj_table else 0, _ptr(o)))
        return ComplexField(pm, o)

    def bin(self, x, ybar, ind, conj, nbins, eit):
        """local part of the cotangent of the table: a host array of ``nbins`` doubles"""
        pm, rt = self.pm, self.pm.rt
        x, ybar = pm._as_field(x, True), pm._as_field(ybar, True)
        e = _upload(numpy.asarray(eit).ravel(), numpy.complex128) if eit is not None else None
        rt.sync_stream()
        out = rt.empty((nbins,))
        C.check(C.lib.vpm_digitized_bin(rt.ctx, C.i3(pm._cshape3), int(pm.Nmesh[-1]), _ptr(x.t), _ptr(ybar.t),
                                        _ptr(ind), _ptr(conj), int(nbins), _ptr(e), _ptr(out)))
        return out.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# fused force evaluation (the `gravity` graph of fastpm.py:409-438 as three calls)
# ---------------------------------------------------------------------------------------------
class ForceState(object):
    """what one fused force evaluation leaves on the tape for its vjp / jvp"""
    __slots__ = ('xl', 'layout', 'F', 'potk', 'phi', 'scale', 'grad_tables')

    def __init__(self, xl, layout, F, potk, phi, scale, grad_tables):
        self.xl = xl
        self.layout = layout
        self.F = F            # the three real force meshes, or None when they are recomputed
        self.potk = potk      # slab path with recompute: 1 complex mesh instead of 3 real ones
        self.phi = phi        # single-GPU path: the real potential; F_d = -D_d phi by stencil
        self.scale = scale
        self.grad_tables = grad_tables

    def meshes(self, pm):
        if self.F is not None:
            return self.F
        if self.phi is not None:
            return _fd4(pm, self.phi)
        return [pm._c2r(pm._transfer(self.potk, tables=_only(d, self.grad_tables[d])), 1.0, consume=True)
                for d in range(3)]


def _only(d, table):
    t = [None, None, None]
    t[d] = table
    return t


def _h3(pm):
    return (ctypes.c_double * 3)(*[float(pm.BoxSize[d] / pm.Nmesh[d]) for d in range(3)])


class _Padded(object):
    """a slab array with two halo planes on each side of axis 0"""
    __slots__ = ('t',)

    def __init__(self, t):
        self.t = t


def _fd4(pm, phi):
    """F_d = -D_d phi: the three force meshes from the real potential (4-point stencil)"""
    rt = pm.rt
    rt.sync_stream()
    F = [rt.empty(pm.rshape) for _ in range(3)]
    if isinstance(phi, _Padded):
        C.check(C.lib.vpm_fd4_neg_gradient_slab(rt.ctx, C.i3(pm.rshape), _h3(pm), 2, _ptr(phi.t), _ptr(F[0]),
                                                _ptr(F[1]), _ptr(F[2])))
    else:
        C.check(C.lib.vpm_fd4_neg_gradient(rt.ctx, C.i3(pm._n3), _h3(pm), _ptr(phi.t), _ptr(F[0]),
                                           _ptr(F[1]), _ptr(F[2])))
    return [RealField(pm, t) for t in F]


def _potential(pm, potk, consume=False):
    """the real potential c2r(potk) in the form the stencil wants: the periodic mesh on one rank,
    the slab with its halo planes (one extra 0.5 MB exchange) on several"""
    if pm.P == 1:
        return pm._c2r(potk, 1.0, consume=consume)
    n0l, N1, N2 = pm.rshape
    pad = pm.rt.empty((n0l + 4, N1, N2))
    pm.rt.sync_stream()
    pm._c2r_slab(potk, 1.0, consume=consume, out=pad[2:2 + n0l])
    pm._halo_fill(pad)
    return _Padded(pad)


def _force_tables(pm, order=1):
    """per-axis finite-difference gradient tables -i (8 sin w - sin 2w) / (6 h) (fastpm.py:316-323),
    evaluated once on the host with numpy -- the numbers the oracle uses"""
    key = ('force_tables', order)
    if key not in pm._tables:
        tabs = []
        for d in range(3):
            h = pm.BoxSize[d] / pm.Nmesh[d]
            w = pm.freq_index(d) * (2 * numpy.pi / pm.BoxSize[d]) * h
            a = 1 / (6.0 * h) * (8 * numpy.sin(w) - numpy.sin(2 * w))
            tabs.append(_upload(-1j * a))
        pm._tables[key] = tabs
    return pm._tables[key]


_HALO_STENCIL = os.environ.get('VMAD_B200_HALO_STENCIL', '1') == '1'


def _use_stencil(pm):
    """the real-space gradient needs +-2 planes of neighbours: the periodic mesh itself on one
    rank, a halo exchange with the neighbouring slabs on several (else the k-space gradients)"""
    if not getattr(pm, 'force_stencil', True) or min(pm.rshape) < 2:
        return False
    if pm.P == 1:
        return True
    # several ranks: halo planes come from the immediate neighbours over peer memory; the put
    # kernel moves 16-byte words, so a slab must hold an even number of doubles (odd meshes
    # keep the k-space gradients)
    n0l, N1, N2 = pm.rshape
    return _HALO_STENCIL and n0l >= 2 and (n0l * N1 * N2) % 2 == 0 and (N1 * N2) % 2 == 0 \
        and pm._peer() is not None


def drift_pos(pm, dx, p, fac, q):
    """``dx1 = dx + fac p`` and ``x1 = q + dx1`` in one pass (vpm_drift_pos)"""
    rt = pm.rt
    dx, p, q = asdevice(dx), asdevice(p), asdevice(q)
    rt.sync_stream()
    dx1 = rt.empty(tuple(dx.t.shape))
    x1 = rt.empty(tuple(dx.t.shape))
    C.check(C.lib.vpm_drift_pos(rt.ctx, dx.t.numel(), _ptr(dx.t), _ptr(p.t), float(fac), _ptr(q.t),
                                _ptr(dx1), _ptr(x1)))
    return DeviceArray(dx1), DeviceArray(x1)


def force_apl(pm, dx, q, x=None, kick=None, want_f=True):
    """f = -grad laplace^-1 (rho Nc / Np) read out at x = q + dx; returns (f, potk, state).
    ``x``: the positions when the caller already has them (fused drift).  ``kick=(p, fac)``:
    returns ``(f, potk, state, p + fac f)`` with the kick folded into the force readout;
    ``want_f=False`` then returns ``f = None`` (a stage whose force nobody reads)."""
    if pm.ndim != 3:
        raise ValueError("the fused force operator is 3-D only")
    rt = pm.rt
    q = asdevice(q)
    if x is None:
        dx = asdevice(dx)
        x = q + dx
    layout = pm.decompose(x)
    xl = layout.exchange(x)
    rho = pm.paint(xl, 1.0)
    rhok = pm._r2c(rho, 1.0)                                   # un-normalised cuFFT output
    npart = pm._global_count(len(q))
    scale = 1.0 / npart                                        # (Nc / Np) * (1 / Nc)
    tabs = _force_tables(pm)
    if _use_stencil(pm):
        # one inverse FFT (of the potential) + a streaming 4-point stencil for the 3 gradients
        potk = pm._transfer(rhok, scale=scale, laplace=True)
        phi = _potential(pm, potk)                             # potk is an output: preserved
        F = _fd4(pm, phi)
        # small meshes: the three force meshes stay on the tape (24 B/cell) and the reverse sweep
        # skips the stencil pass; large ones recompute them from the potential
        keep = not getattr(pm, 'force_recompute', False) and 24 * phi.t.numel() <= (512 << 20)
        st = ForceState(xl, layout, F if keep else None, None, phi, scale, tabs)
        if kick is not None:
            f, p1 = pm.readout3(F[0], F[1], F[2], xl, gather=layout, kick=kick, want_value=want_f)
            return f, potk, st, p1
        return pm.readout3(F[0], F[1], F[2], xl, gather=layout), potk, st
    potk, g = pm.force_kspace(rhok, scale, tabs, want_p=True)
    if pm.P > 1:
        F = pm._c2r_slab_many(g, 1.0)                  # three inverse transforms, one all-to-all
    else:
        F = [pm._c2r(gd, 1.0, consume=True) for gd in g]   # g are temporaries of this call
    if getattr(pm, 'force_recompute', False):
        # trade 3 inverse FFTs in the reverse sweep for 2/3 of the mesh memory on the tape
        st = ForceState(xl, layout, None, potk, None, scale, tabs)
    else:
        st = ForceState(xl, layout, F, None, None, scale, tabs)
    if kick is not None:
        f, p1 = pm.readout3(F[0], F[1], F[2], xl, gather=layout, kick=kick, want_value=want_f)
        return f, potk, st, p1
    return pm.readout3(F[0], F[1], F[2], xl, gather=layout), potk, st


def force_vjp(pm, st, _f, _potk, f_scale=1.0, base=None, update=None):
    """reverse sweep of force_apl: cotangents of (f, potk) -> cotangent of dx.
    ``update=(y, fac)``: returns ``(g, y + fac g)`` (y may be a zero literal), the second array
    written by the same kernel when the layout is one-rank.
    ``f_scale``: the force cotangent is ``f_scale * _f`` (the factor of the kick it comes from,
    applied while exchanging); ``base``: added to the result (the cotangent reaching dx from its
    other consumers), inside the store of the last kernel where possible."""
    rt = pm.rt
    lay, xl = st.layout, st.xl
    f_zero = _is_zero_literal(_f)
    n = xl.t.shape[0]
    one_rank = isinstance(lay, Layout) and lay.newlength == n == lay.oldlength
    if f_zero:
        _fl = DeviceArray(torch.zeros((n, 3), dtype=_F8, device=rt.device))
    elif f_scale == 1.0:
        _fl = lay.exchange(asdevice(_f))
    elif one_rank:
        _f = asdevice(_f)
        rt.sync_stream()
        t = rt.empty((n, 3))
        C.check(C.lib.vpm_take_rows_scaled(rt.ctx, _ptr(_f.t), _ptr(lay.indices), n, 3, float(f_scale), _ptr(t)))
        _fl = DeviceArray(t, cellindex=lay)
    else:
        _fl = lay.exchange(asdevice(_f), scale=f_scale)
    # readout.vjp (mesh half) x3: three paints of the columns of the force cotangent
    halo = st.phi is not None and pm.P > 1
    Fbar = pm.paint_cols3(xl, _fl, pad=2 if halo else 0)
    pbar = None if _is_zero_literal(_potk) else pm._as_field(_potk, True)
    if halo:
        # slab: the axis-0 difference of the adjoint needs the neighbours' boundary planes of Fbar_0
        pm._halo_fill(Fbar[0])
        phibar = rt.empty(pm.rshape)
        C.check(C.lib.vpm_fd4_neg_gradient_adjoint_slab(rt.ctx, C.i3(pm.rshape), _h3(pm), 2, _ptr(Fbar[0]),
                                                        _ptr(Fbar[1]), _ptr(Fbar[2]), _ptr(phibar)))
        del Fbar
        potk_bar = pm._r2c(RealField(pm, phibar), 1.0)
        if pbar is not None:
            potk_bar = potk_bar + pbar
        rhok_bar = pm._transfer(potk_bar, scale=st.scale, laplace=True)
    elif st.phi is not None:
        # adjoint of the stencil, then ONE forward FFT (c2r.vjp of the un-normalised inverse)
        phibar = rt.empty(pm.rshape)
        C.check(C.lib.vpm_fd4_neg_gradient_adjoint(rt.ctx, C.i3(pm._n3), _h3(pm), _ptr(Fbar[0].t),
                                                   _ptr(Fbar[1].t), _ptr(Fbar[2].t), _ptr(phibar)))
        del Fbar
        potk_bar = pm._r2c(RealField(pm, phibar), 1.0)
        if pbar is not None:
            potk_bar = potk_bar + pbar
        rhok_bar = pm._transfer(potk_bar, scale=st.scale, laplace=True)
    else:
        gbar = pm._r2c_slab_many(Fbar, 1.0) if pm.P > 1 else [pm._r2c(Fd, 1.0) for Fd in Fbar]
        del Fbar
        out = rt.empty(pm.cshape, _C16)
        C.check(C.lib.vpm_force_kspace_vjp(
            rt.ctx, C.i3(pm._cshape3), _ptr(gbar[0].t), _ptr(gbar[1].t), _ptr(gbar[2].t),
            _ptr(pbar.t) if pbar is not None else None, float(st.scale), _ptr(pm._k_table(0)),
            _ptr(pm._k_table(1)), _ptr(pm._k_table(2)), _ptr(st.grad_tables[0]),
            _ptr(st.grad_tables[1]), _ptr(st.grad_tables[2]), _ptr(out)))
        rhok_bar = ComplexField(pm, out)
    # r2c.vjp of the un-normalised transform is the un-normalised inverse
    rho_bar = pm._c2r(rhok_bar, 1.0, consume=True)
    # particle half: readout.vjp x3 and paint.vjp in one pass
    F = st.meshes(pm)
    rt.sync_stream()
    g = rt.empty((n, 3))
    if base is not None and _is_zero_literal(base):
        base = None
    y = fac = None
    if update is not None:
        y, fac = update
        y = None if _is_zero_literal(y) else asdevice(y)
    if one_rank:        # gather, the accumulation of `base` and the update of `y` in the store
        bt = asdevice(base).t if base is not None else None
        yo = rt.empty((n, 3)) if y is not None else None
        C.check(C.lib.vpm_readout_grad4(rt.ctx, ctypes.byref(pm._mesh), _ptr(F[0].t), _ptr(F[1].t),
                                        _ptr(F[2].t), _ptr(rho_bar.t), _ptr(xl.t), _ptr(_fl.t), n,
                                        _ptr(lay.indices), _ptr(bt), _ptr(g),
                                        _ptr(y.t) if y is not None else None,
                                        float(fac) if y is not None else 0.0, _ptr(yo)))
        out = DeviceArray(g)
        if update is None:
            return out
        return out, (DeviceArray(yo) if y is not None else out * float(fac))
    home = lay.home_rows() if hasattr(lay, 'home_rows') else None
    if home is not None:
        # several ranks: local home rows (+ base) stored by the kernel, remote partial sums added after
        bt = asdevice(base).t if base is not None else None
        acc = bt.clone() if bt is not None else torch.zeros((lay.oldlength, 3), dtype=_F8, device=rt.device)
        C.check(C.lib.vpm_readout_grad4_home(rt.ctx, ctypes.byref(pm._mesh), _ptr(F[0].t), _ptr(F[1].t),
                                             _ptr(F[2].t), _ptr(rho_bar.t), _ptr(xl.t), _ptr(_fl.t), n,
                                             _ptr(home), _ptr(g), _ptr(bt), _ptr(acc)))
        lay.gather_remote(g, acc, 1.0)
        out = DeviceArray(acc)
    else:
        C.check(C.lib.vpm_readout_grad4(rt.ctx, ctypes.byref(pm._mesh), _ptr(F[0].t), _ptr(F[1].t),
                                        _ptr(F[2].t), _ptr(rho_bar.t), _ptr(xl.t), _ptr(_fl.t), n,
                                        None, None, _ptr(g), None, 0.0, None))
        out = lay.gather(DeviceArray(g), init=base)
    if update is None:
        return out
    return out, (y._axpby(1.0, float(fac), out.t) if y is not None else out * float(fac))


def lpt1_apl(pm, rhok, q):
    """Zel'dovich displacement (fastpm.py:336-351) read out at the constant positions ``q``:
    ONE inverse FFT of the potential + the 4-point stencil (the Fourier symbol of
    fourier_space_neg_gradient(order=1), as in force_apl) + the three-mesh readout; the layout
    and the sorted copy of ``q`` are cached with ``q``.  Across ranks the stencil reads the
    neighbours' boundary planes (halo exchange, as the force does).  Returns ``(dx1, state)``."""
    q = asdevice(q)
    layout = pm.decompose(q)
    xl = layout.exchange(q)
    phi = _potential(pm, pm._transfer(pm._as_field(rhok, True), scale=1.0, laplace=True), consume=True)
    F = _fd4(pm, phi)
    return pm.readout3(F[0], F[1], F[2], xl, gather=layout), (layout, xl)


def _fd4_adjoint(pm, Fbar):
    """phibar = adjoint of F_d = -D_d phi applied to the three cotangent meshes; across ranks
    ``Fbar`` is the padded [3, n0l + 4, N1, N2] array of ``paint_cols3(pad=2)``"""
    rt = pm.rt
    phibar = rt.empty(pm.rshape)
    if pm.P > 1:
        pm._halo_fill(Fbar[0])
        C.check(C.lib.vpm_fd4_neg_gradient_adjoint_slab(rt.ctx, C.i3(pm.rshape), _h3(pm), 2, _ptr(Fbar[0]),
                                                        _ptr(Fbar[1]), _ptr(Fbar[2]), _ptr(phibar)))
    else:
        C.check(C.lib.vpm_fd4_neg_gradient_adjoint(rt.ctx, C.i3(pm._n3), _h3(pm), _ptr(Fbar[0].t),
                                                   _ptr(Fbar[1].t), _ptr(Fbar[2].t), _ptr(phibar)))
    return RealField(pm, phibar)


def lpt1_vjp(pm, state, _dx1):
    """cotangent of dx1 -> cotangent of rhok (``q`` is a constant here)"""
    layout, xl = state
    _fl = layout.exchange(asdevice(_dx1))
    Fbar = pm.paint_cols3(xl, _fl, pad=2 if pm.P > 1 else 0)
    phibar = _fd4_adjoint(pm, Fbar)
    del Fbar
    # c2r.vjp of the un-normalised inverse is the un-normalised forward transform
    return pm._transfer(pm._r2c(phibar, 1.0), scale=1.0, laplace=True)


def lpt1_jvp(pm, state, rhok_):
    if _is_zero_literal(rhok_):
        return 0
    layout, xl = state
    phi_ = _potential(pm, pm._transfer(pm._as_field(rhok_, True), scale=1.0, laplace=True), consume=True)
    F_ = _fd4(pm, phi_)
    return pm.readout3(F_[0], F_[1], F_[2], xl, gather=layout)


def _lpt2_potential(pm, rhok):
    """phi = c2r(laplace rhok) and the three meshes F_j = -D_j phi"""
    phi = pm._c2r(pm._transfer(pm._as_field(rhok, True), scale=1.0, laplace=True), 1.0, consume=True)
    return phi, _fd4(pm, phi)


def _lpt2_bilinear(pm, F, G, scale):
    rt = pm.rt
    rt.sync_stream()
    out = rt.empty(pm.rshape)
    if pm.P > 1:        # F, G: tensors with two halo planes per side (_lpt2_gradients_slab)
        g = [None, None, None] if G is None else [_ptr(t) for t in G]
        C.check(C.lib.vpm_lpt2_source_slab(rt.ctx, C.i3(pm.rshape), _h3(pm), 2, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]),
                                           g[0], g[1], g[2], float(scale), _ptr(out)))
        return RealField(pm, out)
    g = [None, None, None] if G is None else [_ptr(t.t) for t in G]
    C.check(C.lib.vpm_lpt2_source(rt.ctx, C.i3(pm._n3), _h3(pm), _ptr(F[0].t), _ptr(F[1].t), _ptr(F[2].t),
                                  g[0], g[1], g[2], float(scale), _ptr(out)))
    return RealField(pm, out)


def _lpt2_potential_slab(pm, rhok):
    """several ranks: the potential with FOUR halo planes per side (the source is two stencils
    deep), as a [n0l + 8, N1, N2] tensor: one inverse FFT + one halo exchange"""
    rt = pm.rt
    n0l, N1, N2 = pm.rshape
    phi4 = rt.empty((n0l + 8, N1, N2))
    rt.sync_stream()
    pm._c2r_slab(pm._transfer(pm._as_field(rhok, True), scale=1.0, laplace=True), 1.0, consume=True,
                 out=phi4[4:4 + n0l])
    pm._halo_fill(phi4, width=4)
    return phi4


def _lpt2_gradients_slab(pm, phi4):
    """F_j = -D_j phi on the local planes AND two planes beyond on each side: three
    [n0l + 4, N1, N2] tensors = the gradient meshes with their own halo planes (no exchange)"""
    rt = pm.rt
    n0l, N1, N2 = pm.rshape
    rt.sync_stream()
    F = [rt.empty((n0l + 4, N1, N2)) for _ in range(3)]
    C.check(C.lib.vpm_fd4_neg_gradient_slab(rt.ctx, C.i3((n0l + 4, N1, N2)), _h3(pm), 2, _ptr(phi4), _ptr(F[0]),
                                            _ptr(F[1]), _ptr(F[2])))
    return F


def lpt2src_apl(pm, rhok):
    """source of the second-order displacement (fastpm.py:353-387) from ONE inverse FFT: the six
    second derivatives are stencils of the gradient meshes (csrc/stencil.cu).  Returns
    ``(source, state)``; only the potential stays on the tape."""
    if pm.P > 1:
        phi4 = _lpt2_potential_slab(pm, rhok)
        return _lpt2_bilinear(pm, _lpt2_gradients_slab(pm, phi4), None, 0.5), phi4
    phi, F = _lpt2_potential(pm, rhok)
    return _lpt2_bilinear(pm, F, None, 0.5), phi


def lpt2src_vjp(pm, phi, _src):
    rt = pm.rt
    sbar = pm._as_field(_src, False)
    if pm.P > 1:
        n0l, N1, N2 = pm.rshape
        F = _lpt2_gradients_slab(pm, phi)
        rt.sync_stream()
        W = rt.empty((6, n0l + 4, N1, N2))
        C.check(C.lib.vpm_lpt2_adjoint_w_slab(rt.ctx, C.i3(pm.rshape), _h3(pm), 2, _ptr(F[0]), _ptr(F[1]),
                                              _ptr(F[2]), _ptr(sbar.t), _ptr(W)))
        del F
        # the second half differentiates W_00 and W_01 along axis 0: their boundary planes travel
        pm._halo_fill(W[0])
        pm._halo_fill(W[5])
        rt.sync_stream()
        Fb = rt.empty((3, n0l + 4, N1, N2))
        C.check(C.lib.vpm_lpt2_adjoint_f_slab(rt.ctx, C.i3(pm.rshape), _h3(pm), 2, _ptr(W), _ptr(Fb[0]), _ptr(Fb[1]),
                                              _ptr(Fb[2])))
        del W
        phibar = _fd4_adjoint(pm, Fb)
        del Fb
        return pm._transfer(pm._r2c(phibar, 1.0), scale=1.0, laplace=True)
    F = _fd4(pm, phi)
    rt.sync_stream()
    work = rt.empty((6,) + tuple(pm.rshape))
    Fb = [rt.empty(pm.rshape) for _ in range(3)]
    C.check(C.lib.vpm_lpt2_source_adjoint(rt.ctx, C.i3(pm._n3), _h3(pm), _ptr(F[0].t), _ptr(F[1].t),
                                          _ptr(F[2].t), _ptr(sbar.t), _ptr(work), _ptr(Fb[0]), _ptr(Fb[1]),
                                          _ptr(Fb[2])))
    del work, F
    phibar = rt.empty(pm.rshape)
    C.check(C.lib.vpm_fd4_neg_gradient_adjoint(rt.ctx, C.i3(pm._n3), _h3(pm), _ptr(Fb[0]), _ptr(Fb[1]),
                                               _ptr(Fb[2]), _ptr(phibar)))
    del Fb
    return pm._transfer(pm._r2c(RealField(pm, phibar), 1.0), scale=1.0, laplace=True)


def lpt2src_jvp(pm, phi, rhok_):
    if _is_zero_literal(rhok_):
        return 0
    if pm.P > 1:
        F_ = _lpt2_gradients_slab(pm, _lpt2_potential_slab(pm, rhok_))
        return _lpt2_bilinear(pm, _lpt2_gradients_slab(pm, phi), F_, 1.0)
    _, F_ = _lpt2_potential(pm, rhok_)
    return _lpt2_bilinear(pm, _fd4(pm, phi), F_, 1.0)


def force_jvp(pm, st, dx_):
    """forward-mode sweep of force_apl: tangent of dx -> tangents of (f, potk)"""
    rt = pm.rt
    lay, xl = st.layout, st.xl
    if _is_zero_literal(dx_):
        return 0, 0
    xl_ = lay.exchange(asdevice(dx_))
    rho_ = pm.paint_jvp(xl, mass=1.0, v_pos=xl_)
    rhok_ = pm._r2c(rho_, 1.0)
    if st.phi is not None:
        potk_ = pm._transfer(rhok_, scale=st.scale, laplace=True)
        F_ = _fd4(pm, _potential(pm, potk_))
    else:
        potk_, g_ = pm.force_kspace(rhok_, st.scale, st.grad_tables, want_p=True)
        F_ = [pm._c2r(g_[d], 1.0, consume=True) for d in range(3)]
    F = st.meshes(pm)
    cols = [pm._readout_jvp(F[d], xl, F_[d], xl_, None) for d in range(3)]
    return lay.gather(stack(cols, axis=-1)), potk_
