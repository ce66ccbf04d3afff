"""numpy fp64 restatement of the ``pmesh.pm`` protocol used by vmad.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  pmesh itself (C/Cython +
PFFT + MPI, un-vendored, un-pinned: ``/root/reference/.travis.yml:21``) cannot be
run here; every method below names the ``/root/reference/vmad/lib/fastpm.py``
call site it serves.  The formulas are the published pmesh semantics as listed in
SURVEY.md Appendix A; where pmesh's source would be needed to settle a detail the
choice made here is stated, and the CUDA path mirrors it operation for operation.

Conventions fixed here (and mirrored bit-for-bit by vmad_b200/csrc):

* cell coordinate   ``u_d = x_d * (Nmesh_d / BoxSize_d)``  -- one rounded multiply by a
  precomputed fp64 scale, never an FMA with a translate;
* cell index        ``i_d = floor(u_d)``; fractional offset ``f_d = u_d - i_d``; the
  *integer* is wrapped periodically, ``i_d mod Nmesh_d`` (never ``fmod`` of x);
* CIC weights       ``(1 - f_d)`` on ``i_d`` and ``f_d`` on ``i_d + 1``; the 3-D weight is
  ``(w0 * w1) * w2`` and a particle's contribution is ``mass * weight``;
* cell key          ``(i_0 * N_1 + i_1) * N_2 + i_2`` of the wrapped floor cell;
* layout            the stable argsort of the keys (ties keep input order);
  ``exchange`` = take by that permutation, ``gather`` = scatter-add back (the pair
  are mutual adjoints, fastpm.py:281-308);
* FFT               ``r2c = rfftn / Ntot``, ``c2r = irfftn * Ntot`` (fastpm.py:50,64;
  ``c2r(r2c(x)) == x`` is pinned by test_fastpm.py:25-37);
* wavenumbers       integer frequency ``j`` for ``j < N/2`` and ``j - N`` for ``j >= N/2`` on
  every axis (the Nyquist mode is stored as -N/2, also on the half axis).
"""
import numpy

try:  # scipy.fft is threaded; numpy.fft is the fallback
    import scipy.fft as _fft
    _FFT_KW = dict(workers=-1)
except Exception:  # pragma: no cover
    _fft = numpy.fft
    _FFT_KW = {}


def set_fft_workers(n):
    """Number of threads scipy.fft may use (-1: all cores)."""
    if 'workers' in _FFT_KW:
        _FFT_KW['workers'] = n


# ---------------------------------------------------------------------------------------------
# optional C restatement of the CIC loops (oracle/cic.c, built by oracle/Makefile).  Same
# formulas; used when present so the CPU baseline is a compiled loop.  ``set_cic_threads(1)``
# (the default) keeps the strictly sequential particle order; the numpy versions below remain
# the fallback and the cross-check (tests/test_oracle.py).
# ---------------------------------------------------------------------------------------------
import ctypes as _ct
import os as _os

_CIC = None
_CIC_THREADS = 1


def _load_cic():
    global _CIC
    path = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         '_build', 'libcic_oracle.so')
    if _CIC is None and _os.path.exists(path) and not _os.environ.get('ORACLE_NO_C'):
        lib = _ct.CDLL(path)
        P, L, D, I = _ct.c_void_p, _ct.c_int64, _ct.c_double, _ct.c_int
        lib.cic_paint.argtypes = [P, P, D, L, P, P, P, I]
        lib.cic_paint_gradient.argtypes = [P, P, D, P, L, P, P, P, I]
        lib.cic_readout.argtypes = [P, P, L, P, P, P, I]
        lib.cic_readout_gradient.argtypes = [P, P, L, P, P, P, I]
        lib.scatter_add_rows.argtypes = [P, P, L, L, P]
        lib.cic_cell_keys.argtypes = [P, L, P, P, P]
        for f in (lib.cic_paint, lib.cic_paint_gradient, lib.cic_readout,
                  lib.cic_readout_gradient, lib.scatter_add_rows, lib.cic_cell_keys):
            f.restype = None
        _CIC = lib
    return _CIC


def set_cic_threads(n):
    """OpenMP threads for the C loops (1 = sequential, deterministic)."""
    global _CIC_THREADS
    _CIC_THREADS = max(1, int(n))


def use_c(flag):
    """force the numpy (False) or C (True, if built) implementation of the CIC loops"""
    global _CIC
    if flag:
        _os.environ.pop('ORACLE_NO_C', None)
        _load_cic()
    else:
        _os.environ['ORACLE_NO_C'] = '1'
        _CIC = None


def _c3(pm, pos):
    """3-D geometry + contiguous [np, 3] positions for the C loops"""
    pos = numpy.ascontiguousarray(pos, dtype='f8').reshape(-1, pm.ndim)
    pad = 3 - pm.ndim
    if pad:
        pos = numpy.concatenate([numpy.zeros((len(pos), pad)), pos], axis=1)
    n = numpy.array([1] * pad + list(pm.Nmesh), dtype='i8')
    sc = numpy.array([1.0] * pad + list(pm.scale), dtype='f8')
    return pos, n, sc


def _ptr(a):
    return None if a is None else a.ctypes.data


class SelfComm(object):
    """A one-rank stand-in for ``MPI.COMM_SELF`` (fastpm.py:196,251,422,596)."""
    rank = 0
    size = 1

    def allreduce(self, x, op=None):
        return x

    def barrier(self):
        pass

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1


class KList(list):
    """The ``k`` argument handed to ``Field.apply`` callbacks (fastpm.py:141,329)."""

    def normp(self, p=2, zeromode=None):
        r = 0
        for ki in self:
            r = r + numpy.abs(ki) ** p
        r = numpy.array(r, dtype='f8')  # broadcast to the full shape, own copy
        if zeromode is not None:
            r[r == 0] = zeromode
        return r


def _unwrap(x):
    return x.value if isinstance(x, Field) else x


class Field(object):
    """Common behaviour of RealField / ComplexField: an ndarray plus its ``pm``."""
    __array_priority__ = 20.0

    def __init__(self, pm, value):
        self.pm = pm
        self.value = value

    # ndarray-ish surface (fastpm.py:30-31,35,150,176; testing.py:21)
    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return self.value.dtype

    @property
    def size(self):
        return self.value.size

    @property
    def ndim(self):
        return self.value.ndim

    @property
    def real(self):
        return self.value.real

    @property
    def imag(self):
        return self.value.imag

    @property
    def flat(self):
        return self.value.flat

    def __array__(self, dtype=None, copy=None):
        if dtype is None:
            return self.value
        return self.value.astype(dtype)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, index):
        return self.value[index]

    def __setitem__(self, index, v):
        self.value[index] = _unwrap(v)

    def copy(self):
        return type(self)(self.pm, self.value.copy())

    def sum(self, *args, **kwargs):
        return self.value.sum(*args, **kwargs)

    def _wrap(self, v):
        v = numpy.asarray(v)
        if numpy.iscomplexobj(v) and v.shape == self.pm.cshape:
            return ComplexField(self.pm, v)
        if v.shape == self.pm.rshape and not numpy.iscomplexobj(v):
            return RealField(self.pm, v)
        if v.shape == self.value.shape:
            return type(self)(self.pm, v)
        return v

    def __add__(self, o): return self._wrap(self.value + _unwrap(o))
    def __radd__(self, o): return self._wrap(_unwrap(o) + self.value)
    def __sub__(self, o): return self._wrap(self.value - _unwrap(o))
    def __rsub__(self, o): return self._wrap(_unwrap(o) - self.value)
    def __mul__(self, o): return self._wrap(self.value * _unwrap(o))
    def __rmul__(self, o): return self._wrap(_unwrap(o) * self.value)
    def __truediv__(self, o): return self._wrap(self.value / _unwrap(o))
    def __rtruediv__(self, o): return self._wrap(_unwrap(o) / self.value)
    def __pow__(self, o): return self._wrap(self.value ** _unwrap(o))
    def __neg__(self): return self._wrap(-self.value)
    def __pos__(self): return self._wrap(+self.value)
    def __abs__(self): return self._wrap(abs(self.value))
    def __gt__(self, o): return self.value > _unwrap(o)
    def __lt__(self, o): return self.value < _unwrap(o)

    def apply(self, func, kind='wavenumber', out=None):
        """``Field.apply(func(k, v), kind=, out=)`` (fastpm.py:79,83,87,153-154,174,188)."""
        coords = self._coords(kind)
        r = func(coords, self.value)
        if out is Ellipsis:
            self.value[...] = r
            return self
        if out is None:
            res = numpy.empty_like(self.value)
            res[...] = r
            return type(self)(self.pm, res)
        if isinstance(out, Field):
            out.value[...] = r
            return out
        out[...] = r
        return out


class RealField(Field):
    def _coords(self, kind):
        pm = self.pm
        xs = []
        for d in range(pm.ndim):
            shape = [1] * pm.ndim
            shape[d] = pm.Nmesh[d]
            i = numpy.arange(pm.Nmesh[d])
            if kind == 'index':
                x = i
            else:
                x = i * (pm.BoxSize[d] / pm.Nmesh[d])
            xs.append(x.reshape(shape))
        return KList(xs)

    def r2c(self, out=None):
        """fastpm.py:50,56 -- forward transform with pmesh's 1/Ntot normalisation."""
        pm = self.pm
        c = _fft.rfftn(self.value, **_FFT_KW) / pm.Ntot
        return ComplexField(pm, c)

    def c2r_vjp(self, out=None):
        """fastpm.py:67 -- adjoint of c2r:  ``_x_k = Ntot * r2c(_y)_k``.

        Convention for cotangents of Hermitian-compressed complex fields (pinned by
        the reference's own test ``Test_apply_digitized_amp_tf``,
        test_fastpm.py:101-116, whose vjp multiplies by ``_expand_hermitian``,
        fastpm.py:174): a complex cotangent is the gradient with respect to the
        *full* spectrum restricted to the stored modes, i.e.
        ``dL = sum_stored w_k Re(conj(_x_k) dx_k)`` with w = 2 off the
        self-conjugate planes.  Hence no weights appear in the transform adjoints;
        they appear where a reduction over modes is taken (cnorm, cdot,
        apply_digitized)."""
        pm = self.pm
        c = _fft.rfftn(self.value, **_FFT_KW)  # == Ntot * r2c
        return ComplexField(pm, c)

    def cnorm(self):
        """fastpm.py:15 -- sum of squares."""
        return self.pm.comm.allreduce((self.value ** 2).sum())

    def cdot(self, other):
        """fastpm.py:21."""
        return self.pm.comm.allreduce((self.value * _unwrap(other)).sum())

    def cdot_vjp(self, v):
        return self * v

    # -- particle <-> mesh, gather side (fastpm.py:252,256,263) ---------------
    def readout(self, pos, layout=None, resampler=None, out=None):
        pm = self.pm
        if layout is not None:
            pos = layout.exchange(pos)
        value = _cic_readout(pm, self.value, pos)
        if layout is not None:
            value = layout.gather(value)
        return value

    def readout_vjp(self, pos, v, layout=None, resampler=None):
        """returns ``(_mesh, _pos)``:  ``_mesh = paint(pos, mass=v)``,
        ``_pos[p, d] = v_p * sum_c mesh[c] dW/dx_d``."""
        pm = self.pm
        if layout is not None:
            pos1 = layout.exchange(pos)
            v1 = layout.exchange(v)
        else:
            pos1, v1 = pos, v
        _mesh = RealField(pm, _cic_paint(pm, pos1, v1))
        g = _cic_readout_gradient(pm, self.value, pos1)
        _pos = g * numpy.asarray(v1)[:, None]
        if layout is not None:
            _pos = layout.gather(_pos)
        return _mesh, _pos

    def readout_jvp(self, pos, v_self=None, v_pos=None, layout=None, resampler=None):
        pm = self.pm
        if layout is not None:
            pos1 = layout.exchange(pos)
            if v_pos is not None:
                v_pos = layout.exchange(v_pos)
        else:
            pos1 = pos
        r = 0
        if v_self is not None:
            r = r + _cic_readout(pm, _unwrap(v_self), pos1)
        if v_pos is not None:
            g = _cic_readout_gradient(pm, self.value, pos1)
            r = r + (g * numpy.asarray(v_pos)).sum(axis=-1)
        if layout is not None and not numpy.isscalar(r):
            r = layout.gather(r)
        return r


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

    def c2r(self, out=None):
        """fastpm.py:64,70 -- un-normalised inverse transform."""
        pm = self.pm
        r = _fft.irfftn(self.value, s=tuple(pm.Nmesh), **_FFT_KW) * pm.Ntot
        return RealField(pm, r)

    def r2c_vjp(self, out=None):
        """fastpm.py:53 -- adjoint of r2c:  ``_x = c2r(_y) / Ntot`` (cotangent
        convention: see ``RealField.c2r_vjp``)."""
        pm = self.pm
        r = _fft.irfftn(self.value, s=tuple(pm.Nmesh), **_FFT_KW)  # == c2r / Ntot
        return RealField(pm, r)

    def cnorm(self):
        pm = self.pm
        w = pm.hermitian_weights()
        return pm.comm.allreduce(((self.value.real ** 2 + self.value.imag ** 2) * w).sum())

    def cdot(self, other):
        """fastpm.py:519,527 -- sum over the *full* spectrum of x conj(y)."""
        pm = self.pm
        w = pm.hermitian_weights()
        return pm.comm.allreduce((self.value * numpy.conj(_unwrap(other)) * w).sum())

    def cdot_vjp(self, v):
        """fastpm.py:522-523 -- gradient of ``Re cdot(x, self)`` w.r.t. ``x`` (full-
        spectrum cotangent convention, so no Hermitian weight)."""
        return ComplexField(self.pm, self.value * v)

    def _expand_hermitian(self, i, v):
        """fastpm.py:174,188 -- weight each stored mode by how many modes of the
        full spectrum it stands for."""
        N = self.pm.Nmesh[-1]
        mask = (i[-1] != 0) & (numpy.abs(i[-1]) != N // 2)
        return v * (1 + mask)


class Layout(object):
    """Result of ``pm.decompose`` (fastpm.py:272): particle routing.

    One rank: every particle has exactly one copy, on rank 0.  The exchange order
    is the stable cell-key sort (``indices``), so exchanged particles arrive
    cell-sorted; ``gather`` sums copies back into the original order.
    """

    def __init__(self, keys, indices, oldlength):
        self.keys = keys            # int64 cell key of every input particle
        self.indices = indices      # int64 permutation: slot -> input particle
        self.oldlength = oldlength
        self.newlength = len(indices)

    def exchange(self, data):
        """fastpm.py:289,301,308."""
        data = numpy.asarray(data)
        return numpy.take(data, self.indices, axis=0)

    def gather(self, data, mode='sum'):
        """fastpm.py:286,293,304."""
        data = numpy.asarray(data)
        out = numpy.zeros((self.oldlength,) + data.shape[1:], dtype=data.dtype)
        lib = _load_cic()
        if lib is not None and data.dtype == numpy.dtype('f8'):
            d = numpy.ascontiguousarray(data)
            idx = numpy.ascontiguousarray(self.indices, dtype='i8')
            ncomp = int(numpy.prod(data.shape[1:])) if data.ndim > 1 else 1
            lib.scatter_add_rows(_ptr(d), _ptr(idx), len(d), ncomp, _ptr(out))
            return out
        numpy.add.at(out, self.indices, data)
        return out


class ParticleMesh(object):
    """``ParticleMesh(Nmesh, BoxSize, comm, dtype='f8', resampler='cic')``
    (fastpm.py:542-546; tests test_fastpm.py:28).  Single rank."""

    def __init__(self, Nmesh, BoxSize, comm=None, dtype='f8', resampler='cic', np=None):
        Nmesh = numpy.array(Nmesh, dtype='i8').ravel()
        ndim = len(Nmesh)
        if ndim not in (2, 3):
            raise ValueError("only 2-D and 3-D meshes")
        box = numpy.empty(ndim, dtype='f8')
        box[...] = BoxSize
        if numpy.dtype(dtype) != numpy.dtype('f8'):
            raise ValueError("the oracle is fp64 only")
        if resampler != 'cic':
            raise ValueError("only the CIC resampler is restated")
        self.Nmesh = Nmesh
        self.BoxSize = box
        self.ndim = ndim
        self.comm = comm if comm is not None else SelfComm()
        if getattr(self.comm, 'size', 1) != 1 and hasattr(self.comm, 'Get_size') and self.comm.Get_size() != 1:
            raise ValueError("the oracle ParticleMesh is single-rank")
        self.dtype = numpy.dtype(dtype)
        self.resampler = resampler
        self.rshape = tuple(int(n) for n in Nmesh)
        self.cshape = tuple(int(n) for n in Nmesh[:-1]) + (int(Nmesh[-1]) // 2 + 1,)
        self.Ntot = int(Nmesh.prod())
        # the one rounded fp64 scale used for every position -> cell conversion
        self.scale = self.Nmesh / self.BoxSize

    def __vmad_hyper_key__(self):
        """memo key when the repo's VM memoises autooperator models (immutable configuration)"""
        return ('oracle-pm', id(self))

    # -- helpers -------------------------------------------------------------
    def freq_index(self, d):
        """Integer frequency along axis ``d`` of the stored complex array."""
        N = int(self.Nmesh[d])
        j = numpy.arange(self.cshape[d], dtype='i8')
        j[j >= (N + 1) // 2] -= N  # even N: Nyquist stored as -N/2; odd N: no Nyquist
        return j

    def hermitian_weights(self):
        """1 on the self-conjugate planes of the half axis (0 and Nyquist), else 2;
        broadcastable against the complex array."""
        N = int(self.Nmesh[-1])
        j = numpy.arange(self.cshape[-1])
        w = numpy.where((j == 0) | ((N % 2 == 0) & (j == N // 2)), 1.0, 2.0)
        shape = [1] * self.ndim
        shape[-1] = len(j)
        return w.reshape(shape)

    # -- factories (fastpm.py:29,39,52,55,66,69,228,262,532) -------------------
    def create(self, type=None, value=None, mode=None):
        type = type if type is not None else mode
        if type == 'real':
            f = RealField(self, numpy.zeros(self.rshape, dtype='f8'))
        elif type == 'complex':
            f = ComplexField(self, numpy.zeros(self.cshape, dtype='c16'))
        else:
            raise ValueError("type must be 'real' or 'complex'")
        if value is not None:
            f.value[...] = _unwrap(value)
        return f

    def generate_uniform_particle_grid(self, shift=0.5, dtype='f8'):
        """``(i + shift) * BoxSize / Nmesh`` for every cell, C order, ``[Np, ndim]``."""
        cell = self.BoxSize / self.Nmesh
        axes = [(numpy.arange(n) + shift) * cell[d] for d, n in enumerate(self.rshape)]
        grid = numpy.meshgrid(*axes, indexing='ij')
        return numpy.stack([g.ravel() for g in grid], axis=-1).astype(dtype)

    def generate_whitenoise(self, seed, unitary=False, type='complex', mean=0.0, mode=None):
        """Seeded white noise.  pmesh's Gadget-compatible generator is not
        reproducible without its source; the stand-in is the
        transform of a ``numpy.random.default_rng(seed)`` standard-normal real
        field, scaled to unit variance per complex mode."""
        type = mode if mode is not None else type
        rng = numpy.random.default_rng(seed)
        g = rng.standard_normal(self.rshape)
        c = _fft.rfftn(g, **_FFT_KW) / self.Ntot ** 0.5
        if unitary:
            a = numpy.abs(c)
            a[a == 0] = 1.0
            c = c / a
        c[(0,) * self.ndim] = mean
        cf = ComplexField(self, c)
        if type == 'real':
            return cf.c2r()
        return cf

    # -- domain decomposition (fastpm.py:272) ---------------------------------
    def cell_index(self, pos):
        """Floor cell ``i`` (unwrapped int64) and fractional offset ``f`` of every
        particle, from the *same* ``u``."""
        pos = numpy.asarray(pos, dtype='f8')
        u = pos * self.scale  # one rounded multiply per coordinate
        fl = numpy.floor(u)
        return fl.astype('i8'), u - fl

    def cell_keys(self, pos):
        lib = _load_cic()
        if lib is not None:
            p3, n, sc = _c3(self, pos)
            keys = numpy.empty(len(p3), dtype='i8')
            lib.cic_cell_keys(_ptr(p3), len(p3), _ptr(n), _ptr(sc), _ptr(keys))
            return keys
        i, _ = self.cell_index(pos)
        key = numpy.zeros(len(i), dtype='i8')
        for d in range(self.ndim):
            key = key * self.Nmesh[d] + numpy.mod(i[:, d], self.Nmesh[d])
        return key

    def decompose(self, pos, smoothing=None):
        keys = self.cell_keys(pos)
        indices = numpy.argsort(keys, kind='stable')
        return Layout(keys, indices, len(keys))

    # -- particle <-> mesh, scatter side (fastpm.py:224,229,238) --------------
    def paint(self, pos, mass=1.0, layout=None, hold=False, out=None, resampler=None):
        if layout is not None:
            pos = layout.exchange(pos)
            if numpy.ndim(mass) != 0:
                mass = layout.exchange(mass)
        return RealField(self, _cic_paint(self, pos, mass))

    def paint_vjp(self, v, pos, layout=None, mass=1.0, resampler=None):
        """returns ``(_pos, _mass)``:  ``_mass_p = readout(v, pos)``,
        ``_pos[p, d] = mass_p * sum_c v[c] dW/dx_d``."""
        v = _unwrap(v)
        if layout is not None:
            pos1 = layout.exchange(pos)
            mass1 = layout.exchange(mass) if numpy.ndim(mass) != 0 else mass
        else:
            pos1, mass1 = pos, mass
        _mass = _cic_readout(self, v, pos1)
        g = _cic_readout_gradient(self, v, pos1)
        if numpy.ndim(mass1) != 0:
            _pos = g * numpy.asarray(mass1)[:, None]
        else:
            _pos = g * mass1
        if layout is not None:
            _pos = layout.gather(_pos)
            _mass = layout.gather(_mass)
        return _pos, _mass

    def paint_jvp(self, pos, v_mass=None, mass=1.0, v_pos=None, layout=None, resampler=None):
        """``mesh_ = sum_d paint(pos, mass * v_pos[:, d], gradient=d) + paint(pos, v_mass)``."""
        if layout is not None:
            pos = layout.exchange(pos)
            if numpy.ndim(mass) != 0:
                mass = layout.exchange(mass)
            if v_pos is not None:
                v_pos = layout.exchange(v_pos)
            if v_mass is not None and numpy.ndim(v_mass) != 0:
                v_mass = layout.exchange(v_mass)
        out = numpy.zeros(self.rshape, dtype='f8')
        if v_pos is not None:
            out += _cic_paint_gradient(self, pos, mass, numpy.asarray(v_pos))
        if v_mass is not None:
            out += _cic_paint(self, pos, v_mass)
        return RealField(self, out)


# ---------------------------------------------------------------------------
# CIC window (pmesh ``ResampleWindow`` 'cic': linear kernel, support 2).
# Written as eight vectorised corner passes; accumulation is in particle order
# inside each corner (numpy.bincount), corners in (a, b, c) lexicographic order.
# ---------------------------------------------------------------------------
def _corner_iter(pm, pos):
    """Yield ``(flat_index, weight, dweights)`` for each of the 2^ndim corners.

    ``dweights[d]`` is dW/dx_d (including the ``Nmesh_d / BoxSize_d`` chain factor).
    """
    i, f = pm.cell_index(pos)
    nd = pm.ndim
    N = pm.Nmesh
    w = [(1.0 - f[:, d], f[:, d]) for d in range(nd)]
    dw = [(-pm.scale[d], pm.scale[d]) for d in range(nd)]
    for corner in numpy.ndindex(*([2] * nd)):
        flat = numpy.zeros(len(i), dtype='i8')
        for d in range(nd):
            flat = flat * N[d] + numpy.mod(i[:, d] + corner[d], N[d])
        wd = [w[d][corner[d]] for d in range(nd)]
        weight = wd[0] * wd[1]
        if nd == 3:
            weight = weight * wd[2]
        grads = []
        for d in range(nd):
            g = None
            for e in range(nd):
                t = dw[d][corner[d]] if e == d else wd[e]
                g = t if g is None else g * t
            grads.append(g)
        yield flat, weight, grads


def _cic_paint(pm, pos, mass):
    lib = _load_cic()
    if lib is not None:
        p3, n, sc = _c3(pm, pos)
        marr = numpy.ascontiguousarray(mass, dtype='f8') if numpy.ndim(mass) != 0 else None
        out = numpy.zeros(pm.Ntot, dtype='f8')
        lib.cic_paint(_ptr(p3), _ptr(marr), 1.0 if marr is not None else float(mass), len(p3),
                      _ptr(n), _ptr(sc), _ptr(out), _CIC_THREADS)
        return out.reshape(pm.rshape)
    pos = numpy.asarray(pos, dtype='f8').reshape(-1, pm.ndim)
    mass = numpy.asarray(mass, dtype='f8') if numpy.ndim(mass) != 0 else float(mass)
    out = numpy.zeros(pm.Ntot, dtype='f8')
    for flat, weight, _ in _corner_iter(pm, pos):
        out += numpy.bincount(flat, weights=mass * weight, minlength=pm.Ntot)
    return out.reshape(pm.rshape)


def _cic_paint_gradient(pm, pos, mass, v_pos):
    lib = _load_cic()
    if lib is not None:
        p3, n, sc = _c3(pm, pos)
        v3 = _c3(pm, v_pos)[0]
        marr = numpy.ascontiguousarray(mass, dtype='f8') if numpy.ndim(mass) != 0 else None
        out = numpy.zeros(pm.Ntot, dtype='f8')
        lib.cic_paint_gradient(_ptr(p3), _ptr(marr), 1.0 if marr is not None else float(mass),
                               _ptr(v3), len(p3), _ptr(n), _ptr(sc), _ptr(out), _CIC_THREADS)
        return out.reshape(pm.rshape)
    pos = numpy.asarray(pos, dtype='f8').reshape(-1, pm.ndim)
    mass = numpy.asarray(mass, dtype='f8') if numpy.ndim(mass) != 0 else float(mass)
    out = numpy.zeros(pm.Ntot, dtype='f8')
    for flat, _, grads in _corner_iter(pm, pos):
        c = 0
        for d in range(pm.ndim):
            c = c + v_pos[:, d] * grads[d]
        out += numpy.bincount(flat, weights=mass * c, minlength=pm.Ntot)
    return out.reshape(pm.rshape)


def _cic_readout(pm, mesh, pos):
    lib = _load_cic()
    if lib is not None:
        p3, n, sc = _c3(pm, pos)
        m = numpy.ascontiguousarray(mesh, dtype='f8')
        out = numpy.empty(len(p3), dtype='f8')
        lib.cic_readout(_ptr(m), _ptr(p3), len(p3), _ptr(n), _ptr(sc), _ptr(out), _CIC_THREADS)
        return out
    pos = numpy.asarray(pos, dtype='f8').reshape(-1, pm.ndim)
    m = numpy.asarray(mesh, dtype='f8').ravel()
    out = numpy.zeros(len(pos), dtype='f8')
    for flat, weight, _ in _corner_iter(pm, pos):
        out += m[flat] * weight
    return out


def _cic_readout_gradient(pm, mesh, pos):
    lib = _load_cic()
    if lib is not None:
        p3, n, sc = _c3(pm, pos)
        m = numpy.ascontiguousarray(mesh, dtype='f8')
        out = numpy.empty((len(p3), 3), dtype='f8')
        lib.cic_readout_gradient(_ptr(m), _ptr(p3), len(p3), _ptr(n), _ptr(sc), _ptr(out),
                                 _CIC_THREADS)
        return out[:, 3 - pm.ndim:]
    pos = numpy.asarray(pos, dtype='f8').reshape(-1, pm.ndim)
    m = numpy.asarray(mesh, dtype='f8').ravel()
    out = numpy.zeros((len(pos), pm.ndim), dtype='f8')
    for flat, _, grads in _corner_iter(pm, pos):
        mv = m[flat]
        for d in range(pm.ndim):
            out[:, d] += mv * grads[d]
    return out
