"""FastPM operator library: the names, signatures and semantics of ``vmad.lib.fastpm``.

Every operator here answers to the same name, argument order, ``ain`` / ``aout`` and vjp / jvp
conventions as its counterpart in ``/root/reference/vmad/lib/fastpm.py`` (cited per operator),
so model / autooperator graphs written against the reference library run unchanged.  The
operators speak the pmesh object protocol and therefore run on any
backend that implements it; on ``vmad_b200.pm`` (the B200 backend) the k-space operators take
fused single-pass kernels because the library's own filters (``fourier_space_laplace``,
``fourier_space_neg_gradient``) are recognisable objects rather than anonymous closures.

Additions named by BASELINE.json ``north_star`` (absent from the reference):
the fused operators ``force``, ``kick`` and ``drift`` -- see the end of this file.
"""
import itertools

import numpy

from ..background import MatterDominated
from ..vm import autooperator, operator
from ..vm.ops import ZeroGradient
from . import linalg


def __getattr__(name):
    """``fastpm.ParticleMesh`` (the reference re-exports pmesh's, fastpm.py:6), resolved on first
    use: importing the operator library must not load torch / libvmadpm.so (the library also runs
    on other backends of the pmesh protocol, and bench.py's CPU arm must not map the CUDA library)"""
    if name == 'ParticleMesh':
        from ..pm import ParticleMesh
        return ParticleMesh
    raise AttributeError(name)


def _is_zero(x):
    return ZeroGradient.is_zero(x)


def _localize(q, pm):
    """particle arrays enter the graph once: put them where the mesh lives"""
    f = getattr(pm, 'asparticles', None)
    return f(q) if f is not None else q


# ---------------------------------------------------------------------------------------------
# field <-> scalar, entry points
# ---------------------------------------------------------------------------------------------
@operator
class to_scalar:
    """fastpm.py:9-21 -- ``y = sum x^2`` (``Field.cnorm``)"""
    ain = {'x': 'RealField'}
    aout = {'y': '*'}

    def apl(node, x):
        return dict(y=x.cnorm())

    def vjp(node, _y, x):
        return dict(_x=x * (2 * _y))

    def jvp(node, x_, x):
        return dict(y_=x.cdot(x_) * 2)


@operator
class as_complex_field:
    """fastpm.py:23-42 -- ``[..., 2]`` real pairs -> ComplexField"""
    ain = {'x': '*'}
    aout = {'y': 'ComplexField'}

    def apl(node, x, pm):
        x = numpy.asarray(x)
        return dict(y=pm.create(type='complex', value=x[..., 0] + 1j * x[..., 1]))

    def vjp(node, _y, pm):
        _y = numpy.asarray(_y)
        return dict(_x=numpy.stack([_y.real, _y.imag], axis=-1))

    def jvp(node, x_, pm):
        x_ = numpy.asarray(x_)
        return dict(y_=pm.create(type='complex', value=x_[..., 0] + 1j * x_[..., 1]))


# ---------------------------------------------------------------------------------------------
# transforms
# ---------------------------------------------------------------------------------------------
@operator
class r2c:
    """fastpm.py:44-56"""
    ain = {'x': 'RealField'}
    aout = {'y': 'ComplexField'}

    def apl(node, x):
        return dict(y=x.r2c(), pm=x.pm)

    def vjp(node, _y, pm):
        return dict(_x=pm.create(type='complex', value=_y).r2c_vjp())

    def jvp(node, x_, pm):
        return dict(y_=pm.create(type='real', value=x_).r2c())


@operator
class c2r:
    """fastpm.py:58-70"""
    ain = {'x': 'ComplexField'}
    aout = {'y': 'RealField'}

    def apl(node, x):
        return dict(y=x.c2r(), pm=x.pm)

    def vjp(node, _y, pm):
        return dict(_x=pm.create(type='real', value=_y).c2r_vjp())

    def jvp(node, x_, pm):
        return dict(y_=pm.create(type='complex', value=x_).c2r())


# ---------------------------------------------------------------------------------------------
# transfer functions
# ---------------------------------------------------------------------------------------------
class Filter(object):
    """A transfer function of ``k`` that the device backend can apply with a fused kernel.

    Calling the object with the ``k`` list (what ``Field.apply`` hands to its callback) returns
    the multiplier like any plain function would, so it also works on a generic backend.
    """

    def __call__(self, k):
        raise NotImplementedError

    def apply_device(self, x, conj):
        """``x * self(k)`` (or its conjugate) on a device ComplexField"""
        raise NotImplementedError


class _Laplace(Filter):
    """fastpm.py:328-334 -- the Green's function ``-1 / k^2`` with the zero mode set to 0"""

    def __call__(self, k):
        k2 = k.normp(2)
        bad = k2 == 0
        k2[bad] = 1
        k2 = -1 / k2
        k2[bad] = 0
        return k2

    def apply_device(self, x, conj):
        return x.pm._transfer(x, laplace=True)  # real multiplier: its own conjugate


fourier_space_laplace = _Laplace()


class _NegGradient(Filter):
    """fastpm.py:310-326 -- ``-i k_d`` (order 0, with the Nyquist mask) or the 4-point finite
    difference kernel ``-i (8 sin w - sin 2w) / (6 h)``, ``w = k_d h`` (order 1)"""

    def __init__(self, dir, pm, order):
        if order not in (0, 1):
            raise ValueError("only order 0 and 1 are supported")
        self.dir = dir
        self.order = order
        self.BoxSize = float(pm.BoxSize[dir])
        self.Nmesh = int(pm.Nmesh[dir])

    def __call__(self, k):
        kd = k[self.dir]
        if self.order == 0:
            nyquist = numpy.pi / self.BoxSize * (self.Nmesh - 0.01)
            return -1j * kd * (abs(kd) < nyquist)
        cellsize = self.BoxSize / self.Nmesh
        w = kd * cellsize
        a = 1 / (6.0 * cellsize) * (8 * numpy.sin(w) - numpy.sin(2 * w))
        return -1j * a

    def apply_device(self, x, conj):
        pm = x.pm
        d = self.dir
        key = ('neg_gradient', d, self.order, self.BoxSize, self.Nmesh)
        if key not in pm._tables:
            k1 = [None] * pm.ndim
            k1[d] = pm.freq_index(d) * (2 * numpy.pi / pm.BoxSize[d])
            pm._axis_table(key, d, self(k1))
        tables = [None] * pm.ndim
        tables[d] = pm._tables[key]
        return pm._transfer(x, tables=tables, conj=conj)


def fourier_space_neg_gradient(dir, pm, order):
    return _NegGradient(dir, pm, order)


def _apply_tf(x, tf, kind, conj):
    if isinstance(tf, Filter) and kind == 'wavenumber' and hasattr(x.pm, '_transfer'):
        return tf.apply_device(x, conj)
    cached = getattr(x, 'apply_transfer_function', None)
    if cached is not None:
        return cached(tf, kind, conj)
    if conj:
        return x.apply(lambda k, v: v * numpy.conj(tf(k)), kind=kind)
    return x.apply(lambda k, v: v * tf(k), kind=kind)


@operator
class apply_transfer:
    """fastpm.py:72-87 -- ``y_k = x_k tf(k)``; the vjp multiplies by ``conj(tf)``"""
    ain = {'x': 'ComplexField'}
    aout = {'y': 'ComplexField'}

    def apl(node, x, tf, kind='wavenumber'):
        return dict(y=_apply_tf(x, tf, kind, False))

    def vjp(node, _y, tf, kind):
        return dict(_x=_apply_tf(_y, tf, kind, True))

    def jvp(node, x_, tf, kind):
        return dict(y_=_apply_tf(x_, tf, kind, False))


def _take_default(array, ind, conj, default):
    """``array[ind]`` with out-of-range bins replaced by ``default`` and conjugation where
    flagged (fastpm.py:89-95)"""
    array = numpy.asarray(array)
    ind1 = ind.clip(0, len(array) - 1)
    r = array[ind1]
    if conj.any():
        r = numpy.where(conj, numpy.conj(r), r)
    return numpy.where(ind != ind1, default, r)


def _times_table(x, table):
    """field times a host multiplier table of the field's (local) shape"""
    f = getattr(x.pm, '_transfer_table', None)
    if f is not None:
        return f(x, numpy.ascontiguousarray(table))
    return x * table


@operator
class apply_digitized:
    """fastpm.py:97-213 -- transfer function given per bin: ``y = x * tf[bin(k)]`` (mode
    'amplitude') or ``y = x * exp(i tf[bin(k)])`` (mode 'phase'); modes whose bin falls outside
    ``tf`` are left alone.  ``digitizer(k) -> (ind, conj)`` is user code and is evaluated on the
    host (like any ``tf(k)``); on the device backend the per-mode multiply and the binned
    reduction of the vjp are kernels, the per-bin table ``tf`` stays a (small) host array."""
    ain = 'x', 'tf'
    aout = 'y'

    @staticmethod
    def isotropic_wavenumber(kedges):
        def digitizer(k):
            k2 = k.normp(2)
            ind = numpy.digitize(k2 ** 0.5, kedges) - 1
            return ind, numpy.zeros_like(ind, dtype='?')
        return digitizer

    def apl(node, x, tf, digitizer, kind='wavenumber', mode='amplitude'):
        assert mode in ('amplitude', 'phase')
        shape = tuple(x.shape)
        ind, conj = digitizer(x._coords(kind))
        ind = numpy.broadcast_to(numpy.asarray(ind, dtype='intp'), shape)
        conj = numpy.broadcast_to(numpy.asarray(conj, dtype='?'), shape)
        tf = numpy.asarray(tf)
        eit = None if mode == 'amplitude' else numpy.exp(1j * tf)
        dev = getattr(x.pm, 'digitized', None)
        if dev is not None:
            # device backend: the bins (like any tf(k), evaluated once by the user's digitizer on
            # the host) become device index arrays; the multiply and, in the vjp, the binned
            # reduction are kernels (vpm_digitized_mul / vpm_digitized_bin)
            ind, conj = dev.bins(ind, conj)
            y = dev.mul(x, ind, conj, tf if eit is None else eit, 1.0)
            return dict(y=y, ind=ind, conj=conj, eit=eit)
        y = _times_table(x, _take_default(tf if eit is None else eit, ind, conj, 1))
        return dict(y=y, ind=ind, conj=conj, eit=eit)

    def vjp(node, _y, x, tf, ind, conj, eit, mode):
        tf = numpy.asarray(tf)
        pm = x.pm
        _y = pm.create(type='complex', value=_y)
        dev = getattr(pm, 'digitized', None)
        if dev is not None:
            _x = dev.mul(_y, ind, conj, tf if eit is None else eit, 1.0, conj_table=eit is not None)
            _tf = pm.comm.allreduce(dev.bin(x, _y, ind, conj, len(tf), eit))
            return dict(_tf=_tf, _x=_x)
        ind1 = ind.clip(0, len(tf) - 1)
        mask = ind1 != ind
        # every stored mode stands for 1 or 2 modes of the full spectrum (fastpm.py:174,188)
        i_last = x._coords('index')[-1]
        mult = 1 + ((i_last != 0) & (numpy.abs(i_last) != int(pm.Nmesh[-1]) // 2))
        xh = numpy.asarray(x)
        yh = numpy.asarray(_y)
        if mode == 'amplitude':
            weights = yh * numpy.conj(xh) * mult
            _x = _times_table(_y, _take_default(tf, ind, conj, 1))
        else:
            y = xh * _take_default(eit, ind, conj, 0)
            weights = yh * numpy.conj(1j * y) * mult
            _x = _times_table(_y, _take_default(eit.conj(), ind, conj, 1))
        weights = numpy.where(mask, 0, weights)
        _tf = numpy.bincount(ind1.ravel(), weights=weights.real.ravel(), minlength=len(tf))
        _tf = pm.comm.allreduce(_tf)
        return dict(_tf=_tf, _x=_x)

    def jvp(node, tf_, x_, x, tf, ind, conj, eit, mode):
        y_ = 0
        tf = numpy.asarray(tf)
        dev = getattr(x.pm, 'digitized', None)
        if dev is not None:
            times = lambda f, table, dflt: dev.mul(f, ind, conj, table, dflt)
        else:
            times = lambda f, table, dflt: _times_table(f, _take_default(table, ind, conj, dflt))
        if mode == 'amplitude':
            if not _is_zero(tf_):
                y_ = y_ + times(x, numpy.asarray(tf_), 0)
            if not _is_zero(x_):
                y_ = y_ + times(x.pm.create(type='complex', value=x_), tf, 1)
        else:
            if not _is_zero(tf_):
                y_ = y_ + times(x, eit * 1j * numpy.asarray(tf_), 0)
            if not _is_zero(x_):
                y_ = y_ + times(x.pm.create(type='complex', value=x_), eit, 1)
        return dict(y_=y_)


# ---------------------------------------------------------------------------------------------
# particles <-> mesh
# ---------------------------------------------------------------------------------------------
@operator
class paint:
    """fastpm.py:215-240 -- CIC deposit of ``mass`` at ``x``"""
    aout = {'mesh': 'RealField'}
    ain = {'x': 'ndarray', 'layout': 'Layout', 'mass': 'ndarray'}

    def apl(node, x, mass, layout, pm):
        return dict(mesh=pm.paint(x, mass=mass, layout=layout, hold=False))

    def vjp(node, _mesh, x, mass, layout, pm):
        _mesh = pm.create(type='real', value=_mesh)
        _x, _mass = pm.paint_vjp(_mesh, x, layout=layout, mass=mass)
        return dict(_layout=0, _x=_x, _mass=_mass)

    def jvp(node, x_, x, layout, mass, layout_, mass_, pm):
        if _is_zero(x_):
            x_ = None
        if _is_zero(mass_):
            mass_ = None
        return dict(mesh_=pm.paint_jvp(x, v_mass=mass_, mass=mass, v_pos=x_, layout=layout))


@operator
class readout:
    """fastpm.py:242-264 -- CIC interpolation of ``mesh`` at ``x`` (the reference's host-side
    ``allreduce(len(x))`` at :251 is dropped: its result is unused and it is a sync point)"""
    aout = {'value': 'ndarray'}
    ain = {'x': 'ndarray', 'mesh': 'RealField', 'layout': 'Layout'}

    def apl(node, mesh, x, layout, resampler=None):
        return dict(value=mesh.readout(x, layout=layout, resampler=resampler), pm=mesh.pm)

    def vjp(node, _value, x, layout, mesh, pm, resampler=None):
        _mesh, _x = mesh.readout_vjp(x, _value, layout=layout, resampler=resampler)
        return dict(_mesh=_mesh, _x=_x, _layout=0)

    def jvp(node, x_, mesh_, x, layout, layout_, mesh, pm, resampler=None):
        if _is_zero(mesh_):
            mesh_ = None
        if _is_zero(x_):
            x_ = None
        mesh = pm.create(type='real', value=mesh)
        return dict(value_=mesh.readout_jvp(x, v_self=mesh_, v_pos=x_, layout=layout,
                                            resampler=resampler))


@operator
class decompose:
    """fastpm.py:266-278 -- particle routing; carries no gradient"""
    aout = {'layout': 'Layout'}
    ain = {'x': 'ndarray'}

    def apl(node, x, pm):
        return dict(layout=pm.decompose(x))

    def vjp(node, _layout):
        return dict(_x=0)

    def jvp(node, x_):
        return dict(layout_=0)


@operator
class gather:
    """fastpm.py:280-293 -- return exchanged values to their home order, summing copies"""
    aout = 'y'
    ain = 'x', 'layout'

    def apl(node, x, layout):
        return dict(y=layout.gather(x))

    def vjp(node, _y, layout):
        return dict(_layout=0, _x=layout.exchange(_y))

    def jvp(node, layout, layout_, x_):
        return dict(y_=layout.gather(x_))


@operator
class exchange:
    """fastpm.py:295-308 -- send values to the ranks (and the order) that own their cells"""
    aout = 'y'
    ain = 'x', 'layout'

    def apl(node, x, layout):
        return dict(y=layout.exchange(x))

    def vjp(node, _y, layout):
        return dict(_layout=0, _x=layout.gather(_y))

    def jvp(node, layout, layout_, x_):
        return dict(y_=layout.exchange(x_))


# ---------------------------------------------------------------------------------------------
# Lagrangian perturbation theory
# ---------------------------------------------------------------------------------------------
@autooperator('rhok,q->dx1')
def lpt1(rhok, q, pm):
    """fastpm.py:336-351 -- Zel'dovich displacement at ``q``; note that ``q`` is a
    differentiable input of this operator"""
    p = apply_transfer(rhok, fourier_space_laplace)
    layout = decompose(q, pm)
    r1 = []
    for d in range(pm.ndim):
        dx1_c = apply_transfer(p, fourier_space_neg_gradient(d, pm, order=1))
        dx1_r = c2r(dx1_c)
        r1.append(readout(dx1_r, q, layout))
    return dict(dx1=linalg.stack(r1, axis=-1))


@autooperator('rhok->rho_lpt2')
def lpt2src(rhok, pm):
    """fastpm.py:353-387 -- source of the second-order displacement,
    ``3/7 (sum_{i<j} P_ii P_jj - P_ij^2)`` with ``P_ij`` the second derivatives of the potential"""
    if pm.ndim != 3:
        raise ValueError("LPT 2 only works in 3d")
    D1 = [1, 2, 0]
    D2 = [2, 0, 1]
    potk = apply_transfer(rhok, fourier_space_laplace)

    def second_derivative(i, j):
        t = apply_transfer(potk, fourier_space_neg_gradient(i, pm, order=1))
        t = apply_transfer(t, fourier_space_neg_gradient(j, pm, order=1))
        return c2r(t)

    Pii = [second_derivative(d, d) for d in range(pm.ndim)]
    source = None
    for d in range(pm.ndim):
        term = Pii[D1[d]] * Pii[D2[d]]
        source = term if source is None else source + term
    for d in range(pm.ndim):
        Pij = second_derivative(D1[d], D2[d])
        source = source + (-Pij * Pij)
    return dict(rho_lpt2=(3.0 / 7) * source)


@autooperator('wnk->c')
def induce_correlation(wnk, powerspectrum, pm):
    """fastpm.py:389-396 -- colour white noise with ``sqrt(P(k) / V)``"""
    volume = pm.BoxSize.prod()

    def tf(k):
        kk = sum(ki ** 2 for ki in k) ** 0.5
        return (powerspectrum(kk) / volume) ** 0.5

    return dict(c=apply_transfer(wnk, tf))


@operator
class lpt1_at_constant_q:
    """``lpt1`` (fastpm.py:336-351) for the case every caller in the library has -- ``q`` a
    constant of the graph -- as ONE operator: Green's function, ONE inverse FFT of the potential,
    the real-space 4-point stencil that ``fourier_space_neg_gradient(order=1)`` is the symbol of,
    and the three-mesh readout, instead of 4 transfers + 3 inverse FFTs + 3 readouts + stack
    (13 nodes).  Device backend (across ranks the stencil reads the neighbours' boundary planes);
    ``lpt1`` is the portable graph."""
    ain = {'rhok': 'ComplexField'}
    aout = {'dx1': 'ndarray'}

    def apl(node, rhok, q, pm):
        from .. import pm as dev
        dx1, state = dev.lpt1_apl(pm, rhok, q)
        return dict(dx1=dx1, state=state)

    def vjp(node, _dx1, state, pm):
        from .. import pm as dev
        return dict(_rhok=dev.lpt1_vjp(pm, state, _dx1))

    def jvp(node, rhok_, state, pm):
        from .. import pm as dev
        return dict(dx1_=dev.lpt1_jvp(pm, state, rhok_))


@operator
class lpt2src_real_space:
    """``lpt2src`` (fastpm.py:353-387) as ONE operator: Green's function, ONE inverse FFT, the
    gradient stencil, and a second stencil pass that forms the six second derivatives and their
    3/7 (sum P_ii P_jj - P_ij^2) combination -- instead of 13 transfers, 6 inverse FFTs and 9
    elementwise nodes; the reverse sweep is two stencil kernels, the gradient-stencil adjoint and
    ONE forward FFT.  Device backend (across ranks: four halo planes of the potential, then halo
    planes of two of the six adjoint meshes); ``lpt2src`` is the portable graph."""
    ain = {'rhok': 'ComplexField'}
    aout = {'rho_lpt2': 'RealField'}

    def apl(node, rhok, pm):
        from .. import pm as dev
        src, phi = dev.lpt2src_apl(pm, rhok)
        return dict(rho_lpt2=src, phi=phi)

    def vjp(node, _rho_lpt2, phi, pm):
        from .. import pm as dev
        return dict(_rhok=dev.lpt2src_vjp(pm, phi, _rho_lpt2))

    def jvp(node, rhok_, phi, pm):
        from .. import pm as dev
        return dict(rho_lpt2_=dev.lpt2src_jvp(pm, phi, rhok_))


@autooperator('rhok->dx1,dx2')
def lpt(rhok, q, pm):
    """fastpm.py:398-406"""
    q = _localize(q, pm)
    one, src = lpt1, lpt2src
    if _fusable(pm) and USE_FUSED_STAGE:
        from .. import pm as dev
        if dev._use_stencil(pm):
            one = lpt1_at_constant_q           # across ranks: halo planes over peer memory
            if getattr(pm, 'P', 1) == 1 or pm.rshape[0] >= 4:
                src = lpt2src_real_space       # two stencils deep: four halo planes across ranks
    dx1 = one(rhok, q, pm)
    rhok2 = r2c(src(rhok, pm))
    dx2 = one(rhok2, q, pm)
    return dict(dx1=dx1, dx2=dx2)


# ---------------------------------------------------------------------------------------------
# force, leapfrog, n-body
# ---------------------------------------------------------------------------------------------
def _gravity_body(dx, q, pm):
    """shared by ``gravity`` (fastpm.py:409-438) and ``FastPMSimulation.gravity`` (:582-612)"""
    q = _localize(q, pm)
    x = q + dx
    layout = decompose(x, pm)
    xl = exchange(x, layout)
    rho = paint(xl, 1.0, None, pm)
    # mean density 1 (the particle count is reduced once, when the graph is built)
    fac = 1.0 * pm.Nmesh.prod() / pm.comm.allreduce(len(q))
    rho = rho * fac
    rhok = r2c(rho)
    p = apply_transfer(rhok, fourier_space_laplace)
    r1 = []
    for d in range(pm.ndim):
        dx1_c = apply_transfer(p, fourier_space_neg_gradient(d, pm, order=1))
        dx1_r = c2r(dx1_c)
        dx1l = readout(dx1_r, xl, None)
        r1.append(gather(dx1l, layout))
    return dict(f=linalg.stack(r1, axis=-1), potk=p)


@autooperator('dx->f,potk')
def gravity(dx, q, pm):
    return _gravity_body(dx, q, pm)


def KickFactor(pt, ai, ac, af):
    """fastpm.py:440-441"""
    return 1 / (ac ** 2 * pt.E(ac)) * (pt.Gf(af) - pt.Gf(ai)) / pt.gf(ac)


def DriftFactor(pt, ai, ac, af):
    """fastpm.py:443-444"""
    return 1 / (ac ** 3 * pt.E(ac)) * (pt.Gp(af) - pt.Gp(ai)) / pt.gp(ac)


def _leapfrog_body(dx, p, stages, pt, force_fn, fused=False, stage_fn=None):
    """kick-drift-kick stepping of fastpm.py:455-473 around a force callback.

    ``fused`` (device backend) uses the single-pass kick / drift operators and merges the
    closing half-kick of a step with the opening half-kick of the next one -- both add the SAME
    force to ``p`` with nothing in between, so ``p + f c1 + f c2 = p + f (c1 + c2)`` (one pass
    over the particles instead of two; equal to rounding).  ``stage_fn(dx, p, cd, ck)``: one
    fused drift-force-kick stage instead of the three operators."""
    Om0 = pt.Om0
    pairs = list(zip(stages[:-1], stages[1:]))
    f, potk = force_fn(dx)
    if not fused:
        for ai, af in pairs:
            ac = (ai * af) ** 0.5
            p = p + f * (KickFactor(pt, ai, ai, ac) * 1.5 * Om0)       # kick
            dx = dx + p * DriftFactor(pt, ai, ac, af)                  # drift
            f, potk = force_fn(dx)                                     # force
            p = p + f * (KickFactor(pt, ac, af, af) * 1.5 * Om0)       # kick
        return dict(dx=dx, p=p, f=f * (1.5 * Om0))
    opening = [KickFactor(pt, ai, ai, (ai * af) ** 0.5) * 1.5 * Om0 for ai, af in pairs]
    closing = [KickFactor(pt, (ai * af) ** 0.5, af, af) * 1.5 * Om0 for ai, af in pairs]
    if pairs:
        p = kick(p, f, opening[0])
    for i, (ai, af) in enumerate(pairs):
        ac = (ai * af) ** 0.5
        c = closing[i] + (opening[i + 1] if i + 1 < len(pairs) else 0.0)
        if stage_fn is not None:
            dx, p, f, potk = stage_fn(dx, p, DriftFactor(pt, ai, ac, af), c, i + 1 == len(pairs))
            continue
        dx = drift(dx, p, DriftFactor(pt, ai, ac, af))
        f, potk = force_fn(dx)
        p = kick(p, f, c)
    return dict(dx=dx, p=p, f=f * (1.5 * Om0))


@autooperator('dx_i,p_i->dx,p,f')
def leapfrog(dx_i, p_i, q, stages, pt, pm):
    """fastpm.py:446-474"""
    q = _localize(q, pm)
    if _fusable(pm):
        return _leapfrog_body(dx_i, p_i, stages, pt, lambda dx: force(dx, q, pm), fused=True,
                              stage_fn=(lambda dx, p, cd, ck, last: drift_force_kick(dx, p, cd, ck, q, pm, want_f=last))
                              if USE_FUSED_STAGE else None)
    return _leapfrog_body(dx_i, p_i, stages, pt, lambda dx: gravity(dx, q, pm))


def _firststep_body(rhok, q, a, pt, pm):
    """2LPT initial displacement and momentum at ``a`` (fastpm.py:479-495)"""
    dx1, dx2 = lpt(rhok, q, pm)
    E, D1, D2, f1, f2 = pt.E(a), pt.D1(a), pt.D2(a), pt.f1(a), pt.f2(a)
    dx1 = dx1 * D1
    dx2 = dx2 * D2
    p = dx1 * (a ** 2 * f1 * E) + dx2 * (a ** 2 * f2 * E)
    return dict(dx=dx1 + dx2, p=p)


@autooperator('rhok->dx,p')
def firststep(rhok, q, a, pt, pm):
    """fastpm.py:476-495"""
    return _firststep_body(rhok, q, a, pt, pm)


def _support(stages):
    stages = numpy.array(stages)
    mid = (stages[1:] * stages[:-1]) ** 0.5
    support = numpy.concatenate([mid, stages])
    support.sort()
    return stages, support


@autooperator('rhok->dx,p,f')
def nbody(rhok, q, stages, cosmology, pm):
    """fastpm.py:497-510"""
    stages, support = _support(stages)
    pt = MatterDominated(cosmology.Om0, a=support)
    q = _localize(q, pm)
    dx, p = firststep(rhok, q, stages[0], pt, pm)
    dx, p, f = leapfrog(dx, p, q, stages, pt, pm)
    return dict(dx=dx, p=p, f=f)


@operator
class cdot:
    """fastpm.py:512-527 -- real part of ``sum_k x1 conj(x2)`` over the full spectrum"""
    ain = {'x1': 'ComplexField', 'x2': 'ComplexField'}
    aout = {'y': '*'}

    def apl(node, x1, x2):
        return dict(y=x1.cdot(x2).real)

    def vjp(node, x1, x2, _y):
        return dict(_x1=x2.cdot_vjp(_y), _x2=x1.cdot_vjp(_y))

    def jvp(node, x1_, x2_, x1, x2):
        return dict(y_=x1.cdot(x2_).real + x2.cdot(x1_).real)


class FastPMSimulation(object):
    """fastpm.py:529-646 -- the same pipeline as ``nbody`` with the force mesh ``B`` times finer"""

    def __init__(self, stages, cosmology, pm, B=1, q=None):
        if q is None:
            q = pm.generate_uniform_particle_grid()
        stages, support = _support(stages)
        self.pt = MatterDominated(cosmology.Om0, a=support)
        self.stages = stages
        self.pm = pm
        self.fpm = type(pm)(Nmesh=pm.Nmesh * B, BoxSize=pm.BoxSize, dtype=pm.dtype,
                            comm=pm.comm, resampler=pm.resampler)
        self.q = _localize(q, pm)
        self.B = B
        # the models of the instance-bound autooperators below are memoised HERE (vm/ops.py:
        # AutoOperator._build), so they -- and the device arrays they hold -- die with the simulation
        self._vmad_model_cache = {}
        self._serial = next(FastPMSimulation._serials)

    _serials = itertools.count()

    def __vmad_hyper_key__(self):
        return ('sim', self._serial, numpy.asarray(self.stages).tobytes(), self.B)

    def KickFactor(self, ai, ac, af):
        return KickFactor(self.pt, ai, ac, af)

    def DriftFactor(self, ai, ac, af):
        return DriftFactor(self.pt, ai, ac, af)

    @autooperator('rhok->dx, p')
    def firststep(self, rhok):
        return _firststep_body(rhok, self.q, self.stages[0], self.pt, self.pm)

    @autooperator('dx->f,potk')
    def gravity(self, dx):
        return _gravity_body(dx, self.q, self.fpm)

    @autooperator('rhok->dx,p,f')
    def run(self, rhok):
        dx, p = self.firststep(rhok)
        if _fusable(self.fpm):
            return _leapfrog_body(dx, p, self.stages, self.pt,
                                  lambda dx: force(dx, self.q, self.fpm), fused=True,
                                  stage_fn=(lambda dx, p, cd, ck, last: drift_force_kick(dx, p, cd, ck, self.q, self.fpm, want_f=last))
                                  if USE_FUSED_STAGE else None)
        return _leapfrog_body(dx, p, self.stages, self.pt, lambda dx: self.gravity(dx))


# ---------------------------------------------------------------------------------------------
# fused operators named by the north star: force, kick, drift (not in the reference, where
# they are sub-graphs of stdlib mul / add nodes and the `gravity` autooperator, fastpm.py:409-474)
# ---------------------------------------------------------------------------------------------
USE_FUSED = True   # leapfrog / nbody / FastPMSimulation use force / kick / drift on the device
USE_FUSED_STAGE = True   # ... and one drift_force_kick operator per stage instead of the three


def set_fused(flag, stage=True):
    """switch the leapfrog between the fused operators and the reference's node-by-node graph
    (``stage``: also fuse each drift-force-kick stage into one operator); drops the memoised
    models so that the next build sees the new setting"""
    global USE_FUSED, USE_FUSED_STAGE
    USE_FUSED = bool(flag)
    USE_FUSED_STAGE = bool(stage)
    for op in (leapfrog, nbody, lpt, firststep, FastPMSimulation.run, FastPMSimulation.firststep):
        op._model_cache.clear()


def _fusable(pm):
    return USE_FUSED and hasattr(pm, 'rt') and pm.ndim == 3


@operator
class force:
    """``f, potk = gravity(dx, q, pm)`` (fastpm.py:409-438) as ONE operator with hand-written
    vjp and jvp: decompose + exchange + paint + FFT + fused Green/gradient pass + 3 inverse FFTs
    + fused 3-mesh readout + gather.  Device backend only; ``gravity`` is the portable graph."""
    ain = {'dx': 'ndarray'}
    aout = [('f', 'ndarray'), ('potk', 'ComplexField')]

    def apl(node, dx, q, pm):
        from .. import pm as dev
        f, potk, state = dev.force_apl(pm, dx, q)
        return dict(f=f, potk=potk, state=state)

    def vjp(node, _f, _potk, state, pm):
        from .. import pm as dev
        return dict(_dx=dev.force_vjp(pm, state, _f, _potk))

    def jvp(node, dx_, state, pm):
        from .. import pm as dev
        f_, potk_ = dev.force_jvp(pm, state, dx_)
        return dict(f_=f_, potk_=potk_)


@operator
class drift_force_kick:
    """One leapfrog stage of fastpm.py:462-471 as ONE operator::

        dx1 = dx + p * cd;   f, potk = gravity(dx1, q, pm);   p1 = p + f * ck

    The drift is fused with the position update ``q + dx1`` of the force evaluation; in the
    reverse sweep the factor ``ck`` rides on the exchange of the momentum cotangent and the
    cotangent of ``dx1`` from its other consumer is added inside the store of the force
    gradient, so the elementwise traffic of a stage drops from 19 particle-array passes to 12
    and three tape nodes become one.  Device backend only."""
    ain = [('dx', 'ndarray'), ('p', 'ndarray')]
    aout = [('dx1', 'ndarray'), ('p1', 'ndarray'), ('f', 'ndarray'), ('potk', 'ComplexField')]

    def apl(node, dx, p, cd, ck, q, pm, want_f=True):
        from .. import pm as dev
        dx1, x1 = dev.drift_pos(pm, dx, p, cd, q)
        # the kick rides on the store of the force readout; a stage whose force nobody reads
        # (every stage but the last) does not write it at all
        f, potk, state, p1 = dev.force_apl(pm, dx1, q, x=x1, kick=(p, ck), want_f=want_f)
        return dict(dx1=dx1, p1=p1, f=f if f is not None else 0, potk=potk, state=state)

    def vjp(node, _dx1, _p1, _f, _potk, state, cd, ck, pm):
        from .. import pm as dev
        # f feeds p1 (factor ck) and is an output itself; g = cotangent of dx1 from both of its
        # consumers; dx1 = dx + cd p, so _p = _p1 + cd g comes out of the same kernel
        if _is_zero(_p1):
            g, _p = dev.force_vjp(pm, state, _f, _potk, base=_dx1, update=(_p1, cd))
        elif _is_zero(_f):
            g, _p = dev.force_vjp(pm, state, _p1, _potk, f_scale=ck, base=_dx1, update=(_p1, cd))
        else:
            g, _p = dev.force_vjp(pm, state, _axpy(_f, _p1, ck), _potk, base=_dx1, update=(_p1, cd))
        return dict(_dx=g, _p=_p)

    def jvp(node, dx_, p_, state, cd, ck, pm):
        from .. import pm as dev
        if _is_zero(p_):
            dx1_ = dx_
        elif _is_zero(dx_):
            dx1_ = p_ * cd
        else:
            dx1_ = _axpy(dx_, p_, cd)
        f_, potk_ = dev.force_jvp(pm, state, dx1_)
        if _is_zero(f_):
            p1_ = p_
        elif _is_zero(p_):
            p1_ = f_ * ck
        else:
            p1_ = _axpy(p_, f_, ck)
        return dict(dx1_=dx1_, p1_=p1_, f_=f_, potk_=potk_)


def _axpy(y, x, fac):
    """y + x * fac in one pass on the device"""
    ax = getattr(y, '_axpby', None)
    if ax is not None and hasattr(x, 't') and x.t.shape == y.t.shape:
        return ax(1.0, float(fac), x.t)
    return y + x * fac


@operator
class kick:
    """``p + f * fac`` (fastpm.py:459-460,470-471)"""
    ain = 'p', 'f'
    aout = 'y'

    def apl(node, p, f, fac):
        return dict(y=_axpy(p, f, fac))

    def vjp(node, _y, fac):
        return dict(_p=_y, _f=_y * fac)

    def jvp(node, p_, f_, fac):
        if _is_zero(f_):
            return dict(y_=p_)
        if _is_zero(p_):
            return dict(y_=f_ * fac)
        return dict(y_=_axpy(p_, f_, fac))


@operator
class drift:
    """``dx + p * fac`` (fastpm.py:463-464)"""
    ain = 'dx', 'p'
    aout = 'y'

    def apl(node, dx, p, fac):
        return dict(y=_axpy(dx, p, fac))

    def vjp(node, _y, fac):
        return dict(_dx=_y, _p=_y * fac)

    def jvp(node, dx_, p_, fac):
        if _is_zero(p_):
            return dict(y_=dx_)
        if _is_zero(dx_):
            return dict(y_=p_ * fac)
        return dict(y_=_axpy(dx_, p_, fac))
