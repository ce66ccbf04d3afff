"""CPU (oracle backend): apply_digitized re-hosted from the reference's tests, and the
Chi-square problem (objective / gradient / Gauss-Newton product) against finite differences."""
import numpy
import pytest

from vmad_b200 import Builder, autooperator
from vmad_b200.contrib.chisquare import MPIChiSquareProblem
from vmad_b200.lib import fastpm, linalg


def fd_check(m, x0, make, bases, eps=1e-4, rtol=1e-5):
    (y,), (g,) = m.compute_with_vjp(init=dict(x=make(x0)), v=dict(_y=1.0))
    g = numpy.asarray(g)
    for b in bases:
        yl = m.compute('y', init=dict(x=make(x0 - b * eps / 2)))
        yr = m.compute('y', init=dict(x=make(x0 + b * eps / 2)))
        fd = (yr - yl) / eps
        assert abs((g * b).sum() - fd) < rtol * max(1.0, abs(fd)), ((g * b).sum(), fd)
        _, y_ = m.compute_with_jvp('y', init=dict(x=make(x0)), v=dict(x_=make(b)))
        assert abs(y_ - fd) < rtol * max(1.0, abs(fd)), (y_, fd)


@pytest.mark.parametrize("mode", ['amplitude', 'phase'])
def test_apply_digitized_wrt_x_and_tf(oracle_pm, mode):
    """test_fastpm.py:68-133 re-hosted"""
    pm = oracle_pm([4, 4, 4], 4.0)
    kedges = numpy.linspace(0, 2 * numpy.pi / 4.0 * 3, 8)
    digitizer = fastpm.apply_digitized.isotropic_wavenumber(kedges)
    tf0 = numpy.linspace(-1, 2, len(kedges) - 1) if mode == 'phase' else numpy.arange(len(kedges) - 1) + 1.0
    x0 = pm.generate_whitenoise(seed=300, unitary=True, type='real').value
    rng = numpy.random.default_rng(0)

    with Builder() as m:
        x = m.input('x')
        c = fastpm.apply_digitized(fastpm.r2c(x), tf=tf0, digitizer=digitizer, kind='wavenumber', mode=mode)
        m.output(y=fastpm.to_scalar(fastpm.c2r(c)))
    fd_check(m, x0, lambda a: pm.create('real', value=a), [rng.standard_normal(x0.shape) for _ in range(3)])

    c0 = pm.generate_whitenoise(seed=300, unitary=True, type='complex', mean=1.0)
    with Builder() as m2:
        t = m2.input('x')
        c = fastpm.apply_digitized(c0, tf=t, digitizer=digitizer, kind='wavenumber', mode=mode)
        m2.output(y=fastpm.to_scalar(fastpm.c2r(c)))
    bases = [b for b in numpy.eye(len(tf0))]
    if mode == 'amplitude':
        fd_check(m2, tf0, lambda a: a, bases)


def test_chisquare_problem_on_oracle(oracle_pm):
    from helpers import Cosmology, power
    N = 8
    pm = oracle_pm([N] * 3, 25.0)
    q = pm.generate_uniform_particle_grid(shift=0.0)
    stages = numpy.linspace(0.1, 1.0, 3)
    rng = numpy.random.default_rng(3)
    data = pm.create('real', value=1.0 + 0.3 * rng.standard_normal(pm.rshape))

    @autooperator('x->model')
    def forward(x):
        rhok = fastpm.induce_correlation(x, power, pm)
        dx, p, f = fastpm.nbody(rhok, q, stages, Cosmology, pm)
        xf = linalg.add(dx, q)
        layout = fastpm.decompose(xf, pm)
        return dict(model=fastpm.paint(xf, 1.0, layout, pm))

    @autooperator('model->r')
    def residual(model):
        return dict(r=(model - data) * 0.5)

    problem = MPIChiSquareProblem(pm.comm, forward, [residual])
    wn = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    s = pm.create('complex', value=wn)
    y0 = problem.objective(s)
    g = numpy.asarray(problem.gradient(s))
    v = pm.create('complex', value=numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5)
    eps = 1e-5
    fd = (problem.objective(s + v * eps) - problem.objective(s - v * eps)) / (2 * eps)
    w = pm.hermitian_weights()
    # cotangents of complex fields are gradients w.r.t. the full spectrum (see the oracle)
    lhs = (w * (g.real * v.value.real + g.imag * v.value.imag)).sum()
    assert abs(lhs - fd) < 1e-5 * max(1.0, abs(fd)), (lhs, fd)
    # Gauss-Newton product: v^T (2 J^T J) v = 2 |J v|^2 >= 0, and matches a jvp
    Hv = numpy.asarray(problem.hessian_vector_product(s, v))
    vHv = (w * (Hv.real * v.value.real + Hv.imag * v.value.imag)).sum()
    with Builder() as mr:
        xx = mr.input('x')
        mr.output(y=residual(forward(xx)))
    _, r_ = mr.compute_with_jvp('y', init=dict(x=s), v=dict(x_=v))
    assert abs(vHv - 2 * (numpy.asarray(r_) ** 2).sum()) < 1e-8 * abs(vHv)
    assert y0 > 0
