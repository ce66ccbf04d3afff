"""GPU: the FastPM pipelines on the B200 backend against (a) the golden vectors produced by the
reference's code and (b) the numpy oracle on the same seeded inputs at larger sizes.
Bar (north_star): multi-step FastPM forward / vjp / jvp within 1e-10 relative L2."""
import numpy
import pytest

from conftest import rel_l2
from helpers import Cosmology, host, load_golden, nbody_model, power, run_golden_case

pytestmark = pytest.mark.gpu


def put(x):
    from vmad_b200.pm import DeviceArray
    return DeviceArray.from_host(x)


@pytest.fixture(params=[(True, True), (True, False), (False, False)], ids=['stage', 'fused', 'unfused'])
def fused(request):
    """'stage': one drift_force_kick operator per leapfrog stage (default); 'fused': separate
    force / kick / drift operators; 'unfused': the reference's node-by-node graph"""
    from vmad_b200.lib import fastpm
    fastpm.set_fused(*request.param)
    yield request.param[0]
    fastpm.set_fused(True)


@pytest.mark.parametrize("name", ["nbody_8_2step", "nbody_16_3step"])
def test_golden_on_device(device_pm, fused, name):
    g = load_golden(name)
    pm = device_pm([int(g['N'])] * 3, g['BoxSize'])
    errs = run_golden_case(pm, g, tol=1e-10, put=put)
    print(errs)


@pytest.mark.parametrize("N,nsteps", [(32, 5), (64, 3)])
def test_nbody_vs_oracle(oracle_pm, device_pm, fused, N, nsteps):
    """config-1/2 style: white noise -> induce_correlation -> nbody, forward + vjp + jvp"""
    L = 100.0 * N / 32
    opm = oracle_pm([N] * 3, L)
    dpm = device_pm([N] * 3, L)
    rng = numpy.random.default_rng(300 + N)
    wn = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    tan = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    q = opm.generate_uniform_particle_grid(shift=0.0)
    stages = numpy.linspace(0.1, 1.0, nsteps + 1)
    cot = [rng.standard_normal(q.shape) for _ in range(3)]
    res = {}
    for name, pm, mv in [('oracle', opm, lambda x: x), ('device', dpm, put)]:
        m = nbody_model(pm, q, stages)
        x = pm.create('complex', value=wn)
        (dx, p, f), (gx,) = m.compute_with_vjp(
            init=dict(x=x), v=dict(_dx=mv(cot[0]), _p=mv(cot[1]), _f=mv(cot[2])))
        _, (dx_, p_, f_) = m.compute_with_jvp(['dx', 'p', 'f'], init=dict(x=x),
                                              v=dict(x_=pm.create('complex', value=tan)))
        res[name] = [host(a) for a in (dx, p, f, gx, dx_, p_, f_)]
    names = ['dx', 'p', 'f', 'vjp', 'jvp_dx', 'jvp_p', 'jvp_f']
    errs = {n: rel_l2(a, b) for n, a, b in zip(names, res['device'], res['oracle'])}
    print(errs)
    # displacements must be non-trivial (about a cell at a = 1) for the test to mean something
    assert numpy.std(res["oracle"][0]) > 0.5 * L / N
    assert all(e < 1e-10 for e in errs.values()), errs


def test_force_operator_matches_gravity(device_pm):
    """the fused force operator against the node-by-node gravity graph: f, potk, vjp, jvp"""
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    N = 32
    pm = device_pm([N] * 3, 100.0)
    rng = numpy.random.default_rng(5)
    q = pm.generate_uniform_particle_grid(shift=0.0)
    dx = put(rng.standard_normal((N ** 3, 3)) * 2.5)
    cf = put(rng.standard_normal((N ** 3, 3)))
    cp = pm.create('complex', value=rng.standard_normal(pm.cshape) + 1j * rng.standard_normal(pm.cshape))
    tan = put(rng.standard_normal((N ** 3, 3)))
    out = {}
    variants = {'gravity': (fastpm.gravity, True, False),
                'force_stencil': (fastpm.force, True, False),      # real-space gradients (default)
                'force_kspace': (fastpm.force, False, False),      # k-space gradients, meshes kept
                'force_kspace_rc': (fastpm.force, False, True)}    # ... meshes recomputed
    for name, (op, stencil, recompute) in variants.items():
        pm.force_stencil = stencil
        pm.force_recompute = recompute
        with Builder() as m:
            x = m.input('x')
            f, potk = op(x, q, pm)
            m.output(f=f, potk=potk)
        (f, potk), (g,) = m.compute_with_vjp(init=dict(x=dx), v=dict(_f=cf, _potk=cp))
        _, (f_, potk_) = m.compute_with_jvp(['f', 'potk'], init=dict(x=dx), v=dict(x_=tan))
        out[name] = [host(a) for a in (f, potk, g, f_, potk_)]
    for variant in ('force_stencil', 'force_kspace', 'force_kspace_rc'):
        for a, b, what in zip(out[variant], out['gravity'], ['f', 'potk', 'vjp', 'jvp_f', 'jvp_potk']):
            print(variant, what, rel_l2(a, b))
            assert rel_l2(a, b) < 1e-12, (variant, what, rel_l2(a, b))


def test_chisquare_gauss_newton_vs_oracle(oracle_pm, device_pm):
    """config-4 shape (contrib Chi-square: objective, gradient, Gauss-Newton J^T J v) on the
    device against the same problem on the oracle backend"""
    from vmad_b200 import autooperator
    from vmad_b200.contrib.chisquare import MPIChiSquareProblem
    from vmad_b200.lib import fastpm, linalg
    N = 16
    L = 50.0
    rng = numpy.random.default_rng(3)
    data_h = 1.0 + 0.3 * rng.standard_normal((N, N, N))
    wn = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    vh = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    stages = numpy.linspace(0.1, 1.0, 6)
    out = {}
    for name, PM in [('oracle', oracle_pm), ('device', device_pm)]:
        pm = PM([N] * 3, L)
        q = pm.generate_uniform_particle_grid(shift=0.0)
        data = pm.create('real', value=data_h)

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
        s = pm.create('complex', value=wn)
        v = pm.create('complex', value=vh)
        out[name] = [float(problem.objective(s)), host(problem.gradient(s)),
                     host(problem.hessian_vector_product(s, v))]
    assert abs(out['device'][0] - out['oracle'][0]) < 1e-11 * abs(out['oracle'][0])
    assert rel_l2(out['device'][1], out['oracle'][1]) < 1e-10
    assert rel_l2(out['device'][2], out['oracle'][2]) < 1e-10


def test_apply_digitized_vs_oracle(oracle_pm, device_pm):
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    N = 8
    kedges = numpy.linspace(0, 2 * numpy.pi / 4.0 * 3, 8)
    digitizer = fastpm.apply_digitized.isotropic_wavenumber(kedges)
    rng = numpy.random.default_rng(9)
    xh = rng.standard_normal((N, N, N))
    out = {}
    for mode in ('amplitude', 'phase'):
        tf0 = numpy.linspace(-1, 2, len(kedges) - 1) if mode == 'phase' else numpy.arange(len(kedges) - 1) + 1.0
        for name, PM in [('oracle', oracle_pm), ('device', device_pm)]:
            pm = PM([N] * 3, 4.0)
            c0 = pm.create('real', value=xh).r2c()
            with Builder() as m:
                t = m.input('x')
                c = fastpm.apply_digitized(c0, tf=t, digitizer=digitizer, kind='wavenumber', mode=mode)
                m.output(y=fastpm.to_scalar(fastpm.c2r(c)))
            (y,), (g,) = m.compute_with_vjp(init=dict(x=tf0), v=dict(_y=1.0))
            out[name] = (float(y), numpy.asarray(g))
        assert abs(out['device'][0] - out['oracle'][0]) < 1e-12 * abs(out['oracle'][0])
        assert numpy.allclose(out['device'][1], out['oracle'][1], rtol=1e-10, atol=1e-12)


def test_cuda_graph_replay_matches_eager(oracle_pm, device_pm):
    """The whole forward+vjp captured into a CUDA graph (vmad_b200/cudagraph.py) and replayed with
    NEW input values through the same buffers equals the eager VM run and the oracle."""
    import torch
    from vmad_b200.cudagraph import GraphedCall
    N, nsteps = 16, 3
    L = 100.0 * N / 32
    opm = oracle_pm([N] * 3, L)
    dpm = device_pm([N] * 3, L)
    rng = numpy.random.default_rng(7)
    wn1 = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    wn2 = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    q = opm.generate_uniform_particle_grid(shift=0.0)
    stages = numpy.linspace(0.1, 1.0, nsteps + 1)
    cot = rng.standard_normal(q.shape)
    # a constant of the graph must be a device array (or a read-only host array): a writeable
    # numpy q cannot key the model memo, so every call would rebuild the model and upload q again
    # -- as the reference does -- which a capture forbids
    m = nbody_model(dpm, dpm.asparticles(q), stages)
    x = dpm.create('complex', value=wn1)
    c = put(cot)

    def call():
        (dx,), (gx,) = m.compute_with_vjp(init=dict(x=x), v=dict(_dx=c))
        return dx, gx

    step = GraphedCall(call)
    x.t.copy_(torch.from_numpy(wn2).to(x.t.device))       # new values, same buffer
    dx, gx = step()
    got = [host(dx).copy(), host(gx).copy()]
    dx, gx = call()                                       # eager, same inputs
    eager = [host(dx), host(gx)]
    om = nbody_model(opm, q, stages)
    (odx,), (ogx,) = om.compute_with_vjp(init=dict(x=opm.create('complex', value=wn2)), v=dict(_dx=cot))
    for a, b, o in zip(got, eager, [host(odx), host(ogx)]):
        assert rel_l2(a, b) < 1e-12
        assert rel_l2(a, o) < 1e-10


@pytest.mark.parametrize("B", [1, 2])
def test_fastpm_simulation_vs_oracle(oracle_pm, device_pm, fused, B):
    """FastPMSimulation (fastpm.py:529-646): the force mesh B times finer than the particle grid;
    forward + vjp of ``run`` on the device == the oracle"""
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    N = 16
    L = 100.0 * N / 32
    rng = numpy.random.default_rng(19)
    wn = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    stages = [0.1, 0.5, 1.0]
    res = {}
    for name, pm, mv in [('oracle', oracle_pm([N] * 3, L), lambda x: x), ('device', device_pm([N] * 3, L), put)]:
        q = oracle_pm([N] * 3, L).generate_uniform_particle_grid(shift=0.0)
        cot = numpy.random.default_rng(5).standard_normal(q.shape)
        sim = fastpm.FastPMSimulation(stages, Cosmology, pm, B=B, q=q)
        with Builder() as m:
            x = m.input('x')
            rhok = fastpm.induce_correlation(x, power, pm)
            dx, p, f = sim.run(rhok)
            m.output(dx=dx, p=p, f=f)
        (dx, p, f), (gx,) = m.compute_with_vjp(init=dict(x=pm.create('complex', value=wn)),
                                               v=dict(_dx=mv(cot), _p=mv(cot * 0.5), _f=mv(cot * 0.1)))
        res[name] = [host(a) for a in (dx, p, f, gx)]
    errs = [rel_l2(a, b) for a, b in zip(res['device'], res['oracle'])]
    assert all(e < 1e-10 for e in errs), errs


def test_scalar_reductions_and_entry_points_vs_oracle(oracle_pm, device_pm):
    """to_scalar, cdot and as_complex_field (fastpm.py:9-42,512-527) at operator level:
    values, vjp and jvp on the device == the oracle"""
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    N = 12
    rng = numpy.random.default_rng(23)
    pairs = rng.standard_normal((N, N, N // 2 + 1, 2))
    c2v = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    tan = rng.standard_normal(pairs.shape)
    res = {}
    for name, pm in [('oracle', oracle_pm([N] * 3, 30.0)), ('device', device_pm([N] * 3, 30.0))]:
        with Builder() as m:
            a, c2 = m.input('a', 'c2')
            c1 = fastpm.as_complex_field(a, pm)
            s1 = fastpm.to_scalar(fastpm.c2r(c1))
            s2 = fastpm.cdot(c1, c2)
            m.output(y=s1 + s2 * 3.0)
        init = dict(a=pairs, c2=pm.create('complex', value=c2v))
        y, (ga, gc2) = m.compute_with_vjp(init=init, v=dict(_y=1.0))
        _, (y_,) = m.compute_with_jvp(['y'], init=init, v=dict(a_=tan, c2_=0))
        res[name] = [numpy.asarray(host(v), dtype='c16' if numpy.iscomplexobj(host(v)) else 'f8')
                     for v in (y[0] if isinstance(y, (list, tuple)) else y, ga, gc2, y_)]
    for a, b in zip(res['device'], res['oracle']):
        assert rel_l2(numpy.atleast_1d(a), numpy.atleast_1d(b)) < 1e-12


def test_lpt_real_space_operators_vs_graph(oracle_pm, device_pm):
    """lpt2src_real_space / lpt1_at_constant_q (one inverse FFT + stencils) against the reference
    graph (k-space second derivatives, six inverse FFTs) on the oracle: value, vjp, jvp, on an
    anisotropic mesh"""
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    N, L = [12, 8, 16], [30.0, 16.0, 48.0]
    opm, dpm = oracle_pm(N, L), device_pm(N, L)
    rng = numpy.random.default_rng(31)
    rk = numpy.fft.rfftn(rng.standard_normal(N)) / numpy.prod(N)
    tk = numpy.fft.rfftn(rng.standard_normal(N)) / numpy.prod(N)
    sbar = rng.standard_normal(N)
    q = opm.generate_uniform_particle_grid(shift=0.25)
    dbar = rng.standard_normal(q.shape)
    res = {}
    for name, pm, src, one, mv in [('oracle', opm, fastpm.lpt2src, fastpm.lpt1, lambda a: a),
                                   ('device', dpm, fastpm.lpt2src_real_space, fastpm.lpt1_at_constant_q, put)]:
        qq = q if name == 'oracle' else dpm.asparticles(q)
        with Builder() as m:
            x = m.input('x')
            m.output(s=src(x, pm), d=one(x, qq, pm))
        init = dict(x=pm.create('complex', value=rk))
        (s, d), (gx,) = m.compute_with_vjp(init=init, v=dict(_s=pm.create('real', value=sbar), _d=mv(dbar)))
        _, (s_, d_) = m.compute_with_jvp(['s', 'd'], init=init, v=dict(x_=pm.create('complex', value=tk)))
        res[name] = [host(a) for a in (s, d, gx, s_, d_)]
    for a, b in zip(res['device'], res['oracle']):
        assert rel_l2(a, b) < 1e-12


def test_neg_gradient_order0_vs_oracle(oracle_pm, device_pm):
    """fourier_space_neg_gradient(order=0) (fastpm.py:316-319: -i k with the Nyquist mask):
    apply_transfer forward, vjp (conjugate filter) and jvp on the device against the oracle"""
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    N, L = 16, 37.0
    rng = numpy.random.default_rng(17)
    xh = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    ch = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))
    for d in range(3):
        out = {}
        for name, PM in [('oracle', oracle_pm), ('device', device_pm)]:
            pm = PM([N] * 3, L)
            with Builder() as m:
                x = m.input('x')
                m.output(y=fastpm.apply_transfer(x, fastpm.fourier_space_neg_gradient(d, pm, order=0)))
            x0 = pm.create('complex', value=xh)
            (y,), (g,) = m.compute_with_vjp(init=dict(x=x0), v=dict(_y=pm.create('complex', value=ch)))
            _, (y_,) = m.compute_with_jvp(['y'], init=dict(x=x0), v=dict(x_=pm.create('complex', value=ch)))
            out[name] = [host(y), host(g), host(y_)]
        for a, b in zip(out['device'], out['oracle']):
            assert rel_l2(a, b) < 1e-12
        assert numpy.abs(out['oracle'][0]).max() > 0


def test_lpt1_gradient_wrt_q_vs_oracle(oracle_pm, device_pm):
    """lpt1's spec is 'rhok,q->dx1' (fastpm.py:336-351): the readout's position cotangent flows to
    ``q``.  Forward, vjp w.r.t. BOTH inputs and jvp on the device against the oracle."""
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    N, L = 16, 50.0
    rng = numpy.random.default_rng(23)
    rk = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5 * 20.0
    qh = rng.uniform(0, L, size=(3000, 3))
    cot = rng.standard_normal(qh.shape)
    tq = rng.standard_normal(qh.shape)
    out = {}
    for name, PM, mv in [('oracle', oracle_pm, lambda a: a), ('device', device_pm, put)]:
        pm = PM([N] * 3, L)
        with Builder() as m:
            r, q = m.input('r', 'q')
            m.output(dx1=fastpm.lpt1(r, q, pm))
        r0 = pm.create('complex', value=rk)
        (dx1,), (gr, gq) = m.compute_with_vjp(init=dict(r=r0, q=mv(qh)), v=dict(_dx1=mv(cot)))
        _, (dx1_,) = m.compute_with_jvp(['dx1'], init=dict(r=r0, q=mv(qh)), v=dict(r_=0, q_=mv(tq)))
        out[name] = [host(dx1), host(gr), host(gq), host(dx1_)]
    for what, a, b in zip(('dx1', '_rhok', '_q', 'jvp_q'), out['device'], out['oracle']):
        assert rel_l2(a, b) < 1e-11, (what, rel_l2(a, b))
    assert numpy.abs(out['oracle'][2]).max() > 0
