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


@pytest.mark.parametrize("name", ["nbody_8_2step", "nbody_16_3step"])
def test_golden_on_device(device_pm, name):
    g = load_golden(name)
    pm = device_pm([int(g['N'])] * 3, g['BoxSize'])
    errs = run_golden_case(pm, g, tol=1e-10, put=put)
    print(errs)


@pytest.mark.parametrize("N,nsteps", [(32, 5), (64, 3)])
def test_nbody_vs_oracle(oracle_pm, device_pm, N, nsteps):
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
