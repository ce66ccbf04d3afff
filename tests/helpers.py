"""Shared model builders for the parity tests (mirror tests/golden/make_golden.py, but on
vmad_b200's own VM and operator library, over whichever pm backend is handed in)."""
import os

import numpy

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Cosmology(object):
    Om0 = 0.3075


def power(k):
    k = numpy.where(k == 0, 1.0, k)
    return 1.5e6 * k / (1.0 + (k / 0.05) ** 2) ** 2


def load_golden(name):
    return numpy.load(os.path.join(GOLDEN, name + '.npz'))


def host(x):
    """device array / oracle field / ndarray -> ndarray"""
    if hasattr(x, 'numpy') and not isinstance(x, numpy.ndarray):
        return x.numpy()
    return numpy.asarray(x)


def nbody_model(pm, q, stages):
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    with Builder() as m:
        x = m.input('x')
        rhok = fastpm.induce_correlation(x, power, pm)
        dx, p, f = fastpm.nbody(rhok, q, stages, Cosmology, pm)
        m.output(dx=dx, p=p, f=f)
    return m


def lpt_model(pm, q):
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    with Builder() as m:
        x = m.input('x')
        dx1, dx2 = fastpm.lpt(x, q, pm)
        m.output(dx1=dx1, dx2=dx2)
    return m


def gravity_model(pm, q, op=None):
    from vmad_b200 import Builder
    from vmad_b200.lib import fastpm
    op = op or fastpm.gravity
    with Builder() as m:
        x = m.input('x')
        f, potk = op(x, q, pm)
        m.output(f=f, potk=potk)
    return m


def run_golden_case(pm, g, tol, put=lambda x: x):
    """run the three golden pipelines on ``pm``; ``put`` moves host arrays to the backend"""
    from conftest import rel_l2
    q = g['q']
    stages = g['stages']
    wn = pm.create('complex', value=g['wn'])
    m = nbody_model(pm, q, stages)
    cot = dict(_dx=put(g['cot_dx']), _p=put(g['cot_p']), _f=put(g['cot_f']))
    (dx, p, f), (gx,) = m.compute_with_vjp(init=dict(x=wn), v=cot)
    errs = {}
    errs['dx'] = rel_l2(host(dx), g['dx'])
    errs['p'] = rel_l2(host(p), g['p'])
    errs['f'] = rel_l2(host(f), g['f'])
    errs['vjp_x'] = rel_l2(host(gx), g['vjp_x'])
    tan = pm.create('complex', value=g['tan'])
    _, (dx_, p_, f_) = m.compute_with_jvp(['dx', 'p', 'f'], init=dict(x=wn), v=dict(x_=tan))
    errs['jvp_dx'] = rel_l2(host(dx_), g['jvp_dx'])
    errs['jvp_p'] = rel_l2(host(p_), g['jvp_p'])
    errs['jvp_f'] = rel_l2(host(f_), g['jvp_f'])

    m2 = lpt_model(pm, q)
    rk = pm.create('complex', value=g['rhok'])
    (dx1, dx2), (gl,) = m2.compute_with_vjp(init=dict(x=rk),
                                            v=dict(_dx1=put(g['cot_dx']), _dx2=put(g['cot_p'])))
    errs['lpt_dx1'] = rel_l2(host(dx1), g['lpt_dx1'])
    errs['lpt_dx2'] = rel_l2(host(dx2), g['lpt_dx2'])
    errs['lpt_vjp'] = rel_l2(host(gl), g['lpt_vjp'])

    m3 = gravity_model(pm, q)
    (f, potk), (gg,) = m3.compute_with_vjp(init=dict(x=put(g['dx'])),
                                           v=dict(_f=put(g['cot_f']), _potk=0))
    errs['grav_f'] = rel_l2(host(f), g['grav_f'])
    errs['grav_potk'] = rel_l2(host(potk), g['grav_potk'])
    errs['grav_vjp'] = rel_l2(host(gg), g['grav_vjp'])
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, "parity failures (rel L2 > %g): %s" % (tol, bad)
    return errs
