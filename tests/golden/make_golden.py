"""Generate the golden vectors of tests/golden/*.npz.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

The vectors are produced by the REFERENCE's own code -- the unmodified vmad core
(/root/reference/vmad/core) driving the unmodified /root/reference/vmad/lib/fastpm.py --
with the numpy oracle (oracle/pmesh, oracle/fastpm) standing in for the absent pmesh and
fastpm packages.  They pin the operator wiring (LPT, gravity, leapfrog, n-body, their vjp and
jvp) that vmad_b200 must reproduce; numeric parity with pmesh itself stays "unpinned"
(oracle/__init__.py).
"""
import os
import sys
import warnings

import numpy

warnings.simplefilter('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import oracle  # noqa: E402
oracle.activate()

from vmad import Builder  # noqa: E402  (the reference's vmad)
from vmad.lib import fastpm, linalg  # noqa: E402
from pmesh.pm import ParticleMesh  # noqa: E402  (the oracle)


class Cosmology(object):
    Om0 = 0.3075


def power(k):
    k = numpy.where(k == 0, 1.0, k)
    return 1.5e6 * k / (1.0 + (k / 0.05) ** 2) ** 2


def case(N, nsteps, seed, shift):
    pm = ParticleMesh([N, N, N], 100.0 * N / 32)
    rng = numpy.random.default_rng(seed)
    wn = pm.create('real', value=rng.standard_normal(pm.rshape)).r2c()
    q = pm.generate_uniform_particle_grid(shift=shift)
    stages = numpy.linspace(0.1, 1.0, nsteps + 1)
    out = dict(N=N, nsteps=nsteps, BoxSize=pm.BoxSize, q=q, stages=stages, wn=wn.value,
               Om0=Cosmology.Om0)

    with Builder() as m:
        x = m.input('x')
        rhok = fastpm.induce_correlation(x, power, pm)
        dx, p, f = fastpm.nbody(rhok, q, stages, Cosmology, pm)
        m.output(dx=dx, p=p, f=f)

    cot = dict(_dx=rng.standard_normal(q.shape), _p=rng.standard_normal(q.shape),
               _f=rng.standard_normal(q.shape))
    (dx, p, f), (gx,) = m.compute_with_vjp(init=dict(x=wn), v=cot)
    out.update(dx=dx, p=p, f=f, cot_dx=cot['_dx'], cot_p=cot['_p'], cot_f=cot['_f'],
               vjp_x=numpy.asarray(gx))
    with Builder() as m0:
        x = m0.input('x')
        m0.output(rhok=fastpm.induce_correlation(x, power, pm))
    out['rhok'] = numpy.asarray(m0.compute('rhok', init=dict(x=wn)))

    tan = pm.create('real', value=rng.standard_normal(pm.rshape)).r2c()
    _, (dx_, p_, f_) = m.compute_with_jvp(['dx', 'p', 'f'], init=dict(x=wn), v=dict(x_=tan))
    out.update(tan=tan.value, jvp_dx=dx_, jvp_p=p_, jvp_f=f_)

    # LPT alone (config 1: Zel'dovich / 2LPT forward + vjp)
    with Builder() as m2:
        x = m2.input('x')
        dx1, dx2 = fastpm.lpt(x, q, pm)
        m2.output(dx1=dx1, dx2=dx2)
    rk = pm.create('complex', value=out['rhok'])
    (dx1, dx2), (g,) = m2.compute_with_vjp(init=dict(x=rk), v=dict(_dx1=cot['_dx'], _dx2=cot['_p']))
    out.update(lpt_dx1=dx1, lpt_dx2=dx2, lpt_vjp=numpy.asarray(g))

    # one force evaluation
    with Builder() as m3:
        x = m3.input('x')
        f, potk = fastpm.gravity(x, q, pm)
        m3.output(f=f, potk=potk)
    (f, potk), (g,) = m3.compute_with_vjp(init=dict(x=out['dx']), v=dict(_f=cot['_f'], _potk=0))
    out.update(grav_f=f, grav_potk=numpy.asarray(potk), grav_vjp=g)
    return out


if __name__ == '__main__':
    for name, args in [('nbody_8_2step', (8, 2, 300, 0.0)), ('nbody_16_3step', (16, 3, 301, 0.5))]:
        r = case(*args)
        path = os.path.join(HERE, name + '.npz')
        numpy.savez_compressed(path, **r)
        print(path, os.path.getsize(path) // 1024, 'KiB')
