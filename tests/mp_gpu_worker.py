"""Worker for the multi-GPU tests (launched by torchrun from tests/test_gpu_multi.py): slab
decomposition over WORLD_SIZE GPUs must reproduce the 1-rank oracle results."""
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def rel(a, b):
    return float(numpy.linalg.norm(numpy.asarray(a) - numpy.asarray(b)) / numpy.linalg.norm(b))


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from vmad_b200.dist import TorchComm
    from vmad_b200.pm import DeviceArray, ParticleMesh
    import oracle
    oracle.activate(shims=False)
    from pmesh.pm import ParticleMesh as OraclePM
    from helpers import nbody_model

    comm = TorchComm()
    P, rank = comm.size, comm.rank
    N = 32
    L = 100.0
    pm = ParticleMesh([N] * 3, L, comm=comm)
    opm = OraclePM([N] * 3, L)
    rng = numpy.random.default_rng(11)
    errs = {}

    # ---- slab FFT against numpy ---------------------------------------------------------
    r = rng.standard_normal((N, N, N))
    c = rng.standard_normal(opm.cshape) + 1j * rng.standard_normal(opm.cshape)
    n0l, n1l = N // P, N // P
    rs, cs = slice(rank * n0l, (rank + 1) * n0l), slice(rank * n1l, (rank + 1) * n1l)
    dr = pm.create('real', value=r)
    dc = pm.create('complex', value=c)
    errs['r2c'] = rel(dr.r2c().numpy(), opm.create('real', value=r).r2c().value[:, cs])
    errs['c2r'] = rel(dc.c2r().numpy(), opm.create('complex', value=c).c2r().value[rs])
    errs['c2r_keeps_input'] = 0.0 if numpy.array_equal(dc.numpy(), c[:, cs]) else 1.0
    errs['cnorm'] = abs(dc.cnorm() - opm.create('complex', value=c).cnorm()) / abs(dc.cnorm())
    errs['rnorm'] = abs(dr.cnorm() - (r ** 2).sum()) / (r ** 2).sum()

    # ---- routing, paint, readout: every rank holds an arbitrary share of the particles -----
    x_all = rng.uniform(-0.3 * L, 1.3 * L, size=(20000, 3))
    m_all = rng.uniform(0.5, 1.5, size=20000)
    mine = slice(rank, None, P)            # interleaved ownership: positions unrelated to slabs
    x, m = x_all[mine], m_all[mine]
    lay = pm.decompose(x)
    assert lay.oldlength == len(x)
    errs['paint'] = rel(pm.paint(x, m, layout=lay).numpy(), opm.paint(x_all, m_all).value[rs])
    f = rng.standard_normal((N, N, N))
    df = pm.create('real', value=f)
    of = opm.create('real', value=f)
    errs['readout'] = rel(df.readout(x, layout=lay).numpy(), of.readout(x_all)[mine])
    v_all = rng.standard_normal(20000)
    _mesh, _x = df.readout_vjp(x, v_all[mine], layout=lay)
    omesh, ox = of.readout_vjp(x_all, v_all)
    errs['readout_vjp_mesh'] = rel(_mesh.numpy(), omesh.value[rs])
    errs['readout_vjp_x'] = rel(_x.numpy(), ox[mine])
    a = rng.standard_normal((20000, 3))
    errs['gather_exchange'] = rel(lay.gather(lay.exchange(a[mine])).numpy() /
                                  numpy.maximum(1.0, lay.gather(lay.exchange(numpy.ones(len(x)))).numpy()[:, None]),
                                  a[mine])

    # ---- the whole pipeline: nbody forward + vjp, P ranks == 1 rank (oracle) --------------------
    nsteps = 3
    wn = numpy.fft.rfftn(rng.standard_normal((N, N, N))) / N ** 1.5
    q_all = opm.generate_uniform_particle_grid(shift=0.0)
    cot_all = rng.standard_normal(q_all.shape)
    stages = numpy.linspace(0.1, 1.0, nsteps + 1)
    mo = nbody_model(opm, q_all, stages)
    (odx, op_, of_), (og,) = mo.compute_with_vjp(
        init=dict(x=opm.create('complex', value=wn)), v=dict(_dx=cot_all, _p=cot_all * 0.5, _f=cot_all * 0.25))
    q = pm.generate_uniform_particle_grid(shift=0.0)
    npl = len(q)
    ps = slice(rank * npl, (rank + 1) * npl)
    md = nbody_model(pm, q, stages)
    put = DeviceArray.from_host
    (ddx, dp, dfo), (dg,) = md.compute_with_vjp(
        init=dict(x=pm.create('complex', value=wn)),
        v=dict(_dx=put(cot_all[ps]), _p=put(cot_all[ps] * 0.5), _f=put(cot_all[ps] * 0.25)))
    errs['nbody_dx'] = rel(ddx.numpy(), odx[ps])
    errs['nbody_p'] = rel(dp.numpy(), op_[ps])
    errs['nbody_f'] = rel(dfo.numpy(), of_[ps])
    errs['nbody_vjp'] = rel(dg.numpy(), numpy.asarray(og)[:, cs])
    # memory-lean mode of the largest runs (potential recomputed, no composed routing indices)
    pm.force_recompute = True
    md = nbody_model(pm, q, stages)
    (ddx, dp, dfo), (dg,) = md.compute_with_vjp(
        init=dict(x=pm.create('complex', value=wn)),
        v=dict(_dx=put(cot_all[ps]), _p=put(cot_all[ps] * 0.5), _f=put(cot_all[ps] * 0.25)))
    errs['lean_dx'] = rel(ddx.numpy(), odx[ps])
    errs['lean_vjp'] = rel(dg.numpy(), numpy.asarray(og)[:, cs])
    pm.force_recompute = False

    bad = {k: v for k, v in errs.items() if not v < 1e-10}
    print("rank %d/%d: %s" % (rank, P, {k: '%.1e' % v for k, v in errs.items()}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        print("rank %d FAILED: %s" % (rank, bad), flush=True)
        sys.exit(1)


if __name__ == '__main__':
    main()
