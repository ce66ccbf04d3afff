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
    # ---- bit-exact exchange routing against the P-rank restatement of pmesh's decompose ------------
    from pmesh import domain
    orc = domain.decompose_ranks([x_all[slice(r, None, P)] for r in range(P)], [N] * 3, L)

    def check_routing(lay, tag):
        class Bad(object):
            n = 0
            what = []

            def __iadd__(self, failed):
                import inspect
                if failed:
                    self.n += 1
                    self.what.append(inspect.stack()[1].lineno)
                return self
        bad = Bad()
        if hasattr(lay, 'plan'):        # fixed-capacity segments: compare segment by segment
            plan = lay.plan
            sidx = lay.send_idx.cpu().numpy()
            idx = lay.indices.cpu().numpy()
            for d in range(P):
                want = orc['send_idx'][rank][orc['send_start'][rank][d]:orc['send_start'][rank][d + 1]]
                seg = sidx[plan.sbase[d]:plan.sbase[d + 1]]
                bad += not numpy.array_equal(seg[:len(want)], want)
                bad += not (seg[len(want):] == -1).all()
            bad += not numpy.array_equal(lay.scnt.cpu().numpy()[:P], orc['counts'][rank])
            bad += not numpy.array_equal(lay.rcnt.cpu().numpy(), orc['counts'][:, rank])
            # sorted slot -> physical receive row -> (source, k) -> position in the packed receive order
            nrecv = int(orc['counts'][:, rank].sum())
            rb = numpy.array(plan.rbase)
            src = numpy.searchsorted(rb, idx[:nrecv], side='right') - 1
            k = idx[:nrecv] - rb[src]
            packed = numpy.concatenate([[0], numpy.cumsum(orc['counts'][:, rank])])[src] + k
            bad += not (k < orc['counts'][src, rank]).all()
            bad += not numpy.array_equal(packed, orc['perm'][rank])
        else:
            bad += not numpy.array_equal(lay.send_idx.cpu().numpy(), orc['send_idx'][rank])
            bad += not numpy.array_equal(numpy.array(lay.send_start), orc['send_start'][rank])
            bad += not numpy.array_equal(numpy.array(lay.recv_counts), orc['counts'][:, rank])
            bad += not numpy.array_equal(lay.indices.cpu().numpy(), orc['perm'][rank])
        errs['routing_bit_exact_' + tag] = float(bad.n)
        if bad.n:
            msg = "rank %d routing (%s): failed checks at lines %s" % (rank, tag, bad.what)
            if hasattr(lay, 'plan'):
                msg += "\nrank %d cap %s sbase %s rbase %s scnt %s rcnt %s oracle counts %s" % (
                    rank, lay.plan.cap.tolist(), lay.plan.sbase, lay.plan.rbase, lay.scnt.cpu().tolist(),
                    lay.rcnt.cpu().tolist(), orc['counts'].tolist())
            print(msg, flush=True)
            out = os.path.join(ROOT, 'gpurun_out')
            if os.path.isdir(out):
                with open(os.path.join(out, 'mp_worker_debug_rank%d.txt' % rank), 'a') as f:
                    f.write(msg + "\n")

    check_routing(lay, 'exact')
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

    # ---- fixed-capacity routing (counts stay on the device): same integers, same results ---------
    pm.plan_static_routing()
    lay = pm.decompose(x)
    check_routing(lay, 'static')
    errs['static_paint'] = rel(pm.paint(x, m, layout=lay).numpy(), opm.paint(x_all, m_all).value[rs])
    errs['static_readout'] = rel(df.readout(x, layout=lay).numpy(), of.readout(x_all)[mine])
    _mesh, _x = df.readout_vjp(x, v_all[mine], layout=lay)
    errs['static_readout_vjp_mesh'] = rel(_mesh.numpy(), omesh.value[rs])
    errs['static_readout_vjp_x'] = rel(_x.numpy(), ox[mine])
    md = nbody_model(pm, q, stages)
    cots = dict(_dx=put(cot_all[ps]), _p=put(cot_all[ps] * 0.5), _f=put(cot_all[ps] * 0.25))
    x0 = pm.create('complex', value=wn)
    (ddx, dp, dfo), (dg,) = md.compute_with_vjp(init=dict(x=x0), v=cots)
    pm.check_routing()
    errs['static_nbody_dx'] = rel(ddx.numpy(), odx[ps])
    errs['static_nbody_p'] = rel(dp.numpy(), op_[ps])
    errs['static_nbody_f'] = rel(dfo.numpy(), of_[ps])
    errs['static_nbody_vjp'] = rel(dg.numpy(), numpy.asarray(og)[:, cs])
    # ... and the whole P-rank forward + vjp captured into a CUDA graph on every rank and replayed
    from vmad_b200.cudagraph import GraphedCall

    def call():
        (a_, b_, c_), (g_,) = md.compute_with_vjp(init=dict(x=x0), v=cots)
        return a_, g_

    step = GraphedCall(call, warmup=1)
    for _ in range(3):
        gdx, gg = step()
    torch.cuda.synchronize()
    pm.check_routing()
    errs['graph_nbody_dx'] = rel(gdx.numpy(), odx[ps])
    errs['graph_nbody_vjp'] = rel(gg.numpy(), numpy.asarray(og)[:, cs])
    # an overflowing plan is reported, collectively
    small = numpy.full((P, P), 8, dtype='i8')
    pm.plan_static_routing(cap=small)
    pm.decompose(x)
    try:
        pm.check_routing()
        errs['overflow_detected'] = 1.0
    except Exception:
        errs['overflow_detected'] = 0.0

    bad = {k: v for k, v in errs.items() if not v < 1e-10}
    print("rank %d/%d: %s" % (rank, P, {k: '%.1e' % v for k, v in errs.items()}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        print("rank %d FAILED: %s" % (rank, bad), flush=True)
        sys.exit(1)


if __name__ == '__main__':
    main()
