"""CPU, gloo, world_size 2: host-side logic of the multi-GPU path (no CUDA involved)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gloo_world2_plumbing():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29517',
           os.path.join(ROOT, 'tests', 'mp_cpu_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert r.stdout.count(' ok') == 2, tail


def test_routing_plan_segments_tile_the_areas():
    """fixed-capacity routing (vmad_b200.pm.RoutingPlan): for every capacity matrix the send
    segments of a rank tile its send order, the segments the sources write tile the receiver's
    receive order, and the gather direction lands every block back in its sender's send order"""
    import numpy
    from vmad_b200.pm import RoutingPlan
    rng = numpy.random.default_rng(5)
    for P in (2, 3, 4, 8):
        cap = rng.integers(0, 50, size=(P, P)) * 4096
        plans = [RoutingPlan(cap, r) for r in range(P)]
        for r, pl in enumerate(plans):
            assert pl.sbase == [0] + list(numpy.cumsum(cap[r, :]))
            assert pl.rbase == [0] + list(numpy.cumsum(cap[:, r]))
            assert pl.cap_send == cap[r, :].sum() and pl.cap_recv == cap[:, r].sum()
            assert pl.max_rows >= max(pl.cap_send, pl.cap_recv)
        for d in range(P):
            # exchange: source s writes rows [fwd_base_s[d], + cap[s][d]) of d's receive order
            covered = numpy.zeros(plans[d].cap_recv, dtype=int)
            for s in range(P):
                lo = plans[s].fwd_base[d]
                assert lo == plans[d].rbase[s]
                covered[lo:lo + cap[s, d]] += 1
            assert (covered == 1).all()
        for s in range(P):
            # gather: rank d returns its block to rows [back_base_d[s], + cap[s][d]) of s's send order
            covered = numpy.zeros(plans[s].cap_send, dtype=int)
            for d in range(P):
                lo = plans[d].back_base[s]
                assert lo == plans[s].sbase[d]
                covered[lo:lo + cap[s, d]] += 1
            assert (covered == 1).all()
