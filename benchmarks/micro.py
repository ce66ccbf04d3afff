"""CIC paint / readout microbenchmark (BASELINE.json configs[2]): kernel-level Gparticles/s and
achieved algorithmic GB/s, timed with CUDA events on the launching stream."""
import argparse
import json
import sys
import os

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmad_b200.pm import ParticleMesh, DeviceArray  # noqa: E402
from vmad_b200 import _capi as C  # noqa: E402
from vmad_b200 import pm as pmmod  # noqa: E402


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(numpy.median(ts)), float(numpy.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nmesh', type=int, default=256)
    ap.add_argument('--log2np', type=int, nargs='+', default=[24])
    ap.add_argument('--dist', default='uniform', choices=['uniform', 'lattice'])
    ap.add_argument('--peak', type=float, default=6540.5)
    args = ap.parse_args()
    N = args.nmesh
    L = 100.0 * N / 32
    pm = ParticleMesh([N, N, N], L)
    Nc = N ** 3
    g = torch.Generator(device='cuda').manual_seed(1)
    for l2 in args.log2np:
        n = 1 << l2
        if args.dist == 'uniform':
            x = torch.rand((n, 3), generator=g, device='cuda', dtype=torch.float64) * L
        else:
            side = round(n ** (1 / 3))
            n = side ** 3
            q = (torch.stack(torch.meshgrid(*[torch.arange(side, device='cuda', dtype=torch.float64)] * 3,
                                            indexing='ij'), -1).reshape(-1, 3) * (L / side))
            x = q + torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64) * (1.5 * L / N)
        x = DeviceArray(x)
        m = DeviceArray(torch.rand((n,), generator=g, device='cuda', dtype=torch.float64) + 0.5)
        f = pm.create('real')
        f.t.copy_(torch.randn((N, N, N), generator=g, device='cuda', dtype=torch.float64))
        lay = pm.decompose(x)
        xs = lay.exchange(x)
        ms = lay.exchange(m)
        vs = lay.exchange(m)
        rows = []

        def rec(name, fn, bytes_):
            med, best = timeit(fn)
            rows.append(dict(op=name, np=n, ms=med * 1e3, ms_best=best * 1e3, gpart_s=n / med / 1e9,
                             gbs=bytes_ / med / 1e9, frac=bytes_ / med / 1e9 / args.peak))

        rec('decompose(sort)', lambda: pm.decompose(x), 48.0 * n)
        rec('exchange[n,3]', lambda: lay.exchange(x), 52.0 * n)
        rec('gather[n,3]', lambda: lay.gather(xs), 52.0 * n)
        rec('paint_sorted', lambda: pm.paint(xs, 1.0), 24.0 * n + 8.0 * Nc)
        pmmod.set_paint_algorithm('tile')
        rec('paint_tile(determ.)', lambda: pm.paint(xs, 1.0), 24.0 * n + 8.0 * Nc)
        m3 = DeviceArray(torch.rand((n, 3), generator=g, device='cuda', dtype=torch.float64))
        m3s = lay.exchange(m3)
        pmmod.set_paint_algorithm('gather')
        rec('paint_gather', lambda: pm.paint(xs, 1.0), 24.0 * n + 8.0 * Nc)
        rec('paint_gather_mass', lambda: pm.paint(xs, ms), 32.0 * n + 8.0 * Nc)
        rec('paint_gather_cols3', lambda: pm.paint_cols3(xs, m3s), 48.0 * n + 24.0 * Nc)
        pmmod.set_paint_algorithm('aggregated')
        rec('paint_agg_cols3', lambda: pm.paint_cols3(xs, m3s), 48.0 * n + 24.0 * Nc)
        rec('paint_agg_unsorted', lambda: pm._paint_any_order(x), 24.0 * n + 8.0 * Nc)
        rec('paint_rows(ablation)', lambda: pm.paint_rows(xs, 1.0), 24.0 * n + 8.0 * Nc)
        rec('paint_sorted_mass', lambda: pm.paint(xs, ms), 32.0 * n + 8.0 * Nc)
        rec('paint_atomic', lambda: pm.paint_atomic(x, 1.0), 24.0 * n + 8.0 * Nc)
        rec('paint_atomic_sortedx', lambda: pm.paint_atomic(xs, 1.0), 24.0 * n + 8.0 * Nc)
        rec('readout_sorted', lambda: f.readout(xs), 32.0 * n + 8.0 * Nc)
        rec('readout_unsorted', lambda: f.readout(x), 32.0 * n + 8.0 * Nc)
        rec('readout3_sorted', lambda: pm.readout3(f, f, f, xs), 48.0 * n + 24.0 * Nc)
        rec('readout_vjp_sorted', lambda: f.readout_vjp(xs, vs), 56.0 * n + 16.0 * Nc)
        rec('paint_vjp_sorted', lambda: pm.paint_vjp(f, xs, mass=1.0), 56.0 * n + 8.0 * Nc)
        rec('paint_jvp_sorted', lambda: pm.paint_jvp(xs, mass=1.0, v_pos=xs), 48.0 * n + 8.0 * Nc)
        rec('readout_jvp_sorted', lambda: f.readout_jvp(xs, v_self=f, v_pos=xs), 56.0 * n + 16.0 * Nc)
        r = f.r2c()
        rec('r2c', lambda: f.r2c(), 16.06 * Nc)
        rec('c2r', lambda: r.c2r(), 16.06 * Nc)
        rec('laplace', lambda: pm._transfer(r, laplace=True), 16.06 * Nc)
        a = DeviceArray(torch.rand((n, 3), generator=g, device='cuda', dtype=torch.float64))
        rec('axpy(kick)', lambda: a + xs * 0.5, 72.0 * n * 2 / 3 * 1.5)
        for row in rows:
            print(json.dumps(row))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
