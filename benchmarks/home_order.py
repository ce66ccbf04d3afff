"""Sorted-order pipeline (cell sort + exchange + aggregated paint + scatter-store readouts) against
the same work done in HOME order (lattice order of the Lagrangian grid, which is cell-coherent in a
FastPM run): paint by plain reductions, readouts with coalesced particle arrays.

    python benchmarks/home_order.py 512 0.3 1.5 3.0      # mesh size, rms displacements in cells
"""
import ctypes
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmad_b200 import _capi as C
from vmad_b200.pm import ParticleMesh, DeviceArray, _ptr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sigmas = [float(s) for s in sys.argv[2:]] or [0.3, 1.5, 3.0]
L = 100.0 * N / 32
pm = ParticleMesh([N] * 3, L)
rt = pm.rt
g = torch.Generator(device='cuda').manual_seed(1)
q = pm.generate_uniform_particle_grid(shift=0.0)
n = len(q)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, iters=5):
    ts = []
    for i in range(iters + 1):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rt.sync_stream()
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[1:])[len(ts[1:]) // 2]


def rnd():
    return torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64)


F = [torch.randn(tuple(pm.rshape), generator=g, device='cuda', dtype=torch.float64) for _ in range(4)]
mesh = ctypes.byref(pm._mesh)
for sigma in sigmas:
    x = DeviceArray(q.t + rnd() * (sigma * L / N))
    p, fbar, base, ybar = rnd(), rnd(), rnd(), rnd()
    rec = dict(nmesh=N, sigma_cells=sigma)
    # ---- sorted pipeline ----
    lay = pm.decompose(x)
    rec['sort_ms'] = timeit(lambda: pm.decompose(x))
    xl = lay.exchange(x)
    rec['exchange_ms'] = timeit(lambda: lay.exchange(x))
    rho_s = pm.paint(xl, 1.0)
    rec['paint_sorted_ms'] = timeit(lambda: pm.paint(xl, 1.0))
    val = rt.empty((n, 3)); yo = rt.empty((n, 3))
    f_s = lambda: C.check(C.lib.vpm_readout3(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(xl.t), n,
                                             _ptr(lay.indices), None, _ptr(p), 0.5, _ptr(yo)))
    f_s(); yo_s = yo.clone()
    rec['readout3_sorted_scatter_ms'] = timeit(f_s)
    fl = lay.exchange(DeviceArray(fbar))
    rec['exchange_f_ms'] = timeit(lambda: lay.exchange(DeviceArray(fbar)))
    m3 = pm.paint_cols3(xl, fl)
    rec['paint3_sorted_ms'] = timeit(lambda: pm.paint_cols3(xl, fl))
    gg = rt.empty((n, 3)); y2 = rt.empty((n, 3))
    g_s = lambda: C.check(C.lib.vpm_readout_grad4(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(F[3]),
                                                  _ptr(xl.t), _ptr(fl.t), n, _ptr(lay.indices), _ptr(base), _ptr(gg),
                                                  _ptr(ybar), 0.25, _ptr(y2)))
    g_s(); gg_s = gg.clone()
    rec['grad4_sorted_scatter_ms'] = timeit(g_s)
    # ---- home order ----
    rho_h = rt.empty(tuple(pm.rshape))
    f_p = lambda: C.check(C.lib.vpm_paint_atomic(rt.ctx, mesh, _ptr(x.t), None, 1.0, n, _ptr(rho_h)))
    f_p()
    rec['paint_home_atomic_ms'] = timeit(f_p)
    rec['paint_rel'] = float((rho_h - rho_s.t).norm() / rho_s.t.norm())
    f_h = lambda: C.check(C.lib.vpm_readout3(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(x.t), n,
                                             None, None, _ptr(p), 0.5, _ptr(yo)))
    f_h()
    rec['readout3_rel'] = float((yo - yo_s).norm() / yo_s.norm())
    rec['readout3_home_ms'] = timeit(f_h)
    g_h = lambda: C.check(C.lib.vpm_readout_grad4(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(F[3]),
                                                  _ptr(x.t), _ptr(fbar), n, None, _ptr(base), _ptr(gg),
                                                  _ptr(ybar), 0.25, _ptr(y2)))
    g_h()
    rec['grad4_rel'] = float((gg - gg_s).norm() / gg_s.norm())
    rec['grad4_home_ms'] = timeit(g_h)
    rec['forward_sorted_ms'] = rec['sort_ms'] + rec['exchange_ms'] + rec['paint_sorted_ms'] + rec['readout3_sorted_scatter_ms']
    rec['forward_home_ms'] = rec['paint_home_atomic_ms'] + rec['readout3_home_ms']
    rec['reverse_sorted_ms'] = rec['exchange_f_ms'] + rec['paint3_sorted_ms'] + rec['grad4_sorted_scatter_ms']
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) and v > 1e-6 else v) for k, v in rec.items()}), flush=True)
    del lay, xl, fl, m3
