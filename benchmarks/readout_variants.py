"""Ablation of the implementation variants of vpm_readout3 / vpm_readout_grad4 (csrc/readout.cu)
on cell-sorted particles with home-order (scattered) stores, the configuration of the FastPM step.

    python benchmarks/readout_variants.py 512 0.3 1.5
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
sigmas = [float(s) for s in sys.argv[2:]] or [0.3, 1.5]
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
VARIANTS = list(range(24)) + [b | (mb << 6) for mb in (5, 6, 7, 8) for b in (16, 20)] + [48 | (5 << 6), 48 | (4 << 6), 16 | (4 << 6)]
if os.environ.get('VARIANTS'):
    VARIANTS = [int(t) for t in os.environ['VARIANTS'].split(',')]
NAMES = {1: 'tile', 2: 'early', 4: 'split', 8: 'pair=aligned', 16: 'pair=8B', 32: 'stream'}
for sigma in sigmas:
    x = DeviceArray(q.t + rnd() * (sigma * L / N))
    p, fbar, base, ybar = rnd(), rnd(), rnd(), rnd()
    lay = pm.decompose(x)
    xl = lay.exchange(x)
    fl = lay.exchange(DeviceArray(fbar))
    yo = rt.empty((n, 3)); gg = rt.empty((n, 3)); y2 = rt.empty((n, 3))
    f3 = lambda: C.check(C.lib.vpm_readout3(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(xl.t), n,
                                            _ptr(lay.indices), None, _ptr(p), 0.5, _ptr(yo)))
    g4 = lambda: C.check(C.lib.vpm_readout_grad4(rt.ctx, mesh, _ptr(F[0]), _ptr(F[1]), _ptr(F[2]), _ptr(F[3]),
                                                 _ptr(xl.t), _ptr(fl.t), n, _ptr(lay.indices), _ptr(base), _ptr(gg),
                                                 _ptr(ybar), 0.25, _ptr(y2)))
    ref = None
    for v in VARIANTS:
        C.check(C.lib.vpm_readout_set_variant(v))
        f3(); g4()
        cur = (yo.clone(), gg.clone(), y2.clone())
        if ref is None:
            ref = cur
        same = all(bool(torch.equal(a, b)) for a, b in zip(ref, cur))
        rec = dict(nmesh=N, sigma_cells=sigma, variant=v,
                   bits=('+'.join(NAMES[b] for b in (1, 2, 4, 8, 16, 32) if v & b) or 'baseline') + (' minblocks=%d' % (v >> 6) if v >> 6 else ''),
                   readout3_ms=round(timeit(f3), 4), grad4_ms=round(timeit(g4), 4), bit_identical_to_variant0=same)
        print(json.dumps(rec), flush=True)
