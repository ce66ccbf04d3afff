"""Sweep the output-tile shape of the sorted paint kernel (256^3, 2^24 lattice+noise / uniform)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from vmad_b200.pm import ParticleMesh, DeviceArray
from vmad_b200 import _capi as C
from benchmarks.micro import timeit

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = 100.0 * N / 32
pm = ParticleMesh([N] * 3, L)
g = torch.Generator(device='cuda').manual_seed(1)
q = torch.stack(torch.meshgrid(*[torch.arange(N, device='cuda', dtype=torch.float64)] * 3, indexing='ij'), -1).reshape(-1, 3) * (L / N)
cases = {'lattice+1.5': q + torch.randn(q.shape, generator=g, device='cuda', dtype=torch.float64) * (1.5 * L / N),
         'lattice+0.1': q + torch.randn(q.shape, generator=g, device='cuda', dtype=torch.float64) * (0.1 * L / N)}
for name, xt in cases.items():
    x = DeviceArray(xt)
    lay = pm.decompose(x)
    xs = lay.exchange(x)
    n = len(xs)
    for tile in [(4, 8, 32), (4, 8, 64), (2, 8, 64), (4, 4, 64), (8, 8, 32), (4, 16, 32), (2, 16, 32), (4, 8, 128), (2, 8, 128), (8, 8, 16), (4, 8, 16)]:
        C.lib.vpm_paint_set_tile(*tile)
        med, best = timeit(lambda: pm.paint(xs, 1.0))
        print(json.dumps(dict(case=name, tile=tile, ms=med * 1e3, gpart_s=n / med / 1e9,
                              frac=(24.0 * n + 8.0 * N ** 3) / med / 1e9 / 6540.5)))
    C.lib.vpm_paint_set_tile(0, 0, 0)
