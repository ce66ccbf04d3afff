"""One sorted paint at 256^3 / 2^24 lattice+noise particles (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmad_b200.pm import ParticleMesh, DeviceArray
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = 100.0 * N / 32
pm = ParticleMesh([N] * 3, L)
g = torch.Generator(device='cuda').manual_seed(1)
side = N
q = torch.stack(torch.meshgrid(*[torch.arange(side, device='cuda', dtype=torch.float64)] * 3, indexing='ij'), -1).reshape(-1, 3) * (L / side)
x = DeviceArray(q + torch.randn(q.shape, generator=g, device='cuda', dtype=torch.float64) * (1.5 * L / N))
lay = pm.decompose(x)
xs = lay.exchange(x)
for _ in range(3):
    f = pm.paint(xs, 1.0)
    v = f.readout(xs)
torch.cuda.synchronize()
print(float(f.sum()), len(xs))
