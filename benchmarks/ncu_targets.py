"""One force evaluation + its reverse sweep (every hot kernel of a FastPM step once), bracketed by
cudaProfilerStart/Stop: the target of the per-kernel ``ncu --set full`` captures.

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/r2_force_256 python benchmarks/ncu_targets.py 256

Particles: the lattice displaced by a Gaussian of 1.5 cells (the occupancy statistics of a FastPM
step), in lattice (home) order, so that the cell sort, the exchange and the home-order stores see
the access pattern of the real step.
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmad_b200 import pm as vpm
from vmad_b200.pm import ParticleMesh, DeviceArray

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = 100.0 * N / 32
pm = ParticleMesh([N] * 3, L)
pm.force_recompute = False
g = torch.Generator(device='cuda').manual_seed(1)
q = pm.generate_uniform_particle_grid(shift=0.0)
n = len(q)
dx = DeviceArray(torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64) * (1.5 * L / N))
p = DeviceArray(torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64))
fbar = DeviceArray(torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64))
base = DeviceArray(torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64))
ybar = DeviceArray(torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64))


def once():
    dx1, x1 = vpm.drift_pos(pm, dx, p, 0.01, q)
    f, potk, st, p1 = vpm.force_apl(pm, dx1, q, x=x1, kick=(p, 0.5), want_f=False)
    gdx, yo = vpm.force_vjp(pm, st, fbar, 0, f_scale=0.5, base=base, update=(ybar, 0.25))
    return p1, gdx, yo


for _ in range(2):
    r = once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(reps):
    r = once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(N, n, float(r[0].t.sum()), float(r[1].t.sum()))
