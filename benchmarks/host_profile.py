"""cProfile of one forward+vjp bench step (host side), after warm-up."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import torch
import bench
from vmad_b200.pm import DeviceArray, ParticleMesh

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pm = ParticleMesh([N] * 3, 100.0 * N / 32)
wn_h, cot_h = bench.make_inputs(N)
q = pm.generate_uniform_particle_grid(shift=0.0)
model = bench.build_model(pm, q, numpy.linspace(0.1, 1.0, nsteps + 1))
wn = pm.create('complex', value=wn_h)
cot = DeviceArray.from_host(cot_h)


def step():
    r = model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot))
    torch.cuda.synchronize()
    return r


for _ in range(3):
    step()
t0 = time.perf_counter()
step()
print("wall per step: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
# GPU-only time: enqueue everything, measure with events
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(35)
st.sort_stats('cumulative').print_stats(45)
