"""CUDA-graph replay of the whole forward+vjp vs the eager run: timing and agreement."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch, bench
from vmad_b200.pm import DeviceArray, ParticleMesh
from vmad_b200.cudagraph import GraphedCall
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pm = ParticleMesh([N] * 3, 100.0 * N / 32)
wn_h, cot_h = bench.make_inputs(N)
q = pm.generate_uniform_particle_grid(shift=0.0)
model = bench.build_model(pm, q, numpy.linspace(0.1, 1.0, 11))
wn = pm.create('complex', value=wn_h)
cot = DeviceArray.from_host(cot_h)
def step():
    (dx,), (g,) = model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot))
    return dx, g
for _ in range(3): dx0, g0 = step()
torch.cuda.synchronize()
dx0 = dx0.numpy().copy(); g0 = g0.numpy().copy()
gs = GraphedCall(step)
def timeit(fn, n=10):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return numpy.median(ts), min(ts)
print("eager  ms:", timeit(step))
print("graph  ms:", timeit(gs))
dx1, g1 = gs()
torch.cuda.synchronize()
rel = lambda a, b: float(numpy.linalg.norm(a - b) / numpy.linalg.norm(b))
print("graph vs eager: dx %.2e grad %.2e" % (rel(dx1.numpy(), dx0), rel(g1.numpy(), g0)))
# new input values through the same buffers
wn2 = numpy.roll(wn_h, 3, axis=0)
wn.t.copy_(torch.from_numpy(wn2).to('cuda'))
dx2, g2 = gs(); torch.cuda.synchronize(); a = dx2.numpy().copy(); b = g2.numpy().copy()
dx3, g3 = step(); torch.cuda.synchronize()
print("new values, graph vs eager: dx %.2e grad %.2e" % (rel(a, dx3.numpy()), rel(b, g3.numpy())))
