import os, sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
import numpy, torch, bench
from vmad_b200.pm import DeviceArray, ParticleMesh
N=128
pm = ParticleMesh([N]*3, 100.0*N/32)
wn_h, cot_h = bench.make_inputs(N)
q = pm.generate_uniform_particle_grid(shift=0.0)
model = bench.build_model(pm, q, numpy.linspace(0.1,1.0,11))
wn = pm.create('complex', value=wn_h); cot = DeviceArray.from_host(cot_h)
def step(): return model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot))
for i in range(8):
    t0=time.perf_counter()
    if i == 4:
        pr = cProfile.Profile(); pr.enable()
    step(); torch.cuda.synchronize()
    if i == 4:
        pr.disable(); pstats.Stats(pr).sort_stats('tottime').print_stats(12)
    print(i, '%.1f ms' % ((time.perf_counter()-t0)*1e3), 'alloc %.2f GB reserved %.2f GB' % (torch.cuda.memory_allocated()/1e9, torch.cuda.memory_reserved()/1e9))
