import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy, torch, bench
from vmad_b200.pm import DeviceArray, ParticleMesh
from vmad_b200 import _capi as C
N=128
pm = ParticleMesh([N]*3, 100.0*N/32)
wn_h, cot_h = bench.make_inputs(N)
q = pm.generate_uniform_particle_grid(shift=0.0)
model = bench.build_model(pm, q, numpy.linspace(0.1,1.0,11))
wn = pm.create('complex', value=wn_h); cot = DeviceArray.from_host(cot_h)
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
def step(): return model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot))
def run(tag, do_flush, sampler):
    s = bench.ClockSampler(0)
    if sampler: s.start()
    for _ in range(3): step()
    torch.cuda.synchronize()
    tot=0; w0=time.perf_counter()
    for _ in range(3):
        if do_flush: flush.fill_(1)
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); h0=time.perf_counter(); step(); h1=time.perf_counter(); e1.record(); e1.synchronize()
        tot+=e0.elapsed_time(e1)
        print(tag, 'host enqueue %.1f ms, event %.1f ms' % ((h1-h0)*1e3, e0.elapsed_time(e1)))
    if sampler: print(s.stop())
run('plain', False, False)
run('flush', True, False)
run('sampler', False, True)
run('plain2', False, False)
