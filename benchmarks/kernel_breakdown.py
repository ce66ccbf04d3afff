"""Per-kernel GPU time of one forward+vjp bench step, from CUPTI (torch.profiler)."""
import os, sys, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch, bench
from torch.profiler import profile, ProfilerActivity
from vmad_b200.pm import DeviceArray, ParticleMesh

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pm = ParticleMesh([N] * 3, 100.0 * N / 32)
pm.force_recompute = N >= 512
wn_h, cot_h = bench.make_inputs(N)
q = pm.generate_uniform_particle_grid(shift=0.0)
model = bench.build_model(pm, q, numpy.linspace(0.1, 1.0, 11))
wn = pm.create('complex', value=wn_h)
cot = DeviceArray.from_host(cot_h)
del wn_h, cot_h


def step():
    r = model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot))
    torch.cuda.synchronize()
    return r


for _ in range(2):
    step()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r'\(.*', '', ev.name).replace('void ', '').replace('vpm::', '')[:60]
        a = agg[name]
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
tot = sum(v for _, v in agg.values())
print("N=%d total GPU kernel time %.2f ms" % (N, tot / 1e3))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-62s n=%4d %9.3f ms %5.1f%%" % (k, n, v / 1e3, 100 * v / tot))
