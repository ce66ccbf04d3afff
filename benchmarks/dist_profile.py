"""CUPTI kernel totals + cProfile of one multi-GPU forward+vjp step (torchrun, >= 2 GPUs)."""
import os, sys, time, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import bench
from vmad_b200.dist import TorchComm
from vmad_b200.pm import DeviceArray, ParticleMesh
comm = TorchComm(); rank = comm.rank; world = comm.size
N = 128
dims = bench.global_shape(N, world)
pm = ParticleMesh(dims, [100.0 * n / 32 for n in dims], comm=comm)
wn_g, cot_g = bench.make_inputs(N, dims=dims)
q = pm.generate_uniform_particle_grid(shift=0.0); npl = len(q)
wn = pm.create('complex', value=wn_g)
cot = DeviceArray.from_host(cot_g[rank * npl:(rank + 1) * npl])
model = bench.build_model(pm, q, numpy.linspace(0.1, 1.0, 11))
def step():
    r = model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot)); torch.cuda.synchronize(); return r
for _ in range(3): step()
import gc
gc.collect(); gc.set_debug(gc.DEBUG_SAVEALL); step(); n = gc.collect()
if rank == 0:
    hist = collections.Counter(type(o).__name__ for o in gc.garbage)
    print("cyclic garbage after one step: %d objects" % n, hist.most_common(12))
    big = [o for o in gc.garbage if isinstance(o, torch.Tensor)]
    print("tensors in garbage:", [(tuple(t.shape), t.dtype) for t in big][:20])
gc.set_debug(0); gc.garbage.clear()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter(); step(); wall = time.perf_counter() - t0
if rank == 0:
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r'\(.*', '', ev.name).replace('void ', '').replace('vpm::', '')[:50]
            a = agg[name]; a[0] += 1; a[1] += ev.device_time
    tot = sum(v for _, v in agg.values())
    print("wall %.1f ms; total GPU kernel time %.2f ms" % (wall * 1e3, tot / 1e3))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("%-52s n=%4d %9.3f ms %5.1f%%" % (k, n, v / 1e3, 100 * v / tot))
    cpu = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CPU:
            a = cpu[ev.name[:50]]; a[0] += 1; a[1] += ev.cpu_time
    for k, (n, v) in sorted(cpu.items(), key=lambda kv: -kv[1][1])[:12]:
        print("CPU %-48s n=%4d %9.3f ms" % (k, n, v / 1e3))
import cProfile, pstats
dist.barrier()
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
if rank == 0:
    st = pstats.Stats(pr)
    st.sort_stats('tottime').print_stats(30)
    st.sort_stats('cumulative').print_stats(50)
dist.barrier(); dist.destroy_process_group()
