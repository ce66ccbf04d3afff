import os, sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
import numpy, torch, torch.distributed as dist
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
def step(): return model.compute_with_vjp(init=dict(x=wn), v=dict(_dx=cot))
for _ in range(3): step()
torch.cuda.synchronize(); dist.barrier()
for i in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); h0 = time.perf_counter(); step(); h1 = time.perf_counter(); e1.record(); e1.synchronize()
    if rank == 0: print('host %.1f ms event %.1f ms' % ((h1 - h0) * 1e3, e0.elapsed_time(e1)), flush=True)
if rank == 0:
    pr = cProfile.Profile(); pr.enable()
step(); torch.cuda.synchronize()
if rank == 0:
    pr.disable(); pstats.Stats(pr).sort_stats('tottime').print_stats(18)
dist.barrier(); dist.destroy_process_group()
