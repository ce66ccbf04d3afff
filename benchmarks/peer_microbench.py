"""step-by-step check of the peer exchange primitives (2+ GPUs, torchrun)"""
import os, sys, ctypes
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from vmad_b200.dist import TorchComm, PeerExchange
from vmad_b200 import _capi as C
from vmad_b200.pm import runtime
comm = TorchComm()
P, rank = comm.size, comm.rank
def say(*a):
    print('[%d]' % rank, *a, flush=True)
px = PeerExchange(comm, 1 << 20)
say('opened', [hex(x) for x in px.buf_ptrs[0][:P]], [hex(x) for x in px.data_flag_ptrs[:P]])
rt = runtime()
e, b = px.begin()
say('begin', e, b)
px.put(e, torch.zeros(2, dtype=torch.float64, device='cuda'), 0, 0, 0, 0, 0); rt.synchronize(); say('empty put signalled')
px.commit(e); rt.synchronize(); say('waited')
px.release(e); rt.synchronize(); say('released')
n = 1024
src = torch.arange(P * n * 2, dtype=torch.float64, device='cuda') + 1000 * rank
e, b = px.begin()
px.put(e, src, 1, n, 0, n, rank * n); rt.synchronize(); say('put done')
px.commit(e); rt.synchronize(); say('committed')
out = torch.empty(P * n * 2, dtype=torch.float64, device='cuda')
C.check(C.lib.vpm_axpby(rt.ctx, P * n * 2, ctypes.c_double(1.0), px.recv_ptr(b), ctypes.c_double(0.0), None, ctypes.c_double(0.0), ctypes.c_void_p(out.data_ptr())))
rt.synchronize()
px.release(e)
exp = torch.cat([torch.arange(rank * n * 2, (rank + 1) * n * 2, dtype=torch.float64) + 1000 * s for s in range(P)])
say('recv ok:', bool((out.cpu() == exp).all()))
for it in range(20):
    e, b = px.begin(); px.put(e, src, 1, n, 0, n, rank * n); px.commit(e); px.release(e)
rt.synchronize(); say('20 rounds ok')
# ---- timing: full protocol per transfer, transpose-shaped payloads of growing size vs empty
px.ensure(1200 << 20)
def loop(big, nrows, rows, inner, iters):
    comm.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        e, b = px.begin(); px.put(e, big, nrows, inner, P * inner, inner, rank * rows * inner); px.commit(e); px.release(e)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for rows, inner, iters in ((128, (128 // P) * 65, 100), (256, (256 // P) * 129, 50), (512, (512 // P) * 65 * 2, 30), (512, (512 // P) * 65 * 4, 20)):
    big = torch.randn(rows * P * inner * 2, dtype=torch.float64, device='cuda')
    t0 = loop(big, 0, rows, inner, iters)
    t = loop(big, rows, rows, inner, iters)
    mb = rows * inner * 16 / 1e6
    if rank == 0:
        say('%7.1f MB to each of %d peers (+ the same kept locally): %8.1f us per transfer (empty: %.1f us) -> %.0f GB/s of '
            'remote stores per GPU' % (mb, P - 1, t, t0, mb * (P - 1) * 1e6 / ((t - t0) * 1e-6) / 1e9))
    del big
px.close(); say('closed')
dist.destroy_process_group()
