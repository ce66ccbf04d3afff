"""Bandwidth of the transpose puts of the slab FFT at the shapes of an nmesh^3 problem on P ranks
(torchrun): forward pattern (strided source rows -> packed blocks) and inverse pattern (packed
blocks -> rows with a destination stride), whole flag protocol included, CUDA events, max over ranks.

    torchrun --nproc-per-node 8 benchmarks/peer_put_bw.py 512
"""
import json
import os
import sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from vmad_b200.dist import TorchComm, PeerExchange  # noqa: E402
from vmad_b200.pm import runtime  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
comm = TorchComm()
P, rank = comm.size, comm.rank
rt = runtime()
n0l, n1l, Nh = N // P, N // P, N // 2 + 1
inner = n1l * Nh
blk = n0l * inner
px = PeerExchange(comm, 16 * P * blk)
a = torch.randn(2 * P * blk, dtype=torch.float64, device='cuda')


def timed(fn):
    for _ in range(3):
        fn()
    comm.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def fwd():      # r2c: [n0l, P, inner] -> block d packed in rank d's area
    _, b = px.begin()
    px.put(b, a, n0l, inner, P * inner, inner, rank * n0l * inner)


def inv():      # c2r: block d = planes of rank d -> rows with stride P * inner in d's area
    _, b = px.begin()
    px.put(b, a, n0l, inner, inner, blk, rank * inner, d_row=P * inner)


def inv_packed():   # c2r, round-1 form: contiguous block, unpack pass afterwards
    _, b = px.begin()
    px.put(b, a, 1, blk, 0, blk, rank * blk)


def empty():
    _, b = px.begin()
    px.put(b, a, 0, inner, 0, 0, 0)


t0 = timed(empty)
rows = []
for name, fn in (('forward (strided source)', fwd), ('inverse (strided destination)', inv), ('inverse (packed block)', inv_packed)):
    t = timed(fn)
    mb = blk * 16 / 1e6
    rows.append(dict(pattern=name, ms=t, empty_ms=t0, mb_per_peer=mb, remote_gbs=mb * (P - 1) / (t * 1e-3) / 1e3,
                     remote_gbs_net_of_protocol=mb * (P - 1) / ((t - t0) * 1e-3) / 1e3))
if rank == 0:
    print(json.dumps(dict(nmesh=N, ranks=P, iters=iters, rows=rows)))
px.close()
dist.destroy_process_group()
