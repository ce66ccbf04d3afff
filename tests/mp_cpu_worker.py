"""Worker for tests/test_dist_cpu.py (gloo, world_size 2, CPU tensors): the host-side plumbing
of the multi-GPU path -- communicator shim, variable-count row exchange, block transpose."""
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    dist.init_process_group('gloo')
    from vmad_b200.dist import TorchComm, slab
    comm = TorchComm()
    r, P = comm.rank, comm.size
    assert P == 2
    # allreduce of scalars / arrays / complex; mpi operators of the library on top of it
    assert comm.allreduce(r + 1) == 3
    assert abs(comm.allreduce(0.5 * (r + 1)) - 1.5) < 1e-15
    assert comm.allreduce(complex(r, 1)) == complex(1, 2)
    assert numpy.array_equal(comm.allreduce(numpy.arange(3.0) * (r + 1)), numpy.arange(3.0) * 3)
    from vmad_b200 import Builder
    from vmad_b200.lib import mpi
    with Builder() as m:
        x = m.input('x')
        m.output(y=mpi.allreduce(mpi.allbcast(x, comm) * (r + 1.0), comm))
    (y,), (gx,) = m.compute_with_vjp(init=dict(x=2.0), v=dict(_y=1.0))
    assert y == 6.0 and gx == 3.0       # d/dx sum_r (r+1) x, all-reduced by allbcast.vjp
    # slab arithmetic
    assert slab(8, 2, r) == (4 * r, 4)
    try:
        slab(7, 2, r)
        raise AssertionError("uneven slabs must be refused")
    except ValueError:
        pass
    # variable-count row exchange (Layout.exchange / gather): rank s sends s + 1 + d rows to d
    send_counts = [r + 1 + d for d in range(P)]
    recv_counts = comm.alltoall_counts(send_counts)
    assert recv_counts == [s + 1 + r for s in range(P)]
    # the whole count matrix (shared-memory board on one node; repeated rounds reuse the tables)
    for it in range(5):
        m = comm.allgather_counts([c + it for c in send_counts])
        assert m.shape == (P, P) and m.tolist() == [[s + 1 + d + it for d in range(P)] for s in range(P)]
    assert comm._board, "single-node ranks should use the host board"
    rows = torch.cat([torch.full((c, 3), 10.0 * r + d, dtype=torch.float64)
                      for d, c in enumerate(send_counts)])
    got = comm.alltoallv_rows(rows, send_counts, recv_counts)
    expect = torch.cat([torch.full((s + 1 + r, 3), 10.0 * s + r, dtype=torch.float64) for s in range(P)])
    assert torch.equal(got, expect)
    back = comm.alltoallv_rows(got, recv_counts, send_counts)       # the reverse route (gather)
    assert torch.equal(back, rows)
    # block transpose of the slab FFT: global complex array [N0, N1, K], rows sharded -> columns sharded
    N0, N1, K = 4, 6, 3
    rng = numpy.random.default_rng(0)
    full = rng.standard_normal((N0, N1, K)) + 1j * rng.standard_normal((N0, N1, K))
    n0l, n1l = N0 // P, N1 // P
    mine = torch.from_numpy(full[r * n0l:(r + 1) * n0l])                      # [n0l, N1, K]
    send = mine.reshape(n0l, P, n1l, K).permute(1, 0, 2, 3).contiguous()      # pack: [P, n0l, n1l, K]
    recv = comm.alltoall_blocks(send)                                         # [P, n0l, n1l, K]
    cols = recv.reshape(N0, n1l, K).numpy()
    assert numpy.array_equal(cols, full[:, r * n1l:(r + 1) * n1l])
    back = comm.alltoall_blocks(torch.from_numpy(cols).reshape(P, n0l, n1l, K))
    rowsb = back.permute(1, 0, 2, 3).reshape(n0l, N1, K).numpy()              # unpack
    assert numpy.array_equal(rowsb, full[r * n0l:(r + 1) * n0l])
    comm.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % r)


if __name__ == '__main__':
    main()
