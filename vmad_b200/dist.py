"""Multi-GPU plumbing: a ``comm`` shim over torch.distributed and the slab bookkeeping.

One process per GPU (torchrun); NCCL over NVLink on the GPUs, gloo on CPU for the host-logic
tests.  The shim offers what vmad expects of ``pm.comm`` (``rank``, ``size``, ``allreduce``,
``barrier``: fastpm.py:196,251,422,596; lib/mpi.py:24,30,53,57; chisquare.py:32) plus the two
exchange patterns that replace pmesh's MPI calls:

* ``alltoallv_rows``  -- rows of a ``[n, ...]`` array, variable counts per peer
  (``Layout.exchange`` / ``Layout.gather``: MPI_Alltoallv inside pmesh's domain module);
* ``alltoall_blocks`` -- equal blocks ``[P, ...]`` (the transpose inside the slab FFT).

Both are single ``torch.distributed.all_to_all_single`` calls (NCCL grouped send/recv).
"""
import numpy
import torch
import torch.distributed as dist


def slab(n, size, rank):
    """(start, count) of the planes of an axis of length n owned by ``rank``"""
    if n % size:
        raise ValueError("mesh axis of length %d does not divide over %d ranks" % (n, size))
    c = n // size
    return rank * c, c


class TorchComm(object):
    def __init__(self, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        if device is None:
            backend = dist.get_backend(group)
            device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' \
                else torch.device('cpu')
        self.device = device
        # small host-side exchanges (the per-peer particle counts of a decompose) go through a
        # CPU group when the main group is NCCL: no device tensor, no stream synchronisation
        self._cpu_group = None
        if self.device.type == 'cuda' and self.size > 1:
            try:
                self._cpu_group = dist.new_group(backend='gloo')
            except Exception:
                self._cpu_group = None

    # -- what vmad asks of a communicator ---------------------------------------------------
    def allreduce(self, x, op=None):
        """sum over ranks of a python scalar, a complex, or a numpy array (returned as such)"""
        if isinstance(x, complex):
            r = self.allreduce(numpy.array([x.real, x.imag]))
            return complex(r[0], r[1])
        if isinstance(x, (int, numpy.integer)) and not isinstance(x, bool):
            t = torch.tensor([int(x)], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, group=self.group)
            return int(t.item())
        if isinstance(x, (float, numpy.floating)):
            t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, group=self.group)
            return float(t.item())
        a = numpy.ascontiguousarray(x)
        t = torch.from_numpy(a.copy()).to(self.device)
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def barrier(self):
        dist.barrier(group=self.group)

    # -- exchanges --------------------------------------------------------------------------
    def alltoall_counts(self, send_counts):
        if self._cpu_group is not None:
            s = torch.tensor([int(c) for c in send_counts], dtype=torch.int64)
            r = torch.empty_like(s)
            dist.all_to_all_single(r, s, group=self._cpu_group)
            return [int(v) for v in r.tolist()]
        s = torch.tensor([int(c) for c in send_counts], dtype=torch.int64, device=self.device)
        r = torch.empty_like(s)
        dist.all_to_all_single(r, s, group=self.group)
        return [int(v) for v in r.tolist()]

    def alltoallv_rows(self, send, send_counts, recv_counts):
        """send[:send_counts[0]] goes to rank 0, the next send_counts[1] rows to rank 1, ...;
        returns the rows received, ordered by source rank"""
        send = send.contiguous()
        out = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype,
                          device=send.device)
        dist.all_to_all_single(out, send, output_split_sizes=[int(c) for c in recv_counts],
                               input_split_sizes=[int(c) for c in send_counts], group=self.group)
        return out

    def alltoall_blocks(self, send):
        """send[d] goes to rank d; returns out with out[s] = what rank s sent here"""
        assert send.shape[0] == self.size
        send = send.contiguous()
        cplx = send.is_complex()
        s = torch.view_as_real(send) if cplx else send
        out = torch.empty_like(s)
        dist.all_to_all_single(out, s, group=self.group)
        return torch.view_as_complex(out) if cplx else out
