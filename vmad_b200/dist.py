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


class HostBoard(object):
    """A P x (1 + width) int64 table in node-local shared memory (a file under /dev/shm mapped
    by every rank) for the tiny host-side all-gathers of a decompose (per-peer particle
    counts): a rank writes its row, then its sequence number; readers spin until all P sequence
    numbers arrived.  Two tables alternate so that a fast rank can post the next round while a
    slow one still reads.  Microseconds instead of a gloo collective; single node only (the
    stores are ordered as written on x86; the sequence number is written last)."""

    def __init__(self, comm, width=16):
        import mmap
        import os
        self.P, self.rank, self.width = comm.size, comm.rank, width
        nbytes = 8 * 2 * self.P * (1 + width)
        name = [None]
        if self.rank == 0:
            path = '/dev/shm/vmad_b200_board_%d_%x' % (os.getpid(), int.from_bytes(os.urandom(4), 'little'))
            with open(path, 'wb') as f:
                f.write(b'\0' * nbytes)
            name[0] = path
        dist.broadcast_object_list(name, src=dist.get_global_rank(comm.group, 0) if comm.group else 0,
                                   group=comm.group)
        self._f = open(name[0], 'r+b')
        self._map = mmap.mmap(self._f.fileno(), nbytes)
        self.table = numpy.frombuffer(self._map, dtype='i8').reshape(2, self.P, 1 + width)
        comm.barrier()
        if self.rank == 0:
            os.unlink(name[0])
        self.seq = 0

    def allgather(self, row):
        row = numpy.asarray(row, dtype='i8')
        n = len(row)
        assert n <= self.width
        self.seq += 1
        t = self.table[self.seq & 1]
        t[self.rank, 1:1 + n] = row
        t[self.rank, 0] = self.seq
        flags = t[:, 0]
        spins = 0
        import time
        t0 = None
        while (flags < self.seq).any():
            spins += 1
            if spins > 20000:       # ~10 ms of busy polling, then back off
                if t0 is None:
                    t0 = time.monotonic()
                time.sleep(50e-6)
                if time.monotonic() - t0 > 600.0:
                    raise RuntimeError("host board: a peer did not post its counts")
        return t[:, 1:1 + n].copy()


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
        # python scalars (particle counts, norms already reduced on the device) never touch the
        # GPU on one node: all-gather the 8-byte words through the host board, sum in rank order
        board = self._host_board() if isinstance(x, (int, float, numpy.integer, numpy.floating)) \
            and not isinstance(x, bool) else None
        if board:
            if isinstance(x, (int, numpy.integer)):
                return int(board.allgather([int(x)]).sum())
            bits = numpy.array([float(x)], dtype='f8').view('i8')
            return float(board.allgather(bits).view('f8').sum())
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
    def allgather_counts(self, row):
        """every rank's ``row`` (a short list of ints): [P, len(row)] int64 on the host"""
        row = [int(c) for c in row]
        if self.size == 1:
            return numpy.array([row], dtype='i8')
        board = self._host_board()
        if board and len(row) <= board.width:
            return board.allgather(row)
        group = self._cpu_group if self._cpu_group is not None else self.group
        dev = 'cpu' if self._cpu_group is not None else self.device
        mine = torch.tensor(row, dtype=torch.int64, device=dev)
        out = [torch.empty_like(mine) for _ in range(self.size)]
        dist.all_gather(out, mine, group=group)
        return numpy.array([o.tolist() for o in out], dtype='i8')

    def _host_board(self):
        if getattr(self, '_board', None) is None:
            self._board = False
            import platform
            # the board relies on stores becoming visible in program order (x86 TSO); on other
            # hosts (Grace / ARM) the counts go through the gloo group instead
            if self.size > 1 and platform.machine() in ('x86_64', 'AMD64') and self._single_node():
                try:
                    self._board = HostBoard(self)
                except OSError:
                    self._board = False
        return self._board

    def _single_node(self):
        import socket
        names = [None] * self.size
        dist.all_gather_object(names, socket.gethostname(), group=self.group)
        return len(set(names)) == 1

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


class PeerExchange(object):
    """Double-buffered receive areas + flag blocks of every rank, mapped into this process
    through CUDA IPC, for the peer-memory transposes of the slab FFT and the particle exchange
    (vmad_b200/csrc/peer.cu).  One per communicator; ``nbytes`` is the capacity of one
    receive area.

    Protocol of a transfer on receive buffer b, all inside the put kernel(s): acknowledge my
    previous transfer to every peer (its consumer precedes this launch on the stream) -> wait
    until every peer acknowledged the previous transfer that used buffer b (it has finished
    reading the area about to be overwritten) -> store blocks into the peers' areas -> last
    thread block raises my data flag everywhere and waits for all data flags here: when the
    kernel completes my receive area is complete and the consumer (next kernel on the stream)
    reads it in place.  The transfer numbers live in the flag block ON THE DEVICE; the host only
    alternates the buffer, so captured CUDA graphs containing transfers can be replayed.
    """

    def __init__(self, comm, nbytes):
        import ctypes
        from . import _capi as C
        self.C = C
        self.ctypes = ctypes
        self.comm = comm
        self.P, self.rank = comm.size, comm.rank
        if self.P > 8:
            raise ValueError("peer exchange supports up to 8 ranks")
        self._own, self._opened = [], []
        self.seq = 0
        self._allocate(nbytes)

    def _allocate(self, nbytes):
        """collective: (re)create the receive areas and flag blocks, exchange the IPC handles"""
        import ctypes
        C, comm = self.C, self.comm
        self.nbytes = int(nbytes)
        ctx = self._ctx()
        self._own, self._opened = [], []
        mine = []
        for size in (self.nbytes, self.nbytes, 8 * C.PEER_FLAG_WORDS):
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            C.check(C.lib.vpm_peer_alloc(ctx, size, ctypes.byref(ptr), handle))
            self._own.append(ptr.value)
            mine.append(handle.raw)
        C.check(C.lib.vpm_sync(ctx))
        everyone = [None] * self.P
        dist.all_gather_object(everyone, mine, group=comm.group)
        ptrs = [[0] * self.P for _ in range(3)]
        for d in range(self.P):
            for k in range(3):
                if d == self.rank:
                    ptrs[k][d] = self._own[k]
                else:
                    ptr = ctypes.c_void_p()
                    C.check(C.lib.vpm_peer_open(ctx, everyone[d][k], ctypes.byref(ptr)))
                    self._opened.append(ptr.value)
                    ptrs[k][d] = ptr.value
        A = ctypes.c_uint64 * 8
        pad = lambda v: A(*(list(v) + [0] * (8 - len(v))))
        self.buf_ptrs = [pad(ptrs[0]), pad(ptrs[1])]
        self.data_flag_ptrs = pad(ptrs[2])
        self.my_bufs = [self._own[0], self._own[1]]
        self.my_flags = self._own[2]
        comm.barrier()

    def close(self):
        C = self.C
        if self._own:
            ctx = self._ctx()
            C.lib.vpm_sync(ctx)
            self.comm.barrier()
            for p in self._opened:
                C.lib.vpm_peer_close(ctx, p)
            self.comm.barrier()
            for p in self._own:
                C.lib.vpm_peer_free(ctx, p)
            self._own, self._opened = [], []

    def _ctx(self):
        from .pm import runtime
        rt = runtime()
        rt.sync_stream()
        return rt.ctx

    def begin(self):
        """next transfer: (sequence number, receive buffer)"""
        self.seq += 1
        return self.seq, self.seq % 2

    def _sync(self, b, first, last):
        """flag protocol riding on a put launch (include/vmadpm.h: vpm_peer_sync)"""
        C, ct = self.C, self.ctypes
        sy = C.vpm_peer_sync()
        sy.peer_flags = ct.cast(self.data_flag_ptrs, ct.POINTER(ct.c_uint64))
        sy.my_flags = self.my_flags
        sy.rank = self.rank
        mode = C.PEER_WAIT_ACK
        if first:
            mode |= C.PEER_RAISE_ACK
        if last:
            mode |= C.PEER_SIGNAL | C.PEER_WAIT_DATA
        sy.mode = mode
        sy.buf = b
        return ct.byref(sy)

    def put(self, b, src, rows, inner, s_row, s_dst, dst_off, first=True, last=True, d_row=None):
        """block d of ``src`` (rows x inner complex elements, row stride s_row, block stride s_dst) into
        rank d's receive area at dst_off, rows ``d_row`` apart there (default: packed)"""
        C, ct = self.C, self.ctypes
        C.check(C.lib.vpm_peer_put_strided(self._ctx(), self.P, int(rows), int(inner), int(s_row), int(s_dst),
                                           int(inner if d_row is None else d_row),
                                           ct.c_void_p(src.data_ptr()), self.buf_ptrs[b], int(dst_off),
                                           self._sync(b, first, last)))

    def put_to(self, b, src_ptr, count, mask, dst_off, first, last):
        """``count`` contiguous complex128 elements at device address ``src_ptr`` to the ranks in
        ``mask`` (halo planes for the real-space stencils)"""
        C, ct = self.C, self.ctypes
        C.check(C.lib.vpm_peer_put_to(self._ctx(), self.P, int(mask), int(count), ct.c_void_p(int(src_ptr)),
                                      self.buf_ptrs[b], int(dst_off), self._sync(b, first, last)))

    def put_blocks(self, b, src_ptr, count, blocks):
        """``blocks`` = [(dest rank, source offset, destination offset)] in complex128 elements, each
        ``count`` elements long, all in one launch and one pass of the flag protocol"""
        C, ct = self.C, self.ctypes
        nb = len(blocks)
        dest = (ct.c_int32 * nb)(*[int(v[0]) for v in blocks])
        so = (ct.c_int64 * nb)(*[int(v[1]) for v in blocks])
        do = (ct.c_int64 * nb)(*[int(v[2]) for v in blocks])
        C.check(C.lib.vpm_peer_put_blocks(self._ctx(), self.P, nb, int(count), ct.c_void_p(int(src_ptr)), dest, so, do,
                                          self.buf_ptrs[b], self._sync(b, True, True)))

    def put_rows(self, b, src, idx, n, ncomp, idx_is_src, seg_start, dst_base, skip=-1, seg_count=None,
                 mail=None, sentinel=None):
        """particle rows to their destination ranks (vpm_peer_put_rows); ``seg_count`` / ``mail``
        (device int32 tensors) and ``sentinel`` (one inert position per destination rank) belong to
        the fixed-capacity routing"""
        C, ct = self.C, self.ctypes
        P = self.P
        seg = (ct.c_int32 * (P + 1))(*[int(v) for v in seg_start])
        base = (ct.c_int64 * P)(*[int(v) for v in dst_base])
        sent = (ct.c_double * (3 * P))(*[float(v) for row in sentinel for v in row]) if sentinel is not None else None
        C.check(C.lib.vpm_peer_put_rows(self._ctx(), P, int(n), int(ncomp), ct.c_void_p(src.data_ptr()),
                                        ct.c_void_p(idx.data_ptr()), int(idx_is_src), int(skip), seg, base,
                                        ct.c_void_p(seg_count.data_ptr()) if seg_count is not None else None,
                                        ct.c_void_p(mail.data_ptr()) if mail is not None else None, sent,
                                        self.buf_ptrs[b], self._sync(b, True, True)))

    def read_mail(self, b, out):
        """what the peers posted with the last transfer on buffer b -> ``out`` (device int32 [P])"""
        C, ct = self.C, self.ctypes
        C.check(C.lib.vpm_peer_read_mail(self._ctx(), self.P, ct.c_void_p(self.my_flags), int(b),
                                         ct.c_void_p(out.data_ptr())))

    def ensure(self, nbytes):
        """collective: grow the receive areas (every rank must call with the same value)"""
        if nbytes > self.nbytes:
            self.close()
            self._allocate(int(1.25 * nbytes))

    def recv_ptr(self, b, offset_elems=0):
        """device pointer (as c_void_p) to complex element ``offset_elems`` of my receive area b"""
        return self.ctypes.c_void_p(self.my_bufs[b] + 16 * int(offset_elems))
