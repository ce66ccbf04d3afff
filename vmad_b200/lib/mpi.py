"""``allreduce`` / ``allbcast`` on a duck-typed communicator (vmad/lib/mpi.py:19-62).

universal   : a value replicated on all ranks;  distributed : a value partitioned over ranks.
``comm`` needs ``allreduce(x)`` and ``barrier()``; ``vmad_b200.pm.SelfComm`` (one rank) and
``vmad_b200.dist.TorchComm`` (torch.distributed, NCCL or gloo) both qualify.
"""
from ..vm import operator


@operator
class allreduce:
    """distributed -> universal sum; the vjp of a sum over ranks is the identity per rank"""
    ain = [('x', '*')]
    aout = [('y', '*')]

    def apl(node, x, comm):
        return comm.allreduce(x)

    def vjp(node, _y, comm):
        return dict(_x=_y)

    def jvp(node, x_, comm):
        return dict(y_=comm.allreduce(x_))


@operator
class allbcast:
    """universal -> distributed (a no-op forward); its vjp sums the cotangents over ranks"""
    ain = [('x', '*')]
    aout = [('y', '*')]

    def apl(node, x, comm):
        comm.barrier()
        return x

    def vjp(node, _y, comm):
        return dict(_x=comm.allreduce(_y))

    def jvp(node, x_):
        return dict(y_=x_)
