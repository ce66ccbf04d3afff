"""Reductions over ranks as differentiable operators (the two operators of vmad/lib/mpi.py).

A value is either *universal* (identical on every rank) or *distributed* (every rank holds its
own part).  ``allreduce`` turns distributed into universal by summing over ranks; ``allbcast``
is the formal inverse, universal -> distributed, which does nothing in the forward direction.
They are each other's adjoints, so the vjp of one is the other's forward action.

``comm`` is duck-typed: ``allreduce(x)`` and ``barrier()`` are all that is asked of it
(``vmad_b200.pm.SelfComm`` for one rank, ``vmad_b200.dist.TorchComm`` over torch.distributed).
"""
from ..vm import operator


def _sum_over_ranks(x, comm):
    return comm.allreduce(x)


@operator
class allreduce:
    ain = [('x', '*')]
    aout = [('y', '*')]

    def apl(node, x, comm):
        return dict(y=_sum_over_ranks(x, comm))

    def vjp(node, _y, comm):
        # every rank's part enters the sum with weight one
        return dict(_x=_y)

    def jvp(node, x_, comm):
        return dict(y_=_sum_over_ranks(x_, comm))


@operator
class allbcast:
    ain = [('x', '*')]
    aout = [('y', '*')]

    def apl(node, x, comm):
        comm.barrier()  # the value is already the same everywhere; just line the ranks up
        return dict(y=x)

    def vjp(node, _y, comm):
        # the universal value feeds every rank's part: its cotangent collects all of them
        return dict(_x=_sum_over_ranks(_y, comm))

    def jvp(node, x_):
        return dict(y_=x_)
