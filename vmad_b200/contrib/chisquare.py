"""Chi-square problems on top of a forward operator (mirror of vmad/contrib/chisquare.py).

``MPIChiSquareProblem`` wraps ``y = sum_i |R_i(F(s))|^2`` -- F a forward autooperator (e.g. a
FastPM simulation), R_i residual operators -- into objective / gradient / Gauss-Newton
Hessian-vector product callables (chisquare.py:34-99).  The reference derives from
``abopt.abopt2.Problem`` / ``VectorSpace`` (an optimiser package that is not part of vmad and
is absent here); when abopt is importable those bases are used, otherwise two minimal
stand-ins with the same constructor arguments.  The optimisers themselves and the multi-epoch
driver scaffolding around them (``EpochScheduler``, ``MAPInversion``: chisquare.py:101-246,
pure bookkeeping over an abopt optimiser object) are out of scope.
"""
from ..vm import Builder, autooperator
from ..lib import linalg, mpi

try:  # pragma: no cover - abopt is not installed in this image
    from abopt.abopt2 import Problem as BaseProblem, VectorSpace
except Exception:
    class VectorSpace(object):
        def __init__(self, addmul=None, dot=None):
            if addmul is not None:
                self.addmul = addmul
            if dot is not None:
                self.dot = dot

    class BaseProblem(object):
        def __init__(self, objective, gradient, hessian_vector_product=None, vs=None, **kwargs):
            self.objective = objective
            self.gradient = gradient
            self.hessian_vector_product = hessian_vector_product
            self.vs = vs


@autooperator
class MPIChiSquareOperator:
    """chisquare.py:7-15 -- ``sum over ranks of sum(x * x)``"""
    ain = [('x', '*')]
    aout = [('y', '*')]

    def main(self, x, comm):
        chi2 = linalg.sum(x * x)
        chi2 = mpi.allreduce(chi2, comm)
        return dict(y=chi2)


def _local_sum(x):
    s = getattr(x, 'sum', None)
    if s is not None and hasattr(x, 't'):
        return s()
    import numpy
    return numpy.sum(x)


class MPIVectorSpace(VectorSpace):
    """chisquare.py:19-32"""

    def __init__(self, comm):
        self.comm = comm

    def addmul(self, a, b, c, p=1):
        """a + b * c ** p, following the type of b"""
        if p != 1:
            c = c ** p
        c = b * c
        if not (isinstance(a, (int, float)) and a == 0):
            c = c + a
        return c

    def dot(self, a, b):
        return self.comm.allreduce(_local_sum(a * b))


class MPIChiSquareProblem(BaseProblem):
    """chisquare.py:34-99"""

    def __init__(self, comm, forward_operator, residuals, vectorspace=None):
        self.residuals = residuals
        self.forward_operator = forward_operator
        self.comm = comm
        if vectorspace is None:
            vectorspace = MPIVectorSpace(comm)

        def outputs(fx):
            return fx if isinstance(fx, (list, tuple)) else [fx]

        with Builder() as m:
            x = m.input('x')
            y = 0
            fx = outputs(forward_operator(x))
            for residual in self.residuals:
                r = residual(*fx)
                y = linalg.add(y, MPIChiSquareOperator(r, comm))
            m.output(y=y)
        self.model = m

        def objective(x):
            return m.compute(vout='y', init=dict(x=x))

        def gradient(x):
            y, [vjp] = m.compute_with_vjp(init=dict(x=x), v=dict(_y=1.0))
            return vjp

        def hessian_vector_product(x, v):
            """Gauss-Newton: 2 J^T J v, J the Jacobian of the stacked residuals"""
            Dv = 0
            replay = forward_operator.precompute(x=x)
            for residual in self.residuals:
                with Builder() as mr:
                    xx = mr.input('x')
                    r = residual(*outputs(replay(xx)))
                    mr.output(y=r)
                y, [Dv1] = mr.compute_with_gnDp(vout='y', init=dict(x=x), v=dict(x_=v))
                Dv = Dv + Dv1
            return Dv * 2

        BaseProblem.__init__(self, vs=vectorspace, objective=objective, gradient=gradient,
                             hessian_vector_product=hessian_vector_product)
