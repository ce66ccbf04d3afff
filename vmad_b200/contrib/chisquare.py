"""Chi-square problems on top of a forward operator (mirror of vmad/contrib/chisquare.py).

``MPIChiSquareProblem`` wraps ``y = sum_i |R_i(F(s))|^2`` -- F a forward autooperator (e.g. a
FastPM simulation), R_i residual operators -- into objective / gradient / Gauss-Newton
Hessian-vector product callables (chisquare.py:34-99).  The reference derives from
``abopt.abopt2.Problem`` / ``VectorSpace`` (an optimiser package that is not part of vmad and
is absent here); when abopt is importable those bases are used, otherwise two minimal
stand-ins with the same constructor arguments.  The optimisers themselves are out of scope.
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


class EpochScheduler(dict):
    """chisquare.py:101-151 -- per-epoch hyper-parameter schedule"""

    def __call__(self, argname, values):
        self.schedule(argname, values)

    @property
    def min_nepochs(self):
        return max(list(self.keys()) + [0]) + 1

    def schedule(self, argname, values):
        if isinstance(values, (list, tuple)):
            items = enumerate(values)
        elif isinstance(values, dict):
            items = values.items()
        else:
            items = enumerate([values])
        for epoch, value in items:
            args = self.get(epoch, {})
            args[argname] = value
            self[epoch] = args

    def get_args(self, epoch):
        args = {}
        for e in sorted(self.keys()):
            if e <= epoch:
                args.update(self[e])
        return args


class Epoch(object):
    def __init__(self, epoch, nepochs):
        self.epoch = epoch
        self.nepochs = nepochs

    def __repr__(self):
        return "Epoch %d / %d" % (self.epoch, self.nepochs)


class MAPInversion(object):
    """chisquare.py:160-246 -- multi-epoch MAP reconstruction driver around an optimiser object
    that offers ``minimize(problem, x0, monitor=)`` (abopt's; not provided here)"""

    def __init__(self, optimizer, ForwardOperator):
        self.optimizer = optimizer
        self.ForwardOperator = ForwardOperator
        self.schedule_optimizer = EpochScheduler()
        self.schedule_problem = EpochScheduler()

    @classmethod
    def create_like(kls, other, epochs=None):
        self = object.__new__(kls)
        self.__dict__.update(other.__dict__)
        self.schedule_optimizer = EpochScheduler()
        self.schedule_problem = EpochScheduler()
        if epochs is None:
            self.schedule_optimizer.update(other.schedule_optimizer)
            self.schedule_problem.update(other.schedule_problem)
        else:
            for i, epoch in enumerate(epochs):
                self.schedule_optimizer[i] = other.schedule_optimizer.get_args(epoch)
                self.schedule_problem[i] = other.schedule_problem.get_args(epoch)
        return self

    def apply(self, d, s0, epochs=None, monitor_epoch=None, monitor_progress=None):
        if epochs is None:
            epochs = range(self.nepochs)
        s1 = s0
        for epoch in epochs:
            if monitor_epoch:
                monitor_epoch(Epoch(epoch, max(epochs) + 1))
            problem = self.get_problem(d, epoch)
            self.optimizer.__dict__.update(self.schedule_optimizer.get_args(epoch))
            state = self.optimizer.minimize(problem, s1, monitor=monitor_progress)
            s1 = state['x']
        return s1

    @property
    def nepochs(self):
        return max(self.schedule_optimizer.min_nepochs, self.schedule_problem.min_nepochs)

    def get_problem(self, d, epoch=0):
        return self.problem_factory(d=d, **self.schedule_problem.get_args(epoch))

    def problem_factory(self, d, **args):
        raise NotImplementedError
