"""Standard operators of the VM: arithmetic on symbols, ``terminal``, and the debugging helpers
(behaviour of vmad/core/stdlib/operators.py:14-152, utils.py:9-86, __init__.py:4-13)."""
from builtins import abs as _abs
from builtins import eval as _eval

from .ops import ZeroGradient, operator


@operator
class terminal:
    """marks a model output"""
    ain = 'x'
    aout = 'y'

    def apl(node, x):
        return dict(y=x)

    def vjp(node, _y):
        return dict(_x=_y)

    def jvp(node, x_):
        return dict(y_=x_)


def _linear_unary(name, f):
    """y = f(x) with f linear: the vjp and jvp are f itself"""
    def apl(node, x):
        return dict(y=f(x))

    def vjp(node, _y):
        return dict(_x=f(_y))

    def jvp(node, x_):
        return dict(y_=f(x_))

    return operator(type(name, (), dict(ain='x', aout='y', apl=apl, vjp=vjp, jvp=jvp)))


def _linear_binary(name, f, sign2):
    """y = f(x1, x2) = x1 + sign2 * x2"""
    def apl(node, x1, x2):
        return dict(y=f(x1, x2))

    def vjp(node, _y):
        return dict(_x1=_y, _x2=_y if sign2 > 0 else -_y)

    def jvp(node, x1_, x2_):
        return dict(y_=f(x1_, x2_))

    return operator(type(name, (), dict(ain=('x1', 'x2'), aout='y', apl=apl, vjp=vjp, jvp=jvp)))


neg = _linear_unary('neg', lambda x: -x)
pos = _linear_unary('pos', lambda x: +x)
add = _linear_binary('add', lambda a, b: a + b, +1)
sub = _linear_binary('sub', lambda a, b: a - b, -1)


@operator
class abs:
    ain = 'x'
    aout = 'y'

    def apl(node, x):
        return dict(y=_abs(x))

    def vjp(node, _y, x):
        return dict(_x=_y * (x > 0) + -_y * (x < 0))

    def jvp(node, x_, x):
        return dict(y_=x_ * (x > 0) + -x_ * (x < 0))


@operator
class mul:
    ain = 'x1', 'x2'
    aout = 'y'

    def apl(node, x1, x2):
        return dict(y=x1 * x2)

    def rcd(node, x1, x2, y):
        # the partner of a literal factor is never needed: do not pin it on the tape
        if node.is_literal('x1'):
            x2 = 0
        if node.is_literal('x2'):
            x1 = 0
        return dict(x1=x1, x2=x2)

    def vjp(node, _y, x1, x2):
        return dict(_x1=_y * x2, _x2=_y * x1)

    def jvp(node, x1_, x2_, x1, x2):
        return dict(y_=x1_ * x2 + x1 * x2_)


@operator
class div:
    ain = 'x1', 'x2'
    aout = 'y'

    def apl(node, x1, x2):
        x2inv = 1 / x2
        return dict(y=x1 * x2inv, x2inv=x2inv)

    def rcd(node, x1, x2inv, y, x2):
        x1fac = 0 if node.is_literal('x1') else x2inv
        x2fac = 0 if node.is_literal('x2') else x1 * x2inv * x2inv
        return dict(x1fac=x1fac, x2fac=x2fac)

    def vjp(node, _y, x1fac, x2fac):
        return dict(_x1=_y * x1fac, _x2=-_y * x2fac)

    def jvp(node, x1_, x2_, x1fac, x2fac):
        return dict(y_=x1_ * x1fac - x2fac * x2_)


@operator
class mod:
    ain = 'x1', 'x2'
    aout = 'y'

    def apl(node, x1, x2):
        return dict(y=x1 % x2, n=x1 // x2)

    def vjp(node, _y, x1, x2, n):
        return dict(_x1=_y, _x2=(-n) * _y)

    def jvp(node, x1_, x2_, x1, x2, n):
        return dict(y_=x1_ - n * x2_)


@operator
class pow:
    ain = 'x1', 'x2'
    aout = 'y'

    def apl(node, x1, x2):
        if not node.is_literal('x2'):
            from numpy import log
            logx1 = log(x1)
        else:
            logx1 = 0
        return dict(y=x1 ** x2, logx1=logx1)

    def _fac(x1, x2):
        import numpy as np
        x2 = np.array(x2, dtype='f8')
        with np.errstate(divide='ignore'):
            fac = np.asarray(x1 ** (x2 - 1), dtype='f8').copy()
        fac[np.isinf(fac)] = 0
        return x2, fac

    def vjp(node, _y, x1, x2, logx1):
        x2, fac = pow.prototype._fac(x1, x2)
        return dict(_x1=x2 * _y * fac, _x2=_y * x1 ** x2 * logx1)

    def jvp(node, x1_, x2_, x1, x2, logx1):
        x2, fac = pow.prototype._fac(x1, x2)
        return dict(y_=x2 * x1_ * fac + x2_ * x1 ** x2 * logx1)


@operator
class eval:
    """evaluate ``function(x)`` (or the expression string, with ``x`` bound) at run time"""
    ain = 'x'
    aout = 'y'

    def apl(node, x, function):
        if callable(function):
            return dict(y=function(x))
        return dict(y=_eval(function, dict(x=x)))

    def vjp(node):
        return ZeroGradient

    def jvp(node):
        return ZeroGradient


@operator
class watchpoint:
    """call ``monitor(x)`` when the model reaches this point"""
    ain = {'x': '*'}
    aout = {}

    def apl(node, x, monitor=print):
        monitor(x)
        return dict(y=x)

    def vjp(node):
        return ZeroGradient

    def jvp(node):
        return


@operator
class assert_isinstance:
    ain = 'obj'
    aout = []

    def apl(node, obj, class_or_tuple):
        if not isinstance(obj, class_or_tuple):
            raise TypeError('Expecting an instance of %s, got %s'
                            % (repr(class_or_tuple), repr(type(obj))))

    def vjp(node):
        return ZeroGradient

    def jvp(node):
        return ZeroGradient


@operator
class assert_true:
    ain = 'x'
    aout = []

    def apl(node, x, func):
        if not func(x):
            raise AssertionError('Assertion failed on %s(%s).' % (func, x))

    def vjp(node):
        return ZeroGradient

    def jvp(node):
        return ZeroGradient
