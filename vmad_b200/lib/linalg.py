"""The part of vmad/lib/linalg.py that the particle-mesh path uses.

``stack`` (linalg.py:139-149; used at fastpm.py:349,437,611), ``to_scalar`` (:114-126), ``sum``
(:255-277), ``take`` (:151-178), ``copy``, ``reshape``, ``broadcast_to`` and the re-exported
stdlib arithmetic.  The generic numpy ufunc / einsum operators of the reference are outside the
hot path and are not provided.

Values may be numpy arrays or device arrays: the numpy functions used here (``numpy.stack``,
``numpy.take``, ``numpy.sum``) dispatch to the device through ``DeviceArray.__array_function__``.
"""
import numpy

from ..vm import operator
from ..vm.std import abs, add, div, mod, mul, pow  # noqa: F401  (same re-exports as the reference)


def _on_device(x):
    return hasattr(x, 't') and hasattr(x, '_cellindex')


@operator
class to_scalar:
    """``y = sum |x|^2``"""
    ain = {'x': 'ndarray'}
    aout = {'y': '*'}

    def apl(node, x):
        if _on_device(x):
            return dict(y=x.dot(x))
        return dict(y=(x * numpy.conj(x)).sum())

    def vjp(node, _y, x):
        return dict(_x=x * (2 * numpy.conj(_y)))

    def jvp(node, x_, x):
        if _on_device(x):
            return dict(y_=2 * x.dot(x_))
        return dict(y_=(x_ * numpy.conj(x) + numpy.conj(x_) * x).sum())


@operator
class copy:
    ain = {'x': 'ndarray'}
    aout = {'y': 'ndarray'}

    def apl(node, x):
        return dict(y=numpy.copy(x))

    def vjp(node, _y):
        return dict(_x=numpy.copy(_y))

    def jvp(node, x_):
        return dict(y_=numpy.copy(x_))


@operator
class stack:
    """``numpy.stack`` of a list of arrays; the vjp un-stacks with ``numpy.take``"""
    ain = {'x': 'ndarray'}
    aout = {'y': 'ndarray'}

    def apl(node, x, axis):
        return dict(y=numpy.stack(x, axis=axis))

    def vjp(node, _y, axis):
        n = numpy.shape(_y)[axis]
        return dict(_x=[numpy.take(_y, i, axis=axis) for i in range(n)])

    def jvp(node, x_, axis, x):
        zero = [xi_ for xi_ in x_ if not (isinstance(xi_, (int, float)) and xi_ == 0)]
        if len(zero) != len(x_):
            # a missing tangent is a zero array of the matching shape
            x_ = [xi * 0 if (isinstance(xi_, (int, float)) and xi_ == 0) else xi_
                  for xi, xi_ in zip(x, x_)]
        return dict(y_=numpy.stack(x_, axis=axis))


@operator
class take:
    ain = {'x': 'ndarray'}
    aout = {'y': 'ndarray'}

    def apl(node, x, i, axis):
        if axis is None:
            raise AssertionError('axis keyword in linalg.take cannot be None.')
        if _on_device(x):
            y = numpy.take(x, i, axis=axis)
            return dict(y=y, yshape=y.shape, xshape=x.shape, i=i, axis=axis)
        i = numpy.array(i, dtype='intp')
        y = numpy.take(x, i, axis=axis)
        return dict(y=y, yshape=numpy.shape(y), xshape=numpy.shape(x), i=i, axis=axis)

    def vjp(node, _y, i, axis, xshape, yshape):
        _y = numpy.broadcast_to(numpy.asarray(_y), yshape)
        _x = numpy.zeros(xshape)
        index = [slice(None)] * len(xshape)
        index[axis] = i
        numpy.add.at(_x, tuple(index), _y)
        return dict(_x=_x)

    def jvp(node, x_, i, axis):
        return dict(y_=numpy.take(x_, i, axis=axis))


@operator
class reshape:
    ain = {'x': '*'}
    aout = {'y': '*'}

    def apl(node, x, shape):
        return dict(y=numpy.reshape(x, shape))

    def rcd(node, x, shape, y):
        return dict(xshape=numpy.shape(x), shape=shape)

    def vjp(node, _y, xshape):
        return dict(_x=numpy.reshape(_y, xshape))

    def jvp(node, x_, shape):
        return dict(y_=numpy.reshape(x_, shape))


@operator
class broadcast_to:
    ain = 'x'
    aout = 'y'

    def apl(node, x, shape):
        return dict(y=numpy.broadcast_to(x, shape), xshape=numpy.shape(x), yshape=shape)

    def vjp(node, _y, xshape, yshape):
        _x = numpy.broadcast_to(numpy.asarray(_y), yshape)
        extra = len(yshape) - len(xshape)
        if extra:
            _x = numpy.sum(_x, axis=tuple(range(extra)))
        for axis, size in enumerate(xshape):
            if size == 1:
                _x = numpy.sum(_x, axis=axis, keepdims=True)
        return dict(_x=_x)

    def jvp(node, x_, yshape):
        return dict(y_=numpy.broadcast_to(x_, yshape))


@operator
class sum:
    ain = {'x': '*'}
    aout = {'y': '*'}

    def apl(node, x, axis=None):
        if _on_device(x):
            return dict(y=x.sum(axis=axis))
        return dict(y=numpy.sum(x, axis=axis, dtype='f8'))

    def rcd(node, x, y, axis=None):
        if _on_device(x):
            return dict(xshape=x.shape, axis=axis, like=x)
        return dict(xshape=numpy.shape(x), axis=axis, like=None)

    def vjp(node, _y, xshape, axis, like):
        if like is not None:
            # d(sum x)/dx = 1: a constant array on the device (like * 0 + _y in one pass)
            return dict(_x=like._axpby(0.0, c=float(_y)))
        _x = numpy.ones(xshape)
        if axis is not None:
            shape = list(numpy.shape(_y))
            shape.insert(axis, 1)
            _y = numpy.reshape(_y, shape)
        return dict(_x=_x * _y)

    def jvp(node, x_, axis, like):
        if _on_device(x_):
            return dict(y_=x_.sum(axis=axis))
        return dict(y_=numpy.sum(x_, axis=axis, dtype='f8'))
