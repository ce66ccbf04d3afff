"""Behaviour of the tape VM (vmad_b200.vm) on toy scalar operators; no backend needed."""
import gc

import numpy
import pytest

from vmad_b200 import Builder, autooperator, operator
from vmad_b200.vm import ZeroGradient
from vmad_b200.vm import stdlib


@operator
class scale:
    ain = 'x'
    aout = 'y'

    def apl(node, x, a):
        return dict(y=x * a)

    def vjp(node, _y, a):
        return dict(_x=_y * a)

    def jvp(node, x_, a):
        return dict(y_=x_ * a)


def test_compute_vjp_jvp_gndp():
    with Builder() as m:
        a = m.input('a')
        b = a * a + scale(a, 3.0)      # b = a^2 + 3a
        m.output(b=b)
    b = m.compute('b', init=dict(a=2.0))
    assert b == 10.0
    (b,), (ga,) = m.compute_with_vjp(init=dict(a=2.0), v=dict(_b=1.0))
    assert ga == 7.0
    b, b_ = m.compute_with_jvp('b', init=dict(a=2.0), v=dict(a_=1.0))
    assert b_ == 7.0
    b, (gn,) = m.compute_with_gnDp('b', init=dict(a=2.0), v=dict(a_=1.0))
    assert gn == 49.0


def test_zero_bypass_and_literals():
    calls = []

    @operator
    class probe:
        ain = 'x'
        aout = 'y'

        def apl(node, x):
            return x

        def vjp(node, _y):
            calls.append('vjp')
            return _y

        def jvp(node, x_):
            calls.append('jvp')
            return x_

    with Builder() as m:
        a = m.input('a')
        m.output(c=probe(a), d=probe(a))
    (c, d), tape = m.compute(['c', 'd'], init=dict(a=3), return_tape=True)
    _a = tape.get_vjp().compute('_a', init=dict(_c=ZeroGradient, _d=ZeroGradient))
    assert _a == 0 and calls == []
    _a = tape.get_vjp().compute('_a', init=dict(_c=2.0, _d=ZeroGradient))
    assert _a == 2.0 and calls == ['vjp']


def test_unused_nodes_do_not_run():
    @operator
    class boom:
        ain = 'x'
        aout = 'y'

        def apl(node, x):
            raise AssertionError("shall not run")

    with Builder() as m:
        a = m.input('a')
        boom(a)
        m.output(b=a + 1, c=boom(a))
    assert m.compute('b', init=dict(a=1)) == 2


def test_list_arguments_and_extra_record():
    @operator
    class total:
        ain = 'x'
        aout = 'y'

        def apl(node, x):
            return dict(y=sum(x), n=len(x))

        def vjp(node, _y, n):
            return dict(_x=[_y] * n)

        def jvp(node, x_):
            return dict(y_=sum(x_))

    with Builder() as m:
        a, b = m.input('a', 'b')
        m.output(y=total([a, b, a]))
    (y,), (ga, gb) = m.compute_with_vjp(init=dict(a=1.0, b=2.0), v=dict(_y=1.0))
    assert y == 4.0 and ga == 2.0 and gb == 1.0


def test_autooperator_bind_precompute_forgetful_and_memo():
    builds = []

    @autooperator('x->y')
    def poly(x, n):
        builds.append(n)
        y = x
        for _ in range(n - 1):
            y = y * x
        return dict(y=y)

    with Builder() as m:
        a = m.input('a')
        m.output(b=poly(a, 3) + poly(a, 3))
    (b,), (ga,) = m.compute_with_vjp(init=dict(a=2.0), v=dict(_b=1.0))
    assert b == 16.0 and ga == 24.0
    m.compute_with_vjp(init=dict(a=3.0), v=dict(_b=1.0))
    assert builds == [3]            # the inner model is built once per hyper-argument set
    bound = poly.bind(n=2)
    with Builder() as m2:
        a = m2.input('a')
        m2.output(b=bound(a))
    assert m2.compute('b', init=dict(a=5.0)) == 25.0
    replay = poly.precompute(x=2.0, n=2)
    with Builder() as m3:
        a = m3.input('a')
        m3.output(b=replay(a, 2))          # same signature as the original (as in the reference)
    assert m3.compute('b', init=dict(a=100.0)) == 4.0  # replays the recorded tape
    with Builder() as m4:
        a = m4.input('a')
        m4.output(b=poly.forgetful(a, 3))
    (b,), (ga,) = m4.compute_with_vjp(init=dict(a=2.0), v=dict(_b=1.0))
    assert b == 8.0 and ga == 12.0


def test_member_autooperator():
    class Thing(object):
        def __init__(self, k):
            self.k = k

        @autooperator('x->y')
        def times(self, x):
            return dict(y=x * self.k)

    t = Thing(4.0)
    with Builder() as m:
        a = m.input('a')
        m.output(b=t.times(a))
    (b,), (ga,) = m.compute_with_vjp(init=dict(a=2.0), v=dict(_b=1.0))
    assert b == 8.0 and ga == 4.0


def test_stdlib_arithmetic_gradients():
    with Builder() as m:
        a, b = m.input('a', 'b')
        m.output(y=(a - b) / b + (-a) * 2 + a ** 2)
    a0, b0 = 3.0, 2.0
    (y,), (ga, gb) = m.compute_with_vjp(init=dict(a=a0, b=b0), v=dict(_y=1.0))
    assert numpy.isclose(y, (a0 - b0) / b0 - 2 * a0 + a0 ** 2)
    assert numpy.isclose(ga, 1 / b0 - 2 + 2 * a0)
    assert numpy.isclose(gb, -a0 / b0 ** 2)
    assert stdlib.terminal is not None


def test_no_cyclic_garbage_per_step():
    """derived models pin device memory through their literals: they must die by refcount"""
    @autooperator('x->y')
    def inner(x):
        return dict(y=x * x + x)

    with Builder() as m:
        a = m.input('a')
        m.output(b=inner(inner(a)))
    m.compute_with_vjp(init=dict(a=1.5), v=dict(_b=1.0))
    gc.collect()
    gc.disable()
    try:
        r = m.compute_with_vjp(init=dict(a=1.5), v=dict(_b=1.0))
        m.compute_with_jvp('b', init=dict(a=1.5), v=dict(a_=1.0))
        del r
        assert gc.collect() == 0
    finally:
        gc.enable()


def test_errors():
    from vmad_b200.vm.errors import BadArgument, MissingArgument, UnexpectedOutput
    with Builder() as m:
        a = m.input('a')
        with pytest.raises(MissingArgument):
            scale(a)
        with pytest.raises(BadArgument):
            scale(a, 2.0, nothing=1)
        m.output(b=scale(a, 2.0))
    with pytest.raises(UnexpectedOutput):
        m.compute('c', init=dict(a=1))


def test_autooperator_memo_is_keyed_by_value():
    """a hyper-argument mutated between two calls must not replay the stale model (the
    reference rebuilds on every call, vmad/core/autooperator.py:133-156)"""
    import numpy

    class Cfg(object):
        def __init__(self, k):
            self.k = k

    @autooperator('x->y')
    def f(x, cfg):
        return dict(y=x * cfg.k)

    def run(cfg):
        with Builder() as m:
            x = m.input('x')
            m.output(y=f(x, cfg))
        return m.compute('y', init=dict(x=3.0))

    cfg = Cfg(2)
    assert run(cfg) == 6.0
    cfg.k = 5                       # in-place mutation of the same object
    assert run(cfg) == 15.0
    n0 = len(f._model_cache)
    assert run(cfg) == 15.0         # ... while an unchanged one is served from the memo
    assert len(f._model_cache) == n0

    # a class used as a namespace (the way cosmologies are passed, fastpm.py:504)
    class Cosmo(object):
        Om0 = 0.3

    @autooperator('x->y')
    def g(x, cosmology):
        return dict(y=x * cosmology.Om0)

    def run_g():
        with Builder() as m:
            x = m.input('x')
            m.output(y=g(x, Cosmo))
        return m.compute('y', init=dict(x=1.0))

    assert run_g() == 0.3
    Cosmo.Om0 = 0.25
    assert run_g() == 0.25

    # large writeable arrays cannot be fingerprinted: no memo, always rebuilt
    @autooperator('x->y')
    def h(x, table):
        return dict(y=x * float(table[0]))

    table = numpy.ones(10000)

    def run_h():
        with Builder() as m:
            x = m.input('x')
            m.output(y=h(x, table))
        return m.compute('y', init=dict(x=2.0))

    assert run_h() == 2.0
    table[0] = 4.0
    assert run_h() == 8.0
    assert len(h._model_cache) == 0

    # closures are keyed by the values they captured
    def make(scale_by):
        def tf(v):
            return v * scale_by
        return tf

    @autooperator('x->y')
    def k(x, fn):
        return dict(y=x * fn(1.0))

    def run_k(fn):
        with Builder() as m:
            x = m.input('x')
            m.output(y=k(x, fn))
        return m.compute('y', init=dict(x=1.0))

    assert run_k(make(2.0)) == 2.0
    assert run_k(make(3.0)) == 3.0
