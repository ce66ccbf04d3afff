"""Symbols, nodes, models, execution and tapes of the vmad-compatible tape VM.

vmad (the reference) is a tape-based autodiff "virtual machine": a model is a linear list of
nodes executed in build order (vmad/core/model.py:11, vmad/core/context.py:61-88); running it
with ``return_tape=True`` records, per executed node, the arguments its vjp/jvp need
(vmad/core/tape.py:16-58).  This module provides the same public behaviour --
``Builder``/``Model.input``/``output``/``compute``/``compute_with_vjp``/``compute_with_jvp``/
``compute_with_gnDp`` and ``Tape.get_vjp``/``get_jvp`` -- with one structural difference that
matters for a GPU backend: execution goes through a cached *plan* (dead-node elimination from
the requested outputs plus a precomputed free list per step) instead of re-deriving liveness
with an O(nodes^2) scan on every run (vmad/core/context.py:31-44), because here the host loop,
not the numerics, is the bottleneck.
"""
import itertools
import sys
import weakref
from collections import OrderedDict

from .errors import (DuplicatedOutput, InferError, ModelError, OverwritePrecaution,
                     ResolveError, UnexpectedOutput, makeExecutionError)

_counter = itertools.count()

_raise_internal_errors = True


def set_raise_internal_errors(flag):
    """True: node errors propagate unchanged; False: they are wrapped in an ExecutionError that
    names the line where the node was declared (vmad/core/context.py:8-18)."""
    global _raise_internal_errors
    _raise_internal_errors = flag


def _std():
    from . import std
    return std


# ---------------------------------------------------------------------------------------------
# symbols
# ---------------------------------------------------------------------------------------------
class BaseSymbol(object):
    """Something that resolves to a python value inside an execution context."""
    __array_priority__ = 15.0  # numpy scalars defer to our reflected operators

    def __init__(self):
        self._nrefs = 0

    # a symbol doubles as its own reference record
    @property
    def symbol(self):
        return self

    def _add_reference(self, node):
        self._nrefs += 1
        return self

    def _has_reference(self):
        return self._nrefs > 0

    def _resolve(self, ctx):
        raise NotImplementedError

    def _store(self, ctx, value):
        pass

    def _names(self, acc):
        """collect the context names this symbol reads"""
        return acc

    def eval(self, function):
        return _std().eval(self, function)

    def __add__(self, o): return _std().add(self, o)
    def __radd__(self, o): return _std().add(o, self)
    def __sub__(self, o): return _std().sub(self, o)
    def __rsub__(self, o): return _std().sub(o, self)
    def __mul__(self, o): return _std().mul(self, o)
    def __rmul__(self, o): return _std().mul(o, self)
    def __truediv__(self, o): return _std().div(self, o)
    def __rtruediv__(self, o): return _std().div(o, self)
    def __mod__(self, o): return _std().mod(self, o)
    def __rmod__(self, o): return _std().mod(o, self)
    def __neg__(self): return _std().neg(self)
    def __pos__(self): return _std().pos(self)
    def __abs__(self): return _std().abs(self)

    def __pow__(self, o, modulo=None):
        if modulo is not None:
            raise ValueError("pow with modulo is not supported")
        return _std().pow(self, o)

    def __floordiv__(self, o):
        raise TypeError("floor div is not supported in autodiff")

    __rfloordiv__ = __floordiv__


class Symbol(BaseSymbol):
    """A named value; bound to the model it was created for."""

    def __init__(self, name, model=None):
        assert isinstance(name, str)
        BaseSymbol.__init__(self)
        self._name = name
        # weak reference to a real model (so that dropping a model frees its nodes, symbols and
        # the tensors its literals pin by reference counting alone -- a vjp model is built, run
        # once and dropped on every backward pass), strong reference to a transient one
        self._model_ref = None
        if model is not None:
            model.anchor(self)

    @property
    def _model(self):
        r = self._model_ref
        if r is None or isinstance(r, Model):
            return r
        return r()

    def __repr__(self):
        return self._name

    @property
    def _vjp_name(self):
        return '_' + self._name

    @property
    def _jvp_name(self):
        return self._name + '_'

    def _resolve(self, ctx):
        try:
            return ctx[self._name]
        except KeyError:
            raise ResolveError("Symbol %s does not exist in the context" % self._name)

    def _store(self, ctx, value):
        ctx[self._name] = value

    def _names(self, acc):
        acc.add(self._name)
        return acc


class Literal(BaseSymbol):
    """A constant; takes no part in gradient propagation (vmad/core/autodiff.py:93-94)."""

    def __init__(self, value):
        BaseSymbol.__init__(self)
        self._value = value

    def __repr__(self):
        return "L(%s)" % type(self._value).__name__

    def _resolve(self, ctx):
        return self._value

    def __getattr__(self, attrname):
        if attrname.startswith('_'):
            raise AttributeError(attrname)
        return AttrSymbol(self, attrname)


class AttrSymbol(Literal):
    """Attribute of another symbol's value, resolved at run time; always a literal."""

    def __init__(self, parent, attrname):
        Literal.__init__(self, None)
        self._parent = parent
        self._attrname = attrname

    def _resolve(self, ctx):
        return getattr(self._parent._resolve(ctx), self._attrname)

    def _names(self, acc):
        return self._parent._names(acc)

    def __repr__(self):
        return "%r.%s" % (self._parent, self._attrname)


class ZeroLiteral(Literal):
    """Marks an absent gradient; resolves to the int 0."""

    def __init__(self):
        Literal.__init__(self, 0)

    def __repr__(self):
        return "[ZERO]"


class List(BaseSymbol):
    """A python list of symbols used as one argument (e.g. ``linalg.stack([a, b, c])``)."""

    def __init__(self, items):
        BaseSymbol.__init__(self)
        self._items = list(items)

    def __repr__(self):
        return "L%s" % (self._items,)

    def __len__(self):
        return len(self._items)

    def __getitem__(self, i):
        return self._items[i]

    def __iter__(self):
        return iter(self._items)

    def _resolve(self, ctx):
        return [v._resolve(ctx) for v in self._items]

    def _store(self, ctx, value):
        for var, v in zip(self._items, value):
            var._store(ctx, v)

    def _add_reference(self, node):
        self._nrefs += 1
        for v in self._items:
            v._add_reference(node)
        return self

    def _has_reference(self):
        return self._nrefs > 0 or any(v._has_reference() for v in self._items)

    def _names(self, acc):
        for v in self._items:
            v._names(acc)
        return acc


class ListPlaceholder(object):
    """Declares that an operator output is a list of ``size`` symbols."""

    def __init__(self, size, varname=None):
        self.size = size
        self.varname = varname

    def as_list(self, default_varname, model=None):
        name = self.varname if self.varname is not None else default_varname
        return List([Symbol("%s-%d" % (name, i), model=model) for i in range(self.size)])


def assymbol(obj):
    if type(obj) in (list, tuple):
        return List([assymbol(i) for i in obj])
    if isinstance(obj, BaseSymbol):
        return obj
    return Literal(obj)


# ---------------------------------------------------------------------------------------------
# nodes
# ---------------------------------------------------------------------------------------------
class Node(object):
    """One application of a primitive; first argument of every apl / vjp / jvp / rcd function.

    ``node[argname]`` is the input symbol, ``node.is_literal(argname)`` tells whether the
    argument is a constant (vmad/core/node.py:111-112).
    """
    __slots__ = ('primitive', 'operator', 'prototype', 'name', '_where', '_varin', '_varout',
                 '_in_names', '_out_names')

    def __init__(self, primitive, where):
        self.primitive = primitive
        self.operator = primitive.operator
        self.prototype = primitive.operator.prototype
        self.name = primitive.name
        self._where = where
        self._varin = OrderedDict()
        self._varout = OrderedDict()
        self._in_names = None
        self._out_names = None

    def __getitem__(self, key):
        return self._varin[key]

    @property
    def varin(self):
        return self._varin

    @property
    def varout(self):
        return self._varout

    def __repr__(self):
        return "%s @ %s : %s " % (self.name, self._where[0], self._where[1])

    def is_literal(self, argname):
        return isinstance(self._varin[argname], Literal)

    def find_primitive_type(self, func):
        return getattr(self.operator, func)

    def in_names(self):
        if self._in_names is None:
            acc = set()
            for v in self._varin.values():
                v._names(acc)
            self._in_names = frozenset(acc)
        return self._in_names

    def out_names(self):
        if self._out_names is None:
            acc = set()
            for v in self._varout.values():
                v._names(acc)
            self._out_names = frozenset(acc)
        return self._out_names

    def call(self, kwargs):
        r = self.primitive.impl(self, **kwargs)
        if not isinstance(r, dict):
            nout = len(self._varout)
            if nout == 1:
                r = {next(iter(self._varout)): r}
            elif nout == 0:
                if r is not None:
                    raise ValueError("Return value of the primitive is not None, while no "
                                     "output arguments are defined")
                r = {}
            elif isinstance(r, (tuple, list)) and len(r) == nout:
                r = dict(zip(self._varout, r))
        return r

    def record(self, kwargs, r):
        """what goes on the tape for this node: inputs take priority over results
        (vmad/core/node.py:74-97)"""
        d = dict(r)
        d.update(kwargs)
        return self.primitive.record_impl(self, **d)


# ---------------------------------------------------------------------------------------------
# tape
# ---------------------------------------------------------------------------------------------
class Record(object):
    __slots__ = ('node', 'impl_kwargs')

    def __init__(self, node, impl_kwargs):
        self.node = node
        self.impl_kwargs = impl_kwargs

    def __repr__(self):
        return '%s / %s' % (self.node, list(self.impl_kwargs))


class Tape(list):
    def __init__(self, model, init):
        list.__init__(self)
        self.model = model
        self.init = init
        self.out = None
        self._completed = False

    def finalize(self, out):
        assert isinstance(out, dict)
        self.out = out.copy()
        self._completed = True

    def append(self, node, impl_kwargs):
        assert not self._completed
        list.append(self, Record(node, impl_kwargs))

    def get_vjp_vout(self):
        return ['_' + name for name in self.init.keys()]

    def get_jvp_vout(self):
        return [name + '_' for name in self.out.keys()]

    def get_vjp(self):
        assert self._completed
        from .diff import vjpmodel
        return vjpmodel(self)

    def get_jvp(self):
        assert self._completed
        from .diff import jvpmodel
        return jvpmodel(self)

    def compute_jvjp(self, vout, aout, init):
        jvp = self.get_jvp()
        t = jvp.compute([a + '_' for a in aout], init)
        vjp = self.get_vjp()
        return vjp.compute(vout, init=dict(('_' + a, t1) for a, t1 in zip(aout, t)))


# ---------------------------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------------------------
class Model(list):
    """A linear program: nodes in build order, declared inputs and outputs."""

    def __init__(self):
        list.__init__(self)
        self._vin = []
        self._vout = []
        self._uid = next(_counter)
        self._symbols = []
        self._plans = {}
        self._weakself = weakref.ref(self)

    def __hash__(self):
        return self._uid

    def __eq__(self, other):
        return self is other

    def __ne__(self, other):
        return self is not other

    def __repr__(self):
        return "Model: %s => %s" % (self._vin, self._vout)

    # -- construction ------------------------------------------------------------------------
    def anchor(self, symbol):
        old = symbol._model
        if old is not None and old is not self and not isinstance(old, ModelInTransient):
            raise ModelError("cannot change the model after the symbol is no longer in transient.")
        if isinstance(self, ModelInTransient):
            symbol._model_ref = self
        else:
            symbol._model_ref = self._weakself
        self._symbols.append(symbol)

    def append(self, node):
        list.append(self, node)
        self._plans.clear()

    def extend(self, other):
        self._vin.extend(other._vin)
        self._vout.extend(other._vout)
        for s in other._symbols:
            s._model_ref = None
            self.anchor(s)
        list.extend(self, other)
        self._plans.clear()

    def input(self, *args):
        v = []
        for name in args:
            r = name if isinstance(name, Symbol) else Symbol(name)
            self.anchor(r)
            v.append(r)
        self._vin.extend(v)
        return v[0] if len(args) == 1 else v

    def output(self, **kwargs):
        for varname, oldvar in kwargs.items():
            if any(var._name == varname for var in self._vout):
                raise DuplicatedOutput("Variable %s is already marked as an output" % varname)
            var = Symbol(varname, model=self)
            _std().terminal.apl.create_node(dict(x=oldvar), dict(y=var), model=self)
            self._vout.append(var)

    def compile(self):
        pass

    def unique_name(self, base):
        return '%s-%d' % (base, next(_counter))

    @staticmethod
    def get_model(symbol):
        return symbol._model

    @staticmethod
    def consolidate(models):
        """Merge the models touched by one node creation; a real model absorbs transient ones."""
        real = None
        for m in models:
            if m is None or isinstance(m, ModelInTransient):
                continue
            if real is None:
                real = m
            elif real is not m:
                raise InferError("Multiple non transient models are found")
        if real is None:
            real = models[0] if models else ModelInTransient()
        for m in models:
            if m is not real and m is not None:
                real.extend(m)
        return real

    # -- execution ---------------------------------------------------------------------------
    def _plan(self, vout):
        """(steps, drop_first) for the requested outputs.

        steps: list of (node, is_terminal, names_to_free_after).  Only nodes that feed a
        requested output (or have no outputs at all, i.e. exist for their side effect) run.
        """
        key = tuple(vout)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        terminal = _std().terminal.apl
        wanted = set(vout)
        needed = set()
        keep = [False] * len(self)
        for i in range(len(self) - 1, -1, -1):
            node = self[i]
            outs = node.out_names()
            if node.primitive is terminal:
                k = bool(outs & wanted)
            elif not node._varout:
                k = True
            else:
                k = bool(outs & needed)
            if k:
                keep[i] = True
                needed -= outs
                needed |= node.in_names()
        steps = []
        live = set()
        for i in range(len(self) - 1, -1, -1):
            if not keep[i]:
                continue
            node = self[i]
            ins, outs = node.in_names(), node.out_names()
            free = [n for n in (ins | outs) if n not in live]
            steps.append((node, node.primitive is terminal, free))
            live |= ins
        steps.reverse()
        plan = (steps, live)
        self._plans[key] = plan
        return plan

    def compute(self, vout, init, return_tape=False, return_dict=False, monitor=None):
        """Run the model from the initial values ``init`` (vmad/core/model.py:127-163)."""
        single = isinstance(vout, str)
        if single:
            vout = [vout]
        init = OrderedDict(init)
        declared = set(var._name for var in self._vout)
        for name in vout:
            if name not in declared:
                raise UnexpectedOutput("Requested vout %s is not defined by the model as an "
                                       "output; available ones are %s" % (name, declared))
        tape = Tape(self, init) if return_tape else None
        steps, live_at_start = self._plan(vout)
        ctx = dict(init)
        for name in list(ctx):
            if name not in live_at_start:
                del ctx[name]
        results = {}
        for node, is_terminal, free in steps:
            kwargs = {}
            for argname, var in node._varin.items():
                kwargs[argname] = var._resolve(ctx)
            if _raise_internal_errors:
                r = node.call(kwargs)
            else:
                try:
                    r = node.call(kwargs)
                except Exception as e:
                    raise makeExecutionError("Error computing node : %s" % (node,), e)
            if tape is not None:
                tape.append(node, node.record(kwargs, r))
            for argname, var in node._varout.items():
                var._store(ctx, r[argname] if argname in r else 0)
            if is_terminal:
                for var in node._varout.values():
                    results[var._name] = ctx[var._name]
            if monitor is not None:
                monitor(node, node._varin, node._varout, ctx)
            for name in free:
                ctx.pop(name, None)
        out = OrderedDict((name, results[name]) for name in vout)
        if return_tape:
            tape.finalize(out)
        if not return_dict:
            out = list(out.values())
            if single:
                out = out[0]
        if return_tape:
            return out, tape
        return out

    def compute_with_vjp(self, init, v, return_dict=False, monitor=None):
        """Forward, then back-propagate the cotangents ``v`` (keys ``'_' + output name``);
        returns ``(outputs, [gradient per init key])`` (vmad/core/model.py:165-212)."""
        init = OrderedDict(init)
        v = OrderedDict(v)
        assert all(name.startswith('_') for name in v)
        vout = [name[1:] for name in v]
        out, tape = self.compute(vout, init=init, return_tape=True, monitor=monitor)
        vjpvout = tape.get_vjp_vout()
        vjpout = tape.get_vjp().compute(vjpvout, init=v, monitor=monitor)
        if return_dict:
            return OrderedDict(zip(vout, out)), OrderedDict(zip(vjpvout, vjpout))
        return out, vjpout

    def compute_with_jvp(self, vout, init, v, return_dict=False, monitor=None):
        """Forward, then push the tangents ``v`` (keys ``input name + '_'``) forward
        (vmad/core/model.py:214-260)."""
        init = OrderedDict(init)
        v = OrderedDict(v)
        assert all(name.endswith('_') for name in v)
        out, tape = self.compute(vout, init=init, return_tape=True, monitor=monitor)
        jvpvout = tape.get_jvp_vout()
        jvpout = tape.get_jvp().compute(jvpvout, init=v, monitor=monitor)
        if return_dict:
            vl = [vout] if isinstance(vout, str) else vout
            ol = [out] if isinstance(vout, str) else out
            return OrderedDict(zip(vl, ol)), OrderedDict(zip(jvpvout, jvpout))
        if not isinstance(vout, (tuple, list)):
            jvpout = jvpout[0]
        return out, jvpout

    def compute_with_gnDp(self, vout, init, v, return_dict=False, monitor=None):
        """Gauss-Newton Hessian-vector product ``J^T J v`` (vmad/core/model.py:262-311)."""
        init = OrderedDict(init)
        v = OrderedDict(v)
        assert all(name.endswith('_') for name in v)
        out, tape = self.compute(vout, init=init, return_tape=True, monitor=monitor)
        jvpvout = tape.get_jvp_vout()
        jvpout = tape.get_jvp().compute(jvpvout, init=v, monitor=monitor)
        vjpvout = tape.get_vjp_vout()
        vjpvin = ['_' + name[:-1] for name in jvpvout]
        gn = tape.get_vjp().compute(vjpvout, init=OrderedDict(zip(vjpvin, jvpout)), monitor=monitor)
        if return_dict:
            vl = [vout] if isinstance(vout, str) else vout
            ol = [out] if isinstance(vout, str) else out
            return OrderedDict(zip(vl, ol)), OrderedDict(zip(vjpvout, gn))
        return out, gn


class ModelInTransient(Model):
    """Holds nodes whose arguments did not yet touch a real model; merged on first contact."""
    pass


class Builder(Model):
    """``with Builder() as m:`` -- the model under construction."""

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.compile()


def where_called(depth):
    """(filename, lineno) of the user code that created a node, cheaply (no source lookup)"""
    try:
        f = sys._getframe(depth)
        return (f.f_code.co_filename, f.f_lineno)
    except ValueError:
        return ("Unknown", "Unknown")
