"""``@operator`` / ``@autooperator``: the plugin contract of vmad, re-implemented.

The contract (vmad/core/operator.py:218-262, vmad/core/primitive.py:144-178):

* an operator class declares ``ain`` / ``aout`` and functions ``apl(node, ...)``,
  ``vjp(node, ...)``, ``jvp(node, ...)`` and optionally ``rcd(node, ...)``;
* **the python signature is the interface**: argument names are read from the function's code
  object; positional call arguments map to ``apl``'s names in order, defaults are filled in;
* vjp inputs are ``'_' + aout`` names that appear in the vjp signature, outputs ``'_' + ain``;
  jvp inputs ``ain + '_'``, outputs ``aout + '_'``;
* ``apl`` may return extra keys; the default record keeps every apl input/result whose name
  appears in the vjp or jvp signature (this is how ``pm`` reaches ``r2c.vjp``,
  vmad/lib/fastpm.py:50-51);
* a vjp / jvp whose vector inputs are all zero is bypassed and yields ``ZeroGradient``.

``@autooperator`` turns a model-building function into an operator whose vjp / jvp replay the
inner tape (vmad/core/autooperator.py).  Unlike the reference, the inner model is memoised per
set of hyper-arguments instead of being rebuilt on every call (autooperator.py:133-156), since
on a GPU backend model construction is pure host overhead.  The memo key is BY VALUE
(``_hyper_key``): mutating a hyper-argument between calls yields the model of the new value, as
the reference's rebuild would; hyper-arguments that cannot be fingerprinted disable the memo
for that call.  ``AutoOperator.invalidate()`` drops the memo.
"""
import inspect
from collections import OrderedDict

import numpy

from .errors import BadArgument, MissingArgument, ModelError, OverwritePrecaution
from .graph import (Builder, List, ListPlaceholder, Literal, Model, Node, Symbol, assymbol,
                    where_called)


class ZeroGradientType(int):
    """The universal zero of vjp / jvp propagation (an int equal to 0)."""

    def __new__(kls):
        return int.__new__(kls, 0)

    def __iter__(self):
        while True:
            yield ZeroGradient

    def __array__(self, dtype=None, copy=None):
        return numpy.array(0, dtype=dtype)

    def __repr__(self):
        return "ZeroGradient"

    @staticmethod
    def is_zero(obj):
        if isinstance(obj, ZeroGradientType):
            return True
        t = type(obj)
        if t is int or t is float:
            return obj == 0
        if t is list or t is tuple:
            return all(ZeroGradientType.is_zero(o) for o in obj)
        return False


ZeroGradient = ZeroGradientType()


def _to_ordereddict(spec):
    """``['a', ('b', '*')]``, ``'a'``, ``('a', 'b')`` or a dict -> OrderedDict name -> pattern"""
    if isinstance(spec, (list, tuple)):
        spec = [a if isinstance(a, tuple) else (a, '*') for a in spec]
    elif not isinstance(spec, dict):
        spec = [(spec, '*')]
    return OrderedDict(spec)


def _default_args(func):
    try:
        sig = inspect.signature(func)
    except (TypeError, ValueError):
        return {}
    return {k: p.default for k, p in sig.parameters.items()
            if p.default is not inspect.Parameter.empty}


def _argnames_of(impl):
    code = impl.__code__
    return code.co_varnames[1:code.co_argcount]


class EmptyPrimitive(object):
    argnames = []
    ain = OrderedDict()
    aout = OrderedDict()


class Primitive(object):
    """apl, vjp or jvp of one operator; calling ``create_node`` appends a node to a model."""

    def __init__(self, func, opr, ain, aout, argnames, outnames, impl, record_impl):
        self.name = opr.prototype.__name__ + '-' + func
        self.func = func
        self.operator = opr
        self.ain = ain
        self.aout = aout
        self.argnames = tuple(argnames)
        self.outnames = list(outnames)
        self.impl = impl
        self._defaults = _default_args(getattr(impl, 'original', impl))
        self.record_impl = record_impl

    def __eq__(self, other):
        return (isinstance(other, Primitive) and self.operator is other.operator
                and self.func == other.func)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((id(self.operator), self.func))

    # -- node creation ------------------------------------------------------------------------
    def create_node(self, kwargs, kwout=None, stacklevel=None, model=None, where=None):
        if kwout is None:
            kwout = {}
        node = Node(self, where if where is not None else ("Unknown", "Unknown"))
        varin = OrderedDict()
        for argname in self.argnames:
            if argname not in kwargs:
                if argname in self.ain:
                    raise MissingArgument("Argument %s is missing." % argname)
                continue  # a default of the implementation function applies
            varin[argname] = assymbol(kwargs[argname])
        if model is None:
            models = []
            for v in itertools_chain(varin.values(), kwout.values()):
                _collect_models(v, models)
            model = Model.consolidate(models)
        base = model.unique_name(self.name)
        for argname, var in varin.items():
            node._varin[argname] = var._add_reference(node)
        for argname in self.outnames:
            if argname not in kwout:
                var = Symbol(base + '-' + argname, model=model)
            elif isinstance(kwout[argname], ListPlaceholder):
                var = kwout[argname].as_list(base + '-' + argname, model=model)
            else:
                var = assymbol(kwout[argname])
                if var._has_reference():
                    raise OverwritePrecaution("Overwritting used symbols is not supported. "
                                              "Because it breaks vjp.")
                _anchor(var, model)
            node._varout[argname] = var
        model.append(node)
        return OrderedDict((argname, node._varout[argname]) for argname in self.outnames)

    def _parse_args(self, args, kwargs):
        """positional -> named by signature order; split off output placeholders"""
        if len(args) > len(self.argnames):
            raise BadArgument("Argument list longer than total number of args")
        kwargs = dict(kwargs)
        for argname, arg in zip(self.argnames, args):
            if argname in kwargs:
                raise BadArgument("argument %s already provided as keyword" % argname)
            kwargs[argname] = arg
        for argname, value in self._defaults.items():
            if argname not in kwargs and argname in self.argnames:
                kwargs[argname] = value
        kwout = {}
        for argname in list(kwargs):
            if argname in self.outnames and argname not in self.argnames:
                kwout[argname] = kwargs.pop(argname)
            elif argname not in self.argnames:
                raise BadArgument("Argument %s is passed in , but not declared as an argument"
                                  % argname)
        for argname in self.argnames:
            if argname not in kwargs:
                raise MissingArgument("Argument %s is missing." % argname)
        return kwargs, kwout


def itertools_chain(*its):
    for it in its:
        for x in it:
            yield x


def _collect_models(var, acc):
    if isinstance(var, Symbol):
        m = var._model
        if m is not None and not any(m is a for a in acc):
            acc.append(m)
    elif isinstance(var, (List, list, tuple)):
        for v in var:
            _collect_models(v, acc)
    elif isinstance(var, Literal) and getattr(var, '_parent', None) is not None:
        _collect_models(var._parent, acc)


def _anchor(var, model):
    if isinstance(var, List):
        for v in var:
            _anchor(v, model)
    elif isinstance(var, Symbol) and var._model is not model:
        model.anchor(var)


# ---------------------------------------------------------------------------------------------
# recording and zero bypass
# ---------------------------------------------------------------------------------------------
def record_copy_all(node, **kwargs):
    return kwargs


def record_copy_autodiff(node, **kwargs):
    """keep what the vjp / jvp signatures name (vmad/core/operator.py:165-178)"""
    opr = node.operator
    names = opr._autodiff_argnames
    return {k: v for k, v in kwargs.items() if k in names}


def zerobypass(impl):
    def zerobypassimpl(__node__, **kwargs):
        prim = __node__.primitive
        if all(ZeroGradientType.is_zero(kwargs[a]) for a in prim.ain):
            return {a: ZeroGradient for a in prim.aout}
        return impl(__node__, **kwargs)
    zerobypassimpl.original = impl
    return zerobypassimpl


def default_jvp(node, **kwargs):
    """no jvp given: apply ``apl`` to the tangents (right for linear operators)"""
    apl = node.operator.apl
    d = {}
    for argname, value in kwargs.items():
        d[argname[:-1] if argname.endswith('_') else argname] = value
    r = apl.impl(node, **d)
    if not isinstance(r, dict):
        return r
    return {(k + '_' if k in apl.aout else k): v for k, v in r.items()}


def _make_primitive(opr, func, impl, argnames=None, outnames=None, record_impl=record_copy_all):
    assert func in ('apl', 'vjp', 'jvp')
    if argnames is None:
        argnames = _argnames_of(impl)
    ain, aout = OrderedDict(), OrderedDict()
    if func == 'apl':
        ain, aout = opr.ain, opr.aout
    elif func == 'vjp':
        for arg in opr.ain:
            aout['_' + arg] = opr.ain[arg]
        for arg in opr.aout:
            if '_' + arg in argnames:
                ain['_' + arg] = opr.aout[arg]
        impl = zerobypass(impl)
    else:
        for arg in opr.ain:
            if arg + '_' in argnames:
                ain[arg + '_'] = opr.ain[arg]
        for arg in opr.aout:
            aout[arg + '_'] = opr.aout[arg]
        impl = zerobypass(impl)
    if outnames is None:
        outnames = list(aout.keys())
    return Primitive(func, opr, ain, aout, argnames, outnames, impl, record_impl)


# ---------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------
class BaseOperator(object):
    def __init__(self, prototype):
        self.prototype = prototype
        self.ain = _to_ordereddict(prototype.ain)
        self.aout = _to_ordereddict(prototype.aout)
        for name, value in prototype.__dict__.items():
            if isinstance(value, staticmethod):
                setattr(self, name, getattr(prototype, name))

    def __call__(self, *args, **kwargs):
        kwargs.pop('__stacklevel__', None)
        apl = self.apl
        kw, kwout = apl._parse_args(args, kwargs)
        vout = apl.create_node(kw, kwout, where=where_called(2))
        if len(self.aout) == 1:
            return next(iter(vout.values()))
        return list(vout.values())

    def __get__(self, instance, owner):
        if instance is None:
            return self
        return InstanceOperator(self, instance)


class Operator(BaseOperator):
    def __init__(self, prototype):
        BaseOperator.__init__(self, prototype)
        p = prototype
        record_impl = getattr(p, 'rcd', None) or record_copy_autodiff
        self._apl = _make_primitive(self, 'apl', p.apl, record_impl=record_impl)
        if hasattr(p, 'vjp'):
            self._vjp = _make_primitive(self, 'vjp', p.vjp)
        else:
            self._vjp = EmptyPrimitive
        if hasattr(p, 'jvp'):
            self._jvp = _make_primitive(self, 'jvp', p.jvp)
        else:
            names = [a + '_' if a in self.ain else a for a in self._apl.argnames]
            self._jvp = _make_primitive(self, 'jvp', default_jvp, argnames=names)
        self._autodiff_argnames = frozenset(self._vjp.argnames) | frozenset(self._jvp.argnames)

    @property
    def apl(self):
        return self._apl

    @property
    def vjp(self):
        return self._vjp

    @property
    def jvp(self):
        return self._jvp


class InstanceOperator(object):
    """an operator accessed through an instance: the instance becomes the first argument"""

    def __init__(self, base, instance):
        self.base = base
        self.instance = instance

    def __call__(self, *args, **kwargs):
        return self.base(self.instance, *args, **kwargs)

    def build(self, **kwargs):
        d = dict(kwargs)
        d[self.base.argnames[0]] = self.instance
        return self.base.build(**d)

    def __getattr__(self, attrname):
        return getattr(self.base, attrname)


def operator(kls):
    """class decorator: turn an operator prototype into an ``Operator``"""
    return Operator(kls)


# ---------------------------------------------------------------------------------------------
# auto operators
# ---------------------------------------------------------------------------------------------
TAPENAME = '$$tape$$'


class _Uncacheable(Exception):
    """a hyper-argument whose state cannot be fingerprinted: the model is rebuilt on every call,
    which is what the reference always does (vmad/core/autooperator.py:133-156)"""


def _hyper_key(value, depth=0):
    """Memo key of one hyper-argument -- BY VALUE, so that mutating a hyper-argument between two
    calls selects (or builds) a different model instead of replaying a stale one:

    * python / numpy scalars, strings, small arrays, lists and tuples: their contents;
    * objects that declare ``__vmad_hyper_key__()`` (ParticleMesh, frozen device arrays,
      MatterDominated, FastPMSimulation ...): what they return -- they vouch that everything a
      model build reads from them is covered by that key;
    * plain functions: identity + the keys of their closure cells and defaults;
    * read-only numpy arrays: identity (they cannot be mutated in place);
    * any other object or class: the keys of its public data attributes (a config object whose
      ``k`` changes from 2 to 5 changes key).

    Whatever does not fit (a writeable large array, an object holding one, recursion deeper
    than 3) raises ``_Uncacheable`` and the caller rebuilds, like the reference.
    """
    if value is None or isinstance(value, (bool, int, float, complex, str, bytes)):
        return ('v', type(value).__name__, value)
    if isinstance(value, numpy.generic):
        return ('v', 'np', value.item())
    if isinstance(value, numpy.ndarray):
        if value.size <= 4096 and value.dtype.kind in 'fiubc':
            return ('a', value.shape, value.dtype.str, value.tobytes())
        if not value.flags.writeable and (value.base is None or not getattr(value.base, 'flags', value.flags).writeable):
            return ('ro', id(value))
        raise _Uncacheable("large writeable array")
    if depth > 3:
        raise _Uncacheable("nesting too deep")
    if isinstance(value, (list, tuple)):
        if len(value) > 4096:
            raise _Uncacheable("long sequence")
        return ('t', type(value).__name__, tuple(_hyper_key(v, depth + 1) for v in value))
    if isinstance(value, dict):
        if len(value) > 256:
            raise _Uncacheable("large dict")
        return ('d', tuple(sorted((repr(k), _hyper_key(v, depth + 1)) for k, v in value.items())))
    hook = getattr(value, '__vmad_hyper_key__', None)
    if hook is not None and not inspect.isclass(value):
        try:
            return ('k', type(value).__name__, hook())
        except TypeError as e:          # the object declines (e.g. a mutable device array)
            raise _Uncacheable(str(e))
    if hasattr(value, '__array__') or hasattr(value, '__cuda_array_interface__') or \
            type(value).__module__.split('.')[0] in ('torch', 'scipy', 'pandas'):
        raise _Uncacheable("array-like or foreign object of type %s" % type(value).__name__)
    if inspect.isfunction(value) or inspect.ismethod(value):
        fn = value.__func__ if inspect.ismethod(value) else value
        cells = tuple(_hyper_key(c.cell_contents, depth + 1) for c in (fn.__closure__ or ()))
        defaults = tuple(_hyper_key(d, depth + 1) for d in (fn.__defaults__ or ()))
        owner = (_hyper_key(value.__self__, depth + 1),) if inspect.ismethod(value) else ()
        return ('f', id(fn), cells, defaults, owner)
    if inspect.isbuiltin(value) or isinstance(value, numpy.ufunc):
        return ('b', id(value))
    try:
        attrs = vars(value)
    except TypeError:
        raise _Uncacheable("opaque object of type %s" % type(value).__name__)
    items = []
    for name in sorted(attrs):
        if name.startswith('__') and name.endswith('__'):
            continue
        a = attrs[name]
        if inspect.isroutine(a) or isinstance(a, (staticmethod, classmethod, property)):
            continue
        items.append((name, _hyper_key(a, depth + 1)))
    return ('o', id(value) if inspect.isclass(value) else type(value).__qualname__, tuple(items))


class AutoOperator(BaseOperator):
    """An operator defined by a model-building function ``main(model, **args)``."""

    def __init__(self, prototype, argnames):
        BaseOperator.__init__(self, prototype)
        impl = prototype.main
        if argnames is None:
            argnames = _argnames_of(impl)
        self.argnames = tuple(argnames)
        self.hyperargs = {}
        self.__bound_tape__ = None
        self.__bound_model__ = None
        self.__tapename__ = TAPENAME
        self._prims = {}
        self._model_cache = OrderedDict()

    @property
    def forgetful(self):
        """same operator, but the inner tape is not kept; vjp / jvp recompute the forward"""
        obj = AutoOperator(self.prototype, self.argnames)
        obj.__bound_tape__ = self.__bound_tape__
        obj.__bound_model__ = self.__bound_model__
        obj.__tapename__ = None
        return obj

    # -- primitives -------------------------------------------------------------------------------
    @property
    def apl(self):
        return self.get_apl(self.__tapename__)

    @property
    def vjp(self):
        return self.get_vjp(self.__tapename__)

    @property
    def jvp(self):
        return self.get_jvp(self.__tapename__)

    def get_apl(self, tapename):
        key = ('apl', tapename)
        if key not in self._prims:
            def apl(node, **kwargs):
                return node.operator._compute(kwargs, tapename=tapename)

            def rcd(node, **kwargs):
                return kwargs
            self._prims[key] = _make_primitive(self, 'apl', apl, argnames=self.argnames,
                                               outnames=list(self.aout.keys()), record_impl=rcd)
        return self._prims[key]

    def get_vjp(self, tapename):
        key = ('vjp', tapename)
        if key not in self._prims:
            argnames = list(self.argnames)
            if tapename is not None:
                argnames.append(tapename)
            argnames.extend('_' + a for a in self.aout)

            def vjp(node, **kwargs):
                if tapename is not None:
                    tape = kwargs.pop(tapename)
                else:
                    tape = node.operator._compute(kwargs, TAPENAME)[TAPENAME]
                v = [(a, kwargs[a]) for a in node.primitive.ain if a.startswith('_')]
                vjpvout = tape.get_vjp_vout()
                vjpout = tape.get_vjp().compute(vjpvout, init=v)
                return dict(zip(vjpvout, vjpout))
            self._prims[key] = _make_primitive(self, 'vjp', vjp, argnames=argnames)
        return self._prims[key]

    def get_jvp(self, tapename):
        key = ('jvp', tapename)
        if key not in self._prims:
            argnames = list(self.argnames)
            if tapename is not None:
                argnames.append(tapename)
            argnames.extend(a + '_' for a in self.ain)

            def jvp(node, **kwargs):
                if tapename is not None:
                    tape = kwargs[tapename]
                else:
                    tape = node.operator._compute(kwargs, TAPENAME)[TAPENAME]
                v = [(a, kwargs[a]) for a in node.primitive.ain if a.endswith('_')]
                jvpvout = tape.get_jvp_vout()
                jvpout = tape.get_jvp().compute(jvpvout, init=v)
                return dict(zip(jvpvout, jvpout))
            self._prims[key] = _make_primitive(self, 'jvp', jvp, argnames=argnames)
        return self._prims[key]

    # -- evaluation -------------------------------------------------------------------------------
    def _compute(self, kwargs, tapename):
        if self.__bound_tape__ is not None:
            tape = self.__bound_tape__
            y = dict(tape.out)
        else:
            m = self._build(kwargs)
            init = [(a, kwargs[a]) for a in self.ain]
            vout = list(self.aout.keys())
            if tapename is not None:
                y, tape = m.compute(vout, init=init, return_dict=True, return_tape=True)
            else:
                y = m.compute(vout, init=init, return_dict=True, return_tape=False)
            y = dict(y)
        if tapename is not None:
            y[tapename] = tape
        return y

    def _build(self, kwargs):
        if self.__bound_model__ is not None:
            return self.__bound_model__
        hyper = [(a, kwargs[a]) for a in self.argnames if a not in self.ain and a in kwargs]
        try:
            key = tuple((a, _hyper_key(v)) for a, v in hyper)
        except _Uncacheable:
            return self._build_model(hyper)
        # an instance-bound autooperator (FastPMSimulation.run ...) memoises on the instance, so
        # that dropping the simulation drops its models (and the device arrays they hold)
        cache = self._model_cache
        if hyper:
            own = getattr(hyper[0][1], '_vmad_model_cache', None)
            if isinstance(own, dict):
                cache = own.setdefault(id(self), OrderedDict())
        hit = cache.get(key)
        if hit is not None:
            cache.move_to_end(key)
            return hit[0]
        m = self._build_model(hyper)
        # keep the hyper-arguments alive next to the model so identity-based key parts stay unique
        cache[key] = (m, [v for _, v in hyper] if cache is self._model_cache else None)
        while len(cache) > 64:
            cache.popitem(last=False)
        return m

    def _build_model(self, hyper):
        impl = self.prototype.main
        model_args = dict(hyper)
        with Builder() as m:
            for argname in self.ain:
                model_args[argname] = m.input(argname)
            r = impl(m, **model_args)
            for argname in self.aout:
                if argname not in r:
                    raise ModelError("output arg '%s' is not produced by the model" % argname)
            m.output(**r)
        return m

    def invalidate(self):
        """drop every memoised model of this operator (after changing global state that a model
        build reads and that no hyper-argument carries)"""
        self._model_cache.clear()

    def build(__self__, **kwargs):
        """the model of this operator for the given hyper-parameters"""
        self = __self__
        for argname in kwargs:
            if argname in self.ain:
                raise ModelError("argname %s is an input, shall not be used to produce a model"
                                 % argname)
        return self._build(kwargs)

    def precompute(self, **kwargs):
        """run once; the returned operator replays that tape whatever its inputs
        (vmad/core/autooperator.py:170-185)"""
        y = self._compute(kwargs, TAPENAME)
        m = self._build(kwargs)
        obj = AutoOperator(self.prototype, argnames=self.argnames)
        obj.__bound_tape__ = y[TAPENAME]
        obj.__bound_model__ = m
        return obj

    def bind(self, **hyperargs):
        """fix the hyper-parameters; the returned operator takes only the remaining arguments"""
        for argname in hyperargs:
            if argname in self.ain:
                raise ModelError("argname %s is an input, shall not be used to produce a model"
                                 % argname)
        m = self._build(hyperargs)
        argnames = tuple(a for a in self.argnames if a not in hyperargs)
        obj = AutoOperator(self.prototype, argnames=argnames)
        obj.__bound_model__ = m
        obj.hyperargs = hyperargs
        return obj


def _autograd(func, ain, aout):
    in_names = [a[0] if isinstance(a, tuple) else a for a in ain]

    def main(__model__, *args, **kwargs):
        kwargs = dict(kwargs)
        kwargs.update(zip(in_names, args))
        r = func(**kwargs)
        if isinstance(r, tuple):
            r = dict(zip(aout, r))
        elif not isinstance(r, dict) and len(aout) == 1:
            r = {aout[0]: r}
        return r

    code = func.__code__
    argnames = code.co_varnames[:code.co_argcount]
    prototype = type(func.__name__, (), dict(ain=ain, aout=aout, main=main))
    return AutoOperator(prototype, argnames=argnames)


def _parse_spec(args):
    if len(args) == 1:
        if callable(args[0]):
            func = args[0]
            ann = getattr(func, '__annotations__', {})
            if not ann:
                raise ValueError("Must provide ain, aout or function annotations.")
            ain = [(k, v) for k, v in ann.items() if k != 'return']
            return ain, ann['return'], func
        spec = args[0]
        left, right = spec.split('->')
        return [a.strip() for a in left.split(',')], [a.strip() for a in right.split(',')], None
    split = args.index('->')
    return list(args[:split]), list(args[split + 1:]), None


def autooperator(*args):
    """``@autooperator('x, y->z')`` on a function, or ``@autooperator`` on a class with
    ``ain`` / ``aout`` / ``main`` (vmad/core/autooperator.py:261-345)."""
    if inspect.isclass(args[0]):
        return AutoOperator(args[0], argnames=None)
    ain, aout, func = _parse_spec(args)
    if isinstance(aout, str):
        aout = [aout]

    def wrapped(f):
        return _autograd(f, ain, aout)

    return wrapped if func is None else wrapped(func)
