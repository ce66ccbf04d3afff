"""Tape -> vjp model / jvp model (the job of vmad/core/autodiff.py:157-236).

Both walk the tape (backwards for vjp, forwards for jvp) and emit one node of the recorded
operator's vjp / jvp primitive per record, fed with the recorded arguments as literals and the
cotangent / tangent symbols.  Partial gradients of a value used more than once are summed with
the stdlib ``add`` in the order the uses are met walking backwards; absent gradients are the
``ZeroLiteral`` (int 0), and gradients with respect to literal inputs are never produced.

The helpers are module-level functions (no self-referential closures): a derived model is
built, run once and dropped on every backward pass, and it must be freed by reference counting
alone because its literals pin device memory.
"""
from .graph import List, Literal, Symbol, ZeroLiteral
from .ops import EmptyPrimitive

_ZERO = ZeroLiteral()


def _std():
    from . import std
    return std


def _lookup(var, table, literal_is_zero):
    """symbol (or nested list of symbols) -> its cotangent / tangent symbol, or ZERO"""
    if isinstance(var, List):
        return [_lookup(v, table, literal_is_zero) for v in var]
    if literal_is_zero and isinstance(var, Literal):
        return _ZERO
    return table.get(id(var), _ZERO)


def _new_partial(var, model, base, sink):
    """fresh symbol(s) receiving the partial gradient of one use of ``var``"""
    if isinstance(var, List):
        return List([_new_partial(v, model, base, sink) for v in var])
    if isinstance(var, Literal):
        return Symbol(model.unique_name(base + '-ignored'), model=model)
    s = Symbol(model.unique_name('_' + var._name + '#'), model=model)
    sink.append((var, s))
    return s


def _new_tangent(var, model, table):
    if isinstance(var, List):
        return List([_new_tangent(v, model, table) for v in var])
    s = Symbol(model.unique_name(var._name + '_'), model=model)
    table[id(var)] = s
    return s


def vjpmodel(tape):
    model = type(tape.model)()
    grads = {}  # id(forward symbol) -> symbol holding its (so far accumulated) cotangent
    for var in tape.model._vout:
        grads[id(var)] = model.input(var._vjp_name)

    add = _std().add.apl
    for record in reversed(tape):
        node = record.node
        prim = node.operator.vjp
        if prim is EmptyPrimitive:
            continue
        kwargs = dict(record.impl_kwargs)
        for argname, var in node.varout.items():
            kwargs['_' + argname] = _lookup(var, grads, False)
        kwout = {}
        partials = []
        for argname, var in node.varin.items():
            if '_' + argname in prim.outnames:
                kwout['_' + argname] = _new_partial(var, model, node.name, partials)
        prim.create_node(kwargs, kwout, model=model, where=node._where)
        for var, part in partials:
            have = grads.get(id(var))
            if have is None:
                grads[id(var)] = part
            else:
                total = Symbol(model.unique_name('_' + var._name + '+'), model=model)
                add.create_node(dict(x1=have, x2=part), dict(y=total), model=model)
                grads[id(var)] = total

    for var in tape.model._vin:
        model.output(**{var._vjp_name: grads.get(id(var), _ZERO)})
    return model


def jvpmodel(tape):
    model = type(tape.model)()
    tangents = {}
    for var in tape.model._vin:
        tangents[id(var)] = model.input(var._jvp_name)

    for record in tape:
        node = record.node
        prim = node.operator.jvp
        kwargs = dict(record.impl_kwargs)
        for argname, var in node.varin.items():
            kwargs[argname + '_'] = _lookup(var, tangents, True)
        kwout = {}
        for argname, var in node.varout.items():
            if argname + '_' in prim.outnames:
                kwout[argname + '_'] = _new_tangent(var, model, tangents)
        prim.create_node(kwargs, kwout, model=model, where=node._where)

    for var in tape.model._vout:
        model.output(**{var._jvp_name: tangents.get(id(var), _ZERO)})
    return model
