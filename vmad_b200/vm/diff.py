"""Tape -> vjp model / jvp model (the job of vmad/core/autodiff.py:157-236).

Both walk the tape (backwards for vjp, forwards for jvp) and emit one node of the recorded
operator's vjp / jvp primitive per record, fed with the recorded arguments as literals and the
cotangent / tangent symbols.  Partial gradients of a value used more than once are summed with
the stdlib ``add`` in the order the uses are met walking backwards; absent gradients are the
``ZeroLiteral`` (int 0), and gradients with respect to literal inputs are never produced.
"""
from .graph import List, Literal, Symbol, ZeroLiteral
from .ops import EmptyPrimitive


def _std():
    from . import std
    return std


def vjpmodel(tape):
    model = type(tape.model)()
    zero = ZeroLiteral()
    grads = {}  # id(forward symbol) -> symbol holding its (so far accumulated) cotangent

    for var in tape.model._vout:
        grads[id(var)] = model.input(var._vjp_name)

    def cotangent(var):
        if isinstance(var, List):
            return [cotangent(v) for v in var]
        return grads.get(id(var), zero)

    def new_partial(var, base, sink):
        """symbol(s) that receive the partial gradient of one use of ``var``"""
        if isinstance(var, List):
            return List([new_partial(v, base, sink) for v in var])
        if isinstance(var, Literal):
            return Symbol(model.unique_name(base + '-ignored'), model=model)
        s = Symbol(model.unique_name('_' + var._name + '#'), model=model)
        sink.append((var, s))
        return s

    add = _std().add.apl
    for record in reversed(tape):
        node = record.node
        prim = node.operator.vjp
        if prim is EmptyPrimitive:
            continue
        kwargs = dict(record.impl_kwargs)
        for argname, var in node.varout.items():
            kwargs['_' + argname] = cotangent(var)
        kwout = {}
        partials = []
        for argname, var in node.varin.items():
            if '_' + argname in prim.outnames:
                kwout['_' + argname] = new_partial(var, node.name, partials)
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
        model.output(**{var._vjp_name: grads.get(id(var), zero)})
    return model


def jvpmodel(tape):
    model = type(tape.model)()
    zero = ZeroLiteral()
    tangents = {}

    for var in tape.model._vin:
        tangents[id(var)] = model.input(var._jvp_name)

    def tangent(var):
        if isinstance(var, List):
            return [tangent(v) for v in var]
        if isinstance(var, Literal):
            return zero
        return tangents.get(id(var), zero)

    def new_out(var):
        if isinstance(var, List):
            return List([new_out(v) for v in var])
        s = Symbol(model.unique_name(var._name + '_'), model=model)
        tangents[id(var)] = s
        return s

    for record in tape:
        node = record.node
        prim = node.operator.jvp
        kwargs = dict(record.impl_kwargs)
        for argname, var in node.varin.items():
            kwargs[argname + '_'] = tangent(var)
        kwout = {}
        for argname, var in node.varout.items():
            if argname + '_' in prim.outnames:
                kwout[argname + '_'] = new_out(var)
        prim.create_node(kwargs, kwout, model=model, where=node._where)

    for var in tape.model._vout:
        model.output(**{var._jvp_name: tangents.get(id(var), zero)})
    return model
