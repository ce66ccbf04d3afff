"""vmad-compatible tape VM: ``operator``, ``autooperator``, ``Builder``, ``ListPlaceholder``."""
from .graph import Builder, ListPlaceholder, Model, Symbol, Literal, List, set_raise_internal_errors
from .ops import ZeroGradient, autooperator, operator
from . import std as stdlib
