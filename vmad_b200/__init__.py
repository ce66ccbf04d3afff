"""vmad_b200: B200-native particle-mesh backend behind vmad's operator contract.

``from vmad_b200 import operator, autooperator, Builder`` mirrors ``from vmad import ...``
(vmad/__init__.py:1-4); ``vmad_b200.lib.fastpm`` mirrors ``vmad.lib.fastpm``.
"""
from .vm import Builder, ListPlaceholder, autooperator, operator  # noqa: F401
