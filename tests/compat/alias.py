"""Expose vmad_b200's tape VM under the reference's module names.

Loaded (as a pytest plugin: ``-p alias``) by tests/test_reference_compat.py in a scratch
directory, so that the REFERENCE's own test files and operator libraries
(/root/reference/vmad/core/tests, /root/reference/vmad/lib) import ``vmad.*`` and get the
re-implemented VM.  ``vmad.lib`` resolves to the reference's own lib directory: its linalg /
unary / binary / mpi / fastpm modules execute unmodified on top of vmad_b200.vm, and pmesh /
fastpm.background / mpi4py / nbodykit resolve to the oracle and its shims.
"""
import importlib.util
import os
import sys
import types

ROOT = os.environ['VMAD_B200_ROOT']
REF = os.environ.get('VMAD_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
oracle.activate()

from vmad_b200.vm import diff, errors, graph, ops, std  # noqa: E402


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pub(module):
    return {k: v for k, v in module.__dict__.items() if not k.startswith('__')}


vmad = _mod('vmad', operator=ops.operator, autooperator=ops.autooperator, Builder=graph.Builder,
            ListPlaceholder=graph.ListPlaceholder)
vmad.__path__ = []
core = _mod('vmad.core')
core.__path__ = []
vmad.core = core
for name, src in [('model', graph), ('symbol', graph), ('context', graph), ('tape', graph),
                  ('node', graph), ('operator', ops), ('primitive', ops), ('autooperator', ops),
                  ('error', errors), ('autodiff', diff), ('stdlib', std)]:
    setattr(core, name, _mod('vmad.core.' + name, **_pub(src)))
core.stdlib.__path__ = []
lib = _mod('vmad.lib')
lib.__path__ = [os.path.join(REF, 'vmad', 'lib')]
vmad.lib = lib

spec = importlib.util.spec_from_file_location('vmad.testing', os.path.join(REF, 'vmad', 'testing.py'))
testing = importlib.util.module_from_spec(spec)
sys.modules['vmad.testing'] = testing
spec.loader.exec_module(testing)
vmad.testing = testing
# pytest >= 8 no longer calls nose-style setup()
for _k in ('BaseScalarTest', 'BaseVectorTest'):
    getattr(testing, _k).setup_method = lambda self, method=None: self.setup()
