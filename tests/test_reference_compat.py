"""The reference's OWN test files, run against vmad_b200's re-implemented tape VM.

Only possible where /root/reference exists (the build container); skipped elsewhere.  The
reference's test modules are copied to a scratch directory at test time (never into the repo)
and collected with tests/compat/alias.py mapping ``vmad.core.*`` onto ``vmad_b200.vm`` -- so
the reference's operator libraries (vmad/lib/linalg.py, unary.py, binary.py, mpi.py and
fastpm.py) execute unmodified on our VM, with the numpy oracle standing in for pmesh.
"""
import os
import shutil
import subprocess
import sys

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'vmad')),
                                reason="the reference tree is only present in the build container")

SUITES = {
    'core': ['core/tests/test_operator.py', 'core/tests/test_model.py',
             'core/tests/test_autooperator.py', 'core/tests/test_stdlib.py',
             'core/tests/test_error.py', 'core/tests/test_symbol.py'],
    'lib': ['lib/tests/test_linalg.py', 'lib/tests/test_unary.py', 'lib/tests/test_binary.py',
            'lib/tests/test_mpi.py', 'tests/test_vmad3.py'],
    'fastpm': ['lib/tests/test_fastpm.py'],
}


@pytest.mark.parametrize('suite', sorted(SUITES))
def test_reference_suite_on_our_vm(tmp_path, suite):
    files = []
    for rel in SUITES[suite]:
        dst = tmp_path / ('ref_' + rel.replace('/', '_'))
        shutil.copy(os.path.join(REF, 'vmad', rel), dst)
        files.append(str(dst))
    shutil.copy(os.path.join(ROOT, 'tests', 'compat', 'alias.py'), tmp_path / 'alias.py')
    env = dict(os.environ, VMAD_B200_ROOT=ROOT, VMAD_REFERENCE=REF, PYTHONPATH=str(tmp_path))
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-p', 'alias', '-x',
                        '-W', 'ignore', '--rootdir', str(tmp_path)] + files,
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert ' passed' in r.stdout and ' failed' not in r.stdout, tail
