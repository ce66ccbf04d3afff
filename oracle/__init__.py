"""CPU oracle for the vmad particle-mesh (FastPM) hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``vmad_b200/`` imports it, and
the product path raises if its CUDA library is missing -- it never falls back
to this code.

What it restates
----------------
``/root/reference/vmad/lib/fastpm.py`` performs all of its arithmetic through
the third-party package ``pmesh`` (``from pmesh.pm import ParticleMesh``,
fastpm.py:6) and takes its scalar growth factors from
``fastpm.background.MatterDominated`` (fastpm.py:5).  Neither package is in
``/root/reference`` nor installed (they are un-pinned conda dependencies,
``.travis.yml:21``), so the reference's own implementation of the path cannot
be executed.  ``oracle/pmesh`` and ``oracle/fastpm`` restate, in numpy fp64,
exactly the object protocol that fastpm.py calls under
those import names, so that the *reference's own* ``vmad/lib/fastpm.py`` runs
unmodified on top of it (``oracle.activate()``).

PARITY STATUS: "parity unpinned" at the pmesh boundary.  The reference's tests
hold no stored numeric vectors; what they do pin --
finite-difference vjp/jvp consistency and a handful of closed-form identities
(``vmad/lib/tests/test_fastpm.py``) -- is re-hosted in ``tests/`` and passes
on this oracle.  The golden vectors under ``tests/golden`` were produced by the
reference's ``vmad/lib/fastpm.py`` operator wiring running on this oracle
(``tests/golden/make_golden.py``).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def activate(shims=True):
    """Make ``import pmesh.pm`` / ``import fastpm.background`` resolve to the oracle.

    With ``shims`` the minimal stand-ins for ``mpi4py.MPI.COMM_SELF``,
    ``nbodykit.cosmology.Planck15`` and ``abopt.abopt2`` (all absent from this
    image) are made importable as well, so the reference test modules import.
    """
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    if shims:
        p = os.path.join(_HERE, 'shims')
        if p not in sys.path:
            sys.path.insert(0, p)
