"""Stand-in for nbodykit (absent): only ``cosmology.Planck15.Om0`` is needed
(/root/reference/vmad/lib/tests/test_fastpm.py:322,339).  TEST INFRASTRUCTURE ONLY."""
