"""CPU: vmad_b200's own VM + operator library, running on the numpy oracle backend, reproduce
the golden vectors that the REFERENCE vmad + reference fastpm.py produced on the same oracle.
This pins the host logic (tape VM, autodiff, operator wiring, growth factors) without a GPU."""
import pytest

from helpers import load_golden, run_golden_case


@pytest.mark.parametrize("name", ["nbody_8_2step", "nbody_16_3step"])
def test_golden_on_oracle_backend(oracle_pm, name):
    g = load_golden(name)
    pm = oracle_pm([int(g['N'])] * 3, g['BoxSize'])
    errs = run_golden_case(pm, g, tol=1e-10)
    print(errs)
