"""Randomised differential test of the ensemble kernel against the oracle (scripts/fuzz_forces.py):
random nuclei of 1..260 nucleons -- clustered, spread, with coincident and skip-range pairs --
with random strengths and time steps."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_nuclei_against_oracle(seed, ensemble_kernel):
    import fuzz_forces
    res = fuzz_forces.run(seed, trials=10)
    assert res["nuclei"] == 480
    assert res["worst_pos_err"] <= 1e-5, res
    assert res["worst_force_err_l2"] <= 1e-5, res
