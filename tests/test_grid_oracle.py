"""CPU: neat_b200.grid.grid_points (the point order and float arithmetic neat_sdf_grid reproduces in-kernel, checked
against each other on the GPU in test_sdf_grid_vs_oracle_and_module) vs the unmodified reference's get_grid_uniform
(oracle/make_golden_grid.py -> tests/golden/grid.npz)."""
import os

import numpy as np

from neat_b200 import grid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid.npz")


def test_grid_points_match_reference_bit_exact():
    g = np.load(GOLD)
    assert np.array_equal(grid.grid_points(7, (-1.3, 1.7)).numpy(), g["points_7"])
    big = grid.grid_points(100, (-1.5, 1.5))
    assert np.array_equal(big.numpy()[::997], g["points_100_stride997"])
    assert np.array_equal(big.double().sum(0).numpy(), g["points_100_sum"])
    # the reference evaluates the 10^6-point grid in 10 chunks of 100,000 (plots.py:106); neat_sdf_grid is one launch
    assert g["chunks_100"].tolist() == [100000] * 10 and g["chunks_7"].tolist() == [343]
