"""world_size 2 and 4 under gloo on CPU: the sharding / permutation / swap scheduling of dist.py with a
numpy local engine, against the single-process oracle on the same circuit."""
import numpy as np
import pytest

from dist_helpers import mixed_circuit, run_ranks
from oracle import np_oracle


@pytest.mark.parametrize("world,n,seed", [(2, 7, 1), (2, 8, 2), (4, 8, 3)])
def test_sharded_matches_single_process(world, n, seed):
    res = run_ranks(world, n, seed, "gloo", port=29611 + seed)
    ref = np_oracle.StateVector(n)
    ref.apply_ops(mixed_circuit(n, seed))
    np.testing.assert_allclose(res["state"], ref.get_state(), rtol=0, atol=1e-12)
    assert abs(res["norm2"] - 1.0) < 1e-12
    for w in range(n):
        assert abs(res["z"][w] - ref.expval_named("PauliZ", [w])) < 1e-12
    assert res["swaps"] > 0
