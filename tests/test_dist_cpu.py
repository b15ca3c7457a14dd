"""world_size 2 and 4 under gloo on CPU: the sharding / permutation / swap scheduling of dist.py with a
numpy local engine, against the single-process oracle on the same circuit."""
import numpy as np
import pytest

from dist_helpers import mixed_circuit, multicall_tapes, run_ranks, run_ranks_measure, run_ranks_multicall, check_measurements
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


@pytest.mark.parametrize("world,n,seed", [(4, 7, 3), (4, 6, 6), (4, 6, 7)])
def test_multicall_schedule_is_rank_independent(world, n, seed):
    """Regression (these seeds made the ranks' wire -> bit maps diverge): a controlled gate whose control AND target are global must be deferred on every rank
    (the control-value skip is applied only when the op executes), so the wire -> bit map stays equal."""
    res = run_ranks_multicall(world, n, seed, port=29650 + seed)
    assert all(p == res["phys"][0] for p in res["phys"]), res["phys"]
    ref = np_oracle.StateVector(n)
    for t in multicall_tapes(n, seed):
        ref.apply_ops(t)
    np.testing.assert_allclose(res["state"], ref.get_state(), rtol=0, atol=1e-12)


def test_schedule_only_many_seeds():
    """The scheduling half of apply_ops (which ops are deferred, which wires swap) run for every rank of a
    world in ONE process with the data movement stubbed out: the wire -> bit map must not depend on the rank."""
    import torch

    from dist_helpers import NumpyEngine
    from pennylane_lightning_b200.dist import DistStateVector

    class FakeDist:
        def __init__(self, w, r):
            self.w, self.r = w, r

        def get_world_size(self, g=None):
            return self.w

        def get_rank(self, g=None):
            return self.r

    for seed in range(1, 25):
        for world, n in ((4, 6), (2, 6), (4, 7), (8, 7)):
            maps = []
            for r in range(world):
                sv = DistStateVector.__new__(DistStateVector)
                sv.dist, sv.torch, sv.group = FakeDist(world, r), torch, None
                sv.world, sv.rank, sv.g = world, r, int(np.log2(world))
                sv.n, sv.nloc, sv.dtype = n, n - sv.g, np.dtype(np.complex128)
                sv.engine = NumpyEngine(sv.nloc, np.complex128)
                sv.n_swaps = sv.swap_bytes = 0
                sv.swap_mode = "nccl"

                def fake_swap(gw, lw, sv=sv):
                    sv.phys[gw], sv.phys[lw] = sv.phys[lw], sv.phys[gw]

                sv._swap = fake_swap
                sv.reset()
                for t in multicall_tapes(n, seed):
                    sv.apply_ops(t)
                maps.append(list(sv.phys))
            assert all(m == maps[0] for m in maps), (world, n, seed, maps)


@pytest.mark.parametrize("world,n,seed", [(2, 7, 1), (4, 8, 2)])
def test_sharded_measurements(world, n, seed):
    """Pauli-word / Hamiltonian expval (X/Y on global wires are swapped in), probs marginals in the given wire
    order, and sampling on the sharded state, against the single-process oracle on the gathered state."""
    check_measurements(run_ranks_measure(world, n, seed, "gloo", port=29690 + seed), n)


@pytest.mark.parametrize("world,n,seed", [(2, 6, 3), (4, 7, 4)])
def test_sharded_adjoint_jacobian(world, n, seed):
    """Adjoint Jacobian with lambda, H lambda and mu sharded over the ranks and kept on one wire map
    (AdjointJacobianGPUMPI.hpp semantics) against the single-process oracle's adjoint loop."""
    from dist_helpers import check_adjoint, run_ranks_adjoint

    check_adjoint(run_ranks_adjoint(world, n, seed, "gloo", port=29730 + seed), n, seed)
