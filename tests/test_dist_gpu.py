"""Sharded execution on >=2 GPUs (NCCL): full amplitudes vs the single-GPU engine and the reference."""
import numpy as np
import pytest

from dist_helpers import check_measurements, mixed_circuit, run_ranks, run_ranks_measure

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("swap", ["peer", "peer-nofuse", "peer-multi", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_gpu_matches_single_gpu_and_reference(plb, ref, world, swap):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    n, seed = 18, 10 + world
    res = run_ranks(world, n, seed, "nccl", port=29700 + world + {"peer": 10, "peer-multi": 20, "nccl": 0, "peer-nofuse": 30}[swap], swap=swap)
    ops = mixed_circuit(n, seed)
    single = plb.StateVector(n)
    single.apply_ops(ops, fuse=True)
    r = ref.StateVector(n)
    r.apply_ops(ops)
    np.testing.assert_allclose(res["state"], r.get_state(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(res["state"], single.get_state(), rtol=0, atol=1e-12)
    assert abs(res["norm2"] - 1.0) < 1e-12
    assert res["swaps"] > 0
    if swap == "peer":
        assert res["fused_swaps"] == res["swaps"]  # every exchange rode on a pass's store phase


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_measurements_gpu(world):
    """expval of Pauli words / a Hamiltonian (X, Y on global wires swapped in through a routed op-less pass),
    probs marginals and sampling on a state sharded over GPUs."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    n = 16
    check_measurements(run_ranks_measure(world, n, 3, "nccl", port=29760 + world), n)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_adjoint_jacobian_gpu(world):
    """Adjoint Jacobian with lambda / H lambda / mu sharded over GPUs (routed exchanges keep the vectors on one
    wire map) against the oracle's single-process adjoint loop."""
    from dist_helpers import check_adjoint, run_ranks_adjoint

    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    n, seed = 14, 5
    check_adjoint(run_ranks_adjoint(world, n, seed, "nccl", port=29780 + world), n, seed)
