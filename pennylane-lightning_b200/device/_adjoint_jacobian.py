"""Adjoint-Jacobian class of the ``lightning.b200`` device (counterpart of lightning_gpu/_adjoint_jacobian.py:57-180):
serialised tape -> AdjointJacobianC64/C128 of ``lightning_b200_ops``; ``batch_obs`` splits the observables over the
GPUs of the box (AdjointJacobian.batched)."""
import numpy as np
from pennylane_lightning.lightning_base._adjoint_jacobian import LightningBaseAdjointJacobian

from .. import lightning_b200_ops as _ops


class LightningB200AdjointJacobian(LightningBaseAdjointJacobian):
    def __init__(self, qubit_state, batch_obs=False):
        self._dp = _ops.DevPool()
        super().__init__(qubit_state, batch_obs)

    def _adjoint_jacobian_dtype(self):
        alg = _ops.algorithms
        if self.dtype == np.complex64:
            return alg.AdjointJacobianC64(), alg.create_ops_listC64
        return alg.AdjointJacobianC128(), alg.create_ops_listC128

    def calculate_jacobian(self, tape):
        if self._handle_raises(tape, is_jacobian=True):
            return np.array([], dtype=self.dtype)
        split_obs = self._dp.getTotalDevices() if self._batch_obs else False
        data = self._process_jacobian_tape(tape, split_obs, False)
        if not data:  # no trainable parameters
            return np.array([], dtype=self.dtype)
        tp = data["tp_shift"]
        call = self._jacobian_lightning.batched if self._batch_obs else self._jacobian_lightning
        jac = np.array(call(data["state_vector"], data["obs_serialized"], data["ops_serialized"], tp))
        nonempty = bool(len(jac))
        # rows of split observables (one engine row per Hamiltonian chunk) are summed back per measured observable
        rows = np.asarray(data["obs_indices"])
        n_obs = len(np.unique(rows))
        jac = jac.reshape((len(rows), -1))
        summed = np.zeros((n_obs, jac.shape[1]), dtype=jac.dtype)
        np.add.at(summed, rows, jac)
        jac = summed.reshape(-1, len(tp)) if nonempty else summed
        # columns go back to the positions of the tape's trainable parameters
        full = np.zeros((jac.shape[0], data["all_params"]))
        full[:, data["record_tp_rows"]] = jac
        return self._adjoint_jacobian_processing(full)

    # calculate_vjp is inherited: LightningBaseAdjointJacobian turns dy into ONE weighted observable and calls
    # calculate_jacobian (lightning_base/_adjoint_jacobian.py:229-288) - a single adjoint sweep on the device.
