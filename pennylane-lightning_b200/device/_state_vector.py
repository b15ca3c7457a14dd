"""State-vector class of the ``lightning.b200`` device: routes PennyLane operations to StateVectorC64/C128 of
``lightning_b200_ops``.  Counterpart of pennylane_lightning/lightning_gpu/_state_vector.py:64-390 (without the MPI
handler: the sharded mode lives below the C ABI, pennylane_lightning_b200.dist)."""
import numpy as np
import pennylane as qp
from pennylane.exceptions import DeviceError
from pennylane.measurements import MidMeasureMP
from pennylane.ops import Conditional
from pennylane.ops.op_math import Adjoint
from pennylane.wires import Wires
from pennylane_lightning.lightning_base._state_vector import LightningBaseStateVector

from .. import lightning_b200_ops as _ops
from ._measurements import LightningB200Measurements

# operators whose cached matrix is keyed by a hash passed as the (single) parameter
_HASHED = ("BlockEncode", "ControlledQubitUnitary", "DiagonalQubitUnitary", "MultiControlledX", "OrbitalRotation",
           "PSWAP", "QubitUnitary")


class LightningB200StateVector(LightningBaseStateVector):
    """Holds a device-resident state vector; every gate call lands in the engine's lazy queue and the whole tape
    runs as fused passes at the first read."""

    def __init__(self, num_wires, dtype=np.complex128, rng=None, use_async=False):
        super().__init__(num_wires, dtype, rng)
        self._device_name = "lightning.b200"
        self._use_async = use_async
        self._num_local_wires = num_wires
        self._qubit_state = self._state_dtype()(num_wires)

    def _state_dtype(self):
        return _ops.StateVectorC128 if self.dtype == np.complex128 else _ops.StateVectorC64

    # ---- host <-> device
    def syncD2H(self, state_vector, use_async=False):
        self._qubit_state.DeviceToHost(state_vector.ravel(order="C"), use_async)

    def syncH2D(self, state_vector, use_async=False):
        self._qubit_state.HostToDevice(np.ascontiguousarray(state_vector, dtype=self.dtype).ravel(order="C"), use_async)

    @property
    def state(self):
        out = np.zeros(2 ** self._num_wires, dtype=self.dtype)
        self.syncD2H(out)
        return out

    @staticmethod
    def _operation_is_sparse(operation):  # no sparse-matrix gates on the device
        return False

    # ---- state preparation
    def _apply_state_vector(self, state, device_wires: Wires, **kwargs):
        if isinstance(state, self._qubit_state.__class__):
            raise DeviceError("lightning.b200 does not adopt an external device state vector.")
        state = state.toarray().ravel() if hasattr(state, "toarray") else np.asarray(state)
        if len(device_wires) == self._num_wires and Wires(sorted(device_wires)) == device_wires:
            self.syncH2D(state.reshape(-1))
            return
        self._qubit_state.setStateVector(np.ascontiguousarray(state, dtype=self.dtype).ravel(), list(device_wires),
                                         kwargs.get("use_async", False))

    # ---- gates
    @staticmethod
    def _params(op):
        if op.name == "PCPhase":  # the subspace dimension travels as a second parameter
            return np.array([op.parameters[0], float(op.hyperparameters["dimension"][0])])
        return op.parameters

    def _apply_lightning_controlled(self, operation, adjoint):
        base = operation.base
        if isinstance(base, Adjoint):
            base, adjoint = base.base, not adjoint
        cw, cv, tw = list(operation.control_wires), operation.control_values, list(operation.target_wires)
        method = getattr(self._qubit_state, base.name, None)
        if method is not None:
            method(cw, cv, tw, adjoint, self._params(base))
        else:
            self._qubit_state.applyControlledMatrix(qp.matrix(base), cw, cv, tw, adjoint)

    def _apply_lightning(self, operations, mid_measurements=None, postselect_mode=None):
        sv = self._qubit_state
        for op in operations:
            if isinstance(op, qp.Identity):
                continue
            base, inv = (op.base, True) if isinstance(op, Adjoint) else (op, False)
            wires = list(op.wires)
            if isinstance(op, Conditional):
                if op.meas_val.concretize(mid_measurements):
                    self._apply_lightning([op.base])
            elif isinstance(op, MidMeasureMP):
                self._apply_lightning_midmeasure(LightningB200Measurements(self).measure_final_state, op, mid_measurements,
                                                 postselect_mode=postselect_mode)
            elif isinstance(op, qp.PauliRot):
                word = op._hyperparameters["pauli_word"]  # pylint: disable=protected-access
                keep = [(w, p) for w, p in zip(wires, word) if p != "I"]
                sv.applyPauliRot([w for w, _ in keep], inv, op.parameters, "".join(p for _, p in keep))
            elif getattr(sv, base.name, None) is not None:
                getattr(sv, base.name)(wires, inv, self._params(base))
            elif isinstance(base, qp.ops.Controlled):
                self._apply_lightning_controlled(base, inv)
            else:  # anything with a matrix
                mat = qp.matrix(op)
                if len(mat) == 0:
                    raise ValueError("Unsupported operation")
                real = np.float32 if self.dtype == np.complex64 else np.float64
                param = [[real(op.hash)]] if op.name in _HASHED else []
                sv.apply(base.name, wires, False, param, np.ascontiguousarray(mat, dtype=self.dtype).ravel(order="C"))
