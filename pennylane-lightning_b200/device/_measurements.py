"""Measurements class of the ``lightning.b200`` device (counterpart of lightning_gpu/_measurements.py:44-98):
binds MeasurementsC64/C128 of ``lightning_b200_ops`` and adds the fused Pauli-sentence expectation value."""
import numpy as np
import pennylane as qp
from pennylane_lightning.lightning_base._measurements import LightningBaseMeasurements

from .. import lightning_b200_ops as _ops


class LightningB200Measurements(LightningBaseMeasurements):  # pylint: disable=too-few-public-methods
    def __init__(self, qubit_state):
        super().__init__(qubit_state)
        self._measurement_lightning = self._measurement_dtype()(qubit_state.state_vector)
        if qubit_state._rng:  # pylint: disable=protected-access
            self._measurement_lightning.set_random_seed(int(qubit_state._rng.integers(0, 2**31 - 1)))

    def _measurement_dtype(self):
        return _ops.MeasurementsC128 if self.dtype == np.complex128 else _ops.MeasurementsC64

    def _expval_pauli_sentence(self, measurementprocess):
        """sum_k c_k <P_k> for an observable with a Pauli representation: one fused engine call."""
        pwords, coeffs = zip(*measurementprocess.obs.pauli_rep.items())
        return self._measurement_lightning.expval([qp.pauli.pauli_word_to_string(p) for p in pwords],
                                                  [p.wires.tolist() for p in pwords], np.real(coeffs))
