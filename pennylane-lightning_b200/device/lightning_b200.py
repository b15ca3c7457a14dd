"""The ``lightning.b200`` PennyLane device (counterpart of lightning_gpu/lightning_gpu.py:124-429)."""
from dataclasses import replace
from functools import partial
from pathlib import Path

import numpy as np
import pennylane as qp
from pennylane.devices import ExecutionConfig
from pennylane.devices.capabilities import OperatorProperties
from pennylane.devices.modifiers import simulator_tracking, single_tape_support
from pennylane.devices.preprocess import (
    decompose,
    device_resolve_dynamic_wires,
    validate_device_wires,
    validate_measurements,
    validate_observables,
)
from pennylane.exceptions import DeviceError
from pennylane.ops import MidMeasure
from pennylane.transforms import defer_measurements, dynamic_one_shot
from pennylane_lightning.lightning_base.lightning_base import LightningBase, resolve_mcm_method
from pennylane_lightning.lightning_base._adjoint_jacobian import adjoint_transforms, supports_adjoint  # noqa: F401

from .. import lightning_b200_ops as _ops
from ._adjoint_jacobian import LightningB200AdjointJacobian
from ._measurements import LightningB200Measurements
from ._state_vector import LightningB200StateVector

# operators executed through their matrix (no named kernel)
_to_matrix_ops = {name: OperatorProperties(controllable=(name == "BlockEncode")) for name in (
    "BlockEncode", "ControlledQubitUnitary", "ECR", "ISWAP", "SISWAP", "SQISW", "OrbitalRotation", "QubitCarry",
    "QubitSum", "DiagonalQubitUnitary")}


def stopping_condition(op, allow_mcms=True):
    """Whether ``lightning.b200`` executes the operator natively (mid-circuit measurements are never decomposed)."""
    if isinstance(op, MidMeasure):
        return allow_mcms
    return _supports_operation(op.name)


allow_mcms_stopping_condition = partial(stopping_condition, allow_mcms=True)
no_mcms_stopping_condition = partial(stopping_condition, allow_mcms=False)


def check_gpu_resources():
    if not _ops.DevPool.getTotalDevices():
        raise ValueError("No supported CUDA-capable device found")
    if not _ops.is_gpu_supported():
        raise ValueError(f"CUDA device is an unsupported version: {_ops.get_gpu_arch()} (lightning.b200 needs sm_100)")


@simulator_tracking
@single_tape_support
class LightningB200(LightningBase):
    """PennyLane device backed by the B200-native state-vector engine.

    Args:
        wires: number (or labels) of the wires; ``None`` sizes the register per circuit
        c_dtype: ``np.complex128`` (default) or ``np.complex64``
        seed: seed of the sampler
        batch_obs: split the observables of an adjoint Jacobian over the GPUs of the box
    """

    _device_options = ("c_dtype", "batch_obs")
    _new_API = True
    _CPP_BINARY_AVAILABLE = True
    _backend_info = staticmethod(_ops.backend_info)
    config_filepath = Path(__file__).parent / "lightning_b200.toml"
    pennylane_requires = ">=0.41"

    def __init__(self, wires=None, *, c_dtype=np.complex128, shots=None, seed="global", batch_obs=False,
                 use_async=False):
        check_gpu_resources()
        super().__init__(wires=wires, c_dtype=c_dtype, shots=shots, seed=seed, batch_obs=batch_obs)
        self._use_async = use_async
        self._statevector = None
        self._sv_init_kwargs = {"use_async": use_async}

    @property
    def name(self):
        return "lightning.b200"

    def _set_lightning_classes(self):
        self.LightningStateVector = LightningB200StateVector
        self.LightningMeasurements = LightningB200Measurements
        self.LightningAdjointJacobian = LightningB200AdjointJacobian

    def setup_execution_config(self, config=None, circuit=None):
        config = config or ExecutionConfig()
        for option in config.device_options:
            if option not in self._device_options:
                raise DeviceError(f"device option {option} not present on {self}")
        adjointish = config.gradient_method in ("best", "adjoint")
        updated = {}
        if config.gradient_method == "best":
            updated["gradient_method"] = "adjoint"
        if config.use_device_jacobian_product is None:
            updated["use_device_jacobian_product"] = adjointish
        if config.use_device_gradient is None:
            updated["use_device_gradient"] = adjointish
        if (config.use_device_gradient or updated.get("use_device_gradient")) and config.grad_on_execution is None:
            updated["grad_on_execution"] = True
        options = dict(config.device_options)
        for option in self._device_options:
            options.setdefault(option, getattr(self, f"_{option}", None))
        updated["mcm_config"] = resolve_mcm_method(config.mcm_config, circuit, "lightning.b200")
        return replace(config, **updated, device_options=options)

    def preprocess_transforms(self, execution_config=None):
        cfg = execution_config or ExecutionConfig()
        pipeline = qp.CompilePipeline()
        gate_set = self.capabilities.gate_set()
        deferred = cfg.mcm_config.mcm_method == "deferred"
        stop = no_mcms_stopping_condition if deferred else allow_mcms_stopping_condition
        if not deferred:
            gate_set |= {"MidMeasureMP"}
        pipeline.add_transform(validate_measurements, name=self.name)
        pipeline.add_transform(validate_observables, self.capabilities.supports_observable, name=self.name)
        if deferred:
            pipeline.add_transform(defer_measurements, allow_postselect=False)
        pipeline.add_transform(decompose, stopping_condition=stop, skip_initial_state_prep=True, name=self.name,
                               device_wires=self.wires, target_gates=gate_set)
        pipeline.add_transform(device_resolve_dynamic_wires, wires=self.wires, allow_resets=not deferred)
        pipeline.add_transform(validate_device_wires, self.wires, name=self.name)
        if cfg.mcm_config.mcm_method == "one-shot":
            pipeline.add_transform(dynamic_one_shot, postselect_mode=cfg.mcm_config.postselect_mode)
        pipeline.add_transform(qp.transforms.broadcast_expand)
        if cfg.gradient_method == "adjoint":
            pipeline += adjoint_transforms(self, not deferred)
        return pipeline

    def execute(self, circuits, execution_config=None):
        cfg = execution_config or ExecutionConfig()
        results = []
        for circuit in circuits:
            if self._wire_map is not None:
                [circuit], _ = qp.map_wires(circuit, self._wire_map)
            results.append(self.simulate(self.dynamic_wires_from_circuit(circuit), self._statevector,
                                         postselect_mode=cfg.mcm_config.postselect_mode,
                                         mcm_method=cfg.mcm_config.mcm_method))
        return tuple(results)

    def supports_derivatives(self, execution_config=None, circuit=None):
        if execution_config is None and circuit is None:
            return True
        if execution_config and execution_config.gradient_method in {"adjoint", "best"}:
            return supports_adjoint(self, circuit)
        return False

    @staticmethod
    def get_c_interface():
        # the Catalyst QuantumDevice plugin is not part of this repository
        raise NotImplementedError("lightning.b200 has no Catalyst runtime plugin")


_supports_operation = LightningB200.capabilities.supports_operation
