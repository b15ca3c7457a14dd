"""PennyLane device ``lightning.b200``: the Python layer that lets PennyLane's lightning_base drive the B200 engine
(SURVEY 8 f1).  Sibling of pennylane_lightning/lightning_gpu/ in the reference tree: the classes of this package
plug the pybind11 module ``lightning_b200_ops`` into ``LightningBase`` / ``LightningBaseStateVector`` /
``LightningBaseMeasurements`` / ``LightningBaseAdjointJacobian``.

Importing the device class needs ``pennylane`` and ``pennylane_lightning.lightning_base`` (the reference's Python
layer); the engine itself (``pennylane_lightning_b200``) does not, so the import is deferred to first use.
"""


def __getattr__(name):
    if name == "LightningB200":
        from .lightning_b200 import LightningB200

        return LightningB200
    raise AttributeError(name)
