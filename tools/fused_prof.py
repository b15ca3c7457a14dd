"""Small driver for ncu: one fused pass sequence of the config-2 circuit family at n qubits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pennylane_lightning_b200 as plb
from pennylane_lightning_b200 import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
dtype = np.complex128 if (len(sys.argv) < 3 or sys.argv[2] == "c128") else np.complex64
fuse = not (len(sys.argv) > 3 and sys.argv[3] == "nofuse")
ops = circuits.random_circuit(n, 20, 1234)
sv = plb.StateVector(n, dtype, 0, torch.cuda.current_stream().cuda_stream)
blob = plb.OpsBlob(ops)
sv.apply_ops(blob, fuse=fuse)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sv.apply_ops(blob, fuse=fuse)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"n={n} fuse={fuse} gates={len(ops)} stats={sv.last_apply_stats()} {ms:.2f} ms -> {len(ops)/ms*1e3:.1f} gates/s")
