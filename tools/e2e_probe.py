"""Where the end-to-end step (bench.py e2e) spends its time: host marshalling, scheduling, GPU, expvals."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pennylane_lightning_b200 as plb
from pennylane_lightning_b200 import circuits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ops = circuits.random_circuit(n, 20, 1234)
sv = plb.StateVector(n, np.complex128, 0, torch.cuda.current_stream().cuda_stream)
sv.apply_ops(plb.OpsBlob(ops), fuse=True); torch.cuda.synchronize()
zw = [[w] for w in range(n)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); sv.reset(); ta = time.perf_counter()
    blob = plb.OpsBlob(ops); tb = time.perf_counter()
    sv.apply_ops(blob, fuse=True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ez = sv.expval_pauli_words_each(["Z"] * n, zw); t3 = time.perf_counter()
    print(f"reset call {1e3*(ta-t0):.1f} ms | OpsBlob {1e3*(tb-ta):.1f} | apply_ops call {1e3*(t1-tb):.1f} | "
          f"GPU drained at +{1e3*(t2-t1):.1f} | expvals {1e3*(t3-t2):.1f} | step total {1e3*(t3-t0):.1f} ms", flush=True)
