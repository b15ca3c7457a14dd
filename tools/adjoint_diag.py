"""Diagnostic: per-call time and JIT statistics of the 24q / 1000-parameter adjoint, bench.py's way (torch stream)
and tools/adjoint_bench.py's way (default stream)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pennylane_lightning_b200 as plb
from pennylane_lightning_b200 import circuits

n, n_params = 24, 1000
ops, tp = circuits.hardware_efficient_ansatz(n, n_params, 99)
co, words, wires = circuits.pauli_hamiltonian(n, 100, 99)
for way in ("torch_stream", "default_stream"):
    for dt in (np.complex128, np.complex64):
        if way == "torch_stream":
            sv = plb.StateVector(n, dt, torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        else:
            sv = plb.StateVector(n, dt)
        ham = circuits.hamiltonian_observable(plb, co, words, wires)
        blob = plb.OpsBlob(ops)
        sv.apply_ops(blob)
        sv.sync()
        for it in range(6):
            if it == 2:
                plb.jit_wait()
            s0 = plb.jit_stats()
            t0 = time.perf_counter()
            jac = sv.adjoint_jacobian([ham], blob, tp)
            dt_s = time.perf_counter() - t0
            s1 = plb.jit_stats()
            print(way, np.dtype(dt).name, it, f"{dt_s*1e3:.1f} ms", "jit+", s1["jit_launches"] - s0["jit_launches"], "interp+",
                  s1["interpreter_launches"] - s0["interpreter_launches"], "compiled", s1["compiled"], "seen", s1["structures_seen"], flush=True)
        del sv
