"""Small driver for ncu: fused pass sequence of the config-2 circuit family at n qubits.
argv: n [c128|c64] [fuse|nofuse] [reps]   (PLB200_JIT=sync|0 selects specialised kernels / interpreter)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pennylane_lightning_b200 as plb
from pennylane_lightning_b200 import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
dtype = np.complex128 if (len(sys.argv) < 3 or sys.argv[2] == "c128") else np.complex64
fuse = not (len(sys.argv) > 3 and sys.argv[3] == "nofuse")
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kind = os.environ.get("PLB200_PROF_TAPE", "random")
if kind == "light":   # one pass, one round, four cheap gates: the memory floor of the pass kernel
    ops = [circuits.op("RX", [w], [0.3 + w]) for w in (3, 9, 14, 20)]
elif kind == "light2":  # two rounds
    ops = [circuits.op("RX", [w], [0.3 + w]) for w in (3, 5, 7, 9, 11, 13, 15, 17)]
else:
    ops = circuits.random_circuit(n, 20, 1234)
sv = plb.StateVector(n, dtype, 0, torch.cuda.current_stream().cuda_stream)
blob = plb.OpsBlob(ops)
for _ in range(2):
    sv.apply_ops(blob, fuse=fuse)
torch.cuda.synchronize()
plb.jit_wait()
sv.apply_ops(blob, fuse=fuse)
torch.cuda.synchronize()
best = None
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sv.apply_ops(blob, fuse=fuse)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
S = (1 << n) * (16 if dtype == np.complex128 else 8)
passes = sv.last_apply_stats()[1]
print(f"n={n} {sys.argv[2] if len(sys.argv) > 2 else 'c128'} fuse={fuse} jit_mode={plb.jit_mode()} minb={os.environ.get('PLB200_JIT_MINB','-')} "
      f"gates={len(ops)} stats={sv.last_apply_stats()} {best:.2f} ms -> {len(ops)/best*1e3:.1f} gates/s; "
      f"hbm frac {passes*2*S/(best*1e-3)/1e9/6550.1:.3f}; jit {plb.jit_stats()}", flush=True)
