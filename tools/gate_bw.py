"""Per-gate HBM bandwidth sweep (un-fused kernels), gate type x target index bit."""
import json, sys
import numpy as np, torch
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pennylane_lightning_b200 as plb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dtype = np.complex128 if (len(sys.argv) < 3 or sys.argv[2] == "c128") else np.complex64
sv = plb.StateVector(n, dtype, 0, torch.cuda.current_stream().cuda_stream)
S = (1 << n) * (16 if dtype == np.complex128 else 8)
def bench(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
w = lambda bit: n - 1 - bit
res = []
for bit in [0, 1, 2, 3, 4, 5, 8, 12, 20, n - 1]:
    for name, mk, bytes_ in [
        ("RX", lambda b: sv.apply("RX", [w(b)], False, [0.3]), 2 * S),
        ("RZ", lambda b: sv.apply("RZ", [w(b)], False, [0.3]), 2 * S),
        ("Hadamard", lambda b: sv.apply("Hadamard", [w(b)]), 2 * S),
        ("PauliZ", lambda b: sv.apply("PauliZ", [w(b)]), S),
        ("CNOT(c=b,t=b+7)", lambda b: sv.apply("CNOT", [w(b), w((b + 7) % n)]), S),
        ("CNOT(c=b+7,t=b)", lambda b: sv.apply("CNOT", [w((b + 7) % n), w(b)]), S),
        ("CRZ(c=b+7,t=b)", lambda b: sv.apply("CRZ", [w((b + 7) % n), w(b)], False, [0.3]), S),
        ("SWAP(b,b+7)", lambda b: sv.apply("SWAP", [w(b), w((b + 7) % n)]), S),
        ("IsingXX(b,b+7)", lambda b: sv.apply("IsingXX", [w(b), w((b + 7) % n)], False, [0.3]), 2 * S),
    ]:
        ms = bench(lambda: mk(bit))
        res.append(dict(gate=name, bit=bit, ms=ms, GBs=bytes_ / ms / 1e6))
        print(f"{name:18s} bit={bit:2d} {ms:8.3f} ms  {bytes_/ms/1e6:8.1f} GB/s  frac={bytes_/ms/1e6/6550:.3f}", flush=True)
json.dump(res, open("gpurun_out/gate_bw_%d_%s.json" % (n, "c128" if dtype == np.complex128 else "c64"), "w"))
