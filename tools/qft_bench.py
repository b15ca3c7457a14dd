"""BASELINE.json configs[2]: n-qubit QFT (PennyLane decomposition) on one B200; analytic check
amp[k] = 2^{-n/2} exp(2 pi i x k / 2^n) on the first 2^16 indices."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pennylane_lightning_b200 as plb
from pennylane_lightning_b200 import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
dtype = np.complex128 if (len(sys.argv) < 3 or sys.argv[2] == "c128") else np.complex64
free, total = torch.cuda.mem_get_info()
need = (1 << n) * (16 if dtype == np.complex128 else 8)
print(f"free {free/2**30:.1f} GiB, state needs {need/2**30:.1f} GiB", flush=True)
if need > free * 0.98:
    raise SystemExit("state does not fit")
rng = np.random.default_rng(7)
x = int(rng.integers(0, 1 << 62)) % (1 << n)
bits = [(x >> (n - 1 - w)) & 1 for w in range(n)]
ops = circuits.qft(n)
sv = plb.StateVector(n, dtype, 0, torch.cuda.current_stream().cuda_stream)
res = {}
for fuse in (True, False):
    sv.set_basis_state(bits, list(range(n)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sv.apply_ops(ops, fuse=fuse)
    sv.sync()
    dt = time.perf_counter() - t0
    head = sv.get_state(1 << 16)
    k = np.arange(1 << 16, dtype=np.float64)
    # exact phase: (x*k mod 2^n) / 2^n with integer arithmetic
    ph = np.array([((x * int(kk)) % (1 << n)) / float(1 << n) for kk in range(1 << 16)])
    exact = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ph)
    err = float(np.max(np.abs(head - exact)) / 2.0 ** (-n / 2))
    res["fused" if fuse else "unfused"] = dict(seconds=dt, gates=len(ops), gates_per_s=len(ops) / dt,
                                                stats=sv.last_apply_stats(), max_rel_err=err)
    print(n, "fused" if fuse else "unfused", res["fused" if fuse else "unfused"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/qft_{n}_{'c128' if dtype == np.complex128 else 'c64'}.json", "w"))
