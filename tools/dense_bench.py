"""Dense k-qubit matrix (QubitUnitary) throughput: tensor-core path vs the scalar mat-vec kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pennylane_lightning_b200 as plb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
out = {}
for dtype, tag, ks in ((np.complex128, "c128", (5, 6)), (np.complex64, "c64", (5, 6, 7))):
    sv = plb.StateVector(n, dtype, 0, torch.cuda.current_stream().cuda_stream)
    S = (1 << n) * (16 if dtype == np.complex128 else 8)
    for k in ks:
        rng = np.random.default_rng(k)
        a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        u, _ = np.linalg.qr(a)
        wires = list(range(2, 2 + k))
        for mode in ("mma", "scalar"):
            if mode == "scalar":
                os.environ["PLB200_DENSE_MMA"] = "0"
            else:
                os.environ.pop("PLB200_DENSE_MMA", None)
            sv.apply_matrix(u, wires)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                sv.apply_matrix(u, wires)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            flops = 8.0 * (1 << k) * (1 << n)  # 4 D real FMA per amplitude
            out[f"{tag}_k{k}_{mode}"] = dict(ms=ms, hbm_frac=2 * S / (ms * 1e-3) / 1e9 / 6550.1, tflops=flops / (ms * 1e-3) / 1e12)
            print(tag, k, mode, out[f"{tag}_k{k}_{mode}"], flush=True)
    del sv
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/dense_bench_{n}q.json", "w"), indent=1)
