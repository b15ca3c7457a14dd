"""Index-bit swap microbenchmark (2+ GPUs, torchrun): time of one global<->local bit swap per local bit
position, peer (in-place NVLink loads/stores) vs NCCL (pack -> send/recv -> unpack), as GB/s per direction
per GPU.   torchrun --nproc-per-node 2 tools/swap_bench.py [local_qubits]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pennylane_lightning_b200.dist import DistStateVector

nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
g = int(np.log2(world))
n = nloc + g
res = {}
for mode in ("peer", "nccl"):
    try:
        sv = DistStateVector(n, np.complex128, swap=mode)
    except Exception as exc:  # e.g. no room for NCCL staging buffers
        res[mode] = f"unavailable: {exc}"
        continue
    out = {}
    gw = 0  # wire 0 is global (top physical bit)
    for lb in (nloc - 1, nloc // 2, 3, 0):
        lw = next(w for w in range(n) if sv.phys[w] == lb)
        gw = next(w for w in range(n) if sv.phys[w] == n - 1)
        times = []
        for it in range(3):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sv._swap(gw, lw)
            e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            gw, lw = lw, gw  # swap back next time
        t = torch.tensor([min(times[1:])], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = (1 << (nloc - 1)) * 16
        out[f"bit{lb}"] = dict(ms=float(t.item()), GBps_per_direction=nbytes / (float(t.item()) * 1e-3) / 1e9)
    res[mode] = out
    del sv
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(dict(local_qubits=nloc, world=world, bytes_per_swap_per_gpu=(1 << (nloc - 1)) * 16, **res)))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/swap_bench_{nloc}q_{world}gpu.json", "w"))
dist.destroy_process_group()
