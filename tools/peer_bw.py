"""NVLink peer-memory bandwidth of SM-issued copies (2 GPUs, torchrun): every rank pushes (remote stores) or pulls
(remote loads) half a slab to / from its partner's ping-pong slab at the same time.
  torchrun --nproc-per-node 2 tools/peer_bw.py [local_qubits]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pennylane_lightning_b200 import _capi
from pennylane_lightning_b200.dist import DistStateVector

nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
sv = DistStateVector(nloc + int(np.log2(world)), np.complex128)
eng = sv.engine
lib = _capi.lib()
lib.plb200_sv_alt_ptr.restype = C.c_void_p
mine = eng.sv.device_ptr
peer_alt = eng.peers_alt[rank ^ 1]
nbytes = (1 << (nloc - 1)) * 16
res = {}
for mode in ("push", "pull"):
    for unroll in (4, 8):
        ts = []
        for it in range(4):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if mode == "push":
                _capi._check(lib.plb200_sv_peer_copy(eng.sv._h, C.c_void_p(peer_alt), C.c_void_p(mine), C.c_int64(nbytes), unroll))
            else:
                _capi._check(lib.plb200_sv_peer_copy(eng.sv._h, C.c_void_p(mine), C.c_void_p(peer_alt), C.c_int64(nbytes), unroll))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ts[1:])], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"{mode}_u{unroll}"] = dict(ms=float(t.item()), GBps_per_direction=nbytes / (float(t.item()) * 1e-3) / 1e9)
if rank == 0:
    print(json.dumps(dict(local_qubits=nloc, bytes=nbytes, **res)))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/peer_bw_{nloc}q.json", "w"), indent=1)
sv.close()
dist.destroy_process_group()
