"""Dump the specialised (NVRTC) source of every fused pass of a benchmark tape and, with --sass, compile it
with nvcc for sm_100a here (no GPU needed) and print registers / spills / instruction mix per pass."""
import collections, ctypes as C, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pennylane_lightning_b200 as plb
from pennylane_lightning_b200 import circuits

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
prec = 64 if (len(sys.argv) < 3 or sys.argv[2] == "c128") else 32
kind = sys.argv[3] if len(sys.argv) > 3 else "random"
out = sys.argv[4] if len(sys.argv) > 4 else "/tmp/plb_jit_dump"
os.makedirs(out, exist_ok=True)
ops = {"random": lambda: circuits.random_circuit(n, 20, 1234), "qft": lambda: circuits.qft(n),
       "sel": lambda: circuits.strongly_entangling_layers(n, 4, 42)[0]}[kind]()
blob = plb.OpsBlob(ops)
npass = C.c_int64()
lib = plb.lib()
rc = lib.plb200_jit_dump_sources(C.c_int64(n), prec, blob.ptr(), out.encode(), C.byref(npass))
assert rc == 0, lib.plb200_last_error()
print(f"{npass.value} passes -> {out}")
if "--sass" in sys.argv:
    for i in range(npass.value):
        src = f"{out}/pass_{i}.cu"
        r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
                            "-cubin", "-o", f"{out}/pass_{i}.cubin", src], capture_output=True, text=True)
        info = " ".join(l.strip() for l in r.stderr.splitlines() if "registers" in l or "spill" in l)
        sass = subprocess.run(["cuobjdump", "-sass", f"{out}/pass_{i}.cubin"], capture_output=True, text=True).stdout
        mix = collections.Counter()
        for l in sass.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                mix[m.group(1).split(".")[0]] += 1
        tot = sum(mix.values())
        print(f"pass {i}: {info}\n   SASS {tot} instr: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(12)))
