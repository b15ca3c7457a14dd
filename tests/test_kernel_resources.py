"""CPU-side guards on the compiled tile kernels (cuobjdump on the in-tree object file): the design points
DESIGN.md section 6 measured — 80 registers (24 warps/SM), a lean kernel of ~4k SASS instructions that
decodes ops on the uniform datapath — are properties of the binary and can regress silently."""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
OBJ = os.path.join(os.path.dirname(HERE), "pennylane-lightning_b200", "lib", "fusion.o")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(OBJ) and os.path.exists(CUOBJDUMP)),
                                reason="needs the built fusion.o and cuobjdump")


def _res_usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", OBJ], capture_output=True, text=True, check=True).stdout
    res = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            res[name] = (int(m.group(1)), int(m.group(2)))
    return res


def _kernels(res, precision, cfg, ext):
    key = f"tile_kernelI{'7double2' if precision == 64 else '6float2'}NS_4tile6{cfg}CfgIS2_EELb{int(ext)}"
    return {k: v for k, v in res.items() if key in k}


def test_forward_tile_kernels_fit_80_registers():
    res = _res_usage()
    for precision in (64, 32):
        for ext in (False, True):
            ks = _kernels(res, precision, "Fwd", ext)
            assert len(ks) == 1, (precision, ext, list(res))
            (reg, stack), = ks.values()
            assert reg <= 80, f"c{2 * precision} ext={ext}: {reg} registers (> 80: fewer than 24 warps/SM)"
            assert stack <= 256, f"c{2 * precision} ext={ext}: {stack} B of stack (spills)"


def test_adjoint_tile_kernels_exist_and_fit():
    res = _res_usage()
    for precision in (64, 32):
        for ext in (False, True):
            ks = _kernels(res, precision, "Adj", ext)
            assert len(ks) == 1
            (reg, _), = ks.values()
            assert reg <= 128


def test_lean_c128_kernel_is_small_and_uniform():
    sass = subprocess.run([CUOBJDUMP, "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    lean = [b for b in blocks if "tile_kernelI7double2NS_4tile6FwdCfgIS2_EELb0" in b.split("\n", 1)[0]]
    assert len(lean) == 1
    n_instr = len(re.findall(r"^\s*/\*[0-9a-f]{4,}\*/", lean[0], flags=re.M))
    assert n_instr < 5000, n_instr  # 3952 when measured; the extended kernel carries the rare op kinds
    assert lean[0].count("LDCU") > 50  # op descriptors are decoded on the uniform datapath
    assert "DFMA" in lean[0] and "BAR.SYNC" in lean[0]
