"""CPU checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol declared in
include/vdetr_b200.h; the product refuses to run without CUDA tensors (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import vdetr_b200._C as C
    header = open(os.path.join(ROOT, "include", "vdetr_b200.h")).read()
    declared = set(re.findall(r"\b(vdetr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = C.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vdetr_b200.h but not exported"
    assert declared == set(C.EXPORTS), (declared ^ set(C.EXPORTS))
    assert lib.vdetr_version().decode().endswith("sm_100a")
    assert "bad argument" in lib.vdetr_error_string(0x1001).decode()


def test_ops_reject_cpu_tensors():
    from vdetr_b200 import ops
    import vdetr_b200.pointnet2_utils as pu
    q = torch.zeros(1, 4, 4, 64)
    with pytest.raises(RuntimeError):
        ops.rpe_attention(q, torch.zeros(1, 8, 1, 64), torch.zeros(1, 8, 1, 64))
    with pytest.raises(RuntimeError):
        pu.ball_query(0.2, 4, torch.zeros(1, 8, 3), torch.zeros(1, 2, 3))
    with pytest.raises(RuntimeError):
        pu.grouping_operation(torch.zeros(1, 2, 8), torch.zeros(1, 2, 2, dtype=torch.int32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "v-detr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("CPU oracle is too slow", ""), f"{f} mentions the oracle"


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/vdetr_b200.h must compile as C (no C++, no torch or CUDA types in the signatures)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "vdetr_b200.h"\nint main(void) { VdetrXattnShape s; VdetrPeerCtx c; (void)s; (void)c; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-c", str(src), "-I", os.path.join(ROOT, "include"), "-o",
                        str(tmp_path / "abi.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
