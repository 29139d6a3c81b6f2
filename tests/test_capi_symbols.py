"""The C-ABI library loads without a GPU and exports every symbol include/pnp_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from mujoco_panda_pnp_b200 import _lib
from mujoco_panda_pnp_b200.tree import KinematicTree, PnpTreeStruct

HEADER = os.path.join(ROOT, "include", "pnp_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnp_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    decl = declared_functions()
    assert len(decl) >= 20
    assert sorted(_lib.SIGNATURES) == decl


def test_library_exports_every_declared_symbol(cuda_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (pnp_[a-z0-9_]+)", out))
    assert set(declared_functions()) <= exported
    for name in declared_functions():
        assert hasattr(cuda_lib, name)
    assert cuda_lib.pnp_abi_version() == 3


def test_struct_layouts_match_header(cuda_lib):
    # sizeof(PnpTree) = 8 + (21+63+3+9+7+7+7)*8
    assert ctypes.sizeof(PnpTreeStruct) == 8 + 117 * 8
    assert ctypes.sizeof(_lib.PnpIkParams) == 32
    assert ctypes.sizeof(_lib.PnpRewardParams) == 40
    # the specialised tree baked into the library is the packaged asset's tree
    s = PnpTreeStruct()
    assert cuda_lib.pnp_get_specialized_tree(ctypes.byref(s)) == 0
    assert bytes(s) == bytes(KinematicTree.from_mjcf().to_struct())


def test_header_constants_match_binding():
    """#define values of the header (kinematics selectors, flag bits, counter slots) = the binding's constants."""
    text = open(HEADER).read()
    defs = {m.group(1): int(m.group(2).rstrip("u")) for m in re.finditer(r"#define\s+(PNP_[A-Z0-9_]+)\s+(-?\d+u?)\b", text)}
    for name in ("PNP_KIN_AUTO", "PNP_KIN_GENERIC", "PNP_KIN_SPECIALIZED", "PNP_KIN_SPEC_LANE", "PNP_KIN_SPEC_PAIR",
                 "PNP_IK_CONVERGED", "PNP_IK_SUCCESS"):
        assert defs[name] == getattr(_lib, name), name
    assert _lib.KINEMATICS == {"auto": 0, "generic": 1, "specialized": 2, "spec_lane": 3, "spec_pair": 4}
    assert [defs[f"PNP_IK_CNT_{k}"] for k in ("N", "CONVERGED", "SUCCESS", "ITERATIONS")] == [0, 1, 2, 3]


def test_packed_sass_is_present():
    """The shipped library really contains Blackwell's packed FP32 instructions and the bulk-copy engine ops
    the design relies on (FFMA2/FMUL2/FADD2, UBLKCP) - a build that silently lost them would still pass parity."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.check_output([cuobjdump, "-sass", _lib.LIB_PATH], text=True)
    for op in ("FFMA2", "FMUL2", "FADD2", "UBLKCP", "SYNCS"):
        assert op in sass, op
    assert sass.count("FFMA2") > 400


def test_argument_errors_do_not_need_a_gpu(cuda_lib):
    assert cuda_lib.pnp_set_tree(None) == -1
    assert b"NULL" in cuda_lib.pnp_last_error()
    assert cuda_lib.pnp_get_specialized_tree(None) == -1
    assert cuda_lib.pnp_launch_count() >= 0


def test_sass_is_sm100a_only():
    """The shipped cubin targets sm_100a and nothing else (no multi-arch fat binary)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.check_output([cuobjdump, "-lelf", _lib.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
