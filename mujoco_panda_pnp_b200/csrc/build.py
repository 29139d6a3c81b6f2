"""Build libpnp_b200.so in-tree with nvcc for sm_100a.

    python -m mujoco_panda_pnp_b200.csrc.build [--force] [--verbose]

The specialised-kinematics header is regenerated first (tools/gen_spec_kinematics.py) so the
library always matches the packaged kinematic tree.
"""

from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libpnp_b200.so")
SOURCES = ["pnp_capi.cu"]
DEPS = ["pnp_capi.cu", "pnp_kernels.cuh", "pnp_common.cuh", "pnp_vec.cuh", "pnp_host_api.inc",
        os.path.join("generated", "spec_kinematics.cuh"), os.path.join(ROOT, "include", "pnp_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d if os.path.isabs(d) else os.path.join(HERE, d)) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    gen = os.path.join(ROOT, "tools", "gen_spec_kinematics.py")
    if os.path.exists(gen):
        # regenerate only when the content would change (keeps mtimes stable)
        if subprocess.call([sys.executable, gen, "--check"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) != 0:
            subprocess.check_call([sys.executable, gen])
    if not force and up_to_date():
        return OUT
    extra = os.environ.get("PNP_NVCC_EXTRA", "").split()  # development: -D switches of kernel variants
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", OUT, *SOURCES]
    res = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed, see output above")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
