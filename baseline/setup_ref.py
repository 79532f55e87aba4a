"""Stage the UNMODIFIED reference under the git-ignored `baseline/_ref/` so that it travels to the GPU box.

    python baseline/setup_ref.py            # copy + pre-build the reference's CUDA extension (≈5 min, no GPU needed)
    python baseline/setup_ref.py --no-ext   # copy only

What it does:
  1. copies the reference's Python/C++/CUDA sources from /root/reference to baseline/_ref/GenS/ byte for byte
     (nothing under baseline/_ref is tracked by git: the copy is a run-time baseline, never product source);
  2. JIT-builds the reference's only native component, `gridsample_grad2`
     (models/modules/grid_sample_cuda/cuda_gridsample.py:5), for sm_100a exactly as the reference does
     (`torch.utils.cpp_extension.load`, CWD = the reference root) into baseline/_ref/_ext/, and leaves the
     resulting `gridsample_grad2.so` there.  nvcc cross-compiles without a GPU; the .so is git-ignored but is
     shipped by gpurun, so the GPU box never spends lease time on the 5-minute build.

`baseline/ref_runtime.py` is the loader the GPU tests and `bench.py` use: it imports the staged tree
with the pre-built extension (no rebuild on the box) and the harness shims the reference needs here
(pyhocon / mcubes are absent, MnasNet weights cannot be downloaded).
"""
from __future__ import annotations

import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
TREE = os.path.join(DST, "GenS")
EXT = os.path.join(DST, "_ext")

KEEP_DIRS = ("models", "utils", "confs")
KEEP_FILES = ("LICENSE.txt",)


def stage_sources() -> None:
    if not os.path.isdir(REF_SRC):
        raise SystemExit(f"{REF_SRC} is not present: run this in the build container")
    os.makedirs(DST, exist_ok=True)
    if os.path.isdir(TREE):
        shutil.rmtree(TREE)
    os.makedirs(TREE)
    for d in KEEP_DIRS:
        shutil.copytree(os.path.join(REF_SRC, d), os.path.join(TREE, d),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in KEEP_FILES:
        shutil.copy2(os.path.join(REF_SRC, f), os.path.join(TREE, f))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write("Unmodified copy of /root/reference (prstrive/GenS) made by baseline/setup_ref.py.\n"
                 "Run-time baseline only; git-ignored; never imported by gens_b200/.\n")


def build_extension() -> str:
    """Builds gridsample_grad2 the way the reference does (cpp_extension.load with CWD-relative sources)."""
    os.makedirs(EXT, exist_ok=True)
    os.environ["TORCH_EXTENSIONS_DIR"] = EXT
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    cwd = os.getcwd()
    os.chdir(TREE)
    try:
        from torch.utils import cpp_extension
        t0 = time.time()
        cpp_extension.load(
            name="gridsample_grad2",
            sources=["models/modules/grid_sample_cuda/gridsample_cuda.cpp",
                     "models/modules/grid_sample_cuda/gridsample_cuda.cu"],
            verbose=True)
        print(f"gridsample_grad2 built in {time.time() - t0:.0f} s")
    finally:
        os.chdir(cwd)
    so = os.path.join(EXT, "gridsample_grad2", "gridsample_grad2.so")
    assert os.path.exists(so), so
    return so


if __name__ == "__main__":
    stage_sources()
    if "--no-ext" not in sys.argv:
        print(build_extension())
