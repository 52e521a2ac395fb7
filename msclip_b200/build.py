"""Build libmsclip_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m msclip_b200.build [--force]

The shared library is the product's only compute path: it is git-ignored but travels to the GPU box
with the working tree.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
# two builds of the same sources: MMA operands in bf16 (default) or IEEE fp16 (-DMSCLIP_FP16)
# MSCLIP_LIB_SUFFIX: build / load an experimental variant next to the default libraries (A/B runs, e.g. "_tanh")
SUFFIX = os.environ.get("MSCLIP_LIB_SUFFIX", "")
LIB_PATHS = {"bf16": os.path.join(HERE, f"libmsclip_b200{SUFFIX}.so"), "fp16": os.path.join(HERE, f"libmsclip_b200_fp16{SUFFIX}.so")}
LIB_PATH = LIB_PATHS["bf16"]
SOURCES = ["runtime.cu", "gemm.cu", "conv_gemm.cu", "elementwise.cu", "conv.cu", "front.cu", "attention.cu", "loss.cu", "engine.cu", "api.cu",
           "backward.cu", "wgrad.cu", "attention_bwd.cu", "engine_train.cu", "preprocess.cu", "tokenizer.cu"]
HEADERS = ["unicode_tables.inc", "common.cuh", "gemm_common.cuh", "rowops.cuh", "kernels.h", "engine.h", os.path.join("..", "..", "include", "msclip_b200.h"),
           os.path.join("..", "..", "include", "msclip_b200_ops.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _mtime(path: str) -> float:
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def needs_build(precision: str = "bf16") -> bool:
    lib = _mtime(LIB_PATHS[precision])
    if lib == 0.0:
        return True
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(_mtime(d) > lib for d in deps)


def build_all(force: bool = False, verbose: bool = False):
    return [build(force, verbose, p) for p in ("bf16", "fp16")]


def build(force: bool = False, verbose: bool = False, precision: str = "bf16") -> str:
    lib_path = LIB_PATHS[precision]
    if not force and not needs_build(precision):
        return lib_path
    obj_dir = os.path.join(OBJ_DIR, precision + SUFFIX)
    os.makedirs(obj_dir, exist_ok=True)
    flags = NVCC_FLAGS + (["-DMSCLIP_FP16"] if precision == "fp16" else [])
    if os.environ.get("MSCLIP_QGELU_EXP") == "1":       # A/B build: two-MUFU (ex2 + rcp) QuickGELU in the fc1 epilogue
        flags = flags + ["-DMSCLIP_QGELU_EXP"]
    nvcc = _nvcc()
    newest_header = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if not force and _mtime(obj) > max(_mtime(os.path.join(CSRC, src)), newest_header):
            return obj
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib_path] + objs + ["-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    for path in build_all(force="--force" in sys.argv, verbose="--quiet" not in sys.argv):
        print("built", path)
