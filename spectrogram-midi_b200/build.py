"""Build libaegis_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libaegis_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")
NVCC_FLAGS = [*os.environ.get("AEGIS_NVCC_EXTRA", "").split(), 
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr", "--extended-lambda",
]


def _nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: cannot build libaegis_b200.so")
    return path


def sources() -> list:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, "..", "include", "aegis_b200.h")]
    return any(os.path.getmtime(d) > lib_m for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        extra = ["-fmad=false"] if os.path.basename(src) in ("trend.cu", "notes_fin.cu") else []  # keep the reference's IEEE op order
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, sources()))
    log = "\n".join(f"== {o}\n{e}" for o, e in results)
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as f:
        f.write(log)
    if verbose:
        print(log)
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *[o for o, _ in results], "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    try:
        print("BUILD OK", build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    except Exception as e:   # the last line of the output always says how it went, whatever the caller pipes it through
        print(e, file=sys.stderr)
        print("BUILD FAILED")
        sys.exit(1)
