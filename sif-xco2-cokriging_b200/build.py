"""Build libcokrig_b200.so (sm_100a only) in-tree with nvcc.

    python sif-xco2-cokriging_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The library lands next to the ctypes binding
(cokrig_b200/libcokrig_b200.so); it is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "cokrig_b200", "libcokrig_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["ck_matern.cu", "ck_chol.cu", "ck_vario.cu", "ck_local.cu", "ck_mg.cu", "ck_mgctx.cu", "ck_ozaki.cu", "ck_xcor.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _nccl_include() -> list:
    """ck_mgctx.cu needs <nccl.h> for types and prototypes only (the library is bound at run time with dlopen): the system
    header if there is one, else the one that ships with torch's NCCL wheel."""
    if os.path.exists("/usr/include/nccl.h"):
        return []
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            inc = os.path.join(root, "include")
            if os.path.exists(os.path.join(inc, "nccl.h")):
                return ["-I", inc]
    except Exception:  # noqa: BLE001
        pass
    return []


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "cokrig.h"))
    files.append(os.path.abspath(__file__))
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant / defines: a tuning or profiling build next to the product library (tools only), e.g.
    build(variant="prof", defines=("CK_LOCAL_PROFILE",)) -> cokrig_b200/libvariant_prof.so, selected at run time with
    COKRIG_B200_LIB=<path>."""
    out = OUT if not variant else os.path.join(HERE, "cokrig_b200", f"libvariant_{variant}.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _deps_mtime():
        return out
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", (f"_{variant}" if variant else "") + ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *_nccl_include(), *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # --cudart shared: the CUDA runtime is NOT embedded in the product binary (it resolves to the libcudart.so.12 that
    # torch has already loaded, so the library and torch share one runtime instance)
    cmd = [nvcc, "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    var = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var, defines=defs))
