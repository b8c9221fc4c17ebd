"""Builds libmvtb.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Each csrc/*.cu is compiled to an object file (in parallel; objects are kept under csrc/_obj and reused while
newer than every source and header), then linked into one shared library."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libmvtb.so")
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "..", "include", "mvtb.h")]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    deps = sources() + _headers()
    return all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps)


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _compile_one(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    deps = [src] + _headers()
    if os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in deps):
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (os.path.basename(src), res.stdout, res.stderr))
    return obj, res.stderr


def build_library(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile_one(s, verbose), sources()))
    res = subprocess.run([_nvcc(), "-shared", "-o", OUT] + [o for o, _ in results] + ["-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(log for _, log in results))
    return OUT


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
