"""Build libmonodetr_b200.so IN-TREE for sm_100a with plain nvcc (no torch, no JIT cache).

    python -m monodetr_b200.build [--force]

Every csrc/*.cu is compiled to an object (cached by mtime under csrc/_obj/) and linked into
monodetr_b200/libmonodetr_b200.so.  nvcc cross-compiles without a GPU, so this runs in the
authoring container; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(HERE, "libmonodetr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"] + (["-DMDB_TIMELINE"] if os.environ.get("MDB_TIMELINE") else [])   # profiling build: tools/diag_timeline.py


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, force):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    deps = [src] + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    if force or _newer(obj, deps):
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        return obj, True
    return obj, False


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if not srcs:
        raise RuntimeError("no CUDA sources found")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    rebuilt = [os.path.basename(o) for o, r in results if r]
    if rebuilt or _newer(SO, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", SO, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(f"[monodetr_b200.build] compiled {rebuilt or 'nothing'}; linked {SO}")
    elif verbose:
        print(f"[monodetr_b200.build] up to date: {SO}")
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv)
