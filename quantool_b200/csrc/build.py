"""Builds quantool_b200/lib/libquantool_b200.so (the C-ABI library) with nvcc for sm_100a.

In-tree build so the .so travels to the GPU box with the snapshot.  nvcc cross-compiles here
without a GPU.  `python -m quantool_b200.csrc.build [--force]`.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OBJ = os.path.join(PKG, "lib", "obj")
SO = os.path.join(PKG, "lib", "libquantool_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
          "-I", HERE, "--expt-relaxed-constexpr"]

# per-file extra flags.  gguf.cu: bit-exact llama.cpp operation order => no FMA contraction.
SOURCES = {
    "common.cu": [],
    "gguf.cu": ["--fmad=false"],
    "quant.cu": [],
    "linalg.cu": [],
    "gptq.cu": [],
    "hessian.cu": [],
    "split.cu": [],
    "smooth.cu": [],
    "awq.cu": [],
    "forward.cu": [],
    "tgemm.cu": [],
}


def _register(name, flags=()):
    SOURCES[name] = list(flags)


def _stamp(path, flags):
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".cuh", ".h")) or f == os.path.basename(path):
            with open(os.path.join(HERE, f), "rb") as fh:
                h.update(fh.read())
    with open(os.path.join(ROOT, "include", "quantool_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _compile(name, flags, force, verbose):
    src = os.path.join(HERE, name)
    obj = os.path.join(OBJ, name.replace(".cu", ".o"))
    stamp_file = obj + ".stamp"
    allflags = ARCH + COMMON + flags
    stamp = _stamp(src, allflags)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False
    cmd = [NVCC] + allflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {name}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, True


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    present = {n: f for n, f in SOURCES.items() if os.path.exists(os.path.join(HERE, n))}
    with ThreadPoolExecutor(max_workers=min(8, len(present))) as ex:
        results = list(ex.map(lambda kv: _compile(kv[0], kv[1], force, verbose), present.items()))
    objs = [o for o, _ in results]
    if force or any(c for _, c in results) or not os.path.exists(SO):
        cmd = [NVCC] + ARCH + ["-shared", "-o", SO] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
