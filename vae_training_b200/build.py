"""Builds libvae_b200.so (sm_100a) in-tree with nvcc.  Run: python -m vae_training_b200.build

Each .cu is compiled to an object in parallel, then linked; objects live in csrc/_obj/
(git-ignored).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.environ.get("VAE_B200_LIB") or os.path.join(HERE, "libvae_b200.so")      # override: A/B builds
if os.environ.get("VAE_B200_LIB"):
    OBJ = OBJ + "_" + os.path.basename(LIB).replace(".", "_")
HEADER = os.path.join(HERE, "..", "include", "vae_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2",
] + [f"-D{k}={os.environ[k]}" for k in ("VAE_SMALL_MAXB", "VAE_SMALL_GM_MODELS") if os.environ.get(k)] \
  + [f"-D{d}" for d in os.environ.get("VAE_B200_DEFINES", "").split() if d]      # e.g. VAE_TC_PROFILE (A/B + profiling builds)


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libvae_b200.so cannot be built")
    return nvcc


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER]
    return max(os.path.getmtime(h) for h in hs)


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return src, proc.returncode, proc.stdout + proc.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = find_nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    todo, objs = [], []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm):
            todo.append((src, obj))
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            results = list(ex.map(lambda so: _compile(nvcc, so[0], so[1], verbose), todo))
        for src, rc, out in results:
            if verbose or rc != 0:
                sys.stderr.write(out)
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}")
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("nvcc failed linking libvae_b200.so")
    build_xla_ffi_shim()
    return LIB


def build_xla_ffi_shim(include_dir: str = "") -> str:
    """g++ of csrc/xla_ffi_shim.cc -> libvae_b200_xla_ffi.so (the jax.ffi handlers over the C ABI).  The handlers
    compile only where jaxlib's xla/ffi/api/ffi.h is found (include_dir, or `jax.ffi.include_dir()` when jax is
    importable); otherwise the library exports just vae_xla_ffi_available() == 0."""
    src = os.path.join(CSRC, "xla_ffi_shim.cc")
    out = os.path.join(HERE, "libvae_b200_xla_ffi.so")
    if not include_dir:
        try:
            import jax.ffi as jffi             # absent in this image: the shim then builds without handlers
            include_dir = jffi.include_dir()
        except Exception:
            include_dir = ""
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(HEADER)) and not include_dir:
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", src, "-o", out]
    if include_dir:
        cmd += ["-I", include_dir, "-L", HERE, "-l:libvae_b200.so", "-Wl,-rpath,$ORIGIN"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("g++ failed on xla_ffi_shim.cc")
    return out


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
