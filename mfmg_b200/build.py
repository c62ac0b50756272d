"""In-tree builds: the sm_100a CUDA library (C ABI), the host-setup helper and the CPU oracle.

All artefacts are built next to their sources so they travel with the `gpurun` snapshot:
  mfmg_b200/csrc/libmfmg_b200.so          nvcc -gencode arch=compute_100a,code=sm_100a
  mfmg_b200/hostsetup/libmfmg_b200_host.so g++ -fopenmp   (setup / problem generation only)
  oracle/libmfmg_oracle.so, oracle/libstdrand.so  gcc/g++ (test infrastructure)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mfmg_b200", "csrc")
HOSTSETUP = os.path.join(ROOT, "mfmg_b200", "hostsetup")
ORACLE = os.path.join(ROOT, "oracle")

CUDA_LIB = os.path.join(CSRC, "libmfmg_b200.so")
HOST_LIB = os.path.join(HOSTSETUP, "libmfmg_b200_host.so")
ORACLE_LIB = os.path.join(ORACLE, "libmfmg_oracle.so")
STDRAND_LIB = os.path.join(ORACLE, "libstdrand.so")


def _host_cxx() -> str:
    # /opt/gcc/bin/g++ (the CXX in this image's environment) lacks libgomp.spec; the distro
    # compiler has OpenMP.
    for cand in ("/usr/bin/g++", shutil.which("g++") or ""):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no g++ found")


def _host_cc() -> str:
    for cand in ("/usr/bin/gcc", shutil.which("gcc") or ""):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no gcc found")


def _nvcc() -> str:
    for cand in ("/usr/local/cuda/bin/nvcc", shutil.which("nvcc") or ""):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(sources: list[str]) -> str:
    h = hashlib.sha256()
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _newer(target: str, sources: list[str]) -> bool:
    """True if `target` exists and was built from exactly these source contents (a content
    stamp, not mtimes: the gpurun snapshot does not preserve mtimes)."""
    stamp = target + ".stamp"
    if not (os.path.exists(target) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _digest(sources)


def _stamp(target: str, sources: list[str]) -> None:
    with open(target + ".stamp", "w") as f:
        f.write(_digest(sources))


def _run(cmd: list[str], cwd: str | None = None) -> None:
    res = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))


def cuda_sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "mfmg_b200.h"))
    if not force and _newer(CUDA_LIB, deps):
        return CUDA_LIB
    nccl_inc = "/usr/include"
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-ccbin", _host_cxx(), "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", nccl_inc,
        "--expt-relaxed-constexpr", "--threads", "0",
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", CUDA_LIB] + srcs + ["-lnccl"]
    _run(cmd)
    _stamp(CUDA_LIB, deps)
    return CUDA_LIB


def build_host(force: bool = False) -> str:
    src = os.path.join(HOSTSETUP, "assemble.cpp")
    if not force and _newer(HOST_LIB, [src]):
        return HOST_LIB
    _run([_host_cxx(), "-O3", "-march=x86-64-v2", "-fPIC", "-fopenmp", "-std=c++17", "-shared",
          "-o", HOST_LIB, src])
    _stamp(HOST_LIB, [src])
    return HOST_LIB


def build_oracle(force: bool = False) -> str:
    src_c = os.path.join(ORACLE, "mfmg_oracle.c")
    src_r = os.path.join(ORACLE, "stdrand.cpp")
    if force or not _newer(ORACLE_LIB, [src_c]):
        _run([_host_cc(), "-O2", "-fPIC", "-fopenmp", "-std=gnu11", "-ffp-contract=off", "-shared",
              "-o", ORACLE_LIB, src_c, "-lm"])
        _stamp(ORACLE_LIB, [src_c])
    if force or not _newer(STDRAND_LIB, [src_r]):
        _run([_host_cxx(), "-O2", "-fPIC", "-std=c++17", "-shared", "-o", STDRAND_LIB, src_r])
        _stamp(STDRAND_LIB, [src_r])
    return ORACLE_LIB


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)
    build_oracle(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", CUDA_LIB, HOST_LIB, ORACLE_LIB, STDRAND_LIB)
