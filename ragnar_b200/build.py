"""In-tree build of libragnar_cuda.so (CUDA sm_100a + C-ABI) and the pybind11
module ``ragnar`` (host side mirroring the reference's Python interface).

    python -m ragnar_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  Outputs land next to this file
(git-ignored, but they travel to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOST = CSRC / "host"
OBJ = PKG / "_build"
LIB = PKG / "libragnar_cuda.so"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")
PYMOD = PKG / f"ragnar{EXT_SUFFIX}"

CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))
NVCC = str(CUDA_HOME / "bin" / "nvcc")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # unfused float/double arithmetic like the reference's default build; every
    # FMA in the kernels is an explicit fmaf()/fma()
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]
GXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-Wall"]
EXPORT = ["-DRGC_BUILDING"]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def _run(cmd, verbose):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    return proc.stdout + proc.stderr


def build(force: bool = False, verbose: bool = False) -> dict:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.hpp")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "ragnar_cuda.h"]
    inc = ["-I", str(ROOT / "include"), "-I", str(CSRC)]
    jobs = []
    objs = []
    for src in sorted(CSRC.glob("*.cu")):
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src, *headers]):
            jobs.append([NVCC, *NVCC_FLAGS, *inc, "-c", str(src), "-o", str(obj)])
    for src in sorted(CSRC.glob("*.cpp")):
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src, *headers]):
            jobs.append(["g++", *GXX_FLAGS, *inc, "-I", str(CUDA_HOME / "include"),
                         "-c", str(src), "-o", str(obj)])
    logs = {}
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        for cmd, out in zip(jobs, pool.map(lambda c: _run(c, verbose), jobs)):
            logs[Path(cmd[-1]).name] = out
    (OBJ / "ptxas.log").write_text("\n".join(f"== {k}\n{v}" for k, v in logs.items()))
    if force or jobs or not LIB.exists():
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs), "-cudart", "static",
              "-Xlinker", "--exclude-libs,ALL", "-ldl", "-lpthread", "-lz"], verbose)
    # pybind11 module
    import pybind11

    host_srcs = sorted(HOST.glob("*.cpp"))
    host_hdrs = list(HOST.glob("*.hpp")) + [ROOT / "include" / "ragnar_cuda.h"]
    if host_srcs and (force or not _newer(PYMOD, [*host_srcs, *host_hdrs, LIB])):
        host_objs = []
        hjobs = []
        for src in host_srcs:
            obj = OBJ / ("host_" + src.stem + ".o")
            host_objs.append(obj)
            if force or not _newer(obj, [src, *host_hdrs]):
                hjobs.append(["g++", *GXX_FLAGS, "-I", str(ROOT / "include"), "-I", str(HOST),
                              "-I", pybind11.get_include(),
                              "-I", sysconfig.get_paths()["include"],
                              "-c", str(src), "-o", str(obj)])
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
            list(pool.map(lambda c: _run(c, verbose), hjobs))
        _run(["g++", "-shared", "-o", str(PYMOD), *map(str, host_objs),
              "-L", str(PKG), "-lragnar_cuda", "-Wl,-rpath,$ORIGIN"], verbose)
    return {"lib": LIB, "module": PYMOD if host_srcs else None, "compiled": list(logs)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    info = build(a.force, a.verbose)
    print(info)
