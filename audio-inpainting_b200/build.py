"""Build recipe for libainmf.so (sm_100a) and, for the CPU test-suite only, the kernel-logic emulator.

    python audio-inpainting_b200/build.py            # nvcc -> audio-inpainting_b200/libainmf.so
    python audio-inpainting_b200/build.py --emu      # g++  -> tests/_emu/libainmf_emu.so (test harness)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "comm.cu", "stft.cu", "mask.cu", "impute.cu", "nmf_cd.cu", "nmf_wside.cu", "nmf_mukl.cu", "nmf_coop.cu", "nmf_small.cu", "rng.cu", "pcm.cu", "gaps.cu", "tc_host.cu", "nmf_tc.cu", "nmf_ts.cu"]
# diagnostics (descriptor probe, sweep unit test, tcgen05 issue-rate microbenchmark): their own library, loaded by
# tests/ and tools/ only -- nothing of it is linked into or exported from the product library
DIAG_SOURCES = ["diag/tc_probe.cu", "diag/sweep_test.cu", "diag/mma_bench.cu", "tc_host.cu"]
LIB = os.path.join(HERE, "libainmf.so")
DIAG_LIB = os.path.join(HERE, "libainmf_diag.so")
EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libainmf_emu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _all_deps() -> list[str]:
    deps = [os.path.join(ROOT, "include", "ainmf.h"), os.path.abspath(__file__)]
    for d, _, files in os.walk(CSRC):
        deps += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp", ".inc"))]
    return deps


def _compile(sources: list[str], objdir: str, extra: list[str], logname: str, verbose: bool) -> list[str]:
    """nvcc -c every source that is newer than its object (or whose headers are), in parallel."""
    os.makedirs(objdir, exist_ok=True)
    hdrs = [d for d in _all_deps() if not d.endswith(".cu")]
    objs, procs = [], []
    for src in sources:
        obj = os.path.join(objdir, src.replace("/", "_") + ".o")
        objs.append(obj)
        if not _newer(obj, hdrs + [os.path.join(CSRC, src)]):
            continue
        cmd = [NVCC, *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    if procs:
        with open(os.path.join(objdir, logname), "a") as f:
            f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return objs


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into one shared library (static cudart), plus the diagnostics library."""
    objdir = os.path.join(HERE, "build")
    if force:
        import shutil
        shutil.rmtree(objdir, ignore_errors=True)
    if force or _newer(LIB, _all_deps()):
        objs = _compile(SOURCES, objdir, [], "ptxas.log", verbose)
        subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-lcudart_static", "-ldl", "-lpthread", "-lrt"])
    if force or _newer(DIAG_LIB, _all_deps()):
        objs = _compile(DIAG_SOURCES, os.path.join(objdir, "diag"), ["-DAINMF_DIAG_LIB"], "ptxas.log", verbose)
        subprocess.check_call([NVCC, "-shared", "-o", DIAG_LIB, *objs, "-lcudart_static", "-ldl", "-lpthread", "-lrt"])
    return LIB


def build_emulator(force: bool = False) -> str:
    """TEST HARNESS: the same sources compiled for the host against csrc/emu/cuda_emu.h."""
    if not force and not _newer(EMU_LIB, _all_deps()):
        return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DAINMF_EMU", "-x", "c++",
           "-Wno-unknown-pragmas", "-Wno-attributes", "-fno-strict-aliasing"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES + [d for d in DIAG_SOURCES if d not in SOURCES]]
    cmd += ["-o", EMU_LIB]
    subprocess.check_call(cmd)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emulator(force=True))
    else:
        print(build_library(force=True, verbose="-v" in sys.argv))
