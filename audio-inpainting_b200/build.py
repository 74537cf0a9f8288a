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
SOURCES = ["api.cu", "comm.cu", "stft.cu", "mask.cu", "impute.cu", "nmf_cd.cu", "pcm.cu", "gaps.cu", "tc_host.cu", "tc_probe.cu", "nmf_tc.cu", "nmf_ts.cu", "sweep_test.cu"]
LIB = os.path.join(HERE, "libainmf.so")
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
        deps += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    return deps


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into one shared library (static cudart)."""
    if not force and not _newer(LIB, _all_deps()):
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        cmd = [NVCC, *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-lcudart_static", "-ldl", "-lpthread", "-lrt"])
    return LIB


def build_emulator(force: bool = False) -> str:
    """TEST HARNESS: the same sources compiled for the host against csrc/emu/cuda_emu.h."""
    if not force and not _newer(EMU_LIB, _all_deps()):
        return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DAINMF_EMU", "-x", "c++",
           "-Wno-unknown-pragmas", "-Wno-attributes", "-fno-strict-aliasing"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", EMU_LIB]
    subprocess.check_call(cmd)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emulator(force=True))
    else:
        print(build_library(force=True, verbose="-v" in sys.argv))
