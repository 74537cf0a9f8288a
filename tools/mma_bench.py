"""Developer microbenchmark (GPU): cycles per tcgen05.mma as a function of N, M, kind, operand source and the number of
independent accumulators the instructions rotate over (csrc/diag/mma_bench.cu).  Prints a markdown table.

    python tools/mma_bench.py > gpurun_out/mma_bench.md
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import diag  # noqa: E402

fn = diag.lib().ainmf_diag_mma_bench
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 10 + [C.c_void_p, C.c_void_p]
torch.zeros(1, device="cuda")


def run(M, N, bf16, ts, n_acc, cg2, blocks, issue=0, iters=64, per=16):
    out = torch.zeros(2 * blocks + 2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        rc = fn(M, N, bf16, ts, n_acc, cg2, iters, per, blocks, issue, out.data_ptr(), None)
        assert rc == 0, rc
        torch.cuda.synchronize()
    o = out.cpu().numpy()[: 2 * (blocks // (2 if cg2 else 1))].reshape(-1, 2)
    n = iters * per
    return float(np.median(o[:, 0])) / n, float(np.median(o[:, 1])) / n


ISSUE = {0: "if (thread == 0) around the loop", 1: "whole warp, elect.sync per MMA", 2: "whole warp, one elect.sync around the loop"}
print("| CTAs | group | issue path | kind | A from | M | N | accumulators | cycles/MMA (to completion) | cycles/MMA (issue only) | law max(M,128)*N/256/cg |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for blocks in (148,):
    for issue in (0, 1, 2):
        for bf16 in (0, 1):
            for ts in (1, 0):
                for M in (128, 64):
                    for N in (32, 64, 128, 256):
                        for n_acc in (1, 2, 4):
                            if n_acc * N > 480:
                                continue
                            if M == 64 and (ts == 1 or n_acc == 2):
                                continue
                            tot, iss = run(M, N, bf16, ts, n_acc, 0, blocks, issue)
                            print(f"| {blocks} | 1 | {ISSUE[issue]} | {'bf16 K16' if bf16 else 'tf32 K8'} | {'TMEM' if ts else 'smem'} | {M} | {N} | {n_acc} | {tot:.1f} | {iss:.1f} | {max(M, 128) * N / 256:.0f} |", flush=True)
if "--cg2" in sys.argv:
    for blocks in (2, 148):
        for bf16 in (0, 1):
            for M in (128, 256):
                for N in (64, 128, 256):
                    for n_acc in (1, 2):
                        if n_acc * N > 480:
                            continue
                        tot, iss = run(M, N, bf16, 0, n_acc, 1, blocks)
                        print(f"| {blocks} | 2 | if (thread == 0) around the loop | {'bf16 K16' if bf16 else 'tf32 K8'} | smem | {M} | {N} | {n_acc} | {tot:.1f} | {iss:.1f} | {max(M // 2, 128) * N / 256 / 1:.0f} |", flush=True)
