"""Developer microbenchmark (GPU): the hardware floor of tcgen05.mma -- groups of 8 fully unrolled instructions inside one
elect.sync region (csrc/diag/mma_bench.cu: mma_bench_unrolled_kernel; SASS = UTCHMMA back to back).  Markdown table."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import diag  # noqa: E402

fn = diag.lib().ainmf_diag_mma_bench_unrolled
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_void_p]
torch.zeros(1, device="cuda")
print("| CTAs | kind | A from | M | N | accumulators | cycles/MMA (to completion) | cycles/MMA (issue only) | law max(M,128)*N/256 |")
print("|---|---|---|---|---|---|---|---|---|")
for blocks in (1, 148):
    for bf16 in (0, 1, 2, 3):
        for ts, n_accs in ((1, (1, 2)), (0, (1,))):
            if bf16 == 3 and ts == 0:
                continue
            for M in (128, 64):
                if M == 64 and (ts or bf16 >= 2):
                    continue
                for N in (16, 32, 64, 128, 256):
                    for n_acc in (n_accs if bf16 != 3 else (1,)):
                        if n_acc * N > 480:
                            continue
                        iters = 128
                        out = torch.zeros(2 * blocks, dtype=torch.int64, device="cuda")
                        for _ in range(2):
                            rc = fn(M, N, bf16, ts, n_acc, iters, blocks, out.data_ptr(), None)
                            assert rc == 0, rc
                            torch.cuda.synchronize()
                        o = out.cpu().numpy().reshape(-1, 2)
                        n = iters * 8
                        print(f"| {blocks} | {['tf32 K8', 'bf16 K16', 'alternating tf32/bf16', '4 tf32 then 4 bf16'][bf16]} | {'TMEM' if ts else 'smem'} | {M} | {N} | {n_acc} | {np.median(o[:, 0]) / n:.1f} | {np.median(o[:, 1]) / n:.1f} | {max(M, 128) * N / 256:.0f} |", flush=True)
