"""Developer microbenchmark: tcgen05.mma kind::tf32 issue/throughput (cycles per M128 x N x K8 instruction)."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ainmf
lib = ainmf._lib.lib(); ainmf._lib.handle(0)
fn = lib.ainmf_tc_mma_bench
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p]
for blocks in (148,):
    for N in (64, 128, 256):
        for ts in (1, 3):
            for per, commit_each in ((12, 0),):
                iters = 200
                out = torch.zeros(2 * blocks, dtype=torch.int64, device="cuda")
                for _ in range(2):
                    rc = fn(N, ts, iters, per, commit_each, blocks, out.data_ptr(), None)
                    assert rc == 0, rc
                    torch.cuda.synchronize()
                o = out.cpu().numpy().reshape(blocks, 2)
                n = iters * per
                print(f"blocks {blocks:3d} N {N:3d} {['SS tf32','TS tf32','SS bf16','TS bf16'][ts]} per {per:2d} commit {commit_each}: total {np.median(o[:,0])/n:7.1f} cyc/mma, issue {np.median(o[:,1])/n:7.1f} cyc/mma")
