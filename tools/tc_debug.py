"""Developer helper (GPU): prints small tcgen05 probe outputs for operand-layout debugging."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_tc import run_probe  # noqa
np.set_printoptions(linewidth=200, precision=1, suppress=True)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N, Kd = 64, 32
def go(A, B, name):
    Ain = A if mode == 0 else np.ascontiguousarray(A.T)
    Bin = B if mode == 0 else np.ascontiguousarray(B.T)
    D = run_probe(mode, 0, N, Kd, Ain, Bin)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    print(f"--- {name}: max|D|={np.abs(D).max():.1f} zeros={np.mean(D==0):.2f} maxerr={np.abs(D-ref).max():.1f}")
    print("D[:6,:8]=\n", D[:6, :8]); print("D[32:36,:4]=\n", D[32:36,:4]); print("ref[:2,:8]=", ref[:2,:8])
    return D
m = np.arange(128, dtype=np.float32)[:, None]; n = np.arange(N, dtype=np.float32)[:, None]; k = np.arange(Kd, dtype=np.float32)[None, :]
one_a = np.ones((128, Kd), np.float32); one_b = np.ones((N, Kd), np.float32)
go(one_a, one_b, "ones")
go(one_a * m, one_b, "A=m")
go(one_a, one_b * n, "B=n")
go(one_a * k, one_b, "A=k")
A = np.zeros((128, Kd), np.float32); A[5, 3] = 1
go(A, one_b * 0 + (k + 1), "A onehot(5,3), B=k+1")
