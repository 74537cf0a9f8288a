"""Developer helper (GPU): timing of the incremental coordinate sweep in isolation (csrc/diag/sweep_test.cu)."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import diag
lib = diag.lib()
fn = lib.ainmf_test_sweep_inc_timed
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
for KP in (64, 128):
    for mult in (1, 2, 4):
        rows = 128 * 148 * mult
        g = torch.Generator(device="cuda").manual_seed(0)
        A = torch.rand((rows, KP), device="cuda", generator=g)
        Hm = torch.rand((KP + 40, KP), device="cuda", generator=g)
        G = (Hm.T @ Hm).contiguous()
        Bm = torch.rand((rows, KP), device="cuda", generator=g) * KP
        viol = torch.zeros(rows // 128, device="cuda")
        cyc = torch.zeros(rows // 128, dtype=torch.int64, device="cuda")
        for _ in range(2):
            rc = fn(A.data_ptr(), G.data_ptr(), Bm.data_ptr(), rows, KP, viol.data_ptr(), cyc.data_ptr(), None)
            torch.cuda.synchronize()
        c = cyc.cpu().numpy()
        print(f"KP={KP} ctas={rows//128} ({mult}/SM): sweep cycles per CTA median {np.median(c):.0f} min {c.min()} max {c.max()} -> per step {np.median(c)/KP:.0f}")
