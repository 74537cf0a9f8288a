"""Tensor-core building blocks on the GPU: TMA + SWIZZLE_128B descriptors + tcgen05.mma kind::tf32 + TMEM loads,
through the diagnostic entry point ainmf_tc_probe (csrc/tc_probe.cu)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def trunc_tf32(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def run_probe(mode, split, N, Kd, A, B):
    import ainmf
    import diag
    lib = diag.lib()
    fn = lib.ainmf_tc_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    ainmf._lib.handle(0)
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    D = torch.zeros((128, N), dtype=torch.float32, device="cuda")
    rc = fn(mode, split, N, Kd, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), None)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return D.cpu().numpy()


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N", [64, 128])
def test_probe_single_pass_and_split(mode, N):
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required")
    rng = np.random.default_rng(mode * 10 + N)
    Kd = 96
    A = rng.standard_normal((128, Kd)).astype(np.float32)       # logical A [M][K], B [N][K]
    B = rng.standard_normal((N, Kd)).astype(np.float32)
    Ain = A if mode != 1 else np.ascontiguousarray(A.T)
    Bin = B if mode != 1 else np.ascontiguousarray(B.T)
    ref64 = A.astype(np.float64) @ B.astype(np.float64).T
    ref_tr = trunc_tf32(A).astype(np.float64) @ trunc_tf32(B).astype(np.float64).T
    scale = np.abs(ref64).max()
    D1 = run_probe(mode, 0, N, Kd, Ain, Bin)
    e_tr = np.abs(D1 - ref_tr).max() / scale
    e_64 = np.abs(D1 - ref64).max() / scale
    print(f"mode {mode} N {N}: single pass vs truncated-input ref {e_tr:.2e}, vs exact {e_64:.2e}")
    assert e_64 < 5e-3, "layout/descriptor error (not a rounding-sized difference)"
    D3 = run_probe(mode, 1, N, Kd, Ain, Bin)
    e3 = np.abs(D3 - ref64).max() / scale
    print(f"mode {mode} N {N}: 3-pass split vs exact {e3:.2e}")
    assert e3 < 2e-6


def bf16_rn(a):
    """float32 -> bfloat16 (round to nearest even) as uint16"""
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r


@pytest.mark.parametrize("N", [64, 128])
def test_probe_tf32_main_plus_bf16_cross_terms(N):
    """A from TMEM; a*b ~= a_hi*b_hi (tf32) + [a_lo | a_hi].[b_hi ; b_lo] (one bf16 K=16 instruction per 8 elements)."""
    import ainmf
    rng = np.random.default_rng(N)
    Kd = 96
    A = rng.standard_normal((128, Kd)).astype(np.float32)
    B = rng.standard_normal((N, Kd)).astype(np.float32)
    Bhi = trunc_tf32(B)
    Blo = (B - Bhi).astype(np.float32)
    # cross tile: per group of 8 k: 16 bf16 = [b_hi(8), b_lo(8)] packed into 8 float-sized words
    g = Kd // 8
    bh = bf16_rn(np.ascontiguousarray(B)).reshape(N, g, 8)
    bl = bf16_rn(np.ascontiguousarray(Blo)).reshape(N, g, 8)
    Bx = np.concatenate([bh, bl], axis=2).reshape(N, g * 16).copy().view(np.float32).reshape(N, Kd)
    import diag
    lib = diag.lib()
    fn = lib.ainmf_tc_probe_x
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    ainmf._lib.handle(0)
    Ad, Bd, Bxd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(Bx).cuda()
    D = torch.zeros((128, N), dtype=torch.float32, device="cuda")
    assert fn(3, 1, N, Kd, Ad.data_ptr(), Bd.data_ptr(), Bxd.data_ptr(), D.data_ptr(), None) == 0
    torch.cuda.synchronize()
    ref64 = A.astype(np.float64) @ B.astype(np.float64).T
    e = np.abs(D.cpu().numpy() - ref64).max() / np.abs(ref64).max()
    print(f"N {N}: tf32 + bf16 cross terms vs exact {e:.2e}")
    assert e < 4e-6
