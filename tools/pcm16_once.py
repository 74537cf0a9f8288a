"""Developer helper (GPU): one ainmf_inpaint_host_pcm16 call and one ainmf_inpaint_host call on a small stereo batch in two
chunks (a small case to run on its own)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ainmf
from ainmf import _capi
L = ainmf._lib.lib(); h = ainmf._lib.handle(0); vp = C.c_void_p
rng = np.random.default_rng(3)
B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 300, 12000
t = np.arange(N) / 8000.0
P = np.empty((B, N, 2), np.int16)
for b in range(B):
    x = np.sin(2 * np.pi * rng.uniform(200, 1500) * t) + 0.05 * rng.standard_normal(N)
    x = x / np.abs(x).max() * 0.9
    if b % 5:
        s0 = int(rng.integers(500, N - 3000)); x[s0:s0 + 1500] = 0
    P[b, :, 0] = np.round(x * 32767); P[b, :, 1] = np.round(0.5 * x * 32767)
pin = torch.from_numpy(P).pin_memory(); pout = torch.zeros((B, N), dtype=torch.int16).pin_memory()
peak, nb, er, ni = np.zeros(B, np.float32), np.zeros(B, np.int32), np.zeros(B, np.float32), np.zeros(B, np.int32)
p = _capi.default_params(L, batch=B, n_samples=N, n_fft=256, hop=64, rank=16, max_iter=6, tol=0.0, seed=42, threshold=1e-4, frac_num=9, frac_den=10)
for cap in (0, 40 << 20):
    rc = L.ainmf_inpaint_host_pcm16(h, C.byref(p), vp(pin.data_ptr()), 2, vp(pout.data_ptr()), peak.ctypes.data_as(vp), nb.ctypes.data_as(vp), er.ctypes.data_as(vp), ni.ctypes.data_as(vp), cap)
    ainmf._lib.check(rc, 0)
    print("pcm16 cap", cap, "ok: restored clips", int((nb > 0).sum()), "peak", float(peak.max()), "out max", int(np.abs(pout.numpy()).max()))
xh = (torch.from_numpy(P[:, :, 0].astype(np.float32)) / 32767.0).pin_memory(); yh = torch.zeros((B, N)).pin_memory()
rc = L.ainmf_inpaint_host(h, C.byref(p), vp(xh.data_ptr()), vp(yh.data_ptr()), nb.ctypes.data_as(vp), er.ctypes.data_as(vp), ni.ctypes.data_as(vp), 0)
ainmf._lib.check(rc, 0)
print("float ok: restored clips", int((nb > 0).sum()))
