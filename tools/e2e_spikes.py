"""Developer helper (GPU): 16 back-to-back ainmf_inpaint_host calls on the c4 batch, the time of each, and (AINMF_HOST_TIMES=1)
the host's setup / enqueue / drain split per call -- to see where an occasional slow call spends its time."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ainmf
from ainmf import _capi
os.environ["AINMF_HOST_TIMES"] = "1"
L = ainmf._lib.lib(); h = ainmf._lib.handle(0)
wl = dict(bench.WORKLOADS["c4"]); B, N, K = wl["clips"], wl["N"], wl["K"]
x = bench.synth_device(wl, 0, B, torch.device("cuda", 0))
xh = torch.empty((B, N), dtype=torch.float32).pin_memory(); xh.copy_(x)
yh = torch.empty((B, N), dtype=torch.float32).pin_memory()
del x; torch.cuda.empty_cache()
nbh = np.zeros(B, np.int32); errh = np.zeros(B, np.float32); nih = np.zeros(B, np.int32)
p = _capi.default_params(L, batch=B, n_samples=N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4,
                         seed=wl["seed"], threshold=wl["thr"], frac_num=wl["num"], frac_den=wl["den"])
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    t0 = time.perf_counter()
    rc = L.ainmf_inpaint_host(h, C.byref(p), C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()), nbh.ctypes.data_as(C.c_void_p),
                              errh.ctypes.data_as(C.c_void_p), nih.ctypes.data_as(C.c_void_p), 0)
    t1 = time.perf_counter()
    ainmf._lib.check(rc, 0)
    print(f"call {i}: {(t1 - t0)*1e3:.1f} ms (+ check {(time.perf_counter() - t1)*1e3:.2f} ms)", file=sys.stderr, flush=True)
