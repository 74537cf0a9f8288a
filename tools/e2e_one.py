"""Developer helper (GPU): ainmf_inpaint_host on one c2 / c3 clip, with the AINMF_HOST_TRACE timeline of the last call."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ainmf
from ainmf import _capi
L = ainmf._lib.lib(); h = ainmf._lib.handle(0)
for name in sys.argv[1:] or ["c2"]:
    wl = dict(bench.WORKLOADS[name]); N, K = wl["N"], wl["K"]
    xh = torch.from_numpy(bench.synth_host(wl, 0)[None].copy()).pin_memory()
    yh = torch.empty((1, N), dtype=torch.float32).pin_memory()
    nbh = np.zeros(1, np.int32); errh = np.zeros(1, np.float32); nih = np.zeros(1, np.int32)
    p = _capi.default_params(L, batch=1, n_samples=N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4,
                             seed=wl["seed"], threshold=wl["thr"], frac_num=wl["num"], frac_den=wl["den"])
    def step():
        rc = L.ainmf_inpaint_host(h, C.byref(p), C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()), nbh.ctypes.data_as(C.c_void_p),
                                  errh.ctypes.data_as(C.c_void_p), nih.ctypes.data_as(C.c_void_p), 0)
        ainmf._lib.check(rc, 0)
    for _ in range(3): step()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "e2e ms per call", [round(t, 2) for t in ts], flush=True)
    os.environ["AINMF_HOST_TRACE"] = "1"; step(); os.environ.pop("AINMF_HOST_TRACE")
