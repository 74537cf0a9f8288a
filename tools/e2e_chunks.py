"""Developer helper (GPU): end-to-end time of ainmf_inpaint_host on the c4 batch for several chunk sizes (AINMF_HOST_CHUNK),
next to the device-resident time of the same batch and of one chunk."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ainmf
from ainmf import _capi
L = ainmf._lib.lib(); h = ainmf._lib.handle(0)
wl = dict(bench.WORKLOADS["c4"]); B, N, K = wl["clips"], wl["N"], wl["K"]
dev = torch.device("cuda", 0)
x = bench.synth_device(wl, 0, B, dev)
def dev_step(xx):
    return ainmf.ops.nmf_inpaint(xx, wl["n_fft"], wl["hop"], K, 200, 1e-4, wl["seed"], wl["thr"], wl["num"], wl["den"], -1, -1, 1, None, None)
for nb in (512, 148, 74):
    xx = x[:nb]
    for _ in range(2): dev_step(xx)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): dev_step(xx)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"device-resident {nb} clips: {dt*1e3:.1f} ms = {dt*1e3/nb:.4f} ms/clip", flush=True)
xh = torch.empty((B, N), dtype=torch.float32).pin_memory(); xh.copy_(x)
yh = torch.empty((B, N), dtype=torch.float32).pin_memory()
t0 = time.perf_counter(); x.copy_(xh, non_blocking=True); torch.cuda.synchronize(); print(f"H2D 903 MB: {(time.perf_counter()-t0)*1e3:.1f} ms")
t0 = time.perf_counter(); yh.copy_(x, non_blocking=True); torch.cuda.synchronize(); print(f"D2H 903 MB: {(time.perf_counter()-t0)*1e3:.1f} ms")
del x; ainmf.ops._workspaces.clear(); torch.cuda.empty_cache()
nbh = np.zeros(B, np.int32); errh = np.zeros(B, np.float32); nih = np.zeros(B, np.int32)
p = _capi.default_params(L, batch=B, n_samples=N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4,
                         seed=wl["seed"], threshold=wl["thr"], frac_num=wl["num"], frac_den=wl["den"])
for ch in sys.argv[1:] or ["512", "296", "148", "74", "0"]:
    os.environ.pop("AINMF_HOST_CHUNK", None); os.environ.pop("AINMF_HOST_FIRST", None)
    if ch.startswith("f"): os.environ["AINMF_HOST_FIRST"] = ch[1:]      # fN: first chunk of N clips, the rest as by default
    elif ch != "0": os.environ["AINMF_HOST_CHUNK"] = ch
    def step():
        rc = L.ainmf_inpaint_host(h, C.byref(p), C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()), nbh.ctypes.data_as(C.c_void_p),
                                  errh.ctypes.data_as(C.c_void_p), nih.ctypes.data_as(C.c_void_p), 0)
        ainmf._lib.check(rc, 0)
    for _ in range(2): step()
    torch.cuda.synchronize(); t0 = time.perf_counter(); each = []
    for _ in range(3): t1 = time.perf_counter(); step(); each.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"e2e chunk {ch}: {dt*1e3:.1f} ms  (calls: {', '.join('%.1f' % e for e in each)})", flush=True)
    if os.environ.get("TRACE_ONE"):
        os.environ["AINMF_HOST_TRACE"] = "1"; step(); torch.cuda.synchronize(); os.environ.pop("AINMF_HOST_TRACE")
