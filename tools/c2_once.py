"""Developer helper (GPU): one c2 / c3 clip through the op (AINMF_COOP_DEBUG=1 prints the cooperative kernel's cycles per phase)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench, ainmf
for name in sys.argv[1:] or ["c2"]:
    wl = dict(bench.WORKLOADS[name])
    x = torch.from_numpy(bench.synth_host(wl, 0)[None]).cuda()
    f = lambda: ainmf.ops.nmf_inpaint(x, wl["n_fft"], wl["hop"], wl["K"], 200, 1e-4, wl["seed"], wl["thr"], wl["num"], wl["den"], -1, -1, 1, None, None)
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): out = f()
    torch.cuda.synchronize()
    print(name, "ms per call", (time.perf_counter() - t0) / 5 * 1e3, "n_iter", int(out[6][0]))
