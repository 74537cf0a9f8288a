#!/usr/bin/env python
"""bench.py -- audio-seconds restored per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c5|c1|c2|c3] [--legs ...]

Headline workload ("c4", BASELINE.json configs[3]): a batch of 10 s / 44.1 kHz clips with random-fragment masks
(generate_part1_data.create_random_mask semantics), n_fft 1024 / hop 256, K = 64, seed 42, 200 CD iterations,
tol 1e-4; 512 clips per GPU, so 8 GPUs process the named 4096 clips (weak scaling: clips are independent, no
data-path collective).  One step = the whole path (STFT -> frame mask -> imputation -> NMF fit -> recombine ->
iSTFT) over the batch.  `value` is measured with the batch resident in HBM; `e2e` goes through the C ABI's
host-buffer entry point (pinned host -> device -> host inside the timed region).

The same JSON line carries, as objects next to the headline keys (--legs, default all):
  "c5"            BASELINE configs[4] at EVERY N: the 1-hour signal, K = 128, 2048/512 -- one GPU at N = 1, time-frame-
                  sharded H + all-reduce of the W partial sums at N > 1 (strong scaling); with its own e2e, roofline, parity
                  against the oracle on a prefix and (N = 1) a same-box CPU figure;
  "c4_full_4096"  (N = 1) all 4096 clips of configs[3] on one GPU through ainmf_inpaint_host (chunked by free memory);
  "mu_kl"         (N = 1) the multiplicative-update / Kullback-Leibler solver (north_star (3)'s ratio form) on the c4 shape;
  "latency_cases" (N = 1) configs[0], [1], [2]: main4_NMF.py's 50 chained refits on the real segment, one 10 s clip with a
                  2 s gap at 2048/512, one clip with the random-fragment mask; each with the CPU path beside it.
`--impl reference` times the reference's own CPU path (scipy + sklearn through oracle/libcalls.py) for the same workloads.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
WORKLOADS = {
    # name: (samples, n_fft, hop, K, threshold, num, den, frac, seed, clips/GPU)
    "c4": dict(N=441000, n_fft=1024, hop=256, K=64, thr=0.01, num=4, den=5, frac=0.8, seed=42, clips=512,
               desc="BASELINE configs[3]: 10 s 44.1 kHz clips, random-fragment masks, K=64, 200 CD iterations"),
    "c2": dict(N=441000, n_fft=2048, hop=512, K=40, thr=1e-4, num=9, den=10, frac=0.9, seed=42, clips=1,
               desc="BASELINE configs[1]: one 10 s clip with a 2 s gap, n_fft 2048 / hop 512, K=40 (latency case)"),
    "c3": dict(N=441000, n_fft=1024, hop=256, K=40, thr=0.01, num=4, den=5, frac=0.8, seed=42, clips=1,
               desc="BASELINE configs[2]: one 10 s clip with the random-fragment mask of generate_part1_data, n_fft 1024 / hop 256, "
                    "K=40 (latency case)"),
}
GOLDEN = os.path.join(ROOT, "tests", "golden")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def mask_gaps(n_samples, clip_seed, mask_ratio=0.25, max_gap_len=400):
    """(starts, lens) of generate_part1_data.create_random_mask(n, 0.25) with np.random.seed(clip_seed)."""
    rs = np.random.RandomState(clip_seed)
    num = int(n_samples * mask_ratio / max_gap_len * 2)
    starts = np.empty(num, np.int64)
    lens = np.empty(num, np.int64)
    for i in range(num):
        lens[i] = rs.randint(50, max_gap_len)
        starts[i] = rs.randint(0, n_samples - lens[i])
    return starts, lens


def synth_host(wl, b):
    """Clip b of the synthetic workload on the host (float32, peak-normalised, masked) -- used for the CPU legs."""
    N = wl["N"]
    rs = np.random.RandomState(1000 + b)
    f = rs.uniform(100.0, 8000.0, 8)
    a = rs.uniform(0.05, 0.3, 8)
    ph = rs.uniform(0, 2 * np.pi, 8)
    t = np.arange(N, dtype=np.float64) / SR
    x = sum(a[j] * np.sin(2 * np.pi * f[j] * t + ph[j]) for j in range(8)) + 0.02 * rs.standard_normal(N)
    x = x.astype(np.float32)
    x = x / np.max(np.abs(x))
    if wl["thr"] > 1e-3:
        s, l = mask_gaps(N, b)
        for i in range(len(s)):
            x[s[i]:s[i] + l[i]] = 0
    else:
        c = N // 2
        x[c - SR:c + SR] = 0
    return x


def synth_device(wl, b0, B, device):
    """Same family of clips generated on the device (float32 sin instead of float64: a different but equally valid
    draw; parity is checked on clips produced by synth_host)."""
    import torch
    N = wl["N"]
    t = torch.arange(N, device=device, dtype=torch.float32) / SR
    x = torch.zeros((B, N), device=device, dtype=torch.float32)
    fr = np.empty((B, 8)); am = np.empty((B, 8)); ph = np.empty((B, 8))
    for i in range(B):
        rs = np.random.RandomState(1000 + b0 + i)
        fr[i], am[i], ph[i] = rs.uniform(100.0, 8000.0, 8), rs.uniform(0.05, 0.3, 8), rs.uniform(0, 2 * np.pi, 8)
    fr_d = torch.tensor(fr, device=device, dtype=torch.float32)
    am_d = torch.tensor(am, device=device, dtype=torch.float32)
    ph_d = torch.tensor(ph, device=device, dtype=torch.float32)
    for j in range(8):
        x += am_d[:, j:j + 1] * torch.sin(2 * np.pi * fr_d[:, j:j + 1] * t[None, :] + ph_d[:, j:j + 1])
    g = torch.Generator(device=device).manual_seed(1234 + b0)
    x += 0.02 * torch.randn((B, N), device=device, generator=g)
    x /= x.abs().amax(dim=1, keepdim=True)
    if wl["thr"] > 1e-3:
        import ainmf
        gl = [mask_gaps(N, b0 + i) for i in range(B)]
        starts = torch.from_numpy(np.stack([g[0] for g in gl])).to(device)
        lens = torch.from_numpy(np.stack([g[1] for g in gl])).to(device)
        ainmf.ops.apply_gaps_(x, starts, lens)              # the zeroing loop of generate_part1_data.py:44-46 on the device
    else:
        c = N // 2
        x[:, c - SR:c + SR] = 0
    return x.contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def alg_bytes_per_iter(F, T, K):
    """V read twice, W and Ht read twice + written once, two Grams (float32)."""
    return 8.0 * F * T + 12.0 * K * (F + T) + 8.0 * K * K


def cpu_leg(wl, n_clips, first_clip=0):
    """The reference's own calls (scipy.signal.stft -> mask loop -> mean-impute -> sklearn NMF cd -> istft) on the
    host cores -- oracle/libcalls.py, kind "port": the glue restated, the arithmetic run by the same wheels."""
    from oracle import libcalls
    t0 = time.perf_counter()
    its = 0
    outs = []
    for b in range(first_clip, first_clip + n_clips):
        x = synth_host(wl, b)
        y, st = libcalls.restore_columns(x, SR, n_fft=wl["n_fft"], hop=wl["hop"], threshold=wl["thr"], frac=wl["frac"],
                                         K=wl["K"], seed=wl["seed"], return_all=True)
        its += st.get("n_iter", 0)
        outs.append((x, y, st))
    dt = time.perf_counter() - t0
    return dt, its, outs


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs are meant to use every host core (OpenBLAS inside
    sklearn / numpy), so lift the limit at run time.  Returns the thread count in effect."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=cores)
        n = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
        return n
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", cores))


def newest_traffic(workload):
    """ncu dram bytes per launch (profiles/*_traffic.json written by profiles/summarize.py), newest capture for this
    workload; None when no capture of the current kernels has been committed."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for fn in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if fn.endswith("_traffic.json"):
            try:
                tj = json.load(open(os.path.join(pdir, fn)))
            except Exception:
                continue
            if tj.get("workload", "c4") == workload:
                tj["file"] = fn
                best = tj
    return best


C5 = dict(N=158760000, n_fft=2048, hop=512, K=128, thr=1e-4, num=9, den=10, frac=0.9, seed=0,
          desc="BASELINE configs[4]: one 1-hour 44.1 kHz signal, 2 s gaps every 30 s, n_fft 2048 / hop 512, K=128, "
               "200 CD iterations; N>1: time-frame-sharded H + all-reduce of the W partial sums (strong scaling)")


def c5_signal_host(N):
    """0.1*N(0,1) + a bed of 8 sinusoids, zeroed on [s, s+2) s for s = 15, 45, ..., peak-normalised."""
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal(N)).astype(np.float32)
    rs = np.random.RandomState(7)
    fr, am = rs.uniform(100.0, 8000.0, 8), rs.uniform(0.05, 0.3, 8)
    chunk = 1 << 22
    for a in range(0, N, chunk):
        b = min(N, a + chunk)
        t = np.arange(a, b, dtype=np.float64) / SR
        for j in range(8):
            x[a:b] += (am[j] * np.sin(2 * np.pi * fr[j] * t)).astype(np.float32)
    s = 15
    while (s + 2) * SR <= N:
        x[s * SR:(s + 2) * SR] = 0
        s += 30
    return x / np.abs(x).max()


def c5_signal_device(device, N):
    """The same family generated on the device (the 1-hour signal is 635 MB; parity is checked on a host-generated prefix)."""
    import torch
    g = torch.Generator(device=device).manual_seed(0)
    x = 0.1 * torch.randn(N, device=device, generator=g)
    rs = np.random.RandomState(7)
    fr, am = rs.uniform(100.0, 8000.0, 8), rs.uniform(0.05, 0.3, 8)
    chunk = 1 << 24
    for a in range(0, N, chunk):
        b = min(N, a + chunk)
        t = torch.arange(a, b, device=device, dtype=torch.float64) / SR
        for j in range(8):
            x[a:b] += (am[j] * torch.sin(2 * np.pi * fr[j] * t)).float()
    s = 15
    while (s + 2) * SR <= N:
        x[s * SR:(s + 2) * SR] = 0
        s += 30
    return x / x.abs().max()


def c5_cpu_sample(seconds=240.0, iters=6):
    """BASELINE.md section 3 for configs[4], bounded: the reference's calls on a `seconds`-long prefix of the signal with
    `iters` CD iterations; the once-per-signal stages and the per-iteration cost both scale linearly with the length, so
    the hour costs (outside + per_iter * 200) * 3600 / seconds.  Returns (audio-s/s for the full config, detail dict)."""
    from oracle import libcalls
    from scipy import signal
    N = int(seconds * SR)
    x = c5_signal_host(N)
    t0 = time.perf_counter()
    Z, mag, phase, _ = libcalls.stft_mag_phase(x, SR, C5["n_fft"], C5["hop"])
    bad = libcalls.column_mask(x, mag.shape[1], C5["hop"], C5["thr"], C5["frac"])
    cur = libcalls.impute(mag, bad)
    t1 = time.perf_counter()
    W, H, n_iter, err = libcalls.nmf_fit(cur, C5["K"], C5["seed"], iters, 1e-4)
    t2 = time.perf_counter()
    final = mag.copy()
    final[:, bad] = (W @ H)[:, bad]
    Zr = final * np.exp(1j * phase)
    _, y = signal.istft(Zr, SR, nperseg=C5["n_fft"], noverlap=C5["n_fft"] - C5["hop"])
    t3 = time.perf_counter()
    outside, per_iter = (t1 - t0) + (t3 - t2), (t2 - t1) / max(n_iter, 1)
    full_s = (outside + per_iter * 200) * (C5["N"] / SR) / seconds
    return (C5["N"] / SR) / full_s, dict(prefix_seconds=seconds, iterations_timed=n_iter, outside_loop_s=outside,
                                         s_per_iteration=per_iter, extrapolated_seconds_for_the_hour=full_s,
                                         cpu_seconds_spent=t3 - t0)


class Ctx:
    """Process-wide state of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: ainmf has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        import ainmf
        from ainmf import _capi
        self.ainmf, self.capi, self.ops = ainmf, _capi, ainmf.ops
        self.L = ainmf._lib.lib()
        self.h = ainmf._lib.handle(self.local_rank)
        self.args = args

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], device=self.device, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def free(self):
        self.ops._workspaces.clear()
        self.torch.cuda.empty_cache()

    def profile_iteration(self, step):
        """Live CUDA-event time of the iteration's kernels during one extra step (outside the timed legs)."""
        self.L.ainmf_profile(self.h, 1, None, None)
        step()
        self.torch.cuda.synchronize()
        pms, pcn = (C.c_double * 6)(), (C.c_int64 * 6)()
        self.L.ainmf_profile(self.h, 0, pms, pcn)
        names = ["gram_Ht", "xht_gram_tc", "w_side_fused", "gram_W", "h_step_tc", "stop_rule"]
        return {k: float(pms[i]) for i, k in enumerate(names)}, max(int(pcn[4]), 1)


KERNEL_NAMES = {"h_step_tc": "h_step_ts_kernel (X^T.W contraction on tcgen05 + H coordinate sweep, one launch per iteration)",
                "xht_gram_tc": "xht_ts_kernel (X.Ht contraction + Gram of Ht on tcgen05, one launch per iteration)",
                "w_side_fused": "w_side_kernel (W sweep, W^T W and the H step's operands)"}
NCU_NAMES = {"h_step_tc": "h_step_ts_kernel", "xht_gram_tc": "xht_ts_kernel", "w_side_fused": "w_side_kernel"}


def roofline_object(kern_ms, n_it, F, T, K, units, workload, note_iteration, traffic_units=None):
    """The contract's roofline object for the dominant kernel + the same accounting per kernel and for the whole iteration.
    units = spectrograms one launch processes (clips; 1 for the long signal, whose T is this rank's slice)."""
    peak, peak_src = peaks()
    kb = {"gram_Ht": 4.0 * T * K + 4.0 * K * K, "xht_gram_tc": 4.0 * F * T + 4.0 * T * K + 4.0 * F * K,
          "w_side_fused": 12.0 * F * K + 8.0 * K * K, "gram_W": 4.0 * F * K + 4.0 * K * K,
          "h_step_tc": 4.0 * F * T + 4.0 * F * K + 8.0 * T * K + 4.0 * K * K, "stop_rule": 0.0}
    tot = max(sum(kern_ms.values()), 1e-12)
    kernels = {k: {"ms_per_launch": v / n_it, "share": v / tot,
                   "hbm_frac": (kb[k] * units / (v / n_it * 1e-3) / 1e9 / peak) if v > 0 else None} for k, v in kern_ms.items()}
    dom = max(kern_ms, key=lambda k: kern_ms[k])
    dom_ms = kern_ms[dom] / n_it
    dom_ach = kb[dom] * units / (dom_ms * 1e-3) / 1e9
    iter_ms = tot / n_it
    bytes_iter = alg_bytes_per_iter(F, T, K) * units
    ach = bytes_iter / (iter_ms * 1e-3) / 1e9
    traffic, tfile = None, None
    tj = newest_traffic(workload)
    if tj and NCU_NAMES.get(dom) in tj.get("dram_bytes_per_launch", {}):
        n_units = tj.get("units") or tj.get("clips")       # clips (c4) or frames (c5) of the captured launch
        now = traffic_units if traffic_units is not None else units
        traffic = tj["dram_bytes_per_launch"][NCU_NAMES[dom]] / n_units * now if n_units else None
        tfile = tj["file"]
    return {"bound": "hbm", "kernel": KERNEL_NAMES.get(dom, dom), "achieved": dom_ach, "peak": peak, "unit": "GB/s",
            "frac": dom_ach / peak, "traffic": traffic, "traffic_source": tfile, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": kb[dom] * units, "ms_per_launch": dom_ms,
            "iteration": {"kernel": note_iteration, "achieved": ach, "frac": ach / peak,
                          "algorithmic_bytes_per_iteration": bytes_iter, "ms_per_iteration": iter_ms},
            "kernels": kernels}


def coop_roofline_object(fit_ms, n_iter, F, T, K):
    """One clip through nmf_coop_kernel: the launch is the whole fit; per iteration it reads X twice from L2 (a clip's
    spectrogram fits the 126 MB L2 and stays there by design), so the HBM fraction is a latency-case figure, not a bandwidth
    claim; the FP32 rate (4 F T K flops per iteration in the two products) is what the kernel is made of."""
    peak, peak_src = peaks()
    iter_ms = fit_ms / n_iter
    bytes_iter = alg_bytes_per_iter(F, T, K)
    ach = bytes_iter / (iter_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "nmf_coop_kernel (the whole fit of one clip as one cooperative launch: 148 CTAs, 4 grid barriers per iteration, FFMA)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_iter * n_iter, "ms_per_launch": fit_ms,
            "iteration": {"kernel": "one iteration inside the cooperative launch (X.Ht | Gram, W sweep, Xt.W | Gram, H sweep, stop rule)",
                          "achieved": ach, "frac": ach / peak, "algorithmic_bytes_per_iteration": bytes_iter, "ms_per_iteration": iter_ms,
                          "fp32_tflops": 4.0 * F * T * K * 2 / (iter_ms * 1e-3) / 1e12},
            "kernels": {"nmf_coop_kernel": {"ms_per_launch": fit_ms, "share": 1.0, "hbm_frac": ach / peak}}}


def parity_vs_oracle(ctx, x_host, wl, K, max_iter=200):
    """One host signal through the op and through the oracle: the gates of north_star as numbers."""
    from oracle import libcalls
    torch = ctx.torch
    yo, st = libcalls.restore_columns(x_host, SR, n_fft=wl["n_fft"], hop=wl["hop"], threshold=wl["thr"], frac=wl["frac"], K=K,
                                      seed=wl["seed"], max_iter=max_iter, return_all=True)
    yg, ig, ng, _, _, eg, itg = ctx.ops.nmf_inpaint(torch.from_numpy(x_host[None]).to(ctx.device), wl["n_fft"], wl["hop"], K, max_iter,
                                                    1e-4, wl["seed"], wl["thr"], wl["num"], wl["den"], -1, -1, 1, None, None)
    bad, n = st["bad"], int(ng[0])
    y = yg[0].cpu().numpy()
    m = np.zeros(len(x_host), bool)
    for c in bad:
        m[max(0, c * wl["hop"] - wl["n_fft"] // 2):min(len(x_host), c * wl["hop"] + wl["n_fft"] // 2)] = True
    return {"mask_bit_exact": bool(n == len(bad) and np.array_equal(ig[0, :n].cpu().numpy(), bad)),
            "objective_rel_diff": abs(float(eg[0]) - st["err"]) / st["err"],
            "snr_vs_oracle_db": float(libcalls.snr_db(yo, y)), "snr_restored_samples_db": float(libcalls.snr_db(yo[m], y[m])),
            "n_iter": [int(itg[0]), st["n_iter"]], "max_iter": max_iter}


# =====================================================================================================
# clip batches (configs[3]; also the single-clip cases configs[1], configs[2])
# =====================================================================================================
def leg_clips(ctx, name, wl, steps, warmup, cpu_baseline=True, headline=False):
    torch, L, h, ops = ctx.torch, ctx.L, ctx.h, ctx.ops
    B, N, K = wl["clips"], wl["N"], wl["K"]
    T, F, _ = ctx.capi.stft_geometry(L, N, wl["n_fft"], wl["hop"])
    if B == 1:
        x = torch.from_numpy(synth_host(wl, 0)[None]).to(ctx.device)        # the latency cases run the host-generated clip
    else:
        x = synth_device(wl, ctx.rank * B, B, ctx.device)
    torch.cuda.synchronize()

    def step():
        return ops.nmf_inpaint(x, wl["n_fft"], wl["hop"], K, 200, 1e-4, wl["seed"], wl["thr"], wl["num"], wl["den"],
                               -1, -1, 1, None, None)

    t0 = time.perf_counter()
    out = step()
    torch.cuda.synchronize()
    cold_ms = (time.perf_counter() - t0) * 1e3          # first call: seeded N(0,1) tables (host RNG), FFT tables, workspace
    for _ in range(max(warmup - 1, 0)):
        out = step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0 and headline:
        sampler.start()
    launches0 = L.ainmf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    ctx.barrier()
    launches = L.ainmf_launch_count() - launches0
    clocks = sampler.stop() if (ctx.rank == 0 and headline) else None
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    y, idx, nb, Wf, Hf, err, nit = out
    iters_done = int(nit.sum())
    kern_ms, n_it = ctx.profile_iteration(step)
    if n_it == 1 and kern_ms["xht_gram_tc"] == 0.0 and kern_ms["gram_Ht"] == 0.0 and kern_ms["h_step_tc"] > 0.0:
        # one clip: the whole fit is ONE cooperative launch (nmf_coop.cu), timed as a whole under the H-step slot
        roofline = coop_roofline_object(kern_ms["h_step_tc"], max(int(nit.max()), 1), F, T, K)
    else:
        roofline = roofline_object(kern_ms, n_it, F, T, K, B, name,
                                   "whole CD iteration = xht_ts_kernel + reduce_splits + w_side_kernel + w_finish_kernel + h_step_ts_kernel + stop_kernel")

    # ---- end-to-end leg through the C ABI with HOST buffers -----------------------------------------
    xh = torch.empty((B, N), dtype=torch.float32).pin_memory()
    xh.copy_(x)
    yh = torch.empty((B, N), dtype=torch.float32).pin_memory()
    nbh = np.zeros(B, np.int32); errh = np.zeros(B, np.float32); nih = np.zeros(B, np.int32)
    p = ctx.capi.default_params(L, batch=B, n_samples=N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4,
                                seed=wl["seed"], threshold=wl["thr"], frac_num=wl["num"], frac_den=wl["den"])
    del out, y, idx, Wf, Hf
    ctx.free()

    def e2e_step():
        rc = L.ainmf_inpaint_host(h, C.byref(p), C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()),
                                  nbh.ctypes.data_as(C.c_void_p), errh.ctypes.data_as(C.c_void_p),
                                  nih.ctypes.data_as(C.c_void_p), 0)
        ctx.ainmf._lib.check(rc, ctx.local_rank)

    for _ in range(warmup):
        e2e_step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
    ctx.barrier()
    # ---- the same from 16-bit file samples to 16-bit file samples (ainmf_inpaint_host_pcm16): half the bytes over PCIe ----
    e2e_pcm = None
    if headline and B > 1:
        ph = torch.empty((B, N), dtype=torch.int16).pin_memory()
        ph.copy_((xh * 32767.0).to(torch.int16))
        qh = torch.empty((B, N), dtype=torch.int16).pin_memory()
        pkh = np.zeros(B, np.float32)

        def pcm_step():
            rc = L.ainmf_inpaint_host_pcm16(h, C.byref(p), C.c_void_p(ph.data_ptr()), 1, C.c_void_p(qh.data_ptr()),
                                            pkh.ctypes.data_as(C.c_void_p), nbh.ctypes.data_as(C.c_void_p),
                                            errh.ctypes.data_as(C.c_void_p), nih.ctypes.data_as(C.c_void_p), 0)
            ctx.ainmf._lib.check(rc, ctx.local_rank)

        for _ in range(warmup):
            pcm_step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            pcm_step()
        torch.cuda.synchronize()
        pcm_ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
        ctx.barrier()
        e2e_pcm = {"value": ctx.world * B * N / SR * steps / (pcm_ms * 1e-3), "unit": "audio-s/s", "ms_per_step": pcm_ms / steps,
                   "h2d_bytes_per_step": B * N * 2, "d2h_bytes_per_step": B * N * 2 + B * 16,
                   "api": "ainmf_inpaint_host_pcm16 (int16 samples in, int16 samples out; load_damaged_data and save_result on the device)",
                   "iterations_done": int(nih.sum())}
        del ph, qh
    audio_s_step = ctx.world * B * N / SR
    line = {
        "metric": "audio_seconds_restored_per_second", "value": audio_s_step * steps / (ms_total * 1e-3), "unit": "audio-s/s",
        "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {wl['desc']}", "clips_per_gpu": B, "n_samples": N, "n_fft": wl["n_fft"],
                   "hop": wl["hop"], "F": F, "T": T, "rank": K, "max_iter": 200, "tol": 1e-4, "solver": "cd",
                   "l2": ("inputs (%.0f MB/GPU of waveform, %.1f GB of spectrogram) exceed the 126 MB L2" % (B * N * 4 / 1e6, B * F * T * 4 / 1e9))
                         if B > 64 else "latency case: one clip (%.1f MB of spectrogram) fits the L2 and stays there across the 200 iterations by design; "
                                        "not a bandwidth figure" % (F * T * 4 / 1e6)},
        "nmf_iters_per_s": ctx.world * B * 200 / (sum(kern_ms.values()) * 1e-3) if sum(kern_ms.values()) > 0 else None,
        "nmf_iterations_done_last_step": iters_done,
        "cold_first_call_ms": cold_ms,
        "e2e": {"value": audio_s_step * steps / (e2e_ms * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": B * N * 4,
                "d2h_bytes_per_step": B * N * 4 + B * 12, "ms_per_step": e2e_ms / steps,
                "api": "ainmf_inpaint_host (C ABI, pinned host buffers)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": None, "parity": None,
    }
    if e2e_pcm:
        line["e2e_pcm16"] = e2e_pcm
    if ctx.rank == 0 and ctx.world == 1 and cpu_baseline:
        n_cpu = 8 if B > 1 else 1
        cores = use_all_host_threads()
        dt, its, outs = cpu_leg(wl, n_cpu)
        line["cpu_baseline"] = {"value": n_cpu * N / SR / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} clip(s) of the workload, serially, through scipy.signal.stft/istft + sklearn NMF(cd) "
                                          f"(oracle/libcalls.py); OpenBLAS threads = {cores}, coordinate sweep single-threaded",
                                "nmf_iters_per_s": its / dt, "seconds": dt}
        line["parity"] = [parity_vs_oracle(ctx, o[0], wl, K) for o in outs[:2]]
    del xh, yh
    ctx.free()
    return line


def leg_c4_full(ctx, steps=1):
    """All 4096 clips of configs[3] on ONE GPU: host buffers, chunked by the library (strong-scaling point of the c4 curve)."""
    torch, L, h = ctx.torch, ctx.L, ctx.h
    wl = dict(WORKLOADS["c4"])
    Btot, N, K = 4096, wl["N"], wl["K"]
    try:
        xh = torch.empty((Btot, N), dtype=torch.float32).pin_memory()
        yh = torch.empty((Btot, N), dtype=torch.float32).pin_memory()
    except Exception as e:                                   # host without 14.5 GB of pinnable memory
        return {"unavailable": f"cannot pin 2 x {Btot * N * 4 / 1e9:.1f} GB of host memory: {e}"}
    for b0 in range(0, Btot, 512):
        xh[b0:b0 + 512].copy_(synth_device(wl, b0, 512, ctx.device))
    torch.cuda.synchronize()
    ctx.free()
    nbh = np.zeros(Btot, np.int32); errh = np.zeros(Btot, np.float32); nih = np.zeros(Btot, np.int32)
    p = ctx.capi.default_params(L, batch=Btot, n_samples=N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4,
                                seed=wl["seed"], threshold=wl["thr"], frac_num=wl["num"], frac_den=wl["den"])

    def call():
        rc = L.ainmf_inpaint_host(h, C.byref(p), C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()), nbh.ctypes.data_as(C.c_void_p),
                                  errh.ctypes.data_as(C.c_void_p), nih.ctypes.data_as(C.c_void_p), 0)
        ctx.ainmf._lib.check(rc, ctx.local_rank)

    call()                                                   # warm (workspace, pinned staging)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    out = {"clips": Btot, "api": "ainmf_inpaint_host (host buffers, chunked by free device memory)", "seconds_per_call": dt,
           "value": Btot * N / SR / dt, "unit": "audio-s/s", "h2d_bytes": Btot * N * 4, "d2h_bytes": Btot * N * 4,
           "nmf_iterations_done": int(nih.sum()), "clips_with_bad_frames": int((nbh > 0).sum())}
    del xh, yh
    ctx.free()
    return out


# =====================================================================================================
# configs[0]: main4_NMF.py on the real 50 ms segment (the shipped golden fixture): 50 chained refits with early stop
# =====================================================================================================
def leg_c1(ctx, steps=5, cpu_baseline=True):
    torch = ctx.torch
    g = np.load(os.path.join(GOLDEN, "c1_part0.npz"))
    ainmf = ctx.ainmf
    lab = ainmf.SpectralInpainter.__new__(ainmf.SpectralInpainter)
    ainmf.SpectralInpainter.__init__(lab, filename=None, duration=0.05, device=str(ctx.device))
    lab.sr = int(g["sr"])
    lab.raw_audio = g["raw"].copy()
    lab.apply_mask(0.2)
    launches0 = ctx.L.ainmf_launch_count()
    lab.restore_with_nmf(n_components=40, n_iter=50)
    launches = ctx.L.ainmf_launch_count() - launches0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = lab.restore_with_nmf(n_components=40, n_iter=50)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    dur = len(lab.raw_audio) / lab.sr
    res = {"workload": "c1: BASELINE configs[0]: main4_NMF.py on the 50 ms real segment (257 x 19, K=40, 50 refits with early stop), "
                       "through ainmf.SpectralInpainter.restore_with_nmf (host arrays in and out)",
           "seconds_per_call": dt, "value": dur / dt, "unit": "audio-s/s", "n_iter_last_refit": lab.n_iter_,
           "launches_per_call": int(launches), "data": "tests/golden/c1_part0.npz (the reference's own segment)"}
    if cpu_baseline:
        from oracle import libcalls
        cores = use_all_host_threads()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            yo, st = libcalls.part0_restore(g["raw"], lab.corrupted_audio, lab.sr, lab.gap_start, lab.gap_end, return_all=True)
            best = min(best, time.perf_counter() - t0)
        res["cpu_baseline"] = {"value": dur / best, "unit": "audio-s/s", "cores": cores, "kind": "port", "seconds": best,
                               "sample": "the whole config (best of 3) through oracle/libcalls.part0_restore"}
        res["parity"] = {"snr_vs_oracle_db": float(libcalls.snr_db(yo, out)), "n_iter_last_refit": [lab.n_iter_, st["n_iters"][-1]],
                         "objective_rel_diff": abs(lab.reconstruction_err_ - st["err"]) / st["err"]}
        res["speedup_vs_cpu"] = best / dt
    return res


# =====================================================================================================
# configs[4]: the 1-hour signal
# =====================================================================================================
def leg_c5(ctx, steps, warmup, seconds=0.0, cpu_baseline=True, clocks=False):
    """N=1: torch.ops.ainmf.nmf_inpaint; N>1: TimeShardedInpainter (time-frame split + all-reduce of the W partial sums)."""
    torch, L, h = ctx.torch, ctx.L, ctx.h
    wl = dict(C5)
    if seconds > 0:
        wl["N"] = int(seconds * SR)
    N, K, world = wl["N"], wl["K"], ctx.world
    T, F, _ = ctx.capi.stft_geometry(L, N, wl["n_fft"], wl["hop"])
    x = c5_signal_device(ctx.device, N)
    if world > 1:
        from ainmf.sharding import TimeShardedInpainter
        tsi = TimeShardedInpainter(ctx.device)
        pl = tsi.plan(N, wl["n_fft"], wl["hop"])
        xl = x[pl["x_begin"]:pl["x_end"]].clone()
        del x
        torch.cuda.empty_cache()

        def step():
            return tsi.restore(xl, N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4, seed=wl["seed"],
                               threshold=wl["thr"], frac=(wl["num"], wl["den"]))
        src = xl
    else:
        xb = x[None]

        def step():
            out = ctx.ops.nmf_inpaint(xb, wl["n_fft"], wl["hop"], K, 200, 1e-4, wl["seed"], wl["thr"], wl["num"], wl["den"],
                                      -1, -1, 1, None, None)
            return out[0][0], dict(n_bad=out[2], n_iter=out[6], err=out[5])
        src = x

    t0 = time.perf_counter()
    y, info = step()
    torch.cuda.synchronize()
    cold_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(max(warmup - 1, 0)):
        y, info = step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0 and clocks:
        sampler.start()
    launches0 = L.ainmf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y, info = step()
    e1.record()
    ctx.barrier()
    launches = L.ainmf_launch_count() - launches0
    clk = sampler.stop() if (ctx.rank == 0 and clocks) else None
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    kern_ms, n_it = ctx.profile_iteration(step)
    kernels_per_step = ctx.max_over_ranks(sum(kern_ms.values()))
    # end to end: pinned host -> device -> host around the same call
    xh = torch.empty(src.shape, dtype=torch.float32).pin_memory()
    xh.copy_(src)
    yh = torch.empty(y.shape, dtype=torch.float32).pin_memory()

    def e2e_step():
        src.copy_(xh, non_blocking=True)
        yy, _ = step()
        yh.copy_(yy, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(warmup):
        e2e_step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    e2e_ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
    ctx.barrier()
    Tl = T // world
    audio_s = N / SR
    step_ms = ms / steps
    res = {
        "metric": "audio_seconds_restored_per_second", "value": audio_s * steps / (ms * 1e-3), "unit": "audio-s/s",
        "n_gpus": world, "n_ranks": world, "steps": steps, "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c5: " + wl["desc"], "n_samples": N, "n_fft": wl["n_fft"], "hop": wl["hop"], "F": F, "T": T,
                   "frames_per_rank": Tl, "rank": K, "max_iter": 200, "tol": 1e-4, "solver": "cd",
                   "l2": "the spectrogram slice per GPU (%.2f GB) exceeds the 126 MB L2" % (F * Tl * 4 / 1e9)},
        "nmf_iters_per_s": 200 * steps / (ms * 1e-3),
        "n_iter": int(info["n_iter"][0]), "n_bad": int(info["n_bad"][0]), "objective": float(info["err"][0]),
        "cold_first_call_ms": cold_ms,
        "ms_per_iteration_kernels": kernels_per_step / n_it,
        "ms_per_step_outside_iteration_kernels": step_ms - kernels_per_step,
        "exposed_note": "ms_per_step minus the CUDA-event time of the iteration's kernels (max over ranks): STFT, mask, imputation, "
                        "initial factors, objective, iSTFT, launch gaps and, at N > 1, the all-reduces between the kernels",
        "e2e": {"value": audio_s * steps / (e2e_ms * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": int(src.numel() * 4),
                "d2h_bytes_per_step": int(y.numel() * 4), "ms_per_step": e2e_ms / steps},
        "gpu_launches": int(launches), "clocks": clk,
        "transport": {0: "single rank", 1: "ncclAllReduce between the kernels of an iteration",
                      2: "peer mailboxes over NVLink (CUDA IPC): push kernel + ordered slot sum, no library collective "
                         "inside the iteration"}[int(L.ainmf_comm_transport(h))],
        "roofline": roofline_object(kern_ms, n_it, F, Tl, K, 1, "c5",
                                    "whole CD iteration on this rank (its frame slice; the W side is replicated, not sharded)", traffic_units=Tl),
        "cpu_baseline": None, "parity": None,
    }
    del xh, yh, y, src
    if world == 1:
        del xb, x
    ctx.free()
    if ctx.rank == 0 and world == 1 and cpu_baseline:
        cores = use_all_host_threads()
        val, det = c5_cpu_sample()
        res["cpu_baseline"] = {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                               "sample": "a %.0f s prefix of the signal: STFT, mask, imputation, recombination and iSTFT once + %d CD "
                                         "iterations through scipy + sklearn (oracle/libcalls.py), extrapolated linearly to the hour at 200 "
                                         "iterations (BASELINE.md section 3); OpenBLAS threads = %d" % (det["prefix_seconds"], det["iterations_timed"], cores),
                               **det}
        res["parity"] = parity_vs_oracle(ctx, c5_signal_host(int(60 * SR)), wl, K, max_iter=30)
        res["parity"]["note"] = "60 s prefix, 30 iterations (a full-length oracle run takes ~20 min of CPU)"
        ctx.free()
    return res


# =====================================================================================================
# reference arm: the reference's own CPU path (scipy + sklearn via oracle/libcalls.py), rank 0 only
# =====================================================================================================
def run_reference(args, rank):
    if rank != 0:
        return
    cores = use_all_host_threads()
    name = args.workload
    base = {"impl": "reference", "metric": "audio_seconds_restored_per_second", "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    if name == "c5":
        val, det, t = 0.0, None, 0.0
        for _ in range(max(args.steps, 1)):
            v, det = c5_cpu_sample(seconds=120.0, iters=6)
            val += v
            t += det["cpu_seconds_spent"]
        val /= max(args.steps, 1)
        line = dict(base, value=val, ms_per_step=1e3 * t / max(args.steps, 1), scaling="strong",
                    config={"workload": "c5: " + C5["desc"], "n_fft": C5["n_fft"], "hop": C5["hop"], "rank": C5["K"], "max_iter": 200, "tol": 1e-4},
                    cpu_baseline={"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                  "sample": "per step: a 120 s prefix, every stage once + 6 CD iterations, extrapolated to the hour at 200 iterations", **det},
                    e2e={"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    if name == "c1":
        class _NoGpu:
            pass
        from oracle import libcalls
        g = np.load(os.path.join(GOLDEN, "c1_part0.npz"))
        cor, gs, ge = libcalls.part0_apply_mask(g["raw"])
        t = 0.0
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            libcalls.part0_restore(g["raw"], cor, int(g["sr"]), gs, ge)
            if i >= args.warmup:
                t += time.perf_counter() - t0
        dur = len(g["raw"]) / int(g["sr"])
        val = dur * args.steps / t
        line = dict(base, value=val, ms_per_step=1e3 * t / args.steps, scaling="weak", data="tests/golden/c1_part0.npz",
                    config={"workload": "c1: BASELINE configs[0]: main4_NMF.py on the 50 ms real segment"},
                    cpu_baseline={"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": "the whole config per step"},
                    e2e={"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    wl = dict(WORKLOADS[name])
    clips_per_step = 2 if wl["clips"] > 1 else 1
    for _ in range(args.warmup):
        cpu_leg(wl, 1)
    t, its = 0.0, 0
    for s in range(args.steps):
        dt, it, _ = cpu_leg(wl, clips_per_step, first_clip=s * clips_per_step)
        t += dt
        its += it
    audio_s = args.steps * clips_per_step * wl["N"] / SR
    val = audio_s / t
    line = dict(base, value=val, ms_per_step=1e3 * t / args.steps, scaling="weak",
                config={"workload": f"{name}: {wl['desc']}", "n_fft": wl["n_fft"], "hop": wl["hop"], "rank": wl["K"],
                        "max_iter": 200, "tol": 1e-4, "clips_per_step": clips_per_step},
                nmf_iters_per_s=its / t,
                cpu_baseline={"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                              "sample": f"{clips_per_step} clip(s) of the workload per step through scipy.signal.stft/istft + "
                                        f"sklearn NMF(cd) (oracle/libcalls.py), OpenBLAS on {cores} threads, sweep single-threaded"},
                e2e={"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    if name == "c4" and "c5" in args.legs:
        v5, det = c5_cpu_sample(seconds=120.0, iters=6)
        line["c5"] = {"impl": "reference", "value": v5, "unit": "audio-s/s", "cores": cores, "kind": "port",
                      "sample": "a 120 s prefix, every stage once + 6 CD iterations, extrapolated to the hour at 200 iterations", **det}
    print(json.dumps(line), flush=True)


def leg_mu_kl(ctx, cpu):
    """north_star (3)'s multiplicative-update form (solver 'mu-kl': W.H, X / (W.H) and both contractions fused per half-step,
    nmf_mukl.cu) on spectrograms of the c4 shape: ms per iteration over a batch, as a fraction of the HBM roofline (an
    iteration reads X twice: 8 F T bytes + factors) and of the FP32 rate it is actually bound by (8 F T K flops);
    sklearn's solver='mu', beta_loss='kullback-leibler' on one spectrogram beside it, with the parity of that one."""
    torch = ctx.torch
    F, T, K, B, iters = 513, 1724, 64, 256, 10
    g = torch.Generator(device=ctx.device).manual_seed(1)
    X = torch.rand((B, F, T), device=ctx.device, generator=g) ** 2
    ctx.ops.nmf_fit(X, K, 2, 0.0, 42, None, None, "mu-kl")
    ctx.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); ctx.ops.nmf_fit(X, K, 2, 0.0, 42, None, None, "mu-kl")
    e1.record(); W, H, err, nit = ctx.ops.nmf_fit(X, K, 2 + iters, 0.0, 42, None, None, "mu-kl")
    e2.record(); torch.cuda.synchronize()
    ms_it = (e1.elapsed_time(e2) - e0.elapsed_time(e1)) / iters          # the fixed costs (transpose, init, error) cancel
    bytes_it = (8.0 * F * T + 12.0 * K * (F + T)) * B
    flops_it = 8.0 * F * T * K * B
    peak, src = peaks()
    out = {"workload": f"{B} spectrograms {F} x {T}, K = {K}: multiplicative update, Kullback-Leibler divergence (ratio form), fused FFMA kernels",
           "ms_per_iteration": ms_it, "nmf_iters_per_s": 1e3 * B / ms_it, "hbm_frac": bytes_it / (ms_it * 1e-3) / 1e9 / peak,
           "fp32_tflops": flops_it / (ms_it * 1e-3) / 1e12, "bound": "fp32 FMA pipe (8 F T K flops against 8 F T bytes: 64 flop/B at K = 64)"}
    if cpu:
        from oracle import libcalls
        cores = use_all_host_threads()
        x1 = X[0].cpu().numpy()
        t0 = time.perf_counter()
        Wo, Ho, no, eo = libcalls.nmf_fit(x1, K, seed=42, max_iter=12, tol=0.0, solver="mu", beta_loss="kullback-leibler")
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"nmf_iters_per_s": 12 / dt, "cores": cores, "kind": "port",
                               "sample": "12 iterations of sklearn NMF(solver='mu', beta_loss='kullback-leibler') on one spectrogram"}
        out["parity"] = {"objective_rel_diff": abs(float(err[0]) - eo) / eo, "n_iter": [int(nit[0]), no]}
        out["speedup_vs_cpu"] = out["nmf_iters_per_s"] / out["cpu_baseline"]["nmf_iters_per_s"]
    del X
    ctx.free()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=0.0, help="c5 only: signal length override")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU (default: the workload's)")
    ap.add_argument("--legs", default="c5,c4_full,latency,mu_kl", help="extra objects of the default (c4) line: comma list of c5, c4_full, "
                                                                 "latency, mu_kl; 'none' for the headline only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.legs = [] if args.legs == "none" else [s for s in args.legs.split(",") if s]
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    ctx = Ctx(args)
    cpu = not args.no_cpu_baseline
    if args.workload == "c5":
        line = leg_c5(ctx, args.steps, args.warmup, args.seconds, cpu, clocks=True)
    elif args.workload == "c1":
        r = leg_c1(ctx, max(args.steps, 1), cpu)
        line = {"metric": "audio_seconds_restored_per_second", "value": r["value"], "unit": "audio-s/s", "n_gpus": 1, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds_per_call"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": r["data"], "config": {"workload": r["workload"]},
                "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 2205 * 4, "d2h_bytes_per_step": 2205 * 4},
                "gpu_launches": r["launches_per_call"] * args.steps, "cpu_baseline": r.get("cpu_baseline"), "parity": r.get("parity")}
    else:
        wl = dict(WORKLOADS[args.workload])
        if args.clips > 0:
            wl["clips"] = args.clips
        line = leg_clips(ctx, args.workload, wl, args.steps, args.warmup, cpu, headline=True)
        if args.workload == "c4":
            if "c5" in args.legs:
                line["c5"] = leg_c5(ctx, min(args.steps, 3), 3, args.seconds, cpu)
            if ctx.world == 1 and "c4_full" in args.legs:
                line["c4_full_4096"] = leg_c4_full(ctx)
            if ctx.world == 1 and "latency" in args.legs:
                lat = {"c1": leg_c1(ctx, 3, cpu)}
                for nm in ("c2", "c3"):
                    r = leg_clips(ctx, nm, dict(WORKLOADS[nm]), 5, 3, cpu)
                    lat[nm] = {k: r[k] for k in ("value", "unit", "ms_per_step", "config", "cold_first_call_ms", "e2e", "cpu_baseline", "parity", "nmf_iters_per_s")}
                    lat[nm]["ms_per_iteration"] = r["roofline"]["iteration"]["ms_per_iteration"]
                line["latency_cases"] = lat
            if ctx.world == 1 and "mu_kl" in args.legs:
                line["mu_kl"] = leg_mu_kl(ctx, cpu)
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
