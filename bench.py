#!/usr/bin/env python
"""bench.py -- audio-seconds restored per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c2|c5]

Default workload ("c4", BASELINE.json configs[3]): a batch of 10 s / 44.1 kHz clips with random-fragment masks
(generate_part1_data.create_random_mask semantics), n_fft 1024 / hop 256, K = 64, seed 42, 200 CD iterations,
tol 1e-4; 512 clips per GPU, so 8 GPUs process the named 4096 clips (weak scaling: clips are independent, no
data-path collective).  One step = the whole path (STFT -> frame mask -> imputation -> NMF fit -> recombine ->
iSTFT) over the batch.  `value` is measured with the batch resident in HBM; `e2e` goes through the C ABI's
host-buffer entry point (pinned host -> device -> host inside the timed region).
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
}


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


def run_reference(args, wl, rank):
    if rank != 0:
        return
    cores = use_all_host_threads()
    clips_per_step = 2 if wl["clips"] > 1 else 1
    for _ in range(args.warmup):
        cpu_leg(wl, 1)
    t = 0.0
    its = 0
    for s in range(args.steps):
        dt, it, _ = cpu_leg(wl, clips_per_step, first_clip=s * clips_per_step)
        t += dt
        its += it
    audio_s = args.steps * clips_per_step * wl["N"] / SR
    val = audio_s / t
    line = {
        "impl": "reference", "metric": "audio_seconds_restored_per_second", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "n_fft": wl["n_fft"], "hop": wl["hop"], "rank": wl["K"],
                   "max_iter": 200, "tol": 1e-4, "clips_per_step": clips_per_step},
        "nmf_iters_per_s": its / t,
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{clips_per_step} clip(s) of the workload per step through scipy.signal.stft/istft + "
                                   f"sklearn NMF(cd) (oracle/libcalls.py), OpenBLAS on {cores} threads, sweep single-threaded"},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


C5 = dict(N=158760000, n_fft=2048, hop=512, K=128, thr=1e-4, num=9, den=10, frac=0.9, seed=0,
          desc="BASELINE configs[4]: one 1-hour 44.1 kHz signal, 2 s gaps every 30 s, n_fft 2048 / hop 512, K=128, "
               "200 CD iterations; N>1: time-frame-sharded H + all-reduce of the W partial sums (strong scaling)")


def c5_signal_device(device, N):
    """0.1*N(0,1) + a bed of 8 sinusoids, zeroed on [s, s+2) s for s = 15, 45, ..., peak-normalised."""
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


def run_c5(args, rank, local_rank, world):
    """The 1-hour signal (BASELINE configs[4]).  N=1: torch.ops.ainmf.nmf_inpaint; N>1: TimeShardedInpainter."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import ainmf
    from ainmf import _capi
    L = ainmf._lib.lib()
    h = ainmf._lib.handle(local_rank)
    wl = dict(C5)
    if args.seconds > 0:
        wl["N"] = int(args.seconds * SR)
    N, K = wl["N"], wl["K"]
    T, F, _ = _capi.stft_geometry(L, N, wl["n_fft"], wl["hop"])
    x = c5_signal_device(device, N)
    if world > 1:
        from ainmf.sharding import TimeShardedInpainter
        tsi = TimeShardedInpainter(device)
        pl = tsi.plan(N, wl["n_fft"], wl["hop"])
        xl = x[pl["x_begin"]:pl["x_end"]].clone()
        del x
        torch.cuda.empty_cache()

        def step():
            return tsi.restore(xl, N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4, seed=wl["seed"],
                               threshold=wl["thr"], frac=(wl["num"], wl["den"]))
    else:
        xb = x[None]

        def step():
            out = ainmf.ops.nmf_inpaint(xb, wl["n_fft"], wl["hop"], K, 200, 1e-4, wl["seed"], wl["thr"], wl["num"], wl["den"],
                                        -1, -1, 1, None, None)
            return out[0][0], dict(n_bad=out[2], n_iter=out[6], err=out[5])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        y, info = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.ainmf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y, info = step()
    e1.record()
    barrier()
    launches = L.ainmf_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # kernel-level timing
    L.ainmf_profile(h, 1, None, None)
    step()
    torch.cuda.synchronize()
    pms, pcn = (C.c_double * 6)(), (C.c_int64 * 6)()
    L.ainmf_profile(h, 0, pms, pcn)
    names = ["gram_Ht", "xht_gram_tc", "w_side_fused", "gram_W", "h_step_tc", "stop_rule"]
    kern_ms = {k: float(pms[i]) for i, k in enumerate(names)}
    n_it = max(int(pcn[4]), 1)
    iter_ms = sum(kern_ms.values()) / n_it          # this rank's kernels; the all-reduce sits between them
    # end to end: pinned host -> device -> host around the same call
    src = xl if world > 1 else x
    xh = torch.empty(src.shape, dtype=torch.float32).pin_memory()
    xh.copy_(src)
    yh = torch.empty(y.shape, dtype=torch.float32).pin_memory()

    def e2e_step():
        src.copy_(xh, non_blocking=True)
        yy, _ = step()
        yh.copy_(yy, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    barrier()
    if rank == 0:
        peak, peak_src = peaks()
        Tl = T // world
        bytes_iter = alg_bytes_per_iter(F, Tl, K)        # per GPU: its slice of V and Ht, replicated W
        achieved = bytes_iter / (iter_ms * 1e-3) / 1e9
        h_bytes = 4.0 * F * Tl + 4.0 * F * K + 8.0 * Tl * K + 4.0 * K * K
        audio_s = N / SR
        line = {
            "metric": "audio_seconds_restored_per_second", "value": audio_s * args.steps / (float(ms[0]) * 1e-3),
            "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(ms[0]) / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5: " + wl["desc"], "n_samples": N, "n_fft": wl["n_fft"], "hop": wl["hop"], "F": F, "T": T,
                       "rank": K, "max_iter": 200, "tol": 1e-4, "solver": "cd",
                       "l2": "the spectrogram slice per GPU (%.2f GB) exceeds the 126 MB L2" % (F * Tl * 4 / 1e9)},
            "nmf_iters_per_s": 200 * args.steps / (float(ms[0]) * 1e-3),
            "n_iter": int(info["n_iter"][0]), "n_bad": int(info["n_bad"][0]), "objective": float(info["err"][0]),
            "e2e": {"value": audio_s * args.steps / (float(e2e_ms[0]) * 1e-3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": int(src.numel() * 4), "d2h_bytes_per_step": int(y.numel() * 4),
                    "ms_per_step": float(e2e_ms[0]) / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "h_step_ts_kernel on this rank's frame slice (X^T.W contraction on tcgen05 + H coordinate sweep)",
                         "achieved": h_bytes / (kern_ms["h_step_tc"] / n_it * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": h_bytes / (kern_ms["h_step_tc"] / n_it * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": h_bytes, "ms_per_launch": kern_ms["h_step_tc"] / n_it,
                         "iteration": {"kernel": "whole CD iteration on this rank (the W side is replicated, not sharded)",
                                       "achieved": achieved, "frac": achieved / peak,
                                       "algorithmic_bytes_per_iteration": bytes_iter, "ms_per_iteration": iter_ms},
                         "kernels_ms_per_launch": {k: v / n_it for k, v in kern_ms.items()}},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=0.0, help="c5 only: signal length override")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload == "c5":
        if args.warmup < 3:
            args.warmup = 3
        if args.impl == "reference":
            raise SystemExit("--impl reference supports the clip workloads (c4, c2)")
        return run_c5(args, int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
                      int(os.environ.get("WORLD_SIZE", "1")))
    wl = dict(WORKLOADS[args.workload])
    if args.clips > 0:
        wl["clips"] = args.clips
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ainmf has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import ainmf
    from ainmf import _capi
    ops = ainmf.ops
    L = ainmf._lib.lib()
    h = ainmf._lib.handle(local_rank)

    B, N = wl["clips"], wl["N"]
    T, F, _ = _capi.stft_geometry(L, N, wl["n_fft"], wl["hop"])
    K = wl["K"]
    x = synth_device(wl, rank * B, B, device)
    torch.cuda.synchronize()

    def step():
        return ops.nmf_inpaint(x, wl["n_fft"], wl["hop"], K, 200, 1e-4, wl["seed"], wl["thr"], wl["num"], wl["den"],
                               -1, -1, 1, None, None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg (`value`) -------------------------------------------------------------
    for _ in range(args.warmup):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.ainmf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    launches = L.ainmf_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms[0])
    y, idx, nb, Wf, Hf, err, nit = out
    iters_done = int(nit.sum())

    # ---- kernel-level timing of the NMF iteration (CUDA events on the launching stream, outside the timed legs) ----
    L.ainmf_profile(h, 1, None, None)
    step()
    torch.cuda.synchronize()
    pms = (C.c_double * 6)()
    pcn = (C.c_int64 * 6)()
    L.ainmf_profile(h, 0, pms, pcn)
    kern_names = ["gram_Ht", "xht_gram_tc", "w_side_fused", "gram_W", "h_step_tc", "stop_rule"]
    kern_ms = {k: float(pms[i]) for i, k in enumerate(kern_names)}
    n_it = max(int(pcn[4]), 1)
    iter_ms = sum(kern_ms.values()) / n_it
    peak, peak_src = peaks()
    bytes_iter = alg_bytes_per_iter(F, T, K) * B
    achieved = bytes_iter / (iter_ms * 1e-3) / 1e9
    # per-kernel algorithmic bytes (float32): what each launch must move at least
    # (gram_Ht / gram_W are separate launches only on the FFMA path, K < 64: on the tensor-core path the Gram of Ht comes
    # out of the X.Ht kernel and W^T W out of the fused W-side kernel, so their entries read 0)
    kb = {"gram_Ht": 4.0 * T * K + 4.0 * K * K, "xht_gram_tc": 4.0 * F * T + 4.0 * T * K + 4.0 * F * K,
          "w_side_fused": 12.0 * F * K + 8.0 * K * K, "gram_W": 4.0 * F * K + 4.0 * K * K,
          "h_step_tc": 4.0 * F * T + 4.0 * F * K + 8.0 * T * K + 4.0 * K * K, "stop_rule": 0.0}
    kernels = {k: {"ms_per_launch": kern_ms[k] / n_it, "share": kern_ms[k] / max(sum(kern_ms.values()), 1e-12),
                   "hbm_frac": ((kb[k] * B / (kern_ms[k] / n_it * 1e-3) / 1e9) / peak) if kern_ms[k] > 0 else None}
               for k in kern_names}

    # roofline of the dominant kernel (the contract's object) + the same accounting for the whole iteration
    dom = max(kern_names, key=lambda k: kern_ms[k])
    dom_kernel = {"h_step_tc": "h_step_ts_kernel (X^T.W contraction on tcgen05 + H coordinate sweep, one launch per iteration)",
                  "xht_gram_tc": "xht_ts_kernel (X.Ht contraction + Gram of Ht on tcgen05, one launch per iteration)",
                  "w_side_fused": "w_side_kernel"}.get(dom, dom)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01d_traffic.json")
    if os.path.exists(tpath) and args.workload == "c4":
        tj = json.load(open(tpath))
        key = {"h_step_tc": "h_step_ts_kernel", "xht_gram_tc": "xht_ts_kernel", "w_side_fused": "w_side_kernel"}.get(dom)
        if key in tj["dram_bytes_per_launch"]:
            traffic = tj["dram_bytes_per_launch"][key] / tj["clips"] * B     # ncu dram read+write per launch, scaled to B clips
    dom_ms = kern_ms[dom] / n_it
    dom_achieved = kb[dom] * B / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_kernel, "achieved": dom_achieved, "peak": peak, "unit": "GB/s",
                "frac": dom_achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kb[dom] * B, "ms_per_launch": dom_ms,
                "limiter": "tensor pipe: 2 tcgen05.mma per 8 contraction elements (tf32 main term + bf16 cross terms), ~277 cycles per 128 frames x 8 bins",
                "iteration": {"kernel": "whole CD iteration = xht_ts_kernel + reduce_splits + w_side_kernel + w_finish_kernel + h_step_ts_kernel + stop_kernel",
                              "achieved": achieved, "frac": achieved / peak, "algorithmic_bytes_per_iteration": bytes_iter,
                              "ms_per_iteration": iter_ms},
                "kernels": kernels}

    # ---- end-to-end leg through the C ABI with HOST buffers -----------------------------------------
    xh = torch.empty((B, N), dtype=torch.float32).pin_memory()
    xh.copy_(x)
    yh = torch.empty((B, N), dtype=torch.float32).pin_memory()
    nbh = np.zeros(B, np.int32); errh = np.zeros(B, np.float32); nih = np.zeros(B, np.int32)
    p = _capi.default_params(L, batch=B, n_samples=N, n_fft=wl["n_fft"], hop=wl["hop"], rank=K, max_iter=200, tol=1e-4,
                             seed=wl["seed"], threshold=wl["thr"], frac_num=wl["num"], frac_den=wl["den"])
    del out, y, idx, Wf, Hf
    ops._workspaces.clear()
    torch.cuda.empty_cache()

    def e2e_step():
        rc = L.ainmf_inpaint_host(h, C.byref(p), C.c_void_p(xh.data_ptr()), C.c_void_p(yh.data_ptr()),
                                  nbh.ctypes.data_as(C.c_void_p), errh.ctypes.data_as(C.c_void_p),
                                  nih.ctypes.data_as(C.c_void_p), 0)
        ainmf._lib.check(rc, local_rank)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    barrier()
    audio_s_step = world * B * N / SR
    value = audio_s_step * args.steps / (ms_total * 1e-3)
    e2e_value = audio_s_step * args.steps / (float(e2e_ms[0]) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity spot check + CPU baseline (rank 0, N=1 only; bounded sample) -------------------------
    parity, cpu = None, None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import libcalls
        n_cpu = 8 if B > 1 else 1
        cores = use_all_host_threads()
        dt, its, outs = cpu_leg(wl, n_cpu)
        cpu = {"value": n_cpu * N / SR / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} clips of the workload, serially, through scipy.signal.stft/istft + sklearn NMF(cd) "
                         f"(oracle/libcalls.py); OpenBLAS threads = {cores}, coordinate sweep single-threaded",
               "nmf_iters_per_s": its / dt, "seconds": dt}
        xs = np.stack([o[0] for o in outs[:2]])
        yg, ig, ng, _, _, eg, itg = ops.nmf_inpaint(torch.from_numpy(xs).to(device), wl["n_fft"], wl["hop"], K, 200, 1e-4,
                                                    wl["seed"], wl["thr"], wl["num"], wl["den"], -1, -1, 1, None, None)
        parity = []
        for i in range(xs.shape[0]):
            _, yo, st = outs[i]
            bad = st["bad"]
            n = int(ng[i])
            parity.append({"mask_bit_exact": bool(n == len(bad) and np.array_equal(ig[i, :n].cpu().numpy(), bad)),
                           "objective_rel_diff": abs(float(eg[i]) - st["err"]) / st["err"],
                           "snr_vs_oracle_db": float(libcalls.snr_db(yo, yg[i].cpu().numpy())),
                           "n_iter": [int(itg[i]), st["n_iter"]]})

    line = {
        "metric": "audio_seconds_restored_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "clips_per_gpu": B, "n_samples": N, "n_fft": wl["n_fft"],
                   "hop": wl["hop"], "F": F, "T": T, "rank": K, "max_iter": 200, "tol": 1e-4, "solver": "cd",
                   "l2": "inputs (%.0f MB/GPU of waveform, %.1f GB of spectrogram) exceed the 126 MB L2" % (B * N * 4 / 1e6, B * F * T * 4 / 1e9)},
        "nmf_iters_per_s": world * B * 200 / (sum(kern_ms.values()) * 1e-3) if sum(kern_ms.values()) > 0 else None,
        "nmf_iterations_done_last_step": iters_done,
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": B * N * 4, "d2h_bytes_per_step": B * N * 4 + B * 12,
                "ms_per_step": float(e2e_ms[0]) / args.steps, "api": "ainmf_inpaint_host (C ABI, pinned host buffers)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
