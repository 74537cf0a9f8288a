#!/usr/bin/env python
"""Turn the raw ncu outputs of profiles/run_profiles.sh (in gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py r02            # c4: gpurun_out/launches_r02.csv, gpurun_out/prof_r02.ncu-rep
    python profiles/summarize.py c5_r02 c5 51680   # c5 capture on a 600 s prefix: workload name and frames per launch

Writes <tag>_launches.csv (per-kernel totals and shares), <tag>_ncu_full.csv (selected metrics of the captured launches)
and <tag>_traffic.json (dram bytes per launch, what bench.py puts into roofline.traffic, scaled by units).
"""
import csv, collections, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
workload = sys.argv[2] if len(sys.argv) > 2 else "c4"
units = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

# ---- launch list ---------------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("ainmf::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
allus = sum(tot.values())
with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
    f.write("kernel,launches,total_us,share\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"\"{k}\",{cnt[k]},{v:.1f},{v / allus:.4f}\n")

# ---- full capture ---------------------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", os.path.join(G, f"prof_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tmem.sum", "sm__inst_executed_pipe_tc.sum",
        "smsp__inst_executed_pipe_uniform.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = [next((i for i, x in enumerate(h) if x == k or x.endswith("." + k)), None) for k in keep]
with open(os.path.join(P, f"{tag}_ncu_full.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([u[i] if i is not None else "" for i in idx])
    for r in rr[2:]:
        w.writerow([r[i] if i is not None else "" for i in idx])
print(open(os.path.join(P, f"{tag}_launches.csv")).read())
print(open(os.path.join(P, f"{tag}_ncu_full.csv")).read())

# ---- dram traffic per launch (bench.py: roofline.traffic) -------------------------------------------------------
ki = h.index("Kernel Name")
ri = next(i for i, x in enumerate(h) if x.endswith("dram__bytes_read.sum"))
wi = next(i for i, x in enumerate(h) if x.endswith("dram__bytes_write.sum"))
def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
traffic = {}
for r in rr[2:]:
    name = r[ki].split("<")[0].split("(")[0].replace("void ", "").replace("ainmf::", "")
    traffic[name] = to_bytes(r[ri], u[ri]) + to_bytes(r[wi], u[wi])
json.dump({"source": f"profiles/{tag}_ncu_full.csv (ncu --set full, one launch of each kernel)", "workload": workload, "units": units,
           "unit_is": "clips per launch" if workload != "c5" else "frames per launch", "dram_bytes_per_launch": traffic},
          open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_traffic.json")).read())
# tcgen05-related counters the capture holds (names vary by ncu version): printed so that they can be cited
for i, x in enumerate(h):
    if any(s_ in x for s_ in ("tensor", "_tc", "tmem", "uniform")):
        print(x, [r[i] for r in rr[2:]])
