"""Print a one-line summary of bench.py JSON outputs (developer helper for gpurun sessions)."""
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable", e); continue
    r = d.get("roofline", {})
    k = r.get("kernels") or r.get("kernels_ms_per_launch") or {}
    ks = {n: round(v["ms_per_launch"] if isinstance(v, dict) else v, 4) for n, v in k.items()}
    it = r.get("iteration", r)
    print(p, "value %.0f e2e %.0f frac(dominant) %.3f frac(iteration) %.3f ms/iter %.4f" % (d["value"], d["e2e"]["value"], r.get("frac", 0), it.get("frac", 0), it.get("ms_per_iteration", 0)),
          "obj", d.get("objective"), "parity", d.get("parity"))
    print("   ", ks)
