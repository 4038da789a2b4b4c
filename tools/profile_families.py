"""Per-kernel-family device time (CUDA events around every launch, llz_ctx_profile) of one BASELINE configuration.
    python tools/profile_families.py c1|c2|c3|c4|c5 [L]
Event-bracketed launches serialise the host, so the SUM is not a throughput figure; the per-family averages are."""
import importlib, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import __graft_entry__ as e
import configs as cf
pkg = e.load_package(); wl = importlib.import_module("lambda_lanczos_b200.workloads")
ctx = pkg.Context(0)
env = cf.Env(pkg, wl, ctx)
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 28
run = {"c1": lambda: cf.run_c1(env), "c3": lambda: cf.run_c3(env), "c4": lambda: cf.run_c4(env, L, 100 if L >= 30 else 0),
       "c5": lambda: cf.run_c5(env, L, 20)}[which]
d = run(); d.pop("_csr", None)
print("unprofiled:", json.dumps({k: d[k] for k in ("iterations", "seconds", "iterations_per_s", "frac_of_measured_peak") if k in d}))
ctx.profile(True)
d = run(); d.pop("_csr", None)
fams = ("spmv", "halo", "exchange", "orth", "project", "reduce", "update", "scale", "combine", "dot", "recurrence")
tot = 0.0
for f in fams:
    ms, cnt, by = ctx.profile_read(f)
    if cnt:
        tot += ms
        print(f"{f:10s} launches {cnt:7d}  total {ms:9.3f} ms  avg {ms/cnt*1e3:9.2f} us  {by/(ms*1e-3)/1e9 if ms > 0 else 0:8.0f} GB/s (algorithmic)")
print(f"sum of families {tot:.3f} ms; profiled run {d['seconds']*1e3:.3f} ms")
ctx.profile(False)
