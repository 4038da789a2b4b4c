"""Time operator creation (CSR vs SELL) at the bench size; run on the GPU box."""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e
pkg = e.load_package(); wl = importlib.import_module("lambda_lanczos_b200.workloads")
ctx = pkg.Context(0)
csr = wl.laplacian2d_csr(4096)
for rep in range(3):
    for name, make in (("csr", pkg.Operator.csr), ("sell", pkg.Operator.sell)):
        t0 = time.perf_counter(); op = make(ctx, *csr); ctx.synchronize(); t1 = time.perf_counter()
        op.close(); ctx.synchronize(); t2 = time.perf_counter()
        print(f"{name}: create {t1-t0:.3f} s, destroy {t2-t1:.3f} s", flush=True)
