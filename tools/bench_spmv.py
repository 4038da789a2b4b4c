"""SpMV micro-benchmark on the GPU box: achieved algorithmic GB/s of every operator kind at BASELINE.json's shapes.
    python tools/bench_spmv.py [laplacian] [peierls] [xxz26] [xxz28] [random]
Times 20 applies with the context's CUDA-event profiler (the "spmv" family)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e
pkg = e.load_package(); wl = importlib.import_module("lambda_lanczos_b200.workloads")
ctx = pkg.Context(0)
which = sys.argv[1:] or ["laplacian", "peierls", "random", "xxz26"]

def run(name, op, dtype):
    n = op.n
    x = pkg.Vector.from_host(ctx, wl.start_vector(n, dtype)); y = pkg.Vector(ctx, dtype, n)
    for _ in range(3): op.apply(x, y)
    ctx.synchronize(); ctx.profile(True)
    for _ in range(20): op.apply(x, y)
    ctx.synchronize(); ms, cnt, by = ctx.profile_read("spmv"); ctx.profile(False)
    itemsize = np.dtype(dtype).itemsize
    alg = op.bytes() + 2 * n * itemsize
    print(f"{name:28s} [{op.storage():22s}] n={n:10d} {ms/cnt*1e3:9.1f} us/apply  A_bytes={op.bytes()/1e6:9.1f} MB  algorithmic {alg/(ms/cnt*1e-3)/1e9:7.0f} GB/s"
          f"  (x,y-only {2*n*itemsize/(ms/cnt*1e-3)/1e9:6.0f} GB/s)", flush=True)

for w in which:
    if w == "laplacian":
        csr = wl.laplacian2d_csr(4096)
        run("laplacian4096 csr", pkg.Operator.csr(ctx, *csr), np.float64)
        run("laplacian4096 auto", pkg.Operator.sell(ctx, *csr), np.float64)
        run("laplacian4096 sell sigma=1", pkg.Operator.sell(ctx, *csr, sigma=1), np.float64)
    elif w == "peierls":
        csr = wl.peierls_csr(2896, 2896)
        run("peierls2896 csr c128", pkg.Operator.csr(ctx, *csr), np.complex128)
        run("peierls2896 auto c128", pkg.Operator.sell(ctx, *csr), np.complex128)
        run("peierls2896 sell sigma=1 c128", pkg.Operator.sell(ctx, *csr, sigma=1), np.complex128)
    elif w == "random":
        csr = wl.random_symmetric_csr(100000)
        run("random100k csr", pkg.Operator.csr(ctx, *csr), np.float64)
        run("random100k sell", pkg.Operator.sell(ctx, *csr), np.float64)
        csr = wl.random_symmetric_csr(4000000)
        run("random4M csr", pkg.Operator.csr(ctx, *csr), np.float64)
        run("random4M sell", pkg.Operator.sell(ctx, *csr), np.float64)
    elif w.startswith("xxz"):
        L = int(w[3:])
        run(f"xxz L={L} f64", pkg.Operator.xxz(ctx, L), np.float64)
        if L <= 28:
            run(f"xxz L={L} c128", pkg.Operator.xxz(ctx, L, dtype=np.complex128), np.complex128)
