"""XXZ apply micro-benchmark on the GPU box: the block kernel under several splits / CTA sizes and the per-state kernel.
    python tools/bench_xxz.py L [variants...]     variant = state | block[:m], e.g. block:12
Times 20 applies with the context's CUDA-event profiler; GB/s counts 2 n s bytes (x read once, y written once)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e
pkg = e.load_package(); wl = importlib.import_module("lambda_lanczos_b200.workloads")
ctx = pkg.Context(0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 28
variants = sys.argv[2:] or ["state", "block", "block:11", "block:13"]
dtypes = [np.float64, np.complex128] if L <= 28 else [np.float64]
for dtype in dtypes:
    ref = None
    for v in variants:
        parts = v.split(":")
        for k in ("LLZ_XXZ_KERNEL", "LLZ_XXZ_M"):
            os.environ.pop(k, None)
        if parts[0] == "state":
            os.environ["LLZ_XXZ_KERNEL"] = "state"
        if len(parts) > 1:
            os.environ["LLZ_XXZ_M"] = parts[1]
        op = pkg.Operator.xxz(ctx, L, dtype=dtype)
        n = op.n
        x = pkg.Vector.from_host(ctx, wl.start_vector(n, dtype)); y = pkg.Vector(ctx, dtype, n)
        for _ in range(3): op.apply(x, y)
        ctx.synchronize(); ctx.profile(True)
        for _ in range(20): op.apply(x, y)
        ctx.synchronize(); ms, cnt, by = ctx.profile_read("spmv"); ctx.profile(False)
        s = np.dtype(dtype).itemsize
        yh = y.download()
        if ref is None:
            ref = yh
        same = bool(np.array_equal(yh.view(np.uint8), ref.view(np.uint8)))
        print(f"xxz L={L} {np.dtype(dtype).name:10s} {v:16s} n={n} {ms/cnt*1e3:9.1f} us/apply  A_bytes={op.bytes()/1e6:8.2f} MB  "
              f"{(op.bytes() + 2*n*s)/(ms/cnt*1e-3)/1e9:7.0f} GB/s algorithmic  ({(4*n + 2*n*s)/(ms/cnt*1e-3)/1e9:7.0f} GB/s by the r01 count)  "
              f"bit-identical to first: {same}", flush=True)
        del op, x, y
        ctx.release_cache()
