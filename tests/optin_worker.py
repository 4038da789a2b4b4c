"""Worker of tests/test_gpu_optin_paths.py: a fixed set of small operator applies and Lanczos runs on cuda:0, results
printed as one JSON line.  The test runs it once per environment switch (LLZ_SELL_TMA=1, LLZ_SPMV=v, LLZ_BASIS_VMM=0,
LLZ_FUSED_ORTH=0, LLZ_XXZ_KERNEL=state — all read when the library / context / operator is created, hence a process
each) and compares with the default path."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def main():
    pkg = entry.load_package()
    wl = importlib.import_module("lambda_lanczos_b200.workloads")
    ctx = pkg.Context(0)
    out = {}
    mats = {"laplacian": wl.laplacian2d_csr(97, 83), "random": wl.random_symmetric_csr(5003, 9),
            "peierls": wl.peierls_csr(31, 29), "ragged": wl.random_symmetric_csr(4001, 40)}
    for name, csr in mats.items():
        n = csr[0].size - 1
        x = wl.start_vector(n, csr[2].dtype, seed=11)
        y_sell = pkg.Operator.sell(ctx, *csr).matvec(x)
        y_csr = pkg.Operator.csr(ctx, *csr).matvec(x)
        out[f"sell/{name}"] = digest(y_sell)
        out[f"csr/{name}"] = digest(y_csr)
        out[f"csr_vals/{name}"] = [float(v) for v in np.abs(y_csr[:64])]
    for L in (12, 15):
        x = wl.start_vector(wl.xxz_dim(L) if hasattr(wl, "xxz_dim") else pkg.Operator.xxz(ctx, L).n, np.float64, seed=5)
        out[f"xxz/{L}"] = digest(pkg.Operator.xxz(ctx, L).matvec(x))
    csr = wl.random_symmetric_csr(20000)
    for fmt in ("csr", "sell"):
        op = getattr(pkg.Operator, fmt)(ctx, *csr)
        eng = pkg.LambdaLanczos(op, 20000, True, 2)
        eng.init_vector = wl.start_vector(20000)
        ev, vec = eng.run()
        out[f"lanczos/{fmt}"] = {"ev": [float(v) for v in ev], "its": eng.getIterationCounts(), "vec": digest(vec)}
    lap = wl.laplacian2d_csr(40)
    eng = pkg.LambdaLanczos(pkg.Operator.sell(ctx, *lap), 1600, False, 4)
    eng.init_vector = wl.start_vector(1600)
    ev, vec = eng.run()
    out["lanczos/laplacian4"] = {"ev": [float(v) for v in ev], "its": eng.getIterationCounts(), "vec": digest(vec)}
    opc = pkg.Operator.xxz(ctx, 12, dtype=np.complex128)
    it, psi = pkg.Exponentiator(opc, opc.n).run(-0.2j, wl.neel_state(12))
    out["expm/xxz12"] = {"its": it, "vec": digest(psi)}
    print("OPTIN_RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
