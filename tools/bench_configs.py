#!/usr/bin/env python
"""All five BASELINE.json configurations on the GPU(s): iterations, time-to-eigenpair, iterations/s and the achieved
fraction of the HBM roofline from the bytes model of SURVEY.md §8d, one JSON line per configuration.

    python tools/bench_configs.py [c1] [c2] [c3] [c4] [c5] [--L4 28] [--L5 28] [--steps5 100] [--cpu]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_configs.py c4 --L4 30     (row-sharded)

Not the driver's bench (that is bench.py, config 2); this is the table for RESULTS.md.  `--cpu` also times the compiled
reference (oracle/_ref) on config 1 on the host cores (the only configuration it finishes in seconds).
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--L4", type=int, default=28)
    ap.add_argument("--L5", type=int, default=28)
    ap.add_argument("--steps5", type=int, default=100)
    ap.add_argument("--cap", type=int, default=192, help="max_iteration for the configurations that cannot converge in HBM")
    ap.add_argument("--cap4", type=int, default=0, help="max_iteration for config 4 (0 = to convergence; a single GPU holds ~120 vectors of L=30)")
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as entry

    pkg = entry.load_package()
    wl = importlib.import_module("lambda_lanczos_b200.workloads")
    ctx = pkg.Context(local_rank)
    if world > 1:
        blob = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            blob.copy_(torch.frombuffer(bytearray(pkg.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(blob, src=0)
        ctx.join(rank, world, blob.cpu().numpy().tobytes())
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    def sync():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    def lanczos_case(name, make_op, n, dtype, find_max, num_eigs, max_iteration=None, start=None, extra=None, repeat=2):
        row0, nl = wl.partition(n, rank, world)
        op = make_op(row0, nl)
        s = np.dtype(dtype).itemsize
        if start is None:
            start = wl.start_vector(n, dtype)[row0:row0 + nl]
        best = None
        for rep in range(repeat):  # first repetition maps the basis memory (kept by the context afterwards)
            eng = pkg.LambdaLanczos(op, n, find_max, num_eigs)
            eng.init_vector = start
            if max_iteration:
                eng.max_iteration = max_iteration
            eng.want_eigenvectors = False
            sync()
            t0 = time.perf_counter()
            ev, _ = eng.run()
            sync()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, ev, eng.getIterationCounts(), eng.stats.seconds_host)
        dt, ev, counts, host_s = best
        # SURVEY.md §8d bytes model, per GPU
        a_bytes, q, total = op.bytes(), 0, 0.0
        for m in counts:
            total += m * a_bytes + (m * (m + 1) + (2 * q + 7) * m) * nl * s + (m + 5) * nl * s
            q = min(num_eigs, q + 5)
        d = {"config": name, "n": n, "n_gpus": world, "dtype": str(np.dtype(dtype)), "iterations": counts, "seconds": dt,
             "iterations_per_s": sum(counts) / dt, "host_seconds": host_s, "eigenvalues": [float(x) for x in ev],
             "model_GB_per_gpu": total / 1e9, "model_GBps_per_gpu": total / dt / 1e9, "frac_of_measured_peak": total / dt / 1e9 / peak,
             "frac_of_8TBps": total / dt / 1e9 / 8000.0, "first_run_seconds_incl_basis_mapping": None}
        if extra:
            d.update(extra)
        emit(d)
        return d

    for c in args.configs:
        if c == "c1":
            n = 100000
            full = wl.random_symmetric_csr(n)

            def mk(row0, nl):
                return pkg.Operator.sell(ctx, *wl.csr_row_block(*full, row0, nl), row0=row0, n_cols=n)

            d = lanczos_case("config1 random symmetric CSR n=100k ~17 nnz/row, max eigenpair, to convergence", mk, n, np.float64, True, 1, repeat=3)
            if args.cpu and rank == 0:
                import oracle

                impl = oracle.best()
                t0 = time.perf_counter()
                r = impl.lanczos(*full, find_max=True, num_eigs=1, init=wl.start_vector(n))
                dt = time.perf_counter() - t0
                emit({"config": "config1 on the host CPU", "impl": impl.kind, "iterations": r.iter_counts, "seconds": dt,
                      "iterations_per_s": sum(r.iter_counts) / dt, "eigenvalues": [float(x) for x in r.eigenvalues],
                      "rel_diff_vs_gpu": abs(r.eigenvalues[0] - d["eigenvalues"][0]) / abs(r.eigenvalues[0])})
        elif c == "c2":
            nx = 4096

            def mk(row0, nl):
                return pkg.Operator.sell(ctx, *wl.laplacian2d_csr_rows(nx, row0, nl), row0=row0, n_cols=nx * nx)

            lanczos_case(f"config2 Laplacian 4096^2 double, 4 smallest, max_iteration={args.cap}", mk, nx * nx, np.float64, False, 4, args.cap)
        elif c == "c3":
            lx = 2896
            n = lx * lx
            full = wl.peierls_csr(lx, lx, flux=0.05, trap=0.02)

            def mk(row0, nl):
                return pkg.Operator.sell(ctx, *wl.csr_row_block(*full, row0, nl), row0=row0, n_cols=n)

            lanczos_case(f"config3 Peierls tight-binding {lx}^2 (n={n}) complex128, 2 lowest, max_iteration={args.cap}", mk, n, np.complex128, False, 2,
                         args.cap)
            del full
        elif c == "c4":
            L = args.L4

            def mk(row0, nl):
                return pkg.Operator.xxz(ctx, L)

            n = int(round(np.exp(sum(np.log(np.arange(L // 2 + 1, L + 1))) - sum(np.log(np.arange(1, L // 2 + 1))))))
            op0 = pkg.Operator.xxz(ctx, L)
            n = op0.n_global
            del op0
            row0, nl = wl.partition(n, rank, world)
            rs = np.random.RandomState(1 + rank)  # start vector generated per block (a 155M-entry global vector per rank is wasteful)
            start = rs.uniform(-1, 1, nl)
            lanczos_case(f"config4 XXZ chain L={L} Sz=0 (dim {n}) matrix-free, ground state" + (f", max_iteration={args.cap4}" if args.cap4 else ""),
                         mk, n, np.float64, False, 1, args.cap4 or None, start=start, extra={"note": "start vector seeded per row block"})
        elif c == "c5":
            L = args.L5
            op = pkg.Operator.xxz(ctx, L, dtype=np.complex128)
            n = op.n_global
            row0, nl = wl.partition(n, rank, world)
            # Neel state |0101...>: rank of the state computed from the Lin-table formula is not needed on the host — the
            # start is a unit vector; find its index with the combinatorial number system
            neel = sum(1 << b for b in range(0, L, 2))
            from math import comb

            idx, j = 0, 0
            for b in range(L):
                if neel >> b & 1:
                    j += 1
                    idx += comb(b, j)
            psi = np.zeros(nl, dtype=np.complex128)
            if row0 <= idx < row0 + nl:
                psi[idx - row0] = 1
            ex = pkg.Exponentiator(op, n)
            lib = pkg.lib()
            import ctypes as C

            vin, vout = pkg.Vector.from_host(ctx, psi), pkg.Vector(ctx, np.complex128, nl)
            din, dout = C.c_void_p(), C.c_void_p()
            lib.llz_vec_device_ptr(vin.h, C.byref(din))
            lib.llz_vec_device_ptr(vout.h, C.byref(dout))
            its = []
            sync()
            t0 = time.perf_counter()
            for step in range(args.steps5):  # device-resident time evolution: output feeds the next step
                it = C.c_int64(0)
                st = lib.llz_expm_run(ctx.h, op.h, C.c_int(3), (C.c_double * 2)(0.0, -0.1), din, dout, C.c_int(0), C.c_double(-1.0), C.c_int(0),
                                      C.c_int64(0), C.c_int(0), C.byref(it))
                assert st == 0, lib.llz_last_error()
                its.append(int(it.value))
                din, dout = dout, din
            sync()
            dt = time.perf_counter() - t0
            fin = (vin if args.steps5 % 2 == 0 else vout)
            nrm = fin.norm()
            total = sum((op.bytes() + 8 * nl * 16) * m + (m + 1) * nl * 16 for m in its)
            emit({"config": f"config5 Exponentiator exp(-i H 0.1) on XXZ L={L} (dim {n}) complex128, {args.steps5} steps from the Neel state",
                  "n_gpus": world, "iterations_per_step": its[:5] + ["..."] + its[-2:], "total_iterations": sum(its), "seconds": dt,
                  "steps_per_s": args.steps5 / dt, "iterations_per_s": sum(its) / dt, "final_norm": nrm,
                  "model_GBps_per_gpu": total / dt / 1e9, "frac_of_measured_peak": total / dt / 1e9 / peak})
    if dist is not None:
        sync()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
