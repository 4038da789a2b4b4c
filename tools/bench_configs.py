#!/usr/bin/env python
"""All five BASELINE.json configurations on the GPU(s): iterations, time-to-eigenpair, iterations/s and the achieved
fraction of the HBM roofline from the bytes model of SURVEY.md §8d, one JSON line per configuration.

    python tools/bench_configs.py [c1] [c2] [c3] [c4] [c5] [--L4 28] [--L5 28] [--steps5 100] [--cpu]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_configs.py c4 --L4 30     (row-sharded)

The runners live in tools/configs.py (bench.py uses the same ones for its `extra` entries); this is the stand-alone
table for RESULTS.md.  `--cpu` also times the compiled reference (oracle/_ref) on config 1 on the host cores (the only
configuration it finishes in seconds).
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
sys.path.insert(0, os.path.join(ROOT, "tools"))


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
    import configs as cf

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
        d.pop("_csr", None)
        if rank == 0:
            print(json.dumps(d), flush=True)

    env = cf.Env(pkg, wl, ctx, rank, world, sync=sync, peak=peak)
    for c in args.configs:
        ctx.release_cache()
        if c == "c1":
            d = cf.run_c1(env)
            full = d.pop("_csr")
            emit(d)
            if args.cpu and rank == 0:
                import oracle

                impl = oracle.best(fast=True)
                t0 = time.perf_counter()
                r = impl.lanczos(*full, find_max=True, num_eigs=1, init=wl.start_vector(100000))
                dt = time.perf_counter() - t0
                emit({"config": "config1 on the host CPU", "impl": impl.kind, "iterations": r.iter_counts, "seconds": dt,
                      "iterations_per_s": sum(r.iter_counts) / dt, "eigenvalues": [float(x) for x in r.eigenvalues],
                      "rel_diff_vs_gpu": abs(r.eigenvalues[0] - d["eigenvalues"][0]) / abs(r.eigenvalues[0])})
        elif c == "c2":
            nx = 4096
            row0, nl = wl.partition(nx * nx, rank, world)
            op = pkg.Operator.sell(ctx, *wl.laplacian2d_csr_rows(nx, row0, nl), row0=row0, n_cols=nx * nx)
            emit(cf.lanczos_case(env, f"config2: Laplacian 4096^2 double, 4 smallest, max_iteration={args.cap}", op, nx * nx, np.float64,
                                 False, 4, args.cap))
        elif c == "c3":
            emit(cf.run_c3(env, args.cap))
        elif c == "c4":
            emit(cf.run_c4(env, args.L4, args.cap4))
        elif c == "c5":
            emit(cf.run_c5(env, args.L5, args.steps5))
    if dist is not None:
        sync()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
