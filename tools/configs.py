"""Runners for the BASELINE.json configurations other than the headline one (config 2 is bench.py's own workload).

Shared by ``bench.py`` (the ``extra`` entries of its JSON line) and ``tools/bench_configs.py`` (the stand-alone table).
Every runner returns one dict: iterations, seconds (wall clock around a synchronised region, after a warm-up run that
also maps the basis memory), iterations/s, and the achieved fraction of the HBM roofline from the bytes model of
SURVEY.md §8d,  B_run = m A_bytes + (m(m+1) + (2q+7) m) n s + (m + r) n s  per Lanczos run (Exponentiator:
(A_bytes + 8 n s) per iteration + (m+1) n s per step), per GPU.
"""
from __future__ import annotations

import time
from math import comb

import numpy as np


def lanczos_bytes(counts, n_local, s, a_bytes, num_eigs, nroot=5):
    total, q = 0.0, 0
    for m in counts:
        total += m * a_bytes + (m * (m + 1) + (2 * q + 7) * m) * n_local * s + (m + nroot) * n_local * s
        q = min(num_eigs, q + nroot)
    return total


class Env:
    """What a runner needs: the package, the workloads module, the context, rank/world and a group-wide sync."""

    def __init__(self, pkg, wl, ctx, rank=0, world=1, sync=None, peak=6541.8):
        self.pkg, self.wl, self.ctx, self.rank, self.world, self.peak = pkg, wl, ctx, rank, world, peak
        self.sync = sync or ctx.synchronize


def lanczos_case(env: Env, name, op, n, dtype, find_max, num_eigs, max_iteration=None, start=None, repeat=2, extra=None):
    wl, pkg = env.wl, env.pkg
    row0, nl = wl.partition(n, env.rank, env.world)
    s = np.dtype(dtype).itemsize
    if start is None:
        start = wl.start_vector(n, dtype)[row0:row0 + nl]
    best = None
    for _ in range(repeat):  # the first repetition maps the basis memory (kept by the context afterwards)
        eng = pkg.LambdaLanczos(op, n, find_max, num_eigs)
        eng.init_vector = start
        if max_iteration:
            eng.max_iteration = max_iteration
        eng.want_eigenvectors = False
        env.sync()
        t0 = time.perf_counter()
        ev, _ = eng.run()
        env.sync()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, ev, eng.getIterationCounts(), eng.stats.seconds_host)
    dt, ev, counts, host_s = best
    total = lanczos_bytes(counts, nl, s, op.bytes(), num_eigs)
    capped = bool(max_iteration) and all(c >= max_iteration for c in counts)
    d = {"config": name, "n": int(n), "n_gpus": env.world, "dtype": str(np.dtype(dtype)), "iterations": counts,
         "seconds": dt, "iterations_per_s": sum(counts) / dt, "converged": not capped,
         "time_to_eigenpair_s": None if capped else dt, "host_seconds": host_s,
         "eigenvalues": [float(x) for x in ev], "model_GBps_per_gpu": total / dt / 1e9,
         "frac_of_measured_peak": total / dt / 1e9 / env.peak, "frac_of_8TBps": total / dt / 1e9 / 8000.0}
    if extra:
        d.update(extra)
    return d


def run_c1(env: Env, fmt="sell"):
    wl, pkg = env.wl, env.pkg
    n = 100000
    full = wl.random_symmetric_csr(n)
    row0, nl = wl.partition(n, env.rank, env.world)
    make = pkg.Operator.sell if fmt == "sell" else pkg.Operator.csr
    op = make(env.ctx, *wl.csr_row_block(*full, row0, nl), row0=row0, n_cols=n)
    d = lanczos_case(env, "config1: random symmetric CSR n=100k ~17 nnz/row, double, max eigenpair, to convergence", op, n,
                     np.float64, True, 1, repeat=3)
    d["_csr"] = full
    return d


def run_c3(env: Env, cap=192, lx=2896):
    wl, pkg = env.wl, env.pkg
    n = lx * lx
    full = wl.peierls_csr(lx, lx, flux=0.05, trap=0.02)
    row0, nl = wl.partition(n, env.rank, env.world)
    op = pkg.Operator.sell(env.ctx, *wl.csr_row_block(*full, row0, nl), row0=row0, n_cols=n)
    del full
    return lanczos_case(env, f"config3: Peierls tight-binding {lx}^2 (n={n}) complex128, 2 lowest, max_iteration={cap}", op, n,
                        np.complex128, False, 2, cap)


def run_c4(env: Env, L=30, cap=0):
    wl, pkg = env.wl, env.pkg
    op = pkg.Operator.xxz(env.ctx, L)
    n = op.n_global
    row0, nl = wl.partition(n, env.rank, env.world)
    rs = np.random.RandomState(1 + env.rank)  # start vector seeded per row block (no 155M-entry global vector per rank)
    start = rs.uniform(-1, 1, nl)
    return lanczos_case(env, f"config4: XXZ chain L={L} Sz=0 (dim {n}) matrix-free, double, ground state" +
                        (f", max_iteration={cap}" if cap else ", to convergence"), op, n, np.float64, False, 1, cap or None,
                        start=start, extra={"note": "start vector seeded per row block"})


def neel_index(L):
    neel = sum(1 << b for b in range(0, L, 2))
    idx, j = 0, 0
    for b in range(L):
        if neel >> b & 1:
            j += 1
            idx += comb(b, j)
    return idx


def run_c5(env: Env, L=28, steps=100):
    """Device-resident time evolution: the output vector of a step is the input of the next (llz_expm_run, host = 0)."""
    import ctypes as C

    wl, pkg = env.wl, env.pkg
    op = pkg.Operator.xxz(env.ctx, L, dtype=np.complex128)
    n = op.n_global
    row0, nl = wl.partition(n, env.rank, env.world)
    idx = neel_index(L)
    psi = np.zeros(nl, dtype=np.complex128)
    if row0 <= idx < row0 + nl:
        psi[idx - row0] = 1
    lib = pkg.lib()

    def evolve(nsteps):
        vin, vout = pkg.Vector.from_host(env.ctx, psi), pkg.Vector(env.ctx, np.complex128, nl)
        din, dout = C.c_void_p(), C.c_void_p()
        lib.llz_vec_device_ptr(vin.h, C.byref(din))
        lib.llz_vec_device_ptr(vout.h, C.byref(dout))
        its = []
        env.sync()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            it = C.c_int64(0)
            st = lib.llz_expm_run(env.ctx.h, op.h, C.c_int(3), (C.c_double * 2)(0.0, -0.1), din, dout, C.c_int(0),
                                  C.c_double(-1.0), C.c_int(0), C.c_int64(0), C.c_int(0), C.byref(it))
            assert st == 0, lib.llz_last_error()
            its.append(int(it.value))
            din, dout = dout, din
        env.sync()
        dt = time.perf_counter() - t0
        fin = vin if nsteps % 2 == 0 else vout
        return its, dt, fin.norm()

    evolve(3)  # warm-up: maps the basis memory
    its, dt, nrm = evolve(steps)
    total = sum((op.bytes() + 8 * nl * 16) * m + (m + 1) * nl * 16 for m in its)
    return {"config": f"config5: Exponentiator exp(-i H 0.1) on XXZ L={L} (dim {n}) complex128, {steps} steps from the Neel state, "
                      "device-resident", "n": int(n), "n_gpus": env.world, "iterations_per_step": sorted(set(its)),
            "iterations": sum(its), "seconds": dt, "steps_per_s": steps / dt, "iterations_per_s": sum(its) / dt,
            "time_for_all_steps_s": dt, "final_norm": nrm, "model_GBps_per_gpu": total / dt / 1e9,
            "frac_of_measured_peak": total / dt / 1e9 / env.peak, "frac_of_8TBps": total / dt / 1e9 / 8000.0}
