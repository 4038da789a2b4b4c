"""Row-sharded (multi-GPU) path.

CPU (`-m "not gpu"`): world_size-2 gloo run of the host-side sharding logic (partition, halo planning through the C ABI,
the request/pack/exchange protocol) against the unsharded product.
GPU (`-m gpu`, needs >= 2 devices, else skipped): the CUDA path on 2 ranks over NCCL against an un-joined single-GPU
context — tests/mgpu_worker.py holds the checks.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def launch(mode, nproc, port, timeout, extra_env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mgpu_worker.py"), "--mode", mode]
    env = dict(os.environ, OMP_NUM_THREADS="1", **(extra_env or {}))
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_host_logic_gloo(pkg, world):
    r = launch("cpu", world, 29611 + world, 600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MGPU_OK") == world


def test_partition_covers_everything(pkg, wl):
    import ctypes as C

    lib = pkg.lib()
    for n in (1, 7, 1000, 155117520, 16777216):
        for G in (1, 2, 3, 8):
            prev = 0
            for r in range(G):
                a, b = C.c_int64(0), C.c_int64(0)
                assert lib.llz_partition(C.c_int64(n), r, G, C.byref(a), C.byref(b)) == 0
                assert a.value == prev and (a.value, b.value) == wl.partition(n, r, G)
                prev = a.value + b.value
            assert prev == n
    a, b = C.c_int64(0), C.c_int64(0)
    assert lib.llz_partition(C.c_int64(10), 3, 3, C.byref(a), C.byref(b)) == 1  # rank out of range


def test_halo_plan_rejects_bad_columns(pkg):
    import ctypes as C

    lib = pkg.lib()
    rowptr = np.array([0, 2], dtype=np.int64)
    colidx = np.array([0, 9], dtype=np.int32)
    bounds = np.array([0, 1, 4], dtype=np.int64)
    n_halo = C.c_int64(0)
    st = lib.llz_halo_plan(C.c_int64(1), C.c_int64(0), rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p), 2,
                           bounds.ctypes.data_as(C.c_void_p), None, None, C.c_int64(0), C.byref(n_halo), None)
    assert st == 1 and b"outside" in lib.llz_last_error()


def _gpu_count():
    # (no `import torch` here: this process already holds libllz.so; count the devices out of process)
    try:
        return len([l for l in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout.splitlines() if l.startswith("GPU ")])
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_cuda_path(world):
    """CSR, SELL and matrix-free XXZ operators, LambdaLanczos and the Exponentiator on `world` row blocks: applies
    bit-identical to the single-GPU apply, runs within the north-star tolerances of the single-GPU run AND of the oracle
    (tests/mgpu_worker.py).  Every GPU count the box offers is exercised (gpurun --gpus N)."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs (run with gpurun --gpus {world})")
    for p2p in ("1", "0"):  # peer-memory channels, then the NCCL-only path
        r = launch("gpu", world, 29631 + int(p2p) + 2 * world, 420, {"LLZ_P2P": p2p})
        err = r.stderr
        if "Traceback" in err:
            err = err[err.index("Traceback"):]
        assert r.returncode == 0, r.stdout[-2000:] + err[:3000]
        assert r.stdout.count("MGPU_OK") == world
