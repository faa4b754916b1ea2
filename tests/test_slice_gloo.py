"""Position-slice mode (sharding.iqsim_sliced): world_size-2 gloo test on CPU with an oracle-backed slab
search, and (gpu-marked) the same with two ranks sharing cuda:0.  Both must reproduce the single-process
run bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleSliceBackend:
    """CPU stand-in for GpuSliceBackend: same contract, distances from the oracle."""

    def __init__(self, ti_crop, tilesize, disabled_crop):
        self.ti = np.asarray(ti_crop, dtype=np.float64)
        self.dis = None if disabled_crop is None else np.asarray(disabled_crop).astype(bool)

    def distance(self, mask, simdev):
        from oracle import iq_oracle as O
        D = O.fastdistance(self.ti, np.asarray(simdev, dtype=np.float64), mask.astype(float), method="direct")
        if self.dis is not None:
            D[self.dis] = np.inf
        self.D = D.astype(np.float32).ravel(order="F")
        return float(self.D.min())

    def select(self, tol, gmin):
        idx = np.flatnonzero(self.D.astype(np.float64) <= (1.0 + tol) * float(np.float32(gmin))).astype(np.int64)
        return idx, self.D[idx]


def _worker(rank, world, port, ti, use_gpu, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # two physical GPUs: one rank per GPU over NCCL (the candidate exchange runs on device tensors); otherwise gloo,
    # both ranks sharing cuda:0 (or no GPU at all with the oracle-backed slab search)
    nccl = use_gpu == "nccl"
    device = rank if nccl else 0
    if nccl:
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import iqb200  # noqa: F401
    from iqb200 import sharding
    factory = None if use_gpu else (lambda t, ts, d: OracleSliceBackend(t, ts, d))
    reals = sharding.iqsim_sliced(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2, seed=3,
                                  device=device, backend_factory=factory)
    if rank == 0:
        q.put([np.asarray(r) for r in reals])
    dist.barrier()
    dist.destroy_process_group()


def _run(use_gpu):
    r = np.random.default_rng(1)
    ti = np.asfortranarray(r.integers(0, 3, (20, 15, 9)).astype(np.float64))
    ti[3, 4, 2] = np.nan
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rank, 2, port, ti, use_gpu, q)) for rank in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=280)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return ti, got


def test_slab_split():
    sys.path.insert(0, ROOT)
    import iqb200  # noqa: F401
    from iqb200 import sharding
    for n in (1, 6, 85):
        for w in (1, 2, 8):
            b = [sharding.slab(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))


@pytest.mark.timeout(300)
def test_two_rank_position_slices_equal_oracle_iqsim():
    from oracle import iq_oracle as O
    import iqb200
    ti, got = _run(use_gpu=False)
    want = O.iqsim(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2,
                   rng=np.random.default_rng(3), method="direct", cut_fn=O.graphcut_c)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_two_rank_position_slices_equal_single_gpu_iqsim():
    import iqb200
    ti, got = _run(use_gpu=True)
    want = iqb200.iqsim(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2,
                        rng=np.random.default_rng(3), cut="host")
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_two_rank_position_slices_over_nccl_on_two_gpus():
    """One rank per GPU, NCCL: all-reduce(min) of the local minima and the tensor all-gather of the candidate records
    run on device tensors.  Needs two physical GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import iqb200
    ti, got = _run(use_gpu="nccl")
    want = iqb200.iqsim(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2,
                        rng=np.random.default_rng(3))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
