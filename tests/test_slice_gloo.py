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
    """CPU stand-in for GpuSliceBackend: same contract, distances from the oracle (FP32 maps, like the device)."""

    def __init__(self, ti_crop, tilesize, disabled_crop, auxti_crop=()):
        self.ti = np.asarray(ti_crop, dtype=np.float64)
        self.aux = [np.asarray(a, dtype=np.float64) for a in auxti_crop]
        self.dis = None if disabled_crop is None else np.asarray(disabled_crop).astype(bool)

    def _map(self, img, kern, w):
        from oracle import iq_oracle as O
        D = O.fastdistance(img, np.asarray(kern, dtype=np.float64), w, method="direct")
        if self.dis is not None:
            D[self.dis] = np.inf
        return np.ascontiguousarray(D.astype(np.float32).ravel(order="F"))

    def distance(self, mask, simdev, softdevs=()):
        self.maps = [self._map(self.ti, simdev, mask.astype(float))]
        self.maps += [self._map(a, sd, np.ones(mask.shape)) for a, sd in zip(self.aux, softdevs)]
        self.D = self.maps[0]
        return float(self.D.min())

    def select(self, tol, gmin):
        idx = np.flatnonzero(self.D.astype(np.float64) <= (1.0 + tol) * float(np.float32(gmin))).astype(np.int64)
        return idx, self.D[idx]

    # relaxation path: the contract of iq_slice_minmax / _hist / _kth / _pick
    def _keys(self, s):
        return (self.maps[s].view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(self.maps[s].size, dtype=np.uint64)

    def minmax(self):
        en = slice(None) if self.dis is None else ~self.dis.ravel(order="F")
        lo = np.array([m[en].min() if m[en].size else np.inf for m in self.maps], dtype=np.float32).view(np.uint32)
        hi = np.array([m[en].max() if m[en].size else 0.0 for m in self.maps], dtype=np.float32).view(np.uint32)
        return lo, hi

    def hist(self, reqs):
        out = np.zeros((len(reqs), 256), dtype=np.int64)
        for i, (s, level, prefix) in enumerate(reqs):
            bits = self.maps[s].view(np.uint32).astype(np.uint64)
            sel = np.ones(bits.size, bool) if level == 0 else (bits >> np.uint64(32 - 8 * level)) == np.uint64(prefix)
            out[i] = np.bincount(((bits[sel] >> np.uint64(24 - 8 * level)) & np.uint64(255)).astype(np.int64), minlength=256)
        return out

    def kth(self, src, k_local):
        return int(np.sort(self._keys(src))[k_local - 1])

    def pick(self, thr):
        ok = np.ones(self.maps[0].size, bool)
        for s, t in enumerate(thr):
            ok &= self._keys(s) <= np.uint64(t)
        idx = np.flatnonzero(ok).astype(np.int64)
        return idx, np.stack([self.maps[s][idx] for s in range(len(thr))])


def _soft_pair(ti):
    """Integer-valued auxiliary variable and auxiliary TI (exact in FP32: the CPU stand-in, the oracle's FP64 maps and
    the device agree bit for bit; many ties, which is what the position tie-break of the distributed select needs)."""
    r = np.random.default_rng(11)
    auxti = np.asfortranarray(np.round(np.nan_to_num(ti) * 2 + r.integers(0, 2, ti.shape)).astype(np.float64))
    aux = np.asfortranarray(r.integers(0, 6, ti.shape).astype(np.float64))
    return aux, auxti


def _worker(rank, world, port, ti, use_gpu, q, soft=False):
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
    factory = None if use_gpu else (lambda t, ts, d, a=(): OracleSliceBackend(t, ts, d, a))
    reals = sharding.iqsim_sliced(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2, seed=3,
                                  soft=[_soft_pair(ti)] if soft else (), device=device, backend_factory=factory)
    if rank == 0:
        q.put([np.asarray(r) for r in reals])
    dist.barrier()
    dist.destroy_process_group()



def _wait_result(q, procs, timeout):
    """Result of rank 0, failing fast (instead of waiting out the timeout) when a rank has died."""
    import queue
    import time
    t0 = time.time()
    while True:
        try:
            return q.get(timeout=1.0)
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError("worker exit codes %r" % dead if dead else "no result within %d s" % timeout)


def _run(use_gpu, soft=False):
    r = np.random.default_rng(1)
    ti = np.asfortranarray(r.integers(0, 3, (20, 15, 9)).astype(np.float64))
    ti[3, 4, 2] = np.nan
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rank, 2, port, ti, use_gpu, q, soft)) for rank in range(2)]
    for p in procs:
        p.start()
    got = _wait_result(q, procs, 280)
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


def test_select_over_slabs_equals_global_partial_sort():
    """The distributed radix select (all-gathered digit histograms + position tie-break by slab order) returns, on every
    rank, a local threshold that admits exactly the local members of the global k smallest (value, position) keys."""
    sys.path.insert(0, ROOT)
    import iqb200  # noqa: F401
    from iqb200 import sharding
    r = np.random.default_rng(5)
    for trial in range(12):
        world = int(r.integers(1, 5))
        n = int(r.integers(40, 400))
        # few distinct values -> ties across slabs; some +Inf (disabled) entries; one trial all zeros
        vals = (r.integers(0, 4 if trial % 2 else 50, n) * (0.0 if trial == 3 else 1.0)).astype(np.float32)
        vals[r.random(n) < 0.1] = np.inf
        cuts = np.sort(r.integers(0, n + 1, world - 1)) if world > 1 else np.zeros(0, int)
        bounds = [0] + [int(c) for c in cuts] + [n]
        backs = []
        for w in range(world):
            b = OracleSliceBackend(np.zeros((1, 1)), (1, 1), None)
            b.maps = [np.ascontiguousarray(vals[bounds[w]:bounds[w + 1]])]
            backs.append(b if b.maps[0].size else None)
        gkeys = (vals.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        nfinite = int(np.isfinite(vals).sum())
        for k in sorted({1, 2, max(1, nfinite // 10), max(1, nfinite // 2), max(1, nfinite)}):
            want = np.zeros(n, bool)
            want[np.argsort(gkeys, kind="stable")[:k]] = True
            for w in range(world):
                last = {}

                def gather(local, w=w):
                    # what the all-gather would deliver: every rank's histograms of the same requests
                    out = []
                    for x, bk in enumerate(backs):
                        h = bk.hist(last["reqs"]) if bk is not None else np.zeros((len(last["reqs"]), 256), np.int64)
                        if local.shape[0] == h.shape[0] + 1:
                            h = np.concatenate([h, np.zeros((1, 256), np.int64)])
                        out.append(h)
                    return np.stack(out)

                class Spy:
                    def hist(self_, reqs):
                        last["reqs"] = reqs
                        return backs[w].hist(reqs)

                    def kth(self_, src, kl):
                        return backs[w].kth(src, kl)

                if backs[w] is None:
                    continue
                # ranks without positions still take part in the exchange; here only the populated ones are checked
                thr, _ = sharding.select_over_slabs(Spy(), w, world, [0], [k], gather)
                lk = backs[w]._keys(0)
                got = np.zeros(lk.size, bool) if thr[0] is None else lk <= np.uint64(thr[0])
                assert np.array_equal(got, want[bounds[w]:bounds[w + 1]]), (trial, world, k, w)


@pytest.mark.timeout(300)
def test_two_rank_position_slices_with_soft_data_equal_oracle_iqsim():
    """Relaxation path over slabs (all-gathered radix histograms) == the oracle's single-process iqsim."""
    from oracle import iq_oracle as O
    ti, got = _run(use_gpu=False, soft=True)
    want = O.iqsim(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2, soft=[_soft_pair(ti)],
                   rng=np.random.default_rng(3), method="direct", cut_fn=O.graphcut_c)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_two_rank_position_slices_with_soft_data_equal_single_gpu_iqsim():
    import iqb200
    ti, got = _run(use_gpu=True, soft=True)
    want = iqb200.iqsim(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2, soft=[_soft_pair(ti)],
                        rng=np.random.default_rng(3), cut="host")
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


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_two_rank_position_slices_with_soft_data_over_nccl_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import iqb200
    ti, got = _run(use_gpu="nccl", soft=True)
    want = iqb200.iqsim(ti, (8, 6, 4), None, overlap=(0.25, 0.34, 0.5), tol=0.1, path="random", nreal=2, soft=[_soft_pair(ti)],
                        rng=np.random.default_rng(3))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
