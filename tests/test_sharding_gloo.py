"""world_size-2 gloo test (CPU) of the realization sharding: the union of the shards must equal the
single-process run bit for bit.  The per-rank simulation is the CPU oracle here (no GPU in this tier);
the GPU path uses the same split/stream logic (imagequilting.jl_b200/sharding.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_run(trainimg, tilesize, simsize, nreal, rng, _real_range, **kw):
    """Oracle with the product's sharding contract: draw the whole stream, simulate rows r0:r1."""
    from oracle import iq_oracle as O
    r0, r1 = _real_range
    geo = O.geometry(np.shape(trainimg), tilesize, simsize, kw.get("overlap"))
    skipped, datainds = O.findskipped(kw.get("hard") or {}, geo)
    path = O.genpath(rng, geo["ntiles"], kw.get("path", "raster"), datainds)
    nvis = len([p for p in path if p not in skipped])
    u = rng.random(nreal * nvis).reshape(nreal, nvis)

    class Replay:  # hands the pre-drawn rows to the oracle in its own draw order
        def __init__(self, rows):
            self.vals = list(rows.ravel())
            self.i = 0

        def random(self):
            v = self.vals[self.i]
            self.i += 1
            return v

        def permutation(self, n):
            raise AssertionError("path is replayed")

    class FixedPath(Replay):
        pass

    rep = Replay(u[r0:r1])
    orig_genpath = O.genpath
    O.genpath = lambda rng_, extent, kind, di: list(path)
    try:
        return O.iqsim(trainimg, tilesize, simsize, nreal=r1 - r0, rng=rep, **kw)
    finally:
        O.genpath = orig_genpath


def _worker(rank, world, port, ti, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import iqb200  # noqa: F401
    from iqb200 import sharding
    reals = sharding.iqsim_sharded(ti, (8, 8), None, nreal=5, seed=11, run_fn=_oracle_run, path="random")
    if rank == 0:
        q.put([np.asarray(r) for r in reals])
    dist.barrier()
    dist.destroy_process_group()


def test_split_covers_everything():
    sys.path.insert(0, ROOT)
    import iqb200  # noqa: F401
    from iqb200 import sharding
    for nreal in (1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            blocks = [sharding.split(nreal, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == nreal
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))



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


@pytest.mark.timeout(300)
def test_two_rank_gloo_equals_single_process():
    from oracle import iq_oracle as O
    r = np.random.default_rng(0)
    ti = np.asfortranarray(r.integers(0, 3, (24, 20)).astype(np.float64))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rank, 2, port, ti, q)) for rank in range(2)]
    for p in procs:
        p.start()
    got = _wait_result(q, procs, 240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = O.iqsim(ti, (8, 8), None, nreal=5, rng=np.random.default_rng(11), path="random")
    assert len(got) == 5
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
