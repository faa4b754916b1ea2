"""torchrun --nproc-per-node N scripts/sharded_check.py : realization sharding over N GPUs (NCCL gather) equals the
single-process run bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import iqb200
from iqb200 import sharding, synth

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ti = synth.gaussian_field((64, 56, 24), (6, 6, 3), 4)
kw = dict(overlap=(0.25, 0.25, 0.25), path="random", device=local)
got = sharding.iqsim_sharded(ti, (16, 14, 8), None, nreal=7, seed=21, **kw)
if rank == 0:
    want = iqb200.iqsim(ti, (16, 14, 8), nreal=7, rng=np.random.default_rng(21), **kw)
    ok = len(got) == 7 and all(np.array_equal(a, b) for a, b in zip(got, want))
    print("sharded == single process:", ok, "ranks", dist.get_world_size(), flush=True)
    assert ok
dist.barrier()
dist.destroy_process_group()
