"""World-size-2 run of the multi-rank host logic on CPU (gloo): every rank regenerates ITS shard of
the counter-based state stream, evaluates it (here with the host emulation of the device pipeline,
since this box has no GPU), and the gathered result equals the single-process result on the union;
the bench's max-over-ranks timing reduction is exercised too."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util

sys.path.insert(0, os.path.join(util.ROOT, "tests", "hostemu"))
import emu  # noqa: E402

pytestmark = pytest.mark.skipif(not emu.available(), reason="host emulation library not buildable")

TOTAL = 203      # deliberately not divisible by the world size


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.shard import shard_range
    from mujoco_inversedynamicstest_b200.states import generate_states
    model = mjb.Model.from_mjb(util.golden("humanoid")[0])
    first, count = shard_range(TOTAL, rank, world)
    qpos, qvel, qacc = generate_states(model, count, first=first)
    out = emu.run(model, qpos, qvel, qacc)
    mine = torch.from_numpy(out["qfrc_inverse"])
    sizes = [shard_range(TOTAL, r, world)[1] for r in range(world)]
    gathered = [torch.empty((s, mine.shape[1]), dtype=torch.float64) for s in sizes]
    dist.all_gather(gathered, mine) if len(set(sizes)) == 1 else _gather_uneven(gathered, mine, rank, world)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)       # bench.py: max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.cat(gathered).numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.destroy_process_group()


def _gather_uneven(gathered, mine, rank, world):
    for r in range(world):
        buf = mine.clone() if r == rank else gathered[r]
        dist.broadcast(buf, src=r)
        gathered[r].copy_(buf)


def test_two_ranks_cover_the_batch_exactly(tmp_path):
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.shard import shard_range, weak_shard
    from mujoco_inversedynamicstest_b200.states import generate_states
    # partition properties
    for total in (0, 1, 7, 203, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
    assert weak_shard(1 << 20, 3) == (3 << 20, 1 << 20)

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    assert np.load(tmp_path / "tmax.npy")[0] == 11.0

    model = mjb.Model.from_mjb(util.golden("humanoid")[0])
    qpos, qvel, qacc = generate_states(model, TOTAL)
    single = emu.run(model, qpos, qvel, qacc)["qfrc_inverse"]
    assert np.array_equal(gathered, single)
