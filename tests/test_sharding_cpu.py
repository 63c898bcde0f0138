"""N>1 host logic on CPU: world_size-2 gloo group (what bench.py does under torchrun with nccl)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ksim_gym_zbot_b200  # noqa: F401
from ksim_gym_zbot_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(4097, world, rank)
    # each rank "processes" its block; rank 1 is slower -> the job time is the max
    value, tmax = sharding.aggregate_throughput((hi - lo) * 4 * 10, 100.0 * (1 + rank))
    # the oracle is deterministic per environment: sharded == unsharded on the same global env ids
    import zb_oracle
    from helpers import random_actions, random_states
    from ksim_gym_zbot_b200 import model as M

    m = M.build_zbot_syn()
    n = 10
    qg, vg = random_states(m, n, 5)
    ag = random_actions(n, 6)
    slo, shi = sharding.shard_range(n, world, rank)
    o = zb_oracle.Oracle(m, M.make_opts(dt=0.005), np.float32)
    st = o.new_state(shi - slo)
    o.reset(st, qg[slo:shi], vg[slo:shi], np.full(shi - slo, 1e9))
    o.step(st, ag[slo:shi], 2)
    parts = [None] * world
    dist.all_gather_object(parts, (slo, shi, st["qpos"]))
    q.put((rank, lo, hi, value, tmax, parts, sharding.rank_seed(3, rank)))
    dist.destroy_process_group()


def test_shard_range_partition():
    for n, w in [(4096, 1), (4097, 2), (65536, 8), (5, 8)]:
        blocks = [sharding.shard_range(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_aggregate_single_rank():
    v, t = sharding.aggregate_throughput(1000.0, 500.0)
    assert v == 2000.0 and t == 500.0


def test_world2_gloo(zmodel):
    import zb_oracle
    from helpers import random_actions, random_states
    from ksim_gym_zbot_b200 import model as M

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, v0, t0, parts0, s0), (r1, lo1, hi1, v1, t1, parts1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 2049, 2049, 4097)
    assert t0 == t1 == 200.0 and v0 == v1 == 4097 * 40 / 0.2  # max over ranks, sum of units
    assert s0 != s1
    # sharded result == single-process result on the same global environments
    n = 10
    qg, vg = random_states(zmodel, n, 5)
    o = zb_oracle.Oracle(zmodel, M.make_opts(dt=0.005), np.float32)
    st = o.new_state(n)
    o.reset(st, qg, vg, np.full(n, 1e9))
    o.step(st, random_actions(n, 6), 2)
    got = np.concatenate([p[2] for p in sorted(parts0)])
    np.testing.assert_array_equal(got, st["qpos"])
