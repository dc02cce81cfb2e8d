"""Multi-rank tests of the shard/exchange layer (goetia_b200/shard.py).

CPU: world_size-2 and -3 gloo runs check the plan, the outbox/inbox layout and the two all-to-alls
against the oracle's tables (kernels emulated in numpy).  GPU (needs >= 2 devices): the same check
with the real kernels over NCCL.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "shard_worker.py")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(mode, kind, world, extra_env=None, timeout=300):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, WORKER, mode, str(kind)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    return [p.returncode for p in procs], outs


def test_plan_partitions_every_table():
    from goetia_b200.shard import ShardPlan
    from oracle.binding import Port
    for kind in (0, 1, 2):
        for world in (1, 2, 3, 8):
            sizes = Port.primes_near(4, 8_000_000)
            plan = ShardPlan(kind, sizes, world, 1_000_000, 14)
            assert plan.nb <= 1024
            for t, size in enumerate(sizes):
                bs = np.nonzero(plan.table == t)[0]
                # slices tile the table exactly, in order
                assert int(plan.slot0[bs[0]]) == 0
                assert np.array_equal(plan.slot0[bs][1:], (plan.slot0[bs] + plan.slots[bs])[:-1])
                assert int(plan.slot0[bs[-1]] + plan.slots[bs[-1]]) == size
                # owners are contiguous runs in rank order and match own_lo / own_hi
                assert np.all(np.diff(plan.owner[bs]) >= 0)
                for r in range(world):
                    mine = bs[plan.owner[bs] == r]
                    if mine.size:
                        assert int(plan.slot0[mine[0]]) == int(plan.own_lo[r, t])
                        assert int(plan.slot0[mine[-1]] + plan.slots[mine[-1]]) == int(plan.own_hi[r, t])
                    else:
                        assert int(plan.own_lo[r, t]) == int(plan.own_hi[r, t])
                assert int(plan.own_hi[world - 1, t]) == size
                # no 32-bit table word is shared between two ranks
                assert np.all(plan.own_lo[:, t] % 32 == 0)
            # capacities cover the expected share with slack
            assert np.all(plan.cap >= 1_000_000 * plan.slots / np.asarray(sizes, dtype=np.float64)[plan.table])


def test_plan_of_the_headline_workload():
    """C3 (BitStorage, 4 x 8e9 bits) at 1 / 2 / 4 / 8 ranks: the bucket count stays at 120 (k_bucket's sweet spot),
    every rank gets at least 3 slices of every table, the heaviest rank holds at most 4/30 of a table, and the peer
    layout's inboxes hold `world` regions plus the overflow lists."""
    from goetia_b200.shard import ShardPlan
    from oracle.binding import Port
    sizes = Port.primes_near(4, int(8e9))
    for world in (1, 2, 4, 8):
        plan = ShardPlan(0, sizes, world, 900_000_000 // max(1, world // 2), 0)
        assert plan.nb == 120 and plan.shift == 28
        for t, size in enumerate(sizes):
            per_rank = [(plan.owner[plan.table == t] == r).sum() for r in range(world)]
            assert min(per_rank) >= 2 and max(per_rank) <= -(-30 // world)
            assert sum(int(plan.own_hi[r, t] - plan.own_lo[r, t]) for r in range(world)) == size
        lay = plan.peer_layout()
        for q in range(world):
            assert int(lay["region"][q]) == int(plan.cap[plan.owned[q]].sum())
            assert int(lay["inbox_bytes"][q]) >= world * int(lay["region"][q]) * 4
        assert sorted(plan.fill_perm().tolist()) == list(range(plan.nb + world))


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("world", [2, 3])
def test_exchange_layout_gloo(kind, world):
    rcs, outs = launch("cpu", kind, world)
    assert rcs == [0] * world, "\n".join(outs)
    assert "tables bit-exact" in outs[0]


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("world", [2, 3])
def test_peer_layout_gloo(kind, world):
    """The peer transport's inbox layout, fill exchange and overflow lists on CPU (stores emulated by all-to-all)."""
    rcs, outs = launch("cpu_p2p", kind, world)
    assert rcs == [0] * world, "\n".join(outs)
    assert "tables bit-exact" in outs[0]


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("world", [2, 3])
def test_owner_routed_requests_gloo(kind, world):
    """ShardRouter (route / ship / back) on CPU: routed queries and the first-toucher flags of a tracked insert equal the
    oracle's serial answers; a request that reaches a rank not holding its slot fails the run."""
    rcs, outs = launch("cpu_route", kind, world, {"SHARD_READS": "120"})
    assert rcs == [0] * world, "\n".join(outs)
    assert "routes exact" in outs[0]


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["p2p", "ce", "nccl"])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_sharded_insert_nccl(kind, transport):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    rcs, outs = launch("cuda", kind, world, {"SHARD_TABLE_X": "40000000", "SHARD_READS": "40000", "SHARD_SLICE_LOG2": "16",
                                             "SHARD_ROUNDS": "3", "SHARD_TRANSPORT": transport})
    assert rcs == [0] * world, "\n".join(outs)
    assert "tables bit-exact" in outs[0]


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["p2p", "ce"])
@pytest.mark.parametrize("kind", [0, 1])
def test_sharded_insert_foreign_bucket_overflow(kind, transport):
    """Heavily duplicated reads with capacities sized for the uniform share: buckets of foreign slices overflow
    and the excess travels through the overflow lists (bucket.cuh, post_foreign) -- tables stay bit-exact."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    rcs, outs = launch("cuda", kind, 2, {"SHARD_TABLE_X": "40000000", "SHARD_READS": "40000", "SHARD_SLICE_LOG2": "16",
                                         "SHARD_ROUNDS": "2", "SHARD_TRANSPORT": transport, "SHARD_BUDGET_X": "1"})
    assert rcs == [0, 0], "\n".join(outs)
    assert "tables bit-exact" in outs[0]
