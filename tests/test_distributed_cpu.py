"""CPU tests of the multi-GPU plumbing with the gloo backend, world_size 2: sharding arithmetic,
global chain ids, the accumulator all-reduce and the sample gather.  The sampling itself needs a GPU; here every rank
fabricates per-chain accumulators that depend only on the GLOBAL chain id, so the reduced totals must equal the
single-process ones whatever the sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rwm_pt_pytorch_b200 import distributed as D


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 1025, 16384):
        for world in (1, 2, 3, 4, 8):
            blocks = [D.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    assert D.chain_id_base(5, 8) == 40
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)


def _fake_stats(first_chain, n_chains, K):
    ids = np.arange(first_chain, first_chain + n_chains, dtype=np.float64)
    return {"accept_count": float((ids % 7).sum()), "chain_steps": float(n_chains * 100),
            "sq_jump_sum": float((ids * 0.25).sum()), "swap_attempts": float(n_chains // K * 30),
            "swap_accepts": float((ids % 3).sum()), "sq_beta_jump_sum": float((ids % 5).sum() * 0.01)}


def _worker(rank, world, port, n_ladders, K, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = D.shard_range(n_ladders, rank, world)
        base = D.chain_id_base(start, K)
        total = D.allreduce_statistics(_fake_stats(base, count * K, K))
        counts = [D.shard_range(n_ladders, r, world)[1] for r in range(world)]
        local = torch.arange(start, start + count, dtype=torch.float32).view(-1, 1, 1).expand(count, 3, 2).contiguous()
        everyone = D.gather_samples(local, counts)
        only0 = D.gather_samples(local, counts, dst=0)
        q.put((rank, total, everyone[:, 0, 0].tolist(), None if only0 is None else only0[:, 2, 1].tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allreduce_and_gather():
    n_ladders, K, world = 11, 8, 2          # uneven shards: 6 + 5 ladders
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_ladders, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_stats(0, n_ladders * K, K)
    for rank, total, everyone, only0 in res:
        for k in D.STAT_KEYS:
            assert total[k] == pytest.approx(want[k], rel=1e-12), k
        assert everyone == [float(i) for i in range(n_ladders)]
        assert (only0 == [float(i) for i in range(n_ladders)]) if rank == 0 else (only0 is None)
    summ = D.pooled_summary(res[0][1])
    assert summ["acceptance_rate"] == pytest.approx(want["accept_count"] / want["chain_steps"])
    assert summ["swap_acceptance_rate"] == pytest.approx(want["swap_accepts"] / want["swap_attempts"])


def test_single_process_paths_are_identity():
    st = _fake_stats(0, 16, 8)
    assert D.allreduce_statistics(st) == {k: float(st[k]) for k in D.STAT_KEYS}
    x = torch.zeros(3, 2, 2)
    assert D.gather_samples(x, [3]) is x


class _FakeBatch:
    """CPU stand-in for the sampler's batch (only the fields the statistics helpers read)."""

    def __init__(self, L, K, seed):
        g = torch.Generator().manual_seed(seed)
        self.L, self.K, self.n_chains = L, K, L * K
        self.accept_count = torch.randint(0, 1000, (L * K,), generator=g, dtype=torch.int64)
        self.sq_jump_sum = torch.rand(L * K, generator=g, dtype=torch.float64) * 50
        self.swap_accepts = torch.randint(0, 40, (L, max(K - 1, 1)), generator=g, dtype=torch.int64)
        self.beta = torch.tensor([0.5 ** k for k in range(K)], dtype=torch.float32).repeat(L)

    def post_burn_in_steps(self):
        return 400

    def swap_rounds(self):
        return 40


class _FakeAlgo:
    def __init__(self, L, K, seed):
        self._batch = _FakeBatch(L, K, seed)


@pytest.mark.parametrize("K", [1, 8])
def test_device_side_statistics_match_the_host_side_ones(K):
    """The sync-free tensor form used inside bench.py's timed loop carries the same six accumulators as the dict form."""
    algo = _FakeAlgo(6, K, 3)
    want = D.local_statistics(algo)
    got = D.statistics_from_tensor(D.allreduce_statistics_tensor(D.local_statistics_tensor(algo)))
    for k in D.STAT_KEYS:
        assert got[k] == pytest.approx(want[k], rel=1e-12), k


class ParallelTemperingFake:
    """Records what make_sharded passes (class name contains 'ParallelTempering' like the real sampler)."""
    def __init__(self, dim, **kw):
        self.dim, self.kw = dim, kw


class RWMFake(ParallelTemperingFake):
    pass


RWMFake.__name__ = "RandomWalkFake"


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pt, shard_pt = D.make_sharded(ParallelTemperingFake, 11, 20, device="cpu", swap_every=10)
        rw, shard_rw = D.make_sharded(RWMFake, 4097, 20, device="cpu")
        q.put((rank, pt.kw, shard_pt, rw.kw, shard_rw))
    finally:
        dist.destroy_process_group()


def test_make_sharded_hands_every_rank_its_block_and_global_ids():
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, pt0, s0, rw0, t0), (_, pt1, s1, rw1, t1) = res
    assert (pt0["num_ladders"], pt0["ladder_id_base"], s0) == (6, 0, (0, 6)) and (pt1["num_ladders"], pt1["ladder_id_base"], s1) == (5, 6, (6, 5))
    assert (rw0["num_chains"], rw0["chain_id_base"]) == (2049, 0) and (rw1["num_chains"], rw1["chain_id_base"]) == (2048, 2049)
    assert pt0["swap_every"] == 10 and "chain_id_base" not in pt0 and "ladder_id_base" not in rw0
    # single process: the whole batch is one shard
    algo, shard = D.make_sharded(ParallelTemperingFake, 7, 3, device="cpu")
    assert shard == (0, 7) and algo.kw["num_ladders"] == 7 and algo.kw["ladder_id_base"] == 0
