"""world_size-2 gloo tests of the multi-GPU partitioning logic (the N>1 path of bench.py / distributed.py).
The path shards independent hyper-parameter sets; the only collective is the all-gather of results."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from gpyrn_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_cover_exactly_once():
    for B in (1, 7, 8, 8192, 8191):
        for W in (1, 2, 4, 8):
            for mode in ("strided", "block"):
                got = np.concatenate([D.shard_indices(B, W, r, mode) for r in range(W)])
                assert sorted(got.tolist()) == list(range(B))
                sizes = [len(D.shard_indices(B, W, r, mode)) for r in range(W)]
                assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_indices(8, 2, 2)


def test_balanced_assignment():
    rng = np.random.default_rng(0)
    for B, W in ((16, 8), (17, 4), (8192, 8), (3, 4)):
        costs = rng.integers(4, 90, B)
        parts = D.balanced_assignment(costs, W)
        assert sorted(np.concatenate(parts).tolist()) == list(range(B))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
        sums = [costs[p].sum() for p in parts]
        naive = [costs[D.shard_indices(B, W, r)].sum() for r in range(W)]
        if B >= 2 * W:
            assert max(sums) - min(sums) <= max(naive) - min(naive) + costs.max()


def test_balanced_assignment_with_lockstep_cost_model():
    """Blocks evaluated in lock-step: a block pays step_times[n-1] per iteration while n of its sets are active.
    The refinement must keep the partition valid and not increase the largest modelled block cost."""
    st = [24.65, 40.15, 56.8, 74.25]
    assert D.lockstep_cost([10, 10, 10, 10], st) == pytest.approx(10 * st[3])
    assert D.lockstep_cost([5, 20], st) == pytest.approx(5 * st[1] + 15 * st[0])
    rng = np.random.default_rng(1)
    for trial in range(5):
        costs = rng.integers(30, 100, 32)
        base = D.balanced_assignment(costs, 8)
        ref = D.balanced_assignment(costs, 8, step_times=st)
        assert sorted(np.concatenate(ref).tolist()) == list(range(32)) and all(len(b) == 4 for b in ref)
        worst = lambda blocks: max(D.lockstep_cost(costs[b], st) for b in blocks)
        assert worst(ref) <= worst(base) + 1e-9
    rec = [44, 54, 40, 55, 42, 60, 40, 62, 44, 43, 53, 58, 38, 51, 41, 34, 53, 50, 67, 77, 53, 69, 79, 40, 67, 42, 71,
           63, 94, 68, 53, 73]                      # profiles/iterations_c4.json
    c = [D.lockstep_cost(np.array(rec)[b], st) for b in D.balanced_assignment(rec, 8, step_times=st)]
    assert max(c) / min(c) < 1.03                   # 1.12 with the iteration sums alone


def test_shared_counter_single_process():
    c = D.SharedCounter(3, "solo")
    assert [c.next(), c.next(), c(), c.next(), c.next()] == [0, 1, 2, -1, -1]
    assert c.handed == [0, 1, 2]
    c = D.SharedCounter(11, "grains", first=4, grain=3)       # reservations: [0..4) [4..7) [7..10) [10..11)
    assert [c.next() for _ in range(13)] == list(range(11)) + [-1, -1] and c.trips >= 4
    assert D.dealing_grains(16, 2, 0) == (8, 1) and D.dealing_grains(64, 8, 4) == (4, 1)
    assert D.dealing_grains(16384, 2, 13107) == (8192, 64)
    P = np.arange(12.0).reshape(4, 3)

    def evaluate(PP, counter):
        taken = np.zeros(len(PP), bool)
        while (i := counter.next()) >= 0:
            taken[i] = True
        return PP.sum(axis=1), np.full(len(PP), 7, np.int32), np.zeros(len(PP), np.int32), taken

    e, it, st, owner = D.elbo_pool_sharded(None, P, evaluate=evaluate)
    assert np.allclose(e, P.sum(axis=1)) and np.all(it == 7) and np.all(owner == 0)


def test_gather_single_process():
    idx = np.array([0, 2, 4])
    out = D.gather_results(idx, {"elbo": np.array([1.0, 2.0, 3.0])}, 5)
    assert np.allclose(out["elbo"], [1, 0, 2, 0, 3])


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = D.shard_indices(B, world, rank, "strided")
    vals = {"elbo": -1000.0 - idx.astype(float), "iters": (idx % 7 + 4).astype(np.int64)}
    out = D.gather_results(idx, vals, B)
    ok = np.allclose(out["elbo"], -1000.0 - np.arange(B)) and np.array_equal(out["iters"], np.arange(B) % 7 + 4)

    class Fake:   # stands in for an inference bound to this rank's GPU
        def ELBO_batch(self, P, max_iter=None, return_info=False):
            e = P.sum(axis=1)
            return e, np.full(len(e), 5, np.int32), np.zeros(len(e), np.int32)

    P = np.arange(B * 3, dtype=float).reshape(B, 3)
    e, it, st = D.elbo_batch_sharded(Fake(), P, mode="block")
    ok = ok and np.allclose(e, P.sum(axis=1)) and np.all(it == 5) and np.all(st == 0)
    # dynamic dealing: both ranks pull set indices from ONE shared counter (TCPStore add), a slow and a fast rank;
    # every set is evaluated exactly once, the all-reduce hands every rank the complete result
    import time as _time

    def evaluate(PP, counter):
        e = np.zeros(len(PP)); it = np.zeros(len(PP), np.int32); st = np.zeros(len(PP), np.int32)
        taken = np.zeros(len(PP), bool)
        while True:
            i = counter.next()
            if i < 0:
                break
            _time.sleep(0.002 * (1 + 4 * rank))           # rank 1 is five times slower
            e[i], it[i], taken[i] = PP[i].sum(), 4 + i % 3, True
        return e, it, st, taken

    e, it, st, owner = D.elbo_pool_sharded(None, P, key="t1", evaluate=evaluate, slots=4)     # 4 in flight: dynamic tail
    ok = ok and np.allclose(e, P.sum(axis=1)) and np.array_equal(it, 4 + np.arange(B) % 3) and np.all(st == 0)
    ok = ok and set(np.unique(owner).tolist()) <= {0, 1} and np.all(owner >= 0)
    if B >= 32:
        ok = ok and (owner == 0).sum() > (owner == 1).sum()       # the fast rank took more sets
    # a second pool must start from a fresh counter
    e2, _, _, owner2 = D.elbo_pool_sharded(None, P[::-1].copy(), key="t2", evaluate=evaluate)
    ok = ok and np.allclose(e2, P[::-1].sum(axis=1))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 64])
def test_gloo_world2_gather(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
