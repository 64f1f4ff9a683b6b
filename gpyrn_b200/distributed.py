"""Sharding of independent ELBO evaluations over the GPUs of one box.

The path has no data-path exchange: every hyper-parameter set is an independent ``ELBOcalc``
(SURVEY.md 8e), the data (time, y, yerr) is replicated.  Two ways to split a pool of B sets over G ranks:

* dynamic (``elbo_pool_sharded``, the default of the benchmark): every rank keeps ``slots`` evaluations in flight on
  its GPU and takes the next set from ONE shared counter whenever a slot frees up (``SharedCounter``: an atomic
  fetch-add on the process group's rendezvous store -- no NCCL, no recorded costs).  Sets differ several-fold in
  iteration count, so this is what keeps the GPUs equally busy (the reference's pattern is a process pool over
  independent walkers, gpyrn/examples/example_4.py:66-68);
* static (``elbo_batch_sharded``): rank r evaluates a contiguous or strided block.

Either way the only collective is the reduction / gather of the B ELBO values (+ iteration counts) at the end -- NCCL
over NVLink when the ranks hold GPUs, gloo in the CPU tests of the partitioning logic.
"""
import itertools

import numpy as np

_pool_serial = itertools.count()


class SharedCounter:
    """Atomic work counter shared by all ranks of the process group: ``next()`` returns every index of
    ``range(total)`` exactly once across the whole group and -1 once the pool is exhausted.  Backed by ``store.add`` of
    the rendezvous store (TCPStore), an atomic fetch-add served by rank 0; a plain local counter without a process
    group.

    Indices are reserved in grains, one store round trip each: the FIRST reservation takes ``first`` consecutive
    indices (a rank's initial fill: its workspace slots, at most its fair share ceil(total / world) -- a GPU whose
    memory could hold the whole pool must not drain it), every later one ``grain`` (small, so that the tail of the
    pool goes to whichever GPU frees a slot first).  A pool of thousands of cheap sets uses a coarser grain than one
    of a few expensive sets: a round trip costs tens of microseconds."""

    def __init__(self, total, key, store=None, first=1, grain=1):
        import torch.distributed as dist
        self.total = int(total)
        self.key = f"gprn_pool/{key}"
        self.store = store
        self.local = 0
        self.first, self.grain = max(1, int(first)), max(1, int(grain))
        self.reserve = []              # indices reserved and not handed out yet
        self.trips = 0
        self.handed = []
        if store is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.store = dist.distributed_c10d._get_default_store()

    def _reserve(self, n):
        if self.store is None:
            lo = self.local
            self.local += n
        else:
            lo = self.store.add(self.key, n) - n
        self.trips += 1
        return range(lo, min(lo + n, self.total))

    def next(self):
        if not self.reserve:
            if self.trips and self.local >= self.total and self.store is None:
                return -1
            self.reserve = list(self._reserve(self.first if self.trips == 0 else self.grain))[::-1]
            if not self.reserve:
                return -1
        i = self.reserve.pop()
        self.handed.append(i)
        return i

    __call__ = next


def dealing_grains(B, world, slots):
    """(first, grain) of SharedCounter for a pool of B sets over `world` ranks with `slots` evaluations in flight per
    rank (0: unknown / as many as fit): the initial fill is a contiguous block of min(slots, fair share) sets, the
    rest is dealt in grains of 1 (few large sets) up to 64 (thousands of small ones)."""
    share = -(-B // max(1, world))
    first = share if slots <= 0 else min(slots, share)
    grain = max(1, min(64, first // 16))
    return first, grain


def shard_indices(B, world_size, rank, mode="strided"):
    """Indices of the hyper-parameter sets evaluated by ``rank``.

    ``strided`` (round-robin) evens out the spread of iteration counts across ranks when neighbouring
    sets are similar; ``block`` keeps contiguous ranges."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    if mode == "strided":
        return np.arange(rank, B, world_size)
    if mode == "block":
        base, rem = divmod(B, world_size)
        start = rank * base + min(rank, rem)
        return np.arange(start, start + base + (1 if rank < rem else 0))
    raise ValueError(f"unknown shard mode {mode!r}")


def lockstep_cost(iters, step_times):
    """Cost of evaluating a block of sets in lock-step (one batched call): while n sets are still active an iteration
    takes ``step_times[n-1]``; the sets drop out one by one as they converge."""
    it = np.sort(np.asarray(iters, dtype=float))
    n = it.size
    cost, prev = 0.0, 0.0
    for k in range(n):                       # n - k sets are active between it[k-1] and it[k]
        cost += (it[k] - prev) * step_times[min(n - k, len(step_times)) - 1]
        prev = it[k]
    return cost


def balanced_assignment(costs, world_size, step_times=None):
    """Equal-count partition of the sets over ranks that evens out the cost (iteration counts of a previous pass
    are a good predictor: the fixed point a set converges to moves little between passes).

    Sets are sorted by cost and dealt in serpentine order, so every rank receives the same number of sets (+-1)
    and the per-rank cost sums differ by at most one cost gap per round.  With ``step_times`` (measured time of one
    lock-step iteration with 1, 2, ... active sets) the blocks are then refined by pairwise swaps that lower the
    largest ``lockstep_cost`` -- a rank whose block ends in a long single-set tail pays more per iteration than one
    whose sets converge together.  Deterministic.  Returns a list of index arrays."""
    costs = np.asarray(costs, dtype=float)
    order = np.argsort(-costs, kind="stable")
    out = [[] for _ in range(world_size)]
    for pos, i in enumerate(order):
        rnd, k = divmod(pos, world_size)
        out[k if rnd % 2 == 0 else world_size - 1 - k].append(int(i))
    if step_times is not None and world_size > 1:
        cost = lambda blk: lockstep_cost(costs[blk], step_times)
        for _ in range(200):
            bc = [cost(b) for b in out]
            worst = int(np.argmax(bc))
            best = None
            for other in range(world_size):
                if other == worst:
                    continue
                for ia, a in enumerate(out[worst]):
                    for ib, b in enumerate(out[other]):
                        na = out[worst][:ia] + [b] + out[worst][ia + 1:]
                        nb = out[other][:ib] + [a] + out[other][ib + 1:]
                        m = max(cost(na), cost(nb))
                        if m < bc[worst] - 1e-9 and (best is None or m < best[0]):
                            best = (m, other, na, nb)
            if best is None:
                break
            out[worst], out[best[1]] = best[2], best[3]
    return [np.array(sorted(o), dtype=np.int64) for o in out]


def gather_results(local_idx, local_vals, B, group=None, device=None):
    """All-gather per-rank results into full length-B arrays on every rank.

    ``local_vals`` is a dict name -> 1-D array aligned with ``local_idx``.  Uses torch.distributed
    (backend of the default/``group`` process group: nccl on GPUs, gloo on CPU); without an
    initialised process group it just scatters the local values (single process)."""
    import torch
    import torch.distributed as dist

    out = {}
    if not (dist.is_available() and dist.is_initialized()):
        for k, v in local_vals.items():
            full = np.zeros(B, dtype=np.asarray(v).dtype)
            full[local_idx] = v
            out[k] = full
        return out
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if backend == "nccl" else torch.device("cpu"))
    nmax = -(-B // world)
    idx_t = torch.full((nmax,), -1, dtype=torch.int64, device=dev)
    idx_t[: len(local_idx)] = torch.as_tensor(np.asarray(local_idx), dtype=torch.int64, device=dev)
    idx_all = [torch.empty_like(idx_t) for _ in range(world)]
    dist.all_gather(idx_all, idx_t, group=group)
    for k, v in local_vals.items():
        v = np.asarray(v)
        tdt = torch.float64 if v.dtype.kind == "f" else torch.int64
        buf = torch.zeros((nmax,), dtype=tdt, device=dev)
        buf[: len(local_idx)] = torch.as_tensor(v.astype(np.float64 if tdt == torch.float64 else np.int64), device=dev)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        full = np.zeros(B, dtype=v.dtype)
        for ids, vals in zip(idx_all, parts):
            ids = ids.cpu().numpy()
            ok = ids >= 0
            full[ids[ok]] = vals.cpu().numpy()[ok].astype(v.dtype)
        out[k] = full
    return out


def reduce_disjoint(arrays, group=None):
    """Combine per-rank result arrays whose supports are disjoint (every set was evaluated by exactly one rank, the
    others hold zeros): one all-reduce(SUM) per array -- the one collective of the path.  Accepts torch tensors
    (reduced in place, on their device) or numpy arrays (staged through the backend's device); returns the inputs'
    kind.  No-op without a process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arrays
    backend = dist.get_backend(group)
    out = []
    for a in arrays:
        if isinstance(a, torch.Tensor):
            dist.all_reduce(a, op=dist.ReduceOp.SUM, group=group)
            out.append(a)
        else:
            dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            t = torch.as_tensor(np.ascontiguousarray(a)).to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            out.append(t.cpu().numpy().astype(np.asarray(a).dtype))
    return out


def elbo_pool_sharded(gprn, parameters, max_iter=None, slots=0, key=None, group=None, evaluate=None):
    """Evaluate B hyper-parameter sets across all ranks with DYNAMIC dealing: every rank pulls the next set from a
    shared counter whenever one of its ``slots`` workspace slots frees up, then one all-reduce combines the results.
    Every rank returns the full (elbo[B], iters[B], status[B], owner[B]) -- ``owner`` is the rank that evaluated a set.

    ``gprn`` is this rank's ``inference`` bound to its GPU; every rank must pass the same ``parameters`` and call this
    the same number of times (``key`` names the counter of this call; default: a per-process call serial).
    ``evaluate(P, work_source)`` -> (elbo, iters, status, taken) replaces ``gprn.ELBO_batch`` in the CPU tests."""
    import torch.distributed as dist

    P = np.atleast_2d(np.asarray(parameters, dtype=float))
    B = P.shape[0]
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    first, grain = dealing_grains(B, world, slots)
    counter = SharedCounter(B, next(_pool_serial) if key is None else key, first=first, grain=grain)
    if evaluate is None:
        # a rank never holds more than its fair share in flight: a GPU whose memory fits the whole pool would
        # otherwise drain the counter before the other ranks get going
        cap = slots if (slots or world == 1) else first
        elbo, iters, status, taken = gprn.ELBO_batch(P, max_iter=max_iter, return_info=True, slots=cap,
                                                     work_source=counter)
    else:
        elbo, iters, status, taken = evaluate(P, counter)
    taken = np.asarray(taken, dtype=bool)
    elbo = np.where(taken, elbo, 0.0)
    owner = np.where(taken, rank + 1, 0).astype(np.int64)
    elbo, iters, status, owner = reduce_disjoint([elbo, np.where(taken, iters, 0).astype(np.int64),
                                                  np.where(taken, status, 0).astype(np.int64), owner], group=group)
    return elbo, iters, status, owner - 1


def elbo_batch_sharded(gprn, parameters, max_iter=None, mode="strided", group=None):
    """Evaluate B hyper-parameter sets across all ranks of the process group; every rank returns the
    full (elbo[B], iters[B], status[B]).  ``gprn`` is this rank's ``inference`` bound to its GPU."""
    import torch.distributed as dist

    P = np.atleast_2d(np.asarray(parameters, dtype=float))
    B = P.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    idx = shard_indices(B, world, rank, mode)
    if len(idx):
        elbo, iters, status = gprn.ELBO_batch(P[idx], max_iter=max_iter, return_info=True)
    else:
        elbo, iters, status = np.zeros(0), np.zeros(0, np.int32), np.zeros(0, np.int32)
    res = gather_results(idx, {"elbo": elbo, "iters": iters.astype(np.int64), "status": status.astype(np.int64)}, B,
                         group=group)
    return res["elbo"], res["iters"], res["status"]
