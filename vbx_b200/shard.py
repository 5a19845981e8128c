"""Sharding of a batch of independent recordings over the GPUs of one node.

The reference parallelises by launching one process per recording (AMI_run.sh:53-58); here each rank owns a
subset of the recordings, runs the EM loop on its own GPU with no data-path collective, and a single all-reduce
combines the per-iteration ELBO sums.  The compute callable is injected so the host logic is testable with gloo on CPU.
"""
import numpy as np


def partition(lengths, world_size):
    """Greedy longest-processing-time partition by frame count.  Returns a list (per rank) of recording indices,
    each sorted ascending so that shards keep the caller's order."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind='stable')
    load = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for b in order:
        r = int(np.argmin(load))
        shards[r].append(int(b))
        load[r] += lengths[b]
    return [sorted(s) for s in shards]


def run_sharded(lengths, run_local, rank, world_size, max_iters, all_reduce=None, gather=None):
    """Run `run_local(indices) -> dict(Li [n,max_iters] float64 NaN-padded, ...)` on this rank's shard.

    all_reduce(np.ndarray float64) -> np.ndarray : sum over ranks (NCCL/gloo wrapper supplied by the caller).
    gather(obj) -> list of objs on every rank (optional; used to reassemble per-recording outputs).
    Returns dict(indices, local (run_local's result), elbo_sum [max_iters] global, n_active [max_iters] global)."""
    shards = partition(lengths, world_size)
    mine = shards[rank]
    local = run_local(mine)
    Li = np.asarray(local['Li'], dtype=np.float64).reshape(len(mine), max_iters) if len(mine) else np.zeros((0, max_iters))
    stats = np.zeros(2 * max_iters, dtype=np.float64)
    stats[:max_iters] = np.nansum(Li, axis=0)
    stats[max_iters:] = np.sum(~np.isnan(Li), axis=0)
    if all_reduce is not None and world_size > 1:
        stats = all_reduce(stats)
    out = dict(indices=mine, local=local, elbo_sum=stats[:max_iters], n_active=stats[max_iters:].astype(np.int64))
    if gather is not None:
        out['all_indices'] = gather(mine)
    return out
