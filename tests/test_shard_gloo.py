"""Multi-process host logic (N>1 path) on CPU with the gloo backend, world_size 2: LPT sharding, per-rank runs,
the single all-reduce of the ELBO trace.  The compute callable here is the C oracle (tests may use it)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vbx_b200 import shard, synth


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    lens = rng.integers(200, 3000, size=257)
    for ws in (1, 2, 4, 8):
        parts = shard.partition(lens, ws)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lens)))
        loads = np.array([lens[p].sum() for p in parts])
        assert loads.max() - loads.min() <= lens.max()


def _worker(rank, world, port, lens, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import c_oracle
    d = synth.make_batch(lens, R=32, S=4, seed=2, dtype=np.float64)
    iters = 4

    def run_local(idx):
        offs = np.concatenate([[0], np.cumsum([lens[i] for i in idx])]) if idx else np.zeros(1, dtype=np.int64)
        fea = np.concatenate([d['fea'][d['offsets'][i]:d['offsets'][i + 1]] for i in idx]) if idx else np.zeros((0, 32))
        g0 = np.concatenate([d['gamma0'][d['offsets'][i]:d['offsets'][i + 1]] for i in idx]) if idx else np.zeros((0, 4))
        return c_oracle.vbx_oracle_batch(fea, d['Phi'], offs, g0, np.full(4, 0.25), 0.3, 17.0, 0.9, iters, -np.inf)

    def all_reduce(a):
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t)
        return t.numpy()

    out = shard.run_sharded(lens, run_local, rank, world, iters, all_reduce=all_reduce)
    if rank == 0:
        q.put((out['elbo_sum'], out['n_active']))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    lens = [50, 120, 33, 80, 64, 7]
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    elbo_sum, n_active = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from oracle import c_oracle
    d = synth.make_batch(lens, R=32, S=4, seed=2, dtype=np.float64)
    ref = c_oracle.vbx_oracle_batch(d['fea'], d['Phi'], d['offsets'], d['gamma0'], np.full(4, 0.25), 0.3, 17.0, 0.9, 4, -np.inf)
    np.testing.assert_allclose(elbo_sum, ref['Li'].sum(0), rtol=1e-12)
    assert np.all(n_active == len(lens))
