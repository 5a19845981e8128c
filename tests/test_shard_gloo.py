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


def test_strong_scaling_job_is_the_same_for_every_world_size():
    """bench.py --workload c4 (strong scaling): recording i of the job is generated from seed + i whichever rank owns it,
    and the LPT shards cover the job exactly once - so 1, 2, 4 and 8 ranks time the SAME 192 recordings."""
    import importlib.util
    import os
    import torch
    from vbx_b200 import shard
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w = dict(bench.WORKLOADS['tinystrong'])
    lengths = bench.workload_lengths(w, seed=1000)
    dev = torch.device('cpu')
    whole = bench.make_device_shard(lengths, list(range(len(lengths))), w['S'], seed=17, device=dev)
    offs = np.concatenate([[0], np.cumsum(lengths)])
    for world in (2, 4):
        parts = shard.partition(lengths, world)
        assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
        loads = [int(lengths[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= int(lengths.max())               # LPT balance
        for p in parts:
            mine = bench.make_device_shard(lengths, p, w['S'], seed=17, device=dev)
            o = 0
            for i in p:
                n = int(lengths[i])
                assert torch.equal(mine['X'][o:o + n], whole['X'][offs[i]:offs[i + 1]])
                assert torch.equal(mine['gamma0'][o:o + n], whole['gamma0'][offs[i]:offs[i + 1]])
                o += n
