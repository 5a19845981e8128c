#!/usr/bin/env python
"""Benchmark of the VB-HMM EM hot path (BASELINE.json metric: x-vectors/s through 10 EM iterations).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one JSON line on rank 0)
    python bench.py --impl reference ...                     # the reference algorithm on the host cores

A step = one pass of the hot path over one synthetic batch: rho = X.V (projection) followed by 10 EM
iterations (epsilon = -inf, the reference then never breaks, VBx/VBx.py:122).  `value` is measured with
X / gamma0 resident in HBM; `e2e` goes through the host-buffer API (pinned host X and gamma0 copied in,
gamma/pi/Li copied out, inside the timed region).  Multi-GPU: recordings are independent, every rank owns
its own batch (weak scaling) and one NCCL all-reduce combines the per-iteration ELBO sums.
"""
import os
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # before numpy: the CPU baseline runs one process per core
os.environ.setdefault("OMP_NUM_THREADS", "1")
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'x-vectors/sec through 10 VB-HMM EM iters'
UNIT = 'x-vectors/s'

# name -> (B, T spec, S, iters, Fa, Fb, loopP)    SURVEY.md 8(d) / BASELINE.json configs
WORKLOADS = {
    # north_star headline: 4096 recordings x T=1000, D=256 / R=128 / S=16, 10 iterations, one GPU
    'headline': dict(B=4096, T=1000, S=16, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
    'c2': dict(B=256, T=1000, S=16, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
    'c3': dict(B=4096, T=(200, 3000), S=16, iters=20, Fa=0.3, Fb=17.0, loopP=0.99),
    'c4': dict(B=24, T=12000, S=30, iters=40, Fa=0.2, Fb=6.0, loopP=0.35),      # per GPU share of 192 recordings
    'c5s4': dict(B=1024, T=2000, S=4, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
    'c5s8': dict(B=1024, T=2000, S=8, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
    'c5s16': dict(B=1024, T=2000, S=16, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
    'c5s32': dict(B=1024, T=2000, S=32, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
    'c5s64': dict(B=1024, T=2000, S=64, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
    'tiny': dict(B=32, T=300, S=16, iters=10, Fa=0.3, Fb=17.0, loopP=0.99),
}
D_RAW, R_DIM = 256, 128


def workload_lengths(w, seed):
    if isinstance(w['T'], tuple):
        rng = np.random.default_rng(seed)
        return rng.integers(w['T'][0], w['T'][1] + 1, size=w['B']).astype(np.int64)
    return np.full(w['B'], w['T'], dtype=np.int64)


# ------------------------------------------------------------------------------------------------
# synthetic data on the device (same generative model as vbx_b200/synth.py, vectorised in torch)
# ------------------------------------------------------------------------------------------------
def make_device_batch(lengths, S, seed, device):
    import torch
    from vbx_b200 import synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    B, Tm = len(lengths), int(max(lengths))
    Phi = torch.from_numpy(synth.plda_phi(R_DIM)).to(device=device, dtype=torch.float32)
    V0 = torch.from_numpy(synth.projection_basis(D_RAW, R_DIM)).to(device=device, dtype=torch.float32)
    lens_d = torch.from_numpy(np.asarray(lengths)).to(device)
    keep = (torch.arange(Tm, device=device)[None, :] < lens_d[:, None]).reshape(-1)
    ragged = not bool(keep.all())
    n_spk = torch.randint(2, 9, (B,), device=device, generator=g)
    means = torch.randn((B, 8, R_DIM), device=device, generator=g) * Phi.sqrt()
    X_parts, G_parts = [], []
    chunk = max(1, (1 << 22) // Tm)            # recordings per chunk, bounds temporaries to ~1 GB
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        nb = b1 - b0
        switch = torch.rand((nb, Tm), device=device, generator=g) >= 0.99
        switch[:, 0] = True
        jump = (torch.rand((nb, Tm), device=device, generator=g) * n_spk[b0:b1, None]).long()
        idx = torch.where(switch, torch.arange(Tm, device=device)[None, :].expand(nb, Tm), torch.zeros((), dtype=torch.long, device=device))
        last = torch.cummax(idx, dim=1).values
        z = torch.gather(jump, 1, last)                                        # sticky Markov path
        fea = torch.gather(means[b0:b1], 1, z[:, :, None].expand(nb, Tm, R_DIM))
        fea = fea + torch.randn((nb, Tm, R_DIM), device=device, generator=g)
        noise = torch.randn((nb * Tm, D_RAW), device=device, generator=g)
        noise = noise - (noise @ V0) @ V0.T
        X = fea.reshape(-1, R_DIM) @ V0.T + 0.5 * noise
        gam = -torch.log(torch.rand((nb * Tm, S), device=device, generator=g).clamp_min(1e-12))
        gam = gam / gam.sum(1, keepdim=True)                                  # flat Dirichlet rows, VBx/VBx.py:82-83
        if ragged:
            k = keep[b0 * Tm:b1 * Tm]
            X, gam = X[k], gam[k]
        X_parts.append(X)
        G_parts.append(gam)
        del fea, noise, z, jump, switch, idx, last
    X = torch.cat(X_parts) if len(X_parts) > 1 else X_parts[0]
    gamma0 = torch.cat(G_parts) if len(G_parts) > 1 else G_parts[0]
    V = (V0 * Phi.sqrt()[None, :]).contiguous()
    return dict(X=X.contiguous(), V=V, V0=V0, Phi=Phi, gamma0=gamma0.contiguous())


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        if not sm:   # region shorter than the sampling period: use every sample we have
            sm = [float(l.split(',')[1]) for _, l in self.lines if len(l.split(',')) >= 9] or [float('nan')]
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm, one recording per task on all host cores
# (the reference's own one-process-per-recording model, AMI_run.sh:53-58)
# ------------------------------------------------------------------------------------------------
def _cpu_task(args):
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    X, V0, Phi, g0, S, iters, Fa, Fb, loopP = args
    from oracle import vbx_oracle as po
    fea = X.astype(np.float64) @ V0.astype(np.float64)        # the caller-side projection, VBx/vbhmm.py:153
    t = time.perf_counter()
    po.vbx_oracle(fea, Phi.astype(np.float64), loopProb=loopP, Fa=Fa, Fb=Fb, pi=S, gamma=g0.astype(np.float64),
                  maxIters=iters, epsilon=-np.inf)
    return time.perf_counter() - t


def cpu_baseline(sample, w, cores=None, repeats=1):
    """sample: list of (X [T,D], g0 [T,S]) numpy arrays + shared V0, Phi.  Returns (x-vec/s, cores, wall s)."""
    import multiprocessing as mp
    recs, V0, Phi = sample
    cores = cores or os.cpu_count() or 1
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    tasks = [(X, V0, Phi, g0, w['S'], w['iters'], w['Fa'], w['Fb'], w['loopP']) for X, g0 in recs]
    frames = sum(X.shape[0] for X, _ in recs)
    ctx = mp.get_context('fork')
    best = None
    with ctx.Pool(min(cores, len(tasks))) as pool:
        pool.map(_cpu_warm, range(min(cores, len(tasks))))
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_task, tasks, chunksize=1)
            wall = time.perf_counter() - t0
            best = wall if best is None else min(best, wall)
    return frames / best, min(cores, len(tasks)), best


def _cpu_warm(_):
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    from oracle import vbx_oracle  # noqa: F401
    return 0


def host_sample(w, n_rec, seed):
    """A bounded sample of the workload generated on the host with the numpy generator (same model)."""
    from vbx_b200 import synth
    lens = workload_lengths(w, seed)[:n_rec]
    d = synth.make_batch(lens, R=R_DIM, S=w['S'], seed=seed, D=D_RAW, dtype=np.float32)
    recs = [(d['X'][lo:hi], d['gamma0'][lo:hi]) for lo, hi in zip(d['offsets'][:-1], d['offsets'][1:])]
    return recs, synth.projection_basis(D_RAW, R_DIM), d['Phi']


# ------------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def run_reference(args, w, wname):
    """--impl reference: the reference's CPU algorithm (oracle port, float64 numpy, log-domain recursions) on all
    host cores; each step = a bounded sample of the workload."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_rec = max(8, cores)
    sample = host_sample(w, n_rec, seed=1)
    frames = sum(x.shape[0] for x, _ in sample[0])
    vals, walls = [], []
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    tasks = [(X, sample[1], sample[2], g0, w['S'], w['iters'], w['Fa'], w['Fb'], w['loopP']) for X, g0 in sample[0]]
    with ctx.Pool(min(cores, len(tasks))) as pool:
        pool.map(_cpu_warm, range(min(cores, len(tasks))))
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_task, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                walls.append(dt)
    ms = 1e3 * float(np.mean(walls))
    value = frames / (ms / 1e3)
    used = min(cores, len(tasks))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(w, wname, args.gpus, note=f'bounded sample: {len(tasks)} recordings ({frames} x-vectors) per step'),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': used, 'kind': 'port',
                         'sample': f'{len(tasks)} recordings x {w["iters"]} iterations of the workload per step, one per process '
                                   f'(oracle/vbx_oracle.py: float64 numpy restatement of VBx/VBx.py, OPENBLAS_NUM_THREADS=1)'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(w, wname, n_gpus, note=None):
    T = w['T']
    cfg = {'workload': f'{wname}: B={w["B"]} recordings/GPU x T={"U[%d,%d]" % T if isinstance(T, tuple) else T} '
                       f'x D={D_RAW} -> R={R_DIM}, S={w["S"]}, {w["iters"]} EM iterations',
           'recordings_per_gpu': w['B'], 'frames_per_recording': list(T) if isinstance(T, tuple) else T, 'D': D_RAW,
           'R': R_DIM, 'S': w['S'], 'em_iterations': w['iters'], 'Fa': w['Fa'], 'Fb': w['Fb'], 'loopProb': w['loopP'],
           'parallelism': f'recordings sharded over {n_gpus} GPU(s), one NCCL all-reduce of the ELBO trace',
           'l2': 'inputs larger than L2 (rho alone exceeds 126 MB)' if w['B'] * (np.mean(T) if isinstance(T, tuple) else T) * R_DIM * 4 > 2 * 126e6
                 else 'L2 flushed between steps (256 MB scratch write)'}
    if note:
        cfg['note'] = note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='headline', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--fb-spl', type=int, default=0)
    ap.add_argument('--projection', type=int, default=0)
    ap.add_argument('--fb-classic', type=int, default=0, help='1 = normalise-every-frame forward-backward sweep (A/B against the look-ahead kernel)')
    ap.add_argument('--front', default='project', choices=['project', 'xvectors'],
                    help="what feeds the EM loop: 'project' = rho = X.V (the headline definition, SURVEY 8d); 'xvectors' = the "
                         "real-data chain vbx_prepare_xvectors (x-vector transform + PLDA projection, two tcgen05 passes)")
    ap.add_argument('--extra', default='', help='comma separated extra workloads to time (kernel-only) in the same run')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    w, wname = WORKLOADS[args.workload], args.workload
    if args.impl == 'reference':
        return run_reference(args, w, wname)

    import torch
    import torch.distributed as dist
    from vbx_b200.batch import VbxBatch
    from vbx_b200.host_pipeline import HostPipeline

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the VB-HMM path has no CPU fallback; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()

    def dbg(msg):
        if os.environ.get('VBX_BENCH_DEBUG'):
            print(f'[rank {rank}] {msg}', file=sys.stderr, flush=True)

    def time_workload(w, wname, steps, warmup, with_clocks):
        lengths = workload_lengths(w, seed=1000 + rank)
        data = make_device_batch(lengths, w['S'], seed=17 + rank, device=device)
        N = int(lengths.sum())
        dbg(f'{wname}: data on device, N={N}')
        vb = VbxBatch(lengths, R_DIM, w['S'], device=device)
        if args.fb_spl:
            vb.set_option('fb_states_per_lane', args.fb_spl)
        if args.projection:
            vb.set_option('projection', args.projection)
        if args.fb_classic:
            vb.set_option('fb_classic', 1)
        vb.set_option('timing', 1)
        S = vb.S
        rho = torch.empty((N, R_DIM), dtype=torch.float32, device=device)
        gamma = torch.zeros((N, S), dtype=torch.float32, device=device)
        pi = torch.empty((len(lengths), S), dtype=torch.float32, device=device)
        pi0 = torch.zeros(S, device=device)
        pi0[:w['S']] = 1.0 / w['S']
        elbo_sum = torch.zeros(w['iters'], dtype=torch.float64, device=device)
        flush = None
        if N * R_DIM * 4 <= 2 * 126e6:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        vb.n_states = None if w['S'] == S else torch.full((len(lengths),), w['S'], dtype=torch.int32, device=device)

        model = None
        if args.front == 'xvectors':       # synthetic model with the shapes of VBx/models/ResNet101_16kHz
            gen = torch.Generator(device='cpu').manual_seed(5)
            rnd = lambda *shape: torch.randn(*shape, generator=gen)
            q, _ = torch.linalg.qr(rnd(R_DIM, R_DIM))
            model = [t.to(device).contiguous() for t in (
                rnd(D_RAW) * 0.5, rnd(D_RAW, R_DIM) / D_RAW ** 0.5, rnd(R_DIM) * 0.05, rnd(R_DIM) * 0.02,
                q * (2.0 + 18.0 * torch.rand(R_DIM, generator=gen))[:, None])] + [data['Phi']]

        def step():
            if model is None:
                vb.prepare_project(data['X'], data['V'], data['Phi'], out=rho)
            else:
                vb.prepare_xvectors(data['X'], *model, out=rho)
            gamma[:, :w['S']].copy_(data['gamma0'])
            pi.copy_(pi0.expand_as(pi))
            out = vb.run(gamma, pi, Fa=w['Fa'], Fb=w['Fb'], loopProb=w['loopP'], maxIters=w['iters'], epsilon=-float('inf'))
            s = out['Li'].sum(0)
            if world > 1:
                dist.all_reduce(s)       # the one collective of the path: global ELBO trace
            elbo_sum.copy_(s)
            return out

        for _ in range(warmup):
            if flush is not None:
                flush.zero_()
            step()
        torch.cuda.synchronize()
        dbg('warm-up done')
        vb.timings(reset=True)
        l0 = vb.launches
        sampler = ClockSampler(local_rank) if (with_clocks and rank == 0) else None
        if sampler:
            sampler.start()
            time.sleep(0.25)
        torch.cuda.synchronize()
        barrier()                      # every rank enters the timed region together
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_wall0 = time.time()
        for i in range(steps):
            if flush is not None:
                flush.zero_()
            evs[i][0].record()
            out = step()
            evs[i][1].record()
        torch.cuda.synchronize()
        barrier()
        t_wall1 = time.time()
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        per_step = [a.elapsed_time(b) for a, b in evs]
        if os.environ.get('VBX_BENCH_DEBUG'):
            print(f'[rank {rank}] per-step ms: {[round(x, 3) for x in per_step]} wall {1e3 * (t_wall1 - t_wall0) / steps:.3f} ms/step', file=sys.stderr, flush=True)
        ms = sum(per_step) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            nt = torch.tensor([N], dtype=torch.float64, device=device)
            dist.all_reduce(nt)
            N_total = int(nt.item())
        else:
            N_total = N
        timings = vb.timings(reset=True)
        launches = (vb.launches - l0) / steps
        assert bool(torch.isfinite(elbo_sum).all()), 'non-finite ELBO in the benchmark run'
        res = dict(ms=ms, N=N, N_total=N_total, timings=timings, launches=launches, clocks=clocks, lengths=lengths,
                   data=data, vb=vb, S=S, out=out, steps=steps)
        return res

    res = time_workload(w, wname, args.steps, args.warmup, with_clocks=True)
    dbg(f'timed region done: {res["ms"]:.3f} ms/step')
    ms, N, N_total = res['ms'], res['N'], res['N_total']
    value = N_total / (ms / 1e3)

    # ---- roofline of the dominant kernel (CUDA events recorded inside the C ABI on the launching stream) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_gbs = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s (B200_PROFILING.md)'
    S = res['S']
    alg_bytes = {   # algorithmic bytes per frame per launch
        'project': 4 * D_RAW + 4 * R_DIM if args.front == 'project' else 4 * D_RAW + 3 * 4 * R_DIM,
        'prepare': 4 * R_DIM,
        'mstep_partial': 4 * R_DIM + 4 * S,
        'loglik': 4 * R_DIM + 4 * S + 4,
        'forward_backward': 5 * 4 * S + 12,
    }
    per_kernel = {}
    for k, (tms, cnt) in res['timings'].items():
        if cnt:
            per_kernel[k] = {'ms_per_launch': tms / cnt, 'launches_per_step': cnt / res['steps'], 'ms_per_step': tms / res['steps']}
            if k in alg_bytes:
                per_kernel[k]['gbs'] = alg_bytes[k] * N / (tms / cnt * 1e-3) / 1e9
    dom = max((k for k in per_kernel if k in alg_bytes), key=lambda k: per_kernel[k]['ms_per_step'])
    roof = {'bound': 'hbm', 'kernel': dom, 'achieved': per_kernel[dom]['gbs'], 'peak': peak_gbs, 'unit': 'GB/s',
            'frac': per_kernel[dom]['gbs'] / peak_gbs, 'traffic': None, 'peak_source': peak_src,
            'algorithmic_bytes_per_launch': alg_bytes[dom] * N, 'avg_launch_ms': per_kernel[dom]['ms_per_launch'],
            'share_of_step': per_kernel[dom]['ms_per_step'] / ms}
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            if tr.get('workload') == wname and dom in tr.get('dram_bytes_per_launch', {}):
                roof['traffic'] = tr['dram_bytes_per_launch'][dom]
                roof['traffic_source'] = tr.get('source')
        except Exception:
            pass
    step_bytes = N * (4 * D_RAW + 4 * R_DIM + 4 * w['S'] + 2 * 4 * R_DIM * w['iters'])   # N*(1600+1024*iters) at S=16
    whole = {'algorithmic_bytes_per_step': step_bytes, 'achieved_gbs': step_bytes / (ms * 1e-3) / 1e9 * (N_total / N) / max(world, 1),
             'frac_of_peak': step_bytes / (ms * 1e-3) / 1e9 / peak_gbs}

    # ---- end-to-end through the host-buffer API (pinned host inputs, H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e and args.front == 'project':
        hp = HostPipeline(res['lengths'], D_RAW, R_DIM, w['S'], device=device)
        Xh = torch.empty((N, D_RAW), dtype=torch.float32).pin_memory()
        Gh = torch.empty((N, w['S']), dtype=torch.float32).pin_memory()
        Xh.copy_(res['data']['X'])
        Gh.copy_(res['data']['gamma0'])
        torch.cuda.synchronize()
        kw = dict(Fa=w['Fa'], Fb=w['Fb'], loopProb=w['loopP'], maxIters=w['iters'], epsilon=-float('inf'))
        for _ in range(2):
            hp.run(Xh, res['data']['V'], res['data']['Phi'], Gh, **kw)
        barrier()
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 5))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_e2e):
            o = hp.run(Xh, res['data']['V'], res['data']['Phi'], Gh, **kw)
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ems = e0.elapsed_time(e1) / n_e2e
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        # parity of the two paths on the same data (device-resident vs host-pipelined)
        dmax = float((o['gamma'].to(device) - res['out']['gamma'][:, :w['S']]).abs().max())
        e2e = {'value': N_total / (ems / 1e3), 'unit': UNIT, 'ms_per_step': ems, 'h2d_bytes_per_step': hp.h2d_bytes,
               'd2h_bytes_per_step': hp.d2h_bytes, 'chunks': hp.n_chunks, 'max_abs_gamma_diff_vs_resident': dmax,
               'api': 'vbx_b200.host_pipeline.HostPipeline.run (pinned host X, gamma0 -> gamma, pi, Li on the host)'}
        del Xh, Gh, hp
        dbg('e2e done')

    # ---- CPU baseline on this box's host cores (rank 0, N=1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_rec = max(8, min(cores, 64))
        sample = host_sample(w, n_rec, seed=1)
        v, used, wall = cpu_baseline(sample, w, cores=cores)
        cpu = {'value': v, 'unit': UNIT, 'cores': used, 'kind': 'port', 'wall_s': wall,
               'sample': f'{len(sample[0])} recordings of the workload ({sum(x.shape[0] for x, _ in sample[0])} x-vectors, '
                         f'{w["iters"]} iterations), one process per recording on {used} cores; oracle/vbx_oracle.py = float64 numpy '
                         'restatement of VBx/VBx.py (log-domain recursions, same per-frame Python overhead as the reference)'}
        try:
            from oracle import c_oracle
            t0 = time.perf_counter()
            recs = sample[0][:8]
            fea = np.concatenate([x.astype(np.float64) @ sample[1] for x, _ in recs])
            g0 = np.concatenate([g for _, g in recs])
            offs = np.concatenate([[0], np.cumsum([x.shape[0] for x, _ in recs])])
            c_oracle.vbx_oracle_batch(fea, sample[2], offs, g0, np.full(w['S'], 1.0 / w['S']), w['Fa'], w['Fb'], w['loopP'], w['iters'], -np.inf)
            cpu['c_oracle_single_thread'] = {'value': fea.shape[0] / (time.perf_counter() - t0), 'unit': UNIT,
                                             'note': 'oracle/vbx_oracle_c.c (O(S) scaled recursion, float64), 1 thread, 8 recordings'}
        except Exception as ex:   # the C oracle is optional here
            cpu['c_oracle_single_thread'] = {'error': str(ex)}

    extra = {}
    for name in [x for x in args.extra.split(',') if x]:
        r2 = time_workload(WORKLOADS[name], name, max(3, args.steps // 2), 3, with_clocks=False)
        extra[name] = {'value': r2['N_total'] / (r2['ms'] / 1e3), 'unit': UNIT, 'ms_per_step': r2['ms'],
                       'config': workload_config(WORKLOADS[name], name, world),
                       'kernels_ms_per_step': {k: v[0] / r2['steps'] for k, v in r2['timings'].items() if v[1]}}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic (seeded sticky-Markov speakers in PLDA space, SURVEY.md 8d; generated on the device)',
            'config': workload_config(w, wname, world, note=None if args.front == 'project' else
                                      'front end = vbx_prepare_xvectors (x-vector transform + PLDA projection) instead of rho = X.V; not the headline definition'),
            'target': {'north_star_x_vectors_per_s': 1e7, 'ratio': value / 1e7 / max(world, 1)},
            'roofline': roof, 'whole_step': whole, 'kernels': per_kernel, 'gpu_launches': res['launches'] * args.steps,
            'gpu_launches_per_step': res['launches'], 'clocks': res['clocks'], 'e2e': e2e, 'cpu_baseline': cpu,
        }
        if extra:
            line['other_workloads'] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
