"""The steps either side of the VB-HMM call in the reference's driver, batched:

  before:  x-vector transform  l2_norm(lda^T . l2_norm(x - mean1) - mean2)            VBx/vbhmm.py:125-129
           PLDA simultaneous diagonalisation (host, once per model)                   VBx/vbhmm.py:107-113
           projection into the PLDA space  (x - plda_mu) . plda_tr^T [:, :lda_dim]    VBx/vbhmm.py:153
           soft initialisation from hard AHC labels  softmax(onehot * smoothing)      VBx/vbhmm.py:150-152
  after:   hard labels  argsort(-q)[:, 0] (and 2nd best)                               VBx/vbhmm.py:160-162
           merge_adjacent_labels + RTTM lines                                          VBx/diarization_lib.py:113-135, vbhmm.py:48-51

The dense algebra runs on the device through torch (library GEMMs: not the hot path); the label merge is a
host-side O(T) pass exactly like the reference's.  AHC itself (VBx/vbhmm.py:131-146) is out of scope.
"""
import numpy as np
import torch


def diagonalise_plda(plda_mu, plda_tr, plda_psi):
    """VBx/vbhmm.py:107-113: generalised eigen-problem B v = lambda W v; returns (mu, tr, psi) with psi descending."""
    from scipy.linalg import eigh
    W = np.linalg.inv(plda_tr.T.dot(plda_tr))
    B = np.linalg.inv((plda_tr.T / plda_psi).dot(plda_tr))
    acvar, wccn = eigh(B, W)
    return plda_mu, wccn.T[::-1].copy(), acvar[::-1].copy()


def l2_norm_rows(x):
    """VBx/diarization_lib.py:172-187 for a matrix of row vectors."""
    return x / torch.linalg.vector_norm(x, dim=1, keepdim=True)


def xvector_transform(x_raw, mean1, mean2, lda):
    """VBx/vbhmm.py:129 on the device.  x_raw [N,256] -> [N,128] (float64 like the reference's h5 arrays)."""
    x = l2_norm_rows(x_raw - mean1[None, :])
    x = x @ lda - mean2[None, :]
    return l2_norm_rows(x)


def plda_project(x, plda_mu, plda_tr, lda_dim):
    """VBx/vbhmm.py:153: fea = (x - plda_mu) . plda_tr^T, first lda_dim columns."""
    return ((x - plda_mu[None, :]) @ plda_tr.T)[:, :lda_dim]


def soft_init(labels, n_states, smoothing):
    """VBx/vbhmm.py:150-152: qinit = softmax(onehot(labels) * smoothing) (rows), float32 on labels' device."""
    q = torch.zeros((labels.shape[0], n_states), dtype=torch.float32, device=labels.device)
    q.scatter_(1, labels.long()[:, None], float(smoothing))
    return torch.softmax(q, dim=1)


def hard_labels(gamma, second=False):
    """VBx/vbhmm.py:160-162: most (and second most) likely speaker per frame.  Stable ordering like np.argsort(-q)."""
    order = torch.argsort(gamma, dim=1, descending=True, stable=True)
    return (order[:, 0], order[:, 1]) if second and gamma.shape[1] > 1 else order[:, 0]


def merge_adjacent_labels(starts, ends, labels):
    """Compact labelled segments: merge adjacent/overlapping segments with equal labels, split the overlap of
    differently labelled neighbours in the middle.  Same result as VBx/diarization_lib.py:113-135."""
    starts = np.asarray(starts, dtype=np.float64)
    ends = np.asarray(ends, dtype=np.float64)
    labels = np.asarray(labels)
    if len(labels) == 0:
        return starts.copy(), ends.copy(), labels.copy()
    touching = np.isclose(ends[:-1], starts[1:]) | (ends[:-1] > starts[1:])
    cut = np.nonzero(~touching | (labels[1:] != labels[:-1]))[0]        # a new segment starts at cut+1
    first = np.concatenate([[0], cut + 1])
    last = np.concatenate([cut, [len(labels) - 1]])
    s, e, l = starts[first].copy(), ends[last].copy(), labels[first].copy()
    over = np.nonzero(s[1:] < e[:-1])[0]
    mid = (e[over] + s[over + 1]) / 2.0
    e[over] = mid
    s[over + 1] = mid
    return s, e, l


def rttm_lines(recording, starts, ends, labels):
    """VBx/vbhmm.py:48-51."""
    return [f'SPEAKER {recording} 1 {s:03f} {e - s:03f} <NA> <NA> {int(l) + 1} <NA> <NA>'
            for s, e, l in zip(starts, ends, labels)]


def diarize_recording(x_raw, seg_times, ahc_labels, transform, plda, Fa, Fb, loopP, lda_dim=128, smoothing=5.0,
                      max_iters=40, epsilon=1e-6, device=None, recording='rec', chain='tcgen05',
                      plda_is_diagonal=False, threshold=-0.015):
    """One recording end to end on the device, the AHC+VB branch of VBx/vbhmm.py:120-172.
    transform = (mean1, mean2, lda); plda = (mu, tr, psi) as read from the Kaldi model (diagonalised here as in
    VBx/vbhmm.py:136-143 unless plda_is_diagonal says it already is).
    ahc_labels=None runs the AHC initialisation on the device too (vbx_ahc, VBx/vbhmm.py:131-146, `threshold` is
    the --threshold bias); otherwise the given labels seed gamma.
    chain='tcgen05' runs the x-vector transform and the PLDA projection in the fused tensor-core kernels
    (vbx_prepare_xvectors; needs lda_dim == 128 and a raw dimension that is a multiple of 32), chain='float64'
    evaluates them with float64 torch ops on the device.  Returns (rttm lines, labels, gamma)."""
    from .batch import VbxBatch
    from . import ahc as _ahc
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    f32 = lambda a: f64(a).float().contiguous()
    mu, tr, psi = plda if plda_is_diagonal else diagonalise_plda(*plda)
    T = int(np.asarray(x_raw).shape[0])
    if chain not in ('tcgen05', 'float64'):
        raise ValueError("chain must be 'tcgen05' or 'float64'")
    if chain == 'tcgen05' and (lda_dim != 128 or tr.shape[0] != 128):
        raise ValueError("chain='tcgen05' needs a 128-dimensional PLDA space")
    # the front end does not depend on the number of speakers: run it on a planning-only batch first
    front = VbxBatch([T], lda_dim, 1, device=dev)
    if chain == 'tcgen05':
        mean1, mean2, lda = transform
        rho, x = front.prepare_xvectors(f32(x_raw), f32(mean1), f32(lda), f32(mean2), f32(mu), f32(tr), f32(psi))
        Phi = f32(psi)
    else:
        mean1, mean2, lda = (f64(a) for a in transform)
        x = xvector_transform(f64(x_raw), mean1, mean2, lda).contiguous()
        fea = plda_project(x, f64(mu), f64(tr), lda_dim)
        Phi = f64(psi[:lda_dim]).float()
        rho = None
    if ahc_labels is None:
        ahc_labels = _ahc.ahc_batch(front, x, threshold=threshold)[0][0]
    front.close()
    S = int(np.max(ahc_labels)) + 1
    q0 = soft_init(torch.from_numpy(np.asarray(ahc_labels)).to(dev), S, smoothing)
    vb = VbxBatch([T], lda_dim, S, device=dev)
    vb.set_option('gemm', 1)
    g = torch.zeros((T, vb.S), dtype=torch.float32, device=dev)
    g[:, :S] = q0
    p = torch.zeros((1, vb.S), dtype=torch.float32, device=dev)
    p[0, :S] = 1.0 / S
    if rho is not None:       # the speaker count was not known when the front end ran: hand its features to the new plan
        fea = rho / torch.sqrt(Phi)[None, :]
    vb.prepare_scale(fea.float().contiguous(), Phi)
    vb.run(g, p, Fa=Fa, Fb=Fb, loopProb=loopP, maxIters=max_iters, epsilon=epsilon)
    labels = vb.hard_labels(g).cpu().numpy().astype(np.int64)      # only the labels leave the device
    s, e, l = merge_adjacent_labels(seg_times[:, 0], seg_times[:, 1], labels)
    vb.close()
    return rttm_lines(recording, s, e, l), labels, g[:, :S]


def diarize_batch(recordings, transform, plda, Fa, Fb, loopP, lda_dim=128, threshold=-0.015, smoothing=5.0, init='AHC+VB',
                  max_iters=40, epsilon=1e-6, device=None, chain='auto', output_2nd=False):
    """Every recording of an archive in ONE batch on the device - the body of the loop VBx/vbhmm.py:120-179 for all
    recordings at once: x-vector transform + PLDA projection (vbx_prepare_xvectors), AHC initialisation (vbx_ahc), the
    VB-HMM with the reference's stop rule (vbx_run), hard labels (vbx_hard_labels); merging and RTTM lines on the host.

    recordings: {name: (x_raw [T,Dx] float array, seg_times [T,2])} in archive order.  transform = (mean1, mean2, lda),
    plda = (mu, tr, psi) as read from the Kaldi model (diagonalised here as VBx/vbhmm.py:107-113 does).
    init: 'AHC' (clustering only) or 'AHC+VB' (VBx/vbhmm.py:131,147).  chain: 'tcgen05' (fused tensor-core front end,
    needs lda_dim == 128 and a 128-dim PLDA), 'float64' (float64 torch ops), 'auto' = tcgen05 when the shapes allow.
    Returns {name: dict(rttm, labels, labels2nd or None, n_speakers, iterations)}."""
    from .batch import VbxBatch
    from . import ahc as _ahc
    if init not in ('AHC', 'AHC+VB'):
        raise ValueError('Wrong option for args.initialization.')          # VBx/vbhmm.py:163-164
    if not torch.cuda.is_available():
        from ._lib import VbxError
        raise VbxError('diarize_batch(): no CUDA device - vbx_b200 has no CPU fallback')
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    names = list(recordings)
    lens = np.array([np.asarray(recordings[n][0]).shape[0] for n in names], dtype=np.int64)
    if len(names) == 0:
        return {}
    f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    f32 = lambda a: f64(a).float().contiguous()
    mu, tr, psi = diagonalise_plda(*plda)
    mean1, mean2, lda = transform
    Dx = int(np.asarray(recordings[names[0]][0]).shape[1])
    if chain == 'auto':
        chain = 'tcgen05' if (lda_dim == 128 and tr.shape[0] == 128 and lda.shape[1] == 128 and Dx % 32 == 0) else 'float64'
    x_all = np.concatenate([np.asarray(recordings[n][0], dtype=np.float64) for n in names])
    front = VbxBatch(lens, 128 if chain == 'tcgen05' else 4, 1, device=dev, exact_stop=False)
    if chain == 'tcgen05':
        rho, x = front.prepare_xvectors(f32(x_all), f32(mean1), f32(lda), f32(mean2), f32(mu), f32(tr), f32(psi))
        Phi = f32(psi)
        fea = (rho / torch.sqrt(Phi)[None, :]).contiguous()
    else:
        x = xvector_transform(f64(x_all), f64(mean1), f64(mean2), f64(lda)).contiguous()
        fea = plda_project(x, f64(mu), f64(tr), lda_dim).float().contiguous()
        Phi = f64(psi[:lda_dim]).float().contiguous()
    ahc_labels, _, _ = _ahc.ahc_batch(front, x, threshold=threshold)      # VBx/vbhmm.py:131-146
    front.close()
    offs = np.concatenate([[0], np.cumsum(lens)])
    out = {}
    labels1 = [l.astype(np.int64) for l in ahc_labels]
    labels2 = [None] * len(names)
    iters = [0] * len(names)
    if init.endswith('VB'):
        ns = np.array([int(l.max()) + 1 if len(l) else 1 for l in ahc_labels], dtype=np.int32)
        if ns.max() > 64:
            bad = names[int(ns.argmax())]
            raise ValueError(f'recording {bad!r}: AHC produced {int(ns.max())} clusters; the VB-HMM kernels hold at most 64 HMM states '
                             '(raise --threshold, or run that recording with --init AHC)')
        R = int(fea.shape[1])
        pad = (-R) % 4
        if pad:                     # inert zero features (see api.VBx): labels do not depend on them
            fea = torch.cat([fea, torch.zeros((fea.shape[0], pad), device=dev)], dim=1).contiguous()
            Phi = torch.cat([Phi, torch.zeros(pad, device=dev)]).contiguous()
        vb = VbxBatch(lens, R + pad, ns, device=dev)
        lab_d = torch.from_numpy(np.concatenate(ahc_labels)).to(dev)
        g = torch.zeros((int(lens.sum()), vb.S), dtype=torch.float32, device=dev)
        p = torch.zeros((len(names), vb.S), dtype=torch.float32, device=dev)
        for b in range(len(names)):              # VBx/vbhmm.py:150-152: qinit = softmax(onehot * smoothing)
            g[offs[b]:offs[b + 1], :ns[b]] = soft_init(lab_d[offs[b]:offs[b + 1]], int(ns[b]), smoothing)
            p[b, :ns[b]] = 1.0 / ns[b]
        vb.prepare_scale(fea, Phi)
        res = vb.run(g, p, Fa=Fa, Fb=Fb, loopProb=loopP, maxIters=max_iters, epsilon=epsilon)     # VBx/vbhmm.py:154-158
        first, second = vb.hard_labels(g, second=True)                     # VBx/vbhmm.py:160-162
        first, second = first.cpu().numpy().astype(np.int64), second.cpu().numpy().astype(np.int64)
        iters = res['n_iters'].cpu().numpy().tolist()
        for b in range(len(names)):
            labels1[b] = first[offs[b]:offs[b + 1]]
            if ns[b] > 1:
                labels2[b] = second[offs[b]:offs[b + 1]]
        vb.close()
    for b, n in enumerate(names):
        seg = np.asarray(recordings[n][1], dtype=np.float64)
        s, e, l = merge_adjacent_labels(seg[:, 0], seg[:, 1], labels1[b])   # VBx/vbhmm.py:169
        item = dict(rttm=rttm_lines(n, s, e, l), labels=labels1[b], labels2nd=labels2[b], iterations=int(iters[b]),
                    n_speakers=int(len(set(labels1[b].tolist()))), rttm2nd=None)
        if output_2nd and labels2[b] is not None:
            s2, e2, l2 = merge_adjacent_labels(seg[:, 0], seg[:, 1], labels2[b])   # VBx/vbhmm.py:174-179
            item['rttm2nd'] = rttm_lines(n, s2, e2, l2)
        out[n] = item
    return out
