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
