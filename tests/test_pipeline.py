"""The 'next' rows either side of the hot path against the ES2005a golden produced by the reference."""
import os

import numpy as np
import pytest
import torch

from vbx_b200 import pipeline

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def es():
    return np.load(os.path.join(GOLD, 'es2005a.npz'))


def test_merge_adjacent_labels_matches_reference_output(es):
    s, e, l = pipeline.merge_adjacent_labels(es['seg_times'][:, 0], es['seg_times'][:, 1], es['labels'])
    np.testing.assert_allclose(s, es['rttm_starts'])
    np.testing.assert_allclose(e, es['rttm_ends'])
    assert np.array_equal(l, es['rttm_labels'])
    lines = pipeline.rttm_lines('ES2005a', s, e, l)
    assert lines[0] == 'SPEAKER ES2005a 1 0.000000 7.560000 <NA> <NA> %d <NA> <NA>' % (int(l[0]) + 1)
    assert len(lines) == 50


def test_merge_edge_cases():
    s, e, l = pipeline.merge_adjacent_labels([0.0, 1.0, 2.5, 2.6], [1.2, 2.0, 3.0, 4.0], np.array([0, 1, 1, 1]))
    np.testing.assert_allclose(s, [0.0, 1.1, 2.5])
    np.testing.assert_allclose(e, [1.1, 2.0, 4.0])
    assert l.tolist() == [0, 1, 1]
    s, e, l = pipeline.merge_adjacent_labels([], [], np.array([], dtype=int))
    assert len(s) == 0 and len(l) == 0


def test_soft_init_and_hard_labels_cpu(es):
    lab = torch.from_numpy(es['labels_ahc'].astype(np.int64))
    q = pipeline.soft_init(lab, int(lab.max()) + 1, float(es['smoothing']))
    assert torch.allclose(q.sum(1), torch.ones(len(lab)))
    assert torch.equal(q.argmax(1), lab)
    g = torch.from_numpy(es['gamma'])
    assert np.array_equal(pipeline.hard_labels(g).numpy(), es['labels'])
    a, b = pipeline.hard_labels(g, second=True)
    assert np.array_equal(b.numpy(), np.argsort(-es['gamma'], axis=1, kind='stable')[:, 1])


def test_xvector_transform_matches_reference_cpu(es):
    """VBx/vbhmm.py:129 (float64, CPU tensors here; the same code runs on the device)."""
    from vbx_b200 import formats
    ref_dir = '/root/reference/VBx/models/ResNet101_16kHz'
    if not os.path.exists(ref_dir):
        pytest.skip('reference model files only exist in the build container')
    mean1, mean2, lda = formats.read_xvec_transform(os.path.join(ref_dir, 'transform.h5'))
    x = pipeline.xvector_transform(torch.from_numpy(es['x_raw'].astype(np.float64)), torch.from_numpy(mean1),
                                   torch.from_numpy(mean2), torch.from_numpy(lda))
    np.testing.assert_allclose(x.numpy(), es['x_lda'], atol=1e-12)
    mu, tr, psi = formats.read_kaldi_plda(os.path.join(ref_dir, 'plda'))
    mu, tr, psi = pipeline.diagonalise_plda(mu, tr, psi)
    fea = pipeline.plda_project(x, torch.from_numpy(mu), torch.from_numpy(tr), 128)
    np.testing.assert_allclose(fea.numpy(), es['fea'], atol=1e-9)
    np.testing.assert_allclose(psi[:128], es['Phi'], rtol=1e-12)


@pytest.mark.gpu
def test_es2005a_rttm_end_to_end_on_gpu(es):
    """fea -> VB-HMM on the B200 -> labels -> merged segments == the reference's committed system output
    exp/ES2005a.rttm (50 segments, speaker ids up to renaming)."""
    from vbx_b200.batch import VbxBatch
    dev = torch.device('cuda:0')
    lab = torch.from_numpy(es['labels_ahc'].astype(np.int64)).to(dev)
    S = int(lab.max()) + 1
    T = len(lab)
    vb = VbxBatch([T], 128, S, device=dev)
    g = torch.zeros((T, vb.S), device=dev)
    g[:, :S] = pipeline.soft_init(lab, S, float(es['smoothing']))
    p = torch.zeros((1, vb.S), device=dev)
    p[0, :S] = 1.0 / S
    vb.prepare_scale(torch.from_numpy(es['fea'].astype(np.float32)).to(dev), torch.from_numpy(es['Phi'].astype(np.float32)).to(dev))
    vb.run(g, p, Fa=float(es['Fa']), Fb=float(es['Fb']), loopProb=float(es['loopProb']), maxIters=40, epsilon=1e-6)
    labels = pipeline.hard_labels(g[:, :S]).cpu().numpy()
    s, e, l = pipeline.merge_adjacent_labels(es['seg_times'][:, 0], es['seg_times'][:, 1], labels)
    np.testing.assert_allclose(s, es['rttm_starts'])
    np.testing.assert_allclose(e, es['rttm_ends'])
    mapping = {}
    for mine, ref in zip(l, es['rttm_ref_labels']):
        assert mapping.setdefault(int(mine), int(ref)) == int(ref)
    assert len(set(mapping.values())) == len(mapping) == 5
    vb.close()
