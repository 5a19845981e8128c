"""Property tests of the host-side logic either side of the GPU path (hypothesis, CPU only): sharding over ranks, the
choice of sub-batches, label compaction, the DER diagnostic, archive / model file round trips.  The properties are the
size-independent ones the domain offers: every recording on exactly one rank, balance inside the LPT bound, idempotence
of the compaction, permutation invariance of the scoring, write -> read = identity."""
import os
import tempfile

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from vbx_b200 import formats, shard
from vbx_b200.api import DER
from vbx_b200.parts import auto_parts
from vbx_b200.pipeline import merge_adjacent_labels, rttm_lines

settings.register_profile('repo', derandomize=True, deadline=None, database=None)    # same examples on every run, no .hypothesis/ directory
settings.load_profile('repo')

LENGTHS = st.lists(st.integers(min_value=1, max_value=20000), min_size=0, max_size=300)


@settings(max_examples=200, deadline=None)
@given(LENGTHS, st.integers(min_value=1, max_value=8))
def test_partition_covers_every_recording_once_and_is_balanced(lengths, world):
    shards = shard.partition(lengths, world)
    assert len(shards) == world
    flat = [b for s in shards for b in s]
    assert sorted(flat) == list(range(len(lengths)))                  # every recording on exactly one rank
    assert all(s == sorted(s) for s in shards)                        # shards keep the caller's order
    loads = [int(sum(lengths[b] for b in s)) for s in shards]
    if lengths:
        # greedy LPT: the heaviest rank exceeds the lightest by at most one recording (the last one it received)
        assert max(loads) - min(loads) <= max(lengths)
        # Graham's list-scheduling bound: the heaviest rank carries at most the mean load plus (1 - 1/m) of one recording
        assert max(loads) <= sum(lengths) / world + max(lengths) * (1.0 - 1.0 / world) + 1e-9
    assert shard.partition(lengths, world) == shards                  # deterministic


@settings(max_examples=100, deadline=None)
@given(LENGTHS)
def test_one_rank_owns_everything(lengths):
    assert shard.partition(lengths, 1) == [list(range(len(lengths)))]


def test_partition_is_independent_of_the_rank_that_asks():
    rng = np.random.default_rng(3)
    lengths = rng.integers(200, 12000, size=192)
    views = [shard.partition(lengths, 8) for _ in range(8)]          # every rank computes the same plan locally
    assert all(v == views[0] for v in views)
    loads = np.array([lengths[s].sum() for s in views[0]])
    assert loads.max() / loads.mean() < 1.01                         # config 4: 192 recordings over 8 GPUs, < 1 % imbalance


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 6000), st.integers(1, 5000), st.integers(1, 64))
def test_auto_parts_only_cuts_large_batches_of_short_recordings(n_rec, t, s):
    lengths = np.full(n_rec, t)
    parts = auto_parts(lengths, s)
    assert parts in (1, 2)
    if parts == 2:
        assert n_rec >= 2048 and n_rec * t >= 2_000_000 and t < 4096
    if n_rec < 2048 or t >= 4096:
        assert parts == 1


# ---------------------------------------------------------------------------------------------------------------------
# label compaction (VBx/diarization_lib.py:113-135 restated as the sequential procedure it describes)
# ---------------------------------------------------------------------------------------------------------------------
def compact_sequentially(starts, ends, labels):
    out = []
    for s, e, l in zip(starts, ends, labels):
        if out and l == out[-1][2] and (np.isclose(out[-1][1], s) or out[-1][1] > s):
            out[-1][1] = e
        else:
            out.append([s, e, l])
    for a, b in zip(out[:-1], out[1:]):
        if b[0] < a[1]:
            a[1] = b[0] = (a[1] + b[0]) / 2.0
    return (np.array([o[0] for o in out], dtype=np.float64), np.array([o[1] for o in out], dtype=np.float64),
            np.array([o[2] for o in out], dtype=np.int64))


@st.composite
def segmentations(draw):
    n = draw(st.integers(0, 60))
    step = draw(st.lists(st.sampled_from([0.0, 0.12, 0.24, 0.5, 1.5]), min_size=n, max_size=n))
    dur = draw(st.lists(st.sampled_from([0.24, 0.36, 1.44, 2.0]), min_size=n, max_size=n))
    lab = draw(st.lists(st.integers(0, 3), min_size=n, max_size=n))
    starts = np.cumsum(np.array(step, dtype=np.float64)) if n else np.zeros(0)
    ends = starts + np.array(dur, dtype=np.float64) if n else np.zeros(0)
    ends = np.maximum.accumulate(ends) if n else ends                 # the x-vector windows of a recording advance monotonically
    return starts, ends, np.array(lab, dtype=np.int64)


@settings(max_examples=300, deadline=None)
@given(segmentations())
def test_label_compaction_matches_the_sequential_procedure_and_is_idempotent(seg):
    starts, ends, labels = seg
    s, e, l = merge_adjacent_labels(starts, ends, labels)
    s2, e2, l2 = compact_sequentially(starts, ends, labels)
    np.testing.assert_allclose(s, s2)
    np.testing.assert_allclose(e, e2)
    np.testing.assert_array_equal(l, l2)
    assert len(l) <= len(labels)
    if len(l):
        assert np.all(e[:-1] <= s[1:] + 1e-12)                        # no overlap is left
        assert s[0] == starts[0] and e[-1] == ends[-1]                # the covered span is unchanged
    s3, e3, l3 = merge_adjacent_labels(s, e, l)                       # compacting a compact segmentation changes nothing
    np.testing.assert_allclose(s3, s)
    np.testing.assert_allclose(e3, e)
    np.testing.assert_array_equal(l3, l)
    lines = rttm_lines('rec', s, e, l)
    assert len(lines) == len(l) and all(x.startswith('SPEAKER rec 1 ') for x in lines)


# ---------------------------------------------------------------------------------------------------------------------
# DER diagnostic (VBx/VBx.py:129-143)
# ---------------------------------------------------------------------------------------------------------------------
@settings(max_examples=100, deadline=None)
@given(st.integers(1, 200), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_der_properties(n_frames, n_spk, seed):
    rng = np.random.default_rng(seed)
    S = n_spk + int(rng.integers(0, 3))
    ref = rng.integers(0, n_spk, size=n_frames)
    q = rng.dirichlet(np.ones(S), size=n_frames)
    d = DER(q, ref)
    assert -1e-12 <= d <= 1.0 + 1e-12
    perm = rng.permutation(S)
    assert abs(DER(q[:, perm], ref) - d) < 1e-12                      # the mapping is optimised: state order is irrelevant
    onehot = np.zeros((n_frames, S))
    onehot[np.arange(n_frames), ref] = 1.0
    assert abs(DER(onehot, ref)) < 1e-12                              # perfect posteriors
    assert abs(DER(onehot[:, perm], ref)) < 1e-12
    assert DER(q, ref, expected=False) >= -1e-12
    assert DER(onehot, ref, xentropy=True) < 1e-9
    assert DER(q, ref, xentropy=True) >= DER(onehot, ref, xentropy=True)
    with pytest.raises(ValueError):
        DER(q[:-1] if n_frames > 1 else np.zeros((2, S)), ref)


# ---------------------------------------------------------------------------------------------------------------------
# files: write -> read is the identity
# ---------------------------------------------------------------------------------------------------------------------
@settings(max_examples=50, deadline=None)
@given(st.integers(0, 20), st.integers(1, 64), st.integers(0, 2 ** 31 - 1))
def test_ark_round_trip(n, dim, seed):
    rng = np.random.default_rng(seed)
    keys = [f'rec{(i // 3):02d}_{i:04d}-{i * 24:08d}-{i * 24 + 144:08d}' for i in range(n)]
    vecs = rng.standard_normal((n, dim)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'x.ark')
        formats.write_vec_flt_ark(path, keys, vecs)
        back = list(formats.read_vec_flt_ark(path))
        assert [k for k, _ in back] == keys
        for (_, v), w in zip(back, vecs):
            assert v.dtype == np.float32
            np.testing.assert_array_equal(v, w)
        by_rec = formats.read_xvectors_by_recording(path)
        assert sum(len(k) for k, _ in by_rec.values()) == n
        for rec, (ks, x) in by_rec.items():
            assert all(k.rsplit('_', 1)[0] == rec for k in ks)        # VBx/vbhmm.py:119
            assert x.shape == (len(ks), dim)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 24), st.integers(0, 2 ** 31 - 1))
def test_plda_round_trips_binary_and_text(dim, seed):
    rng = np.random.default_rng(seed)
    mean = rng.standard_normal(dim)
    tr = rng.standard_normal((dim, dim))
    psi = np.abs(rng.standard_normal(dim)) + 0.1
    with tempfile.TemporaryDirectory() as d:
        pb, pt = os.path.join(d, 'plda'), os.path.join(d, 'plda.txt')
        formats.write_kaldi_plda_binary(pb, mean, tr, psi)
        formats.write_kaldi_plda_text(pt, mean, tr, psi)
        mb, tb, sb = formats.read_kaldi_plda(pb)
        mt, tt, s_t = formats.read_kaldi_plda(pt)
    np.testing.assert_array_equal(mb, mean)                           # binary: exact float64
    np.testing.assert_array_equal(tb, tr)
    np.testing.assert_array_equal(sb, psi)
    np.testing.assert_allclose(mt, mean, rtol=1e-6, atol=1e-9)        # text: as many digits as the writer prints
    np.testing.assert_allclose(tt, tr, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(s_t, psi, rtol=1e-6, atol=1e-9)


@settings(max_examples=50, deadline=None)
@given(segmentations())
def test_rttm_round_trip(seg):
    starts, ends, labels = seg
    s, e, l = merge_adjacent_labels(starts, ends, labels)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'r.rttm')
        with open(path, 'w') as fp:
            formats.write_rttm(fp, 'rec', l, s, e)
        back = formats.read_rttm(path)
    assert len(back) == len(l)
    for row, (a, b, c) in zip(back, zip(s, e, l)):
        assert row[0] == 'rec'
        assert abs(row[1] - a) < 1e-6 and abs(row[2] - (b - a)) < 1e-6 and int(row[3]) == int(c) + 1
