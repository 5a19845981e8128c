"""Minimal h5py for the reference driver: `with h5py.File(path, 'r') as f: np.array(f['mean1'])` (VBx/vbhmm.py:125-128)."""
from vbx_b200.formats import read_xvec_transform


class File:
    def __init__(self, path, mode='r'):
        mean1, mean2, lda = read_xvec_transform(path)
        self._d = {'mean1': mean1, 'mean2': mean2, 'lda': lda}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __getitem__(self, key):
        return self._d[key]
