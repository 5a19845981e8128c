"""Minimal kaldi_io for the reference driver: what VBx/vbhmm.py:117 and VBx/kaldi_utils.py:21-22,37 use."""
from vbx_b200.formats import read_vec_flt_ark  # noqa: F401  (VBx/vbhmm.py:117)


class BadSampleSize(Exception):
    pass


class UnknownMatrixHeader(Exception):
    pass


def open_or_fd(file, mode='rb'):
    """VBx/kaldi_utils.py:37 passes a path or an open file object."""
    return open(file, mode) if isinstance(file, str) else file
