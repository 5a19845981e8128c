"""kaldi_io.kaldi_io private readers imported by VBx/kaldi_utils.py:22 (text / compressed matrices)."""
import numpy as np


def _read_mat_ascii(fd):
    rows = []
    while True:
        line = fd.readline()
        if not line:
            raise ValueError('unterminated text matrix')
        line = line.decode() if isinstance(line, bytes) else line
        if not line.strip():
            continue
        parts = line.strip().split()
        last = parts[-1] == ']'
        if last:
            parts = parts[:-1]
        if parts:
            rows.append(np.array(parts, dtype='float32'))
        if last:
            return np.vstack(rows) if rows else np.zeros((0, 0), dtype='float32')


def _read_compressed_mat(fd, header):
    raise NotImplementedError('compressed Kaldi matrices are not used by the shipped models')
