"""Minimal fastcluster for the reference driver (VBx/vbhmm.py:140-141): average linkage of a condensed distance
vector; scipy implements the same algorithm and output layout."""
from scipy.cluster.hierarchy import linkage as _linkage


def linkage(X, method='single', metric='euclidean', preserve_input=True):
    return _linkage(X, method=method, metric=metric)
