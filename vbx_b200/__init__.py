"""vbx_b200: the VB-HMM EM loop of BUTSpeechFIT/VBx (VBx/VBx.py) as hand-written sm_100a CUDA kernels
behind a C ABI (include/vbx_b200.h), with a drop-in `VBx()` and a batched `VbxBatch` / `vbx_batch`."""
from ._lib import VbxError, LIB_PATH  # noqa: F401


def __getattr__(name):
    # torch is imported lazily so that `import vbx_b200.formats` etc. stay light
    if name in ('VBx', 'DER'):
        from . import api
        return getattr(api, name)
    if name in ('VbxBatch', 'vbx_batch'):
        from . import batch
        return getattr(batch, name)
    raise AttributeError(name)
