"""CPU-side checks of the boundary: the shared library loads and exports every symbol the header declares,
the drop-in keeps the reference's signature, and the product path refuses to run without a GPU."""
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from vbx_b200 import build
    build.build_library()
    import vbx_b200._lib as L
    return L.load()


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'vbx_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vbx_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol(lib):
    import vbx_b200._lib as L
    names = header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/vbx_b200.h but not exported'
    assert sorted(L.EXPORTS) == names
    assert lib.vbx_version().startswith(b'vbx_b200')


def test_padded_states(lib):
    got = [lib.vbx_padded_states(n) for n in (1, 3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 64)]
    assert got == [4, 4, 4, 8, 8, 16, 16, 32, 32, 32, 64, 64]
    assert lib.vbx_padded_states(0) == -1 and lib.vbx_padded_states(65) == -1


def test_dropin_signature_is_the_references():
    """VBx/VBx.py:27-29 - names, order and defaults."""
    from vbx_b200.api import VBx
    sig = inspect.signature(VBx)
    want = [('X', inspect._empty), ('Phi', inspect._empty), ('loopProb', 0.9), ('Fa', 1.0), ('Fb', 1.0), ('pi', 10),
            ('gamma', None), ('maxIters', 10), ('epsilon', 1e-4), ('alphaQInit', 1.0), ('ref', None),
            ('plot', False), ('return_model', False), ('alpha', None), ('invL', None)]
    assert [(p.name, p.default) for p in sig.parameters.values()] == want


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from vbx_b200 import VBx, VbxError
    from vbx_b200.batch import VbxBatch
    with pytest.raises(VbxError):
        VbxBatch([10], 128, 4)
    with pytest.raises(VbxError):
        VBx(np.zeros((10, 128)), np.ones(128), pi=3, gamma=np.full((10, 3), 1 / 3))


def test_create_without_device_reports_status(lib):
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    assert lib.vbx_create(0, ctypes.byref(h)) == -4      # VBX_ERR_NO_DEVICE
    assert not h.value


def test_product_code_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'vbx_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` (the oracle port on the host cores) must work on a CPU-only box."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'tiny',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['unit'] == 'x-vectors/s'
    assert line['cpu_baseline']['kind'] in ('port', 'reference') and line['e2e']['h2d_bytes_per_step'] == 0
    assert 'note' not in line['config']                    # config must equal the b200 arm's for the same workload
