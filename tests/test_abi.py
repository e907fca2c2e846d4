# -*- coding: utf-8 -*-
"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/ssq_b200.h declares (no compute calls here)."""
import os
import re
import pytest

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, 'include', 'ssq_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(ssqb_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
    from ssqueezepy_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing symbol %s" % n
    assert sorted(_lib.SYMBOLS) == names
    assert b'sm_100a' in lib.ssqb_version()


def test_no_cpu_fallback():
    """Without a CUDA device the public API must fail loudly, never compute."""
    import numpy as np
    import torch
    import ssqueezepy_b200 as S
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    x = np.zeros(256, dtype='float32')
    for call in (lambda: S.cwt(x, 'morlet'), lambda: S.ssq_cwt(x, 'morlet'),
                 lambda: S.stft(x), lambda: S.ssq_stft(x),
                 lambda: S.ssqueeze_fast(x.astype('complex64')[None], x.astype('complex64')[None],
                                         np.array([.1, .2]), 1., gamma=1e-6)):
        with pytest.raises(RuntimeError):
            call()


def test_product_never_imports_the_oracle():
    import subprocess, sys
    code = ("import sys; import ssqueezepy_b200; "
            "bad=[m for m in sys.modules if m.split('.')[0]=='oracle']; "
            "assert not bad, bad")
    subprocess.check_call([sys.executable, '-c', code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'ssqueezepy_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src, (dirpath, f)
