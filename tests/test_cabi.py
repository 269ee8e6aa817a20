"""The C-ABI library loads on a CPU-only box and exports every symbol include/bm.h
declares; without a GPU it refuses to create a context instead of falling back."""
import os
import re

import pytest

from boltzmann_machines import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'bm.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bm_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _native.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert b'sm_100a' in lib.bm_version()


def test_no_cpu_fallback_without_gpu():
    if _native.device_count() > 0:
        pytest.skip('a GPU is visible')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        _native.Context()
    with pytest.raises(RuntimeError):
        from boltzmann_machines.rbm import BernoulliRBM
        BernoulliRBM(n_visible=4, n_hidden=3, model_path='/tmp/_bm_nofallback/').init()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, 'boltzmann-machines_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(d, f)
