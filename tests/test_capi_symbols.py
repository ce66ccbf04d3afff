"""The C-ABI library loads without a GPU and exports every symbol include/vmadpm.h declares
(no compute calls here); the product path refuses to run without a CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'vmadpm.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vpm_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound():
    from vmad_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 40
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in vmadpm.h but not exported: %s" % missing
    unbound = [n for n in names if n not in _capi.SIGNATURES]
    assert not unbound, "declared in vmadpm.h but no ctypes signature: %s" % unbound
    extra = [n for n in _capi.SIGNATURES if n not in names]
    assert not extra, "bound in _capi.py but not declared in vmadpm.h: %s" % extra


def test_version_and_error_channel():
    from vmad_b200 import _capi
    assert _capi.lib.vpm_version() >= 100
    assert _capi.launch_count() >= 0
    # NULL context -> non-zero status and a message, never a crash
    st = _capi.lib.vpm_sync(None)
    assert st != 0 and 'NULL' in _capi.last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from vmad_b200 import _capi
    from vmad_b200.pm import ParticleMesh
    with pytest.raises(_capi.VpmError):
        ParticleMesh([8, 8, 8], 8.0)
    h = ctypes.c_void_p()
    assert _capi.lib.vpm_ctx_create(ctypes.byref(h), 0, None) != 0
    assert 'no CPU path' in _capi.last_error()


def test_product_does_not_import_the_oracle():
    """nothing under vmad_b200/ may import, call or link the oracle"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'vmad_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(import|from)\s+oracle\b', src, flags=re.M) or 'libcic_oracle' in src:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
