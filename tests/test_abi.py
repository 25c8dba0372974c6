"""The C ABI: libnrpb200.so loads without a GPU and exports every function include/nrp_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'nrp_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nrp_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_the_documented_surface():
    names = _declared_functions()
    for required in ('nrp_ctx_create', 'nrp_ctx_destroy', 'nrp_bg_set', 'nrp_load_parts', 'nrp_build', 'nrp_result_counts',
                     'nrp_result_flags', 'nrp_result_groups', 'nrp_result_last_single', 'nrp_result_edges',
                     'nrp_result_timings', 'nrp_last_error'):
        assert required in names


def test_library_exports_every_declared_symbol():
    from nrpcalc_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert lib.nrp_version() >= 100


def test_binding_matches_header():
    from nrpcalc_b200 import _native
    lib = _native.load_library()
    for name in _declared_functions():
        assert getattr(lib, name).argtypes is not None or name in ('nrp_version', 'nrp_last_error'), name


def test_no_gpu_means_loud_failure(scratch_cwd):
    """Without a CUDA device the product must raise, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is visible')
    from nrpcalc_b200 import _native
    with pytest.raises(_native.NativeError):
        _native.Context(0)
    import nrpcalc_b200
    with pytest.raises((RuntimeError, _native.NativeError)):
        nrpcalc_b200.finder(['ACGTTGCATGCATGCAAAGTCC', 'TTGCATGCATGCAAAGTCCAAA'], 12, verbose=False)


def test_tail_library_exports_every_declared_symbol():
    """libnrptail.so (host tail, plain C++) exports what include/nrp_tail.h declares."""
    from nrpcalc_b200 import native_tail
    if not os.path.exists(native_tail.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    text = open(os.path.join(ROOT, 'include', 'nrp_tail.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = sorted(set(re.findall(r'\b(nrp_tail_[a-z_0-9]+)\s*\(', text)))
    assert len(names) == 10
    lib = ctypes.CDLL(native_tail.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), name
