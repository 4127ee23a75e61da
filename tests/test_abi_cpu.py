"""The C-ABI library loads without a GPU and exports every symbol include/htd_b200.h declares;
argument validation (no compute) returns the documented error codes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cdll():
    from htd_b200 import build
    return ctypes.CDLL(build.build())


@pytest.fixture(scope='module')
def hooks_cdll(cdll):
    from htd_b200 import build
    return ctypes.CDLL(build.LIB_HOOKS)


def _declared(hooks=False):
    """Names the header declares: the product interface, or only the -DHTD_DEBUG_HOOKS block."""
    src = open(os.path.join(ROOT, 'include', 'htd_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    blocks = re.findall(r'#ifdef HTD_DEBUG_HOOKS(.*?)#endif', src, flags=re.S)
    if hooks:
        src = '\n'.join(blocks)
    else:
        src = re.sub(r'#ifdef HTD_DEBUG_HOOKS.*?#endif', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(htd_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(cdll):
    names = _declared()
    assert len(names) >= 17 and 'htd_pgraph_gemm' in names and 'htd_roi_align_bwd' in names
    for n in names:
        assert hasattr(cdll, n), f'{n} declared in include/htd_b200.h but not exported'


def test_measurement_hooks_are_not_in_the_product_library(cdll, hooks_cdll):
    """htd_debug_* exist only in the library built with -DHTD_DEBUG_HOOKS, which exports the
    whole product interface as well."""
    hooks = _declared(hooks=True)
    assert 'htd_debug_set_bwd_variant' in hooks and 'htd_debug_set_option' in hooks
    for n in hooks:
        assert not hasattr(cdll, n), f'{n} leaked into the product library'
        assert hasattr(hooks_cdll, n)
    for n in _declared():
        assert hasattr(hooks_cdll, n)


def test_python_binding_covers_the_header():
    from htd_b200 import _lib
    missing = set(_declared()) - set(_lib.SIGNATURES) - {'htd_abi_version', 'htd_last_error',
                                                             'htd_roi_plan_rows_bound',
                                                             'htd_pgraph_max_tiles',
                                                             'htd_multiclass_nms_workspace_bytes',
                                                             'htd_multiclass_soft_nms_workspace_bytes',
                                                             'htd_dense_gemm_workspace_bytes',
                                                             'htd_ba_mlp_supported',
                                                             'htd_ba_mlp_workspace_floats',
                                                             'htd_roi_align_bwd_uses_tensor_pipe'}
    assert not missing, missing


def test_argument_validation_without_gpu(cdll):
    cdll.htd_last_error.restype = ctypes.c_char_p
    assert cdll.htd_abi_version() == 1
    # K > 0 with null pointers -> HTD_ERR_INVALID_ARGUMENT (1), message set, nothing launched
    rc = cdll.htd_level_assign(None, 4, 4, ctypes.c_float(56.0), None, None)
    assert rc == 1 and b'null' in cdll.htd_last_error()
    rc = cdll.htd_pgraph_plan(None, None, 4, 40, 4, 64, 0, None, None, None, None, None, None)
    assert rc == 1 and b'bad sizes' in cdll.htd_last_error()
    # empty inputs are fine
    assert cdll.htd_level_assign(None, 0, 4, ctypes.c_float(56.0), None, None) == 0


def test_struct_layout_matches_header():
    from htd_b200 import _lib
    assert ctypes.sizeof(_lib.HtdGemmGroup) == 48
    assert ctypes.sizeof(_lib.HtdLevel) == 24
    assert ctypes.sizeof(_lib.HtdBwdSource) == 8 * 8 + 4 * 4       # 8 pointers + K, dy_per_level, ring_edge, addvec_dtype


def test_backward_kernel_choice_is_a_host_decision(cdll, hooks_cdll):
    """htd_roi_align_bwd_uses_tensor_pipe: bf16 dy with C a multiple of 64 in 64..256 takes the
    tensor-pipe gather unless it is switched off; everything else the exact scalar gather."""
    f = cdll.htd_roi_align_bwd_uses_tensor_pipe
    F32, BF16 = 0, 1
    assert f(256, 7, BF16) == 1 and f(64, 7, BF16) == 1 and f(128, 8, BF16) == 1
    assert f(256, 7, F32) == 0                     # fp32 gradients: exact kernel
    assert f(96, 7, BF16) == 0 and f(320, 7, BF16) == 0 and f(32, 7, BF16) == 0
    fh = hooks_cdll.htd_roi_align_bwd_uses_tensor_pipe
    hooks_cdll.htd_debug_set_bwd_variant(0)
    try:
        assert fh(256, 7, BF16) == 0
    finally:
        hooks_cdll.htd_debug_set_bwd_variant(-1)
    assert fh(256, 7, BF16) == 1


def test_backward_source_validation_without_gpu(cdll):
    """A malformed add-vector dtype, and a bf16 add vector on a path that cannot take it, are
    rejected before anything is launched."""
    from htd_b200 import _lib
    cdll.htd_last_error.restype = ctypes.c_char_p
    lv = (_lib.HtdLevel * 1)()
    lv[0].data, lv[0].H, lv[0].W, lv[0].spatial_scale = 0x1000, 8, 8, 0.25   # never dereferenced
    src = (_lib.HtdBwdSource * 1)()
    q = src[0]
    q.rois = q.boxes = q.offsets = q.ranges = q.weights = q.dy = q.addvec = 0x1000
    q.K, q.dy_per_level, q.ring_edge = 4, 0, -1
    q.addvec_dtype = 7
    rc = cdll.htd_roi_align_bwd_multi(lv, 1, 1, 256, 1, 0, src, 1, 7, 1, None)
    assert rc == 1 and b'addvec_dtype' in cdll.htd_last_error()
    q.addvec_dtype = 1                                        # bf16 add vector with fp32 gradients
    rc = cdll.htd_roi_align_bwd_multi(lv, 1, 1, 256, 0, 0, src, 1, 7, 0, None)
    assert rc == 1 and b'bf16 addvec' in cdll.htd_last_error()
    rc = cdll.htd_roi_align_bwd_multi(lv, 1, 1, 256, 1, 0, src, 1, 8, 1, None)    # pooled == 8
    assert rc == 1 and b'bf16 addvec' in cdll.htd_last_error()


def test_product_does_not_import_the_oracle():
    """The shipped package must never route through oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, 'htd_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
            assert 'from oracle' not in src and 'import oracle' not in src, fn
