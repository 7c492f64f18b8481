"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/xnrs_b200.h declares."""
import ctypes
import os
import re
import subprocess

from xnrs_b200 import _lib


def test_header_prototypes_parse():
    protos = _lib.parse_header()
    assert len(protos) >= 35
    for must in ('xnrs_gemm', 'xnrs_addpool_fwd', 'xnrs_addpool_bwd', 'xnrs_mha_fwd', 'xnrs_mha_bwd', 'xnrs_gru_fwd',
                 'xnrs_score_loss', 'xnrs_infonce_rows', 'xnrs_eval_impressions', 'xnrs_adam_step',
                 'xnrs_gather_rows', 'xnrs_expand_titles'):
        assert must in protos
    # every prototype ends with the stream argument (no hidden streams), except the host-side queries / switches
    for name, (_, args) in protos.items():
        if name not in ('xnrs_version', 'xnrs_last_error', 'xnrs_launch_count', 'xnrs_device_is_sm100', 'xnrs_set_option',
                        'xnrs_last_gemm_kernel', 'xnrs_gemm_simt_fallbacks'):
            assert args and args[-1] is ctypes.c_void_p, name


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = _lib.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    for name in _lib.parse_header():
        assert hasattr(handle, name), f'{name} declared in the header but not exported'
    lib = _lib.lib()
    assert lib.xnrs_version() == 100
    assert lib.xnrs_launch_count() >= 0


def test_library_contains_sm100a_code_only():
    out = subprocess.run(['cuobjdump', '-lelf', _lib.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_product_never_imports_the_oracle():
    root = os.path.join(_lib.ROOT, 'xnrs_b200')
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
