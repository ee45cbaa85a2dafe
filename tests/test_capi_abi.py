"""CPU: the C-ABI shared library loads and exports every symbol include/spatialvla_b200.h declares; argument
validation paths that run before any CUDA call behave as documented (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from spatialvla_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    L.build_library()
    return L.load_library()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "spatialvla_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(svla_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(L.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert lib.svla_abi_version() == 3
    assert lib.svla_launch_count() >= 0


def test_gemm_args_struct_matches_header_layout():
    # 11 pointers/int64 before m: a w bias colscale res_bf16 res2_bf16 res_f32 res_mod out_bf16 out_f32 out_relu
    assert L.SvlaGemmArgs.m.offset == 11 * 8
    assert L.SvlaGemmArgs.nb.offset == 17 * 8
    assert L.SvlaGemmArgs.a2.offset == 17 * 8 + 10 * 4 and C.sizeof(L.SvlaGemmArgs) == 17 * 8 + 10 * 4 + 5 * 8       # K extension appended
    assert L.SvlaGemmTnArgs.groups.offset == 56 and C.sizeof(L.SvlaTnGroup) == 40 and L.SvlaAttnBwdArgs.lse.offset == 24 * 8
    assert L.SvlaAttnArgs.batch.offset == 12 * 8
    assert L.SvlaAttnArgs.kv_start.offset == 152 and L.SvlaAttnArgs.causal_prefix.offset == 160 and L.SvlaAttnArgs.lse.offset == 168
    assert L.SvlaAttnArgs.window.offset == 184          # sliding-window predicate appended (ABI v3)
    assert C.sizeof(L.SvlaAttnArgs) == 192 and L.SvlaAttnBwdArgs.fwd_lse2.offset == 26 * 8 + 10 * 4 and C.sizeof(L.SvlaAttnBwdArgs) == 26 * 8 + 40 + 16 + 8 and L.SvlaAttnBwdArgs.window.offset == 26 * 8 + 40 + 16


def test_validation_errors_before_any_cuda_call(lib):
    g = L.SvlaGemmArgs()
    assert lib.svla_gemm(C.byref(g), None) == -1
    assert b"null operand" in lib.svla_last_error()
    a = L.SvlaAttnArgs()
    assert lib.svla_attention(C.byref(a), None) == -1
    nb_bad = (C.c_int32 * 7)(16, 32, 8, 16, 16, 100, 2)
    x = np.zeros((1, 7)); e = np.zeros(200); ids = np.zeros((1, 3), dtype=np.int32)
    assert lib.svla_tok_encode_host(x.ctypes.data, e.ctypes.data, C.cast(nb_bad, C.c_void_p), ids.ctypes.data, 1, -1.0, 1.0, 1, None, 0, 0) == -1
    assert b"bins per axis" in lib.svla_last_error()
    nb = (C.c_int32 * 7)(16, 32, 8, 16, 16, 16, 2)
    assert lib.svla_tok_encode_host(None, e.ctypes.data, C.cast(nb, C.c_void_p), None, 0, -1.0, 1.0, 1, None, 0, 0) == 0   # empty batch
    with pytest.raises(L.SvlaError):
        L.check(-1, "x")


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spatialvla_b200.ops import CudaOps
    with pytest.raises(L.SvlaError):
        CudaOps("cuda:0")
    from spatialvla_b200 import SpatialVLAForConditionalGeneration, get_config_dict
    with pytest.raises(L.SvlaError):
        SpatialVLAForConditionalGeneration(get_config_dict("tiny"), {})
    # nothing under spatialvla_b200/ imports the oracle
    pkg = os.path.join(ROOT, "spatialvla_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in re.sub(r'""".*?"""', "", open(os.path.join(pkg, fn)).read(), flags=re.S), fn
