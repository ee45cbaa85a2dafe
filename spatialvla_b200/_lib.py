"""ctypes binding of libspatialvla_b200.so (the C ABI declared in include/spatialvla_b200.h).

The product path has NO fallback: if the shared library is missing or a symbol is absent the import of the
compute path raises.  `build_library()` is what `__graft_entry__.build()` calls (nvcc, sm_100a only).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "lib", "libspatialvla_b200.so")
SOURCES = ["capi.cu", "gemm_tcgen05.cu", "gemm_skinny.cu", "attention.cu", "attention_tc.cu", "attention_bwd_tc.cu", "decode_mega.cu",
           "fused_ops.cu", "tokenizer.cu", "image_ops.cu", "train_ops.cu", "train_mma.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_RELU, ACT_SOFTCAP, ACT_SOFTPLUS = range(6)
GEMM_GEGLU, GEMM_ACCUM_F32, GEMM_CONV3X3 = 1, 2, 4


class SvlaGemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("colscale", C.c_void_p),
        ("res_bf16", C.c_void_p), ("res2_bf16", C.c_void_p), ("res_f32", C.c_void_p), ("res_mod", C.c_int64),
        ("out_bf16", C.c_void_p), ("out_f32", C.c_void_p), ("out_relu_bf16", C.c_void_p),
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldo", C.c_int64),
        ("nb", C.c_int32), ("h", C.c_int32), ("wd", C.c_int32), ("c", C.c_int32),
        ("alpha", C.c_float), ("act_param", C.c_float), ("act", C.c_int32), ("flags", C.c_int32),
        ("block_n", C.c_int32), ("impl", C.c_int32),
        ("a2", C.c_void_p), ("w2", C.c_void_p), ("k2", C.c_int64), ("lda2", C.c_int64), ("ldw2", C.c_int64),
    ]


class SvlaSkinnyArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("out_bf16", C.c_void_p), ("out_f32", C.c_void_p),
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64), ("ldx", C.c_int64), ("ldw", C.c_int64), ("ldo", C.c_int64),
        ("partial_stride", C.c_int64), ("alpha", C.c_float), ("act_param", C.c_float), ("act", C.c_int32),
        ("flags", C.c_int32), ("splits", C.c_int32),
    ]


class SvlaAttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_bs", C.c_int64), ("q_ss", C.c_int64), ("k_bs", C.c_int64), ("k_ss", C.c_int64),
        ("v_bs", C.c_int64), ("v_ss", C.c_int64), ("o_bs", C.c_int64), ("o_ss", C.c_int64),
        ("batch", C.c_int32), ("hq", C.c_int32), ("hkv", C.c_int32), ("sq", C.c_int32), ("sk", C.c_int32),
        ("d", C.c_int32), ("scale", C.c_float), ("softcap", C.c_float), ("causal", C.c_int32),
        ("relpos_table", C.c_void_p), ("relpos_win", C.c_int32), ("relpos_head_major", C.c_int32),
        ("kv_start", C.c_void_p), ("causal_prefix", C.c_int32), ("lse", C.c_void_p), ("lse_stride", C.c_int64),
        ("window", C.c_int32),
    ]


class SvlaAttnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("dout", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("q_bs", C.c_int64), ("q_ss", C.c_int64), ("k_bs", C.c_int64), ("k_ss", C.c_int64), ("v_bs", C.c_int64), ("v_ss", C.c_int64),
        ("o_bs", C.c_int64), ("o_ss", C.c_int64), ("do_bs", C.c_int64), ("do_ss", C.c_int64), ("dq_bs", C.c_int64), ("dq_ss", C.c_int64),
        ("dk_bs", C.c_int64), ("dk_ss", C.c_int64), ("dv_bs", C.c_int64), ("dv_ss", C.c_int64),
        ("lse", C.c_void_p), ("delta", C.c_void_p),
        ("batch", C.c_int32), ("hq", C.c_int32), ("hkv", C.c_int32), ("sq", C.c_int32), ("sk", C.c_int32), ("d", C.c_int32),
        ("scale", C.c_float), ("softcap", C.c_float), ("causal", C.c_int32), ("causal_prefix", C.c_int32),
        ("fwd_lse2", C.c_void_p), ("lse_stride", C.c_int64), ("window", C.c_int32),
    ]


class SvlaTnGroup(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("ld", C.c_int64), ("row0", C.c_int32), ("rows", C.c_int32), ("col_start", C.c_int32),
                ("col_stride", C.c_int32), ("ncols", C.c_int32), ("pad", C.c_int32)]


class SvlaGemmTnArgs(C.Structure):
    _fields_ = [("s", C.c_void_p), ("y", C.c_void_p), ("m", C.c_int64), ("lds", C.c_int64), ("ldy", C.c_int64),
                ("r", C.c_int32), ("n", C.c_int32), ("scale", C.c_float), ("n_groups", C.c_int32), ("groups", SvlaTnGroup * 4)]


_P, _I, _L, _F, _D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes): every symbol include/spatialvla_b200.h declares
SIGNATURES = {
    "svla_last_error": (C.c_char_p, []),
    "svla_abi_version": (_I, []),
    "svla_launch_count": (C.c_longlong, []),
    "svla_gemm": (_I, [C.POINTER(SvlaGemmArgs), _P]),
    "svla_gemm_skinny": (_I, [C.POINTER(SvlaSkinnyArgs), _P]),
    "svla_gemm_skinny_splits": (_I, [_L, _L]),
    "svla_attention": (_I, [C.POINTER(SvlaAttnArgs), _P]),
    "svla_decode_attention": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _P, _P]),
    "svla_decode_attention_fused": (_I, [_P, _I, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P, _P]),
    "svla_layernorm": (_I, [_P, _P, _P, _F, _L, _I, _P, _P, _I, _P]),
    "svla_rmsnorm_residual": (_I, [_P, _P, _P, _P, _F, _L, _I, _P, _I, _L, _P]),
    "svla_rmsnorm_residual_hilo": (_I, [_P, _P, _P, _P, _F, _L, _I, _P, _P, _I, _L, _P]),
    "svla_decode_attention_fused_ex": (_I, [_P, _I, _L, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P, _I, _P]),
    "svla_decode_mega_supported": (_I, [_I, _I, _I, _I, _I, _I, _I]),
    "svla_decode_mega_scratch_bytes": (_L, [_I, _I, _I, _I, _I]),
    "svla_decode_mega_maps_bytes": (_L, [_I]),
    "svla_decode_mega_plan": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "svla_decode_mega_step": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P]),
    "svla_rope_kv": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _I, _L, _P, _P]),
    "svla_embed_tokens": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _L, _L, _L, _L, _I, _F, _P, _P]),
    "svla_argmax_rows": (_I, [_P, _L, _L, _L, _L, _P, _L, _P]),
    "svla_cross_entropy_rows": (_I, [_P, _L, _L, _L, _P, _L, _P, _P, _L, _P, _P]),
    "svla_adamw_step": (_I, [_P, _P, _P, _P, _L, _D, _D, _D, _D, _D, _L, _D, _P, _D, _P]),
    "svla_sumsq": (_I, [_P, _L, _P, _P]),
    "svla_cross_entropy_bwd": (_I, [_P, _L, _L, _L, _P, _L, _P, _L, _P, _F, _P, _L, _P]),
    "svla_rmsnorm_train_fwd": (_I, [_P, _P, _P, _P, _F, _L, _I, _P, _P, _P]),
    "svla_rmsnorm_bwd": (_I, [_P, _P, _P, _I, _P, _F, _L, _I, _P, _P, _P]),
    "svla_layernorm_bwd": (_I, [_P, _P, _P, _P, _F, _L, _I, _I, _P, _P, _P, _P]),
    "svla_geglu_fwd": (_I, [_P, _P, _L, _L, _P]),
    "svla_geglu_bwd": (_I, [_P, _P, _P, _L, _L, _P]),
    "svla_gelu_tanh_fwd": (_I, [_P, _P, _L, _P]),
    "svla_gelu_tanh_bwd": (_I, [_P, _P, _P, _L, _P]),
    "svla_rope_bwd": (_I, [_P, _I, _I, _I, _I, _I, _F, _P]),
    "svla_rows_cast": (_I, [_P, _P, _F, _L, _I, _P, _P]),
    "svla_lora_pack": (_I, [_P, _P, _P, _I, _I, _P]),
    "svla_fill_zero": (_I, [_P, _L, _P]),
    "svla_attention_bwd": (_I, [C.POINTER(SvlaAttnBwdArgs), _P]),
    "svla_gemm_tn": (_I, [C.POINTER(SvlaGemmTnArgs), _P]),
    "svla_siglip_patchify": (_I, [_P, _P, _I, _I, _P]),
    "svla_zoe_patchify": (_I, [_P, _P, _I, _P]),
    "svla_image_preprocess": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P]),
    "svla_barycentric_gather": (_I, [_P, _L, _P, _P, _P, _L, _I, _P]),
    "svla_beit_assemble": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "svla_readout_concat": (_I, [_P, _P, _I, _I, _I, _P]),
    "svla_pixel_shuffle": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "svla_im2col3x3_s2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "svla_bilinear_nhwc": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "svla_relu_bf16": (_I, [_P, _P, _L, _P]),
    "svla_zoe_router_embed": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "svla_zoe_attractor": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "svla_softplus_f32": (_I, [_P, _P, _L, _P]),
    "svla_zoe_select_head": (_I, [_P, _I, _I, _I, _P, _P, _L, _P, _P]),
    "svla_zoe_depth_tail": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "svla_zoe_depth_tail_fused": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
    "svla_ego3d_encode": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _P]),
    "svla_tok_encode": (_I, [_P, _P, _P, _P, _L, _D, _D, _I, _P, _I, _I, _P]),
    "svla_tok_decode": (_I, [_P, _P, _P, _L, _P, _L, _I, _P, _P]),
    "svla_tok_encode_host": (_I, [_P, _P, _P, _P, _L, _D, _D, _I, _P, _I, _I]),
    "svla_tok_decode_host": (_I, [_P, _P, _P, _L, _P, _L, _I, _P]),
}


class SvlaError(RuntimeError):
    pass


def build_library(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library (no GPU needed).  One object per source
    (compiled in parallel, rebuilt only when the source or a shared header is newer), then one link step."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    objdir = os.path.join(os.path.dirname(LIB_PATH), "obj")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")] + [os.path.join(_HERE, "..", "include", "spatialvla_b200.h")]
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    src_time = max(os.path.getmtime(os.path.join(CSRC, s)) for s in SOURCES)
    if os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(hdr_time, src_time):
        return LIB_PATH                  # e.g. on the GPU box: the prebuilt library travels, the object files do not
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        s, o = os.path.join(CSRC, src), os.path.join(objdir, src.replace(".cu", ".o"))
        if os.path.exists(o) and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_time):
            return o, False
        cmd = [nvcc] + flags + ["-c", "-o", o, s]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return o, True

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load_library():
    """Load the C-ABI library, bind and type every declared symbol. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvlaError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the spatialvla_b200 compute path has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.svla_abi_version() != 3:
        raise SvlaError("libspatialvla_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load_library().svla_last_error()
        raise SvlaError(f"{what or 'svla call'} failed ({rc}): {msg.decode() if msg else ''}")
