"""Per-kernel parity cases: every C-ABI op (CUDA, through spatialvla_b200.ops.CudaOps) against its torch oracle
(oracle.ops_ref.RefOps) on the same seeded inputs.  Used by tests/test_kernels_gpu.py (pytest -m gpu) and by
tools/gpu_selftest.py (one subprocess per case, so one faulting kernel cannot hide the others).

Tolerances (written here, used by both): bf16 outputs: max|err| <= 1e-2 * max|ref| (one bf16 ulp is 2^-8 relative,
plus fp32 accumulation-order noise); fp32 outputs: 1e-3 * max|ref| unless stated; integer outputs: exact.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ops_ref import RefOps  # noqa: E402
from spatialvla_b200._lib import (ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_RELU, ACT_SOFTCAP, ACT_SOFTPLUS)  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
TOL_BF16, TOL_F32 = 1e-2, 1e-3


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _randn(g, *shape, dtype=F32, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def _err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    denom = max(float(ref.abs().max()), 1e-20)
    return float((got - ref).abs().max()) / denom


class Result:
    def __init__(self, name):
        self.name, self.items = name, []

    def add(self, what, err, tol):
        self.items.append((what, float(err), float(tol)))

    @property
    def ok(self):
        return all(e <= t and e == e for _, e, t in self.items)

    def __str__(self):
        return f"{self.name}: " + ", ".join(f"{w} err={e:.3e} (tol {t:.0e}){'' if e <= t else ' FAIL'}" for w, e, t in self.items)


def _both(fn, dev):
    """Run `fn(ops, to)` with the CUDA ops and the reference ops; `to` moves inputs to the op device."""
    from spatialvla_b200.ops import CudaOps
    cu = CudaOps(dev)
    rf = RefOps()
    out_c = fn(cu, lambda t: None if t is None else t.to(dev))
    torch.cuda.synchronize()
    out_r = fn(rf, lambda t: None if t is None else t.clone())
    return out_c, out_r


# ------------------------------------------------------------------------------------------------ GEMM
def gemm_case(name, M, N, K, *, lda=None, bias=False, act=ACT_NONE, act_param=0.0, colscale=False, res_bf16=False,
              res2=False, res_f32=False, res_mod=0, out=("bf16",), accumulate=False, geglu=False, alpha=1.0,
              block_n=0, conv=None, seed=0, impl=None):
    def case(dev="cuda:0"):
        g = _gen(seed)
        if conv is not None:
            nb, h, w, c = conv
            cpad = (c + 63) // 64 * 64
            a = _randn(g, nb, h, w, c, dtype=BF16)
            wt = torch.zeros(N, 9, cpad)
            wt[:, :, :c] = torch.randn(N, 9, c, generator=g) / (9 * c) ** 0.5
            wt = wt.reshape(N, 9 * cpad).to(BF16)
            m = nb * h * w
        else:
            ld = lda or K
            a = _randn(g, M, ld, dtype=BF16)          # sliced to [:, :K] after the move so the row stride survives
            wt = (_randn(g, N, K) / K ** 0.5).to(BF16)
            m = M
        ncol = N // 2 if geglu else N
        ldo = (ncol + 7) // 8 * 8 if ncol % 8 else ncol
        t = {"bias": _randn(g, N) if bias else None, "colscale": (_randn(g, N) * 0.5 + 1.0) if colscale else None,
             "res_bf16": _randn(g, m, ldo, dtype=BF16) if res_bf16 else None,
             "res2_bf16": _randn(g, m, ldo, dtype=BF16) if res2 else None,
             "res_f32": _randn(g, res_mod or m, ldo) if res_f32 else None,
             "init_f32": _randn(g, m, ldo) if accumulate else None}

        def run(ops, to):
            A, W = to(a), to(wt)
            if conv is None and A.shape[1] != K:
                A = A[:, :K]
            o_b = ops.zeros((m, ldo), BF16) if "bf16" in out else None
            o_f = (to(t["init_f32"]) if accumulate else ops.zeros((m, ldo), F32)) if "f32" in out else None
            o_r = ops.zeros((m, ldo), BF16) if "relu" in out else None
            ops.gemm(A, W, out_bf16=o_b, out_f32=o_f, out_relu=o_r, bias=to(t["bias"]), colscale=to(t["colscale"]),
                     res_bf16=to(t["res_bf16"]), res2_bf16=to(t["res2_bf16"]), res_f32=to(t["res_f32"]), res_mod=res_mod,
                     act=act, act_param=act_param, alpha=alpha, geglu=geglu, accumulate=accumulate, conv=conv,
                     block_n=block_n, impl=impl)
            return o_b, o_f, o_r
        (cb, cf, cr), (rb, rf_, rr) = _both(run, dev)
        res = Result(name)
        if cb is not None:
            res.add("bf16", _err(cb, rb), TOL_BF16)
        if cf is not None:
            res.add("f32", _err(cf, rf_), TOL_F32)
        if cr is not None:
            res.add("relu", _err(cr, rr), TOL_BF16)
        return res
    case.__name__ = name
    return case


GEMM_CASES = [
    gemm_case("gemm_basic_256", 256, 256, 128),
    gemm_case("gemm_bn128", 256, 256, 128, block_n=128),
    gemm_case("gemm_bn64", 128, 64, 64, block_n=64, out=("bf16", "f32")),
    gemm_case("gemm_bn32_n2", 64, 2, 128, out=("f32",), bias=True),
    gemm_case("gemm_n16_n40", 1000, 40, 32, out=("bf16",)),
    gemm_case("gemm_ragged_bias_gelu", 300, 200, 104, bias=True, act=ACT_GELU_TANH, out=("bf16", "f32")),
    gemm_case("gemm_gelu_erf_colscale_accum", 577, 128, 256, bias=True, act=ACT_GELU_ERF, colscale=True, out=("f32",), accumulate=True),
    gemm_case("gemm_posemb_resmod", 1024, 1152, 640, bias=True, res_f32=True, res_mod=256, out=("f32",)),
    gemm_case("gemm_geglu", 512, 1024, 256, geglu=True),
    gemm_case("gemm_softcap_tail", 64, 8194, 512, act=ACT_SOFTCAP, act_param=30.0, out=("f32",)),
    gemm_case("gemm_softplus_f32", 288, 64, 64, bias=True, act=ACT_SOFTPLUS, out=("f32",)),
    gemm_case("gemm_strided_a", 64, 128, 512, lda=512 * 7),
    gemm_case("gemm_res_bf16_relu_copy", 384, 256, 256, bias=True, res_bf16=True, res2=True, out=("bf16", "relu")),
    gemm_case("gemm_long_k", 256, 256, 9216),
    gemm_case("gemm_multiwave", 4096 + 77, 4304, 1152, bias=True, act=ACT_GELU_TANH),
    gemm_case("gemm_multiwave_bn128", 20000, 1152, 576, block_n=128, out=("f32",), accumulate=True),
    gemm_case("conv_24_c64", 0, 64, 0, conv=(2, 24, 24, 64), bias=True, act=ACT_RELU),
    gemm_case("conv_12_c128_res", 0, 64, 0, conv=(3, 12, 12, 128), bias=True, res_bf16=True, res2=True, out=("bf16", "relu")),
    gemm_case("conv_48_c32", 0, 32, 0, conv=(1, 48, 48, 32), bias=True),
    gemm_case("conv_96_c256", 0, 256, 0, conv=(2, 96, 96, 256)),
    gemm_case("conv_odd_20x28", 0, 128, 0, conv=(2, 20, 28, 64), bias=True),
]
PAIR_CASES = [   # cta_group::2 kernel forced (impl=3) and the 1-CTA kernel forced (impl=2) on the same shapes
    gemm_case("pair_basic_512", 512, 256, 128, impl=3),
    gemm_case("pair_bn128_ragged_odd_tiles", 128 * 3 + 5, 200, 104, bias=True, act=ACT_GELU_TANH, out=("bf16", "f32"), block_n=128, impl=3),
    gemm_case("pair_geglu", 1024, 1024, 256, geglu=True, impl=3),
    gemm_case("pair_multiwave_accum", 20000, 1152, 576, out=("f32",), accumulate=True, bias=True, colscale=True, impl=3),
    gemm_case("pair_long_k", 512, 512, 9216, impl=3),
    gemm_case("pair_conv_96_c256", 0, 256, 0, conv=(2, 96, 96, 256), bias=True, res_bf16=True, out=("bf16", "relu"), impl=3),
    gemm_case("pair_conv_odd_20x28", 0, 128, 0, conv=(3, 20, 28, 64), bias=True, impl=3),
    gemm_case("single_forced_multiwave", 4096 + 77, 4304, 1152, bias=True, act=ACT_GELU_TANH, impl=2),
]
TMA_EPI_CASES = [   # TMA-store epilogue (one output, no residuals): ragged M/N, partial 32-column chunk, odd pair tiles
    gemm_case("tmaepi_bf16_ragged", 300, 200, 104, bias=True, act=ACT_GELU_TANH),
    gemm_case("tmaepi_bf16_erf_colscale_alpha", 577, 1048, 256, bias=True, act=ACT_GELU_ERF, colscale=True, alpha=0.5),
    gemm_case("tmaepi_bf16_pair_odd_tiles", 128 * 5 + 17, 336, 192, bias=True, impl=3),
    gemm_case("tmaepi_bf16_pair_bn128", 128 * 4, 320, 128, block_n=128, impl=3, act=ACT_RELU),
    gemm_case("tmaepi_f32_plain_ragged", 300, 200, 104, bias=True, out=("f32",)),
    gemm_case("tmaepi_f32_softcap", 130, 8196, 512, act=ACT_SOFTCAP, act_param=30.0, out=("f32",)),
    gemm_case("tmaepi_f32_accum_ragged", 577, 1028, 256, bias=True, colscale=True, out=("f32",), accumulate=True),
    gemm_case("tmaepi_f32_accum_pair_multiwave", 36928 // 4 + 9, 1024, 1024, bias=True, colscale=True, out=("f32",), accumulate=True, impl=3),
    gemm_case("tmaepi_bf16_multiwave_4304", 16384 // 2, 4304, 1152, bias=True, act=ACT_GELU_TANH),
]
ROWTILE_CASES = [   # row-tile 3x3 conv (impl=4): shifted-descriptor taps over one 130-pixel halo row per (chunk, kernel row)
    gemm_case("rowtile_n32_w384", 0, 32, 0, conv=(1, 5, 384, 128), bias=True, act=ACT_RELU, impl=4),
    gemm_case("rowtile_n32_ragged_w150", 0, 32, 0, conv=(2, 7, 150, 64), bias=True, impl=4),
    gemm_case("rowtile_n24_w130_c72", 0, 24, 0, conv=(2, 3, 130, 72), bias=True, impl=4),
    gemm_case("rowtile_n64_w256", 0, 64, 0, conv=(2, 6, 256, 128), bias=True, act=ACT_RELU, impl=4),
    gemm_case("rowtile_n128_w192_c256", 0, 128, 0, conv=(2, 9, 192, 256), bias=True, impl=4),
    gemm_case("rowtile_n128_res_relu_copy", 0, 128, 0, conv=(1, 4, 140, 128), bias=True, res_bf16=True, res2=True, out=("bf16", "relu"), impl=4),
    gemm_case("rowtile_small_w20", 0, 32, 0, conv=(3, 20, 20, 64), bias=True, impl=4),
    gemm_case("rowtile_auto_dispatch", 0, 32, 0, conv=(1, 4, 384, 128), bias=True, act=ACT_RELU),
]
SIMT_CASES = [
    gemm_case("simt_ragged", 300, 200, 104, bias=True, act=ACT_GELU_TANH, out=("bf16", "f32"), impl=1),
    gemm_case("simt_conv", 0, 64, 0, conv=(2, 24, 24, 64), bias=True, act=ACT_RELU, impl=1),
    gemm_case("simt_geglu", 128, 256, 64, geglu=True, impl=1),
]


def skinny_case(name, M, N, K, *, mode="partial", splits=0, bias=False, act=ACT_NONE, act_param=0.0, ldx=None, seed=0, tiled=False,
                pair=False):
    def case(dev="cuda:0"):
        from spatialvla_b200.ops import tile_weight
        g = _gen(seed)
        x_full = _randn(g, M, ldx or K, dtype=BF16)
        wt = (_randn(g, N, K) / K ** 0.5).to(BF16)
        b = _randn(g, N) if bias else None

        def run(ops, to):
            X = to(x_full)[:, :K]
            W = to(tile_weight(wt)) if tiled else to(wt)
            tn = N if tiled else None
            if mode == "partial":
                S = splits or ops.skinny_splits(N, K) if ops.name == "cuda" else (splits or 1)
                out = ops.zeros((S, M, N), F32)
                ops.gemm_skinny(X, W, out_f32=out, tiled_n=tn, pair=pair)
                return (out.sum(0),)
            if mode == "geglu":
                out = ops.zeros((M, N // 2), BF16)
                ops.gemm_skinny(X, W, out_bf16=out, geglu=True, bias=to(b), tiled_n=tn, pair=pair)
                return (out,)
            of, ob = ops.zeros((M, N), F32), ops.zeros((M, N), BF16)
            ops.gemm_skinny(X, W, out_f32=of, out_bf16=ob, bias=to(b), act=act, act_param=act_param, tiled_n=tn, pair=pair)
            return of, ob
        c, r = _both(run, dev)
        res = Result(name)
        for i, (a_, b_) in enumerate(zip(c, r)):
            res.add(f"out{i}", _err(a_, b_), TOL_BF16 if a_.dtype == BF16 else TOL_F32)
        return res
    case.__name__ = name
    return case


def skinny_consumers_case(dev="cuda:0"):
    """split-K partial sums consumed by svla_rmsnorm_residual and svla_rope_kv (fp32 partial input)."""
    g = _gen(21)
    rows, cols, S = 64, 2304, 5
    x, parts = _randn(g, rows, cols), _randn(g, S, rows, cols)
    wp, wq = _randn(g, cols) * 0.1, _randn(g, cols) * 0.1
    B, hq, hkv, d, smax, pos0 = 64, 8, 4, 256, 290, 281
    qkvp = _randn(g, 3, B, (hq + 2 * hkv) * d)

    def run(ops, to):
        xx, ob = to(x), ops.zeros((rows, cols), BF16)
        ops.rmsnorm_residual(xx, branch=to(parts), w_post=to(wp), w_pre=to(wq), eps=1e-6, out_bf16=ob)
        q, kc, vc = ops.zeros((B, hq * d), BF16), ops.zeros((B, smax, hkv, d), BF16), ops.zeros((B, smax, hkv, d), BF16)
        ops.rope_kv(to(qkvp), q, kc, vc, batch=B, s=1, hq=hq, hkv=hkv, d=d, smax=smax, pos0=pos0, theta=10000.0)
        return xx, ob, q, kc, vc
    c, r = _both(run, dev)
    res = Result("skinny_consumers")
    for nm, a_, b_, tol in zip(("x", "h", "q", "kcache", "vcache"), c, r, (1e-5, TOL_BF16, TOL_BF16, TOL_BF16, TOL_BF16)):
        res.add(nm, _err(a_, b_), tol)
    return res


def _hilo_pair(v):
    """fp32 -> the hi/lo bf16 pair [2, rows, cols] of the decode chain (hi = bf16(v), lo = bf16(v - hi))"""
    hi = v.to(BF16)
    return torch.stack([hi, (v - hi.float()).to(BF16)]).contiguous()


def _hilo_sum(t):
    return t[0].float() + t[1].float()


def skinny_hilo_case(name, M, N, K, *, mode="partial", act=ACT_NONE, act_param=0.0, seed=0, tiled=False, pair=False):
    """X_HILO / OUT_HILO modes of svla_gemm_skinny: hi/lo activation pairs in, fp32 partial sums / fp32 / hi/lo pairs out.
    Checked (1) against the re-statement on the same pair and (2) against the fp64 product of the UNSPLIT fp32 activations,
    where the pair must be far inside what one bf16 plane can give (2^-9 per element)."""
    def case(dev="cuda:0"):
        from spatialvla_b200.ops import tile_weight
        g = _gen(seed)
        xf = _randn(g, M, K)
        wt = (_randn(g, N, K) / K ** 0.5).to(BF16)
        exact = (xf.double() @ wt.double().t()).float()

        def run(ops, to):
            X = to(_hilo_pair(xf))
            W = to(tile_weight(wt)) if tiled else to(wt)
            tn = N if tiled else None
            if mode == "partial":
                S = ops.skinny_splits(N, K) if ops.name == "cuda" else 1
                out = ops.zeros((S, M, N), F32)
                ops.gemm_skinny(X, W, out_f32=out, tiled_n=tn, pair=pair)
                return (out.sum(0),)
            if mode == "geglu":
                out = ops.zeros((2, M, N // 2), BF16)
                ops.gemm_skinny(X, W, out_bf16=out, geglu=True, tiled_n=tn, pair=pair)
                return (_hilo_sum(out), out[0].float())
            of, ob = ops.zeros((M, N), F32), ops.zeros((2, M, N), BF16)
            ops.gemm_skinny(X, W, out_f32=of, out_bf16=ob, act=act, act_param=act_param, tiled_n=tn, pair=pair)
            return of, _hilo_sum(ob), ob[0].float()
        c, r = _both(run, dev)
        res = Result(name)
        for i, (a_, b_) in enumerate(zip(c, r)):
            res.add(f"out{i}", _err(a_, b_), 1e-4 if i < len(c) - 1 or mode == "partial" else TOL_BF16)
        if mode == "partial":
            res.add("vs_fp64_unsplit", _err(c[0], exact), 5e-5)
        elif mode == "geglu":
            ref = torch.nn.functional.gelu(exact[:, 0::2].double(), approximate="tanh") * exact[:, 1::2].double()
            res.add("vs_fp64_unsplit", _err(c[0], ref), 5e-5)
        return res
    case.__name__ = name
    return case


def decode_hilo_consumers_case(dev="cuda:0"):
    """hi/lo outputs of the sandwich norm and of the fused decode attention (the producers of the X_HILO GEMM operands)."""
    g = _gen(23)
    rows, cols, S = 64, 2304, 5
    x, parts = _randn(g, rows, cols), _randn(g, S, rows, cols)
    wp, wq = _randn(g, cols) * 0.1, _randn(g, cols) * 0.1
    B, hq, hkv, d, smax, ctx, splits = 3, 8, 4, 256, 290, 285, 4
    part = _randn(g, splits, B, (hq + 2 * hkv) * d, scale=0.6)
    kc0, vc0 = _randn(g, B, smax, hkv, d, dtype=BF16), _randn(g, B, smax, hkv, d, dtype=BF16)
    pads = torch.tensor([0, 6, 97], dtype=torch.int32)

    def run(ops, to):
        xx, ob = to(x), ops.zeros((2, rows, cols), BF16)
        ops.rmsnorm_residual(xx, branch=to(parts), w_post=to(wp), w_pre=to(wq), eps=1e-6, out_bf16=ob)
        x1, ob1 = to(x[:1]), ops.zeros((2, 1, cols), BF16)
        ops.rmsnorm_residual(x1, w_pre=to(wq), eps=1e-6, out_bf16=ob1)
        kc, vc = to(kc0), to(vc0)
        out = ops.zeros((2, B, hq * d), BF16)
        ops.decode_attention_fused(to(part), kc, vc, out, batch=B, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, theta=10000.0,
                                   scale=1 / 16, softcap=50.0, kv_start=to(pads))
        return xx, _hilo_sum(ob), ob[0].float(), _hilo_sum(ob1), _hilo_sum(out), out[0].float(), kc[:, ctx - 1].clone()
    c, r = _both(run, dev)
    res = Result("decode_hilo_consumers")
    for nm, a_, b_, tol in zip(("x", "h_pair", "h_hi", "h_pair_1row", "attn_pair", "attn_hi", "k_row"), c, r,
                               (1e-5, 5e-5, TOL_BF16, 5e-5, 5e-4, TOL_BF16, TOL_BF16)):
        res.add(nm, _err(a_, b_), tol)
    return res


SKINNY_CASES = [
    skinny_hilo_case("skinny_hilo_qkv_partial", 64, 4096, 2304, seed=41),
    skinny_hilo_case("skinny_hilo_down_partial_m33", 33, 2304, 9216, seed=42),
    skinny_hilo_case("skinny_hilo_geglu", 64, 18432, 2304, mode="geglu", seed=43),
    skinny_hilo_case("skinny_hilo_geglu_m1_tiled", 1, 1024, 512, mode="geglu", seed=44, tiled=True),
    skinny_hilo_case("skinny_hilo_head_softcap_m64", 64, 8194, 2304, mode="plain", act=ACT_SOFTCAP, act_param=30.0, seed=45),
    skinny_hilo_case("skinny_hilo_head_m17_ragged", 17, 300, 200, mode="plain", seed=46),
    skinny_hilo_case("skinny_hilo_m32_partial", 32, 512, 1024, seed=47),
    skinny_hilo_case("skinny_hilo_m8_partial", 8, 2304, 2048, seed=56),
    skinny_hilo_case("skinny_hilo_m5_geglu", 5, 18432, 2304, mode="geglu", seed=57),
    skinny_hilo_case("skinny_hilo_m1_head_softcap", 1, 8194, 2304, mode="plain", act=ACT_SOFTCAP, act_param=30.0, seed=58),
    skinny_hilo_case("skinny_hilo_m9_partial", 9, 512, 1024, seed=59),
    # CTA pairs (cta_group::2; opt-in, measured slower on the decode chain): same results, odd tile count = one surplus CTA
    skinny_hilo_case("skinny_hilo_pair_qkv_partial", 64, 4096, 2304, seed=48, pair=True),
    skinny_hilo_case("skinny_hilo_pair_geglu_m20", 20, 18432, 2304, mode="geglu", seed=49, pair=True),
    skinny_hilo_case("skinny_hilo_pair_head_odd_tiles", 64, 8194, 2304, mode="plain", act=ACT_SOFTCAP, act_param=30.0, seed=50, pair=True, tiled=True),
    decode_hilo_consumers_case,
    skinny_case("skinny_qkv_partial", 64, 4096, 2304),
    skinny_case("skinny_o_partial_split8", 64, 2304, 2048, splits=8),
    skinny_case("skinny_down_partial", 64, 2304, 9216),
    skinny_case("skinny_geglu", 64, 18432, 2304, mode="geglu"),
    skinny_case("skinny_head_softcap_tail", 64, 8194, 2304, mode="plain", act=ACT_SOFTCAP, act_param=30.0),
    skinny_case("skinny_m1_bias", 1, 1000, 512, mode="plain", bias=True),
    skinny_case("skinny_m7_strided_x", 7, 256, 128, mode="plain", ldx=128 * 5),
    skinny_case("skinny_m100", 100, 384, 192, mode="plain", act=ACT_RELU),
    skinny_case("skinny_m128_partial", 128, 512, 1024, splits=3),
    skinny_case("skinny_tiled_qkv_partial", 64, 4096, 2304, tiled=True, seed=31),
    skinny_case("skinny_tiled_geglu", 64, 18432, 2304, mode="geglu", tiled=True, seed=32),
    skinny_case("skinny_tiled_head_ragged", 64, 8194, 2304, mode="plain", act=ACT_SOFTCAP, act_param=30.0, tiled=True, seed=33),
    skinny_case("skinny_tiled_ragged_k", 5, 300, 200, mode="plain", bias=True, tiled=True, seed=34),
    skinny_case("skinny_pair_down_partial", 64, 2304, 9216, pair=True, seed=35),
    skinny_case("skinny_pair_m100_ragged", 100, 384 + 128, 192, mode="plain", act=ACT_RELU, bias=True, pair=True, seed=36),
    skinny_case("skinny_pair_m20_geglu_odd_tiles", 20, 128 * 5, 256, mode="geglu", pair=True, seed=37),
    skinny_consumers_case,
]


# ------------------------------------------------------------------------------------------------ attention
def attn_case(name, B, hq, hkv, sq, sk, d, *, scale=None, softcap=0.0, causal=False, relpos_win=0, packed_qkv=False,
              smax=None, seed=0, head_major=False, kv_start=None, causal_prefix=0, window=0):
    def case(dev="cuda:0"):
        g = _gen(seed)
        sc = scale if scale is not None else d ** -0.5
        if packed_qkv:      # [B*S, 3*hq*d] like the ViT towers
            D = hq * d
            qkv = _randn(g, B * sq, 3 * D, dtype=BF16)
            tab = _randn(g, (2 * relpos_win - 1) ** 2 + 3, hq, scale=0.5) if relpos_win else None
            if tab is not None and head_major:
                tab = tab.t().contiguous()          # [hq, nrel]: the layout the engine packs at load time

            def run(ops, to):
                t = to(qkv)
                out = ops.zeros((B * sq, D), BF16)
                st = (sq * 3 * D, 3 * D)
                ops.attention(t, t[:, D:], t[:, 2 * D:], out, batch=B, hq=hq, hkv=hkv, sq=sq, sk=sk, d=d, q_strides=st,
                              k_strides=st, v_strides=st, o_strides=(sq * D, D), scale=sc, softcap=softcap, causal=causal,
                              relpos_table=to(tab), relpos_win=relpos_win, relpos_head_major=head_major)
                return out
        else:               # Gemma2 layout: q [B*sq, hq*d], cache [B, smax, hkv, d]
            sm = smax or sk
            q = _randn(g, B * sq, hq * d, dtype=BF16)
            kc = _randn(g, B, sm, hkv, d, dtype=BF16)
            vc = _randn(g, B, sm, hkv, d, dtype=BF16)

            def run(ops, to):
                out = ops.zeros((B * sq, hq * d), BF16)
                kvs = (sm * hkv * d, hkv * d)
                ops.attention(to(q), to(kc), to(vc), out, batch=B, hq=hq, hkv=hkv, sq=sq, sk=sk, d=d,
                              q_strides=(sq * hq * d, hq * d), k_strides=kvs, v_strides=kvs, o_strides=(sq * hq * d, hq * d),
                              scale=sc, softcap=softcap, causal=causal, causal_prefix=causal_prefix, window=window,
                              kv_start=None if kv_start is None else to(torch.tensor(kv_start, dtype=torch.int32)))
                return out
        c, r = _both(run, dev)
        res = Result(name)
        res.add("out", _err(c, r), 1.5e-2)
        return res
    case.__name__ = name
    return case


def decode_attn_case(dev="cuda:0"):
    g = _gen(3)
    B, hq, hkv, d, smax, ctx = 3, 4, 2, 256, 300, 271
    q, kc, vc = _randn(g, B, hq * d, dtype=BF16), _randn(g, B, smax, hkv, d, dtype=BF16), _randn(g, B, smax, hkv, d, dtype=BF16)

    def run(ops, to):
        out = ops.zeros((B, hq * d), BF16)
        ops.decode_attention(to(q), to(kc), to(vc), out, batch=B, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, scale=1 / 16, softcap=50.0)
        return out
    c, r = _both(run, dev)
    res = Result("decode_attention")
    res.add("out", _err(c, r), 1.5e-2)

    def run_padded(ops, to):
        out = ops.zeros((B, hq * d), BF16)
        ops.decode_attention(to(q), to(kc), to(vc), out, batch=B, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, scale=1 / 16, softcap=50.0,
                             kv_start=to(torch.tensor([0, 7, 200], dtype=torch.int32)))
        return out
    c, r = _both(run_padded, dev)
    res.add("out[left_padded]", _err(c, r), 1.5e-2)
    return res


def decode_attn_fused_case(dev="cuda:0"):
    """RoPE + cache append + attention in one launch vs the rope_kv -> decode_attention composition of the reference ops."""
    res = Result("decode_attention_fused")
    for tag, (B, hq, hkv, smax, ctx, splits) in {"gqa2_ctx271": (3, 4, 2, 300, 271, 3), "first_token": (2, 4, 2, 16, 1, 1),
                                                  "mha_ctx33": (2, 2, 2, 40, 33, 2), "chunk_edge_ctx65": (1, 8, 4, 65, 65, 4),
                                                  "ctx64": (2, 8, 4, 290, 64, 4), "left_padded_ctx285": (3, 8, 4, 290, 285, 4)}.items():
        g = _gen(ctx)
        d = 256
        W = (hq + 2 * hkv) * d
        part = _randn(g, splits, B, W, scale=0.6)
        kc0, vc0 = _randn(g, B, smax, hkv, d, dtype=BF16), _randn(g, B, smax, hkv, d, dtype=BF16)

        pads = torch.tensor([0, 6, 97], dtype=torch.int32) if tag.startswith("left_padded") else None

        def run(ops, to):
            kc, vc = to(kc0), to(vc0)
            out = ops.zeros((B, hq * d), BF16)
            ops.decode_attention_fused(to(part), kc, vc, out, batch=B, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, theta=10000.0,
                                       scale=1 / 16, softcap=50.0, kv_start=to(pads))
            return out, kc[:, ctx - 1].clone(), vc[:, ctx - 1].clone()
        (co, ck, cv), (ro, rk, rv) = _both(run, dev)
        res.add(f"out[{tag}]", _err(co, ro), 1.5e-2)
        res.add(f"k_row[{tag}]", _err(ck, rk), TOL_BF16)
        res.add(f"v_row[{tag}]", _err(cv, rv), TOL_BF16)
    return res


def decode_attn_fused_window_case(dev="cuda:0"):
    """fused decode step on a sliding-window layer: only the last `window` cache slots receive weight (plain and hi/lo outputs,
    left-padded rows whose padding lies inside / outside the window)."""
    res = Result("decode_attention_fused_window")
    for tag, (B, hq, hkv, smax, ctx, splits, window, pads) in {"w48_ctx271": (3, 8, 4, 300, 271, 3, 48, None),
                                                                "w200_pads_inside": (3, 8, 4, 290, 285, 4, 200, [0, 6, 97]),
                                                                "w1_self_only": (2, 4, 2, 64, 40, 2, 1, None),
                                                                "w_ge_ctx": (2, 2, 2, 40, 33, 2, 33, None)}.items():
        g = _gen(ctx + window)
        d = 256
        part = _randn(g, splits, B, (hq + 2 * hkv) * d, scale=0.6)
        kc0, vc0 = _randn(g, B, smax, hkv, d, dtype=BF16), _randn(g, B, smax, hkv, d, dtype=BF16)
        pt = None if pads is None else torch.tensor(pads, dtype=torch.int32)

        def run(ops, to):
            kc, vc = to(kc0), to(vc0)
            out, pair = ops.zeros((B, hq * d), BF16), ops.zeros((2, B, hq * d), BF16)
            kw = dict(batch=B, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, theta=10000.0, scale=1 / 16, softcap=50.0, kv_start=to(pt), window=window)
            ops.decode_attention_fused(to(part), kc, vc, out, **kw)
            ops.decode_attention_fused(to(part), kc, vc, pair, **kw)
            return out, _hilo_sum(pair)
        (co, cp), (ro, rp) = _both(run, dev)
        res.add(f"out[{tag}]", _err(co, ro), 1.5e-2)
        res.add(f"pair[{tag}]", _err(cp, rp), 5e-4)
    return res


ATTN_CASES = [
    attn_case("attn_siglip_d72", 2, 2, 2, 256, 256, 72, packed_qkv=True),
    attn_case("attn_beit_d64_relpos", 1, 2, 2, 577, 577, 64, packed_qkv=True, relpos_win=24),
    attn_case("attn_router_d32", 2, 4, 4, 145, 145, 32, packed_qkv=True),
    attn_case("attn_gemma_prefill_d256", 2, 4, 2, 278, 278, 256, scale=1 / 16, softcap=50.0, smax=290),
    attn_case("attn_gemma_causal_d256", 1, 2, 1, 70, 70, 256, scale=1 / 16, softcap=50.0, causal=True, smax=80),
    attn_case("attn_d128_generic", 1, 2, 2, 100, 130, 128, causal=False, smax=130),
    # tcgen05 path (attention_tc.cu): batch > 1 through the 3-D tensor maps, ragged tails, sharp softmax (lazy O
    # rescale fires), large soft-cap arguments (libm tanh branch), causal tile skipping, GQA, sq != sk
    attn_case("attn_tc_beit_batch3", 3, 4, 4, 577, 577, 64, packed_qkv=True, relpos_win=24, seed=5),
    attn_case("attn_tc_beit_head_major", 2, 5, 5, 577, 577, 64, packed_qkv=True, relpos_win=24, seed=7, head_major=True),
    attn_case("attn_tc_beit_head_major_win5", 2, 3, 3, 26, 26, 64, packed_qkv=True, relpos_win=5, seed=8, head_major=True),
    attn_case("attn_tc_d64_sharp_relpos", 2, 3, 3, 577, 577, 64, packed_qkv=True, relpos_win=24, scale=1.0, seed=6),
    attn_case("attn_tc_d64_plain_ragged", 2, 2, 2, 130, 200, 64, smax=210, seed=7),
    attn_case("attn_tc_d64_tiny", 1, 1, 1, 1, 1, 64, smax=8, seed=8),
    attn_case("attn_tc_d256_sharp_softcap", 3, 8, 4, 278, 278, 256, scale=0.5, softcap=50.0, smax=290, seed=9),
    attn_case("attn_tc_d256_causal_offset", 2, 4, 2, 200, 260, 256, scale=0.25, softcap=50.0, causal=True, smax=300, seed=10),
    # left-padded prompts: keys [0, kv_start[b]) masked for every query (tile fully / partly / not masked; pad > one tile)
    attn_case("attn_tc_d256_left_padded", 4, 4, 2, 278, 278, 256, scale=1 / 16, softcap=50.0, smax=290, seed=11, kv_start=[0, 3, 64, 131]),
    # causal continuation over a left-padded cache (sq < sk, every query still sees at least one real key)
    attn_case("attn_tc_d256_left_padded_causal", 2, 2, 1, 100, 150, 256, scale=1 / 16, softcap=50.0, causal=True, smax=160, seed=12, kv_start=[5, 50]),
    # prefix-LM mask of the training forward (model/modeling_spatialvla.py:292-305): keys < prefix visible to every query, the
    # suffix causal; prefix inside the first tile / spanning tiles / == sk (degenerates to bidirectional); L = 291 like config #5
    attn_case("attn_tc_d256_prefix_lm", 3, 4, 2, 291, 291, 256, scale=1 / 16, softcap=50.0, causal=True, smax=291, seed=14, causal_prefix=278),
    attn_case("attn_tc_d256_prefix_lm_short", 2, 2, 1, 200, 200, 256, scale=1 / 16, softcap=50.0, causal=True, smax=208, seed=15, causal_prefix=7),
    attn_case("attn_tc_d256_prefix_lm_all", 1, 2, 2, 150, 150, 256, scale=1 / 16, softcap=50.0, causal=True, smax=150, seed=16, causal_prefix=150),
    attn_case("attn_mma_d128_prefix_lm", 2, 2, 2, 100, 100, 128, causal=True, smax=100, seed=17, causal_prefix=70),
    attn_case("attn_mma_d128_left_padded", 2, 2, 2, 100, 130, 128, smax=130, seed=13, kv_start=[9, 77]),
    # sliding-window layers (even Gemma2 layers): key slot j masked for query slot i when i - j >= window
    attn_case("attn_tc_d256_window_bidirectional", 2, 8, 4, 278, 278, 256, scale=1 / 16, softcap=50.0, smax=290, seed=51, window=48),
    attn_case("attn_tc_d256_window_causal_gqa", 2, 8, 4, 200, 200, 256, scale=1 / 16, softcap=50.0, causal=True, smax=208, seed=52, window=70),
    attn_case("attn_tc_d256_window_causal_offset_left_padded", 3, 2, 1, 150, 214, 256, scale=1 / 16, softcap=50.0, causal=True, smax=214, seed=53,
              window=130, kv_start=[0, 5, 60]),
    attn_case("attn_tc_d256_window_prefix_lm", 2, 2, 1, 200, 200, 256, scale=1 / 16, softcap=50.0, causal=True, smax=208, seed=54, causal_prefix=90,
              window=64),
    attn_case("attn_mma_d128_window", 2, 2, 2, 100, 130, 128, smax=130, seed=55, kv_start=[9, 77], window=40),
    decode_attn_case,
    decode_attn_fused_case,
    decode_attn_fused_window_case,
]


# ------------------------------------------------------------------------------------------------ fused ops
def layernorm_case(dev="cuda:0"):
    res = Result("layernorm")
    for cols, rows in ((144, 37), (1152, 300), (4304, 5), (128, 290), (1024, 577 * 2), (2304, 260), (512, 1000)):
        g = _gen(cols)
        x, ga, be = _randn(g, rows, cols), _randn(g, cols) * 0.1 + 1, _randn(g, cols) * 0.1

        def run(ops, to):
            ob, of = ops.zeros((rows, cols), BF16), ops.zeros((rows, cols), F32)
            ops.layernorm(to(x), to(ga), to(be), 1e-6, out_bf16=ob, out_f32=of, relu=(cols == 128))
            return ob, of
        (cb, cf), (rb, rf_) = _both(run, dev)
        res.add(f"bf16[{cols}]", _err(cb, rb), TOL_BF16)
        res.add(f"f32[{cols}]", _err(cf, rf_), 1e-4)
    return res


def rmsnorm_case(dev="cuda:0"):
    res = Result("rmsnorm_residual")
    for cols, rows in ((2304, 70), (512, 9), (2304, 2100), (512, 2500), (1024, 2049)):
        g = _gen(cols)
        x, br = _randn(g, rows, cols), _randn(g, rows, cols, scale=3.0)
        wp, wq = _randn(g, cols) * 0.1, _randn(g, cols) * 0.1

        def run(ops, to):
            xx, ob = to(x), ops.zeros((rows, cols), BF16)
            ops.rmsnorm_residual(xx, branch=to(br), w_post=to(wp), w_pre=to(wq), eps=1e-6, out_bf16=ob)
            ob2 = ops.zeros((rows, cols), BF16)
            ops.rmsnorm_residual(xx, w_pre=to(wp), eps=1e-6, out_bf16=ob2)
            return xx, ob, ob2
        (cx, cb, cb2), (rx, rb, rb2) = _both(run, dev)
        res.add(f"x[{cols}]", _err(cx, rx), 1e-5)
        res.add(f"h[{cols}]", _err(cb, rb), TOL_BF16)
        res.add(f"h2[{cols}]", _err(cb2, rb2), TOL_BF16)
    return res


def rope_case(dev="cuda:0"):
    res = Result("rope_kv")
    for (B, S, hq, hkv, d, smax, pos0, padded) in ((2, 37, 4, 2, 256, 64, 11, False), (8, 150, 8, 4, 256, 160, 0, False),
                                                   (5, 300, 2, 1, 64, 300, 0, False), (3, 40, 4, 2, 256, 64, 0, True),
                                                   (8, 150, 8, 4, 256, 160, 0, True)):
        g = _gen(5 + S)
        qkv = _randn(g, B * S, (hq + 2 * hkv) * d, dtype=BF16)
        pads = (torch.arange(B, dtype=torch.int32) * 5) % 17 if padded else None      # left-padded rows (row 0 unpadded)

        def run(ops, to):
            q, kc, vc = ops.zeros((B * S, hq * d), BF16), ops.zeros((B, smax, hkv, d), BF16), ops.zeros((B, smax, hkv, d), BF16)
            ops.rope_kv(to(qkv), q, kc, vc, batch=B, s=S, hq=hq, hkv=hkv, d=d, smax=smax, pos0=pos0, theta=10000.0, row_pads=to(pads))
            return q, kc, vc
        c, r = _both(run, dev)
        for nm, a, b in zip(("q", "k", "v"), c, r):
            res.add(f"{nm}[{B}x{S}{'p' if padded else ''}]", _err(a, b), TOL_BF16 if nm != "v" else 0.0)
    return res


def embed_case(dev="cuda:0"):
    g = _gen(6)
    B, S, H, V, n_act, n_img = 3, 40, 512, 9216, 8194, 16
    act_lo, img_tok = 1022, 1021
    ids = torch.randint(3, 1000, (B, S), generator=g)
    ids[:, 2:2 + n_img] = img_tok
    ids[:, -3:] = torch.randint(act_lo, act_lo + n_act, (B, 3), generator=g)
    emb, sp, img = _randn(g, V, H, dtype=BF16), _randn(g, n_act, H, dtype=BF16), _randn(g, B, n_img, H)

    def run(ops, to):
        x, st = ops.zeros((B * S, H), F32), ops.zeros((1,), torch.int32)
        ops.embed_tokens(to(ids), to(emb), to(sp), to(img), x, image_token=img_tok, act_lo=act_lo, n_act=n_act, n_img=n_img,
                         normalizer=float(torch.tensor(H ** 0.5)), status=st)
        return x, st
    (cx, cs), (rx, rs) = _both(run, dev)
    res = Result("embed_tokens")
    res.add("x", _err(cx, rx), 1e-6)
    res.add("status", abs(int(cs.item()) - int(rs.item())), 0)
    return res


def argmax_case(dev="cuda:0"):
    g = _gen(7)
    lg = _randn(g, 64, 8194)
    lg[3, 100] = lg[3, 7000] = 9.0          # tie -> first index wins (torch.argmax)
    lg[5, 8193] = 11.0

    def run(ops, to):
        out = ops.zeros((64, 4), torch.int64)
        ops.argmax_rows(to(lg), out[:, 2], id_offset=257153)
        return out
    c, r = _both(run, dev)
    res = Result("argmax_rows")
    res.add("ids", float((c.cpu() != r).sum()), 0)
    return res


def patchify_case(dev="cuda:0"):
    g = _gen(8)
    px = torch.rand(2, 3, 224, 224, generator=g)
    px = torch.nn.functional.avg_pool2d(px, 3, 1, 1)

    def run(ops, to):
        a, b = ops.zeros((2 * 256, 640), BF16), ops.zeros((2 * 576, 768), BF16)
        ops.siglip_patchify(to(px), a)
        ops.zoe_patchify(to(px), b)
        return a, b
    (ca, cb), (ra, rb) = _both(run, dev)
    res = Result("patchify")
    res.add("siglip", _err(ca, ra), 0.0)
    res.add("zoe_bicubic", _err(cb, rb), TOL_BF16)
    return res


def assemble_concat_case(dev="cuda:0"):
    g = _gen(9)
    B, n, c = 2, 576, 128
    patches, cls = _randn(g, B * n, c), _randn(g, c)

    def run(ops, to):
        x = ops.zeros((B * (n + 1), c), F32)
        ops.beit_assemble(to(patches), to(cls), x, batch=B, n=n, c=c)
        a = ops.zeros((B * n, 2 * c), BF16)
        ops.readout_concat(x, a, batch=B, n=n, c=c)
        return x, a
    (cx, ca), (rx, ra) = _both(run, dev)
    res = Result("beit_assemble+readout_concat")
    res.add("x", _err(cx, rx), 0.0)
    res.add("concat", _err(ca, ra), 0.0)
    return res


def shuffle_im2col_case(dev="cuda:0"):
    g = _gen(10)
    B, h, w, c = 2, 24, 24, 64
    gt = _randn(g, B * h * w, 16 * c, dtype=BF16)
    x = _randn(g, B, h, w, c, dtype=BF16)

    def run(ops, to):
        o4 = ops.zeros((B * h * 4 * w * 4, c), BF16)
        ops.pixel_shuffle(to(gt), o4, batch=B, h=h, w=w, c=c, f=4)
        o2 = ops.zeros((B * h * 2 * w * 2, c), BF16)
        ops.pixel_shuffle(to(gt)[:, : 4 * c].contiguous(), o2, batch=B, h=h, w=w, c=c, f=2)
        col = ops.zeros((B * (h // 2) * (w // 2), 9 * c), BF16)
        ops.im2col3x3_s2(to(x), col, batch=B, h=h, w=w, c=c)
        return o4, o2, col
    c_, r_ = _both(run, dev)
    res = Result("pixel_shuffle+im2col_s2")
    for nm, a, b in zip(("shuffle4", "shuffle2", "im2col"), c_, r_):
        res.add(nm, _err(a, b), 0.0)
    return res


def bilinear_case(dev="cuda:0"):
    g = _gen(11)
    res = Result("bilinear_nhwc")
    for (h, w, oh, ow, c) in ((12, 12, 24, 24, 128), (24, 24, 48, 48, 64), (7, 9, 21, 20, 8), (48, 48, 96, 96, 256), (5, 6, 50, 41, 72),
                              (20, 20, 14, 14, 16), (9, 33, 9, 70, 136)):
        x, ad = _randn(g, 2, h, w, c, dtype=BF16), _randn(g, 2, oh, ow, c, dtype=BF16)

        def run(ops, to):
            o, orl = ops.zeros((2, oh, ow, c), BF16), ops.zeros((2, oh, ow, c), BF16)
            ops.bilinear_nhwc(to(x), o, batch=2, h=h, w=w, c=c, oh=oh, ow=ow, add=to(ad), out_relu=orl)
            rl = ops.zeros((2, h, w, c), BF16)
            ops.relu_bf16(to(x), rl)
            return o, orl, rl
        (co, cr, cl), (ro, rr, rl_) = _both(run, dev)
        res.add(f"out[{h}->{oh}]", _err(co, ro), TOL_BF16)
        res.add(f"relu[{h}->{oh}]", _err(cr, rr), TOL_BF16)
        res.add(f"relu_bf16[{h}]", _err(cl, rl_), 0.0)
    return res


def zoe_tail_case(dev="cuda:0"):
    g = _gen(12)
    B, h, w, oh, ow, na, nb, nh = 2, 12, 12, 24, 24, 16, 64, 40
    attr = _randn(g, B * oh * ow, na, dtype=BF16)
    prev = torch.nn.functional.softplus(_randn(g, B, h, w, nb))
    conv = _randn(g, B * 144, 128)
    t, e = _randn(g, B * oh * ow, nh, dtype=BF16), _randn(g, B, h, w, nh, dtype=BF16)
    b1, w2, b2 = _randn(g, nh) * 0.1, _randn(g, 4, nh) * 0.3, torch.tensor([0.0, 0.0, -4.0, 3.0])
    sp_in = _randn(g, 1000, dtype=BF16)

    def run(ops, to):
        bins = ops.zeros((B * oh * ow, nb), F32)
        ops.zoe_attractor(to(attr), to(prev), bins, batch=B, h=h, w=w, oh=oh, ow=ow, na=na, nbins=nb)
        e32, eb = ops.zeros((B * 145, 128), F32), ops.zeros((B * 145, 128), BF16)
        ops.zoe_router_embed(to(conv), e32, eb, batch=B, n=144, c=128)
        depth = ops.zeros((B, oh, ow), F32)
        ops.zoe_depth_tail(to(t), to(e), to(b1), to(w2), to(b2), to(prev), depth, batch=B, h=h, w=w, oh=oh, ow=ow, nh=nh,
                           nbins=nb, min_temp=0.0212, max_temp=50.0)
        sp = ops.zeros((1000,), F32)
        ops.softplus_f32(to(sp_in), sp)
        return bins, e32, eb, depth, sp
    c_, r_ = _both(run, dev)
    res = Result("zoe_metric_tail")
    for nm, a, b, tol in zip(("attractor", "router_embed", "router_embed_bf16", "depth_tail", "softplus"), c_, r_,
                             (1e-4, 1e-4, TOL_BF16, 2e-3, 1e-5)):
        res.add(nm, _err(a, b), tol)
    # fused tail (first CLB layer inside the kernel): ragged tiles (oh, ow not multiples of 16), batch 3, and the exact x2 geometry
    for (B2, h2, w2_, oh2, ow2) in ((3, 11, 13, 21, 27), (2, 24, 24, 48, 48), (1, 5, 5, 37, 35)):
        g2 = _gen(100 + oh2)
        xr = _randn(g2, B2 * oh2 * ow2, 32, dtype=BF16)
        wa = (_randn(g2, 40, 32) * 0.3).to(BF16)
        e2 = _randn(g2, B2, h2, w2_, nh, dtype=BF16)
        prev2 = torch.nn.functional.softplus(_randn(g2, B2, h2, w2_, nb)) * 3.0

        def run2(ops, to):
            d2 = ops.zeros((B2, oh2, ow2), F32)
            ops.zoe_depth_tail_fused(to(xr), to(wa), to(e2), to(b1), to(w2), to(b2), to(prev2), d2, batch=B2, h=h2, w=w2_, oh=oh2, ow=ow2,
                                     min_temp=0.0212, max_temp=50.0)
            return d2
        cf, rf2 = _both(run2, dev)
        res.add(f"depth_tail_fused[{h2}x{w2_}->{oh2}x{ow2}]", _err(cf, rf2), 2e-3)
    return res


def ego3d_case(dev="cuda:0"):
    from spatialvla_b200.configs import default_intrinsic_224
    g = _gen(13)
    B = 3
    depth = torch.rand(B, 384, 384, generator=g) * 4.7 + 0.3
    depth = torch.nn.functional.avg_pool2d(depth[:, None], 9, 1, 4)[:, 0].contiguous()
    K1 = torch.tensor(default_intrinsic_224())
    Kb = K1[None].repeat(B, 1, 1) * torch.tensor([1.0, 1.1, 0.9])[:, None, None]
    Kb[:, 2, 2] = 1.0
    res = Result("ego3d_encode")
    for nm, K in (("K3x3", K1), ("Kbatched", Kb.contiguous())):
        def run(ops, to):
            xyz, enc = ops.zeros((B * 256, 12), F32), ops.zeros((B * 256, 208), BF16)
            ops.ego3d_encode(to(depth), to(K), xyz, enc, n_freqs=8)
            return xyz, enc
        (cx, ce), (rx, re_) = _both(run, dev)
        res.add(f"xyz[{nm}]", _err(cx, rx), 1e-5)
        # sin/cos of 2^7 * x amplify the fp32 xyz difference: compare with an absolute bf16-level tolerance
        res.add(f"enc[{nm}]", float((ce.float().cpu() - re_.float()).abs().max()), 2e-2)
    return res


def tokenizer_case(dev="cuda:0"):
    """Ids EXACTLY equal to the golden vectors minted from the live reference (atan2-free exact angular binning, no forgiven rows);
    decode with the host-tabulated bin-centre sin / cos: rotation / gripper 0 ulp, x y z within 1 ulp of the golden (the golden was
    minted with this container's libm; a different host CPU may select another glibc sin / cos variant) and 0 ulp against numpy on
    the same host.  The library-atan2 path (no table) is measured beside it and its mismatch count reported."""
    from spatialvla_b200.action_tokenizer import edge_trig_table
    from spatialvla_b200.ops import CudaOps
    ops = CudaOps(dev)
    res = Result("tokenizer")
    for name in ("gauss", "uniform"):
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"tokenizer_{name}.npz"))
        keys = ("theta_bins", "phi_bins", "r_bins", "roll_bins", "pitch_bins", "yaw_bins")
        edges = np.concatenate([gold[f"edge_{k}"] for k in keys])
        nb = [len(gold[f"edge_{k}"]) - 1 for k in keys] + [2]
        acts = torch.from_numpy(gold["actions"]).to(dev)
        ed = torch.from_numpy(edges).to(dev)
        trig, pn, pg = edge_trig_table(gold["edge_theta_bins"][1:-1], gold["edge_phi_bins"][1:-1])
        ids = torch.zeros(acts.shape[0], 3, dtype=torch.int32, device=dev)
        ops.tok_encode(acts, ed, nb, ids, trig=torch.from_numpy(trig).to(dev), phi_nonpos=pn, phi_neg=pg)
        n_bad = int((ids.cpu().numpy() != gold["local_ids"]).any(1).sum())
        res.add(f"encode_mismatch_rows[{name}]", n_bad, 0)
        ids_lib = torch.zeros_like(ids)
        ops.tok_encode(acts, ed, nb, ids_lib)                                  # library atan2: informational
        res.add(f"encode_mismatch_rows_library_atan2[{name}]", int((ids_lib.cpu().numpy() != gold["local_ids"]).any(1).sum()), 64)
        cen = [0.5 * (gold[f"edge_{k}"][:-1] + gold[f"edge_{k}"][1:]) for k in ("theta_bins", "phi_bins")]
        ctrig = np.ascontiguousarray(np.concatenate([np.stack([np.sin(c), np.cos(c)], 1) for c in cen]))
        dids = torch.from_numpy(gold["decode_ids"]).to(dev)
        out = torch.zeros(dids.shape[0], 7, dtype=torch.float64, device=dev)
        ops.tok_decode(dids, ed, nb, int(gold["begin"]), out, center_trig=torch.from_numpy(ctrig).to(dev))
        ref = gold["decode_actions"]
        ulp = np.abs(out.cpu().numpy() - ref) / np.maximum(np.spacing(np.abs(ref)), 1e-300)
        res.add(f"decode_ulp_xyz_vs_golden[{name}]", float(ulp[:, :3].max()), 1)
        res.add(f"decode_ulp_rot_grip[{name}]", float(ulp[:, 3:].max()), 0)
        from oracle import tokenizer_ref as T
        num_bins = {"translation": {k: nb[i] for i, k in enumerate(keys[:3])}, "rotation": {k: nb[3 + i] for i, k in enumerate(keys[3:])},
                    "gripper": 2}
        pol = {"translation": {k: gold[f"edge_{k}"] for k in keys[:3]}, "rotation": {k: gold[f"edge_{k}"] for k in keys[3:]}}
        here = T.decode(gold["decode_ids"] - int(gold["begin"]), pol, num_bins)
        res.add(f"decode_maxabs_vs_numpy_same_host[{name}]", float(np.abs(out.cpu().numpy() - here).max()), 0)
        out_lib = torch.zeros_like(out)
        ops.tok_decode(dids, ed, nb, int(gold["begin"]), out_lib)              # library sincos
        ulp = np.abs(out_lib.cpu().numpy() - ref) / np.maximum(np.spacing(np.abs(ref)), 1e-300)
        res.add(f"decode_ulp_xyz_library_sincos[{name}]", float(ulp[:, :3].max()), 4)
        oob = torch.from_numpy(gold["oob_ids"]).to(dev)
        out2 = torch.zeros(oob.shape[0], 7, dtype=torch.float64, device=dev)
        ops.tok_decode(oob, ed, nb, int(gold["begin"]), out2, center_trig=torch.from_numpy(ctrig).to(dev))
        res.add(f"decode_oob[{name}]", float(np.abs(out2.cpu().numpy() - gold["oob_actions"]).max()), 1e-15)
    return res


def cross_entropy_case(dev="cuda:0"):
    """svla_cross_entropy_rows vs nn.CrossEntropyLoss: the real (odd) vocabulary width so rows are not 16-byte aligned, a padded
    row stride, ignored rows, an argmax tie, a dominant logit (online max rescale), chunked calls with the summary on the last."""
    res = Result("cross_entropy_rows")
    for (rows, cols, ld, chunk) in ((37, 265347, 265347, 37), (9, 9216, 9220, 4), (5, 3, 3, 5), (300, 1001, 1001, 128)):
        g = _gen(rows + cols)
        lg = torch.zeros(rows, ld)
        lg[:, :cols] = _randn(g, rows, cols, scale=4.0)
        lab = torch.randint(0, cols, (rows,), generator=g)
        lab[1::5] = -100
        lg[0, cols - 1] = 60.0                       # dominant last column: exp(m_old - m_new) underflows
        lab[0] = cols - 1
        lg[2, 1] = lg[2, cols - 2] = 25.0            # tie -> first index
        lab[2] = 1

        def run(ops, to):
            l_, y = to(lg), to(lab)
            rl, ra, sm = ops.zeros((rows,), F32), ops.zeros((rows,), torch.int64), ops.zeros((3,), F32)
            for r0 in range(0, rows, chunk):
                r1 = min(rows, r0 + chunk)
                ops.cross_entropy_rows(l_[r0:r1, :cols], y, rl, ra, row_offset=r0, summary=sm if r1 == rows else None)
            return rl, ra, sm
        (crl, cra, csm), (rrl, rra, rsm) = _both(run, dev)
        # fp32 accumulation of up to 265 347 terms against the float64 reference: 1e-4 absolute on losses of ~20 (5e-6 relative)
        res.add(f"row_loss[{cols}]", float((crl.cpu() - rrl).abs().max()), 1e-4)
        res.add(f"argmax[{cols}]", float((cra.cpu() != rra).sum()), 0)
        res.add(f"mean[{cols}]", abs(float(csm[0]) - float(rsm[0])), 1e-4)
        res.add(f"count_hits[{cols}]", float((csm[1:].cpu() - rsm[1:]).abs().max()), 0)
    return res


def cross_entropy_bwd_case(dev="cuda:0"):
    """svla_cross_entropy_bwd vs (softmax - onehot) * softcap' / count in float64: odd vocabulary (unaligned fp32 rows), K padding
    columns zeroed, ignored rows zero, chunked calls, with and without soft-capping."""
    res = Result("cross_entropy_bwd")
    for (rows, cols, chunk, cap) in ((19, 265347, 19, 30.0), (9, 9216, 4, 30.0), (5, 3, 5, 0.0), (70, 1001, 32, 30.0)):
        g = _gen(rows * 3 + cols)
        lg = _randn(g, rows, cols, scale=6.0)
        if cap:
            lg = cap * torch.tanh(lg / cap)                       # what the lm_head GEMM epilogue hands over
        lab = torch.randint(0, cols, (rows,), generator=g)
        lab[1::5] = -100
        ldo = (cols + 7) // 8 * 8

        def run(ops, to):
            l_, y = to(lg), to(lab)
            rl, ra, sm = ops.zeros((rows,), F32), ops.zeros((rows,), torch.int64), ops.zeros((3,), F32)
            ops.cross_entropy_rows(l_, y, rl, ra, summary=sm)
            dz = ops.zeros((rows, ldo), BF16) + 7.0               # poison: every element must be written
            for r0 in range(0, rows, chunk):
                r1 = min(rows, r0 + chunk)
                ops.cross_entropy_bwd(l_[r0:r1], y, rl, sm, dz[r0:r1], row_offset=r0, softcap=cap)
            return dz
        c, r = _both(run, dev)
        c, r = c.float().cpu(), r.float()
        res.add(f"dz[{cols}]", float((c - r).abs().max()) / float(r.abs().max()), 1e-2)          # bf16 output
        res.add(f"pad+ignored[{cols}]", float(c[:, cols:].abs().max() if ldo > cols else 0.0) + float(c[1::5].abs().max()), 0)
        # the bf16 rounding must not bias the row sums: sum_j dz / softcap' = 0 per live row (softmax - onehot sums to zero)
        if not cap:
            res.add(f"rowsum[{cols}]", float(c[0::5].sum(-1).abs().max()), 2e-2)
    return res


def adamw_case(dev="cuda:0"):
    """svla_adamw_step vs the torch.optim.AdamW arithmetic (oracle/ops_ref.py, itself checked against torch.optim.AdamW on the CPU):
    three steps on flat buffers whose length is not a multiple of 4, with weight decay and a gradient scale."""
    res = Result("adamw_step")
    for (n, wd, gs) in ((1_000_003, 0.01, 0.5), (64, 0.0, 1.0), (7, 0.1, 1.0)):
        g = _gen(n)
        p0 = _randn(g, n)
        grads = [_randn(g, n, scale=0.1 * (s + 1)) for s in range(3)]

        def run(ops, to):
            p, m, v = to(p0), ops.zeros((n,), F32), ops.zeros((n,), F32)
            for s, gr in enumerate(grads):
                ops.adamw_step(p, to(gr), m, v, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=wd, step=s + 1, grad_scale=gs)
            return p, m, v
        (cp, cm, cv), (rp, rm, rv) = _both(run, dev)
        res.add(f"p[{n}]", _err(cp, rp), 2e-6)
        res.add(f"m[{n}]", _err(cm, rm), 2e-6)
        res.add(f"v[{n}]", _err(cv, rv), 2e-6)
        res.add(f"moved[{n}]", 0.0 if float((rp - p0).abs().max()) > 1e-4 else 1.0, 0)
    return res


# ------------------------------------------------------------------------------------------------ fine-tune step kernels
def gemm_lora_ext_case(name, M, N, K, K2, *, pair=None, geglu=False, block_n=0, out=("bf16",), seed=0):
    """K extension of the tcgen05 GEMM: acc = A W^T + A2 W2^T in one TMEM tile (un-merged LoRA Linear and its input gradient)."""
    def case(dev="cuda:0"):
        g = _gen(seed + 77)
        a, wt = _randn(g, M, K, dtype=BF16), (_randn(g, N, K) / K ** 0.5).to(BF16)
        k2p = (K2 + 63) // 64 * 64                                    # operand buffers padded with zero columns, like the step's pool
        a2, w2 = torch.zeros(M, k2p, dtype=BF16), torch.zeros(N, k2p, dtype=BF16)
        a2[:, :K2] = _randn(g, M, K2, dtype=BF16)
        w2[:, :K2] = (_randn(g, N, K2) * 0.3).to(BF16)
        ncol = N // 2 if geglu else N

        def run(ops, to):
            ob = ops.zeros((M, ncol), BF16) if "bf16" in out else None
            of = ops.zeros((M, ncol), F32) if "f32" in out else None
            ops.gemm(to(a), to(wt), out_bf16=ob, out_f32=of, geglu=geglu, a2=to(a2), w2=to(w2), block_n=block_n, impl=pair)
            return ob, of
        (cb, cf), (rb, rf_) = _both(run, dev)
        res = Result(name)
        if cb is not None:
            res.add("bf16", _err(cb, rb), TOL_BF16)
        if cf is not None:
            res.add("f32", _err(cf, rf_), TOL_F32)
        # the extension term must matter: without it the result is clearly different
        base = (a.float() @ wt.float().t())
        ext = a2.float() @ w2.float().t()
        res.add("ext_is_visible", 0.0 if float(ext.abs().max()) > 0.05 * float(base.abs().max()) else 1.0, 0)
        return res
    case.__name__ = name
    return case


LORA_GEMM_CASES = [
    gemm_lora_ext_case("gemm_ext_r32", 300, 256, 128, 32, out=("bf16", "f32")),
    gemm_lora_ext_case("gemm_ext_r96_bn128", 640, 512, 192, 96, block_n=128),
    gemm_lora_ext_case("gemm_ext_r64_geglu", 512, 1024, 256, 64, geglu=True),
    gemm_lora_ext_case("gemm_ext_pair", 1024, 2304, 320, 96, pair=3),
    gemm_lora_ext_case("gemm_ext_multiwave", 9312, 4096, 2304, 96),
    gemm_lora_ext_case("gemm_ext_bn64_f32", 200, 64, 72, 32, out=("f32",)),
]


def train_norm_case(dev="cuda:0"):
    res = Result("train_norms")
    for (rows, cols) in ((37, 2304), (1001, 512), (64, 1152)):
        g = _gen(rows + cols)
        x, br = _randn(g, rows, cols), _randn(g, rows, cols, scale=3.0)
        wp, wq = _randn(g, cols, scale=0.2), _randn(g, cols, scale=0.2)
        dy_b, dy_f = _randn(g, rows, cols, dtype=BF16), _randn(g, rows, cols)
        acc0 = _randn(g, rows, cols)

        def fwd(ops, to):
            xo, h = ops.zeros((rows, cols), F32), ops.zeros((rows, cols), BF16)
            ops.rmsnorm_train_fwd(to(x), branch=to(br), w_post=to(wp), w_pre=to(wq), eps=1e-6, x_out=xo, h=h)
            h0 = ops.zeros((rows, cols), BF16)
            ops.rmsnorm_train_fwd(to(x), w_pre=to(wq), eps=1e-6, h=h0)
            return xo, h, h0
        (cx, ch, ch0), (rx, rh, rh0) = _both(fwd, dev)
        res.add(f"fwd_x[{rows}x{cols}]", _err(cx, rx), 1e-5)
        res.add(f"fwd_h[{rows}x{cols}]", _err(ch, rh), TOL_BF16)
        res.add(f"fwd_h_only[{rows}x{cols}]", _err(ch0, rh0), TOL_BF16)

        def bwd(ops, to):
            acc, ob = to(acc0), ops.zeros((rows, cols), BF16)
            ops.rmsnorm_bwd(to(x), to(wq), to(dy_b), eps=1e-6, dx_accum=acc)
            ops.rmsnorm_bwd(to(br), to(wp), to(dy_f), eps=1e-6, dx_bf16=ob)
            return acc, ob
        (ca, cb), (ra, rb) = _both(bwd, dev)
        res.add(f"rms_bwd_accum[{rows}x{cols}]", _err(ca, ra), 1e-4)
        res.add(f"rms_bwd_bf16[{rows}x{cols}]", _err(cb, rb), TOL_BF16)
        # labelled-row indirection of the final norm
        nsel = min(rows, 13)
        sel = torch.randperm(rows, generator=g)[:nsel].sort().values
        dyr = _randn(g, nsel, cols)

        def bwd_idx(ops, to):
            acc = ops.zeros((rows, cols), F32)
            ops.rmsnorm_bwd(to(x), to(wq), to(dyr), eps=1e-6, row_idx=to(sel), dx_accum=acc)
            return acc
        ci, ri = _both(bwd_idx, dev)
        res.add(f"rms_bwd_rows[{rows}x{cols}]", _err(ci, ri), 1e-4)
        gm, bt = _randn(g, cols, scale=0.5) + 1.0, _randn(g, cols, scale=0.3)
        for relu in (False, True):
            def ln(ops, to):
                acc, cp, ob = to(acc0), ops.zeros((rows, cols), BF16), ops.zeros((rows, cols), BF16)
                ops.layernorm_bwd(to(x), to(gm), to(bt), to(dy_b), eps=1e-6, relu=relu, dx_accum=acc, copy_bf16=cp, dx_bf16=ob)
                return acc, cp, ob
            (ca, cc, cb), (ra, rc, rb) = _both(ln, dev)
            res.add(f"ln_bwd_accum[{rows}x{cols},relu={int(relu)}]", _err(ca, ra), 1e-4)
            res.add(f"ln_bwd_copy[{rows}x{cols},relu={int(relu)}]", _err(cc, rc), TOL_BF16)
            res.add(f"ln_bwd_bf16[{rows}x{cols},relu={int(relu)}]", _err(cb, rb), TOL_BF16)
    return res


def train_elementwise_case(dev="cuda:0"):
    res = Result("train_elementwise")
    g = _gen(5)
    rows, inter = 333, 1024
    gu, dact = _randn(g, rows, 2 * inter, dtype=BF16, scale=1.5), _randn(g, rows, inter, dtype=BF16)

    def geglu(ops, to):
        act, dgu = ops.zeros((rows, inter), BF16), ops.zeros((rows, 2 * inter), BF16)
        ops.geglu_fwd(to(gu), act)
        ops.geglu_bwd(to(gu), to(dact), dgu)
        return act, dgu
    (ca, cd), (ra, rd) = _both(geglu, dev)
    res.add("geglu_fwd", _err(ca, ra), TOL_BF16)
    res.add("geglu_bwd", _err(cd, rd), TOL_BF16)
    z, df = _randn(g, 200, 4304, dtype=BF16, scale=2.0), _randn(g, 200, 4304, dtype=BF16)

    def gelu(ops, to):
        f, dz = ops.zeros(z.shape, BF16), ops.zeros(z.shape, BF16)
        ops.gelu_tanh_fwd(to(z), f)
        ops.gelu_tanh_bwd(to(z), to(df), dz)
        return f, dz
    (cf, cz), (rf_, rz) = _both(gelu, dev)
    res.add("gelu_fwd", _err(cf, rf_), TOL_BF16)
    res.add("gelu_bwd", _err(cz, rz), TOL_BF16)
    B, S, hq, hkv, d = 3, 19, 4, 2, 256
    dq = _randn(g, B * S, (hq + 2 * hkv) * d, dtype=BF16)

    def rope(ops, to):
        t = to(dq).clone()
        ops.rope_bwd(t, batch=B, s=S, hq=hq, hkv=hkv, d=d, theta=10000.0)
        return t
    cr, rr = _both(rope, dev)
    res.add("rope_bwd", _err(cr, rr), TOL_BF16)
    res.add("rope_bwd_v_untouched", float((cr.cpu()[:, (hq + hkv) * d:].float() - dq[:, (hq + hkv) * d:].float()).abs().max()), 0)
    src = _randn(g, 500, 512)
    idx = torch.randint(0, 500, (77,), generator=g)

    def cast(ops, to):
        o1, o2 = ops.zeros((77, 512), BF16), ops.zeros((500, 512), BF16)
        ops.rows_cast(to(src), o1, row_idx=to(idx), scale=48.0 / 47.0)
        ops.rows_cast(to(src), o2)
        return o1, o2
    (c1, c2), (r1, r2) = _both(cast, dev)
    res.add("rows_cast_gather", _err(c1, r1), TOL_BF16)
    res.add("rows_cast", _err(c2, r2), TOL_BF16)
    return res


def attn_bwd_case(name, B, hq, hkv, sq, sk, d, *, softcap=0.0, causal=False, prefix=0, seed=0, with_lse=True, window=0):
    """svla_attention_bwd (3 launches) vs the closed form oracle/backward_ref.softcap_attention_bwd (itself pinned on autograd);
    q/k/v packed like the step's tensors (q in its own tensor, k / v as rows of a [B, sk, hkv, d] cache, gradients as column blocks
    of one [tokens, (hq + 2 hkv) d] tensor)."""
    def case(dev="cuda:0"):
        g = _gen(seed + 31)
        q = _randn(g, B * sq, hq * d, dtype=BF16)
        kc, vc = _randn(g, B, sk, hkv, d, dtype=BF16), _randn(g, B, sk, hkv, d, dtype=BF16)
        dout = _randn(g, B * sq, hq * d, dtype=BF16)
        scale = d ** -0.5
        W = (hq + 2 * hkv) * d

        def run(ops, to):
            Q, K, V, dO = to(q), to(kc), to(vc), to(dout)
            out = ops.zeros((B * sq, hq * d), BF16)
            kvs = (sk * hkv * d, hkv * d)
            # with_lse: the forward keeps the row log-sum-exp and the backward runs the tcgen05 sweeps; without: warp-MMA kernels
            lse = ops.zeros((B, hq, (sq + 63) // 64 * 64), F32) if with_lse else None
            ops.attention(Q, K, V, out, batch=B, hq=hq, hkv=hkv, sq=sq, sk=sk, d=d, q_strides=(sq * hq * d, hq * d), k_strides=kvs,
                          v_strides=kvs, o_strides=(sq * hq * d, hq * d), scale=scale, softcap=softcap, causal=causal, causal_prefix=prefix,
                          lse=lse, window=window)
            assert sq == sk
            dqkv = ops.zeros((B * sq, W), BF16) + 5.0                       # poison: every element must be written
            ops.attention_bwd(Q, K, V, out, dO, dqkv, dqkv[:, hq * d:], dqkv[:, (hq + hkv) * d:], batch=B, hq=hq, hkv=hkv, sq=sq, sk=sk,
                              d=d, q_strides=(sq * hq * d, hq * d), k_strides=kvs, v_strides=kvs, o_strides=(sq * hq * d, hq * d),
                              do_strides=(sq * hq * d, hq * d), dq_strides=(sq * W, W), dk_strides=(sk * W, W), dv_strides=(sk * W, W),
                              scale=scale, softcap=softcap, causal=causal, causal_prefix=prefix, lse=lse, window=window)
            return dqkv if lse is None else (dqkv, lse)
        c, r = _both(run, dev)
        res = Result(name)
        if with_lse:
            (c, cl), (r, rl) = c, r
            res.add("fwd_lse2", float((cl.cpu()[:, :, :sq] - rl[:, :, :sq]).abs().max()), 3e-2)       # bf16 operands, log2 units
        c, r = c.float().cpu(), r.float()
        for nm, lo, hi in (("dq", 0, hq * d), ("dk", hq * d, (hq + hkv) * d), ("dv", (hq + hkv) * d, W)):
            res.add(nm, float((c[:, lo:hi] - r[:, lo:hi]).abs().max()) / max(float(r[:, lo:hi].abs().max()), 1e-20), 2e-2)
        return res
    case.__name__ = name
    return case


ATTN_BWD_CASES = [
    # warp-MMA kernels (no forward log-sum-exp): the fallback path
    attn_bwd_case("attn_bwd_mma_gemma_prefixlm", 2, 4, 2, 291, 291, 256, softcap=50.0, causal=True, prefix=278, with_lse=False),
    attn_bwd_case("attn_bwd_mma_siglip_d72", 2, 4, 4, 256, 256, 72, with_lse=False),
    attn_bwd_case("attn_bwd_mma_d64_causal_prefix", 2, 4, 1, 200, 200, 64, causal=True, prefix=50, softcap=20.0, with_lse=False),
    # tcgen05 sweeps (dQ / dK / dV) fed by the forward kernel's log-sum-exp
    attn_bwd_case("attn_bwd_mma_gemma_window", 2, 4, 2, 200, 200, 256, softcap=50.0, causal=True, prefix=90, with_lse=False, window=64, seed=7),
    # sliding-window layers in the tcgen05 sweeps
    attn_bwd_case("attn_bwd_gemma_window_prefixlm", 2, 4, 2, 291, 291, 256, softcap=50.0, causal=True, prefix=278, window=100, seed=8),
    attn_bwd_case("attn_bwd_gemma_window_causal", 1, 8, 4, 200, 200, 256, softcap=50.0, causal=True, window=48, seed=9),
    attn_bwd_case("attn_bwd_gemma_window_bidirectional", 2, 2, 2, 150, 150, 256, softcap=50.0, window=70, seed=10),
    attn_bwd_case("attn_bwd_tiny_tile", 1, 1, 1, 40, 40, 64),
    attn_bwd_case("attn_bwd_two_tiles_d64", 1, 2, 2, 128, 128, 64),
    attn_bwd_case("attn_bwd_gqa4_d256", 1, 4, 1, 150, 150, 256, softcap=50.0, causal=True, prefix=30),
    attn_bwd_case("attn_bwd_gemma_prefixlm", 2, 4, 2, 291, 291, 256, softcap=50.0, causal=True, prefix=278),
    attn_bwd_case("attn_bwd_gemma_causal", 1, 8, 4, 130, 130, 256, softcap=50.0, causal=True),
    attn_bwd_case("attn_bwd_gemma_bidirectional", 2, 2, 2, 77, 77, 256, softcap=50.0),
    attn_bwd_case("attn_bwd_siglip_d72", 2, 16, 16, 256, 256, 72),
    attn_bwd_case("attn_bwd_d72_ragged", 3, 2, 2, 100, 100, 72),
    attn_bwd_case("attn_bwd_d64_gqa_ragged", 2, 4, 1, 200, 200, 64),
    attn_bwd_case("attn_bwd_d128", 1, 2, 2, 96, 96, 128),
]


def gemm_tn_case(dev="cuda:0"):
    """svla_gemm_tn: rank-r reductions over the token dimension into windows of a fp32 gradient arena (fused q|k|v blocks, interleaved
    gate/up columns, a column-clipped window (K = 204 of a 208-wide operand), accumulation on top of existing values)."""
    res = Result("gemm_tn")
    g = _gen(11)
    M = 2000
    # fused qkv gB: 3 adapters x r=32 rows, each with its own column window of dY [M, 4096]
    s, y = _randn(g, M, 128, dtype=BF16), _randn(g, M, 1024, dtype=BF16)
    s[:, 96:] = 0
    init = _randn(g, 3 * 32 * 512)

    def qkv(ops, to):
        arena = to(init).clone()
        v0, v1, v2 = arena[:32 * 512].view(32, 512), arena[32 * 512:32 * 768].view(32, 256), arena[32 * 768:32 * 1024].view(32, 256)
        ops.gemm_tn(to(s), to(y), [(v0, 0, 32, 0, 1, 512), (v1, 32, 32, 512, 1, 256), (v2, 64, 32, 768, 1, 256)], r=96, n=1024, scale=0.5)
        return arena
    c, r = _both(qkv, dev)
    res.add("qkv_windows", _err(c, r), 2e-3)
    # gate/up: two adapters, interleaved columns; gA: one dense group with a clipped column count
    s2, y2 = _randn(g, M, 64, dtype=BF16), _randn(g, M, 2048, dtype=BF16)
    x208 = _randn(g, M, 208, dtype=BF16)

    def gu(ops, to):
        gg, gup, ga = ops.zeros((32, 1024), F32), ops.zeros((32, 1024), F32), ops.zeros((64, 204), F32)
        ops.gemm_tn(to(s2), to(y2), [(gg, 0, 32, 0, 2, 1024), (gup, 32, 32, 1, 2, 1024)], r=64, n=2048)
        ops.gemm_tn(to(s2), to(x208), [(ga, 0, 64, 0, 1, 204)], r=64, n=208, scale=2.0)
        return gg, gup, ga
    (c0, c1, c2), (r0, r1, r2) = _both(gu, dev)
    res.add("gate_stride2", _err(c0, r0), 2e-3)
    res.add("up_stride2", _err(c1, r1), 2e-3)
    res.add("clipped_cols", _err(c2, r2), 2e-3)
    # long contraction, r = 32, config-#5-sized token count
    M2 = 9312
    s3, y3 = _randn(g, M2, 64, dtype=BF16), _randn(g, M2, 2304, dtype=BF16)

    def big(ops, to):
        o = ops.zeros((32, 2304), F32)
        ops.gemm_tn(to(s3)[:, :64], to(y3), [(o, 0, 32, 0, 1, 2304)], r=32, n=2304)
        return o
    cb, rb = _both(big, dev)
    res.add("m9312_r32", _err(cb, rb), 2e-3)
    return res


def lora_pack_case(dev="cuda:0"):
    """svla_lora_pack through the step's own layout builder: every operand of a tiny adapted model against the torch re-statement."""
    from spatialvla_b200.lora import LoRAStepLayout
    from spatialvla_b200.configs import get_config_dict
    res = Result("lora_pack")
    cfg = get_config_dict("tiny")

    def run(ops, to):
        lay = LoRAStepLayout(cfg, ops, r=32, alpha=32.0, seed=3)
        lay.param.copy_(to(torch.randn(lay.param.numel(), generator=_gen(9))))
        lay.pack()
        return lay.pool
    c, r = _both(run, dev)
    res.add("pool", float((c.float().cpu() - r.float()).abs().max()), 0)
    res.add("pool_nonzero", 0.0 if float(r.float().abs().sum()) > 0 else 1.0, 0)
    return res


TRAIN_CASES = [train_norm_case, train_elementwise_case, gemm_tn_case, lora_pack_case]

FUSED_CASES = [layernorm_case, rmsnorm_case, rope_case, embed_case, argmax_case, cross_entropy_case, cross_entropy_bwd_case, adamw_case, patchify_case, assemble_concat_case,
               shuffle_im2col_case, bilinear_case, zoe_tail_case, ego3d_case, tokenizer_case]

ALL_CASES = {c.__name__: c for c in (SIMT_CASES + GEMM_CASES + PAIR_CASES + TMA_EPI_CASES + ROWTILE_CASES + SKINNY_CASES + ATTN_CASES + FUSED_CASES
                                     + LORA_GEMM_CASES + ATTN_BWD_CASES + TRAIN_CASES)}
