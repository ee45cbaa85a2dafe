#!/usr/bin/env python
"""Micro-benchmark of svla_attention_bwd (stats + dQ + dK/dV launches) and svla_gemm_tn at the shapes of config #5 (B = 32)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatialvla_b200.ops import CudaOps  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    ops = CudaOps("cuda:0")
    out = []
    for name, (B, hq, hkv, S, d, cap, causal, prefix) in {"gemma": (32, 8, 4, 291, 256, 50.0, True, 0), "gemma_prefix": (32, 8, 4, 291, 256, 50.0, True, 278),
                                                           "siglip": (32, 16, 16, 256, 72, 0.0, False, 0)}.items():
        q = torch.randn(B * S, hq * d, device="cuda").to(BF16)
        kc, vc = torch.randn(B, S, hkv, d, device="cuda").to(BF16), torch.randn(B, S, hkv, d, device="cuda").to(BF16)
        do = torch.randn(B * S, hq * d, device="cuda").to(BF16)
        o = torch.zeros_like(q)
        W = (hq + 2 * hkv) * d
        dqkv = torch.zeros(B * S, W, device="cuda", dtype=BF16)
        kvs, qs = (S * hkv * d, hkv * d), (S * hq * d, hq * d)
        kw = dict(batch=B, hq=hq, hkv=hkv, sq=S, sk=S, d=d, q_strides=qs, k_strides=kvs, v_strides=kvs, o_strides=qs, scale=d ** -0.5,
                  softcap=cap, causal=causal, causal_prefix=prefix)
        lse = torch.zeros(B, hq, (S + 63) // 64 * 64, device="cuda")
        ops.attention(q, kc, vc, o, lse=lse, **kw)
        fwd = timeit(lambda: ops.attention(q, kc, vc, o, lse=lse, **kw))
        bw = dict(do_strides=qs, dq_strides=(S * W, W), dk_strides=(S * W, W), dv_strides=(S * W, W), **kw)
        bwd_mma = timeit(lambda: ops.attention_bwd(q, kc, vc, o, do, dqkv, dqkv[:, hq * d:], dqkv[:, (hq + hkv) * d:], **bw))
        bwd_tc = timeit(lambda: ops.attention_bwd(q, kc, vc, o, do, dqkv, dqkv[:, hq * d:], dqkv[:, (hq + hkv) * d:], lse=lse, **bw))
        fl = 4.0 * B * hq * S * S * d
        out.append({"case": name, "fwd_ms": round(fwd, 4), "fwd_tflops": round(fl / fwd / 1e9, 1), "bwd_mma_ms": round(bwd_mma, 4),
                    "bwd_tcgen05_ms": round(bwd_tc, 4), "bwd_tcgen05_tflops_5units": round(2.5 * fl / bwd_tc / 1e9, 1)})
    for (M, r, n) in ((9312, 64, 18432), (9312, 96, 4096), (9312, 32, 2304), (8192, 32, 1152), (8192, 32, 4304)):
        s, y = torch.randn(M, (r + 63) // 64 * 64, device="cuda").to(BF16), torch.randn(M, n, device="cuda").to(BF16)
        dst = torch.zeros(r, n, device="cuda")
        ms = timeit(lambda: ops.gemm_tn(s, y, [(dst, 0, r, 0, 1, n)], r=r, n=n))
        out.append({"case": f"gemm_tn M={M} r={r} n={n}", "ms": round(ms, 4), "GBs": round((M * n * 2 + M * r * 2) / ms / 1e6, 1),
                    "tflops": round(2.0 * M * r * n / ms / 1e9, 1)})
    for row in out:
        print(json.dumps(row))


if __name__ == "__main__":
    main()
