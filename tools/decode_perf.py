"""Decode-step micro-benchmark: the weight-streaming skinny GEMMs (row-major vs tile-major weights), the fused
RoPE + cache + attention kernel and the sandwich norm, timed the way they run in production: captured in a CUDA graph over
26 distinct layers (so weights / caches come from HBM, not L2), CUDA events around the replay.
Usage: python tools/decode_perf.py [substr]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spatialvla_b200.ops import CudaOps, tile_weight

dev = "cuda:0"
ops = CudaOps(dev)
BF16, F32 = torch.bfloat16, torch.float32
only = sys.argv[1] if len(sys.argv) > 1 else ""
L, B = 26, 64


def graph_time(fn, iters=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for name, N, K, kw in [("dec_qkv", 4096, 2304, {}), ("dec_o", 2304, 2048, {}), ("dec_gateup", 18432, 2304, {"geglu": True}),
                       ("dec_down", 2304, 9216, {}), ("dec_head", 8194, 2304, {"head": True})]:
    if only and only not in name: continue
    ws = [(torch.randn(N, K, device=dev) / 48).to(BF16) for _ in range(L)]
    wt = [tile_weight(w) for w in ws]
    x = torch.randn(B, K, device=dev).to(BF16)
    splits = 1 if kw else ops.skinny_splits(N, K)
    if kw.get("geglu"):
        out = torch.empty(B, N // 2, device=dev, dtype=BF16)
        run = lambda w, t: ops.gemm_skinny(x, w, out_bf16=out, geglu=True, tiled_n=t)
    elif kw.get("head"):
        out = torch.empty(B, N, device=dev, dtype=F32)
        run = lambda w, t: ops.gemm_skinny(x, w, out_f32=out, tiled_n=t)
    else:
        out = torch.empty(splits, B, N, device=dev, dtype=F32)
        run = lambda w, t: ops.gemm_skinny(x, w, out_f32=out, tiled_n=t)
    for tag, mats, t in (("rowmajor", ws, None), ("tiled", wt, N)):
        ms = graph_time(lambda: [run(w, t) for w in mats])
        us = ms * 1e3 / L
        print({"name": name, "layout": tag, "splits": splits, "us": round(us, 2), "weight_GBs": round(N * K * 2 / us / 1e3, 1)}, flush=True)
    del ws, wt

if not only or "attn" in only:
    hq, hkv, d, smax, ctx = 8, 4, 256, 290, 285
    kc = [torch.randn(B, smax, hkv, d, device=dev).to(BF16) for _ in range(L)]
    vc = [torch.randn(B, smax, hkv, d, device=dev).to(BF16) for _ in range(L)]
    part = torch.randn(4, B, (hq + 2 * hkv) * d, device=dev)
    out = torch.empty(B, hq * d, device=dev, dtype=BF16)
    q = torch.randn(B, hq * d, device=dev).to(BF16)
    ms = graph_time(lambda: [ops.decode_attention_fused(part, kc[i], vc[i], out, batch=B, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, theta=1e4,
                                                        scale=1 / 16, softcap=50.0) for i in range(L)])
    us = ms * 1e3 / L
    print({"name": "decode_attention_fused", "us": round(us, 2), "kv_GBs": round(2 * B * ctx * hkv * d * 2 / us / 1e3, 1)}, flush=True)
    ms = graph_time(lambda: [ops.decode_attention(q, kc[i], vc[i], out, batch=B, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, scale=1 / 16, softcap=50.0)
                             for i in range(L)])
    us = ms * 1e3 / L
    print({"name": "decode_attention_unfused(+rope_kv separately)", "us": round(us, 2), "kv_GBs": round(2 * B * ctx * hkv * d * 2 / us / 1e3, 1)}, flush=True)

if not only or "norm" in only:
    H = 2304
    x = torch.randn(B, H, device=dev)
    w1, w2 = torch.randn(H, device=dev) * 0.1, torch.randn(H, device=dev) * 0.1
    h = torch.empty(B, H, device=dev, dtype=BF16)
    for splits in (1, 4, 8):
        br = [torch.randn(splits, B, H, device=dev) for _ in range(L)]
        ms = graph_time(lambda: [ops.rmsnorm_residual(x, branch=b_, w_post=w1, w_pre=w2, eps=1e-6, out_bf16=h) for b_ in br])
        print({"name": "rmsnorm_residual_decode", "splits": splits, "us": round(ms * 1e3 / L, 2)}, flush=True)
