"""Micro-benchmark of the ZoeDepth metric-bins tail at the 4B-224 sizes (B=64): attractor stage at 192x192, the conditional
log-binomial tail 192 -> 384 (GEMM 32->40 + svla_zoe_depth_tail vs the fused kernel), bilinear x2 up-samplings.
CUDA events, L2 flushed between runs.  Usage: python tools/zoe_tail_perf.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spatialvla_b200.ops import CudaOps

dev = "cuda:0"
ops = CudaOps(dev)
BF16, F32 = torch.bfloat16, torch.float32
flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


B, h, oh, nh, nb = 64, 192, 384, 40, 64
x = torch.randn(B * oh * oh, 32, device=dev).to(BF16)
wa = (torch.randn(nh, 32, device=dev) * 0.3).to(BF16)
e = torch.randn(B, h, h, nh, device=dev).to(BF16)
bins = torch.nn.functional.softplus(torch.randn(B, h, h, nb, device=dev))
b1, w2, b2 = torch.randn(nh, device=dev) * 0.1, torch.randn(4, nh, device=dev) * 0.3, torch.tensor([0.0, 0.0, -4.0, 3.0], device=dev)
depth = torch.empty(B, oh, oh, device=dev)
t = torch.empty(B * oh * oh, nh, device=dev, dtype=BF16)


def unfused():
    ops.gemm(x, wa, out_bf16=t)
    ops.zoe_depth_tail(t, e, b1, w2, b2, bins, depth, batch=B, h=h, w=h, oh=oh, ow=oh, nh=nh, nbins=nb, min_temp=0.0212, max_temp=50.0)


ms_u = timeit(unfused)
ms_f = timeit(lambda: ops.zoe_depth_tail_fused(x, wa, e, b1, w2, b2, bins, depth, batch=B, h=h, w=h, oh=oh, ow=oh, min_temp=0.0212, max_temp=50.0))
alg = x.numel() * 2 + e.numel() * 2 + bins.numel() * 4 + depth.numel() * 4
print({"name": "clb_tail_192_to_384", "unfused_gemm_plus_tail_ms": round(ms_u, 3), "fused_ms": round(ms_f, 3),
       "fused_GBs_algorithmic": round(alg / ms_f / 1e6, 1)}, flush=True)

attr = torch.randn(B * h * h, 16, device=dev).to(BF16)
prev = torch.nn.functional.softplus(torch.randn(B, 96, 96, nb, device=dev))
out = torch.empty(B * h * h, nb, device=dev)
ms = timeit(lambda: ops.zoe_attractor(attr, prev, out, batch=B, h=96, w=96, oh=h, ow=h, na=16, nbins=nb))
print({"name": "attractor_96_to_192_na16", "ms": round(ms, 3), "GBs": round((attr.numel() * 2 + prev.numel() * 4 + out.numel() * 4) / ms / 1e6, 1)}, flush=True)

for (hh, c) in ((96, 256), (192, 128)):
    src = torch.randn(B, hh, hh, c, device=dev).to(BF16)
    dst = torch.empty(B, 2 * hh, 2 * hh, c, device=dev, dtype=BF16)
    ms = timeit(lambda: ops.bilinear_nhwc(src, dst, batch=B, h=hh, w=hh, c=c, oh=2 * hh, ow=2 * hh))
    print({"name": f"bilinear_{hh}_to_{2 * hh}_c{c}", "ms": round(ms, 3), "GBs": round((src.numel() + dst.numel()) * 2 / ms / 1e6, 1)}, flush=True)
