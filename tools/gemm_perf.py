"""GEMM / conv / attention micro-benchmark on the shapes of the 4B-224 path at batch 64 (CUDA events, L2 flushed)."""
import json, os, sys, time
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

rows = []
shapes = [("siglip_qkv", 16384, 3456, 1152), ("siglip_out", 16384, 1152, 1152), ("siglip_fc1", 16384, 4304, 1152), ("siglip_fc2", 16384, 1152, 4304),
          ("beit_qkv", 36928, 3072, 1024), ("beit_out", 36928, 1024, 1024), ("beit_fc1", 36928, 4096, 1024), ("beit_fc2", 36928, 1024, 4096),
          ("gemma_qkv", 17792, 4096, 2304), ("gemma_o", 17792, 2304, 2048), ("gemma_gateup", 17792, 18432, 2304), ("gemma_down", 17792, 2304, 9216),
          ("dec_qkv", 64, 4096, 2304), ("dec_o", 64, 2304, 2048), ("dec_gateup", 64, 18432, 2304), ("dec_down", 64, 2304, 9216), ("dec_head", 64, 8194, 2304),
          ("cublas_ref_8192", 8192, 8192, 8192)]
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, M, N, K in shapes:
    if only and only not in name: continue
    a = torch.randn(M, K, device=dev).to(BF16); w = torch.randn(N, K, device=dev).to(BF16)
    out = torch.empty(M, N, device=dev, dtype=BF16)
    for bn, impl in ([(0, 0)] if M <= 64 else [(128, 2), (256, 2), (128, 3), (256, 3)]):
        ms = timeit(lambda: ops.gemm(a, w, out_bf16=out, block_n=bn, impl=impl))
        tf = 2.0 * M * N * K / ms / 1e9
        gbs = (M * K + N * K + M * N) * 2 / ms / 1e6
        rows.append({"name": name, "M": M, "N": N, "K": K, "bn": bn, "cta_group": {0: "auto", 2: 1, 3: 2}[impl], "ms": round(ms, 4), "tflops": round(tf, 1), "gbs": round(gbs, 1)})
        print(rows[-1], flush=True)
    if name.startswith("cublas") or name in ("gemma_gateup", "beit_fc1"):
        ms = timeit(lambda: torch.matmul(a, w.t()))
        print({"name": name + "_torch_matmul", "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}, flush=True)
# mainloop probes: one tile per SM (148 tiles) and a single tile, K sweep -> time per 64-wide k-block
for name, M, N in [("probe_148tiles_bn256", 128 * 148, 256), ("probe_1tile_bn256", 128, 256), ("probe_148tiles_m64_bn128", 64, 128 * 148)]:
    if only and only not in name: continue
    for K in (1024, 4096, 16384):
        a = torch.randn(M, K, device=dev).to(BF16); w = torch.randn(N, K, device=dev).to(BF16)
        out = torch.empty(M, N, device=dev, dtype=BF16)
        bn = 128 if "bn128" in name else 256
        ms = timeit(lambda: ops.gemm(a, w, out_bf16=out, block_n=bn, impl=2))
        print({"name": name, "K": K, "ms": round(ms, 4), "us_per_kblock": round(ms * 1e3 / (K / 64), 3), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}, flush=True)
# conv
for name, (nb, h, w_, c), N in [("fusion_conv_96", (64, 96, 96, 256), 256), ("fusion_conv_48", (64, 48, 48, 256), 256), ("rel_conv1_192", (64, 192, 192, 256), 128), ("rel_conv2_384", (64, 384, 384, 128), 32)]:
    if only and only not in name: continue
    x = torch.randn(nb, h, w_, c, device=dev).to(BF16); wt = torch.randn(N, 9 * c, device=dev).to(BF16)
    out = torch.empty(nb * h * w_, N, device=dev, dtype=BF16)
    for impl in (2, 3, 4):
        if impl == 4 and N > 128: continue
        ms = timeit(lambda: ops.gemm(x, wt, conv=(nb, h, w_, c), out_bf16=out, impl=impl), iters=3)
        print({"name": name, "kernel": {2: "patch-tile 1-CTA", 3: "patch-tile 2-CTA", 4: "row-tile"}[impl], "ms": round(ms, 3),
               "tflops": round(2.0 * nb * h * w_ * N * 9 * c / ms / 1e9, 1)}, flush=True)
# attention
for name, B, hq, hkv, S, d, kw in [("attn_siglip", 64, 16, 16, 256, 72, {}), ("attn_beit", 64, 16, 16, 577, 64, {"relpos": 24}), ("attn_gemma", 64, 8, 4, 278, 256, {"softcap": 50.0})]:
    if only and only not in name: continue
    D = hq * d
    if hq == hkv:
        qkv = torch.randn(B * S, 3 * D, device=dev).to(BF16); out = torch.empty(B * S, D, device=dev, dtype=BF16)
        tab = torch.randn(hq, (2 * 24 - 1) ** 2 + 3, device=dev) if "relpos" in kw else None      # head-major, as the engine packs it
        st = (S * 3 * D, 3 * D)
        fn = lambda: ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], out, batch=B, hq=hq, hkv=hkv, sq=S, sk=S, d=d, q_strides=st, k_strides=st, v_strides=st, o_strides=(S * D, D), scale=d ** -0.5, relpos_table=tab, relpos_win=24 if tab is not None else 0, relpos_head_major=tab is not None)
    else:
        q = torch.randn(B * S, D, device=dev).to(BF16); kc = torch.randn(B, 290, hkv, d, device=dev).to(BF16); vc = torch.randn_like(kc); out = torch.empty_like(q)
        kvs = (290 * hkv * d, hkv * d)
        fn = lambda: ops.attention(q, kc, vc, out, batch=B, hq=hq, hkv=hkv, sq=S, sk=S, d=d, q_strides=(S * D, D), k_strides=kvs, v_strides=kvs, o_strides=(S * D, D), scale=1 / 16, softcap=50.0)
    ms = timeit(fn)
    print({"name": name, "ms": round(ms, 3), "tflops": round(4.0 * B * hq * S * S * d / ms / 1e9, 1)}, flush=True)

# ---- epilogue variants on the BEiT fc1 shape
from spatialvla_b200._lib import ACT_GELU_ERF, ACT_GELU_TANH, ACT_RELU
if not only or "epi" in only:
    only_tag = sys.argv[2] if len(sys.argv) > 2 else ""
    M, N, K = (int(v) for v in sys.argv[3:6]) if len(sys.argv) >= 6 else (36928, 4096, 1024)     # epi <tag|''> M N K
    a = torch.randn(M, K, device=dev).to(BF16); w = (torch.randn(N, K, device=dev) / 32).to(BF16)
    bias = torch.randn(N, device=dev); cs = torch.rand(N, device=dev)
    outb = torch.empty(M, N, device=dev, dtype=BF16); outf = torch.zeros(M, N, device=dev, dtype=F32)
    for tag, kw in [("plain_bf16", dict(out_bf16=outb)), ("bias", dict(out_bf16=outb, bias=bias)), ("erf", dict(out_bf16=outb, act=ACT_GELU_ERF)),
                    ("bias_erf", dict(out_bf16=outb, bias=bias, act=ACT_GELU_ERF)), ("bias_tanh", dict(out_bf16=outb, bias=bias, act=ACT_GELU_TANH)),
                    ("relu", dict(out_bf16=outb, act=ACT_RELU)), ("f32_out", dict(out_f32=outf)), ("f32_accum_colscale", dict(out_f32=outf, accumulate=True, colscale=cs, bias=bias))]:
        if only_tag and tag != only_tag: continue
        ms = timeit(lambda: ops.gemm(a, w, block_n=int(os.environ.get("BN", "256")), **kw))
        print({"name": "epi_" + tag, "M": M, "N": N, "K": K, "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}, flush=True)

# ---- decode kernels measured from CUDA-graph replays (true GPU time, no host launch gaps)
def graph_time(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps

if not only or "dec" in only:
    for name, N, K, geglu in [("dec_qkv", 4096, 2304, False), ("dec_o", 2304, 2048, False), ("dec_gateup", 18432, 2304, True), ("dec_down", 2304, 9216, False), ("dec_head", 8194, 2304, False)]:
        x = torch.randn(64, K, device=dev).to(BF16)
        ws = [(torch.randn(N, K, device=dev) / 48).to(BF16) for _ in range(20)]      # distinct weights per call: no L2 reuse
        S = ops.skinny_splits(N, K)
        it = [0]
        if geglu:
            out = torch.empty(64, N // 2, device=dev, dtype=BF16)
            def fn():
                ops.gemm_skinny(x, ws[it[0] % 20], out_bf16=out, geglu=True); it[0] += 1
        else:
            out = torch.empty(S, 64, N, device=dev, dtype=F32)
            def fn():
                ops.gemm_skinny(x, ws[it[0] % 20], out_f32=out); it[0] += 1
        us = graph_time(fn) * 1e3
        print({"name": name + "_skinny_graph", "splits": S, "us": round(us, 2), "gbs": round(N * K * 2 / us / 1e3, 1)}, flush=True)
    q = torch.randn(64, 2048, device=dev).to(BF16); kc = torch.randn(20, 64, 290, 4, 256, device=dev).to(BF16); vc = torch.randn_like(kc); o = torch.empty_like(q)
    it = [0]
    def fn():
        ops.decode_attention(q, kc[it[0] % 20], vc[it[0] % 20], o, batch=64, hq=8, hkv=4, d=256, smax=290, ctx=285, scale=1 / 16, softcap=50.0); it[0] += 1
    us = graph_time(fn) * 1e3
    print({"name": "decode_attention_graph", "us": round(us, 2), "gbs": round(64 * 285 * 4 * 256 * 2 * 2 / us / 1e3, 1)}, flush=True)
    xs = torch.randn(64, 2304, device=dev); br = torch.randn(8, 64, 2304, device=dev); w1 = torch.randn(2304, device=dev); h = torch.empty(64, 2304, device=dev, dtype=BF16)
    us = graph_time(lambda: ops.rmsnorm_residual(xs, branch=br, w_post=w1, w_pre=w1, out_bf16=h)) * 1e3
    print({"name": "rmsnorm_residual_decode_graph", "us": round(us, 2)}, flush=True)
    qkvp = torch.randn(4, 64, 4096, device=dev); qq = torch.empty(64, 2048, device=dev, dtype=BF16)
    us = graph_time(lambda: ops.rope_kv(qkvp, qq, kc[0], vc[0], batch=64, s=1, hq=8, hkv=4, d=256, smax=290, pos0=280, theta=10000.0)) * 1e3
    print({"name": "rope_kv_decode_graph", "us": round(us, 2)}, flush=True)
