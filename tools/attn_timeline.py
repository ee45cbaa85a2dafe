"""In-kernel timeline of ONE CTA of the tcgen05 flash-attention kernel (clock64 stamps of the TMA producer, the MMA issuer and two
softmax warps), on the batch-64 shapes of the 4B-224 path.  Builds a profiling variant of the attention sources with
-DSVLA_ATTN_TIMELINE into gpurun_out/ (the product library carries no instrumentation) and routes svla_attention through it.
Usage: python tools/attn_timeline.py [siglip|beit|gemma]"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from spatialvla_b200.ops import CudaOps
from spatialvla_b200 import _lib

out_dir = os.path.join(ROOT, "gpurun_out"); os.makedirs(out_dir, exist_ok=True)
so = os.path.join(out_dir, "libsvla_attn_timeline.so")
srcs = [os.path.join(_lib.CSRC, s) for s in ("capi.cu", "attention.cu", "attention_tc.cu")]
subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + _lib.NVCC_FLAGS + ["-DSVLA_ATTN_TIMELINE", "-o", so] + srcs, check=True)
var = C.CDLL(so)
var.svla_attention.restype = C.c_int
var.svla_attention.argtypes = [C.POINTER(_lib.SvlaAttnArgs), C.c_void_p]
var.svla_last_error.restype = C.c_char_p

dev = "cuda:0"; ops = CudaOps(dev); BF16 = torch.bfloat16


class Routed:
    def __init__(self, main): self._main = main
    def __getattr__(self, name): return getattr(var if name == "svla_attention" else self._main, name)


ops.lib = Routed(ops.lib)


def dump(name, ntiles):
    buf = (C.c_ulonglong * 2048)()
    torch.cuda.synchronize()
    assert var.svla_dbg_attn_timeline(buf) == 0
    v = list(buf)
    t0 = v[512]
    f = lambda x: (x - t0) if x else -1
    print("====", name, "(SM clocks relative to the issuer's first stamp)")
    for r, role in enumerate(["producer", "issuer", "softmax_w2", "softmax_w6"]):
        base = 512 * r
        print(role, "head [start, setup sync, q_full, loop end, last PV seen, epilogue end, CTA sync, exit]", [f(v[base + i]) for i in range(8)])
        for j in range(ntiles):
            print("   tile", j, [f(v[base + 8 + j * 8 + i]) for i in range(6)])


print("per tile -- producer: [k_empty done, v_empty done]; issuer: [k_full done, -, QK issued, p_full(j)+v_full seen, PV issued]; "
      "softmax: [s_full done, S loaded, scores+max+exchange done, exp done, rescale done, P stored + arrive]")
which = sys.argv[1] if len(sys.argv) > 1 else ""
for name, B, hq, hkv, S, d, kw in [("attn_siglip", 64, 16, 16, 256, 72, {}), ("attn_beit", 64, 16, 16, 577, 64, {"relpos": 24}), ("attn_gemma", 64, 8, 4, 278, 256, {"softcap": 50.0})]:
    if which and which not in name: continue
    D = hq * d
    if hq == hkv:
        qkv = torch.randn(B * S, 3 * D, device=dev).to(BF16); out = torch.empty(B * S, D, device=dev, dtype=BF16)
        tab = torch.randn(hq, (2 * 24 - 1) ** 2 + 3, device=dev) if "relpos" in kw else None      # head-major, as the engine packs it
        st = (S * 3 * D, 3 * D)
        fn = lambda: ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], out, batch=B, hq=hq, hkv=hkv, sq=S, sk=S, d=d, q_strides=st, k_strides=st, v_strides=st, o_strides=(S * D, D), scale=d ** -0.5, relpos_table=tab, relpos_win=24 if tab is not None else 0, relpos_head_major=tab is not None)
    else:
        q = torch.randn(B * S, D, device=dev).to(BF16); kc = torch.randn(B, 290, hkv, d, device=dev).to(BF16); vc = torch.randn_like(kc); out = torch.empty_like(q)
        kvs = (290 * hkv * d, hkv * d)
        fn = lambda: ops.attention(q, kc, vc, out, batch=B, hq=hq, hkv=hkv, sq=S, sk=S, d=d, q_strides=(S * D, D), k_strides=kvs, v_strides=kvs, o_strides=(S * D, D), scale=1 / 16, softcap=50.0, causal=True)
    fn(); fn(); torch.cuda.synchronize()
    dump(name, (S + 63) // 64)
