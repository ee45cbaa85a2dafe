"""Forward half of the LoRA training step at the size of BASELINE.json config #5: SpatialVLA-4B-224, per-GPU batch 32,
sequence = prefix 278 + 12 action ids + EOS (L = 291), token types 0/1, labels -100 on the prefix; prefix-LM mask.
Times forward(labels=...) end to end (host inputs, CUDA events, L2 flushed) and its loss tail (labelled-row gather, full-vocabulary
lm_head GEMM with soft-cap, cross-entropy kernel) on the launching stream.  Usage: python tools/loss_forward_perf.py [--batch 32]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import synth_inputs
from spatialvla_b200 import get_config_dict
from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
from spatialvla_b200.weights import synth_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--config", default="4b-224")
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()

dev = "cuda:0"
F32 = torch.float32
cfg = get_config_dict(args.config)
sd = synth_state_dict(cfg, seed=0, device=dev, on_device_rng=True, dtype=torch.bfloat16)
model = SpatialVLAForConditionalGeneration(cfg, sd, device=dev)
del sd
torch.cuda.empty_cache()
eng, ops = model.engine, model.ops
B = args.batch
px, ids, K = synth_inputs(cfg, B)
g = torch.Generator().manual_seed(5)
lo = cfg["action_token_begin_idx"]
suffix = torch.cat([torch.randint(lo, lo + cfg["spatial_token_num"], (B, 12), generator=g), torch.full((B, 1), cfg["eos_token_id"])], 1)
full = torch.cat([ids, suffix], 1)
P, L = ids.shape[1], full.shape[1]
tt = torch.cat([torch.zeros(B, P, dtype=torch.int64), torch.ones(B, L - P, dtype=torch.int64)], 1)
labels = torch.where(tt == 1, full, torch.full_like(full, -100))
ones = torch.ones(B, L, dtype=torch.int64)
flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


out = {}


def fwd():
    out["o"] = model.forward(input_ids=full, pixel_values=px, intrinsic=K, labels=labels, token_type_ids=tt, attention_mask=ones)


n0 = ops.launch_count()
fwd()
torch.cuda.synchronize()
launches = ops.launch_count() - n0
ms_fwd = timed(fwd, args.iters)
o = out["o"]
R, V, H = o.label_rows.numel(), eng.t["vocab_size"], eng.t["hidden_size"]

# the loss tail alone, on resident hidden states
h = torch.randn(B * L, H, device=dev).to(torch.bfloat16)
rows = o.label_rows
lab = labels[:, 1:][labels[:, 1:] != -100].to(dev).contiguous()
ms_tail = timed(lambda: eng.labelled_loss(h, rows, lab), args.iters)
lg = ops.empty((R, V), F32)
w = eng.lm_head_full()
from spatialvla_b200._lib import ACT_SOFTCAP
hr = h.index_select(0, rows)
ms_gemm = timed(lambda: ops.gemm(hr, w, out_f32=lg, act=ACT_SOFTCAP, act_param=30.0), args.iters)
rl, ra, sm = ops.empty((R,), F32), ops.empty((R,), torch.int64), ops.empty((3,), F32)
ms_ce = timed(lambda: ops.cross_entropy_rows(lg, lab, rl, ra, summary=sm), args.iters)
# backward of the loss tail: cross-entropy backward kernel (fp32 logits in, bf16 dz out) + dh = dz @ W_head (K = vocabulary)
wt = eng.lm_head_full_t()
dz = ops.empty((R, wt.shape[1]), torch.bfloat16)
ms_ce_bwd = timed(lambda: ops.cross_entropy_bwd(lg, lab, rl, sm, dz, softcap=30.0), args.iters)
dh = ops.empty((R, H), F32)
ms_dh = timed(lambda: ops.gemm(dz, wt, out_f32=dh), args.iters)
ms_dh64 = timed(lambda: ops.gemm(dz, wt, out_f32=dh, block_n=64), args.iters)
ms_tail_fb = timed(lambda: eng.labelled_loss_backward(h, rows, lab), args.iters)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
print(json.dumps({
    "workload": f"forward(labels) {args.config} B={B} L={L} prefix={P} prefix-LM mask, {R} labelled rows x V={V}",
    "loss": round(float(o.loss), 5), "token_accuracy": round(float(o.token_accuracy), 5), "launches": launches,
    "forward_ms_p50": round(ms_fwd, 3), "samples_per_s": round(B / ms_fwd * 1e3, 1),
    "loss_tail_ms": round(ms_tail, 3),
    "lm_head_gemm_ms": round(ms_gemm, 3), "lm_head_TFLOPs": round(2.0 * R * V * H / ms_gemm / 1e9, 1),
    "cross_entropy_ms": round(ms_ce, 3), "cross_entropy_GBs": round(R * V * 4 / ms_ce / 1e6, 1),
    "cross_entropy_algorithmic_bytes": R * V * 4,
    "loss_tail_fwd_bwd_ms": round(ms_tail_fb, 3),
    "cross_entropy_bwd_ms": round(ms_ce_bwd, 3), "cross_entropy_bwd_GBs": round(R * V * 6 / ms_ce_bwd / 1e6, 1),
    "dh_gemm_ms": round(ms_dh, 3), "dh_gemm_TFLOPs": round(2.0 * R * V * H / ms_dh / 1e9, 1),
    "dh_gemm_bn64_ms": round(ms_dh64, 3), "dh_gemm_bn64_TFLOPs": round(2.0 * R * V * H / ms_dh64 / 1e9, 1),
    "dh_gemm_note": "M=%d N=%d K=%d: 72 output tiles at the default BN=128, 144 at BN=64 (what the engine uses); no split-K yet" % (R, H, V), "hbm_peak_GBs": peaks.get("hbm_gbs"),
    "note": "CUDA events on the launching stream, L2 flushed (256 MiB write) before every repetition, p50 of %d" % args.iters,
}), flush=True)
