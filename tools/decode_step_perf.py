"""In-situ timing of ONE Gemma2 decode step (B=64, ctx 279) exactly as predict_action runs it: embed + 26 layers x
[qkv skinny GEMM, fused RoPE/cache/attention, o skinny GEMM, sandwich norm, gate/up skinny GEMM, down skinny GEMM, sandwich
norm] + action-slice head + argmax, captured in a CUDA graph (PDL chain) and replayed; CUDA events around the replays, the
step's 8.1 GB of weights >> L2.  Prints us/step and GB/s against the algorithmic bytes of SURVEY.md §8d.
Usage: python tools/decode_step_perf.py [reps]     env: SVLA_SKINNY_STAGES, SVLA_PDL"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spatialvla_b200 import get_config_dict
from spatialvla_b200.engine import SpatialVLAEngine
from spatialvla_b200.ops import CudaOps
from spatialvla_b200.weights import synth_state_dict

dev = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = get_config_dict("4b-224")
cfg = dict(cfg, use_vision_zoe=False)          # language model only (the vision towers are not part of a decode step)
sd = synth_state_dict(cfg, seed=0, device=dev, on_device_rng=True, dtype=torch.bfloat16)
ops = CudaOps(dev)
eng = SpatialVLAEngine(cfg, sd, ops)
del sd
torch.cuda.empty_cache()
B, P, H = int(os.environ.get("B", "64")), 278, cfg["text_config"]["hidden_size"]
g = torch.Generator().manual_seed(0)
ids = torch.randint(3, 250000, (B, P), generator=g).to(dev)
x, _ = eng.embed(ids)
cache = eng.new_cache(B, P + 12)
eng.gemma_forward(x, B, P, cache, bidirectional=True)
tok = torch.randint(cfg["action_token_begin_idx"], cfg["action_token_begin_idx"] + 8194, (B, 1), generator=g).to(dev)
out_tok = torch.zeros(B, dtype=torch.int64, device=dev)


def step():
    cache["len"] = P
    xx, _ = eng.embed(tok)
    rows = eng.gemma_forward(xx, B, 1, cache, bidirectional=False, hilo_out=True)
    lg = eng.action_logits(rows, B)
    ops.argmax_rows(lg, out_tok, id_offset=cfg["action_token_begin_idx"])


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step()
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        step()
    launches = ops.launch_count() - n0
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
t = cfg["text_config"]
L_, nh, nkv, hd, FF = t["num_hidden_layers"], t["num_attention_heads"], t["num_key_value_heads"], t["head_dim"], t["intermediate_size"]
w_bytes = L_ * 2 * (H * (nh + 2 * nkv) * hd + nh * hd * H + 3 * H * FF) + 8194 * H * 2
kv_bytes = L_ * 2 * B * (P + 1) * nkv * hd * 2
print({"name": "decode_step", "batch": B, "hilo": eng.decode_hilo, "stages_env": os.environ.get("SVLA_SKINNY_STAGES", ""), "pdl": os.environ.get("SVLA_PDL", "1"),
       "launches": launches, "us_median": round(ms * 1e3, 1), "us_min": round(ts[0] * 1e3, 1),
       "weight_GB": round(w_bytes / 1e9, 3), "kv_GB": round(kv_bytes / 1e9, 3),
       "GBs": round((w_bytes + kv_bytes) / ms / 1e6, 1), "roofline_us_at_6555GBs": round((w_bytes + kv_bytes) / 6555.8e3, 1)}, flush=True)
