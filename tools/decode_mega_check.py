"""Persistent decode kernel (csrc/decode_mega.cu) vs the 7-launches-per-layer chain on one B200: same prefilled cache, same
embedded token, one decode step each way -> hidden states / appended cache rows must be bit-identical; then in-situ timing of
both (CUDA-graph replay of embed + step + action head + argmax, median of `reps`).
Usage: python tools/decode_mega_check.py [config=4b-224] [batches=64,1] [reps=20]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spatialvla_b200 import get_config_dict
from spatialvla_b200.engine import SpatialVLAEngine
from spatialvla_b200.ops import CudaOps
from spatialvla_b200.weights import synth_state_dict

dev = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "4b-224"
batches = [int(b) for b in (sys.argv[2] if len(sys.argv) > 2 else "64,1").split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cfg = dict(get_config_dict(name), use_vision_zoe=False)
sd = synth_state_dict(cfg, seed=0, device=dev, on_device_rng=True, dtype=torch.bfloat16)
ops = CudaOps(dev)
eng = SpatialVLAEngine(cfg, sd, ops)
del sd
torch.cuda.empty_cache()
eng.decode_hilo = False          # the persistent kernel implements the plain bf16 chain (no hi/lo activation pairs)
t = cfg["text_config"]
H = t["hidden_size"]
P = 278 if name != "tiny" else 70
g = torch.Generator().manual_seed(0)
lo, n_act = cfg["action_token_begin_idx"], cfg["spatial_token_num"]
ok_all = True
for B in batches:
    ids = torch.randint(3, lo - 10, (B, P), generator=g).to(dev)
    pads = None
    x, _ = eng.embed(ids)
    cache = eng.new_cache(B, P + 12)
    cache.pop("small_table", None)
    eng.gemma_forward(x, B, P, cache, bidirectional=True)
    tok = torch.randint(lo, lo + n_act, (B, 1), generator=g).to(dev)
    res = {}
    for mode in ("chain", "mega"):
        eng.mega_decode = mode == "mega"
        c2 = {"k": cache["k"].clone(), "v": cache["v"].clone(), "smax": cache["smax"], "len": P}
        hs = []
        for step in range(3):                       # three consecutive steps: the appended rows feed the next step
            xx, _ = eng.embed(tok)
            hs.append(eng.gemma_forward(xx, B, 1, c2, bidirectional=False).float().clone())
        torch.cuda.synchronize()
        res[mode] = (torch.stack(hs), c2["k"][:, :, P:P + 3].float().clone(), c2["v"][:, :, P:P + 3].float().clone())
    dh = (res["chain"][0] - res["mega"][0]).abs().max().item()
    dk = (res["chain"][1] - res["mega"][1]).abs().max().item()
    dv = (res["chain"][2] - res["mega"][2]).abs().max().item()
    ref = res["chain"][0].abs().max().item()
    same = dh == 0.0 and dk == 0.0 and dv == 0.0
    ok_all &= same
    print({"check": "mega_vs_chain", "config": name, "batch": B, "max_abs_h": dh, "max_abs_k": dk, "max_abs_v": dv, "h_scale": ref,
           "bit_identical": same, "finite": bool(torch.isfinite(res["mega"][0]).all())}, flush=True)

    out_tok = torch.zeros(B, dtype=torch.int64, device=dev)

    def step():
        cache["len"] = P
        xx, _ = eng.embed(tok)
        rows = eng.gemma_forward(xx, B, 1, cache, bidirectional=False)
        lg = eng.action_logits(rows, B)
        ops.argmax_rows(lg, out_tok, id_offset=lo)

    for mode in ("chain", "mega"):
        eng.mega_decode = mode == "mega"
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
        L_, nh, nkv, hd, FF = t["num_hidden_layers"], t["num_attention_heads"], t["num_key_value_heads"], t["head_dim"], t["intermediate_size"]
        w_bytes = L_ * 2 * (H * (nh + 2 * nkv) * hd + nh * hd * H + 3 * H * FF) + n_act * H * 2
        kv_bytes = L_ * 2 * B * (P + 1) * nkv * hd * 2
        print({"name": "decode_step", "mode": mode, "config": name, "batch": B, "launches": launches, "us_median": round(ms * 1e3, 1),
               "us_min": round(ts[0] * 1e3, 1), "GBs": round((w_bytes + kv_bytes) / ms / 1e6, 1),
               "roofline_us_at_6555GBs": round((w_bytes + kv_bytes) / 6555.8e3, 1)}, flush=True)
    if os.environ.get("SVLA_DECODE_MEGA_TIMING") is not None:
        eng.mega_decode = True
        step()
        torch.cuda.synchronize()
        plan = eng._mega_plan()
        tm = plan["scratch"][-64 * 8:].view(torch.int64).cpu().tolist()
        names = {0: "n0_end", 1: "attn_start", 2: "attn_end", 3: "norm1_start", 4: "norm1_end", 5: "norm2_start", 6: "norm2_end",
                 8: "qkv_end", 9: "o_end", 10: "gateup_end", 11: "down_end"}
        names.update({19: "attn_item0_done"})
        for gi, gn in enumerate(("qkv", "o", "gateup", "down")):
            names.update({20 + 4 * gi: gn + "_epoch_seen", 21 + 4 * gi: gn + "_acc_ready", 22 + 4 * gi: gn + "_stored"})
        order = [20, 21, 22, 8, 1, 19, 2, 24, 25, 26, 9, 3, 4, 28, 29, 30, 10, 32, 33, 34, 11, 5, 6]
        base = tm[8]
        print({"layer_timeline_us(cta0)": {names[i]: round((tm[i] - base) / 1e3, 2) for i in order}}, flush=True)
print("MEGA_CHECK", "OK" if ok_all else "MISMATCH")
