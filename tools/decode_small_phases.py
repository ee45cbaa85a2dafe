"""Phase timing of the persistent small-batch decode kernel (svla_decode_step_small): CTA 0's globaltimer stamps in layer 5.
Usage: SVLA_DECODE_SMALL_TIMING=1 python tools/decode_small_phases.py [batch]"""
import os, sys
os.environ.setdefault("SVLA_DECODE_SMALL_TIMING", "1")
os.environ.setdefault("SVLA_DECODE_SMALL", "1")            # the persistent kernel is opt-in
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spatialvla_b200 import get_config_dict
from spatialvla_b200.engine import SpatialVLAEngine
from spatialvla_b200.ops import CudaOps
from spatialvla_b200.weights import synth_state_dict

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = dict(get_config_dict("4b-224"), use_vision_zoe=False)
sd = synth_state_dict(cfg, seed=0, device=dev, on_device_rng=True, dtype=torch.bfloat16)
ops = CudaOps(dev)
eng = SpatialVLAEngine(cfg, sd, ops)
del sd
P = 278
g = torch.Generator().manual_seed(0)
ids = torch.randint(3, 250000, (B, P), generator=g).to(dev)
x, _ = eng.embed(ids)
cache = eng.new_cache(B, P + 12)
eng.gemma_forward(x, B, P, cache, bidirectional=True)
tok = torch.randint(cfg["action_token_begin_idx"], cfg["action_token_begin_idx"] + 8194, (B, 1), generator=g).to(dev)
names = ["A norm", "A gemv qkv", "barrier", "B attention", "barrier", "C combine", "C gemv o", "barrier", "D norm", "D gemv gate/up",
         "barrier", "E load act", "E gemv down", "barrier"]
for rep in range(3):
    cache["len"] = P
    xx, _ = eng.embed(tok)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.gemma_decode_small(xx, B, cache)
    e1.record()
    torch.cuda.synchronize()
    n = int(ops.lib.svla_decode_step_small_scratch_floats(B, 2304, 8, 4, 256, 9216))
    sc = cache["small_scratch"]
    ts = sc[n - 64 - 0:].view(torch.int64)[: 32].cpu().tolist() if False else sc[n - 64:].contiguous().view(torch.int64)[:16].cpu().tolist()
    d = [(ts[i + 1] - ts[i]) / 1e3 for i in range(len(names))]
    print(f"rep {rep}: step {e0.elapsed_time(e1) * 1e3:.0f} us; layer 5 phases (us): " + ", ".join(f"{nm} {v:.1f}" for nm, v in zip(names, d)) +
          f" | layer total {sum(d):.1f}")
