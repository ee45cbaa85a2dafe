#!/usr/bin/env python
"""Where does the bf16 logit noise come from?  B = 4 observations at full size: teacher-forced action-slice logits of the GPU path
(a) end to end, (b) with the ORACLE's fp32 image features fed to the GPU language stage, against the fp32 oracle."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_ref as R  # noqa: E402
from oracle.gen_golden_full import full_inputs  # noqa: E402
from spatialvla_b200.configs import get_config_dict  # noqa: E402
from spatialvla_b200.engine import SpatialVLAEngine  # noqa: E402
from spatialvla_b200.ops import CudaOps  # noqa: E402
from spatialvla_b200.weights import synth_state_dict  # noqa: E402


def main():
    dev = "cuda:0"
    cfg = get_config_dict("4b-224")
    sd = synth_state_dict(cfg, seed=0)
    px, ids, K, _ = full_inputs(cfg)
    B, n_new = 4, 13
    px, ids = px[:B], ids[:B]
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    ref_toks, ref_logits, raux = R.predict_action_ref(sd, cfg, ids, px, K, n_new, force_head=0, return_aux=True)
    print(f"oracle {time.time() - t0:.0f}s")
    eng = SpatialVLAEngine(cfg, sd, CudaOps(dev))
    eng.force_head = 0
    with torch.no_grad():
        feats, aux = eng.image_features(px.to(dev), K.to(dev), return_aux=True)
        _, lg_a = eng.generate_actions(ids.to(dev), px.to(dev), K.to(dev), n_new, forced_tokens=ref_toks.to(dev), return_logits=True)
        logs = []
        eng.language_stage(ids.to(dev), raux["image_features"].to(dev).contiguous(), n_new, forced_tokens=ref_toks.to(dev), logs=logs)
        lg_b = torch.stack(logs, 1)
        # (c) oracle depth -> GPU Ego3D / projector: isolates ZoeDepth
    for name, lg in (("end to end", lg_a), ("oracle image features -> GPU Gemma2", lg_b)):
        d = (lg.cpu() - ref_logits).abs()
        print(f"{name:40s}: logits rms {float(d.pow(2).mean().sqrt()):.5f} max {float(d.max()):.5f}")
    fe = (feats.cpu() - raux["image_features"]).abs()
    print(f"image features: rel max {float(fe.max() / raux['image_features'].abs().max()):.5f} rms {float(fe.pow(2).mean().sqrt()):.6f} "
          f"(ref rms {float(raux['image_features'].pow(2).mean().sqrt()):.5f})")
    print(f"siglip rel max {float((aux['siglip'].view_as(raux['siglip']).cpu() - raux['siglip']).abs().max() / raux['siglip'].abs().max()):.5f}; "
          f"depth max |d| {float((aux['depth384'].cpu() - raux['depth384']).abs().max()):.5f} m; "
          f"pos3d-free feature check: xyz max |d| {float((aux['xyz'].cpu() - raux['xyz']).abs().max()):.5f}")


if __name__ == "__main__":
    main()
