"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/full_4b_b64.npz: the fp32 oracle (oracle/model_ref.py, pinned on the live reference
at the tiny size) run OFFLINE on the full SpatialVLA-4B-224 configuration for the 64 observations of BASELINE.json config #2, in
chunks of 8 (64 x 13 = 832 teacher-forced positions: the last prompt position + 12 decode positions per sample).
Stored per position: the oracle's greedy token, its top-1 / top-2 logits (margin) and the post-softcap logits at 96 fixed action
columns + the arg-max column; plus the calibration the parity report prints beside the new implementation's agreement
(BASELINE.md §5): the SAME oracle run in bf16 arithmetic (torch.autocast on the CPU) against its own fp32 run on the first chunk.
ZoeDepth's metric head is pinned to 0 for every chunk (the router votes over the batch; the GPU test pins the same head).
Usage: python -m oracle.gen_golden_full   (~15 min on 8 cores, 20 GB of RAM)"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import model_ref as R  # noqa: E402
from spatialvla_b200.configs import default_intrinsic_224, get_config_dict  # noqa: E402
from spatialvla_b200.weights import synth_state_dict  # noqa: E402

B_TOTAL, CHUNK, N_NEW, SEED, N_COLS = 64, 8, 13, 11, 96


def full_inputs(cfg, B=B_TOTAL, seed=SEED):
    g = torch.Generator().manual_seed(seed)
    px = torch.rand(B, 3, 224, 224, generator=g)
    ids = torch.cat([torch.full((B, 256), cfg["image_token_index"]), torch.full((B, 1), 2),
                     torch.randint(3, 250000, (B, 20), generator=g), torch.full((B, 1), 108)], 1)
    K = torch.tensor(default_intrinsic_224())
    cols = torch.randperm(cfg["spatial_token_num"], generator=g)[:N_COLS].sort().values
    return px, ids, K, cols


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = get_config_dict("4b-224")
    t0 = time.time()
    sd = synth_state_dict(cfg, seed=0)
    print(f"weights {time.time() - t0:.0f}s", flush=True)
    px, ids, K, cols = full_inputs(cfg)
    toks, top2, sub, dlog = [], [], [], []
    calib = None
    for c0 in range(0, B_TOTAL, CHUNK):
        t1 = time.time()
        sl = slice(c0, c0 + CHUNK)
        tk, lg, aux = R.predict_action_ref(sd, cfg, ids[sl], px[sl], K, N_NEW, force_head=0, return_aux=True)
        toks.append(tk)
        top2.append(lg.topk(2, -1).values)
        sub.append(lg[..., cols])
        dlog.append(aux["zoe"]["domain_logits"])
        print(f"chunk {c0 // CHUNK}: {time.time() - t1:.0f}s", flush=True)
        if c0 == 0:
            t2 = time.time()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                tk16, lg16 = R.predict_action_ref(sd, cfg, ids[sl], px[sl], K, N_NEW, force_head=0, forced_tokens=tk)
            lg16 = lg16.float()
            d = (lg16 - lg).abs()
            m = lg.topk(2, -1).values
            calib = np.array([float((tk16 == tk).float().mean()), float(d.mean()), float((d <= 2e-2 + 2e-2 * lg.abs()).float().mean()),
                              float((m[..., 0] - m[..., 1]).median()), float(tk.numel())])
            print(f"calibration (oracle bf16-autocast vs fp32, {tk.numel()} positions): agreement {calib[0]:.4f}, mean|dlogit| {calib[1]:.4f}, "
                  f"within tolerance {calib[2]:.4f}, median margin {calib[3]:.4f} ({time.time() - t2:.0f}s)", flush=True)
    out = os.path.join(ROOT, "tests", "golden", "full_4b_b64.npz")
    np.savez_compressed(out, tokens=torch.cat(toks).numpy(), top2=torch.cat(top2).numpy().astype(np.float32),
                        logits_sub=torch.cat(sub).numpy().astype(np.float32), cols=cols.numpy(), domain_logits=torch.cat(dlog).numpy(),
                        calibration=calib, seed=np.int64(SEED), n_new=np.int64(N_NEW))
    print("wrote", out, f"{time.time() - t0:.0f}s total")


if __name__ == "__main__":
    main()
