#!/usr/bin/env python
"""Parity report of the B200 path against the fp32 oracle at the sizes north_star quotes (run on the GPU box; the oracle side is the
committed golden tests/golden/full_4b_b64.npz minted offline by oracle/gen_golden_full.py):
  * SpatialVLA-4B-224, batch 64, teacher-forced on the oracle's tokens: 64 x 13 = 832 positions -- raw action-slice
    argmax agreement (gate >= 99.5 %), element-wise logit coverage |d| <= 2e-2 + 2e-2 |ref|, with the calibration line
    (the oracle's own bf16 run against its fp32 run) printed beside it;
  * tokenizer: mismatching rows on the golden vectors and on 1 M random actions (gate == 0), decode ulp.
Used by tests/test_e2e_gpu.py::test_full_size_parity_832_positions and committed as profiles/parity_r2_*.txt."""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def full_size_parity(device="cuda:0", log=print):
    from oracle.gen_golden_full import N_NEW, full_inputs
    from spatialvla_b200.configs import get_config_dict
    from spatialvla_b200.engine import SpatialVLAEngine
    from spatialvla_b200.ops import CudaOps
    from spatialvla_b200.weights import synth_state_dict
    g = np.load(os.path.join(GOLD, "full_4b_b64.npz"))
    cfg = get_config_dict("4b-224")
    t0 = time.time()
    sd = synth_state_dict(cfg, seed=0)                       # the weights the oracle used (bf16-rounded, CPU generator)
    px, ids, K, cols = full_inputs(cfg)
    assert np.array_equal(cols.numpy(), g["cols"])
    eng = SpatialVLAEngine(cfg, sd, CudaOps(device))
    del sd
    ref_toks = torch.from_numpy(g["tokens"])
    B = ref_toks.shape[0]
    eng.force_head = 0                                       # the golden pins metric head 0 (chunked oracle; the router votes per batch)
    with torch.no_grad():
        toks, logits = eng.generate_actions(ids.to(device), px.to(device), K.to(device), N_NEW, forced_tokens=ref_toks.to(device),
                                            return_logits=True)
    eng.force_head = None
    toks, logits = toks.cpu(), logits.cpu()
    agree = (toks == ref_toks)
    top2 = torch.from_numpy(g["top2"])
    margin = top2[..., 0] - top2[..., 1]
    sub_ref = torch.from_numpy(g["logits_sub"])
    sub = logits[..., cols]
    d = (sub - sub_ref).abs()
    cover = float((d <= 2e-2 + 2e-2 * sub_ref.abs()).float().mean())
    top1 = logits.gather(-1, (ref_toks - cfg["action_token_begin_idx"]).unsqueeze(-1)).squeeze(-1)
    d1 = (top1 - top2[..., 0]).abs()
    cal = g["calibration"]
    res = {"positions": int(agree.numel()), "agreement": float(agree.float().mean()), "mismatches": int((~agree).sum()),
           "mismatch_margins": [round(float(m), 5) for m in margin[~agree]], "logit_cover": cover, "logit_max_abs": float(d.max()),
           "logit_rms": float(d.pow(2).mean().sqrt()), "top1_max_abs": float(d1.max()), "median_margin": float(margin.median()),
           "calibration": {"agreement": float(cal[0]), "mean_abs_dlogit": float(cal[1]), "cover": float(cal[2]), "median_margin": float(cal[3]),
                           "positions": int(cal[4])}, "seconds": round(time.time() - t0, 1)}
    log(f"[4B-224 parity, batch {B}, teacher-forced on the oracle's tokens] {res['positions']} positions: action-slice argmax agreement "
        f"{res['agreement']:.4f} ({res['mismatches']} mismatches, margins {res['mismatch_margins']}); logits |d| <= 2e-2 + 2e-2|ref| on "
        f"{cover:.4%} of {d.numel()} sampled logits (max |d| {res['logit_max_abs']:.4f}, rms {res['logit_rms']:.4f}, top-1 max |d| "
        f"{res['top1_max_abs']:.4f}); oracle median top-1/top-2 margin {res['median_margin']:.4f}")
    log(f"[calibration: the fp32 oracle's OWN bf16-autocast run vs its fp32 run, {int(cal[4])} positions] agreement {cal[0]:.4f}, mean |dlogit| "
        f"{cal[1]:.4f}, within tolerance {cal[2]:.4%}, median margin {cal[3]:.4f}   (survey, reference code bf16 vs fp32, 556 positions: "
        f"91.4 % agreement, 79.9 % of logits within 2e-2, BASELINE.md §5)")
    # the router's own (free) vote over the 64 observations equals the oracle's
    with torch.no_grad():
        eng.zoedepth(px.to(device))
    res["router_head"], res["router_head_oracle"] = eng.last_router_head, int(np.argmax(g["domain_logits"].sum(0)))
    log(f"[router] device-side vote over the batch: head {res['router_head']}, oracle: head {res['router_head_oracle']}")
    return res


def tokenizer_parity(device="cuda:0", log=print):
    from fakes import FakeTokenizer
    from oracle import tokenizer_ref as T
    from spatialvla_b200 import SpatialActionTokenizer
    out = {}
    nb = {"translation": {"theta_bins": 16, "phi_bins": 32, "r_bins": 8}, "rotation": {"roll_bins": 16, "pitch_bins": 16, "yaw_bins": 16},
          "gripper": 2}
    for name in ("gauss", "uniform"):
        g = np.load(os.path.join(GOLD, f"tokenizer_{name}.npz"))
        pol = {"translation": {k: g[f"edge_{k}"].tolist() for k in ("theta_bins", "phi_bins", "r_bins")},
               "rotation": {k: g[f"edge_{k}"].tolist() for k in ("roll_bins", "pitch_bins", "yaw_bins")}}
        tk = SpatialActionTokenizer(FakeTokenizer(int(g["begin"])), nb, bin_policy=pol)
        begin = tk.action_token_begin_idx
        ids = (tk.encode_ids(torch.from_numpy(g["actions"]).to(device)) - begin).cpu().numpy()
        bad = int((ids != g["local_ids"]).any(1).sum())
        dec = tk.decode_ids(torch.from_numpy(g["decode_ids"]).to(device)).cpu().numpy()
        ref = g["decode_actions"]
        ulp = np.abs(dec - ref) / np.maximum(np.spacing(np.abs(ref)), 1e-300)
        rng = np.random.default_rng(0)
        acts = rng.uniform(-1, 1, size=(1_000_000, 7))
        acts[:, 6] = rng.integers(0, 2, size=acts.shape[0])
        want = T.encode(acts, pol, nb)
        got = (tk.encode_ids(torch.from_numpy(acts).to(device)) - begin).cpu().numpy()
        bad1m = int((got != want).any(1).sum())
        d1m = tk.decode_ids(torch.from_numpy(want + begin).to(device)).cpu().numpy()
        r1m = T.decode(want, pol, nb)
        ulp1m = np.abs(d1m - r1m) / np.maximum(np.spacing(np.abs(r1m)), 1e-300)
        out[name] = {"golden_rows": int(ids.shape[0]), "golden_mismatching_rows": bad, "golden_decode_max_ulp_xyz": float(ulp[:, :3].max()),
                     "golden_decode_max_ulp_rot_grip": float(ulp[:, 3:].max()), "random_1M_mismatching_rows": bad1m,
                     "random_1M_decode_max_ulp": float(ulp1m.max())}
        log(f"[tokenizer, {name} grid] golden (live reference, {ids.shape[0]} actions incl. on-edge rows): {bad} mismatching rows; decode vs "
            f"golden max {ulp[:, :3].max():.0f} ulp (x, y, z) / {ulp[:, 3:].max():.0f} ulp (rotation, gripper); 1 M random actions vs numpy "
            f"on this host: {bad1m} mismatching rows, decode max {ulp1m.max():.0f} ulp")
    return out


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    tokenizer_parity()
    full_size_parity()
