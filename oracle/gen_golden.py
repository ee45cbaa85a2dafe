"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/*.npz from the LIVE reference (/root/reference imported
through oracle/compat.py).  Run in the build container:  python -m oracle.gen_golden
The fixtures pin (a) oracle/tokenizer_ref.py + the CUDA tokenizer, (b) oracle/model_ref.py + the CUDA path
on the TINY config with `synth_state_dict(TINY, seed=0)` weights (regenerated bit-identically at test time).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import compat  # noqa: E402
from spatialvla_b200.configs import get_config_dict, default_intrinsic_224  # noqa: E402
from spatialvla_b200.weights import synth_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
BEGIN = 257153


class FakeHFTokenizer:
    """Minimal stand-in for the Gemma tokenizer surface the reference tokenizer touches
    (add_tokens / convert_tokens_to_ids / vocab_size / __len__): base vocab 257 153."""

    def __init__(self, base=BEGIN):
        self.base = base
        self.added = {}

    @property
    def vocab_size(self):
        return self.base

    def __len__(self):
        return self.base + len(self.added)

    def add_tokens(self, toks, special_tokens=False):
        n = 0
        for t in toks:
            if t not in self.added:
                self.added[t] = self.base + len(self.added)
                n += 1
        return n

    def convert_tokens_to_ids(self, tok):
        if isinstance(tok, (list, tuple, np.ndarray)):
            return [self.added[str(t)] for t in tok]
        return self.added[str(tok)]


def edge_case_actions(pol):
    rows = [np.zeros(7), np.ones(7), -np.ones(7), np.array([1, -1, 1, -1, 1, -1, 0.5]),
            np.array([0, 0, 1, 0, 0, 0, 0.49999]), np.array([0, 0, -1, 0.3, 0.3, 0.3, 1.0]),
            np.array([1e-300, 0, 0, 0, 0, 0, 0]), np.array([2.5, -3.0, 0.1, 1.5, -1.5, 0.0, 7.0])]
    for key in ("roll_bins", "pitch_bins", "yaw_bins"):
        for e in pol["rotation"][key]:
            for d in (0.0, 1e-12, -1e-12):
                r = np.zeros(7)
                r[3:6] = e + d
                rows.append(r)
    # points whose radius / polar angle sit exactly on translation edges
    for e in pol["translation"]["r_bins"]:
        rows.append(np.array([min(e, 1.0), 0, 0, 0, 0, 0, 0]))
        rows.append(np.array([0, 0, min(e, 1.0), 0, 0, 0, 0]))
    for e in pol["translation"]["theta_bins"]:
        rows.append(np.array([0.5 * np.sin(e), 0, 0.5 * np.cos(e), 0, 0, 0, 0]))
    for e in pol["translation"]["phi_bins"]:
        rows.append(np.array([0.5 * np.cos(e), 0.5 * np.sin(e), 0.1, 0, 0, 0, 1]))
    return np.stack(rows).astype(np.float64)


def gen_tokenizer():
    _, tok_mod, _, _ = compat.import_reference()
    action_cfg = json.load(open(os.path.join(compat.REFERENCE_ROOT, "scripts", "action_config.json")))
    gs = json.load(open(os.path.join(compat.REFERENCE_ROOT, "scripts", "gs_spatialvla_plus.json")))
    for name, gs_params, min_sigma in (("gauss", gs, 0.5), ("uniform", None, 0.0)):
        tk = tok_mod.SpatialActionTokenizer(FakeHFTokenizer(), num_bins=action_cfg["num_bins"], gs_params=gs_params,
                                            use_spherical=True, min_sigma=min_sigma)
        pol = tk.bin_policy
        assert tk.action_token_begin_idx == BEGIN and tk.vocab_size == 8194
        rng = np.random.default_rng(0)
        acts = rng.uniform(-1, 1, size=(20000, 7))
        acts[:, 6] = rng.integers(0, 2, size=20000)
        acts = np.concatenate([acts, edge_case_actions(pol)], 0)
        toks = tk(acts)
        ids = np.vectorize(lambda s: int(s[7:12]))(toks).astype(np.int32)       # '<ACTION%05d>' -> local id
        all_ids = np.stack([np.arange(4096) , np.arange(4096) + 4096,
                            8192 + (np.arange(4096) % 2)], 1) + BEGIN
        dec_all = tk.decode_token_ids_to_actions(all_ids)
        oob = np.array([[BEGIN - 5, BEGIN + 100, BEGIN + 9000], [BEGIN + 5000, BEGIN + 10, 0],
                        [0, 0, 0], [BEGIN + 8193, BEGIN + 8193, BEGIN + 8193]])
        dec_oob = tk.decode_token_ids_to_actions(oob)
        edges = {f"edge_{k}": np.asarray(v, dtype=np.float64) for bt in pol.values() for k, v in bt.items()}
        np.savez_compressed(os.path.join(GOLD, f"tokenizer_{name}.npz"), actions=acts, local_ids=ids,
                            decode_ids=all_ids.astype(np.int64), decode_actions=dec_all, oob_ids=oob.astype(np.int64),
                            oob_actions=dec_oob, begin=np.int64(BEGIN), min_sigma=np.float64(min_sigma), **edges)
        print(f"tokenizer_{name}: {acts.shape[0]} actions, ids range {ids.min()}..{ids.max()}")


def tiny_inputs(B=2, T=6, seed=0):
    cfg = get_config_dict("tiny")
    g = torch.Generator().manual_seed(seed)
    px_u8 = torch.randint(0, 256, (B, 3, 224, 224), generator=g, dtype=torch.uint8)
    # smooth the noise a little so that resampling paths see structure, keep exact u8 storage
    px_u8 = (torch.nn.functional.avg_pool2d(px_u8.float(), 5, 1, 2)).round().clamp(0, 255).to(torch.uint8)
    ids = torch.cat([torch.full((B, 256), cfg["image_token_index"]), torch.full((B, 1), 2),
                     torch.randint(3, 1000, (B, T), generator=g), torch.full((B, 1), 108)], 1)
    K = torch.tensor(default_intrinsic_224(), dtype=torch.float32)
    return cfg, px_u8, ids, K


def gen_model(n_new=6):
    cfg, px_u8, ids, K = tiny_inputs()
    px = px_u8.float() / 255.0
    model = compat.build_reference_model(cfg)
    sd = synth_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    lo, hi = cfg["action_token_begin_idx"], cfg["action_token_begin_idx"] + cfg["spatial_token_num"]
    from oracle.model_ref import process_zoe
    with torch.no_grad():
        sig = model.vision_tower((px - 0.5) / 0.5).last_hidden_state
        zo = model.vision_zoe_model(pixel_values=process_zoe(px), output_hidden_states=False)
        depth384 = zo.predicted_depth
        depth224 = torch.nn.functional.interpolate(depth384.unsqueeze(1), size=(286, 286), mode="bicubic",
                                                   align_corners=True)[..., 31:-31, 31:-31]
        xyz = model.backproject_patch(K, depth224, patch_size=14, reso=cfg["ego3d_patch_reso"])
        pos3d = model.position_embedding_3d(xyz)
        feats = model.get_image_features(px, K)
    toks, logits = compat.reference_greedy(model, ids, px, K, n_new, lo, hi)
    # teacher-forced second pass on a fixed token stream (tests feed the same stream to both sides)
    forced = toks.clone()
    toks_f, logits_f = compat.reference_greedy(model, ids, px, K, n_new, lo, hi, forced_tokens=forced)
    assert torch.equal(toks, toks_f)
    np.savez_compressed(
        os.path.join(GOLD, "tiny_model.npz"),
        pixel_u8=px_u8.numpy(), input_ids=ids.numpy(), intrinsic=K.numpy(),
        siglip=sig[:, ::4].numpy(), depth384_s4=depth384[:, ::4, ::4].numpy(),
        depth384_mean=depth384.mean((1, 2)).numpy(), depth384_std=depth384.std((1, 2)).numpy(),
        domain_logits=zo.domain_logits.numpy(), xyz=xyz.numpy(), pos3d=pos3d[:, ::4].numpy(),
        image_features=feats[:, ::4].numpy(), tokens=toks.numpy(), logits=logits.numpy().astype(np.float32),
        n_new=np.int64(n_new))
    print("tiny_model: tokens", toks.tolist(), "domain_logits", zo.domain_logits.tolist())


def padded_inputs(seed=1):
    """3 rows with prompts of 6 / 3 / 1 text tokens, LEFT-padded (pad id 0, attention_mask 0) to a common length."""
    cfg, px_u8, _, K = tiny_inputs(B=3, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    rows, masks = [], []
    Tmax = 6
    for T in (6, 3, 1):
        real = torch.cat([torch.full((256,), cfg["image_token_index"]), torch.tensor([2]), torch.randint(3, 1000, (T,), generator=g),
                          torch.tensor([108])])
        pad = Tmax - T
        rows.append(torch.cat([torch.zeros(pad, dtype=torch.int64), real]))
        masks.append(torch.cat([torch.zeros(pad, dtype=torch.int64), torch.ones(real.numel(), dtype=torch.int64)]))
    return cfg, px_u8, torch.stack(rows), torch.stack(masks), K


def gen_model_padded(n_new=6):
    """Left-padded batch through the live reference (masks by its own _update_causal_mask) -> tests/golden/tiny_model_padded.npz"""
    cfg, px_u8, ids, am, K = padded_inputs()
    px = px_u8.float() / 255.0
    model = compat.build_reference_model(cfg)
    model.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
    lo, hi = cfg["action_token_begin_idx"], cfg["action_token_begin_idx"] + cfg["spatial_token_num"]
    toks, logits = compat.reference_greedy_padded(model, ids, am, px, K, n_new, lo, hi)
    # size-independent property the tests reuse: a padded row decodes exactly like the same sample alone, unpadded
    # (ZoeDepth's router is batch-coupled, so the comparison forces the batch's head)
    np.savez_compressed(os.path.join(GOLD, "tiny_model_padded.npz"), pixel_u8=px_u8.numpy(), input_ids=ids.numpy(),
                        attention_mask=am.numpy(), intrinsic=K.numpy(), tokens=toks.numpy(),
                        logits=logits.numpy().astype(np.float32), n_new=np.int64(n_new))
    print("tiny_model_padded: tokens", toks.tolist())


WINDOW = 48


def gen_model_window(n_new=6):
    """Gemma2's sliding-window layers (even layer_idx, model/modeling_gemma2.py:343,441-473) through the live reference: the tiny
    configuration with text_config.sliding_window = 48 << prompt length 264, so the window predicate is active in the bidirectional
    prefill (tril(diagonal=-window) over the full mask) AND in every decode step (the reference's sliding cache keeps the last
    `window` slots) -> tests/golden/tiny_model_window.npz; plus the same for the left-padded batch."""
    cfg, px_u8, ids, K = tiny_inputs()
    cfg["text_config"]["sliding_window"] = WINDOW
    px = px_u8.float() / 255.0
    model = compat.build_reference_model(cfg)
    model.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
    lo, hi = cfg["action_token_begin_idx"], cfg["action_token_begin_idx"] + cfg["spatial_token_num"]
    toks, logits = compat.reference_greedy(model, ids, px, K, n_new, lo, hi)
    B, P = ids.shape
    with torch.no_grad():          # every prefill position (no cache): post-softcap logits at 64 fixed action columns
        full = model(input_ids=ids, pixel_values=px, intrinsic=K, attention_mask=torch.zeros(B, 1, P, P), use_cache=False).logits
    cols = torch.arange(lo, hi, (hi - lo) // 64)[:64]
    _, px_u8p, idsp, amp, Kp = padded_inputs()
    toks_p, logits_p = compat.reference_greedy_padded(model, idsp, amp, px_u8p.float() / 255.0, Kp, n_new, lo, hi)
    np.savez_compressed(os.path.join(GOLD, "tiny_model_window.npz"), window=np.int64(WINDOW), tokens=toks.numpy(),
                        logits=logits.numpy().astype(np.float32), prefill_cols=cols.numpy(),
                        prefill_logits=full[:, :, cols].float().numpy(), tokens_padded=toks_p.numpy(),
                        logits_padded=logits_p.numpy().astype(np.float32), n_new=np.int64(n_new))
    print("tiny_model_window: tokens", toks.tolist(), "padded", toks_p.tolist())


def train_inputs(B=2, T=6, n_act=6, seed=2):
    """Training-shaped samples (train/monkey_patch.py:21-75, data/dataset.py:145-153): prefix = image tokens + BOS + text +
    newline (token type 0, labels -100), suffix = action tokens + EOS (token type 1, labels = ids)."""
    cfg, px_u8, ids, K = tiny_inputs(B=B, T=T, seed=seed)
    g = torch.Generator().manual_seed(seed + 200)
    lo = cfg["action_token_begin_idx"]
    suffix = torch.cat([torch.randint(lo, lo + cfg["spatial_token_num"], (B, n_act), generator=g),
                        torch.full((B, 1), cfg["eos_token_id"])], 1)
    full = torch.cat([ids, suffix], 1)
    P = ids.shape[1]
    tt = torch.cat([torch.zeros(B, P, dtype=torch.int64), torch.ones(B, suffix.shape[1], dtype=torch.int64)], 1)
    labels = torch.where(tt == 1, full, torch.full_like(full, -100))
    return cfg, px_u8, full, tt, labels, K


def gen_model_train():
    """forward(labels=...) of the live reference under its three masks -> tests/golden/tiny_model_train.npz:
    prefix-LM (token_type_ids + 2-D attention_mask), triangular (token_type_ids, no attention_mask), bidirectional (labels only)."""
    cfg, px_u8, ids, tt, labels, K = train_inputs()
    px = px_u8.float() / 255.0
    model = compat.build_reference_model(cfg)
    model.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
    B, L = ids.shape
    out = {}
    bi, ti = torch.nonzero(labels[:, 1:] != -100, as_tuple=True)
    with torch.no_grad():
        for name, kw in (("prefix_lm", dict(token_type_ids=tt, attention_mask=torch.ones(B, L, dtype=torch.int64))),
                         ("causal", dict(token_type_ids=tt)), ("bidirectional", dict())):
            o = model(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels, use_cache=False, **kw)
            out["loss_" + name] = o.loss.float().numpy()
            out["logits_" + name] = o.logits.float()[bi, ti].numpy().astype(np.float32)      # labelled rows, full vocabulary
            print(f"tiny_model_train[{name}]: loss {float(o.loss):.6f}")
    np.savez_compressed(os.path.join(GOLD, "tiny_model_train.npz"), pixel_u8=px_u8.numpy(), input_ids=ids.numpy(),
                        token_type_ids=tt.numpy(), labels=labels.numpy(), intrinsic=K.numpy(), **out)


ADAPT_BINS = {"translation": {"theta_bins": 4, "phi_bins": 5, "r_bins": 3}, "rotation": {"roll_bins": 3, "pitch_bins": 4, "yaw_bins": 2},
              "gripper": 2}
ADAPT_GS0 = {k: {"mu": 0.05 * i - 0.1, "sigma": 0.3 + 0.05 * i} for i, k in enumerate(("theta", "phi", "r", "roll", "pitch", "yaw"))}
ADAPT_GS1 = {k: {"mu": -0.04 * i + 0.1, "sigma": 0.45 - 0.03 * i} for i, k in enumerate(("theta", "phi", "r", "roll", "pitch", "yaw"))}


def gen_adaption():
    """spatial_embedding_adaption of the live reference (model/action_tokenizer.py:372-430) on a small grid
    -> tests/golden/embedding_adaption.npz (old embeddings, re-sampled embeddings, new bin edges)."""
    _, tok_mod, _, _ = compat.import_reference()
    n_tok = 4 * 5 * 3 + 3 * 4 * 2 + 2
    w = torch.randn(n_tok, 6, generator=torch.Generator().manual_seed(0))
    tk = tok_mod.SpatialActionTokenizer(FakeHFTokenizer(1000), num_bins=ADAPT_BINS, gs_params=ADAPT_GS0, min_sigma=0.1)
    emb = torch.nn.Embedding(n_tok, 6)
    emb.weight.data.copy_(w)
    tk.spatial_embedding_adaption(ADAPT_GS1, emb, min_sigma=0.2, adpt_feature=True)
    edges = {f"edge_{k}": np.asarray(tk.bin_policy[bt][k]) for bt in ("translation", "rotation") for k in ADAPT_BINS[bt]}
    np.savez_compressed(os.path.join(GOLD, "embedding_adaption.npz"), before=w.numpy(), after=emb.weight.data.numpy(), **edges)
    print("embedding_adaption: NaN rows", int(torch.isnan(emb.weight.data).any(1).sum()), "of", n_tok)


GRAD_KEYS = ("language_model.model.layers.2.self_attn.q_proj.weight", "language_model.model.layers.0.mlp.gate_proj.weight",
             "language_model.model.layers.1.mlp.down_proj.weight", "language_model.model.layers.2.input_layernorm.weight",
             "multi_modal_projector.linear.weight", "vision_tower.vision_model.encoder.layers.1.mlp.fc2.weight",
             "vision_tower.vision_model.encoder.layers.0.self_attn.out_proj.weight",
             "position_embedding_3d.position_embedding_head.0.weight", "position_embedding_3d.position_embedding_head.3.weight")


def gen_model_train_grads():
    """loss.backward() of the live reference (prefix-LM training mask) -> tests/golden/tiny_model_train_grads.npz: dLoss/dW of a
    LoRA-target sample across the towers (train/spatialvla_finetune.py:262-270), every 3rd row / column to keep the file small.
    Pins oracle/model_ref.loss_and_grads_ref, the checker of next round's backward kernels."""
    cfg, px_u8, ids, tt, labels, K = train_inputs()
    px = px_u8.float() / 255.0
    model = compat.build_reference_model(cfg)
    model.load_state_dict(synth_state_dict(cfg, seed=0), strict=True)
    params = dict(model.named_parameters())
    for p_ in params.values():
        p_.requires_grad_(False)
    for k in GRAD_KEYS:
        params[k].requires_grad_(True)
    B, L = ids.shape
    o = model(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels, use_cache=False, token_type_ids=tt,
              attention_mask=torch.ones(B, L, dtype=torch.int64))
    o.loss.backward()
    out = {"loss": o.loss.detach().float().numpy()}
    for k in GRAD_KEYS:
        gk = params[k].grad.float()
        out["grad:" + k] = (gk[::3, ::3] if gk.dim() == 2 else gk).numpy()
        out["norm:" + k] = gk.norm().numpy()
        print(f"grad {k}: norm {float(gk.norm()):.6e}")
    np.savez_compressed(os.path.join(GOLD, "tiny_model_train_grads.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "adaption":
        gen_adaption()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "grads":
        gen_model_train_grads()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "padded":
        gen_model_padded()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        gen_model_train()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "window":
        gen_model_window()
        sys.exit(0)
    gen_tokenizer()
    gen_model()
    gen_model_padded()
    gen_model_window()
    gen_model_train()
    gen_model_train_grads()
    gen_adaption()
