"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (spatialvla_b200/).

Compatibility shim that lets the UNMODIFIED reference (`/root/reference/model`, pinned to
transformers==4.47.0) import and run its `forward()` under the transformers 5.5.0 that this image ships.
It only exists in the build container (the GPU box has no /root/reference); it is used
  * by `oracle/gen_golden.py` to mint the golden vectors committed under `tests/golden/`,
  * by `tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent) to pin the
    restatement in `oracle/model_ref.py` / `oracle/tokenizer_ref.py` to the real reference.

Shims (all monkey-patches in *this* process; the reference tree is never edited) -- SURVEY.md §8(c):
  1. transformers.cache_utils.HybridCache          (model/modeling_gemma2.py:24, modeling_spatialvla.py:25)
  2. transformers.modeling_utils.PretrainedConfig  (model/modeling_spatialvla.py:27)
  3. processing_utils._validate_images_text_input_order (model/processing_spatialvla.py:21)
  4. paligemma.processing_paligemma.make_batched_images (model/processing_spatialvla.py:25)
  5. Gemma2ForCausalLM._tied_weights_keys as dict  (model/modeling_gemma2.py:888)
Config patches: text_config.rope_theta, text_config._attn_implementation="eager", config.pad_token_id.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("SVLA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    sys.dont_write_bytecode = True  # /root/reference is read-only
    import transformers
    import transformers.cache_utils as cu
    import transformers.modeling_utils as mu
    import transformers.processing_utils as pu
    import transformers.models.paligemma.processing_paligemma as pp

    if not hasattr(cu, "HybridCache"):
        class HybridCache(cu.DynamicCache):
            pass
        cu.HybridCache = HybridCache
    if not hasattr(mu, "PretrainedConfig"):
        mu.PretrainedConfig = transformers.PretrainedConfig
    if not hasattr(pu, "_validate_images_text_input_order"):
        pu._validate_images_text_input_order = lambda images, text: (images, text)
    if not hasattr(pp, "make_batched_images"):
        def make_batched_images(images):
            if isinstance(images, (list, tuple)) and images and isinstance(images[0], (list, tuple)):
                return [img for sub in images for img in sub]
            if isinstance(images, (list, tuple)):
                return list(images)
            return [images]
        pp.make_batched_images = make_batched_images
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def import_reference():
    """Returns the reference's `model` package modules (configuration, modeling, tokenizer)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_shims()
    import importlib
    cfg_mod = importlib.import_module("model.configuration_spatialvla")
    tok_mod = importlib.import_module("model.action_tokenizer")
    g2 = importlib.import_module("model.modeling_gemma2")
    g2.Gemma2ForCausalLM._tied_weights_keys = {"lm_head.weight": "model.embed_tokens.weight"}
    mdl = importlib.import_module("model.modeling_spatialvla")
    return cfg_mod, tok_mod, g2, mdl


def build_reference_model(cfg_dict: dict):
    """Instantiate the reference SpatialVLAForConditionalGeneration from a plain config dict
    (see oracle/configs.py) with the config attribute patches of SURVEY.md §8(c)."""
    import copy
    import torch
    cfg_mod, _, _, mdl = import_reference()
    d = copy.deepcopy(cfg_dict)
    d["text_config"]["tie_word_embeddings"] = False
    cfg = cfg_mod.SpatialVLAConfig(**d, tie_word_embeddings=False)
    cfg.text_config.rope_theta = 10000.0
    cfg.text_config._attn_implementation = "eager"
    cfg._attn_implementation = "eager"
    cfg.pad_token_id = 0
    with torch.no_grad():
        model = mdl.SpatialVLAForConditionalGeneration(cfg).eval()
    return model


def reference_greedy(model, input_ids, pixel_values, intrinsic, n_new, act_lo, act_hi, forced_tokens=None):
    """Greedy loop around the reference forward() reproducing HF-4.47 generate's masks:
    bidirectional prefill, causal decode, 1-indexed positions, argmax restricted to [act_lo, act_hi).
    Returns (tokens [B,n_new] int64, logits [B,n_new,act_hi-act_lo] fp32: the action-slice logits the
    argmax was taken on: position 0 = last prompt position)."""
    import torch
    from transformers.cache_utils import DynamicCache
    B, P = input_ids.shape
    cache = DynamicCache(config=model.config.text_config)
    toks, logs = [], []
    with torch.no_grad():
        out = model(input_ids=input_ids, pixel_values=pixel_values, intrinsic=intrinsic,
                    attention_mask=torch.zeros(B, 1, P, P), past_key_values=cache, use_cache=True,
                    cache_position=torch.arange(P))
        for t in range(n_new):
            sl = out.logits[:, -1, act_lo:act_hi].float()
            logs.append(sl)
            nxt = sl.argmax(-1, keepdim=True) + act_lo
            toks.append(nxt)
            if t == n_new - 1:
                break
            feed = nxt if forced_tokens is None else forced_tokens[:, t:t + 1]
            L = P + t + 1
            out = model(input_ids=feed, attention_mask=torch.zeros(B, 1, 1, L), past_key_values=cache,
                        use_cache=True, cache_position=torch.tensor([L - 1]))
    return torch.cat(toks, 1), torch.stack(logs, 1)


def reference_greedy_padded(model, input_ids, attention_mask, pixel_values, intrinsic, n_new, act_lo, act_hi):
    """Left-padded batch through the reference forward(): the 2-D attention_mask goes to the reference's own
    `_update_causal_mask` (model/modeling_spatialvla.py:258-306: bidirectional prompt, padded key columns -> min), and the
    position ids are the ones HF-4.47 generate derives from the mask (model/modeling_gemma2.py:1042-1051 cumsum - 1, pads -> 1)
    plus the +1 of model/modeling_spatialvla.py:473-474.  Returns (tokens [B,n_new], action-slice logits [B,n_new,n_act])."""
    import torch
    from transformers.cache_utils import DynamicCache
    B, P = input_ids.shape
    cache = DynamicCache(config=model.config.text_config)
    am = attention_mask.to(torch.int64)
    toks, logs = [], []
    with torch.no_grad():
        pos = (am.cumsum(-1) - 1).masked_fill(am == 0, 1) + 1
        out = model(input_ids=input_ids, pixel_values=pixel_values, intrinsic=intrinsic, attention_mask=am, position_ids=pos,
                    past_key_values=cache, use_cache=True, cache_position=torch.arange(P))
        for t in range(n_new):
            sl = out.logits[:, -1, act_lo:act_hi].float()
            logs.append(sl)
            nxt = sl.argmax(-1, keepdim=True) + act_lo
            toks.append(nxt)
            if t == n_new - 1:
                break
            am = torch.cat([am, torch.ones(B, 1, dtype=torch.int64)], 1)
            L = P + t + 1
            pos = ((am.cumsum(-1) - 1).masked_fill(am == 0, 1) + 1)[:, -1:]
            out = model(input_ids=nxt, attention_mask=am, position_ids=pos, past_key_values=cache, use_cache=True,
                        cache_position=torch.tensor([L - 1]))
    return torch.cat(toks, 1), torch.stack(logs, 1)
