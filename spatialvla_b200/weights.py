"""HF state_dict key layout of SpatialVLAForConditionalGeneration and a deterministic synthetic
weight generator used by tests / bench (there is no network for checkpoints).

`state_dict_spec(cfg)` lists every persistent key with its shape in the reference's own layout
(reference: model/modeling_spatialvla.py:163-192, model/modeling_gemma2.py:351-354,444-448 plus the HF
Siglip / ZoeDepth / BEiT modules it instantiates); it is pinned against the live reference state_dict in
tests/test_oracle_vs_reference.py.  `synth_state_dict` fills it from per-key seeds so that any subset of keys
can be generated independently, bit-identically on every machine with the same torch build.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch


def _zoe(cfg):
    z = cfg["vision_zoe_config"]
    b = z["backbone_config"]
    return z, b


def state_dict_spec(cfg: dict) -> "OrderedDict[str, tuple]":
    v, t = cfg["vision_config"], cfg["text_config"]
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    D, I, P = v["hidden_size"], v["intermediate_size"], v["patch_size"]
    npos = (v["image_size"] // P) ** 2
    p = "vision_tower.vision_model."
    spec[p + "embeddings.patch_embedding.weight"] = (D, 3, P, P)
    spec[p + "embeddings.patch_embedding.bias"] = (D,)
    spec[p + "embeddings.position_embedding.weight"] = (npos, D)
    for i in range(v["num_hidden_layers"]):
        q = f"{p}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            spec[q + ln + ".weight"] = (D,)
            spec[q + ln + ".bias"] = (D,)
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            spec[q + f"self_attn.{nm}.weight"] = (D, D)
            spec[q + f"self_attn.{nm}.bias"] = (D,)
        spec[q + "mlp.fc1.weight"] = (I, D)
        spec[q + "mlp.fc1.bias"] = (I,)
        spec[q + "mlp.fc2.weight"] = (D, I)
        spec[q + "mlp.fc2.bias"] = (D,)
    spec[p + "post_layernorm.weight"] = (D,)
    spec[p + "post_layernorm.bias"] = (D,)

    H = t["hidden_size"]
    spec["multi_modal_projector.linear.weight"] = (v["projection_dim"], D)
    spec["multi_modal_projector.linear.bias"] = (v["projection_dim"],)

    V, FF = t["vocab_size"], t["intermediate_size"]
    nh, nkv, hd = t["num_attention_heads"], t["num_key_value_heads"], t["head_dim"]
    p = "language_model.model."
    spec[p + "embed_tokens.weight"] = (V, H)
    for i in range(t["num_hidden_layers"]):
        q = f"{p}layers.{i}."
        spec[q + "self_attn.q_proj.weight"] = (nh * hd, H)
        spec[q + "self_attn.k_proj.weight"] = (nkv * hd, H)
        spec[q + "self_attn.v_proj.weight"] = (nkv * hd, H)
        spec[q + "self_attn.o_proj.weight"] = (H, nh * hd)
        spec[q + "mlp.gate_proj.weight"] = (FF, H)
        spec[q + "mlp.up_proj.weight"] = (FF, H)
        spec[q + "mlp.down_proj.weight"] = (H, FF)
        for ln in ("input_layernorm", "post_attention_layernorm", "pre_feedforward_layernorm",
                   "post_feedforward_layernorm"):
            spec[q + ln + ".weight"] = (H,)
    spec[p + "norm.weight"] = (H,)
    spec["language_model.lm_head.weight"] = (V, H)

    if cfg.get("use_vision_zoe", True):
        z, b = _zoe(cfg)
        C, BI, BP = b["hidden_size"], b["intermediate_size"], b["patch_size"]
        bh = b["num_attention_heads"]
        win = b["image_size"] // BP
        nrel = (2 * win - 1) ** 2 + 3
        p = "vision_zoe_model.backbone."
        spec[p + "embeddings.cls_token"] = (1, 1, C)
        spec[p + "embeddings.patch_embeddings.projection.weight"] = (C, 3, BP, BP)
        spec[p + "embeddings.patch_embeddings.projection.bias"] = (C,)
        for i in range(b["num_hidden_layers"]):
            q = f"{p}encoder.layer.{i}."
            spec[q + "lambda_1"] = (C,)
            spec[q + "lambda_2"] = (C,)
            spec[q + "attention.attention.query.weight"] = (C, C)
            spec[q + "attention.attention.query.bias"] = (C,)
            spec[q + "attention.attention.key.weight"] = (C, C)
            spec[q + "attention.attention.value.weight"] = (C, C)
            spec[q + "attention.attention.value.bias"] = (C,)
            spec[q + "attention.attention.relative_position_bias.relative_position_bias_table"] = (nrel, bh)
            spec[q + "attention.output.dense.weight"] = (C, C)
            spec[q + "attention.output.dense.bias"] = (C,)
            spec[q + "intermediate.dense.weight"] = (BI, C)
            spec[q + "intermediate.dense.bias"] = (BI,)
            spec[q + "output.dense.weight"] = (C, BI)
            spec[q + "output.dense.bias"] = (C,)
            for ln in ("layernorm_before", "layernorm_after"):
                spec[q + ln + ".weight"] = (C,)
                spec[q + ln + ".bias"] = (C,)
        nk = z["neck_hidden_sizes"]
        F = z["fusion_hidden_size"]
        p = "vision_zoe_model.neck."
        for s, (ch, fac) in enumerate(zip(nk, z["reassemble_factors"])):
            q = f"{p}reassemble_stage.layers.{s}."
            spec[q + "projection.weight"] = (ch, C, 1, 1)
            spec[q + "projection.bias"] = (ch,)
            if fac > 1:
                spec[q + "resize.weight"] = (ch, ch, int(fac), int(fac))
                spec[q + "resize.bias"] = (ch,)
            elif fac < 1:
                spec[q + "resize.weight"] = (ch, ch, 3, 3)
                spec[q + "resize.bias"] = (ch,)
        for s in range(len(nk)):
            spec[f"{p}reassemble_stage.readout_projects.{s}.0.weight"] = (C, 2 * C)
            spec[f"{p}reassemble_stage.readout_projects.{s}.0.bias"] = (C,)
        for s, ch in enumerate(nk):
            spec[f"{p}convs.{s}.weight"] = (F, ch, 3, 3)
        for s in range(len(nk)):
            q = f"{p}fusion_stage.layers.{s}."
            spec[q + "projection.weight"] = (F, F, 1, 1)
            spec[q + "projection.bias"] = (F,)
            for r in ("residual_layer1", "residual_layer2"):
                for c in ("convolution1", "convolution2"):
                    spec[q + f"{r}.{c}.weight"] = (F, F, 3, 3)
                    spec[q + f"{r}.{c}.bias"] = (F,)
        R = z["num_relative_features"]
        p = "vision_zoe_model.relative_head."
        spec[p + "conv1.weight"] = (F // 2, F, 3, 3)
        spec[p + "conv1.bias"] = (F // 2,)
        spec[p + "conv2.weight"] = (R, F // 2, 3, 3)
        spec[p + "conv2.bias"] = (R,)
        spec[p + "conv3.weight"] = (1, R, 1, 1)
        spec[p + "conv3.bias"] = (1,)
        BN, E = z["bottleneck_features"], z["bin_embedding_dim"]
        TH, TI = z["patch_transformer_hidden_size"], z["patch_transformer_intermediate_size"]
        p = "vision_zoe_model.metric_head."
        spec[p + "conv2.weight"] = (BN, BN, 1, 1)
        spec[p + "conv2.bias"] = (BN,)
        for i in range(z["num_patch_transformer_layers"]):
            q = f"{p}patch_transformer.transformer_encoder.{i}."
            for nm in ("query", "key", "value", "out_proj"):
                spec[q + f"self_attn.{nm}.weight"] = (TH, TH)
                spec[q + f"self_attn.{nm}.bias"] = (TH,)
            spec[q + "linear1.weight"] = (TI, TH)
            spec[q + "linear1.bias"] = (TI,)
            spec[q + "linear2.weight"] = (TH, TI)
            spec[q + "linear2.bias"] = (TH,)
            for ln in ("norm1", "norm2"):
                spec[q + ln + ".weight"] = (TH,)
                spec[q + ln + ".bias"] = (TH,)
        spec[p + "patch_transformer.embedding_convPxP.weight"] = (TH, BN, 1, 1)
        spec[p + "patch_transformer.embedding_convPxP.bias"] = (TH,)
        spec[p + "mlp_classifier.linear1.weight"] = (128, 128)
        spec[p + "mlp_classifier.linear1.bias"] = (128,)
        spec[p + "mlp_classifier.linear2.weight"] = (2, 128)
        spec[p + "mlp_classifier.linear2.bias"] = (2,)
        names = [c["name"] for c in z["bin_configurations"]]
        for c in z["bin_configurations"]:
            q = f"{p}seed_bin_regressors.{c['name']}."
            spec[q + "conv1.weight"] = (E // 2, BN, 1, 1)
            spec[q + "conv1.bias"] = (E // 2,)
            spec[q + "conv2.weight"] = (c["n_bins"], E // 2, 1, 1)
            spec[q + "conv2.bias"] = (c["n_bins"],)
        spec[p + "seed_projector.conv1.weight"] = (E // 2, BN, 1, 1)
        spec[p + "seed_projector.conv1.bias"] = (E // 2,)
        spec[p + "seed_projector.conv2.weight"] = (E, E // 2, 1, 1)
        spec[p + "seed_projector.conv2.bias"] = (E,)
        for s in range(4):
            q = f"{p}projectors.{s}."
            spec[q + "conv1.weight"] = (E // 2, F, 1, 1)
            spec[q + "conv1.bias"] = (E // 2,)
            spec[q + "conv2.weight"] = (E, E // 2, 1, 1)
            spec[q + "conv2.bias"] = (E,)
        for nmm in names:
            for s in range(len(z["num_attractors"])):
                q = f"{p}attractors.{nmm}.{s}."
                spec[q + "conv1.weight"] = (E, E, 1, 1)
                spec[q + "conv1.bias"] = (E,)
                # HF quirk: every attractor layer is built with the default n_attractors=16
                # (HF zoedepth/modeling_zoedepth.py: Attractor(config, n_bins=n_attractors[i], ...))
                spec[q + "conv2.weight"] = (16, E, 1, 1)
                spec[q + "conv2.bias"] = (16,)
        for c in z["bin_configurations"]:
            q = f"{p}conditional_log_binomial.{c['name']}.mlp."
            bott = (R + E) // 4
            spec[q + "0.weight"] = (bott, R + E, 1, 1)
            spec[q + "0.bias"] = (bott,)
            spec[q + "2.weight"] = (4, bott, 1, 1)
            spec[q + "2.bias"] = (4,)
        nin = cfg["ego3d_patch_reso"] ** 2 * 3 * (2 * cfg["n_freqs"] + 1)
        p = "position_embedding_3d.position_embedding_head."
        spec[p + "0.weight"] = (D, nin)
        spec[p + "0.bias"] = (D,)
        spec[p + "1.weight"] = (D,)
        spec[p + "1.bias"] = (D,)
        spec[p + "3.weight"] = (D, D)
        spec[p + "3.bias"] = (D,)
    if cfg.get("use_spatial_token", True):
        spec["spatial_embed_tokens.weight"] = (cfg["spatial_token_num"], H)
    return spec


def _is_norm_weight(key: str) -> bool:
    k = key.rsplit(".", 2)
    leaf = ".".join(k[-2:])
    names = ("layer_norm1.weight", "layer_norm2.weight", "post_layernorm.weight", "layernorm_before.weight",
             "layernorm_after.weight", "norm1.weight", "norm2.weight", "position_embedding_head.1.weight")
    return any(key.endswith(n) for n in names) or leaf in names


def synth_tensor(key: str, shape: tuple, seed: int = 0, device="cpu", on_device_rng: bool = False) -> torch.Tensor:
    """One synthetic fp32 parameter. Fan-in scaled normals keep activations O(1) through every sub-model so
    that parity tests are sensitive (a sigma=0.02 init makes ZoeDepth's output nearly constant).
    `on_device_rng` draws with the device generator (fast for the 4 B-parameter bench model; values then differ
    from the CPU stream, so parity tests never use it)."""
    gdev = device if on_device_rng else "cpu"
    g = torch.Generator(device=gdev)
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    n = 1
    for s in shape:
        n *= s

    def randn(std=1.0):
        return torch.randn(n, generator=g, dtype=torch.float32, device=gdev).reshape(shape) * std

    if key.startswith("language_model.model.layers") and key.endswith("layernorm.weight") \
            or key == "language_model.model.norm.weight":
        out = randn(0.1)                       # Gemma RMSNorm scale is (1 + w)
    elif _is_norm_weight(key):
        out = 1.0 + randn(0.1)
    elif key.endswith("lambda_1") or key.endswith("lambda_2"):
        out = 0.1 * (1.0 + randn(0.2))         # BEiT layer-scale, init value 0.1
    elif key.endswith("relative_position_bias_table"):
        out = randn(0.5)
    elif key.endswith("cls_token"):
        out = randn(0.5)
    elif key.endswith("position_embedding.weight"):
        out = randn(0.1)
    elif key.endswith("embed_tokens.weight"):
        out = randn(0.02)
    elif "conditional_log_binomial" in key and key.endswith("mlp.2.bias"):
        # low temperature -> peaked bin distribution -> depth varies strongly per pixel (sensitive parity tests)
        out = torch.tensor([0.0, 0.0, -4.0, 3.0], device=gdev) + randn(0.05)
    elif key.endswith(".bias"):
        out = randn(0.05)
    elif len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        if "resize.weight" in key and len(shape) == 4 and shape[2] in (2, 4):
            fan_in = shape[0]                  # ConvTranspose2d weight is (Cin, Cout, k, k); k==stride
        out = randn(1.0 / math.sqrt(fan_in))
    else:
        out = randn(0.05)
    return out.to(device)


def synth_state_dict(cfg: dict, seed: int = 0, prefix_filter=None, device="cpu", bf16_round: bool = True,
                     on_device_rng: bool = False, dtype=torch.float32):
    """All (or a prefix-filtered subset of) parameters. With `bf16_round` every GEMM operand matrix is rounded
    to bf16 and stored back as fp32, so the fp32 oracle and the bf16 B200 path share identical weights
    (weight quantisation is common-mode; SURVEY.md §7 'bf16-vs-fp32 flips')."""
    sd = OrderedDict()
    for key, shape in state_dict_spec(cfg).items():
        if prefix_filter is not None and not any(key.startswith(p) for p in prefix_filter):
            continue
        t = synth_tensor(key, shape, seed, device=device, on_device_rng=on_device_rng)
        if bf16_round and t.dim() >= 2:
            t = t.to(torch.bfloat16)
            t = t if dtype == torch.bfloat16 else t.to(torch.float32)
        sd[key] = t.to(device)
    return sd
