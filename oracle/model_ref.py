"""TEST INFRASTRUCTURE ONLY -- the CPU oracle. Never imported by spatialvla_b200/ (the product path).

fp32 torch restatement of the reference's `predict_action` arithmetic as plain functions over an HF-keyed
state_dict.  It exists because /root/reference does not travel to the GPU box and because the reference's own
`generate()` path does not run under transformers 5.5.  Each function cites the reference lines
it follows.  PINNING: tests/test_oracle_vs_reference.py runs this file against the live reference (through
oracle/compat.py) in the build container, and tests/golden/*.npz hold outputs of the live reference minted by
oracle/gen_golden.py -- tests/test_oracle_golden.py checks this file against them everywhere.
Third-party arithmetic (not under /root/reference): transformers==4.47.0 Siglip / ZoeDepth / BEiT modules; the
restatement follows the installed transformers 5.5.0 sources (`HF:` citations), which the live-reference
comparison exercises directly.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- helpers
def _lin(x, sd, key, bias=True):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias") if bias else None)


def _ln(x, sd, key, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def _conv(x, sd, key, stride=1, padding=0):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


# ----------------------------------------------------------------------------------------------- SigLIP
def siglip_forward(sd, cfg, pixel_values):
    """HF:models/siglip/modeling_siglip.py:116-187 (embeddings), :252-362 (encoder layer), :586-625.
    pixel_values already normalised to (x-0.5)/0.5 (model/modeling_spatialvla.py:309). -> (B, 256, D)"""
    v = cfg["vision_config"]
    p = "vision_tower.vision_model."
    nh = v["num_attention_heads"]
    eps = v.get("layer_norm_eps", 1e-6)
    x = F.conv2d(pixel_values, sd[p + "embeddings.patch_embedding.weight"], sd[p + "embeddings.patch_embedding.bias"],
                 stride=v["patch_size"])
    x = x.flatten(2).transpose(1, 2) + sd[p + "embeddings.position_embedding.weight"][None]
    B, S, D = x.shape
    hd = D // nh
    for i in range(v["num_hidden_layers"]):
        q = f"{p}encoder.layers.{i}."
        h = _ln(x, sd, q + "layer_norm1", eps)
        qq = _lin(h, sd, q + "self_attn.q_proj").view(B, S, nh, hd).transpose(1, 2)
        kk = _lin(h, sd, q + "self_attn.k_proj").view(B, S, nh, hd).transpose(1, 2)
        vv = _lin(h, sd, q + "self_attn.v_proj").view(B, S, nh, hd).transpose(1, 2)
        att = torch.softmax((qq @ kk.transpose(-1, -2)) * hd ** -0.5, dim=-1)
        ctx = (att @ vv).transpose(1, 2).reshape(B, S, D)
        x = x + _lin(ctx, sd, q + "self_attn.out_proj")
        h = _ln(x, sd, q + "layer_norm2", eps)
        x = x + _lin(gelu_tanh(_lin(h, sd, q + "mlp.fc1")), sd, q + "mlp.fc2")
    return _ln(x, sd, p + "post_layernorm", eps)


# ----------------------------------------------------------------------------------------------- ZoeDepth
def process_zoe(pixel_values):
    """model/modeling_spatialvla.py:99-110: reflect-pad 31, bicubic(align_corners) -> 384x384, (x-.5)/.5"""
    x = F.pad(pixel_values, (31, 31, 31, 31), mode="reflect")
    x = F.interpolate(x, size=(384, 384), mode="bicubic", align_corners=True)
    return (x - 0.5) / 0.5


def beit_rel_pos_bias(table, win):
    """HF:models/beit/modeling_beit.py:511-590. At the native window the bilinear re-interpolation of the
    table is the identity, so bias[h, i, j] = table[index[i, j], h]. -> (nH, win*win+1, win*win+1)"""
    nrel = (2 * win - 1) ** 2 + 3
    coords = torch.stack(torch.meshgrid(torch.arange(win), torch.arange(win), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += win - 1
    rel[:, :, 1] += win - 1
    rel[:, :, 0] *= 2 * win - 1
    idx = torch.zeros((win * win + 1,) * 2, dtype=torch.long)
    idx[1:, 1:] = rel.sum(-1)
    idx[0, 0:] = nrel - 3
    idx[0:, 0] = nrel - 2
    idx[0, 0] = nrel - 1
    return table[idx.view(-1)].view(win * win + 1, win * win + 1, -1).permute(2, 0, 1).contiguous()


def beit_forward(sd, cfg, x384):
    """HF:models/beit/modeling_beit.py:92-222 (embeddings: CLS + patch conv, no abs-pos), :225-306 (attention:
    q,v bias, k no bias, rel-pos bias), :448-508 (pre-LN block with layer scale), :1340-1460 (backbone taps).
    -> list of 4 hidden states (B, 577, C)"""
    z = cfg["vision_zoe_config"]
    b = z["backbone_config"]
    p = "vision_zoe_model.backbone."
    nh, ps = b["num_attention_heads"], b["patch_size"]
    eps = b.get("layer_norm_eps", 1e-12)
    win = b["image_size"] // ps
    x = F.conv2d(x384, sd[p + "embeddings.patch_embeddings.projection.weight"],
                 sd[p + "embeddings.patch_embeddings.projection.bias"], stride=ps)
    x = x.flatten(2).transpose(1, 2)
    B = x.shape[0]
    x = torch.cat([sd[p + "embeddings.cls_token"].expand(B, -1, -1), x], 1)
    S, C = x.shape[1], x.shape[2]
    hd = C // nh
    taps = [int(s.replace("stage", "")) for s in b["out_features"]]
    outs = []
    for i in range(b["num_hidden_layers"]):
        q = f"{p}encoder.layer.{i}."
        h = _ln(x, sd, q + "layernorm_before", eps)
        qq = _lin(h, sd, q + "attention.attention.query").view(B, S, nh, hd).transpose(1, 2)
        kk = _lin(h, sd, q + "attention.attention.key", bias=False).view(B, S, nh, hd).transpose(1, 2)
        vv = _lin(h, sd, q + "attention.attention.value").view(B, S, nh, hd).transpose(1, 2)
        sc = (qq @ kk.transpose(-1, -2)) / math.sqrt(hd)
        sc = sc + beit_rel_pos_bias(
            sd[q + "attention.attention.relative_position_bias.relative_position_bias_table"], win)[None]
        ctx = (torch.softmax(sc, -1) @ vv).transpose(1, 2).reshape(B, S, C)
        x = x + sd[q + "lambda_1"] * _lin(ctx, sd, q + "attention.output.dense")
        h = _ln(x, sd, q + "layernorm_after", eps)
        h = _lin(F.gelu(_lin(h, sd, q + "intermediate.dense")), sd, q + "output.dense")
        x = x + sd[q + "lambda_2"] * h
        if (i + 1) in taps:
            outs.append(x)
    return outs


def _preact_residual(x, sd, key):
    """HF:models/zoedepth/modeling_zoedepth.py:182-238 (no batch norm, bias on)"""
    h = _conv(F.relu(x), sd, key + ".convolution1", padding=1)
    h = _conv(F.relu(h), sd, key + ".convolution2", padding=1)
    return h + x


def zoe_neck(sd, cfg, hidden_states, ph, pw):
    """HF:models/zoedepth/modeling_zoedepth.py:55-149 (reassemble, readout 'project'), :152-175 + :241-275
    (fusion), :278-329 (neck). -> (fused list [4], features[-1])"""
    z = cfg["vision_zoe_config"]
    p = "vision_zoe_model.neck."
    feats = []
    for s, (hs, fac) in enumerate(zip(hidden_states, z["reassemble_factors"])):
        cls, tok = hs[:, 0], hs[:, 1:]
        B, N, C = tok.shape
        cat = torch.cat([tok, cls[:, None].expand_as(tok)], -1)
        h = F.gelu(_lin(cat, sd, f"{p}reassemble_stage.readout_projects.{s}.0"))
        h = h.permute(0, 2, 1).reshape(B, C, ph, pw)
        h = _conv(h, sd, f"{p}reassemble_stage.layers.{s}.projection")
        if fac > 1:
            h = F.conv_transpose2d(h, sd[f"{p}reassemble_stage.layers.{s}.resize.weight"],
                                   sd[f"{p}reassemble_stage.layers.{s}.resize.bias"], stride=int(fac))
        elif fac < 1:
            h = _conv(h, sd, f"{p}reassemble_stage.layers.{s}.resize", stride=int(1 / fac), padding=1)
        feats.append(h)
    feats = [F.conv2d(f, sd[f"{p}convs.{s}.weight"], None, padding=1) for s, f in enumerate(feats)]
    fused_list, fused = [], None
    for li, f in enumerate(feats[::-1]):
        q = f"{p}fusion_stage.layers.{li}."
        if fused is None:
            h = f
        else:
            r = f
            if fused.shape != r.shape:
                r = F.interpolate(r, size=fused.shape[2:], mode="bilinear", align_corners=False)
            h = fused + _preact_residual(r, sd, q + "residual_layer1")
        h = _preact_residual(h, sd, q + "residual_layer2")
        h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
        fused = _conv(h, sd, q + "projection")
        fused_list.append(fused)
    return fused_list, feats[-1]


def zoe_relative_head(sd, fused_last):
    """HF:models/zoedepth/modeling_zoedepth.py:332-373 -> (relative_depth (B,H,W), features (B,32,H,W))"""
    p = "vision_zoe_model.relative_head."
    h = _conv(fused_last, sd, p + "conv1", padding=1)
    h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
    feat = F.relu(_conv(h, sd, p + "conv2", padding=1))
    rel = F.relu(_conv(feat, sd, p + "conv3"))
    return rel.squeeze(1), feat


def _log_binom(n, k, eps=1e-7):
    n = n + eps
    k = k + eps
    return n * torch.log(n) - k * torch.log(k) - (n - k) * torch.log(n - k + eps)


def zoe_metric_head(sd, cfg, outconv, bottleneck, feature_blocks, force_head=None):
    """HF:models/zoedepth/modeling_zoedepth.py:965-1103 (multi-head variant): batch-level router (:1059-1067),
    softplus seed regressor (:494-547), projector (:749-772), un-normed attractors with the inv_attractor
    defaults alpha=300, gamma=2 and mean over 16 attractors (:551-570, :665-746), conditional log-binomial
    (:383-491). -> (depth (B,1,H,W), domain_logits (B,2), chosen head index)"""
    z = cfg["vision_zoe_config"]
    p = "vision_zoe_model.metric_head."
    x = _conv(bottleneck, sd, p + "conv2")
    # patch transformer router
    e = _conv(x, sd, p + "patch_transformer.embedding_convPxP").flatten(2)
    e = F.pad(e, (1, 0)).permute(0, 2, 1)
    B, S, E = e.shape
    pos = torch.arange(0, S, dtype=torch.float32).unsqueeze(1)
    idx = torch.arange(0, E, 2, dtype=torch.float32).unsqueeze(0)
    div = torch.exp(idx * (-torch.log(torch.tensor(10000.0)) / E))
    pe = pos * div
    e = e + torch.cat([torch.sin(pe), torch.cos(pe)], 1)[None]
    nh = z["patch_transformer_num_attention_heads"]
    hd = E // nh
    for i in range(z["num_patch_transformer_layers"]):
        q = f"{p}patch_transformer.transformer_encoder.{i}."
        qq = _lin(e, sd, q + "self_attn.query").view(B, S, nh, hd).transpose(1, 2)
        kk = _lin(e, sd, q + "self_attn.key").view(B, S, nh, hd).transpose(1, 2)
        vv = _lin(e, sd, q + "self_attn.value").view(B, S, nh, hd).transpose(1, 2)
        att = torch.softmax(qq @ kk.transpose(-1, -2) / math.sqrt(hd), -1)
        ctx = (att @ vv).transpose(1, 2).reshape(B, S, E)
        e = _ln(e + _lin(ctx, sd, q + "self_attn.out_proj"), sd, q + "norm1", 1e-5)
        h = _lin(F.relu(_lin(e, sd, q + "linear1")), sd, q + "linear2")
        e = _ln(e + h, sd, q + "norm2", 1e-5)
    emb = e[:, 0]
    domain_logits = _lin(F.relu(_lin(emb, sd, p + "mlp_classifier.linear1")), sd, p + "mlp_classifier.linear2")
    vote = torch.softmax(domain_logits.sum(0, keepdim=True), -1)
    head = int(torch.argmax(vote, -1)) if force_head is None else int(force_head)
    conf = z["bin_configurations"][head]
    name = conf["name"]
    # seed bins (softplus -> un-normed)
    q = f"{p}seed_bin_regressors.{name}."
    prev_bin = F.softplus(_conv(F.relu(_conv(x, sd, q + "conv1")), sd, q + "conv2"))
    prev_emb = _conv(F.relu(_conv(x, sd, p + "seed_projector.conv1")), sd, p + "seed_projector.conv2")
    bin_centers = prev_bin
    for s, feat in enumerate(feature_blocks):
        q = f"{p}projectors.{s}."
        emb_s = _conv(F.relu(_conv(feat, sd, q + "conv1")), sd, q + "conv2")
        hh = emb_s + F.interpolate(prev_emb, emb_s.shape[-2:], mode="bilinear", align_corners=True)
        q = f"{p}attractors.{name}.{s}."
        A = F.softplus(_conv(F.relu(_conv(hh, sd, q + "conv1")), sd, q + "conv2"))
        c = F.interpolate(prev_bin, A.shape[-2:], mode="bilinear", align_corners=True)
        delta = torch.zeros_like(c)
        for a in range(A.shape[1]):
            dx = A[:, a:a + 1] - c
            delta = delta + dx / (1 + 300.0 * dx.pow(2))
        delta = delta / A.shape[1]
        prev_bin = c + delta
        bin_centers = prev_bin
        prev_emb = emb_s
    last = outconv
    bin_centers = F.interpolate(bin_centers, last.shape[-2:], mode="bilinear", align_corners=True)
    emb_up = F.interpolate(prev_emb, last.shape[-2:], mode="bilinear", align_corners=True)
    q = f"{p}conditional_log_binomial.{name}.mlp."
    pt = F.softplus(_conv(F.gelu(_conv(torch.cat([last, emb_up], 1), sd, q + "0")), sd, q + "2"))
    pr = pt[:, :2] + 1e-4
    pr = pr[:, 0] / (pr[:, 0] + pr[:, 1])
    tm = pt[:, 2:] + 1e-4
    tm = (tm[:, 0] / (tm[:, 0] + tm[:, 1])).unsqueeze(1)
    tm = (z["max_temp"] - z["min_temp"]) * tm + z["min_temp"]
    K = conf["n_bins"]
    k_idx = torch.arange(0, K).view(1, -1, 1, 1)
    km1 = torch.tensor([K - 1]).view(1, -1, 1, 1)
    pr = pr.unsqueeze(1)
    omp = torch.clamp(1 - pr, 1e-4, 1)
    prc = torch.clamp(pr, 1e-4, 1)
    y = _log_binom(km1, k_idx) + k_idx * torch.log(prc) + (km1 - k_idx) * torch.log(omp)
    prob = torch.softmax(y / tm, dim=1)
    depth = torch.sum(prob * bin_centers, dim=1, keepdim=True)
    return depth, domain_logits, head


def zoedepth_forward(sd, cfg, x384, force_head=None, return_aux=False):
    """HF:models/zoedepth/modeling_zoedepth.py:1252-1346 -> predicted_depth (B, 384, 384)"""
    b = cfg["vision_zoe_config"]["backbone_config"]
    hs = beit_forward(sd, cfg, x384)
    ph = pw = b["image_size"] // b["patch_size"]
    fused, bottleneck = zoe_neck(sd, cfg, hs, ph, pw)
    rel, feat = zoe_relative_head(sd, fused[-1])
    depth, dlog, head = zoe_metric_head(sd, cfg, feat, bottleneck, fused, force_head)
    if return_aux:
        return depth.squeeze(1), {"hidden": hs, "fused": fused, "bottleneck": bottleneck, "rel": rel,
                                  "outconv": feat, "domain_logits": dlog, "head": head}
    return depth.squeeze(1)


# ----------------------------------------------------------------------------------------------- Ego3D
def backproject_patch(K, depth, patch_size=14, reso=2):
    """model/modeling_spatialvla.py:181-185 (uv_h), :195-223. depth (B,1,H,W), K (3,3)|(B,3,3) -> (B, hp*wp, 12)"""
    b, c, h, w = depth.shape
    hp, wp = h // patch_size, w // patch_size
    step = patch_size // reso
    y, x = torch.meshgrid(torch.arange(0, h, step), torch.arange(0, w, step), indexing="ij")
    y, x = y + patch_size / reso / 2, x + patch_size / reso / 2
    uv_h = torch.stack([x, y, torch.ones_like(x)], 0).reshape(3, -1).float()
    pd = F.interpolate(depth, size=(hp * reso, wp * reso), mode="area").reshape(b, c, -1)
    p_cam = (torch.linalg.inv(K.float()) @ uv_h) * pd
    return p_cam.reshape(b, 3, hp, reso, wp, reso).permute(0, 2, 4, 3, 5, 1).reshape(b, hp * wp, -1)


def ego3d_encoding(xyz, n_freqs=8):
    """model/modeling_spatialvla.py:74-91"""
    center = torch.tensor([0.0, 0.0, 2.0]).repeat(xyz.shape[-1] // 3)
    freq = 2 ** torch.linspace(0, n_freqs - 1, n_freqs)
    xn = (xyz - center) / 2.0
    xf = xn.unsqueeze(-1) * freq
    return torch.cat([xn.unsqueeze(-1), torch.sin(xf), torch.cos(xf)], -1).reshape(*xyz.shape[:2], -1)


def ego3d_forward(sd, cfg, xyz):
    """model/modeling_spatialvla.py:59-64, :93-97"""
    p = "position_embedding_3d.position_embedding_head."
    h = _lin(ego3d_encoding(xyz, cfg["n_freqs"]), sd, p + "0")
    h = F.relu(_ln(h, sd, p + "1", 1e-5))
    return _lin(h, sd, p + "3")


def depth_to_224(depth384):
    """model/modeling_spatialvla.py:318-323"""
    d = F.interpolate(depth384.unsqueeze(1), size=(286, 286), mode="bicubic", align_corners=True)
    return d[..., 31:-31, 31:-31]


def image_features(sd, cfg, pixel_values, intrinsic, force_head=None, return_aux=False):
    """model/modeling_spatialvla.py:308-333 -> (B, 256, H_text)"""
    sig = siglip_forward(sd, cfg, (pixel_values - 0.5) / 0.5)
    aux = {"siglip": sig}
    if cfg.get("use_vision_zoe", True):
        with torch.no_grad():       # the reference runs ZoeDepth + back-projection under no_grad (:315-326): no gradient flows here
            depth384, zaux = zoedepth_forward(sd, cfg, process_zoe(pixel_values), force_head, return_aux=True)
            depth = depth_to_224(depth384)
            xyz = backproject_patch(intrinsic, depth, cfg["vision_config"]["patch_size"], cfg["ego3d_patch_reso"])
        pos = ego3d_forward(sd, cfg, xyz)
        aux.update({"depth384": depth384, "xyz": xyz, "pos3d": pos, "zoe": zaux})
        sig = sig + pos
    feat = _lin(sig, sd, "multi_modal_projector.linear") / (cfg["text_config"]["hidden_size"] ** 0.5)
    if return_aux:
        return feat, aux
    return feat


# ----------------------------------------------------------------------------------------------- Gemma2
def _rms(x, w, eps):
    """model/modeling_gemma2.py:60-77"""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + w)


def _rope(x, pos, theta):
    """model/modeling_gemma2.py:95-154: half-rotation layout, fp32 cos/sin. x (B,h,S,hd), pos (S,) or (B,S), 1-indexed"""
    hd = x.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    fr = pos.float()[..., None] * inv
    emb = torch.cat([fr, fr], -1)
    cos, sin = (emb.cos()[None, None], emb.sin()[None, None]) if pos.dim() == 1 else (emb.cos()[:, None], emb.sin()[:, None])
    x1, x2 = x[..., : hd // 2], x[..., hd // 2:]
    return x * cos + torch.cat([-x2, x1], -1) * sin


def embed_inputs(sd, cfg, input_ids, image_feats=None):
    """model/modeling_spatialvla.py:361-387 + model/modeling_gemma2.py:741-742 (x sqrt(H))."""
    t = cfg["text_config"]
    emb = sd["language_model.model.embed_tokens.weight"][input_ids].clone()
    lo = cfg["action_token_begin_idx"]
    if cfg.get("use_spatial_token", True):
        sel = (input_ids >= lo) & (input_ids < lo + cfg["spatial_token_num"])
        emb[sel] = sd["spatial_embed_tokens.weight"][input_ids[sel] - lo]
    if image_feats is not None:
        m = input_ids == cfg["image_token_index"]
        if int(m.sum()) * emb.shape[-1] != image_feats.numel():
            raise ValueError("Number of images does not match number of special image tokens in the input text.")
        emb[m] = image_feats.reshape(-1, emb.shape[-1]).to(emb.dtype)
    return emb * torch.tensor(t["hidden_size"] ** 0.5, dtype=emb.dtype)


def padded_positions(pos_start, S, pads):
    """Position ids of cache slots [pos_start, pos_start + S) for rows with `pads` leading padding tokens, as HF generate
    builds them (model/modeling_gemma2.py:1042-1051: cumsum(attention_mask) - 1, pads -> 1) plus the +1 of
    model/modeling_spatialvla.py:473-474: the k-th real token of a row sits at position k + 1, padding slots at 2."""
    slot = torch.arange(pos_start, pos_start + S)[None, :]
    pos = slot - pads[:, None] + 1
    return torch.where(slot < pads[:, None], torch.full_like(pos, 2), pos)


def gemma2_forward(sd, cfg, x, pos_start, kv_cache, bidirectional, pads=None, causal_prefix=0):
    """model/modeling_gemma2.py:364-413 (attention), :451-506 (sandwich-norm layer), :680-793 (stack),
    :993-997 (lm_head + final softcap is applied by the caller on the rows it keeps).
    x (B,S,H) already scaled; positions are pos_start+1 ... (1-indexed, model/modeling_spatialvla.py:371-372).
    kv_cache: list of [k, v] per layer (appended in place). Prefill is bidirectional over the prompt
    (model/modeling_spatialvla.py:291-297), decode rows attend to everything cached.
    pads (B,) int64: leading padding tokens per row (left-padded batch, attention_mask = 0 on them): their key columns are
    masked for every query (model/modeling_spatialvla.py:298-303) and positions restart at 1 on the first real token."""
    t = cfg["text_config"]
    nh, nkv, hd = t["num_attention_heads"], t["num_key_value_heads"], t["head_dim"]
    eps, theta = t["rms_norm_eps"], t.get("rope_theta", 10000.0)
    scale = t["query_pre_attn_scalar"] ** -0.5
    cap = t["attn_logit_softcapping"]
    B, S, H = x.shape
    pos = torch.arange(pos_start, pos_start + S) + 1 if pads is None else padded_positions(pos_start, S, pads)
    for i in range(t["num_hidden_layers"]):
        q = f"language_model.model.layers.{i}."
        h = _rms(x, sd[q + "input_layernorm.weight"], eps)
        qq = F.linear(h, sd[q + "self_attn.q_proj.weight"]).view(B, S, nh, hd).transpose(1, 2)
        kk = F.linear(h, sd[q + "self_attn.k_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
        vv = F.linear(h, sd[q + "self_attn.v_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
        qq, kk = _rope(qq, pos, theta), _rope(kk, pos, theta)
        if kv_cache[i] is None:
            kv_cache[i] = [kk, vv]
        else:
            kv_cache[i] = [torch.cat([kv_cache[i][0], kk], 2), torch.cat([kv_cache[i][1], vv], 2)]
        K, V = kv_cache[i]
        K = K.repeat_interleave(nh // nkv, 1)
        V = V.repeat_interleave(nh // nkv, 1)
        sc = (qq @ K.transpose(2, 3)) * scale
        if cap:
            sc = torch.tanh(sc / cap) * cap
        if not bidirectional and S > 1:
            L = K.shape[2]
            # causal_prefix > 0: prefix-LM mask of the training forward -- token_type_ids == 0 columns are unmasked on top of the
            # triangular mask (model/modeling_spatialvla.py:292-293,304-305)
            mask = torch.arange(L)[None, :] > torch.clamp(torch.arange(S)[:, None] + (L - S), min=causal_prefix - 1)
            sc = sc.masked_fill(mask, float("-inf"))
        win = t.get("sliding_window")
        if win and i % 2 == 0 and K.shape[2] > win:
            # sliding-window layers (even layer_idx, model/modeling_gemma2.py:343,441): key slot j is masked for query slot i when
            # i - j >= window (:461-471, tril(diagonal=-window) on top of whatever mask the layer received -- also on top of the
            # bidirectional prefix mask); a decode row therefore sees the last `window` slots (HF's HybridCache keeps exactly those)
            L = K.shape[2]
            qslot = torch.arange(S)[:, None] + (L - S)
            sc = sc.masked_fill((qslot - torch.arange(L)[None, :]) >= win, float("-inf"))
        if pads is not None:
            sc = sc.masked_fill((torch.arange(K.shape[2])[None, :] < pads[:, None])[:, None, None, :], float("-inf"))
        ctx = (torch.softmax(sc, -1) @ V).transpose(1, 2).reshape(B, S, nh * hd)
        a = F.linear(ctx, sd[q + "self_attn.o_proj.weight"])
        x = x + _rms(a, sd[q + "post_attention_layernorm.weight"], eps)
        h = _rms(x, sd[q + "pre_feedforward_layernorm.weight"], eps)
        m = F.linear(gelu_tanh(F.linear(h, sd[q + "mlp.gate_proj.weight"])) * F.linear(h, sd[q + "mlp.up_proj.weight"]),
                     sd[q + "mlp.down_proj.weight"])
        x = x + _rms(m, sd[q + "post_feedforward_layernorm.weight"], eps)
    return _rms(x, sd["language_model.model.norm.weight"], eps)


def lm_head_slice(sd, cfg, h, lo, hi):
    """model/modeling_gemma2.py:993-997 restricted to vocabulary rows [lo, hi)."""
    cap = cfg["text_config"]["final_logit_softcapping"]
    lg = F.linear(h, sd["language_model.lm_head.weight"][lo:hi])
    if cap:
        lg = torch.tanh(lg / cap) * cap
    return lg


def left_pads(attention_mask):
    """(B,P) 0/1 mask -> (B,) number of leading zeros; raises unless every row is zeros followed by ones (left padding, the
    Gemma tokenizer's padding side) with at least one real token."""
    am = attention_mask.to(torch.int64)
    pads = (am == 0).sum(1)
    P = am.shape[1]
    expect = (torch.arange(P)[None, :] >= pads[:, None]).to(torch.int64)
    if not torch.equal(am.cpu(), expect.cpu()) or int(pads.max()) >= P:
        raise NotImplementedError("only left-padded batches (attention_mask = 0...01...1) are supported")
    return pads


def predict_action_ref(sd, cfg, input_ids, pixel_values, intrinsic, n_new, forced_tokens=None, force_head=None,
                       return_aux=False, attention_mask=None):
    """Greedy action-token decode = model/modeling_spatialvla.py:484-492 with the action-restricted argmax
    of SURVEY.md §7.  Returns (tokens (B,n_new) int64, logits (B,n_new,n_action) fp32 post-softcap)."""
    lo = cfg["action_token_begin_idx"]
    hi = lo + cfg["spatial_token_num"]
    with torch.no_grad():
        feats, aux = image_features(sd, cfg, pixel_values, intrinsic, force_head, return_aux=True)
        x = embed_inputs(sd, cfg, input_ids, feats)
        B, P, _ = x.shape
        cache = [None] * cfg["text_config"]["num_hidden_layers"]
        pads = None if attention_mask is None else left_pads(attention_mask)
        h = gemma2_forward(sd, cfg, x, 0, cache, bidirectional=True, pads=pads)
        toks, logs = [], []
        for step in range(n_new):
            lg = lm_head_slice(sd, cfg, h[:, -1], lo, hi)
            logs.append(lg)
            nxt = lg.argmax(-1) + lo
            toks.append(nxt)
            if step == n_new - 1:
                break
            feed = nxt if forced_tokens is None else forced_tokens[:, step]
            x = embed_inputs(sd, cfg, feed[:, None])
            h = gemma2_forward(sd, cfg, x, P + step, cache, bidirectional=False, pads=pads)
    out = (torch.stack(toks, 1), torch.stack(logs, 1))
    if return_aux:
        aux["image_features"] = feats
        return out + (aux,)
    return out


def generate_ref(sd, cfg, input_ids, pixel_values, intrinsic, max_new_tokens, eos_id, pad_id, force_head=None):
    """HF greedy `generate` as the reference's predict_action calls it (model/modeling_spatialvla.py:484-492; GenerationMixin
    4.47 `_sample` with do_sample=False): full-vocabulary argmax of the post-softcap logits, finished rows emit pad_id, a row
    finishes on eos_id, stop when all rows are finished or after max_new_tokens.  Returns the NEW tokens int64 [B, n]."""
    V = cfg["text_config"]["vocab_size"]
    with torch.no_grad():
        feats = image_features(sd, cfg, pixel_values, intrinsic, force_head)
        x = embed_inputs(sd, cfg, input_ids, feats)
        B, P, _ = x.shape
        cache = [None] * cfg["text_config"]["num_hidden_layers"]
        h = gemma2_forward(sd, cfg, x, 0, cache, bidirectional=True)
        unfinished = torch.ones(B, dtype=torch.bool)
        out = []
        for step in range(max_new_tokens):
            nxt = lm_head_slice(sd, cfg, h[:, -1], 0, V).argmax(-1)
            nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_id))
            out.append(nxt)
            unfinished = unfinished & (nxt != eos_id)
            if step == max_new_tokens - 1 or not bool(unfinished.any()):
                break
            h = gemma2_forward(sd, cfg, embed_inputs(sd, cfg, nxt[:, None]), P + step, cache, bidirectional=False)
    return torch.stack(out, 1)


def prefix_length(token_type_ids):
    """(B, L) token types -> p such that every row is p zeros followed by L - p ones (prefix / suffix of the training samples,
    train/monkey_patch.py:21-75); raises for any other pattern."""
    tt = token_type_ids.to(torch.int64).cpu()
    p = int((tt[0] == 0).sum())
    expect = (torch.arange(tt.shape[1])[None, :] >= p).to(torch.int64).expand_as(tt)
    if not torch.equal(tt, expect):
        raise NotImplementedError("token_type_ids must be 0...01...1 with the same prefix length in every row")
    return p


def forward_loss_ref(sd, cfg, input_ids, pixel_values, intrinsic, labels, token_type_ids=None, attention_mask=None,
                     force_head=None, ignore_index=-100, pad_token_id=0, image_feats=None):
    """forward() with labels = model/modeling_spatialvla.py:335-430.  Mask (`_update_causal_mask`, :258-306): training
    (token_type_ids and labels given) = triangular, plus -- only when a 2-D attention_mask is passed -- every token_type 0
    column unmasked (prefix-LM); labels without token_type_ids = the inference mask (bidirectional).  Loss = shifted
    nn.CrossEntropyLoss over the full vocabulary, post-softcap logits, ignore_index rows dropped (:413-430).
    Returns (loss 0-dim fp32, flat row indices b*L+t of the labelled rows, their labels, their logits fp32 [R, V])."""
    with torch.no_grad():
        return _forward_loss(sd, cfg, input_ids, pixel_values, intrinsic, labels, token_type_ids, attention_mask, force_head,
                             ignore_index, pad_token_id, image_feats)


def _forward_loss(sd, cfg, input_ids, pixel_values, intrinsic, labels, token_type_ids, attention_mask, force_head, ignore_index,
                  pad_token_id, image_feats):
    if attention_mask is not None and bool((attention_mask == 0).any()):
        raise NotImplementedError("padded batches are not covered by the labelled forward")
    feats = image_feats             # precomputed get_image_features output (full-size tests reuse the one they already have)
    if feats is None and pixel_values is not None:
        feats = image_features(sd, cfg, pixel_values, intrinsic, force_head)
    x = embed_inputs(sd, cfg, input_ids, feats)
    B, L, _ = x.shape
    cache = [None] * cfg["text_config"]["num_hidden_layers"]
    if token_type_ids is not None:
        prefix = prefix_length(token_type_ids) if attention_mask is not None else 0
        h = gemma2_forward(sd, cfg, x, 0, cache, bidirectional=False, causal_prefix=prefix)
    else:
        h = gemma2_forward(sd, cfg, x, 0, cache, bidirectional=True)
    if bool((labels == pad_token_id).any()):                       # :392-397
        labels = torch.where(input_ids == pad_token_id, torch.full_like(labels, ignore_index), labels)
    shift_labels = labels[:, 1:]
    bi, ti = torch.nonzero(shift_labels != ignore_index, as_tuple=True)
    rows = bi * L + ti
    lab = shift_labels[bi, ti]
    lg = lm_head_slice(sd, cfg, h.reshape(B * L, -1)[rows], 0, cfg["text_config"]["vocab_size"]).float()
    loss = F.cross_entropy(lg, lab) if rows.numel() else torch.tensor(float("nan"))
    return loss, rows, lab, lg


def training_metrics_ref(logit_rows, row_labels, actions, ranges, decode_fn):
    """train/monkey_patch.py:267-324 on the labelled rows (positions with label -100 never pass its action-range mask, so the
    block only ever looks at labelled rows).  ranges = {"translation": (start, end), "rotation": ..., "gripper": ...} inclusive
    token-id ranges of the sub-tokenizers; decode_fn: (n,3) int64 numpy global ids -> (n,7) float64 actions."""
    pred = logit_rows.argmax(-1)
    lo, hi = ranges["translation"][0], ranges["gripper"][1]
    mask = (row_labels >= lo) & (row_labels <= hi)
    gt, pr = row_labels[mask], pred[mask]
    out = {"accuracy": float((gt == pr).sum().float() / mask.sum().float())}
    for name in ("translation", "rotation", "gripper"):
        m = (gt >= ranges[name][0]) & (gt <= ranges[name][1])
        out[name + "_accuracy"] = float((gt[m] == pr[m]).sum().float() / m.sum().float())
    pred_actions = torch.tensor(decode_fn(pr.numpy().reshape(-1, 3)))
    gt_actions = torch.as_tensor(actions).reshape(-1, 7).to(torch.float32)
    out["l1_loss"] = float(F.l1_loss(pred_actions.to(torch.float32), gt_actions))
    return out


def loss_and_grads_ref(sd, cfg, input_ids, pixel_values, intrinsic, labels, grad_keys, token_type_ids=None, attention_mask=None,
                       force_head=None):
    """Oracle of the BACKWARD half of the training step (next round's kernels are checked against it):
    torch autograd through the restated forward, dLoss/dW for the weights named in grad_keys (fp32).  A LoRA adapter on a
    Linear W (PEFT: W + (alpha/r) B A, train/spatialvla_finetune.py:262-302) gets its gradients from dW by the chain rule,
    dB = (alpha/r) dW A^T and dA = (alpha/r) B^T dW, so full-weight gradients pin every adapter gradient.  ZoeDepth is under
    no_grad exactly like the reference (model/modeling_spatialvla.py:315-326)."""
    sd2 = dict(sd)
    leaves = {}
    for k in grad_keys:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        sd2[k] = leaves[k]
    with torch.enable_grad():
        loss, rows, lab, lg = _forward_loss(sd2, cfg, input_ids, pixel_values, intrinsic, labels, token_type_ids, attention_mask,
                                            force_head, -100, 0, None)
        loss.backward()
    return loss.detach(), {k: v.grad.detach() for k, v in leaves.items()}


def loss_tail_grads_ref(sd, cfg, h_rows, row_labels):
    """d(mean CE)/d(final-normed hidden rows) by autograd through lm_head + soft-cap + nn.CrossEntropyLoss
    (model/modeling_gemma2.py:993-997, model/modeling_spatialvla.py:413-430).  h_rows fp32 [R, H] -> (loss, dh [R, H])."""
    hr = h_rows.detach().float().clone().requires_grad_(True)
    with torch.enable_grad():
        lg = lm_head_slice(sd, cfg, hr, 0, cfg["text_config"]["vocab_size"]).float()
        loss = F.cross_entropy(lg, row_labels)
        loss.backward()
    return loss.detach(), hr.grad.detach()
