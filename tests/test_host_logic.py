"""CPU: the host orchestration (weight repack + kernel sequence) run through the torch op re-statements
(oracle/ops_ref.py) against the fp32 oracle and the golden vectors; config / processor / sharding helpers."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model_ref as R
from oracle.gen_golden import tiny_inputs
from oracle.ops_ref import RefOps
from spatialvla_b200.engine import SpatialVLAEngine, pack_conv3x3, pack_conv3x3_im2col, pack_deconv
from spatialvla_b200.weights import synth_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tiny():
    cfg, px_u8, ids, K = tiny_inputs()
    sd = synth_state_dict(cfg, seed=0)
    return cfg, px_u8.float() / 255.0, ids, K, sd, SpatialVLAEngine(cfg, sd, RefOps())


def test_engine_matches_oracle_and_golden(tiny):
    cfg, px, ids, K, sd, eng = tiny
    g = np.load(os.path.join(GOLD, "tiny_model.npz"))
    with torch.no_grad():
        feats, aux = eng.image_features(px, K, return_aux=True)
        toks, logits = eng.generate_actions(ids, px, K, int(g["n_new"]), return_logits=True)
    assert eng.last_router_head == int(np.argmax(g["domain_logits"].sum(0)))
    assert np.abs(aux["siglip"].view(2, 256, -1)[:, ::4].numpy() - g["siglip"]).max() < 3e-2 * np.abs(g["siglip"]).max()
    d = aux["depth384"][:, ::4, ::4].numpy()
    assert np.abs(d - g["depth384_s4"]).max() < 5e-3            # metric depth, metres; bf16 operands
    assert np.abs(aux["xyz"].numpy() - g["xyz"]).max() < 5e-3
    assert np.abs(feats[:, ::4].numpy() - g["image_features"]).max() < 2e-2 * np.abs(g["image_features"]).max()
    assert np.array_equal(toks.numpy(), g["tokens"])
    assert np.abs(logits.numpy() - g["logits"]).max() < 6e-2


def test_decode_hilo_chain_is_closer_to_the_fp32_oracle(tiny):
    """The decode chain carries hi/lo bf16 activation pairs and the last prompt row is re-evaluated as a decode step
    (engine.gemma_forward / language_stage): same tokens as the plain bf16 chain, logits closer to the fp32 oracle on the
    language stage (the oracle's own image features are fed so that only the Gemma2 arithmetic differs)."""
    cfg, px, ids, K, sd, eng = tiny
    tk, lg, aux = R.predict_action_ref(sd, cfg, ids, px, K, 6, force_head=0, return_aux=True)
    feats = aux["image_features"]
    err = {}
    try:
        for mode in (True, False):
            eng.decode_hilo = mode
            logs = []
            with torch.no_grad():
                toks = eng.language_stage(ids, feats, 6, forced_tokens=tk, logs=logs)
            assert torch.equal(toks, tk)
            err[mode] = float((torch.stack(logs, 1) - lg).pow(2).mean().sqrt())
    finally:
        del eng.decode_hilo                     # back to the class default
    assert err[True] < 0.6 * err[False], err


def test_language_stage_batch1_equals_its_row_in_a_batch(tiny):
    """A single observation decodes like the same observation inside a batch (the re-evaluation of the last prompt row starts from
    a COPY of its embedding: at batch 1 the row slice is already contiguous and would alias the in-place residual stream)."""
    cfg, px, ids, K, sd, eng = tiny
    with torch.no_grad():
        feats = eng.image_features(px, K)
        logs2, logs1 = [], []
        t2 = eng.language_stage(ids, feats, 4, logs=logs2)
        t1 = eng.language_stage(ids[1:2], feats[1:2].contiguous(), 4, logs=logs1)
    assert torch.equal(t1[0], t2[1])
    assert float((torch.stack(logs1, 1)[0] - torch.stack(logs2, 1)[1]).abs().max()) < 1e-3


def test_sliding_window_layers_vs_reference_golden():
    """Gemma2's sliding-window / global layer alternation (even layer_idx windowed, model/modeling_gemma2.py:343,441-473) through
    the engine's kernel sequence: tiny config with sliding_window = 48 << 264 prompt tokens, so the window predicate is active in
    the prefill attention and in every fused decode step; golden minted from the live reference (tests/golden/tiny_model_window.npz)."""
    g = np.load(os.path.join(GOLD, "tiny_model_window.npz"))
    cfg, px_u8, ids, K = tiny_inputs()
    cfg["text_config"]["sliding_window"] = int(g["window"])
    sd = synth_state_dict(cfg, seed=0)
    eng = SpatialVLAEngine(cfg, sd, RefOps())
    px = px_u8.float() / 255.0
    with torch.no_grad():
        toks, logits = eng.generate_actions(ids, px, K, int(g["n_new"]), return_logits=True)
    assert np.array_equal(toks.numpy(), g["tokens"])
    assert np.abs(logits.numpy() - g["logits"]).max() < 6e-2
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    m = SpatialVLAForConditionalGeneration(cfg, sd, ops=RefOps())
    out = m.forward(input_ids=ids, pixel_values=px, intrinsic=K)                  # every prefill position
    assert np.abs(out.logits[:, :, torch.from_numpy(g["prefill_cols"])].numpy() - g["prefill_logits"]).max() < 8e-2


def test_forward_logits_api(tiny):
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    cfg, px, ids, K, sd, eng = tiny
    m = SpatialVLAForConditionalGeneration(cfg, sd, ops=RefOps())
    out = m.forward(input_ids=ids, pixel_values=px, intrinsic=K, num_logits_to_keep=1)
    lo = cfg["action_token_begin_idx"]
    toks, logits = R.predict_action_ref(sd, cfg, ids, px, K, 1, force_head=m.engine.last_router_head)
    assert out.logits.shape == (2, 1, cfg["text_config"]["vocab_size"])
    assert (out.logits[:, 0, lo:lo + cfg["spatial_token_num"]] - logits[:, 0]).abs().max() < 6e-2
    pa = m.predict_action({"input_ids": ids, "pixel_values": px, "intrinsic": K}, max_new_tokens=3)
    assert pa.shape == (2, 3) and pa.dtype == torch.int64
    with pytest.raises(ValueError):
        m.predict_action({"input_ids": ids[:, 5:], "pixel_values": px, "intrinsic": K})
    with pytest.raises(NotImplementedError):      # all-padding rows / right padding / holes: rejected, never mis-computed
        m.predict_action({"input_ids": ids, "pixel_values": px, "intrinsic": K, "attention_mask": torch.zeros_like(ids)})
    am = torch.ones_like(ids)
    am[:, -1] = 0
    with pytest.raises(NotImplementedError):
        m.predict_action({"input_ids": ids, "pixel_values": px, "intrinsic": K, "attention_mask": am})


def test_left_padded_batch_matches_reference_golden():
    """predict_action on a LEFT-padded batch through the public API (host orchestration over the torch op re-statements)
    against the golden vectors the live reference produced for the same padded inputs."""
    from spatialvla_b200.configs import get_config_dict
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    g = np.load(os.path.join(GOLD, "tiny_model_padded.npz"))
    cfg = get_config_dict("tiny")
    sd = synth_state_dict(cfg, seed=0)
    m = SpatialVLAForConditionalGeneration(cfg, sd, ops=RefOps())
    ids, am = torch.from_numpy(g["input_ids"]), torch.from_numpy(g["attention_mask"])
    px, K = torch.from_numpy(g["pixel_u8"]).float() / 255.0, torch.from_numpy(g["intrinsic"])
    toks, logits = m.predict_action({"input_ids": ids, "attention_mask": am, "pixel_values": px, "intrinsic": K},
                                    max_new_tokens=int(g["n_new"]), return_logits=True)
    assert np.array_equal(toks.numpy(), g["tokens"])
    assert np.abs(logits.numpy() - g["logits"]).max() < 6e-2
    # forward(): logits of the last prompt position with the 2-D mask, then one cached decode step that keeps masking the pads
    lo = cfg["action_token_begin_idx"]
    fw = m.forward(input_ids=ids, pixel_values=px, intrinsic=K, attention_mask=am, num_logits_to_keep=1)
    assert (fw.logits[:, 0, lo:lo + 8194] - torch.from_numpy(g["logits"][:, 0])).abs().max() < 6e-2
    nxt = torch.from_numpy(g["tokens"][:, :1])
    fw2 = m.forward(input_ids=nxt, past_key_values=fw.past_key_values, num_logits_to_keep=1)
    assert (fw2.logits[:, 0, lo:lo + 8194] - torch.from_numpy(g["logits"][:, 1])).abs().max() < 6e-2


def test_weight_packing_layouts():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 10, 12, generator=g)            # NCHW
    w = torch.randn(5, 8, 3, 3, generator=g)
    ref = F.conv2d(x, w, padding=1)
    wp = pack_conv3x3(w).view(5, 9, 64)[:, :, :8]
    xp = F.pad(x, (1, 1, 1, 1))
    acc = sum(torch.einsum("nchw,oc->nohw", xp[:, :, t // 3:t // 3 + 10, t % 3:t % 3 + 12], wp[:, t]) for t in range(9))
    assert torch.allclose(acc, ref, atol=1e-4)
    ops = RefOps()
    col = torch.zeros(2 * 5 * 6, 9 * 8, dtype=torch.bfloat16)
    ops.im2col3x3_s2(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), col, batch=2, h=10, w=12, c=8)
    xb = x.to(torch.bfloat16).float()
    ref2 = F.conv2d(xb, w, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, 5)
    assert torch.allclose(col.float() @ pack_conv3x3_im2col(w).t(), ref2, atol=1e-3)
    wt, bt = torch.randn(8, 6, 4, 4, generator=g), torch.randn(6, generator=g)
    wg, bg = pack_deconv(wt, bt)
    ref3 = F.conv_transpose2d(x, wt, bt, stride=4)
    gm = (x.permute(0, 2, 3, 1).reshape(-1, 8) @ wg.t() + bg).to(torch.bfloat16)
    out = torch.zeros(2, 40, 48, 6, dtype=torch.bfloat16)
    # pixel_shuffle wants c % 8 == 0 on the GPU; the torch re-statement has no such limit
    ops.pixel_shuffle(gm, out, batch=2, h=10, w=12, c=6, f=4)
    assert torch.allclose(out.float().permute(0, 3, 1, 2), ref3, atol=5e-2)


def test_config_roundtrip():
    from spatialvla_b200 import SpatialVLAConfig, get_config_dict
    c = SpatialVLAConfig(**get_config_dict("tiny"))
    d = c.to_engine_dict()
    assert d["text_config"]["head_dim"] == 256 and d["vision_zoe_config"]["backbone_config"]["hidden_size"] == 128
    assert d["text_config"]["rope_theta"] == 10000.0
    assert c.text_config.num_image_tokens == 256
    c2 = SpatialVLAConfig(**json.loads(json.dumps({k: v for k, v in c.to_dict().items() if k in
                                                   ("vision_config", "text_config", "vision_zoe_config", "image_token_index",
                                                    "action_token_begin_idx", "spatial_token_num", "use_spatial_token",
                                                    "ego3d_patch_reso", "n_freqs", "use_vision_zoe")})))
    assert c2.to_engine_dict()["vision_config"]["hidden_size"] == d["vision_config"]["hidden_size"]


def test_shard_bounds():
    from spatialvla_b200.parallel import shard_batch, shard_bounds
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    b = {"input_ids": torch.arange(10).view(5, 2), "pixel_values": torch.zeros(5, 3, 4, 4), "intrinsic": torch.eye(3)}
    s = shard_batch(b, 1, 2)
    assert s["input_ids"].shape[0] == 2 and s["intrinsic"].shape == (3, 3)


def test_from_pretrained_checkpoint_roundtrip(tmp_path):
    """Checkpoint interchange: a directory holding config.json + *.safetensors in the reference's
    own key layout loads through `from_pretrained`; like the reference (model/modeling_spatialvla.py:524-525) the last
    `spatial_token_num` rows of embed_tokens are overwritten with spatial_embed_tokens; predict_action on the loaded model
    reproduces the golden tokens of the live reference."""
    from safetensors.torch import save_file
    from spatialvla_b200 import SpatialVLAConfig, get_config_dict
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    from spatialvla_b200.weights import state_dict_spec
    cfg = get_config_dict("tiny")
    sd = synth_state_dict(cfg, seed=0)
    assert set(sd) == set(state_dict_spec(cfg))                 # the reference's key layout, nothing else
    SpatialVLAConfig(**cfg).save_pretrained(tmp_path)
    keys = sorted(sd)
    half = len(keys) // 2                                   # two shards, like a sharded hub checkpoint
    save_file({k: sd[k].contiguous() for k in keys[:half]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: sd[k].contiguous() for k in keys[half:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    m = SpatialVLAForConditionalGeneration.from_pretrained(str(tmp_path), ops=RefOps())
    assert m.config.spatial_token_num == cfg["spatial_token_num"] and m.engine_config["text_config"]["head_dim"] == 256
    g = np.load(os.path.join(GOLD, "tiny_model.npz"))
    cfg_, px_u8, ids, K = tiny_inputs()
    toks = m.predict_action({"input_ids": ids, "pixel_values": px_u8.float() / 255.0, "intrinsic": K}, max_new_tokens=int(g["n_new"]))
    assert np.array_equal(toks.numpy(), g["tokens"])
    n = cfg["spatial_token_num"]
    assert torch.equal(m.engine.gem["embed"][-n:].float(), sd["spatial_embed_tokens.weight"].to(torch.bfloat16).float())
    with pytest.raises(OSError):
        SpatialVLAForConditionalGeneration.from_pretrained(str(tmp_path / "missing"), ops=RefOps())


def test_tied_lm_head_checkpoint_loads_and_follows_the_spatial_overwrite(tmp_path):
    """Gemma2 ties lm_head to embed_tokens (model/modeling_gemma2.py:888, model/modeling_spatialvla.py:171-172): a checkpoint saved
    by the reference with tie_word_embeddings=True has NO lm_head key, and the post-load overwrite of embed_tokens[-n:] with
    spatial_embed_tokens (:524-525) therefore also replaces the action-slice rows of the head."""
    from safetensors.torch import save_file
    from spatialvla_b200 import SpatialVLAConfig, get_config_dict
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    cfg = get_config_dict("tiny")
    cfg["text_config"]["tie_word_embeddings"] = True
    sd = {k: v for k, v in synth_state_dict(get_config_dict("tiny"), seed=0).items() if k != "language_model.lm_head.weight"}
    SpatialVLAConfig(**cfg).save_pretrained(tmp_path)
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "model.safetensors"))
    m = SpatialVLAForConditionalGeneration.from_pretrained(str(tmp_path), ops=RefOps())
    n, lo = cfg["spatial_token_num"], cfg["action_token_begin_idx"]
    V = cfg["text_config"]["vocab_size"]
    assert lo + n == V                                        # the action slice is the tail of the vocabulary
    want = sd["spatial_embed_tokens.weight"].to(torch.bfloat16).float()
    assert torch.equal(m.engine.gem["head_act"].float(), want)
    assert torch.equal(m.engine.lm_head_full()[-n:].float(), want)
    assert torch.equal(m.engine.lm_head_full()[:lo].float(), sd["language_model.model.embed_tokens.weight"][:lo].to(torch.bfloat16).float())
    # a plain dict without the key (no config flag) aliases as well instead of raising KeyError
    eng = SpatialVLAEngine(get_config_dict("tiny"), sd, RefOps())
    assert eng.gem["head_act"].shape == (n, cfg["text_config"]["hidden_size"])


def test_reference_generate_semantics_full_vocab_argmax_and_eos_stop(tiny):
    """predict_action(reference_generate=True) = the reference's own decoding rule (model/modeling_spatialvla.py:484-492):
    full-vocabulary argmax, EOS stop with padded finished rows, max_new_tokens bound; against the oracle's restatement of HF's
    greedy loop.  With random weights no EOS appears by itself, so the EOS id is set to a token the rows really emit."""
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    cfg, px, ids, K, sd, eng = tiny
    m = SpatialVLAForConditionalGeneration(cfg, sd, ops=RefOps())
    free = R.generate_ref(sd, cfg, ids, px, K, 5, eos_id=-7, pad_id=0, force_head=None)
    got = m.generate({"input_ids": ids, "pixel_values": px, "intrinsic": K}, max_new_tokens=5, eos_token_id=-7, pad_token_id=0)
    assert got.shape == (2, ids.shape[1] + 5) and torch.equal(got[:, :ids.shape[1]], ids) and torch.equal(got[:, ids.shape[1]:], free)
    eos = int(free[0, 1])                                    # row 0 stops after its 2nd token
    want = R.generate_ref(sd, cfg, ids, px, K, 5, eos_id=eos, pad_id=0)
    new = m.predict_action({"input_ids": ids, "pixel_values": px, "intrinsic": K}, max_new_tokens=5, reference_generate=True) \
        if cfg.get("eos_token_id") == eos else m.generate({"input_ids": ids, "pixel_values": px, "intrinsic": K}, max_new_tokens=5,
                                                          eos_token_id=eos, pad_token_id=0)[:, ids.shape[1]:]
    assert torch.equal(new, want)
    assert int(want[0, 1]) == eos and (want.shape[1] == 2 or bool((want[0, 2:] == 0).all()))
    with pytest.raises(NotImplementedError):
        m.generate({"input_ids": ids, "pixel_values": px, "intrinsic": K}, do_sample=True)


def test_labelled_forward_host_logic_matches_oracle_and_golden(tiny):
    """forward(labels=...) orchestration (mask selection, label shift / ignore / pad masking, row gather, chunked lm_head +
    cross entropy) through the torch op re-statements, against the fp32 oracle and the live-reference golden losses."""
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    cfg, _, _, _, sd, _ = tiny
    g = np.load(os.path.join(GOLD, "tiny_model_train.npz"))
    ids, tt, labels = (torch.from_numpy(g[k]) for k in ("input_ids", "token_type_ids", "labels"))
    px, K = torch.from_numpy(g["pixel_u8"]).float() / 255.0, torch.from_numpy(g["intrinsic"])
    B, L = ids.shape
    m = SpatialVLAForConditionalGeneration(cfg, sd, ops=RefOps())
    ones = torch.ones(B, L, dtype=torch.int64)
    for name, kw in (("prefix_lm", dict(token_type_ids=tt, attention_mask=ones)), ("causal", dict(token_type_ids=tt)),
                     ("bidirectional", dict())):
        out = m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels, **kw)
        ref_loss, rows, lab, ref_lg = R.forward_loss_ref(sd, cfg, ids, px, K, labels, force_head=m.engine.last_router_head, **kw)
        assert torch.equal(out.label_rows, rows) and out.logits.shape == ref_lg.shape
        assert (out.logits - ref_lg).abs().max() < 6e-2, name
        assert abs(float(out.loss) - float(ref_loss)) < 1e-2 and abs(float(out.loss) - float(g["loss_" + name])) < 1e-2, name
        assert abs(float(out.row_loss.mean()) - float(out.loss)) < 1e-5
        assert 0.0 <= float(out.token_accuracy) <= 1.0
    # chunked path: two rows per lm_head / cross-entropy launch gives the same loss, logits are not kept
    m.engine.loss_chunk_rows = 4
    out2 = m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels)
    assert abs(float(out2.loss) - float(out.loss)) < 1e-6 and out2.logits is None
    m.engine.loss_chunk_rows = 4096
    # pad-id labels are ignored where the input token is the pad token (model/modeling_spatialvla.py:389-397)
    ids_p, lab_p = ids.clone(), labels.clone()
    ids_p[:, -1], lab_p[:, -1] = 0, 0
    out3 = m.forward(input_ids=ids_p, pixel_values=px, intrinsic=K, labels=lab_p, token_type_ids=tt, attention_mask=ones)
    assert out3.label_rows.numel() == rows.numel() - B
    # nothing labelled -> NaN like nn.CrossEntropyLoss; unsupported patterns raise
    assert torch.isnan(m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=torch.full_like(labels, -100)).loss)
    am = ones.clone()
    am[0, 0] = 0
    with pytest.raises(NotImplementedError):
        m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels, token_type_ids=tt, attention_mask=am)
    bad_tt = tt.clone()
    bad_tt[1, 3] = 1
    with pytest.raises(NotImplementedError):
        m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels, token_type_ids=bad_tt, attention_mask=ones)
    with pytest.raises(ValueError):
        m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels[:, :-1])


def _metric_tokenizer(cfg):
    """Real SpatialActionTokenizer over the fake HF tokenizer; its host decode (a CUDA kernel behind the C ABI) is replaced by
    the oracle restatement so the test runs without a GPU."""
    from fakes import FakeTokenizer
    from oracle import tokenizer_ref as T
    from spatialvla_b200.action_tokenizer import SpatialActionTokenizer
    nb = {"translation": {"theta_bins": 16, "phi_bins": 32, "r_bins": 8}, "rotation": {"roll_bins": 16, "pitch_bins": 16, "yaw_bins": 16},
          "gripper": 2, "total": 8194}
    tk = SpatialActionTokenizer(FakeTokenizer(base=cfg["action_token_begin_idx"]), nb)
    assert tk.action_token_begin_idx == cfg["action_token_begin_idx"]
    tk.decode_token_ids_to_actions = lambda ids: T.decode(np.asarray(ids) - tk.action_token_begin_idx, tk.bin_policy, nb)
    ranges = {k: (getattr(tk, k + "_tokenizer").token_start_idx, getattr(tk, k + "_tokenizer").token_end_idx)
              for k in ("translation", "rotation", "gripper")}
    return tk, ranges


def test_action_metrics_match_reference_metric_block(tiny):
    """SpatialVLAForConditionalGeneration.action_metrics vs the restated metric block of the reference's training step
    (train/monkey_patch.py:267-324) on the same labelled-row logits."""
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    cfg, _, _, _, sd, _ = tiny
    g = np.load(os.path.join(GOLD, "tiny_model_train.npz"))
    ids, tt, labels = (torch.from_numpy(g[k]) for k in ("input_ids", "token_type_ids", "labels"))
    px, K = torch.from_numpy(g["pixel_u8"]).float() / 255.0, torch.from_numpy(g["intrinsic"])
    m = SpatialVLAForConditionalGeneration(cfg, sd, ops=RefOps())
    tk, ranges = _metric_tokenizer(cfg)
    # teach the "model" two of the labels so that the accuracies are not all zero: labels := its own argmax on some rows
    out = m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=labels, token_type_ids=tt)
    lab2 = labels.clone()
    B, L = ids.shape
    for r, a in zip(out.label_rows[:5].tolist(), out.row_argmax[:5].tolist()):
        if ranges["translation"][0] <= a <= ranges["gripper"][1]:
            lab2[r // L, r % L + 1] = a
    out = m.forward(input_ids=ids, pixel_values=px, intrinsic=K, labels=lab2, token_type_ids=tt)
    actions = torch.rand(B, 2, 7, generator=torch.Generator().manual_seed(3)) * 2 - 1
    got = m.action_metrics(out, actions, tk)
    ref = R.training_metrics_ref(out.logits, out.row_labels, actions, ranges, tk.decode_token_ids_to_actions)
    assert set(got) == {"accuracy", "translation_accuracy", "rotation_accuracy", "gripper_accuracy", "l1_loss"}
    for k in ref:
        assert (np.isnan(got[k]) and np.isnan(ref[k])) or abs(got[k] - ref[k]) < 1e-6, (k, got[k], ref[k])
    assert abs(got["accuracy"] - float(out.token_accuracy) * out.row_labels.numel() / 12) < 1e-6      # EOS rows are not action rows


def test_loss_tail_backward_matches_autograd_oracle(tiny):
    """engine.labelled_loss_backward (cross-entropy backward kernel + dz @ W_head GEMM, here through the torch op re-statements)
    against autograd through the oracle's lm_head + soft-cap + cross entropy."""
    cfg, _, _, _, sd, eng = tiny
    g = torch.Generator().manual_seed(11)
    H, V = cfg["text_config"]["hidden_size"], cfg["text_config"]["vocab_size"]
    h = (torch.randn(40, H, generator=g) * 2.0).to(torch.bfloat16)
    rows = torch.tensor([3, 4, 9, 17, 18, 30, 39])
    lab = torch.randint(0, V, (7,), generator=g)
    for chunk in (4096, 3):                          # single chunk (logits kept) and the recompute path
        eng.loss_chunk_rows = chunk
        summary, row_loss, dh = eng.labelled_loss_backward(h, rows, lab)
        ref_loss, ref_dh = R.loss_tail_grads_ref(sd, cfg, h[rows].float(), lab)
        assert abs(float(summary[0]) - float(ref_loss)) < 5e-3
        assert dh.shape == ref_dh.shape and (dh - ref_dh).abs().max() < 2e-2 * ref_dh.abs().max(), chunk
    eng.loss_chunk_rows = 4096
