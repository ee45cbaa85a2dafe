"""`SpatialVLAForConditionalGeneration` -- the reference's model-side API (model/modeling_spatialvla.py:162-526)
over the B200 engine: `predict_action`, `forward` (logits), `get_image_features`, `backproject_patch`,
`from_pretrained`.  There is no torch.nn module tree and no CPU path: weights are repacked once into kernel layouts
(spatialvla_b200/engine.py) and every FLOP runs in libspatialvla_b200.so.

Differences from the reference, all deliberate and documented in DESIGN.md:
  * `predict_action` decodes a fixed number of action tokens (3 per action x action_chunk_size, default 12) with the
    argmax restricted to the spatial-action slice of the vocabulary instead of HF `generate(max_new_tokens=256)` +
    EOS stop (north_star: "argmax restricted to the action vocabulary"); `max_new_tokens` can be passed.
  * inputs stay fp32 (the reference rounds pixel_values *and* the intrinsic matrix to bf16 at :489).
  * LEFT-padded batches (attention_mask = 0...01...1, the Gemma tokenizer's padding side) are supported with the reference's
    semantics (padded key columns masked, positions restart on the first real token); any other mask pattern raises
    NotImplementedError instead of silently mis-computing.
  * `forward(labels=..., token_type_ids=...)` -- the forward half of the training step -- computes the
    reference's loss with its training masks (triangular, or prefix-LM when a 2-D attention_mask is passed) but materialises
    logits only for the labelled rows ([R, V] instead of [B, L, V]); padded batches raise NotImplementedError on this path
    (the reference un-masks padded prefix columns there, :304-305).  There is no backward pass yet.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from .configuration_spatialvla import SpatialVLAConfig
from .engine import SpatialVLAEngine

BF16, F32 = torch.bfloat16, torch.float32


@dataclass
class SpatialVLACausalLMOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: torch.Tensor = None
    past_key_values: Optional[dict] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None
    image_hidden_states: Optional[torch.Tensor] = None
    # labelled forward only (this build materialises logits for the labelled rows, not [B, L, V]):
    label_rows: Optional[torch.Tensor] = None        # int64 [R] flat positions b*L + t whose next token is labelled
    row_loss: Optional[torch.Tensor] = None          # fp32 [R] per-row cross entropy
    row_labels: Optional[torch.Tensor] = None        # int64 [R] the labels of those rows (shifted)
    row_argmax: Optional[torch.Tensor] = None        # int64 [R] full-vocabulary argmax of those rows
    token_accuracy: Optional[torch.Tensor] = None    # 0-dim fp32: argmax == label over the labelled rows


class SpatialVLAForConditionalGeneration:
    config_class = SpatialVLAConfig

    def __init__(self, config, state_dict=None, device="cuda:0", ops=None, action_chunk_size: int = 4):
        """config: SpatialVLAConfig or plain engine dict. state_dict: HF-keyed tensors (any float dtype); when None,
        deterministic synthetic weights are generated (tests / bench: there is no network for checkpoints)."""
        if isinstance(config, SpatialVLAConfig):
            self.config, cfg = config, config.to_engine_dict()
        else:
            self.config, cfg = None, dict(config)
        self.engine_config = cfg
        if ops is None:
            from .ops import CudaOps          # raises when the CUDA extension / device is missing -- no fallback
            ops = CudaOps(device)
        self.ops = ops
        self.device = ops.device
        self.dtype = BF16
        if state_dict is None:
            from .weights import synth_state_dict
            state_dict = synth_state_dict(cfg, seed=0)
        self.engine = SpatialVLAEngine(cfg, state_dict, ops)
        self.action_chunk_size = action_chunk_size
        self.vocab_size = cfg["text_config"]["vocab_size"]

    # ------------------------------------------------------------------------------------------ construction
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, config=None, device="cuda:0", **kwargs):
        """Loads `config.json` + `*.safetensors` from a local directory written by the reference
        (model/modeling_spatialvla.py:494-526).  Like the reference, the last `spatial_token_num` rows of
        `embed_tokens` are overwritten with `spatial_embed_tokens` (:524-525)."""
        path = str(pretrained_model_name_or_path)
        if not os.path.isdir(path):
            raise OSError(f"{path} is not a local directory (this build has no hub access)")
        if config is None:
            config = SpatialVLAConfig.from_pretrained(path)
        from safetensors.torch import load_file
        sd = {}
        for fn in sorted(os.listdir(path)):
            if fn.endswith(".safetensors"):
                sd.update(load_file(os.path.join(path, fn)))
        if not sd:
            raise OSError(f"no *.safetensors under {path}")
        if config.use_spatial_token:
            n = config.spatial_token_num
            sd["language_model.model.embed_tokens.weight"][-n:] = sd["spatial_embed_tokens.weight"]
        return cls(config, sd, device=device, **kwargs)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def cuda(self):
        return self

    # ------------------------------------------------------------------------------------------ hot path
    def _prepare(self, model_inputs):
        ids = model_inputs["input_ids"]
        px = model_inputs.get("pixel_values") if hasattr(model_inputs, "get") else model_inputs["pixel_values"]
        K = model_inputs.get("intrinsic") if hasattr(model_inputs, "get") else model_inputs["intrinsic"]
        am = model_inputs.get("attention_mask") if hasattr(model_inputs, "get") else None
        pads = None
        if am is not None and am.dim() == 2 and bool((am == 0).any()):
            pads = self._left_pads(am, ids.shape[1])
        self.raise_if_bad_batch()                   # deferred check of the previous device-resident batch (no sync on this one)
        if px is not None and not ids.is_cuda:
            # host ids (what the processor returns): audit EVERY row before anything is launched (:379-385)
            n_img = (ids == self.engine_config["image_token_index"]).sum(1)
            if bool((n_img != 256).any()) or ids.shape[0] != px.shape[0]:
                raise ValueError(
                    "Number of images does not match number of special image tokens in the input text. "
                    f"Got {int(n_img.sum())} image tokens in the text but {px.shape[0] * 256} tokens from image embeddings.")
        ids = ids.to(self.device, torch.int64).contiguous()
        if px is not None:
            px = px.to(self.device, F32).contiguous()
            if ids.shape[0] != px.shape[0]:
                raise ValueError("Number of images does not match number of special image tokens in the input text. "
                                 f"Got {ids.shape[0]} prompt rows but {px.shape[0]} images.")
        if K is not None:
            K = K.to(self.device, F32).contiguous()
        return ids, px, K, pads

    def raise_if_bad_batch(self):
        """Device-resident `input_ids` are audited by the embedding kernel (every row must hold exactly 256 image tokens and only
        valid ids); its status word is read here -- after the batch has been decoded, or at the start of the next call -- so the
        hot path itself carries no device->host synchronisation.  Raises the reference's ValueError (:379-385)."""
        st = getattr(self.engine, "last_status", None)
        if st is None:
            return
        self.engine.last_status = None
        code = int(st.item())
        if code == 1:
            raise ValueError("Number of images does not match number of special image tokens in the input text.")
        if code == 2:
            raise ValueError("input_ids outside [0, vocab_size)")

    def _left_pads(self, attention_mask, P):
        """(B,P) 0/1 mask with zeros -> int32 [B] device tensor of leading pad counts.  Only left padding is supported: every
        row must be zeros followed by ones with at least one real token (model/modeling_spatialvla.py:298-303 masks the padded
        key columns; HF generate restarts the position ids on the first real token, model/modeling_gemma2.py:1042-1051)."""
        am = attention_mask.to(torch.int64)
        pads = (am == 0).sum(1)
        expect = (torch.arange(P, device=am.device)[None, :] >= pads[:, None]).to(torch.int64)
        if am.shape[1] != P or not bool(torch.equal(am, expect)) or int(pads.max()) >= P:
            raise NotImplementedError("only left-padded batches (attention_mask = 0...01...1) are supported on this path")
        return pads.to(device=self.device, dtype=torch.int32).contiguous()

    @torch.no_grad()
    def generate(self, model_inputs=None, max_new_tokens: int = 256, do_sample: bool = False, eos_token_id=None, pad_token_id=None,
                 **inputs):
        """HF-`generate` semantics of the reference's predict_action (model/modeling_spatialvla.py:484-492): greedy, argmax over the
        FULL vocabulary, stop at EOS (finished rows padded), at most `max_new_tokens` new tokens.  Returns prompt + generated ids
        like `GenerationMixin.generate`; `predict_action(..., reference_generate=True)` strips the prompt as the reference does."""
        if do_sample:
            raise NotImplementedError("sampling is not part of the reference's predict_action path (do_sample=False at :491)")
        model_inputs = dict(model_inputs) if model_inputs is not None else {}
        model_inputs.update(inputs)
        ids, px, K, pads = self._prepare(model_inputs)
        eos = eos_token_id if eos_token_id is not None else self.engine_config.get("eos_token_id")
        pad = pad_token_id if pad_token_id is not None else self.engine_config.get("pad_token_id")
        eos = 1 if eos is None else int(eos)
        pad = eos if pad is None else int(pad)                      # HF falls back to the EOS id when no pad id is configured
        new = self.engine.generate_reference(ids, px, K, int(max_new_tokens), eos, pad, pads=pads)
        self.raise_if_bad_batch()
        return torch.cat([ids, new], 1)

    @torch.no_grad()
    def predict_action(self, model_inputs, max_new_tokens: Optional[int] = None, return_logits: bool = False,
                       reference_generate: bool = False):
        """model_inputs: dict-like with input_ids (B,P), pixel_values (B,3,224,224) in [0,1], intrinsic (3,3)|(B,3,3)
        -> LongTensor (B, n_new) of generated action-token ids on the model device (prompt stripped, :492).
        reference_generate=True: the reference's exact decoding rule instead (full-vocabulary argmax, EOS stop, up to 256 tokens)."""
        if reference_generate:
            P = model_inputs["input_ids"].shape[1]
            return self.generate(model_inputs, max_new_tokens=256 if max_new_tokens is None else int(max_new_tokens))[:, P:]
        ids, px, K, pads = self._prepare(model_inputs)
        n_new = int(max_new_tokens) if max_new_tokens is not None else 3 * self.action_chunk_size
        out = self.engine.generate_actions(ids, px, K, n_new, return_logits=return_logits, pads=pads)
        return out

    @torch.no_grad()
    def get_image_features(self, pixel_values: torch.Tensor, intrinsic: torch.Tensor):
        """(B,3,224,224) in [0,1], K -> (B, 256, H_text) fp32, already / sqrt(H) (model/modeling_spatialvla.py:308-333)"""
        return self.engine.image_features(pixel_values.to(self.device, F32).contiguous(), intrinsic.to(self.device, F32))

    @torch.no_grad()
    def predict_depth(self, pixel_values: torch.Tensor):
        """ZoeDepth metric depth at 384x384 for [0,1] 224x224 images (the tensor :317 produces)."""
        return self.engine.zoedepth(pixel_values.to(self.device, F32).contiguous())

    @torch.no_grad()
    def backproject_patch(self, K: torch.Tensor, depth384: torch.Tensor, patch_size=14, reso=2) -> torch.Tensor:
        """Fused kernel for model/modeling_spatialvla.py:318-323 + :195-223: takes the 384x384 ZoeDepth output
        (B,384,384), resamples/crops to 224, area-pools and back-projects -> (B, 256, 12) fp32."""
        if patch_size != 14 or reso != 2:
            raise NotImplementedError("the fused Ego3D kernel is specialised for patch 14 / reso 2 (4B-224 config)")
        B = depth384.shape[0]
        xyz = self.ops.empty((B * 256, 12), F32)
        enc = self.ops.empty((B * 256, self.engine.ego_kpad), BF16)
        self.ops.ego3d_encode(depth384.to(self.device, F32).contiguous(), K.to(self.device, F32).contiguous(), xyz, enc,
                              n_freqs=self.engine_config["n_freqs"])
        return xyz.view(B, 256, 12)

    @torch.no_grad()
    def forward(self, input_ids=None, pixel_values=None, actions=None, intrinsic=None, attention_mask=None,
                position_ids=None, past_key_values=None, token_type_ids=None, cache_position=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                num_logits_to_keep: int = 0):
        """Inference forward (model/modeling_spatialvla.py:335-442): full-vocabulary post-softcap logits (fp32) for the
        last `num_logits_to_keep` positions (0 = all). `past_key_values` is this engine's cache dict."""
        if labels is not None:
            if past_key_values is not None or inputs_embeds is not None or position_ids is not None or output_attentions or output_hidden_states:
                raise NotImplementedError("labelled forward: past_key_values / inputs_embeds / position_ids / output_* are not supported")
            return self._forward_with_labels(input_ids, pixel_values, intrinsic, attention_mask, token_type_ids, labels)
        # token_type_ids without labels: the reference ignores them (is_training is False, :357)
        if inputs_embeds is not None or position_ids is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("inputs_embeds / position_ids / output_attentions / output_hidden_states")
        eng = self.engine
        ids, px, K, pads = self._prepare({"input_ids": input_ids, "pixel_values": pixel_values, "intrinsic": intrinsic,
                                          "attention_mask": attention_mask if (attention_mask is not None and attention_mask.dim() == 2
                                                                               and attention_mask.shape[1] == input_ids.shape[1]) else None})
        B, S = ids.shape
        feats = eng.image_features(px, K) if px is not None else None
        x, status = eng.embed(ids, feats)
        cache = past_key_values if past_key_values is not None else eng.new_cache(B, S + 256)
        prefill = cache["len"] == 0
        if prefill:
            cache["pads"] = pads             # decode calls on this cache keep masking the prompt's padding slots
        h = eng.gemma_forward(x, B, S, cache, bidirectional=prefill, pads=cache.get("pads"))
        H = h.shape[-1]
        keep = S if num_logits_to_keep in (0, None) else int(num_logits_to_keep)
        rows = h.view(B, S, H)[:, S - keep:].reshape(B * keep, H)
        V = self.vocab_size
        cap = eng.t["final_logit_softcapping"]
        logits = self.ops.empty((B * keep, V), F32)
        from ._lib import ACT_NONE, ACT_SOFTCAP
        self.ops.gemm(rows, eng.lm_head_full(), out_f32=logits, act=ACT_SOFTCAP if cap else ACT_NONE, act_param=cap or 0.0)
        if int(status.item()) == 1:
            raise ValueError("Number of images does not match number of special image tokens in the input text.")
        return SpatialVLACausalLMOutputWithPast(logits=logits.view(B, keep, V), past_key_values=cache,
                                                image_hidden_states=feats)

    def _forward_with_labels(self, input_ids, pixel_values, intrinsic, attention_mask, token_type_ids, labels):
        """Loss forward = model/modeling_spatialvla.py:335-430 with labels.  Mask (`_update_causal_mask`, :258-306):
        token_type_ids given (is_training) -> triangular, plus every token_type 0 column unmasked when a 2-D attention_mask
        is passed (prefix-LM); labels alone -> the bidirectional inference mask.  Loss: shifted nn.CrossEntropyLoss over the
        full vocabulary on the post-softcap logits, ignore_index rows dropped, pad-token labels masked (:389-397)."""
        eng = self.engine
        ignore = self.engine_config.get("ignore_index", -100)
        ignore = -100 if ignore is None else ignore
        pad_id = self.engine_config.get("pad_token_id")
        pad_id = -1 if pad_id is None else pad_id
        B, L = input_ids.shape
        if tuple(labels.shape) != (B, L):
            raise ValueError(f"labels shape {tuple(labels.shape)} != input_ids shape {(B, L)}")
        if attention_mask is not None and (attention_mask.dim() != 2 or bool((attention_mask == 0).any())):
            raise NotImplementedError("labelled forward: only unpadded batches (attention_mask None or all ones)")
        bidirectional, prefix = True, 0
        if token_type_ids is not None:
            bidirectional = False
            if attention_mask is not None:
                tt = token_type_ids.to("cpu", torch.int64)
                if tuple(tt.shape) != (B, L):
                    raise ValueError("token_type_ids shape != input_ids shape")
                prefix = int((tt[0] == 0).sum())
                if not bool(torch.equal(tt, (torch.arange(L)[None, :] >= prefix).to(torch.int64).expand(B, L))):
                    raise NotImplementedError("token_type_ids must be 0...01...1 with the same prefix length in every row")
        # label bookkeeping on the host (labels arrive from the data collator on the CPU): shifted positions that carry a label
        lab = labels.to("cpu", torch.int64)
        ids_cpu = input_ids.to("cpu", torch.int64)
        if bool((lab == pad_id).any()):
            lab = torch.where(ids_cpu == pad_id, torch.full_like(lab, ignore), lab)
        shift = lab[:, 1:]
        bi, ti = torch.nonzero(shift != ignore, as_tuple=True)
        row_labels = shift[bi, ti]
        if row_labels.numel() and (int(row_labels.min()) < 0 or int(row_labels.max()) >= self.vocab_size):
            raise ValueError("labels outside [0, vocab_size)")
        ids, px, K, _ = self._prepare({"input_ids": input_ids, "pixel_values": pixel_values, "intrinsic": intrinsic})
        feats = eng.image_features(px, K) if px is not None else None
        x, status = eng.embed(ids, feats)
        cache = eng.new_cache(B, L)
        h = eng.gemma_forward(x, B, L, cache, bidirectional=bidirectional, causal_prefix=prefix)
        if int(status.item()) == 1:
            raise ValueError("Number of images does not match number of special image tokens in the input text.")
        rows = (bi * L + ti).to(self.device)
        if rows.numel() == 0:                     # nn.CrossEntropyLoss over zero rows is NaN
            nan = torch.full((), float("nan"), dtype=F32, device=self.device)
            return SpatialVLACausalLMOutputWithPast(loss=nan, logits=None, image_hidden_states=feats, label_rows=rows)
        row_labels = row_labels.to(self.device).contiguous()
        summary, row_loss, row_argmax, logits = eng.labelled_loss(h, rows, row_labels, ignore_index=ignore)
        return SpatialVLACausalLMOutputWithPast(loss=summary[0], logits=logits, image_hidden_states=feats, label_rows=rows,
                                                row_loss=row_loss, row_labels=row_labels, row_argmax=row_argmax,
                                                token_accuracy=summary[2] / summary[1])

    @staticmethod
    def action_metrics(outputs, actions, action_tokenizer):
        """The metric block of the reference's training step (train/monkey_patch.py:267-324) from a labelled forward:
        accuracy of the full-vocabulary argmax on the positions whose label is an action token, the same per token group
        (translation / rotation / gripper) and the L1 distance between the de-tokenised predictions and `actions`.
        outputs: result of forward(labels=...); actions: (..., 7) ground-truth (normalised) actions, one per labelled action
        triple; action_tokenizer: SpatialActionTokenizer (sub-tokenizer id ranges + decode).  Returns a dict of floats."""
        gt, pred = outputs.row_labels, outputs.row_argmax
        tk = action_tokenizer
        lo, hi = tk.translation_tokenizer.token_start_idx, tk.gripper_tokenizer.token_end_idx
        mask = (gt >= lo) & (gt <= hi)
        gt, pred = gt[mask], pred[mask]
        correct = gt == pred
        out = {"accuracy": correct.sum().float() / mask.sum().float()}
        for name, sub_tk in (("translation", tk.translation_tokenizer), ("rotation", tk.rotation_tokenizer),
                             ("gripper", tk.gripper_tokenizer)):
            m = (gt >= sub_tk.token_start_idx) & (gt <= sub_tk.token_end_idx)
            out[name + "_accuracy"] = correct[m].sum().float() / m.sum().float()
        ids = pred.reshape(-1, 3)
        if ids.is_cuda and hasattr(tk, "decode_ids"):
            pred_actions = tk.decode_ids(ids).to(F32)                                  # device kernel, no host round trip
        else:
            pred_actions = torch.as_tensor(tk.decode_token_ids_to_actions(ids.cpu().numpy())).to(F32)
        gt_actions = torch.as_tensor(actions).reshape(-1, 7).to(device=pred_actions.device, dtype=F32)
        out["l1_loss"] = torch.nn.functional.l1_loss(pred_actions, gt_actions)
        return {k: float(v) for k, v in out.items()}

    __call__ = forward
