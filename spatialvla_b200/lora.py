"""LoRA parameter arena for the fine-tune step of BASELINE.json config #5.

The reference wraps the model with PEFT (`train/spatialvla_finetune.py:262-302`: r = 32, alpha = 32, `init_lora_weights="gaussian"`,
targets by module-name suffix) and lets DeepSpeed ZeRO-1 reduce the adapter gradients (`scripts/zero1.json`).  Here every adapter
lives in ONE flat contiguous buffer (parameters) with a same-shaped gradient buffer, so that the data-parallel step needs exactly
one collective: `parallel.allreduce_gradients(arena)` = a single NCCL all-reduce over `arena.grad` (59.2 M elements at the 4B-224
size: 118 MB in bf16, 237 MB in fp32).  The forward of an adapted Linear is W x + (alpha / r) B (A x); for inference the adapters
are folded into the base weights (`merged_state_dict`) and the unchanged engine runs them.  The backward kernels that will fill
`arena.grad` are next round's work; the chain rule they implement (`accumulate_from_weight_grad`) is checked against autograd.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

# train/spatialvla_finetune.py:264-270 (lora_target == "linear"); "+emb" adds spatial_embed_tokens, "+h" adds lm_head (:271-287)
LORA_TARGETS = {
    "linear": ("q_proj", "o_proj", "k_proj", "v_proj", "gate_proj", "up_proj", "down_proj", "fc1", "fc2", "out_proj", "linear",
               "position_embedding_head.0", "position_embedding_head.3"),
}
LORA_TARGETS["linear+emb"] = LORA_TARGETS["linear"] + ("spatial_embed_tokens",)
LORA_TARGETS["linear+emb+h"] = LORA_TARGETS["linear"] + ("lm_head", "spatial_embed_tokens")


def lora_target_keys(spec: Dict[str, tuple], lora_target: str = "linear") -> List[Tuple[str, int, int]]:
    """[(state_dict weight key, out_features, in_features)] of the modules PEFT adapts: 2-D `.weight` tensors whose module name
    equals a target or ends with '.' + target (peft/tuners/tuners_utils.py suffix rule), in state_dict order."""
    if lora_target not in LORA_TARGETS:
        raise ValueError(f"don't support lora targets {lora_target}")
    targets = LORA_TARGETS[lora_target]
    out = []
    for k, shape in spec.items():
        if not k.endswith(".weight") or len(shape) != 2:
            continue
        mod = k[: -len(".weight")]
        if any(mod == t or mod.endswith("." + t) for t in targets):
            out.append((k, int(shape[0]), int(shape[1])))
    return out


def lora_numel(spec: Dict[str, tuple], r: int = 32, lora_target: str = "linear") -> int:
    return sum(r * (o + i) for _, o, i in lora_target_keys(spec, lora_target))


class LoRAArena:
    """All adapters of a model in one flat buffer.  `A[key]` is the (r, in) view, `B[key]` the (out, r) view of the base weight
    `key`; `gA` / `gB` are the matching views of the gradient buffer.  PEFT's gaussian init: A ~ N(0, 1/r), B = 0."""

    def __init__(self, spec: Dict[str, tuple], r: int = 32, alpha: float = 32.0, lora_target: str = "linear", device="cpu",
                 dtype=torch.float32, seed: int = 0):
        self.r, self.alpha, self.scale = int(r), float(alpha), float(alpha) / int(r)
        self.keys = lora_target_keys(spec, lora_target)
        n = sum(self.r * (o + i) for _, o, i in self.keys)
        self.param = torch.zeros(n, dtype=dtype, device=device)
        self.grad = torch.zeros(n, dtype=dtype, device=device)
        self.A, self.B, self.gA, self.gB = {}, {}, {}, {}
        g = torch.Generator().manual_seed(seed)
        off = 0
        for k, o, i in self.keys:
            na, nb = self.r * i, o * self.r
            self.A[k], self.gA[k] = self.param[off:off + na].view(self.r, i), self.grad[off:off + na].view(self.r, i)
            self.A[k].copy_((torch.randn(self.r, i, generator=g) / self.r).to(dtype))      # std 1/r (peft LoraLayer.reset, "gaussian")
            off += na
            self.B[k], self.gB[k] = self.param[off:off + nb].view(o, self.r), self.grad[off:off + nb].view(o, self.r)
            off += nb
        assert off == n

    def numel(self) -> int:
        return self.param.numel()

    def zero_grad(self):
        self.grad.zero_()

    def delta(self, key: str) -> torch.Tensor:
        """(alpha / r) * B @ A in fp32: what the adapter adds to the base weight `key`."""
        return self.scale * (self.B[key].float() @ self.A[key].float())

    def merged_state_dict(self, sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Base state_dict with every adapter folded in (W + (alpha/r) B A, PEFT merge_and_unload): feed it to
        SpatialVLAForConditionalGeneration for inference with the fine-tuned weights."""
        out = dict(sd)
        for k, _, _ in self.keys:
            out[k] = (sd[k].float() + self.delta(k).to(sd[k].device)).to(sd[k].dtype)
        return out

    def accumulate_from_weight_grad(self, key: str, dW: torch.Tensor):
        """Chain rule of W' = W + s B A for a full-weight gradient dW = dL/dW':  gB += s dW A^T,  gA += s B^T dW.
        Host-side torch helper for tests and one-off conversions only: the training step never forms dW -- its kernels will write
        gA / gB of this arena directly (oracle/backward_ref.lora_linear_bwd is the formula they are checked against)."""
        dW = dW.float()
        self.gB[key] += (self.scale * (dW @ self.A[key].float().t())).to(self.grad.dtype)
        self.gA[key] += (self.scale * (self.B[key].float().t() @ dW)).to(self.grad.dtype)
