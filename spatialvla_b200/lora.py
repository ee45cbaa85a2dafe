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


# ============================================================================================================ step layout
class FusedLinear:
    """One GEMM of the engine that carries adapters: a fused projection (q|k|v, interleaved gate/up) holds several PEFT adapters."""

    def __init__(self, name, k_in, n_out, adapters):
        self.name, self.k_in, self.n_out = name, int(k_in), int(n_out)
        self.adapters = adapters                 # [(state_dict key, in_features, out_features, col_start, col_stride)]
        self.R = 32 * len(adapters)              # set properly by the layout (r * adapters)
        self.Rp = 0


class LoRAStepLayout:
    """Every adapter of the fine-tune step (PEFT targets of train/spatialvla_finetune.py:262-270) in ONE flat fp32 arena
    (`param`, `grad`, AdamW `exp_avg` / `exp_avg_sq`: the buffers the single gradient all-reduce and the optimizer kernel see), plus
    the bf16 OPERAND POOL the GEMMs of the step read, rebuilt from the arena by one `svla_lora_pack` launch per step.

    Arena, per adapter: A [r, in] then B^T [r, out] (B is kept transposed so that both gradient reductions of `svla_gemm_tn` land
    in place: gA = (s dY B)^T X is [r, in], gB^T = (s X A^T)^T dY is [r, out]).
    Pool, per fused linear with adapters i = 0.. (R = r * count, Rp = R rounded up to 64, zero padded):
        A_cat  [Rp, k_in]   rows r*i.. = A_i                          W operand of  u = s x A_cat^T          (forward)
        At_cat [k_in, Rp]   = A_cat^T                                  W2 operand of dx = dy W + v At_cat^T   (backward K extension)
        B_blk  [n_out, Rp]  row n of adapter i, cols r*i.. = B_i[n]    W2 operand of  y = x W^T + u B_blk^T   (forward K extension)
        Bt_blk [Rp, n_out]  = B_blk^T                                  W operand of  v = s dy Bt_blk^T        (backward)
    Block structure (zeros outside an adapter's rows / columns) makes the fused q|k|v and gate/up projections one GEMM each.
    Adapters PEFT creates on modules the step never differentiates (ZoeDepth's `out_proj`, frozen under no_grad in the reference,
    model/modeling_spatialvla.py:315-326) keep their arena slots -- the parameter count equals the reference's -- but no operands."""

    def __init__(self, cfg: dict, ops, r: int = 32, alpha: float = 32.0, seed: int = 0, lora_target: str = "linear"):
        from .weights import state_dict_spec
        self.cfg, self.ops, self.r, self.alpha, self.scale = cfg, ops, int(r), float(alpha), float(alpha) / int(r)
        dev = ops.device
        spec = state_dict_spec(cfg)
        self.keys = lora_target_keys(spec, lora_target)                 # every adapter PEFT would create, state_dict order
        shapes = {k: (o, i) for k, o, i in self.keys}
        self.fused = self._fused_linears(cfg, shapes)
        active = [a[0] for fl in self.fused.values() for a in fl.adapters]
        assert len(set(active)) == len(active) and all(k in shapes for k in active)
        order = active + [k for k, _, _ in self.keys if k not in set(active)]
        # ---- arena
        self.slots, off = {}, 0
        for k in order:
            o, i = shapes[k]
            self.slots[k] = (off, off + self.r * i)                     # offsets of A [r, in] and B^T [r, out]
            off += self.r * (i + o)
        self.n = off
        # arena prefix that belongs to the language model: its gradients are complete before the vision backward starts
        gem_keys = [a[0] for nm, fl in self.fused.items() if nm.startswith("gem.") for a in fl.adapters]
        self.n_language = max(self.slots[k][1] + self.r * shapes[k][0] for k in gem_keys) if gem_keys else 0
        z = lambda: torch.zeros(self.n, dtype=torch.float32, device=dev)          # noqa: E731
        self.param, self.grad, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        g = torch.Generator().manual_seed(seed)
        for k in order:                                                 # PEFT "gaussian": A ~ N(0, 1/r) (std 1/r), B = 0
            o, i = shapes[k]
            self.A(k).copy_((torch.randn(self.r, i, generator=g) / self.r).to(dev))
        # ---- operand pool + pack records
        poff, recs = 0, []
        for fl in self.fused.values():
            fl.R = self.r * len(fl.adapters)
            fl.Rp = (fl.R + 63) // 64 * 64
            fl.off = {}
            for nm, numel in (("A_cat", fl.Rp * fl.k_in), ("At_cat", fl.k_in * fl.Rp), ("B_blk", fl.n_out * fl.Rp), ("Bt_blk", fl.Rp * fl.n_out)):
                fl.off[nm] = poff
                poff += (numel + 127) // 128 * 128                      # 256-byte aligned operands
            for ai, (k, fin, fout, c0, cs) in enumerate(fl.adapters):
                a_off, b_off = self.slots[k]
                r0 = self.r * ai
                recs.append((a_off, fl.off["A_cat"] + r0 * fl.k_in, fl.k_in, 1, self.r, fin))
                recs.append((a_off, fl.off["At_cat"] + r0, 1, fl.Rp, self.r, fin))
                recs.append((b_off, fl.off["Bt_blk"] + r0 * fl.n_out + c0, fl.n_out, cs, self.r, fout))
                recs.append((b_off, fl.off["B_blk"] + c0 * fl.Rp + r0, 1, cs * fl.Rp, self.r, fout))
        self.pool = torch.zeros(poff, dtype=torch.bfloat16, device=dev)
        self.records = recs
        tiles, t0 = [], 0
        for (_, _, _, _, rows, cols) in recs:
            tiles.append(t0)
            t0 += ((rows + 31) // 32) * ((cols + 31) // 32)
        self.total_tiles = t0
        import numpy as np
        tab = np.zeros(len(recs), dtype=[("src", "<i8"), ("dst", "<i8"), ("si", "<i8"), ("sj", "<i8"), ("rows", "<i4"), ("cols", "<i4"),
                                         ("tile0", "<i4"), ("pad", "<i4")])
        for n_, ((so, do, si, sj, rows, cols), tl) in enumerate(zip(recs, tiles)):
            tab[n_] = (so, do, si, sj, rows, cols, tl, 0)
        self.desc_table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        for fl in self.fused.values():
            fl.A_cat = self.pool[fl.off["A_cat"]:fl.off["A_cat"] + fl.Rp * fl.k_in].view(fl.Rp, fl.k_in)
            fl.At_cat = self.pool[fl.off["At_cat"]:fl.off["At_cat"] + fl.k_in * fl.Rp].view(fl.k_in, fl.Rp)
            fl.B_blk = self.pool[fl.off["B_blk"]:fl.off["B_blk"] + fl.n_out * fl.Rp].view(fl.n_out, fl.Rp)
            fl.Bt_blk = self.pool[fl.off["Bt_blk"]:fl.off["Bt_blk"] + fl.Rp * fl.n_out].view(fl.Rp, fl.n_out)
            # gradient windows of svla_gemm_tn: (dst view, row0, rows, col_start, col_stride, ncols)
            fl.groups_A = [(self.gA(k), self.r * ai, self.r, 0, 1, fin) for ai, (k, fin, fout, c0, cs) in enumerate(fl.adapters)]
            fl.groups_B = [(self.gBt(k), self.r * ai, self.r, c0, cs, fout) for ai, (k, fin, fout, c0, cs) in enumerate(fl.adapters)]

    # ---- which GEMMs of the engine carry which PEFT adapters
    @staticmethod
    def _fused_linears(cfg, shapes):
        t, v = cfg["text_config"], cfg["vision_config"]
        H, FF = t["hidden_size"], t["intermediate_size"]
        nq, nkv = t["num_attention_heads"] * t["head_dim"], t["num_key_value_heads"] * t["head_dim"]
        D, DI = v["hidden_size"], v["intermediate_size"]
        out = {}

        def add(name, k_in, n_out, parts):
            out[name] = FusedLinear(name, k_in, n_out, [(k, shapes[k][1], shapes[k][0], c0, cs) for k, c0, cs in parts])
        for li in range(t["num_hidden_layers"]):
            p = f"language_model.model.layers.{li}."
            add(f"gem.{li}.qkv", H, nq + 2 * nkv, [(p + "self_attn.q_proj.weight", 0, 1), (p + "self_attn.k_proj.weight", nq, 1),
                                                   (p + "self_attn.v_proj.weight", nq + nkv, 1)])
            add(f"gem.{li}.o", nq, H, [(p + "self_attn.o_proj.weight", 0, 1)])
            add(f"gem.{li}.gu", H, 2 * FF, [(p + "mlp.gate_proj.weight", 0, 2), (p + "mlp.up_proj.weight", 1, 2)])   # interleaved rows
            add(f"gem.{li}.down", FF, H, [(p + "mlp.down_proj.weight", 0, 1)])
        for li in range(v["num_hidden_layers"]):
            p = f"vision_tower.vision_model.encoder.layers.{li}."
            add(f"sig.{li}.qkv", D, 3 * D, [(p + "self_attn.q_proj.weight", 0, 1), (p + "self_attn.k_proj.weight", D, 1),
                                            (p + "self_attn.v_proj.weight", 2 * D, 1)])
            add(f"sig.{li}.o", D, D, [(p + "self_attn.out_proj.weight", 0, 1)])
            add(f"sig.{li}.fc1", D, DI, [(p + "mlp.fc1.weight", 0, 1)])
            add(f"sig.{li}.fc2", DI, D, [(p + "mlp.fc2.weight", 0, 1)])
        add("proj", D, H, [("multi_modal_projector.linear.weight", 0, 1)])
        if cfg.get("use_vision_zoe", True):
            kin = (12 * (2 * cfg["n_freqs"] + 1) + 7) // 8 * 8
            add("ego.0", kin, D, [("position_embedding_3d.position_embedding_head.0.weight", 0, 1)])
            add("ego.3", D, D, [("position_embedding_3d.position_embedding_head.3.weight", 0, 1)])
        return out

    # ---- arena views
    def _shape(self, k):
        for kk, o, i in self.keys:
            if kk == k:
                return o, i
        raise KeyError(k)

    def A(self, k):
        o, i = self._shape(k)
        a, _ = self.slots[k]
        return self.param[a:a + self.r * i].view(self.r, i)

    def Bt(self, k):
        o, i = self._shape(k)
        _, b = self.slots[k]
        return self.param[b:b + self.r * o].view(self.r, o)

    def gA(self, k):
        o, i = self._shape(k)
        a, _ = self.slots[k]
        return self.grad[a:a + self.r * i].view(self.r, i)

    def gBt(self, k):
        o, i = self._shape(k)
        _, b = self.slots[k]
        return self.grad[b:b + self.r * o].view(self.r, o)

    def numel(self):
        return self.n

    def randomize_B(self, std=0.02, seed=1):
        """Tests: PEFT starts B at zero (gA would be identically zero); give every active adapter a non-trivial B."""
        g = torch.Generator().manual_seed(seed)
        for fl in self.fused.values():
            for (k, fin, fout, _, _) in fl.adapters:
                self.Bt(k).copy_((torch.randn(self.r, fout, generator=g) * std).to(self.param.device))

    # ---- per step
    def zero_grad(self):
        self.ops.fill_zero(self.grad)

    def pack(self):
        """fp32 arena -> bf16 operand pool (one launch; the structural zeros of the pool were written once at construction)."""
        self.ops.lora_pack(self.param, self.pool, self)

    # ---- interchange
    def delta(self, k):
        return self.scale * (self.Bt(k).float().t() @ self.A(k).float())

    def merged_state_dict(self, sd):
        """Base state_dict with every adapter folded in (W + (alpha / r) B A): what PEFT's merge_and_unload produces; the inference
        engine (and the oracle) consume it unchanged."""
        out = dict(sd)
        for k, _, _ in self.keys:
            out[k] = (sd[k].float() + self.delta(k).to(sd[k].device)).to(sd[k].dtype)
        return out

    def peft_state_dict(self):
        """PEFT-style adapter tensors: `<module>.lora_A.default.weight` [r, in], `<module>.lora_B.default.weight` [out, r]."""
        out = {}
        for k, _, _ in self.keys:
            mod = "base_model.model." + k[: -len(".weight")]
            out[mod + ".lora_A.default.weight"] = self.A(k).detach().clone()
            out[mod + ".lora_B.default.weight"] = self.Bt(k).detach().t().contiguous()
        return out
