"""The LoRA fine-tune step of BASELINE.json config #5 over the C-ABI kernels: forward with un-merged
adapters, full backward, gradient all-reduce over the flat arena, gradient-norm clipping and AdamW.

Reference: `train/spatialvla_finetune.py:262-302` (PEFT LoRA r = 32, alpha = 32 on every Linear of Gemma2 / SigLIP / projector /
Ego3D head; ZoeDepth frozen and under no_grad, `model/modeling_spatialvla.py:315-326`), `model/modeling_spatialvla.py:335-430`
(labelled forward, shifted cross entropy), `train/monkey_patch.py:222-326` + HF Trainer (step: backward, clip_grad_norm 1.0,
AdamW betas 0.9 / 0.999, eps 1e-8, weight decay 0), `scripts/zero1.json` (DeepSpeed ZeRO-1 gradient reduction).

B200 formulation (no autograd, no recomputation, no per-layer Python modules):
  * an adapted Linear is ONE tcgen05 GEMM whose K loop is extended by the rank-r term: y = x W^T + u B^T with u = s x A^T
    (`svla_gemm` K extension); its input gradient is the same kernel on the cached transposed weight: dx = dy W + v A, v = s dy B;
    the base dW is never formed; gA = v^T x and gB^T = u^T dy are token-dimension reductions (`svla_gemm_tn`) that accumulate
    straight into the fp32 gradient arena;
  * every activation the backward needs is kept (B = 32 per GPU: ~32 GB of the 180 GB HBM) -- the reference recomputes them
    (gradient checkpointing) because it has to;
  * attention backward in the flash formulation (`svla_attention_bwd`), norms / GeGLU / GELU / RoPE backward as fused memory-bound
    kernels, residual-stream gradients in fp32;
  * ONE collective per step: all-reduce of the 59.2 M-element gradient arena (NCCL over NVLink), then clip + AdamW in one kernel.
"""
from __future__ import annotations

import torch

from ._lib import ACT_NONE, ACT_SOFTCAP
from .lora import LoRAStepLayout

BF16, F32 = torch.bfloat16, torch.float32


class LoRATrainer:
    def __init__(self, engine, r: int = 32, alpha: float = 32.0, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: float = 1.0, seed: int = 0):
        self.eng, self.ops = engine, engine.ops
        self.lay = LoRAStepLayout(engine.cfg, engine.ops, r=r, alpha=alpha, seed=seed)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.step_count = 0
        self._transpose_weights()
        self.norm_sq = torch.zeros(1, dtype=F32, device=self.ops.device)
        self.timers = None

    # ------------------------------------------------------------------------------------------ one-off: W^T for the dX GEMMs
    def _transpose_weights(self):
        """dx = dy W needs W as the [N' = in, K' = out] operand of the K-major GEMM: one transposed bf16 copy per adapted weight,
        made once (5.2 GB at the 4B size; the base weights never change during LoRA fine-tuning)."""
        e = self.eng
        t = lambda w: w.t().contiguous()                     # noqa: E731  init-time plumbing
        self.gem_t = [{"wqkv": t(L["wqkv"]), "wo": t(L["wo"]), "wgu": t(L["wgu"]), "wd": t(L["wd"])} for L in e.gem["layers"]]
        self.sig_t = [{"wqkv": t(L["wqkv"]), "wo": t(L["wo"]), "w1": t(L["w1"]), "w2": t(L["w2"])} for L in e.sig["layers"]]
        self.proj_t = t(e.proj_w)
        self.ego3_t = t(e.ego["w3"]) if e.use_zoe else None

    # ------------------------------------------------------------------------------------------ adapted Linear
    def _lin_fwd(self, x, W, fl, rows, *, out_dtype=BF16, **kw):
        """y = epilogue(x W^T + u B_blk^T), u = s x A_cat^T.  Returns (y, u); u is kept for gB."""
        ops = self.ops
        u = ops.empty((rows, fl.Rp), BF16)
        ops.gemm(x, fl.A_cat, out_bf16=u, alpha=self.lay.scale)
        y = ops.empty((rows, fl.n_out), out_dtype)
        if out_dtype == BF16:
            ops.gemm(x, W, out_bf16=y, a2=u, w2=fl.B_blk, **kw)
        else:
            ops.gemm(x, W, out_f32=y, a2=u, w2=fl.B_blk, **kw)
        return y, u

    def _lin_bwd(self, dy, Wt, fl, x, u, rows, need_dx=True):
        """dx = dy W + v A_cat (bf16), v = s dy B_blk;  gA += v^T x,  gB^T += u^T dy  (into the gradient arena)."""
        ops = self.ops
        v = ops.empty((rows, fl.Rp), BF16)
        ops.gemm(dy, fl.Bt_blk, out_bf16=v, alpha=self.lay.scale)
        dx = None
        if need_dx:
            dx = ops.empty((rows, fl.k_in), BF16)
            ops.gemm(dy, Wt, out_bf16=dx, a2=v, w2=fl.At_cat)
        ops.gemm_tn(v, x, fl.groups_A, r=fl.R, n=fl.k_in)
        ops.gemm_tn(u, dy, fl.groups_B, r=fl.R, n=fl.n_out)
        return dx

    # ------------------------------------------------------------------------------------------ forward (keeps activations)
    def _siglip_fwd(self, px, B):
        ops, e, s, v, F_ = self.ops, self.eng, self.eng.sig, self.eng.v, self.lay.fused
        D, nh = v["hidden_size"], v["num_attention_heads"]
        S, M = 256, B * 256
        eps = v.get("layer_norm_eps", 1e-6)
        a = ops.empty((M, e.sig_kpad), BF16)
        ops.siglip_patchify(px, a)
        x = ops.empty((M, D), F32)
        ops.gemm(a, s["patch_w"], bias=s["patch_b"], res_f32=s["pos"], res_mod=S, out_f32=x)
        saved = []
        hd = D // nh
        st = (S * 3 * D, 3 * D)
        for li, L_ in enumerate(s["layers"]):
            c = {"x_in": x}
            c["h1"] = ops.empty((M, D), BF16)
            ops.layernorm(x, L_["ln1_g"], L_["ln1_b"], eps, out_bf16=c["h1"])
            c["qkv"], c["u_qkv"] = self._lin_fwd(c["h1"], L_["wqkv"], F_[f"sig.{li}.qkv"], M, bias=L_["bqkv"])
            c["ctx"] = ops.empty((M, D), BF16)
            c["lse"] = ops.empty((B, nh, (S + 63) // 64 * 64), F32)      # row log-sum-exp of the forward, read by the backward sweeps
            qkv = c["qkv"]
            ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], c["ctx"], batch=B, hq=nh, hkv=nh, sq=S, sk=S, d=hd, q_strides=st, k_strides=st,
                          v_strides=st, o_strides=(S * D, D), scale=hd ** -0.5, lse=c["lse"])
            c["x_mid"], c["u_o"] = self._lin_fwd(c["ctx"], L_["wo"], F_[f"sig.{li}.o"], M, out_dtype=F32, bias=L_["bo"], res_f32=x)
            c["h2"] = ops.empty((M, D), BF16)
            ops.layernorm(c["x_mid"], L_["ln2_g"], L_["ln2_b"], eps, out_bf16=c["h2"])
            c["z"], c["u_fc1"] = self._lin_fwd(c["h2"], L_["w1"], F_[f"sig.{li}.fc1"], M, bias=L_["b1"])
            c["f"] = ops.empty(c["z"].shape, BF16)
            ops.gelu_tanh_fwd(c["z"], c["f"])
            x, c["u_fc2"] = self._lin_fwd(c["f"], L_["w2"], F_[f"sig.{li}.fc2"], M, out_dtype=F32, bias=L_["b2"], res_f32=c["x_mid"])
            saved.append(c)
        sig = ops.empty((M, D), F32)
        ops.layernorm(x, s["post_g"], s["post_b"], eps, out_f32=sig)
        return sig, {"layers": saved, "x_last": x}

    def _vision_fwd(self, px, K, B):
        """-> (image features fp32 [B, 256, H] already / sqrt(H), saved activations).  ZoeDepth runs the inference engine's own
        kernels without keeping anything: it is frozen and under no_grad in the reference (model/modeling_spatialvla.py:315-326)."""
        ops, e, F_ = self.ops, self.eng, self.lay.fused
        D, H = e.v["hidden_size"], e.t["hidden_size"]
        M = B * 256
        sig, sv = self._siglip_fwd(px, B)
        if e.use_zoe:
            depth = e.zoedepth(px)
            xyz = ops.empty((M, 12), F32)
            sv["enc"] = ops.empty((M, e.ego_kpad), BF16)
            ops.ego3d_encode(depth, K, xyz, sv["enc"], n_freqs=e.cfg["n_freqs"])
            sv["h0"], sv["u_e0"] = self._lin_fwd(sv["enc"], e.ego["w0"], F_["ego.0"], M, out_dtype=F32, bias=e.ego["b0"])
            sv["hb"] = ops.empty((M, D), BF16)
            ops.layernorm(sv["h0"], e.ego["ln_g"], e.ego["ln_b"], 1e-5, out_bf16=sv["hb"], relu=True)
            sv["src"], sv["u_e3"] = self._lin_fwd(sv["hb"], e.ego["w3"], F_["ego.3"], M, bias=e.ego["b3"], res_f32=sig)
        else:
            sv["src"] = ops.empty((M, D), BF16)
            ops.rows_cast(sig, sv["src"])
        feats, sv["u_proj"] = self._lin_fwd(sv["src"], e.proj_w, F_["proj"], M, out_dtype=F32, bias=e.proj_b, colscale=e.proj_scale)
        return feats.view(B, 256, H), sv

    def _gemma_fwd(self, x, B, S, causal, prefix):
        ops, e, g, t, F_ = self.ops, self.eng, self.eng.gem, self.eng.t, self.lay.fused
        H, nh, nkv, hd, FF = t["hidden_size"], t["num_attention_heads"], t["num_key_value_heads"], t["head_dim"], t["intermediate_size"]
        eps, theta = t["rms_norm_eps"], float(t.get("rope_theta", 10000.0))
        scale, cap = t["query_pre_attn_scalar"] ** -0.5, t["attn_logit_softcapping"] or 0.0
        M = B * S
        # sliding-window layers (even layer_idx, model/modeling_gemma2.py:343,441-473): the predicate is passed to the forward and
        # backward attention kernels once the sequence exceeds the window (never at the 291 tokens of config #5 against 4096)
        win = int(t.get("sliding_window") or 0)
        win = win if S > win else 0
        nL = len(g["layers"])
        kc_all, vc_all = ops.empty((nL, B, S, nkv, hd), BF16), ops.empty((nL, B, S, nkv, hd), BF16)
        h1 = ops.empty((M, H), BF16)
        ops.rmsnorm_train_fwd(x, w_pre=g["layers"][0]["ln_in"], eps=eps, h=h1)
        saved = []
        kvs = (S * nkv * hd, nkv * hd)
        for li, L_ in enumerate(g["layers"]):
            c = {"x_in": x, "h1": h1, "kc": kc_all[li], "vc": vc_all[li]}
            qkv, c["u_qkv"] = self._lin_fwd(h1, L_["wqkv"], F_[f"gem.{li}.qkv"], M)
            c["q"] = ops.empty((M, nh * hd), BF16)
            ops.rope_kv(qkv, c["q"], c["kc"], c["vc"], batch=B, s=S, hq=nh, hkv=nkv, d=hd, smax=S, pos0=0, theta=theta)
            c["ctx"] = ops.empty((M, nh * hd), BF16)
            c["lse"] = ops.empty((B, nh, (S + 63) // 64 * 64), F32)
            ops.attention(c["q"], c["kc"], c["vc"], c["ctx"], batch=B, hq=nh, hkv=nkv, sq=S, sk=S, d=hd, q_strides=(S * nh * hd, nh * hd),
                          k_strides=kvs, v_strides=kvs, o_strides=(S * nh * hd, nh * hd), scale=scale, softcap=cap, causal=causal,
                          causal_prefix=prefix if causal else 0, lse=c["lse"], window=win if li % 2 == 0 else 0)
            c["br"], c["u_o"] = self._lin_fwd(c["ctx"], L_["wo"], F_[f"gem.{li}.o"], M, out_dtype=F32)
            c["x_mid"], c["h2"] = ops.empty((M, H), F32), ops.empty((M, H), BF16)
            ops.rmsnorm_train_fwd(x, branch=c["br"], w_post=L_["ln_post_attn"], w_pre=L_["ln_pre_ff"], eps=eps, x_out=c["x_mid"], h=c["h2"])
            c["gu"], c["u_gu"] = self._lin_fwd(c["h2"], L_["wgu"], F_[f"gem.{li}.gu"], M)
            c["act"] = ops.empty((M, FF), BF16)
            ops.geglu_fwd(c["gu"], c["act"])
            c["br2"], c["u_d"] = self._lin_fwd(c["act"], L_["wd"], F_[f"gem.{li}.down"], M, out_dtype=F32)
            nxt = g["layers"][li + 1]["ln_in"] if li + 1 < nL else g["final"]
            x_out, h1 = ops.empty((M, H), F32), ops.empty((M, H), BF16)
            ops.rmsnorm_train_fwd(c["x_mid"], branch=c["br2"], w_post=L_["ln_post_ff"], w_pre=nxt, eps=eps, x_out=x_out, h=h1)
            x = x_out
            saved.append(c)
        return h1, {"layers": saved, "x_last": x}

    # ------------------------------------------------------------------------------------------ backward
    def _gemma_bwd(self, sv, dx, B, S, causal, prefix):
        """dx fp32 [M, H]: gradient w.r.t. the last layer's output residual stream; on return w.r.t. the embedded inputs."""
        ops, e, g, t, F_ = self.ops, self.eng, self.eng.gem, self.eng.t, self.lay.fused
        nh, nkv, hd = t["num_attention_heads"], t["num_key_value_heads"], t["head_dim"]
        eps, theta = t["rms_norm_eps"], float(t.get("rope_theta", 10000.0))
        scale, cap = t["query_pre_attn_scalar"] ** -0.5, t["attn_logit_softcapping"] or 0.0
        M = B * S
        Wd = (nh + 2 * nkv) * hd
        kvs = (S * nkv * hd, nkv * hd)
        qs = (S * nh * hd, nh * hd)
        win = int(t.get("sliding_window") or 0)
        win = win if S > win else 0              # as in the forward: even layers are windowed once the sequence exceeds the window
        for li in reversed(range(len(g["layers"]))):
            L_, T_, c = g["layers"][li], self.gem_t[li], sv["layers"][li]
            H = dx.shape[1]
            dbr2 = ops.empty((M, H), BF16)
            ops.rmsnorm_bwd(c["br2"], L_["ln_post_ff"], dx, eps=eps, dx_bf16=dbr2)
            dact = self._lin_bwd(dbr2, T_["wd"], F_[f"gem.{li}.down"], c["act"], c["u_d"], M)
            dgu = ops.empty(c["gu"].shape, BF16)
            ops.geglu_bwd(c["gu"], dact, dgu)
            dh2 = self._lin_bwd(dgu, T_["wgu"], F_[f"gem.{li}.gu"], c["h2"], c["u_gu"], M)
            ops.rmsnorm_bwd(c["x_mid"], L_["ln_pre_ff"], dh2, eps=eps, dx_accum=dx)
            dbr = ops.empty((M, H), BF16)
            ops.rmsnorm_bwd(c["br"], L_["ln_post_attn"], dx, eps=eps, dx_bf16=dbr)
            dctx = self._lin_bwd(dbr, T_["wo"], F_[f"gem.{li}.o"], c["ctx"], c["u_o"], M)
            dqkv = ops.empty((M, Wd), BF16)
            ops.attention_bwd(c["q"], c["kc"], c["vc"], c["ctx"], dctx, dqkv, dqkv[:, nh * hd:], dqkv[:, (nh + nkv) * hd:], batch=B, hq=nh,
                              hkv=nkv, sq=S, sk=S, d=hd, q_strides=qs, k_strides=kvs, v_strides=kvs, o_strides=qs, do_strides=qs,
                              dq_strides=(S * Wd, Wd), dk_strides=(S * Wd, Wd), dv_strides=(S * Wd, Wd), scale=scale, softcap=cap,
                              causal=causal, causal_prefix=prefix if causal else 0, lse=c["lse"],
                              window=win if li % 2 == 0 else 0)
            ops.rope_bwd(dqkv, batch=B, s=S, hq=nh, hkv=nkv, d=hd, theta=theta)
            dh1 = self._lin_bwd(dqkv, T_["wqkv"], F_[f"gem.{li}.qkv"], c["h1"], c["u_qkv"], M)
            ops.rmsnorm_bwd(c["x_in"], L_["ln_in"], dh1, eps=eps, dx_accum=dx)
            sv["layers"][li] = None                      # the layer's activations are dead: let the allocator reuse them
        return dx

    def _vision_bwd(self, sv, dfeat_b, B):
        """dfeat_b bf16 [B*256, H]: gradient w.r.t. the projector's pre-scale output."""
        ops, e, s, v, F_ = self.ops, self.eng, self.eng.sig, self.eng.v, self.lay.fused
        D, nh = v["hidden_size"], v["num_attention_heads"]
        S, M = 256, B * 256
        eps = v.get("layer_norm_eps", 1e-6)
        dsrc = self._lin_bwd(dfeat_b, self.proj_t, F_["proj"], sv["src"], sv["u_proj"], M)
        if e.use_zoe:
            dhb = self._lin_bwd(dsrc, self.ego3_t, F_["ego.3"], sv["hb"], sv["u_e3"], M)
            dh0 = ops.empty((M, D), BF16)
            ops.layernorm_bwd(sv["h0"], e.ego["ln_g"], e.ego["ln_b"], dhb, eps=1e-5, relu=True, dx_bf16=dh0)
            self._lin_bwd(dh0, None, F_["ego.0"], sv["enc"], sv["u_e0"], M, need_dx=False)
        dx, dxb = ops.empty((M, D), F32), ops.empty((M, D), BF16)
        ops.fill_zero(dx)
        ops.layernorm_bwd(sv["x_last"], s["post_g"], s["post_b"], dsrc, eps=eps, dx_accum=dx, copy_bf16=dxb)
        hd = D // nh
        st = (S * 3 * D, 3 * D)
        os_ = (S * D, D)
        for li in reversed(range(len(s["layers"]))):
            L_, T_, c = s["layers"][li], self.sig_t[li], sv["layers"][li]
            df = self._lin_bwd(dxb, T_["w2"], F_[f"sig.{li}.fc2"], c["f"], c["u_fc2"], M)
            dz = ops.empty(c["z"].shape, BF16)
            ops.gelu_tanh_bwd(c["z"], df, dz)
            dh2 = self._lin_bwd(dz, T_["w1"], F_[f"sig.{li}.fc1"], c["h2"], c["u_fc1"], M)
            ops.layernorm_bwd(c["x_mid"], L_["ln2_g"], L_["ln2_b"], dh2, eps=eps, dx_accum=dx, copy_bf16=dxb)
            dctx = self._lin_bwd(dxb, T_["wo"], F_[f"sig.{li}.o"], c["ctx"], c["u_o"], M)
            dqkv = ops.empty((M, 3 * D), BF16)
            qkv = c["qkv"]
            ops.attention_bwd(qkv, qkv[:, D:], qkv[:, 2 * D:], c["ctx"], dctx, dqkv, dqkv[:, D:], dqkv[:, 2 * D:], batch=B, hq=nh, hkv=nh,
                              sq=S, sk=S, d=hd, q_strides=st, k_strides=st, v_strides=st, o_strides=os_, do_strides=os_, dq_strides=st,
                              dk_strides=st, dv_strides=st, scale=hd ** -0.5, lse=c["lse"])
            last = li == 0                       # nothing trainable below the first block (the patch embedding is not a LoRA target)
            dh1 = self._lin_bwd(dqkv, T_["wqkv"], F_[f"sig.{li}.qkv"], c["h1"], c["u_qkv"], M, need_dx=not last)
            if not last:
                ops.layernorm_bwd(c["x_in"], L_["ln1_g"], L_["ln1_b"], dh1, eps=eps, dx_accum=dx, copy_bf16=dxb)
            sv["layers"][li] = None

    # ------------------------------------------------------------------------------------------ the step
    def forward_backward(self, input_ids, pixel_values, intrinsic, labels, token_type_ids=None, attention_mask=None,
                         on_language_grads_ready=None):
        """Loss forward + full backward: fills `self.lay.grad` (sum over this rank's batch of d(mean CE)/d(adapter)).
        Mask selection as `forward(labels=...)` (model/modeling_spatialvla.py:258-306): token_type_ids -> triangular, plus the
        prefix columns when a 2-D attention_mask is passed; labels alone -> bidirectional.  Returns the fp32 [3] loss summary
        (mean loss, labelled rows, argmax hits) on the device."""
        ops, e = self.ops, self.eng
        B, L = input_ids.shape
        dev = ops.device
        ignore = e.cfg.get("ignore_index", -100)
        ignore = -100 if ignore is None else ignore
        causal, prefix = False, 0
        if token_type_ids is not None:
            causal = True
            if attention_mask is not None:
                tt = token_type_ids.to("cpu", torch.int64)
                prefix = int((tt[0] == 0).sum())
                if not bool(torch.equal(tt, (torch.arange(L)[None, :] >= prefix).to(torch.int64).expand(B, L))):
                    raise NotImplementedError("token_type_ids must be 0...01...1 with the same prefix length in every row")
        if attention_mask is not None and bool((attention_mask == 0).any()):
            raise NotImplementedError("training step: only unpadded batches")
        # label bookkeeping on the host (the collator's tensors live there): shifted positions that carry a label
        lab, ids_cpu = labels.to("cpu", torch.int64), input_ids.to("cpu", torch.int64)
        pad_id = e.cfg.get("pad_token_id")
        pad_id = -1 if pad_id is None else pad_id
        if bool((lab == pad_id).any()):
            lab = torch.where(ids_cpu == pad_id, torch.full_like(lab, ignore), lab)
        shift = lab[:, 1:]
        bi, ti = torch.nonzero(shift != ignore, as_tuple=True)
        rows = (bi * L + ti).to(dev)
        row_labels = shift[bi, ti].to(dev).contiguous()
        ib, it = torch.nonzero(ids_cpu == e.cfg["image_token_index"], as_tuple=True)
        img_rows = (ib * L + it).to(dev)
        if img_rows.numel() != B * 256:
            raise ValueError("Number of images does not match number of special image tokens in the input text.")
        ids = input_ids.to(dev, torch.int64).contiguous()
        px = pixel_values.to(dev, F32).contiguous()
        K = intrinsic.to(dev, F32).contiguous()

        self.lay.zero_grad()
        self.lay.pack()
        feats, vsv = self._vision_fwd(px, K, B)
        x0, _ = e.embed(ids, feats)
        h, gsv = self._gemma_fwd(x0, B, L, causal, prefix)
        summary, row_loss, dh = e.labelled_loss_backward(h, rows, row_labels, ignore_index=ignore)
        H = e.t["hidden_size"]
        dx = ops.empty((B * L, H), F32)
        ops.fill_zero(dx)
        ops.rmsnorm_bwd(gsv["x_last"], e.gem["final"], dh, eps=e.t["rms_norm_eps"], row_idx=rows, dx_accum=dx)
        dx = self._gemma_bwd(gsv, dx, B, L, causal, prefix)
        if on_language_grads_ready is not None:
            on_language_grads_ready()            # the Gemma2 segment of the gradient arena is final: its all-reduce can start now
        # d(embeddings): only the image-token rows lead to trainable parameters (embed_tokens is frozen)
        normalizer = float(torch.tensor(H ** 0.5, dtype=F32))
        dfeat_b = ops.empty((B * 256, H), BF16)
        ops.rows_cast(dx, dfeat_b, row_idx=img_rows, scale=normalizer * float(torch.tensor(1.0 / (H ** 0.5), dtype=F32)))
        self._vision_bwd(vsv, dfeat_b, B)
        return summary

    def optimizer_step(self, world_size: int = 1):
        """Clip (global L2 norm of the averaged gradient <= max_grad_norm, HF Trainer default 1.0) + AdamW, device-side."""
        self.step_count += 1
        ops, lay = self.ops, self.lay
        ops.fill_zero(self.norm_sq)
        ops.sumsq(lay.grad, self.norm_sq)
        ops.adamw_step(lay.param, lay.grad, lay.exp_avg, lay.exp_avg_sq, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                       weight_decay=self.weight_decay, step=self.step_count, grad_scale=1.0 / world_size, sumsq=self.norm_sq,
                       max_grad_norm=self.max_grad_norm)

    def step(self, batch: dict, group=None):
        """One data-parallel training step on this rank's shard: forward/backward, ONE gradient all-reduce, clip + AdamW.
        Returns the local loss summary tensor (device)."""
        from . import parallel
        red = parallel.GradientReducer(self.lay, self.lay.n_language, group=group)
        summary = self.forward_backward(batch["input_ids"], batch["pixel_values"], batch["intrinsic"], batch["labels"],
                                        token_type_ids=batch.get("token_type_ids"), attention_mask=batch.get("attention_mask"),
                                        on_language_grads_ready=red.first_segment_ready)
        world = red.finish()                     # sums; the 1 / world average is folded into the optimizer kernel's gradient scale
        self.optimizer_step(world_size=world)
        return summary
