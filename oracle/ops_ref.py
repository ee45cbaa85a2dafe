"""TEST INFRASTRUCTURE ONLY.  Per-op torch (CPU, fp32 arithmetic) re-statement of every C-ABI kernel with the same
Python signatures as spatialvla_b200.ops.CudaOps.  Two uses, both in tests/:
  * the per-kernel oracle of the `-m gpu` parity tests (CUDA op vs this, same inputs);
  * injected as the op backend to run the host orchestration (spatialvla_b200/engine.py) on CPU and compare it
    with oracle/model_ref.py and the golden vectors -- the product never imports this file.
Rounding points (bf16 outputs) mirror the kernels so the CPU run predicts the B200 path's numerics.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from spatialvla_b200._lib import (ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_RELU, ACT_SOFTCAP, ACT_SOFTPLUS)

BF16, F32 = torch.bfloat16, torch.float32


def _act(v, act, p):
    if act == ACT_GELU_TANH:
        return F.gelu(v, approximate="tanh")
    if act == ACT_GELU_ERF:
        return F.gelu(v)
    if act == ACT_RELU:
        return F.relu(v)
    if act == ACT_SOFTCAP:
        return p * torch.tanh(v / p)
    if act == ACT_SOFTPLUS:
        return F.softplus(v)
    return v


def _store_bf16(out, v, n):
    """bf16 store of v[:, :n]; a 3-D `out` [2, M, N] is a hi/lo pair: hi = bf16(v), lo = bf16(v - hi)"""
    if out.dim() == 3:
        hi = v.to(BF16)
        out[0][:, :n] = hi
        out[1][:, :n] = (v - hi.float()).to(BF16)
    else:
        out[:, :n] = v.to(BF16)


def _rows(t):
    return t if t.dim() == 2 else t.view(-1, t.shape[-1])


class RefOps:
    name = "ref"

    def __init__(self, device="cpu"):
        self.device = torch.device(device)
        self.launches = 0

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def launch_count(self):
        return self.launches

    # ---- G1
    def gemm(self, a, w, *, n=None, k=None, out_bf16=None, out_f32=None, out_relu=None, bias=None, colscale=None,
             res_bf16=None, res2_bf16=None, res_f32=None, res_mod=0, act=ACT_NONE, act_param=0.0, alpha=1.0,
             geglu=False, accumulate=False, conv=None, block_n=0, impl=None, a2=None, w2=None):
        self.launches += 1
        N = int(n if n is not None else w.shape[0])
        wf = w[:N].float()
        if conv is not None:
            nb, h, wd, c = conv
            cpad = (c + 63) // 64 * 64
            x = a.float().view(nb, h, wd, c).permute(0, 3, 1, 2)
            w4 = wf.view(N, 9, cpad)[:, :, :c].reshape(N, 3, 3, c).permute(0, 3, 1, 2)
            acc = F.conv2d(x, w4, None, padding=1).permute(0, 2, 3, 1).reshape(nb * h * wd, N)
        else:
            K = int(k if k is not None else a.shape[1])
            acc = a[:, :K].float() @ wf[:, :K].t()
            if a2 is not None:                       # K extension: second operand pair accumulated into the same tile
                acc = acc + a2.float() @ w2[:N].float().t()
        v = acc * alpha
        if bias is not None:
            v = v + bias[:N]
        if geglu:
            g, u = v[:, 0::2], v[:, 1::2]
            _rows(out_bf16)[:, : N // 2] = (F.gelu(g, approximate="tanh") * u).to(BF16)
            return
        v = _act(v, act, act_param)
        if colscale is not None:
            v = v * colscale[:N]
        if res_bf16 is not None:
            v = v + _rows(res_bf16)[:, :N].float()
        if res2_bf16 is not None:
            v = v + _rows(res2_bf16)[:, :N].float()
        if res_f32 is not None:
            r = _rows(res_f32)[:, :N]
            if res_mod:
                idx = torch.arange(v.shape[0]) % res_mod
                r = r[idx]
            v = v + r
        if out_f32 is not None:
            o = _rows(out_f32)
            if accumulate:
                v = v + o[:, :N]
            o[:, :N] = v
        if out_bf16 is not None:
            _rows(out_bf16)[:, :N] = v.to(BF16)
        if out_relu is not None:
            _rows(out_relu)[:, :N] = F.relu(v).to(BF16)

    # ---- fine-tune step: training forward pieces and backward kernels (closed forms = oracle/backward_ref.py)
    def rmsnorm_train_fwd(self, x_in, *, branch=None, w_post=None, w_pre=None, eps=1e-6, x_out=None, h=None):
        self.launches += 1
        x = x_in
        if branch is not None:
            r = torch.rsqrt(branch.pow(2).mean(-1, keepdim=True) + eps)
            x_out.copy_(x_in + branch * r * (1.0 + w_post))
            x = x_out
        if w_pre is not None:
            r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
            h.copy_((x * r * (1.0 + w_pre)).to(BF16))

    def rmsnorm_bwd(self, x, w, dy, *, eps=1e-6, row_idx=None, dx_accum=None, dx_bf16=None):
        from .backward_ref import rmsnorm_bwd
        self.launches += 1
        xr = x if row_idx is None else x[row_idx]
        dx, _ = rmsnorm_bwd(xr, w, dy.float(), eps)
        if dx_accum is not None:
            if row_idx is None:
                dx_accum += dx
            else:
                dx_accum.index_add_(0, row_idx, dx)
        if dx_bf16 is not None:
            if row_idx is None:
                dx_bf16.copy_(dx.to(BF16))
            else:
                dx_bf16[row_idx] = dx.to(BF16)

    def layernorm_bwd(self, x, gamma, beta, dy, *, eps, relu=False, dx_accum=None, copy_bf16=None, dx_bf16=None):
        self.launches += 1
        mu = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt((x - mu).pow(2).mean(-1, keepdim=True) + eps)
        xh = (x - mu) * rstd
        g = dy.float() * gamma
        if relu:
            g = torch.where(xh * gamma + beta > 0, g, torch.zeros_like(g))
        dx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        if dx_bf16 is not None:
            dx_bf16.copy_(dx.to(BF16))
        if dx_accum is not None:
            dx_accum += dx
            if copy_bf16 is not None:
                copy_bf16.copy_(dx_accum.to(BF16))

    def geglu_fwd(self, gu, act):
        self.launches += 1
        g, u = gu[:, 0::2].float(), gu[:, 1::2].float()
        act.copy_((F.gelu(g, approximate="tanh") * u).to(BF16))

    def geglu_bwd(self, gu, dact, dgu):
        from .backward_ref import geglu_bwd
        self.launches += 1
        dg, du = geglu_bwd(gu[:, 0::2], gu[:, 1::2], dact)
        dgu[:, 0::2] = dg.to(BF16)
        dgu[:, 1::2] = du.to(BF16)

    def gelu_tanh_fwd(self, z, f):
        self.launches += 1
        f.copy_(F.gelu(z.float(), approximate="tanh").to(BF16))

    def gelu_tanh_bwd(self, z, df, dz):
        from .backward_ref import gelu_tanh_grad
        self.launches += 1
        dz.copy_((df.float() * gelu_tanh_grad(z.float())).to(BF16))

    def rope_bwd(self, dqkv, *, batch, s, hq, hkv, d, theta):
        from .backward_ref import rope_bwd
        self.launches += 1
        t = dqkv.float().view(batch, s, hq + 2 * hkv, d)
        pos = torch.arange(1, s + 1)
        qk = t[:, :, : hq + hkv].permute(0, 2, 1, 3)                          # [B, heads, S, d]
        t[:, :, : hq + hkv] = rope_bwd(qk, pos, theta).permute(0, 2, 1, 3)
        dqkv.copy_(t.reshape(dqkv.shape).to(BF16))

    def rows_cast(self, src, out, *, row_idx=None, scale=1.0):
        self.launches += 1
        out.copy_(((src if row_idx is None else src[row_idx]) * scale).to(BF16))

    def lora_pack(self, arena, pool, plan):
        """plan.records: list of (src_off, dst_off, stride_i, stride_j, rows, cols)"""
        self.launches += 1
        flat = pool.view(-1)
        for (so, do, si, sj, rows, cols) in plan.records:
            src = arena[so:so + rows * cols].view(rows, cols)
            idx = (do + torch.arange(rows)[:, None] * si + torch.arange(cols)[None, :] * sj).reshape(-1)
            flat[idx] = src.reshape(-1).to(BF16)

    def fill_zero(self, t):
        t.zero_()

    def attention_bwd(self, q, k, v, out, dout, dq, dk, dv, *, batch, hq, hkv, sq, sk, d, q_strides, k_strides, v_strides, o_strides,
                      do_strides, dq_strides, dk_strides, dv_strides, scale, softcap=0.0, causal=False, causal_prefix=0, lse=None, window=0):
        from .backward_ref import softcap_attention_bwd
        self.launches += 3
        Q = self._strided(q, q_strides[0], q_strides[1], batch, sq, hq, d).float().permute(0, 2, 1, 3)
        K = self._strided(k, k_strides[0], k_strides[1], batch, sk, hkv, d).float().permute(0, 2, 1, 3)
        V = self._strided(v, v_strides[0], v_strides[1], batch, sk, hkv, d).float().permute(0, 2, 1, 3)
        dO = self._strided(dout, do_strides[0], do_strides[1], batch, sq, hq, d).float().permute(0, 2, 1, 3)
        G = hq // hkv
        Kr, Vr = K.repeat_interleave(G, 1), V.repeat_interleave(G, 1)
        mask = None
        if causal:
            mask = torch.arange(sk)[None, :] > torch.clamp(torch.arange(sq)[:, None] + (sk - sq), min=causal_prefix - 1)
        if window:
            wm = (torch.arange(sq)[:, None] + (sk - sq) - torch.arange(sk)[None, :]) >= window
            mask = wm if mask is None else (mask | wm)
        gq, gk, gv = softcap_attention_bwd(Q, Kr, Vr, dO, scale, softcap, mask)
        gk = gk.view(batch, hkv, G, sk, d).sum(2)
        gv = gv.view(batch, hkv, G, sk, d).sum(2)
        self._strided(dq, dq_strides[0], dq_strides[1], batch, sq, hq, d).copy_(gq.permute(0, 2, 1, 3).to(BF16))
        self._strided(dk, dk_strides[0], dk_strides[1], batch, sk, hkv, d).copy_(gk.permute(0, 2, 1, 3).to(BF16))
        self._strided(dv, dv_strides[0], dv_strides[1], batch, sk, hkv, d).copy_(gv.permute(0, 2, 1, 3).to(BF16))

    def gemm_tn(self, s, y, groups, *, r, n, scale=1.0):
        """groups: list of (dst fp32 [rows, >= ncols] view, row0, rows, col_start, col_stride, ncols); dst += scale * S[:, rows]^T Y[:, cols]"""
        self.launches += 1
        full = (s[:, :r].float().t() @ y[:, :n].float()) * scale
        for (dst, row0, rows, c0, cs, nc) in groups:
            dst[:, :nc] += full[row0:row0 + rows, c0:c0 + (nc - 1) * cs + 1:cs]

    # ---- G1s
    def skinny_splits(self, n, k):
        return max(1, min(148 // max(1, (n + 127) // 128), max(1, ((k + 63) // 64) // 4)))

    def gemm_skinny(self, x, w, *, out_bf16=None, out_f32=None, bias=None, act=ACT_NONE, act_param=0.0, alpha=1.0,
                    geglu=False, splits=1, tiled_n=None, pair=False):
        self.launches += 1
        if x.dim() == 3:                    # hi/lo activation pair [2, M, K]: hi @ w.T + lo @ w.T in fp32
            x = x[0].float() + x[1].float()
        if tiled_n is not None:             # tile-major [nt, kb, 128, 64] copy of a [tiled_n, K] matrix
            nt, kb = w.shape[0], w.shape[1]
            w = w.permute(0, 2, 1, 3).reshape(nt * 128, kb * 64)[:tiled_n, : x.shape[1]]
        xf, wf = x.float(), w.float()
        if out_f32 is not None and out_f32.dim() == 3:
            S, K = out_f32.shape[0], x.shape[1]
            kb = (K + 63) // 64
            per = (kb + S - 1) // S
            for sidx in range(S):
                lo, hi = sidx * per * 64, min(K, (sidx + 1) * per * 64)
                out_f32[sidx, :, : w.shape[0]] = xf[:, lo:hi] @ wf[:, lo:hi].t()
            return
        v = (xf @ wf.t()) * alpha
        if bias is not None:
            v = v + bias
        if geglu:
            _store_bf16(out_bf16, F.gelu(v[:, 0::2], approximate="tanh") * v[:, 1::2], w.shape[0] // 2)
            return
        v = _act(v, act, act_param)
        if out_f32 is not None:
            out_f32[:, : w.shape[0]] = v
        if out_bf16 is not None:
            _store_bf16(out_bf16, v, w.shape[0])

    # ---- G2 / G3
    @staticmethod
    def _strided(t, bs, ss, batch, s, heads, d):
        """(B, S, H, D) view over the tensor's own storage, exactly the addressing the kernel uses."""
        return t.as_strided((batch, s, heads, d), (bs, ss, d, 1), t.storage_offset())

    def attention(self, q, k, v, out, *, batch, hq, hkv, sq, sk, d, q_strides, k_strides, v_strides, o_strides,
                  scale, softcap=0.0, causal=False, relpos_table=None, relpos_win=0, relpos_head_major=False, kv_start=None,
                  causal_prefix=0, lse=None, window=0):
        self.launches += 1
        Q = self._strided(q, *q_strides, batch, sq, hq, d).float().permute(0, 2, 1, 3)
        K = self._strided(k, *k_strides, batch, sk, hkv, d).float().permute(0, 2, 1, 3).repeat_interleave(hq // hkv, 1)
        V = self._strided(v, *v_strides, batch, sk, hkv, d).float().permute(0, 2, 1, 3).repeat_interleave(hq // hkv, 1)
        s = (Q @ K.transpose(-1, -2)) * scale
        if softcap:
            s = softcap * torch.tanh(s / softcap)
        if relpos_table is not None:
            if relpos_head_major:
                relpos_table = relpos_table.t()
            from oracle.model_ref import beit_rel_pos_bias
            s = s + beit_rel_pos_bias(relpos_table, relpos_win)[None]
        if kv_start is not None:
            s = s.masked_fill((torch.arange(sk)[None, :] < kv_start.long()[:, None])[:, None, None, :], float("-inf"))
        if causal:
            # prefix-LM (training forward, model/modeling_spatialvla.py:292-305): keys < causal_prefix stay visible
            m = torch.arange(sk)[None, :] > torch.clamp(torch.arange(sq)[:, None] + (sk - sq), min=causal_prefix - 1)
            s = s.masked_fill(m, float("-inf"))
        if window:                  # sliding-window layer: key slot j masked for query slot i when i - j >= window
            s = s.masked_fill((torch.arange(sq)[:, None] + (sk - sq) - torch.arange(sk)[None, :]) >= window, float("-inf"))
        # the kernel rounds the un-normalised probabilities to bf16 and divides by the fp32 row sum
        mx = s.max(-1, keepdim=True).values
        e = torch.exp(s - mx)
        o = (e.to(BF16).float() @ V) / e.sum(-1, keepdim=True)
        self._strided(out, *o_strides, batch, sq, hq, d).copy_(o.permute(0, 2, 1, 3).to(BF16))
        if lse is not None:                      # log2-domain log-sum-exp of every row, kept for the backward pass
            lse[:, :, :sq] = torch.logsumexp(s, -1) * 1.4426950408889634

    def decode_attention(self, q, kcache, vcache, out, *, batch, hq, hkv, d, smax, ctx, scale, softcap=0.0, kv_start=None, window=0):
        self.launches += 1
        Q = q.float().view(batch, hq, 1, d)
        K = kcache.float().view(batch, smax, hkv, d)[:, :ctx].permute(0, 2, 1, 3).repeat_interleave(hq // hkv, 1)
        V = vcache.float().view(batch, smax, hkv, d)[:, :ctx].permute(0, 2, 1, 3).repeat_interleave(hq // hkv, 1)
        s = (Q @ K.transpose(-1, -2)) * scale
        if softcap:
            s = softcap * torch.tanh(s / softcap)
        if kv_start is not None:
            s = s.masked_fill((torch.arange(ctx)[None, :] < kv_start.long()[:, None])[:, None, None, :], float("-inf"))
        if window:                  # re-statement only (the fused kernel is the windowed decode path on the device)
            s = s.masked_fill((torch.arange(ctx) < ctx - window)[None, None, None, :], float("-inf"))
        p = torch.softmax(s, -1).to(BF16).float()
        out.view(batch, hq, d)[:] = (p @ V).squeeze(2).to(BF16)

    def decode_attention_fused(self, qkv_partials, kcache, vcache, out, *, batch, hq, hkv, d, smax, ctx, theta, scale, softcap=0.0,
                               kv_start=None, window=0):
        if out.dim() == 3:                      # hi/lo output pair: the kernel's own precision (fp32 q and probabilities, bf16 cache)
            q = torch.empty(batch, hq * d, dtype=F32)
            self.rope_kv(qkv_partials, q, kcache, vcache, batch=batch, s=1, hq=hq, hkv=hkv, d=d, smax=smax, pos0=ctx - 1, theta=theta,
                         row_pads=kv_start)
            Q = q.view(batch, hq, 1, d)
            K = kcache.float().view(batch, smax, hkv, d)[:, :ctx].permute(0, 2, 1, 3).repeat_interleave(hq // hkv, 1)
            V = vcache.float().view(batch, smax, hkv, d)[:, :ctx].permute(0, 2, 1, 3).repeat_interleave(hq // hkv, 1)
            s = (Q @ K.transpose(-1, -2)) * scale
            if softcap:
                s = softcap * torch.tanh(s / softcap)
            if kv_start is not None:
                s = s.masked_fill((torch.arange(ctx)[None, :] < kv_start.long()[:, None])[:, None, None, :], float("-inf"))
            if window:
                s = s.masked_fill((torch.arange(ctx) < ctx - window)[None, None, None, :], float("-inf"))
            _store_bf16(out, (torch.softmax(s, -1) @ V).squeeze(2).reshape(batch, hq * d), hq * d)
            return
        q = torch.empty(batch, hq * d, dtype=BF16)
        self.rope_kv(qkv_partials, q, kcache, vcache, batch=batch, s=1, hq=hq, hkv=hkv, d=d, smax=smax, pos0=ctx - 1, theta=theta,
                     row_pads=kv_start)
        self.launches -= 1                      # one launch on the device
        self.decode_attention(q, kcache, vcache, out, batch=batch, hq=hq, hkv=hkv, d=d, smax=smax, ctx=ctx, scale=scale, softcap=softcap,
                              kv_start=kv_start, window=window)

    # ---- fused memory-bound ops
    def layernorm(self, x, gamma, beta, eps, *, out_bf16=None, out_f32=None, relu=False):
        self.launches += 1
        y = F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
        if relu:
            y = F.relu(y)
        if out_f32 is not None:
            out_f32.copy_(y.view_as(out_f32))
        if out_bf16 is not None:
            out_bf16.copy_(y.view_as(out_bf16).to(BF16))

    def rmsnorm_residual(self, x, *, branch=None, w_post=None, w_pre=None, eps=1e-6, out_bf16=None):
        self.launches += 1

        def rms(t, w):
            return t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + w)
        if branch is not None:
            if branch.dim() == 3:
                branch = branch.sum(0)
            x.add_(rms(branch.reshape(x.shape), w_post))
        if w_pre is not None:
            if out_bf16.dim() == 3 and out_bf16.shape[0] == 2 and out_bf16.shape[1] == x.numel() // x.shape[-1]:
                _store_bf16(out_bf16, rms(x, w_pre).view_as(out_bf16[0]), x.shape[-1])     # hi/lo pair of the decode chain
            else:
                out_bf16.copy_(rms(x, w_pre).view_as(out_bf16).to(BF16))

    def rope_kv(self, qkv, q_out, kcache, vcache, *, batch, s, hq, hkv, d, smax, pos0, theta, row_pads=None):
        self.launches += 1
        if qkv.dtype == F32 and qkv.dim() == 3:
            qkv = qkv.sum(0)
        t = qkv.float().view(batch, s, hq + 2 * hkv, d)
        slot = torch.arange(pos0, pos0 + s)[None, :].expand(batch, s)
        if row_pads is None:
            pos = slot.float() + 1.0
        else:       # left-padded rows: positions restart at 1 on the first real token, padding slots sit at 2
            pads = row_pads.long()[:, None]
            pos = torch.where(slot < pads, torch.full_like(slot, 2), slot - pads + 1).float()
        inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
        fr = pos[..., None] * inv
        cos, sin = torch.cat([fr, fr], -1).cos()[:, :, None], torch.cat([fr, fr], -1).sin()[:, :, None]
        qk = t[:, :, : hq + hkv]
        x1, x2 = qk[..., : d // 2], qk[..., d // 2:]
        rot = qk * cos + torch.cat([-x2, x1], -1) * sin
        q_out.view(batch, s, hq, d).copy_(rot[:, :, :hq].to(q_out.dtype))
        kcache.view(batch, smax, hkv, d)[:, pos0:pos0 + s] = rot[:, :, hq:].to(BF16)
        vcache.view(batch, smax, hkv, d)[:, pos0:pos0 + s] = t[:, :, hq + hkv:].to(BF16)

    def embed_tokens(self, ids, embed, spatial_embed, image_feats, x, *, image_token, act_lo, n_act, n_img,
                     normalizer, status):
        self.launches += 1
        B, S = ids.shape
        H = x.shape[-1]
        out = embed[ids.clamp(0, embed.shape[0] - 1)].float()
        if spatial_embed is not None:
            sel = (ids >= act_lo) & (ids < act_lo + n_act)
            out[sel] = spatial_embed[ids[sel] - act_lo].float()
        if image_feats is not None:
            m = ids == image_token
            rank = torch.cumsum(m.long(), 1) - 1
            if bool((m.sum(1) != n_img).any()):            # too many OR too few image tokens in some row
                status.fill_(1)
                return
            bidx = torch.arange(B)[:, None].expand(B, S)
            out[m] = image_feats.view(B, n_img, H)[bidx[m], rank[m]]
        x.view(B, S, H).copy_(out * normalizer)

    def argmax_rows(self, logits, out_ids, *, id_offset=0):
        self.launches += 1
        out_ids.copy_(logits.argmax(-1) + id_offset)

    def cross_entropy_rows(self, logits, labels, row_loss, row_argmax, *, row_offset=0, summary=None, ignore_index=-100):
        """nn.CrossEntropyLoss pieces of model/modeling_spatialvla.py:413-430 on already selected rows (chunked like the kernel)."""
        self.launches += 1 if summary is None else 2
        r = logits.shape[0]
        lab = labels[row_offset:row_offset + r]
        row_loss[row_offset:row_offset + r] = F.cross_entropy(logits.double(), lab, ignore_index=ignore_index, reduction="none").float()
        row_argmax[row_offset:row_offset + r] = logits.argmax(-1)
        if summary is not None:
            n = row_offset + r
            valid = labels[:n] != ignore_index
            summary[0] = row_loss[:n][valid].sum() / valid.sum()
            summary[1] = valid.sum()
            summary[2] = (row_argmax[:n][valid] == labels[:n][valid]).sum()

    def cross_entropy_bwd(self, logits, labels, row_loss, summary, dz, *, row_offset=0, softcap=0.0, ignore_index=-100):
        """d(mean CE)/d(pre-softcap logits): (softmax - onehot) * (1 - (logit/cap)^2) / count (model/modeling_spatialvla.py:413-430,
        model/modeling_gemma2.py:993-997), zero for ignored rows and for the K-padding columns."""
        self.launches += 1
        r, c = logits.shape
        lab = labels[row_offset:row_offset + r]
        live = lab != ignore_index
        lg = logits.double()
        g = torch.softmax(lg, -1)
        g[live, lab[live]] -= 1.0
        if softcap:
            g = g * (1.0 - (lg / softcap) ** 2)
        g = g / float(summary[1]) * live[:, None]
        dz.zero_()
        dz[:, :c] = g.to(BF16)

    def sumsq(self, x, out):
        self.launches += 1
        out += (x.double() ** 2).sum().float()

    def adamw_step(self, param, grad, exp_avg, exp_avg_sq, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step, grad_scale=1.0,
                   sumsq=None, max_grad_norm=0.0):
        """torch.optim.AdamW single-tensor arithmetic (torch/optim/adamw.py: _single_tensor_adamw), in place, preceded by
        torch.nn.utils.clip_grad_norm_ on the scaled gradient when `sumsq` (sum of squares of the un-scaled gradient) is given."""
        self.launches += 1
        if sumsq is not None:
            total = float(sumsq.sqrt()) * grad_scale
            grad_scale = grad_scale * min(1.0, max_grad_norm / (total + 1e-6))
        g = grad * grad_scale
        param.mul_(1 - lr * weight_decay)
        exp_avg.lerp_(g, 1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        denom = (exp_avg_sq.sqrt() / (bc2 ** 0.5)).add_(eps)
        param.addcdiv_(exp_avg, denom, value=-(lr / bc1))

    def siglip_patchify(self, px, a):
        self.launches += 1
        B = px.shape[0]
        x = (px - 0.5) / 0.5
        col = F.unfold(x, kernel_size=14, stride=14).transpose(1, 2).reshape(B * 256, 588)
        a.zero_()
        a[:, :588] = col.to(BF16)

    def zoe_patchify(self, px, a):
        self.launches += 1
        from oracle.model_ref import process_zoe
        B = px.shape[0]
        col = F.unfold(process_zoe(px), kernel_size=16, stride=16).transpose(1, 2).reshape(B * 576, 768)
        a.copy_(col.to(BF16))

    def image_preprocess(self, images, tmp, out, tab_h, tab_v, lut):
        """Pillow's two-pass fixed-point resampler from the SAME host-built tables (the tables themselves are pinned against
        oracle/image_ref.py and Pillow in tests/test_image_preprocess.py)."""
        self.launches += 1
        cur = images.to(torch.int64)
        for axis, tab in ((2, tab_h), (1, tab_v)):
            if tab is None:
                continue
            bounds, kk, _ = tab
            rows = []
            for o in range(bounds.shape[0]):
                lo, n = int(bounds[o, 0]), int(bounds[o, 1])
                k = kk[o, :n].to(torch.int64)
                sl = cur.narrow(axis, lo, n)
                shape = [1, 1, 1, 1]
                shape[axis] = n
                acc = (sl * k.view(shape)).sum(axis) + (1 << 21)
                rows.append((acc >> 22).clamp(0, 255))
            cur = torch.stack(rows, axis)
        idx = cur.permute(0, 3, 1, 2)                                  # [B, 3, oh, ow]
        out.copy_(torch.stack([lut[c][idx[:, c]] for c in range(3)], 1))

    def barycentric_gather(self, src, rows, weights, out):
        self.launches += 1
        ok = rows[:, 0] >= 0
        r = rows.clamp(min=0).long()
        acc = torch.zeros(out.shape, dtype=torch.float64)
        for v in range(4):
            acc = acc + weights[:, v:v + 1] * src[r[:, v]].double()
        acc[~ok] = float("nan")
        out.copy_(acc.float())

    def beit_assemble(self, patches, cls, x, *, batch, n, c):
        self.launches += 1
        xv = x.view(batch, n + 1, c)
        xv[:, 0] = cls.view(1, c)
        xv[:, 1:] = patches.view(batch, n, c)

    def readout_concat(self, hs, a, *, batch, n, c):
        self.launches += 1
        h = hs.view(batch, n + 1, c)
        a.view(batch, n, 2 * c)[:, :, :c] = h[:, 1:].to(BF16)
        a.view(batch, n, 2 * c)[:, :, c:] = h[:, :1].expand(batch, n, c).to(BF16)

    def pixel_shuffle(self, g, out, *, batch, h, w, c, f):
        self.launches += 1
        t = g.view(batch, h, w, f, f, c).permute(0, 1, 3, 2, 4, 5).reshape(batch, h * f, w * f, c)
        out.view(batch, h * f, w * f, c).copy_(t)

    def im2col3x3_s2(self, x, a, *, batch, h, w, c):
        self.launches += 1
        xn = x.float().view(batch, h, w, c).permute(0, 3, 1, 2)
        col = F.unfold(xn, kernel_size=3, stride=2, padding=1)               # (B, c*9, L) ordered (c, tap)
        L_ = col.shape[-1]
        col = col.view(batch, c, 9, L_).permute(0, 3, 2, 1).reshape(batch * L_, 9 * c)
        a.copy_(col.to(BF16))

    def bilinear_nhwc(self, x, out, *, batch, h, w, c, oh, ow, add=None, out_relu=None):
        self.launches += 1
        xn = x.float().view(batch, h, w, c).permute(0, 3, 1, 2)
        y = F.interpolate(xn, size=(oh, ow), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        if add is not None:
            y = y + add.float().view(batch, oh, ow, c)
        if out is not None:
            out.view(batch, oh, ow, c).copy_(y.to(BF16))
        if out_relu is not None:
            out_relu.view(batch, oh, ow, c).copy_(F.relu(y).to(BF16))

    def relu_bf16(self, x, out):
        self.launches += 1
        out.copy_(F.relu(x.float()).to(BF16))

    def zoe_router_embed(self, conv, e, e_bf16, *, batch, n, c):
        self.launches += 1
        S = n + 1
        pos = torch.arange(0, S, dtype=F32).unsqueeze(1)
        idx = torch.arange(0, c, 2, dtype=F32).unsqueeze(0)
        div = torch.exp(idx * (-torch.log(torch.tensor(10000.0)) / c))
        pe = torch.cat([torch.sin(pos * div), torch.cos(pos * div)], 1)
        ev = e.view(batch, S, c)
        ev[:, 0] = pe[0]
        ev[:, 1:] = conv.view(batch, n, c) + pe[1:]
        if e_bf16 is not None:
            e_bf16.view(batch, S, c).copy_(ev.to(BF16))

    def zoe_attractor(self, attr, prev, out, *, batch, h, w, oh, ow, na, nbins):
        self.launches += 1
        A = F.softplus(attr.float().view(batch, oh, ow, na)).permute(0, 3, 1, 2)
        c = F.interpolate(prev.view(batch, h, w, nbins).permute(0, 3, 1, 2), (oh, ow), mode="bilinear",
                          align_corners=True)
        delta = torch.zeros_like(c)
        for i in range(na):
            dx = A[:, i:i + 1] - c
            delta = delta + dx / (1 + 300.0 * dx * dx)
        out.view(batch, oh, ow, nbins).copy_((c + delta / na).permute(0, 2, 3, 1))

    def zoe_select_head(self, dlog, arena_ptrs, active, head_out, *, forced=-1):
        """arena_ptrs: on the CPU stand-in, a list of the per-head uint8 arenas themselves."""
        self.launches += 1
        head = int(forced) if forced >= 0 else int(torch.argmax(dlog.sum(0)))
        active.copy_(arena_ptrs[head])
        head_out.fill_(head)

    def softplus_f32(self, x, out):
        self.launches += 1
        out.copy_(F.softplus(x.float()).view_as(out))

    def zoe_depth_tail(self, t, e, b1, w2, b2, bins, depth, *, batch, h, w, oh, ow, nh, nbins, min_temp, max_temp):
        self.launches += 1
        eu = F.interpolate(e.float().view(batch, h, w, nh).permute(0, 3, 1, 2), (oh, ow), mode="bilinear",
                           align_corners=True).permute(0, 2, 3, 1)
        hid = F.gelu(t.float().view(batch, oh, ow, nh) + eu + b1)
        o4 = F.softplus(hid @ w2.t() + b2)
        pr = (o4[..., 0] + 1e-4) / (o4[..., 0] + 1e-4 + o4[..., 1] + 1e-4)
        tm = (o4[..., 2] + 1e-4) / (o4[..., 2] + 1e-4 + o4[..., 3] + 1e-4)
        tm = (max_temp - min_temp) * tm + min_temp
        kk = torch.arange(0, nbins, dtype=F32)
        nn_ = torch.tensor(float(nbins - 1)) + 1e-7
        k_ = kk + 1e-7
        lb = nn_ * torch.log(nn_) - k_ * torch.log(k_) - (nn_ - k_) * torch.log(nn_ - k_ + 1e-7)
        lp = torch.log(pr.clamp(1e-4, 1)).unsqueeze(-1)
        lq = torch.log((1 - pr).clamp(1e-4, 1)).unsqueeze(-1)
        y = (lb + kk * lp + (nbins - 1 - kk) * lq) / tm.unsqueeze(-1)
        p = torch.softmax(y, -1)
        c = F.interpolate(bins.view(batch, h, w, nbins).permute(0, 3, 1, 2), (oh, ow), mode="bilinear",
                          align_corners=True).permute(0, 2, 3, 1)
        depth.view(batch, oh, ow).copy_((p * c).sum(-1))

    def zoe_depth_tail_fused(self, x, wa, e, b1, w2, b2, bins, depth, *, batch, h, w, oh, ow, min_temp, max_temp):
        t = x.float() @ wa.float().t()                      # fp32 inside the kernel: no bf16 round trip of the hidden pre-activation
        self.zoe_depth_tail(t, e, b1, w2, b2, bins, depth, batch=batch, h=h, w=w, oh=oh, ow=ow, nh=wa.shape[0],
                            nbins=bins.shape[-1], min_temp=min_temp, max_temp=max_temp)

    @staticmethod
    def zoe_depth_tail_fused_supported(nx, nh, nbins, h, oh):
        return nx == 32 and nh == 40 and nbins == 64 and oh >= 1.4 * h

    def ego3d_encode(self, depth384, intrinsic, xyz, enc, *, n_freqs):
        self.launches += 1
        from oracle.model_ref import backproject_patch, depth_to_224, ego3d_encoding
        p = backproject_patch(intrinsic, depth_to_224(depth384), 14, 2)
        xyz.copy_(p.view_as(xyz))
        e = ego3d_encoding(p, n_freqs).reshape(-1, 12 * (2 * n_freqs + 1))
        enc.zero_()
        enc[:, : e.shape[1]] = e.to(BF16)

    # ---- tokenizer
    def tok_encode(self, actions, edges, nbins_host, ids, *, min_action=-1.0, max_action=1.0, use_spherical=True, trig=None,
                   phi_nonpos=0, phi_neg=0):
        self.launches += 1
        from oracle import tokenizer_ref as T
        pol, nb = _policy_from_flat(edges.numpy(), nbins_host)
        ids.copy_(torch.from_numpy(T.encode(actions.numpy(), pol, nb, min_action, max_action)).to(ids.dtype))

    def tok_decode(self, ids, edges, nbins_host, begin, actions, *, use_spherical=True, center_trig=None):
        self.launches += 1
        from oracle import tokenizer_ref as T
        pol, nb = _policy_from_flat(edges.numpy(), nbins_host)
        actions.copy_(torch.from_numpy(T.decode(ids.numpy() - begin, pol, nb)))


def _policy_from_flat(edges, nbins):
    names = (("translation", "theta_bins"), ("translation", "phi_bins"), ("translation", "r_bins"),
             ("rotation", "roll_bins"), ("rotation", "pitch_bins"), ("rotation", "yaw_bins"))
    pol = {"translation": {}, "rotation": {}}
    nb = {"translation": {}, "rotation": {}, "gripper": int(nbins[6])}
    off = 0
    for (bt, bk), n in zip(names, nbins[:6]):
        pol[bt][bk] = edges[off:off + n + 1]
        nb[bt][bk] = int(n)
        off += n + 1
    return pol, nb
