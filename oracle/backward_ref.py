"""TEST INFRASTRUCTURE ONLY -- closed-form backward restatements of the memory-bound Gemma2 pieces, the checkers of the backward
kernels of the LoRA step (none of these kernels exists yet).  Each function is the formula a CUDA kernel will
implement, written out (no autograd inside), and tests/test_backward_ref.py pins every one of them on torch autograd through the
forward restatement in oracle/model_ref.py -- which itself is pinned on the live reference's loss.backward()
(tests/golden/tiny_model_train_grads.npz).  Reference forward lines cited per function."""
from __future__ import annotations

import math

import torch


def rmsnorm_bwd(x, w, dy, eps):
    """Gemma2RMSNorm (model/modeling_gemma2.py:60-77): y = x * rsqrt(mean(x^2) + eps) * (1 + w), statistics in fp32.
    Returns (dx, dw): dx = r * (g - x * r^2 * mean(g * x)) with g = dy * (1 + w), r = rsqrt(mean(x^2) + eps); dw = sum_rows dy * x * r."""
    x, dy = x.float(), dy.float()
    r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    g = dy * (1.0 + w.float())
    dx = r * (g - x * (r * r) * (g * x).mean(-1, keepdim=True))
    dw = (dy * x * r).reshape(-1, x.shape[-1]).sum(0)
    return dx, dw


def sandwich_bwd(x, branch, w_post, dx_out, eps):
    """Residual update of a Gemma2 decoder layer (model/modeling_gemma2.py:475-496): x_out = x + rms(branch; w_post).
    Given d(x_out) returns (d x, d branch, d w_post)."""
    dbranch, dw = rmsnorm_bwd(branch, w_post, dx_out, eps)
    return dx_out.float(), dbranch, dw


def gelu_tanh_grad(z):
    """d/dz of gelu_pytorch_tanh (model/modeling_gemma2.py:80-92 activation): 0.5 (1 + t) + 0.5 z (1 - t^2) u'(z),
    t = tanh(u), u = sqrt(2/pi) (z + 0.044715 z^3)."""
    k = math.sqrt(2.0 / math.pi)
    u = k * (z + 0.044715 * z.pow(3))
    t = torch.tanh(u)
    return 0.5 * (1.0 + t) + 0.5 * z * (1.0 - t * t) * k * (1.0 + 3 * 0.044715 * z * z)


def geglu_bwd(gate, up, dact):
    """act = gelu_tanh(gate) * up (model/modeling_gemma2.py:91-92) -> (d gate, d up)"""
    gate, up, dact = gate.float(), up.float(), dact.float()
    return dact * up * gelu_tanh_grad(gate), dact * torch.nn.functional.gelu(gate, approximate="tanh")


def rope_bwd(dy, pos, theta):
    """apply_rotary_pos_emb (model/modeling_gemma2.py:140-154): y = x cos + rotate_half(x) sin is an orthogonal map per (position,
    frequency) pair, so dx = dy cos - rotate_half(dy) sin (rotation by the negative angle).  dy [..., S, D], pos [S] or [B, S]."""
    d = dy.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    fr = pos.float()[..., None] * inv
    emb = torch.cat([fr, fr], -1)
    cos, sin = (emb.cos()[None, None], emb.sin()[None, None]) if pos.dim() == 1 else (emb.cos()[:, None], emb.sin()[:, None])
    d1, d2 = dy[..., : d // 2], dy[..., d // 2:]
    return dy * cos - torch.cat([-d2, d1], -1) * sin


def softcap_attention_bwd(q, k, v, do, scale, cap, mask):
    """Eager Gemma2 attention (model/modeling_gemma2.py:169-195): s = q k^T * scale, c = cap tanh(s / cap), p = softmax(c + mask),
    o = p v.  q [B,H,Sq,D], k/v [B,H,Sk,D] (already repeated for GQA), mask bool [Sq,Sk] (True = masked) or None.
    Returns (dq, dk, dv):  dv = p^T do;  dp = do v^T;  dc = p (dp - sum(dp p));  ds = dc (1 - (c / cap)^2);  dq = ds k scale;
    dk = ds^T q scale -- the flash-attention backward recurrence with one extra factor for the soft-cap."""
    q, k, v, do = q.float(), k.float(), v.float(), do.float()
    s = (q @ k.transpose(-1, -2)) * scale
    c = cap * torch.tanh(s / cap) if cap else s
    if mask is not None:
        c = c.masked_fill(mask, float("-inf"))
    p = torch.softmax(c, -1)
    dv = p.transpose(-1, -2) @ do
    dp = do @ v.transpose(-1, -2)
    dc = p * (dp - (dp * p).sum(-1, keepdim=True))
    if cap:
        cc = torch.where(torch.isinf(c), torch.zeros_like(c), c)
        dc = dc * (1.0 - (cc / cap) ** 2)
    dq = (dc @ k) * scale
    dk = (dc.transpose(-1, -2) @ q) * scale
    return dq, dk, dv


def lora_linear_bwd(x, dy, w, A, B, s):
    """y = x W^T + s (x A^T) B^T (PEFT LoRA Linear, train/spatialvla_finetune.py:262-302) with frozen W:
    dx = dy W + s (dy B) A;  gA = s (dy B)^T x;  gB = s dy^T (x A^T).  The base dW is never formed."""
    x2, dy2 = x.float().reshape(-1, x.shape[-1]), dy.float().reshape(-1, dy.shape[-1])
    t = dy2 @ B.float()                      # [rows, r]
    dx = dy2 @ w.float() + s * (t @ A.float())
    gA = s * (t.t() @ x2)
    gB = s * (dy2.t() @ (x2 @ A.float().t()))
    return dx.reshape(x.shape), gA, gB
