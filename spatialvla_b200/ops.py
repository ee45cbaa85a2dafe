"""Thin torch-tensor front end of the C ABI: validates shapes/dtypes, passes raw device pointers + the current
CUDA stream.  PyTorch is plumbing only (device memory, streams); every op below runs a hand-written sm_100a
kernel from libspatialvla_b200.so.  There is no CPU path here -- tests inject `oracle.ops_ref.RefOps`, a torch
re-statement with the same method signatures, to validate the host orchestration without a GPU.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import (ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_RELU, ACT_SOFTCAP, ACT_SOFTPLUS,  # noqa: F401
                   GEMM_GEGLU, GEMM_ACCUM_F32, GEMM_CONV3X3)

BF16, F32 = torch.bfloat16, torch.float32


def _ptr(t):
    return None if t is None else t.data_ptr()


def _req(cond, msg):
    if not cond:
        raise ValueError(msg)


def tile_weight(w):
    """[N, K] bf16 -> tile-major [ceil(N/128), ceil(K/64), 128, 64] (zero padded): the decode GEMM streams one contiguous
    16 KB block per pipeline stage instead of 128 row segments of 128 bytes."""
    n, k = w.shape
    nt, kb = (n + 127) // 128, (k + 63) // 64
    p = torch.zeros(nt * 128, kb * 64, dtype=w.dtype, device=w.device)
    p[:n, :k] = w
    return p.view(nt, 128, kb, 64).permute(0, 2, 1, 3).contiguous()


class CudaOps:
    """All methods are asynchronous on torch's current stream."""

    name = "cuda"

    def __init__(self, device="cuda:0", gemm_impl: int = 0):
        if not torch.cuda.is_available():
            raise L.SvlaError("spatialvla_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load_library()
        self.device = torch.device(device)
        self.gemm_impl = gemm_impl    # 0 = tcgen05 (product); 1 = SIMT debugging kernel (tests only)

    # ---- helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def launch_count(self) -> int:
        return int(self.lib.svla_launch_count())

    # ---- G1
    def gemm(self, a, w, *, n=None, k=None, out_bf16=None, out_f32=None, out_relu=None, bias=None, colscale=None,
             res_bf16=None, res2_bf16=None, res_f32=None, res_mod=0, act=ACT_NONE, act_param=0.0, alpha=1.0,
             geglu=False, accumulate=False, conv=None, block_n=0, impl=None, a2=None, w2=None):
        """value = act(alpha * (a @ w.T + a2 @ w2.T) + bias) * colscale + residuals; see include/spatialvla_b200.h (svla_gemm).
        a2 bf16 [M, K2] / w2 bf16 [N, K2]: the K extension (un-merged LoRA term accumulated in the same TMEM tile).
        a: bf16 [M, K] (row stride = lda) or, with conv=(nb,h,w,c), a contiguous NHWC bf16 tensor.
        w: bf16 [N, ldw]. Outputs are caller-allocated [M, >=N] row-major tensors sharing one row stride."""
        g = L.SvlaGemmArgs()
        _req(a.dtype == BF16 and w.dtype == BF16, "gemm: operands must be bf16")
        _req(w.dim() == 2 and w.stride(1) == 1, "gemm: w must be [N, ldw] row-major")
        N = int(n if n is not None else w.shape[0])
        flags = 0
        if conv is not None:
            nb, h, wd, c = conv
            _req(a.is_contiguous() and a.numel() == nb * h * wd * c, "gemm(conv): a must be contiguous NHWC")
            M, K, lda = nb * h * wd, int(w.shape[1]), c
            g.nb, g.h, g.wd, g.c = nb, h, wd, c
            flags |= GEMM_CONV3X3
        else:
            _req(a.dim() == 2 and a.stride(1) == 1, "gemm: a must be [M, lda] row-major")
            M, K, lda = int(a.shape[0]), int(k if k is not None else a.shape[1]), int(a.stride(0))
        if geglu:
            flags |= GEMM_GEGLU
        if accumulate:
            flags |= GEMM_ACCUM_F32
        outs = [t for t in (out_bf16, out_f32, out_relu, res_bf16, res2_bf16, res_f32) if t is not None]
        _req(len(outs) > 0, "gemm: no output")
        ldo = None
        for t in outs:
            _req(t.dim() >= 2 and t.stride(-1) == 1, "gemm: outputs/residuals must be row-major")
            t2 = t if t.dim() == 2 else t.view(-1, t.shape[-1])
            ld = int(t2.stride(0))
            ldo = ld if ldo is None else ldo
            _req(ld == ldo, "gemm: outputs/residuals must share one row stride")
        for t, dt in ((out_bf16, BF16), (out_relu, BF16), (res_bf16, BF16), (res2_bf16, BF16), (out_f32, F32),
                      (res_f32, F32), (bias, F32), (colscale, F32)):
            _req(t is None or t.dtype == dt, "gemm: dtype mismatch in epilogue tensors")
        g.a, g.w = _ptr(a), _ptr(w)
        g.bias, g.colscale = _ptr(bias), _ptr(colscale)
        g.res_bf16, g.res2_bf16, g.res_f32, g.res_mod = _ptr(res_bf16), _ptr(res2_bf16), _ptr(res_f32), int(res_mod)
        g.out_bf16, g.out_f32, g.out_relu_bf16 = _ptr(out_bf16), _ptr(out_f32), _ptr(out_relu)
        g.m, g.n, g.k = M, N, K
        g.lda, g.ldw, g.ldo = lda, int(w.stride(0)), ldo
        g.alpha, g.act_param, g.act, g.flags = float(alpha), float(act_param), int(act), flags
        g.block_n = int(block_n)
        g.impl = int(self.gemm_impl if impl is None else impl)
        if a2 is not None:
            _req(w2 is not None and a2.dtype == BF16 and w2.dtype == BF16 and a2.dim() == 2 and w2.dim() == 2 and a2.stride(1) == 1
                 and w2.stride(1) == 1 and a2.shape[0] == M and w2.shape[0] >= N and a2.shape[1] == w2.shape[1], "gemm: bad K-extension operands")
            g.a2, g.w2, g.k2, g.lda2, g.ldw2 = _ptr(a2), _ptr(w2), int(a2.shape[1]), int(a2.stride(0)), int(w2.stride(0))
        L.check(self.lib.svla_gemm(C.byref(g), self._stream()), "svla_gemm")

    # ---- fine-tune step: training forward pieces and backward kernels (csrc/train_ops.cu, csrc/train_mma.cu)
    def rmsnorm_train_fwd(self, x_in, *, branch=None, w_post=None, w_pre=None, eps=1e-6, x_out=None, h=None):
        _req(x_in.dtype == F32 and x_in.is_contiguous(), "rmsnorm_train_fwd: x_in must be contiguous fp32")
        _req(branch is None or (branch.dtype == F32 and branch.is_contiguous() and x_out is not None and x_out.is_contiguous()), "rmsnorm_train_fwd: branch / x_out")
        rows, cols = x_in.numel() // x_in.shape[-1], x_in.shape[-1]
        L.check(self.lib.svla_rmsnorm_train_fwd(_ptr(x_in), _ptr(branch), _ptr(w_post), _ptr(w_pre), float(eps), rows, cols, _ptr(x_out),
                                                _ptr(h), self._stream()), "svla_rmsnorm_train_fwd")

    def rmsnorm_bwd(self, x, w, dy, *, eps=1e-6, row_idx=None, dx_accum=None, dx_bf16=None):
        """dy bf16 / fp32 [rows, cols]; row_idx int64 [rows] (dy row i <-> x / dx row row_idx[i]) or None."""
        _req(x.dtype == F32 and x.is_contiguous() and dy.is_contiguous() and dy.dtype in (BF16, F32), "rmsnorm_bwd: x fp32, dy bf16 / fp32, contiguous")
        _req(row_idx is None or (row_idx.dtype == torch.int64 and row_idx.is_contiguous() and row_idx.numel() == dy.shape[0]), "rmsnorm_bwd: row_idx")
        for t, dt in ((dx_accum, F32), (dx_bf16, BF16)):
            _req(t is None or (t.dtype == dt and t.is_contiguous() and t.shape[-1] == x.shape[-1]), "rmsnorm_bwd: outputs")
        rows, cols = dy.numel() // dy.shape[-1], dy.shape[-1]
        L.check(self.lib.svla_rmsnorm_bwd(_ptr(x), _ptr(w), _ptr(dy), int(dy.dtype == F32), _ptr(row_idx), float(eps), rows, cols,
                                          _ptr(dx_accum), _ptr(dx_bf16), self._stream()), "svla_rmsnorm_bwd")

    def layernorm_bwd(self, x, gamma, beta, dy, *, eps, relu=False, dx_accum=None, copy_bf16=None, dx_bf16=None):
        _req(x.dtype == F32 and x.is_contiguous() and dy.dtype == BF16 and dy.is_contiguous() and dy.shape == x.shape, "layernorm_bwd: x fp32 / dy bf16")
        rows, cols = x.numel() // x.shape[-1], x.shape[-1]
        L.check(self.lib.svla_layernorm_bwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(dy), float(eps), rows, cols, int(relu), _ptr(dx_accum),
                                            _ptr(copy_bf16), _ptr(dx_bf16), self._stream()), "svla_layernorm_bwd")

    def geglu_fwd(self, gu, act):
        _req(gu.dtype == BF16 and act.dtype == BF16 and gu.is_contiguous() and act.is_contiguous() and gu.shape[1] == 2 * act.shape[1], "geglu_fwd: shapes")
        L.check(self.lib.svla_geglu_fwd(_ptr(gu), _ptr(act), act.shape[0], act.shape[1], self._stream()), "svla_geglu_fwd")

    def geglu_bwd(self, gu, dact, dgu):
        _req(all(t.dtype == BF16 and t.is_contiguous() for t in (gu, dact, dgu)) and gu.shape == dgu.shape and gu.shape[1] == 2 * dact.shape[1], "geglu_bwd: shapes")
        L.check(self.lib.svla_geglu_bwd(_ptr(gu), _ptr(dact), _ptr(dgu), dact.shape[0], dact.shape[1], self._stream()), "svla_geglu_bwd")

    def gelu_tanh_fwd(self, z, f):
        _req(z.dtype == BF16 and f.dtype == BF16 and z.is_contiguous() and f.is_contiguous() and z.shape == f.shape, "gelu_tanh_fwd: shapes")
        L.check(self.lib.svla_gelu_tanh_fwd(_ptr(z), _ptr(f), z.numel(), self._stream()), "svla_gelu_tanh_fwd")

    def gelu_tanh_bwd(self, z, df, dz):
        _req(all(t.dtype == BF16 and t.is_contiguous() and t.shape == z.shape for t in (z, df, dz)), "gelu_tanh_bwd: shapes")
        L.check(self.lib.svla_gelu_tanh_bwd(_ptr(z), _ptr(df), _ptr(dz), z.numel(), self._stream()), "svla_gelu_tanh_bwd")

    def rope_bwd(self, dqkv, *, batch, s, hq, hkv, d, theta):
        _req(dqkv.dtype == BF16 and dqkv.is_contiguous() and tuple(dqkv.shape) == (batch * s, (hq + 2 * hkv) * d), "rope_bwd: dqkv bf16 [B*S, (hq+2hkv)d]")
        L.check(self.lib.svla_rope_bwd(_ptr(dqkv), batch, s, hq, hkv, d, float(theta), self._stream()), "svla_rope_bwd")

    def rows_cast(self, src, out, *, row_idx=None, scale=1.0):
        _req(src.dtype == F32 and src.is_contiguous() and out.dtype == BF16 and out.is_contiguous() and src.shape[-1] == out.shape[-1], "rows_cast: shapes")
        _req(row_idx is None or (row_idx.dtype == torch.int64 and row_idx.is_contiguous() and row_idx.numel() == out.shape[0]), "rows_cast: row_idx")
        L.check(self.lib.svla_rows_cast(_ptr(src), _ptr(row_idx), float(scale), out.shape[0], out.shape[-1], _ptr(out), self._stream()), "svla_rows_cast")

    def lora_pack(self, arena, pool, plan):
        """plan: LoRAStepLayout (desc_table = packed device records, records = the same on the host, total_tiles)."""
        _req(arena.dtype == F32 and arena.is_contiguous() and pool.dtype == BF16 and pool.is_contiguous(), "lora_pack: arena fp32 / pool bf16")
        L.check(self.lib.svla_lora_pack(_ptr(arena), _ptr(pool), _ptr(plan.desc_table), len(plan.records), int(plan.total_tiles), self._stream()),
                "svla_lora_pack")

    def fill_zero(self, t):
        _req(t.is_contiguous(), "fill_zero: contiguous tensor")
        L.check(self.lib.svla_fill_zero(_ptr(t), t.numel() * t.element_size(), self._stream()), "svla_fill_zero")

    def attention_bwd(self, q, k, v, out, dout, dq, dk, dv, *, batch, hq, hkv, sq, sk, d, q_strides, k_strides, v_strides, o_strides,
                      do_strides, dq_strides, dk_strides, dv_strides, scale, softcap=0.0, causal=False, causal_prefix=0, lse=None, window=0):
        """Backward of `attention` (no relpos / kv_start): strides = (batch stride, token stride) in elements, head h at column h*d.
        lse: the forward call's `lse` output -> the tcgen05 sweeps (dQ, dK, dV); None -> warp-MMA kernels that recompute it."""
        a = L.SvlaAttnBwdArgs()
        for t in (q, k, v, out, dout, dq, dk, dv):
            _req(t.dtype == BF16, "attention_bwd: bf16 only")
        a.q, a.k, a.v, a.out, a.dout, a.dq, a.dk, a.dv = (_ptr(t) for t in (q, k, v, out, dout, dq, dk, dv))
        (a.q_bs, a.q_ss), (a.k_bs, a.k_ss), (a.v_bs, a.v_ss), (a.o_bs, a.o_ss) = q_strides, k_strides, v_strides, o_strides
        (a.do_bs, a.do_ss), (a.dq_bs, a.dq_ss), (a.dk_bs, a.dk_ss), (a.dv_bs, a.dv_ss) = do_strides, dq_strides, dk_strides, dv_strides
        sp = int(lse.shape[2]) if lse is not None else sq
        stats = torch.empty((2, batch, hq, sp), dtype=F32, device=self.device)
        a.lse, a.delta = _ptr(stats[0]), _ptr(stats[1])
        if lse is not None:
            _req(lse.dtype == F32 and lse.dim() == 3 and lse.is_contiguous() and lse.shape[0] == batch and lse.shape[1] == hq and sp >= sq,
                 "attention_bwd: lse must be the forward's fp32 [batch, hq, >= sq]")
            a.fwd_lse2, a.lse_stride = _ptr(lse), sp
        a.batch, a.hq, a.hkv, a.sq, a.sk, a.d = batch, hq, hkv, sq, sk, d
        a.scale, a.softcap, a.causal, a.causal_prefix = float(scale), float(softcap or 0.0), int(bool(causal)), int(causal_prefix)
        a.window = int(window or 0)
        L.check(self.lib.svla_attention_bwd(C.byref(a), self._stream()), "svla_attention_bwd")
        return stats

    def gemm_tn(self, s, y, groups, *, r, n, scale=1.0):
        """dst_g += scale * S[:, row0:row0+rows]^T Y[:, col_start::col_stride][:ncols] for each group
        (dst fp32 view [rows, >= ncols] with unit column stride, row0, rows, col_start, col_stride, ncols); up to 4 groups."""
        a = L.SvlaGemmTnArgs()
        _req(s.dtype == BF16 and y.dtype == BF16 and s.dim() == 2 and y.dim() == 2 and s.stride(1) == 1 and y.stride(1) == 1
             and s.shape[0] == y.shape[0], "gemm_tn: bf16 [M, *] operands with one row count")
        _req(1 <= len(groups) <= 4, "gemm_tn: 1..4 groups")
        a.s, a.y, a.m, a.lds, a.ldy = _ptr(s), _ptr(y), int(s.shape[0]), int(s.stride(0)), int(y.stride(0))
        a.r, a.n, a.scale, a.n_groups = int(r), int(n), float(scale), len(groups)
        for i, (dst, row0, rows, c0, cs, nc) in enumerate(groups):
            _req(dst.dtype == F32 and dst.dim() == 2 and dst.stride(1) == 1 and dst.shape[0] == rows and dst.shape[1] >= nc, "gemm_tn: bad group dst")
            g = a.groups[i]
            g.dst, g.ld, g.row0, g.rows, g.col_start, g.col_stride, g.ncols = _ptr(dst), int(dst.stride(0)), int(row0), int(rows), int(c0), int(cs), int(nc)
        L.check(self.lib.svla_gemm_tn(C.byref(a), self._stream()), "svla_gemm_tn")

    # ---- G1s
    def skinny_splits(self, n, k):
        return int(self.lib.svla_gemm_skinny_splits(int(n), int(k)))

    def gemm_skinny(self, x, w, *, out_bf16=None, out_f32=None, bias=None, act=ACT_NONE, act_param=0.0, alpha=1.0,
                    geglu=False, splits=1, tiled_n=None, pair=False):
        """Decode GEMM (M <= 128). splits > 1 (or out_f32 with 3 dims) writes raw fp32 partial sums out_f32[s, M, N].
        tiled_n: w is the tile-major copy made by `tile_weight` of a [tiled_n, K] matrix.
        x of shape [2, M, K] (contiguous) is a hi/lo activation pair (X_HILO, M <= 64): the result is hi @ w.T + lo @ w.T;
        out_bf16 of shape [2, M, N] is written as such a pair (OUT_HILO)."""
        g = L.SvlaSkinnyArgs()
        x_hilo = x.dim() == 3
        if x_hilo:
            _req(x.shape[0] == 2 and x.is_contiguous() and x.shape[1] <= 64, "gemm_skinny: hi/lo activations must be contiguous [2, M <= 64, K]")
            x = x.view(2 * x.shape[1], x.shape[2])
        out_hilo = out_bf16 is not None and out_bf16.dim() == 3
        if out_hilo:
            _req(out_bf16.shape[0] == 2 and out_bf16.is_contiguous(), "gemm_skinny: hi/lo output must be contiguous [2, M, N]")
        _req(x.dtype == BF16 and w.dtype == BF16 and x.dim() == 2 and x.stride(1) == 1, "gemm_skinny: bf16 row-major operands")
        if tiled_n is not None:
            _req(w.dim() == 4 and w.shape[2] == 128 and w.shape[3] == 64 and w.is_contiguous(), "gemm_skinny: tiled w must be [nt, kb, 128, 64]")
            _req(w.shape[0] == (tiled_n + 127) // 128 and w.shape[1] == (x.shape[1] + 63) // 64, "gemm_skinny: tiled w does not match N / K")
        partial = out_f32 is not None and out_f32.dim() == 3
        _req(partial or splits == 1, "gemm_skinny: split-K needs a [splits, M, N] fp32 output")
        g.x, g.w, g.bias = _ptr(x), _ptr(w), _ptr(bias)
        g.out_bf16, g.out_f32 = _ptr(out_bf16), _ptr(out_f32)
        g.m, g.n, g.k = int(x.shape[0]) // (2 if x_hilo else 1), int(w.shape[0] if tiled_n is None else tiled_n), int(x.shape[1])
        g.ldx, g.ldw = int(x.stride(0)), int(w.stride(0) if tiled_n is None else 64)
        o = out_f32 if out_f32 is not None else out_bf16
        g.ldo = int(o.stride(-2))
        g.partial_stride = int(out_f32.stride(0)) if partial else 0
        g.alpha, g.act_param, g.act = float(alpha), float(act_param), int(act)
        g.flags = (1 if geglu else 0) | (2 if partial else 0) | (4 if tiled_n is not None else 0) | (8 if x_hilo else 0) | (16 if out_hilo else 0) | (32 if pair else 0)
        g.splits = int(out_f32.shape[0]) if partial else 1
        L.check(self.lib.svla_gemm_skinny(C.byref(g), self._stream()), "svla_gemm_skinny")

    # ---- G2 / G3
    def attention(self, q, k, v, out, *, batch, hq, hkv, sq, sk, d, q_strides, k_strides, v_strides, o_strides,
                  scale, softcap=0.0, causal=False, relpos_table=None, relpos_win=0, relpos_head_major=False, kv_start=None,
                  causal_prefix=0, lse=None, window=0):
        """strides = (batch stride, token stride) in elements; head h lives at column offset h*d.
        lse: fp32 [batch, hq, >= sq] -- the kernel also stores the log2-domain log-sum-exp of every query row (training forward).
        relpos_table: fp32 [(2*win-1)^2+3, hq] (HF layout) or, with relpos_head_major, its transpose [hq, (2*win-1)^2+3]."""
        a = L.SvlaAttnArgs()
        for t in (q, k, v, out):
            _req(t.dtype == BF16, "attention: bf16 only")
        a.q, a.k, a.v, a.out = _ptr(q), _ptr(k), _ptr(v), _ptr(out)
        a.q_bs, a.q_ss = q_strides
        a.k_bs, a.k_ss = k_strides
        a.v_bs, a.v_ss = v_strides
        a.o_bs, a.o_ss = o_strides
        a.batch, a.hq, a.hkv, a.sq, a.sk, a.d = batch, hq, hkv, sq, sk, d
        a.scale, a.softcap, a.causal = float(scale), float(softcap or 0.0), int(bool(causal))
        a.window = int(window or 0)       # sliding-window layer: key slot j masked for query slot i when i - j >= window
        _req(relpos_table is None or relpos_table.dtype == F32, "attention: relpos table must be fp32")
        _req(relpos_table is None or relpos_table.is_contiguous(), "attention: relpos table must be contiguous")
        a.relpos_table, a.relpos_win, a.relpos_head_major = _ptr(relpos_table), int(relpos_win), int(bool(relpos_head_major))
        _req(kv_start is None or (kv_start.dtype == torch.int32 and kv_start.is_contiguous() and kv_start.numel() == batch),
             "attention: kv_start must be int32 [batch]")
        a.kv_start = _ptr(kv_start)
        _req(causal_prefix == 0 or (causal and 0 < causal_prefix <= sk), "attention: causal_prefix needs causal=True and 0 < prefix <= sk")
        a.causal_prefix = int(causal_prefix)       # prefix-LM: keys < causal_prefix visible to every query
        if lse is not None:
            _req(lse.dtype == F32 and lse.dim() == 3 and lse.is_contiguous() and lse.shape[0] == batch and lse.shape[1] == hq and lse.shape[2] >= sq,
                 "attention: lse must be fp32 [batch, hq, >= sq]")
            a.lse, a.lse_stride = _ptr(lse), int(lse.shape[2])
        L.check(self.lib.svla_attention(C.byref(a), self._stream()), "svla_attention")

    def decode_attention(self, q, kcache, vcache, out, *, batch, hq, hkv, d, smax, ctx, scale, softcap=0.0, kv_start=None):
        L.check(self.lib.svla_decode_attention(_ptr(q), _ptr(kcache), _ptr(vcache), _ptr(out), batch, hq, hkv, d,
                                               smax, ctx, float(scale), float(softcap or 0.0), _ptr(kv_start), self._stream()),
                "svla_decode_attention")

    def decode_attention_fused(self, qkv_partials, kcache, vcache, out, *, batch, hq, hkv, d, smax, ctx, theta, scale, softcap=0.0,
                               kv_start=None, window=0):
        """RoPE + KV-cache append + attention of one decode step; qkv_partials fp32 [splits, batch, (hq+2hkv)*d].
        out bf16 [batch, hq*d], or a contiguous hi/lo pair [2, batch, hq*d]; window > 0: only the last `window` cache slots."""
        _req(qkv_partials.dtype == F32 and qkv_partials.dim() == 3 and qkv_partials.stride(2) == 1 and
             qkv_partials.stride(1) == qkv_partials.shape[2], "decode_attention_fused: qkv must be fp32 [splits, batch, W]")
        hi, lo = out, None
        if out.dim() == 3:
            _req(out.shape[0] == 2 and out.is_contiguous(), "decode_attention_fused: hi/lo output must be contiguous [2, batch, hq*d]")
            hi, lo = out[0], out[1]
        L.check(self.lib.svla_decode_attention_fused_ex(_ptr(qkv_partials), int(qkv_partials.shape[0]), int(qkv_partials.stride(0)),
                                                        _ptr(kcache), _ptr(vcache), _ptr(hi), _ptr(lo), batch, hq, hkv, d, smax, ctx,
                                                        float(theta), float(scale), float(softcap or 0.0), _ptr(kv_start),
                                                        int(window or 0), self._stream()), "svla_decode_attention_fused_ex")

    # ---- G4 persistent small-batch decode step
    # ---- persistent tensor-core decode step (csrc/decode_mega.cu)
    def decode_mega_supported(self, batch, hidden, hq, hkv, d, ff, ctx):
        return bool(self.lib.svla_decode_mega_supported(batch, hidden, hq, hkv, d, ff, ctx))

    def decode_mega_plan(self, layer_weights, norm_weights, *, hidden, hq, hkv, d, ff):
        """layer_weights: per layer (wqkv, wo, wgu, wd) bf16 row-major device tensors; norm_weights: per layer (ln_in,
        ln_post_attn, ln_pre_ff, ln_post_ff) fp32 device tensors.  Returns the plan dict `decode_mega_step` takes: device
        copies of the TMA descriptors and of the norm-pointer table, plus the zero-initialised scratch buffer."""
        n_layers = len(layer_weights)
        for ws in layer_weights:
            for w in ws:
                _req(w.dtype == BF16 and w.is_contiguous() and w.dim() == 2, "decode_mega_plan: weights must be contiguous bf16 matrices")
        scratch = torch.zeros(int(self.lib.svla_decode_mega_scratch_bytes(hidden, hq, hkv, d, ff)), dtype=torch.uint8, device=self.device)
        nbytes = int(self.lib.svla_decode_mega_maps_bytes(n_layers))
        host = (C.c_uint8 * nbytes)()
        wptrs = (C.c_void_p * (4 * n_layers))(*[w.data_ptr() for ws in layer_weights for w in ws])
        L.check(self.lib.svla_decode_mega_plan(C.cast(host, C.c_void_p), C.cast(wptrs, C.c_void_p), n_layers, hidden,
                                               hq, hkv, d, ff, _ptr(scratch)), "svla_decode_mega_plan")
        maps = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(self.device)
        _req(maps.data_ptr() % 64 == 0, "decode_mega_plan: descriptor table is not 64-byte aligned")
        for ns in norm_weights:
            for n in ns:
                _req(n.dtype == F32 and n.is_contiguous(), "decode_mega_plan: norm weights must be contiguous fp32")
        norm_tab = torch.tensor([n.data_ptr() for ns in norm_weights for n in ns], dtype=torch.int64).to(self.device)
        return {"maps": maps, "norm_tab": norm_tab, "scratch": scratch, "n_layers": n_layers, "keep": (layer_weights, norm_weights),
                "dims": (hidden, hq, hkv, d, ff)}

    def decode_mega_step(self, plan, x, final_w, h_out, kcache, vcache, *, batch, smax, ctx, theta, scale, softcap, eps, kv_start=None):
        """One decode step of every layer in one launch. kcache/vcache: bf16 [layers, batch, smax, hkv, d] contiguous."""
        hidden, hq, hkv, d, ff = plan["dims"]
        _req(x.dtype == F32 and x.is_contiguous() and tuple(x.shape) == (batch, hidden), "decode_mega_step: x must be fp32 [batch, hidden]")
        _req(h_out.dtype == BF16 and h_out.is_contiguous() and h_out.numel() == batch * hidden, "decode_mega_step: h_out must be bf16 [batch, hidden]")
        _req(kcache.dtype == BF16 and kcache.is_contiguous() and vcache.is_contiguous() and tuple(kcache.shape) == (plan["n_layers"], batch, smax, hkv, d)
             and vcache.shape == kcache.shape, "decode_mega_step: cache must be bf16 [layers, batch, smax, hkv, d]")
        L.check(self.lib.svla_decode_mega_step(_ptr(plan["maps"]), _ptr(plan["norm_tab"]), plan["n_layers"], _ptr(x), _ptr(final_w), _ptr(h_out),
                                               _ptr(kcache), _ptr(vcache), int(kcache.stride(0)), _ptr(plan["scratch"]), batch, hidden, hq, hkv, d,
                                               ff, smax, ctx, float(theta), float(scale), float(softcap or 0.0), float(eps), _ptr(kv_start),
                                               self._stream()), "svla_decode_mega_step")

    # ---- memory-bound fused ops
    def layernorm(self, x, gamma, beta, eps, *, out_bf16=None, out_f32=None, relu=False):
        _req(x.dtype == F32 and x.is_contiguous(), "layernorm: x must be contiguous fp32")
        rows, cols = x.numel() // x.shape[-1], x.shape[-1]
        L.check(self.lib.svla_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), float(eps), rows, cols, _ptr(out_bf16),
                                        _ptr(out_f32), int(relu), self._stream()), "svla_layernorm")

    def rmsnorm_residual(self, x, *, branch=None, w_post=None, w_pre=None, eps=1e-6, out_bf16=None):
        """branch may be [rows, cols] or split-K partial sums [splits, rows, cols] (summed while read)."""
        _req(x.dtype == F32 and x.is_contiguous(), "rmsnorm_residual: x must be contiguous fp32")
        rows, cols = x.numel() // x.shape[-1], x.shape[-1]
        npart, pstride = 1, 0
        if branch is not None and branch.dim() == 3:
            npart, pstride = int(branch.shape[0]), int(branch.stride(0))
        if out_bf16 is not None and out_bf16.dim() == 3 and out_bf16.shape[0] == 2 and out_bf16.shape[1] == rows:
            # hi/lo activation pair [2, rows, cols] of the decode chain
            _req(out_bf16.is_contiguous(), "rmsnorm_residual: hi/lo output must be contiguous [2, rows, cols]")
            L.check(self.lib.svla_rmsnorm_residual_hilo(_ptr(x), _ptr(branch), _ptr(w_post), _ptr(w_pre), float(eps), rows, cols,
                                                        _ptr(out_bf16[0]), _ptr(out_bf16[1]), npart, pstride, self._stream()),
                    "svla_rmsnorm_residual_hilo")
            return
        L.check(self.lib.svla_rmsnorm_residual(_ptr(x), _ptr(branch), _ptr(w_post), _ptr(w_pre), float(eps), rows,
                                               cols, _ptr(out_bf16), npart, pstride, self._stream()), "svla_rmsnorm_residual")

    def rope_kv(self, qkv, q_out, kcache, vcache, *, batch, s, hq, hkv, d, smax, pos0, theta, row_pads=None):
        """qkv: bf16 [tokens, W] or fp32 split-K partial sums [splits, tokens, W]."""
        if qkv.dtype == F32:
            _req(qkv.dim() == 3, "rope_kv: fp32 qkv must be [splits, tokens, W]")
            L.check(self.lib.svla_rope_kv(None, _ptr(q_out), _ptr(kcache), _ptr(vcache), batch, s, hq, hkv, d, smax, pos0,
                                          float(theta), _ptr(qkv), int(qkv.shape[0]), int(qkv.stride(0)), _ptr(row_pads), self._stream()),
                    "svla_rope_kv")
            return
        L.check(self.lib.svla_rope_kv(_ptr(qkv), _ptr(q_out), _ptr(kcache), _ptr(vcache), batch, s, hq, hkv, d, smax,
                                      pos0, float(theta), None, 1, 0, _ptr(row_pads), self._stream()), "svla_rope_kv")

    def embed_tokens(self, ids, embed, spatial_embed, image_feats, x, *, image_token, act_lo, n_act, n_img,
                     normalizer, status):
        _req(ids.dtype == torch.int64 and ids.is_contiguous(), "embed_tokens: ids must be contiguous int64")
        B, S = ids.shape
        L.check(self.lib.svla_embed_tokens(_ptr(ids), _ptr(embed), _ptr(spatial_embed), _ptr(image_feats), _ptr(x), B,
                                           S, x.shape[-1], embed.shape[0], int(image_token), int(act_lo), int(n_act),
                                           int(n_img), float(normalizer), _ptr(status), self._stream()),
                "svla_embed_tokens")

    def argmax_rows(self, logits, out_ids, *, id_offset=0):
        """out_ids: int64 view with one element per row (any stride)."""
        _req(logits.dtype == F32 and logits.dim() == 2 and logits.stride(1) == 1, "argmax_rows: bad logits")
        _req(out_ids.dtype == torch.int64 and out_ids.dim() == 1, "argmax_rows: out_ids must be 1-D int64")
        L.check(self.lib.svla_argmax_rows(_ptr(logits), logits.shape[0], logits.shape[1], logits.stride(0),
                                          int(id_offset), _ptr(out_ids), out_ids.stride(0), self._stream()),
                "svla_argmax_rows")

    def cross_entropy_rows(self, logits, labels, row_loss, row_argmax, *, row_offset=0, summary=None, ignore_index=-100):
        """logits fp32 [rows, cols] = entries [row_offset, row_offset+rows) of labels int64 [n] / row_loss fp32 [n] / row_argmax
        int64 [n]; summary fp32 [3] (last chunk only) <- mean loss over the non-ignored entries so far, their count, argmax hits."""
        _req(logits.dtype == F32 and logits.dim() == 2 and logits.stride(1) == 1, "cross_entropy_rows: bad logits")
        n = labels.shape[0]
        _req(labels.dtype == torch.int64 and labels.dim() == 1 and labels.is_contiguous(), "cross_entropy_rows: labels must be contiguous int64")
        _req(row_loss.dtype == F32 and row_loss.is_contiguous() and row_loss.shape == (n,), "cross_entropy_rows: row_loss fp32 [n]")
        _req(row_argmax.dtype == torch.int64 and row_argmax.is_contiguous() and row_argmax.shape == (n,), "cross_entropy_rows: row_argmax int64 [n]")
        _req(0 <= row_offset and row_offset + logits.shape[0] <= n, "cross_entropy_rows: chunk outside the label array")
        _req(summary is None or (summary.dtype == F32 and summary.numel() == 3), "cross_entropy_rows: summary fp32 [3]")
        L.check(self.lib.svla_cross_entropy_rows(_ptr(logits), logits.shape[0], logits.shape[1], logits.stride(0), _ptr(labels),
                                                 int(ignore_index), _ptr(row_loss), _ptr(row_argmax), int(row_offset), _ptr(summary),
                                                 self._stream()),
                "svla_cross_entropy_rows")

    def cross_entropy_bwd(self, logits, labels, row_loss, summary, dz, *, row_offset=0, softcap=0.0, ignore_index=-100):
        """dz bf16 [rows, ldo >= cols, even] <- d(mean CE)/d(pre-softcap logits) for the chunk logits fp32 [rows, cols] (entries
        [row_offset, +rows) of labels / row_loss from cross_entropy_rows; summary = its whole-batch summary)."""
        _req(logits.dtype == F32 and logits.dim() == 2 and logits.stride(1) == 1, "cross_entropy_bwd: bad logits")
        _req(dz.dtype == BF16 and dz.dim() == 2 and dz.stride(1) == 1 and dz.shape[0] == logits.shape[0]
             and dz.shape[1] >= logits.shape[1] and dz.shape[1] % 2 == 0 and dz.stride(0) == dz.shape[1], "cross_entropy_bwd: bad dz")
        _req(labels.dtype == torch.int64 and labels.is_contiguous() and row_loss.dtype == F32 and row_loss.is_contiguous()
             and row_loss.shape == labels.shape and row_offset + logits.shape[0] <= labels.shape[0], "cross_entropy_bwd: labels / row_loss")
        _req(summary.dtype == F32 and summary.numel() == 3, "cross_entropy_bwd: summary fp32 [3]")
        L.check(self.lib.svla_cross_entropy_bwd(_ptr(logits), logits.shape[0], logits.shape[1], logits.stride(0), _ptr(labels),
                                                int(ignore_index), _ptr(row_loss), int(row_offset), _ptr(summary), float(softcap or 0.0),
                                                _ptr(dz), dz.shape[1], self._stream()),
                "svla_cross_entropy_bwd")

    def adamw_step(self, param, grad, exp_avg, exp_avg_sq, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step, grad_scale=1.0,
                   sumsq=None, max_grad_norm=0.0):
        """In-place AdamW on flat fp32 buffers (torch.optim.AdamW arithmetic); step counts from 1.  sumsq: device fp32 scalar holding
        sum(grad^2) (ops.sumsq) -> clip_grad_norm_(max_grad_norm) of the scaled gradient is applied inside the kernel."""
        for t in (param, grad, exp_avg, exp_avg_sq):
            _req(t.dtype == F32 and t.dim() == 1 and t.is_contiguous() and t.numel() == param.numel(), "adamw_step: flat fp32 buffers of one size")
        _req(sumsq is None or (sumsq.dtype == F32 and sumsq.numel() == 1), "adamw_step: sumsq fp32 scalar")
        L.check(self.lib.svla_adamw_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr), float(beta1),
                                         float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), _ptr(sumsq),
                                         float(max_grad_norm), self._stream()),
                "svla_adamw_step")

    def sumsq(self, x, out):
        _req(x.dtype == F32 and x.is_contiguous() and out.dtype == F32 and out.numel() == 1, "sumsq: fp32 buffer / scalar")
        L.check(self.lib.svla_sumsq(_ptr(x), x.numel(), _ptr(out), self._stream()), "svla_sumsq")

    def siglip_patchify(self, px, a):
        _req(px.dtype == F32 and px.is_contiguous() and tuple(px.shape[1:]) == (3, 224, 224), "siglip_patchify: px")
        L.check(self.lib.svla_siglip_patchify(_ptr(px), _ptr(a), px.shape[0], a.shape[1], self._stream()),
                "svla_siglip_patchify")

    def zoe_patchify(self, px, a):
        _req(px.dtype == F32 and px.is_contiguous() and tuple(px.shape[1:]) == (3, 224, 224), "zoe_patchify: px")
        L.check(self.lib.svla_zoe_patchify(_ptr(px), _ptr(a), px.shape[0], self._stream()), "svla_zoe_patchify")

    def image_preprocess(self, images, tmp, out, tab_h, tab_v, lut):
        """images uint8 [B, H, W, 3] -> out fp32 [B, 3, oh, ow]; tab_* = (bounds int32 [o, 2], kk int32 [o, ksize], ksize) or None
        when that axis already has the output size; tmp uint8 [B, H, ow, 3] (horizontal pass result) or None."""
        _req(images.dtype == torch.uint8 and images.is_contiguous() and images.dim() == 4 and images.shape[-1] == 3, "image_preprocess: uint8 [B,H,W,3]")
        _req(out.dtype == F32 and out.is_contiguous() and lut.dtype == F32 and lut.numel() == 768, "image_preprocess: out / lut")
        B, H, W, _ = images.shape
        bh, kh, ksh = tab_h if tab_h is not None else (None, None, 0)
        bv, kv, ksv = tab_v if tab_v is not None else (None, None, 0)
        L.check(self.lib.svla_image_preprocess(_ptr(images), B, H, W, _ptr(tmp), _ptr(out), out.shape[2], out.shape[3], _ptr(bh), _ptr(kh),
                                               int(ksh), _ptr(bv), _ptr(kv), int(ksv), _ptr(lut), self._stream()), "svla_image_preprocess")

    def barycentric_gather(self, src, rows, weights, out):
        """out[t] = sum_v weights[t, v] * src[rows[t, v]] (fp64 accumulate); rows int32 [T, 4] (-1 row -> NaN), weights fp64 [T, 4]."""
        _req(src.dtype == F32 and src.is_contiguous() and out.dtype == F32 and out.is_contiguous() and src.shape[1] == out.shape[1],
             "barycentric_gather: fp32 src / out of one width")
        _req(rows.dtype == torch.int32 and rows.is_contiguous() and tuple(rows.shape) == (out.shape[0], 4) and
             weights.dtype == torch.float64 and weights.is_contiguous() and tuple(weights.shape) == (out.shape[0], 4), "barycentric_gather: rows / weights")
        L.check(self.lib.svla_barycentric_gather(_ptr(src), src.shape[0], _ptr(rows), _ptr(weights), _ptr(out), out.shape[0], out.shape[1],
                                                 self._stream()), "svla_barycentric_gather")

    def beit_assemble(self, patches, cls, x, *, batch, n, c):
        L.check(self.lib.svla_beit_assemble(_ptr(patches), _ptr(cls), _ptr(x), batch, n, c, self._stream()),
                "svla_beit_assemble")

    def readout_concat(self, hs, a, *, batch, n, c):
        L.check(self.lib.svla_readout_concat(_ptr(hs), _ptr(a), batch, n, c, self._stream()), "svla_readout_concat")

    def pixel_shuffle(self, g, out, *, batch, h, w, c, f):
        L.check(self.lib.svla_pixel_shuffle(_ptr(g), _ptr(out), batch, h, w, c, f, self._stream()),
                "svla_pixel_shuffle")

    def im2col3x3_s2(self, x, a, *, batch, h, w, c):
        L.check(self.lib.svla_im2col3x3_s2(_ptr(x), _ptr(a), batch, h, w, c, self._stream()), "svla_im2col3x3_s2")

    def bilinear_nhwc(self, x, out, *, batch, h, w, c, oh, ow, add=None, out_relu=None):
        L.check(self.lib.svla_bilinear_nhwc(_ptr(x), _ptr(add), _ptr(out), _ptr(out_relu), batch, h, w, c, oh, ow,
                                            self._stream()), "svla_bilinear_nhwc")

    def relu_bf16(self, x, out):
        L.check(self.lib.svla_relu_bf16(_ptr(x), _ptr(out), x.numel(), self._stream()), "svla_relu_bf16")

    def zoe_router_embed(self, conv, e, e_bf16, *, batch, n, c):
        L.check(self.lib.svla_zoe_router_embed(_ptr(conv), _ptr(e), _ptr(e_bf16), batch, n, c, self._stream()),
                "svla_zoe_router_embed")

    def zoe_attractor(self, attr, prev, out, *, batch, h, w, oh, ow, na, nbins):
        L.check(self.lib.svla_zoe_attractor(_ptr(attr), _ptr(prev), _ptr(out), batch, h, w, oh, ow, na, nbins,
                                            self._stream()), "svla_zoe_attractor")

    def zoe_select_head(self, dlog, arena_ptrs, active, head_out, *, forced=-1):
        """dlog fp32 [B, n_heads]; arena_ptrs int64 device tensor [n_heads] of device pointers to the per-head byte arenas;
        active uint8 [bytes]; head_out int32 [1].  Device-side argmax of the batch-summed logits + copy of that head's arena."""
        _req(dlog.dtype == F32 and dlog.dim() == 2 and dlog.is_contiguous(), "zoe_select_head: dlog fp32 [B, n_heads]")
        _req(arena_ptrs.dtype == torch.int64 and arena_ptrs.numel() == dlog.shape[1] and active.dtype == torch.uint8, "zoe_select_head: arenas")
        L.check(self.lib.svla_zoe_select_head(_ptr(dlog), int(dlog.shape[0]), int(dlog.shape[1]), int(forced), _ptr(arena_ptrs), _ptr(active),
                                              int(active.numel()), _ptr(head_out), self._stream()), "svla_zoe_select_head")

    def softplus_f32(self, x, out):
        L.check(self.lib.svla_softplus_f32(_ptr(x), _ptr(out), x.numel(), self._stream()), "svla_softplus_f32")

    def zoe_depth_tail(self, t, e, b1, w2, b2, bins, depth, *, batch, h, w, oh, ow, nh, nbins, min_temp, max_temp):
        L.check(self.lib.svla_zoe_depth_tail(_ptr(t), _ptr(e), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(bins), _ptr(depth),
                                             batch, h, w, oh, ow, nh, nbins, float(min_temp), float(max_temp),
                                             self._stream()), "svla_zoe_depth_tail")

    def zoe_depth_tail_fused(self, x, wa, e, b1, w2, b2, bins, depth, *, batch, h, w, oh, ow, min_temp, max_temp):
        """x bf16 [B*oh*ow, 32] relative-head features, wa bf16 [40, 32]: the first CLB layer runs inside the kernel."""
        _req(x.dtype == BF16 and wa.dtype == BF16 and x.is_contiguous() and wa.is_contiguous(), "zoe_depth_tail_fused: bf16 contiguous x / wa")
        L.check(self.lib.svla_zoe_depth_tail_fused(_ptr(x), _ptr(wa), _ptr(e), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(bins), _ptr(depth),
                                                   batch, h, w, oh, ow, int(x.shape[-1]), int(wa.shape[0]), int(bins.shape[-1]),
                                                   float(min_temp), float(max_temp), self._stream()), "svla_zoe_depth_tail_fused")

    @staticmethod
    def zoe_depth_tail_fused_supported(nx, nh, nbins, h, oh):
        return nx == 32 and nh == 40 and nbins == 64 and oh >= 1.4 * h

    def ego3d_encode(self, depth384, intrinsic, xyz, enc, *, n_freqs):
        _req(depth384.dtype == F32 and depth384.is_contiguous() and tuple(depth384.shape[1:]) == (384, 384),
             "ego3d_encode: depth must be fp32 [B,384,384]")
        _req(intrinsic.dtype == F32 and intrinsic.is_contiguous(), "ego3d_encode: intrinsic must be fp32")
        k_stride = 9 if intrinsic.dim() == 3 else 0
        _req(k_stride == 0 or intrinsic.shape[0] == depth384.shape[0], "ego3d_encode: per-sample K batch mismatch")
        L.check(self.lib.svla_ego3d_encode(_ptr(depth384), _ptr(intrinsic), k_stride, _ptr(xyz), _ptr(enc),
                                           depth384.shape[0], enc.shape[1], n_freqs, self._stream()),
                "svla_ego3d_encode")

    # ---- M8 tokenizer (device buffers)
    def tok_encode(self, actions, edges, nbins_host, ids, *, min_action=-1.0, max_action=1.0, use_spherical=True, trig=None,
                   phi_nonpos=0, phi_neg=0):
        """trig: float64 [n_theta_interior + n_phi_interior, 4] device table of action_tokenizer.edge_trig_table (exact angular
        binning); None = the library atan2 path (within a few ulp of an edge it may differ from glibc)."""
        nb = (C.c_int32 * 7)(*nbins_host)
        _req(trig is None or (trig.dtype == torch.float64 and trig.is_contiguous()
                              and trig.numel() == 4 * (nbins_host[0] - 1 + nbins_host[1] - 1)), "tok_encode: bad trig table")
        L.check(self.lib.svla_tok_encode(_ptr(actions), _ptr(edges), C.cast(nb, C.c_void_p), _ptr(ids),
                                         actions.shape[0], float(min_action), float(max_action), int(use_spherical),
                                         _ptr(trig), int(phi_nonpos), int(phi_neg), self._stream()), "svla_tok_encode")

    def tok_decode(self, ids, edges, nbins_host, begin, actions, *, use_spherical=True, center_trig=None):
        """center_trig: float64 [theta bins + phi bins, 2] device table of (sin, cos) at the bin centres (host libm); None = sincos."""
        nb = (C.c_int32 * 7)(*nbins_host)
        _req(center_trig is None or (center_trig.dtype == torch.float64 and center_trig.is_contiguous()
                                     and center_trig.numel() == 2 * (nbins_host[0] + nbins_host[1])), "tok_decode: bad centre table")
        L.check(self.lib.svla_tok_decode(_ptr(ids), _ptr(edges), C.cast(nb, C.c_void_p), int(begin), _ptr(actions),
                                         ids.shape[0], int(use_spherical), _ptr(center_trig), self._stream()), "svla_tok_decode")
