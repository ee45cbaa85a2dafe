/*
 * spatialvla_b200 -- C ABI of the B200 (sm_100a) kernels behind the SpatialVLA `predict_action` path.
 *
 * The reference (tomputer-g/SpatialVLA) is pure Python on top of torch/ATen; it has no FFI of its own.  The
 * drop-in boundary is therefore its HuggingFace-facing Python API (SpatialVLAProcessor,
 * SpatialVLAForConditionalGeneration.predict_action, SpatialActionTokenizer), mirrored in
 * spatialvla_b200/*.py, and THIS header is the thin C ABI those classes call through ctypes: every entry
 * point replaces the torch op sequence of the reference lines cited next to it.
 *
 * Conventions: all pointers are DEVICE pointers owned by the caller unless the name ends in `_host`;
 * every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 * synchronisation and returns 0 on success or a negative code (`svla_last_error()` gives the text).
 * bf16 = raw uint16 storage.  Row-major everywhere; images/feature maps are NHWC.
 */
#ifndef SPATIALVLA_B200_H
#define SPATIALVLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVLA_ABI_VERSION 3

const char* svla_last_error(void);
int svla_abi_version(void);
/* Number of kernels launched by this library since load (bench.py's `gpu_launches`). */
long long svla_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * G1: D = epilogue(A[M,K] * W[N,K]^T)  -- tcgen05.mma (cta_group::1, 128xBNx16, fp32 accumulators in TMEM),
 * operands staged by TMA (SWIZZLE_128B), warp-specialised persistent kernel.
 * Replaces every nn.Linear / 1x1 conv / patch-embed conv / 3x3 conv of the path:
 *   model/modeling_gemma2.py:80-92,351-354,993 ; HF siglip/modeling_siglip.py:124-130,252-327 ;
 *   HF beit/modeling_beit.py:209,225-306,400-440 ; HF zoedepth/modeling_zoedepth.py:129-149,199-216,301-303,
 *   349-352 ; model/modeling_spatialvla.py:59-64,121-128.
 * value = act(alpha * acc + bias[n]) * colscale[n] + res_bf16[m,n] + res2_bf16[m,n] + res_f32[res_mod ? m % res_mod : m, n]
 * --------------------------------------------------------------------------------------------------- */
enum {
  SVLA_ACT_NONE = 0, SVLA_ACT_GELU_TANH = 1, SVLA_ACT_GELU_ERF = 2, SVLA_ACT_RELU = 3,
  SVLA_ACT_SOFTCAP = 4, /* act_param * tanh(x / act_param) */
  SVLA_ACT_SOFTPLUS = 5
};
enum {
  SVLA_GEMM_GEGLU = 1,       /* out[m, n/2] = gelu_tanh(v[m,2j]) * v[m,2j+1] (gate/up rows interleaved in W) */
  SVLA_GEMM_ACCUM_F32 = 2,   /* out_f32[m,n] += value (fp32 residual stream read-modify-write)            */
  SVLA_GEMM_CONV3X3 = 4      /* A is an NHWC tensor, implicit 3x3/pad 1/stride 1 convolution               */
};

typedef struct SvlaGemmArgs {
  const void* a;          /* bf16 [M, lda] (or NHWC [nb, h, w, c] in conv mode) */
  const void* w;          /* bf16 [N, ldw]; conv mode: [N][9][cpad] with cpad = roundup(c, 64) */
  const float* bias;      /* fp32 [N] or NULL */
  const float* colscale;  /* fp32 [N] or NULL */
  const void* res_bf16;   /* bf16 [M, ldo] residual or NULL */
  const void* res2_bf16;  /* second bf16 [M, ldo] residual or NULL */
  const float* res_f32;   /* fp32 [M or res_mod, ldo] residual or NULL */
  int64_t res_mod;        /* > 0: res_f32 row = m % res_mod (learned position embeddings broadcast over the batch) */
  void* out_bf16;         /* bf16 [M, ldo] or NULL */
  float* out_f32;         /* fp32 [M, ldo] or NULL */
  void* out_relu_bf16;    /* bf16 [M, ldo]: relu(value) or NULL */
  int64_t m, n, k;
  int64_t lda, ldw, ldo;  /* in elements */
  int32_t nb, h, wd, c;   /* conv mode geometry */
  float alpha;
  float act_param;
  int32_t act;
  int32_t flags;
  int32_t block_n;        /* 0 = auto; else 32 | 64 | 128 | 256 */
  int32_t impl;           /* 0 = tcgen05, kernel variant chosen by shape (product path); 1 = SIMT debugging kernel (tests only);
                             2 / 3 = force the 1-CTA / 2-CTA (cta_group::2) kernel; 4 = force the row-tile 3x3 conv kernel (n <= 128) */
  /* K extension (NULL = none): acc = A W^T + A2 W2^T in ONE accumulator -- the un-merged LoRA Linear y = x W^T + (s x A^T) B^T
   * (train/spatialvla_finetune.py:262-302, peft) and its input gradient dx = dy W + (s dy B) A cost k2 / 64 extra k-blocks of the
   * base GEMM.  a2 bf16 [M, lda2], w2 bf16 [N, ldw2], k2 columns each (zero columns are free: TMA zero-fills every tail). */
  const void* a2;
  const void* w2;
  int64_t k2, lda2, ldw2;
} SvlaGemmArgs;

int svla_gemm(const SvlaGemmArgs* args, void* stream);

/* G1s: skinny GEMM for the decode steps (M <= 128 activation rows): weight-streaming, swap-AB (UMMA M walks the weight
 * rows), split-K across CTAs so every SM streams weights. out[m,n] = act(alpha * sum_k x[m,k] w[n,k] + bias[n]).
 * flags: 1 = GEGLU (as svla_gemm), 2 = PARTIAL: raw fp32 partial sums out_f32[split * partial_stride + m*ldo + n] for
 * `splits` K-slices (the consumers svla_rmsnorm_residual / svla_rope_kv add them); 4 = W_TILED: w holds the same matrix
 * tile-major, bf16 [ceil(N/128)][ceil(K/64)][128][64] zero-padded (ldw ignored), so that every 16 KB pipeline stage is one
 * contiguous HBM read;
 * 8 = X_HILO (m <= 64): x holds TWO bf16 planes [2][M, ldx] -- hi = bf16(v) and lo = bf16(v - hi), 16 mantissa bits together --
 * and the result is hi.W^T + lo.W^T accumulated in fp32: the weights are streamed once, the activation tile is twice as wide
 * (the decode rows' own bf16 rounding before each of the 4 x 26 Linear layers is what dominates the bf16-vs-fp32 logit noise of
 * this path, tools/parity_report.py); 16 = OUT_HILO: out_bf16 is written as two such planes [2][M, ldo] (GEGLU or plain);
 * 32 = PAIR: CTA pairs (tcgen05 cta_group::2, 256 weight rows per MMA, half an activation tile staged per CTA) for the 64 / 128
 * column tiles -- same results, measured slower on the decode chain, opt-in.
 * model/modeling_gemma2.py:80-92,351-354,993 */
typedef struct SvlaSkinnyArgs {
  const void* x;          /* bf16 [M, ldx] */
  const void* w;          /* bf16 [N, ldw] */
  const float* bias;      /* fp32 [N] or NULL */
  void* out_bf16;         /* bf16 [M, ldo] or NULL */
  float* out_f32;         /* fp32 [M, ldo] (or [splits][M][ldo] with PARTIAL) or NULL */
  int64_t m, n, k, ldx, ldw, ldo, partial_stride;
  float alpha, act_param;
  int32_t act, flags, splits;
} SvlaSkinnyArgs;
int svla_gemm_skinny(const SvlaSkinnyArgs* args, void* stream);
int svla_gemm_skinny_splits(int64_t n, int64_t k);   /* the split count the library would pick for an [n, k] weight */

/* ---------------------------------------------------------------------------------------------------
 * G2: softmax(scale * Q K^T [softcap] [+ relpos bias] [mask]) V, flash style, fp32 softmax.
 *   model/modeling_gemma2.py:169-195 (GQA, tanh soft-capping; bidirectional prefill per
 *   model/modeling_spatialvla.py:291-297) ; HF siglip/modeling_siglip.py:252-312 ;
 *   HF beit/modeling_beit.py:225-306,511-590 (relative position bias looked up from the table on the fly) ;
 *   HF zoedepth/modeling_zoedepth.py:783-845.
 * q: element (b, s, h, d) at q + b*q_bs + s*q_ss + h*D + d   (strides in elements), same for k/v/out.
 * --------------------------------------------------------------------------------------------------- */
typedef struct SvlaAttnArgs {
  const void* q; const void* k; const void* v; void* out;       /* bf16 */
  int64_t q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs, o_ss;
  int32_t batch, hq, hkv, sq, sk, d;
  float scale;            /* applied before softcap */
  float softcap;          /* 0 = none */
  int32_t causal;         /* 1: key j visible to query i iff j <= i + (sk - sq) */
  const float* relpos_table; /* fp32 [(2*win-1)^2+3, hq] or NULL (BEiT) */
  int32_t relpos_win;
  int32_t relpos_head_major; /* 1: relpos_table is stored transposed, [hq, (2*win-1)^2+3] (one contiguous row per head: the
                                layout the engine packs once at load time so that a CTA reads its head's table coalesced) */
  const int32_t* kv_start;   /* device int32 [batch] or NULL: keys j < kv_start[b] are masked for every query of batch row b
                                (left-padded prompts: attention_mask == 0 columns, model/modeling_spatialvla.py:298-303) */
  int32_t causal_prefix;     /* with causal = 1: keys j < causal_prefix are visible to EVERY query (prefix-LM mask of the
                                training forward: token_type_ids == 0 columns unmasked on top of the triangular mask,
                                model/modeling_spatialvla.py:292-305); 0 = plain causal */
  float* lse;                /* NULL, or fp32 [batch, hq, lse_stride] (lse_stride >= sq): the kernel also writes log2(sum_j 2^(s_ij)) of every
                                query row, s = log2(e) * (soft-capped, scaled, masked score) -- what svla_attention_bwd needs from the
                                forward pass.  Only the tcgen05 kernels provide it (every shape of the training step). */
  int64_t lse_stride;
  int32_t window;            /* > 0: key slot j is masked for query slot i = query index + (sk - sq) when i - j >= window, on top of every
                                other mask (Gemma2's sliding-window layers, model/modeling_gemma2.py:461-471: tril(diagonal=-window)
                                also over the bidirectional prefix mask); 0 = global layer */
} SvlaAttnArgs;

int svla_attention(const SvlaAttnArgs* args, void* stream);

/* G3: single-query decode attention over the KV cache (model/modeling_gemma2.py:387-395,169-195).
 * q bf16 [B, hq*D]; kcache/vcache bf16 [B, smax, hkv, D]; attends to positions [0, ctx). out bf16 [B, hq*D] */
int svla_decode_attention(const void* q, const void* kcache, const void* vcache, void* out, int batch, int hq,
                          int hkv, int d, int smax, int ctx, float scale, float softcap, const int32_t* kv_start,
                          void* stream);

/* G3f: decode step after the qkv projection in ONE launch: RoPE of the new token's q/k (1-indexed position ctx), append of
 * its k/v at cache slot ctx-1, soft-capped softmax attention over slots [0, ctx) (model/modeling_gemma2.py:95-154,169-195,
 * 376-395). qkv_f32 = n_partials split-K partial sums [n_partials][B][(hq+2hkv)*D] of svla_gemm_skinny (partial_stride in
 * elements); kcache/vcache bf16 [B, smax, hkv, D]; out bf16 [B, hq*D]. D = 256, hq/hkv in {1, 2}. */
int svla_decode_attention_fused(const float* qkv_f32, int n_partials, int64_t partial_stride, void* kcache, void* vcache,
                                void* out, int batch, int hq, int hkv, int d, int smax, int ctx, float theta, float scale,
                                float softcap, const int32_t* kv_start, void* stream);
/* Extended entry.  out_lo != NULL: the output leaves as two bf16 planes, out_hi = bf16(o) and out_lo = bf16(o - out_hi), for the
 * X_HILO mode of svla_gemm_skinny, and the rotated query stays in fp32.  window > 0: sliding-window layer (even Gemma2 layers,
 * model/modeling_gemma2.py:343,441-473): only the last `window` cache slots [ctx - window, ctx) receive weight. */
int svla_decode_attention_fused_ex(const float* qkv_f32, int n_partials, int64_t partial_stride, void* kcache, void* vcache,
                                   void* out_hi, void* out_lo, int batch, int hq, int hkv, int d, int smax, int ctx, float theta,
                                   float scale, float softcap, const int32_t* kv_start, int window, void* stream);

/* G4: one whole Gemma2 decode step (all layers) for batch <= 64 in ONE persistent tensor-core launch: one CTA per SM
 * (cooperative launch), a TMA weight ring that streams the [128 x 64] weight tiles of ALL phases of ALL layers without waiting
 * for phase boundaries, swap-AB tcgen05.mma with split-K, and worker warps for RoPE + KV append + attention and the sandwich
 * norms; phases are separated by in-kernel grid barriers (model/modeling_gemma2.py:80-92,169-195,351-413,451-506,680-793).
 * It replaces the 7-launches-per-layer chain (svla_gemm_skinny / svla_decode_attention_fused / svla_rmsnorm_residual) and
 * produces bit-identical hidden states.
 *   svla_decode_mega_plan: HOST-side, once per model: writes the TMA descriptors of the 4 weight matrices of every layer
 *     (weights[4*l + {0,1,2,3}] = device pointers of wqkv [(hq+2hkv)d, hidden], wo [hidden, hq*d], wgu [2ff, hidden] with
 *     gate/up rows interleaved, wd [hidden, ff]; bf16 row-major) and of the activation buffers inside scratch_dev into HOST
 *     memory maps_host (svla_decode_mega_maps_bytes); the caller copies them to 64-byte aligned device memory.
 *   svla_decode_mega_step: x fp32 [batch, hidden] residual stream (embedded token times sqrt(hidden); updated in place),
 *     norm_w_dev: DEVICE array of 4*n_layers float pointers (ln_in, ln_post_attn, ln_pre_ff, ln_post_ff per layer), h_out bf16
 *     [batch, hidden] final-normed hidden state, kcache/vcache bf16 [layer][batch, smax, hkv, d] with cache_layer_stride
 *     elements between layers; ctx = cache slots after this step (the new token sits at slot ctx-1, RoPE position
 *     ctx - kv_start[b]).  scratch_dev: svla_decode_mega_scratch_bytes bytes, zero-initialised once, the same buffer the plan saw. */
int svla_decode_mega_supported(int batch, int hidden, int hq, int hkv, int d, int ff, int ctx);
int64_t svla_decode_mega_scratch_bytes(int hidden, int hq, int hkv, int d, int ff);
int64_t svla_decode_mega_maps_bytes(int n_layers);
int svla_decode_mega_plan(void* maps_host, const void* const* weights, int n_layers, int hidden, int hq, int hkv, int d, int ff,
                          void* scratch_dev);
int svla_decode_mega_step(const void* maps_dev, const void* norm_w_dev, int n_layers, float* x, const float* final_norm_w,
                          void* h_out_bf16, void* kcache, void* vcache, int64_t cache_layer_stride, void* scratch_dev, int batch,
                          int hidden, int hq, int hkv, int d, int ff, int smax, int ctx, float theta, float scale, float softcap,
                          float eps, const int32_t* kv_start, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Memory-bound fused kernels
 * --------------------------------------------------------------------------------------------------- */
/* M1 LayerNorm over fp32 rows (HF siglip :330-362, beit :448-508, zoedepth :843-877, spatialvla :59-64).
 * out_bf16 / out_f32 may each be NULL; relu applies to both outputs. */
int svla_layernorm(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int cols,
                   void* out_bf16, float* out_f32, int relu, void* stream);

/* M2 Gemma2 sandwich norm (model/modeling_gemma2.py:60-77,475-496):
 *   if branch:  x += rms(branch) * (1 + w_post)        (x fp32 residual, updated in place)
 *   if w_pre:   out_bf16 = rms(x) * (1 + w_pre)                                                     */
int svla_rmsnorm_residual(float* x, const float* branch, const float* w_post, const float* w_pre, float eps,
                          int64_t rows, int cols, void* out_bf16, int n_partials, int64_t partial_stride, void* stream);
/* n_partials > 1: branch = sum_s branch[s * partial_stride + ...] (split-K partial sums of svla_gemm_skinny) */
/* hi/lo variant for the decode chain (rows < 2048): out_hi = bf16(v), out_lo = bf16(v - out_hi) */
int svla_rmsnorm_residual_hilo(float* x, const float* branch, const float* w_post, const float* w_pre, float eps,
                               int64_t rows, int cols, void* out_hi_bf16, void* out_lo_bf16, int n_partials,
                               int64_t partial_stride, void* stream);

/* M3 RoPE + KV-cache write (model/modeling_gemma2.py:95-154,376-395; positions 1-indexed per
 * model/modeling_spatialvla.py:371-372). qkv bf16 [B*S, (hq+2hkv)*D] -> q_out bf16 [B*S, hq*D] (rotated),
 * kcache/vcache bf16 [B, smax, hkv, D] rows [pos0, pos0+S).
 * row_pads: device int32 [B] or NULL -- leading padding tokens per row of a left-padded batch: cache slot i of row b gets
 * position i - row_pads[b] + 1 (padding slots: 2), the position ids HF generate derives from the attention mask
 * (model/modeling_gemma2.py:1042-1051 + model/modeling_spatialvla.py:473-474). */
int svla_rope_kv(const void* qkv, void* q_out, void* kcache, void* vcache, int batch, int s, int hq, int hkv,
                 int d, int smax, int pos0, float theta, const float* qkv_f32, int n_partials, int64_t partial_stride,
                 const int32_t* row_pads, void* stream);
/* qkv_f32 != NULL: read qkv as the sum of n_partials fp32 partial buffers instead of the bf16 `qkv` */

/* M6 embedding gather (model/modeling_spatialvla.py:361-387, model/modeling_gemma2.py:741-742):
 * text ids -> embed (bf16 [V,H]); ids in [act_lo, act_lo+n_act) -> spatial_embed (bf16 [n_act,H]);
 * the k-th image token of row b -> image_feats fp32 [B, n_img, H] row k; everything * sqrt(H) -> x fp32 */
int svla_embed_tokens(const int64_t* ids, const void* embed, const void* spatial_embed, const float* image_feats,
                      float* x, int batch, int s, int hdim, int64_t vocab, int64_t image_token, int64_t act_lo,
                      int64_t n_act, int n_img, float normalizer, int* status_flag, void* stream);

/* M7 argmax over fp32 logits rows (greedy step of HF generate, restricted to the action slice):
 * out_ids[b*out_stride] = argmax_j logits[b, j] + id_offset (first max wins, like torch.argmax) */
int svla_argmax_rows(const float* logits, int64_t rows, int64_t cols, int64_t ld, int64_t id_offset,
                     int64_t* out_ids, int64_t out_stride, void* stream);

/* M8: cross entropy over full-vocabulary logit rows -- the loss of the training / evaluation forward
 * (model/modeling_spatialvla.py:413-430: shift, drop ignore_index rows, nn.CrossEntropyLoss mean).  The logit rows of a big
 * batch are produced chunk by chunk: logits fp32 [rows, ld] (post-softcap) holds entries [row_offset, row_offset + rows) of the
 * FULL arrays labels int64 / row_loss fp32 / row_argmax int64.  row_loss[i] = logsumexp(logits row) - logit[labels[i]] (0 where
 * labels[i] == ignore_index, NaN where a label is out of range), row_argmax[i] = first maximal column.  summary (NULL except
 * on the last chunk) fp32 [3] = { mean of row_loss over the non-ignored entries of [0, row_offset + rows), their count, entries
 * whose argmax equals the label } (token accuracy of train/monkey_patch.py:267-324). One launch per chunk + one for the summary. */
int svla_cross_entropy_rows(const float* logits, int64_t rows, int64_t cols, int64_t ld, const int64_t* labels,
                            int64_t ignore_index, float* row_loss, int64_t* row_argmax, int64_t row_offset, float* summary,
                            void* stream);

/* M8b: backward of M8 -- gradient of the mean cross entropy (model/modeling_spatialvla.py:413-430) w.r.t. the PRE-soft-cap logits
 * z (model/modeling_gemma2.py:993-997: logit = cap * tanh(z / cap)), emitted as the bf16 A operand of the dh = dz * W_head GEMM:
 * dz[i, j] = (softmax(logits[i])[j] - [j == labels[i]]) * (1 - (logits[i, j] / cap)^2) / count, zero rows for ignored labels and
 * zero columns in [cols, ldo) (K padding; ldo even).  logits fp32 [rows, ld] = entries [row_offset, row_offset + rows) of labels /
 * row_loss as in svla_cross_entropy_rows; row_loss is that call's output (the row's logsumexp is row_loss + logit[label]);
 * summary = its fp32 [3] summary of the WHOLE batch (count read on the device); softcap 0 = no soft-capping. rows <= 65535. */
int svla_cross_entropy_bwd(const float* logits, int64_t rows, int64_t cols, int64_t ld, const int64_t* labels,
                           int64_t ignore_index, const float* row_loss, int64_t row_offset, const float* summary,
                           float softcap, void* dz_bf16, int64_t ldo, void* stream);

/* AdamW step on flat fp32 buffers (the LoRA arena of the fine-tune config: HF Trainer default adamw_torch, scripts/spatialvla_4b_finetune/
 * finetune_lora.sh; arithmetic order of torch.optim.AdamW): grad is multiplied by grad_scale first (1 / world size or a clipping
 * coefficient), p *= 1 - lr * wd, m = lerp(m, g, 1 - beta1), v = beta2 v + (1 - beta2) g^2,
 * p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps).  step counts from 1.  Buffers 16-byte aligned.
 * Hyper-parameters are doubles: the derived scalars (1 - beta2, lr / (1 - beta1^step), ...) are formed in double on the host, as
 * torch forms them from Python floats, and only then rounded to fp32.
 */
int svla_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                    double beta2, double eps, double weight_decay, int64_t step, double grad_scale, const float* grad_sumsq_dev,
                    double max_grad_norm, void* stream);
/* grad_sumsq_dev (NULL = no clipping): DEVICE scalar holding sum(grad^2) of the un-scaled buffer (svla_sumsq): the gradient is also
 * multiplied by min(1, max_grad_norm / (grad_scale * sqrt(sumsq) + 1e-6)) -- torch.nn.utils.clip_grad_norm_ as HF Trainer applies it
 * (max_grad_norm 1.0) -- without a host read of the norm.   *out += sum(x^2). */
int svla_sumsq(const float* x, int64_t n, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fine-tune step (BASELINE.json config #5): training-mode forward pieces and the backward kernels.
 * The reference trains through torch autograd (train/spatialvla_finetune.py:262-302 PEFT LoRA r = 32 on every Linear of Gemma2,
 * SigLIP, projector and Ego3D head; HF Trainer / train/monkey_patch.py:222-326); base weights and norm weights are frozen, so
 * only activation gradients and the rank-r adapter gradients are formed.  Closed forms: oracle/backward_ref.py.
 * --------------------------------------------------------------------------------------------------- */
/* Gemma2 sandwich norm, out of place (model/modeling_gemma2.py:60-77,475-496): x_out = x_in + rms(branch) (1 + w_post) [branch,
 * w_post, x_out all NULL: skipped]; h_bf16 = rms(x) (1 + w_pre) of the updated row [w_pre, h_bf16 NULL: skipped]. */
int svla_rmsnorm_train_fwd(const float* x_in, const float* branch, const float* w_post, const float* w_pre, float eps, int64_t rows,
                           int cols, float* x_out, void* h_bf16, void* stream);
/* RMSNorm backward: dx = r (g - x r^2 mean(g x)), g = dy (1 + w).  dy bf16 [rows, cols] or fp32 (dy_is_f32); row_idx (NULL = identity):
 * dy row i belongs to row row_idx[i] of x / dx.  dx_accum[row] += dx (fp32, may be NULL), dx_bf16[row] = dx (may be NULL). */
int svla_rmsnorm_bwd(const float* x, const float* w, const void* dy, int dy_is_f32, const int64_t* row_idx, float eps, int64_t rows,
                     int cols, float* dx_accum, void* dx_bf16, void* stream);
/* LayerNorm backward (HF siglip :330-362; model/modeling_spatialvla.py:59-64 with relu = 1: the ReLU mask is recomputed from x):
 * dx_accum += dx (NULL = skip), copy_bf16 = bf16(updated dx_accum) (NULL = skip), dx_bf16 = dx (NULL = skip). dy bf16. */
int svla_layernorm_bwd(const float* x, const float* gamma, const float* beta, const void* dy_bf16, float eps, int64_t rows, int cols,
                       int relu, float* dx_accum, void* copy_bf16, void* dx_bf16, void* stream);
/* GeGLU (model/modeling_gemma2.py:80-92) on the un-fused gate/up GEMM output gu bf16 [rows, 2 inter] (col 2j = gate_j, 2j + 1 = up_j):
 * act[rows, inter] = gelu_tanh(gate) up;  backward dgu = (dact up gelu'(gate), dact gelu(gate)) interleaved. */
int svla_geglu_fwd(const void* gu, void* act, int64_t rows, int64_t inter, void* stream);
int svla_geglu_bwd(const void* gu, const void* dact, void* dgu, int64_t rows, int64_t inter, void* stream);
/* gelu_pytorch_tanh and its derivative on bf16 (SigLIP MLP, HF siglip :316-327): f = gelu(z); dz = df gelu'(z). n % 8 == 0 */
int svla_gelu_tanh_fwd(const void* z, void* f, int64_t n, void* stream);
int svla_gelu_tanh_bwd(const void* z, const void* df, void* dz, int64_t n, void* stream);
/* RoPE backward (model/modeling_gemma2.py:140-154): inverse rotation, in place, of the q and k column blocks of the gradient of the
 * rotated qkv tensor bf16 [batch * s, (hq + 2 hkv) d]; the v block is untouched; position of row t is t % s + 1. */
int svla_rope_bwd(void* dqkv, int batch, int s, int hq, int hkv, int d, float theta, void* stream);
/* out_bf16[i, :] = scale * src[row_idx ? row_idx[i] : i, :]  (fp32 rows -> bf16 GEMM operand; gathers the image-token rows of the
 * embedding gradient, model/modeling_spatialvla.py:361-387 backwards) */
int svla_rows_cast(const float* src, const int64_t* row_idx, float scale, int64_t rows, int cols, void* out_bf16, void* stream);
/* LoRA operand packing: one launch turns the fp32 master copy of every adapter (flat arena) into the bf16 operand layouts of the
 * GEMMs.  descs_dev: device array of n_desc records { int64 src_off, dst_off, stride_i, stride_j; int32 rows, cols, tile0, pad }:
 * pool[dst_off + i stride_i + j stride_j] = bf16(arena[src_off + i cols + j]); tile0 = running count of 32x32 tiles (ascending). */
int svla_lora_pack(const float* arena, void* pool_bf16, const void* descs_dev, int n_desc, int total_tiles, void* stream);
int svla_fill_zero(void* ptr, int64_t bytes, void* stream);   /* cudaMemsetAsync */

/* Backward of G2 (flash formulation; see csrc/train_mma.cu).  q/k/v/out as in SvlaAttnArgs (out = the forward result), dout = dL/dout;
 * dq/dk/dv bf16 with their own strides (they may be column blocks of one [tokens, (hq + 2 hkv) d] tensor); lse / delta: fp32
 * [batch, hq, sq] scratch written by the first launch.  Masks: causal / causal_prefix / window as in SvlaAttnArgs (no kv_start, no relpos). */
typedef struct SvlaAttnBwdArgs {
  const void* q; const void* k; const void* v; const void* out; const void* dout;
  void* dq; void* dk; void* dv;
  int64_t q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs, o_ss, do_bs, do_ss, dq_bs, dq_ss, dk_bs, dk_ss, dv_bs, dv_ss;
  float* lse; float* delta;
  int32_t batch, hq, hkv, sq, sk, d;
  float scale, softcap;
  int32_t causal, causal_prefix;
  const float* fwd_lse2;     /* NULL, or the forward kernel's `lse` output (log2 domain), fp32 [batch, hq, lse_stride]: the backward then runs on
                                the tcgen05 kernels (csrc/attention_bwd_tc.cu: dQ, dK and dV sweeps with TMEM accumulators) and `lse` / `delta`
                                must be fp32 [batch, hq, lse_stride] scratch (delta is written, lse unused); NULL = the warp-MMA kernels, which
                                recompute the row statistics themselves */
  int64_t lse_stride;
  int32_t window;            /* sliding-window predicate of the forward (SvlaAttnArgs.window); 0 = none */
} SvlaAttnBwdArgs;
int svla_attention_bwd(const SvlaAttnBwdArgs* args, void* stream);

/* Rank-r gradient reduction over the token dimension: for every group g, dst_g[r - row0, j] += scale * sum_m S[m, r] Y[m, col_start +
 * j col_stride] for r in [row0, row0 + rows), j in [0, ncols)  (fp32 reductions, dst must be zero-initialised or hold a running sum).
 * gA = (s dY B)^T X and gB^T = (s X A^T)^T dY of a LoRA Linear (oracle/backward_ref.lora_linear_bwd); groups address the adapters of a
 * fused q|k|v or interleaved gate/up projection inside the gradient arena.  S bf16 [m, lds] (r <= 128 columns used), Y bf16 [m, ldy]. */
typedef struct SvlaTnGroup { float* dst; int64_t ld; int32_t row0, rows, col_start, col_stride, ncols, pad; } SvlaTnGroup;
typedef struct SvlaGemmTnArgs {
  const void* s; const void* y;
  int64_t m, lds, ldy;
  int32_t r, n;
  float scale;
  int32_t n_groups;
  SvlaTnGroup groups[4];
} SvlaGemmTnArgs;
int svla_gemm_tn(const SvlaGemmTnArgs* args, void* stream);

/* M9 image preprocessing.
 * siglip: (x-0.5)/0.5 + im2col for the 14x14/14 patch conv (model/modeling_spatialvla.py:309;
 *         HF siglip :124-130,176-187): px fp32 [B,3,224,224] -> a bf16 [B*256, kpad] (kpad >= 588, zero padded)
 * zoe   : reflect-pad 31 -> bicubic(align_corners) 286->384 -> (x-0.5)/0.5 -> im2col 16x16/16
 *         (model/modeling_spatialvla.py:99-110; HF beit :209): -> a bf16 [B*576, 768] */
int svla_siglip_patchify(const float* px, void* a, int batch, int kpad, void* stream);
int svla_zoe_patchify(const float* px, void* a, int batch, void* stream);

/* M9b observation frames on the device (model/processing_spatialvla.py:174 -> HF SiglipImageProcessor @ transformers 4.47 -> Pillow
 * ImagingResample): images uint8 [B, h, w, 3] -> out fp32 [B, 3, oh, ow] = value_lut[c][resize_bicubic_u8(image)].  Pillow's
 * separable fixed-point resampler, horizontal pass first (into tmp_u8 [B, h, ow, 3]; may be NULL when w == ow), then vertical;
 * bounds_* int32 [o, 2] = (first source index, taps), kk_* int32 [o, ksize_*] 22-bit fixed-point coefficients (host-built per
 * input size); value_lut fp32 [3, 256] = what the image processor turns each uint8 level into (rescale, optional normalise). */
int svla_image_preprocess(const void* images_u8, int batch, int h, int w, void* tmp_u8, float* out, int oh, int ow,
                          const int32_t* bounds_h, const int32_t* kk_h, int ksize_h, const int32_t* bounds_v, const int32_t* kk_v,
                          int ksize_v, const float* value_lut, void* stream);

/* Fine-tune-time re-gridding of the spatial embeddings (model/action_tokenizer.py:390-430: scipy griddata 'linear' = Delaunay
 * interpolation of the old padded grid at the new bin centres): out[t, :] = sum_{v<4} weights[t, v] * src[rows[t, v], :] with double
 * accumulation in scipy's order; rows[t, 0] < 0 (target outside the old hull) gives a NaN row like the reference.  The triangulation /
 * point location is E-independent host work (SpatialActionTokenizer.adaption_plan, same Qhull call as griddata); this is the
 * 8 192 x 2 304 gather.  src fp32 [n_src, e], rows int32 [n_out, 4], weights fp64 [n_out, 4], out fp32 [n_out, e]. */
int svla_barycentric_gather(const float* src, int64_t n_src, const int32_t* rows, const double* weights, float* out, int64_t n_out,
                            int e, void* stream);

/* BEiT token assembly: x[b,0,:] = cls, x[b,1+i,:] = patches[b*n+i,:] (HF beit :190-222). fp32 */
int svla_beit_assemble(const float* patches, const float* cls, float* x, int batch, int n, int c, void* stream);

/* ZoeDepth reassemble read-out 'project' input (HF zoedepth :55-110): a bf16 [B*n, 2c] = [tok_i | cls] */
int svla_readout_concat(const float* hs, void* a, int batch, int n, int c, void* stream);

/* ConvTranspose2d with kernel == stride == f as GEMM + this scatter (HF zoedepth :129-149):
 * g bf16 [B*h*w, f*f*c] (col = (i*f+j)*c + co) -> out bf16 NHWC [B, h*f, w*f, c] */
int svla_pixel_shuffle(const void* g, void* out, int batch, int h, int w, int c, int f, void* stream);

/* im2col for the single 3x3 stride-2 pad-1 conv of the neck (HF zoedepth :129-149):
 * x bf16 NHWC [B,h,w,c] -> a bf16 [B*(h/2)*(w/2), 9*c] (col = tap*c + ci) */
int svla_im2col3x3_s2(const void* x, void* a, int batch, int h, int w, int c, void* stream);

/* Bilinear resize of NHWC bf16 maps, align_corners=True (HF zoedepth :264-272,350,716-728,1094-1095):
 * out = resize(x) [+ add (same shape as out)], optional relu copy. */
int svla_bilinear_nhwc(const void* x, const void* add, void* out, void* out_relu, int batch, int h, int w, int c,
                       int oh, int ow, void* stream);

/* elementwise relu(x) -> out (bf16), n elements (pre-activation residual units, HF zoedepth :182-238) */
int svla_relu_bf16(const void* x, void* out, int64_t n, void* stream);

/* ZoeDepth patch-transformer input (HF zoedepth :905-963): e fp32 [B, 1+n, c]: row 0 = PE[0] (zero CLS),
 * row 1+i = conv[b*n+i] + PE[1+i] with the 1-D sinusoidal encoding. conv fp32 [B*n, c]; e_bf16 = bf16 copy */
int svla_zoe_router_embed(const float* conv, float* e, void* e_bf16, int batch, int n, int c, void* stream);

/* ZoeDepth metric-bins tail (HF zoedepth :551-570,665-746): softplus attractors, bilinear(align_corners)
 * upsampling of the previous bin centres, inverse-attractor update with alpha=300, gamma=2, mean over na.
 * attr bf16 [B*oh*ow, na] (pre-softplus), prev fp32 NHWC [B, h, w, nbins] -> out fp32 NHWC [B, oh, ow, nbins];
 * na <= 16, nbins % 4 == 0 (the reference has na in {16, 8, 4, 1}, nbins = 64) */
int svla_zoe_attractor(const void* attr, const float* prev, float* out, int batch, int h, int w, int oh, int ow,
                       int na, int nbins, void* stream);

/* ZoeDepth metric-head router decision on the device (HF zoedepth :1059-1067 picks argmax_h sum_b domain_logits[b, h] and reads
 * it back with `.item()`): the weights of every metric-bins head are packed by the caller into one byte arena per head (identical
 * layout); this call copies the arena of the selected head into `active_arena`, which the kernels of the bins stage read, and
 * writes the index to head_out (may be NULL).  No host synchronisation: the whole predict_action step is ONE CUDA graph.
 * head_arenas_dev: DEVICE array of n_heads device pointers; forced_head >= 0 overrides the vote (tests, per-shard parity). */
int svla_zoe_select_head(const float* domain_logits, int batch, int n_heads, int forced_head, const void* const* head_arenas_dev,
                         void* active_arena, int64_t bytes, int* head_out, void* stream);

/* softplus(x) bf16 -> fp32 (seed bin regressor, HF zoedepth :494-547) */
int svla_softplus_f32(const void* x, float* out, int64_t n, void* stream);

/* ZoeDepth conditional log-binomial + expectation (HF zoedepth :383-491,1094-1101) at full resolution:
 * t bf16 [B*oh*ow, nh] = W_a*last (no bias); e bf16 NHWC [B, h, w, nh] = W_b*bin_embedding (half res);
 * b1 fp32 [nh]; w2 fp32 [4, nh]; b2 fp32 [4]; bins fp32 NHWC [B, h, w, nbins]
 * -> depth fp32 [B, oh, ow] = sum_k softmax_k(logbinom(p)/T) * bilinear(bins)_k */
int svla_zoe_depth_tail(const void* t, const void* e, const float* b1, const float* w2, const float* b2,
                        const float* bins, float* depth, int batch, int h, int w, int oh, int ow, int nh,
                        int nbins, float min_temp, float max_temp, void* stream);
/* Same tail with the first MLP layer folded in: t = W_a * x is computed per pixel inside the kernel from the relative-head
 * features x bf16 [B*oh*ow, nx] and wa bf16 [nh, nx] (no bias), so the [pixels, nh] tensor never reaches HBM; the
 * half-resolution operands are staged per 16x16 output tile in shared memory.  Specialised for nx = 32, nh = 40, nbins = 64
 * (the ZoeDepth-NK head) and up-sampling ratios (oh >= ~1.4 h). */
int svla_zoe_depth_tail_fused(const void* x, const void* wa, const void* e, const float* b1, const float* w2,
                              const float* b2, const float* bins, float* depth, int batch, int h, int w, int oh, int ow,
                              int nx, int nh, int nbins, float min_temp, float max_temp, void* stream);

/* M5 Ego3D (model/modeling_spatialvla.py:41-97,181-223,318-323): depth384 fp32 [B,384,384] ->
 * bicubic(align_corners) 384->286, crop 31 -> 7x7 area mean -> inv(K) uv d -> xyz fp32 [B,256,12]
 * -> [(x-c)/2, sin(2^k .), cos(2^k .)] -> enc bf16 [B*256, kpad] (kpad >= 204, zero padded).
 * intrinsic fp32 [3,3] (k_stride = 0) or [B,3,3] (k_stride = 9). */
int svla_ego3d_encode(const float* depth384, const float* intrinsic, int k_stride, float* xyz, void* enc, int batch,
                      int kpad, int n_freqs, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * M8 SpatialActionTokenizer grid lookup and inverse (model/action_tokenizer.py:105-137,177-202,227-243,
 * 305-333), float64 arithmetic, bit-exact ids.
 * edges: 6 fp64 arrays concatenated [theta(nt+1) | phi(np+1) | r(nr+1) | roll | pitch | yaw];
 * nbins int32[7] = {theta, phi, r, roll, pitch, yaw, gripper}.
 * encode: actions fp64 [n,7] -> ids int32 [n,3] LOCAL ids (0..vocab-1); decode: ids int64 [n,3]
 * (global ids, `begin` subtracted inside like the reference) -> actions fp64 [n,7].
 * use_spherical = 0 bins the Cartesian translation directly (model/action_tokenizer.py:110-113,135).
 * nbins is always read on the HOST (7 ints of configuration).
 * The `_host` variants take HOST buffers and do the H2D/D2H copies themselves (the e2e call). */
/* edge_trig (NULL = library atan2 path): fp64 [(theta bins - 1) + (phi bins - 1)][4] = (cos_hi, cos_lo, sin_hi, sin_lo) of the
 * rounding boundary m = (pred(e) + e) / 2 below each INTERIOR theta / phi edge e, double-double (built on the host with 200-bit
 * arithmetic, spatialvla_b200/action_tokenizer.py::edge_trig_table); phi_nonpos / phi_neg = number of interior phi edges <= 0 / < 0.
 * With the table the angular bins are decided WITHOUT atan2: `fl(atan2(a, b)) >= e` of np.digitize (model/action_tokenizer.py:115-118)
 * is the sign of a cos m - b sin m, evaluated in double and, within 8 roundings of zero, with error-free transformations (2^-104):
 * the ids a correctly rounded atan2 yields, independent of the math library's last-ulp behaviour. */
int svla_tok_encode(const double* actions, const double* edges, const int32_t* nbins, int32_t* ids, int64_t n,
                    double min_action, double max_action, int use_spherical, const double* edge_trig, int phi_nonpos,
                    int phi_neg, void* stream);
/* center_trig (NULL = library sincos, <= 4 ulp from glibc): fp64 [(theta bins) + (phi bins)][2] = (sin, cos) of every bin CENTRE
 * 0.5 (e[i] + e[i+1]) -- the only angles the inverse evaluates (model/action_tokenizer.py:99-103,129-135) -- tabulated once on the
 * host with the reference's own libm (numpy), so the decoded x, y, z are the reference's doubles bit for bit. */
int svla_tok_decode(const int64_t* ids, const double* edges, const int32_t* nbins, int64_t begin, double* actions,
                    int64_t n, int use_spherical, const double* center_trig, void* stream);
int svla_tok_encode_host(const double* actions_host, const double* edges_host, const int32_t* nbins_host,
                         int32_t* ids_host, int64_t n, double min_action, double max_action, int use_spherical,
                         const double* edge_trig_host, int phi_nonpos, int phi_neg);
int svla_tok_decode_host(const int64_t* ids_host, const double* edges_host, const int32_t* nbins_host, int64_t begin,
                         double* actions_host, int64_t n, int use_spherical, const double* center_trig_host);

#ifdef __cplusplus
}
#endif
#endif /* SPATIALVLA_B200_H */
