/* b200fm C ABI -- the drop-in boundary of the B200-native 4M hot path.
 *
 * The reference (apple/ml-4m) is pure Python/PyTorch: it has no FFI of its own, so each entry point below names the
 * reference Python call site it replaces (paths relative to the reference root).  All pointers are DEVICE pointers
 * unless a name ends in `_host`; sizes are plain integers; `stream` is a cudaStream_t passed as void*.  No torch types
 * cross this boundary.  Every function returns 0 on success, non-zero on failure; b200fm_last_error() then holds a
 * message (thread-local).  Kernels are compiled for sm_100a only.
 *
 * Binding (Python side, what a maintainer adds to the reference): see INTEGRATION.md -- `ctypes.CDLL("libb200fm.so")`
 * with tensors passed as `t.data_ptr()` and `torch.cuda.current_stream().cuda_stream`.
 */
#ifndef B200FM_H_
#define B200FM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FM_ABI_VERSION 1

/* GEMM operand layouts (row-major storage everywhere) */
#define B200FM_LAYOUT_NT 0 /* A [M,K], B [N,K]  : y = x W^T        (nn.Linear forward)  */
#define B200FM_LAYOUT_NN 1 /* A [M,K], B [K,N]  : dx = dy W        (nn.Linear dgrad)    */
#define B200FM_LAYOUT_TN 2 /* A [K,M], B [K,N]  : dW = dy^T x      (nn.Linear wgrad)    */

/* GEMM epilogues */
#define B200FM_EPI_BF16 0   /* out0 bf16 [M,N] = alpha * (acc (+bias))                                             */
#define B200FM_EPI_F32 1    /* out0 fp32 [M,N] = alpha * (acc (+bias))                                            */
#define B200FM_EPI_RESID 2  /* out0 fp32 [M,N] = resid fp32 + bf16(acc (+bias))      (x = x + proj(...))          */
#define B200FM_EPI_SWIGLU 3 /* B = [fc1; fc3] [2N,K]; out0 bf16 [M,2N] = [a | b]; out1 bf16 [M,N] = silu(a) * b   */
#define B200FM_EPI_GELU 4   /* out0 bf16 [M,N] = pre-activation; out1 bf16 [M,N] = gelu(out0)                     */
#define B200FM_EPI_TANH 5   /* as EPI_GELU with tanh (ViT tokenizer post_mlp, vq/models/vit_models.py:494-496)     */

const char* b200fm_last_error(void);
int b200fm_abi_version(void);
int b200fm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);

/* Runtime options (defaults from the environment variable B200FM_<NAME upper-case>): "pdl" (1: programmatic dependent launch),
 * "gemm_cta_pairs" (1: cta_group::2 GEMM tiles), "ln_bwd_v2" (1: LayerNorm backward with all loads of a row issued before its reductions; 0: plain kernel), "sm_reserve" (0: number
 * of SMs the persistent GEMM grids leave free, for a concurrent gradient all-reduce kernel), "gemv" (1: NT problems with <= 8 rows --
 * the linears of the K/V-cached decode loop -- run on an HBM-bound weight-streaming kernel instead of a tcgen05 tile), "gemv_prefetch"
 * (1: that kernel pulls its weight rows into L2 before it waits for the preceding kernel of the stream), "gemm_tma_store" (1: bf16 GEMM outputs are written by TMA stores), "comm_slim" (1: the gradient all-reduce runs as many
 * 128-thread CTAs co-resident with the compute kernels instead of a few wide ones on reserved SMs), "gemm_debug" (measurement
 * only, results are WRONG when non-zero: 1 = the GEMM epilogue reads the accumulator but stores nothing, 2 = no epilogue).  Changing an
 * option affects launches issued afterwards; meant for A/B measurements inside one process.                          */
int b200fm_set_option(const char* name, int value);
int b200fm_get_option(const char* name, int* value);

/* ---- dense contractions (tcgen05) ---------------------------------------------------------------------------
 * Replaces every nn.Linear / F.linear on the path: fourm/models/fm_utils.py:155-157,163,178 (qkv, proj),
 * :190-194,200-201,217 (q, kv, proj), :136-144 (GatedMlp fc1/fc3/fc2), :116-125 (Mlp), fourm/models/fm.py:154,679
 * (decoder_proj_context), fourm/models/decoder_embeddings.py:141-152,257-268 (to_logits),
 * fourm/models/encoder_embeddings.py:301 (patch proj), and their autograd backward (dgrad = NN, wgrad = TN).
 * bf16 operands, fp32 accumulation (same contract as the reference under torch.autocast(bf16)).
 * lda/ldb/ld0/ld1/ldr are row strides in ELEMENTS.  lda, ldb multiples of 8.  alpha_dev: optional device scalar
 * multiplied into alpha (EPI_F32 / EPI_BF16), lets a loss scale stay on the device.                                 */
int b200fm_gemm_bf16(int layout, int epilogue, int M, int N, int K, const void* A, long long lda, const void* B,
                     long long ldb, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
                     const float* resid, long long ldr, float alpha, const float* alpha_dev, void* stream);

/* Same with a DEVICE-side problem size (the masked-token head under a captured CUDA graph: per-modality row counts never visit the
 * host).  dyn_mode 1 (NT / NN): the number of output rows is min(*dyn_dev, M); dyn_mode 2 (TN, EPI_F32): the contraction length is
 * min(*dyn_dev, K) -- operand rows in [*dyn_dev, roundup(*dyn_dev, 64)) must be zero in one operand and finite in the other.
 * Grid, tensor maps and the split-K plan are sized for the bounds M / K; dyn_dev == NULL behaves like b200fm_gemm_bf16.        */
int b200fm_gemm_bf16_dyn(int layout, int epilogue, int M, int N, int K, const void* A, long long lda, const void* B, long long ldb,
                         void* out0, long long ld0, void* out1, long long ld1, const float* bias, const float* resid, long long ldr,
                         float alpha, const float* alpha_dev, const int* dyn_dev, int dyn_mode, void* stream);

/* ---- LayerNorm (fourm/models/fm_utils.py:93-108; F.layer_norm, fp32 statistics) -----------------------------
 * x fp32 [rows, D] -> y (bf16 if y_is_bf16 else fp32) [rows, D]; mean/rstd fp32 [rows] saved for backward (may be NULL). */
int b200fm_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean,
                         float* rstd, int rows, int D, float eps, void* stream);
/* Residual add fused into the norm: x_out fp32 = x + add (bf16 branch output of the previous sub-layer), y = LayerNorm(x_out).
 * This is how `x = x + proj(...)` / `x = x + fc2(...)` (fm_utils.py:332-334, 363-365) are realised: the GEMM writes its bf16
 * output, the NEXT norm adds it to the fp32 stream while it reads the stream anyway.                                        */
int b200fm_add_layernorm_fwd(const float* x, const void* add_bf16, float* x_out, const float* gamma, const float* beta, void* y,
                             int y_is_bf16, float* mean, float* rstd, int rows, int D, float eps, void* stream);
/* dy (bf16 if dy_is_bf16 else fp32) is the gradient w.r.t. the LN output; dx_out fp32 = (dres ? dres : 0) + LN backward;
 * dx_bf16 (optional) = bf16 copy of dx_out; dgamma/dbeta fp32 [D] are ACCUMULATED into (atomics), may be NULL.   */
int b200fm_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                         const float* rstd, const float* dres, float* dx_out, void* dx_bf16, float* dgamma, float* dbeta,
                         int rows, int D, void* stream);

/* Per-head LayerNorm over head_dim 64 on q / k (qk_norm presets: NormAttention / NormCrossAttention, fm_utils.py:222-307).
 * x bf16 [rows, ldx], head h = columns [h*64, h*64+64) (may be a column slice of a packed qkv); y bf16 [rows, ldy];
 * stats fp32 [rows, H, 2] = (mean, rstd).  Backward: dx bf16, dgamma / dbeta fp32 [64] ACCUMULATED into (may be NULL).     */
int b200fm_headnorm_fwd(const void* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy, float* stats,
                        long long rows, int H, float eps, void* stream);
int b200fm_headnorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* gamma, const float* stats, void* dx,
                        long long lddx, float* dgamma, float* dbeta, long long rows, int H, void* stream);

/* ---- fused multi-head attention (fourm/models/fm_utils.py:160-180, 197-219; vq/models/vit_models.py:165-197) --
 * q rows [B*Nq, ldq], k/v rows [B*Nk, ldk/ldv] bf16, head h occupies columns [h*64, h*64+64) from each base pointer
 * (so q/k/v may alias one packed qkv buffer).  mask: uint8/bool, 1 = masked (reference convention), addressed as
 * mask[b*mask_b_stride + i*mask_q_stride + j]; NULL = no mask.  Masked scores are filled with -FLT_MAX-like
 * finite value (fully masked rows -> uniform attention, like masked_fill(-finfo.max)).
 * out bf16 [B*Nq, ldo]; lse ("stats") fp32 [B, H, Nq, 2] = (row max of s*scale*log2e, 1/rowsum), saved for backward.
 * head_dim is fixed at 64.  attention_fwd: Nk <= 256 keeps all keys of a (b, h, 128-query) item resident; Nk > 256 (the
 * generation callers, generate.py:407-445, 886-913) streams 128-key tiles with the running-max rescale.  attention_bwd:
 * Nq, Nk <= 256; needs dsum_ws: fp32 [B, H, Nq] scratch (receives rowsum(dO o O)).                                           */
int b200fm_attention_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                         const uint8_t* mask, long long mask_b_stride, long long mask_q_stride, void* out, long long ldo,
                         float* lse, int B, int H, int Nq, int Nk, float scale, void* stream);
/* One query row per sequence (the attention calls of the K/V-cached decode step, generate.py:886-901): q bf16 [B, >= H*64], k / v
 * bf16 rows [B*Nk, >= H*64] (e.g. the two column halves of a [B, Nk, 2D] cache), mask uint8 [B, Nk] (1 = masked, stride
 * mask_b_stride; may be null), out bf16 [B, >= H*64].  Exact fp32 softmax; a fully masked row is uniform (masked_fill semantics). */
int b200fm_attention_decode(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                            const uint8_t* mask, long long mask_b_stride, void* out, long long ldo, int B, int H, int Nk,
                            float scale, void* stream);
int b200fm_attention_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                         const uint8_t* mask, long long mask_b_stride, long long mask_q_stride, const void* out,
                         long long ldo, const void* dout, long long lddo, const float* lse, float* dsum_ws, void* dq,
                         long long lddq, void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Nq, int Nk,
                         float scale, void* stream);

/* ---- element-wise / row-wise kernels around the GEMMs -------------------------------------------------------
 * swiglu_bwd: fm_utils.py:143 backward. ab bf16 [R,2H] = [a|b] saved by EPI_SWIGLU, dg bf16 [R,H] -> dab bf16 [R,2H].   */
int b200fm_swiglu_bwd(const void* ab, long long ld_ab, const void* dg, long long ld_dg, void* dab, long long ld_dab,
                      long long R, int H, void* stream);
/* act_bwd: dpre = dact * f'(pre) on bf16 [n]; act 0 = GELU(erf) (fm_utils.py:116-125 Mlp), 1 = tanh (vit post_mlp).     */
int b200fm_act_bwd(int act, const void* pre, const void* dact, void* dpre, long long n, void* stream);
/* cross_entropy: fm.py:597 F.cross_entropy on fp32 logits [n,V] (row stride ld); loss_rows fp32 [n] (per-row NLL);
 * dlogits (optional) bf16 [n,V] = softmax - onehot, UNSCALED (fold 1/n and the upstream grad in via gemm alpha_dev).    */
int b200fm_cross_entropy(const float* logits, long long ld, const int64_t* targets, float* loss_rows, void* dlogits,
                         long long ldd, long long n, int V, void* stream);
/* cross_entropy with a device-side row count: rows >= *n_dev (< n_max) get loss 0; their dlogits rows are zeroed up to the next
 * multiple of 64.  masked_mean: mean_out = sum(x[:n]) / max(n, 1) (0 for n == 0: fm.py:593-595), inv_n_out = 1 / max(n, 1).  */
int b200fm_cross_entropy_dyn(const float* logits, long long ld, const int64_t* targets, float* loss_rows, void* dlogits,
                             long long ldd, long long n_max, int V, const int* n_dev, void* stream);
int b200fm_masked_mean(const float* x, const int* n_dev, long long n_max, float* mean_out, float* inv_n_out, void* stream);
/* Fused masked-token head (fm.py:589-600 `to_logits` + F.cross_entropy): loss_rows[i] = logsumexp(h_i W^T) - (h_i W^T)[target_i] and
 * (optional) dlogits bf16 [M, V] = softmax(h_i W^T) - onehot(target_i) WITHOUT materialising the fp32 logits: the logits tile stays in
 * TMEM / registers; pass 1 writes per-(row, 128-column slot) (max, sum exp) partials into ws (fp32 [M, 2 * head_ce_ws_slots(V)]) and the
 * target logit, a row reduction produces lse / loss, pass 2 recomputes the tile and writes the bf16 gradient directly.
 * n_dev (optional): device-side row count as in gemm_bf16_dyn; rows [n, roundup64(n)) of dlogits are zero-filled.           */
int b200fm_head_ce_ws_slots(int V);
int b200fm_head_ce(const void* h, long long ldh, const void* W, long long ldw, const int64_t* targets, const int* n_dev, int M, int V,
                   int K, float* ws, float* tlogit, float* lse, float* loss_rows, void* dlogits, long long ldd, void* stream);
/* colsum_bf16: out[c] += sum_r x[r,c] (bias gradients of nn.Linear layers with bias).                                   */
int b200fm_colsum_bf16(const void* x, long long ld, float* out, long long R, int N, void* stream);
/* cast_f32_bf16: bf16 shadow of fp32 master weights / activations (what autocast does per call in the reference).       */
int b200fm_cast_f32_bf16(const float* x, void* y, long long n, void* stream);
/* patchify: encoder_embeddings.py:301 rearrange 'b d (nh ph) (nw pw) -> b (nh nw) (ph pw d)', fp32 -> bf16 rows.       */
int b200fm_patchify(const float* img, void* out, int B, int C, int H, int W, int P, void* stream);
/* adamw: torch.optim.AdamW single-tensor step (optim_factory.py:239-240) fused with the bf16 shadow refresh.
 * g is multiplied by grad_scale first (DDP mean / loss scaling).                                                         */
int b200fm_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* Multi-tensor form: ONE launch for all tensors of a param group (same lr / weight decay / step).  table_dev: device array of
 * per-tensor pointers; chunk_tensor_dev / chunk_offset_dev (device, n_chunks entries) assign b200fm_adamw_chunk_elems()
 * consecutive elements of one tensor to each CTA.                                                                          */
typedef struct b200fm_adamw_tensor {
    float* p;            /* fp32 master weight, updated in place */
    const float* g;      /* fp32 gradient                        */
    float* m;            /* exp_avg                               */
    float* v;            /* exp_avg_sq                            */
    void* shadow_bf16;   /* optional bf16 mirror of p (refreshed) */
    long long n;         /* elements                              */
} b200fm_adamw_tensor;
int b200fm_adamw_multi(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                       int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                       void* stream);
int b200fm_adamw_chunk_elems(void);
/* As adamw_multi with the per-step scalars in DEVICE memory, hyper_dev = {lr, 1 - beta1^t, sqrt(1 - beta2^t)}: a captured CUDA graph
 * replays the same launch while the host refreshes the three floats (learning-rate schedule, bias correction) before each replay. */
int b200fm_adamw_multi_dev(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                           int n_chunks, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                           const float* hyper_dev, void* stream);
/* Either of the two above (hyper_dev == NULL: lr / step as arguments) plus *gnorm_sq += sum of the squared scaled gradients of the
 * group -- the gradient norm the reference logs (utils/native_scaler.py:56-65) without another pass over the gradients.  The caller
 * zeroes *gnorm_sq once per step.                                                                                           */
int b200fm_adamw_multi_gnorm(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                             int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             const float* hyper_dev, float* gnorm_sq, void* stream);
/* Exponential moving average of a model in one launch (ModelEmaV2.update, fourm/utils/timm/model_ema.py:123-127; every step of
 * run_training_vqvae.py:1169-1171): per table entry p (EMA tensor) = decay * p + (1 - decay) * g (model tensor), optional bf16 mirror
 * of p refreshed; m / v of the entries are ignored.  one_minus_decay is passed separately because the reference computes 1 - decay in
 * double before rounding it to fp32; products and sum are rounded like the torch expression (bit-identical).        */
int b200fm_ema_multi(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev, int n_chunks,
                     float decay, float one_minus_decay, void* stream);

/* ---- modality-masked token selection + embedding gather / scatter ------------------------------------------------
 * Replaces cat_{encoder,decoder}_tensors + forward_mask_{encoder,decoder} + adapt_decoder_attention_mask
 * (fourm/models/fm.py:245-475) and the embedding module forwards (fourm/models/encoder_embeddings.py:87-121, 184-211,
 * 280-309; fourm/models/decoder_embeddings.py:98-139, 226-255).  One b200fm_segment per modality, in concatenation order
 * (mod_dict order on the encoder side; the Python-`random` shuffled order of fm.py:306 on the decoder side).          */
#define B200FM_MAX_SEGMENTS 24
/* `mode` flags of select_plan / embed_rows / embed_rows_bwd */
#define B200FM_MODE_DECODER 1  /* decoder side: teacher-forcing shift for sequences, mask token for images, targets + dam   */
#define B200FM_MODE_IDENTITY 2 /* keep every position in place (materialise a whole modality: embedding modules' forward)    */
#define B200FM_MODE_NO_SUM 4   /* x0 = x instead of x + emb                                                                 */
#define B200FM_KIND_IMG 0     /* pixel patches: x rows come from the patch-projection GEMM (x_rows)  */
#define B200FM_KIND_TOK_IMG 1 /* tokenised image: ids [B, L]                                         */
#define B200FM_KIND_SEQ 2     /* token sequence: ids [B, L], decoder side uses the teacher-forcing shift */
#define B200FM_KIND_SEQ_EMB 3 /* pre-computed sequence features (T5-XXL, encoder_embeddings.py:312-421): x rows from the
                                 emb_proj GEMM (x_rows) like KIND_IMG, positions ranked among valid inputs like KIND_SEQ */
typedef struct b200fm_segment {
    const uint8_t* mask;      /* [B, L] bool, 1 = masked: input_mask (encoder) / target_mask (decoder)            */
    const void* ids;          /* [B, L] int32 or int64 token ids (NULL for KIND_IMG)                              */
    const int32_t* dam;       /* [B, L] decoder_attention_mask (decoder side only)                                */
    const float* token_emb;   /* [V, D] fp32 table                                                                */
    const float* pos_emb;     /* [P, D] fp32 positional table                                                     */
    const float* mod_emb;     /* [D]                                                                              */
    const void* x_rows;       /* KIND_IMG / KIND_SEQ_EMB: bf16 [B*L, D] projected patches / features              */
    float* d_token_emb;       /* backward: [V, D] fp32 gradient table, accumulated into (may be NULL)             */
    float* d_mod_emb;         /* backward: [D] fp32, accumulated into (may be NULL)                               */
    void* dx_rows;            /* backward, KIND_IMG: bf16 [B*L, D] gradient rows (pre-zeroed by the caller)       */
    long long padding_idx;    /* nn.Embedding padding_idx (its gradient row stays zero); -1 = none                */
    int L;                    /* positions per sample in the raw tensors                                          */
    int kind;
    int mod_id;               /* generate_uint15_hash(name), fourm/utils/misc.py:39-41                            */
    int max_length;           /* decoder sequences: positions >= max_length wrap to 0 (decoder_embeddings.py:128) */
    int ids_is_i64;
    int reserved;
    float* d_pos_emb;         /* backward: [P, D] fp32 gradient of a LEARNABLE positional table (sincos_pos_emb=False), accumulated
                                 into at row pos_id (NULL for the sincos buffers)                                 */
} b200fm_segment;

/* Stable partition "first n_keep valid positions, then masked ones, in order" per sample (== argsort(mask+arange*1e-6)[:, :n_keep]).
 * Outputs [B, n_keep]: src_seg (segment index), src_pos (position inside the segment), pos_id (positional-table row, -1 = none),
 * pad_mask (1 = padded slot: the reference's encoder_mask / decoder_mask), mod_mask (int16, -1 on pads), mod_raw (before the
 * -1 assignment: needed by the attention mask), and on the decoder side target_ids (int64, 0 on pads) and dam (int32).     */
int b200fm_select_plan(const b200fm_segment* segs, int n_seg, int mode, int B, int n_keep, int32_t* src_seg, int32_t* src_pos,
                       int32_t* pos_id, uint8_t* pad_mask, int16_t* mod_mask, int16_t* mod_raw, int64_t* target_ids,
                       int32_t* dam_out, void* stream);
/* As select_plan with the concatenation order as DEVICE data: segs[] is given in a fixed (canonical) order and order_dev[i] names the
 * segment that comes i-th (the Python-random decoder shuffle of fm.py:306); src_seg stores indices into segs[].  order_dev NULL ==
 * select_plan.  Lets a captured CUDA graph replay the step with a fresh shuffle.                                                   */
int b200fm_select_plan_ordered(const b200fm_segment* segs, int n_seg, int mode, int B, int n_keep, int32_t* src_seg, int32_t* src_pos,
                               int32_t* pos_id, uint8_t* pad_mask, int16_t* mod_mask, int16_t* mod_raw, int64_t* target_ids,
                               int32_t* dam_out, const int32_t* order_dev, void* stream);
/* mask_out uint8 [B, M, M], 1 = masked: (j >= cumsum(dam)[i]) | (mod_raw[i] != mod_raw[j])  (or triu(1) if causal).          */
int b200fm_decoder_attention_mask(const int32_t* dam, const int16_t* mod_raw, uint8_t* mask_out, int B, int M, int causal, int sep,
                                  void* stream);
/* x0 fp32 [B, n_keep, D] = x + emb, emb_out (optional) fp32 = pos + mod; padded slots are zero.                               */
int b200fm_embed_rows(const b200fm_segment* segs, int n_seg, int mode, const int32_t* src_seg, const int32_t* src_pos,
                      const int32_t* pos_id, const uint8_t* pad_mask, const float* mask_token, float* x0, float* emb_out, int B,
                      int n_keep, int D, void* stream);
/* Backward of embed_rows: dx0 (and demb, optional) fp32 [B, n_keep, D] -> scatter-add into d_token_emb / d_mod_emb / dx_rows of
 * each segment and d_mask_token (decoder image modalities); d_pos_emb rows by pos_id when a segment's table is learnable.                                                                 */
int b200fm_embed_rows_bwd(const b200fm_segment* segs, int n_seg, int mode, const int32_t* src_seg, const int32_t* src_pos,
                          const int32_t* pos_id, const uint8_t* pad_mask, const float* dx0, const float* demb, float* d_mask_token, int B, int n_keep,
                          int D, void* stream);
/* Masked-token head index sets (fm.py:589-600 `y[decoder_mod_mask == idx]`): rows_out int32 [n_mods, n_rows] (row-major order,
 * first counts[m] entries valid), counts int32 [n_mods].  mod_ids_dev: device int32 [n_mods].                              */
int b200fm_head_rows(const int16_t* mod_mask, long long n_rows, const int* mod_ids_dev, int n_mods, int32_t* rows_out,
                     int32_t* counts, void* stream);
int b200fm_gather_rows_bf16(const void* src, const int32_t* rows, void* out, long long n, int D, void* stream);
int b200fm_gather_i64(const int64_t* src, const int32_t* rows, int64_t* out, long long n, void* stream);
/* Device-side counts (n = min(*n_dev, n)): gather_rows zero-fills output rows [n, roundup(n, 128)).                          */
int b200fm_gather_rows_bf16_dyn(const void* src, const int32_t* rows, void* out, long long n, int D, const int* n_dev, void* stream);
int b200fm_gather_i64_dyn(const int64_t* src, const int32_t* rows, int64_t* out, long long n, const int* n_dev, void* stream);
int b200fm_scatter_rows_bf16_dyn(const void* src, const int32_t* rows, void* dst, long long n, int D, const int* n_dev, void* stream);
/* Nucleus sampling of one token per row (generate.py:332-371 top_k_top_p_filtering + softmax(. / temperature) + multinomial, per
 * generated token): logits fp32 [rows, V] (row stride ld), u fp32 [rows] uniform in [0, 1) from the caller's generator, out int64 [rows].
 * Keeps token i iff the probability mass ranked strictly before it is <= top_p (the reference's rule; top_p <= 0 or >= 1: all tokens),
 * then draws from softmax(kept / temperature) by inverse CDF in index order.  temperature > 0.                          */
int b200fm_sample_top_p(const float* logits, long long ld, int rows, int V, float top_p, float temperature, const float* u,
                        int64_t* out, void* stream);
/* K/V cache append of the decode loop (replaces the growing torch.cat of generate.py's per-token forward): cache bf16 [B, L, row_w];
 * cache[b, *pos_dev, col0 : col0 + width] = src[b, 0 : width] (src bf16 [B, >= width], row stride ld_src).  *pos_dev int64 on the device. */
int b200fm_kv_append(const void* src, long long ld_src, void* cache, long long L, long long row_w, const int64_t* pos_dev, int B,
                     int width, int col0, void* stream);
/* dst bf16 [*, D] row rows[i] = src bf16 [n, D] row i (backward of gather_rows_bf16; distinct destination rows).           */
int b200fm_scatter_rows_bf16(const void* src, const int32_t* rows, void* dst, long long n, int D, void* stream);
/* dst fp32 [*, D] rows[i] += src bf16 [n, D] row i (distinct destination rows).                                             */
int b200fm_scatter_add_rows(const void* src_bf16, const int32_t* rows, float* dst, long long n, int D, void* stream);

/* ---- VQ codebook scan (fourm/vq/quantizers/quantize_lucid.py:388-407 cosine, :263-284 Euclidean) -------------
 * z fp32 [n, d], codebook fp32 [K, d] (d <= 64, multiple of 4).  cosine != 0: both sides are l2-normalised (eps 1e-12,
 * F.normalize) and the arg-max of the dot product is taken; else arg-max of -(|z|^2 - 2 z.e + |e|^2).  fp32 FMA,
 * ties -> lowest index.  idx_out int64 [n].  quant_out (optional) fp32 [n, d] = codebook[idx] (un-normalised rows). */
int b200fm_vq_argmax(const float* z, const float* codebook, int64_t* idx_out, float* quant_out, long long n, int K, int d,
                     int cosine, void* stream);
/* Same through HOST buffers (pageable or pinned): H2D of z, scan, D2H of idx -- the call save_vq_tokens.py would make. */
int b200fm_vq_argmax_host(const float* z_host, const float* codebook_dev, int64_t* idx_host, long long n, int K, int d,
                          int cosine, void* stream);

/* Training-side codebook statistics (quantize_lucid.py:404-419 / 286-292) without the one-hot [n, K]: bins[idx[r]] += 1,
 * embed_sum[idx[r], :] += (cosine ? l2norm(z[r]) : z[r]).  bins fp32 [K], embed_sum fp32 [K, d] are ACCUMULATED into (zero
 * them first; they may be two slices of one packed buffer so that sync_codebook needs a single all-reduce).               */
int b200fm_vq_ema_stats(const float* z, const long long* idx, long long n, int K, int d, int cosine, float* bins, float* embed_sum,
                        void* stream);
/* Cosine codebook EMA (quantize_lucid.py:413-425), in place: cluster_size = cs*decay + bins*(1-decay);
 * embed[k] = embed[k]*decay + (bins[k] == 0 ? l2norm(embed[k]) : l2norm(embed_sum[k] / bins[k])) * (1-decay).             */
int b200fm_vq_ema_update_cosine(float* embed, float* cluster_size, const float* bins, const float* embed_sum, int K, int d, float decay,
                                void* stream);

/* ---- fp32-faithful inference ("precise mode": callers that run the reference without bf16 autocast, save_vq_tokens.py:288) ------
 * split_limbs: x fp32 [rows, K] (row stride ldx) -> bf16 [rows, terms * K]: the limb layout of the A (role 0) or B (role 1) operand
 * of a bf16 GEMM over the contraction length terms * K that reproduces the fp32 product (terms 3: ~2^-16 relative, terms 6: fp32
 * class).  attention_f32: softmax(q k^T * scale (masked)) v in fp32 FMA arithmetic, head_dim 64, same mask addressing as attention_fwd. */
int b200fm_split_limbs(const float* x, long long ldx, void* out_bf16, long long rows, int K, int terms, int role, void* stream);
int b200fm_attention_f32(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, const uint8_t* mask,
                         long long mask_b_stride, long long mask_q_stride, float* out, long long ldo, int B, int H, int Nq, int Nk,
                         float scale, void* stream);

/* ---- batch assembly on the device (fourm/data/masking.py:236-266 UnifiedMasking.image_mask; RGB loader normalisation) -------------
 * mask_images: one row per (sample, image-like modality): noise fp32 [rows, L] (the reference's torch.rand), in_budget / tgt_budget
 * int32 [rows] (tgt_budget NULL or < 0 = the reference's target_budget None) -> input_mask / target_mask uint8 [rows, L] (1 = masked),
 * decoder_attention_mask int32 [rows, L].  Bit-exact with the reference for the same noise.
 * patchify_u8: uint8 RGB [B,3,H,W] -> bf16 patches like b200fm_patchify, normalised on the fly (x/255 - mean[c]) / std[c]
 * (mean3 / std3: HOST arrays of 3 floats), so a batch travels as 1 byte per value instead of 4.                                  */
int b200fm_mask_images(const float* noise, const int32_t* in_budget, const int32_t* tgt_budget, uint8_t* input_mask,
                       uint8_t* target_mask, int32_t* dam, long long rows, int L, void* stream);
int b200fm_patchify_u8(const uint8_t* img, void* out, int B, int C, int H, int W, int P, const float* mean3, const float* std3,
                       void* stream);

/* ---- data-parallel gradient all-reduce over NVLink peer memory ------------------------------------------------------------
 * Replaces the NCCL all-reduce torch DDP issues for the reference's train step (run_training_4m.py:512; fp32 gradients, mean).
 * Every rank allocates one arena with comm_alloc (cudaMalloc, so it can be exported), exports it with comm_ipc_export (64-byte
 * cudaIpcMemHandle_t), opens the peers' handles with comm_ipc_open.  The first comm_flag_bytes() bytes of each arena are the
 * flag block; gradients live behind it at IDENTICAL offsets on all ranks.
 * allreduce_f32: two-shot all-reduce (P2P loads of the caller's 1/world shard from all ranks in rank order, sum, * scale, P2P
 * stores into all ranks) of data[offset_elems .. offset_elems + n_elems) in ONE launch of n_ctas CTAs, bracketed by per-CTA
 * cross-rank flag barriers.  peer_data / peer_flags: HOST arrays of `world` device pointers (index = rank; own entry = own
 * arena).  seq: strictly increasing per call (same value on all ranks).  All ranks must issue the same calls in the same order. */
int b200fm_comm_flag_bytes(void);
int b200fm_comm_alloc(long long bytes, void** ptr);
int b200fm_comm_free(void* ptr);
int b200fm_comm_ipc_export(const void* ptr, void* handle64);
int b200fm_comm_ipc_open(const void* handle64, void** ptr);
int b200fm_comm_ipc_close(void* ptr);
int b200fm_allreduce_f32(void* const* peer_data, void* const* peer_flags, int rank, int world, long long offset_elems,
                         long long n_elems, float scale, unsigned int seq, int n_ctas, void* stream);
/* As above with seq + *seq_base_dev as the sequence number (seq_base_dev: device uint32 the caller advances once per step), so the
 * launch can be captured in a CUDA graph and replayed.                                                                         */
int b200fm_allreduce_f32_seq(void* const* peer_data, void* const* peer_flags, int rank, int world, long long offset_elems,
                             long long n_elems, float scale, unsigned int seq, const unsigned int* seq_base_dev, int n_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FM_H_ */
