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

const char* b200fm_last_error(void);
int b200fm_abi_version(void);
int b200fm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);

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

/* ---- LayerNorm (fourm/models/fm_utils.py:93-108; F.layer_norm, fp32 statistics) -----------------------------
 * x fp32 [rows, D] -> y (bf16 if y_is_bf16 else fp32) [rows, D]; mean/rstd fp32 [rows] saved for backward (may be NULL). */
int b200fm_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean,
                         float* rstd, int rows, int D, float eps, void* stream);
/* dy (bf16 if dy_is_bf16 else fp32) is the gradient w.r.t. the LN output; dx_out fp32 = (dres ? dres : 0) + LN backward;
 * dx_bf16 (optional) = bf16 copy of dx_out; dgamma/dbeta fp32 [D] are ACCUMULATED into (atomics), may be NULL.   */
int b200fm_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                         const float* rstd, const float* dres, float* dx_out, void* dx_bf16, float* dgamma, float* dbeta,
                         int rows, int D, void* stream);

/* ---- fused multi-head attention (fourm/models/fm_utils.py:160-180, 197-219; vq/models/vit_models.py:165-197) --
 * q rows [B*Nq, ldq], k/v rows [B*Nk, ldk/ldv] bf16, head h occupies columns [h*64, h*64+64) from each base pointer
 * (so q/k/v may alias one packed qkv buffer).  mask: uint8/bool, 1 = masked (reference convention), addressed as
 * mask[b*mask_b_stride + i*mask_q_stride + j]; NULL = no mask.  Masked scores are filled with -FLT_MAX-like
 * finite value (fully masked rows -> uniform attention, like masked_fill(-finfo.max)).
 * out bf16 [B*Nq, ldo]; lse fp32 [B, H, Nq] (saved for backward; natural log).  head_dim is fixed at 64.        */
int b200fm_attention_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                         const uint8_t* mask, long long mask_b_stride, long long mask_q_stride, void* out, long long ldo,
                         float* lse, int B, int H, int Nq, int Nk, float scale, void* stream);
int b200fm_attention_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                         const uint8_t* mask, long long mask_b_stride, long long mask_q_stride, const void* out,
                         long long ldo, const void* dout, long long lddo, const float* lse, void* dq, long long lddq,
                         void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Nq, int Nk, float scale,
                         void* stream);

/* ---- VQ codebook scan (fourm/vq/quantizers/quantize_lucid.py:388-407 cosine, :263-284 Euclidean) -------------
 * z fp32 [n, d], codebook fp32 [K, d] (d <= 64, multiple of 4).  cosine != 0: both sides are l2-normalised (eps 1e-12,
 * F.normalize) and the arg-max of the dot product is taken; else arg-max of -(|z|^2 - 2 z.e + |e|^2).  fp32 FMA,
 * ties -> lowest index.  idx_out int64 [n].  quant_out (optional) fp32 [n, d] = codebook[idx] (un-normalised rows). */
int b200fm_vq_argmax(const float* z, const float* codebook, int64_t* idx_out, float* quant_out, long long n, int K, int d,
                     int cosine, void* stream);
/* Same through HOST buffers (pageable or pinned): H2D of z, scan, D2H of idx -- the call save_vq_tokens.py would make. */
int b200fm_vq_argmax_host(const float* z_host, const float* codebook_dev, int64_t* idx_host, long long n, int K, int d,
                          int cosine, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FM_H_ */
