// fp32-faithful inference path ("precise mode") for callers that run the reference WITHOUT bf16 autocast: the VQ tokenizers
// (save_vq_tokens.py:288 runs VQ.tokenize in fp32) -- their arg-min over 16k codes flips on ~25 % of the positions when the ViT
// encoder runs with bf16 operands, so token-level parity needs (near-)fp32 contractions.
//
// Contractions stay on the tcgen05 bf16 GEMM: every fp32 operand is split into bf16 limbs x = h + m (+ l) (each limb the bf16
// rounding of the remaining residual) and the product is assembled from limb products along a LONGER contraction dimension:
//   2 limbs, 3 terms:  [Ah | Am | Ah] . [Bh | Bh | Bm]^T                          (drops Am.Bm: ~2^-16 relative)
//   3 limbs, 6 terms:  [Ah | Am | Al | Ah | Am | Ah] . [Bh | Bh | Bh | Bm | Bm | Bl]^T   (drops terms <= 2^-24: fp32 class)
// limb products are exact in the fp32 accumulator.  The attention core (4 N^2 d FLOPs, ~5 % of a ViT) runs as a plain fp32 FMA kernel.
#include <cfloat>

#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

// out [rows, terms * K] bf16; role 0 = A operand, 1 = B operand; terms in {3, 6}
B200FM_DEVINL void limbs_of(float v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    m = __float2bfloat16_rn(r1);
    l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}
__global__ void split_limbs_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long rows, int K, int terms, int role) {
    pdl_enter();
    const long long total = rows * K;
    const long long ldo = (long long)terms * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K; const int c = (int)(i % K);
        __nv_bfloat16 h, m, l;
        limbs_of(x[r * ldx + c], h, m, l);
        __nv_bfloat16* o = out + r * ldo + c;
        if (terms == 3) {
            if (role == 0) { o[0] = h; o[K] = m; o[2 * K] = h; }
            else           { o[0] = h; o[K] = h; o[2 * K] = m; }
        } else {
            if (role == 0) { o[0] = h; o[K] = m; o[2 * K] = l; o[3 * K] = h; o[4 * K] = m; o[5 * K] = h; }
            else           { o[0] = h; o[K] = h; o[2 * K] = h; o[3 * K] = m; o[4 * K] = m; o[5 * K] = l; }
        }
    }
}
// 16-byte version (K % 8 == 0, 16-byte aligned rows): one thread = 8 consecutive elements, two float4 loads, `terms` uint4 stores
__global__ void __launch_bounds__(256)
split_limbs_vec_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long rows, int K, int terms, int role) {
    pdl_enter();
    const int kc = K / 8;
    const long long total = rows * kc;
    const long long ldo = (long long)terms * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / kc; const int c = (int)(i % kc) * 8;
        const float4 v0 = __ldcs(reinterpret_cast<const float4*>(x + r * ldx + c)), v1 = __ldcs(reinterpret_cast<const float4*>(x + r * ldx + c) + 1);
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        __align__(16) __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) limbs_of(v[e], h[e], m[e], l[e]);
        const uint4 H = *reinterpret_cast<const uint4*>(h), M = *reinterpret_cast<const uint4*>(m), L = *reinterpret_cast<const uint4*>(l);
        __nv_bfloat16* o = out + r * ldo + c;
        auto st = [&](int slot, const uint4& w) { *reinterpret_cast<uint4*>(o + (long long)slot * K) = w; };
        if (terms == 3) {
            if (role == 0) { st(0, H); st(1, M); st(2, H); }
            else           { st(0, H); st(1, H); st(2, M); }
        } else {
            if (role == 0) { st(0, H); st(1, M); st(2, L); st(3, H); st(4, M); st(5, H); }
            else           { st(0, H); st(1, H); st(2, H); st(3, M); st(4, M); st(5, L); }
        }
    }
}

// fp32 attention, head_dim 64: one thread per query row (q and the output accumulator in registers), 64-key tiles of K / V in
// shared memory (every lane reads the same address: broadcast), online softmax.  mask: optional uint8, 1 = masked (-FLT_MAX fill).
constexpr int kF32Tile = 64;
__global__ void __launch_bounds__(64)
attention_f32_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk, const float* __restrict__ v,
                     long long ldv, const uint8_t* __restrict__ mask, long long mask_b_stride, long long mask_q_stride, float* __restrict__ out,
                     long long ldo, int Nq, int Nk, float scale) {
    pdl_enter();
    __shared__ __align__(16) float ks[kF32Tile][64];
    __shared__ __align__(16) float vs[kF32Tile][64];
    const int b = blockIdx.z, h = blockIdx.y, row = blockIdx.x * 64 + threadIdx.x;
    const bool ok = row < Nq;
    float qr[64], o[64];
    const float* qp = q + ((long long)b * Nq + (ok ? row : 0)) * ldq + h * 64;
#pragma unroll
    for (int d = 0; d < 64; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(qp + d);
        qr[d] = t.x * scale; qr[d + 1] = t.y * scale; qr[d + 2] = t.z * scale; qr[d + 3] = t.w * scale;
        o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
    }
    const uint8_t* mrow = mask ? mask + b * mask_b_stride + (ok ? row : 0) * mask_q_stride : nullptr;
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < Nk; k0 += kF32Tile) {
        __syncthreads();
        for (int i = threadIdx.x; i < kF32Tile * 16; i += 64) {
            const int j = i >> 4, d4 = (i & 15) * 4;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (k0 + j < Nk) {
                kk = *reinterpret_cast<const float4*>(k + ((long long)b * Nk + k0 + j) * ldk + h * 64 + d4);
                vv = *reinterpret_cast<const float4*>(v + ((long long)b * Nk + k0 + j) * ldv + h * 64 + d4);
            }
            *reinterpret_cast<float4*>(&ks[j][d4]) = kk;
            *reinterpret_cast<float4*>(&vs[j][d4]) = vv;
        }
        __syncthreads();
        const int nj = min(kF32Tile, Nk - k0);
        // 16-key chunks, fully unrolled: the chunk's scores stay in registers (an indexed 64-entry array would live in local memory)
#pragma unroll 1
        for (int c0 = 0; c0 < kF32Tile; c0 += 16) {
            if (c0 >= nj) break;
            float s[16];
            float tmax = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int j = c0 + jj;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;                  // four independent chains: FMA latency is hidden by ILP
#pragma unroll
                for (int d = 0; d < 64; d += 4) {
                    const float4 kk = *reinterpret_cast<const float4*>(&ks[j][d]);
                    a0 = fmaf(qr[d], kk.x, a0); a1 = fmaf(qr[d + 1], kk.y, a1); a2 = fmaf(qr[d + 2], kk.z, a2); a3 = fmaf(qr[d + 3], kk.w, a3);
                }
                float acc = (a0 + a1) + (a2 + a3);
                if (j >= nj) acc = -INFINITY;                                    // tile padding: excluded exactly
                else if (mrow && mrow[k0 + j]) acc = -FLT_MAX;                    // masked_fill(-finfo.max), fm_utils.py:169
                s[jj] = acc;
                tmax = fmaxf(tmax, acc);
            }
            const float m_new = fmaxf(m, tmax);
            const float corr = (m == -INFINITY) ? 0.f : expf(m - m_new);
            l *= corr;
#pragma unroll
            for (int d = 0; d < 64; ++d) o[d] *= corr;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int j = c0 + jj;
                const float p = (j < nj) ? expf(s[jj] - m_new) : 0.f;
                l += p;
#pragma unroll
                for (int d = 0; d < 64; d += 4) {
                    const float4 vv = *reinterpret_cast<const float4*>(&vs[j][d]);
                    o[d] = fmaf(p, vv.x, o[d]); o[d + 1] = fmaf(p, vv.y, o[d + 1]); o[d + 2] = fmaf(p, vv.z, o[d + 2]); o[d + 3] = fmaf(p, vv.w, o[d + 3]);
                }
            }
            m = m_new;
        }
    }
    if (ok) {
        const float inv = 1.0f / l;
        float* op = out + ((long long)b * Nq + row) * ldo + h * 64;
#pragma unroll
        for (int d = 0; d < 64; d += 4) *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    }
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_split_limbs(const float* x, long long ldx, void* out_bf16, long long rows, int K, int terms, int role, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0 || K == 0) return 0;
    B200FM_CHECK(x && out_bf16, "split_limbs: null pointer");
    B200FM_CHECK((terms == 3 || terms == 6) && (role == 0 || role == 1), "split_limbs: terms must be 3 or 6, role 0 (A) or 1 (B)");
    const bool vec = K % 8 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0;
    const long long total = vec ? rows * (K / 8) : rows * K;
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    if (vec) B200FM_LAUNCH(split_limbs_vec_kernel, dim3(grid), dim3(256), 0, stream, 1, x, ldx, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, K, terms, role);
    else B200FM_LAUNCH(split_limbs_kernel, dim3(grid), dim3(256), 0, stream, 1, x, ldx, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, K, terms, role);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_attention_f32(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
                                    const uint8_t* mask, long long mask_b_stride, long long mask_q_stride, float* out, long long ldo, int B,
                                    int H, int Nq, int Nk, float scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0 || H == 0 || Nq == 0) return 0;
    B200FM_CHECK(q && k && v && out && Nk >= 1, "attention_f32: bad arguments");
    B200FM_CHECK(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "attention_f32: row strides must be multiples of 4 floats");
    B200FM_CHECK(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                 "attention_f32: pointers must be 16-byte aligned");
    B200FM_LAUNCH(attention_f32_kernel, dim3((Nq + 63) / 64, H, B), dim3(64), 0, stream, 1, q, ldq, k, ldk, v, ldv, mask, mask_b_stride, mask_q_stride, out,
                  ldo, Nq, Nk, scale);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
