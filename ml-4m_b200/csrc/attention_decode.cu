// Attention for ONE query row per sequence (head_dim 64): the shape of every attention call inside the K/V-cached autoregressive
// decode step (b200fm/decode.py; reference: the per-token decoder forward of fourm/models/generate.py:886-901, whose attention is
// fm_utils.py:160-180 / 197-219 over the whole prefix).  The tile kernel (attention_fwd.cu) spends a 128-query TMEM tile, TMA
// descriptors and mbarrier round trips on a single row; here a CTA of 256 threads per (batch, head) does it with plain loads:
//   pass 1: thread t scores the keys t, t + 256, ... (q in registers, one 128-byte K row per key), scores in shared memory;
//   block max / sum (fp32, exact softmax like the reference: masked keys are filled with a large negative BEFORE the softmax, so a
//   fully masked row becomes uniform);
//   pass 2: thread = (8 output dimensions, one of 32 key groups): 8 threads read one 128-byte V row, every thread accumulates its 8
//   dimensions over the keys of its group (Nk / 32 iterations, 16-byte loads), the 32 partial rows are summed through shared memory.
// Latency-bound by construction (B x H CTAs, a few KB each): what matters is that it is ONE short kernel; the host side uses it up to
// a few thousand keys (self-attention cache, short contexts) and the tile kernels beyond.  HBM bytes: Nk * 256 per (b, h).
#include <cfloat>

#include "../../include/b200fm.h"
#include "attention_common.cuh"
#include "common.cuh"

namespace b200fm {

constexpr int kDecThreads = 256;

__global__ void __launch_bounds__(kDecThreads)
attention_decode_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k, long long ldk,
                        const __nv_bfloat16* __restrict__ v, long long ldv, const uint8_t* __restrict__ mask, long long mask_b_stride,
                        __nv_bfloat16* __restrict__ out, long long ldo, int Nk, float scale) {
    pdl_enter();
    extern __shared__ float sc[];                         // [Nk] scores, then probabilities
    __shared__ float red[kDecThreads / 32];
    __shared__ float part[kDecThreads / 8][64 + 1];
    const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // q row of this head: 64 bf16 = 8 x 16 B, pre-scaled
    float qf[64];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(q + (long long)b * ldq + h * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 w = __ldg(qp + i);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(ww[e]);
                qf[i * 8 + e * 2] = f.x * scale;
                qf[i * 8 + e * 2 + 1] = f.y * scale;
            }
        }
    }
    const uint8_t* mrow = mask ? mask + b * mask_b_stride : nullptr;
    float mx = -FLT_MAX;
    for (int j = t; j < Nk; j += kDecThreads) {
        const uint4* kp = reinterpret_cast<const uint4*>(k + ((long long)b * Nk + j) * ldk + h * 64);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 w = __ldg(kp + i);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(ww[e]);
                a0 = fmaf(qf[i * 8 + e * 2], f.x, a0);
                a1 = fmaf(qf[i * 8 + e * 2 + 1], f.y, a1);
            }
        }
        float s = a0 + a1;
        if (mrow != nullptr && mrow[j] != 0) s = -FLT_MAX;               // masked_fill(mask, -finfo.max), fm_utils.py:169
        sc[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    {
        float m2 = lane < kDecThreads / 32 ? red[lane] : -FLT_MAX;
        mx = warp_max(m2);
    }
    float sum = 0.f;
    for (int j = t; j < Nk; j += kDecThreads) {
        const float p = __expf(sc[j] - mx);                               // all keys masked: every p = 1 -> uniform, like the reference
        sc[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    __syncthreads();                                                      // red[] reuse + sc[] complete
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    float tot = lane < kDecThreads / 32 ? red[lane] : 0.f;
    tot = warp_sum(tot);
    const float inv = 1.0f / tot;
    // pass 2: dg = 8 output dimensions, kg = key group
    const int dg = t & 7, kg = t >> 3;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const __nv_bfloat16* vp = v + (long long)b * Nk * ldv + h * 64 + dg * 8;
    const bool vec = (ldv % 8) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0;
#pragma unroll 4
    for (int j = kg; j < Nk; j += kDecThreads / 8) {
        const float p = sc[j];
        uint32_t ww[4];
        if (vec) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(vp + (long long)j * ldv));
            ww[0] = w.x; ww[1] = w.y; ww[2] = w.z; ww[3] = w.w;
        } else {
            const uint16_t* r = reinterpret_cast<const uint16_t*>(vp + (long long)j * ldv);
#pragma unroll
            for (int e = 0; e < 4; ++e) ww[e] = (uint32_t)r[2 * e] | ((uint32_t)r[2 * e + 1] << 16);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(ww[e]);
            acc[2 * e] = fmaf(p, f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(p, f.y, acc[2 * e + 1]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[kg][dg * 8 + e] = acc[e];
    __syncthreads();
    if (t < 64) {
        float o = 0.f;
#pragma unroll 8
        for (int g = 0; g < kDecThreads / 8; ++g) o += part[g][t];
        out[(long long)b * ldo + h * 64 + t] = __float2bfloat16_rn(o * inv);
    }
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_attention_decode(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                       const uint8_t* mask, long long mask_b_stride, void* out, long long ldo, int B, int H, int Nk,
                                       float scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0 || H == 0) return 0;
    B200FM_CHECK(q && k && v && out && Nk >= 1, "attention_decode: bad arguments");
    B200FM_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k)) & 15) == 0,
                 "attention_decode: q / k rows must be 16-byte aligned (strides multiples of 8 elements)");
    const size_t smem = (size_t)Nk * sizeof(float);
    B200FM_CHECK(smem <= 200 * 1024, "attention_decode: Nk=%d does not fit the score buffer (max 51200 keys)", Nk);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        B200FM_CUDA(cudaFuncSetAttribute(attention_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    B200FM_LAUNCH(attention_decode_kernel, dim3(H, B), dim3(kDecThreads), smem, stream, 1, reinterpret_cast<const __nv_bfloat16*>(q), ldq,
                  reinterpret_cast<const __nv_bfloat16*>(k), ldk, reinterpret_cast<const __nv_bfloat16*>(v), ldv, mask, mask_b_stride,
                  reinterpret_cast<__nv_bfloat16*>(out), ldo, Nk, scale);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
