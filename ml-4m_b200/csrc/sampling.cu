// Fused nucleus (top-p) sampling of one token per row of fp32 logits: the sampling step of the autoregressive generation loop
// (fourm/models/generate.py:332-371 top_k_top_p_filtering + softmax(filtered / temperature) + torch.multinomial, called once per
// generated token at :1003-1008).  The reference (and the overlay's torch path) spends a full sort of the vocabulary, a cumsum, a
// scatter, two softmaxes and a multinomial -- ~10 launches -- per token; here one CTA per row keeps the row in shared memory:
//   1. row max and Z = sum exp(l - max);
//   2. nucleus: the reference keeps token i iff the probability mass ranked strictly before it is <= top_p.  With
//      M(x) = sum of p_j over { l_j > x } that is M(l_i) <= top_p * Z, a monotone predicate in l_i: the cut is found by a 32-step
//      bisection over the order-preserving integer image of the floats (exact: no sort, ties are kept or dropped together);
//   3. one draw from softmax(kept / temperature) by inverse CDF in index order with the caller's uniform number u in [0, 1)
//      (torch.rand on the device generator): same distribution as multinomial, a different use of the random stream.
// top_p <= 0 keeps everything.  HBM bytes: 4 V per row read once.
#include <cfloat>

#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

constexpr int kSampleThreads = 1024;

B200FM_DEVINL uint32_t float_key(float f) {              // order-preserving map float -> uint32
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

B200FM_DEVINL float block_sum(float v, float* red, int warp, int lane) {
    v = warp_sum(v);
    __syncthreads();                                     // previous use of red[] is over
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = lane < kSampleThreads / 32 ? red[lane] : 0.f;
    return warp_sum(s);                                  // every warp computes the same total
}

__global__ void __launch_bounds__(kSampleThreads)
sample_top_p_kernel(const float* __restrict__ logits, long long ld, int V, float top_p, float inv_temperature, const float* __restrict__ u,
                    long long* __restrict__ out) {
    pdl_enter();
    extern __shared__ float sl[];                        // [V] logits of this row
    __shared__ float red[kSampleThreads / 32];
    __shared__ float scan[kSampleThreads / 32];
    const int row = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const float* lr = logits + (long long)row * ld;
    float mx = -FLT_MAX;
    for (int i = t; i < V; i += kSampleThreads) {
        const float l = lr[i];
        sl[i] = l;
        mx = fmaxf(mx, l);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = lane < kSampleThreads / 32 ? red[lane] : -FLT_MAX;
    mx = warp_max(mx);
    float z = 0.f;
    for (int i = t; i < V; i += kSampleThreads) z += __expf(sl[i] - mx);
    z = block_sum(z, red, warp, lane);
    // nucleus cut K*: the smallest key K with M(K) = sum_{key_i > K} p_i <= top_p * Z; kept = { key_i >= K* }
    uint32_t cut = 0u;
    if (top_p > 0.f && top_p < 1.f) {
        const float target = top_p * z;
        uint32_t lo = 0u, hi = 0xffffffffu;              // M(hi) = 0 <= target always holds
        for (int it = 0; it < 32; ++it) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            float m = 0.f;
            for (int i = t; i < V; i += kSampleThreads) {
                const float l = sl[i];
                if (float_key(l) > mid) m += __expf(l - mx);
            }
            m = block_sum(m, red, warp, lane);
            if (m <= target) hi = mid; else lo = mid + 1u;
            if (lo >= hi) break;
        }
        cut = hi;
    }
    // inverse-CDF draw over the kept tokens, weights exp((l - max) / T); thread t owns the contiguous segment [t * seg, (t + 1) * seg)
    const int seg = (V + kSampleThreads - 1) / kSampleThreads;
    const int i0 = t * seg, i1 = min(V, i0 + seg);
    float local = 0.f;
    for (int i = i0; i < i1; ++i) {
        const float l = sl[i];
        if (float_key(l) >= cut) local += __expf((l - mx) * inv_temperature);
    }
    // block exclusive scan of `local`
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    __syncthreads();
    if (lane == 31) scan[warp] = incl;
    __syncthreads();
    float wbase = 0.f, total = 0.f;
    for (int w = 0; w < kSampleThreads / 32; ++w) {
        const float s = scan[w];
        if (w < warp) wbase += s;
        total += s;
    }
    const float base = wbase + incl - local;
    const float r = fminf(u[row], 0.99999994f) * total;
    __shared__ int found;
    if (t == 0) found = 0;
    __syncthreads();
    if (local > 0.f && r >= base && r < base + local) {
        float c = base;
        int pick = -1, last_kept = -1;
        for (int i = i0; i < i1; ++i) {
            const float l = sl[i];
            if (float_key(l) >= cut) {
                last_kept = i;
                c += __expf((l - mx) * inv_temperature);
                if (r < c) { pick = i; break; }
            }
        }
        if (pick < 0) pick = last_kept;                  // rounding at the segment's end
        if (atomicExch(&found, 1) == 0) out[row] = pick;
    }
    __syncthreads();
    if (t == 0 && found == 0) {                          // r landed on a segment boundary by rounding: take the arg-max (always kept)
        int best = 0;
        for (int i = 1; i < V; ++i) if (sl[i] > sl[best]) best = i;
        out[row] = best;
    }
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_sample_top_p(const float* logits, long long ld, int rows, int V, float top_p, float temperature, const float* u,
                                   int64_t* out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return 0;
    B200FM_CHECK(logits && u && out && V >= 1, "sample_top_p: bad arguments");
    B200FM_CHECK(temperature > 0.f, "sample_top_p: temperature must be > 0 (use arg-max for temperature 0)");
    const size_t smem = (size_t)V * sizeof(float);
    B200FM_CHECK(smem <= 200 * 1024, "sample_top_p: V=%d does not fit shared memory (max 51200)", V);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        B200FM_CUDA(cudaFuncSetAttribute(sample_top_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    B200FM_LAUNCH(sample_top_p_kernel, dim3(rows), dim3(kSampleThreads), smem, stream, 1, logits, ld, V, top_p, 1.0f / temperature, u,
                  reinterpret_cast<long long*>(out));
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
