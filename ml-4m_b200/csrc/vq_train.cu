// Training-side codebook maintenance of the 4M tokenizers (fourm/vq/quantizers/quantize_lucid.py:404-426 cosine, :286-299
// Euclidean).  The reference materialises a one-hot [n, K] and runs a second [d, n] x [n, K] GEMM for the per-code latent
// sums; here the assignment indices from the scan kernel drive a scatter: bins[idx] += 1, embed_sum[idx, :] += z (fp32
// atomics; n x d elements against K x d accumulators), then ONE pass over the codebook applies the EMA.  Between the two
// kernels the caller may all-reduce the packed [K * (d + 1)] statistics buffer (sync_codebook / DDP, :411, :419).
#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

// one warp per latent row; lane l owns components l, l + 32, ...  (d <= 256)
__global__ void __launch_bounds__(256)
vq_ema_stats_kernel(const float* __restrict__ z, const long long* __restrict__ idx, long long n, int K, int d, int cosine,
                    float* __restrict__ bins, float* __restrict__ embed_sum) {
    pdl_enter();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long r = static_cast<long long>(blockIdx.x) * 8 + warp; r < n; r += static_cast<long long>(gridDim.x) * 8) {
        const long long k = idx[r];
        if (k < 0 || k >= K) continue;
        float v[8];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            v[i] = c < d ? z[r * d + c] : 0.f;
            ss += v[i] * v[i];
        }
        float scale = 1.0f;
        if (cosine) {                                            // l2norm(flatten): x / max(|x|, 1e-12)  (F.normalize)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            scale = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < d) atomicAdd(embed_sum + k * d + c, v[i] * scale);
        }
        if (lane == 0) atomicAdd(bins + k, 1.0f);
    }
}

// cosine codebook EMA (quantize_lucid.py:413-425): one warp per code
__global__ void __launch_bounds__(256)
vq_ema_update_cosine_kernel(float* __restrict__ embed, float* __restrict__ cluster_size, const float* __restrict__ bins,
                            const float* __restrict__ embed_sum, int K, int d, float decay) {
    pdl_enter();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = blockIdx.x * 8 + warp; k < K; k += gridDim.x * 8) {
        const float b = bins[k];
        float e[8], s[8];
        float ne = 0.f, ns = 0.f;
        const float binv = 1.0f / (b == 0.f ? 1.0f : b);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            e[i] = c < d ? embed[static_cast<long long>(k) * d + c] : 0.f;
            s[i] = c < d ? embed_sum[static_cast<long long>(k) * d + c] * binv : 0.f;
            ne += e[i] * e[i];
            ns += s[i] * s[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ne += __shfl_xor_sync(0xffffffffu, ne, o);
            ns += __shfl_xor_sync(0xffffffffu, ns, o);
        }
        const float ie = 1.0f / fmaxf(sqrtf(ne), 1e-12f), is = 1.0f / fmaxf(sqrtf(ns), 1e-12f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < d) {
                const float target = (b == 0.f) ? e[i] * ie : s[i] * is;          // unused code keeps its (normalised) direction
                embed[static_cast<long long>(k) * d + c] = e[i] * decay + target * (1.0f - decay);
            }
        }
        if (lane == 0) cluster_size[k] = cluster_size[k] * decay + b * (1.0f - decay);
    }
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_vq_ema_stats(const float* z, const long long* idx, long long n, int K, int d, int cosine, float* bins,
                                   float* embed_sum, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(z && idx && bins && embed_sum, "vq_ema_stats: null pointer");
    B200FM_CHECK(K > 0 && d > 0 && d <= 256, "vq_ema_stats: K=%d d=%d (need 0 < d <= 256)", K, d);
    const long long blocks = (n + 7) / 8;
    const int grid = static_cast<int>(blocks < 148 * 8 ? blocks : 148 * 8);
    B200FM_LAUNCH(vq_ema_stats_kernel, dim3(grid), dim3(256), 0, stream, 1, z, idx, n, K, d, cosine, bins, embed_sum);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_vq_ema_update_cosine(float* embed, float* cluster_size, const float* bins, const float* embed_sum, int K, int d,
                                           float decay, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    B200FM_CHECK(embed && cluster_size && bins && embed_sum, "vq_ema_update_cosine: null pointer");
    B200FM_CHECK(K > 0 && d > 0 && d <= 256, "vq_ema_update_cosine: K=%d d=%d (need 0 < d <= 256)", K, d);
    const int blocks = (K + 7) / 8;
    B200FM_LAUNCH(vq_ema_update_cosine_kernel, dim3(blocks < 148 * 8 ? blocks : 148 * 8), dim3(256), 0, stream, 1, embed, cluster_size, bins,
                  embed_sum, K, d, decay);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
