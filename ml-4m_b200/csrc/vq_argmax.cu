// Fused VQ codebook scan: arg-max over the codebook of the cosine similarity (or negative squared L2 distance) without
// ever materialising the [n, K] score matrix the reference writes twice (fourm/vq/quantizers/quantize_lucid.py:402-407
// cosine, :275-284 Euclidean).  True fp32 FMA arithmetic (the reference runs a full-precision SGEMM: TF32 is off in
// save_vq_tokens.py / the trainers), ties -> lowest index like torch.argmax.
//
// Algorithmic HBM traffic is tiny (z: n*d*4 B in, idx: n*8 B out, codebook K*d*4 B stays L2-resident), so this kernel is
// bound by the fp32 FMA pipe: 2*n*K*d FLOP.  Layout: the codebook is pre-normalised and TRANSPOSED once per call into
// et[d][Kpad] so a 128-code chunk is 32 contiguous 512 B rows -> 16 B cp.async straight into smem (double buffered);
// each CTA owns 128 latents, each thread an 8x8 (latent x code) register tile.
#include <cfloat>

#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

constexpr int kVqTile = 128;     // latents per CTA and codes per chunk
constexpr int kVqThreads = 256;
constexpr int kVqDMax = 64;

// et[j][k] = (cosine ? e[k][j] / max(|e_k|, 1e-12) : e[k][j]);  ee[k] = |e_k|^2 (Euclidean only); padded codes: 0.
__global__ void vq_prep_codebook(const float* __restrict__ e, float* __restrict__ et, float* __restrict__ ee, int K, int Kpad,
                                 int d, int cosine) {
    pdl_enter();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kpad) return;
    if (k >= K) {
        for (int j = 0; j < d; ++j) et[(size_t)j * Kpad + k] = 0.f;
        ee[k] = 0.f;
        return;
    }
    float ss = 0.f;
    for (int j = 0; j < d; ++j) { const float v = e[(size_t)k * d + j]; ss += v * v; }
    const float inv = cosine ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;     // F.normalize: x / max(||x||, eps)
    for (int j = 0; j < d; ++j) et[(size_t)j * Kpad + k] = cosine ? e[(size_t)k * d + j] * inv : e[(size_t)k * d + j];
    ee[k] = ss;
}

B200FM_DEVINL void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
B200FM_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
B200FM_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D>
__global__ void __launch_bounds__(kVqThreads, 2)
vq_scan_kernel(const float* __restrict__ z, const float* __restrict__ et, const float* __restrict__ ee,
               const float* __restrict__ e_raw, int64_t* __restrict__ idx_out, float* __restrict__ quant_out, long long n,
               int K, int Kpad, int cosine) {
    pdl_enter();
    extern __shared__ __align__(16) float smem_f[];
    float* zs = smem_f;                                   // [D][128]  (latents transposed, normalised)
    float* es = zs + D * kVqTile;                         // 2 x [D][128]
    float* ees = es + 2 * D * kVqTile;                    // 2 x [128]
    float* zz_s = ees + 2 * kVqTile;                      // [128] |z|^2 (Euclidean)
    __shared__ float red_s[kVqTile][17];
    __shared__ int red_i[kVqTile][17];

    const int tid = threadIdx.x;
    const int tx = tid & 15;        // code group: codes tx*8 .. tx*8+7 of the chunk
    const int ty = tid >> 4;        // latent group: latents ty*8 .. ty*8+7 of the tile
    const long long row0 = (long long)blockIdx.x * kVqTile;
    const int num_chunks = Kpad / kVqTile;

    auto load_chunk = [&](int c, int buf) {
        float* dst = es + buf * D * kVqTile;
        const float* src = et + (size_t)c * kVqTile;
        for (int i = tid; i < D * (kVqTile / 4); i += kVqThreads) {
            const int j = i / (kVqTile / 4), q = i % (kVqTile / 4);
            cp_async16(dst + j * kVqTile + q * 4, src + (size_t)j * Kpad + q * 4);
        }
        if (tid < kVqTile / 4) cp_async16(ees + buf * kVqTile + tid * 4, ee + (size_t)c * kVqTile + tid * 4);
        cp_async_commit();
    };
    load_chunk(0, 0);

    // latent tile: one pass, two threads per latent (half a row each), normalise, store transposed
    {
        const int r = tid >> 1, half = tid & 1;
        const long long row = row0 + r;
        float v[D / 2];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < D / 2; j += 4) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < n) x = *reinterpret_cast<const float4*>(z + row * D + half * (D / 2) + j);
            v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
            ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        const float inv = cosine ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
#pragma unroll
        for (int j = 0; j < D / 2; ++j) zs[(half * (D / 2) + j) * kVqTile + r] = v[j] * inv;
        if (half == 0) zz_s[r] = ss;
    }

    float best[8];
    int besti[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { best[i] = -FLT_MAX; besti[i] = 0x7fffffff; }

    for (int c = 0; c < num_chunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < num_chunks) { load_chunk(c + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* eb = es + buf * D * kVqTile;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 8
        for (int kk = 0; kk < D; ++kk) {
            const float4 za = *reinterpret_cast<const float4*>(zs + kk * kVqTile + ty * 8);
            const float4 zb = *reinterpret_cast<const float4*>(zs + kk * kVqTile + ty * 8 + 4);
            const float4 ea = *reinterpret_cast<const float4*>(eb + kk * kVqTile + tx * 8);
            const float4 ec = *reinterpret_cast<const float4*>(eb + kk * kVqTile + tx * 8 + 4);
            const float zr[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
            const float er[8] = {ea.x, ea.y, ea.z, ea.w, ec.x, ec.y, ec.z, ec.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(zr[i], er[j], acc[i][j]);
        }
        const int code0 = c * kVqTile + tx * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int code = code0 + j;
            if (code < K) {
                const float eej = ees[buf * kVqTile + tx * 8 + j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    // Euclidean: -(|z|^2 - 2 z.e + |e|^2), evaluated in the reference's order (quantize_lucid.py:275-279)
                    const float s = cosine ? acc[i][j] : -((zz_s[ty * 8 + i] - 2.0f * acc[i][j]) + eej);
                    if (s > best[i]) { best[i] = s; besti[i] = code; }      // ascending codes + strict > : lowest index wins
                }
            }
        }
        __syncthreads();
    }

    // reduce the 16 code-group candidates of every latent: max score, ties -> lowest index
#pragma unroll
    for (int i = 0; i < 8; ++i) { red_s[ty * 8 + i][tx] = best[i]; red_i[ty * 8 + i][tx] = besti[i]; }
    __syncthreads();
    if (tid < kVqTile) {
        float b = red_s[tid][0];
        int bi = red_i[tid][0];
#pragma unroll
        for (int t = 1; t < 16; ++t) {
            const float s = red_s[tid][t];
            const int si = red_i[tid][t];
            if (s > b || (s == b && si < bi)) { b = s; bi = si; }
        }
        const long long row = row0 + tid;
        if (row < n) {
            idx_out[row] = bi;
            red_i[tid][16] = bi;
        }
    }
    if (quant_out != nullptr) {
        __syncthreads();
        for (int i = tid; i < kVqTile * (D / 4); i += kVqThreads) {
            const int r = i / (D / 4), q = i % (D / 4);
            const long long row = row0 + r;
            if (row < n)
                *reinterpret_cast<float4*>(quant_out + row * D + q * 4) =
                    *reinterpret_cast<const float4*>(e_raw + (size_t)red_i[r][16] * D + q * 4);
        }
    }
}

template <int D>
static int launch_vq(const float* z, const float* et, const float* ee, const float* e_raw, int64_t* idx, float* quant,
                     long long n, int K, int Kpad, int cosine, cudaStream_t stream) {
    const size_t smem = (size_t)(3 * D * kVqTile + 3 * kVqTile) * sizeof(float);
    auto kern = vq_scan_kernel<D>;
    static bool configured = false;
    if (!configured) {
        B200FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long long grid = (n + kVqTile - 1) / kVqTile;
    B200FM_LAUNCH(kern, dim3((unsigned)grid), dim3(kVqThreads), smem, stream, 1, z, et, ee, e_raw, idx, quant, n, K, Kpad, cosine);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_vq_argmax(const float* z, const float* codebook, int64_t* idx_out, float* quant_out, long long n, int K,
                                int d, int cosine, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    B200FM_CHECK(K > 0 && d > 0, "vq_argmax: empty codebook (K=%d d=%d)", K, d);
    B200FM_CHECK(d == 8 || d == 16 || d == 32 || d == 64, "vq_argmax: latent dim %d not in {8,16,32,64}", d);
    B200FM_CHECK(n >= 0 && n < (1ll << 31) * kVqTile, "vq_argmax: bad n %lld", n);
    if (n == 0) return 0;
    B200FM_CHECK(z && codebook && idx_out, "vq_argmax: null pointer");
    B200FM_CHECK((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(codebook) & 15) == 0,
                 "vq_argmax: z / codebook must be 16-byte aligned");
    const int Kpad = (K + kVqTile - 1) / kVqTile * kVqTile;
    float* ws = nullptr;
    B200FM_CUDA(cudaMallocAsync(&ws, (size_t)(d + 1) * Kpad * sizeof(float), stream));
    float* et = ws;
    float* ee = ws + (size_t)d * Kpad;
    B200FM_LAUNCH(vq_prep_codebook, dim3((Kpad + 255) / 256), dim3(256), 0, stream, 1, codebook, et, ee, K, Kpad, d, cosine);
    int rc = 0;
    switch (d) {
        case 8: rc = launch_vq<8>(z, et, ee, codebook, idx_out, quant_out, n, K, Kpad, cosine, stream); break;
        case 16: rc = launch_vq<16>(z, et, ee, codebook, idx_out, quant_out, n, K, Kpad, cosine, stream); break;
        case 32: rc = launch_vq<32>(z, et, ee, codebook, idx_out, quant_out, n, K, Kpad, cosine, stream); break;
        default: rc = launch_vq<64>(z, et, ee, codebook, idx_out, quant_out, n, K, Kpad, cosine, stream); break;
    }
    cudaFreeAsync(ws, stream);
    return rc;
}

extern "C" int b200fm_vq_argmax_host(const float* z_host, const float* codebook_dev, int64_t* idx_host, long long n, int K,
                                     int d, int cosine, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(z_host && idx_host, "vq_argmax_host: null host pointer");
    float* zd = nullptr;
    int64_t* id = nullptr;
    B200FM_CUDA(cudaMallocAsync(&zd, (size_t)n * d * sizeof(float), stream));
    B200FM_CUDA(cudaMallocAsync(&id, (size_t)n * sizeof(int64_t), stream));
    B200FM_CUDA(cudaMemcpyAsync(zd, z_host, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, stream));
    int rc = b200fm_vq_argmax(zd, codebook_dev, id, nullptr, n, K, d, cosine, stream_);
    if (rc == 0) {
        B200FM_CUDA(cudaMemcpyAsync(idx_host, id, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
        B200FM_CUDA(cudaStreamSynchronize(stream));
    }
    cudaFreeAsync(zd, stream);
    cudaFreeAsync(id, stream);
    return rc;
}
