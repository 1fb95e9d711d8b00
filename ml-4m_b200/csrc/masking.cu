// On-GPU input / target masking of image-like modalities: UnifiedMasking.image_mask of the reference
// (fourm/data/masking.py:236-266) runs per sample on CPU dataloader workers (torch.rand + argsort + two gathers + argmin); here one
// CTA handles one (sample, modality) row on the device, so only token ids and per-row budgets have to travel -- masks are born in HBM.
//
//   pi = argsort(noise)                                       (ascending; equal noise values keep index order)
//   input_mask[p]  = !(pi[p] <  n_in)                         (masking.py:250-252: gather of [0]*n_in + [1]*rest by pi)
//   target_mask[p] = !(n_in <= pi[p] < n_in + n_tgt)          (:254-259; n_tgt < 0 means "None": target_mask = ~input_mask)
//   decoder_attention_mask = 0 except at the FIRST target position, which holds the number of targets (:261-264)
// Integer / bool outputs: bit-exact with the reference for the same noise.
#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

constexpr int kMaskMaxL = 1024;

__global__ void __launch_bounds__(256)
image_mask_kernel(const float* __restrict__ noise, const int32_t* __restrict__ in_budget, const int32_t* __restrict__ tgt_budget,
                  uint8_t* __restrict__ input_mask, uint8_t* __restrict__ target_mask, int32_t* __restrict__ dam, int L) {
    pdl_enter();
    __shared__ float nz[kMaskMaxL];
    __shared__ int pi[kMaskMaxL];
    __shared__ int first_tgt, n_tgt_s;
    const long long row = blockIdx.x;
    const float* nr = noise + row * L;
    for (int i = threadIdx.x; i < L; i += 256) nz[i] = nr[i];
    if (threadIdx.x == 0) { first_tgt = L; n_tgt_s = 0; }
    __syncthreads();
    // rank of every element = its slot in the ascending order (ties by index); pi[rank] = index
    for (int i = threadIdx.x; i < L; i += 256) {
        const float v = nz[i];
        int rank = 0;
        for (int j = 0; j < L; ++j) {
            const float u = nz[j];
            rank += (u < v || (u == v && j < i)) ? 1 : 0;
        }
        pi[rank] = i;
    }
    __syncthreads();
    const int n_in = in_budget[row];
    const int n_tgt = tgt_budget ? tgt_budget[row] : -1;
    int my_cnt = 0, my_first = L;
    for (int p = threadIdx.x; p < L; p += 256) {
        const int idx = pi[p];
        const bool is_in = idx < n_in;
        const bool is_tgt = n_tgt < 0 ? !is_in : (idx >= n_in && idx < n_in + n_tgt);
        input_mask[row * L + p] = is_in ? 0 : 1;
        target_mask[row * L + p] = is_tgt ? 0 : 1;
        dam[row * L + p] = 0;
        if (is_tgt) { ++my_cnt; my_first = my_first < p ? my_first : p; }
    }
    atomicAdd(&n_tgt_s, my_cnt);
    atomicMin(&first_tgt, my_first);
    __syncthreads();
    // argmin(target_mask + arange * 1e-6) is position 0 when nothing is a target (masking.py:263 then writes 0 there)
    if (threadIdx.x == 0) dam[row * L + (first_tgt < L ? first_tgt : 0)] = n_tgt_s;
}

// uint8 RGB [B, C, H, W] -> bf16 patches [B * nh * nw, P * P * C] in '(ph pw c)' order with the loader's normalisation
// (x / 255 - mean[c]) / std[c] applied on the fly: the fp32 image (4 bytes / value over PCIe and through HBM) never exists.
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int Himg, int Wimg, int P,
                                   float m0, float m1, float m2, float s0, float s1, float s2) {
    pdl_enter();
    const int nh = Himg / P, nw = Wimg / P;
    const long long total = (long long)B * nh * nw * P * P * C;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(o % C);
        long long t = o / C;
        const int pw = (int)(t % P); t /= P;
        const int ph = (int)(t % P); t /= P;
        const int iw = (int)(t % nw); t /= nw;
        const int ih = (int)(t % nh);
        const int b = (int)(t / nh);
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        const float x = (float)img[(((long long)b * C + c) * Himg + ih * P + ph) * Wimg + iw * P + pw];
        out[o] = __float2bfloat16_rn((x / 255.0f - mean) / sd);       // ToTensor (x / 255) then Normalize: same operation order
    }
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_mask_images(const float* noise, const int32_t* in_budget, const int32_t* tgt_budget, uint8_t* input_mask,
                                  uint8_t* target_mask, int32_t* dam, long long rows, int L, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return 0;
    B200FM_CHECK(noise && in_budget && input_mask && target_mask && dam, "mask_images: null pointer");
    B200FM_CHECK(L >= 1 && L <= kMaskMaxL, "mask_images: L=%d outside [1, %d]", L, kMaskMaxL);
    B200FM_LAUNCH(image_mask_kernel, dim3((unsigned)rows), dim3(256), 0, stream, 1, noise, in_budget, tgt_budget, input_mask, target_mask, dam, L);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_patchify_u8(const uint8_t* img, void* out, int B, int C, int H, int W, int P, const float* mean3, const float* std3,
                                  void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return 0;
    B200FM_CHECK(img && out && mean3 && std3, "patchify_u8: null pointer (mean3 / std3 are HOST arrays of 3 floats)");
    B200FM_CHECK(C == 3, "patchify_u8: C=%d (RGB only)", C);
    B200FM_CHECK(P > 0 && H % P == 0 && W % P == 0, "Image sizes %dx%d must be divisible by patch sizes %dx%d", H, W, P, P);
    const long long total = (long long)B * C * H * W;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    B200FM_LAUNCH(patchify_u8_kernel, dim3(grid), dim3(256), 0, stream, 1, img, reinterpret_cast<__nv_bfloat16*>(out), B, C, H, W, P, mean3[0], mean3[1],
                  mean3[2], std3[0], std3[1], std3[2]);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
