// HBM-bound element-wise and row-wise kernels around the GEMMs: gated-MLP / GELU backward, the masked-token head's
// cross-entropy (forward + gradient in one pass), bias-gradient column sums, fp32->bf16 weight shadow casts, pixel
// patchify, fused AdamW.  All use 16-byte vector accesses; grids are sized in multiples of the SM count.
#include <cfloat>

#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

static int sm_count_ew() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
static int ew_grid(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)sm_count_ew() * 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ---- SwiGLU backward (fm_utils.py:143: fc2(silu(fc1 x) * fc3 x)) ------------------------------------------------
// ab bf16 [R, 2H] = [a | b] (saved pre-activations), dg bf16 [R, H] -> dab bf16 [R, 2H] = [da | db]
//   s = sigmoid(a); da = dg * b * s * (1 + a * (1 - s)); db = dg * a * s
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ ab, const __nv_bfloat16* __restrict__ dg,
                                  __nv_bfloat16* __restrict__ dab, long long R, int H, long long ld_ab, long long ld_dg,
                                  long long ld_dab) {
    pdl_enter();
    const int hv = H / 8;
    const long long total = R * hv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / hv;
        const int c = (int)(i % hv) * 8;
        const uint4 av = *reinterpret_cast<const uint4*>(ab + row * ld_ab + c);
        const uint4 bv = *reinterpret_cast<const uint4*>(ab + row * ld_ab + H + c);
        const uint4 gv = *reinterpret_cast<const uint4*>(dg + row * ld_dg + c);
        const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t da[4], db[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = unpack_bf16x2(aw[e]), b = unpack_bf16x2(bw[e]), g = unpack_bf16x2(gw[e]);
            float dax, dbx, day, dby;
            swiglu_grad(a.x, b.x, g.x, dax, dbx);
            swiglu_grad(a.y, b.y, g.y, day, dby);
            da[e] = pack_bf16x2(dax, day);
            db[e] = pack_bf16x2(dbx, dby);
        }
        *reinterpret_cast<uint4*>(dab + row * ld_dab + c) = make_uint4(da[0], da[1], da[2], da[3]);
        *reinterpret_cast<uint4*>(dab + row * ld_dab + H + c) = make_uint4(db[0], db[1], db[2], db[3]);
    }
}

// ---- GELU / tanh backward: dpre = dact * f'(pre), bf16 [R, N] -----------------------------------------------------
template <int ACT>   // 0 = GELU (erf), 1 = tanh
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dact,
                               __nv_bfloat16* __restrict__ dpre, long long n8) {
    pdl_enter();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 pv = reinterpret_cast<const uint4*>(pre)[i], gv = reinterpret_cast<const uint4*>(dact)[i];
        const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16x2(pw[e]), g = unpack_bf16x2(gw[e]);
            float d0, d1;
            if (ACT == 0) {
                d0 = 0.5f * (1.0f + erff(x.x * 0.70710678f)) + x.x * 0.3989422804f * __expf(-0.5f * x.x * x.x);
                d1 = 0.5f * (1.0f + erff(x.y * 0.70710678f)) + x.y * 0.3989422804f * __expf(-0.5f * x.y * x.y);
            } else {
                const float t0 = tanhf(x.x), t1 = tanhf(x.y);
                d0 = 1.0f - t0 * t0; d1 = 1.0f - t1 * t1;
            }
            o[e] = pack_bf16x2(g.x * d0, g.y * d1);
        }
        reinterpret_cast<uint4*>(dpre)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---- cross-entropy over fp32 logits (fm.py:597 F.cross_entropy, fp32 under autocast) --------------------------------
// One CTA per row.  loss_row[i] = logsumexp(l_i) - l_i[target_i];  dlogits bf16 [n, V] = softmax(l_i) - onehot(target_i)
// (unscaled: the 1/n mean factor and the upstream gradient are folded into the following GEMMs via alpha_dev).
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const float* __restrict__ logits, long long ld, const int64_t* __restrict__ targets,
                     float* __restrict__ loss_rows, __nv_bfloat16* __restrict__ dlogits, long long ldd, int V, const int* __restrict__ n_dev) {
    pdl_enter();
    __shared__ float red[8];
    __shared__ float bc;
    const long long row = blockIdx.x;
    if (n_dev != nullptr) {
        // device-side row count (graph-captured head): rows >= n carry no sample.  Their loss is 0 and, up to the next multiple
        // of 64 (the K block of the weight-gradient GEMM that contracts over rows), their gradient rows are ZERO.
        const long long n = max(0, __ldg(n_dev));
        if (row >= n) {
            if (row < ((n + 63) & ~63ll)) {
                if (threadIdx.x == 0) loss_rows[row] = 0.f;
                if (dlogits != nullptr)
                    for (int j = threadIdx.x * 2; j < V; j += 512) {
                        if (j + 1 < V) *reinterpret_cast<uint32_t*>(dlogits + row * ldd + j) = 0u; else dlogits[row * ldd + j] = __float2bfloat16_rn(0.f);
                    }
            } else if (threadIdx.x == 0) loss_rows[row] = 0.f;
            return;
        }
    }
    const float* l = logits + row * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int j = tid; j < V; j += 256) m = fmaxf(m, l[j]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (tid == 0) { float x = red[0]; for (int w = 1; w < 8; ++w) x = fmaxf(x, red[w]); bc = x; }
    __syncthreads();
    m = bc;
    float s = 0.f;
    for (int j = tid; j < V; j += 256) s += __expf(l[j] - m);
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { float x = 0.f; for (int w = 0; w < 8; ++w) x += red[w]; bc = x; }
    __syncthreads();
    s = bc;
    const int t = (int)targets[row];
    if (tid == 0) loss_rows[row] = (m + logf(s)) - l[t];
    if (dlogits != nullptr) {
        const float inv = 1.0f / s;
        __nv_bfloat16* d = dlogits + row * ldd;
        for (int j = tid * 2; j < V; j += 512) {
            const float p0 = __expf(l[j] - m) * inv - (j == t ? 1.0f : 0.0f);
            if (j + 1 < V) {
                const float p1 = __expf(l[j + 1] - m) * inv - (j + 1 == t ? 1.0f : 0.0f);
                *reinterpret_cast<uint32_t*>(d + j) = pack_bf16x2(p0, p1);
            } else {
                d[j] = __float2bfloat16_rn(p0);
            }
        }
    }
}

// ---- column sums (bias gradients): out[c] += sum_r x[r, c], x bf16 [R, N] ------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out, long long R, int N, int rows_per_block) {
    pdl_enter();
    const int c = blockIdx.x * 256 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    if (c >= N) return;
    float s = 0.f;
    const long long r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
    for (long long r = r0; r < r1; ++r) s += __bfloat162float(x[r * ld + c]);
    atomicAdd(out + c, s);
}

// ---- fp32 -> bf16 cast (weight shadows; also activations) ----------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
    pdl_enter();
    const long long n4 = n / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) y[n4 * 4 + threadIdx.x] = __float2bfloat16_rn(x[n4 * 4 + threadIdx.x]);
}

// ---- patchify (encoder_embeddings.py:301): img fp32 [B,C,H,W] -> bf16 [B*nh*nw, ph*pw*C] in '(ph pw c)' order ------
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int Himg, int Wimg, int P) {
    pdl_enter();
    const int nh = Himg / P, nw = Wimg / P;
    const long long total = (long long)B * nh * nw * P * P * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int c = (int)(t % C); t /= C;
        const int pw = (int)(t % P); t /= P;
        const int ph = (int)(t % P); t /= P;
        const int iw = (int)(t % nw); t /= nw;
        const int ih = (int)(t % nh); t /= nh;
        const int b = (int)t;
        out[i] = __float2bfloat16_rn(img[(((long long)b * C + c) * Himg + ih * P + ph) * Wimg + iw * P + pw]);
    }
}

// ---- fused AdamW (torch.optim.AdamW semantics, optim_factory.py:239-240), optional bf16 shadow of the new weight ------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             __nv_bfloat16* __restrict__ shadow, long long n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2_sqrt, float grad_scale) {
    pdl_enter();
    const float decay = 1.0f - lr * wd, step = lr / bc1, ob1 = 1.0f - beta1, ob2 = 1.0f - beta2;
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 g4 = reinterpret_cast<const float4*>(g)[i];
        float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        float* pp = &p4.x; float* mm = &m4.x; float* vv = &v4.x; const float* gg = &g4.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gi = gg[e] * grad_scale;
            mm[e] = beta1 * mm[e] + ob1 * gi;
            vv[e] = beta2 * vv[e] + ob2 * gi * gi;
            pp[e] = pp[e] * decay - step * (mm[e] / (sqrtf(vv[e]) / bc2_sqrt + eps));
        }
        reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
        if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = n4 * 4 + threadIdx.x;
        const float gi = g[i] * grad_scale;
        const float mi = beta1 * m[i] + ob1 * gi, vi = beta2 * v[i] + ob2 * gi * gi;
        const float pi = p[i] * decay - step * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        p[i] = pi; m[i] = mi; v[i] = vi;
        if (shadow) shadow[i] = __float2bfloat16_rn(pi);
    }
}

// ---- multi-tensor AdamW: one launch for every parameter tensor of a param group --------------------------------------
// table: one b200fm_adamw_tensor per tensor; chunk_tensor / chunk_offset map each CTA to (tensor, first element).
constexpr int kMtChunk = 8192;      // elements per CTA (256 threads x 8 float4)
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const b200fm_adamw_tensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                   const long long* __restrict__ chunk_offset, float lr, float beta1, float beta2, float eps, float wd, float bc1,
                   float bc2_sqrt, float grad_scale, const float* __restrict__ hyper_dev, float* __restrict__ gnorm_sq) {
    pdl_enter();
    float gsq = 0.f;                 // sum of squared (scaled) gradients this thread touched: the logged gradient norm for free
    if (hyper_dev != nullptr) {      // captured in a CUDA graph: the per-step scalars live in device memory {lr, 1 - b1^t, sqrt(1 - b2^t)}
        lr = __ldg(hyper_dev); bc1 = __ldg(hyper_dev + 1); bc2_sqrt = __ldg(hyper_dev + 2);
    }
    const b200fm_adamw_tensor t = table[chunk_tensor[blockIdx.x]];
    const long long off = chunk_offset[blockIdx.x];
    const long long end = off + kMtChunk < t.n ? off + kMtChunk : t.n;
    float* p = t.p; const float* g = t.g; float* m = t.m; float* v = t.v;
    __nv_bfloat16* shadow = reinterpret_cast<__nv_bfloat16*>(t.shadow_bf16);
    const float decay = 1.0f - lr * wd, step = lr / bc1, ob1 = 1.0f - beta1, ob2 = 1.0f - beta2;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                     (shadow == nullptr || (reinterpret_cast<uintptr_t>(shadow) & 7) == 0);
    if (vec) {
        const long long e4 = off + ((end - off) & ~3ll);
        for (long long i = off + threadIdx.x * 4ll; i < e4; i += 1024) {
            const float4 g4 = *reinterpret_cast<const float4*>(g + i);
            float4 p4 = *reinterpret_cast<float4*>(p + i), m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
            float* pp = &p4.x; float* mm = &m4.x; float* vv = &v4.x; const float* gg = &g4.x;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float gi = gg[e] * grad_scale;
                gsq = fmaf(gi, gi, gsq);
                mm[e] = beta1 * mm[e] + ob1 * gi;
                vv[e] = beta2 * vv[e] + ob2 * gi * gi;
                pp[e] = pp[e] * decay - step * (mm[e] / (sqrtf(vv[e]) / bc2_sqrt + eps));
            }
            *reinterpret_cast<float4*>(p + i) = p4; *reinterpret_cast<float4*>(m + i) = m4; *reinterpret_cast<float4*>(v + i) = v4;
            if (shadow) *reinterpret_cast<uint2*>(shadow + i) = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
        }
        for (long long i = e4 + threadIdx.x; i < end; i += 256) {
            const float gi = g[i] * grad_scale;
            gsq = fmaf(gi, gi, gsq);
            const float mi = beta1 * m[i] + ob1 * gi, vi = beta2 * v[i] + ob2 * gi * gi;
            const float pi = p[i] * decay - step * (mi / (sqrtf(vi) / bc2_sqrt + eps));
            p[i] = pi; m[i] = mi; v[i] = vi;
            if (shadow) shadow[i] = __float2bfloat16_rn(pi);
        }
    } else {
        for (long long i = off + threadIdx.x; i < end; i += 256) {
            const float gi = g[i] * grad_scale;
            gsq = fmaf(gi, gi, gsq);
            const float mi = beta1 * m[i] + ob1 * gi, vi = beta2 * v[i] + ob2 * gi * gi;
            const float pi = p[i] * decay - step * (mi / (sqrtf(vi) / bc2_sqrt + eps));
            p[i] = pi; m[i] = mi; v[i] = vi;
            if (shadow) shadow[i] = __float2bfloat16_rn(pi);
        }
    }
    if (gnorm_sq != nullptr) {       // one atomic per CTA (order-dependent in the last bits: a logged quantity)
        __shared__ float red[8];
        gsq = warp_sum(gsq);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gsq;
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(gnorm_sq, ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7])));
    }
}

// ---- multi-tensor exponential moving average of a model (ModelEmaV2.update, fourm/utils/timm/model_ema.py:123-127, called every
// step by run_training_vqvae.py:1169-1171): ema = decay * ema + (1 - decay) * model over every fp32 entry of the state_dict in ONE
// launch (the reference issues 3 element-wise kernels per tensor).  Same table / chunk maps as the AdamW kernel: p = EMA tensor,
// g = model tensor, shadow_bf16 = optional bf16 mirror of the EMA weight.  The two products and the sum are rounded separately
// (no fma contraction): bit-identical to the torch expression.  12 B per element (+2 with a mirror): HBM-bound.
__global__ void __launch_bounds__(256)
ema_multi_kernel(const b200fm_adamw_tensor* __restrict__ table, const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_offset,
                 float decay, float one_minus_decay) {
    pdl_enter();
    const b200fm_adamw_tensor t = table[chunk_tensor[blockIdx.x]];
    const long long off = chunk_offset[blockIdx.x];
    const long long end = off + kMtChunk < t.n ? off + kMtChunk : t.n;
    float* e = t.p; const float* m = t.g;
    __nv_bfloat16* shadow = reinterpret_cast<__nv_bfloat16*>(t.shadow_bf16);
    const bool vec = ((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(m)) & 15) == 0 && (shadow == nullptr || (reinterpret_cast<uintptr_t>(shadow) & 7) == 0);
    long long i = off + threadIdx.x * 4ll;
    const long long e4 = vec ? off + ((end - off) & ~3ll) : off;
    for (; i < e4; i += 1024) {
        float4 a = *reinterpret_cast<float4*>(e + i);
        const float4 b = *reinterpret_cast<const float4*>(m + i);
        a.x = __fadd_rn(__fmul_rn(decay, a.x), __fmul_rn(one_minus_decay, b.x));
        a.y = __fadd_rn(__fmul_rn(decay, a.y), __fmul_rn(one_minus_decay, b.y));
        a.z = __fadd_rn(__fmul_rn(decay, a.z), __fmul_rn(one_minus_decay, b.z));
        a.w = __fadd_rn(__fmul_rn(decay, a.w), __fmul_rn(one_minus_decay, b.w));
        *reinterpret_cast<float4*>(e + i) = a;
        if (shadow) *reinterpret_cast<uint2*>(shadow + i) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    }
    for (long long j = e4 + threadIdx.x; j < end; j += 256) {
        const float a = __fadd_rn(__fmul_rn(decay, e[j]), __fmul_rn(one_minus_decay, m[j]));
        e[j] = a;
        if (shadow) shadow[j] = __float2bfloat16_rn(a);
    }
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_ema_multi(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev, int n_chunks,
                                float decay, float one_minus_decay, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_chunks == 0) return 0;
    B200FM_CHECK(table_dev && chunk_tensor_dev && chunk_offset_dev, "ema_multi: null pointer");
    B200FM_LAUNCH(ema_multi_kernel, dim3(n_chunks), dim3(256), 0, stream, 1, table_dev, chunk_tensor_dev, chunk_offset_dev, decay, one_minus_decay);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_adamw_multi(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                                  int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                  float grad_scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_chunks == 0) return 0;
    B200FM_CHECK(table_dev && chunk_tensor_dev && chunk_offset_dev, "adamw_multi: null pointer");
    B200FM_CHECK(step >= 1, "adamw_multi: step must be >= 1");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
    B200FM_LAUNCH(adamw_multi_kernel, dim3(n_chunks), dim3(256), 0, stream, 1, table_dev, chunk_tensor_dev, chunk_offset_dev, lr, beta1, beta2, eps, weight_decay, bc1,
                                                   bc2s, grad_scale, static_cast<const float*>(nullptr), static_cast<float*>(nullptr));
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_adamw_multi_dev(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                                      int n_chunks, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                      const float* hyper_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_chunks == 0) return 0;
    B200FM_CHECK(table_dev && chunk_tensor_dev && chunk_offset_dev && hyper_dev, "adamw_multi_dev: null pointer");
    B200FM_LAUNCH(adamw_multi_kernel, dim3(n_chunks), dim3(256), 0, stream, 1, table_dev, chunk_tensor_dev, chunk_offset_dev, 0.f, beta1, beta2, eps, weight_decay, 1.f,
                                                   1.f, grad_scale, hyper_dev, static_cast<float*>(nullptr));
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

// As adamw_multi (hyper_dev == nullptr: lr / step given here) or adamw_multi_dev (hyper_dev != nullptr), and *gnorm_sq += sum of the squared
// (scaled) gradients of the group: the gradient norm the training loop logs (native_scaler.py:56-65) without a separate pass over the gradients.
extern "C" int b200fm_adamw_multi_gnorm(const b200fm_adamw_tensor* table_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                                        int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                        const float* hyper_dev, float* gnorm_sq, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_chunks == 0) return 0;
    B200FM_CHECK(table_dev && chunk_tensor_dev && chunk_offset_dev && gnorm_sq, "adamw_multi_gnorm: null pointer");
    B200FM_CHECK(hyper_dev != nullptr || step >= 1, "adamw_multi_gnorm: step must be >= 1");
    const float bc1 = hyper_dev ? 1.f : 1.0f - powf(beta1, (float)step);
    const float bc2s = hyper_dev ? 1.f : sqrtf(1.0f - powf(beta2, (float)step));
    B200FM_LAUNCH(adamw_multi_kernel, dim3(n_chunks), dim3(256), 0, stream, 1, table_dev, chunk_tensor_dev, chunk_offset_dev, lr, beta1, beta2, eps, weight_decay, bc1,
                                                   bc2s, grad_scale, hyper_dev, gnorm_sq);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_adamw_chunk_elems(void) { return kMtChunk; }


extern "C" int b200fm_swiglu_bwd(const void* ab, long long ld_ab, const void* dg, long long ld_dg, void* dab, long long ld_dab,
                                 long long R, int H, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (R == 0) return 0;
    B200FM_CHECK(ab && dg && dab, "swiglu_bwd: null pointer");
    B200FM_CHECK(H % 8 == 0 && ld_ab % 8 == 0 && ld_dg % 8 == 0 && ld_dab % 8 == 0, "swiglu_bwd: H and strides must be multiples of 8");
    B200FM_LAUNCH(swiglu_bwd_kernel, dim3(ew_grid(R * (H / 8), 256)), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(ab), reinterpret_cast<const __nv_bfloat16*>(dg),
                                                                  reinterpret_cast<__nv_bfloat16*>(dab), R, H, ld_ab, ld_dg, ld_dab);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_act_bwd(int act, const void* pre, const void* dact, void* dpre, long long n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(pre && dact && dpre, "act_bwd: null pointer");
    B200FM_CHECK(n % 8 == 0, "act_bwd: element count must be a multiple of 8");
    B200FM_CHECK(act == 0 || act == 1, "act_bwd: act must be 0 (gelu) or 1 (tanh)");
    const int grid = ew_grid(n / 8, 256);
    if (act == 0) B200FM_LAUNCH((act_bwd_kernel<0>), dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(pre), reinterpret_cast<const __nv_bfloat16*>(dact), reinterpret_cast<__nv_bfloat16*>(dpre), n / 8);
    else B200FM_LAUNCH((act_bwd_kernel<1>), dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(pre), reinterpret_cast<const __nv_bfloat16*>(dact), reinterpret_cast<__nv_bfloat16*>(dpre), n / 8);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_cross_entropy(const float* logits, long long ld, const int64_t* targets, float* loss_rows, void* dlogits,
                                    long long ldd, long long n, int V, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(logits && targets && loss_rows, "cross_entropy: null pointer");
    B200FM_CHECK(V > 0 && (dlogits == nullptr || ldd % 2 == 0), "cross_entropy: bad V / dlogits stride");
    B200FM_LAUNCH(cross_entropy_kernel, dim3((unsigned)n), dim3(256), 0, stream, 1, logits, ld, targets, loss_rows, reinterpret_cast<__nv_bfloat16*>(dlogits), ldd, V,
                  static_cast<const int*>(nullptr));
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_cross_entropy_dyn(const float* logits, long long ld, const int64_t* targets, float* loss_rows, void* dlogits,
                                        long long ldd, long long n_max, int V, const int* n_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_max == 0) return 0;
    B200FM_CHECK(logits && targets && loss_rows && n_dev, "cross_entropy_dyn: null pointer");
    B200FM_CHECK(V > 0 && (dlogits == nullptr || ldd % 2 == 0), "cross_entropy_dyn: bad V / dlogits stride");
    B200FM_LAUNCH(cross_entropy_kernel, dim3((unsigned)n_max), dim3(256), 0, stream, 1, logits, ld, targets, loss_rows, reinterpret_cast<__nv_bfloat16*>(dlogits), ldd, V, n_dev);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

// mean of the first *n_dev entries of x (0 when n == 0, like the reference's `torch.zeros(1)` term for an empty modality, fm.py:593-595);
// inv_n_out (optional) = 1 / max(n, 1): the factor the backward GEMMs fold in through alpha_dev.  Single CTA, deterministic order.
__global__ void __launch_bounds__(1024)
masked_mean_kernel(const float* __restrict__ x, const int* __restrict__ n_dev, long long n_max, float* __restrict__ mean_out, float* __restrict__ inv_n_out) {
    pdl_enter();
    __shared__ float red[32];
    const long long n = min((long long)max(0, __ldg(n_dev)), n_max);
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += 1024) s += x[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 32; ++w) t += red[w];
        const float inv = 1.0f / (float)(n > 0 ? n : 1);
        *mean_out = t * inv;
        if (inv_n_out) *inv_n_out = inv;
    }
}

extern "C" int b200fm_masked_mean(const float* x, const int* n_dev, long long n_max, float* mean_out, float* inv_n_out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    B200FM_CHECK(x && n_dev && mean_out, "masked_mean: null pointer");
    B200FM_LAUNCH(masked_mean_kernel, dim3(1), dim3(1024), 0, stream, 1, x, n_dev, n_max, mean_out, inv_n_out);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_colsum_bf16(const void* x, long long ld, float* out, long long R, int N, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (R == 0 || N == 0) return 0;
    B200FM_CHECK(x && out, "colsum: null pointer");
    const int rpb = 256;
    dim3 grid((N + 255) / 256, (unsigned)((R + rpb - 1) / rpb));
    B200FM_LAUNCH(colsum_bf16_kernel, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(x), ld, out, R, N, rpb);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_cast_f32_bf16(const float* x, void* y, long long n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(x && y, "cast: null pointer");
    B200FM_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0, "cast: misaligned buffers");
    B200FM_LAUNCH(cast_f32_bf16_kernel, dim3(ew_grid(n / 4 + 1, 256)), dim3(256), 0, stream, 1, x, reinterpret_cast<__nv_bfloat16*>(y), n);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_patchify(const float* img, void* out, int B, int C, int H, int W, int P, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return 0;
    B200FM_CHECK(img && out, "patchify: null pointer");
    B200FM_CHECK(P > 0 && H % P == 0 && W % P == 0, "Image sizes %dx%d must be divisible by patch sizes %dx%d", H, W, P, P);
    const long long total = (long long)B * C * H * W;
    B200FM_LAUNCH(patchify_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, stream, 1, img, reinterpret_cast<__nv_bfloat16*>(out), B, C, H, W, P);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float beta1,
                            float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(p && g && m && v, "adamw: null pointer");
    B200FM_CHECK(step >= 1, "adamw: step must be >= 1");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
    B200FM_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                 (shadow_bf16 == nullptr || (reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0), "adamw: buffers must be 16-byte aligned");
    B200FM_LAUNCH(adamw_kernel, dim3(ew_grid(n / 4 + 1, 256)), dim3(256), 0, stream, 1, p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n, lr, beta1, beta2, eps,
                                                    weight_decay, bc1, bc2s, grad_scale);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
