// LayerNorm forward / backward over the fp32 residual stream (fourm/models/fm_utils.py:93-108 -> F.layer_norm; fp32
// statistics under autocast).  HBM-bound: one warp per row, 16 B vector loads, the bf16 copy that the next
// GEMM consumes is produced here so the normalised activations never round-trip in fp32.
//   fwd bytes/row: read D*4, write D*2 (+8)          bwd bytes/row: read D*(2+4) (+D*4 residual grad), write D*4 (+D*2)
#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

constexpr int kLnWarps = 8;

// VEC = D / 128 float4 per lane
template <int VEC, bool OUT_BF16>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ add, float* __restrict__ x_out,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     void* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, float eps) {
    pdl_enter();
    constexpr int D = VEC * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 g[VEC], b[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        g[i] = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
        b[i] = beta ? reinterpret_cast<const float4*>(beta)[i * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
        float4 v[VEC];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            v[i] = xr[i * 32 + lane];
            if (add != nullptr) {      // pending residual branch: x := x + y (fp32 stream += bf16 branch output), written back
                const uint2 u = reinterpret_cast<const uint2*>(add + (size_t)row * D)[i * 32 + lane];
                const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
                v[i].x += a.x; v[i].y += a.y; v[i].z += c.x; v[i].w += c.y;
                reinterpret_cast<float4*>(x_out + (size_t)row * D)[i * 32 + lane] = v[i];
            }
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
            q += (a * a + c * c) + (d * d + e * e);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
        if (lane == 0) { if (mean_out) mean_out[row] = mean; if (rstd_out) rstd_out[row] = rstd; }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float o0 = (v[i].x - mean) * rstd * g[i].x + b[i].x;
            const float o1 = (v[i].y - mean) * rstd * g[i].y + b[i].y;
            const float o2 = (v[i].z - mean) * rstd * g[i].z + b[i].z;
            const float o3 = (v[i].w - mean) * rstd * g[i].w + b[i].w;
            if constexpr (OUT_BF16) {
                reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + (size_t)row * D)[i * 32 + lane] =
                    make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
            } else {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)row * D)[i * 32 + lane] = make_float4(o0, o1, o2, o3);
            }
        }
    }
}

template <int VEC, bool DY_BF16>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ dres,
                     float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int rows) {
    pdl_enter();
    constexpr int D = VEC * 128;
    __shared__ float red[kLnWarps][128 + 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 g[VEC], pg[VEC], pb[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        g[i] = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
        pg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
        float4 xh[VEC], dyv[VEC];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float4 xv = xr[i * 32 + lane];
            if constexpr (DY_BF16) {
                const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)row * D)[i * 32 + lane];
                const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
                dyv[i] = make_float4(a.x, a.y, c.x, c.y);
            } else {
                dyv[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + (size_t)row * D)[i * 32 + lane];
            }
            xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            const float4 gy = make_float4(dyv[i].x * g[i].x, dyv[i].y * g[i].y, dyv[i].z * g[i].z, dyv[i].w * g[i].w);
            c1 += (gy.x + gy.y) + (gy.z + gy.w);
            c2 += (gy.x * xh[i].x + gy.y * xh[i].y) + (gy.z * xh[i].z + gy.w * xh[i].w);
            pg[i].x += dyv[i].x * xh[i].x; pg[i].y += dyv[i].y * xh[i].y; pg[i].z += dyv[i].z * xh[i].z; pg[i].w += dyv[i].w * xh[i].w;
            pb[i].x += dyv[i].x; pb[i].y += dyv[i].y; pb[i].z += dyv[i].z; pb[i].w += dyv[i].w;
        }
        c1 = warp_sum(c1) * (1.0f / D);
        c2 = warp_sum(c2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float4 o;
            o.x = rstd * (dyv[i].x * g[i].x - c1 - xh[i].x * c2);
            o.y = rstd * (dyv[i].y * g[i].y - c1 - xh[i].y * c2);
            o.z = rstd * (dyv[i].z * g[i].z - c1 - xh[i].z * c2);
            o.w = rstd * (dyv[i].w * g[i].w - c1 - xh[i].w * c2);
            if (dres) {
                const float4 r = reinterpret_cast<const float4*>(dres + (size_t)row * D)[i * 32 + lane];
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            reinterpret_cast<float4*>(dx_out + (size_t)row * D)[i * 32 + lane] = o;
            if (dx_bf16)
                reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[i * 32 + lane] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
    if (dgamma == nullptr && dbeta == nullptr) return;
    // block reduce the per-warp column partials (128 columns at a time), then one atomic per column per block
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? dgamma : dbeta;
        if (dst == nullptr) continue;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            __syncthreads();
            *reinterpret_cast<float4*>(&red[warp][lane * 4]) = pass == 0 ? pg[i] : pb[i];
            __syncthreads();
            if (threadIdx.x < 128) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kLnWarps; ++w) s += red[w][threadIdx.x];
                atomicAdd(dst + i * 128 + threadIdx.x, s);
            }
        }
    }
}

// Default backward (option "ln_bwd_v2" = 1; 0 selects the kernel above): the residual-gradient row is requested together with x and dy,
// BEFORE the two warp reductions, so one row costs one memory round trip instead of two.  Stand-alone at 16384 x 768: 35.4 us = 5.68 TB/s
// vs 47.4 us (tools/ln_bench.py); one CTA of 8 warps per SM either way (register-limited: the column partials live in registers).
template <int VEC, bool DY_BF16>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_v2_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ dres,
                     float* __restrict__ dx_out, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int rows) {
    pdl_enter();
    constexpr int D = VEC * 128;
    __shared__ float red[kLnWarps][128 + 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 g[VEC], pg[VEC], pb[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        g[i] = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
        pg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
        float4 xh[VEC], dyv[VEC], rv[VEC];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i)       // the residual gradient does not depend on the reductions: fetch it with the first phase
            rv[i] = dres ? __ldcs(reinterpret_cast<const float4*>(dres + (size_t)row * D) + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float4 xv = xr[i * 32 + lane];
            if constexpr (DY_BF16) {
                const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)row * D)[i * 32 + lane];
                const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
                dyv[i] = make_float4(a.x, a.y, c.x, c.y);
            } else {
                dyv[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + (size_t)row * D)[i * 32 + lane];
            }
            xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            const float4 gy = make_float4(dyv[i].x * g[i].x, dyv[i].y * g[i].y, dyv[i].z * g[i].z, dyv[i].w * g[i].w);
            c1 += (gy.x + gy.y) + (gy.z + gy.w);
            c2 += (gy.x * xh[i].x + gy.y * xh[i].y) + (gy.z * xh[i].z + gy.w * xh[i].w);
            pg[i].x += dyv[i].x * xh[i].x; pg[i].y += dyv[i].y * xh[i].y; pg[i].z += dyv[i].z * xh[i].z; pg[i].w += dyv[i].w * xh[i].w;
            pb[i].x += dyv[i].x; pb[i].y += dyv[i].y; pb[i].z += dyv[i].z; pb[i].w += dyv[i].w;
        }
        c1 = warp_sum(c1) * (1.0f / D);
        c2 = warp_sum(c2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float4 o;
            o.x = rstd * (dyv[i].x * g[i].x - c1 - xh[i].x * c2);
            o.y = rstd * (dyv[i].y * g[i].y - c1 - xh[i].y * c2);
            o.z = rstd * (dyv[i].z * g[i].z - c1 - xh[i].z * c2);
            o.w = rstd * (dyv[i].w * g[i].w - c1 - xh[i].w * c2);
            o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w;
            reinterpret_cast<float4*>(dx_out + (size_t)row * D)[i * 32 + lane] = o;
            if (dx_bf16)
                reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[i * 32 + lane] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
    if (dgamma == nullptr && dbeta == nullptr) return;
    // block reduce the per-warp column partials (128 columns at a time), then one atomic per column per block
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? dgamma : dbeta;
        if (dst == nullptr) continue;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            __syncthreads();
            *reinterpret_cast<float4*>(&red[warp][lane * 4]) = pass == 0 ? pg[i] : pb[i];
            __syncthreads();
            if (threadIdx.x < 128) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kLnWarps; ++w) s += red[w][threadIdx.x];
                atomicAdd(dst + i * 128 + threadIdx.x, s);
            }
        }
    }
}

// ---- per-head LayerNorm over head_dim = 64 (qk_norm presets: NormAttention / NormCrossAttention, fm_utils.py:244-245, 290-291)
// x bf16 [rows, ldx]: head h occupies columns [h*64, h*64+64) (x may be a column slice of a packed qkv buffer);
// y bf16 [rows, ldy] = bf16(LN_fp32(x) * gamma + beta) -- F.layer_norm under autocast computes in fp32, the following
// q @ k^T casts to bf16.  One warp per (row, head), two elements per lane, statistics kept for the backward.
__global__ void __launch_bounds__(256)
headnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                    __nv_bfloat16* __restrict__ y, long long ldy, float2* __restrict__ stats, long long rows, int H, float eps) {
    pdl_enter();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float g0 = gamma[2 * lane], g1 = gamma[2 * lane + 1];
    const float b0 = beta ? beta[2 * lane] : 0.f, b1 = beta ? beta[2 * lane + 1] : 0.f;
    const long long total = rows * H;
    for (long long i = static_cast<long long>(blockIdx.x) * 8 + warp; i < total; i += static_cast<long long>(gridDim.x) * 8) {
        const long long row = i / H;
        const int h = static_cast<int>(i % H);
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + row * ldx + h * 64 + 2 * lane));
        float s = v.x + v.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 64.0f);
        const float d0 = v.x - mean, d1 = v.y - mean;
        float q = d0 * d0 + d1 * d1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / 64.0f) + eps);
        *reinterpret_cast<uint32_t*>(y + row * ldy + h * 64 + 2 * lane) = pack_bf16x2(d0 * rstd * g0 + b0, d1 * rstd * g1 + b1);
        if (lane == 0 && stats) stats[i] = make_float2(mean, rstd);
    }
}

// dy bf16 [rows, lddy] -> dx bf16 [rows, lddx]; dgamma / dbeta fp32 [64] are ACCUMULATED into (may be NULL)
__global__ void __launch_bounds__(256)
headnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ x, long long ldx,
                    const float* __restrict__ gamma, const float2* __restrict__ stats, __nv_bfloat16* __restrict__ dx, long long lddx,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int H) {
    pdl_enter();
    __shared__ float red[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float g0 = gamma[2 * lane], g1 = gamma[2 * lane + 1];
    float ag0 = 0.f, ag1 = 0.f, ab0 = 0.f, ab1 = 0.f;
    const long long total = rows * H;
    for (long long i = static_cast<long long>(blockIdx.x) * 8 + warp; i < total; i += static_cast<long long>(gridDim.x) * 8) {
        const long long row = i / H;
        const int h = static_cast<int>(i % H);
        const float2 st = stats[i];
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + row * ldx + h * 64 + 2 * lane));
        const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + row * lddy + h * 64 + 2 * lane));
        const float xh0 = (v.x - st.x) * st.y, xh1 = (v.y - st.x) * st.y;
        const float e0 = d.x * g0, e1 = d.y * g1;
        float c1 = e0 + e1, c2 = e0 * xh0 + e1 * xh1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        }
        c1 *= (1.0f / 64.0f); c2 *= (1.0f / 64.0f);
        *reinterpret_cast<uint32_t*>(dx + row * lddx + h * 64 + 2 * lane) =
            pack_bf16x2(st.y * (e0 - c1 - xh0 * c2), st.y * (e1 - c1 - xh1 * c2));
        ag0 += d.x * xh0; ag1 += d.y * xh1; ab0 += d.x; ab1 += d.y;
    }
    red[warp][2 * lane] = ag0; red[warp][2 * lane + 1] = ag1;
    red[warp][64 + 2 * lane] = ab0; red[warp][64 + 2 * lane + 1] = ab1;
    __syncthreads();
    if (threadIdx.x < 128) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        if (threadIdx.x < 64) { if (dgamma) atomicAdd(dgamma + threadIdx.x, t); }
        else if (dbeta) atomicAdd(dbeta + threadIdx.x - 64, t);
    }
}


}  // namespace b200fm


using namespace b200fm;

static int ln_grid(int rows) {
    int blocks = (rows + kLnWarps - 1) / kLnWarps;
    const int cap = 148 * 4;
    return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

extern "C" int b200fm_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean,
                                    float* rstd, int rows, int D, float eps, void* stream_) {
    return b200fm_add_layernorm_fwd(x, nullptr, nullptr, gamma, beta, y, y_is_bf16, mean, rstd, rows, D, eps, stream_);
}

extern "C" int b200fm_add_layernorm_fwd(const float* x, const void* add_bf16, float* x_out, const float* gamma, const float* beta, void* y,
                                        int y_is_bf16, float* mean, float* rstd, int rows, int D, float eps, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return 0;
    B200FM_CHECK(x && gamma && y, "layernorm_fwd: null pointer");
    B200FM_CHECK((add_bf16 == nullptr) == (x_out == nullptr), "add_layernorm_fwd: add and x_out go together");
    const __nv_bfloat16* add = reinterpret_cast<const __nv_bfloat16*>(add_bf16);
    B200FM_CHECK(D % 128 == 0 && D >= 128 && D <= 2048, "layernorm_fwd: D=%d unsupported (need multiple of 128 in [128, 2048])", D);
    const int grid = ln_grid(rows);
#define LN_FWD(V)                                                                                                        \
    case V:                                                                                                              \
        if (y_is_bf16) B200FM_LAUNCH((layernorm_fwd_kernel<V, true>), dim3(grid), dim3(kLnWarps * 32), 0, stream, 1, x, add, x_out, gamma, beta, y, mean, rstd, rows, eps); \
        else B200FM_LAUNCH((layernorm_fwd_kernel<V, false>), dim3(grid), dim3(kLnWarps * 32), 0, stream, 1, x, add, x_out, gamma, beta, y, mean, rstd, rows, eps);          \
        break;
    switch (D / 128) {
        LN_FWD(1) LN_FWD(2) LN_FWD(3) LN_FWD(4) LN_FWD(5) LN_FWD(6) LN_FWD(8) LN_FWD(10) LN_FWD(12) LN_FWD(16)
        default: B200FM_CHECK(false, "layernorm_fwd: D=%d has no instantiation", D);
    }
#undef LN_FWD
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                                    const float* rstd, const float* dres, float* dx_out, void* dx_bf16, float* dgamma,
                                    float* dbeta, int rows, int D, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0) return 0;
    B200FM_CHECK(dy && x && gamma && mean && rstd && dx_out, "layernorm_bwd: null pointer");
    B200FM_CHECK(D % 128 == 0 && D >= 128 && D <= 2048, "layernorm_bwd: D=%d unsupported", D);
    int grid = ln_grid(rows);
    if (grid > 148 * 2) grid = 148 * 2;       // fewer blocks -> fewer column atomics
    const bool v2 = option(kOptLnBwdV2) != 0;     // default: all loads of a row hoisted ahead of the reductions
#define LN_BWD(V)                                                                                                          \
    case V:                                                                                                                \
        if (v2 && dy_is_bf16) B200FM_LAUNCH((layernorm_bwd_v2_kernel<V, true>), dim3(grid), dim3(kLnWarps * 32), 0, stream, 1, dy, x, gamma, mean, rstd, dres, dx_out, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, rows); \
        else if (dy_is_bf16) B200FM_LAUNCH((layernorm_bwd_kernel<V, true>), dim3(grid), dim3(kLnWarps * 32), 0, stream, 1, dy, x, gamma, mean, rstd, dres, dx_out, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, rows); \
        else B200FM_LAUNCH((layernorm_bwd_kernel<V, false>), dim3(grid), dim3(kLnWarps * 32), 0, stream, 1, dy, x, gamma, mean, rstd, dres, dx_out, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, rows);          \
        break;
    switch (D / 128) {
        LN_BWD(1) LN_BWD(2) LN_BWD(3) LN_BWD(4) LN_BWD(5) LN_BWD(6) LN_BWD(8) LN_BWD(10) LN_BWD(12) LN_BWD(16)
        default: B200FM_CHECK(false, "layernorm_bwd: D=%d has no instantiation", D);
    }
#undef LN_BWD
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_headnorm_fwd(const void* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy,
                                   float* stats, long long rows, int H, float eps, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0 || H == 0) return 0;
    B200FM_CHECK(x && gamma && y, "headnorm_fwd: null pointer");
    B200FM_CHECK(ldx % 2 == 0 && ldy % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 3) == 0,
                 "headnorm_fwd: rows must be 4-byte aligned (even strides)");
    const long long blocks = (rows * H + 7) / 8;
    const int grid = static_cast<int>(blocks < 148 * 8 ? blocks : 148 * 8);
    B200FM_LAUNCH(headnorm_fwd_kernel, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(x), ldx, gamma, beta,
                  reinterpret_cast<__nv_bfloat16*>(y), ldy, reinterpret_cast<float2*>(stats), rows, H, eps);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_headnorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* gamma, const float* stats,
                                   void* dx, long long lddx, float* dgamma, float* dbeta, long long rows, int H, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows == 0 || H == 0) return 0;
    B200FM_CHECK(dy && x && gamma && stats && dx, "headnorm_bwd: null pointer");
    B200FM_CHECK(lddy % 2 == 0 && ldx % 2 == 0 && lddx % 2 == 0, "headnorm_bwd: strides must be even");
    const long long blocks = (rows * H + 7) / 8;
    const int grid = static_cast<int>(blocks < 148 * 4 ? blocks : 148 * 4);
    B200FM_LAUNCH(headnorm_bwd_kernel, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(dy), lddy,
                  reinterpret_cast<const __nv_bfloat16*>(x), ldx, gamma, reinterpret_cast<const float2*>(stats),
                  reinterpret_cast<__nv_bfloat16*>(dx), lddx, dgamma, dbeta, rows, H);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
