// Small-M "GEMM" = weight streaming: y[M <= 8, N] = x[M, K] W[N, K]^T (+ bias), the shape of every Linear in the K/V-cached
// autoregressive decode loop (b200fm/decode.py: one new token per sequence; M = batch x CFG variants, 1..4 rows).  A 128 x BN
// tcgen05 tile wastes > 98 % of its MMA rows here and, worse, only N / BN CTAs (16 for N = 2048) pull weights from HBM: the 4M-XL
// decode step measured 0.7 TB/s through the tensor-core GEMM.  This kernel is HBM-bound by construction: every warp owns output
// columns, streams the corresponding weight rows with 16-byte loads (4 in flight per lane), keeps the few activation rows in
// shared memory, and reduces with shuffles.  Epilogues: bf16, fp32, fp32 residual add, SwiGLU (a = fc1 row, b = fc3 row, g = silu(a) * b with the
// tcgen05 epilogue's rounding points).  Dispatched from b200fm_gemm_bf16 (NT layout, M <= 8): callers do not see it.
#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

constexpr int kGemvMaxM = 8;
constexpr int kGemvWarps = 8;

B200FM_DEVINL float silu_gv(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

struct GemvArgs {
    const __nv_bfloat16* A; long long lda;
    const __nv_bfloat16* W; long long ldb;
    void* out0; long long ld0;
    void* out1; long long ld1;
    const float* bias;
    const float* resid; long long ldr;
    const float* alpha_dev;
    float alpha;
    int M, N, K, n_half;
    int prefetch;
};

B200FM_DEVINL uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int MM>
B200FM_DEVINL void dot8(const uint4& w, const uint4* arow, int chunk, int kchunks, float (&acc)[MM]) {
    const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
#pragma unroll
    for (int m = 0; m < MM; ++m) {
        const uint4 a = arow[m * kchunks + chunk];
        const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
        float s = acc[m];
        s = fmaf(w0.x, a0.x, s); s = fmaf(w0.y, a0.y, s); s = fmaf(w1.x, a1.x, s); s = fmaf(w1.y, a1.y, s);
        s = fmaf(w2.x, a2.x, s); s = fmaf(w2.y, a2.y, s); s = fmaf(w3.x, a3.x, s); s = fmaf(w3.y, a3.y, s);
        acc[m] = s;
    }
}

// Weight rows are streamed in batches of 8 x 16 B per lane (one 4 KB row of a K = 2048 layer = ONE batch): all loads of a batch are
// issued before the first FMA, so a warp keeps 4 KB (8 KB with the SwiGLU row pair) in flight instead of 2 KB, and the first batch is
// requested BEFORE the activation rows are staged in shared memory -- the kernel's critical path is one memory round trip.
constexpr int kGemvBatch = 8;

template <int NR>
B200FM_DEVINL void load_batch(const uint4* const (&wp)[NR], int c0, int kchunks, uint4 (&w)[NR][kGemvBatch]) {
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < kGemvBatch; ++j) {
            const int c = c0 + j * 32;
            w[r][j] = c < kchunks ? ldg_stream(wp[r] + c) : make_uint4(0u, 0u, 0u, 0u);
        }
}
template <int MM, int NR>
B200FM_DEVINL void fma_batch(const uint4 (&w)[NR][kGemvBatch], const uint4* arow, int c0, int kchunks, float (&acc)[NR][MM]) {
#pragma unroll
    for (int j = 0; j < kGemvBatch; ++j) {
        const int c = c0 + j * 32;
        if (c < kchunks) {
#pragma unroll
            for (int r = 0; r < NR; ++r) dot8<MM>(w[r][j], arow, c, kchunks, acc[r]);
        }
    }
}

template <int MM, int EPI>
__global__ void __launch_bounds__(kGemvWarps * 32)
gemv_kernel(const GemvArgs a) {
    constexpr int NR = EPI == B200FM_EPI_SWIGLU ? 2 : 1;                  // weight rows per output column
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The weights do not depend on the preceding kernels of a decode step (only the activation rows do): pull this warp's first rows
    // into L2 BEFORE waiting for the predecessor.  In a chain of programmatically dependent launches this kernel is resident while
    // the previous one still runs, so the HBM stream of layer i+1 overlaps the execution of layer i.  (A prefetch is a hint: if an
    // earlier kernel does rewrite the weights, L2 stays coherent.)
    if (a.prefetch) {
        const long long row_bytes = (long long)a.K * 2;
        int n = blockIdx.x * kGemvWarps + warp;
#pragma unroll 1
        for (int it = 0; it < 2 && n < a.N; ++it, n += gridDim.x * kGemvWarps) {
            const char* r0 = reinterpret_cast<const char*>(a.W + (long long)n * a.ldb);
            for (long long off = lane * 128ll; off < row_bytes; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + off));
            if constexpr (NR == 2) {
                const char* r1 = reinterpret_cast<const char*>(a.W + (long long)(a.n_half + n) * a.ldb);
                for (long long off = lane * 128ll; off < row_bytes; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + off));
            }
        }
    }
    pdl_wait();
    extern __shared__ uint4 smem_a[];                    // [MM][K / 8] chunks of 8 bf16
    const int kchunks = a.K / 8;
    int n = blockIdx.x * kGemvWarps + warp;
    uint4 w[NR][kGemvBatch];
    const uint4* wp[NR];
    if (n < a.N) {                                       // first batch of weight loads in flight across the activation staging
        wp[0] = reinterpret_cast<const uint4*>(a.W + (long long)n * a.ldb);
        if constexpr (NR == 2) wp[1] = reinterpret_cast<const uint4*>(a.W + (long long)(a.n_half + n) * a.ldb);
        load_batch<NR>(wp, lane, kchunks, w);
    }
    for (int i = threadIdx.x; i < MM * kchunks; i += kGemvWarps * 32) {
        const int m = i / kchunks, c = i % kchunks;
        smem_a[i] = m < a.M ? *reinterpret_cast<const uint4*>(a.A + (long long)m * a.lda + c * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const float alpha = a.alpha * (a.alpha_dev ? __ldg(a.alpha_dev) : 1.0f);
    for (; n < a.N; n += gridDim.x * kGemvWarps) {
        float acc[NR][MM];
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int m = 0; m < MM; ++m) acc[r][m] = 0.f;
        for (int c0 = lane; c0 < kchunks; c0 += 32 * kGemvBatch) {
            fma_batch<MM, NR>(w, smem_a, c0, kchunks, acc);
            // next batch: of this row, or the first batch of the warp's next row
            const int c1 = c0 + 32 * kGemvBatch;
            if (c1 < kchunks) {
                load_batch<NR>(wp, c1, kchunks, w);
            } else {
                const int n1 = n + gridDim.x * kGemvWarps;
                if (n1 < a.N) {
                    wp[0] = reinterpret_cast<const uint4*>(a.W + (long long)n1 * a.ldb);
                    if constexpr (NR == 2) wp[1] = reinterpret_cast<const uint4*>(a.W + (long long)(a.n_half + n1) * a.ldb);
                    load_batch<NR>(wp, lane, kchunks, w);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int m = 0; m < MM; ++m) acc[r][m] = warp_sum(acc[r][m]);
        if constexpr (EPI == B200FM_EPI_SWIGLU) {
            if (lane < a.M) {
                float va = 0.f, vb = 0.f;
#pragma unroll
                for (int m = 0; m < MM; ++m) if (m == lane) { va = acc[0][m]; vb = acc[NR - 1][m]; }
                if (a.bias) { va += a.bias[n]; vb += a.bias[a.n_half + n]; }
                __nv_bfloat16* ab = reinterpret_cast<__nv_bfloat16*>(a.out0) + (long long)lane * a.ld0;
                const __nv_bfloat16 pa = __float2bfloat16_rn(va), pb = __float2bfloat16_rn(vb);
                ab[n] = pa; ab[a.n_half + n] = pb;
                // same rounding points as the tcgen05 SwiGLU epilogue (gemm.cu): silu and the product are each rounded to bf16
                const float g = bf16_round(silu_gv(__bfloat162float(pa))) * __bfloat162float(pb);
                reinterpret_cast<__nv_bfloat16*>(a.out1)[(long long)lane * a.ld1 + n] = __float2bfloat16_rn(g);
            }
        } else {
            if (lane < a.M) {
                float v = 0.f;
#pragma unroll
                for (int m = 0; m < MM; ++m) if (m == lane) v = acc[0][m];
                if (a.bias) v += a.bias[n];
                if constexpr (EPI == B200FM_EPI_RESID) {          // out = resid + bf16(acc + bias), like the tile kernel's epilogue
                    reinterpret_cast<float*>(a.out0)[(long long)lane * a.ld0 + n] = a.resid[(long long)lane * a.ldr + n] + bf16_round(v);
                } else {
                    v *= alpha;
                    if constexpr (EPI == B200FM_EPI_F32) reinterpret_cast<float*>(a.out0)[(long long)lane * a.ld0 + n] = v;
                    else reinterpret_cast<__nv_bfloat16*>(a.out0)[(long long)lane * a.ld0 + n] = __float2bfloat16_rn(v);
                }
            }
        }
    }
}

template <int MM, int EPI>
static int launch_gemv_t(const GemvArgs& a, cudaStream_t stream) {
    auto kern = gemv_kernel<MM, EPI>;
    const size_t smem = (size_t)MM * (a.K / 8) * sizeof(uint4);
    static size_t configured = 0;
    if (smem > configured && smem > 48 * 1024) {
        B200FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int want = (a.N + kGemvWarps - 1) / kGemvWarps;
    const int cap = usable_sm_count() * (EPI == B200FM_EPI_SWIGLU ? 2 : 3);   // resident CTAs per SM at 124 / 85 registers: one wave, rows pipelined inside the warps
    B200FM_LAUNCH(kern, dim3(want < cap ? want : cap), dim3(kGemvWarps * 32), smem, stream, 1, a);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

bool gemv_applicable(int layout, int epilogue, int M, int N, int K, long long lda, long long ldb) {
    if (layout != 0 || M > kGemvMaxM || M < 1) return false;
    if (epilogue != B200FM_EPI_BF16 && epilogue != B200FM_EPI_F32 && epilogue != B200FM_EPI_SWIGLU && epilogue != B200FM_EPI_RESID) return false;
    if ((K % 8) != 0 || (lda % 8) != 0 || (ldb % 8) != 0) return false;
    const int mm = M <= 2 ? 2 : (M <= 4 ? 4 : 8);
    return (size_t)mm * K * 2 <= 200 * 1024 && N >= 64;
}

int launch_gemv(int epilogue, int M, int N, int K, const void* A, long long lda, const void* W, long long ldb, void* out0, long long ld0,
                void* out1, long long ld1, const float* bias, const float* resid, long long ldr, float alpha, const float* alpha_dev,
                cudaStream_t stream) {
    GemvArgs a;
    a.resid = resid; a.ldr = ldr;
    a.A = reinterpret_cast<const __nv_bfloat16*>(A); a.lda = lda; a.W = reinterpret_cast<const __nv_bfloat16*>(W); a.ldb = ldb;
    a.out0 = out0; a.ld0 = ld0; a.out1 = out1; a.ld1 = ld1; a.bias = bias; a.alpha = alpha; a.alpha_dev = alpha_dev;
    a.M = M; a.N = N; a.K = K; a.n_half = N; a.prefetch = option(kOptGemvPrefetch);
#define B200FM_GEMV_CASE(MM_)                                                                                     \
    if (epilogue == B200FM_EPI_BF16) return launch_gemv_t<MM_, B200FM_EPI_BF16>(a, stream);                       \
    if (epilogue == B200FM_EPI_F32) return launch_gemv_t<MM_, B200FM_EPI_F32>(a, stream);                         \
    if (epilogue == B200FM_EPI_RESID) return launch_gemv_t<MM_, B200FM_EPI_RESID>(a, stream);                     \
    return launch_gemv_t<MM_, B200FM_EPI_SWIGLU>(a, stream);
    if (M <= 2) { B200FM_GEMV_CASE(2) }
    if (M <= 4) { B200FM_GEMV_CASE(4) }
    B200FM_GEMV_CASE(8)
#undef B200FM_GEMV_CASE
}

}  // namespace b200fm
