// Attention forward for key counts beyond the resident range (Nk > 256): the generation callers of the same math
// (fourm/models/generate.py:407-445, 628-648, 745-764, 886-913 -- encoder contexts of up to ~1.9 k tokens, head_dim 64).
// Same contract as attention_fwd.cu (mask byte != 0 -> score := -3e38 in the log2 domain, fully masked row -> uniform,
// stats = (row max in the log2 domain, 1/sum)); the keys are streamed in 128-key tiles with the usual running-max
// rescale, so the result is the exact softmax up to fp32 rounding of the rescale factors.
//
// Work item = (batch b, head h, 128-query tile).  CTA = 6 warps: 0-3 softmax / running output (TMEM lane quarter = warp,
// 64 fp32 output accumulators per thread), 4 = TMA producer + TMEM allocator, 5 = MMA issuer.  K/V tiles are double
// buffered; S = Q K_t^T lands in TMEM columns [0,128), the per-tile product P_t V_t in columns [128,192).
#include <cfloat>

#include "../../include/b200fm.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "tmap.cuh"

namespace b200fm {

struct AttnLongSmem {
    static constexpr int kQ = 0;                         // 128 x 64 bf16
    static constexpr int kK = 16384;                     // 2 x (128 keys x 64)
    static constexpr int kV = kK + 2 * 16384;            // 2 x (128 keys x 64)
    static constexpr int kP = kV + 2 * 16384;            // 128 q x 128 keys bf16 = two 64-key swizzle atoms
    static constexpr int kBar = kP + 32768;
    static constexpr int kStage = kBar + 128;            // 4 warps x 2 KB store staging
    static constexpr int kTotal = kStage + 4 * 2048 + 1024;
    static constexpr int kTmemCols = 256;
    static constexpr int kOCol = 128;
};

struct AttnLongArgs {
    const uint8_t* mask;
    long long mask_b_stride, mask_q_stride;
    __nv_bfloat16* out;
    long long ldo;
    float* stats;
    int B, H, Nq, Nk, q_tiles, k_tiles, num_items;
    float scale_log2;
};

__global__ void __launch_bounds__(192, 1)
attention_fwd_long_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const AttnLongArgs args) {
    using SM = AttnLongSmem;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_free = bars + 1;
    uint64_t* kv_full = bars + 2;     // [2]
    uint64_t* kv_free = bars + 4;     // [2]
    uint64_t* s_full = bars + 6;
    uint64_t* p_full = bars + 7;
    uint64_t* o_full = bars + 8;
    uint64_t* o_free = bars + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
            mbar_init(q_full, 1); mbar_init(q_free, 1);
            for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_free[i], 1); }
            mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_free, 4);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, SM::kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0, g = 0;
            for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
                const int h = item % args.H;
                const int qt = (item / args.H) % args.q_tiles;
                const int b = item / (args.H * args.q_tiles);
                mbar_wait(q_free, (it & 1) ^ 1);
                mbar_arrive_expect_tx(q_full, 16384);
                tma_load_3d(smem + SM::kQ, &tmap_q, q_full, h * 64, qt * 128, b, kEvictFirst);
                for (int t = 0; t < args.k_tiles; ++t, ++g) {
                    const uint32_t buf = g & 1, par = (g >> 1) & 1;
                    mbar_wait(&kv_free[buf], par ^ 1);
                    mbar_arrive_expect_tx(&kv_full[buf], 32768);
                    tma_load_3d(smem + SM::kK + buf * 16384, &tmap_k, &kv_full[buf], h * 64, t * 128, b);
                    tma_load_3d(smem + SM::kV + buf * 16384, &tmap_v, &kv_full[buf], h * 64, t * 128, b);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t kIdescS = make_idesc_bf16(128, 128, false, false);
            constexpr uint32_t kIdescO = make_idesc_bf16(128, 64, false, true);
            const uint32_t sq = smem_u32(smem + SM::kQ), sp = smem_u32(smem + SM::kP);
            uint32_t it = 0, g = 0;
            for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
                mbar_wait(q_full, it & 1);
                for (int t = 0; t < args.k_tiles; ++t, ++g) {
                    const uint32_t buf = g & 1, par = (g >> 1) & 1;
                    const uint32_t sk = smem_u32(smem + SM::kK + buf * 16384), sv = smem_u32(smem + SM::kV + buf * 16384);
                    mbar_wait(&kv_full[buf], par);
                    tc_fence_after();
                    // S columns are free: the softmax warps finished reading the previous tile before p_full, which was
                    // awaited below before the previous P V was issued
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base, make_smem_desc(sq + k * 32, 16, 1024), make_smem_desc(sk + k * 32, 16, 1024), kIdescS, k != 0);
                    umma_commit(s_full);
                    if (t == args.k_tiles - 1) umma_commit(q_free);
                    mbar_wait(p_full, g & 1);
                    mbar_wait(o_free, (g & 1) ^ 1);
                    tc_fence_after();
                    const int keys = min(128, args.Nk - t * 128);
                    const int nk_steps = (keys + 15) / 16;
                    for (int kk = 0; kk < nk_steps; ++kk)
                        umma_bf16(tmem_base + SM::kOCol, make_smem_desc(sp + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                                  make_smem_desc(sv + kk * 2048, 16384, 1024), kIdescO, kk != 0);
                    umma_commit(o_full);
                    umma_commit(&kv_free[buf]);
                }
            }
        }
    } else {
        const int r = warp * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        uint8_t* sp = smem + SM::kP;
        uint32_t g = 0;
        for (int item = blockIdx.x; item < args.num_items; item += gridDim.x) {
            const int h = item % args.H;
            const int qt = (item / args.H) % args.q_tiles;
            const int b = item / (args.H * args.q_tiles);
            const int qrow = qt * 128 + r;
            const bool row_ok = qrow < args.Nq;
            const uint8_t* mrow = args.mask ? args.mask + b * args.mask_b_stride + (row_ok ? qrow : 0) * args.mask_q_stride : nullptr;
            float m_run = -INFINITY, l_run = 0.f;
            float o[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) o[j] = 0.f;

            for (int t = 0; t < args.k_tiles; ++t, ++g) {
                const int key0 = t * 128;
                uint32_t mb[4], tail[4];
                uint32_t any_masked = 0u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    mb[c] = attn_mask_bits32(mrow, key0 + c * 32, args.Nk);
                    const int valid = args.Nk - key0 - c * 32;
                    tail[c] = valid >= 32 ? 0u : (valid <= 0 ? 0xffffffffu : (0xffffffffu << valid));
                    any_masked |= mb[c];
                }
                mbar_wait(s_full, g & 1);
                tc_fence_after();
                float mu = -INFINITY;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (key0 + c * 32 < args.Nk) {
                        uint32_t rr[32];
                        tmem_ld_x32(t_lane + c * 32, rr);
                        tmem_ld_wait();
                        const uint32_t ex = mb[c] | tail[c];
                        if (!__any_sync(0xffffffffu, ex != 0u)) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) mu = fmaxf(mu, __uint_as_float(rr[j]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) mu = fmaxf(mu, ((ex >> j) & 1u) ? -INFINITY : __uint_as_float(rr[j]));
                        }
                    }
                }
                float m_t = mu * args.scale_log2;
                if (any_masked) m_t = fmaxf(m_t, kMaskedScore);
                const float m_new = fmaxf(m_run, m_t);                 // finite: every tile holds at least one real key
                const float alpha = fast_exp2(m_run - m_new);          // 0 on the first tile (m_run = -inf)
                const float p_masked = (m_new == kMaskedScore) ? 1.0f : 0.0f;
                const float neg_m = -m_new;
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
                    if (key0 + c * 32 < args.Nk) {
                        uint32_t rr[32];
                        tmem_ld_x32(t_lane + c * 32, rr);
                        tmem_ld_wait();
                        const uint32_t mbits = mb[c], tbits = tail[c];
                        if (!__any_sync(0xffffffffu, (mbits | tbits) != 0u)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float p0 = fast_exp2(fmaf(__uint_as_float(rr[j]), args.scale_log2, neg_m));
                                const float p1 = fast_exp2(fmaf(__uint_as_float(rr[j + 1]), args.scale_log2, neg_m));
                                sum += p0 + p1;
                                pk[j >> 1] = pack_bf16x2(p0, p1);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                float p0 = fast_exp2(fmaf(__uint_as_float(rr[j]), args.scale_log2, neg_m));
                                float p1 = fast_exp2(fmaf(__uint_as_float(rr[j + 1]), args.scale_log2, neg_m));
                                if ((mbits >> j) & 1u) p0 = p_masked;
                                if ((mbits >> (j + 1)) & 1u) p1 = p_masked;
                                if ((tbits >> j) & 1u) p0 = 0.f;
                                if ((tbits >> (j + 1)) & 1u) p1 = 0.f;
                                sum += p0 + p1;
                                pk[j >> 1] = pack_bf16x2(p0, p1);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = 0u;
                    }
                    uint8_t* atom = sp + (c >> 1) * 16384;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4*>(atom + swz128(r, (c & 1) * 4 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(p_full);
                l_run = fmaf(l_run, alpha, sum);
                m_run = m_new;

                mbar_wait(o_full, g & 1);
                tc_fence_after();
                {
                    uint32_t o0[32], o1[32];
                    tmem_ld_x32(t_lane + SM::kOCol, o0);
                    tmem_ld_x32(t_lane + SM::kOCol + 32, o1);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(o_free);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        o[j] = fmaf(o[j], alpha, __uint_as_float(o0[j]));
                        o[32 + j] = fmaf(o[32 + j], alpha, __uint_as_float(o1[j]));
                    }
                }
            }

            const float inv = 1.0f / l_run;
            if (row_ok && args.stats) {
                float2* st = reinterpret_cast<float2*>(args.stats) + ((static_cast<long long>(b) * args.H + h) * args.Nq + qrow);
                *st = make_float2(m_run, inv);
            }
            uint32_t* stg = reinterpret_cast<uint32_t*>(smem + SM::kStage) + warp * 512;
            __nv_bfloat16* base = args.out + static_cast<long long>(b) * args.Nq * args.ldo + h * 64;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(o[2 * j] * inv, o[2 * j + 1] * inv);
            attn_stage_store32(stg, lane, pk, base, args.ldo, qt * 128 + warp * 32, args.Nq);
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(o[32 + 2 * j] * inv, o[32 + 2 * j + 1] * inv);
            attn_stage_store32(stg, lane, pk, base + 32, args.ldo, qt * 128 + warp * 32, args.Nq);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SM::kTmemCols);
    }
}

// called by b200fm_attention_fwd (attention_fwd.cu) when Nk > 256
int launch_attention_fwd_long(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const uint8_t* mask,
                              long long mask_b_stride, long long mask_q_stride, __nv_bfloat16* out, long long ldo, float* stats, int B,
                              int H, int Nq, int Nk, float scale_log2, cudaStream_t stream) {
    AttnLongArgs a;
    a.mask = mask; a.mask_b_stride = mask_b_stride; a.mask_q_stride = mask_q_stride;
    a.out = out; a.ldo = ldo; a.stats = stats;
    a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.q_tiles = (Nq + 127) / 128; a.k_tiles = (Nk + 127) / 128;
    a.num_items = B * H * a.q_tiles;
    a.scale_log2 = scale_log2;
    constexpr int smem = AttnLongSmem::kTotal;
    static bool configured = false;
    if (!configured) {
        B200FM_CUDA(cudaFuncSetAttribute(attention_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int sms = usable_sm_count();
    const int grid = a.num_items < sms ? a.num_items : sms;
    B200FM_LAUNCH(attention_fwd_long_kernel, dim3(grid), dim3(192), smem, stream, 1, tq, tk, tv, a);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200fm
