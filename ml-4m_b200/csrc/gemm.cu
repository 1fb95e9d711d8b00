// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in
// TMEM, double buffered) -> fused epilogues.  This one kernel family serves every dense contraction outside the
// attention core of the 4M block stack (reference call sites: fourm/models/fm_utils.py:155-157,190-194 qkv/q/kv/proj,
// :129-144 GatedMlp, fourm/models/fm.py:154 decoder_proj_context, decoder_embeddings.py:141-152 to_logits,
// encoder_embeddings.py:301 patch proj; fourm/vq/models/vit_models.py Mlp/Attention linears) and their backward
// (dgrad / wgrad).
//
//   C[M,N] = A . B^T-ish, three operand layouts:
//     LAYOUT_NT : A [M,K] row-major (K-major), B [N,K] row-major (K-major)             forward  y = x W^T
//     LAYOUT_NN : A [M,K] row-major (K-major), B [K,N] row-major (N-major / "MN-major") dgrad    dx = dy W
//     LAYOUT_TN : A [K,M] row-major (M-major), B [K,N] row-major (N-major)             wgrad    dW = dy^T x
//
// CTA = 10 warps: warps 0-7 epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2; two warps per quarter keep the
// epilogue off the critical path), warp 8 TMA producer + TMEM allocator, warp 9 MMA issuer.
// Tile 128 x BN x 64, BN in {128, 256}; smem ring of (16 KB + BN*128 B) stages; accumulators 2 x BN TMEM columns.
#include <cstdlib>

#include "../../include/b200fm.h"
#include "common.cuh"
#include "tmap.cuh"

namespace b200fm {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 320;     // warps 0-7 epilogue (lane quarter = warp & 3, column half = warp >> 2), 8 = TMA, 9 = MMA
constexpr int kEpiWarps = 8;

enum GemmLayout { LAYOUT_NT = 0, LAYOUT_NN = 1, LAYOUT_TN = 2 };

struct GemmArgs {
    int M, N, K;
    int num_m_blocks, num_n_blocks;
    void* out0;          // EPI_BF16/GELU(pre)/SWIGLU(ab): bf16 ; EPI_F32/EPI_RESID: fp32
    long long ld0;
    void* out1;          // EPI_SWIGLU: g bf16 [M, H] ; EPI_GELU: act bf16 [M,N]
    long long ld1;
    const float* bias;   // [N] fp32 or nullptr
    const float* resid;  // EPI_RESID: fp32 [M,N]
    long long ldr;
    int n_half;          // EPI_SWIGLU: H (B rows [0,H) = fc1, [H,2H) = fc3); N must equal H
    float alpha;         // EPI_F32 / EPI_BF16: out = alpha * (acc + bias)
    const float* alpha_dev;   // optional device scalar folded into alpha
    int act;             // EPI_GELU family: 0 = GELU(erf), 1 = tanh
    int k_splits;        // EPI_F32 only: > 1 -> each work item covers a K slice and accumulates with fp32 atomics (out pre-zeroed)
    int kb_per_split;
};

// CTA2: a CTA pair (cluster of 2, one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2 -- each CTA stages its own
// 128 rows of A and only HALF of the B tile (BN/2 rows), so the per-SM smem fill per MMA-cycle drops by a third.
template <int BN, bool CTA2>
struct GemmSmem {
    static constexpr int kBBytes = CTA2 ? BN * 64 : BN * 128;
    static constexpr int kStageBytes = kBM * 128 + kBBytes;
    static constexpr int kStages = CTA2 ? (BN == 256 ? 6 : 8) : ((BN == 256) ? 4 : 6);
    static constexpr int kBarBytes = 256;
    static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;   // +1024 alignment slack
};

B200FM_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
B200FM_DEVINL float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

template <int BN, int LAYOUT, int EPI, bool CTA2>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmArgs args) {
    using SM = GemmSmem<BN, CTA2>;
    constexpr int kStages = SM::kStages;
    constexpr bool A_MN = (LAYOUT == LAYOUT_TN);
    constexpr bool B_MN = (LAYOUT != LAYOUT_NT);
    constexpr int kABytes = kBM * 128;
    constexpr int kBBytes = SM::kBBytes;
    constexpr int kBNLocal = CTA2 ? BN / 2 : BN;              // B rows (N) staged by this CTA
    constexpr uint32_t kIdesc = make_idesc_bf16(CTA2 ? 2 * kBM : kBM, BN, A_MN, B_MN);
    constexpr int kTmemCols = 2 * BN;
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;      // 0 = leader (issues the MMAs), 1 = peer
    const int cta_stride = CTA2 ? 2 : 1;
    const int my_first = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int my_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * SM::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m_units = (args.num_m_blocks + cta_stride - 1) / cta_stride;      // CTA2: pairs of 128-row blocks
    const int num_mn = num_m_units * args.num_n_blocks;
    const int num_tiles = num_mn * args.k_splits;               // work items: (k slice, n block, m unit), m fastest
    const int num_kb_total = (args.K + kBK - 1) / kBK;

    if (warp == kEpiWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_a);
            tma_prefetch_desc(&tmap_b);
            for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kEpiWarps * cta_stride); }
            fence_mbar_init();
        }
        __syncwarp();
        if constexpr (CTA2) tmem_alloc_2sm(tmem_slot, kTmemCols); else tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all(); else __syncthreads();      // peer barriers must be initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kEpiWarps) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = my_first; tile < num_tiles; tile += my_step) {
                const int mn = tile % num_mn, ks = tile / num_mn;
                const int m_blk = (mn % num_m_units) * cta_stride + static_cast<int>(rank);
                const int n_blk = mn / num_m_units;
                const int m0 = m_blk * kBM;
                const int n0 = (EPI == B200FM_EPI_SWIGLU) ? n_blk * (BN / 2) : n_blk * BN;
                const int kb_begin = ks * args.kb_per_split;
                const int kb_end = min(kb_begin + args.kb_per_split, num_kb_total);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * SM::kStageBytes;
                    uint8_t* sb = sa + kABytes;
                    const int k0 = kb * kBK;
                    if constexpr (CTA2) {
                        // both CTAs' bytes are counted on the LEADER's full barrier (the leader issues the MMA for the pair)
                        const uint32_t fb = mapa_u32(&full_bar[stage], 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + kBBytes));
                        if constexpr (!A_MN) {
                            tma_load_2d_2sm(sa, &tmap_a, fb, k0, m0);
                        } else {
#pragma unroll
                            for (int c = 0; c < kBM / 64; ++c) tma_load_2d_2sm(sa + c * 8192, &tmap_a, fb, m0 + 64 * c, k0);
                        }
                        if constexpr (!B_MN) {          // this CTA's half of the B tile: kBNLocal rows
                            const int nrow = (EPI == B200FM_EPI_SWIGLU) ? (rank == 0 ? n0 : args.n_half + n0) : n0 + static_cast<int>(rank) * kBNLocal;
                            tma_load_2d_2sm(sb, &tmap_b, fb, k0, nrow, kEvictLast);
                            if constexpr (kBNLocal == 256) tma_load_2d_2sm(sb + 128 * 128, &tmap_b, fb, k0, nrow + 128, kEvictLast);
                        } else {
#pragma unroll
                            for (int c = 0; c < kBNLocal / 64; ++c)
                                tma_load_2d_2sm(sb + c * 8192, &tmap_b, fb, n0 + static_cast<int>(rank) * kBNLocal + 64 * c, k0);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], kABytes + kBBytes);
                    if constexpr (!A_MN) {
                        tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);                 // box {64 k, 128 rows}
                    } else {
#pragma unroll
                        for (int c = 0; c < kBM / 64; ++c)                                   // box {64 m, 64 k-rows}
                            tma_load_2d(sa + c * 8192, &tmap_a, &full_bar[stage], m0 + 64 * c, k0);
                    }
                    if constexpr (!B_MN) {
                        if constexpr (EPI == B200FM_EPI_SWIGLU) {                            // box {64 k, BN/2 rows} x 2
                            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0, kEvictLast);
                            tma_load_2d(sb + (BN / 2) * 128, &tmap_b, &full_bar[stage], k0, args.n_half + n0, kEvictLast);
                        } else if constexpr (BN == 256) {                                    // box {64 k, 128 rows} x 2
                            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0, kEvictLast);
                            tma_load_2d(sb + 128 * 128, &tmap_b, &full_bar[stage], k0, n0 + 128, kEvictLast);
                        } else {
                            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0, kEvictLast);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < BN / 64; ++c)                                    // box {64 n, 64 k-rows}
                            tma_load_2d(sb + c * 8192, &tmap_b, &full_bar[stage], n0 + 64 * c, k0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ------------------------------ MMA issuer (single thread) ------------------------------
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int tile = my_first; tile < num_tiles; tile += my_step) {
                const int ks = tile / num_mn;
                const int kb_begin = ks * args.kb_per_split;
                const int num_kb = min(kb_begin + args.kb_per_split, num_kb_total) - kb_begin;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * SM::kStageBytes);
                    const uint32_t sb = sa + kABytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 64 * 128, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 64 * 128, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
                        if constexpr (CTA2) umma_bf16_2sm(d_tmem, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_bf16(d_tmem, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // smem slot reusable once these MMAs retire (CTA2: in both CTAs of the pair)
                    if constexpr (CTA2) umma_commit_2sm(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                // accumulator complete (CTA2: each CTA's epilogue reads its own 128 rows from its own TMEM)
                if constexpr (CTA2) umma_commit_2sm(&tfull_bar[as], 3); else umma_commit(&tfull_bar[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue warps 0..7 ------------------------------
        const int quarter = warp & 3, half = warp >> 2;
        int as = 0; uint32_t aphase = 0;
        const uint32_t tempty_leader0 = CTA2 ? mapa_u32(&tempty_bar[0], 0) : 0u;
        const uint32_t tempty_leader1 = CTA2 ? mapa_u32(&tempty_bar[1], 0) : 0u;
        for (int tile = my_first; tile < num_tiles; tile += my_step) {
            const int mn = tile % num_mn;
            const int m_blk = (mn % num_m_units) * cta_stride + static_cast<int>(rank);
            const int n_blk = mn / num_m_units;
            const int row = m_blk * kBM + quarter * 32 + lane;
            const bool row_ok = row < args.M;
            const uint32_t t_acc = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);

            if constexpr (EPI == B200FM_EPI_SWIGLU) {
                mbar_wait(&tfull_bar[as], aphase);
                tc_fence_after();
                constexpr int HB = BN / 2;
                const int n0 = n_blk * HB;
                __nv_bfloat16* ab = reinterpret_cast<__nv_bfloat16*>(args.out0) + static_cast<long long>(row) * args.ld0;
                __nv_bfloat16* gg = reinterpret_cast<__nv_bfloat16*>(args.out1) + static_cast<long long>(row) * args.ld1;
#pragma unroll 1
                for (int c = half * (HB / 32); c < (half + 1) * (HB / 32); ++c) {
                    uint32_t ra[16], rb[16];
                    tmem_ld_x16(t_acc + c * 16, ra);
                    tmem_ld_x16(t_acc + HB + c * 16, rb);
                    tmem_ld_wait();
                    const int n = n0 + c * 16;
                    if (row_ok && n < args.N) {
                        uint32_t pa[8], pb[8], pg[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float a0 = __uint_as_float(ra[2 * j]), a1 = __uint_as_float(ra[2 * j + 1]);
                            float b0 = __uint_as_float(rb[2 * j]), b1 = __uint_as_float(rb[2 * j + 1]);
                            if (args.bias) { a0 += args.bias[n + 2 * j]; a1 += args.bias[n + 2 * j + 1];
                                             b0 += args.bias[args.n_half + n + 2 * j]; b1 += args.bias[args.n_half + n + 2 * j + 1]; }
                            pa[j] = pack_bf16x2(a0, a1);
                            pb[j] = pack_bf16x2(b0, b1);
                            // reference numerics (fm_utils.py:143 under autocast): silu and the product are each rounded to bf16
                            const float2 ar = unpack_bf16x2(pa[j]), br = unpack_bf16x2(pb[j]);
                            pg[j] = pack_bf16x2(bf16_round(silu_f(ar.x)) * br.x, bf16_round(silu_f(ar.y)) * br.y);
                        }
                        if (n + 16 <= args.N) {
                            uint4* pa4 = reinterpret_cast<uint4*>(ab + n);
                            uint4* pb4 = reinterpret_cast<uint4*>(ab + args.n_half + n);
                            uint4* pg4 = reinterpret_cast<uint4*>(gg + n);
                            pa4[0] = make_uint4(pa[0], pa[1], pa[2], pa[3]); pa4[1] = make_uint4(pa[4], pa[5], pa[6], pa[7]);
                            pb4[0] = make_uint4(pb[0], pb[1], pb[2], pb[3]); pb4[1] = make_uint4(pb[4], pb[5], pb[6], pb[7]);
                            pg4[0] = make_uint4(pg[0], pg[1], pg[2], pg[3]); pg4[1] = make_uint4(pg[4], pg[5], pg[6], pg[7]);
                        } else {
                            for (int j = 0; j < 16 && n + j < args.N; ++j) {
                                const uint32_t wa = pa[j >> 1], wb = pb[j >> 1], wg = pg[j >> 1];
                                const uint16_t ha = (j & 1) ? (wa >> 16) : (wa & 0xffff);
                                const uint16_t hb = (j & 1) ? (wb >> 16) : (wb & 0xffff);
                                const uint16_t hg = (j & 1) ? (wg >> 16) : (wg & 0xffff);
                                reinterpret_cast<uint16_t*>(ab)[n + j] = ha;
                                reinterpret_cast<uint16_t*>(ab)[args.n_half + n + j] = hb;
                                reinterpret_cast<uint16_t*>(gg)[n + j] = hg;
                            }
                        }
                    }
                }
            } else {
                const int n0 = n_blk * BN;
                const float alpha = args.alpha * (args.alpha_dev ? __ldg(args.alpha_dev) : 1.0f);
                constexpr int kChunks = BN / 64;                 // 32-column chunks per epilogue warp
                const int c0 = half * kChunks;
                [[maybe_unused]] float4 rnext[8];
                [[maybe_unused]] bool rvec = false;
                if constexpr (EPI == B200FM_EPI_RESID) {
                    // the residual tile does not depend on the accumulator: fetch the first chunk before waiting for the MMAs,
                    // and every following chunk one iteration ahead
                    rvec = (args.ld0 & 3) == 0 && (args.ldr & 3) == 0;
                    const int n = n0 + c0 * 32;
                    if (rvec && row_ok && n + 32 <= args.N) {
                        const float4* rs4 = reinterpret_cast<const float4*>(args.resid + static_cast<long long>(row) * args.ldr + n);
#pragma unroll
                        for (int q = 0; q < 8; ++q) rnext[q] = __ldg(rs4 + q);
                    }
                }
                mbar_wait(&tfull_bar[as], aphase);
                tc_fence_after();
#pragma unroll 1
                for (int c = c0; c < c0 + kChunks; ++c) {
                    uint32_t r[32];
                    tmem_ld_x32(t_acc + c * 32, r);
                    const int n = n0 + c * 32;
                    [[maybe_unused]] float4 rcur[8];
                    if constexpr (EPI == B200FM_EPI_RESID) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) rcur[q] = rnext[q];
                        const int nn = n + 32;
                        if (c + 1 < c0 + kChunks && rvec && row_ok && nn + 32 <= args.N) {
                            const float4* rs4 = reinterpret_cast<const float4*>(args.resid + static_cast<long long>(row) * args.ldr + nn);
#pragma unroll
                            for (int q = 0; q < 8; ++q) rnext[q] = __ldg(rs4 + q);
                        }
                    }
                    tmem_ld_wait();
                    if (!row_ok || n >= args.N) continue;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    const bool full = (n + 32 <= args.N);
                    if (args.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (full || n + j < args.N) v[j] += args.bias[n + j];
                    }
                    if constexpr (EPI == B200FM_EPI_BF16 || EPI == B200FM_EPI_F32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= alpha;
                    }
                    if constexpr (EPI == B200FM_EPI_BF16 || EPI == B200FM_EPI_GELU) {
                        __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(args.out0) + static_cast<long long>(row) * args.ld0 + n;
                        uint32_t p[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) p[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                        if (full) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(o0)[q] = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
                        } else {
                            for (int j = 0; j < 32 && n + j < args.N; ++j)
                                reinterpret_cast<uint16_t*>(o0)[j] = (j & 1) ? (p[j >> 1] >> 16) : (p[j >> 1] & 0xffff);
                        }
                        if constexpr (EPI == B200FM_EPI_GELU) {
                            __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(args.out1) + static_cast<long long>(row) * args.ld1 + n;
                            uint32_t g[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float2 pr = unpack_bf16x2(p[j]);       // activation sees the bf16-rounded pre-activation
                                g[j] = args.act == 0 ? pack_bf16x2(gelu_erf(pr.x), gelu_erf(pr.y)) : pack_bf16x2(tanhf(pr.x), tanhf(pr.y));
                            }
                            if (full) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(o1)[q] = make_uint4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
                            } else {
                                for (int j = 0; j < 32 && n + j < args.N; ++j)
                                    reinterpret_cast<uint16_t*>(o1)[j] = (j & 1) ? (g[j >> 1] >> 16) : (g[j >> 1] & 0xffff);
                            }
                        }
                    } else if constexpr (EPI == B200FM_EPI_F32) {
                        float* o0 = reinterpret_cast<float*>(args.out0) + static_cast<long long>(row) * args.ld0 + n;
                        if (args.k_splits > 1) {
                            if (full && (args.ld0 & 3) == 0) {
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    atomicAdd(reinterpret_cast<float4*>(o0) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                            } else {
                                for (int j = 0; j < 32 && n + j < args.N; ++j) atomicAdd(o0 + j, v[j]);
                            }
                        } else if (full && (args.ld0 & 3) == 0) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                reinterpret_cast<float4*>(o0)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        } else {
                            for (int j = 0; j < 32 && n + j < args.N; ++j) o0[j] = v[j];
                        }
                    } else {   // EPI_RESID: out = resid + bf16_round(acc + bias)   (fp32 residual stream)
                        float* o0 = reinterpret_cast<float*>(args.out0) + static_cast<long long>(row) * args.ld0 + n;
                        if (full && rvec) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 x = rcur[q];
                                reinterpret_cast<float4*>(o0)[q] = make_float4(x.x + bf16_round(v[4 * q]), x.y + bf16_round(v[4 * q + 1]),
                                                                              x.z + bf16_round(v[4 * q + 2]), x.w + bf16_round(v[4 * q + 3]));
                            }
                        } else {
                            const float* rs = args.resid + static_cast<long long>(row) * args.ldr + n;
                            for (int j = 0; j < 32 && n + j < args.N; ++j) o0[j] = rs[j] + bf16_round(v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_cluster(as == 0 ? tempty_leader0 : tempty_leader1);   // the leader's MMA thread waits for both CTAs
                else mbar_arrive(&tempty_bar[as]);
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all(); else __syncthreads();       // the peer may still read this CTA's smem / signal its barriers
    if (warp == kEpiWarps) {
        tc_fence_after();
        if constexpr (CTA2) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <int BN, int LAYOUT, int EPI, bool CTA2>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t stream) {
    auto kern = gemm_kernel<BN, LAYOUT, EPI, CTA2>;
    constexpr int smem = GemmSmem<BN, CTA2>::kTotal;
    static bool configured = false;
    if (!configured) {
        B200FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int units = (CTA2 ? (a.num_m_blocks + 1) / 2 : a.num_m_blocks) * a.num_n_blocks * a.k_splits;
    if constexpr (CTA2) {
        const int clusters = units < sm_count() / 2 ? units : sm_count() / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        B200FM_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, a));
    } else {
        const int grid = units < sm_count() ? units : sm_count();
        kern<<<grid, kGemmThreads, smem, stream>>>(ta, tb, a);
    }
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

static bool use_cta_pairs() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200FM_GEMM_CTA_PAIRS");
        v = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_gemm_bf16(int layout, int epilogue, int M, int N, int K, const void* A, long long lda, const void* B,
                                long long ldb, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
                                const float* resid, long long ldr, float alpha, const float* alpha_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    B200FM_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    B200FM_CHECK(layout >= 0 && layout <= 2, "gemm: bad layout %d", layout);
    B200FM_CHECK(epilogue >= 0 && epilogue <= 5, "gemm: bad epilogue %d", epilogue);
    const int act = (epilogue == B200FM_EPI_TANH) ? 1 : 0;
    if (epilogue == B200FM_EPI_TANH) epilogue = B200FM_EPI_GELU;
    B200FM_CHECK(A && B && out0, "gemm: null pointer");
    B200FM_CHECK((lda % 8) == 0 && (ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 elements (16 B rows) for TMA, got %lld %lld", lda, ldb);
    if (epilogue == B200FM_EPI_BF16 || epilogue == B200FM_EPI_GELU || epilogue == B200FM_EPI_SWIGLU)
        B200FM_CHECK((ld0 % 8) == 0 && (reinterpret_cast<uintptr_t>(out0) & 15) == 0, "gemm: bf16 output must be 16 B aligned with ld %% 8 == 0");
    if (epilogue == B200FM_EPI_GELU || epilogue == B200FM_EPI_SWIGLU)
        B200FM_CHECK(out1 && (ld1 % 8) == 0 && (reinterpret_cast<uintptr_t>(out1) & 15) == 0, "gemm: second output missing or misaligned");
    if (epilogue == B200FM_EPI_RESID) B200FM_CHECK(resid != nullptr, "gemm: residual epilogue needs resid");
    if (epilogue == B200FM_EPI_SWIGLU) B200FM_CHECK(layout == LAYOUT_NT && (N % 8) == 0, "gemm: swiglu epilogue needs the NT layout and N %% 8 == 0");

    GemmArgs a;
    a.M = M; a.N = N; a.K = K;
    a.out0 = out0; a.ld0 = ld0; a.out1 = out1; a.ld1 = ld1; a.bias = bias; a.resid = resid; a.ldr = ldr;
    a.n_half = N; a.alpha = alpha; a.alpha_dev = alpha_dev; a.act = act;
    a.num_m_blocks = (M + kBM - 1) / kBM;

    // tile width: 256 when there is enough N to fill it and enough tiles to fill the machine, else 128
    a.k_splits = 1;
    a.kb_per_split = (K + kBK - 1) / kBK;
    int BN = 256;
    if (epilogue == B200FM_EPI_SWIGLU) {
        BN = 256;                                   // 128 a-columns + 128 b-columns per tile
        a.num_n_blocks = (N + 127) / 128;
    } else {
        const int tiles256 = a.num_m_blocks * ((N + 255) / 256);
        const int num_kb = (K + kBK - 1) / kBK;
        if (epilogue == B200FM_EPI_F32 && bias == nullptr && N > 128 && tiles256 < 2 * sm_count() && num_kb >= 32) {
            // few output tiles, long K (weight gradients): keep the 128x256 tile and split K across CTAs (fp32 atomics)
            BN = 256;
            int splits = (2 * sm_count() + tiles256 - 1) / tiles256;
            if (splits > num_kb / 8) splits = num_kb / 8;
            if (splits < 1) splits = 1;
            a.kb_per_split = (num_kb + splits - 1) / splits;
            a.k_splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;
        } else if (N <= 128 || tiles256 < sm_count()) {
            BN = 128;
        }
        a.num_n_blocks = (N + BN - 1) / BN;
    }
    if (a.k_splits > 1) B200FM_CUDA(cudaMemsetAsync(out0, 0, sizeof(float) * (size_t)M * (size_t)ld0 - sizeof(float) * (size_t)(ld0 - N), stream));

    CUtensorMap ta, tb;
    int rc;
    if (layout == LAYOUT_TN) rc = make_tmap_2d(&ta, A, TmapDtype::BF16, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, 64, true);
    else rc = make_tmap_2d(&ta, A, TmapDtype::BF16, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, 128, true);
    if (rc) return rc;
    if (layout == LAYOUT_NT) {
        const uint64_t rows = (epilogue == B200FM_EPI_SWIGLU) ? 2ull * N : (uint64_t)N;
        rc = make_tmap_2d(&tb, B, TmapDtype::BF16, (uint64_t)K, rows, (uint64_t)ldb * 2, 64, 128, true);
    } else {
        rc = make_tmap_2d(&tb, B, TmapDtype::BF16, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, 64, true);
    }
    if (rc) return rc;

    // CTA pairs whenever there are at least two 128-row blocks (a lone block would leave the peer CTA idle)
    // (BN == 256 only: the B tensor map's 128-row box is exactly one CTA's half of the tile)
    const bool pairs = use_cta_pairs() && a.num_m_blocks >= 2 && BN == 256;
#define B200FM_GEMM_CASE(BN_, L_, E_)                                                                  \
    if (BN == BN_ && layout == L_ && epilogue == E_)                                                    \
        return pairs ? launch_gemm<BN_, L_, E_, true>(ta, tb, a, stream) : launch_gemm<BN_, L_, E_, false>(ta, tb, a, stream);
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_BF16)
    B200FM_GEMM_CASE(128, LAYOUT_NT, B200FM_EPI_BF16)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_F32)
    B200FM_GEMM_CASE(128, LAYOUT_NT, B200FM_EPI_F32)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_RESID)
    B200FM_GEMM_CASE(128, LAYOUT_NT, B200FM_EPI_RESID)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_SWIGLU)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_GELU)
    B200FM_GEMM_CASE(128, LAYOUT_NT, B200FM_EPI_GELU)
    B200FM_GEMM_CASE(256, LAYOUT_NN, B200FM_EPI_BF16)
    B200FM_GEMM_CASE(128, LAYOUT_NN, B200FM_EPI_BF16)
    B200FM_GEMM_CASE(256, LAYOUT_NN, B200FM_EPI_F32)
    B200FM_GEMM_CASE(128, LAYOUT_NN, B200FM_EPI_F32)
    B200FM_GEMM_CASE(256, LAYOUT_TN, B200FM_EPI_F32)
    B200FM_GEMM_CASE(128, LAYOUT_TN, B200FM_EPI_F32)
    B200FM_GEMM_CASE(256, LAYOUT_TN, B200FM_EPI_BF16)
    B200FM_GEMM_CASE(128, LAYOUT_TN, B200FM_EPI_BF16)
#undef B200FM_GEMM_CASE
    B200FM_CHECK(false, "gemm: unsupported combination layout=%d epilogue=%d BN=%d", layout, epilogue, BN);
}
