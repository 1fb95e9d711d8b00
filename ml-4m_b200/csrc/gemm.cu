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
    // device-side problem size (CUDA-graph friendly masked-token head: the per-modality row counts never visit the host):
    // dyn_mode 1: the number of output rows M is *dyn_dev (<= M), dyn_mode 2 (LAYOUT_TN): the contraction length K is *dyn_dev (<= K).
    // The launch (grid, tensor maps, split-K plan) is sized for the upper bounds; tiles / K blocks beyond the device value are skipped.
    const int* dyn_dev;
    int dyn_mode;
    // fused masked-token head (EPI_CE_STATS / EPI_CE_GRAD, fm.py:589-600): the logits tile never leaves the SM
    const long long* targets;     // int64 [M]
    const float* lse;             // EPI_CE_GRAD: fp32 [M] log-sum-exp of every row
    int tma_store;                // bf16 outputs leave through TMA stores (tensor maps tmap_o0..2) instead of LDS + STG
    int debug;                    // option "gemm_debug" (measurement only): 1 = epilogue reads TMEM but stores nothing, 2 = epilogue skipped
};

constexpr int EPI_CE_STATS = 6;   // out0 = float2 ws[M][ld0] per-(row, column slot) (max, sum exp(l - max)); out1 = fp32 tlogit[M] (target logit)
constexpr int EPI_CE_GRAD = 7;    // out0 = bf16 [M, ld0] softmax(l) - onehot(target), rows [M_dyn, roundup64(M_dyn)) zero-filled

// CTA2: a CTA pair (cluster of 2, one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2 -- each CTA stages its own
// 128 rows of A and only HALF of the B tile (BN/2 rows), so the per-SM smem fill per MMA-cycle drops by a third.
template <int BN, bool CTA2>
struct GemmSmem {
    static constexpr int kBBytes = CTA2 ? BN * 64 : BN * 128;
    static constexpr int kStageBytes = kBM * 128 + kBBytes;
    static constexpr int kStages = CTA2 ? (BN == 256 ? 5 : 6) : ((BN == 256) ? 3 : 5);
    static constexpr int kBarBytes = 1024;                                    // barriers + TMEM slot; keeps the staging area 1 KB aligned
    // per epilogue warp: fp32 staging 32 rows x (32 + 4 pad) floats (4608 B) OR three 2 KB boxes (32 rows x 32 bf16, 64B-swizzled) that
    // rotate as sources of TMA stores
    static constexpr int kEpiStageBytes = 3 * 2048;
    static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 8 * kEpiStageBytes + 1024;   // +1024 alignment slack
};

B200FM_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
B200FM_DEVINL float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// ---- warp-private smem staging: the accumulator arrives with thread = row (TMEM lane); going through a small smem tile turns
// the global accesses into row-contiguous segments (full 32 B sectors, 64-128 B per row) instead of 32 strided 16 B pieces.
//
// bf16: 32 rows x 16 packed words, XOR-swizzled 16 B quads (conflict-free for both the row-wise write and the 8-row read)
B200FM_DEVINL void stage_store_bf16(uint32_t* stg, int lane, const uint32_t (&p)[16], __nv_bfloat16* out, long long ld, int row_base, int n,
                                    int M, int N, bool vec_ok, int dbg = 0) {
    if (dbg == 4) {      // measurement: thread = row writes its 64 contiguous bytes straight from registers (no shared-memory staging)
        const int grow = row_base + lane;
        if (grow < M && n + 32 <= N && vec_ok) {
            uint4* dst = reinterpret_cast<uint4*>(out + static_cast<long long>(grow) * ld + n);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        }
        return;
    }
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(stg + lane * 16 + 4 * (q ^ sw)) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = i * 8 + (lane >> 2), quad = lane & 3;
        const uint4 w = *reinterpret_cast<const uint4*>(stg + rl * 16 + 4 * (quad ^ ((rl >> 1) & 3)));
        const int grow = row_base + rl, gcol = n + quad * 8;
        if (grow < M && gcol < N && dbg != 3) {          // dbg 3 (measurement): staging traffic only, no global stores
            __nv_bfloat16* dst = out + static_cast<long long>(grow) * ld + gcol;
            if (vec_ok && gcol + 8 <= N) {
                *reinterpret_cast<uint4*>(dst) = w;
            } else {
                const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
                for (int e = 0; e < 8 && gcol + e < N; ++e)
                    reinterpret_cast<uint16_t*>(dst)[e] = (e & 1) ? (ww[e >> 1] >> 16) : (ww[e >> 1] & 0xffff);
            }
        }
    }
    __syncwarp();
}
// fp32: 32 rows x 36 floats (4 pad): row-wise float4 writes and 4-row x 32-column float4 reads are both conflict-free
B200FM_DEVINL void stage_write_f32(float* stg, int lane, const uint32_t (&r)[32]) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + lane * 36 + 4 * q) =
            make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
    __syncwarp();
}

// bf16 through the TMA: the warp writes its 32 x 32 tile (thread = row, 64 B) into a 2 KB box with the 64B-swizzle pattern (16-byte
// chunk index ^ ((row >> 1) & 3): conflict-free for the row-wise writes), one lane issues the bulk tensor store.  Compared with
// stage_store_bf16 the tile crosses the shared-memory / L1 data path twice (STS + TMA read) instead of three times (STS + LDS + STG) --
// that path, not the tensor pipe, is what bounds this kernel (tools/gemm_probe.py) -- and rows / columns beyond the tensor are clipped
// by the hardware.  Three boxes per warp rotate: a box is rewritten only after the store issued from it three stores ago has been read.
B200FM_DEVINL void tma_stage_store_bf16(uint32_t* stg, uint32_t& box_i, int lane, const uint32_t (&p)[16], const CUtensorMap* map, int col, int row) {
    uint32_t* buf = stg + (box_i % 3u) * 512u;
    ++box_i;
    if (lane == 0) tma_store_wait_read<2>();
    __syncwarp();
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(buf + lane * 16 + 4 * (q ^ sw)) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map, buf, col, row);
        tma_store_commit();
    }
}

template <int BN, int LAYOUT, int EPI, bool CTA2>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_o0,
            const __grid_constant__ CUtensorMap tmap_o1, const __grid_constant__ CUtensorMap tmap_o2, const GemmArgs args) {
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
    pdl_trigger();      // TMEM is held: a dependent grid can no longer starve this one of columns
    pdl_wait();         // everything below touches global memory
    const uint32_t tmem_base = *tmem_slot;
    int M_ = args.M, K_ = args.K, num_m_blocks = args.num_m_blocks;
    if (args.dyn_dev != nullptr) {
        const int d = max(0, __ldg(args.dyn_dev));
        if (args.dyn_mode == 1) { M_ = min(d, args.M); num_m_blocks = (M_ + kBM - 1) / kBM; }
        else K_ = min(d, args.K);
    }
    const int num_m_units = (num_m_blocks + cta_stride - 1) / cta_stride;           // CTA2: pairs of 128-row blocks
    const int num_mn = num_m_units * args.num_n_blocks;
    const int num_tiles = num_mn * args.k_splits;               // work items: (k slice, m unit, n block), n fastest
    const int num_kb_total = (K_ + kBK - 1) / kBK;

    if (warp == kEpiWarps) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = my_first; tile < num_tiles; tile += my_step) {
                const int mn = tile % num_mn, ks = tile / num_mn;
                const int m_blk = (mn / args.num_n_blocks) * cta_stride + static_cast<int>(rank);   // n fastest: concurrent CTAs share the A rows,
                const int n_blk = mn % args.num_n_blocks;                                           // B (weights) stays L2-resident -> A streams from HBM once
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
        uint32_t box_i = 0;                                               // TMA-store boxes this warp has used (3 rotate)
        for (int tile = my_first; tile < num_tiles; tile += my_step) {
            const int mn = tile % num_mn;
            const int m_blk = (mn / args.num_n_blocks) * cta_stride + static_cast<int>(rank);
            const int n_blk = mn % args.num_n_blocks;
            const int row_base = m_blk * kBM + quarter * 32;            // first of this warp's 32 rows
            const int row = row_base + lane;
            const bool row_ok = row < M_;
            // a K slice that lies entirely beyond the (device-side) contraction length: no MMA ran, the accumulator is undefined
            const bool empty_acc = (tile / num_mn) * args.kb_per_split >= num_kb_total;
            const uint32_t t_acc = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
            float* stg_f = reinterpret_cast<float*>(smem + kStages * SM::kStageBytes + SM::kBarBytes) + warp * (SM::kEpiStageBytes / 4);
            uint32_t* stg_u = reinterpret_cast<uint32_t*>(stg_f);
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();

            if constexpr (EPI == B200FM_EPI_SWIGLU) {
                constexpr int HB = BN / 2;
                const int n0 = n_blk * HB;
                __nv_bfloat16* ab = reinterpret_cast<__nv_bfloat16*>(args.out0);
                __nv_bfloat16* gg = reinterpret_cast<__nv_bfloat16*>(args.out1);
                const bool vec_ok = (args.ld0 & 7) == 0 && (args.ld1 & 7) == 0 && (args.n_half & 7) == 0;
#pragma unroll 1
                for (int c = half * (HB / 64); c < (half + 1) * (HB / 64) && args.debug != 2; ++c) {        // 32 gate columns per iteration
                    uint32_t ra[32], rb[32];
                    tmem_ld_x32(t_acc + c * 32, ra);
                    tmem_ld_x32(t_acc + HB + c * 32, rb);
                    tmem_ld_wait();
                    if (args.debug == 1) continue;
                    const int n = n0 + c * 32;
                    uint32_t pa[16], pb[16], pg[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a0 = __uint_as_float(ra[2 * j]), a1 = __uint_as_float(ra[2 * j + 1]);
                        float b0 = __uint_as_float(rb[2 * j]), b1 = __uint_as_float(rb[2 * j + 1]);
                        if (args.bias && n + 2 * j + 1 < args.N) {
                            a0 += args.bias[n + 2 * j]; a1 += args.bias[n + 2 * j + 1];
                            b0 += args.bias[args.n_half + n + 2 * j]; b1 += args.bias[args.n_half + n + 2 * j + 1];
                        }
                        pa[j] = pack_bf16x2(a0, a1);
                        pb[j] = pack_bf16x2(b0, b1);
                        // reference numerics (fm_utils.py:143 under autocast): silu and the product are each rounded to bf16
                        const float2 ar = unpack_bf16x2(pa[j]), br = unpack_bf16x2(pb[j]);
                        pg[j] = pack_bf16x2(bf16_round(silu_f(ar.x)) * br.x, bf16_round(silu_f(ar.y)) * br.y);
                    }
                    if (args.tma_store) {
                        tma_stage_store_bf16(stg_u, box_i, lane, pa, &tmap_o0, n, row_base);
                        tma_stage_store_bf16(stg_u, box_i, lane, pb, &tmap_o1, n, row_base);
                        tma_stage_store_bf16(stg_u, box_i, lane, pg, &tmap_o2, n, row_base);
                    } else {
                        stage_store_bf16(stg_u, lane, pa, ab, args.ld0, row_base, n, M_, args.N, vec_ok, args.debug);
                        stage_store_bf16(stg_u, lane, pb, ab + args.n_half, args.ld0, row_base, n, M_, args.N, vec_ok, args.debug);
                        stage_store_bf16(stg_u, lane, pg, gg, args.ld1, row_base, n, M_, args.N, vec_ok, args.debug);
                    }
                }
            } else {
                const int n0 = n_blk * BN;
                const float alpha = args.alpha * (args.alpha_dev ? __ldg(args.alpha_dev) : 1.0f);
                constexpr int kChunks = BN / 64;                 // 32-column chunks per epilogue warp
                const int c0 = half * kChunks;
                if constexpr (EPI == EPI_CE_STATS || EPI == EPI_CE_GRAD) {
                    // ---- fused cross-entropy head: thread = row holds its logits of this tile in registers, 32 columns at a time ----
                    constexpr float kLog2e = 1.4426950408889634f;
                    const int tgt = row_ok ? static_cast<int>(args.targets[row]) : -1;
                    if constexpr (EPI == EPI_CE_STATS) {
                        float run_m = -INFINITY, run_s = 0.f;
#pragma unroll 1
                        for (int c = c0; c < c0 + kChunks; ++c) {
                            const int n = n0 + c * 32;
                            if (n >= args.N) break;               // warp-uniform
                            uint32_t r[32];
                            tmem_ld_x32(t_acc + c * 32, r);
                            tmem_ld_wait();
                            float cm = -INFINITY;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (n + j < args.N) cm = fmaxf(cm, __uint_as_float(r[j]));
                            const float m_new = fmaxf(run_m, cm);
                            float cs = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (n + j < args.N) cs += exp2f((__uint_as_float(r[j]) - m_new) * kLog2e);
                            run_s = run_s * exp2f((run_m - m_new) * kLog2e) + cs;      // run_m = -inf on the first chunk: factor 0
                            run_m = m_new;
                            if (tgt >= n && tgt < n + 32) {
                                float tl = 0.f;
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (n + j == tgt) tl = __uint_as_float(r[j]);
                                reinterpret_cast<float*>(args.out1)[row] = tl;
                            }
                        }
                        if (row_ok && n0 + c0 * 32 < args.N)
                            reinterpret_cast<float2*>(args.out0)[static_cast<long long>(row) * args.ld0 + n_blk * 2 + half] = make_float2(run_m, run_s);
                    } else {
                        const float lse2 = row_ok ? args.lse[row] * kLog2e : 0.f;
                        const int m_fill = min((M_ + 63) & ~63, args.M);          // rows [M_, m_fill) are written as zeros (see header)
#pragma unroll 1
                        for (int c = c0; c < c0 + kChunks; ++c) {
                            const int n = n0 + c * 32;
                            if (n >= args.N) break;
                            uint32_t r[32];
                            tmem_ld_x32(t_acc + c * 32, r);
                            tmem_ld_wait();
                            uint32_t p[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float v0 = 0.f, v1 = 0.f;
                                if (row_ok) {
                                    v0 = exp2f(fmaf(__uint_as_float(r[2 * j]), kLog2e, -lse2)) - (n + 2 * j == tgt ? 1.0f : 0.0f);
                                    v1 = exp2f(fmaf(__uint_as_float(r[2 * j + 1]), kLog2e, -lse2)) - (n + 2 * j + 1 == tgt ? 1.0f : 0.0f);
                                }
                                p[j] = pack_bf16x2(v0, v1);
                            }
                            stage_store_bf16(stg_u, lane, p, reinterpret_cast<__nv_bfloat16*>(args.out0), args.ld0, row_base, n, m_fill, args.N, (args.ld0 & 7) == 0);
                        }
                    }
                } else
#pragma unroll 1
                for (int c = c0; c < c0 + kChunks && args.debug != 2; ++c) {
                    uint32_t r[32];
                    tmem_ld_x32(t_acc + c * 32, r);
                    tmem_ld_wait();
                    const int n = n0 + c * 32;
                    if (n >= args.N) continue;                   // warp-uniform
                    if (args.debug == 1) continue;
                    if (empty_acc) {                             // warp-uniform; only reachable with a device-side K (dyn_mode 2)
                        if (args.k_splits > 1) continue;         // nothing to add
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = 0u;  // K == 0: the product is exactly zero
                    }
                    if constexpr (EPI == B200FM_EPI_BF16 || EPI == B200FM_EPI_GELU) {
                        uint32_t p[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float v0 = __uint_as_float(r[2 * j]), v1 = __uint_as_float(r[2 * j + 1]);
                            if (args.bias) {
                                if (n + 2 * j < args.N) v0 += args.bias[n + 2 * j];
                                if (n + 2 * j + 1 < args.N) v1 += args.bias[n + 2 * j + 1];
                            }
                            if constexpr (EPI == B200FM_EPI_BF16) { v0 *= alpha; v1 *= alpha; }
                            p[j] = pack_bf16x2(v0, v1);
                        }
                        const bool vec0 = (args.ld0 & 7) == 0;
                        if (EPI == B200FM_EPI_BF16 && args.tma_store) tma_stage_store_bf16(stg_u, box_i, lane, p, &tmap_o0, n, row_base);
                        else stage_store_bf16(stg_u, lane, p, reinterpret_cast<__nv_bfloat16*>(args.out0), args.ld0, row_base, n, M_, args.N, vec0, args.debug);
                        if constexpr (EPI == B200FM_EPI_GELU) {
                            uint32_t g[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float2 pr = unpack_bf16x2(p[j]);       // activation sees the bf16-rounded pre-activation
                                g[j] = args.act == 0 ? pack_bf16x2(gelu_erf(pr.x), gelu_erf(pr.y)) : pack_bf16x2(tanhf(pr.x), tanhf(pr.y));
                            }
                            stage_store_bf16(stg_u, lane, g, reinterpret_cast<__nv_bfloat16*>(args.out1), args.ld1, row_base, n, M_, args.N, (args.ld1 & 7) == 0);
                        }
                    } else {
                        // fp32 outputs: transpose through smem, then each lane owns 4 consecutive columns of 8 rows
                        stage_write_f32(stg_f, lane, r);
                        const int c4 = (lane & 7) * 4;
                        const int gcol = n + c4;
                        const bool vec = (args.ld0 & 3) == 0 && (EPI != B200FM_EPI_RESID || (args.ldr & 3) == 0) && gcol + 4 <= args.N;
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (args.bias && gcol < args.N) {
                            b4.x = args.bias[gcol];
                            if (gcol + 1 < args.N) b4.y = args.bias[gcol + 1];
                            if (gcol + 2 < args.N) b4.z = args.bias[gcol + 2];
                            if (gcol + 3 < args.N) b4.w = args.bias[gcol + 3];
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rl = i * 4 + (lane >> 3);
                            const int grow = row_base + rl;
                            float4 a = *reinterpret_cast<const float4*>(stg_f + rl * 36 + c4);
                            a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
                            if (grow >= M_ || gcol >= args.N) continue;
                            float* o0 = reinterpret_cast<float*>(args.out0) + static_cast<long long>(grow) * args.ld0 + gcol;
                            if constexpr (EPI == B200FM_EPI_F32) {
                                a.x *= alpha; a.y *= alpha; a.z *= alpha; a.w *= alpha;
                                if (vec) {
                                    if (args.k_splits > 1) atomicAdd(reinterpret_cast<float4*>(o0), a);
                                    else *reinterpret_cast<float4*>(o0) = a;
                                } else {
                                    const float av[4] = {a.x, a.y, a.z, a.w};
                                    for (int e = 0; e < 4 && gcol + e < args.N; ++e) {
                                        if (args.k_splits > 1) atomicAdd(o0 + e, av[e]); else o0[e] = av[e];
                                    }
                                }
                            } else {   // EPI_RESID: out = resid + bf16_round(acc + bias)   (fp32 residual stream)
                                const float* rs = args.resid + static_cast<long long>(grow) * args.ldr + gcol;
                                if (vec) {
                                    const float4 x = __ldg(reinterpret_cast<const float4*>(rs));
                                    *reinterpret_cast<float4*>(o0) = make_float4(x.x + bf16_round(a.x), x.y + bf16_round(a.y), x.z + bf16_round(a.z), x.w + bf16_round(a.w));
                                } else {
                                    const float av[4] = {a.x, a.y, a.z, a.w};
                                    for (int e = 0; e < 4 && gcol + e < args.N; ++e) o0[e] = rs[e] + bf16_round(av[e]);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            (void)row; (void)row_ok;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_cluster(as == 0 ? tempty_leader0 : tempty_leader1);   // the leader's MMA thread waits for both CTAs
                else mbar_arrive(&tempty_bar[as]);
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (lane == 0) tma_store_wait_read<0>();                          // the boxes must outlive the TMA's reads
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
static int sm_count() { return usable_sm_count(); }

template <int BN, int LAYOUT, int EPI, bool CTA2>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap (&to)[3], const GemmArgs& a, cudaStream_t stream) {
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
        B200FM_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(kGemmThreads), smem, stream, 2, ta, tb, to[0], to[1], to[2], a));
    } else {
        const int grid = units < sm_count() ? units : sm_count();
        B200FM_CUDA(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), smem, stream, 1, ta, tb, to[0], to[1], to[2], a));
    }
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

static bool use_cta_pairs() { return option(kOptGemmCtaPairs) != 0; }

// gemv.cu: weight-streaming kernel for M <= 8 rows (autoregressive decode)
bool gemv_applicable(int layout, int epilogue, int M, int N, int K, long long lda, long long ldb);
int launch_gemv(int epilogue, int M, int N, int K, const void* A, long long lda, const void* W, long long ldb, void* out0, long long ld0,
                void* out1, long long ld1, const float* bias, const float* resid, long long ldr, float alpha, const float* alpha_dev,
                cudaStream_t stream);

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_gemm_bf16(int layout, int epilogue, int M, int N, int K, const void* A, long long lda, const void* B,
                                long long ldb, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
                                const float* resid, long long ldr, float alpha, const float* alpha_dev, void* stream_) {
    return b200fm_gemm_bf16_dyn(layout, epilogue, M, N, K, A, lda, B, ldb, out0, ld0, out1, ld1, bias, resid, ldr, alpha, alpha_dev, nullptr, 0,
                                stream_);
}

static int gemm_impl(int layout, int epilogue, int M, int N, int K, const void* A, long long lda, const void* B,
                     long long ldb, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
                     const float* resid, long long ldr, float alpha, const float* alpha_dev, const int* dyn_dev,
                     int dyn_mode, const long long* targets, const float* lse, void* stream_);

extern "C" int b200fm_gemm_bf16_dyn(int layout, int epilogue, int M, int N, int K, const void* A, long long lda, const void* B,
                                    long long ldb, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
                                    const float* resid, long long ldr, float alpha, const float* alpha_dev, const int* dyn_dev,
                                    int dyn_mode, void* stream_) {
    B200FM_CHECK(epilogue >= 0 && epilogue <= 5, "gemm: bad epilogue %d", epilogue);
    return gemm_impl(layout, epilogue, M, N, K, A, lda, B, ldb, out0, ld0, out1, ld1, bias, resid, ldr, alpha, alpha_dev, dyn_dev, dyn_mode,
                     nullptr, nullptr, stream_);
}

static int gemm_impl(int layout, int epilogue, int M, int N, int K, const void* A, long long lda, const void* B,
                     long long ldb, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
                     const float* resid, long long ldr, float alpha, const float* alpha_dev, const int* dyn_dev,
                     int dyn_mode, const long long* targets, const float* lse, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    B200FM_CHECK(dyn_dev == nullptr || (dyn_mode == 1 && layout != LAYOUT_TN) || (dyn_mode == 2 && layout == LAYOUT_TN && epilogue == B200FM_EPI_F32),
                 "gemm: dyn_mode %d does not fit layout %d / epilogue %d (1: rows of NT/NN, 2: contraction length of TN + EPI_F32)", dyn_mode, layout, epilogue);
    B200FM_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    B200FM_CHECK(layout >= 0 && layout <= 2, "gemm: bad layout %d", layout);
    B200FM_CHECK(epilogue >= 0 && epilogue <= 7, "gemm: bad epilogue %d", epilogue);
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
    if (dyn_dev == nullptr && option(kOptGemv) != 0 && gemv_applicable(layout, epilogue, M, N, K, lda, ldb))
        return launch_gemv(epilogue, M, N, K, A, lda, B, ldb, out0, ld0, out1, ld1, bias, resid, ldr, alpha, alpha_dev, stream);

    GemmArgs a;
    a.M = M; a.N = N; a.K = K;
    a.out0 = out0; a.ld0 = ld0; a.out1 = out1; a.ld1 = ld1; a.bias = bias; a.resid = resid; a.ldr = ldr;
    a.n_half = N; a.alpha = alpha; a.alpha_dev = alpha_dev; a.act = act;
    a.dyn_dev = dyn_dev; a.dyn_mode = dyn_dev ? dyn_mode : 0;
    a.targets = targets; a.lse = lse; a.debug = option(kOptGemmDebug);
    a.num_m_blocks = (M + kBM - 1) / kBM;

    // tile width: 256 when there is enough N to fill it and enough tiles to fill the machine, else 128
    a.k_splits = 1;
    a.kb_per_split = (K + kBK - 1) / kBK;
    int BN = 256;
    if (epilogue == B200FM_EPI_SWIGLU) {
        BN = 256;                                   // 128 a-columns + 128 b-columns per tile
        a.num_n_blocks = (N + 127) / 128;
    } else if (epilogue == EPI_CE_STATS || epilogue == EPI_CE_GRAD) {
        BN = 256;                                   // fixed: the statistics workspace is indexed by (n block, column half)
        a.num_n_blocks = (N + 255) / 256;
    } else {
        const int tiles256 = a.num_m_blocks * ((N + 255) / 256);
        const int num_kb = (K + kBK - 1) / kBK;
        if (epilogue == B200FM_EPI_F32 && bias == nullptr && N > 128 && tiles256 < 2 * sm_count() && num_kb >= 32) {
            // few output tiles, long K (weight gradients): keep the 128x256 tile and split K across CTAs (fp32 atomics).
            // Work items = output units x K slices are dealt round-robin to the resident CTAs (pairs), so the slice count is
            // chosen to fill whole rounds: best machine fill first, fewer slices (less atomic traffic) on ties.
            BN = 256;
            const bool will_pair = use_cta_pairs() && a.num_m_blocks >= 2;
            const int units = (will_pair ? (a.num_m_blocks + 1) / 2 : a.num_m_blocks) * ((N + 255) / 256);
            const int slots = will_pair ? sm_count() / 2 : sm_count();
            int best_s = 1;
            double best_score = -1.0;
            for (int s = 1; s <= num_kb / 8 && s <= 32; ++s) {
                const int per = (num_kb + s - 1) / s;
                const int real = (num_kb + per - 1) / per;                 // slices actually produced
                if (real != s) continue;
                const long long items = 1ll * units * s;
                const long long rounds = (items + slots - 1) / slots;
                const double fill = static_cast<double>(items) / static_cast<double>(rounds * slots);
                const double score = fill - 0.004 * s;
                if (score > best_score) { best_score = score; best_s = s; }
            }
            a.kb_per_split = (num_kb + best_s - 1) / best_s;
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

    // bf16 outputs through TMA stores (EPI_BF16: out0; EPI_SWIGLU: a | b halves of out0 and the gate): 32 x 32 boxes, 64B swizzle.
    // Not with a device-side row count (the hardware clips at the tensor's static bounds) nor with rows that are not 16-byte multiples.
    CUtensorMap to[3] = {ta, ta, ta};
    a.tma_store = 0;
    if (option(kOptGemmTmaStore) != 0 && dyn_dev == nullptr && a.debug == 0 && (epilogue == B200FM_EPI_BF16 || epilogue == B200FM_EPI_SWIGLU) &&
        (ld0 % 8) == 0 && (reinterpret_cast<uintptr_t>(out0) & 15) == 0) {
        if (epilogue == B200FM_EPI_BF16) {
            rc = make_tmap_2d_sw(&to[0], out0, TmapDtype::BF16, (uint64_t)N, (uint64_t)M, (uint64_t)ld0 * 2, 32, 32, 64);
            if (rc) return rc;
            a.tma_store = 1;
        } else if ((ld1 % 8) == 0 && (reinterpret_cast<uintptr_t>(out1) & 15) == 0 && (N % 8) == 0) {
            rc = make_tmap_2d_sw(&to[0], out0, TmapDtype::BF16, (uint64_t)N, (uint64_t)M, (uint64_t)ld0 * 2, 32, 32, 64);
            if (rc) return rc;
            rc = make_tmap_2d_sw(&to[1], reinterpret_cast<const __nv_bfloat16*>(out0) + N, TmapDtype::BF16, (uint64_t)N, (uint64_t)M, (uint64_t)ld0 * 2, 32, 32, 64);
            if (rc) return rc;
            rc = make_tmap_2d_sw(&to[2], out1, TmapDtype::BF16, (uint64_t)N, (uint64_t)M, (uint64_t)ld1 * 2, 32, 32, 64);
            if (rc) return rc;
            a.tma_store = 1;
        }
    }

    // CTA pairs whenever there are at least two 128-row blocks (a lone block would leave the peer CTA idle)
    // (BN == 256 only: the B tensor map's 128-row box is exactly one CTA's half of the tile)
    const bool pairs = use_cta_pairs() && a.num_m_blocks >= 2 && BN == 256;
#define B200FM_GEMM_CASE(BN_, L_, E_)                                                                  \
    if (BN == BN_ && layout == L_ && epilogue == E_)                                                    \
        return pairs ? launch_gemm<BN_, L_, E_, true>(ta, tb, to, a, stream) : launch_gemm<BN_, L_, E_, false>(ta, tb, to, a, stream);
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_BF16)
    B200FM_GEMM_CASE(128, LAYOUT_NT, B200FM_EPI_BF16)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_F32)
    B200FM_GEMM_CASE(128, LAYOUT_NT, B200FM_EPI_F32)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_RESID)
    B200FM_GEMM_CASE(128, LAYOUT_NT, B200FM_EPI_RESID)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_SWIGLU)
    B200FM_GEMM_CASE(256, LAYOUT_NT, B200FM_EPI_GELU)
    B200FM_GEMM_CASE(256, LAYOUT_NT, EPI_CE_STATS)
    B200FM_GEMM_CASE(256, LAYOUT_NT, EPI_CE_GRAD)
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


// ---- fused masked-token head: logits GEMM with the cross-entropy in its epilogue (no [rows, V] fp32 logits tensor) -----------------
namespace b200fm {
// one warp per row: combine the per-(n block, column half) partials into the row's log-sum-exp and its loss
__global__ void __launch_bounds__(256)
ce_reduce_kernel(const float2* __restrict__ ws, int slots, int V, const float* __restrict__ tlogit, const int* __restrict__ n_dev, long long M,
                 float* __restrict__ lse, float* __restrict__ loss_rows) {
    pdl_enter();
    const long long row = (blockIdx.x * 256ll + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const long long n = n_dev ? min((long long)max(0, __ldg(n_dev)), M) : M;
    if (row >= n) { if (lane == 0) { loss_rows[row] = 0.f; lse[row] = 0.f; } return; }
    float m = -INFINITY;
    for (int s = lane; s < slots; s += 32)
        if ((s >> 1) * 256 + (s & 1) * 128 < V) m = fmaxf(m, ws[row * slots + s].x);
    m = warp_max(m);
    float acc = 0.f;
    for (int s = lane; s < slots; s += 32)
        if ((s >> 1) * 256 + (s & 1) * 128 < V) { const float2 p = ws[row * slots + s]; acc += p.y * __expf(p.x - m); }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float l = m + logf(acc);
        lse[row] = l;
        loss_rows[row] = l - tlogit[row];
    }
}
}  // namespace b200fm

extern "C" int b200fm_head_ce_ws_slots(int V) { return 2 * ((V + 255) / 256); }

extern "C" int b200fm_head_ce(const void* h, long long ldh, const void* W, long long ldw, const int64_t* targets, const int* n_dev, int M,
                              int V, int K, float* ws, float* tlogit, float* lse, float* loss_rows, void* dlogits, long long ldd,
                              void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (M == 0) return 0;
    B200FM_CHECK(h && W && targets && ws && tlogit && lse && loss_rows, "head_ce: null pointer");
    B200FM_CHECK(dlogits == nullptr || (ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0), "head_ce: dlogits must be 16 B aligned with ldd %% 8 == 0");
    const int slots = b200fm_head_ce_ws_slots(V);
    int rc = gemm_impl(LAYOUT_NT, EPI_CE_STATS, M, V, K, h, ldh, W, ldw, ws, slots, tlogit, 0, nullptr, nullptr, 0, 1.0f, nullptr, n_dev, n_dev ? 1 : 0,
                       reinterpret_cast<const long long*>(targets), nullptr, stream_);
    if (rc) return rc;
    B200FM_LAUNCH(ce_reduce_kernel, dim3((unsigned)((M * 32ll + 255) / 256)), dim3(256), 0, stream, 1, reinterpret_cast<const float2*>(ws), slots, V, tlogit, n_dev,
                  (long long)M, lse, loss_rows);
    B200FM_CUDA(cudaGetLastError());
    if (dlogits == nullptr) return 0;
    return gemm_impl(LAYOUT_NT, EPI_CE_GRAD, M, V, K, h, ldh, W, ldw, dlogits, ldd, nullptr, 0, nullptr, nullptr, 0, 1.0f, nullptr, n_dev, n_dev ? 1 : 0,
                     reinterpret_cast<const long long*>(targets), lse, stream_);
}
