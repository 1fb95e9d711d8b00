// Fused multi-head attention forward for sm_100a (head_dim 64): S = Q K^T on tcgen05 into TMEM, masked softmax in
// registers (one thread per query row), P (bf16) staged in 128B-swizzled smem, O = P V on tcgen05, normalised in the
// epilogue.  Replaces the materialised softmax(QK^T*scale masked_fill) V of fourm/models/fm_utils.py:160-180 (self),
// :197-219 (cross) and fourm/vq/models/vit_models.py:186-192 -- the [B,h,Nq,Nk] score tensors are never written.
//
// One work item = (batch b, head h, 128-query tile).  All keys of the item (Nk <= 128*NKT, NKT in {1,2}) are resident, so
// the softmax is exact two-pass-over-TMEM, no online rescaling.  CTA = 6 warps: 0-3 softmax/epilogue (TMEM lane quarter =
// warp), 4 = TMA producer + TMEM allocator, 5 = MMA issuer.  NKT=1 uses ~83 KB smem / 256 TMEM columns so two CTAs share
// an SM and overlap each other's softmax (CUDA cores) with MMA/TMA.
//
// Mask semantics (reference): mask byte != 0 -> score := -3e38 (finite, like masked_fill(-finfo.max)): a fully masked
// row yields UNIFORM attention over the Nk real keys; tile padding keys (j >= Nk) are excluded exactly (p = 0).
#include <cfloat>

#include "../../include/b200fm.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "tmap.cuh"

namespace b200fm {

template <int NKT>
struct AttnFwdSmem {
    static constexpr int kQ = 0;
    static constexpr int kK = 16384;
    static constexpr int kV = kK + NKT * 16384;
    static constexpr int kP = kV + NKT * 16384;
    static constexpr int kBar = kP + NKT * 32768;        // P: [128 q][NKT*128 keys] bf16 = NKT x two 64-key swizzle atoms
    static constexpr int kStage = kBar + 128;            // 4 warps x 2 KB store staging
    static constexpr int kTotal = kStage + 4 * 2048 + 1024;
    static constexpr int kTmemCols = NKT == 1 ? 256 : 512;
    static constexpr int kOCol = NKT * 128;
};

struct AttnFwdArgs {
    const uint8_t* mask;
    long long mask_b_stride, mask_q_stride;
    __nv_bfloat16* out;
    long long ldo;
    float* stats;          // [B, H, Nq, 2] : (row max of t = s*scale*log2e, 1/sum)
    int B, H, Nq, Nk, q_tiles, num_items;
    float scale_log2;      // scale * log2(e)
};

template <int NKT>
__global__ void __launch_bounds__(192, NKT == 1 ? 2 : 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnFwdArgs args) {
    using SM = AttnFwdSmem<NKT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBar);
    uint64_t* full_qk = bars + 0;
    uint64_t* full_v = bars + 1;
    uint64_t* free_qk = bars + 2;
    uint64_t* free_v = bars + 3;
    uint64_t* s_full = bars + 4;
    uint64_t* p_full = bars + 5;
    uint64_t* o_full = bars + 6;
    uint64_t* o_empty = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
            mbar_init(full_qk, 1); mbar_init(full_v, 1); mbar_init(free_qk, 1); mbar_init(free_v, 1);
            mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_empty, 4);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, SM::kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // TMEM is held: a dependent grid can no longer starve this one of columns
    pdl_wait();         // everything below touches global memory
    const int nk_steps = (args.Nk + 15) / 16;          // 16-key MMA steps actually needed for P V

    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
                const int h = item % args.H;
                const int qt = (item / args.H) % args.q_tiles;
                const int b = item / (args.H * args.q_tiles);
                const uint32_t par = it & 1;
                mbar_wait(free_qk, par ^ 1);
                mbar_arrive_expect_tx(full_qk, 16384 + NKT * 16384);
                tma_load_3d(smem + SM::kQ, &tmap_q, full_qk, h * 64, qt * 128, b, kEvictFirst);
#pragma unroll
                for (int t = 0; t < NKT; ++t) tma_load_3d(smem + SM::kK + t * 16384, &tmap_k, full_qk, h * 64, t * 128, b);
                mbar_wait(free_v, par ^ 1);
                mbar_arrive_expect_tx(full_v, NKT * 16384);
#pragma unroll
                for (int t = 0; t < NKT; ++t) tma_load_3d(smem + SM::kV + t * 16384, &tmap_v, full_v, h * 64, t * 128, b);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t kIdescS = make_idesc_bf16(128, 128, false, false);
            constexpr uint32_t kIdescO = make_idesc_bf16(128, 64, false, true);
            const uint32_t sq = smem_u32(smem + SM::kQ), sk = smem_u32(smem + SM::kK), sv = smem_u32(smem + SM::kV),
                           sp = smem_u32(smem + SM::kP);
            uint32_t it = 0;
            for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
                const uint32_t par = it & 1;
                mbar_wait(full_qk, par);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < NKT; ++t)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + t * 128, make_smem_desc(sq + k * 32, 16, 1024),
                                  make_smem_desc(sk + t * 16384 + k * 32, 16, 1024), kIdescS, k != 0);
                umma_commit(s_full);
                umma_commit(free_qk);
                mbar_wait(p_full, par);
                mbar_wait(full_v, par);
                mbar_wait(o_empty, par ^ 1);
                tc_fence_after();
                for (int kk = 0; kk < nk_steps; ++kk)
                    umma_bf16(tmem_base + SM::kOCol, make_smem_desc(sp + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                              make_smem_desc(sv + kk * 2048, 16384, 1024), kIdescO, kk != 0);
                umma_commit(o_full);
                umma_commit(free_v);
            }
        }
    } else {
        const int r = warp * 32 + lane;                   // query row inside the tile == TMEM lane
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        uint8_t* sp = smem + SM::kP;
        uint32_t it = 0;
        // The mask bytes of an item's row are requested one item AHEAD (right after the previous item's softmax) and turned into bits only
        // after the wait for the scores: in round 1 the softmax warps spent ~15 % of their samples stalled on these loads (ncu source page).
        uint4 mra[NKT * 4], mrb[NKT * 4];
        uint32_t mb[NKT * 4];
        bool mfast[NKT * 4];
        auto issue_mask = [&](int item_) {
            const int qt_ = (item_ / args.H) % args.q_tiles, b_ = item_ / (args.H * args.q_tiles);
            const int qrow_ = qt_ * 128 + r;
            const uint8_t* mrow_ = args.mask ? args.mask + b_ * args.mask_b_stride + (qrow_ < args.Nq ? qrow_ : 0) * args.mask_q_stride : nullptr;
#pragma unroll
            for (int c = 0; c < NKT * 4; ++c) mfast[c] = attn_mask_issue32(mrow_, c * 32, args.Nk, mra[c], mrb[c], mb[c]);
        };
        if (static_cast<int>(blockIdx.x) < args.num_items) issue_mask(blockIdx.x);
        for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
            const int h = item % args.H;
            const int qt = (item / args.H) % args.q_tiles;
            const int b = item / (args.H * args.q_tiles);
            const uint32_t par = it & 1;
            const int qrow = qt * 128 + r;
            const bool row_ok = qrow < args.Nq;
            uint32_t tail[NKT * 4];
#pragma unroll
            for (int c = 0; c < NKT * 4; ++c) {
                const int valid = args.Nk - c * 32;                                // keys of this chunk that exist
                tail[c] = valid >= 32 ? 0u : (valid <= 0 ? 0xffffffffu : (0xffffffffu << valid));
            }
            mbar_wait(s_full, par);
            tc_fence_after();
            uint32_t any_masked = 0u;
#pragma unroll
            for (int c = 0; c < NKT * 4; ++c) {
                if (mfast[c]) mb[c] = attn_mask_bits_from_raw(mra[c], mrb[c]);       // keys >= Nk report 0
                any_masked |= mb[c];
            }
            // pass 1: row maximum over the unmasked keys, on the raw scores (scale > 0 keeps the order); a masked key counts as
            // kMaskedScore in the log2 domain, exactly like masked_fill(-finfo.max) followed by the softmax's max
            float mu = -INFINITY;
#pragma unroll
            for (int c = 0; c < NKT * 4; ++c) {
                if (c * 32 < args.Nk) {
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
            float m = mu * args.scale_log2;                                        // -inf when every key is masked
            if (any_masked) m = fmaxf(m, kMaskedScore);
            const float p_masked = (m == kMaskedScore) ? 1.0f : 0.0f;              // exp2(kMaskedScore - m): uniform row iff all masked
            const float neg_m = -m;
            // pass 2: p = exp2(s * scale_log2 - m) in one fma + ex2, row sum, bf16 P into swizzled smem
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < NKT * 4; ++c) {
                uint32_t pk[16];
                if (c * 32 < args.Nk) {
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
            if (item + static_cast<int>(gridDim.x) < args.num_items) issue_mask(item + gridDim.x);    // next item's mask, hidden behind P V + epilogue

            const float inv = 1.0f / sum;
            if (row_ok && args.stats) {
                float2* st = reinterpret_cast<float2*>(args.stats) + ((static_cast<long long>(b) * args.H + h) * args.Nq + qrow);
                *st = make_float2(m, inv);
            }
            mbar_wait(o_full, par);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld_x32(t_lane + SM::kOCol, o0);
            tmem_ld_x32(t_lane + SM::kOCol + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
            {
                uint32_t* stg = reinterpret_cast<uint32_t*>(smem + SM::kStage) + warp * 512;
                __nv_bfloat16* base = args.out + static_cast<long long>(b) * args.Nq * args.ldo + h * 64;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(o0[2 * j]) * inv, __uint_as_float(o0[2 * j + 1]) * inv);
                attn_stage_store32(stg, lane, pk, base, args.ldo, qt * 128 + warp * 32, args.Nq);
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(o1[2 * j]) * inv, __uint_as_float(o1[2 * j + 1]) * inv);
                attn_stage_store32(stg, lane, pk, base + 32, args.ldo, qt * 128 + warp * 32, args.Nq);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SM::kTmemCols);
    }
}

template <int NKT>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdArgs& a,
                           cudaStream_t stream) {
    auto kern = attention_fwd_kernel<NKT>;
    constexpr int smem = AttnFwdSmem<NKT>::kTotal;
    static bool configured = false;
    if (!configured) {
        B200FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int sms = usable_sm_count();
    const int cap = sms * (NKT == 1 ? 2 : 1);
    const int grid = a.num_items < cap ? a.num_items : cap;
    B200FM_LAUNCH(kern, dim3(grid), dim3(192), smem, stream, 1, tq, tk, tv, a);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

// attention_fwd_long.cu: streamed-key variant for Nk > 256
int launch_attention_fwd_long(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const uint8_t* mask,
                              long long mask_b_stride, long long mask_q_stride, __nv_bfloat16* out, long long ldo, float* stats, int B,
                              int H, int Nq, int Nk, float scale_log2, cudaStream_t stream);

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_attention_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                    const uint8_t* mask, long long mask_b_stride, long long mask_q_stride, void* out,
                                    long long ldo, float* stats, int B, int H, int Nq, int Nk, float scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0 || H == 0 || Nq == 0) return 0;
    B200FM_CHECK(q && k && v && out, "attention_fwd: null pointer");
    B200FM_CHECK(Nk >= 1, "attention_fwd: Nk=%d", Nk);
    B200FM_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention_fwd: row strides must be multiples of 8 elements");
    B200FM_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention_fwd: out must be 16-byte aligned");
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_tmap_3d(&tq, q, TmapDtype::BF16, (uint64_t)H * 64, Nq, B, (uint64_t)ldq * 2, (uint64_t)Nq * ldq * 2, 64, 128, true))) return rc;
    if ((rc = make_tmap_3d(&tk, k, TmapDtype::BF16, (uint64_t)H * 64, Nk, B, (uint64_t)ldk * 2, (uint64_t)Nk * ldk * 2, 64, 128, true))) return rc;
    if ((rc = make_tmap_3d(&tv, v, TmapDtype::BF16, (uint64_t)H * 64, Nk, B, (uint64_t)ldv * 2, (uint64_t)Nk * ldv * 2, 64, 128, true))) return rc;
    if (Nk > 256)
        return launch_attention_fwd_long(tq, tk, tv, mask, mask_b_stride, mask_q_stride, reinterpret_cast<__nv_bfloat16*>(out), ldo, stats, B, H,
                                         Nq, Nk, scale * 1.4426950408889634f, stream);
    AttnFwdArgs a;
    a.mask = mask; a.mask_b_stride = mask_b_stride; a.mask_q_stride = mask_q_stride;
    a.out = reinterpret_cast<__nv_bfloat16*>(out); a.ldo = ldo; a.stats = stats;
    a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.q_tiles = (Nq + 127) / 128; a.num_items = B * H * a.q_tiles;
    a.scale_log2 = scale * 1.4426950408889634f;
    return Nk <= 128 ? launch_attn_fwd<1>(tq, tk, tv, a, stream) : launch_attn_fwd<2>(tq, tk, tv, a, stream);
}
