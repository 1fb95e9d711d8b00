// Fused multi-head attention backward for sm_100a (head_dim 64).  Autograd counterpart of attention_fwd.cu for the
// reference's materialised attention (fourm/models/fm_utils.py:160-180, 197-219): given dO it recomputes
// P = softmax(mask(Q K^T * scale)) from the saved row statistics and produces dQ, dK, dV without ever writing a
// [B,h,Nq,Nk] tensor.
//
// One work item = (batch b, head h); the CTA walks key tiles (outer) x query tiles (inner), 128 x 128 each (<= 2 x 2):
//     S  = Q K^T,  dP = dO V^T                 (tcgen05, TMEM cols [0,128) and [128,256))
//     P  = exp2(S*scale*log2e - m) / sum,  dS = P o (dP - D) * scale,  masked positions: dS = 0 (masked_fill blocks grad)
//     dV += P^T dO,  dK += dS^T Q              (M = keys: the bf16 P / dS smem tiles are consumed as MN-major A operands)
//     dQ += dS K                               (same dS tile consumed as a K-major A operand)
// dV, dK live in TMEM across the query loop, dQ (one 64-column accumulator per query tile) across the key loop.
// CTA = 10 warps: 0-7 softmax-backward math + epilogues (thread = row; the two warps of a lane quarter split the key columns),
// 8 = TMA producer / TMEM allocator (loads run one buffer ahead of the MMAs), 9 = MMA issuer.
#include <cfloat>

#include "../../include/b200fm.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "tmap.cuh"

namespace b200fm {

// smem: Q/dO buffers (double-buffered across work items when NQT == 1), K/V double-buffered across key tiles / items,
// P and dS tiles.  NQT=1: 2*32 + 2*32 + 64 = 192 KB;  NQT=2: 64 + 2*32 + 64 = 192 KB.
template <int NQT>
struct AttnBwdSmem {
    static constexpr int kQBufs = NQT == 1 ? 2 : 1;
    static constexpr int kQBufBytes = NQT * 2 * 16384;                  // Q tiles then dO tiles of one item
    static constexpr int kQ = 0;
    static constexpr int kKV = kQBufs * kQBufBytes;                     // 2 x (K 16 KB + V 16 KB)
    static constexpr int kP = kKV + 2 * 32768;                          // 32 KB  [128 q][128 keys] bf16, two 64-key swizzle atoms
    static constexpr int kDS = kP + 32768;                              // 32 KB
    static constexpr int kBar = kDS + 32768;
    static constexpr int kStage = kBar + 256;                           // math warps x 2 KB store staging
    static constexpr int kTotal8 = kStage + 8 * 2048 + 1024;
    static constexpr int kTotal16 = kStage + 16 * 2048 + 1024;
};
constexpr int kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;
// MW math warps (8 or 16: softmax-backward math + epilogues; lane quarter = warp & 3, key-column group = warp >> 2), then one TMA producer
// / TMEM allocator warp and one MMA issuer warp.  The math warps are busy ~90 % of an item at low IPC (dependent TMEM-load -> exp2 -> STS
// chains, ncu source page r2_attn2): MW = 16 halves the work per warp and doubles the warps the schedulers can switch between.

struct AttnBwdArgs {
    const uint8_t* mask;
    long long mask_b_stride, mask_q_stride;
    const __nv_bfloat16* out;  long long ldo;
    const __nv_bfloat16* dout; long long lddo;
    const float* stats;
    const float* dsum;         // [B, H, Nq] fp32: D = rowsum(dO o O), produced by attn_bwd_prep_kernel
    __nv_bfloat16* dq; long long lddq;
    __nv_bfloat16* dk; long long lddk;
    __nv_bfloat16* dv; long long lddv;
    int B, H, Nq, Nk, nkt, num_items;
    float scale, scale_log2;
};

B200FM_DEVINL void pack32(const uint32_t (&a)[32], uint32_t (&p)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = pack_bf16x2(__uint_as_float(a[2 * j]), __uint_as_float(a[2 * j + 1]));
}

// D[b, h, q] = sum_d dO[b, q, h, d] * O[b, q, h, d]: 8 lanes per (row, head), 16 B per lane, fully coalesced.
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ out, long long ldo, const __nv_bfloat16* __restrict__ dout, long long lddo,
                     float* __restrict__ dsum, int B, int H, int Nq) {
    pdl_enter();
    const long long total = static_cast<long long>(B) * Nq * H;
    const int sub = threadIdx.x & 7;
    const long long stride = (gridDim.x * 256ll) >> 3;
    constexpr int U = 4;                                   // groups per thread in flight: 8 independent 16-byte loads
    // the loop condition is warp-uniform (first group of the warp): the full-mask shuffles below need all 32 lanes in the loop
    for (long long gw = (blockIdx.x * 256ll + (threadIdx.x & ~31)) >> 3; gw < total; gw += U * stride) {
        const long long g0 = gw + ((threadIdx.x & 31) >> 3);
        uint4 a[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long g = g0 + u * stride;
            if (g < total) {
                const int h = static_cast<int>(g % H);
                const long long row = g / H;                   // b * Nq + q
                a[u] = __ldg(reinterpret_cast<const uint4*>(out + row * ldo + h * 64) + sub);
                d[u] = __ldg(reinterpret_cast<const uint4*>(dout + row * lddo + h * 64) + sub);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long g = g0 + u * stride;
            const bool ok = g < total;                         // uniform across the 8 lanes of a group
            const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, dw[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
            float acc = 0.f;
            if (ok) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(dw[e]);
                    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (ok && sub == 0) {
                const int h = static_cast<int>(g % H);
                const long long row = g / H, b = row / Nq, q = row % Nq;
                dsum[(b * H + h) * Nq + q] = acc;
            }
        }
    }
}

template <int NQT, int MW>
__global__ void __launch_bounds__((MW + 2) * 32, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                     const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                     const AttnBwdArgs args) {
    using SM = AttnBwdSmem<NQT>;
    constexpr int QB = SM::kQBufs;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBar);
    uint64_t* full_q = bars + 0;    // [2]
    uint64_t* free_q = bars + 2;    // [2]
    uint64_t* full_kv = bars + 4;   // [2]
    uint64_t* free_kv = bars + 6;   // [2]
    uint64_t* sdp_full = bars + 8;  uint64_t* pds_full = bars + 9;
    uint64_t* dkv_full = bars + 10; uint64_t* dkv_free = bars + 11;
    uint64_t* dq_full = bars + 12;  uint64_t* dq_free = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    constexpr int kBwdMathWarps = MW;
    constexpr int CPW = 16 / MW;                 // 32-key chunks of a 128-key tile per math warp: 4 chunks over MW / 4 column groups (2 or 1)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkt = args.nkt;

    if (warp == kBwdMathWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
            for (int i = 0; i < 2; ++i) { mbar_init(&full_q[i], 1); mbar_init(&free_q[i], 1); mbar_init(&full_kv[i], 1); mbar_init(&free_kv[i], 1); }
            mbar_init(sdp_full, 1); mbar_init(pds_full, kBwdMathWarps * 32); mbar_init(dkv_full, 1); mbar_init(dkv_free, kBwdMathWarps);
            mbar_init(dq_full, 1); mbar_init(dq_free, kBwdMathWarps);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // TMEM is held: a dependent grid can no longer starve this one of columns
    pdl_wait();         // everything below touches global memory

    if (warp == kBwdMathWarps) {
        // ------------------------------ TMA producer: runs ahead of the MMAs by one buffer ------------------------------
        if (lane == 0) {
            uint32_t it = 0, kvc = 0;
            for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
                const int h = item % args.H, b = item / args.H;
                const uint32_t qb = it % QB, qn = it / QB;                  // buffer and its use count
                mbar_wait(&free_q[qb], (qn & 1) ^ 1);
                mbar_arrive_expect_tx(&full_q[qb], NQT * 2 * 16384);
                uint8_t* sq = smem + SM::kQ + qb * SM::kQBufBytes;
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) {
                    tma_load_3d(sq + qt * 16384, &tmap_q, &full_q[qb], h * 64, qt * 128, b);
                    tma_load_3d(sq + (NQT + qt) * 16384, &tmap_do, &full_q[qb], h * 64, qt * 128, b);
                }
                for (int kt = 0; kt < nkt; ++kt, ++kvc) {
                    const uint32_t kb = kvc & 1, kn = kvc >> 1;
                    mbar_wait(&free_kv[kb], (kn & 1) ^ 1);
                    mbar_arrive_expect_tx(&full_kv[kb], 2 * 16384);
                    tma_load_3d(smem + SM::kKV + kb * 32768, &tmap_k, &full_kv[kb], h * 64, kt * 128, b);
                    tma_load_3d(smem + SM::kKV + kb * 32768 + 16384, &tmap_v, &full_kv[kb], h * 64, kt * 128, b);
                }
            }
        }
    } else if (warp == kBwdMathWarps + 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            constexpr uint32_t kIdS = make_idesc_bf16(128, 128, false, false);    // S, dP : K-major x K-major
            constexpr uint32_t kIdT = make_idesc_bf16(128, 64, true, true);       // dV, dK: MN-major A (P^T / dS^T), MN-major B
            constexpr uint32_t kIdQ = make_idesc_bf16(128, 64, false, true);      // dQ    : K-major A (dS), MN-major B (K)
            const uint32_t sP = smem_u32(smem + SM::kP), sDS = smem_u32(smem + SM::kDS);
            uint32_t it = 0, kvc = 0, stepc = 0;
            for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
                const uint32_t qb = it % QB, qn = it / QB;
                const uint32_t sQbase = smem_u32(smem + SM::kQ + qb * SM::kQBufBytes);
                mbar_wait(&full_q[qb], qn & 1);
                for (int kt = 0; kt < nkt; ++kt, ++kvc) {
                    const uint32_t kb = kvc & 1, kn = kvc >> 1;
                    const uint32_t sK = smem_u32(smem + SM::kKV + kb * 32768), sV = sK + 16384;
                    mbar_wait(&full_kv[kb], kn & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int qt = 0; qt < NQT; ++qt, ++stepc) {
                        const uint32_t sQ = sQbase + qt * 16384, sDO = sQbase + (NQT + qt) * 16384;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + kColS, make_smem_desc(sQ + k * 32, 16, 1024), make_smem_desc(sK + k * 32, 16, 1024), kIdS, k != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + kColDP, make_smem_desc(sDO + k * 32, 16, 1024), make_smem_desc(sV + k * 32, 16, 1024), kIdS, k != 0);
                        umma_commit(sdp_full);
                        mbar_wait(pds_full, stepc & 1);
                        if (qt == 0) mbar_wait(dkv_free, (kvc & 1) ^ 1);
                        if (kt == 0 && qt == 0) mbar_wait(dq_free, (it & 1) ^ 1);
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {      // K dimension = 128 query rows, 16 per step
                            umma_bf16(tmem_base + kColDV, make_smem_desc(sP + kk * 2048, 16384, 1024), make_smem_desc(sDO + kk * 2048, 16384, 1024), kIdT, (qt | kk) != 0);
                            umma_bf16(tmem_base + kColDK, make_smem_desc(sDS + kk * 2048, 16384, 1024), make_smem_desc(sQ + kk * 2048, 16384, 1024), kIdT, (qt | kk) != 0);
                        }
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)        // K dimension = 128 keys
                            umma_bf16(tmem_base + kColDQ + qt * 64, make_smem_desc(sDS + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                                      make_smem_desc(sK + kk * 2048, 16384, 1024), kIdQ, (kt | kk) != 0);
                    }
                    umma_commit(dkv_full);
                    umma_commit(&free_kv[kb]);
                }
                umma_commit(dq_full);
                umma_commit(&free_q[qb]);
            }
        }
    } else {
        // ------------------------------ math + epilogue warps 0..7 ------------------------------
        const int quarter = warp & 3, cg = warp >> 2;          // cg: column group (CPW chunks of 32 keys)
        const int r = quarter * 32 + lane;                                  // query (or key) row inside the tile == TMEM lane
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        uint8_t* sP = smem + SM::kP;
        uint8_t* sDS = smem + SM::kDS;
        uint32_t* stg = reinterpret_cast<uint32_t*>(smem + SM::kStage) + warp * 512;
        uint32_t it = 0, kvc = 0, stepc = 0;
        // per query row: saved statistics and D = rowsum(dO o O).  They are requested one item AHEAD (during the previous item's dQ
        // epilogue): in round 1 the math warps spent ~40 % of their stall samples on exactly these loads (ncu source page, FSETP on m).
        float m_n[NQT], inv_n[NQT], D_n[NQT];
        auto prefetch_rows = [&](int item_) {
            const int h_ = item_ % args.H, b_ = item_ / args.H;
#pragma unroll
            for (int qt = 0; qt < NQT; ++qt) {
                const int qrow = qt * 128 + r;
                m_n[qt] = 0.f; inv_n[qt] = 0.f; D_n[qt] = 0.f;
                if (qrow < args.Nq) {
                    const long long si = (static_cast<long long>(b_) * args.H + h_) * args.Nq + qrow;
                    const float2 st = __ldg(reinterpret_cast<const float2*>(args.stats) + si);
                    m_n[qt] = st.x; inv_n[qt] = st.y;
                    D_n[qt] = __ldg(args.dsum + si);
                }
            }
        };
        if (static_cast<int>(blockIdx.x) < args.num_items) prefetch_rows(blockIdx.x);
        for (int item = blockIdx.x; item < args.num_items; item += gridDim.x, ++it) {
            const int h = item % args.H, b = item / args.H;
            float m_[NQT], inv_[NQT], D_[NQT];
#pragma unroll
            for (int qt = 0; qt < NQT; ++qt) { m_[qt] = m_n[qt]; inv_[qt] = inv_n[qt]; D_[qt] = D_n[qt]; }
            for (int kt = 0; kt < nkt; ++kt, ++kvc) {
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt, ++stepc) {
                    const int qrow = qt * 128 + r;
                    const bool row_ok = qrow < args.Nq;
                    const uint8_t* mrow = args.mask ? args.mask + b * args.mask_b_stride + (row_ok ? qrow : 0) * args.mask_q_stride : nullptr;
                    // p = exp2(s*scale_log2 - m) * inv = exp2(fma(s, scale_log2, log2(inv) - m)); rows past Nq get exp2(-inf) = 0.
                    // A masked key has p = exp2(kMaskedScore - m) * inv: inv on a fully masked row (m == kMaskedScore), else 0.
                    const float m = m_[qt], inv = inv_[qt];
                    const float shift = row_ok ? (__log2f(inv) - m) : -INFINITY;
                    const float p_masked = (row_ok && m == kMaskedScore) ? inv : 0.f;
                    const float neg_ds = -D_[qt] * args.scale;                       // ds = p * (dp*scale - D*scale)
                    uint32_t mb[CPW], tail[CPW];
                    uint4 mra[CPW], mrb[CPW];
                    bool mfast[CPW];
#pragma unroll
                    for (int cc = 0; cc < CPW; ++cc) {
                        const int col0 = kt * 128 + (cg * CPW + cc) * 32;
                        mfast[cc] = attn_mask_issue32(mrow, col0, args.Nk, mra[cc], mrb[cc], mb[cc]);   // loads in flight across the wait below
                        const int valid = args.Nk - col0;
                        tail[cc] = valid >= 32 ? 0u : (valid <= 0 ? 0xffffffffu : (0xffffffffu << valid));
                    }
                    mbar_wait(sdp_full, stepc & 1);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < CPW; ++cc)
                        if (mfast[cc]) mb[cc] = attn_mask_bits_from_raw(mra[cc], mrb[cc]);
#pragma unroll
                    for (int cc = 0; cc < CPW; ++cc) {
                        const int c = cg * CPW + cc;                          // this warp's 32-key chunk of the 128-key tile
                        const int col0 = kt * 128 + c * 32;
                        uint32_t pk[16], dk_[16];
                        if (col0 < args.Nk) {
                            uint32_t ss[32], dd[32];
                            tmem_ld_x32(t_lane + kColS + c * 32, ss);
                            tmem_ld_x32(t_lane + kColDP + c * 32, dd);
                            tmem_ld_wait();
                            const uint32_t mbits = mb[cc], tbits = tail[cc];
                            if (!__any_sync(0xffffffffu, (mbits | tbits) != 0u)) {
#pragma unroll
                                for (int j = 0; j < 32; j += 2) {
                                    const float p0 = fast_exp2(fmaf(__uint_as_float(ss[j]), args.scale_log2, shift));
                                    const float p1 = fast_exp2(fmaf(__uint_as_float(ss[j + 1]), args.scale_log2, shift));
                                    pk[j >> 1] = pack_bf16x2(p0, p1);
                                    dk_[j >> 1] = pack_bf16x2(p0 * fmaf(__uint_as_float(dd[j]), args.scale, neg_ds),
                                                             p1 * fmaf(__uint_as_float(dd[j + 1]), args.scale, neg_ds));
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; j += 2) {
                                    float p[2], ds[2];
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        float pe = fast_exp2(fmaf(__uint_as_float(ss[j + e]), args.scale_log2, shift));
                                        float de = pe * fmaf(__uint_as_float(dd[j + e]), args.scale, neg_ds);
                                        if ((mbits >> (j + e)) & 1u) { pe = p_masked; de = 0.f; }       // masked_fill: no gradient to the score
                                        if ((tbits >> (j + e)) & 1u) { pe = 0.f; de = 0.f; }
                                        p[e] = pe;
                                        ds[e] = de;
                                    }
                                    pk[j >> 1] = pack_bf16x2(p[0], p[1]);
                                    dk_[j >> 1] = pack_bf16x2(ds[0], ds[1]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) { pk[j] = 0u; dk_[j] = 0u; }
                        }
                        const uint32_t aoff = (c >> 1) * 16384;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t off = aoff + swz128(r, (c & 1) * 4 + q);
                            *reinterpret_cast<uint4*>(sP + off) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                            *reinterpret_cast<uint4*>(sDS + off) = make_uint4(dk_[4 * q], dk_[4 * q + 1], dk_[4 * q + 2], dk_[4 * q + 3]);
                        }
                    }
                    fence_proxy_async_smem();
                    tc_fence_before();
                    mbar_arrive(pds_full);
                }
                // dV, dK of this key tile are complete.  MW = 8: column group 0 writes dV, 1 writes dK (64 columns = 128 B per row each);
                // MW = 16: groups 0, 1 write the two 32-column halves of dV, groups 2, 3 those of dK.
                mbar_wait(dkv_full, kvc & 1);
                tc_fence_after();
                {
                    constexpr int EC = MW == 8 ? 2 : 1;                        // 32-column pieces per warp
                    const bool is_dv = MW == 8 ? cg == 0 : cg < 2;
                    const int piece0 = MW == 8 ? 0 : (cg & 1);
                    uint32_t a[EC][32];
                    const uint32_t col = (is_dv ? kColDV : kColDK) + piece0 * 32;
#pragma unroll
                    for (int e = 0; e < EC; ++e) tmem_ld_x32(t_lane + col + e * 32, a[e]);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dkv_free);
                    __nv_bfloat16* base = (is_dv ? args.dv + static_cast<long long>(b) * args.Nk * args.lddv
                                                 : args.dk + static_cast<long long>(b) * args.Nk * args.lddk) + h * 64 + piece0 * 32;
                    const long long ld = is_dv ? args.lddv : args.lddk;
#pragma unroll
                    for (int e = 0; e < EC; ++e) {
                        uint32_t pk[16];
                        pack32(a[e], pk);
                        attn_stage_store32(stg, lane, pk, base + e * 32, ld, kt * 128 + quarter * 32, args.Nk);
                    }
                }
            }
            if (item + static_cast<int>(gridDim.x) < args.num_items) prefetch_rows(item + gridDim.x);   // hidden behind the dQ epilogue
            mbar_wait(dq_full, it & 1);
            tc_fence_after();
            if (cg < 2) {                                                    // two column groups write the 2 x 32 dQ columns
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) {
                    uint32_t a0[32];
                    tmem_ld_x32(t_lane + kColDQ + qt * 64 + cg * 32, a0);
                    tmem_ld_wait();
                    uint32_t pk[16];
                    pack32(a0, pk);
                    attn_stage_store32(stg, lane, pk, args.dq + static_cast<long long>(b) * args.Nq * args.lddq + h * 64 + cg * 32, args.lddq,
                                       qt * 128 + quarter * 32, args.Nq);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_free);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kBwdMathWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int NQT, int MW>
static int launch_attn_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tk, const CUtensorMap& tv,
                           const AttnBwdArgs& a, cudaStream_t stream) {
    auto kern = attention_bwd_kernel<NQT, MW>;
    constexpr int smem = MW == 8 ? AttnBwdSmem<NQT>::kTotal8 : AttnBwdSmem<NQT>::kTotal16;
    static bool configured = false;
    if (!configured) {
        B200FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int sms = usable_sm_count();
    const int grid = a.num_items < sms ? a.num_items : sms;
    B200FM_LAUNCH(kern, dim3(grid), dim3((MW + 2) * 32), smem, stream, 1, tq, tdo, tk, tv, a);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_attention_bwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                    const uint8_t* mask, long long mask_b_stride, long long mask_q_stride, const void* out,
                                    long long ldo, const void* dout, long long lddo, const float* stats, float* dsum_ws, void* dq,
                                    long long lddq, void* dk, long long lddk, void* dv, long long lddv, int B, int H, int Nq, int Nk,
                                    float scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0 || H == 0 || Nq == 0) return 0;
    B200FM_CHECK(q && k && v && out && dout && stats && dq && dk && dv, "attention_bwd: null pointer");
    B200FM_CHECK(Nk >= 1 && Nk <= 256 && Nq <= 256, "attention_bwd: (Nq=%d, Nk=%d) outside the supported range (<= 256)", Nq, Nk);
    B200FM_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
                 "attention_bwd: row strides must be multiples of 8 elements");
    CUtensorMap tq, tdo, tk, tv;
    int rc;
    if ((rc = make_tmap_3d(&tq, q, TmapDtype::BF16, (uint64_t)H * 64, Nq, B, (uint64_t)ldq * 2, (uint64_t)Nq * ldq * 2, 64, 128, true))) return rc;
    if ((rc = make_tmap_3d(&tdo, dout, TmapDtype::BF16, (uint64_t)H * 64, Nq, B, (uint64_t)lddo * 2, (uint64_t)Nq * lddo * 2, 64, 128, true))) return rc;
    if ((rc = make_tmap_3d(&tk, k, TmapDtype::BF16, (uint64_t)H * 64, Nk, B, (uint64_t)ldk * 2, (uint64_t)Nk * ldk * 2, 64, 128, true))) return rc;
    if ((rc = make_tmap_3d(&tv, v, TmapDtype::BF16, (uint64_t)H * 64, Nk, B, (uint64_t)ldv * 2, (uint64_t)Nk * ldv * 2, 64, 128, true))) return rc;
    AttnBwdArgs a;
    a.mask = mask; a.mask_b_stride = mask_b_stride; a.mask_q_stride = mask_q_stride;
    a.out = reinterpret_cast<const __nv_bfloat16*>(out); a.ldo = ldo;
    a.dout = reinterpret_cast<const __nv_bfloat16*>(dout); a.lddo = lddo;
    a.stats = stats;
    a.dq = reinterpret_cast<__nv_bfloat16*>(dq); a.lddq = lddq;
    a.dk = reinterpret_cast<__nv_bfloat16*>(dk); a.lddk = lddk;
    a.dv = reinterpret_cast<__nv_bfloat16*>(dv); a.lddv = lddv;
    a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.nkt = (Nk + 127) / 128; a.num_items = B * H;
    a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
    B200FM_CHECK(dsum_ws != nullptr, "attention_bwd: dsum_ws (fp32 [B, H, Nq] workspace) is required");
    float* dsum = dsum_ws;
    a.dsum = dsum;
    {
        const long long groups = (long long)B * Nq * H;
        long long blocks = (groups * 8 + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        B200FM_LAUNCH(attn_bwd_prep_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, 1, a.out, ldo, a.dout, lddo, dsum, B, H, Nq);
    }
    if (option(kOptAttnBwdWarps) == 16)
        return Nq <= 128 ? launch_attn_bwd<1, 16>(tq, tdo, tk, tv, a, stream) : launch_attn_bwd<2, 16>(tq, tdo, tk, tv, a, stream);
    return Nq <= 128 ? launch_attn_bwd<1, 8>(tq, tdo, tk, tv, a, stream) : launch_attn_bwd<2, 8>(tq, tdo, tk, tv, a, stream);
}
