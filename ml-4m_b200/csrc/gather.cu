// Modality-masked token selection and embedding gather/scatter (the "K5" path of SURVEY.md): replaces
//   cat_encoder_tensors + forward_mask_encoder   (fourm/models/fm.py:245-277, 338-390)
//   cat_decoder_tensors + forward_mask_decoder   (fourm/models/fm.py:279-336, 392-438)
//   adapt_decoder_attention_mask                 (fourm/models/fm.py:440-475)
//   the per-modality embedding forwards          (fourm/models/encoder_embeddings.py:87-121, 184-211, 280-309;
//                                                 fourm/models/decoder_embeddings.py:98-139, 226-255)
// The reference materialises [B, 2204, D] fp32 token and embedding tensors (x2), argsorts a float key and gathers; here a
// PLAN kernel does the stable partition with an integer prefix sum (bit-exact with argsort(mask + arange*1e-6), SURVEY.md
// v4) and an EMBED kernel writes only the kept [B, N, D] rows straight from the tables.  Index/mask outputs are integer
// exact; fp32 values are produced with the reference's operation order (emb = pos + mod; x0 = x + emb).
#include "../../include/b200fm.h"
#include "common.cuh"

namespace b200fm {

struct SegTable {
    b200fm_segment seg[B200FM_MAX_SEGMENTS];
    int offset[B200FM_MAX_SEGMENTS + 1];   // start of each segment in the concatenated position space
    int n_seg;
    int decoder;                           // 1: decoder side (teacher-forcing shift, mask token, targets); 0: encoder side
    int identity;                          // 1: keep every position in place (materialise a whole modality, generation path)
    int sum_mode;                          // 1: x0 = x + emb (training path); 0: x0 = x (embedding modules' own forward)
};

// effective number of positions a segment contributes: decoder sequences lose one (teacher-forcing shift, fm.py:312-319)
__host__ __device__ inline bool seg_rows_from_gemm(const b200fm_segment& s) { return s.kind == B200FM_KIND_IMG || s.kind == B200FM_KIND_SEQ_EMB; }
__host__ __device__ inline bool seg_seq_positions(const b200fm_segment& s) { return s.kind == B200FM_KIND_SEQ || s.kind == B200FM_KIND_SEQ_EMB; }
__host__ __device__ inline int seg_len(const b200fm_segment& s, int decoder) { return (decoder && s.kind == B200FM_KIND_SEQ) ? s.L - 1 : s.L; }

B200FM_DEVINL bool seg_masked(const b200fm_segment& s, int decoder, int b, int l) {
    const uint8_t* m = s.mask + (long long)b * s.L;
    if (decoder && s.kind == B200FM_KIND_SEQ) return (m[l + 1] | m[l]) != 0;      // fm.py:317
    return m[l] != 0;
}

// One CTA per sample.  Stable partition of the concatenated positions: valid (mask False) first, in order, then masked,
// in order; the first n_keep are kept (fm.py:364-367 / 410-413).
__global__ void __launch_bounds__(1024)
plan_kernel(const SegTable tab, int n_keep, int32_t* __restrict__ src_seg, int32_t* __restrict__ src_pos, int32_t* __restrict__ pos_id,
            uint8_t* __restrict__ pad_mask, int16_t* __restrict__ mod_mask, int16_t* __restrict__ mod_raw,
            int64_t* __restrict__ target_ids, int32_t* __restrict__ dam_out, const int32_t* __restrict__ order_dev) {
    pdl_enter();
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    // concatenation order: tab.seg[] as given, or -- order_dev != NULL -- tab.seg[order_dev[i]] is the i-th segment (the Python-random
    // decoder shuffle of fm.py:306 as DATA, so a captured CUDA graph can replay with a new order); src_seg always stores the index
    // into tab.seg[] so that the embedding kernels need no order.
    __shared__ int ord_s[B200FM_MAX_SEGMENTS];
    __shared__ int off_s[B200FM_MAX_SEGMENTS + 1];
    if (threadIdx.x == 0) {
        int off = 0;
        for (int i = 0; i < tab.n_seg; ++i) {
            const int sidx = order_dev ? order_dev[i] : i;
            ord_s[i] = sidx;
            off_s[i] = off;
            off += seg_len(tab.seg[sidx], tab.decoder);
        }
        for (int i = tab.n_seg; i <= B200FM_MAX_SEGMENTS; ++i) off_s[i] = off;
    }
    __syncthreads();
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = off_s[tab.n_seg];
    // pass A: number of valid positions in the whole row (needed to place the masked ones behind them)
    int cnt = 0;
    for (int p = tid; p < total; p += 1024) {
        int s = 0;
        while (p >= off_s[s + 1]) ++s;
        cnt += seg_masked(tab.seg[ord_s[s]], tab.decoder, b, p - off_s[s]) ? 0 : 1;
    }
    cnt = (int)warp_sum((float)cnt);   // counts <= 2^24: exact in fp32
    if (lane == 0) warp_tot[warp] = cnt;
    __syncthreads();
    int n_valid = 0;
    for (int w = 0; w < 32; ++w) n_valid += warp_tot[w];
    __syncthreads();
    if (tid == 0) carry_s = 0;
    __syncthreads();
    // pass B: chunked exclusive scan over positions, in order
    for (int base = 0; base < total; base += 1024) {
        const int p = base + tid;
        int s = 0, l = 0, valid = 0;
        if (p < total) {
            while (p >= off_s[s + 1]) ++s;
            l = p - off_s[s];
            s = ord_s[s];
            valid = seg_masked(tab.seg[s], tab.decoder, b, l) ? 0 : 1;
        }
        int incl = valid;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
        const int carry = carry_s;
        const int rank_valid = carry + wbase + incl - valid;          // exclusive rank among valid positions
        if (p < total) {
            const int slot = tab.identity ? p : (valid ? rank_valid : n_valid + (p - rank_valid));   // masked: rank among masked = p - #valid before p
            if (slot < n_keep) {
                const b200fm_segment& sg = tab.seg[s];
                const long long o = (long long)b * n_keep + slot;
                src_seg[o] = s;
                src_pos[o] = l;
                pad_mask[o] = (valid || tab.identity) ? 0 : 1;
                mod_raw[o] = (int16_t)sg.mod_id;
                mod_mask[o] = (valid || tab.identity) ? (int16_t)sg.mod_id : (int16_t)-1;   // fm.py:387 / 432
                // positional index: rank among the segment's valid RAW positions (encoder: input_mask, decoder: target_mask),
                // encoder_embeddings.py:110-112, decoder_embeddings.py:125-128; image modalities use the patch index.
                int pid = l;
                if (seg_seq_positions(sg)) {
                    const uint8_t* m = sg.mask + (long long)b * sg.L;
                    int rk = 0;
                    for (int i = 0; i <= l; ++i) rk += m[i] ? 0 : 1;
                    pid = m[l] ? -1 : rk - 1;                                    // -1: positional part zeroed (masked raw position)
                    if (sg.max_length > 0 && pid >= sg.max_length) pid = 0;       // decoder_embeddings.py:128 (encoder passes 0)
                }
                pos_id[o] = pid;
                if (tab.decoder) {
                    long long tgt = 0;
                    if (valid) {
                        const int li = (sg.kind == B200FM_KIND_SEQ) ? l + 1 : l;   // ids shifted left for sequences (fm.py:313)
                        tgt = sg.ids_is_i64 ? reinterpret_cast<const int64_t*>(sg.ids)[(long long)b * sg.L + li]
                                            : (long long)reinterpret_cast<const int32_t*>(sg.ids)[(long long)b * sg.L + li];
                    }
                    target_ids[o] = tgt;                                          // fm.py:430 target_ids[pad] = 0
                    dam_out[o] = sg.dam[(long long)b * sg.L + l];                  // gathered as is, also for pads (fm.py:425)
                }
            }
        }
        __syncthreads();
        if (tid == 1023) carry_s = carry + wbase + incl;
        __syncthreads();
    }
}

// adapt_decoder_attention_mask (fm.py:440-475): mask[b,i,j] = (j >= cumsum(dam)[b,i]) | (mod[b,i] != mod[b,j]) with the
// PRE-padding modality ids, or the causal triu(1).  One CTA per sample.
__global__ void __launch_bounds__(256)
decoder_mask_kernel(const int32_t* __restrict__ dam, const int16_t* __restrict__ mod_raw, uint8_t* __restrict__ out, int M,
                    int causal, int sep) {
    pdl_enter();
    extern __shared__ int sm_i[];
    int* cs = sm_i;                 // [M]
    int* md = sm_i + M;             // [M]
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < M; i += 256) md[i] = mod_raw[(long long)b * M + i];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < M; ++i) { acc += dam[(long long)b * M + i]; cs[i] = acc; }
    }
    __syncthreads();
    uint8_t* o = out + (long long)b * M * M;
    for (int idx = threadIdx.x; idx < M * M; idx += 256) {
        const int i = idx / M, j = idx % M;
        bool m = causal ? (j > i) : (j >= cs[i]);
        if (sep) m = m || (md[i] != md[j]);
        o[idx] = m ? 1 : 0;
    }
}

// One warp per kept row: x0 = x + emb, emb = pos (+) mod.  x comes from the token table, the projected pixel patches,
// or the mask token (decoder image modalities, fm.py:322).  Padded rows are zero (fm.py:385-386 / 428-429).
template <int VEC>
__global__ void __launch_bounds__(256)
embed_kernel(const SegTable tab, const int32_t* __restrict__ src_seg, const int32_t* __restrict__ src_pos,
             const int32_t* __restrict__ pos_id, const uint8_t* __restrict__ pad_mask, const float* __restrict__ mask_token,
             float* __restrict__ x0, float* __restrict__ emb_out, long long rows, int n_keep) {
    pdl_enter();
    constexpr int D = VEC * 128;
    const int lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * 8ll + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8ll) {
        float4* xo = reinterpret_cast<float4*>(x0 + row * D);
        float4* eo = emb_out ? reinterpret_cast<float4*>(emb_out + row * D) : nullptr;
        if (pad_mask[row]) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) { xo[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f); if (eo) eo[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f); }
            continue;
        }
        const b200fm_segment& sg = tab.seg[src_seg[row]];
        const int b = (int)(row / n_keep), l = src_pos[row], pid = pos_id[row];
        const float4* pe = pid >= 0 ? reinterpret_cast<const float4*>(sg.pos_emb + (long long)pid * D) : nullptr;
        const float4* me = reinterpret_cast<const float4*>(sg.mod_emb);
        const float4* xs = nullptr;
        const uint2* xb = nullptr;
        if (tab.decoder && sg.kind != B200FM_KIND_SEQ) {
            xs = reinterpret_cast<const float4*>(mask_token);
        } else if (seg_rows_from_gemm(sg)) {
            xb = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(sg.x_rows) + ((long long)b * sg.L + l) * D);
        } else {
            const long long id = sg.ids_is_i64 ? reinterpret_cast<const int64_t*>(sg.ids)[(long long)b * sg.L + l]
                                               : (long long)reinterpret_cast<const int32_t*>(sg.ids)[(long long)b * sg.L + l];
            xs = reinterpret_cast<const float4*>(sg.token_emb + id * D);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int c = i * 32 + lane;
            const float4 m4 = me[c];
            float4 e = pe ? pe[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            e.x += m4.x; e.y += m4.y; e.z += m4.z; e.w += m4.w;
            float4 x;
            if (xb) {
                const uint2 u = xb[c];
                const float2 a = unpack_bf16x2(u.x), d = unpack_bf16x2(u.y);
                x = make_float4(a.x, a.y, d.x, d.y);
            } else {
                x = xs[c];
            }
            xo[c] = tab.sum_mode ? make_float4(x.x + e.x, x.y + e.y, x.z + e.z, x.w + e.w) : x;
            if (eo) eo[c] = e;
        }
    }
}

// Backward scatter: token-table rows (vector fp32 atomics; padding_idx rows skipped like nn.Embedding), pixel-patch rows
// (bf16 copy for the patch-projection wgrad), learnable positional tables (d_pos_emb[pos_id] += dx0 + demb; the sincos
// buffers of the shipped configs pass NULL), one warp per kept row.
template <int VEC>
__global__ void __launch_bounds__(256)
embed_bwd_scatter_kernel(const SegTable tab, const int32_t* __restrict__ src_seg, const int32_t* __restrict__ src_pos,
                         const int32_t* __restrict__ pos_id, const uint8_t* __restrict__ pad_mask, const float* __restrict__ dx0,
                         const float* __restrict__ demb, long long rows, int n_keep) {
    pdl_enter();
    constexpr int D = VEC * 128;
    const int lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * 8ll + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8ll) {
        if (pad_mask[row]) continue;
        const b200fm_segment& sg = tab.seg[src_seg[row]];
        const int b = (int)(row / n_keep), l = src_pos[row];
        const float4* g = reinterpret_cast<const float4*>(dx0 + row * D);
        if (sg.d_pos_emb != nullptr && pos_id != nullptr) {
            const int pid = pos_id[row];
            if (pid >= 0) {
                float4* dst = reinterpret_cast<float4*>(sg.d_pos_emb + (long long)pid * D);
                const float4* ge = demb ? reinterpret_cast<const float4*>(demb + row * D) : nullptr;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    float4 v = tab.sum_mode ? g[i * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ge) { const float4 e = ge[i * 32 + lane]; v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w; }
                    atomicAdd(dst + i * 32 + lane, v);
                }
            }
        }
        if (tab.decoder && sg.kind != B200FM_KIND_SEQ) continue;              // mask-token rows: handled by the modality sums
        if (seg_rows_from_gemm(sg)) {
            if (sg.dx_rows == nullptr) continue;
            uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(sg.dx_rows) + ((long long)b * sg.L + l) * D);
#pragma unroll
            for (int i = 0; i < VEC; ++i) { const float4 v = g[i * 32 + lane]; dst[i * 32 + lane] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)); }
        } else {
            if (sg.d_token_emb == nullptr) continue;
            const long long id = sg.ids_is_i64 ? reinterpret_cast<const int64_t*>(sg.ids)[(long long)b * sg.L + l]
                                               : (long long)reinterpret_cast<const int32_t*>(sg.ids)[(long long)b * sg.L + l];
            if (id == sg.padding_idx) continue;
            float4* dst = reinterpret_cast<float4*>(sg.d_token_emb + id * D);
#pragma unroll
            for (int i = 0; i < VEC; ++i) atomicAdd(dst + i * 32 + lane, g[i * 32 + lane]);
        }
    }
}

// Per-modality column sums: d_mod_emb[s] += sum over kept rows of segment s of (dx0 + demb);  decoder image modalities
// additionally feed d_mask_token.  grid = (n_seg, D / 128, row_splits), block = 128 threads (one column each).
__global__ void __launch_bounds__(128)
embed_bwd_modsum_kernel(const SegTable tab, const int32_t* __restrict__ src_seg, const uint8_t* __restrict__ pad_mask,
                        const float* __restrict__ dx0, const float* __restrict__ demb, float* __restrict__ d_mask_token,
                        long long rows, int D) {
    pdl_enter();
    const int s = blockIdx.x, col = blockIdx.y * 128 + threadIdx.x;
    const long long per = (rows + gridDim.z - 1) / gridDim.z;
    const long long r0 = blockIdx.z * per, r1 = r0 + per < rows ? r0 + per : rows;
    float acc_mod = 0.f, acc_tok = 0.f;
    for (long long r = r0; r < r1; ++r) {
        if (pad_mask[r] || src_seg[r] != s) continue;
        const float g = dx0[r * D + col];
        acc_tok += g;
        acc_mod += (tab.sum_mode ? g : 0.f) + (demb ? demb[r * D + col] : 0.f);
    }
    const b200fm_segment& sg = tab.seg[s];
    if (sg.d_mod_emb != nullptr && acc_mod != 0.f) atomicAdd(sg.d_mod_emb + col, acc_mod);
    if (tab.decoder && sg.kind != B200FM_KIND_SEQ && d_mask_token != nullptr && acc_tok != 0.f) atomicAdd(d_mask_token + col, acc_tok);
}

// Row index lists per modality for the masked-token head: rows (b*M + n) with mod_mask == id, in row-major order
// (the order of y[decoder_mod_mask == idx], fm.py:591).  Single CTA; n_rows <= a few 10^4.
__global__ void __launch_bounds__(1024)
head_rows_kernel(const int16_t* __restrict__ mod_mask, long long n_rows, const int* __restrict__ mod_ids, int n_mods,
                 int32_t* __restrict__ rows_out, int32_t* __restrict__ counts) {
    pdl_enter();
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int mi = 0; mi < n_mods; ++mi) {
        const int id = mod_ids[mi];
        if (tid == 0) carry_s = 0;
        __syncthreads();
        for (long long base = 0; base < n_rows; base += 1024) {
            const long long p = base + tid;
            const int hit = (p < n_rows && mod_mask[p] == id) ? 1 : 0;
            int incl = hit;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
            const int carry = carry_s;
            if (hit) rows_out[(long long)mi * n_rows + carry + wbase + incl - 1] = (int32_t)p;
            __syncthreads();
            if (tid == 1023) carry_s = carry + wbase + incl;
            __syncthreads();
        }
        if (tid == 0) counts[mi] = carry_s;
        __syncthreads();
    }
}

// out[i] = src[rows[i]] : row gather used to feed the per-modality logits GEMMs (bf16 rows) and their targets (int64).
// n_dev != NULL: the row count lives on the device (n = min(*n_dev, n)); output rows [n, roundup(n, 128)) are ZERO-filled because a
// following weight-gradient GEMM contracts over whole 64-row blocks.
__global__ void gather_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ rows, __nv_bfloat16* __restrict__ out,
                                        long long n, int D8, const int* __restrict__ n_dev) {
    pdl_enter();
    long long fill = n;
    if (n_dev != nullptr) { const long long d = max(0, __ldg(n_dev)); fill = min((d + 127) & ~127ll, n); n = min(d, n); }
    const long long total = fill * D8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D8; const int c = (int)(i % D8);
        reinterpret_cast<uint4*>(out)[i] = r < n ? reinterpret_cast<const uint4*>(src)[(long long)rows[r] * D8 + c] : make_uint4(0u, 0u, 0u, 0u);
    }
}
__global__ void gather_i64_kernel(const int64_t* __restrict__ src, const int32_t* __restrict__ rows, int64_t* __restrict__ out, long long n,
                                  const int* __restrict__ n_dev) {
    pdl_enter();
    if (n_dev != nullptr) n = min((long long)max(0, __ldg(n_dev)), n);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = src[rows[i]];
}
// dst[rows[i]] += src[i] (fp32 accumulate of bf16 rows; each destination row is hit by at most one source row per call)
__global__ void scatter_add_rows_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ rows, float* __restrict__ dst,
                                        long long n, int D4) {
    pdl_enter();
    const long long total = n * D4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D4; const int c = (int)(i % D4);
        const uint2 u = reinterpret_cast<const uint2*>(src)[i];
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        float4* d = reinterpret_cast<float4*>(dst) + (long long)rows[r] * D4 + c;
        float4 v = *d;
        v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
        *d = v;
    }
}

// dst[rows[i]] = src[i] for bf16 rows (destination rows are distinct)
__global__ void scatter_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ rows, __nv_bfloat16* __restrict__ dst,
                                         long long n, int D8, const int* __restrict__ n_dev) {
    pdl_enter();
    if (n_dev != nullptr) n = min((long long)max(0, __ldg(n_dev)), n);
    const long long total = n * D8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / D8; const int c = (int)(i % D8);
        reinterpret_cast<uint4*>(dst)[(long long)rows[r] * D8 + c] = reinterpret_cast<const uint4*>(src)[i];
    }
}

static int fill_table(SegTable& t, const b200fm_segment* segs, int n_seg, int mode) {
    B200FM_CHECK(n_seg >= 1 && n_seg <= B200FM_MAX_SEGMENTS, "segment count %d outside [1, %d]", n_seg, B200FM_MAX_SEGMENTS);
    B200FM_CHECK(mode >= 0 && mode < 8, "bad mode flags %d", mode);
    const int decoder = (mode & B200FM_MODE_DECODER) ? 1 : 0;
    t.n_seg = n_seg; t.decoder = decoder; t.identity = (mode & B200FM_MODE_IDENTITY) ? 1 : 0; t.sum_mode = (mode & B200FM_MODE_NO_SUM) ? 0 : 1;
    int off = 0;
    for (int i = 0; i < n_seg; ++i) {
        t.seg[i] = segs[i];
        t.offset[i] = off;
        B200FM_CHECK(segs[i].mask != nullptr && segs[i].L > 0, "segment %d: missing mask or empty", i);
        off += seg_len(segs[i], decoder);
    }
    for (int i = n_seg; i <= B200FM_MAX_SEGMENTS; ++i) t.offset[i] = off;
    return 0;
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_select_plan(const b200fm_segment* segs, int n_seg, int mode, int B, int n_keep, int32_t* src_seg,
                                  int32_t* src_pos, int32_t* pos_id, uint8_t* pad_mask, int16_t* mod_mask, int16_t* mod_raw,
                                  int64_t* target_ids, int32_t* dam_out, void* stream_) {
    return b200fm_select_plan_ordered(segs, n_seg, mode, B, n_keep, src_seg, src_pos, pos_id, pad_mask, mod_mask, mod_raw, target_ids, dam_out,
                                      nullptr, stream_);
}

extern "C" int b200fm_select_plan_ordered(const b200fm_segment* segs, int n_seg, int mode, int B, int n_keep, int32_t* src_seg,
                                          int32_t* src_pos, int32_t* pos_id, uint8_t* pad_mask, int16_t* mod_mask, int16_t* mod_raw,
                                          int64_t* target_ids, int32_t* dam_out, const int32_t* order_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return 0;
    SegTable t;
    if (int rc = fill_table(t, segs, n_seg, mode)) return rc;
    const int decoder = t.decoder;
    B200FM_CHECK(n_keep >= 1 && n_keep <= t.offset[n_seg], "select_plan: n_keep=%d exceeds the %d available positions", n_keep, t.offset[n_seg]);
    B200FM_CHECK(src_seg && src_pos && pos_id && pad_mask && mod_mask && mod_raw, "select_plan: null output");
    if (decoder) {
        B200FM_CHECK(target_ids && dam_out, "select_plan: decoder side needs target_ids and dam outputs");
        for (int i = 0; i < n_seg; ++i) B200FM_CHECK(segs[i].ids && segs[i].dam, "select_plan: decoder segment %d needs ids and dam", i);
    }
    B200FM_LAUNCH(plan_kernel, dim3(B), dim3(1024), 0, stream, 1, t, n_keep, src_seg, src_pos, pos_id, pad_mask, mod_mask, mod_raw, target_ids, dam_out, order_dev);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_decoder_attention_mask(const int32_t* dam, const int16_t* mod_raw, uint8_t* mask_out, int B, int M, int causal,
                                             int sep, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0 || M == 0) return 0;
    B200FM_CHECK(dam && mod_raw && mask_out, "decoder_attention_mask: null pointer");
    B200FM_CHECK(M <= 4096, "decoder_attention_mask: M=%d too large", M);
    B200FM_LAUNCH(decoder_mask_kernel, dim3(B), dim3(256), 2 * M * sizeof(int), stream, 1, dam, mod_raw, mask_out, M, causal, sep);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

#define B200FM_VEC_SWITCH(D, CALL)                                                                     \
    switch ((D) / 128) {                                                                               \
        case 1: { constexpr int V = 1; CALL; } break;  case 2: { constexpr int V = 2; CALL; } break;   \
        case 3: { constexpr int V = 3; CALL; } break;  case 4: { constexpr int V = 4; CALL; } break;   \
        case 6: { constexpr int V = 6; CALL; } break;  case 8: { constexpr int V = 8; CALL; } break;   \
        case 10: { constexpr int V = 10; CALL; } break; case 12: { constexpr int V = 12; CALL; } break; \
        case 16: { constexpr int V = 16; CALL; } break;                                                \
        default: B200FM_CHECK(false, "embedding dim %d has no instantiation", (D));                    \
    }

extern "C" int b200fm_embed_rows(const b200fm_segment* segs, int n_seg, int mode, const int32_t* src_seg, const int32_t* src_pos,
                                 const int32_t* pos_id, const uint8_t* pad_mask, const float* mask_token, float* x0, float* emb_out,
                                 int B, int n_keep, int D, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return 0;
    SegTable t;
    if (int rc = fill_table(t, segs, n_seg, mode)) return rc;
    const int decoder = t.decoder;
    B200FM_CHECK(D % 128 == 0, "embed_rows: D=%d must be a multiple of 128", D);
    B200FM_CHECK(src_seg && src_pos && pos_id && pad_mask && x0, "embed_rows: null pointer");
    B200FM_CHECK(!decoder || mask_token, "embed_rows: decoder side needs the mask token");
    const long long rows = (long long)B * n_keep;
    const int grid = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
    B200FM_VEC_SWITCH(D, (B200FM_LAUNCH((embed_kernel<V>), dim3(grid), dim3(256), 0, stream, 1, t, src_seg, src_pos, pos_id, pad_mask, mask_token, x0, emb_out, rows, n_keep)));
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_embed_rows_bwd(const b200fm_segment* segs, int n_seg, int mode, const int32_t* src_seg, const int32_t* src_pos,
                                     const int32_t* pos_id, const uint8_t* pad_mask, const float* dx0, const float* demb, float* d_mask_token, int B, int n_keep,
                                     int D, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0) return 0;
    SegTable t;
    if (int rc = fill_table(t, segs, n_seg, mode)) return rc;
    B200FM_CHECK(D % 128 == 0, "embed_rows_bwd: D=%d must be a multiple of 128", D);
    B200FM_CHECK(src_seg && src_pos && pad_mask && dx0, "embed_rows_bwd: null pointer");
    const long long rows = (long long)B * n_keep;
    const int grid = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
    B200FM_VEC_SWITCH(D, (B200FM_LAUNCH((embed_bwd_scatter_kernel<V>), dim3(grid), dim3(256), 0, stream, 1, t, src_seg, src_pos, pos_id, pad_mask, dx0, demb, rows, n_keep)));
    B200FM_CUDA(cudaGetLastError());
    const int splits = (int)(rows / 512 > 0 ? (rows / 512 > 32 ? 32 : rows / 512) : 1);
    B200FM_LAUNCH(embed_bwd_modsum_kernel, dim3(dim3(n_seg, D / 128, splits)), dim3(128), 0, stream, 1, t, src_seg, pad_mask, dx0, demb, d_mask_token, rows, D);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_head_rows(const int16_t* mod_mask, long long n_rows, const int* mod_ids_dev, int n_mods, int32_t* rows_out,
                                int32_t* counts, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_rows == 0 || n_mods == 0) return 0;
    B200FM_CHECK(mod_mask && mod_ids_dev && rows_out && counts, "head_rows: null pointer");
    B200FM_LAUNCH(head_rows_kernel, dim3(1), dim3(1024), 0, stream, 1, mod_mask, n_rows, mod_ids_dev, n_mods, rows_out, counts);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_gather_rows_bf16(const void* src, const int32_t* rows, void* out, long long n, int D, void* stream_) {
    return b200fm_gather_rows_bf16_dyn(src, rows, out, n, D, nullptr, stream_);
}
extern "C" int b200fm_gather_rows_bf16_dyn(const void* src, const int32_t* rows, void* out, long long n, int D, const int* n_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(src && rows && out && D % 8 == 0, "gather_rows_bf16: bad arguments");
    const long long total = n * (D / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    B200FM_LAUNCH(gather_rows_bf16_kernel, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(src), rows, reinterpret_cast<__nv_bfloat16*>(out), n, D / 8, n_dev);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

// K/V cache append of the autoregressive decode loop: cache[b, *pos, col0 : col0 + width] = src[b, 0 : width] (bf16), the position read
// from device memory so the launch can sit in a CUDA graph.  One CTA per sequence; 16-byte moves when the geometry allows.
namespace b200fm {
__global__ void __launch_bounds__(256)
kv_append_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ cache, long long L, long long row_w,
                 const long long* __restrict__ pos_dev, int width, int col0, int vec) {
    pdl_enter();
    const long long pos = __ldg(pos_dev);
    if (pos < 0 || pos >= L) return;
    const __nv_bfloat16* s = src + blockIdx.x * ld_src;
    __nv_bfloat16* d = cache + (blockIdx.x * L + pos) * row_w + col0;
    if (vec) {
        for (int i = threadIdx.x; i < width / 8; i += blockDim.x) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
    } else {
        for (int i = threadIdx.x; i < width; i += blockDim.x) d[i] = s[i];
    }
}
}  // namespace b200fm

extern "C" int b200fm_kv_append(const void* src, long long ld_src, void* cache, long long L, long long row_w, const int64_t* pos_dev, int B,
                                int width, int col0, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (B == 0 || width == 0) return 0;
    B200FM_CHECK(src && cache && pos_dev && col0 >= 0 && col0 + width <= row_w, "kv_append: bad arguments");
    const int vec = (width % 8 == 0 && col0 % 8 == 0 && row_w % 8 == 0 && ld_src % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(cache) & 15) == 0) ? 1 : 0;
    B200FM_LAUNCH(kv_append_kernel, dim3(B), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(src), ld_src,
                  reinterpret_cast<__nv_bfloat16*>(cache), L, row_w, reinterpret_cast<const long long*>(pos_dev), width, col0, vec);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_gather_i64(const int64_t* src, const int32_t* rows, int64_t* out, long long n, void* stream_) {
    return b200fm_gather_i64_dyn(src, rows, out, n, nullptr, stream_);
}
extern "C" int b200fm_gather_i64_dyn(const int64_t* src, const int32_t* rows, int64_t* out, long long n, const int* n_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(src && rows && out, "gather_i64: null pointer");
    B200FM_LAUNCH(gather_i64_kernel, dim3((int)((n + 255) / 256)), dim3(256), 0, stream, 1, src, rows, out, n, n_dev);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_scatter_add_rows(const void* src_bf16, const int32_t* rows, float* dst, long long n, int D, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(src_bf16 && rows && dst && D % 4 == 0, "scatter_add_rows: bad arguments");
    const long long total = n * (D / 4);
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    B200FM_LAUNCH(scatter_add_rows_kernel, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(src_bf16), rows, dst, n, D / 4);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b200fm_scatter_rows_bf16(const void* src, const int32_t* rows, void* dst, long long n, int D, void* stream_) {
    return b200fm_scatter_rows_bf16_dyn(src, rows, dst, n, D, nullptr, stream_);
}
extern "C" int b200fm_scatter_rows_bf16_dyn(const void* src, const int32_t* rows, void* dst, long long n, int D, const int* n_dev, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    B200FM_CHECK(src && rows && dst && D % 8 == 0, "scatter_rows_bf16: bad arguments");
    const long long total = n * (D / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    B200FM_LAUNCH(scatter_rows_bf16_kernel, dim3(grid), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(src), rows, reinterpret_cast<__nv_bfloat16*>(dst), n, D / 8, n_dev);
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
