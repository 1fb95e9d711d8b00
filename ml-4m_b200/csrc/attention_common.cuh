// Helpers shared by the attention forward / backward kernels.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace b200fm {

// Finite stand-in for masked_fill(mask, -finfo.max) (fm_utils.py:169): large enough that exp2(masked - max) == 0 whenever
// the row has one unmasked key, equal for all masked keys so a fully masked row is uniform.  Applied AFTER the
// scale*log2e multiplication, so no overflow to -inf can occur.
constexpr float kMaskedScore = -3.0e38f;

// 2^x on the SFU (ex2.approx.ftz): exact 0 for x <= -150, 1 for x == 0 -- all the softmax needs.
B200FM_DEVINL float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

B200FM_DEVINL uint32_t nonzero_bytes_to_bits(uint32_t w) {
    // per byte: bit 7 of ((b & 0x7f) + 0x7f) | b is set iff b != 0; the multiply gathers the four flags into bits 24..27
    const uint32_t nz = ((w | ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu)) >> 7) & 0x01010101u;
    return (nz * 0x01020408u) >> 24;
}

// Bit j set <=> key (col0 + j) is masked for this query row; keys >= Nk report 0.  mrow may be nullptr (no mask).
B200FM_DEVINL uint32_t attn_mask_bits32(const uint8_t* mrow, int col0, int Nk) {
    if (mrow == nullptr) return 0u;
    const uint8_t* p = mrow + col0;
    uint32_t bits = 0u;
    if (col0 + 32 <= Nk && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        bits = nonzero_bytes_to_bits(a.x) | (nonzero_bytes_to_bits(a.y) << 4) | (nonzero_bytes_to_bits(a.z) << 8) |
               (nonzero_bytes_to_bits(a.w) << 12) | (nonzero_bytes_to_bits(b.x) << 16) | (nonzero_bytes_to_bits(b.y) << 20) |
               (nonzero_bytes_to_bits(b.z) << 24) | (nonzero_bytes_to_bits(b.w) << 28);
    } else {
        for (int j = 0; j < 32 && col0 + j < Nk; ++j) bits |= (__ldg(p + j) != 0 ? 1u : 0u) << j;
    }
    return bits;
}

// Split form of attn_mask_bits32 for software pipelining: `issue` starts the two 16-byte loads of an aligned, complete 32-key chunk and
// returns true (the words are consumed later by attn_mask_bits_from_raw, after the global-load latency has been hidden behind a barrier
// wait or the previous item's epilogue); otherwise it computes the bits right away (byte loads) into `bits` and returns false.
B200FM_DEVINL bool attn_mask_issue32(const uint8_t* mrow, int col0, int Nk, uint4& a, uint4& b, uint32_t& bits) {
    bits = 0u;
    if (mrow == nullptr) return false;
    const uint8_t* p = mrow + col0;
    if (col0 + 32 <= Nk && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        a = __ldg(reinterpret_cast<const uint4*>(p));
        b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        return true;
    }
    for (int j = 0; j < 32 && col0 + j < Nk; ++j) bits |= (__ldg(p + j) != 0 ? 1u : 0u) << j;
    return false;
}
B200FM_DEVINL uint32_t attn_mask_bits_from_raw(const uint4& a, const uint4& b) {
    return nonzero_bytes_to_bits(a.x) | (nonzero_bytes_to_bits(a.y) << 4) | (nonzero_bytes_to_bits(a.z) << 8) |
           (nonzero_bytes_to_bits(a.w) << 12) | (nonzero_bytes_to_bits(b.x) << 16) | (nonzero_bytes_to_bits(b.y) << 20) |
           (nonzero_bytes_to_bits(b.z) << 24) | (nonzero_bytes_to_bits(b.w) << 28);
}

// Warp-private smem staging for bf16 row tiles (same scheme as the GEMM epilogue): the thread = row register layout coming out
// of TMEM is turned into row-contiguous 64 B global segments.  stg: 32 rows x 16 packed words (2 KB), XOR-swizzled 16 B quads.
B200FM_DEVINL void attn_stage_store32(uint32_t* stg, int lane, const uint32_t (&p)[16], __nv_bfloat16* base, long long ld, int row0,
                                      int n_rows) {
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(stg + lane * 16 + 4 * (q ^ sw)) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = i * 8 + (lane >> 2), quad = lane & 3;
        const uint4 w = *reinterpret_cast<const uint4*>(stg + rl * 16 + 4 * (quad ^ ((rl >> 1) & 3)));
        if (row0 + rl < n_rows) *reinterpret_cast<uint4*>(base + static_cast<long long>(row0 + rl) * ld + quad * 8) = w;
    }
    __syncwarp();
}

}  // namespace b200fm
