// Host-side construction of 2D TMA descriptors (cuTensorMapEncodeTiled fetched through the runtime's
// driver-entry-point query, so the library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200fm {

enum class TmapDtype { BF16, F32 };

// 2D row-major tensor [outer, inner] with `row_stride_bytes` between rows; box = [box_outer, box_inner]; 128B swizzle
// (box_inner * elem_size must be 128 B) or none.  Returns 0 on success (error text via set_last_error).
int make_tmap_2d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle128);

// same with an explicit swizzle width in bytes (0, 64 or 128; box_inner * elem_size must equal it when non-zero)
int make_tmap_2d_sw(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);

// 3D tensor [d2, d1, inner] (inner contiguous; stride1/stride2 in bytes); box = [1, box_d1, box_inner].  Used for
// [batch, tokens, features] activations so that rows past the end of one sample are zero-filled (never the next sample's).
int make_tmap_3d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t inner, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box_inner, uint32_t box_d1, bool swizzle128);

}  // namespace b200fm
