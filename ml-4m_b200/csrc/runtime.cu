// Host runtime glue for the C ABI: last-error string, device properties, TMA descriptor encoding.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/b200fm.h"
#include "common.cuh"
#include "tmap.cuh"

namespace b200fm {

// ---- runtime options: defaults from the environment (B200FM_<NAME>), changeable in-process through b200fm_set_option so that two
// variants can be measured back to back in ONE process on ONE box (box-to-box variance of the pool is ~3 %)
struct OptionSlot { const char* name; const char* env; int def; int value; bool init; };
static OptionSlot g_options[kOptCount] = {
    {"pdl", "B200FM_PDL", 1, 1, false},                        // programmatic dependent launch on every kernel
    {"gemm_cta_pairs", "B200FM_GEMM_CTA_PAIRS", 1, 1, false},  // tcgen05 cta_group::2 GEMM tiles
    {"ln_bwd_v2", "B200FM_LN_BWD_V2", 1, 1, false},            // LayerNorm backward: 1 = all loads of a row hoisted ahead of the reductions (35.4 us vs 47.4 us stand-alone at 16384 x 768: tools/ln_bench.py), 0 = plain
    {"sm_reserve", "B200FM_SM_RESERVE", 0, 0, false},          // SMs the persistent GEMM grids leave free (concurrent all-reduce kernel)
    {"gemv", "B200FM_GEMV", 1, 1, false},                      // NT GEMMs with <= 8 rows run on the weight-streaming kernel (gemv.cu)
    {"gemv_prefetch", "B200FM_GEMV_PREFETCH", 1, 1, false},    // gemv.cu: L2-prefetch the weight rows BEFORE waiting for the predecessor grid
    {"gemm_tma_store", "B200FM_GEMM_TMA_STORE", 1, 1, false},  // gemm.cu: bf16 outputs leave through TMA stores (2 passes over the smem / L1 data path instead of 3)
    {"gemm_debug", "B200FM_GEMM_DEBUG", 0, 0, false},          // MEASUREMENT ONLY (wrong results): 1 = GEMM epilogue stores nothing, 2 = epilogue skipped
    {"comm_slim", "B200FM_COMM_SLIM", 1, 1, false},            // comm.cu: all-reduce CTAs of 128 threads / 64 registers, co-resident with the persistent kernels (no SM reservation); 0 = few 512-thread CTAs on reserved SMs
    {"attn_bwd_warps", "B200FM_ATTN_BWD_WARPS", 8, 8, false},  // attention_bwd.cu: softmax-backward math warps per CTA (8 or 16)
};

int option(int id) {
    OptionSlot& o = g_options[id];
    if (!o.init) {
        const char* e = getenv(o.env);
        o.value = (e && e[0] != '\0') ? atoi(e) : o.def;
        o.init = true;
    }
    return o.value;
}

bool pdl_enabled() { return option(kOptPdl) != 0; }

int usable_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    // "sm_reserve": SMs left to a concurrently running gradient all-reduce kernel (comm.cu); even, so CTA-pair grids stay whole
    const int reserve = option(kOptSmReserve);
    const int n = sms - (reserve > 0 ? ((reserve + 1) & ~1) : 0);
    return n >= 16 ? n : 16;
}

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, bool swizzle128) {
    return make_tmap_2d_sw(out, base, dt, inner, outer, row_stride_bytes, box_inner, box_outer, swizzle128 ? 128 : 0);
}

int make_tmap_2d_sw(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                    uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
    const bool swizzle128 = swizzle_bytes == 128;
    EncodeTiledFn fn = get_encode_fn();
    B200FM_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    const uint32_t es = dt == TmapDtype::BF16 ? 2 : 4;
    B200FM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer %p not 16-byte aligned", base);
    B200FM_CHECK(row_stride_bytes % 16 == 0, "TMA row stride %llu B not a multiple of 16", (unsigned long long)row_stride_bytes);
    B200FM_CHECK(swizzle_bytes == 0 || swizzle_bytes == 64 || swizzle_bytes == 128, "TMA swizzle must be 0, 64 or 128 bytes");
    B200FM_CHECK(swizzle_bytes == 0 || box_inner * es == (uint32_t)swizzle_bytes, "%dB swizzle needs a %d-byte inner box (got %u)", swizzle_bytes, swizzle_bytes, box_inner * es);
    B200FM_CHECK(box_inner <= 256 && box_outer <= 256, "TMA box dims must be <= 256");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, dt == TmapDtype::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200FM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu box=%ux%u)",
                 (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
                 box_outer);
    return 0;
}

int make_tmap_3d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t inner, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box_inner, uint32_t box_d1, bool swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    B200FM_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    const uint32_t es = dt == TmapDtype::BF16 ? 2 : 4;
    B200FM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer %p not 16-byte aligned", base);
    B200FM_CHECK(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "TMA strides (%llu, %llu) B not multiples of 16",
                 (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes);
    B200FM_CHECK(!swizzle128 || box_inner * es == 128, "128B swizzle needs a 128-byte inner box (got %u)", box_inner * es);
    cuuint64_t dims[3] = {inner, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box_inner, box_d1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, dt == TmapDtype::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200FM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
    return 0;
}

}  // namespace b200fm

extern "C" {

const char* b200fm_last_error(void) { return b200fm::g_err; }

int b200fm_abi_version(void) { return B200FM_ABI_VERSION; }

int b200fm_set_option(const char* name, int value) {
    B200FM_CHECK(name != nullptr, "set_option: null name");
    for (int i = 0; i < b200fm::kOptCount; ++i)
        if (strcmp(b200fm::g_options[i].name, name) == 0) {
            b200fm::g_options[i].value = value;
            b200fm::g_options[i].init = true;
            return 0;
        }
    B200FM_CHECK(false, "set_option: unknown option '%s'", name);
    return 1;
}

int b200fm_get_option(const char* name, int* value) {
    B200FM_CHECK(name != nullptr && value != nullptr, "get_option: null pointer");
    for (int i = 0; i < b200fm::kOptCount; ++i)
        if (strcmp(b200fm::g_options[i].name, name) == 0) {
            *value = b200fm::option(i);
            return 0;
        }
    B200FM_CHECK(false, "get_option: unknown option '%s'", name);
    return 1;
}

int b200fm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin) {
    cudaDeviceProp p;
    B200FM_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
    return 0;
}

}  // extern "C"
