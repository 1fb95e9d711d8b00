// Shared device-side building blocks for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers (inline PTX),
// UMMA descriptor construction, small math helpers.  No CUTLASS dependency; descriptor bit layouts follow the PTX ISA
// (and were cross-checked against cute/arch/mma_sm100_desc.hpp's field tables).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200fm {

#define B200FM_DEVINL __device__ __forceinline__

// ------------------------------------------------------------------------------------------------------------
// programmatic dependent launch: every kernel of the library is launched with programmatic stream serialisation, lets
// the next kernel in the stream start its prologue (pdl_trigger, first instruction) and orders ALL of its own global
// memory traffic after the previous grid (pdl_wait: every CTA executes it, before its first global access).  A kernel
// that skipped pdl_wait could finish before its predecessor and break the chain for the kernel after it.
// ------------------------------------------------------------------------------------------------------------
B200FM_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
B200FM_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
B200FM_DEVINL void pdl_enter() { pdl_trigger(); pdl_wait(); }
enum { kOptPdl = 0, kOptGemmCtaPairs = 1, kOptLnBwdV2 = 2, kOptSmReserve = 3, kOptGemv = 4, kOptGemvPrefetch = 5, kOptGemmTmaStore = 6, kOptGemmDebug = 7, kOptCommSlim = 8, kOptAttnBwdWarps = 9, kOptCount = 10 };
int option(int id);      // runtime.cu: value of a runtime option (env default, b200fm_set_option override)
bool pdl_enabled();      // option "pdl" (env B200FM_PDL, default 1)
int usable_sm_count();   // SM count of the current device minus option "sm_reserve" (rounded up to even), at least 16
#define B200FM_LAUNCH(...) (void)::b200fm::launch_pdl(__VA_ARGS__)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------------------
// error reporting for the C ABI (thread-local last-error string; entry points return cudaError_t-like ints)
// ------------------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define B200FM_CHECK(cond, ...)                                  \
    do {                                                         \
        if (!(cond)) {                                           \
            ::b200fm::set_last_error(__VA_ARGS__);               \
            return 1;                                            \
        }                                                        \
    } while (0)
#define B200FM_CUDA(expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            ::b200fm::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 2;                                                                                  \
        }                                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------------------
// generic helpers
// ------------------------------------------------------------------------------------------------------------
B200FM_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

B200FM_DEVINL uint32_t elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred;
}

B200FM_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
B200FM_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

B200FM_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
B200FM_DEVINL float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
B200FM_DEVINL float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------------------
// mbarrier (shared::cta, 64-bit objects)
// ------------------------------------------------------------------------------------------------------------
B200FM_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
B200FM_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
B200FM_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
B200FM_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
B200FM_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
B200FM_DEVINL uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a protocol bug traps (surfaces as a CUDA error on the host) instead of hanging the GPU box.
#ifndef B200FM_WATCHDOG_CYCLES
#define B200FM_WATCHDOG_CYCLES (4000000000ll)   // ~2 s at 1.9 GHz
#endif
B200FM_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > B200FM_WATCHDOG_CYCLES) {
            printf("b200fm: mbarrier watchdog fired (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, parity);
            __trap();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2D tiled maps
// ------------------------------------------------------------------------------------------------------------
B200FM_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// L2 cache-policy hints (createpolicy encodings used by CUTLASS' TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

B200FM_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c_inner, int32_t c_outer,
                               uint64_t hint = kEvictNormal) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "l"(hint)
        : "memory");
}
B200FM_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                               uint64_t hint = kEvictNormal) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
        : "memory");
}
B200FM_DEVINL void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c_inner, int32_t c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
                 : "memory");
}
B200FM_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
B200FM_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
B200FM_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------------------
// d/da, d/db of silu(a) * b given the incoming gradient g (fm_utils.py:143): s = sigmoid(a)
B200FM_DEVINL void swiglu_grad(float a, float b, float g, float& da, float& db) {
    const float s = 1.0f / (1.0f + __expf(-a));
    da = g * b * s * (1.0f + a * (1.0f - s));
    db = g * a * s;
}
B200FM_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
B200FM_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Whole-warp (.sync.aligned).  ncols: power of two in [32, 512].  Writes the TMEM base address to *smem_slot.
B200FM_DEVINL void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
B200FM_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs with fp32 accumulate.  One thread issues.
B200FM_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
B200FM_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x N consecutive 32-bit columns (thread i gets lane base+i).
B200FM_DEVINL void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
B200FM_DEVINL void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
B200FM_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: two SMs of one TPC cooperate on a 256-row MMA; the leader (cluster rank 0) issues ----
B200FM_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
B200FM_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
B200FM_DEVINL uint32_t mapa_u32(const void* local, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(local)), "r"(rank));
    return out;
}
// Used to hand a TMEM accumulator back to the leader's MMA thread: the tcgen05 fences order the TMEM reads, no global or
// shared data is published through this arrive, so the default (release.cta) form is enough -- a cluster-scope release
// would add a MEMBAR that waits for all of the epilogue's outstanding global stores.
B200FM_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier given by its shared::cluster address (the leader's barrier)
B200FM_DEVINL void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int32_t c_inner, int32_t c_outer,
                                   uint64_t hint = kEvictNormal) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer), "l"(hint)
        : "memory");
}
B200FM_DEVINL void tmem_alloc_2sm(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
B200FM_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
B200FM_DEVINL void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// arrive on the mbarrier at this smem offset in every CTA of `cta_mask` once the issuing thread's prior MMAs retired
B200FM_DEVINL void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// ---- shared-memory matrix descriptor (PTX "matrix descriptor" for tcgen05.mma) ----
// bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
// [46,48) version = 1 (Blackwell) | [49,52) base offset (0: tiles are 1024 B aligned) | [61,64) layout: 2 = SWIZZLE_128B.
//
// K-major operand tile, 128B swizzle:  rows (M or N) at a 128 B pitch, 64 bf16 of K per row; 8-row groups 1024 B apart
//   (SBO = 1024); LBO unused for swizzled K-major (set to 1 as CUTLASS does).  Advancing K by 16 elements inside the
//   swizzle atom = +32 B on the start address.
// MN-major operand tile, 128B swizzle: K rows at a 128 B pitch, 64 bf16 of M/N per row; 8-K-row groups 1024 B apart
//   (SBO = 1024); successive 64-wide M/N chunks LBO bytes apart (= rows_of_K_in_tile * 128).  Advancing K by 16 = +2048 B.
B200FM_DEVINL uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// ---- instruction descriptor for kind::f16, bf16 x bf16 -> fp32 ----
// [4,6) D format: 1 = F32 | [7,10) A format: 1 = BF16 | [10,13) B format: 1 = BF16 | [15] A major (0 = K, 1 = MN) |
// [16] B major | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

// Byte offset of logical (row, 16-byte chunk) inside a 128B-swizzled tile whose rows are 128 B apart and whose base is
// 1024 B aligned: chunk index is XORed with (row mod 8)  (Swizzle<3,4,3>).
B200FM_DEVINL uint32_t swz128(uint32_t row, uint32_t chunk16) { return row * 128u + ((chunk16 ^ (row & 7u)) << 4); }

}  // namespace b200fm
