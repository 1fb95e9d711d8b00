// Gradient all-reduce of the data-parallel 4M train step over NVLink peer memory -- replaces the NCCL all-reduce that torch DDP
// issues for the reference (run_training_4m.py:512 `DistributedDataParallel(model, ...)`; gradients are averaged, fp32).
//
// Why not NCCL here: the block stack runs persistent 148-CTA tcgen05 GEMM grids, so an NCCL ring kernel (16-32 CTAs, scheduled
// whenever an SM frees up) either waits behind them or steals SMs from statically scheduled tiles; round 1 measured the 1.44 GB
// fp32 all-reduce as fully exposed (+4.3 ms at 2 GPUs, +6.5 ms at 8).  This kernel is a two-shot all-reduce (reduce-scatter by
// P2P loads, all-gather by P2P stores) in ONE launch on a FIXED, small number of CTAs; the compute kernels leave exactly that many
// SMs free while a reduction is in flight (runtime option "sm_reserve"), so neither side ever waits for the other's SMs.
//
//   every rank owns one arena (cudaMalloc, exported with cudaIpcGetMemHandle) holding all gradients at identical offsets, plus a
//   small flag block.  For the chunk [off, off+n):
//     phase 0  CTA c of rank r tells CTA c of every peer "my chunk is final" (flag store, release.sys) and waits for theirs
//     phase 1  rank r reduces the r-th 1/W of the chunk: v = sum_p peer[p][i] (16-byte loads over NVLink, 16 in flight per thread),
//              v *= 1/W, stores v into ALL W arenas (its own included): one owner per element -> bit-identical replicas
//     phase 2  fence.sys, flag to every peer "my stores have landed", wait for theirs -> the chunk is complete everywhere
//   flags carry a monotonically increasing sequence number, so they are never reset.
#include "../../include/b200fm.h"
#include <cstring>

#include "common.cuh"

namespace b200fm {

constexpr int kCommMaxWorld = 8;
constexpr int kCommMaxCtas = 256;
constexpr int kCommThreads = 512;
// flag block layout (uint32): [phase 0|1][cta][src rank]
constexpr int kCommFlagWords = 2 * kCommMaxCtas * kCommMaxWorld;

struct CommPeers {
    float* data[kCommMaxWorld];
    uint32_t* flags[kCommMaxWorld];
};

B200FM_DEVINL void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
B200FM_DEVINL uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer gradients are read exactly once: keep them out of L1 and do not let them displace the GEMM operands in L2
B200FM_DEVINL float4 ld_peer_f4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
B200FM_DEVINL void st_peer_f4(float4* p, const float4& v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// signal the same-numbered CTA of every rank, then wait until every rank's CTA has signalled this one
B200FM_DEVINL void cta_barrier_all_ranks(const CommPeers& pr, int rank, int world, int phase, uint32_t seq) {
    __syncthreads();
    const int slot = (phase * kCommMaxCtas + blockIdx.x) * kCommMaxWorld;
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(pr.flags[threadIdx.x] + slot + rank, seq);
        const uint32_t* mine = pr.flags[rank] + slot + threadIdx.x;
        // sequence numbers only grow; a peer that is already one reduction ahead has necessarily passed this one
        while (static_cast<int32_t>(ld_acquire_sys(mine) - seq) < 0) __nanosleep(64);
    }
    __syncthreads();
}

// Two launch shapes.  "wide": few CTAs of 512 threads with 16 loads in flight per thread -- they need SMs of their own (the persistent
// GEMM / attention grids leave `sm_reserve` SMs free for them).  "slim" (THREADS = 128, <= 64 registers, no shared memory): small enough
// to be CO-RESIDENT with a persistent kernel's CTA on the same SM (320 x 168 + 128 x 64 registers fit the file), so no SM has to be taken
// away from the backward pass; the bytes in flight come from the number of CTAs (32-64) instead.
template <int W, int THREADS, int UTOT>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 8 : 1)
allreduce_f32_kernel(const CommPeers pr, int rank, long long off, long long n4, float scale, uint32_t seq, const uint32_t* seq_base_dev) {
    // CUDA-graph replays: the sequence number must differ per replay, so its step-dependent part is read from device memory
    if (seq_base_dev != nullptr) seq += *reinterpret_cast<const volatile uint32_t*>(seq_base_dev);
    cta_barrier_all_ranks(pr, rank, W, 0, seq);
    // my shard of the chunk, in float4 units
    const long long per = (n4 + W - 1) / W;
    const long long lo = per * rank, hi = (lo + per < n4) ? lo + per : n4;
    const float4* src[W];
    float4* dst[W];
#pragma unroll
    for (int p = 0; p < W; ++p) {
        // start with the NEXT rank's arena so the W ranks do not all hit the same peer at the same time
        const int q = (rank + 1 + p) % W;
        src[p] = reinterpret_cast<const float4*>(pr.data[q] + off);
        dst[p] = reinterpret_cast<float4*>(pr.data[q] + off);
    }
    // NVLink round trips are ~2 us: the few CTAs this kernel is allowed need many 16-byte loads in flight per thread.
    // U float4 per peer per thread -> U * W independent loads before the first use (16 for every W).
    constexpr int U = UTOT / W >= 1 ? UTOT / W : 1;
    const long long stride = (long long)gridDim.x * THREADS;
    for (long long i0 = lo + (long long)blockIdx.x * THREADS + threadIdx.x; i0 < hi; i0 += U * stride) {
        float4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
#pragma unroll
                for (int p = 0; p < W; ++p) v[u][p] = ld_peer_f4(src[p] + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                // every element is summed by exactly ONE rank (its shard owner) and the result is stored to all ranks, so the
                // replicas end up bit-identical whatever the summation order is; the owner's order is fixed -> reproducible
                float4 s = v[u][0];
#pragma unroll
                for (int p = 1; p < W; ++p) { s.x += v[u][p].x; s.y += v[u][p].y; s.z += v[u][p].z; s.w += v[u][p].w; }
                s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
#pragma unroll
                for (int p = 0; p < W; ++p) st_peer_f4(dst[p] + i, s);
            }
        }
    }
    cta_barrier_all_ranks(pr, rank, W, 1, seq);
}

}  // namespace b200fm

using namespace b200fm;

extern "C" int b200fm_comm_flag_bytes(void) { return kCommFlagWords * (int)sizeof(uint32_t); }

extern "C" int b200fm_comm_alloc(long long bytes, void** ptr) {
    B200FM_CHECK(bytes > 0 && ptr != nullptr, "comm_alloc: bad arguments");
    B200FM_CUDA(cudaMalloc(ptr, (size_t)bytes));
    B200FM_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    return 0;
}

extern "C" int b200fm_comm_free(void* ptr) {
    if (ptr) B200FM_CUDA(cudaFree(ptr));
    return 0;
}

extern "C" int b200fm_comm_ipc_export(const void* ptr, void* handle64) {
    B200FM_CHECK(ptr && handle64, "comm_ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    B200FM_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr)));
    return 0;
}

extern "C" int b200fm_comm_ipc_open(const void* handle64, void** ptr) {
    B200FM_CHECK(ptr && handle64, "comm_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    B200FM_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int b200fm_comm_ipc_close(void* ptr) {
    if (ptr) B200FM_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int b200fm_allreduce_f32(void* const* peer_data, void* const* peer_flags, int rank, int world, long long offset_elems,
                                    long long n_elems, float scale, unsigned int seq, int n_ctas, void* stream_) {
    return b200fm_allreduce_f32_seq(peer_data, peer_flags, rank, world, offset_elems, n_elems, scale, seq, nullptr, n_ctas, stream_);
}

extern "C" int b200fm_allreduce_f32_seq(void* const* peer_data, void* const* peer_flags, int rank, int world, long long offset_elems,
                                        long long n_elems, float scale, unsigned int seq, const unsigned int* seq_base_dev, int n_ctas,
                                        void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    B200FM_CHECK(peer_data && peer_flags, "allreduce_f32: null pointer table");
    B200FM_CHECK(world >= 2 && world <= kCommMaxWorld && rank >= 0 && rank < world, "allreduce_f32: world=%d rank=%d (2..%d ranks)", world, rank, kCommMaxWorld);
    B200FM_CHECK(n_ctas >= 1 && n_ctas <= kCommMaxCtas, "allreduce_f32: n_ctas=%d outside [1, %d]", n_ctas, kCommMaxCtas);
    B200FM_CHECK(offset_elems % 4 == 0 && n_elems % 4 == 0 && n_elems > 0, "allreduce_f32: offset/length must be multiples of 4 floats");
    CommPeers pr;
    for (int p = 0; p < kCommMaxWorld; ++p) {
        pr.data[p] = p < world ? reinterpret_cast<float*>(peer_data[p]) : nullptr;
        pr.flags[p] = p < world ? reinterpret_cast<uint32_t*>(peer_flags[p]) : nullptr;
        B200FM_CHECK(p >= world || (pr.data[p] && pr.flags[p]), "allreduce_f32: null peer pointer %d", p);
    }
    const long long n4 = n_elems / 4;
    // plain launch (no programmatic dependent launch): the kernel must not start before the producers of the chunk have finished
    const bool slim = option(kOptCommSlim) != 0;
#define B200FM_AR_CASE(W_)                                                                                                   \
    case W_:                                                                                                                 \
        if (slim) allreduce_f32_kernel<W_, 128, 8><<<n_ctas, 128, 0, stream>>>(pr, rank, offset_elems, n4, scale, seq, seq_base_dev); \
        else allreduce_f32_kernel<W_, kCommThreads, 16><<<n_ctas, kCommThreads, 0, stream>>>(pr, rank, offset_elems, n4, scale, seq, seq_base_dev); \
        break;
    switch (world) {
        B200FM_AR_CASE(2) B200FM_AR_CASE(3) B200FM_AR_CASE(4) B200FM_AR_CASE(5) B200FM_AR_CASE(6) B200FM_AR_CASE(7) B200FM_AR_CASE(8)
        default: B200FM_CHECK(false, "allreduce_f32: unsupported world size %d", world);
    }
#undef B200FM_AR_CASE
    B200FM_CUDA(cudaGetLastError());
    return 0;
}
