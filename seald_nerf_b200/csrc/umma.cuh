// umma.cuh — thin PTX wrappers for the Blackwell (sm_100a) tensor-core path: tcgen05.mma with accumulators in tensor
// memory (TMEM), mbarriers, 1-D bulk copies (cp.async.bulk, the TMA engine without a tensor map) and the shared-memory
// matrix descriptors of the canonical K-major no-swizzle layout.
//
// Canonical K-major operand tile used everywhere in this library (rows = M or N index, K contiguous in 16-byte chunks):
//
//     element (row r, column k)  ->  byte offset  ((k / 8) * ROWS + r) * 16 + (k % 8) * 2
//
// i.e. [K/8][ROWS][8 halves]: an 8-row x 8-column core matrix is 128 contiguous bytes (rows 16 B apart), the next 8 rows
// follow 128 B later (stride byte offset, SBO) and the next 8 columns ROWS*16 B later (leading byte offset, LBO).  One
// tcgen05.mma.kind::f16 consumes K = 16 (two chunks); the k-th MMA of a tile starts 2*k*LBO further.  A thread that owns
// one row writes its 16-byte chunks at consecutive addresses across the warp: conflict-free.
#pragma once
#include "common.cuh"

namespace seald {
namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug traps (launch error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const uint32_t parity) {
    const uint32_t a = smem_addr(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (spins > (1u << 24)) __trap();
    }
}

// ---- bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0, both addresses 16-byte aligned) ----------
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, const uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory ---------------------------------------------------------------------------------
// One full warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, const uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(const uint32_t base, const uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i), columns [col, col + 32).
__device__ __forceinline__ void tmem_ld32(const uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(const uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 in [0,14), LBO >> 4 in
// [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64)).
__device__ __forceinline__ uint64_t smem_desc(const uint32_t saddr, const uint32_t lbo_bytes, const uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// Instruction descriptor for kind::f16: fp16 A and B (both K-major), fp32 accumulator, M x N tile
// (cute::UMMA::InstrDescriptor: c_format [4,6) = 1 (F32), a/b_format = 0 (F16), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4).
__host__ __device__ constexpr uint32_t instr_desc_f16(const uint32_t M, const uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void mma_f16(const uint32_t tmem_d, const uint64_t desc_a, const uint64_t desc_b, const uint32_t idesc,
                                        const uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All tcgen05.mma issued so far by this thread -> one arrival on `bar` when they have completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

}  // namespace umma
}  // namespace seald
