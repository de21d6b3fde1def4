// umma_probe.cu — measurement kernel: how fast does one SM execute back-to-back tcgen05.mma (128 x N x 16, fp16 -> fp32) for a
// given operand placement?  One CTA per SM, one thread issues `iters` MMAs on (uninitialised) operands and times them with clock64.
// It decides the operand layout of the field kernels (profiles/r2_umma_probe.md): mode
//   0  A and B in shared memory, canonical K-major no-swizzle layout (umma.cuh; what round 1 used)
//   1  A and B in shared memory, K-major SWIZZLE_128B layout
//   2  A in tensor memory, B in shared memory no-swizzle
//   3  A in tensor memory, B in shared memory SWIZZLE_128B
#include "umma.cuh"

namespace seald {

__device__ __forceinline__ uint64_t smem_desc_sw128(const uint32_t saddr) {
    // K-major, 128-byte swizzle: rows of 64 halves (128 B), 8-row groups 1024 B apart (SBO); LBO unused (1); layout type 2 in [61,64)
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((1024u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mma_f16_ts(const uint32_t tmem_d, const uint32_t tmem_a, const uint64_t desc_b, const uint32_t idesc,
                                           const uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) k_umma_probe(const int mode, const int iters, const int n_cols, const int issuers,
                                                        long long* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ long long s_t[4][2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { umma::mbar_init(&bar, (uint32_t)issuers); umma::mbar_fence_init(); }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
    for (uint32_t i = tid; i < 2 * 32768 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if ((tid & 31) == 0 && warp < issuers) {
        const uint32_t a_addr = umma::smem_addr(smem), b_addr = a_addr + 32768;
        const uint32_t idesc = umma::instr_desc_f16(128, (uint32_t)n_cols);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const int k = i & 7;  // the 8 K = 16 slices of a 128 x 128 operand tile
            uint64_t da, db;
            if (mode & 1) {
                // swizzled: slice k lives in 64-column slab k / 4 (16 KiB each), 32 B further per slice inside the 128-byte row
                da = smem_desc_sw128(a_addr + (k >> 2) * 16384 + (k & 3) * 32);
                db = smem_desc_sw128(b_addr + (k >> 2) * 16384 + (k & 3) * 32);
            } else {
                da = umma::smem_desc(a_addr + k * 2 * (128 * 16), 128 * 16, 128);
                db = umma::smem_desc(b_addr + k * 2 * (n_cols * 16), n_cols * 16, 128);
            }
            const uint32_t d = tmem + (uint32_t)warp * ((mode & 2) ? 64u : 128u);  // one accumulator per issuing warp (64-column ones beside A in TMEM)
            if (mode & 2) mma_f16_ts(d, tmem + 256 + k * 8, db, idesc, i > 0 ? 1u : 0u);  // A: 128 lanes x 8 columns (16 halves) per slice
            else umma::mma_f16(d, da, db, idesc, i > 0 ? 1u : 0u);
        }
        umma::mma_commit(&bar);
        const long long t1 = clock64();
        umma::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        s_t[warp][0] = t1 - t0;
        s_t[warp][1] = t2 - t0;
    }
    __syncthreads();
    if (tid == 0) {
        long long a = 0, b = 0;
        for (int w = 0; w < issuers; w++) { a = max(a, s_t[w][0]); b = max(b, s_t[w][1]); }
        out[blockIdx.x * 2] = a;
        out[blockIdx.x * 2 + 1] = b;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace seald

using namespace seald;

extern "C" int seald_umma_probe(int mode, int iters, int n_cols, int issuers, int64_t* out_cycles, seald_stream_t stream) {
    if (!out_cycles || iters <= 0 || mode < 0 || mode > 3 || n_cols < 16 || n_cols > 128 || n_cols % 16 || issuers < 1 || issuers > 4)
        return SEALD_E_BADARG;
    if ((mode & 2) && (n_cols > 64 || issuers > 4)) return SEALD_E_BADARG;
    const size_t smem = 2 * 32768 + 1024;
    cudaError_t e = cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_umma_probe<<<SEALD_NUM_SMS, 128, smem, to_stream(stream)>>>(mode, iters, n_cols, issuers, reinterpret_cast<long long*>(out_cycles));
    return launch_status();
}
