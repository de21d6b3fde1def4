// wgrad_umma.cu — weight gradients dW[out][in] += sum_rows G[row][out] * A[row][in] of all MLP layers on the Blackwell
// tensor cores (replaces the CUTLASS split-K GEMMs of ffmlp_backward, ffmlp/src/ffmlp.cu:801-877, and the mma.sync kernel
// k_wgrad of field.cu).
//
// The contraction runs over SAMPLE ROWS, so both operands are "MN-major" for the MMA: G^T is M(out) x K(rows), A is
// N(in) x K(rows), and the saved activations / activation gradients are row-major [rows][width] in HBM, i.e. contiguous
// along M resp. N.  tcgen05.mma reads MN-major operands directly (instruction descriptor a_major = b_major = 1) from the
// canonical no-swizzle layout
//
//     element (mn, k)  ->  (k / 8) * LBO + (mn / 8) * 128 B + (k % 8) * 16 B + (mn % 8) * 2 B
//
// which is a pure permutation of 16-byte pieces of the row-major tile: staging is plain cp.async (LDGSTS), 16 bytes per
// thread, no transposes, no ldmatrix, no register accumulators.  One CTA streams a balanced share of one layer's rows
// through a 3-stage ring of 64-row chunks (4 MMAs of 128 x N x 16 each), accumulates the fp32 128 x N tile in tensor
// memory and flushes it once with vector atomics into the flat gradient buffer.  The kernel is a pure HBM stream
// (4.7 KB per sample over all 13 layers); per chunk a thread issues 8 cp.async, one fence and one barrier.
//
// TILE-IMAGE operands (ldg == 0 and lda == 0 in the job): the tcgen05 deformation kernels (field_umma.cu) save activations and
// activation gradients exactly as their K-major shared-memory A tiles, [tile][width / 8][128 rows][8 halves].  Read with the roles
// of the dimensions swapped that IS the MN-major operand of this kernel (k = sample row: 16 B apart inside a core matrix, next 8
// rows LBO = 128 B, next 8 columns SBO = 2048 B), so a whole 128-row tile of an operand is ONE contiguous block of HBM: a producer
// thread fetches G and A of a tile with two cp.async.bulk copies (the TMA engine, <= 32 KiB each) into a 3-stage ring that
// completes on an mbarrier, an issuer thread fires the 8 MMAs of the tile, nobody else touches the data.  Dead rows of the last
// live tile are zeros in the saved images, so no row masking is needed.
#include <cstdlib>

#include "umma.cuh"

namespace seald {

struct UWgradJob {
    const __half* G;  // [rows][ldg], n_out columns used (multiple of 8, <= 128)
    const __half* A;  // [rows][lda], n_in columns used (multiple of 16, <= 128)
    float* dW;        // [n_real][ldw]
    int n_out, n_in;
    int ldg, lda, ldw;
    int n_real, k_real;
};
constexpr int kUWMaxJobs = 16;
struct UWgradJobs {
    UWgradJob j[kUWMaxJobs];
    int first_cta[kUWMaxJobs + 1];
    int n_jobs;
};

constexpr int kUWChunk = 64;     // rows per stage
constexpr int kUWStages = 3;
constexpr int kUWMaxStages = 6;      // row-major path inside the 192 KiB ring
constexpr int kUWMaxTileStages = 8;  // tile-image path: stages of (G tile + A tile) bytes
constexpr int kUWThreads = 128;  // 4 warps: all stage, thread 0 issues the MMAs, all flush (thread t = output row t)
constexpr uint32_t kUWOperandBytes = (kUWChunk / 8) * 16 * 128;  // 16 KiB: [k/8][mn/8 = 16][k%8][mn%8] for a 128-wide operand
constexpr uint32_t kUWStageBytes = 2 * kUWOperandBytes;
constexpr size_t kUWSmem = kUWStages * kUWStageBytes + 256;
// tile-image path: a stage holds one 128-row tile of G (<= 32 KiB) and of A (<= 32 KiB)
constexpr int kUWTileRows = 128;
constexpr uint32_t kUWTileOperandBytes = 128 * kUWTileRows * 2;
constexpr uint32_t kUWTileStageBytes = 2 * kUWTileOperandBytes;
constexpr size_t kUWSmemTile = kUWStages * kUWTileStageBytes + 256;

// instruction descriptor: fp16 x fp16 -> fp32, A and B MN-major (bits 15 / 16), M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_mn_major(const uint32_t n) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(kUWThreads) k_wgrad_umma(const __grid_constant__ UWgradJobs jobs, const int M, const int* __restrict__ m_dev,
                                                           const uint32_t ring_bytes, int* __restrict__ found_inf) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar_empty = reinterpret_cast<uint64_t*>(smem + ring_bytes);  // [<= 8] the MMAs reading a stage have completed
    uint64_t* bar_done = bar_empty + kUWMaxTileStages;                     // all MMAs of this CTA have completed
    uint64_t* bar_full = bar_done + 1;                                     // [<= 8] tile path: the bulk copies of a stage have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + kUWMaxTileStages);
    // row-major path: as many 64-row stages as the ring holds (3 in the 96 KiB ring, 6 beside tile-image jobs)
    const int row_stages = (int)(ring_bytes / kUWStageBytes);

    int job = 0;
    while (job + 1 < jobs.n_jobs && (int)blockIdx.x >= jobs.first_cta[job + 1]) job++;
    const UWgradJob& jb = jobs.j[job];
    const int part = (int)blockIdx.x - jobs.first_cta[job], parts = jobs.first_cta[job + 1] - jobs.first_cta[job];
    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    const bool tile_image = (jb.ldg == 0);
    // row-major path: a balanced range of rows; tile-image path: a balanced range of 128-row tiles
    const int rows_per_cta = ((m_used + parts - 1) / parts + kUWChunk - 1) / kUWChunk * kUWChunk;
    const int n_tiles = (m_used + kUWTileRows - 1) / kUWTileRows;
    const int tiles_per_cta = (n_tiles + parts - 1) / parts;
    const int m_begin = tile_image ? part * tiles_per_cta : part * rows_per_cta;                              // (tiles resp. rows)
    const int m_end = tile_image ? min(n_tiles, m_begin + tiles_per_cta) : min(m_used, m_begin + rows_per_cta);
    if (m_begin >= m_end) return;  // (uniform per CTA, before any barrier / allocation)

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kUWMaxTileStages; i++) { umma::mbar_init(bar_empty + i, 1); umma::mbar_init(bar_full + i, 1); }
        umma::mbar_init(bar_done, 1);
        umma::mbar_fence_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 128);
    // row-major path: zero the ring once (columns of G beyond n_out are never written again).  Tile path: nothing to zero — the
    // M = 128 MMA reads 128 - n_out rows of whatever follows a narrow G tile, but those only reach output rows >= n_out, which are
    // never flushed (rows of D are independent).
    if (!tile_image) {
        for (uint32_t i = tid; i < row_stages * kUWStageBytes / 16; i += kUWThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t smem_base = umma::smem_addr(smem);
    const uint32_t idesc = idesc_mn_major((uint32_t)jb.n_in);

    if (tile_image) {
        // ---- tile-image operands: producer thread (bulk copies) and issuer thread (MMAs) talk through mbarriers only
        // a stage = the tile's G image followed by its A image; as many stages as the ring holds (<= 8: narrow head layers get a
        // deeper pipeline, 3 x 64 KiB for the 128-wide layers), leaving room for the M = 128 over-read behind the last G tile
        const int n = m_end - m_begin;
        const uint32_t g_bytes = (uint32_t)jb.n_out * kUWTileRows * 2, a_bytes = (uint32_t)jb.n_in * kUWTileRows * 2;
        const uint32_t stage_bytes = g_bytes + a_bytes;
        const int n_st = min(kUWMaxTileStages, (int)((ring_bytes - (kUWTileOperandBytes - g_bytes)) / stage_bytes));
        if (tid == 32) {
            const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(jb.G) + (size_t)m_begin * g_bytes;
            const unsigned char* asrc = reinterpret_cast<const unsigned char*>(jb.A) + (size_t)m_begin * a_bytes;
            for (int i = 0; i < n; i++) {
                const int st = i % n_st;
                if (i >= n_st) umma::mbar_wait(bar_empty + st, ((i / n_st) - 1) & 1);  // the MMAs of tile i - n_st are done
                unsigned char* sg = smem + (size_t)st * stage_bytes;
                umma::mbar_arrive_expect_tx(bar_full + st, stage_bytes);
                umma::bulk_load(sg, gsrc + (size_t)i * g_bytes, g_bytes, bar_full + st);
                umma::bulk_load(sg + g_bytes, asrc + (size_t)i * a_bytes, a_bytes, bar_full + st);
            }
        } else if (tid == 0) {
            for (int i = 0; i < n; i++) {
                const int st = i % n_st;
                umma::mbar_wait(bar_full + st, (i / n_st) & 1);
                umma::fence_after_sync();
                const uint32_t g0 = smem_base + st * stage_bytes, a0 = g0 + g_bytes;
#pragma unroll
                for (int k = 0; k < kUWTileRows / 16; k++) {
                    // K = 16 sample rows = two k-groups 128 B apart (LBO); next 8 columns 2048 B further (SBO)
                    const uint64_t da = umma::smem_desc(g0 + k * 256, 128, 2048);
                    const uint64_t db = umma::smem_desc(a0 + k * 256, 128, 2048);
                    umma::mma_f16(tmem_d, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                umma::mma_commit(bar_empty + st);
            }
            umma::mma_commit(bar_done);
        }
        __syncwarp();
    } else {
    // ---- staging: a chunk of operand X is 64 rows x (w / 8) 16-byte pieces.  Lanes 0-7 of a warp take rows r..r+7 of one piece
    //      column (one 128-byte core matrix in shared memory), the next 8 lanes the neighbouring column: a warp reads full
    //      32-byte sectors and writes 512 contiguous bytes.  No index decode: (row-in-block, column group) are fixed per thread.
    const int g_cols = jb.n_out / 8, a_cols = jb.n_in / 8, cols = g_cols + a_cols;  // <= 32 columns
    const int rsub = tid & 7, cgrp = tid >> 3;                                       // 16 column groups x 8 rows per pass

    const int n_chunks = (m_end - m_begin + kUWChunk - 1) / kUWChunk;

    auto stage_chunk = [&](const int chunk) {
        const int s = chunk % row_stages;
        unsigned char* sg = smem + (size_t)s * kUWStageBytes;
        unsigned char* sa = sg + kUWOperandBytes;
        const int row0 = m_begin + chunk * kUWChunk;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int call = cgrp + 16 * j;
            if (call >= cols) break;
            const bool is_g = call < g_cols;
            const int c = is_g ? call : call - g_cols;
            unsigned char* dcol = (is_g ? sg : sa) + c * 128 + rsub * 16;
            const __half* scol = is_g ? jb.G + c * 8 : jb.A + c * 8;
            const int ld = is_g ? jb.ldg : jb.lda;
#pragma unroll
            for (int rb = 0; rb < kUWChunk / 8; rb++) {
                const int row = row0 + rb * 8 + rsub;
                unsigned char* dst = dcol + rb * 2048;
                if (row < m_end) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_addr(dst)), "l"(scol + (size_t)row * ld) : "memory");
                } else {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ---- pipeline: chunk c is staged while chunks c-1, c-2 are in flight / being multiplied
    for (int c = 0; c < n_chunks + row_stages - 1; c++) {
        if (c < n_chunks) {
            if (c >= row_stages) umma::mbar_wait(bar_empty + (c % row_stages), ((c / row_stages) - 1) & 1);  // the MMAs of chunk c - stages are done
            stage_chunk(c);
        } else {
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const int ready = c - (row_stages - 1);  // this chunk's copies (issued stages - 1 iterations ago) have landed after wait_group<stages - 1>
        if (ready >= 0) {
            if (row_stages == kUWMaxStages) asm volatile("cp.async.wait_group %0;" ::"n"(kUWMaxStages - 1) : "memory");
            else asm volatile("cp.async.wait_group %0;" ::"n"(kUWStages - 1) : "memory");
            umma::fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
                const int s = ready % row_stages;
                const uint32_t g0 = smem_base + s * kUWStageBytes, a0 = g0 + kUWOperandBytes;
#pragma unroll
                for (int k = 0; k < kUWChunk / 16; k++) {
                    // K = 16 rows = two k-groups of 2048 B; SBO (next 8 mn) = 128 B, LBO (next 8 k) = 2048 B
                    const uint64_t da = umma::smem_desc(g0 + k * 4096, 2048, 128);
                    const uint64_t db = umma::smem_desc(a0 + k * 4096, 2048, 128);
                    umma::mma_f16(tmem_d, da, db, idesc, (ready > 0 || k > 0) ? 1u : 0u);
                }
                umma::mma_commit(bar_empty + s);
                if (ready == n_chunks - 1) umma::mma_commit(bar_done);
            }
        }
    }

    }  // row-major path

    // ---- flush: thread t owns output row t (TMEM lane t), n_in fp32 columns
    umma::mbar_wait(bar_done, 0);
    umma::fence_after_sync();
    const uint32_t t_lane = tmem_d + ((uint32_t)(tid & ~31) << 16);
    const bool row_ok = tid < jb.n_real;
    float* out_row = jb.dW + (size_t)tid * jb.ldw;
    const bool vec_ok = (jb.ldw % 4 == 0) && (((uintptr_t)jb.dW & 15) == 0);
    bool bad = false;  // a non-finite partial sum makes the accumulated gradient non-finite: GradScaler's found_inf (optional)
    for (int q = 0; q < jb.n_in; q += 16) {
        uint32_t v[16];
        umma::tmem_ld16(t_lane + q, v);
        umma::wait_ld();
        if (row_ok) {
            if (found_inf) {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    if (q + i < jb.k_real) bad |= !isfinite(__uint_as_float(v[i]));
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const int col = q + i;
                if (vec_ok && col + 3 < jb.k_real) {
                    atomicAdd(reinterpret_cast<float4*>(out_row + col),
                              make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (col + e < jb.k_real) atomicAdd(out_row + col + e, __uint_as_float(v[i + e]));
                }
            }
        }
    }
    if (found_inf && __any_sync(0xffffffffu, bad) && (tid & 31) == 0) atomicOr(found_inf, 0x3f800000);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        umma::tmem_dealloc(tmem_d, 128);
    }
}

}  // namespace seald

using namespace seald;

// Same job table as seald_mlp_wgrad (include/seald_b200.h); N = out columns of G (16 / 64 / 128), K = in columns of A.
static int wgrad_umma_launch(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, int32_t* found_inf, seald_stream_t stream);

extern "C" int seald_mlp_wgrad_umma(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, seald_stream_t stream) {
    return wgrad_umma_launch(jobs, n_jobs, M, m_dev, nullptr, stream);
}

// + GradScaler's overflow flag: *found_inf |= 0x3f800000 when a partial sum added to a weight gradient is not finite
extern "C" int seald_mlp_wgrad_umma_flag(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, int32_t* found_inf,
                                         seald_stream_t stream) {
    return wgrad_umma_launch(jobs, n_jobs, M, m_dev, found_inf, stream);
}

static int wgrad_umma_launch(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, int32_t* found_inf, seald_stream_t stream) {
    if (M == 0 || n_jobs == 0) return 0;
    if (!jobs || n_jobs < 0 || n_jobs > kUWMaxJobs) return SEALD_E_BADARG;
    UWgradJobs js;
    js.n_jobs = n_jobs;
    int cost = 0;
    bool any_tile = false;
    // share of CTAs per byte: the cp.async path pays a barrier per 64-row chunk, so next to the bulk-copy jobs its jobs get row_w / 2 times
    // the CTAs per byte (measured, SEALD_WGRAD_ROWW)
    static const int row_w = [] { const char* e = getenv("SEALD_WGRAD_ROWW"); return e ? atoi(e) : 8; }();
    for (int i = 0; i < n_jobs; i++) {
        const seald_wgrad_job& a = jobs[i];
        if (!a.G || !a.A || !a.dW) return SEALD_E_BADARG;
        if (a.N % 8 || a.N <= 0 || a.N > 128 || a.K % 16 || a.K <= 0 || a.K > 128 || a.ldg % 8 || a.lda % 8 || a.n_real > 128) return SEALD_E_UNSUPPORTED;
        if ((a.ldg == 0) != (a.lda == 0)) return SEALD_E_UNSUPPORTED;  // tile-image operands come in pairs (both saved by field_umma.cu)
        any_tile = any_tile || a.ldg == 0;
        if (((uintptr_t)a.G % 16) || ((uintptr_t)a.A % 16)) return SEALD_E_ALIGN;
        UWgradJob& j = js.j[i];
        j.G = (const __half*)a.G; j.A = (const __half*)a.A; j.dW = a.dW;
        j.n_out = a.N; j.n_in = a.K; j.ldg = a.ldg; j.lda = a.lda; j.ldw = a.ldw; j.n_real = a.n_real; j.k_real = a.k_real;
        cost += (a.N + a.K) * (a.ldg == 0 ? 2 : row_w);
    }
    // one balanced wave, CTAs shared out by bytes per row: 2 CTAs per SM with the 96 KiB ring of the row-major path, 1 CTA per SM
    // with the 192 KiB ring of whole-tile bulk copies
    int budget = (any_tile ? 1 : 2) * SEALD_NUM_SMS;
    const int max_by_rows = (int)div_up(M, 256u) * n_jobs;
    if (budget > max_by_rows) budget = max_by_rows;
    if (budget < n_jobs) budget = n_jobs;
    int total = 0;
    for (int i = 0; i < n_jobs; i++) {
        int parts = (int)((long long)budget * (js.j[i].n_out + js.j[i].n_in) * (js.j[i].ldg == 0 ? 2 : row_w) / cost);
        if (parts < 1) parts = 1;
        js.first_cta[i] = total;
        total += parts;
    }
    js.first_cta[n_jobs] = total;
    const size_t smem = any_tile ? kUWSmemTile : kUWSmem;
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUWSmemTile);
    if (e != cudaSuccess) return (int)e;
    k_wgrad_umma<<<total, kUWThreads, smem, to_stream(stream)>>>(js, (int)M, m_dev, (uint32_t)(smem - 256), found_inf);
    return launch_status();
}
