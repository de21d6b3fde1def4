// field_umma.cu — the D-NeRF deformation MLP (76 -> 8 x Linear(128) + ReLU -> 3, dnerf/network.py:123-143) on the
// Blackwell tensor cores: tcgen05.mma with fp32 accumulators in tensor memory.
//
// One persistent CTA per SM works on G = 2 or 4 128-sample tiles at a time ("groups"); each group is an independent pipeline of
// 4 epilogue warps + ONE ISSUER WARP, with its own TMEM accumulator (128 columns) and its own pair of mbarriers:
//
//   issuer warp g (one elected thread): waits for the group's A tile and for the layer's weights, issues the tcgen05.mma's of that
//                                 tile-layer (K/16 instructions of shape 128 x 128 x 16, last layer 128 x 16 x 16) and commits them to
//                                 the group's mbarrier (and to the weight stage's "free" barrier).
//   loader warp (one thread)    : streams the layer weights global -> shared with cp.async.bulk (3 stages of 32 KiB, pre-packed in
//                                 the canonical K-major layout of umma.cuh), refilling a stage once every group's MMAs on it are done.
//   4 warps per group (8 for G = 2: two warpgroups, each half of the output columns — the small-batch layer step is epilogue bound:
//                                 backward 34.7 -> 25.8 us at 31.7k samples)
//                               : thread t owns sample row t of its tile = TMEM lane t.  Layer 0 input: frequency
//                                 encoding of xyz (63) and t (13) written straight into the A-operand tile.  After each
//                                 layer: tcgen05.ld the fp32 row, ReLU, pack to fp16, store as the next layer's A tile
//                                 (and, when training, into fwd_buf for the backward pass).  Last layer: dx, x' = x + dx,
//                                 x01 = (x' + bound) / (2 bound).
//
// Why one issuer per group (profiles/r2_umma_probe.md): a single thread gets one 128 x 128 x 16 MMA accepted every ~165 cycles
// whatever the operand placement (shared memory with or without swizzle, A in tensor memory), 2.5x the 64-cycle floor of the tensor
// pipe; four threads issuing into four accumulators reach one MMA per 77 cycles.  Round 1's single control thread also served the
// groups one after the other (~180 cycles per mbarrier wait with the pipe idle), which put the kernel at 36% of the tensor peak.
//
// Activations never leave the SM between layers; per 128-sample tile the tensor pipe does 8 layers x 4.2 MFLOP while the
// only HBM traffic is 12 B in / 24 B out per sample (inference).  Measured bound (profiles/): shared-memory bandwidth —
// an SS-mode 128x128x16 MMA reads 8 KiB of operands per 64 cycles (= the 128 B/cycle port), plus 32 KiB of epilogue
// stores and the weight refills per tile-layer, ~830 cycles against the 512-cycle MMA floor.  Numerics as the mma.sync kernels of field.cu: fp16
// operands and layer outputs, fp32 accumulation.
#include <cstdlib>

#include "encoders.cuh"
#include "grid_device.cuh"
#include "umma.cuh"

namespace seald {

constexpr int kUW = 128;            // hidden width
constexpr int kUK0 = 80;            // layer-0 K (76 real inputs, zero padded)
constexpr int kUNLast = 16;         // last layer N (3 real outputs, zero padded)
constexpr int kUStages = 3;
constexpr uint32_t kTileBytes = kUW * kUW * 2;  // 32 KiB: one A tile / one weight stage

template <int G>
struct UmmaSmem {
    // G == 2 (training-size batches: one or two tiles per SM, the step of a layer is dominated by its epilogue): TWO warpgroups per
    // tile, each taking half of the 128 output columns (warps w and w + 4 of a group share a tensor-memory lane quadrant)
    static constexpr int EW = (G == 2) ? 2 : 1;
    static constexpr int EPI = 128 * EW;                    // epilogue threads per group
    static constexpr int THREADS = G * EPI + (G + 1) * 32;  // G groups x 4 EW epilogue warps + G issuer warps + 1 loader warp
    static constexpr uint32_t TMEM_COLS = G * kUW;
    static constexpr size_t A_OFF = 0;
    static constexpr size_t W_OFF = G * kTileBytes;
    static constexpr size_t BAR_OFF = W_OFF + kUStages * kTileBytes;
    static constexpr size_t BYTES = BAR_OFF + 128;
    static constexpr size_t LP_OFF = BYTES;                                    // density kernel: the 16 hash-grid level records
    static constexpr size_t BYTES_DENS = LP_OFF + 16 * sizeof(LevelParams);
};

// ---- density query (k_deform_forward_umma<.., DENS = true>): the sigma head rides the same pipeline as extra tile-layers ----
constexpr int kSK0 = 32;  // sigma head input: 16 levels x 2 features
constexpr int kSW = 64;   // sigma head hidden width
constexpr int kSNLast = 16;
struct DensityArgs {
    const __half* packed_sigma;  // sigma head weights as operand tiles (seald_field_umma_pack_sigma)
    int n_sigma;
    const __half* table;         // fp16 hash table (16-byte aligned)
    const int* offsets;          // [17]
    float S;
    uint32_t H, gridtype, interp;
    bool align_corners;
    float density_scale;
    float* sigma;                // [M] or null
    const int* indices;          // optional scatter: tmp[indices[row]] = sigma * store_scale (occupancy refresh)
    float store_scale;
    float* tmp;
};
__host__ __device__ __forceinline__ uint32_t sigma_layer_bytes(const int s, const int n_sigma) {
    return (uint32_t)((s == 0 ? kSK0 : kSW) * (s == n_sigma - 1 ? kSNLast : kSW) * 2);
}
__host__ __device__ __forceinline__ size_t sigma_layer_offset(const int s, const int n_sigma) {
    size_t o = 0;
    for (int i = 0; i < s; i++) o += sigma_layer_bytes(i, n_sigma);
    return o;
}
// shape of tile-layer l of the chain deformation net (n_layers) -> sigma head (n_sigma)
__device__ __forceinline__ void chain_shape(const int l, const int n_layers, const int n_sigma, int& K, int& N) {
    if (l < n_layers) {
        K = (l == 0) ? kUK0 : kUW;
        N = (l == n_layers - 1) ? kUNLast : kUW;
    } else {
        const int sl = l - n_layers;
        K = (sl == 0) ? kSK0 : kSW;
        N = (sl == n_sigma - 1) ? kSNLast : kSW;
    }
}

__host__ __device__ __forceinline__ uint32_t umma_layer_bytes(const int l, const int n_layers) {
    if (l == 0) return kUK0 * kUW * 2;
    if (l == n_layers - 1) return kUW * kUNLast * 2;
    return kTileBytes;
}
__host__ __device__ __forceinline__ size_t umma_layer_offset(const int l, const int n_layers) {
    size_t o = 0;
    for (int i = 0; i < l; i++) o += umma_layer_bytes(i, n_layers);
    return o;
}

// src [n_real][ld] row-major fp16 (nn.Linear.weight) -> dst [K/8][n_pad][8] (rows >= n_real zero); all layers in one launch
struct PackJobs {
    const __half* src[12];
    int n_layers;
};
__device__ __forceinline__ void pack_umma_body(const PackJobs& jobs, __half* __restrict__ packed, const int l) {
    const int n_layers = jobs.n_layers;
    const bool last = (l == n_layers - 1);
    const int K = (l == 0) ? kUK0 : kUW, n_pad = last ? kUNLast : kUW, n_real = last ? 3 : kUW, ld = K;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
    if (i >= K / 8 * n_pad) return;
    const int c = i / n_pad, n = i - c * n_pad;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < n_real) v = *reinterpret_cast<const uint4*>(jobs.src[l] + (size_t)n * ld + c * 8);
    __half* dst = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(packed) + umma_layer_offset(l, n_layers));
    *reinterpret_cast<uint4*>(dst + (size_t)i * 8) = v;
}
__global__ void k_pack_umma(const PackJobs jobs, __half* __restrict__ packed) { pack_umma_body(jobs, packed, blockIdx.y); }

// sigma head: layer s [N][K] row-major (ld = K) -> [K/8][N][8]
__global__ void k_pack_umma_sigma(const PackJobs jobs, __half* __restrict__ packed) {
    const int s_l = blockIdx.y, n_sigma = jobs.n_layers;
    const int K = (s_l == 0) ? kSK0 : kSW, N = (s_l == n_sigma - 1) ? kSNLast : kSW;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K / 8 * N) return;
    const int c = i / N, n = i - c * N;
    const uint4 v = *reinterpret_cast<const uint4*>(jobs.src[s_l] + (size_t)n * K + c * 8);
    __half* dst = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(packed) + sigma_layer_offset(s_l, n_sigma));
    *reinterpret_cast<uint4*>(dst + (size_t)i * 8) = v;
}

template <bool SAVE, int G, bool DENS>
__global__ void __launch_bounds__(UmmaSmem<G>::THREADS, 1) k_deform_forward_umma(const float* __restrict__ xyz, const float* __restrict__ time,
                                                                      const __half* __restrict__ packed, const int n_layers, const int M,
                                                                      const int* __restrict__ m_dev, const float bound, const int t0_mode,
                                                                      float* __restrict__ deform, float* __restrict__ x01,
                                                                      __half* __restrict__ in_buf, __half* __restrict__ fwd_buf,
                                                                      const DensityArgs da) {
    extern __shared__ __align__(128) unsigned char smem[];
    using SM = UmmaSmem<G>;
    constexpr int ISSUE0 = G * 4 * SM::EW;  // warps ISSUE0 .. ISSUE0 + G - 1 issue the MMAs of group 0 .. G - 1
    constexpr int LOADER = ISSUE0 + G;      // streams the weights
    unsigned char* s_a = smem + SM::A_OFF;
    unsigned char* s_w = smem + SM::W_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
    uint64_t* bar_full = bars;            // [3] weights of a stage have landed
    uint64_t* bar_wfree = bars + 3;       // [3] every group's MMAs on a stage have completed (count G)
    uint64_t* bar_aready = bars + 6;      // [G] the group's A tile is written
    uint64_t* bar_mma = bars + 6 + G;     // [G] the group's MMAs of the current layer have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * G);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    const int n_tiles = (m_used + kUW - 1) / kUW;
    const int n_pairs = (n_tiles + G - 1) / G;  // work units of G tiles
    const int my_pairs = (n_pairs > (int)blockIdx.x) ? (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    // DENS: the chain continues through the hash grid (gathered by the epilogue threads into the A tile) and the sigma head
    const int n_chain = DENS ? n_layers + da.n_sigma : n_layers;
    const int n_items = my_pairs * n_chain;
    LevelParams* s_lp = reinterpret_cast<LevelParams*>(smem + SM::LP_OFF);
    // (The groups of a CTA share the weight ring and therefore stay within two tile-layers of each other: they reach the gather
    // together and the tensor pipe idles meanwhile — why this variant loses to the two-kernel query on big batches,
    // profiles/r2_occupancy.md.  Staggering their start would dead-lock on the ring.)
    if constexpr (DENS) {
        if (tid < 16) s_lp[tid] = make_level(da.offsets, tid, da.S, da.H, 3, da.gridtype, da.align_corners);
    }

    if (tid == 0) {
        for (int i = 0; i < 3; i++) { umma::mbar_init(bar_full + i, 1); umma::mbar_init(bar_wfree + i, G); }
        for (int i = 0; i < G; i++) { umma::mbar_init(bar_aready + i, SM::EPI); umma::mbar_init(bar_mma + i, 1); }
        umma::mbar_fence_init();
    }
    if (warp == LOADER) umma::tmem_alloc(tmem_slot, SM::TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == LOADER) {
        // =============================== loader: weights of item i into stage i % 3, two items ahead of the MMAs ===============================
        if (lane == 0) {
            for (int i = 0; i < n_items; i++) {
                const int l = i % n_chain, s = i % kUStages;
                if (i >= kUStages) umma::mbar_wait(bar_wfree + s, ((i / kUStages) - 1) & 1);  // every group is done with item i - 3
                uint32_t bytes;
                const unsigned char* src;
                if (!DENS || l < n_layers) {
                    bytes = umma_layer_bytes(l, n_layers);
                    src = reinterpret_cast<const unsigned char*>(packed) + umma_layer_offset(l, n_layers);
                } else {
                    bytes = sigma_layer_bytes(l - n_layers, da.n_sigma);
                    src = reinterpret_cast<const unsigned char*>(da.packed_sigma) + sigma_layer_offset(l - n_layers, da.n_sigma);
                }
                umma::mbar_arrive_expect_tx(bar_full + s, bytes);
                umma::bulk_load(s_w + (size_t)s * kTileBytes, src, bytes, bar_full + s);
            }
        }
    } else if (warp >= ISSUE0) {
        // =============================== issuer of group g: its tile's MMAs, layer after layer ===============================
        const int g = warp - ISSUE0;
        if (lane == 0) {
            const uint32_t a0 = umma::smem_addr(s_a) + g * kTileBytes, w_addr = umma::smem_addr(s_w);
            const uint32_t d = tmem_base + g * kUW;
            for (int i = 0; i < n_items; i++) {
                const int l = i % n_chain, s = i % kUStages;
                int K_l, N_l;
                chain_shape(l, n_layers, DENS ? da.n_sigma : 0, K_l, N_l);
                const int ksteps = K_l / 16;
                const uint32_t n_rows = (uint32_t)N_l;
                const uint32_t idesc = umma::instr_desc_f16(128, n_rows);
                const uint32_t w_lbo = n_rows * 16;
                umma::mbar_wait(bar_full + s, (i / kUStages) & 1);
                umma::mbar_wait(bar_aready + g, i & 1);
                umma::fence_after_sync();
                const uint32_t w0 = w_addr + s * kTileBytes;
                for (int k = 0; k < ksteps; k++) {
                    const uint64_t da = umma::smem_desc(a0 + k * 2 * (kUW * 16), kUW * 16, 128);
                    const uint64_t db = umma::smem_desc(w0 + k * 2 * w_lbo, w_lbo, 128);
                    umma::mma_f16(d, da, db, idesc, k > 0 ? 1u : 0u);
                }
                umma::mma_commit(bar_mma + g);    // -> the group's epilogue warps
                umma::mma_commit(bar_wfree + s);  // -> the loader (stage s may be refilled once all G groups arrived)
            }
        }
    } else {
        // =============================== epilogue groups: one thread per sample row ===============================
        const int g = tid / SM::EPI, te = tid - g * SM::EPI, r = te & 127, half = te >> 7;  // half: which share of the columns (EW == 2)
        constexpr int QN = 4 / SM::EW, C8N = (kUK0 / 8) / SM::EW;
        unsigned char* a_tile = s_a + (size_t)g * kTileBytes;
        const uint32_t t_lane = tmem_base + ((uint32_t)(r & ~31) << 16) + g * kUW;  // this warp's 32-lane quadrant, group's columns
        const float tval = *time;
        const bool t_is_zero = (tval == 0.0f);
        int row = 0, tile = 0;
        bool save_tile = false;
        const size_t tiles_cap = (size_t)(M + kUW - 1) / kUW;  // tiles the saved-activation buffers hold per layer
        float px = 0.f, py = 0.f, pz = 0.f;
        for (int i = 0; i < n_items; i++) {
            const int l = i % n_chain;
            if (l == 0) {
                const int pair = (int)blockIdx.x + (i / n_chain) * (int)gridDim.x;
                tile = G * pair + g;
                row = tile * kUW + r;
                save_tile = SAVE && tile < n_tiles;
                const bool live = row < m_used;
                if (live) { px = xyz[(size_t)row * 3]; py = xyz[(size_t)row * 3 + 1]; pz = xyz[(size_t)row * 3 + 2]; }
                const float xv[3] = {px, py, pz};
                // 8 columns of the encoded input -> one 16-byte chunk of the A tile (c8 is a compile-time constant after unrolling: the
                // channel -> (frequency, phase, dimension) decode of freq_channel folds away)
                auto encode_chunk = [&](const int c8) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float v[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int c = c8 * 8 + j * 2 + e;
                            float val = 0.0f;
                            if (live) {
                                if (c < 63) val = freq_channel(xv, 3, c);
                                else if (c < 76) val = freq_channel(&tval, 1, c - 63);
                            }
                            v[e] = val;
                        }
                        const __half2 h = __floats2half2_rn(v[0], v[1]);
                        w[j] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    const uint4 u = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(a_tile + ((size_t)c8 * kUW + r) * 16) = u;
                    // saved in the tile-image layout (umma.cuh): the shared-memory tile as it is, 512 contiguous bytes per warp store;
                    // dead rows of a live tile are zeros (their values never reach a weight gradient)
                    if (save_tile) *reinterpret_cast<uint4*>(in_buf + (((size_t)tile * (kUK0 / 8) + c8) * kUW + r) * 8) = u;
                };
                if (SM::EW == 1 || half == 0) {
#pragma unroll
                    for (int c8 = 0; c8 < C8N; c8++) encode_chunk(c8);
                } else {
#pragma unroll
                    for (int c8 = C8N; c8 < 2 * C8N; c8++) encode_chunk(c8);
                }
                umma::fence_proxy_async();
                umma::mbar_arrive(bar_aready + g);
            }
            umma::mbar_wait(bar_mma + g, i & 1);
            umma::fence_after_sync();
            if (l < n_layers - 1) {
#pragma unroll 1
                for (int q = half * QN; q < (half + 1) * QN; q++) {
                    uint32_t v[32];
                    umma::tmem_ld32(t_lane + q * 32, v);
                    umma::wait_ld();
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float a = fmaxf(__uint_as_float(v[c4 * 8 + j * 2]), 0.0f);
                            const float b = fmaxf(__uint_as_float(v[c4 * 8 + j * 2 + 1]), 0.0f);
                            const __half2 h = __floats2half2_rn(a, b);
                            w[j] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        const uint4 u = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(a_tile + ((size_t)(q * 4 + c4) * kUW + r) * 16) = u;
                        if (save_tile)
                            *reinterpret_cast<uint4*>(fwd_buf + ((((size_t)l * tiles_cap + tile) * (kUW / 8) + (q * 4 + c4)) * kUW + r) * 8) =
                                (row < m_used) ? u : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                umma::fence_before_sync();
                umma::fence_proxy_async();
                umma::mbar_arrive(bar_aready + g);
            } else if (DENS && l >= n_layers) {
                // ---- sigma head (NeRFNetwork.density, dnerf/network.py:196-208): hidden layers of 64, then [sigma, geo_feat]
                if (l < n_chain - 1) {
#pragma unroll 1
                    for (int q = half * (2 / SM::EW); q < (half + 1) * (2 / SM::EW); q++) {
                        uint32_t v[32];
                        umma::tmem_ld32(t_lane + q * 32, v);
                        umma::wait_ld();
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) {
                            uint32_t w[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const float a = fmaxf(__uint_as_float(v[c4 * 8 + j * 2]), 0.0f);
                                const float b = fmaxf(__uint_as_float(v[c4 * 8 + j * 2 + 1]), 0.0f);
                                const __half2 h = __floats2half2_rn(a, b);
                                w[j] = *reinterpret_cast<const uint32_t*>(&h);
                            }
                            *reinterpret_cast<uint4*>(a_tile + ((size_t)(q * 4 + c4) * kUW + r) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                    umma::fence_before_sync();
                    umma::fence_proxy_async();
                    umma::mbar_arrive(bar_aready + g);
                } else if (half == 0) {
                    uint32_t v[16];
                    umma::tmem_ld16(t_lane, v);
                    umma::wait_ld();
                    umma::fence_before_sync();
                    if (row < m_used) {
                        const float sg = da.density_scale * expf(__half2float(__float2half_rn(__uint_as_float(v[0]))));
                        if (da.sigma) da.sigma[row] = sg;
                        if (da.tmp) da.tmp[da.indices[row]] = sg * da.store_scale;
                    }
                }
            } else if (DENS) {
                // ---- last deformation layer -> x01 -> hash-grid features (16 levels x 2, fp16) as the sigma head's A tile: this
                // thread's row, levels [half * 16 / EW, (half + 1) * 16 / EW), two levels in flight (gridencoder.cu:100-216)
                uint32_t v[16];
                umma::tmem_ld16(t_lane, v);
                umma::wait_ld();
                umma::fence_before_sync();
                float x[3];
                bool live = row < m_used;
                {
                    const float xin[3] = {px, py, pz};
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const float dx = __half2float(__float2half_rn(__uint_as_float(v[d])));
                        const float xp = t_is_zero ? xin[d] : xin[d] + dx;
                        x[d] = (xp + bound) / (2 * bound);
                        if (x[d] < 0 || x[d] > 1) live = false;  // outside [0,1]: zero features (gridencoder.cu:117-131)
                    }
                }
                constexpr int LV = 16 / SM::EW;  // levels per thread
                constexpr int GL = 2;
#pragma unroll 1
                for (int c8 = 0; c8 < LV / 4; c8++) {  // 4 levels = 8 halves = one 16-byte chunk of the A tile
                    uint32_t w[4];
#pragma unroll
                    for (int pass = 0; pass < 4 / GL; pass++) {
                        CellGather<__half, 3, 2> cg[GL];
                        float pos[GL][3], deriv[GL][3];
                        if (live) {
#pragma unroll
                            for (int gl = 0; gl < GL; gl++) {
                                const LevelParams lp = s_lp[half * LV + c8 * 4 + pass * GL + gl];
                                uint32_t pos_grid[3];
                                locate<3>(x, lp, da.align_corners, da.interp, pos[gl], deriv[gl], pos_grid);
                                cg[gl].issue(da.table, da.gridtype, da.align_corners, lp, pos_grid);
                            }
                        }
#pragma unroll
                        for (int gl = 0; gl < GL; gl++) {
                            float res[2] = {0.f, 0.f};
                            if (live) {
                                float val[8][2];
                                cg[gl].resolve(val);
#pragma unroll
                                for (uint32_t idx = 0; idx < 8; idx++) {
                                    float wt = 1;
#pragma unroll
                                    for (uint32_t d = 0; d < 3; d++) wt *= ((idx >> d) & 1u) ? pos[gl][d] : 1 - pos[gl][d];
                                    res[0] += wt * val[idx][0];
                                    res[1] += wt * val[idx][1];
                                }
                            }
                            __half hh[2];
                            Row<__half, 2>::store(hh, res);
                            w[pass * GL + gl] = *reinterpret_cast<const uint32_t*>(hh);
                        }
                    }
                    *reinterpret_cast<uint4*>(a_tile + ((size_t)(half * (LV / 4) + c8) * kUW + r) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                umma::fence_proxy_async();
                umma::mbar_arrive(bar_aready + g);
            } else if (half == 0) {
                uint32_t v[16];
                umma::tmem_ld16(t_lane, v);
                umma::wait_ld();
                umma::fence_before_sync();
                if (row < m_used) {
                    const float xin[3] = {px, py, pz};
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        float dx = __half2float(__float2half_rn(__uint_as_float(v[d])));
                        float xp;
                        if (t_is_zero) {
                            xp = xin[d];
                            if (t0_mode == 1) dx = 0.0f;
                        } else {
                            xp = xin[d] + dx;
                        }
                        deform[(size_t)row * 3 + d] = dx;
                        x01[(size_t)row * 3 + d] = (xp + bound) / (2 * bound);
                    }
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == LOADER) {
        __syncwarp();
        umma::tmem_dealloc(tmem_base, SM::TMEM_COLS);
    }
}


// =====================================================================================================================
// Deformation MLP backward (dgrad chain) on tcgen05: G_{l-1} = (G_l W_l) * ReLU'(h_l), l = n-1 .. 1, starting from
// G_{n-1} = dL/d(dx) = grad_x01 / (2 bound) (zero at t == 0, dnerf/network.py:140-141).  Same CTA organisation as the forward
// kernel; the B operand of step l is W_l^T in the canonical K-major layout (K = out features), pre-packed by
// k_pack_umma_T.  The epilogue multiplies the fp32 accumulator row by the ReLU mask of the saved activation, writes G_{l-1}
// as fp16 both to bwd_buf (the weight-gradient kernel reads it) and into the A-operand tile of the next step.
// No input gradient: xyz and t are leaves without grad (dnerf/network.py:129-134).
// =====================================================================================================================
__host__ __device__ __forceinline__ uint32_t umma_layerT_bytes(const int l, const int n_layers) {
    return (l == n_layers - 1) ? (uint32_t)(kUNLast * kUW * 2) : kTileBytes;  // layers 1 .. n-1 (layer 0 has no dgrad)
}
__host__ __device__ __forceinline__ size_t umma_layerT_offset(const int l, const int n_layers) {
    size_t o = 0;
    for (int i = 1; i < l; i++) o += umma_layerT_bytes(i, n_layers);
    return o;
}

// dst chunk (c, n) = W_l[8c .. 8c+7][n] (rows >= n_real zero): the K-major tile of W_l^T, K = out features
__device__ __forceinline__ void pack_umma_T_body(const PackJobs& jobs, __half* __restrict__ packedT, const int l) {
    const int n_layers = jobs.n_layers;
    const bool last = (l == n_layers - 1);
    const int K = last ? kUNLast : kUW, n_real = last ? 3 : kUW;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K / 8 * kUW) return;
    const int c = i / kUW, n = i - c * kUW;
    __half v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int k = c * 8 + e;
        v[e] = (k < n_real) ? jobs.src[l][(size_t)k * kUW + n] : __float2half_rn(0.0f);
    }
    __half* dst = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(packedT) + umma_layerT_offset(l, n_layers));
    *reinterpret_cast<uint4*>(dst + (size_t)i * 8) = *reinterpret_cast<const uint4*>(v);
}
__global__ void k_pack_umma_T(const PackJobs jobs, __half* __restrict__ packedT) { pack_umma_T_body(jobs, packedT, blockIdx.y + 1); }
__global__ void k_pack_umma_both(const PackJobs jobs, __half* __restrict__ packed, __half* __restrict__ packedT) {
    if ((int)blockIdx.y < jobs.n_layers) pack_umma_body(jobs, packed, blockIdx.y);
    else pack_umma_T_body(jobs, packedT, (int)blockIdx.y - jobs.n_layers + 1);
}

template <int G>
__global__ void __launch_bounds__(UmmaSmem<G>::THREADS, 1) k_deform_backward_umma(const float* __restrict__ grad_x01, const float* __restrict__ time,
                                                                                  const __half* __restrict__ packedT, const int n_layers,
                                                                                  const int M, const int* __restrict__ m_dev, const float bound,
                                                                                  const __half* __restrict__ fwd_buf, __half* __restrict__ bwd_buf,
                                                                                  __half* __restrict__ gout_buf) {
    extern __shared__ __align__(128) unsigned char smem[];
    using SM = UmmaSmem<G>;
    constexpr int ISSUE0 = G * 4 * SM::EW, LOADER = ISSUE0 + G;  // same warp roles as the forward kernel
    unsigned char* s_a = smem + SM::A_OFF;
    unsigned char* s_w = smem + SM::W_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
    uint64_t* bar_full = bars;
    uint64_t* bar_wfree = bars + 3;
    uint64_t* bar_aready = bars + 6;
    uint64_t* bar_mma = bars + 6 + G;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * G);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    const int n_tiles = (m_used + kUW - 1) / kUW;
    const int n_units = (n_tiles + G - 1) / G;
    const int my_units = (n_units > (int)blockIdx.x) ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_steps = n_layers - 1;  // steps per unit: differentiate through matmuls n-1 .. 1
    const int n_items = my_units * n_steps;

    if (tid == 0) {
        for (int i = 0; i < 3; i++) { umma::mbar_init(bar_full + i, 1); umma::mbar_init(bar_wfree + i, G); }
        for (int i = 0; i < G; i++) { umma::mbar_init(bar_aready + i, SM::EPI); umma::mbar_init(bar_mma + i, 1); }
        umma::mbar_fence_init();
    }
    if (warp == LOADER) umma::tmem_alloc(tmem_slot, SM::TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == LOADER) {
        if (lane == 0) {
            for (int i = 0; i < n_items; i++) {
                const int l = n_layers - 1 - (i % n_steps), st = i % kUStages;
                if (i >= kUStages) umma::mbar_wait(bar_wfree + st, ((i / kUStages) - 1) & 1);
                const uint32_t bytes = umma_layerT_bytes(l, n_layers);
                umma::mbar_arrive_expect_tx(bar_full + st, bytes);
                umma::bulk_load(s_w + (size_t)st * kTileBytes, reinterpret_cast<const unsigned char*>(packedT) + umma_layerT_offset(l, n_layers), bytes,
                                bar_full + st);
            }
        }
    } else if (warp >= ISSUE0) {
        const int g = warp - ISSUE0;
        if (lane == 0) {
            const uint32_t a0 = umma::smem_addr(s_a) + g * kTileBytes, w_addr = umma::smem_addr(s_w);
            const uint32_t d = tmem_base + g * kUW;
            const uint32_t idesc = umma::instr_desc_f16(128, kUW);
            for (int i = 0; i < n_items; i++) {
                const int st = i % kUStages;
                const int ksteps = ((i % n_steps) == 0 ? kUNLast : kUW) / 16;
                umma::mbar_wait(bar_full + st, (i / kUStages) & 1);
                umma::mbar_wait(bar_aready + g, i & 1);
                umma::fence_after_sync();
                const uint32_t w0 = w_addr + st * kTileBytes;
                for (int k = 0; k < ksteps; k++) {
                    const uint64_t da = umma::smem_desc(a0 + k * 2 * (kUW * 16), kUW * 16, 128);
                    const uint64_t db = umma::smem_desc(w0 + k * 2 * (kUW * 16), kUW * 16, 128);
                    umma::mma_f16(d, da, db, idesc, k > 0 ? 1u : 0u);
                }
                umma::mma_commit(bar_mma + g);
                umma::mma_commit(bar_wfree + st);
            }
        }
    } else {
        const int g = tid / SM::EPI, te = tid - g * SM::EPI, r = te & 127, half = te >> 7;  // half: which share of the columns (EW == 2)
        constexpr int QN = 4 / SM::EW;
        unsigned char* a_tile = s_a + (size_t)g * kTileBytes;
        const uint32_t t_lane = tmem_base + ((uint32_t)(r & ~31) << 16) + g * kUW;
        // at t == 0 the deformation is replaced by zeros (network.py:140-141): no gradient reaches the net
        const float inv = (time && *time == 0.0f) ? 0.0f : 1.0f / (2 * bound);
        int row = 0, tile = 0;
        bool live_tile = false;
        const size_t tiles_cap = (size_t)(M + kUW - 1) / kUW;
        for (int i = 0; i < n_items; i++) {
            const int s = i % n_steps;
            const int l = n_layers - 1 - s;
            if (s == 0) {
                const int unit = (int)blockIdx.x + (i / n_steps) * (int)gridDim.x;
                tile = G * unit + g;
                row = tile * kUW + r;
                live_tile = tile < n_tiles;
                const bool live = row < m_used;
                float gv[3] = {0.f, 0.f, 0.f};
                if (live) { gv[0] = grad_x01[(size_t)row * 3] * inv; gv[1] = grad_x01[(size_t)row * 3 + 1] * inv; gv[2] = grad_x01[(size_t)row * 3 + 2] * inv; }
                const __half2 h01 = __floats2half2_rn(gv[0], gv[1]), h2 = __floats2half2_rn(gv[2], 0.0f);
                const uint4 u0 = make_uint4(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h2), 0u, 0u);
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                if (half == 0) {
                    *reinterpret_cast<uint4*>(a_tile + ((size_t)0 * kUW + r) * 16) = u0;
                    *reinterpret_cast<uint4*>(a_tile + ((size_t)1 * kUW + r) * 16) = z;
                    if (live_tile) {  // tile-image layout, dead rows zero (u0 is zero for them)
                        *reinterpret_cast<uint4*>(gout_buf + (((size_t)tile * 2 + 0) * kUW + r) * 8) = u0;
                        *reinterpret_cast<uint4*>(gout_buf + (((size_t)tile * 2 + 1) * kUW + r) * 8) = z;
                    }
                }
                umma::fence_proxy_async();
                umma::mbar_arrive(bar_aready + g);
            }
            const bool live = row < m_used;
            // tile images (umma.cuh): chunk (column group c, row r) of tile t of layer j at (((j * tiles_cap + t) * 16 + c) * 128 + r) * 8
            const __half* hsave = fwd_buf + ((((size_t)(l - 1) * tiles_cap + tile) * (kUW / 8)) * kUW + r) * 8;  // h_l: the activation that fed matmul l
            __half* gsave = bwd_buf + ((((size_t)(l - 1) * tiles_cap + tile) * (kUW / 8)) * kUW + r) * 8;        // G_{l-1}
            uint4 hm[4];
            auto load_mask = [&](const int q) {
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++)
                    hm[c4] = live ? __ldg(reinterpret_cast<const uint4*>(hsave + (size_t)(q * 4 + c4) * kUW * 8)) : make_uint4(0u, 0u, 0u, 0u);
            };
            load_mask(half * QN);  // in flight while the MMAs finish
            umma::mbar_wait(bar_mma + g, i & 1);
            umma::fence_after_sync();
            const bool feeds_next = (s + 1 < n_steps);
#pragma unroll 1
            for (int q = half * QN; q < (half + 1) * QN; q++) {
                uint32_t v[32];
                umma::tmem_ld32(t_lane + q * 32, v);
                umma::wait_ld();
                uint4 out[4];
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    const uint32_t hw[4] = {hm[c4].x, hm[c4].y, hm[c4].z, hm[c4].w};
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const __half2 hh = *reinterpret_cast<const __half2*>(&hw[j]);
                        const float a = __low2float(hh) > 0.f ? __uint_as_float(v[c4 * 8 + j * 2]) : 0.f;
                        const float b = __high2float(hh) > 0.f ? __uint_as_float(v[c4 * 8 + j * 2 + 1]) : 0.f;
                        const __half2 p = __floats2half2_rn(a, b);
                        w[j] = *reinterpret_cast<const uint32_t*>(&p);
                    }
                    out[c4] = make_uint4(w[0], w[1], w[2], w[3]);
                }
                if (q + 1 < (half + 1) * QN) load_mask(q + 1);
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    if (feeds_next) *reinterpret_cast<uint4*>(a_tile + ((size_t)(q * 4 + c4) * kUW + r) * 16) = out[c4];
                    if (live_tile) *reinterpret_cast<uint4*>(gsave + (size_t)(q * 4 + c4) * kUW * 8) = out[c4];  // (dead rows: mask 0 -> zeros)
                }
            }
            umma::fence_before_sync();
            if (feeds_next) {
                umma::fence_proxy_async();
                umma::mbar_arrive(bar_aready + g);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == LOADER) {
        __syncwarp();
        umma::tmem_dealloc(tmem_base, SM::TMEM_COLS);
    }
}

}  // namespace seald

using namespace seald;

extern "C" uint64_t seald_field_umma_deform_bytes(int n_layers) {
    if (n_layers < 2) return 0;
    return (uint64_t)umma_layer_offset(n_layers, n_layers);
}

// weights: HOST array of device pointers to the fp16 staging copies (same as seald_field_deform_forward): layer 0
// [128][80], hidden [128][128], last [3][128].  packed: device buffer of seald_field_umma_deform_bytes(n_layers) bytes.
extern "C" int seald_field_umma_pack_deform(const void* const* weights, int n_layers, void* packed, seald_stream_t stream) {
    if (!weights || !packed || n_layers < 2 || n_layers > 12) return SEALD_E_BADARG;
    PackJobs jobs;
    jobs.n_layers = n_layers;
    for (int l = 0; l < n_layers; l++) {
        if (!weights[l]) return SEALD_E_BADARG;
        if ((uintptr_t)weights[l] & 15) return SEALD_E_ALIGN;
        jobs.src[l] = reinterpret_cast<const __half*>(weights[l]);
    }
    k_pack_umma<<<dim3(div_up(kUW / 8 * kUW, 256), n_layers), 256, 0, to_stream(stream)>>>(jobs, reinterpret_cast<__half*>(packed));
    return launch_status();
}

extern "C" uint64_t seald_field_umma_deform_bytes_T(int n_layers) {
    if (n_layers < 2) return 0;
    return (uint64_t)umma_layerT_offset(n_layers, n_layers);
}

// packedT: the transposed tiles W_l^T (layers 1 .. n-1) for the backward kernel, seald_field_umma_deform_bytes_T(n) bytes
extern "C" int seald_field_umma_pack_deform_T(const void* const* weights, int n_layers, void* packedT, seald_stream_t stream) {
    if (!weights || !packedT || n_layers < 2 || n_layers > 12) return SEALD_E_BADARG;
    PackJobs jobs;
    jobs.n_layers = n_layers;
    for (int l = 0; l < n_layers; l++) {
        if (!weights[l]) return SEALD_E_BADARG;
        jobs.src[l] = reinterpret_cast<const __half*>(weights[l]);
    }
    k_pack_umma_T<<<dim3(div_up(kUW / 8 * kUW, 256), n_layers - 1), 256, 0, to_stream(stream)>>>(jobs, reinterpret_cast<__half*>(packedT));
    return launch_status();
}

// both layouts in one launch pair-free call (the optimiser refreshes them every step)
extern "C" int seald_field_umma_pack_deform_both(const void* const* weights, int n_layers, void* packed, void* packedT, seald_stream_t stream) {
    if (!weights || !packed || !packedT || n_layers < 2 || n_layers > 12) return SEALD_E_BADARG;
    PackJobs jobs;
    jobs.n_layers = n_layers;
    for (int l = 0; l < n_layers; l++) {
        if (!weights[l]) return SEALD_E_BADARG;
        if ((uintptr_t)weights[l] & 15) return SEALD_E_ALIGN;
        jobs.src[l] = reinterpret_cast<const __half*>(weights[l]);
    }
    k_pack_umma_both<<<dim3(div_up(kUW / 8 * kUW, 256), 2 * n_layers - 1), 256, 0, to_stream(stream)>>>(jobs, reinterpret_cast<__half*>(packed),
                                                                                                      reinterpret_cast<__half*>(packedT));
    return launch_status();
}

extern "C" int seald_field_deform_backward_umma(const float* grad_x01, const float* time_dev, const void* packedT, int n_layers, uint32_t M,
                                                const int32_t* m_dev, float bound, const void* fwd_buf, void* bwd_buf, void* gout_buf,
                                                seald_stream_t stream) {
    if (M == 0) return 0;
    if (!grad_x01 || !packedT || !fwd_buf || !bwd_buf || !gout_buf || n_layers < 2 || n_layers > 12) return SEALD_E_BADARG;
    if (((uintptr_t)packedT & 15) != 0) return SEALD_E_ALIGN;
    const uint32_t n_tiles = div_up(M, (uint32_t)kUW);
    cudaStream_t st = to_stream(stream);
    auto launch = [&](auto kernel, const int G, const size_t smem, const int threads) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const uint32_t units = div_up(n_tiles, (uint32_t)G);
        const uint32_t grid = units < (uint32_t)SEALD_NUM_SMS ? units : (uint32_t)SEALD_NUM_SMS;
        kernel<<<grid, threads, smem, st>>>(grad_x01, time_dev, (const __half*)packedT, n_layers, (int)M, m_dev, bound, (const __half*)fwd_buf,
                                            (__half*)bwd_buf, (__half*)gout_buf);
        return 0;
    };
    static const int forced_g = getenv("SEALD_UMMA_G") ? atoi(getenv("SEALD_UMMA_G")) : 0;  // measurement switch
    const bool big = forced_g ? forced_g == 4 : n_tiles >= 4u * SEALD_NUM_SMS;
    const int rc = big ? launch(k_deform_backward_umma<4>, 4, UmmaSmem<4>::BYTES, UmmaSmem<4>::THREADS)
                       : launch(k_deform_backward_umma<2>, 2, UmmaSmem<2>::BYTES, UmmaSmem<2>::THREADS);
    if (rc) return rc;
    return launch_status();
}

extern "C" int seald_field_deform_forward_umma(const float* xyz, const float* time_dev, const void* packed, int n_layers, uint32_t M,
                                               const int32_t* m_dev, float bound, int t0_mode, float* deform, float* x01, void* in_buf,
                                               void* fwd_buf, seald_stream_t stream) {
    if (M == 0) return 0;
    if (!xyz || !time_dev || !packed || !deform || !x01 || n_layers < 2 || n_layers > 12) return SEALD_E_BADARG;
    if ((in_buf == nullptr) != (fwd_buf == nullptr)) return SEALD_E_BADARG;
    if (((uintptr_t)packed & 15) != 0) return SEALD_E_ALIGN;
    const uint32_t n_tiles = div_up(M, (uint32_t)kUW);
    cudaStream_t st = to_stream(stream);
    auto launch = [&](auto kernel, const int G, const size_t smem, const int threads) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const uint32_t units = div_up(n_tiles, (uint32_t)G);
        const uint32_t grid = units < (uint32_t)SEALD_NUM_SMS ? units : (uint32_t)SEALD_NUM_SMS;
        kernel<<<grid, threads, smem, st>>>(xyz, time_dev, (const __half*)packed, n_layers, (int)M, m_dev, bound, t0_mode, deform, x01,
                                            (__half*)in_buf, (__half*)fwd_buf, DensityArgs{});
        return 0;
    };
    // enough tiles to give every SM four at a time: G = 4 (tensor pipe saturated); otherwise spread over more SMs with G = 2
    static const int forced_g = getenv("SEALD_UMMA_G") ? atoi(getenv("SEALD_UMMA_G")) : 0;  // measurement switch
    const bool big = forced_g ? forced_g == 4 : n_tiles >= 4u * SEALD_NUM_SMS;
    int rc;
    if (fwd_buf) {
        rc = big ? launch(k_deform_forward_umma<true, 4, false>, 4, UmmaSmem<4>::BYTES, UmmaSmem<4>::THREADS)
                 : launch(k_deform_forward_umma<true, 2, false>, 2, UmmaSmem<2>::BYTES, UmmaSmem<2>::THREADS);
    } else {
        rc = big ? launch(k_deform_forward_umma<false, 4, false>, 4, UmmaSmem<4>::BYTES, UmmaSmem<4>::THREADS)
                 : launch(k_deform_forward_umma<false, 2, false>, 2, UmmaSmem<2>::BYTES, UmmaSmem<2>::THREADS);
    }
    if (rc) return rc;
    return launch_status();
}

// ---- density query in ONE launch: deformation net -> hash grid -> sigma head, nothing but xyz in and sigma out touches HBM ----
extern "C" uint64_t seald_field_umma_sigma_bytes(int n_sigma) {
    if (n_sigma < 2) return 0;
    return (uint64_t)sigma_layer_offset(n_sigma, n_sigma);
}

// weights: HOST array of device pointers to the fp16 staging copies of the sigma head (layer 0 [64][32], hidden [64][64], last [16][64])
extern "C" int seald_field_umma_pack_sigma(const void* const* weights, int n_sigma, void* packed, seald_stream_t stream) {
    if (!weights || !packed || n_sigma < 2 || n_sigma > 12) return SEALD_E_BADARG;
    PackJobs jobs;
    jobs.n_layers = n_sigma;
    for (int l = 0; l < n_sigma; l++) {
        if (!weights[l]) return SEALD_E_BADARG;
        if ((uintptr_t)weights[l] & 15) return SEALD_E_ALIGN;
        jobs.src[l] = reinterpret_cast<const __half*>(weights[l]);
    }
    k_pack_umma_sigma<<<dim3(div_up(kSW / 8 * kSW, 256), n_sigma), 256, 0, to_stream(stream)>>>(jobs, reinterpret_cast<__half*>(packed));
    return launch_status();
}

extern "C" int seald_field_density_umma(const float* xyz, const float* time_dev, const void* packed, int n_layers, const void* packed_sigma,
                                        int n_sigma, const void* table, const int32_t* offsets, uint32_t D, uint32_t C, uint32_t L, float S,
                                        uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, uint32_t M, const int32_t* m_dev,
                                        float bound, int t0_mode, float density_scale, float* sigma, const int32_t* indices, float store_scale,
                                        float* tmp, seald_stream_t stream) {
    if (M == 0) return 0;
    if (!xyz || !time_dev || !packed || !packed_sigma || !table || !offsets || n_layers < 2 || n_layers > 12 || n_sigma < 2 || n_sigma > 12)
        return SEALD_E_BADARG;
    if (!sigma && !tmp) return SEALD_E_BADARG;
    if (tmp && !indices) return SEALD_E_BADARG;
    if (D != 3 || C != 2 || L != 16 || gridtype > 1 || interp > 1) return SEALD_E_UNSUPPORTED;
    if (((uintptr_t)packed & 15) || ((uintptr_t)packed_sigma & 15) || ((uintptr_t)table & 15)) return SEALD_E_ALIGN;
    const uint32_t n_tiles = div_up(M, (uint32_t)kUW);
    cudaStream_t st = to_stream(stream);
    const DensityArgs da{(const __half*)packed_sigma, n_sigma, (const __half*)table, offsets, S, H, gridtype, interp, align_corners != 0,
                         density_scale, sigma, indices, store_scale, tmp};
    auto launch = [&](auto kernel, const int G, const size_t smem, const int threads) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const uint32_t units = div_up(n_tiles, (uint32_t)G);
        const uint32_t grid = units < (uint32_t)SEALD_NUM_SMS ? units : (uint32_t)SEALD_NUM_SMS;
        kernel<<<grid, threads, smem, st>>>(xyz, time_dev, (const __half*)packed, n_layers, (int)M, m_dev, bound, t0_mode, nullptr, nullptr, nullptr,
                                            nullptr, da);
        return 0;
    };
    static const int forced_g = getenv("SEALD_UMMA_G") ? atoi(getenv("SEALD_UMMA_G")) : 0;  // measurement switch
    const bool big = forced_g ? forced_g == 4 : n_tiles >= 4u * SEALD_NUM_SMS;
    const int rc = big ? launch(k_deform_forward_umma<false, 4, true>, 4, UmmaSmem<4>::BYTES_DENS, UmmaSmem<4>::THREADS)
                       : launch(k_deform_forward_umma<false, 2, true>, 2, UmmaSmem<2>::BYTES_DENS, UmmaSmem<2>::THREADS);
    if (rc) return rc;
    return launch_status();
}
