// mlp.cuh — fused bias-free MLP building blocks (fp16 operands, fp32 accumulation) shared by the FFMLP-compatible
// kernels (mlp.cu) and the fused D-NeRF field kernels (field.cu).
//
// One CTA (8 warps) owns a tile of 128 rows (samples); each warp owns 16 rows end to end.  Activations live in
// REGISTERS between layers: the m16n8k16 accumulator fragment of layer l is exactly the A-operand fragment of
// layer l+1 after a ReLU and an fp16 pack, so hidden activations never touch shared memory.  Weights are streamed
// layer by layer into a double-buffered shared-memory stage with cp.async (32 KiB per 128x128 layer) and read with
// ldmatrix; the first layer's input tile and the last layer's (<=16 wide) output tile go through shared memory so
// that input/output stages (frequency encoding, SH, exp / sigmoid epilogues ...) are ordinary per-row code.
//
// Reference semantics: ffmlp/src/ffmlp.cu:332-410 (fused forward), :411-518 (fused backward); unlike the
// reference (fp16 wmma accumulators, ffmlp.cu:68) accumulation is fp32.
#pragma once
#include "common.cuh"

namespace seald {

constexpr int kTileRows = 128;   // rows per CTA tile
constexpr int kMlpThreads = 256; // 8 warps x 16 rows
constexpr int kPad = 8;          // smem row padding in halves (16 bytes): conflict-free ldmatrix
constexpr int kMaxLayers = 12;

struct MlpWeights {
    const __half* w[kMaxLayers];  // layer l: [N_l][K_l] row-major fp16 (nn.Linear.weight layout)
    int n_layers;                 // number of matmuls (>= 2)
    int k0;                       // input width, padded to a multiple of 16 (columns beyond the real width are zero)
    int k0_ld;                    // leading dimension (in halves) of layer 0's weight rows
    int n_out;                    // real rows of the last layer's weight (<= 16); rows beyond are treated as zero
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
// d += a(16x16, row) * b(16x8, col), fp16 in / fp32 accumulate
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint32_t pack_half2(const float a, const float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Stage a weight matrix [N][K] (row-major, leading dimension ld halves) into smem rows of stride (Kpad + kPad).
// Rows >= n_real are zero-filled.  K must be a multiple of 8.  All threads of the CTA participate.
__device__ __forceinline__ void stage_weights(__half* s_w, const int s_stride, const __half* __restrict__ w, const int N, const int n_real,
                                              const int K, const int ld) {
    const int chunks_per_row = K / 8;
    for (int i = threadIdx.x; i < N * chunks_per_row; i += blockDim.x) {
        const int r = i / chunks_per_row, c = i - r * chunks_per_row;
        __half* dst = s_w + r * s_stride + c * 8;
        if (r < n_real) {
            cp_async16(dst, w + (size_t)r * ld + c * 8);
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
    }
}

// ===================================================================================================
// Forward over one 128-row tile.
//   s_in  : [128][K0MAX + kPad] halves, input tile (filled by the caller, columns [0,k0))
//   s_w   : 2 x [WIDTH][WMAX + kPad] halves, WMAX = max(K0MAX, WIDTH)
//   s_out : [128][kOutStride] floats, last layer pre-activation (16 columns)
//   fwd_buf (optional): [n_layers-1][M][WIDTH] halves, post-ReLU activations for backward
// Must be called by all 256 threads; ends with a __syncthreads() so s_out is readable by everyone.
// ===================================================================================================
constexpr int kOutStride = 17;

template <int WIDTH, int K0MAX>
struct MlpSmem {
    static constexpr int WMAX = (K0MAX > WIDTH ? K0MAX : WIDTH);
    static constexpr int IN_STRIDE = K0MAX + kPad;
    static constexpr int W_STRIDE = WMAX + kPad;
    static constexpr int IN_HALVES = kTileRows * IN_STRIDE;
    static constexpr int W_HALVES = WIDTH * W_STRIDE;  // one stage
    static constexpr size_t BYTES = (size_t)(IN_HALVES + 2 * W_HALVES) * sizeof(__half) + (size_t)kTileRows * kOutStride * sizeof(float);
};

template <int WIDTH, int K0MAX, bool SAVE>
__device__ __forceinline__ void mlp_forward_tile(const MlpWeights& mw, const __half* s_in, __half* s_w, float* s_out, __half* __restrict__ fwd_buf,
                                                 const int M, const int row0) {
    using SM = MlpSmem<WIDTH, K0MAX>;
    constexpr int NT = WIDTH / 8;   // n8 tiles of a hidden layer
    constexpr int KT = WIDTH / 16;  // k16 tiles of a hidden layer
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wrow = warp * 16;  // first tile row of this warp
    const int n_layers = mw.n_layers;

    // stage layer 0
    stage_weights(s_w, SM::W_STRIDE, mw.w[0], WIDTH, WIDTH, mw.k0, mw.k0_ld);
    cp_async_commit();

    uint32_t areg[KT][4];
    float acc[NT][4];

    for (int l = 0; l < n_layers - 1; l++) {
        __half* wcur = s_w + (l & 1) * SM::W_HALVES;
        __half* wnext = s_w + ((l + 1) & 1) * SM::W_HALVES;
        cp_async_wait<0>();
        __syncthreads();  // layer l's weights visible; everyone is done with the other stage
        // prefetch the next layer
        if (l + 1 < n_layers - 1) {
            stage_weights(wnext, SM::W_STRIDE, mw.w[l + 1], WIDTH, WIDTH, WIDTH, WIDTH);
        } else {
            stage_weights(wnext, SM::W_STRIDE, mw.w[l + 1], 16, mw.n_out, WIDTH, WIDTH);
        }
        cp_async_commit();

#pragma unroll
        for (int n = 0; n < NT; n++) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f; }

        if (l == 0) {
            const int kt0 = mw.k0 / 16;
            for (int kk = 0; kk < kt0; kk++) {
                uint32_t a[4];
                ldmatrix_x4(a, s_in + (wrow + (lane & 7) + 8 * ((lane >> 3) & 1)) * SM::IN_STRIDE + kk * 16 + 8 * (lane >> 4));
#pragma unroll
                for (int nn = 0; nn < NT; nn += 2) {
                    uint32_t b[4];
                    ldmatrix_x4(b, wcur + (8 * nn + (lane & 7) + 8 * (lane >> 4)) * SM::W_STRIDE + kk * 16 + 8 * ((lane >> 3) & 1));
                    mma_16816(acc[nn], a, b[0], b[1]);
                    mma_16816(acc[nn + 1], a, b[2], b[3]);
                }
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < KT; kk++) {
#pragma unroll
                for (int nn = 0; nn < NT; nn += 2) {
                    uint32_t b[4];
                    ldmatrix_x4(b, wcur + (8 * nn + (lane & 7) + 8 * (lane >> 4)) * SM::W_STRIDE + kk * 16 + 8 * ((lane >> 3) & 1));
                    mma_16816(acc[nn], areg[kk], b[0], b[1]);
                    mma_16816(acc[nn + 1], areg[kk], b[2], b[3]);
                }
            }
        }
        // ReLU, pack to the next layer's A fragments, optionally save for backward
#pragma unroll
        for (int kk = 0; kk < KT; kk++) {
            areg[kk][0] = pack_half2(fmaxf(acc[2 * kk][0], 0.f), fmaxf(acc[2 * kk][1], 0.f));
            areg[kk][1] = pack_half2(fmaxf(acc[2 * kk][2], 0.f), fmaxf(acc[2 * kk][3], 0.f));
            areg[kk][2] = pack_half2(fmaxf(acc[2 * kk + 1][0], 0.f), fmaxf(acc[2 * kk + 1][1], 0.f));
            areg[kk][3] = pack_half2(fmaxf(acc[2 * kk + 1][2], 0.f), fmaxf(acc[2 * kk + 1][3], 0.f));
        }
        if (SAVE) {
            __half* base = fwd_buf + (size_t)l * M * WIDTH;
            const int r0 = row0 + wrow + g, r1 = r0 + 8;
#pragma unroll
            for (int kk = 0; kk < KT; kk++) {
                if (r0 < M) {
                    *reinterpret_cast<uint32_t*>(base + (size_t)r0 * WIDTH + kk * 16 + 2 * t) = areg[kk][0];
                    *reinterpret_cast<uint32_t*>(base + (size_t)r0 * WIDTH + kk * 16 + 8 + 2 * t) = areg[kk][2];
                }
                if (r1 < M) {
                    *reinterpret_cast<uint32_t*>(base + (size_t)r1 * WIDTH + kk * 16 + 2 * t) = areg[kk][1];
                    *reinterpret_cast<uint32_t*>(base + (size_t)r1 * WIDTH + kk * 16 + 8 + 2 * t) = areg[kk][3];
                }
            }
        }
    }
    // last layer: N = 16 (two n8 tiles)
    {
        const int l = n_layers - 1;
        __half* wcur = s_w + (l & 1) * SM::W_HALVES;
        cp_async_wait<0>();
        __syncthreads();
        float o[2][4];
#pragma unroll
        for (int n = 0; n < 2; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f; }
#pragma unroll
        for (int kk = 0; kk < KT; kk++) {
            uint32_t b[4];
            ldmatrix_x4(b, wcur + ((lane & 7) + 8 * (lane >> 4)) * SM::W_STRIDE + kk * 16 + 8 * ((lane >> 3) & 1));
            mma_16816(o[0], areg[kk], b[0], b[1]);
            mma_16816(o[1], areg[kk], b[2], b[3]);
        }
#pragma unroll
        for (int n = 0; n < 2; n++) {
            s_out[(wrow + g) * kOutStride + 8 * n + 2 * t] = o[n][0];
            s_out[(wrow + g) * kOutStride + 8 * n + 2 * t + 1] = o[n][1];
            s_out[(wrow + g + 8) * kOutStride + 8 * n + 2 * t] = o[n][2];
            s_out[(wrow + g + 8) * kOutStride + 8 * n + 2 * t + 1] = o[n][3];
        }
    }
    __syncthreads();
}

// ===================================================================================================
// Backward over one 128-row tile.
//   s_g   : [128][16 + kPad] halves, dL/d(out) tile (filled by the caller; columns >= n_out must be zero)
//   s_w   : same double-buffered stage as forward
//   fwd_buf : [n_layers-1][M_ld][WIDTH] saved activations;  bwd_buf : [n_layers-1][M_ld][WIDTH] dL/d(pre-activation)
//   M_ld = allocated rows per layer slab, M = rows that are live (rows >= M are neither read nor written)
//   dinput (optional): fp16, row (r - row0 + din_row0) receives dL/d(input) columns [0, k0) of tile row r; may point
//                      to global ([M][din_ld], din_row0 = row0) or to a shared-memory tile (din_row0 = 0)
// Ends with __syncthreads().
// ===================================================================================================
constexpr int kGStride = 16 + kPad;

// Element (row r, column c) of a saved [rows][W] fp16 tensor in the 128-row TILE-IMAGE layout (umma.cuh / wgrad_umma.cu):
// [tile][W / 8][128 rows][8 halves] — what the tcgen05 weight-gradient kernel fetches with one bulk copy per tile.
template <int W>
__device__ __forceinline__ size_t tile_img_off(const int r, const int c) {
    return (((size_t)(r >> 7) * (W / 8) + (c >> 3)) * 128 + (r & 127)) * 8 + (c & 7);
}

// TILED: fwd_buf / bwd_buf are tile images (M_ld = rows rounded up to 128); every row of a live tile is written (dead rows carry
// zeros: their g_out and saved activations are zero), so the weight-gradient kernel needs no row mask.
template <int WIDTH, int K0MAX, bool TILED = false>
__device__ __forceinline__ void mlp_backward_tile(const MlpWeights& mw, const __half* s_g, __half* s_w, const __half* __restrict__ fwd_buf,
                                                  __half* __restrict__ bwd_buf, __half* dinput, const int din_ld, const int din_row0,
                                                  const int M_ld, const int M, const int row0) {
    using SM = MlpSmem<WIDTH, K0MAX>;
    constexpr int NT = WIDTH / 8;
    constexpr int KT = WIDTH / 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wrow = warp * 16;
    const int n_layers = mw.n_layers;
    const int r0 = row0 + wrow + g, r1 = r0 + 8;

    // stage the last layer's weight [16][WIDTH]
    stage_weights(s_w, SM::W_STRIDE, mw.w[n_layers - 1], 16, mw.n_out, WIDTH, WIDTH);
    cp_async_commit();

    uint32_t greg[KT][4];
    float acc[NT][4];

    // step s = 0 .. n_layers-2 : produces G of hidden layer (n_layers-2-s)
    for (int s = 0; s < n_layers - 1; s++) {
        const int l = n_layers - 1 - s;  // the matmul being differentiated through (its weight is staged in s & 1)
        __half* wcur = s_w + (s & 1) * SM::W_HALVES;
        __half* wnext = s_w + ((s + 1) & 1) * SM::W_HALVES;
        cp_async_wait<0>();
        __syncthreads();
        // prefetch the weight of matmul l-1 (needed next step, or for dinput when l-1 == 0)
        if (l - 1 >= 1) {
            stage_weights(wnext, SM::W_STRIDE, mw.w[l - 1], WIDTH, WIDTH, WIDTH, WIDTH);
        } else if (dinput) {
            stage_weights(wnext, SM::W_STRIDE, mw.w[0], WIDTH, WIDTH, mw.k0, mw.k0_ld);
        }
        cp_async_commit();

#pragma unroll
        for (int n = 0; n < NT; n++) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f; }

        if (s == 0) {
            // dh = g_out[16 rows x 16] * W_last[16][WIDTH]
            uint32_t a[4];
            ldmatrix_x4(a, s_g + (wrow + (lane & 7) + 8 * ((lane >> 3) & 1)) * kGStride + 8 * (lane >> 4));
#pragma unroll
            for (int nn = 0; nn < NT; nn += 2) {
                uint32_t b[4];
                ldmatrix_x4_trans(b, wcur + ((lane & 7) + 8 * ((lane >> 3) & 1)) * SM::W_STRIDE + 8 * nn + 8 * (lane >> 4));
                mma_16816(acc[nn], a, b[0], b[1]);
                mma_16816(acc[nn + 1], a, b[2], b[3]);
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < KT; kk++) {
#pragma unroll
                for (int nn = 0; nn < NT; nn += 2) {
                    uint32_t b[4];
                    ldmatrix_x4_trans(b, wcur + (16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)) * SM::W_STRIDE + 8 * nn + 8 * (lane >> 4));
                    mma_16816(acc[nn], greg[kk], b[0], b[1]);
                    mma_16816(acc[nn + 1], greg[kk], b[2], b[3]);
                }
            }
        }
        // dh_l = G_l W_l ; G_{l-1} = dh_l * ReLU'(h_l).  fwd_buf[j] holds h_{j+1}; bwd_buf[j] receives G_j = dL/dz_j.
        const __half* hsave = fwd_buf + (size_t)(l - 1) * M_ld * WIDTH;
        __half* gsave = bwd_buf + (size_t)(l - 1) * M_ld * WIDTH;
#pragma unroll
        for (int kk = 0; kk < KT; kk++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {  // the two n8 tiles of this k16 group
                const int col = kk * 16 + 8 * h + 2 * t;
                uint32_t m0 = 0, m1 = 0;
                if constexpr (TILED) {
                    m0 = *reinterpret_cast<const uint32_t*>(hsave + tile_img_off<WIDTH>(r0, col));
                    m1 = *reinterpret_cast<const uint32_t*>(hsave + tile_img_off<WIDTH>(r1, col));
                } else {
                    if (r0 < M) m0 = *reinterpret_cast<const uint32_t*>(hsave + (size_t)r0 * WIDTH + col);
                    if (r1 < M) m1 = *reinterpret_cast<const uint32_t*>(hsave + (size_t)r1 * WIDTH + col);
                }
                const __half2 h0 = *reinterpret_cast<const __half2*>(&m0), h1 = *reinterpret_cast<const __half2*>(&m1);
                const float* c = acc[2 * kk + h];
                const uint32_t p0 = pack_half2(__low2float(h0) > 0.f ? c[0] : 0.f, __high2float(h0) > 0.f ? c[1] : 0.f);
                const uint32_t p1 = pack_half2(__low2float(h1) > 0.f ? c[2] : 0.f, __high2float(h1) > 0.f ? c[3] : 0.f);
                greg[kk][2 * h] = p0;
                greg[kk][2 * h + 1] = p1;
                if constexpr (TILED) {
                    *reinterpret_cast<uint32_t*>(gsave + tile_img_off<WIDTH>(r0, col)) = p0;
                    *reinterpret_cast<uint32_t*>(gsave + tile_img_off<WIDTH>(r1, col)) = p1;
                } else {
                    if (r0 < M) *reinterpret_cast<uint32_t*>(gsave + (size_t)r0 * WIDTH + col) = p0;
                    if (r1 < M) *reinterpret_cast<uint32_t*>(gsave + (size_t)r1 * WIDTH + col) = p1;
                }
            }
        }
    }
    // dinput = G_1 * W_0   ([16 x WIDTH] * [WIDTH][k0])
    if (dinput) {
        const int s = n_layers - 1;
        __half* wcur = s_w + (s & 1) * SM::W_HALVES;
        cp_async_wait<0>();
        __syncthreads();
        const int nt0 = mw.k0 / 8;
        for (int nn = 0; nn < nt0; nn += 2) {
            float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kk = 0; kk < KT; kk++) {
                uint32_t b[4];
                ldmatrix_x4_trans(b, wcur + (16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)) * SM::W_STRIDE + 8 * nn + 8 * (lane >> 4));
                mma_16816(d0, greg[kk], b[0], b[1]);
                mma_16816(d1, greg[kk], b[2], b[3]);
            }
            const int c0 = 8 * nn + 2 * t, c1 = c0 + 8;
            const size_t q0 = (size_t)(r0 - row0 + din_row0), q1 = q0 + 8;  // dinput may be a shared-memory tile (din_row0 = 0)
            if (r0 < M) {
                *reinterpret_cast<uint32_t*>(dinput + q0 * din_ld + c0) = pack_half2(d0[0], d0[1]);
                *reinterpret_cast<uint32_t*>(dinput + q0 * din_ld + c1) = pack_half2(d1[0], d1[1]);
            }
            if (r1 < M) {
                *reinterpret_cast<uint32_t*>(dinput + q1 * din_ld + c0) = pack_half2(d0[2], d0[3]);
                *reinterpret_cast<uint32_t*>(dinput + q1 * din_ld + c1) = pack_half2(d1[2], d1[3]);
            }
        }
    } else {
        cp_async_wait<0>();
    }
    __syncthreads();
}

}  // namespace seald
