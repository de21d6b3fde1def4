// field.cu — the D-NeRF field as fused tensor-core kernels (reference: dnerf/network.py:123-169 forward, :171-208
// density; SealDNeRF/network.py:125-212 is the same computation).
//
//   deform   : freq(x) (63) || freq(t) (13)  ->  8 x Linear(128, bias=False) + ReLU  ->  dx (3);  x' = x + dx
//   canonical: GridEncoder(x')  (grid_encoder.cu, separate launch: it is a gather, not a contraction)
//   sigma    : 32 -> 64 -> 16 ; sigma = exp(h0) (fp32), geo = h[1:16]
//   colour   : SH4(d) (16) || geo (15) -> 64 -> 64 -> 3 -> sigmoid
//
// The frequency / SH encodings are computed in the MLP input stage (they never exist in HBM unless saved for the
// weight-gradient GEMMs), hidden activations stay in registers (mlp.cuh), trunc_exp / sigmoid are epilogues.
// Numerics follow the reference under autocast: fp16 operands and layer outputs, fp32 accumulation (cuBLAS),
// exp in fp32 on the fp16-rounded pre-activation (activation.py:5-17), sigmoid on the fp16 output.
#include <cstdlib>
#include <string>

#include "encoders.cuh"
#include "grid_device.cuh"
#include "mlp.cuh"

namespace seald {

constexpr int kDeformW = 128, kDeformK0 = 80;  // 63 + 13 = 76 real inputs, padded to 80
constexpr int kHeadW = 64, kHeadK0 = 32;       // sigma: 32 grid features; colour: 16 SH + 15 geo + 1 pad

using DeformSmem = MlpSmem<kDeformW, kDeformK0>;
using HeadSmem = MlpSmem<kHeadW, kHeadK0>;

__device__ __forceinline__ float half_round(const float v) { return __half2float(__float2half_rn(v)); }

// ---------------------------------------------------------------------------------------------------
// deformation network forward
//   xyz [M,3] fp32, time: device pointer to one float
//   t0_mode: what happens when *time == 0 — 1: forward() semantics (dx := 0, network.py:140-141),
//            2: density() semantics (x' = x but dx is still reported, network.py:188-190)
//   outputs: deform [M,3] fp32 (fp16-rounded values), x01 [M,3] fp32 = (x' + bound) / (2 bound)  (grid.py:149)
//   in_buf [M,80] fp16 (optional, training): the encoded input, for the first layer's weight gradient
// ---------------------------------------------------------------------------------------------------
template <bool SAVE>
__global__ void __launch_bounds__(kMlpThreads, 2) k_deform_forward(const float* __restrict__ xyz, const float* __restrict__ time,
                                                                   const MlpWeights mw, const int M, const int* __restrict__ m_dev,
                                                                   const float bound, const int t0_mode, float* __restrict__ deform,
                                                                   float* __restrict__ x01, __half* __restrict__ in_buf,
                                                                   __half* __restrict__ fwd_buf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_in = reinterpret_cast<__half*>(smem_raw);
    __half* s_w = s_in + DeformSmem::IN_HALVES;
    float* s_out = reinterpret_cast<float*>(s_w + 2 * DeformSmem::W_HALVES);

    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;  // rows beyond the live sample count are skipped
    const float tval = *time;
    const bool t_is_zero = (tval == 0.0f);
    const int n_tiles = (m_used + kTileRows - 1) / kTileRows;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows;
        // ---- input stage: 80 columns per row, 128 rows -> each thread fills 40 entries
        for (int i = threadIdx.x; i < kTileRows * kDeformK0; i += kMlpThreads) {
            const int r = i / kDeformK0, c = i - r * kDeformK0;
            const int row = row0 + r;
            float v = 0.0f;
            if (row < m_used) {
                if (c < 63) {
                    const float* x = xyz + (size_t)row * 3;
                    const float xv[3] = {x[0], x[1], x[2]};
                    v = freq_channel(xv, 3, c);
                } else if (c < 76) {
                    v = freq_channel(&tval, 1, c - 63);
                }
            }
            const __half h = __float2half_rn(v);
            s_in[r * DeformSmem::IN_STRIDE + c] = h;
            if (SAVE && row < m_used) in_buf[(size_t)row * kDeformK0 + c] = h;
        }
        mlp_forward_tile<kDeformW, kDeformK0, SAVE>(mw, s_in, s_w, s_out, fwd_buf, M, row0);
        // ---- epilogue: dx = fp16(z[0:3]); x' = x + dx; x01 = (x' + bound) / (2 bound)
        if (threadIdx.x < kTileRows) {
            const int row = row0 + threadIdx.x;
            if (row < m_used) {
                const float* z = s_out + threadIdx.x * kOutStride;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    float dx = half_round(z[d]);
                    const float x = xyz[(size_t)row * 3 + d];
                    float xp;
                    if (t_is_zero) {
                        xp = x;
                        if (t0_mode == 1) dx = 0.0f;
                    } else {
                        xp = x + dx;
                    }
                    deform[(size_t)row * 3 + d] = dx;
                    x01[(size_t)row * 3 + d] = (xp + bound) / (2 * bound);
                }
            }
        }
        __syncthreads();
    }
}

// deformation network backward: g_out = grad_x01 / (2 bound) (fp16), no input gradient (xyz / t are leaves without grad)
__global__ void __launch_bounds__(kMlpThreads, 2) k_deform_backward(const float* __restrict__ grad_x01, const float* __restrict__ time,
                                                                    const MlpWeights mw, const int M,
                                                                    const int* __restrict__ m_dev, const float bound,
                                                                    const __half* __restrict__ fwd_buf, __half* __restrict__ bwd_buf,
                                                                    __half* __restrict__ gout_buf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_g = reinterpret_cast<__half*>(smem_raw);  // [128][24]
    __half* s_w = s_g + DeformSmem::IN_HALVES;           // keep the forward layout (s_in region is large enough)
    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    const int n_tiles = (m_used + kTileRows - 1) / kTileRows;
    // at t == 0 the deformation is replaced by zeros (network.py:140-141): no gradient reaches the net
    const float inv = (time && *time == 0.0f) ? 0.0f : 1.0f / (2 * bound);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows;
        for (int i = threadIdx.x; i < kTileRows * 16; i += kMlpThreads) {
            const int r = i >> 4, c = i & 15;
            const int row = row0 + r;
            float v = 0.0f;
            if (row < m_used && c < 3) v = grad_x01[(size_t)row * 3 + c] * inv;
            const __half h = __float2half_rn(v);
            s_g[r * kGStride + c] = h;
            if (row < m_used) gout_buf[(size_t)row * 16 + c] = h;
        }
        mlp_backward_tile<kDeformW, kDeformK0>(mw, s_g, s_w, fwd_buf, bwd_buf, nullptr, 0, 0, M, m_used, row0);
    }
}

// ---------------------------------------------------------------------------------------------------
// sigma + colour heads forward
//   feat [M,32] fp16 (grid features), dirs [M,3] fp32
//   outputs: sigma [M] fp32 (= density_scale * exp(fp16(h0))), rgb [M,3] fp32 (fp16-rounded sigmoid)
//   training buffers: hs [M,16] fp16 (sigma-net output), cin [M,32] fp16 (colour-net input),
//                     fwd_s [1][M][64], fwd_c [2][M][64]
// ---------------------------------------------------------------------------------------------------
template <bool SAVE>
__global__ void __launch_bounds__(kMlpThreads, 2) k_heads_forward(const __half* __restrict__ feat, const float* __restrict__ dirs,
                                                                  const MlpWeights mw_s, const MlpWeights mw_c, const int M,
                                                                  const int* __restrict__ m_dev, const float density_scale,
                                                                  float* __restrict__ sigma, float* __restrict__ rgb,
                                                                  __half* __restrict__ hs, __half* __restrict__ cin,
                                                                  __half* __restrict__ fwd_s, __half* __restrict__ fwd_c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_in = reinterpret_cast<__half*>(smem_raw);
    __half* s_w = s_in + HeadSmem::IN_HALVES;
    float* s_out = reinterpret_cast<float*>(s_w + 2 * HeadSmem::W_HALVES);
    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    const int n_tiles = (m_used + kTileRows - 1) / kTileRows;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows;
        // ---- stage the grid features: 128 rows x 64 bytes
        for (int i = threadIdx.x; i < kTileRows * 4; i += kMlpThreads) {
            const int r = i >> 2, c = i & 3;
            const int row = row0 + r;
            __half* dst = s_in + r * HeadSmem::IN_STRIDE + c * 8;
            if (row < m_used) cp_async16(dst, feat + (size_t)row * 32 + c * 8);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
        mlp_forward_tile<kHeadW, kHeadK0, SAVE>(mw_s, s_in, s_w, s_out, fwd_s, M, row0);
        // ---- sigma epilogue + colour input stage (one thread per row)
        if (threadIdx.x < kTileRows) {
            const int r = threadIdx.x, row = row0 + r;
            const float* z = s_out + r * kOutStride;
            __half* ci = s_in + r * HeadSmem::IN_STRIDE;
            if (row < m_used) {
                __align__(16) __half hrow[16];
#pragma unroll
                for (int c = 0; c < 16; c++) hrow[c] = __float2half_rn(z[c]);
                sigma[row] = density_scale * expf(__half2float(hrow[0]));
                float sh[16];
                const float* d = dirs + (size_t)row * 3;
                sh_eval<4>(d[0], d[1], d[2], sh);
#pragma unroll
                for (int c = 0; c < 16; c++) ci[c] = __float2half_rn(sh[c]);
#pragma unroll
                for (int c = 1; c < 16; c++) ci[15 + c] = hrow[c];
                ci[31] = __float2half_rn(0.0f);
                if (SAVE) {
                    *reinterpret_cast<uint4*>(hs + (size_t)row * 16) = *reinterpret_cast<const uint4*>(hrow);
                    *reinterpret_cast<uint4*>(hs + (size_t)row * 16 + 8) = *reinterpret_cast<const uint4*>(hrow + 8);
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        *reinterpret_cast<uint4*>(cin + (size_t)row * 32 + 8 * c) = *reinterpret_cast<const uint4*>(ci + 8 * c);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(ci + 8 * c) = make_uint4(0, 0, 0, 0);
            }
        }
        mlp_forward_tile<kHeadW, kHeadK0, SAVE>(mw_c, s_in, s_w, s_out, fwd_c, M, row0);
        if (threadIdx.x < kTileRows) {
            const int row = row0 + threadIdx.x;
            if (row < m_used) {
                const float* z = s_out + threadIdx.x * kOutStride;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float zh = half_round(z[c]);
                    rgb[(size_t)row * 3 + c] = half_round(1.0f / (1.0f + expf(-zh)));
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// The same heads with the weights RESIDENT in shared memory and one WARP per 16-row slab, no block-wide barrier after the
// weights are in: the tile kernel above re-stages 20 KiB of weights for every 128-row tile (8 KiB of input) behind ~10
// __syncthreads and runs its per-row epilogues on half the CTA while the other half waits.  Here a warp loads its 16 feature rows,
// runs sigma net -> epilogue (exp, SH of the view direction, colour input) -> colour net -> sigmoid on its own, synchronising with
// __syncwarp only, so the 16 warps of an SM overlap each other's loads, MMAs and epilogues.  Same MMA order per output element as
// mlp_forward_tile: results are bit-identical to the tile kernel (tests/test_gpu_field.py).  SIGMA_ONLY = NeRFNetwork.density.
// ---------------------------------------------------------------------------------------------------
constexpr int kHrK0Stride = kHeadK0 + kPad;                       // 40 halves: rows of a K = 32 weight / input tile
constexpr int kHrWStride = kHeadW + kPad;                         // 72 halves: rows of a K = 64 weight
constexpr int kHrWarpBytes = 16 * kHrK0Stride * 2 + 16 * kOutStride * 4;  // per-warp input tile + output scratch

__host__ __device__ inline int head_weight_halves(const int n_layers) {  // layer 0 [64][40], hidden [64][72] x (n - 2), last [16][72]
    return kHeadW * kHrK0Stride + (n_layers - 2) * kHeadW * kHrWStride + 16 * kHrWStride;
}
__device__ __forceinline__ const __half* head_layer_ptr(const __half* base, const int l) {
    return l == 0 ? base : base + kHeadW * kHrK0Stride + (l - 1) * kHeadW * kHrWStride;
}

// One warp, 16 rows: input tile s_in [16][40] halves -> s_out [16][17] floats (last layer pre-activation); sw = this net's weights.
template <bool SAVE, bool TILED>
__device__ __forceinline__ void warp_head_mlp(const __half* sw, const int n_layers, const __half* s_in, float* s_out,
                                              __half* __restrict__ fwd_buf, const int M, const int m_used, const int row0) {
    constexpr int NT = kHeadW / 8, KT = kHeadW / 16;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    uint32_t areg[KT][4];
    float acc[NT][4];
    for (int l = 0; l < n_layers - 1; l++) {
        const __half* wcur = head_layer_ptr(sw, l);
#pragma unroll
        for (int n = 0; n < NT; n++) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f; }
        if (l == 0) {
#pragma unroll
            for (int kk = 0; kk < kHeadK0 / 16; kk++) {
                uint32_t a[4];
                ldmatrix_x4(a, s_in + ((lane & 7) + 8 * ((lane >> 3) & 1)) * kHrK0Stride + kk * 16 + 8 * (lane >> 4));
#pragma unroll
                for (int nn = 0; nn < NT; nn += 2) {
                    uint32_t b[4];
                    ldmatrix_x4(b, wcur + (8 * nn + (lane & 7) + 8 * (lane >> 4)) * kHrK0Stride + kk * 16 + 8 * ((lane >> 3) & 1));
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
                    ldmatrix_x4(b, wcur + (8 * nn + (lane & 7) + 8 * (lane >> 4)) * kHrWStride + kk * 16 + 8 * ((lane >> 3) & 1));
                    mma_16816(acc[nn], areg[kk], b[0], b[1]);
                    mma_16816(acc[nn + 1], areg[kk], b[2], b[3]);
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < KT; kk++) {
            areg[kk][0] = pack_half2(fmaxf(acc[2 * kk][0], 0.f), fmaxf(acc[2 * kk][1], 0.f));
            areg[kk][1] = pack_half2(fmaxf(acc[2 * kk][2], 0.f), fmaxf(acc[2 * kk][3], 0.f));
            areg[kk][2] = pack_half2(fmaxf(acc[2 * kk + 1][0], 0.f), fmaxf(acc[2 * kk + 1][1], 0.f));
            areg[kk][3] = pack_half2(fmaxf(acc[2 * kk + 1][2], 0.f), fmaxf(acc[2 * kk + 1][3], 0.f));
        }
        if (SAVE && TILED) {  // tile images, every row of a live tile (M = rows rounded up to 128; dead rows are zeros here)
            __half* base = fwd_buf + (size_t)l * M * kHeadW;
            const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
            for (int kk = 0; kk < KT; kk++) {
                *reinterpret_cast<uint32_t*>(base + tile_img_off<kHeadW>(r0, kk * 16 + 2 * t)) = areg[kk][0];
                *reinterpret_cast<uint32_t*>(base + tile_img_off<kHeadW>(r0, kk * 16 + 8 + 2 * t)) = areg[kk][2];
                *reinterpret_cast<uint32_t*>(base + tile_img_off<kHeadW>(r1, kk * 16 + 2 * t)) = areg[kk][1];
                *reinterpret_cast<uint32_t*>(base + tile_img_off<kHeadW>(r1, kk * 16 + 8 + 2 * t)) = areg[kk][3];
            }
        } else if (SAVE) {
            __half* base = fwd_buf + (size_t)l * M * kHeadW;
            const int r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
            for (int kk = 0; kk < KT; kk++) {
                if (r0 < m_used) {
                    *reinterpret_cast<uint32_t*>(base + (size_t)r0 * kHeadW + kk * 16 + 2 * t) = areg[kk][0];
                    *reinterpret_cast<uint32_t*>(base + (size_t)r0 * kHeadW + kk * 16 + 8 + 2 * t) = areg[kk][2];
                }
                if (r1 < m_used) {
                    *reinterpret_cast<uint32_t*>(base + (size_t)r1 * kHeadW + kk * 16 + 2 * t) = areg[kk][1];
                    *reinterpret_cast<uint32_t*>(base + (size_t)r1 * kHeadW + kk * 16 + 8 + 2 * t) = areg[kk][3];
                }
            }
        }
    }
    const __half* wlast = head_layer_ptr(sw, n_layers - 1);
    float o[2][4];
#pragma unroll
    for (int n = 0; n < 2; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f; }
#pragma unroll
    for (int kk = 0; kk < KT; kk++) {
        uint32_t b[4];
        ldmatrix_x4(b, wlast + ((lane & 7) + 8 * (lane >> 4)) * kHrWStride + kk * 16 + 8 * ((lane >> 3) & 1));
        mma_16816(o[0], areg[kk], b[0], b[1]);
        mma_16816(o[1], areg[kk], b[2], b[3]);
    }
#pragma unroll
    for (int n = 0; n < 2; n++) {
        s_out[g * kOutStride + 8 * n + 2 * t] = o[n][0];
        s_out[g * kOutStride + 8 * n + 2 * t + 1] = o[n][1];
        s_out[(g + 8) * kOutStride + 8 * n + 2 * t] = o[n][2];
        s_out[(g + 8) * kOutStride + 8 * n + 2 * t + 1] = o[n][3];
    }
}

// TILED (training, tcgen05 weight gradients): fwd_s / fwd_c / cin and a copy of the input features (feat_img) are saved as 128-row
// tile images with M_ld = M rounded up to 128 rows per layer; every row of a live tile is written (dead rows: zeros).
// GRID: the 32 input features are not read from HBM but ENCODED here — the hash-grid forward (grid_encoder.cu's k_grid_forward_pair,
// same gathers, same interpolation order: bit-identical features) for 16 levels x 2 features, fp16 table: lane = (row, half of the
// levels), 8 levels each in two passes of 4 with all gathers in flight, the 16 halves written straight into the warp's input tile.
// One launch and one HBM round trip of the features (64 B / sample each way) less per field evaluation.
struct GridArgs {
    const float* x01;     // [M,3] positions in [0,1]
    const __half* table;  // fp16 table (16-byte aligned)
    const int* offsets;   // [L+1]
    float S;
    uint32_t H, gridtype, interp;
    bool align_corners;
};

template <bool SAVE, bool SIGMA_ONLY, bool TILED, bool GRID>
__global__ void __launch_bounds__(kMlpThreads, 2) k_heads_forward_warp(const __half* __restrict__ feat, const GridArgs ga,
                                                                       const float* __restrict__ dirs,
                                                                       const MlpWeights mw_s, const MlpWeights mw_c, const int M,
                                                                       const int* __restrict__ m_dev, const float density_scale,
                                                                       float* __restrict__ sigma, float* __restrict__ rgb,
                                                                       __half* __restrict__ hs, __half* __restrict__ cin,
                                                                       __half* __restrict__ fwd_s, __half* __restrict__ fwd_c,
                                                                       __half* __restrict__ geo, __half* __restrict__ feat_img) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sw_s = reinterpret_cast<__half*>(smem_raw);
    __half* sw_c = sw_s + head_weight_halves(mw_s.n_layers);
    unsigned char* warp_base = reinterpret_cast<unsigned char*>(sw_c + (SIGMA_ONLY ? 0 : head_weight_halves(mw_c.n_layers)));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __half* s_in = reinterpret_cast<__half*>(warp_base + (size_t)warp * kHrWarpBytes);
    float* s_out = reinterpret_cast<float*>(s_in + 16 * kHrK0Stride);
    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    __shared__ LevelParams s_lp[16];
    if constexpr (GRID) {
        if (threadIdx.x < 16) s_lp[threadIdx.x] = make_level(ga.offsets, threadIdx.x, ga.S, ga.H, 3, ga.gridtype, ga.align_corners);
    }

    // ---- the weights, once per CTA
#pragma unroll 1
    for (int net = 0; net < (SIGMA_ONLY ? 1 : 2); net++) {
        const MlpWeights& mw = net ? mw_c : mw_s;
        __half* base = net ? sw_c : sw_s;
        for (int l = 0; l < mw.n_layers; l++) {
            __half* dst = const_cast<__half*>(head_layer_ptr(base, l));
            if (l == 0) stage_weights(dst, kHrK0Stride, mw.w[0], kHeadW, kHeadW, mw.k0, mw.k0_ld);
            else if (l < mw.n_layers - 1) stage_weights(dst, kHrWStride, mw.w[l], kHeadW, kHeadW, kHeadW, kHeadW);
            else stage_weights(dst, kHrWStride, mw.w[l], 16, mw.n_out, kHeadW, kHeadW);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int M_ld = TILED ? (M + 127) / 128 * 128 : M;                                  // rows per saved layer
    const int n_slabs = (SAVE && TILED) ? (m_used + 127) / 128 * 8 : (m_used + 15) / 16;   // tile images: whole live tiles
    for (int slab = blockIdx.x * (kMlpThreads / 32) + warp; slab < n_slabs; slab += gridDim.x * (kMlpThreads / 32)) {
        const int row0 = slab * 16;
        if constexpr (GRID) {
            // ---- encode: lane = (row r, levels [8 h, 8 h + 8))
            const int r = lane >> 1, h = lane & 1, row = row0 + r;
            float x[3] = {0.f, 0.f, 0.f};
            bool live = row < m_used;
            if (live) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    x[d] = ga.x01[(size_t)row * 3 + d];
                    if (x[d] < 0 || x[d] > 1) live = false;  // outside [0,1]: zero features (gridencoder.cu:117-131)
                }
            }
            constexpr int GL = 2;  // levels in flight per pass (4 would need > 128 registers next to the MLP fragments: spills)
            uint32_t w[8];
#pragma unroll
            for (int pass = 0; pass < 8 / GL; pass++) {
                CellGather<__half, 3, 2> cg[GL];
                float pos[GL][3], deriv[GL][3];
                if (live) {
#pragma unroll
                    for (int g = 0; g < GL; g++) {
                        const LevelParams lp = s_lp[h * 8 + pass * GL + g];
                        uint32_t pos_grid[3];
                        locate<3>(x, lp, ga.align_corners, ga.interp, pos[g], deriv[g], pos_grid);
                        cg[g].issue(ga.table, ga.gridtype, ga.align_corners, lp, pos_grid);
                    }
                }
#pragma unroll
                for (int g = 0; g < GL; g++) {
                    float res[2] = {0.f, 0.f};
                    if (live) {
                        float val[8][2];
                        cg[g].resolve(val);
#pragma unroll
                        for (uint32_t idx = 0; idx < 8; idx++) {
                            float wt = 1;
#pragma unroll
                            for (uint32_t d = 0; d < 3; d++) wt *= ((idx >> d) & 1u) ? pos[g][d] : 1 - pos[g][d];
                            res[0] += wt * val[idx][0];
                            res[1] += wt * val[idx][1];
                        }
                    }
                    __half hh[2];
                    Row<__half, 2>::store(hh, res);
                    w[pass * GL + g] = *reinterpret_cast<const uint32_t*>(hh);
                }
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint4 v = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                *reinterpret_cast<uint4*>(s_in + r * kHrK0Stride + h * 16 + q * 8) = v;
                if (SAVE && TILED) *reinterpret_cast<uint4*>(feat_img + tile_img_off<kHeadK0>(row, h * 16 + q * 8)) = v;
            }
        } else {
            // ---- 16 feature rows x 64 bytes: two 16-byte chunks per lane
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int idx = lane * 2 + c, r = idx >> 2, ch = idx & 3;
                const int row = row0 + r;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (row < m_used) v = __ldg(reinterpret_cast<const uint4*>(feat + (size_t)row * 32 + ch * 8));
                *reinterpret_cast<uint4*>(s_in + r * kHrK0Stride + ch * 8) = v;
                if (SAVE && TILED) *reinterpret_cast<uint4*>(feat_img + tile_img_off<kHeadK0>(row, ch * 8)) = v;
            }
        }
        __syncwarp();
        warp_head_mlp<SAVE, TILED>(sw_s, mw_s.n_layers, s_in, s_out, fwd_s, M_ld, m_used, row0);
        __syncwarp();
        // ---- sigma epilogue + colour input (lanes 0-15: one row each)
        if (lane < 16) {
            const int r = lane, row = row0 + r;
            const float* z = s_out + r * kOutStride;
            __half* ci = s_in + r * kHrK0Stride;
            if (row < m_used) {
                __align__(16) __half hrow[16];
#pragma unroll
                for (int c = 0; c < 16; c++) hrow[c] = __float2half_rn(z[c]);
                sigma[row] = density_scale * expf(__half2float(hrow[0]));
                if constexpr (SIGMA_ONLY) {
                    if (geo) {
#pragma unroll
                        for (int c = 1; c < 16; c++) geo[(size_t)row * 15 + c - 1] = hrow[c];
                    }
                } else {
                    float sh[16];
                    const float* d = dirs + (size_t)row * 3;
                    sh_eval<4>(d[0], d[1], d[2], sh);
#pragma unroll
                    for (int c = 0; c < 16; c++) ci[c] = __float2half_rn(sh[c]);
#pragma unroll
                    for (int c = 1; c < 16; c++) ci[15 + c] = hrow[c];
                    ci[31] = __float2half_rn(0.0f);
                    if (SAVE) {
                        *reinterpret_cast<uint4*>(hs + (size_t)row * 16) = *reinterpret_cast<const uint4*>(hrow);
                        *reinterpret_cast<uint4*>(hs + (size_t)row * 16 + 8) = *reinterpret_cast<const uint4*>(hrow + 8);
                        if constexpr (!TILED) {
#pragma unroll
                            for (int c = 0; c < 4; c++)
                                *reinterpret_cast<uint4*>(cin + (size_t)row * 32 + 8 * c) = *reinterpret_cast<const uint4*>(ci + 8 * c);
                        }
                    }
                }
            } else if constexpr (!SIGMA_ONLY) {
#pragma unroll
                for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(ci + 8 * c) = make_uint4(0, 0, 0, 0);
            }
            if constexpr (SAVE && TILED && !SIGMA_ONLY) {  // colour input as a tile image, dead rows of a live tile included (zeros)
#pragma unroll
                for (int c = 0; c < 4; c++)
                    *reinterpret_cast<uint4*>(cin + tile_img_off<kHeadK0>(row, 8 * c)) = *reinterpret_cast<const uint4*>(ci + 8 * c);
            }
        }
        __syncwarp();
        if constexpr (!SIGMA_ONLY) {
            warp_head_mlp<SAVE, TILED>(sw_c, mw_c.n_layers, s_in, s_out, fwd_c, M_ld, m_used, row0);
            __syncwarp();
            if (lane < 16) {
                const int row = row0 + lane;
                if (row < m_used) {
                    const float* z = s_out + lane * kOutStride;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float zh = half_round(z[c]);
                        rgb[(size_t)row * 3 + c] = half_round(1.0f / (1.0f + expf(-zh)));
                    }
                }
            }
            __syncwarp();
        }
    }
}

// Density-only variant (NeRFNetwork.density, dnerf/network.py:193-206): sigma MLP only.
__global__ void __launch_bounds__(kMlpThreads, 2) k_sigma_forward(const __half* __restrict__ feat, const MlpWeights mw_s, const int M,
                                                                  const float density_scale, float* __restrict__ sigma,
                                                                  __half* __restrict__ geo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_in = reinterpret_cast<__half*>(smem_raw);
    __half* s_w = s_in + HeadSmem::IN_HALVES;
    float* s_out = reinterpret_cast<float*>(s_w + 2 * HeadSmem::W_HALVES);
    const int n_tiles = (M + kTileRows - 1) / kTileRows;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows;
        for (int i = threadIdx.x; i < kTileRows * 4; i += kMlpThreads) {
            const int r = i >> 2, c = i & 3;
            const int row = row0 + r;
            __half* dst = s_in + r * HeadSmem::IN_STRIDE + c * 8;
            if (row < M) cp_async16(dst, feat + (size_t)row * 32 + c * 8);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
        mlp_forward_tile<kHeadW, kHeadK0, false>(mw_s, s_in, s_w, s_out, nullptr, M, row0);
        if (threadIdx.x < kTileRows) {
            const int row = row0 + threadIdx.x;
            if (row < M) {
                const float* z = s_out + threadIdx.x * kOutStride;
                sigma[row] = density_scale * expf(half_round(z[0]));
                if (geo) {
#pragma unroll
                    for (int c = 1; c < 16; c++) geo[(size_t)row * 15 + c - 1] = __float2half_rn(z[c]);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// sigma + colour heads backward
//   grad_sigma [M] fp32, grad_rgb [M,3] fp32 (already multiplied by the loss scale)
//   outputs: dfeat [M,32] fp16, and for the weight-gradient GEMMs: gout_c [M,16], gout_s [M,16], bwd_c [2][M][64], bwd_s [1][M][64]
// ---------------------------------------------------------------------------------------------------
template <bool TILED>
__global__ void __launch_bounds__(kMlpThreads, 2) k_heads_backward(const float* __restrict__ grad_sigma, const float* __restrict__ grad_rgb,
                                                                   const float* __restrict__ rgb, const __half* __restrict__ hs,
                                                                   const MlpWeights mw_s, const MlpWeights mw_c, const int M,
                                                                   const int* __restrict__ m_dev, const float density_scale,
                                                                   const __half* __restrict__ fwd_s, const __half* __restrict__ fwd_c,
                                                                   __half* __restrict__ bwd_s, __half* __restrict__ bwd_c,
                                                                   __half* __restrict__ gout_s, __half* __restrict__ gout_c,
                                                                   __half* __restrict__ dfeat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_in = reinterpret_cast<__half*>(smem_raw);  // colour dinput tile [128][40] / g_out tile [128][24]
    __half* s_w = s_in + HeadSmem::IN_HALVES;
    __half* s_g = reinterpret_cast<__half*>(s_w + 2 * HeadSmem::W_HALVES);  // [128][24] halves (fits in the s_out region)
    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    const int n_tiles = (m_used + kTileRows - 1) / kTileRows;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows;
        // ---- colour g_out: d sigmoid
        if (threadIdx.x < kTileRows) {
            const int r = threadIdx.x, row = row0 + r;
            __align__(16) __half g16[16];
#pragma unroll
            for (int c = 0; c < 16; c++) g16[c] = __float2half_rn(0.0f);
            if (row < m_used) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float y = rgb[(size_t)row * 3 + c];
                    g16[c] = __float2half_rn(grad_rgb[(size_t)row * 3 + c] * y * (1.0f - y));
                }
                if constexpr (!TILED) {
                    *reinterpret_cast<uint4*>(gout_c + (size_t)row * 16) = *reinterpret_cast<const uint4*>(g16);
                    *reinterpret_cast<uint4*>(gout_c + (size_t)row * 16 + 8) = *reinterpret_cast<const uint4*>(g16 + 8);
                }
            }
            if constexpr (TILED) {  // tile image [tile][2][128][8], dead rows of the live tile as zeros
                *reinterpret_cast<uint4*>(gout_c + tile_img_off<16>(row, 0)) = *reinterpret_cast<const uint4*>(g16);
                *reinterpret_cast<uint4*>(gout_c + tile_img_off<16>(row, 8)) = *reinterpret_cast<const uint4*>(g16 + 8);
            }
            *reinterpret_cast<uint4*>(s_g + r * kGStride) = *reinterpret_cast<const uint4*>(g16);
            *reinterpret_cast<uint4*>(s_g + r * kGStride + 8) = *reinterpret_cast<const uint4*>(g16 + 8);
        }
        // colour backward; dL/d(colour input) lands in the s_in tile
        mlp_backward_tile<kHeadW, kHeadK0, TILED>(mw_c, s_g, s_w, fwd_c, bwd_c, s_in, HeadSmem::IN_STRIDE, 0, TILED ? (M + 127) / 128 * 128 : M, m_used, row0);
        // ---- sigma g_out: trunc_exp backward on column 0, d geo on columns 1..15
        if (threadIdx.x < kTileRows) {
            const int r = threadIdx.x, row = row0 + r;
            __align__(16) __half g16[16];
#pragma unroll
            for (int c = 0; c < 16; c++) g16[c] = __float2half_rn(0.0f);
            if (row < m_used) {
                const float h0 = __half2float(hs[(size_t)row * 16]);
                g16[0] = __float2half_rn(grad_sigma[row] * density_scale * expf(fminf(fmaxf(h0, -15.0f), 15.0f)));
#pragma unroll
                for (int c = 1; c < 16; c++) g16[c] = s_in[r * HeadSmem::IN_STRIDE + 15 + c];
                if constexpr (!TILED) {
                    *reinterpret_cast<uint4*>(gout_s + (size_t)row * 16) = *reinterpret_cast<const uint4*>(g16);
                    *reinterpret_cast<uint4*>(gout_s + (size_t)row * 16 + 8) = *reinterpret_cast<const uint4*>(g16 + 8);
                }
            }
            if constexpr (TILED) {
                *reinterpret_cast<uint4*>(gout_s + tile_img_off<16>(row, 0)) = *reinterpret_cast<const uint4*>(g16);
                *reinterpret_cast<uint4*>(gout_s + tile_img_off<16>(row, 8)) = *reinterpret_cast<const uint4*>(g16 + 8);
            }
            *reinterpret_cast<uint4*>(s_g + r * kGStride) = *reinterpret_cast<const uint4*>(g16);
            *reinterpret_cast<uint4*>(s_g + r * kGStride + 8) = *reinterpret_cast<const uint4*>(g16 + 8);
        }
        mlp_backward_tile<kHeadW, kHeadK0, TILED>(mw_s, s_g, s_w, fwd_s, bwd_s, dfeat, 32, row0, TILED ? (M + 127) / 128 * 128 : M, m_used, row0);
    }
}

// ---------------------------------------------------------------------------------------------------
// Weight gradients: dW[N][K] (+)= sum_m G[m][N]^T * A[m][K]   (reference: CUTLASS split-K GEMMs, ffmlp.cu:801-877)
// One CTA reduces a chunk of rows with mma.sync (fp32 accumulate) and adds its partial with fp32 atomics.
//   G [M][ldg] fp16 (N columns used), A [M][lda] fp16 (K columns used); dW [n_real][ldw] fp32, rows >= n_real dropped,
//   columns >= k_real dropped.  blockIdx.y selects the job.
// ---------------------------------------------------------------------------------------------------
struct WgradJob {
    const __half* G;
    const __half* A;
    float* dW;
    int N, K;        // padded GEMM dims (multiples of 16)
    int ldg, lda, ldw;
    int n_real, k_real;
};
constexpr int kMaxWgradJobs = 16;
struct WgradJobs {
    WgradJob j[kMaxWgradJobs];
    int first_cta[kMaxWgradJobs + 1];  // job i owns CTAs [first_cta[i], first_cta[i+1]): a share proportional to its bytes per row
    int n_jobs;
};

constexpr int kWgChunk = 32;   // rows staged per step
constexpr int kWgStages = 4;   // cp.async pipeline depth: three chunks in flight while one is consumed (the loop is load-latency bound)

// Warp -> output tiles: the N/16 m-tiles are spread over the 8 warps (N in {16,32,64,128}); the warps sharing an m-tile
// split the K/8 n-tiles in even-sized contiguous ranges (pairs of n-tiles share one ldmatrix.x4).
__global__ void __launch_bounds__(256) k_wgrad(const __grid_constant__ WgradJobs jobs, const int M, const int* __restrict__ m_dev) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // one wave of CTAs; every job's rows are split evenly over its CTAs, so all CTAs stream about the same number of bytes
    int job = 0;
    while (job + 1 < jobs.n_jobs && (int)blockIdx.x >= jobs.first_cta[job + 1]) job++;
    const WgradJob& jb = jobs.j[job];
    const int part = (int)blockIdx.x - jobs.first_cta[job], parts = jobs.first_cta[job + 1] - jobs.first_cta[job];
    const int N = jb.N, K = jb.K;
    const int gs = N + kPad, as = K + kPad;  // smem strides
    __half* const s_base = reinterpret_cast<__half*>(smem_raw);
    const int stage_halves = kWgChunk * (gs + as);
    auto s_gt = [&](const int i) { return s_base + i * stage_halves; };
    auto s_at = [&](const int i) { return s_base + i * stage_halves + kWgChunk * gs; };

    const int m_used = m_dev ? min(M, max(*m_dev, 0)) : M;
    const int rows_per_cta = ((m_used + parts - 1) / parts + kWgChunk - 1) / kWgChunk * kWgChunk;
    const int m_begin = part * rows_per_cta;
    const int m_end = min(m_used, m_begin + rows_per_cta);
    if (m_begin >= m_end) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int MT = N / 16, NT = K / 8;
    const int wpm = 8 / MT;                        // warps per m-tile
    const int mt = warp / wpm;
    const int cnt = (((NT + wpm - 1) / wpm) + 1) & ~1;  // n-tiles per warp (even)
    const int nt0 = (warp % wpm) * cnt;
    const int n_pairs = max(0, min(cnt, NT - nt0)) / 2;  // <= 8

    float acc[8][2][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int h = 0; h < 2; h++) acc[i][h][0] = acc[i][h][1] = acc[i][h][2] = acc[i][h][3] = 0.0f;

    // Staging: every thread owns up to 4 fixed (row-in-chunk, 16-byte column) slots of a chunk; source pointers and
    // shared-memory offsets are computed once, a stage is then 4 predicated cp.async per thread (the integer divisions of a
    // per-copy index decode made the loop issue bound).
    const int gchunks = N / 8, achunks = K / 8, chunks = gchunks + achunks;
    constexpr int kSlots = 4;  // kWgChunk * chunks / 256 <= 32 * 32 / 256
    const __half* src[kSlots];
    int dst[kSlots], srow[kSlots], sld[kSlots];
#pragma unroll
    for (int k = 0; k < kSlots; k++) {
        const int i = threadIdx.x + 256 * k;
        const int r = i / chunks, c = i - r * chunks;
        const bool is_g = c < gchunks;
        const int cc = is_g ? c : c - gchunks;
        srow[k] = (i < kWgChunk * chunks) ? r : (1 << 28);  // inactive slots never pass the row test
        sld[k] = kWgChunk * (is_g ? jb.ldg : jb.lda);
        src[k] = (is_g ? jb.G + (size_t)(m_begin + r) * jb.ldg : jb.A + (size_t)(m_begin + r) * jb.lda) + cc * 8;
        dst[k] = is_g ? (r * gs + cc * 8) : (kWgChunk * gs + r * as + cc * 8);
    }
    int m_next = m_begin;  // first row of the next chunk to stage (stages are issued in order)
    auto stage = [&](int buf, int) {
        __half* sb = s_base + buf * stage_halves;
#pragma unroll
        for (int k = 0; k < kSlots; k++) {
            if (m_next + srow[k] < m_end) cp_async16(sb + dst[k], src[k]);
            else if (srow[k] < kWgChunk) *reinterpret_cast<uint4*>(sb + dst[k]) = make_uint4(0, 0, 0, 0);
            src[k] += sld[k];
        }
        m_next += kWgChunk;
        cp_async_commit();
    };

    // per-lane smem offsets of the two ldmatrix patterns (halves)
    //  A = G^T (m = n_idx, k = sample): matrices (s 0-7, n 0-7), (s 0-7, n 8-15), (s 8-15, n 0-7), (s 8-15, n 8-15), transposed
    const int a_off = ((lane & 7) + 8 * (lane >> 4)) * gs + 16 * mt + 8 * ((lane >> 3) & 1);
    //  B (k = sample, n = k_idx), two n-tiles: (s 0-7, c 0-7), (s 8-15, c 0-7), (s 0-7, c 8-15), (s 8-15, c 8-15), transposed
    const int b_off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * as + 8 * nt0 + 8 * (lane >> 4);

    const int n_steps = (m_end - m_begin + kWgChunk - 1) / kWgChunk;
#pragma unroll
    for (int s = 0; s < kWgStages - 1; s++) {
        if (s < n_steps) stage(s, m_begin + s * kWgChunk);
        else cp_async_commit();
    }
    for (int s = 0; s < n_steps; s++) {
        const int buf = s % kWgStages;
        cp_async_wait<kWgStages - 2>();  // chunk s has landed (one group is committed per iteration)
        __syncthreads();                 // ... for every thread, and everyone is done with the buffer refilled below
        if (s + kWgStages - 1 < n_steps) stage((s + kWgStages - 1) % kWgStages, m_begin + (s + kWgStages - 1) * kWgChunk);
        else cp_async_commit();
        if (n_pairs > 0) {
            const __half* gt = s_gt(buf) + a_off;
            const __half* at = s_at(buf) + b_off;
#pragma unroll
            for (int kk = 0; kk < kWgChunk / 16; kk++) {
                uint32_t a[4];
                ldmatrix_x4_trans(a, gt + 16 * kk * gs);
#pragma unroll
                for (int pr = 0; pr < 8; pr++) {
                    if (pr < n_pairs) {
                        uint32_t b[4];
                        ldmatrix_x4_trans(b, at + 16 * kk * as + 16 * pr);
                        mma_16816(acc[pr][0], a, b[0], b[1]);
                        mma_16816(acc[pr][1], a, b[2], b[3]);
                    }
                }
            }
        }
    }
    // flush partial sums with fp32 (vector) atomics
    const bool vec_ok = (jb.ldw % 2 == 0) && (((uintptr_t)jb.dW % 8) == 0);
#pragma unroll
    for (int pr = 0; pr < 8; pr++) {
        if (pr < n_pairs) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int c0 = 8 * (nt0 + 2 * pr + h) + 2 * t;
#pragma unroll
                for (int half_ = 0; half_ < 2; half_++) {
                    const int r = 16 * mt + g + 8 * half_;
                    if (r >= jb.n_real || c0 >= jb.k_real) continue;
                    const float v0 = acc[pr][h][2 * half_], v1 = acc[pr][h][2 * half_ + 1];
                    float* dst = jb.dW + (size_t)r * jb.ldw + c0;
                    if (vec_ok && c0 + 1 < jb.k_real) {
                        atomicAdd(reinterpret_cast<float2*>(dst), make_float2(v0, v1));
                    } else {
                        atomicAdd(dst, v0);
                        if (c0 + 1 < jb.k_real) atomicAdd(dst + 1, v1);
                    }
                }
            }
        }
    }
}

}  // namespace seald

// ===================================================================================================
// C-ABI
// ===================================================================================================
using namespace seald;

namespace {

int make_weights(MlpWeights& mw, const void* const* w, int n_layers, int k0, int k0_ld, int n_out) {
    if (!w || n_layers < 2 || n_layers > kMaxLayers) return SEALD_E_BADARG;
    for (int i = 0; i < n_layers; i++) {
        if (!w[i] || ((uintptr_t)w[i] % 16) != 0) return w[i] ? SEALD_E_ALIGN : SEALD_E_BADARG;
        mw.w[i] = reinterpret_cast<const __half*>(w[i]);
    }
    mw.n_layers = n_layers;
    mw.k0 = k0;
    mw.k0_ld = k0_ld;
    mw.n_out = n_out;
    return 0;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e == cudaSuccess ? 0 : (int)e;
}

inline int tiles_grid(uint32_t M, int ctas_per_sm) {
    const int n_tiles = (int)((M + kTileRows - 1) / kTileRows);
    const int cap = SEALD_NUM_SMS * ctas_per_sm;
    return n_tiles < cap ? (n_tiles > 0 ? n_tiles : 1) : cap;
}

}  // namespace

extern "C" int seald_field_deform_forward(const float* xyz, const float* time_dev, const void* const* weights, int n_layers, uint32_t M,
                                          const int32_t* m_dev, float bound, int t0_mode, float* deform, float* x01, void* in_buf,
                                          void* fwd_buf, seald_stream_t stream) {
    if (M == 0) return 0;
    if (!xyz || !time_dev || !deform || !x01) return SEALD_E_BADARG;
    if ((in_buf == nullptr) != (fwd_buf == nullptr)) return SEALD_E_BADARG;
    MlpWeights mw;
    int rc = make_weights(mw, weights, n_layers, kDeformK0, kDeformK0, 3);
    if (rc) return rc;
    const size_t smem = DeformSmem::BYTES;
    cudaStream_t st = to_stream(stream);
    if (fwd_buf) {
        if ((rc = set_smem(k_deform_forward<true>, smem))) return rc;
        k_deform_forward<true><<<tiles_grid(M, 2), kMlpThreads, smem, st>>>(xyz, time_dev, mw, (int)M, m_dev, bound, t0_mode, deform, x01,
                                                                           (__half*)in_buf, (__half*)fwd_buf);
    } else {
        if ((rc = set_smem(k_deform_forward<false>, smem))) return rc;
        k_deform_forward<false><<<tiles_grid(M, 2), kMlpThreads, smem, st>>>(xyz, time_dev, mw, (int)M, m_dev, bound, t0_mode, deform, x01,
                                                                            nullptr, nullptr);
    }
    return launch_status();
}

extern "C" int seald_field_deform_backward(const float* grad_x01, const float* time_dev, const void* const* weights, int n_layers, uint32_t M, const int32_t* m_dev,
                                           float bound, const void* fwd_buf, void* bwd_buf, void* gout_buf, seald_stream_t stream) {
    if (M == 0) return 0;
    if (!grad_x01 || !fwd_buf || !bwd_buf || !gout_buf) return SEALD_E_BADARG;
    MlpWeights mw;
    int rc = make_weights(mw, weights, n_layers, kDeformK0, kDeformK0, 3);
    if (rc) return rc;
    const size_t smem = DeformSmem::BYTES;
    if ((rc = set_smem(k_deform_backward, smem))) return rc;
    k_deform_backward<<<tiles_grid(M, 2), kMlpThreads, smem, to_stream(stream)>>>(grad_x01, time_dev, mw, (int)M, m_dev, bound, (const __half*)fwd_buf,
                                                                                 (__half*)bwd_buf, (__half*)gout_buf);
    return launch_status();
}

static int heads_forward_impl(const void* feat, const float* dirs, const void* const* w_sigma, int n_sigma, const void* const* w_color,
                              int n_color, uint32_t M, const int32_t* m_dev, float density_scale, float* sigma, float* rgb, void* hs,
                              void* cin, void* fwd_s, void* fwd_c, void* feat_img, seald_stream_t stream, const GridArgs* grid = nullptr) {
    if (M == 0) return 0;
    if ((!feat && !grid) || !dirs || !sigma || !rgb) return SEALD_E_BADARG;
    const GridArgs ga = grid ? *grid : GridArgs{};
    const bool save = hs != nullptr;
    const bool tiled = feat_img != nullptr;
    if (save && (!cin || !fwd_s || !fwd_c)) return SEALD_E_BADARG;
    if (tiled && !save) return SEALD_E_BADARG;
    MlpWeights ms, mc;
    int rc = make_weights(ms, w_sigma, n_sigma, kHeadK0, kHeadK0, 16);
    if (rc) return rc;
    if ((rc = make_weights(mc, w_color, n_color, kHeadK0, kHeadK0, 3))) return rc;
    cudaStream_t st = to_stream(stream);
    // weights resident + one warp per 16-row slab (default); SEALD_HEADS_IMPL=tile: the 128-row tile kernel (measurement switch, and
    // the fallback when the nets are too deep for their weights to stay in shared memory)
    static const bool tile_impl = getenv("SEALD_HEADS_IMPL") && std::string(getenv("SEALD_HEADS_IMPL")) == "tile";
    const size_t smem_res = (size_t)(head_weight_halves(n_sigma) + head_weight_halves(n_color)) * 2 + (kMlpThreads / 32) * kHrWarpBytes;
    if ((tiled || grid) && (tile_impl || smem_res > 96 * 1024)) return SEALD_E_UNSUPPORTED;  // tile images / fused encoder: warp kernel only
    if (grid && save && !tiled) return SEALD_E_UNSUPPORTED;
    if (!tile_impl && smem_res <= 96 * 1024) {
        const uint32_t slabs = tiled ? div_up(M, 128u) * 8u : div_up(M, 16u);
        uint32_t blocks = div_up(slabs, (uint32_t)(kMlpThreads / 32));
        if (blocks > 2u * SEALD_NUM_SMS) blocks = 2u * SEALD_NUM_SMS;
#define SEALD_LAUNCH_HEADS(SAVE_, TILED_, GRID_, HS, CIN, FS, FC, FIMG)                                                                  \
    do {                                                                                                                             \
        auto k = k_heads_forward_warp<SAVE_, false, TILED_, GRID_>;                                                                  \
        if ((rc = set_smem(k, smem_res))) return rc;                                                                                 \
        k<<<blocks, kMlpThreads, smem_res, st>>>((const __half*)feat, ga, dirs, ms, mc, (int)M, m_dev, density_scale, sigma, rgb,    \
                                                 (__half*)(HS), (__half*)(CIN), (__half*)(FS), (__half*)(FC), nullptr, (__half*)(FIMG)); \
    } while (0)
        if (save && tiled) {
            if (grid) SEALD_LAUNCH_HEADS(true, true, true, hs, cin, fwd_s, fwd_c, feat_img);
            else SEALD_LAUNCH_HEADS(true, true, false, hs, cin, fwd_s, fwd_c, feat_img);
        } else if (save) {
            SEALD_LAUNCH_HEADS(true, false, false, hs, cin, fwd_s, fwd_c, nullptr);
        } else {
            if (grid) SEALD_LAUNCH_HEADS(false, false, true, nullptr, nullptr, nullptr, nullptr, nullptr);
            else SEALD_LAUNCH_HEADS(false, false, false, nullptr, nullptr, nullptr, nullptr, nullptr);
        }
#undef SEALD_LAUNCH_HEADS
        return launch_status();
    }
    const size_t smem = HeadSmem::BYTES;
    if (save) {
        if ((rc = set_smem(k_heads_forward<true>, smem))) return rc;
        k_heads_forward<true><<<tiles_grid(M, 2), kMlpThreads, smem, st>>>((const __half*)feat, dirs, ms, mc, (int)M, m_dev, density_scale, sigma, rgb,
                                                                          (__half*)hs, (__half*)cin, (__half*)fwd_s, (__half*)fwd_c);
    } else {
        if ((rc = set_smem(k_heads_forward<false>, smem))) return rc;
        k_heads_forward<false><<<tiles_grid(M, 2), kMlpThreads, smem, st>>>((const __half*)feat, dirs, ms, mc, (int)M, m_dev, density_scale, sigma, rgb,
                                                                           nullptr, nullptr, nullptr, nullptr);
    }
    return launch_status();
}

extern "C" int seald_field_heads_forward(const void* feat, const float* dirs, const void* const* w_sigma, int n_sigma, const void* const* w_color,
                                         int n_color, uint32_t M, const int32_t* m_dev, float density_scale, float* sigma, float* rgb, void* hs,
                                         void* cin, void* fwd_s, void* fwd_c, seald_stream_t stream) {
    return heads_forward_impl(feat, dirs, w_sigma, n_sigma, w_color, n_color, M, m_dev, density_scale, sigma, rgb, hs, cin, fwd_s, fwd_c, nullptr,
                              stream);
}

extern "C" int seald_field_heads_forward_tiled(const void* feat, const float* dirs, const void* const* w_sigma, int n_sigma,
                                               const void* const* w_color, int n_color, uint32_t M, const int32_t* m_dev, float density_scale,
                                               float* sigma, float* rgb, void* hs, void* cin, void* fwd_s, void* fwd_c, void* feat_img,
                                               seald_stream_t stream) {
    if (!feat_img) return SEALD_E_BADARG;
    return heads_forward_impl(feat, dirs, w_sigma, n_sigma, w_color, n_color, M, m_dev, density_scale, sigma, rgb, hs, cin, fwd_s, fwd_c, feat_img,
                              stream);
}

static int sigma_forward_impl(const void* feat, const void* const* w_sigma, int n_sigma, uint32_t M, float density_scale, float* sigma,
                              void* geo, seald_stream_t stream, const GridArgs* grid) {
    if (M == 0) return 0;
    if ((!feat && !grid) || !sigma) return SEALD_E_BADARG;
    const GridArgs ga = grid ? *grid : GridArgs{};
    MlpWeights ms;
    int rc = make_weights(ms, w_sigma, n_sigma, kHeadK0, kHeadK0, 16);
    if (rc) return rc;
    static const bool tile_impl = getenv("SEALD_HEADS_IMPL") && std::string(getenv("SEALD_HEADS_IMPL")) == "tile";
    const size_t smem_res = (size_t)head_weight_halves(n_sigma) * 2 + (kMlpThreads / 32) * kHrWarpBytes;
    if (!tile_impl && smem_res <= 96 * 1024) {
        const uint32_t slabs = div_up(M, 16u);
        uint32_t blocks = div_up(slabs, (uint32_t)(kMlpThreads / 32));
        if (blocks > 2u * SEALD_NUM_SMS) blocks = 2u * SEALD_NUM_SMS;
        auto k = grid ? k_heads_forward_warp<false, true, false, true> : k_heads_forward_warp<false, true, false, false>;
        if ((rc = set_smem(k, smem_res))) return rc;
        k<<<blocks, kMlpThreads, smem_res, to_stream(stream)>>>((const __half*)feat, ga, nullptr, ms, ms, (int)M, nullptr, density_scale, sigma,
                                                                nullptr, nullptr, nullptr, nullptr, nullptr, (__half*)geo, nullptr);
        return launch_status();
    }
    if (grid) return SEALD_E_UNSUPPORTED;
    const size_t smem = HeadSmem::BYTES;
    if ((rc = set_smem(k_sigma_forward, smem))) return rc;
    k_sigma_forward<<<tiles_grid(M, 2), kMlpThreads, smem, to_stream(stream)>>>((const __half*)feat, ms, (int)M, density_scale, sigma, (__half*)geo);
    return launch_status();
}

extern "C" int seald_field_sigma_forward(const void* feat, const void* const* w_sigma, int n_sigma, uint32_t M, float density_scale, float* sigma,
                                         void* geo, seald_stream_t stream) {
    return sigma_forward_impl(feat, w_sigma, n_sigma, M, density_scale, sigma, geo, stream, nullptr);
}

// hash-grid encoder (16 levels x 2 features, 3-D, fp16 table) fused into the heads / the density head
static int make_grid_args(GridArgs& ga, const float* x01, const void* table, const int32_t* offsets, uint32_t D, uint32_t C, uint32_t L, float S,
                          uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp) {
    if (!x01 || !table || !offsets) return SEALD_E_BADARG;
    if (D != 3 || C != 2 || L != 16 || gridtype > 1 || interp > 1) return SEALD_E_UNSUPPORTED;
    if ((uintptr_t)table % 16) return SEALD_E_ALIGN;
    ga = GridArgs{x01, (const __half*)table, offsets, S, H, gridtype, interp, align_corners != 0};
    return 0;
}

extern "C" int seald_field_grid_heads_forward(const float* x01, const void* table, const int32_t* offsets, uint32_t D, uint32_t C, uint32_t L,
                                              float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, const float* dirs,
                                              const void* const* w_sigma, int n_sigma, const void* const* w_color, int n_color, uint32_t M,
                                              const int32_t* m_dev, float density_scale, float* sigma, float* rgb, void* hs, void* cin,
                                              void* fwd_s, void* fwd_c, void* feat_img, seald_stream_t stream) {
    GridArgs ga;
    int rc = make_grid_args(ga, x01, table, offsets, D, C, L, S, H, gridtype, align_corners, interp);
    if (rc) return rc;
    return heads_forward_impl(nullptr, dirs, w_sigma, n_sigma, w_color, n_color, M, m_dev, density_scale, sigma, rgb, hs, cin, fwd_s, fwd_c, feat_img,
                              stream, &ga);
}

extern "C" int seald_field_grid_sigma_forward(const float* x01, const void* table, const int32_t* offsets, uint32_t D, uint32_t C, uint32_t L,
                                              float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                                              const void* const* w_sigma, int n_sigma, uint32_t M, float density_scale, float* sigma, void* geo,
                                              seald_stream_t stream) {
    GridArgs ga;
    int rc = make_grid_args(ga, x01, table, offsets, D, C, L, S, H, gridtype, align_corners, interp);
    if (rc) return rc;
    return sigma_forward_impl(nullptr, w_sigma, n_sigma, M, density_scale, sigma, geo, stream, &ga);
}

static int heads_backward_impl(const float* grad_sigma, const float* grad_rgb, const float* rgb, const void* hs,
                               const void* const* w_sigma, int n_sigma, const void* const* w_color, int n_color, uint32_t M,
                               const int32_t* m_dev, float density_scale, const void* fwd_s, const void* fwd_c, void* bwd_s, void* bwd_c,
                               void* gout_s, void* gout_c, void* dfeat, bool tiled, seald_stream_t stream) {
    if (M == 0) return 0;
    if (!grad_sigma || !grad_rgb || !rgb || !hs || !fwd_s || !fwd_c || !bwd_s || !bwd_c || !gout_s || !gout_c || !dfeat) return SEALD_E_BADARG;
    MlpWeights ms, mc;
    int rc = make_weights(ms, w_sigma, n_sigma, kHeadK0, kHeadK0, 16);
    if (rc) return rc;
    if ((rc = make_weights(mc, w_color, n_color, kHeadK0, kHeadK0, 3))) return rc;
    const size_t smem = HeadSmem::BYTES;
    auto k = tiled ? k_heads_backward<true> : k_heads_backward<false>;
    if ((rc = set_smem(k, smem))) return rc;
    k<<<tiles_grid(M, 2), kMlpThreads, smem, to_stream(stream)>>>(grad_sigma, grad_rgb, rgb, (const __half*)hs, ms, mc, (int)M, m_dev, density_scale,
                                                                 (const __half*)fwd_s, (const __half*)fwd_c, (__half*)bwd_s, (__half*)bwd_c,
                                                                 (__half*)gout_s, (__half*)gout_c, (__half*)dfeat);
    return launch_status();
}

extern "C" int seald_field_heads_backward(const float* grad_sigma, const float* grad_rgb, const float* rgb, const void* hs,
                                          const void* const* w_sigma, int n_sigma, const void* const* w_color, int n_color, uint32_t M,
                                          const int32_t* m_dev, float density_scale, const void* fwd_s, const void* fwd_c, void* bwd_s, void* bwd_c,
                                          void* gout_s, void* gout_c, void* dfeat, seald_stream_t stream) {
    return heads_backward_impl(grad_sigma, grad_rgb, rgb, hs, w_sigma, n_sigma, w_color, n_color, M, m_dev, density_scale, fwd_s, fwd_c, bwd_s, bwd_c,
                               gout_s, gout_c, dfeat, false, stream);
}

extern "C" int seald_field_heads_backward_tiled(const float* grad_sigma, const float* grad_rgb, const float* rgb, const void* hs,
                                                const void* const* w_sigma, int n_sigma, const void* const* w_color, int n_color, uint32_t M,
                                                const int32_t* m_dev, float density_scale, const void* fwd_s, const void* fwd_c, void* bwd_s,
                                                void* bwd_c, void* gout_s, void* gout_c, void* dfeat, seald_stream_t stream) {
    return heads_backward_impl(grad_sigma, grad_rgb, rgb, hs, w_sigma, n_sigma, w_color, n_color, M, m_dev, density_scale, fwd_s, fwd_c, bwd_s, bwd_c,
                               gout_s, gout_c, dfeat, true, stream);
}

extern "C" int seald_mlp_wgrad(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, seald_stream_t stream) {
    if (M == 0 || n_jobs == 0) return 0;
    if (!jobs || n_jobs < 0 || n_jobs > kMaxWgradJobs) return SEALD_E_BADARG;
    WgradJobs js;
    js.n_jobs = n_jobs;
    int maxNK = 0;
    for (int i = 0; i < n_jobs; i++) {
        const seald_wgrad_job& a = jobs[i];
        if (!a.G || !a.A || !a.dW) return SEALD_E_BADARG;
        if ((a.N != 16 && a.N != 32 && a.N != 64 && a.N != 128) || a.K % 16 || a.K <= 0 || a.K > 128 || a.ldg % 8 || a.lda % 8 || a.ldg == 0 || a.lda == 0) return SEALD_E_UNSUPPORTED;  // (ld 0 = tile images: wgrad_umma.cu only)
        if (((uintptr_t)a.G % 16) || ((uintptr_t)a.A % 16)) return SEALD_E_ALIGN;
        WgradJob& j = js.j[i];
        j.G = (const __half*)a.G; j.A = (const __half*)a.A; j.dW = a.dW;
        j.N = a.N; j.K = a.K; j.ldg = a.ldg; j.lda = a.lda; j.ldw = a.ldw; j.n_real = a.n_real; j.k_real = a.k_real;
        if (a.N + a.K > maxNK) maxNK = a.N + a.K;
    }
    const size_t smem = (size_t)kWgStages * kWgChunk * (maxNK + 2 * kPad) * sizeof(__half);
    int rc = set_smem(k_wgrad, smem);
    if (rc) return rc;
    // ONE balanced wave: 2 resident CTAs per SM (98 registers x 256 threads), shared out over the jobs in proportion to the
    // bytes a job streams per row (N + K halves); but never fewer than ~256 rows per CTA (the 64 KiB atomic flush must amortise)
    int cost = 0;
    for (int i = 0; i < n_jobs; i++) cost += js.j[i].N + js.j[i].K;
    int budget = 2 * SEALD_NUM_SMS;
    const int max_by_rows = (int)div_up(M, 256u) * n_jobs;
    if (budget > max_by_rows) budget = max_by_rows;
    if (budget < n_jobs) budget = n_jobs;
    int total = 0;
    for (int i = 0; i < n_jobs; i++) {
        int parts = (int)((long long)budget * (js.j[i].N + js.j[i].K) / cost);
        if (parts < 1) parts = 1;
        js.first_cta[i] = total;
        total += parts;
    }
    js.first_cta[n_jobs] = total;
    k_wgrad<<<total, 256, smem, to_stream(stream)>>>(js, (int)M, m_dev);
    return launch_status();
}
