// grid_encoder.cu — multiresolution hash / tiled grid encoder for sm_100a.
//
// Semantics follow the reference's gridencoder extension (gridencoder/src/gridencoder.cu:50-84 index/hash,
// :137-159 scale/pos, :166-191 interpolation, :201-244 dy_dx, :283-339 scatter, :344-369 input grad), but the
// kernels are laid out differently:
//   * forward : one thread per point, looping over levels, all 2^D corner gathers of a level group issued
//               before use (memory-level parallelism), row-vector loads (half2 / uint2 / uint4), fp32
//               accumulation, output written straight into [B, L*C] with 16-byte stores (no [L,B,C] + permute).
//   * backward: grid = (point chunks, levels), level fastest.  Every corner contribution is a fire-and-forget vector reduction
//               (red.global.add.v2/.v4.f32) into the fp32 gradient table; on the coarse levels the lanes of a warp that hit the
//               same row are summed first (match.any) so contended rows see one reduction per warp.  Input gradients are
//               recomputed from the table with fp32 accumulation (or taken from dy_dx when the caller kept it).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "grid_device.cuh"

namespace seald {

// =================================================================================================
// forward
// =================================================================================================
template <typename T, uint32_t D, uint32_t C, uint32_t G, bool DYDX>
__global__ void __launch_bounds__(256) k_grid_forward(const float* __restrict__ inputs, const T* __restrict__ table,
                                                      const int* __restrict__ offsets, T* __restrict__ outputs,
                                                      T* __restrict__ dy_dx, const uint32_t B, const uint32_t L,
                                                      const float S, const uint32_t H, const uint32_t gridtype,
                                                      const bool align_corners, const uint32_t interp, const int* __restrict__ b_dev) {
    __shared__ LevelParams s_lp[kMaxLevels];
    if (threadIdx.x < L) s_lp[threadIdx.x] = make_level(offsets, threadIdx.x, S, H, D, gridtype, align_corners);
    __syncthreads();
    const uint32_t Bn = b_dev ? min(B, (uint32_t)max(*b_dev, 0)) : B;  // live rows

    constexpr uint32_t NC = 1u << D;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < Bn; b += gridDim.x * blockDim.x) {
        float x[D];
        bool oob = false;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            x[d] = inputs[(size_t)b * D + d];
            if (x[d] < 0 || x[d] > 1) oob = true;
        }
        T* out_row = outputs + (size_t)b * L * C;
        T* dy_row = DYDX ? dy_dx + (size_t)b * L * D * C : nullptr;

        for (uint32_t l0 = 0; l0 < L; l0 += G) {
            float res[G][C];
            float gres[G][D][C];
            float pos[G][D], deriv[G][D];
            float val[G][NC][C];
            float scale[G];

            if (!oob) {
                // issue every gather of the level group first
#pragma unroll
                for (uint32_t g = 0; g < G; g++) {
                    const LevelParams lp = s_lp[l0 + g];
                    scale[g] = lp.scale;
                    uint32_t pos_grid[D];
                    locate<D>(x, lp, align_corners, interp, pos[g], deriv[g], pos_grid);
                    const T* tl = table + (size_t)lp.offset * C;
#pragma unroll
                    for (uint32_t idx = 0; idx < NC; idx++) {
                        uint32_t pg[D];
#pragma unroll
                        for (uint32_t d = 0; d < D; d++) pg[d] = pos_grid[d] + ((idx >> d) & 1u);
                        const uint32_t row = grid_row<D>(lp, pg);
                        Row<T, C>::load(tl + (size_t)row * C, val[g][idx]);
                    }
                }
            }
#pragma unroll
            for (uint32_t g = 0; g < G; g++) {
#pragma unroll
                for (uint32_t c = 0; c < C; c++) res[g][c] = 0.0f;
#pragma unroll
                for (uint32_t d = 0; d < D; d++)
#pragma unroll
                    for (uint32_t c = 0; c < C; c++) gres[g][d][c] = 0.0f;
                if (oob) continue;
#pragma unroll
                for (uint32_t idx = 0; idx < NC; idx++) {
                    float w = 1;
#pragma unroll
                    for (uint32_t d = 0; d < D; d++) w *= ((idx >> d) & 1u) ? pos[g][d] : 1 - pos[g][d];
#pragma unroll
                    for (uint32_t c = 0; c < C; c++) res[g][c] += w * val[g][idx][c];
                }
                if constexpr (DYDX) {
#pragma unroll
                    for (uint32_t gd = 0; gd < D; gd++) {
#pragma unroll
                        for (uint32_t idx = 0; idx < NC; idx++) {
                            if ((idx >> gd) & 1u) continue;  // idx = "left" corner in dim gd
                            float w = scale[g];
#pragma unroll
                            for (uint32_t d = 0; d < D; d++)
                                if (d != gd) w *= ((idx >> d) & 1u) ? pos[g][d] : 1 - pos[g][d];
#pragma unroll
                            for (uint32_t c = 0; c < C; c++)
                                gres[g][gd][c] += w * (val[g][idx | (1u << gd)][c] - val[g][idx][c]) * deriv[g][gd];
                        }
                    }
                }
            }
            // ---- store outputs of this level group
            if constexpr (G * C * sizeof(T) == 16) {
                T packed[G * C];
#pragma unroll
                for (uint32_t g = 0; g < G; g++) Row<T, C>::store(packed + g * C, res[g]);
                *reinterpret_cast<uint4*>(out_row + (size_t)l0 * C) = *reinterpret_cast<const uint4*>(packed);
            } else {
#pragma unroll
                for (uint32_t g = 0; g < G; g++) Row<T, C>::store(out_row + (size_t)(l0 + g) * C, res[g]);
            }
            if constexpr (DYDX) {
#pragma unroll
                for (uint32_t g = 0; g < G; g++)
#pragma unroll
                    for (uint32_t d = 0; d < D; d++) Row<T, C>::store(dy_row + ((size_t)(l0 + g) * D + d) * C, gres[g][d]);
            }
        }
    }
}


template <typename T, uint32_t D, uint32_t C, uint32_t G, bool DYDX>
__global__ void __launch_bounds__(256) k_grid_forward_pair(const float* __restrict__ inputs, const T* __restrict__ table,
                                                           const int* __restrict__ offsets, T* __restrict__ outputs,
                                                           T* __restrict__ dy_dx, const uint32_t B, const uint32_t L,
                                                           const float S, const uint32_t H, const uint32_t gridtype,
                                                           const bool align_corners, const uint32_t interp, const int* __restrict__ b_dev) {
    __shared__ LevelParams s_lp[kMaxLevels];
    if (threadIdx.x < L) s_lp[threadIdx.x] = make_level(offsets, threadIdx.x, S, H, D, gridtype, align_corners);
    __syncthreads();
    const uint32_t Bn = b_dev ? min(B, (uint32_t)max(*b_dev, 0)) : B;  // live rows
    constexpr uint32_t NC = 1u << D;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < Bn; b += gridDim.x * blockDim.x) {
        float x[D];
        bool oob = false;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            x[d] = inputs[(size_t)b * D + d];
            if (x[d] < 0 || x[d] > 1) oob = true;
        }
        T* out_row = outputs + (size_t)b * L * C;
        T* dy_row = DYDX ? dy_dx + (size_t)b * L * D * C : nullptr;
        for (uint32_t l0 = 0; l0 < L; l0 += G) {
            float res[G][C];
            float pos[G][D], deriv[G][D];
            CellGather<T, D, C> cg[G];
            if (!oob) {
#pragma unroll
                for (uint32_t g = 0; g < G; g++) {
                    const LevelParams lp = s_lp[l0 + g];
                    uint32_t pos_grid[D];
                    locate<D>(x, lp, align_corners, interp, pos[g], deriv[g], pos_grid);
                    cg[g].issue(table, gridtype, align_corners, lp, pos_grid);
                }
            }
#pragma unroll
            for (uint32_t g = 0; g < G; g++) {
                float gres[D][C];
#pragma unroll
                for (uint32_t c = 0; c < C; c++) res[g][c] = 0.0f;
#pragma unroll
                for (uint32_t d = 0; d < D; d++)
#pragma unroll
                    for (uint32_t c = 0; c < C; c++) gres[d][c] = 0.0f;
                if (!oob) {
                    float val[NC][C];
                    cg[g].resolve(val);
#pragma unroll
                    for (uint32_t idx = 0; idx < NC; idx++) {
                        float w = 1;
#pragma unroll
                        for (uint32_t d = 0; d < D; d++) w *= ((idx >> d) & 1u) ? pos[g][d] : 1 - pos[g][d];
#pragma unroll
                        for (uint32_t c = 0; c < C; c++) res[g][c] += w * val[idx][c];
                    }
                    if constexpr (DYDX) {
                        const float scale = s_lp[l0 + g].scale;
#pragma unroll
                        for (uint32_t gd = 0; gd < D; gd++) {
#pragma unroll
                            for (uint32_t idx = 0; idx < NC; idx++) {
                                if ((idx >> gd) & 1u) continue;
                                float w = scale;
#pragma unroll
                                for (uint32_t d = 0; d < D; d++)
                                    if (d != gd) w *= ((idx >> d) & 1u) ? pos[g][d] : 1 - pos[g][d];
#pragma unroll
                                for (uint32_t c = 0; c < C; c++)
                                    gres[gd][c] += w * (val[idx | (1u << gd)][c] - val[idx][c]) * deriv[g][gd];
                            }
                        }
                    }
                }
                if constexpr (DYDX) {
#pragma unroll
                    for (uint32_t d = 0; d < D; d++) Row<T, C>::store(dy_row + ((size_t)(l0 + g) * D + d) * C, gres[d]);
                }
            }
            if constexpr (G * C * sizeof(T) == 16) {
                T packed[G * C];
#pragma unroll
                for (uint32_t g = 0; g < G; g++) Row<T, C>::store(packed + g * C, res[g]);
                *reinterpret_cast<uint4*>(out_row + (size_t)l0 * C) = *reinterpret_cast<const uint4*>(packed);
            } else {
#pragma unroll
                for (uint32_t g = 0; g < G; g++) Row<T, C>::store(out_row + (size_t)(l0 + g) * C, res[g]);
            }
        }
    }
}

// =================================================================================================
// backward
// =================================================================================================
template <typename T, int C>
struct VecAtomic;

template <int C>
struct VecAtomic<float, C> {
    static __device__ __forceinline__ void add(float* p, const float (&v)[C]) {
        if constexpr (C % 2 == 0) {
#pragma unroll
            for (int c = 0; c < C; c += 2) atomicAdd(reinterpret_cast<float2*>(p + c), make_float2(v[c], v[c + 1]));  // red.global.add.v2.f32
        } else {
#pragma unroll
            for (int c = 0; c < C; c++) atomicAdd(p + c, v[c]);
        }
    }
};

template <int C>
struct VecAtomic<__half, C> {
    static __device__ __forceinline__ void add(__half* p, const float (&v)[C]) {
        if constexpr (C % 2 == 0) {
#pragma unroll
            for (int c = 0; c < C; c += 2) atomicAdd(reinterpret_cast<__half2*>(p + c), __floats2half2_rn(v[c], v[c + 1]));
        } else {
#pragma unroll
            for (int c = 0; c < C; c++) atomicAdd(p + c, __float2half_rn(v[c]));
        }
    }
};

// Sum `v` over the lanes of the warp that target the same row; returns true on the lane that must
// issue the atomic for the group.
template <int C>
__device__ __forceinline__ bool warp_aggregate(const uint32_t key, float (&v)[C], const uint32_t lane) {
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const uint32_t nmax = __reduce_max_sync(0xffffffffu, (uint32_t)__popc(peers));
    if (nmax == 1) return true;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.0f;
    uint32_t rem = peers;
    for (uint32_t i = 0; i < nmax; i++) {
        const int src = rem ? (__ffs(rem) - 1) : (int)lane;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const float o = __shfl_sync(0xffffffffu, v[c], src);
            if (rem) acc[c] += o;
        }
        rem &= rem - 1;
    }
#pragma unroll
    for (int c = 0; c < C; c++) v[c] = acc[c];
    return lane == (uint32_t)(__ffs(peers) - 1);
}

// Scatter of the table gradient.  One CTA = (chunk of points, level), the level varying fastest over blockIdx.x so the CTAs
// that re-read one chunk's grad rows run together (L2 hits).  Every corner contribution is a fire-and-forget vector reduction
// (red.global.add.v2.f32); the two corners that differ in dimension 0 go out as ONE red.global.add.v4.f32 when they are the two
// halves of an aligned 16-byte group (rows r, r ^ 1: every dense-level cell with even r, every hashed cell with even x).
// The `agg_levels` coarsest levels sum the lanes of a warp that hit the same row first (match.any): consecutive samples of a ray
// share coarse cells.  found_inf (optional): set to 1.0f's bit pattern when a consumed grad element is inf/nan — the table
// gradient is non-finite exactly when one of them is (weights are in [0,1], the sums are fp32), so the optimiser's overflow
// check does not have to re-read the whole table gradient.
template <typename T, typename TG, uint32_t D, uint32_t C, bool V4>
__device__ __forceinline__ void grid_scatter_body(const uint32_t bid, const T* __restrict__ grad, const float* __restrict__ inputs,
                                                      const int* __restrict__ offsets, TG* __restrict__ grad_table,
                                                      const uint32_t B, const uint32_t L, const float S, const uint32_t H,
                                                      const uint32_t gridtype, const bool align_corners, const uint32_t interp,
                                                      const uint32_t points_per_cta, const uint32_t agg_levels,
                                                      const int* __restrict__ b_dev, int* __restrict__ found_inf) {
    const uint32_t level = bid % L;
    const uint32_t chunk = bid / L;
    const uint32_t Bn = b_dev ? min(B, (uint32_t)max(*b_dev, 0)) : B;  // live rows
    const uint32_t b_begin = chunk * points_per_cta;
    const uint32_t b_end = min(Bn, b_begin + points_per_cta);
    if (b_begin >= b_end) return;
    const LevelParams lp = make_level(offsets, level, S, H, D, gridtype, align_corners);
    const bool agg = level < agg_levels;
    constexpr uint32_t NP = 1u << (D - 1);
    const uint32_t lane = threadIdx.x & 31u;
    TG* gl = grad_table + (size_t)lp.offset * C;
    bool bad = false;

    // all lanes of a warp iterate together (warp_aggregate needs the full warp)
    for (uint32_t b0 = b_begin + (threadIdx.x & ~31u); b0 < b_end; b0 += blockDim.x) {
        const uint32_t b = b0 + lane;
        bool valid = b < b_end;
        float x[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            x[d] = valid ? inputs[(size_t)b * D + d] : 0.0f;
            if (x[d] < 0 || x[d] > 1) valid = false;  // gridencoder.cu:276-281
        }
        float g[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) g[c] = 0.0f;
        if (valid) {
            Row<T, C>::load(grad + ((size_t)b * L + level) * C, g);
#pragma unroll
            for (uint32_t c = 0; c < C; c++) bad |= !isfinite(g[c]);
        }
        float pos[D], deriv[D];
        uint32_t pos_grid[D];
        locate<D>(x, lp, align_corners, interp, pos, deriv, pos_grid);
        uint32_t rows0[NP], rows1[NP];
        pair_rows<D>(lp, pos_grid, rows0, rows1);
#pragma unroll
        for (uint32_t j = 0; j < NP; j++) {
            const uint32_t r0 = rows0[j], r1 = rows1[j];
            // same product order as the reference (w = 1 * w_0 * w_1 * ...)
            float w0 = 1 - pos[0], w1 = pos[0];
#pragma unroll
            for (uint32_t d = 1; d < D; d++) {
                const float f = ((j >> (d - 1)) & 1u) ? pos[d] : 1 - pos[d];
                w0 *= f; w1 *= f;
            }
            float a0[C], a1[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) { a0[c] = w0 * g[c]; a1[c] = w1 * g[c]; }
            if (agg) {
                const bool lead0 = warp_aggregate<C>(valid ? r0 : (0xffffffffu - lane), a0, lane);
                if (valid && lead0) VecAtomic<TG, C>::add(gl + (size_t)r0 * C, a0);
                const bool lead1 = warp_aggregate<C>(valid ? r1 : (0xffffffffu - lane), a1, lane);
                if (valid && lead1) VecAtomic<TG, C>::add(gl + (size_t)r1 * C, a1);
            } else if (valid) {
                if constexpr (V4) {  // TG = float, C = 2, 16-byte aligned table
                    const uint32_t ra0 = lp.offset + r0, ra1 = lp.offset + r1;
                    if ((ra0 ^ ra1) == 1u) {
                        const bool odd = ra0 & 1u;
                        const float4 v = odd ? make_float4(a1[0], a1[1], a0[0], a0[1]) : make_float4(a0[0], a0[1], a1[0], a1[1]);
                        atomicAdd(reinterpret_cast<float4*>(grad_table + (size_t)(ra0 & ~1u) * C), v);  // red.global.add.v4.f32
                    } else {
                        VecAtomic<TG, C>::add(gl + (size_t)r0 * C, a0);
                        VecAtomic<TG, C>::add(gl + (size_t)r1 * C, a1);
                    }
                } else {
                    VecAtomic<TG, C>::add(gl + (size_t)r0 * C, a0);
                    VecAtomic<TG, C>::add(gl + (size_t)r1 * C, a1);
                }
            }
        }
    }
    if (found_inf && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(found_inf, 0x3f800000);
}

template <typename T, typename TG, uint32_t D, uint32_t C, bool V4>
__global__ void __launch_bounds__(256) k_grid_scatter(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                      const int* __restrict__ offsets, TG* __restrict__ grad_table,
                                                      const uint32_t B, const uint32_t L, const float S, const uint32_t H,
                                                      const uint32_t gridtype, const bool align_corners, const uint32_t interp,
                                                      const uint32_t points_per_cta, const uint32_t agg_levels,
                                                      const int* __restrict__ b_dev, int* __restrict__ found_inf) {
    grid_scatter_body<T, TG, D, C, V4>(blockIdx.x, grad, inputs, offsets, grad_table, B, L, S, H, gridtype, align_corners, interp, points_per_cta, agg_levels, b_dev, found_inf);
}

// grad_x[b, d] = sum_{l,c} grad[b,l,c] * d out[b,l,c] / d x[b,d]; recomputed from the table (fp32 accumulate)
template <typename T, uint32_t D, uint32_t C, bool PAIR>
__device__ __forceinline__ void grid_input_backward_body(const uint32_t bid, const uint32_t nblocks, const T* __restrict__ grad, const float* __restrict__ inputs,
                                                                       const T* __restrict__ table, const int* __restrict__ offsets,
                                                                       float* __restrict__ grad_x, const uint32_t B, const uint32_t L,
                                                                       const float S, const uint32_t H, const uint32_t gridtype,
                                                                       const bool align_corners, const uint32_t interp,
                                                                       const int* __restrict__ b_dev) {
    __shared__ LevelParams s_lp[kMaxLevels];
    if (threadIdx.x < L) s_lp[threadIdx.x] = make_level(offsets, threadIdx.x, S, H, D, gridtype, align_corners);
    __syncthreads();
    const uint32_t Bn = b_dev ? min(B, (uint32_t)max(*b_dev, 0)) : B;
    constexpr uint32_t NC = 1u << D;
    for (uint32_t b = bid * blockDim.x + threadIdx.x; b < Bn; b += nblocks * blockDim.x) {
        float x[D];
        bool oob = false;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            x[d] = inputs[(size_t)b * D + d];
            if (x[d] < 0 || x[d] > 1) oob = true;
        }
        float gx[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) gx[d] = 0.0f;
        if (!oob) {
#pragma unroll 2
            for (uint32_t l = 0; l < L; l++) {
                const LevelParams lp = s_lp[l];
                float pos[D], deriv[D];
                uint32_t pos_grid[D];
                locate<D>(x, lp, align_corners, interp, pos, deriv, pos_grid);
                float val[NC][C];
                if constexpr (PAIR) {
                    CellGather<T, D, C> cg;
                    cg.issue(table, gridtype, align_corners, lp, pos_grid);
                    cg.resolve(val);
                } else {
                    const T* tl = table + (size_t)lp.offset * C;
#pragma unroll
                    for (uint32_t idx = 0; idx < NC; idx++) {
                        uint32_t pg[D];
#pragma unroll
                        for (uint32_t d = 0; d < D; d++) pg[d] = pos_grid[d] + ((idx >> d) & 1u);
                        const uint32_t row = grid_row<D>(lp, pg);
                        Row<T, C>::load(tl + (size_t)row * C, val[idx]);
                    }
                }
                float g[C];
                Row<T, C>::load(grad + ((size_t)b * L + l) * C, g);
#pragma unroll
                for (uint32_t gd = 0; gd < D; gd++) {
                    float acc = 0.0f;
#pragma unroll
                    for (uint32_t idx = 0; idx < NC; idx++) {
                        if ((idx >> gd) & 1u) continue;
                        float w = lp.scale;
#pragma unroll
                        for (uint32_t d = 0; d < D; d++)
                            if (d != gd) w *= ((idx >> d) & 1u) ? pos[d] : 1 - pos[d];
#pragma unroll
                        for (uint32_t c = 0; c < C; c++) acc += w * (val[idx | (1u << gd)][c] - val[idx][c]) * g[c];
                    }
                    gx[gd] += acc * deriv[gd];
                }
            }
        }
#pragma unroll
        for (uint32_t d = 0; d < D; d++) grad_x[(size_t)b * D + d] = gx[d];
    }
}

// gridencoder.cu:344-369 with fp32 accumulation and the [B, L*C] grad layout.
template <typename T, uint32_t D, uint32_t C, bool PAIR>
__global__ void __launch_bounds__(256) k_grid_input_backward_recompute(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                                       const T* __restrict__ table, const int* __restrict__ offsets,
                                                                       float* __restrict__ grad_x, const uint32_t B, const uint32_t L,
                                                                       const float S, const uint32_t H, const uint32_t gridtype,
                                                                       const bool align_corners, const uint32_t interp,
                                                                       const int* __restrict__ b_dev) {
    grid_input_backward_body<T, D, C, PAIR>(blockIdx.x, gridDim.x, grad, inputs, table, offsets, grad_x, B, L, S, H, gridtype, align_corners, interp, b_dev);
}

// Table scatter and input gradient of one backward pass in ONE launch ("horizontal" fusion): CTAs [0, n_input) run the per-point input
// gradient (all levels of a point, paired gathers), the others the per-(chunk, level) scatter.  The two are independent, both latency
// bound at training-batch size, and a kernel boundary costs ~5 us inside the step graph.
template <typename T, typename TG, uint32_t D, uint32_t C, bool V4, bool PAIR>
__global__ void __launch_bounds__(256) k_grid_backward_both(const uint32_t n_input, const T* __restrict__ table, float* __restrict__ grad_x,
                                                            const T* __restrict__ grad, const float* __restrict__ inputs,
                                                      const int* __restrict__ offsets, TG* __restrict__ grad_table,
                                                      const uint32_t B, const uint32_t L, const float S, const uint32_t H,
                                                      const uint32_t gridtype, const bool align_corners, const uint32_t interp,
                                                      const uint32_t points_per_cta, const uint32_t agg_levels,
                                                      const int* __restrict__ b_dev, int* __restrict__ found_inf) {
    if (blockIdx.x < n_input) {
        grid_input_backward_body<T, D, C, PAIR>(blockIdx.x, n_input, grad, inputs, table, offsets, grad_x, B, L, S, H, gridtype, align_corners, interp,
                                                b_dev);
    } else {
        grid_scatter_body<T, TG, D, C, V4>(blockIdx.x - n_input, grad, inputs, offsets, grad_table, B, L, S, H, gridtype, align_corners, interp, points_per_cta, agg_levels, b_dev, found_inf);
    }
}

template <typename T, uint32_t D, uint32_t C>
__global__ void k_grid_input_backward_dydx(const T* __restrict__ grad, const T* __restrict__ dy_dx, float* __restrict__ grad_x,
                                           const uint32_t B, const uint32_t L, const int* __restrict__ b_dev) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    const uint32_t Bn = b_dev ? min(B, (uint32_t)max(*b_dev, 0)) : B;
    if (t >= Bn * D) return;
    const uint32_t b = t / D;
    const uint32_t d = t - b * D;
    const T* dy = dy_dx + (size_t)b * L * D * C;
    const T* g = grad + (size_t)b * L * C;
    float result = 0;
    for (uint32_t l = 0; l < L; l++) {
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++) result += (float)g[l * C + ch] * (float)dy[l * D * C + d * C + ch];
    }
    grad_x[t] = result;
}

template <uint32_t D>
__global__ void k_grid_debug_indices(const float* __restrict__ inputs, const int* __restrict__ offsets, uint32_t* __restrict__ indices,
                                     float* __restrict__ scales, uint32_t* __restrict__ resolutions, const uint32_t B, const uint32_t L,
                                     const float S, const uint32_t H, const uint32_t gridtype, const bool align_corners) {
    constexpr uint32_t NC = 1u << D;
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t < L) {
        const LevelParams lp = make_level(offsets, t, S, H, D, gridtype, align_corners);
        scales[t] = lp.scale;
        resolutions[t] = lp.resolution;
    }
    if (t >= B * L) return;
    const uint32_t b = t / L, l = t - b * L;
    const LevelParams lp = make_level(offsets, l, S, H, D, gridtype, align_corners);
    float x[D], pos[D], deriv[D];
    uint32_t pos_grid[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) x[d] = inputs[(size_t)b * D + d];
    locate<D>(x, lp, align_corners, 0, pos, deriv, pos_grid);
#pragma unroll
    for (uint32_t idx = 0; idx < NC; idx++) {
        uint32_t pg[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) pg[d] = pos_grid[d] + ((idx >> d) & 1u);
        indices[(size_t)t * NC + idx] = grid_row<D>(lp, pg);
    }
}

// ---- host dispatch ------------------------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
int launch_forward(const float* x, const void* table, const int* offsets, void* out, void* dy_dx, uint32_t B, uint32_t L, float S,
                   uint32_t H, uint32_t gridtype, bool align, uint32_t interp, const int* b_dev, cudaStream_t st) {
    constexpr uint32_t G16 = 16 / (C * sizeof(T));  // levels per 16-byte output store
    // levels per pass: all gathers of G levels are in flight before the first use; 4-D cells have 16 corners, so two levels fill the registers
    constexpr uint32_t G = (G16 >= 2 && G16 <= 4) ? (D <= 3 ? G16 : (D == 4 ? 2u : 1u)) : 1;
    const uint32_t threads = 256;
    const uint32_t blocks = div_up(B, threads);
    if constexpr (G > 1) {
        const bool vec_ok = (L % G) == 0 && ((L * C * sizeof(T)) % 16 == 0) && ((uintptr_t)out % 16 == 0);
        if constexpr (RowPack<T, C>::ok && D >= 2 && D <= 4) {
            if (vec_ok && ((uintptr_t)table % 16 == 0)) {
                if (dy_dx) k_grid_forward_pair<T, D, C, G, true><<<blocks, threads, 0, st>>>(x, (const T*)table, offsets, (T*)out, (T*)dy_dx, B, L, S, H, gridtype, align, interp, b_dev);
                else k_grid_forward_pair<T, D, C, G, false><<<blocks, threads, 0, st>>>(x, (const T*)table, offsets, (T*)out, (T*)dy_dx, B, L, S, H, gridtype, align, interp, b_dev);
                return launch_status();
            }
        }
        if (vec_ok) {
            if (dy_dx) k_grid_forward<T, D, C, G, true><<<blocks, threads, 0, st>>>(x, (const T*)table, offsets, (T*)out, (T*)dy_dx, B, L, S, H, gridtype, align, interp, b_dev);
            else k_grid_forward<T, D, C, G, false><<<blocks, threads, 0, st>>>(x, (const T*)table, offsets, (T*)out, (T*)dy_dx, B, L, S, H, gridtype, align, interp, b_dev);
            return launch_status();
        }
    }
    if (dy_dx) k_grid_forward<T, D, C, 1, true><<<blocks, threads, 0, st>>>(x, (const T*)table, offsets, (T*)out, (T*)dy_dx, B, L, S, H, gridtype, align, interp, b_dev);
    else k_grid_forward<T, D, C, 1, false><<<blocks, threads, 0, st>>>(x, (const T*)table, offsets, (T*)out, (T*)dy_dx, B, L, S, H, gridtype, align, interp, b_dev);
    return launch_status();
}

template <typename T, uint32_t D>
int dispatch_forward_C(uint32_t C, const float* x, const void* table, const int* offsets, void* out, void* dy_dx, uint32_t B, uint32_t L,
                       float S, uint32_t H, uint32_t gridtype, bool align, uint32_t interp, const int* b_dev, cudaStream_t st) {
    switch (C) {
        case 1: return launch_forward<T, D, 1>(x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 2: return launch_forward<T, D, 2>(x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 4: return launch_forward<T, D, 4>(x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 8: return launch_forward<T, D, 8>(x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        default: return SEALD_E_UNSUPPORTED;
    }
}

template <typename T>
int dispatch_forward_D(uint32_t D, uint32_t C, const float* x, const void* table, const int* offsets, void* out, void* dy_dx, uint32_t B,
                       uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align, uint32_t interp, const int* b_dev, cudaStream_t st) {
    switch (D) {
        case 2: return dispatch_forward_C<T, 2>(C, x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 3: return dispatch_forward_C<T, 3>(C, x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 4: return dispatch_forward_C<T, 4>(C, x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 5: return dispatch_forward_C<T, 5>(C, x, table, offsets, out, dy_dx, B, L, S, H, gridtype, align, interp, b_dev, st);
        default: return SEALD_E_UNSUPPORTED;
    }
}

// Number of coarse levels whose contributions are summed inside the warp before the reduction goes out: levels whose cells are
// large against the sample spacing of a marched ray (resolution <= 128: a ray at dt = 2*sqrt(3)/1024 puts >= 2 consecutive
// samples in a cell).  Uniformly random point sets (the encoder benchmark) gain nothing from it; SEALD_GRID_AGG_LEVELS overrides.
inline uint32_t default_agg_levels(uint32_t B, uint32_t L, float S, uint32_t H) {
    static int env = -2;
    if (env == -2) {
        const char* e = getenv("SEALD_GRID_AGG_LEVELS");
        env = e ? atoi(e) : -1;
    }
    if (env >= 0) return (uint32_t)env;
    if (B >= (1u << 18)) return 0;
    uint32_t n = 0;
    for (uint32_t l = 0; l < L; l++)
        if (exp2f(l * S) * H <= 128.0f) n = l + 1;
    return n;
}

template <typename T, typename TG, uint32_t D, uint32_t C>
int launch_scatter(const void* grad, const float* x, const int* offsets, void* grad_table, uint32_t B, uint32_t L, float S, uint32_t H,
                   uint32_t gridtype, bool align, uint32_t interp, const int* b_dev, int* found_inf, cudaStream_t st) {
    static int env_ppc = -1;
    if (env_ppc < 0) {
        const char* e = getenv("SEALD_GRID_SCATTER_PPC");
        env_ppc = e ? atoi(e) : 0;
    }
    // a training batch (tens of thousands of samples) is latency bound: one point per thread puts 4x more CTAs in flight (measured
    // 0.058 -> 0.032 ms at 31.7k samples); the 2^22-point sweeps amortise the level set-up over 4 points per thread
    const uint32_t ppc = env_ppc > 0 ? (uint32_t)env_ppc : (B < (1u << 18) ? 256u : 1024u);
    const uint32_t grid = div_up(B, ppc) * L;
    const uint32_t agg = default_agg_levels(B, L, S, H);
    if constexpr (std::is_same<TG, float>::value && C == 2) {
        if ((uintptr_t)grad_table % 16 == 0) {
            k_grid_scatter<T, TG, D, C, true><<<grid, 256, 0, st>>>((const T*)grad, x, offsets, (TG*)grad_table, B, L, S, H, gridtype, align, interp, ppc, agg, b_dev, found_inf);
            return launch_status();
        }
    }
    k_grid_scatter<T, TG, D, C, false><<<grid, 256, 0, st>>>((const T*)grad, x, offsets, (TG*)grad_table, B, L, S, H, gridtype, align, interp, ppc, agg, b_dev, found_inf);
    return launch_status();
}

template <typename T, uint32_t D, uint32_t C>
int launch_input_backward(const void* grad, const float* x, const void* table, const int* offsets, const void* dy_dx, float* grad_x, uint32_t B,
                          uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align, uint32_t interp, const int* b_dev, cudaStream_t st) {
    if (dy_dx) {
        k_grid_input_backward_dydx<T, D, C><<<div_up(B * D, 256u), 256, 0, st>>>((const T*)grad, (const T*)dy_dx, grad_x, B, L, b_dev);
        return launch_status();
    }
    if constexpr (RowPack<T, C>::ok && D >= 2 && D <= 4) {
        if ((uintptr_t)table % 16 == 0) {
            k_grid_input_backward_recompute<T, D, C, true><<<div_up(B, 256u), 256, 0, st>>>((const T*)grad, x, (const T*)table, offsets, grad_x, B, L, S, H, gridtype, align, interp, b_dev);
            return launch_status();
        }
    }
    k_grid_input_backward_recompute<T, D, C, false><<<div_up(B, 256u), 256, 0, st>>>((const T*)grad, x, (const T*)table, offsets, grad_x, B, L, S, H, gridtype, align, interp, b_dev);
    return launch_status();
}

// grad_table == nullptr: input gradient only; grad_x == nullptr: table gradient only
template <typename T, typename TG, uint32_t D, uint32_t C>
int launch_backward(const void* grad, const float* x, const void* table, const int* offsets, void* grad_table, const void* dy_dx,
                    float* grad_x, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align, uint32_t interp,
                    const int* b_dev, int* found_inf, cudaStream_t st) {
    int rc = 0;
    // both halves, fp32 C = 2 gradient table, paired gathers: ONE launch (k_grid_backward_both); SEALD_GRID_BWD_SPLIT=1 keeps two
    if constexpr (std::is_same<TG, float>::value && C == 2 && RowPack<T, C>::ok && D >= 2 && D <= 4) {
        static const bool split = getenv("SEALD_GRID_BWD_SPLIT") && atoi(getenv("SEALD_GRID_BWD_SPLIT")) != 0;
        if (!split && grad_table && grad_x && !dy_dx && table && (uintptr_t)grad_table % 16 == 0 && (uintptr_t)table % 16 == 0) {
            const uint32_t ppc = B < (1u << 18) ? 256u : 1024u;
            const uint32_t n_scatter = div_up(B, ppc) * L, n_input = div_up(B, 256u);
            const uint32_t agg = default_agg_levels(B, L, S, H);
            k_grid_backward_both<T, TG, D, C, true, true><<<n_input + n_scatter, 256, 0, st>>>(
                n_input, (const T*)table, grad_x, (const T*)grad, x, offsets, (TG*)grad_table, B, L, S, H, gridtype, align, interp, ppc, agg, b_dev,
                found_inf);
            return launch_status();
        }
    }
    if (grad_table) rc = launch_scatter<T, TG, D, C>(grad, x, offsets, grad_table, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
    if (rc) return rc;
    if (grad_x) rc = launch_input_backward<T, D, C>(grad, x, table, offsets, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
    return rc;
}

template <typename T, typename TG, uint32_t D>
int dispatch_backward_C(uint32_t C, const void* grad, const float* x, const void* table, const int* offsets, void* grad_table,
                        const void* dy_dx, float* grad_x, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,
                        uint32_t interp, const int* b_dev, int* found_inf, cudaStream_t st) {
    switch (C) {
        case 1: return launch_backward<T, TG, D, 1>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        case 2: return launch_backward<T, TG, D, 2>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        case 4: return launch_backward<T, TG, D, 4>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        case 8: return launch_backward<T, TG, D, 8>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        default: return SEALD_E_UNSUPPORTED;
    }
}

template <typename T, typename TG>
int dispatch_backward_D(uint32_t D, uint32_t C, const void* grad, const float* x, const void* table, const int* offsets, void* grad_table,
                        const void* dy_dx, float* grad_x, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,
                        uint32_t interp, const int* b_dev, int* found_inf, cudaStream_t st) {
    switch (D) {
        case 2: return dispatch_backward_C<T, TG, 2>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        case 3: return dispatch_backward_C<T, TG, 3>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        case 4: return dispatch_backward_C<T, TG, 4>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        case 5: return dispatch_backward_C<T, TG, 5>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, found_inf, st);
        default: return SEALD_E_UNSUPPORTED;
    }
}

}  // namespace seald

using namespace seald;

extern "C" int seald_grid_encode_forward(const float* x01, const void* table, const int32_t* offsets, void* out, void* dy_dx, uint32_t B,
                                         uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                         uint32_t interp, int dtype, const int32_t* b_dev, seald_stream_t stream) {
    if (B == 0) return 0;
    if (!x01 || !table || !offsets || !out) return SEALD_E_BADARG;
    if (L == 0 || L > kMaxLevels || gridtype > 1 || interp > 1) return SEALD_E_UNSUPPORTED;
    cudaStream_t st = to_stream(stream);
    if (dtype == SEALD_F16) return dispatch_forward_D<__half>(D, C, x01, table, offsets, out, dy_dx, B, L, S, H, gridtype, align_corners != 0, interp, b_dev, st);
    if (dtype == SEALD_F32) return dispatch_forward_D<float>(D, C, x01, table, offsets, out, dy_dx, B, L, S, H, gridtype, align_corners != 0, interp, b_dev, st);
    return SEALD_E_UNSUPPORTED;
}

static int grid_backward_any(const void* grad_out, const float* x01, const void* table, const int32_t* offsets, void* grad_table,
                             const void* dy_dx, float* grad_x, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             uint32_t gridtype, int align_corners, uint32_t interp, int dtype, int grad_table_dtype, const int32_t* b_dev,
                             int32_t* found_inf, seald_stream_t stream) {
    if (B == 0) return 0;
    if (!grad_out || !x01 || !offsets || (!grad_table && !grad_x)) return SEALD_E_BADARG;
    if (grad_x && !dy_dx && !table) return SEALD_E_BADARG;
    if (L == 0 || L > kMaxLevels || gridtype > 1 || interp > 1) return SEALD_E_UNSUPPORTED;
    cudaStream_t st = to_stream(stream);
    const bool al = align_corners != 0;
    if (dtype == SEALD_F16 && grad_table_dtype == SEALD_F16)
        return dispatch_backward_D<__half, __half>(D, C, grad_out, x01, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, al, interp, b_dev, found_inf, st);
    if (dtype == SEALD_F16 && grad_table_dtype == SEALD_F32)
        return dispatch_backward_D<__half, float>(D, C, grad_out, x01, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, al, interp, b_dev, found_inf, st);
    if (dtype == SEALD_F32 && grad_table_dtype == SEALD_F32)
        return dispatch_backward_D<float, float>(D, C, grad_out, x01, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, al, interp, b_dev, found_inf, st);
    return SEALD_E_UNSUPPORTED;
}

extern "C" int seald_grid_encode_backward(const void* grad_out, const float* x01, const void* table, const int32_t* offsets,
                                          void* grad_table, const void* dy_dx, float* grad_x, uint32_t B, uint32_t D, uint32_t C,
                                          uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                                          int dtype, int grad_table_dtype, const int32_t* b_dev, seald_stream_t stream) {
    if (!grad_table) return SEALD_E_BADARG;
    return grid_backward_any(grad_out, x01, table, offsets, grad_table, dy_dx, grad_x, B, D, C, L, S, H, gridtype, align_corners, interp, dtype,
                             grad_table_dtype, b_dev, nullptr, stream);
}

extern "C" int seald_grid_encode_backward_both(const void* grad_out, const float* x01, const void* table, const int32_t* offsets,
                                               void* grad_table, float* grad_x, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                               uint32_t gridtype, int align_corners, uint32_t interp, int dtype, int grad_table_dtype,
                                               const int32_t* b_dev, int32_t* found_inf, seald_stream_t stream) {
    if (!grad_table || !grad_x || !table) return SEALD_E_BADARG;
    return grid_backward_any(grad_out, x01, table, offsets, grad_table, nullptr, grad_x, B, D, C, L, S, H, gridtype, align_corners, interp, dtype,
                             grad_table_dtype, b_dev, found_inf, stream);
}

extern "C" int seald_grid_encode_backward_table(const void* grad_out, const float* x01, const int32_t* offsets, void* grad_table, uint32_t B,
                                                uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                                uint32_t interp, int dtype, int grad_table_dtype, const int32_t* b_dev, int32_t* found_inf,
                                                seald_stream_t stream) {
    if (!grad_table) return SEALD_E_BADARG;
    return grid_backward_any(grad_out, x01, nullptr, offsets, grad_table, nullptr, nullptr, B, D, C, L, S, H, gridtype, align_corners, interp,
                             dtype, grad_table_dtype, b_dev, found_inf, stream);
}

extern "C" int seald_grid_encode_backward_input(const void* grad_out, const float* x01, const void* table, const int32_t* offsets,
                                                const void* dy_dx, float* grad_x, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                                uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                                                const int32_t* b_dev, seald_stream_t stream) {
    if (!grad_x) return SEALD_E_BADARG;
    return grid_backward_any(grad_out, x01, table, offsets, nullptr, dy_dx, grad_x, B, D, C, L, S, H, gridtype, align_corners, interp, dtype,
                             dtype == SEALD_F16 ? SEALD_F16 : SEALD_F32, b_dev, nullptr, stream);
}

extern "C" int seald_grid_debug_indices(const float* x01, const int32_t* offsets, uint32_t* indices, float* scales, uint32_t* resolutions,
                                        uint32_t B, uint32_t D, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                        seald_stream_t stream) {
    if (!x01 || !offsets || !indices || !scales || !resolutions) return SEALD_E_BADARG;
    if (L == 0 || L > kMaxLevels) return SEALD_E_UNSUPPORTED;
    cudaStream_t st = to_stream(stream);
    const uint32_t n = (B * L > L ? B * L : L);
    const uint32_t blocks = div_up(n, 256u);
    const bool al = align_corners != 0;
    switch (D) {
        case 2: k_grid_debug_indices<2><<<blocks, 256, 0, st>>>(x01, offsets, indices, scales, resolutions, B, L, S, H, gridtype, al); break;
        case 3: k_grid_debug_indices<3><<<blocks, 256, 0, st>>>(x01, offsets, indices, scales, resolutions, B, L, S, H, gridtype, al); break;
        case 4: k_grid_debug_indices<4><<<blocks, 256, 0, st>>>(x01, offsets, indices, scales, resolutions, B, L, S, H, gridtype, al); break;
        case 5: k_grid_debug_indices<5><<<blocks, 256, 0, st>>>(x01, offsets, indices, scales, resolutions, B, L, S, H, gridtype, al); break;
        default: return SEALD_E_UNSUPPORTED;
    }
    return launch_status();
}
