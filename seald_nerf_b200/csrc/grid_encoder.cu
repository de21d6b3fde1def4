// grid_encoder.cu — multiresolution hash / tiled grid encoder for sm_100a.
//
// Semantics follow the reference's gridencoder extension (gridencoder/src/gridencoder.cu:50-84 index/hash,
// :137-159 scale/pos, :166-191 interpolation, :201-244 dy_dx, :283-339 scatter, :344-369 input grad), but the
// kernels are laid out differently:
//   * forward : one thread per point, looping over levels, all 2^D corner gathers of a level group issued
//               before use (memory-level parallelism), row-vector loads (half2 / uint2 / uint4), fp32
//               accumulation, output written straight into [B, L*C] with 16-byte stores (no [L,B,C] + permute).
//   * backward: grid = (point chunks, levels).  Levels whose table slice fits in shared memory are
//               accumulated in a CTA-private fp32 copy (shared atomics) and flushed once; the others use
//               warp-aggregated (match.any) vector atomics.  Input gradients are recomputed from the table
//               with fp32 accumulation (or taken from dy_dx when the caller kept it).
#include "common.cuh"

namespace seald {

struct LevelParams {
    float scale;
    uint32_t resolution;
    uint32_t hashmap_size;
    uint32_t offset;  // in rows
};

constexpr int kMaxLevels = 32;

// gridencoder.cu:137-139 — evaluated on device in fp32 with the same expression shape.
__device__ __forceinline__ LevelParams make_level(const int* __restrict__ offsets, uint32_t level, float S, uint32_t H) {
    LevelParams p;
    p.offset = (uint32_t)offsets[level];
    p.hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    p.scale = exp2f(level * S) * H - 1.0f;
    p.resolution = (uint32_t)ceil(p.scale) + 1;
    return p;
}

// gridencoder.cu:50-84
template <uint32_t D>
__device__ __forceinline__ uint32_t grid_row(const uint32_t gridtype, const bool align_corners, const uint32_t hashmap_size,
                                             const uint32_t resolution, const uint32_t pos_grid[D]) {
    uint32_t stride = 1;
    uint32_t index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {
        constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        uint32_t result = 0;
#pragma unroll
        for (uint32_t i = 0; i < D; ++i) result ^= pos_grid[i] * primes[i];
        index = result;
    }
    return index % hashmap_size;
}

// ---- row-vector load/store of C channels --------------------------------------------------------
template <typename T, int C>
struct Row;

template <int C>
struct Row<__half, C> {
    static __device__ __forceinline__ void load(const __half* p, float (&v)[C]) {
        if constexpr (C == 1) {
            v[0] = __half2float(__ldg(p));
        } else if constexpr (C == 2) {
            const __half2 h = __ldg(reinterpret_cast<const __half2*>(p));
            const float2 f = __half22float2(h);
            v[0] = f.x; v[1] = f.y;
        } else if constexpr (C == 4) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                v[2 * i] = a.x; v[2 * i + 1] = a.y;
            }
        }
    }
    static __device__ __forceinline__ void store(__half* p, const float (&v)[C]) {
#pragma unroll
        for (int c = 0; c < C; c++) p[c] = __float2half_rn(v[c]);
    }
};

template <int C>
struct Row<float, C> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[C]) {
        if constexpr (C == 1) {
            v[0] = __ldg(p);
        } else if constexpr (C == 2) {
            const float2 f = __ldg(reinterpret_cast<const float2*>(p));
            v[0] = f.x; v[1] = f.y;
        } else {
#pragma unroll
            for (int i = 0; i < C / 4; i++) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(p) + i);
                v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
            }
        }
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[C]) {
#pragma unroll
        for (int c = 0; c < C; c++) p[c] = v[c];
    }
};

__device__ __forceinline__ float smoothstep_f(float v) { return v * v * (3.0f - 2.0f * v); }
__device__ __forceinline__ float smoothstep_d(float v) { return 6 * v * (1.0f - v); }

// Per-point, per-level position: returns false when the point is outside [0,1]^D.
template <uint32_t D>
__device__ __forceinline__ void locate(const float (&x)[D], const LevelParams& lp, const bool align_corners,
                                       const uint32_t interp, float (&pos)[D], float (&deriv)[D], uint32_t (&pos_grid)[D]) {
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        pos[d] = x[d] * lp.scale + (align_corners ? 0.0f : 0.5f);
        pos_grid[d] = floorf(pos[d]);
        pos[d] -= (float)pos_grid[d];
        if (interp == 1) {
            deriv[d] = smoothstep_d(pos[d]);
            pos[d] = smoothstep_f(pos[d]);
        } else {
            deriv[d] = 1.0f;
        }
    }
}

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
    if (threadIdx.x < L) s_lp[threadIdx.x] = make_level(offsets, threadIdx.x, S, H);
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
                        const uint32_t row = grid_row<D>(gridtype, align_corners, lp.hashmap_size, lp.resolution, pg);
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

// grid = (point chunks, L).  TG = dtype of grad_table.
template <typename T, typename TG, uint32_t D, uint32_t C, bool PRIV>
__global__ void __launch_bounds__(256) k_grid_backward(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                       const int* __restrict__ offsets, TG* __restrict__ grad_table,
                                                       const uint32_t B, const uint32_t L, const float S, const uint32_t H,
                                                       const uint32_t gridtype, const bool align_corners,
                                                       const uint32_t interp, const uint32_t smem_rows_max,
                                                       const uint32_t points_per_cta, const int* __restrict__ b_dev) {
    extern __shared__ float s_acc[];  // [hashmap_size * C] when the level is privatised
    const uint32_t level = blockIdx.y;
    const LevelParams lp = make_level(offsets, level, S, H);
    const bool privatised = lp.hashmap_size <= smem_rows_max;
    if (privatised != PRIV) return;  // the other launch handles this level
    constexpr uint32_t NC = 1u << D;
    const uint32_t lane = threadIdx.x & 31u;

    if (privatised) {
        for (uint32_t i = threadIdx.x; i < lp.hashmap_size * C; i += blockDim.x) s_acc[i] = 0.0f;
        __syncthreads();
    }
    TG* gl = grad_table + (size_t)lp.offset * C;

    const uint32_t Bn = b_dev ? min(B, (uint32_t)max(*b_dev, 0)) : B;  // live rows
    const uint32_t b_begin = blockIdx.x * points_per_cta;
    const uint32_t b_end = min(Bn, b_begin + points_per_cta);
    if (b_begin >= b_end) return;
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
        if (valid) {
            Row<T, C>::load(grad + ((size_t)b * L + level) * C, g);
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; c++) g[c] = 0.0f;
        }
        float pos[D], deriv[D];
        uint32_t pos_grid[D];
        locate<D>(x, lp, align_corners, interp, pos, deriv, pos_grid);
#pragma unroll
        for (uint32_t idx = 0; idx < NC; idx++) {
            float w = 1;
            uint32_t pg[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                w *= ((idx >> d) & 1u) ? pos[d] : 1 - pos[d];
                pg[d] = pos_grid[d] + ((idx >> d) & 1u);
            }
            const uint32_t row = grid_row<D>(gridtype, align_corners, lp.hashmap_size, lp.resolution, pg);
            float wv[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) wv[c] = w * g[c];
            // lanes of a warp that hit the same row (consecutive samples of a ray share coarse cells) are summed first
            const uint32_t key = valid ? row : (0xffffffffu - lane);
            const bool leader = warp_aggregate<C>(key, wv, lane);
            if (valid && leader) {
                if (privatised) {
#pragma unroll
                    for (uint32_t c = 0; c < C; c++) atomicAdd(&s_acc[row * C + c], wv[c]);
                } else {
                    VecAtomic<TG, C>::add(gl + (size_t)row * C, wv);
                }
            }
        }
    }
    if (privatised) {
        __syncthreads();
        for (uint32_t r = threadIdx.x; r < lp.hashmap_size; r += blockDim.x) {
            float v[C];
            bool nz = false;
#pragma unroll
            for (uint32_t c = 0; c < C; c++) { v[c] = s_acc[r * C + c]; nz |= (v[c] != 0.0f); }
            if (nz) VecAtomic<TG, C>::add(gl + (size_t)r * C, v);
        }
    }
}

// grad_x[b, d] = sum_{l,c} grad[b,l,c] * d out[b,l,c] / d x[b,d]; recomputed from the table (fp32 accumulate)
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) k_grid_input_backward_recompute(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                                       const T* __restrict__ table, const int* __restrict__ offsets,
                                                                       float* __restrict__ grad_x, const uint32_t B, const uint32_t L,
                                                                       const float S, const uint32_t H, const uint32_t gridtype,
                                                                       const bool align_corners, const uint32_t interp,
                                                                       const int* __restrict__ b_dev) {
    __shared__ LevelParams s_lp[kMaxLevels];
    if (threadIdx.x < L) s_lp[threadIdx.x] = make_level(offsets, threadIdx.x, S, H);
    __syncthreads();
    const uint32_t Bn = b_dev ? min(B, (uint32_t)max(*b_dev, 0)) : B;
    constexpr uint32_t NC = 1u << D;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < Bn; b += gridDim.x * blockDim.x) {
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
                const T* tl = table + (size_t)lp.offset * C;
                float val[NC][C];
#pragma unroll
                for (uint32_t idx = 0; idx < NC; idx++) {
                    uint32_t pg[D];
#pragma unroll
                    for (uint32_t d = 0; d < D; d++) pg[d] = pos_grid[d] + ((idx >> d) & 1u);
                    const uint32_t row = grid_row<D>(gridtype, align_corners, lp.hashmap_size, lp.resolution, pg);
                    Row<T, C>::load(tl + (size_t)row * C, val[idx]);
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
        const LevelParams lp = make_level(offsets, t, S, H);
        scales[t] = lp.scale;
        resolutions[t] = lp.resolution;
    }
    if (t >= B * L) return;
    const uint32_t b = t / L, l = t - b * L;
    const LevelParams lp = make_level(offsets, l, S, H);
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
        indices[(size_t)t * NC + idx] = grid_row<D>(gridtype, align_corners, lp.hashmap_size, lp.resolution, pg);
    }
}

// ---- host dispatch ------------------------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
int launch_forward(const float* x, const void* table, const int* offsets, void* out, void* dy_dx, uint32_t B, uint32_t L, float S,
                   uint32_t H, uint32_t gridtype, bool align, uint32_t interp, const int* b_dev, cudaStream_t st) {
    constexpr uint32_t G16 = 16 / (C * sizeof(T));  // levels per 16-byte output store
    constexpr uint32_t G = (D <= 3 && G16 >= 2 && G16 <= 4) ? G16 : 1;
    const uint32_t threads = 256;
    const uint32_t blocks = div_up(B, threads);
    if constexpr (G > 1) {
        const bool vec_ok = (L % G) == 0 && ((L * C * sizeof(T)) % 16 == 0) && ((uintptr_t)out % 16 == 0);
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

constexpr uint32_t kBwdSmemBytes = 160 * 1024;  // privatised accumulator budget per CTA

template <typename T, typename TG, uint32_t D, uint32_t C>
int launch_backward(const void* grad, const float* x, const void* table, const int* offsets, void* grad_table, const void* dy_dx,
                    float* grad_x, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align, uint32_t interp,
                    const int* b_dev, cudaStream_t st) {
    auto kern = k_grid_backward<T, TG, D, C, true>;
    auto kern_direct = k_grid_backward<T, TG, D, C, false>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    // Shared-memory privatisation of the small (coarse) levels only pays when many points hit each cell (large batches,
    // e.g. the 2^22-point encoder benchmark); for a training batch every level uses warp-aggregated global atomics.
    const bool use_priv = B >= (1u << 18);
    const uint32_t smem_rows_max = use_priv ? kBwdSmemBytes / (C * sizeof(float)) : 0u;
    int rc = 0;
    if (use_priv) {
        // chunks: enough CTAs per level to fill the machine ~2x over all levels, but large enough to amortise the flush
        uint32_t chunks = div_up(2u * SEALD_NUM_SMS, L);
        uint32_t ppc = div_up(B, chunks);
        ppc = div_up(ppc < 2048u ? 2048u : ppc, 256u) * 256u;
        chunks = div_up(B, ppc);
        dim3 grid(chunks, L);
        kern<<<grid, 256, kBwdSmemBytes, st>>>((const T*)grad, x, offsets, (TG*)grad_table, B, L, S, H, gridtype, align, interp, smem_rows_max, ppc, b_dev);
        rc = launch_status();
        if (rc) return rc;
    }
    // direct (warp-aggregated atomics) levels: small chunks, no shared memory, full occupancy
    const uint32_t ppc_d = 1024;
    dim3 grid_d(div_up(B, ppc_d), L);
    kern_direct<<<grid_d, 256, 0, st>>>((const T*)grad, x, offsets, (TG*)grad_table, B, L, S, H, gridtype, align, interp, smem_rows_max, ppc_d, b_dev);
    rc = launch_status();
    if (rc) return rc;
    if (grad_x) {
        if (dy_dx) {
            k_grid_input_backward_dydx<T, D, C><<<div_up(B * D, 256u), 256, 0, st>>>((const T*)grad, (const T*)dy_dx, grad_x, B, L, b_dev);
        } else {
            k_grid_input_backward_recompute<T, D, C><<<div_up(B, 256u), 256, 0, st>>>((const T*)grad, x, (const T*)table, offsets, grad_x, B, L, S, H, gridtype, align, interp, b_dev);
        }
        rc = launch_status();
    }
    return rc;
}

template <typename T, typename TG, uint32_t D>
int dispatch_backward_C(uint32_t C, const void* grad, const float* x, const void* table, const int* offsets, void* grad_table,
                        const void* dy_dx, float* grad_x, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,
                        uint32_t interp, const int* b_dev, cudaStream_t st) {
    switch (C) {
        case 1: return launch_backward<T, TG, D, 1>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 2: return launch_backward<T, TG, D, 2>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 4: return launch_backward<T, TG, D, 4>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 8: return launch_backward<T, TG, D, 8>(grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
        default: return SEALD_E_UNSUPPORTED;
    }
}

template <typename T, typename TG>
int dispatch_backward_D(uint32_t D, uint32_t C, const void* grad, const float* x, const void* table, const int* offsets, void* grad_table,
                        const void* dy_dx, float* grad_x, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align,
                        uint32_t interp, const int* b_dev, cudaStream_t st) {
    switch (D) {
        case 2: return dispatch_backward_C<T, TG, 2>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 3: return dispatch_backward_C<T, TG, 3>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 4: return dispatch_backward_C<T, TG, 4>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
        case 5: return dispatch_backward_C<T, TG, 5>(C, grad, x, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, align, interp, b_dev, st);
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

extern "C" int seald_grid_encode_backward(const void* grad_out, const float* x01, const void* table, const int32_t* offsets,
                                          void* grad_table, const void* dy_dx, float* grad_x, uint32_t B, uint32_t D, uint32_t C,
                                          uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                                          int dtype, int grad_table_dtype, const int32_t* b_dev, seald_stream_t stream) {
    if (B == 0) return 0;
    if (!grad_out || !x01 || !offsets || !grad_table) return SEALD_E_BADARG;
    if (grad_x && !dy_dx && !table) return SEALD_E_BADARG;
    if (L == 0 || L > kMaxLevels || gridtype > 1 || interp > 1) return SEALD_E_UNSUPPORTED;
    cudaStream_t st = to_stream(stream);
    const bool al = align_corners != 0;
    if (dtype == SEALD_F16 && grad_table_dtype == SEALD_F16)
        return dispatch_backward_D<__half, __half>(D, C, grad_out, x01, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, al, interp, b_dev, st);
    if (dtype == SEALD_F16 && grad_table_dtype == SEALD_F32)
        return dispatch_backward_D<__half, float>(D, C, grad_out, x01, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, al, interp, b_dev, st);
    if (dtype == SEALD_F32 && grad_table_dtype == SEALD_F32)
        return dispatch_backward_D<float, float>(D, C, grad_out, x01, table, offsets, grad_table, dy_dx, grad_x, B, L, S, H, gridtype, al, interp, b_dev, st);
    return SEALD_E_UNSUPPORTED;
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
