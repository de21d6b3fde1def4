// grid_device.cuh — device-side building blocks of the multiresolution hash grid (reference: gridencoder/src/gridencoder.cu:50-84
// get_grid_index, :100-216 forward): level parameters, row addressing with per-level fast paths, per-point location inside a level,
// row loads / stores and the paired corner gather.  Shared by the encoder kernels (grid_encoder.cu) and the fused encoder + heads
// kernel (field.cu).
#pragma once
#include "common.cuh"

namespace seald {

struct LevelParams {
    float scale;
    uint32_t resolution;
    uint32_t hashmap_size;
    uint32_t offset;  // in rows
    // how grid_row addresses this level (functions of the level only, hoisted out of the per-corner arithmetic):
    uint32_t stride1;   // resolution (+1 without align_corners): row stride of dimension 1 of the dense index
    uint32_t n_dense;   // dimensions the dense index consumes before its stride passes hashmap_size (gridencoder.cu:57-60)
    uint32_t use_hash;  // hashed level: gridtype 0 and the dense index does not fit
    uint32_t mask;      // hashmap_size - 1 when hashmap_size is a power of two (x % size == x & mask), else 0
};

constexpr int kMaxLevels = 32;

// gridencoder.cu:137-139 — evaluated on device in fp32 with the same expression shape.
__device__ __forceinline__ LevelParams make_level(const int* __restrict__ offsets, uint32_t level, float S, uint32_t H, const uint32_t D,
                                                  const uint32_t gridtype, const bool align_corners) {
    LevelParams p;
    p.offset = (uint32_t)offsets[level];
    p.hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    p.scale = exp2f(level * S) * H - 1.0f;
    p.resolution = (uint32_t)ceil(p.scale) + 1;
    p.stride1 = align_corners ? p.resolution : p.resolution + 1;
    uint32_t stride = 1, nd = 0;
    for (uint32_t d = 0; d < D && stride <= p.hashmap_size; d++) {  // the loop of gridencoder.cu:57-60 on the strides alone
        stride *= p.stride1;
        nd++;
    }
    p.n_dense = nd;
    p.use_hash = (gridtype == 0 && stride > p.hashmap_size) ? 1u : 0u;
    p.mask = (p.hashmap_size & (p.hashmap_size - 1)) == 0 ? p.hashmap_size - 1 : 0u;
    return p;
}

// gridencoder.cu:50-84 (get_grid_index): same index, with everything that depends on the level alone taken from LevelParams — the
// reference's `index % hashmap_size` (an integer division per corner: half of this file's instructions when written naively)
// becomes a mask on the power-of-two hashed levels and a never-taken compare on the dense ones.
template <uint32_t D>
__device__ __forceinline__ uint32_t grid_row(const LevelParams& lp, const uint32_t pos_grid[D]) {
    uint32_t index;
    if (lp.use_hash) {
        constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        index = 0;
#pragma unroll
        for (uint32_t i = 0; i < D; ++i) index ^= pos_grid[i] * primes[i];
        if (lp.mask) return index & lp.mask;
    } else {
        uint32_t stride = 1;
        index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if (d < lp.n_dense) index += pos_grid[d] * stride;
            stride *= lp.stride1;
        }
        if (index < lp.hashmap_size) return index;
    }
    return index % lp.hashmap_size;
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
// forward, paired corner loads
// =================================================================================================
// The two corners of a cell that differ in dimension 0 are neighbours in memory: on a dense level they are rows r and
// r + 1, on a hashed level (prime[0] == 1) rows h ^ x and h ^ (x + 1), which differ only in the low bits x ^ (x + 1).
// Three times out of four both lie inside one aligned 16-byte group of the table, so ONE 16-byte load fetches the pair and
// a 4/8-byte load is issued only by the lanes whose pair straddles a group.  The gathers are L1-wavefront bound (one
// wavefront per lane per load), so this removes ~3/8 of the wavefronts.  Rows are RB = C * sizeof(T) = 4 or 8 bytes.
template <typename T, uint32_t C>
struct RowPack {
    static constexpr uint32_t RB = C * sizeof(T);
    static constexpr uint32_t W = RB / 4;        // 32-bit words per row
    static constexpr uint32_t R16 = 16 / RB;     // rows per 16-byte group
    static constexpr bool ok = (RB == 4 || RB == 8);
    static __device__ __forceinline__ void pick(const uint4& u, const uint32_t k, uint32_t (&w)[W]) {
        if constexpr (W == 1) {
            const uint32_t lo = (k & 1u) ? u.y : u.x;
            const uint32_t hi = (k & 1u) ? u.w : u.z;
            w[0] = (k & 2u) ? hi : lo;
        } else {
            w[0] = (k & 1u) ? u.z : u.x;
            w[1] = (k & 1u) ? u.w : u.y;
        }
    }
    static __device__ __forceinline__ void load_row(const T* p, uint32_t (&w)[W]) {
        if constexpr (W == 1) {
            w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
        } else {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
            w[0] = u.x; w[1] = u.y;
        }
    }
    static __device__ __forceinline__ void unpack(const uint32_t (&w)[W], float (&v)[C]) {
        if constexpr (sizeof(T) == 2) {
#pragma unroll
            for (uint32_t i = 0; i < W; i++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                v[2 * i] = f.x; v[2 * i + 1] = f.y;
            }
        } else {
#pragma unroll
            for (uint32_t i = 0; i < W; i++) v[i] = __uint_as_float(w[i]);
        }
    }
};

// Rows (relative to the level) of the 2^(D-1) corner pairs of a cell: pair j = corners (x, y + j0, z + j1, ..) and (x + 1, ..).
// Two straight-line fast paths chosen per LEVEL (uniform over the warp): hashed levels with a power-of-two table (row = (x ^ K) &
// mask, K = xor of the other dimensions' products, shared by the pair) and dense levels whose largest corner index stays inside
// the level (no wrap: row = base + strides, pair = r, r + 1).  Anything else (tiled grids, non-power-of-two hashed levels,
// align_corners wrap at x = 1) takes the reference's index arithmetic per corner (grid_row).
template <uint32_t D>
__device__ __forceinline__ void pair_rows(const LevelParams& lp, const uint32_t (&pos_grid)[D], uint32_t (&r0)[1u << (D - 1)],
                                          uint32_t (&r1)[1u << (D - 1)]) {
    constexpr uint32_t NP = 1u << (D - 1);
    bool fast = false;
    if (lp.use_hash) {
        if (lp.mask) {
            constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
            uint32_t lo[D], hi[D];
#pragma unroll
            for (uint32_t d = 1; d < D; d++) {
                lo[d] = pos_grid[d] * primes[d];
                hi[d] = lo[d] + primes[d];
            }
#pragma unroll
            for (uint32_t j = 0; j < NP; j++) {
                uint32_t k = 0;
#pragma unroll
                for (uint32_t d = 1; d < D; d++) k ^= ((j >> (d - 1)) & 1u) ? hi[d] : lo[d];
                r0[j] = (pos_grid[0] ^ k) & lp.mask;
                r1[j] = ((pos_grid[0] + 1) ^ k) & lp.mask;
            }
            fast = true;
        }
    } else if (lp.n_dense == D) {
        uint32_t st[D];
        st[0] = 1;
#pragma unroll
        for (uint32_t d = 1; d < D; d++) st[d] = st[d - 1] * lp.stride1;
        uint32_t base = 0, top = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            base += pos_grid[d] * st[d];
            top += st[d];
        }
        if (base + top < lp.hashmap_size) {
#pragma unroll
            for (uint32_t j = 0; j < NP; j++) {
                uint32_t r = base;
#pragma unroll
                for (uint32_t d = 1; d < D; d++) r += ((j >> (d - 1)) & 1u) ? st[d] : 0u;
                r0[j] = r;
                r1[j] = r + 1;
            }
            fast = true;
        }
    }
    if (!fast) {
#pragma unroll
        for (uint32_t j = 0; j < NP; j++) {
            uint32_t pg[D];
            pg[0] = pos_grid[0];
#pragma unroll
            for (uint32_t d = 1; d < D; d++) pg[d] = pos_grid[d] + ((j >> (d - 1)) & 1u);
            r0[j] = grid_row<D>(lp, pg);
            pg[0] = pos_grid[0] + 1;
            r1[j] = grid_row<D>(lp, pg);
        }
    }
}

// Gathers the 2^D corner rows of one cell with paired loads.  issue(): all loads in flight; resolve(): rows as floats.
template <typename T, uint32_t D, uint32_t C>
struct CellGather {
    using RP = RowPack<T, C>;
    static constexpr uint32_t NP = 1u << (D - 1);
    uint4 u[NP];
    uint32_t w1[NP][RP::W];
    uint32_t code;      // per pair 4 bits: position of row 0 / row 1 inside the 16-byte group (2 + 2); NP <= 8
    uint32_t straddle;  // bit j: pair j does not fit one group, w1[j] holds row 1

    // tl = base of the whole table (16-byte aligned, readable up to the next 16-byte boundary past its last row)
    __device__ __forceinline__ void issue(const T* __restrict__ tl, const uint32_t gridtype, const bool align_corners, const LevelParams& lp,
                                          const uint32_t (&pos_grid)[D]) {
        code = 0;
        straddle = 0;
        // absolute rows: 16-byte groups are aligned relative to the table base (level offsets need not be)
        uint32_t r0[NP], r1[NP];
        pair_rows<D>(lp, pos_grid, r0, r1);
#pragma unroll
        for (uint32_t j = 0; j < NP; j++) { r0[j] += lp.offset; r1[j] += lp.offset; }
#pragma unroll
        for (uint32_t j = 0; j < NP; j++) {
            u[j] = __ldg(reinterpret_cast<const uint4*>(tl + (size_t)(r0[j] & ~(RP::R16 - 1)) * C));
            const bool far = (r0[j] ^ r1[j]) >= RP::R16;
            if (far) RP::load_row(tl + (size_t)r1[j] * C, w1[j]);
            code |= ((r0[j] & (RP::R16 - 1)) | ((r1[j] & (RP::R16 - 1)) << 2)) << (4 * j);
            straddle |= (far ? 1u : 0u) << j;
        }
    }
    // val[idx][c], idx bit d = +1 in dimension d (same corner numbering as the reference)
    __device__ __forceinline__ void resolve(float (&val)[1u << D][C]) const {
#pragma unroll
        for (uint32_t j = 0; j < NP; j++) {
            uint32_t a[RP::W], b[RP::W];
            RP::pick(u[j], (code >> (4 * j)) & 3u, a);
            RP::pick(u[j], (code >> (4 * j + 2)) & 3u, b);
            if ((straddle >> j) & 1u) {
#pragma unroll
                for (uint32_t i = 0; i < RP::W; i++) b[i] = w1[j][i];
            }
            RP::unpack(a, val[2 * j]);
            RP::unpack(b, val[2 * j + 1]);
        }
    }
};

}  // namespace seald
