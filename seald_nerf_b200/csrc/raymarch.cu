// raymarch.cu — occupancy-bitfield ray marching and alpha compositing for sm_100a.
//
// Integer/index results (Morton codes, bitfield bytes, per-ray sample counts) are bit-exact with the
// reference's raymarching extension, and so are the fp32 sample positions: every floating-point expression
// below keeps the operand order and literal types of the reference (raymarching/src/raymarching.cu:42-81
// helpers, :108-144 slab test, :335-479 training march, :730-804 inference march, :516-576 / :620-681 /
// :833-904 compositing) so that nvcc forms the same FMA contractions; the file must NOT be built with
// --use_fast_math.  What differs is the parallel structure:
//   * march_rays_train fuses the AABB slab test, and reserves output ranges with ONE atomic per warp after a
//     warp prefix sum (ray-ordered packing inside a warp) instead of two atomics per ray;
//   * rays[] rows are written in ray order (row n describes ray n), which makes the layout deterministic;
//   * the inference kernels can take the live-ray count from device memory so the render loop needs no host
//     synchronisation, and seald_compact_alive replaces the boolean-mask compaction.
#include <cstdlib>
#include <limits>

#include "common.cuh"
#include "seal.cuh"

namespace seald {

__device__ __forceinline__ constexpr float SQRT3() { return 1.7320508075688772f; }
__device__ __forceinline__ constexpr float RPI() { return 0.3183098861837907f; }

__device__ __forceinline__ float signf(const float x) { return copysignf(1.0, x); }
__device__ __forceinline__ void swapf(float& a, float& b) { float c = a; a = b; b = c; }

// raymarching.cu:42-54
__device__ __forceinline__ int mip_from_pos(const float x, const float y, const float z, const float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabs(y), fabs(z)));
    int exponent;
    frexpf(mx, &exponent);
    return fminf(max_cascade - 1, fmaxf(0, exponent));
}
__device__ __forceinline__ int mip_from_dt(const float dt, const float H, const float max_cascade) {
    const float mx = dt * H * 0.5;
    int exponent;
    frexpf(mx, &exponent);
    return fminf(max_cascade - 1, fmaxf(0, exponent));
}

// raymarching.cu:56-81
__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D_enc(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t morton3D_dec(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

// raymarching.cu:108-144
__device__ __forceinline__ void slab_test(const float ox, const float oy, const float oz, const float dx, const float dy,
                                          const float dz, const float* __restrict__ aabb, const float min_near, float& near_o,
                                          float& far_o) {
    const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
    float near = (aabb[0] - ox) * rdx;
    float far = (aabb[3] - ox) * rdx;
    if (near > far) swapf(near, far);
    float near_y = (aabb[1] - oy) * rdy;
    float far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) swapf(near_y, far_y);
    if (near > far_y || near_y > far) {
        near_o = far_o = std::numeric_limits<float>::max();
        return;
    }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - oz) * rdz;
    float far_z = (aabb[5] - oz) * rdz;
    if (near_z > far_z) swapf(near_z, far_z);
    if (near > far_z || near_z > far) {
        near_o = far_o = std::numeric_limits<float>::max();
        return;
    }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_o = near;
    far_o = far;
}

__global__ void k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ aabb,
                           const uint32_t N, const float min_near, float* nears, float* fars) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float* o = rays_o + (size_t)n * 3;
    const float* d = rays_d + (size_t)n * 3;
    float near, far;
    slab_test(o[0], o[1], o[2], d[0], d[1], d[2], aabb, min_near, near, far);
    nears[n] = near;
    fars[n] = far;
}

// raymarching.cu:163-198
__global__ void k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float radius, const uint32_t N,
                               float* coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    rays_o += (size_t)n * 3;
    rays_d += (size_t)n * 3;
    coords += (size_t)n * 2;
    const float ox = rays_o[0], oy = rays_o[1], oz = rays_o[2];
    const float dx = rays_d[0], dy = rays_d[1], dz = rays_d[2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float C = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * C)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2(sqrtf(x * x + z * z), y);
    const float phi = atan2(z, x);
    coords[0] = 2 * theta * RPI() - 1;
    coords[1] = phi * RPI();
}

__global__ void k_morton3D(const int* __restrict__ coords, const uint32_t N, int* indices) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    coords += (size_t)n * 3;
    indices[n] = morton3D_enc(coords[0], coords[1], coords[2]);
}

__global__ void k_morton3D_invert(const int* __restrict__ indices, const uint32_t N, int* coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    coords += (size_t)n * 3;
    const int ind = indices[n];
    coords[0] = morton3D_dec(ind >> 0);
    coords[1] = morton3D_dec(ind >> 1);
    coords[2] = morton3D_dec(ind >> 2);
}

// One thread packs 32 cells -> 4 bytes (one 128-byte coalesced read per 4 lanes), raymarching.cu:281-288.
__global__ void k_packbits(const float* __restrict__ grid, const uint32_t n_bytes, const float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t w = threadIdx.x + blockIdx.x * blockDim.x;  // 32-bit word of the bitfield
    const uint32_t n_words = n_bytes / 4;
    if (w < n_words) {
        const float4* g = reinterpret_cast<const float4*>(grid) + (size_t)w * 8;
        uint32_t bits = 0;
#pragma unroll
        for (uint32_t i = 0; i < 8; i++) {
            const float4 v = __ldg(g + i);
            bits |= (v.x > thresh ? 1u : 0u) << (4 * i);
            bits |= (v.y > thresh ? 1u : 0u) << (4 * i + 1);
            bits |= (v.z > thresh ? 1u : 0u) << (4 * i + 2);
            bits |= (v.w > thresh ? 1u : 0u) << (4 * i + 3);
        }
        reinterpret_cast<uint32_t*>(bitfield)[w] = bits;
    } else {
        // tail bytes (n_bytes % 4)
        const uint32_t n = n_words * 4 + (w - n_words);
        if (n >= n_bytes) return;
        uint8_t bits = 0;
        for (uint32_t i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? ((uint8_t)1 << i) : 0;
        bitfield[n] = bits;
    }
}

__global__ void k_packbits_bytes(const float* __restrict__ grid, const uint32_t N, const float thresh, uint8_t* bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    grid += (size_t)n * 8;
    uint8_t bits = 0;
#pragma unroll
    for (uint8_t i = 0; i < 8; i++) bits |= (grid[i] > thresh) ? ((uint8_t)1 << i) : 0;
    bitfield[n] = bits;
}

// ------------------------------------------------------------------------------------------------
// The per-ray state of the march (raymarching.cu:335-351)
// ------------------------------------------------------------------------------------------------
struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};

struct MarchConst {
    float bound, dt_gamma, dt_min, dt_max, rH, H3;
    uint32_t C, H;
};

__device__ __forceinline__ MarchConst make_march_const(const float bound, const float dt_gamma, const uint32_t max_steps, const uint32_t C,
                                                        const uint32_t H) {
    MarchConst mc;
    mc.bound = bound;
    mc.dt_gamma = dt_gamma;
    mc.rH = 1 / (float)H;
    mc.H3 = H * H * H;
    mc.dt_min = 2 * SQRT3() / max_steps;
    mc.dt_max = 2 * SQRT3() * (1 << (C - 1)) / H;
    mc.C = C;
    mc.H = H;
    return mc;
}

// One probe of the occupancy grid at parameter t (raymarching.cu:360-379).  Returns occupancy; outputs the
// clamped position, dt, and what the empty-space skip needs.
struct Probe {
    float x, y, z, dt, mip_bound;
    int nx, ny, nz;
};

// the sample position at parameter t (raymarching.cu:360-362); ONE definition so that every kernel rounds it the same way
__device__ __forceinline__ void ray_point(const float ox, const float oy, const float oz, const float dx, const float dy, const float dz,
                                          const float bound, const float t, float& x, float& y, float& z) {
    x = clampf(ox + t * dx, -bound, bound);
    y = clampf(oy + t * dy, -bound, bound);
    z = clampf(oz + t * dz, -bound, bound);
}

__device__ __forceinline__ bool probe_grid(const Ray& r, const MarchConst& mc, const uint8_t* __restrict__ grid, const float t, Probe& p) {
    float x, y, z;
    ray_point(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, mc.bound, t, x, y, z);

    const float dt = clampf(t * mc.dt_gamma, mc.dt_min, mc.dt_max);

    const int level = max(mip_from_pos(x, y, z, mc.C), mip_from_dt(dt, mc.H, mc.C));

    const float mip_bound = fminf(scalbnf(1.0f, level), mc.bound);
    const float mip_rbound = 1 / mip_bound;

    const uint32_t H = mc.H;
    const int nx = clampf(0.5 * (x * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
    const int ny = clampf(0.5 * (y * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
    const int nz = clampf(0.5 * (z * mip_rbound + 1) * H, 0.0f, (float)(H - 1));

    const uint32_t index = level * mc.H3 + morton3D_enc(nx, ny, nz);
    const bool occ = grid[index / 8] & (1 << (index % 8));
    p.x = x; p.y = y; p.z = z; p.dt = dt; p.mip_bound = mip_bound;
    p.nx = nx; p.ny = ny; p.nz = nz;
    return occ;
}

// Parameter at which the ray leaves the current voxel (raymarching.cu:389-394).
__device__ __forceinline__ float voxel_exit(const Ray& r, const MarchConst& mc, const Probe& p, const float t) {
    const float rH = mc.rH;
    const float tx = (((p.nx + 0.5f + 0.5f * signf(r.dx)) * rH * 2 - 1) * p.mip_bound - p.x) * r.rdx;
    const float ty = (((p.ny + 0.5f + 0.5f * signf(r.dy)) * rH * 2 - 1) * p.mip_bound - p.y) * r.rdy;
    const float tz = (((p.nz + 0.5f + 0.5f * signf(r.dz)) * rH * 2 - 1) * p.mip_bound - p.z) * r.rdz;
    return t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
}

// Skip to the next voxel (raymarching.cu:388-399).
__device__ __forceinline__ float skip_voxel(const Ray& r, const MarchConst& mc, const Probe& p, float t) {
    const float tt = voxel_exit(r, mc, p, t);
    do {
        t += clampf(t * mc.dt_gamma, mc.dt_min, mc.dt_max);
    } while (t < tt);
    return t;
}

// ------------------------------------------------------------------------------------------------
// Occupied-region guard.  `occ` = world-space AABB of every occupied cell of the bitfield, grown by a guard band of whole
// cells (seald_occupancy_aabb).  A sample can only be produced where the ray is inside an occupied cell, so
//   * a ray whose [near, far] segment misses `occ` has no samples at all (exactly what the full walk would find), and
//   * no sample lies beyond the parameter at which the ray leaves `occ`: the walk may stop there.
// Both only remove probes that land in empty cells; the sequence of probed chain elements up to the last sample - and with
// it every count, position and delta - is untouched.  Returns false for a miss; otherwise lowers `far` to the exit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool clip_to_occupied(const Ray& r, const float* __restrict__ occ, const float near, float& far) {
    float t0 = near, t1 = far;
    const float o[3] = {r.ox, r.oy, r.oz}, rd[3] = {r.rdx, r.rdy, r.rdz}, d[3] = {r.dx, r.dy, r.dz};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float lo = __ldg(occ + a), hi = __ldg(occ + 3 + a);
        if (d[a] == 0.0f) {
            if (o[a] < lo || o[a] > hi) return false;
        } else {
            float ta = (lo - o[a]) * rd[a], tb = (hi - o[a]) * rd[a];
            if (ta > tb) swapf(ta, tb);
            t0 = fmaxf(t0, ta);
            t1 = fminf(t1, tb);
        }
    }
    if (!(t0 <= t1)) return false;
    far = fminf(far, t1);
    return true;
}

// per-cascade integer cell ranges of the occupied cells: ranges[c][0..2] = min ix,iy,iz ; [3..5] = max (Morton order grid)
__global__ void k_occupancy_ranges(const uint8_t* __restrict__ bitfield, const uint32_t C, const uint32_t H, int* __restrict__ ranges) {
    const uint32_t n_bytes = C * H * H * H / 8;
    const uint32_t H3 = H * H * H;
    int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
    uint32_t cas = 0xffffffffu;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += gridDim.x * blockDim.x) {
        const uint32_t bits = bitfield[i];
        if (!bits) continue;
        const uint32_t c = (i * 8) / H3;
        if (c != cas) {
            if (cas != 0xffffffffu) {
#pragma unroll
                for (int a = 0; a < 3; a++) { atomicMin(ranges + cas * 6 + a, lo[a]); atomicMax(ranges + cas * 6 + 3 + a, hi[a]); }
            }
            cas = c;
#pragma unroll
            for (int a = 0; a < 3; a++) { lo[a] = 1 << 30; hi[a] = -1; }
        }
        for (uint32_t b = 0; b < 8; b++) {
            if (!(bits & (1u << b))) continue;
            const uint32_t m = (i * 8 + b) - c * H3;
            const int ix = morton3D_dec(m), iy = morton3D_dec(m >> 1), iz = morton3D_dec(m >> 2);
            lo[0] = min(lo[0], ix); lo[1] = min(lo[1], iy); lo[2] = min(lo[2], iz);
            hi[0] = max(hi[0], ix); hi[1] = max(hi[1], iy); hi[2] = max(hi[2], iz);
        }
    }
    // one set of atomics per WARP and cascade (a well-trained grid has hundreds of thousands of occupied bytes: per-thread atomics on
    // the same six words serialise — 0.6 ms per frame measured, 38 ms per refresh)
    for (uint32_t c = 0; c < C; c++) {
        const bool mine = (cas == c);
        if (!__any_sync(0xffffffffu, mine)) continue;
        int l[3], h[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            l[a] = __reduce_min_sync(0xffffffffu, mine ? lo[a] : (1 << 30));
            h[a] = __reduce_max_sync(0xffffffffu, mine ? hi[a] : -1);
        }
        if ((threadIdx.x & 31u) == 0) {
#pragma unroll
            for (int a = 0; a < 3; a++) { atomicMin(ranges + c * 6 + a, l[a]); atomicMax(ranges + c * 6 + 3 + a, h[a]); }
        }
    }
}

__global__ void k_occupancy_init(int* __restrict__ ranges, const uint32_t C) {
    const uint32_t i = threadIdx.x;
    if (i < C * 6) ranges[i] = (i % 6 < 3) ? (1 << 30) : -1;
}

// union over the cascades of the occupied cell boxes, grown by `guard` cells of that cascade, in world coordinates
// (cell ix of cascade c spans [(ix / H * 2 - 1) * mip_bound, ((ix + 1) / H * 2 - 1) * mip_bound], raymarching.cu:372-376)
__global__ void k_occupancy_finalize(const int* __restrict__ ranges, const uint32_t C, const uint32_t H, const float bound, const int guard,
                                     float* __restrict__ aabb6) {
    if (threadIdx.x != 0) return;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (uint32_t c = 0; c < C; c++) {
        const float mip_bound = fminf(scalbnf(1.0f, (int)c), bound);
        for (int a = 0; a < 3; a++) {
            const int l = ranges[c * 6 + a], h = ranges[c * 6 + 3 + a];
            if (h < l) continue;
            lo[a] = fminf(lo[a], ((float)(l - guard) / (float)H * 2.0f - 1.0f) * mip_bound);
            hi[a] = fmaxf(hi[a], ((float)(h + 1 + guard) / (float)H * 2.0f - 1.0f) * mip_bound);
        }
    }
    for (int a = 0; a < 3; a++) { aabb6[a] = lo[a]; aabb6[3 + a] = hi[a]; }
}

// ------------------------------------------------------------------------------------------------
// training march
// ------------------------------------------------------------------------------------------------
template <bool SEAL>
__global__ void k_march_rays_train(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
                                   const float bound, const float dt_gamma, const uint32_t max_steps, const uint32_t N, const uint32_t C,
                                   const uint32_t H, const uint32_t M, const float* __restrict__ nears, const float* __restrict__ fars,
                                   const float* __restrict__ aabb, const float min_near, float* __restrict__ nears_out,
                                   float* __restrict__ fars_out, float* __restrict__ xyzs, float* __restrict__ dirs,
                                   float* __restrict__ deltas, int* __restrict__ rays, int* __restrict__ counter,
                                   const float* __restrict__ noises, const __grid_constant__ seald_seal_mapper mp,
                                   uint8_t* __restrict__ seal_mask, const float* __restrict__ occ) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    const bool active = n < N;

    Ray r;
    float near = 0.f, far = 0.f, noise = 0.f;
    if (active) {
        const float* o = rays_o + (size_t)n * 3;
        const float* d = rays_d + (size_t)n * 3;
        r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
        r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
        r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
        if (nears) {
            near = nears[n];
            far = fars[n];
        } else {
            slab_test(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, aabb, min_near, near, far);
            if (nears_out) { nears_out[n] = near; fars_out[n] = far; }
        }
        noise = noises[n];
    }
    const MarchConst mc = make_march_const(bound, dt_gamma, max_steps, C, H);

    float t0 = near;
    t0 += clampf(t0 * dt_gamma, mc.dt_min, mc.dt_max) * noise;
    bool may_hit = active;
    if (occ && active) may_hit = clip_to_occupied(r, occ, near, far);

    // pass 1: count
    float t = t0;
    uint32_t num_steps = 0;
    if (may_hit) {
        Probe p;
        while (t < far && num_steps < max_steps) {
            if (probe_grid(r, mc, grid, t, p)) {
                num_steps++;
                t += p.dt;
            } else {
                t = skip_voxel(r, mc, p, t);
            }
        }
    }

    // reserve the output range: warp prefix sum + one atomic per warp (raymarching.cu:405-406 does two per ray)
    const uint32_t incl = warp_inclusive_scan(num_steps, lane);
    const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t warp_rays = __popc(__ballot_sync(0xffffffffu, active));
    uint32_t base = 0;
    if (lane == 0) {
        base = atomicAdd(counter, warp_total);
        atomicAdd(counter + 1, warp_rays);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!active) return;
    const uint32_t point_index = base + incl - num_steps;

    rays[n * 3] = n;
    rays[n * 3 + 1] = point_index;
    rays[n * 3 + 2] = num_steps;

    if (num_steps == 0) return;
    if (point_index + num_steps > M) {  // overflow: dropped ray; its rows below M get the reference's zero fill (see the warp kernel)
        for (size_t s = point_index; s < M; s++) {
            xyzs[s * 3] = 0.f; xyzs[s * 3 + 1] = 0.f; xyzs[s * 3 + 2] = 0.f;
            dirs[s * 3] = 0.f; dirs[s * 3 + 1] = 0.f; dirs[s * 3 + 2] = 0.f;
            deltas[s * 2] = 0.f; deltas[s * 2 + 1] = 0.f;
            if (SEAL) seal_mask[s] = 0;
        }
        return;
    }

    xyzs += (size_t)point_index * 3;
    dirs += (size_t)point_index * 3;
    deltas += (size_t)point_index * 2;
    if (SEAL) seal_mask += point_index;

    // pass 2: write
    t = t0;
    uint32_t step = 0;
    float last_t = t;
    Probe p;
    while (t < far && step < num_steps) {
        if (probe_grid(r, mc, grid, t, p)) {
            if (SEAL) {
                // Seal proxy mapping fused in: the edited sample is remapped before it is stored (and encoded)
                float sx = p.x, sy = p.y, sz = p.z, sdx = r.dx, sdy = r.dy, sdz = r.dz;
                seal_mask[step] = seal_map_sample(mp, sx, sy, sz, sdx, sdy, sdz) ? 1 : 0;
                xyzs[0] = sx; xyzs[1] = sy; xyzs[2] = sz;
                dirs[0] = sdx; dirs[1] = sdy; dirs[2] = sdz;
            } else {
                xyzs[0] = p.x; xyzs[1] = p.y; xyzs[2] = p.z;
                dirs[0] = r.dx; dirs[1] = r.dy; dirs[2] = r.dz;
            }
            t += p.dt;
            deltas[0] = p.dt;
            deltas[1] = t - last_t;
            last_t = t;
            xyzs += 3; dirs += 3; deltas += 2;
            step++;
        } else {
            t = skip_voxel(r, mc, p, t);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// training march, one WARP per ray.
//
// Every parameter value the reference loop can visit is an element of ONE chain t_{k+1} = t_k + clamp(t_k * dt_gamma,
// dt_min, dt_max): both the sample step (raymarching.cu:386) and the empty-space skip (:396-398) advance t with that
// recurrence.  So a warp (a) generates 32 chain elements (the fp32 adds stay sequential, hence bit-exact), (b) probes
// the occupancy grid at all 32 in parallel, (c) replays the reference's control flow over ballots: runs of occupied
// elements become samples, an empty element jumps to the first element >= its voxel-exit time.  Sample parameters are
// kept in shared memory, the CTA reserves its output range with one atomic, and the samples are written cooperatively
// (coalesced) - no second march.  Same results as the thread-per-ray kernel, ~20x less latency at 4096 rays.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kMarchWarps = 8;
constexpr uint32_t kMaxStepsSmem = 1024;

// The sequential-chain walk of one ray by one warp (see k_march_rays_train_warp): returns the sample count, sample parameters in s_t.
__device__ __forceinline__ uint32_t walk_sequential(const Ray& r, const MarchConst& mc, const uint8_t* __restrict__ grid, const float t0,
                                                    const float far, const float dt_gamma, const uint32_t max_steps, const uint32_t lane,
                                                    float* __restrict__ s_t) {
    constexpr uint32_t FULL = 0xffffffffu;
    uint32_t count = 0;
    {
        float tb = t0;         // chain value at the start of the current block of 32 (warp uniform)
        uint32_t lp = 0;       // walker position inside the block
        bool has_pend = false; // an empty-space skip is looking for its landing element
        float tt_pend = 0.f;
        bool done = false;
        const bool const_dt = (dt_gamma == 0.0f);
        const float dt_c = clampf(0.0f, mc.dt_min, mc.dt_max);
        while (!done) {
            // (a) 32 chain elements, sequential adds
            float mine = tb, t = tb;
            if (const_dt) {
#pragma unroll
                for (uint32_t i = 0; i < 32; i++) {
                    if (i == lane) mine = t;
                    t += dt_c;
                }
            } else {
#pragma unroll 8
                for (uint32_t i = 0; i < 32; i++) {
                    if (i == lane) mine = t;
                    t += clampf(t * dt_gamma, mc.dt_min, mc.dt_max);
                }
            }
            tb = t;
            // (b) probe all of them
            Probe p;
            const bool occ = probe_grid(r, mc, grid, mine, p);
            const float tt = occ ? 0.0f : voxel_exit(r, mc, p, mine);
            const uint32_t valid_mask = __ballot_sync(FULL, mine < far);
            const uint32_t occ_mask = __ballot_sync(FULL, occ) & valid_mask;
            // (c) replay the control flow
            while (true) {
                if (lp >= 32) break;
                if (has_pend) {
                    const uint32_t ge = __ballot_sync(FULL, !(mine < tt_pend)) & (FULL << lp);  // `while (t < tt)` ends here
                    if (ge == 0) break;  // lands in a later block
                    lp = __ffs(ge) - 1;
                    has_pend = false;
                }
                if (!((valid_mask >> lp) & 1u)) { done = true; break; }  // t >= far
                if ((occ_mask >> lp) & 1u) {
                    const uint32_t inv = ~(occ_mask >> lp);
                    uint32_t n_run = inv ? (uint32_t)(__ffs(inv) - 1) : 32u;
                    n_run = min(n_run, max_steps - count);
                    if (lane >= lp && lane < lp + n_run) s_t[count + lane - lp] = mine;
                    count += n_run;
                    lp += n_run;
                    if (count >= max_steps) { done = true; break; }
                } else {
                    tt_pend = __shfl_sync(FULL, tt, lp);
                    has_pend = true;
                    lp += 1;  // do { t += dt } while (t < tt): at least one step
                }
            }
            lp = 0;
            if (!(tb < far)) done = true;  // every later element is beyond far
        }
    }
    return count;
}


template <bool SEAL>
__global__ void __launch_bounds__(kMarchWarps * 32) k_march_rays_train_warp(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid, const float bound,
    const float dt_gamma, const uint32_t max_steps, const uint32_t N, const uint32_t C, const uint32_t H, const uint32_t M,
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ aabb, const float min_near,
    float* __restrict__ nears_out, float* __restrict__ fars_out, float* __restrict__ xyzs, float* __restrict__ dirs,
    float* __restrict__ deltas, int* __restrict__ rays, int* __restrict__ counter, const float* __restrict__ noises,
    const __grid_constant__ seald_seal_mapper mp, uint8_t* __restrict__ seal_mask, const float* __restrict__ occ) {
    __shared__ float s_t[kMarchWarps][kMaxStepsSmem];
    __shared__ uint32_t s_cnt[kMarchWarps];
    __shared__ uint32_t s_off[kMarchWarps];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n = blockIdx.x * kMarchWarps + warp;
    const bool active = n < N;
    constexpr uint32_t FULL = 0xffffffffu;

    Ray r;
    float near = 0.f, far = 0.f, noise = 0.f;
    if (active) {
        const float* o = rays_o + (size_t)n * 3;
        const float* d = rays_d + (size_t)n * 3;
        r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
        r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
        r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
        if (nears) {
            near = nears[n];
            far = fars[n];
        } else {
            slab_test(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, aabb, min_near, near, far);
            if (nears_out && lane == 0) { nears_out[n] = near; fars_out[n] = far; }
        }
        noise = noises[n];
    }
    const MarchConst mc = make_march_const(bound, dt_gamma, max_steps, C, H);
    float t0 = near;
    t0 += clampf(t0 * dt_gamma, mc.dt_min, mc.dt_max) * noise;

    bool may_hit = active;
    if (occ && active) may_hit = clip_to_occupied(r, occ, near, far);

    uint32_t count = 0;
    if (may_hit && t0 < far) count = walk_sequential(r, mc, grid, t0, far, dt_gamma, max_steps, lane, s_t[warp]);
    if (lane == 0) s_cnt[warp] = active ? count : 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (uint32_t w = 0; w < kMarchWarps; w++) { s_off[w] = total; total += s_cnt[w]; }
        const uint32_t n_rays = min(kMarchWarps, N - blockIdx.x * kMarchWarps);
        const uint32_t base = atomicAdd(counter, total);
        atomicAdd(counter + 1, n_rays);
        for (uint32_t w = 0; w < kMarchWarps; w++) s_off[w] += base;
    }
    __syncthreads();
    if (!active) return;
    const uint32_t point_index = s_off[warp];
    if (lane == 0) {
        rays[n * 3] = n;
        rays[n * 3 + 1] = point_index;
        rays[n * 3 + 2] = count;
    }
    if (count == 0) return;
    if (point_index + count > M) {
        // overflow: the ray is dropped like in the reference (raymarching.cu:408), whose wrapper zero-filled the buffers
        // (raymarching.py:205-207).  Rows [point_index, M) belong to no ray but are still evaluated by the field kernels:
        // give them the reference's zeros instead of last step's samples.
        for (size_t s = (size_t)point_index + lane; s < M; s += 32) {
            xyzs[s * 3] = 0.f; xyzs[s * 3 + 1] = 0.f; xyzs[s * 3 + 2] = 0.f;
            dirs[s * 3] = 0.f; dirs[s * 3 + 1] = 0.f; dirs[s * 3 + 2] = 0.f;
            deltas[s * 2] = 0.f; deltas[s * 2 + 1] = 0.f;
            if (SEAL) seal_mask[s] = 0;
        }
        return;
    }
    __syncwarp();
    for (uint32_t j = lane; j < count; j += 32) {
        const float t = s_t[warp][j];
        const float x = clampf(r.ox + t * r.dx, -bound, bound);
        const float y = clampf(r.oy + t * r.dy, -bound, bound);
        const float z = clampf(r.oz + t * r.dz, -bound, bound);
        const float dt = clampf(t * dt_gamma, mc.dt_min, mc.dt_max);
        float last_t = t0;
        if (j > 0) {
            const float tp = s_t[warp][j - 1];
            last_t = tp + clampf(tp * dt_gamma, mc.dt_min, mc.dt_max);
        }
        const float t_new = t + dt;
        const size_t s = (size_t)point_index + j;
        if (SEAL) {
            float sx = x, sy = y, sz = z, sdx = r.dx, sdy = r.dy, sdz = r.dz;
            seal_mask[s] = seal_map_sample(mp, sx, sy, sz, sdx, sdy, sdz) ? 1 : 0;
            xyzs[s * 3] = sx; xyzs[s * 3 + 1] = sy; xyzs[s * 3 + 2] = sz;
            dirs[s * 3] = sdx; dirs[s * 3 + 1] = sdy; dirs[s * 3 + 2] = sdz;
        } else {
            xyzs[s * 3] = x; xyzs[s * 3 + 1] = y; xyzs[s * 3 + 2] = z;
            dirs[s * 3] = r.dx; dirs[s * 3 + 1] = r.dy; dirs[s * 3 + 2] = r.dz;
        }
        deltas[s * 2] = dt;
        deltas[s * 2 + 1] = t_new - last_t;
    }
}

// ------------------------------------------------------------------------------------------------
// training march with a constant step (dt_gamma == 0, the D-NeRF default), one warp per ray, NO sequential adds.
//
// The chain t_{k+1} = fl(t_k + dt) is piecewise linear in k: inside one binade every t_k is a multiple of ulp(t) and
// fl(t + dt) = t + I * ulp with I = dt / ulp rounded to nearest (a constant unless dt sits exactly half-way between two
// multiples).  So the chain is described by <= kChainSegs segments {k0, t(k0), I * ulp}; the element that crosses into the next
// binade is produced by ONE real fp32 add.  t_k = fma(k - k0, inc, t(k0)) is then exact, i.e. bit-identical to the
// reference's running sum (checked against sequential adds for every case the tests march).  With that:
//   phase A  every chain element of the ray is probed in parallel (no dependency between blocks of 32, the bitfield loads
//            pipeline); an empty element computes by itself WHERE its empty-space skip lands: the first element with
//            !(t_j < tt), found from the closed form and verified with exact comparisons.  succ[k] = occupied flag | landing;
//   phase B  the reference's control flow is a pointer chase through succ[] in shared memory (~1 LDS per visited element);
//   write    samples are recomputed from their chain indices and stored cooperatively (coalesced).
// Rays whose chain has a tie / leaves the supported exponent range fall back to the sequential walk of the legacy kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kChainSegs = 6;
constexpr uint32_t kChainChunk = 1024;   // chain elements resolved per phase A/B pass
constexpr int kChainMaxIndex = 32767;    // succ[] stores absolute indices in 15 bits

struct Chain {
    int k[kChainSegs];
    float t[kChainSegs];
    float inc[kChainSegs];
    int k_end;  // first chain index with !(t < far)
    bool ok;
};

__device__ __forceinline__ float chain_at(const Chain& c, const int j) {
    float ts = c.t[0], inc = c.inc[0];
    int ks = 0;
#pragma unroll
    for (int i = 1; i < kChainSegs; i++)
        if (j >= c.k[i]) { ts = c.t[i]; inc = c.inc[i]; ks = c.k[i]; }
    return __fmaf_rn((float)(j - ks), inc, ts);
}

__device__ __forceinline__ void chain_build(Chain& c, const float t0, const float dt, const float far) {
#pragma unroll
    for (int i = 0; i < kChainSegs; i++) { c.k[i] = 0x7fffffff; c.t[i] = 0.f; c.inc[i] = 0.f; }
    c.ok = true;
    c.k_end = 0;
    const uint32_t bd = __float_as_uint(dt);
    const int ed = (int)(bd >> 23);
    const uint32_t D = (bd & 0x7fffffu) | 0x800000u;
    if (ed == 0 || ed >= 255) { c.ok = false; return; }
    float t = t0;
    int k = 0;
    bool open = true;
#pragma unroll
    for (int s = 0; s < kChainSegs; s++) {
        if (open) {
            if (!(t < far)) {
                c.k_end = k;
                open = false;
            } else {
                const uint32_t bt = __float_as_uint(t);
                const int et = (int)(bt >> 23);
                const int shift = et - ed;
                if (et < 24 || et >= 255 || shift < 1 || shift > 23) { c.ok = false; return; }
                const uint32_t Mt = (bt & 0x7fffffu) | 0x800000u;
                const uint32_t q = D >> shift, rem = D & ((1u << shift) - 1u), half = 1u << (shift - 1);
                if (rem == half) { c.ok = false; return; }  // tie: the increment alternates with the parity of the sum
                const uint32_t I = q + (rem > half ? 1u : 0u);
                const uint32_t n_safe = (0xffffffu - Mt) / I;  // elements k .. k + n_safe stay inside the binade
                const float inc = __uint_as_float((uint32_t)(et - 23) << 23) * (float)I;
                c.k[s] = k; c.t[s] = t; c.inc[s] = inc;
                const float tl = __fmaf_rn((float)n_safe, inc, t);
                if (!(tl < far)) {
                    int j = __float2int_ru((far - t) / inc);
                    j = max(0, min(j, (int)n_safe));
                    while (j > 0 && !(__fmaf_rn((float)(j - 1), inc, t) < far)) j--;
                    while (__fmaf_rn((float)j, inc, t) < far) j++;
                    c.k_end = k + j;
                    open = false;
                } else {
                    t = tl + dt;  // the one real add that crosses into the next binade
                    k += (int)n_safe + 1;
                }
            }
        }
    }
    if (open) {
        if (t < far) c.ok = false; else c.k_end = k;
    }
    if (c.k_end > kChainMaxIndex) c.ok = false;
}

// probe_grid with the cascade selection short-cut for a single cascade and the cell index in fp32: 0.5 * v is exact and
// v * 0.5 * H has one rounding in fp32 exactly like the double product rounded to float, so the integers are the same
__device__ __forceinline__ bool probe_grid_fast(const Ray& r, const MarchConst& mc, const uint8_t* __restrict__ grid, const float t, Probe& p) {
    const float x = clampf(r.ox + t * r.dx, -mc.bound, mc.bound);
    const float y = clampf(r.oy + t * r.dy, -mc.bound, mc.bound);
    const float z = clampf(r.oz + t * r.dz, -mc.bound, mc.bound);
    const float dt = clampf(t * mc.dt_gamma, mc.dt_min, mc.dt_max);
    int level = 0;
    if (mc.C > 1) level = max(mip_from_pos(x, y, z, mc.C), mip_from_dt(dt, mc.H, mc.C));
    const float mip_bound = fminf(scalbnf(1.0f, level), mc.bound);
    const float mip_rbound = 1 / mip_bound;
    const float fH = (float)mc.H, top = (float)(mc.H - 1);
    const int nx = clampf((0.5f * (x * mip_rbound + 1)) * fH, 0.0f, top);
    const int ny = clampf((0.5f * (y * mip_rbound + 1)) * fH, 0.0f, top);
    const int nz = clampf((0.5f * (z * mip_rbound + 1)) * fH, 0.0f, top);
    const uint32_t index = level * mc.H3 + morton3D_enc(nx, ny, nz);
    const bool occ = __ldg(grid + index / 8) & (1 << (index % 8));
    p.x = x; p.y = y; p.z = z; p.dt = dt; p.mip_bound = mip_bound;
    p.nx = nx; p.ny = ny; p.nz = nz;
    return occ;
}

// per-ray record shared by the CTA: phase A is pooled over all warps (a ray that misses costs nothing, a long ray is
// probed by the whole CTA), so the kernel's duration follows the CTA's mean ray, not the longest ray of the batch
struct ChainRay {
    Ray r;
    int seg_k[kChainSegs + 1];
    float seg_t[kChainSegs];
    float seg_inc[kChainSegs];
    int k_end;
};

__device__ __forceinline__ float chain_at(const ChainRay& c, const int j) {
    float ts = c.seg_t[0], inc = c.seg_inc[0];
    int ks = 0;
#pragma unroll
    for (int i = 1; i < kChainSegs; i++)
        if (j >= c.seg_k[i]) { ts = c.seg_t[i]; inc = c.seg_inc[i]; ks = c.seg_k[i]; }
    return __fmaf_rn((float)(j - ks), inc, ts);
}

template <bool SEAL>
__global__ void __launch_bounds__(kMarchWarps * 32, 4) k_march_rays_train_chain(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid, const float bound,
    const uint32_t max_steps, const uint32_t N, const uint32_t C, const uint32_t H, const uint32_t M,
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ aabb, const float min_near,
    float* __restrict__ nears_out, float* __restrict__ fars_out, float* __restrict__ xyzs, float* __restrict__ dirs,
    float* __restrict__ deltas, int* __restrict__ rays, int* __restrict__ counter, const float* __restrict__ noises,
    const __grid_constant__ seald_seal_mapper mp, uint8_t* __restrict__ seal_mask, const float* __restrict__ occ) {
    // per ray: succ[kChainChunk] + sample indices[kMaxStepsSmem] (uint16), or - fallback rays - kMaxStepsSmem sample parameters (fp32)
    __shared__ __align__(16) uint16_t s_buf[kMarchWarps][kChainChunk + kMaxStepsSmem];
    __shared__ ChainRay s_ray[kMarchWarps];
    __shared__ uint32_t s_cnt[kMarchWarps];
    __shared__ uint32_t s_off[kMarchWarps];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n = blockIdx.x * kMarchWarps + warp;
    const bool active = n < N;
    uint16_t* s_samp = s_buf[warp] + kChainChunk;
    float* s_t = reinterpret_cast<float*>(s_buf[warp]);

    Ray r;
    float near = 0.f, far = 0.f, noise = 0.f;
    if (active) {
        const float* o = rays_o + (size_t)n * 3;
        const float* d = rays_d + (size_t)n * 3;
        r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
        r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
        r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
        if (nears) {
            near = nears[n];
            far = fars[n];
        } else {
            slab_test(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, aabb, min_near, near, far);
            if (nears_out && lane == 0) { nears_out[n] = near; fars_out[n] = far; }
        }
        noise = noises[n];
    }
    const MarchConst mc = make_march_const(bound, 0.0f, max_steps, C, H);
    const float dt_c = clampf(0.0f, mc.dt_min, mc.dt_max);
    float t0 = near;
    t0 += dt_c * noise;

    bool may_hit = active;
    if (occ && active) may_hit = clip_to_occupied(r, occ, near, far);

    uint32_t count = 0;
    bool use_chain = true;
    int my_k_end = 0;
    {
        Chain ch;
        ch.k_end = 0;
        if (may_hit && t0 < far) {
            chain_build(ch, t0, dt_c, far);
            use_chain = ch.ok;
            if (!use_chain) count = walk_sequential(r, mc, grid, t0, far, 0.0f, max_steps, lane, s_t);
        }
        my_k_end = use_chain ? ch.k_end : 0;
        if (lane == 0) {
            ChainRay& cr = s_ray[warp];
            cr.r = r;
#pragma unroll
            for (int i = 0; i < kChainSegs; i++) { cr.seg_k[i] = ch.k[i]; cr.seg_t[i] = ch.t[i]; cr.seg_inc[i] = ch.inc[i]; }
            cr.seg_k[0] = 0;
            cr.seg_k[kChainSegs] = 0x7fffffff;
            cr.k_end = my_k_end;
        }
    }
    __syncthreads();
    int k_max = 0;
#pragma unroll
    for (uint32_t w = 0; w < kMarchWarps; w++) k_max = max(k_max, s_ray[w].k_end);

    int kw = 0;  // this warp's walker on its own ray
    bool done = false;
    for (int c0 = 0; c0 < k_max; c0 += (int)kChainChunk) {
        // ---- phase A, pooled: blocks of 32 chain elements of every ray of the CTA are dealt round-robin to the warps
        uint32_t rot = 0;
        for (uint32_t w = 0; w < kMarchWarps; w++) {
            const ChainRay& cr = s_ray[w];
            const int k_end = cr.k_end;
            const int ke = min(k_end, c0 + (int)kChainChunk);
            if (ke <= c0) continue;
            const uint32_t nb = (uint32_t)(ke - c0 + 31) / 32u;
            const Ray rr = cr.r;
            for (uint32_t blk = (warp - rot) & (kMarchWarps - 1); blk < nb; blk += kMarchWarps) {
                const int kb = c0 + (int)blk * 32;
                const int k = kb + (int)lane;
                // segment of the block's first element (warp uniform); elements past its end take the general path
                int si = 0;
#pragma unroll
                for (int i = 1; i < kChainSegs; i++) si += (kb >= cr.seg_k[i]) ? 1 : 0;
                const int ks = cr.seg_k[si], kn = cr.seg_k[si + 1];
                const float ts = cr.seg_t[si], inc = cr.seg_inc[si];
                auto T = [&](const int q) {
                    if (q < kn) return __fmaf_rn((float)(q - ks), inc, ts);
                    float t2 = ts, i2 = inc;
                    int k2 = ks;
                    for (int i = si + 1; i < kChainSegs; i++)
                        if (q >= cr.seg_k[i]) { t2 = cr.seg_t[i]; i2 = cr.seg_inc[i]; k2 = cr.seg_k[i]; }
                    return __fmaf_rn((float)(q - k2), i2, t2);
                };
                bool is_occ = false;
                uint32_t v = 0;
                if (k < k_end) {
                    const float t = T(k);
                    Probe p;
                    is_occ = probe_grid_fast(rr, mc, grid, t, p);
                    if (!is_occ) {
                        const float tt = voxel_exit(rr, mc, p, t);
                        int j = ks + min(__float2int_ru(__fdividef(tt - ts, inc)), 1 << 20);
                        j = min(max(j, k + 1), k_end);
                        while (j < k_end && T(j) < tt) j++;
                        while (j > k + 1 && !(T(j - 1) < tt)) j--;
                        v = (uint32_t)j;
                    }
                }
                // an occupied element stores the length of the run of occupied elements it starts inside this block
                const uint32_t occ_mask = __ballot_sync(0xffffffffu, is_occ);
                if (is_occ) {
                    const uint32_t inv = ~(occ_mask >> lane);
                    v = 0x8000u | (inv ? (uint32_t)(__ffs(inv) - 1) : 32u);
                }
                if (k < k_end) s_buf[w][k - c0] = (uint16_t)v;
            }
            rot += nb;
        }
        __syncthreads();
        // ---- pointer jumping through runs of empty elements (in place: every value an element can read is a later element of
        // its own path that is reached through empties only, so the walk below visits the same occupied elements in the same order)
#pragma unroll 1
        for (int round = 0; round < 3; round++) {
            for (uint32_t w = 0; w < kMarchWarps; w++) {
                const int nk = min(s_ray[w].k_end, c0 + (int)kChainChunk) - c0;
                uint16_t* sw = s_buf[w];
                for (int e = (int)threadIdx.x; e < nk; e += (int)(kMarchWarps * 32)) {
                    const uint32_t v = sw[e];
                    const int ve = (int)v - c0;
                    if (!(v & 0x8000u) && ve < nk) {
                        const uint32_t u = sw[ve];
                        if (!(u & 0x8000u)) sw[e] = (uint16_t)u;
                    }
                }
            }
            __syncthreads();
        }
        // ---- phase B: the reference's loop as a pointer chase through this warp's own ray (warp uniform)
        if (use_chain && !done) {
            const int c1 = min(my_k_end, c0 + (int)kChainChunk);
            const uint16_t* s_succ = s_buf[warp];
            while (kw < c1) {
                const uint32_t v = s_succ[kw - c0];
                if (v & 0x8000u) {
                    const uint32_t n_run = min(v & 0xffu, max_steps - count);
                    if (lane < n_run) s_samp[count + lane] = (uint16_t)(kw + (int)lane);
                    count += n_run;
                    kw += (int)n_run;
                    if (count >= max_steps) { done = true; break; }
                } else {
                    kw = (int)v;
                }
            }
        }
        __syncthreads();
    }
    if (lane == 0) s_cnt[warp] = active ? count : 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (uint32_t w = 0; w < kMarchWarps; w++) { s_off[w] = total; total += s_cnt[w]; }
        const uint32_t n_rays = min(kMarchWarps, N - blockIdx.x * kMarchWarps);
        const uint32_t base = atomicAdd(counter, total);
        atomicAdd(counter + 1, n_rays);
        for (uint32_t w = 0; w < kMarchWarps; w++) s_off[w] += base;
    }
    __syncthreads();
    if (!active) return;
    const uint32_t point_index = s_off[warp];
    if (lane == 0) {
        rays[n * 3] = n;
        rays[n * 3 + 1] = point_index;
        rays[n * 3 + 2] = count;
    }
    if (count == 0) return;
    if (point_index + count > M) {  // overflow: dropped ray, zero fill of the unowned rows (see k_march_rays_train_warp)
        for (size_t s = (size_t)point_index + lane; s < M; s += 32) {
            xyzs[s * 3] = 0.f; xyzs[s * 3 + 1] = 0.f; xyzs[s * 3 + 2] = 0.f;
            dirs[s * 3] = 0.f; dirs[s * 3 + 1] = 0.f; dirs[s * 3 + 2] = 0.f;
            deltas[s * 2] = 0.f; deltas[s * 2 + 1] = 0.f;
            if (SEAL) seal_mask[s] = 0;
        }
        return;
    }
    __syncwarp();
    for (uint32_t j = lane; j < count; j += 32) {
        float t, last_t = t0;
        if (use_chain) {
            t = chain_at(s_ray[warp], (int)s_samp[j]);
            if (j > 0) last_t = chain_at(s_ray[warp], (int)s_samp[j - 1]) + dt_c;
        } else {
            t = s_t[j];
            if (j > 0) last_t = s_t[j - 1] + dt_c;
        }
        const float x = clampf(r.ox + t * r.dx, -bound, bound);
        const float y = clampf(r.oy + t * r.dy, -bound, bound);
        const float z = clampf(r.oz + t * r.dz, -bound, bound);
        const float t_new = t + dt_c;
        const size_t s = (size_t)point_index + j;
        if (SEAL) {
            float sx = x, sy = y, sz = z, sdx = r.dx, sdy = r.dy, sdz = r.dz;
            seal_mask[s] = seal_map_sample(mp, sx, sy, sz, sdx, sdy, sdz) ? 1 : 0;
            xyzs[s * 3] = sx; xyzs[s * 3 + 1] = sy; xyzs[s * 3 + 2] = sz;
            dirs[s * 3] = sdx; dirs[s * 3 + 1] = sdy; dirs[s * 3 + 2] = sdz;
        } else {
            xyzs[s * 3] = x; xyzs[s * 3 + 1] = y; xyzs[s * 3 + 2] = z;
            dirs[s * 3] = r.dx; dirs[s * 3 + 1] = r.dy; dirs[s * 3 + 2] = r.dz;
        }
        deltas[s * 2] = dt_c;
        deltas[s * 2 + 1] = t_new - last_t;
    }
}

// ------------------------------------------------------------------------------------------------
// compositing (training)
// ------------------------------------------------------------------------------------------------
__global__ void k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                      const int* __restrict__ rays, const uint32_t M, const uint32_t N, const float T_thresh,
                                      float* weights_sum, float* depth, float* image) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    uint32_t index = rays[n * 3];
    uint32_t offset = rays[n * 3 + 1];
    uint32_t num_steps = rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) {
        weights_sum[index] = 0;
        depth[index] = 0;
        image[index * 3] = 0;
        image[index * 3 + 1] = 0;
        image[index * 3 + 2] = 0;
        return;
    }
    sigmas += offset;
    rgbs += (size_t)offset * 3;
    deltas += (size_t)offset * 2;

    uint32_t step = 0;
    float T = 1.0f;
    float r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
    while (step < num_steps) {
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float weight = alpha * T;
        r += weight * rgbs[0];
        g += weight * rgbs[1];
        b += weight * rgbs[2];
        t += deltas[1];
        d += weight * t;
        ws += weight;
        T *= 1.0f - alpha;
        if (T < T_thresh) break;
        sigmas++;
        rgbs += 3;
        deltas += 2;
        step++;
    }
    weights_sum[index] = ws;
    depth[index] = d;
    image[index * 3] = r;
    image[index * 3 + 1] = g;
    image[index * 3 + 2] = b;
}

__global__ void k_composite_train_bwd(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image,
                                      const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                      const int* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ image,
                                      const uint32_t M, const uint32_t N, const float T_thresh, float* grad_sigmas, float* grad_rgbs) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    uint32_t index = rays[n * 3];
    uint32_t offset = rays[n * 3 + 1];
    uint32_t num_steps = rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;

    grad_weights_sum += index;
    grad_image += (size_t)index * 3;
    weights_sum += index;
    image += (size_t)index * 3;
    sigmas += offset;
    rgbs += (size_t)offset * 3;
    deltas += (size_t)offset * 2;
    grad_sigmas += offset;
    grad_rgbs += (size_t)offset * 3;

    uint32_t step = 0;
    float T = 1.0f;
    const float r_final = image[0], g_final = image[1], b_final = image[2], ws_final = weights_sum[0];
    float r = 0, g = 0, b = 0, ws = 0;
    while (step < num_steps) {
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float weight = alpha * T;
        r += weight * rgbs[0];
        g += weight * rgbs[1];
        b += weight * rgbs[2];
        ws += weight;
        T *= 1.0f - alpha;
        grad_rgbs[0] = grad_image[0] * weight;
        grad_rgbs[1] = grad_image[1] * weight;
        grad_rgbs[2] = grad_image[2] * weight;
        grad_sigmas[0] = deltas[0] * (
            grad_image[0] * (T * rgbs[0] - (r_final - r)) +
            grad_image[1] * (T * rgbs[1] - (g_final - g)) +
            grad_image[2] * (T * rgbs[2] - (b_final - b)) +
            grad_weights_sum[0] * (1 - ws_final)
        );
        if (T < T_thresh) break;
        sigmas++;
        rgbs += 3;
        deltas += 2;
        grad_sigmas++;
        grad_rgbs += 3;
        step++;
    }
}

// ------------------------------------------------------------------------------------------------
// compositing (training), one WARP per ray.
//
// The reference walks a ray's samples with one thread (raymarching.cu:516-576, :620-681): a chain of dependent global
// loads.  Here the 32 lanes load 32 consecutive samples at once (coalesced) and evaluate alpha = 1 - exp(-sigma * delta)
// in parallel; the front-to-back recurrence itself (T, the colour / depth / weight sums, the early stop at T < T_thresh)
// is then replayed IN ORDER over warp shuffles, every lane doing the same fp32 operations in the reference's expression
// order - so sums, the termination sample and the zero pattern of the gradients are exactly those of the serial loop.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kCompWarps = 8;

__global__ void __launch_bounds__(kCompWarps * 32) k_composite_train_fwd_warp(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                             const float* __restrict__ deltas, const int* __restrict__ rays,
                                                                             const uint32_t M, const uint32_t N, const float T_thresh,
                                                                             float* __restrict__ weights_sum, float* __restrict__ depth,
                                                                             float* __restrict__ image) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t n = blockIdx.x * kCompWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    float r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
    if (num_steps != 0 && offset + num_steps <= M) {
        float T = 1.0f;
        bool done = false;
        for (uint32_t base = 0; base < num_steps && !done; base += 32) {
            const uint32_t i = base + lane;
            float alpha = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, d1 = 0.f;
            if (i < num_steps) {
                const size_t s = (size_t)offset + i;
                const float2 dl = *reinterpret_cast<const float2*>(deltas + s * 2);
                alpha = 1.0f - __expf(-sigmas[s] * dl.x);
                d1 = dl.y;
                c0 = rgbs[s * 3]; c1 = rgbs[s * 3 + 1]; c2 = rgbs[s * 3 + 2];
            }
            const uint32_t cnt = min(32u, num_steps - base);
            for (uint32_t j = 0; j < cnt; j++) {
                const float a = __shfl_sync(FULL, alpha, j);
                const float x0 = __shfl_sync(FULL, c0, j), x1 = __shfl_sync(FULL, c1, j), x2 = __shfl_sync(FULL, c2, j);
                const float weight = a * T;
                r += weight * x0;
                g += weight * x1;
                b += weight * x2;
                t += __shfl_sync(FULL, d1, j);
                d += weight * t;
                ws += weight;
                T *= 1.0f - a;
                if (T < T_thresh) { done = true; break; }
            }
        }
    }
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[index * 3] = r;
        image[index * 3 + 1] = g;
        image[index * 3 + 2] = b;
    }
}

__global__ void __launch_bounds__(kCompWarps * 32) k_composite_train_bwd_warp(
    const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image, const float* __restrict__ sigmas,
    const float* __restrict__ rgbs, const float* __restrict__ deltas, const int* __restrict__ rays, const float* __restrict__ weights_sum,
    const float* __restrict__ image, const uint32_t M, const uint32_t N, const float T_thresh, float* __restrict__ grad_sigmas,
    float* __restrict__ grad_rgbs) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t n = blockIdx.x * kCompWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;
    const float gi0 = grad_image[(size_t)index * 3], gi1 = grad_image[(size_t)index * 3 + 1], gi2 = grad_image[(size_t)index * 3 + 2];
    const float gws = grad_weights_sum[index];
    const float r_final = image[(size_t)index * 3], g_final = image[(size_t)index * 3 + 1], b_final = image[(size_t)index * 3 + 2];
    const float ws_final = weights_sum[index];
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
    bool done = false;
    for (uint32_t base = 0; base < num_steps && !done; base += 32) {
        const uint32_t i = base + lane;
        float alpha = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, d0 = 0.f;
        const size_t s = (size_t)offset + i;
        if (i < num_steps) {
            d0 = deltas[s * 2];
            alpha = 1.0f - __expf(-sigmas[s] * d0);
            c0 = rgbs[s * 3]; c1 = rgbs[s * 3 + 1]; c2 = rgbs[s * 3 + 2];
        }
        const uint32_t cnt = min(32u, num_steps - base);
        float my_w = 0.f, my_gs = 0.f;
        uint32_t processed = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            const float a = __shfl_sync(FULL, alpha, j);
            const float x0 = __shfl_sync(FULL, c0, j), x1 = __shfl_sync(FULL, c1, j), x2 = __shfl_sync(FULL, c2, j);
            const float dd = __shfl_sync(FULL, d0, j);
            const float weight = a * T;
            r += weight * x0;
            g += weight * x1;
            b += weight * x2;
            ws += weight;
            T *= 1.0f - a;
            const float gsig = dd * (
                gi0 * (T * x0 - (r_final - r)) +
                gi1 * (T * x1 - (g_final - g)) +
                gi2 * (T * x2 - (b_final - b)) +
                gws * (1 - ws_final)
            );
            if (lane == j) { my_w = weight; my_gs = gsig; }
            processed = j + 1;
            if (T < T_thresh) { done = true; break; }
        }
        if (lane < processed) {
            grad_rgbs[s * 3] = gi0 * my_w;
            grad_rgbs[s * 3 + 1] = gi1 * my_w;
            grad_rgbs[s * 3 + 2] = gi2 * my_w;
            grad_sigmas[s] = my_gs;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Training tail in ONE kernel, one warp per ray: composite forward -> background blend + MSE loss and its gradient
// (dnerf/utils.py:74-85, dnerf/renderer.py:325-326) -> composite backward.  Same arithmetic, in the same order, as
// k_composite_train_fwd_warp + k_mse_loss_bg + k_composite_train_bwd_warp; what disappears are two launches, the
// grad_image / grad_weights_sum round trip and the zero-fill of grad_sigmas / grad_rgbs (samples behind the early stop
// get their zeros here).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCompWarps * 32) k_composite_train_loss_fused(
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas, const int* __restrict__ rays,
    const uint32_t M, const uint32_t N, const float T_thresh, const float* __restrict__ bg, const float* __restrict__ gt,
    const float inv_count, const float* __restrict__ loss_scale, float* __restrict__ weights_sum, float* __restrict__ depth,
    float* __restrict__ image, float* __restrict__ pred, float* __restrict__ loss_sum, float* __restrict__ grad_sigmas,
    float* __restrict__ grad_rgbs) {
    // The recurrence over a block of 32 samples is replayed from shared memory (broadcast reads that do not depend on the
    // loop-carried transmittance, so they pipeline) and WITHOUT a branch per sample: after the early stop the weights are
    // multiplied by zero instead (x + 0 * c == x exactly), which keeps every sum identical to the serial loop.
    __shared__ float s_loss[kCompWarps];
    __shared__ float s_v[kCompWarps][5][32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n = blockIdx.x * kCompWarps + warp;
    float(*sv)[32] = s_v[warp];
    float local = 0.0f;
    if (n < N) {
        const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
        const bool has = num_steps != 0 && offset + num_steps <= M;
        // ---- forward (k_composite_train_fwd_warp)
        float r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        if (has) {
            float T = 1.0f;
            bool live = true;
            for (uint32_t base = 0; base < num_steps && live; base += 32) {
                const uint32_t i = base + lane;
                float alpha = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, d1 = 0.f;
                if (i < num_steps) {
                    const size_t s = (size_t)offset + i;
                    const float2 dl = *reinterpret_cast<const float2*>(deltas + s * 2);
                    alpha = 1.0f - __expf(-sigmas[s] * dl.x);
                    d1 = dl.y;
                    c0 = rgbs[s * 3]; c1 = rgbs[s * 3 + 1]; c2 = rgbs[s * 3 + 2];
                }
                __syncwarp();
                sv[0][lane] = alpha; sv[1][lane] = c0; sv[2][lane] = c1; sv[3][lane] = c2; sv[4][lane] = d1;
                __syncwarp();
#pragma unroll
                for (uint32_t j = 0; j < 32; j++) {
                    const float a = sv[0][j];
                    const float weight = live ? a * T : 0.0f;  // samples past the block end have a == 0
                    r += weight * sv[1][j];
                    g += weight * sv[2][j];
                    b += weight * sv[3][j];
                    t += sv[4][j];
                    d += weight * t;
                    ws += weight;
                    T *= 1.0f - a;
                    live = live && !(T < T_thresh);
                }
            }
        }
        // ---- loss and its gradient (k_mse_loss_bg)
        const float scale = loss_scale ? *loss_scale : 1.0f;
        const float img[3] = {r, g, b};
        float gi[3], gws = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float bgc = bg ? bg[(size_t)index * 3 + c] : 1.0f;
            const float p = img[c] + (1 - ws) * bgc;
            const float diff = p - gt[(size_t)index * 3 + c];
            local += diff * diff;
            gi[c] = 2.0f * diff * inv_count * scale;
            gws -= bgc * gi[c];
            if (pred && lane == 0) pred[(size_t)index * 3 + c] = p;
        }
        if (lane == 0) {
            weights_sum[index] = ws;
            depth[index] = d;
            image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
        } else {
            local = 0.0f;  // every lane holds the same value: count the ray once
        }
        // ---- backward (k_composite_train_bwd_warp), zeros behind the early stop
        if (has) {
            const float r_final = r, g_final = g, b_final = b, ws_final = ws;
            const float gws_term = gws * (1 - ws_final);
            float T = 1.0f, ra = 0, ga = 0, ba = 0;
            bool live = true;
            for (uint32_t base = 0; base < num_steps; base += 32) {
                const uint32_t i = base + lane;
                const size_t s = (size_t)offset + i;
                float my_w = 0.f, my_gs = 0.f;
                bool my_live = false;
                if (live) {
                    float alpha = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, d0 = 0.f;
                    if (i < num_steps) {
                        d0 = deltas[s * 2];
                        alpha = 1.0f - __expf(-sigmas[s] * d0);
                        c0 = rgbs[s * 3]; c1 = rgbs[s * 3 + 1]; c2 = rgbs[s * 3 + 2];
                    }
                    __syncwarp();
                    sv[0][lane] = alpha; sv[1][lane] = c0; sv[2][lane] = c1; sv[3][lane] = c2; sv[4][lane] = d0;
                    __syncwarp();
#pragma unroll
                    for (uint32_t j = 0; j < 32; j++) {
                        const float a = sv[0][j];
                        const float x0 = sv[1][j], x1 = sv[2][j], x2 = sv[3][j];
                        const float weight = a * T;
                        const bool was_live = live;
                        if (was_live) {  // (uniform) the serial loop stops updating its running sums at the early stop
                            ra += weight * x0;
                            ga += weight * x1;
                            ba += weight * x2;
                            T *= 1.0f - a;
                        }
                        const float gsig = sv[4][j] * (
                            gi[0] * (T * x0 - (r_final - ra)) +
                            gi[1] * (T * x1 - (g_final - ga)) +
                            gi[2] * (T * x2 - (b_final - ba)) +
                            gws_term
                        );
                        if (lane == j) { my_w = weight; my_gs = gsig; my_live = was_live; }
                        live = was_live && !(T < T_thresh);
                    }
                }
                if (i < num_steps) {
                    grad_rgbs[s * 3] = my_live ? gi[0] * my_w : 0.0f;
                    grad_rgbs[s * 3 + 1] = my_live ? gi[1] * my_w : 0.0f;
                    grad_rgbs[s * 3 + 2] = my_live ? gi[2] * my_w : 0.0f;
                    grad_sigmas[s] = my_live ? my_gs : 0.0f;
                }
            }
        } else if (num_steps != 0 && offset < M) {
            // Sample-buffer overflow: this ray's range straddles the end of the buffer, the ray is dropped (raymarching.cu:523-535)
            // but its rows [offset, M) are still seen by the field backward (it runs over min(M, counter) rows).  The reference
            // zero-fills the gradients before the launch (raymarching.py:283-284); nobody else writes these rows here, so do it.
            for (size_t s = (size_t)offset + lane; s < M; s += 32) {
                grad_rgbs[s * 3] = 0.0f; grad_rgbs[s * 3 + 1] = 0.0f; grad_rgbs[s * 3 + 2] = 0.0f;
                grad_sigmas[s] = 0.0f;
            }
        }
    }
    if (lane == 0) s_loss[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.0f;
#pragma unroll
        for (uint32_t w = 0; w < kCompWarps; w++) v += s_loss[w];
        atomicAdd(loss_sum, v * inv_count);
    }
}

// ------------------------------------------------------------------------------------------------
// inference round (run_cuda eval branch, dnerf/renderer.py:349-376): march n_step samples per live ray, [field], composite.
// (These two kernels keep the reference's buffer contract for the drop-in ops; FusedRenderer's loop runs the SAMPLE-PACKED pair
// further down, k_march_round_pack / k_composite_round_pack.)
//
// The contract of the sample buffers is the reference's (row n * n_step + s belongs to the n-th entry of the alive list; unused
// slots are zero, delta == 0 being the "ray ended" marker, raymarching.cu:850), and so is every number a ray produces; the kernels
// are organised for the memory system instead of one strided 12-byte store / load per thread and sample:
//
//   k_march_round      a CTA of 128 rays walks the bitfield (thread per ray: after the first round the live rays sit inside the
//                      occupied region and find their n_step samples within a few probes) and collects the samples in shared memory;
//                      the CTA's slice of xyzs / dirs / deltas is contiguous in the output, so it is written with coalesced 128-bit
//                      stores, terminator slots included (no zero-fill pass over the buffers).  Live count and n_step are read on
//                      the device, so the host never synchronises inside the loop.
//   k_composite_round  the CTA's slice of sigmas / rgbs / deltas is staged with coalesced loads, each thread then runs the
//                      front-to-back recurrence of its ray from shared memory.  COMPACT: the survivors are appended to the next
//                      round's alive list (warp-aggregated, one atomic per warp) and the last CTA to finish derives the next round's
//                      schedule (n_alive, n_step = clamp(N / n_alive, 1, 8), steps marched) — composite + boolean-mask compaction +
//                      host-side bookkeeping of the reference in one launch.  Otherwise (drop-in op): dead rays are marked -1 in
//                      place, as composite_rays does.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kRoundThreads = 128;

template <bool SEAL>
__global__ void __launch_bounds__(kRoundThreads) k_march_round(
    uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive, const float* __restrict__ rays_t, const float* __restrict__ rays_o,
    const float* __restrict__ rays_d, const float bound, const float dt_gamma, const uint32_t max_steps, const uint32_t C, const uint32_t H,
    const uint8_t* __restrict__ grid, const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
    float* __restrict__ deltas, const float* __restrict__ noises, const int* __restrict__ n_alive_dev, const int* __restrict__ n_step_dev,
    const __grid_constant__ seald_seal_mapper mp, uint8_t* __restrict__ seal_mask, const float* __restrict__ occ, const uint32_t smem_rows) {
    extern __shared__ __align__(16) float s_round[];
    if (n_alive_dev) n_alive = min(n_alive, (uint32_t)max(*n_alive_dev, 0));
    if (n_step_dev) n_step = (uint32_t)max(*n_step_dev, 0);
    const uint32_t first = blockIdx.x * blockDim.x;  // first alive-list entry of this CTA
    if (first >= n_alive || n_step == 0) return;
    const uint32_t n_rays = min((uint32_t)blockDim.x, n_alive - first);
    const uint32_t rows = n_rays * n_step;            // sample rows this CTA owns: [first * n_step, first * n_step + rows)
    const bool staged = rows <= smem_rows;
    float* s_xyz = s_round;                           // [rows][3]
    float* s_dir = s_round + (size_t)smem_rows * 3;   // [rows][3]
    float* s_del = s_round + (size_t)smem_rows * 6;   // [rows][2]
    if (staged) {
        for (uint32_t i = threadIdx.x; i < rows * 8; i += blockDim.x) {
            if (i < rows * 3) s_xyz[i] = 0.f;
            else if (i < rows * 6) s_dir[i - rows * 3] = 0.f;
            else s_del[i - rows * 6] = 0.f;
        }
        __syncthreads();
    }
    const uint32_t n = first + threadIdx.x;
    if (threadIdx.x < n_rays) {
        const int index = rays_alive[n];
        const float noise = noises ? noises[n] : 0.0f;
        Ray r;
        const float* o = rays_o + (size_t)index * 3;
        const float* d = rays_d + (size_t)index * 3;
        r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
        r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
        r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
        const MarchConst mc = make_march_const(bound, dt_gamma, max_steps, C, H);
        float t = rays_t[index];
        float far = fars[index];
        t += clampf(t * dt_gamma, mc.dt_min, mc.dt_max) * noise;
        float last_t = t;
        if (occ && !clip_to_occupied(r, occ, t, far)) far = t;  // cannot meet an occupied cell any more: no samples, the ray ends
        const uint32_t row0 = staged ? threadIdx.x * n_step : n * n_step;
        float* oxyz = staged ? s_xyz : xyzs;
        float* odir = staged ? s_dir : dirs;
        float* odel = staged ? s_del : deltas;
        uint32_t step = 0;
        Probe p;
        while (t < far && step < n_step) {
            if (probe_grid(r, mc, grid, t, p)) {
                float sx = p.x, sy = p.y, sz = p.z, sdx = r.dx, sdy = r.dy, sdz = r.dz;
                if (SEAL) seal_mask[(size_t)n * n_step + step] = seal_map_sample(mp, sx, sy, sz, sdx, sdy, sdz) ? 1 : 0;
                const size_t row = row0 + step;
                oxyz[row * 3] = sx; oxyz[row * 3 + 1] = sy; oxyz[row * 3 + 2] = sz;
                odir[row * 3] = sdx; odir[row * 3 + 1] = sdy; odir[row * 3 + 2] = sdz;
                t += p.dt;
                odel[row * 2] = p.dt;
                odel[row * 2 + 1] = t - last_t;
                last_t = t;
                step++;
            } else {
                t = skip_voxel(r, mc, p, t);
            }
        }
        for (; step < n_step; step++) {  // terminator slots
            if (SEAL) seal_mask[(size_t)n * n_step + step] = 0;
            if (!staged) {
                const size_t row = row0 + step;
                oxyz[row * 3] = 0.f; oxyz[row * 3 + 1] = 0.f; oxyz[row * 3 + 2] = 0.f;
                odir[row * 3] = 0.f; odir[row * 3 + 1] = 0.f; odir[row * 3 + 2] = 0.f;
                odel[row * 2] = 0.f; odel[row * 2 + 1] = 0.f;
            }
        }
    }
    if (!staged) return;
    __syncthreads();
    // coalesced copy-out: the CTA's rows are contiguous in each output array (row base is a multiple of 4 floats when n_step * 128 is)
    const size_t base = (size_t)first * n_step;
    auto copy_out = [&](const float* __restrict__ src, float* __restrict__ dst, const uint32_t count) {
        if ((((uintptr_t)dst) & 15) == 0 && (count & 3) == 0) {
            for (uint32_t i = threadIdx.x; i < count / 4; i += blockDim.x) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
        } else {
            for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
        }
    };
    copy_out(s_xyz, xyzs + base * 3, rows * 3);
    copy_out(s_dir, dirs + base * 3, rows * 3);
    copy_out(s_del, deltas + base * 2, rows * 2);
}

template <bool COMPACT>
__global__ void __launch_bounds__(kRoundThreads) k_composite_round(
    uint32_t n_alive, uint32_t n_step, const float T_thresh, int* __restrict__ rays_alive, float* __restrict__ rays_t,
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas, float* __restrict__ weights_sum,
    float* __restrict__ depth, float* __restrict__ image, const int* __restrict__ n_alive_dev, const int* __restrict__ n_step_dev,
    int* __restrict__ next_alive, int* __restrict__ state, int* __restrict__ counters, const uint32_t budget, const uint32_t max_steps,
    const uint32_t max_n_step, const uint32_t smem_rows) {
    extern __shared__ __align__(16) float s_round[];
    if (n_alive_dev) n_alive = min(n_alive, (uint32_t)max(*n_alive_dev, 0));
    if (n_step_dev) n_step = (uint32_t)max(*n_step_dev, 0);
    const uint32_t first = blockIdx.x * blockDim.x;
    const bool cta_live = first < n_alive && n_step > 0;
    bool survive = false;
    int index = -1;
    if (cta_live) {
        const uint32_t n_rays = min((uint32_t)blockDim.x, n_alive - first);
        const uint32_t rows = n_rays * n_step;
        const bool staged = rows <= smem_rows;
        float* s_sig = s_round;                           // [rows]
        float* s_rgb = s_round + smem_rows;               // [rows][3]
        float* s_del = s_round + (size_t)smem_rows * 4;   // [rows][2]
        const size_t base = (size_t)first * n_step;
        if (staged) {
            for (uint32_t i = threadIdx.x; i < rows; i += blockDim.x) s_sig[i] = sigmas[base + i];
            for (uint32_t i = threadIdx.x; i < rows * 3; i += blockDim.x) s_rgb[i] = rgbs[base * 3 + i];
            for (uint32_t i = threadIdx.x; i < rows * 2; i += blockDim.x) s_del[i] = deltas[base * 2 + i];
            __syncthreads();
        }
        if (threadIdx.x < n_rays) {
            const uint32_t n = first + threadIdx.x;
            index = rays_alive[n];
            const float* sg = staged ? s_sig + threadIdx.x * n_step : sigmas + (size_t)n * n_step;
            const float* cl = staged ? s_rgb + (size_t)threadIdx.x * n_step * 3 : rgbs + (size_t)n * n_step * 3;
            const float* dl = staged ? s_del + (size_t)threadIdx.x * n_step * 2 : deltas + (size_t)n * n_step * 2;
            float t = rays_t[index];
            float ws = weights_sum[index], d = depth[index];
            float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
            uint32_t step = 0;
            for (; step < n_step; step++) {
                const float d0 = dl[step * 2];
                if (d0 == 0) break;                         // terminator slot: the march found no further sample
                const float alpha = 1.0f - __expf(-sg[step] * d0);
                const float T = 1 - ws;
                const float weight = alpha * T;
                ws += weight;
                t += dl[step * 2 + 1];
                d += weight * t;
                r += weight * cl[step * 3];
                g += weight * cl[step * 3 + 1];
                b += weight * cl[step * 3 + 2];
                if (T < T_thresh) break;                    // (tested on the transmittance BEFORE this sample, raymarching.cu:880-886)
            }
            survive = step == n_step;
            if (survive) rays_t[index] = t;
            else if (!COMPACT) rays_alive[n] = -1;
            weights_sum[index] = ws;
            depth[index] = d;
            image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
        }
    }
    if (!COMPACT) return;
    // ---- survivors -> next alive list (one atomic per warp), then the last CTA computes the next round's schedule
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t bal = __ballot_sync(0xffffffffu, survive);
    if (bal) {
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(bal) - 1)) base = (uint32_t)atomicAdd(counters, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        if (survive) next_alive[base + __popc(bal & ((1u << lane) - 1u))] = index;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(counters + 1, 1) + 1;
        if (done == (int)gridDim.x) {
            __threadfence();
            int n_new = atomicAdd(counters, 0);
            const int step = state[3] + state[1];
            state[4] += state[2];
            state[5] += (state[0] > 0) ? 1 : 0;
            if ((uint32_t)step >= max_steps) n_new = 0;
            int ns = 1;
            if (n_new > 0) ns = max(min((int)(budget / (uint32_t)n_new), (int)max_n_step), 1);
            state[0] = n_new;
            state[1] = ns;
            state[2] = n_new * ns;
            state[3] = step;
            counters[0] = 0;
            counters[1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SAMPLE-PACKED round (FusedRenderer's loop; the drop-in ops above keep the reference's n_step-rows-per-ray contract).
//
// With n_step rows reserved per live ray, 22% of the rows a frame hands to the field are empty (round 0: 640 000 rows for 98 774
// samples; rays that leave the occupied region mid-round).  Here a round's samples are stored back to back:
//
//   k_march_round_pack      thread per ray walks the bitfield and parks (t, dt, dt_ray) of its samples in a global scratch slot (12 B
//                           per sample, L2 resident; shared memory for 128 x 32 slots would cut the march to 16 warps per SM: measured
//                           +1.3 ms per frame); the CTA's sample count is scanned, ONE compare-and-swap reserves that many rows of the round's
//                           buffers (state[6] = rows so far = the live count the field kernels read), and the rows are written by a
//                           thread-per-row loop: coalesced stores, positions recomputed from t with the march's own expression
//                           (ray_point), the Seal proxy mapping applied there.  ray_rows[n] = (first row, count) of alive entry n.
//                           A CTA that no longer fits the buffers writes nothing and marks its rays deferred (count -1): they
//                           survive the round untouched - so n_step may be chosen optimistically (round 0).
//   k_composite_round_pack  the front-to-back recurrence over the ray's rows; survivors (n_step samples, still transparent, or
//                           deferred) appended to the next alive list; the last CTA derives the next round's schedule and clears
//                           the row counter.
// Every number a ray produces is the reference's: how its samples are cut into rounds and where they sit does not enter.
// ------------------------------------------------------------------------------------------------
// one bit per 8^3 block of cells: the Morton order makes a block 512 consecutive cells = 64 consecutive bytes of the bitfield
__global__ void k_coarse_bits(const uint8_t* __restrict__ bitfield, const uint32_t n_blocks, uint32_t* __restrict__ coarse) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    bool any = false;
    if (j < n_blocks) {
        const uint4* p = reinterpret_cast<const uint4*>(bitfield + (size_t)j * 64);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 v = p[i];
            any = any || (v.x | v.y | v.z | v.w) != 0u;
        }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, any);
    if ((threadIdx.x & 31u) == 0 && j < n_blocks) coarse[j >> 5] = bal;
}

template <bool SEAL>
__global__ void __launch_bounds__(kRoundThreads) k_march_round_pack(
    uint32_t n_alive, const uint32_t n_step_bound, const int* __restrict__ rays_alive, const float* __restrict__ rays_t,
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float bound, const float dt_gamma, const uint32_t max_steps,
    const uint32_t C, const uint32_t H, const uint8_t* __restrict__ grid, const float* __restrict__ fars, float* __restrict__ xyzs,
    float* __restrict__ dirs, float* __restrict__ deltas, float* __restrict__ noises, const int* __restrict__ n_alive_dev,
    const int* __restrict__ n_step_dev, const __grid_constant__ seald_seal_mapper mp, uint8_t* __restrict__ seal_mask,
    const float* __restrict__ occ, int* __restrict__ row_counter, const uint32_t cap, int2* __restrict__ ray_rows,
    float* __restrict__ stage, const uint32_t* __restrict__ coarse) {
    // stage: [n_alive * n_step][3] = t of the sample, dt, t_after - t_of_previous (slot (n, s) of alive entry n)
    // coarse (optional, C == 1, H <= 128): one bit per 8^3 cells of the bitfield (k_coarse_bits).  The samples of a ray are the chain
    // elements t_k whose cell is occupied, however the walk gets from one to the next — so an empty 8^3 block may be left in one
    // step (same adds, 1/8 of the probes) exactly like the reference leaves an empty cell.
    __shared__ uint32_t s_off[kRoundThreads + 1];
    __shared__ uint32_t s_coarse[128];
    __shared__ uint32_t s_wsum[kRoundThreads / 32];
    __shared__ int s_base;
    n_alive = min(n_alive, (uint32_t)max(*n_alive_dev, 0));
    const uint32_t n_step = min((uint32_t)max(*n_step_dev, 0), n_step_bound);
    const uint32_t first = blockIdx.x * blockDim.x;
    if (first >= n_alive || n_step == 0) return;
    const uint32_t n_rays = min((uint32_t)blockDim.x, n_alive - first);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (coarse) {
        const uint32_t n_words = (H / 8) * (H / 8) * (H / 8) / 32;
        for (uint32_t i = tid; i < n_words; i += blockDim.x) s_coarse[i] = coarse[i];
        __syncthreads();
    }
    uint32_t count = 0;
    int my_index = -1;
    if (tid < n_rays) {
        const uint32_t n = first + tid;
        const int index = rays_alive[n];
        my_index = index;
        // perturbed start (raymarching.cu:741): per RAY (the first alive list is not the identity, k_render_init) and applied the first
        // time the ray is really marched — the entry is cleared below unless this CTA is deferred, later rounds add an exact 0
        const float noise = noises ? noises[index] : 0.0f;
        Ray r;
        const float* o = rays_o + (size_t)index * 3;
        const float* d = rays_d + (size_t)index * 3;
        r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
        r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
        r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
        const MarchConst mc = make_march_const(bound, dt_gamma, max_steps, C, H);
        float t = rays_t[index];
        float far = fars[index];
        t += clampf(t * dt_gamma, mc.dt_min, mc.dt_max) * noise;
        float last_t = t;
        if (occ && !clip_to_occupied(r, occ, t, far)) far = t;
        float* slot = stage + (size_t)n * n_step * 3;
        Probe p;
        while (t < far && count < n_step) {
            if (coarse) {
                // the cell of probe_grid (cascade level 0: C == 1), then its 8^3 block
                float x, y, z;
                ray_point(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, mc.bound, t, x, y, z);
                const float mip_bound = fminf(1.0f, mc.bound), mip_rbound = 1 / mip_bound;
                const int nx = clampf(0.5 * (x * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
                const int ny = clampf(0.5 * (y * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
                const int nz = clampf(0.5 * (z * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
                const uint32_t ci = morton3D_enc(nx >> 3, ny >> 3, nz >> 3);
                if (!((s_coarse[ci >> 5] >> (ci & 31u)) & 1u)) {
                    const float rHc = 8.0f * mc.rH;
                    const float tx = ((((nx >> 3) + 0.5f + 0.5f * signf(r.dx)) * rHc * 2 - 1) * mip_bound - x) * r.rdx;
                    const float ty = ((((ny >> 3) + 0.5f + 0.5f * signf(r.dy)) * rHc * 2 - 1) * mip_bound - y) * r.rdy;
                    const float tz = ((((nz >> 3) + 0.5f + 0.5f * signf(r.dz)) * rHc * 2 - 1) * mip_bound - z) * r.rdz;
                    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
                    do {
                        t += clampf(t * mc.dt_gamma, mc.dt_min, mc.dt_max);
                    } while (t < tt);
                    continue;
                }
            }
            if (probe_grid(r, mc, grid, t, p)) {
                slot[count * 3] = t;
                t += p.dt;
                slot[count * 3 + 1] = p.dt;
                slot[count * 3 + 2] = t - last_t;
                last_t = t;
                count++;
            } else {
                t = skip_voxel(r, mc, p, t);
            }
        }
    }
    // ---- exclusive scan of the counts over the CTA, one reservation
    const uint32_t incl = warp_inclusive_scan(count, lane);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (uint32_t w = 0; w < kRoundThreads / 32; w++) before += (w < warp) ? s_wsum[w] : 0u;
    s_off[tid] = before + incl - count;
    if (tid == kRoundThreads - 1) {
        const uint32_t total = before + incl;
        s_off[kRoundThreads] = total;
        // ONE atomicAdd (a compare-and-swap loop convoys: every success makes all other CTAs retry, ~0.4 us per CTA, serialised).
        // A reservation that passes `cap` is not rolled back: every later one starts beyond cap and fails too, so the written rows
        // stay contiguous from 0; the counter then overshoots (the field clamps its live count to the buffer, seald_composite_rays_pack
        // clears it) and the rows between the last success and cap hold stale samples nobody composites.
        int base = 0;
        if (total) {
            base = atomicAdd(row_counter, (int)total);
            if ((uint32_t)base + total > cap) {
                base = -1;
                atomicAdd(row_counter + 1, (int)n_rays);  // state[7]: rays deferred this round
            }
        }
        s_base = base;
    }
    __syncthreads();
    const int base = s_base;
    if (tid < n_rays) ray_rows[first + tid] = (base < 0) ? make_int2(0, -1) : make_int2(base + (int)s_off[tid], (int)count);
    if (base < 0) return;
    if (noises && my_index >= 0) noises[my_index] = 0.0f;
    // ---- the CTA's rows, one thread per row (consecutive threads -> consecutive rows)
    const uint32_t total = s_off[kRoundThreads];
    for (uint32_t i = tid; i < total; i += blockDim.x) {
        uint32_t lo = 0, hi = kRoundThreads;  // last ray whose first row is <= i (rays without samples share their successor's offset)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= i) lo = mid; else hi = mid;
        }
        const uint32_t sidx = i - s_off[lo];
        const float* slot = stage + ((size_t)(first + lo) * n_step + sidx) * 3;
        const int index = rays_alive[first + lo];
        const float* o = rays_o + (size_t)index * 3;
        const float* d = rays_d + (size_t)index * 3;
        float sdx = d[0], sdy = d[1], sdz = d[2], sx, sy, sz;
        ray_point(o[0], o[1], o[2], sdx, sdy, sdz, bound, slot[0], sx, sy, sz);
        const size_t row = (size_t)base + i;
        if (SEAL) seal_mask[row] = seal_map_sample(mp, sx, sy, sz, sdx, sdy, sdz) ? 1 : 0;
        xyzs[row * 3] = sx; xyzs[row * 3 + 1] = sy; xyzs[row * 3 + 2] = sz;
        dirs[row * 3] = sdx; dirs[row * 3 + 1] = sdy; dirs[row * 3 + 2] = sdz;
        *reinterpret_cast<float2*>(deltas + row * 2) = make_float2(slot[1], slot[2]);
    }
}

// Frame prologue of the packed loop in ONE launch (the reference: near_far_from_aabb + five torch fills / copies, and every ray of the
// frame in the first alive list): near / far (k_near_far's slab test), rays_t = near, zeroed accumulators, and an alive list that only
// holds the rays whose [near, far] meets the occupied region — for a 800x800 frame of the benchmark scene 85% of the rays never
// produce a sample; they keep weights_sum = 0 (background) exactly as if they had been marched.  The last CTA writes the round state:
// n_step of round 0 = clamp(budget / n_alive, n_step_min, max_n_step).
__global__ void __launch_bounds__(256) k_render_init(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                      const float* __restrict__ aabb, const float* __restrict__ occ, const uint32_t N,
                                                      const float min_near, float* __restrict__ nears, float* __restrict__ fars,
                                                      float* __restrict__ rays_t, float* __restrict__ weights_sum, float* __restrict__ depth,
                                                      float* __restrict__ image, int* __restrict__ alive, int* __restrict__ state,
                                                      int* __restrict__ counters, const uint32_t budget, const uint32_t n_step_min,
                                                      const uint32_t max_n_step) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = false;
    if (n < N) {
        const float* o = rays_o + (size_t)n * 3;
        const float* d = rays_d + (size_t)n * 3;
        Ray r;
        r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
        r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
        r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
        float near, far;
        slab_test(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, aabb, min_near, near, far);
        nears[n] = near;
        fars[n] = far;
        rays_t[n] = near;
        weights_sum[n] = 0.f;
        depth[n] = 0.f;
        image[(size_t)n * 3] = 0.f; image[(size_t)n * 3 + 1] = 0.f; image[(size_t)n * 3 + 2] = 0.f;
        float f = far;
        live = near < far && (!occ || clip_to_occupied(r, occ, near, f));
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t bal = __ballot_sync(0xffffffffu, live);
    if (bal) {
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(bal) - 1)) base = (uint32_t)atomicAdd(counters, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        if (live) alive[base + __popc(bal & ((1u << lane) - 1u))] = (int)n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(counters + 1, 1) + 1;
        if (done == (int)gridDim.x) {
            __threadfence();
            const int n_alive = atomicAdd(counters, 0);
            int ns = (int)n_step_min;
            if (n_alive > 0) ns = max(min((int)(budget / (uint32_t)n_alive), (int)max_n_step), (int)n_step_min);
            state[0] = n_alive;
            state[1] = ns;
            state[2] = 0; state[3] = 0; state[4] = 0; state[5] = 0; state[6] = 0; state[7] = 0;
            counters[0] = 0;
            counters[1] = 0;
            __threadfence();
        }
    }
}

__global__ void __launch_bounds__(kRoundThreads) k_composite_round_pack(
    uint32_t n_alive, const float T_thresh, const int* __restrict__ rays_alive, float* __restrict__ rays_t, const float* __restrict__ sigmas,
    const float* __restrict__ rgbs, const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ depth,
    float* __restrict__ image, int* __restrict__ next_alive, int* __restrict__ state, int* __restrict__ counters,
    const int2* __restrict__ ray_rows, const uint32_t budget, const uint32_t max_steps, const uint32_t max_n_step, const uint32_t cap) {
    n_alive = min(n_alive, (uint32_t)max(state[0], 0));
    const int n_step = max(state[1], 0);
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    bool survive = false;
    int index = -1;
    if (n < n_alive && n_step > 0) {
        index = rays_alive[n];
        const int2 rr = ray_rows[n];
        if (rr.y < 0) {
            survive = true;  // deferred: the round's buffers were full, nothing marched
        } else {
            const float* sg = sigmas + rr.x;
            const float* cl = rgbs + (size_t)rr.x * 3;
            const float* dl = deltas + (size_t)rr.x * 2;
            float t = rays_t[index];
            float ws = weights_sum[index], d = depth[index];
            float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
            int step = 0;
            for (; step < rr.y; step++) {
                const float2 dd = *reinterpret_cast<const float2*>(dl + step * 2);
                const float alpha = 1.0f - __expf(-sg[step] * dd.x);
                const float T = 1 - ws;
                const float weight = alpha * T;
                ws += weight;
                t += dd.y;
                d += weight * t;
                r += weight * cl[step * 3];
                g += weight * cl[step * 3 + 1];
                b += weight * cl[step * 3 + 2];
                if (T < T_thresh) break;  // (tested on the transmittance BEFORE this sample, raymarching.cu:880-886)
            }
            survive = (step == n_step);   // a full round of samples and still transparent; fewer samples = the ray ended
            if (survive) rays_t[index] = t;
            weights_sum[index] = ws;
            depth[index] = d;
            image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
        }
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t bal = __ballot_sync(0xffffffffu, survive);
    if (bal) {
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(bal) - 1)) base = (uint32_t)atomicAdd(counters, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        if (survive) next_alive[base + __popc(bal & ((1u << lane) - 1u))] = index;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(counters + 1, 1) + 1;
        if (done == (int)gridDim.x) {
            __threadfence();
            int n_new = atomicAdd(counters, 0);
            // (a round that deferred rays does not count towards max_steps: those rays did not move)
            const int deferred = state[7];
            const int step = state[3] + (deferred ? 0 : state[1]);
            state[4] += min(state[6], (int)cap);        // rows the field evaluated this round
            state[5] += (state[0] > 0) ? 1 : 0;
            if ((uint32_t)step >= max_steps) n_new = 0;
            int ns = 1;
            if (n_new > 0) ns = max(min((int)(budget / (uint32_t)n_new), (int)max_n_step), 1);
            state[0] = n_new;
            state[1] = ns;
            state[2] += deferred;                       // rays deferred so far (statistics)
            state[3] = step;
            state[6] = 0;
            state[7] = 0;
            counters[0] = 0;
            counters[1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// order-preserving compaction of the alive list: three tiny kernels (count per tile, scan of tile totals,
// scatter).  Tiles of 1024 entries.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTile = 1024;

__global__ void k_compact_count(const int* __restrict__ rays_alive, uint32_t n_alive, const int* __restrict__ n_alive_dev, int* __restrict__ tile_counts) {
    if (n_alive_dev) n_alive = min(n_alive, (uint32_t)max(*n_alive_dev, 0));
    __shared__ uint32_t s_warp[32];
    const uint32_t i = blockIdx.x * kTile + threadIdx.x;
    const bool keep = i < n_alive && rays_alive[i] >= 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = s_warp[threadIdx.x];
        v = warp_inclusive_scan(v, threadIdx.x);
        if (threadIdx.x == 31) tile_counts[blockIdx.x] = v;
    }
}

// single CTA: exclusive scan of tile counts in place; writes the total to n_out_dev
__global__ void k_compact_scan(int* __restrict__ tile_counts, const uint32_t n_tiles, int* __restrict__ n_out_dev) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? (uint32_t)tile_counts[i] : 0u;
        const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
        uint32_t incl = warp_inclusive_scan(v, lane);
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint32_t s = s_warp[lane];
            s = warp_inclusive_scan(s, lane);
            s_warp[lane] = s;
        }
        __syncthreads();
        const uint32_t prefix = (w ? s_warp[w - 1] : 0u) + s_carry;
        if (i < n_tiles) tile_counts[i] = (int)(prefix + incl - v);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = prefix + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out_dev = (int)s_carry;
}

__global__ void k_compact_scatter(const int* __restrict__ rays_alive, uint32_t n_alive, const int* __restrict__ n_alive_dev,
                                  const int* __restrict__ tile_offsets, int* __restrict__ out) {
    if (n_alive_dev) n_alive = min(n_alive, (uint32_t)max(*n_alive_dev, 0));
    __shared__ uint32_t s_warp[32];
    const uint32_t i = blockIdx.x * kTile + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const int v = i < n_alive ? rays_alive[i] : -1;
    const bool keep = v >= 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    if (w == 0) {
        uint32_t s = s_warp[lane];
        s = warp_inclusive_scan(s, lane);
        s_warp[lane] = s;
    }
    __syncthreads();
    if (keep) {
        const uint32_t pos = (uint32_t)tile_offsets[blockIdx.x] + (w ? s_warp[w - 1] : 0u) + __popc(bal & ((1u << lane) - 1u));
        out[pos] = v;
    }
}


// Device-side schedule of the inference loop (dnerf/renderer.py:350-376): after a round has been composited and the alive
// list compacted, advance the step counter and derive the next round's parameters without a host round trip.
//   state[0] = n_alive (in: count after compaction; forced to 0 once `max_steps` is reached), state[1] = n_step of the NEXT
//   round = clamp(N / n_alive, 1, 8), state[2] = n_alive * n_step (live sample rows), state[3] = steps marched so far,
//   state[4] = live sample rows evaluated so far, state[5] = non-empty rounds so far (statistics).
__global__ void k_render_schedule(int* __restrict__ state, const int* __restrict__ n_alive_new, const uint32_t N, const uint32_t max_steps,
                                  const uint32_t max_n_step) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int step = state[3] + state[1];
    state[4] += state[2];
    state[5] += (state[0] > 0) ? 1 : 0;
    int n_alive = *n_alive_new;
    if ((uint32_t)step >= max_steps) n_alive = 0;
    int n_step = 1;
    if (n_alive > 0) n_step = max(min((int)(N / (uint32_t)n_alive), (int)max_n_step), 1);
    state[0] = n_alive;
    state[1] = n_step;
    state[2] = n_alive * n_step;
    state[3] = step;
}

}  // namespace seald

using namespace seald;

extern "C" int seald_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb6, uint32_t N, float min_near,
                                        float* nears, float* fars, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !aabb6 || !nears || !fars) return SEALD_E_BADARG;
    k_near_far<<<div_up(N, 128u), 128, 0, to_stream(stream)>>>(rays_o, rays_d, aabb6, N, min_near, nears, fars);
    return launch_status();
}

extern "C" int seald_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !coords) return SEALD_E_BADARG;
    k_sph_from_ray<<<div_up(N, 128u), 128, 0, to_stream(stream)>>>(rays_o, rays_d, radius, N, coords);
    return launch_status();
}

extern "C" int seald_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!coords || !indices) return SEALD_E_BADARG;
    k_morton3D<<<div_up(N, 256u), 256, 0, to_stream(stream)>>>(coords, N, indices);
    return launch_status();
}

extern "C" int seald_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!coords || !indices) return SEALD_E_BADARG;
    k_morton3D_invert<<<div_up(N, 256u), 256, 0, to_stream(stream)>>>(indices, N, coords);
    return launch_status();
}

extern "C" int seald_packbits(const float* grid, uint32_t n_bytes, float thresh, uint8_t* bitfield, seald_stream_t stream) {
    if (n_bytes == 0) return 0;
    if (!grid || !bitfield) return SEALD_E_BADARG;
    if (((uintptr_t)grid % 16 == 0) && ((uintptr_t)bitfield % 4 == 0)) {
        const uint32_t n_thr = n_bytes / 4 + (n_bytes % 4);
        k_packbits<<<div_up(n_thr, 256u), 256, 0, to_stream(stream)>>>(grid, n_bytes, thresh, bitfield);
    } else {
        k_packbits_bytes<<<div_up(n_bytes, 128u), 128, 0, to_stream(stream)>>>(grid, n_bytes, thresh, bitfield);
    }
    return launch_status();
}

static int march_train_impl(const float* rays_o, const float* rays_d, const uint8_t* bitfield, float bound, float dt_gamma,
                            uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears, const float* fars,
                            const float* aabb6, float min_near, float* nears_out, float* fars_out, float* xyzs, float* dirs, float* deltas,
                            int32_t* rays, int32_t* counter, const float* noises, const seald_seal_mapper* mapper, uint8_t* mask,
                            const float* occ, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !bitfield || !xyzs || !dirs || !deltas || !rays || !counter || !noises) return SEALD_E_BADARG;
    if ((nears == nullptr) != (fars == nullptr)) return SEALD_E_BADARG;
    if (!nears && !aabb6) return SEALD_E_BADARG;
    if ((nears_out == nullptr) != (fars_out == nullptr)) return SEALD_E_BADARG;
    if (C == 0 || H == 0 || max_steps == 0) return SEALD_E_BADARG;
    static const seald_seal_mapper no_mapper = {};
    const seald_seal_mapper& mp = mapper ? *mapper : no_mapper;
    cudaStream_t st = to_stream(stream);
    // small batches (training: 4096 rays) are latency bound: one warp per ray.  Large batches (whole images) have enough
    // rays to fill the machine with one thread per ray, which does less total work.
    static const bool no_chain = getenv("SEALD_MARCH_CHAIN") && atoi(getenv("SEALD_MARCH_CHAIN")) == 0;  // measurement switch
    static const uint32_t chain_max = getenv("SEALD_MARCH_CHAIN_MAX") ? (uint32_t)atoll(getenv("SEALD_MARCH_CHAIN_MAX")) : 65536u;
    if (N <= chain_max && max_steps <= kMaxStepsSmem && dt_gamma == 0.0f && !no_chain) {
        // constant step: closed-form chain, every element probed in parallel + pointer chase
        auto k = mapper ? k_march_rays_train_chain<true> : k_march_rays_train_chain<false>;
        k<<<div_up(N, kMarchWarps), kMarchWarps * 32, 0, st>>>(rays_o, rays_d, bitfield, bound, max_steps, N, C, H, M, nears, fars, aabb6, min_near,
                                                              nears_out, fars_out, xyzs, dirs, deltas, rays, counter, noises, mp, mask, occ);
        return launch_status();
    }
    if (N <= 65536u && max_steps <= kMaxStepsSmem) {
        auto k = mapper ? k_march_rays_train_warp<true> : k_march_rays_train_warp<false>;
        k<<<div_up(N, kMarchWarps), kMarchWarps * 32, 0, st>>>(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, aabb6,
                                                              min_near, nears_out, fars_out, xyzs, dirs, deltas, rays, counter, noises, mp, mask, occ);
        return launch_status();
    }
    const uint32_t threads = (N >= 4u * SEALD_NUM_SMS * 128u) ? 128u : 32u;
    auto k = mapper ? k_march_rays_train<true> : k_march_rays_train<false>;
    k<<<div_up(N, threads), threads, 0, st>>>(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, aabb6, min_near,
                                             nears_out, fars_out, xyzs, dirs, deltas, rays, counter, noises, mp, mask, occ);
    return launch_status();
}

static int check_fusable_mapper(const seald_seal_mapper* mp, const uint8_t* mask) {
    if (!mp || !mask) return SEALD_E_BADARG;
    if (mp->type != SEALD_SEAL_BBOX && mp->type != SEALD_SEAL_BRUSH) return SEALD_E_UNSUPPORTED;  // anchor: batch-wide early exit, seal.cu
    if (mp->n_bounds <= 0 || mp->n_tris <= 0 || !mp->bounds || !mp->tris) return SEALD_E_BADARG;
    if (mp->type == SEALD_SEAL_BRUSH && mp->attenuation_mode == SEALD_SEAL_ATT_LINEAR && (mp->n_border <= 0 || !mp->border)) return SEALD_E_BADARG;
    return 0;
}

extern "C" int seald_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* bitfield, float bound, float dt_gamma,
                                      uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                      const float* fars, const float* aabb6, float min_near, float* nears_out, float* fars_out,
                                      float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                                      const float* occ_aabb6, seald_stream_t stream) {
    return march_train_impl(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, aabb6, min_near, nears_out, fars_out,
                            xyzs, dirs, deltas, rays, counter, noises, nullptr, nullptr, occ_aabb6, stream);
}

extern "C" int seald_march_rays_train_seal(const float* rays_o, const float* rays_d, const uint8_t* bitfield, float bound, float dt_gamma,
                                           uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                           const float* fars, const float* aabb6, float min_near, float* nears_out, float* fars_out,
                                           float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                                           const seald_seal_mapper* mapper, uint8_t* mask, const float* occ_aabb6, seald_stream_t stream) {
    if (int rc = check_fusable_mapper(mapper, mask)) return rc;
    return march_train_impl(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, aabb6, min_near, nears_out, fars_out,
                            xyzs, dirs, deltas, rays, counter, noises, mapper, mask, occ_aabb6, stream);
}

extern "C" int seald_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays, uint32_t M,
                                                  uint32_t N, float T_thresh, float* weights_sum, float* depth, float* image,
                                                  seald_stream_t stream) {
    if (N == 0) return 0;
    if (!rays || !weights_sum || !depth || !image) return SEALD_E_BADARG;
    if (M > 0 && (!sigmas || !rgbs || !deltas)) return SEALD_E_BADARG;
    if (((uintptr_t)deltas & 7) == 0) {
        k_composite_train_fwd_warp<<<div_up(N, kCompWarps), kCompWarps * 32, 0, to_stream(stream)>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh,
                                                                                                   weights_sum, depth, image);
        return launch_status();
    }
    const uint32_t threads = (N >= 4u * SEALD_NUM_SMS * 128u) ? 128u : 32u;
    k_composite_train_fwd<<<div_up(N, threads), threads, 0, to_stream(stream)>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
    return launch_status();
}

extern "C" int seald_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                                   const float* rgbs, const float* deltas, const int32_t* rays, const float* weights_sum,
                                                   const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                                   float* grad_rgbs, seald_stream_t stream) {
    if (N == 0 || M == 0) return 0;
    if (!grad_weights_sum || !grad_image || !sigmas || !rgbs || !deltas || !rays || !weights_sum || !image || !grad_sigmas || !grad_rgbs)
        return SEALD_E_BADARG;
    k_composite_train_bwd_warp<<<div_up(N, kCompWarps), kCompWarps * 32, 0, to_stream(stream)>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas,
                                                                                               rays, weights_sum, image, M, N, T_thresh,
                                                                                               grad_sigmas, grad_rgbs);
    return launch_status();
}

static uint32_t round_smem_rows(uint32_t n_step_max) {
    // rows of shared-memory staging per CTA: 128 rays x n_step (32 B per row), capped at 64 KiB; beyond that the kernels fall back to
    // direct (strided) access
    const uint32_t rows = kRoundThreads * (n_step_max ? n_step_max : 1u);
    return rows <= 2048u ? rows : 0u;
}

static int march_impl(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                      const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* bitfield,
                      const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas, const float* noises,
                      const int32_t* n_alive_dev, const int32_t* n_step_dev, const seald_seal_mapper* mapper, uint8_t* mask, const float* occ,
                      seald_stream_t stream) {
    (void)nears;
    if (n_alive == 0 || n_step == 0) return 0;
    if (!rays_alive || !rays_t || !rays_o || !rays_d || !bitfield || !fars || !xyzs || !dirs || !deltas) return SEALD_E_BADARG;
    if (C == 0 || H == 0 || max_steps == 0) return SEALD_E_BADARG;
    static const seald_seal_mapper no_mapper = {};
    const seald_seal_mapper& mp = mapper ? *mapper : no_mapper;
    const uint32_t smem_rows = round_smem_rows(n_step);   // n_step is the launch's upper bound when the live value is on the device
    const size_t smem = (size_t)smem_rows * 8 * sizeof(float);
    auto k = mapper ? k_march_round<true> : k_march_round<false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    k<<<div_up(n_alive, kRoundThreads), kRoundThreads, smem, to_stream(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound,
                                                                                 dt_gamma, max_steps, C, H, bitfield, fars, xyzs, dirs, deltas,
                                                                                 noises, n_alive_dev, n_step_dev, mp, mask, occ, smem_rows);
    return launch_status();
}

extern "C" int seald_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                                const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                                const uint8_t* bitfield, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                const float* noises, const int32_t* n_alive_dev, const int32_t* n_step_dev, const float* occ_aabb6,
                                seald_stream_t stream) {
    return march_impl(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, bitfield, nears, fars, xyzs, dirs,
                      deltas, noises, n_alive_dev, n_step_dev, nullptr, nullptr, occ_aabb6, stream);
}

extern "C" int seald_march_rays_seal(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                                     const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                                     const uint8_t* bitfield, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                     const float* noises, const int32_t* n_alive_dev, const int32_t* n_step_dev,
                                     const seald_seal_mapper* mapper, uint8_t* mask, const float* occ_aabb6, seald_stream_t stream) {
    if (int rc = check_fusable_mapper(mapper, mask)) return rc;
    return march_impl(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, bitfield, nears, fars, xyzs, dirs,
                      deltas, noises, n_alive_dev, n_step_dev, mapper, mask, occ_aabb6, stream);
}

extern "C" int seald_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t, const float* sigmas,
                                    const float* rgbs, const float* deltas, float* weights_sum, float* depth, float* image,
                                    const int32_t* n_alive_dev, const int32_t* n_step_dev, seald_stream_t stream) {
    if (n_alive == 0 || n_step == 0) return 0;
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth || !image) return SEALD_E_BADARG;
    const uint32_t smem_rows = round_smem_rows(n_step);
    const size_t smem = (size_t)smem_rows * 6 * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_composite_round<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    k_composite_round<false><<<div_up(n_alive, kRoundThreads), kRoundThreads, smem, to_stream(stream)>>>(
        n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, n_alive_dev, n_step_dev, nullptr, nullptr,
        nullptr, 0, 0, 0, smem_rows);
    return launch_status();
}

extern "C" int seald_composite_rays_compact(uint32_t n_alive, uint32_t n_step, float T_thresh, const int32_t* rays_alive, float* rays_t,
                                            const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                                            float* image, int32_t* next_alive, int32_t* state, int32_t* counters2, uint32_t budget,
                                            uint32_t max_steps, uint32_t max_n_step, seald_stream_t stream) {
    if (n_alive == 0 || n_step == 0) return 0;
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth || !image || !next_alive || !state || !counters2)
        return SEALD_E_BADARG;
    if (budget == 0 || max_n_step == 0 || rays_alive == next_alive) return SEALD_E_BADARG;
    const uint32_t smem_rows = round_smem_rows(n_step);
    const size_t smem = (size_t)smem_rows * 6 * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_composite_round<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    // n_alive / n_step are launch bounds; the live values are state[0] / state[1]
    k_composite_round<true><<<div_up(n_alive, kRoundThreads), kRoundThreads, smem, to_stream(stream)>>>(
        n_alive, n_step, T_thresh, const_cast<int32_t*>(rays_alive), rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, state, state + 1,
        next_alive, state, counters2, budget, max_steps, max_n_step, smem_rows);
    return launch_status();
}

// sample-packed round (k_march_round_pack / k_composite_round_pack).  state: int32[8] device = {n_alive, n_step, rays deferred so
// far, steps marched, rows evaluated so far, non-empty rounds, ROWS OF THIS ROUND (the field kernels' live count; zero on entry),
// rays deferred this round (zero on entry)};
// n_alive / n_step arguments are launch bounds.  mapper may be NULL (then mask is ignored).
extern "C" int seald_march_rays_pack(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                                     const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                                     const uint8_t* bitfield, const float* fars, float* xyzs, float* dirs, float* deltas, float* noises,
                                     int32_t* state, uint32_t cap_rows, int32_t* ray_rows, float* stage, const seald_seal_mapper* mapper,
                                     uint8_t* mask, const float* occ_aabb6, const uint32_t* coarse_bits, seald_stream_t stream) {
    if (n_alive == 0 || n_step == 0) return 0;
    if (!rays_alive || !rays_t || !rays_o || !rays_d || !bitfield || !fars || !xyzs || !dirs || !deltas || !state || !ray_rows || !stage)
        return SEALD_E_BADARG;
    if (C == 0 || H == 0 || max_steps == 0 || cap_rows < kRoundThreads * n_step) return SEALD_E_BADARG;
    if (coarse_bits && (C != 1 || H % 32 != 0 || H > 128)) return SEALD_E_UNSUPPORTED;
    if (mapper) {
        if (int rc = check_fusable_mapper(mapper, mask)) return rc;
    }
    static const seald_seal_mapper no_mapper = {};
    const seald_seal_mapper& mp = mapper ? *mapper : no_mapper;
    auto k = mapper ? k_march_round_pack<true> : k_march_round_pack<false>;
    k<<<div_up(n_alive, kRoundThreads), kRoundThreads, 0, to_stream(stream)>>>(
        n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, bitfield, fars, xyzs, dirs, deltas, noises, state,
        state + 1, mp, mask, occ_aabb6, state + 6, cap_rows, reinterpret_cast<int2*>(ray_rows), stage, coarse_bits);
    return launch_status();
}

// Frame prologue of the packed loop (k_render_init): nears / fars / rays_t / zeroed weights_sum, depth, image for all N rays, alive =
// the rays that can produce a sample (any order), state[0..7] initialised with n_step = clamp(budget / n_alive, n_step_min, max_n_step).
// counters2: two zero ints (left zero).  occ_aabb6 may be NULL (then only rays that miss the scene box are dropped).
extern "C" int seald_render_init_pack(const float* rays_o, const float* rays_d, const float* aabb6, const float* occ_aabb6, uint32_t N,
                                      float min_near, float* nears, float* fars, float* rays_t, float* weights_sum, float* depth,
                                      float* image, int32_t* alive, int32_t* state, int32_t* counters2, uint32_t budget,
                                      uint32_t n_step_min, uint32_t max_n_step, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !aabb6 || !nears || !fars || !rays_t || !weights_sum || !depth || !image || !alive || !state || !counters2)
        return SEALD_E_BADARG;
    if (budget == 0 || n_step_min == 0 || max_n_step < n_step_min) return SEALD_E_BADARG;
    k_render_init<<<div_up(N, 256u), 256, 0, to_stream(stream)>>>(rays_o, rays_d, aabb6, occ_aabb6, N, min_near, nears, fars, rays_t, weights_sum,
                                                                  depth, image, alive, state, counters2, budget, n_step_min, max_n_step);
    return launch_status();
}

// coarse occupancy of one cascade level for seald_march_rays_pack: bit j of coarse_bits = any cell of the j-th 8^3 block (Morton order)
// occupied.  H % 32 == 0; bitfield 16-byte aligned; coarse_bits: (H / 8)^3 / 32 words.
extern "C" int seald_occupancy_coarse_bits(const uint8_t* bitfield, uint32_t H, uint32_t* coarse_bits, seald_stream_t stream) {
    if (!bitfield || !coarse_bits || H == 0 || H % 32 != 0) return SEALD_E_BADARG;
    if ((uintptr_t)bitfield % 16) return SEALD_E_ALIGN;
    const uint32_t n_blocks = (H / 8) * (H / 8) * (H / 8);
    k_coarse_bits<<<div_up(n_blocks, 256u), 256, 0, to_stream(stream)>>>(bitfield, n_blocks, coarse_bits);
    return launch_status();
}

// Frame epilogue in one launch (the reference: seven torch ops, dnerf/renderer.py:378-384): image = acc + (1 - weights_sum) * bg,
// depth = clamp(depth - near, 0) / (far - near) (or raw), weights_sum copied out; written to separate tensors and / or as packed rows
// {r, g, b, depth, weights_sum} (the all-gather input of the tile-sharded multi-GPU frame).  Same roundings as the torch expressions
// (no contraction).
__global__ void k_render_finish(const float* __restrict__ image, const float* __restrict__ weights_sum, const float* __restrict__ depth,
                                const float* __restrict__ nears, const float* __restrict__ fars, const uint32_t N, const float bg,
                                const int normalize, float* __restrict__ out_image, float* __restrict__ out_depth,
                                float* __restrict__ out_ws, float* __restrict__ out_packed) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ws = weights_sum[n];
    const float a = __fmul_rn(__fsub_rn(1.0f, ws), bg);
    const float r = __fadd_rn(image[(size_t)n * 3], a), g = __fadd_rn(image[(size_t)n * 3 + 1], a), b = __fadd_rn(image[(size_t)n * 3 + 2], a);
    float d = depth[n];
    if (normalize) d = __fdiv_rn(fmaxf(__fsub_rn(d, nears[n]), 0.0f), __fsub_rn(fars[n], nears[n]));
    if (out_image) { out_image[(size_t)n * 3] = r; out_image[(size_t)n * 3 + 1] = g; out_image[(size_t)n * 3 + 2] = b; }
    if (out_depth) out_depth[n] = d;
    if (out_ws) out_ws[n] = ws;
    if (out_packed) {
        float* o = out_packed + (size_t)n * 5;
        o[0] = r; o[1] = g; o[2] = b; o[3] = d; o[4] = ws;
    }
}

extern "C" int seald_render_finish(const float* image, const float* weights_sum, const float* depth, const float* nears, const float* fars,
                                   uint32_t N, float bg, int normalize_depth, float* out_image, float* out_depth, float* out_ws,
                                   float* out_packed5, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!image || !weights_sum || !depth || (normalize_depth && (!nears || !fars))) return SEALD_E_BADARG;
    if (!out_image && !out_depth && !out_ws && !out_packed5) return SEALD_E_BADARG;
    k_render_finish<<<div_up(N, 256u), 256, 0, to_stream(stream)>>>(image, weights_sum, depth, nears, fars, N, bg, normalize_depth, out_image,
                                                                    out_depth, out_ws, out_packed5);
    return launch_status();
}

extern "C" int seald_composite_rays_pack(uint32_t n_alive, float T_thresh, const int32_t* rays_alive, float* rays_t, const float* sigmas,
                                         const float* rgbs, const float* deltas, float* weights_sum, float* depth, float* image,
                                         int32_t* next_alive, int32_t* state, int32_t* counters2, const int32_t* ray_rows, uint32_t budget,
                                         uint32_t max_steps, uint32_t max_n_step, uint32_t cap_rows, seald_stream_t stream) {
    if (n_alive == 0) return 0;
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth || !image || !next_alive || !state || !counters2 ||
        !ray_rows)
        return SEALD_E_BADARG;
    if (budget == 0 || max_n_step == 0 || rays_alive == next_alive) return SEALD_E_BADARG;
    k_composite_round_pack<<<div_up(n_alive, kRoundThreads), kRoundThreads, 0, to_stream(stream)>>>(
        n_alive, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, next_alive, state, counters2,
        reinterpret_cast<const int2*>(ray_rows), budget, max_steps, max_n_step, cap_rows);
    return launch_status();
}

extern "C" int seald_composite_train_loss_fused(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays, uint32_t M,
                                                uint32_t N, float T_thresh, const float* bg, const float* gt, float inv_count,
                                                const float* loss_scale, float* weights_sum, float* depth, float* image, float* pred,
                                                float* loss_sum, float* grad_sigmas, float* grad_rgbs, seald_stream_t stream) {
    if (N == 0) return 0;
    if (!rays || !gt || !weights_sum || !depth || !image || !loss_sum || !grad_sigmas || !grad_rgbs) return SEALD_E_BADARG;
    if (M > 0 && (!sigmas || !rgbs || !deltas)) return SEALD_E_BADARG;
    if ((uintptr_t)deltas & 7) return SEALD_E_ALIGN;
    k_composite_train_loss_fused<<<div_up(N, kCompWarps), kCompWarps * 32, 0, to_stream(stream)>>>(
        sigmas, rgbs, deltas, rays, M, N, T_thresh, bg, gt, inv_count, loss_scale, weights_sum, depth, image, pred, loss_sum, grad_sigmas,
        grad_rgbs);
    return launch_status();
}

extern "C" int seald_occupancy_aabb(const uint8_t* bitfield, uint32_t C, uint32_t H, float bound, int32_t guard_cells, int32_t* scratch,
                                    float* aabb6, seald_stream_t stream) {
    if (!bitfield || !scratch || !aabb6 || C == 0 || C > 16 || H == 0 || guard_cells < 1) return SEALD_E_BADARG;
    cudaStream_t st = to_stream(stream);
    k_occupancy_init<<<1, 128, 0, st>>>(scratch, C);
    const uint32_t n_bytes = C * H * H * H / 8;
    const uint32_t blocks = min(div_up(n_bytes, 256u), 4u * SEALD_NUM_SMS);
    k_occupancy_ranges<<<blocks, 256, 0, st>>>(bitfield, C, H, scratch);
    k_occupancy_finalize<<<1, 32, 0, st>>>(scratch, C, H, bound, guard_cells, aabb6);
    return launch_status();
}

extern "C" int seald_render_schedule(int32_t* state, const int32_t* n_alive_new, uint32_t N, uint32_t max_steps, uint32_t max_n_step,
                                     seald_stream_t stream) {
    if (!state || !n_alive_new || N == 0 || max_n_step == 0) return SEALD_E_BADARG;
    k_render_schedule<<<1, 32, 0, to_stream(stream)>>>(state, n_alive_new, N, max_steps, max_n_step);
    return launch_status();
}

extern "C" int seald_compact_alive(const int32_t* rays_alive, uint32_t n_alive, const int32_t* n_alive_dev, int32_t* out, int32_t* n_out_dev,
                                   int32_t* scratch, seald_stream_t stream) {
    if (!n_out_dev) return SEALD_E_BADARG;
    cudaStream_t st = to_stream(stream);
    if (n_alive == 0) {
        cudaError_t e = cudaMemsetAsync(n_out_dev, 0, sizeof(int32_t), st);
        return (int)e;
    }
    if (!rays_alive || !out || !scratch) return SEALD_E_BADARG;
    const uint32_t n_tiles = div_up(n_alive, kTile);
    k_compact_count<<<n_tiles, kTile, 0, st>>>(rays_alive, n_alive, n_alive_dev, scratch);
    k_compact_scan<<<1, 1024, 0, st>>>(scratch, n_tiles, n_out_dev);
    k_compact_scatter<<<n_tiles, kTile, 0, st>>>(rays_alive, n_alive, n_alive_dev, scratch, out);
    return launch_status();
}
