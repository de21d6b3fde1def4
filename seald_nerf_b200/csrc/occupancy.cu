// occupancy.cu — device side of the occupancy-grid refresh (NeRFRenderer.update_extra_state, dnerf/renderer.py:453-555).
//
// The reference builds, per time frame, the cell sample points with ~10 torch kernels (meshgrid, cat, morton3D, scale, two
// rand_like passes), queries the density field, scatters into a 512 MiB temporary and does three full-tensor passes for the
// decayed maximum.  Here one kernel produces the jittered sample points + their Morton cell indices, the field kernels run on
// preallocated buffers, one kernel stores the scaled densities into a per-frame temporary and one applies
// grid = max(grid * decay, tmp) where both are >= 0 (dnerf/renderer.py:541-543) and resets the temporary.
// Arithmetic keeps the reference's fp32 operation order (explicit _rn intrinsics: no FMA contraction), so with the same
// uniform random numbers the points are bit-identical to the torch expressions.
#include "common.cuh"

namespace seald {

__device__ __forceinline__ uint32_t occ_expand_bits(uint32_t v) {  // raymarching.cu:56-62
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ uint32_t occ_compact_bits(uint32_t x) {  // inverse of occ_expand_bits (raymarching.cu:71-80)
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xC30C30C3u;
    x = (x | (x >> 4)) & 0x0F00F00Fu;
    x = (x | (x >> 8)) & 0xFF0000FFu;
    x = (x | (x >> 16)) & 0x0000FFFFu;
    return x;
}

// coords == nullptr: the full sweep.  The reference enumerates the cells as j = (x * H + y) * H + z, the order of
// custom_meshgrid(X, Y, Z) (dnerf/renderer.py:481-483), and its rand_like() jitter row j belongs to that cell.  The points are
// PRODUCED x-fastest instead (thread i -> cell x = i % H, y = (i / H) % H, z = i / H^2, jitter still read from row j): the results go
// to the grid through their Morton index, so the order of the batch is free — and 32 consecutive points of a warp then lie on one
// x-line, where the hash grid's rows are contiguous (dense levels: row = x + ..., hashed levels: x ^ const), so the encoder that
// follows shares sectors inside a warp: 25 instead of 72 sector requests per point, the quantity that bounds it
// (profiles/r2_occupancy.md).  coords != nullptr: coords[j] (the random / re-sampled cells of the partial pass, :507-518).
__global__ void k_occ_cell_points(const int* __restrict__ coords, const float* __restrict__ rand3, const uint32_t n, const uint32_t H,
                                  const float span, const float half_cell, float* __restrict__ xyzs, int* __restrict__ indices) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t c[3];
    uint32_t jr = j;  // row of the reference's random numbers that belongs to this point
    if (coords) {
        c[0] = (uint32_t)coords[j * 3]; c[1] = (uint32_t)coords[j * 3 + 1]; c[2] = (uint32_t)coords[j * 3 + 2];
    } else {
        c[0] = j % H; c[1] = (j / H) % H; c[2] = j / (H * H);
        jr = (c[0] * H + c[1]) * H + c[2];
    }
    if (indices) indices[j] = (int)(occ_expand_bits(c[0]) | (occ_expand_bits(c[1]) << 1) | (occ_expand_bits(c[2]) << 2));
    // torch divides a CUDA tensor by a host scalar as a * (1 / b) with the reciprocal rounded to fp32 (BinaryDivTrueKernel.cu)
    const float inv_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
#pragma unroll
    for (int d = 0; d < 3; d++) {
        // xyzs = 2 * coords.float() / (H - 1) - 1;  cas_xyzs = xyzs * (bound - half);  cas_xyzs += (rand * 2 - 1) * half
        const float x = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c[d]), inv_hm1), 1.0f);
        const float jit = __fmul_rn(__fsub_rn(__fmul_rn(__ldg(rand3 + (size_t)jr * 3 + d), 2.0f), 1.0f), half_cell);
        xyzs[(size_t)j * 3 + d] = __fadd_rn(__fmul_rn(x, span), jit);
    }
}

// Partial pass (dnerf/renderer.py:504-518): points [0, n) are the uniformly drawn cells `rand_coords` (torch.randint, int64), points
// [n, 2n) re-sample OCCUPIED cells: the reference takes occ_indices = nonzero(grid > 0)[rand_mask]; here the rand_mask[j]-th occupied
// cell is found by binary search in the inclusive prefix sum `csum` of (grid > 0) — the same cell, without materialising the index
// list — and decoded from its Morton index (morton3D_invert).  Jitter and scaling as in k_occ_cell_points (same fp32 operation order).
__global__ void k_occ_partial_points(const int64_t* __restrict__ rand_coords, const int64_t* __restrict__ rand_mask,
                                     const int* __restrict__ csum, const uint32_t n_cells, const float* __restrict__ rand3, const uint32_t n,
                                     const uint32_t H, const float span, const float half_cell, float* __restrict__ xyzs,
                                     int* __restrict__ indices) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * n) return;
    uint32_t c[3], idx;
    if (j < n) {
        c[0] = (uint32_t)rand_coords[(size_t)j * 3]; c[1] = (uint32_t)rand_coords[(size_t)j * 3 + 1]; c[2] = (uint32_t)rand_coords[(size_t)j * 3 + 2];
        idx = occ_expand_bits(c[0]) | (occ_expand_bits(c[1]) << 1) | (occ_expand_bits(c[2]) << 2);
    } else {
        const int want = (int)rand_mask[j - n] + 1;  // the (k+1)-th occupied cell = first position whose prefix sum reaches k + 1
        uint32_t lo = 0, hi = n_cells - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(csum + mid) >= want) hi = mid; else lo = mid + 1;
        }
        idx = lo;
        c[0] = occ_compact_bits(idx); c[1] = occ_compact_bits(idx >> 1); c[2] = occ_compact_bits(idx >> 2);
    }
    indices[j] = (int)idx;
    const float inv_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float x = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c[d]), inv_hm1), 1.0f);
        const float jit = __fmul_rn(__fsub_rn(__fmul_rn(rand3[(size_t)j * 3 + d], 2.0f), 1.0f), half_cell);
        xyzs[(size_t)j * 3 + d] = __fadd_rn(__fmul_rn(x, span), jit);
    }
}

// tmp[indices[j]] = sigma[j] * density_scale (duplicates: one writer wins, as in `tmp_grid[t, cas, indices] = sigmas`)
__global__ void k_occ_store(const float* __restrict__ sigma, const int* __restrict__ indices, const uint32_t n, const float density_scale,
                            float* __restrict__ tmp) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    tmp[indices[j]] = __fmul_rn(sigma[j], density_scale);
}

// grid = max(grid * decay, tmp) where grid >= 0 and tmp >= 0; tmp reset to -1 for the next frame.  float4 over the frame.
__global__ void k_occ_ema_max(float* __restrict__ grid, float* __restrict__ tmp, const uint32_t n4, const float decay) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 g = reinterpret_cast<float4*>(grid)[i];
    const float4 t = reinterpret_cast<const float4*>(tmp)[i];
    bool any = false;
    if (g.x >= 0 && t.x >= 0) { g.x = fmaxf(__fmul_rn(g.x, decay), t.x); any = true; }
    if (g.y >= 0 && t.y >= 0) { g.y = fmaxf(__fmul_rn(g.y, decay), t.y); any = true; }
    if (g.z >= 0 && t.z >= 0) { g.z = fmaxf(__fmul_rn(g.z, decay), t.z); any = true; }
    if (g.w >= 0 && t.w >= 0) { g.w = fmaxf(__fmul_rn(g.w, decay), t.w); any = true; }
    if (any) reinterpret_cast<float4*>(grid)[i] = g;
    if (t.x != -1.0f || t.y != -1.0f || t.z != -1.0f || t.w != -1.0f) reinterpret_cast<float4*>(tmp)[i] = make_float4(-1.f, -1.f, -1.f, -1.f);
}

}  // namespace seald

using namespace seald;

extern "C" int seald_occ_cell_points(const int32_t* coords, const float* rand3, uint32_t n, uint32_t H, float span, float half_cell,
                                     float* xyzs, int32_t* indices, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!rand3 || !xyzs || H < 2 || H > 1024) return SEALD_E_BADARG;
    k_occ_cell_points<<<div_up(n, 256u), 256, 0, to_stream(stream)>>>(coords, rand3, n, H, span, half_cell, xyzs, indices);
    return launch_status();
}

extern "C" int seald_occ_partial_points(const int64_t* rand_coords, const int64_t* rand_mask, const int32_t* csum, uint32_t n_cells,
                                        const float* rand3, uint32_t n, uint32_t H, float span, float half_cell, float* xyzs,
                                        int32_t* indices, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!rand_coords || !rand_mask || !csum || !rand3 || !xyzs || !indices || H < 2 || H > 1024 || n_cells == 0) return SEALD_E_BADARG;
    k_occ_partial_points<<<div_up(2 * n, 256u), 256, 0, to_stream(stream)>>>(rand_coords, rand_mask, csum, n_cells, rand3, n, H, span, half_cell,
                                                                            xyzs, indices);
    return launch_status();
}

extern "C" int seald_occ_store(const float* sigma, const int32_t* indices, uint32_t n, float density_scale, float* tmp, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!sigma || !indices || !tmp) return SEALD_E_BADARG;
    k_occ_store<<<div_up(n, 256u), 256, 0, to_stream(stream)>>>(sigma, indices, n, density_scale, tmp);
    return launch_status();
}

extern "C" int seald_occ_ema_max(float* grid, float* tmp, uint32_t n, float decay, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!grid || !tmp || n % 4 || ((uintptr_t)grid | (uintptr_t)tmp) % 16) return SEALD_E_BADARG;
    k_occ_ema_max<<<div_up(n / 4, 256u), 256, 0, to_stream(stream)>>>(grid, tmp, n / 4, decay);
    return launch_status();
}
