// seal.cuh — device side of the Seal editing proxy mapping (SealNeRF/seal_utils.py runtime).
//
// One sample at a time, no shared memory, no synchronisation: the mapper's triangles / bounds / border points are a
// few hundred bytes to a few KB and stay hot in L1 through the read-only path, so `seal_map_sample` can be called from
// inside any march kernel right before a sample is stored ("remapped before encoding").  Expression order follows the
// reference's torch code (einsum / cross products in fp32) so that the inside/outside decision is identical away from
// triangle edges.
#pragma once
#include "common.cuh"

namespace seald {

// Moller-Trumbore, one origin against all F triangles, for the direction d AND -d at once (seal_utils.py:638-693):
// inside <=> (some triangle is hit along +d) && (some triangle is hit along -d).
__device__ __forceinline__ bool seal_point_in_mesh(const float* __restrict__ tris, const int F, const float px, const float py, const float pz,
                                                   const float dx, const float dy, const float dz) {
    constexpr float eps = 1e-8f;
    bool hit_p = false, hit_n = false;
    for (int f = 0; f < F; f++) {
        const float* t = tris + (size_t)f * 9;
        const float v0x = __ldg(t + 0), v0y = __ldg(t + 1), v0z = __ldg(t + 2);
        const float e1x = __ldg(t + 3) - v0x, e1y = __ldg(t + 4) - v0y, e1z = __ldg(t + 5) - v0z;
        const float e2x = __ldg(t + 6) - v0x, e2y = __ldg(t + 7) - v0y, e2z = __ldg(t + 8) - v0z;
        // N = E1 x E2
        const float nx = e1y * e2z - e1z * e2y;
        const float ny = e1z * e2x - e1x * e2z;
        const float nz = e1x * e2y - e1y * e2x;
        const float a0x = px - v0x, a0y = py - v0y, a0z = pz - v0z;
        const float dn = dx * nx + dy * ny + dz * nz;
        const float a0n = a0x * nx + a0y * ny + a0z * nz;
        // DA0 = A0 x d
        const float cx = a0y * dz - a0z * dy;
        const float cy = a0z * dx - a0x * dz;
        const float cz = a0x * dy - a0y * dx;
        const float de2 = cx * e2x + cy * e2y + cz * e2z;
        const float de1 = cx * e1x + cy * e1y + cz * e1z;
        if (!hit_p) {
            const float inv = 1.0f / -(dn + eps);
            const float u = de2 * inv, v = -de1 * inv, tt = a0n * inv;
            hit_p = (tt >= 0.0f) && (u >= 0.0f) && (v >= 0.0f) && ((u + v) <= 1.0f);
        }
        if (!hit_n) {
            // direction -d: d.N and A0 x d change sign
            const float inv = 1.0f / -(-dn + eps);
            const float u = -de2 * inv, v = de1 * inv, tt = a0n * inv;
            hit_n = (tt >= 0.0f) && (u >= 0.0f) && (v >= 0.0f) && ((u + v) <= 1.0f);
        }
        if (hit_p && hit_n) return true;
    }
    return false;
}

// SealMapper.map_mask (seal_utils.py:132-153): inside one of the AABBs (strictly), no zero coordinate, inside the mesh.
__device__ __forceinline__ bool seal_map_mask(const seald_seal_mapper& mp, const float x, const float y, const float z) {
    if (x == 0.0f || y == 0.0f || z == 0.0f) return false;  // `points.all(1)`
    bool in_bound = false;
    for (int i = 0; i < mp.n_bounds && !in_bound; i++) {
        const float* b = mp.bounds + i * 6;
        in_bound = (__ldg(b + 3) > x) && (x > __ldg(b + 0)) && (__ldg(b + 4) > y) && (y > __ldg(b + 1)) && (__ldg(b + 5) > z) && (z > __ldg(b + 2));
    }
    if (!in_bound) return false;
    return seal_point_in_mesh(mp.tris, mp.n_tris, x, y, z, mp.test_dir[0], mp.test_dir[1], mp.test_dir[2]);
}

// project_points (seal_utils.py:736-744)
__device__ __forceinline__ void seal_project(const float* n, const float* p0, const float x, const float y, const float z, float& ox, float& oy,
                                             float& oz) {
    const float vx = x - p0[0], vy = y - p0[1], vz = z - p0[2];
    const float s = (vx * n[0] + vy * n[1] + vz * n[2]) / (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    ox = x - s * n[0];
    oy = y - s * n[1];
    oz = z - s * n[2];
}

// bbox (seal_utils.py:244-286) and brush (:415-461) map_to_origin for ONE sample; returns the map mask.
// x,y,z / dx,dy,dz are updated in place.  (The anchor mapper needs a batch-wide "any" and lives in seal.cu.)
__device__ __forceinline__ bool seal_map_sample(const seald_seal_mapper& mp, float& x, float& y, float& z, float& dx, float& dy, float& dz) {
    const bool m = seal_map_mask(mp, x, y, z);
    if (mp.type == SEALD_SEAL_BBOX) {
        if (m) {
            const float* T = mp.transform;
            const float tx = T[0] * x + T[1] * y + T[2] * z + T[3];
            const float ty = T[4] * x + T[5] * y + T[6] * z + T[7];
            const float tz = T[8] * x + T[9] * y + T[10] * z + T[11];
            const float* R = mp.rotation;
            const float rx = R[0] * dx + R[1] * dy + R[2] * dz;
            const float ry = R[3] * dx + R[4] * dy + R[5] * dz;
            const float rz = R[6] * dx + R[7] * dy + R[8] * dz;
            x = (tx - mp.center[0]) * mp.scale[0] + mp.center[0];
            y = (ty - mp.center[1]) * mp.scale[1] + mp.center[1];
            z = (tz - mp.center[2]) * mp.scale[2] + mp.center[2];
            dx = rx; dy = ry; dz = rz;
        } else if (mp.has_map_source) {
            const float* e = mp.empty_bound;
            if (e[3] > x && x > e[0] && e[4] > y && y > e[1] && e[5] > z && z > e[2]) {
                x = mp.map_source[0]; y = mp.map_source[1]; z = mp.map_source[2];
            }
        }
    } else if (mp.type == SEALD_SEAL_BRUSH) {
        if (m && mp.attenuation_mode == SEALD_SEAL_ATT_LINEAR) {
            float qx, qy, qz;
            seal_project(mp.normal_expand, mp.center, x, y, z, qx, qy, qz);
            float best = 3.4e38f;
            for (int i = 0; i < mp.n_border; i++) {
                const float* b = mp.border + (size_t)i * 3;
                const float ex = qx - __ldg(b), ey = qy - __ldg(b + 1), ez = qz - __ldg(b + 2);
                best = fminf(best, ex * ex + ey * ey + ez * ez);
            }
            const float dist = sqrtf(best);
            x -= mp.normal_expand[0]; y -= mp.normal_expand[1]; z -= mp.normal_expand[2];
            const float ad = mp.attenuation_distance;
            if (ad > dist) {
                const float k = fabsf(ad - dist) / ad;
                x += k * mp.normal_expand[0]; y += k * mp.normal_expand[1]; z += k * mp.normal_expand[2];
            }
        }
    }
    return m;
}

}  // namespace seald
