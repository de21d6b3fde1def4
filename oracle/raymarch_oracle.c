/*
 * raymarch_oracle.c — CPU restatement of the reference's ray-marching / compositing kernels.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker; never by the product path.
 *
 * Each function follows one kernel of /root/reference/raymarching/src/raymarching.cu (cited per function).
 * GPU arithmetic is reproduced operation by operation: where nvcc contracts a multiply-add into an FMA
 * (checked against the SASS of the reference built for sm_100a, oracle/_ref/_ref_raymarching.so) fmaf() is
 * written explicitly; everything else is plain IEEE fp32, so build with -ffp-contract=off.
 * Pinning: tests/golden/raymarch_*.npz hold outputs of the reference extension itself run on a B200
 * (tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file against them.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float signf(float x) { return copysignf(1.0f, x); }

/* raymarching.cu:56-81 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

void oracle_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) { /* raymarching.cu:214-226 */
    for (uint32_t n = 0; n < N; n++) indices[n] = (int32_t)morton3D(coords[3 * n], coords[3 * n + 1], coords[3 * n + 2]);
}
void oracle_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) { /* raymarching.cu:237-254 */
    for (uint32_t n = 0; n < N; n++) {
        const int32_t ind = indices[n];
        coords[3 * n] = (int32_t)morton3D_invert((uint32_t)(ind >> 0));
        coords[3 * n + 1] = (int32_t)morton3D_invert((uint32_t)(ind >> 1));
        coords[3 * n + 2] = (int32_t)morton3D_invert((uint32_t)(ind >> 2));
    }
}
void oracle_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield) { /* raymarching.cu:268-289 */
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* raymarching.cu:92-145 ((a - o) * rd is a separate subtract and multiply on the GPU too) */
void oracle_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                               float* nears, float* fars) {
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
        const float dx = rays_d[3 * n], dy = rays_d[3 * n + 1], dz = rays_d[3 * n + 2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, t;
        if (near > far) { t = near; near = far; far = t; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { t = near_y; near_y = far_y; far_y = t; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { t = near_z; near_z = far_z; far_z = t; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:42-54 */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, H3;
    uint32_t C, H;
    const uint8_t* grid;
} march_t;

typedef struct { float x, y, z, dt, mip_bound; int nx, ny, nz; } probe_t;

static void march_init(march_t* m, const float* o, const float* d, const uint8_t* grid, float bound, float dt_gamma,
                       uint32_t max_steps, uint32_t C, uint32_t H) {
    m->ox = o[0]; m->oy = o[1]; m->oz = o[2];
    m->dx = d[0]; m->dy = d[1]; m->dz = d[2];
    m->rdx = 1 / m->dx; m->rdy = 1 / m->dy; m->rdz = 1 / m->dz;
    m->bound = bound; m->dt_gamma = dt_gamma;
    m->rH = 1 / (float)H;
    m->H3 = (float)(H * H * H);
    m->dt_min = 2 * 1.7320508075688772f / (float)max_steps;                           /* raymarching.cu:345 */
    m->dt_max = 2 * 1.7320508075688772f * (float)(1 << (C - 1)) / (float)H;           /* raymarching.cu:346 */
    m->C = C; m->H = H; m->grid = grid;
}

/* raymarching.cu:360-379; FMAs: x = fma(d,t,o); (x*rb+1) = fma; index = fma(level,H3,morton) */
static int probe(const march_t* m, float t, probe_t* p) {
    const float x = clampf(fmaf(t, m->dx, m->ox), -m->bound, m->bound);
    const float y = clampf(fmaf(t, m->dy, m->oy), -m->bound, m->bound);
    const float z = clampf(fmaf(t, m->dz, m->oz), -m->bound, m->bound);
    const float dt = clampf(t * m->dt_gamma, m->dt_min, m->dt_max);
    const int a = mip_from_pos(x, y, z, (float)m->C), b = mip_from_dt(dt, (float)m->H, (float)m->C);
    const int level = a > b ? a : b;
    const float mip_bound = fminf(scalbnf(1.0f, level), m->bound);
    const float mip_rbound = 1 / mip_bound;
    const double Hd = (double)m->H;
    const float Hm1 = (float)(m->H - 1);
    const int nx = (int)clampf((float)(0.5 * (double)fmaf(x, mip_rbound, 1.0f) * Hd), 0.0f, Hm1);
    const int ny = (int)clampf((float)(0.5 * (double)fmaf(y, mip_rbound, 1.0f) * Hd), 0.0f, Hm1);
    const int nz = (int)clampf((float)(0.5 * (double)fmaf(z, mip_rbound, 1.0f) * Hd), 0.0f, Hm1);
    const uint32_t index = (uint32_t)fmaf((float)level, m->H3, (float)morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    p->x = x; p->y = y; p->z = z; p->dt = dt; p->mip_bound = mip_bound; p->nx = nx; p->ny = ny; p->nz = nz;
    return (m->grid[index / 8] & (1 << (index % 8))) != 0;
}

/* raymarching.cu:388-399; FMAs as in the SASS: fma(sign,0.5,n+0.5); *rH; fma(.,2,-1); fma(mip_bound,.,-x); *rd */
static float skip_voxel(const march_t* m, const probe_t* p, float t) {
    float sx = fmaf(signf(m->dx), 0.5f, (float)p->nx + 0.5f) * m->rH;
    float sy = fmaf(signf(m->dy), 0.5f, (float)p->ny + 0.5f) * m->rH;
    float sz = fmaf(signf(m->dz), 0.5f, (float)p->nz + 0.5f) * m->rH;
    const float tx = fmaf(p->mip_bound, fmaf(sx, 2.0f, -1.0f), -p->x) * m->rdx;
    const float ty = fmaf(p->mip_bound, fmaf(sy, 2.0f, -1.0f), -p->y) * m->rdy;
    const float tz = fmaf(p->mip_bound, fmaf(sz, 2.0f, -1.0f), -p->z) * m->rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
        t += clampf(t * m->dt_gamma, m->dt_min, m->dt_max);
    } while (t < tt);
    return t;
}

/* raymarching.cu:312-480.  Deterministic variant of the output-range reservation: rays are packed in ray order
 * (the reference's atomicAdd order is arbitrary; counts and per-ray samples are what is compared).
 * rays row n = (n, offset, count).  counter[0] += samples, counter[1] += N. */
void oracle_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                             const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                             const float* noises) {
    for (uint32_t n = 0; n < N; n++) {
        march_t m;
        march_init(&m, rays_o + 3 * n, rays_d + 3 * n, grid, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n], noise = noises[n];
        const float t0 = fmaf(clampf(near * dt_gamma, m.dt_min, m.dt_max), noise, near); /* :351, FFMA in SASS */
        float t = t0;
        uint32_t num_steps = 0;
        probe_t p;
        while (t < far && num_steps < max_steps) {
            if (probe(&m, t, &p)) { num_steps++; t += p.dt; }
            else t = skip_voxel(&m, &p, t);
        }
        const uint32_t point_index = (uint32_t)counter[0];
        counter[0] += (int32_t)num_steps;
        counter[1] += 1;
        rays[3 * n] = (int32_t)n; rays[3 * n + 1] = (int32_t)point_index; rays[3 * n + 2] = (int32_t)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps > M) continue;
        float* xo = xyzs + (size_t)point_index * 3;
        float* dq = dirs + (size_t)point_index * 3;
        float* de = deltas + (size_t)point_index * 2;
        t = t0;
        uint32_t step = 0;
        float last_t = t;
        while (t < far && step < num_steps) {
            if (probe(&m, t, &p)) {
                xo[0] = p.x; xo[1] = p.y; xo[2] = p.z;
                dq[0] = m.dx; dq[1] = m.dy; dq[2] = m.dz;
                t += p.dt;
                de[0] = p.dt; de[1] = t - last_t;
                last_t = t;
                xo += 3; dq += 3; de += 2; step++;
            } else t = skip_voxel(&m, &p, t);
        }
    }
}

/* raymarching.cu:701-805 */
void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                       const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                       const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                       const float* noises) {
    (void)nears;
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float noise = noises[n];
        march_t m;
        march_init(&m, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, grid, bound, dt_gamma, max_steps, C, H);
        float* xo = xyzs + (size_t)n * n_step * 3;
        float* dq = dirs + (size_t)n * n_step * 3;
        float* de = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        const float far = fars[index];
        uint32_t step = 0;
        t = fmaf(clampf(t * dt_gamma, m.dt_min, m.dt_max), noise, t);
        float last_t = t;
        probe_t p;
        while (t < far && step < n_step) {
            if (probe(&m, t, &p)) {
                xo[0] = p.x; xo[1] = p.y; xo[2] = p.z;
                dq[0] = m.dx; dq[1] = m.dy; dq[2] = m.dz;
                t += p.dt;
                de[0] = p.dt; de[1] = t - last_t;
                last_t = t;
                xo += 3; dq += 3; de += 2; step++;
            } else t = skip_voxel(&m, &p, t);
        }
    }
}

/* GPU __expf(x) = ex2.approx(x * log2(e)); on the CPU we use exp2f of the same fp32 product (<= 2 ulp apart). */
static inline float gpu_expf(float x) { return exp2f(x * 1.4426950408889634f); }

/* raymarching.cu:501-577 (accumulations are FFMAs on the GPU) */
void oracle_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                         uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num_steps = (uint32_t)rays[3 * n + 2];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        const float* s = sigmas + offset; const float* c = rgbs + (size_t)offset * 3; const float* dl = deltas + (size_t)offset * 2;
        uint32_t step = 0;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        while (step < num_steps) {
            const float alpha = 1.0f - gpu_expf(-s[0] * dl[0]);
            const float weight = alpha * T;
            r = fmaf(weight, c[0], r); g = fmaf(weight, c[1], g); b = fmaf(weight, c[2], b);
            t += dl[1];
            d = fmaf(weight, t, d);
            ws += weight;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; step++;
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* raymarching.cu:602-682 */
void oracle_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                          const float* rgbs, const float* deltas, const int32_t* rays, const float* weights_sum,
                                          const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                          float* grad_rgbs) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num_steps = (uint32_t)rays[3 * n + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float gws = grad_weights_sum[index];
        const float* gi = grad_image + (size_t)index * 3;
        const float r_final = image[index * 3], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2], ws_final = weights_sum[index];
        const float* s = sigmas + offset; const float* c = rgbs + (size_t)offset * 3; const float* dl = deltas + (size_t)offset * 2;
        float* gs = grad_sigmas + offset; float* gc = grad_rgbs + (size_t)offset * 3;
        uint32_t step = 0;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        while (step < num_steps) {
            const float alpha = 1.0f - gpu_expf(-s[0] * dl[0]);
            const float weight = alpha * T;
            r = fmaf(weight, c[0], r); g = fmaf(weight, c[1], g); b = fmaf(weight, c[2], b);
            ws += weight;
            T *= 1.0f - alpha;
            gc[0] = gi[0] * weight; gc[1] = gi[1] * weight; gc[2] = gi[2] * weight;
            gs[0] = dl[0] * (gi[0] * (T * c[0] - (r_final - r)) + gi[1] * (T * c[1] - (g_final - g)) +
                             gi[2] * (T * c[2] - (b_final - b)) + gws * (1 - ws_final));
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; gs++; gc += 3; step++;
        }
    }
}

/* raymarching.cu:819-905 */
void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t, const float* sigmas,
                           const float* rgbs, const float* deltas, float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float* s = sigmas + (size_t)n * n_step; const float* c = rgbs + (size_t)n * n_step * 3; const float* dl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index], weight_sum = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - gpu_expf(-s[0] * dl[0]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t += dl[1];
            d = fmaf(weight, t, d);
            r = fmaf(weight, c[0], r); g = fmaf(weight, c[1], g); b = fmaf(weight, c[2], b);
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = weight_sum; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}
