// seal.cu — Seal editing proxy mapping as stand-alone ops: map_to_origin for the bbox / brush / anchor mappers and
// map_color (HSV shift, RGB recolour with the batch-mean brightness, image decal).  The fused march variants live in
// raymarch.cu and share seal.cuh.  Reference: SealNeRF/seal_utils.py:48-81,132-153,244-286,415-461,522-578,736-777 and
// SealNeRF/color_utils.py:31-63.
#include "seal.cuh"

namespace seald {

__global__ void k_seal_map(const __grid_constant__ seald_seal_mapper mp, const float* __restrict__ points, const float* __restrict__ dirs,
                           uint32_t M, const int* __restrict__ m_dev, float* __restrict__ points_out, float* __restrict__ dirs_out,
                           uint8_t* __restrict__ mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float x = points[(size_t)i * 3], y = points[(size_t)i * 3 + 1], z = points[(size_t)i * 3 + 2];
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (dirs) { dx = dirs[(size_t)i * 3]; dy = dirs[(size_t)i * 3 + 1]; dz = dirs[(size_t)i * 3 + 2]; }
    bool m = false;
    if (!m_dev || i < (uint32_t)max(*m_dev, 0)) m = seal_map_sample(mp, x, y, z, dx, dy, dz);
    points_out[(size_t)i * 3] = x; points_out[(size_t)i * 3 + 1] = y; points_out[(size_t)i * 3 + 2] = z;
    if (dirs_out) { dirs_out[(size_t)i * 3] = dx; dirs_out[(size_t)i * 3 + 1] = dy; dirs_out[(size_t)i * 3 + 2] = dz; }
    mask[i] = m ? 1 : 0;
}

// anchor pass 1: does ANY point fall inside the mapper's mesh (the reference's early exit, seal_utils.py:526-528)
__global__ void k_seal_any(const __grid_constant__ seald_seal_mapper mp, const float* __restrict__ points, uint32_t M,
                           const int* __restrict__ m_dev, int* __restrict__ any_flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool m = false;
    if (i < M && (!m_dev || i < (uint32_t)max(*m_dev, 0)))
        m = seal_map_mask(mp, points[(size_t)i * 3], points[(size_t)i * 3 + 1], points[(size_t)i * 3 + 2]);
    if (__syncthreads_or(m) && threadIdx.x == 0) atomicOr(any_flag, 1);
}

// anchor pass 2 (seal_utils.py:530-578): cone filter + pull towards the anchor plane, over ALL points
__global__ void k_seal_anchor(const __grid_constant__ seald_seal_mapper mp, const float* __restrict__ points, const float* __restrict__ dirs,
                              uint32_t M, const int* __restrict__ m_dev, const int* __restrict__ any_flag, float* __restrict__ points_out,
                              float* __restrict__ dirs_out, uint8_t* __restrict__ mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float x = points[(size_t)i * 3], y = points[(size_t)i * 3 + 1], z = points[(size_t)i * 3 + 2];
    bool valid = false;
    if (*any_flag && (!m_dev || i < (uint32_t)max(*m_dev, 0))) {
        float qx, qy, qz;
        seal_project(mp.v_h, mp.v_anchor, x, y, z, qx, qy, qz);
        const float wx = qx - x, wy = qy - y, wz = qz - z;  // v_points_to_plane
        const float dist = sqrtf(wx * wx + wy * wy + wz * wz);
        const float os = dist / mp.len_h;
        const float px = qx - os * mp.v_offset[0], py = qy - os * mp.v_offset[1], pz = qz - os * mp.v_offset[2];
        const float ax = px - mp.v_anchor[0], ay = py - mp.v_anchor[1], az = pz - mp.v_anchor[2];
        const float pop = sqrtf(ax * ax + ay * ay + az * az);
        const bool cone = (pop <= mp.radius) && (dist / (mp.radius - pop) < mp.len_h / mp.radius * 1.1f);
        const bool side = (wx * mp.v_h[0] + wy * mp.v_h[1] + wz * mp.v_h[2]) > 0.0f;
        valid = cone && side;
        if (valid) {
            const float k = -((mp.len_h - dist) / 10.0f);
            const float mx = px - k * mp.v_h[0] / mp.len_h, my = py - k * mp.v_h[1] / mp.len_h, mz = pz - k * mp.v_h[2] / mp.len_h;
            x = (mx - mp.v_anchor[0]) * mp.scale[0] + mp.v_anchor[0];
            y = (my - mp.v_anchor[1]) * mp.scale[1] + mp.v_anchor[1];
            z = (mz - mp.v_anchor[2]) * mp.scale[2] + mp.v_anchor[2];
        }
    }
    points_out[(size_t)i * 3] = x; points_out[(size_t)i * 3 + 1] = y; points_out[(size_t)i * 3 + 2] = z;
    if (dirs_out && dirs_out != dirs) {
        dirs_out[(size_t)i * 3] = dirs[(size_t)i * 3]; dirs_out[(size_t)i * 3 + 1] = dirs[(size_t)i * 3 + 1];
        dirs_out[(size_t)i * 3 + 2] = dirs[(size_t)i * 3 + 2];
    }
    mask[i] = valid ? 1 : 0;
}

// ---- colour (color_utils.py:31-63) --------------------------------------------------------------------------------------
__device__ __forceinline__ float pymod(const float a, const float b) {  // torch `%`: result takes the sign of the divisor
    float r = fmodf(a, b);
    if (r != 0.0f && ((r < 0.0f) != (b < 0.0f))) r += b;
    return r;
}

__device__ __forceinline__ void rgb2hsv(const float r, const float g, const float b, float& h, float& s, float& v) {
    const float cmax = fmaxf(r, fmaxf(g, b)), cmin = fminf(r, fminf(g, b));
    const float delta = cmax - cmin;
    if (delta == 0.0f) h = 0.0f;
    else if (r >= g && r >= b) h = pymod((g - b) / delta, 6.0f);   // first maximum wins ties, like torch.max
    else if (g >= b) h = (b - r) / delta + 2.0f;
    else h = (r - g) / delta + 4.0f;
    h /= 6.0f;
    s = (cmax == 0.0f) ? 0.0f : delta / cmax;
    v = cmax;
}

__device__ __forceinline__ void hsv2rgb(const float h, const float s, const float v, float& r, float& g, float& b) {
    const float c = v * s;
    const float x = c * (-fabsf(pymod(h * 6.0f, 2.0f) - 1.0f) + 1.0f);
    const float m = v - c;
    const int idx = ((int)(h * 6.0f) & 255) % 6;  // .type(torch.uint8) % 6
    float rr, gg, bb;
    switch (idx) {
        case 0: rr = c; gg = x; bb = 0; break;
        case 1: rr = x; gg = c; bb = 0; break;
        case 2: rr = 0; gg = c; bb = x; break;
        case 3: rr = 0; gg = x; bb = c; break;
        case 4: rr = x; gg = 0; bb = c; break;
        default: rr = c; gg = 0; bb = x; break;
    }
    r = rr + m; g = gg + m; b = bb + m;
}

__device__ __forceinline__ void apply_hsv(const seald_seal_color& cm, float& r, float& g, float& b) {
    float h, s, v;
    rgb2hsv(r, g, b, h, s, v);
    hsv2rgb(h + cm.hsv[0], s + cm.hsv[1], v + cm.hsv[2], r, g, b);
}

// modify_rgb (seal_utils.py:761-777): hue/saturation of the target colour, brightness = target V + (own V - batch mean V) + offset
__device__ __forceinline__ void apply_rgb(const float tr, const float tg, const float tb, const float mean_v, const float light_offset, float& r,
                                          float& g, float& b) {
    float h, s, v, th, ts, tv;
    rgb2hsv(r, g, b, h, s, v);
    rgb2hsv(tr, tg, tb, th, ts, tv);
    const float nv = fminf(1.0f, fmaxf(0.0f, tv + (v - mean_v) + light_offset));
    hsv2rgb(th, ts, nv, r, g, b);
}

__device__ __forceinline__ void block_sum2_atomic(float a, float b, float* out) {
    __shared__ float s_a[32], s_b[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (lane == 0) { s_a[w] = a; s_b[w] = b; }
    __syncthreads();
    if (w == 0) {
        const uint32_t nw = (blockDim.x + 31) >> 5;
        a = lane < nw ? s_a[lane] : 0.f;
        b = lane < nw ? s_b[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0 && b != 0.f) { atomicAdd(out, a); atomicAdd(out + 1, b); }
    }
}

// stage 0: statistics for the rgb recolour (V after the optional hsv shift); stage 1: statistics for the image decal
// (V of the colours as stored, i.e. after stage-0 apply).
template <int STAGE>
__global__ void k_seal_color_stats(const __grid_constant__ seald_seal_color cm, const uint8_t* __restrict__ mask, const float* __restrict__ rgbs,
                                   uint32_t M, const int* __restrict__ m_dev, float* __restrict__ scratch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float sv = 0.f, cnt = 0.f;
    if (i < M && (!m_dev || i < (uint32_t)max(*m_dev, 0)) && mask[i]) {
        float r = rgbs[(size_t)i * 3], g = rgbs[(size_t)i * 3 + 1], b = rgbs[(size_t)i * 3 + 2];
        if (STAGE == 0 && cm.has_hsv) apply_hsv(cm, r, g, b);
        sv = fmaxf(r, fmaxf(g, b));
        cnt = 1.f;
    }
    block_sum2_atomic(sv, cnt, scratch + 2 * STAGE);
}

template <int STAGE>
__global__ void k_seal_color_apply(const __grid_constant__ seald_seal_color cm, const float* __restrict__ points, const uint8_t* __restrict__ mask,
                                   float* __restrict__ rgbs, uint32_t M, const int* __restrict__ m_dev, const float* __restrict__ scratch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || (m_dev && i >= (uint32_t)max(*m_dev, 0)) || !mask[i]) return;
    float r = rgbs[(size_t)i * 3], g = rgbs[(size_t)i * 3 + 1], b = rgbs[(size_t)i * 3 + 2];
    if (STAGE == 0) {
        if (cm.has_hsv) apply_hsv(cm, r, g, b);
        if (cm.has_rgb) apply_rgb(cm.rgb[0], cm.rgb[1], cm.rgb[2], scratch[0] / scratch[1], cm.rgb_light_offset, r, g, b);
    } else {
        // image decal (seal_utils.py:58-79): project onto the image plane, nearest texel, alpha blend
        float qx, qy, qz;
        seal_project(cm.v_norm, cm.v_o, points[(size_t)i * 3], points[(size_t)i * 3 + 1], points[(size_t)i * 3 + 2], qx, qy, qz);
        const float px = qx - cm.v_o[0], py = qy - cm.v_o[1], pz = qz - cm.v_o[2];
        const float wx = cm.v_w[0] - cm.v_o[0], wy = cm.v_w[1] - cm.v_o[1], wz = cm.v_w[2] - cm.v_o[2];
        const float hx = cm.v_h[0] - cm.v_o[0], hy = cm.v_h[1] - cm.v_o[1], hz = cm.v_h[2] - cm.v_o[2];
        const float lw = sqrtf(wx * wx + wy * wy + wz * wz), lh = sqrtf(hx * hx + hy * hy + hz * hz);
        const float fw = floorf((px * wx + py * wy + pz * wz) / (lw * lw) * cm.img_w);
        const float fh = floorf((px * hx + py * hy + pz * hz) / (lh * lh) * cm.img_h);
        const int iw = (int)fminf(fmaxf(0.0f, fw), (float)(cm.img_w - 1));
        const int ih = (int)fminf(fmaxf(0.0f, fh), (float)(cm.img_h - 1));
        const float* texel = cm.image + ((size_t)ih * cm.img_w + iw) * 3;
        const float a = __ldg(cm.image_mask + (size_t)ih * cm.img_w + iw);
        float mr = r, mg = g, mb = b;
        apply_rgb(__ldg(texel), __ldg(texel + 1), __ldg(texel + 2), scratch[2] / scratch[3], cm.rgb_light_offset, mr, mg, mb);
        r = a * mr + (1.0f - a) * r;
        g = a * mg + (1.0f - a) * g;
        b = a * mb + (1.0f - a) * b;
    }
    rgbs[(size_t)i * 3] = r; rgbs[(size_t)i * 3 + 1] = g; rgbs[(size_t)i * 3 + 2] = b;
}

}  // namespace seald

using namespace seald;

static int check_mapper(const seald_seal_mapper* mp) {
    if (!mp) return SEALD_E_BADARG;
    if (mp->type < SEALD_SEAL_BBOX || mp->type > SEALD_SEAL_ANCHOR) return SEALD_E_UNSUPPORTED;
    if (mp->n_bounds <= 0 || mp->n_tris <= 0 || !mp->bounds || !mp->tris) return SEALD_E_BADARG;
    if (mp->type == SEALD_SEAL_BRUSH) {
        if (mp->attenuation_mode != SEALD_SEAL_ATT_LINEAR && mp->attenuation_mode != SEALD_SEAL_ATT_DRY) return SEALD_E_UNSUPPORTED;
        if (mp->attenuation_mode == SEALD_SEAL_ATT_LINEAR && (mp->n_border <= 0 || !mp->border)) return SEALD_E_BADARG;
    }
    return 0;
}

extern "C" int seald_seal_map_to_origin(const seald_seal_mapper* mapper, const float* points, const float* dirs, uint32_t M,
                                        const int32_t* m_dev, float* points_out, float* dirs_out, uint8_t* mask, int32_t* scratch,
                                        seald_stream_t stream) {
    if (int rc = check_mapper(mapper)) return rc;
    if (M == 0) return 0;
    if (!points || !points_out || !mask) return SEALD_E_BADARG;
    if (dirs_out && !dirs) return SEALD_E_BADARG;
    cudaStream_t st = to_stream(stream);
    const uint32_t threads = 128, blocks = div_up(M, threads);
    if (mapper->type == SEALD_SEAL_ANCHOR) {
        if (!scratch) return SEALD_E_BADARG;
        cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(int32_t), st);
        if (e != cudaSuccess) return (int)e;
        k_seal_any<<<blocks, threads, 0, st>>>(*mapper, points, M, m_dev, scratch);
        k_seal_anchor<<<blocks, threads, 0, st>>>(*mapper, points, dirs, M, m_dev, scratch, points_out, dirs_out, mask);
    } else {
        k_seal_map<<<blocks, threads, 0, st>>>(*mapper, points, dirs, M, m_dev, points_out, dirs_out, mask);
    }
    return launch_status();
}

extern "C" int seald_seal_map_color(const seald_seal_color* color, const float* points, const uint8_t* mask, float* rgbs, uint32_t M,
                                    const int32_t* m_dev, float* scratch, seald_stream_t stream) {
    if (!color) return SEALD_E_BADARG;
    if (M == 0 || !(color->has_hsv || color->has_rgb || color->has_image)) return 0;
    if (!mask || !rgbs || !scratch) return SEALD_E_BADARG;
    if (color->has_image && (!points || !color->image || !color->image_mask || color->img_h <= 0 || color->img_w <= 0)) return SEALD_E_BADARG;
    cudaStream_t st = to_stream(stream);
    const uint32_t threads = 256, blocks = div_up(M, threads);
    cudaError_t e = cudaMemsetAsync(scratch, 0, 4 * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    if (color->has_hsv || color->has_rgb) {
        if (color->has_rgb) k_seal_color_stats<0><<<blocks, threads, 0, st>>>(*color, mask, rgbs, M, m_dev, scratch);
        k_seal_color_apply<0><<<blocks, threads, 0, st>>>(*color, points, mask, rgbs, M, m_dev, scratch);
    }
    if (color->has_image) {
        k_seal_color_stats<1><<<blocks, threads, 0, st>>>(*color, mask, rgbs, M, m_dev, scratch);
        k_seal_color_apply<1><<<blocks, threads, 0, st>>>(*color, points, mask, rgbs, M, m_dev, scratch);
    }
    return launch_status();
}
