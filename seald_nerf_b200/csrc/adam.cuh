// adam.cuh — the streaming Adam pass over a flat fp32 slab (torch.optim.Adam without amsgrad / weight decay), shared by k_adam
// (train.cu) and the single-launch optimiser k_mlp_tail<table> (optim_tail.cu).
#pragma once
#include "common.cuh"

namespace seald {

// One CTA's grid-stride share (bid of nblocks) of the slab; skip: GradScaler found an overflow (only the gradient is cleared).
__device__ __forceinline__ void adam_slab_body(const uint32_t bid, const uint32_t nblocks, float* __restrict__ p, float* __restrict__ g,
                                               float* __restrict__ m, float* __restrict__ v, const size_t n, const float beta1,
                                               const float beta2, const float eps, const float bc2_sqrt, const float inv_scale,
                                               const float step_size, const bool skip, __half* __restrict__ p16, const int zero_grad) {
    // 128-bit path over the aligned body (n is a multiple of 4 for the slabs the trainer passes)
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) && (!p16 || ((uintptr_t)p16 & 7) == 0);
    const size_t n4 = vec ? n / 4 : 0;
    // Two independent float4 lanes per thread and iteration (more bytes in flight per thread); streaming loads / stores: every
    // element is touched exactly once per step.  An entry whose gradient AND both moments are exactly zero (a table row no sample
    // has reached yet: most of the coarse dense levels outside the occupied region) is left alone - its update is exactly zero
    // (p -= lr * 0 / (0 + eps)) - which skips the read of p and all five writes for it.
    const size_t nth = (size_t)nblocks * blockDim.x;
    for (size_t i0 = threadIdx.x + (size_t)bid * blockDim.x; i0 < n4; i0 += 2 * nth) {
        float4 g4[2], m4[2], v4[2];
        bool live[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const size_t i = i0 + u * nth;
            live[u] = i < n4;
            if (live[u]) {
                g4[u] = __ldcs(reinterpret_cast<const float4*>(g) + i);
                if (!skip) {
                    m4[u] = __ldcs(reinterpret_cast<const float4*>(m) + i);
                    v4[u] = __ldcs(reinterpret_cast<const float4*>(v) + i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (!live[u]) continue;
            const size_t i = i0 + u * nth;
            const bool g_zero = g4[u].x == 0.f && g4[u].y == 0.f && g4[u].z == 0.f && g4[u].w == 0.f;
            if (!skip) {
                const bool untouched = g_zero && m4[u].x == 0.f && m4[u].y == 0.f && m4[u].z == 0.f && m4[u].w == 0.f && v4[u].x == 0.f &&
                                       v4[u].y == 0.f && v4[u].z == 0.f && v4[u].w == 0.f;
                if (untouched) continue;
                float4 p4 = __ldcs(reinterpret_cast<const float4*>(p) + i);
                float* pp = reinterpret_cast<float*>(&p4); float* gg = reinterpret_cast<float*>(&g4[u]);
                float* mm = reinterpret_cast<float*>(&m4[u]); float* vv = reinterpret_cast<float*>(&v4[u]);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float gi = gg[k] * inv_scale;
                    mm[k] = beta1 * mm[k] + (1.0f - beta1) * gi;
                    vv[k] = beta2 * vv[k] + (1.0f - beta2) * gi * gi;
                    pp[k] = pp[k] - step_size * (mm[k] / (sqrtf(vv[k]) / bc2_sqrt + eps));
                }
                __stcs(reinterpret_cast<float4*>(p) + i, p4);
                __stcs(reinterpret_cast<float4*>(m) + i, m4[u]);
                __stcs(reinterpret_cast<float4*>(v) + i, v4[u]);
                if (p16) {
                    const __half2 lo = __floats2half2_rn(pp[0], pp[1]), hi = __floats2half2_rn(pp[2], pp[3]);
                    uint2 w;
                    w.x = *reinterpret_cast<const uint32_t*>(&lo);
                    w.y = *reinterpret_cast<const uint32_t*>(&hi);
                    reinterpret_cast<uint2*>(p16)[i] = w;  // (the fp16 table is re-read by the next step's gathers: default caching)
                }
            }
            if (zero_grad && !g_zero) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (size_t i = n4 * 4 + threadIdx.x + (size_t)bid * blockDim.x; i < n; i += (size_t)nblocks * blockDim.x) {
        if (!skip) {
            const float gi = g[i] * inv_scale;
            const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
            const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            const float pi = p[i] - step_size * (mi / denom);
            p[i] = pi;
            if (p16) p16[i] = __float2half_rn(pi);
        }
        if (zero_grad) g[i] = 0.0f;
    }
}

}  // namespace seald
