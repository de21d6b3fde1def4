// encoders.cuh — device-side frequency and spherical-harmonics encodings shared by the stand-alone encoder
// kernels (encoders.cu) and the fused field kernels (field.cu).
#pragma once
#include "common.cuh"
#include "sh_tables.cuh"

namespace seald {

__device__ __forceinline__ constexpr float PI_F() { return 3.141592653589793f; }

// One output channel c of the frequency encoding of x[0..D) (freqencoder/src/freqencoder.cu:45-57):
// c < D: identity; else col = c / D - 1, d = c % D, freq = col / 2, phase = (col % 2) * pi/2, sin via __sinf.
__device__ __forceinline__ float freq_channel(const float* x, const uint32_t D, const uint32_t c) {
    if (c < D) return x[c];
    const uint32_t col = c / D - 1;
    const uint32_t d = c % D;
    const uint32_t freq = col / 2;
    const float phase_shift = (col % 2) * (PI_F() / 2);
    return __sinf(scalbnf(x[d], freq) + phase_shift);
}

// Real spherical harmonics up to degree 4 (16 values); polynomial forms as in shencoder/src/shencoder.cu:49-70.
template <int DEG>
__device__ __forceinline__ void sh_eval(const float x, const float y, const float z, float* out) {
    out[0] = 0.28209479177387814f;
    if constexpr (DEG >= 2) {
        out[1] = -0.48860251190291987f * y;
        out[2] = 0.48860251190291987f * z;
        out[3] = -0.48860251190291987f * x;
    }
    if constexpr (DEG >= 3) {
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        out[4] = 1.0925484305920792f * xy;
        out[5] = -1.0925484305920792f * yz;
        out[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
        out[7] = -1.0925484305920792f * xz;
        out[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
        if constexpr (DEG >= 4) {
            out[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
            out[10] = 2.8906114426405538f * xy * z;
            out[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
            out[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
            out[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
            out[14] = 1.4453057213202769f * z * (x2 - y2);
            out[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
        }
    }
}

// d out / d(x,y,z), each [DEG*DEG]
template <int DEG>
__device__ __forceinline__ void sh_grad(const float x, const float y, const float z, float* dx, float* dy, float* dz) {
    dx[0] = dy[0] = dz[0] = 0.0f;
    if constexpr (DEG >= 2) {
        dx[1] = 0.0f; dx[2] = 0.0f; dx[3] = -0.48860251190291992f;
        dy[1] = -0.48860251190291992f; dy[2] = 0.0f; dy[3] = 0.0f;
        dz[1] = 0.0f; dz[2] = 0.48860251190291992f; dz[3] = 0.0f;
    }
    if constexpr (DEG >= 3) {
        dx[4] = 1.0925484305920792f * y; dx[5] = 0.0f; dx[6] = 0.0f; dx[7] = -1.0925484305920792f * z; dx[8] = 1.0925484305920792f * x;
        dy[4] = 1.0925484305920792f * x; dy[5] = -1.0925484305920792f * z; dy[6] = 0.0f; dy[7] = 0.0f; dy[8] = -1.0925484305920792f * y;
        dz[4] = 0.0f; dz[5] = -1.0925484305920792f * y; dz[6] = 1.8923493915151202f * z; dz[7] = -1.0925484305920792f * x; dz[8] = 0.0f;
    }
    if constexpr (DEG >= 4) {
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        dx[9] = -3.5402615395598609f * xy;
        dx[10] = 2.8906114426405538f * yz;
        dx[11] = 0.0f;
        dx[12] = 0.0f;
        dx[13] = 0.45704579946446572f - 2.2852289973223288f * z2;
        dx[14] = 2.8906114426405538f * xz;
        dx[15] = -1.7701307697799304f * x2 + 1.7701307697799304f * y2;
        dy[9] = -1.7701307697799304f * x2 + 1.7701307697799304f * y2;
        dy[10] = 2.8906114426405538f * xz;
        dy[11] = 0.45704579946446572f - 2.2852289973223288f * z2;
        dy[12] = 0.0f;
        dy[13] = 0.0f;
        dy[14] = -2.8906114426405538f * yz;
        dy[15] = 3.5402615395598609f * xy;
        dz[9] = 0.0f;
        dz[10] = 2.8906114426405538f * xy;
        dz[11] = -4.5704579946446566f * yz;
        dz[12] = 5.597644988851731f * z2 - 1.1195289977703462f;
        dz[13] = -4.5704579946446566f * xz;
        dz[14] = 1.4453057213202769f * x2 - 1.4453057213202769f * y2;
        dz[15] = 0.0f;
    }
}

// Bands l = 4 .. DEG - 1 (degrees 5 .. 8 of the reference, shencoder.cu:71-135) from the generated z-polynomials: Y_l^{+-m} =
// Q_lm(z) * Re / Im((x + i y)^m).  out / dx / dy / dz are indexed like the reference: l * l + l + m.  The reference hard-codes the
// same polynomials expanded; here they are evaluated as a product of a polynomial in z and one in (x, y), so values agree to fp32
// round-off and the partial derivatives (x, y, z treated as independent, like the reference's dy_dx) are those of the same form.
template <int DEG, bool GRAD>
__device__ __forceinline__ void sh_high_bands(const float x, const float y, const float z, float* out, float* dx, float* dy, float* dz) {
    float c[8], s[8], zp[8];
    c[0] = 1.0f; s[0] = 0.0f; zp[0] = 1.0f;
#pragma unroll
    for (int m = 1; m < DEG; m++) {
        c[m] = x * c[m - 1] - y * s[m - 1];
        s[m] = x * s[m - 1] + y * c[m - 1];
        zp[m] = zp[m - 1] * z;
    }
#pragma unroll
    for (int l = 4; l < DEG; l++) {
#pragma unroll
        for (int m = 0; m <= l; m++) {
            float q = 0.0f, dq = 0.0f;
#pragma unroll
            for (int k = l - m; k >= 0; k -= 2) {  // Q_lm has the parity of l - m
                q += kShQ[l][m][k] * zp[k];
                if (GRAD && k >= 1) dq += (float)k * kShQ[l][m][k] * zp[k - 1];
            }
            const int ip = l * l + l + m, im = l * l + l - m;
            if (m == 0) {
                out[ip] = q;
                if (GRAD) { dx[ip] = 0.0f; dy[ip] = 0.0f; dz[ip] = dq; }
            } else {
                out[ip] = q * c[m];
                out[im] = q * s[m];
                if (GRAD) {
                    const float fm = (float)m;
                    dx[ip] = q * fm * c[m - 1]; dy[ip] = -q * fm * s[m - 1]; dz[ip] = dq * c[m];
                    dx[im] = q * fm * s[m - 1]; dy[im] = q * fm * c[m - 1]; dz[im] = dq * s[m];
                }
            }
        }
    }
}

}  // namespace seald
