// encoders.cu — stand-alone frequency / spherical-harmonics encoders (fp32), the drop-in replacements of the
// reference's freqencoder and shencoder extensions.  In the fused field kernels (field.cu) the same device
// functions run in the MLP input stage and these outputs never touch HBM.
#include "encoders.cuh"

namespace seald {

// freqencoder.cu:30-58: one thread per output element (coalesced row-major store).
__global__ void k_freq_fwd(const float* __restrict__ inputs, const uint32_t B, const uint32_t D, const uint32_t deg, const uint32_t C,
                           float* __restrict__ outputs) {
    const size_t t = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
    if (t >= (size_t)B * C) return;
    const uint32_t b = t / C;
    const uint32_t c = t - (size_t)b * C;
    outputs[t] = freq_channel(inputs + (size_t)b * D, D, c);
}

// freqencoder.cu:63-94
__global__ void k_freq_bwd(const float* __restrict__ grad, const float* __restrict__ outputs, const uint32_t B, const uint32_t D,
                           const uint32_t deg, const uint32_t C, float* __restrict__ grad_inputs) {
    const size_t t = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
    if (t >= (size_t)B * D) return;
    const uint32_t b = t / D;
    const uint32_t d = t - (size_t)b * D;
    grad += (size_t)b * C;
    outputs += (size_t)b * C;
    float result = grad[d];
    grad += D;
    outputs += D;
    for (uint32_t f = 0; f < deg; f++) {
        result += scalbnf(1.0f, f) * (grad[d] * outputs[D + d] - grad[D + d] * outputs[d]);
        grad += 2 * D;
        outputs += 2 * D;
    }
    grad_inputs[t] = result;
}

template <int DEG>
__global__ void k_sh_fwd(const float* __restrict__ inputs, float* __restrict__ outputs, const uint32_t B, float* __restrict__ dy_dx) {
    const uint32_t b = threadIdx.x + blockIdx.x * blockDim.x;
    if (b >= B) return;
    constexpr int C2 = DEG * DEG;
    constexpr int LOW = DEG < 4 ? DEG : 4;  // bands 0..3: the hard-coded forms; 4..7: sh_high_bands
    const float x = inputs[(size_t)b * 3], y = inputs[(size_t)b * 3 + 1], z = inputs[(size_t)b * 3 + 2];
    float o[C2];
    float* out = outputs + (size_t)b * C2;
    if (!dy_dx) {
        sh_eval<LOW>(x, y, z, o);
        if constexpr (DEG > 4) sh_high_bands<DEG, false>(x, y, z, o, nullptr, nullptr, nullptr);
#pragma unroll
        for (int i = 0; i < C2; i++) out[i] = o[i];
    } else {
        float gx[C2], gy[C2], gz[C2];
        sh_eval<LOW>(x, y, z, o);
        sh_grad<LOW>(x, y, z, gx, gy, gz);
        if constexpr (DEG > 4) sh_high_bands<DEG, true>(x, y, z, o, gx, gy, gz);
#pragma unroll
        for (int i = 0; i < C2; i++) out[i] = o[i];
        float* g = dy_dx + (size_t)b * 3 * C2;  // [B, D, C2] (shencoder.cu:139-141)
#pragma unroll
        for (int i = 0; i < C2; i++) { g[i] = gx[i]; g[C2 + i] = gy[i]; g[2 * C2 + i] = gz[i]; }
    }
}

// shencoder.cu:359-381 (the reference accumulates into a pre-zeroed buffer; we overwrite)
__global__ void k_sh_bwd(const float* __restrict__ grad, const uint32_t B, const uint32_t D, const uint32_t C2,
                         const float* __restrict__ dy_dx, float* __restrict__ grad_inputs) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    const uint32_t b = t / D;
    if (b >= B) return;
    const uint32_t d = t - b * D;
    grad += (size_t)b * C2;
    dy_dx += (size_t)b * D * C2 + (size_t)d * C2;
    float acc = 0.0f;
    for (uint32_t ch = 0; ch < C2; ch++) acc += grad[ch] * dy_dx[ch];
    grad_inputs[t] = acc;
}

}  // namespace seald

using namespace seald;

extern "C" int seald_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs,
                                         seald_stream_t stream) {
    if (B == 0) return 0;
    if (!inputs || !outputs) return SEALD_E_BADARG;
    if (D == 0 || C != D + 2 * D * deg) return SEALD_E_BADARG;
    const size_t n = (size_t)B * C;
    k_freq_fwd<<<(uint32_t)div_up(n, (size_t)256), 256, 0, to_stream(stream)>>>(inputs, B, D, deg, C, outputs);
    return launch_status();
}

extern "C" int seald_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                          float* grad_inputs, seald_stream_t stream) {
    if (B == 0) return 0;
    if (!grad || !outputs || !grad_inputs) return SEALD_E_BADARG;
    if (D == 0 || C != D + 2 * D * deg) return SEALD_E_BADARG;
    const size_t n = (size_t)B * D;
    k_freq_bwd<<<(uint32_t)div_up(n, (size_t)256), 256, 0, to_stream(stream)>>>(grad, outputs, B, D, deg, C, grad_inputs);
    return launch_status();
}

extern "C" int seald_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, float* dy_dx,
                                       seald_stream_t stream) {
    if (B == 0) return 0;
    if (!inputs || !outputs) return SEALD_E_BADARG;
    if (D != 3) return SEALD_E_UNSUPPORTED;
    cudaStream_t st = to_stream(stream);
    const uint32_t blocks = div_up(B, 256u);
    switch (degree) {
        case 1: k_sh_fwd<1><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        case 2: k_sh_fwd<2><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        case 3: k_sh_fwd<3><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        case 4: k_sh_fwd<4><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        case 5: k_sh_fwd<5><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        case 6: k_sh_fwd<6><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        case 7: k_sh_fwd<7><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        case 8: k_sh_fwd<8><<<blocks, 256, 0, st>>>(inputs, outputs, B, dy_dx); break;
        default: return SEALD_E_UNSUPPORTED;
    }
    return launch_status();
}

extern "C" int seald_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx,
                                        float* grad_inputs, seald_stream_t stream) {
    if (B == 0) return 0;
    if (!grad || !dy_dx || !grad_inputs) return SEALD_E_BADARG;
    if (D != 3 || degree < 1 || degree > 8) return SEALD_E_UNSUPPORTED;
    (void)inputs;
    k_sh_bwd<<<div_up(B * D, 256u), 256, 0, to_stream(stream)>>>(grad, B, D, degree * degree, dy_dx, grad_inputs);
    return launch_status();
}

// ---- trunc_exp (activation.py:5-17): y = exp(x) in fp32; dx = g * exp(clamp(x, -15, 15)) ------------------------------------
namespace seald {
__global__ void k_trunc_exp_fwd(const float* __restrict__ x, float* __restrict__ y, const size_t n) {
    for (size_t i = threadIdx.x + (size_t)blockIdx.x * blockDim.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = expf(x[i]);
}
__global__ void k_trunc_exp_bwd(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ gx, const size_t n) {
    for (size_t i = threadIdx.x + (size_t)blockIdx.x * blockDim.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        gx[i] = g[i] * expf(clampf(x[i], -15.0f, 15.0f));
}
}  // namespace seald

extern "C" int seald_trunc_exp_forward(const float* x, float* y, uint64_t n, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!x || !y) return SEALD_E_BADARG;
    const uint32_t blocks = (uint32_t)(n / 256 + 1 < 8u * SEALD_NUM_SMS ? n / 256 + 1 : 8u * SEALD_NUM_SMS);
    seald::k_trunc_exp_fwd<<<blocks, 256, 0, seald::to_stream(stream)>>>(x, y, (size_t)n);
    return seald::launch_status();
}

extern "C" int seald_trunc_exp_backward(const float* grad, const float* x, float* grad_x, uint64_t n, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!grad || !x || !grad_x) return SEALD_E_BADARG;
    const uint32_t blocks = (uint32_t)(n / 256 + 1 < 8u * SEALD_NUM_SMS ? n / 256 + 1 : 8u * SEALD_NUM_SMS);
    seald::k_trunc_exp_bwd<<<blocks, 256, 0, seald::to_stream(stream)>>>(grad, x, grad_x, (size_t)n);
    return seald::launch_status();
}
