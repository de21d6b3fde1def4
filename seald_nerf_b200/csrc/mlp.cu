// mlp.cu — FFMLP-compatible fused MLP (reference: ffmlp/ffmlp.py:99-168, ffmlp/src/ffmlp.cu:635-895) on the building
// blocks of mlp.cuh.  Bias-free, ReLU hidden activations, no output activation; weights are ONE flat fp16 buffer laid
// out [W0: hidden x in][W1..: hidden x hidden][Wout: 16 x hidden], each row-major [out][in] (ffmlp.cu:632).
#include "mlp.cuh"

namespace seald {

constexpr int kFfK0Max = 128;  // input_dim <= 128 (multiple of 16)

template <int WIDTH, bool SAVE>
__global__ void __launch_bounds__(kMlpThreads, (WIDTH <= 64 ? 2 : 1)) k_ffmlp_forward(const __half* __restrict__ inputs, const MlpWeights mw,
                                                                                     const int B, __half* __restrict__ fwd_buf,
                                                                                     __half* __restrict__ outputs) {
    using SM = MlpSmem<WIDTH, kFfK0Max>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_in = reinterpret_cast<__half*>(smem_raw);
    __half* s_w = s_in + SM::IN_HALVES;
    float* s_out = reinterpret_cast<float*>(s_w + 2 * SM::W_HALVES);
    const int n_tiles = (B + kTileRows - 1) / kTileRows;
    const int chunks = mw.k0 / 8;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows;
        for (int i = threadIdx.x; i < kTileRows * chunks; i += kMlpThreads) {
            const int r = i / chunks, c = i - r * chunks;
            __half* dst = s_in + r * SM::IN_STRIDE + c * 8;
            if (row0 + r < B) cp_async16(dst, inputs + (size_t)(row0 + r) * mw.k0 + c * 8);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
        mlp_forward_tile<WIDTH, kFfK0Max, SAVE>(mw, s_in, s_w, s_out, fwd_buf, B, row0);
        for (int i = threadIdx.x; i < kTileRows * 16; i += kMlpThreads) {
            const int r = i >> 4, c = i & 15;
            if (row0 + r < B) outputs[(size_t)(row0 + r) * 16 + c] = __float2half_rn(s_out[r * kOutStride + c]);
        }
        __syncthreads();
    }
}

template <int WIDTH>
__global__ void __launch_bounds__(kMlpThreads, (WIDTH <= 64 ? 2 : 1)) k_ffmlp_backward(const __half* __restrict__ grad, const MlpWeights mw,
                                                                                      const int B, const __half* __restrict__ fwd_buf,
                                                                                      __half* __restrict__ bwd_buf,
                                                                                      __half* __restrict__ grad_inputs) {
    using SM = MlpSmem<WIDTH, kFfK0Max>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_g = reinterpret_cast<__half*>(smem_raw);
    __half* s_w = s_g + SM::IN_HALVES;
    const int n_tiles = (B + kTileRows - 1) / kTileRows;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kTileRows;
        for (int i = threadIdx.x; i < kTileRows * 2; i += kMlpThreads) {
            const int r = i >> 1, c = i & 1;
            __half* dst = s_g + r * kGStride + c * 8;
            if (row0 + r < B) cp_async16(dst, grad + (size_t)(row0 + r) * 16 + c * 8);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
        mlp_backward_tile<WIDTH, kFfK0Max>(mw, s_g, s_w, fwd_buf, bwd_buf, grad_inputs, mw.k0, row0, B, B, row0);
    }
}

}  // namespace seald

using namespace seald;

namespace {

int ff_weights(MlpWeights& mw, const void* weights, uint32_t in_dim, uint32_t hidden, uint32_t num_layers) {
    // FFMLP(num_layers = n) owns n + 1 matrices (ffmlp.py:121-122)
    const int n_mat = (int)num_layers + 1;
    if (!weights || n_mat < 2 || n_mat > kMaxLayers) return SEALD_E_BADARG;
    if (in_dim == 0 || in_dim % 16 || in_dim > (uint32_t)kFfK0Max) return SEALD_E_UNSUPPORTED;
    if (((uintptr_t)weights % 16) != 0) return SEALD_E_ALIGN;
    const __half* w = reinterpret_cast<const __half*>(weights);
    mw.w[0] = w;
    w += (size_t)hidden * in_dim;
    for (int i = 1; i < n_mat - 1; i++) { mw.w[i] = w; w += (size_t)hidden * hidden; }
    mw.w[n_mat - 1] = w;
    mw.n_layers = n_mat;
    mw.k0 = (int)in_dim;
    mw.k0_ld = (int)in_dim;
    mw.n_out = 16;
    return 0;
}

template <int WIDTH>
int ff_forward(const __half* in, const MlpWeights& mw, int B, __half* fwd_buf, __half* out, cudaStream_t st) {
    const size_t smem = MlpSmem<WIDTH, kFfK0Max>::BYTES;
    const int per_sm = WIDTH <= 64 ? 2 : 1;
    const int n_tiles = (B + kTileRows - 1) / kTileRows;
    const int grid = n_tiles < SEALD_NUM_SMS * per_sm ? n_tiles : SEALD_NUM_SMS * per_sm;
    cudaError_t e;
    if (fwd_buf) {
        e = cudaFuncSetAttribute(k_ffmlp_forward<WIDTH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k_ffmlp_forward<WIDTH, true><<<grid, kMlpThreads, smem, st>>>(in, mw, B, fwd_buf, out);
    } else {
        e = cudaFuncSetAttribute(k_ffmlp_forward<WIDTH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k_ffmlp_forward<WIDTH, false><<<grid, kMlpThreads, smem, st>>>(in, mw, B, nullptr, out);
    }
    return launch_status();
}

template <int WIDTH>
int ff_backward(const __half* grad, const MlpWeights& mw, int B, const __half* fwd_buf, __half* bwd_buf, __half* grad_inputs, cudaStream_t st) {
    const size_t smem = MlpSmem<WIDTH, kFfK0Max>::BYTES;
    const int per_sm = WIDTH <= 64 ? 2 : 1;
    const int n_tiles = (B + kTileRows - 1) / kTileRows;
    const int grid = n_tiles < SEALD_NUM_SMS * per_sm ? n_tiles : SEALD_NUM_SMS * per_sm;
    cudaError_t e = cudaFuncSetAttribute(k_ffmlp_backward<WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_ffmlp_backward<WIDTH><<<grid, kMlpThreads, smem, st>>>(grad, mw, B, fwd_buf, bwd_buf, grad_inputs);
    return launch_status();
}

}  // namespace

// Replaces ffmlp_forward / ffmlp_inference (ffmlp/src/ffmlp.h:8-9): fwd_buf == NULL is the inference variant.
extern "C" int seald_ffmlp_forward(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                   uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation, void* fwd_buf,
                                   void* outputs, seald_stream_t stream) {
    if (B == 0) return 0;
    if (!inputs || !outputs) return SEALD_E_BADARG;
    if (activation != 0 /*relu*/ || output_activation != 6 /*none*/ || output_dim != 16) return SEALD_E_UNSUPPORTED;
    MlpWeights mw;
    int rc = ff_weights(mw, weights, input_dim, hidden_dim, num_layers);
    if (rc) return rc;
    cudaStream_t st = to_stream(stream);
    switch (hidden_dim) {
        case 16: return ff_forward<16>((const __half*)inputs, mw, (int)B, (__half*)fwd_buf, (__half*)outputs, st);
        case 32: return ff_forward<32>((const __half*)inputs, mw, (int)B, (__half*)fwd_buf, (__half*)outputs, st);
        case 64: return ff_forward<64>((const __half*)inputs, mw, (int)B, (__half*)fwd_buf, (__half*)outputs, st);
        case 128: return ff_forward<128>((const __half*)inputs, mw, (int)B, (__half*)fwd_buf, (__half*)outputs, st);
        default: return SEALD_E_UNSUPPORTED;
    }
}

// Replaces ffmlp_backward (ffmlp/src/ffmlp.h:11): fused activation-gradient pass + weight-gradient GEMMs.
// grad [B,16] f16; bwd_buf [num_layers,B,hidden] f16 scratch; grad_inputs [B,input_dim] f16 or NULL;
// grad_weights: flat fp32 buffer with the layout of `weights`, ACCUMULATED into (caller pre-zeroes).
extern "C" int seald_mlp_wgrad(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, seald_stream_t stream);

extern "C" int seald_ffmlp_backward(const void* grad, const void* inputs, const void* weights, const void* fwd_buf, uint32_t B,
                                    uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                    uint32_t output_activation, void* bwd_buf, void* grad_inputs, float* grad_weights, seald_stream_t stream) {
    if (B == 0) return 0;
    if (!grad || !inputs || !fwd_buf || !bwd_buf || !grad_weights) return SEALD_E_BADARG;
    if (activation != 0 || output_activation != 6 || output_dim != 16) return SEALD_E_UNSUPPORTED;
    MlpWeights mw;
    int rc = ff_weights(mw, weights, input_dim, hidden_dim, num_layers);
    if (rc) return rc;
    cudaStream_t st = to_stream(stream);
    switch (hidden_dim) {
        case 16: rc = ff_backward<16>((const __half*)grad, mw, (int)B, (const __half*)fwd_buf, (__half*)bwd_buf, (__half*)grad_inputs, st); break;
        case 32: rc = ff_backward<32>((const __half*)grad, mw, (int)B, (const __half*)fwd_buf, (__half*)bwd_buf, (__half*)grad_inputs, st); break;
        case 64: rc = ff_backward<64>((const __half*)grad, mw, (int)B, (const __half*)fwd_buf, (__half*)bwd_buf, (__half*)grad_inputs, st); break;
        case 128: rc = ff_backward<128>((const __half*)grad, mw, (int)B, (const __half*)fwd_buf, (__half*)bwd_buf, (__half*)grad_inputs, st); break;
        default: return SEALD_E_UNSUPPORTED;
    }
    if (rc) return rc;
    // weight gradients: dW_l = G_l^T h_l
    const int n_mat = (int)num_layers + 1;
    const int H = (int)hidden_dim;
    const __half* fb = (const __half*)fwd_buf;
    const __half* bb = (const __half*)bwd_buf;
    seald_wgrad_job jobs[kMaxLayers];
    float* gw = grad_weights;
    for (int l = 0; l < n_mat; l++) {
        seald_wgrad_job& j = jobs[l];
        const bool last = (l == n_mat - 1);
        j.G = last ? grad : (const void*)(bb + (size_t)l * B * H);
        j.ldg = last ? 16 : H;
        j.N = last ? 16 : H;
        j.n_real = j.N;
        j.A = (l == 0) ? inputs : (const void*)(fb + (size_t)(l - 1) * B * H);
        j.lda = (l == 0) ? (int)input_dim : H;
        j.K = j.lda;
        j.k_real = j.K;
        j.ldw = j.K;
        j.dW = gw;
        gw += (size_t)j.N * j.K;
    }
    return seald_mlp_wgrad(jobs, n_mat, B, nullptr, stream);
}
