// train.cu — the element-wise glue of a training step as fused kernels: background blend + MSE loss + its gradient
// (dnerf/utils.py:74-85, dnerf/renderer.py:325-326), fp32 -> fp16 weight staging, gradient finiteness check, dynamic
// loss scaling (torch.cuda.amp.GradScaler semantics, nerf/utils.py:884-886) and Adam (main_dnerf.py:129).
#include "common.cuh"
#include "adam.cuh"

namespace seald {

// pred = image + (1 - ws) * bg ; loss_sum += sum (pred - gt)^2 ; grads of  loss_scale * mean((pred-gt)^2)
__global__ void k_mse_loss_bg(const float* __restrict__ image, const float* __restrict__ weights_sum, const float* __restrict__ bg,
                              const float* __restrict__ gt, const uint32_t N, const float inv_count, const float* __restrict__ loss_scale,
                              float* __restrict__ pred, float* __restrict__ loss_sum, float* __restrict__ grad_image,
                              float* __restrict__ grad_ws) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    float local = 0.0f;
    if (n < N) {
        const float scale = loss_scale ? *loss_scale : 1.0f;
        const float ws = weights_sum[n];
        float gws = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float b = bg ? bg[(size_t)n * 3 + c] : 1.0f;
            const float p = image[(size_t)n * 3 + c] + (1 - ws) * b;
            const float diff = p - gt[(size_t)n * 3 + c];
            local += diff * diff;
            const float gi = 2.0f * diff * inv_count * scale;
            if (pred) pred[(size_t)n * 3 + c] = p;
            grad_image[(size_t)n * 3 + c] = gi;
            gws -= b * gi;
        }
        grad_ws[n] = gws;
    }
    // block reduction of the loss
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ float s_part[32];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(loss_sum, v * inv_count);
    }
}

// Seal local pre-training loss (SealNeRF/trainer.py:448-462): L1Loss(sigma, gt_sigma) + L1Loss(rgb, gt_rgb), both 'mean' reductions
// (over M resp. 3M elements); loss_sum += the loss, grads of loss_scale * loss (sign(0) = 0 like torch's l1_loss backward).
__global__ void k_l1_pretrain_loss(const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ gt_sigma,
                                   const float* __restrict__ gt_rgb, const uint32_t M, const float* __restrict__ loss_scale,
                                   float* __restrict__ loss_sum, float* __restrict__ grad_sigma, float* __restrict__ grad_rgb) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    const float inv_s = 1.0f / (float)M, inv_c = 1.0f / (3.0f * (float)M);
    float local = 0.0f;
    if (n < M) {
        const float scale = loss_scale ? *loss_scale : 1.0f;
        const float ds = sigma[n] - gt_sigma[n];
        local += fabsf(ds) * inv_s;
        grad_sigma[n] = (ds > 0.0f ? 1.0f : (ds < 0.0f ? -1.0f : 0.0f)) * inv_s * scale;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float dc = rgb[(size_t)n * 3 + c] - gt_rgb[(size_t)n * 3 + c];
            local += fabsf(dc) * inv_c;
            grad_rgb[(size_t)n * 3 + c] = (dc > 0.0f ? 1.0f : (dc < 0.0f ? -1.0f : 0.0f)) * inv_c * scale;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ float s_part[32];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(loss_sum, v);
    }
}

__global__ void k_cast_pad_f16(const float* __restrict__ src, __half* __restrict__ dst, const uint32_t rows, const uint32_t cols,
                               const uint32_t ld) {
    const size_t i = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
    if (i >= (size_t)rows * ld) return;
    const uint32_t r = i / ld, c = i - (size_t)r * ld;
    dst[i] = __float2half_rn(c < cols ? src[(size_t)r * cols + c] : 0.0f);
}

// Batched variant: every MLP weight of the field in one launch (13 matrices -> one kernel instead of 13).
struct CastSeg { const float* src; __half* dst; uint32_t rows, cols, ld, first; };  // first = running element offset
constexpr int kMaxCastSegs = 32;
struct CastSegs { CastSeg s[kMaxCastSegs]; int n; uint32_t total; };

__global__ void k_cast_pad_f16_batch(const CastSegs segs) {
    const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= segs.total) return;
    int k = 0;
    while (k + 1 < segs.n && i >= segs.s[k + 1].first) k++;
    const CastSeg sg = segs.s[k];
    const uint32_t j = i - sg.first;
    const uint32_t r = j / sg.ld, c = j - r * sg.ld;
    sg.dst[j] = __float2half_rn(c < sg.cols ? sg.src[(size_t)r * sg.cols + c] : 0.0f);
}

// found_inf[0] |= bits of 1.0f if any gradient is inf/nan (non-zero as an int AND a positive float: the flag is summed over the
// ranks by a float all-reduce)
__global__ void k_grad_finite_check(const float* __restrict__ g, const size_t n, int* __restrict__ found_inf) {
    bool bad = false;
    const size_t n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (size_t i = threadIdx.x + (size_t)blockIdx.x * blockDim.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    }
    for (size_t i = n4 * 4 + threadIdx.x + (size_t)blockIdx.x * blockDim.x; i < n; i += (size_t)gridDim.x * blockDim.x) bad |= !isfinite(g[i]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(found_inf, 0x3f800000);
}

// torch.optim.Adam (no amsgrad / weight decay) on a flat fp32 slab, gradients un-scaled by 1 / *loss_scale, the whole
// update skipped when *found_inf != 0 (GradScaler.step).  Optionally refreshes an fp16 copy and zeroes the gradient.
__global__ void __launch_bounds__(256, 4) k_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, const size_t n,
                       const float lr, const float beta1, const float beta2, const float eps, float bc1, float bc2_sqrt,
                       const int* __restrict__ step_dev, const int step_add, const float* __restrict__ loss_scale,
                       const int* __restrict__ found_inf, __half* __restrict__ p16, const int zero_grad, const float* __restrict__ lr_scale) {
    const bool skip = found_inf && *found_inf != 0;
    if (step_dev) {  // step number kept on the device (graph replay): bias corrections computed here
        __shared__ float s_bc[2];
        if (threadIdx.x == 0) {
            const double st = (double)max(*step_dev + step_add, 1);
            s_bc[0] = (float)(1.0 - pow((double)beta1, st));
            s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, st));
        }
        __syncthreads();
        bc1 = s_bc[0];
        bc2_sqrt = s_bc[1];
    }
    const float inv_scale = loss_scale ? 1.0f / *loss_scale : 1.0f;
    const float lr_eff = lr_scale ? lr * *lr_scale : lr;  // LambdaLR factor kept on the device (main_dnerf.py:134)
    adam_slab_body(blockIdx.x, gridDim.x, p, g, m, v, n, beta1, beta2, eps, bc2_sqrt, inv_scale, lr_eff / bc1, skip, p16, zero_grad);
}

// optimizer.step() is skipped by GradScaler when a non-finite gradient was found: the step counter only advances otherwise
__global__ void k_adam_advance(int* step_dev, const int* __restrict__ found_inf) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && !(found_inf && *found_inf)) *step_dev += 1;
}

// GradScaler.update(): found_inf -> scale *= backoff, tracker = 0; else tracker++ and scale *= growth every `interval`.
// stash (optional, 4 words): {found_inf, step counter, loss scale} as they were for THIS step's update, kept for the part of the
// optimiser that the trainer defers into the next step (the hash-table Adam pass runs beside the next march).
__global__ void k_loss_scale_update(float* loss_scale, int* found_inf, int* growth_tracker, const float growth, const float backoff,
                                    const int interval, int* step_dev, int* stash) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (stash) {
            stash[0] = *found_inf;
            stash[1] = step_dev ? *step_dev : 0;
            stash[2] = __float_as_int(*loss_scale);
        }
        if (step_dev && !*found_inf) *step_dev += 1;  // optimizer.step() ran: one more completed update
        if (*found_inf) {
            *loss_scale *= backoff;
            *growth_tracker = 0;
        } else {
            const int t = *growth_tracker + 1;
            if (t >= interval) {
                *loss_scale *= growth;
                *growth_tracker = 0;
            } else {
                *growth_tracker = t;
            }
        }
        *found_inf = 0;
    }
}

// Start of a training step (dnerf/renderer.py:285,291-292): pick the occupancy frame of this time stamp
// t_idx = clamp(floor(time * T), 0, T-1), copy its bitfield and occupied-cell box into the step's static buffers and reset
// the sample counter - one launch instead of a chain of small tensor ops.
__global__ void k_select_frame(const float* __restrict__ time, const uint32_t T, const uint4* __restrict__ bitfield_all, const uint32_t frame_vec16,
                               uint4* __restrict__ bitfield_out, const float* __restrict__ occ_all, float* __restrict__ occ_out,
                               int* __restrict__ counter) {
    int t_idx = (int)floorf(*time * (float)T);
    t_idx = min(max(t_idx, 0), (int)T - 1);
    const uint4* src = bitfield_all + (size_t)t_idx * frame_vec16;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < frame_vec16; i += gridDim.x * blockDim.x) bitfield_out[i] = src[i];
    if (blockIdx.x == 0) {
        if (threadIdx.x < 6 && occ_all && occ_out) occ_out[threadIdx.x] = occ_all[(size_t)t_idx * 6 + threadIdx.x];
        if (threadIdx.x < 2 && counter) counter[threadIdx.x] = 0;
    }
}

// Start of a fused training step in ONE launch: select_frame + the zero of the loss accumulator + the step's perturbation noise
// (noises [n_noise] uniform in [0, 1): the reference draws torch.rand(N) per march, raymarching.py:179-181).  The noise is a
// counter-based hash of (ctr[0], ray) — two rounds of a 64-bit mix (splitmix64 finaliser), the top 24 bits as the mantissa — so a
// CUDA-graph replay produces a fresh stream without host RNG state; ctr[0] is advanced by the LAST CTA to finish (ticket in ctr[1]),
// after every CTA has read it.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_step_begin(const float* __restrict__ time, const uint32_t T, const uint4* __restrict__ bitfield_all, const uint32_t frame_vec16,
                             uint4* __restrict__ bitfield_out, const float* __restrict__ occ_all, float* __restrict__ occ_out,
                             int* __restrict__ counter, float* __restrict__ loss_sum, float* __restrict__ noises, const uint32_t n_noise,
                             unsigned long long* __restrict__ ctr) {
    int t_idx = (int)floorf(*time * (float)T);
    t_idx = min(max(t_idx, 0), (int)T - 1);
    const uint4* src = bitfield_all + (size_t)t_idx * frame_vec16;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < frame_vec16; i += gridDim.x * blockDim.x) bitfield_out[i] = src[i];
    if (noises) {
        const uint64_t base = mix64(ctr[0] * 0x9E3779B97F4A7C15ull + 0x1234567ull);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_noise; i += gridDim.x * blockDim.x)
            noises[i] = (float)(mix64(base + (uint64_t)i * 0x9E3779B97F4A7C15ull) >> 40) * (1.0f / 16777216.0f);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x < 6 && occ_all && occ_out) occ_out[threadIdx.x] = occ_all[(size_t)t_idx * 6 + threadIdx.x];
        if (threadIdx.x < 2 && counter) counter[threadIdx.x] = 0;
        if (threadIdx.x == 0 && loss_sum) *loss_sum = 0.0f;
    }
    if (noises) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned long long done = atomicAdd(ctr + 1, 1ull) + 1ull;
            if (done == gridDim.x) { ctr[1] = 0ull; ctr[0] += 1ull; }
        }
    }
}

}  // namespace seald

using namespace seald;

extern "C" int seald_step_begin(const float* time_dev, uint32_t T, const uint8_t* bitfield_all, uint32_t frame_bytes, uint8_t* bitfield_out,
                                const float* occ_all, float* occ_out, int32_t* counter, float* loss_sum, float* noises, uint32_t n_noise,
                                uint64_t* noise_ctr2, seald_stream_t stream) {
    if (!time_dev || !bitfield_all || !bitfield_out || T == 0 || frame_bytes == 0) return SEALD_E_BADARG;
    if (noises && !noise_ctr2) return SEALD_E_BADARG;
    if (frame_bytes % 16 || ((uintptr_t)bitfield_all & 15) || ((uintptr_t)bitfield_out & 15)) return SEALD_E_ALIGN;
    const uint32_t n16 = frame_bytes / 16;
    const uint32_t blocks = div_up(n16, 256u) < (uint32_t)SEALD_NUM_SMS ? div_up(n16, 256u) : (uint32_t)SEALD_NUM_SMS;
    k_step_begin<<<blocks, 256, 0, to_stream(stream)>>>(time_dev, T, reinterpret_cast<const uint4*>(bitfield_all), n16,
                                                       reinterpret_cast<uint4*>(bitfield_out), occ_all, occ_out, counter, loss_sum, noises, n_noise,
                                                       reinterpret_cast<unsigned long long*>(noise_ctr2));
    return launch_status();
}

extern "C" int seald_select_frame(const float* time_dev, uint32_t T, const uint8_t* bitfield_all, uint32_t frame_bytes, uint8_t* bitfield_out,
                                  const float* occ_all, float* occ_out, int32_t* counter, seald_stream_t stream) {
    if (!time_dev || !bitfield_all || !bitfield_out || T == 0 || frame_bytes == 0) return SEALD_E_BADARG;
    if (frame_bytes % 16 || ((uintptr_t)bitfield_all & 15) || ((uintptr_t)bitfield_out & 15)) return SEALD_E_ALIGN;
    const uint32_t n16 = frame_bytes / 16;
    const uint32_t blocks = div_up(n16, 256u) < (uint32_t)SEALD_NUM_SMS ? div_up(n16, 256u) : (uint32_t)SEALD_NUM_SMS;
    k_select_frame<<<blocks, 256, 0, to_stream(stream)>>>(time_dev, T, reinterpret_cast<const uint4*>(bitfield_all), n16,
                                                         reinterpret_cast<uint4*>(bitfield_out), occ_all, occ_out, counter);
    return launch_status();
}

extern "C" int seald_mse_loss_bg(const float* image, const float* weights_sum, const float* bg, const float* gt, uint32_t N, float inv_count,
                                 const float* loss_scale, float* pred, float* loss_sum, float* grad_image, float* grad_ws,
                                 seald_stream_t stream) {
    if (N == 0) return 0;
    if (!image || !weights_sum || !gt || !loss_sum || !grad_image || !grad_ws) return SEALD_E_BADARG;
    k_mse_loss_bg<<<div_up(N, 256u), 256, 0, to_stream(stream)>>>(image, weights_sum, bg, gt, N, inv_count, loss_scale, pred, loss_sum, grad_image,
                                                                 grad_ws);
    return launch_status();
}

extern "C" int seald_l1_pretrain_loss(const float* sigma, const float* rgb, const float* gt_sigma, const float* gt_rgb, uint32_t M,
                                      const float* loss_scale, float* loss_sum, float* grad_sigma, float* grad_rgb, seald_stream_t stream) {
    if (M == 0) return 0;
    if (!sigma || !rgb || !gt_sigma || !gt_rgb || !loss_sum || !grad_sigma || !grad_rgb) return SEALD_E_BADARG;
    k_l1_pretrain_loss<<<div_up(M, 256u), 256, 0, to_stream(stream)>>>(sigma, rgb, gt_sigma, gt_rgb, M, loss_scale, loss_sum, grad_sigma, grad_rgb);
    return launch_status();
}

extern "C" int seald_cast_pad_f16(const float* src, void* dst, uint32_t rows, uint32_t cols, uint32_t ld, seald_stream_t stream) {
    if (rows == 0 || ld == 0) return 0;
    if (!src || !dst || ld < cols) return SEALD_E_BADARG;
    const size_t n = (size_t)rows * ld;
    k_cast_pad_f16<<<(uint32_t)div_up(n, (size_t)256), 256, 0, to_stream(stream)>>>(src, (__half*)dst, rows, cols, ld);
    return launch_status();
}

extern "C" int seald_cast_pad_f16_batch(const void* const* src, void* const* dst, const uint32_t* rows, const uint32_t* cols,
                                        const uint32_t* ld, int n, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!src || !dst || !rows || !cols || !ld || n < 0 || n > kMaxCastSegs) return SEALD_E_BADARG;
    CastSegs segs;
    uint32_t total = 0;
    for (int i = 0; i < n; i++) {
        if (!src[i] || !dst[i] || ld[i] < cols[i]) return SEALD_E_BADARG;
        segs.s[i] = CastSeg{(const float*)src[i], (__half*)dst[i], rows[i], cols[i], ld[i], total};
        total += rows[i] * ld[i];
    }
    segs.n = n;
    segs.total = total;
    k_cast_pad_f16_batch<<<div_up(total, 256u), 256, 0, to_stream(stream)>>>(segs);
    return launch_status();
}

extern "C" int seald_grad_finite_check(const float* g, uint64_t n, int32_t* found_inf, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!g || !found_inf) return SEALD_E_BADARG;
    if ((uintptr_t)g % 16) return SEALD_E_ALIGN;
    const uint32_t blocks = (uint32_t)(n / 4 / 256 + 1 < 4u * SEALD_NUM_SMS ? n / 4 / 256 + 1 : 4u * SEALD_NUM_SMS);
    k_grad_finite_check<<<blocks, 256, 0, to_stream(stream)>>>(g, (size_t)n, found_inf);
    return launch_status();
}

extern "C" int seald_adam_advance(int32_t* step_dev, const int32_t* found_inf, seald_stream_t stream) {
    if (!step_dev) return SEALD_E_BADARG;
    k_adam_advance<<<1, 32, 0, to_stream(stream)>>>(step_dev, found_inf);
    return launch_status();
}

static int adam_launch(float* p, float* g, float* m, float* v, uint64_t n, float lr, float beta1, float beta2, float eps, uint32_t step,
                       const int32_t* step_dev, const float* loss_scale, const int32_t* found_inf, void* p16, int zero_grad,
                       uint32_t max_blocks, seald_stream_t stream, const float* lr_scale = nullptr) {
    if (n == 0) return 0;
    if (!p || !g || !m || !v || (step == 0 && !step_dev)) return SEALD_E_BADARG;
    const uint32_t step_in = step;
    if (step == 0) step = 1;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    if (max_blocks == 0) max_blocks = 8u * SEALD_NUM_SMS;
    const uint32_t blocks = (uint32_t)(n / 256 + 1 < max_blocks ? n / 256 + 1 : max_blocks);
    // with step_dev: the step number of THIS update is *step_dev + step (step = 0: the counter was advanced already by
    // seald_adam_advance; step = 1: it counts completed updates and is advanced by seald_loss_scale_update afterwards)
    k_adam<<<blocks, 256, 0, to_stream(stream)>>>(p, g, m, v, (size_t)n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), step_dev,
                                                  step_dev ? (int)step_in : 0, loss_scale, found_inf, (__half*)p16, zero_grad, lr_scale);
    return launch_status();
}

extern "C" int seald_adam_step_lr(float* p, float* g, float* m, float* v, uint64_t n, float lr, const float* lr_scale_dev, float beta1,
                                  float beta2, float eps, uint32_t step, const int32_t* step_dev, const float* loss_scale,
                                  const int32_t* found_inf, void* p16, int zero_grad, uint32_t max_blocks, seald_stream_t stream) {
    return adam_launch(p, g, m, v, n, lr, beta1, beta2, eps, step, step_dev, loss_scale, found_inf, p16, zero_grad, max_blocks, stream,
                       lr_scale_dev);
}

extern "C" int seald_adam_step(float* p, float* g, float* m, float* v, uint64_t n, float lr, float beta1, float beta2, float eps, uint32_t step,
                               const int32_t* step_dev, const float* loss_scale, const int32_t* found_inf, void* p16, int zero_grad,
                               seald_stream_t stream) {
    return adam_launch(p, g, m, v, n, lr, beta1, beta2, eps, step, step_dev, loss_scale, found_inf, p16, zero_grad, 0, stream);
}

extern "C" int seald_adam_step_ex(float* p, float* g, float* m, float* v, uint64_t n, float lr, float beta1, float beta2, float eps,
                                  uint32_t step, const int32_t* step_dev, const float* loss_scale, const int32_t* found_inf, void* p16,
                                  int zero_grad, uint32_t max_blocks, seald_stream_t stream) {
    return adam_launch(p, g, m, v, n, lr, beta1, beta2, eps, step, step_dev, loss_scale, found_inf, p16, zero_grad, max_blocks, stream);
}

extern "C" int seald_loss_scale_update(float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth, float backoff,
                                       int interval, int32_t* step_dev, seald_stream_t stream) {
    if (!loss_scale || !found_inf || !growth_tracker) return SEALD_E_BADARG;
    k_loss_scale_update<<<1, 32, 0, to_stream(stream)>>>(loss_scale, found_inf, growth_tracker, growth, backoff, interval, step_dev, nullptr);
    return launch_status();
}

extern "C" int seald_loss_scale_update_stash(float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth, float backoff,
                                             int interval, int32_t* step_dev, int32_t* stash, seald_stream_t stream) {
    if (!loss_scale || !found_inf || !growth_tracker || !stash) return SEALD_E_BADARG;
    k_loss_scale_update<<<1, 32, 0, to_stream(stream)>>>(loss_scale, found_inf, growth_tracker, growth, backoff, interval, step_dev, stash);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Step input generation on the device: get_rays (nerf/utils.py:54-137) for the sampled pixels of ONE camera
// + the ground-truth gather of NeRFDataset.collate (dnerf/provider.py:340-343) in one kernel.  The reference runs ~15 small
// torch kernels per step for this (meshgrid, gather x2, stack, norm, div, bmm, expand, gather of the image); with the dataset
// preloaded on the GPU (provider `preload`) the host then only sends a frame index per step.
//   pixel (w, h) = (ind % W, ind / W), centre at +0.5; dir = ((i - cx) / fx, (j - cy) / fy, 1) normalised, rotated by pose[:3,:3];
//   origin = pose[:3, 3].  The division by the scalar intrinsics is a multiplication by the fp32 reciprocal, as torch does it.
//   gt: C = 3 copies RGB; C = 4 blends rgb * a + bg * (1 - a) (dnerf/utils.py:61-66) with bg[n] (NULL = white).
//   frame_dev selects pose / time / image from the resident tables; time_out receives times[frame].
// ------------------------------------------------------------------------------------------------
namespace seald {
__global__ void k_get_rays_gather(const float* __restrict__ poses, const float* __restrict__ times, const float* __restrict__ images,
                                  const int* __restrict__ frame_dev, const long long* __restrict__ inds, const uint32_t N, const uint32_t H,
                                  const uint32_t W, const uint32_t C, const float inv_fx, const float inv_fy, const float cx, const float cy,
                                  const float* __restrict__ bg, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                  float* __restrict__ gt, float* __restrict__ time_out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = frame_dev ? *frame_dev : 0;
    const float* P = poses + (size_t)frame * 16;
    if (n == 0 && time_out && times) *time_out = times[frame];
    if (n >= N) return;
    const long long ind = inds[n];
    const uint32_t w = (uint32_t)(ind % W), h = (uint32_t)(ind / W);
    const float i = __fadd_rn((float)w, 0.5f), j = __fadd_rn((float)h, 0.5f);
    const float xs = __fmul_rn(__fsub_rn(i, cx), inv_fx), ys = __fmul_rn(__fsub_rn(j, cy), inv_fy), zs = 1.0f;
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xs, xs), __fmul_rn(ys, ys)), __fmul_rn(zs, zs)));
    const float dx = __fdiv_rn(xs, nrm), dy = __fdiv_rn(ys, nrm), dz = __fdiv_rn(zs, nrm);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        rays_d[(size_t)n * 3 + r] = dx * P[r * 4] + dy * P[r * 4 + 1] + dz * P[r * 4 + 2];  // directions @ R^T
        rays_o[(size_t)n * 3 + r] = P[r * 4 + 3];
    }
    if (gt && images) {
        const float* px = images + ((size_t)frame * H * W + (size_t)ind) * C;
        if (C == 4) {
            const float a = px[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float b = bg ? bg[(size_t)n * 3 + c] : 1.0f;
                gt[(size_t)n * 3 + c] = __fadd_rn(__fmul_rn(px[c], a), __fmul_rn(b, __fsub_rn(1.0f, a)));
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++) gt[(size_t)n * 3 + c] = px[c];
        }
    }
}
}  // namespace seald

extern "C" int seald_get_rays_gather(const float* poses, const float* times, const float* images, const int32_t* frame_dev,
                                     const int64_t* inds, uint32_t N, uint32_t H, uint32_t W, uint32_t C, float fx, float fy, float cx,
                                     float cy, const float* bg, float* rays_o, float* rays_d, float* gt, float* time_out,
                                     seald_stream_t stream) {
    if (N == 0) return 0;
    if (!poses || !inds || !rays_o || !rays_d || H == 0 || W == 0 || fx == 0.0f || fy == 0.0f) return SEALD_E_BADARG;
    if (gt && images && C != 3 && C != 4) return SEALD_E_UNSUPPORTED;
    seald::k_get_rays_gather<<<seald::div_up(N, 256u), 256, 0, seald::to_stream(stream)>>>(poses, times, images, frame_dev, (const long long*)inds, N, H, W, C,
                                                                                          1.0f / fx, 1.0f / fy, cx, cy, bg, rays_o, rays_d, gt,
                                                                                          time_out);
    return seald::launch_status();
}
