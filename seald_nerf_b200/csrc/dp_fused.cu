// dp_fused.cu — the data-parallel exchange of a training step fused with the optimiser, over NVLink peer memory.
//
// Every rank holds the flat fp32 gradient buffer [hash table | MLP weights | overflow flag] and the fp16 working copy of the
// table in SYMMETRIC memory (same layout on every GPU, each GPU's copy mapped into every process, plus an NVSwitch multicast
// mapping).  Rank r owns table elements [r * shard, (r + 1) * shard).  Two kernels per rank replace
// reduce-scatter -> all-reduce -> Adam -> all-gather:
//
//   k_dp_reduce_shard      grad_shard = sum over ranks of grads[shard]: 16-byte loads straight from every peer's buffer (or
//                          multimem.ld_reduce: the sum happens IN the switch).  Launched on the trainer's second stream as
//                          soon as every rank finished its table scatter, so the NVLink transfer runs underneath the
//                          tensor-core backward of the deformation net (which does not touch the table gradient).
//   k_dp_adam_weights      sums the overflow flag and the small MLP region over the peers and runs Adam (torch.optim.Adam +
//                          GradScaler semantics exactly as k_adam, train.cu) on the replicated MLP weights.
//   k_dp_adam_shard_broadcast  Adam on the local fp32 p/m/v of the shard, refreshed fp16 rows stored into EVERY rank's table
//                          (multimem.st: one store, replicated by the switch; else one store per peer).  The trainer defers this
//                          kernel into the beginning of the next step, where it runs beside the march + deformation forward.
//
// Ordering is the caller's (trainer.py): a cross-rank barrier before each kernel (all table gradients / all MLP gradients
// complete) and one before the table is read or the gradient buffer is cleared again — that last one sits behind the next
// step's march + deformation forward.
#include <cstdlib>

#include "common.cuh"

namespace seald {

constexpr int kMaxRanks = 16;

struct DpPeers {
    const float* grads[kMaxRanks];
    __half* table16[kMaxRanks];
};

__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}

__device__ __forceinline__ void mc_st_b128(void* mc, const uint4 u) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(u.x)),
                 "f"(__uint_as_float(u.y)), "f"(__uint_as_float(u.z)), "f"(__uint_as_float(u.w))
                 : "memory");
}

template <bool MC>
__device__ __forceinline__ float4 reduce_f32x4(const DpPeers& peers, const float* mc_grads, const int world, const size_t i) {
    if constexpr (MC) {
        return mc_ld_reduce_f32x4(mc_grads + i);
    } else {
        // all peer loads in flight before the first add (remote latency is microseconds): fixed unroll, predicated on the world size
        float4 b[kMaxRanks];
#pragma unroll
        for (int r = 0; r < kMaxRanks; r++)
            if (r < world) b[r] = *reinterpret_cast<const float4*>(peers.grads[r] + i);  // (weak load: ordered by the caller's barrier)
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxRanks; r++)
            if (r < world) { a.x += b[r].x; a.y += b[r].y; a.z += b[r].z; a.w += b[r].w; }
        return a;
    }
}

__device__ __forceinline__ void adam4(float4& p4, float4& m4, float4& v4, const float4 g4, const float inv_scale, const float beta1,
                                      const float beta2, const float eps, const float step_size, const float bc2_sqrt) {
    float* pp = reinterpret_cast<float*>(&p4);
    float* mm = reinterpret_cast<float*>(&m4);
    float* vv = reinterpret_cast<float*>(&v4);
    const float* gg = reinterpret_cast<const float*>(&g4);
#pragma unroll
    for (int k = 0; k < 4; k++) {  // same expression order as k_adam
        const float gi = gg[k] * inv_scale;
        mm[k] = beta1 * mm[k] + (1.0f - beta1) * gi;
        vv[k] = beta2 * vv[k] + (1.0f - beta2) * gi * gi;
        pp[k] = pp[k] - step_size * (mm[k] / (sqrtf(vv[k]) / bc2_sqrt + eps));
    }
}

// ---- phase A: this rank's shard of the summed table gradient ---------------------------------------------------------------
template <bool MC>
__global__ void __launch_bounds__(256) k_dp_reduce_shard(const DpPeers peers, const float* __restrict__ mc_grads, const int world,
                                                         const size_t shard_off, const size_t shard_len, float* __restrict__ out) {
    const size_t tid = threadIdx.x + (size_t)blockIdx.x * blockDim.x, nth = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = shard_len / 4;
    constexpr int U = 4;  // independent 16-byte loads per peer in flight per thread (remote latency is microseconds)
    for (size_t j0 = tid; j0 < n4; j0 += nth * U) {
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MC) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const size_t j = j0 + (size_t)u * nth;
                if (j < n4) acc[u] = mc_ld_reduce_f32x4(mc_grads + shard_off + j * 4);
            }
        } else {
            for (int r = 0; r < world; r++) {
                const float4* src = reinterpret_cast<const float4*>(peers.grads[r] + shard_off);
                float4 b[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const size_t j = j0 + (size_t)u * nth;
                    b[u] = j < n4 ? src[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < U; u++) { acc[u].x += b[u].x; acc[u].y += b[u].y; acc[u].z += b[u].z; acc[u].w += b[u].w; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t j = j0 + (size_t)u * nth;
            if (j < n4) reinterpret_cast<float4*>(out)[j] = acc[u];
        }
    }
}

// ---- phase B1: overflow decision + Adam on the replicated MLP weights (gradient summed over the peers here: 0.5 MB) ------------
template <bool MC>
__global__ void __launch_bounds__(256) k_dp_adam_weights(const DpPeers peers, const float* __restrict__ mc_grads, const int world,
                                                         float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                         const size_t w_off, const size_t n_weights, const size_t flag_off,
                                                         const float lr_net, const float beta1, const float beta2, const float eps,
                                                         const int* __restrict__ step_dev, const float* __restrict__ loss_scale,
                                                         int* __restrict__ found_inf_out) {
    __shared__ float s_bc[2];
    __shared__ int s_skip;
    if (threadIdx.x == 0) {
        const float4 f = reduce_f32x4<MC>(peers, mc_grads, world, flag_off);  // > 0 (or nan) iff some rank saw a non-finite gradient
        const bool skip = !(f.x == 0.0f);
        s_skip = skip ? 1 : 0;
        const double st = (double)max(*step_dev + 1, 1);
        s_bc[0] = (float)(1.0 - pow((double)beta1, st));
        s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, st));
        if (blockIdx.x == 0) *found_inf_out = skip ? 0x3f800000 : 0;
    }
    __syncthreads();
    if (s_skip) return;
    const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
    const float inv_scale = 1.0f / *loss_scale;
    const size_t tid = threadIdx.x + (size_t)blockIdx.x * blockDim.x, nth = (size_t)gridDim.x * blockDim.x;
    const float step_size = lr_net / bc1;
    const size_t n4 = n_weights / 4;
    for (size_t j = tid; j < n4; j += nth) {
        const size_t i = w_off + j * 4;
        const float4 g4 = reduce_f32x4<MC>(peers, mc_grads, world, i);
        float4 p4 = *reinterpret_cast<const float4*>(p + i), m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
        adam4(p4, m4, v4, g4, inv_scale, beta1, beta2, eps, step_size, bc2_sqrt);
        *reinterpret_cast<float4*>(p + i) = p4;
        *reinterpret_cast<float4*>(m + i) = m4;
        *reinterpret_cast<float4*>(v + i) = v4;
    }
}

// ---- phase B2: Adam on this rank's shard of the hash table + its fp16 rows to every rank (8 elements per thread and iteration) ----
template <bool MC>
__global__ void __launch_bounds__(256) k_dp_adam_shard_broadcast(const DpPeers peers, __half* __restrict__ mc_table16, const int world,
                                                                 float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                                 const float* __restrict__ grad_shard, const size_t shard_off,
                                                                 const size_t shard_len, const float lr, const float beta1,
                                                                 const float beta2, const float eps, const int* __restrict__ step_dev,
                                                                 const float* __restrict__ loss_scale, const int* __restrict__ found_inf, const float* __restrict__ lr_scale) {
    if (*found_inf != 0) return;
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
        const double st = (double)max(*step_dev + 1, 1);
        s_bc[0] = (float)(1.0 - pow((double)beta1, st));
        s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, st));
    }
    __syncthreads();
    const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
    const float inv_scale = 1.0f / *loss_scale;
    const size_t tid = threadIdx.x + (size_t)blockIdx.x * blockDim.x, nth = (size_t)gridDim.x * blockDim.x;
    const float step_size = (lr_scale ? lr * *lr_scale : lr) / bc1;  // LambdaLR factor of the step this (deferred) pass belongs to
    const size_t n8 = shard_len / 8;
    for (size_t j = tid; j < n8; j += nth) {
        const size_t i = shard_off + j * 8;
        const float4 ga = *reinterpret_cast<const float4*>(grad_shard + j * 8), gb = *reinterpret_cast<const float4*>(grad_shard + j * 8 + 4);
        float4 pa = *reinterpret_cast<const float4*>(p + i), pb = *reinterpret_cast<const float4*>(p + i + 4);
        float4 ma = *reinterpret_cast<const float4*>(m + i), mb = *reinterpret_cast<const float4*>(m + i + 4);
        float4 va = *reinterpret_cast<const float4*>(v + i), vb = *reinterpret_cast<const float4*>(v + i + 4);
        adam4(pa, ma, va, ga, inv_scale, beta1, beta2, eps, step_size, bc2_sqrt);
        adam4(pb, mb, vb, gb, inv_scale, beta1, beta2, eps, step_size, bc2_sqrt);
        *reinterpret_cast<float4*>(p + i) = pa; *reinterpret_cast<float4*>(p + i + 4) = pb;
        *reinterpret_cast<float4*>(m + i) = ma; *reinterpret_cast<float4*>(m + i + 4) = mb;
        *reinterpret_cast<float4*>(v + i) = va; *reinterpret_cast<float4*>(v + i + 4) = vb;
        const __half2 h0 = __floats2half2_rn(pa.x, pa.y), h1 = __floats2half2_rn(pa.z, pa.w);
        const __half2 h2 = __floats2half2_rn(pb.x, pb.y), h3 = __floats2half2_rn(pb.z, pb.w);
        uint4 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
        u.z = *reinterpret_cast<const uint32_t*>(&h2); u.w = *reinterpret_cast<const uint32_t*>(&h3);
        if constexpr (MC) {
            mc_st_b128(mc_table16 + i, u);
        } else {
            for (int r = 0; r < world; r++) *reinterpret_cast<uint4*>(peers.table16[r] + i) = u;
        }
    }
}

}  // namespace seald

using namespace seald;

static int fill_peers(DpPeers& peers, const void* const* peer_grads, void* const* peer_table16, int world) {
    if (!peer_grads || world < 1 || world > kMaxRanks) return SEALD_E_BADARG;
    for (int r = 0; r < kMaxRanks; r++) {
        peers.grads[r] = r < world ? (const float*)peer_grads[r] : nullptr;
        peers.table16[r] = (r < world && peer_table16) ? (__half*)peer_table16[r] : nullptr;
        if (r < world && !peers.grads[r]) return SEALD_E_BADARG;
        if (r < world && peer_table16 && !peers.table16[r]) return SEALD_E_BADARG;
    }
    return 0;
}

static uint32_t dp_blocks() {
    static int env_blocks = -1;
    if (env_blocks < 0) {
        const char* e = getenv("SEALD_DP_BLOCKS");
        env_blocks = e ? atoi(e) : 0;
    }
    return env_blocks > 0 ? (uint32_t)env_blocks : 4u * SEALD_NUM_SMS;
}

extern "C" int seald_dp_reduce_shard(const void* const* peer_grads, const void* mc_grads, int world, uint64_t shard_off, uint64_t shard_len,
                                     float* grad_shard, seald_stream_t stream) {
    if (!grad_shard || (shard_off | shard_len) % 8) return SEALD_E_BADARG;
    DpPeers peers;
    int rc = fill_peers(peers, peer_grads, nullptr, world);
    if (rc) return rc;
    cudaStream_t st = to_stream(stream);
    if (mc_grads) k_dp_reduce_shard<true><<<dp_blocks(), 256, 0, st>>>(peers, (const float*)mc_grads, world, shard_off, shard_len, grad_shard);
    else k_dp_reduce_shard<false><<<dp_blocks(), 256, 0, st>>>(peers, nullptr, world, shard_off, shard_len, grad_shard);
    return launch_status();
}

extern "C" int seald_dp_adam_weights(const void* const* peer_grads, const void* mc_grads, int world, float* p, float* m, float* v,
                                     uint64_t w_off, uint64_t n_weights, uint64_t flag_off, float lr_net, float beta1, float beta2, float eps,
                                     const int32_t* step_dev, const float* loss_scale, int32_t* found_inf_out, seald_stream_t stream) {
    if (!p || !m || !v || !step_dev || !loss_scale || !found_inf_out) return SEALD_E_BADARG;
    if ((w_off | n_weights | flag_off) % 4) return SEALD_E_BADARG;  // 16-byte vectors throughout
    DpPeers peers;
    int rc = fill_peers(peers, peer_grads, nullptr, world);
    if (rc) return rc;
    cudaStream_t st = to_stream(stream);
    const uint32_t blocks = (uint32_t)div_up<uint64_t>(n_weights / 4 + 1, 256);
    if (mc_grads)
        k_dp_adam_weights<true><<<blocks, 256, 0, st>>>(peers, (const float*)mc_grads, world, p, m, v, w_off, n_weights, flag_off, lr_net, beta1,
                                                        beta2, eps, step_dev, loss_scale, found_inf_out);
    else
        k_dp_adam_weights<false><<<blocks, 256, 0, st>>>(peers, nullptr, world, p, m, v, w_off, n_weights, flag_off, lr_net, beta1, beta2, eps,
                                                         step_dev, loss_scale, found_inf_out);
    return launch_status();
}

extern "C" int seald_dp_adam_shard_broadcast(void* const* peer_table16, void* mc_table16, int world, float* p, float* m, float* v,
                                             const float* grad_shard, uint64_t shard_off, uint64_t shard_len, float lr, float beta1,
                                             float beta2, float eps, const int32_t* step_dev, const float* loss_scale,
                                             const int32_t* found_inf, const float* lr_scale_dev, seald_stream_t stream) {
    if (!peer_table16 || !p || !m || !v || !grad_shard || !step_dev || !loss_scale || !found_inf) return SEALD_E_BADARG;
    if ((shard_off | shard_len) % 8 || world < 1 || world > kMaxRanks) return SEALD_E_BADARG;
    DpPeers peers;
    for (int r = 0; r < kMaxRanks; r++) {
        peers.grads[r] = nullptr;
        peers.table16[r] = r < world ? (__half*)peer_table16[r] : nullptr;
        if (r < world && !peers.table16[r]) return SEALD_E_BADARG;
    }
    cudaStream_t st = to_stream(stream);
    const uint32_t blocks = 4u * SEALD_NUM_SMS;
    if (mc_table16)
        k_dp_adam_shard_broadcast<true><<<blocks, 256, 0, st>>>(peers, (__half*)mc_table16, world, p, m, v, grad_shard, shard_off, shard_len, lr,
                                                                beta1, beta2, eps, step_dev, loss_scale, found_inf, lr_scale_dev);
    else
        k_dp_adam_shard_broadcast<false><<<blocks, 256, 0, st>>>(peers, nullptr, world, p, m, v, grad_shard, shard_off, shard_len, lr, beta1, beta2,
                                                                 eps, step_dev, loss_scale, found_inf, lr_scale_dev);
    return launch_status();
}
