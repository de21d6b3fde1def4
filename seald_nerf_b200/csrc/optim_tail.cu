// optim_tail.cu — everything a training step does after the weight-gradient GEMMs, for the MLP weights, in ONE launch:
//
//     GradScaler.unscale_ / found-inf check  ->  Adam on the 13 nn.Linear weights (main_dnerf.py:129)  ->  fp16 staging copies and
//     the tcgen05 operand tiles of the refreshed weights  ->  GradScaler.update (nerf/utils.py:884-886)  ->  lr_scheduler.step()
//     (LambdaLR 0.1 ** min(iter / iters, 1), main_dnerf.py:134)
//
// The reference runs this as  scaler.step(optimizer); scaler.update(); lr_scheduler.step()  = one multi-tensor Adam + ~10 tiny
// kernels and two host synchronisations (found_inf.item()).  Round 1 of this library used five launches (finite check, Adam, fp16
// cast, tile repack, loss-scale update); they were ~15 us of a 400 us step, all launch latency.  Here:
//
//   phase 0   every thread looks at its share of the MLP gradients; a CTA that sees inf/nan raises the step's overflow flag
//             (the table scatter has already raised it for the table gradient)
//   barrier   grid-wide (atomic arrive + spin; the grid is at most one CTA per SM, so every CTA is resident)
//   phase 1   Adam on the same elements (skipped as a whole on overflow), gradient cleared, and the new value scattered as fp16 to
//             the three places the forward/backward kernels read it from: row-padded staging copy, K-major UMMA tile, and the
//             transposed UMMA tile of the deformation backward — all pure permutations of the weight matrix
//   phase 2   the LAST CTA to finish applies GradScaler.update, advances the step / scheduler counters, computes the next
//             learning-rate factor, stashes {found_inf, step, loss scale} for the hash-table pass that follows, resets the flags
//
// The hash-table Adam pass (k_adam, train.cu) runs after it (or beside the next step's march) with the stashed values.
#include "adam.cuh"
#include "common.cuh"

namespace seald {

constexpr int kTailMaxSegs = 16;
struct TailSeg {
    uint32_t first;      // offset of the matrix inside the MLP region of the flat buffers
    uint32_t rows, cols; // nn.Linear.weight [rows = out][cols = in]
    uint32_t ld;         // row pitch of the fp16 staging copy
    __half* dst16;       // staging copy [rows][ld]
    __half* packed;      // K-major operand tile [cols_pad / 8][n_pad][8] of this layer (deformation net only, else nullptr)
    uint32_t n_pad;      // rows of that tile (128, or 16 for the last layer)
    __half* packedT;     // transposed tile [rows_pad / 8][128][8] (deformation layers 1 .. n-1, else nullptr)
};
struct TailSegs {
    TailSeg s[kTailMaxSegs];
    int n;
    uint32_t total;
};

struct TailState {
    int* step_dev;         // completed optimiser updates
    float* loss_scale;
    int* found_inf;        // bits of a positive float when any gradient overflowed (summed across ranks as a float elsewhere)
    int* growth_tracker;
    int* stash;            // [4]: {found_inf, step, loss-scale bits, -} of THIS step, for the table pass
    float* lr_scale;       // learning-rate factor of the CURRENT step (read), replaced by the next step's (written)
    int* sched_step;       // lr_scheduler.last_epoch
    int* sync;             // [2]: barrier arrivals, finished CTAs
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The hash-table pass of the same optimiser step (optional: p == nullptr): with it the kernel is the WHOLE scaler.step(optimizer) /
// scaler.update() / lr_scheduler.step() of a training step in one launch — every CTA first takes part in the overflow check, then
// runs its grid-stride share of torch.optim.Adam over the table (adam.cuh, same arithmetic as k_adam) next to the MLP weights.
struct TableAdam {
    float *p, *g, *m, *v;
    size_t n;
    float lr;
    __half* p16;
    int flags_final;  // 1: *found_inf is already final (table scatter + weight-gradient flush raised it): no check pass, no grid barrier
};

// Data parallel (PEERS): the gradient of element i is the sum over the ranks' gradient buffers (peer-mapped, read in rank order on
// every rank: identical sums), the overflow decision is "some rank raised its flag (table scatter) or the summed MLP gradient is not
// finite".  The local gradient / flag are NOT cleared here — peers may still be reading them; the trainer clears the buffer behind
// the next step's cross-rank barrier — and the decision lives in a rank-local word (found_local).
constexpr int kTailMaxRanks = 16;
struct TailPeers {
    const float* g[kTailMaxRanks];  // each rank's flat gradient buffer
    int world;
    size_t w_off, flag_off;         // MLP region / overflow flag (a float) inside it
    int* found_local;
    int flags_only;                 // 1: every rank has already raised its flag for its own gradients (table scatter + weight-gradient
                                    // flush): the decision is the sum of the flags alone — no check pass, no grid barrier
};

template <bool PEERS>
__device__ __forceinline__ float tail_grad(const float* __restrict__ g, const TailPeers& pr, const uint32_t i) {
    if constexpr (PEERS) {
        float s = 0.0f;
        for (int r = 0; r < pr.world; r++) s += pr.g[r][pr.w_off + i];
        return s;
    } else {
        return g[i];
    }
}

template <bool TABLE, bool PEERS>
__global__ void __launch_bounds__(256, TABLE ? 4 : 1) k_mlp_tail(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                  const __grid_constant__ TailSegs segs, const TailState st, const float lr,
                                                  const float beta1, const float beta2, const float eps, const float growth, const float backoff,
                                                  const int interval, const int sched_iters, const TableAdam tb, const __grid_constant__ TailPeers pr) {
    int* const found = PEERS ? pr.found_local : st.found_inf;
    const uint32_t n = segs.total;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;

    bool skip;
    if (PEERS && pr.flags_only) {
        // decision = "some rank raised its flag" (each rank flags its own gradients: table scatter + weight-gradient flush); every CTA
        // reads the W flags itself — the same answer everywhere, no check pass, no grid barrier
        __shared__ int s_found;
        if (threadIdx.x == 0) {
            float f = 0.0f;
            for (int r = 0; r < pr.world; r++) f += pr.g[r][pr.flag_off];
            s_found = (f == 0.0f) ? 0 : 0x3f800000;
        }
        __syncthreads();
        skip = s_found != 0;
        if (i0 == 0) *found = s_found;  // (read again in phase 2 only, behind the ticket)
    } else if (TABLE && tb.flags_final) {
        // the overflow flag was raised by the kernels that produced the gradients (seald_grid_encode_backward_* for the table,
        // seald_mlp_wgrad_umma_flag for the weights): nothing to check, nothing to wait for
        skip = ld_acquire(found) != 0;
    } else {
        // ---- phase 0: overflow check over this thread's gradients -------------------------------------------------------------
        bool bad = false;
        for (uint32_t i = i0; i < n; i += stride) bad |= !isfinite(tail_grad<PEERS>(g, pr, i));
        if constexpr (PEERS) {
            if (i0 == 0) {  // the ranks' overflow flags (raised by their table scatters)
                float f = 0.0f;
                for (int r = 0; r < pr.world; r++) f += pr.g[r][pr.flag_off];
                bad |= !(f == 0.0f);
            }
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        if (threadIdx.x == 0) {
            if (any_bad) atomicOr(found, 0x3f800000);
            __threadfence();
            atomicAdd(st.sync, 1);
            while (ld_acquire(st.sync) < (int)gridDim.x) __nanosleep(32);
        }
        __syncthreads();
        skip = ld_acquire(found) != 0;
    }
    // ---- phase 1: Adam + fp16 copies ----------------------------------------------------------------------------------------------
    if (!skip) {
        const double stp = (double)(*st.step_dev + 1);
        const float bc1 = (float)(1.0 - pow((double)beta1, stp));
        const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, stp));
        const float inv_scale = 1.0f / *st.loss_scale;
        const float step_size = (lr * (st.lr_scale ? *st.lr_scale : 1.0f)) / bc1;
        int k = 0;
        for (uint32_t i = i0; i < n; i += stride) {
            while (k + 1 < segs.n && i >= segs.s[k + 1].first) k++;
            const TailSeg& sg = segs.s[k];
            const float gi = tail_grad<PEERS>(g, pr, i) * inv_scale;
            const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
            const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
            const float pi = p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
            m[i] = mi; v[i] = vi; p[i] = pi;
            if constexpr (!PEERS) g[i] = 0.0f;
            const uint32_t j = i - sg.first;
            const uint32_t r = j / sg.cols, c = j - r * sg.cols;
            const __half h = __float2half_rn(pi);
            sg.dst16[(size_t)r * sg.ld + c] = h;
            if (sg.packed) sg.packed[((size_t)(c >> 3) * sg.n_pad + r) * 8 + (c & 7)] = h;
            if (sg.packedT) sg.packedT[((size_t)(r >> 3) * 128 + c) * 8 + (r & 7)] = h;
        }
    } else if constexpr (!PEERS) {
        for (uint32_t i = i0; i < n; i += stride) g[i] = 0.0f;
    }
    if constexpr (TABLE) {
        // the table: step number / loss scale / lr factor of THIS step (phase 2 changes them only after every CTA is done)
        const double stp = (double)(*st.step_dev + 1);
        const float bc1 = (float)(1.0 - pow((double)beta1, stp));
        const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, stp));
        const float lr_eff = st.lr_scale ? tb.lr * *st.lr_scale : tb.lr;
        adam_slab_body(blockIdx.x, gridDim.x, tb.p, tb.g, tb.m, tb.v, tb.n, beta1, beta2, eps, bc2_sqrt, 1.0f / *st.loss_scale, lr_eff / bc1, skip,
                       tb.p16, 1);
    }

    // ---- phase 2: the last CTA closes the step ---------------------------------------------------------------------------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(st.sync + 1, 1) + 1;
        if (done == (int)gridDim.x) {
            const int found_v = *found;
            st.stash[0] = found_v;
            st.stash[1] = *st.step_dev;
            st.stash[2] = __float_as_int(*st.loss_scale);
            st.stash[3] = __float_as_int(st.lr_scale ? *st.lr_scale : 1.0f);
            if (found_v) {
                *st.loss_scale *= backoff;
                *st.growth_tracker = 0;
            } else {
                *st.step_dev += 1;
                const int t = *st.growth_tracker + 1;
                if (t >= interval) { *st.loss_scale *= growth; *st.growth_tracker = 0; }
                else *st.growth_tracker = t;
            }
            if (st.sched_step) {  // lr_scheduler.step() runs every iteration, skipped or not (nerf/utils.py:888-889)
                const int e = *st.sched_step + 1;
                *st.sched_step = e;
                if (st.lr_scale && sched_iters > 0) *st.lr_scale = (float)pow(0.1, fmin((double)e / (double)sched_iters, 1.0));
            }
            *found = 0;
            st.sync[0] = 0;
            st.sync[1] = 0;
            __threadfence();
        }
    }
}

// shadow -= (1 - decay) * (shadow - param): torch_ema.ExponentialMovingAverage.update (the reference's `ema_decay=0.95`,
// main_dnerf.py:136, updated once per epoch, nerf/utils.py:909-910)
__global__ void k_ema_update(float* __restrict__ shadow, const float* __restrict__ param, const size_t n, const float one_minus_decay) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float s = shadow[i];
        shadow[i] = s - one_minus_decay * (s - param[i]);
    }
}

}  // namespace seald

using namespace seald;

static int tail_launch(float* p, float* g, float* m, float* v, const seald_tail_seg* segs, int n_segs, float lr, float beta1, float beta2,
                       float eps, int32_t* step_dev, float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth,
                       float backoff, int interval, int32_t* stash, float* lr_scale, int32_t* sched_step, int sched_iters,
                       int32_t* sync2, const TableAdam& tb, seald_stream_t stream, const TailPeers* peers = nullptr) {
    if (!p || !g || !m || !v || !segs || n_segs <= 0 || n_segs > kTailMaxSegs) return SEALD_E_BADARG;
    if (!step_dev || !loss_scale || !found_inf || !growth_tracker || !stash || !sync2) return SEALD_E_BADARG;
    TailSegs ts;
    uint32_t total = 0;
    for (int i = 0; i < n_segs; i++) {
        const seald_tail_seg& a = segs[i];
        if (!a.dst16 || a.rows == 0 || a.cols == 0 || a.ld < a.cols || a.first != total) return SEALD_E_BADARG;
        if (a.packedT && (a.cols > 128)) return SEALD_E_UNSUPPORTED;
        ts.s[i] = TailSeg{a.first, a.rows, a.cols, a.ld, (__half*)a.dst16, (__half*)a.packed, a.n_pad, (__half*)a.packedT};
        total += a.rows * a.cols;
    }
    ts.n = n_segs;
    ts.total = total;
    TailState st{step_dev, loss_scale, found_inf, growth_tracker, stash, lr_scale, sched_step, sync2};
    if (tb.p) {
        // with the table pass: 4 CTAs per SM (launch bounds: <= 64 registers), all resident for the grid barrier — the kernel follows the
        // weight-gradient kernel in stream order, so it finds the SMs empty
        if (!tb.g || !tb.m || !tb.v) return SEALD_E_BADARG;
        if ((((uintptr_t)tb.p | (uintptr_t)tb.g | (uintptr_t)tb.m | (uintptr_t)tb.v) & 15) || ((uintptr_t)tb.p16 & 7)) return SEALD_E_ALIGN;
        if (peers) return SEALD_E_UNSUPPORTED;
        k_mlp_tail<true, false><<<4u * SEALD_NUM_SMS, 256, 0, to_stream(stream)>>>(p, g, m, v, ts, st, lr, beta1, beta2, eps, growth, backoff,
                                                                                 interval, sched_iters, tb, TailPeers{});
        return launch_status();
    }
    // at most one CTA per SM: the grid barrier needs every CTA resident
    uint32_t blocks = div_up(total, 256u * 4u);
    if (blocks > (uint32_t)SEALD_NUM_SMS) blocks = SEALD_NUM_SMS;
    if (blocks == 0) blocks = 1;
    if (peers)
        k_mlp_tail<false, true><<<blocks, 256, 0, to_stream(stream)>>>(p, g, m, v, ts, st, lr, beta1, beta2, eps, growth, backoff, interval,
                                                                     sched_iters, tb, *peers);
    else
        k_mlp_tail<false, false><<<blocks, 256, 0, to_stream(stream)>>>(p, g, m, v, ts, st, lr, beta1, beta2, eps, growth, backoff, interval,
                                                                      sched_iters, tb, TailPeers{});
    return launch_status();
}

extern "C" int seald_mlp_tail_dp(const void* const* peer_grads, int world, uint64_t w_off, uint64_t flag_off, int flags_only, int32_t* found_local, float* p,
                                 float* m, float* v, const seald_tail_seg* segs, int n_segs, float lr, float beta1, float beta2, float eps,
                                 int32_t* step_dev, float* loss_scale, int32_t* growth_tracker, float growth, float backoff, int interval,
                                 int32_t* stash, float* lr_scale, int32_t* sched_step, int sched_iters, int32_t* sync2, seald_stream_t stream) {
    if (!peer_grads || world <= 0 || world > kTailMaxRanks || !found_local) return SEALD_E_BADARG;
    TailPeers pr;
    for (int r = 0; r < world; r++) {
        if (!peer_grads[r]) return SEALD_E_BADARG;
        pr.g[r] = (const float*)peer_grads[r];
    }
    pr.world = world; pr.w_off = (size_t)w_off; pr.flag_off = (size_t)flag_off; pr.found_local = found_local; pr.flags_only = flags_only ? 1 : 0;
    // (g: unused in this mode, the sums are read from the peers; found_inf: the local word doubles as the required argument)
    return tail_launch(p, p, m, v, segs, n_segs, lr, beta1, beta2, eps, step_dev, loss_scale, found_local, growth_tracker, growth, backoff, interval,
                       stash, lr_scale, sched_step, sched_iters, sync2, TableAdam{}, stream, &pr);
}

extern "C" int seald_mlp_tail(float* p, float* g, float* m, float* v, const seald_tail_seg* segs, int n_segs, float lr, float beta1, float beta2,
                              float eps, int32_t* step_dev, float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth,
                              float backoff, int interval, int32_t* stash, float* lr_scale, int32_t* sched_step, int sched_iters,
                              int32_t* sync2, seald_stream_t stream) {
    return tail_launch(p, g, m, v, segs, n_segs, lr, beta1, beta2, eps, step_dev, loss_scale, found_inf, growth_tracker, growth, backoff, interval,
                       stash, lr_scale, sched_step, sched_iters, sync2, TableAdam{}, stream);
}

extern "C" int seald_optimizer_step(float* p, float* g, float* m, float* v, const seald_tail_seg* segs, int n_segs, float lr, float beta1,
                                    float beta2, float eps, int32_t* step_dev, float* loss_scale, int32_t* found_inf, int32_t* growth_tracker,
                                    float growth, float backoff, int interval, int32_t* stash, float* lr_scale, int32_t* sched_step,
                                    int sched_iters, int32_t* sync2, float* table_p, float* table_g, float* table_m, float* table_v,
                                    uint64_t table_n, float table_lr, void* table_p16, int flags_final, seald_stream_t stream) {
    if (!table_p || table_n == 0) return SEALD_E_BADARG;
    return tail_launch(p, g, m, v, segs, n_segs, lr, beta1, beta2, eps, step_dev, loss_scale, found_inf, growth_tracker, growth, backoff, interval,
                       stash, lr_scale, sched_step, sched_iters, sync2,
                       TableAdam{table_p, table_g, table_m, table_v, (size_t)table_n, table_lr, (__half*)table_p16, flags_final ? 1 : 0}, stream);
}

extern "C" int seald_ema_update(float* shadow, const float* param, uint64_t n, float decay, seald_stream_t stream) {
    if (n == 0) return 0;
    if (!shadow || !param || !(decay >= 0.0f && decay <= 1.0f)) return SEALD_E_BADARG;
    const uint32_t blocks = (uint32_t)(div_up<uint64_t>(n, 256) < 8ull * SEALD_NUM_SMS ? div_up<uint64_t>(n, 256) : 8ull * SEALD_NUM_SMS);
    k_ema_update<<<blocks, 256, 0, to_stream(stream)>>>(shadow, param, (size_t)n, 1.0f - decay);
    return launch_status();
}
