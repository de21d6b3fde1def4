/*
 * seald_b200.h — C-ABI of libseald_b200.so: the sm_100a implementation of SealD-NeRF's
 * dynamic-scene render/train hot path.
 *
 * Every entry point replaces one pybind11 function (or fused group of them) of the reference's
 * native extensions; the reference interface it stands in for is cited per function
 * (paths relative to the reference checkout).
 *
 * Conventions (all functions):
 *   - extern "C", plain pointers + sizes, no torch types.  All pointers are DEVICE pointers owned by
 *     the caller ("caller allocates", like the reference: raymarching.py:205-208, grid.py:47-52).
 *   - never allocate, never synchronise; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 = ok, negative = argument error (SEALD_E_*), positive = cudaError_t of the launch.
 *   - dtype enum: SEALD_F32 = 0, SEALD_F16 = 1 (the reference dispatches on at::ScalarType).
 */
#ifndef SEALD_B200_H_
#define SEALD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEALD_F32 0
#define SEALD_F16 1

#define SEALD_E_BADARG (-1)   /* null pointer / zero size where not allowed            */
#define SEALD_E_UNSUPPORTED (-2) /* D / C / width / degree outside the supported set      */
#define SEALD_E_ALIGN (-3)    /* pointer not aligned for the vector width the kernel uses */

typedef void* seald_stream_t; /* cudaStream_t */

/* Library/ABI version and the SM architecture the kernels were built for (100). */
int seald_version(void);
int seald_sm_arch(void);
/* Human readable text for a negative status code (positive codes: cudaGetErrorString). */
const char* seald_strerror(int status);

/* ------------------------------------------------------------------------------------------------
 * Multiresolution hash / tiled grid encoder.
 * Replaces grid_encode_forward / grid_encode_backward (gridencoder/src/gridencoder.h:12-13,
 * kernels gridencoder/src/gridencoder.cu:88,249,344).
 *   x01      [B, D] fp32, already mapped to [0,1]
 *   table    [offsets[L], C]  dtype (fp16 or fp32); when 16-byte aligned, rows of 4 or 8 bytes are fetched in aligned 16-byte
 *            groups, so the allocation must be readable up to the next 16-byte boundary past its last row (GridEncoder
 *            rounds every level to a multiple of 8 rows, grid.py:125, which guarantees it)
 *   offsets  [L+1] int32
 *   out      [B, L*C] dtype — written directly in the layout GridEncoder.forward returns
 *            (the reference writes [L,B,C] and permutes, grid.py:47,57)
 *   dy_dx    [B, L, D, C] dtype or NULL
 *   gridtype 0 = hash, 1 = tiled; interp 0 = linear, 1 = smoothstep
 * ------------------------------------------------------------------------------------------------ */
/*   b_dev    optional device int32: live row count; rows >= *b_dev are neither read nor written */
int seald_grid_encode_forward(const float* x01, const void* table, const int32_t* offsets, void* out, void* dy_dx,
                              uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                              uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                              const int32_t* b_dev, seald_stream_t stream);

/* grad_out [B, L*C] dtype.  grad_table [offsets[L], C] of grad_table_dtype (fp16: half2 atomics like
 * gridencoder.cu:325-331; fp32: float atomics) is ACCUMULATED into (caller pre-zeroes, grid.py:77).
 * grad_x [B, D] fp32 or NULL.  If grad_x != NULL: uses dy_dx when given (gridencoder.cu:344-369), else
 * recomputes the corner differences from `table` (fp32 accumulation). */
int seald_grid_encode_backward(const void* grad_out, const float* x01, const void* table, const int32_t* offsets,
                               void* grad_table, const void* dy_dx, float* grad_x,
                               uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                               uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                               int grad_table_dtype, const int32_t* b_dev, seald_stream_t stream);

/* Table gradient AND input gradient (recomputed from the table, no dy_dx) of one backward pass in ONE launch where the fast paths
 * apply (fp32 gradient table, C = 2, D 2..4, 16-byte aligned tables: CTA roles inside k_grid_backward_both; otherwise two launches),
 * + the GradScaler overflow flag of seald_grid_encode_backward_table. */
int seald_grid_encode_backward_both(const void* grad_out, const float* x01, const void* table, const int32_t* offsets, void* grad_table,
                                    float* grad_x, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                    int align_corners, uint32_t interp, int dtype, int grad_table_dtype, const int32_t* b_dev,
                                    int32_t* found_inf, seald_stream_t stream);
/* The two halves of seald_grid_encode_backward as separate launches (they are independent: the fused trainer runs the table
 * scatter beside the deformation-net backward on a second stream).  found_inf (optional, device int32): set to a non-zero
 * word (the bits of 1.0f) when a consumed grad_out element is inf/nan; the table gradient is non-finite exactly then, so
 * GradScaler's overflow check (nerf/utils.py:884-886 scaler.step) need not re-read the table gradient. */
int seald_grid_encode_backward_table(const void* grad_out, const float* x01, const int32_t* offsets, void* grad_table,
                                     uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                     uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                                     int grad_table_dtype, const int32_t* b_dev, int32_t* found_inf, seald_stream_t stream);
int seald_grid_encode_backward_input(const void* grad_out, const float* x01, const void* table, const int32_t* offsets,
                                     const void* dy_dx, float* grad_x, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                     float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                                     const int32_t* b_dev, seald_stream_t stream);

/* Debug/parity op: emits the uint32 table row index (before *C) of every (point, level, corner)
 * [B, L, 2^D] and the per-level (scale, resolution) the device computed [L] each. */
int seald_grid_debug_indices(const float* x01, const int32_t* offsets, uint32_t* indices, float* scales,
                             uint32_t* resolutions, uint32_t B, uint32_t D, uint32_t L, float S, uint32_t H,
                             uint32_t gridtype, int align_corners, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Ray marching utilities.  Replace raymarching/src/raymarching.h:7-11.
 * ------------------------------------------------------------------------------------------------ */
int seald_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb6, uint32_t N,
                             float min_near, float* nears, float* fars, seald_stream_t stream);
int seald_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                       seald_stream_t stream);
int seald_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, seald_stream_t stream);
int seald_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, seald_stream_t stream);
/* grid [n_bytes*8] fp32 -> bitfield [n_bytes]; bit i of byte n = grid[8n+i] > thresh. */
int seald_packbits(const float* grid, uint32_t n_bytes, float thresh, uint8_t* bitfield, seald_stream_t stream);

/* Training march.  Replaces march_rays_train (raymarching.h:13, raymarching.cu:312-490).
 * Same contract: xyzs/dirs [M,3], deltas [M,2] (caller pre-zeroes), rays [N,3] = (ray id, offset, count),
 * counter[2] += (samples, rays).  nears/fars may be NULL: then aabb6 must be given and the slab test of
 * near_far_from_aabb (raymarching.cu:92-145) is fused in (optionally also written to nears_out/fars_out).
 * Offsets are reserved with one atomic per CTA after a block prefix sum, so rays of one CTA are packed in
 * ray order (the reference uses one atomic per ray, raymarching.cu:405-406). */
int seald_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* bitfield, float bound,
                           float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                           const float* nears, const float* fars, const float* aabb6, float min_near,
                           float* nears_out, float* fars_out,
                           float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                           const float* noises, const float* occ_aabb6, seald_stream_t stream);
/* World-space box of all occupied cells of `bitfield` (all cascades), grown by `guard_cells` (>= 1) cells; scratch: 6*C int32.
 * Passing it to the march functions as occ_aabb6 (optional, NULL = off) lets rays that cannot meet an occupied cell end
 * without walking the empty volume, and every walk stop where the ray leaves the box.  Only probes of empty cells are
 * removed: counts, positions and deltas are unchanged.  Must be recomputed whenever the bitfield changes. */
int seald_occupancy_aabb(const uint8_t* bitfield, uint32_t C, uint32_t H, float bound, int32_t guard_cells, int32_t* scratch,
                         float* aabb6, seald_stream_t stream);

/* Replaces composite_rays_train_forward/backward (raymarching.h:14-15, raymarching.cu:501-693). */
int seald_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                       const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                       float* weights_sum, float* depth, float* image, seald_stream_t stream);
int seald_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                        const float* rgbs, const float* deltas, const int32_t* rays,
                                        const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                        float T_thresh, float* grad_sigmas, float* grad_rgbs, seald_stream_t stream);

/* composite forward + background blend / MSE loss + composite backward of a training step in one kernel (the arithmetic of
 * seald_composite_rays_train_forward + seald_mse_loss_bg + seald_composite_rays_train_backward): also writes the zeros of
 * grad_sigmas / grad_rgbs behind each ray's early stop, so the caller does not pre-zero them.  bg [N,3] or NULL (white),
 * gt [N,3]; loss_sum += mean squared error * (3N * inv_count); gradients are those of loss_scale * that loss. */
int seald_composite_train_loss_fused(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays, uint32_t M,
                                     uint32_t N, float T_thresh, const float* bg, const float* gt, float inv_count,
                                     const float* loss_scale, float* weights_sum, float* depth, float* image, float* pred,
                                     float* loss_sum, float* grad_sigmas, float* grad_rgbs, seald_stream_t stream);

/* Inference march / composite.  Replace march_rays / composite_rays (raymarching.h:17-18,
 * raymarching.cu:701-914).  n_alive_dev / n_step_dev (optional, may be NULL): device int32 holding the live ray count and
 * the samples per ray of this round; when given they override the host values inside the kernel (the host n_alive is then
 * only the launch bound), which lets run_cuda loop without the host-synced boolean-mask compaction and `n_step` choice
 * (dnerf/renderer.py:353-372).  seald_march_rays writes delta = 0 terminator slots itself (no pre-zeroing needed). */
int seald_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                     const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t C, uint32_t H, const uint8_t* bitfield, const float* nears, const float* fars,
                     float* xyzs, float* dirs, float* deltas, const float* noises, const int32_t* n_alive_dev,
                     const int32_t* n_step_dev, const float* occ_aabb6, seald_stream_t stream);
int seald_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                         const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                         float* depth, float* image, const int32_t* n_alive_dev, const int32_t* n_step_dev,
                         seald_stream_t stream);

/* composite_rays + the compaction `rays_alive = rays_alive[rays_alive >= 0]` + the loop bookkeeping of run_cuda
 * (dnerf/renderer.py:349-376: n_alive, n_step = clamp(N / n_alive, 1, 8), step += n_step) in ONE launch for a device-resident render
 * loop.  n_alive / n_step are launch bounds; the live values are state[0] / state[1].  Survivors are appended to next_alive (any
 * order: a ray's result does not depend on its position) and the last CTA rewrites state[0..5] = {n_alive, n_step of the next round
 * = clamp(budget / n_alive, 1, max_n_step), n_alive * n_step, steps marched, sample rows evaluated, non-empty rounds}; n_alive becomes 0
 * once max_steps is reached.  counters2: two zero-initialised ints (left zero). */
int seald_composite_rays_compact(uint32_t n_alive, uint32_t n_step, float T_thresh, const int32_t* rays_alive, float* rays_t,
                                 const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                                 float* image, int32_t* next_alive, int32_t* state, int32_t* counters2, uint32_t budget, uint32_t max_steps,
                                 uint32_t max_n_step, seald_stream_t stream);
/* Device-side schedule of the render loop (dnerf/renderer.py:350-376).  state[8] int32 = {n_alive, n_step, n_alive*n_step,
 * steps done, samples evaluated so far, non-empty rounds so far, -, -}: adds the finished round's n_step to the step counter, takes the compacted count *n_alive_new (0 once
 * max_steps is reached) and derives the next round's n_step = clamp(N / n_alive, 1, max_n_step) (the reference uses 8). */
int seald_render_schedule(int32_t* state, const int32_t* n_alive_new, uint32_t N, uint32_t max_steps, uint32_t max_n_step,
                          seald_stream_t stream);
/* Order-preserving compaction of rays_alive (entries >= 0 kept): out[0..*n_out) ; replaces
 * `rays_alive = rays_alive[rays_alive >= 0]` (dnerf/renderer.py:372).  scratch: >= ceil(n/1024)+1 int32. */
int seald_compact_alive(const int32_t* rays_alive, uint32_t n_alive, const int32_t* n_alive_dev, int32_t* out,
                        int32_t* n_out_dev, int32_t* scratch, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Seal editing proxy mapping (SealNeRF/seal_utils.py runtime: SealMapper.map_mask :132-153,
 * SealBBoxMapper.map_to_origin :244-286, SealBrushMapper.map_to_origin :415-461,
 * SealAnchorMapper.map_to_origin :522-578, SealMapper.map_color :48-81).
 * The descriptor mirrors the tensors the reference keeps in `map_data` / `map_triangles` / `map_test_dir`;
 * the struct itself lives in HOST memory, its pointer members are DEVICE arrays.
 * ------------------------------------------------------------------------------------------------ */
#define SEALD_SEAL_BBOX 0
#define SEALD_SEAL_BRUSH 1
#define SEALD_SEAL_ANCHOR 2
#define SEALD_SEAL_ATT_LINEAR 0
#define SEALD_SEAL_ATT_DRY 1

typedef struct {
    int32_t type;              /* SEALD_SEAL_*                                                        */
    int32_t n_bounds;          /* B of map_bound [B,2,3] (min row, max row)                           */
    int32_t n_tris;            /* F of map_triangles [F,3,3]                                          */
    int32_t n_border;          /* brush: P of border_points [P,3]                                     */
    const float* bounds;       /* device [B,2,3]                                                      */
    const float* tris;         /* device [F,3,3]                                                      */
    const float* border;       /* device [P,3] or NULL                                                */
    float test_dir[3];         /* map_test_dir, or trimesh's magic direction (seal_utils.py:686-688)  */
    float transform[12];       /* bbox: rows 0..2 of map_data['transform'] (inverse 4x4, row-major)   */
    float rotation[9];         /* bbox: map_data['rotation'] (row-major)                              */
    float scale[3];            /* bbox / anchor: map_data['scale']                                    */
    float center[3];           /* bbox: from_center; brush: plane point                               */
    int32_t has_map_source;    /* bbox: 'map_source' in map_data                                      */
    float empty_bound[6];      /* bbox: map_data['empty_bound'] (min xyz, max xyz)                    */
    float map_source[3];
    float normal_expand[3];    /* brush                                                               */
    float attenuation_distance;
    int32_t attenuation_mode;  /* SEALD_SEAL_ATT_*                                                    */
    float v_anchor[3], v_offset[3], v_h[3]; /* anchor                                                 */
    float len_h, radius;
} seald_seal_mapper;

typedef struct {
    int32_t has_hsv, has_rgb, has_image;
    float hsv[3];              /* map_data['hsv']                                                     */
    float rgb[3];              /* map_data['rgb']                                                     */
    float rgb_light_offset;
    int32_t img_h, img_w;
    const float* image;        /* device [H,W,3]                                                      */
    const float* image_mask;   /* device [H,W]                                                        */
    float v_norm[3], v_o[3], v_w[3], v_h[3];
} seald_seal_color;

/* (points', dirs', mask) = mapper.map_to_origin(points, dirs).  points/dirs [M,3] fp32 (dirs may be NULL), outputs may
 * alias the inputs; mask [M] bytes (0/1, torch.bool layout).  m_dev: optional live row count.  scratch: one int32
 * (anchor only: the batch-wide `map_mask.any()` early exit, seal_utils.py:526-528). */
int seald_seal_map_to_origin(const seald_seal_mapper* mapper, const float* points, const float* dirs, uint32_t M,
                             const int32_t* m_dev, float* points_out, float* dirs_out, uint8_t* mask, int32_t* scratch,
                             seald_stream_t stream);
/* rgbs[mask] = mapper.map_color(points[mask], dirs[mask], rgbs[mask]) in place; rgbs [M,3] fp32, points = MAPPED points
 * (only read for the image projection).  scratch: 4 floats (sums for the batch mean of V, seal_utils.py:771-773). */
int seald_seal_map_color(const seald_seal_color* color, const float* points, const uint8_t* mask, float* rgbs, uint32_t M,
                         const int32_t* m_dev, float* scratch, seald_stream_t stream);
/* Inference / training march with the proxy mapping fused in: same contracts as seald_march_rays /
 * seald_march_rays_train, but xyzs/dirs receive the MAPPED samples and mask [M] their map mask (replaces
 * march_rays + seal_mapper.map_to_origin, SealDNeRF/renderer.py:245-253 and :150-158).  bbox and brush mappers. */
int seald_march_rays_seal(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                          const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                          uint32_t C, uint32_t H, const uint8_t* bitfield, const float* nears, const float* fars,
                          float* xyzs, float* dirs, float* deltas, const float* noises, const int32_t* n_alive_dev,
                          const int32_t* n_step_dev, const seald_seal_mapper* mapper, uint8_t* mask, const float* occ_aabb6,
                          seald_stream_t stream);

/* SAMPLE-PACKED inference round (csrc/raymarch.cu k_march_round_pack / k_composite_round_pack; same per-ray numbers as
 * march_rays / composite_rays, raymarching.cu:701-914, different buffer contract).  A round's samples are stored back to back instead of
 * n_step rows per ray (22% of an 800x800 frame's rows are empty terminator slots otherwise): ray_rows[n] = {first row, count} of alive
 * entry n (count -1: the round's buffers were full, the ray was not marched and survives), state[6] = rows written this round = the live
 * count for the field kernels (must be 0 on entry; seald_composite_rays_pack clears it; it exceeds cap_rows in a round that deferred rays:
 * the field kernels clamp to their buffer); state[7] = rays deferred this round (0 on entry), state[2] = rays deferred so far.  state[0..5] as seald_composite_rays_compact.
 * n_alive / n_step are launch bounds (cap_rows >= 128 * n_step).  mapper: optional fused Seal proxy mapping (mask [cap]). */
int seald_march_rays_pack(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                          const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                          const uint8_t* bitfield, const float* fars, float* xyzs, float* dirs, float* deltas,
                          float* noises /* optional, per RAY: applied at a ray's first real march, then cleared — pass it every round */,
                          int32_t* state, uint32_t cap_rows, int32_t* ray_rows /* [n_alive, 2] */,
                          float* stage /* scratch, 3 floats per (alive entry, step): >= 3 * max over rounds of n_alive * n_step */,
                          const seald_seal_mapper* mapper, uint8_t* mask, const float* occ_aabb6,
                          const uint32_t* coarse_bits /* optional (C == 1, H <= 128): seald_occupancy_coarse_bits — empty 8^3 blocks are
                                                         left in one step; the samples are the same chain elements */,
                          seald_stream_t stream);
int seald_occupancy_coarse_bits(const uint8_t* bitfield, uint32_t H, uint32_t* coarse_bits /* (H/8)^3 / 32 words */, seald_stream_t stream);
/* frame prologue of the packed loop in one launch: near_far_from_aabb (raymarching.cu:108-144) for all N rays, rays_t = near, zeroed
 * weights_sum / depth / image, `alive` = the rays whose [near, far] meets the occupied region (any order; the others keep weights_sum 0
 * = background, as if marched), state[0..7] = {n_alive, n_step = clamp(budget / n_alive, n_step_min, max_n_step), 0, ...}. */
int seald_render_init_pack(const float* rays_o, const float* rays_d, const float* aabb6, const float* occ_aabb6, uint32_t N,
                           float min_near, float* nears, float* fars, float* rays_t, float* weights_sum, float* depth, float* image,
                           int32_t* alive, int32_t* state, int32_t* counters2, uint32_t budget, uint32_t n_step_min,
                           uint32_t max_n_step, seald_stream_t stream);
/* frame epilogue of run_cuda (dnerf/renderer.py:378-384) in one launch: image = acc + (1 - weights_sum) * bg (scalar background),
 * depth = clamp(depth - near, 0) / (far - near) when normalize_depth (else raw, SealDNeRF/renderer.py:284), weights_sum copied; any of
 * the outputs may be NULL; out_packed5: rows {r, g, b, depth, weights_sum} (all-gather input of the tile-sharded frame). */
int seald_render_finish(const float* image, const float* weights_sum, const float* depth, const float* nears, const float* fars,
                        uint32_t N, float bg, int normalize_depth, float* out_image, float* out_depth, float* out_ws,
                        float* out_packed5, seald_stream_t stream);
int seald_composite_rays_pack(uint32_t n_alive, float T_thresh, const int32_t* rays_alive, float* rays_t, const float* sigmas,
                              const float* rgbs, const float* deltas, float* weights_sum, float* depth, float* image,
                              int32_t* next_alive, int32_t* state, int32_t* counters2, const int32_t* ray_rows, uint32_t budget,
                              uint32_t max_steps, uint32_t max_n_step, uint32_t cap_rows, seald_stream_t stream);
int seald_march_rays_train_seal(const float* rays_o, const float* rays_d, const uint8_t* bitfield, float bound,
                                float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                const float* nears, const float* fars, const float* aabb6, float min_near,
                                float* nears_out, float* fars_out, float* xyzs, float* dirs, float* deltas, int32_t* rays,
                                int32_t* counter, const float* noises, const seald_seal_mapper* mapper, uint8_t* mask,
                                const float* occ_aabb6, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Small encoders.  Replace freq_encode_forward/backward (freqencoder/src/freqencoder.h:7-10) and
 * sh_encode_forward/backward (shencoder/src/shencoder.h:7-8); SH degree 1..4 (the D-NeRF setting).
 * ------------------------------------------------------------------------------------------------ */
int seald_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs,
                              seald_stream_t stream);
int seald_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg,
                               uint32_t C, float* grad_inputs, seald_stream_t stream);
int seald_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree,
                            float* dy_dx, seald_stream_t stream);
int seald_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree,
                             const float* dy_dx, float* grad_inputs, seald_stream_t stream);

/* trunc_exp (activation.py:5-17): y = exp(x) in fp32; grad_x = grad * exp(clamp(x, -15, 15)). */
int seald_trunc_exp_forward(const float* x, float* y, uint64_t n, seald_stream_t stream);
int seald_trunc_exp_backward(const float* grad, const float* x, float* grad_x, uint64_t n, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused D-NeRF field (tensor-core MLPs, fp16 operands / fp32 accumulation).
 * Replaces the per-layer cuBLAS path of NeRFNetwork.forward / .density (dnerf/network.py:123-208;
 * SealDNeRF/network.py:125-212) and the freq / SH encoder launches feeding it.
 * Weights are passed as HOST arrays of device pointers to fp16 nn.Linear.weight matrices [out][in]
 * (row-major, 16-byte aligned): deform layer 0 zero-padded to [128][80], hidden [128][128], last [3][128];
 * sigma {[64][32], [16][64]}; colour {[64][32] (column 31 zero), [64][64], [3][64]}.
 * m_dev (optional): device int32 with the live sample count; rows >= *m_dev are skipped (they are zero padding
 * whose outputs no ray reads, raymarching.py:205-207).
 * ------------------------------------------------------------------------------------------------ */
int seald_field_deform_forward(const float* xyz, const float* time_dev, const void* const* weights, int n_layers, uint32_t M,
                               const int32_t* m_dev, float bound, int t0_mode, float* deform, float* x01, void* in_buf /*[M,80] f16 or NULL*/,
                               void* fwd_buf /*[n_layers-1,M,128] f16 or NULL*/, seald_stream_t stream);
/* Same operation on the Blackwell tensor cores (tcgen05.mma, accumulators in tensor memory; csrc/field_umma.cu).
 * `packed`: the deformation weights re-laid out by seald_field_umma_pack_deform into the canonical K-major operand tiles
 * ([K/8][N][8] fp16 per layer, seald_field_umma_deform_bytes(n_layers) bytes, 16-byte aligned) so that one bulk copy per
 * layer brings them into shared memory.  Same outputs as seald_field_deform_forward; the optional training buffers hold whole
 * 128-row TILE IMAGES (the shared-memory A tile as it is: [tile][width/8][128][8] fp16, dead rows of a live tile zero):
 * in_buf ceil128(M) x 80, fwd_buf [n_layers-1] x ceil128(M) x 128 halves — the layout seald_mlp_wgrad_umma reads with ld = 0. */
uint64_t seald_field_umma_deform_bytes(int n_layers);
int seald_field_umma_pack_deform(const void* const* weights, int n_layers, void* packed, seald_stream_t stream);
int seald_field_deform_forward_umma(const float* xyz, const float* time_dev, const void* packed, int n_layers, uint32_t M,
                                    const int32_t* m_dev, float bound, int t0_mode, float* deform, float* x01, void* in_buf,
                                    void* fwd_buf, seald_stream_t stream);
/* NeRFNetwork.density (dnerf/network.py:171-208: deformation net -> hash grid -> sigma head) in ONE launch on tcgen05: the
 * deformation kernel's tile pipeline continues past its last layer — the epilogue threads turn the deformed position into the
 * 16 x 2 hash-grid features (gridencoder.cu:100-216) written straight into the A-operand tile, and the sigma head's layers are
 * further tile-layers of the same chain; only xyz (12 B) in and sigma (4 B) out per sample touch HBM.  `packed_sigma`: the
 * sigma weights as operand tiles (seald_field_umma_pack_sigma, seald_field_umma_sigma_bytes(n_sigma) bytes).  D = 3, C = 2,
 * L = 16 only (SEALD_E_UNSUPPORTED otherwise).  sigma = density_scale * exp(h[0]); optional scatter for the occupancy refresh
 * (dnerf/renderer.py:497-499): tmp[indices[i]] = sigma[i] * store_scale.  `sigma` or `tmp` may be NULL, not both. */
uint64_t seald_field_umma_sigma_bytes(int n_sigma);
int seald_field_umma_pack_sigma(const void* const* weights, int n_sigma, void* packed_sigma, seald_stream_t stream);
int seald_field_density_umma(const float* xyz, const float* time_dev, const void* packed, int n_layers, const void* packed_sigma,
                             int n_sigma, const void* table, const int32_t* offsets, uint32_t D, uint32_t C, uint32_t L, float S,
                             uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, uint32_t M, const int32_t* m_dev,
                             float bound, int t0_mode, float density_scale, float* sigma, const int32_t* indices, float store_scale,
                             float* tmp, seald_stream_t stream);

/* Backward (dgrad chain) on tcgen05: the buffers of seald_field_deform_backward as tile images (fwd_buf as written by
 * seald_field_deform_forward_umma; bwd_buf [n_layers-1] x ceil128(M) x 128, gout_buf ceil128(M) x 16 halves); packedT = the transposed
 * weight tiles W_l^T of layers 1..n-1 (seald_field_umma_pack_deform_T, seald_field_umma_deform_bytes_T(n_layers) bytes). */
uint64_t seald_field_umma_deform_bytes_T(int n_layers);
int seald_field_umma_pack_deform_T(const void* const* weights, int n_layers, void* packedT, seald_stream_t stream);
int seald_field_umma_pack_deform_both(const void* const* weights, int n_layers, void* packed, void* packedT, seald_stream_t stream);
int seald_field_deform_backward_umma(const float* grad_x01, const float* time_dev, const void* packedT, int n_layers, uint32_t M,
                                     const int32_t* m_dev, float bound, const void* fwd_buf, void* bwd_buf, void* gout_buf,
                                     seald_stream_t stream);
int seald_field_deform_backward(const float* grad_x01, const float* time_dev /*NULL or device float: t == 0 => zero gradient*/,
                                const void* const* weights, int n_layers, uint32_t M, const int32_t* m_dev,
                                float bound, const void* fwd_buf, void* bwd_buf /*[n_layers-1,M,128] f16*/, void* gout_buf /*[M,16] f16*/,
                                seald_stream_t stream);
int seald_field_heads_forward(const void* feat /*[M,32] f16*/, const float* dirs, const void* const* w_sigma, int n_sigma,
                              const void* const* w_color, int n_color, uint32_t M, const int32_t* m_dev, float density_scale, float* sigma,
                              float* rgb, void* hs /*[M,16] f16 or NULL*/, void* cin /*[M,32] f16*/, void* fwd_s /*[n_sigma-1,M,64]*/,
                              void* fwd_c /*[n_color-1,M,64]*/, seald_stream_t stream);
int seald_field_sigma_forward(const void* feat, const void* const* w_sigma, int n_sigma, uint32_t M, float density_scale, float* sigma,
                              void* geo /*[M,15] f16 or NULL*/, seald_stream_t stream);
int seald_field_heads_backward(const float* grad_sigma, const float* grad_rgb, const float* rgb, const void* hs, const void* const* w_sigma,
                               int n_sigma, const void* const* w_color, int n_color, uint32_t M, const int32_t* m_dev, float density_scale,
                               const void* fwd_s, const void* fwd_c, void* bwd_s, void* bwd_c, void* gout_s /*[M,16]*/, void* gout_c /*[M,16]*/,
                               void* dfeat /*[M,32] f16*/, seald_stream_t stream);

/* The same two kernels with every tensor the weight-gradient GEMMs read saved as 128-row TILE IMAGES ([tile][width/8][128][8] fp16,
 * ceil128(M) rows per layer, dead rows of a live tile zero; see seald_mlp_wgrad_umma): fwd_s, fwd_c, cin and feat_img (a tile-image
 * copy of the input features, [ceil128(M), 32]) in the forward; fwd_s / fwd_c read and bwd_s, bwd_c, gout_s, gout_c written in the
 * backward.  hs and dfeat stay row-major. */
int seald_field_heads_forward_tiled(const void* feat, const float* dirs, const void* const* w_sigma, int n_sigma, const void* const* w_color,
                                    int n_color, uint32_t M, const int32_t* m_dev, float density_scale, float* sigma, float* rgb, void* hs,
                                    void* cin, void* fwd_s, void* fwd_c, void* feat_img, seald_stream_t stream);
int seald_field_heads_backward_tiled(const float* grad_sigma, const float* grad_rgb, const float* rgb, const void* hs,
                                     const void* const* w_sigma, int n_sigma, const void* const* w_color, int n_color, uint32_t M,
                                     const int32_t* m_dev, float density_scale, const void* fwd_s, const void* fwd_c, void* bwd_s, void* bwd_c,
                                     void* gout_s, void* gout_c, void* dfeat, seald_stream_t stream);

/* Hash-grid encoder FUSED into the heads / the density head (one launch instead of seald_grid_encode_forward + seald_field_heads_forward
 * resp. seald_field_sigma_forward; the 32 features never reach HBM): D = 3, C = 2, L = 16, fp16 table (16-byte aligned), same gathers
 * and interpolation order as seald_grid_encode_forward — bit-identical features.  Training (hs != NULL) saves tile images like
 * seald_field_heads_forward_tiled (feat_img required); hs == NULL: inference.  Other grid shapes: SEALD_E_UNSUPPORTED (use the two calls). */
int seald_field_grid_heads_forward(const float* x01, const void* table, const int32_t* offsets, uint32_t D, uint32_t C, uint32_t L, float S,
                                   uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, const float* dirs,
                                   const void* const* w_sigma, int n_sigma, const void* const* w_color, int n_color, uint32_t M,
                                   const int32_t* m_dev, float density_scale, float* sigma, float* rgb, void* hs, void* cin, void* fwd_s,
                                   void* fwd_c, void* feat_img, seald_stream_t stream);
int seald_field_grid_sigma_forward(const float* x01, const void* table, const int32_t* offsets, uint32_t D, uint32_t C, uint32_t L, float S,
                                   uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp, const void* const* w_sigma, int n_sigma,
                                   uint32_t M, float density_scale, float* sigma, void* geo, seald_stream_t stream);

/* Weight gradients dW[N][K] += G[M][N]^T A[M][K] (fp16 in, fp32 atomics out); replaces the CUTLASS split-K GEMMs of
 * ffmlp_backward (ffmlp/src/ffmlp.cu:801-877).  Up to 16 jobs per launch. */
typedef struct {
    const void* G; /* [M][ldg] f16 */
    const void* A; /* [M][lda] f16 */
    float* dW;     /* [n_real][ldw] f32, accumulated into */
    int N, K;      /* GEMM dims padded to multiples of 16 (<= 128) */
    int ldg, lda, ldw;
    int n_real, k_real; /* rows / columns of dW actually written */
} seald_wgrad_job;
int seald_mlp_wgrad(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, seald_stream_t stream);
/* Same contract on the Blackwell tensor cores (csrc/wgrad_umma.cu): tcgen05.mma over MN-major operand tiles, fp32 accumulators in
 * tensor memory, one vector-atomic flush per CTA.  Row-major operands are staged with cp.async.  ldg == lda == 0 declares TILE-IMAGE
 * operands as the tcgen05 deformation kernels save them: [tile][width/8][128 rows][8 halves] per 128-row tile (dead rows of a live
 * tile zero); such a tile is fetched with one cp.async.bulk per operand. */
int seald_mlp_wgrad_umma(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, seald_stream_t stream);
/* The same + GradScaler's overflow flag: *found_inf |= bits of 1.0f when a partial sum added to a weight gradient is not finite (then the
 * accumulated gradient is not finite either): replaces a separate seald_grad_finite_check pass over the weight gradients. */
int seald_mlp_wgrad_umma_flag(const seald_wgrad_job* jobs, int n_jobs, uint32_t M, const int32_t* m_dev, int32_t* found_inf,
                              seald_stream_t stream);

/* FFMLP-compatible fused MLP.  Replace ffmlp_forward / ffmlp_inference / ffmlp_backward (ffmlp/src/ffmlp.h:8-11).
 * inputs [B,input_dim] f16 (input_dim % 16 == 0, <= 128), weights: flat f16 buffer (ffmlp.cu:632 layout),
 * outputs [B,16] f16, fwd_buf [num_layers,B,hidden] f16 (NULL = inference), hidden in {16,32,64,128},
 * activation 0 (relu) / output_activation 6 (none) as in ffmlp.py:89-96.  Backward accumulates fp32 weight gradients
 * (layout of `weights`) into grad_weights and optionally writes grad_inputs [B,input_dim] f16. */
int seald_ffmlp_forward(const void* inputs, const void* weights, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                        uint32_t num_layers, uint32_t activation, uint32_t output_activation, void* fwd_buf, void* outputs,
                        seald_stream_t stream);
int seald_ffmlp_backward(const void* grad, const void* inputs, const void* weights, const void* fwd_buf, uint32_t B, uint32_t input_dim,
                         uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                         void* bwd_buf, void* grad_inputs, float* grad_weights, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training-step glue (replaces the torch element-wise kernels around the hot path: dnerf/utils.py:74-85 loss,
 * torch.cuda.amp.GradScaler + torch.optim.Adam as configured in main_dnerf.py:129,136).
 * ------------------------------------------------------------------------------------------------ */
/* Start of a step: t_idx = clamp(floor(*time * T), 0, T-1) (dnerf/renderer.py:285); copies frame t_idx of bitfield_all
 * [T, frame_bytes] (and row t_idx of occ_all [T,6], optional) into the step's static buffers and zeroes counter[2] (optional). */
int seald_select_frame(const float* time_dev, uint32_t T, const uint8_t* bitfield_all, uint32_t frame_bytes, uint8_t* bitfield_out,
                       const float* occ_all, float* occ_out, int32_t* counter, seald_stream_t stream);
/* Start of a fused training step in one launch: seald_select_frame + *loss_sum = 0 (optional) + the step's perturbation noise
 * (optional: noises [n_noise] uniform in [0,1), what the reference draws with torch.rand per march, raymarching.py:179-181; a
 * counter-based hash of (noise_ctr2[0], index) so that CUDA-graph replays draw fresh numbers; noise_ctr2 = two zero-initialised
 * uint64, [0] advances by one per launch). */
int seald_step_begin(const float* time_dev, uint32_t T, const uint8_t* bitfield_all, uint32_t frame_bytes, uint8_t* bitfield_out,
                     const float* occ_all, float* occ_out, int32_t* counter, float* loss_sum, float* noises, uint32_t n_noise,
                     uint64_t* noise_ctr2, seald_stream_t stream);
/* pred = image + (1 - ws) * bg (bg NULL = white); loss_sum += mean squared error (inv_count = 1/(3N) or 1/(3N*world));
 * grad_image / grad_ws = d(loss_scale * mse)/d(image, ws). */
int seald_mse_loss_bg(const float* image, const float* weights_sum, const float* bg, const float* gt, uint32_t N, float inv_count,
                      const float* loss_scale, float* pred, float* loss_sum, float* grad_image, float* grad_ws, seald_stream_t stream);
/* Seal local pre-training loss (SealNeRF/trainer.py:448-462, pretrain_step): loss_sum += L1Loss(sigma, gt_sigma) +
 * L1Loss(rgb, gt_rgb) ('mean' over M resp. 3M elements); grad_sigma [M] / grad_rgb [M,3] = d(loss_scale * loss)/d(sigma, rgb). */
int seald_l1_pretrain_loss(const float* sigma, const float* rgb, const float* gt_sigma, const float* gt_rgb, uint32_t M,
                           const float* loss_scale, float* loss_sum, float* grad_sigma, float* grad_rgb, seald_stream_t stream);
/* dst[rows][ld] f16 = src[rows][cols] f32, zero padded columns. */
int seald_cast_pad_f16(const float* src, void* dst, uint32_t rows, uint32_t cols, uint32_t ld, seald_stream_t stream);
/* n matrices in one launch; src/dst/rows/cols/ld are HOST arrays of length n (<= 32). */
int seald_cast_pad_f16_batch(const void* const* src, void* const* dst, const uint32_t* rows, const uint32_t* cols, const uint32_t* ld, int n,
                             seald_stream_t stream);
int seald_grad_finite_check(const float* g, uint64_t n, int32_t* found_inf, seald_stream_t stream);
/* step_dev (optional device int32) overrides `step`: the step counter lives on the device so the optimiser can be replayed
 * from a CUDA graph; seald_adam_advance increments it unless *found_inf (GradScaler skips optimizer.step() then). */
int seald_adam_advance(int32_t* step_dev, const int32_t* found_inf, seald_stream_t stream);
int seald_adam_step(float* p, float* g, float* m, float* v, uint64_t n, float lr, float beta1, float beta2, float eps, uint32_t step,
                    const int32_t* step_dev, const float* loss_scale, const int32_t* found_inf, void* p16, int zero_grad,
                    seald_stream_t stream);
/* Same with a cap on the grid (0 = default): a pass that shares the GPU with other kernels (the trainer runs the table pass
 * beside the next step's march) must not fill every SM, or the latency-bound kernel next to it cannot be scheduled. */
int seald_adam_step_ex(float* p, float* g, float* m, float* v, uint64_t n, float lr, float beta1, float beta2, float eps, uint32_t step,
                       const int32_t* step_dev, const float* loss_scale, const int32_t* found_inf, void* p16, int zero_grad,
                       uint32_t max_blocks, seald_stream_t stream);
/* step_dev (optional): completed-update counter, incremented here when no inf was found (then seald_adam_advance is not
 * needed and seald_adam_step is called with step = 1: "this is update number *step_dev + 1"). */
int seald_loss_scale_update(float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth, float backoff, int interval,
                            int32_t* step_dev, seald_stream_t stream);
/* Same, and also records {found_inf, *step_dev, bits of *loss_scale} (before they change) into stash[0..2]: the values the
 * deferred part of this step's optimiser (the hash-table pass, run beside the next step's march) must use. */
int seald_loss_scale_update_stash(float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth, float backoff,
                                  int interval, int32_t* step_dev, int32_t* stash, seald_stream_t stream);

/* The learning rate as lr * *lr_scale_dev (NULL: lr): the LambdaLR factor of main_dnerf.py:134 lives on the device so a
 * graph-captured step follows the schedule without being recaptured.  Otherwise seald_adam_step_ex. */
int seald_adam_step_lr(float* p, float* g, float* m, float* v, uint64_t n, float lr, const float* lr_scale_dev, float beta1, float beta2,
                       float eps, uint32_t step, const int32_t* step_dev, const float* loss_scale, const int32_t* found_inf, void* p16,
                       int zero_grad, uint32_t max_blocks, seald_stream_t stream);

/* The MLP-weight tail of a training step in ONE launch (csrc/optim_tail.cu): overflow check of the MLP gradients (ORed into
 * *found_inf, which the table scatter may already have raised) -> grid barrier -> torch.optim.Adam on the weights
 * (p, g, m, v point at the MLP region; skipped on overflow; gradients cleared) with the refreshed values written as fp16 into the
 * row-padded staging copy and the tcgen05 operand tiles (seald_field_umma_pack_deform* layouts) -> GradScaler.update()
 * (nerf/utils.py:884-886) -> lr_scheduler.step() (factor 0.1 ** min(iter / sched_iters, 1) written to *lr_scale; sched_iters <= 0
 * or lr_scale == NULL: constant).  stash[4] receives {found_inf, step, loss-scale bits, lr-factor bits} as they applied to THIS
 * step, for the hash-table pass (seald_adam_step_lr with the stash as step_dev / loss_scale / found_inf / lr_scale_dev).
 * segs: the weight matrices in buffer order (first = running element offset); sync2: two zero-initialised ints. */
typedef struct seald_tail_seg {
    uint32_t first, rows, cols, ld;
    void* dst16;    /* fp16 staging copy [rows][ld] */
    void* packed;   /* K-major operand tile of a deformation layer ([cols_pad/8][n_pad][8]) or NULL */
    uint32_t n_pad;
    void* packedT;  /* transposed operand tile ([rows_pad/8][128][8]) or NULL */
} seald_tail_seg;
int seald_mlp_tail(float* p, float* g, float* m, float* v, const seald_tail_seg* segs, int n_segs, float lr, float beta1, float beta2,
                   float eps, int32_t* step_dev, float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth,
                   float backoff, int interval, int32_t* stash, float* lr_scale, int32_t* sched_step, int sched_iters, int32_t* sync2,
                   seald_stream_t stream);

/* seald_mlp_tail for the data-parallel trainer: the gradient of element i is the sum over the ranks' peer-mapped gradient buffers
 * (peer_grads[r] + w_off + i, summed in rank order), the step overflows when some rank's flag (peer_grads[r][flag_off], a float) is
 * set or a summed gradient is not finite; the decision is kept in the rank-local word *found_local (zero-initialised, reset here).
 * Local gradients and flags are not cleared (peers may still read them).  p / m / v: this rank's MLP region.  flags_only != 0: every rank
 * has already flagged its own gradients (table scatter + seald_mlp_wgrad_umma_flag), the decision is the sum of the flags alone and the kernel
 * runs without its check pass and grid barrier. */
int seald_mlp_tail_dp(const void* const* peer_grads, int world, uint64_t w_off, uint64_t flag_off, int flags_only, int32_t* found_local, float* p, float* m,
                      float* v, const seald_tail_seg* segs, int n_segs, float lr, float beta1, float beta2, float eps, int32_t* step_dev,
                      float* loss_scale, int32_t* growth_tracker, float growth, float backoff, int interval, int32_t* stash, float* lr_scale,
                      int32_t* sched_step, int sched_iters, int32_t* sync2, seald_stream_t stream);

/* seald_mlp_tail + the hash-table pass of the same optimiser step (seald_adam_step_lr over table_p / g / m / v [table_n], lr table_lr,
 * fp16 copy table_p16, gradient cleared) in ONE launch: every CTA takes part in the overflow check, then runs its share of Adam over the
 * table next to the MLP weights; GradScaler.update / lr_scheduler.step by the last CTA.  The whole
 * scaler.step(optimizer); scaler.update(); lr_scheduler.step() of nerf/utils.py:884-889. */
int seald_optimizer_step(float* p, float* g, float* m, float* v, const seald_tail_seg* segs, int n_segs, float lr, float beta1, float beta2,
                         float eps, int32_t* step_dev, float* loss_scale, int32_t* found_inf, int32_t* growth_tracker, float growth,
                         float backoff, int interval, int32_t* stash, float* lr_scale, int32_t* sched_step, int sched_iters, int32_t* sync2,
                         float* table_p, float* table_g, float* table_m, float* table_v, uint64_t table_n, float table_lr, void* table_p16,
                         int flags_final /* 1: *found_inf is final already (scatter + seald_mlp_wgrad_umma_flag): no check pass / barrier */,
                         seald_stream_t stream);

/* torch_ema.ExponentialMovingAverage.update: shadow -= (1 - decay) * (shadow - param) (ema_decay = 0.95, main_dnerf.py:136;
 * once per epoch, nerf/utils.py:909-910). */
int seald_ema_update(float* shadow, const float* param, uint64_t n, float decay, seald_stream_t stream);


/* Measurement kernel (csrc/umma_probe.cu): cycles one SM needs for `iters` back-to-back tcgen05.mma 128 x n_cols x 16 (fp16 -> fp32)
 * with the operands placed as `mode` says (0 smem/smem no-swizzle, 1 smem/smem 128-byte swizzle, 2 A in tensor memory + B no-swizzle,
 * 3 A in tensor memory + B swizzled; modes 2/3: n_cols <= 64), issued by `issuers` (1..4) warps at once, each into its own accumulator.  out_cycles[2 * sm] = issue time, [2 * sm + 1] = time until the last MMA completed. */
int seald_umma_probe(int mode, int iters, int n_cols, int issuers, int64_t* out_cycles, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel exchange fused with the optimiser over NVLink peer memory (no counterpart in the reference, which is
 * single-GPU: this is the "hash-table and MLP gradients summed over the GPUs" step of BASELINE.json's north_star,
 * SURVEY.md §8e).  peer_grads[r] / peer_table16[r]: rank r's flat fp32 gradient buffer and fp16 table mapped into this
 * process (symmetric memory); mc_*: NVSwitch multicast mappings of the same buffers, or NULL.  The calling rank owns table
 * elements [shard_off, shard_off + shard_len).
 *   seald_dp_reduce_shard   grad_shard[0..shard_len) = sum over the ranks of grads[shard_off ..) (peer loads, or in-switch
 *                           reduction with multicast).
 *   seald_dp_adam_weights   grads[flag_off] (a float, > 0 = overflow) is summed over the ranks first: if set, nothing is
 *                           updated (GradScaler.step) and found_inf_out says so.  Otherwise Adam runs on the local fp32
 *                           p/m/v of the replicated MLP region [w_off, w_off + n_weights), its gradient summed over the peers.
 *   seald_dp_adam_shard_broadcast  Adam on the local fp32 p/m/v of the shard (gradient = grad_shard, skipped when
 *                           *found_inf != 0) and the shard's fp16 rows stored into EVERY rank's table.  step_dev / loss_scale /
 *                           found_inf are the values stashed by seald_loss_scale_update_stash: the trainer runs this
 *                           kernel at the beginning of the NEXT step, beside the march.
 * The caller orders the ranks with barriers (see csrc/dp_fused.cu).
 * ------------------------------------------------------------------------------------------------ */
int seald_dp_reduce_shard(const void* const* peer_grads, const void* mc_grads, int world, uint64_t shard_off,
                          uint64_t shard_len, float* grad_shard, seald_stream_t stream);
int seald_dp_adam_weights(const void* const* peer_grads, const void* mc_grads, int world, float* p, float* m, float* v,
                          uint64_t w_off, uint64_t n_weights, uint64_t flag_off, float lr_net, float beta1, float beta2, float eps,
                          const int32_t* step_dev, const float* loss_scale, int32_t* found_inf_out, seald_stream_t stream);
int seald_dp_adam_shard_broadcast(void* const* peer_table16, void* mc_table16, int world, float* p, float* m, float* v,
                                  const float* grad_shard, uint64_t shard_off, uint64_t shard_len, float lr, float beta1,
                                  float beta2, float eps, const int32_t* step_dev, const float* loss_scale,
                                  const int32_t* found_inf, const float* lr_scale_dev /* NULL: lr as given */, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Occupancy-grid refresh on the device (NeRFRenderer.update_extra_state, dnerf/renderer.py:453-555).
 *   seald_occ_cell_points  sample point of cell j: (2 c / (H-1) - 1) * span + (rand * 2 - 1) * half_cell in the reference's fp32
 *                          operation order; c = coords[j] or, when coords == NULL (n = H^3, the full sweep), every cell of
 *                          custom_meshgrid(X, Y, Z) (:477-497) with ITS row of rand3 (row (x H + y) H + z), emitted x-fastest
 *                          (point j = cell (j % H, (j / H) % H, j / H^2)) so that a warp of the encoder that follows works on one
 *                          x-line; indices[j] = Morton code of c (optional).
 *   seald_occ_partial_points  the 2n sample points of the partial pass (:504-518): points [0, n) from the drawn cells rand_coords
 *                          (torch.randint, int64 [n,3]); points [n, 2n) from the rand_mask[j]-th OCCUPIED cell
 *                          (== nonzero(grid > 0)[rand_mask], :509-511), found by binary search in csum = inclusive prefix sum of
 *                          (grid > 0) [n_cells] and decoded with morton3D_invert; indices[2n] = Morton codes; rand3 [2n,3] jitter.
 *   seald_occ_store        tmp[indices[j]] = sigma[j] * density_scale          (`tmp_grid[t, cas, indices] = sigmas`, :499)
 *   seald_occ_ema_max      grid = max(grid * decay, tmp) where grid >= 0 and tmp >= 0 (:541-543); tmp is reset to -1.
 * ------------------------------------------------------------------------------------------------ */
int seald_occ_cell_points(const int32_t* coords, const float* rand3, uint32_t n, uint32_t H, float span, float half_cell,
                          float* xyzs, int32_t* indices, seald_stream_t stream);
int seald_occ_partial_points(const int64_t* rand_coords, const int64_t* rand_mask, const int32_t* csum, uint32_t n_cells,
                             const float* rand3, uint32_t n, uint32_t H, float span, float half_cell, float* xyzs, int32_t* indices,
                             seald_stream_t stream);
int seald_occ_store(const float* sigma, const int32_t* indices, uint32_t n, float density_scale, float* tmp, seald_stream_t stream);
int seald_occ_ema_max(float* grid, float* tmp, uint32_t n, float decay, seald_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Step input generation on the device: get_rays for the sampled pixels of one camera
 * (nerf/utils.py:54-137, N > 0, patch_size 1, no error_map: pixel centres at +0.5, normalised directions rotated by
 * pose[:3,:3], origin pose[:3,3]) + the ground-truth gather / alpha blend of NeRFDataset.collate and Trainer.train_step
 * (dnerf/provider.py:340-343, dnerf/utils.py:61-66).
 *   poses [F,4,4], times [F] (or NULL), images [F,H*W,C] (C = 3 or 4; or NULL) resident tables; frame_dev: device int32
 *   frame index (NULL = 0); inds [N] int64 flat pixel indices h*W+w; bg [N,3] or NULL (white) for C = 4;
 *   outputs rays_o, rays_d, gt [N,3]; time_out [1] = times[frame] (optional).
 * ------------------------------------------------------------------------------------------------ */
int seald_get_rays_gather(const float* poses, const float* times, const float* images, const int32_t* frame_dev,
                          const int64_t* inds, uint32_t N, uint32_t H, uint32_t W, uint32_t C, float fx, float fy, float cx,
                          float cy, const float* bg, float* rays_o, float* rays_d, float* gt, float* time_out,
                          seald_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SEALD_B200_H_ */
