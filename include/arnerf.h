/*
 * include/arnerf.h -- C ABI of libarnerf.so, the B200 (sm_100a) implementation of the AR-NeRF rendering hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces one function of the reference's pybind11 module `vren`
 * (models/csrc/binding.cpp:234-250) or one tiny-cuda-nn call site of models/networks.py, and is what a
 * maintainer's FFI stub binds (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the library never allocates persistent device memory: all buffers (including scratch) come from the caller;
 *   - every kernel is launched on the `stream` argument (a cudaStream_t passed as void*), never on the legacy stream,
 *     so calls are CUDA-graph capturable and re-entrant;
 *   - return value: 0 on success, negative ARN_E_* otherwise; arn_last_error() returns a thread-local message;
 *   - outputs are written completely by the kernels (no pre-zeroing by the caller is needed) unless stated.
 *   - "in place" tensors keep the reference's in-place semantics: density_bitfield, hits_t (test march),
 *     alive_indices, opacity/depth/rgb (test compositing).
 *   - one process per GPU, one calling thread per device: the side stream / event pool of the pipelined field evaluation, the
 *     arn_train_set_* switches (thread-local) and the peer-exchange settings are per process, not per calling stream.  Kernel
 *     attributes that depend on the device (shared-memory limits of the MLP kernels) are set per device on first use.
 */
#ifndef ARNERF_H_
#define ARNERF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARN_VERSION 200 /* round 2 */

#define ARN_OK 0
#define ARN_E_INVALID (-1) /* bad argument (null pointer, negative size, unsupported configuration) */
#define ARN_E_CUDA (-2)    /* a CUDA runtime call or kernel launch failed */

typedef void* arn_stream_t; /* cudaStream_t */

int arn_version(void);
const char* arn_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
int64_t arn_launch_count(void);
/* Per-kernel device timing for bench.py's roofline numbers: while enabled, every kernel launch of this library is
 * bracketed by CUDA events on its launch stream.  arn_profile_report synchronises, writes one "kernel calls total_ms"
 * line per kernel into buf_host and clears the log.  Keep it disabled during CUDA-graph capture. */
/* Kernel-variant switch for A/B measurements and for the tests that hold the variants to each other; every variant of
 * a kernel computes the same results.  Names: "march_warp" (1 = warp-per-ray window march, 0 = thread-per-ray loop),
 * "hash_bw_mode" (8/16/32/64 = run-walking hash-grid backward, the value being the SHORTEST run of consecutive samples a lane
 * group walks -- the samples are cut into equal runs over one resident wave of blocks; 16 is the default; 1 = warp-segmented
 * backward: a lane per sample, equal cells of neighbouring lanes merged by a segmented warp scan -- measured 98 vs 79 us on the
 * W1 batch; 0 = one reduction per sample and corner), "hash_bw_blocks" (blocks per SM of that wave, 0 = what the occupancy calculator allows), "adam_vec" (1 = 128-bit Adam kernel), "pipeline_parts" (field evaluations of >= 64 K samples are split
 * into this many consecutive tile ranges, hash-grid and MLP kernels of neighbouring ranges overlapped on two streams; default
 * 1 = off: measured on B200 the overlap loses to the per-launch fixed costs, 0.384 / 0.409 / 0.469 ms per step at 1 / 2 / 3),
 * "mlp_wide" (bit 0: forward MLP with four warpgroups per CTA sharing one weight image, bit 1: backward with three; default 3),
 * "pdl" (1 = programmatic dependent launch along the serial kernel chains; default 0: measured slower, see arn_common.cuh). */
int arn_set_tunable(const char* name, int value);
/* Measurement aid (bench.py): n_reductions red.global.add.v4.f32 to pseudo-random 16-byte slots of buf (n_floats floats, 16-byte
 * aligned) and nothing else -- timed by the caller, it is the MEASURED rate at which the L2 retires scattered 16-byte
 * reductions, the roof the hash-grid backward is reported against. */
int arn_dbg_l2_red_peak(float* buf, int64_t n_floats, int64_t n_reductions, arn_stream_t stream);
int arn_profile_enable(int on);
int arn_profile_report(char* buf_host, int capacity);

/* ------------------------------------------------------------------------------------------------------------
 * Ray / box and ray / sphere intersection.
 * Replaces vren.ray_aabb_intersect   (binding.cpp:4-17   -> intersection.cu:59-100, kernel :25-56)
 *          vren.ray_sphere_intersect (binding.cpp:19-32  -> intersection.cu:156-197, kernel :124-153).
 * hit_cnt (R) i32, hits_t (R,max_hits,2) f32, hits_idx (R,max_hits) i64; unfilled slots are -1 and, exactly as the
 * reference's torch::sort on t1 leaves them, come FIRST.  One thread per ray loops over the voxels (deterministic).
 * ---------------------------------------------------------------------------------------------------------- */
int arn_ray_aabb_intersect(const float* rays_o, const float* rays_d, int64_t n_rays,
                           const float* centers, const float* half_sizes, int n_voxels, int max_hits,
                           int32_t* hit_cnt, float* hits_t, int64_t* hits_idx, arn_stream_t stream);
int arn_ray_sphere_intersect(const float* rays_o, const float* rays_d, int64_t n_rays,
                             const float* centers, const float* radii, int n_spheres, int max_hits,
                             int32_t* hit_cnt, float* hits_t, int64_t* hits_idx, arn_stream_t stream);
/* Fused fast path used by render(): single box, max_hits = 1, plus the near-plane clamp of rendering.py:29-31
 * (0 <= t1 < near  ->  t1 = near).  hits_t (R,1,2).  Pass near < 0 to skip the clamp. */
int arn_ray_aabb_near(const float* rays_o, const float* rays_d, int64_t n_rays, const float* center_host,
                      const float* half_size_host, float near, float* hits_t, arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Occupancy-grid utilities.
 * Replaces vren.morton3D (binding.cpp:46-50 -> raymarching.cu:72-88), vren.morton3D_invert (:53-57 -> :103-119),
 *          vren.packbits (:34-43 -> raymarching.cu:143-162, kernel :122-141).
 * grid_dtype: 0 = float32, 1 = float16, 2 = float64 (the reference dispatches on the same three).
 * ---------------------------------------------------------------------------------------------------------- */
int arn_morton3d(const int32_t* coords, int64_t n, int32_t* indices, arn_stream_t stream);
int arn_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords, arn_stream_t stream);
int arn_packbits(const void* density_grid, int grid_dtype, float threshold, uint8_t* density_bitfield,
                 int64_t n_bytes, arn_stream_t stream);

/* Rays of one sampled training batch on the device: replaces `poses[img_idxs]`, `directions[pix_idxs]` and
 * get_rays (train.py:121-126, datasets/ray_utils.py:46-70).  directions (H*W,3) f32 as get_ray_directions builds them,
 * or NULL with K_host (3x3 row-major intrinsics, host) + width to compute ((u-cx+.5)/fx, (v-cy+.5)/fy, 1) from the pixel
 * index (ray_utils.py:33-35).  poses (n_images,3,4) f32 camera-to-world; img_idxs (n) i64 or NULL with img_single (the
 * 'same_image' sampling strategy, datasets/base.py:27-28); pix_idxs (n) i64.  rays_o, rays_d (n,3) f32, rays_d un-normalised. */
int arn_gather_rays(const float* directions, const float* K_host, int width, const float* poses, const int64_t* img_idxs,
                    int64_t img_single, const int64_t* pix_idxs, int64_t n, float* rays_o, float* rays_d, arn_stream_t stream);
/* The whole sampled batch of datasets/base.py:22-36 + train.py:121-126 from the two index vectors: arn_gather_rays plus the
 * pixels `rays[img_idxs, pix_idxs]` of the DEVICE-resident training images (n_images, pixels_per_image, channels) f32
 * (channels = 3 rgb, 4 with HDR-NeRF's exposure) into pixels_out (n, channels) -- so that a training step's host -> device
 * traffic is the 2 x 8 bytes of indices per ray the dataset draws, not rays and colours. */
int arn_gather_batch(const float* directions, const float* K_host, int width, const float* poses, const int64_t* img_idxs,
                     int64_t img_single, const int64_t* pix_idxs, int64_t n, const float* images, int64_t pixels_per_image,
                     int channels, float* rays_o, float* rays_d, float* pixels_out, arn_stream_t stream);

/* Occupancy refresh, the arithmetic of NGP.update_density_grid (networks.py:253-281) around the density evaluation:
 *   arn_grid_cell_positions : xyzs_w = (coords/(G-1)*2-1)*(s - s/G) + (rnd*2-1)*(s/G) for n_cells cells (networks.py:263-267;
 *                             rnd is the caller's torch.rand_like draw, positions are bit-identical with the torch expression);
 *   arn_density_grid_update : density_grid = where(grid < 0, grid, max(grid*decay, tmp)) IN PLACE (decay_cells != NULL: a
 *                             per-cell decay, the `erode` path), threshold = min(mean(grid[grid > 0]), density_threshold)
 *                             accumulated in double on the device, then packbits with that threshold -- no host round trip
 *                             (the reference does .item() here).  scratch: ARN_GRID_UPDATE_SCRATCH_BYTES; the threshold used
 *                             is left in the last 4 bytes of the scratch (float). */
#define ARN_GRID_UPDATE_PARTS 2048
#define ARN_GRID_UPDATE_SCRATCH_BYTES (ARN_GRID_UPDATE_PARTS * 16 + 16)
int arn_grid_cell_positions(const int32_t* coords, const float* rnd, int64_t n_cells, int grid_size, float s, float* xyzs,
                            arn_stream_t stream);
/* Steady-state cell selection of the refresh (networks.py:181-207 sample_uniform_and_occupied_cells + :263-267 positions) for
 * ONE cascade: coords1 (M,3) i32 = the caller's torch.randint(G) draw (M uniform cells), u (M) i64 = the caller's second
 * randint draw (the k-th occupied cell with k = u mod count, count = #(density_grid > density_threshold): the distribution of
 * the reference's nonzero()[randint(count)]), rnd (2M,3) = the caller's torch.rand draw.  Out, in draw order (uniform half
 * first, as the reference concatenates them): indices (2M) i64 morton codes and xyzs (2M,3) jittered positions.  No host
 * round trip, three small launches (occupancy bit masks + chunk prefix instead of cumsum / searchsorted over the grid).
 * scratch: ARN_GRID_SAMPLE_SCRATCH_INTS(G^3) int32, 16-byte aligned.  M: multiple of 256.  grid_size <= 160. */
#define ARN_GRID_SAMPLE_SCRATCH_INTS(n_cells) ((((n_cells) + 1023) / 1024) * 34 + 4)
int arn_grid_sample_cells(const float* density_grid, float density_threshold, int grid_size, float s, const int32_t* coords1,
                          const int64_t* u, int64_t M, const float* rnd, void* scratch, int64_t* indices, float* xyzs,
                          arn_stream_t stream);
/* The same selection with the 2M cells ordered along the morton curve (counting sort over the cells; the draws of one cell are
 * neighbours in arbitrary order, each with the jitter of its own draw): the density evaluation that follows gathers neighbouring
 * cells' hash-grid corners together (1 M cells: 167 us instead of 295).  The sort costs about what it saves, so it pays where the
 * selection is computed AHEAD of the refresh -- it depends on the previous refresh's density_grid and the draws only.
 * scratch: ARN_GRID_SAMPLE_SORTED_SCRATCH_INTS(G^3, M) int32, 16-byte aligned. */
#define ARN_GRID_SAMPLE_SORTED_SCRATCH_INTS(n_cells, M) ((((n_cells) + 1023) / 1024) * 36 + 12 + (n_cells) + 2 * (M))
int arn_grid_sample_cells_sorted(const float* density_grid, float density_threshold, int grid_size, float s, const int32_t* coords1,
                                 const int64_t* u, int64_t M, const float* rnd, void* scratch, int64_t* indices, float* xyzs,
                                 arn_stream_t stream);
/* dst[indices[i]] = src[i], i < n (density_grid_tmp[c, indices] = density, networks.py:268): an index outside [0, n_dst) is
 * skipped, a cell that occurs twice keeps one of its values (as index_put_ does). */
int arn_grid_scatter(float* dst, int64_t n_dst, const int64_t* indices, const float* src, int64_t n, arn_stream_t stream);
int arn_density_grid_update(float* density_grid, const float* density_tmp, const float* decay_cells, float decay,
                            float density_threshold, int64_t n_cells, uint8_t* density_bitfield, void* scratch,
                            arn_stream_t stream);
/* NGP.mark_invisible_cells (networks.py:209-250) for ONE cascade, all cameras in one launch (the reference chunks the cells
 * and materialises (N_cams,3,chunk) tensors): cell i (coords (n_cells,3) i32, indices (n_cells) i64 = morton codes) projects
 * its centre x_w = (coords/(G-1)*2-1)*(s - s/G) through every camera -- w2c (n_cams,12) f32 on the device, per camera the
 * world-to-camera rotation row-major (9) then the translation (3), i.e. poses[:, :3, :3]^T and -R^T t -- and K_host (3x3
 * row-major, host).  count_grid[idx] = (#cameras with the cell inside the image at depth >= near) / n_cams;
 * density_grid[idx] = 0 if count > 0 and no camera has the cell inside the image closer than near, else -1. */
int arn_mark_invisible_cells(const int32_t* coords, const int64_t* indices, int64_t n_cells, int grid_size, float s,
                             const float* w2c, int n_cams, const float* K_host, float img_w, float img_h, float near,
                             float* density_grid, float* count_grid, arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training ray march.  Replaces vren.raymarching_train (binding.cpp:60-81 -> raymarching.cu:283-332, kernel :166-280).
 * The reference's one kernel (count, two atomics, re-march into worst-case R*max_samples buffers) is split:
 *   arn_march_train_count : per-ray sample counts (bit-exact with the reference's pass 1), then an exclusive scan:
 *                           rays_a[r] = (r, start[r], N[r]) (i64, canonical order) and counter = (total, n_rays).
 *   arn_march_train_emit  : re-march into EXACTLY-sized outputs (caller reads counter[0] to size them).
 * hits_t is (R,2).  Sample values are bit-exact with the reference's for the same ray.
 * ---------------------------------------------------------------------------------------------------------- */
int arn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                          const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                          float exp_step_factor, const float* noise, int max_samples,
                          int64_t* rays_a, int32_t* counter, arn_stream_t stream);
int arn_march_train_emit(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                         const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                         float exp_step_factor, const float* noise, int max_samples, const int64_t* rays_a,
                         float* xyzs, float* dirs, float* deltas, float* ts, int64_t capacity, arn_stream_t stream);
/* Same pair with a caller-provided scratch t_scratch (n_rays * max_samples floats, NOT zeroed, only the marched
 * prefix of each row is touched): pass 1 records each sample's t, so pass 2 is one thread per sample (coalesced
 * stores, no second march).  t_scratch == NULL falls back to the re-march of the plain entry points. */
int arn_march_train_count_ex(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                             const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                             float exp_step_factor, const float* noise, int max_samples,
                             int64_t* rays_a, int32_t* counter, float* t_scratch, int32_t* count_scratch,
                             arn_stream_t stream);
int arn_march_train_emit_ex(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                            const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                            float exp_step_factor, const float* noise, int max_samples, const int64_t* rays_a,
                            const float* t_scratch, float* xyzs, float* dirs, float* deltas, float* ts,
                            int64_t capacity, arn_stream_t stream);

/* Test-time march.  Replaces vren.raymarching_test (binding.cpp:84-106 -> raymarching.cu:407-454, kernel :335-404).
 * hits_t (R,2) is updated in place at [r][0]; outputs are padded (n_alive, N_samples, .) and zero-filled past
 * N_eff_samples[n] by the kernel.  Keeps the reference's use of `cascades` as calc_dt's scale (raymarching.cu:370). */
int arn_march_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_indices,
                   int64_t n_alive, const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                   float exp_step_factor, int n_samples, int max_samples,
                   float* xyzs, float* dirs, float* deltas, float* ts, int32_t* n_eff_samples, arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Compositing.  Replaces vren.composite_train_fw (binding.cpp:109-126 -> volumerendering.cu:47-83, kernel :5-44),
 *               vren.composite_train_bw (:129-163 -> :153-201, kernel :86-150),
 *               vren.composite_test_fw  (:166-194 -> :251-284, kernel :204-248).
 * One warp per ray, warp-segmented product/sum scans, early termination at T <= T_threshold.
 * dL_dws may be NULL (no gradient reaches ws).  n_samples = sigmas.numel().
 * ---------------------------------------------------------------------------------------------------------- */
int arn_composite_train_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                           const int64_t* rays_a, int64_t n_rays, int64_t n_samples, float T_threshold,
                           int64_t* total_samples, float* opacity, float* depth, float* rgb, float* ws,
                           arn_stream_t stream);
int arn_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb, const float* dL_dws,
                           const float* sigmas, const float* rgbs, const float* ws, const float* deltas,
                           const float* ts, const int64_t* rays_a, const float* opacity, const float* depth,
                           const float* rgb, int64_t n_rays, int64_t n_samples, float T_threshold,
                           float* dL_dsigmas, float* dL_drgbs, arn_stream_t stream);
int arn_composite_test_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                          int64_t* alive_indices, int64_t n_alive, int n_samples, float T_threshold,
                          const int32_t* n_eff_samples, float* opacity, float* depth, float* rgb, arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Distortion loss.  Replaces vren.distortion_loss_fw (binding.cpp:197-209 -> losses.cu:62-107) and
 *                   vren.distortion_loss_bw (:212-231 -> losses.cu:143-181).
 * ---------------------------------------------------------------------------------------------------------- */
int arn_distortion_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays,
                      int64_t n_samples, float* loss, float* ws_inclusive_scan, float* wts_inclusive_scan,
                      arn_stream_t stream);
int arn_distortion_bw(const float* dL_dloss, const float* ws_inclusive_scan, const float* wts_inclusive_scan,
                      const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays,
                      int64_t n_samples, float* dL_dws, arn_stream_t stream);

/* Per-ray segment sums used by RayMarcher.backward (replaces torch_scatter.segment_csr, custom_functions.py:104-112):
 * dL_drays_o[r] = sum_s dL_dxyzs[s], dL_drays_d[r] = sum_s (dL_dxyzs[s]*ts[s] + dL_ddirs[s]).  dL_ddirs may be NULL. */
int arn_march_train_bw(const float* dL_dxyzs, const float* dL_ddirs, const float* ts, const int64_t* rays_a,
                       int64_t n_rays, float* dL_drays_o, float* dL_drays_d, arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Field: multiresolution hash grid (L=16, F=2) + density MLP 32-64-16 + SH-4 + colour MLP 32-64-64-16.
 * Replaces the tiny-cuda-nn modules built in models/networks.py:37-78 and evaluated in :95-108 / :133-165.
 * Numeric contract ("fp16 operands, fp32 accumulate") is stated in DESIGN.md and oracle/oracle_field.c.
 * ---------------------------------------------------------------------------------------------------------- */
#define ARN_N_LEVELS 16
#define ARN_DENSITY_MLP_PARAMS 3072 /* 64*32 + 16*64            (networks.py:37-57)  */
#define ARN_RGB_MLP_PARAMS 7168     /* 64*32 + 64*64 + 16*64    (networks.py:68-78)  */

/* HOST function: level table in strict float32 (tiny-cuda-nn grid.h rule).
 * scale/res/size have n_levels entries, offset has n_levels+1 (offset[n_levels] = total entries). */
int arn_hashgrid_geometry(int n_levels, int base_resolution, float per_level_scale, int log2_hashmap_size,
                          float* scale_host, uint32_t* res_host, uint32_t* size_host, uint32_t* offset_host);

/* fp32 master parameters -> fp16 working copy (tiny-cuda-nn casts every forward).  n elements. */
int arn_cast_f32_to_f16(const float* src, void* dst_f16, int64_t n, arn_stream_t stream);

/* Level table passed by value to the field kernels (all HOST pointers, ARN_N_LEVELS entries; offset has +1). */
typedef struct {
    const float* scale_host;
    const uint32_t* res_host;
    const uint32_t* size_host;
    const uint32_t* offset_host;
} arn_levels_t;

/* Workspace layout of one field evaluation over n samples (all device buffers owned by the caller).  Every buffer holds
 * m = 128 * ceil(n / 128) rows: the kernels move whole 128-row tiles with bulk (TMA) copies.
 *   feat  (m,32) f16   encoded features                      hid   (m,64) f16   density hidden layer
 *   h     (m,16) f32   density-net output (h[:,0] = log sigma), plain row-major
 *   in32  (m,32) f16   colour-net input [sh16 | h16]         hid1, hid2 (m,64) f16  colour hidden layers
 * feat/hid/in32/hid1/hid2 (and dfeat_scratch (m,32) f32 of the backward) are opaque "tile images": within a row the
 * 16-byte chunks are stored in the order of the shared-memory swizzle of that row width, so that a tile is one contiguous
 * block the tensor-core kernels copy straight into an operand buffer.  They are produced and consumed by this library
 * only.  For a density-only evaluation (NGP.density) dirs, in32, hid1, hid2, rgbs and params_rgb_f16 are NULL.
 * Inference (no backward will follow): hid = in32 = hid1 = hid2 = NULL skips every activation store of the forward
 * (tensor-core path); h = NULL skips the h output as well. */
#define ARN_FIELD_SCRATCH_SLABS 1280
#define ARN_FIELD_SCRATCH_BYTES (20480 + ARN_FIELD_SCRATCH_SLABS * 10240 * 4)
typedef struct {
    void* feat; void* hid; float* h; void* in32; void* hid1; void* hid2;
    /* ARN_FIELD_SCRATCH_BYTES of scratch: [0, 20480) the MLP weights as swizzled tcgen05 operand tiles (rebuilt every
     * call), then one 10240-float slab of weight gradients per backward CTA, summed in slab order after the kernel. */
    void* wimg;
} arn_field_ws_t;

/* Forward.  xyzs (n,3) world coordinates, normalised inside with (x - xyz_min)/(xyz_max - xyz_min) (networks.py:104);
 * dirs (n,3) un-normalised (normalised inside, networks.py:144).  params_xyz_f16 = [3072 MLP | table], params_rgb_f16
 * = 7168, both already cast.  rgb_act: 1 = Sigmoid, 0 = None.  sigmas (n) f32 = exp(h0); rgbs (n,3) f32. */
int arn_field_fw(const float* xyzs, const float* dirs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                 arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act,
                 arn_field_ws_t ws, float* sigmas, float* rgbs, arn_stream_t stream);

/* Backward.  dL_dsigmas (n) / dL_drgbs (n,3) f32 (either may be NULL = zero).  grad_params_xyz (3072 + 2*entries) f32
 * and grad_params_rgb (7168) f32 are ACCUMULATED INTO (caller zeroes them).  dL_dxyzs (n,3) f32 optional (NULL to
 * skip): gradient w.r.t. the world coordinates (needed by render_surface_normal, rendering.py:301-313).
 * dfeat_scratch (n,32) f32 is scratch. */
int arn_field_bw(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                 arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act,
                 arn_field_ws_t ws, const float* sigmas, const float* rgbs, const float* dL_dsigmas,
                 const float* dL_drgbs, float loss_scale, float* dfeat_scratch, float* grad_params_xyz,
                 float* grad_params_rgb, float* dL_dxyzs, arn_stream_t stream);

/* Tensor-core implementation behind arn_field_fw / arn_field_bw (tcgen05 + TMEM, weights by bulk TMA copy). */
int arn_field_fw_tc(const float* xyzs, const float* dirs, int64_t n, const float* xyz_min_host,
                    const float* xyz_max_host, arn_levels_t levels, const void* params_xyz_f16,
                    const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws, float* sigmas, float* rgbs,
                    arn_stream_t stream);

int arn_field_bw_tc(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                    arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act,
                    arn_field_ws_t ws, const float* sigmas, const float* rgbs, const float* dL_dsigmas,
                    const float* dL_drgbs, float loss_scale, float* dfeat_scratch, float* grad_params_xyz,
                    float* grad_params_rgb, float* dL_dxyzs, arn_stream_t stream);

/* CUDA-core ("simt") implementation of the same two calls: operation order identical to oracle/oracle_field.c, kept
 * as the on-device cross-check of the tensor-core path (tests only; same arguments as arn_field_fw / arn_field_bw). */
int arn_field_fw_simt(const float* xyzs, const float* dirs, int64_t n, const float* xyz_min_host,
                      const float* xyz_max_host, arn_levels_t levels, const void* params_xyz_f16,
                      const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws, float* sigmas, float* rgbs,
                      arn_stream_t stream);
int arn_field_bw_simt(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                      arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act,
                      arn_field_ws_t ws, const float* sigmas, const float* rgbs, const float* dL_dsigmas,
                      const float* dL_drgbs, float loss_scale, float* dfeat_scratch, float* grad_params_xyz,
                      float* grad_params_rgb, float* dL_dxyzs, arn_stream_t stream);

/* Stand-alone pieces of the field (exported for parity tests and for the hash-encode GB/s measurement). */
int arn_hash_encode_fw(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                       arn_levels_t levels, const void* table_f16, void* feat_f16, arn_stream_t stream);
int arn_hash_encode_bw(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                       arn_levels_t levels, const void* table_f16, const float* dfeat, float* table_grad,
                       float* dL_dxyzs, arn_stream_t stream);
int arn_sh4(const float* dirs, int64_t n, void* out_f16, arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused training step (forward + loss + backward of one batch of rays in ONE host call, no host synchronisation).
 * Replaces, for the trainer, the Python sequence of train.py:174-198 / rendering.py:255-298 / losses.py:63-82:
 *   render(train) -> NeRFLoss('raw' rgb + opacity entropy [+ depth]) -> backward.
 * The marched sample count never leaves the device: per-sample buffers are sized by `capacity` (n_rays*max_samples is
 * always enough) and the kernels read the count from counter[0].  Gradients are ACCUMULATED into grad_xyz / grad_rgb;
 * the caller all-reduces them (N > 1) and calls arn_adam_step.  Results match the eager path (tests/test_gpu_parity.py).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    /* batch (device) */
    const float* rays_o; const float* rays_d; const float* rgb_target; const float* noise; int64_t n_rays;
    /* scene / model */
    const uint8_t* density_bitfield; int cascades; int grid_size; float scale; float exp_step_factor; int max_samples;
    float T_threshold; float near;
    const float* center_host; const float* half_size_host; const float* xyz_min_host; const float* xyz_max_host;
    arn_levels_t levels; const void* params_xyz_f16; const void* params_rgb_f16; int rgb_act;
    /* loss (losses.py:41-82, 'raw') and scaling */
    const float* bg_host; float lambda_opacity; float lambda_depth; float grad_scale; float loss_scale;
    /* workspace (device), per ray */
    float* hits_t; int64_t* rays_a; int32_t* counter; float* t_scratch; int32_t* count_scratch; int64_t* total_samples;
    float* opacity; float* depth; float* rgb; float* rgb_final; float* dL_dopacity; float* dL_ddepth; float* dL_drgb;
    /* workspace (device), per sample, `capacity` entries */
    int64_t capacity; float* xyzs; float* dirs; float* deltas; float* ts; float* sigmas; float* rgbs; float* ws_out;
    float* dL_dsigmas; float* dL_drgbs; float* dfeat; arn_field_ws_t ws;
    /* outputs */
    float* grad_xyz; float* grad_rgb; float* loss_out;
} arn_train_t;
int arn_train_fwbw(const arn_train_t* cfg_host, arn_stream_t stream);
/* (loss_out is zeroed by arn_train_march and accumulated by arn_train_fwbw_marched: give every march set that may be in
 * flight its own accumulator.) */
/* The same step in its two halves.  arn_train_march (ray/box, march count + scan + emit) reads only the rays, the noise and
 * the occupancy bitfield -- not the weights -- and fills rays_a, counter, xyzs, dirs, deltas, ts (+ the march scratch);
 * arn_train_fwbw_marched does the rest on those buffers.  A trainer can therefore march batch k+1 on a second stream
 * while batch k is in its field / optimizer kernels (ar_nerf_b200/trainer.py: next_rays=). */
int arn_train_march(const arn_train_t* cfg_host, arn_stream_t stream);
int arn_train_fwbw_marched(const arn_train_t* cfg_host, arn_stream_t stream);
/* Fork point for that second stream: the next arn_train_fwbw_marched calls of this thread record `cuda_event` (a cudaEvent_t)
 * on their stream right after launching stage `stage` (0 = before the field forward, 1 = after the field forward, 2 =
 * after compositing fw + bw, 3 = after the MLP backward i.e. in front of the hash-grid backward, 4 = after the whole call),
 * so that the caller can make the side stream wait there: the ray march is issue-bound and shares an SM best with the
 * memory-bound tail of the step (hash-grid reductions, Adam).  cuda_event == NULL switches the recording off. */
int arn_train_set_fork(int stage, void* cuda_event);
/* The opposite direction: the next arn_train_fwbw_marched calls of this thread make their stream WAIT for `cuda_event`
 * in front of stage `stage` (same numbering; 2 = in front of the MLP backward, the first kernel that writes parameter
 * gradients) -- e.g. for a gradient buffer that a side stream zeroes while the forward runs.  NULL switches it off. */
int arn_train_set_join(int stage, void* cuda_event);
/* Level-major hash-grid backward: the next field backward calls of this thread (arn_train_fwbw_marched, arn_field_bw_tc[_dyn])
 * walk the 16 levels in n_groups launches over the level ranges [level_begin[g], level_begin[g+1]) (level_begin[0] = 0,
 * level_begin[n_groups] = 16) and record cuda_events[g] (cudaEvent_t, may be NULL) on their stream behind group g.  Behind
 * event g the table gradient of those levels is final -- and behind event 0 the MLP weight gradients as well -- so an
 * optimizer / multi-GPU exchange for a finished group can run on another stream while the next group is still reducing
 * (SURVEY 5 "pipeline per hash level"; ar_nerf_b200/trainer.py).  n_groups = 0 restores the single launch. */
int arn_train_set_level_groups(int n_groups, const int* level_begin_host, void* const* cuda_events_host);
/* Compositing forward with the NeRFLoss epilogue of the fused step (one launch instead of two; rays_a in canonical ray
 * order).  Same outputs as arn_composite_train_fw followed by arn_nerf_loss. */
int arn_composite_train_fw_loss(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                const int64_t* rays_a, int64_t n_rays, int64_t n_samples, float T_threshold,
                                int64_t* total_samples, float* opacity, float* depth, float* rgb, float* ws,
                                const float* target, const float* bg_host, float lambda_opacity, float lambda_depth,
                                float grid_scale, float grad_scale, float* rgb_out, float* dL_drgb, float* dL_dopacity,
                                float* dL_ddepth, float* loss_out, arn_stream_t stream);
/* The loss kernel on its own (per-ray rgb/opacity/depth -> scalar loss + gradients; bg_host = 3 floats). */
int arn_nerf_loss(const float* rgb, const float* opacity, const float* depth, const float* target, int64_t n_rays,
                  const float* bg_host, float lambda_opacity, float lambda_depth, float grid_scale, float grad_scale,
                  float* rgb_out, float* dL_drgb, float* dL_dopacity, float* dL_ddepth, float* loss_out,
                  arn_stream_t stream);
/* Device-count forms (n = capacity, real count read from *n_dev) of the per-sample entry points. */
int arn_march_train_emit_dyn(const float* rays_o, const float* rays_d, int64_t n_rays, int cascades, int grid_size,
                             float scale, float exp_step_factor, int max_samples, const int64_t* rays_a,
                             const float* t_scratch, const int32_t* counter, float* xyzs, float* dirs, float* deltas,
                             float* ts, int64_t capacity, arn_stream_t stream);
int arn_hash_encode_fw_dyn(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host,
                           const float* xyz_max_host, arn_levels_t levels, const void* table_f16, void* feat_f16,
                           arn_stream_t stream);
int arn_hash_encode_bw_dyn(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host,
                           const float* xyz_max_host, arn_levels_t levels, const void* table_f16, const float* dfeat,
                           float* table_grad, float* dL_dxyzs, arn_stream_t stream);
int arn_field_fw_tc_dyn(const float* xyzs, const float* dirs, int64_t n, const int32_t* n_dev, const float* xyz_min_host,
                        const float* xyz_max_host, arn_levels_t levels, const void* params_xyz_f16,
                        const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws, float* sigmas, float* rgbs,
                        arn_stream_t stream);
int arn_field_bw_tc_dyn(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host,
                        const float* xyz_max_host, arn_levels_t levels, const void* params_xyz_f16,
                        const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws, const float* sigmas,
                        const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                        float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs,
                        arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * One iteration of the test-time render loop (rendering.py:189-236: raymarching_test -> model -> composite_test_fw ->
 * alive_indices[alive_indices >= 0]) in ONE host call with no host round trip inside.  Same schedule, same per-ray
 * arithmetic as the reference's loop: the march is the bit-exact test march, samples are evaluated as a compact list
 * (the reference's valid_mask indexing), compositing is composite_test_fw's loop, the alive list keeps its order.
 * After the call the caller reads counts[0] (valid samples of this iteration; 0 -> the reference breaks) and
 * counts_alive[0] (rays still alive) to drive N_samples of the next iteration.  hits_t (R,2), opacity/depth/rgb in place.
 * capacity >= n_alive * n_samples bounds every per-sample buffer; ws.hid etc. may be NULL (inference). */
typedef struct {
    const float* rays_o; const float* rays_d; float* hits_t; const int64_t* alive; int64_t n_alive;
    const uint8_t* density_bitfield; int cascades; int grid_size; float scale; float exp_step_factor; int n_samples; int max_samples;
    float T_threshold;
    const float* xyz_min_host; const float* xyz_max_host; arn_levels_t levels; const void* params_xyz_f16; const void* params_rgb_f16; int rgb_act;
    /* workspace */
    int64_t capacity; float* deltas; float* ts; int32_t* n_eff; int64_t* rays_a; int32_t* counts; int32_t* counts_alive;
    float* xyzs; float* dirs; float* sigmas; float* rgbs; arn_field_ws_t ws;
    /* in/out */
    float* opacity; float* depth; float* rgb; int64_t* alive_out; int64_t* total_samples;
    /* device-driven loop only: numerator of the schedule N_samples = max(min(schedule_rays // N_alive, 64), min_samples).
     * 0 (or anything <= N_rays) = N_rays, the reference's schedule; k * N_rays asks for k times the samples per iteration --
     * fewer, larger iterations, the same per-ray results (a ray's samples and their compositing order do not depend on the
     * slicing), more samples evaluated behind the point where a ray terminates.  capacity >= schedule_rays. */
    int64_t schedule_rays;
} arn_test_iter_t;
int arn_render_test_iter(const arn_test_iter_t* cfg_host, arn_stream_t stream);
/* Far clamp of a frame's rays, once, in front of the test loop: every ray is marched to its end with the test march's
 * own arithmetic and hits_t[r][1] is pulled back to the chain point behind the ray's last occupied sample (to
 * hits_t[r][0] if it meets no occupied cell).  The loop then produces exactly the samples, N_eff values and kill pattern
 * of the unclamped march, without any iteration walking a ray's empty exit stretch.  hits_t (R,2) in place. */
int arn_march_test_far_clamp(const float* rays_o, const float* rays_d, float* hits_t, int64_t n_rays,
                             const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                             float exp_step_factor, int max_samples, arn_stream_t stream);
/* The same iteration with the loop's control state on the device, so that a caller can queue iterations without reading
 * anything back (the reference's loop costs three host synchronisations per iteration, rendering.py:186,198,219).
 * state (8 x int32, device) = {n_alive, N_samples, samples requested so far, active, iterations done, iterations that started
 * active, 0, 0}: the iteration reads
 * state_in and writes state_out -- the schedule of rendering.py:184-206 (stop when nothing was marched, nobody is alive or
 * budget_samples (= kwargs max_samples) is spent; N_samples = max(min(N_rays // N_alive, 64), min_samples)); pass two
 * alternating buffers.  Initial state for a frame of N_rays: {N_rays, S0, S0, 1, 0, 0, 0, 0} with S0 = max(1, min_samples) (or
 * active = 0 when budget_samples <= 0).  An inactive state turns the call into no-ops, so iterations may be queued
 * speculatively and the state read back late.  cfg->n_alive = N_rays (whole frame), cfg->alive/alive_out alternate as
 * before, cfg->n_samples is ignored; n_upper >= the device-side n_alive (grid sizing only; N_rays is always valid);
 * partial: (N_rays + 127) / 128 int32 of scratch; capacity >= N_rays * min_samples.  Every launch is a grid-stride loop over
 * device-side counts with no host synchronisation, so a run of calls can be captured into a CUDA graph with n_upper = N_rays
 * (ar_nerf_b200/rendering.py replays 8 iterations per graph launch). */
int arn_render_test_step(const arn_test_iter_t* cfg_host, const int32_t* state_in, int32_t* state_out, int32_t* partial,
                         int min_samples, int budget_samples, int64_t n_upper, arn_stream_t stream);
/* The frame's samples marched ONCE.  The test march is resumable and deterministic (an iteration starts at the chain point
 * behind the previous iteration's last sample), so a ray's sample sequence does not depend on how the loop slices it:
 * arn_march_test_all marches every ray of the frame to its end with the test march's own arithmetic and records the
 * parameter t of every occupied sample, sample-major (ts_all[s * n_rays + r]), at most `stride` per ray (the loop never
 * asks a ray for more than max_samples + 63: stride = budget + 64 is always enough), the count in totals[r], and zeroes
 * cursor[r].  arn_render_test_step_pre is arn_render_test_step with the march replaced by a slice of ts_all at the ray's
 * cursor (dt = calc_dt(t) as the march computes it): same samples, same N_eff, same kill pattern, no marching inside the
 * loop.  hits_t is not advanced in this form. */
int arn_march_test_all(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                       const uint8_t* density_bitfield, int cascades, int grid_size, float scale, float exp_step_factor,
                       int max_samples, int stride, float* ts_all, int32_t* totals, int32_t* cursor, arn_stream_t stream);
int arn_render_test_step_pre(const arn_test_iter_t* cfg_host, const int32_t* state_in, int32_t* state_out, int32_t* partial,
                             const float* ts_all, const int32_t* totals, int32_t* cursor, int min_samples,
                             int budget_samples, int64_t n_upper, arn_stream_t stream);
/* The same iteration in FOUR launches instead of seven (slice + emit | hash grid | MLP | compositing + survivors + control
 * state): the two scans are replaced by atomics whose order does not matter -- the compact sample list and the next alive list
 * come out in a scheduling-dependent ORDER, every per-ray result, the kill pattern and total_samples are those of
 * arn_render_test_step_pre.  A small frame (one rank's share of a frame rendered by several GPUs) is bound by the number of
 * kernels of its ~50 iterations.  sync: 4 x int32 of device scratch, zero before a frame's first iteration (every iteration
 * leaves it zero); cfg->counts / the `partial` scratch are not used. */
int arn_render_test_step_fused(const arn_test_iter_t* cfg_host, const int32_t* state_in, int32_t* state_out, int32_t* sync,
                               const float* ts_all, const int32_t* totals, int32_t* cursor, int min_samples,
                               int budget_samples, int64_t n_upper, arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimizer step.  Replaces apex FusedAdam(lr, betas=(0.9,0.999), eps=1e-15, weight_decay=0) (train.py:146).
 * One fused pass: Adam update of the fp32 master, optional un-scaling of the gradient by inv_grad_scale, refresh of
 * the fp16 working copy (dst_f16 may be NULL), and zeroing of the gradient for the next step (zero_grad != 0).
 * step is the 1-based step count (bias correction as torch.optim.Adam).
 * ---------------------------------------------------------------------------------------------------------- */
int arn_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* dst_f16, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step, float inv_grad_scale, int zero_grad,
                  arn_stream_t stream);
/* The same update for a large tensor (>= 4096 elements, multiple of 4, 16-byte aligned) and a small one (<= 2^20 elements,
 * updated element by element by extra blocks) in ONE launch: same hyper-parameters, step and zero_grad for both. */
int arn_adam_step2(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* dst_f16, int64_t n,
                   float* params2, float* grads2, float* exp_avg2, float* exp_avg_sq2, void* dst2_f16, int64_t n2,
                   float lr, float beta1, float beta2, float eps, int step, float inv_grad_scale, int zero_grad,
                   arn_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU: gradient exchange fused with the optimizer over NVLink peer memory (one process per GPU; replaces the
 * DDP all-reduce of train.py's Lightning trainer + apex FusedAdam for the hash table).
 *   arn_p2p_alloc / free    device buffer whose base pointer can be exported (cudaMalloc, zero-filled)
 *   arn_p2p_export / open / close   CUDA IPC handle (64 bytes) of such a buffer / mapping of a peer's buffer
 *   arn_p2p_signal          publish `value` in slot `slot` of every rank's flag array (after a system-scope fence)
 *   arn_p2p_wait            wait until every rank has published >= value in `slot` of MY flag array (bounded spin)
 *   arn_p2p_adam_exchange   for the slice [lo, lo+count) of the flat parameter: sum the ranks' gradients read from peer
 *                           memory (rank order), Adam (arn_adam_step's arithmetic), write the fp16 result into every rank's
 *                           fp16 working copy.  params/exp_avg/exp_avg_sq point at this rank's slice (local memory).
 * Flag arrays hold ARN_P2P_FLAG_SLOTS * ARN_P2P_MAX_RANKS uint64 and must come from arn_p2p_alloc (zeroed). */
#define ARN_P2P_MAX_RANKS 8
#define ARN_P2P_FLAG_SLOTS 24
int arn_p2p_alloc(void** ptr_host, int64_t bytes);
int arn_p2p_free(void* ptr);
int arn_p2p_export(void* ptr, unsigned char* handle64_host);
int arn_p2p_open(const unsigned char* handle64_host, void** ptr_host);
int arn_p2p_close(void* ptr);
int arn_p2p_signal(void* const* peer_flags_host, int n_ranks, int rank, int slot, uint64_t value, arn_stream_t stream);
int arn_p2p_wait(const void* my_flags, int n_ranks, int slot, uint64_t value, arn_stream_t stream);
/* arn_p2p_signal followed by arn_p2p_wait on the same slot and value, as one launch. */
int arn_p2p_barrier(void* const* peer_flags_host, const void* my_flags, int n_ranks, int rank, int slot, uint64_t value,
                    arn_stream_t stream);
/* Waits are bounded spins.  A wait that exceeds the budget (arn_p2p_set_timeout, seconds; default 120 -- ranks may drift apart
 * by a checkpoint write or a validation pass; call a host-side barrier first if more is expected) does NOT trap: it stores
 * a non-zero code (0x100 | rank it waited for: arn_p2p_wait, 0x200 | rank: arn_p2p_barrier) into the error word registered
 * with arn_p2p_set_error_word -- the DEVICE address of a zeroed uint32 in pinned, mapped host memory -- and returns; exchange
 * kernels launched behind it see the word and update nothing.  The host reads the word whenever it likes (no
 * synchronisation needed) and decides: raise, or fall back to the NCCL form of the same sharding.
 * arn_p2p_set_grid: thread blocks per SM of the exchange kernel (default 8); a caller that overlaps the exchange of one level
 * group with the hash-grid backward of the next (arn_train_set_level_groups) asks for 1-2 so that both hold SM slots. */
/* arn_p2p_adam_exchange through the NVSwitch's multicast (NVLS): mc_grads / mc_p16 are MULTICAST addresses of the ranks'
 * gradient / fp16 buffers (same offsets on every rank; torch symmetric memory hands them out).  The gradient sum of a float4
 * group is ONE multimem.ld_reduce (the switch adds the ranks' copies: the reduction order is the fabric's), the fp16 result
 * ONE multimem.st into every rank's copy -- 43 / 26 MB per link direction and step at 8 ranks instead of 60 / 60. */
int arn_p2p_adam_exchange_mc(const void* mc_grads, void* mc_p16, int64_t lo, int64_t count, float* params_slice,
                             float* exp_avg_slice, float* exp_avg_sq_slice, float lr, float beta1, float beta2, float eps,
                             int step, float inv_grad_scale, arn_stream_t stream);
int arn_p2p_set_timeout(double seconds);
int arn_p2p_set_error_word(void* err_word);
int arn_p2p_set_grid(int blocks_per_sm);
int arn_p2p_adam_exchange(void* const* peer_grads_host, void* const* peer_p16_host, int n_ranks, int64_t lo, int64_t count,
                          float* params_slice, float* exp_avg_slice, float* exp_avg_sq_slice, float lr, float beta1,
                          float beta2, float eps, int step, float inv_grad_scale, arn_stream_t stream);


/* ------------------------------------------------------------------------------------------------------------
 * Spherical-Gaussian shadow and shading of an inserted object (AR insertion frame, BASELINE.json configs[4]).
 * Replaces insert/sg_shadow.py:103-116 SGShadow.calc_shadow_factor, :118-153 SGShadow.calc_self_shadow_light_dacay and
 * insert/render_utils.py:321-375 SG_render_core (with :266-318 SGProduct / SGHemisphereIntegral / SGIrradiance).
 * Tables (device, fp32): coeff_cl = the PCA coefficient volume CHANNEL-LAST (D,H,W,C) -- the reference's
 * coeff_volume (1,C,D,H,W) permuted once by the caller; components (C,envH,envW), mean (envH,envW) = the PCA basis over the
 * light direction; fh_tab (fh_h = sharpness rows, fh_w = angle columns) = insert/data/fh_pretab.npy.
 * lSGs (L,7) = axis(3) | sharpness | rgb; pts (n,3) world positions; model_pos_host (3), rot_inv_host (9, row-major, may
 * be NULL), scale = model radius.  light_scratch: ARN_SG_SCRATCH_FLOATS(L, C) floats of device scratch (per-light tables
 * built by a one-block prologue kernel).  All arithmetic fp32 in the reference's operation order; grid_sample's
 * bilinear / padding_mode='border' rules (align_corners as the reference calls it: True for the volume only).
 * ---------------------------------------------------------------------------------------------------------- */
#define ARN_SG_MAX_LIGHTS 64
#define ARN_SG_MAX_COMPONENTS 128
#define ARN_SG_SCRATCH_FLOATS(n_lights, n_components) ((n_lights) * ((n_components) + 12) + 3)
typedef struct {
    const float* coeff_cl; int D; int H; int W; int C;
    const float* components; const float* mean; int envH; int envW;
    const float* fh_tab; int fh_h; int fh_w;
    float vol_range; float angle_decay_fac; float shadow_pow_fac; float self_shadow_pow_fac;
} arn_sg_tables_t;
/* factor (n): the shadow the object casts on scene points (lSGs already rotated into the model frame by the caller when
 * rot_inv_host is given, as main.py:496-499 does). */
int arn_sg_shadow_factor(const arn_sg_tables_t* tables_host, const float* lSGs, int n_lights, const float* pts, int64_t n,
                         const float* model_pos_host, const float* rot_inv_host, float scale, float* light_scratch,
                         float* factor, arn_stream_t stream);
/* Self-shadow attenuation of the lights per pixel fused with SG_render_core: radiance (n,3).  lSGs_axis (L,7; may be NULL =
 * lSGs) supplies the axes used for the environment lookup (the lights rotated into the model frame), lSGs the axes /
 * sharpness / colours that are shaded.  self_shadow = 0: SG_render_core(..., self_shadow=False) on lSGs as they are.
 * lSGs_out (n,L,7), optional: the attenuated lights (what calc_self_shadow_light_dacay returns); radiance == NULL computes
 * only those.  albedo (n,3), metal (n), rough (n), normal (n,3) (any length), vdirs (n,3) unit view directions. */
int arn_sg_shade(const arn_sg_tables_t* tables_host, const float* lSGs, const float* lSGs_axis, int n_lights, const float* pts,
                 int64_t n, const float* model_pos_host, const float* rot_inv_host, float scale, const float* albedo,
                 const float* metal, const float* rough, const float* normal, const float* vdirs, int clamp01, int self_shadow,
                 float* light_scratch, float* lSGs_out, float* radiance, arn_stream_t stream);

/* SG_render_core on lights given by the caller: per_pixel = 1: lSGs (n,L,7), one set per pixel (self_shadow=True: what
 * calc_self_shadow_light_dacay returned); per_pixel = 0: lSGs (L,7) shared by all pixels (self_shadow=False). */
int arn_sg_shade_px(const float* lSGs, int n_lights, int per_pixel, int64_t n, const float* albedo, const float* metal,
                    const float* rough, const float* normal, const float* vdirs, int clamp01, float* radiance,
                    arn_stream_t stream);

/* Shadow field, the SH alternative to the SG shadow: insert/shadow_fields.py:59-78 soft_shadow_map with :92-101 / :112-121
 * fetch_sh (SimplifySF / ComplexSF) and insert/insert_utils.py:153-154 SH_product0.  sf_cl = the field's volume CHANNEL-LAST
 * (D,H,W,K) -- the reference's sf_vol (1,K,D,H,W) permuted once by the caller --, K <= ARN_SF_MAX_COEFFS SH coefficients;
 * model_sh_host = the lighting's SH (K,3) row-major on the host.  Out (either may be NULL): sh_out (n,K) = fetch_sh of the
 * points relative to the model, shadow (n) = the soft shadow factor. */
#define ARN_SF_MAX_COEFFS 16
int arn_sf_soft_shadow(const float* sf_cl, int D, int H, int W, int K, float vol_range, const float* model_sh_host, const float* pts,
                       int64_t n, const float* model_pos_host, const float* rot_inv_host, float scale, float* sh_out, float* shadow,
                       arn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ARNERF_H_ */
