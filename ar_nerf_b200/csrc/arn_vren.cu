// libarnerf.so -- kernels replacing the reference's `vren` extension (models/csrc/*.cu), hand-written for sm_100a.
// Compiled with -fmad=false: every fused multiply-add on a bit-exact path is explicit (arn_march_core.h).
//
//   intersection   one thread per ray, fused near clamp for the single-box case used by render()
//   march (train)  pass 1: one thread per ray counts samples (bit-exact with raymarching.cu:184-234) and records the
//                  sample parameters t into a caller-provided scratch; a single-CTA scan lays rays out canonically;
//                  pass 2 is then embarrassingly parallel: one thread per SAMPLE, coalesced stores, no re-march.
//   march (test)   one thread per alive ray, zero-fills its own padding (no memset launches)
//   compositing    one warp per ray: multiplicative/additive warp scans, ballot for early termination
#include "arn_common.cuh"
#include "arn_march_core.h"

namespace arn {

// ---------------------------------------------------------------------------------------------- intersection
// intersection.cu:5-22
__device__ __forceinline__ float2 slab_test(const float* o, const float* inv, const float* c, const float* h) {
    float t1 = -INFINITY, t2 = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float tmin = __fmul_rn(__fsub_rn(__fsub_rn(c[k], h[k]), o[k]), inv[k]);
        const float tmax = __fmul_rn(__fsub_rn(__fadd_rn(c[k], h[k]), o[k]), inv[k]);
        const float a = fminf(tmin, tmax), b = fmaxf(tmin, tmax);
        t1 = (k == 0) ? a : fmaxf(t1, a);
        t2 = (k == 0) ? b : fminf(t2, b);
    }
    if (t1 > t2) return make_float2(-1.0f, -1.0f);
    return make_float2(t1, t2);
}

// intersection.cu:103-121 ; dot(a,b) is contracted by nvcc as fma(a.z,b.z, fma(a.x,b.x, a.y*b.y)) (reference SASS)
__device__ __forceinline__ float2 sphere_test(const float* o, const float* d, const float* c, float radius) {
    const float cx = __fsub_rn(o[0], c[0]), cy = __fsub_rn(o[1], c[1]), cz = __fsub_rn(o[2], c[2]);
    const float a = __fmaf_rn(d[2], d[2], __fmaf_rn(d[0], d[0], __fmul_rn(d[1], d[1])));
    const float half_b = __fmaf_rn(d[2], cz, __fmaf_rn(d[0], cx, __fmul_rn(d[1], cy)));
    const float cc = __fmaf_rn(-radius, radius, __fmaf_rn(cz, cz, __fmaf_rn(cx, cx, __fmul_rn(cy, cy))));
    const float disc = __fmaf_rn(half_b, half_b, -__fmul_rn(a, cc));
    if (disc < 0) return make_float2(-1.0f, -1.0f);
    const float sq = __fsqrt_rn(disc);
    return make_float2(__fdiv_rn(__fsub_rn(-half_b, sq), a), __fdiv_rn(__fadd_rn(-half_b, sq), a));
}

template <bool SPHERE>
__global__ void __launch_bounds__(256) intersect_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        int64_t n_rays, const float* __restrict__ centers,
                                                        const float* __restrict__ extents, int n_prims, int max_hits,
                                                        int32_t* __restrict__ hit_cnt, float* __restrict__ hits_t,
                                                        int64_t* __restrict__ hits_idx) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
    const float d[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
    const float inv[3] = {__fdiv_rn(1.0f, d[0]), __fdiv_rn(1.0f, d[1]), __fdiv_rn(1.0f, d[2])};
    float* ht = hits_t + r * max_hits * 2;
    int64_t* hi = hits_idx + r * max_hits;
    for (int k = 0; k < max_hits; k++) { ht[2 * k] = -1.0f; ht[2 * k + 1] = -1.0f; hi[k] = -1; }
    int cnt = 0;
    for (int v = 0; v < n_prims; v++) {
        const float c[3] = {centers[3 * v], centers[3 * v + 1], centers[3 * v + 2]};
        float2 tt;
        if (SPHERE) tt = sphere_test(o, d, c, extents[v]);
        else { const float h[3] = {extents[3 * v], extents[3 * v + 1], extents[3 * v + 2]}; tt = slab_test(o, inv, c, h); }
        if (tt.y > 0) {  // intersection.cu:49-55
            if (cnt < max_hits) { ht[2 * cnt] = fmaxf(tt.x, 0.0f); ht[2 * cnt + 1] = tt.y; hi[cnt] = v; }
            cnt++;
        }
    }
    hit_cnt[r] = cnt;
    // near -> far on t1, ascending with the -1 fills (what torch::sort does at intersection.cu:95-97)
    for (int i = 1; i < max_hits; i++) {
        const float a0 = ht[2 * i], a1 = ht[2 * i + 1]; const int64_t ai = hi[i];
        int j = i - 1;
        while (j >= 0 && ht[2 * j] > a0) { ht[2 * j + 2] = ht[2 * j]; ht[2 * j + 3] = ht[2 * j + 1]; hi[j + 1] = hi[j]; j--; }
        ht[2 * j + 2] = a0; ht[2 * j + 3] = a1; hi[j + 1] = ai;
    }
}

// rendering.py:29-31 fused: single box, max_hits 1, near clamp.
__global__ void __launch_bounds__(256) aabb_near_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        int64_t n_rays, float cx, float cy, float cz, float hx, float hy,
                                                        float hz, float near, float2* __restrict__ hits_t) {
    pdl_enter();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
    const float inv[3] = {__fdiv_rn(1.0f, rays_d[3 * r]), __fdiv_rn(1.0f, rays_d[3 * r + 1]), __fdiv_rn(1.0f, rays_d[3 * r + 2])};
    const float c[3] = {cx, cy, cz}, h[3] = {hx, hy, hz};
    const float2 tt = slab_test(o, inv, c, h);
    float2 out = make_float2(-1.0f, -1.0f);
    if (tt.y > 0) {
        out.x = fmaxf(tt.x, 0.0f); out.y = tt.y;
        if (out.x >= 0 && out.x < near) out.x = near;
    }
    hits_t[r] = out;
}

// ---------------------------------------------------------------------------------------------- ray generation
// train.py:121-126 + datasets/ray_utils.py:46-70 for one sampled batch: directions[pix_idxs] rotated by poses[img_idxs]
// (rays_d = R d, k ascending, fp32 like the reference's batched matmul under autocast(float32)), rays_o = the camera
// centre.  K != NULL recomputes the direction from the pixel index (ray_utils.py:33-35) instead of reading a table.
__global__ void __launch_bounds__(256) gather_rays_kernel(const float* __restrict__ directions, const float* __restrict__ poses,
                                                          const int64_t* __restrict__ img_idxs, int64_t img_single,
                                                          const int64_t* __restrict__ pix_idxs, int64_t n, int width, float fx, float fy, float cx,
                                                          float cy, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                                          const float* __restrict__ images, int64_t pixels_per_image, int channels,
                                                          float* __restrict__ pixels_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t pix = pix_idxs[i];
    if (images) {  // rays[img_idxs, pix_idxs] of datasets/base.py:32: the pixel's colour (+ exposure) row
        const float* src = images + ((img_idxs ? img_idxs[i] : img_single) * pixels_per_image + pix) * channels;
        for (int c = 0; c < channels; c++) pixels_out[i * channels + c] = src[c];
    }
    float d[3];
    if (directions) { d[0] = directions[3 * pix]; d[1] = directions[3 * pix + 1]; d[2] = directions[3 * pix + 2]; }
    else {
        const float u = (float)(pix % width), v = (float)(pix / width);
        d[0] = __fdiv_rn(__fadd_rn(__fsub_rn(u, cx), 0.5f), fx); d[1] = __fdiv_rn(__fadd_rn(__fsub_rn(v, cy), 0.5f), fy); d[2] = 1.0f;
    }
    const float* P = poses + 12 * (img_idxs ? img_idxs[i] : img_single);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        rays_d[3 * i + r] = __fmaf_rn(d[2], P[4 * r + 2], __fmaf_rn(d[1], P[4 * r + 1], __fmul_rn(d[0], P[4 * r])));
        rays_o[3 * i + r] = P[4 * r + 3];
    }
}

// ---------------------------------------------------------------------------------------------- grid utilities
__global__ void __launch_bounds__(256) morton3d_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t* __restrict__ indices) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    indices[i] = (int32_t)arn_morton3d((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1], (uint32_t)coords[3 * i + 2]);
}
__global__ void __launch_bounds__(256) morton3d_invert_kernel(const int32_t* __restrict__ indices, int64_t n, int32_t* __restrict__ coords) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t ind = indices[i];
    coords[3 * i + 0] = (int32_t)arn_morton3d_invert((uint32_t)(ind >> 0));
    coords[3 * i + 1] = (int32_t)arn_morton3d_invert((uint32_t)(ind >> 1));
    coords[3 * i + 2] = (int32_t)arn_morton3d_invert((uint32_t)(ind >> 2));
}

// raymarching.cu:122-141.  One thread packs 4 output bytes from 32 consecutive cells (128-bit loads for f32),
// so a warp reads 4 KB contiguous and writes 128 B contiguous.
template <typename T>
__global__ void __launch_bounds__(256) packbits_kernel(const T* __restrict__ grid, float thr, uint8_t* __restrict__ bits, int64_t n_bytes) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // word index (4 bytes)
    const int64_t b0 = w * 4;
    if (b0 >= n_bytes) return;
    uint32_t word = 0;
    const int nb = (int)min((int64_t)4, n_bytes - b0);
    for (int b = 0; b < nb; b++) {
        uint32_t byte = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) byte |= ((float)grid[(b0 + b) * 8 + i] > thr) ? (1u << i) : 0u;
        word |= byte << (8 * b);
    }
    if (nb == 4 && ((uintptr_t)(bits + b0) & 3) == 0) *reinterpret_cast<uint32_t*>(bits + b0) = word;
    else for (int b = 0; b < nb; b++) bits[b0 + b] = (uint8_t)(word >> (8 * b));
}
// double: compared in double precision against (double)thr, as the reference's template does (scalar_t > float)
__global__ void __launch_bounds__(256) packbits_f64_kernel(const double* __restrict__ grid, float thr, uint8_t* __restrict__ bits, int64_t n_bytes) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_bytes) return;
    uint32_t byte = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) byte |= (grid[n * 8 + i] > (double)thr) ? (1u << i) : 0u;
    bits[n] = (uint8_t)byte;
}

// ---------------------------------------------------------------------------------------------- occupancy refresh
// networks.py:263-267: cell centre in world units plus a uniform jitter of half a cell, in torch's operation order
// (int -> float, division by the scalar G-1 -- which torch's CUDA kernel performs as a multiplication by the float
// reciprocal --, *2, -1, *(s - half); rand*2-1, *half; +): positions are bit-identical with
//   xyzs_w = (coords / (G-1) * 2 - 1) * (s - half);  xyzs_w += (rand * 2 - 1) * half
__global__ void __launch_bounds__(256) cell_positions_kernel(const int32_t* __restrict__ coords, const float* __restrict__ rnd, int64_t n3,
                                                             float inv_gm1, float s_minus_half, float half, float* __restrict__ xyzs) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one coordinate per thread
    if (e >= n3) return;
    const float c = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)coords[e], inv_gm1), 2.0f), 1.0f), s_minus_half);
    const float j = __fmul_rn(__fsub_rn(__fmul_rn(rnd[e], 2.0f), 1.0f), half);
    xyzs[e] = __fadd_rn(c, j);
}

// NGP.mark_invisible_cells (networks.py:209-250): a cell of cascade c is kept (density 0) when at least one camera sees its
// centre inside the image at depth >= near and no camera has it inside the image closer than near; otherwise -1.  One
// thread per cell walks all cameras (world-to-camera rows staged through shared memory); count_grid gets the covered
// fraction.  Operation order (plain IEEE mul/add, this file is compiled with -fmad=false):
//   x_w = ((coord * 1/(G-1)) * 2 - 1) * (s - s/G);  x_c[i] = ((R[i][0] x + R[i][1] y) + R[i][2] z) + T[i];
//   uvd[i] = (K[i][0] x_c + K[i][1] y_c) + K[i][2] z_c;  uv = uvd[:2] / uvd[2]
constexpr int kCamTile = 128;
__global__ void __launch_bounds__(256) mark_invisible_kernel(const int32_t* __restrict__ coords, const int64_t* __restrict__ indices, int64_t n_cells,
                                                             float inv_gm1, float s_minus_half, const float* __restrict__ w2c, int n_cams,
                                                             float k00, float k01, float k02, float k10, float k11, float k12, float k20, float k21,
                                                             float k22, float img_w, float img_h, float near, float* __restrict__ density_grid,
                                                             float* __restrict__ count_grid) {
    __shared__ float cam[kCamTile * 12];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n_cells;
    float x = 0.f, y = 0.f, z = 0.f;
    if (live) {
        x = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)coords[3 * i], inv_gm1), 2.0f), 1.0f), s_minus_half);
        y = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)coords[3 * i + 1], inv_gm1), 2.0f), 1.0f), s_minus_half);
        z = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)coords[3 * i + 2], inv_gm1), 2.0f), 1.0f), s_minus_half);
    }
    int covered = 0; bool too_near = false;
    for (int c0 = 0; c0 < n_cams; c0 += kCamTile) {
        const int nc = min(kCamTile, n_cams - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < nc * 12; e += blockDim.x) cam[e] = w2c[(int64_t)c0 * 12 + e];
        __syncthreads();
        if (!live) continue;
        for (int c = 0; c < nc; c++) {
            const float* M = cam + 12 * c;  // R row-major (9) | T (3)
            const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[0], x), __fmul_rn(M[1], y)), __fmul_rn(M[2], z)), M[9]);
            const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[3], x), __fmul_rn(M[4], y)), __fmul_rn(M[5], z)), M[10]);
            const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[6], x), __fmul_rn(M[7], y)), __fmul_rn(M[8], z)), M[11]);
            const float u0 = __fadd_rn(__fadd_rn(__fmul_rn(k00, xc), __fmul_rn(k01, yc)), __fmul_rn(k02, zc));
            const float v0 = __fadd_rn(__fadd_rn(__fmul_rn(k10, xc), __fmul_rn(k11, yc)), __fmul_rn(k12, zc));
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(k20, xc), __fmul_rn(k21, yc)), __fmul_rn(k22, zc));
            const float u = __fdiv_rn(u0, d), v = __fdiv_rn(v0, d);
            const bool in_image = d >= 0.0f && u >= 0.0f && u < img_w && v >= 0.0f && v < img_h;
            covered += (in_image && d >= near) ? 1 : 0;
            too_near |= in_image && d < near;
        }
    }
    if (!live) return;
    const float count = __fdiv_rn((float)covered, (float)n_cams);
    const int64_t o = indices[i];
    count_grid[o] = count;
    density_grid[o] = (count > 0.0f && !too_near) ? 0.0f : -1.0f;
}

// ---- Steady-state cell selection of the occupancy refresh (networks.py:181-207 + :263-267) in three small launches.
// M uniform cells come from the caller's randint draw; M occupied cells are "the k-th occupied cell, k = u mod count" of the
// caller's second draw (the reference indexes nonzero(grid > thr) with randint(count): the same distribution).  The k-th
// occupied cell is found through occupancy bit masks (one word per 32 cells) and a prefix over 1024-cell chunks -- a binary
// search in shared memory and one 128-byte read of the chunk's masks -- instead of cumsum + searchsorted over the whole
// grid; the jittered position of the cell (networks.py:263-267) is written by the same thread.
// arn_grid_sample_cells_sorted returns the same cells ordered along the morton curve (a counting sort over the cells: the
// density evaluation that follows gathers 1 M cells' hash-grid corners in 167 us instead of 295 when neighbours in the list are
// neighbours in space).  The sort costs about what it saves, so it only pays where the selection is computed AHEAD of the
// refresh, off the step's critical path: it depends on the previous refresh's density_grid and on random draws only
// (NGP.update_density_grid prefetches it on a side stream).
constexpr int kCellChunk = 1024;  // cells per chunk = 32 mask words
constexpr int kMaxChunks = 4096;
__global__ void __launch_bounds__(256) occ_mask_kernel(const float* __restrict__ grid, float thr, int64_t n_cells, uint32_t* __restrict__ masks,
                                                       int32_t* __restrict__ chunk_count) {
    const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per chunk
    const int lane = threadIdx.x & 31;
    if (chunk * kCellChunk >= n_cells) return;
    uint32_t mine = 0u; int cnt = 0;
    for (int w = 0; w < 32; w++) {
        const int64_t i = chunk * kCellChunk + w * 32 + lane;
        const unsigned m = __ballot_sync(kFull, i < n_cells && grid[i] > thr);
        if (lane == w) mine = m;
        cnt += __popc(m);
    }
    masks[chunk * 32 + lane] = mine;
    if (lane == 0) chunk_count[chunk] = cnt;
}
// exclusive prefix of up to 4096 ints in one CTA: out[i] = sum(in[0..i)), out[n] = total
__global__ void __launch_bounds__(1024) small_scan_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out) {
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int v[4]; int sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int i = threadIdx.x * 4 + k; v[k] = i < n ? in[i] : 0; sum += v[k]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFull, w, o); if (lane >= o) w += u; }
        warp_tot[lane] = w;
    }
    __syncthreads();
    int run = inc - sum + (wid ? warp_tot[wid - 1] : 0);
#pragma unroll
    for (int k = 0; k < 4; k++) { const int i = threadIdx.x * 4 + k; if (i < n) out[i] = run; run += v[k]; }
    if (threadIdx.x == 0) out[n] = warp_tot[31];
}
// jittered position of a cell (networks.py:263-267), bit-identical with the torch expression
__device__ __forceinline__ void cell_position(uint32_t idx, const float* __restrict__ rnd3, float inv_gm1, float s_minus_half, float half,
                                              float* __restrict__ out3) {
    const uint32_t c[3] = {arn_morton3d_invert(idx), arn_morton3d_invert(idx >> 1), arn_morton3d_invert(idx >> 2)};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float cc = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn((float)(int32_t)c[d], inv_gm1), 2.0f), 1.0f), s_minus_half);
        const float jj = __fmul_rn(__fsub_rn(__fmul_rn(rnd3[d], 2.0f), 1.0f), half);
        out3[d] = __fadd_rn(cc, jj);
    }
}
// cell index + jittered position of every draw (j < M: uniform; j >= M: k-th occupied).  SORTED: the index goes to idx_tmp
// and into the histogram over the cells instead (place_cells_kernel writes the outputs in curve order).
template <bool SORTED>
__global__ void __launch_bounds__(256) pick_cells_kernel(const int32_t* __restrict__ coords1, const int64_t* __restrict__ u, int64_t M,
                                                         const uint32_t* __restrict__ masks, const int32_t* __restrict__ chunk_prefix, int n_chunks,
                                                         const float* __restrict__ rnd, float inv_gm1, float s_minus_half, float half,
                                                         int64_t* __restrict__ indices, float* __restrict__ xyzs,
                                                         uint32_t* __restrict__ idx_tmp, int32_t* __restrict__ hist) {
    __shared__ int s_prefix[kMaxChunks + 1];
    const bool occupied_half = (int64_t)blockIdx.x * blockDim.x >= M;  // M is a multiple of the block size (checked by the host)
    if (occupied_half) {
        for (int b = threadIdx.x; b <= n_chunks; b += blockDim.x) s_prefix[b] = chunk_prefix[b];
        __syncthreads();
    }
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * M) return;
    uint32_t idx;
    if (!occupied_half) {
        idx = arn_morton3d((uint32_t)coords1[3 * j], (uint32_t)coords1[3 * j + 1], (uint32_t)coords1[3 * j + 2]);
    } else {
        const int total = s_prefix[n_chunks];
        const int k = (int)(u[j - M] % (int64_t)max(total, 1));
        int lo = 0, hi = n_chunks - 1;  // last chunk with prefix <= k
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_prefix[mid] <= k) lo = mid; else hi = mid - 1; }
        int rem = k - s_prefix[lo];
        const uint4* mw = reinterpret_cast<const uint4*>(masks + (int64_t)lo * 32);
        uint32_t m[32];
#pragma unroll
        for (int q = 0; q < 8; q++) { const uint4 v = __ldg(mw + q); m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w; }
        int w = 31; uint32_t word = m[31]; bool found = false;
#pragma unroll
        for (int q = 0; q < 32; q++) {
            const int c = __popc(m[q]);
            if (!found) { if (rem < c) { found = true; w = q; word = m[q]; } else rem -= c; }
        }
        const int bit = (total > 0 && found) ? (int)__fns(word, 0, rem + 1) : 0;  // (an empty grid degenerates to one cell, harmless: see networks.py)
        idx = (uint32_t)lo * kCellChunk + (uint32_t)w * 32u + (uint32_t)(bit & 31);
    }
    if (SORTED) {
        idx_tmp[j] = idx;
        atomicAdd(hist + idx, 1);
        return;
    }
    indices[j] = (int64_t)idx;
    cell_position(idx, rnd + 3 * j, inv_gm1, s_minus_half, half, xyzs + 3 * j);
}
// counting sort over the cells, pass 2: totals of the histogram's 1024-cell chunks (one warp per chunk) ...
__global__ void __launch_bounds__(256) hist_chunk_sum_kernel(const int32_t* __restrict__ hist, int64_t n_cells, int n_chunks, int32_t* __restrict__ chunk_tot) {
    const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    int sum = 0;
    for (int w = 0; w < 32; w++) { const int64_t i = chunk * kCellChunk + w * 32 + lane; sum += i < n_cells ? hist[i] : 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
    if (lane == 0) chunk_tot[chunk] = sum;
}
// ... pass 3: the histogram becomes the table of first output slots (exclusive prefix inside the chunk + the chunk's base) ...
__global__ void __launch_bounds__(256) hist_chunk_scan_kernel(int32_t* __restrict__ hist, int64_t n_cells, int n_chunks, const int32_t* __restrict__ chunk_base) {
    const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    int run = chunk_base[chunk];
    for (int w = 0; w < 32; w++) {
        const int64_t i = chunk * kCellChunk + w * 32 + lane;
        const int v = i < n_cells ? hist[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += t; }
        if (i < n_cells) hist[i] = run + inc - v;
        run += __shfl_sync(kFull, inc, 31);
    }
}
// ... pass 4: every draw takes the next slot of its cell (draws of one cell land next to each other in arbitrary order, each with
// the jitter of its own draw)
__global__ void __launch_bounds__(256) place_cells_kernel(const uint32_t* __restrict__ idx_tmp, int32_t* __restrict__ cursor, int64_t n,
                                                          const float* __restrict__ rnd, float inv_gm1, float s_minus_half, float half,
                                                          int64_t* __restrict__ indices, float* __restrict__ xyzs) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t idx = idx_tmp[j];
    const int64_t p = atomicAdd(cursor + idx, 1);
    indices[p] = (int64_t)idx;
    cell_position(idx, rnd + 3 * j, inv_gm1, s_minus_half, half, xyzs + 3 * p);
}
// dst[indices[i]] = src[i] (density_grid_tmp[c, indices] = density of networks.py:268; a cell drawn twice keeps one of its values)
__global__ void __launch_bounds__(256) scatter_f32_kernel(float* __restrict__ dst, const int64_t* __restrict__ indices, const float* __restrict__ src, int64_t n,
                                                          int64_t n_dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = indices[i];
    if (o >= 0 && o < n_dst) dst[o] = src[i];
}

// networks.py:273-279: grid = where(grid < 0, grid, max(grid * decay, tmp)) in place, plus the sum / count of the positive
// cells of the result (for the mean that caps the occupancy threshold).  Per-block partials in double, fixed order.
__global__ void __launch_bounds__(256) grid_update_kernel(float* __restrict__ grid, const float* __restrict__ tmp, const float* __restrict__ decay_cells,
                                                          float decay, int64_t n, double* __restrict__ part_sum, int64_t* __restrict__ part_cnt) {
    __shared__ double s_sum[8]; __shared__ int64_t s_cnt[8];
    double sum = 0.0; int64_t cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grid[i];
        float ng = g;
        if (!(g < 0.0f)) { ng = fmaxf(__fmul_rn(g, decay_cells ? decay_cells[i] : decay), tmp[i]); grid[i] = ng; }
        if (ng > 0.0f) { sum += (double)ng; cnt++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(kFull, sum, o); cnt += __shfl_xor_sync(kFull, cnt, o); }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0; int64_t b = 0;
        for (int k = 0; k < 8; k++) { a += s_sum[k]; b += s_cnt[k]; }
        part_sum[blockIdx.x] = a; part_cnt[blockIdx.x] = b;
    }
}
// threshold = min(mean of the positive cells, density_threshold) (networks.py:280-281), written to thr_out[0]
__global__ void __launch_bounds__(256) grid_threshold_kernel(const double* __restrict__ part_sum, const int64_t* __restrict__ part_cnt, int n_parts,
                                                             float density_threshold, float* __restrict__ thr_out) {
    __shared__ double s_sum[256]; __shared__ int64_t s_cnt[256];
    double a = 0.0; int64_t b = 0;
    for (int k = threadIdx.x; k < n_parts; k += 256) { a += part_sum[k]; b += part_cnt[k]; }
    s_sum[threadIdx.x] = a; s_cnt[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.0; b = 0;
        for (int k = 0; k < 256; k++) { a += s_sum[k]; b += s_cnt[k]; }
        // no positive cell: torch's mean of an empty selection is NaN and min(NaN, thr) in Python keeps NaN -> nothing is occupied
        const float mean = b > 0 ? (float)(a / (double)b) : NAN;
        thr_out[0] = mean < density_threshold ? mean : (b > 0 ? density_threshold : NAN);
    }
}
// packbits with the threshold read from device memory (no host round trip between the refresh and the packing)
__global__ void __launch_bounds__(256) packbits_devthr_kernel(const float* __restrict__ grid, const float* __restrict__ thr_dev, uint8_t* __restrict__ bits, int64_t n_bytes) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b0 = w * 4;
    if (b0 >= n_bytes) return;
    const float thr = thr_dev[0];
    const int nb = (int)min((int64_t)4, n_bytes - b0);
    for (int b = 0; b < nb; b++) {
        uint32_t byte = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) byte |= (grid[(b0 + b) * 8 + i] > thr) ? (1u << i) : 0u;
        bits[b0 + b] = (uint8_t)byte;
    }
}

// ---------------------------------------------------------------------------------------------- march (train)
// Pass 1 (raymarching.cu:184-234).  Counts go to rays_a[r][2]; if t_scratch != nullptr the parameter t of sample i
// of ray r is recorded at t_scratch[r*max_samples + i] so that pass 2 needs no second march.
__global__ void __launch_bounds__(128) march_train_count_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                const float* __restrict__ hits_t, int64_t n_rays,
                                                                const uint8_t* __restrict__ bitfield, ArnMarchConsts c,
                                                                const float* __restrict__ noise, int64_t* __restrict__ rays_a,
                                                                float* __restrict__ t_scratch, int32_t* __restrict__ counts) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
    const float t1 = arn_jitter_start(c, hits_t[2 * r], noise[r]);
    const float t2 = hits_t[2 * r + 1];
    float* rec = t_scratch ? t_scratch + r * c.max_samples : nullptr;
    float t = t1; int N = 0;
    while (0 <= t && t < t2 && N < c.max_samples) {
        float x, y, z, dt;
        if (arn_march_eval(c, ray, bitfield, t, x, y, z, dt)) {
            if (rec) rec[N] = t;
            t = __fadd_rn(t, dt); N++;
        }
    }
    if (counts) counts[r] = N; else rays_a[3 * r + 2] = N;
}

// Pass 1, warp-cooperative form (arn_march_core.h, "Window form"): one WARP per ray, 32 chain points per turn,
// ballot/popc compaction of the occupied visited points into t_scratch.  Same counts and t values as the kernel above.
// CONST_DT: exp_step_factor == 0 (calc_dt is the constant dt_lo).  FAST: cascades == 1 and grid_size <= 256.
template <bool CONST_DT, bool FAST>
__global__ void __launch_bounds__(256) march_train_count_warp_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                     const float* __restrict__ hits_t, int64_t n_rays,
                                                                     const uint8_t* __restrict__ bitfield, ArnMarchConsts c,
                                                                     const float* __restrict__ noise, int64_t* __restrict__ rays_a,
                                                                     float* __restrict__ t_scratch, int32_t* __restrict__ counts) {
    pdl_enter();
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rays) return;  // warp-uniform
    const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
    const float t1 = arn_jitter_start(c, hits_t[2 * r], noise[r]);
    const float t2 = hits_t[2 * r + 1];
    float* rec = t_scratch ? t_scratch + r * c.max_samples : nullptr;
    int N = 0;
    if (0 <= t1 && t1 < t2) {
        // lane j <- j steps along the chain from t1; afterwards every lane advances 32 steps per window
        float t = t1;
#pragma unroll
        for (int j = 0; j < 31; j++) {
            const float tn = __fadd_rn(t, CONST_DT ? c.dt_lo : arn_calc_dt(c, t));
            if (lane > j) t = tn;
        }
        float pending = -INFINITY;  // skip target carried over from the previous window
        for (;;) {
            float x, y, z, dt, tgt;
            const bool occ = arn_march_probe<FAST, FAST>(c, ray, bitfield, t, x, y, z, dt, tgt);
            const unsigned valid = __ballot_sync(kFull, t < t2);  // a prefix: the chain is increasing
            const unsigned occm = __ballot_sync(kFull, occ);
            const int s0 = __popc(__ballot_sync(kFull, t < pending));  // first lane not passed over by the carried skip
            // next visit after this lane: first k > lane with t_k >= tgt (32: beyond the window); occupied: lane + 1
            int lo = lane + 1, hi = 32;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi) >> 1;
                const float tv = __shfl_sync(kFull, t, mid & 31);
                if (lo < hi) { if (tv < tgt) lo = mid + 1; else hi = mid; }
            }
            int R = occ ? lane + 1 : lo;
            // lanes reachable from s0 along the next pointers (pointer doubling)
            unsigned M = 1u << lane;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const unsigned Mo = __shfl_sync(kFull, M, R & 31);
                const int Ro = __shfl_sync(kFull, R, R & 31);
                if (R < 32) { M |= Mo; R = Ro; }
            }
            const unsigned vis = s0 < 32 ? __shfl_sync(kFull, M, s0 & 31) : 0u;
            unsigned emit = vis & occm & valid;
            const int rem = c.max_samples - N;
            bool done = valid != kFull;
            if (__popc(emit) >= rem) {  // the sample budget ends the march inside this window
                done = true;
                const int rank_all = __popc(emit & ((1u << lane) - 1u));
                emit = __ballot_sync(kFull, ((emit >> lane) & 1u) && rank_all < rem);
            }
            if (rec && ((emit >> lane) & 1u)) rec[N + __popc(emit & ((1u << lane) - 1u))] = t;
            N += __popc(emit);
            if (done) break;
            // carry: the last visited lane either is lane 31 and occupied, or is empty with a target beyond the window
            if (vis) {
                const int last = 31 - __clz(vis);
                pending = __shfl_sync(kFull, occ ? -INFINITY : tgt, last);
            }
#pragma unroll
            for (int k = 0; k < 32; k++) t = __fadd_rn(t, CONST_DT ? c.dt_lo : arn_calc_dt(c, t));
        }
    }
    if (lane == 0) {
        if (counts) counts[r] = N; else rays_a[3 * r + 2] = N;
    }
}

// Exclusive scan of compact per-ray counts -> rays_a[r] = (r, start, N), counter = (total, n_rays).  One CTA per 1024
// rays: it sums the counts of all earlier rays (coalesced 4-byte loads), scans its own 1024 and writes its rays_a rows
// as 3072 consecutive 8-byte stores.  (A single CTA walking the 24-byte rows of rays_a is bound by one SM's
// sector throughput: 20 us for 8192 rays.)
__global__ void __launch_bounds__(1024) rays_scan_compact_kernel(const int32_t* __restrict__ counts, int64_t n_rays, int64_t* __restrict__ rays_a,
                                                                 int32_t* __restrict__ counter) {
    pdl_enter();
    __shared__ int64_t warp_pre[32], warp_inc[32];
    __shared__ int64_t s_start[1024];
    __shared__ int32_t s_cnt[1024];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r_base = (int64_t)blockIdx.x * 1024, r = r_base + threadIdx.x;
    int64_t pre = 0;
    for (int64_t i = threadIdx.x; i < r_base; i += 1024) pre += counts[i];
    const int32_t mine = r < n_rays ? counts[r] : 0;
    int64_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += u; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(kFull, pre, o);
    if (lane == 31) warp_inc[wid] = inc;
    if (lane == 0) warp_pre[wid] = pre;
    __syncthreads();
    if (wid == 0) {
        int64_t s = warp_inc[lane], p = warp_pre[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, s, o); if (lane >= o) s += u; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(kFull, p, o);
        warp_inc[lane] = s;  // inclusive over warps
        warp_pre[lane] = p;  // every entry: total of the earlier blocks
    }
    __syncthreads();
    const int64_t before = warp_pre[0];
    s_start[threadIdx.x] = before + (wid ? warp_inc[wid - 1] : 0) + inc - mine;
    s_cnt[threadIdx.x] = mine;
    __syncthreads();
    const int64_t rows = min((int64_t)1024, n_rays - r_base);
    for (int64_t e = threadIdx.x; e < 3 * rows; e += 1024) {
        const int ray = (int)(e / 3), f = (int)(e % 3);
        rays_a[3 * r_base + e] = f == 0 ? r_base + ray : (f == 1 ? s_start[ray] : (int64_t)s_cnt[ray]);
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const int64_t tot = before + warp_inc[31];
        counter[0] = (int32_t)(tot > 0x7fffffff ? 0x7fffffff : tot);
        counter[1] = (int32_t)(n_rays > 0x7fffffff ? 0x7fffffff : n_rays);
    }
}

// Exclusive scan of the counts: rays_a[r] = (r, start, N); counter = (total, n_rays).  Single CTA: thread i owns the
// K = ceil(n_rays/1024) consecutive rays [i*K, (i+1)*K) (K independent loads in flight), one block scan of the
// per-thread totals, one pass of stores.
__global__ void __launch_bounds__(1024) rays_scan_kernel(int64_t* __restrict__ rays_a, int64_t n_rays, int32_t* __restrict__ counter) {
    __shared__ int64_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t K = (n_rays + blockDim.x - 1) / blockDim.x;
    const int64_t r0 = min(n_rays, (int64_t)threadIdx.x * K), r1 = min(n_rays, r0 + K);
    int64_t mine = 0;
    if (K <= 8) {
        int64_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (r0 + k < r1) ? rays_a[3 * (r0 + k) + 2] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++) mine += v[k];
    } else {
        for (int64_t r = r0; r < r1; r++) mine += rays_a[3 * r + 2];
    }
    int64_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int64_t s = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, s, o); if (lane >= o) s += u; }
        warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    int64_t start = (wid ? warp_sums[wid - 1] : 0) + inc - mine;
    for (int64_t r = r0; r < r1; r++) {
        const int64_t nr = rays_a[3 * r + 2];
        rays_a[3 * r] = r; rays_a[3 * r + 1] = start;
        start += nr;
    }
    if (threadIdx.x == 0) {
        const int64_t tot = warp_sums[31];
        counter[0] = (int32_t)(tot > 0x7fffffff ? 0x7fffffff : tot);
        counter[1] = (int32_t)(n_rays > 0x7fffffff ? 0x7fffffff : n_rays);
    }
}

// Pass 2, parallel form: one thread per sample.  The ray of sample s is found by binary search over the starts.
__global__ void __launch_bounds__(256) march_train_emit_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                               int64_t n_rays, ArnMarchConsts c, const int64_t* __restrict__ rays_a,
                                                               const float* __restrict__ t_scratch, int64_t total,
                                                               const int32_t* __restrict__ total_dev,
                                                               float* __restrict__ xyzs, float* __restrict__ dirs,
                                                               float* __restrict__ deltas, float* __restrict__ ts) {
    pdl_enter();
    if (total_dev) total = min(total, (int64_t)*total_dev);  // device-side count (fused step): `total` is the capacity
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
    // largest r with start[r] <= s and N[r] > 0 : starts are non-decreasing, so search the last start <= s
    int64_t lo = 0, hi = n_rays - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (rays_a[3 * mid + 1] <= s) lo = mid; else hi = mid - 1;
    }
    const int64_t r = lo;  // rays with N == 0 share their start with the next ray; "last start <= s" skips them
    const int i = (int)(s - rays_a[3 * r + 1]);
    const float t = t_scratch[r * c.max_samples + i];
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    xyzs[3 * s] = __fmaf_rn(dx, t, ox); xyzs[3 * s + 1] = __fmaf_rn(dy, t, oy); xyzs[3 * s + 2] = __fmaf_rn(dz, t, oz);
    dirs[3 * s] = dx; dirs[3 * s + 1] = dy; dirs[3 * s + 2] = dz;
    ts[s] = t; deltas[s] = arn_calc_dt(c, t);
    }
}

// Pass 2, re-march form (raymarching.cu:236-279), used when the caller gives no t scratch.
__global__ void __launch_bounds__(128) march_train_remarch_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                  const float* __restrict__ hits_t, int64_t n_rays,
                                                                  const uint8_t* __restrict__ bitfield, ArnMarchConsts c,
                                                                  const float* __restrict__ noise, const int64_t* __restrict__ rays_a,
                                                                  float* __restrict__ xyzs, float* __restrict__ dirs,
                                                                  float* __restrict__ deltas, float* __restrict__ ts) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
    const float t1 = arn_jitter_start(c, hits_t[2 * r], noise[r]);
    const float t2 = hits_t[2 * r + 1];
    const int64_t start = rays_a[3 * r + 1]; const int N = (int)rays_a[3 * r + 2];
    float t = t1; int samples = 0;
    while (t < t2 && samples < N) {
        float x, y, z, dt;
        if (arn_march_eval(c, ray, bitfield, t, x, y, z, dt)) {
            const int64_t s = start + samples;
            xyzs[3 * s] = x; xyzs[3 * s + 1] = y; xyzs[3 * s + 2] = z;
            dirs[3 * s] = ray.dx; dirs[3 * s + 1] = ray.dy; dirs[3 * s + 2] = ray.dz;
            ts[s] = t; deltas[s] = dt;
            t = __fadd_rn(t, dt); samples++;
        }
    }
}

// ---------------------------------------------------------------------------------------------- march (test)
// raymarching.cu:335-404
__global__ void __launch_bounds__(128) march_test_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         float* __restrict__ hits_t, const int64_t* __restrict__ alive,
                                                         int64_t n_alive, const uint8_t* __restrict__ bitfield, ArnMarchConsts c,
                                                         int S, float* __restrict__ xyzs, float* __restrict__ dirs,
                                                         float* __restrict__ deltas, float* __restrict__ ts, int32_t* __restrict__ n_eff) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int64_t r = alive[n];
    const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
    float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
    int s = 0;
    float t_resume = t; bool moved = false;
    while (t < t2 && s < S) {
        float x, y, z, dt;
        if (arn_march_eval(c, ray, bitfield, t, x, y, z, dt)) {
            const int64_t o = n * S + s;
            xyzs[3 * o] = x; xyzs[3 * o + 1] = y; xyzs[3 * o + 2] = z;
            dirs[3 * o] = ray.dx; dirs[3 * o + 1] = ray.dy; dirs[3 * o + 2] = ray.dz;
            ts[o] = t; deltas[o] = dt;
            t = __fadd_rn(t, dt);
            t_resume = t; moved = true;  // raymarching.cu:386: only an occupied step moves the resume point
            s++;
        }
    }
    if (moved) hits_t[2 * r] = t_resume;
    n_eff[n] = s;
    for (int k = s; k < S; k++) {  // zero padding (the reference memsets the whole buffers, :421-426)
        const int64_t o = n * S + k;
        xyzs[3 * o] = 0.f; xyzs[3 * o + 1] = 0.f; xyzs[3 * o + 2] = 0.f;
        dirs[3 * o] = 0.f; dirs[3 * o + 1] = 0.f; dirs[3 * o + 2] = 0.f;
        ts[o] = 0.f; deltas[o] = 0.f;
    }
}

// Test-time render iteration, compact form (arn_render_test_iter): the march records only t / dt per (alive ray, slot)
// and the per-ray count; a scan lays the valid samples out contiguously, and this kernel materialises their positions.
__global__ void __launch_bounds__(128) march_test_lite_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                              float* __restrict__ hits_t, const int64_t* __restrict__ alive,
                                                              int64_t n_alive, const uint8_t* __restrict__ bitfield, ArnMarchConsts c,
                                                              int S, float* __restrict__ deltas, float* __restrict__ ts, int32_t* __restrict__ n_eff) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int64_t r = alive[n];
    const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
    float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
    int s = 0;
    float t_resume = t; bool moved = false;
    while (t < t2 && s < S) {
        float x, y, z, dt;
        if (arn_march_eval(c, ray, bitfield, t, x, y, z, dt)) {
            const int64_t o = n * S + s;
            ts[o] = t; deltas[o] = dt;
            t = __fadd_rn(t, dt);
            t_resume = t; moved = true;  // raymarching.cu:386: only an occupied step moves the resume point
            s++;
        }
    }
    if (moved) hits_t[2 * r] = t_resume;
    n_eff[n] = s;
}
// one thread per (alive ray, slot): valid slots write their sample into the compact list at start[ray] + slot
__global__ void __launch_bounds__(256) emit_test_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        const int64_t* __restrict__ alive, int64_t n_alive, int S,
                                                        const int64_t* __restrict__ rays_a, const float* __restrict__ ts,
                                                        float* __restrict__ xyzs, float* __restrict__ dirs) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_alive * S) return;
    const int64_t n = e / S; const int sl = (int)(e % S);
    if (sl >= (int)rays_a[3 * n + 2]) return;
    const int64_t r = alive[n], o = rays_a[3 * n + 1] + sl;
    const float t = ts[e];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    xyzs[3 * o] = __fmaf_rn(dx, t, rays_o[3 * r]); xyzs[3 * o + 1] = __fmaf_rn(dy, t, rays_o[3 * r + 1]); xyzs[3 * o + 2] = __fmaf_rn(dz, t, rays_o[3 * r + 2]);
    dirs[3 * o] = dx; dirs[3 * o + 1] = dy; dirs[3 * o + 2] = dz;
}

// ---------------------------------------------------------------------------------------------- compositing
// alpha = 1 - __expf(-sigma*delta): SASS of the reference is FMUL s*d ; FMUL -1.44269502 ; MUFU.EX2 (volumerendering.cu:27)
__device__ __forceinline__ float alpha_of(float sigma, float delta) {
    return __fsub_rn(1.0f, __expf(-sigma * delta));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(kFull, v, o); if (lane >= o) v += u; }
    return v;
}
__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(kFull, v, o); if (lane >= o) v *= u; }
    return v;
}

// NeRFLoss ('raw', losses.py:63-82) + background blend (rendering.py:287-296) as the per-ray epilogue of the compositing
// kernel: the values are in lane 0's registers when the ray is done, so the separate loss kernel and its re-read of the
// per-ray outputs disappear from the fused training step.  Same arithmetic as nerf_loss_kernel (arn_train.cu).
struct LossEpilogue {
    const float* target; float bg[3]; float lambda_opacity, lambda_depth, grid_scale, grad_scale;
    float* rgb_out; float* dL_drgb; float* dL_dopacity; float* dL_ddepth; float* loss_out;
};
// store: this lane writes the per-ray results; every calling lane gets the gradients (g_rgb, g_op, g_d) back
__device__ __forceinline__ float ray_loss(const LossEpilogue& L, int64_t r, int64_t n_rays, const float c[3], float o, float dep, bool store,
                                          const float tgt[3], float g_rgb[3], float& g_op, float& g_d) {
    const float inv_r = 1.0f / (float)n_rays, inv_3r = inv_r / 3.0f;
    float loss = 0.0f;
    g_op = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float est = c[k] + L.bg[k] * (1.0f - o);
        if (store && L.rgb_out) L.rgb_out[3 * r + k] = est;
        const float den = est + 1e-3f;
        const float e = (est - tgt[k]) / den;
        loss += e * e * inv_3r;
        const float g = 2.0f * e / den * inv_3r * L.grad_scale;
        g_rgb[k] = g;
        if (store) L.dL_drgb[3 * r + k] = g;
        g_op -= L.bg[k] * g;
    }
    const float oe = o + 1e-10f;
    loss += L.lambda_opacity * (-oe * logf(oe)) * inv_r;
    g_op += L.lambda_opacity * (-logf(oe) - 1.0f) * inv_r * L.grad_scale;
    if (store) L.dL_dopacity[r] = g_op;
    g_d = 0.0f;
    if (L.lambda_depth != 0.0f) {
        const float v = dep / L.grid_scale + 1e-10f;
        loss += -L.lambda_depth * logf(fminf(v, 1.0f)) * inv_r;
        if (v < 1.0f) g_d = -L.lambda_depth / v / L.grid_scale * inv_r * L.grad_scale;
    }
    if (store) L.dL_ddepth[r] = g_d;
    return loss;
}

// Samples per lane and per turn of the training compositing loops: a warp takes kSPL * 32 consecutive samples of its ray at
// once -- lane l the samples kSPL*l .. kSPL*l + kSPL-1, scanned inside the lane, the lanes' totals by ONE warp scan per quantity.
// A batch's rays are few and long (W1: 2 000 of 8 192 rays hit the object, 120 samples each on average, 437 at most) and a turn
// is a chain of dependent shuffles.  Measured on W1 (fused forward + loss + backward launch, 64-thread blocks): kSPL = 1 / 2 / 4 / 8
// -> 25.3 / 20.2 / 21.4 / 29.1 us (more samples per lane = fewer turns but more registers = fewer rays resident).
constexpr int kSPL = 2;
// ... and the blocks are SMALL (two rays): a zero-sample ray's warp is gone at once, but its block holds its registers until the
// block's longest ray is done -- with eight rays per block three quarters of the resident warps were such idle ones
// (256 / 64 / 32 threads per block at kSPL = 4: 28.9 / 21.4 / 24.1 us).
constexpr int kCompBlock = 64;

// volumerendering.cu:86-150 for one ray, by one warp.  The reference's in-kernel serial thrust scan of dL_dws*ws (:118-121)
// becomes a warp reduction (total) plus a running warp scan.  Shared by composite_train_bw_kernel and by the fused
// forward + loss + backward of the training step.
__device__ __forceinline__ void composite_bw_ray(int lane, int64_t start, int N, float R, float G, float B, float O, float D,
                                                 float gR, float gG, float gB, float gO, float gD, const float* __restrict__ dL_dws,
                                                 const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ws,
                                                 const float* __restrict__ deltas, const float* __restrict__ ts, float T_thr,
                                                 float* __restrict__ dL_dsigmas, float* __restrict__ dL_drgbs) {
    float ww_total = 0.0f;
    if (dL_dws) {
        float p = 0.0f;
        for (int i = lane; i < N; i += 32) p += dL_dws[start + i] * ws[start + i];
        ww_total = warp_sum(p);
    }
    float T = 1.0f, r = 0.f, g = 0.f, b = 0.f, d = 0.f, ww = 0.f;
    bool done = false; int base = 0;
    for (; base < N && !done; base += 32 * kSPL) {
        const int i0 = base + kSPL * lane;
        float sg[kSPL], dl[kSPL], cr[kSPL], cg[kSPL], cb[kSPL], ct[kSPL], gw[kSPL], wsv[kSPL];
#pragma unroll
        for (int k = 0; k < kSPL; k++) {
            const bool in = i0 + k < N; const int64_t q = start + i0 + k;
            sg[k] = in ? sigmas[q] : 0.f; dl[k] = in ? deltas[q] : 0.f; ct[k] = in ? ts[q] : 0.f;
            cr[k] = in ? rgbs[3 * q] : 0.f; cg[k] = in ? rgbs[3 * q + 1] : 0.f; cb[k] = in ? rgbs[3 * q + 2] : 0.f;
            gw[k] = (in && dL_dws) ? dL_dws[q] : 0.f; wsv[k] = (in && dL_dws) ? ws[q] : 0.f;
        }
        float al[kSPL], pp[kSPL];  // alpha, product of (1 - alpha) over the lane's samples up to and including k
#pragma unroll
        for (int k = 0; k < kSPL; k++) {
            al[k] = i0 + k < N ? alpha_of(sg[k], dl[k]) : 0.0f;
            const float om = __fsub_rn(1.0f, al[k]);
            pp[k] = k ? pp[k - 1] * om : om;
        }
        const float incl = warp_incl_prod(pp[kSPL - 1], lane);
        float excl = __shfl_up_sync(kFull, incl, 1); if (lane == 0) excl = 1.0f;
        const float Tl = T * excl;  // transmittance in front of the lane's first sample
        float Tb[kSPL], Ta[kSPL]; int first = kSPL;
#pragma unroll
        for (int k = kSPL - 1; k >= 0; k--) {
            Tb[k] = k ? Tl * pp[k - 1] : Tl; Ta[k] = Tl * pp[k];
            if (i0 + k < N && Ta[k] <= T_thr) first = k;
        }
        const unsigned term = __ballot_sync(kFull, first < kSPL);
        int last = 32 * kSPL - 1;
        if (term) { const int tl = __ffs(term) - 1; last = kSPL * tl + __shfl_sync(kFull, first, tl); }
        float w[kSPL], s_r[kSPL], s_g[kSPL], s_b[kSPL], s_d[kSPL], s_w[kSPL];  // in-lane inclusive sums
#pragma unroll
        for (int k = 0; k < kSPL; k++) {
            const bool use = i0 + k < N && kSPL * lane + k <= last;
            w[k] = use ? al[k] * Tb[k] : 0.0f;
            s_r[k] = (k ? s_r[k - 1] : 0.f) + w[k] * cr[k]; s_g[k] = (k ? s_g[k - 1] : 0.f) + w[k] * cg[k];
            s_b[k] = (k ? s_b[k - 1] : 0.f) + w[k] * cb[k]; s_d[k] = (k ? s_d[k - 1] : 0.f) + w[k] * ct[k];
            s_w[k] = (k ? s_w[k - 1] : 0.f) + gw[k] * wsv[k];
        }
        const float i_r = warp_incl_sum(s_r[kSPL - 1], lane), i_g = warp_incl_sum(s_g[kSPL - 1], lane), i_b = warp_incl_sum(s_b[kSPL - 1], lane);
        const float i_d = warp_incl_sum(s_d[kSPL - 1], lane), i_w = dL_dws ? warp_incl_sum(s_w[kSPL - 1], lane) : 0.0f;
        const float e_r = r + (i_r - s_r[kSPL - 1]), e_g = g + (i_g - s_g[kSPL - 1]), e_b = b + (i_b - s_b[kSPL - 1]);
        const float e_d = d + (i_d - s_d[kSPL - 1]), e_w = ww + (i_w - s_w[kSPL - 1]);
#pragma unroll
        for (int k = 0; k < kSPL; k++) {
            if (i0 + k >= N) continue;
            const int64_t q = start + i0 + k;
            float o_r = 0.f, o_g = 0.f, o_b = 0.f, o_s = 0.f;
            if (kSPL * lane + k <= last) {
                const float pr = e_r + s_r[k], pg = e_g + s_g[k], pb = e_b + s_b[k], pd = e_d + s_d[k], pww = e_w + s_w[k];
                o_r = gR * w[k]; o_g = gG * w[k]; o_b = gB * w[k];
                o_s = dl[k] * (gR * (cr[k] * Ta[k] - (R - pr)) + gG * (cg[k] * Ta[k] - (G - pg)) + gB * (cb[k] * Ta[k] - (B - pb)) +
                               gO * (1.0f - O) + gD * (ct[k] * Ta[k] - (D - pd)) + Ta[k] * gw[k] - (ww_total - pww));
            }
            dL_drgbs[3 * q] = o_r; dL_drgbs[3 * q + 1] = o_g; dL_drgbs[3 * q + 2] = o_b;
            dL_dsigmas[q] = o_s;
        }
        if (term) done = true;
        T = __shfl_sync(kFull, Ta[kSPL - 1], 31);
        r += __shfl_sync(kFull, i_r, 31); g += __shfl_sync(kFull, i_g, 31); b += __shfl_sync(kFull, i_b, 31);
        d += __shfl_sync(kFull, i_d, 31); ww += __shfl_sync(kFull, i_w, 31);
    }
    for (int i = base + lane; i < N; i += 32) {  // zero-init in the reference (:171-172)
        const int64_t q = start + i;
        dL_drgbs[3 * q] = 0.f; dL_drgbs[3 * q + 1] = 0.f; dL_drgbs[3 * q + 2] = 0.f; dL_dsigmas[q] = 0.f;
    }
}

// volumerendering.cu:5-44, one warp per rays_a row.
template <bool LOSS>
__global__ void __launch_bounds__(kCompBlock) composite_train_fw_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                 const float* __restrict__ deltas, const float* __restrict__ ts,
                                                                 const int64_t* __restrict__ rays_a, int64_t n_rays, float T_thr,
                                                                 int64_t* __restrict__ total_samples, float* __restrict__ opacity,
                                                                 float* __restrict__ depth, float* __restrict__ rgb, float* __restrict__ ws,
                                                                 const LossEpilogue L, int64_t n_samples,
                                                                 float* __restrict__ bw_dsigmas, float* __restrict__ bw_drgbs) {
    pdl_enter();
    __shared__ float s_loss[kCompBlock / 32];
    if (LOSS && threadIdx.x < kCompBlock / 32) s_loss[threadIdx.x] = 0.0f;
    if (LOSS) __syncthreads();
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool live = n < n_rays;
    if (!LOSS && !live) return;
    int64_t ray_idx = 0, start = 0; int N = 0;
    if (live) {
        ray_idx = rays_a[3 * n]; start = rays_a[3 * n + 1]; N = (int)rays_a[3 * n + 2];
        // never read past the sample buffers (a caller-chosen sample capacity smaller than the march: the tail is dropped)
        N = (int)max((int64_t)0, min((int64_t)N, n_samples - start));
    }
    float T = 1.0f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_o = 0.f;
    int64_t samples = N; bool done = false;
    int base = 0;
    float tgt[3] = {0.f, 0.f, 0.f};  // the ray's target colour, requested in front of the sample loop (used by the loss behind it)
    if (LOSS && live) { tgt[0] = L.target[3 * ray_idx]; tgt[1] = L.target[3 * ray_idx + 1]; tgt[2] = L.target[3 * ray_idx + 2]; }
    for (; base < N && !done; base += 32 * kSPL) {
        const int i0 = base + kSPL * lane;
        float al[kSPL], pp[kSPL], cr[kSPL], cg[kSPL], cb[kSPL], ct[kSPL];
#pragma unroll
        for (int k = 0; k < kSPL; k++) {
            const bool in = i0 + k < N; const int64_t q = start + i0 + k;
            const float sg = in ? sigmas[q] : 0.f, dl = in ? deltas[q] : 0.f;
            ct[k] = in ? ts[q] : 0.f; cr[k] = in ? rgbs[3 * q] : 0.f; cg[k] = in ? rgbs[3 * q + 1] : 0.f; cb[k] = in ? rgbs[3 * q + 2] : 0.f;
            al[k] = in ? alpha_of(sg, dl) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < kSPL; k++) { const float om = __fsub_rn(1.0f, al[k]); pp[k] = k ? pp[k - 1] * om : om; }
        const float incl = warp_incl_prod(pp[kSPL - 1], lane);
        float excl = __shfl_up_sync(kFull, incl, 1); if (lane == 0) excl = 1.0f;
        const float Tl = T * excl;  // transmittance in front of the lane's first sample
        int first = kSPL;
#pragma unroll
        for (int k = kSPL - 1; k >= 0; k--) if (i0 + k < N && Tl * pp[k] <= T_thr) first = k;
        const unsigned term = __ballot_sync(kFull, first < kSPL);
        int last = 32 * kSPL - 1;  // the terminating sample still contributes (:37-40)
        if (term) { const int tl = __ffs(term) - 1; last = kSPL * tl + __shfl_sync(kFull, first, tl); }
        float p_r = 0.f, p_g = 0.f, p_b = 0.f, p_d = 0.f, p_o = 0.f;
#pragma unroll
        for (int k = 0; k < kSPL; k++) {
            const bool in = i0 + k < N;
            const float w = (in && kSPL * lane + k <= last) ? al[k] * (k ? Tl * pp[k - 1] : Tl) : 0.0f;
            if (in) ws[start + i0 + k] = w;
            p_r += w * cr[k]; p_g += w * cg[k]; p_b += w * cb[k]; p_d += w * ct[k]; p_o += w;
        }
        acc_r += warp_sum(p_r); acc_g += warp_sum(p_g); acc_b += warp_sum(p_b);
        acc_d += warp_sum(p_d); acc_o += warp_sum(p_o);
        if (term) { done = true; samples = base + last; }  // break happens before samples++ (:40-41)
        T = __shfl_sync(kFull, Tl * pp[kSPL - 1], 31);
    }
    for (int i = base + lane; i < N; i += 32) ws[start + i] = 0.0f;  // after termination (reference: zero-init, :59)
    if (lane == 0 && live) {
        opacity[ray_idx] = acc_o; depth[ray_idx] = acc_d;
        rgb[3 * ray_idx] = acc_r; rgb[3 * ray_idx + 1] = acc_g; rgb[3 * ray_idx + 2] = acc_b;
        total_samples[ray_idx] = samples;
    }
    if (LOSS && live) {
        // every lane evaluates the ray's loss terms (the sums are warp-uniform), lane 0 stores them; with bw_dsigmas the warp
        // then runs the ray's backward at once (composite_train_bw's arithmetic on the samples it has just read): the fused
        // training step needs no second compositing launch and no re-read of rays_a / the per-ray outputs
        const float c[3] = {acc_r, acc_g, acc_b};
        float g_rgb[3], g_op, g_d;
        const float l = ray_loss(L, ray_idx, n_rays, c, acc_o, acc_d, lane == 0, tgt, g_rgb, g_op, g_d);
        if (lane == 0) s_loss[threadIdx.x >> 5] = l;
        if (bw_dsigmas && N > 0)
            composite_bw_ray(lane, start, N, acc_r, acc_g, acc_b, acc_o, acc_d, g_rgb[0], g_rgb[1], g_rgb[2], g_op, g_d, nullptr, sigmas, rgbs, ws, deltas, ts,
                             T_thr, bw_dsigmas, bw_drgbs);
    }
    if (LOSS) {  // one atomic per CTA
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.0f;
#pragma unroll
            for (int k = 0; k < kCompBlock / 32; k++) v += s_loss[k];
            atomicAdd(L.loss_out, v);
        }
    }
}

// volumerendering.cu:86-150, one warp per row.  The reference's in-kernel serial thrust scan of dL_dws*ws (:118-121)
// becomes a warp reduction (total) plus a running warp scan.
__global__ void __launch_bounds__(kCompBlock) composite_train_bw_kernel(const float* __restrict__ dL_dopacity, const float* __restrict__ dL_ddepth,
                                                                 const float* __restrict__ dL_drgb, const float* __restrict__ dL_dws,
                                                                 const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                 const float* __restrict__ ws, const float* __restrict__ deltas,
                                                                 const float* __restrict__ ts, const int64_t* __restrict__ rays_a,
                                                                 const float* __restrict__ opacity, const float* __restrict__ depth,
                                                                 const float* __restrict__ rgb, int64_t n_rays, float T_thr,
                                                                 float* __restrict__ dL_dsigmas, float* __restrict__ dL_drgbs, int64_t n_samples) {
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= n_rays) return;
    const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1];
    const int N = (int)max((int64_t)0, min(rays_a[3 * n + 2], n_samples - start));  // same clamp as the forward
    if (N <= 0) return;
    composite_bw_ray(lane, start, N, rgb[3 * ray_idx], rgb[3 * ray_idx + 1], rgb[3 * ray_idx + 2], opacity[ray_idx], depth[ray_idx],
                     dL_drgb[3 * ray_idx], dL_drgb[3 * ray_idx + 1], dL_drgb[3 * ray_idx + 2], dL_dopacity[ray_idx], dL_ddepth[ray_idx], dL_dws,
                     sigmas, rgbs, ws, deltas, ts, T_thr, dL_dsigmas, dL_drgbs);
}

// volumerendering.cu:204-248.  Chunks are short (S <= 64, usually 1..8): one thread per alive ray, registers for the
// running sums, one read-modify-write of the per-ray outputs at the end.
__global__ void __launch_bounds__(256) composite_test_fw_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                const float* __restrict__ deltas, const float* __restrict__ ts,
                                                                int64_t* __restrict__ alive, int64_t n_alive, int S, float T_thr,
                                                                const int32_t* __restrict__ n_eff, float* __restrict__ opacity,
                                                                float* __restrict__ depth, float* __restrict__ rgb) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int ne = n_eff[n];
    if (ne == 0) { alive[n] = -1; return; }
    const int64_t r = alive[n];
    float O = opacity[r];
    float T = __fsub_rn(1.0f, O);
    float cr = rgb[3 * r], cg = rgb[3 * r + 1], cb = rgb[3 * r + 2], D = depth[r];
    for (int s = 0; s < ne; s++) {
        const int64_t o = n * S + s;
        const float a = alpha_of(sigmas[o], deltas[o]);
        const float w = __fmul_rn(a, T);
        cr = __fmaf_rn(w, rgbs[3 * o], cr); cg = __fmaf_rn(w, rgbs[3 * o + 1], cg); cb = __fmaf_rn(w, rgbs[3 * o + 2], cb);
        D = __fmaf_rn(w, ts[o], D);
        O = __fadd_rn(O, w);
        T = __fmul_rn(T, __fsub_rn(1.0f, a));
        if (T <= T_thr) { alive[n] = -1; break; }
    }
    opacity[r] = O; depth[r] = D; rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
}

// composite_test_fw_kernel on the compact sample list: sigmas / rgbs at rays_a[n].start + s, deltas / ts in their padded
// slots.  Also writes keep[n] = 1 if the ray stays alive (the input of the alive-list compaction) and adds the
// effective samples of the iteration to *total (rendering.py:203).
__global__ void __launch_bounds__(256) composite_test_compact_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                     const float* __restrict__ deltas, const float* __restrict__ ts,
                                                                     const int64_t* __restrict__ alive, int64_t n_alive, int S, float T_thr,
                                                                     const int64_t* __restrict__ rays_a, float* __restrict__ opacity,
                                                                     float* __restrict__ depth, float* __restrict__ rgb, int32_t* __restrict__ keep,
                                                                     unsigned long long* __restrict__ total) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int ne = 0; bool live = false;
    if (n < n_alive) {
        ne = (int)rays_a[3 * n + 2];
        if (ne > 0) {
            live = true;
            const int64_t r = alive[n], c0 = rays_a[3 * n + 1];
            float O = opacity[r];
            float T = __fsub_rn(1.0f, O);
            float cr = rgb[3 * r], cg = rgb[3 * r + 1], cb = rgb[3 * r + 2], D = depth[r];
            for (int s = 0; s < ne; s++) {
                const int64_t o = n * S + s, q = c0 + s;
                const float a = alpha_of(sigmas[q], deltas[o]);
                const float w = __fmul_rn(a, T);
                cr = __fmaf_rn(w, rgbs[3 * q], cr); cg = __fmaf_rn(w, rgbs[3 * q + 1], cg); cb = __fmaf_rn(w, rgbs[3 * q + 2], cb);
                D = __fmaf_rn(w, ts[o], D);
                O = __fadd_rn(O, w);
                T = __fmul_rn(T, __fsub_rn(1.0f, a));
                if (T <= T_thr) { live = false; break; }
            }
            opacity[r] = O; depth[r] = D; rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
        }
        keep[n] = live ? 1 : 0;
    }
    // effective samples of this iteration: warp sum, one atomic per warp
    unsigned v = (unsigned)ne;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, (unsigned long long)v);
}
// alive_out[start[n]] = alive[n] for the rays that stay alive (order preserved: alive_indices[alive_indices >= 0])
__global__ void __launch_bounds__(256) alive_scatter_kernel(const int64_t* __restrict__ alive, int64_t n_alive, const int64_t* __restrict__ rays_a,
                                                            int64_t* __restrict__ alive_out) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    if (rays_a[3 * n + 2] > 0) alive_out[rays_a[3 * n + 1]] = alive[n];
}

// ---------------------------------------------------------------------------------------------- test loop, device-driven
// arn_render_test_step: the same iteration with the loop's control state ON THE DEVICE, so the host can queue iterations
// without reading anything back.  state = {n_alive, N_samples, samples requested so far, active, iterations done}; an
// iteration reads state_in and its last kernel writes state_out (double-buffered by the caller: a CTA of that kernel may
// still be reading state_in).  Every kernel is a grid-stride loop over the device-side count (the host sizes the grids
// from an upper bound of n_alive).  Producers of per-ray counts also write the sum of every chunk of 128 rays (`partial`),
// from which the scan kernels get their offset without re-reading all earlier counts.
constexpr int kStN = 0, kStS = 1, kStDone = 2, kStActive = 3, kStIters = 4, kStLive = 5;  // kStLive: iterations that started active

__device__ __forceinline__ int block_sum_128(int v, int* sm4) {  // sum over a 128-thread group (4 warps), result in every thread
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    const int w = (threadIdx.x >> 5) & 3;
    if ((threadIdx.x & 31) == 0) sm4[w] = v;
    __syncthreads();
    const int t = sm4[0] + sm4[1] + sm4[2] + sm4[3];
    __syncthreads();
    return t;
}

// Far clamp of a frame's rays, once, in front of the loop: every ray is marched to its end (same chain, same visits as the
// loop's march) and hits_t[r][1] is pulled back to the chain point behind its LAST occupied sample (to hits_t[r][0] for
// a ray that meets no occupied cell).  Nothing the loop computes changes -- the samples, their order, every N_eff and
// the kill pattern are those of the unclamped march, whose `t < t2` test now ends where it would have found nothing
// more -- but no iteration has to walk a ray's empty exit stretch (twice: once behind its last samples, once to find
// zero samples), which left each iteration waiting for a handful of threads doing ~150 dependent probes.
template <bool FAST>
__global__ void __launch_bounds__(128) march_test_far_clamp_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                   float* __restrict__ hits_t, int64_t n_rays,
                                                                   const uint8_t* __restrict__ bitfield, ArnMarchConsts c) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
    float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
    if (!(t < t2)) return;
    float t_end = t;
    while (t < t2) {
        float x, y, z, dt;
        if (arn_march_eval_t<FAST>(c, ray, bitfield, t, x, y, z, dt)) { t = __fadd_rn(t, dt); t_end = t; }
    }
    if (t_end < t2) hits_t[2 * r + 1] = t_end;
}

template <bool FAST>
__global__ void __launch_bounds__(128) march_test_dyn_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                             float* __restrict__ hits_t, const int64_t* __restrict__ alive,
                                                             const int32_t* __restrict__ state, const uint8_t* __restrict__ bitfield,
                                                             ArnMarchConsts c, float* __restrict__ deltas, float* __restrict__ ts,
                                                             int32_t* __restrict__ n_eff, int32_t* __restrict__ partial) {
    __shared__ int sm4[4];
    const int64_t n_alive = state[kStN]; const int S = state[kStS];
    for (int64_t base = (int64_t)blockIdx.x * 128; base < n_alive; base += (int64_t)gridDim.x * 128) {
        const int64_t n = base + threadIdx.x;
        int s = 0;
        if (n < n_alive) {
            const int64_t r = alive[n];
            const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
            float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
            float t_resume = t; bool moved = false;
            while (t < t2 && s < S) {
                float x, y, z, dt;
                if (arn_march_eval_t<FAST>(c, ray, bitfield, t, x, y, z, dt)) {
                    const int64_t o = n * S + s;
                    ts[o] = t; deltas[o] = dt;
                    t = __fadd_rn(t, dt);
                    t_resume = t; moved = true;  // raymarching.cu:386: only an occupied step moves the resume point
                    s++;
                }
            }
            if (moved) hits_t[2 * r] = t_resume;
            n_eff[n] = s;
        }
        const int tot = block_sum_128(s, sm4);
        if (threadIdx.x == 0) partial[base >> 7] = tot;
    }
}

// The same march with one WARP per alive ray (arn_march_core.h, "Window form": 32 chain points probed per turn, visited
// occupied points compacted with ballot/popc) for the iterations in which few rays take many samples each -- one thread
// walking 64 samples is a serial chain of probes (~1 us per sample), a warp takes them a window at a time.  Same samples,
// same resume point; writes no chunk sums (the scan re-reads the few counts instead).
template <bool CONST_DT, bool FAST>
__global__ void __launch_bounds__(256) march_test_warp_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                              float* __restrict__ hits_t, const int64_t* __restrict__ alive,
                                                              const int32_t* __restrict__ state, const uint8_t* __restrict__ bitfield,
                                                              ArnMarchConsts c, float* __restrict__ deltas, float* __restrict__ ts,
                                                              int32_t* __restrict__ n_eff) {
    const int64_t n_alive = state[kStN]; const int S = state[kStS];
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_alive; n += n_warps) {
        const int64_t r = alive[n];
        const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
        const float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        int N = 0;
        float t_resume = t1; bool moved = false;
        if (t1 < t2) {
            float t = t1;
#pragma unroll
            for (int j = 0; j < 31; j++) {
                const float tn = __fadd_rn(t, CONST_DT ? c.dt_lo : arn_calc_dt(c, t));
                if (lane > j) t = tn;
            }
            float pending = -INFINITY;
            for (;;) {
                float x, y, z, dt, tgt;
                const bool occ = arn_march_probe<FAST, FAST>(c, ray, bitfield, t, x, y, z, dt, tgt);
                const unsigned valid = __ballot_sync(kFull, t < t2);
                const unsigned occm = __ballot_sync(kFull, occ);
                const int s0 = __popc(__ballot_sync(kFull, t < pending));
                int lo = lane + 1, hi = 32;
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const int mid = (lo + hi) >> 1;
                    const float tv = __shfl_sync(kFull, t, mid & 31);
                    if (lo < hi) { if (tv < tgt) lo = mid + 1; else hi = mid; }
                }
                int R = occ ? lane + 1 : lo;
                unsigned M = 1u << lane;
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const unsigned Mo = __shfl_sync(kFull, M, R & 31);
                    const int Ro = __shfl_sync(kFull, R, R & 31);
                    if (R < 32) { M |= Mo; R = Ro; }
                }
                const unsigned vis = s0 < 32 ? __shfl_sync(kFull, M, s0 & 31) : 0u;
                unsigned emit = vis & occm & valid;
                const int rem = S - N;
                bool done = valid != kFull;
                if (__popc(emit) >= rem) {  // the iteration's sample budget ends the march inside this window
                    done = true;
                    const int rank_all = __popc(emit & ((1u << lane) - 1u));
                    emit = __ballot_sync(kFull, ((emit >> lane) & 1u) && rank_all < rem);
                }
                if ((emit >> lane) & 1u) {
                    const int64_t o = n * S + N + __popc(emit & ((1u << lane) - 1u));
                    ts[o] = t; deltas[o] = dt;
                }
                if (emit) {  // raymarching.cu:386: the resume point follows the last sample taken
                    const int last = 31 - __clz(emit);
                    t_resume = __shfl_sync(kFull, __fadd_rn(t, dt), last);
                    moved = true;
                }
                N += __popc(emit);
                if (done) break;
                if (vis) {
                    const int last = 31 - __clz(vis);
                    pending = __shfl_sync(kFull, occ ? -INFINITY : tgt, last);
                }
#pragma unroll
                for (int k = 0; k < 32; k++) t = __fadd_rn(t, CONST_DT ? c.dt_lo : arn_calc_dt(c, t));
            }
        }
        if (lane == 0) {
            if (moved) hits_t[2 * r] = t_resume;
            n_eff[n] = N;
        }
    }
}

// ---- The frame's samples marched ONCE (arn_march_test_all): the test march is resumable and deterministic -- an iteration
// starts at the chain point behind the previous iteration's last sample -- so the sequence of samples of a ray does not
// depend on how the loop slices it.  One thread per ray marches the whole ray and records the parameter t of every
// occupied sample, sample-major (ts_all[s * n_rays + r]: neighbouring rays write neighbouring words), up to `stride`
// samples (the loop never asks a ray for more than max_samples + 63); an iteration of the loop then only slices the
// next N_samples of each alive ray (neff / emit kernels below) instead of marching -- no iteration waits for a thread
// that crosses an empty stretch, and no empty cell is probed twice.  dt of a sample is calc_dt(t), as the march computes it.
template <bool FAST>
__global__ void __launch_bounds__(128) march_test_all_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                             const float* __restrict__ hits_t, int64_t n_rays,
                                                             const uint8_t* __restrict__ bitfield, ArnMarchConsts c, int stride,
                                                             float* __restrict__ ts_all, int32_t* __restrict__ totals, int32_t* __restrict__ cursor) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const ArnRay ray = arn_load_ray(rays_o + 3 * r, rays_d + 3 * r);
    float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
    int s = 0;
    while (t < t2 && s < stride) {
        float x, y, z, dt;
        if (arn_march_eval_t<FAST>(c, ray, bitfield, t, x, y, z, dt)) {
            ts_all[(int64_t)s * n_rays + r] = t;
            t = __fadd_rn(t, dt);
            s++;
        }
    }
    totals[r] = s; cursor[r] = 0;
}

// N_eff of the iteration for every alive ray (what the march would have returned) + chunk sums for the scan
__global__ void __launch_bounds__(128) neff_test_pre_kernel(const int64_t* __restrict__ alive, const int32_t* __restrict__ state,
                                                            const int32_t* __restrict__ totals, const int32_t* __restrict__ cursor,
                                                            int32_t* __restrict__ n_eff, int32_t* __restrict__ partial) {
    pdl_enter();
    __shared__ int sm4[4];
    const int64_t n_alive = state[kStN]; const int S = state[kStS];
    for (int64_t base = (int64_t)blockIdx.x * 128; base < n_alive; base += (int64_t)gridDim.x * 128) {
        const int64_t n = base + threadIdx.x;
        int s = 0;
        if (n < n_alive) {
            const int64_t r = alive[n];
            s = min(S, totals[r] - cursor[r]);
            n_eff[n] = s;
        }
        const int tot = block_sum_128(s, sm4);
        if (threadIdx.x == 0) partial[base >> 7] = tot;
    }
}

// the iteration's samples sliced out of ts_all: padded (deltas, ts) for the compositing kernel, compact (xyzs, dirs) for the field
__global__ void __launch_bounds__(256) emit_test_pre_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                            const int64_t* __restrict__ alive, const int32_t* __restrict__ state,
                                                            const int64_t* __restrict__ rays_a, const float* __restrict__ ts_all,
                                                            const int32_t* __restrict__ cursor, int64_t n_rays, ArnMarchConsts c,
                                                            float* __restrict__ deltas, float* __restrict__ ts,
                                                            float* __restrict__ xyzs, float* __restrict__ dirs) {
    pdl_enter();
    const int64_t n_alive = state[kStN]; const int S = state[kStS];
    const int64_t total = n_alive * S;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = e / S; const int sl = (int)(e - n * S);
        if (sl >= (int)rays_a[3 * n + 2]) continue;
        const int64_t r = alive[n], o = rays_a[3 * n + 1] + sl;
        const float t = ts_all[(int64_t)(cursor[r] + sl) * n_rays + r];
        ts[e] = t; deltas[e] = arn_calc_dt(c, t);
        const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
        xyzs[3 * o] = __fmaf_rn(dx, t, rays_o[3 * r]); xyzs[3 * o + 1] = __fmaf_rn(dy, t, rays_o[3 * r + 1]); xyzs[3 * o + 2] = __fmaf_rn(dz, t, rays_o[3 * r + 2]);
        dirs[3 * o] = dx; dirs[3 * o + 1] = dy; dirs[3 * o + 2] = dz;
    }
}

// Offsets of the compact sample list: rays_a[n] = (n, start, N) for the n_alive rays, counts = (valid samples, n_alive).
__global__ void __launch_bounds__(1024) scan_test_dyn_kernel(const int32_t* __restrict__ counts_in, const int32_t* __restrict__ partial,
                                                             const int32_t* __restrict__ state, int64_t* __restrict__ rays_a,
                                                             int32_t* __restrict__ counter) {
    pdl_enter();
    __shared__ int64_t warp_inc[32];
    __shared__ int64_t s_pre;
    const int64_t n_rays = state[kStN];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t r_base = (int64_t)blockIdx.x * 1024; r_base < n_rays || (r_base == 0 && n_rays == 0); r_base += (int64_t)gridDim.x * 1024) {
        const int64_t r = r_base + threadIdx.x;
        int64_t pre = 0;
        if (partial) { for (int64_t j = threadIdx.x; j < (r_base >> 7); j += 1024) pre += partial[j]; }
        else { for (int64_t j = threadIdx.x; j < r_base; j += 1024) pre += counts_in[j]; }
        const int32_t mine = r < n_rays ? counts_in[r] : 0;
        int64_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += u; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(kFull, pre, o);
        __syncthreads();  // the previous turn's readers of warp_inc / s_pre are done
        if (lane == 31) warp_inc[wid] = inc;
        if (threadIdx.x == 0) s_pre = 0;
        __syncthreads();
        if (lane == 0 && pre) atomicAdd((unsigned long long*)&s_pre, (unsigned long long)pre);
        if (wid == 0) {
            int64_t v = warp_inc[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, v, o); if (lane >= o) v += u; }
            warp_inc[lane] = v;  // inclusive over warps
        }
        __syncthreads();
        const int64_t before = s_pre;
        const int64_t start = before + (wid ? warp_inc[wid - 1] : 0) + inc - mine;
        if (r < n_rays) { rays_a[3 * r] = r; rays_a[3 * r + 1] = start; rays_a[3 * r + 2] = mine; }
        if (threadIdx.x == 0 && n_rays <= r_base + 1024) {  // the CTA that holds the last ray (or the only CTA of an empty list)
            const int64_t tot = before + warp_inc[31];
            counter[0] = (int32_t)(tot > 0x7fffffff ? 0x7fffffff : tot);
            counter[1] = (int32_t)n_rays;
        }
    }
}

__global__ void __launch_bounds__(256) emit_test_dyn_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                            const int64_t* __restrict__ alive, const int32_t* __restrict__ state,
                                                            const int64_t* __restrict__ rays_a, const float* __restrict__ ts,
                                                            float* __restrict__ xyzs, float* __restrict__ dirs) {
    const int64_t n_alive = state[kStN]; const int S = state[kStS];
    const int64_t total = n_alive * S;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = e / S; const int sl = (int)(e - n * S);
        if (sl >= (int)rays_a[3 * n + 2]) continue;
        const int64_t r = alive[n], o = rays_a[3 * n + 1] + sl;
        const float t = ts[e];
        const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
        xyzs[3 * o] = __fmaf_rn(dx, t, rays_o[3 * r]); xyzs[3 * o + 1] = __fmaf_rn(dy, t, rays_o[3 * r + 1]); xyzs[3 * o + 2] = __fmaf_rn(dz, t, rays_o[3 * r + 2]);
        dirs[3 * o] = dx; dirs[3 * o + 1] = dy; dirs[3 * o + 2] = dz;
    }
}

// composite_test_compact_kernel over the device-side count; keep flags + their sums per 128 rays
__global__ void __launch_bounds__(128) composite_test_dyn_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                 const float* __restrict__ deltas, const float* __restrict__ ts,
                                                                 const int64_t* __restrict__ alive, const int32_t* __restrict__ state, float T_thr,
                                                                 const int64_t* __restrict__ rays_a, float* __restrict__ opacity,
                                                                 float* __restrict__ depth, float* __restrict__ rgb, int32_t* __restrict__ keep,
                                                                 int32_t* __restrict__ partial, unsigned long long* __restrict__ total,
                                                                 int32_t* __restrict__ cursor) {
    pdl_enter();
    __shared__ int sm4[4];
    const int64_t n_alive = state[kStN]; const int S = state[kStS];
    for (int64_t base = (int64_t)blockIdx.x * 128; base < n_alive; base += (int64_t)gridDim.x * 128) {
        const int64_t n = base + threadIdx.x;
        int ne = 0; bool live = false;
        if (n < n_alive) {
            ne = (int)rays_a[3 * n + 2];
            if (ne > 0) {
                live = true;
                const int64_t r = alive[n], c0 = rays_a[3 * n + 1];
                if (cursor) cursor[r] += ne;  // pre-marched frame: the ray's next slice starts behind these samples
                float O = opacity[r];
                float T = __fsub_rn(1.0f, O);
                float cr = rgb[3 * r], cg = rgb[3 * r + 1], cb = rgb[3 * r + 2], D = depth[r];
                for (int s = 0; s < ne; s++) {
                    const int64_t o = n * S + s, q = c0 + s;
                    const float a = alpha_of(sigmas[q], deltas[o]);
                    const float w = __fmul_rn(a, T);
                    cr = __fmaf_rn(w, rgbs[3 * q], cr); cg = __fmaf_rn(w, rgbs[3 * q + 1], cg); cb = __fmaf_rn(w, rgbs[3 * q + 2], cb);
                    D = __fmaf_rn(w, ts[o], D);
                    O = __fadd_rn(O, w);
                    T = __fmul_rn(T, __fsub_rn(1.0f, a));
                    if (T <= T_thr) { live = false; break; }
                }
                opacity[r] = O; depth[r] = D; rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
            }
            keep[n] = live ? 1 : 0;
        }
        unsigned v = (unsigned)ne;  // effective samples of this iteration: warp sum, one atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, (unsigned long long)v);
        const int kept = block_sum_128(live ? 1 : 0, sm4);
        if (threadIdx.x == 0) partial[base >> 7] = kept;
    }
}

// alive_out = alive[keep != 0] (order preserved: alive_indices[alive_indices >= 0]) and the control state of the NEXT
// iteration, the schedule of rendering.py:184-206: stop when nothing was marched, nobody is alive or the sample budget is
// spent; otherwise N_samples = max(min(N_rays // N_alive, 64), min_samples).
__global__ void __launch_bounds__(1024) alive_compact_dyn_kernel(const int64_t* __restrict__ alive, const int32_t* __restrict__ keep,
                                                                 const int32_t* __restrict__ partial, const int32_t* __restrict__ state_in,
                                                                 const int32_t* __restrict__ counts, int64_t* __restrict__ alive_out,
                                                                 int32_t* __restrict__ counts_alive, int32_t* __restrict__ state_out,
                                                                 int64_t n_rays_total, int min_samples, int budget) {
    pdl_enter();
    __shared__ int64_t warp_inc[32];
    __shared__ int64_t s_pre;
    const int64_t n = state_in[kStN];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t r_base = (int64_t)blockIdx.x * 1024; r_base < n || (r_base == 0 && n == 0); r_base += (int64_t)gridDim.x * 1024) {
        const int64_t r = r_base + threadIdx.x;
        int64_t pre = 0;
        for (int64_t j = threadIdx.x; j < (r_base >> 7); j += 1024) pre += partial[j];
        const int32_t mine = r < n ? keep[r] : 0;
        int64_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += u; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(kFull, pre, o);
        __syncthreads();
        if (lane == 31) warp_inc[wid] = inc;
        if (threadIdx.x == 0) s_pre = 0;
        __syncthreads();
        if (lane == 0 && pre) atomicAdd((unsigned long long*)&s_pre, (unsigned long long)pre);
        if (wid == 0) {
            int64_t v = warp_inc[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int64_t u = __shfl_up_sync(kFull, v, o); if (lane >= o) v += u; }
            warp_inc[lane] = v;
        }
        __syncthreads();
        const int64_t before = s_pre;
        if (mine) alive_out[before + (wid ? warp_inc[wid - 1] : 0) + inc - mine] = alive[r];
        if (threadIdx.x == 0 && n <= r_base + 1024) {
            const int64_t n_keep = before + warp_inc[31];
            counts_alive[0] = (int32_t)n_keep; counts_alive[1] = (int32_t)n;
            int active = state_in[kStActive], done = state_in[kStDone], S = 1;
            int64_t n_next = 0;
            if (active && counts[0] == 0) active = 0;                       // rendering.py:206: nothing was marched
            if (active) { n_next = n_keep; if (n_next == 0 || done >= budget) active = 0; }
            if (active) {
                const int64_t q = n_rays_total / n_next;
                S = (int)(q < 64 ? q : 64); if (S < min_samples) S = min_samples;
                done += S;
            } else n_next = 0;
            state_out[kStN] = (int32_t)n_next; state_out[kStS] = S; state_out[kStDone] = done; state_out[kStActive] = active;
            state_out[kStIters] = state_in[kStIters] + 1;
            state_out[kStLive] = state_in[kStLive] + (state_in[kStActive] ? 1 : 0);
        }
    }
}

// ---- The pre-marched iteration in FOUR launches (arn_render_test_step_fused): a small frame -- one rank's share of a frame
// rendered by several GPUs -- is bound by the NUMBER of kernels of its ~50 iterations, not by their work.  The two scans that
// cost the seven-launch form three kernels are replaced by atomics whose ORDER does not matter:
//   slice_emit      N_eff of every alive ray, its place in the compact sample list (block scan + ONE atomicAdd per block: the list
//                   is ordered inside a block of 128 rays, blocks land in arrival order), and the samples themselves
//   (hash grid, MLP on the compact list; the sample count is read on the device)
//   composite_keep  compositing + ray kill as before, survivors appended to the next alive list by warp-aggregated atomicAdd,
//                   the LAST block to finish (ticket) writes the loop's next control state and resets the counters.
// A ray's samples, its compositing and the counts do not depend on where the ray sits in either list: pixels, kill pattern and
// total_samples are those of the seven-launch form (tests hold them identical); only the lists' order is scheduling-dependent.
// sync (4 x int32, zero before the first iteration): {samples of this iteration, -, survivors, ticket}.
__global__ void __launch_bounds__(128) test_slice_emit_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                              const int64_t* __restrict__ alive, const int32_t* __restrict__ state,
                                                              const float* __restrict__ ts_all, const int32_t* __restrict__ totals,
                                                              const int32_t* __restrict__ cursor, int64_t n_rays, ArnMarchConsts c,
                                                              int32_t* __restrict__ n_eff, int64_t* __restrict__ rays_a, int32_t* __restrict__ sync,
                                                              float* __restrict__ deltas, float* __restrict__ ts, float* __restrict__ xyzs,
                                                              float* __restrict__ dirs, int64_t capacity) {
    __shared__ int s_excl[129];
    __shared__ int s_warp[4];
    __shared__ int s_base;
    const int64_t n_alive = state[kStN]; const int S = state[kStS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int64_t base = (int64_t)blockIdx.x * 128; base < n_alive; base += (int64_t)gridDim.x * 128) {
        const int64_t n = base + tid;
        int s = 0;
        if (n < n_alive) { const int64_t r = alive[n]; s = min(S, totals[r] - cursor[r]); n_eff[n] = s; }
        int inc = s;  // inclusive scan over the block
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        const int before = (wid > 0 ? s_warp[0] : 0) + (wid > 1 ? s_warp[1] : 0) + (wid > 2 ? s_warp[2] : 0);
        const int tot = s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
        s_excl[tid] = before + inc - s;
        if (tid == 0) { s_excl[128] = tot; s_base = tot ? atomicAdd(&sync[0], tot) : 0; }
        __syncthreads();
        const int64_t blk0 = s_base;
        if (n < n_alive) { rays_a[3 * n] = n; rays_a[3 * n + 1] = blk0 + s_excl[tid]; rays_a[3 * n + 2] = s; }
        // the block's samples, one per thread and turn: which ray of the block sample e belongs to = binary search of the offsets
        for (int e = tid; e < tot; e += 128) {
            int lo = 0, hi = 127;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_excl[mid] <= e) lo = mid; else hi = mid - 1; }
            // rays with no samples share their successor's offset: the LAST ray whose offset is <= e owns it (lo), and it has samples
            const int sl = e - s_excl[lo];
            const int64_t nj = base + lo, r = alive[nj], o = blk0 + e;
            const float t = ts_all[(int64_t)(cursor[r] + sl) * n_rays + r];
            ts[nj * S + sl] = t; deltas[nj * S + sl] = arn_calc_dt(c, t);
            if (o < capacity) {
                const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
                xyzs[3 * o] = __fmaf_rn(dx, t, rays_o[3 * r]); xyzs[3 * o + 1] = __fmaf_rn(dy, t, rays_o[3 * r + 1]); xyzs[3 * o + 2] = __fmaf_rn(dz, t, rays_o[3 * r + 2]);
                dirs[3 * o] = dx; dirs[3 * o + 1] = dy; dirs[3 * o + 2] = dz;
            }
        }
        __syncthreads();  // s_excl / s_base are rewritten by the next turn
    }
}

__global__ void __launch_bounds__(128) test_composite_keep_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                  const float* __restrict__ deltas, const float* __restrict__ ts,
                                                                  const int64_t* __restrict__ alive, const int32_t* __restrict__ state_in, float T_thr,
                                                                  const int64_t* __restrict__ rays_a, float* __restrict__ opacity,
                                                                  float* __restrict__ depth, float* __restrict__ rgb,
                                                                  unsigned long long* __restrict__ total, int32_t* __restrict__ cursor,
                                                                  int64_t* __restrict__ alive_out, int32_t* __restrict__ sync,
                                                                  int32_t* __restrict__ counts_alive, int32_t* __restrict__ state_out,
                                                                  int64_t n_rays_total, int min_samples, int budget) {
    __shared__ int s_last;
    const int64_t n_alive = state_in[kStN]; const int S = state_in[kStS];
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * 128; base < n_alive; base += (int64_t)gridDim.x * 128) {
        const int64_t n = base + threadIdx.x;
        int ne = 0; bool live = false; int64_t r = 0;
        if (n < n_alive) {
            ne = (int)rays_a[3 * n + 2];
            r = alive[n];
            if (ne > 0) {
                live = true;
                const int64_t c0 = rays_a[3 * n + 1];
                cursor[r] += ne;  // the ray's next slice starts behind these samples
                float O = opacity[r];
                float T = __fsub_rn(1.0f, O);
                float cr = rgb[3 * r], cg = rgb[3 * r + 1], cb = rgb[3 * r + 2], D = depth[r];
                for (int s = 0; s < ne; s++) {
                    const int64_t o = n * S + s, q = c0 + s;
                    const float a = alpha_of(sigmas[q], deltas[o]);
                    const float w = __fmul_rn(a, T);
                    cr = __fmaf_rn(w, rgbs[3 * q], cr); cg = __fmaf_rn(w, rgbs[3 * q + 1], cg); cb = __fmaf_rn(w, rgbs[3 * q + 2], cb);
                    D = __fmaf_rn(w, ts[o], D);
                    O = __fadd_rn(O, w);
                    T = __fmul_rn(T, __fsub_rn(1.0f, a));
                    if (T <= T_thr) { live = false; break; }
                }
                opacity[r] = O; depth[r] = D; rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
            }
        }
        unsigned v = (unsigned)ne;  // effective samples of this iteration: warp sum, one atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0 && v) atomicAdd(total, (unsigned long long)v);
        // survivors: one atomicAdd per warp, the warp's survivors stay in their order
        const unsigned keep = __ballot_sync(kFull, live);
        int pos = 0;
        if (lane == 0 && keep) pos = atomicAdd(&sync[2], __popc(keep));
        pos = __shfl_sync(kFull, pos, 0);
        if (live) alive_out[pos + __popc(keep & ((1u << lane) - 1u))] = r;
    }
    // the last block to get here writes the loop's next control state (the schedule of rendering.py:184-206) and resets the counters
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd(&sync[3], 1) == (int)gridDim.x - 1; }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const int64_t n_keep = atomicAdd(&sync[2], 0);
        const int marched = atomicAdd(&sync[0], 0);
        counts_alive[0] = (int32_t)n_keep; counts_alive[1] = (int32_t)n_alive;
        int active = state_in[kStActive], done = state_in[kStDone], Sn = 1;
        int64_t n_next = 0;
        if (active && marched == 0) active = 0;                          // rendering.py:206: nothing was marched
        if (active) { n_next = n_keep; if (n_next == 0 || done >= budget) active = 0; }
        if (active) {
            const int64_t q = n_rays_total / n_next;
            Sn = (int)(q < 64 ? q : 64); if (Sn < min_samples) Sn = min_samples;
            done += Sn;
        } else n_next = 0;
        state_out[kStN] = (int32_t)n_next; state_out[kStS] = Sn; state_out[kStDone] = done; state_out[kStActive] = active;
        state_out[kStIters] = state_in[kStIters] + 1;
        state_out[kStLive] = state_in[kStLive] + (state_in[kStActive] ? 1 : 0);
        sync[0] = 0; sync[2] = 0; sync[3] = 0;
        __threadfence();
    }
}

// ---------------------------------------------------------------------------------------------- distortion loss
// losses.cu:7-59,62-107: one warp per row; inclusive scans kept for the backward, loss reduced in the same pass.
__global__ void __launch_bounds__(256) distortion_fw_kernel(const float* __restrict__ ws, const float* __restrict__ deltas,
                                                            const float* __restrict__ ts, const int64_t* __restrict__ rays_a,
                                                            int64_t n_rays, float* __restrict__ loss, float* __restrict__ wsi,
                                                            float* __restrict__ wtsi) {
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= n_rays) return;
    const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    float cw = 0.f, cwt = 0.f, acc = 0.f;
    for (int base = 0; base < N; base += 32) {
        const int i = base + lane; const bool in = i < N; const int64_t s = start + i;
        const float w = in ? ws[s] : 0.f, t = in ? ts[s] : 0.f, dl = in ? deltas[s] : 0.f;
        const float wt = w * t;
        const float wi = cw + warp_incl_sum(w, lane), wti = cwt + warp_incl_sum(wt, lane);
        const float we = wi - w, wte = wti - wt;
        if (in) { wsi[s] = wi; wtsi[s] = wti; }
        acc += in ? (2.0f * (wti * we - wi * wte) + (1.0f / 3.0f) * w * w * dl) : 0.f;
        cw = __shfl_sync(kFull, wi, 31); cwt = __shfl_sync(kFull, wti, 31);
    }
    acc = warp_sum(acc);
    if (lane == 0) loss[ray_idx] = acc;
}

// losses.cu:110-140 -- every sample is independent given the scans: one thread per sample of a row's segment.
__global__ void __launch_bounds__(256) distortion_bw_kernel(const float* __restrict__ dL_dloss, const float* __restrict__ wsi,
                                                            const float* __restrict__ wtsi, const float* __restrict__ ws,
                                                            const float* __restrict__ deltas, const float* __restrict__ ts,
                                                            const int64_t* __restrict__ rays_a, int64_t n_rays, float* __restrict__ dL_dws) {
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= n_rays) return;
    const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    if (N <= 0) return;
    const int64_t end = start + N - 1;
    const float ws_sum = wsi[end], wts_sum = wtsi[end], gl = dL_dloss[ray_idx];
    for (int64_t s = start + lane; s <= end; s += 32) {
        float v = gl * 2.0f * ((s == start ? 0.0f : (ts[s] * wsi[s - 1] - wtsi[s - 1])) +
                               (wts_sum - wtsi[s] - ts[s] * (ws_sum - wsi[s])));
        v += gl * (2.0f / 3.0f) * ws[s] * deltas[s];
        dL_dws[s] = v;
    }
}

// custom_functions.py:104-112 (torch_scatter.segment_csr): one warp per row.
__global__ void __launch_bounds__(256) march_train_bw_kernel(const float* __restrict__ dL_dxyzs, const float* __restrict__ dL_ddirs,
                                                             const float* __restrict__ ts, const int64_t* __restrict__ rays_a,
                                                             int64_t n_rays, float* __restrict__ dL_do, float* __restrict__ dL_dd) {
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= n_rays) return;
    const int64_t start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    float o[3] = {0, 0, 0}, d[3] = {0, 0, 0};
    for (int i = lane; i < N; i += 32) {
        const int64_t s = start + i; const float t = ts[s];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float gx = dL_dxyzs[3 * s + k];
            o[k] += gx; d[k] += gx * t + (dL_ddirs ? dL_ddirs[3 * s + k] : 0.0f);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { o[k] = warp_sum(o[k]); d[k] = warp_sum(d[k]); }
    if (lane == 0) {
        for (int k = 0; k < 3; k++) { dL_do[3 * n + k] = o[k]; dL_dd[3 * n + k] = d[k]; }  // segment_csr is by row
    }
}

}  // namespace arn

using namespace arn;

// ================================================================================================ C ABI
extern "C" ARN_API int arn_ray_aabb_intersect(const float* rays_o, const float* rays_d, int64_t n_rays, const float* centers,
                                      const float* half_sizes, int n_voxels, int max_hits, int32_t* hit_cnt, float* hits_t,
                                      int64_t* hits_idx, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && n_voxels >= 0 && max_hits >= 1, "bad sizes");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && centers && half_sizes && hit_cnt && hits_t && hits_idx, "null pointer");
    ARN_LAUNCH("intersect_kernel", (cudaStream_t)stream, intersect_kernel<false><<<ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, n_rays, centers, half_sizes,
                                                                                    n_voxels, max_hits, hit_cnt, hits_t, hits_idx));
    return check_launch("ray_aabb_intersect");
}

extern "C" ARN_API int arn_ray_sphere_intersect(const float* rays_o, const float* rays_d, int64_t n_rays, const float* centers,
                                        const float* radii, int n_spheres, int max_hits, int32_t* hit_cnt, float* hits_t,
                                        int64_t* hits_idx, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && n_spheres >= 0 && max_hits >= 1, "bad sizes");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && centers && radii && hit_cnt && hits_t && hits_idx, "null pointer");
    ARN_LAUNCH("intersect_kernel", (cudaStream_t)stream, intersect_kernel<true><<<ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, n_rays, centers, radii,
                                                                                   n_spheres, max_hits, hit_cnt, hits_t, hits_idx));
    return check_launch("ray_sphere_intersect");
}

extern "C" ARN_API int arn_ray_aabb_near(const float* rays_o, const float* rays_d, int64_t n_rays, const float* center_host,
                                 const float* half_size_host, float near, float* hits_t, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0, "bad sizes");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && center_host && half_size_host && hits_t, "null pointer");
    float ch[6];  // the box is 24 bytes of module state: passed from the host so it travels as kernel arguments
    for (int k = 0; k < 3; k++) { ch[k] = center_host[k]; ch[3 + k] = half_size_host[k]; }
    ARN_LAUNCH_PDL("aabb_near_kernel", (cudaStream_t)stream, (aabb_near_kernel), ceil_div(n_rays, 256), 256, 0, rays_o, rays_d, n_rays, ch[0], ch[1], ch[2], ch[3],
                                                                            ch[4], ch[5], near, reinterpret_cast<float2*>(hits_t));
    return check_launch("ray_aabb_near");
}

extern "C" ARN_API int arn_morton3d(const int32_t* coords, int64_t n, int32_t* indices, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(coords && indices, "null pointer");
    ARN_LAUNCH("morton3d_kernel", (cudaStream_t)stream, morton3d_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(coords, n, indices));
    return check_launch("morton3d");
}
extern "C" ARN_API int arn_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(coords && indices, "null pointer");
    ARN_LAUNCH("morton3d_invert_kernel", (cudaStream_t)stream, morton3d_invert_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(indices, n, coords));
    return check_launch("morton3d_invert");
}
extern "C" ARN_API int arn_packbits(const void* density_grid, int grid_dtype, float threshold, uint8_t* density_bitfield,
                            int64_t n_bytes, arn_stream_t stream) {
    ARN_REQUIRE(n_bytes >= 0, "bad size");
    if (n_bytes == 0) return ARN_OK;
    ARN_REQUIRE(density_grid && density_bitfield, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t words = (n_bytes + 3) / 4;
    if (grid_dtype == 0) ARN_LAUNCH("packbits_kernel", st, packbits_kernel<float><<<ceil_div(words, 256), 256, 0, st>>>((const float*)density_grid, threshold, density_bitfield, n_bytes));
    else if (grid_dtype == 1) ARN_LAUNCH("packbits_kernel", st, packbits_kernel<__half><<<ceil_div(words, 256), 256, 0, st>>>((const __half*)density_grid, threshold, density_bitfield, n_bytes));
    else if (grid_dtype == 2) ARN_LAUNCH("packbits_f64_kernel", st, packbits_f64_kernel<<<ceil_div(n_bytes, 256), 256, 0, st>>>((const double*)density_grid, threshold, density_bitfield, n_bytes));
    else { set_error("arn_packbits: grid_dtype must be 0 (f32), 1 (f16) or 2 (f64)"); return ARN_E_INVALID; }
    return check_launch("packbits");
}

extern "C" ARN_API int arn_gather_rays(const float* directions, const float* K_host, int width, const float* poses, const int64_t* img_idxs,
                                       int64_t img_single, const int64_t* pix_idxs, int64_t n, float* rays_o, float* rays_d, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(poses && pix_idxs && rays_o && rays_d, "null pointer");
    ARN_REQUIRE(directions || (K_host && width > 0), "either the direction table or the intrinsics + image width are needed");
    float fx = 1.f, fy = 1.f, cx = 0.f, cy = 0.f;
    if (!directions) { fx = K_host[0]; fy = K_host[4]; cx = K_host[2]; cy = K_host[5]; }
    ARN_LAUNCH("gather_rays_kernel", (cudaStream_t)stream, gather_rays_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        directions, poses, img_idxs, img_single, pix_idxs, n, width, fx, fy, cx, cy, rays_o, rays_d, nullptr, 0, 0, nullptr));
    return check_launch("gather_rays");
}

extern "C" ARN_API int arn_gather_batch(const float* directions, const float* K_host, int width, const float* poses, const int64_t* img_idxs,
                                        int64_t img_single, const int64_t* pix_idxs, int64_t n, const float* images, int64_t pixels_per_image,
                                        int channels, float* rays_o, float* rays_d, float* pixels_out, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(poses && pix_idxs && rays_o && rays_d && images && pixels_out, "null pointer");
    ARN_REQUIRE(directions || (K_host && width > 0), "either directions or K_host + width");
    ARN_REQUIRE(pixels_per_image > 0 && channels >= 1 && channels <= 8, "bad image layout");
    float fx = 1.f, fy = 1.f, cx = 0.f, cy = 0.f;
    if (!directions) { fx = K_host[0]; fy = K_host[4]; cx = K_host[2]; cy = K_host[5]; }
    ARN_LAUNCH("gather_rays_kernel", (cudaStream_t)stream, gather_rays_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        directions, poses, img_idxs, img_single, pix_idxs, n, width, fx, fy, cx, cy, rays_o, rays_d, images, pixels_per_image, channels, pixels_out));
    return check_launch("gather_batch");
}

extern "C" ARN_API int arn_grid_cell_positions(const int32_t* coords, const float* rnd, int64_t n_cells, int grid_size, float s, float* xyzs,
                                               arn_stream_t stream) {
    ARN_REQUIRE(n_cells >= 0 && grid_size >= 2, "bad size");
    if (n_cells == 0) return ARN_OK;
    ARN_REQUIRE(coords && rnd && xyzs, "null pointer");
    const float half = s / (float)grid_size;  // Python: half_grid_size = s / G (double), s - half (double), both then cast by torch to float32
    const float s_minus_half = (float)((double)s - (double)s / (double)grid_size);
    const float half_f = (float)((double)s / (double)grid_size);
    (void)half;
    ARN_LAUNCH("cell_positions_kernel", (cudaStream_t)stream, cell_positions_kernel<<<ceil_div(3 * n_cells, 256), 256, 0, (cudaStream_t)stream>>>(
        coords, rnd, 3 * n_cells, 1.0f / (float)(grid_size - 1), s_minus_half, half_f, xyzs));
    return check_launch("grid_cell_positions");
}

extern "C" ARN_API int arn_mark_invisible_cells(const int32_t* coords, const int64_t* indices, int64_t n_cells, int grid_size, float s,
                                                const float* w2c, int n_cams, const float* K_host, float img_w, float img_h, float near,
                                                float* density_grid, float* count_grid, arn_stream_t stream) {
    ARN_REQUIRE(n_cells >= 0 && grid_size >= 2 && n_cams >= 1, "bad size");
    if (n_cells == 0) return ARN_OK;
    ARN_REQUIRE(coords && indices && w2c && K_host && density_grid && count_grid, "null pointer");
    const float s_minus_half = (float)((double)s - (double)s / (double)grid_size);  // Python double arithmetic, cast by torch to float32
    ARN_LAUNCH("mark_invisible_kernel", (cudaStream_t)stream, mark_invisible_kernel<<<ceil_div(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(
        coords, indices, n_cells, 1.0f / (float)(grid_size - 1), s_minus_half, w2c, n_cams, K_host[0], K_host[1], K_host[2], K_host[3], K_host[4],
        K_host[5], K_host[6], K_host[7], K_host[8], img_w, img_h, near, density_grid, count_grid));
    return check_launch("mark_invisible_cells");
}

static int grid_sample_cells_impl(const float* density_grid, float density_threshold, int grid_size, float s, const int32_t* coords1,
                                  const int64_t* u, int64_t M, const float* rnd, void* scratch, int64_t* indices, float* xyzs, bool sorted,
                                  arn_stream_t stream) {
    ARN_REQUIRE(grid_size >= 2 && grid_size <= 1024 && M > 0 && M % 256 == 0, "bad sizes (M must be a positive multiple of 256)");
    ARN_REQUIRE(density_grid && coords1 && u && rnd && scratch && indices && xyzs, "null pointer");
    const int64_t n_cells = (int64_t)grid_size * grid_size * grid_size;
    const int n_chunks = (int)((n_cells + kCellChunk - 1) / kCellChunk);
    ARN_REQUIRE(n_chunks <= kMaxChunks - 1, "grid too large for the single-CTA chunk scan (grid_size <= 160)");
    ARN_REQUIRE(((uintptr_t)scratch & 15) == 0, "scratch must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // scratch: masks (32 per chunk) | chunk_count (n_chunks) | chunk_prefix (n_chunks + 1) | pad | sorted only: chunk_tot (n_chunks) |
    // chunk_base (n_chunks + 1) | hist (n_cells) | idx_tmp (2 M)
    uint32_t* masks = (uint32_t*)scratch;
    int32_t* chunk_count = (int32_t*)(masks + (size_t)n_chunks * 32);
    int32_t* chunk_prefix = chunk_count + n_chunks;
    int32_t* chunk_tot = (int32_t*)scratch + (size_t)n_chunks * 34 + 4;
    int32_t* chunk_base = chunk_tot + n_chunks;
    int32_t* hist = chunk_base + n_chunks + 4;
    uint32_t* idx_tmp = (uint32_t*)(hist + n_cells);
    ARN_LAUNCH("occ_mask_kernel", st, occ_mask_kernel<<<ceil_div((int64_t)n_chunks * 32, 256), 256, 0, st>>>(density_grid, density_threshold, n_cells, masks, chunk_count));
    if (int e = check_launch("occ_mask")) return e;
    ARN_LAUNCH("small_scan_kernel", st, small_scan_kernel<<<1, 1024, 0, st>>>(chunk_count, n_chunks, chunk_prefix));
    if (int e = check_launch("small_scan")) return e;
    const float s_minus_half = (float)((double)s - (double)s / (double)grid_size);
    const float half_f = (float)((double)s / (double)grid_size);
    const float inv_gm1 = 1.0f / (float)(grid_size - 1);
    if (!sorted) {
        ARN_LAUNCH("pick_cells_kernel", st, pick_cells_kernel<false><<<ceil_div(2 * M, 256), 256, 0, st>>>(coords1, u, M, masks, chunk_prefix, n_chunks, rnd,
                   inv_gm1, s_minus_half, half_f, indices, xyzs, nullptr, nullptr));
        return check_launch("pick_cells");
    }
    ARN_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_cells * sizeof(int32_t), st));
    ARN_LAUNCH("pick_cells_kernel", st, pick_cells_kernel<true><<<ceil_div(2 * M, 256), 256, 0, st>>>(coords1, u, M, masks, chunk_prefix, n_chunks, rnd,
               inv_gm1, s_minus_half, half_f, nullptr, nullptr, idx_tmp, hist));
    if (int e = check_launch("pick_cells (sorted)")) return e;
    ARN_LAUNCH("hist_chunk_sum_kernel", st, hist_chunk_sum_kernel<<<ceil_div((int64_t)n_chunks * 32, 256), 256, 0, st>>>(hist, n_cells, n_chunks, chunk_tot));
    if (int e = check_launch("hist_chunk_sum")) return e;
    ARN_LAUNCH("small_scan_kernel", st, small_scan_kernel<<<1, 1024, 0, st>>>(chunk_tot, n_chunks, chunk_base));
    if (int e = check_launch("small_scan")) return e;
    ARN_LAUNCH("hist_chunk_scan_kernel", st, hist_chunk_scan_kernel<<<ceil_div((int64_t)n_chunks * 32, 256), 256, 0, st>>>(hist, n_cells, n_chunks, chunk_base));
    if (int e = check_launch("hist_chunk_scan")) return e;
    ARN_LAUNCH("place_cells_kernel", st, place_cells_kernel<<<ceil_div(2 * M, 256), 256, 0, st>>>(idx_tmp, hist, 2 * M, rnd, inv_gm1, s_minus_half, half_f, indices, xyzs));
    return check_launch("place_cells");
}
extern "C" ARN_API int arn_grid_sample_cells(const float* density_grid, float density_threshold, int grid_size, float s, const int32_t* coords1,
                                             const int64_t* u, int64_t M, const float* rnd, void* scratch, int64_t* indices, float* xyzs,
                                             arn_stream_t stream) {
    return grid_sample_cells_impl(density_grid, density_threshold, grid_size, s, coords1, u, M, rnd, scratch, indices, xyzs, false, stream);
}
extern "C" ARN_API int arn_grid_sample_cells_sorted(const float* density_grid, float density_threshold, int grid_size, float s, const int32_t* coords1,
                                                    const int64_t* u, int64_t M, const float* rnd, void* scratch, int64_t* indices, float* xyzs,
                                                    arn_stream_t stream) {
    return grid_sample_cells_impl(density_grid, density_threshold, grid_size, s, coords1, u, M, rnd, scratch, indices, xyzs, true, stream);
}
extern "C" ARN_API int arn_grid_scatter(float* dst, int64_t n_dst, const int64_t* indices, const float* src, int64_t n, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0 && n_dst >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(dst && indices && src, "null pointer");
    ARN_LAUNCH("scatter_f32_kernel", (cudaStream_t)stream, scatter_f32_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, indices, src, n, n_dst));
    return check_launch("grid_scatter");
}

extern "C" ARN_API int arn_density_grid_update(float* density_grid, const float* density_tmp, const float* decay_cells, float decay,
                                               float density_threshold, int64_t n_cells, uint8_t* density_bitfield, void* scratch,
                                               arn_stream_t stream) {
    ARN_REQUIRE(n_cells > 0 && n_cells % 8 == 0, "n_cells must be a positive multiple of 8");
    ARN_REQUIRE(density_grid && density_tmp && density_bitfield && scratch, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_parts = (int)min((int64_t)ARN_GRID_UPDATE_PARTS, (n_cells + 255) / 256);
    double* part_sum = (double*)scratch; int64_t* part_cnt = (int64_t*)(part_sum + ARN_GRID_UPDATE_PARTS);
    float* thr = (float*)(part_cnt + ARN_GRID_UPDATE_PARTS);
    ARN_LAUNCH("grid_update_kernel", st, grid_update_kernel<<<n_parts, 256, 0, st>>>(density_grid, density_tmp, decay_cells, decay, n_cells, part_sum, part_cnt));
    if (int e = check_launch("grid_update")) return e;
    ARN_LAUNCH("grid_threshold_kernel", st, grid_threshold_kernel<<<1, 256, 0, st>>>(part_sum, part_cnt, n_parts, density_threshold, thr));
    if (int e = check_launch("grid_threshold")) return e;
    const int64_t n_bytes = n_cells / 8;
    ARN_LAUNCH("packbits_devthr_kernel", st, packbits_devthr_kernel<<<ceil_div((n_bytes + 3) / 4, 256), 256, 0, st>>>(density_grid, thr, density_bitfield, n_bytes));
    return check_launch("packbits_devthr");
}

static int check_march_cfg(int cascades, int grid_size, int max_samples) {
    if (cascades < 1 || cascades > 16) { set_error("march: cascades out of range [1,16]"); return ARN_E_INVALID; }
    if (grid_size < 1 || grid_size > 1024) { set_error("march: grid_size out of range [1,1024]"); return ARN_E_INVALID; }
    if (max_samples < 1) { set_error("march: max_samples must be >= 1"); return ARN_E_INVALID; }
    return ARN_OK;
}

// t_scratch / count_scratch are carried through a second entry point so the published signature stays the reference's
// argument list.  count_scratch (n_rays i32, optional): compact per-ray counts for the multi-CTA scan.
extern "C" ARN_API int arn_march_train_count_ex(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                                        const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                                        float exp_step_factor, const float* noise, int max_samples, int64_t* rays_a,
                                        int32_t* counter, float* t_scratch, int32_t* count_scratch, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0, "bad size");
    ARN_REQUIRE(counter, "null counter");
    if (int e = check_march_cfg(cascades, grid_size, max_samples)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays == 0) { ARN_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int32_t), st)); return ARN_OK; }
    ARN_REQUIRE(rays_o && rays_d && hits_t && density_bitfield && noise && rays_a, "null pointer");
    const ArnMarchConsts c = arn_march_consts(cascades, grid_size, scale, scale, exp_step_factor, max_samples);
    if (!tunable(kTunMarchWarp)) {
        ARN_LAUNCH("march_train_count_kernel", st, march_train_count_kernel<<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, hits_t, n_rays, density_bitfield, c, noise, rays_a, t_scratch, count_scratch));
    } else {
        const bool const_dt = exp_step_factor == 0.0f && c.dt_hi >= 0.0f;  // calc_dt is the constant dt_lo
        const bool fast = cascades == 1 && grid_size <= 256;
        const int grid = ceil_div(n_rays * 32, 256);
#define ARN_MARCH_WARP(CD, FA) ARN_LAUNCH_PDL("march_train_count_warp_kernel", st, (march_train_count_warp_kernel<CD, FA>), grid, 256, 0, rays_o, rays_d, hits_t, n_rays, density_bitfield, c, noise, rays_a, t_scratch, count_scratch)
        if (const_dt && fast) ARN_MARCH_WARP(true, true);
        else if (const_dt) ARN_MARCH_WARP(true, false);
        else if (fast) ARN_MARCH_WARP(false, true);
        else ARN_MARCH_WARP(false, false);
#undef ARN_MARCH_WARP
    }
    if (int e = check_launch("march_train_count")) return e;
    if (count_scratch) ARN_LAUNCH_PDL("rays_scan_compact_kernel", st, (rays_scan_compact_kernel), ceil_div(n_rays, 1024), 1024, 0, count_scratch, n_rays, rays_a, counter);
    else ARN_LAUNCH("rays_scan_kernel", st, rays_scan_kernel<<<1, 1024, 0, st>>>(rays_a, n_rays, counter));
    return check_launch("rays_scan");
}
extern "C" ARN_API int arn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                                     const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                                     float exp_step_factor, const float* noise, int max_samples, int64_t* rays_a,
                                     int32_t* counter, arn_stream_t stream) {
    return arn_march_train_count_ex(rays_o, rays_d, hits_t, n_rays, density_bitfield, cascades, grid_size, scale,
                                    exp_step_factor, noise, max_samples, rays_a, counter, nullptr, nullptr, stream);
}

extern "C" ARN_API int arn_march_train_emit_ex(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                                       const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                                       float exp_step_factor, const float* noise, int max_samples, const int64_t* rays_a,
                                       const float* t_scratch, float* xyzs, float* dirs, float* deltas, float* ts,
                                       int64_t capacity, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && capacity >= 0, "bad size");
    if (int e = check_march_cfg(cascades, grid_size, max_samples)) return e;
    if (n_rays == 0 || capacity == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && hits_t && density_bitfield && noise && rays_a && xyzs && dirs && deltas && ts, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const ArnMarchConsts c = arn_march_consts(cascades, grid_size, scale, scale, exp_step_factor, max_samples);
    if (t_scratch) {
        ARN_LAUNCH_PDL("march_train_emit_kernel", st, (march_train_emit_kernel), ceil_div(capacity, 256), 256, 0, rays_o, rays_d, n_rays, c, rays_a, t_scratch, capacity, nullptr, xyzs, dirs, deltas, ts);
        return check_launch("march_train_emit");
    }
    ARN_LAUNCH("march_train_remarch_kernel", st, march_train_remarch_kernel<<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, hits_t, n_rays, density_bitfield, c, noise, rays_a, xyzs, dirs, deltas, ts));
    return check_launch("march_train_remarch");
}
// Device-count form used by the fused training step: `capacity` bounds the outputs, the real count is counter[0].
extern "C" ARN_API int arn_march_train_emit_dyn(const float* rays_o, const float* rays_d, int64_t n_rays, int cascades, int grid_size, float scale,
                                                float exp_step_factor, int max_samples, const int64_t* rays_a, const float* t_scratch,
                                                const int32_t* counter, float* xyzs, float* dirs, float* deltas, float* ts, int64_t capacity,
                                                arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && capacity >= 0, "bad size");
    if (int e = check_march_cfg(cascades, grid_size, max_samples)) return e;
    if (n_rays == 0 || capacity == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && rays_a && t_scratch && counter && xyzs && dirs && deltas && ts, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const ArnMarchConsts c = arn_march_consts(cascades, grid_size, scale, scale, exp_step_factor, max_samples);
    const int grid = (int)min((int64_t)148 * 16, (capacity + 255) / 256);
    ARN_LAUNCH_PDL("march_train_emit_kernel", st, (march_train_emit_kernel), grid, 256, 0, rays_o, rays_d, n_rays, c, rays_a, t_scratch, capacity, counter, xyzs, dirs, deltas, ts);
    return check_launch("march_train_emit_dyn");
}

extern "C" ARN_API int arn_march_train_emit(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                                    const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                                    float exp_step_factor, const float* noise, int max_samples, const int64_t* rays_a,
                                    float* xyzs, float* dirs, float* deltas, float* ts, int64_t capacity, arn_stream_t stream) {
    return arn_march_train_emit_ex(rays_o, rays_d, hits_t, n_rays, density_bitfield, cascades, grid_size, scale, exp_step_factor,
                                   noise, max_samples, rays_a, nullptr, xyzs, dirs, deltas, ts, capacity, stream);
}

extern "C" ARN_API int arn_march_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_indices,
                              int64_t n_alive, const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                              float exp_step_factor, int n_samples, int max_samples, float* xyzs, float* dirs, float* deltas,
                              float* ts, int32_t* n_eff_samples, arn_stream_t stream) {
    ARN_REQUIRE(n_alive >= 0 && n_samples >= 1, "bad size");
    if (int e = check_march_cfg(cascades, grid_size, max_samples)) return e;
    if (n_alive == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && hits_t && alive_indices && density_bitfield && xyzs && dirs && deltas && ts && n_eff_samples, "null pointer");
    // raymarching.cu:370,399: the test kernel passes `cascades` where calc_dt expects `scale`
    const ArnMarchConsts c = arn_march_consts(cascades, grid_size, scale, (float)cascades, exp_step_factor, max_samples);
    ARN_LAUNCH("march_test_kernel", (cudaStream_t)stream, march_test_kernel<<<ceil_div(n_alive, 128), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, hits_t, alive_indices, n_alive,
                                                                              density_bitfield, c, n_samples, xyzs, dirs, deltas, ts, n_eff_samples));
    return check_launch("march_test");
}

extern "C" ARN_API int arn_composite_train_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                      const int64_t* rays_a, int64_t n_rays, int64_t n_samples, float T_threshold,
                                      int64_t* total_samples, float* opacity, float* depth, float* rgb, float* ws, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && n_samples >= 0, "bad size");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rays_a && total_samples && opacity && depth && rgb, "null pointer");
    ARN_REQUIRE(n_samples == 0 || (sigmas && rgbs && deltas && ts && ws), "null pointer");
    ARN_LAUNCH_PDL("composite_train_fw_kernel", (cudaStream_t)stream, (composite_train_fw_kernel<false>), ceil_div(n_rays * 32, kCompBlock), kCompBlock, 0, sigmas, rgbs, deltas, ts, rays_a, n_rays, T_threshold,
                                                                                          total_samples, opacity, depth, rgb, ws, LossEpilogue{}, n_samples, nullptr, nullptr);
    return check_launch("composite_train_fw");
}

extern "C" int arn_composite_train_fw_loss_ex(const float*, const float*, const float*, const float*, const int64_t*, int64_t, int64_t, float, int64_t*, float*,
                                              float*, float*, float*, const float*, const float*, float, float, float, float, float*, float*, float*, float*,
                                              float*, int, float*, float*, arn_stream_t);
// Compositing + NeRFLoss in one launch (the fused training step); rays_a must be in canonical ray order (ray_idx == row).
extern "C" ARN_API int arn_composite_train_fw_loss(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                                   const int64_t* rays_a, int64_t n_rays, int64_t n_samples, float T_threshold,
                                                   int64_t* total_samples, float* opacity, float* depth, float* rgb, float* ws,
                                                   const float* target, const float* bg_host, float lambda_opacity, float lambda_depth,
                                                   float grid_scale, float grad_scale, float* rgb_out, float* dL_drgb, float* dL_dopacity,
                                                   float* dL_ddepth, float* loss_out, arn_stream_t stream) {
    return arn_composite_train_fw_loss_ex(sigmas, rgbs, deltas, ts, rays_a, n_rays, n_samples, T_threshold, total_samples, opacity, depth, rgb, ws, target, bg_host,
                                          lambda_opacity, lambda_depth, grid_scale, grad_scale, rgb_out, dL_drgb, dL_dopacity, dL_ddepth, loss_out, 1, nullptr, nullptr,
                                          stream);
}
// zero_loss = 0: *loss_out has been zeroed by the caller (the fused step does it in arn_train_march, off the serial chain);
// bw_dsigmas / bw_drgbs != NULL: the compositing backward of every ray runs in the same launch, behind the ray's loss terms
extern "C" int arn_composite_train_fw_loss_ex(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                              const int64_t* rays_a, int64_t n_rays, int64_t n_samples, float T_threshold,
                                              int64_t* total_samples, float* opacity, float* depth, float* rgb, float* ws,
                                              const float* target, const float* bg_host, float lambda_opacity, float lambda_depth,
                                              float grid_scale, float grad_scale, float* rgb_out, float* dL_drgb, float* dL_dopacity,
                                              float* dL_ddepth, float* loss_out, int zero_loss, float* bw_dsigmas, float* bw_drgbs, arn_stream_t stream) {
    ARN_REQUIRE(n_rays > 0 && n_samples >= 0, "bad size");
    ARN_REQUIRE(!bw_dsigmas == !bw_drgbs, "the fused backward needs both sample-gradient buffers");
    ARN_REQUIRE(rays_a && total_samples && opacity && depth && rgb && sigmas && rgbs && deltas && ts && ws, "null pointer");
    ARN_REQUIRE(target && bg_host && dL_drgb && dL_dopacity && dL_ddepth && loss_out, "null pointer (loss)");
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_loss) ARN_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    LossEpilogue L{target, {bg_host[0], bg_host[1], bg_host[2]}, lambda_opacity, lambda_depth, grid_scale, grad_scale, rgb_out, dL_drgb, dL_dopacity,
                   dL_ddepth, loss_out};
    ARN_LAUNCH_PDL("composite_train_fw_loss_kernel", st, (composite_train_fw_kernel<true>), ceil_div(n_rays * 32, kCompBlock), kCompBlock, 0, sigmas, rgbs, deltas, ts, rays_a, n_rays, T_threshold,
                                                                                          total_samples, opacity, depth, rgb, ws, L, n_samples, bw_dsigmas, bw_drgbs);
    return check_launch("composite_train_fw_loss");
}

extern "C" ARN_API int arn_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb, const float* dL_dws,
                                      const float* sigmas, const float* rgbs, const float* ws, const float* deltas, const float* ts,
                                      const int64_t* rays_a, const float* opacity, const float* depth, const float* rgb,
                                      int64_t n_rays, int64_t n_samples, float T_threshold, float* dL_dsigmas, float* dL_drgbs,
                                      arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && n_samples >= 0, "bad size");
    if (n_rays == 0 || n_samples == 0) return ARN_OK;
    ARN_REQUIRE(dL_dopacity && dL_ddepth && dL_drgb && sigmas && rgbs && ws && deltas && ts && rays_a && opacity && depth && rgb && dL_dsigmas && dL_drgbs,
                "null pointer");
    ARN_LAUNCH("composite_train_bw_kernel", (cudaStream_t)stream, composite_train_bw_kernel<<<ceil_div(n_rays * 32, kCompBlock), kCompBlock, 0, (cudaStream_t)stream>>>(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws,
                                                                                          deltas, ts, rays_a, opacity, depth, rgb, n_rays,
                                                                                          T_threshold, dL_dsigmas, dL_drgbs, n_samples));
    return check_launch("composite_train_bw");
}

extern "C" ARN_API int arn_composite_test_fw(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                     int64_t* alive_indices, int64_t n_alive, int n_samples, float T_threshold,
                                     const int32_t* n_eff_samples, float* opacity, float* depth, float* rgb, arn_stream_t stream) {
    ARN_REQUIRE(n_alive >= 0 && n_samples >= 1, "bad size");
    if (n_alive == 0) return ARN_OK;
    ARN_REQUIRE(sigmas && rgbs && deltas && ts && alive_indices && n_eff_samples && opacity && depth && rgb, "null pointer");
    ARN_LAUNCH("composite_test_fw_kernel", (cudaStream_t)stream, composite_test_fw_kernel<<<ceil_div(n_alive, 256), 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, deltas, ts, alive_indices, n_alive, n_samples,
                                                                                     T_threshold, n_eff_samples, opacity, depth, rgb));
    return check_launch("composite_test_fw");
}

extern "C" ARN_API int arn_distortion_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays,
                                 int64_t n_samples, float* loss, float* ws_inclusive_scan, float* wts_inclusive_scan, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && n_samples >= 0, "bad size");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rays_a && loss, "null pointer");
    ARN_REQUIRE(n_samples == 0 || (ws && deltas && ts && ws_inclusive_scan && wts_inclusive_scan), "null pointer");
    ARN_LAUNCH("distortion_fw_kernel", (cudaStream_t)stream, distortion_fw_kernel<<<ceil_div(n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(ws, deltas, ts, rays_a, n_rays, loss, ws_inclusive_scan, wts_inclusive_scan));
    return check_launch("distortion_fw");
}
extern "C" ARN_API int arn_distortion_bw(const float* dL_dloss, const float* ws_inclusive_scan, const float* wts_inclusive_scan, const float* ws,
                                 const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_rays, int64_t n_samples,
                                 float* dL_dws, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && n_samples >= 0, "bad size");
    if (n_rays == 0 || n_samples == 0) return ARN_OK;
    ARN_REQUIRE(dL_dloss && ws_inclusive_scan && wts_inclusive_scan && ws && deltas && ts && rays_a && dL_dws, "null pointer");
    ARN_LAUNCH("distortion_bw_kernel", (cudaStream_t)stream, distortion_bw_kernel<<<ceil_div(n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts,
                                                                                     rays_a, n_rays, dL_dws));
    return check_launch("distortion_bw");
}
extern "C" ARN_API int arn_march_train_bw(const float* dL_dxyzs, const float* dL_ddirs, const float* ts, const int64_t* rays_a, int64_t n_rays,
                                  float* dL_drays_o, float* dL_drays_d, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0, "bad size");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(dL_dxyzs && ts && rays_a && dL_drays_o && dL_drays_d, "null pointer");
    ARN_LAUNCH("march_train_bw_kernel", (cudaStream_t)stream, march_train_bw_kernel<<<ceil_div(n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(dL_dxyzs, dL_ddirs, ts, rays_a, n_rays, dL_drays_o, dL_drays_d));
    return check_launch("march_train_bw");
}


// ---------------------------------------------------------------------------------------------------------------------
// One iteration of the test-time render loop (rendering.py:189-236) in one host call, no host round trip inside:
//   march the alive rays by <= S samples (t, dt, count) -> scan -> compact sample list -> field (inference, count on the
//   device) -> front-to-back compositing with ray kill -> order-preserving compaction of the alive list.
// The caller reads counts_out = (valid samples of this iteration, rays still alive) once per iteration to drive the
// reference's schedule (N_samples = max(min(N_rays // N_alive, 64), min_samples)).
extern "C" int arn_field_fw_tc_dyn(const float*, const float*, int64_t, const int32_t*, const float*, const float*, arn_levels_t, const void*, const void*,
                                   int, arn_field_ws_t, float*, float*, arn_stream_t);
extern "C" ARN_API int arn_render_test_iter(const arn_test_iter_t* c, arn_stream_t stream) {
    ARN_REQUIRE(c, "null config");
    ARN_REQUIRE(c->n_alive > 0 && c->n_samples >= 1 && c->capacity >= c->n_alive * c->n_samples, "bad sizes");
    if (int e = check_march_cfg(c->cascades, c->grid_size, c->max_samples)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = c->n_alive; const int S = c->n_samples;
    // raymarching.cu:370,399: the test kernel passes `cascades` where calc_dt expects `scale`
    const ArnMarchConsts mc = arn_march_consts(c->cascades, c->grid_size, c->scale, (float)c->cascades, c->exp_step_factor, c->max_samples);
    ARN_LAUNCH("march_test_lite_kernel", st, march_test_lite_kernel<<<ceil_div(n, 128), 128, 0, st>>>(c->rays_o, c->rays_d, c->hits_t, c->alive, n, c->density_bitfield, mc, S,
                                                                                                  c->deltas, c->ts, c->n_eff));
    if (int e = check_launch("march_test_lite")) return e;
    ARN_LAUNCH_PDL("rays_scan_compact_kernel", st, (rays_scan_compact_kernel), ceil_div(n, 1024), 1024, 0, c->n_eff, n, c->rays_a, c->counts);
    if (int e = check_launch("rays_scan")) return e;
    ARN_LAUNCH("emit_test_kernel", st, emit_test_kernel<<<ceil_div(n * S, 256), 256, 0, st>>>(c->rays_o, c->rays_d, c->alive, n, S, c->rays_a, c->ts, c->xyzs, c->dirs));
    if (int e = check_launch("emit_test")) return e;
    if (int e = arn_field_fw_tc_dyn(c->xyzs, c->dirs, c->capacity, c->counts, c->xyz_min_host, c->xyz_max_host, c->levels, c->params_xyz_f16,
                                    c->params_rgb_f16, c->rgb_act, c->ws, c->sigmas, c->rgbs, stream)) return e;
    ARN_LAUNCH("composite_test_compact_kernel", st, composite_test_compact_kernel<<<ceil_div(n, 256), 256, 0, st>>>(c->sigmas, c->rgbs, c->deltas, c->ts, c->alive, n, S, c->T_threshold,
                                                                                                               c->rays_a, c->opacity, c->depth, c->rgb, c->n_eff,
                                                                                                               (unsigned long long*)c->total_samples));
    if (int e = check_launch("composite_test_compact")) return e;
    // n_eff now holds the keep flags; the scan reuses rays_a and leaves (rays kept, n) in counts_alive
    ARN_LAUNCH_PDL("rays_scan_compact_kernel", st, (rays_scan_compact_kernel), ceil_div(n, 1024), 1024, 0, c->n_eff, n, c->rays_a, c->counts_alive);
    if (int e = check_launch("rays_scan")) return e;
    ARN_LAUNCH("alive_scatter_kernel", st, alive_scatter_kernel<<<ceil_div(n, 256), 256, 0, st>>>(c->alive, n, c->rays_a, c->alive_out));
    return check_launch("alive_scatter");
}

// Far clamp of the rays of a frame, once, in front of the test loop (march_test_far_clamp_kernel): hits_t (R,2) in place.
extern "C" ARN_API int arn_march_test_far_clamp(const float* rays_o, const float* rays_d, float* hits_t, int64_t n_rays,
                                                const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                                                float exp_step_factor, int max_samples, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0, "bad size");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && hits_t && density_bitfield, "null pointer");
    if (int e = check_march_cfg(cascades, grid_size, max_samples)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const ArnMarchConsts mc = arn_march_consts(cascades, grid_size, scale, (float)cascades, exp_step_factor, max_samples);  // the test march's dt (Q2)
    if (cascades == 1 && grid_size <= 256)
        ARN_LAUNCH("march_test_far_clamp_kernel", st, march_test_far_clamp_kernel<true><<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, hits_t, n_rays, density_bitfield, mc));
    else
        ARN_LAUNCH("march_test_far_clamp_kernel", st, march_test_far_clamp_kernel<false><<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, hits_t, n_rays, density_bitfield, mc));
    return check_launch("march_test_far_clamp");
}

// Numerator of the device-driven loop's schedule: the frame's ray count (the reference's N_rays // N_alive) unless the caller
// asks for more samples per iteration (cfg->schedule_rays > N_rays; pixels do not depend on the slicing).
static inline int64_t schedule_rays(const arn_test_iter_t* c) { return c->schedule_rays > c->n_alive ? c->schedule_rays : c->n_alive; }

// The iteration above with the loop control on the device (kernels: "test loop, device-driven").  c->n_alive = N_rays of the
// frame (the numerator of the schedule), c->n_samples is ignored; n_upper bounds the device-side n_alive (grid sizing only).
extern "C" ARN_API int arn_render_test_step(const arn_test_iter_t* c, const int32_t* state_in, int32_t* state_out, int32_t* partial,
                                            int min_samples, int budget_samples, int64_t n_upper, arn_stream_t stream) {
    ARN_REQUIRE(c && state_in && state_out && partial, "null pointer");
    ARN_REQUIRE(c->n_alive > 0 && min_samples >= 1 && n_upper >= 0 && n_upper <= c->n_alive, "bad sizes");
    ARN_REQUIRE(c->capacity >= c->n_alive * (int64_t)min_samples && c->capacity >= schedule_rays(c), "capacity must hold max(N_rays * min_samples, schedule_rays) samples");
    if (int e = check_march_cfg(c->cascades, c->grid_size, c->max_samples)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nu = n_upper > 0 ? n_upper : 1;
    const int64_t sched = schedule_rays(c);
    const int64_t samples_upper = nu * min_samples > sched ? nu * min_samples : sched;  // n * S <= max(schedule_rays, n * min_samples)
    const ArnMarchConsts mc = arn_march_consts(c->cascades, c->grid_size, c->scale, (float)c->cascades, c->exp_step_factor, c->max_samples);
    const int g128 = (int)min((int64_t)148 * 16, (nu + 127) / 128), g1024 = (int)min((int64_t)148 * 4, (nu + 1023) / 1024);  // grid-stride: full residency is enough
    // few rays, many samples each (N_samples >= 9 once n_alive <= N_rays / 9): one warp per ray
    const bool warp_march = tunable(kTunMarchWarp) != 0 && nu <= 65536 && nu * 9 <= c->n_alive;
    if (warp_march) {
        const int gw = (int)min((int64_t)148 * 16, (nu + 7) / 8);
        const bool cd = c->exp_step_factor == 0.0f, fa = c->cascades == 1 && c->grid_size <= 256;
#define ARN_TEST_WARP(CD, FA) ARN_LAUNCH("march_test_warp_kernel", st, (march_test_warp_kernel<CD, FA><<<gw, 256, 0, st>>>(c->rays_o, c->rays_d, c->hits_t, c->alive, state_in, c->density_bitfield, mc, c->deltas, c->ts, c->n_eff)))
        if (cd && fa) ARN_TEST_WARP(true, true); else if (cd) ARN_TEST_WARP(true, false); else if (fa) ARN_TEST_WARP(false, true); else ARN_TEST_WARP(false, false);
#undef ARN_TEST_WARP
        if (int e = check_launch("march_test_warp")) return e;
    } else {
        if (c->cascades == 1 && c->grid_size <= 256)
            ARN_LAUNCH("march_test_dyn_kernel", st, march_test_dyn_kernel<true><<<g128, 128, 0, st>>>(c->rays_o, c->rays_d, c->hits_t, c->alive, state_in, c->density_bitfield, mc,
                                                                                                  c->deltas, c->ts, c->n_eff, partial));
        else
            ARN_LAUNCH("march_test_dyn_kernel", st, march_test_dyn_kernel<false><<<g128, 128, 0, st>>>(c->rays_o, c->rays_d, c->hits_t, c->alive, state_in, c->density_bitfield, mc,
                                                                                                   c->deltas, c->ts, c->n_eff, partial));
        if (int e = check_launch("march_test_dyn")) return e;
    }
    ARN_LAUNCH_PDL("scan_test_dyn_kernel", st, (scan_test_dyn_kernel), g1024, 1024, 0, c->n_eff, warp_march ? nullptr : partial, state_in, c->rays_a, c->counts);
    if (int e = check_launch("scan_test_dyn")) return e;
    const int g_emit = (int)min((int64_t)148 * 32, (samples_upper + 255) / 256);
    ARN_LAUNCH("emit_test_dyn_kernel", st, emit_test_dyn_kernel<<<g_emit, 256, 0, st>>>(c->rays_o, c->rays_d, c->alive, state_in, c->rays_a, c->ts, c->xyzs, c->dirs));
    if (int e = check_launch("emit_test_dyn")) return e;
    if (int e = arn_field_fw_tc_dyn(c->xyzs, c->dirs, samples_upper, c->counts, c->xyz_min_host, c->xyz_max_host, c->levels, c->params_xyz_f16,
                                    c->params_rgb_f16, c->rgb_act, c->ws, c->sigmas, c->rgbs, stream)) return e;
    ARN_LAUNCH_PDL("composite_test_dyn_kernel", st, (composite_test_dyn_kernel), g128, 128, 0, c->sigmas, c->rgbs, c->deltas, c->ts, c->alive, state_in, c->T_threshold,
                                                                                            c->rays_a, c->opacity, c->depth, c->rgb, c->n_eff, partial,
                                                                                            (unsigned long long*)c->total_samples, nullptr);
    if (int e = check_launch("composite_test_dyn")) return e;
    ARN_LAUNCH_PDL("alive_compact_dyn_kernel", st, (alive_compact_dyn_kernel), g1024, 1024, 0, c->alive, c->n_eff, partial, state_in, c->counts, c->alive_out,
                                                                                           c->counts_alive, state_out, sched, min_samples, budget_samples);
    return check_launch("alive_compact_dyn");
}

// The frame's samples marched once (march_test_all_kernel): ts_all (stride x n_rays floats, sample-major), totals / cursor (n_rays).
extern "C" ARN_API int arn_march_test_all(const float* rays_o, const float* rays_d, const float* hits_t, int64_t n_rays,
                                          const uint8_t* density_bitfield, int cascades, int grid_size, float scale,
                                          float exp_step_factor, int max_samples, int stride, float* ts_all, int32_t* totals,
                                          int32_t* cursor, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0 && stride >= 1, "bad size");
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rays_o && rays_d && hits_t && density_bitfield && ts_all && totals && cursor, "null pointer");
    if (int e = check_march_cfg(cascades, grid_size, max_samples)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const ArnMarchConsts mc = arn_march_consts(cascades, grid_size, scale, (float)cascades, exp_step_factor, max_samples);  // the test march's dt (Q2)
    if (cascades == 1 && grid_size <= 256)
        ARN_LAUNCH("march_test_all_kernel", st, march_test_all_kernel<true><<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, hits_t, n_rays, density_bitfield, mc, stride, ts_all, totals, cursor));
    else
        ARN_LAUNCH("march_test_all_kernel", st, march_test_all_kernel<false><<<ceil_div(n_rays, 128), 128, 0, st>>>(rays_o, rays_d, hits_t, n_rays, density_bitfield, mc, stride, ts_all, totals, cursor));
    return check_launch("march_test_all");
}

// arn_render_test_step on a frame marched by arn_march_test_all: the iteration slices its samples instead of marching.
extern "C" ARN_API int arn_render_test_step_pre(const arn_test_iter_t* c, const int32_t* state_in, int32_t* state_out, int32_t* partial,
                                                const float* ts_all, const int32_t* totals, int32_t* cursor, int min_samples,
                                                int budget_samples, int64_t n_upper, arn_stream_t stream) {
    ARN_REQUIRE(c && state_in && state_out && partial && ts_all && totals && cursor, "null pointer");
    ARN_REQUIRE(c->n_alive > 0 && min_samples >= 1 && n_upper >= 0 && n_upper <= c->n_alive, "bad sizes");
    ARN_REQUIRE(c->capacity >= c->n_alive * (int64_t)min_samples && c->capacity >= schedule_rays(c), "capacity must hold max(N_rays * min_samples, schedule_rays) samples");
    if (int e = check_march_cfg(c->cascades, c->grid_size, c->max_samples)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nu = n_upper > 0 ? n_upper : 1;
    const int64_t sched = schedule_rays(c);
    const int64_t samples_upper = nu * min_samples > sched ? nu * min_samples : sched;  // n * S <= max(schedule_rays, n * min_samples)
    const ArnMarchConsts mc = arn_march_consts(c->cascades, c->grid_size, c->scale, (float)c->cascades, c->exp_step_factor, c->max_samples);
    const int g128 = (int)min((int64_t)148 * 16, (nu + 127) / 128), g1024 = (int)min((int64_t)148 * 4, (nu + 1023) / 1024);  // grid-stride: full residency is enough
    ARN_LAUNCH_PDL("neff_test_pre_kernel", st, (neff_test_pre_kernel), g128, 128, 0, c->alive, state_in, totals, cursor, c->n_eff, partial);
    if (int e = check_launch("neff_test_pre")) return e;
    ARN_LAUNCH_PDL("scan_test_dyn_kernel", st, (scan_test_dyn_kernel), g1024, 1024, 0, c->n_eff, partial, state_in, c->rays_a, c->counts);
    if (int e = check_launch("scan_test_dyn")) return e;
    const int g_emit = (int)min((int64_t)148 * 32, (samples_upper + 255) / 256);
    ARN_LAUNCH_PDL("emit_test_pre_kernel", st, (emit_test_pre_kernel), g_emit, 256, 0, c->rays_o, c->rays_d, c->alive, state_in, c->rays_a, ts_all, cursor, c->n_alive, mc,
                                                                                     c->deltas, c->ts, c->xyzs, c->dirs);
    if (int e = check_launch("emit_test_pre")) return e;
    if (int e = arn_field_fw_tc_dyn(c->xyzs, c->dirs, samples_upper, c->counts, c->xyz_min_host, c->xyz_max_host, c->levels, c->params_xyz_f16,
                                    c->params_rgb_f16, c->rgb_act, c->ws, c->sigmas, c->rgbs, stream)) return e;
    ARN_LAUNCH_PDL("composite_test_dyn_kernel", st, (composite_test_dyn_kernel), g128, 128, 0, c->sigmas, c->rgbs, c->deltas, c->ts, c->alive, state_in, c->T_threshold,
                                                                                            c->rays_a, c->opacity, c->depth, c->rgb, c->n_eff, partial,
                                                                                            (unsigned long long*)c->total_samples, cursor);
    if (int e = check_launch("composite_test_dyn")) return e;
    ARN_LAUNCH_PDL("alive_compact_dyn_kernel", st, (alive_compact_dyn_kernel), g1024, 1024, 0, c->alive, c->n_eff, partial, state_in, c->counts, c->alive_out,
                                                                                           c->counts_alive, state_out, sched, min_samples, budget_samples);
    return check_launch("alive_compact_dyn");
}


// arn_render_test_step_pre in four launches (kernels above): sync = 4 x int32 of device scratch, zero before the first iteration
// of a frame (the last block of every iteration leaves it zero again).  The alive lists and the compact sample list come out in
// a scheduling-dependent ORDER; every per-ray result, the kill pattern and total_samples are those of the seven-launch form.
extern "C" ARN_API int arn_render_test_step_fused(const arn_test_iter_t* c, const int32_t* state_in, int32_t* state_out, int32_t* sync,
                                                  const float* ts_all, const int32_t* totals, int32_t* cursor, int min_samples,
                                                  int budget_samples, int64_t n_upper, arn_stream_t stream) {
    ARN_REQUIRE(c && state_in && state_out && sync && ts_all && totals && cursor, "null pointer");
    ARN_REQUIRE(c->n_alive > 0 && min_samples >= 1 && n_upper >= 0 && n_upper <= c->n_alive, "bad sizes");
    ARN_REQUIRE(c->capacity >= c->n_alive * (int64_t)min_samples && c->capacity >= schedule_rays(c), "capacity must hold max(N_rays * min_samples, schedule_rays) samples");
    if (int e = check_march_cfg(c->cascades, c->grid_size, c->max_samples)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nu = n_upper > 0 ? n_upper : 1;
    const int64_t sched = schedule_rays(c);
    const int64_t samples_upper = nu * min_samples > sched ? nu * min_samples : sched;  // n * S <= max(schedule_rays, n * min_samples)
    const ArnMarchConsts mc = arn_march_consts(c->cascades, c->grid_size, c->scale, (float)c->cascades, c->exp_step_factor, c->max_samples);
    const int g128 = (int)min((int64_t)148 * 16, (nu + 127) / 128);
    ARN_LAUNCH("test_slice_emit_kernel", st, test_slice_emit_kernel<<<g128, 128, 0, st>>>(c->rays_o, c->rays_d, c->alive, state_in, ts_all, totals, cursor, c->n_alive, mc,
                                                                                         c->n_eff, c->rays_a, sync, c->deltas, c->ts, c->xyzs, c->dirs, c->capacity));
    if (int e = check_launch("test_slice_emit")) return e;
    if (int e = arn_field_fw_tc_dyn(c->xyzs, c->dirs, samples_upper, sync, c->xyz_min_host, c->xyz_max_host, c->levels, c->params_xyz_f16,
                                    c->params_rgb_f16, c->rgb_act, c->ws, c->sigmas, c->rgbs, stream)) return e;
    ARN_LAUNCH("test_composite_keep_kernel", st, test_composite_keep_kernel<<<g128, 128, 0, st>>>(c->sigmas, c->rgbs, c->deltas, c->ts, c->alive, state_in, c->T_threshold,
                                                                                                 c->rays_a, c->opacity, c->depth, c->rgb, (unsigned long long*)c->total_samples,
                                                                                                 cursor, c->alive_out, sync, c->counts_alive, state_out, sched, min_samples,
                                                                                                 budget_samples));
    return check_launch("test_composite_keep");
}
