/*
 * oracle/oracle_vren.c -- CPU restatement of the reference's `vren` CUDA extension (models/csrc).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libarnerf.so) never links or calls it.
 *
 * Parity status: PINNED against the real reference kernels.  oracle/_ref/vren.so is the unmodified
 * reference compiled for sm_100a (oracle/build_ref.sh); tests/golden/make_golden.py runs it on a B200
 * and commits its outputs as tests/golden/ (npz files); tests/test_oracle_golden.py checks this file against
 * them (bit-exact for the integer / marching work, 1e-5 for the __expf-based compositing).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Floating point: compiled with -ffp-contract=off; the fused multiply-adds that nvcc 12.9 emits for the
 * reference (verified in the sm_100a SASS, see DESIGN.md "FMA sites") are written as explicit fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SQRT3 1.73205080757f

/* ---- helper_math.h:280-283 clamp(f,a,b) = fmaxf(a, fminf(f,b)) ---- */
static inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
/* raymarching.cu:7 */
static inline float signf_(float x) { return copysignf(1.0f, x); }

/* raymarching.cu:11-13.  SASS: FMUL t*esf ; FMNMX(min) with hi ; FMNMX(max) with lo ;
 * lo = 1.7320508f / (float)max_samples (IEEE div), hi = (scale * 3.4641016f) / (float)grid_size. */
static inline float calc_dt(float t, float esf, int max_samples, int grid_size, float scale) {
    const float lo = SQRT3 / (float)max_samples;
    const float hi = (scale * (SQRT3 * 2)) / (float)grid_size;
    return clampf(t * esf, lo, hi);
}

/* raymarching.cu:19-23 */
static inline int mip_from_pos(float x, float y, float z, int cascades) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent; frexpf(mx, &exponent);
    int m = exponent + 1; if (m < 0) m = 0; if (m > cascades - 1) m = cascades - 1;
    return m;
}
/* raymarching.cu:29-32 */
static inline int mip_from_dt(float dt, int grid_size, int cascades) {
    int exponent; frexpf(dt * (float)grid_size, &exponent);
    int m = exponent; if (m < 0) m = 0; if (m > cascades - 1) m = cascades - 1;
    return m;
}

/* raymarching.cu:35-60 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

/* raymarching.cu:62-70 */
void orc_morton3d(int n, const int32_t* coords, int32_t* indices) {
    for (int i = 0; i < n; i++)
        indices[i] = (int32_t)morton3D((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1], (uint32_t)coords[3 * i + 2]);
}
/* raymarching.cu:90-101 */
void orc_morton3d_invert(int n, const int32_t* indices, int32_t* coords) {
    for (int i = 0; i < n; i++) {
        const int32_t ind = indices[i];
        coords[3 * i + 0] = (int32_t)morton3D_invert((uint32_t)(ind >> 0));
        coords[3 * i + 1] = (int32_t)morton3D_invert((uint32_t)(ind >> 1));
        coords[3 * i + 2] = (int32_t)morton3D_invert((uint32_t)(ind >> 2));
    }
}
/* raymarching.cu:122-141 (float instantiation; strict >, LSB-first) */
void orc_packbits(int n_bytes, const float* density_grid, float thr, uint8_t* bitfield) {
    for (int n = 0; n < n_bytes; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (density_grid[8 * (size_t)n + i] > thr) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* ------------------------------------------------------------------------------------------------
 * One evaluation of the marching loop body (raymarching.cu:205-232 == :246-277 == :367-402).
 * Returns 1 if the cell at t is occupied (then *t is NOT advanced: caller does t += dt), else 0 and
 * *t has been advanced past the empty cell by the do/while skip (raymarching.cu:225-232).
 * `dt_scale` is what the kernel passes as calc_dt's last argument: `scale` in the train kernel,
 * `cascades` in the test kernel (raymarching.cu:370,399 -- the reference's own quirk, kept).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float ox, oy, oz, dx, dy, dz, dxi, dyi, dzi;
} ray_t;

static inline int march_eval(const ray_t* r, const uint8_t* bitfield, int cascades, int grid_size, float scale,
                             float dt_scale, float esf, int max_samples, float* t_io,
                             float* x_o, float* y_o, float* z_o, float* dt_o) {
    const float t = *t_io;
    const uint32_t grid_size3 = (uint32_t)grid_size * grid_size * grid_size;
    const float grid_size_inv = 1.0f / (float)grid_size;
    const float G = (float)grid_size;
    /* :205  FFMA(d,t,o) */
    const float x = fmaf(r->dx, t, r->ox), y = fmaf(r->dy, t, r->oy), z = fmaf(r->dz, t, r->oz);
    const float dt = calc_dt(t, esf, max_samples, grid_size, dt_scale);
    const int mp = mip_from_pos(x, y, z, cascades), md = mip_from_dt(dt, grid_size, cascades);
    const int mip = mp > md ? mp : md;
    /* :211-212 */
    const float mip_bound = fminf(scalbnf(1.0f, mip - 1), scale);
    const float mip_bound_inv = 1.0f / mip_bound;
    /* :215-217  FFMA(x,inv,1) ; FMUL 0.5 ; FMUL G ; min(G-1) ; max(0) ; F2I.TRUNC (NaN -> 0) */
    const float fx = clampf((0.5f * fmaf(x, mip_bound_inv, 1.0f)) * G, 0.0f, G - 1.0f);
    const float fy = clampf((0.5f * fmaf(y, mip_bound_inv, 1.0f)) * G, 0.0f, G - 1.0f);
    const float fz = clampf((0.5f * fmaf(z, mip_bound_inv, 1.0f)) * G, 0.0f, G - 1.0f);
    const int nx = isnan(fx) ? 0 : (int)fx, ny = isnan(fy) ? 0 : (int)fy, nz = isnan(fz) ? 0 : (int)fz;
    /* :219-220 */
    const uint32_t idx = (uint32_t)mip * grid_size3 + morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const int occ = bitfield[idx / 8] & (1 << (idx % 8));
    *x_o = x; *y_o = y; *z_o = z; *dt_o = dt;
    if (occ) return 1;
    /* :225-227  I2F ; FADD .5 ; FFMA(sgn,.5,.) ; FMUL Ginv ; FFMA(.,2,-1) ; FFMA(mip_bound,.,-x) ; FMUL d_inv */
    const float tx = fmaf(mip_bound, fmaf(fmaf(signf_(r->dx), 0.5f, (float)nx + 0.5f) * grid_size_inv, 2.0f, -1.0f), -x) * r->dxi;
    const float ty = fmaf(mip_bound, fmaf(fmaf(signf_(r->dy), 0.5f, (float)ny + 0.5f) * grid_size_inv, 2.0f, -1.0f), -y) * r->dyi;
    const float tz = fmaf(mip_bound, fmaf(fmaf(signf_(r->dz), 0.5f, (float)nz + 0.5f) * grid_size_inv, 2.0f, -1.0f), -z) * r->dzi;
    /* :229-232 */
    const float t_target = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    float tt = t;
    do { tt += calc_dt(tt, esf, max_samples, grid_size, dt_scale); } while (tt < t_target);
    *t_io = tt;
    return 0;
}

static inline void load_ray(ray_t* r, const float* o, const float* d) {
    r->ox = o[0]; r->oy = o[1]; r->oz = o[2];
    r->dx = d[0]; r->dy = d[1]; r->dz = d[2];
    r->dxi = 1.0f / r->dx; r->dyi = 1.0f / r->dy; r->dzi = 1.0f / r->dz; /* :189 */
}

/* raymarching.cu:166-280 pass 1 (:184-234): per-ray sample counts.  n_samples[r] for every ray. */
void orc_march_train_count(int n_rays, const float* rays_o, const float* rays_d, const float* hits_t,
                           const uint8_t* bitfield, int cascades, int grid_size, float scale, float esf,
                           const float* noise, int max_samples, int32_t* n_samples) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int r = 0; r < n_rays; r++) {
        ray_t ray; load_ray(&ray, rays_o + 3 * r, rays_d + 3 * r);
        float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        if (t1 >= 0) { /* :192-195  t1 += dt*noise -> FFMA */
            const float dt = calc_dt(t1, esf, max_samples, grid_size, scale);
            t1 = fmaf(dt, noise[r], t1);
        }
        float t = t1; int N = 0;
        while (0 <= t && t < t2 && N < max_samples) { /* :204 */
            float x, y, z, dt;
            if (march_eval(&ray, bitfield, cascades, grid_size, scale, scale, esf, max_samples, &t, &x, &y, &z, &dt)) {
                t += dt; N++;
            }
        }
        n_samples[r] = N;
    }
}

/* raymarching.cu:166-280 pass 2 (:236-279).  The reference reserves [start,start+N) and a rays_a row with two
 * independent atomics (:237-238), so its row/segment order is scheduling-dependent.  The restatement uses the
 * canonical order rays_a[r] = (r, exclusive_scan(N)[r], N[r]); parity is per ray (tests sort reference rows).
 * Outputs must be sized by sum(n_samples) (from orc_march_train_count). Returns total samples. */
int64_t orc_march_train_emit(int n_rays, const float* rays_o, const float* rays_d, const float* hits_t,
                             const uint8_t* bitfield, int cascades, int grid_size, float scale, float esf,
                             const float* noise, int max_samples, const int32_t* n_samples,
                             int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts) {
    int64_t total = 0;
    for (int r = 0; r < n_rays; r++) {
        rays_a[3 * r] = r; rays_a[3 * r + 1] = total; rays_a[3 * r + 2] = n_samples[r];
        total += n_samples[r];
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int r = 0; r < n_rays; r++) {
        ray_t ray; load_ray(&ray, rays_o + 3 * r, rays_d + 3 * r);
        float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        if (t1 >= 0) {
            const float dt = calc_dt(t1, esf, max_samples, grid_size, scale);
            t1 = fmaf(dt, noise[r], t1);
        }
        const int64_t start = rays_a[3 * r + 1]; const int N = n_samples[r];
        float t = t1; int samples = 0;
        while (t < t2 && samples < N) { /* :245 */
            float x, y, z, dt;
            if (march_eval(&ray, bitfield, cascades, grid_size, scale, scale, esf, max_samples, &t, &x, &y, &z, &dt)) {
                const int64_t s = start + samples;
                xyzs[3 * s] = x; xyzs[3 * s + 1] = y; xyzs[3 * s + 2] = z;
                dirs[3 * s] = ray.dx; dirs[3 * s + 1] = ray.dy; dirs[3 * s + 2] = ray.dz;
                ts[s] = t; deltas[s] = dt;
                t += dt; samples++;
            }
        }
    }
    return total;
}

/* raymarching.cu:335-404.  Outputs (n_alive, S, .) must be zero-initialised by the caller (:421-426).
 * hits_t (R,2) is updated in place at [r][0] (:386). */
void orc_march_test(int n_alive, const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive,
                    const uint8_t* bitfield, int cascades, int grid_size, float scale, float esf, int N_samples,
                    int max_samples, float* xyzs, float* dirs, float* deltas, float* ts, int32_t* n_eff) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int n = 0; n < n_alive; n++) {
        const int64_t r = alive[n];
        ray_t ray; load_ray(&ray, rays_o + 3 * r, rays_d + 3 * r);
        float t = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        int s = 0;
        while (t < t2 && s < N_samples) {
            float x, y, z, dt;
            /* :370,399 calc_dt(..., cascades) -- `cascades` where `scale` belongs */
            if (march_eval(&ray, bitfield, cascades, grid_size, scale, (float)cascades, esf, max_samples, &t, &x, &y, &z, &dt)) {
                const size_t o = (size_t)n * N_samples + s;
                xyzs[3 * o] = x; xyzs[3 * o + 1] = y; xyzs[3 * o + 2] = z;
                dirs[3 * o] = ray.dx; dirs[3 * o + 1] = ray.dy; dirs[3 * o + 2] = ray.dz;
                ts[o] = t; deltas[o] = dt;
                t += dt;
                hits_t[2 * r] = t;
                s++;
            }
        }
        n_eff[n] = s;
    }
}

/* ------------------------------------------------------------------------------------------------
 * intersection.cu:5-22,25-56,59-100.  hits sorted near->far by t1 ascending with the -1 fill values,
 * exactly what torch::sort on hits_t[...,0] does (:95-97): unfilled (-1) slots sort FIRST.
 * With more hits than max_hits the reference keeps whichever won the atomic race; the restatement
 * keeps the first max_hits in voxel order (one legal outcome).
 * ---------------------------------------------------------------------------------------------- */
static void sort_hits(int max_hits, float* ht, int64_t* hi) {
    for (int i = 1; i < max_hits; i++) { /* stable insertion sort on t1 */
        const float a0 = ht[2 * i], a1 = ht[2 * i + 1]; const int64_t ai = hi[i];
        int j = i - 1;
        while (j >= 0 && ht[2 * j] > a0) { ht[2 * j + 2] = ht[2 * j]; ht[2 * j + 3] = ht[2 * j + 1]; hi[j + 1] = hi[j]; j--; }
        ht[2 * j + 2] = a0; ht[2 * j + 3] = a1; hi[j + 1] = ai;
    }
}

void orc_ray_aabb_intersect(int n_rays, const float* rays_o, const float* rays_d, int n_vox, const float* centers,
                            const float* half_sizes, int max_hits, int32_t* hit_cnt, float* hits_t, int64_t* hits_idx) {
#pragma omp parallel for
    for (int r = 0; r < n_rays; r++) {
        float* ht = hits_t + (size_t)r * max_hits * 2; int64_t* hi = hits_idx + (size_t)r * max_hits;
        for (int k = 0; k < max_hits; k++) { ht[2 * k] = -1.0f; ht[2 * k + 1] = -1.0f; hi[k] = -1; }
        int cnt = 0;
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]}; /* :41 */
        for (int v = 0; v < n_vox; v++) {
            const float* c = centers + 3 * v; const float* h = half_sizes + 3 * v;
            float a1[3], a2[3];
            for (int k = 0; k < 3; k++) { /* :12-16 (sub, sub, mul: nothing to contract) */
                const float tmin = ((c[k] - h[k]) - o[k]) * inv[k];
                const float tmax = ((c[k] + h[k]) - o[k]) * inv[k];
                a1[k] = fminf(tmin, tmax); a2[k] = fmaxf(tmin, tmax);
            }
            float t1 = fmaxf(fmaxf(a1[0], a1[1]), a1[2]);
            float t2 = fminf(fminf(a2[0], a2[1]), a2[2]);
            if (t1 > t2) { t1 = -1.0f; t2 = -1.0f; } /* :20 */
            if (t2 > 0) { /* :49-55 */
                if (cnt < max_hits) { ht[2 * cnt] = fmaxf(t1, 0.0f); ht[2 * cnt + 1] = t2; hi[cnt] = v; }
                cnt++;
            }
        }
        hit_cnt[r] = cnt;
        sort_hits(max_hits, ht, hi);
    }
}

/* intersection.cu:103-153,156-197.  Unused by the reference's callers; kept for API completeness.
 * dot() contraction order follows the SASS: fma(a.z,b.z, fma(a.x,b.x, a.y*b.y)). */
void orc_ray_sphere_intersect(int n_rays, const float* rays_o, const float* rays_d, int n_sph, const float* centers,
                              const float* radii, int max_hits, int32_t* hit_cnt, float* hits_t, int64_t* hits_idx) {
    for (int r = 0; r < n_rays; r++) {
        float* ht = hits_t + (size_t)r * max_hits * 2; int64_t* hi = hits_idx + (size_t)r * max_hits;
        for (int k = 0; k < max_hits; k++) { ht[2 * k] = -1.0f; ht[2 * k + 1] = -1.0f; hi[k] = -1; }
        int cnt = 0;
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        for (int s = 0; s < n_sph; s++) {
            const float* c = centers + 3 * s;
            const float co[3] = {o[0] - c[0], o[1] - c[1], o[2] - c[2]};
            const float a = fmaf(d[2], d[2], fmaf(d[0], d[0], d[1] * d[1]));
            const float half_b = fmaf(d[2], co[2], fmaf(d[0], co[0], d[1] * co[1]));
            const float cc = fmaf(-radii[s], radii[s], fmaf(co[2], co[2], fmaf(co[0], co[0], co[1] * co[1])));
            const float disc = fmaf(half_b, half_b, -(a * cc));
            float t1 = -1.0f, t2 = -1.0f;
            if (!(disc < 0)) { const float sq = sqrtf(disc); t1 = (-half_b - sq) / a; t2 = (-half_b + sq) / a; }
            if (t2 > 0) {
                if (cnt < max_hits) { ht[2 * cnt] = fmaxf(t1, 0.0f); ht[2 * cnt + 1] = t2; hi[cnt] = s; }
                cnt++;
            }
        }
        hit_cnt[r] = cnt;
        sort_hits(max_hits, ht, hi);
    }
}

/* ------------------------------------------------------------------------------------------------
 * volumerendering.cu.  __expf(x) = ex2.approx(x*log2e): SASS is FMUL s*d ; FMUL -1.44269502 ; MUFU.EX2.
 * MUFU.EX2 is an approximation, so these are tolerance-matched (1e-4 rel in the north star; the
 * oracle-vs-reference golden test holds 2e-6 abs on w).
 * ---------------------------------------------------------------------------------------------- */
static inline float alpha_of(float sigma, float delta) { return 1.0f - exp2f((sigma * delta) * -1.44269502f); }

/* volumerendering.cu:5-44.  opacity/depth/rgb/ws/total_samples must be zero-initialised (:56-60). */
void orc_composite_train_fw(int n_rays, const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                            const int64_t* rays_a, float T_threshold, int64_t* total_samples, float* opacity,
                            float* depth, float* rgb, float* ws) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int n = 0; n < n_rays; n++) {
        const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
        int samples = 0; float T = 1.0f;
        while (samples < N) {
            const int64_t s = start + samples;
            const float a = alpha_of(sigmas[s], deltas[s]);
            const float w = a * T;
            rgb[3 * ray_idx + 0] = fmaf(w, rgbs[3 * s + 0], rgb[3 * ray_idx + 0]);
            rgb[3 * ray_idx + 1] = fmaf(w, rgbs[3 * s + 1], rgb[3 * ray_idx + 1]);
            rgb[3 * ray_idx + 2] = fmaf(w, rgbs[3 * s + 2], rgb[3 * ray_idx + 2]);
            depth[ray_idx] = fmaf(w, ts[s], depth[ray_idx]);
            opacity[ray_idx] += w;
            ws[s] = w;
            T *= 1.0f - a;
            if (T <= T_threshold) break; /* :40 -- before samples++ (:41): terminated rays count one less */
            samples++;
        }
        total_samples[ray_idx] = samples;
    }
}

/* volumerendering.cu:86-150 (+ host op dL_dws*ws :174, in-kernel inclusive scan :118-121).
 * dL_dsigmas/dL_drgbs must be zero-initialised (:171-172). */
void orc_composite_train_bw(int n_rays, const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb,
                            const float* dL_dws, const float* sigmas, const float* rgbs, const float* ws,
                            const float* deltas, const float* ts, const int64_t* rays_a, const float* opacity,
                            const float* depth, const float* rgb, float T_threshold, float* dL_dsigmas, float* dL_drgbs) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int n = 0; n < n_rays; n++) {
        const int64_t ray_idx = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
        if (N <= 0) continue; /* reference reads index start-1 here (:122) but never uses it */
        float* scan = (float*)malloc(sizeof(float) * (size_t)N);
        float acc = 0.0f;
        for (int i = 0; i < N; i++) { acc += dL_dws[start + i] * ws[start + i]; scan[i] = acc; }
        const float scan_sum = scan[N - 1];
        int samples = 0;
        const float R = rgb[3 * ray_idx], G = rgb[3 * ray_idx + 1], B = rgb[3 * ray_idx + 2];
        const float O = opacity[ray_idx], D = depth[ray_idx];
        const float gR = dL_drgb[3 * ray_idx], gG = dL_drgb[3 * ray_idx + 1], gB = dL_drgb[3 * ray_idx + 2];
        float T = 1.0f, r = 0.0f, g = 0.0f, b = 0.0f, d = 0.0f;
        while (samples < N) {
            const int64_t s = start + samples;
            const float a = alpha_of(sigmas[s], deltas[s]);
            const float w = a * T;
            r = fmaf(w, rgbs[3 * s], r); g = fmaf(w, rgbs[3 * s + 1], g); b = fmaf(w, rgbs[3 * s + 2], b);
            d = fmaf(w, ts[s], d);
            T *= 1.0f - a;
            dL_drgbs[3 * s] = gR * w; dL_drgbs[3 * s + 1] = gG * w; dL_drgbs[3 * s + 2] = gB * w;
            dL_dsigmas[s] = deltas[s] * (
                gR * (rgbs[3 * s] * T - (R - r)) +
                gG * (rgbs[3 * s + 1] * T - (G - g)) +
                gB * (rgbs[3 * s + 2] * T - (B - b)) +
                dL_dopacity[ray_idx] * (1 - O) +
                dL_ddepth[ray_idx] * (ts[s] * T - (D - d)) +
                T * dL_dws[s] - (scan_sum - scan[samples]));
            if (T <= T_threshold) break;
            samples++;
        }
        free(scan);
    }
}

/* volumerendering.cu:204-248.  In place on opacity/depth/rgb and alive (-1 = dead). */
void orc_composite_test_fw(int n_alive, int S, const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                           int64_t* alive, float T_threshold, const int32_t* n_eff, float* opacity, float* depth, float* rgb) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int n = 0; n < n_alive; n++) {
        if (n_eff[n] == 0) { alive[n] = -1; continue; }
        const int64_t r = alive[n];
        int s = 0; float T = 1 - opacity[r];
        while (s < n_eff[n]) {
            const size_t o = (size_t)n * S + s;
            const float a = alpha_of(sigmas[o], deltas[o]);
            const float w = a * T;
            rgb[3 * r] = fmaf(w, rgbs[3 * o], rgb[3 * r]);
            rgb[3 * r + 1] = fmaf(w, rgbs[3 * o + 1], rgb[3 * r + 1]);
            rgb[3 * r + 2] = fmaf(w, rgbs[3 * o + 2], rgb[3 * r + 2]);
            depth[r] = fmaf(w, ts[o], depth[r]);
            opacity[r] += w;
            T *= 1.0f - a;
            if (T <= T_threshold) { alive[n] = -1; break; }
            s++;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * losses.cu:7-41,44-59,62-107 (forward) and :110-140,143-181 (backward).
 * ---------------------------------------------------------------------------------------------- */
void orc_distortion_fw(int n_rays, int64_t n, const float* ws, const float* deltas, const float* ts, const int64_t* rays_a,
                       float* loss, float* ws_inclusive_scan, float* wts_inclusive_scan) {
    memset(ws_inclusive_scan, 0, sizeof(float) * (size_t)n);
    memset(wts_inclusive_scan, 0, sizeof(float) * (size_t)n);
    for (int r = 0; r < n_rays; r++) {
        const int64_t ray_idx = rays_a[3 * r], start = rays_a[3 * r + 1]; const int N = (int)rays_a[3 * r + 2];
        float wi = 0.0f, wti = 0.0f, acc = 0.0f;
        for (int i = 0; i < N; i++) {
            const int64_t s = start + i;
            const float we = wi, wte = wti;        /* exclusive scans (:31-38) */
            const float wt = ws[s] * ts[s];        /* :70 */
            wi += ws[s]; wti += wt;                /* inclusive scans (:22-29) */
            ws_inclusive_scan[s] = wi; wts_inclusive_scan[s] = wti;
            /* :91-92 elementwise torch expression, then thrust::reduce (:54-57) */
            const float l = 2 * (wti * we - wi * wte) + ((1.0f / 3) * ws[s]) * ws[s] * deltas[s];
            acc += l;
        }
        loss[ray_idx] = acc;
    }
}

void orc_distortion_bw(int n_rays, const float* dL_dloss, const float* ws_inclusive_scan, const float* wts_inclusive_scan,
                       const float* ws, const float* deltas, const float* ts, const int64_t* rays_a, float* dL_dws) {
    for (int r = 0; r < n_rays; r++) {
        const int64_t ray_idx = rays_a[3 * r], start = rays_a[3 * r + 1]; const int N = (int)rays_a[3 * r + 2];
        if (N <= 0) continue;
        const int64_t end = start + N - 1;
        const float ws_sum = ws_inclusive_scan[end], wts_sum = wts_inclusive_scan[end];
        for (int64_t s = start; s <= end; s++) {
            float v = dL_dloss[ray_idx] * 2 * (
                (s == start ? 0.0f : (ts[s] * ws_inclusive_scan[s - 1] - wts_inclusive_scan[s - 1])) +
                (wts_sum - wts_inclusive_scan[s] - ts[s] * (ws_sum - ws_inclusive_scan[s])));
            v += dL_dloss[ray_idx] * (2.0f / 3) * ws[s] * deltas[s];
            dL_dws[s] = v;
        }
    }
}
