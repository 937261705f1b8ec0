// Ray-march arithmetic shared by the train/test march kernels (and by the host-side unit harness in tests/).
//
// Bit-exactness contract: sample counts, indices and sample values must equal those of the reference kernels
// (/root/reference/models/csrc/raymarching.cu:11-32,166-280,335-404) as nvcc 12.9 compiles them for sm_100a
// (-O2, default -fmad=true).  The reference lets nvcc pick the fused multiply-adds; here every rounding is explicit
// (ARN_FMA / ARN_MUL / ARN_ADD map to __fmaf_rn / __fmul_rn / __fadd_rn, which ptxas never re-associates or fuses),
// following the FFMA sites read from the reference's SASS ("FMA sites").
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ARN_HD __host__ __device__ __forceinline__
#else
#define ARN_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define ARN_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define ARN_MUL(a, b) __fmul_rn((a), (b))
#define ARN_ADD(a, b) __fadd_rn((a), (b))
#define ARN_DIV(a, b) __fdiv_rn((a), (b))
#else  // host build (tests only): compile with -ffp-contract=off
#define ARN_FMA(a, b, c) fmaf((a), (b), (c))
#define ARN_MUL(a, b) ((a) * (b))
#define ARN_ADD(a, b) ((a) + (b))
#define ARN_DIV(a, b) ((a) / (b))
#endif

#define ARN_SQRT3 1.73205080757f

// Loop-invariant quantities of one march launch.
struct ArnMarchConsts {
    int cascades;
    int grid_size;
    uint32_t grid_size3;
    int max_samples;
    float scale;    // mip_bound clamp (raymarching.cu:211)
    float esf;      // exp_step_factor
    float dt_lo;    // SQRT3 / max_samples              (IEEE division)
    float dt_hi;    // (dt_scale * (SQRT3*2)) / grid    (constant folded first, IEEE division)
    float G;        // (float)grid_size
    float Gm1;      // G - 1
    float Ginv;     // 1 / G (IEEE)
    float mip0_bound_inv;  // 1 / min(0.5, scale) (IEEE): mip_bound_inv of cascade 0
};

// dt_scale: `scale` for the train kernel, `(float)cascades` for the test kernel (reference quirk, raymarching.cu:370).
ARN_HD ArnMarchConsts arn_march_consts(int cascades, int grid_size, float scale, float dt_scale, float esf,
                                       int max_samples) {
    ArnMarchConsts c;
    c.cascades = cascades; c.grid_size = grid_size;
    c.grid_size3 = (uint32_t)grid_size * (uint32_t)grid_size * (uint32_t)grid_size;
    c.max_samples = max_samples; c.scale = scale; c.esf = esf;
    c.G = (float)grid_size;
    c.dt_lo = ARN_DIV(ARN_SQRT3, (float)max_samples);
    c.dt_hi = ARN_DIV(ARN_MUL(dt_scale, ARN_SQRT3 * 2), c.G);
    c.Gm1 = ARN_ADD(c.G, -1.0f);
    c.Ginv = ARN_DIV(1.0f, c.G);
    c.mip0_bound_inv = ARN_DIV(1.0f, fminf(0.5f, scale));
    return c;
}

// raymarching.cu:11-13  clamp(t*esf, lo, hi) = fmaxf(lo, fminf(t*esf, hi))
ARN_HD float arn_calc_dt(const ArnMarchConsts& c, float t) { return fmaxf(c.dt_lo, fminf(ARN_MUL(t, c.esf), c.dt_hi)); }

// Exponent e such that |x| = m * 2^e with m in [0.5,1) (frexpf), 0 for x == 0.  Finite inputs only.
ARN_HD int arn_frexp_exponent(float x) {
    union { float f; uint32_t u; } v; v.f = x;
    uint32_t bits = v.u & 0x7fffffffu;
    if (bits == 0) return 0;
    int adj = 0;
    if (bits < 0x00800000u) { v.f = ARN_MUL(fabsf(x), 16777216.0f); bits = v.u & 0x7fffffffu; adj = -24; }
    return (int)(bits >> 23) - 126 + adj;
}

ARN_HD uint32_t arn_expand_bits(uint32_t v) {  // raymarching.cu:35-42
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
ARN_HD uint32_t arn_morton3d(uint32_t x, uint32_t y, uint32_t z) {  // :44-50
    return arn_expand_bits(x) | (arn_expand_bits(y) << 1) | (arn_expand_bits(z) << 2);
}
// Same value for v < 256 (the first step of arn_expand_bits is then the identity): grids up to 256^3.
ARN_HD uint32_t arn_expand_bits8(uint32_t v) {
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
ARN_HD uint32_t arn_morton3d_8(uint32_t x, uint32_t y, uint32_t z) {
    return arn_expand_bits8(x) | (arn_expand_bits8(y) << 1) | (arn_expand_bits8(z) << 2);
}
ARN_HD uint32_t arn_morton3d_invert(uint32_t x) {  // :52-60
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

struct ArnRay {
    float ox, oy, oz, dx, dy, dz, dxi, dyi, dzi, sx, sy, sz;  // s* = copysignf(1, d*)
};

ARN_HD ArnRay arn_load_ray(const float* o, const float* d) {
    ArnRay r;
    r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
    r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
    r.dxi = ARN_DIV(1.0f, r.dx); r.dyi = ARN_DIV(1.0f, r.dy); r.dzi = ARN_DIV(1.0f, r.dz);  // :189
    r.sx = copysignf(1.0f, r.dx); r.sy = copysignf(1.0f, r.dy); r.sz = copysignf(1.0f, r.dz);
    return r;
}

// float -> int as F2I.TRUNC does it (NaN -> 0); inputs are already clamped to [0, G-1].
ARN_HD int arn_f2i(float f) { return (f != f) ? 0 : (int)f; }

// Jittered start (raymarching.cu:192-195): t1 += dt*noise  ->  one FFMA.
ARN_HD float arn_jitter_start(const ArnMarchConsts& c, float t1, float noise) {
    if (t1 >= 0) t1 = ARN_FMA(arn_calc_dt(c, t1), noise, t1);
    return t1;
}

// One probe of the loop body at parameter t (raymarching.cu:205-229): position, step, occupancy of the cell that holds
// o + t*d and -- for an empty cell -- the parameter t_target up to which the reference's skip loop advances.
// ONE_CASCADE / SMALL_GRID are compile-time shortcuts with identical results: cascades == 1 makes both mip clamps
// return 0 whatever the position, grid_size <= 256 makes the first morton step the identity.
#if defined(__cplusplus)
template <bool ONE_CASCADE = false, bool SMALL_GRID = false>
#endif
ARN_HD bool arn_march_probe(const ArnMarchConsts& c, const ArnRay& r, const uint8_t* __restrict__ bitfield, float t,
                            float& x, float& y, float& z, float& dt, float& t_target) {
    x = ARN_FMA(r.dx, t, r.ox); y = ARN_FMA(r.dy, t, r.oy); z = ARN_FMA(r.dz, t, r.oz);
    dt = arn_calc_dt(c, t);
    // mip_from_pos (:19-23) / mip_from_dt (:29-32)
    int mip = 0;
    if (!ONE_CASCADE) {
        const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        int mp = arn_frexp_exponent(mx) + 1; mp = mp < 0 ? 0 : mp; mp = mp > c.cascades - 1 ? c.cascades - 1 : mp;
        int md = arn_frexp_exponent(ARN_MUL(dt, c.G)); md = md < 0 ? 0 : md; md = md > c.cascades - 1 ? c.cascades - 1 : md;
        mip = mp > md ? mp : md;
    }
    // :211-212  mip_bound = min(2^(mip-1), scale) ; exact power of two built from its exponent bits
    union { uint32_t u; float f; } p2; p2.u = (uint32_t)(127 + mip - 1) << 23;
    const float mip_bound = fminf(p2.f, c.scale);
    const float mip_bound_inv = ONE_CASCADE ? c.mip0_bound_inv : ARN_DIV(1.0f, mip_bound);
    // :215-217
    const float fx = fmaxf(0.0f, fminf(ARN_MUL(ARN_MUL(0.5f, ARN_FMA(x, mip_bound_inv, 1.0f)), c.G), c.Gm1));
    const float fy = fmaxf(0.0f, fminf(ARN_MUL(ARN_MUL(0.5f, ARN_FMA(y, mip_bound_inv, 1.0f)), c.G), c.Gm1));
    const float fz = fmaxf(0.0f, fminf(ARN_MUL(ARN_MUL(0.5f, ARN_FMA(z, mip_bound_inv, 1.0f)), c.G), c.Gm1));
    const int nx = arn_f2i(fx), ny = arn_f2i(fy), nz = arn_f2i(fz);
    // :219-220
    const uint32_t idx = (uint32_t)mip * c.grid_size3 + (SMALL_GRID ? arn_morton3d_8((uint32_t)nx, (uint32_t)ny, (uint32_t)nz)
                                                                    : arn_morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const bool occ = (bitfield[idx >> 3] >> (idx & 7u)) & 1u;
    if (occ) { t_target = t; return true; }
    // :225-227
    const float tx = ARN_MUL(ARN_FMA(mip_bound, ARN_FMA(ARN_MUL(ARN_FMA(r.sx, 0.5f, ARN_ADD((float)nx, 0.5f)), c.Ginv), 2.0f, -1.0f), -x), r.dxi);
    const float ty = ARN_MUL(ARN_FMA(mip_bound, ARN_FMA(ARN_MUL(ARN_FMA(r.sy, 0.5f, ARN_ADD((float)ny, 0.5f)), c.Ginv), 2.0f, -1.0f), -y), r.dyi);
    const float tz = ARN_MUL(ARN_FMA(mip_bound, ARN_FMA(ARN_MUL(ARN_FMA(r.sz, 0.5f, ARN_ADD((float)nz, 0.5f)), c.Ginv), 2.0f, -1.0f), -z), r.dzi);
    // :229
    t_target = ARN_ADD(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    return false;
}

// One evaluation of the loop body at parameter t (raymarching.cu:205-232).
// Occupied: returns true, t unchanged (caller records the sample and does t += dt).
// Empty:    returns false, t advanced by the reference's do/while skip (:230-232).
ARN_HD bool arn_march_eval(const ArnMarchConsts& c, const ArnRay& r, const uint8_t* __restrict__ bitfield, float& t,
                           float& x, float& y, float& z, float& dt) {
    float t_target;
    if (arn_march_probe<false, false>(c, r, bitfield, t, x, y, z, dt, t_target)) return true;
    do { t = ARN_ADD(t, arn_calc_dt(c, t)); } while (t < t_target);
    return false;
}

#if defined(__cplusplus)
// The same evaluation with the compile-time shortcuts of arn_march_probe (identical results).
template <bool FAST>
ARN_HD bool arn_march_eval_t(const ArnMarchConsts& c, const ArnRay& r, const uint8_t* __restrict__ bitfield, float& t,
                             float& x, float& y, float& z, float& dt) {
    float t_target;
    if (arn_march_probe<FAST, FAST>(c, r, bitfield, t, x, y, z, dt, t_target)) return true;
    do { t = ARN_ADD(t, arn_calc_dt(c, t)); } while (t < t_target);
    return false;
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Window form of the train march (warp-cooperative kernel, arn_vren.cu).
//
// Every advance of the reference loop -- the occupied step `t += dt` and each turn of the skip loop -- is the same map
// t -> t + calc_dt(t), so the parameters the loop can ever visit form ONE chain t_0 = t1, t_{k+1} = t_k + calc_dt(t_k)
// that does not depend on the occupancy grid; the grid only selects WHICH chain points are visited:
//   visit k, occupied  -> sample, next visit k+1
//   visit k, empty     -> next visit = first k' > k with t_{k'} >= t_target(k)
// A warp therefore takes 32 consecutive chain points at a time (lane j builds t_{k0+j} by j sequential steps, so every
// value is the reference's own rounding sequence), probes all 32 cells at once, turns the rule above into a "next"
// pointer per lane (binary search over the monotone chain), finds the visited lanes by pointer doubling from the first
// lane, and compacts the occupied visited lanes with ballot/popc.  A skip that leaves the window is carried into the
// next window as `pending` (chain points below it are passed over).  Counts and sample parameters are bit-identical
// with the one-thread-per-ray loop; tests/test_host_march.py runs a lane-by-lane host emulation of exactly this
// procedure (tests/host_march_harness.cpp) against the oracle.
