// libarnerf.so -- spherical-Gaussian shadow and shading of an inserted object (BASELINE config 5).
//
// Replaces, for the per-pixel work of one AR frame,
//   insert/sg_shadow.py:103-116  SGShadow.calc_shadow_factor            (shadow the object casts on the scene: one factor per pixel)
//   insert/sg_shadow.py:118-153  SGShadow.calc_self_shadow_light_dacay  (per pixel, per light attenuation on the object itself)
//   insert/render_utils.py:321-375 SG_render_core                       (Cook-Torrance shading under the 32 SG lights)
// which the reference runs as ~25 torch ops with (px, 32, 7) temporaries.  Here: a one-block prologue builds the per-light
// tables (sg_shadow.py:35-53 light_axis_to_cood: environment-map fetch of the PCA components at the light axis; the f_h row
// of the light's sharpness; sg_shadow.py:70-74 calc_inte_L), then ONE kernel per call does everything per pixel: trilinear
// fetch of the 32 PCA coefficients (channel-last volume: a corner is 128 contiguous bytes), the 32x32 product with the
// per-light components out of shared memory, the f_h table fetch, and either the shadow factor or the whole SG shading
// sum.  fp32 throughout, operation order of the reference; torch.nn.functional.grid_sample's bilinear / border rules.
#include "arn_common.cuh"

namespace arn {

constexpr int kSgMaxLights = ARN_SG_MAX_LIGHTS, kSgMaxComp = ARN_SG_MAX_COMPONENTS;
constexpr float kPi = 3.14159265358979323846f;
// per-light record in the scratch (floats)
constexpr int kLMean = 0, kLLam = 1, kLCol = 2, kLAxis = 5, kLRow0 = 8, kLWy1 = 9, kLRow1Ok = 10, kLFhN = 11, kLRec = 12;

__device__ __forceinline__ float unnorm(float x, int size, bool align) {
    return align ? (x + 1.0f) / 2.0f * (float)(size - 1) : ((x + 1.0f) * (float)size - 1.0f) / 2.0f;
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// grid_sample(img (C,H,W), (gx, gy)), bilinear, border, align_corners = False, channel c
__device__ float env_fetch(const float* __restrict__ img, int H, int W, float gx, float gy) {
    const float ix = clampf(unnorm(gx, W, false), 0.0f, (float)(W - 1)), iy = clampf(unnorm(gy, H, false), 0.0f, (float)(H - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    float v = img[y0 * W + x0] * (wx0 * wy0);
    if (x0 + 1 < W) v += img[y0 * W + x0 + 1] * (wx1 * wy0);
    if (y0 + 1 < H) v += img[(y0 + 1) * W + x0] * (wx0 * wy1);
    if (x0 + 1 < W && y0 + 1 < H) v += img[(y0 + 1) * W + x0 + 1] * (wx1 * wy1);
    return v;
}

// scratch: n_lights x (C + kLRec) floats, then inte_L (3)
__global__ void sg_light_tables_kernel(arn_sg_tables_t tb, const float* __restrict__ lSGs_axis, const float* __restrict__ lSGs, int n_lights,
                                       float* __restrict__ scratch) {
    const int C = tb.C, rec = C + kLRec;
    for (int e = threadIdx.x; e < n_lights * (C + 1); e += blockDim.x) {   // sg_shadow.py:35-53
        const int l = e / (C + 1), c = e % (C + 1);
        const float phi = acosf(lSGs_axis[7 * l + 1]), theta = atan2f(lSGs_axis[7 * l + 2], lSGs_axis[7 * l]);
        const float gy = phi / kPi * 2.0f - 1.0f, gx = theta / kPi;
        if (c < C) scratch[l * rec + kLRec + c] = env_fetch(tb.components + (size_t)c * tb.envH * tb.envW, tb.envH, tb.envW, gx, gy);
        else scratch[l * rec + kLMean] = env_fetch(tb.mean, tb.envH, tb.envW, gx, gy);
    }
    for (int l = threadIdx.x; l < n_lights; l += blockDim.x) {
        float* r = scratch + l * rec;
        const float lam = lSGs[7 * l + 3];
        r[kLLam] = lam;
        for (int k = 0; k < 3; k++) { r[kLCol + k] = lSGs[7 * l + 4 + k]; r[kLAxis + k] = lSGs[7 * l + k]; }
        // row of the f_h table: logspace -1..4 -> -1..1 (sg_shadow.py:57-58), grid_sample's y rule
        const float gy = (log10f(fabsf(lam + 1e-6f)) - 1.5f) / 2.5f;
        const float iy = clampf(unnorm(gy, tb.fh_h, false), 0.0f, (float)(tb.fh_h - 1));
        const float fy = floorf(iy);
        r[kLRow0] = fy; r[kLWy1] = iy - fy; r[kLRow1Ok] = ((int)fy + 1 < tb.fh_h) ? 1.0f : 0.0f;
        r[kLFhN] = 2.0f * kPi / lam * (1.0f - expf(-1.0f * lam));                    // sg_shadow.py:142-143
    }
    __syncthreads();
    if (threadIdx.x < 3) {   // sg_shadow.py:70-74, summed in light order
        float s = 0.0f;
        for (int l = 0; l < n_lights; l++) {
            const float lam = lSGs[7 * l + 3];
            s += 2.0f * kPi * (lSGs[7 * l + 4 + threadIdx.x] / lam) * (1.0f - expf(-1.0f * lam));
        }
        scratch[n_lights * rec + threadIdx.x] = s;
    }
}

struct SgFrame { float pos[3]; float rot[9]; int has_rot; float scale; };

// sg_shadow.py:80-101: ssdf of one point against every light, clipped to [-pi/2, pi/2] (:111), left in shared memory at
// ssdf[l * blockDim.x + threadIdx.x].  The PCA coefficients are fetched 32 at a time (registers: static indices) and each
// chunk is folded into the per-light sums straight away, so any number of components up to ARN_SG_MAX_COMPONENTS works
// without spilling (the insertion tool runs 128: insert/main.py:107).
__device__ __forceinline__ void ssdf_of_point(const arn_sg_tables_t& tb, const SgFrame& fr, const float* __restrict__ pt,
                                              const float* __restrict__ lights, int n_lights, float* __restrict__ ssdf) {
    const int rec = tb.C + kLRec;
    float m[3] = {pt[0] - fr.pos[0], pt[1] - fr.pos[1], pt[2] - fr.pos[2]};
    if (fr.has_rot) {
        const float a = fr.rot[0] * m[0] + fr.rot[1] * m[1] + fr.rot[2] * m[2];
        const float b = fr.rot[3] * m[0] + fr.rot[4] * m[1] + fr.rot[5] * m[2];
        const float c = fr.rot[6] * m[0] + fr.rot[7] * m[1] + fr.rot[8] * m[2];
        m[0] = a; m[1] = b; m[2] = c;
    }
    float p[3];
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = m[k] / fr.scale / tb.vol_range;
    const float dis = fmaxf(sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]), 1.0f);
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = p[k] / dis;
    const float delta = (asinf(1.0f / tb.vol_range) - asinf(1.0f / (dis * tb.vol_range))) * tb.angle_decay_fac;
    // grid_sample 3-D: x -> W, y -> H, z -> D; bilinear, border, align_corners = True
    const int S[3] = {tb.W, tb.H, tb.D};
    int i0[3]; float w1[3]; bool ok1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float i = clampf(unnorm(p[k], S[k], true), 0.0f, (float)(S[k] - 1));
        const float f = floorf(i);
        i0[k] = (int)f; w1[k] = i - f; ok1[k] = i0[k] + 1 < S[k];
    }
    float cw[8]; size_t coff[8];
#pragma unroll
    for (int corner = 0; corner < 8; corner++) {
        const int dx = corner & 1, dy = (corner >> 1) & 1, dz = corner >> 2;
        const bool ok = !((dx && !ok1[0]) || (dy && !ok1[1]) || (dz && !ok1[2]));
        cw[corner] = ok ? (dx ? w1[0] : 1.0f - w1[0]) * (dy ? w1[1] : 1.0f - w1[1]) * (dz ? w1[2] : 1.0f - w1[2]) : 0.0f;
        coff[corner] = ok ? ((size_t)((i0[2] + dz) * tb.H + (i0[1] + dy)) * tb.W + (i0[0] + dx)) * tb.C : 0;
    }
    for (int l = 0; l < n_lights; l++) ssdf[l * blockDim.x + threadIdx.x] = 0.0f;
    for (int c0 = 0; c0 < tb.C; c0 += 32) {
        float pca[32];
#pragma unroll
        for (int c = 0; c < 32; c++) pca[c] = 0.0f;
#pragma unroll
        for (int corner = 0; corner < 8; corner++) {
            if (cw[corner] == 0.0f) continue;  // (outside the volume's last cell, or a zero weight: contributes nothing)
            const float4* src = reinterpret_cast<const float4*>(tb.coeff_cl + coff[corner] + c0);
            const float w = cw[corner];
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++) {
                if (c0 + 4 * c4 < tb.C) {
                    const float4 v = __ldg(src + c4);
                    pca[4 * c4] += v.x * w; pca[4 * c4 + 1] += v.y * w; pca[4 * c4 + 2] += v.z * w; pca[4 * c4 + 3] += v.w * w;
                }
            }
        }
        for (int l = 0; l < n_lights; l++) {
            // (records are 16-byte aligned: C and the header are multiples of 4 floats; one 128-bit broadcast load per 4 components)
            const float4* comp = reinterpret_cast<const float4*>(lights + l * rec + kLRec + c0);
            float acc = 0.0f;
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++) {
                if (c0 + 4 * c4 < tb.C) {
                    const float4 q = comp[c4];
                    acc += pca[4 * c4] * q.x; acc += pca[4 * c4 + 1] * q.y; acc += pca[4 * c4 + 2] * q.z; acc += pca[4 * c4 + 3] * q.w;
                }
            }
            ssdf[l * blockDim.x + threadIdx.x] += acc;
        }
    }
    for (int l = 0; l < n_lights; l++) {
        const float v = ssdf[l * blockDim.x + threadIdx.x] + lights[l * rec + kLMean] + delta;
        ssdf[l * blockDim.x + threadIdx.x] = clampf(v, -kPi / 2.0f, kPi / 2.0f);
    }
}

// f_h of (pixel, light) from the clipped ssdf: table fetch (sg_shadow.py:55-64)
__device__ __forceinline__ float fh_of(const arn_sg_tables_t& tb, const float* __restrict__ rec, float ssdf) {
    const float ix = clampf(unnorm(ssdf / (kPi / 2.0f), tb.fh_w, false), 0.0f, (float)(tb.fh_w - 1));
    const float fx = floorf(ix);
    const int x0 = (int)fx, y0 = (int)rec[kLRow0];
    const float wx1 = ix - fx, wx0 = 1.0f - wx1, wy1 = rec[kLWy1], wy0 = 1.0f - wy1;
    const float* row = tb.fh_tab + (size_t)y0 * tb.fh_w;
    const bool x1 = x0 + 1 < tb.fh_w, y1 = rec[kLRow1Ok] != 0.0f;
    float v = __ldg(row + x0) * (wx0 * wy0);
    if (x1) v += __ldg(row + x0 + 1) * (wx1 * wy0);
    if (y1) v += __ldg(row + tb.fh_w + x0) * (wx0 * wy1);
    if (x1 && y1) v += __ldg(row + tb.fh_w + x0 + 1) * (wx1 * wy1);
    return v;
}

__global__ void __launch_bounds__(128) sg_shadow_factor_kernel(arn_sg_tables_t tb, SgFrame fr, const float* __restrict__ scratch, int n_lights,
                                                               const float* __restrict__ pts, int64_t n, float* __restrict__ factor) {
    extern __shared__ float sm[];
    const int rec = tb.C + kLRec;
    for (int e = threadIdx.x; e < n_lights * rec + 3; e += blockDim.x) sm[e] = scratch[e];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* ssdf = sm + n_lights * rec + 4;
    ssdf_of_point(tb, fr, pts + 3 * i, sm, n_lights, ssdf);
    float col[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < n_lights; l++) {                       // fhs @ lcols (sg_shadow.py:65-67)
        const float* r = sm + l * rec;
        const float fh = fh_of(tb, r, ssdf[l * blockDim.x + threadIdx.x]);
        col[0] += fh * r[kLCol]; col[1] += fh * r[kLCol + 1]; col[2] += fh * r[kLCol + 2];
    }
    const float* inte_L = sm + n_lights * rec;
    float f[3];
#pragma unroll
    for (int k = 0; k < 3; k++) f[k] = clampf(fabsf(col[k] / inte_L[k]), 0.0f, 1.0f);
    const float lum = 0.2989f * f[0] + 0.5870f * f[1] + 0.1140f * f[2];
    factor[i] = powf(lum, tb.shadow_pow_fac);
}

struct Sg { float ax[3]; float lam; float col[3]; };

__device__ __forceinline__ Sg sg_product(const Sg& a, const Sg& b) {   // render_utils.py:266-278
    Sg o;
    const float lm = a.lam + b.lam;
    float um[3];
#pragma unroll
    for (int k = 0; k < 3; k++) um[k] = (a.lam * a.ax[k] + b.lam * b.ax[k]) / lm;
    const float ul = sqrtf(um[0] * um[0] + um[1] * um[1] + um[2] * um[2]);
    const float inv = 1.0f / ul;
#pragma unroll
    for (int k = 0; k < 3; k++) o.ax[k] = um[k] * inv;
    o.lam = lm * ul;
    const float e = expf(lm * (ul - 1.0f));
#pragma unroll
    for (int k = 0; k < 3; k++) o.col[k] = a.col[k] * b.col[k] * e;
    return o;
}
// render_utils.py:280-300: the scalar in front of the colour
__device__ __forceinline__ float sg_hemi(const Sg& s, const float n[3]) {
    const float cos_b = s.ax[0] * n[0] + s.ax[1] * n[1] + s.ax[2] * n[2];
    const float lam = fmaxf(s.lam, 1e-6f), il = 1.0f / lam;
    const float t = sqrtf(lam) * (1.6988f + 10.8438f * il) / (1.0f + 6.2201f * il + 10.2415f * il * il);
    const float inv_a = expf(-t);
    float sv;
    if (cos_b >= 0.0f) {
        const float inv_b = expf(-t * fmaxf(cos_b, 0.0f));
        sv = (1.0f - inv_a * inv_b) / (1.0f - inv_a + inv_b - inv_a * inv_b);
    } else {
        const float b = expf(t * fminf(cos_b, 0.0f));
        sv = (b - inv_a) / ((1.0f - inv_a) * (b + 1.0f));
    }
    const float A_b = 2.0f * kPi / lam * (expf(-lam) - expf(-2.0f * lam));
    const float A_u = 2.0f * kPi / lam * (1.0f - expf(-lam));
    return A_b * (1.0f - sv) + A_u * sv;
}
// render_utils.py:304-318 for one light: irr += Hemi(sg x cosSG) - 31.7003 Hemi(sg)
__device__ __forceinline__ void sg_irradiance_add(const Sg& s, const float n[3], float irr[3]) {
    Sg cosg; cosg.ax[0] = n[0]; cosg.ax[1] = n[1]; cosg.ax[2] = n[2]; cosg.lam = 0.0315f; cosg.col[0] = cosg.col[1] = cosg.col[2] = 32.7080f;
    const Sg p = sg_product(s, cosg);
    const float h1 = sg_hemi(p, n), h2 = sg_hemi(s, n);
#pragma unroll
    for (int k = 0; k < 3; k++) irr[k] += h1 * p.col[k] - 31.7003f * (h2 * s.col[k]);
}

// SG_render_core :322-330: view / normal / distribution SG of a pixel
struct ShadePix { float nn[3]; float ndv; float m2; Sg D; };
__device__ __forceinline__ void shade_begin(const float* __restrict__ normal, const float* __restrict__ vdirs, const float* __restrict__ rough,
                                            int64_t i, ShadePix& sp) {
    const float nl = sqrtf(normal[3 * i] * normal[3 * i] + normal[3 * i + 1] * normal[3 * i + 1] + normal[3 * i + 2] * normal[3 * i + 2]);
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { sp.nn[k] = normal[3 * i + k] / nl; v[k] = -vdirs[3 * i + k]; }
    sp.ndv = sp.nn[0] * v[0] + sp.nn[1] * v[1] + sp.nn[2] * v[2];
    const float rg = rough[i];
    sp.m2 = rg * rg;
#pragma unroll
    for (int k = 0; k < 3; k++) sp.D.ax[k] = sp.ndv * sp.nn[k] * 2.0f - v[k];                 // reflect_dir :191-192
    sp.D.lam = 2.0f / sp.m2 / (4.0f * fmaxf(sp.ndv, 1e-6f));                                  // pos_dot_eps :12-13
    sp.D.col[0] = sp.D.col[1] = sp.D.col[2] = 1.0f / (kPi * sp.m2);
}
// SG_render_core :348-375
__device__ __forceinline__ void shade_end(const ShadePix& sp, const float spec_irr[3], const float diff_irr[3], const float* __restrict__ albedo,
                                          const float* __restrict__ metal, const float* __restrict__ rough, int64_t i, int clamp01,
                                          float* __restrict__ radiance) {
    const float NdotV = fmaxf(sp.ndv, 0.0f), NdotL = NdotV;
    const float mt = metal[i], rg = rough[i];
    const float p5 = powf(1.0f - NdotV, 5.0f);
    const float sq = sp.m2 * fmaxf(1.0f / (NdotV * NdotV) - 1.0f, 0.0f);
    const float G = 1.0f / (0.5f * (sqrtf(1.0f + sq) - 1.0f) * 2.0f + 1.0f);           // GeometryBlender :68-72
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float al = albedo[3 * i + k];
        const float F0 = 0.04f * (1.0f - mt) + al * mt;                                    // get_F0 :51-52
        const float Fr = F0 + (1.0f - F0) * p5;                                            // fresnelSchlick :55-57
        const float Moi = Fr * G / (4.0f * NdotL * NdotV + 1e-6f);
        const float spec = Moi * fmaxf(spec_irr[k], 0.0f);
        const float diff = al / kPi * fmaxf(diff_irr[k], 0.0f);
        const float kS = F0 + (fmaxf(1.0f - rg, F0) - F0) * p5;                            // fresnelSchlickRoughness :59-61
        const float kD = (1.0f - kS) * (1.0f - mt);
        const float rad = kD * diff + spec;
        radiance[3 * i + k] = clamp01 ? clampf(rad, 0.0f, 1.0f) : fmaxf(rad, 0.0f);
    }
}

__global__ void __launch_bounds__(128) sg_shade_kernel(arn_sg_tables_t tb, SgFrame fr, const float* __restrict__ scratch, int n_lights,
                                                       const float* __restrict__ pts, int64_t n, const float* __restrict__ albedo,
                                                       const float* __restrict__ metal, const float* __restrict__ rough,
                                                       const float* __restrict__ normal, const float* __restrict__ vdirs, int clamp01,
                                                       int self_shadow, int shade, float* __restrict__ lSGs_out, float* __restrict__ radiance) {
    extern __shared__ float sm[];
    const int rec = tb.C + kLRec;
    for (int e = threadIdx.x; e < n_lights * rec + 3; e += blockDim.x) sm[e] = scratch[e];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* ssdf = sm + n_lights * rec + 4;
    if (self_shadow) ssdf_of_point(tb, fr, pts + 3 * i, sm, n_lights, ssdf);
    ShadePix sp;
    if (shade) shade_begin(normal, vdirs, rough, i, sp);
    float spec_irr[3] = {0.f, 0.f, 0.f}, diff_irr[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < n_lights; l++) {
        const float* r = sm + l * rec;
        Sg L; L.lam = r[kLLam];
#pragma unroll
        for (int k = 0; k < 3; k++) { L.ax[k] = r[kLAxis + k]; L.col[k] = r[kLCol + k]; }
        if (self_shadow) {   // sg_shadow.py:131-152
            const float fh = fh_of(tb, r, ssdf[l * blockDim.x + threadIdx.x]);
            const float decay = powf(clampf(fabsf(fh / r[kLFhN]), 0.0f, 1.0f), tb.self_shadow_pow_fac);
#pragma unroll
            for (int k = 0; k < 3; k++) L.col[k] *= decay;
            if (lSGs_out) {
                float* o = lSGs_out + ((size_t)i * n_lights + l) * 7;
                o[0] = L.ax[0]; o[1] = L.ax[1]; o[2] = L.ax[2]; o[3] = L.lam; o[4] = L.col[0]; o[5] = L.col[1]; o[6] = L.col[2];
            }
        }
        if (shade) {
            sg_irradiance_add(sg_product(sp.D, L), sp.nn, spec_irr);
            sg_irradiance_add(L, sp.nn, diff_irr);
        }
    }
    if (shade) shade_end(sp, spec_irr, diff_irr, albedo, metal, rough, i, clamp01, radiance);
}

// SG_render_core on lights that are an INPUT per pixel ((n, L, 7): already attenuated, self_shadow=True) or shared ((L, 7))
__global__ void __launch_bounds__(128) sg_shade_px_kernel(const float* __restrict__ lSGs, int n_lights, int per_pixel, int64_t n,
                                                          const float* __restrict__ albedo, const float* __restrict__ metal,
                                                          const float* __restrict__ rough, const float* __restrict__ normal,
                                                          const float* __restrict__ vdirs, int clamp01, float* __restrict__ radiance) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ShadePix sp;
    shade_begin(normal, vdirs, rough, i, sp);
    float spec_irr[3] = {0.f, 0.f, 0.f}, diff_irr[3] = {0.f, 0.f, 0.f};
    const float* base = lSGs + (per_pixel ? (size_t)i * n_lights * 7 : 0);
    for (int l = 0; l < n_lights; l++) {
        const float* r = base + 7 * l;
        Sg L; L.ax[0] = r[0]; L.ax[1] = r[1]; L.ax[2] = r[2]; L.lam = r[3]; L.col[0] = r[4]; L.col[1] = r[5]; L.col[2] = r[6];
        sg_irradiance_add(sg_product(sp.D, L), sp.nn, spec_irr);
        sg_irradiance_add(L, sp.nn, diff_irr);
    }
    shade_end(sp, spec_irr, diff_irr, albedo, metal, rough, i, clamp01, radiance);
}

// per-light tables + inte_L (padded to 4 floats) + one ssdf per (light, thread of the 128-thread block)
static size_t sg_smem_bytes(int n_lights, int C) { return ((size_t)n_lights * (C + kLRec) + 4 + (size_t)n_lights * 128) * sizeof(float); }
// ---- Shadow field, the SH alternative to the SG shadow (insert/shadow_fields.py:59-78 soft_shadow_map): K SH coefficients of the
// object's visibility fetched trilinearly at the point (grid_sample, border, align_corners = True; channel-last volume: a
// corner is K contiguous floats), SH_product0 against the lighting's SH per colour, ratio to the unshadowed DC term, pow 10.
struct SfLight { float sh[ARN_SF_MAX_COEFFS * 3]; };  // model_sh9[k][c]
__global__ void __launch_bounds__(256) sf_soft_shadow_kernel(const float* __restrict__ sf_cl, int D, int H, int W, int K, float vol_range, SgFrame fr,
                                                             SfLight ml, const float* __restrict__ pts, int64_t n, float* __restrict__ sh_out,
                                                             float* __restrict__ shadow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m[3] = {pts[3 * i] - fr.pos[0], pts[3 * i + 1] - fr.pos[1], pts[3 * i + 2] - fr.pos[2]};
    if (fr.has_rot) {
        const float a = fr.rot[0] * m[0] + fr.rot[1] * m[1] + fr.rot[2] * m[2];
        const float b = fr.rot[3] * m[0] + fr.rot[4] * m[1] + fr.rot[5] * m[2];
        const float c = fr.rot[6] * m[0] + fr.rot[7] * m[1] + fr.rot[8] * m[2];
        m[0] = a; m[1] = b; m[2] = c;
    }
    const int S[3] = {W, H, D};
    int i0[3]; float w1[3]; bool ok1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float p = m[k] / fr.scale / vol_range;                      // shadow_fields.py:93 (no normalisation beyond the volume)
        const float x = clampf(unnorm(p, S[k], true), 0.0f, (float)(S[k] - 1));
        const float f = floorf(x);
        i0[k] = (int)f; w1[k] = x - f; ok1[k] = i0[k] + 1 < S[k];
    }
    float sh[ARN_SF_MAX_COEFFS];
#pragma unroll
    for (int k = 0; k < ARN_SF_MAX_COEFFS; k++) sh[k] = 0.0f;
#pragma unroll
    for (int corner = 0; corner < 8; corner++) {
        const int dx = corner & 1, dy = (corner >> 1) & 1, dz = corner >> 2;
        if ((dx && !ok1[0]) || (dy && !ok1[1]) || (dz && !ok1[2])) continue;
        const float w = (dx ? w1[0] : 1.0f - w1[0]) * (dy ? w1[1] : 1.0f - w1[1]) * (dz ? w1[2] : 1.0f - w1[2]);
        const float* src = sf_cl + ((size_t)((i0[2] + dz) * H + (i0[1] + dy)) * W + (i0[0] + dx)) * K;
#pragma unroll
        for (int k = 0; k < ARN_SF_MAX_COEFFS; k++) if (k < K) sh[k] += __ldg(src + k) * w;
    }
    if (sh_out) {
#pragma unroll
        for (int k = 0; k < ARN_SF_MAX_COEFFS; k++) if (k < K) sh_out[i * K + k] = sh[k];
    }
    if (!shadow) return;
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float dot = 0.0f;
#pragma unroll
        for (int k = 0; k < ARN_SF_MAX_COEFFS; k++) if (k < K) dot += sh[k] * ml.sh[3 * k + c];
        acc += clampf(0.2821f * dot / ml.sh[c], 0.0f, 1.0f);               // SH_product0 (insert_utils.py:153-154) over the unshadowed DC term
    }
    shadow[i] = powf(acc / 3.0f, 10.0f);                                   // mean over the colours, "to augment shadow effect" (:76)
}

static int check_tables(const arn_sg_tables_t* tb, int n_lights) {
    if (!tb || !tb->coeff_cl || !tb->components || !tb->mean || !tb->fh_tab) { set_error("arn_sg: null table pointer"); return ARN_E_INVALID; }
    if (tb->C < 4 || tb->C > kSgMaxComp || tb->C % 4 || n_lights < 1 || n_lights > kSgMaxLights || tb->D < 1 || tb->H < 1 || tb->W < 1 ||
        tb->envH < 1 || tb->envW < 1 || tb->fh_h < 1 || tb->fh_w < 1 || !(tb->vol_range > 0.0f)) {
        set_error("arn_sg: unsupported table geometry (components: multiple of 4 up to %d, lights up to %d)", kSgMaxComp, kSgMaxLights);
        return ARN_E_INVALID;
    }
    if (((uintptr_t)tb->coeff_cl & 15) != 0) { set_error("arn_sg: coeff_cl must be 16-byte aligned"); return ARN_E_INVALID; }
    return ARN_OK;
}
static SgFrame make_frame(const float* pos, const float* rot, float scale) {
    SgFrame f{};
    for (int k = 0; k < 3; k++) f.pos[k] = pos ? pos[k] : 0.0f;
    f.has_rot = rot != nullptr;
    for (int k = 0; k < 9; k++) f.rot[k] = rot ? rot[k] : 0.0f;
    f.scale = scale;
    return f;
}

}  // namespace arn

using namespace arn;

extern "C" ARN_API int arn_sg_shadow_factor(const arn_sg_tables_t* tables_host, const float* lSGs, int n_lights, const float* pts, int64_t n,
                                            const float* model_pos_host, const float* rot_inv_host, float scale, float* light_scratch,
                                            float* factor, arn_stream_t stream) {
    if (int e = check_tables(tables_host, n_lights)) return e;
    ARN_REQUIRE(n >= 0 && scale > 0.0f, "bad size / scale");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(lSGs && pts && model_pos_host && light_scratch && factor, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const arn_sg_tables_t tb = *tables_host;
    ARN_LAUNCH("sg_light_tables_kernel", st, sg_light_tables_kernel<<<1, 256, 0, st>>>(tb, lSGs, lSGs, n_lights, light_scratch));
    if (int e = check_launch("sg_light_tables")) return e;
    const size_t smem = sg_smem_bytes(n_lights, tb.C);
    ARN_CUDA(cudaFuncSetAttribute(sg_shadow_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem_bytes(kSgMaxLights, kSgMaxComp)));
    ARN_LAUNCH("sg_shadow_factor_kernel", st, sg_shadow_factor_kernel<<<ceil_div(n, 128), 128, smem, st>>>(tb, make_frame(model_pos_host, rot_inv_host, scale), light_scratch,
                                                                                                          n_lights, pts, n, factor));
    return check_launch("sg_shadow_factor");
}

extern "C" ARN_API int arn_sg_shade(const arn_sg_tables_t* tables_host, const float* lSGs, const float* lSGs_axis, int n_lights, const float* pts,
                                    int64_t n, const float* model_pos_host, const float* rot_inv_host, float scale, const float* albedo,
                                    const float* metal, const float* rough, const float* normal, const float* vdirs, int clamp01, int self_shadow,
                                    float* light_scratch, float* lSGs_out, float* radiance, arn_stream_t stream) {
    if (int e = check_tables(tables_host, n_lights)) return e;
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    const bool shade = radiance != nullptr;
    ARN_REQUIRE(lSGs && light_scratch && (shade || lSGs_out), "null pointer");
    if (shade) ARN_REQUIRE(albedo && metal && rough && normal && vdirs, "null pointer (G-buffer)");
    if (self_shadow) ARN_REQUIRE(pts && model_pos_host && scale > 0.0f, "self shadow needs the points and the model frame");
    ARN_REQUIRE(self_shadow || !lSGs_out, "lSGs_out is the self-shadow attenuation: needs self_shadow");
    cudaStream_t st = (cudaStream_t)stream;
    const arn_sg_tables_t tb = *tables_host;
    ARN_LAUNCH("sg_light_tables_kernel", st, sg_light_tables_kernel<<<1, 256, 0, st>>>(tb, lSGs_axis ? lSGs_axis : lSGs, lSGs, n_lights, light_scratch));
    if (int e = check_launch("sg_light_tables")) return e;
    const size_t smem = sg_smem_bytes(n_lights, tb.C);
    ARN_CUDA(cudaFuncSetAttribute(sg_shade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem_bytes(kSgMaxLights, kSgMaxComp)));
    ARN_LAUNCH("sg_shade_kernel", st, sg_shade_kernel<<<ceil_div(n, 128), 128, smem, st>>>(tb, make_frame(model_pos_host, rot_inv_host, scale > 0.0f ? scale : 1.0f),
                                                                                         light_scratch, n_lights, pts, n, albedo, metal, rough, normal, vdirs,
                                                                                         clamp01, self_shadow, shade ? 1 : 0, lSGs_out, radiance));
    return check_launch("sg_shade");
}

extern "C" ARN_API int arn_sg_shade_px(const float* lSGs, int n_lights, int per_pixel, int64_t n, const float* albedo, const float* metal,
                                       const float* rough, const float* normal, const float* vdirs, int clamp01, float* radiance,
                                       arn_stream_t stream) {
    ARN_REQUIRE(n >= 0 && n_lights >= 1, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(lSGs && albedo && metal && rough && normal && vdirs && radiance, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    ARN_LAUNCH("sg_shade_px_kernel", st, sg_shade_px_kernel<<<ceil_div(n, 128), 128, 0, st>>>(lSGs, n_lights, per_pixel, n, albedo, metal, rough, normal, vdirs,
                                                                                             clamp01, radiance));
    return check_launch("sg_shade_px");
}

extern "C" ARN_API int arn_sf_soft_shadow(const float* sf_cl, int D, int H, int W, int K, float vol_range, const float* model_sh_host, const float* pts,
                                          int64_t n, const float* model_pos_host, const float* rot_inv_host, float scale, float* sh_out, float* shadow,
                                          arn_stream_t stream) {
    ARN_REQUIRE(n >= 0 && D >= 1 && H >= 1 && W >= 1 && K >= 1 && K <= ARN_SF_MAX_COEFFS && vol_range > 0.0f && scale > 0.0f, "bad sizes");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(sf_cl && pts && model_pos_host && (sh_out || shadow) && (!shadow || model_sh_host), "null pointer");
    SfLight ml{};
    if (model_sh_host) for (int k = 0; k < 3 * K; k++) ml.sh[k] = model_sh_host[k];
    cudaStream_t st = (cudaStream_t)stream;
    ARN_LAUNCH("sf_soft_shadow_kernel", st, sf_soft_shadow_kernel<<<ceil_div(n, 256), 256, 0, st>>>(sf_cl, D, H, W, K, vol_range, make_frame(model_pos_host, rot_inv_host, scale),
                                                                                                   ml, pts, n, sh_out, shadow));
    return check_launch("sf_soft_shadow");
}
