/*
 * oracle/oracle_field.c -- CPU restatement of the field evaluated by models/networks.py NGP
 * (reference call sites networks.py:37-57 hash grid + density MLP, :59-66 SH-4, :68-78 colour MLP,
 *  :95-108 density(), :133-165 forward(), custom_functions.py:162-173 TruncExp).
 *
 * TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).
 *
 * PARITY UNPINNED: the arithmetic lives in NVlabs/tiny-cuda-nn, which is neither vendored in
 * /root/reference nor version-pinned by it (README.md:39 only), and the reference has no test or golden
 * vector for any encoder/MLP output.  This file restates tiny-cuda-nn's published algorithm (multiresolution
 * hash encoding: Mueller et al. 2022, sec. 3; grid.h `grid_index`/`kernel_grid`; spherical_harmonics.h;
 * FullyFusedMLP weight layout) as recorded in SURVEY.md Appendix A, under ONE explicit numeric contract that
 * the CUDA product implements as well:
 *
 *   "fp16 operands, fp32 accumulate": the hash table, the MLP weights and every tensor that enters a matrix
 *   product (encoded features, SH coefficients, hidden activations, h as colour-net input, scaled output
 *   gradients) are rounded to IEEE fp16 (round-to-nearest-even); all sums are fp32; network outputs
 *   (h[16], colour pre-activations, sigma, rgb) and every gradient buffer are fp32.
 *   Backward scales the incoming output gradients by `loss_scale` (tiny-cuda-nn default 128) before the
 *   fp16 rounding and divides parameter/input gradients back at the end.
 *
 * An independent fp32-autograd PyTorch restatement (oracle/field_torch.py) cross-checks the hand-derived
 * backward here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 f16;

#define N_LEVELS 16
#define N_FEAT 2
#define ENC_DIM 32

/* Level geometry, SURVEY Appendix A.2 (tiny-cuda-nn grid.h: grid_scale / grid_resolution / offsets table),
 * strict float32 on the host.  tbl_* arrays have N_LEVELS entries (offsets N_LEVELS+1). */
void orc_hashgrid_geometry(int n_levels, int base_resolution, float per_level_scale, int log2_hashmap_size,
                           float* tbl_scale, uint32_t* tbl_res, uint32_t* tbl_size, uint32_t* tbl_offset) {
    const float log2_pls = log2f(per_level_scale);
    uint32_t offset = 0;
    for (int l = 0; l < n_levels; l++) {
        const float scale = exp2f((float)l * log2_pls) * (float)base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1;
        const uint32_t max_params = UINT32_MAX / 2;
        uint32_t params = max_params;
        if ((double)res * res * res < (double)max_params) params = res * res * res;
        params = (params + 7u) / 8u * 8u; /* next multiple of 8 */
        const uint32_t cap = 1u << log2_hashmap_size;
        if (params > cap) params = cap;
        tbl_scale[l] = scale; tbl_res[l] = res; tbl_size[l] = params; tbl_offset[l] = offset;
        offset += params;
    }
    tbl_offset[n_levels] = offset;
}

static inline uint32_t grid_index(uint32_t hashmap_size, uint32_t res, const uint32_t p[3]) {
    uint32_t stride = 1, index = 0;
    for (int d = 0; d < 3 && stride <= hashmap_size; d++) { index += p[d] * stride; stride *= res; }
    if (hashmap_size < stride) index = (p[0] * 1u) ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u);
    return index % hashmap_size;
}

/* Forward of the encoding (Appendix A.3).  x01 (N,3) in [0,1]; table fp16 (total_entries,2); out fp16 (N,32). */
void orc_hash_encode_fw(int64_t n, const float* x01, const float* tbl_scale, const uint32_t* tbl_res,
                        const uint32_t* tbl_size, const uint32_t* tbl_offset, const f16* table, f16* feat) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        for (int l = 0; l < N_LEVELS; l++) {
            float w[3]; uint32_t g[3];
            for (int d = 0; d < 3; d++) {
                const float pos = fmaf(tbl_scale[l], x01[3 * i + d], 0.5f);
                const float fl = floorf(pos);
                w[d] = pos - fl; g[d] = (uint32_t)(int32_t)fl;
            }
            float acc0 = 0.0f, acc1 = 0.0f;
            for (int c = 0; c < 8; c++) {
                uint32_t p[3]; float wt = 1.0f;
                for (int d = 0; d < 3; d++) {
                    if (c & (1 << d)) { wt *= w[d]; p[d] = g[d] + 1; } else { wt *= 1.0f - w[d]; p[d] = g[d]; }
                }
                const uint32_t idx = tbl_offset[l] + grid_index(tbl_size[l], tbl_res[l], p);
                acc0 = fmaf(wt, (float)table[2 * (size_t)idx], acc0);
                acc1 = fmaf(wt, (float)table[2 * (size_t)idx + 1], acc1);
            }
            feat[ENC_DIM * i + 2 * l] = (f16)acc0; feat[ENC_DIM * i + 2 * l + 1] = (f16)acc1;
        }
    }
}

/* Backward of the encoding: table_grad (total_entries,2) += weight * dfeat ; optional dL/dx01 (N,3).
 * The oracle accumulates in DOUBLE so that it is the order-independent reference for the (order-dependent) fp32
 * atomics of the GPU kernel. */
void orc_hash_encode_bw(int64_t n, const float* x01, const float* tbl_scale, const uint32_t* tbl_res,
                        const uint32_t* tbl_size, const uint32_t* tbl_offset, const f16* table, const float* dfeat,
                        double* table_grad, float* dx01) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float gx[3] = {0, 0, 0};
        for (int l = 0; l < N_LEVELS; l++) {
            float w[3]; uint32_t g[3];
            for (int d = 0; d < 3; d++) {
                const float pos = fmaf(tbl_scale[l], x01[3 * i + d], 0.5f);
                const float fl = floorf(pos);
                w[d] = pos - fl; g[d] = (uint32_t)(int32_t)fl;
            }
            const float d0 = dfeat[ENC_DIM * i + 2 * l], d1 = dfeat[ENC_DIM * i + 2 * l + 1];
            for (int c = 0; c < 8; c++) {
                uint32_t p[3]; float wt = 1.0f;
                for (int d = 0; d < 3; d++) {
                    if (c & (1 << d)) { wt *= w[d]; p[d] = g[d] + 1; } else { wt *= 1.0f - w[d]; p[d] = g[d]; }
                }
                const uint32_t idx = tbl_offset[l] + grid_index(tbl_size[l], tbl_res[l], p);
                if (table_grad) {
#pragma omp atomic
                    table_grad[2 * (size_t)idx] += (double)(wt * d0);
#pragma omp atomic
                    table_grad[2 * (size_t)idx + 1] += (double)(wt * d1);
                }
                if (dx01) {
                    const float v = (float)table[2 * (size_t)idx] * d0 + (float)table[2 * (size_t)idx + 1] * d1;
                    for (int d = 0; d < 3; d++) {
                        float wd = 1.0f;
                        for (int e = 0; e < 3; e++) if (e != d) wd *= (c & (1 << e)) ? w[e] : 1.0f - w[e];
                        gx[d] += ((c & (1 << d)) ? 1.0f : -1.0f) * wd * v * tbl_scale[l];
                    }
                }
            }
        }
        if (dx01) { dx01[3 * i] = gx[0]; dx01[3 * i + 1] = gx[1]; dx01[3 * i + 2] = gx[2]; }
    }
}

/* SH degree 4 of the normalised direction (Appendix A.4; networks.py:144-145: d/|d|, (d+1)/2, tcnn maps back 2u-1). */
void orc_sh4(int64_t n, const float* dirs, f16* out, int out_stride) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = ((dx / nrm + 1.0f) / 2.0f) * 2.0f - 1.0f;
        const float y = ((dy / nrm + 1.0f) / 2.0f) * 2.0f - 1.0f;
        const float z = ((dz / nrm + 1.0f) / 2.0f) * 2.0f - 1.0f;
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        f16* o = out + (size_t)out_stride * i;
        o[0] = (f16)(0.28209479177387814f);
        o[1] = (f16)(-0.48860251190291987f * y);
        o[2] = (f16)(0.48860251190291987f * z);
        o[3] = (f16)(-0.48860251190291987f * x);
        o[4] = (f16)(1.0925484305920792f * xy);
        o[5] = (f16)(-1.0925484305920792f * yz);
        o[6] = (f16)(0.94617469575755997f * z2 - 0.31539156525251999f);
        o[7] = (f16)(-1.0925484305920792f * xz);
        o[8] = (f16)(0.54627421529603959f * x2 - 0.54627421529603959f * y2);
        o[9] = (f16)(0.59004358992664352f * y * (-3.0f * x2 + y2));
        o[10] = (f16)(2.8906114426405538f * xy * z);
        o[11] = (f16)(0.45704579946446572f * y * (1.0f - 5.0f * z2));
        o[12] = (f16)(0.3731763325901154f * z * (5.0f * z2 - 3.0f));
        o[13] = (f16)(0.45704579946446572f * x * (1.0f - 5.0f * z2));
        o[14] = (f16)(1.4453057213202769f * z * (x2 - y2));
        o[15] = (f16)(0.59004358992664352f * x * (-x2 + 3.0f * y2));
    }
}

/* y[j] = sum_k W[j][k] * x[k], W fp16 row-major [n_out][n_in] (Appendix A.5), fp32 accumulate, k ascending. */
static inline void matvec(const f16* W, int n_out, int n_in, const f16* x, float* y) {
    for (int j = 0; j < n_out; j++) {
        float acc = 0.0f;
        for (int k = 0; k < n_in; k++) acc = fmaf((float)W[j * n_in + k], (float)x[k], acc);
        y[j] = acc;
    }
}
/* dx[k] = sum_j W[j][k] * g[j] */
static inline void matvec_t(const f16* W, int n_out, int n_in, const f16* g, float* dx) {
    for (int k = 0; k < n_in; k++) dx[k] = 0.0f;
    for (int j = 0; j < n_out; j++) {
        const float gj = (float)g[j];
        for (int k = 0; k < n_in; k++) dx[k] = fmaf((float)W[j * n_in + k], gj, dx[k]);
    }
}

/* Density net 32 -> 64 (ReLU) -> 16 (networks.py:37-57).  Wd fp16: W1[64][32] then W2[16][64].
 * Saves hid (N,64) fp16; h (N,16) fp32; sigma = exp(h0) (TruncExp fw, custom_functions.py:166-167). */
void orc_density_mlp_fw(int64_t n, const f16* feat, const f16* Wd, f16* hid, float* h, float* sigma) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float a[64];
        matvec(Wd, 64, 32, feat + 32 * i, a);
        f16* hi = hid + 64 * i;
        for (int j = 0; j < 64; j++) hi[j] = (f16)fmaxf(a[j], 0.0f);
        matvec(Wd + 64 * 32, 16, 64, hi, h + 16 * i);
        sigma[i] = expf(h[16 * i]);
    }
}

/* Colour net 32 -> 64 -> 64 (ReLU) -> 16, Sigmoid on the first 3 (networks.py:68-78,146).
 * Input = [sh16 | fp16(h16)].  Wc fp16: W1[64][32], W2[64][64], W3[16][64].
 * rgb_act: 1 = Sigmoid, 0 = None.  Saves in (N,32), hid1, hid2 (N,64) fp16, rgb (N,3) fp32. */
void orc_rgb_mlp_fw(int64_t n, const f16* sh, const float* h, const f16* Wc, int rgb_act,
                    f16* in32, f16* hid1, f16* hid2, float* rgb) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        f16* in = in32 + 32 * i;
        for (int j = 0; j < 16; j++) { in[j] = sh[16 * i + j]; in[16 + j] = (f16)h[16 * i + j]; }
        float a[64], o[16];
        matvec(Wc, 64, 32, in, a);
        f16* h1 = hid1 + 64 * i; for (int j = 0; j < 64; j++) h1[j] = (f16)fmaxf(a[j], 0.0f);
        matvec(Wc + 64 * 32, 64, 64, h1, a);
        f16* h2 = hid2 + 64 * i; for (int j = 0; j < 64; j++) h2[j] = (f16)fmaxf(a[j], 0.0f);
        matvec(Wc + 64 * 32 + 64 * 64, 16, 64, h2, o);
        for (int j = 0; j < 3; j++) rgb[3 * i + j] = rgb_act ? 1.0f / (1.0f + expf(-o[j])) : o[j];
    }
}

/* Backward of both nets.  Inputs dL_dsigma (N), dL_drgb (N,3) fp32 (unscaled).
 * Outputs: dWc (7168) double +=, dWd (3072) double += (both already divided by loss_scale),
 *          dfeat (N,32) fp32 = dL/dfeat (unscaled) for the encoding backward. */
/* |W|^T g for non-negative g (the all-paths magnitude propagation of field_mlp_bw_l1 below) */
static inline void matvec_t_abs(const f16* W, int n_out, int n_in, const float* g, float* dx) {
    for (int k = 0; k < n_in; k++) dx[k] = 0.0f;
    for (int j = 0; j < n_out; j++)
        for (int k = 0; k < n_in; k++) dx[k] += fabsf((float)W[j * n_in + k]) * g[j];
}

static void field_mlp_bw_impl(int64_t n, const float* dL_dsigma, const float* dL_drgb, const float* rgb, const float* h,
                              const f16* feat, const f16* hid, const f16* in32, const f16* hid1, const f16* hid2,
                              const f16* Wd, const f16* Wc, int rgb_act, float loss_scale,
                              double* dWd, double* dWc, float* dfeat) {
    const f16* Wc1 = Wc; const f16* Wc2 = Wc + 2048; const f16* Wc3 = Wc + 2048 + 4096;
    const f16* Wd1 = Wd; const f16* Wd2 = Wd + 2048;
    const float inv_scale = 1.0f / loss_scale;
#pragma omp parallel
    {
        double* lWc = (double*)calloc(7168, sizeof(double)); /* double: order-independent reference sums */
        double* lWd = (double*)calloc(3072, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            f16 g3[16], g2[64], g1[64], gh[16], gd[64];
            float t[64];
            /* colour output layer */
            for (int j = 0; j < 16; j++) {
                float g = 0.0f;
                if (j < 3) {
                    const float y = rgb[3 * i + j];
                    g = dL_drgb[3 * i + j] * (rgb_act ? y * (1.0f - y) : 1.0f);
                }
                g3[j] = (f16)(g * loss_scale);
            }
            for (int j = 0; j < 16; j++) for (int k = 0; k < 64; k++)
                lWc[2048 + 4096 + j * 64 + k] += (float)g3[j] * (float)hid2[64 * i + k];
            matvec_t(Wc3, 16, 64, g3, t);
            for (int k = 0; k < 64; k++) g2[k] = (f16)((float)hid2[64 * i + k] > 0.0f ? t[k] : 0.0f);
            for (int j = 0; j < 64; j++) for (int k = 0; k < 64; k++)
                lWc[2048 + j * 64 + k] += (float)g2[j] * (float)hid1[64 * i + k];
            matvec_t(Wc2, 64, 64, g2, t);
            for (int k = 0; k < 64; k++) g1[k] = (f16)((float)hid1[64 * i + k] > 0.0f ? t[k] : 0.0f);
            for (int j = 0; j < 64; j++) for (int k = 0; k < 32; k++)
                lWc[j * 32 + k] += (float)g1[j] * (float)in32[32 * i + k];
            matvec_t(Wc1, 64, 32, g1, t); /* t[16..31] = scaled dL/dh from the colour branch */
            /* density output layer: dL/dh0 += dL/dsigma * exp(clamp(h0,-15,15))  (custom_functions.py:170-173) */
            for (int j = 0; j < 16; j++) {
                float g = t[16 + j];
                if (j == 0) g += (dL_dsigma[i] * expf(fminf(fmaxf(h[16 * i], -15.0f), 15.0f))) * loss_scale;
                gh[j] = (f16)g;
            }
            for (int j = 0; j < 16; j++) for (int k = 0; k < 64; k++)
                lWd[2048 + j * 64 + k] += (float)gh[j] * (float)hid[64 * i + k];
            matvec_t(Wd2, 16, 64, gh, t);
            for (int k = 0; k < 64; k++) gd[k] = (f16)((float)hid[64 * i + k] > 0.0f ? t[k] : 0.0f);
            for (int j = 0; j < 64; j++) for (int k = 0; k < 32; k++)
                lWd[j * 32 + k] += (float)gd[j] * (float)feat[32 * i + k];
            matvec_t(Wd1, 64, 32, gd, t);
            for (int k = 0; k < 32; k++) dfeat[32 * i + k] = t[k] * inv_scale;
        }
#pragma omp critical
        {
            for (int k = 0; k < 7168; k++) dWc[k] += lWc[k] * (double)inv_scale;
            for (int k = 0; k < 3072; k++) dWd[k] += lWd[k] * (double)inv_scale;
        }
        free(lWc); free(lWd);
    }
}
void orc_field_mlp_bw(int64_t n, const float* dL_dsigma, const float* dL_drgb, const float* rgb, const float* h,
                      const f16* feat, const f16* hid, const f16* in32, const f16* hid1, const f16* hid2,
                      const f16* Wd, const f16* Wc, int rgb_act, float loss_scale,
                      double* dWd, double* dWc, float* dfeat) {
    field_mlp_bw_impl(n, dL_dsigma, dL_drgb, rgb, h, feat, hid, in32, hid1, hid2, Wd, Wc, rgb_act, loss_scale, dWd, dWc, dfeat);
}
/* The yardstick of the parity tests (tests/test_gpu_parity.py assert_sum): the SAME walk with magnitudes -- every weight
 * replaced by |W|, every upstream gradient by |g|, no fp16 rounding -- so that each output is the sum of the absolute values
 * of ALL products that reach it along all paths of the backward graph ("all-paths L1").  It bounds every intermediate
 * cancellation: a gradient vector that is itself a cancelling sum (dL/dfeat = W1^T g, |W1|^T|g| >> |W1^T g|) carries its own
 * uncertainty into the sums it feeds.  dWd / dWc receive the L1 of every weight gradient, dfeat the L1 of dL/dfeat. */
void orc_field_mlp_bw_l1(int64_t n, const float* dL_dsigma, const float* dL_drgb, const float* rgb, const float* h,
                         const f16* feat, const f16* hid, const f16* in32, const f16* hid1, const f16* hid2,
                         const f16* Wd, const f16* Wc, int rgb_act, float loss_scale,
                         double* dWd, double* dWc, float* dfeat) {
    const f16* Wc1 = Wc; const f16* Wc2 = Wc + 2048; const f16* Wc3 = Wc + 2048 + 4096;
    const f16* Wd1 = Wd; const f16* Wd2 = Wd + 2048;
    (void)loss_scale;
#pragma omp parallel
    {
        double* lWc = (double*)calloc(7168, sizeof(double));
        double* lWd = (double*)calloc(3072, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            float g3[16], g2[64], g1[64], gh[16], gd[64], t[64];
            for (int j = 0; j < 16; j++) {
                float g = 0.0f;
                if (j < 3) { const float y = rgb[3 * i + j]; g = fabsf(dL_drgb[3 * i + j] * (rgb_act ? y * (1.0f - y) : 1.0f)); }
                g3[j] = g;
            }
            for (int j = 0; j < 16; j++) for (int k = 0; k < 64; k++) lWc[2048 + 4096 + j * 64 + k] += g3[j] * fabsf((float)hid2[64 * i + k]);
            matvec_t_abs(Wc3, 16, 64, g3, t);
            for (int k = 0; k < 64; k++) g2[k] = (float)hid2[64 * i + k] > 0.0f ? t[k] : 0.0f;
            for (int j = 0; j < 64; j++) for (int k = 0; k < 64; k++) lWc[2048 + j * 64 + k] += g2[j] * fabsf((float)hid1[64 * i + k]);
            matvec_t_abs(Wc2, 64, 64, g2, t);
            for (int k = 0; k < 64; k++) g1[k] = (float)hid1[64 * i + k] > 0.0f ? t[k] : 0.0f;
            for (int j = 0; j < 64; j++) for (int k = 0; k < 32; k++) lWc[j * 32 + k] += g1[j] * fabsf((float)in32[32 * i + k]);
            matvec_t_abs(Wc1, 64, 32, g1, t);
            for (int j = 0; j < 16; j++) {
                float g = t[16 + j];
                if (j == 0) g += fabsf(dL_dsigma[i] * expf(fminf(fmaxf(h[16 * i], -15.0f), 15.0f)));
                gh[j] = g;
            }
            for (int j = 0; j < 16; j++) for (int k = 0; k < 64; k++) lWd[2048 + j * 64 + k] += gh[j] * fabsf((float)hid[64 * i + k]);
            matvec_t_abs(Wd2, 16, 64, gh, t);
            for (int k = 0; k < 64; k++) gd[k] = (float)hid[64 * i + k] > 0.0f ? t[k] : 0.0f;
            for (int j = 0; j < 64; j++) for (int k = 0; k < 32; k++) lWd[j * 32 + k] += gd[j] * fabsf((float)feat[32 * i + k]);
            matvec_t_abs(Wd1, 64, 32, gd, t);
            for (int k = 0; k < 32; k++) dfeat[32 * i + k] = t[k];
        }
#pragma omp critical
        {
            for (int k = 0; k < 7168; k++) dWc[k] += lWc[k];
            for (int k = 0; k < 3072; k++) dWd[k] += lWd[k];
        }
        free(lWc); free(lWd);
    }
}
/* fp32 -> fp16 parameter cast done every forward (Appendix A.5). */
void orc_cast_f16(int64_t n, const float* src, f16* dst) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) dst[i] = (f16)src[i];
}
