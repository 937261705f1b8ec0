// Host build of the PRODUCT's marching arithmetic (ar_nerf_b200/csrc/arn_march_core.h) so its bit-exactness against
// the oracle can be checked without a GPU (tests/test_host_march.py).  Test artefact only: nothing ships from here.
#include <stdint.h>
#include "../ar_nerf_b200/csrc/arn_march_core.h"

extern "C" void h_march_train(int n_rays, const float* o, const float* d, const float* hits_t, const uint8_t* bits, int cascades,
                              int grid, float scale, float esf, const float* noise, int max_samples, int32_t* counts,
                              float* t_rec /* n_rays*max_samples */) {
    const ArnMarchConsts c = arn_march_consts(cascades, grid, scale, scale, esf, max_samples);
    for (int r = 0; r < n_rays; r++) {
        const ArnRay ray = arn_load_ray(o + 3 * r, d + 3 * r);
        float t = arn_jitter_start(c, hits_t[2 * r], noise[r]);
        const float t2 = hits_t[2 * r + 1];
        int N = 0;
        while (0 <= t && t < t2 && N < max_samples) {
            float x, y, z, dt;
            if (arn_march_eval(c, ray, bits, t, x, y, z, dt)) { t_rec[(size_t)r * max_samples + N] = t; t = ARN_ADD(t, dt); N++; }
        }
        counts[r] = N;
    }
}

extern "C" void h_march_test(int n_alive, const float* o, const float* d, float* hits_t, const int64_t* alive, const uint8_t* bits,
                             int cascades, int grid, float scale, float esf, int S, int max_samples, float* ts, float* deltas,
                             float* xyzs, int32_t* n_eff) {
    const ArnMarchConsts c = arn_march_consts(cascades, grid, scale, (float)cascades, esf, max_samples);
    for (int n = 0; n < n_alive; n++) {
        const int64_t r = alive[n];
        const ArnRay ray = arn_load_ray(o + 3 * r, d + 3 * r);
        float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
        int s = 0;
        while (t < t2 && s < S) {
            float x, y, z, dt;
            if (arn_march_eval(c, ray, bits, t, x, y, z, dt)) {
                const size_t q = (size_t)n * S + s;
                ts[q] = t; deltas[q] = dt; xyzs[3 * q] = x; xyzs[3 * q + 1] = y; xyzs[3 * q + 2] = z;
                t = ARN_ADD(t, dt); hits_t[2 * r] = t; s++;
            }
        }
        n_eff[n] = s;
    }
}

// march_test_all_kernel (arn_vren.cu): every ray marched once to its end, t of every occupied sample recorded sample-major
// (ts_all[s * n_rays + r]), at most `stride` per ray; dt of a sample is calc_dt(t).  fast = the compile-time shortcuts.
extern "C" void h_march_test_all(int n_rays, const float* o, const float* d, const float* hits_t, const uint8_t* bits, int cascades,
                                 int grid, float scale, float esf, int max_samples, int stride, int fast, float* ts_all, float* dt_all,
                                 int32_t* totals) {
    const ArnMarchConsts c = arn_march_consts(cascades, grid, scale, (float)cascades, esf, max_samples);
    for (int r = 0; r < n_rays; r++) {
        const ArnRay ray = arn_load_ray(o + 3 * r, d + 3 * r);
        float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
        int s = 0;
        while (t < t2 && s < stride) {
            float x, y, z, dt;
            const bool occ = fast ? arn_march_eval_t<true>(c, ray, bits, t, x, y, z, dt) : arn_march_eval_t<false>(c, ray, bits, t, x, y, z, dt);
            if (occ) {
                ts_all[(size_t)s * n_rays + r] = t; dt_all[(size_t)s * n_rays + r] = arn_calc_dt(c, t);
                t = ARN_ADD(t, dt); s++;
            }
        }
        totals[r] = s;
    }
}

extern "C" int h_frexp_exponent(float x) { return arn_frexp_exponent(x); }

// Lane-by-lane emulation of march_train_count_warp_kernel (arn_vren.cu): same window procedure, the warp primitives
// (ballot, shfl) replaced by loops over 32-entry arrays.
static inline int popc32(uint32_t v) { return __builtin_popcount(v); }
extern "C" void h_march_train_window(int n_rays, const float* o, const float* d, const float* hits_t, const uint8_t* bits, int cascades,
                                     int grid, float scale, float esf, const float* noise, int max_samples, int32_t* counts,
                                     float* t_rec /* n_rays*max_samples */) {
    const ArnMarchConsts c = arn_march_consts(cascades, grid, scale, scale, esf, max_samples);
    for (int r = 0; r < n_rays; r++) {
        const ArnRay ray = arn_load_ray(o + 3 * r, d + 3 * r);
        const float t1 = arn_jitter_start(c, hits_t[2 * r], noise[r]);
        const float t2 = hits_t[2 * r + 1];
        float* rec = t_rec + (size_t)r * max_samples;
        int N = 0;
        if (0 <= t1 && t1 < t2) {
            float pending = -INFINITY;
            float t[32];
            t[0] = t1;
            for (int j = 1; j < 32; j++) t[j] = ARN_ADD(t[j - 1], arn_calc_dt(c, t[j - 1]));
            const bool fast = cascades == 1 && grid <= 256;
            for (;;) {
                float tgt[32]; bool occ[32]; int R[32]; uint32_t M[32];
                uint32_t valid = 0, occm = 0; int s0 = 0;
                for (int l = 0; l < 32; l++) {
                    float x, y, z, dt;
                    occ[l] = fast ? arn_march_probe<true, true>(c, ray, bits, t[l], x, y, z, dt, tgt[l])
                                  : arn_march_probe<false, false>(c, ray, bits, t[l], x, y, z, dt, tgt[l]);
                    if (t[l] < t2) valid |= 1u << l;
                    if (occ[l]) occm |= 1u << l;
                    if (t[l] < pending) s0++;
                }
                for (int l = 0; l < 32; l++) {
                    int lo = l + 1, hi = 32;
                    for (int it = 0; it < 5; it++) {
                        const int mid = (lo + hi) >> 1;
                        const float tv = t[mid & 31];
                        if (lo < hi) { if (tv < tgt[l]) lo = mid + 1; else hi = mid; }
                    }
                    R[l] = occ[l] ? l + 1 : lo; M[l] = 1u << l;
                }
                for (int it = 0; it < 5; it++) {
                    uint32_t Mo[32]; int Ro[32];
                    for (int l = 0; l < 32; l++) { Mo[l] = M[R[l] & 31]; Ro[l] = R[R[l] & 31]; }
                    for (int l = 0; l < 32; l++) if (R[l] < 32) { M[l] |= Mo[l]; R[l] = Ro[l]; }
                }
                const uint32_t vis = s0 < 32 ? M[s0 & 31] : 0u;
                uint32_t emit = vis & occm & valid;
                const int rem = max_samples - N;
                bool done = valid != 0xffffffffu;
                if (popc32(emit) >= rem) {
                    done = true;
                    uint32_t e2 = 0;
                    for (int l = 0; l < 32; l++) if (((emit >> l) & 1u) && popc32(emit & ((1u << l) - 1u)) < rem) e2 |= 1u << l;
                    emit = e2;
                }
                for (int l = 0; l < 32; l++) if ((emit >> l) & 1u) rec[N + popc32(emit & ((1u << l) - 1u))] = t[l];
                N += popc32(emit);
                if (done) break;
                if (vis) { const int last = 31 - __builtin_clz(vis); pending = occ[last] ? -INFINITY : tgt[last]; }
                for (int l = 0; l < 32; l++)
                    for (int k = 0; k < 32; k++) t[l] = ARN_ADD(t[l], arn_calc_dt(c, t[l]));
            }
        }
        counts[r] = N;
    }
}


// Lane-by-lane emulation of march_test_warp_kernel (arn_vren.cu): the window procedure above with the test march's start
// (hits_t[r][0], no jitter), the iteration's sample budget S and the resume point behind the last sample taken.
extern "C" void h_march_test_window(int n_alive, const float* o, const float* d, float* hits_t, const int64_t* alive, const uint8_t* bits,
                                    int cascades, int grid, float scale, float esf, int S, int max_samples, float* ts, float* deltas,
                                    int32_t* n_eff) {
    const ArnMarchConsts c = arn_march_consts(cascades, grid, scale, (float)cascades, esf, max_samples);
    const bool fast = cascades == 1 && grid <= 256;
    for (int n = 0; n < n_alive; n++) {
        const int64_t r = alive[n];
        const ArnRay ray = arn_load_ray(o + 3 * r, d + 3 * r);
        const float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        int N = 0; float t_resume = t1; bool moved = false;
        if (t1 < t2) {
            float pending = -INFINITY;
            float t[32];
            t[0] = t1;
            for (int j = 1; j < 32; j++) t[j] = ARN_ADD(t[j - 1], arn_calc_dt(c, t[j - 1]));
            for (;;) {
                float tgt[32], dts[32]; bool occ[32]; int R[32]; uint32_t M[32];
                uint32_t valid = 0, occm = 0; int s0 = 0;
                for (int l = 0; l < 32; l++) {
                    float x, y, z;
                    occ[l] = fast ? arn_march_probe<true, true>(c, ray, bits, t[l], x, y, z, dts[l], tgt[l])
                                  : arn_march_probe<false, false>(c, ray, bits, t[l], x, y, z, dts[l], tgt[l]);
                    if (t[l] < t2) valid |= 1u << l;
                    if (occ[l]) occm |= 1u << l;
                    if (t[l] < pending) s0++;
                }
                for (int l = 0; l < 32; l++) {
                    int lo = l + 1, hi = 32;
                    for (int it = 0; it < 5; it++) {
                        const int mid = (lo + hi) >> 1;
                        const float tv = t[mid & 31];
                        if (lo < hi) { if (tv < tgt[l]) lo = mid + 1; else hi = mid; }
                    }
                    R[l] = occ[l] ? l + 1 : lo; M[l] = 1u << l;
                }
                for (int it = 0; it < 5; it++) {
                    uint32_t Mo[32]; int Ro[32];
                    for (int l = 0; l < 32; l++) { Mo[l] = M[R[l] & 31]; Ro[l] = R[R[l] & 31]; }
                    for (int l = 0; l < 32; l++) if (R[l] < 32) { M[l] |= Mo[l]; R[l] = Ro[l]; }
                }
                const uint32_t vis = s0 < 32 ? M[s0 & 31] : 0u;
                uint32_t emit = vis & occm & valid;
                const int rem = S - N;
                bool done = valid != 0xffffffffu;
                if (popc32(emit) >= rem) {
                    done = true;
                    uint32_t e2 = 0;
                    for (int l = 0; l < 32; l++) if (((emit >> l) & 1u) && popc32(emit & ((1u << l) - 1u)) < rem) e2 |= 1u << l;
                    emit = e2;
                }
                for (int l = 0; l < 32; l++) if ((emit >> l) & 1u) {
                    const size_t q = (size_t)n * S + N + popc32(emit & ((1u << l) - 1u));
                    ts[q] = t[l]; deltas[q] = dts[l];
                }
                if (emit) { const int last = 31 - __builtin_clz(emit); t_resume = ARN_ADD(t[last], dts[last]); moved = true; }
                N += popc32(emit);
                if (done) break;
                if (vis) { const int last = 31 - __builtin_clz(vis); pending = occ[last] ? -INFINITY : tgt[last]; }
                for (int l = 0; l < 32; l++)
                    for (int k = 0; k < 32; k++) t[l] = ARN_ADD(t[l], arn_calc_dt(c, t[l]));
            }
        }
        if (moved) hits_t[2 * r] = t_resume;
        n_eff[n] = N;
    }
}
