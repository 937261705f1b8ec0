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

extern "C" int h_frexp_exponent(float x) { return arn_frexp_exponent(x); }
