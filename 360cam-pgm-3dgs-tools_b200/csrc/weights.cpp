// Host-side construction of the interpolation weight tables.
//
// The reference resamples with cv2.remap (gs360_DualFisheyeDistortionCalibration.py:2001-2008);
// to land on the same pixel values the kernels use OpenCV's published weight scheme
// (modules/imgproc/src/imgwarp.cpp, initInterTab1D / initInterTab2D): 32 sub-pixel
// fractions, cubic kernel with A = -0.75 evaluated in float32, 2-D weight = float32 product,
// and for 8-bit images 15-bit fixed-point weights whose sum is forced to 32768.
//
// Compile with -ffp-contract=off: every operation must round to float32 on its own.
#include "r360_tables.h"

namespace r360 {

static void cubic_row(float t, float* w) {
    const float a = -0.75f;
    w[0] = ((a * (t + 1.0f) - 5.0f * a) * (t + 1.0f) + 8.0f * a) * (t + 1.0f) - 4.0f * a;
    w[1] = ((a + 2.0f) * t - (a + 3.0f)) * t * t + 1.0f;
    const float u = 1.0f - t;
    w[2] = ((a + 2.0f) * u - (a + 3.0f)) * u * u + 1.0f;
    w[3] = 1.0f - w[0] - w[1] - w[2];
}

static int round_half_even(float v) {
    // |v| <= 32768 here; nearbyintf honours the default round-to-nearest-even mode
    return (int)__builtin_nearbyintf(v);
}

void build_weight_tables(WeightTables* out) {
    const float step = 1.0f / 32.0f;
    for (int f = 0; f < 32; ++f) cubic_row((float)f * step, out->cubic_1d + 4 * f);

    for (int fy = 0; fy < 32; ++fy) {
        for (int fx = 0; fx < 32; ++fx) {
            short* e = out->cubic_fixed + (fy * 32 + fx) * 16;
            int q[16];
            int sum = 0;
            for (int ky = 0; ky < 4; ++ky) {
                for (int kx = 0; kx < 4; ++kx) {
                    const float w = out->cubic_1d[4 * fy + ky] * out->cubic_1d[4 * fx + kx];
                    int v = round_half_even(w * 32768.0f);
                    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
                    q[ky * 4 + kx] = v;
                    sum += v;
                }
            }
            const int excess = sum - 32768;
            if (excess != 0) {
                // OpenCV folds the rounding excess into the extreme weight among taps
                // (ky, kx) in {2, 3} x {2, 3}: largest if the sum is short, smallest if over.
                int hi = 2 * 4 + 2, lo = 2 * 4 + 2;
                for (int ky = 2; ky < 4; ++ky) {
                    for (int kx = 2; kx < 4; ++kx) {
                        const int idx = ky * 4 + kx;
                        if (q[idx] < q[lo]) lo = idx;
                        else if (q[idx] > q[hi]) hi = idx;
                    }
                }
                q[excess < 0 ? hi : lo] -= excess;
            }
            for (int k = 0; k < 16; ++k) e[k] = (short)q[k];
        }
    }
}

}  // namespace r360
