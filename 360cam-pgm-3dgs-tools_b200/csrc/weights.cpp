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

// interpolateLanczos4: 8 taps of the a = 4 windowed sinc, sin(pi d)sin(pi d / 4) expanded with the
// angle-addition table cs[] so that only one sin/cos pair (in double) is needed; each coefficient
// is rounded to float32, then the row is normalised in float32.  A zero fraction is the exact
// unit impulse.
static void lanczos4_row(float t, float* w) {
    if (t < 1.1920929e-07f) {
        for (int k = 0; k < 8; ++k) w[k] = k == 3 ? 1.0f : 0.0f;
        return;
    }
    const double q = 0.70710678118654752440084436210485, pi = 3.1415926535897932384626433832795;
    const double cs[8][2] = {{1, 0}, {-q, -q}, {0, 1}, {q, -q}, {-1, 0}, {q, q}, {0, -1}, {-q, q}};
    const double a0 = -((double)t + 3.0) * pi * 0.25;
    const double s0 = __builtin_sin(a0), c0 = __builtin_cos(a0);
    float sum = 0.0f;
    for (int k = 0; k < 8; ++k) {
        const float d = t + 3.0f - (float)k;
        if (__builtin_fabsf(d) >= 1e-6f) {
            const double a = -(double)d * pi * 0.25;
            w[k] = (float)((cs[k][0] * s0 + cs[k][1] * c0) / (a * a));
        } else {
            w[k] = 1e30f;
        }
        sum += w[k];
    }
    const float inv = 1.0f / sum;
    for (int k = 0; k < 8; ++k) w[k] *= inv;
}

static int round_half_even(float v) {
    // |v| <= 32768 here; nearbyintf honours the default round-to-nearest-even mode
    return (int)__builtin_nearbyintf(v);
}

// 15-bit fixed-point K x K tables from the 1-D float rows (initInterTab2D).
static void fixed_tables(const float* rows, int K, short* out) {
    for (int fy = 0; fy < 32; ++fy) {
        for (int fx = 0; fx < 32; ++fx) {
            short* e = out + (fy * 32 + fx) * K * K;
            int q[64];
            int sum = 0;
            for (int ky = 0; ky < K; ++ky) {
                for (int kx = 0; kx < K; ++kx) {
                    const float w = rows[K * fy + ky] * rows[K * fx + kx];
                    int v = round_half_even(w * 32768.0f);
                    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
                    q[ky * K + kx] = v;
                    sum += v;
                }
            }
            const int excess = sum - 32768;
            if (excess != 0) {
                // OpenCV folds the rounding excess into the extreme weight among taps
                // (ky, kx) in {K/2, K/2+1}^2: largest if the sum is short, smallest if over.
                const int h = K / 2;
                int hi = h * K + h, lo = h * K + h;
                for (int ky = h; ky < h + 2; ++ky) {
                    for (int kx = h; kx < h + 2; ++kx) {
                        const int idx = ky * K + kx;
                        if (q[idx] < q[lo]) lo = idx;
                        else if (q[idx] > q[hi]) hi = idx;
                    }
                }
                q[excess < 0 ? hi : lo] -= excess;
            }
            for (int k = 0; k < K * K; ++k) e[k] = (short)q[k];
        }
    }
}

void build_weight_tables(WeightTables* out) {
    const float step = 1.0f / 32.0f;
    for (int f = 0; f < 32; ++f) {
        cubic_row((float)f * step, out->cubic_1d + 4 * f);
        lanczos4_row((float)f * step, out->lanczos_1d + 8 * f);
    }
    fixed_tables(out->cubic_1d, 4, out->cubic_fixed);
    fixed_tables(out->lanczos_1d, 8, out->lanczos_fixed);
}

}  // namespace r360
