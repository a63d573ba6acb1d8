// cv2-compatible interpolation weight tables (plain C++ header, no CUDA types).
#pragma once

namespace r360 {

struct WeightTables {
    short cubic_fixed[32 * 32 * 16];   // [fy][fx][ky][kx], each entry sums to 32768
    float cubic_1d[32 * 4];            // [f][k], A = -0.75
    float lanczos_1d[32 * 8];          // [f][k], Lanczos a = 4
    short lanczos_fixed[32 * 32 * 64]; // [fy][fx][ky][kx], each entry sums to 32768
};

// Fills `out` on the host (weights.cpp).
void build_weight_tables(WeightTables* out);

}  // namespace r360
