// What-if: bilinear sampling on the TEXTURE unit instead of shared memory + integer arithmetic.
//
// cv2.remap's bilinear arithmetic on 8-bit data is exact integer math on 1/32-pixel fractions:
//   out = (sum_taps (32-fx | fx) * (32-fy | fy) * p + 512) >> 10
// The texture unit filters with 8-bit fractions (1.8 fixed point), which hold k/32 exactly, so a texture fetch at a
// coordinate snapped to 1/32 px computes the same weighted sum -- if its internal precision and rounding allow.
// This probe answers the two questions that decide whether a texture path is worth building:
//   (1) accuracy: fraction of pixels whose texture result, rounded to 8 bits, equals / is within 1 LSB of the integer
//       formula, on uniform noise (worst case), for all 32 x 32 fraction pairs;
//   (2) throughput: a perspective-like 2:1 minifying gather (the 8K panorama -> 1600 px view geometry: neighbouring
//       output pixels ~2 source pixels apart, rows drifting) of RGBA8 texels from a pitch-linear 2-D texture, one
//       fetch per output pixel, against the 4 texels/clk/SM nominal rate -- plus the cost of the RGB -> RGBX
//       expansion pass a 3-channel source needs first (textures have 1, 2 or 4 channels).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tex_probe tools/tex_probe.cu && tools/tex_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int W = 7680, H = 3840, PAD = 8;          // panorama; PAD wrap columns each side (seam), rows clamp (poles)
constexpr int OUT = 1600, VIEWS = 12;

// RGB (3 bytes) -> RGBX texels with wrap padding: the expansion pass a texture path pays once per uploaded frame
__global__ void expand_kernel(const uint8_t* __restrict__ rgb, uchar4* __restrict__ rgbx, int pitch_px) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W + 2 * PAD) return;
    int sx = x - PAD;
    sx = sx < 0 ? sx + W : sx >= W ? sx - W : sx;
    const uint8_t* p = rgb + ((size_t)y * W + sx) * 3;
    rgbx[(size_t)y * pitch_px + x] = make_uchar4(p[0], p[1], p[2], 255);
}

// One fetch per output pixel.  The map is a cheap stand-in with the right access pattern: view v looks at a
// different part of the panorama, output pixels step ~2 source pixels in x and y with a slow sinusoidal drift, and
// coordinates are snapped to 1/32 px exactly as the remap kernels do.
__global__ void __launch_bounds__(256) gather_kernel(cudaTextureObject_t tex, uint8_t* __restrict__ out, int frames) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), j = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int v = blockIdx.z % VIEWS, f = blockIdx.z / VIEWS;
    if (i >= OUT || j >= OUT || f >= frames) return;
    const float cx = 400.0f + 560.0f * v, cy = 300.0f + 40.0f * (v % 3);
    float sx = cx + 1.95f * i + 0.07f * j + 6.0f * __sinf(0.004f * j);
    float sy = cy + 1.95f * j - 0.05f * i + 6.0f * __cosf(0.004f * i);
    sx = rintf(sx * 32.0f) * (1.0f / 32.0f);
    sy = rintf(sy * 32.0f) * (1.0f / 32.0f);
    if (sx >= W) sx -= W;
    const float4 t = tex2D<float4>(tex, sx + PAD + 0.5f, sy + 0.5f);          // unnormalised coordinates, texel centres at +0.5
    uint8_t* o = out + (((size_t)(f * VIEWS + v) * OUT + j) * OUT + i) * 3;
    o[0] = (uint8_t)__float2int_rn(t.x * 255.0f);
    o[1] = (uint8_t)__float2int_rn(t.y * 255.0f);
    o[2] = (uint8_t)__float2int_rn(t.z * 255.0f);
}

// Accuracy: every pixel of a 1024 x 1024 output samples at (ix + fx/32, iy + fy/32) with (fx, fy) = pixel index mod 32
__global__ void accuracy_kernel(cudaTextureObject_t tex, float4* __restrict__ res) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const float sx = 100.0f + 3.0f * i + (i & 31) * (1.0f / 32.0f), sy = 50.0f + 3.0f * j + (j & 31) * (1.0f / 32.0f);
    res[j * 1024 + i] = tex2D<float4>(tex, sx + PAD + 0.5f, sy + 0.5f);
}

int main() {
    const int pitch_px = ((W + 2 * PAD) * 4 + 511) / 512 * 512 / 4;             // pitch-linear textures want 512-byte rows here
    std::vector<uint8_t> host((size_t)W * H * 3);
    uint32_t s = 12345u;
    for (auto& b : host) { s = s * 1664525u + 1013904223u; b = (uint8_t)(s >> 24); }
    uint8_t* d_rgb; uchar4* d_rgbx; uint8_t* d_out; float4* d_res;
    const int frames = 4;
    CHECK(cudaMalloc(&d_rgb, host.size()));
    CHECK(cudaMalloc(&d_rgbx, (size_t)pitch_px * H * 4));
    CHECK(cudaMalloc(&d_out, (size_t)frames * VIEWS * OUT * OUT * 3));
    CHECK(cudaMalloc(&d_res, 1024 * 1024 * sizeof(float4)));
    CHECK(cudaMemcpy(d_rgb, host.data(), host.size(), cudaMemcpyHostToDevice));

    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    const dim3 eg((W + 2 * PAD + 255) / 256, H);
    expand_kernel<<<eg, 256>>>(d_rgb, d_rgbx, pitch_px);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    for (int k = 0; k < 10; ++k) expand_kernel<<<eg, 256>>>(d_rgb, d_rgbx, pitch_px);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms_expand = 0;
    CHECK(cudaEventElapsedTime(&ms_expand, e0, e1));
    ms_expand /= 10;

    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d_rgbx;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = W + 2 * PAD;
    rd.res.pitch2D.height = H;
    rd.res.pitch2D.pitchInBytes = (size_t)pitch_px * 4;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CHECK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));

    // ---- accuracy against the integer formula
    accuracy_kernel<<<dim3(4, 1024), 256>>>(tex, d_res);
    CHECK(cudaDeviceSynchronize());
    std::vector<float4> res(1024 * 1024);
    CHECK(cudaMemcpy(res.data(), d_res, res.size() * sizeof(float4), cudaMemcpyDeviceToHost));
    long long equal = 0, within1 = 0, total = 0;
    double worst = 0;
    for (int j = 0; j < 1024; ++j)
        for (int i = 0; i < 1024; ++i) {
            const int ix = 100 + 3 * i, iy = 50 + 3 * j, fx = i & 31, fy = j & 31;
            const float4 t = res[j * 1024 + i];
            const float got[3] = {t.x, t.y, t.z};
            for (int c = 0; c < 3; ++c) {
                auto px = [&](int x, int y) { return (int)host[((size_t)y * W + x) * 3 + c]; };
                const int acc = (32 - fx) * (32 - fy) * px(ix, iy) + fx * (32 - fy) * px(ix + 1, iy) +
                                (32 - fx) * fy * px(ix, iy + 1) + fx * fy * px(ix + 1, iy + 1);
                const int want = (acc + 512) >> 10;
                const int g = (int)lrintf(got[c] * 255.0f);
                const double err = fabs(got[c] * 255.0 - acc / 1024.0);
                if (err > worst) worst = err;
                equal += g == want; within1 += abs(g - want) <= 1; ++total;
            }
        }
    printf("{\"probe\": \"texture bilinear vs cv2 integer formula, noise, all 32x32 fractions\", \"samples\": %lld, \"equal\": %.6f, "
           "\"within_1_lsb\": %.6f, \"worst_abs_error_of_unrounded_value_lsb\": %.5f}\n", total, (double)equal / total,
           (double)within1 / total, worst);

    // ---- throughput
    const dim3 gg((OUT + 31) / 32, (OUT + 7) / 8, VIEWS * frames);
    for (int k = 0; k < 3; ++k) gather_kernel<<<gg, 256>>>(tex, d_out, frames);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    for (int k = 0; k < 10; ++k) gather_kernel<<<gg, 256>>>(tex, d_out, frames);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 10;
    const double pix = (double)frames * VIEWS * OUT * OUT;
    int sms = 0, khz = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    printf("{\"probe\": \"texture gather, 2:1 minification, RGBA8 pitch-linear, %d frames x %d views x %d^2\", \"ms\": %.4f, "
           "\"Gpix_per_s\": %.1f, \"fetches_per_clk_per_sm_at_max_clock\": %.3f, \"expand_rgb_to_rgbx_ms_per_frame\": %.4f, "
           "\"expand_GBps\": %.1f, \"Gpix_per_s_including_expansion\": %.1f}\n", frames, VIEWS, OUT, ms, pix / ms / 1e6,
           pix / (ms * 1e-3) / ((double)sms * khz * 1e3), ms_expand, ((double)W * H * 7) / ms_expand / 1e6,
           pix / (ms + frames * ms_expand) / 1e6);
    return 0;
}
