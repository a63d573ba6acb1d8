// Micro-probe: DFMA and double atan2 throughput per SM on the current GPU (design input for
// where the float64 projection may run).  nvcc -arch=sm_100a -O3 fp64_probe.cu -o fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f;
}
__global__ void ffma(float* out, int iters) {
    float a = threadIdx.x * 1e-3f, b = 1.0000001f, c = 1e-9f, d = a + 1, e = a + 2, f = a + 3;
    for (int i = 0; i < iters; ++i) { a = fmaf(a, b, c); d = fmaf(d, b, c); e = fmaf(e, b, c); f = fmaf(f, b, c); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f;
}
__global__ void datan2(double* out, int iters) {
    double x = 0.3 + threadIdx.x * 1e-3, y = 0.7, acc = 0;
    for (int i = 0; i < iters; ++i) { acc += atan2(x, y); x += 1e-3; y -= 1e-4; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double* d; cudaMalloc(&d, sms * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int it = 1 << 16;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dfma<<<sms * 8, 256>>>(d, it); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("DFMA: %.1f GDFMA/s  = %.1f lanes/clk/SM at %d MHz\n", 4.0 * it * sms * 8 * 256 / ms / 1e6,
                        4.0 * it * sms * 8 * 256 / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
        cudaEventRecord(e0); ffma<<<sms * 8, 256>>>((float*)d, it); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("FFMA: %.1f GFFMA/s  = %.1f lanes/clk/SM\n", 4.0 * it * sms * 8 * 256 / ms / 1e6,
                        4.0 * it * sms * 8 * 256 / (ms * 1e-3) / sms / (clk * 1e3));
        cudaEventRecord(e0); datan2<<<sms * 8, 256>>>(d, it / 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("atan2(double): %.2f G/s = %.3f per clk per SM\n", 1.0 * (it / 16) * sms * 8 * 256 / ms / 1e6,
                        1.0 * (it / 16) * sms * 8 * 256 / (ms * 1e-3) / sms / (clk * 1e3));
    }
    return 0;
}
