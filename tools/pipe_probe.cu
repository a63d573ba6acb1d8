// Micro-benchmarks of the SM pipes the remap kernels lean on (B200, sm_100a): how many warp-instructions per clock
// per SM the machine sustains for the integer dot products, byte permutes, funnel shifts and shared-memory loads of
// the sampling loops, alone and in the mixes the kernels use.  The numbers feed the cycle models in DESIGN.md.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu && tools/pipe_probe
//
// Each kernel runs ITER iterations of an unrolled body of independent chains (8 per thread) in 1024-thread blocks,
// one block per SM x 2; throughput = executed warp-instructions / (SM cycles), cycles from clock64 on one SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITER = 2000;
constexpr int CH = 8;     // independent chains per thread

enum Op { DP2A, DP4A, PRMT, SHF, IMAD, LOP3, IADD3, FFMA, SEL, I2IP, MIX_PRMT_DP2A, MIX_PRMT_SHF_DP2A, MIX_PRMT_IMAD, N_ALU_OPS };
const char* kOpNames[] = {"dp2a", "dp4a", "prmt", "shf.r.wrap", "imad", "lop3", "iadd3", "ffma", "selp", "cvt.pack.sat(I2IP)",
                          "prmt+dp2a (1:1)", "2 prmt + 1 shf + 2 dp2a (bicubic mix)", "prmt+imad (1:1)"};

template <int OP>
__global__ void __launch_bounds__(1024) alu_kernel(unsigned* out, unsigned seed, long long* cycles) {
    unsigned v[CH], w = seed * 2654435761u + threadIdx.x, s = (threadIdx.x & 3) * 8;
    float f[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { v[c] = seed + c * 977u + threadIdx.x; f[c] = (float)v[c]; }
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == DP2A) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(v[c]) : "r"(w), "r"(v[(c + 1) % CH]));
            if (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(v[c]) : "r"(w), "r"(v[(c + 1) % CH]));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5210;" : "+r"(v[c]) : "r"(w));
            if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(w), "r"(s));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(w), "r"(s));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(w), "r"(s));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(w));
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(1.0001f), "f"(0.5f));
            if (OP == SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.u32 %0, %0, %1, p;}" : "+r"(v[c]) : "r"(w), "r"(s));
            if (OP == I2IP) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(w), "r"(s));
            if (OP == MIX_PRMT_DP2A) {
                asm volatile("prmt.b32 %0, %0, %1, 0x5210;" : "+r"(v[c]) : "r"(w));
                asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(v[(c + 3) % CH]) : "r"(w), "r"(v[c]));
            }
            if (OP == MIX_PRMT_SHF_DP2A) {
                asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(w), "r"(s));
                asm volatile("prmt.b32 %0, %0, %1, 0x5210;" : "+r"(v[c]) : "r"(w));
                asm volatile("prmt.b32 %0, %0, %1, 0x6310;" : "+r"(v[(c + 1) % CH]) : "r"(v[c]));
                asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(v[(c + 3) % CH]) : "r"(w), "r"(v[c]));
                asm volatile("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(v[(c + 5) % CH]) : "r"(w), "r"(v[c]));
            }
            if (OP == MIX_PRMT_IMAD) {
                asm volatile("prmt.b32 %0, %0, %1, 0x5210;" : "+r"(v[c]) : "r"(w));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[(c + 3) % CH]) : "r"(w), "r"(s));
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= v[c] ^ __float_as_uint(f[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// Shared-memory loads: lane l of a warp reads `width` bytes at byte offset base + l * stride (+ row term), the
// pattern of the tap loads (stride 6 = 8K panorama -> 1600 px view, 3 = 1:1 magnification); table-like random
// 16-byte entries for the bicubic weight table.
enum LdsPat { L32_S4, L32_S6, L32_S8, L32_S12, L64_S6, L64_S8, L64_S12, L128_S16, L128_RANDOM, L32_RANDOM, SHFL_DOWN, N_LDS };
const char* kLdsNames[] = {"lds.32 lane stride 4 B (conflict-free)", "lds.32 lane stride 6 B", "lds.32 lane stride 8 B", "lds.32 lane stride 12 B",
                           "lds.64 lane stride 6 B (8 B aligned)", "lds.64 lane stride 8 B", "lds.64 lane stride 12 B (8 B aligned)",
                           "lds.128 lane stride 16 B", "lds.128 random 16 B entries of a 16 KB table", "lds.32 random words of 16 KB", "shfl.down"};

template <int PAT>
__global__ void __launch_bounds__(1024) lds_kernel(unsigned* out, unsigned seed, long long* cycles) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i * 2654435761u + seed;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned addr;
    if (PAT == L32_S4) addr = lane * 4;
    if (PAT == L32_S6) addr = (lane * 6) & ~3;
    if (PAT == L32_S8) addr = lane * 8;
    if (PAT == L32_S12) addr = lane * 12;
    if (PAT == L64_S6) addr = (lane * 6) & ~7;
    if (PAT == L64_S8) addr = lane * 8;
    if (PAT == L64_S12) addr = (lane * 12) & ~7;
    if (PAT == L128_S16) addr = lane * 16;
    if (PAT == L128_RANDOM) addr = ((lane * 2654435761u + warp * 40503u + seed) >> 7) & 0x3ff0;
    if (PAT == L32_RANDOM) addr = ((lane * 2654435761u + warp * 40503u + seed) >> 7) & 0x3ffc;
    if (PAT == SHFL_DOWN) addr = 0;
    addr += (warp & 7) * 512;            // rows of a patch: 128-byte aligned offsets, same banks
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    unsigned acc[CH] = {};
    unsigned ac[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) ac[c] = base + ((addr + c * 1024) & 0x3fff);
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const unsigned a = ac[c];
            if (PAT <= L32_S12 || PAT == L32_RANDOM) {
                unsigned v;
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
                acc[c] ^= v;
            } else if (PAT <= L64_S12) {
                unsigned x, y;
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a & ~7u) : "memory");
                acc[c] ^= x + y;
            } else if (PAT <= L128_RANDOM) {
                unsigned x, y, z, w;
                asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a & ~15u) : "memory");
                acc[c] ^= x + y + z + w;
            } else {
                acc[c] ^= __shfl_down_sync(0xffffffffu, acc[(c + 1) % CH] + it, 1);
            }
        }
    }
    const long long t1 = clock64();
    unsigned r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// A row of bicubic taps, fetched the two ways the sampler can: four aligned 32-bit words from (addr & ~3), or three
// aligned 64-bit words from (addr & ~7) plus the selects.  Lane stride `stride` bytes; rows_per_warp: how many patch
// rows (128-byte aligned pitch: same banks) the lanes of one warp are spread over.
template <int WIDE, int STRIDE, int ROWS>
__global__ void __launch_bounds__(1024) taprow_kernel(unsigned* out, unsigned seed, long long* cycles) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i * 2654435761u + seed;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
    // lanes drift over ROWS patch rows like a rotated view: lane l sits in row (l * ROWS) / 32
    const unsigned addr = base + (unsigned)(lane * STRIDE + (seed & 3)) + (unsigned)((lane * ROWS) / 32) * 1024u + (warp & 3) * 256u;
    const bool odd = (addr & 4u) != 0;
    const unsigned sh = (addr & 3u) << 3;
    unsigned acc[4] = {};
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned a = addr + c * 4096u - (c == 3 ? 4096u * 3 : 0);    // four tap rows (kept inside the 16 KB)
            unsigned q0, q1, q2, q3;
            if (WIDE) {
                unsigned x0, y0, x1, y1, x2, y2;
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x0), "=r"(y0) : "r"(a & ~7u) : "memory");
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x1), "=r"(y1) : "r"((a & ~7u) + 8) : "memory");
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x2), "=r"(y2) : "r"((a & ~7u) + 16) : "memory");
                q0 = odd ? y0 : x0; q1 = odd ? x1 : y0; q2 = odd ? y1 : x1; q3 = odd ? x2 : y1;
            } else {
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(q0) : "r"(a & ~3u) : "memory");
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(q1) : "r"((a & ~3u) + 4) : "memory");
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(q2) : "r"((a & ~3u) + 8) : "memory");
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(q3) : "r"((a & ~3u) + 12) : "memory");
            }
            acc[c] ^= __funnelshift_r(q0, q1, sh) + __funnelshift_r(q1, q2, sh) + __funnelshift_r(q2, q3, sh);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] ^ acc[1] ^ acc[2] ^ acc[3];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int instr_per_body, unsigned* out, long long* d_cycles, int warps_per_sm) {
    long long cyc = 0;
    launch();
    CHECK(cudaDeviceSynchronize());
    launch();
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost));
    const double warp_instr = (double)ITER * CH * instr_per_body * warps_per_sm;
    printf("{\"probe\": \"%s\", \"warps_per_sm\": %d, \"cycles\": %lld, \"warp_instr_per_clk_per_sm\": %.3f, \"clk_per_warp_instr\": %.3f}\n",
           name, warps_per_sm, cyc, warp_instr / cyc, cyc / warp_instr);
    fflush(stdout);
}

int main() {
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned* out;
    long long* d_cycles;
    CHECK(cudaMalloc(&out, (size_t)sms * 2 * 1024 * 4));
    CHECK(cudaMalloc(&d_cycles, 8));
    const int grid = sms;                 // one 1024-thread block (32 warps) per SM
#define ALU(OP, N) run(kOpNames[OP], [&] { alu_kernel<OP><<<grid, 1024>>>(out, 7u, d_cycles); }, N, out, d_cycles, 32)
    ALU(DP2A, 1); ALU(DP4A, 1); ALU(PRMT, 1); ALU(SHF, 1); ALU(IMAD, 1); ALU(LOP3, 1); ALU(IADD3, 1); ALU(FFMA, 1); ALU(SEL, 2); ALU(I2IP, 1);
    ALU(MIX_PRMT_DP2A, 2); ALU(MIX_PRMT_SHF_DP2A, 5); ALU(MIX_PRMT_IMAD, 2);
#define LDS(PAT) run(kLdsNames[PAT], [&] { lds_kernel<PAT><<<grid, 1024, 16384 + 4096>>>(out, 7u, d_cycles); }, 1, out, d_cycles, 32)
    LDS(L32_S4); LDS(L32_S6); LDS(L32_S8); LDS(L32_S12); LDS(L64_S6); LDS(L64_S8); LDS(L64_S12); LDS(L128_S16); LDS(L128_RANDOM); LDS(L32_RANDOM); LDS(SHFL_DOWN);
#define TAP(WIDE, STRIDE, ROWS) run("tap row: " #WIDE " (1 = 3 x lds.64 + selects, 0 = 4 x lds.32), lane stride " #STRIDE " B, lanes over " #ROWS " patch rows; per tap ROW", \
        [&] { taprow_kernel<WIDE, STRIDE, ROWS><<<grid, 1024, 16384 + 4096>>>(out, 7u, d_cycles); }, 1, out, d_cycles, 32)
    // CH = 8 in run(): the kernel does 4 rows per iteration, so halve the printed per-instruction numbers' meaning:
    // "warp_instr" here counts tap rows x 2
    TAP(0, 6, 1); TAP(1, 6, 1); TAP(0, 6, 3); TAP(1, 6, 3); TAP(0, 3, 1); TAP(1, 3, 1); TAP(0, 12, 1); TAP(1, 12, 1);
    return 0;
}
