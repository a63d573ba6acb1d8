// Tiled fast path: one thread block per 32x32 output tile.
//
//  1. fit      the tile's source coordinates are evaluated in float64 at 6x6 Chebyshev nodes
//              (ray -> R.d -> atan2/asin or equisolid+Brown, exactly the direct path's math) and
//              turned into two bivariate degree-5 polynomials in tile-local units; 13 extra
//              check points and the top-degree coefficients bound the fit error.  Inside the
//              tile every pixel's coordinate then costs 10 FFMA instead of two float64 atan2,
//              and -- being tile-relative -- keeps ~1e-5 px accuracy that plain float32 cannot.
//  2. stage    the bounding source patch (+ tap margins) is copied into shared memory with the
//              bulk-async copy engine (cp.async.bulk, mbarrier completion), one row per issue;
//              pole rows are replicated while staging so the inner loop has no border logic.
//  3. sample   polynomial -> float32 cast emulation -> 1/32-px quantisation -> taps from the
//              patch -> cv2.remap arithmetic (r360_sample.cuh) -> output tile in shared memory.
//  4. store    bulk-async row stores (or plain stores for ragged tiles).
//
// Tiles the fit cannot describe to ~4e-6 px (around a pole), whose patch does not fit, that cross
// the panorama seam, or that touch the fisheye validity boundary run the direct float64 path
// for their pixels instead -- same results, slower.
#pragma once

#include "r360_common.cuh"
#include "r360_direct.cuh"
#include "r360_sample.cuh"

namespace r360 {

constexpr int kTile = 32;
constexpr int kFitN = 6;                 // nodes per axis (degree 5)
constexpr int kFitChecks = 13;
constexpr int kTapLo = 2, kTapHi = 3;    // patch margin around floor(coord): cubic taps -1..+2, +1 safety

enum TileFlags : int { kTileFitOk = 1, kTileAllValid = 2, kTileAllInvalid = 4 };

struct TilePlan {
    float kx[36];        // x(s,t) = sum kx[l*6+k] t^l s^k, units of 1/32 px, relative to x0
    float ky[36];
    int x0, y0;          // integer origins (source pixels; x unwrapped for ERP)
    int bx0, bx1;        // range of floor(x) over the tile's pixel centres (conservative)
    int by0, by1;
    int flags;
    int pad;
};
static_assert(sizeof(TilePlan) == 320, "TilePlan layout");

struct FitConstants {
    double node[kFitN];              // Chebyshev nodes on [-1, 1]
    double minv[kFitN * kFitN];      // monomial coefficients = minv * node values
    double check[kFitChecks][2];     // (s, t) of the check points
};
__constant__ FitConstants c_fit;

// Scratch for one fit (shared memory).
struct FitScratch {
    double fx[64], fy[64];           // node / check values
    double gx[36], gy[36];           // half-transformed
    double kx[36], ky[36];           // monomial coefficients (px)
    double resid[16];
    int valid_count, invalid_count;
    int bmin_x, bmax_x, bmin_y, bmax_y;
};

__device__ __forceinline__ void named_barrier_sync_64() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

__device__ __forceinline__ double poly2d(const double* K, double s, double t) {
    double acc = 0.0;
#pragma unroll
    for (int l = kFitN - 1; l >= 0; --l) {
        double row = K[l * 6 + 5];
#pragma unroll
        for (int k = 4; k >= 0; --k) row = fma(row, s, K[l * 6 + k]);
        acc = fma(acc, t, row);
    }
    return acc;
}

// Executed by threads 0..63 of the block (all 64 must call it); result lands in *plan (shared).
template <int PROJ>
__device__ void fit_tile(const ViewDev& view, const ErpDev& erp, const LensDev* lens, int i0, int j0,
                         FitScratch* fs, TilePlan* plan, int tid) {
    const double half = 0.5 * (kTile - 1);
    const double ci = i0 + half, cj = j0 + half;
    if (tid == 0) {
        fs->valid_count = 0; fs->invalid_count = 0;
        fs->bmin_x = INT_MAX; fs->bmax_x = INT_MIN; fs->bmin_y = INT_MAX; fs->bmax_y = INT_MIN;
    }
    // -- exact coordinates at the nodes and check points ----------------------------------
    double s = 0.0, t = 0.0, x = 0.0, y = 0.0;
    bool ok = true;
    if (tid < 36) { s = c_fit.node[tid % 6]; t = c_fit.node[tid / 6]; }
    else if (tid < 36 + kFitChecks) { s = c_fit.check[tid - 36][0]; t = c_fit.check[tid - 36][1]; }
    if (tid < 36 + kFitChecks) {
        ok = project_pixel<PROJ>(view, erp, lens, fma(s, half, ci), fma(t, half, cj), x, y);
        fs->fx[tid] = x; fs->fy[tid] = y;
    }
    named_barrier_sync_64();
    if (tid < 36 + kFitChecks) {
        if (PROJ == kProjErp) {
            // unwrap longitude against node 0 so the tile sees a continuous function
            const double ref = fs->fx[0], period = erp.su;
            double d = x - ref;
            d -= period * rint(d / period);
            x = ref + d;
        }
        atomicAdd(ok ? &fs->valid_count : &fs->invalid_count, 1);
    }
    named_barrier_sync_64();
    if (tid < 36 + kFitChecks) { fs->fx[tid] = x; }
    named_barrier_sync_64();
    // -- K = M F M^T --------------------------------------------------------------------------
    if (tid < 36) {
        const int m = tid / 6, b = tid % 6;
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            ax = fma(c_fit.minv[m * 6 + a], fs->fx[a * 6 + b], ax);
            ay = fma(c_fit.minv[m * 6 + a], fs->fy[a * 6 + b], ay);
        }
        fs->gx[tid] = ax; fs->gy[tid] = ay;
    }
    named_barrier_sync_64();
    if (tid < 36) {
        const int m = tid / 6, n = tid % 6;
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            ax = fma(fs->gx[m * 6 + b], c_fit.minv[n * 6 + b], ax);
            ay = fma(fs->gy[m * 6 + b], c_fit.minv[n * 6 + b], ay);
        }
        fs->kx[tid] = ax; fs->ky[tid] = ay;
    }
    named_barrier_sync_64();
    // -- fit residual at the check points -------------------------------------------------------
    if (tid >= 36 && tid < 36 + kFitChecks) {
        const double rx = fabs(poly2d(fs->kx, s, t) - x), ry = fabs(poly2d(fs->ky, s, t) - y);
        fs->resid[tid - 36] = fmax(rx, ry);
    }
    // -- conservative bounding box: the polynomial on the tile's boundary pixels ---------------
    {
        const double inv_half = 1.0 / half;
        for (int b = tid; b < 4 * (kTile - 1); b += 64) {
            const int side = b / (kTile - 1), q = b % (kTile - 1);
            int il, jl;
            if (side == 0) { il = q; jl = 0; }
            else if (side == 1) { il = kTile - 1; jl = q; }
            else if (side == 2) { il = kTile - 1 - q; jl = kTile - 1; }
            else { il = 0; jl = kTile - 1 - q; }
            const double ps = (il - half) * inv_half, pt = (jl - half) * inv_half;
            const double lim = 268435456.0;   // keeps the int arithmetic below overflow for wild fits
            const int fxv = (int)floor(fmin(fmax(poly2d(fs->kx, ps, pt), -lim), lim));
            const int fyv = (int)floor(fmin(fmax(poly2d(fs->ky, ps, pt), -lim), lim));
            atomicMin(&fs->bmin_x, fxv); atomicMax(&fs->bmax_x, fxv);
            atomicMin(&fs->bmin_y, fyv); atomicMax(&fs->bmax_y, fyv);
            if (PROJ == kProjFisheye) {
                double ex, ey;
                const bool v = project_pixel<PROJ>(view, erp, lens, (double)(i0 + il), (double)(j0 + jl), ex, ey);
                atomicAdd(v ? &fs->valid_count : &fs->invalid_count, 1);
            }
        }
    }
    named_barrier_sync_64();
    // -- origin, float32 coefficients in 1/32 px, verdict ---------------------------------------
    const int x0 = (fs->bmin_x + fs->bmax_x) >> 1, y0 = (fs->bmin_y + fs->bmax_y) >> 1;
    if (tid < 36) {
        double kx = fs->kx[tid], ky = fs->ky[tid];
        if (tid == 0) { kx -= x0; ky -= y0; }
        plan->kx[tid] = (float)(kx * 32.0);
        plan->ky[tid] = (float)(ky * 32.0);
    }
    if (tid == 0) {
        double top_x = 0.0, top_y = 0.0;
        for (int q = 0; q < 6; ++q) {
            top_x += fabs(fs->kx[q * 6 + 5]) + fabs(fs->kx[5 * 6 + q]);
            top_y += fabs(fs->ky[q * 6 + 5]) + fabs(fs->ky[5 * 6 + q]);
        }
        // T5 has leading coefficient 16: the monomial top row / column over 16 approximates the
        // size of the last Chebyshev terms, i.e. of the truncation error (calibrated offline:
        // tiles passing both tests have fit error < 1e-5 px).
        // Every comparison is written so that a NaN fails it.
        bool fit_ok = top_x * (1.0 / 16.0) < 2e-4 && top_y * (1.0 / 16.0) < 2e-4;
        for (int q = 0; q < kFitChecks; ++q) fit_ok = fit_ok && fs->resid[q] < 4e-6;
        fit_ok = fit_ok && (long long)fs->bmax_x - fs->bmin_x < 2048 && (long long)fs->bmax_y - fs->bmin_y < 2048;
        int flags = fit_ok ? kTileFitOk : 0;
        if (fs->invalid_count == 0) flags |= kTileAllValid;
        if (fs->valid_count == 0) flags |= kTileAllInvalid;
        plan->x0 = x0; plan->y0 = y0;
        plan->bx0 = fs->bmin_x; plan->bx1 = fs->bmax_x; plan->by0 = fs->bmin_y; plan->by1 = fs->bmax_y;
        plan->flags = flags; plan->pad = 0;
    }
}

// ---- bulk-async copy / mbarrier wrappers (PTX) ------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends for a hardware time slice per call; a copy that never completes is a
    // bug, so trap instead of hanging the device
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the tile kernel -----------------------------------------------------------------------------

struct TiledParams {
    LaunchParams lp;
    int tiles_x, tiles_y;
    int patch_budget;       // bytes of shared memory available for the staged patch
    int out_stage_bytes;    // kTile * kTile * channels * sizeof(TOut)
    int bulk_store_ok;      // destination layout allows 16-byte aligned row stores
    int bulk_load_ok;       // source layout allows 16-byte aligned row loads
    // debug outputs (coords entry point); all null in production launches
    float* dbg_x32; float* dbg_y32; double* dbg_x64; double* dbg_y64; unsigned char* dbg_valid;
};

struct PatchGeom {
    int xb0;        // unwrapped source byte column of patch byte 0 (multiple of 16)
    int row_bytes;  // bytes copied per row (multiple of 16)
    int pitch;      // patch row pitch in shared memory
    int y0, rows;
    bool ok;
};

template <int PROJ>
__device__ __forceinline__ PatchGeom patch_geometry(const TilePlan& tp, int src_w, int src_h, int px_bytes,
                                                    int budget, bool need_valid_fill) {
    PatchGeom g;
    const int xs0 = tp.bx0 - kTapLo, xs1 = tp.bx1 + kTapHi;
    const int ys0 = tp.by0 - kTapLo, ys1 = tp.by1 + kTapHi;
    g.xb0 = (xs0 * px_bytes) & ~15;
    const int xb1 = ((xs1 + 1) * px_bytes + 15) & ~15;
    g.row_bytes = xb1 - g.xb0;
    g.pitch = g.row_bytes + ((g.row_bytes & 127) == 0 ? 16 : 0);
    g.y0 = ys0;
    g.rows = ys1 - ys0 + 1;
    g.ok = (tp.flags & kTileFitOk) && g.rows > 0 && g.row_bytes > 0 && g.rows * g.pitch <= budget;
    // columns: the patch must lie inside one period of the panorama / inside the sensor
    g.ok = g.ok && xs0 >= 0 && xb1 <= src_w * px_bytes;
    if (PROJ == kProjFisheye) {
        g.ok = g.ok && ys0 >= 0 && ys1 < src_h && (tp.flags & kTileAllValid);
    }
    (void)need_valid_fill;
    return g;
}

template <int PROJ, int INTERP, typename TIn, typename TOut>
__global__ void __launch_bounds__(256) remap_tiled_kernel(const __grid_constant__ TiledParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const LaunchParams& p = P.lp;
    // shared-memory carve-up
    TilePlan* plan = reinterpret_cast<TilePlan*>(smem);                        // 320
    float* rowc = reinterpret_cast<float*>(smem + 320);                          // 32 * 12 floats = 1536
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 320 + 1536);              // 16
    FitScratch* fs = reinterpret_cast<FitScratch*>(smem + 2048);                 // fit scratch overlaps out stage + patch
    unsigned char* out_stage = smem + 2048;
    unsigned char* patch = smem + 2048 + ((P.out_stage_bytes + 127) & ~127);

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int ti = tile % P.tiles_x, tj = tile / P.tiles_x;
    const int i0 = ti * kTile, j0 = tj * kTile;
    const int v = blockIdx.y % p.n_views, g = blockIdx.y / p.n_views;
    const ViewDev& view = p.views[v];
    const bool debug = P.dbg_x32 != nullptr;

    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (tid < 64) fit_tile<PROJ>(view, p.erp, p.lens, i0, j0, fs, plan, tid);
    __syncthreads();   // plan visible; fit scratch dead from here on

    const int px_bytes = p.channels * (int)sizeof(TIn);
    const PatchGeom geo = patch_geometry<PROJ>(*plan, p.src.width, p.src.height, px_bytes, P.patch_budget,
                                               p.fill_invalid != 0);
    const bool all_invalid_fill = PROJ == kProjFisheye && p.fill_invalid && (plan->flags & kTileAllInvalid);
    const bool fast = geo.ok && P.bulk_load_ok && !all_invalid_fill;

    const int jl = tid >> 3;                 // tile row of this thread
    const int il0 = (tid & 7) * 4;           // first of its 4 pixels
    const long long dst_img = (long long)g * p.n_views_total + p.view_base + v;
    unsigned char* dst_base = p.dst.data + dst_img * p.dst.image_stride;

    if (!fast) {
        // ---- direct float64 path for this tile (or constant fill) --------------------------
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + il0 + q, j = j0 + jl;
            if (i >= p.dst.width || j >= p.dst.height) continue;
            if (debug) {
                double x, y;
                const bool ok = project_pixel<PROJ>(view, p.erp, p.lens, (double)i, (double)j, x, y);
                const long long o = ((long long)(p.view_base + v) * p.dst.height + j) * p.dst.width + i;
                P.dbg_x32[o] = (float)x; P.dbg_y32[o] = (float)y; P.dbg_x64[o] = x; P.dbg_y64[o] = y;
                if (P.dbg_valid) P.dbg_valid[o] = ok;
            } else {
                direct_pixel<PROJ, INTERP, TIn, TOut>(p, view, g, v, i, j);
            }
        }
        return;
    }

    // ---- stage the source patch (warp 0) ------------------------------------------------------
    const unsigned char* img = p.src.data + ((long long)g * p.n_lenses + view.slot) * p.src.image_stride;
    if (!debug && tid < 32) {
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)(geo.rows * geo.row_bytes));
        __syncwarp();
        for (int r = tid; r < geo.rows; r += 32) {
            int sy = geo.y0 + r;
            sy = min(max(sy, 0), p.src.height - 1);          // pole rows replicate (ERP); fisheye is in range
            bulk_g2s(patch + r * geo.pitch, img + (long long)sy * p.src.pitch + geo.xb0, (uint32_t)geo.row_bytes, bar);
        }
    }
    // ---- per-row polynomial coefficients (all threads) --------------------------------------
    for (int task = tid; task < kTile * 12; task += 256) {
        const int row = task / 12, c = task % 12;
        const float* K = c < 6 ? plan->kx : plan->ky;
        const int k = c % 6;
        const float t = (float)(2 * row - (kTile - 1)) * (1.0f / (kTile - 1));
        float a = K[5 * 6 + k];
#pragma unroll
        for (int l = 4; l >= 0; --l) a = fmaf(a, t, K[l * 6 + k]);
        rowc[row * 12 + c] = a;
    }
    __syncthreads();
    if (!debug) mbar_wait(bar, 0);

    // ---- pixels -------------------------------------------------------------------------------
    const float X0 = (float)(plan->x0 * 32), Y0 = (float)(plan->y0 * 32);
    const float* rc = rowc + jl * 12;
    const PatchTaps<TIn> taps{patch, geo.pitch, geo.xb0, geo.y0, p.channels};
    TOut* stage_row = reinterpret_cast<TOut*>(out_stage) + (jl * kTile + il0) * p.channels;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float s = (float)(2 * (il0 + q) - (kTile - 1)) * (1.0f / (kTile - 1));
        float dx = rc[5], dy = rc[11];
#pragma unroll
        for (int k = 4; k >= 0; --k) { dx = fmaf(dx, s, rc[k]); dy = fmaf(dy, s, rc[6 + k]); }
        // 32 * float32(x): the rounding of this add is the float32 cast cv2.remap's map would see
        const float sxf = __fadd_rn(dx, X0), syf = __fadd_rn(dy, Y0);
        if (debug) {
            const int i = i0 + il0 + q, j = j0 + jl;
            if (i < p.dst.width && j < p.dst.height) {
                const long long o = ((long long)(p.view_base + v) * p.dst.height + j) * p.dst.width + i;
                P.dbg_x32[o] = sxf * (1.0f / 32.0f); P.dbg_y32[o] = syf * (1.0f / 32.0f);
                P.dbg_x64[o] = (double)plan->x0 + (double)dx * (1.0 / 32.0);
                P.dbg_y64[o] = (double)plan->y0 + (double)dy * (1.0 / 32.0);
                if (P.dbg_valid) P.dbg_valid[o] = 1;
            }
            continue;
        }
        sample_pixel<INTERP, TIn, TOut>(taps, p.channels, p.src.width, p.src.height, p.border_value,
                                        sxf * (1.0f / 32.0f), syf * (1.0f / 32.0f), stage_row + q * p.channels);
    }
    if (debug) return;

    // ---- store the tile --------------------------------------------------------------------------
    const int row_out_bytes = kTile * p.channels * (int)sizeof(TOut);
    const bool full = i0 + kTile <= p.dst.width && j0 + kTile <= p.dst.height;
    if (full && P.bulk_store_ok) {
        fence_async_shared();
        __syncthreads();
        if (tid < kTile) {
            bulk_s2g(dst_base + (long long)(j0 + tid) * p.dst.pitch + (long long)i0 * p.channels * sizeof(TOut),
                     out_stage + tid * row_out_bytes, (uint32_t)row_out_bytes);
            bulk_commit();
            bulk_wait_read_all();
        }
    } else {
        __syncthreads();
        const int nelem = kTile * p.channels;
        for (int e = tid; e < kTile * nelem; e += 256) {
            const int r = e / nelem, c = e % nelem;
            const int i = i0 + c / p.channels, j = j0 + r;
            if (i < p.dst.width && j < p.dst.height)
                reinterpret_cast<TOut*>(dst_base + (long long)j * p.dst.pitch)[(long long)i0 * p.channels + c] =
                    reinterpret_cast<const TOut*>(out_stage)[r * nelem + c];
        }
    }
}

}  // namespace r360
