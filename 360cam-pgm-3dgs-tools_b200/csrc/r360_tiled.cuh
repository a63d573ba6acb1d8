// Tiled fast path.
//
// PLAN (once per view set, cached by the caller; r360_plan_* in the ABI)
//   For every 32x32 output tile of every view, `plan_kernel` evaluates the source coordinates in
//   float64 at 6x6 Chebyshev nodes (ray -> R.d -> atan2/asin, or equisolid + Brown: exactly the
//   direct path's math), converts them into two bivariate degree-5 polynomials in tile-local
//   units of 1/32 px, bounds the fit error with 13 extra check points and the top-degree
//   coefficients, and derives the bounding source patch.  368 bytes per tile (0.36 B per output
//   pixel, against 8 B for a float map) -- the analogue of the reference building its maps once
//   (DF:1857-1907) and applying them to every frame (DF:1996-2014).
//   Tiles the polynomial cannot describe to a few 1e-6 px (around a pole), whose patch does not
//   fit in shared memory, that cross the panorama seam or touch the fisheye validity boundary
//   are put on a fallback list.
//
// REMAP (every frame)
//   `remap_tiled_kernel`: one block per (tile, view, frame).  The patch is staged into shared
//   memory with the bulk-async copy engine (cp.async.bulk + mbarrier), pole rows replicated while
//   staging so the inner loop has no border logic; coordinates cost 10 FFMA per pixel and --
//   being tile-relative -- keep ~1e-5 px where plain float32 cannot; sampling is cv2.remap's
//   arithmetic; the output tile leaves through shared memory as bulk-async row stores.
//   `remap_fallback_kernel` runs the direct float64 path on the listed tiles.
#pragma once

#include <climits>

#include <cuda.h>

#include "r360_common.cuh"
#include "r360_direct.cuh"
#include "r360_sample.cuh"
#include "r360_fast_u8.cuh"
#include "r360_fast_u16.cuh"

namespace r360 {

constexpr int kTile = 32;
constexpr int kFitN = 6;                 // nodes per axis (degree 5)
constexpr int kFitChecks = 13;
// Patch margins around the range of floor(coord).  The quantised coordinate rint(32 x) / 32 can round
// up to the next integer, so bilinear taps reach floor(x) + 2 and bicubic taps floor(x) - 1 .. + 3;
// one more column / row is kept on the low side.
__host__ __device__ constexpr int tap_margin_lo(int interp) { return interp == kLanczos4 ? 4 : interp == kCubic ? 2 : 1; }
__host__ __device__ constexpr int tap_margin_hi(int interp) { return interp == kLanczos4 ? 5 : interp == kCubic ? 3 : 2; }

// kModeFast: patch staged with 2-D tensor TMA boxes; kModeFastRows: staged row by row with 1-D bulk
// copies (rows clamped at the poles, or wider than the largest box).
// kModeFastSeam: a panorama tile that straddles the +-180 degree seam: rows are staged in two pieces
// (unwrapped in shared memory) and every pixel wraps its coordinate before the float32 cast.
enum TileMode : int { kModeFallback = 0, kModeFast = 1, kModeFill = 2, kModeFastRows = 3, kModeFastSeam = 4 };

// Box shapes of the tensor-TMA descriptors: widths in bytes (odd multiples of 32 keep consecutive
// rows on different banks; 1024 = 256 uint32 elements is the hardware limit) x heights {32, 8}.
constexpr int kNumBoxWidths = 11;
constexpr int kNumBoxHeights = 2;
// Two families of widths.  Family 0: odd multiples of 32 bytes -- consecutive patch rows start 8 banks apart,
// which suits the paths whose warps read four tile rows at once (bilinear, generic sampler).  Family 1:
// multiples of 128 bytes -- every row starts on bank 0, so on the lane-per-column path of the 8-bit bicubic
// kernel (a warp reads one output row whose taps drift over a few source rows) the bank of a tap depends on
// its column only.  Measured on B200: bicubic 100.2 -> 103.3 Gpix/s with family 1, bilinear 215 -> 201.
__host__ __device__ constexpr int box_width_bytes(int k, int family) {
    if (family == 1) return k < 8 ? 128 * (k + 1) : 1024;
    return k == 0 ? 96 : k == 1 ? 160 : k == 2 ? 224 : k == 3 ? 288 : k == 4 ? 352 : k == 5 ? 416
         : k == 6 ? 544 : k == 7 ? 672 : k == 8 ? 800 : k == 9 ? 928 : 1024;
}
__host__ __device__ constexpr int box_height_rows(int k) { return k == 0 ? 32 : 8; }
// rows actually staged for `rows` needed rows: as many 32-row boxes as fit, the rest in 8-row boxes
__host__ __device__ constexpr int staged_rows(int rows) { return (rows / 32) * 32 + (((rows % 32) + 7) / 8) * 8; }

struct TilePlan {
    // source coordinate in units of 1/32 px, absolute:
    //   32 x = ax[0] + ax[1] * il + ax[2] * jl  +  sum rx[l*6+k] t^l s^k
    // (il, jl) = pixel index inside the tile, s = (2 il - 31) / 31, t = (2 jl - 31) / 31.  The affine
    // part carries the large magnitudes and is evaluated in float64; the residual polynomial (its
    // constant and pure-linear terms are zero) is small and is evaluated in float32.
    float rx[36];
    float ry[36];
    double ax[3];
    double ay[3];
    int py0, rows;       // first source row of the patch (may be < 0: clamped while staging), row count
    int xb0, row_bytes;  // source byte column of patch byte 0 (multiple of 16), bytes needed per row
    int pitch;           // patch row pitch in shared memory
    int mode_slot;       // TileMode | (source slot << 8) | (box width index << 16)
    int pad[2];          // [0] why the tile is on the fallback list (tools/fallback_reasons.py), -1: staged tile whose patch
                         // needs the large ring (large-patch pass), [1] 1 + index of the
                         // tile's per-pixel map in the coordinate pool (0: the polynomial is the map)
};
static_assert(sizeof(TilePlan) == 368, "TilePlan layout");

struct PlanHeader {      // first 256 bytes of the plan workspace
    int n_fallback;
    int n_coords;        // per-pixel maps handed out (may exceed the pool: the surplus tiles are fallback tiles)
    int pad[62];
};

// A tile whose map no degree-5 polynomial follows (the neighbourhood of a pole) but whose patch still fits the ring
// keeps an explicit map instead: 32 x 32 float64 pairs (32 x, 32 y), unwrapped like the polynomial's values, in a
// pool behind the plan records.  The remap kernel reads it where it would evaluate the polynomial; everything after
// that -- seam wrap, float32 cast, staging, sampling of all frames of the batch -- is the ordinary tile path.
constexpr int kCoordTileBytes = kTile * kTile * 16;

struct FitConstants {
    double node[kFitN];              // Chebyshev nodes on [-1, 1]
    double minv[kFitN * kFitN];      // monomial coefficients = minv * node values
    double check[kFitChecks][2];     // (s, t) of the check points
};
__constant__ FitConstants c_fit;

struct PlanParams {
    int proj;
    int out_w, out_h, tiles_x, tiles_y, n_views;
    int src_w, src_h, px_bytes;      // px_bytes = channels * sizeof(source element)
    int patch_budget;                // bytes of shared memory the remap kernel can give to a patch
    int patch_budget_large;          // ... in its one-block-per-SM, one-frame shape: the pass for tiles with large patches
    int bulk_load_ok;                // source layout allows 16-byte aligned row copies
    int tensor_ok;                   // tensor-TMA descriptors can be built for the source layout
    int fill_invalid;
    int interp;                      // decides the tap margins of the patch
    int box_family;                  // which list of tensor-box widths the remap kernel's descriptors use
    ErpDev erp;
    LensDev lens[kMaxLenses];
    const ViewDev* views;            // n_views, device
    TilePlan* plans;                 // n_views * tiles, device
    PlanHeader* header;
    int2* fallback;                  // (view, tile) list, device
    double2* coords;                 // pool of per-pixel maps (kTile * kTile entries each), nullptr: none
    int coords_capacity;             // maps in the pool
};

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

template <int PROJ>
__device__ __forceinline__ void plan_tile(const PlanParams& P) {
    __shared__ double fx[64], fy[64], gx[36], gy[36], kx[36], ky[36], resid[16];
    __shared__ int valid_count, invalid_count, bmin_x, bmax_x, bmin_y, bmax_y;
    __shared__ int s_poly_ok, s_coords_idx;
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, v = blockIdx.y;
    const int i0 = (tile % P.tiles_x) * kTile, j0 = (tile / P.tiles_x) * kTile;
    const ViewDev view = P.views[v];
    const double half = 0.5 * (kTile - 1);
    const double ci = i0 + half, cj = j0 + half;
    if (tid == 0) {
        valid_count = 0; invalid_count = 0;
        bmin_x = INT_MAX; bmax_x = INT_MIN; bmin_y = INT_MAX; bmax_y = INT_MIN;
    }
    // -- exact coordinates at the nodes and check points ----------------------------------------
    double s = 0.0, t = 0.0, x = 0.0, y = 0.0;
    bool ok = true;
    const bool active = tid < 36 + kFitChecks;
    if (tid < 36) { s = c_fit.node[tid % 6]; t = c_fit.node[tid / 6]; }
    else if (active) { s = c_fit.check[tid - 36][0]; t = c_fit.check[tid - 36][1]; }
    if (active) {
        ok = project_pixel<PROJ>(view, P.erp, P.lens, fma(s, half, ci), fma(t, half, cj), x, y);
        fx[tid] = x; fy[tid] = y;
    }
    __syncthreads();
    if (active) {
        if (PROJ == kProjErp) {
            // unwrap longitude against node 0 so the tile sees a continuous function
            const double ref = fx[0], period = P.erp.su;
            double d = x - ref;
            d -= period * rint(d / period);
            x = ref + d;
        }
        atomicAdd(ok ? &valid_count : &invalid_count, 1);
    }
    __syncthreads();
    if (active) fx[tid] = x;
    __syncthreads();
    // -- K = M F M^T ----------------------------------------------------------------------------
    if (tid < 36) {
        const int m = tid / 6, b = tid % 6;
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            ax = fma(c_fit.minv[m * 6 + a], fx[a * 6 + b], ax);
            ay = fma(c_fit.minv[m * 6 + a], fy[a * 6 + b], ay);
        }
        gx[tid] = ax; gy[tid] = ay;
    }
    __syncthreads();
    if (tid < 36) {
        const int m = tid / 6, n = tid % 6;
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            ax = fma(gx[m * 6 + b], c_fit.minv[n * 6 + b], ax);
            ay = fma(gy[m * 6 + b], c_fit.minv[n * 6 + b], ay);
        }
        kx[tid] = ax; ky[tid] = ay;
    }
    __syncthreads();
    // -- fit residual at the check points ---------------------------------------------------------
    if (tid >= 36 && active)
        resid[tid - 36] = fmax(fabs(poly2d(kx, s, t) - x), fabs(poly2d(ky, s, t) - y));
    // -- bounding box of floor(coord): the polynomial on the tile's boundary pixels (the maps have
    //    no interior extrema away from the poles, and pole tiles fail the fit) -----------------
    {
        const double inv_half = 1.0 / half, lim = 268435456.0;
        for (int b = tid; b < 4 * (kTile - 1); b += 64) {
            const int side = b / (kTile - 1), q = b % (kTile - 1);
            int il, jl;
            if (side == 0) { il = q; jl = 0; }
            else if (side == 1) { il = kTile - 1; jl = q; }
            else if (side == 2) { il = kTile - 1 - q; jl = kTile - 1; }
            else { il = 0; jl = kTile - 1 - q; }
            const double ps = (il - half) * inv_half, pt = (jl - half) * inv_half;
            const int fxv = (int)floor(fmin(fmax(poly2d(kx, ps, pt), -lim), lim));
            const int fyv = (int)floor(fmin(fmax(poly2d(ky, ps, pt), -lim), lim));
            atomicMin(&bmin_x, fxv); atomicMax(&bmax_x, fxv);
            atomicMin(&bmin_y, fyv); atomicMax(&bmax_y, fyv);
            if (PROJ != kProjErp) {
                double ex, ey;
                const bool vld = project_pixel<PROJ>(view, P.erp, P.lens, (double)(i0 + il), (double)(j0 + jl), ex, ey);
                atomicAdd(vld ? &valid_count : &invalid_count, 1);
            }
        }
    }
    __syncthreads();
    // -- origin, float32 coefficients, verdict ------------------------------------------------------
    TilePlan* out = P.plans + (long long)v * (P.tiles_x * P.tiles_y) + tile;
    if (tid < 36) {
        // residual = everything but the constant and the two pure-linear terms
        const bool affine_term = tid == 0 || tid == 1 || tid == 6;
        out->rx[tid] = affine_term ? 0.0f : (float)(kx[tid] * 32.0);
        out->ry[tid] = affine_term ? 0.0f : (float)(ky[tid] * 32.0);
    }
    if (tid == 36) {
        // K00 + K01 s + K10 t with s = (2 il - 31) / 31, t likewise, in units of 1/32 px
        const double g = 2.0 / (kTile - 1);
        out->ax[0] = (kx[0] - kx[1] - kx[6]) * 32.0; out->ax[1] = kx[1] * g * 32.0; out->ax[2] = kx[6] * g * 32.0;
        out->ay[0] = (ky[0] - ky[1] - ky[6]) * 32.0; out->ay[1] = ky[1] * g * 32.0; out->ay[2] = ky[6] * g * 32.0;
    }
    if (tid == 0) {
        double top_x = 0.0, top_y = 0.0;
        for (int q = 0; q < 6; ++q) {
            top_x += fabs(kx[q * 6 + 5]) + fabs(kx[5 * 6 + q]);
            top_y += fabs(ky[q * 6 + 5]) + fabs(ky[5 * 6 + q]);
        }
        // T5 has leading coefficient 16: the monomial top row / column over 16 approximates the
        // size of the last Chebyshev terms, i.e. of the truncation error (calibrated offline:
        // tiles passing both tests have fit error < 1e-5 px).  Every
        // comparison is written so that a NaN fails it.
        bool ok_fit = top_x * (1.0 / 16.0) < 2e-4 && top_y * (1.0 / 16.0) < 2e-4;
        for (int q = 0; q < kFitChecks; ++q) ok_fit = ok_fit && resid[q] < 4e-6;
        s_poly_ok = ok_fit ? 1 : 0;
        s_coords_idx = -1;
    }
    __syncthreads();
    // -- no polynomial (panorama tiles next to a pole): the exact map of all 1024 pixels, first its bounding box ----
    const bool exact = PROJ == kProjErp && P.coords != nullptr && !s_poly_ok;
    if (exact) {
        if (tid == 0) { bmin_x = INT_MAX; bmax_x = INT_MIN; bmin_y = INT_MAX; bmax_y = INT_MIN; }
        __syncthreads();
        const double ref = fx[0], period = P.erp.su, lim = 268435456.0;
        for (int px = tid; px < kTile * kTile; px += 64) {
            double ex, ey;
            project_pixel<PROJ>(view, P.erp, P.lens, (double)(i0 + (px & (kTile - 1))), (double)(j0 + px / kTile), ex, ey);
            double d = ex - ref;
            d -= period * rint(d / period);
            const int fxv = (int)floor(fmin(fmax(ref + d, -lim), lim));
            const int fyv = (int)floor(fmin(fmax(ey, -lim), lim));
            atomicMin(&bmin_x, fxv); atomicMax(&bmax_x, fxv);
            atomicMin(&bmin_y, fyv); atomicMax(&bmax_y, fyv);
        }
        __syncthreads();
    }
    if (tid == 0) {
        const bool poly_ok = s_poly_ok != 0;
        bool fit_ok = poly_ok || exact;
        fit_ok = fit_ok && (long long)bmax_x - bmin_x < 2048 && (long long)bmax_y - bmin_y < 2048;

        // patch geometry
        const int lo = tap_margin_lo(P.interp), hi = tap_margin_hi(P.interp);
        const int xs0 = bmin_x - lo, xs1 = bmax_x + hi;
        const int ys0 = bmin_y - lo, ys1 = bmax_y + hi;
        int mode = kModeFallback, wbox = 0;
        int xb0 = 0, row_bytes = 0, pitch = 0, rows = 0;
        // why a tile is left to the direct path (pad[0] of its record; tools/fallback_reasons.py reads it):
        // 1 the polynomial does not fit, 2 the patch does not fit the ring, 3 columns outside one period / the
        // sensor (and no seam handling for this geometry), 4 rows outside the sensor or invalid pixels, 5 the
        // map is fine but the tile spans 2048 source pixels or more, 6 the coordinate pool is full
        int reason = fit_ok ? 3 : (poly_ok || exact) ? 5 : 1;
        if (fit_ok) {
            xb0 = (xs0 * P.px_bytes) & ~15;
            const int xb1 = ((xs1 + 1) * P.px_bytes + 15) & ~15;
            row_bytes = xb1 - xb0;
            rows = ys1 - ys0 + 1;
            // columns must lie inside one period of the panorama / inside the sensor
            const int row_total = P.src_w * P.px_bytes;
            bool fast = P.bulk_load_ok && xs0 >= 0 && xb1 <= row_total;
            if (PROJ != kProjErp && fast && !(ys0 >= 0 && ys1 < P.src_h && invalid_count == 0)) { fast = false; reason = 4; }
            // seam: only when the coordinate period equals the image width (halfpixel convention)
            const bool seam = PROJ == kProjErp && P.bulk_load_ok && !fast && P.erp.su == (double)P.src_w &&
                              row_bytes <= row_total && xb0 > -row_total && xb1 < 2 * row_total;
            if (fast) {
                while (wbox < kNumBoxWidths && box_width_bytes(wbox, P.box_family) < row_bytes) ++wbox;
                const bool rows_inside = ys0 >= 0 && ys1 < P.src_h;
                if (P.tensor_ok && wbox < kNumBoxWidths && rows_inside &&
                    staged_rows(rows) * box_width_bytes(wbox, P.box_family) <= P.patch_budget_large) {
                    mode = kModeFast;                                    // tensor boxes: pitch = box width
                    pitch = box_width_bytes(wbox, P.box_family);
                } else {
                    wbox = 0;
                    pitch = row_bytes + ((row_bytes & 127) == 0 ? 16 : 0);   // keep rows off the same banks
                    if (rows * pitch <= P.patch_budget_large) mode = kModeFastRows;
                    else reason = 2;
                }
            } else if (seam) {
                pitch = row_bytes + ((row_bytes & 127) == 0 ? 16 : 0);
                if (rows * pitch <= P.patch_budget_large) mode = kModeFastSeam;
                else reason = 2;
            }
        }
        if (PROJ != kProjErp && P.fill_invalid && valid_count == 0) mode = kModeFill;
        // a patch that only the large ring holds: the tile leaves the main walk for the large-patch pass (pad[0] = -1)
        const bool large = (mode == kModeFast || mode == kModeFastRows || mode == kModeFastSeam) &&
                           (mode == kModeFast ? staged_rows(rows) : rows) * pitch > P.patch_budget;
        if (exact && mode != kModeFallback) {
            const int idx = atomicAdd(&P.header->n_coords, 1);
            if (idx < P.coords_capacity) s_coords_idx = idx;
            else { mode = kModeFallback; reason = 6; }
        }
        out->py0 = ys0; out->rows = rows; out->xb0 = xb0; out->row_bytes = row_bytes; out->pitch = pitch;
        out->mode_slot = mode | (view.slot << 8) | (wbox << 16);
        out->pad[0] = mode == kModeFallback ? reason : large ? -1 : 0; out->pad[1] = s_coords_idx + 1;
        if (mode == kModeFallback) {
            const int idx = atomicAdd(&P.header->n_fallback, 1);
            P.fallback[idx] = make_int2(v, tile);
        }
    }
    __syncthreads();
    if (exact && s_coords_idx >= 0) {
        // the map itself, in the units and the unwrapping of the polynomial's values (1/32 px, continuous across the seam)
        double2* cm = P.coords + (long long)s_coords_idx * (kTile * kTile);
        const double ref = fx[0], period = P.erp.su;
        for (int px = tid; px < kTile * kTile; px += 64) {
            double ex, ey;
            project_pixel<PROJ>(view, P.erp, P.lens, (double)(i0 + (px & (kTile - 1))), (double)(j0 + px / kTile), ex, ey);
            double d = ex - ref;
            d -= period * rint(d / period);
            cm[px] = make_double2((ref + d) * 32.0, ey * 32.0);
        }
    }
}

__global__ void __launch_bounds__(64) plan_kernel(const __grid_constant__ PlanParams P) {
    if (P.proj == kProjErp) plan_tile<kProjErp>(P);
    else if (P.proj == kProjFisheye) plan_tile<kProjFisheye>(P);
    else plan_tile<kProjUndistort>(P);
}

// Sort key of every (view, tile) for the remap kernel's walk: source slot, then the centre row of the patch;
// fallback tiles (not staged) get INT_MAX and drop off the end of the list, tiles of the large-patch pass INT_MAX - 1.
__global__ void __launch_bounds__(256) order_key_kernel(const TilePlan* plans, int n, int* keys, int* patch_bytes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TilePlan& p = plans[i];
    const int mode = p.mode_slot & 0xff, slot = (p.mode_slot >> 8) & 0xff;
    const bool staged = mode == kModeFast || mode == kModeFastRows || mode == kModeFastSeam;
    keys[i] = mode == kModeFallback ? INT_MAX : p.pad[0] == -1 ? INT_MAX - 1
            : (slot << 24) + (staged ? max(0, min(p.py0 + p.rows / 2, (1 << 24) - 1)) : 0);
    // bytes the tile's patch takes in the ring (one frame): the host sizes the batch shape by their mean
    patch_bytes[i] = staged ? (mode == kModeFast ? staged_rows(p.rows) : p.rows) * p.pitch : 0;
}

// ---- bulk-async copy / mbarrier wrappers (PTX) ------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar_saddr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0xF4240;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar_saddr), "r"(parity) : "memory");
    return done != 0;
}
// Wait on a barrier given by its shared-memory address.  try_wait suspends the warp for a hardware time slice
// per call; a copy that never completes is a bug, so trap instead of hanging the device.
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, uint32_t parity) {
    if (mbar_try_wait_s(bar_saddr, parity)) return;
    for (unsigned spins = 0; !mbar_try_wait_s(bar_saddr, parity); ++spins)
        if (spins > 4000000u) __trap();
}
// The producer's wait (a slot to be released): it sleeps between polls so that the polling does not take issue
// slots from the consumer warps of its scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    if (mbar_try_wait_s(a, parity)) return;
    for (unsigned spins = 0; !mbar_try_wait_s(a, parity); ++spins) {
        __nanosleep(128);
        if (spins > 8000000u) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar_saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { mbar_arrive_s(smem_u32(bar)); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// L2 policy of the patch loads: the source frames are re-read by every view that overlaps them (and by the
// neighbouring tiles' halos) -> evict_last; the outputs are written once and never read -> streamed (st.global.cs).
// Measured on B200 (16 x 8K frames, 12 views): DRAM reads 2.84 GB with evict_last, 4.84 GB with evict_normal.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_by_kind(int kind) {
    uint64_t pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else pol = l2_policy_evict_last();
    return pol;
}
__device__ __forceinline__ void tensor_g2s_3d(void* smem_dst, const void* tmap, int x, int y, int z, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
                 "[%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_global_streaming(void* gptr, const int4& v) {
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- the remap kernel -------------------------------------------------------------------------------
//
// Persistent and warp-specialised.  Each block walks work items (tile, view, block of FR consecutive frames):
//   * one producer warp runs ahead.  Per item it reads the tile's plan record, carves space for the patches
//     (the same source rectangle of each of the item's frames) out of a shared-memory ring (variable-size,
//     FIFO), issues the tensor-TMA boxes (or per-row bulk copies) that complete on the slot's mbarrier, and
//     publishes a slot record.  Items whose FR patches do not fit the budget, and the frames left over when
//     the batch is not a multiple of FR, are published as one-frame slots; fallback tiles are skipped;
//   * consumer warps work in teams of eight (a block has one or two teams, slot k belongs to team k mod
//     teams).  A team's warps each own 4 rows of the tile and never synchronise with one another: wait for
//     the slot's mbarrier, generate the coordinates, tap addresses and weights of their pixels ONCE and sample
//     every frame of the item with them (what a video / batch shares between frames: the map), write their rows;
//     the last reader of a slot releases it to the producer through a second mbarrier (8 arrivals).
//   The producer ends the stream with one `kModeExit` slot per team.

// FIFO allocator for variable-size patches in the shared-memory ring.  The bytes in flight form
// one circular interval [tail, head) of length `used`; allocations are contiguous (an allocation
// that does not fit before the end of the ring skips the end and is charged for the skipped
// bytes) and are released in allocation order.
struct PatchRing {
    int head = 0, tail = 0, used = 0;
    // On success returns true with the byte offset and the bytes to give back on release.
    __host__ __device__ bool try_alloc(int need, int capacity, int& off, int& charge) {
        if (used == 0) { head = 0; tail = 0; }
        if (need == 0) { off = head; charge = 0; return true; }
        if (used == 0 || head > tail) {
            if (head + need <= capacity) { off = head; charge = need; }
            else if (need <= tail) { off = 0; charge = need + (capacity - head); }
            else return false;
        } else if (head < tail) {
            if (head + need <= tail) { off = head; charge = need; } else return false;
        } else {
            return false;                       // head == tail with bytes in use: full
        }
        used += charge;
        head = off + need;
        return true;
    }
    __host__ __device__ void release(int charge, int capacity) {
        used -= charge;
        tail += charge;
        if (tail >= capacity) tail -= capacity;
    }
};

constexpr int kMaxFramesPerItem = 4;

struct TiledParams {
    ImageSetDev src, dst;
    int channels;
    int n_views;            // views per source group
    int n_groups;           // source groups (frames) in this launch
    int n_lenses;
    int tiles_x, tiles_y;
    int out_stage_bytes;    // kTile * kTile * channels * sizeof(TOut), rounded up to 128 (one frame of one team)
    int ring_bytes;         // shared-memory ring for patches (multiple of 128)
    int bulk_store_ok;      // destination layout allows 16-byte aligned row stores
    int use_table;          // copy the fixed-point cubic table into shared memory
    int frames_per_item;    // FR: must equal the kernel's template argument
    int multi_budget;       // an item takes FR frames at once when FR * (patch bytes) <= this
    float border_value;
    long long dst_fstride;  // bytes between the same view of consecutive frames (n_views * image stride)
    const TilePlan* plans;  // whole plan (all views)
    unsigned int* work_counter;   // zeroed before the launch: the next item of the walk (claimed by the producers)
    const int2* order;      // (view, tile) of the staged tiles sorted by source row: the order the items are walked in
    int n_order;            // entries of `order` (tiles that are not on the fallback list)
    const double2* coords;  // pool of per-pixel maps (tiles with TilePlan::pad[1] > 0)
    int l2_policy;          // patch loads: 0 evict_last, 1 evict_normal, 2 evict_first (R360_L2_POLICY, experiments)
};

// Shape of a block: one or two consumer teams of eight warps (a template argument of the kernel) and how many blocks
// per SM the register allocation must allow.
constexpr int kConsumerWarps = 8;                       // warps of a team (256 threads <-> 32 rows x 8 lanes)
constexpr int kTeamThreads = kConsumerWarps * 32;
constexpr int kMaxTeams = 2;
constexpr int kSlots = 4;                               // slots in flight per block (a multiple of the team count); eight
                                                        // measured 1-7 % slower on every kernel (profiles/README.md)
static_assert(kSlots % kMaxTeams == 0, "slot k belongs to team k mod teams");
__host__ __device__ constexpr int tiled_threads(int teams) { return teams * kTeamThreads + 32; }
// blocks per SM the registers are budgeted for: the 8-bit bicubic kernel shares the SM with its 32 KB weight
// table (2 blocks), lanczos4 with a 128 KB one (1 block); bilinear, 16-bit and float samplers want the registers
// (and, with several frames per item, the shared memory) of 2 blocks -- measured: bilinear 8-bit, two frames per
// item, 266 Gpix/s at 2 blocks per SM against 208 at 4; nearest runs 4 blocks of one team.  Two teams halve it.
__host__ __device__ constexpr int tiled_default_blocks(int elem_bytes, int interp) {
    return (elem_bytes == 1 && interp == kLanczos4) ? 1 : (interp == kLinear || interp == kCubic || elem_bytes > 1) ? 2 : 4;
}
__host__ __device__ constexpr int tiled_min_blocks(int elem_bytes, int interp, int teams) {
    return (tiled_default_blocks(elem_bytes, interp) + teams - 1) / teams;
}
constexpr int kModeExit = 255;                          // slot record that ends a team's stream

struct SlotInfo {            // per slot, written by the producer
    uint32_t patch_saddr;    // shared-memory address of the first frame's patch
    uint32_t bias;           // fast samplers: tap address bias (r360_fast_u8.cuh)
    int size;                // ring bytes to give back on release
    int mode;                // TileMode
    int pitch, xb0, py0;     // patch geometry
    int full_tile;           // the tile lies completely inside the destination image
    int i0, j0;
    long long dst_tile;      // byte offset of the tile's first output pixel inside the destination batch (first frame)
    long long dst_off;       // byte offset of the destination image (first frame, view)
    int nf;                  // frames of this slot (1 or FR)
    uint32_t fstride;        // shared-memory bytes between the patches of consecutive frames
};
static_assert(sizeof(SlotInfo) == 64, "SlotInfo layout");

constexpr int kTableBytes = 32 * 32 * 16 * 2;              // bicubic: cv2's 15-bit 4 x 4 table
constexpr int kLanczosTableBytes = 32 * 32 * 64 * 2;       // lanczos4: the 8 x 8 one (TiledParams::use_table == 2)
__host__ __device__ constexpr int table_bytes(int use_table) { return use_table == 2 ? kLanczosTableBytes : use_table ? kTableBytes : 0; }
// fixed part of the dynamic shared memory: barriers | slot records | per-team row coefficients (32 rows x 12
// floats) | plan records
constexpr int kSmemSlots = (2 * kSlots * 8 + 63) / 64 * 64;
constexpr int kSmemRowc = kSmemSlots + kSlots * 64;
constexpr int kSmemPlans = kSmemRowc + kMaxTeams * 1536;
constexpr int kTiledFixedSmem = (kSmemPlans + kSlots * 368 + 127) / 128 * 128;
static_assert(kTiledFixedSmem % 128 == 0 && kTableBytes % 128 == 0 && kSmemPlans % 16 == 0, "ring alignment");

// One descriptor per box shape, all over the same 3-D view of the source batch:
// (row bytes / 4 as uint32, rows, images) with strides (pitch, image stride).
struct TensorMaps {
    CUtensorMap m[kNumBoxWidths * kNumBoxHeights];     // [width index][height index]
};

// Residual polynomial of one pixel: Horner in s over the row's 12 coefficients (6 for x, 6 for y).
__device__ __forceinline__ void residual_xy(const float4& c0, const float4& c1, const float4& c2, float s, float& dx, float& dy) {
    dx = c1.y; dy = c2.w;
    dx = fmaf(dx, s, c1.x); dx = fmaf(dx, s, c0.w); dx = fmaf(dx, s, c0.z); dx = fmaf(dx, s, c0.y); dx = fmaf(dx, s, c0.x);
    dy = fmaf(dy, s, c2.z); dy = fmaf(dy, s, c2.y); dy = fmaf(dy, s, c2.x); dy = fmaf(dy, s, c1.w); dy = fmaf(dy, s, c1.z);
}

// ---- 8-bit RGB, lane = pixel column, 4 rows per lane ---------------------------------------------------
// With adjacent lanes on adjacent pixels the tap loads of a warp fall into neighbouring words (few bank conflicts),
// and the finished row leaves straight from registers: three lanes out of four hold one 32-bit word of the 96-byte
// row after a shuffle, so the store is contiguous.  The weights and the tap address of a pixel serve all NF frames
// of the slot.  Bilinear and bicubic full tiles both take this path (measured on B200, 16 x 8K frames -> 12 views,
// two frames per item: bilinear 266 Gpix/s here against 244 with four pixels per lane through the output stage).
template <int INTERP, int NF>
__device__ __forceinline__ void column_rows_u8c3(const TilePlan* plan, const float* rowc_warp, uint32_t bias, uint32_t pitch,
                                                 uint32_t tab, uint32_t fstride, float s, double dlane, double drow,
                                                 unsigned char* out_word, long long dst_pitch, long long dst_fstride, int m) {
    const double axi = plan->ax[1], ayi = plan->ay[1], axj = plan->ax[2], ayj = plan->ay[2];
    double ax_r = fma(axi, dlane, fma(axj, drow, plan->ax[0]));
    double ay_r = fma(ayi, dlane, fma(ayj, drow, plan->ay[0]));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float* rq = rowc_warp + q * 12;
        const float4 c0 = *reinterpret_cast<const float4*>(rq), c1 = *reinterpret_cast<const float4*>(rq + 4),
                     c2 = *reinterpret_cast<const float4*>(rq + 8);
        float dx, dy;
        residual_xy(c0, c1, c2, s, dx, dy);
        const float sx = __double2float_rn(ax_r + (double)dx), sy = __double2float_rn(ay_r + (double)dy);
        ax_r += axj; ay_r += ayj;
        uint32_t own[NF];
        if constexpr (INTERP == kCubic) {
            const BicubicPrep pp = bicubic_prep_u8c3(bias, pitch, tab, round_bits(sx), round_bits(sy));
#pragma unroll
            for (int f = 0; f < NF; ++f) own[f] = bicubic_taps_u8c3(pp, pitch, (uint32_t)f * fstride);
        } else {
            const BilinearPrep pp = bilinear_prep_u8c3(bias, pitch, round_bits(sx), round_bits(sy));
#pragma unroll
            for (int f = 0; f < NF; ++f) own[f] = bilinear_taps_u8c3(pp, pitch, (uint32_t)f * fstride);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const uint32_t nxt = __shfl_down_sync(0xffffffffu, own[f], 1);
            const uint32_t word = (own[f] >> (8 * m)) | (nxt << (24 - 8 * m));
            // three lanes out of four store: a predicated store (no divergent branch around it; no "memory"
            // clobber: the outputs are never read back, and the next row's tap loads may move above it)
            asm volatile("{\n.reg .pred p;\nsetp.lt.u32 p, %2, 3;\n@p st.global.cs.b32 [%0], %1;\n}"
                         ::"l"(out_word + f * dst_fstride), "r"(word), "r"((uint32_t)m));
        }
        out_word += dst_pitch;
    }
}

// ---- 16-bit RGB, lane = pixel column, 4 rows per lane (BASELINE config 4) ------------------------------
// The samplers of r360_fast_u16.cuh; a finished pixel is three 16-bit values, two lanes (a pixel pair) own three
// 32-bit words of the row: the even lane stores words 0 and 1 (it fetches the odd lane's first word with a
// shuffle), the odd lane word 2.
template <typename TOut>
__device__ __forceinline__ uint32_t out_bits16(float a);
template <>
__device__ __forceinline__ uint32_t out_bits16<uint16_t>(float a) { return (uint32_t)Finish<uint16_t, uint16_t>::run(a); }
template <>
__device__ __forceinline__ uint32_t out_bits16<__half>(float a) { return (uint32_t)__half_as_ushort(Finish<uint16_t, __half>::run(a)); }

template <int INTERP, typename TOut, int NF>
__device__ __forceinline__ void column_rows_u16c3(const TilePlan* plan, const float* rowc_warp, uint32_t bias, uint32_t pitch,
                                                  uint32_t fstride, float s, double dlane, double drow, unsigned char* out_row,
                                                  long long dst_pitch, long long dst_fstride, int lane) {
    const double axi = plan->ax[1], ayi = plan->ay[1], axj = plan->ax[2], ayj = plan->ay[2];
    double ax_r = fma(axi, dlane, fma(axj, drow, plan->ax[0]));
    double ay_r = fma(ayi, dlane, fma(ayj, drow, plan->ay[0]));
    const bool odd = (lane & 1) != 0;
    // even lane: words 0 and 1 of its pair at byte 12 * (lane / 2); odd lane: word 2
    unsigned char* out_a = out_row + 12 * (lane >> 1) + (odd ? 8 : 0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float* rq = rowc_warp + q * 12;
        const float4 c0 = *reinterpret_cast<const float4*>(rq), c1 = *reinterpret_cast<const float4*>(rq + 4),
                     c2 = *reinterpret_cast<const float4*>(rq + 8);
        float dx, dy;
        residual_xy(c0, c1, c2, s, dx, dy);
        const float sx = __double2float_rn(ax_r + (double)dx), sy = __double2float_rn(ay_r + (double)dy);
        ax_r += axj; ay_r += ayj;
        float acc[NF][3];
        if constexpr (INTERP == kCubic) {
            const BicubicColPrepU16 pp = bicubic_col_prep_u16c3(bias, pitch, g_tables.cubic_1d, round_bits(sx), round_bits(sy));
#pragma unroll
            for (int f = 0; f < NF; ++f) bicubic_col_taps_u16c3(pp, pitch, (uint32_t)f * fstride, acc[f]);
        } else {
            const BilinearColPrepU16 pp = bilinear_col_prep_u16c3(bias, pitch, round_bits(sx), round_bits(sy));
#pragma unroll
            for (int f = 0; f < NF; ++f) bilinear_col_taps_u16c3(pp, pitch, (uint32_t)f * fstride, acc[f]);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const uint32_t a = out_bits16<TOut>(acc[f][0]) | (out_bits16<TOut>(acc[f][1]) << 16), b = out_bits16<TOut>(acc[f][2]);
            const uint32_t a_other = __shfl_xor_sync(0xffffffffu, a, 1);
            const uint32_t first = odd ? ((a >> 16) | (b << 16)) : a;          // word 2 | word 0
            const uint32_t second = b | (a_other << 16);                        // word 1 (even lanes only)
            unsigned char* o = out_a + f * dst_fstride;
            asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(o), "r"(first));
            asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %2, 0;\n@p st.global.cs.b32 [%0], %1;\n}"
                         ::"l"(o + 4), "r"(second), "r"((uint32_t)odd));
        }
        out_a += dst_pitch;
    }
}

// ---- 8-bit RGB, lane = a PAIR of adjacent pixel columns, 16 pairs x 2 rows per warp pass ----------------
// At the 2:1 minification of an 8K -> 1600 px view neighbouring pixels are 6 source bytes apart: with one pixel per
// lane a tap load of the warp spreads over 48 words = two shared-memory wavefronts at best (2.4 measured).  With two
// adjacent pixels per lane the lanes of a row are 12 bytes = three words apart -- an odd word stride, so 16 lanes
// hit 16 different banks -- and the second half of the warp works on the next tile row, two source rows further
// down, which with patch pitches that are odd multiples of 32 bytes lands 16 banks away: the two halves
// interleave (tools/pipe_probe.cu: 1 wavefront for a 12-byte lane stride against 2 for 6 bytes).  Same samplers,
// same arithmetic; the finished 4-pixel group of two lanes leaves as three 32-bit words.
// Measured on B200 (16 x 8K frames -> 12 views, two frames per item; shared-memory wavefronts per launch / time):
// bilinear 363 M / 1.88 ms with one pixel per lane, 352 M / 1.82 ms with pairs: used for bilinear.  Bicubic goes the
// other way -- 828 M / 3.48 ms with one pixel per lane on 128-byte pitches against 936 M / 3.70 ms with pairs (its four
// words per tap row make the two half-warps collide more often than the wider stride saves) -- and stays there.
template <int INTERP, int NF>
__device__ __forceinline__ void pair_rows_u8c3(const TilePlan* plan, const float* rowc_warp, uint32_t bias, uint32_t pitch,
                                               uint32_t tab, uint32_t fstride, int lane, int warp, unsigned char* out_rows,
                                               long long dst_pitch, long long dst_fstride) {
    const int half = lane >> 4, pair = lane & 15;              // tile row within the pass, pixel pair within the row
    const bool odd = (pair & 1) != 0;
    const double axi = plan->ax[1], ayi = plan->ay[1], axj = plan->ax[2], ayj = plan->ay[2];
    const float s0 = (float)(4 * pair - (kTile - 1)) * (1.0f / (kTile - 1)), s1 = (float)(4 * pair + 2 - (kTile - 1)) * (1.0f / (kTile - 1));
    const double di0 = (double)(2 * pair);
    // even pair lane: words 0 and 1 of its 4-pixel group at byte 12 * (pair / 2); odd pair lane: word 2
    unsigned char* out_a = out_rows + (long long)half * dst_pitch + 12 * (pair >> 1) + (odd ? 8 : 0);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int row = 2 * pass + half;                        // 0..3 within the warp's four rows
        const float* rq = rowc_warp + row * 12;
        const float4 c0 = *reinterpret_cast<const float4*>(rq), c1 = *reinterpret_cast<const float4*>(rq + 4),
                     c2 = *reinterpret_cast<const float4*>(rq + 8);
        const double djl = (double)(warp * 4 + row);
        const double ax_b = fma(axj, djl, plan->ax[0]), ay_b = fma(ayj, djl, plan->ay[0]);
        uint32_t px[2][NF];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float dx, dy;
            residual_xy(c0, c1, c2, e ? s1 : s0, dx, dy);
            const float sx = __double2float_rn(fma(axi, di0 + (double)e, ax_b) + (double)dx);
            const float sy = __double2float_rn(fma(ayi, di0 + (double)e, ay_b) + (double)dy);
            if constexpr (INTERP == kCubic) {
                const BicubicPrep pp = bicubic_prep_u8c3(bias, pitch, tab, round_bits(sx), round_bits(sy));
#pragma unroll
                for (int f = 0; f < NF; ++f) px[e][f] = bicubic_taps_u8c3(pp, pitch, (uint32_t)f * fstride);
            } else {
                const BilinearPrep pp = bilinear_prep_u8c3(bias, pitch, round_bits(sx), round_bits(sy));
#pragma unroll
                for (int f = 0; f < NF; ++f) px[e][f] = bilinear_taps_u8c3(pp, pitch, (uint32_t)f * fstride);
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            // group of four pixels A0 A1 (even lane) B0 B1 (odd lane), each R | G << 8 | B << 16:
            //   word 0 = A0 | A1 << 24, word 1 = A1 >> 8 | B0 << 16, word 2 = B0 >> 16 | B1 << 8
            const uint32_t b0 = __shfl_down_sync(0xffffffffu, px[0][f], 1);
            const uint32_t first = odd ? ((px[0][f] >> 16) | (px[1][f] << 8)) : (px[0][f] | (px[1][f] << 24));
            const uint32_t second = (px[1][f] >> 8) | (b0 << 16);
            unsigned char* o = out_a + f * dst_fstride;
            asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(o), "r"(first));
            asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %2, 0;\n@p st.global.cs.b32 [%0], %1;\n}"
                         ::"l"(o + 4), "r"(second), "r"((uint32_t)odd));
        }
        out_a += 2 * dst_pitch;
    }
}

template <int INTERP, typename TIn, typename TOut, int FR, int TEAMS>
__global__ void __launch_bounds__(tiled_threads(TEAMS), tiled_min_blocks((int)sizeof(TIn), INTERP, TEAMS)) remap_tiled_kernel(const __grid_constant__ TiledParams P,
                                                                    const __grid_constant__ TensorMaps maps) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [kSlots]
    uint64_t* empty = full + kSlots;                               // [kSlots]
    SlotInfo* slots = reinterpret_cast<SlotInfo*>(smem + kSmemSlots);            // [kSlots]
    float* rowc_all = reinterpret_cast<float*>(smem + kSmemRowc);                // [team][32 rows][12]
    TilePlan* planbuf = reinterpret_cast<TilePlan*>(smem + kSmemPlans);          // [kSlots]
    unsigned char* table = smem + kTiledFixedSmem;
    unsigned char* stage_all = table + table_bytes(P.use_table);   // [team][FR][out_stage_bytes]
    constexpr int n_teams = TEAMS;
    unsigned char* ring = stage_all + n_teams * FR * P.out_stage_bytes;

    constexpr bool kFastU8 = std::is_same<TIn, uint8_t>::value && std::is_same<TOut, uint8_t>::value &&
                             (INTERP == kLinear || INTERP == kCubic);
    constexpr bool kFastLanczosU8 = std::is_same<TIn, uint8_t>::value && std::is_same<TOut, uint8_t>::value && INTERP == kLanczos4;
    // nearest on 8-bit sources with 3 channels (views) or 1 channel (the dual-fisheye tool's masks, DF:2031-2043)
    constexpr bool kFastNearestU8 = std::is_same<TIn, uint8_t>::value && std::is_same<TOut, uint8_t>::value && INTERP == kNearest;
    constexpr bool kFastU16 = std::is_same<TIn, uint16_t>::value && (INTERP == kLinear || INTERP == kCubic);
    const int tid = threadIdx.x;
    const int n_tiles = P.tiles_x * P.tiles_y;
    const int n_fblocks = (P.n_groups + FR - 1) / FR;
    const int total = n_fblocks * P.n_order;                        // the host keeps this below 2^31

    if (tid == 0) {
        for (int q = 0; q < kSlots; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], kConsumerWarps); }
        fence_mbar_init();
    }
    if (P.use_table == 2) {
        // lanczos4: 128-byte entries (eight 16-byte tap rows); row ky of entry e sits in slot (ky + e) mod 8 so that
        // the 32 entries a warp reads for the same ky spread over the eight bank groups
        const int4* src = reinterpret_cast<const int4*>(g_tables.lanczos_fixed);
        for (int q = tid; q < kLanczosTableBytes / 16; q += blockDim.x) {
            const int e = q >> 3, ky = q & 7;
            reinterpret_cast<int4*>(table)[e * 8 + ((ky + e) & 7)] = __ldg(src + q);
        }
    } else if (P.use_table) {
        const int4* src = reinterpret_cast<const int4*>(g_tables.cubic_fixed);
        // Two planes (tap rows 0,1 | rows 2,3) with a 16-byte entry stride: a warp's 32 random entries
        // then spread over all 8 bank groups instead of the 4 a 32-byte stride would reach.
        for (int q = tid; q < kTableBytes / 16; q += blockDim.x)
            reinterpret_cast<int4*>(table)[table_entry_index((uint32_t)q >> 6, ((uint32_t)q >> 1) & 31u) + (q & 1) * (kTableBytes / 32)] = __ldg(src + q);
    }
    __syncthreads();

    if (tid >= n_teams * kTeamThreads) {
        // ================= producer warp: keep it short, it is the serial part of the pipeline ======
        const int lane = tid - n_teams * kTeamThreads;
        PatchRing ringst;                // identical in every lane
        int oldest = 0, k = 0;
        const uint64_t policy = l2_policy_by_kind(P.l2_policy);
        // Items are walked in the plan's source-row order (all views interleaved): the blocks of the grid then read
        // from one band of source rows at a time, so a frame's bytes come out of DRAM once and every view that
        // overlaps the band finds them in L2.  Item = (entry of the order list, frame block); the (view, tile) pair
        // is prefetched two items ahead, the tile's geometry one item ahead.
        // Items are CLAIMED from a counter in global memory, three ahead, not assigned by a grid stride: whatever the
        // blocks' speeds, the items in flight are always one contiguous stretch of the list, so neighbouring tiles are
        // loaded by different SMs at the same moment and share their lines through L2.  Measured on B200 (16 x 8K
        // frames, 12 views, DRAM reads per launch): with a static grid stride the blocks drift apart over their ~400
        // items each and the stretch in flight grows to many bands -- bilinear (296 blocks) 3.6 GB, a single
        // four-frame group alone 0.41 GB; claimed items: see profiles/README.md.  (One contiguous run of the list per
        // block was 9.1 GB and 98 against 145 Gpix/s.)
        const int item_end = total;
        auto claim = [&]() -> int {
            int v = 0;
            if (lane == 0) v = (int)atomicAdd(P.work_counter, 1u);
            return __shfl_sync(0xffffffffu, v, 0);
        };
        int item = claim(), item_n1 = claim(), item_n2 = claim();
        int2 vt = make_int2(0, 0), vt_next = make_int2(0, 0);       // (view, tile) of this item / the next one
        int gb = item / P.n_order, gb_next = item_n1 / P.n_order;
        if (item < item_end) vt = __ldg(P.order + (item - gb * P.n_order));
        if (item_n1 < item_end) vt_next = __ldg(P.order + (item_n1 - gb_next * P.n_order));
        // geometry of the item about to be processed: lanes 0 and 1 hold the record's last two
        // 16-byte pieces (py0 rows xb0 row_bytes | pitch mode_slot - -)
        int4 geo = make_int4(0, 0, 0, 0);
        const TilePlan* gp = P.plans + (long long)vt.x * n_tiles + vt.y;
        if (item < item_end && lane < 2) geo = __ldg(reinterpret_cast<const int4*>(gp) + 21 + lane);
        for (; item < item_end;) {
            const int py0 = __shfl_sync(0xffffffffu, geo.x, 0), rows_needed = __shfl_sync(0xffffffffu, geo.y, 0);
            const int xb0 = __shfl_sync(0xffffffffu, geo.z, 0), row_bytes = __shfl_sync(0xffffffffu, geo.w, 0);
            const int pitch = __shfl_sync(0xffffffffu, geo.x, 1), mode_slot = __shfl_sync(0xffffffffu, geo.y, 1);
            const TilePlan* gp_cur = gp;
            const int cur_gb = gb, cur_v = vt.x;
            const int cur_tj = vt.y / P.tiles_x, cur_ti = vt.y - cur_tj * P.tiles_x;
            // advance: the next item's geometry and the (view, tile) pair of the one after it
            {
                item = item_n1; item_n1 = item_n2;
                vt = vt_next; gb = gb_next;
                gb_next = item_n1 / P.n_order;
                if (item_n1 < item_end) vt_next = __ldg(P.order + (item_n1 - gb_next * P.n_order));
                item_n2 = item_n1 < item_end ? claim() : item_end;
            }
            gp = P.plans + (long long)vt.x * n_tiles + vt.y;
            if (item < item_end && lane < 2) geo = __ldg(reinterpret_cast<const int4*>(gp) + 21 + lane);

            const int mode = mode_slot & 0xff, src_slot = (mode_slot >> 8) & 0xff, wbox = mode_slot >> 16;
            if (mode == kModeFallback) continue;                     // remap_fallback_kernel owns this tile
            const bool staged = mode == kModeFast || mode == kModeFastRows || mode == kModeFastSeam;
            const int rows = staged ? rows_needed : 0;
            const int srows = mode == kModeFast ? staged_rows(rows) : rows;      // rows written to the ring
            const int need1 = (srows * pitch + 127) & ~127;                       // one frame's patch
            const uint32_t patch_bytes1 = !staged ? 0u : (uint32_t)(mode == kModeFast ? srows * pitch : rows * row_bytes);
            // frames of this item: all FR in one slot when they fit, else one slot per frame
            const int g_first = cur_gb * FR;
            const int avail = min(FR, P.n_groups - g_first);
            const bool multi = FR > 1 && avail == FR && FR * need1 <= P.multi_budget;
            const int nf = multi ? FR : 1, n_sub = multi ? 1 : avail;
            const int i0 = cur_ti * kTile, j0 = cur_tj * kTile;
            for (int sub = 0; sub < n_sub; ++sub, ++k) {
                const int slot = k % kSlots;
                const int g0 = g_first + sub;                                     // first frame of the slot
                const int need = need1 * nf;
                // ---- find room: wait for the oldest slots to be released until the patches fit ---------
                int off = 0, charge = 0;
                while ((k - oldest) >= kSlots || !ringst.try_alloc(need, P.ring_bytes, off, charge)) {
                    mbar_wait_relaxed(&empty[oldest % kSlots], (oldest / kSlots) & 1);
                    ringst.release(slots[oldest % kSlots].size, P.ring_bytes);
                    ++oldest;
                }
                unsigned char* patch = ring + off;
                __syncwarp();                                              // slots[] reads above are done
                if (lane == 0) {
                    SlotInfo si;
                    si.patch_saddr = smem_u32(patch);
                    si.bias = 0;
                    if (kFastU8) {
                        si.bias = patch_bias_u8c3(si.patch_saddr, pitch, xb0, py0);
                        if (INTERP == kCubic) si.bias -= 3u + (uint32_t)pitch;
                    }
                    if (kFastLanczosU8) si.bias = patch_bias_u8c3(si.patch_saddr, pitch, xb0, py0) - (9u + 3u * (uint32_t)pitch);
                    if (kFastNearestU8) si.bias = patch_bias_nearest_u8(si.patch_saddr, pitch, xb0, py0, (uint32_t)P.channels);
                    if (kFastU16) {
                        si.bias = patch_bias_u16c3(si.patch_saddr, pitch, xb0, py0);
                        if (INTERP == kCubic) si.bias -= 6u + (uint32_t)pitch;
                    }
                    si.size = charge; si.mode = mode; si.pitch = pitch; si.xb0 = xb0; si.py0 = py0;
                    si.full_tile = (i0 + kTile <= P.dst.width && j0 + kTile <= P.dst.height) ? 1 : 0;
                    si.i0 = i0; si.j0 = j0;
                    si.dst_off = ((long long)g0 * P.n_views + cur_v) * P.dst.image_stride;
                    si.dst_tile = si.dst_off + (long long)j0 * P.dst.pitch + (long long)i0 * P.channels * (int)sizeof(TOut);
                    si.nf = nf; si.fstride = (uint32_t)need1;
                    slots[slot] = si;
                    mbar_expect_tx(&full[slot], (uint32_t)sizeof(TilePlan) + patch_bytes1 * (uint32_t)nf);     // arrive + expect
                    bulk_g2s(&planbuf[slot], gp_cur, (uint32_t)sizeof(TilePlan), &full[slot]);
                }
                __syncwarp();
                if (mode == kModeFast) {
                    // a few tensor boxes per frame: 32-row boxes first, then 8-row boxes
                    const int n32 = rows / 32, nb = n32 + ((rows % 32) + 7) / 8;
                    for (int t = lane; t < nb * nf; t += 32) {
                        const int f = t / nb, bx = t - f * nb;
                        const int r0 = bx < n32 ? bx * 32 : n32 * 32 + (bx - n32) * 8;
                        const CUtensorMap* tm = &maps.m[wbox * kNumBoxHeights + (bx < n32 ? 0 : 1)];
                        tensor_g2s_3d(patch + f * need1 + r0 * pitch, tm, xb0 >> 2, py0 + r0, (g0 + f) * P.n_lenses + src_slot,
                                      &full[slot], policy);
                    }
                } else if (mode == kModeFastRows) {
                    for (int t = lane; t < rows * nf; t += 32) {
                        const int f = t / rows, r = t - f * rows;
                        const unsigned char* img = P.src.data + ((long long)(g0 + f) * P.n_lenses + src_slot) * P.src.image_stride;
                        const int sy = min(max(py0 + r, 0), P.src.height - 1);            // pole rows replicate
                        bulk_g2s(patch + f * need1 + r * pitch, img + (long long)sy * P.src.pitch + xb0, (uint32_t)row_bytes, &full[slot]);
                    }
                } else if (mode == kModeFastSeam) {
                    // the patch spans the seam: its first bytes come from the end of the image row, the rest
                    // from its beginning (both pieces are multiples of 16 bytes)
                    const int row_total = P.src.width * P.channels * (int)sizeof(TIn);
                    const int start = xb0 < 0 ? xb0 + row_total : xb0;                   // wrapped first byte
                    const int len1 = min(row_bytes, row_total - start);
                    for (int t = lane; t < rows * nf; t += 32) {
                        const int f = t / rows, r = t - f * rows;
                        const unsigned char* img = P.src.data + ((long long)(g0 + f) * P.n_lenses + src_slot) * P.src.image_stride;
                        const int sy = min(max(py0 + r, 0), P.src.height - 1);
                        const unsigned char* row = img + (long long)sy * P.src.pitch;
                        unsigned char* dst_row = patch + f * need1 + r * pitch;
                        bulk_g2s(dst_row, row + start, (uint32_t)len1, &full[slot]);
                        if (len1 < row_bytes) bulk_g2s(dst_row + len1, row, (uint32_t)(row_bytes - len1), &full[slot]);
                    }
                }
            }
        }
        // ---- end of stream: one exit record per team (slots k, k + 1, ... belong to teams k mod n_teams, ...) ----
        for (int e = 0; e < n_teams; ++e, ++k) {
            const int slot = k % kSlots;
            while ((k - oldest) >= kSlots) {
                mbar_wait_relaxed(&empty[oldest % kSlots], (oldest / kSlots) & 1);
                ++oldest;
            }
            __syncwarp();
            if (lane == 0) {
                slots[slot].mode = kModeExit;
                mbar_arrive(&full[slot]);
            }
            __syncwarp();
        }
        return;
    }

    // ================= consumer warps ================================================================
    const int team = tid / kTeamThreads;
    const int ttid = tid - team * kTeamThreads;
    const int warp = ttid >> 5, lane = tid & 31;
    const int jl = ttid >> 3;                // tile row of this thread (4 rows per warp)
    const int il0 = (ttid & 7) * 4;          // first of its 4 pixels
    const int row_out_bytes = kTile * P.channels * (int)sizeof(TOut);
    const float s0 = (float)(2 * il0 - (kTile - 1)) * (1.0f / (kTile - 1));
    const float ds = 2.0f / (kTile - 1);
    const float trow = (float)(2 * jl - (kTile - 1)) * (1.0f / (kTile - 1));
    const double dil0 = (double)il0, djl = (double)jl;
    float* rowc = rowc_all + team * (kTile * 12);
    float* rc = rowc + jl * 12;
    // residual-coefficient tasks of this lane: coefficient c0 = lane & 7 always, c1 = c0 + 8 for c0 < 4
    const int ctask0 = lane & 7, ctask1 = ctask0 + 8;
    const int koff0 = ctask0 < 6 ? ctask0 : 36 + ctask0 - 6;             // float offset into rx[36] ry[36]
    const int koff1 = 36 + ctask1 - 6;
    // this lane's share of the warp's 4 output rows when they leave as 16-byte chunks
    const int chunks_per_row = row_out_bytes >> 4;                        // 6 for 8-bit RGB
    const int st_r = lane / chunks_per_row, st_c = lane - st_r * chunks_per_row;
    const int st_dr = 32 / chunks_per_row, st_dc = 32 - st_dr * chunks_per_row;
    const uint32_t full_s = smem_u32(full), empty_s = smem_u32(empty);
    unsigned char* stage = stage_all + team * FR * P.out_stage_bytes;     // [FR][tile]; a warp only touches its own 4 rows
    TOut* stage_row = reinterpret_cast<TOut*>(stage) + (jl * kTile + il0) * P.channels;
    const int stage_fstride = P.out_stage_bytes / (int)sizeof(TOut);      // elements between the frames of the stage
    // per-lane constants of the lane-per-column path (kept in registers across tiles: the int -> double
    // conversions run on the slow conversion unit)
    const uint32_t col_tab = smem_u32(table);
    const float col_s = (float)(2 * lane - (kTile - 1)) * (1.0f / (kTile - 1));
    const double col_dlane = (double)lane, col_drow = (double)(warp * 4);
    for (int k = team;; k += n_teams) {
        const int slot = k % kSlots;
        mbar_wait_s(full_s + slot * 8, (k / kSlots) & 1);
        const SlotInfo* si = &slots[slot];
        const int mode = si->mode;
        if (mode == kModeExit) break;
        const int nf = si->nf;
        const uint32_t fstride = si->fstride;
        const TilePlan* plan = &planbuf[slot];
        const int cmap = plan->pad[1];           // > 0: the tile has a per-pixel map instead of a polynomial
        if (mode != kModeFill && cmap == 0) {
            // the 12 residual coefficients of this lane's row, spread over the 8 lanes that share the row
            const float* K = plan->rx + koff0;
            float a = K[30];
#pragma unroll
            for (int l = 4; l >= 0; --l) a = fmaf(a, trow, K[l * 6]);
            rc[ctask0] = a;
            if (ctask0 < 4) {
                const float* K1 = plan->rx + koff1;
                float b = K1[30];
#pragma unroll
                for (int l = 4; l >= 0; --l) b = fmaf(b, trow, K1[l * 6]);
                rc[ctask1] = b;
            }
        }
        __syncwarp();
        if constexpr (kFastU8) {
            // lane-per-column tiles: whole tile inside the image, vector stores possible
            if (mode != kModeFill && mode != kModeFastSeam && cmap == 0 && P.channels == 3 && si->full_tile && P.bulk_store_ok) {
                const int m = lane & 3;
                unsigned char* out_word = P.dst.data + si->dst_tile + (long long)(warp * 4) * P.dst.pitch + ((lane >> 2) * 3 + m) * 4;
                if constexpr (INTERP == kLinear) {
                    unsigned char* out_rows = P.dst.data + si->dst_tile + (long long)(warp * 4) * P.dst.pitch;
                    if (FR > 1 && nf == FR)
                        pair_rows_u8c3<INTERP, FR>(plan, rowc + warp * 48, si->bias, (uint32_t)si->pitch, col_tab, fstride, lane, warp,
                                                   out_rows, P.dst.pitch, P.dst_fstride);
                    else
                        pair_rows_u8c3<INTERP, 1>(plan, rowc + warp * 48, si->bias, (uint32_t)si->pitch, col_tab, fstride, lane, warp,
                                                  out_rows, P.dst.pitch, P.dst_fstride);
                } else if (FR > 1 && nf == FR)
                    column_rows_u8c3<INTERP, FR>(plan, rowc + warp * 48, si->bias, (uint32_t)si->pitch, col_tab, fstride, col_s,
                                                 col_dlane, col_drow, out_word, P.dst.pitch, P.dst_fstride, m);
                else
                    column_rows_u8c3<INTERP, 1>(plan, rowc + warp * 48, si->bias, (uint32_t)si->pitch, col_tab, fstride, col_s,
                                                col_dlane, col_drow, out_word, P.dst.pitch, P.dst_fstride, m);
                __syncwarp();
                if (lane == 0) mbar_arrive_s(empty_s + slot * 8);
                continue;
            }
        }
        if constexpr (kFastU16) {
            // lane-per-column tiles of 16-bit RGB sources (4-byte aligned destination rows)
            if (mode != kModeFill && mode != kModeFastSeam && cmap == 0 && P.channels == 3 && si->full_tile && (P.dst.pitch & 3) == 0 &&
                (P.dst.image_stride & 3) == 0) {
                unsigned char* out_row = P.dst.data + si->dst_tile + (long long)(warp * 4) * P.dst.pitch;
                if (FR > 1 && nf == FR)
                    column_rows_u16c3<INTERP, TOut, FR>(plan, rowc + warp * 48, si->bias, (uint32_t)si->pitch, fstride, col_s, col_dlane,
                                                        col_drow, out_row, P.dst.pitch, P.dst_fstride, lane);
                else
                    column_rows_u16c3<INTERP, TOut, 1>(plan, rowc + warp * 48, si->bias, (uint32_t)si->pitch, fstride, col_s, col_dlane,
                                                       col_drow, out_row, P.dst.pitch, P.dst_fstride, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive_s(empty_s + slot * 8);
                continue;
            }
        }
        if (mode == kModeFill) {
            for (int f = 0; f < nf; ++f)
                for (int q = 0; q < 4 * P.channels; ++q) stage_row[f * stage_fstride + q] = Finish<TIn, TOut>::run(P.border_value);
        } else {
            // residual polynomial in float32, affine part in float64 (absolute, exact to ~1e-10 px);
            // the final double -> float conversion IS the float32 cast of the map cv2.remap would get.
            // A tile on the +-180 degree seam (kModeFastSeam) wraps its coordinate into (-0.5, W - 0.5] before
            // that cast; the taps of a wrapped pixel sit one image width further along the (unwrapped) patch:
            // woff[q] is that byte offset, zero everywhere else.
            const float4 c0 = *reinterpret_cast<const float4*>(rc), c1 = *reinterpret_cast<const float4*>(rc + 4),
                         c2 = *reinterpret_cast<const float4*>(rc + 8);
            const double axi = plan->ax[1], ayi = plan->ay[1];
            double ax_q = fma(axi, dil0, fma(plan->ax[2], djl, plan->ax[0]));
            double ay_q = fma(ayi, dil0, fma(plan->ay[2], djl, plan->ay[0]));
            const bool seam = mode == kModeFastSeam;
            const int row_total = P.src.width * P.channels * (int)sizeof(TIn);
            const double period32 = 32.0 * P.src.width;
            float sxf[4], syf[4];
            int woff[4];
            // a tile without a polynomial reads its four (32 x, 32 y) pairs from the coordinate pool instead
            const double2* cm = cmap > 0 ? P.coords + ((long long)(cmap - 1) * (kTile * kTile) + jl * kTile + il0) : nullptr;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double sx, sy;
                if (cm != nullptr) {
                    const double2 c = __ldg(cm + q);
                    sx = c.x; sy = c.y;
                } else {
                    const float s = fmaf((float)q, ds, s0);
                    float dx, dy;
                    residual_xy(c0, c1, c2, s, dx, dy);
                    sx = ax_q + (double)dx;
                    sy = ay_q + (double)dy;
                    ax_q += axi; ay_q += ayi;
                }
                woff[q] = 0;
                if (seam) {
                    if (sx > period32 - 16.0) { sx -= period32; woff[q] = row_total; }
                    else if (sx <= -16.0) { sx += period32; woff[q] = -row_total; }
                }
                sxf[q] = __double2float_rn(sx);
                syf[q] = __double2float_rn(sy);
            }
            bool done = false;
            if constexpr (kFastU8) {
                if (P.channels == 3) {
                    const uint32_t bias = si->bias, pitch = (uint32_t)si->pitch;
                    if constexpr (INTERP == kLinear) {
                        BilinearPrep pp[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) pp[q] = bilinear_prep_u8c3(bias + (uint32_t)woff[q], pitch, round_bits(sxf[q]), round_bits(syf[q]));
                        auto frames = [&](auto nfc) {
                            constexpr int NF = decltype(nfc)::value;
                            uint32_t w[NF][3];          // every frame sampled before the first stage store (which the
#pragma unroll                                          // compiler must assume to alias the patches)
                            for (int f = 0; f < NF; ++f) {
                                uint32_t px[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) px[q] = bilinear_taps_u8c3(pp[q], pitch, (uint32_t)f * fstride);
                                pack4_rgb(px[0], px[1], px[2], px[3], w[f][0], w[f][1], w[f][2]);
                            }
#pragma unroll
                            for (int f = 0; f < NF; ++f) {
                                uint32_t* o = reinterpret_cast<uint32_t*>(stage_row + f * stage_fstride);
                                o[0] = w[f][0]; o[1] = w[f][1]; o[2] = w[f][2];
                            }
                        };
                        if (FR > 1 && nf == FR) frames(std::integral_constant<int, FR>{});
                        else frames(std::integral_constant<int, 1>{});
                    } else {
                        // bicubic tiles off the lane-per-column path (partial tiles, unaligned destinations)
                        const uint32_t tab = smem_u32(table);
                        for (int f = 0; f < nf; ++f) {
                            uint32_t px[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                px[q] = bicubic_u8c3(bias + (uint32_t)woff[q] + (uint32_t)f * fstride, pitch, tab, round_bits(sxf[q]), round_bits(syf[q]));
                            uint32_t w0, w1, w2;
                            pack4_rgb(px[0], px[1], px[2], px[3], w0, w1, w2);
                            uint32_t* o = reinterpret_cast<uint32_t*>(stage_row + f * stage_fstride);
                            o[0] = w0; o[1] = w1; o[2] = w2;
                        }
                    }
                    done = true;
                }
            }
            if constexpr (kFastLanczosU8) {
                if (P.channels == 3) {
                    const uint32_t pitch = (uint32_t)si->pitch;
                    for (int f = 0; f < nf; ++f) {
                        const uint32_t bias = si->bias + (uint32_t)f * fstride;
                        uint32_t px[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            px[q] = P.use_table == 2
                                        ? lanczos4_u8c3<true>(bias + (uint32_t)woff[q], pitch, nullptr, smem_u32(table), round_bits(sxf[q]), round_bits(syf[q]))
                                        : lanczos4_u8c3<false>(bias + (uint32_t)woff[q], pitch, g_tables.lanczos_fixed, 0u, round_bits(sxf[q]), round_bits(syf[q]));
                        uint32_t w0, w1, w2;
                        pack4_rgb(px[0], px[1], px[2], px[3], w0, w1, w2);
                        uint32_t* o = reinterpret_cast<uint32_t*>(stage_row + f * stage_fstride);
                        o[0] = w0; o[1] = w1; o[2] = w2;
                    }
                    done = true;
                }
            }
            if constexpr (kFastNearestU8) {
                if (P.channels == 3 || P.channels == 1) {
                    const uint32_t pitch = (uint32_t)si->pitch;
                    uint32_t ux[4], uy[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { ux[q] = round_bits(sxf[q] * (1.0f / 32.0f)); uy[q] = round_bits(syf[q] * (1.0f / 32.0f)); }
                    for (int f = 0; f < nf; ++f) {
                        const uint32_t bias = si->bias + (uint32_t)f * fstride;
                        uint32_t px[4];
                        if (P.channels == 3) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) px[q] = nearest_u8c3(bias + (uint32_t)woff[q], pitch, ux[q], uy[q]);
                            uint32_t w0, w1, w2;
                            pack4_rgb(px[0], px[1], px[2], px[3], w0, w1, w2);
                            uint32_t* o = reinterpret_cast<uint32_t*>(stage_row + f * stage_fstride);
                            o[0] = w0; o[1] = w1; o[2] = w2;
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) px[q] = nearest_u8c1(bias + (uint32_t)woff[q], pitch, ux[q], uy[q]);
                            *reinterpret_cast<uint32_t*>(stage_row + f * stage_fstride) = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
                        }
                    }
                    done = true;
                }
            }
            if constexpr (kFastU16) {
                if (P.channels == 3) {
                    const uint32_t bias = si->bias, pitch = (uint32_t)si->pitch;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if constexpr (INTERP == kLinear) {
                            const BilinearPrepU16 pp = bilinear_prep_u16c3(bias + (uint32_t)woff[q], pitch, round_bits(sxf[q]), round_bits(syf[q]));
                            for (int f = 0; f < nf; ++f) {
                                float acc[3];
                                bilinear_taps_u16c3(pp, pitch, (uint32_t)f * fstride, acc);
#pragma unroll
                                for (int c = 0; c < 3; ++c) stage_row[f * stage_fstride + q * 3 + c] = Finish<TIn, TOut>::run(acc[c]);
                            }
                        } else {
                            const BicubicPrepU16 pp = bicubic_prep_u16c3(bias + (uint32_t)woff[q], pitch, g_tables.cubic_1d, round_bits(sxf[q]), round_bits(syf[q]));
                            for (int f = 0; f < nf; ++f) {
                                float acc[3];
                                bicubic_taps_u16c3(pp, pitch, (uint32_t)f * fstride, acc);
#pragma unroll
                                for (int c = 0; c < 3; ++c) stage_row[f * stage_fstride + q * 3 + c] = Finish<TIn, TOut>::run(acc[c]);
                            }
                        }
                    }
                    done = true;
                }
            }
            if (!done) {
                unsigned char* patch = smem + (si->patch_saddr - smem_u32(smem));
#pragma unroll 1
                for (int f = 0; f < nf; ++f) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const PatchTaps<TIn> taps{patch + f * fstride, si->pitch, si->xb0 - woff[q], si->py0, P.channels};
                        sample_pixel<INTERP, TIn, TOut>(taps, P.channels, P.src.width, P.src.height, P.border_value,
                                                        sxf[q] * (1.0f / 32.0f), syf[q] * (1.0f / 32.0f),
                                                        stage_row + f * stage_fstride + q * P.channels);
                    }
                }
            }
        }
        // ---- this warp's 4 rows of every frame leave as 16-byte vector stores; the patches are released ----
        const bool vec_store = si->full_tile && P.bulk_store_ok;
        const long long dst_tile = si->dst_tile, dst_off = si->dst_off;
        const int i0 = si->i0, j0 = si->j0;
        __syncwarp();
        if (lane == 0) mbar_arrive_s(empty_s + slot * 8);  // all 8 warps arrived -> the producer may reuse the bytes
        for (int f = 0; f < nf; ++f) {
            const unsigned char* stage_f = stage + f * P.out_stage_bytes;
            if (vec_store) {
                unsigned char* dst_rows = P.dst.data + dst_tile + f * P.dst_fstride + (long long)(warp * 4) * P.dst.pitch;
                const unsigned char* src_rows = stage_f + warp * 4 * row_out_bytes;
                // chunk e = lane, lane + 32, ... of the 4 * chunks_per_row chunks; (r, c) advanced without dividing
                for (int r = st_r, c = st_c; r < 4;) {
                    st_global_streaming(dst_rows + (long long)r * P.dst.pitch + c * 16,
                                        *reinterpret_cast<const int4*>(src_rows + r * row_out_bytes + c * 16));
                    r += st_dr; c += st_dc;
                    if (c >= chunks_per_row) { c -= chunks_per_row; ++r; }
                }
            } else {
                unsigned char* dst_base = P.dst.data + dst_off + f * P.dst_fstride;
                const int nelem = kTile * P.channels;
                for (int e = lane; e < 4 * nelem; e += 32) {
                    const int r = warp * 4 + e / nelem, c = e % nelem;
                    const int i = i0 + c / P.channels, j = j0 + r;
                    if (i < P.dst.width && j < P.dst.height)
                        reinterpret_cast<TOut*>(dst_base + (long long)j * P.dst.pitch)[(long long)i0 * P.channels + c] =
                            reinterpret_cast<const TOut*>(stage_f)[r * nelem + c];
                }
            }
        }
        __syncwarp();                              // the stage rows are free again
    }
}

// Debug twin: what the tiled kernel samples at, written as maps (r360_plan_coords).  One block per
// (tile, view); only fast tiles write (the caller pre-fills everything from the direct path).
struct TiledCoordParams {
    int out_w, out_h, tiles_x, tiles_y;
    double period32;        // 32 * source width (seam tiles wrap their x coordinate)
    const TilePlan* plans;
    const double2* coords;
    float* x32; float* y32; double* x64; double* y64; unsigned char* valid;
};

__global__ void __launch_bounds__(256) coords_tiled_kernel(const __grid_constant__ TiledCoordParams P) {
    __shared__ TilePlan plan;
    __shared__ float rowc[kTile * 12];
    const int tid = threadIdx.x, tile = blockIdx.x, v = blockIdx.y;
    const int i0 = (tile % P.tiles_x) * kTile, j0 = (tile / P.tiles_x) * kTile;
    const int4* gp = reinterpret_cast<const int4*>(P.plans + (long long)v * (P.tiles_x * P.tiles_y) + tile);
    if (tid < (int)(sizeof(TilePlan) / 16)) reinterpret_cast<int4*>(&plan)[tid] = __ldg(gp + tid);
    __syncthreads();
    const int mode = plan.mode_slot & 0xff;
    if (mode != kModeFast && mode != kModeFastRows && mode != kModeFastSeam) return;
    const double2* cm = plan.pad[1] > 0 ? P.coords + (long long)(plan.pad[1] - 1) * (kTile * kTile) : nullptr;
    for (int task = tid; task < kTile * 12; task += 256) {
        const int row = task / 12, c = task % 12;
        const float* K = (c < 6 ? plan.rx : plan.ry) + (c < 6 ? c : c - 6);
        const float t = (float)(2 * row - (kTile - 1)) * (1.0f / (kTile - 1));
        float a = K[30];
#pragma unroll
        for (int l = 4; l >= 0; --l) a = fmaf(a, t, K[l * 6]);
        rowc[row * 12 + c] = a;
    }
    __syncthreads();
    const int jl = tid >> 3, il0 = (tid & 7) * 4;
    const float* rc = rowc + jl * 12;
    const double bx = fma(plan.ax[2], (double)jl, plan.ax[0]), by = fma(plan.ay[2], (double)jl, plan.ay[0]);
    for (int q = 0; q < 4; ++q) {
        const int i = i0 + il0 + q, j = j0 + jl;
        if (i >= P.out_w || j >= P.out_h) continue;
        const float s = (float)(2 * (il0 + q) - (kTile - 1)) * (1.0f / (kTile - 1));
        float dx = rc[5], dy = rc[11];
#pragma unroll
        for (int kk = 4; kk >= 0; --kk) { dx = fmaf(dx, s, rc[kk]); dy = fmaf(dy, s, rc[6 + kk]); }
        double sx = fma(plan.ax[1], (double)(il0 + q), bx) + (double)dx;
        double sy = fma(plan.ay[1], (double)(il0 + q), by) + (double)dy;
        if (cm != nullptr) { const double2 c = cm[jl * kTile + il0 + q]; sx = c.x; sy = c.y; }
        if (mode == kModeFastSeam) {
            if (sx > P.period32 - 16.0) sx -= P.period32;
            else if (sx <= -16.0) sx += P.period32;
        }
        const long long o = ((long long)v * P.out_h + j) * P.out_w + i;
        P.x32[o] = __double2float_rn(sx) * (1.0f / 32.0f); P.y32[o] = __double2float_rn(sy) * (1.0f / 32.0f);
        P.x64[o] = sx * (1.0 / 32.0);
        P.y64[o] = sy * (1.0 / 32.0);
        if (P.valid) P.valid[o] = 1;
    }
}

// ---- fallback tiles: direct float64 path ---------------------------------------------------------
//
// One block per four rows of a listed (view, tile) and per run of `frames_per_block` frames; a warp owns one
// tile row at a time (lane = column), so the gathers of neighbouring lanes share lines and the 32 results of a row
// leave as one contiguous piece.  The kernel is latency-bound -- a float64 projection (~6k cycles of dependent
// arithmetic) and then taps gathered from L2 -- so a thread projects its pixel ONCE and samples it from four frames
// of the batch at a time: the loads of the four frames are independent and overlap.  Measured on B200, 8K bicubic,
// 508 fallback tiles x 8 frames: 340 us with one frame per thread and a projection per frame, see profiles/README.md.

struct FallbackParams {
    LaunchParams lp;               // lp.views is unused: views come from the device array
    const ViewDev* views;          // all views of the plan
    const int2* list;              // (view, tile)
    int tiles_x;
    int n_groups, frames_per_block;
};

constexpr int kFallbackThreads = 128;
constexpr int kFallbackRows = 4;                       // tile rows per block: one per warp
constexpr int kFallbackFrames = 4;                     // frames sampled together

template <int PROJ, int INTERP, typename TIn, typename TOut>
__global__ void __launch_bounds__(kFallbackThreads) remap_fallback_kernel(const __grid_constant__ FallbackParams F) {
    const LaunchParams& p = F.lp;
    constexpr int kParts = kTile / kFallbackRows;
    const int2 entry = F.list[blockIdx.x / kParts];
    const int part = blockIdx.x % kParts;
    const int v = entry.x, tile = entry.y;
    const int g_first = blockIdx.y * F.frames_per_block, g_end = min(g_first + F.frames_per_block, F.n_groups);
    const int i0 = (tile % F.tiles_x) * kTile, j0 = (tile / F.tiles_x) * kTile + part * kFallbackRows;
    const ViewDev view = F.views[v];
    const int i = i0 + (threadIdx.x & 31);
    if (i >= p.dst.width) return;
#pragma unroll 1
    for (int jl = threadIdx.x >> 5; jl < kFallbackRows; jl += kFallbackThreads / 32) {
        const int j = j0 + jl;
        if (j >= p.dst.height) break;
        double x, y;
        const bool ok = project_pixel<PROJ>(view, p.erp, p.lens, (double)i, (double)j, x, y);
        const float xf = (float)x, yf = (float)y;
        int g = g_first;
#pragma unroll 1
        for (; g + kFallbackFrames <= g_end; g += kFallbackFrames)
            direct_sample<PROJ, INTERP, TIn, TOut, kFallbackFrames>(p, view, g, v, i, j, xf, yf, ok);   // lp.view_base is 0 here
#pragma unroll 1
        for (; g < g_end; ++g)
            direct_sample<PROJ, INTERP, TIn, TOut, 1>(p, view, g, v, i, j, xf, yf, ok);
    }
}

}  // namespace r360
