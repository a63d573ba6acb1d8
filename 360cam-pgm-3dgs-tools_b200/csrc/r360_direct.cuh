// Direct path: one thread per output pixel, float64 projection, taps gathered through L1/L2.
//
// This is the reference-grade device path: it handles every layout, border and projection the
// ABI accepts, and the tiled fast path hands it the tiles it cannot stage (poles, sensor edges).
// Arithmetic per pixel:
//   ray (gs360_GUI.py:377-392) -> lon/lat -> ERP pixel (gs360_GUI.py:419-424)            [ERP]
//   ray -> equisolid + Brown -> sensor pixel, validity (DF:1794-1821)                     [fisheye]
//   normalised sensor coordinates / zoom -> Brown -> sensor pixel, validity (DF:1008-1051) [undistort]
//   float32 cast -> 1/32-px quantisation -> cv2.remap weights (DF:2001-2008)              [sampling]
#pragma once

#include "r360_common.cuh"
#include "r360_sample.cuh"

namespace r360 {

// ---- projection of one output pixel to (x, y, valid) in float64 -----------------------------

template <int PROJ>
__device__ __forceinline__ bool project_pixel(const ViewDev& view, const ErpDev& erp, const LensDev* lens,
                                              double fi, double fj, double& x, double& y) {
    double dx, dy, dz;
    ray_at(view, fi, fj, dx, dy, dz);
    if (PROJ == kProjErp) {
        double lon, lat;
        erp_lonlat(dx, dy, dz, lon, lat);
        erp_xy(erp, lon, lat, x, y);
        return true;
    }
    if (PROJ == kProjUndistort) return undistort_xy(lens[view.slot], dx, dy, x, y);
    return fisheye_xy(lens[view.slot], dx, dy, dz, x, y);
}

// Sampling half of one output pixel, for NF consecutive frames of the batch starting at frame g: (x, y) is the
// source coordinate project_pixel returned, already cast to float32 as cv2's map would be.
template <int PROJ, int INTERP, typename TIn, typename TOut, int NF>
__device__ __forceinline__ void direct_sample(const LaunchParams& p, const ViewDev& view, int g, int v, int i, int j,
                                              float x, float y, bool valid) {
    const long long dst_img = (long long)g * p.n_views_total + p.view_base + v;
    const long long dst_fstride = (long long)p.n_views_total * p.dst.image_stride / (long long)sizeof(TOut);   // elements
    TOut* dst = reinterpret_cast<TOut*>(p.dst.data + dst_img * p.dst.image_stride + (long long)j * p.dst.pitch)
                + (long long)i * p.channels;
    if (PROJ != kProjErp && p.fill_invalid && !valid) {
        for (int f = 0; f < NF; ++f)
            for (int c = 0; c < p.channels; ++c) dst[f * dst_fstride + c] = Finish<TIn, TOut>::run(p.border_value);
        return;
    }
    const unsigned char* img = p.src.data + ((long long)g * p.n_lenses + view.slot) * p.src.image_stride;
    const long long src_fstride = (long long)p.n_lenses * p.src.image_stride;
    if (PROJ == kProjErp) {
        const ErpGlobalTaps<TIn> taps{img, p.src.pitch, p.src.width, p.src.height, p.channels};
        sample_pixel_frames<INTERP, TIn, TOut, NF>(taps, src_fstride, p.channels, p.src.width, p.src.height, 0.f, x, y, dst, dst_fstride);
    } else {
        const ConstBorderGlobalTaps<TIn> taps{img, p.src.pitch, p.src.width, p.src.height, p.channels, p.border_value};
        sample_pixel_frames<INTERP, TIn, TOut, NF>(taps, src_fstride, p.channels, p.src.width, p.src.height, p.border_value,
                                                   x, y, dst, dst_fstride);
    }
}

// One output pixel, start to finish.
template <int PROJ, int INTERP, typename TIn, typename TOut>
__device__ __forceinline__ void direct_pixel(const LaunchParams& p, const ViewDev& view, int g, int v, int i, int j) {
    double x, y;
    const bool valid = project_pixel<PROJ>(view, p.erp, p.lens, (double)i, (double)j, x, y);
    direct_sample<PROJ, INTERP, TIn, TOut, 1>(p, view, g, v, i, j, (float)x, (float)y, valid);
}

template <int PROJ, int INTERP, typename TIn, typename TOut>
__global__ void __launch_bounds__(256) remap_direct_kernel(const __grid_constant__ LaunchParams p) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);
    const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= p.dst.width || j >= p.dst.height) return;
    const int v = blockIdx.z % p.n_views;
    const int g = blockIdx.z / p.n_views;
    direct_pixel<PROJ, INTERP, TIn, TOut>(p, p.views[v], g, v, i, j);
}

__global__ void __launch_bounds__(256) coords_kernel(const __grid_constant__ CoordParams p, int proj) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);
    const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= p.out_w || j >= p.out_h) return;
    const int v = blockIdx.z;
    double x, y;
    const bool valid = proj == kProjErp
        ? project_pixel<kProjErp>(p.views[v], p.erp, p.lens, (double)i, (double)j, x, y)
        : proj == kProjFisheye ? project_pixel<kProjFisheye>(p.views[v], p.erp, p.lens, (double)i, (double)j, x, y)
                               : project_pixel<kProjUndistort>(p.views[v], p.erp, p.lens, (double)i, (double)j, x, y);
    const long long o = ((p.view_base + v) * p.out_h + j) * (long long)p.out_w + i;
    if (p.x32) p.x32[o] = (float)x;
    if (p.y32) p.y32[o] = (float)y;
    if (p.x64) p.x64[o] = x;
    if (p.y64) p.y64[o] = y;
    if (p.valid) p.valid[o] = valid ? 1 : 0;
}

}  // namespace r360
