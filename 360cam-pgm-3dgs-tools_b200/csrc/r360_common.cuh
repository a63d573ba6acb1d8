// Shared device/host definitions of the remap360 kernels (sm_100a).
//
// Projection math follows gs360_GUI.py:342-424 (ERP) and
// cli_tools/gs360_DualFisheyeDistortionCalibration.py:975-1005, :1759-1823 (equisolid + Brown);
// sampling arithmetic is cv2.remap's (the call at DF:2001-2008), see weights.cpp.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "r360_tables.h"

namespace r360 {

constexpr int kMaxViewsPerLaunch = 16;
constexpr int kMaxLenses = 4;

// kProjUndistort: fisheye image -> "undistorted" fisheye image (DF:1008-1051); its views carry the
// normalised sensor-plane coordinates (x / zoom, y / zoom, 1) instead of a world ray.
enum Proj : int { kProjErp = 0, kProjFisheye = 1, kProjUndistort = 2 };
// Radial law of a fisheye source: r = 2 f sin(theta / 2) (Metashape calibrations, v360 `equisolid`) or
// r = f theta (v360 `fisheye`, the --fisheye-projection equidistant of gs360_Video2Frames.py:466-473).
enum LensModel : int { kLensEquisolid = 0, kLensEquidistant = 1 };
enum Interp : int { kNearest = 0, kLinear = 1, kCubic = 2, kLanczos4 = 3 };

// A view is a linear map from output pixel indices to an (unnormalised) world ray:
//   d(i, j) = c0 + i * ci + j * cj
// with d = R * (tan(hfov/2) * ((2i+1)/w - 1), -tan(vfov/2) * ((2j+1)/h - 1), 1).
//
// kind == kRayFisheye (v360 `output=fisheye`, equidistant): the pixel's flat coordinates
//   (u, v) = (f[0] * i + f[1], f[2] * j + f[3])      [= h_fov/180 * ((2i+1)/w - 1), v likewise]
// give the angle from the view axis alpha = pi/2 * hypot(u, v) and the azimuth atan2(v, u); the
// camera ray (sin(alpha) cos(phi), -sin(alpha) sin(phi), cos(alpha)) is rotated by the matrix
// whose COLUMNS are c0, ci, cj.
enum RayKind : int { kRayLinear = 0, kRayFisheye = 1 };
struct ViewDev {
    double c0[3];
    double ci[3];
    double cj[3];
    double f[4];
    int32_t slot;
    int32_t kind;
};

struct ErpDev {          // x = (lon/2pi + 0.5) * su + ou ;  y = (0.5 - lat/pi) * sv + ov
    double su, ou, sv, ov;
};

struct LensDev {         // one fisheye calibration
    double cx0, cy0;     // width/2 + cx, height/2 + cy
    double f, b1, b2;
    double k1, k2, k3, k4, p1, p2;
    double xmax, ymax;   // width - 1, height - 1
    double cos_theta_max;
    double sin_half_theta_max;   // kProjUndistort: valid <=> min(r / 2, 1) <= sin(theta_max / 2)
    int32_t model;               // LensModel
    int32_t pad;
};

struct ImageSetDev {
    unsigned char* data;
    long long pitch;
    long long image_stride;
    int width, height;
};

struct LaunchParams {
    ImageSetDev src;
    ImageSetDev dst;
    int channels;
    int n_views;          // views in this launch (<= kMaxViewsPerLaunch)
    int view_base;        // index of views[0] within the call's view list
    int n_views_total;    // dst images per source group
    int n_lenses;         // source images per group
    int n_groups;
    int fill_invalid;
    float border_value;
    ErpDev erp;
    LensDev lens[kMaxLenses];
    ViewDev views[kMaxViewsPerLaunch];
};

struct CoordParams {
    int out_w, out_h, n_views;
    ErpDev erp;
    LensDev lens[kMaxLenses];
    ViewDev views[kMaxViewsPerLaunch];
    float* x32; float* y32; double* x64; double* y64; unsigned char* valid;
    long long view_base;
};

#ifdef __CUDACC__

// (fi, fj) may be fractional: the tile fitter samples between pixel centres.
__device__ __forceinline__ void ray_at(const ViewDev& v, double fi, double fj, double& dx, double& dy, double& dz) {
    if (v.kind == kRayFisheye) {
        const double u = fma(fi, v.f[0], v.f[1]), w = fma(fj, v.f[2], v.f[3]);
        const double r = sqrt(fma(u, u, w * w));
        double sa, ca;
        sincos(1.5707963267948966 * r, &sa, &ca);
        const double k = r > 1e-12 ? sa / r : 1.5707963267948966;
        const double cx = u * k, cy = -w * k;
        dx = fma(cx, v.c0[0], fma(cy, v.ci[0], ca * v.cj[0]));
        dy = fma(cx, v.c0[1], fma(cy, v.ci[1], ca * v.cj[1]));
        dz = fma(cx, v.c0[2], fma(cy, v.ci[2], ca * v.cj[2]));
        return;
    }
    dx = fma(fi, v.ci[0], fma(fj, v.cj[0], v.c0[0]));
    dy = fma(fi, v.ci[1], fma(fj, v.cj[1], v.c0[1]));
    dz = fma(fi, v.ci[2], fma(fj, v.cj[2], v.c0[2]));
}

// ERP: lon = atan2(x, z), lat = asin(y / |d|) written as atan2(y, hypot(x, z)).
// Returns longitude / latitude in radians.
__device__ __forceinline__ void erp_lonlat(double dx, double dy, double dz, double& lon, double& lat) {
    lon = atan2(dx, dz);
    lat = atan2(dy, sqrt(fma(dx, dx, dz * dz)));
}

__device__ __forceinline__ void erp_xy(const ErpDev& e, double lon, double lat, double& x, double& y) {
    const double inv_2pi = 0.15915494309189533577;
    const double inv_pi = 0.31830988618379067154;
    x = fma(fma(lon, inv_2pi, 0.5), e.su, e.ou);
    y = fma(fma(-lat, inv_pi, 0.5), e.sv, e.ov);
}

// Brown distortion in normalised image coordinates, then the affine sensor model
// (DF:975-1005, DF:1812-1817).  Returns r^2.
__device__ __forceinline__ double brown_to_sensor(const LensDev& L, double xn, double yn, double& x, double& y) {
    const double r2 = fma(xn, xn, yn * yn);
    const double r4 = r2 * r2;
    const double radial = 1.0 + L.k1 * r2 + L.k2 * r4 + L.k3 * (r4 * r2) + L.k4 * (r4 * r4);
    double xd = xn * radial, yd = yn * radial;
    if (L.p1 != 0.0 || L.p2 != 0.0) {
        const double xy = xn * yn;
        xd += L.p1 * (r2 + 2.0 * xn * xn) + 2.0 * L.p2 * xy;
        yd += L.p2 * (r2 + 2.0 * yn * yn) + 2.0 * L.p1 * xy;
    }
    x = L.cx0 + xd * L.f + xd * L.b1 + yd * L.b2;
    y = L.cy0 + yd * L.f;
    return r2;
}

// Equisolid fisheye with Brown distortion.  With n = |d|:
//   2 sin(theta/2) / rho = sqrt(2 / (n (n + dz)))   (theta from +z, rho = hypot(dx, dy) / n)
// so no trigonometry is needed.  valid = theta <= theta_max and inside the sensor.
__device__ __forceinline__ bool fisheye_xy(const LensDev& L, double dx, double dy, double dz,
                                           double& x, double& y) {
    const double n2 = fma(dx, dx, fma(dy, dy, dz * dz));
    const double n = sqrt(n2);
    double s;
    if (L.model == kLensEquidistant) {
        // theta / rho with rho = hypot(dx, dy) / n: the ray's angle from the axis spread linearly over the radius
        const double h = sqrt(fma(dx, dx, dy * dy));
        s = h > 1e-12 * n ? atan2(h, dz) / h : (dz > 0.0 ? 1.0 / n : 0.0);
    } else {
        const double den = n * (n + dz);
        s = den > 1e-24 * n2 ? sqrt(2.0 / den) : 0.0;
    }
    const double xn = dx * s;
    const double yn = -dy * s;
    brown_to_sensor(L, xn, yn, x, y);
    const bool in_fov = dz >= L.cos_theta_max * n;
    return in_fov && x >= 0.0 && x <= L.xmax && y >= 0.0 && y <= L.ymax;
}

// Fisheye -> undistorted fisheye (DF:1008-1051): (xn, yn) are the output pixel's normalised
// coordinates divided by the zoom; the source pixel is where the Brown model sends them.  The
// reference's validity test theta = 2 asin(clip(r / 2, 0, 1)) <= theta_max is monotone in r.
__device__ __forceinline__ bool undistort_xy(const LensDev& L, double xn, double yn, double& x, double& y) {
    const double r2 = brown_to_sensor(L, xn, yn, x, y);
    const bool in_fov = fmin(0.5 * sqrt(fmax(r2, 0.0)), 1.0) <= L.sin_half_theta_max;
    return in_fov && x >= 0.0 && x <= L.xmax && y >= 0.0 && y <= L.ymax;
}

// ---- element I/O -------------------------------------------------------------------------

template <typename T> struct Elem;
template <> struct Elem<uint8_t> {
    static __device__ __forceinline__ float to_float(uint8_t v) { return (float)v; }
};
template <> struct Elem<uint16_t> {
    static __device__ __forceinline__ float to_float(uint16_t v) { return (float)v; }
};
template <> struct Elem<__half> {
    static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
};
template <> struct Elem<float> {
    static __device__ __forceinline__ float to_float(float v) { return v; }
};

// float accumulator -> output element, cv2 saturate_cast semantics for integers
template <typename TIn, typename TOut> struct Finish;
template <> struct Finish<uint16_t, uint16_t> {
    static __device__ __forceinline__ uint16_t run(float a) {
        return (uint16_t)min(max(__float2int_rn(a), 0), 65535);
    }
};
template <> struct Finish<uint16_t, __half> {
    static __device__ __forceinline__ __half run(float a) {
        return __float2half_rn(__fmul_rn(a, 1.0f / 65535.0f));
    }
};
template <> struct Finish<__half, __half> {
    static __device__ __forceinline__ __half run(float a) { return __float2half_rn(a); }
};
template <> struct Finish<float, float> {
    static __device__ __forceinline__ float run(float a) { return a; }
};
template <> struct Finish<uint8_t, uint8_t> {
    static __device__ __forceinline__ uint8_t run(float a) {
        return (uint8_t)min(max(__float2int_rn(a), 0), 255);
    }
};

#endif  // __CUDACC__

}  // namespace r360
