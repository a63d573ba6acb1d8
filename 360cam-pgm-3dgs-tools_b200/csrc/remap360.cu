// remap360 C ABI (include/remap360.h): argument checking, view/lens preparation in float64,
// weight-table upload and kernel dispatch.  Device code lives in r360_direct.cuh (per-pixel
// float64 path) and r360_tiled.cuh (tile-staged fast path).
#include "../../include/remap360.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "r360_common.cuh"
#include "r360_direct.cuh"
#include "r360_tiled.cuh"
#include "r360_color.cuh"

using namespace r360;

namespace {

std::atomic<int64_t> g_launches{0};
thread_local char tl_cuda_error[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
    std::snprintf(tl_cuda_error, sizeof(tl_cuda_error), "%s: %s", what, cudaGetErrorString(e));
    return R360_E_CUDA;
}
#define R360_CUDA(call)                                         \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);   \
    } while (0)

// ---- one-time per-device state ---------------------------------------------------------

std::mutex g_init_mutex;
bool g_device_ready[64] = {};

// Chebyshev nodes, the inverse monomial Vandermonde matrix at those nodes (Gauss-Jordan in long
// double) and the check points of the tile fitter (r360_tiled.cuh).
void make_fit_constants(FitConstants* f) {
    const int n = kFitN;
    long double v[kFitN][2 * kFitN];
    for (int k = 0; k < n; ++k) {
        const long double t = cosl(M_PIl * (k + 0.5L) / n);
        f->node[k] = (double)t;
        long double pw = 1.0L;
        for (int m = 0; m < n; ++m) { v[k][m] = pw; pw *= t; }
        for (int m = 0; m < n; ++m) v[k][n + m] = (k == m) ? 1.0L : 0.0L;
    }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(v[r][c]) > fabsl(v[piv][c])) piv = r;
        for (int m = 0; m < 2 * n; ++m) { const long double tmp = v[c][m]; v[c][m] = v[piv][m]; v[piv][m] = tmp; }
        const long double d = v[c][c];
        for (int m = 0; m < 2 * n; ++m) v[c][m] /= d;
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const long double fct = v[r][c];
            for (int m = 0; m < 2 * n; ++m) v[r][m] -= fct * v[c][m];
        }
    }
    for (int m = 0; m < n; ++m)
        for (int k = 0; k < n; ++k) f->minv[m * n + k] = (double)v[m][n + k];
    const double cp[kFitChecks][2] = {{-1, -1}, {1, -1}, {-1, 1}, {1, 1}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {0, 0},
                                      {-.5, -.5}, {.5, -.5}, {-.5, .5}, {.5, .5}};
    for (int q = 0; q < kFitChecks; ++q) { f->check[q][0] = cp[q][0]; f->check[q][1] = cp[q][1]; }
}

int ensure_device_ready() {
    int dev = 0;
    R360_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return R360_E_NO_DEVICE;
    std::lock_guard<std::mutex> lock(g_init_mutex);
    if (g_device_ready[dev]) return R360_OK;
    cudaDeviceProp prop;
    R360_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        std::snprintf(tl_cuda_error, sizeof(tl_cuda_error),
                      "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
        return R360_E_NO_DEVICE;
    }
    static WeightTables host_tables;   // built once, identical for every device
    static bool built = false;
    if (!built) {
        build_weight_tables(&host_tables);
        built = true;
    }
    R360_CUDA(cudaMemcpyToSymbol(g_tables, &host_tables, sizeof(WeightTables)));
    // L2 persisting carve-out: the patch loads carry an evict_last policy, which only protects lines while the
    // device has a persisting region to keep them in (experiments: R360_L2_PERSIST_MB, default = leave the device alone)
    if (const char* env = std::getenv("R360_L2_PERSIST_MB")) {
        const size_t want = (size_t)std::max(0, std::atoi(env)) << 20;
        const size_t cap = (size_t)prop.persistingL2CacheMaxSize;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(want, cap));
    }
    static FitConstants fit;
    make_fit_constants(&fit);
    R360_CUDA(cudaMemcpyToSymbol(c_fit, &fit, sizeof(FitConstants)));
    g_device_ready[dev] = true;
    return R360_OK;
}

// ---- float64 preparation of views and lenses -----------------------------------------------

double clamp_fov_rad(double deg) {
    // gs360_GUI.py:437-438 / DF:1781-1782
    const double d = std::fmin(std::fmax(deg, 1e-3), 179.9);
    return d * (M_PI / 180.0);
}

void make_view(const r360_view& v, int out_w, int out_h, ViewDev* o) {
    const double yaw = v.yaw_deg * (M_PI / 180.0), pitch = v.pitch_deg * (M_PI / 180.0),
                 roll = v.roll_deg * (M_PI / 180.0);
    const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch),
                 cr = std::cos(roll), sr = std::sin(roll);
    // R = Ryaw * Rpitch * Rroll with the signs of gs360_GUI.py:351-374
    const double ry[3][3] = {{cy, 0, sy}, {0, 1, 0}, {-sy, 0, cy}};
    const double rp[3][3] = {{1, 0, 0}, {0, cp, sp}, {0, -sp, cp}};
    const double rr[3][3] = {{cr, -sr, 0}, {sr, cr, 0}, {0, 0, 1}};
    double t[3][3], R[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            t[a][b] = 0;
            for (int k = 0; k < 3; ++k) t[a][b] += rp[a][k] * rr[k][b];
        }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            R[a][b] = 0;
            for (int k = 0; k < 3; ++k) R[a][b] += ry[a][k] * t[k][b];
        }
    std::memset(o, 0, sizeof(*o));
    o->slot = v.src_slot;
    if (v.projection == R360_OUT_FISHEYE) {
        // v360 output=fisheye: flat coordinates scaled by fov / 180 (prepare_fisheye_out), R's columns in c0/ci/cj
        o->kind = kRayFisheye;
        for (int a = 0; a < 3; ++a) { o->c0[a] = R[a][0]; o->ci[a] = R[a][1]; o->cj[a] = R[a][2]; }
        const double su = std::fmin(std::fmax(v.hfov_deg, 1e-3), 360.0) / 180.0;
        const double sv = std::fmin(std::fmax(v.vfov_deg, 1e-3), 360.0) / 180.0;
        o->f[0] = su * 2.0 / out_w; o->f[1] = su * (1.0 / out_w - 1.0);
        o->f[2] = sv * 2.0 / out_h; o->f[3] = sv * (1.0 / out_h - 1.0);
        return;
    }
    const double tx = std::tan(clamp_fov_rad(v.hfov_deg) * 0.5);
    const double ty = std::tan(clamp_fov_rad(v.vfov_deg) * 0.5);
    // camera ray (tx*u, -ty*vv, 1), u = (2i+1)/w - 1, vv = (2j+1)/h - 1
    const double u0 = 1.0 / out_w - 1.0, v0 = 1.0 / out_h - 1.0;
    const double du = 2.0 / out_w, dv = 2.0 / out_h;
    for (int a = 0; a < 3; ++a) {
        o->c0[a] = R[a][0] * (tx * u0) - R[a][1] * (ty * v0) + R[a][2];
        o->ci[a] = R[a][0] * (tx * du);
        o->cj[a] = -R[a][1] * (ty * dv);
    }
}

void make_erp(int W, int H, int convention, ErpDev* e) {
    if (convention == R360_CONV_V360) {
        e->su = W - 1.0; e->ou = 0.0; e->sv = H - 1.0; e->ov = 0.0;
    } else {
        e->su = W; e->ou = -0.5; e->sv = H; e->ov = -0.5;
    }
}

void make_lens(const r360_fisheye_calib& c, LensDev* L) {
    L->cx0 = c.width * 0.5 + c.cx;
    L->cy0 = c.height * 0.5 + c.cy;
    L->f = c.f; L->b1 = c.b1; L->b2 = c.b2;
    L->k1 = c.k1; L->k2 = c.k2; L->k3 = c.k3; L->k4 = c.k4; L->p1 = c.p1; L->p2 = c.p2;
    L->xmax = c.width - 1.0;
    L->ymax = c.height - 1.0;
    const double half = std::fmax(1.0, std::fmin(360.0, c.lens_fov_deg)) * 0.5;   // DF:1800
    L->cos_theta_max = std::cos(half * (M_PI / 180.0));
    L->sin_half_theta_max = std::sin(half * 0.5 * (M_PI / 180.0));
    L->model = c.model == R360_LENS_EQUIDISTANT ? kLensEquidistant : kLensEquisolid;
    L->pad = 0;
}

// Fisheye -> undistorted fisheye (DF:1021-1029): the output pixel (i, j) has normalised coordinates
//   y = (j - cy0) / f / zoom,   x = (i - cx0 - (j - cy0) / f * b2) / (f + b1) / zoom,
// a linear function of (i, j): it rides in the view's "ray" slots (dz is unused).
int make_undistort_view(const r360_undistort& u, const r360_fisheye_calib& c, ViewDev* o) {
    const double cx0 = c.width * 0.5 + c.cx, cy0 = c.height * 0.5 + c.cy;
    const double den_y = c.f, den_x = c.f + c.b1;
    if (std::fabs(den_y) < 1e-12 || std::fabs(den_x) < 1e-12) return R360_E_INVALID_ARG;     // DF:1018-1020
    const double zoom = std::fmax(1e-6, u.zoom);                                             // DF:1145
    std::memset(o, 0, sizeof(*o));
    o->ci[0] = 1.0 / (den_x * zoom);
    o->cj[0] = -c.b2 / (den_y * den_x * zoom);
    o->c0[0] = (-cx0 + cy0 * c.b2 / den_y) / (den_x * zoom);
    o->cj[1] = 1.0 / (den_y * zoom);
    o->c0[1] = -cy0 / (den_y * zoom);
    o->c0[2] = 1.0;
    o->slot = u.src_slot;
    return R360_OK;
}

int build_persp_views(const r360_view* views, int n_views, int out_w, int out_h, std::vector<ViewDev>* out) {
    if (!views || n_views <= 0 || out_w <= 0 || out_h <= 0) return R360_E_INVALID_ARG;
    out->resize(n_views);
    for (int v = 0; v < n_views; ++v) {
        if (views[v].projection != R360_OUT_RECTILINEAR && views[v].projection != R360_OUT_FISHEYE)
            return R360_E_INVALID_ARG;
        make_view(views[v], out_w, out_h, &(*out)[v]);
    }
    return R360_OK;
}

int build_undistort_views(const r360_undistort* items, int n_items, const r360_fisheye_calib* calib, int n_lenses,
                          std::vector<ViewDev>* out) {
    if (!items || n_items <= 0 || !calib) return R360_E_INVALID_ARG;
    if (n_lenses < 1) return R360_E_INVALID_ARG;
    if (n_lenses > R360_MAX_LENSES) return R360_E_TOO_MANY;
    out->resize(n_items);
    for (int v = 0; v < n_items; ++v) {
        if (items[v].src_slot < 0 || items[v].src_slot >= n_lenses) return R360_E_INVALID_ARG;
        const int rc = make_undistort_view(items[v], calib[items[v].src_slot], &(*out)[v]);
        if (rc != R360_OK) return rc;
    }
    return R360_OK;
}

int elem_size(int dtype) {
    switch (dtype) {
        case R360_U8: return 1;
        case R360_U16: case R360_F16: return 2;
        case R360_F32: return 4;
        default: return 0;
    }
}

int check_images(const r360_images* im) {
    if (!im || !im->data) return R360_E_INVALID_ARG;
    if (im->width <= 0 || im->height <= 0 || im->count <= 0) return R360_E_INVALID_ARG;
    if (im->channels < 1 || im->channels > 4) return R360_E_UNSUPPORTED;
    const int es = elem_size(im->dtype);
    if (!es) return R360_E_UNSUPPORTED;
    if (im->pitch_bytes < (int64_t)im->width * im->channels * es) return R360_E_INVALID_ARG;
    if (im->pitch_bytes % es) return R360_E_INVALID_ARG;
    if (im->count > 1 && im->image_stride_bytes < im->pitch_bytes * im->height) return R360_E_INVALID_ARG;
    if (im->image_stride_bytes % es) return R360_E_INVALID_ARG;
    return R360_OK;
}

// ---- dispatch ------------------------------------------------------------------------------------

struct Prepared {          // everything the kernels need, derived from the ABI arguments
    int proj, n_lenses, n_views, interp, in_dt, out_dt;
    LaunchParams lp;       // views[] filled per launch chunk
    r360_options opt;
};

template <int PROJ, int INTERP, typename TIn, typename TOut>
int launch_direct(const LaunchParams& p, cudaStream_t stream) {
    dim3 block(256);
    dim3 grid((p.dst.width + 31) / 32, (p.dst.height + 7) / 8, 1);
    // grid.z is limited to 65535: walk the groups in chunks
    const int max_groups = 65535 / p.n_views;
    for (int g0 = 0; g0 < p.n_groups; g0 += max_groups) {
        LaunchParams q = p;
        const int ng = p.n_groups - g0 < max_groups ? p.n_groups - g0 : max_groups;
        q.src.data += (long long)g0 * p.n_lenses * p.src.image_stride;
        q.dst.data += (long long)g0 * p.n_views_total * p.dst.image_stride;
        q.n_groups = ng;
        grid.z = ng * p.n_views;
        remap_direct_kernel<PROJ, INTERP, TIn, TOut><<<grid, block, 0, stream>>>(q);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        R360_CUDA(cudaGetLastError());
    }
    return R360_OK;
}

// type / interpolation / projection switchboard: calls f.template run<PROJ, INTERP, TIn, TOut>()
template <typename F>
int dispatch(int proj, int interp, int in_dt, int out_dt, F&& f) {
#define R360_CASE_T(TIN, TOUT)                                                                              \
    do {                                                                                                    \
        if (proj == kProjErp) {                                                                             \
            if (interp == R360_NEAREST) return f.template run<kProjErp, kNearest, TIN, TOUT>();             \
            if (interp == R360_LINEAR) return f.template run<kProjErp, kLinear, TIN, TOUT>();               \
            if (interp == R360_CUBIC) return f.template run<kProjErp, kCubic, TIN, TOUT>();                 \
            if (interp == R360_LANCZOS4) return f.template run<kProjErp, kLanczos4, TIN, TOUT>();           \
        } else if (proj == kProjFisheye) {                                                                  \
            if (interp == R360_NEAREST) return f.template run<kProjFisheye, kNearest, TIN, TOUT>();         \
            if (interp == R360_LINEAR) return f.template run<kProjFisheye, kLinear, TIN, TOUT>();           \
            if (interp == R360_CUBIC) return f.template run<kProjFisheye, kCubic, TIN, TOUT>();             \
            if (interp == R360_LANCZOS4) return f.template run<kProjFisheye, kLanczos4, TIN, TOUT>();       \
        } else if (proj == kProjUndistort) {                                                                \
            if (interp == R360_NEAREST) return f.template run<kProjUndistort, kNearest, TIN, TOUT>();       \
            if (interp == R360_LINEAR) return f.template run<kProjUndistort, kLinear, TIN, TOUT>();         \
            if (interp == R360_CUBIC) return f.template run<kProjUndistort, kCubic, TIN, TOUT>();           \
            if (interp == R360_LANCZOS4) return f.template run<kProjUndistort, kLanczos4, TIN, TOUT>();     \
        }                                                                                                   \
        return R360_E_INVALID_ARG;                                                                          \
    } while (0)
    if (in_dt == R360_U8 && out_dt == R360_U8) R360_CASE_T(uint8_t, uint8_t);
    if (in_dt == R360_U16 && out_dt == R360_U16) R360_CASE_T(uint16_t, uint16_t);
    if (in_dt == R360_U16 && out_dt == R360_F16) R360_CASE_T(uint16_t, __half);
    if (in_dt == R360_F16 && out_dt == R360_F16) R360_CASE_T(__half, __half);
    if (in_dt == R360_F32 && out_dt == R360_F32) R360_CASE_T(float, float);
#undef R360_CASE_T
    return R360_E_UNSUPPORTED;
}

bool supported_types(int in_dt, int out_dt) {
    return (in_dt == R360_U8 && out_dt == R360_U8) || (in_dt == R360_U16 && out_dt == R360_U16) ||
           (in_dt == R360_U16 && out_dt == R360_F16) || (in_dt == R360_F16 && out_dt == R360_F16) ||
           (in_dt == R360_F32 && out_dt == R360_F32);
}

// Validates the arguments shared by the plan-less and planned entry points and fills `out`.
// With `layout_only` the data pointers and counts are not looked at.
int prepare(int proj, const r360_images* src, const r360_images* dst, const r360_fisheye_calib* calib,
            int n_lenses, const ViewDev* views, int n_views, const r360_options* opt_in, bool layout_only,
            Prepared* out) {
    int rc;
    r360_images s_chk, d_chk;
    if (!src || !dst) return R360_E_INVALID_ARG;
    s_chk = *src; d_chk = *dst;
    if (layout_only) {
        s_chk.data = d_chk.data = reinterpret_cast<void*>(16);
        s_chk.count = n_lenses > 0 ? n_lenses : 1;
        d_chk.count = n_views > 0 ? n_views : 1;
    }
    if ((rc = check_images(&s_chk)) != R360_OK) return rc;
    if ((rc = check_images(&d_chk)) != R360_OK) return rc;
    if (!views || n_views <= 0) return R360_E_INVALID_ARG;
    if (n_lenses < 1) return R360_E_INVALID_ARG;
    if (n_lenses > R360_MAX_LENSES) return R360_E_TOO_MANY;
    if (proj != kProjErp && !calib) return R360_E_INVALID_ARG;
    r360_options opt;
    if (opt_in) opt = *opt_in; else r360_default_options(&opt);
    if (opt.interp < R360_NEAREST || opt.interp > R360_LANCZOS4) return R360_E_INVALID_ARG;
    if (opt.convention != R360_CONV_HALFPIXEL && opt.convention != R360_CONV_V360) return R360_E_INVALID_ARG;
    if (opt.path < R360_PATH_AUTO || opt.path > R360_PATH_TILED) return R360_E_INVALID_ARG;
    if (s_chk.channels != d_chk.channels) return R360_E_INVALID_ARG;
    if (s_chk.count % n_lenses) return R360_E_INVALID_ARG;
    const int n_groups = s_chk.count / n_lenses;
    if ((int64_t)d_chk.count != (int64_t)n_groups * n_views) return R360_E_INVALID_ARG;
    const int out_dt = opt.out_dtype < 0 ? src->dtype : opt.out_dtype;
    if (out_dt != dst->dtype) return R360_E_INVALID_ARG;
    if (!supported_types(src->dtype, out_dt)) return R360_E_UNSUPPORTED;
    for (int v = 0; v < n_views; ++v)
        if (views[v].slot < 0 || views[v].slot >= n_lenses) return R360_E_INVALID_ARG;

    out->proj = proj; out->n_lenses = n_lenses; out->n_views = n_views; out->interp = opt.interp;
    out->in_dt = src->dtype; out->out_dt = out_dt; out->opt = opt;
    LaunchParams& p = out->lp;
    std::memset(&p, 0, sizeof(p));
    p.src = {static_cast<unsigned char*>(src->data), src->pitch_bytes, src->image_stride_bytes, src->width, src->height};
    p.dst = {static_cast<unsigned char*>(dst->data), dst->pitch_bytes, dst->image_stride_bytes, dst->width, dst->height};
    p.channels = src->channels;
    p.n_views_total = n_views;
    p.n_lenses = n_lenses;
    p.n_groups = n_groups;
    p.fill_invalid = proj != kProjErp ? (opt.fill_invalid != 0) : 0;
    double bv = proj != kProjErp ? opt.border_value : 0.0;
    if (src->dtype == R360_U8) bv = std::nearbyint(std::fmin(std::fmax(bv, 0.0), 255.0));          // saturate_cast<uchar>
    else if (src->dtype == R360_U16) bv = std::nearbyint(std::fmin(std::fmax(bv, 0.0), 65535.0));  // saturate_cast<ushort>
    p.border_value = (float)bv;
    if (proj == kProjErp) make_erp(src->width, src->height, opt.convention, &p.erp);
    else for (int l = 0; l < n_lenses; ++l) make_lens(calib[l], &p.lens[l]);
    return R360_OK;
}

struct DirectLauncher {
    const LaunchParams& p; cudaStream_t s;
    template <int PROJ, int INTERP, typename TIn, typename TOut> int run() { return launch_direct<PROJ, INTERP, TIn, TOut>(p, s); }
};

int remap_direct(int proj, const r360_images* src, const r360_images* dst, const r360_fisheye_calib* calib,
                 int n_lenses, const std::vector<ViewDev>& views, const r360_options* opt_in, void* stream) {
    Prepared pr;
    const int n_views = (int)views.size();
    int rc = prepare(proj, src, dst, calib, n_lenses, views.data(), n_views, opt_in, false, &pr);
    if (rc != R360_OK) return rc;
    if (pr.opt.path == R360_PATH_TILED) return R360_E_INVALID_ARG;      // needs a plan
    if ((rc = ensure_device_ready()) != R360_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    LaunchParams& p = pr.lp;
    for (int v0 = 0; v0 < n_views; v0 += kMaxViewsPerLaunch) {
        p.view_base = v0;
        p.n_views = n_views - v0 < kMaxViewsPerLaunch ? n_views - v0 : kMaxViewsPerLaunch;
        for (int v = 0; v < p.n_views; ++v) p.views[v] = views[v0 + v];
        rc = dispatch(proj, pr.interp, pr.in_dt, pr.out_dt, DirectLauncher{p, s});
        if (rc != R360_OK) return rc;
    }
    return R360_OK;
}

// ---- plans ------------------------------------------------------------------------------------------


int env_int(const char* name, int fallback) {
    const char* e = std::getenv(name);
    return e && *e ? std::atoi(e) : fallback;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kMaxRingBytes = 200 * 1024;       // shared-memory ring of one block

struct WorkspaceLayout { size_t header, views, plans, fallback, order, coords, total; int coords_capacity; };

// Per-pixel maps for the tiles no polynomial follows (pole neighbourhoods): room for one tile in sixteen, which covers
// a view set made of pole views only (4.6 % of a pitch-90 view's tiles at 8K); tiles beyond the pool stay on the
// fallback list.
int coords_pool_tiles(size_t tiles_total) { return (int)std::min<size_t>(tiles_total / 16 + 32, 1u << 20); }

WorkspaceLayout workspace_layout(int n_views, int out_w, int out_h) {
    const size_t tiles = (size_t)((out_w + kTile - 1) / kTile) * ((out_h + kTile - 1) / kTile);
    WorkspaceLayout w;
    w.header = 0;
    w.views = align_up(sizeof(PlanHeader), 256);
    w.plans = w.views + align_up(sizeof(ViewDev) * n_views, 256);
    w.fallback = w.plans + align_up(sizeof(TilePlan) * n_views * tiles, 256);
    w.order = w.fallback + align_up(sizeof(int2) * n_views * tiles, 256);      // also scratch for the sort keys
    w.coords = w.order + align_up(sizeof(int2) * n_views * tiles, 256);
    w.coords_capacity = coords_pool_tiles((size_t)n_views * tiles);
    w.total = w.coords + (size_t)w.coords_capacity * kCoordTileBytes;
    return w;
}

}  // namespace

struct r360_plan {
    Prepared pr;
    r360_images src_layout, dst_layout;
    std::vector<ViewDev> views;
    int tiles_x, tiles_y, n_tiles, n_fallback;
    int out_stage_bytes, patch_budget, ring_bytes, smem_bytes, ctas_per_sm, use_table, sm_count;
    int frames_pref, teams_multi_pref, ctas_multi_pref;     // launch shape for batches (choose_shape)
    int multi_pct_pref;                                     // share of the ring one multi-frame item may take
    int mean_patch_bytes;                                   // mean ring bytes of a staged tile's patch (one frame)
    bool bulk_load_ok, bulk_store_ok;
    unsigned char* ws;
    PlanHeader* d_header; ViewDev* d_views; TilePlan* d_plans; int2* d_fallback; int2* d_order; double2* d_coords;
    int coords_capacity, n_coords;                  // per-pixel map pool: size, maps in use
    int n_order;                                    // staged / fill tiles, in the order the remap kernel walks them
    int n_large, patch_budget_large;                // tiles of the large-patch pass (they follow in d_order), its budget
    // tensor-TMA descriptors depend on the source base pointer and batch size: small cache
    bool tensor_ok;
    int box_family;
    mutable std::mutex tm_mutex;
    struct TmEntry { const void* data; int count; int64_t stride; TensorMaps maps; };
    mutable std::vector<TmEntry> tm_cache;
    // The fallback tiles run beside the tiled kernel on this stream (created with the plan, non-blocking).
    cudaStream_t side_stream = nullptr;
    ~r360_plan() { if (side_stream) cudaStreamDestroy(side_stream); }
};

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Descriptors for every box shape over (row bytes / 4 as uint32, rows, images).
int encode_tensor_maps(const r360_images& src, int family, TensorMaps* out) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return R360_E_CUDA;
    const int es = elem_size(src.dtype);
    const cuuint64_t dims[3] = {(cuuint64_t)src.width * src.channels * es / 4, (cuuint64_t)src.height,
                                (cuuint64_t)src.count};
    const cuuint64_t strides[2] = {(cuuint64_t)src.pitch_bytes,
                                   (cuuint64_t)(src.count > 1 ? src.image_stride_bytes : src.pitch_bytes * src.height)};
    const cuuint32_t estr[3] = {1, 1, 1};
    // L2 fetch granularity of the boxes (R360_L2_PROMO = 0 / 64 / 128 / 256 for experiments)
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char* env = std::getenv("R360_L2_PROMO")) {
        const int v = std::atoi(env);
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }
    for (int wk = 0; wk < kNumBoxWidths; ++wk)
        for (int hk = 0; hk < kNumBoxHeights; ++hk) {
            const cuuint32_t box[3] = {(cuuint32_t)box_width_bytes(wk, family) / 4, (cuuint32_t)box_height_rows(hk), 1};
            const CUresult r = enc(&out->m[wk * kNumBoxHeights + hk], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, src.data, dims,
                                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                std::snprintf(tl_cuda_error, sizeof(tl_cuda_error), "cuTensorMapEncodeTiled failed (%d) for box %dx%d",
                              (int)r, box_width_bytes(wk, family), box_height_rows(hk));
                return R360_E_CUDA;
            }
        }
    return R360_OK;
}

bool aligned16(const void* p, int64_t pitch, int64_t stride) {
    return reinterpret_cast<uintptr_t>(p) % 16 == 0 && pitch % 16 == 0 && stride % 16 == 0;
}

struct TiledShape { int fr, teams, ctas, ring, smem, multi_budget; };
bool shape_for(const r360_plan* pl, int fr, int teams, int ctas, int smem_per_sm, TiledShape* out);

int plan_create(int proj, const r360_images* src, const r360_images* dst, const r360_fisheye_calib* calib,
                int n_lenses, const std::vector<ViewDev>& views, const r360_options* opt_in,
                void* workspace, size_t workspace_bytes, void* stream, r360_plan** plan_out) {
    if (!plan_out || !workspace) return R360_E_INVALID_ARG;
    *plan_out = nullptr;
    const int n_views = (int)views.size();
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return R360_E_INVALID_ARG;
    r360_plan* pl = new (std::nothrow) r360_plan();
    if (!pl) return R360_E_INVALID_ARG;
    int rc = prepare(proj, src, dst, calib, n_lenses, views.data(), n_views, opt_in, true, &pl->pr);
    if (rc == R360_OK) rc = ensure_device_ready();
    const WorkspaceLayout wl = workspace_layout(n_views > 0 ? n_views : 1, dst ? dst->width : 1, dst ? dst->height : 1);
    if (rc == R360_OK && workspace_bytes < wl.total) rc = R360_E_INVALID_ARG;
    if (rc != R360_OK) { delete pl; return rc; }

    pl->src_layout = *src; pl->dst_layout = *dst;
    pl->tiles_x = (dst->width + kTile - 1) / kTile;
    pl->tiles_y = (dst->height + kTile - 1) / kTile;
    pl->n_tiles = pl->tiles_x * pl->tiles_y;
    pl->views = views;
    const int out_es = elem_size(pl->pr.out_dt), in_es = elem_size(pl->pr.in_dt);
    pl->out_stage_bytes = (int)align_up((size_t)kTile * kTile * dst->channels * out_es, 128);
    // Shared memory per block: barriers/coefficients/plan records, (8-bit bicubic) the 32 KB weight
    // table, two output tiles, two patch buffers.  The patch budget follows from how many blocks
    // should be resident per SM; tiles whose patch is larger take the fallback path.
    pl->use_table = (pl->pr.in_dt == R360_U8 && pl->pr.interp == R360_CUBIC && dst->channels == 3) ? 1 : 0;
    // lanczos4 on 8-bit RGB: the 128 KB table in shared memory, one block per SM (measured: the kernel runs as fast
    // with 8 warps per SM as with 32 -- it is bound by the table reads, 32 L1 lines per warp load when they go through L1)
    if (pl->pr.in_dt == R360_U8 && pl->pr.out_dt == R360_U8 && pl->pr.interp == R360_LANCZOS4 && dst->channels == 3 &&
        std::getenv("R360_LANCZOS_L1") == nullptr)
        pl->use_table = 2;
    {
        int dev = 0, smem_per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&pl->sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        // blocks per SM: what the kernel's registers are budgeted for (tiled_min_blocks); 8-bit bicubic without the
        // packed sampler (1 / 4 channels) has no table in shared memory but is compiled for 2 blocks all the same
        int want = pl->use_table == 2 ? 1 : tiled_min_blocks(in_es, pl->pr.interp, 1);
        if (const char* env = std::getenv("R360_TILED_CTAS_PER_SM")) want = std::atoi(env) > 0 ? std::atoi(env) : want;
        const int fixed = kTiledFixedSmem + table_bytes(pl->use_table) + pl->out_stage_bytes + 128;   // one team, one frame per item
        for (;; --want) {
            const int per_block = smem_per_sm / want - 1024;       // 1 KB per block is reserved by the driver
            pl->ring_bytes = (per_block - fixed) & ~127;
            if (pl->ring_bytes >= 24 * 1024 || want == 1) break;
        }
        if (pl->ring_bytes < 8192) pl->ring_bytes = 8192;
        if (pl->ring_bytes > kMaxRingBytes) pl->ring_bytes = kMaxRingBytes;
        if (const char* env = std::getenv("R360_RING_KB")) {                   // experiments: leave more of the SM to L1
            const int cap = std::atoi(env) * 1024;
            if (cap >= 8192 && cap < pl->ring_bytes) pl->ring_bytes = cap & ~127;
        }
        pl->ctas_per_sm = want;
        pl->smem_bytes = fixed + pl->ring_bytes;
        // batches: frames per work item / teams / blocks per SM (choose_shape).  8-bit bicubic: four frames per
        // item in ONE block per SM -- one weight table per SM instead of two, and a ring (~187 KB) that holds the
        // four-frame item being sampled AND the next one, so the L2 -> shared-memory latency of an item is hidden
        // behind the previous one.  Measured on B200 (16 x 8K frames -> 12 views, profiles/r02_shape_sweep.jsonl):
        // one team of eight consumer warps 159.4 Gpix/s, two teams (whose second stage set leaves the ring room for
        // the two items in flight only) 156.7, two blocks per SM with two frames per item 151.5.  A patch may use
        // the whole ring of the smallest shape a call can take: one team with kMaxFramesPerItem output stages.
        const bool cubic_u8 = pl->use_table == 1;
        const bool linear_u8 = pl->pr.in_dt == R360_U8 && pl->pr.interp == R360_LINEAR && dst->channels == 3;
        // (8-bit bilinear: four frames per item on one team, 264 against 256 Gpix/s with two; 16-bit patches are twice
        // the size and stay at two frames per item)
        // (16-bit bicubic: two frames per item on two teams in one block per SM, 82.1 against 76.8 Gpix/s on two blocks)
        const bool cubic_u16 = in_es == 2 && pl->pr.interp == R360_CUBIC && dst->channels == 3;
        pl->frames_pref = pl->use_table == 2 ? 1 : (cubic_u8 || linear_u8) ? 4 : 2;
        pl->teams_multi_pref = cubic_u16 ? 2 : 1;
        pl->multi_pct_pref = linear_u8 ? 100 : 75;
        pl->ctas_multi_pref = (cubic_u8 || cubic_u16) ? 1 : want;
        pl->patch_budget = pl->ring_bytes - (kMaxFramesPerItem - 1) * pl->out_stage_bytes;
        if (pl->patch_budget < 8192) pl->patch_budget = 8192;
        // Tiles whose patch is larger than that (panorama tiles a few degrees from a pole span hundreds of columns)
        // are walked in a second, short launch of the one-frame kernel with one block per SM, whose ring is the
        // largest the SM can give; only what exceeds even that goes to the fallback kernel.
        pl->patch_budget_large = std::min((smem_per_sm - 1024 - fixed) & ~127, kMaxRingBytes);
        if (pl->patch_budget_large < pl->patch_budget || env_int("R360_LARGE_PATCH_PASS", 1) == 0) pl->patch_budget_large = pl->patch_budget;
    }
    // the data pointers are not known yet: assume 16-byte aligned bases (checked at remap time)
    pl->bulk_load_ok = src->pitch_bytes % 16 == 0 && src->image_stride_bytes % 16 == 0 &&
                       ((int64_t)src->width * src->channels * in_es) % 16 == 0;
    pl->bulk_store_ok = dst->pitch_bytes % 16 == 0 && dst->image_stride_bytes % 16 == 0;
    pl->tensor_ok = pl->bulk_load_ok && encode_tiled_fn() != nullptr && std::getenv("R360_NO_TENSOR_TMA") == nullptr;
    // lane-per-column bicubic paths (8-bit and 16-bit): rows on identical banks; measured 16-bit bicubic 76.7 against
    // 73.3 Gpix/s with the odd multiples of 32 bytes, 8-bit bilinear (pixel pairs) 265 against 251 the other way round
    pl->box_family = (pl->use_table == 1 || (in_es == 2 && pl->pr.interp == R360_CUBIC && dst->channels == 3)) ? 1 : 0;
    if (const char* env = std::getenv("R360_BOX_FAMILY")) pl->box_family = std::atoi(env) == 1 ? 1 : 0;
    pl->ws = static_cast<unsigned char*>(workspace);
    pl->d_header = reinterpret_cast<PlanHeader*>(pl->ws + wl.header);
    pl->d_views = reinterpret_cast<ViewDev*>(pl->ws + wl.views);
    pl->d_plans = reinterpret_cast<TilePlan*>(pl->ws + wl.plans);
    pl->d_fallback = reinterpret_cast<int2*>(pl->ws + wl.fallback);
    pl->d_order = reinterpret_cast<int2*>(pl->ws + wl.order);
    pl->d_coords = reinterpret_cast<double2*>(pl->ws + wl.coords);
    pl->coords_capacity = env_int("R360_COORD_TILES", 1) != 0 ? wl.coords_capacity : 0;   // 0: experiments without the pool

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto fail = [&](cudaError_t e, const char* what) { delete pl; return cuda_fail(e, what); };
    cudaError_t e;
    if ((e = cudaMemsetAsync(pl->d_header, 0, sizeof(PlanHeader), s)) != cudaSuccess) return fail(e, "cudaMemsetAsync");
    if ((e = cudaMemcpyAsync(pl->d_views, pl->views.data(), sizeof(ViewDev) * n_views, cudaMemcpyHostToDevice, s)) != cudaSuccess)
        return fail(e, "cudaMemcpyAsync(views)");
    PlanParams P;
    std::memset(&P, 0, sizeof(P));
    P.proj = proj; P.out_w = dst->width; P.out_h = dst->height; P.tiles_x = pl->tiles_x; P.tiles_y = pl->tiles_y;
    P.n_views = n_views; P.src_w = src->width; P.src_h = src->height; P.px_bytes = src->channels * in_es;
    P.patch_budget = pl->patch_budget; P.patch_budget_large = pl->patch_budget_large; P.bulk_load_ok = pl->bulk_load_ok; P.tensor_ok = pl->tensor_ok;
    P.fill_invalid = pl->pr.lp.fill_invalid;
    P.interp = pl->pr.interp;
    P.box_family = pl->box_family;
    P.erp = pl->pr.lp.erp;
    std::memcpy(P.lens, pl->pr.lp.lens, sizeof(P.lens));
    P.views = pl->d_views; P.plans = pl->d_plans; P.header = pl->d_header; P.fallback = pl->d_fallback;
    P.coords = pl->coords_capacity > 0 ? pl->d_coords : nullptr; P.coords_capacity = pl->coords_capacity;
    for (int v0 = 0; v0 < n_views; v0 += 65535) {
        PlanParams Q = P;
        const int nv = n_views - v0 < 65535 ? n_views - v0 : 65535;
        Q.views = pl->d_views + v0; Q.plans = pl->d_plans + (size_t)v0 * pl->n_tiles; Q.n_views = nv;
        // the fallback list stores view indices relative to Q.views: only one chunk is ever needed in practice
        if (v0 != 0) { delete pl; return R360_E_TOO_MANY; }
        plan_kernel<<<dim3(pl->n_tiles, nv), 64, 0, s>>>(Q);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "plan_kernel launch");
    }
    PlanHeader h;
    if ((e = cudaMemcpyAsync(&h, pl->d_header, sizeof(h), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail(e, "cudaMemcpyAsync(header)");
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "cudaStreamSynchronize");
    pl->n_fallback = h.n_fallback;
    pl->n_coords = std::min(h.n_coords, pl->coords_capacity);
    if (pl->n_fallback > 0 && cudaStreamCreateWithFlags(&pl->side_stream, cudaStreamNonBlocking) != cudaSuccess) {
        pl->side_stream = nullptr;                                     // the fallback kernel then follows in `s`
        (void)cudaGetLastError();
    }
    // ---- walk order of the remap kernel: (view, tile) sorted by source slot and source row ------------------
    {
        const int n = n_views * pl->n_tiles;
        int* d_keys = reinterpret_cast<int*>(pl->d_order);            // the order region doubles as key scratch
        order_key_kernel<<<(n + 255) / 256, 256, 0, s>>>(pl->d_plans, n, d_keys, d_keys + n);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "order_key_kernel launch");
        std::vector<int> keys(2 * (size_t)n);                          // keys, then the patch bytes of every tile
        if ((e = cudaMemcpyAsync(keys.data(), d_keys, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail(e, "cudaMemcpyAsync(keys)");
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "cudaStreamSynchronize");
        {
            // 8-bit bicubic batches: two consumer teams when the ring they leave (one block per SM, two sets of output
            // stages) still holds the two four-frame items being sampled AND a third being loaded; else one team
            // with the whole ring.  Measured on B200: 8K -> 1600 px preset views (patches 15 KB a frame) 156.7 with
            // two teams, 159.4 with one; 3840x1920 -> 1600 px (5 KB) 196.7 against 182.7; dual fisheye (12 KB) 166.1
            // against 157.7.
            long long sum = 0, cnt = 0;
            for (int i = 0; i < n; ++i)
                if (keys[n + i] > 0 && keys[i] < INT_MAX - 1) { sum += keys[n + i]; ++cnt; }
            pl->mean_patch_bytes = cnt ? (int)(sum / cnt) : 0;
            if (pl->use_table == 1) {
                int dev = 0, smem_per_sm = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
                TiledShape two;
                const bool fits = shape_for(pl, 4, 2, 1, smem_per_sm, &two) && 3LL * 4 * pl->mean_patch_bytes <= two.ring;
                pl->teams_multi_pref = fits ? 2 : 1;
            }
        }
        std::vector<int> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        // rows are grouped in bands of 128: inside a band the tiles keep their (view, tile) order, so the entries the
        // grid loads at the same moment are horizontal neighbours of one view.  Measured on B200 (16 x 8K frames, 12
        // views, DRAM reads per launch / Gpix/s): bicubic 2.02 GB / 145.5 sorted by exact row, 1.60 GB / 148.5 in bands
        // of 128 rows (unique source bytes: 1.27 GB); bilinear 4.71 GB / 272.8 -> 3.71 GB / 281.8 (R360_ORDER_BAND)
        const int band = std::max(1, env_int("R360_ORDER_BAND", 128));
        if (band > 1)
            for (int i = 0; i < n; ++i)
                if (keys[i] < INT_MAX - 1) keys[i] = (keys[i] & ~0xFFFFFF) | ((keys[i] & 0xFFFFFF) / band);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return keys[a] < keys[b]; });
        std::vector<int2> order;
        order.reserve(n);
        pl->n_order = 0;
        for (int i = 0; i < n && keys[idx[i]] != INT_MAX; ++i) {
            order.push_back(make_int2(idx[i] / pl->n_tiles, idx[i] % pl->n_tiles));
            if (keys[idx[i]] < INT_MAX - 1) ++pl->n_order;
        }
        pl->n_large = (int)order.size() - pl->n_order;               // they sort behind the main walk
        if (!order.empty() &&
            (e = cudaMemcpyAsync(pl->d_order, order.data(), sizeof(int2) * order.size(), cudaMemcpyHostToDevice, s)) != cudaSuccess)
            return fail(e, "cudaMemcpyAsync(order)");
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "cudaStreamSynchronize");
    }
    *plan_out = pl;
    return R360_OK;
}

// Work counters of the tiled kernel's item queue: every launch takes the next one of a pool (zeroed in the launch's
// stream right before the kernel), so launches of one plan from several host threads / streams do not share a counter.
constexpr int kWorkCounters = 4096;
__device__ unsigned int g_work_counters[kWorkCounters];
std::atomic<unsigned> g_next_work_counter{0};

unsigned int* next_work_counter() {
    static std::mutex mu;
    static unsigned int* base[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!base[dev]) {
            void* p = nullptr;
            if (cudaGetSymbolAddress(&p, g_work_counters) != cudaSuccess) return nullptr;
            base[dev] = static_cast<unsigned int*>(p);
        }
    }
    return base[dev] + g_next_work_counter.fetch_add(1, std::memory_order_relaxed) % kWorkCounters;
}

// Launch shape of the tiled kernel for one call: frames per work item, consumer teams per block, blocks per SM
// and the shared-memory ring that is left.  Every fast tile of the plan fits the ring of every shape chosen here
// (ring >= plan->patch_budget); what a shape changes is how many frames share one coordinate / weight set-up.


bool shape_for(const r360_plan* pl, int fr, int teams, int ctas, int smem_per_sm, TiledShape* out) {
    const int fixed = kTiledFixedSmem + table_bytes(pl->use_table) + teams * fr * pl->out_stage_bytes + 128;
    const int per_block = smem_per_sm / ctas - 1024;               // 1 KB per block is reserved by the driver
    int ring = (per_block - fixed) & ~127;
    if (ring > kMaxRingBytes) ring = kMaxRingBytes;
    if (ring < pl->patch_budget) return false;
    out->fr = fr; out->teams = teams; out->ctas = ctas; out->ring = ring; out->smem = fixed + ring;
    // an item takes all its frames at once when it leaves room for the next one to load behind it (measured on B200,
    // 8K -> 12 x 1600^2, two frames per item: 50 % of the ring 140.9 / 279.0 Gpix/s bicubic / bilinear, 75 % 143.1 / 280.1)
    // (four-frame bilinear items: 285.9 Gpix/s at 75 %, 289.2 at 100 %)
    const int pct = std::min(100, std::max(10, env_int("R360_MULTI_PCT", pl->multi_pct_pref)));
    out->multi_budget = (int)((long long)ring * pct / 100);
    return true;
}

TiledShape choose_shape(const r360_plan* pl, int n_groups) {
    int dev = 0, smem_per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    // defaults measured on B200 (profiles/README.md, round 2); the environment overrides are for experiments
    int fr = n_groups >= 4 ? pl->frames_pref : n_groups >= 2 ? std::min(2, pl->frames_pref) : 1;
    fr = env_int("R360_FRAMES", fr);
    if (fr != 1 && fr != 2 && fr != 4) fr = 1;
    if (fr > n_groups) fr = n_groups >= 2 ? 2 : 1;
    // two teams exist for two- and four-frame items (the instantiations the library carries); the plan's preference
    // applies to its preferred frame count
    int teams = fr >= 2 ? std::min(kMaxTeams, std::max(1, env_int("R360_TEAMS", fr == pl->frames_pref ? pl->teams_multi_pref : 1))) : 1;
    int ctas = std::max(1, env_int("R360_TILED_CTAS_PER_SM", fr == pl->frames_pref ? pl->ctas_multi_pref : teams == 2 ? 1 : pl->ctas_per_sm));
    TiledShape s;
    for (;;) {
        if (shape_for(pl, fr, teams, ctas, smem_per_sm, &s)) return s;
        if (ctas > 1) --ctas;
        else if (teams > 1) --teams;
        else if (fr > 1) { fr /= 2; teams = 1; ctas = pl->ctas_per_sm; }
        else break;
    }
    // the plan's own single-frame shape always fits (it defined the patch budget)
    s.fr = 1; s.teams = 1; s.ctas = pl->ctas_per_sm; s.ring = pl->ring_bytes; s.smem = pl->smem_bytes; s.multi_budget = 0;
    return s;
}

struct TiledLauncher {
    const r360_plan* pl; const r360_images* src; const r360_images* dst; cudaStream_t s;

    template <int INTERP, typename TIn, typename TOut, int FR, int TEAMS>
    int launch(const TiledParams& Q, const TensorMaps& maps, const TiledShape& shape, long long grid) {
        auto kernel = remap_tiled_kernel<INTERP, TIn, TOut, FR, TEAMS>;
        {
            // opt this instantiation in to the device's full shared memory once per device (the attribute is
            // per context; plans of different channel counts need different amounts of the same kernel)
            static std::mutex mu;
            static bool configured[64] = {};
            int dev = 0;
            R360_CUDA(cudaGetDevice(&dev));
            std::lock_guard<std::mutex> lock(mu);
            if (dev >= 0 && dev < 64 && !configured[dev]) {
                int optin = 0;
                R360_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
                R360_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
                configured[dev] = true;
            }
        }
        kernel<<<dim3((unsigned)grid), tiled_threads(TEAMS), shape.smem, s>>>(Q, maps);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        R360_CUDA(cudaGetLastError());
        return R360_OK;
    }

    template <int PROJ, int INTERP, typename TIn, typename TOut> int run() {
        const LaunchParams& lp = pl->pr.lp;
        const int n_groups = src->count / pl->pr.n_lenses;
        const TiledShape main_shape = choose_shape(pl, n_groups);
        const TiledShape& shape = main_shape;
        TiledParams T;
        std::memset(&T, 0, sizeof(T));
        T.src = {static_cast<unsigned char*>(src->data), src->pitch_bytes, src->image_stride_bytes, src->width, src->height};
        T.dst = {static_cast<unsigned char*>(dst->data), dst->pitch_bytes, dst->image_stride_bytes, dst->width, dst->height};
        T.channels = lp.channels; T.n_views = pl->pr.n_views; T.n_groups = n_groups;
        T.n_lenses = pl->pr.n_lenses; T.tiles_x = pl->tiles_x; T.tiles_y = pl->tiles_y;
        T.out_stage_bytes = pl->out_stage_bytes; T.ring_bytes = shape.ring;
        T.bulk_store_ok = pl->bulk_store_ok; T.use_table = pl->use_table; T.border_value = lp.border_value;
        T.frames_per_item = shape.fr; T.multi_budget = shape.multi_budget;
        T.dst_fstride = (long long)pl->pr.n_views * T.dst.image_stride;
        T.plans = pl->d_plans;
        T.order = pl->d_order; T.n_order = pl->n_order;
        T.coords = pl->d_coords;
        T.l2_policy = env_int("R360_L2_POLICY", 0);

        // Experiment (R360_FALLBACK_OVERLAP=1): the fallback tiles beside the tiled kernel instead of after it -- a side
        // stream forked off `s` here and joined after both launches (capturable; the events are made per call because
        // one plan may be launched from several host threads at once).  Off by default: next to the persistent blocks
        // only one 128-thread fallback block fits per SM, and the latency-bound fallback kernel then takes longer than
        // the tiled kernel it was meant to hide behind (B200, 19 views incl. poles: 121 Gpix/s serial, 72 overlapped).
        struct EventPair {
            cudaEvent_t fork = nullptr, join = nullptr;
            ~EventPair() { if (fork) cudaEventDestroy(fork); if (join) cudaEventDestroy(join); }   // deferred by the runtime
        } ev;
        cudaStream_t fb_stream = s;
        bool forked = false;
        if (pl->n_fallback > 0 && pl->n_order + pl->n_large > 0 && pl->side_stream && env_int("R360_FALLBACK_OVERLAP", 0) != 0 &&
            cudaEventCreateWithFlags(&ev.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ev.join, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventRecord(ev.fork, s) == cudaSuccess && cudaStreamWaitEvent(pl->side_stream, ev.fork, 0) == cudaSuccess) {
            fb_stream = pl->side_stream;
            forked = true;
        }

        // two passes over the same kernel: the main walk in the shape chosen above, then the tiles with large
        // patches, one frame per item, one block per SM with the largest ring (their list follows the main one)
        TiledShape large_shape = shape;
        if (pl->n_large > 0) {
            int dev = 0, smem_per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
            if (!shape_for(pl, 1, 1, 1, smem_per_sm, &large_shape) || large_shape.ring < pl->patch_budget_large) return R360_E_UNSUPPORTED;
        }
        for (int pass = 0; pass < 2; ++pass) {
        const TiledShape& shape = pass == 0 ? main_shape : large_shape;
        const int n_items = pass == 0 ? pl->n_order : pl->n_large;
        if (n_items == 0) continue;
        T.order = pl->d_order + (pass == 0 ? 0 : pl->n_order); T.n_order = n_items;
        T.ring_bytes = shape.ring; T.frames_per_item = shape.fr; T.multi_budget = shape.multi_budget;
        // work items are indexed with 32-bit ints inside the kernel: chunk the groups if needed
        const long long per_group = n_items;
        int max_groups = (int)std::max<long long>(1, (1LL << 30) / per_group);
        if (max_groups > shape.fr) max_groups -= max_groups % shape.fr;          // chunks hold whole frame blocks
        for (int g0 = 0; g0 < n_groups; g0 += max_groups) {
            TiledParams Q = T;
            Q.n_groups = n_groups - g0 < max_groups ? n_groups - g0 : max_groups;
            Q.src.data += (long long)g0 * pl->pr.n_lenses * Q.src.image_stride;
            Q.dst.data += (long long)g0 * pl->pr.n_views * Q.dst.image_stride;
            const long long total = (long long)((Q.n_groups + shape.fr - 1) / shape.fr) * per_group;
            long long grid = (long long)pl->sm_count * shape.ctas;
            if (grid > total) grid = total;
            TensorMaps maps;
            std::memset(&maps, 0, sizeof(maps));
            if (pl->tensor_ok) {
                r360_images chunk = *src;
                chunk.data = Q.src.data;
                chunk.count = Q.n_groups * pl->pr.n_lenses;
                std::lock_guard<std::mutex> lock(pl->tm_mutex);
                const r360_plan::TmEntry* hit = nullptr;
                for (const auto& e : pl->tm_cache)
                    if (e.data == chunk.data && e.count == chunk.count && e.stride == chunk.image_stride_bytes) hit = &e;
                if (!hit) {
                    r360_plan::TmEntry e;
                    e.data = chunk.data; e.count = chunk.count; e.stride = chunk.image_stride_bytes;
                    const int erc = encode_tensor_maps(chunk, pl->box_family, &e.maps);
                    if (erc != R360_OK) return erc;
                    if (pl->tm_cache.size() >= 8) pl->tm_cache.erase(pl->tm_cache.begin());
                    pl->tm_cache.push_back(e);
                    hit = &pl->tm_cache.back();
                }
                maps = hit->maps;
            }
            Q.work_counter = next_work_counter();
            if (!Q.work_counter) return R360_E_CUDA;
            R360_CUDA(cudaMemsetAsync(Q.work_counter, 0, sizeof(unsigned int), s));
            int rc;
            if (shape.fr == 4 && shape.teams == 2) rc = launch<INTERP, TIn, TOut, 4, 2>(Q, maps, shape, grid);
            else if (shape.fr == 4) rc = launch<INTERP, TIn, TOut, 4, 1>(Q, maps, shape, grid);
            else if (shape.fr == 2 && shape.teams == 2) rc = launch<INTERP, TIn, TOut, 2, 2>(Q, maps, shape, grid);
            else if (shape.fr == 2) rc = launch<INTERP, TIn, TOut, 2, 1>(Q, maps, shape, grid);
            else rc = launch<INTERP, TIn, TOut, 1, 1>(Q, maps, shape, grid);
            if (rc != R360_OK) return rc;
        }
        }

        if (pl->n_fallback > 0) {
            FallbackParams F;
            std::memset(&F, 0, sizeof(F));
            F.lp = lp;
            F.lp.src = T.src; F.lp.dst = T.dst; F.lp.view_base = 0; F.lp.n_views = pl->pr.n_views;
            F.views = pl->d_views; F.list = pl->d_fallback; F.tiles_x = pl->tiles_x;
            // frames per block: the projection is shared by the frames of a block and four frames are sampled
            // together, so whole multiples of four while the grid keeps a few waves of blocks
            constexpr int kParts = kTile / kFallbackRows;
            const long long blocks_wanted = 16LL * pl->sm_count;
            long long fpb = (long long)pl->n_fallback * kParts * n_groups / blocks_wanted;
            fpb = std::max<long long>(kFallbackFrames, fpb / kFallbackFrames * kFallbackFrames);
            fpb = std::min<long long>(fpb, n_groups);
            fpb = std::max<long long>(fpb, (n_groups + 65534) / 65535);
            F.n_groups = n_groups; F.frames_per_block = (int)fpb;
            remap_fallback_kernel<PROJ, INTERP, TIn, TOut><<<dim3(pl->n_fallback * kParts, (unsigned)((n_groups + fpb - 1) / fpb)),
                                                             kFallbackThreads, 0, fb_stream>>>(F);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            R360_CUDA(cudaGetLastError());
            if (forked) {
                R360_CUDA(cudaEventRecord(ev.join, fb_stream));
                R360_CUDA(cudaStreamWaitEvent(s, ev.join, 0));
            }
        }
        return R360_OK;
    }
};

// The maps a projection produces, written out for tests (float32 as cv2 would get them + float64).
int launch_coords(CoordParams& p, int proj, const std::vector<ViewDev>& views, cudaStream_t s) {
    const int n_views = (int)views.size();
    for (int v0 = 0; v0 < n_views; v0 += kMaxViewsPerLaunch) {
        p.view_base = v0;
        p.n_views = n_views - v0 < kMaxViewsPerLaunch ? n_views - v0 : kMaxViewsPerLaunch;
        for (int v = 0; v < p.n_views; ++v) p.views[v] = views[v0 + v];
        dim3 grid((p.out_w + 31) / 32, (p.out_h + 7) / 8, p.n_views);
        coords_kernel<<<grid, 256, 0, s>>>(p, proj);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        R360_CUDA(cudaGetLastError());
    }
    return R360_OK;
}

bool same_layout(const r360_images& a, const r360_images& b, bool check_stride) {
    return a.width == b.width && a.height == b.height && a.channels == b.channels && a.dtype == b.dtype &&
           a.pitch_bytes == b.pitch_bytes && (!check_stride || a.image_stride_bytes == b.image_stride_bytes);
}

}  // namespace

// ---- exported functions ----------------------------------------------------------------------------

extern "C" {

int r360_abi_version(void) { return R360_ABI_VERSION; }

const char* r360_error_string(int code) {
    switch (code) {
        case R360_OK: return "ok";
        case R360_E_INVALID_ARG: return "invalid argument";
        case R360_E_UNSUPPORTED: return "unsupported dtype/channel/interpolation combination";
        case R360_E_CUDA: return "CUDA runtime error";
        case R360_E_NO_DEVICE: return "no usable CUDA device (sm_100 required)";
        case R360_E_TOO_MANY: return "too many source slots";
        default: return "unknown error";
    }
}

const char* r360_last_cuda_error(void) { return tl_cuda_error; }

void r360_default_options(r360_options* opt) {
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->interp = R360_CUBIC;            // the reference's default in both tools (PC:730, DF:232)
    opt->convention = R360_CONV_HALFPIXEL;
    opt->path = R360_PATH_AUTO;
    opt->fill_invalid = 1;               // DF: --mask-outside-model defaults to on (DF:258)
    opt->border_value = 0.0;             // DF: --mask-value 0 (DF:262)
    opt->out_dtype = -1;
}

int r360_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    R360_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    R360_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return prop.major == 10 ? R360_OK : R360_E_NO_DEVICE;
}

int r360_remap_erp(const r360_images* src, const r360_images* dst, const r360_view* views, int32_t n_views,
                   const r360_options* opt, void* stream) {
    std::vector<ViewDev> vd;
    if (!dst) return R360_E_INVALID_ARG;
    const int rc = build_persp_views(views, n_views, dst->width, dst->height, &vd);
    return rc != R360_OK ? rc : remap_direct(kProjErp, src, dst, nullptr, 1, vd, opt, stream);
}

int r360_remap_fisheye(const r360_images* src, const r360_images* dst, const r360_fisheye_calib* calib,
                       int32_t n_lenses, const r360_view* views, int32_t n_views, const r360_options* opt,
                       void* stream) {
    std::vector<ViewDev> vd;
    if (!dst) return R360_E_INVALID_ARG;
    const int rc = build_persp_views(views, n_views, dst->width, dst->height, &vd);
    return rc != R360_OK ? rc : remap_direct(kProjFisheye, src, dst, calib, n_lenses, vd, opt, stream);
}

int r360_remap_undistort(const r360_images* src, const r360_images* dst, const r360_fisheye_calib* calib,
                         int32_t n_lenses, const r360_undistort* items, int32_t n_items, const r360_options* opt,
                         void* stream) {
    std::vector<ViewDev> vd;
    const int rc = build_undistort_views(items, n_items, calib, n_lenses, &vd);
    return rc != R360_OK ? rc : remap_direct(kProjUndistort, src, dst, calib, n_lenses, vd, opt, stream);
}

int r360_apply_lut(const r360_images* src, const r360_images* dst, const r360_lut3d* lut,
                   int32_t output_space, int32_t channel_order, void* stream) {
    int rc;
    if ((rc = check_images(src)) != R360_OK) return rc;
    if ((rc = check_images(dst)) != R360_OK) return rc;
    if (!lut || !lut->table_device || lut->size < 2 || lut->size > 256) return R360_E_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(lut->table_device) % 16) return R360_E_INVALID_ARG;
    if (output_space != R360_LUT_PASSTHROUGH && output_space != R360_LUT_SRGB) return R360_E_INVALID_ARG;
    if (channel_order != R360_ORDER_BGR && channel_order != R360_ORDER_RGB) return R360_E_INVALID_ARG;
    if (src->channels < 3) return R360_E_INVALID_ARG;                      // DF:693-697
    if (src->width != dst->width || src->height != dst->height || src->channels != dst->channels ||
        src->dtype != dst->dtype || src->count != dst->count)
        return R360_E_INVALID_ARG;
    if (src->dtype == R360_F16) return R360_E_UNSUPPORTED;
    if (src->height > 65535 || src->count > 65535) return R360_E_INVALID_ARG;
    LutParams p;
    std::memset(&p, 0, sizeof(p));
    p.src = {static_cast<unsigned char*>(src->data), src->pitch_bytes, src->image_stride_bytes, src->width, src->height};
    p.dst = {static_cast<unsigned char*>(dst->data), dst->pitch_bytes, dst->image_stride_bytes, dst->width, dst->height};
    p.channels = src->channels; p.n_images = src->count; p.size = lut->size;
    p.to_srgb = output_space == R360_LUT_SRGB; p.rgb_order = channel_order == R360_ORDER_RGB;
    for (int c = 0; c < 3; ++c) {
        p.dmin[c] = lut->domain_min[c];
        p.span[c] = lut->domain_max[c] - lut->domain_min[c];                 // float32 subtraction, as DF:640
        if (!(p.span[c] > 0.0f)) return R360_E_INVALID_ARG;                 // DF:556-558
    }
    p.table = reinterpret_cast<const float4*>(lut->table_device);
    if ((rc = ensure_device_ready()) != R360_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const dim3 grid((src->width + 255) / 256, src->height, src->count);
    if (src->dtype == R360_U8) lut_kernel<uint8_t><<<grid, 256, 0, s>>>(p);
    else if (src->dtype == R360_U16) lut_kernel<uint16_t><<<grid, 256, 0, s>>>(p);
    else lut_kernel<float><<<grid, 256, 0, s>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    R360_CUDA(cudaGetLastError());
    return R360_OK;
}

int r360_convert_color(const r360_images* src, const r360_images* dst, const r360_color_convert* cc,
                       int32_t channel_order, void* stream) {
    int rc;
    if ((rc = check_images(src)) != R360_OK) return rc;
    if ((rc = check_images(dst)) != R360_OK) return rc;
    if (!cc) return R360_E_INVALID_ARG;
    if (cc->in_trc < R360_TRC_BT709 || cc->in_trc > R360_TRC_LINEAR || cc->out_trc < R360_TRC_BT709 ||
        cc->out_trc > R360_TRC_LINEAR)
        return R360_E_INVALID_ARG;
    if (channel_order != R360_ORDER_BGR && channel_order != R360_ORDER_RGB) return R360_E_INVALID_ARG;
    if (src->channels < 3) return R360_E_INVALID_ARG;
    if (src->width != dst->width || src->height != dst->height || src->channels != dst->channels ||
        src->dtype != dst->dtype || src->count != dst->count)
        return R360_E_INVALID_ARG;
    if (src->dtype == R360_F16) return R360_E_UNSUPPORTED;
    if (src->height > 65535 || src->count > 65535) return R360_E_INVALID_ARG;
    ColorConvertParams p;
    std::memset(&p, 0, sizeof(p));
    p.src = {static_cast<unsigned char*>(src->data), src->pitch_bytes, src->image_stride_bytes, src->width, src->height};
    p.dst = {static_cast<unsigned char*>(dst->data), dst->pitch_bytes, dst->image_stride_bytes, dst->width, dst->height};
    p.channels = src->channels; p.n_images = src->count;
    p.in_trc = cc->in_trc; p.out_trc = cc->out_trc; p.rgb_order = channel_order == R360_ORDER_RGB;
    for (int k = 0; k < 9; ++k) {
        if (!std::isfinite(cc->matrix[k])) return R360_E_INVALID_ARG;
        p.m[k] = cc->matrix[k];
    }
    if ((rc = ensure_device_ready()) != R360_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const dim3 grid((src->width + 255) / 256, src->height, src->count);
    if (src->dtype == R360_U8) color_convert_kernel<uint8_t><<<grid, 256, 0, s>>>(p);
    else if (src->dtype == R360_U16) color_convert_kernel<uint16_t><<<grid, 256, 0, s>>>(p);
    else color_convert_kernel<float><<<grid, 256, 0, s>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    R360_CUDA(cudaGetLastError());
    return R360_OK;
}

int r360_coords(int32_t src_w, int32_t src_h, const r360_fisheye_calib* calib, int32_t n_lenses,
                const r360_view* views, int32_t n_views, int32_t out_w, int32_t out_h,
                const r360_options* opt_in, float* map_x32, float* map_y32, double* map_x64, double* map_y64,
                uint8_t* valid, void* stream) {
    if (!views || n_views <= 0 || out_w <= 0 || out_h <= 0) return R360_E_INVALID_ARG;
    if (!calib && (src_w <= 0 || src_h <= 0)) return R360_E_INVALID_ARG;
    if (calib && n_lenses > R360_MAX_LENSES) return R360_E_TOO_MANY;
    if (calib && n_lenses < 1) return R360_E_INVALID_ARG;
    r360_options opt;
    if (opt_in) opt = *opt_in; else r360_default_options(&opt);
    int rc;
    if ((rc = ensure_device_ready()) != R360_OK) return rc;
    CoordParams p;
    std::memset(&p, 0, sizeof(p));
    p.out_w = out_w; p.out_h = out_h;
    p.x32 = map_x32; p.y32 = map_y32; p.x64 = map_x64; p.y64 = map_y64; p.valid = valid;
    const int proj = calib ? kProjFisheye : kProjErp;
    if (calib) for (int l = 0; l < n_lenses; ++l) make_lens(calib[l], &p.lens[l]);
    else make_erp(src_w, src_h, opt.convention, &p.erp);
    for (int v = 0; v < n_views; ++v)
        if (calib && (views[v].src_slot < 0 || views[v].src_slot >= n_lenses)) return R360_E_INVALID_ARG;
    std::vector<ViewDev> vd;
    if ((rc = build_persp_views(views, n_views, out_w, out_h, &vd)) != R360_OK) return rc;
    return launch_coords(p, proj, vd, static_cast<cudaStream_t>(stream));
}

int r360_coords_undistort(const r360_fisheye_calib* calib, int32_t n_lenses, const r360_undistort* items,
                          int32_t n_items, int32_t out_w, int32_t out_h, float* map_x32, float* map_y32,
                          double* map_x64, double* map_y64, uint8_t* valid, void* stream) {
    if (out_w <= 0 || out_h <= 0) return R360_E_INVALID_ARG;
    std::vector<ViewDev> vd;
    int rc;
    if ((rc = build_undistort_views(items, n_items, calib, n_lenses, &vd)) != R360_OK) return rc;
    if ((rc = ensure_device_ready()) != R360_OK) return rc;
    CoordParams p;
    std::memset(&p, 0, sizeof(p));
    p.out_w = out_w; p.out_h = out_h;
    p.x32 = map_x32; p.y32 = map_y32; p.x64 = map_x64; p.y64 = map_y64; p.valid = valid;
    for (int l = 0; l < n_lenses; ++l) make_lens(calib[l], &p.lens[l]);
    return launch_coords(p, kProjUndistort, vd, static_cast<cudaStream_t>(stream));
}

size_t r360_plan_workspace_bytes(int32_t n_views, int32_t out_w, int32_t out_h) {
    if (n_views <= 0 || out_w <= 0 || out_h <= 0) return 0;
    return workspace_layout(n_views, out_w, out_h).total;
}

int r360_plan_create_erp(const r360_images* src_layout, const r360_images* dst_layout, const r360_view* views,
                         int32_t n_views, const r360_options* opt, void* workspace_device, size_t workspace_bytes,
                         void* stream, r360_plan** plan_out) {
    std::vector<ViewDev> vd;
    if (plan_out) *plan_out = nullptr;
    if (!dst_layout) return R360_E_INVALID_ARG;
    const int rc = build_persp_views(views, n_views, dst_layout->width, dst_layout->height, &vd);
    if (rc != R360_OK) return rc;
    return plan_create(kProjErp, src_layout, dst_layout, nullptr, 1, vd, opt, workspace_device,
                       workspace_bytes, stream, plan_out);
}

int r360_plan_create_fisheye(const r360_images* src_layout, const r360_images* dst_layout,
                             const r360_fisheye_calib* calib, int32_t n_lenses, const r360_view* views,
                             int32_t n_views, const r360_options* opt, void* workspace_device,
                             size_t workspace_bytes, void* stream, r360_plan** plan_out) {
    std::vector<ViewDev> vd;
    if (plan_out) *plan_out = nullptr;
    if (!dst_layout) return R360_E_INVALID_ARG;
    const int rc = build_persp_views(views, n_views, dst_layout->width, dst_layout->height, &vd);
    if (rc != R360_OK) return rc;
    return plan_create(kProjFisheye, src_layout, dst_layout, calib, n_lenses, vd, opt,
                       workspace_device, workspace_bytes, stream, plan_out);
}

int r360_plan_create_undistort(const r360_images* src_layout, const r360_images* dst_layout,
                               const r360_fisheye_calib* calib, int32_t n_lenses, const r360_undistort* items,
                               int32_t n_items, const r360_options* opt, void* workspace_device,
                               size_t workspace_bytes, void* stream, r360_plan** plan_out) {
    std::vector<ViewDev> vd;
    if (plan_out) *plan_out = nullptr;
    const int rc = build_undistort_views(items, n_items, calib, n_lenses, &vd);
    if (rc != R360_OK) return rc;
    return plan_create(kProjUndistort, src_layout, dst_layout, calib, n_lenses, vd, opt,
                       workspace_device, workspace_bytes, stream, plan_out);
}

int r360_plan_info(const r360_plan* plan, int32_t* tiles_per_view, int32_t* n_fallback_tiles) {
    if (!plan) return R360_E_INVALID_ARG;
    if (tiles_per_view) *tiles_per_view = plan->n_tiles;
    if (n_fallback_tiles) *n_fallback_tiles = plan->n_fallback;
    return R360_OK;
}

int r360_plan_info_maps(const r360_plan* plan, int32_t* n_map_tiles, int32_t* map_pool_tiles) {
    if (!plan) return R360_E_INVALID_ARG;
    if (n_map_tiles) *n_map_tiles = plan->n_coords;
    if (map_pool_tiles) *map_pool_tiles = plan->coords_capacity;
    return R360_OK;
}

int r360_remap_planned(const r360_plan* plan, const r360_images* src, const r360_images* dst, void* stream) {
    if (!plan) return R360_E_INVALID_ARG;
    int rc;
    if ((rc = check_images(src)) != R360_OK) return rc;
    if ((rc = check_images(dst)) != R360_OK) return rc;
    if (!same_layout(*src, plan->src_layout, src->count > 1) || !same_layout(*dst, plan->dst_layout, dst->count > 1))
        return R360_E_INVALID_ARG;
    if (src->count % plan->pr.n_lenses) return R360_E_INVALID_ARG;
    if ((int64_t)dst->count != (int64_t)(src->count / plan->pr.n_lenses) * plan->pr.n_views) return R360_E_INVALID_ARG;
    // the plan assumed 16-byte aligned bases wherever it chose bulk copies
    if (plan->bulk_load_ok && !aligned16(src->data, src->pitch_bytes, src->count > 1 ? src->image_stride_bytes : 0))
        return R360_E_INVALID_ARG;
    if (plan->bulk_store_ok && !aligned16(dst->data, dst->pitch_bytes, dst->count > 1 ? dst->image_stride_bytes : 0))
        return R360_E_INVALID_ARG;
    TiledLauncher L{plan, src, dst, static_cast<cudaStream_t>(stream)};
    return dispatch(plan->pr.proj, plan->pr.interp, plan->pr.in_dt, plan->pr.out_dt, L);
}

int r360_plan_coords(const r360_plan* plan, float* map_x32, float* map_y32, double* map_x64, double* map_y64,
                     uint8_t* valid, void* stream) {
    if (!plan || !map_x32 || !map_y32 || !map_x64 || !map_y64) return R360_E_INVALID_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // every pixel from the direct projection first, then the fast tiles overwrite theirs with
    // what the polynomial path computes
    CoordParams p;
    std::memset(&p, 0, sizeof(p));
    p.out_w = plan->dst_layout.width; p.out_h = plan->dst_layout.height;
    p.x32 = map_x32; p.y32 = map_y32; p.x64 = map_x64; p.y64 = map_y64; p.valid = valid;
    p.erp = plan->pr.lp.erp;
    std::memcpy(p.lens, plan->pr.lp.lens, sizeof(p.lens));
    const int n_views = plan->pr.n_views;
    const int crc = launch_coords(p, plan->pr.proj, plan->views, s);
    if (crc != R360_OK) return crc;
    TiledCoordParams T;
    std::memset(&T, 0, sizeof(T));
    T.out_w = p.out_w; T.out_h = p.out_h; T.tiles_x = plan->tiles_x; T.tiles_y = plan->tiles_y;
    T.period32 = 32.0 * plan->src_layout.width;
    T.plans = plan->d_plans;
    T.coords = plan->d_coords;
    T.x32 = map_x32; T.y32 = map_y32; T.x64 = map_x64; T.y64 = map_y64; T.valid = valid;
    coords_tiled_kernel<<<dim3(plan->n_tiles, n_views), 256, 0, s>>>(T);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    R360_CUDA(cudaGetLastError());
    return R360_OK;
}

void r360_plan_destroy(r360_plan* plan) { delete plan; }

int64_t r360_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Test hook (not part of the stable ABI): copies the host-built weight tables out so the CPU
// test-suite can compare them with the oracle's without a GPU.
int r360_debug_weight_tables(int16_t* cubic_fixed_16384, float* cubic_1d_128) {
    static WeightTables t;
    build_weight_tables(&t);
    if (cubic_fixed_16384) std::memcpy(cubic_fixed_16384, t.cubic_fixed, sizeof(t.cubic_fixed));
    if (cubic_1d_128) std::memcpy(cubic_1d_128, t.cubic_1d, sizeof(t.cubic_1d));
    return R360_OK;
}

int r360_debug_weight_tables_lanczos4(int16_t* fixed_65536, float* one_d_256) {
    static WeightTables t;
    build_weight_tables(&t);
    if (fixed_65536) std::memcpy(fixed_65536, t.lanczos_fixed, sizeof(t.lanczos_fixed));
    if (one_d_256) std::memcpy(one_d_256, t.lanczos_1d, sizeof(t.lanczos_1d));
    return R360_OK;
}

// Test hook: drives the shared-memory ring allocator (PatchRing, r360_tiled.cuh) on the host with a
// sequence of patch sizes and at most `slots` allocations in flight; returns -1 if no two live
// allocations ever overlap and none leaves the ring, else the index of the offending request.
int r360_debug_ring_check(const int32_t* sizes, int32_t n, int32_t capacity, int32_t slots) {
    PatchRing ring;
    std::vector<int> off(n), charge(n);
    int oldest = 0;
    for (int k = 0; k < n; ++k) {
        if (sizes[k] > capacity) return k;
        while ((k - oldest) >= slots || !ring.try_alloc(sizes[k], capacity, off[k], charge[k])) {
            if (oldest >= k) return k;                 // nothing left to release and still no room
            ring.release(charge[oldest], capacity);
            ++oldest;
        }
        if (off[k] < 0 || off[k] + sizes[k] > capacity) return k;
        for (int j = oldest; j < k; ++j)
            if (sizes[j] && sizes[k] && !(off[k] + sizes[k] <= off[j] || off[j] + sizes[j] <= off[k])) return k;
    }
    return -1;
}

}  // extern "C"
