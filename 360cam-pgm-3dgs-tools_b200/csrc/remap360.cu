// remap360 C ABI (include/remap360.h): argument checking, view/lens preparation in float64,
// weight-table upload and kernel dispatch.  Device code lives in r360_direct.cuh (per-pixel
// float64 path) and r360_tiled.cuh (tile-staged fast path).
#include "../../include/remap360.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "r360_common.cuh"
#include "r360_direct.cuh"
#include "r360_tiled.cuh"

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
    o->slot = v.src_slot;
    o->pad = 0;
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

// Shared memory per block for the tiled kernel: aim for 4 resident blocks per SM.
constexpr int kSmemPerBlockTarget = 56 * 1024 - 1024;
constexpr int kTiledFixedSmem = 2048;

template <int PROJ, int INTERP, typename TIn, typename TOut>
int launch_tiled(const LaunchParams& p, cudaStream_t stream, const CoordParams* dbg = nullptr) {
    TiledParams T;
    std::memset(&T, 0, sizeof(T));
    T.lp = p;
    T.tiles_x = (p.dst.width + kTile - 1) / kTile;
    T.tiles_y = (p.dst.height + kTile - 1) / kTile;
    T.out_stage_bytes = kTile * kTile * p.channels * (int)sizeof(TOut);
    const int stage = (T.out_stage_bytes + 127) & ~127;
    T.patch_budget = kSmemPerBlockTarget - kTiledFixedSmem - stage;
    if (T.patch_budget < 8192) T.patch_budget = 8192;
    const int smem = kTiledFixedSmem + stage + T.patch_budget;
    auto aligned16 = [](const ImageSetDev& im) {
        return (reinterpret_cast<uintptr_t>(im.data) % 16 == 0) && im.pitch % 16 == 0 && im.image_stride % 16 == 0;
    };
    T.bulk_load_ok = aligned16(p.src) && ((long long)p.src.width * p.channels * sizeof(TIn)) % 16 == 0;
    T.bulk_store_ok = aligned16(p.dst);
    if (dbg) { T.dbg_x32 = dbg->x32; T.dbg_y32 = dbg->y32; T.dbg_x64 = dbg->x64; T.dbg_y64 = dbg->y64; T.dbg_valid = dbg->valid; }
    auto kernel = remap_tiled_kernel<PROJ, INTERP, TIn, TOut>;
    static thread_local const void* configured = nullptr;   // per instantiation (static in a template)
    if (configured != (const void*)kernel) {
        R360_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = (const void*)kernel;
    }
    const int n_tiles = T.tiles_x * T.tiles_y;
    const int max_groups = 65535 / p.n_views;
    for (int g0 = 0; g0 < p.n_groups; g0 += max_groups) {
        TiledParams Q = T;
        const int ng = p.n_groups - g0 < max_groups ? p.n_groups - g0 : max_groups;
        Q.lp.src.data += (long long)g0 * p.n_lenses * p.src.image_stride;
        Q.lp.dst.data += (long long)g0 * p.n_views_total * p.dst.image_stride;
        Q.lp.n_groups = ng;
        dim3 grid(n_tiles, ng * p.n_views, 1);
        kernel<<<grid, 256, smem, stream>>>(Q);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        R360_CUDA(cudaGetLastError());
    }
    return R360_OK;
}

template <int PROJ, typename TIn, typename TOut>
int dispatch_interp(const LaunchParams& p, int interp, bool tiled, cudaStream_t s) {
    switch (interp) {
        case R360_NEAREST: return tiled ? launch_tiled<PROJ, kNearest, TIn, TOut>(p, s) : launch_direct<PROJ, kNearest, TIn, TOut>(p, s);
        case R360_LINEAR: return tiled ? launch_tiled<PROJ, kLinear, TIn, TOut>(p, s) : launch_direct<PROJ, kLinear, TIn, TOut>(p, s);
        case R360_CUBIC: return tiled ? launch_tiled<PROJ, kCubic, TIn, TOut>(p, s) : launch_direct<PROJ, kCubic, TIn, TOut>(p, s);
        default: return R360_E_INVALID_ARG;
    }
}

template <int PROJ>
int dispatch_types(const LaunchParams& p, int in_dt, int out_dt, int interp, bool tiled, cudaStream_t s) {
    if (in_dt == R360_U8 && out_dt == R360_U8) return dispatch_interp<PROJ, uint8_t, uint8_t>(p, interp, tiled, s);
    if (in_dt == R360_U16 && out_dt == R360_U16) return dispatch_interp<PROJ, uint16_t, uint16_t>(p, interp, tiled, s);
    if (in_dt == R360_U16 && out_dt == R360_F16) return dispatch_interp<PROJ, uint16_t, __half>(p, interp, tiled, s);
    if (in_dt == R360_F16 && out_dt == R360_F16) return dispatch_interp<PROJ, __half, __half>(p, interp, tiled, s);
    if (in_dt == R360_F32 && out_dt == R360_F32) return dispatch_interp<PROJ, float, float>(p, interp, tiled, s);
    return R360_E_UNSUPPORTED;
}

int remap_common(int proj, const r360_images* src, const r360_images* dst,
                 const r360_fisheye_calib* calib, int n_lenses,
                 const r360_view* views, int n_views, const r360_options* opt_in, void* stream) {
    int rc;
    if ((rc = check_images(src)) != R360_OK) return rc;
    if ((rc = check_images(dst)) != R360_OK) return rc;
    if (!views || n_views <= 0) return R360_E_INVALID_ARG;
    if (n_lenses < 1) return R360_E_INVALID_ARG;
    if (n_lenses > R360_MAX_LENSES) return R360_E_TOO_MANY;
    if (proj == kProjFisheye && !calib) return R360_E_INVALID_ARG;
    r360_options opt;
    if (opt_in) opt = *opt_in; else r360_default_options(&opt);
    if (opt.interp < R360_NEAREST || opt.interp > R360_CUBIC) return R360_E_INVALID_ARG;
    if (opt.convention != R360_CONV_HALFPIXEL && opt.convention != R360_CONV_V360) return R360_E_INVALID_ARG;
    if (opt.path < R360_PATH_AUTO || opt.path > R360_PATH_TILED) return R360_E_INVALID_ARG;
    if (src->channels != dst->channels) return R360_E_INVALID_ARG;
    if (src->count % n_lenses) return R360_E_INVALID_ARG;
    const int n_groups = src->count / n_lenses;
    if ((int64_t)dst->count != (int64_t)n_groups * n_views) return R360_E_INVALID_ARG;
    const int out_dt = opt.out_dtype < 0 ? src->dtype : opt.out_dtype;
    if (out_dt != dst->dtype) return R360_E_INVALID_ARG;
    for (int v = 0; v < n_views; ++v)
        if (views[v].src_slot < 0 || views[v].src_slot >= n_lenses) return R360_E_INVALID_ARG;
    if ((rc = ensure_device_ready()) != R360_OK) return rc;

    LaunchParams p;
    std::memset(&p, 0, sizeof(p));
    p.src = {static_cast<unsigned char*>(src->data), src->pitch_bytes, src->image_stride_bytes, src->width, src->height};
    p.dst = {static_cast<unsigned char*>(dst->data), dst->pitch_bytes, dst->image_stride_bytes, dst->width, dst->height};
    p.channels = src->channels;
    p.n_views_total = n_views;
    p.n_lenses = n_lenses;
    p.n_groups = n_groups;
    p.fill_invalid = proj == kProjFisheye ? (opt.fill_invalid != 0) : 0;
    double bv = proj == kProjFisheye ? opt.border_value : 0.0;
    if (src->dtype == R360_U8) bv = std::nearbyint(std::fmin(std::fmax(bv, 0.0), 255.0));          // saturate_cast<uchar>
    else if (src->dtype == R360_U16) bv = std::nearbyint(std::fmin(std::fmax(bv, 0.0), 65535.0));  // saturate_cast<ushort>
    p.border_value = (float)bv;
    if (proj == kProjErp) make_erp(src->width, src->height, opt.convention, &p.erp);
    else for (int l = 0; l < n_lenses; ++l) make_lens(calib[l], &p.lens[l]);

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int v0 = 0; v0 < n_views; v0 += kMaxViewsPerLaunch) {
        p.view_base = v0;
        p.n_views = n_views - v0 < kMaxViewsPerLaunch ? n_views - v0 : kMaxViewsPerLaunch;
        for (int v = 0; v < p.n_views; ++v) make_view(views[v0 + v], dst->width, dst->height, &p.views[v]);
        const bool tiled = opt.path != R360_PATH_DIRECT;
        rc = proj == kProjErp ? dispatch_types<kProjErp>(p, src->dtype, out_dt, opt.interp, tiled, s)
                              : dispatch_types<kProjFisheye>(p, src->dtype, out_dt, opt.interp, tiled, s);
        if (rc != R360_OK) return rc;
    }
    return R360_OK;
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
    return remap_common(kProjErp, src, dst, nullptr, 1, views, n_views, opt, stream);
}

int r360_remap_fisheye(const r360_images* src, const r360_images* dst, const r360_fisheye_calib* calib,
                       int32_t n_lenses, const r360_view* views, int32_t n_views, const r360_options* opt,
                       void* stream) {
    return remap_common(kProjFisheye, src, dst, calib, n_lenses, views, n_views, opt, stream);
}

int r360_coords(int32_t src_w, int32_t src_h, const r360_fisheye_calib* calib, int32_t n_lenses,
                const r360_view* views, int32_t n_views, int32_t out_w, int32_t out_h,
                const r360_options* opt_in, float* map_x32, float* map_y32, double* map_x64, double* map_y64,
                uint8_t* valid, void* stream) {
    if (!views || n_views <= 0 || out_w <= 0 || out_h <= 0) return R360_E_INVALID_ARG;
    if (!calib && (src_w <= 0 || src_h <= 0)) return R360_E_INVALID_ARG;
    if (calib && (n_lenses < 1 || n_lenses > R360_MAX_LENSES)) return calib && n_lenses > R360_MAX_LENSES ? R360_E_TOO_MANY : R360_E_INVALID_ARG;
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
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (opt.path != R360_PATH_DIRECT && (!map_x32 || !map_y32 || !map_x64 || !map_y64)) return R360_E_INVALID_ARG;
    for (int v0 = 0; v0 < n_views; v0 += kMaxViewsPerLaunch) {
        p.view_base = v0;
        p.n_views = n_views - v0 < kMaxViewsPerLaunch ? n_views - v0 : kMaxViewsPerLaunch;
        for (int v = 0; v < p.n_views; ++v) make_view(views[v0 + v], out_w, out_h, &p.views[v]);
        if (opt.path != R360_PATH_DIRECT) {
            // what the tiled kernels would sample for an 8-bit 3-channel source of this size
            LaunchParams lp;
            std::memset(&lp, 0, sizeof(lp));
            lp.src.width = calib ? (int)calib[0].width : src_w;
            lp.src.height = calib ? (int)calib[0].height : src_h;
            lp.src.pitch = (long long)lp.src.width * 3;
            lp.dst.width = out_w; lp.dst.height = out_h;
            lp.channels = 3; lp.n_views = p.n_views; lp.view_base = v0; lp.n_views_total = n_views;
            lp.n_lenses = calib ? n_lenses : 1; lp.n_groups = 1; lp.fill_invalid = opt.fill_invalid != 0;
            lp.erp = p.erp;
            std::memcpy(lp.lens, p.lens, sizeof(lp.lens));
            std::memcpy(lp.views, p.views, sizeof(lp.views));
            rc = proj == kProjErp ? launch_tiled<kProjErp, kLinear, uint8_t, uint8_t>(lp, s, &p)
                                  : launch_tiled<kProjFisheye, kLinear, uint8_t, uint8_t>(lp, s, &p);
            if (rc != R360_OK) return rc;
            continue;
        }
        dim3 grid((out_w + 31) / 32, (out_h + 7) / 8, p.n_views);
        coords_kernel<<<grid, 256, 0, s>>>(p, proj);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        R360_CUDA(cudaGetLastError());
    }
    return R360_OK;
}

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

}  // extern "C"
