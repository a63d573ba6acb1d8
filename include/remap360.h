/*
 * remap360 -- C ABI of the B200 panorama / dual-fisheye -> perspective remap library.
 *
 * The reference project (Mistral-Yu/360Cam-PGM-3DGS-Tools) has no FFI: its hot path is
 * reached either by spawning `ffmpeg -vf v360=...` (cli_tools/gs360_360PerspCut.py:310-314,
 * executed at :572) or by `cv2.remap` on NumPy maps
 * (cli_tools/gs360_DualFisheyeDistortionCalibration.py:1759-1823 build, :2001-2014 apply).
 * The entry points below are what a Python binding for those two call sites needs; each one
 * names the reference code it stands in for.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every pointer called "device" is CUDA device memory owned by the caller (e.g. a torch
 *     tensor); the library never allocates, frees or retains caller memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     stream-ordered, no call synchronises the device;
 *   - every function returns R360_OK (0) or a negative R360_E_* code and never throws;
 *   - calls are re-entrant and thread-safe (the only global state is an immutable weight
 *     table uploaded once per device under a lock).
 *   - images are interleaved (HWC), 1..4 channels, rows `pitch_bytes` apart.
 */
#ifndef REMAP360_H
#define REMAP360_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R360_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------------------- */
enum {
    R360_OK = 0,
    R360_E_INVALID_ARG = -1,   /* null pointer, non-positive size, bad enum                */
    R360_E_UNSUPPORTED = -2,   /* dtype / channel / interpolation combination not built    */
    R360_E_CUDA = -3,          /* a CUDA runtime call failed; see r360_last_cuda_error()   */
    R360_E_NO_DEVICE = -4,     /* no CUDA device / wrong architecture (needs sm_100)       */
    R360_E_TOO_MANY = -5       /* more than R360_MAX_LENSES source slots                   */
};

/* ---- enums --------------------------------------------------------------------------- */
enum { R360_U8 = 0, R360_U16 = 1, R360_F16 = 2, R360_F32 = 3 };

/* What an output view is (r360_view.projection). */
enum { R360_OUT_RECTILINEAR = 0, R360_OUT_FISHEYE = 1 };

/* cv2 names at gs360_DualFisheyeDistortionCalibration.py:59-64; ffmpeg `interp=` at
 * gs360_360PerspCut.py:107-109.  Arithmetic is cv2.remap's (1/32-px fractions, A = -0.75). */
enum { R360_NEAREST = 0, R360_LINEAR = 1, R360_CUBIC = 2, R360_LANCZOS4 = 3 };

/* ERP pixel convention: HALFPIXEL x = (lon/2pi+.5)*W - .5 (in-repo
 * geometry, gs360_GUI.py:419-424 with pixel centres); V360 x = (lon/2pi+.5)*(W-1) (the
 * replaced ffmpeg filter). */
enum { R360_CONV_HALFPIXEL = 0, R360_CONV_V360 = 1 };

/* r360_options.path: r360_remap_erp / r360_remap_fisheye always run the direct per-pixel path
 * (AUTO == DIRECT there); the tiled path needs a plan, see r360_plan_create_*.  TILED is
 * rejected by the plan-less entry points. */
enum { R360_PATH_AUTO = 0, R360_PATH_DIRECT = 1, R360_PATH_TILED = 2 };

#define R360_MAX_LENSES 4

/* ---- plain-data descriptors ------------------------------------------------------------ */

/* A batch of equally sized interleaved images in device memory. */
typedef struct r360_images {
    void*   data;            /* device pointer to image 0                                   */
    int32_t width;           /* pixels                                                      */
    int32_t height;
    int32_t channels;        /* 1..4                                                        */
    int32_t dtype;           /* R360_U8 ...                                                 */
    int64_t pitch_bytes;     /* row stride                                                  */
    int64_t image_stride_bytes; /* distance between consecutive images of the batch         */
    int32_t count;           /* number of images                                            */
    int32_t reserved;
} r360_images;

/* One perspective (rectilinear) view.  Replaces the `w:h:yaw:pitch:roll:h_fov:v_fov` options
 * of the v360 filter string (gs360_360PerspCut.py:310-314) and one entry of
 * build_sfm10_specs (gs360_DualFisheyeDistortionCalibration.py:1258-1307).
 * Rotation order and signs: gs360_GUI.py:351-374 (pitch about X, then yaw about Y; +yaw looks
 * right, +pitch looks up); roll is about the view axis.  All views of one call share the
 * output size of `dst`. */
typedef struct r360_view {
    double  yaw_deg;
    double  pitch_deg;
    double  roll_deg;
    double  hfov_deg;
    double  vfov_deg;
    int32_t src_slot;        /* which image of a source group feeds this view (ERP: 0;
                                dual fisheye: 0 = X lens, 1 = Y lens)                        */
    int32_t projection;      /* R360_OUT_RECTILINEAR (0) or R360_OUT_FISHEYE: the view is an
                                equidistant fisheye image, hfov/vfov being v360's h_fov / v_fov of
                                `output=fisheye` (gs360_360PerspCut.py:375-379, preset fisheyeXY) */
} r360_view;

/* Radial law of a fisheye source (r360_fisheye_calib.model).  EQUISOLID r = 2 f sin(theta/2): the
 * Metashape calibrations of the dual-fisheye tool and v360 `input=equisolid`.  EQUIDISTANT
 * r = f theta: v360 `input=fisheye`, i.e. --fisheye-projection equidistant of
 * gs360_Video2Frames.py:466-487. */
enum { R360_LENS_EQUISOLID = 0, R360_LENS_EQUIDISTANT = 1 };

/* Metashape equisolid-fisheye calibration, the fields of SensorCalibration
 * (gs360_DualFisheyeDistortionCalibration.py:67-85) that the projection uses.  An ideal lens of
 * v360's `ih_fov` / `iv_fov` (gs360_Video2Frames.py:483-487) is the same record with zero
 * distortion terms: f + b1 = (W/2) / r(ih_fov/2), f = (H/2) / r(iv_fov/2), r the radial law. */
typedef struct r360_fisheye_calib {
    double width, height;    /* sensor resolution the calibration refers to                 */
    double f, cx, cy;
    double k1, k2, k3, k4;
    double p1, p2;
    double b1, b2;
    double lens_fov_deg;     /* usable lens FOV (DF --lens-fov-deg, default 190)            */
    int32_t model;           /* R360_LENS_EQUISOLID (0) or R360_LENS_EQUIDISTANT            */
    int32_t reserved;
} r360_fisheye_calib;

/* One output of the fisheye -> undistorted-fisheye remap (DF --save-fisheye-output): which lens
 * image it reads and the output zoom (RemapCache.undistort_zoom, DF:1136-1145; the auto value
 * is estimate_auto_undistort_zoom, DF:1054-1117 -- computed by the host layer). */
typedef struct r360_undistort {
    double  zoom;            /* > 0 (clamped to >= 1e-6 like DF:1145)                       */
    int32_t src_slot;        /* image of a source group / index into the calibration array  */
    int32_t reserved;
} r360_undistort;

typedef struct r360_options {
    int32_t interp;          /* R360_NEAREST / LINEAR / CUBIC / LANCZOS4                    */
    int32_t convention;      /* ERP only: R360_CONV_*                                       */
    int32_t path;            /* R360_PATH_*                                                 */
    int32_t fill_invalid;    /* fisheye/undistort: 1 = write border_value where the ray is outside
                                the lens model or the sensor (DF:2009-2014)                 */
    double  border_value;    /* fisheye only: per-tap constant border (cv2 BORDER_CONSTANT) */
    int32_t out_dtype;       /* -1 = same as source; R360_F16 with a U16 source writes
                                value/65535 as half                                         */
    int32_t reserved;
} r360_options;

/* ---- entry points --------------------------------------------------------------------------- */

int         r360_abi_version(void);
const char* r360_error_string(int code);
const char* r360_last_cuda_error(void);   /* thread-local text of the last CUDA failure       */
void        r360_default_options(r360_options* opt);

/* Number of SMs of the current device and whether the library's kernels can run on it. */
int r360_device_info(int* sm_count, int* cc_major, int* cc_minor);

/*
 * Panorama -> perspective views.  Replaces one `ffmpeg -vf v360=input=equirect:
 * output=rectilinear...` process per (source, view) (gs360_360PerspCut.py:286-349, :569-590).
 *
 *   src   `count` ERP frames;
 *   dst   src->count * n_views images, frame-major: image (f * n_views + v) is view v of
 *         frame f; dst->count must equal that product;
 *   x taps wrap at the seam, y taps clamp at the poles.
 */
int r360_remap_erp(const r360_images* src, const r360_images* dst,
                   const r360_view* views, int32_t n_views,
                   const r360_options* opt, void* stream);

/*
 * Dual (or n-) fisheye -> perspective views.  Replaces map build + cv2.remap + mask fill
 * (gs360_DualFisheyeDistortionCalibration.py:1759-1823, :2001-2014; masks :2031-2043 with
 * interp = NEAREST and border_value = 0).
 *
 *   src      groups of `n_lenses` images: image (g * n_lenses + slot); src->count must be a
 *            multiple of n_lenses;
 *   calib    one calibration per slot;
 *   views    yaw is RELATIVE to the lens of views[v].src_slot (DF:1883);
 *   dst      (src->count / n_lenses) * n_views images, group-major.
 */
int r360_remap_fisheye(const r360_images* src, const r360_images* dst,
                       const r360_fisheye_calib* calib, int32_t n_lenses,
                       const r360_view* views, int32_t n_views,
                       const r360_options* opt, void* stream);

/*
 * Fisheye -> undistorted fisheye ("remove the Brown terms, keep the equisolid projection").
 * Replaces build_remap_cache + cv2.remap + mask fill of undistort_prepared_image
 * (gs360_DualFisheyeDistortionCalibration.py:1008-1051, :1120-1170, :1173-1217).  Output pixel
 * (i, j) -- integer coordinates, no half-pixel offset, as in the reference -- samples the source at
 *   y0 = (j - cy0) / f, x0 = (i - cx0 - y0 * b2) / (f + b1), (x, y) = (x0, y0) / zoom,
 *   Brown(x, y) -> (cx0 + xd * (f + b1) + yd * b2, cy0 + yd * f),
 * valid where 2 * asin(min(r / 2, 1)) <= lens_fov / 2 and the source lies on the sensor.
 *
 *   src      groups of `n_lenses` images, as r360_remap_fisheye;
 *   items    n_items outputs per group (typically one per lens);
 *   dst      (src->count / n_lenses) * n_items images, group-major; any size (the reference
 *            writes sensor-sized images);
 *   R360_E_INVALID_ARG if |f| or |f + b1| < 1e-12 (the reference raises, DF:1018-1020).
 */
int r360_remap_undistort(const r360_images* src, const r360_images* dst,
                         const r360_fisheye_calib* calib, int32_t n_lenses,
                         const r360_undistort* items, int32_t n_items,
                         const r360_options* opt, void* stream);

/*
 * Input colour pipeline of the dual-fisheye tool: .cube 3-D LUT (trilinear, RGB) followed by an
 * optional Rec.709 -> linear -> sRGB re-encoding, applied to every lens image before the remap
 * when --input-lut / --input-color-profile osmo360-dlogm is set.  Replaces
 * apply_input_color_pipeline (gs360_DualFisheyeDistortionCalibration.py:684-725: image_to_float01
 * :599-609, apply_cube_lut_trilinear :625-681, rec709_to_srgb :568-596, float01_to_image :612-622).
 * float32 arithmetic in the reference's operation order; src and dst may be the same images.
 *
 *   table_device   size^3 nodes of 4 floats (R, G, B, unused), node (r, g, b) at index
 *                  (b * size + g) * size + r -- the order of a .cube file (red fastest);
 *   output_space   R360_LUT_PASSTHROUGH: clip to [0, 1]; R360_LUT_SRGB: rec709_to_srgb;
 *   channel_order  R360_ORDER_BGR (what cv2.imread returns; the reference flips it, DF:700-701)
 *                  or R360_ORDER_RGB;  channels beyond the first three are copied through.
 *   U8, U16 and F32 images; src and dst layouts must agree.
 */
enum { R360_LUT_PASSTHROUGH = 0, R360_LUT_SRGB = 1 };
enum { R360_ORDER_BGR = 0, R360_ORDER_RGB = 1 };
typedef struct r360_lut3d {
    const float* table_device;
    int32_t size;
    int32_t reserved;
    float   domain_min[3];
    float   domain_max[3];
} r360_lut3d;

int r360_apply_lut(const r360_images* src, const r360_images* dst, const r360_lut3d* lut,
                   int32_t output_space, int32_t channel_order, void* stream);

/*
 * Video colour step of the cutter: `colorspace=iall=bt709:all=smpte170m[:trc=iec61966-2-1]`, the filter
 * gs360_360PerspCut.py:299-309 puts in front of v360 for video sources (gs360_Video2Frames.py:462-464 puts
 * it after).  On R'G'B' frames that filter amounts to: decode `in_trc`, multiply linear RGB by `matrix`
 * (a change of primaries: BT.709 -> SMPTE 170M for the call above), encode `out_trc`, clip to the code
 * range.  float32 arithmetic; ffmpeg's own implementation is 15-bit fixed point behind look-up tables and is
 * not in this image, so results are stated against the formula, not against ffmpeg.
 *
 *   TRC_BT709 is also the SMPTE 170M curve (same constants); TRC_SRGB is IEC 61966-2-1; TRC_LINEAR none.
 *   U8, U16 and F32 images, >= 3 channels (extra channels are copied); src and dst may be the same images.
 */
enum { R360_TRC_BT709 = 0, R360_TRC_SRGB = 1, R360_TRC_LINEAR = 2 };
typedef struct r360_color_convert {
    int32_t in_trc;
    int32_t out_trc;
    float   matrix[9];       /* row-major, linear RGB -> linear RGB */
    int32_t reserved;
} r360_color_convert;

int r360_convert_color(const r360_images* src, const r360_images* dst, const r360_color_convert* cc,
                       int32_t channel_order, void* stream);

/*
 * Test/debug: the source coordinates the kernels sample at, without sampling.  Writes, for
 * view v and output pixel (j, i), element [(v * out_h + j) * out_w + i] of each non-null
 * device array:
 *   map_x32 / map_y32   float   the value handed to the 1/32-px quantiser (what cv2.remap
 *                               would receive as its map);
 *   map_x64 / map_y64   double  the same coordinate before the float32 rounding;
 *   valid               uint8   fisheye only: inside lens FOV and sensor bounds.
 * `calib == NULL` selects the ERP projection of a src_w x src_h panorama.
 */
int r360_coords(int32_t src_w, int32_t src_h,
                const r360_fisheye_calib* calib, int32_t n_lenses,
                const r360_view* views, int32_t n_views,
                int32_t out_w, int32_t out_h, const r360_options* opt,
                float* map_x32, float* map_y32, double* map_x64, double* map_y64,
                uint8_t* valid, void* stream);

/* r360_coords for the undistort projection (maps of DF:1120-1170 before the float32 cast too). */
int r360_coords_undistort(const r360_fisheye_calib* calib, int32_t n_lenses,
                          const r360_undistort* items, int32_t n_items,
                          int32_t out_w, int32_t out_h,
                          float* map_x32, float* map_y32, double* map_x64, double* map_y64,
                          uint8_t* valid, void* stream);

/* ---- planned (tiled) execution -------------------------------------------------------------------
 *
 * The reference builds its remap tables once per view set and applies them to every frame
 * (gs360_DualFisheyeDistortionCalibration.py:1857-1907 build, :1996-2014 apply; ffmpeg's v360
 * does the same inside each process).  A plan is this library's equivalent: per 32x32 output
 * tile, polynomial coordinates fitted in float64 plus the source patch to stage -- 384 bytes per
 * tile (368-byte record + fallback-list and walk-order entries) in a caller-provided device workspace, built on the device by r360_plan_create_*.
 * Panorama tiles next to a pole, whose map no polynomial follows, carry an explicit 16 KB per-pixel map instead (a pool
 * for one tile in sixteen follows the records: 1 KB per tile of workspace on average).
 * r360_remap_planned then runs the tiled fast kernels (bulk-async staging to shared memory),
 * falling back to the direct path tile by tile where the plan says so (tiles that contain a pole, patches larger
 * than shared memory).  Results are the same as
 * r360_remap_erp / r360_remap_fisheye up to 1/32-px bin flips from ~1e-5 px coordinate noise.
 *
 * A plan is tied to the source/destination LAYOUT (size, channels, dtype, pitches, 16-byte
 * alignment of the base pointers), the views, and the options; `data` and `count` of the two
 * layout descriptors are ignored at creation.  Creating a plan synchronises `stream` once.
 */
typedef struct r360_plan r360_plan;    /* opaque host-side handle, owned by the library */

size_t r360_plan_workspace_bytes(int32_t n_views, int32_t out_w, int32_t out_h);

int r360_plan_create_erp(const r360_images* src_layout, const r360_images* dst_layout,
                         const r360_view* views, int32_t n_views, const r360_options* opt,
                         void* workspace_device, size_t workspace_bytes, void* stream,
                         r360_plan** plan_out);

int r360_plan_create_fisheye(const r360_images* src_layout, const r360_images* dst_layout,
                             const r360_fisheye_calib* calib, int32_t n_lenses,
                             const r360_view* views, int32_t n_views, const r360_options* opt,
                             void* workspace_device, size_t workspace_bytes, void* stream,
                             r360_plan** plan_out);

int r360_plan_create_undistort(const r360_images* src_layout, const r360_images* dst_layout,
                               const r360_fisheye_calib* calib, int32_t n_lenses,
                               const r360_undistort* items, int32_t n_items, const r360_options* opt,
                               void* workspace_device, size_t workspace_bytes, void* stream,
                               r360_plan** plan_out);

/* Tiles per view and how many (view, tile) pairs take the direct path. */
int r360_plan_info(const r360_plan* plan, int32_t* tiles_per_view, int32_t* n_fallback_tiles);

/* How many tiles carry an explicit per-pixel map (panorama tiles next to a pole, where no polynomial follows the
 * projection) and how many such maps the workspace has room for. */
int r360_plan_info_maps(const r360_plan* plan, int32_t* n_map_tiles, int32_t* map_pool_tiles);

/* Same contract as r360_remap_erp / r360_remap_fisheye for the layout the plan was made for;
 * R360_E_INVALID_ARG if src/dst do not match that layout. */
int r360_remap_planned(const r360_plan* plan, const r360_images* src, const r360_images* dst,
                       void* stream);

/* Test/debug twin of r360_coords for the planned path (all five arrays as in r360_coords;
 * the four map arrays are required). */
int r360_plan_coords(const r360_plan* plan, float* map_x32, float* map_y32, double* map_x64,
                     double* map_y64, uint8_t* valid, void* stream);

void r360_plan_destroy(r360_plan* plan);

/* Count of kernels this library has launched in the calling process (all threads). */
int64_t r360_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* REMAP360_H */
