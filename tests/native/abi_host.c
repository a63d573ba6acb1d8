/* A plain C caller of include/remap360.h -- no Python, no torch: device memory comes from cudaMalloc.
 *
 *   abi_host check                          header compiles as C, symbols resolve, argument validation (no GPU)
 *   abi_host erp  in.raw W H out.raw size   one ERP frame (u8 x 3) -> 3 views (yaw 0 / 90 / 180+pitch 30), bicubic, direct path
 *   abi_host plan in.raw W H out.raw size   the same through a tile plan (r360_plan_create_erp / r360_remap_planned)
 *
 * Built and driven by tests/test_abi_native.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "remap360.h"

static int fail(const char* what, int code) {
    fprintf(stderr, "%s: %d (%s) %s\n", what, code, r360_error_string(code), r360_last_cuda_error());
    return 1;
}

static int check(void) {
    r360_options opt;
    r360_images img;
    r360_view v;
    if (r360_abi_version() != R360_ABI_VERSION) return fail("abi version", r360_abi_version());
    r360_default_options(&opt);
    if (opt.interp != R360_CUBIC || opt.fill_invalid != 1 || opt.out_dtype != -1) return fail("defaults", -1);
    memset(&img, 0, sizeof img);
    memset(&v, 0, sizeof v);
    v.hfov_deg = v.vfov_deg = 90.0;
    img.width = img.height = 8; img.channels = 3; img.dtype = R360_U8; img.pitch_bytes = 24;
    img.image_stride_bytes = 192; img.count = 1;
    if (r360_remap_erp(&img, &img, &v, 1, &opt, 0) != R360_E_INVALID_ARG) return fail("null data accepted", 0);
    if (r360_plan_workspace_bytes(12, 1600, 1600) < 12u * 2500u * 368u) return fail("workspace size", 0);
    if (sizeof(r360_fisheye_calib) != 14 * sizeof(double) + 8) return fail("calib layout", (int)sizeof(r360_fisheye_calib));
    puts("abi_host check ok");
    return 0;
}

static int run(int planned, const char* in_path, int W, int H, const char* out_path, int size) {
    const size_t in_bytes = (size_t)W * H * 3, view_bytes = (size_t)size * size * 3;
    const int n_views = 3;
    unsigned char *host_in = malloc(in_bytes), *host_out = malloc(view_bytes * n_views);
    void *dev_in = NULL, *dev_out = NULL, *ws = NULL;
    r360_images src, dst;
    r360_view views[3];
    r360_options opt;
    FILE* f = fopen(in_path, "rb");
    int rc;
    if (!f || fread(host_in, 1, in_bytes, f) != in_bytes) return fail("read input", -1);
    fclose(f);
    if (cudaMalloc(&dev_in, in_bytes) || cudaMalloc(&dev_out, view_bytes * n_views)) return fail("cudaMalloc", -3);
    cudaMemcpy(dev_in, host_in, in_bytes, cudaMemcpyHostToDevice);
    memset(&src, 0, sizeof src); memset(&dst, 0, sizeof dst); memset(views, 0, sizeof views);
    src.data = dev_in; src.width = W; src.height = H; src.channels = 3; src.dtype = R360_U8;
    src.pitch_bytes = (int64_t)W * 3; src.image_stride_bytes = (int64_t)in_bytes; src.count = 1;
    dst.data = dev_out; dst.width = size; dst.height = size; dst.channels = 3; dst.dtype = R360_U8;
    dst.pitch_bytes = (int64_t)size * 3; dst.image_stride_bytes = (int64_t)view_bytes; dst.count = n_views;
    views[0].yaw_deg = 0.0;   views[1].yaw_deg = 90.0;  views[2].yaw_deg = 180.0; views[2].pitch_deg = 30.0;
    for (int k = 0; k < n_views; ++k) views[k].hfov_deg = views[k].vfov_deg = 104.2500326978036;
    r360_default_options(&opt);
    if (!planned) {
        rc = r360_remap_erp(&src, &dst, views, n_views, &opt, 0);
        if (rc != R360_OK) return fail("r360_remap_erp", rc);
    } else {
        r360_plan* plan = NULL;
        int32_t tiles = 0, fallback = 0;
        const size_t ws_bytes = r360_plan_workspace_bytes(n_views, size, size);
        if (cudaMalloc(&ws, ws_bytes)) return fail("cudaMalloc workspace", -3);
        opt.path = R360_PATH_TILED;
        rc = r360_plan_create_erp(&src, &dst, views, n_views, &opt, ws, ws_bytes, 0, &plan);
        if (rc != R360_OK) return fail("r360_plan_create_erp", rc);
        r360_plan_info(plan, &tiles, &fallback);
        rc = r360_remap_planned(plan, &src, &dst, 0);
        if (rc != R360_OK) return fail("r360_remap_planned", rc);
        cudaDeviceSynchronize();
        r360_plan_destroy(plan);
        printf("plan: %d tiles per view, %d fallback tiles\n", (int)tiles, (int)fallback);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return fail("kernel", -3);
    cudaMemcpy(host_out, dev_out, view_bytes * n_views, cudaMemcpyDeviceToHost);
    f = fopen(out_path, "wb");
    if (!f || fwrite(host_out, 1, view_bytes * n_views, f) != view_bytes * n_views) return fail("write output", -1);
    fclose(f);
    printf("launches: %lld\n", (long long)r360_launch_count());
    cudaFree(dev_in); cudaFree(dev_out); cudaFree(ws); free(host_in); free(host_out);
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 2 && !strcmp(argv[1], "check")) return check();
    if (argc == 7 && (!strcmp(argv[1], "erp") || !strcmp(argv[1], "plan")))
        return run(!strcmp(argv[1], "plan"), argv[2], atoi(argv[3]), atoi(argv[4]), argv[5], atoi(argv[6]));
    fprintf(stderr, "usage: abi_host check | erp|plan in.raw W H out.raw size\n");
    return 2;
}
