/*
 * sdr_c.h -- C ABI of the B200-native stereo depth engine (libsdr_b200.so).
 *
 * This is the drop-in boundary for the hot path of Amar-Aliaga/Stereo_Depth_Ruler:
 * rectify -> StereoSGBM-compatible disparity (CV_16S, 4 fractional bits) -> reproject
 * with Q -> coloured VoxelGrid cloud. The reference has no FFI layer of its own; its
 * "operator API" is three C++ classes plus the OpenCV / PCL free functions they call.
 * Every entry point below names the reference interface it replaces (file:line in
 * the reference tree). include/sdr/stereo_vision.hpp is the C++ facade with the
 * reference's class and method names over this ABI; INTEGRATION.md shows the binding
 * a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every call returns an sdr_status; the
 * library never throws across the boundary; there is NO CPU fallback -- without a
 * CUDA device sdr_create fails with SDR_ECUDA and nothing else can be called.
 * A handle is thread-compatible, not thread-safe: one ctx <-> one device <-> one stream.
 */
#ifndef SDR_C_H
#define SDR_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SDR_API __attribute__((visibility("default")))
#else
#define SDR_API
#endif

typedef enum sdr_status {
    SDR_OK = 0,
    SDR_EINVAL = 1,          /* bad argument (cv::Exception on bad args in the reference's OpenCV calls) */
    SDR_ESIZE = 2,           /* image size / type mismatch (stereo_rectifier.cpp:27-29 only warns) */
    SDR_ECUDA = 3,           /* CUDA runtime failure or no device */
    SDR_ENOMEM = 4,
    SDR_EIO = 5,             /* file could not be read / parsed (stereo_configuration.cpp:6-8,16-19,37-41) */
    SDR_EVOXEL_OVERFLOW = 6  /* non-fatal: PCL's int32 voxel-index guard tripped, cloud passed through */
} sdr_status;

/* SDR_MEM_DEVICE images are used in place by the kernels of the handle's device. They may live on another GPU of the box when
 * that GPU is a peer (NVLink / PCIe P2P): peer access is enabled on first use; otherwise, or when the pointer is not device
 * memory at all, the call returns SDR_EINVAL (never a fault inside a kernel). */
typedef enum sdr_mem { SDR_MEM_HOST = 0, SDR_MEM_PINNED = 1, SDR_MEM_DEVICE = 2 } sdr_mem;

/* Layout-compatible with the fields of cv::Mat the reference relies on
 * {data, rows, cols, step, type}; rows may be non-contiguous (the SBS ROI views of
 * stereo_displayer.cpp:155-156). */
typedef struct sdr_image {
    void* data;
    int rows, cols, channels, elem_size; /* elem_size: bytes per channel element (1 = u8, 2 = s16, 4 = f32) */
    size_t step_bytes;
    int mem; /* sdr_mem */
} sdr_image;

/* StereoConfiguration (stereo_vision/include/stereo_configuration.hpp:6-16), row-major fp64. */
typedef struct sdr_calibration {
    int image_width, image_height;
    double K_left[9], dist_left[5], K_right[9], dist_right[5];
    double R[9], T[3], E[9], F[9];
    double R1[9], R2[9], P1[12], P2[12], Q[16];
    uint32_t present; /* bit i set when field i (order above, K_left = bit 0) was found in the file */
} sdr_calibration;

/* Same order and meaning as cv::StereoSGBM::create(...) at stereo_disparity.cpp:5-6 /
 * pcd_write.cpp:102-108. */
typedef enum sdr_sgbm_mode { SDR_MODE_SGBM = 0, SDR_MODE_HH = 1, SDR_MODE_SGBM_3WAY = 2 } sdr_sgbm_mode;
typedef struct sdr_sgbm_params {
    int minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff;
    int preFilterCap, uniquenessRatio, speckleWindowSize, speckleRange, mode;
} sdr_sgbm_params;

typedef struct sdr_ctx sdr_ctx;
typedef struct sdr_rectifier sdr_rectifier;
typedef struct sdr_sgbm sdr_sgbm;
typedef struct sdr_cloud sdr_cloud;
typedef struct sdr_pipeline sdr_pipeline;

/* ---- context ----------------------------------------------------------------- */
SDR_API const char* sdr_version(void);
SDR_API int sdr_device_count(void);
SDR_API int sdr_create(int device, sdr_ctx** ctx);
SDR_API void sdr_destroy(sdr_ctx* ctx);
/* Run on a caller-owned CUDA stream (a cudaStream_t passed as void*); NULL restores the ctx's own stream. */
SDR_API int sdr_set_stream(sdr_ctx* ctx, void* cuda_stream);
SDR_API void* sdr_get_stream(sdr_ctx* ctx);
SDR_API int sdr_synchronize(sdr_ctx* ctx);
SDR_API const char* sdr_last_error(sdr_ctx* ctx);
/* Number of this library's kernels launched through ctx since creation. */
SDR_API uint64_t sdr_kernel_launches(sdr_ctx* ctx);

/* ---- calibration: StereoConfiguration::loadFromFile / saveToFile / isValid
 *      (stereo_configuration.cpp:4-46, :49-74, :77-80). OpenCV-dialect YAML 1.0. ---- */
SDR_API int sdr_load_calibration_yaml(const char* path, sdr_calibration* out);
SDR_API int sdr_save_calibration_yaml(const char* path, const sdr_calibration* in);
SDR_API int sdr_calibration_is_valid(const sdr_calibration* c);

/* ---- rectification: StereoRectifier ctor + rectify (stereo_rectifier.cpp:6-14, :17-41) ---- */
SDR_API int sdr_rectifier_create(sdr_ctx* ctx, const sdr_calibration* calib, sdr_rectifier** out);
SDR_API void sdr_rectifier_destroy(sdr_rectifier* r);
/* eye 0 = left, 1 = right. map1: int16[H][W][2], map2: uint16[H][W] (host), exactly what
 * cv::initUndistortRectifyMap(..., CV_16SC2, map1, map2) returns. */
SDR_API int sdr_rectifier_get_maps(sdr_rectifier* r, int eye, int16_t* map1, uint16_t* map2);
/* left_src/right_src: u8 BGR (3 ch) or gray (1 ch); *_dst same type; *_gray optional (NULL) u8 1-ch
 * outputs = cvtColor(dst, BGR2GRAY) fused into the same kernel (stereo_disparity.cpp:19-20). */
SDR_API int sdr_rectify(sdr_rectifier* r, const sdr_image* left_src, const sdr_image* right_src,
                        sdr_image* left_dst, sdr_image* right_dst, sdr_image* left_gray, sdr_image* right_gray);
/* cv::remap(src, dst, map1, map2, INTER_LINEAR) with caller-supplied CV_16SC2 maps (host pointers). */
SDR_API int sdr_remap_with_maps(sdr_ctx* ctx, const sdr_image* src, const int16_t* map1, const uint16_t* map2,
                                sdr_image* dst);
/* cv::cvtColor(BGR2GRAY) (stereo_disparity.cpp:19-20, pcd_write.cpp:88-89). */
SDR_API int sdr_bgr2gray(sdr_ctx* ctx, const sdr_image* bgr, sdr_image* gray);

/* ---- disparity: cv::StereoSGBM::create + compute (stereo_disparity.cpp:5-6,27; pcd_write.cpp:102-111) ---- */
SDR_API int sdr_sgbm_create(sdr_ctx* ctx, const sdr_sgbm_params* params, sdr_sgbm** out);
SDR_API void sdr_sgbm_destroy(sdr_sgbm* m);
/* left/right: u8 1-ch, same size. disp: int16 1-ch (CV_16S, disparity * 16; invalid = (minD-1)*16).
 * Includes the medianBlur(3) and filterSpeckles that StereoSGBM::compute applies internally. */
SDR_API int sdr_sgbm_compute(sdr_sgbm* m, const sdr_image* left, const sdr_image* right, sdr_image* disp);
/* n independent pairs of one size in one submission (frame-parallel batch; frames share no state). */
SDR_API int sdr_sgbm_compute_batch(sdr_sgbm* m, int n, const sdr_image* lefts, const sdr_image* rights,
                                   sdr_image* disps);
/* Row-strip split of one oversize frame (BASELINE config 5). MODE_SGBM_3WAY is four independent stripes by construction
 * (OpenCV's computeDisparity3WaySGBM: stripe i reads source rows [src_start, row_end) and owns output rows
 * [dst_start, row_end)); other modes are a single stripe. Each GPU matches its stripes from the input rows it needs,
 * the raw strips are gathered, and the median / speckle post-filters run once over the assembled image.
 * sdr_sgbm_stripe_rows returns the number of stripes for this size (and one stripe's rows); it needs no device. */
SDR_API int sdr_sgbm_stripe_rows(const sdr_sgbm_params* params, int width, int height, int stripe, int* src_start, int* dst_start,
                                 int* row_end);
/* Raw matcher output (before median / speckle) of stripes [first, first + n): only their output rows of `raw`
 * (int16, full image size) are written. left/right are the full images. */
SDR_API int sdr_sgbm_match_stripes(sdr_sgbm* m, const sdr_image* left, const sdr_image* right, int first_stripe, int n_stripes,
                                   sdr_image* raw);
/* medianBlur(3) + filterSpeckles with the handle's parameters on an assembled raw map. */
SDR_API int sdr_sgbm_postfilter(sdr_sgbm* m, const sdr_image* raw, sdr_image* disp);

/* The same split, native and in one process, over several GPUs of one box (BASELINE config 5). devices[0] assembles the
 * map and runs the median / speckle filters; every device matches a contiguous block of the stripes from just the input
 * rows it needs (host frames: each device uploads its own overlapping row range; frames resident on devices[0]: the
 * rows travel by cudaMemcpyPeerAsync over NVLink) and pushes its raw strip to devices[0] by a peer copy. No collective.
 * Devices beyond the number of stripes (4 for MODE_SGBM_3WAY, 1 otherwise) idle; a device may be listed more than once.
 * left / right: u8 1-ch in host / pinned memory, or dense in the memory of devices[0]; disp: int16, host or devices[0].
 * With a device-resident disp the call is asynchronous: sdr_sgbm_split_synchronize waits for it. */
typedef struct sdr_sgbm_split sdr_sgbm_split;
typedef struct sdr_split_stats {
    int n_parts, n_stripes;
    int device[16], stripes[16], input_rows[16], peer_access[16];
    int64_t input_bytes[16];  /* image bytes uploaded / peer-copied to part k (0: used in place) */
    int64_t gather_bytes[16]; /* raw-strip bytes peer-copied from part k to devices[0] */
    float ms;                 /* device time of the whole call, CUDA events on devices[0]'s stream */
} sdr_split_stats;
SDR_API int sdr_sgbm_split_create(const int* devices, int n_devices, const sdr_sgbm_params* params, sdr_sgbm_split** out);
SDR_API void sdr_sgbm_split_destroy(sdr_sgbm_split* s);
SDR_API int sdr_sgbm_split_compute(sdr_sgbm_split* s, const sdr_image* left, const sdr_image* right, sdr_image* disp);
SDR_API int sdr_sgbm_split_synchronize(sdr_sgbm_split* s);
SDR_API int sdr_sgbm_split_last_stats(sdr_sgbm_split* s, sdr_split_stats* out);
SDR_API const char* sdr_sgbm_split_last_error(sdr_sgbm_split* s);

/* Kernel-selection knobs of one matcher handle. No option changes a result (every kernel variant is bit-exact and tested);
 * the library never reads the process environment. Options: "shared_device" (1: other matchers run on this device at the
 * same time -- the throughput runners set it -- so the two horizontal passes stay separate kernels that leave room for
 * each other; 0, the default and the reference's usage: one matcher, both passes in one launch), "path_hh" (-1 automatic,
 * 0 / 1 force), "fused_sweep", "cost_fusion" (0 disables the cooperative sweep / the fused cost + sweep kernel),
 * "cf_cols" (32 / 64), "cf_prod", "cf_fpl", "v3_stages", "h_legacy", "h_k", "h_stages", "h_warps", "hb_threads",
 * "hb_ring", "hh_threads" (tile shapes, 0 = default). Unknown names return SDR_EINVAL. */
SDR_API int sdr_sgbm_set_option(sdr_sgbm* m, const char* name, int value);

/* Per-kernel device times (ms) of the last compute, measured with CUDA events on the ctx stream.
 * Enable first; names[i] point to static strings. Returns the number of stages written (<= cap). */
SDR_API int sdr_sgbm_set_profiling(sdr_sgbm* m, int enabled);
SDR_API int sdr_sgbm_stage_times(sdr_sgbm* m, int cap, const char** names, float* ms);
/* Test hook: copy an internal buffer of the last single-frame compute to host.
 * what: 0 = C volume int16 [rowsC][W1][D] (without the +P2 bias), 1 = S volume int16 [H][W1][D],
 * 2 = raw matcher output int16 [H][W] (before median / speckle). */
SDR_API int sdr_sgbm_debug_read(sdr_sgbm* m, int what, void* host_dst, size_t bytes);

/* cv::medianBlur(disp, 3) on CV_16S and cv::filterSpeckles, as StereoSGBM::compute chains them. */
SDR_API int sdr_median3_16s(sdr_ctx* ctx, const sdr_image* src, sdr_image* dst);
SDR_API int sdr_filter_speckles(sdr_ctx* ctx, sdr_image* img, int new_val, int max_size, int max_diff);

/* ---- depth: convertTo(CV_32F, 1/16) + cv::reprojectImageTo3D
 *      (stereo_disparity.cpp:34,76-80; pcd_write.cpp:112,116) ---- */
/* Mat::convertTo(CV_32F, 1.0/16.0) of the CV_16S map (stereo_disparity.cpp:34, pcd_write.cpp:112): exact in fp32. */
SDR_API int sdr_disp_to_float(sdr_ctx* ctx, const sdr_image* disp16, sdr_image* disp32f);
/* disp: int16 (CV_16S, divided by 16 internally, i.e. convertTo + reproject fused) or f32 (used as is).
 * Q: 16 doubles row-major. points: f32 3-ch. */
SDR_API int sdr_reproject_to_3d(sdr_ctx* ctx, const sdr_image* disp, const double* Q, int handle_missing,
                                sdr_image* points);

/* ---- cloud: convertCVMatToPCL + pcl::VoxelGrid<PointXYZRGB> (pcd_write.cpp:17-51, :123-130) ---- */
typedef struct sdr_voxel_options {
    float leaf[3];
    int drop_invalid;  /* 0: literal reference semantics (every finite point, incl. the Z=10000 fill of
                          handleMissingValues); 1: drop points whose disparity is invalid (<= (minD-1)) first */
    float z_max;       /* > 0: also drop points with z > z_max (or z <= 0); 0 = off */
    int min_disparity; /* used by drop_invalid */
} sdr_voxel_options;

typedef struct sdr_cloud_view {
    int64_t n;          /* number of output points */
    const float* xyz;   /* DEVICE pointer, f32[n][3] */
    const uint8_t* rgb; /* DEVICE pointer, u8[n][3] (r,g,b) */
    const uint32_t* idx;/* DEVICE pointer, voxel linear index per output point (ascending); NULL on passthrough */
    const uint32_t* cnt;/* DEVICE pointer, points per voxel; NULL on passthrough */
    int32_t min_b[3], div_b[3];
    int passthrough;    /* 1 when the int32 guard tripped (status SDR_EVOXEL_OVERFLOW): cloud = finite input points */
} sdr_cloud_view;

SDR_API int sdr_cloud_create(sdr_ctx* ctx, sdr_cloud** out);
SDR_API void sdr_cloud_destroy(sdr_cloud* c);
/* points: f32 3-ch [H][W]; bgr: u8 3-ch (or NULL: colours 0); disp16: optional int16 disparity for
 * drop_invalid (may be NULL when drop_invalid == 0). */
SDR_API int sdr_voxelgrid(sdr_cloud* c, const sdr_image* points, const sdr_image* bgr, const sdr_image* disp16,
                          const sdr_voxel_options* opt, sdr_cloud_view* view);
/* Fused reprojectImageTo3D + voxelisation straight from the CV_16S disparity (no W*H*3 float image in HBM). */
SDR_API int sdr_depth_cloud(sdr_cloud* c, const sdr_image* disp16, const sdr_image* bgr, const double* Q,
                            int handle_missing, const sdr_voxel_options* opt, sdr_cloud_view* view);
/* Copy the last cloud to host buffers (any may be NULL). */
SDR_API int sdr_cloud_download(sdr_cloud* c, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt);

/* ---- consumers of the depth image: tiny reductions done on the device so the 11 MB point image
 *      does not have to travel back ---- */
/* StereoDisplayer::onMouseMeasure (stereo_displayer.cpp:41-47): distance between the reprojected pixels (x1,y1), (x2,y2)
 * = cv::norm(xyz1 - xyz2): float differences, squares summed in double. points: f32 3-ch. Same units as the cloud (mm;
 * the reference prints dist/10 as cm). Out-of-range clicks return SDR_EINVAL (the reference ignores them, :33). */
SDR_API int sdr_measure_distance(sdr_ctx* ctx, const sdr_image* points, int x1, int y1, int x2, int y2, double* dist);
/* StereoDisplayer::depth_coverage (stereo_displayer.cpp:105-118): pixels with column >= first_col (the reference hard-codes
 * 80), 0 <= Z <= z_max (12000) and Z not NaN; coverage_percent = count / (rows*cols) * 100. */
SDR_API int sdr_depth_coverage(sdr_ctx* ctx, const sdr_image* points, int first_col, float z_min, float z_max, int64_t* count,
                               double* coverage_percent);
/* pcl::io::savePCDFileBinary(path, cloud) for the last cloud of `c` (pcd_write.cpp:134-142): PCD v0.7 header
 * (FIELDS x y z rgb / SIZE 4 4 4 4 / TYPE F F F U / COUNT 1 1 1 1 / WIDTH n / HEIGHT 1 / VIEWPOINT 0 0 0 1 0 0 0 /
 * POINTS n / DATA binary) + packed 16-byte records, rgb = 0xFF<<24 | r<<16 | g<<8 | b. The header is restated from PCL's
 * pcd_io as recalled (no PCL and no sample file here): PARITY UNPINNED. Returns SDR_EINVAL for an empty cloud
 * (the reference prints "No points to save" and returns, :143-146). */
SDR_API int sdr_cloud_save_pcd_binary(sdr_cloud* c, int frame, const char* path);

/* ---- the reference's live path around the matcher (StereoDisparity, stereo_disparity.cpp) ---- */
/* cv::resize(src, dst, Size(), 0.5, 0.5, INTER_AREA) (stereo_disparity.cpp:23-24): u8, 1 or 3 channels, even width and
 * height (OpenCV's exact-factor-2 path, (a+b+c+d+2)>>2); odd sizes return SDR_EINVAL. */
SDR_API int sdr_resize_half_area(sdr_ctx* ctx, const sdr_image* src, sdr_image* dst);
/* cv::ximgproc::createRightMatcher(matcher) (stereo_disparity.cpp:10): parameters of the matcher that, run on (right, left),
 * gives the right-view disparity: minDisparity = -(minD + D) + 1, uniquenessRatio 0, disp12MaxDiff 1000000, no speckle
 * filter, same P1 / P2 / preFilterCap / mode. The cv::ximgproc::DisparityWLSFilter that consumes it (:11-13, :31) is NOT
 * provided: ximgproc is absent from every oracle available here, so its parity could not be pinned. */
SDR_API int sdr_sgbm_right_matcher_params(const sdr_sgbm_params* left, sdr_sgbm_params* right);
/* Visualisation helpers. The handle owns what the reference keeps between frames: prev_vis, prev_depth_vis
 * (stereo_disparity.hpp:11) and the smoothed depth range (function-static in show_depthMap, stereo_disparity.cpp:93). */
typedef struct sdr_vis sdr_vis;
SDR_API int sdr_vis_create(sdr_ctx* ctx, sdr_vis** out);
SDR_API void sdr_vis_destroy(sdr_vis* v);
SDR_API int sdr_vis_reset(sdr_vis* v);
/* StereoDisparity::show_disparityMap (stereo_disparity.cpp:42-73): disparity f32 1-ch -> u8 1-ch. */
SDR_API int sdr_vis_disparity(sdr_vis* v, const sdr_image* disparity32f, int num_disparities, sdr_image* out_u8);
/* StereoDisparity::show_depthMap (stereo_disparity.cpp:83-124): f32 3-ch point image (Z = channel 2) or 1-ch depth ->
 * u8 BGR (COLORMAP_TURBO over the smoothed valid range, blended with the previous frame). */
SDR_API int sdr_vis_depth(sdr_vis* v, const sdr_image* depth32f, sdr_image* out_bgr);
SDR_API int sdr_vis_depth_range(sdr_vis* v, double* zmin_smooth, double* zmax_smooth);

/* ---- frame ingest: what sits between the video source and the BGR side-by-side frame the reference's
 *      loops start from (cv::VideoCapture >> frame, stereo_displayer.cpp:132-156, pcd_write.cpp:59-84). A hardware decoder
 *      (NVDEC) delivers NV12 surfaces and the ZED 2 streams YUYV 4:2:2 over UVC; the frame crosses PCIe in that format (12 / 16
 *      instead of 24 bits per pixel) and is converted on the device: cv::cvtColor(COLOR_YUV2BGR_NV12 / COLOR_YUV2BGR_YUY2),
 *      bit-exact. The bitstream decode itself is not part of this library. ---- */
typedef enum sdr_pixel_format {
    SDR_PIX_BGR = 0,  /* u8 3-ch, what cv::VideoCapture returns */
    SDR_PIX_NV12 = 1, /* u8 1-ch, rows * 3 / 2 rows: Y plane, then interleaved U,V at half resolution (even width and height) */
    SDR_PIX_YUYV = 2  /* u8 2-ch: Y0 U Y1 V (even width) */
} sdr_pixel_format;
/* src: the picture in `format` (NV12: rows = bgr->rows * 3 / 2, 1 channel; YUYV: 2 channels); bgr: u8 3-ch output. */
SDR_API int sdr_ingest_to_bgr(sdr_ctx* ctx, int format, const sdr_image* src, sdr_image* bgr);

/* ---- whole path per frame (stereo_displayer.cpp:155-162 + pcd_write.cpp:82-130) ---- */
typedef struct sdr_pipeline_options {
    sdr_sgbm_params sgbm;
    sdr_voxel_options voxel;
    int handle_missing; /* reprojectImageTo3D(..., handleMissingValues) */
    int batch;          /* frames per submission (>= 1) */
    int want_points;    /* also materialise the W*H*3 float image (computeDepth) */
} sdr_pipeline_options;

SDR_API int sdr_pipeline_create(sdr_ctx* ctx, const sdr_calibration* calib, const sdr_pipeline_options* opt,
                                sdr_pipeline** out);
SDR_API void sdr_pipeline_destroy(sdr_pipeline* p);
/* sbs: n side-by-side BGR frames (u8 3-ch, cols = 2*image_width). Outputs (each optional / NULL):
 * disp[n] int16 1-ch, points[n] f32 3-ch, left_rect[n] u8 3-ch. views[n] receives the clouds
 * (valid until the next run). */
SDR_API int sdr_pipeline_run(sdr_pipeline* p, int n, const sdr_image* sbs, sdr_image* disp, sdr_image* points,
                             sdr_image* left_rect, sdr_cloud_view* views);
/* Format of the SBS frames sdr_pipeline_run is handed (default SDR_PIX_BGR). With SDR_PIX_NV12 / SDR_PIX_YUYV the frames are
 * copied to the device as they are and converted there (sdr_ingest_to_bgr) in front of the rectification. */
SDR_API int sdr_pipeline_set_input_format(sdr_pipeline* p, int format);
/* sdr_sgbm_set_option on the pipeline's matcher. */
SDR_API int sdr_pipeline_set_sgbm_option(sdr_pipeline* p, const char* name, int value);
/* Per-stage device times (ms) of the last run (CUDA events on the ctx stream); enable first. */
SDR_API int sdr_pipeline_set_profiling(sdr_pipeline* p, int enabled);
SDR_API int sdr_pipeline_stage_times(sdr_pipeline* p, int cap, const char** names, float* ms);
/* The same marks as absolute times: ms from `base_cuda_event` (a cudaEvent_t recorded earlier on any stream of the device) to
 * the start of the last run ("begin") and to the end of each stage; which = 0: the pipeline's stages, 1: the matcher's kernels.
 * Laid side by side for several handles this is the timeline of who ran when (tools/probes/timeline.py). */
SDR_API int sdr_pipeline_stage_timeline(sdr_pipeline* p, void* base_cuda_event, int which, int cap, const char** names, float* t_ms);
/* Copy cloud i of the last run to host buffers (any may be NULL); sized by views[i].n. */
SDR_API int sdr_pipeline_cloud_download(sdr_pipeline* p, int i, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt);
/* The clouds of frames 0..n-1 of the last run in one go (arrays of n host pointers, any array may be NULL): the copies are
 * queued back to back and waited for once. */
SDR_API int sdr_pipeline_clouds_download(sdr_pipeline* p, int n, float* const* xyz, uint8_t* const* rgb, uint32_t* const* idx,
                                         uint32_t* const* cnt);
/* The same clouds packed back to back, as they sit on the device: ONE copy per array (any may be NULL) and one wait.
 * offsets[n + 1] (optional) receives the first point of every frame's cloud; frame i owns points [offsets[i], offsets[i+1]).
 * Buffers must hold offsets[n] points (sum of views[i].n of the run); idx / cnt are undefined for passthrough frames. */
SDR_API int sdr_pipeline_clouds_download_packed(sdr_pipeline* p, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt,
                                                int64_t* offsets);

/* ---- frame-parallel stream over the GPUs of one box (BASELINE config 4; the video loop of stereo_displayer.cpp:145-198
 *      with frames sharded by index, no collective: frames share no state) ----
 * One host worker thread per (device, handle): each owns a context, a stream and a pipeline, pulls the next batch of
 * frame indices from a shared counter, asks `source` for the frames, runs the path and hands every result to `sink`.
 * Callbacks are called from worker threads (concurrently for different frames). `source` fills *sbs with a view of
 * frame `index` (u8 BGR, cols = 2 * image_width, host / pinned / device memory that stays valid until the matching
 * sink call) and returns SDR_OK, or non-zero to drop the frame. `sink` receives the CV_16S map in pinned host memory
 * (valid during the call) and the cloud view (device pointers, valid during the call). */
typedef int (*sdr_frame_source)(void* user, int64_t index, sdr_image* sbs);
typedef void (*sdr_frame_sink)(void* user, int64_t index, int device, const sdr_image* disp, const sdr_cloud_view* cloud);
/* The frame's filtered cloud in pinned HOST memory (valid during the call): n points, xyz f32[n][3], rgb u8[n][3]. */
typedef void (*sdr_cloud_sink)(void* user, int64_t index, int device, int64_t n, const float* xyz, const uint8_t* rgb);
typedef struct sdr_stream_stats {
    int64_t frames;        /* frames processed */
    double seconds;        /* wall time of the whole run */
    double frames_per_s;
    int64_t per_device[16];
    uint64_t kernel_launches; /* this library's kernels launched by the run */
    int64_t h2d_bytes, d2h_bytes; /* bytes copied host->device (frames) and device->host (maps, clouds) by the run */
} sdr_stream_stats;
SDR_API int sdr_stream_run(const int* devices, int n_devices, int handles_per_device, const sdr_calibration* calib,
                           const sdr_pipeline_options* opt, int64_t n_frames, sdr_frame_source source, sdr_frame_sink sink, void* user,
                           sdr_stream_stats* stats);
/* The persistent form: contexts, pipelines, device scratch and pinned result buffers are created once (create) and reused
 * by every process() call; one process() at a time. download_clouds != 0: every batch's clouds are also copied to pinned
 * host memory (packed, one copy per array) and handed to cloud_sink. */
typedef struct sdr_stream sdr_stream;
SDR_API int sdr_stream_create(const int* devices, int n_devices, int handles_per_device, const sdr_calibration* calib,
                              const sdr_pipeline_options* opt, int download_clouds, sdr_stream** out);
SDR_API void sdr_stream_destroy(sdr_stream* s);
SDR_API int sdr_stream_process(sdr_stream* s, int64_t n_frames, sdr_frame_source source, sdr_frame_sink sink, sdr_cloud_sink cloud_sink,
                               void* user, sdr_stream_stats* stats);
/* sdr_pipeline_set_input_format on every worker's pipeline (call between create and process; sources then return frames in
 * that format and sdr_stream_process_pool reads its pool as such). */
SDR_API int sdr_stream_set_input_format(sdr_stream* s, int format);
/* sdr_sgbm_set_option on every worker's matcher (call between create and process; results never depend on it). */
SDR_API int sdr_stream_set_sgbm_option(sdr_stream* s, const char* name, int value);
/* process() over a pool of SBS frames at a constant stride, frame i of the stream = pool frame (i mod n_pool), no callback
 * into the host language (BASELINE config 4: 64 unique pairs cycled to 4096); results land in the workers' pinned
 * buffers and are dropped. mem: sdr_mem of the pool. */
SDR_API int sdr_stream_process_pool(sdr_stream* s, int64_t n_frames, const void* pool_base, int64_t n_pool, size_t frame_stride,
                                    size_t step_bytes, int mem, sdr_stream_stats* stats);
/* Binds the calling thread (and the threads it creates later) to the CPUs of the NUMA node `device` is attached to, so that
 * pinned buffers allocated afterwards and the copies that use them stay on the GPU's socket. Returns the node, or -1 when the
 * topology is not visible (nothing changes then). The stream workers do this themselves. */
SDR_API int sdr_bind_host_to_device(int device);

#ifdef __cplusplus
}
#endif
#endif /* SDR_C_H */
