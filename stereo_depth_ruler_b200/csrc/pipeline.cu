// Whole hot path per frame, device-resident end to end: the loop body of
// StereoDisplayer::show_disparity_overlay (stereo_vision/src/stereo_displayer.cpp:155-162:
// split the SBS frame, rectify, disparity, depth) composed with save_and_display_pointcloud
// (point_cloud/src/pcd_write.cpp:82-130: gray, SGBM, /16, reproject, colourise, VoxelGrid).
#include "sdr_internal.h"

#include <cstring>

struct sdr_pipeline {
    sdr_ctx* ctx = nullptr;
    sdr_calibration calib{};
    sdr_pipeline_options opt{};
    sdr_rectifier* rect = nullptr;
    sdr_sgbm* sgbm = nullptr;
    sdr_cloud* cloud = nullptr;
    DevBuf sbs, raw, rectL, rectR, grayL, grayR, disp, points, minkey;
    int input_format = SDR_PIX_BGR;   // what the caller's SBS frames are (sdr_pipeline_set_input_format)
    StageTimer timer;
};

static int check_size(sdr_ctx* ctx, const sdr_image* im, int W, int H, const char* what) {
    if (im->cols != W || im->rows != H)
        return SDR_FAIL(ctx, SDR_ESIZE, "pipeline: %s must be %dx%d (the calibration size), got %dx%d", what, W, H, im->cols, im->rows);
    return SDR_OK;
}

// n caller frames <-> a dense device block [n][rows][row_bytes]. One cudaMemcpyAsync for the whole batch when the caller's
// frames are dense (step == row_bytes) and consecutive in memory (slices of one staging block), one per frame when they
// are only dense, a pitched copy per frame otherwise (ROI views).
static int copy_frames(sdr_ctx* ctx, bool to_device, uint8_t* dev, size_t dev_frame, size_t row_bytes, const sdr_image* im, int n, int rows) {
    cudaStream_t st = ctx->stream;
    const size_t frame = row_bytes * rows;
    bool dense = dev_frame == frame, consecutive = true;
    for (int i = 0; i < n; i++) {
        dense = dense && im[i].step_bytes == row_bytes;
        consecutive = consecutive && (const uint8_t*)im[i].data == (const uint8_t*)im[0].data + (size_t)i * frame;
    }
    if (dense && consecutive) {
        if (to_device) SDR_CUDA(ctx, cudaMemcpyAsync(dev, im[0].data, frame * n, cudaMemcpyDefault, st));
        else SDR_CUDA(ctx, cudaMemcpyAsync(im[0].data, dev, frame * n, cudaMemcpyDefault, st));
        return SDR_OK;
    }
    for (int i = 0; i < n; i++) {
        uint8_t* d = dev + (size_t)i * dev_frame;
        if (im[i].step_bytes == row_bytes) {
            if (to_device) SDR_CUDA(ctx, cudaMemcpyAsync(d, im[i].data, frame, cudaMemcpyDefault, st));
            else SDR_CUDA(ctx, cudaMemcpyAsync(im[i].data, d, frame, cudaMemcpyDefault, st));
        } else if (to_device) {
            SDR_CUDA(ctx, cudaMemcpy2DAsync(d, row_bytes, im[i].data, im[i].step_bytes, row_bytes, rows, cudaMemcpyDefault, st));
        } else {
            SDR_CUDA(ctx, cudaMemcpy2DAsync(im[i].data, im[i].step_bytes, d, row_bytes, row_bytes, rows, cudaMemcpyDefault, st));
        }
    }
    return SDR_OK;
}

extern "C" int sdr_pipeline_create(sdr_ctx* ctx, const sdr_calibration* calib, const sdr_pipeline_options* opt, sdr_pipeline** out) {
    if (!ctx || !calib || !opt || !out) return SDR_EINVAL;
    if (opt->batch < 1) return SDR_FAIL(ctx, SDR_EINVAL, "pipeline: batch must be >= 1");
    sdr_pipeline* p = new sdr_pipeline();
    p->ctx = ctx;
    p->calib = *calib;
    p->opt = *opt;
    int rc = sdr_rectifier_create(ctx, calib, &p->rect);
    if (rc == SDR_OK) rc = sdr_sgbm_create(ctx, &opt->sgbm, &p->sgbm);
    if (rc == SDR_OK) rc = sdr_cloud_create(ctx, &p->cloud);
    if (rc != SDR_OK) {
        sdr_pipeline_destroy(p);
        return rc;
    }
    *out = p;
    return SDR_OK;
}

extern "C" void sdr_pipeline_destroy(sdr_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    if (p->rect) sdr_rectifier_destroy(p->rect);
    if (p->sgbm) sdr_sgbm_destroy(p->sgbm);
    if (p->cloud) sdr_cloud_destroy(p->cloud);
    for (DevBuf* b : {&p->sbs, &p->raw, &p->rectL, &p->rectR, &p->grayL, &p->grayR, &p->disp, &p->points, &p->minkey}) b->release();
    p->timer.release();
    delete p;
}

extern "C" int sdr_pipeline_run(sdr_pipeline* p, int n, const sdr_image* sbs, sdr_image* disp, sdr_image* points, sdr_image* left_rect,
                                sdr_cloud_view* views) {
    if (!p || n <= 0 || !sbs) return SDR_EINVAL;
    sdr_ctx* ctx = p->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n > p->opt.batch) return SDR_FAIL(ctx, SDR_EINVAL, "pipeline: %d frames submitted, batch is %d", n, p->opt.batch);
    const int W = p->calib.image_width, H = p->calib.image_height;
    cudaStream_t st = ctx->stream;
    int rc;
    const int fmt = p->input_format;
    int in_rows;
    size_t in_row_bytes;
    sdr_ingest_layout(fmt, 2 * W, H, &in_rows, &in_row_bytes);
    for (int i = 0; i < n; i++) {
        if ((rc = sdr_check_image(ctx, &sbs[i], fmt == SDR_PIX_BGR ? 3 : fmt == SDR_PIX_NV12 ? 1 : 2, 1, "sbs"))) return rc;
        if (sbs[i].cols != 2 * W || sbs[i].rows != in_rows)
            return SDR_FAIL(ctx, SDR_ESIZE, "pipeline: SBS frame must be %d cols x %d rows in this input format, got %dx%d", 2 * W, in_rows,
                            sbs[i].cols, sbs[i].rows);
    }
    // every caller-owned output is checked before anything is queued: exactly W x H (a smaller Mat, e.g. the half-resolution
    // map of the reference's live path, would be overrun by the copies below)
    for (int i = 0; i < n; i++) {
        if (disp && ((rc = sdr_check_image(ctx, &disp[i], 1, 2, "disp")) || (rc = check_size(ctx, &disp[i], W, H, "disp")))) return rc;
        if (points && ((rc = sdr_check_image(ctx, &points[i], 3, 4, "points")) || (rc = check_size(ctx, &points[i], W, H, "points")))) return rc;
        if (left_rect && ((rc = sdr_check_image(ctx, &left_rect[i], 3, 1, "left_rect")) || (rc = check_size(ctx, &left_rect[i], W, H, "left_rect"))))
            return rc;
    }
    p->timer.begin(st);
    // 1. frames to the device (used in place when they are already there at a constant stride)
    const size_t sbs_row = in_row_bytes, sbs_frame_c = sbs_row * in_rows;
    const uint8_t* S0;
    size_t sstep, sframe;
    bool in_place = sbs[0].mem == SDR_MEM_DEVICE;
    sframe = n > 1 ? (size_t)((const uint8_t*)sbs[1].data - (const uint8_t*)sbs[0].data) : 0;
    for (int i = 0; i < n && in_place; i++)
        in_place = sbs[i].mem == SDR_MEM_DEVICE && sbs[i].step_bytes == sbs[0].step_bytes &&
                   (const uint8_t*)sbs[i].data == (const uint8_t*)sbs[0].data + (size_t)i * sframe;
    if (in_place) {
        S0 = (const uint8_t*)sbs[0].data; sstep = sbs[0].step_bytes;
    } else {
        DevBuf& land = fmt == SDR_PIX_BGR ? p->sbs : p->raw;
        if ((rc = land.reserve(ctx, sbs_frame_c * n))) return rc;
        if ((rc = copy_frames(ctx, true, land.as<uint8_t>(), sbs_frame_c, sbs_row, sbs, n, in_rows))) return rc;
        S0 = land.as<uint8_t>(); sstep = sbs_row; sframe = sbs_frame_c;
    }
    p->timer.mark(st, "h2d");
    if (fmt != SDR_PIX_BGR) {   // decoder / camera format -> the BGR frame cv::VideoCapture would have handed out
        const size_t bgr_row = (size_t)2 * W * 3, bgr_frame_c = bgr_row * H;
        if ((rc = p->sbs.reserve(ctx, bgr_frame_c * n))) return rc;
        if ((rc = sdr_launch_ingest(ctx, fmt, S0, sstep, sframe, p->sbs.as<uint8_t>(), bgr_row, bgr_frame_c, 2 * W, H, n))) return rc;
        S0 = p->sbs.as<uint8_t>(); sstep = bgr_row; sframe = bgr_frame_c;
        p->timer.mark(st, "ingest");
    }
    // 2. rectify both eyes (+ gray) in one launch
    const size_t bgr_frame = (size_t)W * H * 3, gray_frame = (size_t)W * H;
    if ((rc = p->rectL.reserve(ctx, bgr_frame * n)) || (rc = p->rectR.reserve(ctx, bgr_frame * n)) ||
        (rc = p->grayL.reserve(ctx, gray_frame * n)) || (rc = p->grayR.reserve(ctx, gray_frame * n)) ||
        (rc = p->disp.reserve(ctx, gray_frame * 2 * n)) || (rc = p->minkey.reserve(ctx, 256 + 4 * (size_t)n)))
        return rc;
    {
        const uint8_t* src[2] = {S0, S0 + (size_t)W * 3};
        const size_t ss[2] = {sstep, sstep}, sf[2] = {sframe, sframe};
        uint8_t* dst[2] = {p->rectL.as<uint8_t>(), p->rectR.as<uint8_t>()};
        const size_t ds[2] = {(size_t)W * 3, (size_t)W * 3}, df[2] = {bgr_frame, bgr_frame};
        uint8_t* gr[2] = {p->grayL.as<uint8_t>(), p->grayR.as<uint8_t>()};
        const size_t gs[2] = {(size_t)W, (size_t)W}, gf[2] = {gray_frame, gray_frame};
        if ((rc = sdr_launch_rectify(p->rect, n, 3, src, ss, sf, W, H, dst, ds, df, gr, gs, gf))) return rc;
    }
    p->timer.mark(st, "rectify");
    // 3. disparity
    if ((rc = sdr_sgbm_run_device(p->sgbm, n, p->grayL.as<uint8_t>(), W, gray_frame, p->grayR.as<uint8_t>(), W, gray_frame, W, H,
                                  p->disp.as<int16_t>(), (size_t)W * 2, gray_frame * 2)))
        return rc;
    p->timer.mark(st, "sgbm");
    // 4. depth image (optional) and clouds (reprojection fused into the voxel kernels)
    const bool need_points = p->opt.want_points || points;
    if (need_points) {
        if ((rc = p->points.reserve(ctx, (size_t)W * H * 12 * n))) return rc;
        if ((rc = sdr_launch_reproject(ctx, true, p->disp.as<uint8_t>(), (size_t)W * 2, gray_frame * 2, W, H, n, p->calib.Q, p->opt.handle_missing,
                                       p->minkey.as<int>(), p->points.as<float>(), (size_t)W * 12, (size_t)W * H * 12)))
            return rc;
        p->timer.mark(st, "reproject");
    }
    if (views) {
        if ((rc = sdr_cloud_enqueue(p->cloud, n, nullptr, 0, 0, p->disp.as<uint8_t>(), (size_t)W * 2, gray_frame * 2, p->rectL.as<uint8_t>(),
                                    (size_t)W * 3, bgr_frame, W, H, p->calib.Q, p->opt.handle_missing, &p->opt.voxel)))
            return rc;
        p->timer.mark(st, "voxelgrid");
    }
    // 5. outputs (one copy per output kind when the caller's frames are dense and consecutive, e.g. slices of one pinned block)
    if (disp && (rc = copy_frames(ctx, false, p->disp.as<uint8_t>(), gray_frame * 2, (size_t)W * 2, disp, n, H))) return rc;
    if (points && (rc = copy_frames(ctx, false, p->points.as<uint8_t>(), (size_t)W * H * 12, (size_t)W * 12, points, n, H))) return rc;
    if (left_rect && (rc = copy_frames(ctx, false, p->rectL.as<uint8_t>(), bgr_frame, (size_t)W * 3, left_rect, n, H))) return rc;
    p->timer.mark(st, "d2h");
    int status = SDR_OK;
    if (views) {
        rc = sdr_cloud_finish(p->cloud, views);
        if (rc == SDR_EVOXEL_OVERFLOW) status = rc;
        else if (rc) return rc;
    }
    SDR_CUDA(ctx, sdr_wait(ctx));
    if ((rc = sdr_sgbm_take_sweep_error(p->sgbm))) return rc;   // a sweep of THIS run gave up on a halo: its maps are garbage
    return status;
}

extern "C" int sdr_pipeline_set_input_format(sdr_pipeline* p, int format) {
    if (!p) return SDR_EINVAL;
    if (format != SDR_PIX_BGR && format != SDR_PIX_NV12 && format != SDR_PIX_YUYV)
        return SDR_FAIL(p->ctx, SDR_EINVAL, "pipeline: unknown input format %d", format);
    if (format != SDR_PIX_BGR && ((p->calib.image_width & 1) || (format == SDR_PIX_NV12 && (p->calib.image_height & 1))))
        return SDR_FAIL(p->ctx, SDR_EINVAL, "pipeline: this input format needs an even picture size");
    p->input_format = format;
    return SDR_OK;
}

extern "C" int sdr_pipeline_set_sgbm_option(sdr_pipeline* p, const char* name, int value) {
    if (!p || !p->sgbm) return SDR_EINVAL;
    return sdr_sgbm_set_option(p->sgbm, name, value);
}

extern "C" int sdr_pipeline_set_profiling(sdr_pipeline* p, int enabled) {
    if (!p) return SDR_EINVAL;
    p->timer.enabled = enabled != 0;
    sdr_sgbm_timer(p->sgbm)->enabled = enabled != 0;
    return SDR_OK;
}

extern "C" int sdr_pipeline_stage_times(sdr_pipeline* p, int cap, const char** names, float* ms) {
    if (!p) return 0;
    cudaSetDevice(p->ctx->device);
    return p->timer.collect(cap, names, ms);
}

extern "C" int sdr_pipeline_stage_timeline(sdr_pipeline* p, void* base_cuda_event, int which, int cap, const char** names, float* t_ms) {
    if (!p || !base_cuda_event) return 0;
    cudaSetDevice(p->ctx->device);
    StageTimer* t = which == 0 ? &p->timer : sdr_sgbm_timer(p->sgbm);
    return t->collect_since((cudaEvent_t)base_cuda_event, cap, names, t_ms);
}

extern "C" int sdr_pipeline_cloud_download(sdr_pipeline* p, int i, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt) {
    if (!p || !p->cloud) return SDR_EINVAL;
    cudaSetDevice(p->ctx->device);
    return sdr_cloud_download_frame(p->cloud, i, xyz, rgb, idx, cnt, true);
}

extern "C" int sdr_pipeline_clouds_download(sdr_pipeline* p, int n, float* const* xyz, uint8_t* const* rgb, uint32_t* const* idx,
                                            uint32_t* const* cnt) {
    if (!p || !p->cloud || n <= 0) return SDR_EINVAL;
    sdr_ctx* ctx = p->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < n; i++) {
        int rc = sdr_cloud_download_frame(p->cloud, i, xyz ? xyz[i] : nullptr, rgb ? rgb[i] : nullptr, idx ? idx[i] : nullptr,
                                          cnt ? cnt[i] : nullptr, false);
        if (rc) return rc;
    }
    SDR_CUDA(ctx, sdr_wait(ctx));   // one host wait for the whole batch
    return SDR_OK;
}

extern "C" int sdr_pipeline_clouds_download_packed(sdr_pipeline* p, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt, int64_t* offsets) {
    if (!p || !p->cloud) return SDR_EINVAL;
    SDR_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    return sdr_cloud_download_packed(p->cloud, xyz, rgb, idx, cnt, offsets, true);
}
