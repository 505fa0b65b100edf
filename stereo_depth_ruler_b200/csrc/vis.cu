// The reference's live path around the matcher (stereo_vision/src/stereo_disparity.cpp):
//   sdr_resize_half_area      cv::resize(gray, Size(), 0.5, 0.5, INTER_AREA)                      :23-24
//   sdr_vis_disparity         StereoDisparity::show_disparityMap (mask, /numDisp, pow 0.6, x255, temporal blend)   :42-73
//   sdr_vis_depth             StereoDisparity::show_depthMap (valid min/max, EMA range, convertTo u8, TURBO, blend) :83-124
// OpenCV arithmetic restated (probed against cv2 4.13.0, see tests/test_gpu_vis.py):
//   convertTo(CV_8U, a, b)      = saturate_u8(cvRound(fmaf(x, (float)a, (float)b))), round half to even; values the x86
//                                 float->int conversion cannot represent (NaN, +-inf, |v| >= 2^31) become INT_MIN -> 0
//   addWeighted(p, al, s, be)   = saturate_u8(cvRound(fmaf(p, (float)al, s * (float)be)))  on u8
//   pow(x, 0.6) on CV_32F       : cv2's result times 255, rounded, equals the correctly rounded power for every input a
//                                 CV_16S / 16 disparity can produce (exhaustive test); evaluated here in fp64
#include "sdr_internal.h"
#include "turbo_lut.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

__constant__ unsigned char c_turbo[256][3];

__device__ __forceinline__ int cv_round_sat_u8(float v) {
    if (!(v > -2147483648.f && v < 2147483648.f)) return 0;   // cvtss2si gives INT_MIN for NaN / out of range; saturate_cast -> 0
    const int r = __float2int_rn(v);
    return min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256) k_resize_half(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst, size_t dstep,
                                                     int ow, int oh, int ch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;   // i: element of the output row (pixel * channel)
    if (i >= ow * ch) return;
    const int x = i / ch, c = i - x * ch;
    const uint8_t* r0 = src + (size_t)(2 * y) * sstep + (size_t)(2 * x) * ch + c;
    const uint8_t* r1 = r0 + sstep;
    dst[(size_t)y * dstep + i] = (uint8_t)(((int)r0[0] + (int)r0[ch] + (int)r1[0] + (int)r1[ch] + 2) >> 2);
}

__global__ void __launch_bounds__(256) k_vis_disp(const uint8_t* __restrict__ disp, size_t dstep, int W, int H, float inv_numdisp,
                                                  uint8_t* __restrict__ prev, int have_prev, uint8_t* __restrict__ out, size_t ostep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float d = ((const float*)(disp + (size_t)y * dstep))[x];
    if (!(d > 0.f)) d = 0.f;                                    // valid_mask = disparity > 0; setTo(0, ~valid_mask)
    const float n01 = __fmul_rn(d, inv_numdisp);                // convertTo(CV_32F, 1.0f / max(1, numDisp))
    const float g = (float)pow((double)n01, 0.6);               // cv::pow(norm01, 0.6)
    int v = cv_round_sat_u8(__fmul_rn(g, 255.0f));              // convertTo(CV_8U, 255.0)
    if (have_prev) {                                            // addWeighted(prev_vis, 0.63f, show, 1 - 0.63f, 0)
        const float al = 0.63f, be = 1.0f - 0.63f;
        v = cv_round_sat_u8(__fmaf_rn((float)prev[(size_t)y * W + x], al, __fmul_rn((float)v, be)));
    }
    prev[(size_t)y * W + x] = (uint8_t)v;                       // prev_vis = show_disp.clone(); return prev_vis
    out[(size_t)y * ostep + x] = (uint8_t)v;
}

__device__ __forceinline__ int fkey(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
inline float kfloat_host(int k) {
    const int i = k >= 0 ? k : k ^ 0x7FFFFFFF;
    float f;
    memcpy(&f, &i, 4);
    return f;
}

// minMaxLoc(depthZ, mask = (z > 0) & (z < 10000) & (z == z)) as ordered-int keys; mm[0] = min, mm[1] = max, mm[2] = count > 0
__global__ void __launch_bounds__(256) k_vis_minmax(const uint8_t* __restrict__ pts, size_t step, int W, int H, int ch, int* mm) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    int kmin = 0x7FFFFFFF, kmax = (int)0x80000000;
    if (x < W) {
        const float z = ((const float*)(pts + (size_t)y * step))[(size_t)x * ch + (ch == 3 ? 2 : 0)];
        if (z > 0.f && z < 10000.f) { kmin = kmax = fkey(z); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && kmin != 0x7FFFFFFF) {
        atomicMin(&mm[0], kmin);
        atomicMax(&mm[1], kmax);
        mm[2] = 1;
    }
}

__global__ void __launch_bounds__(256) k_vis_depth(const uint8_t* __restrict__ pts, size_t step, int W, int H, int ch, float a, float b,
                                                   uint8_t* __restrict__ prev, int have_prev, uint8_t* __restrict__ out, size_t ostep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float z = ((const float*)(pts + (size_t)y * step))[(size_t)x * ch + (ch == 3 ? 2 : 0)];
    const int g = cv_round_sat_u8(__fmaf_rn(z, a, b));          // depthZ.convertTo(depth8u, CV_8U, a, b)
    uint8_t* pv = prev + ((size_t)y * W + x) * 3;
    uint8_t* o = out + (size_t)y * ostep + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int v = c_turbo[g][c];                                  // applyColorMap(COLORMAP_TURBO), BGR
        if (have_prev) v = cv_round_sat_u8(__fmaf_rn((float)pv[c], 0.63f, __fmul_rn((float)v, 1.0f - 0.63f)));
        pv[c] = (uint8_t)v;
        o[c] = (uint8_t)v;
    }
}

}  // namespace

struct sdr_vis {
    sdr_ctx* ctx = nullptr;
    DevBuf prev_disp, prev_depth, in, out, mm;
    PinnedBuf host_mm;
    int pd_w = 0, pd_h = 0, pz_w = 0, pz_h = 0;      // sizes of prev_vis / prev_depth_vis (0 = empty)
    double zmin_smooth = 1000.0, zmax_smooth = 2000.0;   // the function-static EMA state of show_depthMap (stereo_disparity.cpp:93)
    bool lut_loaded = false;
};

extern "C" int sdr_vis_create(sdr_ctx* ctx, sdr_vis** out) {
    if (!ctx || !out) return SDR_EINVAL;
    sdr_vis* v = new sdr_vis();
    v->ctx = ctx;
    *out = v;
    return SDR_OK;
}

extern "C" void sdr_vis_destroy(sdr_vis* v) {
    if (!v) return;
    cudaSetDevice(v->ctx->device);
    for (DevBuf* b : {&v->prev_disp, &v->prev_depth, &v->in, &v->out, &v->mm}) b->release();
    v->host_mm.release();
    delete v;
}

extern "C" int sdr_vis_reset(sdr_vis* v) {
    if (!v) return SDR_EINVAL;
    v->pd_w = v->pd_h = v->pz_w = v->pz_h = 0;
    v->zmin_smooth = 1000.0;
    v->zmax_smooth = 2000.0;
    return SDR_OK;
}

extern "C" int sdr_vis_disparity(sdr_vis* v, const sdr_image* disp32f, int num_disparities, sdr_image* out) {
    if (!v || !disp32f || !out) return SDR_EINVAL;
    sdr_ctx* ctx = v->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, disp32f, 1, 4, "disparity")) || (rc = sdr_check_image(ctx, out, 1, 1, "vis"))) return rc;
    const int W = disp32f->cols, H = disp32f->rows;
    if (out->cols != W || out->rows != H) return SDR_FAIL(ctx, SDR_ESIZE, "vis_disparity: size mismatch");
    DevView vi, vo;
    if ((rc = sdr_stage_in(ctx, disp32f, v->in, &vi)) || (rc = sdr_stage_out_begin(ctx, out, v->out, &vo))) return rc;
    if ((rc = v->prev_disp.reserve(ctx, (size_t)W * H))) return rc;
    const int have_prev = v->pd_w == W && v->pd_h == H;       // !prev_vis.empty() && prev_vis.size() == show_disp.size()
    const float inv = 1.0f / (float)std::max(1, num_disparities);
    k_vis_disp<<<dim3(div_up(W, 256), H), 256, 0, ctx->stream>>>(vi.p, vi.pitch, W, H, inv, v->prev_disp.as<uint8_t>(), have_prev, vo.p, vo.pitch);
    SDR_LAUNCHED(ctx);
    v->pd_w = W; v->pd_h = H;
    return sdr_stage_out_end(ctx, out, vo);
}

extern "C" int sdr_vis_depth(sdr_vis* v, const sdr_image* depth, sdr_image* out) {
    if (!v || !depth || !out) return SDR_EINVAL;
    sdr_ctx* ctx = v->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const int ch = depth->channels;
    if (ch != 1 && ch != 3) return SDR_FAIL(ctx, SDR_ESIZE, "vis_depth: expected a 1- or 3-channel float image");
    if ((rc = sdr_check_image(ctx, depth, ch, 4, "depth")) || (rc = sdr_check_image(ctx, out, 3, 1, "vis"))) return rc;
    const int W = depth->cols, H = depth->rows;
    if (out->cols != W || out->rows != H) return SDR_FAIL(ctx, SDR_ESIZE, "vis_depth: size mismatch");
    if (!v->lut_loaded) {
        SDR_CUDA(ctx, cudaMemcpyToSymbol(c_turbo, kTurboBGR, sizeof kTurboBGR));
        v->lut_loaded = true;
    }
    DevView vi, vo;
    if ((rc = sdr_stage_in(ctx, depth, v->in, &vi)) || (rc = sdr_stage_out_begin(ctx, out, v->out, &vo))) return rc;
    if ((rc = v->prev_depth.reserve(ctx, (size_t)W * H * 3)) || (rc = v->mm.reserve(ctx, 16)) || (rc = v->host_mm.reserve(ctx, 16))) return rc;
    const int init[3] = {0x7FFFFFFF, (int)0x80000000, 0};
    SDR_CUDA(ctx, cudaMemcpyAsync(v->mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    k_vis_minmax<<<dim3(div_up(W, 256), H), 256, 0, ctx->stream>>>(vi.p, vi.pitch, W, H, ch, v->mm.as<int>());
    SDR_LAUNCHED(ctx);
    SDR_CUDA(ctx, cudaMemcpyAsync(v->host_mm.p, v->mm.p, 12, cudaMemcpyDeviceToHost, ctx->stream));
    SDR_CUDA(ctx, sdr_wait(ctx));
    const int* hm = (const int*)v->host_mm.p;
    double zmin_raw = 0, zmax_raw = 0;                          // minMaxLoc leaves them untouched (0) on an empty mask
    if (hm[2]) { zmin_raw = (double)kfloat_host(hm[0]); zmax_raw = (double)kfloat_host(hm[1]); }
    if (!(zmax_raw > zmin_raw)) { zmin_raw = 1000.0; zmax_raw = 2000.0; }
    const double alpha = 0.1;
    v->zmin_smooth = (1.0 - alpha) * v->zmin_smooth + alpha * zmin_raw;
    v->zmax_smooth = (1.0 - alpha) * v->zmax_smooth + alpha * zmax_raw;
    v->zmin_smooth = std::max(0.0, std::min(v->zmin_smooth, 10000.0));
    v->zmax_smooth = std::max(v->zmin_smooth + 1.0, std::min(v->zmax_smooth, 10000.0));
    const double a = 255.0 / (v->zmax_smooth - v->zmin_smooth), b = -255.0 * v->zmin_smooth / (v->zmax_smooth - v->zmin_smooth);
    const int have_prev = v->pz_w == W && v->pz_h == H;
    k_vis_depth<<<dim3(div_up(W, 256), H), 256, 0, ctx->stream>>>(vi.p, vi.pitch, W, H, ch, (float)a, (float)b, v->prev_depth.as<uint8_t>(),
                                                                 have_prev, vo.p, vo.pitch);
    SDR_LAUNCHED(ctx);
    v->pz_w = W; v->pz_h = H;
    return sdr_stage_out_end(ctx, out, vo);
}

extern "C" int sdr_vis_depth_range(sdr_vis* v, double* zmin_smooth, double* zmax_smooth) {
    if (!v) return SDR_EINVAL;
    if (zmin_smooth) *zmin_smooth = v->zmin_smooth;
    if (zmax_smooth) *zmax_smooth = v->zmax_smooth;
    return SDR_OK;
}

extern "C" int sdr_resize_half_area(sdr_ctx* ctx, const sdr_image* src, sdr_image* dst) {
    if (!ctx || !src || !dst) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const int ch = src->channels;
    if (ch != 1 && ch != 3) return SDR_FAIL(ctx, SDR_ESIZE, "resize_half_area: 1 or 3 channels");
    if ((rc = sdr_check_image(ctx, src, ch, 1, "src")) || (rc = sdr_check_image(ctx, dst, ch, 1, "dst"))) return rc;
    // cv::resize takes its integer-factor INTER_AREA path ((a+b+c+d+2)>>2) only for an exact factor of 2
    if ((src->cols & 1) || (src->rows & 1)) return SDR_FAIL(ctx, SDR_EINVAL, "resize_half_area: width and height must be even (got %dx%d)", src->cols, src->rows);
    if (dst->cols != src->cols / 2 || dst->rows != src->rows / 2) return SDR_FAIL(ctx, SDR_ESIZE, "resize_half_area: dst must be %dx%d", src->cols / 2, src->rows / 2);
    DevBuf bin, bout;
    DevView vi, vo;
    rc = sdr_stage_in(ctx, src, bin, &vi);
    if (rc == SDR_OK) rc = sdr_stage_out_begin(ctx, dst, bout, &vo);
    if (rc == SDR_OK) {
        k_resize_half<<<dim3(div_up(dst->cols * ch, 256), dst->rows), 256, 0, ctx->stream>>>(vi.p, vi.pitch, vo.p, vo.pitch, dst->cols, dst->rows, ch);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = SDR_FAIL(ctx, SDR_ECUDA, "resize_half_area launch failed");
    }
    if (rc == SDR_OK) rc = sdr_stage_out_end(ctx, dst, vo);
    sdr_wait(ctx);          // the scratch buffers are released below
    bin.release(); bout.release();
    return rc;
}

// cv::ximgproc::createRightMatcher(left) for a StereoSGBM (stereo_disparity.cpp:10): the matcher that is run on
// (right, left) to give the right-view disparity the WLS filter consumes.
extern "C" int sdr_sgbm_right_matcher_params(const sdr_sgbm_params* left, sdr_sgbm_params* right) {
    if (!left || !right) return SDR_EINVAL;
    *right = *left;
    right->minDisparity = -(left->minDisparity + left->numDisparities) + 1;
    right->uniquenessRatio = 0;
    right->disp12MaxDiff = 1000000;
    right->speckleWindowSize = 0;
    right->speckleRange = 0;    // StereoSGBM::create default
    return SDR_OK;
}
