// Rectification (cv::initUndistortRectifyMap + cv::remap fused into one bilinear kernel),
// cv::cvtColor(BGR2GRAY) and cv::reprojectImageTo3D for sm_100a; replaces
// StereoRectifier::{StereoRectifier, rectify} (stereo_vision/src/stereo_rectifier.cpp:6-14, :17-41),
// the cvtColor calls at stereo_disparity.cpp:19-20 / pcd_write.cpp:88-89 and
// StereoDisparity::computeDepth (stereo_disparity.cpp:76-80) / pcd_write.cpp:112-116.
// Arithmetic follows SURVEY.md Appendix A.1, A.2, A.7 exactly: the map evaluation is fp64 with
// every operation individually rounded (no FMA contraction), the bilinear blend is the 15-bit
// fixed-point form, and reprojection keeps OpenCV's float-then-reciprocal rounding quirk.
#include "sdr_internal.h"

#include <cmath>
#include <cstring>

namespace {

struct MapParams {  // one eye
    double iR[9];
    double fx, fy, cx, cy, k1, k2, p1, p2, k3;
};

__device__ __forceinline__ void eval_map(const MapParams& m, int j, int i, int& sx, int& sy, int& fxq, int& fyq) {
    const double dj = (double)j, di = (double)i;
    const double _x = __dadd_rn(__dadd_rn(__dmul_rn(dj, m.iR[0]), __dmul_rn(di, m.iR[1])), m.iR[2]);
    const double _y = __dadd_rn(__dadd_rn(__dmul_rn(dj, m.iR[3]), __dmul_rn(di, m.iR[4])), m.iR[5]);
    const double _w = __dadd_rn(__dadd_rn(__dmul_rn(dj, m.iR[6]), __dmul_rn(di, m.iR[7])), m.iR[8]);
    const double w = __ddiv_rn(1.0, _w);
    const double x = __dmul_rn(_x, w), y = __dmul_rn(_y, w);
    const double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y);
    const double r2 = __dadd_rn(x2, y2), _2xy = __dmul_rn(__dmul_rn(2.0, x), y);
    const double kr =
        __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(m.k3, r2), m.k2), r2), m.k1), r2));
    const double xd = __dadd_rn(__dadd_rn(__dmul_rn(x, kr), __dmul_rn(m.p1, _2xy)),
                                __dmul_rn(m.p2, __dadd_rn(r2, __dmul_rn(2.0, x2))));
    const double yd = __dadd_rn(__dadd_rn(__dmul_rn(y, kr), __dmul_rn(m.p1, __dadd_rn(r2, __dmul_rn(2.0, y2)))),
                                __dmul_rn(m.p2, _2xy));
    const double u = __dadd_rn(__dmul_rn(m.fx, xd), m.cx);
    const double v = __dadd_rn(__dmul_rn(m.fy, yd), m.cy);
    // round-half-even == cvRound; a value beyond the int range is INT_MIN there (cvtsd2si), where the GPU's conversion saturates
    const double u32 = __dmul_rn(u, 32.0), v32 = __dmul_rn(v, 32.0);
    const int iu = fabs(u32) < 2147483648.0 ? __double2int_rn(u32) : (int)0x80000000;
    const int iv = fabs(v32) < 2147483648.0 ? __double2int_rn(v32) : (int)0x80000000;
    // CV_16SC2 storage wraps to int16 (saturate_cast in OpenCV; identical while |coord| < 32768)
    sx = max(min(iu >> 5, 32767), -32768);
    sy = max(min(iv >> 5, 32767), -32768);
    fxq = iu & 31;
    fyq = iv & 31;
}

__global__ void __launch_bounds__(256) k_make_maps(MapParams m, int W, int H, int16_t* map1, uint16_t* map2) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int sx, sy, fx, fy;
    eval_map(m, x, y, sx, sy, fx, fy);
    map1[((size_t)y * W + x) * 2] = (int16_t)sx;
    map1[((size_t)y * W + x) * 2 + 1] = (int16_t)sy;
    map2[(size_t)y * W + x] = (uint16_t)(fy * 32 + fx);
}

// Six source bytes starting at byte address p (any alignment) as two words, through aligned 32-bit loads.
__device__ __forceinline__ void load6(const uint8_t* p, uint32_t& lo, uint32_t& hi) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

template <int CH>
__device__ __forceinline__ void sample(const uint8_t* __restrict__ src, size_t step, int sw, int sh, bool aligned, int sx, int sy,
                                       int fx, int fy, int (&out)[CH]) {
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    if (CH == 3 && aligned && sx >= 0 && sx + 3 < sw && sy >= 0 && sy + 1 < sh) {
        uint32_t a0, a1, b0, b1;
        load6(src + (size_t)sy * step + (size_t)sx * 3, a0, a1);
        load6(src + (size_t)(sy + 1) * step + (size_t)sx * 3, b0, b1);
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int p00 = (a0 >> (8 * c)) & 0xFF;
            const int p01 = c == 0 ? (a0 >> 24) : (a1 >> (8 * (c - 1))) & 0xFF;
            const int p10 = (b0 >> (8 * c)) & 0xFF;
            const int p11 = c == 0 ? (b0 >> 24) : (b1 >> (8 * (c - 1))) & 0xFF;
            out[c] = (w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + 16384) >> 15;
        }
        return;
    }
#pragma unroll
    for (int c = 0; c < CH; c++) {
        int acc = 16384;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int xx = sx + (t & 1), yy = sy + (t >> 1);
            const int p = (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * step + (size_t)xx * CH + c] : 0;
            acc += (t == 0 ? w00 : t == 1 ? w01 : t == 2 ? w10 : w11) * p;
        }
        out[c] = acc >> 15;
    }
}

struct RectifyArgs {
    const uint8_t* src[2];  // per eye
    size_t src_step[2], src_frame[2];
    uint8_t* dst[2];
    size_t dst_step[2], dst_frame[2];
    uint8_t* gray[2];  // optional
    size_t gray_step[2], gray_frame[2];
    int sw, sh, W, H;
    const int16_t* map1[2];  // optional precomputed maps (when !FUSED)
    const uint16_t* map2[2];
};

// Each thread produces 4 consecutive destination pixels of one eye: 3 word stores of BGR (+1 of gray).
template <int CH, bool FUSED>
__global__ void __launch_bounds__(128) k_rectify(RectifyArgs a, MapParams mL, MapParams mR) {
    const int eye = blockIdx.z & 1, b = blockIdx.z >> 1;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= a.W) return;
    const MapParams& m = eye ? mR : mL;
    const uint8_t* src = a.src[eye] + (size_t)b * a.src_frame[eye];
    const bool aligned = (((uintptr_t)src | a.src_step[eye]) & 3) == 0;
    uint8_t px[4][CH];
    uint8_t gr[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = min(x0 + i, a.W - 1);
        int sx, sy, fx, fy;
        if (FUSED) {
            eval_map(m, x, y, sx, sy, fx, fy);
        } else {
            const short2 s = *(const short2*)(a.map1[eye] + ((size_t)y * a.W + x) * 2);
            const int f = a.map2[eye][(size_t)y * a.W + x];
            sx = s.x; sy = s.y; fx = f & 31; fy = (f >> 5) & 31;
        }
        int v[CH];
        sample<CH>(src, a.src_step[eye], a.sw, a.sh, aligned, sx, sy, fx, fy, v);
#pragma unroll
        for (int c = 0; c < CH; c++) px[i][c] = (uint8_t)v[c];
        if (CH == 3) gr[i] = (uint8_t)((3735 * v[0] + 19235 * v[1] + 9798 * v[CH - 1] + 16384) >> 15);
        else gr[i] = (uint8_t)v[0];
    }
    uint8_t* drow = a.dst[eye] + (size_t)b * a.dst_frame[eye] + (size_t)y * a.dst_step[eye];
    const bool full = x0 + 3 < a.W;
    if (full && ((((uintptr_t)drow) | a.dst_step[eye]) & 3) == 0) {
        const uint8_t* f = &px[0][0];
        uint32_t* d = (uint32_t*)(drow + (size_t)x0 * CH);
#pragma unroll
        for (int w = 0; w < CH; w++)
            d[w] = (uint32_t)f[4 * w] | ((uint32_t)f[4 * w + 1] << 8) | ((uint32_t)f[4 * w + 2] << 16) | ((uint32_t)f[4 * w + 3] << 24);
    } else {
        for (int i = 0; i < 4 && x0 + i < a.W; i++)
#pragma unroll
            for (int c = 0; c < CH; c++) drow[(size_t)(x0 + i) * CH + c] = px[i][c];
    }
    if (a.gray[eye]) {
        uint8_t* grow = a.gray[eye] + (size_t)b * a.gray_frame[eye] + (size_t)y * a.gray_step[eye];
        if (full && ((((uintptr_t)grow) | a.gray_step[eye]) & 3) == 0)
            *(uint32_t*)(grow + x0) = (uint32_t)gr[0] | ((uint32_t)gr[1] << 8) | ((uint32_t)gr[2] << 16) | ((uint32_t)gr[3] << 24);
        else
            for (int i = 0; i < 4 && x0 + i < a.W; i++) grow[x0 + i] = gr[i];
    }
}

__global__ void __launch_bounds__(256) k_bgr2gray(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst,
                                                  size_t dstep, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* p = src + (size_t)y * sstep + (size_t)x * 3;
    dst[(size_t)y * dstep + x] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
}

// ---- reprojectImageTo3D ---------------------------------------------------------------------
struct QMat {
    double q[16];
};

// minimum disparity of the image, as an order-preserving int key of the float value
__device__ __forceinline__ int float_key(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

template <bool IS16>
__global__ void __launch_bounds__(256) k_min_disp(const uint8_t* __restrict__ disp, size_t step, size_t frame, int W, int H, int* out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    int key = 0x7FFFFFFF;
    if (x < W) {
        const uint8_t* r = disp + (size_t)b * frame + (size_t)y * step;
        const float d = IS16 ? (float)((const int16_t*)r)[x] * 0.0625f : ((const float*)r)[x];
        if (d == d) key = float_key(d);  // NaN never compares less in cv::minMaxLoc
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out + b, key);
}

__global__ void k_set_int(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <bool IS16>
__global__ void __launch_bounds__(256) k_reproject(const uint8_t* __restrict__ disp, size_t step, size_t frame, int W, int H, QMat Q,
                                                   int handle_missing, const int* __restrict__ min_key, float* __restrict__ out,
                                                   size_t ostep, size_t oframe) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const uint8_t* r = disp + (size_t)b * frame + (size_t)y * step;
    const float df = IS16 ? (float)((const int16_t*)r)[x] * 0.0625f : ((const float*)r)[x];
    const double d = (double)df, dx = (double)x, dy = (double)y;
    double h[4];
#pragma unroll
    for (int c = 0; c < 4; c++)
        h[c] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Q.q[c * 4], dx), __dmul_rn(Q.q[c * 4 + 1], dy)), __dmul_rn(Q.q[c * 4 + 2], d)),
                         Q.q[c * 4 + 3]);
    const double iw = __ddiv_rn(1.0, h[3]);
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; c++) o[c] = (float)__dmul_rn((double)(float)h[c], iw);
    if (handle_missing) {
        const double mind = (double)key_float(min_key[b]);
        if (fabs(d - mind) <= 1.1920928955078125e-07) o[2] = 10000.f;
    }
    float* orow = (float*)((uint8_t*)out + (size_t)b * oframe + (size_t)y * ostep) + (size_t)x * 3;
    orow[0] = o[0];
    orow[1] = o[1];
    orow[2] = o[2];
}

__global__ void __launch_bounds__(256) k_coverage(const uint8_t* __restrict__ pts, size_t step, int W, int H, int first_col, float zmin, float zmax,
                                                  unsigned long long* count) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    unsigned ok = 0;
    if (x < W && x >= first_col) {
        const float z = ((const float*)(pts + (size_t)y * step))[(size_t)x * 3 + 2];
        ok = (z >= zmin && z <= zmax && !(z != z)) ? 1u : 0u;
    }
    const unsigned n = __popc(__ballot_sync(0xffffffffu, ok));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, (unsigned long long)n);
}

__global__ void k_pick2(const uint8_t* __restrict__ pts, size_t step, int x1, int y1, int x2, int y2, float* out) {
    const int i = threadIdx.x;
    if (i < 3) out[i] = ((const float*)(pts + (size_t)y1 * step))[(size_t)x1 * 3 + i];
    else if (i < 6) out[i] = ((const float*)(pts + (size_t)y2 * step))[(size_t)x2 * 3 + i - 3];
}

__global__ void __launch_bounds__(256) k_disp_to_float(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst, size_t dstep, int W,
                                                       int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W) ((float*)(dst + (size_t)y * dstep))[x] = (float)((const int16_t*)(src + (size_t)y * sstep))[x] * 0.0625f;
}

// cv::invert(DECOMP_LU) of a 3x3 (partial pivoting), fp64, on the host
bool invert3(const double* m, double* inv) {
    double a[3][6];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            a[i][j] = m[i * 3 + j];
            a[i][j + 3] = i == j ? 1.0 : 0.0;
        }
    for (int i = 0; i < 3; i++) {
        int k = i;
        for (int j = i + 1; j < 3; j++)
            if (std::fabs(a[j][i]) > std::fabs(a[k][i])) k = j;
        if (std::fabs(a[k][i]) < 2.2204460492503131e-16) return false;
        if (k != i)
            for (int j = 0; j < 6; j++) std::swap(a[i][j], a[k][j]);
        const double d = -1.0 / a[i][i];
        for (int j = i + 1; j < 3; j++) {
            const double alpha = a[j][i] * d;
            for (int c = i + 1; c < 6; c++) a[j][c] += alpha * a[i][c];
        }
    }
    for (int i = 2; i >= 0; i--)
        for (int j = 3; j < 6; j++) {
            double s = a[i][j];
            for (int k = i + 1; k < 3; k++) s -= a[i][k] * a[k][j];
            a[i][j] = s / a[i][i];
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) inv[i * 3 + j] = a[i][j + 3];
    return true;
}

bool make_map_params(const double* K, const double* dist, const double* R, const double* P, MapParams& m) {
    double PR[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += P[i * 4 + k] * R[k * 3 + j];
            PR[i * 3 + j] = s;
        }
    if (!invert3(PR, m.iR)) return false;
    m.fx = K[0]; m.fy = K[4]; m.cx = K[2]; m.cy = K[5];
    m.k1 = dist[0]; m.k2 = dist[1]; m.p1 = dist[2]; m.p2 = dist[3]; m.k3 = dist[4];
    return true;
}

}  // namespace

struct sdr_rectifier {
    sdr_ctx* ctx = nullptr;
    int W = 0, H = 0;
    MapParams m[2];
    DevBuf in[2], out[2], gray[2], maps;
};

extern "C" int sdr_rectifier_create(sdr_ctx* ctx, const sdr_calibration* c, sdr_rectifier** out) {
    if (!ctx || !c || !out) return SDR_EINVAL;
    if (c->image_width <= 0 || c->image_height <= 0) return SDR_FAIL(ctx, SDR_EINVAL, "calibration has no image size");
    sdr_rectifier* r = new sdr_rectifier();
    r->ctx = ctx;
    r->W = c->image_width;
    r->H = c->image_height;
    if (!make_map_params(c->K_left, c->dist_left, c->R1, c->P1, r->m[0]) ||
        !make_map_params(c->K_right, c->dist_right, c->R2, c->P2, r->m[1])) {
        delete r;
        return SDR_FAIL(ctx, SDR_EINVAL, "P*R is singular");
    }
    *out = r;
    return SDR_OK;
}

extern "C" void sdr_rectifier_destroy(sdr_rectifier* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    for (int e = 0; e < 2; e++) { r->in[e].release(); r->out[e].release(); r->gray[e].release(); }
    r->maps.release();
    delete r;
}

extern "C" int sdr_rectifier_get_maps(sdr_rectifier* r, int eye, int16_t* map1, uint16_t* map2) {
    if (!r || eye < 0 || eye > 1 || !map1 || !map2) return SDR_EINVAL;
    sdr_ctx* ctx = r->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)r->W * r->H;
    int rc = r->maps.reserve(ctx, n * 6);
    if (rc) return rc;
    int16_t* d1 = r->maps.as<int16_t>();
    uint16_t* d2 = (uint16_t*)(d1 + n * 2);
    k_make_maps<<<dim3(div_up(r->W, 256), r->H), 256, 0, ctx->stream>>>(r->m[eye], r->W, r->H, d1, d2);
    SDR_LAUNCHED(ctx);
    SDR_CUDA(ctx, cudaMemcpyAsync(map1, d1, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SDR_CUDA(ctx, cudaMemcpyAsync(map2, d2, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    SDR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SDR_OK;
}

// Shared by sdr_rectify and the pipeline: both eyes, `batch` frames, device pointers.
int sdr_launch_rectify(sdr_rectifier* r, int batch, int channels, const uint8_t* const src[2], const size_t src_step[2],
                       const size_t src_frame[2], int sw, int sh, uint8_t* const dst[2], const size_t dst_step[2],
                       const size_t dst_frame[2], uint8_t* const gray[2], const size_t gray_step[2], const size_t gray_frame[2]) {
    sdr_ctx* ctx = r->ctx;
    RectifyArgs a{};
    for (int e = 0; e < 2; e++) {
        a.src[e] = src[e]; a.src_step[e] = src_step[e]; a.src_frame[e] = src_frame[e];
        a.dst[e] = dst[e]; a.dst_step[e] = dst_step[e]; a.dst_frame[e] = dst_frame[e];
        a.gray[e] = gray ? gray[e] : nullptr;
        a.gray_step[e] = gray ? gray_step[e] : 0;
        a.gray_frame[e] = gray ? gray_frame[e] : 0;
    }
    a.sw = sw; a.sh = sh; a.W = r->W; a.H = r->H;
    const dim3 grid(div_up(div_up(r->W, 4), 128), r->H, 2 * batch);
    if (channels == 3)
        k_rectify<3, true><<<grid, 128, 0, ctx->stream>>>(a, r->m[0], r->m[1]);
    else
        k_rectify<1, true><<<grid, 128, 0, ctx->stream>>>(a, r->m[0], r->m[1]);
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

extern "C" int sdr_rectify(sdr_rectifier* r, const sdr_image* ls, const sdr_image* rs, sdr_image* ld, sdr_image* rd, sdr_image* lg,
                           sdr_image* rg) {
    if (!r) return SDR_EINVAL;
    sdr_ctx* ctx = r->ctx;
    // stereo_rectifier.cpp:18-21: empty input is an error (the reference prints and returns)
    if (!ls || !rs || !ls->data || !rs->data) return SDR_FAIL(ctx, SDR_EINVAL, "Input images are empty");
    if (!ld || !rd) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    const int ch = ls->channels;
    if (ch != 1 && ch != 3) return SDR_FAIL(ctx, SDR_ESIZE, "rectify: 1 or 3 channels expected");
    int rc;
    const sdr_image* srcs[2] = {ls, rs};
    sdr_image* dsts[2] = {ld, rd};
    sdr_image* grays[2] = {lg, rg};
    if ((lg == nullptr) != (rg == nullptr)) return SDR_FAIL(ctx, SDR_EINVAL, "rectify: give both gray outputs or none");
    DevView vi[2], vo[2], vg[2];
    const uint8_t* sp[2]; uint8_t* dp[2]; uint8_t* gp[2] = {nullptr, nullptr};
    size_t ss[2], sf[2] = {0, 0}, ds[2], df[2] = {0, 0}, gs[2] = {0, 0}, gf[2] = {0, 0};
    for (int e = 0; e < 2; e++) {
        if ((rc = sdr_check_image(ctx, srcs[e], ch, 1, e ? "right_src" : "left_src"))) return rc;
        if ((rc = sdr_check_image(ctx, dsts[e], ch, 1, e ? "right_dst" : "left_dst"))) return rc;
        if (dsts[e]->cols != r->W || dsts[e]->rows != r->H)
            return SDR_FAIL(ctx, SDR_ESIZE, "rectify: destination must be %dx%d (calibration size)", r->W, r->H);
        if (srcs[e]->cols != srcs[0]->cols || srcs[e]->rows != srcs[0]->rows) return SDR_FAIL(ctx, SDR_ESIZE, "rectify: source sizes differ");
        if ((rc = sdr_stage_in(ctx, srcs[e], r->in[e], &vi[e]))) return rc;
        if ((rc = sdr_stage_out_begin(ctx, dsts[e], r->out[e], &vo[e]))) return rc;
        sp[e] = vi[e].p; ss[e] = vi[e].pitch; dp[e] = vo[e].p; ds[e] = vo[e].pitch;
        if (grays[e]) {
            if ((rc = sdr_check_image(ctx, grays[e], 1, 1, "gray"))) return rc;
            if (grays[e]->cols != r->W || grays[e]->rows != r->H) return SDR_FAIL(ctx, SDR_ESIZE, "rectify: gray size");
            if ((rc = sdr_stage_out_begin(ctx, grays[e], r->gray[e], &vg[e]))) return rc;
            gp[e] = vg[e].p; gs[e] = vg[e].pitch;
        }
    }
    if ((rc = sdr_launch_rectify(r, 1, ch, sp, ss, sf, srcs[0]->cols, srcs[0]->rows, dp, ds, df, lg ? gp : nullptr, gs, gf))) return rc;
    for (int e = 0; e < 2; e++) {
        if ((rc = sdr_stage_out_end(ctx, dsts[e], vo[e]))) return rc;
        if (grays[e] && (rc = sdr_stage_out_end(ctx, grays[e], vg[e]))) return rc;
    }
    return SDR_OK;
}

extern "C" int sdr_remap_with_maps(sdr_ctx* ctx, const sdr_image* src, const int16_t* map1, const uint16_t* map2, sdr_image* dst) {
    if (!ctx || !src || !dst || !map1 || !map2) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    const int ch = src->channels;
    if (ch != 1 && ch != 3) return SDR_FAIL(ctx, SDR_ESIZE, "remap: 1 or 3 channels expected");
    int rc;
    if ((rc = sdr_check_image(ctx, src, ch, 1, "src"))) return rc;
    if ((rc = sdr_check_image(ctx, dst, ch, 1, "dst"))) return rc;
    DevBuf bin, bout, bmaps;
    DevView vi, vo;
    const size_t n = (size_t)dst->cols * dst->rows;
    rc = bmaps.reserve(ctx, n * 6);
    if (rc == SDR_OK) rc = sdr_stage_in(ctx, src, bin, &vi);
    if (rc == SDR_OK) rc = sdr_stage_out_begin(ctx, dst, bout, &vo);
    if (rc == SDR_OK) {
        int16_t* d1 = bmaps.as<int16_t>();
        uint16_t* d2 = (uint16_t*)(d1 + n * 2);
        cudaMemcpyAsync(d1, map1, n * 4, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d2, map2, n * 2, cudaMemcpyHostToDevice, ctx->stream);
        RectifyArgs a{};
        a.src[0] = vi.p; a.src_step[0] = vi.pitch; a.dst[0] = vo.p; a.dst_step[0] = vo.pitch;
        a.map1[0] = d1; a.map2[0] = d2;
        a.sw = src->cols; a.sh = src->rows; a.W = dst->cols; a.H = dst->rows;
        MapParams dummy{};
        const dim3 grid(div_up(div_up(a.W, 4), 128), a.H, 1);  // z = 1: eye 0 only
        if (ch == 3) k_rectify<3, false><<<grid, 128, 0, ctx->stream>>>(a, dummy, dummy);
        else k_rectify<1, false><<<grid, 128, 0, ctx->stream>>>(a, dummy, dummy);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = SDR_FAIL(ctx, SDR_ECUDA, "remap launch failed");
    }
    if (rc == SDR_OK) rc = sdr_stage_out_end(ctx, dst, vo);
    cudaStreamSynchronize(ctx->stream);
    bin.release(); bout.release(); bmaps.release();
    return rc;
}

extern "C" int sdr_bgr2gray(sdr_ctx* ctx, const sdr_image* bgr, sdr_image* gray) {
    if (!ctx || !bgr || !gray) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, bgr, 3, 1, "bgr"))) return rc;
    if ((rc = sdr_check_image(ctx, gray, 1, 1, "gray"))) return rc;
    if (bgr->cols != gray->cols || bgr->rows != gray->rows) return SDR_FAIL(ctx, SDR_ESIZE, "bgr2gray: size mismatch");
    DevBuf bin, bout;
    DevView vi, vo;
    rc = sdr_stage_in(ctx, bgr, bin, &vi);
    if (rc == SDR_OK) rc = sdr_stage_out_begin(ctx, gray, bout, &vo);
    if (rc == SDR_OK) {
        k_bgr2gray<<<dim3(div_up(bgr->cols, 256), bgr->rows), 256, 0, ctx->stream>>>(vi.p, vi.pitch, vo.p, vo.pitch, bgr->cols, bgr->rows);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = SDR_FAIL(ctx, SDR_ECUDA, "bgr2gray launch failed");
    }
    if (rc == SDR_OK) rc = sdr_stage_out_end(ctx, gray, vo);
    cudaStreamSynchronize(ctx->stream);
    bin.release(); bout.release();
    return rc;
}

// Device-pointer launcher shared with the pipeline. min_key: int[batch] scratch.
int sdr_launch_reproject(sdr_ctx* ctx, bool is16, const uint8_t* disp, size_t step, size_t frame, int W, int H, int batch,
                         const double* Q, int handle_missing, int* min_key, float* out, size_t ostep, size_t oframe) {
    QMat q;
    memcpy(q.q, Q, sizeof q.q);
    const dim3 grid(div_up(W, 256), H, batch);
    if (handle_missing) {
        k_set_int<<<1, 32, 0, ctx->stream>>>(min_key, batch, 0x7FFFFFFF);
        SDR_LAUNCHED(ctx);
        if (is16) k_min_disp<true><<<grid, 256, 0, ctx->stream>>>(disp, step, frame, W, H, min_key);
        else k_min_disp<false><<<grid, 256, 0, ctx->stream>>>(disp, step, frame, W, H, min_key);
        SDR_LAUNCHED(ctx);
    }
    if (is16) k_reproject<true><<<grid, 256, 0, ctx->stream>>>(disp, step, frame, W, H, q, handle_missing, min_key, out, ostep, oframe);
    else k_reproject<false><<<grid, 256, 0, ctx->stream>>>(disp, step, frame, W, H, q, handle_missing, min_key, out, ostep, oframe);
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

extern "C" int sdr_reproject_to_3d(sdr_ctx* ctx, const sdr_image* disp, const double* Q, int handle_missing, sdr_image* points) {
    if (!ctx || !disp || !Q || !points) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const bool is16 = disp->elem_size == 2;
    if ((rc = sdr_check_image(ctx, disp, 1, is16 ? 2 : 4, "disp"))) return rc;
    if ((rc = sdr_check_image(ctx, points, 3, 4, "points"))) return rc;
    if (disp->cols != points->cols || disp->rows != points->rows) return SDR_FAIL(ctx, SDR_ESIZE, "reproject: size mismatch");
    DevBuf bin, bout, bmin;
    DevView vi, vo;
    rc = bmin.reserve(ctx, 256);
    if (rc == SDR_OK) rc = sdr_stage_in(ctx, disp, bin, &vi);
    if (rc == SDR_OK) rc = sdr_stage_out_begin(ctx, points, bout, &vo);
    if (rc == SDR_OK)
        rc = sdr_launch_reproject(ctx, is16, vi.p, vi.pitch, 0, disp->cols, disp->rows, 1, Q, handle_missing, bmin.as<int>(), (float*)vo.p,
                                  vo.pitch, 0);
    if (rc == SDR_OK) rc = sdr_stage_out_end(ctx, points, vo);
    cudaStreamSynchronize(ctx->stream);
    bin.release(); bout.release(); bmin.release();
    return rc;
}

extern "C" int sdr_measure_distance(sdr_ctx* ctx, const sdr_image* points, int x1, int y1, int x2, int y2, double* dist) {
    if (!ctx || !points || !dist) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, points, 3, 4, "points"))) return rc;
    if (x1 < 0 || x1 >= points->cols || y1 < 0 || y1 >= points->rows || x2 < 0 || x2 >= points->cols || y2 < 0 || y2 >= points->rows)
        return SDR_FAIL(ctx, SDR_EINVAL, "measure_distance: point outside the image");
    float v[6];
    if (points->mem == SDR_MEM_DEVICE) {
        DevBuf tmp;
        if ((rc = tmp.reserve(ctx, 64))) return rc;
        k_pick2<<<1, 32, 0, ctx->stream>>>((const uint8_t*)points->data, points->step_bytes, x1, y1, x2, y2, tmp.as<float>());
        ctx->launches++;
        cudaError_t e = cudaMemcpyAsync(v, tmp.p, sizeof v, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        tmp.release();
        if (e != cudaSuccess) return SDR_FAIL(ctx, SDR_ECUDA, "measure_distance: %s", cudaGetErrorString(e));
    } else {
        const float* a = (const float*)((const uint8_t*)points->data + (size_t)y1 * points->step_bytes) + (size_t)x1 * 3;
        const float* b = (const float*)((const uint8_t*)points->data + (size_t)y2 * points->step_bytes) + (size_t)x2 * 3;
        for (int i = 0; i < 3; i++) { v[i] = a[i]; v[3 + i] = b[i]; }
    }
    double s = 0;
    for (int i = 0; i < 3; i++) {
        const float d = v[i] - v[3 + i];   // Vec3f subtraction
        s += (double)d * (double)d;        // cv::norm accumulates squares in double
    }
    *dist = std::sqrt(s);
    return SDR_OK;
}

extern "C" int sdr_depth_coverage(sdr_ctx* ctx, const sdr_image* points, int first_col, float z_min, float z_max, int64_t* count,
                                  double* coverage_percent) {
    if (!ctx || !points) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, points, 3, 4, "points"))) return rc;
    DevBuf bin, bcnt;
    DevView vi;
    unsigned long long n = 0;
    rc = bcnt.reserve(ctx, 64);
    if (rc == SDR_OK) rc = sdr_stage_in(ctx, points, bin, &vi);
    if (rc == SDR_OK) {
        cudaMemsetAsync(bcnt.p, 0, 8, ctx->stream);
        k_coverage<<<dim3(div_up(points->cols, 256), points->rows), 256, 0, ctx->stream>>>(vi.p, vi.pitch, points->cols, points->rows, first_col, z_min,
                                                                                           z_max, bcnt.as<unsigned long long>());
        ctx->launches++;
        cudaError_t e = cudaMemcpyAsync(&n, bcnt.p, 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = SDR_FAIL(ctx, SDR_ECUDA, "depth_coverage: %s", cudaGetErrorString(e));
    }
    bin.release(); bcnt.release();
    if (rc) return rc;
    if (count) *count = (int64_t)n;
    if (coverage_percent) *coverage_percent = ((double)n / ((double)points->rows * points->cols)) * 100;
    return SDR_OK;
}

extern "C" int sdr_disp_to_float(sdr_ctx* ctx, const sdr_image* disp16, sdr_image* disp32f) {
    if (!ctx || !disp16 || !disp32f) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, disp16, 1, 2, "disp16"))) return rc;
    if ((rc = sdr_check_image(ctx, disp32f, 1, 4, "disp32f"))) return rc;
    if (disp16->cols != disp32f->cols || disp16->rows != disp32f->rows) return SDR_FAIL(ctx, SDR_ESIZE, "disp_to_float: size mismatch");
    DevBuf bin, bout;
    DevView vi, vo;
    rc = sdr_stage_in(ctx, disp16, bin, &vi);
    if (rc == SDR_OK) rc = sdr_stage_out_begin(ctx, disp32f, bout, &vo);
    if (rc == SDR_OK) {
        k_disp_to_float<<<dim3(div_up(disp16->cols, 256), disp16->rows), 256, 0, ctx->stream>>>(vi.p, vi.pitch, vo.p, vo.pitch, disp16->cols,
                                                                                                disp16->rows);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = SDR_FAIL(ctx, SDR_ECUDA, "disp_to_float launch failed");
    }
    if (rc == SDR_OK) rc = sdr_stage_out_end(ctx, disp32f, vo);
    sdr_wait(ctx);
    bin.release(); bout.release();
    return rc;
}
