// Frame ingest: the colour conversion between a hardware video decoder / a UVC camera and the
// BGR side-by-side frame the reference's loops start from. The reference gets its frames from cv::VideoCapture
// (stereo_vision/src/stereo_displayer.cpp:132-156, point_cloud/src/pcd_write.cpp:59-84), which decodes on the CPU and hands
// out BGR; NVDEC delivers NV12 surfaces in device memory and the ZED 2 streams YUYV 4:2:2 over UVC, so on this path the
// frame stays in its native 12 / 16 bit-per-pixel format across PCIe (half / two thirds of the BGR bytes) and becomes BGR
// here. Arithmetic = cv::cvtColor(COLOR_YUV2BGR_NV12 / COLOR_YUV2BGR_YUY2) of OpenCV 4.13 (ITU-R BT.601, 20-bit fixed
// point), pinned exhaustively over all (Y, U, V) against cv2 (tests/test_oracle_vs_cv2.py):
//   y = max(Y - 16, 0) * 1220542;  u = U - 128;  v = V - 128
//   B = sat8((y + 2116026 u + 2^19) >> 20);  G = sat8((y - 409993 u - 852492 v + 2^19) >> 20);  R = sat8((y + 1673527 v + 2^19) >> 20)
// The bitstream decode itself (cuvid) is not built: the image has neither the Video Codec SDK headers nor a test stream.
#include "sdr_internal.h"

namespace {

__device__ __forceinline__ uint32_t sat8(int v) { return (uint32_t)min(max(v, 0), 255); }

struct Chroma {
    int r, g, b;   // the chroma terms incl. the rounding constant, shared by the two pixels of a pair
};
__device__ __forceinline__ Chroma chroma(int U, int V) {
    const int u = U - 128, v = V - 128;
    Chroma c;
    c.r = (1 << 19) + 1673527 * v;
    c.g = (1 << 19) - 852492 * v - 409993 * u;
    c.b = (1 << 19) + 2116026 * u;
    return c;
}
// -> 0x00RRGGBB with B in the low byte (memory order B, G, R)
__device__ __forceinline__ uint32_t bgr_px(int Y, const Chroma& c) {
    const int y = max(Y - 16, 0) * 1220542;
    return sat8((y + c.b) >> 20) | (sat8((y + c.g) >> 20) << 8) | (sat8((y + c.r) >> 20) << 16);
}
// four pixels (24-bit each) -> three little-endian words of the B,G,R byte stream
__device__ __forceinline__ void store4(uint8_t* dst, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
    uint32_t* d = (uint32_t*)dst;
    d[0] = p0 | (p1 << 24);
    d[1] = (p1 >> 8) | (p2 << 16);
    d[2] = (p2 >> 16) | (p3 << 8);
}

// NV12: plane Y [H][W] followed by interleaved UV [H/2][W/2][2] with the same row step. A thread converts 4 pixels of one
// row: one word of Y, one word of UV (two chroma pairs), three words of BGR. W % 4 == 0 on this path.
__global__ void __launch_bounds__(256) k_ingest_nv12(const uint8_t* __restrict__ src, size_t sstep, size_t sframe, uint8_t* __restrict__ dst,
                                                     size_t dstep, size_t dframe, int W, int H) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (q * 4 >= W) return;
    src += (size_t)f * sframe;
    const uint32_t yy = *(const uint32_t*)(src + (size_t)y * sstep + 4 * q);
    const uint32_t uv = *(const uint32_t*)(src + (size_t)(H + (y >> 1)) * sstep + 4 * q);
    const Chroma c0 = chroma(uv & 255, (uv >> 8) & 255), c1 = chroma((uv >> 16) & 255, uv >> 24);
    store4(dst + (size_t)f * dframe + (size_t)y * dstep + 12 * q, bgr_px(yy & 255, c0), bgr_px((yy >> 8) & 255, c0), bgr_px((yy >> 16) & 255, c1),
           bgr_px(yy >> 24, c1));
}

// YUYV (YUY2): [H][W/2][Y0 U Y1 V]. A thread converts 4 pixels: two words in, three words out.
__global__ void __launch_bounds__(256) k_ingest_yuyv(const uint8_t* __restrict__ src, size_t sstep, size_t sframe, uint8_t* __restrict__ dst,
                                                     size_t dstep, size_t dframe, int W, int H) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (q * 4 >= W) return;
    const uint2 w = *(const uint2*)(src + (size_t)f * sframe + (size_t)y * sstep + 8 * q);
    const Chroma c0 = chroma((w.x >> 8) & 255, w.x >> 24), c1 = chroma((w.y >> 8) & 255, w.y >> 24);
    store4(dst + (size_t)f * dframe + (size_t)y * dstep + 12 * q, bgr_px(w.x & 255, c0), bgr_px((w.x >> 16) & 255, c0), bgr_px(w.y & 255, c1),
           bgr_px((w.y >> 16) & 255, c1));
}

// any even width, unaligned steps: one thread per pixel pair, byte accesses
__global__ void __launch_bounds__(256) k_ingest_generic(int fmt, const uint8_t* __restrict__ src, size_t sstep, size_t sframe,
                                                        uint8_t* __restrict__ dst, size_t dstep, size_t dframe, int W, int H) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (p * 2 >= W) return;
    src += (size_t)f * sframe;
    int Y0, Y1, U, V;
    if (fmt == SDR_PIX_NV12) {
        const uint8_t* yr = src + (size_t)y * sstep + 2 * p;
        const uint8_t* cr = src + (size_t)(H + (y >> 1)) * sstep + 2 * p;
        Y0 = yr[0]; Y1 = yr[1]; U = cr[0]; V = cr[1];
    } else {
        const uint8_t* r = src + (size_t)y * sstep + 4 * p;
        Y0 = r[0]; U = r[1]; Y1 = r[2]; V = r[3];
    }
    const Chroma c = chroma(U, V);
    const uint32_t a = bgr_px(Y0, c), b = bgr_px(Y1, c);
    uint8_t* d = dst + (size_t)f * dframe + (size_t)y * dstep + 6 * p;
    d[0] = (uint8_t)a; d[1] = (uint8_t)(a >> 8); d[2] = (uint8_t)(a >> 16);
    d[3] = (uint8_t)b; d[4] = (uint8_t)(b >> 8); d[5] = (uint8_t)(b >> 16);
}

}  // namespace

// Rows and bytes per row of one frame of `cols` x `rows` pixels as it arrives in format fmt.
void sdr_ingest_layout(int fmt, int cols, int rows, int* src_rows, size_t* row_bytes) {
    *src_rows = fmt == SDR_PIX_NV12 ? rows + rows / 2 : rows;
    *row_bytes = fmt == SDR_PIX_BGR ? (size_t)cols * 3 : fmt == SDR_PIX_NV12 ? (size_t)cols : (size_t)cols * 2;
}

// `batch` frames, device pointers. W, H: pixels of the decoded picture.
int sdr_launch_ingest(sdr_ctx* ctx, int fmt, const uint8_t* src, size_t sstep, size_t sframe, uint8_t* dst, size_t dstep, size_t dframe, int W,
                      int H, int batch) {
    if (fmt != SDR_PIX_NV12 && fmt != SDR_PIX_YUYV) return SDR_FAIL(ctx, SDR_EINVAL, "ingest: unknown pixel format %d", fmt);
    if ((W & 1) || (fmt == SDR_PIX_NV12 && (H & 1)))
        return SDR_FAIL(ctx, SDR_EINVAL, "ingest: %s needs an even width%s (got %dx%d)", fmt == SDR_PIX_NV12 ? "NV12" : "YUYV",
                        fmt == SDR_PIX_NV12 ? " and height" : "", W, H);   // cv::cvtColor asserts the same
    const bool words = W % 4 == 0 && sstep % 8 == 0 && sframe % 8 == 0 && ((uintptr_t)src & 7) == 0 && dstep % 4 == 0 && dframe % 4 == 0 &&
                       ((uintptr_t)dst & 3) == 0;
    if (words) {
        const dim3 grid(div_up(W / 4, 256), H, batch);
        if (fmt == SDR_PIX_NV12) k_ingest_nv12<<<grid, 256, 0, ctx->stream>>>(src, sstep, sframe, dst, dstep, dframe, W, H);
        else k_ingest_yuyv<<<grid, 256, 0, ctx->stream>>>(src, sstep, sframe, dst, dstep, dframe, W, H);
    } else {
        k_ingest_generic<<<dim3(div_up(W / 2, 256), H, batch), 256, 0, ctx->stream>>>(fmt, src, sstep, sframe, dst, dstep, dframe, W, H);
    }
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

extern "C" int sdr_ingest_to_bgr(sdr_ctx* ctx, int format, const sdr_image* src, sdr_image* bgr) {
    if (!ctx || !src || !bgr) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (format != SDR_PIX_NV12 && format != SDR_PIX_YUYV) return SDR_FAIL(ctx, SDR_EINVAL, "ingest: pixel format must be SDR_PIX_NV12 or SDR_PIX_YUYV");
    if ((rc = sdr_check_image(ctx, src, format == SDR_PIX_NV12 ? 1 : 2, 1, "src")) || (rc = sdr_check_image(ctx, bgr, 3, 1, "bgr"))) return rc;
    const int W = bgr->cols, H = bgr->rows;
    int srows;
    size_t row_bytes;
    sdr_ingest_layout(format, W, H, &srows, &row_bytes);
    if ((W & 1) || (format == SDR_PIX_NV12 && (H & 1)))
        return SDR_FAIL(ctx, SDR_EINVAL, "ingest: %s needs an even width%s (got %dx%d)", format == SDR_PIX_NV12 ? "NV12" : "YUYV",
                        format == SDR_PIX_NV12 ? " and height" : "", W, H);
    if (src->cols != W || src->rows != srows)
        return SDR_FAIL(ctx, SDR_ESIZE, "ingest: a %dx%d picture arrives as %d rows x %d cols of %d channel(s), got %d x %d", W, H, srows, W,
                        format == SDR_PIX_NV12 ? 1 : 2, src->rows, src->cols);
    DevBuf bin, bout;
    DevView vi, vo;
    rc = sdr_stage_in(ctx, src, bin, &vi);
    if (rc == SDR_OK) rc = sdr_stage_out_begin(ctx, bgr, bout, &vo);
    if (rc == SDR_OK) rc = sdr_launch_ingest(ctx, format, vi.p, vi.pitch, 0, vo.p, vo.pitch, 0, W, H, 1);
    if (rc == SDR_OK) rc = sdr_stage_out_end(ctx, bgr, vo);
    sdr_wait(ctx);   // the scratch buffers are released below
    bin.release(); bout.release();
    return rc;
}
