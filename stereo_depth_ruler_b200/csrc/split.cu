// Row-strip split of ONE oversize frame across the GPUs of a box, native and in one process (BASELINE config 5,
// SURVEY.md 8(e)). MODE_SGBM_3WAY is four independent stripes by construction (OpenCV's computeDisparity3WaySGBM behind
// the reference's matcher, stereo_disparity.cpp:5-6,27): stripe i reads source rows [src_start, row_end) and owns output
// rows [dst_start, row_end). Each device matches a contiguous block of stripes from just the input rows it needs
// (source rows +/- the Sobel / box-window halo), pushes its raw strip into device 0's assembled map over NVLink
// (cudaMemcpyPeerAsync on its own stream, an event hands it to device 0's stream), and device 0 runs the 3x3 median and
// the speckle filter -- the only stages that look across stripe borders -- once. No NCCL: the only exchange is the input
// halo and the 4 MB strips, plain peer copies. Everything is enqueued from the calling thread; nothing blocks until the
// final wait.
#include "sdr_internal.h"

#include <algorithm>
#include <cstring>

namespace {
struct Part {
    int device = 0;
    sdr_ctx* ctx = nullptr;
    sdr_sgbm* m = nullptr;
    DevBuf L, R, raw;
    cudaEvent_t done = nullptr;
    int first = 0, count = 0;   // stripes of this device
    bool peer = false;          // direct peer access to device 0
};
}  // namespace

struct sdr_sgbm_split {
    sdr_sgbm_params params{};
    std::vector<Part> parts;
    cudaEvent_t ev_in = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    sdr_split_stats stats{};
    std::string err;
};

static int split_fail(sdr_sgbm_split* s, int code, const char* what, sdr_ctx* ctx = nullptr) {
    s->err = what;
    if (ctx) { s->err += ": "; s->err += sdr_last_error(ctx); }
    return code;
}

extern "C" int sdr_sgbm_split_create(const int* devices, int n_devices, const sdr_sgbm_params* params, sdr_sgbm_split** out) {
    if (!devices || n_devices <= 0 || n_devices > 16 || !params || !out) return SDR_EINVAL;
    *out = nullptr;
    sdr_sgbm_split* s = new sdr_sgbm_split();
    s->params = *params;
    s->parts.resize(n_devices);
    int rc = SDR_OK;
    for (int k = 0; k < n_devices && rc == SDR_OK; k++) {
        Part& p = s->parts[k];
        p.device = devices[k];
        rc = sdr_create(p.device, &p.ctx);
        if (rc == SDR_OK) rc = sdr_sgbm_create(p.ctx, params, &p.m);
        // several parts may sit on one device (a split over fewer GPUs than stripes is also a valid use): tell the matchers
        for (int j = 0; j < k && rc == SDR_OK; j++)
            if (s->parts[j].device == p.device) {
                sdr_sgbm_set_option(p.m, "shared_device", 1);
                sdr_sgbm_set_option(s->parts[j].m, "shared_device", 1);
            }
        if (rc == SDR_OK && cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming) != cudaSuccess) rc = SDR_ECUDA;
    }
    if (rc == SDR_OK) {
        const int d0 = s->parts[0].device;
        for (int k = 0; k < n_devices; k++) {
            Part& p = s->parts[k];
            if (p.device == d0) { p.peer = true; continue; }
            int can01 = 0, can10 = 0;
            cudaDeviceCanAccessPeer(&can01, d0, p.device);
            cudaDeviceCanAccessPeer(&can10, p.device, d0);
            if (can01 && can10) {
                cudaSetDevice(d0);
                cudaError_t e = cudaDeviceEnablePeerAccess(p.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                cudaSetDevice(p.device);
                cudaError_t f = cudaDeviceEnablePeerAccess(d0, 0);
                if (f == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); f = cudaSuccess; }
                p.peer = e == cudaSuccess && f == cudaSuccess;
            }
            // without peer access cudaMemcpyPeerAsync still works (staged by the driver); stats say which it was
        }
        cudaSetDevice(d0);
        if (cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&s->ev_t0) != cudaSuccess ||
            cudaEventCreate(&s->ev_t1) != cudaSuccess)
            rc = SDR_ECUDA;
    }
    if (rc != SDR_OK) {
        sdr_sgbm_split_destroy(s);
        return rc;
    }
    *out = s;
    return SDR_OK;
}

extern "C" void sdr_sgbm_split_destroy(sdr_sgbm_split* s) {
    if (!s) return;
    for (Part& p : s->parts) {
        if (!p.ctx) continue;
        cudaSetDevice(p.device);
        cudaStreamSynchronize(p.ctx->stream);
        p.L.release(); p.R.release(); p.raw.release();
        if (p.done) cudaEventDestroy(p.done);
        if (p.m) sdr_sgbm_destroy(p.m);
        sdr_destroy(p.ctx);
    }
    if (s->ev_in) cudaEventDestroy(s->ev_in);
    if (s->ev_t0) cudaEventDestroy(s->ev_t0);
    if (s->ev_t1) cudaEventDestroy(s->ev_t1);
    delete s;
}

extern "C" const char* sdr_sgbm_split_last_error(sdr_sgbm_split* s) { return s ? s->err.c_str() : "no handle"; }

extern "C" int sdr_sgbm_split_compute(sdr_sgbm_split* s, const sdr_image* left, const sdr_image* right, sdr_image* disp) {
    if (!s || !left || !right || !disp) return SDR_EINVAL;
    Part& p0 = s->parts[0];
    sdr_ctx* ctx0 = p0.ctx;
    int rc;
    if ((rc = sdr_check_image(ctx0, left, 1, 1, "left")) || (rc = sdr_check_image(ctx0, right, 1, 1, "right")) ||
        (rc = sdr_check_image(ctx0, disp, 1, 2, "disp")))
        return split_fail(s, rc, "split_compute", ctx0);
    const int W = left->cols, H = left->rows;
    if (right->cols != W || right->rows != H || disp->cols != W || disp->rows != H) return split_fail(s, SDR_ESIZE, "split_compute: sizes differ");
    const bool dev_in = left->mem == SDR_MEM_DEVICE;
    if ((right->mem == SDR_MEM_DEVICE) != dev_in) return split_fail(s, SDR_EINVAL, "split_compute: left and right must live in the same memory kind");
    if (dev_in && (left->step_bytes != (size_t)W || right->step_bytes != (size_t)W))
        return split_fail(s, SDR_EINVAL, "split_compute: device-resident inputs must be dense (step == cols)");
    // stripes of this size, contiguous blocks per part (parts beyond the number of stripes idle)
    const int nstripes = sdr_sgbm_stripe_rows(&s->params, W, H, -1, nullptr, nullptr, nullptr);
    if (nstripes <= 0) return split_fail(s, SDR_EINVAL, "split_compute: bad geometry");
    const int nparts = (int)s->parts.size();
    const int SH2 = (s->params.mode == SDR_MODE_SGBM_3WAY ? (s->params.blockSize > 0 ? s->params.blockSize / 2 : 1)
                                                          : (s->params.blockSize > 0 ? s->params.blockSize : 5) / 2);
    {
        const int base = nstripes / nparts, extra = nstripes % nparts;
        int first = 0;
        for (int k = 0; k < nparts; k++) {
            s->parts[k].first = first;
            s->parts[k].count = base + (k < extra ? 1 : 0);
            first += s->parts[k].count;
        }
    }
    sdr_split_stats& st = s->stats;
    st = sdr_split_stats{};
    st.n_parts = nparts;
    st.n_stripes = nstripes;
    const size_t img = (size_t)W * H;
    cudaSetDevice(p0.device);
    if ((rc = p0.raw.reserve(ctx0, img * 2))) return split_fail(s, rc, "split_compute", ctx0);
    if (cudaEventRecord(s->ev_t0, ctx0->stream) != cudaSuccess || cudaEventRecord(s->ev_in, ctx0->stream) != cudaSuccess)
        return split_fail(s, SDR_ECUDA, "split_compute: event record failed");
    sdr_image raw0{p0.raw.p, H, W, 1, 2, (size_t)W * 2, SDR_MEM_DEVICE};
    // parts other than 0 first, so that their work is in flight while device 0's own launches are being queued
    for (int kk = 0; kk < nparts; kk++) {
        const int k = (kk + 1) % nparts;   // 1, 2, ..., n-1, 0
        Part& p = s->parts[k];
        st.stripes[k] = p.count;
        st.device[k] = p.device;
        st.peer_access[k] = p.peer ? 1 : 0;
        if (p.count == 0) continue;
        int src0 = 0, dst0 = 0, end0 = 0, src1 = 0, dst1 = 0, end1 = 0;
        sdr_sgbm_stripe_rows(&s->params, W, H, p.first, &src0, &dst0, &end0);
        sdr_sgbm_stripe_rows(&s->params, W, H, p.first + p.count - 1, &src1, &dst1, &end1);
        const int r0 = std::max(src0 - 1, 0), r1 = std::min(end1 + SH2 + 1, H);   // input rows this part reads
        cudaSetDevice(p.device);
        cudaStream_t stp = p.ctx->stream;
        sdr_image Lk, Rk;
        const bool in_place = dev_in && p.device == p0.device;
        if (in_place) {
            Lk = *left; Rk = *right;
            if (k != 0 && cudaStreamWaitEvent(stp, s->ev_in, 0) != cudaSuccess) return split_fail(s, SDR_ECUDA, "split_compute: wait failed");
        } else {
            if ((rc = p.L.reserve(p.ctx, img)) || (rc = p.R.reserve(p.ctx, img))) return split_fail(s, rc, "split_compute", p.ctx);
            const size_t off = (size_t)r0 * W, rows_b = (size_t)(r1 - r0) * W;
            cudaError_t e;
            if (dev_in) {   // halo + body rows over NVLink from device 0, after whatever produced them on its stream
                e = cudaStreamWaitEvent(stp, s->ev_in, 0);
                if (e == cudaSuccess) e = cudaMemcpyPeerAsync((uint8_t*)p.L.p + off, p.device, (const uint8_t*)left->data + off, p0.device, rows_b, stp);
                if (e == cudaSuccess) e = cudaMemcpyPeerAsync((uint8_t*)p.R.p + off, p.device, (const uint8_t*)right->data + off, p0.device, rows_b, stp);
            } else {        // host frame: every device uploads just its own (overlapping) row range
                e = cudaMemcpy2DAsync((uint8_t*)p.L.p + off, W, (const uint8_t*)left->data + (size_t)r0 * left->step_bytes, left->step_bytes, W,
                                      r1 - r0, cudaMemcpyHostToDevice, stp);
                if (e == cudaSuccess)
                    e = cudaMemcpy2DAsync((uint8_t*)p.R.p + off, W, (const uint8_t*)right->data + (size_t)r0 * right->step_bytes,
                                          right->step_bytes, W, r1 - r0, cudaMemcpyHostToDevice, stp);
            }
            if (e != cudaSuccess) return split_fail(s, SDR_ECUDA, cudaGetErrorString(e));
            st.input_bytes[k] = 2 * (int64_t)rows_b;
            Lk = sdr_image{p.L.p, H, W, 1, 1, (size_t)W, SDR_MEM_DEVICE};
            Rk = sdr_image{p.R.p, H, W, 1, 1, (size_t)W, SDR_MEM_DEVICE};
        }
        st.input_rows[k] = r1 - r0;
        if (p.device == p0.device) {
            // same device: the strip goes straight into the assembled map (other streams of device 0 are ordered by p.done)
            if ((rc = sdr_sgbm_match_stripes(p.m, &Lk, &Rk, p.first, p.count, &raw0))) return split_fail(s, rc, "split_compute: match", p.ctx);
        } else {
            if ((rc = p.raw.reserve(p.ctx, img * 2))) return split_fail(s, rc, "split_compute", p.ctx);
            sdr_image rawk{p.raw.p, H, W, 1, 2, (size_t)W * 2, SDR_MEM_DEVICE};
            if ((rc = sdr_sgbm_match_stripes(p.m, &Lk, &Rk, p.first, p.count, &rawk))) return split_fail(s, rc, "split_compute: match", p.ctx);
            cudaSetDevice(p.device);
            const size_t off = (size_t)dst0 * W * 2, bytes = (size_t)(end1 - dst0) * W * 2;
            if (cudaMemcpyPeerAsync((uint8_t*)p0.raw.p + off, p0.device, (const uint8_t*)p.raw.p + off, p.device, bytes, stp) != cudaSuccess)
                return split_fail(s, SDR_ECUDA, "split_compute: peer copy of the raw strip failed");
            st.gather_bytes[k] = (int64_t)bytes;
        }
        if (k != 0) {
            cudaSetDevice(p.device);
            if (cudaEventRecord(p.done, stp) != cudaSuccess) return split_fail(s, SDR_ECUDA, "split_compute: event record failed");
        }
    }
    cudaSetDevice(p0.device);
    for (int k = 1; k < nparts; k++)
        if (s->parts[k].count && cudaStreamWaitEvent(ctx0->stream, s->parts[k].done, 0) != cudaSuccess)
            return split_fail(s, SDR_ECUDA, "split_compute: wait failed");
    if ((rc = sdr_sgbm_postfilter(p0.m, &raw0, disp))) return split_fail(s, rc, "split_compute: postfilter", ctx0);
    if (cudaEventRecord(s->ev_t1, ctx0->stream) != cudaSuccess) return split_fail(s, SDR_ECUDA, "split_compute: event record failed");
    if (disp->mem != SDR_MEM_DEVICE) {   // host output: postfilter has waited already; the elapsed time is final
        cudaEventSynchronize(s->ev_t1);
    }
    return SDR_OK;
}

extern "C" int sdr_sgbm_split_synchronize(sdr_sgbm_split* s) {
    if (!s) return SDR_EINVAL;
    cudaSetDevice(s->parts[0].device);
    return cudaEventSynchronize(s->ev_t1) == cudaSuccess ? SDR_OK : SDR_ECUDA;
}

extern "C" int sdr_sgbm_split_last_stats(sdr_sgbm_split* s, sdr_split_stats* out) {
    if (!s || !out) return SDR_EINVAL;
    cudaSetDevice(s->parts[0].device);
    if (cudaEventSynchronize(s->ev_t1) != cudaSuccess) return SDR_ECUDA;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1) != cudaSuccess) return SDR_ECUDA;
    s->stats.ms = ms;
    *out = s->stats;
    return SDR_OK;
}
