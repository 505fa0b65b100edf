// Internal declarations shared by the translation units of libsdr_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "sdr_c.h"

struct sdr_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaEvent_t wait_ev = nullptr;   // cudaEventBlockingSync: host waits sleep instead of spinning
    unsigned long long peer_ok = 0;  // bit d: peer access from this device to device d is enabled (images that live on another GPU)
};

// Host wait for everything queued on the ctx stream. Sleeps on a blocking-sync event: with several handles per GPU and
// one process per GPU the waiting threads would otherwise spin on every core of the box.
cudaError_t sdr_wait(sdr_ctx* ctx);

#define SDR_FAIL(ctx, code, ...) (sdr_set_error((ctx), __VA_ARGS__), (int)(code))
void sdr_set_error(sdr_ctx* ctx, const char* fmt, ...);

#define SDR_CUDA(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return SDR_FAIL(ctx, SDR_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                            __FILE__, __LINE__);                                                     \
    } while (0)

// Checks the launch and counts it.
#define SDR_LAUNCHED(ctx)                                                                            \
    do {                                                                                             \
        (ctx)->launches++;                                                                           \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess)                                                                       \
            return SDR_FAIL(ctx, SDR_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                     \
    } while (0)

// Growable device buffer owned by a handle (allocated at first use, reused across frames).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(sdr_ctx* ctx, size_t bytes);
    void release();
    template <class T> T* as() const { return (T*)p; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(sdr_ctx* ctx, size_t bytes);
    void release();
};

// A device-side 2-D view. pitch in bytes.
struct DevView {
    uint8_t* p = nullptr;
    size_t pitch = 0;
};

int sdr_check_image(sdr_ctx* ctx, const sdr_image* im, int channels, int elem_size, const char* what);
// Input: returns a device view of `im` (copying host memory into `scratch` on the ctx stream when needed).
int sdr_stage_in(sdr_ctx* ctx, const sdr_image* im, DevBuf& scratch, DevView* out);
// Output: device view to write into; for host images it points into `scratch` and sdr_stage_out copies back.
int sdr_stage_out_begin(sdr_ctx* ctx, const sdr_image* im, DevBuf& scratch, DevView* out);
int sdr_stage_out_end(sdr_ctx* ctx, const sdr_image* im, const DevView& v);

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---- kernels exposed across translation units (host launchers) ----
int sdr_launch_median3(sdr_ctx* ctx, const int16_t* src, size_t spitch, int16_t* dst, size_t dpitch, int W, int H,
                       int batch, size_t sframe, size_t dframe);

struct SpeckleScratch {
    DevBuf parent, count;
};
int sdr_launch_speckle(sdr_ctx* ctx, int16_t* img, size_t pitch, int W, int H, int batch, size_t frame_bytes,
                       int new_val, int max_size, int max_diff, SpeckleScratch& s);

struct StageTimer {
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    bool enabled = false;
    size_t n = 0;
    void begin(cudaStream_t s);
    void mark(cudaStream_t s, const char* name);
    int collect(int cap, const char** out_names, float* ms);
    int collect_since(cudaEvent_t base, int cap, const char** out_names, float* ms);
    void release();
};

// ---- device-pointer entry points used by the per-frame pipeline ----
int sdr_sgbm_run_device(sdr_sgbm* m, int batch, const uint8_t* L, size_t lpitch, size_t lframe, const uint8_t* R, size_t rpitch,
                        size_t rframe, int W, int H, int16_t* out, size_t opitch, size_t oframe);
int sdr_launch_rectify(sdr_rectifier* r, int batch, int channels, const uint8_t* const src[2], const size_t src_step[2],
                       const size_t src_frame[2], int sw, int sh, uint8_t* const dst[2], const size_t dst_step[2],
                       const size_t dst_frame[2], uint8_t* const gray[2], const size_t gray_step[2], const size_t gray_frame[2]);
int sdr_launch_reproject(sdr_ctx* ctx, bool is16, const uint8_t* disp, size_t step, size_t frame, int W, int H, int batch,
                         const double* Q, int handle_missing, int* min_key, float* out, size_t ostep, size_t oframe);
int sdr_cloud_enqueue(sdr_cloud* c, int frames, const uint8_t* points, size_t pstep, size_t pframe, const uint8_t* disp, size_t dstep,
                      size_t dframe, const uint8_t* bgr, size_t cstep, size_t cframe, int W, int H, const double* Q, int handle_missing,
                      const sdr_voxel_options* opt);
int sdr_cloud_finish(sdr_cloud* c, sdr_cloud_view* views);
int sdr_cloud_download_frame(sdr_cloud* c, int f, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt, bool wait);
int sdr_cloud_download_packed(sdr_cloud* c, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt, int64_t* offsets, bool wait);
StageTimer* sdr_sgbm_timer(sdr_sgbm* m);
// frame ingest (ingest.cu): rows / bytes per row of a cols x rows picture in pixel format fmt; batched conversion to BGR
void sdr_ingest_layout(int fmt, int cols, int rows, int* src_rows, size_t* row_bytes);
int sdr_launch_ingest(sdr_ctx* ctx, int fmt, const uint8_t* src, size_t sstep, size_t sframe, uint8_t* dst, size_t dstep, size_t dframe, int W,
                      int H, int batch);
// after a host wait on the ctx stream: SDR_ECUDA (once) if a cooperative sweep of the finished work timed out on a halo
int sdr_sgbm_take_sweep_error(sdr_sgbm* m);
