// Frame-parallel stream runner: one worker thread per (device, handle), a shared atomic cursor over the frame
// indices, no inter-GPU communication (frames share no state). Host code only.
#include "sdr_internal.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

extern "C" int sdr_stream_run(const int* devices, int n_devices, int handles_per_device, const sdr_calibration* calib,
                              const sdr_pipeline_options* opt, int64_t n_frames, sdr_frame_source source, sdr_frame_sink sink, void* user,
                              sdr_stream_stats* stats) {
    if (!devices || n_devices <= 0 || n_devices > 16 || handles_per_device <= 0 || !calib || !opt || !source || n_frames < 0) return SDR_EINVAL;
    const int B = opt->batch;
    if (B < 1) return SDR_EINVAL;
    const int W = calib->image_width, H = calib->image_height;
    std::atomic<int64_t> cursor{0};
    std::atomic<int> first_error{SDR_OK};
    std::vector<std::atomic<int64_t>> done(n_devices);
    for (auto& d : done) d = 0;
    std::mutex err_mu;
    std::string err_msg;
    // start line: no frame is handed out before every worker has its context and pipeline (so the frames spread over
    // all devices and the clock does not include context creation)
    std::mutex start_mu;
    std::condition_variable start_cv;
    int ready = 0;
    const int n_workers = n_devices * handles_per_device;
    std::chrono::steady_clock::time_point t0;
    auto worker = [&](int di) {
        sdr_ctx* ctx = nullptr;
        sdr_pipeline* pipe = nullptr;
        int rc = sdr_create(devices[di], &ctx);
        if (rc == SDR_OK) rc = sdr_pipeline_create(ctx, calib, opt, &pipe);
        PinnedBuf disp_host;
        if (rc == SDR_OK) rc = disp_host.reserve(ctx, (size_t)B * W * H * 2);
        std::vector<sdr_image> sbs(B), disp(B);
        std::vector<sdr_cloud_view> views(B);
        std::vector<int64_t> idx(B);
        {
            std::unique_lock<std::mutex> lk(start_mu);
            if (++ready == n_workers) {
                t0 = std::chrono::steady_clock::now();
                start_cv.notify_all();
            } else {
                start_cv.wait(lk, [&] { return ready == n_workers; });
            }
        }
        while (rc == SDR_OK && first_error.load() == SDR_OK) {
            const int64_t start = cursor.fetch_add(B);
            if (start >= n_frames) break;
            int n = 0;
            for (int64_t i = start; i < std::min<int64_t>(start + B, n_frames); i++) {
                sdr_image im{};
                if (source(user, i, &im) != SDR_OK) continue;   // dropped frame
                sbs[n] = im;
                disp[n] = sdr_image{(uint8_t*)disp_host.p + (size_t)n * W * H * 2, H, W, 1, 2, (size_t)W * 2, SDR_MEM_PINNED};
                idx[n++] = i;
            }
            if (n == 0) continue;
            rc = sdr_pipeline_run(pipe, n, sbs.data(), disp.data(), nullptr, nullptr, views.data());
            if (rc == SDR_EVOXEL_OVERFLOW) rc = SDR_OK;   // non-fatal: the views say which clouds were passed through
            if (rc != SDR_OK) break;
            if (sink)
                for (int k = 0; k < n; k++) sink(user, idx[k], devices[di], &disp[k], &views[k]);
            done[di] += n;
        }
        if (rc != SDR_OK) {
            int expected = SDR_OK;
            if (first_error.compare_exchange_strong(expected, rc)) {
                std::lock_guard<std::mutex> lk(err_mu);
                err_msg = ctx ? sdr_last_error(ctx) : "sdr_create failed";
            }
        }
        disp_host.release();
        if (pipe) sdr_pipeline_destroy(pipe);
        if (ctx) sdr_destroy(ctx);
    };
    std::vector<std::thread> threads;
    for (int di = 0; di < n_devices; di++)
        for (int h = 0; h < handles_per_device; h++) threads.emplace_back(worker, di);
    for (auto& t : threads) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        *stats = sdr_stream_stats{};
        for (int di = 0; di < n_devices; di++) {
            stats->per_device[di] = done[di];
            stats->frames += done[di];
        }
        stats->seconds = secs;
        stats->frames_per_s = secs > 0 ? stats->frames / secs : 0;
    }
    if (first_error.load() != SDR_OK) fprintf(stderr, "sdr_stream_run: %s\n", err_msg.c_str());
    return first_error.load();
}
