// Frame-parallel stream runner: one worker thread per (device, handle), a shared atomic cursor over the frame
// indices, no inter-GPU communication (frames share no state). Host code only.
//
// The runner is persistent (sdr_stream_create / sdr_stream_process / sdr_stream_destroy): contexts, pipelines, device
// scratch and pinned result buffers are created once and reused by every process() call, the way the reference reuses
// its matcher and maps across the frames of the video loop (stereo_displayer.cpp:145-198). sdr_stream_run is the
// one-shot form. Worker threads are bound to the CPUs of their GPU's NUMA node before they allocate pinned memory, so
// the host side of every copy stays on the socket the GPU hangs off.
#include "sdr_internal.h"

#include <sched.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <thread>

namespace {

// "0-15,32-47" -> cpu_set_t; returns the number of CPUs set
int parse_cpulist(const std::string& s, cpu_set_t* set) {
    CPU_ZERO(set);
    int n = 0;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        int a = -1, b = -1;
        if (sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) {
        } else if (sscanf(tok.c_str(), "%d", &a) == 1) {
            b = a;
        } else {
            continue;
        }
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) {
            if (c >= 0) { CPU_SET(c, set); n++; }
        }
    }
    return n;
}

}  // namespace

// Binds the calling thread to the CPUs of the NUMA node the device is attached to (intersected with the CPUs the
// process may use). Returns the node, or -1 when the topology is not visible (containers often hide it) -- then nothing
// changes. Threads created afterwards inherit the mask; pinned allocations made afterwards land on that node.
extern "C" int sdr_bind_host_to_device(int device) {
    char busid[32] = {0};
    if (cudaDeviceGetPCIBusId(busid, sizeof busid, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = busid; *c; c++) *c = (char)tolower(*c);
    std::ifstream fn(std::string("/sys/bus/pci/devices/") + busid + "/numa_node");
    int node = -1;
    if (!(fn >> node) || node < 0) return -1;
    std::ifstream fc("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string list;
    if (!std::getline(fc, list)) return -1;
    cpu_set_t want, have, both;
    if (parse_cpulist(list, &want) == 0) return -1;
    if (sched_getaffinity(0, sizeof have, &have) != 0) return -1;
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return -1;
    if (sched_setaffinity(0, sizeof both, &both) != 0) return -1;
    return node;
}

struct sdr_stream {
    std::vector<int> devices;
    int handles = 1;
    sdr_calibration calib{};
    sdr_pipeline_options opt{};
    int download_clouds = 0;
    std::atomic<int> input_format{SDR_PIX_BGR};   // applied by every worker at the start of its next job
    std::vector<std::pair<std::string, int>> sgbm_options;   // kernel knobs for every worker's matcher (set between jobs)

    // one job at a time; workers sleep between jobs
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    uint64_t job_id = 0;
    int busy = 0, ready = 0;
    bool quit = false;
    // the current job
    int64_t n_frames = 0;
    sdr_frame_source source = nullptr;
    sdr_frame_sink sink = nullptr;
    sdr_cloud_sink cloud_sink = nullptr;
    void* user = nullptr;
    std::atomic<int64_t> cursor{0};
    std::atomic<int> first_error{SDR_OK};
    std::vector<std::atomic<int64_t>> done;
    std::atomic<uint64_t> launches{0};
    std::atomic<int64_t> h2d_bytes{0}, d2h_bytes{0};
    std::string err_msg;
    int init_error = SDR_OK;
    std::vector<std::thread> threads;

    void worker(int di);
};

void sdr_stream::worker(int di) {
    sdr_bind_host_to_device(devices[di]);
    sdr_ctx* ctx = nullptr;
    sdr_pipeline* pipe = nullptr;
    const int B = opt.batch, W = calib.image_width, H = calib.image_height;
    int rc = sdr_create(devices[di], &ctx);
    if (rc == SDR_OK) rc = sdr_pipeline_create(ctx, &calib, &opt, &pipe);
    if (rc == SDR_OK && handles > 1) rc = sdr_pipeline_set_sgbm_option(pipe, "shared_device", 1);
    PinnedBuf disp_host, xyz_host, rgb_host;
    if (rc == SDR_OK) rc = disp_host.reserve(ctx, (size_t)B * W * H * 2);
    if (rc == SDR_OK && download_clouds) {
        rc = xyz_host.reserve(ctx, (size_t)B * W * H * 12);
        if (rc == SDR_OK) rc = rgb_host.reserve(ctx, (size_t)B * W * H * 3);
    }
    std::vector<sdr_image> sbs(B), disp(B);
    std::vector<sdr_cloud_view> views(B);
    std::vector<int64_t> idx(B), offs(B + 1);
    uint64_t seen = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (rc != SDR_OK && init_error == SDR_OK) {
            init_error = rc;
            err_msg = ctx ? sdr_last_error(ctx) : "sdr_create failed (no usable sm_100 device)";
        }
        ready++;
        cv_done.notify_all();
    }
    while (true) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_job.wait(lk, [&] { return quit || job_id != seen; });
            if (quit) break;
            seen = job_id;
        }
        const uint64_t l0 = ctx ? sdr_kernel_launches(ctx) : 0;
        int jrc = rc;
        const int fmt = input_format.load();
        if (jrc == SDR_OK) jrc = sdr_pipeline_set_input_format(pipe, fmt);
        for (const auto& kv : sgbm_options)
            if (jrc == SDR_OK) jrc = sdr_pipeline_set_sgbm_option(pipe, kv.first.c_str(), kv.second);
        int in_rows;
        size_t in_row_bytes;
        sdr_ingest_layout(fmt, 2 * W, H, &in_rows, &in_row_bytes);
        while (jrc == SDR_OK && first_error.load() == SDR_OK) {
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
            jrc = sdr_pipeline_run(pipe, n, sbs.data(), disp.data(), nullptr, nullptr, views.data());
            if (jrc == SDR_EVOXEL_OVERFLOW) jrc = SDR_OK;   // non-fatal: the views say which clouds were passed through
            if (jrc != SDR_OK) break;
            int64_t bytes_out = (int64_t)n * W * H * 2;
            for (int k = 0; k < n; k++)
                if (sbs[k].mem != SDR_MEM_DEVICE) h2d_bytes += (int64_t)in_rows * (int64_t)in_row_bytes;
            if (download_clouds) {
                jrc = sdr_pipeline_clouds_download_packed(pipe, (float*)xyz_host.p, (uint8_t*)rgb_host.p, nullptr, nullptr, offs.data());
                if (jrc != SDR_OK) break;
                bytes_out += offs[n] * 15;
            }
            d2h_bytes += bytes_out;
            for (int k = 0; k < n; k++) {
                if (sink) sink(user, idx[k], devices[di], &disp[k], &views[k]);
                if (cloud_sink && download_clouds)
                    cloud_sink(user, idx[k], devices[di], offs[k + 1] - offs[k], (const float*)xyz_host.p + offs[k] * 3,
                               (const uint8_t*)rgb_host.p + offs[k] * 3);
            }
            done[di] += n;
        }
        if (jrc != SDR_OK) {
            int expected = SDR_OK;
            if (first_error.compare_exchange_strong(expected, jrc)) {
                std::lock_guard<std::mutex> lk(mu);
                err_msg = ctx ? sdr_last_error(ctx) : "sdr_create failed";
            }
        }
        if (ctx) launches += sdr_kernel_launches(ctx) - l0;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (--busy == 0) cv_done.notify_all();
        }
    }
    disp_host.release();
    xyz_host.release();
    rgb_host.release();
    if (pipe) sdr_pipeline_destroy(pipe);
    if (ctx) sdr_destroy(ctx);
}

extern "C" int sdr_stream_create(const int* devices, int n_devices, int handles_per_device, const sdr_calibration* calib,
                                 const sdr_pipeline_options* opt, int download_clouds, sdr_stream** out) {
    if (!devices || n_devices <= 0 || n_devices > 16 || handles_per_device <= 0 || !calib || !opt || !out || opt->batch < 1) return SDR_EINVAL;
    *out = nullptr;
    sdr_stream* s = new sdr_stream();
    s->devices.assign(devices, devices + n_devices);
    s->handles = handles_per_device;
    s->calib = *calib;
    s->opt = *opt;
    s->download_clouds = download_clouds;
    s->done = std::vector<std::atomic<int64_t>>(n_devices);
    const int n_workers = n_devices * handles_per_device;
    for (int di = 0; di < n_devices; di++)
        for (int h = 0; h < handles_per_device; h++) s->threads.emplace_back([s, di] { s->worker(di); });
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv_done.wait(lk, [&] { return s->ready == n_workers; });
    }
    if (s->init_error != SDR_OK) {
        const int rc = s->init_error;
        fprintf(stderr, "sdr_stream_create: %s\n", s->err_msg.c_str());
        sdr_stream_destroy(s);
        return rc;
    }
    *out = s;
    return SDR_OK;
}

extern "C" void sdr_stream_destroy(sdr_stream* s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->quit = true;
    }
    s->cv_job.notify_all();
    for (auto& t : s->threads) t.join();
    delete s;
}

extern "C" int sdr_stream_process(sdr_stream* s, int64_t n_frames, sdr_frame_source source, sdr_frame_sink sink, sdr_cloud_sink cloud_sink,
                                  void* user, sdr_stream_stats* stats) {
    if (!s || !source || n_frames < 0) return SDR_EINVAL;
    const int n_workers = (int)s->threads.size();
    s->n_frames = n_frames;
    s->source = source;
    s->sink = sink;
    s->cloud_sink = cloud_sink;
    s->user = user;
    s->cursor = 0;
    s->first_error = SDR_OK;
    s->launches = 0;
    s->h2d_bytes = 0;
    s->d2h_bytes = 0;
    for (auto& d : s->done) d = 0;
    const auto t0 = std::chrono::steady_clock::now();
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->busy = n_workers;
        s->job_id++;
    }
    s->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv_done.wait(lk, [&] { return s->busy == 0; });
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        *stats = sdr_stream_stats{};
        for (size_t di = 0; di < s->devices.size(); di++) {
            stats->per_device[di] = s->done[di];
            stats->frames += s->done[di];
        }
        stats->seconds = secs;
        stats->frames_per_s = secs > 0 ? stats->frames / secs : 0;
        stats->kernel_launches = s->launches.load();
        stats->h2d_bytes = s->h2d_bytes.load();
        stats->d2h_bytes = s->d2h_bytes.load();
    }
    if (s->first_error.load() != SDR_OK) fprintf(stderr, "sdr_stream_process: %s\n", s->err_msg.c_str());
    return s->first_error.load();
}

extern "C" int sdr_stream_set_sgbm_option(sdr_stream* s, const char* name, int value) {
    if (!s || !name) return SDR_EINVAL;
    for (auto& kv : s->sgbm_options)
        if (kv.first == name) { kv.second = value; return SDR_OK; }
    s->sgbm_options.emplace_back(name, value);   // unknown names surface as SDR_EINVAL from the next process() call
    return SDR_OK;
}

extern "C" int sdr_stream_set_input_format(sdr_stream* s, int format) {
    if (!s || (format != SDR_PIX_BGR && format != SDR_PIX_NV12 && format != SDR_PIX_YUYV)) return SDR_EINVAL;
    const int W = s->calib.image_width, H = s->calib.image_height;
    if (format != SDR_PIX_BGR && ((W & 1) || (format == SDR_PIX_NV12 && (H & 1)))) return SDR_EINVAL;
    s->input_format = format;
    return SDR_OK;
}

extern "C" int sdr_stream_run(const int* devices, int n_devices, int handles_per_device, const sdr_calibration* calib,
                              const sdr_pipeline_options* opt, int64_t n_frames, sdr_frame_source source, sdr_frame_sink sink, void* user,
                              sdr_stream_stats* stats) {
    if (!source) return SDR_EINVAL;
    sdr_stream* s = nullptr;
    int rc = sdr_stream_create(devices, n_devices, handles_per_device, calib, opt, 0, &s);
    if (rc != SDR_OK) return rc;
    rc = sdr_stream_process(s, n_frames, source, sink, nullptr, user, stats);
    sdr_stream_destroy(s);
    return rc;
}

// ---- a frame source that needs no callback into the host language: a pool of SBS frames at a constant stride, frame i
// of the stream = pool frame (i mod n_pool) (BASELINE config 4: "64 unique pairs cycled to 4096") ----
struct sdr_frame_pool {
    const uint8_t* base;
    int64_t n_pool;
    size_t frame_stride, step_bytes;
    int rows, cols, mem, channels;
};
extern "C" int sdr_frame_pool_source(void* user, int64_t index, sdr_image* sbs) {
    const sdr_frame_pool* p = (const sdr_frame_pool*)user;
    if (!p || p->n_pool <= 0) return SDR_EINVAL;
    *sbs = sdr_image{(void*)(p->base + (size_t)(index % p->n_pool) * p->frame_stride), p->rows, p->cols, p->channels, 1, p->step_bytes, p->mem};
    return SDR_OK;
}
extern "C" int sdr_stream_process_pool(sdr_stream* s, int64_t n_frames, const void* pool_base, int64_t n_pool, size_t frame_stride,
                                       size_t step_bytes, int mem, sdr_stream_stats* stats) {
    if (!s || !pool_base || n_pool <= 0) return SDR_EINVAL;
    const int fmt = s->input_format.load();
    int rows;
    size_t row_bytes;
    sdr_ingest_layout(fmt, 2 * s->calib.image_width, s->calib.image_height, &rows, &row_bytes);
    sdr_frame_pool p{(const uint8_t*)pool_base, n_pool, frame_stride, step_bytes, rows, 2 * s->calib.image_width, mem,
                     fmt == SDR_PIX_BGR ? 3 : fmt == SDR_PIX_NV12 ? 1 : 2};
    return sdr_stream_process(s, n_frames, sdr_frame_pool_source, nullptr, nullptr, &p, stats);
}
