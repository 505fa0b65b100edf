// Context, buffers, host<->device staging, stage timers and the OpenCV-dialect YAML reader /
// writer that stands in for cv::FileStorage in StereoConfiguration::loadFromFile / saveToFile
// (stereo_vision/src/stereo_configuration.cpp:4-46, :49-74 of the reference).
#include "sdr_internal.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

void sdr_set_error(sdr_ctx* ctx, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
}

cudaError_t sdr_wait(sdr_ctx* ctx) {
    if (!ctx->wait_ev) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->wait_ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(ctx->wait_ev, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->wait_ev);
}

int DevBuf::reserve(sdr_ctx* ctx, size_t bytes) {
    if (bytes <= cap) return SDR_OK;
    if (p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // round up so that small size changes do not reallocate
    size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        return SDR_FAIL(ctx, SDR_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return SDR_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
int PinnedBuf::reserve(sdr_ctx* ctx, size_t bytes) {
    if (bytes <= cap) return SDR_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) {
        p = nullptr;
        return SDR_FAIL(ctx, SDR_ENOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = bytes;
    return SDR_OK;
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

void StageTimer::begin(cudaStream_t s) {
    n = 0;
    names.clear();
    if (!enabled) return;
    if (ev.empty()) {
        ev.resize(64);
        for (auto& e : ev) cudaEventCreate(&e);
    }
    cudaEventRecord(ev[0], s);
    n = 1;
}
void StageTimer::mark(cudaStream_t s, const char* name) {
    if (!enabled || n == 0 || n >= ev.size()) return;
    cudaEventRecord(ev[n++], s);
    names.push_back(name);
}
int StageTimer::collect(int cap, const char** out_names, float* ms) {
    if (!enabled || n < 2) return 0;
    cudaEventSynchronize(ev[n - 1]);
    int k = 0;
    for (size_t i = 1; i < n && k < cap; i++, k++) {
        float t = 0;
        cudaEventElapsedTime(&t, ev[i - 1], ev[i]);
        if (out_names) out_names[k] = names[i - 1];
        if (ms) ms[k] = t;
    }
    return k;
}
// Times of begin() and of every mark since `base` (an event recorded earlier on any stream of the device): a timeline that
// can be laid beside those of other handles.
int StageTimer::collect_since(cudaEvent_t base, int cap, const char** out_names, float* ms) {
    if (!enabled || n < 1) return 0;
    cudaEventSynchronize(ev[n - 1]);
    int k = 0;
    for (size_t i = 0; i < n && k < cap; i++, k++) {
        float t = 0;
        cudaEventElapsedTime(&t, base, ev[i]);
        if (out_names) out_names[k] = i == 0 ? "begin" : names[i - 1];
        if (ms) ms[k] = t;
    }
    return k;
}
void StageTimer::release() {
    for (auto& e : ev) cudaEventDestroy(e);
    ev.clear();
}

int sdr_check_image(sdr_ctx* ctx, const sdr_image* im, int channels, int elem_size, const char* what) {
    if (!im || !im->data) return SDR_FAIL(ctx, SDR_EINVAL, "%s: empty image", what);
    if (im->rows <= 0 || im->cols <= 0) return SDR_FAIL(ctx, SDR_ESIZE, "%s: bad size %dx%d", what, im->cols, im->rows);
    if (im->channels != channels || im->elem_size != elem_size)
        return SDR_FAIL(ctx, SDR_ESIZE, "%s: expected %d channel(s) of %d byte(s), got %d x %d", what, channels, elem_size,
                        im->channels, im->elem_size);
    if (im->step_bytes < (size_t)im->cols * channels * elem_size) return SDR_FAIL(ctx, SDR_ESIZE, "%s: step too small", what);
    if (im->mem != SDR_MEM_HOST && im->mem != SDR_MEM_PINNED && im->mem != SDR_MEM_DEVICE)
        return SDR_FAIL(ctx, SDR_EINVAL, "%s: bad mem kind %d", what, im->mem);
    if (im->mem == SDR_MEM_DEVICE) {
        // Device images are used in place by the kernels of ctx->device: the pointer must be device memory, and when it lives on
        // another GPU (a frame pool on GPU 0 read by the workers of a multi-GPU stream) that GPU must be a peer -- enabled here on
        // first use (NVLink / PCIe P2P) or refused, never a fault inside a kernel.
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, im->data) != cudaSuccess) {
            cudaGetLastError();
            return SDR_FAIL(ctx, SDR_EINVAL, "%s: not a CUDA pointer", what);
        }
        if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)
            return SDR_FAIL(ctx, SDR_EINVAL, "%s: mem says device memory, the pointer is not", what);
        if (at.type == cudaMemoryTypeDevice && at.device != ctx->device && at.device >= 0 && at.device < 64 &&
            !(ctx->peer_ok >> at.device & 1ull)) {
            int can = 0;
            cudaSetDevice(ctx->device);
            if (cudaDeviceCanAccessPeer(&can, ctx->device, at.device) != cudaSuccess || !can) {
                cudaGetLastError();
                return SDR_FAIL(ctx, SDR_EINVAL, "%s: the image lives on GPU %d, which GPU %d cannot access (no peer access)", what, at.device,
                                ctx->device);
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            cudaGetLastError();
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return SDR_FAIL(ctx, SDR_ECUDA, "%s: cudaDeviceEnablePeerAccess(%d) failed: %s", what, at.device, cudaGetErrorString(e));
            ctx->peer_ok |= 1ull << at.device;
        }
    }
    return SDR_OK;
}

int sdr_stage_in(sdr_ctx* ctx, const sdr_image* im, DevBuf& scratch, DevView* out) {
    if (im->mem == SDR_MEM_DEVICE) {
        out->p = (uint8_t*)im->data;
        out->pitch = im->step_bytes;
        return SDR_OK;
    }
    const size_t rowb = (size_t)im->cols * im->channels * im->elem_size;
    int rc = scratch.reserve(ctx, rowb * im->rows);
    if (rc) return rc;
    SDR_CUDA(ctx, cudaMemcpy2DAsync(scratch.p, rowb, im->data, im->step_bytes, rowb, im->rows, cudaMemcpyHostToDevice, ctx->stream));
    out->p = (uint8_t*)scratch.p;
    out->pitch = rowb;
    return SDR_OK;
}

int sdr_stage_out_begin(sdr_ctx* ctx, const sdr_image* im, DevBuf& scratch, DevView* out) {
    if (im->mem == SDR_MEM_DEVICE) {
        out->p = (uint8_t*)im->data;
        out->pitch = im->step_bytes;
        return SDR_OK;
    }
    const size_t rowb = (size_t)im->cols * im->channels * im->elem_size;
    int rc = scratch.reserve(ctx, rowb * im->rows);
    if (rc) return rc;
    out->p = (uint8_t*)scratch.p;
    out->pitch = rowb;
    return SDR_OK;
}

int sdr_stage_out_end(sdr_ctx* ctx, const sdr_image* im, const DevView& v) {
    if (im->mem == SDR_MEM_DEVICE) return SDR_OK;
    const size_t rowb = (size_t)im->cols * im->channels * im->elem_size;
    SDR_CUDA(ctx, cudaMemcpy2DAsync(im->data, im->step_bytes, v.p, v.pitch, rowb, im->rows, cudaMemcpyDeviceToHost, ctx->stream));
    SDR_CUDA(ctx, sdr_wait(ctx));
    return SDR_OK;
}

extern "C" const char* sdr_version(void) { return "stereo_depth_ruler_b200 0.1 (sm_100a)"; }

extern "C" int sdr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int sdr_create(int device, sdr_ctx** out) {
    if (!out) return SDR_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SDR_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return SDR_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SDR_ECUDA;
    if (prop.major < 10) return SDR_ECUDA;  // sm_100a code only
    sdr_ctx* c = new sdr_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return SDR_ECUDA;
    }
    c->stream = c->own_stream;
    *out = c;
    return SDR_OK;
}

extern "C" void sdr_destroy(sdr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->wait_ev) cudaEventDestroy(ctx->wait_ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" int sdr_set_stream(sdr_ctx* ctx, void* s) {
    if (!ctx) return SDR_EINVAL;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return SDR_OK;
}
extern "C" void* sdr_get_stream(sdr_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int sdr_synchronize(sdr_ctx* ctx) {
    if (!ctx) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    SDR_CUDA(ctx, sdr_wait(ctx));
    return SDR_OK;
}
extern "C" const char* sdr_last_error(sdr_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" uint64_t sdr_kernel_launches(sdr_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------
// OpenCV FileStorage YAML 1.0 subset: "key: scalar" and "key: !!opencv-matrix" blocks with
// rows / cols / dt / data: [ ... ] (data may span lines). Enough for config/stereo.yaml.
// ------------------------------------------------------------------------------------------
namespace {
struct YMat {
    int rows = 0, cols = 0;
    std::vector<double> data;
};

bool parse_yaml(const std::string& text, std::map<std::string, YMat>& mats, std::map<std::string, double>& scalars) {
    std::istringstream in(text);
    std::string line, cur;
    bool in_data = false;
    std::string data_acc;
    auto trim = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    auto flush_data = [&]() {
        std::string s = data_acc;
        for (auto& ch : s)
            if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
        std::istringstream ds(s);
        std::string tok;
        while (ds >> tok) {
            char* end = nullptr;
            double v = strtod(tok.c_str(), &end);
            if (end == tok.c_str()) {
                if (tok == ".Inf" || tok == ".inf") v = INFINITY;
                else if (tok == "-.Inf" || tok == "-.inf") v = -INFINITY;
                else if (tok == ".Nan" || tok == ".NaN" || tok == ".nan") v = NAN;
                else continue;
            }
            mats[cur].data.push_back(v);
        }
        data_acc.clear();
        in_data = false;
    };
    while (std::getline(in, line)) {
        std::string t = trim(line);
        if (t.empty() || t[0] == '%' || t == "---" || t == "...") continue;
        if (in_data) {
            data_acc += " " + t;
            if (t.find(']') != std::string::npos) flush_data();
            continue;
        }
        size_t colon = t.find(':');
        if (colon == std::string::npos) continue;
        std::string key = trim(t.substr(0, colon)), val = trim(t.substr(colon + 1));
        const bool indented = !line.empty() && (line[0] == ' ' || line[0] == '\t');
        if (!indented) {
            if (val.rfind("!!opencv-matrix", 0) == 0) {
                cur = key;
                mats[cur] = YMat();
            } else if (!val.empty()) {
                char* end = nullptr;
                double v = strtod(val.c_str(), &end);
                if (end != val.c_str()) scalars[key] = v;
                cur.clear();
            } else {
                cur.clear();
            }
        } else if (!cur.empty()) {
            if (key == "rows") mats[cur].rows = atoi(val.c_str());
            else if (key == "cols") mats[cur].cols = atoi(val.c_str());
            else if (key == "data") {
                in_data = true;
                data_acc = val;
                if (val.find(']') != std::string::npos) flush_data();
            }
        }
    }
    return true;
}

struct Field {
    const char* name;
    int rows, cols;
    size_t offset;
};
const Field kFields[] = {
    {"cameraMatrixLeft", 3, 3, offsetof(sdr_calibration, K_left)},
    {"distCoeffsLeft", 1, 5, offsetof(sdr_calibration, dist_left)},
    {"cameraMatrixRight", 3, 3, offsetof(sdr_calibration, K_right)},
    {"distCoeffsRight", 1, 5, offsetof(sdr_calibration, dist_right)},
    {"R", 3, 3, offsetof(sdr_calibration, R)},
    {"T", 3, 1, offsetof(sdr_calibration, T)},
    {"E", 3, 3, offsetof(sdr_calibration, E)},
    {"F", 3, 3, offsetof(sdr_calibration, F)},
    {"R1", 3, 3, offsetof(sdr_calibration, R1)},
    {"R2", 3, 3, offsetof(sdr_calibration, R2)},
    {"P1", 3, 4, offsetof(sdr_calibration, P1)},
    {"P2", 3, 4, offsetof(sdr_calibration, P2)},
    {"Q", 4, 4, offsetof(sdr_calibration, Q)},
};
}  // namespace

extern "C" int sdr_load_calibration_yaml(const char* path, sdr_calibration* out) {
    if (!path || !out) return SDR_EINVAL;
    std::ifstream f(path);
    if (!f) return SDR_EIO;
    std::stringstream ss;
    ss << f.rdbuf();
    std::map<std::string, YMat> mats;
    std::map<std::string, double> scalars;
    parse_yaml(ss.str(), mats, scalars);
    memset(out, 0, sizeof *out);
    out->image_width = scalars.count("imageWidth") ? (int)scalars["imageWidth"] : 0;
    out->image_height = scalars.count("imageHeight") ? (int)scalars["imageHeight"] : 0;
    // stereo_configuration.cpp:15-19: a non-positive size is an error
    if (out->image_width <= 0 || out->image_height <= 0) return SDR_EIO;
    int bit = 0;
    for (const Field& fd : kFields) {
        auto it = mats.find(fd.name);
        if (it != mats.end()) {
            const YMat& m = it->second;
            const size_t want = (size_t)fd.rows * fd.cols;
            // distortion vectors may be stored 1x5 or 5x1, T as 3x1 or 1x3; longer dist vectors are truncated to 5
            if (m.data.size() >= want && (size_t)m.rows * m.cols == m.data.size()) {
                memcpy((char*)out + fd.offset, m.data.data(), want * sizeof(double));
                out->present |= 1u << bit;
            }
        }
        bit++;
    }
    // stereo_configuration.cpp:36-41: the essential matrices must be present
    return sdr_calibration_is_valid(out) && (out->present & (1u << 8)) && (out->present & (1u << 9)) && (out->present & (1u << 10)) &&
                   (out->present & (1u << 11))
               ? SDR_OK
               : SDR_EIO;
}

extern "C" int sdr_calibration_is_valid(const sdr_calibration* c) {
    if (!c) return 0;
    // stereo_configuration.cpp:77-80: K_left, K_right, R, T, Q non-empty
    const uint32_t need = (1u << 0) | (1u << 2) | (1u << 4) | (1u << 5) | (1u << 12);
    return (c->present & need) == need;
}

extern "C" int sdr_save_calibration_yaml(const char* path, const sdr_calibration* in) {
    if (!path || !in) return SDR_EINVAL;
    FILE* f = fopen(path, "w");
    if (!f) return SDR_EIO;
    fprintf(f, "%%YAML:1.0\n---\nimageWidth: %d\nimageHeight: %d\n", in->image_width, in->image_height);
    for (const Field& fd : kFields) {
        const double* d = (const double*)((const char*)in + fd.offset);
        fprintf(f, "%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: d\n   data: [ ", fd.name, fd.rows, fd.cols);
        for (int i = 0; i < fd.rows * fd.cols; i++) fprintf(f, "%s%.17g", i ? ", " : "", d[i]);
        fprintf(f, " ]\n");
    }
    fprintf(f, "imageSize: [ %d, %d ]\n", in->image_width, in->image_height);
    fclose(f);
    return SDR_OK;
}
