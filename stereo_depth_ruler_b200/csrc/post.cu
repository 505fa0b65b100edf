// medianBlur(3) on CV_16S and filterSpeckles, the two post-filters cv::StereoSGBM::compute
// applies internally. Both are HBM-bound elementwise / union-find work.
#include "sdr_internal.h"

namespace {

__device__ __forceinline__ void cswap(int& a, int& b) {
    const int lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

// 3x3 median, BORDER_REPLICATE; invalid values take part as numbers.
__global__ void __launch_bounds__(256) k_median3(const int16_t* __restrict__ src, size_t spitch, size_t sframe,
                                                 int16_t* __restrict__ dst, size_t dpitch, size_t dframe, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const uint8_t* sb = (const uint8_t*)src + (size_t)b * sframe;
    int v[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int16_t* r = (const int16_t*)(sb + (size_t)min(max(y + j - 1, 0), H - 1) * spitch);
#pragma unroll
        for (int i = 0; i < 3; i++) v[j * 3 + i] = r[min(max(x + i - 1, 0), W - 1)];
    }
    // 19-exchange median-of-9 network
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
    cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
    cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
    cswap(v[4], v[2]);
    ((int16_t*)((uint8_t*)dst + (size_t)b * dframe + (size_t)y * dpitch))[x] = (int16_t)v[4];
}

// ---- speckle filter: lock-free union-find over the 4-connected "|a-b| <= maxDiff" graph ----
__device__ __forceinline__ int uf_find(const int* parent, int i) {
    int c = i;
    while (true) {
        const int n = __ldcg(parent + c);
        if (n == c) return c;
        c = n;
    }
}

// find with path halving: every visited node is re-pointed at its grandparent. Racing with the atomicMin links is benign:
// a node only ever points at a smaller index of its own component, and a link lost to such a store is re-established by
// the linker's own loop (it goes on to unite the old parent with its target).
__device__ __forceinline__ int uf_find_halving(int* parent, int i) {
    int c = i;
    while (true) {
        const int n = __ldcg(parent + c);
        if (n == c) return c;
        const int nn = __ldcg(parent + n);
        if (nn != n) parent[c] = nn;
        c = nn;
    }
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find_halving(parent, a);
        b = uf_find_halving(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

__device__ __forceinline__ const int16_t* img_row(const int16_t* img, size_t pitch, size_t frame, int b, int y) {
    return (const int16_t*)((const uint8_t*)img + (size_t)b * frame + (size_t)y * pitch);
}

// Row pass: every pixel's parent starts at the first pixel of its horizontal run (a max-scan of the run-start
// positions along the row), so only vertical links remain to be merged and the trees start out flat.
__global__ void __launch_bounds__(256) k_speckle_rows(const int16_t* img, size_t pitch, size_t frame, int W, int H, int new_val,
                                                      int max_diff, int* parent, int* count) {
    __shared__ int warp_max[8];
    __shared__ int carry_s;
    const int y = blockIdx.x, b = blockIdx.z;
    const int16_t* r = img_row(img, pitch, frame, b, y);
    int* par = parent + (size_t)b * W * H + (size_t)y * W;
    int* cnt = count + (size_t)b * W * H + (size_t)y * W;
    if (threadIdx.x == 0) carry_s = -1;
    __syncthreads();
    for (int x0 = 0; x0 < W; x0 += 256) {
        const int x = x0 + threadIdx.x;
        int start = -1;   // position of the most recent run start at or before x (within this chunk)
        bool valid = false;
        if (x < W) {
            const int v = r[x];
            valid = v != new_val;
            bool conn_left = false;
            if (valid && x > 0) {
                const int u = r[x - 1];
                conn_left = u != new_val && abs(u - v) <= max_diff;
            }
            if (!conn_left) start = x;
        }
        int m = start;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, m, o);
            if ((threadIdx.x & 31) >= o) m = max(m, t);
        }
        if ((threadIdx.x & 31) == 31) warp_max[threadIdx.x >> 5] = m;
        __syncthreads();
        int pre = carry_s;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) pre = max(pre, warp_max[w]);
        m = max(m, pre);
        if (x < W) {
            par[x] = valid ? y * W + m : y * W + x;
            cnt[x] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 255) carry_s = m;
        __syncthreads();
    }
}

// Vertical links. A link at x is implied by the one at x-1 when both rows continue their runs from x-1 to x.
__global__ void __launch_bounds__(256) k_speckle_merge(const int16_t* img, size_t pitch, size_t frame, int W, int H, int new_val,
                                                       int max_diff, int* parent) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W || y + 1 >= H) return;
    int* par = parent + (size_t)b * W * H;
    const int16_t* r0 = img_row(img, pitch, frame, b, y);
    const int16_t* r1 = img_row(img, pitch, frame, b, y + 1);
    const int v = r0[x], u = r1[x];
    if (v == new_val || u == new_val || abs(u - v) > max_diff) return;
    if (x > 0) {
        const int vl = r0[x - 1], ul = r1[x - 1];
        if (vl != new_val && ul != new_val && abs(vl - v) <= max_diff && abs(ul - u) <= max_diff && abs(ul - vl) <= max_diff) return;
    }
    uf_union(par, y * W + x, (y + 1) * W + x);
}

// Flatten to roots and count component sizes (one atomic per distinct root per warp).
__global__ void __launch_bounds__(256) k_speckle_count(const int16_t* img, size_t pitch, size_t frame, int W, int H, int new_val,
                                                       int* parent, int* count) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    const bool in = x < W;
    int* par = parent + (size_t)b * W * H;
    int* cnt = count + (size_t)b * W * H;
    int root = -1;
    if (in) {
        const int v = img_row(img, pitch, frame, b, y)[x];
        if (v != new_val) {
            root = uf_find(par, y * W + x);
            par[y * W + x] = root;
        }
    }
    const unsigned act = __ballot_sync(0xffffffffu, root >= 0);
    if (root >= 0) {
        const unsigned peers = __match_any_sync(act, root);
        const int leader = __ffs(peers) - 1;
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(cnt + root, __popc(peers));
    }
}

__global__ void __launch_bounds__(256) k_speckle_apply(int16_t* img, size_t pitch, size_t frame, int W, int H, int new_val,
                                                       int max_size, const int* parent, const int* count) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    int16_t* r = (int16_t*)((uint8_t*)img + (size_t)b * frame + (size_t)y * pitch);
    if (r[x] == new_val) return;
    const int root = parent[(size_t)b * W * H + y * W + x];
    if (count[(size_t)b * W * H + root] <= max_size) r[x] = (int16_t)new_val;
}

}  // namespace

int sdr_launch_median3(sdr_ctx* ctx, const int16_t* src, size_t spitch, int16_t* dst, size_t dpitch, int W, int H, int batch,
                       size_t sframe, size_t dframe) {
    k_median3<<<dim3(div_up(W, 256), H, batch), 256, 0, ctx->stream>>>(src, spitch, sframe, dst, dpitch, dframe, W, H);
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

int sdr_launch_speckle(sdr_ctx* ctx, int16_t* img, size_t pitch, int W, int H, int batch, size_t frame_bytes, int new_val,
                       int max_size, int max_diff, SpeckleScratch& s) {
    const size_t n = (size_t)W * H;
    int rc;
    if ((rc = s.parent.reserve(ctx, n * batch * sizeof(int)))) return rc;
    if ((rc = s.count.reserve(ctx, n * batch * sizeof(int)))) return rc;
    const dim3 grid(div_up(W, 256), H, batch);
    k_speckle_rows<<<dim3(H, 1, batch), 256, 0, ctx->stream>>>(img, pitch, frame_bytes, W, H, new_val, max_diff, s.parent.as<int>(),
                                                               s.count.as<int>());
    SDR_LAUNCHED(ctx);
    k_speckle_merge<<<grid, 256, 0, ctx->stream>>>(img, pitch, frame_bytes, W, H, new_val, max_diff, s.parent.as<int>());
    SDR_LAUNCHED(ctx);
    k_speckle_count<<<grid, 256, 0, ctx->stream>>>(img, pitch, frame_bytes, W, H, new_val, s.parent.as<int>(), s.count.as<int>());
    SDR_LAUNCHED(ctx);
    k_speckle_apply<<<grid, 256, 0, ctx->stream>>>(img, pitch, frame_bytes, W, H, new_val, max_size, s.parent.as<int>(),
                                                   s.count.as<int>());
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

// ---- C ABI wrappers (exposed for unit parity) ----
namespace {
struct PostScratch {
    DevBuf in, out;
    SpeckleScratch sp;
};
}  // namespace

extern "C" int sdr_median3_16s(sdr_ctx* ctx, const sdr_image* src, sdr_image* dst) {
    if (!ctx || !src || !dst) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, src, 1, 2, "src"))) return rc;
    if ((rc = sdr_check_image(ctx, dst, 1, 2, "dst"))) return rc;
    if (src->rows != dst->rows || src->cols != dst->cols) return SDR_FAIL(ctx, SDR_ESIZE, "median3: size mismatch");
    PostScratch sc;
    DevView vi, vo;
    if ((rc = sdr_stage_in(ctx, src, sc.in, &vi)) == SDR_OK && (rc = sdr_stage_out_begin(ctx, dst, sc.out, &vo)) == SDR_OK) {
        rc = sdr_launch_median3(ctx, (const int16_t*)vi.p, vi.pitch, (int16_t*)vo.p, vo.pitch, src->cols, src->rows, 1, 0, 0);
        if (rc == SDR_OK) rc = sdr_stage_out_end(ctx, dst, vo);
    }
    cudaStreamSynchronize(ctx->stream);
    sc.in.release();
    sc.out.release();
    return rc;
}

extern "C" int sdr_filter_speckles(sdr_ctx* ctx, sdr_image* img, int new_val, int max_size, int max_diff) {
    if (!ctx || !img) return SDR_EINVAL;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, img, 1, 2, "img"))) return rc;
    PostScratch sc;
    DevView v;
    if (img->mem == SDR_MEM_DEVICE) {
        v.p = (uint8_t*)img->data;
        v.pitch = img->step_bytes;
        rc = SDR_OK;
    } else {
        rc = sdr_stage_in(ctx, img, sc.in, &v);
    }
    if (rc == SDR_OK)
        rc = sdr_launch_speckle(ctx, (int16_t*)v.p, v.pitch, img->cols, img->rows, 1, 0, new_val, max_size, max_diff, sc.sp);
    if (rc == SDR_OK && img->mem != SDR_MEM_DEVICE) rc = sdr_stage_out_end(ctx, img, v);
    cudaStreamSynchronize(ctx->stream);
    sc.in.release();
    sc.out.release();
    sc.sp.parent.release();
    sc.sp.count.release();
    return rc;
}
