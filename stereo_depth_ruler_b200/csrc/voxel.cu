// convertCVMatToPCL + pcl::VoxelGrid<PointXYZRGB>::filter (point_cloud/src/pcd_write.cpp:17-51,
// :123-130 of the reference; PCL 1.14 semantics restated in SURVEY.md Appendix A.8), optionally fused
// with cv::reprojectImageTo3D so the W*H*3 float image never reaches HBM.
//
// PCL's algorithm is kept step for step so that the output is the same cloud in the same order:
//   bbox of the finite points -> int32 index guard (passthrough when it trips) -> per point
//   idx = i + j*dx + k*dx*dy -> sort by idx -> one output point per run of equal idx (fp32 sums in
//   sorted order, colour means truncated). The sort is a stable LSD radix sort written here
//   (4 x 8-bit passes: histogram, digit-major scan, ranked scatter).
#include "sdr_internal.h"

#include <cmath>
#include <cstring>

namespace {

struct QMat {
    double q[16];
};

struct VoxParams {  // lives on the device; written by k_vox_params
    float inv[3];
    int minb[3], divb[3], mul[3];
    int overflow;
    unsigned nvalid;
    int bbox_key[6];  // min xyz, max xyz as ordered-int keys
    unsigned nvox;
    int min_disp_key;
    unsigned nsort;   // points that enter the sort: nvalid, or 0 when the index guard trips (passthrough)
};

struct Source {  // where a pixel's 3-D point comes from
    const uint8_t* points;  // f32x3 image, or null
    size_t pstep;
    const uint8_t* disp;  // int16 image (required when points == null or drop_invalid)
    size_t dstep;
    const uint8_t* bgr;  // optional
    size_t cstep;
    size_t pframe, dframe, cframe;  // bytes between consecutive frames of a batch
    int W, H;
    QMat Q;
    int handle_missing, from_disp;
    int drop_invalid, min_valid16;  // drop when disp16 < min_valid16
    float z_max;
};

__device__ __forceinline__ int fkey(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float kfloat(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }
// static_cast<int>(float) as the x86 build of PCL performs it (cvttss2si): a value outside the int range gives INT_MIN, where
// the GPU's conversion saturates. Only reached with absurd leaf sizes whose int64 index guard wraps around and passes.
__device__ __forceinline__ int f2i_x86(float v) { return (v >= 2147483648.f || v < -2147483648.f || v != v) ? (int)0x80000000 : (int)v; }

// xyz of pixel i; returns false when the point does not enter the voxel grid.
__device__ __forceinline__ bool fetch_point(const Source& s, const VoxParams* vp, int fr, int i, float& X, float& Y, float& Z) {
    const int y = i / s.W, x = i - y * s.W;
    int d16 = 0;
    if (s.disp) d16 = ((const int16_t*)(s.disp + (size_t)fr * s.dframe + (size_t)y * s.dstep))[x];
    if (s.from_disp) {
        const float df = (float)d16 * 0.0625f;
        const double d = (double)df, dx = (double)x, dy = (double)y;
        double h[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
            h[c] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(s.Q.q[c * 4], dx), __dmul_rn(s.Q.q[c * 4 + 1], dy)), __dmul_rn(s.Q.q[c * 4 + 2], d)),
                             s.Q.q[c * 4 + 3]);
        const double iw = __ddiv_rn(1.0, h[3]);
        X = (float)__dmul_rn((double)(float)h[0], iw);
        Y = (float)__dmul_rn((double)(float)h[1], iw);
        Z = (float)__dmul_rn((double)(float)h[2], iw);
        if (s.handle_missing && fabs(d - (double)kfloat(vp->min_disp_key)) <= 1.1920928955078125e-07) Z = 10000.f;
    } else {
        const float* p = (const float*)(s.points + (size_t)fr * s.pframe + (size_t)y * s.pstep) + (size_t)x * 3;
        X = p[0]; Y = p[1]; Z = p[2];
    }
    if (!(isfinite(X) && isfinite(Y) && isfinite(Z))) return false;  // pcd_write.cpp:35 + PCL isXYZFinite
    if (s.drop_invalid && d16 < s.min_valid16) return false;
    if (s.z_max > 0.f && !(Z > 0.f && Z <= s.z_max)) return false;
    return true;
}

__global__ void k_vox_reset(VoxParams* vp0) {
    VoxParams* vp = vp0 + blockIdx.x;
    vp->bbox_key[0] = vp->bbox_key[1] = vp->bbox_key[2] = 0x7FFFFFFF;
    vp->bbox_key[3] = vp->bbox_key[4] = vp->bbox_key[5] = (int)0x80000000;
    vp->nvalid = 0;
    vp->nvox = 0;
    vp->nsort = 0;
    vp->overflow = 0;
    vp->min_disp_key = 0x7FFFFFFF;
}

constexpr int VOX_PPT = 8;   // pixels per thread in the reduction kernels: 8x fewer same-address atomics per frame

__global__ void __launch_bounds__(256) k_vox_mindisp(Source s, VoxParams* vp0) {
    const int fr = blockIdx.y;
    VoxParams* vp = vp0 + fr;
    int key = 0x7FFFFFFF;
#pragma unroll
    for (int k = 0; k < VOX_PPT; k++) {
        const int i = (blockIdx.x * VOX_PPT + k) * blockDim.x + threadIdx.x;
        if (i < s.W * s.H) {
            const int y = i / s.W, x = i - y * s.W;
            key = min(key, fkey((float)((const int16_t*)(s.disp + (size_t)fr * s.dframe + (size_t)y * s.dstep))[x] * 0.0625f));
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    __shared__ int wk[8];
    if ((threadIdx.x & 31) == 0) wk[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) key = min(key, wk[w]);
        if (key != 0x7FFFFFFF) atomicMin(&vp->min_disp_key, key);
    }
}

__global__ void __launch_bounds__(256) k_vox_bbox(Source s, VoxParams* vp0, uint32_t* __restrict__ blk_cnt) {
    const int fr = blockIdx.y;
    VoxParams* vp = vp0 + fr;
    int mn[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    unsigned ok = 0;
#pragma unroll 2
    for (int k = 0; k < VOX_PPT; k++) {
        const int i = (blockIdx.x * VOX_PPT + k) * blockDim.x + threadIdx.x;
        if (i < s.W * s.H) {
            float p[3];
            if (fetch_point(s, vp, fr, i, p[0], p[1], p[2])) {
                ok++;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int key = fkey(p[c]);
                    mn[c] = min(mn[c], key);
                    mx[c] = max(mx[c], key);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        ok += __shfl_xor_sync(0xffffffffu, ok, o);
    }
    __shared__ int smn[8][3], smx[8][3];
    __shared__ unsigned sok[8];
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
#pragma unroll
        for (int c = 0; c < 3; c++) { smn[w][c] = mn[c]; smx[w][c] = mx[c]; }
        sok[w] = ok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
#pragma unroll
            for (int c = 0; c < 3; c++) { mn[c] = min(mn[c], smn[w][c]); mx[c] = max(mx[c], smx[w][c]); }
            ok += sok[w];
        }
        blk_cnt[(size_t)fr * gridDim.x + blockIdx.x] = ok;   // points of this block's 2048 pixels that enter the grid (k_vox_keys compacts with it)
        if (ok) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                atomicMin(&vp->bbox_key[c], mn[c]);
                atomicMax(&vp->bbox_key[3 + c], mx[c]);
            }
            atomicAdd(&vp->nvalid, ok);
        }
    }
}

// PCL: inverse leaf, int64 guard, min_b / div_b / divb_mul (all fp32 as in voxel_grid.hpp)
__global__ void k_vox_params(VoxParams* vp0, float lx, float ly, float lz) {
    VoxParams* vp = vp0 + blockIdx.x;
    const float leaf[3] = {lx, ly, lz};
    if (vp->nvalid == 0) {
        for (int c = 0; c < 3; c++) { vp->inv[c] = __fdiv_rn(1.0f, leaf[c]); vp->minb[c] = 0; vp->divb[c] = 1; vp->mul[c] = 1; }
        return;
    }
    long long d[3];
    for (int c = 0; c < 3; c++) {
        const float inv = __fdiv_rn(1.0f, leaf[c]);
        const float mn = kfloat(vp->bbox_key[c]), mx = kfloat(vp->bbox_key[3 + c]);
        vp->inv[c] = inv;
        d[c] = (long long)__fmul_rn(__fsub_rn(mx, mn), inv) + 1;
        vp->minb[c] = f2i_x86(floorf(__fmul_rn(mn, inv)));
        const int maxb = f2i_x86(floorf(__fmul_rn(mx, inv)));
        vp->divb[c] = (int)((unsigned)maxb - (unsigned)vp->minb[c] + 1u);
    }
    // PCL multiplies in int64; for absurd boxes the product wraps (two's complement, as on x86)
    const long long prod = (long long)((unsigned long long)d[0] * (unsigned long long)d[1] * (unsigned long long)d[2]);
    vp->overflow = prod > 2147483647LL ? 1 : 0;
    vp->nsort = vp->overflow ? 0u : vp->nvalid;
    vp->mul[0] = 1;
    vp->mul[1] = vp->divb[0];
    vp->mul[2] = (int)((unsigned)vp->divb[0] * (unsigned)vp->divb[1]);
}

// Keys (voxel index) and values (pixel index) of the points that enter the grid, compacted in ascending pixel order: block
// (bx, fr) owns the same 2048 pixels as in k_vox_bbox, whose per-block counts (scanned) say where its points go. Pixels that
// do not enter the grid (62 % of a typical frame do) are never sorted.
__global__ void __launch_bounds__(256) k_vox_keys(Source s, const VoxParams* vp0, const uint32_t* __restrict__ blk_off, uint32_t* keys,
                                                  uint32_t* vals, int npad) {
    const int fr = blockIdx.y;
    const VoxParams* vp = vp0 + fr;
    if (vp->nsort == 0) return;
    keys += (size_t)fr * npad;
    vals += (size_t)fr * npad;
    __shared__ uint32_t cnt[VOX_PPT * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t key[VOX_PPT];
    unsigned okmask = 0;
    uint32_t rank[VOX_PPT];
#pragma unroll
    for (int k = 0; k < VOX_PPT; k++) {
        const int i = (blockIdx.x * VOX_PPT + k) * blockDim.x + threadIdx.x;
        bool ok = false;
        key[k] = 0;
        if (i < s.W * s.H) {
            float p[3];
            if (fetch_point(s, vp, fr, i, p[0], p[1], p[2])) {
                const int i0 = f2i_x86(__fsub_rn(floorf(__fmul_rn(p[0], vp->inv[0])), (float)vp->minb[0]));
                const int i1 = f2i_x86(__fsub_rn(floorf(__fmul_rn(p[1], vp->inv[1])), (float)vp->minb[1]));
                const int i2 = f2i_x86(__fsub_rn(floorf(__fmul_rn(p[2], vp->inv[2])), (float)vp->minb[2]));
                key[k] = (uint32_t)i0 * (uint32_t)vp->mul[0] + (uint32_t)i1 * (uint32_t)vp->mul[1] + (uint32_t)i2 * (uint32_t)vp->mul[2];
                ok = true;
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        rank[k] = __popc(b & ((1u << lane) - 1));
        if (ok) okmask |= 1u << k;
        if (lane == 0) cnt[k * 8 + warp] = __popc(b);
    }
    __syncthreads();
    uint32_t mine = 0;   // exclusive prefix over the (k, warp) cells = ascending pixel order
    if (threadIdx.x < VOX_PPT * 8)
        for (int j = 0; j < (int)threadIdx.x; j++) mine += cnt[j];
    __syncthreads();
    if (threadIdx.x < VOX_PPT * 8) cnt[threadIdx.x] = mine;
    __syncthreads();
    const uint32_t base = blk_off[(size_t)fr * gridDim.x + blockIdx.x];
#pragma unroll
    for (int k = 0; k < VOX_PPT; k++)
        if (okmask >> k & 1) {
            const uint32_t pos = base + cnt[k * 8 + warp] + rank[k];
            keys[pos] = key[k];
            vals[pos] = (uint32_t)((blockIdx.x * VOX_PPT + k) * blockDim.x + threadIdx.x);
        }
}

// ---- stable LSD radix sort, 8-bit digits ---------------------------------------------------
constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ keys, const VoxParams* __restrict__ vp0, int npad, int shift,
                                                        uint32_t* __restrict__ hist, int nblocks) {
    __shared__ uint32_t h[256];
    const int n = (int)vp0[blockIdx.y].nsort;   // the frame's compacted points; tiles beyond them only report zeros
    keys += (size_t)blockIdx.y * npad;
    hist += (size_t)blockIdx.y * 256 * nblocks;
    const int base = blockIdx.x * RS_TILE;
    if (base >= n) {
        hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = 0;
        return;
    }
    h[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const int i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of m uint32 values by one block (digit-major histogram, or head flags)
__global__ void __launch_bounds__(1024) k_scan_block(uint32_t* data, int m, uint32_t* total, size_t total_stride_bytes) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    data += (size_t)blockIdx.x * m;
    if (total) total = (uint32_t*)((uint8_t*)total + (size_t)blockIdx.x * total_stride_bytes);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < m ? data[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            warp_sums[threadIdx.x] = wi - w;  // exclusive
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
        if (i < m) data[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// Device-wide exclusive scan in three small launches: per-tile sums, scan of the sums (one block), tile rescan.
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 16, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums, uint32_t& block_total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t w = threadIdx.x < SCAN_THREADS / 32 ? warp_sums[threadIdx.x] : 0;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= o) wi += t;
        }
        if (threadIdx.x < SCAN_THREADS / 32) warp_sums[threadIdx.x] = wi - w;
        if (threadIdx.x == SCAN_THREADS / 32 - 1) warp_sums[32] = wi;
    }
    __syncthreads();
    block_total = warp_sums[32];
    return warp_sums[threadIdx.x >> 5] + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ data, int m, uint32_t* __restrict__ sums) {
    __shared__ uint32_t ws[33];
    data += (size_t)blockIdx.y * m;
    sums += (size_t)blockIdx.y * gridDim.x;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < m) v += data[base + k];
    uint32_t total;
    block_excl_scan(v, ws, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_apply(uint32_t* __restrict__ data, int m, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t ws[33];
    data += (size_t)blockIdx.y * m;
    sums += (size_t)blockIdx.y * gridDim.x;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t x[SCAN_ITEMS], v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        x[k] = base + k < m ? data[base + k] : 0;
        v += x[k];
    }
    uint32_t total;
    uint32_t run = block_excl_scan(v, ws, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < m) data[base + k] = run;
        run += x[k];
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           const VoxParams* __restrict__ vp0, int npad, int shift,
                                                           const uint32_t* __restrict__ hist, int nblocks, uint32_t* __restrict__ okeys,
                                                           uint32_t* __restrict__ ovals) {
    constexpr int NW = RS_THREADS / 32;
    __shared__ uint32_t wcount[NW][256];
    const int n = (int)vp0[blockIdx.y].nsort;
    if (blockIdx.x * RS_TILE >= n) return;
    keys += (size_t)blockIdx.y * npad; vals += (size_t)blockIdx.y * npad;
    okeys += (size_t)blockIdx.y * npad; ovals += (size_t)blockIdx.y * npad;
    hist += (size_t)blockIdx.y * 256 * nblocks;
    for (int i = threadIdx.x; i < NW * 256; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp w owns the contiguous sub-tile [w*256, (w+1)*256) of the block's tile: order = (block, warp, round, lane)
    const int base = blockIdx.x * RS_TILE + warp * (RS_ITEMS * 32);
    uint32_t k[RS_ITEMS], v[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int i = base + r * 32 + lane;
        const bool in = i < n;
        k[r] = in ? keys[i] : 0xFFFFFFFFu;
        v[r] = in ? vals[i] : 0;
        const uint32_t d = (k[r] >> shift) & 255u;
        const unsigned act = __ballot_sync(0xffffffffu, in);
        rank[r] = 0;
        if (in) {
            const unsigned peers = __match_any_sync(act, d);
            const uint32_t before = wcount[warp][d];
            rank[r] = before + __popc(peers & ((1u << lane) - 1));
            __syncwarp(act);
            if (lane == __ffs(peers) - 1) wcount[warp][d] = before + __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps + global base of this block
    {
        const int d = threadIdx.x;  // RS_THREADS == 256 digits
        uint32_t run = hist[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (k[r] >> shift) & 255u;
            const uint32_t pos = wcount[warp][d] + rank[r];
            okeys[pos] = k[r];
            ovals[pos] = v[r];
        }
    }
}

// ---- run detection and reduction -------------------------------------------------------------
// first output point of frame fr in the packed per-batch output arrays (every count is final before the reduction starts)
__device__ __forceinline__ size_t cloud_base(const VoxParams* vp0, int fr, size_t npix) {
    size_t base = 0;
    for (int g = 0; g < fr; g++) base += vp0[g].overflow ? npix : (size_t)vp0[g].nvox;
    return base;
}

__global__ void __launch_bounds__(256) k_vox_heads(const uint32_t* __restrict__ keys, const VoxParams* __restrict__ vp0, int npad,
                                                   uint32_t* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const int n = (int)vp0[blockIdx.y].nsort;
    keys += (size_t)blockIdx.y * npad;
    flags += (size_t)blockIdx.y * npad;
    if (i >= n) {
        flags[i] = 0;
        return;
    }
    const uint32_t k = keys[i];
    flags[i] = (i == 0 || keys[i - 1] != k) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_vox_reduce(Source s, VoxParams* vp0, const uint32_t* __restrict__ keys,
                                                    const uint32_t* __restrict__ vals, const uint32_t* __restrict__ slot, int npad,
                                                    float* __restrict__ oxyz, uint8_t* __restrict__ orgb, uint32_t* __restrict__ oidx,
                                                    uint32_t* __restrict__ ocnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, fr = blockIdx.y;
    VoxParams* vp = vp0 + fr;
    const int n = (int)vp->nsort;
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    keys += (size_t)fr * npad; vals += (size_t)fr * npad; slot += (size_t)fr * npad;
    {   // the clouds of a batch are packed back to back (frame fr starts where frame fr-1 ends): one copy brings them all home
        const size_t base = cloud_base(vp0, fr, (size_t)s.W * s.H);
        oxyz += base * 3; orgb += base * 3; oidx += base; ocnt += base;
    }
    // every thread fetches the point at its own sorted position (parallel, coalesced key/val reads); the run heads then
    // add their run in sorted order (= ascending point index) out of shared memory, exactly the order of the sequential sum
    __shared__ float pxyz[256][3];
    __shared__ float prgb[256][3];
    __shared__ uint32_t pkey[256];
    const bool valid = i < n;          // (every compacted key is a real voxel index, 0xFFFFFFFF included when an absurd grid wraps)
    const uint32_t key = valid ? keys[i] : 0u;
    {
        float X = 0.f, Y = 0.f, Z = 0.f, R = 0.f, G = 0.f, B = 0.f;
        if (valid) {
            const int pt = (int)vals[i];
            fetch_point(s, vp, fr, pt, X, Y, Z);
            if (s.bgr) {
                const int y = pt / s.W, x = pt - y * s.W;
                const uint8_t* c = s.bgr + (size_t)fr * s.cframe + (size_t)y * s.cstep + (size_t)x * 3;
                R = (float)c[2]; G = (float)c[1]; B = (float)c[0];
            }
        }
        pxyz[threadIdx.x][0] = X; pxyz[threadIdx.x][1] = Y; pxyz[threadIdx.x][2] = Z;
        prgb[threadIdx.x][0] = R; prgb[threadIdx.x][1] = G; prgb[threadIdx.x][2] = B;
        pkey[threadIdx.x] = key;
    }
    __syncthreads();
    if (!valid || (i > 0 && keys[i - 1] == key)) return;
    float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
    int cnt = 0;
    int j = i;
    for (int t = threadIdx.x; t < 256 && j < n && pkey[t] == key; t++, j++) {   // the part of the run inside this block
        sx = __fadd_rn(sx, pxyz[t][0]); sy = __fadd_rn(sy, pxyz[t][1]); sz = __fadd_rn(sz, pxyz[t][2]);
        sr = __fadd_rn(sr, prgb[t][0]); sg = __fadd_rn(sg, prgb[t][1]); sb = __fadd_rn(sb, prgb[t][2]);
        cnt++;
    }
    for (; j < n && keys[j] == key; j++) {   // a run that continues past the block's last position
        const int pt = (int)vals[j];
        float X, Y, Z;
        fetch_point(s, vp, fr, pt, X, Y, Z);
        sx = __fadd_rn(sx, X); sy = __fadd_rn(sy, Y); sz = __fadd_rn(sz, Z);
        if (s.bgr) {
            const int y = pt / s.W, x = pt - y * s.W;
            const uint8_t* c = s.bgr + (size_t)fr * s.cframe + (size_t)y * s.cstep + (size_t)x * 3;
            sr = __fadd_rn(sr, (float)c[2]); sg = __fadd_rn(sg, (float)c[1]); sb = __fadd_rn(sb, (float)c[0]);
        }
        cnt++;
    }
    const uint32_t o = slot[i];
    const float fc = (float)cnt;
    oxyz[(size_t)o * 3] = __fdiv_rn(sx, fc);
    oxyz[(size_t)o * 3 + 1] = __fdiv_rn(sy, fc);
    oxyz[(size_t)o * 3 + 2] = __fdiv_rn(sz, fc);
    orgb[(size_t)o * 3] = (uint8_t)(uint32_t)__fdiv_rn(sr, fc);
    orgb[(size_t)o * 3 + 1] = (uint8_t)(uint32_t)__fdiv_rn(sg, fc);
    orgb[(size_t)o * 3 + 2] = (uint8_t)(uint32_t)__fdiv_rn(sb, fc);
    oidx[o] = key;
    ocnt[o] = (uint32_t)cnt;
}

// PCL passthrough ("Leaf size is too small"): output = input cloud, organised, NaN where non-finite.
__global__ void __launch_bounds__(256) k_vox_passthrough(Source s, const VoxParams* vp0, float* __restrict__ oxyz, uint8_t* __restrict__ orgb) {
    const int fr = blockIdx.y;
    const VoxParams* vp = vp0 + fr;
    if (!vp->overflow) return;   // the usual case: a small grid of blocks that leave at once
    {
        const size_t base = cloud_base(vp0, fr, (size_t)s.W * s.H);
        oxyz += base * 3;
        orgb += base * 3;
    }
    Source t = s;
    t.drop_invalid = 0;
    t.z_max = 0.f;
    const float nan = __int_as_float(0x7FC00000);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.W * s.H; i += gridDim.x * blockDim.x) {
        float X, Y, Z;
        const bool fin = fetch_point(t, vp, fr, i, X, Y, Z);
        oxyz[(size_t)i * 3] = fin ? X : nan;
        oxyz[(size_t)i * 3 + 1] = fin ? Y : nan;
        oxyz[(size_t)i * 3 + 2] = fin ? Z : nan;
        uint8_t r = 0, g = 0, b = 0;
        if (fin && s.bgr) {
            const int y = i / s.W, x = i - y * s.W;
            const uint8_t* c = s.bgr + (size_t)fr * s.cframe + (size_t)y * s.cstep + (size_t)x * 3;
            r = c[2]; g = c[1]; b = c[0];
        }
        orgb[(size_t)i * 3] = r; orgb[(size_t)i * 3 + 1] = g; orgb[(size_t)i * 3 + 2] = b;
    }
}

}  // namespace

struct sdr_cloud {
    sdr_ctx* ctx = nullptr;
    DevBuf in_points, in_bgr, in_disp;
    DevBuf keys[2], vals[2], hist, flags, params, scan_sums, blk_cnt;
    DevBuf oxyz, orgb, oidx, ocnt;
    PinnedBuf host_params;
    cudaEvent_t done = nullptr;
    int frames = 0;                 // frames of the last enqueue
    std::vector<int64_t> last_n, last_off;   // points of frame f and where they start in the packed output arrays
    std::vector<int> last_passthrough;
    int W = 0, H = 0;
};

extern "C" int sdr_cloud_create(sdr_ctx* ctx, sdr_cloud** out) {
    if (!ctx || !out) return SDR_EINVAL;
    sdr_cloud* c = new sdr_cloud();
    c->ctx = ctx;
    *out = c;
    return SDR_OK;
}

extern "C" void sdr_cloud_destroy(sdr_cloud* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    for (DevBuf* b : {&c->in_points, &c->in_bgr, &c->in_disp, &c->keys[0], &c->keys[1], &c->vals[0], &c->vals[1], &c->hist, &c->flags,
                      &c->params, &c->scan_sums, &c->blk_cnt, &c->oxyz, &c->orgb, &c->oidx, &c->ocnt})
        b->release();
    c->host_params.release();
    if (c->done) cudaEventDestroy(c->done);
    delete c;
}

// `frames` independent exclusive scans of m values each (data[f*m ...]); totals (optional) at total + f*stride bytes.
static int scan_u32(sdr_cloud* c, int frames, uint32_t* data, int m, uint32_t* total, size_t total_stride) {
    sdr_ctx* ctx = c->ctx;
    const int nt = div_up(m, SCAN_TILE);
    int rc = c->scan_sums.reserve(ctx, (size_t)nt * frames * 4);
    if (rc) return rc;
    k_scan_tile_sums<<<dim3(nt, frames), SCAN_THREADS, 0, ctx->stream>>>(data, m, c->scan_sums.as<uint32_t>());
    SDR_LAUNCHED(ctx);
    k_scan_block<<<frames, 1024, 0, ctx->stream>>>(c->scan_sums.as<uint32_t>(), nt, total, total_stride);
    SDR_LAUNCHED(ctx);
    k_scan_tile_apply<<<dim3(nt, frames), SCAN_THREADS, 0, ctx->stream>>>(data, m, c->scan_sums.as<uint32_t>());
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

// Enqueues the voxelisation of `frames` clouds (one per frame of the batch) on the ctx stream; every kernel covers the
// whole batch. The result headers land in c->host_params.
int sdr_cloud_enqueue(sdr_cloud* c, int frames, const uint8_t* points, size_t pstep, size_t pframe, const uint8_t* disp, size_t dstep,
                      size_t dframe, const uint8_t* bgr, size_t cstep, size_t cframe, int W, int H, const double* Q, int handle_missing,
                      const sdr_voxel_options* opt) {
    sdr_ctx* ctx = c->ctx;
    cudaStream_t st = ctx->stream;
    const int n = W * H;
    const int nblocks = div_up(n, RS_TILE);
    const size_t F = (size_t)frames;
    int rc;
    for (int i = 0; i < 2; i++) {
        if ((rc = c->keys[i].reserve(ctx, F * n * 4))) return rc;
        if ((rc = c->vals[i].reserve(ctx, F * n * 4))) return rc;
    }
    if ((rc = c->hist.reserve(ctx, F * 256 * nblocks * 4))) return rc;
    if ((rc = c->flags.reserve(ctx, F * n * 4))) return rc;
    if ((rc = c->params.reserve(ctx, F * sizeof(VoxParams)))) return rc;
    if ((rc = c->blk_cnt.reserve(ctx, F * div_up(n, 256 * VOX_PPT) * 4))) return rc;
    if ((rc = c->oxyz.reserve(ctx, F * n * 12))) return rc;
    if ((rc = c->orgb.reserve(ctx, F * n * 3))) return rc;
    if ((rc = c->oidx.reserve(ctx, F * n * 4))) return rc;
    if ((rc = c->ocnt.reserve(ctx, F * n * 4))) return rc;
    if ((rc = c->host_params.reserve(ctx, F * sizeof(VoxParams)))) return rc;
    c->W = W; c->H = H; c->frames = frames;
    Source s{};
    s.points = points; s.pstep = pstep; s.pframe = pframe; s.disp = disp; s.dstep = dstep; s.dframe = dframe;
    s.bgr = bgr; s.cstep = cstep; s.cframe = cframe; s.W = W; s.H = H;
    s.from_disp = points == nullptr;
    if (Q) memcpy(s.Q.q, Q, sizeof s.Q.q);
    s.handle_missing = s.from_disp ? handle_missing : 0;
    s.drop_invalid = opt->drop_invalid && disp != nullptr;
    s.min_valid16 = opt->min_disparity * 16;
    s.z_max = opt->z_max;
    VoxParams* vp = c->params.as<VoxParams>();
    const dim3 g256(div_up(n, 256), frames);
    const dim3 gppt(div_up(n, 256 * VOX_PPT), frames);   // reduction kernels: VOX_PPT pixels per thread
    k_vox_reset<<<frames, 1, 0, st>>>(vp);
    SDR_LAUNCHED(ctx);
    if (s.from_disp && s.handle_missing) {
        k_vox_mindisp<<<gppt, 256, 0, st>>>(s, vp);
        SDR_LAUNCHED(ctx);
    }
    k_vox_bbox<<<gppt, 256, 0, st>>>(s, vp, c->blk_cnt.as<uint32_t>());
    SDR_LAUNCHED(ctx);
    k_vox_params<<<frames, 1, 0, st>>>(vp, opt->leaf[0], opt->leaf[1], opt->leaf[2]);
    SDR_LAUNCHED(ctx);
    k_scan_block<<<frames, 1024, 0, st>>>(c->blk_cnt.as<uint32_t>(), (int)gppt.x, nullptr, 0);   // where each block's points go
    SDR_LAUNCHED(ctx);
    k_vox_keys<<<gppt, 256, 0, st>>>(s, vp, c->blk_cnt.as<uint32_t>(), c->keys[0].as<uint32_t>(), c->vals[0].as<uint32_t>(), n);
    SDR_LAUNCHED(ctx);
    int cur = 0;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = pass * 8;
        k_rs_hist<<<dim3(nblocks, frames), RS_THREADS, 0, st>>>(c->keys[cur].as<uint32_t>(), vp, n, shift, c->hist.as<uint32_t>(), nblocks);
        SDR_LAUNCHED(ctx);
        if ((rc = scan_u32(c, frames, c->hist.as<uint32_t>(), 256 * nblocks, nullptr, 0))) return rc;
        k_rs_scatter<<<dim3(nblocks, frames), RS_THREADS, 0, st>>>(c->keys[cur].as<uint32_t>(), c->vals[cur].as<uint32_t>(), vp, n, shift,
                                                                   c->hist.as<uint32_t>(), nblocks, c->keys[cur ^ 1].as<uint32_t>(),
                                                                   c->vals[cur ^ 1].as<uint32_t>());
        SDR_LAUNCHED(ctx);
        cur ^= 1;
    }
    k_vox_heads<<<g256, 256, 0, st>>>(c->keys[cur].as<uint32_t>(), vp, n, c->flags.as<uint32_t>());
    SDR_LAUNCHED(ctx);
    if ((rc = scan_u32(c, frames, c->flags.as<uint32_t>(), n, &vp->nvox, sizeof(VoxParams)))) return rc;
    k_vox_reduce<<<g256, 256, 0, st>>>(s, vp, c->keys[cur].as<uint32_t>(), c->vals[cur].as<uint32_t>(), c->flags.as<uint32_t>(), n,
                                       c->oxyz.as<float>(), c->orgb.as<uint8_t>(), c->oidx.as<uint32_t>(), c->ocnt.as<uint32_t>());
    SDR_LAUNCHED(ctx);
    k_vox_passthrough<<<dim3(std::min(div_up(n, 256), 4 * ctx->sm_count), frames), 256, 0, st>>>(s, vp, c->oxyz.as<float>(), c->orgb.as<uint8_t>());
    SDR_LAUNCHED(ctx);
    SDR_CUDA(ctx, cudaMemcpyAsync(c->host_params.p, vp, F * sizeof(VoxParams), cudaMemcpyDeviceToHost, st));
    if (!c->done) SDR_CUDA(ctx, cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming | cudaEventBlockingSync));
    SDR_CUDA(ctx, cudaEventRecord(c->done, st));
    return SDR_OK;
}

// Waits for the enqueued work and fills the views (one per frame). Returns SDR_OK or SDR_EVOXEL_OVERFLOW.
int sdr_cloud_finish(sdr_cloud* c, sdr_cloud_view* views) {
    sdr_ctx* ctx = c->ctx;
    SDR_CUDA(ctx, cudaEventSynchronize(c->done));
    const VoxParams* hp = (const VoxParams*)c->host_params.p;
    const size_t n = (size_t)c->W * c->H;
    c->last_n.assign(c->frames, 0);
    c->last_off.assign(c->frames + 1, 0);
    c->last_passthrough.assign(c->frames, 0);
    int status = SDR_OK;
    for (int f = 0; f < c->frames; f++) {
        c->last_passthrough[f] = hp[f].overflow;
        c->last_n[f] = hp[f].overflow ? (int64_t)n : (int64_t)hp[f].nvox;
        c->last_off[f + 1] = c->last_off[f] + c->last_n[f];
        if (views) {
            sdr_cloud_view* view = views + f;
            const size_t off = (size_t)c->last_off[f];
            view->n = c->last_n[f];
            view->xyz = c->oxyz.as<float>() + off * 3;
            view->rgb = c->orgb.as<uint8_t>() + off * 3;
            view->idx = hp[f].overflow ? nullptr : c->oidx.as<uint32_t>() + off;
            view->cnt = hp[f].overflow ? nullptr : c->ocnt.as<uint32_t>() + off;
            for (int i = 0; i < 3; i++) { view->min_b[i] = hp[f].minb[i]; view->div_b[i] = hp[f].divb[i]; }
            view->passthrough = hp[f].overflow;
        }
        if (hp[f].overflow) status = SDR_EVOXEL_OVERFLOW;
    }
    if (status == SDR_EVOXEL_OVERFLOW)
        sdr_set_error(ctx, "[VoxelGrid] Leaf size is too small for the input dataset. Integer indices would overflow.");
    return status;
}

int sdr_cloud_download_frame(sdr_cloud* c, int f, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt, bool wait) {
    sdr_ctx* ctx = c->ctx;
    if (f < 0 || f >= c->frames || f >= (int)c->last_n.size()) return SDR_FAIL(ctx, SDR_EINVAL, "cloud_download: no such frame %d", f);
    const size_t n = (size_t)c->last_n[f], off = (size_t)c->last_off[f];
    if (xyz) SDR_CUDA(ctx, cudaMemcpyAsync(xyz, c->oxyz.as<float>() + off * 3, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (rgb) SDR_CUDA(ctx, cudaMemcpyAsync(rgb, c->orgb.as<uint8_t>() + off * 3, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    if (!c->last_passthrough[f]) {
        if (idx) SDR_CUDA(ctx, cudaMemcpyAsync(idx, c->oidx.as<uint32_t>() + off, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (cnt) SDR_CUDA(ctx, cudaMemcpyAsync(cnt, c->ocnt.as<uint32_t>() + off, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (wait) SDR_CUDA(ctx, sdr_wait(ctx));
    return SDR_OK;
}

// All clouds of the last batch, packed back to back as they sit on the device: ONE copy per array. offsets[frames + 1]
// (in points) says where each frame's cloud starts. idx / cnt are meaningless for passthrough frames.
int sdr_cloud_download_packed(sdr_cloud* c, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt, int64_t* offsets, bool wait) {
    sdr_ctx* ctx = c->ctx;
    if ((int)c->last_off.size() != c->frames + 1) return SDR_FAIL(ctx, SDR_EINVAL, "cloud_download_packed: no finished batch");
    const size_t total = (size_t)c->last_off[c->frames];
    if (offsets) memcpy(offsets, c->last_off.data(), sizeof(int64_t) * (c->frames + 1));
    if (xyz && total) SDR_CUDA(ctx, cudaMemcpyAsync(xyz, c->oxyz.p, total * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (rgb && total) SDR_CUDA(ctx, cudaMemcpyAsync(rgb, c->orgb.p, total * 3, cudaMemcpyDeviceToHost, ctx->stream));
    if (idx && total) SDR_CUDA(ctx, cudaMemcpyAsync(idx, c->oidx.p, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (cnt && total) SDR_CUDA(ctx, cudaMemcpyAsync(cnt, c->ocnt.p, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (wait) SDR_CUDA(ctx, sdr_wait(ctx));
    return SDR_OK;
}

static int check_opt(sdr_ctx* ctx, const sdr_voxel_options* opt) {
    if (!opt || !(opt->leaf[0] > 0.f) || !(opt->leaf[1] > 0.f) || !(opt->leaf[2] > 0.f))
        return SDR_FAIL(ctx, SDR_EINVAL, "voxel leaf sizes must be positive");
    return SDR_OK;
}

extern "C" int sdr_voxelgrid(sdr_cloud* c, const sdr_image* points, const sdr_image* bgr, const sdr_image* disp16,
                             const sdr_voxel_options* opt, sdr_cloud_view* view) {
    if (!c || !points) return SDR_EINVAL;
    sdr_ctx* ctx = c->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_opt(ctx, opt))) return rc;
    if ((rc = sdr_check_image(ctx, points, 3, 4, "points"))) return rc;
    const int W = points->cols, H = points->rows;
    DevView vp, vc{}, vd{};
    if ((rc = sdr_stage_in(ctx, points, c->in_points, &vp))) return rc;
    if (bgr) {
        if ((rc = sdr_check_image(ctx, bgr, 3, 1, "bgr"))) return rc;
        if (bgr->cols != W || bgr->rows != H) return SDR_FAIL(ctx, SDR_ESIZE, "voxelgrid: colour image size differs");  // pcd_write.cpp:25
        if ((rc = sdr_stage_in(ctx, bgr, c->in_bgr, &vc))) return rc;
    }
    if (disp16) {
        if ((rc = sdr_check_image(ctx, disp16, 1, 2, "disp16"))) return rc;
        if (disp16->cols != W || disp16->rows != H) return SDR_FAIL(ctx, SDR_ESIZE, "voxelgrid: disparity size differs");
        if ((rc = sdr_stage_in(ctx, disp16, c->in_disp, &vd))) return rc;
    } else if (opt->drop_invalid) {
        return SDR_FAIL(ctx, SDR_EINVAL, "voxelgrid: drop_invalid needs the CV_16S disparity");
    }
    if ((rc = sdr_cloud_enqueue(c, 1, vp.p, vp.pitch, 0, vd.p, vd.pitch, 0, vc.p, vc.pitch, 0, W, H, nullptr, 0, opt))) return rc;
    return sdr_cloud_finish(c, view);
}

extern "C" int sdr_depth_cloud(sdr_cloud* c, const sdr_image* disp16, const sdr_image* bgr, const double* Q, int handle_missing,
                               const sdr_voxel_options* opt, sdr_cloud_view* view) {
    if (!c || !disp16 || !Q) return SDR_EINVAL;
    sdr_ctx* ctx = c->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = check_opt(ctx, opt))) return rc;
    if ((rc = sdr_check_image(ctx, disp16, 1, 2, "disp16"))) return rc;
    const int W = disp16->cols, H = disp16->rows;
    DevView vc{}, vd;
    if ((rc = sdr_stage_in(ctx, disp16, c->in_disp, &vd))) return rc;
    if (bgr) {
        if ((rc = sdr_check_image(ctx, bgr, 3, 1, "bgr"))) return rc;
        if (bgr->cols != W || bgr->rows != H) return SDR_FAIL(ctx, SDR_ESIZE, "depth_cloud: colour image size differs");
        if ((rc = sdr_stage_in(ctx, bgr, c->in_bgr, &vc))) return rc;
    }
    if ((rc = sdr_cloud_enqueue(c, 1, nullptr, 0, 0, vd.p, vd.pitch, 0, vc.p, vc.pitch, 0, W, H, Q, handle_missing, opt))) return rc;
    return sdr_cloud_finish(c, view);
}

extern "C" int sdr_cloud_download(sdr_cloud* c, float* xyz, uint8_t* rgb, uint32_t* idx, uint32_t* cnt) {
    if (!c) return SDR_EINVAL;
    SDR_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    return sdr_cloud_download_frame(c, 0, xyz, rgb, idx, cnt, true);
}

extern "C" int sdr_cloud_save_pcd_binary(sdr_cloud* c, int frame, const char* path) {
    if (!c || !path) return SDR_EINVAL;
    sdr_ctx* ctx = c->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (frame < 0 || frame >= c->frames || frame >= (int)c->last_n.size()) return SDR_FAIL(ctx, SDR_EINVAL, "save_pcd: no such frame %d", frame);
    const size_t n = (size_t)c->last_n[frame];
    if (n == 0) return SDR_FAIL(ctx, SDR_EINVAL, "No points to save after downsampling!");   // pcd_write.cpp:143-146
    std::vector<float> xyz(n * 3);
    std::vector<uint8_t> rgb(n * 3);
    int rc = sdr_cloud_download_frame(c, frame, xyz.data(), rgb.data(), nullptr, nullptr, true);
    if (rc) return rc;
    const bool organised = c->last_passthrough[frame] != 0;
    const size_t width = organised ? (size_t)c->W : n, height = organised ? (size_t)c->H : 1;
    FILE* f = fopen(path, "wb");
    if (!f) return SDR_FAIL(ctx, SDR_EIO, "save_pcd: cannot open %s", path);
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n"
               "WIDTH %zu\nHEIGHT %zu\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", width, height, n);
    std::vector<uint8_t> rec(n * 16);
    for (size_t i = 0; i < n; i++) {
        memcpy(&rec[i * 16], &xyz[i * 3], 12);
        const uint32_t packed = 0xFF000000u | ((uint32_t)rgb[i * 3] << 16) | ((uint32_t)rgb[i * 3 + 1] << 8) | (uint32_t)rgb[i * 3 + 2];
        memcpy(&rec[i * 16 + 12], &packed, 4);
    }
    const bool ok = fwrite(rec.data(), 1, rec.size(), f) == rec.size();
    fclose(f);
    return ok ? SDR_OK : SDR_FAIL(ctx, SDR_EIO, "save_pcd: short write to %s", path);
}
