// Semi-global block matching for sm_100a, bit-exact with cv::StereoSGBM::compute
// (the call at stereo_vision/src/stereo_disparity.cpp:27 and point_cloud/src/pcd_write.cpp:111
// of the reference). Stages:
//   k_prefilter   Sobel-x clip + raw channel, Birchfield-Tomasi half-sample intervals
//   k_cost        BT pixel cost -> (2*SH2+1)x(2*SW2+1) box sum -> C volume [rows][W1][D] int16
//   k_path_v      vertical / diagonal path aggregation (chains wrap around the row so every
//                 chain is H steps long), S (+)= L
//   k_path_h      horizontal paths; the last one keeps S in registers and runs WTA,
//                 uniqueness and the sub-pixel parabola
//   k_lrcheck     right-view disparity by keyed atomicMin (cost, then largest x) + LR check
// All cost arithmetic is packed int16x2 on the DPX/video integer units (VIADD.16x2,
// VIMNMX.S16x2, VIMNMX3, VIADDMNMX[.RELU]); a chain owns D disparities spread over G lanes.
#include "sdr_internal.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

struct Seg {
    int src_start, src_end, dst_start, crow0;
    int shift;   // MODE_SGBM_3WAY on tiny frames: source row y of this stripe lands in output row y - shift (see make_geom)
};

struct Geom {
    int W, H, D, minD, W1, minX1, maxX1;
    int nseg, crows;
    int hrow0, hrows;   // output rows of the selected stripes [hrow0, hrow0 + hrows): what the horizontal kernels cover
    Seg seg[4];
    int SW2, SH2, P1, P2, ftzero, uq, d12, three_way;
    int tie8;           // MODE_SGBM_3WAY with uniquenessRatio 0: OpenCV's 8-lane minimum search decides ties (wta_pixel)
    uint32_t uq_magic;  // ceil(2^32 / (100 - uq)): exact n / (100 - uq) for n < 2^32 / 100 by one __umulhi
    int Wp;  // padded width of the right-image byte planes (multiple of 16); PR = [b][y][Wp/16][24 plane-copies][16 B]
};

// One 16-byte chunk of the reversed right-image planes: 24 plane-copies x 16 B, padded to 400 B so that consecutive
// chunks start 4 banks apart in shared memory (the cost kernel reads the same byte offset of 8 consecutive chunks).
constexpr int PR_CHUNK = 400;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
// prmt.b32 with the full 4-bit selectors (msb of a nibble = replicate the sign bit of the selected byte), which
// __byte_perm does not expose
__device__ __forceinline__ uint32_t prmt_full(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ uint32_t dup_lo(uint32_t a) { return __byte_perm(a, a, 0x1010); }
__device__ __forceinline__ uint32_t dup_hi(uint32_t a) { return __byte_perm(a, a, 0x3232); }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// ------------------------------------------------------------------------------------------
// Prefilter. LEFT: one int4 per pixel = packed int16 {-u, u-255, -u1, u0-255} for the Sobel
// channel (.x,.y) and the raw channel (.z,.w). RIGHT: byte planes (v, lo, 255-hi) x 2 channels
// stored reversed (index W-1-x, so that disparity runs forward in memory) in 4 copies shifted by
// 0..3 bytes so that any run of 4 disparities is one aligned 32-bit load.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefilt_px(const uint8_t* img, size_t pitch, int W, int H, int x, int y, int ft,
                                           int& s, int& r) {
    if (x <= 0 || x >= W - 1) {
        s = ft;
        r = ft;
        return;
    }
    const uint8_t* r0 = img + (size_t)y * pitch;
    const uint8_t* rn = img + (size_t)max(y - 1, 0) * pitch;
    const uint8_t* rs = img + (size_t)min(y + 1, H - 1) * pitch;
    int v = 2 * ((int)r0[x + 1] - (int)r0[x - 1]) + (int)rn[x + 1] - (int)rn[x - 1] + (int)rs[x + 1] - (int)rs[x - 1];
    s = min(max(v, -ft), ft) + ft;
    r = r0[x];
}

// One block = 128 positions of one row of one eye. Every pixel's (Sobel, raw) pair is computed once into shared memory
// (134 pixels: the block's range plus the neighbours the half-sample intervals and the shifted copies reach).
// LEFT: thread t writes pixel x0 + t. RIGHT: the block owns 128 consecutive indices i of the reversed planes and writes
// each of the 24 plane-copies as 32 aligned words (copy `sft` holds plane[i + sft], assembled from the byte planes in
// shared memory with a funnel shift) instead of 24 single bytes per pixel.
__global__ void __launch_bounds__(128) k_prefilter(Geom g, const uint8_t* left, size_t lpitch, size_t lframe,
                                                   const uint8_t* right, size_t rpitch, size_t rframe, int4* PL,
                                                   uint8_t* PR, int y0) {
    __shared__ int16_t sS[136], sR[136];               // (Sobel, raw) of pixels xs0 .. xs0 + 133 (0 outside the image)
    __shared__ __align__(4) uint8_t pb[6][136];        // RIGHT: byte planes (v, lo, 255-hi) x 2 channels at j = I0 .. I0 + 130
    const int t = threadIdx.x, y = y0 + blockIdx.y;
    const int b = blockIdx.z >> 1, eye = blockIdx.z & 1;
    const uint8_t* img = eye ? right + (size_t)b * rframe : left + (size_t)b * lframe;
    const size_t pitch = eye ? rpitch : lpitch;
    const int W = g.W;
    // pixel range needed, ascending x. LEFT: x0-1 .. x0+128. RIGHT: j = I0-1 .. I0+131  <=>  x = W-1-j
    const int I0 = blockIdx.x * 128;
    const int xs0 = eye ? (W - 1 - (I0 + 131)) - 1 : I0 - 1;
    const int ns = eye ? 134 : 130;
    for (int k = t; k < ns; k += 128) {
        const int xx = xs0 + k;
        int sv = 0, rv = 0;
        if (xx >= 0 && xx < W) prefilt_px(img, pitch, W, g.H, xx, y, g.ftzero, sv, rv);
        sS[k] = (int16_t)sv;
        sR[k] = (int16_t)rv;
    }
    __syncthreads();
    auto planes = [&](int x, int (&v)[2], int (&lo)[2], int (&hi)[2]) {   // BT half-sample interval at pixel x (inside the image)
        const int k = x - xs0;
        const int s0 = sS[k - 1], s1 = sS[k], s2 = sS[k + 1], r0 = sR[k - 1], r1 = sR[k], r2 = sR[k + 1];
        v[0] = s1; v[1] = r1;
        int vl = x > 0 ? (s1 + s0) / 2 : s1, vr = x < W - 1 ? (s1 + s2) / 2 : s1;
        lo[0] = min(min(vl, vr), s1); hi[0] = max(max(vl, vr), s1);
        vl = x > 0 ? (r1 + r0) / 2 : r1; vr = x < W - 1 ? (r1 + r2) / 2 : r1;
        lo[1] = min(min(vl, vr), r1); hi[1] = max(max(vl, vr), r1);
    };
    if (eye == 0) {
        const int x = I0 + t;
        if (x < W) {
            int v[2], lo[2], hi[2];
            planes(x, v, lo, hi);
            int4 o;
            o.x = (int)pack16(-v[0], v[0] - 255);
            o.y = (int)pack16(-hi[0], lo[0] - 255);
            o.z = (int)pack16(-v[1], v[1] - 255);
            o.w = (int)pack16(-hi[1], lo[1] - 255);
            PL[((size_t)b * g.H + y) * W + x] = o;
        }
        return;
    }
    for (int k = t; k < 132; k += 128) {   // plane bytes at j = I0 + k (k < 131 used; 131 pads the last word)
        const int x = W - 1 - (I0 + k);
        int v[2] = {0, 0}, lo[2] = {0, 0}, hi[2] = {255, 255};
        if (k < 131 && x >= 0 && x < W) planes(x, v, lo, hi);
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            pb[ch * 3 + 0][k] = (uint8_t)v[ch];
            pb[ch * 3 + 1][k] = (uint8_t)lo[ch];
            pb[ch * 3 + 2][k] = (uint8_t)(255 - hi[ch]);
        }
    }
    __syncthreads();
    uint8_t* row = PR + ((size_t)b * g.H + y) * (g.Wp >> 4) * PR_CHUNK;   // chunk (i/16): 16 bytes of each of the 24 plane-copies
    const int w = t & 31, i = I0 + 4 * w;
    if (i < g.Wp) {
        uint8_t* dst = row + (size_t)(i >> 4) * PR_CHUNK + (i & 15);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int pc = (t >> 5) * 6 + k;          // plane-copy index = plane * 4 + shift
            const int pl = pc >> 2, sft = pc & 3;
            const uint32_t* src = (const uint32_t*)pb[pl];
            const uint32_t val = __funnelshift_r(src[w], src[w + 1], 8 * sft);   // bytes plane[4w + sft .. 4w + sft + 3]
            *(uint32_t*)(dst + pc * 16) = val;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Cost volume. One CTA = TX columns (cost space) x RB rows of one segment. Per image row the
// CTA evaluates the BT cost for TX+2*SW2 columns x D disparities into a shared-memory ring of
// 2*SH2+1 rows (u8), sums the ring vertically into Vs (u16), then horizontally into C.
// ------------------------------------------------------------------------------------------
constexpr int COST_TXH = 64;   // columns per CTA including the 2*SW2 halo
constexpr int COST_RB = 90;    // output rows per CTA (2*SH2 warm-up rows are recomputed per band)

// ---- TMA (bulk async copy) + mbarrier helpers: one elected thread moves a row tile global -> shared, the CTA waits on
// the barrier's transaction count ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SDR_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"   // suspend-time hint: sleep in hardware until the phase completes
        "@p bra SDR_MBAR_DONE;\n"
        "bra SDR_MBAR_WAIT;\n"
        "SDR_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// BT cost of 8 consecutive disparities at one pixel (both channels) -> two u8x4 words. pl: the pixel's packed left
// constants; pr: staged right-image chunks (16 B of each of the 24 plane-copies per chunk); o0 / o1: byte offsets of the
// two aligned words inside the stage for plane 0, copy a (plane p adds 64 p).
__device__ __forceinline__ void bt_item(const int4 k, const uint8_t* __restrict__ pr, int o0, int o1, uint32_t& out0, uint32_t& out1) {
    uint32_t acc[4];
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
        const uint32_t kx = ch ? (uint32_t)k.z : (uint32_t)k.x, ky = ch ? (uint32_t)k.w : (uint32_t)k.y;
        const uint32_t negU = dup_lo(kx), Um = dup_hi(kx), negU1 = dup_lo(ky), U0m = dup_hi(ky);
        const int pv = (ch * 3 + 0) * 64, pl = (ch * 3 + 1) * 64, ph = (ch * 3 + 2) * 64;
        const uint32_t wv[2] = {*(const uint32_t*)(pr + o0 + pv), *(const uint32_t*)(pr + o1 + pv)};
        const uint32_t wl[2] = {*(const uint32_t*)(pr + o0 + pl), *(const uint32_t*)(pr + o1 + pl)};
        const uint32_t wh[2] = {*(const uint32_t*)(pr + o0 + ph), *(const uint32_t*)(pr + o1 + ph)};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t sel = (i & 1) ? 0x4342u : 0x4140u;
            const uint32_t V = prmt(wv[i >> 1], 0, sel), V0 = prmt(wl[i >> 1], 0, sel), CV1 = prmt(wh[i >> 1], 0, sel);
            const uint32_t a_ = __vadd2(Um, CV1);                      // u - v1
            const uint32_t c0 = __viaddmax_s16x2_relu(V0, negU, a_);   // max(v0 - u, u - v1, 0)
            const uint32_t b_ = __vadd2(U0m, V ^ 0x00FF00FFu);         // u0 - v
            const uint32_t c1 = __viaddmax_s16x2_relu(V, negU1, b_);   // max(v - u1, u0 - v, 0)
            const uint32_t c = __vmins2(c0, c1);
            if (ch == 0)
                acc[i] = c;
            else
                acc[i] = __vadd2(acc[i], (c >> 2) & 0x3FFF3FFFu);
        }
    }
    out0 = prmt(acc[0], acc[1], 0x6420);
    out1 = prmt(acc[2], acc[3], 0x6420);
}

constexpr int COST_STAGES = 2;

// blockDim.x is a multiple of DQ = D/8: thread = (q, xg); it owns ITEMS columns xh = xg + k*XG of the 64-column tile for
// the whole band, keeps their running vertical sums in registers and their last 2*SH2+1 pixel-cost rows in a
// thread-private shared-memory ring (no synchronisation needed for it). The prefiltered image rows are staged in
// shared memory by TMA bulk copies (two per row: the left constants and the right byte planes of the tile), two rows
// ahead, completion signalled on an mbarrier.
template <int ITEMS, int TXH>   // TXH: columns per CTA including the 2*SW2 halo (64; 32 when the 64-column tile does not fit shared memory)
__global__ void __launch_bounds__(256) k_cost(Geom g, const int4* __restrict__ PL, const uint8_t* __restrict__ PR,
                                              int16_t* __restrict__ C) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[COST_STAGES];
    const int D = g.D, DQ = D >> 3;
    const int n = 2 * g.SH2 + 1, TX = TXH - 2 * g.SW2;
    const int b = blockIdx.z;
    int band = blockIdx.y, si = 0;
    for (; si < g.nseg; si++) {
        const int nb = (g.seg[si].src_end - g.seg[si].src_start + COST_RB - 1) / COST_RB;
        if (band < nb) break;
        band -= nb;
    }
    if (si >= g.nseg) return;
    const Seg sg = g.seg[si];
    const int ya = sg.src_start + band * COST_RB, yb = min(ya + COST_RB, sg.src_end);
    const int ymin = sg.src_start, ymax = g.H - 1;
    const int x0 = blockIdx.x * TX;
    // tile geometry (uniform): cost-space columns [lo, hi), staged right-image chunks [cs, cs + nchunks)
    const int lo = max(x0 - g.SW2, 0), hi = min(x0 - g.SW2 + TXH, g.W1);
    const int j0_min = (g.W - 1 - (hi - 1 + g.minX1)) + g.minD, j0_max = (g.W - 1 - (lo + g.minX1)) + g.minD + 8 * (DQ - 1);
    const int cs = max(j0_min - 3, 0) >> 4, nchunks = ((j0_max + 7) >> 4) - cs + 1;
    const uint32_t pl_bytes = (uint32_t)(hi - lo) * 16, pr_bytes = (uint32_t)nchunks * PR_CHUNK;
    const size_t stage_bytes = (size_t)TXH * 16 + (size_t)((TXH + D + 32) / 16 + 1) * PR_CHUNK;   // as sized by the host
    uint8_t* stage = smem;                                                      // [STAGES][PL 64 x 16 B | PR chunks]
    uint8_t* ring = smem + COST_STAGES * stage_bytes;                           // [n][64][D] u8
    uint16_t* Vs = (uint16_t*)(ring + (size_t)n * TXH * D);                // [64][D] u16
    const int q = threadIdx.x % DQ, xg = threadIdx.x / DQ, XG = blockDim.x / DQ;
    int ix[ITEMS], o0[ITEMS], o1[ITEMS];
    uint32_t V[ITEMS][4];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const int xh = xg + k * XG;
        const int xc = min(max(x0 - g.SW2 + xh, 0), g.W1 - 1);
        ix[k] = min(max(xc - lo, 0), TXH - 1);
        const int j0 = (g.W - 1 - (xc + g.minX1)) + g.minD + 8 * q;
        const int a = j0 & 3, p0 = j0 - a, p1 = p0 + 4;
        o0[k] = ((p0 >> 4) - cs) * PR_CHUNK + (p0 & 15) + a * 16;
        o1[k] = ((p1 >> 4) - cs) * PR_CHUNK + (p1 & 15) + a * 16;
        V[k][0] = V[k][1] = V[k][2] = V[k][3] = 0;
        if (xh < TXH)
            for (int s = 0; s < n; s++) *(uint2*)(ring + ((size_t)s * TXH + xh) * D + 8 * q) = make_uint2(0, 0);
    }
    const int kv0 = ya - g.SH2, NV = (yb - 1 + g.SH2) - kv0 + 1;
    auto issue = [&](int i) {   // one thread: stage virtual row i
        const int rc = min(max(kv0 + i, ymin), ymax);
        uint8_t* st = stage + (size_t)(i % COST_STAGES) * stage_bytes;
        uint64_t* bar = &full[i % COST_STAGES];
        mbar_expect_tx(bar, pl_bytes + pr_bytes);
        bulk_g2s(st, PL + ((size_t)b * g.H + rc) * g.W + lo + g.minX1, pl_bytes, bar);
        bulk_g2s(st + TXH * 16, PR + (((size_t)b * g.H + rc) * (g.Wp >> 4) + cs) * PR_CHUNK, pr_bytes, bar);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < COST_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 0; i < COST_STAGES && i < NV; i++) issue(i);
    int slot = 0;
    for (int i = 0; i < NV; i++) {
        const int kv = kv0 + i;
        const uint8_t* st = stage + (size_t)(i % COST_STAGES) * stage_bytes;
        mbar_wait(&full[i % COST_STAGES], (uint32_t)((i / COST_STAGES) & 1));
        const int4* pls = (const int4*)st;
        const uint8_t* prs = st + TXH * 16;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const int xh = xg + k * XG;
            if (xh < TXH) {
                uint32_t n0, n1;
                bt_item(pls[ix[k]], prs, o0[k], o1[k], n0, n1);
                uint2* rp = (uint2*)(ring + ((size_t)slot * TXH + xh) * D + 8 * q);
                const uint2 old = *rp;
                *rp = make_uint2(n0, n1);
                const uint32_t c0 = ~old.x, c1 = ~old.y;   // 255 - old per byte
                // V += new - old  ==  V + new + (255 - old) - 255 per 16-bit half
                V[k][0] = __vadd2(__vadd2(__vadd2(V[k][0], prmt(n0, 0, 0x4140)), prmt(c0, 0, 0x4140)), 0xFF01FF01u);
                V[k][1] = __vadd2(__vadd2(__vadd2(V[k][1], prmt(n0, 0, 0x4342)), prmt(c0, 0, 0x4342)), 0xFF01FF01u);
                V[k][2] = __vadd2(__vadd2(__vadd2(V[k][2], prmt(n1, 0, 0x4140)), prmt(c1, 0, 0x4140)), 0xFF01FF01u);
                V[k][3] = __vadd2(__vadd2(__vadd2(V[k][3], prmt(n1, 0, 0x4342)), prmt(c1, 0, 0x4342)), 0xFF01FF01u);
            }
        }
        slot = slot + 1 == n ? 0 : slot + 1;
        const int y = kv - g.SH2;
        const bool out_row = y >= ya;   // uniform across the CTA
        if (out_row) {
#pragma unroll
            for (int k = 0; k < ITEMS; k++) {
                const int xh = xg + k * XG;
                if (xh < TXH) *(uint4*)(Vs + (size_t)xh * D + 8 * q) = make_uint4(V[k][0], V[k][1], V[k][2], V[k][3]);
            }
        }
        __syncthreads();   // every thread is done with this stage (and Vs is complete)
        if (threadIdx.x == 0 && i + COST_STAGES < NV) issue(i + COST_STAGES);
        if (out_row) {
            int16_t* Crow = C + ((size_t)b * g.crows + sg.crow0 + (y - sg.src_start)) * g.W1 * D;
#pragma unroll
            for (int k = 0; k < ITEMS; k++) {
                const int xl = xg + k * XG;
                if (xl < TX && x0 + xl < g.W1) {
                    uint4 acc = make_uint4(0, 0, 0, 0);
                    for (int j = 0; j <= 2 * g.SW2; j++) {
                        const uint4 w = *(const uint4*)(Vs + (size_t)(xl + j) * D + 8 * q);
                        acc.x = __vadd2(acc.x, w.x);
                        acc.y = __vadd2(acc.y, w.y);
                        acc.z = __vadd2(acc.z, w.z);
                        acc.w = __vadd2(acc.w, w.w);
                    }
                    *(uint4*)(Crow + (size_t)(x0 + xl) * D + 8 * q) = acc;
                }
            }
            __syncthreads();   // Vs is single-buffered
        }
    }
}

// ------------------------------------------------------------------------------------------
// Path aggregation. A chain's D costs live in G lanes x NP packed registers (D = 2*G*NP).
// State P is the previous L normalised by its minimum, so
//   L[d] = C[d] + min(P[d], min(P[d-1], P[d+1]) + P1, P2)        (== OpenCV's Lr)
// ------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ void load_regs(const int16_t* p, uint32_t (&r)[NP]) {
    if constexpr (NP % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NP / 4; i++) {
            uint4 v = *((const uint4*)p + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
    } else if constexpr (NP % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NP / 2; i++) {
            uint2 v = *((const uint2*)p + i);
            r[2 * i] = v.x; r[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NP; i++) r[i] = *((const uint32_t*)p + i);
    }
}
template <int NP>
__device__ __forceinline__ void store_regs(int16_t* p, const uint32_t (&r)[NP]) {
    if constexpr (NP % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NP / 4; i++) *((uint4*)p + i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    } else if constexpr (NP % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NP / 2; i++) *((uint2*)p + i) = make_uint2(r[2 * i], r[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < NP; i++) *((uint32_t*)p + i) = r[i];
    }
}

// Shared-memory copy of one column's D path costs held by G = 8 lanes x NP registers, laid out so that every 128-bit
// access of a quarter-warp covers 128 contiguous bytes: register quad q of lane lg sits at int16 offset (q*8 + lg)*8.
// (The natural layout -- lane lg's 2*NP values contiguous -- makes lanes lg and lg+4 hit the same banks when NP = 8.)
template <int NP, int G = 8>
__device__ __forceinline__ void load_col_sm(const int16_t* col, int lg, uint32_t (&r)[NP]) {
    if constexpr (NP % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NP / 4; q++) {
            const uint4 v = *(const uint4*)(col + (q * G + lg) * 8);
            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
    } else {
        load_regs<NP>(col + lg * 2 * NP, r);
    }
}
template <int NP, int G = 8>
__device__ __forceinline__ void store_col_sm(int16_t* col, int lg, const uint32_t (&r)[NP]) {
    if constexpr (NP % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NP / 4; q++) *(uint4*)(col + (q * G + lg) * 8) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    } else {
        store_regs<NP>(col + lg * 2 * NP, r);
    }
}

// One path step, in place: on entry P = L of the previous pixel on the path and m = its minimum (packed in both halves);
// on exit P = L of this pixel and m = its minimum.
//   L[d] = C[d] + min( min(P[d], min(P[d-1], P[d+1]) + P1) - m, P2 )
// which is OpenCV's C + min(Lr[d], Lr[d-1]+P1, Lr[d+1]+P1, minLr+P2) - (minLr+P2) with the +P2 bias folded out of C.
template <int NP, int G>
__device__ __forceinline__ void sgm_step(uint32_t (&P)[NP], uint32_t& m, const uint32_t (&c)[NP], uint32_t P1x2, uint32_t P2x2,
                                         uint32_t PADx2, int lg) {
    uint32_t up = __shfl_up_sync(0xffffffffu, P[NP - 1], 1, G);
    uint32_t dn = __shfl_down_sync(0xffffffffu, P[0], 1, G);
    if (lg == 0) up = PADx2;
    if (lg == G - 1) dn = PADx2;
    uint32_t T[NP + 1];
    T[0] = prmt(up, P[0], 0x5432);
#pragma unroll
    for (int k = 1; k < NP; k++) T[k] = prmt(P[k - 1], P[k], 0x5432);
    T[NP] = prmt(P[NP - 1], dn, 0x5432);
    const uint32_t negm = __vadd2(~m, 0x00010001u);
    uint32_t mm = 0x7FFF7FFFu;
#pragma unroll
    for (int k = 0; k < NP; k++) {
        uint32_t t = __vmins2(T[k], T[k + 1]);
        t = __viaddmin_s16x2(t, P1x2, P[k]);      // min(t + P1, P[d])
        t = __viaddmin_s16x2(t, negm, P2x2);      // min(. - m, P2)
        P[k] = __vadd2(c[k], t);
        if (k & 1) mm = __vimin3_s16x2(mm, P[k - 1], P[k]);
    }
    if (NP & 1) mm = __vmins2(mm, P[NP - 1]);
    mm = __vmins2(mm, prmt(mm, mm, 0x1032));
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) mm = __vmins2(mm, __shfl_xor_sync(0xffffffffu, mm, o, G));
    m = mm;
}
// N independent path steps (the sweep's two diagonals) advanced together: the same arithmetic as N sgm_step calls, written
// so that the shuffles of the chains are issued back to back and their arithmetic interleaves. Pays only when few warps
// share a scheduler (narrow strips, one frame at a time: fused kernel 1.018 -> 0.984 ms); with 64-column strips at batch 8
// it is slower (1.99 -> 2.06 ms for two chains, 2.17 for three): profiles/experiments/README.md.
template <int NP, int G, int N>
__device__ __forceinline__ void sgm_step_n(uint32_t* const (&P)[N], uint32_t (&m)[N], const uint32_t (&c)[NP], uint32_t P1x2, uint32_t P2x2,
                                           uint32_t PADx2, int lg) {
    uint32_t up[N], dn[N], negm[N], mm[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        up[j] = __shfl_up_sync(0xffffffffu, P[j][NP - 1], 1, G);
        dn[j] = __shfl_down_sync(0xffffffffu, P[j][0], 1, G);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (lg == 0) up[j] = PADx2;
        if (lg == G - 1) dn[j] = PADx2;
        negm[j] = __vadd2(~m[j], 0x00010001u);
        mm[j] = 0x7FFF7FFFu;
    }
    uint32_t lo[N];   // prmt(P[k-1], P[k]) of the register being updated (old values)
#pragma unroll
    for (int j = 0; j < N; j++) lo[j] = prmt(up[j], P[j][0], 0x5432);
#pragma unroll
    for (int k = 0; k < NP; k++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
            const uint32_t hi = prmt(P[j][k], k + 1 < NP ? P[j][k + 1] : dn[j], 0x5432);
            uint32_t t = __vmins2(lo[j], hi);
            t = __viaddmin_s16x2(t, P1x2, P[j][k]);
            t = __viaddmin_s16x2(t, negm[j], P2x2);
            P[j][k] = __vadd2(c[k], t);
            lo[j] = hi;
            if (k & 1) mm[j] = __vimin3_s16x2(mm[j], P[j][k - 1], P[j][k]);
        }
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (NP & 1) mm[j] = __vmins2(mm[j], P[j][NP - 1]);
        mm[j] = __vmins2(mm[j], prmt(mm[j], mm[j], 0x1032));
    }
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) {
        uint32_t t[N];
#pragma unroll
        for (int j = 0; j < N; j++) t[j] = __shfl_xor_sync(0xffffffffu, mm[j], o, G);
#pragma unroll
        for (int j = 0; j < N; j++) mm[j] = __vmins2(mm[j], t[j]);
    }
#pragma unroll
    for (int j = 0; j < N; j++) m[j] = mm[j];
}
template <int NP, int G>
__global__ void __launch_bounds__(128) k_path_v(Geom g, const int16_t* __restrict__ C, int16_t* __restrict__ S, int dirx,
                                                int diry, int accumulate) {
    constexpr int CPB = 128 / G;
    const int lane = threadIdx.x & 31, lg = lane & (G - 1);
    const int chain = blockIdx.x * CPB + threadIdx.x / G;
    const bool active = chain < g.W1;
    const Seg sg = g.seg[blockIdx.y];
    const int b = blockIdx.z;
    const int D = g.D;
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    int x = min(chain, g.W1 - 1);
    const int nrows = sg.src_end - sg.src_start;
    int y = diry > 0 ? sg.src_start : sg.src_end - 1;
    uint32_t P[NP], c[NP], s[NP], L[NP], cn[NP], sn[NP], m = 0;
#pragma unroll
    for (int k = 0; k < NP; k++) { P[k] = 0; sn[k] = 0; }
    const size_t lane_off = (size_t)lg * 2 * NP;
    const int16_t* Cb = C + ((size_t)b * g.crows + sg.crow0) * g.W1 * D + lane_off;
    int16_t* Sb = S + (size_t)b * g.H * g.W1 * D + lane_off;
    load_regs<NP>(Cb + ((size_t)(y - sg.src_start) * g.W1 + x) * D, cn);
    if (accumulate && y >= sg.dst_start) load_regs<NP>(Sb + ((size_t)y * g.W1 + x) * D, sn);
    for (int step = 0; step < nrows; step++) {
#pragma unroll
        for (int k = 0; k < NP; k++) { c[k] = cn[k]; s[k] = sn[k]; }
        // next position
        int xn = x + dirx, yn = y + diry;
        bool wrap = false;
        if (xn >= g.W1) { xn = 0; wrap = true; }
        if (xn < 0) { xn = g.W1 - 1; wrap = true; }
        if (step + 1 < nrows) {
            load_regs<NP>(Cb + ((size_t)(yn - sg.src_start) * g.W1 + xn) * D, cn);
            if (accumulate && yn >= sg.dst_start) load_regs<NP>(Sb + ((size_t)yn * g.W1 + xn) * D, sn);
        }
        sgm_step<NP, G>(P, m, c, P1x2, P2x2, PADx2, lg);
        if (active && y >= sg.dst_start) {
#pragma unroll
            for (int k = 0; k < NP; k++) L[k] = accumulate ? __vminu2(__vadd2(s[k], P[k]), 0x7FFF7FFFu) : P[k];
            store_regs<NP>(Sb + ((size_t)y * g.W1 + x) * D, L);
        }
        if (wrap) {
#pragma unroll
            for (int k = 0; k < NP; k++) P[k] = 0;
            m = 0;
        }
        x = xn;
        y = yn;
    }
}

// ---- fused top-down (or bottom-up) sweep: vertical + both diagonal paths in ONE pass over C -------------------
// A CTA owns a strip of 32 columns of one frame and walks the rows; a column's D costs live in 8 lanes. The vertical
// state stays in registers, the diagonal states move one column per row through shared memory, and across strip
// borders through a small global halo guarded by per-strip progress flags (neighbours only ever wait for the
// previous row, so every CTA of a launch must be co-resident: the kernel is launched cooperatively).
// C is read once and S written once ("directions may run concurrently and add into S in any order").
template <int NP>
__device__ __forceinline__ void load_regs_cg(const int16_t* p, uint32_t (&r)[NP]) {
#pragma unroll
    for (int i = 0; i < NP; i++) r[i] = __ldcg((const unsigned int*)p + i);
}

constexpr int V3_COLS = 32;
constexpr int V3_STAGES_MAX = 6;       // rows of C in flight per CTA (TMA ring); 3 when S is staged too (keeps 2 CTAs/SM)
constexpr int V3_STAGES_DEFAULT = 6;
constexpr int V3_SPIN_LIMIT = 1 << 22;
#ifndef SDR_HALO_SLEEP_NS
#define SDR_HALO_SLEEP_NS 40   // back-off between two polls of a neighbour's halo words
#endif

// Halo words carry their own sequence number (row + 1) in the upper half of one 64-bit store, so no fence and no
// separate flag is needed: the reader polls the word until the sequence matches (single-copy-atomic 8-byte accesses).
__device__ __forceinline__ void halo_put(unsigned long long* p, uint32_t data, uint32_t seq) {
    const unsigned long long v = ((unsigned long long)seq << 32) | data;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Halo reads: all words of a lane are requested back to back (one L2 round trip), the sequence numbers are checked
// later (halo_ready), and only a stale word triggers the bounded re-poll.
template <int NW>
__device__ __forceinline__ void halo_issue(const unsigned long long* p, unsigned long long (&v)[NW]) {
#pragma unroll
    for (int k = 0; k < NW; k++) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v[k]) : "l"(p + k) : "memory");
}
template <int NW>
__device__ __forceinline__ void halo_complete(const unsigned long long* p, uint32_t seq, int* err, unsigned long long (&v)[NW]) {
    int spins = 0;
    while (true) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < NW; k++) ok = ok && (uint32_t)(v[k] >> 32) == seq;
        if (ok) break;
        ++spins;
        if ((spins & 63) == 0 && *(volatile int*)err) break;
        if (spins > V3_SPIN_LIMIT) { atomicExch(err, 1); break; }   // never hang the GPU on a broken protocol
        if (SDR_HALO_SLEEP_NS > 0) __nanosleep(SDR_HALO_SLEEP_NS);   // leave the issue slots to the warps this one is waiting for
        halo_issue<NW>(p, v);
    }
}

template <int NP, int MINB, bool ACC, int COLS, int G>
__global__ void __launch_bounds__(COLS * G, MINB) k_path_v3(Geom g, const int16_t* __restrict__ C, int16_t* __restrict__ S, int diry, int b0,
                                                       unsigned long long* halo, int* err, int V3_STAGES) {
    constexpr int NW = NP + 1;   // halo words per lane: NP state registers + the path minimum (a column = G lanes x NP registers)
    extern __shared__ __align__(16) int16_t v3sm[];
    const int D = g.D;
    int16_t* Xa = v3sm;                                     // [2][COLS+2][D]  state of the "from (x-1)" diagonal, slot = column+1
    int16_t* Xb = v3sm + 2 * (COLS + 2) * D;             // [2][COLS+2][D]  state of the "from (x+1)" diagonal
    uint32_t* Ma = (uint32_t*)(v3sm + 4 * (COLS + 2) * D);   // [2][COLS+2] path minima of Xa
    uint32_t* Mb = Ma + 2 * (COLS + 2);
    // ring of C (and, when accumulating, S) row tiles filled by TMA bulk copies V3_STAGES rows ahead: the strip's 32
    // columns of one row are one contiguous 32*D*2-byte run in global memory
    int16_t* ringC = (int16_t*)(Mb + 2 * (COLS + 2));        // [STAGES][COLS][D]
    int16_t* ringS = ringC + (size_t)V3_STAGES * COLS * D;   // [STAGES][COLS][D] (ACC only)
    __shared__ __align__(8) uint64_t full[V3_STAGES_MAX];
    const int s = blockIdx.x, nstrips = gridDim.x, f = blockIdx.y, b = b0 + f;
    const int gidx = threadIdx.x / G, lg = threadIdx.x & (G - 1);
    const int x = s * COLS + gidx;
    const bool active = x < g.W1;
    const int xl = min(x, g.W1 - 1);
    for (int i = threadIdx.x; i < (4 * (COLS + 2) * D) / 2 + 4 * (COLS + 2); i += blockDim.x) ((uint32_t*)v3sm)[i] = 0;
    __syncthreads();
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    const size_t lane_off = (size_t)lg * 2 * NP;
    // halo layout: [frame][strip][buf][a|b][G*NW] 64-bit words; (strip, a) = left halo OF that strip, (strip, b) = its right halo
    const int HW = G * NW;
    auto Hptr = [&](int strip, int buf, int which) {
        return halo + ((((size_t)f * nstrips + strip) * 2 + buf) * 2 + which) * HW + (size_t)lg * NW;
    };
    int16_t* Sb = S + (size_t)b * g.H * g.W1 * D + (size_t)xl * D + lane_off;
    const size_t rowstride = (size_t)g.W1 * D;
    const int x0 = s * COLS;
    const uint32_t tile_bytes = (uint32_t)(min(COLS, g.W1 - x0) * D * 2);
    const int16_t* Ctile = C + (size_t)b * g.crows * g.W1 * D + (size_t)x0 * D;    // + row * rowstride
    const int16_t* Stile = S + (size_t)b * g.H * g.W1 * D + (size_t)x0 * D;
    const int ystart = diry > 0 ? 0 : g.H - 1;
    auto issue = [&](int t, int st) {   // one thread: stage row t of the sweep into ring slot st (= t mod V3_STAGES)
        const int yy = ystart + t * diry;
        mbar_expect_tx(&full[st], ACC ? 2 * tile_bytes : tile_bytes);
        bulk_g2s(ringC + (size_t)st * COLS * D, Ctile + (size_t)yy * rowstride, tile_bytes, &full[st]);
        if (ACC) bulk_g2s(ringS + (size_t)st * COLS * D, Stile + (size_t)yy * rowstride, tile_bytes, &full[st]);
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < V3_STAGES; i++) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 0; i < V3_STAGES && i < g.H; i++) issue(i, i);
    const bool left_edge = gidx == 0 && s > 0, right_edge = gidx == COLS - 1 && s + 1 < nstrips;
    uint32_t Pv[NP], pa[NP], pb[NP], c[NP], so[ACC ? NP : 1], acc[NP];
    uint32_t mv = 0, ma = 0, mb = 0;
    unsigned long long hq[NW];   // prefetched halo of the next row (edge groups only)
#pragma unroll
    for (int k = 0; k < NP; k++) Pv[k] = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) hq[k] = 0;
    int y = ystart;
    int stg = 0;
    uint32_t ring_phase = 0;
    for (int t = 0; t < g.H; t++) {
        const int cur = t & 1, prv = cur ^ 1;
        mbar_wait(&full[stg], ring_phase);
        load_regs<NP>(ringC + ((size_t)stg * COLS + gidx) * D + lane_off, c);
        if constexpr (ACC) load_regs<NP>(ringS + ((size_t)stg * COLS + gidx) * D + lane_off, so);
        // previous-row diagonal states: neighbouring column through shared memory, across strips from the prefetched halo
        if (left_edge && t > 0) {
#pragma unroll
            for (int k = 0; k < NP; k++) pa[k] = (uint32_t)hq[k];
            ma = (uint32_t)hq[NP];
        } else {
            load_col_sm<NP, G>(Xa + ((size_t)prv * (COLS + 2) + gidx) * D, lg, pa);
            ma = Ma[prv * (COLS + 2) + gidx];
        }
        if (right_edge && t > 0) {
#pragma unroll
            for (int k = 0; k < NP; k++) pb[k] = (uint32_t)hq[k];
            mb = (uint32_t)hq[NP];
        } else {
            load_col_sm<NP, G>(Xb + ((size_t)prv * (COLS + 2) + gidx + 2) * D, lg, pb);
            mb = Mb[prv * (COLS + 2) + gidx + 2];
        }
        __syncwarp();
        sgm_step<NP, G>(pa, ma, c, P1x2, P2x2, PADx2, lg);
        if (!active) {
#pragma unroll
            for (int k = 0; k < NP; k++) pa[k] = 0;   // columns beyond W1 act as the zero border
            ma = 0;
        }
        if (right_edge) {   // publish first: the right neighbour needs this for its next row
            unsigned long long* hp = Hptr(s + 1, cur, 0);
#pragma unroll
            for (int k = 0; k < NP; k++) halo_put(hp + k, pa[k], (uint32_t)(t + 1));
            halo_put(hp + NP, ma, (uint32_t)(t + 1));
        }
        store_col_sm<NP, G>(Xa + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pa);
        if (lg == 0) Ma[cur * (COLS + 2) + gidx + 1] = ma;
        sgm_step<NP, G>(pb, mb, c, P1x2, P2x2, PADx2, lg);
        if (!active) {
#pragma unroll
            for (int k = 0; k < NP; k++) pb[k] = 0;
            mb = 0;
        }
        if (left_edge) {
            unsigned long long* hp = Hptr(s - 1, cur, 1);
#pragma unroll
            for (int k = 0; k < NP; k++) halo_put(hp + k, pb[k], (uint32_t)(t + 1));
            halo_put(hp + NP, mb, (uint32_t)(t + 1));
        }
        store_col_sm<NP, G>(Xb + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pb);
        if (lg == 0) Mb[cur * (COLS + 2) + gidx + 1] = mb;
        // request the neighbour's state of THIS row now (it publishes early in its row); it is consumed next row, so the
        // L2 round trip overlaps the vertical path below
        if (t + 1 < g.H) {
            if (left_edge) halo_issue<NW>(Hptr(s, cur, 0), hq);
            if (right_edge) halo_issue<NW>(Hptr(s, cur, 1), hq);
        }
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = __vminu2(__vadd2(pa[k], pb[k]), 0x7FFF7FFFu);
        __syncwarp();
        sgm_step<NP, G>(Pv, mv, c, P1x2, P2x2, PADx2, lg);
#pragma unroll
        for (int k = 0; k < NP; k++) {
            acc[k] = __vminu2(__vadd2(acc[k], Pv[k]), 0x7FFF7FFFu);
            if constexpr (ACC) acc[k] = __vminu2(__vadd2(acc[k], so[k]), 0x7FFF7FFFu);
        }
        if (active) store_regs<NP>(Sb + (size_t)y * rowstride, acc);
        if (t + 1 < g.H) {
            if (left_edge) halo_complete<NW>(Hptr(s, cur, 0), (uint32_t)(t + 1), err, hq);
            if (right_edge) halo_complete<NW>(Hptr(s, cur, 1), (uint32_t)(t + 1), err, hq);
        }
        __syncthreads();   // also: every thread is done with ring stage `stg`
        if (threadIdx.x == 0 && t + V3_STAGES < g.H) issue(t + V3_STAGES, stg);
        if (++stg == V3_STAGES) { stg = 0; ring_phase ^= 1; }
        y += diry;
    }
}

// ---- fused cost + top-down sweep ---------------------------------------------------------------------------------
// One CTA = one 32-column strip of one frame, 12 warps in two roles:
//   first warps (producer, CF_PROD threads): BT pixel cost -> vertical running box sum -> horizontal box sum for the strip's
//       32 columns (+2*SW2 halo columns), one image row after the other, into a ring of C row tiles in shared memory;
//       every finished tile is also written to the C volume in HBM (one TMA bulk store) for the horizontal kernels;
//   last 8 warps (consumer, 256 threads): the three top-down paths of k_path_v3 reading C from that ring.
// The sweep alone leaves ~60 % of the issue slots idle (it waits for its neighbours' halos every row) and the cost
// kernel alone is issue-bound: in one kernel the producer fills those slots, and the sweep's read of C from HBM
// (one volume pass) disappears. Registers are split with setmaxnreg (producer 64, consumer 88 per thread).
constexpr int CF_PROD_DEFAULT = 128;
#ifndef SDR_CF_NST
#define SDR_CF_NST 3
#endif
constexpr int CF_NST = SDR_CF_NST;          // C row tiles in flight between producer and consumer

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

struct CfLayout {   // shared-memory carve-up (bytes), computed identically on host and device
    int cw, nchunks_max, stage_bytes;
    size_t off_ma, off_ring_c, off_stage, off_ring8, off_vs, total;
};
__host__ __device__ inline CfLayout cf_layout(int D, int SW2, int SH2, int COLS) {
    CfLayout L;
    L.cw = COLS + 2 * SW2;
    L.nchunks_max = (L.cw + D + 32) / 16 + 1;
    L.stage_bytes = ((L.cw * 16 + L.nchunks_max * PR_CHUNK) + 127) & ~127;
    size_t o = (size_t)4 * (COLS + 2) * D * 2;          // Xa, Xb
    L.off_ma = o;  o += (size_t)4 * (COLS + 2) * 4;     // Ma, Mb
    o = (o + 127) & ~(size_t)127;
    L.off_ring_c = o;  o += (size_t)CF_NST * COLS * D * 2;
    L.off_stage = o;   o += (size_t)2 * L.stage_bytes;
    L.off_ring8 = o;   o += (size_t)(2 * SH2 + 1) * L.cw * D;
    o = (o + 15) & ~(size_t)15;
    L.off_vs = o;      o += (size_t)L.cw * D * 2;
    L.total = o;
    return L;
}

// The producer role of k_wave_fwd (the same code as the producer branch of k_cost_path_v3, which keeps its own inlined copy:
// routed through this function ptxas spills 20 bytes there and the kernel loses 5 %): CF_PROD threads compute the C row tiles of one COLS-column
// strip of frame b, one image row after the other, into the ring `ringC` (CF_NST slots, hand-shake fullC / emptyC with the
// readers) and write every finished tile to the C volume with one TMA bulk store.
// The rows a producer walks: output rows [y0, y0 + nrows) of the image, input rows clamped to [ymin, ymax] (a stripe of
// MODE_SGBM_3WAY replicates its own first row above it but reads on below its last row, as k_cost does), C rows from crow0.
struct CfRows {
    int y0, nrows, ymin, ymax, crow0;
};
template <int NP, int SW2, int CF_PROD, int COLS>
__device__ __forceinline__ void cost_producer(const Geom& g, const int4* __restrict__ PL, const uint8_t* __restrict__ PR, int16_t* __restrict__ C,
                                              int b, int x0, uint8_t* cfsm, const CfLayout& lay, uint64_t (&fullC)[CF_NST], uint64_t (&emptyC)[CF_NST],
                                              uint64_t (&fullIn)[2], uint32_t tile_bytes, const CfRows rw) {
    int16_t* ringC = (int16_t*)(cfsm + lay.off_ring_c);
    constexpr int D = 16 * NP, DQ = D / 8;
    constexpr int XG = CF_PROD / DQ;            // producer column groups
    constexpr int CPT = COLS / XG;              // consecutive output columns per producer thread
    constexpr int NX = 2 * SW2;                 // halo columns of the strip (box-sum window minus one)
    constexpr int CW = COLS + NX;
    static_assert(CF_PROD % DQ == 0 && COLS % XG == 0 && NX <= XG, "producer tiling");
    const int n = 2 * g.SH2 + 1;
    const int tid = threadIdx.x;
    const int q = tid % DQ, xg = tid / DQ;
    uint8_t* stage = cfsm + lay.off_stage;
    uint8_t* ring8 = cfsm + lay.off_ring8;                 // [n][CW][D] u8
    uint16_t* Vs = (uint16_t*)(cfsm + lay.off_vs);         // [CW][D] u16
    const int lo = max(x0 - SW2, 0), hi = min(x0 - SW2 + CW, g.W1);
    const int j0_min = (g.W - 1 - (hi - 1 + g.minX1)) + g.minD, j0_max = (g.W - 1 - (lo + g.minX1)) + g.minD + 8 * (DQ - 1);
    const int cs = max(j0_min - 3, 0) >> 4, nchunks = ((j0_max + 7) >> 4) - cs + 1;
    const uint32_t pl_bytes = (uint32_t)(hi - lo) * 16, pr_bytes = (uint32_t)nchunks * PR_CHUNK;
    constexpr int ITEMS = CPT + 1;
    // item k < CPT: strip column h = CPT*xg + k (h counts from x0 - SW2); item CPT: halo column h = 32 + xg (xg < NX)
    uint32_t tab[ITEMS];   // per item: o0 (low 16 bits) | ix << 16 | "second word is in the next chunk" << 31
    uint32_t V[ITEMS][4];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const int h = k < CPT ? CPT * xg + k : COLS + xg;
        const int xc = min(max(x0 - SW2 + h, 0), g.W1 - 1);
        const int ix = min(max(xc - lo, 0), CW - 1);
        const int j0 = (g.W - 1 - (xc + g.minX1)) + g.minD + 8 * q;
        const int a = j0 & 3, p0 = j0 - a;
        const int o0 = ((p0 >> 4) - cs) * PR_CHUNK + (p0 & 15) + a * 16;
        tab[k] = (uint32_t)o0 | ((uint32_t)ix << 16) | ((p0 & 15) == 12 ? 0x80000000u : 0u);
        V[k][0] = V[k][1] = V[k][2] = V[k][3] = 0;
    }
    const int NV = rw.nrows + 2 * g.SH2;
    auto issue_in = [&](int i) {   // one thread: stage virtual row i (image row clamp(y0 + i - SH2))
        const int rc = min(max(rw.y0 + i - g.SH2, rw.ymin), rw.ymax);
        uint8_t* st = stage + (size_t)(i & 1) * lay.stage_bytes;
        mbar_expect_tx(&fullIn[i & 1], pl_bytes + pr_bytes);
        bulk_g2s(st, PL + ((size_t)b * g.H + rc) * g.W + lo + g.minX1, pl_bytes, &fullIn[i & 1]);
        bulk_g2s(st + CW * 16, PR + (((size_t)b * g.H + rc) * (g.Wp >> 4) + cs) * PR_CHUNK, pr_bytes, &fullIn[i & 1]);
    };
    if (tid == 0) {
        issue_in(0);
        if (NV > 1) issue_in(1);
    }
    int slot = 0, cst = 0;
    uint32_t cphase = 0;
    for (int i = 0; i < NV; i++) {
        const uint8_t* st = stage + (size_t)(i & 1) * lay.stage_bytes;
        mbar_wait(&fullIn[i & 1], (uint32_t)((i >> 1) & 1));
        const int4* pls = (const int4*)st;
        const uint8_t* prs = st + CW * 16;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            if (k < CPT || xg < NX) {
                const int h = k < CPT ? CPT * xg + k : COLS + xg;
                const int o0 = (int)(tab[k] & 0xFFFFu), ix = (int)((tab[k] >> 16) & 0x7FFFu);
                const int o1 = o0 + ((int)tab[k] < 0 ? PR_CHUNK - 12 : 4);
                uint32_t n0, n1;
                bt_item(pls[ix], prs, o0, o1, n0, n1);
                uint2* rp = (uint2*)(ring8 + ((size_t)slot * CW + h) * D + 8 * q);
                const uint2 old = *rp;
                *rp = make_uint2(n0, n1);
                // V += new - old per 16-bit half: new (bytes, zero-extended) minus old (bytes, zero-extended)
                V[k][0] = __vsub2(__vadd2(V[k][0], prmt(n0, 0, 0x4140)), prmt(old.x, 0, 0x4140));
                V[k][1] = __vsub2(__vadd2(V[k][1], prmt(n0, 0, 0x4342)), prmt(old.x, 0, 0x4342));
                V[k][2] = __vsub2(__vadd2(V[k][2], prmt(n1, 0, 0x4140)), prmt(old.y, 0, 0x4140));
                V[k][3] = __vsub2(__vadd2(V[k][3], prmt(n1, 0, 0x4342)), prmt(old.y, 0, 0x4342));
            }
        }
        slot = slot + 1 == n ? 0 : slot + 1;
        const int y = i - 2 * g.SH2;          // output row finished by this virtual row
        if (y >= 0) {
#pragma unroll
            for (int k = 0; k < ITEMS; k++)
                if (k < CPT || xg < NX) {
                    const int h = k < CPT ? CPT * xg + k : COLS + xg;
                    *(uint4*)(Vs + (size_t)h * D + 8 * q) = make_uint4(V[k][0], V[k][1], V[k][2], V[k][3]);
                }
            // the bulk store that last read ring slot cst must be done reading before the slot is rewritten
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(CF_NST - 1) : "memory");
        }
        named_bar_sync(2, CF_PROD);   // Vs complete; every producer thread is done with this input stage
        if (tid == 0 && i + 2 < NV) issue_in(i + 2);
        if (y >= 0) {
            if (y >= CF_NST) mbar_wait(&emptyC[cst], cphase ^ 1);   // consumers released the slot (use number y/NST - 1)
            // horizontal box sum: window of CPT + NX columns = own columns (registers) + NX right neighbours (Vs)
            uint32_t w[CPT + NX][4];
#pragma unroll
            for (int k = 0; k < CPT; k++) { w[k][0] = V[k][0]; w[k][1] = V[k][1]; w[k][2] = V[k][2]; w[k][3] = V[k][3]; }
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const uint4 v = *(const uint4*)(Vs + (size_t)(CPT * xg + CPT + j) * D + 8 * q);
                w[CPT + j][0] = v.x; w[CPT + j][1] = v.y; w[CPT + j][2] = v.z; w[CPT + j][3] = v.w;
            }
            uint32_t acc[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                acc[r] = w[0][r];
#pragma unroll
                for (int j = 1; j <= NX; j++) acc[r] = __vadd2(acc[r], w[j][r]);
            }
            int16_t* ct = ringC + (size_t)cst * COLS * D;
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                if (k > 0) {
#pragma unroll
                    for (int r = 0; r < 4; r++) acc[r] = __vsub2(__vadd2(acc[r], w[k + NX][r]), w[k - 1][r]);
                }
                *(uint4*)(ct + (size_t)(CPT * xg + k) * D + 8 * q) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
            named_bar_sync(3, CF_PROD);   // tile complete (and every read of Vs is done)
            if (tid == 0) {
                bulk_s2g(C + (((size_t)b * g.crows + rw.crow0 + y) * g.W1 + x0) * D, ct, tile_bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                mbar_arrive(&fullC[cst]);
            }
            if (++cst == CF_NST) { cst = 0; cphase ^= 1; }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// register split between the roles (setmaxnreg), per strip width: producer / consumer registers per thread
template <int CF_PROD, int COLS> struct CfRegs {
    static constexpr int prod = COLS <= 16 ? 56 : (2 * CF_PROD == COLS * 8 ? 64 : 48);
    static constexpr int cons = COLS <= 16 ? 96 : (2 * CF_PROD == COLS * 8 ? 88 : 80);
    static constexpr int min_blocks = COLS == 8 ? 4 : COLS == 16 ? 3 : COLS == 32 ? 2 : 1;   // fixes the launch register count: 80 / 80 or 64 / 80
};

template <int NP, int SW2, int CF_PROD, int COLS, int G>   // a column's D = 2 * G * NP costs: G lanes x NP packed registers
__global__ void __launch_bounds__(CF_PROD + COLS * G, CfRegs<CF_PROD, COLS * G / 8>::min_blocks)
k_cost_path_v3(Geom g, const int4* __restrict__ PL, const uint8_t* __restrict__ PR, int16_t* __restrict__ C, int16_t* __restrict__ S, int b0,
               unsigned long long* halo, int* err) {
    constexpr int NW = NP + 1, D = 2 * G * NP, DQ = D / 8;
    constexpr int CF_CONS = COLS * G;           // consumer threads (8 lanes per column)
    constexpr int XG = CF_PROD / DQ;            // producer column groups
    constexpr int CPT = COLS / XG;           // consecutive output columns per producer thread
    constexpr int NX = 2 * SW2;                 // halo columns of the strip (box-sum window minus one)
    constexpr int CW = COLS + NX;
    static_assert(CF_PROD % DQ == 0 && COLS % XG == 0 && NX <= XG, "producer tiling");
    extern __shared__ __align__(128) uint8_t cfsm[];
    __shared__ __align__(8) uint64_t fullC[CF_NST], emptyC[CF_NST], fullIn[2];
    const CfLayout lay = cf_layout(D, SW2, g.SH2, COLS);
    int16_t* Xa = (int16_t*)cfsm;
    int16_t* Xb = Xa + 2 * (COLS + 2) * D;
    uint32_t* Ma = (uint32_t*)(cfsm + lay.off_ma);
    uint32_t* Mb = Ma + 2 * (COLS + 2);
    int16_t* ringC = (int16_t*)(cfsm + lay.off_ring_c);
    const int s = blockIdx.x, nstrips = gridDim.x, f = blockIdx.y, b = b0 + f;
    const int x0 = s * COLS;
    const int n = 2 * g.SH2 + 1;
    // ---- prologue (all threads): zero the diagonal states and the pixel-cost ring, set up the barriers
    for (int i = threadIdx.x; i < (int)(lay.off_ring_c / 4); i += blockDim.x) ((uint32_t*)cfsm)[i] = 0;
    for (int i = threadIdx.x; i < (int)((lay.off_vs - lay.off_ring8) / 4); i += blockDim.x) ((uint32_t*)(cfsm + lay.off_ring8))[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < CF_NST; i++) { mbar_init(&fullC[i], 1); mbar_init(&emptyC[i], CF_CONS / 32); }
        mbar_init(&fullIn[0], 1);
        mbar_init(&fullIn[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tile_bytes = (uint32_t)(min(COLS, g.W1 - x0) * D * 2);

    // Producer warps come first: the SM's warp arbiter favours the higher warp slots, which gives the latency-critical
    // consumer chain priority while the producer's independent arithmetic fills the remaining issue slots.
    if (threadIdx.x < CF_PROD) {
        // =============================== producer: cost rows ===============================
        // register split (per thread): 128 producers 64 / consumers 88 (launch 80); 256 producers 48 / consumers 80 (launch 64);
        // 16-column strips: 128 producers 56 / 128 consumers 96 (launch 80)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(CfRegs<CF_PROD, COLS * G / 8>::prod));
        const int tid = threadIdx.x;
        const int q = tid % DQ, xg = tid / DQ;
        uint8_t* stage = cfsm + lay.off_stage;
        uint8_t* ring8 = cfsm + lay.off_ring8;                 // [n][CW][D] u8
        uint16_t* Vs = (uint16_t*)(cfsm + lay.off_vs);         // [CW][D] u16
        const int lo = max(x0 - SW2, 0), hi = min(x0 - SW2 + CW, g.W1);
        const int j0_min = (g.W - 1 - (hi - 1 + g.minX1)) + g.minD, j0_max = (g.W - 1 - (lo + g.minX1)) + g.minD + 8 * (DQ - 1);
        const int cs = max(j0_min - 3, 0) >> 4, nchunks = ((j0_max + 7) >> 4) - cs + 1;
        const uint32_t pl_bytes = (uint32_t)(hi - lo) * 16, pr_bytes = (uint32_t)nchunks * PR_CHUNK;
        constexpr int ITEMS = CPT + 1;
        // item k < CPT: strip column h = CPT*xg + k (h counts from x0 - SW2); item CPT: halo column h = 32 + xg (xg < NX)
        uint32_t tab[ITEMS];   // per item: o0 (low 16 bits) | ix << 16 | "second word is in the next chunk" << 31
        uint32_t V[ITEMS][4];
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const int h = k < CPT ? CPT * xg + k : COLS + xg;
            const int xc = min(max(x0 - SW2 + h, 0), g.W1 - 1);
            const int ix = min(max(xc - lo, 0), CW - 1);
            const int j0 = (g.W - 1 - (xc + g.minX1)) + g.minD + 8 * q;
            const int a = j0 & 3, p0 = j0 - a;
            const int o0 = ((p0 >> 4) - cs) * PR_CHUNK + (p0 & 15) + a * 16;
            tab[k] = (uint32_t)o0 | ((uint32_t)ix << 16) | ((p0 & 15) == 12 ? 0x80000000u : 0u);
            V[k][0] = V[k][1] = V[k][2] = V[k][3] = 0;
        }
        const int NV = g.H + 2 * g.SH2;
        auto issue_in = [&](int i) {   // one thread: stage virtual row i (image row clamp(i - SH2))
            const int rc = min(max(i - g.SH2, 0), g.H - 1);
            uint8_t* st = stage + (size_t)(i & 1) * lay.stage_bytes;
            mbar_expect_tx(&fullIn[i & 1], pl_bytes + pr_bytes);
            bulk_g2s(st, PL + ((size_t)b * g.H + rc) * g.W + lo + g.minX1, pl_bytes, &fullIn[i & 1]);
            bulk_g2s(st + CW * 16, PR + (((size_t)b * g.H + rc) * (g.Wp >> 4) + cs) * PR_CHUNK, pr_bytes, &fullIn[i & 1]);
        };
        if (tid == 0) {
            issue_in(0);
            if (NV > 1) issue_in(1);
        }
        int slot = 0, cst = 0;
        uint32_t cphase = 0;
        for (int i = 0; i < NV; i++) {
            const uint8_t* st = stage + (size_t)(i & 1) * lay.stage_bytes;
            mbar_wait(&fullIn[i & 1], (uint32_t)((i >> 1) & 1));
            const int4* pls = (const int4*)st;
            const uint8_t* prs = st + CW * 16;
#pragma unroll
            for (int k = 0; k < ITEMS; k++) {
                if (k < CPT || xg < NX) {
                    const int h = k < CPT ? CPT * xg + k : COLS + xg;
                    const int o0 = (int)(tab[k] & 0xFFFFu), ix = (int)((tab[k] >> 16) & 0x7FFFu);
                    const int o1 = o0 + ((int)tab[k] < 0 ? PR_CHUNK - 12 : 4);
                    uint32_t n0, n1;
                    bt_item(pls[ix], prs, o0, o1, n0, n1);
                    uint2* rp = (uint2*)(ring8 + ((size_t)slot * CW + h) * D + 8 * q);
                    const uint2 old = *rp;
                    *rp = make_uint2(n0, n1);
                    // V += new - old per 16-bit half: new (bytes, zero-extended) minus old (bytes, zero-extended)
                    V[k][0] = __vsub2(__vadd2(V[k][0], prmt(n0, 0, 0x4140)), prmt(old.x, 0, 0x4140));
                    V[k][1] = __vsub2(__vadd2(V[k][1], prmt(n0, 0, 0x4342)), prmt(old.x, 0, 0x4342));
                    V[k][2] = __vsub2(__vadd2(V[k][2], prmt(n1, 0, 0x4140)), prmt(old.y, 0, 0x4140));
                    V[k][3] = __vsub2(__vadd2(V[k][3], prmt(n1, 0, 0x4342)), prmt(old.y, 0, 0x4342));
                }
            }
            slot = slot + 1 == n ? 0 : slot + 1;
            const int y = i - 2 * g.SH2;          // output row finished by this virtual row
            if (y >= 0) {
#pragma unroll
                for (int k = 0; k < ITEMS; k++)
                    if (k < CPT || xg < NX) {
                        const int h = k < CPT ? CPT * xg + k : COLS + xg;
                        *(uint4*)(Vs + (size_t)h * D + 8 * q) = make_uint4(V[k][0], V[k][1], V[k][2], V[k][3]);
                    }
                // the bulk store that last read ring slot cst must be done reading before the slot is rewritten
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(CF_NST - 1) : "memory");
            }
            named_bar_sync(2, CF_PROD);   // Vs complete; every producer thread is done with this input stage
            if (tid == 0 && i + 2 < NV) issue_in(i + 2);
            if (y >= 0) {
                if (y >= CF_NST) mbar_wait(&emptyC[cst], cphase ^ 1);   // consumers released the slot (use number y/NST - 1)
                // horizontal box sum: window of CPT + NX columns = own columns (registers) + NX right neighbours (Vs)
                uint32_t w[CPT + NX][4];
#pragma unroll
                for (int k = 0; k < CPT; k++) { w[k][0] = V[k][0]; w[k][1] = V[k][1]; w[k][2] = V[k][2]; w[k][3] = V[k][3]; }
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    const uint4 v = *(const uint4*)(Vs + (size_t)(CPT * xg + CPT + j) * D + 8 * q);
                    w[CPT + j][0] = v.x; w[CPT + j][1] = v.y; w[CPT + j][2] = v.z; w[CPT + j][3] = v.w;
                }
                uint32_t acc[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    acc[r] = w[0][r];
#pragma unroll
                    for (int j = 1; j <= NX; j++) acc[r] = __vadd2(acc[r], w[j][r]);
                }
                int16_t* ct = ringC + (size_t)cst * COLS * D;
#pragma unroll
                for (int k = 0; k < CPT; k++) {
                    if (k > 0) {
#pragma unroll
                        for (int r = 0; r < 4; r++) acc[r] = __vsub2(__vadd2(acc[r], w[k + NX][r]), w[k - 1][r]);
                    }
                    *(uint4*)(ct + (size_t)(CPT * xg + k) * D + 8 * q) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
                named_bar_sync(3, CF_PROD);   // tile complete (and every read of Vs is done)
                if (tid == 0) {
                    bulk_s2g(C + (((size_t)b * g.crows + y) * g.W1 + x0) * D, ct, tile_bytes);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    mbar_arrive(&fullC[cst]);
                }
                if (++cst == CF_NST) { cst = 0; cphase ^= 1; }
            }
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }

    // =============================== consumer: the three top-down paths ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(CfRegs<CF_PROD, COLS * G / 8>::cons));
    const int ct = threadIdx.x - CF_PROD;
    const int gidx = ct / G, lg = ct & (G - 1);
    const int x = x0 + gidx;
    const bool active = x < g.W1;
    const int xl = min(x, g.W1 - 1);
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2 , g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    const size_t lane_off = (size_t)lg * 2 * NP;
    const int HW = G * NW;
    auto Hptr = [&](int strip, int buf, int which) {
        return halo + ((((size_t)f * nstrips + strip) * 2 + buf) * 2 + which) * HW + (size_t)lg * NW;
    };
    int16_t* Sb = S + (size_t)b * g.H * g.W1 * D + (size_t)xl * D + lane_off;
    const size_t rowstride = (size_t)g.W1 * D;
    const bool left_edge = gidx == 0 && s > 0, right_edge = gidx == COLS - 1 && s + 1 < nstrips;
    // the +P2 bias of OpenCV's C is folded out of the stored volume, exactly as in k_cost / k_path_v3
    uint32_t Pv[NP], pa[NP], pb[NP], c[NP], acc[NP];
    uint32_t mv = 0, ma = 0, mb = 0;
    unsigned long long hq[NW];
#pragma unroll
    for (int k = 0; k < NP; k++) Pv[k] = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) hq[k] = 0;
    int stg = 0;
    uint32_t ring_phase = 0;
    for (int t = 0; t < g.H; t++) {
        const int cur = t & 1, prv = cur ^ 1;
        mbar_wait(&fullC[stg], ring_phase);
        load_regs<NP>(ringC + ((size_t)stg * COLS + gidx) * D + lane_off, c);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&emptyC[stg]);   // this warp holds its C values in registers
        if (left_edge && t > 0) {
#pragma unroll
            for (int k = 0; k < NP; k++) pa[k] = (uint32_t)hq[k];
            ma = (uint32_t)hq[NP];
        } else {
            load_col_sm<NP, G>(Xa + ((size_t)prv * (COLS + 2) + gidx) * D, lg, pa);
            ma = Ma[prv * (COLS + 2) + gidx];
        }
        if (right_edge && t > 0) {
#pragma unroll
            for (int k = 0; k < NP; k++) pb[k] = (uint32_t)hq[k];
            mb = (uint32_t)hq[NP];
        } else {
            load_col_sm<NP, G>(Xb + ((size_t)prv * (COLS + 2) + gidx + 2) * D, lg, pb);
            mb = Mb[prv * (COLS + 2) + gidx + 2];
        }
        if constexpr (COLS <= 16) {   // few warps per scheduler: the two diagonal steps interleaved
            __syncwarp();
            {
                uint32_t* const PP[2] = {pa, pb};
                uint32_t mmm[2] = {ma, mb};
                sgm_step_n<NP, G, 2>(PP, mmm, c, P1x2, P2x2, PADx2, lg);
                ma = mmm[0]; mb = mmm[1];
            }
            if (!active) {
#pragma unroll
                for (int k = 0; k < NP; k++) { pa[k] = 0; pb[k] = 0; }
                ma = 0; mb = 0;
            }
            if (right_edge) {
                unsigned long long* hp = Hptr(s + 1, cur, 0);
#pragma unroll
                for (int k = 0; k < NP; k++) halo_put(hp + k, pa[k], (uint32_t)(t + 1));
                halo_put(hp + NP, ma, (uint32_t)(t + 1));
            }
            if (left_edge) {
                unsigned long long* hp = Hptr(s - 1, cur, 1);
#pragma unroll
                for (int k = 0; k < NP; k++) halo_put(hp + k, pb[k], (uint32_t)(t + 1));
                halo_put(hp + NP, mb, (uint32_t)(t + 1));
            }
            store_col_sm<NP, G>(Xa + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pa);
            store_col_sm<NP, G>(Xb + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pb);
            if (lg == 0) { Ma[cur * (COLS + 2) + gidx + 1] = ma; Mb[cur * (COLS + 2) + gidx + 1] = mb; }
            if (t + 1 < g.H) {
                if (left_edge) halo_issue<NW>(Hptr(s, cur, 0), hq);
                if (right_edge) halo_issue<NW>(Hptr(s, cur, 1), hq);
            }
            sgm_step<NP, G>(Pv, mv, c, P1x2, P2x2, PADx2, lg);
#pragma unroll
            for (int k = 0; k < NP; k++) {
                acc[k] = __vminu2(__vadd2(pa[k], pb[k]), 0x7FFF7FFFu);
                acc[k] = __vminu2(__vadd2(acc[k], Pv[k]), 0x7FFF7FFFu);
            }
        } else {
        __syncwarp();
        sgm_step<NP, G>(pa, ma, c, P1x2, P2x2, PADx2, lg);
        if (!active) {
#pragma unroll
            for (int k = 0; k < NP; k++) pa[k] = 0;
            ma = 0;
        }
        if (right_edge) {
            unsigned long long* hp = Hptr(s + 1, cur, 0);
#pragma unroll
            for (int k = 0; k < NP; k++) halo_put(hp + k, pa[k], (uint32_t)(t + 1));
            halo_put(hp + NP, ma, (uint32_t)(t + 1));
        }
        store_col_sm<NP, G>(Xa + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pa);
        if (lg == 0) Ma[cur * (COLS + 2) + gidx + 1] = ma;
        sgm_step<NP, G>(pb, mb, c, P1x2, P2x2, PADx2, lg);
        if (!active) {
#pragma unroll
            for (int k = 0; k < NP; k++) pb[k] = 0;
            mb = 0;
        }
        if (left_edge) {
            unsigned long long* hp = Hptr(s - 1, cur, 1);
#pragma unroll
            for (int k = 0; k < NP; k++) halo_put(hp + k, pb[k], (uint32_t)(t + 1));
            halo_put(hp + NP, mb, (uint32_t)(t + 1));
        }
        store_col_sm<NP, G>(Xb + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pb);
        if (lg == 0) Mb[cur * (COLS + 2) + gidx + 1] = mb;
        if (t + 1 < g.H) {
            if (left_edge) halo_issue<NW>(Hptr(s, cur, 0), hq);
            if (right_edge) halo_issue<NW>(Hptr(s, cur, 1), hq);
        }
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = __vminu2(__vadd2(pa[k], pb[k]), 0x7FFF7FFFu);
        __syncwarp();
        sgm_step<NP, G>(Pv, mv, c, P1x2, P2x2, PADx2, lg);
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = __vminu2(__vadd2(acc[k], Pv[k]), 0x7FFF7FFFu);
        }
        if (active) store_regs<NP>(Sb + (size_t)t * rowstride, acc);
        if (t + 1 < g.H) {
            if (left_edge) halo_complete<NW>(Hptr(s, cur, 0), (uint32_t)(t + 1), err, hq);
            if (right_edge) halo_complete<NW>(Hptr(s, cur, 1), (uint32_t)(t + 1), err, hq);
        }
        named_bar_sync(1, CF_CONS);   // the row's diagonal states are in shared memory
        if (++stg == CF_NST) { stg = 0; ring_phase ^= 1; }
    }
}

// ---- fused cost + vertical path of MODE_SGBM_3WAY -------------------------------------------------------------------
// OpenCV's 3-way mode (the reference's own matcher, stereo_disparity.cpp:5-6, pcd_write.cpp:102-108) runs one top-down path
// per stripe besides the two horizontal ones. CTA = one COLS-column strip of one stripe of one frame: the producer warps are
// cost_producer (as in k_cost_path_v3), the consumer warps (8 lanes per column) run the path step on the C row tile out of the
// shared-memory ring and write S. No diagonal paths, hence no halos between strips, no consumer barrier and an ordinary launch;
// C crosses HBM once (written for the horizontal kernels) instead of written and read back by k_path_v.
__host__ __device__ inline CfLayout cf_layout_v1(int D, int SW2, int SH2, int COLS) {
    CfLayout L;
    L.cw = COLS + 2 * SW2;
    L.nchunks_max = (L.cw + D + 32) / 16 + 1;
    L.stage_bytes = ((L.cw * 16 + L.nchunks_max * PR_CHUNK) + 127) & ~127;
    size_t o = 0;
    L.off_ma = 0;
    L.off_ring_c = o;  o += (size_t)CF_NST * COLS * D * 2;
    o = (o + 127) & ~(size_t)127;
    L.off_stage = o;   o += (size_t)2 * L.stage_bytes;
    L.off_ring8 = o;   o += (size_t)(2 * SH2 + 1) * L.cw * D;
    o = (o + 15) & ~(size_t)15;
    L.off_vs = o;      o += (size_t)L.cw * D * 2;
    L.total = o;
    return L;
}

// CM: columns per consumer lane group (the consumer's share of this kernel's work is small: one path step per column and
// row against a BT item per column and octet, so fewer, busier consumer threads leave the SM to more producer warps).
template <int NP, int SW2, int CF_PROD, int COLS, int CM>
__global__ void __launch_bounds__(CF_PROD + COLS * 8 / CM, NP <= 8 ? (CF_PROD + COLS * 8 / CM <= 256 ? 4 : CF_PROD + COLS * 8 / CM <= 320 ? 3 : 2) : 1)
k_cost_path_v1(Geom g, const int4* __restrict__ PL, const uint8_t* __restrict__ PR, int16_t* __restrict__ C, int16_t* __restrict__ S) {
    constexpr int G = 8, D = 16 * NP;
    constexpr int CF_CONS = COLS * G / CM, CG = COLS / CM;   // consumer threads; column groups (columns gidx + h * CG, h < CM)
    static_assert(COLS % CM == 0 && CF_CONS % 32 == 0, "consumer tiling");
    extern __shared__ __align__(128) uint8_t cfsm[];
    __shared__ __align__(8) uint64_t fullC[CF_NST], emptyC[CF_NST], fullIn[2];
    const CfLayout lay = cf_layout_v1(D, SW2, g.SH2, COLS);
    const int x0 = blockIdx.x * COLS, b = blockIdx.z;
    const Seg sg = g.seg[blockIdx.y];
    const CfRows rw{sg.src_start, sg.src_end - sg.src_start, sg.src_start, g.H - 1, sg.crow0};
    for (int i = threadIdx.x; i < (int)((lay.off_vs - lay.off_ring8) / 4); i += blockDim.x) ((uint32_t*)(cfsm + lay.off_ring8))[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < CF_NST; i++) { mbar_init(&fullC[i], 1); mbar_init(&emptyC[i], CF_CONS / 32); }
        mbar_init(&fullIn[0], 1);
        mbar_init(&fullIn[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tile_bytes = (uint32_t)(min(COLS, g.W1 - x0) * D * 2);
    if (threadIdx.x < CF_PROD) {
        cost_producer<NP, SW2, CF_PROD, COLS>(g, PL, PR, C, b, x0, cfsm, lay, fullC, emptyC, fullIn, tile_bytes, rw);
        return;
    }
    const int ct = threadIdx.x - CF_PROD;
    const int gidx = ct / G, lg = ct & (G - 1);
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    const size_t lane_off = (size_t)lg * 2 * NP;
    const int16_t* ringC = (const int16_t*)(cfsm + lay.off_ring_c);
    int16_t* Sb = S + (size_t)b * g.H * g.W1 * D + lane_off;
    const size_t rowstride = (size_t)g.W1 * D;
    uint32_t Pv[CM][NP], c[CM][NP], mv[CM];
#pragma unroll
    for (int h = 0; h < CM; h++) {
        mv[h] = 0;
#pragma unroll
        for (int k = 0; k < NP; k++) Pv[h][k] = 0;
    }
    int stg = 0;
    uint32_t ring_phase = 0;
    for (int t = 0; t < rw.nrows; t++) {
        mbar_wait(&fullC[stg], ring_phase);
#pragma unroll
        for (int h = 0; h < CM; h++) load_regs<NP>(ringC + ((size_t)stg * COLS + gidx + h * CG) * D + lane_off, c[h]);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&emptyC[stg]);   // this warp holds its C values in registers
        const int y = rw.y0 + t;
#pragma unroll
        for (int h = 0; h < CM; h++) {
            sgm_step<NP, G>(Pv[h], mv[h], c[h], P1x2, P2x2, PADx2, lg);
            const int x = x0 + gidx + h * CG;
            if (x < g.W1 && y >= sg.dst_start) store_regs<NP>(Sb + (size_t)y * rowstride + (size_t)x * D, Pv[h]);
        }
        if (++stg == CF_NST) { stg = 0; ring_phase ^= 1; }
    }
}

// WTA + uniqueness + sub-pixel for one pixel whose final S is in registers (all lanes of the group
// take part; results are uniform across the group).
// STORE = false: `sm` already holds the pixel's S in natural order (a staged row tile), nothing is written.
template <int NP, int G, bool STORE = true>
__device__ __forceinline__ void wta_pixel(const uint32_t (&Sr)[NP], int lg, const Geom& g, int16_t* sm, int& disp_out,
                                          uint32_t& wta_out) {
    // first minimum: keys (S << 8 | d) built with one PRMT per half, reduced with 3-input minima
    uint32_t key = 0xFFFFFFFFu;
    const int d0 = lg * 2 * NP;
    const uint32_t dw0 = (uint32_t)d0 | ((uint32_t)(d0 + 1) << 8);   // bytes: d of the low half, d of the high half, 0, 0
#pragma unroll
    for (int k = 0; k < NP; k++) {
        const uint32_t dw = dw0 + 0x0202u * k;
        key = __vimin3_u32(key, prmt(Sr[k], dw, 0x6104), prmt(Sr[k], dw, 0x6325));
    }
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o, G));
    const int minS = (int)(key >> 8);
    int best = (int)(key & 0xFFu);
    // S[d] < thr  <=>  the uniqueness predicate of the mode; the division by 100 - uq is one
    // multiply-high with a precomputed reciprocal (exact in this range)
    const int q = 100 - g.uq;
    const uint32_t num = g.three_way ? (uint32_t)(100 * minS) : (uint32_t)(minS * 100 + q - 1);
    const int quot = q == 1 ? (int)num : (int)__umulhi(num, g.uq_magic);
    int thr = g.three_way ? (g.uq > 0 ? quot + 1 : 0) : quot;
    thr = min(thr, 0x8000);
    const uint32_t cst = 0x80008000u - pack16(thr, thr);  // per half: 0x8000 - thr, no borrow
    uint32_t nge2 = 0;   // per half: minus the number of entries >= thr
#pragma unroll
    for (int k = 0; k < NP; k++) {
        const uint32_t t = Sr[k] + cst;                     // bit15 / bit31 set <=> S >= thr
        nge2 = __vadd2(nge2, prmt_full(t, t, 0xBB99));      // sign-replicating PRMT: 0xFFFF (= -1) per flagged half
    }
    int nlt = 2 * NP + (int)(int16_t)(nge2 & 0xFFFFu) + (int)(int16_t)(nge2 >> 16);
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) nlt += __shfl_xor_sync(0xffffffffu, nlt, o, G);
    // neighbours of the minimum through shared memory
    if constexpr (STORE) {
        store_regs<NP>(sm + d0, Sr);
        __syncwarp();
    }
    if (g.tie8) {
        // cv2's computeDisparity3WaySGBM looks for the minimum with 8-lane vectors: lane r keeps the LAST d = r (mod 8) that
        // holds its minimum, min_pos() then takes the smallest position among the lanes at the global minimum. It shows only
        // with uniquenessRatio 0 (any uq >= 1 rejects a minimum tied 8 or more disparities apart), i.e. for the matcher
        // createRightMatcher makes (stereo_disparity.cpp:10); pinned against cv2 4.13.0 (oracle/sdr_oracle.c wta_row).
        int pos = 0x7FFF;
        if (lg < 8)
            for (int d = lg; d < g.D; d += 8)
                if (sm[d] == minS) pos = d;
#pragma unroll
        for (int o = G / 2; o >= 1; o >>= 1) pos = min(pos, __shfl_xor_sync(0xffffffffu, pos, o, G));
        best = pos;
    }
    const int Sm = best > 0 ? sm[best - 1] : 0, Sp = best < g.D - 1 ? sm[best + 1] : 0;
    __syncwarp();
    int nexcl = (minS < thr) + (best > 0 && Sm < thr) + (best < g.D - 1 && Sp < thr);
    // A pixel whose costs are ALL saturated (minS == SHRT_MAX: large windows and penalties, eight paths of MODE_HH) keeps cv2's
    // initial bestDisp = -1 in computeDisparitySGBM (no cost is < MAX_COST), which comes out as the invalid value
    if (nlt > nexcl || (!g.three_way && minS >= 32767)) {
        disp_out = (g.minD - 1) * 16;
        wta_out = 0xFFFFFFFFu;
        return;
    }
    int d = best * 16;
    if (best > 0 && best < g.D - 1) {
        // C truncating division ((Sm - Sp) * 16 + den) / (2 * den) by a float quotient plus an exact fix-up
        const int den2 = 2 * max(Sm + Sp - 2 * minS, 1), numr = (Sm - Sp) * 16 + (den2 >> 1);
        int qd = __float2int_rz(__fdividef((float)numr, (float)den2));
        int rem = numr - qd * den2;
        if (numr >= 0) {
            if (rem < 0) qd--;
            else if (rem >= den2) qd++;
        } else {
            if (rem > 0) qd++;
            else if (rem <= -den2) qd--;
        }
        d += qd;
    }
    disp_out = d + g.minD * 16;
    wta_out = ((uint32_t)minS << 16) | (uint32_t)best;
}

// Horizontal chains: one group per output row. dir = +1 (L->R) accumulates into S;
// dir = -1 (R->L) with WTA = true finishes S in registers and emits the raw disparity.
// S2 (optional, WTA kernel only): a second volume to add, the L->R path costs when that pass wrote them to a volume of their own
// instead of read-modify-writing S (one volume pass moved from the HBM-bound L->R kernel to this issue-bound one).
template <int NP, int G, bool WTA, int THREADS = 128>
__global__ void __launch_bounds__(THREADS) k_path_h(Geom g, const int16_t* __restrict__ C, int16_t* __restrict__ S, int dir,
                                                    int accumulate, int16_t* __restrict__ dispraw, uint32_t* __restrict__ wta,
                                                    const int16_t* __restrict__ S2 = nullptr) {
    constexpr int CPB = THREADS / G;
    __shared__ __align__(16) int16_t smS[WTA ? CPB * 2 * NP * G : 8];
    const int lane = threadIdx.x & 31, lg = lane & (G - 1);
    const int row = g.hrow0 + blockIdx.x * CPB + threadIdx.x / G;
    bool active = row < g.hrow0 + g.hrows;
    int y = min(row, g.hrow0 + g.hrows - 1);
    const int b = blockIdx.z;
    const int D = g.D;
    int si = -1;
    for (int i = 0; i < g.nseg; i++)
        if (y >= g.seg[i].dst_start && y < g.seg[i].src_end) si = i;
    if (si < 0) {   // row outside the selected stripes: run on a valid row, write nothing
        si = 0;
        y = g.seg[0].dst_start;
        active = false;
    }
    const Seg sg = g.seg[si];
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    uint32_t P[NP], c[NP], s[NP], L[NP], cn[NP], sn[NP], tn[WTA ? NP : 1], m = 0;
#pragma unroll
    for (int k = 0; k < NP; k++) { P[k] = 0; sn[k] = 0; }
#pragma unroll
    for (int k = 0; k < (WTA ? NP : 1); k++) tn[k] = 0;
    const bool two = WTA && S2 != nullptr;   // (uniform)
    const int16_t* Tb = two ? S2 + (((size_t)b * g.H + y) * g.W1) * g.D + (size_t)lg * 2 * NP : nullptr;
    const size_t lane_off = (size_t)lg * 2 * NP;
    const int16_t* Cb = C + (((size_t)b * g.crows + sg.crow0 + (y - sg.src_start)) * g.W1) * D + lane_off;
    int16_t* Sb = S + (((size_t)b * g.H + y) * g.W1) * D + lane_off;
    int x = dir > 0 ? 0 : g.W1 - 1;
    load_regs<NP>(Cb + (size_t)x * D, cn);
    if (accumulate) load_regs<NP>(Sb + (size_t)x * D, sn);
    if constexpr (WTA) {
        if (two) load_regs<NP>(Tb + (size_t)x * D, tn);
    }
    int16_t* sm = smS + (threadIdx.x / G) * 2 * NP * G;
    for (int step = 0; step < g.W1; step++) {
#pragma unroll
        for (int k = 0; k < NP; k++) { c[k] = cn[k]; s[k] = sn[k]; }
        if constexpr (WTA) {
            if (two) {
#pragma unroll
                for (int k = 0; k < NP; k++) s[k] = __vminu2(__vadd2(s[k], tn[k]), 0x7FFF7FFFu);
            }
        }
        const int xn = x + dir;
        if (step + 1 < g.W1) {
            load_regs<NP>(Cb + (size_t)xn * D, cn);
            if (accumulate) load_regs<NP>(Sb + (size_t)xn * D, sn);
            if constexpr (WTA) {
                if (two) load_regs<NP>(Tb + (size_t)xn * D, tn);
            }
        }
        sgm_step<NP, G>(P, m, c, P1x2, P2x2, PADx2, lg);
#pragma unroll
        for (int k = 0; k < NP; k++) L[k] = accumulate ? __vminu2(__vadd2(s[k], P[k]), 0x7FFF7FFFu) : P[k];
        if constexpr (WTA) {
            int dv;
            uint32_t wv;
            wta_pixel<NP, G>(L, lg, g, sm, dv, wv);
            if (active && lg == 0) {
                dispraw[((size_t)b * g.H + y) * g.W + x + g.minX1] = (int16_t)dv;
                wta[((size_t)b * g.H + y) * g.W1 + x] = wv;
            }
        } else {
            if (active) store_regs<NP>(Sb + (size_t)x * D, L);
        }
        x = xn;
    }
}

// Both horizontal paths of a row in ONE launch: two adjacent 8-lane groups of a warp run the L->R and the R->L chain of
// the same row at the same time and meet in the middle. In the first half of its walk a chain adds its L into S (read-
// modify-write, as k_path_h<.., false>); in the second half the other chain has already been there, so S + L is the final
// sum and the chain runs WTA / uniqueness / sub-pixel on it (as k_path_h<.., true>). Same five volume passes as the two
// separate kernels, but twice as many chains in flight (the step is a long dependent chain: the kernels are bound by the
// number of chains, not by issue or HBM), the WTA work is spread over both chains, and one launch less.
// The two chains of a row advance in lock step (one __syncwarp per step), which orders "partner stored S[x]" before
// "this chain loads S[x]" at the two or three pixels around the meeting point where the one-step-ahead prefetch of S
// would be stale.
template <int NP>
__device__ __forceinline__ void load_regs_l2(const int16_t* p, uint32_t (&r)[NP]) {
    if constexpr (NP % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NP / 4; i++) {
            const uint4 v = __ldcg((const uint4*)p + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
    } else if constexpr (NP % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NP / 2; i++) {
            const uint2 v = __ldcg((const uint2*)p + i);
            r[2 * i] = v.x; r[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NP; i++) r[i] = __ldcg((const unsigned int*)p + i);
    }
}

template <int NP, int G, int THREADS>
__global__ void __launch_bounds__(THREADS) k_path_hh(Geom g, const int16_t* __restrict__ C, int16_t* S, int16_t* __restrict__ dispraw,
                                                     uint32_t* __restrict__ wta) {
    constexpr int CPB = THREADS / G;   // chains per block, two per row
    // G = 8 / 16: the two chains of a row sit in one warp (__syncwarp orders them); G = 32: in the two warps of a 64-thread CTA
    static_assert((G == 8 || G == 16 || (G == 32 && THREADS == 64)) && CPB % 2 == 0, "the two chains of a row must sit in one CTA");
    __shared__ __align__(16) int16_t smS[CPB * 2 * NP * G];
    const int lane = threadIdx.x & 31, lg = lane & (G - 1);
    const int chain = blockIdx.x * CPB + threadIdx.x / G;
    const int row = g.hrow0 + (chain >> 1), dir = (chain & 1) ? -1 : +1;
    bool active = row < g.hrow0 + g.hrows;
    int y = min(row, g.hrow0 + g.hrows - 1);
    const int b = blockIdx.z;
    const int D = g.D, W1 = g.W1;
    int si = -1;
    for (int i = 0; i < g.nseg; i++)
        if (y >= g.seg[i].dst_start && y < g.seg[i].src_end) si = i;
    if (si < 0) {   // row outside the selected stripes: run on a valid row, write nothing
        si = 0;
        y = g.seg[0].dst_start;
        active = false;
    }
    const Seg sg = g.seg[si];
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    uint32_t P[NP], c[NP], s[NP], L[NP], cn[NP], sn[NP], m = 0;
#pragma unroll
    for (int k = 0; k < NP; k++) P[k] = 0;
    const size_t lane_off = (size_t)lg * 2 * NP;
    const int16_t* Cb = C + (((size_t)b * g.crows + sg.crow0 + (y - sg.src_start)) * W1) * D + lane_off;
    int16_t* Sb = S + (((size_t)b * g.H + y) * W1) * D + lane_off;
    const int h1 = (W1 + 1) / 2;    // L->R accumulates x < h1, R->L accumulates x >= h1; each finishes the other's pixels
    const int wta_from = W1 / 2;    // first step at which a chain of the warp is past its accumulating half
    int x = dir > 0 ? 0 : W1 - 1;
    load_regs<NP>(Cb + (size_t)x * D, cn);
    load_regs_l2<NP>(Sb + (size_t)x * D, sn);
    int16_t* sm = smS + (threadIdx.x / G) * 2 * NP * G;
    for (int step = 0; step < W1; step++) {
#pragma unroll
        for (int k = 0; k < NP; k++) { c[k] = cn[k]; s[k] = sn[k]; }
        const int xn = x + dir;
        if (step + 1 < W1) {
            load_regs<NP>(Cb + (size_t)xn * D, cn);
            load_regs_l2<NP>(Sb + (size_t)xn * D, sn);
        }
        sgm_step<NP, G>(P, m, c, P1x2, P2x2, PADx2, lg);
        const bool acc_half = dir > 0 ? x < h1 : x >= h1;
        if (acc_half) {
#pragma unroll
            for (int k = 0; k < NP; k++) L[k] = __vminu2(__vadd2(s[k], P[k]), 0x7FFF7FFFu);
            if (active) store_regs<NP>(Sb + (size_t)x * D, L);
        }
        // the partner's store of this step (and of every earlier step) is visible from here on
        if constexpr (G == 32) __syncthreads(); else __syncwarp();
        if (step >= wta_from) {   // uniform over the warp: every lane takes part in the shuffles of wta_pixel
            // the partner added its L to this pixel 2*step - (W1 - 1) steps ago: 0 or 1 means after the prefetch was issued
            if (!acc_half && 2 * step - (W1 - 1) < 2) load_regs_l2<NP>(Sb + (size_t)x * D, s);
#pragma unroll
            for (int k = 0; k < NP; k++) L[k] = __vminu2(__vadd2(s[k], P[k]), 0x7FFF7FFFu);
            int dv;
            uint32_t wv;
            wta_pixel<NP, G>(L, lg, g, sm, dv, wv);
            if (active && !acc_half && lg == 0) {
                dispraw[((size_t)b * g.H + y) * g.W + x + g.minX1] = (int16_t)dv;
                wta[((size_t)b * g.H + y) * W1 + x] = wv;
            }
        }
        x = xn;
    }
}

#include "sgbm_chain32.inl"

// Horizontal chains with a TMA-fed shared-memory ring (default): a warp owns 4 rows (8 lanes each); lane 0 of every row
// group stages the next chunks of K steps of its row (K*D*2 contiguous bytes of C, and of S) with bulk async copies,
// H_STAGES chunks deep, completion on a per-warp mbarrier. That keeps ~16 KB per warp in flight instead of the 2 KB a
// one-step register prefetch gives, which is what the single-chain-per-row parallelism needs to reach HBM speed.
constexpr int H_STAGES_MAX = 4;
constexpr int H_STAGES_DEFAULT = 2, H_WARPS_DEFAULT = 2, HB_THREADS_DEFAULT = 32;
constexpr bool HB_RING_DEFAULT = false;   // R->L + WTA through the TMA ring: 0.99 ms vs 0.93 ms with the register prefetch   // B200 sweep at 1280x720x128, batch 8 (tools/tune_env.py)

template <int NP, bool WTA>
__global__ void __launch_bounds__(128) k_path_h2(Geom g, const int16_t* __restrict__ C, int16_t* __restrict__ S, int dir, int accumulate,
                                                 int K, int H_STAGES, int16_t* __restrict__ dispraw, uint32_t* __restrict__ wta) {
    constexpr int G = 8, ROWS = 4, NWARP_MAX = 4;
    extern __shared__ __align__(128) uint8_t hsm[];
    __shared__ __align__(8) uint64_t full[NWARP_MAX][H_STAGES_MAX];
    __shared__ __align__(16) int16_t smS[WTA ? NWARP_MAX * ROWS * 2 * NP * G : 8];
    const int D = g.D;
    const int NWARP = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gidx = lane / G, lg = lane & (G - 1);
    const int row = g.hrow0 + (blockIdx.x * NWARP + warp) * ROWS + gidx;
    bool active = row < g.hrow0 + g.hrows;
    int y = min(row, g.hrow0 + g.hrows - 1);
    const int b = blockIdx.z;
    int si = -1;
    for (int i = 0; i < g.nseg; i++)
        if (y >= g.seg[i].dst_start && y < g.seg[i].src_end) si = i;
    if (si < 0) {   // row outside the selected stripes: run on a valid row, write nothing
        si = 0;
        y = g.seg[0].dst_start;
        active = false;
    }
    const Seg sg = g.seg[si];
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    const size_t lane_off = (size_t)lg * 2 * NP;
    const int16_t* Crow = C + (((size_t)b * g.crows + sg.crow0 + (y - sg.src_start)) * g.W1) * D;
    int16_t* Srow = S + (((size_t)b * g.H + y) * g.W1) * D;
    // ring: [warp][stage][C|S][row][K][D]
    const size_t arr_bytes = (size_t)ROWS * K * D * 2, stage_bytes = 2 * arr_bytes;
    uint8_t* ring = hsm + (size_t)warp * H_STAGES * stage_bytes;
    const int nchunks = (g.W1 + K - 1) / K;
    auto chunk_range = [&](int c, int& xlo, int& n) {   // memory range [xlo, xlo + n) of chunk c in traversal order
        if (dir > 0) { xlo = c * K; n = min(K, g.W1 - xlo); }
        else { const int xhi = g.W1 - 1 - c * K; xlo = max(xhi - K + 1, 0); n = xhi - xlo + 1; }
    };
    auto issue = [&](int c, int st) {   // lanes with lg == 0: one per row; st = c mod H_STAGES
        int xlo, n;
        chunk_range(c, xlo, n);
        uint8_t* base = ring + (size_t)st * stage_bytes + (size_t)gidx * K * D * 2;
        const uint32_t bytes = (uint32_t)n * D * 2;
        mbar_expect_tx(&full[warp][st], accumulate ? 2 * bytes : bytes);
        bulk_g2s(base, Crow + (size_t)xlo * D, bytes, &full[warp][st]);
        if (accumulate) bulk_g2s(base + arr_bytes, Srow + (size_t)xlo * D, bytes, &full[warp][st]);
    };
    if (lane == 0) {
        for (int i = 0; i < H_STAGES; i++) mbar_init(&full[warp][i], ROWS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lg == 0)
        for (int c = 0; c < H_STAGES && c < nchunks; c++) issue(c, c);
    uint32_t P[NP], c_[NP], s_[NP], L[NP], m = 0;
#pragma unroll
    for (int k = 0; k < NP; k++) { P[k] = 0; s_[k] = 0; }
    int16_t* sm = smS + (warp * ROWS + gidx) * 2 * NP * G;
    int x = dir > 0 ? 0 : g.W1 - 1;
    int st = 0;
    uint32_t ring_phase = 0;
    for (int c = 0; c < nchunks; c++) {
        int xlo, n;
        chunk_range(c, xlo, n);
        mbar_wait(&full[warp][st], ring_phase);
        const uint8_t* base = ring + (size_t)st * stage_bytes + (size_t)gidx * K * D * 2;
        for (int j = 0; j < n; j++) {
            const int jm = x - xlo;   // position of this step inside the staged run
            load_regs<NP>((const int16_t*)(base + (size_t)jm * D * 2) + lane_off, c_);
            if (accumulate) load_regs<NP>((const int16_t*)(base + arr_bytes + (size_t)jm * D * 2) + lane_off, s_);
            sgm_step<NP, G>(P, m, c_, P1x2, P2x2, PADx2, lg);
#pragma unroll
            for (int k = 0; k < NP; k++) L[k] = accumulate ? __vminu2(__vadd2(s_[k], P[k]), 0x7FFF7FFFu) : P[k];
            if constexpr (WTA) {
                int dv;
                uint32_t wv;
                wta_pixel<NP, G>(L, lg, g, sm, dv, wv);
                if (active && lg == 0) {
                    dispraw[((size_t)b * g.H + y) * g.W + x + g.minX1] = (int16_t)dv;
                    wta[((size_t)b * g.H + y) * g.W1 + x] = wv;
                }
            } else {
                if (active) store_regs<NP>(Srow + (size_t)x * D + lane_off, L);
            }
            x += dir;
        }
        __syncwarp();   // every lane is done reading this stage
        if (lg == 0 && c + H_STAGES < nchunks) issue(c + H_STAGES, st);
        if (++st == H_STAGES) { st = 0; ring_phase ^= 1; }
    }
}

// Right-view disparity + left-right check for one row. disp2 is reduced
// with atomicMin on (cost << 16 | 0xFFFF - x): smallest cost wins, ties go to the largest x,
// which is what OpenCV's descending-x "disp2cost[x2] > minS" scan produces.
__global__ void __launch_bounds__(256) k_lrcheck(Geom g, const int16_t* __restrict__ dispraw, const uint32_t* __restrict__ wta,
                                                 int16_t* __restrict__ out, size_t opitch, size_t oframe, int y0) {
    extern __shared__ uint32_t keys[];
    const int y = y0 + blockIdx.x, b = blockIdx.z;
    const int INV = (g.minD - 1) * 16;
    int16_t* orow = (int16_t*)((uint8_t*)out + (size_t)b * oframe + (size_t)y * opitch);
    // output row y holds the matcher's row ys (ys != y only for the shifted stripes of tiny MODE_SGBM_3WAY frames, Seg::shift)
    int ys = y;
    for (int i = 0; i < g.nseg; i++)
        if (g.seg[i].shift > 0 && y >= g.seg[i].dst_start && y < g.seg[i].src_end) {
            ys = y + g.seg[i].shift;
            if (ys >= g.seg[i].src_end) ys = -1;   // never written by OpenCV: stays invalid
        }
    if (ys < 0) {   // (uniform over the block)
        for (int x = threadIdx.x; x < g.W; x += blockDim.x) orow[x] = (int16_t)INV;
        return;
    }
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) keys[x] = 0xFFFFFFFFu;
    __syncthreads();
    const uint32_t* wr = wta + ((size_t)b * g.H + ys) * g.W1;
    for (int x = threadIdx.x; x < g.W1; x += blockDim.x) {
        const uint32_t w = wr[x];
        if (w != 0xFFFFFFFFu) {
            const int best = (int)(w & 0xFFFFu), minS = (int)(w >> 16);
            const int x2 = x + g.minX1 - best - g.minD;
            // cv2 updates disp2 only when the cost is below its initial MAX_COST = SHRT_MAX: a saturated minimum never enters
            if (minS < 32767 && x2 >= 0 && x2 < g.W) atomicMin(&keys[x2], ((uint32_t)minS << 16) | (uint32_t)(0xFFFF - x));
        }
    }
    __syncthreads();
    const int16_t* dr = dispraw + ((size_t)b * g.H + ys) * g.W;
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) {
        int d1 = INV;
        if (x >= g.minX1 && x < g.maxX1) {
            d1 = dr[x];
            if (d1 != INV) {
                const int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
                const int _x = x - _d, x_ = x - d_;
                bool bad0 = false, bad1 = false;
                if (_x >= 0 && _x < g.W) {
                    const uint32_t k = keys[_x];
                    const int v = k == 0xFFFFFFFFu ? INV : (0xFFFF - (int)(k & 0xFFFFu)) + g.minX1 - _x;
                    bad0 = v >= g.minD && abs(v - _d) > g.d12;
                }
                if (x_ >= 0 && x_ < g.W) {
                    const uint32_t k = keys[x_];
                    const int v = k == 0xFFFFFFFFu ? INV : (0xFFFF - (int)(k & 0xFFFFu)) + g.minX1 - x_;
                    bad1 = v >= g.minD && abs(v - d_) > g.d12;
                }
                if (bad0 && bad1) d1 = INV;
            }
        }
        orow[x] = (int16_t)d1;
    }
}

// Kernel-selection knobs of a matcher handle (sdr_sgbm_set_option). None of them changes a result: every variant is
// bit-exact (tests/test_gpu_sgbm.py runs them against the oracle); they exist for A/B measurements and for callers that
// know how many matchers share the device. 0 = the built-in default unless stated otherwise.
struct SgbmTune {
    int wave = 0;           // MODE_SGBM as two wavefront strip sweeps (k_wave_fwd + k_wave_bwd, 4 volume passes); 2: forward sweep only,
                            // the / path and <- + WTA by the per-direction kernels (validation of the forward sweep alone)
    int fused_sweep = 1;    // cooperative down/diagonal sweep (k_path_v3); 0: one launch per direction (k_path_v)
    int cost_fusion = 1;    // cost volume inside the sweep kernel (k_cost_path_v3)
    int path_hh = -1;       // both horizontal paths in one launch (k_path_hh): 1 / 0, -1 = automatic (on for a lone matcher)
    int shared_device = 0;  // 1: other matchers run on this device at the same time (what "automatic" looks at)
    int cf_cols = 0;        // strip width of the fused kernel: 16 / 32 / 64 (0 = by the number of strips in flight)
    int cf_prod = 0;        // its producer threads: 128 / 256
    int cf_fpl = 0;         // frames per cooperative launch
    int v3_stages = 0;      // TMA ring depth of k_path_v3
    int h_legacy = 0;       // register-prefetch k_path_h instead of the TMA-ring k_path_h2 for L->R
    int h_lanes = 0;        // lanes per row chain in k_path_hh: 8 / 16 / 32, 0 = by the number of rows in flight
    int h_ws = 1;           // few rows in flight: chain warps + tail warps (k_path_hh_ws) instead of k_path_hh with 16 / 32 lanes
    int h_k = 0, h_stages = 0, h_warps = 0, hb_threads = 0, hb_ring = -1, hh_threads = 0;
    int lr_split = 0;       // 1: the L->R pass writes its path costs to a volume of its own (added by the R->L + WTA kernel) instead of
                            // read-modify-writing S. Off: alone the two kernels gain (0.886 + 0.976 -> 0.608 + 1.069 ms per 8 frames),
                            // but with four handles in flight the benchmark loses 2 % (1775 -> 1740 frames/s, same box): the bytes
                            // move into the kernel that other handles' HBM-bound work used to overlap with
    int v3_lanes = 0;       // lanes per column of the sweep kernel at D = 256: 8 / 16 (0 = 16)
    int v3_cols = 0;        // strip width of the sweep kernel k_path_v3: 16 / 32 (0 = by D and the number of strips in flight)
    int v1_cm = 0;          // columns per consumer lane group of the fused cost + vertical path kernel of MODE_SGBM_3WAY: 1 / 2 / 4
};

// ---- dispatch over (NP, G): D = 2*G*NP ----------------------------------------------------
struct PathCfg {
    int NP, G;
};
// D multiple of 16 up to 256. Prefer G = 8 (four chains per warp, three shuffle steps per
// reduction); D = 16*k -> NP = k.
inline PathCfg path_cfg(int D) { return PathCfg{D / 16, 8}; }

template <int NP>
int launch_path_v(sdr_ctx* ctx, const Geom& g, int batch, const int16_t* C, int16_t* S, int dirx, int diry, int acc) {
    constexpr int G = 8;
    dim3 grid(div_up(g.W1, 128 / G), g.nseg, batch);
    k_path_v<NP, G><<<grid, 128, 0, ctx->stream>>>(g, C, S, dirx, diry, acc);
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}
template <int NP>
int launch_path_h(sdr_ctx* ctx, const SgbmTune& tn, const Geom& g, int batch, const int16_t* C, int16_t* S, int dir, int acc, bool wta_on,
                  int16_t* dispraw, uint32_t* wta, const int16_t* S2 = nullptr) {
    constexpr int G = 8;
    const bool legacy = tn.h_legacy != 0;   // A/B switch: register-prefetch version
    // tuning knobs (defaults chosen on a B200 at 1280x720x128, batch 8)
    const int env_k = tn.h_k, env_st = tn.h_stages, env_w = tn.h_warps, env_bt = tn.hb_threads;
    const int K = env_k > 0 ? env_k : std::max(1, 512 / g.D);            // steps per staged chunk: about 1 KB per row (512 B: L->R alone 0.88 -> 0.82 ms at
                                                                          // batch 8, but slower with other handles in flight, at D = 256 and for one frame)
    const int nst = env_st >= 2 && env_st <= H_STAGES_MAX ? env_st : H_STAGES_DEFAULT;
    const int nwarp = env_w == 1 || env_w == 2 || env_w == 4 ? env_w : H_WARPS_DEFAULT;
    const size_t smem = (size_t)nwarp * nst * 2 * 4 * K * g.D * 2;       // [warp][stage][C|S][row][K][D]
    // measured (1280x720x128, batch 8): L->R 1.12 -> 1.01 ms with the ring; the R->L + WTA kernel is bound by its own
    // dependent arithmetic and loses with the extra shared-memory hop (0.94 -> 1.52 ms), so it keeps the register prefetch
    const bool wta_ring = !S2 && (tn.hb_ring >= 0 ? tn.hb_ring != 0 : HB_RING_DEFAULT);   // the ring variant takes one accumulate source
    if (!legacy && (!wta_on || wta_ring) && smem <= ctx->smem_optin) {
        dim3 grid(div_up(g.hrows, 4 * nwarp), 1, batch);
        if (wta_on) {
            SDR_CUDA(ctx, cudaFuncSetAttribute(k_path_h2<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_path_h2<NP, true><<<grid, 32 * nwarp, smem, ctx->stream>>>(g, C, S, dir, acc, K, nst, dispraw, wta);
        } else {
            SDR_CUDA(ctx, cudaFuncSetAttribute(k_path_h2<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_path_h2<NP, false><<<grid, 32 * nwarp, smem, ctx->stream>>>(g, C, S, dir, acc, K, nst, dispraw, wta);
        }
        SDR_LAUNCHED(ctx);
        return SDR_OK;
    }
    const int threads = env_bt == 32 || env_bt == 64 || env_bt == 128 ? env_bt : HB_THREADS_DEFAULT;
    dim3 grid(div_up(g.hrows, threads / G), 1, batch);
    if (wta_on) {
        if (threads == 32) k_path_h<NP, G, true, 32><<<grid, 32, 0, ctx->stream>>>(g, C, S, dir, acc, dispraw, wta, S2);
        else if (threads == 64) k_path_h<NP, G, true, 64><<<grid, 64, 0, ctx->stream>>>(g, C, S, dir, acc, dispraw, wta, S2);
        else k_path_h<NP, G, true, 128><<<grid, 128, 0, ctx->stream>>>(g, C, S, dir, acc, dispraw, wta, S2);
    } else {
        if (threads == 32) k_path_h<NP, G, false, 32><<<grid, 32, 0, ctx->stream>>>(g, C, S, dir, acc, dispraw, wta);
        else if (threads == 64) k_path_h<NP, G, false, 64><<<grid, 64, 0, ctx->stream>>>(g, C, S, dir, acc, dispraw, wta);
        else k_path_h<NP, G, false, 128><<<grid, 128, 0, ctx->stream>>>(g, C, S, dir, acc, dispraw, wta);
    }
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

template <int NP>
int launch_path_hh(sdr_ctx* ctx, const SgbmTune& tn, const Geom& g, int batch, const int16_t* C, int16_t* S, int16_t* dispraw, uint32_t* wta) {
    constexpr int G = 8;
    // Few rows in flight (one frame, or one stripe of a frame split over GPUs): the kernel is bound by the length of a row's
    // chain (W1 dependent steps), not by throughput, so a row's D costs are spread over 16 or 32 lanes instead of 8 -- a
    // quarter of the arithmetic per lane and step, two more shuffle levels per reduction, four times the warps.
    {
        const long rows = (long)g.hrows * batch;
        int lanes = tn.h_lanes;
        if (lanes != 8 && lanes != 16 && lanes != 32) lanes = rows <= 1100 ? 32 : rows <= 2600 ? 16 : 8;
        if (lanes == 32 && NP % 4) lanes = 16;
        if (lanes == 16 && NP % 2) lanes = 8;
        if constexpr (NP % 4 == 0) {
            // latency mode: chains and per-pixel tails in different warps (k_path_hh_ws), as long as every row's CTA is resident at once
            // latency mode (a single frame of up to ~900 rows): beyond ~24 warps per SM the 32-lane chains, two registers per lane,
            // cost more issue slots than they save in latency (4K x 256: 4.1 -> 9.5 ms)
            if (tn.h_ws != 0 && lanes != 8 && rows <= (long)ctx->sm_count * 6) {
                k_path_hh_ws<NP / 4><<<dim3(g.hrows, 1, batch), HWS_THREADS, 0, ctx->stream>>>(g, C, S, dispraw, wta);
                SDR_LAUNCHED(ctx);
                return SDR_OK;
            }
            if (lanes == 32) {
                k_path_hh<NP / 4, 32, 64><<<dim3(g.hrows, 1, batch), 64, 0, ctx->stream>>>(g, C, S, dispraw, wta);
                SDR_LAUNCHED(ctx);
                return SDR_OK;
            }
        }
        if constexpr (NP % 2 == 0) {
            if (lanes == 16) {
                k_path_hh<NP / 2, 16, 64><<<dim3(div_up(g.hrows, 2), 1, batch), 64, 0, ctx->stream>>>(g, C, S, dispraw, wta);
                SDR_LAUNCHED(ctx);
                return SDR_OK;
            }
        }
    }
    const int env_t = tn.hh_threads;   // tuning knob
    const int threads = env_t == 32 || env_t == 64 || env_t == 128 ? env_t : 64;   // 64: 2 warps = 4 rows x 2 directions per CTA
    dim3 grid(div_up(2 * g.hrows, threads / G), 1, batch);
    if (threads == 32) k_path_hh<NP, G, 32><<<grid, 32, 0, ctx->stream>>>(g, C, S, dispraw, wta);
    else if (threads == 128) k_path_hh<NP, G, 128><<<grid, 128, 0, ctx->stream>>>(g, C, S, dispraw, wta);
    else k_path_hh<NP, G, 64><<<grid, 64, 0, ctx->stream>>>(g, C, S, dispraw, wta);
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}

struct V3Scratch {
    DevBuf halo, err;
    PinnedBuf host_err;
};

#include "sgbm_wave.inl"

// Returns SDR_OK, an error, or -1 when the cooperative grid cannot be co-resident (caller falls back).
template <int NP, int COLS, int G = 8>
int launch_path_v3c(sdr_ctx* ctx, const SgbmTune& tn, const Geom& g, int batch, const int16_t* C, int16_t* S, int diry, int acc, V3Scratch& sc,
                    int stages = 0, int* fpl_only = nullptr) {
    constexpr int MINB = NP <= 8 && COLS * G <= 256 ? 2 : 1;
    auto kern = acc ? k_path_v3<NP, MINB, true, COLS, G> : k_path_v3<NP, MINB, false, COLS, G>;
    const int nstrips = div_up(g.W1, COLS);
    const int env_stages = tn.v3_stages > 0 ? tn.v3_stages : stages;   // tuning knob, else the caller's choice
    int nst = acc ? 3 : V3_STAGES_DEFAULT;
    if (env_stages >= 2 && env_stages <= V3_STAGES_MAX) nst = acc ? std::min(env_stages, 3) : env_stages;
    const size_t smem = (size_t)4 * (COLS + 2) * g.D * sizeof(int16_t) + (size_t)4 * (COLS + 2) * sizeof(uint32_t) +
                        (size_t)(acc ? 2 * nst : nst) * COLS * g.D * sizeof(int16_t);
    if (smem > ctx->smem_optin) return -1;
    SDR_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SDR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, COLS * G, smem));
    const int cap = per_sm * ctx->sm_count;
    if (fpl_only) {   // query: frames one cooperative launch of this variant can hold
        *fpl_only = cap / nstrips;
        return SDR_OK;
    }
    if (cap < nstrips) return -1;
    const int fpl = std::min(batch, cap / nstrips);
    int rc;
    const size_t halo_bytes = (size_t)nstrips * 4 * G * (NP + 1) * sizeof(unsigned long long);   // per frame
    if ((rc = sc.halo.reserve(ctx, halo_bytes * fpl))) return rc;
    if (!sc.err.p) {
        if ((rc = sc.err.reserve(ctx, sizeof(int)))) return rc;
        if ((rc = sc.host_err.reserve(ctx, sizeof(int)))) return rc;
        *(int*)sc.host_err.p = 0;
        SDR_CUDA(ctx, cudaMemsetAsync(sc.err.p, 0, sizeof(int), ctx->stream));
    }
    for (int b0 = 0; b0 < batch; b0 += fpl) {
        const int F = std::min(fpl, batch - b0);
        SDR_CUDA(ctx, cudaMemsetAsync(sc.halo.p, 0, halo_bytes * F, ctx->stream));   // sequence numbers back to 0
        Geom gg = g;
        unsigned long long* halo = sc.halo.as<unsigned long long>();
        int* err = sc.err.as<int>();
        void* args[] = {(void*)&gg, (void*)&C, (void*)&S, (void*)&diry, (void*)&b0, (void*)&halo, (void*)&err, (void*)&nst};
        SDR_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kern, dim3(nstrips, F, 1), dim3(COLS * G, 1, 1), args, smem, ctx->stream));
        ctx->launches++;
    }
    SDR_CUDA(ctx, cudaMemcpyAsync(sc.host_err.p, sc.err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return SDR_OK;
}

// Strip width of the sweep: 32 columns (256 threads) by default, 16 columns when the 32-column strips of the whole submission
// leave SMs idle (one 1280x720x128 frame, bottom-up sweep of MODE_HH: 1.02 -> 0.89 ms). Measured and not adopted: 16 columns
// for D = 256, where a 32-column CTA has an SM to itself (1920x1080x256, 2 frames: 2.49 vs 2.55 ms -- the row chain, not
// occupancy, bounds that sweep).
template <int NP>
int launch_path_v3(sdr_ctx* ctx, const SgbmTune& tn, const Geom& g, int batch, const int16_t* C, int16_t* S, int diry, int acc, V3Scratch& sc) {
    bool narrow = (long long)div_up(g.W1, 32) * batch * 2 <= ctx->sm_count;
    int stages = 0;
    if constexpr (NP > 8) {
        // D > 128: a 32-column CTA with 8 lanes x NP registers per column has an SM to itself (150 registers x 256 threads), so a
        // cooperative launch holds sm_count / nstrips frames. Two remedies, measured at 1920x1080x256 MODE_HH:
        //  (a) a column over 16 lanes x NP/2 registers -- the register shape and step latency of D = 128, 16 warps per CTA:
        //      2 frames 2.49 -> 2.20 ms top-down, 2.73 -> 2.21 ms bottom-up;
        //  (b) 16-column strips with a two-row ring, three CTAs per SM, when that saves cooperative launches:
        //      4 frames top-down 4.97 -> 3.70 ms ((a): 4.40), bottom-up 5.46 -> 5.09 ms ((a): 4.42).
        int f32 = 0, f16 = 0, rc;
        if ((rc = launch_path_v3c<NP, 32>(ctx, tn, g, batch, C, S, diry, acc, sc, 0, &f32)) > 0) return rc;
        if ((rc = launch_path_v3c<NP, 16>(ctx, tn, g, batch, C, S, diry, acc, sc, 2, &f16)) > 0) return rc;
        const bool fewer_launches = f16 > 0 && (f32 == 0 || div_up(batch, f16) < div_up(batch, f32));
        if constexpr (NP % 8 == 0) {
            if (tn.v3_lanes != 8 && tn.v3_cols == 0 && !(fewer_launches && !acc)) {
                rc = launch_path_v3c<NP / 2, 32, 16>(ctx, tn, g, batch, C, S, diry, acc, sc);
                if (rc != -1) return rc;
            }
        }
        if (fewer_launches) { narrow = true; stages = 2; }
    }
    if (tn.v3_cols == 16) narrow = true;
    if (tn.v3_cols == 32) narrow = false;
    if (narrow) {
        const int rc = launch_path_v3c<NP, 16>(ctx, tn, g, batch, C, S, diry, acc, sc, stages);
        if (rc != -1) return rc;
    }
    return launch_path_v3c<NP, 32>(ctx, tn, g, batch, C, S, diry, acc, sc);
}

// Fused cost + top-down sweep. Returns SDR_OK, an error, or -1 when this configuration is not covered (caller runs
// k_cost followed by k_path_v3 / k_path_v instead).
template <int NP, int SW2, int CF_PROD, int COLS, int G = 8>
int launch_cost_path_v3(sdr_ctx* ctx, const SgbmTune& tn, const Geom& g, int batch, const int4* PL, const uint8_t* PR, int16_t* C, int16_t* S, V3Scratch& sc) {
    auto kern = k_cost_path_v3<NP, SW2, CF_PROD, COLS, G>;
    constexpr int CF_CONS = COLS * G;
    const int nstrips = div_up(g.W1, COLS);
    const CfLayout lay = cf_layout(g.D, SW2, g.SH2, COLS);
    if (lay.total > ctx->smem_optin) return -1;
    SDR_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
    int per_sm = 0;
    SDR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CF_PROD + CF_CONS, lay.total));
    const int cap = per_sm * ctx->sm_count;
    if (cap < nstrips) return -1;
    // frames per cooperative launch: all that fit by default; fewer (SDR_CF_FPL) leaves part of every SM to the kernels of
    // other handles / streams, which the fused kernel otherwise shuts out (it takes the whole register file at 2 CTAs/SM)
    const int env_fpl = tn.cf_fpl;
    int fpl = std::min(batch, cap / nstrips);
    if (env_fpl > 0) fpl = std::min(fpl, env_fpl);
    int rc;
    const size_t halo_bytes = (size_t)nstrips * 4 * G * (NP + 1) * sizeof(unsigned long long);   // per frame
    if ((rc = sc.halo.reserve(ctx, halo_bytes * fpl))) return rc;
    if (!sc.err.p) {
        if ((rc = sc.err.reserve(ctx, sizeof(int)))) return rc;
        if ((rc = sc.host_err.reserve(ctx, sizeof(int)))) return rc;
        *(int*)sc.host_err.p = 0;
        SDR_CUDA(ctx, cudaMemsetAsync(sc.err.p, 0, sizeof(int), ctx->stream));
    }
    for (int b0 = 0; b0 < batch; b0 += fpl) {
        const int F = std::min(fpl, batch - b0);
        SDR_CUDA(ctx, cudaMemsetAsync(sc.halo.p, 0, halo_bytes * F, ctx->stream));
        Geom gg = g;
        unsigned long long* halo = sc.halo.as<unsigned long long>();
        int* err = sc.err.as<int>();
        void* args[] = {(void*)&gg, (void*)&PL, (void*)&PR, (void*)&C, (void*)&S, (void*)&b0, (void*)&halo, (void*)&err};
        SDR_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kern, dim3(nstrips, F, 1), dim3(CF_PROD + CF_CONS, 1, 1), args, lay.total,
                                                  ctx->stream));
        ctx->launches++;
    }
    SDR_CUDA(ctx, cudaMemcpyAsync(sc.host_err.p, sc.err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return SDR_OK;
}

// Fused cost + vertical path of MODE_SGBM_3WAY. Returns SDR_OK, an error, or -1 when the configuration is not covered.
template <int NP, int SW2, int CF_PROD, int COLS, int CM>
int launch_cost_path_v1(sdr_ctx* ctx, const Geom& g, int batch, const int4* PL, const uint8_t* PR, int16_t* C, int16_t* S) {
    auto kern = k_cost_path_v1<NP, SW2, CF_PROD, COLS, CM>;
    const CfLayout lay = cf_layout_v1(g.D, SW2, g.SH2, COLS);
    if (lay.total > ctx->smem_optin) return -1;
    SDR_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
    kern<<<dim3(div_up(g.W1, COLS), g.nseg, batch), CF_PROD + COLS * 8 / CM, lay.total, ctx->stream>>>(g, PL, PR, C, S);
    SDR_LAUNCHED(ctx);
    return SDR_OK;
}
int launch_cost_path_3way(sdr_ctx* ctx, const SgbmTune& tn, const Geom& g, int batch, const int4* PL, const uint8_t* PR, int16_t* C, int16_t* S) {
    if (g.SW2 != g.SH2 || g.D % 16) return -1;
    const int np = g.D / 16;
    // columns per consumer lane group, measured on B200 (1280x720 batch 8, D = 80 / 128: 1.37 / 1.62 ms with 1, 1.05 / 1.49 ms
    // with 4; 4K x 256: 5.41 / 5.07 / 5.19 ms with 1 / 2 / 4 -- at NP = 16 four columns of state cost too many registers)
    const int cm = tn.v1_cm == 1 || tn.v1_cm == 2 || tn.v1_cm == 4 ? tn.v1_cm : (np > 8 ? 2 : 4);
#define V1_CASE(NPV, SWV, PRODV, COLSV)                                                                        \
    if (np == NPV && g.SW2 == SWV)                                                                              \
        return cm == 1   ? launch_cost_path_v1<NPV, SWV, PRODV, COLSV, 1>(ctx, g, batch, PL, PR, C, S)          \
               : cm == 2 ? launch_cost_path_v1<NPV, SWV, PRODV, COLSV, 2>(ctx, g, batch, PL, PR, C, S)          \
                         : launch_cost_path_v1<NPV, SWV, PRODV, COLSV, 4>(ctx, g, batch, PL, PR, C, S);
    V1_CASE(4, 1, 128, 32) V1_CASE(4, 2, 128, 32) V1_CASE(5, 1, 160, 32) V1_CASE(5, 2, 160, 32)
    V1_CASE(8, 1, 128, 32) V1_CASE(8, 2, 128, 32) V1_CASE(16, 1, 256, 32) V1_CASE(16, 2, 256, 32)
#undef V1_CASE
    return -1;
}

int launch_cost_path_fused(sdr_ctx* ctx, const SgbmTune& tn, const Geom& g, int batch, const int4* PL, const uint8_t* PR, int16_t* C, int16_t* S,
                           V3Scratch& sc) {
    const int env_cols = tn.cf_cols;
    if (g.nseg != 1 || g.SW2 != g.SH2 || g.seg[0].src_start != 0 || g.crows != g.H) return -1;
    const int np = g.D / 16;
    if (g.D % 16) return -1;
    if (np == 16 && tn.cf_cols == 0) {
        // D = 256: 32-column strips, a column over 16 lanes x 8 registers (512 consumer + 256 producer threads, one CTA per SM,
        // ~200 KB of shared memory): the thread and register shape of the 64-column D = 128 kernel
        if (g.SW2 == 2) return launch_cost_path_v3<8, 2, 256, 32, 16>(ctx, tn, g, batch, PL, PR, C, S, sc);
        if (g.SW2 == 1) return launch_cost_path_v3<8, 1, 256, 32, 16>(ctx, tn, g, batch, PL, PR, C, S, sc);
        return -1;
    }
    const int env_prod = tn.cf_prod;   // tuning knob: producer threads (128 | 256)
    const int prod = env_prod == 128 || env_prod == 256 ? env_prod : CF_PROD_DEFAULT;
    // D = 128: 64-column strips, one CTA per SM (512 consumer + 256 producer threads): half as many strip borders to hand
    // halos across, 2.06 -> 1.98 ms at 1280x720, batch 8. SDR_CF_COLS=32 selects the 32-column CTAs (two per SM).
    // Only when the wider strips still fill the device (batch 8 at 1280 wide: 144 CTAs); small batches keep the 32-column
    // CTAs, which give twice as many (a single 1280x720 frame: 1.4 ms against 2.0 ms).
    const bool wide_fills = env_cols == 64 || (long long)div_up(g.W1, 64) * batch * 4 >= 3LL * ctx->sm_count;
    if (env_cols != 32 && np == 8 && wide_fills) {
        int rc64 = -1;
        if (g.SW2 == 2) rc64 = launch_cost_path_v3<8, 2, 256, 64>(ctx, tn, g, batch, PL, PR, C, S, sc);
        else if (g.SW2 == 1) rc64 = launch_cost_path_v3<8, 1, 256, 64>(ctx, tn, g, batch, PL, PR, C, S, sc);
        if (rc64 != -1) return rc64;   // -1: does not fit this device, try the smaller CTAs
    }
    // Few frames in flight (the reference's usage: one frame at a time): 16-column strips put twice as many CTAs on the machine
    // (one 1280x720 frame: 72 instead of 36), each with half the work per row.
    // 8-column strips (one 1280x720 frame = 144 CTAs of 128 producer + 64 consumer threads): option only -- measured slower at
    // 1280x720x128 (1.03 -> 1.49 ms: the producers' halo columns double) and equal at 640x360x64 (profiles/experiments/README.md)
    if (env_cols == 8) {
        int rc8 = -1;
        if (np == 8 && g.SW2 == 2) rc8 = launch_cost_path_v3<8, 2, 128, 8>(ctx, tn, g, batch, PL, PR, C, S, sc);
        else if (np == 8 && g.SW2 == 1) rc8 = launch_cost_path_v3<8, 1, 128, 8>(ctx, tn, g, batch, PL, PR, C, S, sc);
        else if (np == 4 && g.SW2 == 2) rc8 = launch_cost_path_v3<4, 2, 64, 8>(ctx, tn, g, batch, PL, PR, C, S, sc);
        else if (np == 4 && g.SW2 == 1) rc8 = launch_cost_path_v3<4, 1, 64, 8>(ctx, tn, g, batch, PL, PR, C, S, sc);
        if (rc8 != -1) return rc8;
    }
    if (env_cols == 16 || (env_cols == 0 && (long long)div_up(g.W1, 32) * batch * 2 <= ctx->sm_count)) {
        int rc16 = -1;
        if (np == 8 && g.SW2 == 2) rc16 = launch_cost_path_v3<8, 2, 128, 16>(ctx, tn, g, batch, PL, PR, C, S, sc);
        else if (np == 8 && g.SW2 == 1) rc16 = launch_cost_path_v3<8, 1, 128, 16>(ctx, tn, g, batch, PL, PR, C, S, sc);
        else if (np == 4 && g.SW2 == 2) rc16 = launch_cost_path_v3<4, 2, 128, 16>(ctx, tn, g, batch, PL, PR, C, S, sc);
        else if (np == 4 && g.SW2 == 1) rc16 = launch_cost_path_v3<4, 1, 128, 16>(ctx, tn, g, batch, PL, PR, C, S, sc);
        if (rc16 != -1) return rc16;
    }
#define CF_CASE(NPV, SWV)                                                                                                  \
    if (np == NPV && g.SW2 == SWV)                                                                                         \
        return prod == 128 ? launch_cost_path_v3<NPV, SWV, 128, 32>(ctx, tn, g, batch, PL, PR, C, S, sc)                      \
                           : launch_cost_path_v3<NPV, SWV, 256, 32>(ctx, tn, g, batch, PL, PR, C, S, sc);
    CF_CASE(8, 2) CF_CASE(8, 1) CF_CASE(4, 2) CF_CASE(4, 1)
#undef CF_CASE
    return -1;
}

#define NP_SWITCH(np, CALL)                                                              \
    switch (np) {                                                                        \
        case 1: { constexpr int NPc = 1; CALL; } break;                                  \
        case 2: { constexpr int NPc = 2; CALL; } break;                                  \
        case 3: { constexpr int NPc = 3; CALL; } break;                                  \
        case 4: { constexpr int NPc = 4; CALL; } break;                                  \
        case 5: { constexpr int NPc = 5; CALL; } break;                                  \
        case 6: { constexpr int NPc = 6; CALL; } break;                                  \
        case 7: { constexpr int NPc = 7; CALL; } break;                                  \
        case 8: { constexpr int NPc = 8; CALL; } break;                                  \
        case 9: { constexpr int NPc = 9; CALL; } break;                                  \
        case 10: { constexpr int NPc = 10; CALL; } break;                                \
        case 11: { constexpr int NPc = 11; CALL; } break;                                \
        case 12: { constexpr int NPc = 12; CALL; } break;                                \
        case 13: { constexpr int NPc = 13; CALL; } break;                                \
        case 14: { constexpr int NPc = 14; CALL; } break;                                \
        case 15: { constexpr int NPc = 15; CALL; } break;                                \
        case 16: { constexpr int NPc = 16; CALL; } break;                                \
        default: rc = SDR_EINVAL;                                                        \
    }

}  // namespace

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
struct sdr_sgbm {
    sdr_ctx* ctx = nullptr;
    sdr_sgbm_params p{};
    DevBuf inL, inR, PL, PR, C, S, S2, dispraw, wta, matched, out;
    SpeckleScratch speckle;
    V3Scratch v3;
    WaveScratch wave;
    SgbmTune tune;                             // kernel-selection knobs (sdr_sgbm_set_option); results never depend on them
    int stripe_first = 0, stripe_count = -1;   // MODE_SGBM_3WAY: subset of the four stripes to match (-1 = all)
    bool skip_post = false;                    // stop after the raw matcher output (median / speckle run elsewhere)
    DevBuf post_in;
    StageTimer timer;
    Geom g{};
    int last_batch = 0;
};

int sdr_sgbm_take_sweep_error(sdr_sgbm* m);

static int validate_params(sdr_ctx* ctx, const sdr_sgbm_params* p) {
    if (p->numDisparities <= 0 || p->numDisparities % 16 != 0)
        return SDR_FAIL(ctx, SDR_EINVAL, "numDisparities must be a positive multiple of 16 (got %d)", p->numDisparities);
    if (p->numDisparities > 256)
        return SDR_FAIL(ctx, SDR_EINVAL, "numDisparities > 256 is not supported (got %d)", p->numDisparities);
    if (p->blockSize > 0 && p->blockSize % 2 == 0)
        return SDR_FAIL(ctx, SDR_EINVAL, "blockSize must be odd (got %d)", p->blockSize);
    if (p->blockSize > 11) return SDR_FAIL(ctx, SDR_EINVAL, "blockSize > 11 is not supported (got %d)", p->blockSize);
    if (p->mode != SDR_MODE_SGBM && p->mode != SDR_MODE_HH && p->mode != SDR_MODE_SGBM_3WAY)
        return SDR_FAIL(ctx, SDR_EINVAL, "unsupported mode %d (MODE_SGBM=0, MODE_HH=1, MODE_SGBM_3WAY=2)", p->mode);
    if (p->preFilterCap > 95) return SDR_FAIL(ctx, SDR_EINVAL, "preFilterCap > 95 is not supported (got %d)", p->preFilterCap);
    if (p->uniquenessRatio >= 100) return SDR_FAIL(ctx, SDR_EINVAL, "uniquenessRatio must be < 100");
    const int P1 = p->P1 > 0 ? p->P1 : 2;
    const int P2 = std::max(p->P2 > 0 ? p->P2 : 5, P1 + 1);
    const int bs = p->blockSize > 0 ? p->blockSize : 5;
    const int ft = std::max(p->preFilterCap, 15) | 1;
    if ((2 * ft + 63) * bs * bs + P2 + P1 > 32767)
        return SDR_FAIL(ctx, SDR_EINVAL, "P1/P2/blockSize/preFilterCap combination overflows CV_16S path costs");
    return SDR_OK;
}

extern "C" int sdr_sgbm_create(sdr_ctx* ctx, const sdr_sgbm_params* params, sdr_sgbm** out) {
    if (!ctx || !params || !out) return SDR_EINVAL;
    int rc = validate_params(ctx, params);
    if (rc) return rc;
    sdr_sgbm* m = new sdr_sgbm();
    m->ctx = ctx;
    m->p = *params;
    *out = m;
    return SDR_OK;
}

extern "C" void sdr_sgbm_destroy(sdr_sgbm* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    for (DevBuf* b : {&m->inL, &m->inR, &m->PL, &m->PR, &m->C, &m->S, &m->S2, &m->dispraw, &m->wta, &m->matched, &m->out,
                      &m->speckle.parent, &m->speckle.count, &m->v3.halo, &m->v3.err, &m->post_in, &m->wave.haloD, &m->wave.haloH})
        b->release();
    m->v3.host_err.release();
    m->timer.release();
    delete m;
}

static void make_geom(const sdr_sgbm_params& p, int W, int H, Geom& g) {
    g.W = W; g.H = H; g.D = p.numDisparities; g.minD = p.minDisparity;
    const int maxD = g.minD + g.D;
    g.minX1 = std::max(maxD, 0);
    g.maxX1 = W + std::min(g.minD, 0);
    g.W1 = g.maxX1 - g.minX1;
    g.uq = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
    g.d12 = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
    g.P1 = p.P1 > 0 ? p.P1 : 2;
    g.P2 = std::max(p.P2 > 0 ? p.P2 : 5, g.P1 + 1);
    g.ftzero = std::max(p.preFilterCap, 15) | 1;
    g.three_way = p.mode == SDR_MODE_SGBM_3WAY;
    g.tie8 = g.three_way && g.uq == 0;
    g.uq_magic = (uint32_t)(((1ull << 32) + (uint64_t)(100 - g.uq) - 1) / (uint64_t)(100 - g.uq));   // unused when 100 - uq == 1
    g.Wp = ((W + 15) & ~15) + 16;
    if (g.three_way) {
        g.SW2 = g.SH2 = p.blockSize > 0 ? p.blockSize / 2 : 1;
        const int nstripes = 4;
        const int stripe_sz = (H + nstripes - 1) / nstripes;
        // same double expression as OpenCV: 0.1*stripe_sz is not exact, e.g. 0.1*180 > 18
        const int overlap = (p.blockSize / 2 + 1) + (int)std::ceil(0.1 * stripe_sz);
        g.nseg = 0;
        int crow = 0;
        for (int s = 0; s < nstripes; s++) {
            Seg sg;
            sg.src_start = s == 0 ? 0 : std::max(std::min(s * stripe_sz - overlap, H), 0);
            // cv2 quirk (pinned by tests/fuzz/fuzz_oracle_vs_cv2.py over every height 1 ... 39 and window 1 ... 11): when a stripe's
            // warm-up start s * stripe_sz - overlap is negative (frames of <= 4 * (blockSize / 2 + 1) rows) OpenCV clamps the first
            // SOURCE row to 0 but keeps counting OUTPUT rows from the unclamped start: source row y is written to row y - shift, and
            // the last `shift` rows of the stripe's range are never written (they stay invalid). k_lrcheck applies the mapping.
            sg.shift = (s > 0 && s * stripe_sz - overlap < 0) ? -(s * stripe_sz - overlap) : 0;
            sg.dst_start = s * stripe_sz;
            sg.src_end = std::min((s + 1) * stripe_sz, H);
            if (sg.dst_start >= sg.src_end) continue;  // empty stripe (tiny images)
            sg.crow0 = crow;
            crow += sg.src_end - sg.src_start;
            g.seg[g.nseg++] = sg;
        }
        g.crows = crow;
    } else {
        g.SW2 = g.SH2 = (p.blockSize > 0 ? p.blockSize : 5) / 2;
        g.nseg = 1;
        g.seg[0] = Seg{0, H, 0, 0, 0};
        g.crows = H;
    }
}

static int sgbm_run(sdr_sgbm* m, int batch, const uint8_t* L, size_t lpitch, size_t lframe, const uint8_t* R, size_t rpitch,
                    size_t rframe, int W, int H, int16_t* out, size_t opitch, size_t oframe) {
    sdr_ctx* ctx = m->ctx;
    Geom& g = m->g;
    make_geom(m->p, W, H, g);
    if (m->stripe_count >= 0) {   // keep only the selected stripes; C rows are renumbered from 0
        const int first = std::min(m->stripe_first, g.nseg), last = std::min(first + m->stripe_count, g.nseg);
        if (first >= last) return SDR_FAIL(ctx, SDR_EINVAL, "stripe range [%d, %d) is empty (the image has %d stripes)", m->stripe_first,
                                           m->stripe_first + m->stripe_count, g.nseg);
        int crow = 0, k = 0;
        for (int i = first; i < last; i++, k++) {
            g.seg[k] = g.seg[i];
            g.seg[k].crow0 = crow;
            crow += g.seg[k].src_end - g.seg[k].src_start;
        }
        g.nseg = k;
        g.crows = crow;
    }
    g.hrow0 = g.seg[0].dst_start;
    g.hrows = g.seg[g.nseg - 1].src_end - g.hrow0;
    m->last_batch = batch;
    cudaStream_t st = ctx->stream;
    m->timer.begin(st);
    const int INV = (g.minD - 1) * 16;
    int rc = SDR_OK;
    // raw matcher output goes to m->matched, median to m->out (or directly to `out`)
    const size_t img16 = (size_t)W * H * sizeof(int16_t);
    if ((rc = m->matched.reserve(ctx, img16 * batch))) return rc;
    // OpenCV refuses these inputs in MODE_SGBM / MODE_HH (CV_Assert in computeDisparitySGBM: "Your input images are too
    // small for your window size and max disparity, and will result in non-deterministic SGBM results"). Its 3-way path
    // lacks the assert (it only fails on W1 <= 0) but is in the same non-deterministic regime, so it is refused too.
    {
        const int sw2 = (m->p.blockSize > 0 ? m->p.blockSize : 5) / 2;
        if (g.W1 <= 0 || W - (g.minD + g.D) <= sw2)
            return SDR_FAIL(ctx, SDR_EINVAL,
                            "input images are too small for the window size and max disparity: width - (minDisparity + "
                            "numDisparities) = %d must be greater than %d", W - (g.minD + g.D), sw2);
    }
    {
        const size_t vol = (size_t)g.W1 * g.D * sizeof(int16_t);
        if ((rc = m->PL.reserve(ctx, (size_t)batch * H * W * sizeof(int4)))) return rc;
        if ((rc = m->PR.reserve(ctx, (size_t)batch * H * (g.Wp >> 4) * PR_CHUNK))) return rc;
        if ((rc = m->C.reserve(ctx, vol * g.crows * batch))) return rc;
        if ((rc = m->S.reserve(ctx, vol * H * batch))) return rc;
        if ((rc = m->dispraw.reserve(ctx, img16 * batch))) return rc;
        if ((rc = m->wta.reserve(ctx, (size_t)batch * H * g.W1 * sizeof(uint32_t)))) return rc;
        // only the rows the selected stripes read (all of them unless a stripe subset was asked for)
        const int py0 = g.seg[0].src_start, py1 = std::min(g.seg[g.nseg - 1].src_end + g.SH2, H);
        k_prefilter<<<dim3(div_up(W, 128), py1 - py0, batch * 2), 128, 0, st>>>(g, L, lpitch, lframe, R, rpitch, rframe,
                                                                                 m->PL.as<int4>(), m->PR.as<uint8_t>(), py0);
        SDR_LAUNCHED(ctx);
        m->timer.mark(st, "prefilter");
        const int16_t* C = m->C.as<int16_t>();
        int16_t* S = m->S.as<int16_t>();
        const int np = path_cfg(g.D).NP;
        // a fused sweep of an earlier, asynchronous (device-output) call that gave up on a neighbouring strip is reported
        // here, once; the flag is cleared so the handle keeps working (calls that wait on the stream report it themselves)
        if ((rc = sdr_sgbm_take_sweep_error(m))) return rc;
        // MODE_SGBM as two wavefront sweeps (option "wave"): everything up to the raw disparity in two kernels
        bool wave_done = false;
        if (m->p.mode == SDR_MODE_SGBM && m->tune.wave != 0 && g.nseg == 1 && g.SW2 == g.SH2 && (np == 4 || np == 8) && (g.SW2 == 1 || g.SW2 == 2)) {
            const bool bwd = m->tune.wave != 2;
            rc = -1;
#define WAVE_CASE(NPV, SWV)                                                                                                                    \
    if (np == NPV && g.SW2 == SWV)                                                                                                             \
        rc = launch_wave<NPV, SWV>(ctx, g, batch, m->PL.as<int4>(), m->PR.as<uint8_t>(), m->C.as<int16_t>(), S, m->dispraw.as<int16_t>(),       \
                                   m->wta.as<uint32_t>(), m->v3, m->wave, m->timer, bwd);
            WAVE_CASE(8, 2) WAVE_CASE(8, 1) WAVE_CASE(4, 2) WAVE_CASE(4, 1)
#undef WAVE_CASE
            if (rc > 0) return rc;
            if (rc == SDR_OK) {
                wave_done = true;
                if (!bwd) {
                    NP_SWITCH(np, rc = launch_path_v<NPc>(ctx, g, batch, C, S, -1, +1, 1));
                    if (rc) return rc;
                    m->timer.mark(st, "path_v2");
                    NP_SWITCH(np, rc = launch_path_h<NPc>(ctx, m->tune, g, batch, C, S, -1, 1, true, m->dispraw.as<int16_t>(), m->wta.as<uint32_t>()));
                    if (rc) return rc;
                    m->timer.mark(st, "path_h_bwd_wta");
                }
            }
        }
        // MODE_SGBM / MODE_HH: cost volume and top-down sweep in one kernel when the configuration is covered
        bool cost_done = false;
        if (!wave_done) {
        if (m->p.mode != SDR_MODE_SGBM_3WAY && m->tune.fused_sweep && m->tune.cost_fusion) {
            rc = launch_cost_path_fused(ctx, m->tune, g, batch, m->PL.as<int4>(), m->PR.as<uint8_t>(), m->C.as<int16_t>(), S, m->v3);
            if (rc > 0) return rc;
            if (rc == SDR_OK) {
                cost_done = true;
                m->timer.mark(st, "cost_path_v_fused");
            }
        }
        bool v_done = false;   // MODE_SGBM_3WAY: the vertical path ran inside the cost kernel
        if (m->p.mode == SDR_MODE_SGBM_3WAY && m->tune.cost_fusion) {
            rc = launch_cost_path_3way(ctx, m->tune, g, batch, m->PL.as<int4>(), m->PR.as<uint8_t>(), m->C.as<int16_t>(), S);
            if (rc > 0) return rc;
            if (rc == SDR_OK) {
                cost_done = v_done = true;
                m->timer.mark(st, "cost_path_v1_fused");
            }
        }
        if (!cost_done) {
            const int n = 2 * g.SH2 + 1, DQ = g.D / 8;
            const int nthr = (256 / DQ) * DQ, XG = nthr / DQ;
            // 64-column tiles; 32-column tiles when the pixel-cost ring of a large window x disparity range does not fit shared
            // memory (D = 256 with blockSize 11: 233 KB -> 123 KB)
            int txh = COST_TXH;
            size_t smem = 0;
            for (;; txh /= 2) {
                const size_t stage_bytes = (size_t)txh * 16 + (size_t)((txh + g.D + 32) / 16 + 1) * PR_CHUNK;
                smem = COST_STAGES * stage_bytes + (size_t)n * txh * g.D + (size_t)txh * g.D * 2;
                if (smem <= ctx->smem_optin || txh == 32) break;
            }
            if (smem > ctx->smem_optin || txh <= 2 * g.SW2) return SDR_FAIL(ctx, SDR_EINVAL, "cost kernel needs %zu B shared memory", smem);
            const int items = div_up(txh, XG);
            int bands = 0;
            for (int i = 0; i < g.nseg; i++) bands += div_up(g.seg[i].src_end - g.seg[i].src_start, COST_RB);
            const dim3 grid(div_up(g.W1, txh - 2 * g.SW2), bands, batch);
#define COST_LAUNCH(IT)                                                                                                      \
    case IT:                                                                                                                 \
        if (txh == 64) {                                                                                                     \
            SDR_CUDA(ctx, cudaFuncSetAttribute(k_cost<IT, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
            k_cost<IT, 64><<<grid, nthr, smem, st>>>(g, m->PL.as<int4>(), m->PR.as<uint8_t>(), m->C.as<int16_t>());         \
        } else {                                                                                                             \
            SDR_CUDA(ctx, cudaFuncSetAttribute(k_cost<IT, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
            k_cost<IT, 32><<<grid, nthr, smem, st>>>(g, m->PL.as<int4>(), m->PR.as<uint8_t>(), m->C.as<int16_t>());         \
        }                                                                                                                    \
        break;
            switch (items) {
                COST_LAUNCH(1) COST_LAUNCH(2) COST_LAUNCH(3) COST_LAUNCH(4) COST_LAUNCH(5) COST_LAUNCH(6) COST_LAUNCH(7) COST_LAUNCH(8)
                default: return SDR_FAIL(ctx, SDR_EINVAL, "cost kernel: unsupported tiling (items=%d)", items);
            }
#undef COST_LAUNCH
            SDR_LAUNCHED(ctx);
            m->timer.mark(st, "cost");
        }
        // direction list: {dirx, diry}; the first one initialises S
        if (m->p.mode == SDR_MODE_SGBM_3WAY) {
            if (!v_done) {
                NP_SWITCH(np, rc = launch_path_v<NPc>(ctx, g, batch, C, S, 0, +1, 0));
                if (rc) return rc;
                m->timer.mark(st, "path_v");
            }
        } else {
            const int npass = m->p.mode == SDR_MODE_HH ? 2 : 1;
            bool fused = m->tune.fused_sweep != 0;
            for (int pass = cost_done ? 1 : 0; pass < npass && fused; pass++) {
                NP_SWITCH(np, rc = launch_path_v3<NPc>(ctx, m->tune, g, batch, C, S, pass == 0 ? +1 : -1, pass, m->v3));
                if (rc == -1) { fused = false; break; }   // grid cannot be co-resident: separate launches per direction
                if (rc) return rc;
                m->timer.mark(st, pass == 0 ? "path_v_fused_down" : "path_v_fused_up");
            }
            if (!fused) {
                const int dirs[6][2] = {{0, +1}, {+1, +1}, {-1, +1}, {0, -1}, {+1, -1}, {-1, -1}};
                const int nd = m->p.mode == SDR_MODE_HH ? 6 : 3;
                for (int i = cost_done ? 3 : 0; i < nd; i++) {
                    NP_SWITCH(np, rc = launch_path_v<NPc>(ctx, g, batch, C, S, dirs[i][0], dirs[i][1], i > 0));
                    if (rc) return rc;
                    m->timer.mark(st, i == 0 ? "path_v0" : i == 1 ? "path_v1" : i == 2 ? "path_v2" : i == 3 ? "path_v3" : i == 4 ? "path_v4" : "path_v5");
                }
            }
        }
        // One matcher on this device (the reference's usage): the two-direction kernel is 4-13 % faster. Several handles in
        // flight (frame-parallel throughput): the two separate kernels leave room for each other's work.
        const bool use_hh = m->tune.path_hh >= 0 ? m->tune.path_hh == 1 : m->tune.shared_device == 0;
        if (use_hh) {   // both horizontal paths in one launch
            NP_SWITCH(np, rc = launch_path_hh<NPc>(ctx, m->tune, g, batch, C, S, m->dispraw.as<int16_t>(), m->wta.as<uint32_t>()));
            if (rc) return rc;
            m->timer.mark(st, "path_h_both_wta");
        } else if (m->tune.lr_split != 0) {
            // L -> R writes its own volume (C read, L written: 2 volume passes instead of 3 in the HBM-bound kernel); the R -> L +
            // WTA kernel, bound by instruction issue, reads it besides C and S
            if ((rc = m->S2.reserve(ctx, (size_t)g.W1 * g.D * sizeof(int16_t) * H * batch))) return rc;
            int16_t* S2 = m->S2.as<int16_t>();
            NP_SWITCH(np, rc = launch_path_h<NPc>(ctx, m->tune, g, batch, C, S2, +1, 0, false, nullptr, nullptr));
            if (rc) return rc;
            m->timer.mark(st, "path_h_fwd");
            NP_SWITCH(np, rc = launch_path_h<NPc>(ctx, m->tune, g, batch, C, S, -1, 1, true, m->dispraw.as<int16_t>(), m->wta.as<uint32_t>(), S2));
            if (rc) return rc;
            m->timer.mark(st, "path_h_bwd_wta");
        } else {
            NP_SWITCH(np, rc = launch_path_h<NPc>(ctx, m->tune, g, batch, C, S, +1, 1, false, nullptr, nullptr));
            if (rc) return rc;
            m->timer.mark(st, "path_h_fwd");
            NP_SWITCH(np, rc = launch_path_h<NPc>(ctx, m->tune, g, batch, C, S, -1, 1, true, m->dispraw.as<int16_t>(), m->wta.as<uint32_t>()));
            if (rc) return rc;
            m->timer.mark(st, "path_h_bwd_wta");
        }
        }   // !wave_done
        const int ry0 = g.seg[0].dst_start, ry1 = g.seg[g.nseg - 1].src_end;   // output rows of the selected stripes
        k_lrcheck<<<dim3(ry1 - ry0, 1, batch), 256, (size_t)W * sizeof(uint32_t), st>>>(g, m->dispraw.as<int16_t>(), m->wta.as<uint32_t>(),
                                                                                          m->matched.as<int16_t>(), (size_t)W * 2, img16, ry0);
        SDR_LAUNCHED(ctx);
        m->timer.mark(st, "lrcheck");
    }
    if (m->skip_post) return SDR_OK;
    if ((rc = sdr_launch_median3(ctx, m->matched.as<int16_t>(), (size_t)W * 2, out, opitch, W, H, batch, img16, oframe))) return rc;
    m->timer.mark(st, "median3");
    if (m->p.speckleWindowSize > 0) {
        if ((rc = sdr_launch_speckle(ctx, out, opitch, W, H, batch, oframe, INV, m->p.speckleWindowSize, 16 * m->p.speckleRange,
                                     m->speckle)))
            return rc;
        m->timer.mark(st, "speckle");
    }
    return SDR_OK;
}

// The cooperative sweeps (k_path_v3, k_cost_path_v3) set a device flag instead of hanging when a neighbouring strip never
// publishes its halo; the flag is copied to pinned host memory behind every launch. Call after a host wait on the ctx
// stream: returns SDR_ECUDA (once) when the work just waited for contains such a sweep -- its S volume and disparity are
// garbage -- and re-arms the flag so the next compute starts clean.
int sdr_sgbm_take_sweep_error(sdr_sgbm* m) {
    if (!m->v3.host_err.p || !*(volatile int*)m->v3.host_err.p) return SDR_OK;
    *(volatile int*)m->v3.host_err.p = 0;
    if (m->v3.err.p) cudaMemsetAsync(m->v3.err.p, 0, sizeof(int), m->ctx->stream);
    return SDR_FAIL(m->ctx, SDR_ECUDA, "fused path sweep timed out waiting for a neighbouring strip: the disparity of that compute is invalid");
}

int sdr_sgbm_run_device(sdr_sgbm* m, int batch, const uint8_t* L, size_t lpitch, size_t lframe, const uint8_t* R, size_t rpitch,
                        size_t rframe, int W, int H, int16_t* out, size_t opitch, size_t oframe) {
    return sgbm_run(m, batch, L, lpitch, lframe, R, rpitch, rframe, W, H, out, opitch, oframe);
}
StageTimer* sdr_sgbm_timer(sdr_sgbm* m) { return &m->timer; }

extern "C" int sdr_sgbm_compute_batch(sdr_sgbm* m, int n, const sdr_image* lefts, const sdr_image* rights, sdr_image* disps) {
    if (!m || n <= 0 || !lefts || !rights || !disps) return SDR_EINVAL;
    sdr_ctx* ctx = m->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = lefts[0].cols, H = lefts[0].rows;
    int rc;
    for (int i = 0; i < n; i++) {
        if ((rc = sdr_check_image(ctx, &lefts[i], 1, 1, "left"))) return rc;
        if ((rc = sdr_check_image(ctx, &rights[i], 1, 1, "right"))) return rc;
        if ((rc = sdr_check_image(ctx, &disps[i], 1, 2, "disp"))) return rc;
        if (lefts[i].cols != W || lefts[i].rows != H || rights[i].cols != W || rights[i].rows != H || disps[i].cols != W ||
            disps[i].rows != H)
            return SDR_FAIL(ctx, SDR_ESIZE, "left/right/disp sizes differ (pair %d)", i);
    }
    // Fast path: device-resident frames laid out at a constant stride are used in place.
    auto strided = [&](const sdr_image* a, size_t& frame) {
        if (a[0].mem != SDR_MEM_DEVICE) return false;
        frame = n > 1 ? (size_t)((const uint8_t*)a[1].data - (const uint8_t*)a[0].data) : 0;
        for (int i = 0; i < n; i++) {
            if (a[i].mem != SDR_MEM_DEVICE || a[i].step_bytes != a[0].step_bytes) return false;
            if ((const uint8_t*)a[i].data != (const uint8_t*)a[0].data + (size_t)i * frame) return false;
        }
        return true;
    };
    const uint8_t *L, *R;
    size_t lp, lf, rp, rf;
    if (strided(lefts, lf)) {
        L = (const uint8_t*)lefts[0].data; lp = lefts[0].step_bytes;
    } else {
        if ((rc = m->inL.reserve(ctx, (size_t)n * W * H))) return rc;
        for (int i = 0; i < n; i++)
            SDR_CUDA(ctx, cudaMemcpy2DAsync(m->inL.as<uint8_t>() + (size_t)i * W * H, W, lefts[i].data, lefts[i].step_bytes, W, H,
                                            cudaMemcpyDefault, ctx->stream));
        L = m->inL.as<uint8_t>(); lp = W; lf = (size_t)W * H;
    }
    if (strided(rights, rf)) {
        R = (const uint8_t*)rights[0].data; rp = rights[0].step_bytes;
    } else {
        if ((rc = m->inR.reserve(ctx, (size_t)n * W * H))) return rc;
        for (int i = 0; i < n; i++)
            SDR_CUDA(ctx, cudaMemcpy2DAsync(m->inR.as<uint8_t>() + (size_t)i * W * H, W, rights[i].data, rights[i].step_bytes, W, H,
                                            cudaMemcpyDefault, ctx->stream));
        R = m->inR.as<uint8_t>(); rp = W; rf = (size_t)W * H;
    }
    size_t of;
    const bool out_direct = strided(disps, of);
    int16_t* O;
    size_t op;
    if (out_direct) {
        O = (int16_t*)disps[0].data; op = disps[0].step_bytes;
    } else {
        if ((rc = m->out.reserve(ctx, (size_t)n * W * H * 2))) return rc;
        O = m->out.as<int16_t>(); op = (size_t)W * 2; of = (size_t)W * H * 2;
    }
    if ((rc = sgbm_run(m, n, L, lp, lf, R, rp, rf, W, H, O, op, of))) return rc;
    if (!out_direct) {
        bool any_host = false;
        for (int i = 0; i < n; i++) {
            SDR_CUDA(ctx, cudaMemcpy2DAsync(disps[i].data, disps[i].step_bytes, (const uint8_t*)O + (size_t)i * of, op, (size_t)W * 2, H,
                                            cudaMemcpyDefault, ctx->stream));
            any_host |= disps[i].mem != SDR_MEM_DEVICE;
        }
        if (any_host) {
            SDR_CUDA(ctx, sdr_wait(ctx));
            if ((rc = sdr_sgbm_take_sweep_error(m))) return rc;   // this very call, not the next one, reports a broken sweep
        }
    }
    return SDR_OK;
}

extern "C" int sdr_sgbm_compute(sdr_sgbm* m, const sdr_image* left, const sdr_image* right, sdr_image* disp) {
    return sdr_sgbm_compute_batch(m, 1, left, right, disp);
}

extern "C" int sdr_sgbm_stripe_rows(const sdr_sgbm_params* params, int width, int height, int stripe, int* src_start, int* dst_start,
                                    int* row_end) {
    if (!params || width <= 0 || height <= 0) return -1;
    Geom g{};
    make_geom(*params, width, height, g);
    if (stripe >= 0 && stripe < g.nseg) {
        if (src_start) *src_start = g.seg[stripe].src_start;
        if (dst_start) *dst_start = g.seg[stripe].dst_start;
        if (row_end) *row_end = g.seg[stripe].src_end;
    }
    return g.nseg;
}

extern "C" int sdr_sgbm_match_stripes(sdr_sgbm* m, const sdr_image* left, const sdr_image* right, int first_stripe, int n_stripes,
                                      sdr_image* raw) {
    if (!m || !left || !right || !raw || first_stripe < 0 || n_stripes <= 0) return SDR_EINVAL;
    sdr_ctx* ctx = m->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, left, 1, 1, "left")) || (rc = sdr_check_image(ctx, right, 1, 1, "right")) ||
        (rc = sdr_check_image(ctx, raw, 1, 2, "raw")))
        return rc;
    const int W = left->cols, H = left->rows;
    if (right->cols != W || right->rows != H || raw->cols != W || raw->rows != H) return SDR_FAIL(ctx, SDR_ESIZE, "match_stripes: sizes differ");
    DevView vl, vr;
    if ((rc = sdr_stage_in(ctx, left, m->inL, &vl)) || (rc = sdr_stage_in(ctx, right, m->inR, &vr))) return rc;
    m->stripe_first = first_stripe;
    m->stripe_count = n_stripes;
    m->skip_post = true;
    rc = sgbm_run(m, 1, vl.p, vl.pitch, 0, vr.p, vr.pitch, 0, W, H, nullptr, 0, 0);
    m->stripe_count = -1;
    m->skip_post = false;
    if (rc) return rc;
    const int y0 = m->g.seg[0].dst_start, y1 = m->g.seg[m->g.nseg - 1].src_end;
    SDR_CUDA(ctx, cudaMemcpy2DAsync((uint8_t*)raw->data + (size_t)y0 * raw->step_bytes, raw->step_bytes, m->matched.as<int16_t>() + (size_t)y0 * W,
                                    (size_t)W * 2, (size_t)W * 2, y1 - y0, cudaMemcpyDefault, ctx->stream));
    if (raw->mem != SDR_MEM_DEVICE) SDR_CUDA(ctx, sdr_wait(ctx));
    return SDR_OK;
}

extern "C" int sdr_sgbm_postfilter(sdr_sgbm* m, const sdr_image* raw, sdr_image* disp) {
    if (!m || !raw || !disp) return SDR_EINVAL;
    sdr_ctx* ctx = m->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sdr_check_image(ctx, raw, 1, 2, "raw")) || (rc = sdr_check_image(ctx, disp, 1, 2, "disp"))) return rc;
    const int W = raw->cols, H = raw->rows;
    if (disp->cols != W || disp->rows != H) return SDR_FAIL(ctx, SDR_ESIZE, "postfilter: sizes differ");
    DevView vi, vo;
    if ((rc = sdr_stage_in(ctx, raw, m->post_in, &vi)) || (rc = sdr_stage_out_begin(ctx, disp, m->out, &vo))) return rc;
    if ((rc = sdr_launch_median3(ctx, (const int16_t*)vi.p, vi.pitch, (int16_t*)vo.p, vo.pitch, W, H, 1, 0, 0))) return rc;
    if (m->p.speckleWindowSize > 0 &&
        (rc = sdr_launch_speckle(ctx, (int16_t*)vo.p, vo.pitch, W, H, 1, 0, (m->p.minDisparity - 1) * 16, m->p.speckleWindowSize,
                                 16 * m->p.speckleRange, m->speckle)))
        return rc;
    return sdr_stage_out_end(ctx, disp, vo);
}

extern "C" int sdr_sgbm_set_option(sdr_sgbm* m, const char* name, int value) {
    if (!m || !name) return SDR_EINVAL;
    SgbmTune& t = m->tune;
    const struct { const char* n; int* p; } tab[] = {
        {"wave", &t.wave}, {"fused_sweep", &t.fused_sweep}, {"cost_fusion", &t.cost_fusion}, {"path_hh", &t.path_hh}, {"shared_device", &t.shared_device},
        {"cf_cols", &t.cf_cols}, {"cf_prod", &t.cf_prod}, {"cf_fpl", &t.cf_fpl}, {"v3_stages", &t.v3_stages}, {"h_legacy", &t.h_legacy}, {"h_lanes", &t.h_lanes}, {"h_ws", &t.h_ws},
        {"h_k", &t.h_k}, {"h_stages", &t.h_stages}, {"h_warps", &t.h_warps}, {"hb_threads", &t.hb_threads}, {"hb_ring", &t.hb_ring},
        {"hh_threads", &t.hh_threads}, {"v1_cm", &t.v1_cm}, {"v3_cols", &t.v3_cols}, {"v3_lanes", &t.v3_lanes}, {"lr_split", &t.lr_split}};
    for (const auto& e : tab)
        if (!strcmp(e.n, name)) { *e.p = value; return SDR_OK; }
    return SDR_FAIL(m->ctx, SDR_EINVAL, "sdr_sgbm_set_option: unknown option '%s'", name);
}

extern "C" int sdr_sgbm_set_profiling(sdr_sgbm* m, int enabled) {
    if (!m) return SDR_EINVAL;
    m->timer.enabled = enabled != 0;
    return SDR_OK;
}

extern "C" int sdr_sgbm_stage_times(sdr_sgbm* m, int cap, const char** names, float* ms) {
    if (!m) return 0;
    cudaSetDevice(m->ctx->device);
    return m->timer.collect(cap, names, ms);
}

extern "C" int sdr_sgbm_debug_read(sdr_sgbm* m, int what, void* host_dst, size_t bytes) {
    if (!m || !host_dst) return SDR_EINVAL;
    sdr_ctx* ctx = m->ctx;
    SDR_CUDA(ctx, cudaSetDevice(ctx->device));
    const Geom& g = m->g;
    const void* src = nullptr;
    size_t have = 0;
    if (what == 0) { src = m->C.p; have = (size_t)g.crows * g.W1 * g.D * 2; }
    else if (what == 1) { src = m->S.p; have = (size_t)g.H * g.W1 * g.D * 2; }
    else if (what == 2) { src = m->matched.p; have = (size_t)g.H * g.W * 2; }
    else return SDR_FAIL(ctx, SDR_EINVAL, "debug_read: unknown buffer %d", what);
    if (!src || bytes > have) return SDR_FAIL(ctx, SDR_ESIZE, "debug_read: %zu bytes requested, %zu available", bytes, have);
    SDR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SDR_CUDA(ctx, cudaMemcpy(host_dst, src, bytes, cudaMemcpyDeviceToHost));
    return SDR_OK;
}
