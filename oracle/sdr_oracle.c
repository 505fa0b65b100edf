/*
 * sdr_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the arithmetic the reference's hot path executes.
 * The reference (Amar-Aliaga/Stereo_Depth_Ruler) holds none of this arithmetic
 * itself: it calls un-vendored third-party libraries, so each function below
 * restates the published algorithm of the dependency and cites the reference
 * call site it stands in for:
 *
 *   OpenCV calib3d/imgproc (reference pins 4.6.0 through Ubuntu 24.04 apt,
 *   Dockerfile:1,14; the parity target is cv2 4.13.0, which is importable in
 *   the build container and is what this file is pinned against by
 *   tests/test_oracle_vs_cv2.py and the fixtures under tests/golden/):
 *     orc_init_rectify_map   <- cv::initUndistortRectifyMap  stereo_vision/src/stereo_rectifier.cpp:7-11
 *     orc_remap_bilinear     <- cv::remap(INTER_LINEAR)      stereo_vision/src/stereo_rectifier.cpp:39-40
 *     orc_bgr2gray           <- cv::cvtColor(BGR2GRAY)       stereo_vision/src/stereo_disparity.cpp:19-20, point_cloud/src/pcd_write.cpp:88-89
 *     orc_sgbm_compute       <- cv::StereoSGBM::compute      stereo_vision/src/stereo_disparity.cpp:5-6,27, point_cloud/src/pcd_write.cpp:102-111
 *     orc_median3_16s        <- cv::medianBlur(.,3) inside StereoSGBM::compute
 *     orc_filter_speckles    <- cv::filterSpeckles inside StereoSGBM::compute
 *     orc_reproject_to_3d    <- cv::reprojectImageTo3D       stereo_vision/src/stereo_disparity.cpp:76-80, point_cloud/src/pcd_write.cpp:116
 *   PCL 1.14 filters (point_cloud/CMakeLists.txt:4) -- NOT installable here and
 *   not pinned by any reference test: PARITY UNPINNED for
 *     orc_voxelgrid          <- pcl::VoxelGrid<PointXYZRGB>::filter point_cloud/src/pcd_write.cpp:123-130
 *     (together with convertCVMatToPCL, point_cloud/src/pcd_write.cpp:17-51)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library. The product path (libsdr_b200.so) never
 * links or calls it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

enum { ORC_MODE_SGBM = 0, ORC_MODE_HH = 1, ORC_MODE_SGBM_3WAY = 2 };

typedef struct {
    int minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff,
        preFilterCap, uniquenessRatio, speckleWindowSize, speckleRange, mode;
} orc_sgbm_params;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int16_t sat16(int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

/* ------------------------------------------------------------------------- */
/* A.1  initUndistortRectifyMap(K, dist(k1,k2,p1,p2,k3), R, P, CV_16SC2)       */
/* ------------------------------------------------------------------------- */
static void mat3_mul(const double* a, const double* b, double* c) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
            c[i * 3 + j] = s;
        }
}

/* LU with partial pivoting, as cv::invert(DECOMP_LU) does for a 3x3. */
static int mat3_inv_lu(const double* m, double* inv) {
    double a[3][6];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            a[i][j] = m[i * 3 + j];
            a[i][j + 3] = (i == j) ? 1.0 : 0.0;
        }
    for (int i = 0; i < 3; i++) {
        int k = i;
        for (int j = i + 1; j < 3; j++)
            if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
        if (fabs(a[k][i]) < 2.2204460492503131e-16) return -1;
        if (k != i)
            for (int j = 0; j < 6; j++) { double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
        double d = -1.0 / a[i][i];
        for (int j = i + 1; j < 3; j++) {
            double alpha = a[j][i] * d;
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
    return 0;
}

/* K 3x3, dist 5, R 3x3, P 3x4 (row-major fp64). map1: int16[H][W][2], map2: uint16[H][W]. */
ORC_API int orc_init_rectify_map(const double* K, const double* dist, const double* R,
                                 const double* P, int W, int H, int16_t* map1, uint16_t* map2) {
    double P3[9], PR[9], iR[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) P3[i * 3 + j] = P[i * 4 + j];
    mat3_mul(P3, R, PR);
    if (mat3_inv_lu(PR, iR)) return -1;
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
    for (int i = 0; i < H; i++) {
        for (int j = 0; j < W; j++) {
            volatile double _x = j * iR[0] + i * iR[1] + iR[2];
            volatile double _y = j * iR[3] + i * iR[4] + iR[5];
            volatile double _w = j * iR[6] + i * iR[7] + iR[8];
            double w = 1. / _w, x = _x * w, y = _y * w;
            double x2 = x * x, y2 = y * y;
            double r2 = x2 + y2, _2xy = 2 * x * y;
            double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
            double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
            double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
            double u = fx * xd + cx;
            double v = fy * yd + cy;
            /* saturate_cast<int>(double) = cvRound: round-half-even (cvtsd2si), INT_MIN when the value leaves the int range;
             * the integer part is then saturate_cast<short> (pinned by tests/fuzz/fuzz_oracle_vs_cv2.py on absurd calibrations) */
            int iu = fabs(u * 32.0) < 2147483648.0 ? (int)lrint(u * 32.0) : INT32_MIN;
            int iv = fabs(v * 32.0) < 2147483648.0 ? (int)lrint(v * 32.0) : INT32_MIN;
            map1[((size_t)i * W + j) * 2 + 0] = sat16(iu >> 5);
            map1[((size_t)i * W + j) * 2 + 1] = sat16(iv >> 5);
            map2[(size_t)i * W + j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* A.2  remap(INTER_LINEAR, BORDER_CONSTANT 0) on fixed-point maps, BGR2GRAY   */
/* ------------------------------------------------------------------------- */
ORC_API void orc_remap_bilinear(const uint8_t* src, size_t src_step, int src_w, int src_h, int ch,
                                const int16_t* map1, const uint16_t* map2, int W, int H,
                                uint8_t* dst, size_t dst_step) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int sx = map1[((size_t)y * W + x) * 2], sy = map1[((size_t)y * W + x) * 2 + 1];
            int m = map2[(size_t)y * W + x];
            int fx = m & 31, fy = (m >> 5) & 31;
            int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
            int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
            for (int c = 0; c < ch; c++) {
                int p[4];
                for (int t = 0; t < 4; t++) {
                    int xx = sx + (t & 1), yy = sy + (t >> 1);
                    p[t] = (xx >= 0 && xx < src_w && yy >= 0 && yy < src_h)
                               ? src[(size_t)yy * src_step + (size_t)xx * ch + c] : 0;
                }
                int v = (w00 * p[0] + w01 * p[1] + w10 * p[2] + w11 * p[3] + 16384) >> 15;
                dst[(size_t)y * dst_step + (size_t)x * ch + c] = (uint8_t)v;
            }
        }
}

ORC_API void orc_bgr2gray(const uint8_t* bgr, size_t step, int W, int H, uint8_t* gray, size_t gstep) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const uint8_t* p = bgr + (size_t)y * step + (size_t)x * 3;
            gray[(size_t)y * gstep + x] = (uint8_t)((3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15);
        }
}

/* ------------------------------------------------------------------------- */
/* A.3  calcPixelCostBT for one row -> pix[W1][D] (int16)                      */
/* ------------------------------------------------------------------------- */
typedef struct {
    int W, H, minD, maxD, D, minX1, maxX1, W1, ftzero;
    uint8_t *pL, *pR;           /* [2][W] prefiltered rows: channel 0 sobel, 1 raw */
    uint8_t *loL, *hiL, *loR, *hiR; /* [W] half-sample intervals of the current channel */
} bt_ctx;

static void bt_prefilter_row(const bt_ctx* c, const uint8_t* img, size_t step, int y, uint8_t* out) {
    const int W = c->W, H = c->H, ft = c->ftzero;
    const uint8_t* r = img + (size_t)y * step;
    const uint8_t* rn = img + (size_t)imax(y - 1, 0) * step;
    const uint8_t* rs = img + (size_t)imin(y + 1, H - 1) * step;
    for (int x = 1; x < W - 1; x++) {
        int v = 2 * (r[x + 1] - r[x - 1]) + rn[x + 1] - rn[x - 1] + rs[x + 1] - rs[x - 1];
        out[x] = (uint8_t)(iclamp(v, -ft, ft) + ft);
        out[W + x] = r[x];
    }
    out[0] = out[W - 1] = out[W] = out[2 * W - 1] = (uint8_t)ft; /* border columns: ftzero in BOTH channels */
}

static void bt_interval(const uint8_t* p, int W, uint8_t* lo, uint8_t* hi) {
    for (int x = 0; x < W; x++) {
        int v = p[x];
        int vl = x > 0 ? (v + p[x - 1]) / 2 : v;
        int vr = x < W - 1 ? (v + p[x + 1]) / 2 : v;
        lo[x] = (uint8_t)imin(imin(vl, vr), v);
        hi[x] = (uint8_t)imax(imax(vl, vr), v);
    }
}

static void bt_row(bt_ctx* c, const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep,
                   int y, int16_t* pix) {
    const int W = c->W, D = c->D;
    bt_prefilter_row(c, left, lstep, y, c->pL);
    bt_prefilter_row(c, right, rstep, y, c->pR);
    memset(pix, 0, (size_t)c->W1 * D * sizeof(int16_t));
    for (int ch = 0; ch < 2; ch++) {
        const int shift = ch == 0 ? 0 : 2;
        const uint8_t* a = c->pL + ch * W;
        const uint8_t* b = c->pR + ch * W;
        bt_interval(a, W, c->loL, c->hiL);
        bt_interval(b, W, c->loR, c->hiR);
        for (int x = c->minX1; x < c->maxX1; x++) {
            int u = a[x], u0 = c->loL[x], u1 = c->hiL[x];
            int16_t* o = pix + (size_t)(x - c->minX1) * D;
            for (int d = c->minD; d < c->maxD; d++) {
                int xr = x - d;
                int v = b[xr], v0 = c->loR[xr], v1 = c->hiR[xr];
                int c0 = imax(imax(0, u - v1), v0 - u);
                int c1 = imax(imax(0, v - u1), u0 - v);
                o[d - c->minD] = (int16_t)(o[d - c->minD] + (imin(c0, c1) >> shift));
            }
        }
    }
}

static int bt_ctx_init(bt_ctx* c, int W, int H, int minD, int D, int preFilterCap) {
    c->W = W; c->H = H; c->minD = minD; c->maxD = minD + D; c->D = D;
    c->minX1 = imax(c->maxD, 0);
    c->maxX1 = W + imin(minD, 0);
    c->W1 = c->maxX1 - c->minX1;
    c->ftzero = imax(preFilterCap, 15) | 1;
    c->pL = (uint8_t*)malloc((size_t)W * 8);
    if (!c->pL) return -1;
    c->pR = c->pL + 2 * W;
    c->loL = c->pR + 2 * W; c->hiL = c->loL + W; c->loR = c->hiL + W; c->hiR = c->loR + W;
    return 0;
}

/* hsum[x][d] = sum_{j=-SW2..SW2} pix[clamp(x+j,0,W1-1)][d]  (x clamp in COST space) */
static void hsum_row(const int16_t* pix, int W1, int D, int SW2, int16_t* hs) {
    for (int x = 0; x < W1; x++)
        for (int d = 0; d < D; d++) {
            int s = 0;
            for (int j = -SW2; j <= SW2; j++) s += pix[(size_t)iclamp(x + j, 0, W1 - 1) * D + d];
            hs[(size_t)x * D + d] = (int16_t)s;
        }
}

/* ------------------------------------------------------------------------- */
/* WTA + uniqueness + disp2 scatter + parabola for one row, then LR check.     */
/* S: [W1][D] final aggregated costs for the row. three_way selects the        */
/* SIMD-threshold uniqueness predicate of computeDisparity3WaySGBM.            */
/* ------------------------------------------------------------------------- */
typedef struct { int16_t* disp2; int16_t* disp2cost; } row_scratch;

static void wta_row(const int16_t* S, int W, int W1, int D, int minD, int minX1, int maxX1,
                    int uq, int d12, int three_way, int16_t* disp, row_scratch* rs) {
    const int INV = (minD - 1) * 16;
    for (int x = 0; x < W; x++) { disp[x] = (int16_t)INV; rs->disp2[x] = (int16_t)INV; rs->disp2cost[x] = 32767; }
    for (int x = W1 - 1; x >= 0; x--) {
        const int16_t* Sp = S + (size_t)x * D;
        int minS = 32767, best = three_way ? 0 : -1;
        for (int d = 0; d < D; d++)
            if (Sp[d] < minS) { minS = Sp[d]; best = d; }
        if (three_way) {
            /* QUIRK (pinned against cv2 4.13.0, tests/probes/diag3.py: 24 image/parameter cases, 0 differing pixels):
             * computeDisparity3WaySGBM searches the minimum with 8-lane int16 vectors (128-bit universal intrinsics in the
             * pip build). Lane r sees d = r, r+8, ...; an equal value REPLACES the lane's best position (last occurrence
             * per lane), then min_pos() takes the smallest position among the lanes that hold the global minimum.
             * Only visible with uniquenessRatio = 0 (createRightMatcher): any uq >= 1 rejects a pixel whose minimum is
             * tied 8 or more disparities apart. MODE_SGBM / MODE_HH take the first minimum (probed the same way). */
            int pick = D;
            for (int r = 0; r < 8 && r < D; r++) {
                int lm = 32767, lp = -1;
                for (int d = r; d < D; d += 8)
                    if (Sp[d] <= lm) { lm = Sp[d]; lp = d; }
                if (lm == minS && lp < pick) pick = lp;
            }
            if (pick < D) best = pick;
        }
        int d;
        if (three_way) {
            if (uq > 0) {
                int thr = (100 * minS) / (100 - uq) + 1;
                for (d = 0; d < D; d++)
                    if (Sp[d] < thr && abs(best - d) > 1) break;
                if (d < D) continue;
            }
        } else {
            for (d = 0; d < D; d++)
                if (Sp[d] * (100 - uq) < minS * 100 && abs(best - d) > 1) break;
            if (d < D) continue;
        }
        d = best;
        int x2 = x + minX1 - d - minD;
        if (x2 >= 0 && x2 < W && rs->disp2cost[x2] > minS) {
            rs->disp2cost[x2] = (int16_t)minS;
            rs->disp2[x2] = (int16_t)(d + minD);
        }
        if (0 < d && d < D - 1) {
            int denom2 = imax(Sp[d - 1] + Sp[d + 1] - 2 * Sp[d], 1);
            d = d * 16 + ((Sp[d - 1] - Sp[d + 1]) * 16 + denom2) / (denom2 * 2);
        } else
            d *= 16;
        disp[x + minX1] = (int16_t)(d + minD * 16);
    }
    for (int x = minX1; x < maxX1; x++) {
        int d1 = disp[x];
        if (d1 == INV) continue;
        int _d = d1 >> 4;
        int d_ = (d1 + 15) >> 4;
        int _x = x - _d, x_ = x - d_;
        if (0 <= _x && _x < W && rs->disp2[_x] >= minD && abs(rs->disp2[_x] - _d) > d12 &&
            0 <= x_ && x_ < W && rs->disp2[x_] >= minD && abs(rs->disp2[x_] - d_) > d12)
            disp[x] = (int16_t)INV;
    }
}

/* One path update: L[d] = C[d] + min(prev[d], prev[d-1]+P1, prev[d+1]+P1, delta) - delta.
 * C here INCLUDES the +P2 bias (as OpenCV's C buffer), delta = minPrev + P2. Returns min_d L. */
static inline int path_update(const int16_t* C, const int16_t* prev, int minPrev, int D, int P1, int P2,
                              int saturate, int16_t* L) {
    const int delta = minPrev + P2;
    int mn = 32767;
    for (int d = 0; d < D; d++) {
        int pm = d > 0 ? prev[d - 1] : 32767, pp = d < D - 1 ? prev[d + 1] : 32767;
        int v = C[d] + imin(imin(prev[d], pm + P1), imin(pp + P1, delta)) - delta;
        L[d] = saturate ? sat16(v) : (int16_t)v;
        if (L[d] < mn) mn = L[d];
    }
    return mn;
}

/* ------------------------------------------------------------------------- */
/* A.4  computeDisparitySGBM (MODE_SGBM 1 pass = 5 paths, MODE_HH 2 passes = 8) */
/* Optional debug outputs: Cvol [H][W1][D] (with +P2), Svol [H][W1][D] final S. */
/* ------------------------------------------------------------------------- */
static int sgbm_full(const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep, int W, int H,
                     const orc_sgbm_params* p, int16_t* disp, size_t dstep_elems,
                     int16_t* Cvol_out, int16_t* Svol_out, int16_t* Spre_out) {
    const int minD = p->minDisparity, D = p->numDisparities;
    const int uq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    const int d12 = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    const int P1 = p->P1 > 0 ? p->P1 : 2;
    const int P2 = imax(p->P2 > 0 ? p->P2 : 5, P1 + 1);
    const int SW2 = (p->blockSize > 0 ? p->blockSize : 5) / 2, SH2 = SW2;
    const int npasses = p->mode == ORC_MODE_HH ? 2 : 1;
    const int INV = (minD - 1) * 16;
    bt_ctx bc;
    if (bt_ctx_init(&bc, W, H, minD, D, p->preFilterCap)) return -1;
    const int W1 = bc.W1, minX1 = bc.minX1, maxX1 = bc.maxX1;
    if (W1 <= 0) {
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) disp[(size_t)y * dstep_elems + x] = (int16_t)INV;
        free(bc.pL);
        return 0;
    }
    const size_t row = (size_t)W1 * D;
    int16_t* hs = (int16_t*)malloc(row * H * sizeof(int16_t));     /* hsum for every row */
    int16_t* Cv = (int16_t*)malloc(row * H * sizeof(int16_t));     /* full C volume */
    int16_t* Sv = (int16_t*)calloc(row * H, sizeof(int16_t));      /* full S volume */
    int16_t* pix = (int16_t*)malloc(row * sizeof(int16_t));
    /* Lr ring: 2 rows x 4 directions x (W1+2) x D, minLr alike; borders x=-1, x=W1 are zero */
    const size_t lrw = (size_t)(W1 + 2) * D;
    int16_t* Lr = (int16_t*)calloc(2 * 4 * lrw, sizeof(int16_t));
    int16_t* mLr = (int16_t*)calloc((size_t)2 * 4 * (W1 + 2), sizeof(int16_t));
    int16_t* L4 = (int16_t*)malloc((size_t)2 * D * sizeof(int16_t));
    row_scratch rs;
    rs.disp2 = (int16_t*)malloc((size_t)W * 2 * sizeof(int16_t));
    rs.disp2cost = rs.disp2 + W;
    if (!hs || !Cv || !Sv || !pix || !Lr || !mLr || !L4 || !rs.disp2) return -1;

    for (int y = 0; y < H; y++) {
        bt_row(&bc, left, lstep, right, rstep, y, pix);
        hsum_row(pix, W1, D, SW2, hs + row * y);
    }
    for (int y = 0; y < H; y++) {
        int16_t* C = Cv + row * y;
        for (size_t i = 0; i < row; i++) {
            int s = P2;
            for (int k = y - SH2; k <= y + SH2; k++) s += hs[row * iclamp(k, 0, H - 1) + i];
            C[i] = (int16_t)s;
        }
    }
    if (Cvol_out) memcpy(Cvol_out, Cv, row * H * sizeof(int16_t));

#define LR(rowsel, r, x) (Lr + ((size_t)((rowsel) * 4 + (r)) * lrw) + (size_t)((x) + 1) * D)
#define MLR(rowsel, r, x) (mLr[(size_t)((rowsel) * 4 + (r)) * (W1 + 2) + (x) + 1])
    for (int pass = 1; pass <= npasses; pass++) {
        int y1, y2, dy, x1, x2, dx;
        if (pass == 1) { y1 = 0; y2 = H; dy = 1; x1 = 0; x2 = W1; dx = 1; }
        else { y1 = H - 1; y2 = -1; dy = -1; x1 = W1 - 1; x2 = -1; dx = -1; }
        memset(Lr, 0, 2 * 4 * lrw * sizeof(int16_t));
        memset(mLr, 0, (size_t)2 * 4 * (W1 + 2) * sizeof(int16_t));
        int cur = 0;
        for (int y = y1; y != y2; y += dy, cur ^= 1) {
            const int prv = cur ^ 1;
            const int16_t* C = Cv + row * y;
            int16_t* S = Sv + row * y;
            for (int x = x1; x != x2; x += dx) {
                const int16_t* Cp = C + (size_t)x * D;
                int16_t* Sp = S + (size_t)x * D;
                int16_t* L0 = LR(cur, 0, x); int16_t* L1 = LR(cur, 1, x);
                int16_t* L2 = LR(cur, 2, x); int16_t* L3 = LR(cur, 3, x);
                MLR(cur, 0, x) = (int16_t)path_update(Cp, LR(cur, 0, x - dx), MLR(cur, 0, x - dx), D, P1, P2, 0, L0);
                MLR(cur, 1, x) = (int16_t)path_update(Cp, LR(prv, 1, x - 1), MLR(prv, 1, x - 1), D, P1, P2, 0, L1);
                MLR(cur, 2, x) = (int16_t)path_update(Cp, LR(prv, 2, x), MLR(prv, 2, x), D, P1, P2, 0, L2);
                MLR(cur, 3, x) = (int16_t)path_update(Cp, LR(prv, 3, x + 1), MLR(prv, 3, x + 1), D, P1, P2, 0, L3);
                for (int d = 0; d < D; d++) Sp[d] = sat16(Sp[d] + L0[d] + L1[d] + L2[d] + L3[d]);
            }
            if (pass == npasses) {
                if (Spre_out) memcpy(Spre_out + row * y, S, row * sizeof(int16_t)); /* S before the R->L path / WTA */
                if (npasses == 1) {
                    /* 5th path, right-to-left on the same row, right border zero */
                    int16_t* prev = L4; int16_t* curL = L4 + D;
                    memset(prev, 0, D * sizeof(int16_t));
                    int minPrev = 0;
                    for (int x = W1 - 1; x >= 0; x--) {
                        int16_t* Sp = S + (size_t)x * D;
                        minPrev = path_update(C + (size_t)x * D, prev, minPrev, D, P1, P2, 0, curL);
                        for (int d = 0; d < D; d++) Sp[d] = sat16(Sp[d] + curL[d]);
                        int16_t* t = prev; prev = curL; curL = t;
                    }
                }
                wta_row(S, W, W1, D, minD, minX1, maxX1, uq, d12, 0, disp + (size_t)y * dstep_elems, &rs);
            }
            /* the ring row we are leaving must present zero borders to the next row (x=-1 and x=W1 stay 0) */
        }
    }
#undef LR
#undef MLR
    if (Svol_out) memcpy(Svol_out, Sv, row * H * sizeof(int16_t));
    free(hs); free(Cv); free(Sv); free(pix); free(Lr); free(mLr); free(L4); free(rs.disp2); free(bc.pL);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* A.5  computeDisparity3WaySGBM: 4 fixed stripes, paths L->R, top->down, R->L */
/* ------------------------------------------------------------------------- */
static int sgbm_3way(const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep, int W, int H,
                     const orc_sgbm_params* p, int16_t* disp, size_t dstep_elems, int16_t* Svol_out) {
    const int minD = p->minDisparity, D = p->numDisparities;
    const int uq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    const int d12 = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    const int P1 = p->P1 > 0 ? p->P1 : 2;
    const int P2 = imax(p->P2 > 0 ? p->P2 : 5, P1 + 1);
    const int SW2 = p->blockSize > 0 ? p->blockSize / 2 : 1, SH2 = SW2;
    const int INV = (minD - 1) * 16;
    bt_ctx bc;
    if (bt_ctx_init(&bc, W, H, minD, D, p->preFilterCap)) return -1;
    const int W1 = bc.W1, minX1 = bc.minX1, maxX1 = bc.maxX1;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) disp[(size_t)y * dstep_elems + x] = (int16_t)INV;
    if (W1 <= 0) { free(bc.pL); return 0; }
    const int nstripes = 4;
    const int stripe_sz = (int)ceil(H / (double)nstripes);
    const int overlap = (p->blockSize / 2 + 1) + (int)ceil(0.1 * stripe_sz);
    const size_t row = (size_t)W1 * D;
    int16_t* hs = (int16_t*)malloc(row * H * sizeof(int16_t));
    int16_t* pix = (int16_t*)malloc(row * sizeof(int16_t));
    int16_t* C = (int16_t*)malloc(row * sizeof(int16_t));
    int16_t* top = (int16_t*)malloc(row * sizeof(int16_t));
    int16_t* topMin = (int16_t*)malloc((size_t)W1 * sizeof(int16_t));
    int16_t* leftb = (int16_t*)malloc(row * sizeof(int16_t));
    int16_t* S = (int16_t*)malloc(row * sizeof(int16_t));
    int16_t* rbuf = (int16_t*)malloc((size_t)2 * D * sizeof(int16_t));
    int16_t* zero = (int16_t*)calloc(D, sizeof(int16_t));
    row_scratch rs;
    rs.disp2 = (int16_t*)malloc((size_t)W * 2 * sizeof(int16_t));
    rs.disp2cost = rs.disp2 + W;
    if (!hs || !pix || !C || !top || !topMin || !leftb || !S || !rbuf || !zero || !rs.disp2) return -1;
    for (int y = 0; y < H; y++) {
        bt_row(&bc, left, lstep, right, rstep, y, pix);
        hsum_row(pix, W1, D, SW2, hs + row * y);
    }
    for (int s = 0; s < nstripes; s++) {
        int src_start = s == 0 ? 0 : imax(imin(s * stripe_sz - overlap, H), 0);
        int dst_off = s * stripe_sz;
        int src_end = imin((s + 1) * stripe_sz, H);
        memset(top, 0, row * sizeof(int16_t));
        memset(topMin, 0, (size_t)W1 * sizeof(int16_t));
        for (int y = src_start; y < src_end; y++) {
            for (size_t i = 0; i < row; i++) {
                int v = P2;
                for (int k = y - SH2; k <= y + SH2; k++) v += hs[row * iclamp(k, src_start, H - 1) + i];
                C[i] = (int16_t)v;
            }
            const int16_t* lp = zero; int leftMin = 0;
            for (int x = 0; x < W1; x++) {
                int16_t* Lc = leftb + (size_t)x * D;
                leftMin = path_update(C + (size_t)x * D, lp, leftMin, D, P1, P2, 1, Lc);
                lp = Lc;
                int16_t* T = top + (size_t)x * D;
                /* top[x] is updated in place: compute into rbuf then copy (prev = old top[x]) */
                int tm = path_update(C + (size_t)x * D, T, topMin[x], D, P1, P2, 1, rbuf);
                memcpy(T, rbuf, D * sizeof(int16_t));
                topMin[x] = (int16_t)tm;
            }
            int16_t* rp = rbuf; int16_t* rc = rbuf + D;
            memset(rp, 0, D * sizeof(int16_t));
            int rightMin = 0;
            for (int x = W1 - 1; x >= 0; x--) {
                rightMin = path_update(C + (size_t)x * D, rp, rightMin, D, P1, P2, 1, rc);
                int16_t* Sp = S + (size_t)x * D;
                for (int d = 0; d < D; d++)
                    Sp[d] = sat16(leftb[(size_t)x * D + d] + rc[d] + top[(size_t)x * D + d]);
                int16_t* t = rp; rp = rc; rc = t;
            }
            /* QUIRK of tiny frames (pinned against cv2 4.13.0 over every height 1 ... 39 and window 1 ... 11,
             * tests/fuzz/fuzz_oracle_vs_cv2.py): when s * stripe_sz - overlap is negative OpenCV clamps the first SOURCE row of the
             * stripe to 0 but keeps counting OUTPUT rows from the unclamped start, so source row y lands in row y - shift and the
             * last `shift` rows of the stripe's range are never written (they keep the invalid value). */
            const int ustart = s * stripe_sz - overlap;
            const int shift = (s > 0 && ustart < 0) ? -ustart : 0;
            const int out_row = y - shift;
            if (Svol_out && y >= dst_off) memcpy(Svol_out + row * y, S, row * sizeof(int16_t));   /* (indexed by source row) */
            if (out_row >= dst_off)
                wta_row(S, W, W1, D, minD, minX1, maxX1, uq, d12, 1, disp + (size_t)out_row * dstep_elems, &rs);
        }
    }
    free(hs); free(pix); free(C); free(top); free(topMin); free(leftb); free(S); free(rbuf); free(zero);
    free(rs.disp2); free(bc.pL);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* A.6  medianBlur(3) on CV_16S (BORDER_REPLICATE) and filterSpeckles          */
/* ------------------------------------------------------------------------- */
ORC_API void orc_median3_16s(const int16_t* src, size_t sstep, int W, int H, int16_t* dst, size_t dstep) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int16_t v[9];
            int n = 0;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++)
                    v[n++] = src[(size_t)iclamp(y + j, 0, H - 1) * sstep + iclamp(x + i, 0, W - 1)];
            for (int a = 1; a < 9; a++) { /* insertion sort */
                int16_t t = v[a]; int b = a - 1;
                while (b >= 0 && v[b] > t) { v[b + 1] = v[b]; b--; }
                v[b + 1] = t;
            }
            dst[(size_t)y * dstep + x] = v[4];
        }
}

/* In place. 4-connected components over pixels != newVal with |a-b| <= maxDiff between
 * neighbours; components of size <= maxSize become newVal. */
ORC_API int orc_filter_speckles(int16_t* img, size_t step, int W, int H, int newVal, int maxSize, int maxDiff) {
    size_t n = (size_t)W * H;
    int* label = (int*)calloc(n, sizeof(int));
    int* stack = (int*)malloc(n * sizeof(int));
    uint8_t* small = (uint8_t*)malloc(n + 1);
    if (!label || !stack || !small) return -1;
    int cur = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t i = (size_t)y * W + x;
            if (img[(size_t)y * step + x] == newVal) continue;
            if (label[i]) {
                if (small[label[i]]) img[(size_t)y * step + x] = (int16_t)newVal;
                continue;
            }
            cur++;
            int count = 0, sp = 0;
            stack[sp++] = (int)i;
            label[i] = cur;
            while (sp) {
                int q = stack[--sp];
                int qx = q % W, qy = q / W;
                int v = img[(size_t)qy * step + qx];
                count++;
                const int nx[4] = {qx + 1, qx - 1, qx, qx}, ny[4] = {qy, qy, qy + 1, qy - 1};
                for (int t = 0; t < 4; t++) {
                    if (nx[t] < 0 || nx[t] >= W || ny[t] < 0 || ny[t] >= H) continue;
                    size_t j = (size_t)ny[t] * W + nx[t];
                    int u = img[(size_t)ny[t] * step + nx[t]];
                    if (!label[j] && u != newVal && abs(u - v) <= maxDiff) { label[j] = cur; stack[sp++] = (int)j; }
                }
            }
            small[cur] = (uint8_t)(count <= maxSize);
            if (small[cur]) img[(size_t)y * step + x] = (int16_t)newVal;
        }
    free(label); free(stack); free(small);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* StereoSGBM::compute = matcher -> medianBlur(3) -> filterSpeckles            */
/* ------------------------------------------------------------------------- */
static int check_params(const orc_sgbm_params* p, int W, int H) {
    if (W <= 0 || H <= 0) return -2;
    if (p->numDisparities <= 0 || p->numDisparities % 16) return -2;
    if (p->mode != ORC_MODE_SGBM && p->mode != ORC_MODE_HH && p->mode != ORC_MODE_SGBM_3WAY) return -2;
    if (p->blockSize > 0 && p->blockSize % 2 == 0) return -2;
    /* cv2 raises for these in MODE_SGBM / MODE_HH (CV_Assert in computeDisparitySGBM, "input images are too small ...
     * non-deterministic SGBM results"); the 3-way path only fails on W1 <= 0 but is non-deterministic in the same
     * regime, so the oracle (and the CUDA path) refuse it in every mode. */
    if ((W + imin(p->minDisparity, 0)) - imax(p->minDisparity + p->numDisparities, 0) <= 0) return -2;
    if (W - (p->minDisparity + p->numDisparities) <= (p->blockSize > 0 ? p->blockSize : 5) / 2) return -2;
    return 0;
}

/* raw matcher output (no median, no speckle); optional C / final S / pre-WTA S volume dumps (MODE_SGBM / MODE_HH) */
ORC_API int orc_sgbm_match(const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep, int W, int H,
                           const orc_sgbm_params* p, int16_t* disp, int16_t* Cvol_out, int16_t* Svol_out,
                           int16_t* Spre_out) {
    int rc = check_params(p, W, H);
    if (rc) return rc;
    if (p->mode == ORC_MODE_SGBM_3WAY) return sgbm_3way(left, lstep, right, rstep, W, H, p, disp, (size_t)W, Svol_out);
    return sgbm_full(left, lstep, right, rstep, W, H, p, disp, (size_t)W, Cvol_out, Svol_out, Spre_out);
}

ORC_API int orc_sgbm_compute(const uint8_t* left, size_t lstep, const uint8_t* right, size_t rstep, int W, int H,
                             const orc_sgbm_params* p, int16_t* disp) {
    int16_t* raw = (int16_t*)malloc((size_t)W * H * sizeof(int16_t));
    if (!raw) return -1;
    int rc = orc_sgbm_match(left, lstep, right, rstep, W, H, p, raw, NULL, NULL, NULL);
    if (rc) { free(raw); return rc; }
    orc_median3_16s(raw, (size_t)W, W, H, disp, (size_t)W);
    free(raw);
    if (p->speckleWindowSize > 0)
        rc = orc_filter_speckles(disp, (size_t)W, W, H, (p->minDisparity - 1) * 16, p->speckleWindowSize,
                                 16 * p->speckleRange);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* A.7  convertTo(CV_32F, 1/16) + reprojectImageTo3D                           */
/* ------------------------------------------------------------------------- */
ORC_API void orc_disp_to_float(const int16_t* disp, int n, float* out) {
    for (int i = 0; i < n; i++) out[i] = (float)disp[i] * 0.0625f;
}

ORC_API void orc_reproject_to_3d(const float* disp, int W, int H, const double* Q, int handle_missing, float* out) {
    float minDisp = INFINITY;
    if (handle_missing)
        for (size_t i = 0; i < (size_t)W * H; i++)
            if (disp[i] < minDisp) minDisp = disp[i];
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double d = (double)disp[(size_t)y * W + x];
            double h[4];
            for (int c = 0; c < 4; c++) h[c] = Q[c * 4 + 0] * x + Q[c * 4 + 1] * y + Q[c * 4 + 2] * d + Q[c * 4 + 3];
            double iw = 1.0 / h[3];
            float* o = out + ((size_t)y * W + x) * 3;
            /* QUIRK: numerator rounded to f32 before the fp64 multiply by the reciprocal */
            for (int c = 0; c < 3; c++) o[c] = (float)((double)(float)h[c] * iw);
            if (handle_missing && fabs(d - (double)minDisp) <= 1.1920928955078125e-07) o[2] = 10000.f;
        }
}

/* ------------------------------------------------------------------------- */
/* A.8  convertCVMatToPCL + pcl::VoxelGrid<PointXYZRGB>  (PARITY UNPINNED)     */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t idx; uint32_t pt; } vox_key;
static int vox_cmp(const void* a, const void* b) {
    const vox_key* x = (const vox_key*)a; const vox_key* y = (const vox_key*)b;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return x->pt < y->pt ? -1 : (x->pt > y->pt);
}

/* points: f32[n][3]; bgr: u8[n][3] or NULL. Points with a non-finite coordinate are
 * skipped (they are the NaN points of convertCVMatToPCL). drop_mask (optional, u8[n],
 * non-zero = drop before voxelising; measurement mode B of SURVEY 8(d)).
 * Outputs (caller allocates n entries each): out_xyz f32[.][3], out_rgb u8[.][3] (r,g,b),
 * out_idx u32[.], out_cnt u32[.]. Returns number of voxels, or -3 when PCL's int32 index
 * guard trips (PCL then passes the input cloud through unchanged), -1 on allocation failure. */
ORC_API long orc_voxelgrid(const float* points, const uint8_t* bgr, const uint8_t* drop_mask, long n,
                           float leaf_x, float leaf_y, float leaf_z,
                           float* out_xyz, uint8_t* out_rgb, uint32_t* out_idx, uint32_t* out_cnt,
                           int32_t* out_minb_divb /* 6 ints or NULL */) {
    const float inv[3] = {1.0f / leaf_x, 1.0f / leaf_y, 1.0f / leaf_z};
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    long nvalid = 0;
    for (long i = 0; i < n; i++) {
        const float* p = points + i * 3;
        if (drop_mask && drop_mask[i]) continue;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        nvalid++;
        for (int c = 0; c < 3; c++) { if (p[c] < mn[c]) mn[c] = p[c]; if (p[c] > mx[c]) mx[c] = p[c]; }
    }
    if (nvalid == 0) return 0;
    int64_t dx = (int64_t)((mx[0] - mn[0]) * inv[0]) + 1;
    int64_t dy = (int64_t)((mx[1] - mn[1]) * inv[1]) + 1;
    int64_t dz = (int64_t)((mx[2] - mn[2]) * inv[2]) + 1;
    /* PCL multiplies in int64; for absurd boxes (the reference's literal leaf 0.005 on a mm cloud) the
     * product wraps. Two's-complement wrap (what an x86 PCL build does) is made explicit here. */
    if ((int64_t)((uint64_t)dx * (uint64_t)dy * (uint64_t)dz) > (int64_t)INT32_MAX) return -3;
    int minb[3], maxb[3], divb[3];
    for (int c = 0; c < 3; c++) {
        minb[c] = (int)floorf(mn[c] * inv[c]);
        maxb[c] = (int)floorf(mx[c] * inv[c]);
        divb[c] = maxb[c] - minb[c] + 1;
    }
    if (out_minb_divb) for (int c = 0; c < 3; c++) { out_minb_divb[c] = minb[c]; out_minb_divb[3 + c] = divb[c]; }
    const int mul[3] = {1, divb[0], (int)((uint32_t)divb[0] * (uint32_t)divb[1])};
    vox_key* keys = (vox_key*)malloc((size_t)nvalid * sizeof(vox_key));
    if (!keys) return -1;
    long k = 0;
    for (long i = 0; i < n; i++) {
        const float* p = points + i * 3;
        if (drop_mask && drop_mask[i]) continue;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        int ijk0 = (int)(floorf(p[0] * inv[0]) - (float)minb[0]);
        int ijk1 = (int)(floorf(p[1] * inv[1]) - (float)minb[1]);
        int ijk2 = (int)(floorf(p[2] * inv[2]) - (float)minb[2]);
        keys[k].idx = (uint32_t)ijk0 * (uint32_t)mul[0] + (uint32_t)ijk1 * (uint32_t)mul[1] + (uint32_t)ijk2 * (uint32_t)mul[2];
        keys[k].pt = (uint32_t)i;
        k++;
    }
    qsort(keys, (size_t)nvalid, sizeof(vox_key), vox_cmp);
    long nv = 0;
    for (long a = 0; a < nvalid;) {
        long b = a + 1;
        while (b < nvalid && keys[b].idx == keys[a].idx) b++;
        float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0;
        for (long t = a; t < b; t++) {
            const float* p = points + (size_t)keys[t].pt * 3;
            sx += p[0]; sy += p[1]; sz += p[2];
            if (bgr) {
                const uint8_t* c = bgr + (size_t)keys[t].pt * 3;
                sr += (float)c[2]; sg += (float)c[1]; sb += (float)c[0];
            }
        }
        float cnt = (float)(b - a);
        out_xyz[nv * 3 + 0] = sx / cnt; out_xyz[nv * 3 + 1] = sy / cnt; out_xyz[nv * 3 + 2] = sz / cnt;
        out_rgb[nv * 3 + 0] = (uint8_t)(uint32_t)(sr / cnt);
        out_rgb[nv * 3 + 1] = (uint8_t)(uint32_t)(sg / cnt);
        out_rgb[nv * 3 + 2] = (uint8_t)(uint32_t)(sb / cnt);
        out_idx[nv] = keys[a].idx;
        out_cnt[nv] = (uint32_t)(b - a);
        nv++;
        a = b;
    }
    free(keys);
    return nv;
}

/* ------------------------------------------------------------------------- */
/* Consumers of the depth image that live in the reference's own sources:      */
/*   onMouseMeasure  stereo_vision/src/stereo_displayer.cpp:41-47               */
/*   depth_coverage  stereo_vision/src/stereo_displayer.cpp:105-118             */
/* ------------------------------------------------------------------------- */
ORC_API double orc_measure_distance(const float* points, int W, int x1, int y1, int x2, int y2) {
    const float* a = points + ((size_t)y1 * W + x1) * 3;
    const float* b = points + ((size_t)y2 * W + x2) * 3;
    double s = 0;
    for (int i = 0; i < 3; i++) {
        float d = a[i] - b[i];            /* cv::Vec3f xyz1 - xyz2 */
        s += (double)d * (double)d;       /* cv::norm: squares accumulated in double */
    }
    return sqrt(s);
}

ORC_API double orc_depth_coverage(const float* points, int W, int H, long* count_out) {
    int counter = 0, num_dis = 80;
    const int total_pixels = H * W;
    for (int i = 0; i < H; ++i)
        for (int j = num_dis; j < W; ++j) {
            float z = points[((size_t)i * W + j) * 3 + 2];
            if (z >= 0.0 && z <= 12000.0 && !isnan(z)) counter++;
        }
    if (count_out) *count_out = counter;
    return ((double)counter / total_pixels) * 100;
}

/* ---- frame ingest: cv::cvtColor(COLOR_YUV2BGR_NV12 / COLOR_YUV2BGR_YUY2) of OpenCV 4.13 (imgproc color_yuv: ITU-R BT.601
 * constants, 20-bit fixed point), the conversion between a decoder / UVC camera and the BGR frame cv::VideoCapture hands the
 * reference's loops (stereo_displayer.cpp:145-156, pcd_write.cpp:59-84). Pinned exhaustively over (Y,U,V) against cv2
 * (tests/test_oracle_vs_cv2.py). fmt: 1 = NV12 (src rows = H*3/2, Y plane then interleaved UV), 2 = YUYV. ---- */
static inline uint8_t orc_sat8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
static void orc_yuv_px(int Y, int U, int V, uint8_t* bgr) {
    const int u = U - 128, v = V - 128;
    const int y = (Y - 16 > 0 ? Y - 16 : 0) * 1220542;
    bgr[0] = orc_sat8((y + (1 << 19) + 2116026 * u) >> 20);
    bgr[1] = orc_sat8((y + (1 << 19) - 852492 * v - 409993 * u) >> 20);
    bgr[2] = orc_sat8((y + (1 << 19) + 1673527 * v) >> 20);
}
ORC_API int orc_ingest_to_bgr(int fmt, const uint8_t* src, long sstep, int W, int H, uint8_t* dst, long dstep) {
    if ((W & 1) || (fmt == 1 && (H & 1)) || (fmt != 1 && fmt != 2)) return 1;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int Y, U, V;
            if (fmt == 1) {
                Y = src[(long)y * sstep + x];
                const uint8_t* c = src + (long)(H + y / 2) * sstep + (x & ~1);
                U = c[0]; V = c[1];
            } else {
                const uint8_t* r = src + (long)y * sstep + 4 * (x >> 1);
                Y = r[(x & 1) * 2]; U = r[1]; V = r[3];
            }
            orc_yuv_px(Y, U, V, dst + (long)y * dstep + 3 * x);
        }
    return 0;
}
