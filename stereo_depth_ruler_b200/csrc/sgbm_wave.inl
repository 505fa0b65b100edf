// Wavefront form of the MODE_SGBM aggregation (include file of sgbm.cu, inside its anonymous namespace).
//
// OpenCV's one-pass mode sums five paths per pixel: -> (from x-1), \ (from x-1,y-1), | (from y-1), / (from x+1,y-1) and
// <- (from x+1). A strip sweep -- one CTA owns a strip of columns of one frame and walks down the rows -- can carry every
// path whose predecessor lies in the row above or in the strip to ONE side, provided the strips on that side run ahead:
//   k_wave_fwd   strips run left to right (strip s finishes row y before strip s+1 starts it): cost volume + ->, \ and |
//   k_wave_bwd   strips run right to left: / and <-, then winner-take-all on the finished S
// Between the two kernels only C and the partial S cross HBM: written once, read once (4 volume passes; the lock-step
// sweep + two row-parallel horizontal kernels make 7). A horizontal path is a serial chain along the row, so inside a
// strip it is walked by ONE dedicated warp (32 lanes x D/32 disparities, cross-lane minimum by redux.sync: 70 clocks per
// pixel, tools/microbench/chain_step.cu), one row behind the warps that do the column-parallel paths of the same strip.
// Strips are coupled only through per-row halo slots in global memory (the state of a strip's last column: diagonal path and
// horizontal path), each 64-bit word carrying its own sequence number, so a strip waits for its upstream neighbour and
// nobody else: no lock step, no cooperative-launch requirement (upstream strips have the lower block index).

constexpr int WV_COLS = 32;        // columns per strip: two CTAs per SM
constexpr int WV_PROD = 128;       // producer threads of k_wave_fwd
constexpr int WV_CONS = WV_COLS * 8;
constexpr int WV_NSTA = 3;         // partial-S row tiles in flight between the column-parallel warps and the chain warp
constexpr int WB_NST = 5;          // k_wave_bwd: (C, S) row tiles in flight: loader 2 ahead, diagonal, chain, winner-take-all
constexpr int WB_LAG = 2;          // rows between the diagonal step and the winner-take-all of a tile

struct WvLayout {
    CfLayout cf;          // what the producer role needs (stage, pixel-cost ring, vertical sums, C ring)
    size_t off_acc, total;
};
__host__ __device__ inline WvLayout wv_layout(int D, int SW2, int SH2) {
    WvLayout W;
    CfLayout& L = W.cf;
    L.cw = WV_COLS + 2 * SW2;
    L.nchunks_max = (L.cw + D + 32) / 16 + 1;
    L.stage_bytes = ((L.cw * 16 + L.nchunks_max * PR_CHUNK) + 127) & ~127;
    size_t o = (size_t)2 * (WV_COLS + 2) * D * 2;          // Xa: state of the diagonal path, two rows
    L.off_ma = o;  o += (size_t)2 * (WV_COLS + 2) * 4;
    o = (o + 127) & ~(size_t)127;
    L.off_ring_c = o;  o += (size_t)CF_NST * WV_COLS * D * 2;
    W.off_acc = o;     o += (size_t)WV_NSTA * WV_COLS * D * 2;
    L.off_stage = o;   o += (size_t)2 * L.stage_bytes;
    L.off_ring8 = o;   o += (size_t)(2 * SH2 + 1) * L.cw * D;
    o = (o + 15) & ~(size_t)15;
    L.off_vs = o;      o += (size_t)L.cw * D * 2;
    L.total = o;
    W.total = o;
    return W;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_wave_fwd: CTA = strip s of frame f. Warps 0-3 producer (cost rows, cost_producer), warps 4-11 column-parallel paths
// (8 lanes per column: | and \), warp 12 the -> chain. Outputs: C (by the producer) and S = L| + L\ + L-> (one TMA bulk
// store per row tile, issued by the chain warp when the tile is complete).
// ---------------------------------------------------------------------------------------------------------------------
template <int NP, int SW2>
__global__ void __launch_bounds__(WV_PROD + WV_CONS + 32, 2)
k_wave_fwd(Geom g, const int4* __restrict__ PL, const uint8_t* __restrict__ PR, int16_t* __restrict__ C, int16_t* __restrict__ S, int b0,
           unsigned long long* haloD, unsigned long long* haloH, uint32_t seq0, int* err) {
    constexpr int G = 8, NW = NP + 1, D = 16 * NP, COLS = WV_COLS;
    constexpr int NPc = NP / 4, NWc = NPc + 1;
    static_assert(NP % 4 == 0, "the chain warp holds D / 64 registers per lane");
    extern __shared__ __align__(128) uint8_t cfsm[];
    __shared__ __align__(8) uint64_t fullC[CF_NST], emptyC[CF_NST], fullIn[2], accFull[WV_NSTA], accEmpty[WV_NSTA];
    const WvLayout wl = wv_layout(D, SW2, g.SH2);
    const CfLayout& lay = wl.cf;
    int16_t* Xa = (int16_t*)cfsm;
    uint32_t* Ma = (uint32_t*)(cfsm + lay.off_ma);
    int16_t* ringC = (int16_t*)(cfsm + lay.off_ring_c);
    int16_t* accS = (int16_t*)(cfsm + wl.off_acc);
    const int s = blockIdx.x, nstrips = gridDim.x, f = blockIdx.y, b = b0 + f;
    const int x0 = s * COLS;
    for (int i = threadIdx.x; i < (int)(lay.off_ring_c / 4); i += blockDim.x) ((uint32_t*)cfsm)[i] = 0;
    for (int i = threadIdx.x; i < (int)((lay.off_vs - lay.off_ring8) / 4); i += blockDim.x) ((uint32_t*)(cfsm + lay.off_ring8))[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < CF_NST; i++) { mbar_init(&fullC[i], 1); mbar_init(&emptyC[i], WV_CONS / 32 + 1); }
        for (int i = 0; i < WV_NSTA; i++) { mbar_init(&accFull[i], WV_CONS / 32); mbar_init(&accEmpty[i], 1); }
        mbar_init(&fullIn[0], 1);
        mbar_init(&fullIn[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ncols = min(COLS, g.W1 - x0);
    const uint32_t tile_bytes = (uint32_t)(ncols * D * 2);
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);

    if (threadIdx.x < WV_PROD) {
        // =============================== producer: cost rows ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
        cost_producer<NP, SW2, WV_PROD, COLS>(g, PL, PR, C, b, x0, cfsm, lay, fullC, emptyC, fullIn, tile_bytes, CfRows{0, g.H, 0, g.H - 1, 0});
        return;
    }
    if (threadIdx.x >= WV_PROD + WV_CONS) {
        // =============================== chain warp: the -> path of the strip's row ===============================
        const int lane = threadIdx.x & 31;
        const size_t loff = (size_t)lane * 2 * NPc;   // int16 offset of the lane's disparities inside a column
        auto HH = [&](int strip, int row) { return haloH + ((((size_t)f * nstrips + strip) * g.H + row) * 32 + lane) * NWc; };
        int16_t* Srow = S + ((size_t)b * g.H * g.W1 + x0) * D;
        uint32_t P[NPc], c[NPc], a[NPc], cn[NPc], an[NPc], m = 0;
        unsigned long long hq[NWc];
        int cst = 0, ast = 0;
        uint32_t cphase = 0, aphase = 0;
        if (s > 0) halo_issue<NWc>(HH(s, 0), hq);
        for (int y = 0; y < g.H; y++) {
            if (s > 0) {
                halo_complete<NWc>(HH(s, y), seq0 | (uint32_t)(y + 1), err, hq);
#pragma unroll
                for (int k = 0; k < NPc; k++) P[k] = (uint32_t)hq[k];
                m = (uint32_t)hq[NPc];
                if (y + 1 < g.H) halo_issue<NWc>(HH(s, y + 1), hq);   // usually there already: the upstream strip runs ahead
            } else {
#pragma unroll
                for (int k = 0; k < NPc; k++) P[k] = 0;
                m = 0;
            }
            mbar_wait_b(&fullC[cst], cphase);
            mbar_wait_b(&accFull[ast], aphase);
            const int16_t* ct = ringC + (size_t)cst * COLS * D + loff;
            int16_t* at = accS + (size_t)ast * COLS * D + loff;
            ld_chain<NPc>(ct, cn);
            ld_chain<NPc>(at, an);
            for (int x = 0; x < ncols; x++) {
#pragma unroll
                for (int k = 0; k < NPc; k++) { c[k] = cn[k]; a[k] = an[k]; }
                if (x + 1 < ncols) {
                    ld_chain<NPc>(ct + (size_t)(x + 1) * D, cn);
                    ld_chain<NPc>(at + (size_t)(x + 1) * D, an);
                }
                sgm_step32<NPc>(P, m, c, P1x2, P2x2, PADx2, lane);
#pragma unroll
                for (int k = 0; k < NPc; k++) a[k] = __vminu2(__vadd2(a[k], P[k]), 0x7FFF7FFFu);
                st_chain<NPc>(at + (size_t)x * D, a);
            }
            if (s + 1 < nstrips) {   // the downstream strip's chain of this row starts from here
                unsigned long long* hp = HH(s + 1, y);
#pragma unroll
                for (int k = 0; k < NPc; k++) halo_put(hp + k, P[k], seq0 | (uint32_t)(y + 1));
                halo_put(hp + NPc, m, seq0 | (uint32_t)(y + 1));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile's generic-proxy writes -> visible to the bulk store
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&emptyC[cst]);
                bulk_s2g(Srow + (size_t)y * g.W1 * D, accS + (size_t)ast * COLS * D, tile_bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store of row y-1 has read its tile
                if (y >= 1) mbar_arrive(&accEmpty[ast == 0 ? WV_NSTA - 1 : ast - 1]);
            }
            if (++cst == CF_NST) { cst = 0; cphase ^= 1; }
            if (++ast == WV_NSTA) { ast = 0; aphase ^= 1; }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }

    // =============================== column-parallel paths: \ and | ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 80;\n");
    const int ct = threadIdx.x - WV_PROD;
    const int gidx = ct / G, lg = ct & (G - 1);
    const int x = x0 + gidx;
    const bool active = x < g.W1;
    const size_t lane_off = (size_t)lg * 2 * NP;
    auto HD = [&](int strip, int row) { return haloD + ((((size_t)f * nstrips + strip) * g.H + row) * G + lg) * NW; };
    const bool left_edge = gidx == 0 && s > 0, right_edge = gidx == COLS - 1 && s + 1 < nstrips;
    uint32_t Pv[NP], pa[NP], c[NP], acc[NP];
    uint32_t mv = 0, ma = 0;
    unsigned long long hq[NW];
#pragma unroll
    for (int k = 0; k < NP; k++) Pv[k] = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) hq[k] = 0;
    int stg = 0, ast = 0;
    uint32_t ring_phase = 0, aphase = 0;
    for (int t = 0; t < g.H; t++) {
        const int cur = t & 1, prv = cur ^ 1;
        mbar_wait_b(&fullC[stg], ring_phase);
        load_regs<NP>(ringC + ((size_t)stg * COLS + gidx) * D + lane_off, c);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&emptyC[stg]);   // this warp holds its C values in registers
        if (left_edge && t > 0) {
#pragma unroll
            for (int k = 0; k < NP; k++) pa[k] = (uint32_t)hq[k];
            ma = (uint32_t)hq[NP];
        } else {
            load_col_sm<NP>(Xa + ((size_t)prv * (COLS + 2) + gidx) * D, lg, pa);
            ma = Ma[prv * (COLS + 2) + gidx];
        }
        __syncwarp();
        sgm_step<NP, G>(pa, ma, c, P1x2, P2x2, PADx2, lg);
        if (right_edge) {   // the downstream strip's column 0 needs this for ITS row t + 1
            unsigned long long* hp = HD(s + 1, t);
#pragma unroll
            for (int k = 0; k < NP; k++) halo_put(hp + k, pa[k], seq0 | (uint32_t)(t + 1));
            halo_put(hp + NP, ma, seq0 | (uint32_t)(t + 1));
        }
        store_col_sm<NP>(Xa + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pa);
        if (lg == 0) Ma[cur * (COLS + 2) + gidx + 1] = ma;
        if (left_edge && t + 1 < g.H) halo_issue<NW>(HD(s, t), hq);   // consumed next row: the round trip overlaps the vertical path
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = pa[k];
        __syncwarp();
        sgm_step<NP, G>(Pv, mv, c, P1x2, P2x2, PADx2, lg);
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = __vminu2(__vadd2(acc[k], Pv[k]), 0x7FFF7FFFu);
        if (t >= WV_NSTA) mbar_wait_b(&accEmpty[ast], aphase ^ 1);   // the bulk store that read this slot's previous tile is done
        store_regs<NP>(accS + ((size_t)ast * COLS + gidx) * D + lane_off, acc);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&accFull[ast]);
        if (left_edge && t + 1 < g.H) halo_complete<NW>(HD(s, t), seq0 | (uint32_t)(t + 1), err, hq);
        named_bar_sync(1, WV_CONS);   // the row's diagonal states are in shared memory
        if (++stg == CF_NST) { stg = 0; ring_phase ^= 1; }
        if (++ast == WV_NSTA) { ast = 0; aphase ^= 1; }
    }
    (void)active;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_wave_bwd: CTA = strip of frame f, strips run right to left (block index 0 = the rightmost strip). Warps 0-7: the / path
// of the strip's columns (8 lanes per column) added into the staged partial S, then, two rows later, winner-take-all /
// uniqueness / sub-pixel on the tile the chain warp has finished. Warp 8: the <- chain. Warp 9: stages the C and S row tiles
// (two TMA bulk copies per row, WB_NST rows deep).
// ---------------------------------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(WV_CONS + 64, 2)
k_wave_bwd(Geom g, const int16_t* __restrict__ C, const int16_t* __restrict__ S, int b0, unsigned long long* haloD, unsigned long long* haloH,
           uint32_t seq0, int* err, int16_t* __restrict__ dispraw, uint32_t* __restrict__ wta) {
    constexpr int G = 8, NW = NP + 1, D = 16 * NP, COLS = WV_COLS;
    constexpr int NPc = NP / 4, NWc = NPc + 1;
    extern __shared__ __align__(128) uint8_t wbsm[];
    __shared__ __align__(8) uint64_t full[WB_NST], diagDone[WB_NST], chainDone[WB_NST], tileEmpty[WB_NST];
    int16_t* Xb = (int16_t*)wbsm;                                          // [2][COLS+2][D] state of the / path, slot = column + 1
    uint32_t* Mb = (uint32_t*)(wbsm + (size_t)2 * (COLS + 2) * D * 2);     // [2][COLS+2]
    const size_t off_tiles = ((size_t)2 * (COLS + 2) * D * 2 + 2 * (COLS + 2) * 4 + 127) & ~(size_t)127;
    int16_t* tiles = (int16_t*)(wbsm + off_tiles);                         // [WB_NST][C | S][COLS][D]
    const int nstrips = gridDim.x, s = nstrips - 1 - (int)blockIdx.x, f = blockIdx.y, b = b0 + f;
    const int x0 = s * COLS;
    for (int i = threadIdx.x; i < (int)(off_tiles / 4); i += blockDim.x) ((uint32_t*)wbsm)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < WB_NST; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&diagDone[i], WV_CONS / 32);
            mbar_init(&chainDone[i], 1);
            mbar_init(&tileEmpty[i], WV_CONS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ncols = min(COLS, g.W1 - x0);
    const uint32_t tile_bytes = (uint32_t)(ncols * D * 2);
    const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
    const int pad = 32767 - g.P1;
    const uint32_t PADx2 = pack16(pad, pad);
    const size_t tile_elems = (size_t)COLS * D;   // one array of a slot; a slot = C tile followed by S tile

    if (threadIdx.x >= WV_CONS + 32) {
        // =============================== loader ===============================
        if ((threadIdx.x & 31) != 0) return;
        const int16_t* Ct = C + ((size_t)b * g.crows * g.W1 + x0) * D;
        const int16_t* St = S + ((size_t)b * g.H * g.W1 + x0) * D;
        int st = 0;
        uint32_t ph = 0;
        for (int y = 0; y < g.H; y++) {
            if (y >= WB_NST) mbar_wait_b(&tileEmpty[st], ph ^ 1);
            mbar_expect_tx(&full[st], 2 * tile_bytes);
            bulk_g2s(tiles + (size_t)st * 2 * tile_elems, Ct + (size_t)y * g.W1 * D, tile_bytes, &full[st]);
            bulk_g2s(tiles + (size_t)st * 2 * tile_elems + tile_elems, St + (size_t)y * g.W1 * D, tile_bytes, &full[st]);
            if (++st == WB_NST) { st = 0; ph ^= 1; }
        }
        return;
    }
    if (threadIdx.x >= WV_CONS) {
        // =============================== chain warp: the <- path ===============================
        const int lane = threadIdx.x & 31;
        const size_t loff = (size_t)lane * 2 * NPc;
        auto HH = [&](int strip, int row) { return haloH + ((((size_t)f * nstrips + strip) * g.H + row) * 32 + lane) * NWc; };
        const bool has_up = s + 1 < nstrips;
        uint32_t P[NPc], c[NPc], a[NPc], cn[NPc], an[NPc], m = 0;
        unsigned long long hq[NWc];
        int st = 0;
        uint32_t ph = 0;
        if (has_up) halo_issue<NWc>(HH(s, 0), hq);
        for (int y = 0; y < g.H; y++) {
            if (has_up) {
                halo_complete<NWc>(HH(s, y), seq0 | (uint32_t)(y + 1), err, hq);
#pragma unroll
                for (int k = 0; k < NPc; k++) P[k] = (uint32_t)hq[k];
                m = (uint32_t)hq[NPc];
                if (y + 1 < g.H) halo_issue<NWc>(HH(s, y + 1), hq);
            } else {
#pragma unroll
                for (int k = 0; k < NPc; k++) P[k] = 0;
                m = 0;
            }
            mbar_wait_b(&diagDone[st], ph);   // the tile is staged and the / path has been added into it
            const int16_t* ct = tiles + (size_t)st * 2 * tile_elems + loff;
            int16_t* at = tiles + (size_t)st * 2 * tile_elems + tile_elems + loff;
            ld_chain<NPc>(ct + (size_t)(ncols - 1) * D, cn);
            ld_chain<NPc>(at + (size_t)(ncols - 1) * D, an);
            for (int x = ncols - 1; x >= 0; x--) {
#pragma unroll
                for (int k = 0; k < NPc; k++) { c[k] = cn[k]; a[k] = an[k]; }
                if (x > 0) {
                    ld_chain<NPc>(ct + (size_t)(x - 1) * D, cn);
                    ld_chain<NPc>(at + (size_t)(x - 1) * D, an);
                }
                sgm_step32<NPc>(P, m, c, P1x2, P2x2, PADx2, lane);
#pragma unroll
                for (int k = 0; k < NPc; k++) a[k] = __vminu2(__vadd2(a[k], P[k]), 0x7FFF7FFFu);
                st_chain<NPc>(at + (size_t)x * D, a);
            }
            if (s > 0) {
                unsigned long long* hp = HH(s - 1, y);
#pragma unroll
                for (int k = 0; k < NPc; k++) halo_put(hp + k, P[k], seq0 | (uint32_t)(y + 1));
                halo_put(hp + NPc, m, seq0 | (uint32_t)(y + 1));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&chainDone[st]);
            if (++st == WB_NST) { st = 0; ph ^= 1; }
        }
        return;
    }

    // =============================== / path + winner-take-all ===============================
    const int gidx = threadIdx.x / G, lg = threadIdx.x & (G - 1);
    const int x = x0 + gidx;
    const bool active = x < g.W1;
    const size_t lane_off = (size_t)lg * 2 * NP;
    auto HD = [&](int strip, int row) { return haloD + ((((size_t)f * nstrips + strip) * g.H + row) * G + lg) * NW; };
    // upstream = strip s + 1 (its column 0 is this strip's column COLS); downstream = strip s - 1
    const bool up_edge = gidx == COLS - 1 && s + 1 < nstrips, down_edge = gidx == 0 && s > 0;
    uint32_t pb[NP], c[NP], so[NP];
    uint32_t mb = 0;
    unsigned long long hq[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) hq[k] = 0;
    int st = 0, wst = 0;
    uint32_t ph = 0, wph = 0;
    for (int t = 0; t < g.H + WB_LAG; t++) {
        if (t < g.H) {
            const int cur = t & 1, prv = cur ^ 1;
            mbar_wait_b(&full[st], ph);
            int16_t* ctile = tiles + (size_t)st * 2 * tile_elems;
            load_regs<NP>(ctile + (size_t)gidx * D + lane_off, c);
            load_regs<NP>(ctile + tile_elems + (size_t)gidx * D + lane_off, so);
            if (up_edge && t > 0) {
#pragma unroll
                for (int k = 0; k < NP; k++) pb[k] = (uint32_t)hq[k];
                mb = (uint32_t)hq[NP];
            } else {
                load_col_sm<NP>(Xb + ((size_t)prv * (COLS + 2) + gidx + 2) * D, lg, pb);
                mb = Mb[prv * (COLS + 2) + gidx + 2];
            }
            __syncwarp();
            sgm_step<NP, G>(pb, mb, c, P1x2, P2x2, PADx2, lg);
            if (!active) {   // columns beyond W1 act as the zero border of the / path
#pragma unroll
                for (int k = 0; k < NP; k++) pb[k] = 0;
                mb = 0;
            }
            if (down_edge) {
                unsigned long long* hp = HD(s - 1, t);
#pragma unroll
                for (int k = 0; k < NP; k++) halo_put(hp + k, pb[k], seq0 | (uint32_t)(t + 1));
                halo_put(hp + NP, mb, seq0 | (uint32_t)(t + 1));
            }
            store_col_sm<NP>(Xb + ((size_t)cur * (COLS + 2) + gidx + 1) * D, lg, pb);
            if (lg == 0) Mb[cur * (COLS + 2) + gidx + 1] = mb;
            if (up_edge && t + 1 < g.H) halo_issue<NW>(HD(s, t), hq);
#pragma unroll
            for (int k = 0; k < NP; k++) so[k] = __vminu2(__vadd2(so[k], pb[k]), 0x7FFF7FFFu);
            store_regs<NP>(ctile + tile_elems + (size_t)gidx * D + lane_off, so);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&diagDone[st]);
            if (++st == WB_NST) { st = 0; ph ^= 1; }
        }
        if (t >= WB_LAG) {
            const int y = t - WB_LAG;
            mbar_wait_b(&chainDone[wst], wph);
            const int16_t* srow = tiles + (size_t)wst * 2 * tile_elems + tile_elems + (size_t)gidx * D;
            uint32_t Sr[NP];
            load_regs<NP>(srow + lane_off, Sr);
            int dv;
            uint32_t wv;
            wta_pixel<NP, G, false>(Sr, lg, g, const_cast<int16_t*>(srow), dv, wv);
            if (active && lg == 0) {
                dispraw[((size_t)b * g.H + y) * g.W + x + g.minX1] = (int16_t)dv;
                wta[((size_t)b * g.H + y) * g.W1 + x] = wv;
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&tileEmpty[wst]);
            if (++wst == WB_NST) { wst = 0; wph ^= 1; }
        }
        if (t < g.H) {
            if (up_edge && t + 1 < g.H) halo_complete<NW>(HD(s, t), seq0 | (uint32_t)(t + 1), err, hq);
            named_bar_sync(1, WV_CONS);   // the row's / states are in shared memory
        }
    }
}

struct WaveScratch {
    DevBuf haloD, haloH;
    uint32_t epoch = 0;
};

// Returns SDR_OK, an error, or -1 when the configuration is not covered (the caller takes the lock-step path).
template <int NP, int SW2>
int launch_wave(sdr_ctx* ctx, const Geom& g, int batch, const int4* PL, const uint8_t* PR, int16_t* C, int16_t* S, int16_t* dispraw, uint32_t* wta,
                V3Scratch& sc, WaveScratch& ws, StageTimer& timer, bool backward) {
    constexpr int D = 16 * NP, NPc = NP / 4;
    if (g.H > 4094) return -1;   // 12 bits of a halo word's sequence number hold row + 1
    const int nstrips = div_up(g.W1, WV_COLS);
    const WvLayout wl = wv_layout(D, SW2, g.SH2);
    const size_t smem_b = (((size_t)2 * (WV_COLS + 2) * D * 2 + 2 * (WV_COLS + 2) * 4 + 127) & ~(size_t)127) + (size_t)WB_NST * 2 * WV_COLS * D * 2;
    if (wl.total > ctx->smem_optin || smem_b > ctx->smem_optin) return -1;
    auto kf = k_wave_fwd<NP, SW2>;
    auto kb = k_wave_bwd<NP>;
    SDR_CUDA(ctx, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wl.total));
    SDR_CUDA(ctx, cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    int rc;
    const size_t rows = (size_t)batch * nstrips * g.H;
    const size_t bytesD = rows * 8 * (NP + 1) * 8, bytesH = rows * 32 * (NPc + 1) * 8;
    const bool grew = ws.haloD.cap < bytesD || ws.haloH.cap < bytesH;
    if ((rc = ws.haloD.reserve(ctx, bytesD)) || (rc = ws.haloH.reserve(ctx, bytesH))) return rc;
    if (grew || ws.epoch >= (1u << 19) - 2) {   // sequence numbers start over: every slot back to "never written"
        SDR_CUDA(ctx, cudaMemsetAsync(ws.haloD.p, 0, ws.haloD.cap, ctx->stream));
        SDR_CUDA(ctx, cudaMemsetAsync(ws.haloH.p, 0, ws.haloH.cap, ctx->stream));
        ws.epoch = 0;
    }
    if (!sc.err.p) {
        if ((rc = sc.err.reserve(ctx, sizeof(int)))) return rc;
        if ((rc = sc.host_err.reserve(ctx, sizeof(int)))) return rc;
        *(int*)sc.host_err.p = 0;
        SDR_CUDA(ctx, cudaMemsetAsync(sc.err.p, 0, sizeof(int), ctx->stream));
    }
    unsigned long long* hD = ws.haloD.as<unsigned long long>();
    unsigned long long* hH = ws.haloH.as<unsigned long long>();
    int* err = sc.err.as<int>();
    const uint32_t seq_f = (++ws.epoch) << 12;
    k_wave_fwd<NP, SW2><<<dim3(nstrips, batch), WV_PROD + WV_CONS + 32, wl.total, ctx->stream>>>(g, PL, PR, C, S, 0, hD, hH, seq_f, err);
    SDR_LAUNCHED(ctx);
    timer.mark(ctx->stream, "wave_fwd");
    if (backward) {
        const uint32_t seq_b = (++ws.epoch) << 12;
        k_wave_bwd<NP><<<dim3(nstrips, batch), WV_CONS + 64, smem_b, ctx->stream>>>(g, C, S, 0, hD, hH, seq_b, err, dispraw, wta);
        SDR_LAUNCHED(ctx);
        timer.mark(ctx->stream, "wave_bwd_wta");
    }
    SDR_CUDA(ctx, cudaMemcpyAsync(sc.host_err.p, sc.err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return SDR_OK;
}
