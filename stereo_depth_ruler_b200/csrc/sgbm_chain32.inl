// Row chains held by a whole warp (include file of sgbm.cu, inside its anonymous namespace): the path step with the
// cross-lane minimum by redux.sync, winner-take-all with redux.sync, and the latency-mode horizontal kernel built from them.
constexpr int WV_WAIT_LIMIT = 1 << 22;

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// mbarrier wait that cannot hang the GPU: a broken hand-shake inside the CTA ends the kernel with a trap (the launch fails
// loudly) instead of spinning until the watchdog
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
    for (int i = 0; i < WV_WAIT_LIMIT; i++)
        if (mbar_test(bar, parity)) return;
    __trap();
}

// One step of a horizontal path held by a whole warp: lane l owns disparities [2*NPc*l, 2*NPc*(l+1)).
template <int NPc>
__device__ __forceinline__ void sgm_step32(uint32_t (&P)[NPc], uint32_t& m, const uint32_t (&c)[NPc], uint32_t P1x2, uint32_t P2x2, uint32_t PADx2,
                                           int lane) {
    uint32_t up = __shfl_up_sync(0xffffffffu, P[NPc - 1], 1);
    uint32_t dn = __shfl_down_sync(0xffffffffu, P[0], 1);
    if (lane == 0) up = PADx2;
    if (lane == 31) dn = PADx2;
    uint32_t T[NPc + 1];
    T[0] = prmt(up, P[0], 0x5432);
#pragma unroll
    for (int k = 1; k < NPc; k++) T[k] = prmt(P[k - 1], P[k], 0x5432);
    T[NPc] = prmt(P[NPc - 1], dn, 0x5432);
    const uint32_t negm = __vadd2(~m, 0x00010001u);
    uint32_t mm = 0x7FFF7FFFu;
#pragma unroll
    for (int k = 0; k < NPc; k++) {
        uint32_t t = __vmins2(T[k], T[k + 1]);
        t = __viaddmin_s16x2(t, P1x2, P[k]);
        t = __viaddmin_s16x2(t, negm, P2x2);
        P[k] = __vadd2(c[k], t);
        mm = __vmins2(mm, P[k]);
    }
    mm = __vmins2(mm, prmt(mm, mm, 0x1032));   // both halves = the lane's minimum (>= 0), so the u32 order is the int16 order
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(mm));
}
template <int N>
__device__ __forceinline__ void ld_chain(const int16_t* p, uint32_t (&r)[N]) {
    if constexpr (N == 4) { const uint4 v = *(const uint4*)p; r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
    else if constexpr (N == 2) { const uint2 v = *(const uint2*)p; r[0] = v.x; r[1] = v.y; }
    else {   // N = 1, 3 (D = 64, 192): a lane's 4*N bytes are only 4-byte aligned
#pragma unroll
        for (int i = 0; i < N; i++) r[i] = ((const uint32_t*)p)[i];
    }
}
template <int N>
__device__ __forceinline__ void st_chain(int16_t* p, const uint32_t (&r)[N]) {
    if constexpr (N == 4) *(uint4*)p = make_uint4(r[0], r[1], r[2], r[3]);
    else if constexpr (N == 2) *(uint2*)p = make_uint2(r[0], r[1]);
    else {
#pragma unroll
        for (int i = 0; i < N; i++) ((uint32_t*)p)[i] = r[i];
    }
}

// Winner-take-all / uniqueness / sub-pixel for one pixel whose final S is spread over the 32 lanes of a warp (lane l: disparities
// [2*NPc*l, 2*NPc*(l+1))) and also sits in shared memory in natural order (`sm`). Same arithmetic as wta_pixel, the three
// cross-lane reductions are single redux.sync instructions.
template <int NPc>
__device__ __forceinline__ void wta_pixel32(const uint32_t (&Sr)[NPc], int lane, const Geom& g, const int16_t* sm, int& disp_out, uint32_t& wta_out) {
    uint32_t key = 0xFFFFFFFFu;
    const int d0 = lane * 2 * NPc;
    const uint32_t dw0 = (uint32_t)d0 | ((uint32_t)(d0 + 1) << 8);
#pragma unroll
    for (int k = 0; k < NPc; k++) {
        const uint32_t dw = dw0 + 0x0202u * k;
        key = __vimin3_u32(key, prmt(Sr[k], dw, 0x6104), prmt(Sr[k], dw, 0x6325));
    }
    key = __reduce_min_sync(0xffffffffu, key);
    const int minS = (int)(key >> 8);
    int best = (int)(key & 0xFFu);
    const int q = 100 - g.uq;
    const uint32_t num = g.three_way ? (uint32_t)(100 * minS) : (uint32_t)(minS * 100 + q - 1);
    const int quot = q == 1 ? (int)num : (int)__umulhi(num, g.uq_magic);
    int thr = g.three_way ? (g.uq > 0 ? quot + 1 : 0) : quot;
    thr = min(thr, 0x8000);
    const uint32_t cst = 0x80008000u - pack16(thr, thr);
    uint32_t nge2 = 0;
#pragma unroll
    for (int k = 0; k < NPc; k++) {
        const uint32_t t = Sr[k] + cst;
        nge2 = __vadd2(nge2, prmt_full(t, t, 0xBB99));
    }
    int nlt = 2 * NPc + (int)(int16_t)(nge2 & 0xFFFFu) + (int)(int16_t)(nge2 >> 16);
    nlt = __reduce_add_sync(0xffffffffu, nlt);
    if (g.tie8) {   // cv2's 8-lane search of computeDisparity3WaySGBM (see wta_pixel)
        int pos = 0x7FFF;
        if (lane < 8)
            for (int d = lane; d < g.D; d += 8)
                if (sm[d] == minS) pos = d;
        best = __reduce_min_sync(0xffffffffu, pos);
    }
    const int Sm = best > 0 ? sm[best - 1] : 0, Sp = best < g.D - 1 ? sm[best + 1] : 0;
    const int nexcl = (minS < thr) + (best > 0 && Sm < thr) + (best < g.D - 1 && Sp < thr);
    // A pixel whose costs are ALL saturated (minS == SHRT_MAX: large windows and penalties, eight paths of MODE_HH) keeps cv2's
    // initial bestDisp = -1 in computeDisparitySGBM (no cost is < MAX_COST), which comes out as the invalid value
    if (nlt > nexcl || (!g.three_way && minS >= 32767)) {
        disp_out = (g.minD - 1) * 16;
        wta_out = 0xFFFFFFFFu;
        return;
    }
    int d = best * 16;
    if (best > 0 && best < g.D - 1) {
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

// Both horizontal paths + winner-take-all for FEW rows in flight (one frame at a time -- the reference's usage -- or one stripe
// of a frame split over GPUs). There the time is the length of a row's dependent chain, W1 steps, so everything that is not
// the chain is moved off it. CTA = one row, six warps:
//   warps 0 / 1  walk the L->R / R->L chain (32 lanes, minimum by redux.sync: ~70 clocks per step against ~190 with shuffles).
//                As in k_path_hh the chains meet in the middle: on its first half a chain adds its L into S in HBM, on its second
//                half the other chain has been there and S + L is final. C and S come through a shared-memory ring filled by
//                TMA bulk copies (chunks of 1 KB per array, three chunks deep): a register prefetch cannot run that far ahead --
//                a warp has six scoreboards, so sixteen loads in flight share them and a wait for the oldest waits for all
//                (measured: 8, 16 and 24 steps of register prefetch all gave the same 600 ns per step).
//                The two chains synchronise ONCE, between the halves (their first-half stores must be in L2 before the other
//                chain's bulk copies read them); a barrier per step costs more than the step.
//   warps 2 - 5  two per chain: winner-take-all, uniqueness and sub-pixel fit of the finished pixels, taken from a small ring of
//                final S vectors a few pixels behind the chain, alternating pixels; no dependence between pixels.
constexpr int HWS_TSLOTS = 8;     // final-S vectors in flight per chain
constexpr int HWS_NST = 3;        // staged (C, S) chunks in flight per chain
constexpr int HWS_THREADS = 192;

template <int NPc>
__global__ void __launch_bounds__(HWS_THREADS) k_path_hh_ws(Geom g, const int16_t* __restrict__ C, int16_t* S, int16_t* __restrict__ dispraw,
                                                            uint32_t* __restrict__ wta) {
    constexpr int D = 64 * NPc, K = 1024 / D;   // steps per chunk: 2 KB of C, 2 KB of S
    __shared__ __align__(128) int16_t stage[2][HWS_NST][2][K * D];
    __shared__ __align__(16) int16_t ring[2][HWS_TSLOTS][D];
    __shared__ __align__(8) uint64_t fullIn[2][HWS_NST], fullT[2][HWS_TSLOTS], emptyT[2][HWS_TSLOTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = g.hrow0 + blockIdx.x, b = blockIdx.z;
    const int W1 = g.W1;
    int y = row, si = -1;
    bool active = true;
    for (int i = 0; i < g.nseg; i++)
        if (y >= g.seg[i].dst_start && y < g.seg[i].src_end) si = i;
    if (si < 0) {   // row outside the selected stripes: run on a valid row, write nothing
        si = 0;
        y = g.seg[0].dst_start;
        active = false;
    }
    const Seg sg = g.seg[si];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * HWS_NST; i++) mbar_init(&fullIn[0][0] + i, 1);
        for (int i = 0; i < 2 * HWS_TSLOTS; i++) { mbar_init(&fullT[0][0] + i, 1); mbar_init(&emptyT[0][0] + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int h1 = (W1 + 1) / 2;   // L->R accumulates x < h1, R->L accumulates x >= h1; each finishes the other's pixels
    if (warp < 2) {
        // ---- the two chains ----
        const int ch = warp, dir = ch ? -1 : +1;
        const uint32_t P1x2 = pack16(g.P1, g.P1), P2x2 = pack16(g.P2, g.P2);
        const int pad = 32767 - g.P1;
        const uint32_t PADx2 = pack16(pad, pad);
        const size_t loff = (size_t)lane * 2 * NPc;
        const int16_t* Crow = C + (((size_t)b * g.crows + sg.crow0 + (y - sg.src_start)) * W1) * D;
        int16_t* Srow = S + (((size_t)b * g.H + y) * W1) * D;
        uint32_t P[NPc], c[NPc], s[NPc], L[NPc], m = 0;
#pragma unroll
        for (int k = 0; k < NPc; k++) P[k] = 0;
        int cc = 0;          // chunks consumed so far (both halves): stage = cc % NST, parity = (cc / NST) & 1
        int slot = 0;
        uint32_t tphase = 0;
        for (int half = 0; half < 2; half++) {
            // the pixels of this half in walking order: n pixels from xs in direction dir
            const int n = (dir > 0) == (half == 0) ? h1 : W1 - h1;
            const int xs = dir > 0 ? (half == 0 ? 0 : h1) : (half == 0 ? W1 - 1 : h1 - 1);
            const int nchunks = (n + K - 1) / K;
            auto issue_at = [&](int ci, int gidx) {   // gidx = running chunk index of chunk ci
                const int cnt = min(K, n - ci * K);
                const int xlo = dir > 0 ? xs + ci * K : xs - ci * K - cnt + 1;
                const int st = gidx % HWS_NST;
                const uint32_t bytes = (uint32_t)cnt * D * 2;
                mbar_expect_tx(&fullIn[ch][st], 2 * bytes);
                bulk_g2s(&stage[ch][st][0][0], Crow + (size_t)xlo * D, bytes, &fullIn[ch][st]);
                bulk_g2s(&stage[ch][st][1][0], Srow + (size_t)xlo * D, bytes, &fullIn[ch][st]);
            };
            if (lane == 0)
                for (int ci = 0; ci < HWS_NST && ci < nchunks; ci++) issue_at(ci, cc + ci);
            for (int ci = 0; ci < nchunks; ci++, cc++) {
                const int st = cc % HWS_NST;
                const int cnt = min(K, n - ci * K);
                mbar_wait_b(&fullIn[ch][st], (uint32_t)((cc / HWS_NST) & 1));
                const int16_t* cs = &stage[ch][st][0][0] + loff;
                const int16_t* ss = &stage[ch][st][1][0] + loff;
                for (int j = 0; j < cnt; j++) {
                    const int x = xs + dir * (ci * K + j);
                    const int pos = dir > 0 ? j : cnt - 1 - j;   // the chunk is staged in ascending x
                    ld_chain<NPc>(cs + (size_t)pos * D, c);
                    ld_chain<NPc>(ss + (size_t)pos * D, s);
                    sgm_step32<NPc>(P, m, c, P1x2, P2x2, PADx2, lane);
#pragma unroll
                    for (int k = 0; k < NPc; k++) L[k] = __vminu2(__vadd2(s[k], P[k]), 0x7FFF7FFFu);
                    if (half == 0) {
                        if (active) st_chain<NPc>(Srow + (size_t)x * D + loff, L);
                    } else {
                        mbar_wait_b(&emptyT[ch][slot], tphase ^ 1);   // first use of a slot passes at once
                        st_chain<NPc>(&ring[ch][slot][0] + loff, L);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&fullT[ch][slot]);
                        if (++slot == HWS_TSLOTS) { slot = 0; tphase ^= 1; }
                    }
                }
                __syncwarp();   // every lane is done reading this stage
                if (lane == 0 && ci + HWS_NST < nchunks) issue_at(ci + HWS_NST, cc + HWS_NST);
            }
            if (half == 0) {
                // the other chain's bulk copies (async proxy) read what this chain stored: stores to L2, proxy fence, then meet
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                named_bar_sync(1, 64);
            }
        }
        return;
    }
    // ---- the tails: the finished pixels of chain ch (its second half) in the order the chain reaches them, alternating ----
    const int ch = (warp - 2) & 1, tj = (warp - 2) >> 1, dir = ch ? -1 : +1;
    const int n = dir > 0 ? W1 - h1 : h1, xa = dir > 0 ? h1 : h1 - 1;
    for (int i = tj; i < n; i += 2) {
        const int x = xa + dir * i, slot = i % HWS_TSLOTS;
        mbar_wait_b(&fullT[ch][slot], (uint32_t)((i / HWS_TSLOTS) & 1));
        const int16_t* srow = &ring[ch][slot][0];
        uint32_t Sr[NPc];
        ld_chain<NPc>(srow + (size_t)lane * 2 * NPc, Sr);
        int dv;
        uint32_t wv;
        wta_pixel32<NPc>(Sr, lane, g, srow, dv, wv);
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&emptyT[ch][slot]);
            if (active) {
                dispraw[((size_t)b * g.H + y) * g.W + x + g.minX1] = (int16_t)dv;
                wta[((size_t)b * g.H + y) * W1 + x] = wv;
            }
        }
    }
}
