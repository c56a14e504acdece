// FeatureNet.downsample1/2 (3x3 stride-2 pad-1 conv, models/module.py:214,218) as a persistent row-streaming kernel:
// ring of raw input rows (bulk copies) -> staging warps (the producer's InstanceNorm + LeakyReLU, once per input pixel) -> tcgen05 tap GEMMs
// -> TMEM -> epilogue (fp16 value + residual planes, InstanceNorm statistics).
//
// csrc/conv2d_gtc.cu computes the same layer by a per-output-pixel gather: 9 (x2 planes) 16-byte loads and 9 normalisations
// per output pixel and chunk, a serial gather -> MMA -> epilogue chain per tile; it ran at ~1.4 TB/s of the 6.5 TB/s the two
// layers need (0.32 + 0.23 ms per cfg2 map against 0.07 + 0.035 ms of traffic).  Here every input row is loaded ONCE and
// normalised ONCE:
//   * the producer thread fetches the row segment a strip needs (255 pixels, both planes) with one 1-D bulk copy per plane
//     (cp.async.bulk) into a ring of RAW rows.  (A tensor map that de-interleaves the pixel phases -- boxes of 16-byte inner
//     extent -- was measured first: the TMA unit spends ~4 cycles per 16-byte piece, which made it the kernel's bound.)
//   * staging warps read a raw row once, apply the producer's InstanceNorm + LeakyReLU and write the fp16 value and its fp16
//     rounding residual into the OPERAND ring as dense slabs of even / odd pixels, 16 bytes apart -- the tcgen05 K-major
//     no-swizzle operand layout (tc_common.cuh); zeros outside the image: the conv pads the ACTIVATED tensor.  Slab index i
//     holds phase pixel x0-1+i, so that for output pixel x0 + r the taps are
//       dx=0 (input column 2x-1): odd slab, index r      dx=1 (column 2x): even slab, index r+1      dx=2: odd slab, index r+1;
//   * input row R feeds output row R/2 (kernel row 1) if even, rows (R-1)/2 (kernel row 2) and (R+1)/2 (kernel row 0) if odd:
//     the issuer warp issues that row's MMAs (M = 128 output pixels, K = 16 = two slabs, N = 2*Cout: weights | their fp16
//     rounding residuals) into the TMEM slot of the output row (slot = y - y0, TY rows per tile);
//       Cin  8: per kernel row (odd value @r, odd value @r+1), (even value, even residual), (odd residual @r, @r+1)
//       Cin 16: per kernel row and tap (chunk 0, chunk 1) of the value plane, then of the residual plane
//   * the epilogue warps drain a slot when its last input row has been accumulated.
// Weights (host: weights.py pack_conv2d_s2rows): [kernel row 3][image NIMG][k-chunk 2][N/8][8 n][8 k] fp16.
// Requires W even and input statistics; the caller falls back to conv2d_gtc.cu otherwise.
#include <algorithm>
#include <cstdlib>

#include "cds_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TXO = 127;             // output pixels per strip: 255 input pixels = one piece per staging thread (Cin 8)
constexpr int NPX = 2 * TXO + 1;     // raw pixels 1 .. 255 of the segment starting at column 2*(x0-1): phase index (px >> 1)
constexpr int TXB = 132;             // slab pitch in pixels: 2112 B = 64 (mod 128), so the two phases' writes of a quarter warp miss each other's banks
constexpr int SLABB = TXB * 16;
constexpr int NSTGW = 8;             // staging warps (in NGRP groups, each owning every NGRP-th row)
constexpr int NTHREADS = 448;        // warps: 0 producer, 1 MMA, 2-9 staging, 10-13 epilogue (one per TMEM lane quadrant: warp & 3)
constexpr int kStgWarp0 = 2, kEpiWarp0 = kStgWarp0 + NSTGW;
constexpr float kInEps = 1e-5f;

template <int CIN_, int COUT_, int TY_, int NR_, int NRAW_, int NGRP_, int MINB_>
struct S2 {
    static constexpr int CIN = CIN_, COUT = COUT_, C8 = CIN_ / 8, TY = TY_, NR = NR_, NRAW = NRAW_, NGRP = NGRP_, MINB = MINB_;
    static constexpr int RAWP = NPX * CIN_ * 2;                 // one plane of a raw row segment
    static constexpr int RAWB = 2 * RAWP;
    // A ring slot must always be handled by the SAME staging group (ring depths = multiples of NGRP): a parity wait can only
    // tell a phase from its neighbour, so whoever waits on a slot's barriers has to see every one of its phases in turn -- a
    // group arriving at a slot another group is still one phase behind on would sail through the wait.
    static_assert(NSTGW % NGRP_ == 0 && NR_ % NGRP_ == 0 && NRAW_ % NGRP_ == 0 && NR_ > NGRP_, "ring depths must be multiples of the group count");
    static constexpr int WPG = NSTGW / NGRP_;                                       // warps per staging group
    static constexpr int ROUNDS = (NPX * C8 + WPG * 32 - 1) / (WPG * 32);           // 16-byte pieces per staging thread and row
    static constexpr int NC = 2 * COUT_;                       // accumulator columns of an output row: W_hi | W_lo products
    static constexpr int PLANE = 2 * C8 * SLABB;               // value plane, then residual plane: [phase 2][chunk C8] slabs each
    static constexpr int ROWB = 2 * PLANE;
    static constexpr int NIMG = C8 == 1 ? 2 : 3;               // weight images per kernel row
    static constexpr int IMGB = 2 * NC * 16;
    static constexpr int B_BYTES = 3 * NIMG * IMGB;
    static constexpr int ACC_COLS = TY * NC;
    static constexpr int TMEM_COLS = ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512);
    static_assert(ACC_COLS <= 512, "accumulator tile exceeds TMEM");
    static constexpr int NBAR = 2 * NRAW + 2 * NR + 2 * TY + 1;
    static constexpr size_t SMEM = (size_t)NR * ROWB + (size_t)NRAW * RAWB + B_BYTES + NBAR * 8 + 16 + NGRP * 2 * CIN * 4;
};

struct S2Params {
    const __half* in;         // [n][H][W][CIN] raw
    const __half* in_lo;      // residual plane of in, or NULL
    const double* in_stats;   // [n][CIN][2]
    const __half* wgt;
    __half* out;              // [n][Ho][Wo][COUT] raw
    __half* out_lo;           // optional residual plane of out
    double* out_stats;        // [n][COUT][2] or NULL
    int in_act, has_lo;
    int H, W, Ho, Wo, n;
    int xt, yt;
    int dbg;                  // CDS_S2_DEBUG (diagnostics): 1 staging copies without arithmetic, 2 epilogue drains without stores, 4 no MMAs
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    const uint32_t a = tc::smem_u32(bar);
    for (uint32_t spin = 0; !tc::mbar_try_wait(a, parity); ++spin) {
        __nanosleep(32);
        if (spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory"); }
struct Tile {
    int n, y0, ylast, x0, rlo, rhi;
};
template <class C>
__device__ __forceinline__ Tile tile_of(const S2Params& p, int t) {
    Tile q;
    const int tx = t % p.xt, ty = (t / p.xt) % p.yt;
    q.n = t / (p.xt * p.yt);
    q.x0 = tx * TXO;
    q.y0 = ty * C::TY;
    q.ylast = min(q.y0 + C::TY - 1, p.Ho - 1);
    q.rlo = max(2 * q.y0 - 1, 0);
    q.rhi = min(2 * q.ylast + 1, p.H - 1);
    return q;
}

// the MMAs of input row (ring slot at a_row) for kernel row dy into the accumulator at d_col
template <class C>
__device__ __forceinline__ void issue_dy(uint32_t a_row, uint32_t sB_u, int dy, uint32_t d_col, bool first, bool elected) {
    constexpr uint32_t idesc = tc::instr_desc_f16(128, C::NC);
    constexpr uint32_t b_lbo = C::NC * 16;
    const uint32_t b0 = sB_u + (uint32_t)dy * C::NIMG * C::IMGB;
    // slab(plane, phase, chunk)
    auto slab = [&](int pl, int ph, int c) { return a_row + (uint32_t)(pl * C::PLANE + (ph * C::C8 + c) * SLABB); };
    if constexpr (C::C8 == 1) {
        const uint64_t db0 = tc::smem_desc(b0, b_lbo, 128), db1 = tc::smem_desc(b0 + C::IMGB, b_lbo, 128);
        const uint64_t a0 = tc::smem_desc(slab(0, 1, 0), 16, 128);                       // odd value @r | @r+1
        const uint64_t a1 = tc::smem_desc(slab(0, 0, 0) + 16, C::PLANE, 128);            // even value @r+1 | even residual @r+1
        const uint64_t a2 = tc::smem_desc(slab(1, 1, 0), 16, 128);                       // odd residual @r | @r+1
        if (elected) {
            tc::mma_f16(d_col, a0, db0, idesc, !first);
            tc::mma_f16(d_col, a1, db1, idesc, true);
            tc::mma_f16(d_col, a2, db0, idesc, true);
        }
    } else {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ph = dx == 1 ? 0 : 1, off = dx == 0 ? 0 : 16;
                const uint64_t da = tc::smem_desc(slab(pl, ph, 0) + off, SLABB, 128);    // chunk 0 | chunk 1 of the tap
                const uint64_t db = tc::smem_desc(b0 + (uint32_t)dx * C::IMGB, b_lbo, 128);
                if (elected) tc::mma_f16(d_col, da, db, idesc, !(first && pl == 0 && dx == 0));
            }
    }
}

template <class C>
__global__ void __launch_bounds__(NTHREADS, C::MINB) conv2d_s2rows_kernel(const S2Params p) {
    constexpr int CIN = C::CIN, COUT = C::COUT, C8 = C::C8, TY = C::TY, NR = C::NR, NRAW = C::NRAW, NC = C::NC;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                   // operand ring
    uint8_t* sR = smem + (size_t)NR * C::ROWB;            // raw ring
    uint8_t* sB = sR + (size_t)NRAW * C::RAWB;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(sB + C::B_BYTES);   // [NRAW] raw row landed (bulk copies)
    uint64_t* raw_empty = raw_full + NRAW;                               // [NRAW] raw row consumed (staging warps)
    uint64_t* ring_ready = raw_empty + NRAW;                             // [NR] operand row written (staging warps)
    uint64_t* ring_empty = ring_ready + NR;                              // [NR] operand row consumed (MMA commits)
    uint64_t* acc_full = ring_empty + NR;
    uint64_t* acc_empty = acc_full + TY;
    uint64_t* bar_b = acc_empty + TY;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 1);
    float* s_norm = reinterpret_cast<float*>(bar_b + 2);   // [CIN][2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB), sR_u = tc::smem_u32(sR);
    const int ntiles = p.xt * p.yt * p.n;
    // a contiguous run of tiles per CTA: one or two image changes (norm coefficients, statistics flush) instead of one every few tiles
    const int t_begin = (int)((long long)ntiles * blockIdx.x / gridDim.x), t_end = (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    if (threadIdx.x == 32) {
        for (int i = 0; i < NRAW; ++i) { tc::mbar_init(raw_full + i, 1); tc::mbar_init(raw_empty + i, NSTGW / C::NGRP); }
        for (int i = 0; i < NR; ++i) { tc::mbar_init(ring_ready + i, NSTGW / C::NGRP); tc::mbar_init(ring_empty + i, 1); }
        for (int i = 0; i < TY; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 4); }
        tc::mbar_init(bar_b, 1);
        tc::mbar_fence_init();
    }
    // the slabs' tail (indices 129 .. TXB-1) is never staged but sits inside the MMAs' 8-row fetch granularity: finite
    for (int i = threadIdx.x; i < NR * C::ROWB / 16; i += NTHREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- producer: weights once, then one bulk copy per plane and input row ------------------------------------------------
        if (tc::elect_one()) {
            tc::mbar_expect_tx(bar_b, C::B_BYTES);
            tc::bulk_copy_g2s(sB_u, p.wgt, C::B_BYTES, bar_b);
            uint32_t rc = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const Tile q = tile_of<C>(p, t);
                // columns gx1 .. gx1 + 254 of the row, clipped to the image (the staging warps zero what lies outside)
                const int gx1 = 2 * (q.x0 - 1) + 1, gxs = max(gx1, 0), gxe = min(gx1 + NPX, p.W);
                const uint32_t bytes = (uint32_t)(gxe - gxs) * CIN * 2, doff = (uint32_t)(gxs - gx1) * CIN * 2;
                for (int R = q.rlo; R <= q.rhi; ++R, ++rc) {
                    const uint32_t rs = rc % NRAW;
                    if (rc >= (uint32_t)NRAW) mbar_wait_relaxed(raw_empty + rs, ((rc / NRAW) - 1) & 1);
                    tc::mbar_expect_tx(raw_full + rs, p.has_lo ? 2 * bytes : bytes);
                    const size_t src = (((size_t)q.n * p.H + R) * p.W + gxs) * CIN;
                    tc::bulk_copy_g2s(sR_u + rs * C::RAWB + doff, p.in + src, bytes, raw_full + rs);
                    if (p.has_lo) tc::bulk_copy_g2s(sR_u + rs * C::RAWB + C::RAWP + doff, p.in_lo + src, bytes, raw_full + rs);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ------------------------------------------------------------------------------------------------------
        tc::mbar_wait(bar_b, 0);
        tc::tc_fence_after();
        const bool elected = tc::elect_one();
        const uint32_t tmem_u = tc::uniform(tmem);
        uint32_t rc = 0, emp_par = 0;
#pragma unroll 1
        for (int t = t_begin; t < t_end; ++t) {
            const Tile q = tile_of<C>(p, t);
#pragma unroll 1
            for (int R = q.rlo; R <= q.rhi; ++R, ++rc) {
                const uint32_t slot = rc % NR;
                // (output row, kernel row) pairs this input row feeds, oldest output row first
                const int ya = (R & 1) ? (R - 1) / 2 : R / 2, dya = (R & 1) ? 2 : 1;
                const int yb = (R + 1) / 2;   // odd rows only: kernel row 0, the FIRST contribution of output row yb
                const bool has_a = ya >= q.y0 && ya <= q.ylast, has_b = (R & 1) && yb >= q.y0 && yb <= q.ylast;
                // first contribution of an output row: input row 2y-1 (kernel row 0), or row 0 (kernel row 1) for y = 0
                const bool first_a = R == 0;
                if (has_a && first_a) { tc::mbar_wait(acc_empty + (ya - q.y0), ((emp_par >> (ya - q.y0)) & 1) ^ 1); emp_par ^= 1u << (ya - q.y0); }
                if (has_b) { tc::mbar_wait(acc_empty + (yb - q.y0), ((emp_par >> (yb - q.y0)) & 1) ^ 1); emp_par ^= 1u << (yb - q.y0); }
                tc::mbar_wait(ring_ready + slot, (rc / NR) & 1);
                tc::tc_fence_after();
                const uint32_t a_row = sA_u + slot * C::ROWB;
                if (has_a && !(p.dbg & 4)) issue_dy<C>(a_row, sB_u, dya, tmem_u + (uint32_t)(ya - q.y0) * NC, first_a, elected);
                if (has_b && !(p.dbg & 4)) issue_dy<C>(a_row, sB_u, 0, tmem_u + (uint32_t)(yb - q.y0) * NC, true, elected);
                if (p.dbg & 64) {   // timing experiment (no MMAs in flight): plain arrivals instead of commits
                    if (elected) mbar_arrive(ring_empty + slot);
                    if (has_a && ((R & 1) || R == p.H - 1) && elected) mbar_arrive(acc_full + (ya - q.y0));
                } else {
                if (elected) tc::mma_commit(ring_empty + slot);
                // the output row whose last input row this was: 2y+1, or 2y when that is the image's last row
                if (has_a && ((R & 1) || R == p.H - 1) && elected) tc::mma_commit(acc_full + (ya - q.y0));
                }
                __syncwarp();
            }
        }
    } else if (warp >= kStgWarp0 && warp < kEpiWarp0) {
        // ---- staging: InstanceNorm + activation of the producer, value + residual planes ----------------------------------------
        // The warps form NGRP groups; group g owns the rows rc = g (mod NGRP), so NGRP rows are in flight at once (one group over
        // all rows ran them back to back: wait, one round of loads, arithmetic, stores, proxy fence, arrive = ~350 cycles per row of
        // pure latency whatever the thread count) and a thread's pieces of a row are independent work to interleave.
        constexpr int NGRP = C::NGRP, WPG = NSTGW / NGRP;
        const int sw = warp - kStgWarp0, grp = sw / WPG, gtid = (sw % WPG) * 32 + lane;
        float* nm_grp = s_norm + grp * 2 * CIN;
        auto group_sync = [&]() {
            if constexpr (WPG == 1) __syncwarp();
            else named_barrier(1 + grp, WPG * 32);
        };
        uint32_t rc = 0;
        int cur_n = -1;
#pragma unroll 1
        for (int t = t_begin; t < t_end; ++t) {
            const Tile q = tile_of<C>(p, t);
            const int gx1 = 2 * (q.x0 - 1) + 1;
#pragma unroll 1
            for (int R = q.rlo; R <= q.rhi; ++R, ++rc) {
                if ((int)(rc % NGRP) != grp) continue;
                if (q.n != cur_n) {   // group-uniform
                    group_sync();
                    if (gtid < CIN) {
                        const double cnt = (double)p.H * p.W;
                        const double s = p.in_stats[((size_t)q.n * CIN + gtid) * 2], ss = p.in_stats[((size_t)q.n * CIN + gtid) * 2 + 1];
                        const double m = s / cnt;
                        double var = ss / cnt - m * m;
                        if (var < 0.0) var = 0.0;
                        nm_grp[2 * gtid] = (float)m;
                        nm_grp[2 * gtid + 1] = (float)(1.0 / sqrt(var + (double)kInEps));
                    }
                    group_sync();
                    cur_n = q.n;
                }
                const uint32_t slot = rc % NR, rs = rc % NRAW;
                tc::mbar_wait(raw_full + rs, (rc / NRAW) & 1);
                // this thread's pieces of the row: piece u = the u-th 16 bytes of the raw segment (raw pixel 1 + u / C8, chunk
                // u % C8); consecutive lanes read consecutive pieces and write the two phases' slabs alternately
                const uint8_t* rawrow = sR + (size_t)rs * C::RAWB;
                uint4 raw[C::ROUNDS], rawl[C::ROUNDS];
#pragma unroll
                for (int k = 0; k < C::ROUNDS; ++k) {
                    const int u = gtid + k * WPG * 32, gx = gx1 + u / C8;
                    const bool ok = u < NPX * C8 && gx >= 0 && gx < p.W && !(p.dbg & 8);
                    raw[k] = ok ? *reinterpret_cast<const uint4*>(rawrow + (size_t)u * 16) : make_uint4(0, 0, 0, 0);
                    rawl[k] = (ok && p.has_lo) ? *reinterpret_cast<const uint4*>(rawrow + C::RAWP + (size_t)u * 16) : make_uint4(0, 0, 0, 0);
                }
                if (rc >= (uint32_t)NR) tc::mbar_wait(ring_empty + slot, ((rc / NR) - 1) & 1);   // the MMAs of the slot's previous row are done
                tc::tc_fence_after();
                uint8_t* row = sA + (size_t)slot * C::ROWB;
#pragma unroll
                for (int k = 0; k < C::ROUNDS; ++k) {
                    const int u = gtid + k * WPG * 32;
                    if (u >= NPX * C8 || (p.dbg & 8)) break;
                    const int px = 1 + u / C8, c = u % C8;
                    const int ph = px & 1, i = px >> 1;
                    const int gx = gx1 + u / C8;
                    if (gx >= 0 && gx < p.W && !(p.dbg & 1)) {
                        __half2* h = reinterpret_cast<__half2*>(&raw[k]);
                        __half2* hl = reinterpret_cast<__half2*>(&rawl[k]);
                        const float* nm = nm_grp + c * 16;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float2 f = __half22float2(h[j]);
                            const float2 g = __half22float2(hl[j]);
                            f.x = ((f.x + g.x) - nm[4 * j]) * nm[4 * j + 1];
                            f.y = ((f.y + g.y) - nm[4 * j + 2]) * nm[4 * j + 3];
                            if (p.in_act == 1) { f.x = f.x > 0.f ? f.x : 0.1f * f.x; f.y = f.y > 0.f ? f.y : 0.1f * f.y; }
                            h[j] = __floats2half2_rn(f.x, f.y);
                            const float2 back = __half22float2(h[j]);
                            hl[j] = __floats2half2_rn(f.x - back.x, f.y - back.y);
                        }
                    }
                    const size_t o = (size_t)(ph * C8 + c) * SLABB + (size_t)i * 16;
                    *reinterpret_cast<uint4*>(row + o) = raw[k];     // outside the image: the conv's zero padding of the ACTIVATED tensor
                    *reinterpret_cast<uint4*>(row + C::PLANE + o) = rawl[k];
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(ring_ready + slot); mbar_arrive(raw_empty + rs); }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ---- epilogue: one output pixel per thread ------------------------------------------------------------------------------
        const int lg = warp & 3;
        const int r = lg * 32 + lane;
        float st_sum[COUT], st_sq[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) { st_sum[c] = 0.f; st_sq[c] = 0.f; }
        int cur_n = -1;
        auto flush = [&]() {
            if (cur_n < 0 || !p.out_stats) return;
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                const float a = warp_sum(st_sum[c]), qq = warp_sum(st_sq[c]);
                if (lane == 0) {
                    atomicAdd(p.out_stats + ((size_t)cur_n * COUT + c) * 2, (double)a);
                    atomicAdd(p.out_stats + ((size_t)cur_n * COUT + c) * 2 + 1, (double)qq);
                }
                st_sum[c] = 0.f;
                st_sq[c] = 0.f;
            }
        };
        uint32_t full_par = 0;
#pragma unroll 1
        for (int t = t_begin; t < t_end; ++t) {
            const Tile q = tile_of<C>(p, t);
            if (q.n != cur_n) { flush(); cur_n = q.n; }
            const int gx = q.x0 + r;
            const bool live = r < TXO && gx < p.Wo;
#pragma unroll 1
            for (int s = 0; s <= q.ylast - q.y0; ++s) {
                tc::mbar_wait(acc_full + s, (full_par >> s) & 1);
                full_par ^= 1u << s;
                tc::tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)s * NC;
                const size_t m = ((size_t)q.n * p.Ho + (q.y0 + s)) * p.Wo + gx;
#pragma unroll
                for (int c8 = 0; c8 < COUT / 8; ++c8) {
                    uint32_t hi[8], lo[8];
                    if (!(p.dbg & 16)) {
                    tc::tmem_ld8_nowait(taddr + c8 * 8, hi);
                    tc::tmem_ld8_nowait(taddr + COUT + c8 * 8, lo);
                    tc::tmem_ld_wait();
                    }
                    if (c8 == COUT / 8 - 1) {   // every column of the slot is in registers: hand it back
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + s);
                    }
                    if (live && !(p.dbg & 2)) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v[i] = __uint_as_float(hi[i]) + __uint_as_float(lo[i]);
                            st_sum[c8 * 8 + i] += v[i];
                            st_sq[c8 * 8 + i] += v[i] * v[i];
                        }
                        Vec8<__half>::store(p.out + m * COUT + c8 * 8, v);
                        if (p.out_lo) {
                            float res[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) res[i] = v[i] - __half2float(__float2half_rn(v[i]));
                            Vec8<__half>::store(p.out_lo + m * COUT + c8 * 8, res);
                        }
                    }
                }
            }
        }
        flush();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, C::TMEM_COLS);
}

int s2_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <class C>
int launch_s2(S2Params p, cudaStream_t st) {
    static_assert(C::SMEM <= 227 * 1024, "ring + weights do not fit in shared memory");
    auto kern = conv2d_s2rows_kernel<C>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) { cds_set_error("cds_conv2d_3x3s2_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    p.xt = cds_div_up(p.Wo, TXO);
    p.yt = cds_div_up(p.Ho, C::TY);
    const long long ntiles = (long long)p.xt * p.yt * p.n;
    // resident CTAs per SM: registers / shared memory (asked of the runtime) and the TMEM columns each CTA allocates
    static int occ = 0;
    if (!occ && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, C::SMEM) != cudaSuccess || occ < 1)) occ = 1;
    const int per_sm = std::max(1, std::min(occ, 512 / C::TMEM_COLS));
    const int grid = (int)std::min<long long>(ntiles, (long long)s2_sm_count() * per_sm);
    kern<<<grid, NTHREADS, C::SMEM, st>>>(p);
    return cds_check_launch("cds_conv2d_3x3s2_rows");
}

}  // namespace

extern "C" {

int cds_conv2d_3x3s2_rows_supported(int Cin, int Cout, int H, int W) {
    return ((Cin == 8 && Cout == 16) || (Cin == 16 && Cout == 32)) && W % 2 == 0 && W >= 2 && H >= 1;
}
int cds_conv2d_3x3s2_rows_weight_halfs(int Cin, int Cout) { return 3 * (Cin == 8 ? 2 : 3) * 2 * (2 * Cout) * 8; }

// in [n,H,W,Cin] fp16 raw + its statistics / activation (+ in_lo, the fp16 rounding residual plane of in, or NULL)
// -> out [n,ceil(H/2),W/2,Cout] fp16 raw (+ out_lo, its residual plane) + out_stats.  Same contract as cds_conv2d_3x3s2_tc.
int cds_conv2d_3x3s2_rows(const void* in, const void* in_lo, const double* in_stats, int in_act, const void* wgt_packed, int n, int Cin,
                          int Cout, int H, int W, void* out, void* out_lo, double* out_stats, cudaStream_t stream) {
    CDS_REQUIRE(in && in_stats && wgt_packed && out, CDS_EARG, "cds_conv2d_3x3s2_rows: null pointer");
    CDS_REQUIRE(n > 0 && H > 0 && W > 0 && (long long)n * H < (1ll << 31), CDS_ESHAPE, "cds_conv2d_3x3s2_rows: bad shape");
    CDS_REQUIRE(cds_conv2d_3x3s2_rows_supported(Cin, Cout, H, W), CDS_EUNSUPPORTED,
                "cds_conv2d_3x3s2_rows: unsupported layer Cin=%d Cout=%d W=%d (W must be even)", Cin, Cout, W);
    S2Params p{};
    p.in = (const __half*)in; p.in_lo = (const __half*)in_lo; p.in_stats = in_stats; p.wgt = (const __half*)wgt_packed; p.out = (__half*)out; p.out_lo = (__half*)out_lo; p.out_stats = out_stats;
    static const int dbg = [] { const char* e = getenv("CDS_S2_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    p.in_act = in_act; p.has_lo = in_lo != nullptr; p.H = H; p.W = W; p.Ho = (H + 1) / 2; p.Wo = W / 2; p.n = n;
    // One staging group (all eight warps on one row) and a deep raw ring measured best: 0.250 / 0.145 ms at cfg2 (the gather form:
    // 0.320 / 0.229).  Row-owning groups (NGRP = 8 / 4) or loading straight from global memory were no faster: with every stage
    // stubbed out the barrier hand-offs alone (staging -> issuer -> epilogue, ~500 cycles per input row through the one issuer
    // thread) take 0.13 ms, which is what bounds the kernel -- see DESIGN.md section 5.
    // one CTA per SM: two resident CTAs for Cin 8 (72 registers, raw ring of 6) measured 0.259 ms against 0.242
    if (Cin == 8) return launch_s2<S2<8, 16, 8, 4, 16, 1, 1>>(p, stream);
    return launch_s2<S2<16, 32, 4, 3, 8, 1, 1>>(p, stream);
}

}  // extern "C"
