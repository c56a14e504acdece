// DynamicConv (A6) for the TRUNK layers of the feature extractor, second formulation: kernel ROWS folded into the GEMM's N
// dimension, persistent row-streaming pipeline (tcgen05.mma + TMEM + TMA, warp specialised).
//
// Reference: models/dynamic_conv.py:97-122 (att_convs / convs per kernel size, curvature gate, blend).
//
// Why: the tap-GEMM of dynconv_tc.cu issues one K=16 MMA per tap pair with N = 48..96; every such MMA re-reads its 4 KB A
// slab from shared memory, so the kernel sits on the shared-memory operand bandwidth with the tensor pipe ~30 % busy
// (DESIGN.md section 5), and the split-precision trunk (value + residual activation planes, weight residuals: three
// products per layer) triples that cost.  Here an MMA takes ONE input row and ONE horizontal tap pair
// and produces the contributions to ALL the output rows that input row reaches:
//     out[y][x] += sum_c in[R][x + dx - h][c] * W[dy][dx][c]      with y = R - dy + h,  dy = 0..k-1
// i.e. N = k x NPAD columns, column group g <-> output row y = R - h + g (weights of kernel row dy = k-1-g).  The accumulators
// of a tile of TY output rows live side by side in TMEM (slot = y - y0, one region per branch), so the sum over kernel rows is
// done by the tensor core accumulating into neighbouring column ranges and the epilogue reads finished rows exactly as before
// -- no shuffles.  MMAs per 128 pixels: conv00 12 instead of 61, conv01 9 instead of 25, each 2-5x wider, so the A-operand
// traffic per FLOP drops accordingly and the pipe runs on its math rate.
//
// Every input row is consumed by exactly one burst of MMAs, so rows STREAM through a small shared-memory ring:
//   warp 0        TMA producer: one 2 KB box per (plane, 8-channel chunk) of a 128-pixel row segment, zero fill outside
//   warps 5..7    staging: the producer layer's InstanceNorm + LeakyReLU in place (value + residual planes are summed,
//                 normalised in fp32 and re-split), fence.proxy.async, ready
//   warps 1..3    MMA issuers, one per kernel-size branch (one elected lane each; a branch's accumulation order is therefore
//                 fixed: results are bit-reproducible).  Products per row: (A_hi, W_hi), (A_hi, W_lo) and, with
//                 split-precision input, (A_lo, W_hi), all ACCUMULATING into the same columns; a slot is handed back zeroed
//                 by the epilogue (tcgen05.st) instead of being opened with accumulate = 0.
//   warps 8..15   epilogue, two sets of four (one warp per TMEM lane quadrant): curvature -> gate softmax -> blend, fp16
//                 value (+ residual) planes, curvature maps, InstanceNorm statistics for the consumer; zeroes the slot.
// CTAs are persistent over tiles (image item, 128-pixel column strip, TY rows); TMEM (all 512 columns) holds one tile and is
// recycled slot by slot, so the MMAs of the next tile start while the epilogue drains the last rows of this one.
//
// Packed weights (host: weights.py pack_dynamic_conv_kh): per branch b, image t in {hi, lo}, step j:
//   [k-chunk 2][k_b*NPAD/8][8 n][8 k] fp16;  column n = g*NPAD + c: c < Cout feature, Cout..Cout+2 curvature (a,b,c), rest 0.
//   C8 = 1: step j = horizontal taps (2j, 2j+1) (zero weights past the kernel); C8 > 1: tap (2j)/C8, channel chunks (2j)%C8, +1.
#include <algorithm>
#include <cstdlib>
#include <utility>

#include "cds_common.cuh"
#include "tc_common.cuh"
#include "tma_host.h"

namespace {

constexpr int TX = 128;
constexpr int SLAB = TX * 16;          // one 8-channel slab of a row segment
constexpr float kInEps = 1e-5f;
constexpr int NMW = 3;                 // MMA-issuing warps (one per branch; the third idles on two-branch layers)
constexpr int NSTG = 96;               // staging threads (warps 5-7)
constexpr int NTHREADS = 512;          // warps: 0 TMA, 1-3 MMA, 4 idle, 5-7 staging, 8-11 / 12-15 epilogue sets
constexpr int kEpiWarp0 = 8;
// Column groups of NPAD = Cout + 3 curvature columns rounded up to 4 (12 / 20 / 36) instead of 16 (16 / 32 / 48): a quarter
// to a third fewer B-operand bytes and math per MMA, and more output rows per TMEM tile (13 / 11 / 6 instead of 10 / 8 / 5:
// fewer halo rows).  An M = 128 MMA needs N % 16 = 0, so a range of ng groups is issued with N = roundup16(ng * NPAD): the
// up to 12 extra columns multiply either the zero padding behind the branch's last group or, for a range clipped at the
// tile's last row, the next group's real weights -- and land in the first columns of the NEXT slot, which is therefore
// always one the issuer already owns (it waits one slot ahead) and, at the tile's end, a spare slot nobody reads.
constexpr bool KH_TIGHT = true;
__host__ __device__ constexpr int kh_npad(int cout) { return KH_TIGHT ? (cout + 3 + 3) / 4 * 4 : (cout + 3 + 15) / 16 * 16; }
__host__ __device__ constexpr int kh_ncols(int cout, int k) { return KH_TIGHT ? (k * kh_npad(cout) + 15) / 16 * 16 + 16 : k * kh_npad(cout); }

// PX2 (the image layer on 8-bit images): a pixel's 8-channel operand slot holds (RGB of the pixel, RGB of its right neighbour,
// 0, 0) as byte / 256, so ONE K = 16 MMA covers FOUR horizontal taps (two slots two pixels apart) instead of two; the factor
// 256 / 255 of the loader's normalisation is folded into the weights.  byte / 256 is exact in fp16: no residual channels.
template <int K0_, int K1_, int K2_, int CIN_, int COUT_, int TY_, bool SPLIT_, bool PX2_ = false>
struct Kh {
    static constexpr bool PX2 = PX2_;
    static_assert(!PX2_ || (CIN_ == 8 && !SPLIT_), "pixel-pair slots exist for the image layer only");
    static constexpr int NK = K2_ > 0 ? 3 : 2;
    static constexpr int CIN = CIN_, COUT = COUT_, C8 = CIN_ / 8, TY = TY_;
    static constexpr bool SPLIT = SPLIT_;
    static constexpr int PLANES = SPLIT_ ? 2 : 1;
    static constexpr int NPROD = SPLIT_ ? 3 : 2;     // (A_hi, W_hi), (A_hi, W_lo), (A_lo, W_hi)
    __host__ __device__ static constexpr int kb(int b) { return b == 0 ? K0_ : (b == 1 ? K1_ : K2_); }
    __host__ __device__ static constexpr int hb(int b) { return (kb(b) - 1) / 2; }
    static constexpr int KMAX = K2_ > K1_ ? (K2_ > K0_ ? K2_ : K0_) : (K1_ > K0_ ? K1_ : K0_);
    static constexpr int HMAX = (KMAX - 1) / 2;
    static constexpr int TXO = TX - 2 * HMAX;
    static constexpr int NPAD = kh_npad(COUT_);
    static constexpr int SLOTS = KH_TIGHT ? TY_ + 1 : TY_;      // accumulator slots per branch (+ the spare one)
    __host__ __device__ static constexpr int nj(int b) { return PX2_ ? (kb(b) + 3) / 4 : (C8 == 1 ? (kb(b) + 1) / 2 : kb(b) * C8 / 2); }
    __host__ __device__ static constexpr int nbf(int b) { return kh_ncols(COUT_, kb(b)); }        // columns of a branch's weight image
    __host__ __device__ static constexpr int reg(int b) { return b * SLOTS * NPAD; }              // TMEM column base
    static constexpr int ACC_COLS = NK * SLOTS * NPAD;
    static_assert(ACC_COLS <= 512, "accumulator tile exceeds TMEM");
    __host__ __device__ static constexpr uint32_t b_img(int b) { return (uint32_t)nj(b) * 2 * nbf(b) * 16; }   // one image
    __host__ __device__ static constexpr uint32_t b_off(int b, int lo) {
        uint32_t o = 0;
        for (int i = 0; i < b; ++i) o += 2 * b_img(i);
        return o + (lo ? b_img(b) : 0);
    }
    static constexpr uint32_t B_BYTES = b_off(NK, 0);
    static constexpr uint32_t ROWB = PLANES * C8 * SLAB;
    static constexpr int NR = ROWB <= 4096 ? 12 : (ROWB <= 8192 ? 8 : 6);    // ring rows
    static constexpr int NBAR = 3 * NR + 2 * TY + 1;
    // the MMAs read up to 2*HMAX+1 pixels past a row segment: keep one spare slab of finite data behind the ring
    static constexpr size_t SMEM_USED = (size_t)NR * ROWB + SLAB + B_BYTES + NBAR * 8 + 16 + (2 * CIN + 28 + NK * COUT) * 4;
    static constexpr size_t SMEM = SMEM_USED < 120 * 1024 ? 120 * 1024 : SMEM_USED;   // > half an SM: one CTA (it owns all of TMEM)
};

struct KhParams {
    const int* img_index;     // item n reads image img_index[n] of the tensor map (NULL: n)
    const double* in_stats;   // [n][CIN][2] (sum, sumsq) of the input, or NULL: input used as is
    const float* epipole;     // [n][2]
    const __half* wgt;        // packed fp16 B images
    const float* bias;        // [NK][COUT] or NULL
    const float* gate;        // W1f [4][NK], b1 [4], W2 [NK][4]
    __half* out_raw;          // [n][H][W][COUT]
    __half* out_lo;           // optional residual plane of out_raw
    double* out_stats;        // [n][COUT][2] or NULL
    float* norm_curv;         // [n][H][W] or NULL
    float* nc_sq;             // [n][H][W] or NULL
    float* nc_abs;            // [n][H][W] or NULL
    int in_act, nc_mode, H, W, n_images;
    int n_items;              // plain batch: items; pair batch: see pair_v / pair_b
    int pair_v, pair_b;       // pair batch (conv00): items are (side, v, b); the side-0 items v*pair_b + b of b share one image
    int dbg;                  // CDS_KH_DEBUG (timing experiments only): 1 = load one 8-channel chunk per plane instead of all
    int x_pad;                // PX2: the image rows carry this many pad pixels on either side (their slots hold the true neighbours)
    int xt, yt, nz;           // tiles along x, y and item groups
    float epi_scale, inv_temperature;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    const uint32_t a = tc::smem_u32(bar);
    for (uint32_t spin = 0; !tc::mbar_try_wait(a, parity); ++spin) {
        __nanosleep(32);
        if (spin > (1u << 22)) __trap();
    }
}

// tile -> (item group z, row block, column strip); z slowest, so a CTA's consecutive tiles mostly belong to one image
struct Tile {
    int z, y0, x0, tx;
};
template <class C>
__device__ __forceinline__ Tile tile_of(const KhParams& p, int t) {
    Tile q;
    q.tx = t % p.xt;
    const int ty = (t / p.xt) % p.yt;
    q.z = t / (p.xt * p.yt);
    q.x0 = max(0, min(q.tx * C::TXO, p.W - C::TXO));   // last strip overlaps its neighbour; W < TXO: one partial strip
    q.y0 = ty * C::TY;
    return q;
}
// item group z -> first item n, item count cnt, item stride nstr (pair batch: the V reference-side items of a batch item
// share their image's MMAs, GRP at a time; every source-side item is a group of its own)
template <int GRP>
__device__ __forceinline__ void group_of(const KhParams& p, int z, int& n, int& cnt, int& nstr) {
    n = z; cnt = 1; nstr = 0;
    if constexpr (GRP > 1) {
        const int parts = (p.pair_v + GRP - 1) / GRP;
        if (z < p.pair_b * parts) {
            const int v0 = (z / p.pair_b) * GRP;
            n = v0 * p.pair_b + z % p.pair_b;
            cnt = min(GRP, p.pair_v - v0);
            nstr = p.pair_b;
        } else {
            n = p.pair_v * p.pair_b + (z - p.pair_b * parts);
        }
    }
}

// ---- MMA issue ---------------------------------------------------------------------------------------------------------------
template <class C, int W, int B, int P, int J>
__device__ __forceinline__ void issue_one(uint32_t a_row16, uint32_t b16, uint32_t d_col, uint32_t g0, uint32_t idesc, bool elected) {
    // one issuer per branch: a branch's MMAs are issued by ONE thread in program order, so the fp32 accumulation order -- and
    // with it every output bit -- is the same in every run (dealing the MMAs round-robin to all issuers was measured no faster:
    // the burst is bound by the tensor core's shared-memory operand fetch, 4 KB of A per MMA, not by the issue stream)
    if constexpr (B % NMW == W) {
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
        constexpr int hb = C::hb(B);
        constexpr uint32_t plane = P == 2 ? 1u : 0u;
        constexpr uint32_t a_start = plane * (C::C8 * (SLAB >> 4)) +
                                     (C::C8 == 1 ? (uint32_t)(C::HMAX - hb + (C::PX2 ? 4 : 2) * J)
                                                 : (uint32_t)(((2 * J) % C::C8) * (SLAB >> 4) + (C::HMAX - hb + (2 * J) / C::C8)));
        constexpr uint32_t a_lbo = C::C8 == 1 ? (C::PX2 ? 2u : 1u) : (uint32_t)(SLAB >> 4);
        constexpr uint32_t b_start = (C::b_off(B, P == 1 ? 1 : 0) + (uint32_t)J * 2 * C::nbf(B) * 16) >> 4;
        constexpr uint32_t b_lbo = (uint32_t)C::nbf(B);                 // (k_b * NPAD columns) * 16 B >> 4
        const uint64_t da = ((uint64_t)desc_hi << 32) | ((a_row16 + a_start) | (a_lbo << 16));
        const uint64_t db = ((uint64_t)desc_hi << 32) | ((b16 + b_start + g0) | (b_lbo << 16));
        if (elected) tc::mma_f16(d_col, da, db, idesc, true);
    }
}
template <class C, int W, int B, int P, int... J>
__device__ __forceinline__ void issue_prod(uint32_t a_row16, uint32_t b16, uint32_t d_col, uint32_t g0, uint32_t idesc, bool elected,
                                           std::integer_sequence<int, J...>) {
    (issue_one<C, W, B, P, J>(a_row16, b16, d_col, g0, idesc, elected), ...);
}
// input row R (absolute) of the tile starting at output row y0: issuer W's share of branch B's contributions to the tile's rows
template <class C, int W, int B>
__device__ __forceinline__ void issue_branch(const KhParams& p, int R, int y0, uint32_t a_row16, uint32_t b16, uint32_t tmem, bool elected) {
    constexpr int hb = C::hb(B);
    const int ylo = max(y0, R - hb), yhi = min(min(y0 + C::TY - 1, p.H - 1), R + hb);
    if (ylo > yhi) return;   // warp-uniform
    const uint32_t g0 = (uint32_t)(ylo - (R - hb)) * C::NPAD;                                      // first column of the weights
    const uint32_t n = ((uint32_t)(yhi - ylo + 1) * C::NPAD + 15u) & ~15u;                         // N % 16 = 0 (see KH_TIGHT)
    const uint32_t d_col = tmem + (uint32_t)C::reg(B) + (uint32_t)(ylo - y0) * C::NPAD;
    const uint32_t idesc = tc::instr_desc_f16(128, 0) | ((n >> 3) << 17);
    issue_prod<C, W, B, 0>(a_row16, b16, d_col, g0, idesc, elected, std::make_integer_sequence<int, C::nj(B)>{});
    issue_prod<C, W, B, 1>(a_row16, b16, d_col, g0, idesc, elected, std::make_integer_sequence<int, C::nj(B)>{});
    if constexpr (C::NPROD == 3) issue_prod<C, W, B, 2>(a_row16, b16, d_col, g0, idesc, elected, std::make_integer_sequence<int, C::nj(B)>{});
}
template <class C, int W>
__device__ __forceinline__ void issue_row(const KhParams& p, int R, int y0, uint32_t a_row16, uint32_t b16, uint32_t tmem, bool elected) {
    issue_branch<C, W, 0>(p, R, y0, a_row16, b16, tmem, elected);
    issue_branch<C, W, 1>(p, R, y0, a_row16, b16, tmem, elected);
    if constexpr (C::NK == 3) issue_branch<C, W, 2>(p, R, y0, a_row16, b16, tmem, elected);
}

// this warp's 32 lanes x 8 consecutive fp32 accumulator columns <- 0
__device__ __forceinline__ void tmem_zero8(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_zero4(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// this warp's lanes of one accumulator slot (every branch region) <- 0
template <class C>
__device__ __forceinline__ void zero_slot(uint32_t taddr) {
#pragma unroll
    for (int b = 0; b < C::NK; ++b) {
#pragma unroll
        for (int c = 0; c + 8 <= C::NPAD; c += 8) tmem_zero8(taddr + C::reg(b) + c);
        if constexpr (C::NPAD % 8 != 0) tmem_zero4(taddr + C::reg(b) + C::NPAD - 4);
    }
}

template <class C, int GRP>
__global__ void __launch_bounds__(NTHREADS, 1) dynconv_kh_kernel(const __grid_constant__ CUtensorMap tmap, const KhParams p) {
    constexpr int NK = C::NK, HMAX = C::HMAX, TXO = C::TXO, C8 = C::C8, CIN = C::CIN, COUT = C::COUT, NPAD = C::NPAD, TY = C::TY, NR = C::NR;
    constexpr bool STAGING = C::SPLIT || GRP == 1;     // the image layer (pair batch) takes its operand as loaded
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)NR * C::ROWB + SLAB;
    uint64_t* ring_full = reinterpret_cast<uint64_t*>(sB + C::B_BYTES);   // [NR] row landed (TMA)
    uint64_t* ring_ready = ring_full + NR;                                // [NR] row normalised (staging warps)
    uint64_t* ring_empty = ring_ready + NR;                               // [NR] row consumed (MMA commits)
    uint64_t* acc_full = ring_empty + NR;                                 // [TY] output row accumulated
    uint64_t* acc_empty = acc_full + TY;                                  // [TY] output row drained
    uint64_t* bar_b = acc_empty + TY;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 1);
    float* s_norm = reinterpret_cast<float*>(bar_b + 2);                  // [CIN][2] mean, rstd of the staged image
    float* s_gate = s_norm + 2 * CIN;                                     // W1f [4][NK], b1 [4], W2 [NK][4]  (<= 28 floats)
    float* s_bias = s_gate + 28;                                          // [NK][COUT]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);
    const int ntiles = p.xt * p.yt * p.nz;
    // A CTA takes a CONTIGUOUS run of tiles (strips of a row block, then row blocks, then items): it changes item once or twice
    // instead of nearly every tile.  Every item change costs the staging warps a reload of the InstanceNorm coefficients (fp64
    // loads in the pipeline's critical path) and the epilogue warps a flush of their statistics (320 shuffles + 64 fp64 atomics
    // per warp at 32 channels) -- with tiles dealt round-robin that was ~20 % of the quarter-resolution layers' time.
    // (The pair batch keeps round-robin dealing: its shared-image tiles cost several times a single item's, and a contiguous
    // run would give some CTAs only those.)
    const bool contiguous = GRP == 1;
    const int t_begin = contiguous ? (int)((long long)ntiles * blockIdx.x / gridDim.x) : (int)blockIdx.x;
    const int t_end = contiguous ? (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x) : ntiles;
    const int t_step = contiguous ? 1 : (int)gridDim.x;
    const bool staged = STAGING && p.in_stats != nullptr;
    constexpr int EPI_ARRIVALS = 4;   // the four warps of the set that zeroes the slot

    if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 32) {
#pragma unroll
        for (int i = 0; i < NR; ++i) { tc::mbar_init(ring_full + i, 1); tc::mbar_init(ring_ready + i, NSTG / 32); tc::mbar_init(ring_empty + i, NMW); }
#pragma unroll
        for (int i = 0; i < TY; ++i) { tc::mbar_init(acc_full + i, NMW); tc::mbar_init(acc_empty + i, EPI_ARRIVALS); }
        tc::mbar_init(bar_b, 1);
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 8 * NK + 4) s_gate[threadIdx.x - 64] = __ldg(p.gate + threadIdx.x - 64);
    for (int i = threadIdx.x; i < NK * COUT; i += NTHREADS) s_bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    // The tap shifts make an MMA read up to 2*HMAX+1 pixels PAST its row segment, i.e. the first pixels of the next slab /
    // ring slot / the spare slab behind the ring.  Those values only meet discarded rows or ZERO weights (the padding tap of
    // an odd kernel on the last valid pixel) -- but 0 x NaN is NaN, so what they read must be finite: the whole ring starts
    // zeroed (afterwards a slot holds an older row or a row in flight, both finite fp16 data).
    for (int i = threadIdx.x; i < (NR * (int)C::ROWB + SLAB) / 16; i += NTHREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // every MMA accumulates: the accumulator tile starts zeroed (and every slot is handed back zeroed by the epilogue)
    if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 8
        for (int c = 0; c < 512; c += 8) tmem_zero8(taddr + c);
        tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    if (warp == 0) {
        // ---- TMA producer ------------------------------------------------------------------------------------------------
        if (tc::elect_one()) {
            tc::mbar_expect_tx(bar_b, C::B_BYTES);
            tc::bulk_copy_g2s(sB_u, p.wgt, C::B_BYTES, bar_b);
            uint32_t rc = 0;
            for (int t = t_begin; t < t_end; t += t_step) {
                const Tile q = tile_of<C>(p, t);
                int n, cnt, nstr;
                group_of<GRP>(p, q.z, n, cnt, nstr);
                const int img = p.img_index ? __ldg(p.img_index + n) : n;
                const int r0 = max(q.y0 - HMAX, 0), r1 = min(q.y0 + TY - 1 + HMAX, p.H - 1);
                for (int R = r0; R <= r1; ++R, ++rc) {
                    const uint32_t slot = rc % NR;
                    if (rc >= (uint32_t)NR) mbar_wait_relaxed(ring_empty + slot, ((rc / NR) - 1) & 1);
                    tc::mbar_expect_tx(ring_full + slot, (p.dbg & 1) && C8 > 1 ? C::PLANES * SLAB : C::ROWB);
                    const uint32_t dst = sA_u + slot * C::ROWB;
#pragma unroll
                    for (int pl = 0; pl < C::PLANES; ++pl) {
                        if (C8 == 1) {
                            tc::tma_load_4d(dst + pl * SLAB, &tmap, ring_full + slot, 2 * (q.x0 - HMAX + p.x_pad), R, img + pl * p.n_images, 0);
                        } else {
#pragma unroll
                            for (int c8 = 0; c8 < ((p.dbg & 1) ? 1 : C8); ++c8)
                                tc::tma_load_5d(dst + (pl * C8 + c8) * SLAB, &tmap, ring_full + slot, 0, c8, q.x0 - HMAX, R, img + pl * p.n_images);
                        }
                    }
                }
            }
        }
    } else if (warp <= NMW) {
        // ---- MMA issuers: the row's MMAs are dealt round-robin to the NMW warps ------------------------------------------------
        const int mw = warp - 1;
        tc::mbar_wait(bar_b, 0);
        tc::tc_fence_after();
        const bool elected = tc::elect_one();
        const uint32_t tmem_u = tc::uniform(tmem);
        const uint32_t b16 = sB_u >> 4;
        uint32_t rc = 0, it = 0;
#pragma unroll 1
        for (int t = t_begin; t < t_end; t += t_step, ++it) {
            const Tile q = tile_of<C>(p, t);
            const int r0 = max(q.y0 - HMAX, 0), r1 = min(q.y0 + TY - 1 + HMAX, p.H - 1);
            const int ylast = min(q.y0 + TY - 1, p.H - 1);
            int waited = 0;   // slots [0, waited) of this tile have been handed over (drained and zeroed) by the epilogue
#pragma unroll 1
            for (int R = r0; R <= r1; ++R, ++rc) {
                const uint32_t slot = rc % NR;
                // Row R writes the slots up to y = R + HMAX -- and, with tight column groups, the first columns of the slot
                // after its range (N is rounded up to 16): every slot up to R + HMAX + 1 must have been handed over, including
                // the first unused slot of a partial tile.
                {
                    const int upto = min(R + HMAX + (KH_TIGHT ? 1 : 0) - q.y0, TY - 1);
                    for (; waited <= upto; ++waited) tc::mbar_wait(acc_empty + waited, (it & 1) ^ 1);
                }
                tc::mbar_wait((staged ? ring_ready : ring_full) + slot, (rc / NR) & 1);
                tc::tc_fence_after();
                const uint32_t a_row16 = (sA_u + slot * C::ROWB) >> 4;
                if (mw == 0) issue_row<C, 0>(p, R, q.y0, a_row16, b16, tmem_u, elected);
                else if (mw == 1) issue_row<C, 1>(p, R, q.y0, a_row16, b16, tmem_u, elected);
                else issue_row<C, 2>(p, R, q.y0, a_row16, b16, tmem_u, elected);
                if (elected) tc::mma_commit(ring_empty + slot);
                // output rows whose last contribution was row R
                {
                    const int ya = R - HMAX, yb = R == p.H - 1 ? ylast : R - HMAX;
                    for (int y = max(ya, q.y0); y <= min(yb, ylast); ++y)
                        if (elected) tc::mma_commit(acc_full + (y - q.y0));
                }
                __syncwarp();
            }
            // slots below the image's last row are not used by this tile: keep their barriers in step.  The arrival must still
            // be gated by the epilogue's drain of the PREVIOUS tile, like a first touch: an ungated arrival lets this side run
            // two phases ahead, and a parity wait cannot tell phase k from phase k+2 (the epilogue would wait forever)
            for (int s = ylast - q.y0 + 1; s < TY; ++s) {
                if (s >= waited) tc::mbar_wait(acc_empty + s, (it & 1) ^ 1);
                // (tight column groups: the first unused slot caught rounding columns and is cleaned by the epilogue -- which
                // must not start before those MMAs have completed: a commit, not a plain arrival)
                if (elected) {
                    if constexpr (KH_TIGHT) tc::mma_commit(acc_full + s);
                    else mbar_arrive(acc_full + s);
                }
            }
            __syncwarp();
        }
    } else if (warp >= 5 && warp < 8) {
        // ---- staging: the producer layer's InstanceNorm + activation, in place ----------------------------------------------
        if (staged) {
            const int tid = threadIdx.x - 160;   // 0 .. NSTG-1; pixel tid, and pixel tid + NSTG for the first 128 - NSTG threads
            uint32_t rc = 0;
            int cur_n = -1;
#pragma unroll 1
            for (int t = t_begin; t < t_end; t += t_step) {
                const Tile q = tile_of<C>(p, t);
                int n, cnt, nstr;
                group_of<GRP>(p, q.z, n, cnt, nstr);
                if (n != cur_n) {   // block-uniform among the staging warps
                    named_barrier(1, NSTG);
                    if (tid < CIN) {
                        const double cntp = (double)p.H * p.W;
                        const double s = p.in_stats[((size_t)n * CIN + tid) * 2], ss = p.in_stats[((size_t)n * CIN + tid) * 2 + 1];
                        const double m = s / cntp;
                        double var = ss / cntp - m * m;
                        if (var < 0.0) var = 0.0;
                        s_norm[2 * tid] = (float)m;
                        s_norm[2 * tid + 1] = (float)(1.0 / sqrt(var + (double)kInEps));
                    }
                    named_barrier(1, NSTG);
                    cur_n = n;
                }
                const int r0 = max(q.y0 - HMAX, 0), r1 = min(q.y0 + TY - 1 + HMAX, p.H - 1);
#pragma unroll 1
                for (int R = r0; R <= r1; ++R, ++rc) {
                    const uint32_t slot = rc % NR;
                    tc::mbar_wait(ring_full + slot, (rc / NR) & 1);
#pragma unroll 1
                    for (int px = tid; px < TX; px += NSTG) {
                        const int gx = q.x0 - HMAX + px;
                        if (gx < 0 || gx >= p.W) continue;   // zero fill = the conv's padding of the ACTIVATED tensor: stays zero
                        uint8_t* row = sA + (size_t)slot * C::ROWB + (size_t)px * 16;
#pragma unroll
                        for (int c8 = 0; c8 < C8; ++c8) {
                            uint4* qh = reinterpret_cast<uint4*>(row + (size_t)c8 * SLAB);
                            uint4 raw = *qh;
                            __half2* h = reinterpret_cast<__half2*>(&raw);
                            const float* nm = s_norm + c8 * 16;
                            if constexpr (C::SPLIT) {
                                uint4* ql = reinterpret_cast<uint4*>(row + (size_t)(C8 + c8) * SLAB);
                                uint4 rawl = *ql;
                                __half2* hl = reinterpret_cast<__half2*>(&rawl);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float2 f = __half22float2(h[j]), g = __half22float2(hl[j]);
                                    f.x = ((f.x + g.x) - nm[4 * j]) * nm[4 * j + 1];
                                    f.y = ((f.y + g.y) - nm[4 * j + 2]) * nm[4 * j + 3];
                                    if (p.in_act == 1) { f.x = f.x > 0.f ? f.x : 0.1f * f.x; f.y = f.y > 0.f ? f.y : 0.1f * f.y; }
                                    h[j] = __floats2half2_rn(f.x, f.y);
                                    const float2 back = __half22float2(h[j]);
                                    hl[j] = __floats2half2_rn(f.x - back.x, f.y - back.y);
                                }
                                *ql = rawl;
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float2 f = __half22float2(h[j]);
                                    f.x = (f.x - nm[4 * j]) * nm[4 * j + 1];
                                    f.y = (f.y - nm[4 * j + 2]) * nm[4 * j + 3];
                                    if (p.in_act == 1) { f.x = f.x > 0.f ? f.x : 0.1f * f.x; f.y = f.y > 0.f ? f.y : 0.1f * f.y; }
                                    h[j] = __floats2half2_rn(f.x, f.y);
                                }
                            }
                            *qh = raw;
                        }
                    }
                    tc::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's operand reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ring_ready + slot);
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ---- epilogue: gate + blend, one pixel per thread; two sets of four warps --------------------------------------------
        // Two sets of four warps.  A tile with ONE item (every layer but the image layer's shared-image groups): set e drains
        // the slots s = e (mod 2).  A shared-image group (cnt > 1 items on one set of accumulators): both sets read every slot,
        // set e handles the items it = e (mod 2), set 0 hands the slot back.
        const int set = (warp - kEpiWarp0) >> 2;
        const int lg = warp & 3;                     // TMEM lane quadrant this warp may read
        const int r = lg * 32 + lane;                // MMA row = pixel x0 + r (valid while r < TXO)
        constexpr int NIT = GRP > 1 ? (GRP + 1) / 2 : 1;   // items this set handles per slot
        float st_sum[NIT][COUT], st_sq[NIT][COUT];
#pragma unroll
        for (int i = 0; i < NIT; ++i)
#pragma unroll
            for (int c = 0; c < COUT; ++c) { st_sum[i][c] = 0.f; st_sq[i][c] = 0.f; }
        int cur_z = -1, cur_n = 0, cur_cnt = 0, cur_nstr = 0;
        auto flush = [&]() {
            if (cur_z < 0 || !p.out_stats) return;
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const int itx = (GRP > 1 && cur_cnt > 1) ? 2 * i + set : (i == 0 ? 0 : GRP);
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                    const float a = warp_sum(st_sum[i][c]), qq = warp_sum(st_sq[i][c]);
                    if (lane == 0 && itx < cur_cnt) {
                        atomicAdd(p.out_stats + ((size_t)(cur_n + itx * cur_nstr) * COUT + c) * 2, (double)a);
                        atomicAdd(p.out_stats + ((size_t)(cur_n + itx * cur_nstr) * COUT + c) * 2 + 1, (double)qq);
                    }
                    st_sum[i][c] = 0.f;
                    st_sq[i][c] = 0.f;
                }
            }
        };
        uint32_t it = 0;
#pragma unroll 1
        for (int t = t_begin; t < t_end; t += t_step, ++it) {
            const Tile q = tile_of<C>(p, t);
            if (q.z != cur_z) {
                flush();
                cur_z = q.z;
                group_of<GRP>(p, q.z, cur_n, cur_cnt, cur_nstr);
            }
            const int n = cur_n, cnt = cur_cnt, nstr = cur_nstr;
            const bool shared = GRP > 1 && cnt > 1;   // warp-uniform: several items on one set of accumulators
            const int gx = q.x0 + r;
            const bool col_ok = r < TXO && gx < p.W && gx >= q.tx * TXO;   // strips overlap at the right edge: one owner per pixel
            float ex[NIT], ey[NIT];
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const int itx = shared ? min(2 * i + set, cnt - 1) : 0;
                ex[i] = __ldg(p.epipole + 2 * (n + itx * nstr)) * p.epi_scale;
                ey[i] = __ldg(p.epipole + 2 * (n + itx * nstr) + 1) * p.epi_scale;
            }
            // the curvature accumulator is read-modify-written per pixel: fetched one slot ahead (a DRAM round trip otherwise
            // sits in every row's critical path)
            auto load_nc = [&](int s) {
                const int gy = q.y0 + s;
                const bool ok = !shared && s < TY && col_ok && gy < p.H && p.nc_sq && p.nc_mode != 0;
                return ok ? __ldcg(p.nc_sq + ((size_t)n * p.H + gy) * p.W + gx) : 0.f;
            };
            const int s_first = shared ? 0 : set, s_step = shared ? 1 : 2;
            float nc_next = load_nc(s_first);
#pragma unroll 1
            for (int s = 0; s < TY; ++s) {
                const bool mine = shared || (s & 1) == set;
                if (!mine) continue;
                const float nc_old = nc_next;
                nc_next = load_nc(s + s_step);
                tc::mbar_wait(acc_full + s, it & 1);
                tc::tc_fence_after();
                const int gy = q.y0 + s;
                const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)s * NPAD;
                if (gy >= p.H) {
                    // slot not used by this tile; with tight column groups the first unused one caught the rounding columns of
                    // the last row's ranges (real weights): clean it before it is handed back
                    if (!shared || set == 0) {
                        if constexpr (KH_TIGHT) {
                            zero_slot<C>(taddr);
                            tmem_st_wait();
                            tc::tc_fence_before();
                            __syncwarp();
                        }
                        if (lane == 0) mbar_arrive(acc_empty + s);
                    }
                    continue;
                }
                // 1. curvature columns of every branch
                uint32_t ar[NK][4];
#pragma unroll
                for (int b = 0; b < NK; ++b) tmem_ld4_nowait(taddr + C::reg(b) + COUT, ar[b]);
                tc::tmem_ld_wait();
                const bool valid = col_ok;
                float wgt[NIT][NK], ncv[NIT];
#pragma unroll
                for (int i = 0; i < NIT; ++i) {
                    if (i > 0 && !shared) { ncv[i] = 0.f; continue; }   // warp-uniform: a single-item tile has one gate per pixel
                    float uu = (float)gx - ex[i], vv = (float)gy - ey[i];
                    const float rinv = __frcp_rn(sqrtf(uu * uu + vv * vv) + 1e-6f);
                    uu *= rinv;
                    vv *= rinv;
                    float curv[NK];
#pragma unroll
                    for (int b = 0; b < NK; ++b)
                        curv[b] = (__uint_as_float(ar[b][0]) * (uu * uu) + __uint_as_float(ar[b][1]) * (2.f * uu * vv)) +
                                  __uint_as_float(ar[b][2]) * (vv * vv);
                    float hdn[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float tt = s_gate[4 * NK + j];
#pragma unroll
                        for (int b = 0; b < NK; ++b) tt += s_gate[j * NK + b] * curv[b];
                        hdn[j] = fmaxf(tt, 0.f);
                    }
                    float mx = -INFINITY;
#pragma unroll
                    for (int b = 0; b < NK; ++b) {
                        float tt = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) tt += s_gate[4 * NK + 4 + b * 4 + j] * hdn[j];
                        wgt[i][b] = tt * p.inv_temperature;
                        mx = fmaxf(mx, wgt[i][b]);
                    }
                    float den = 0.f, nc = 0.f;
#pragma unroll
                    for (int b = 0; b < NK; ++b) { wgt[i][b] = __expf(wgt[i][b] - mx); den += wgt[i][b]; }
                    const float dinv = __frcp_rn(den);
#pragma unroll
                    for (int b = 0; b < NK; ++b) { wgt[i][b] *= dinv; nc += curv[b] * wgt[i][b]; }
                    ncv[i] = nc;
                }
                // 2. feature columns, 8 channels at a time: blend the branches, store, statistics
#pragma unroll
                for (int c8 = 0; c8 < COUT / 8; ++c8) {
                    uint32_t yr[NK][8];
#pragma unroll
                    for (int b = 0; b < NK; ++b) tc::tmem_ld8_nowait(taddr + C::reg(b) + c8 * 8, yr[b]);
                    tc::tmem_ld_wait();
                    if (c8 == COUT / 8 - 1) {
                        // every accumulator column of the slot is in registers (of both sets, for a shared-image group): hand the
                        // slot back ZEROED -- every MMA accumulates, whichever issuer's MMA arrives first
                        if (shared) named_barrier(2, 256);
                        if (!shared || set == 0) {
                            zero_slot<C>(taddr);
                            tmem_st_wait();
                            tc::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(acc_empty + s);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NIT; ++i) {
                        const int itx = shared ? 2 * i + set : (i == 0 ? 0 : GRP);
                        if (itx < cnt) {   // warp-uniform
                            float out[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) out[c] = 0.f;
#pragma unroll
                            for (int b = 0; b < NK; ++b)
#pragma unroll
                                for (int c = 0; c < 8; ++c) out[c] += wgt[i][b] * (__uint_as_float(yr[b][c]) + s_bias[b * COUT + c8 * 8 + c]);
                            if (valid) {
                                const size_t m = ((size_t)(n + itx * nstr) * p.H + gy) * p.W + gx;
                                Vec8<__half>::store(p.out_raw + m * COUT + c8 * 8, out);
                                if (p.out_lo) {   // split-precision storage: what fp16 rounding just dropped
                                    float res[8];
#pragma unroll
                                    for (int c = 0; c < 8; ++c) res[c] = out[c] - __half2float(__float2half_rn(out[c]));
                                    Vec8<__half>::store(p.out_lo + m * COUT + c8 * 8, res);
                                }
#pragma unroll
                                for (int c = 0; c < 8; ++c) { st_sum[i][c8 * 8 + c] += out[c]; st_sq[i][c8 * 8 + c] += out[c] * out[c]; }
                            }
                        }
                    }
                }
                // 3. curvature maps
#pragma unroll
                for (int i = 0; i < NIT; ++i) {
                    const int itx = shared ? 2 * i + set : (i == 0 ? 0 : GRP);
                    if (itx < cnt && valid) {
                        const size_t m = ((size_t)(n + itx * nstr) * p.H + gy) * p.W + gx;
                        const float nc = ncv[i];
                        if (p.norm_curv) p.norm_curv[m] = nc;
                        if (p.nc_sq) {
                            const float old = !shared ? nc_old : (p.nc_mode != 0 ? p.nc_sq[m] : 0.f);
                            if (p.nc_mode == 0) p.nc_sq[m] = nc * nc;
                            else if (p.nc_mode == 1) p.nc_sq[m] = old + nc * nc;
                            else p.nc_sq[m] = (old + nc * nc) / 3.f;
                        }
                        if (p.nc_abs) p.nc_abs[m] = fabsf(nc);
                    }
                }
            }
        }
        flush();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int kh_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <class C, int GRP>
int launch_kh(const void* x, KhParams p, cudaStream_t st) {
    static_assert(C::SMEM <= 227 * 1024, "rows + weights do not fit in shared memory");
    auto kern = dynconv_kh_kernel<C, GRP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) { cds_set_error("cds_dynamic_conv_kh: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    CUtensorMap tmap;
    const uint64_t H = p.H, W = p.W, NI = (uint64_t)p.n_images * C::PLANES;
    bool ok;
    if (C::C8 == 1) {   // 16-byte pixels viewed as 8-byte elements: a box row is 2 KB (128 pixels)
        const uint64_t Wp = W + 2 * (uint64_t)p.x_pad;
        const uint64_t dims[4] = {2 * Wp, H, NI, 1};
        const uint64_t strides[4] = {0, Wp * 16, H * Wp * 16, NI * H * Wp * 16};
        const uint32_t box[4] = {2 * TX, 1, 1, 1};
        ok = tma::make_u64(&tmap, x, 4, dims, strides, box);
    } else {            // (8 ch, chunk, W, H, image): one box per 8-channel chunk lands as a slab
        const uint64_t dims[5] = {8, (uint64_t)C::C8, W, H, NI};
        const uint64_t strides[5] = {0, 16, (uint64_t)C::CIN * 2, W * C::CIN * 2, H * W * C::CIN * 2};
        const uint32_t box[5] = {8, 1, TX, 1, 1};
        ok = tma::make_f16(&tmap, x, 5, dims, strides, box);
    }
    if (!ok) return CDS_EUNSUPPORTED;
    static const int dbg = [] { const char* e = getenv("CDS_KH_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    p.xt = cds_div_up(p.W, C::TXO);
    p.yt = cds_div_up(p.H, C::TY);
    p.nz = GRP > 1 ? p.pair_b * ((p.pair_v + GRP - 1) / GRP) + p.pair_v * p.pair_b : p.n_items;
    const long long ntiles = (long long)p.xt * p.yt * p.nz;
    const int grid = (int)std::min<long long>(ntiles, kh_sm_count());
    kern<<<grid, NTHREADS, C::SMEM, st>>>(tmap, p);
    return cds_check_launch("cds_dynamic_conv_kh");
}

// layer shapes covered (the trunk of the feature extractor and the three output heads, models/module.py:211-231); 0 = not covered
int kh_layer_id(int Cin, int Cout, int nk, const int* ks) {
    if (!ks) return 0;
    if (Cin == 8 && Cout == 8 && nk == 3 && ks[0] == 3 && ks[1] == 7 && ks[2] == 11) return 1;   // conv00 (image padded to 8)
    if (Cin == 8 && Cout == 8 && nk == 3 && ks[0] == 3 && ks[1] == 5 && ks[2] == 7) return 2;    // conv01
    if (Cin == 16 && Cout == 16 && nk == 2 && ks[0] == 3 && ks[1] == 5) return 4;                // conv10, conv11
    if (Cin == 32 && Cout == 32 && nk == 2 && ks[0] == 1 && ks[1] == 3) return 6;                // conv20, conv21, out1
    if (Cin == 16 && Cout == 16 && nk == 2 && ks[0] == 1 && ks[1] == 3) return 7;                // out2
    if (Cin == 8 && Cout == 8 && nk == 2 && ks[0] == 1 && ks[1] == 3) return 8;                  // out3
    return 0;
}

}  // namespace

extern "C" {

int cds_dynamic_conv_kh_supported(int Cin, int Cout, int H, int W, int num_kernels, const int* ks) {
    if (W < 8 || H < 1) return 0;
    return kh_layer_id(Cin, Cout, num_kernels, ks) != 0;
}

// fp16 elements of the packed weight image: per branch two images (weights, residuals) of nj steps x 2 k-chunks x k*NPAD columns x 8
static int kh_weight_halfs(int Cin, int Cout, int num_kernels, const int* ks, bool px2) {
    const int c8 = Cin / 8;
    long long n = 0;
    for (int b = 0; b < num_kernels; ++b) {
        const int nj = px2 ? (ks[b] + 3) / 4 : (c8 == 1 ? (ks[b] + 1) / 2 : ks[b] * c8 / 2);
        n += 2ll * nj * 2 * kh_ncols(Cout, ks[b]) * 8;
    }
    return (int)n;
}
int cds_dynamic_conv_kh_weight_halfs(int Cin, int Cout, int num_kernels, const int* ks) { return kh_weight_halfs(Cin, Cout, num_kernels, ks, false); }
/* the image layer on 8-bit images (pixel-pair slots, see cds_dynamic_conv_kh_u8) */
int cds_dynamic_conv_kh_u8_weight_halfs(int Cout, int num_kernels, const int* ks) { return kh_weight_halfs(8, Cout, num_kernels, ks, true); }
int cds_dynamic_conv_kh_u8_pad(void) { return 8; }

// layout of the packed weight images (the host packer asks instead of restating it): columns per kernel-row group, columns
// of one image of a k x k branch (groups, then zero padding)
int cds_dynamic_conv_kh_group_cols(int Cout) { return kh_npad(Cout); }
int cds_dynamic_conv_kh_image_cols(int Cout, int k) { return kh_ncols(Cout, k); }

// Same contract as cds_dynamic_conv_tc (+ the pair batch of cds_dynamic_conv_tc_pairs when pair_v > 0): x fp16 [planes *
// n_images, H, W, Cin] (split_in: the residual plane follows), item i reads image img_index[i]; out_raw / out_lo fp16
// [n,H,W,Cout]; statistics, curvature maps as there.  wgt_packed: cds_dynamic_conv_kh_weight_halfs() halfs.
int cds_dynamic_conv_kh(const void* x, int n_images, const int* img_index, const double* in_stats, int in_act, const float* epipole,
                        float epi_scale, const void* wgt_packed, const float* bias, const float* gate, int n, int Cin, int Cout, int H,
                        int W, int num_kernels, const int* kernel_sizes, float temperature, int split_in, void* out_raw, void* out_lo,
                        double* out_stats, float* norm_curv, float* nc_sq, int nc_mode, float* nc_abs, int pair_v, int pair_b,
                        cudaStream_t stream) {
    CDS_REQUIRE(x && epipole && wgt_packed && gate && out_raw && kernel_sizes, CDS_EARG, "cds_dynamic_conv_kh: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && n_images > 0, CDS_ESHAPE, "cds_dynamic_conv_kh: bad batch");
    CDS_REQUIRE(temperature > 0.f, CDS_EARG, "cds_dynamic_conv_kh: temperature must be positive");
    const int lid = kh_layer_id(Cin, Cout, num_kernels, kernel_sizes);
    CDS_REQUIRE(lid != 0 && W >= 8, CDS_EUNSUPPORTED, "cds_dynamic_conv_kh: unsupported layer (Cin=%d Cout=%d W=%d)", Cin, Cout, W);
    CDS_REQUIRE(!split_in || in_stats, CDS_EARG, "cds_dynamic_conv_kh: split-precision input needs the input statistics");
    // output rows per TMEM tile: 512 columns / (branches x (rows + spare) x group columns)
    constexpr int TY8 = KH_TIGHT ? 13 : 10, TY16 = KH_TIGHT ? 11 : 8, TY32 = KH_TIGHT ? 6 : 5;
    constexpr int TY8H = KH_TIGHT ? 20 : 15;   // two-branch head of 8 channels (out3)
    KhParams p{};
    p.img_index = img_index; p.in_stats = in_stats; p.epipole = epipole; p.wgt = (const __half*)wgt_packed; p.bias = bias; p.gate = gate;
    p.out_raw = (__half*)out_raw; p.out_lo = (__half*)out_lo; p.out_stats = out_stats; p.norm_curv = norm_curv; p.nc_sq = nc_sq;
    p.nc_abs = nc_abs; p.in_act = in_act; p.nc_mode = nc_mode; p.H = H; p.W = W; p.n_images = n_images; p.n_items = n;
    p.pair_v = pair_v; p.pair_b = pair_b; p.epi_scale = epi_scale; p.inv_temperature = 1.f / temperature;
    if (pair_v > 0) {
        CDS_REQUIRE(lid == 1 && !in_stats && !split_in && img_index && n == 2 * pair_v * pair_b, CDS_EUNSUPPORTED,
                    "cds_dynamic_conv_kh: the pair batch is implemented for the image layer (3,7,11)");
        return launch_kh<Kh<3, 7, 11, 8, 8, TY8, false>, 4>(x, p, stream);
    }
    if (split_in) {
        CDS_REQUIRE(lid != 1, CDS_EUNSUPPORTED, "cds_dynamic_conv_kh: the image layer takes its residual in spare operand channels");
        CDS_REQUIRE(lid < 7, CDS_EUNSUPPORTED, "cds_dynamic_conv_kh: the stage-2/3 heads take single-plane input");
        if (lid == 2) return launch_kh<Kh<3, 5, 7, 8, 8, TY8, true>, 1>(x, p, stream);
        if (lid == 4) return launch_kh<Kh<3, 5, 0, 16, 16, TY16, true>, 1>(x, p, stream);
        return launch_kh<Kh<1, 3, 0, 32, 32, TY32, true>, 1>(x, p, stream);
    }
    if (lid == 1) return launch_kh<Kh<3, 7, 11, 8, 8, TY8, false>, 1>(x, p, stream);
    if (lid == 2) return launch_kh<Kh<3, 5, 7, 8, 8, TY8, false>, 1>(x, p, stream);
    if (lid == 4) return launch_kh<Kh<3, 5, 0, 16, 16, TY16, false>, 1>(x, p, stream);
    if (lid == 7) return launch_kh<Kh<1, 3, 0, 16, 16, TY16, false>, 1>(x, p, stream);
    if (lid == 8) return launch_kh<Kh<1, 3, 0, 8, 8, TY8H, false>, 1>(x, p, stream);
    return launch_kh<Kh<1, 3, 0, 32, 32, TY32, false>, 1>(x, p, stream);
}

// ---- the image layer on 8-bit images -----------------------------------------------------------------------------------------------
// img [n_images,3,H,W] uint8 -> px2 [n_images,H,W + 2*pad,8] fp16: slot of padded column xp (x = xp - pad) = (R,G,B of pixel x,
// R,G,B of pixel x+1, 0, 0) as byte / 256 (exact in fp16; 256 / 255 is folded into the weights), zeros outside the image
__global__ void __launch_bounds__(256) image_u8_to_px2_kernel(const unsigned char* __restrict__ img, int H, int W, int pad, __half* __restrict__ out) {
    const int Wp = W + 2 * pad, n = blockIdx.z, y = blockIdx.y, xp = blockIdx.x * 256 + threadIdx.x;
    if (xp >= Wp) return;
    const int x = xp - pad;
    const size_t plane = (size_t)H * W;
    const unsigned char* base = img + (size_t)n * 3 * plane + (size_t)y * W;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q)
        if (x + q >= 0 && x + q < W) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[3 * q + c] = (float)__ldg(base + (size_t)c * plane + x + q) * (1.f / 256.f);
        }
    Vec8<__half>::store(out + (((size_t)n * H + y) * Wp + xp) * 8, v);
}

int cds_image_u8_to_px2(const unsigned char* img, int n_images, int H, int W, void* out, cudaStream_t stream) {
    CDS_REQUIRE(img && out && n_images > 0 && n_images <= 65535 && H > 0 && H <= 65535 && W > 0, CDS_EARG, "cds_image_u8_to_px2: bad arguments");
    const int pad = cds_dynamic_conv_kh_u8_pad();
    image_u8_to_px2_kernel<<<dim3(cds_div_up(W + 2 * pad, 256), H, n_images), 256, 0, stream>>>(img, H, W, pad, (__half*)out);
    return cds_check_launch("cds_image_u8_to_px2");
}

// The image layer (3 -> 8 channels, kernel sizes 3, 7, 11; models/module.py:211) on pixel-pair slots built by
// cds_image_u8_to_px2: half the MMAs of cds_dynamic_conv_kh's image layer.  wgt_packed: cds_dynamic_conv_kh_u8_weight_halfs()
// halfs with the loader's 1/255 folded in (host: weights.py pack_dynamic_conv_kh(px2=True)).  pair_v > 0: the cascade's pair
// batch, as cds_dynamic_conv_kh.
int cds_dynamic_conv_kh_u8(const void* px2, int n_images, const int* img_index, const float* epipole, float epi_scale, const void* wgt_packed,
                           const float* gate, int n, int Cout, int H, int W, int num_kernels, const int* kernel_sizes, float temperature,
                           void* out_raw, void* out_lo, double* out_stats, float* norm_curv, float* nc_sq, int nc_mode, float* nc_abs,
                           int pair_v, int pair_b, cudaStream_t stream) {
    CDS_REQUIRE(px2 && epipole && wgt_packed && gate && out_raw && kernel_sizes, CDS_EARG, "cds_dynamic_conv_kh_u8: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && n_images > 0, CDS_ESHAPE, "cds_dynamic_conv_kh_u8: bad batch");
    CDS_REQUIRE(temperature > 0.f, CDS_EARG, "cds_dynamic_conv_kh_u8: temperature must be positive");
    CDS_REQUIRE(kh_layer_id(8, Cout, num_kernels, kernel_sizes) == 1 && W >= 8, CDS_EUNSUPPORTED,
                "cds_dynamic_conv_kh_u8: the image layer only (8 output channels, kernel sizes 3, 7, 11)");
    constexpr int TY8 = KH_TIGHT ? 13 : 10;
    KhParams p{};
    p.img_index = img_index; p.epipole = epipole; p.wgt = (const __half*)wgt_packed; p.gate = gate;
    p.out_raw = (__half*)out_raw; p.out_lo = (__half*)out_lo; p.out_stats = out_stats; p.norm_curv = norm_curv; p.nc_sq = nc_sq;
    p.nc_abs = nc_abs; p.nc_mode = nc_mode; p.H = H; p.W = W; p.n_images = n_images; p.n_items = n;
    p.pair_v = pair_v; p.pair_b = pair_b; p.epi_scale = epi_scale; p.inv_temperature = 1.f / temperature;
    p.x_pad = cds_dynamic_conv_kh_u8_pad();
    if (pair_v > 0) {
        CDS_REQUIRE(img_index && n == 2 * pair_v * pair_b, CDS_EARG, "cds_dynamic_conv_kh_u8: the pair batch needs img_index and n = 2 V B");
        return launch_kh<Kh<3, 7, 11, 8, 8, TY8, false, true>, 4>(px2, p, stream);
    }
    return launch_kh<Kh<3, 7, 11, 8, 8, TY8, false, true>, 1>(px2, p, stream);
}

}  // extern "C"
