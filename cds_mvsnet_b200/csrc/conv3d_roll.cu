// Full-resolution regulariser convolutions (conv0: C -> 8, and the prob head 8 -> 1) as a PERSISTENT, d-rolling
// tcgen05 pipeline with the kw taps folded into the GEMM's N dimension.
//
// Reference semantics: models/module.py:80-122 (Conv3d block k3 p1 s1 -> BN -> ReLU, wired at :305) and :303
// (prob = plain Conv3d(8, 1, 3, padding=1, bias=False)).
//
// Why another kernel: with Cout = 8 a tap GEMM is bound by the tensor core re-reading its 4 KB A operand from shared
// memory for every MMA (>= 32 cycles whatever N is), and conv3d_tc.cu additionally re-stages a 3-plane window per
// tile and runs load -> MMA -> epilogue back to back.  Here
//   * M = 128 consecutive voxels of a row (window origin x0-1), K = 9 (kd,kh) taps x Cin, and the three kw taps sit
//     side by side in N: D[r][kw][co] = sum_{kd,kh,ci} in[r][kd,kh,ci] w[kd,kh,kw,ci,co]; the epilogue forms
//     out[r] = D[r-1][0] + D[r][1] + D[r+1][2] with two warp shuffles per channel (+ a 2-value exchange through
//     shared memory at the warp seams).  3x fewer MMAs / A-operand reads than one MMA per (kd,kh,kw) tap.
//   * a CTA owns a (TY rows x 126 voxels) column of the volume and ROLLS along d: each input plane is fetched by TMA
//     exactly once per column into a 4-slot ring ([slab][TY+2 rows][128 voxels][8 ch], zero-filled outside the volume
//     = the conv padding), so the (kd) reuse never touches L2 again and loads overlap the MMAs of the previous plane;
//   * warp roles (608 threads): warp 0 TMA producer, warps 1-2 MMA issuers on alternating row units (accumulators: a
//     ring of 8 row units in TMEM), warps 3..18 epilogue, four groups of four (one warp per TMEM lane quadrant) (tcgen05.ld -> kw fold -> bias/ReLU -> fp16 NDHWC store, or fp32 logits for the prob head);
//   * CTAs are persistent over columns (grid = min(columns, resident CTAs)), so barrier / TMEM / weight setup is paid once.
// N columns: per kw, Cout columns of fp16-rounded weights then Cout columns of their rounding residual (summed in the
// epilogue: effectively fp32-accurate weights at no extra A traffic); Cout = 8 -> N = 48, prob head -> N = 6 (16).
// Weights (host: weights.py pack_conv3d_roll): [kd][mma][k-chunk 2][N/8][8 n][8 k] fp16; per kd the K slabs are
// Cin = 8: [kh0, kh1], [kh2, zero]; Cin > 8: kh-major, channel-chunk pairs.
#include <algorithm>
#include <utility>

#include "cds_common.cuh"
#include "tc_common.cuh"
#include "tma_host.h"

namespace {

constexpr int TX = 128;
constexpr int TXO = TX - 2;
constexpr int ROW_BYTES = TX * 16;
constexpr int NACC = 8;   // accumulator ring (row units)
constexpr int NEG = 4;    // epilogue warp groups (4 warps each); group g drains the row units u = g (mod NEG) of every plane
constexpr int NMW = 4;    // MMA-issuing warps (one thread's issue stream, ~80 cycles per small MMA, is the limit otherwise)
constexpr int NTHREADS = 32 * (1 + NMW) + NEG * 128;

template <int CIN, int COUT, int TY_>
struct RollCfg {
    static constexpr int TY = TY_;
    static constexpr int C8 = CIN / 8;
    static constexpr int CW = 2 * COUT;                       // columns per kw: weights + residual
    static constexpr int NCOL = 3 * CW;
    static constexpr int NPAD = (NCOL + 15) / 16 * 16;
    static constexpr int MMA_KD = C8 == 1 ? 2 : 3 * C8 / 2;   // MMAs per input plane
    static constexpr int NMMA = 3 * MMA_KD;
    static constexpr int ROWS = TY + 2;
    static constexpr uint32_t SLAB = ROWS * ROW_BYTES;
    static constexpr uint32_t SLOT = C8 * SLAB;
    static constexpr uint32_t B_MMA = 2 * NPAD * 16;
    static constexpr uint32_t B_BYTES = NMMA * B_MMA;
    static constexpr uint32_t XCH_FLOATS = 2 * TY * 4 * 2 * COUT;   // [plane parity][unit][warp][up|down][co]
    static_assert(TY % NEG == 0, "row units must split evenly over the epilogue groups");
    static constexpr uint32_t TMEM_COLS = NACC * NPAD <= 32 ? 32 : (NACC * NPAD <= 64 ? 64 : (NACC * NPAD <= 128 ? 128 : (NACC * NPAD <= 256 ? 256 : 512)));
    static_assert(NACC * NPAD <= 512, "accumulator ring exceeds TMEM");
    // input-plane ring: 3 planes are live, the rest is prefetch depth (a TMA round trip to DRAM is longer than one plane's MMAs)
    static constexpr uint32_t FIXED = B_BYTES + XCH_FLOATS * 4 + (2 * 8 + 2 * NACC + 1) * 8 + 16;
    static constexpr int NR_FIT = (int)((227 * 1024 - FIXED) / SLOT);
    static constexpr int NR = NR_FIT > 8 ? 8 : NR_FIT;
    static_assert(NR >= 4, "input-plane ring needs at least 4 slots");
    static constexpr int NBAR = 2 * NR + 2 * NACC + 1;
    static constexpr size_t SMEM = (size_t)NR * SLOT + B_BYTES + XCH_FLOATS * 4 + NBAR * 8 + 16;
};

struct RollParams {
    const __half* wgt;
    const float* bias;   // [COUT] (conv0) or null (prob head)
    __half* out;         // conv0: [D, H, W, 8] (one batch item)
    float* logits;       // prob head: fp32 [D, H, W]
    // prob head with the fused tail (models/model.py:85-92): softmax over D + expectation + 4-bin confidence in the epilogue
    const float* samples;   // depth hypotheses [D, H, W] of this batch item, or null: logits only
    float* depth_out;       // [H, W]
    float* conf_out;        // [H, W]
    int D, H, W, relu;
    int xt, yt;          // columns along x and y
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epilogue_barrier(int group) { asm volatile("bar.sync %0, 128;\n" ::"r"(group + 1) : "memory"); }
// producer-side wait: not latency critical, so back off instead of stealing issue slots from the epilogue warps
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    const uint32_t a = tc::smem_u32(bar);
    for (uint32_t spin = 0; !tc::mbar_try_wait(a, parity); ++spin) {
        __nanosleep(64);
        if (spin > (1u << 22)) __trap();
    }
}

// All MMAs of G consecutive row units, INTERLEAVED: for every (kd, slab pair) one MMA per unit, so that back-to-back MMAs
// target different accumulators (a chain of dependent small-N MMAs on one accumulator runs at the tensor pipe's
// accumulate latency, not at its throughput).  Every offset is a compile-time constant.
template <class C, int G, int KD, int J>
__device__ __forceinline__ void issue_mma(const uint32_t (&a16)[3], uint32_t urow16, uint32_t b16, const uint32_t (&acc)[G], bool elected) {
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
    constexpr uint32_t idesc = tc::instr_desc_f16(128, C::NPAD);
    constexpr uint32_t off0 = C::C8 == 1 ? (uint32_t)(J == 0 ? 0 : 2 * ROW_BYTES)
                                         : (uint32_t)((2 * (J % (C::C8 / 2))) * C::SLAB + (J / (C::C8 / 2)) * ROW_BYTES);
    constexpr uint32_t lbo = C::C8 == 1 ? (uint32_t)(J == 0 ? ROW_BYTES : 16) : C::SLAB;
    constexpr uint32_t a_const = (off0 >> 4) | ((lbo >> 4) << 16);
    constexpr uint32_t b_const = (((uint32_t)(KD * C::MMA_KD + J) * C::B_MMA) >> 4) | (((uint32_t)(C::NPAD * 16) >> 4) << 16);
    const uint64_t db = ((uint64_t)desc_hi << 32) | (b16 + b_const);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const uint64_t da = ((uint64_t)desc_hi << 32) | (a16[KD] + urow16 + (uint32_t)g * (ROW_BYTES >> 4) + a_const);
        if (elected) tc::mma_f16(acc[g], da, db, idesc, !(KD == 0 && J == 0));
    }
}
template <class C, int G, int KD, int... J>
__device__ __forceinline__ void issue_plane(const uint32_t (&a16)[3], uint32_t urow16, uint32_t b16, const uint32_t (&acc)[G], bool elected,
                                            std::integer_sequence<int, J...>) {
    (issue_mma<C, G, KD, J>(a16, urow16, b16, acc, elected), ...);
}
template <class C, int G>
__device__ __forceinline__ void issue_units(const uint32_t (&a16)[3], uint32_t urow16, uint32_t b16, const uint32_t (&acc)[G], bool elected) {
    issue_plane<C, G, 0>(a16, urow16, b16, acc, elected, std::make_integer_sequence<int, C::MMA_KD>{});
    issue_plane<C, G, 1>(a16, urow16, b16, acc, elected, std::make_integer_sequence<int, C::MMA_KD>{});
    issue_plane<C, G, 2>(a16, urow16, b16, acc, elected, std::make_integer_sequence<int, C::MMA_KD>{});
}

// tmap: 4-D view (2W x 8-byte elements, H, D, C/8) of one batch item's channel-blocked input, box (256, TY+2, 1, 1)
template <int CIN, int COUT, int TY>
__global__ void __launch_bounds__(NTHREADS, 1) conv3d_roll_kernel(const __grid_constant__ CUtensorMap tmap, const RollParams p) {
    using C = RollCfg<CIN, COUT, TY>;
    constexpr int C8 = C::C8, NPAD = C::NPAD, NR = C::NR;
    constexpr int G = 1;   // row units whose MMAs are interleaved by one issuer
    static_assert(TY % (G * NMW) == 0 && NACC % G == 0, "unit groups must tile the row tile and the accumulator ring");
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + NR * C::SLOT;
    float* xch = reinterpret_cast<float*>(sB + C::B_BYTES);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(xch + C::XCH_FLOATS);   // [NR]  plane landed
    uint64_t* bar_empty = bar_full + NR;                                     // [NR]  plane no longer needed
    uint64_t* acc_full = bar_empty + NR;                                     // [NACC] row unit accumulated
    uint64_t* acc_empty = acc_full + NACC;                                   // [NACC] row unit drained (4 warps)
    uint64_t* bar_b = acc_empty + NACC;                                      // weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);
    const int ntiles = p.xt * p.yt;

    if (warp == 0) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    if (threadIdx.x == 32) {
#pragma unroll
        for (int i = 0; i < NR; ++i) { tc::mbar_init(bar_full + i, 1); tc::mbar_init(bar_empty + i, NMW); }
#pragma unroll
        for (int i = 0; i < NACC; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 4); }
        tc::mbar_init(bar_b, 1);
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer: weights once, then every input plane of every column exactly once -------------------------
        if (tc::elect_one()) {
            tc::mbar_expect_tx(bar_b, C::B_BYTES);
            tc::bulk_copy_g2s(sB_u, p.wgt, C::B_BYTES, bar_b);
            uint32_t pc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int x0 = max(0, min((tile % p.xt) * TXO, p.W - TXO));
                const int y0 = (tile / p.xt) * TY;
                for (int pl = -1; pl <= p.D; ++pl, ++pc) {
                    const uint32_t slot = pc % NR;
                    if (pc >= NR) mbar_wait_relaxed(bar_empty + slot, ((pc / NR) - 1) & 1);
                    tc::mbar_expect_tx(bar_full + slot, C::SLOT);
#pragma unroll
                    for (int c8 = 0; c8 < C8; ++c8)
                        tc::tma_load_4d(sA_u + slot * C::SLOT + c8 * C::SLAB, &tmap, bar_full + slot, 2 * (x0 - 1), y0 - 1, pl, c8);
                }
            }
        }
    } else if (warp <= NMW) {
        const uint32_t mw = (uint32_t)warp - 1;   // this issuer takes the row units u = mw (mod NMW)
        // ---- MMA issuer (warp converged, one elected lane issues) -----------------------------------------------------------
        tc::mbar_wait(bar_b, 0);
        tc::tc_fence_after();
        const bool elected = tc::elect_one();
        const uint32_t tmem_u = tc::uniform(tmem);
        const uint32_t b16 = sB_u >> 4;
        uint32_t pc_base = 0, uc0 = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
            for (int d = 0; d < p.D; ++d) {
                // output plane d reads input planes d-1, d, d+1 = plane counters pc_base + d .. + 2
                if (d == 0) {
                    tc::mbar_wait(bar_full + pc_base % NR, (pc_base / NR) & 1);
                    tc::mbar_wait(bar_full + (pc_base + 1) % NR, ((pc_base + 1) / NR) & 1);
                }
                const uint32_t pc2 = pc_base + d + 2;
                tc::mbar_wait(bar_full + pc2 % NR, (pc2 / NR) & 1);
                tc::tc_fence_after();
                uint32_t a16[3];
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) a16[kd] = (sA_u + ((pc_base + d + kd) % NR) * C::SLOT) >> 4;
#pragma unroll 1
                for (uint32_t u = mw * G; u < (uint32_t)TY; u += G * NMW) {
                    const uint32_t uc = uc0 + u;
                    uint32_t acc[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const uint32_t s = (uc + g) % NACC;
                        tc::mbar_wait(acc_empty + s, (((uc + g) / NACC) & 1) ^ 1);   // epilogue has drained this accumulator
                        acc[g] = tmem_u + s * NPAD;
                    }
                    tc::tc_fence_after();
                    issue_units<C, G>(a16, (u * ROW_BYTES) >> 4, b16, acc, elected);
#pragma unroll
                    for (int g = 0; g < G; ++g)
                        if (elected) tc::mma_commit(acc_full + (uc + g) % NACC);
                    __syncwarp();
                }
                // plane d-1 is done with; the column's last output also retires planes D-1 and D
                if (elected) tc::mma_commit(bar_empty + (pc_base + d) % NR);
                if (d == p.D - 1 && elected) {
                    tc::mma_commit(bar_empty + (pc_base + d + 1) % NR);
                    tc::mma_commit(bar_empty + (pc_base + d + 2) % NR);
                }
                __syncwarp();
                uc0 += TY;
            }
            pc_base += p.D + 2;
        }
        // the last planes' release arrivals land in this CTA's shared memory: let them before the CTA may exit
        if (mw == 0 && pc_base > 0) tc::mbar_wait(bar_empty + (pc_base - 1) % NR, ((pc_base - 1) / NR) & 1);
    } else {
        // ---- epilogue warps: fold kw, bias / ReLU, store -----------------------------------------------------------------------
        const int q = warp & 3;              // TMEM lane group this warp may read
        const int grp = (warp - 1 - NMW) >> 2;   // epilogue group: row units u = grp (mod NEG)
        const int r = q * 32 + lane;         // MMA row = window voxel r (x = x0 - 1 + r); rows 1..126 are outputs
        constexpr int UPG = TY / NEG;        // row units per group and plane
        float bias[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) bias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
        uint32_t uc0 = 0, planes = 0;        // unit counter at the start of the current plane
        // fused tail (prob head only): online softmax over the planes of a column, per owned row unit -- running max, sum of
        // e^(l-m), and the two expectations (depth, plane index) rescaled together; same arithmetic, plane order and
        // intrinsics as softmax_regress_kernel (conv3d.cu), so the fused and the separate tail agree bit for bit
        float sm_m[UPG], sm_S[UPG], sm_ed[UPG], sm_ei[UPG];
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int x0 = max(0, min((tile % p.xt) * TXO, p.W - TXO));
            const int y0 = (tile / p.xt) * TY;
            const int x = x0 - 1 + r;
            // the last column strip is shifted left to stay inside the volume and overlaps its neighbour: one owner per voxel
            // (the kw fold associates differently at the warp seams, so two owners would differ in the last bit)
            const bool col_ok = r >= 1 && r <= TXO && x < p.W && x >= (tile % p.xt) * TXO;
#pragma unroll 1
            for (int d = 0; d < p.D; ++d, ++planes, uc0 += TY) {
                float* xb = xch + (planes & 1) * (TY * 4 * 2 * COUT);
                float part[UPG][COUT];
#pragma unroll
                for (int i = 0; i < UPG; ++i) {
                    const int u = i * NEG + grp;
                    const uint32_t uc = uc0 + u;
                    const uint32_t s = uc % NACC;
                    tc::mbar_wait(acc_full + s, (uc / NACC) & 1);
                    tc::tc_fence_after();
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + s * NPAD;
                    float e0[COUT], e1[COUT], e2[COUT];
                    if constexpr (COUT == 8) {
                        uint32_t v[6][8];
#pragma unroll
                        for (int j = 0; j < 6; ++j) tc::tmem_ld8_nowait(taddr + j * 8, v[j]);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            e0[c] = __uint_as_float(v[0][c]) + __uint_as_float(v[1][c]);
                            e1[c] = __uint_as_float(v[2][c]) + __uint_as_float(v[3][c]);
                            e2[c] = __uint_as_float(v[4][c]) + __uint_as_float(v[5][c]);
                        }
                    } else {
                        uint32_t v[8];
                        tc::tmem_ld8_nowait(taddr, v);
                        tc::tmem_ld_wait();
                        e0[0] = __uint_as_float(v[0]) + __uint_as_float(v[1]);
                        e1[0] = __uint_as_float(v[2]) + __uint_as_float(v[3]);
                        e2[0] = __uint_as_float(v[4]) + __uint_as_float(v[5]);
                    }
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + s);   // accumulator read: the MMA warp may refill it
#pragma unroll
                    for (int c = 0; c < COUT; ++c) {
                        const float up = __shfl_up_sync(0xffffffffu, e0[c], 1);
                        const float dn = __shfl_down_sync(0xffffffffu, e2[c], 1);
                        part[i][c] = e1[c] + (lane > 0 ? up : 0.f) + (lane < 31 ? dn : 0.f);
                    }
                    // warp seams: row 32q+31's kw=0 term belongs to row 32(q+1); row 32q's kw=2 term to row 32q-1
                    if (lane == 31) {
#pragma unroll
                        for (int c = 0; c < COUT; ++c) xb[((u * 4 + q) * 2 + 0) * COUT + c] = e0[c];
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < COUT; ++c) xb[((u * 4 + q) * 2 + 1) * COUT + c] = e2[c];
                    }
                }
                epilogue_barrier(grp);   // the group's four warps have published their seam values of this plane
#pragma unroll
                for (int i = 0; i < UPG; ++i) {
                    const int u = i * NEG + grp;
                    if (lane == 0 && q > 0) {
#pragma unroll
                        for (int c = 0; c < COUT; ++c) part[i][c] += xb[((u * 4 + q - 1) * 2 + 0) * COUT + c];
                    }
                    if (lane == 31 && q < 3) {
#pragma unroll
                        for (int c = 0; c < COUT; ++c) part[i][c] += xb[((u * 4 + q + 1) * 2 + 1) * COUT + c];
                    }
                    const int y = y0 + u;
                    if (col_ok && y < p.H) {
                        const size_t vox = ((size_t)d * p.H + y) * p.W + x;
                        if constexpr (COUT == 8) {
                            float o[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float t = part[i][c] + bias[c];
                                o[c] = p.relu ? fmaxf(t, 0.f) : t;
                            }
                            Vec8<__half>::store(p.out + vox * 8, o);
                        } else {
                            const float l = part[i][0];
                            p.logits[vox] = l;
                            if (p.samples) {
                                if (d == 0) { sm_m[i] = -INFINITY; sm_S[i] = 0.f; sm_ed[i] = 0.f; sm_ei[i] = 0.f; }
                                const float dep = __ldg(p.samples + vox);
                                if (l > sm_m[i]) {
                                    const float r = __expf(sm_m[i] - l);   // 0 on the first plane (m = -inf)
                                    sm_S[i] *= r; sm_ed[i] *= r; sm_ei[i] *= r;
                                    sm_m[i] = l;
                                }
                                const float e = __expf(l - sm_m[i]);
                                sm_S[i] += e;
                                sm_ed[i] += e * dep;
                                sm_ei[i] += e * (float)d;
                                if (d == p.D - 1) {
                                    const float inv = 1.f / sm_S[i];
                                    const size_t pix = (size_t)y * p.W + x;
                                    p.depth_out[pix] = sm_ed[i] * inv;
                                    // photometric confidence: the four probabilities around trunc(expected index)
                                    // (models/module.py:382-391); their logits were written by this thread a few planes ago
                                    const int idx = min(max((int)(sm_ei[i] * inv), 0), p.D - 1);
                                    float c = 0.f;
                                    for (int j = idx - 1; j <= idx + 2; ++j)
                                        if (j >= 0 && j < p.D) c += __expf(p.logits[((size_t)j * p.H + y) * p.W + x] - sm_m[i]) * inv;
                                    p.conf_out[pix] = c;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, C::TMEM_COLS);
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int CIN, int COUT, int TY>
int launch_roll(const void* in, const void* wgt, const float* bias, int B, int D, int H, int W, int relu, void* out, cudaStream_t st,
                const float* samples = nullptr, float* depth_out = nullptr, float* conf_out = nullptr) {
    using C = RollCfg<CIN, COUT, TY>;
    static_assert(C::SMEM <= 227 * 1024, "rolling-conv ring does not fit in shared memory");
    auto kern = conv3d_roll_kernel<CIN, COUT, TY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) { cds_set_error("cds_conv3d_k3_roll: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    RollParams p;
    p.wgt = (const __half*)wgt;
    p.bias = bias;
    p.D = D; p.H = H; p.W = W; p.relu = relu;
    p.xt = cds_div_up(W, TXO);
    p.yt = cds_div_up(H, TY);
    const int grid = std::min(p.xt * p.yt, sm_count());   // one persistent CTA per SM
    for (int b = 0; b < B; ++b) {
        const __half* base = (const __half*)in + (size_t)b * D * H * W * CIN;
        CUtensorMap tmap;
        const uint64_t dims[4] = {2 * (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)C::C8};
        const uint64_t strides[4] = {0, (uint64_t)W * 16, (uint64_t)H * W * 16, (uint64_t)D * H * W * 16};
        const uint32_t box[4] = {2 * TX, TY + 2, 1, 1};
        if (!tma::make_u64(&tmap, base, 4, dims, strides, box)) return CDS_EUNSUPPORTED;
        p.out = COUT == 1 ? nullptr : (__half*)out + (size_t)b * D * H * W * COUT;
        p.logits = COUT == 1 ? (float*)out + (size_t)b * D * H * W : nullptr;
        p.samples = samples ? samples + (size_t)b * D * H * W : nullptr;
        p.depth_out = depth_out ? depth_out + (size_t)b * H * W : nullptr;
        p.conf_out = conf_out ? conf_out + (size_t)b * H * W : nullptr;
        kern<<<grid, NTHREADS, C::SMEM, st>>>(tmap, p);
    }
    return cds_check_launch("cds_conv3d_k3_roll");
}

}  // namespace

extern "C" {

// 1 when the rolling kernel covers the layer: stride-1 k3 conv with Cout = 8 (conv0) or the 8 -> 1 prob head
int cds_conv3d_k3_roll_supported(int Cin, int Cout, int D, int H, int W) {
    if (W < 8 || D < 1 || H < 1) return 0;
    return (Cout == 8 && (Cin == 8 || Cin == 16 || Cin == 32)) || (Cin == 8 && Cout == 1);
}

int cds_conv3d_k3_roll_weight_halfs(int Cin, int Cout) {
    const int c8 = Cin / 8, npad = (6 * Cout + 15) / 16 * 16;
    return 3 * (c8 == 1 ? 2 : 3 * c8 / 2) * 2 * npad * 8;
}

// Cout == 8: out [B, D, H, W, 8] fp16 (bias + optional ReLU); Cout == 1: out = fp32 logits [B, D, H, W] (bias may be null)
int cds_conv3d_k3_roll(const void* in, const void* wgt_packed, const float* bias, int B, int Cin, int Cout, int D, int H, int W,
                       int relu, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt_packed && out && (bias || Cout == 1), CDS_EARG, "cds_conv3d_k3_roll: null pointer");
    CDS_REQUIRE(cds_conv3d_k3_roll_supported(Cin, Cout, D, H, W), CDS_EUNSUPPORTED,
                "cds_conv3d_k3_roll: unsupported shape Cin=%d Cout=%d D=%d H=%d W=%d", Cin, Cout, D, H, W);
    if (Cout == 1) return launch_roll<8, 1, 8>(in, wgt_packed, nullptr, B, D, H, W, 0, out, stream);
    if (Cin == 8) return launch_roll<8, 8, 8>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
    if (Cin == 16) return launch_roll<16, 8, 8>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
    return launch_roll<32, 8, 4>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
}

// The regulariser's tail in one kernel (models/module.py:303 prob head, models/model.py:85-92 softmax + depth_regression +
// conf_regression): in [B, D, H, W, 8] fp16 -> logits [B, D, H, W] fp32 (kept: the confidence re-reads four of them),
// depth [B, H, W] = sum_d softmax(logits)_d * samples_d, conf [B, H, W]; samples [B, D, H, W] per-pixel hypotheses.
int cds_prob_head_regress(const void* in, const void* wgt_packed, const float* samples, int B, int D, int H, int W, float* logits,
                          float* depth, float* conf, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt_packed && samples && logits && depth && conf, CDS_EARG, "cds_prob_head_regress: null pointer");
    CDS_REQUIRE(cds_conv3d_k3_roll_supported(8, 1, D, H, W), CDS_EUNSUPPORTED, "cds_prob_head_regress: unsupported shape D=%d H=%d W=%d", D, H, W);
    return launch_roll<8, 1, 8>(in, wgt_packed, nullptr, B, D, H, W, 0, logits, stream, samples, depth, conf);
}

}  // extern "C"
