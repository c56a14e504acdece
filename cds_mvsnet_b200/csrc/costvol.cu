// Cost-volume build (A2) fused with the plane-sweep warp (A1): the warped source features are
// never materialised.  Two sweeps around the visibility net (A3):
//   pass 1  per source view: similarity sum_c ref*warp per plane -> online softmax entropy over D
//   pass 2  visibility-weighted mean over views of ref (.) warp, all C channels kept, written as
//           a channel-blocked channels-last [B, C/8, D, h, w, 8] volume
// Reference: models/model.py:34-60,74 (loop), models/utils/warping.py:84-101 (coordinates, gather).
//
// Layout: features are channels-last [V, B, h, w, C] (one ref and one src map per source view,
// because the ref features depend on the pair's epipole, model.py:154-161); a thread owns one
// (pixel, 8-channel chunk) so each bilinear tap is one 16-byte (fp16) load and the C/8 lanes of
// a pixel reduce the channel sum with warp shuffles.
//
// Both kernels are instruction-issue bound (ncu: 74-78 % issue-active, IPC 3), so the work per gather is trimmed:
//   * the C/8 lanes of a pixel SHARE the projection + bilinear-footprint arithmetic (each lane does it for one plane /
//     one view and broadcasts it by shuffle), interior footprints take a branch-free fast path;
//   * a footprint always names a 2x2 block inside the image (border footprints slide and re-slot their weights), so the
//     four taps are one 64-bit multiply-add per row plus immediates;
//   * fp16 storage never converts a tap to fp32: pass 1 takes per-tap dot products on FHFMA (sm_100 mixed-precision FMA,
//     fp16 x fp16 + fp32), pass 2 blends the taps in packed half and multiplies by the reference chunk on FHFMA;
//   * fp32 storage blends and multiplies as packed fp32x2 FMAs (FFMA2).
#include <cstdlib>
#include <type_traits>
#include "cds_common.cuh"
#include "tc_common.cuh"
#include "tma_host.h"

namespace {

constexpr int kMaxViews = 8;

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2) ------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
    return d;
}

// 8 channels of one pixel as four fp32 pairs
template <typename T>
struct Pix8;
template <>
struct Pix8<__half> {
    __device__ static __forceinline__ void load(const __half* p, float2 (&v)[4]) {
        uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __half22float2(h[i]);
    }
};
template <>
struct Pix8<float> {
    __device__ static __forceinline__ void load(const float* p, float2 (&v)[4]) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p));
        float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
        v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
    }
};

// Bilinear footprint in gather form: pixel offset of the top-left pixel of a 2x2 block that ALWAYS lies inside the
// image (x in [0, w-2], y in [0, h-2]) and the weights of its four pixels.  A footprint that straddles the border is
// slid onto the nearest inside block and its weights move to the slots their pixels now occupy (0 for taps outside
// the image: zero padding, warping.py:100-101), so the four loads sit at fixed offsets from one address.
struct Foot {
    int o00;     // ya * w + xa
    float w00, w01, w10, w11;
};
__device__ __forceinline__ Foot make_foot(float u, float v, int w, int h) {
    Foot f;
    // F2I saturates (NaN -> 0, huge -> INT_MIN/MAX), so wild coordinates fail the interior test and land in make_taps
    int x0 = __float2int_rd(u), y0 = __float2int_rd(v);
    if ((unsigned)x0 < (unsigned)(w - 1) && (unsigned)y0 < (unsigned)(h - 1)) {
        float fx = u - (float)x0, fy = v - (float)y0;
        float gx = 1.f - fx, gy = 1.f - fy;
        f.o00 = y0 * w + x0;
        f.w00 = gx * gy; f.w01 = fx * gy; f.w10 = gx * fy; f.w11 = fx * fy;
    } else {
        Taps t = make_taps(u, v, w, h);   // weights of outside taps are already 0
        const int xa = min(max(t.x0, 0), w - 2), ya = min(max(t.y0, 0), h - 2);
        // column x0 sits in slot x0 - xa: -1 (x0 = -1: only the right tap is inside, it is slot 0 now), 0, or 1 (x0 = w-1:
        // only the left tap is inside, it is slot 1 now); anything further away has no inside tap at all
        if (t.x0 < xa) { t.w00 = t.w01; t.w10 = t.w11; t.w01 = 0.f; t.w11 = 0.f; }
        else if (t.x0 > xa) { t.w01 = t.w00; t.w11 = t.w10; t.w00 = 0.f; t.w10 = 0.f; }
        if (t.y0 < ya) { t.w00 = t.w10; t.w01 = t.w11; t.w10 = 0.f; t.w11 = 0.f; }
        else if (t.y0 > ya) { t.w10 = t.w00; t.w11 = t.w01; t.w00 = 0.f; t.w01 = 0.f; }
        f.o00 = ya * w + xa;
        f.w00 = t.w00; f.w01 = t.w01; f.w10 = t.w10; f.w11 = t.w11;
    }
    return f;
}
// broadcast the footprint computed by lane `src` of the warp
__device__ __forceinline__ Foot shfl_foot(const Foot& f, int src) {
    Foot g;
    g.o00 = __shfl_sync(0xffffffffu, f.o00, src);
    g.w00 = __shfl_sync(0xffffffffu, f.w00, src);
    g.w01 = __shfl_sync(0xffffffffu, f.w01, src);
    g.w10 = __shfl_sync(0xffffffffu, f.w10, src);
    g.w11 = __shfl_sync(0xffffffffu, f.w11, src);
    return g;
}

// The four tap addresses of footprint f for this thread's channel chunk: one 64-bit multiply-add for the block's first
// row, one for its second, and compile-time immediates for the right-hand column.  `pix0` is the pixel index of the
// image's first pixel inside `fea` (view / batch offset), `fea` already points at the thread's channel chunk.
template <typename T, int C>
struct TapAddr {
    const char* r0;
    const char* r1;
    static constexpr int kPix = C * (int)sizeof(T);
    __device__ __forceinline__ TapAddr(const T* fea, unsigned pix0, int w, const Foot& f) {
        const unsigned p = pix0 + (unsigned)f.o00;
        r0 = reinterpret_cast<const char*>(fea) + (size_t)p * kPix;
        r1 = reinterpret_cast<const char*>(fea) + (size_t)(p + (unsigned)w) * kPix;
    }
    __device__ __forceinline__ const T* t00() const { return reinterpret_cast<const T*>(r0); }
    __device__ __forceinline__ const T* t01() const { return reinterpret_cast<const T*>(r0 + kPix); }
    __device__ __forceinline__ const T* t10() const { return reinterpret_cast<const T*>(r1); }
    __device__ __forceinline__ const T* t11() const { return reinterpret_cast<const T*>(r1 + kPix); }
};

// bilinear blend of 8 channels at footprint f
template <typename T, int C>
__device__ __forceinline__ void gather8(const T* __restrict__ fea, unsigned pix0, int w, const Foot& f, float2 (&out)[4]) {
    const TapAddr<T, C> ta(fea, pix0, w, f);
    float2 a[4], b[4], c[4], d[4];
    Pix8<T>::load(ta.t00(), a);
    Pix8<T>::load(ta.t01(), b);
    Pix8<T>::load(ta.t10(), c);
    Pix8<T>::load(ta.t11(), d);
    const float2 w00 = make_float2(f.w00, f.w00), w01 = make_float2(f.w01, f.w01);
    const float2 w10 = make_float2(f.w10, f.w10), w11 = make_float2(f.w11, f.w11);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = ffma2(w11, d[i], ffma2(w10, c[i], ffma2(w01, b[i], fmul2(w00, a[i]))));
}

// ---- similarity of one reference pixel chunk with a footprint: sum_c ref[c] * blend[c] --------------------------
// fp32 storage: blend, then dot.  fp16 storage: the sum is re-associated as sum_t w_t (ref . tap_t), and each tap's
// dot product runs on the mixed-precision FMA of sm_100 (FHFMA: fp16 x fp16 + fp32 -> fp32, exact products, no
// fp16 -> fp32 conversions, ref stays packed in 4 registers).
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float c) {
    float d;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float dot8_f16(const uint4& r, const uint4& t) {
    const uint32_t* rr = &r.x;
    const uint32_t* tt = &t.x;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc = fhfma((uint16_t)(rr[i] & 0xffffu), (uint16_t)(tt[i] & 0xffffu), acc);
        acc = fhfma((uint16_t)(rr[i] >> 16), (uint16_t)(tt[i] >> 16), acc);
    }
    return acc;
}
__device__ __forceinline__ __half2 as_half2(uint32_t x) { return reinterpret_cast<__half2&>(x); }
__device__ __forceinline__ uint32_t as_u32(__half2 x) { return reinterpret_cast<uint32_t&>(x); }

template <typename T, int C>
struct RefChunk;
template <int C>
struct RefChunk<float, C> {
    float2 ref[4];
    __device__ __forceinline__ void load(const float* p) { Pix8<float>::load(p, ref); }
    __device__ __forceinline__ float similarity(const float* fea, unsigned pix0, int w, const Foot& f) const {
        float2 wv[4];
        gather8<float, C>(fea, pix0, w, f, wv);
        float2 s2 = fmul2(ref[0], wv[0]);
#pragma unroll
        for (int i = 1; i < 4; ++i) s2 = ffma2(ref[i], wv[i], s2);
        return s2.x + s2.y;
    }
};
template <int C>
struct RefChunk<__half, C> {
    uint4 ref;
    __device__ __forceinline__ void load(const __half* p) { ref = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ float similarity(const __half* fea, unsigned pix0, int w, const Foot& f) const {
        const TapAddr<__half, C> ta(fea, pix0, w, f);
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(ta.t00()));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(ta.t01()));
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(ta.t10()));
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(ta.t11()));
        return f.w11 * dot8_f16(ref, d) + (f.w10 * dot8_f16(ref, c) + (f.w01 * dot8_f16(ref, b) + f.w00 * dot8_f16(ref, a)));
    }
    // same with the weights as two half2 (w00, w01), (w10, w11): packed-half blend of the taps, then one FHFMA dot product
    __device__ __forceinline__ float similarity_packed(const __half* fea, int w, int o00, uint32_t wa, uint32_t wb) const {
        Foot f;
        f.o00 = o00;
        const TapAddr<__half, C> ta(fea, 0u, w, f);
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(ta.t00()));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(ta.t01()));
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(ta.t10()));
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(ta.t11()));
        const __half2 w00 = __low2half2(as_half2(wa)), w01 = __high2half2(as_half2(wa));
        const __half2 w10 = __low2half2(as_half2(wb)), w11 = __high2half2(as_half2(wb));
        const uint32_t* pa = &a.x; const uint32_t* pb = &b.x; const uint32_t* pc = &c.x; const uint32_t* pd = &d.x;
        uint4 bl;
        uint32_t* o = &bl.x;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            o[i] = as_u32(__hfma2(w11, as_half2(pd[i]), __hfma2(w10, as_half2(pc[i]), __hfma2(w01, as_half2(pb[i]), __hmul2(w00, as_half2(pa[i]))))));
        return dot8_f16(ref, bl);
    }
};

// fp32 storage, FOUR channels per lane (C / 4 lanes per pixel): one 16-byte load per tap and lane, so a warp-wide load
// covers whole 128-byte lines when C = 32 (see aggregate_f32q_kernel)
template <int C>
struct RefQuad {
    float2 ref[2];
    __device__ __forceinline__ void load(const float* p) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        ref[0] = make_float2(a.x, a.y); ref[1] = make_float2(a.z, a.w);
    }
    __device__ __forceinline__ float similarity(const float* fea, unsigned pix0, int w, const Foot& f) const {
        const TapAddr<float, C> ta(fea, pix0, w, f);
        const float4 a = __ldg(reinterpret_cast<const float4*>(ta.t00()));
        const float4 b = __ldg(reinterpret_cast<const float4*>(ta.t01()));
        const float4 c = __ldg(reinterpret_cast<const float4*>(ta.t10()));
        const float4 d = __ldg(reinterpret_cast<const float4*>(ta.t11()));
        const float2 w00 = make_float2(f.w00, f.w00), w01 = make_float2(f.w01, f.w01);
        const float2 w10 = make_float2(f.w10, f.w10), w11 = make_float2(f.w11, f.w11);
        const float2 lo = ffma2(w11, make_float2(d.x, d.y), ffma2(w10, make_float2(c.x, c.y),
                          ffma2(w01, make_float2(b.x, b.y), fmul2(w00, make_float2(a.x, a.y)))));
        const float2 hi = ffma2(w11, make_float2(d.z, d.w), ffma2(w10, make_float2(c.z, c.w),
                          ffma2(w01, make_float2(b.z, b.w), fmul2(w00, make_float2(a.z, a.w)))));
        const float2 s2 = ffma2(ref[1], hi, fmul2(ref[0], lo));
        return s2.x + s2.y;
    }
};

// ---------------------------------------------------------------------------------------------
// pass 1: entropy[v, b, y, x] = H(softmax_d(sum_c ref[c] * warp_d[c]))
// ---------------------------------------------------------------------------------------------
template <typename T, int C, bool PACKED, int CPL = 8>
__global__ void __launch_bounds__(256, 4) entropy_kernel(const T* __restrict__ ref_fea, const T* __restrict__ src_fea,
                                                      const float* __restrict__ coef, const float* __restrict__ depth,
                                                      int V, int B, int D, int h, int w, float* __restrict__ entropy) {
    static_assert(CPL == 8 || (CPL == 4 && std::is_same<T, float>::value && !PACKED), "four channels per lane: fp32 storage only");
    constexpr int LPP = C / CPL;  // lanes per pixel
    // 32-bit indexing (the host checks h*w*C < 2^31): blockIdx.y = (view, batch item), blockIdx.x tiles that image's
    // pixels.  Lanes of one pixel sit in the same warp because LPP divides 32; dead lanes of an image's last block redo
    // its last pixel so that every shuffle below is executed by full warps.
    const int P = h * w;
    const int v = blockIdx.y / B, b = blockIdx.y % B;
    const int gid = blockIdx.x * 256 + threadIdx.x;
    const bool live = gid < P * LPP;
    const int g = live ? gid : P * LPP - 1;
    const int chunk = g % LPP;
    const int pix = g / LPP;
    const int x = pix % w, y = pix / w;
    const int lane_base = (threadIdx.x & 31) & ~(LPP - 1);

    const T* rf = ref_fea + (((size_t)v * B + b) * P + (size_t)y * w + x) * C + chunk * CPL;
    // the thread's channel chunk of the source features; the (view, batch item) image starts at pixel pix0 of it
    const T* sf = src_fea + chunk * CPL;
    const unsigned pix0 = (unsigned)((v * B + b) * P);
    typename std::conditional<CPL == 4, RefQuad<C>, RefChunk<T, C>>::type ref;
    ref.load(rf);
    WarpCoef k = load_coef(coef + ((size_t)b * V + v) * 12);
    float rx, ry, rz;
    pixel_ray(k, (float)x, (float)y, rx, ry, rz);
    const float* dp = depth + (size_t)b * D * P + (size_t)y * w + x;

    // online softmax statistics: m = running max, S = sum e^(s-m), A = sum (s-m) e^(s-m).  After the channel reduction every
    // lane of the pixel holds the similarity of every plane of the step; lane `chunk` folds only plane d0 + chunk into ITS
    // statistics (one update per step instead of LPP), and the LPP partial statistics are merged once at the end.
    float m = -INFINITY, S = 0.f, A = 0.f;
    // the hypothesis of the NEXT step is fetched one step ahead: it streams from DRAM, and the gathers that depend on
    // it would otherwise see two memory latencies back to back
    float dep_next = __ldg(dp + (size_t)min(chunk, D - 1) * P);
    for (int d0 = 0; d0 < D; d0 += LPP) {
        // lane `chunk` of the pixel projects plane d0 + chunk (clamped when D is not a multiple of LPP)
        Foot mine;
        {
            const float dep = dep_next;
            dep_next = __ldg(dp + (size_t)min(d0 + LPP + chunk, D - 1) * P);
            float u, vv;
            project_fast(k, rx, ry, rz, dep, u, vv);
            mine = make_foot(u, vv, w, h);
            mine.o00 += (int)pix0;   // the footprint carries the image's offset: one add per step, not per gather
        }
        uint32_t m_wa = 0, m_wb = 0;   // PACKED (fp16 storage): the weights travel as two half2, 3 shuffles per footprint
        if constexpr (PACKED) {
            m_wa = as_u32(__floats2half2_rn(mine.w00, mine.w01));
            m_wb = as_u32(__floats2half2_rn(mine.w10, mine.w11));
        }
        float my_s = 0.f;
#pragma unroll
        for (int j = 0; j < LPP; ++j) {
            if (d0 + j >= D) break;   // warp-uniform
            float s;
            if constexpr (PACKED) {
                int o00 = mine.o00;
                uint32_t wa = m_wa, wb = m_wb;
                if (LPP > 1) {
                    o00 = __shfl_sync(0xffffffffu, mine.o00, lane_base + j);
                    wa = __shfl_sync(0xffffffffu, m_wa, lane_base + j);
                    wb = __shfl_sync(0xffffffffu, m_wb, lane_base + j);
                }
                s = ref.similarity_packed(sf, w, o00, wa, wb);
            } else {
                Foot f = LPP == 1 ? mine : shfl_foot(mine, lane_base + j);
                s = ref.similarity(sf, 0u, w, f);
            }
#pragma unroll
            for (int o = 1; o < LPP; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (j == chunk) my_s = s;
        }
        if (d0 + chunk < D) {
            const float s = my_s;
            if (s > m) {
                const float delta = m - s;  // <= 0 (or -inf on this lane's first plane)
                const float e = __expf(delta);
                const bool first = S == 0.f;
                A = first ? 0.f : e * (A + delta * S);
                S = first ? 0.f : S * e;
                m = s;
            }
            const float z = s - m;
            const float e = __expf(z);
            S += e;
            A += z * e;
        }
    }
    // merge the lanes' partial statistics: rebasing (m_i, S_i, A_i) to the common maximum M gives S_i e^(m_i - M) and
    // (A_i + (m_i - M) S_i) e^(m_i - M); a lane that saw no plane (D < LPP) contributes nothing
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), So = __shfl_xor_sync(0xffffffffu, S, o), Ao = __shfl_xor_sync(0xffffffffu, A, o);
        const float M = fmaxf(m, mo);
        const float ea = S > 0.f ? __expf(m - M) : 0.f, eb = So > 0.f ? __expf(mo - M) : 0.f;
        const float Sa = S > 0.f ? S : 0.f, da = S > 0.f ? m - M : 0.f, db = So > 0.f ? mo - M : 0.f;
        A = (A + da * Sa) * ea + (Ao + db * So) * eb;
        S = Sa * ea + So * eb;
        m = M;
    }
    if (live && chunk == 0) entropy[((size_t)v * B + b) * P + (size_t)y * w + x] = __logf(S) - A / S;
}

// ---------------------------------------------------------------------------------------------
// pass 2: volume[b, d, y, x, :] = sum_v vis_v * ref_v (.) warp_{v,d} / (sum_v vis_v + 1e-6)
// Lane `chunk` of a pixel owns the projection of views chunk, chunk + LPP, ... (its rays and translations stay in
// registers) and broadcasts each footprint to the pixel's other lanes.
// ---------------------------------------------------------------------------------------------
template <typename T, int C, int VMAX, int MINB>
__global__ void __launch_bounds__(256, MINB) aggregate_kernel(const T* __restrict__ ref_fea, const T* __restrict__ src_fea,
                                                        const float* __restrict__ coef, const float* __restrict__ depth,
                                                        const float* __restrict__ vis, int V, int B, int D, int h, int w,
                                                        T* __restrict__ volume, __half* __restrict__ vol_hi,
                                                        __half* __restrict__ vol_lo) {
    constexpr int LPP = C / 8;
    constexpr int OWN = (VMAX + LPP - 1) / LPP;   // views whose projection this lane may own (V <= VMAX)
    // 32-bit indexing: blockIdx.y = batch item, blockIdx.x tiles its pixels (see entropy_kernel)
    const int P = h * w;
    const int b = blockIdx.y;
    const int gid = blockIdx.x * 256 + threadIdx.x;
    const bool live = gid < P * LPP;
    const int g = live ? gid : P * LPP - 1;
    const int chunk = g % LPP;
    const int pix = g / LPP;
    const int x = pix % w, y = pix / w;
    const size_t pofs = (size_t)pix;
    const int lane_base = (threadIdx.x & 31) & ~(LPP - 1);

    // visibility weights of this pixel
    float vsum = 0.f;
    float vw[VMAX];
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
        vw[v] = (v < V) ? __ldg(vis + ((size_t)v * B + b) * P + pofs) : 0.f;
        if (v < V) vsum += vw[v];  // same accumulation order as the reference loop (model.py:59)
    }
    const float inv = 1.f / (vsum + 1e-6f);
    // per owned view: ray and translation
    float orx[OWN], ory[OWN], orz[OWN], otx[OWN], oty[OWN], otz[OWN];
#pragma unroll
    for (int q = 0; q < OWN; ++q) {
        int v = q * LPP + chunk;
        if (v < V) {
            WarpCoef k = load_coef(coef + ((size_t)b * V + v) * 12);
            pixel_ray(k, (float)x, (float)y, orx[q], ory[q], orz[q]);
            otx[q] = k.t[0]; oty[q] = k.t[1]; otz[q] = k.t[2] + 1e-6f;
        } else {
            orx[q] = ory[q] = orz[q] = otx[q] = oty[q] = 0.f; otz[q] = 1.f;
        }
    }
    const float* dp = depth + (size_t)b * D * P + pofs;
    // channel-blocked volume [B][C/8][D][h][w][8]: one 8-channel slab of a row is contiguous (what conv0's TMA wants)
    const size_t oofs = (((size_t)b * LPP + chunk) * D * P + pofs) * 8;
    T* outp = volume ? volume + oofs : nullptr;
    const T* rbase = ref_fea + ((size_t)b * P + pofs) * C + chunk * 8;
    const T* sbase = src_fea + chunk * 8;
    const size_t vstride = (size_t)B * P * C;
    const unsigned BP = (unsigned)(B * P), bP = (unsigned)(b * P);
    // Up to 4 views: the pixel's reference chunks stay in registers for the whole sweep, already scaled by
    // vis_v / (sum vis + 1e-6), so a plane costs one FMA per channel and view after the blend.  More views: the
    // chunks are re-read per plane (L1-resident after the first one).
    constexpr bool kRefInRegs = VMAX <= 4;
    float2 rv[kRefInRegs ? VMAX : 1][4];
    if (kRefInRegs) {
#pragma unroll
        for (int v = 0; v < VMAX; ++v) {
            const float sc = vw[v] * inv;
            if (v < V) Pix8<T>::load(rbase + (size_t)v * vstride, rv[kRefInRegs ? v : 0]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                rv[kRefInRegs ? v : 0][i] = v < V ? fmul2(rv[kRefInRegs ? v : 0][i], make_float2(sc, sc)) : make_float2(0.f, 0.f);
        }
    }

    float dep_next = __ldg(dp);
    for (int d = 0; d < D; ++d) {
        const float dep = dep_next;
        dep_next = __ldg(dp + (size_t)min(d + 1, D - 1) * P);   // one plane ahead (see entropy_kernel)
        float2 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < OWN; ++q) {
            if (q * LPP < V) {   // warp-uniform
                Foot mine;
                {
                    float px = orx[q] * dep + otx[q];
                    float py = ory[q] * dep + oty[q];
                    float iz = __frcp_rn(orz[q] * dep + otz[q]);
                    mine = make_foot(px * iz, py * iz, w, h);
                    mine.o00 += (int)(bP + (unsigned)(q * LPP + chunk) * BP);   // offset of the owned view's image
                }
#pragma unroll
                for (int j = 0; j < LPP; ++j) {
                    const int v = q * LPP + j;
                    if (v < VMAX && v < V) {   // warp-uniform
                        Foot f = LPP == 1 ? mine : shfl_foot(mine, lane_base + j);
                        float2 wv[4];
                        gather8<T, C>(sbase, 0u, w, f, wv);
                        if (kRefInRegs) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i] = ffma2(rv[kRefInRegs ? v : 0][i], wv[i], acc[i]);
                        } else {
                            float2 ref[4];
                            Pix8<T>::load(rbase + (size_t)v * vstride, ref);   // L1-resident after the first plane
                            const float2 vv2 = make_float2(vw[v] * inv, vw[v] * inv);
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i] = ffma2(fmul2(ref[i], wv[i]), vv2, acc[i]);
                        }
                    }
                }
            }
        }
        float o[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[2 * i] = acc[i].x; o[2 * i + 1] = acc[i].y; }
        if (live) {
            if (outp) Vec8<T>::store(outp + (size_t)d * P * 8, o);
            if (vol_hi) {   // split-precision fp16 volume: value plane + rounding-residual plane (what conv0's tensor cores read)
                Vec8<__half>::store(vol_hi + oofs + (size_t)d * P * 8, o);
                if (vol_lo) {
                    float res[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) res[i] = o[i] - __half2float(__float2half_rn(o[i]));
                    Vec8<__half>::store(vol_lo + oofs + (size_t)d * P * 8, res);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2, fp32 features and the split fp16 volume (stage 1 of the precise cascade), FOUR channels per lane: a pixel's
// C / 4 lanes read one tap with ONE 16-byte load each, so for C = 32 a load instruction of the warp covers four whole
// 128-byte lines (4 pixels x 8 lanes) where the 8-channel form touches eight lines with half of each used -- the sweep
// is bound by L1 wavefronts, and this halves them.  The pixel's 8 lanes share the projections of TWO planes per step
// (lane c projects view c % 4 at plane d + c / 4), so the per-thread cost of a projection is spread over two planes.
// Arithmetic per channel is the 8-channel form's (same blend order, same FMA chain over the views), so the volume is
// bit-identical.
// ---------------------------------------------------------------------------------------------
template <int C, int VMAX, int MINB>
__global__ void __launch_bounds__(256, MINB) aggregate_f32q_kernel(const float* __restrict__ ref_fea, const float* __restrict__ src_fea,
                                                                   const float* __restrict__ coef, const float* __restrict__ depth,
                                                                   const float* __restrict__ vis, int V, int B, int D, int h, int w,
                                                                   __half* __restrict__ vol_hi, __half* __restrict__ vol_lo) {
    constexpr int LPP = C / 4;
    static_assert(LPP <= 32 && LPP % VMAX == 0, "one owned (view, plane) per lane");
    constexpr int PL = LPP / VMAX;   // planes per step: lane `chunk` projects view chunk % VMAX at plane d + chunk / VMAX
    const int P = h * w;
    const int b = blockIdx.y;
    const int gid = blockIdx.x * 256 + threadIdx.x;
    const bool live = gid < P * LPP;
    const int g = live ? gid : P * LPP - 1;
    const int chunk = g % LPP;
    const int pix = g / LPP;
    const int x = pix % w, y = pix / w;
    const size_t pofs = (size_t)pix;
    const int lane_base = (threadIdx.x & 31) & ~(LPP - 1);

    float vsum = 0.f;
    float vw[VMAX];
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
        vw[v] = (v < V) ? __ldg(vis + ((size_t)v * B + b) * P + pofs) : 0.f;
        if (v < V) vsum += vw[v];  // same accumulation order as the reference loop (model.py:59)
    }
    const float inv = 1.f / (vsum + 1e-6f);
    const int own_v = chunk % VMAX, own_p = chunk / VMAX;
    float orx = 0.f, ory = 0.f, orz = 0.f, otx = 0.f, oty = 0.f, otz = 1.f;
    if (own_v < V) {
        WarpCoef k = load_coef(coef + ((size_t)b * V + own_v) * 12);
        pixel_ray(k, (float)x, (float)y, orx, ory, orz);
        otx = k.t[0]; oty = k.t[1]; otz = k.t[2] + 1e-6f;
    }
    const float* dp = depth + (size_t)b * D * P + pofs;
    // channel-blocked volume [B][C/8][D][h][w][8]: this lane writes half of an 8-channel slab
    const size_t oofs = (((size_t)b * (C / 8) + (chunk >> 1)) * D * P + pofs) * 8 + (chunk & 1) * 4;
    const float* rbase = ref_fea + ((size_t)b * P + pofs) * C + chunk * 4;
    const char* sbase = reinterpret_cast<const char*>(src_fea + chunk * 4);
    const size_t vstride = (size_t)B * P * C;
    const unsigned BP = (unsigned)(B * P), bP = (unsigned)(b * P);
    constexpr int kPix = C * (int)sizeof(float);
    // the pixel's reference channels, scaled by vis_v / (sum vis + 1e-6), stay in registers for the whole sweep
    float2 rv[VMAX][2];
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
        const float sc = vw[v] * inv;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < V) r = __ldg(reinterpret_cast<const float4*>(rbase + (size_t)v * vstride));
        rv[v][0] = v < V ? fmul2(make_float2(r.x, r.y), make_float2(sc, sc)) : make_float2(0.f, 0.f);
        rv[v][1] = v < V ? fmul2(make_float2(r.z, r.w), make_float2(sc, sc)) : make_float2(0.f, 0.f);
    }

    float dep_next = __ldg(dp + (size_t)min(own_p, D - 1) * P);
    for (int d0 = 0; d0 < D; d0 += PL) {
        const float dep = dep_next;
        dep_next = __ldg(dp + (size_t)min(d0 + PL + own_p, D - 1) * P);   // one step ahead (see entropy_kernel)
        Foot mine;
        {
            float px = orx * dep + otx;
            float py = ory * dep + oty;
            float iz = __frcp_rn(orz * dep + otz);
            mine = make_foot(px * iz, py * iz, w, h);
            mine.o00 += (int)(bP + (unsigned)own_v * BP);   // offset of the owned view's image
        }
#pragma unroll
        for (int pl = 0; pl < PL; ++pl) {
            const int d = d0 + pl;
            if (d >= D) break;   // warp-uniform
            float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
            for (int v = 0; v < VMAX; ++v) {
                if (v < V) {   // warp-uniform
                    const Foot f = shfl_foot(mine, lane_base + pl * VMAX + v);
                    const char* r0 = sbase + (size_t)(unsigned)f.o00 * kPix;
                    const char* r1 = sbase + (size_t)((unsigned)f.o00 + (unsigned)w) * kPix;
                    const float4 ta = __ldg(reinterpret_cast<const float4*>(r0));
                    const float4 tb = __ldg(reinterpret_cast<const float4*>(r0 + kPix));
                    const float4 tc_ = __ldg(reinterpret_cast<const float4*>(r1));
                    const float4 td = __ldg(reinterpret_cast<const float4*>(r1 + kPix));
                    const float2 w00 = make_float2(f.w00, f.w00), w01 = make_float2(f.w01, f.w01);
                    const float2 w10 = make_float2(f.w10, f.w10), w11 = make_float2(f.w11, f.w11);
                    const float2 lo = ffma2(w11, make_float2(td.x, td.y), ffma2(w10, make_float2(tc_.x, tc_.y),
                                      ffma2(w01, make_float2(tb.x, tb.y), fmul2(w00, make_float2(ta.x, ta.y)))));
                    const float2 hi = ffma2(w11, make_float2(td.z, td.w), ffma2(w10, make_float2(tc_.z, tc_.w),
                                      ffma2(w01, make_float2(tb.z, tb.w), fmul2(w00, make_float2(ta.z, ta.w)))));
                    acc[0] = ffma2(rv[v][0], lo, acc[0]);
                    acc[1] = ffma2(rv[v][1], hi, acc[1]);
                }
            }
            if (live) {
                const float o[4] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y};
                __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
                *reinterpret_cast<uint2*>(vol_hi + oofs + (size_t)d * P * 8) = make_uint2(as_u32(h0), as_u32(h1));
                if (vol_lo) {
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    __half2 l0 = __floats2half2_rn(o[0] - f0.x, o[1] - f0.y), l1 = __floats2half2_rn(o[2] - f1.x, o[3] - f1.y);
                    *reinterpret_cast<uint2*>(vol_lo + oofs + (size_t)d * P * 8) = make_uint2(as_u32(l0), as_u32(l1));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2, fp16 storage.  Same sweep, arithmetic arranged for the fp16 feature maps:
//   * the owner lane folds vis_v / (sum vis + 1e-6) into the four bilinear weights and packs them as two half2
//     (3 shuffles per footprint instead of 5);
//   * the blend of the four taps runs as packed HFMA2 on the raw loaded words -- no fp16 -> fp32 conversions.  Its
//     rounding (fp16 weights, fp16 partial sums) is of the size of the fp16 rounding the stored volume has anyway;
//   * the product with the reference chunk and the sum over views accumulate in fp32 through FHFMA
//     (fp16 x fp16 + fp32), with the pixel's reference chunks held packed in registers (4 per view) for the whole sweep.
// ---------------------------------------------------------------------------------------------
template <int C, int VMAX, int MINB>
__global__ void __launch_bounds__(256, MINB) aggregate_f16_kernel(const __half* __restrict__ ref_fea, const __half* __restrict__ src_fea,
                                                                  const float* __restrict__ coef, const float* __restrict__ depth,
                                                                  const float* __restrict__ vis, int V, int B, int D, int h, int w,
                                                                  __half* __restrict__ volume) {
    constexpr int LPP = C / 8;
    constexpr int OWN = (VMAX + LPP - 1) / LPP;   // views whose projection this lane may own (V <= VMAX)
    const int P = h * w;
    const int b = blockIdx.y;
    const int gid = blockIdx.x * 256 + threadIdx.x;
    const bool live = gid < P * LPP;
    const int g = live ? gid : P * LPP - 1;
    const int chunk = g % LPP;
    const int pix = g / LPP;
    const int x = pix % w, y = pix / w;
    const size_t pofs = (size_t)pix;
    const int lane_base = (threadIdx.x & 31) & ~(LPP - 1);

    float vsum = 0.f;
    for (int v = 0; v < V; ++v) vsum += __ldg(vis + ((size_t)v * B + b) * P + pofs);  // reference order (model.py:59)
    const float inv = 1.f / (vsum + 1e-6f);
    const __half* rbase = ref_fea + ((size_t)b * P + pofs) * C + chunk * 8;
    const size_t vstride = (size_t)B * P * C;
    // per owned view: ray, translation, weight scale vis_v / (sum vis + 1e-6)
    float orx[OWN], ory[OWN], orz[OWN], otx[OWN], oty[OWN], otz[OWN], osc[OWN];
#pragma unroll
    for (int q = 0; q < OWN; ++q) {
        int v = q * LPP + chunk;
        if (v < V) {
            WarpCoef k = load_coef(coef + ((size_t)b * V + v) * 12);
            pixel_ray(k, (float)x, (float)y, orx[q], ory[q], orz[q]);
            otx[q] = k.t[0]; oty[q] = k.t[1]; otz[q] = k.t[2] + 1e-6f;
            osc[q] = __ldg(vis + ((size_t)v * B + b) * P + pofs) * inv;
        } else {
            orx[q] = ory[q] = orz[q] = otx[q] = oty[q] = 0.f; otz[q] = 1.f; osc[q] = 0.f;
        }
    }
    // this thread's chunk of every view's reference features, packed fp16
    uint4 rf[VMAX];
#pragma unroll
    for (int v = 0; v < VMAX; ++v) rf[v] = v < V ? __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)v * vstride)) : make_uint4(0, 0, 0, 0);

    const float* dp = depth + (size_t)b * D * P + pofs;
    __half* outp = volume + (((size_t)b * LPP + chunk) * D * P + pofs) * 8;
    const __half* sbase = src_fea + chunk * 8;
    const unsigned BP = (unsigned)(B * P), bP = (unsigned)(b * P);

    float dep_next = __ldg(dp);
    for (int d = 0; d < D; ++d) {
        const float dep = dep_next;
        dep_next = __ldg(dp + (size_t)min(d + 1, D - 1) * P);   // one plane ahead (see entropy_kernel)
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int q = 0; q < OWN; ++q) {
            if (q * LPP < V) {   // warp-uniform
                int m_o00;
                uint32_t m_wa, m_wb;   // (w00, w01), (w10, w11) scaled by the view's visibility share
                {
                    float px = orx[q] * dep + otx[q];
                    float py = ory[q] * dep + oty[q];
                    float iz = __frcp_rn(orz[q] * dep + otz[q]);
                    Foot mine = make_foot(px * iz, py * iz, w, h);
                    m_o00 = mine.o00 + (int)(bP + (unsigned)(q * LPP + chunk) * BP);   // offset of the owned view's image
                    m_wa = as_u32(__floats2half2_rn(mine.w00 * osc[q], mine.w01 * osc[q]));
                    m_wb = as_u32(__floats2half2_rn(mine.w10 * osc[q], mine.w11 * osc[q]));
                }
#pragma unroll
                for (int j = 0; j < LPP; ++j) {
                    const int v = q * LPP + j;
                    if (v < VMAX && v < V) {   // warp-uniform
                        Foot f;
                        uint32_t wa = m_wa, wb = m_wb;
                        f.o00 = m_o00;
                        if (LPP > 1) {
                            f.o00 = __shfl_sync(0xffffffffu, m_o00, lane_base + j);
                            wa = __shfl_sync(0xffffffffu, m_wa, lane_base + j);
                            wb = __shfl_sync(0xffffffffu, m_wb, lane_base + j);
                        }
                        const TapAddr<__half, C> ta(sbase, 0u, w, f);
                        const uint4 t00 = __ldg(reinterpret_cast<const uint4*>(ta.t00()));
                        const uint4 t01 = __ldg(reinterpret_cast<const uint4*>(ta.t01()));
                        const uint4 t10 = __ldg(reinterpret_cast<const uint4*>(ta.t10()));
                        const uint4 t11 = __ldg(reinterpret_cast<const uint4*>(ta.t11()));
                        const __half2 w00 = __low2half2(as_half2(wa)), w01 = __high2half2(as_half2(wa));
                        const __half2 w10 = __low2half2(as_half2(wb)), w11 = __high2half2(as_half2(wb));
                        const uint32_t* a = &t00.x; const uint32_t* bq = &t01.x; const uint32_t* c = &t10.x; const uint32_t* e = &t11.x;
                        const uint32_t* r = &rf[v].x;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t s = as_u32(__hfma2(w11, as_half2(e[i]), __hfma2(w10, as_half2(c[i]),
                                                      __hfma2(w01, as_half2(bq[i]), __hmul2(w00, as_half2(a[i]))))));
                            acc[2 * i] = fhfma((uint16_t)(r[i] & 0xffffu), (uint16_t)(s & 0xffffu), acc[2 * i]);
                            acc[2 * i + 1] = fhfma((uint16_t)(r[i] >> 16), (uint16_t)(s >> 16), acc[2 * i + 1]);
                        }
                    }
                }
            }
        }
        if (live) Vec8<__half>::store(outp + (size_t)d * P * 8, acc);
    }
}

// mean over views of (ref_nc_sum + src_nc_sum) / 2   (model.py:60,79)
__global__ void nc_mean_kernel(const float* __restrict__ ref_nc, const float* __restrict__ src_nc, int V, long long n,
                               float* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += (__ldg(ref_nc + (size_t)v * n + i) + __ldg(src_nc + (size_t)v * n + i)) / 2.f;
    out[i] = s / (float)V;
}

// ---------------------------------------------------------------------------------------------
// Stage-3 form of both sweeps (C = 8 fp16 features, D = 8 hypotheses around the previous stage's depth): TMA-staged source
// tiles.  The D samples of a pixel lie within a few pixels of each other along its epipolar line, so the footprints of a
// 32 x 8 tile of reference pixels fall, for one source view, into a small box of the source image.  The CTA
//   1. projects its pixels' D samples into every view and reduces the bounding box of the footprints (warp shuffles + smem),
//   2. has ONE thread fetch each view's box (BW x BH pixels, 16 B each) by TMA into shared memory (out-of-image pixels arrive
//      as zeros = grid_sample's zero padding, warping.py:100-101), all views in flight at once,
//   3. gathers the 4 x D taps per view from shared memory (30-cycle LDS instead of dependent L2 round trips: the global form
//      of this sweep is latency-bound at D = 8) and finishes exactly like the global-gather kernels.
// A view whose box does not fit (depth discontinuities inside the tile, wild coordinates) takes the global-gather path for
// that view, so the result never depends on the staging.
// ---------------------------------------------------------------------------------------------
constexpr int S3_TW = 32, S3_TH = 8, S3_D = 8, S3_BW = 48, S3_BH = 16, S3_VMAX = 4;
constexpr int S3_BOX_BYTES = S3_BW * S3_BH * 16;

struct S3Params {
    const __half* ref_fea;   // [V,B,h,w,8]
    const __half* src_fea;   // [V,B,h,w,8] (global-gather path)
    const float* coef;       // [B,V,12]
    const float* depth;      // [B,8,h,w]
    const float* vis;        // [V,B,h,w]   (aggregate)
    float* entropy;          // [V,B,h,w]   (entropy)
    __half* volume;          // [B,1,8,h,w,8] (aggregate)
    int V, B, h, w;
};

// sample position of depth `dep` for a pixel ray: explicit FMAs, so that the bounding-box pass and the sweep compute the
// SAME bits (a coordinate that differed by an ulp across an integer could step outside the staged box)
__device__ __forceinline__ void s3_project(float rx, float ry, float rz, float tx, float ty, float tz, float dep, float& u, float& v) {
    const float iz = __frcp_rn(__fmaf_rn(rz, dep, tz));
    u = __fmul_rn(__fmaf_rn(rx, dep, tx), iz);
    v = __fmul_rn(__fmaf_rn(ry, dep, ty), iz);
}

// the 4 taps (raw 8-channel chunks) and bilinear weights of one sample: from the staged box, or from global memory
struct S3Taps {
    uint4 t00, t01, t10, t11;
    float w00, w01, w10, w11;
};
__device__ __forceinline__ S3Taps s3_taps(bool staged, const uint8_t* box, int bx, int by, const __half* gsrc, unsigned pix0, int w, int h,
                                          float u, float vv) {
    S3Taps t;
    if (staged) {   // block-uniform
        const float xf = floorf(u), yf = floorf(vv);
        const float fx = u - xf, fy = vv - yf, gx = 1.f - fx, gy = 1.f - fy;
        t.w00 = gx * gy; t.w01 = fx * gy; t.w10 = gx * fy; t.w11 = fx * fy;
        const uint8_t* q = box + (((int)yf - by) * S3_BW + ((int)xf - bx)) * 16;
        t.t00 = *reinterpret_cast<const uint4*>(q);
        t.t01 = *reinterpret_cast<const uint4*>(q + 16);
        t.t10 = *reinterpret_cast<const uint4*>(q + S3_BW * 16);
        t.t11 = *reinterpret_cast<const uint4*>(q + S3_BW * 16 + 16);
    } else {
        const Foot f = make_foot(u, vv, w, h);
        const TapAddr<__half, 8> ta(gsrc, pix0, w, f);
        t.t00 = __ldg(reinterpret_cast<const uint4*>(ta.t00())); t.t01 = __ldg(reinterpret_cast<const uint4*>(ta.t01()));
        t.t10 = __ldg(reinterpret_cast<const uint4*>(ta.t10())); t.t11 = __ldg(reinterpret_cast<const uint4*>(ta.t11()));
        t.w00 = f.w00; t.w01 = f.w01; t.w10 = f.w10; t.w11 = f.w11;
    }
    return t;
}

template <int MODE>   // 0: similarity entropy per view; 1: visibility-weighted aggregate
__global__ void __launch_bounds__(256, 4) costvol_s3_tma_kernel(const __grid_constant__ CUtensorMap tmap, const S3Params p) {
    extern __shared__ __align__(128) uint8_t s3_smem[];
    uint8_t (*s_box)[S3_BOX_BYTES] = reinterpret_cast<uint8_t (*)[S3_BOX_BYTES]>(s3_smem);
    __shared__ float s_red[S3_VMAX][8][4];
    __shared__ int s_org[S3_VMAX][3];          // box origin x, y, staged flag
    __shared__ __align__(8) uint64_t s_bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, P = p.h * p.w;
    const int x = min((int)blockIdx.x * S3_TW + lane, p.w - 1), y = min((int)blockIdx.y * S3_TH + warp, p.h - 1);
    const bool live = (int)blockIdx.x * S3_TW + lane < p.w && (int)blockIdx.y * S3_TH + warp < p.h;
    const size_t pofs = (size_t)y * p.w + x;
    if (threadIdx.x == 0) { tc::mbar_init(&s_bar, 1); tc::mbar_fence_init(); }

    float dep[S3_D];
#pragma unroll
    for (int d = 0; d < S3_D; ++d) dep[d] = __ldg(p.depth + ((size_t)b * S3_D + d) * P + pofs);

    bool ascending = true;   // the cascade's hypotheses are (module.py:394-415); any other caller falls back to the global gathers
#pragma unroll
    for (int d = 1; d < S3_D; ++d) ascending = ascending && dep[d - 1] <= dep[d];
    // ---- 1. bounding boxes of the footprints, all views ------------------------------------------------------------------
    float rx[S3_VMAX], ry[S3_VMAX], rz[S3_VMAX], tx[S3_VMAX], ty[S3_VMAX], tz[S3_VMAX];
#pragma unroll
    for (int v = 0; v < S3_VMAX; ++v) {
        if (v < p.V) {
            const WarpCoef k = load_coef(p.coef + ((size_t)b * p.V + v) * 12);
            pixel_ray(k, (float)x, (float)y, rx[v], ry[v], rz[v]);
            tx[v] = k.t[0]; ty[v] = k.t[1]; tz[v] = k.t[2] + 1e-6f;
            // a sample moves MONOTONICALLY along its epipolar line with depth (u, v are Moebius functions of the depth while the
            // denominator keeps its sign), and the hypotheses are ascending: the extremes are taken at the first and last plane
            float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
            const bool same_side = (__fmaf_rn(rz[v], dep[0], tz[v]) > 0.f) == (__fmaf_rn(rz[v], dep[S3_D - 1], tz[v]) > 0.f);
#pragma unroll
            for (int d = 0; d < S3_D; d += S3_D - 1) {
                float u, vv;
                s3_project(rx[v], ry[v], rz[v], tx[v], ty[v], tz[v], dep[d], u, vv);
                // NaN-proof: a NaN coordinate makes the box infinite, which sends the view to the global path
                umin = (u < umin) ? u : (u == u ? umin : -INFINITY); umax = (u > umax) ? u : (u == u ? umax : INFINITY);
                vmin = (vv < vmin) ? vv : (vv == vv ? vmin : -INFINITY); vmax = (vv > vmax) ? vv : (vv == vv ? vmax : INFINITY);
            }
            if (!same_side || !ascending) { umin = vmin = -INFINITY; umax = vmax = INFINITY; }   // pole inside the range / arbitrary planes: global path
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o)); umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, o));
                vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o)); vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            }
            if (lane == 0) { s_red[v][warp][0] = umin; s_red[v][warp][1] = umax; s_red[v][warp][2] = vmin; s_red[v][warp][3] = vmax; }
        } else {
            rx[v] = ry[v] = rz[v] = tx[v] = ty[v] = 0.f; tz[v] = 1.f;
        }
    }
    __syncthreads();
    // ---- 2. one thread: box origins, TMA fetches of every view's box ---------------------------------------------------------
    if (threadIdx.x == 0) {
        uint32_t bytes = 0;
        for (int v = 0; v < p.V; ++v) {
            float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
            for (int i = 0; i < 8; ++i) {
                umin = fminf(umin, s_red[v][i][0]); umax = fmaxf(umax, s_red[v][i][1]);
                vmin = fminf(vmin, s_red[v][i][2]); vmax = fmaxf(vmax, s_red[v][i][3]);
            }
            // footprint columns floor(umin) .. floor(umax) + 1, rows likewise; staged when the box holds them and the
            // coordinates are small integers (everything far outside the image is zero anyway: global path)
            int staged = 0, bx = 0, by = 0;
            if (umin > -65536.f && umax < 65536.f && vmin > -65536.f && vmax < 65536.f) {
                bx = (int)floorf(umin); by = (int)floorf(vmin);
                const int ex = (int)floorf(umax) + 1, ey = (int)floorf(vmax) + 1;
                staged = (ex - bx < S3_BW && ey - by < S3_BH) ? 1 : 0;
            }
            s_org[v][0] = bx; s_org[v][1] = by; s_org[v][2] = staged;
            if (staged) bytes += S3_BOX_BYTES;
        }
        if (bytes) {
            tc::mbar_expect_tx(&s_bar, bytes);
            for (int v = 0; v < p.V; ++v)
                if (s_org[v][2])
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
                                 "r"(tc::smem_u32(&s_box[v][0])), "l"(&tmap), "r"(tc::smem_u32(&s_bar)), "r"(2 * s_org[v][0]), "r"(s_org[v][1]),
                                 "r"(v * p.B + b) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(&s_bar)) : "memory");
        }
    }
    // reference chunks / visibility shares while the boxes are in flight
    uint4 rf[S3_VMAX];
    float sc[S3_VMAX];
    float vsum = 0.f;
#pragma unroll
    for (int v = 0; v < S3_VMAX; ++v) {
        rf[v] = v < p.V ? __ldg(reinterpret_cast<const uint4*>(p.ref_fea + (((size_t)v * p.B + b) * P + pofs) * 8)) : make_uint4(0, 0, 0, 0);
        sc[v] = 0.f;
        if (MODE == 1 && v < p.V) { sc[v] = __ldg(p.vis + ((size_t)v * p.B + b) * P + pofs); vsum += sc[v]; }   // reference order (model.py:59)
    }
    if (MODE == 1) {
        const float inv = 1.f / (vsum + 1e-6f);
#pragma unroll
        for (int v = 0; v < S3_VMAX; ++v) sc[v] *= inv;
    }
    __syncthreads();                 // s_org visible
    tc::mbar_wait(&s_bar, 0);        // every staged box has landed

    // ---- 3. the sweep ------------------------------------------------------------------------------------------------------
    if (MODE == 0) {
#pragma unroll
        for (int v = 0; v < S3_VMAX; ++v) {
            if (v >= p.V) break;   // block-uniform
            const int bx = s_org[v][0], by = s_org[v][1];
            const bool staged = s_org[v][2] != 0;
            const unsigned pix0 = (unsigned)((v * p.B + b) * P);
            float m = -INFINITY, S = 0.f, A = 0.f;
#pragma unroll
            for (int d = 0; d < S3_D; ++d) {
                float u, vv;
                s3_project(rx[v], ry[v], rz[v], tx[v], ty[v], tz[v], dep[d], u, vv);
                const S3Taps t = s3_taps(staged, &s_box[v][0], bx, by, p.src_fea, pix0, p.w, p.h, u, vv);
                const float s = t.w11 * dot8_f16(rf[v], t.t11) + (t.w10 * dot8_f16(rf[v], t.t10) + (t.w01 * dot8_f16(rf[v], t.t01) + t.w00 * dot8_f16(rf[v], t.t00)));
                if (s > m) {
                    const float delta = m - s;
                    const float e = __expf(delta);
                    const bool first = S == 0.f;
                    A = first ? 0.f : e * (A + delta * S);
                    S = first ? 0.f : S * e;
                    m = s;
                }
                const float z = s - m, e = __expf(z);
                S += e;
                A += z * e;
            }
            if (live) p.entropy[((size_t)v * p.B + b) * P + pofs] = __logf(S) - A / S;
        }
    } else {
        __half* outp = p.volume + ((size_t)b * S3_D * P + pofs) * 8;
#pragma unroll 1
        for (int d = 0; d < S3_D; ++d) {
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.f;
            const float depd = __ldg(p.depth + ((size_t)b * S3_D + d) * P + pofs);   // (L1 hit; keeps dep[] out of a dynamic index)
#pragma unroll
            for (int v = 0; v < S3_VMAX; ++v) {
                if (v >= p.V) break;   // block-uniform
                float u, vv;
                s3_project(rx[v], ry[v], rz[v], tx[v], ty[v], tz[v], depd, u, vv);
                const S3Taps t = s3_taps(s_org[v][2] != 0, &s_box[v][0], s_org[v][0], s_org[v][1], p.src_fea, (unsigned)((v * p.B + b) * P), p.w,
                                         p.h, u, vv);
                // same arithmetic as aggregate_f16_kernel: visibility share folded into the weights, packed-half blend, FHFMA
                const __half2 h00 = __floats2half2_rn(t.w00 * sc[v], t.w00 * sc[v]), h01 = __floats2half2_rn(t.w01 * sc[v], t.w01 * sc[v]);
                const __half2 h10 = __floats2half2_rn(t.w10 * sc[v], t.w10 * sc[v]), h11 = __floats2half2_rn(t.w11 * sc[v], t.w11 * sc[v]);
                const uint32_t* a = &t.t00.x; const uint32_t* bq = &t.t01.x; const uint32_t* c = &t.t10.x; const uint32_t* e = &t.t11.x;
                const uint32_t* r = &rf[v].x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t sblend = as_u32(__hfma2(h11, as_half2(e[i]), __hfma2(h10, as_half2(c[i]),
                                                   __hfma2(h01, as_half2(bq[i]), __hmul2(h00, as_half2(a[i]))))));
                    acc[2 * i] = fhfma((uint16_t)(r[i] & 0xffffu), (uint16_t)(sblend & 0xffffu), acc[2 * i]);
                    acc[2 * i + 1] = fhfma((uint16_t)(r[i] >> 16), (uint16_t)(sblend >> 16), acc[2 * i + 1]);
                }
            }
            if (live) Vec8<__half>::store(outp + (size_t)d * P * 8, acc);
        }
    }
}

template <int MODE>
int launch_s3_tma(const S3Params& p, cudaStream_t st) {
    CUtensorMap tmap;
    const uint64_t dims[3] = {2 * (uint64_t)p.w, (uint64_t)p.h, (uint64_t)p.V * p.B};
    const uint64_t strides[3] = {0, (uint64_t)p.w * 16, (uint64_t)p.h * p.w * 16};
    const uint32_t box[3] = {2 * S3_BW, S3_BH, 1};
    if (!tma::make_u64(&tmap, p.src_fea, 3, dims, strides, box)) return CDS_EUNSUPPORTED;
    dim3 grid(cds_div_up(p.w, S3_TW), cds_div_up(p.h, S3_TH), p.B);
    constexpr int smem = S3_VMAX * S3_BOX_BYTES;
    cudaError_t e = cudaFuncSetAttribute(costvol_s3_tma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { cds_set_error("cds_costvol: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    costvol_s3_tma_kernel<MODE><<<grid, 256, smem, st>>>(tmap, p);
    return cds_check_launch(MODE == 0 ? "cds_costvol_entropy" : "cds_costvol_aggregate");
}
// the staged form covers the last cascade stage's shape; CDS_COSTVOL_TMA=0 keeps the global-gather kernels
bool s3_tma_applies(int V, int B, int C, int D, int h, int w, int dtype) {
    static const bool on = [] { const char* e = getenv("CDS_COSTVOL_TMA"); return !(e && e[0] == '0'); }();
    return on && dtype == CDS_F16 && C == 8 && D == S3_D && V <= S3_VMAX && B <= 65535 && w >= 8 && h >= 2;
}

template <typename T>
int launch_entropy(const void* ref, const void* src, const float* coef, const float* depth, int V, int B, int C, int D,
                   int h, int w, float* entropy, cudaStream_t st, bool fast = false) {
    dim3 blocks(cds_div_up((long long)h * w * (C / 8), 256), V * B);
    const T* r = (const T*)ref;
    const T* s = (const T*)src;
    // fp16 storage: fp32 tap weights on per-tap FHFMA dot products.  The `fast` entry (cds_costvol_entropy_fast) takes the
    // packed-half blend of aggregate_f16_kernel: 10-12 % faster, the entropy's error grows from < 2e-4 to 5e-4 (measured on
    // B200).  CDS_ENTROPY_PACKED=1 / 0 forces it on / off for both entries.
    static const int forced = [] { const char* e = getenv("CDS_ENTROPY_PACKED"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    const bool packed = forced < 0 ? fast : forced == 1;
    constexpr bool kHalf = std::is_same<T, __half>::value;
#define CDS_ENT(c)                                                                                                          \
    if (kHalf && packed) entropy_kernel<T, c, kHalf><<<blocks, 256, 0, st>>>(r, s, coef, depth, V, B, D, h, w, entropy); \
    else entropy_kernel<T, c, false><<<blocks, 256, 0, st>>>(r, s, coef, depth, V, B, D, h, w, entropy);                    \
    break;
    if constexpr (!kHalf) {
        // fp32 features with 32 channels (stage 1 of the precise cascade): four channels per lane, 8 lanes per pixel
        static const bool quad = [] { const char* e = getenv("CDS_ENT_QUAD"); return !(e && e[0] == '0'); }();
        if (quad && C == 32) {
            dim3 qb(cds_div_up((long long)h * w * (C / 4), 256), V * B);
            entropy_kernel<float, 32, false, 4><<<qb, 256, 0, st>>>(r, s, coef, depth, V, B, D, h, w, entropy);
            return cds_check_launch("cds_costvol_entropy");
        }
    }
    switch (C) {
        case 8: CDS_ENT(8)
        case 16: CDS_ENT(16)
        case 32: CDS_ENT(32)
        default: cds_set_error("cds_costvol_entropy: C must be 8, 16 or 32 (got %d)", C); return CDS_EUNSUPPORTED;
    }
    return cds_check_launch("cds_costvol_entropy");
}

template <typename T>
int launch_aggregate(const void* ref, const void* src, const float* coef, const float* depth, const float* vis, int V,
                     int B, int C, int D, int h, int w, void* volume, cudaStream_t st, __half* vol_hi = nullptr,
                     __half* vol_lo = nullptr) {
    dim3 blocks(cds_div_up((long long)h * w * (C / 8), 256), B);
    const T* r = (const T*)ref;
    const T* s = (const T*)src;
    T* o = (T*)volume;
    // fp32 storage (and CDS_COSTVOL_BLEND=f32): fp32 blend.  fp16 storage: packed-half blend, see aggregate_f16_kernel.
    // The fp32-feature form with the split fp16 volume (stage 1 of the precise cascade) runs two resident blocks: at three it
    // spills (measured 0.878 -> 0.740 ms at cfg2).
    static const int occ = [] { const char* e = getenv("CDS_AGG_OCC"); return e ? atoi(e) : 0; }();
    static const bool blend32 = [] { const char* e = getenv("CDS_COSTVOL_BLEND"); return e && e[0] == 'f' && e[1] == '3'; }();
    if constexpr (std::is_same<T, __half>::value) {
        if (!blend32) {
#define CDS_AGGH(c, ob)                                                                                                     \
    if (V <= 4 && (occ ? occ : ob) == 4) aggregate_f16_kernel<c, 4, 4><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o); \
    else if (V <= 4 && (occ ? occ : ob) == 3) aggregate_f16_kernel<c, 4, 3><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o); \
    else if (V <= 4) aggregate_f16_kernel<c, 4, 2><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o);        \
    else aggregate_f16_kernel<c, kMaxViews, 2><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o);           \
    break;
            switch (C) {
                case 8: CDS_AGGH(8, 4)
                case 16: CDS_AGGH(16, 3)
                case 32: CDS_AGGH(32, 3)
                default: cds_set_error("cds_costvol_aggregate: C must be 8, 16 or 32 (got %d)", C); return CDS_EUNSUPPORTED;
            }
            return cds_check_launch("cds_costvol_aggregate");
        }
    }
#define CDS_AGG(c)                                                                                                 \
    if (V <= 4 && (occ == 2 || (occ == 0 && vol_hi))) aggregate_kernel<T, c, 4, 2><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o, vol_hi, vol_lo); \
    else if (V <= 4) aggregate_kernel<T, c, 4, 3><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o, vol_hi, vol_lo);   \
    else aggregate_kernel<T, c, kMaxViews, 2><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o, vol_hi, vol_lo);    \
    break;
    if constexpr (std::is_same<T, float>::value) {
        // stage 1 of the precise cascade: fp32 features, split fp16 volume, up to 4 views -> four channels per lane
        static const int quad = [] { const char* e = getenv("CDS_AGG_QUAD"); return e ? atoi(e) : 3; }();
        if (quad && vol_hi && !volume && V <= 4 && C == 32) {
            dim3 qb(cds_div_up((long long)h * w * (C / 4), 256), B);
            if (quad == 2) aggregate_f32q_kernel<32, 4, 2><<<qb, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, vol_hi, vol_lo);
            else if (quad == 4) aggregate_f32q_kernel<32, 4, 4><<<qb, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, vol_hi, vol_lo);
            else aggregate_f32q_kernel<32, 4, 3><<<qb, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, vol_hi, vol_lo);
            return cds_check_launch("cds_costvol_aggregate");
        }
    }
    switch (C) {
        case 8: CDS_AGG(8)
        case 16: CDS_AGG(16)
        case 32: CDS_AGG(32)
        default: cds_set_error("cds_costvol_aggregate: C must be 8, 16 or 32 (got %d)", C); return CDS_EUNSUPPORTED;
    }
    return cds_check_launch("cds_costvol_aggregate");
}

}  // namespace

extern "C" {

int cds_costvol_entropy(const void* ref_fea, const void* src_fea, const float* coef, const float* depth, int V, int B,
                        int C, int D, int h, int w, int dtype, float* entropy, cudaStream_t stream) {
    CDS_REQUIRE(ref_fea && src_fea && coef && depth && entropy, CDS_EARG, "cds_costvol_entropy: null pointer");
    CDS_REQUIRE(V >= 1 && V <= kMaxViews && B > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE,
                "cds_costvol_entropy: bad shape V=%d B=%d D=%d h=%d w=%d (1 <= V <= %d)", V, B, D, h, w, kMaxViews);
    CDS_REQUIRE((long long)h * w * C < (1ll << 31) && (long long)V * B * h * w < (1ll << 31) && (long long)V * B <= 65535, CDS_ESHAPE,
                "cds_costvol_entropy: feature map too large for 32-bit offsets");
    if (s3_tma_applies(V, B, C, D, h, w, dtype)) {
        S3Params p{(const __half*)ref_fea, (const __half*)src_fea, coef, depth, nullptr, entropy, nullptr, V, B, h, w};
        return launch_s3_tma<0>(p, stream);
    }
    if (dtype == CDS_F16) return launch_entropy<__half>(ref_fea, src_fea, coef, depth, V, B, C, D, h, w, entropy, stream);
    if (dtype == CDS_F32) return launch_entropy<float>(ref_fea, src_fea, coef, depth, V, B, C, D, h, w, entropy, stream);
    cds_set_error("cds_costvol_entropy: unknown dtype %d", dtype);
    return CDS_EARG;
}

int cds_costvol_entropy_fast(const void* ref_fea, const void* src_fea, const float* coef, const float* depth, int V, int B,
                             int C, int D, int h, int w, int dtype, float* entropy, cudaStream_t stream) {
    if (dtype != CDS_F16 || s3_tma_applies(V, B, C, D, h, w, dtype))
        return cds_costvol_entropy(ref_fea, src_fea, coef, depth, V, B, C, D, h, w, dtype, entropy, stream);
    CDS_REQUIRE(ref_fea && src_fea && coef && depth && entropy, CDS_EARG, "cds_costvol_entropy_fast: null pointer");
    CDS_REQUIRE(V >= 1 && V <= kMaxViews && B > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE,
                "cds_costvol_entropy_fast: bad shape V=%d B=%d D=%d h=%d w=%d (1 <= V <= %d)", V, B, D, h, w, kMaxViews);
    CDS_REQUIRE((long long)h * w * C < (1ll << 31) && (long long)V * B * h * w < (1ll << 31) && (long long)V * B <= 65535, CDS_ESHAPE,
                "cds_costvol_entropy_fast: feature map too large for 32-bit offsets");
    return launch_entropy<__half>(ref_fea, src_fea, coef, depth, V, B, C, D, h, w, entropy, stream, true);
}

int cds_costvol_aggregate(const void* ref_fea, const void* src_fea, const float* coef, const float* depth,
                          const float* vis, int V, int B, int C, int D, int h, int w, int dtype, void* volume,
                          cudaStream_t stream) {
    CDS_REQUIRE(ref_fea && src_fea && coef && depth && vis && volume, CDS_EARG, "cds_costvol_aggregate: null pointer");
    CDS_REQUIRE(V >= 1 && V <= kMaxViews && B > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE,
                "cds_costvol_aggregate: bad shape V=%d B=%d D=%d h=%d w=%d (1 <= V <= %d)", V, B, D, h, w, kMaxViews);
    CDS_REQUIRE((long long)h * w * C < (1ll << 31) && (long long)V * B * h * w < (1ll << 31) && B <= 65535, CDS_ESHAPE,
                "cds_costvol_aggregate: feature map too large for 32-bit offsets");
    if (s3_tma_applies(V, B, C, D, h, w, dtype)) {
        S3Params p{(const __half*)ref_fea, (const __half*)src_fea, coef, depth, vis, nullptr, (__half*)volume, V, B, h, w};
        return launch_s3_tma<1>(p, stream);
    }
    if (dtype == CDS_F16) return launch_aggregate<__half>(ref_fea, src_fea, coef, depth, vis, V, B, C, D, h, w, volume, stream);
    if (dtype == CDS_F32) return launch_aggregate<float>(ref_fea, src_fea, coef, depth, vis, V, B, C, D, h, w, volume, stream);
    cds_set_error("cds_costvol_aggregate: unknown dtype %d", dtype);
    return CDS_EARG;
}

// fp32 features in, split-precision fp16 volume out (value plane + optional rounding-residual plane)
int cds_costvol_aggregate_split(const float* ref_fea, const float* src_fea, const float* coef, const float* depth, const float* vis,
                                int V, int B, int C, int D, int h, int w, void* vol_hi, void* vol_lo, cudaStream_t stream) {
    CDS_REQUIRE(ref_fea && src_fea && coef && depth && vis && vol_hi, CDS_EARG, "cds_costvol_aggregate_split: null pointer");
    CDS_REQUIRE(V >= 1 && V <= kMaxViews && B > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE,
                "cds_costvol_aggregate_split: bad shape V=%d B=%d D=%d h=%d w=%d (1 <= V <= %d)", V, B, D, h, w, kMaxViews);
    CDS_REQUIRE((long long)h * w * C < (1ll << 31) && (long long)V * B * h * w < (1ll << 31) && B <= 65535, CDS_ESHAPE,
                "cds_costvol_aggregate_split: feature map too large for 32-bit offsets");
    return launch_aggregate<float>(ref_fea, src_fea, coef, depth, vis, V, B, C, D, h, w, nullptr, stream, (__half*)vol_hi, (__half*)vol_lo);
}

int cds_nc_mean(const float* ref_nc_sum, const float* src_nc_sum, int V, long long n, float* out, cudaStream_t stream) {
    CDS_REQUIRE(ref_nc_sum && src_nc_sum && out && V >= 1 && n > 0, CDS_EARG, "cds_nc_mean: bad arguments");
    nc_mean_kernel<<<cds_div_up(n, 256), 256, 0, stream>>>(ref_nc_sum, src_nc_sum, V, n, out);
    return cds_check_launch("cds_nc_mean");
}

}  // extern "C"
