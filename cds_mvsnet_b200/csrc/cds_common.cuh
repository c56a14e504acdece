// Shared device helpers for the cds_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CDS_F32 0
#define CDS_F16 1

// status codes returned through the C ABI (0 = ok, >0 = cudaError_t, <0 = argument check)
#define CDS_OK 0
#define CDS_EARG (-1)
#define CDS_ESHAPE (-2)
#define CDS_EUNSUPPORTED (-3)

void cds_set_error(const char* fmt, ...);
int cds_check_launch(const char* what);

#define CDS_REQUIRE(cond, code, ...)   \
    do {                               \
        if (!(cond)) {                 \
            cds_set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

static inline int cds_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- 8-channel vector access: activations are channels-last with C % 8 == 0 -----------------
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
    __device__ static __forceinline__ void load(const float* p, float (&v)[8]) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p));
        float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ static __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <>
struct Vec8<__half> {
    __device__ static __forceinline__ void load(const __half* p, float (&v)[8]) {
        uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    __device__ static __forceinline__ void store(__half* p, const float (&v)[8]) {
        uint4 r;
        __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }

template <typename T>
__device__ __forceinline__ T from_f32(float x);
template <>
__device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(x); }

// ---- plane-sweep projection: p = R (x,y,1)^T * depth + t ; (u,v) = p.xy / (p.z + 1e-6) ---------
// (reference: models/utils/warping.py:90-94).  coef = {R row-major 9, t 3}.
struct WarpCoef {
    float r[9];
    float t[3];
};
__device__ __forceinline__ WarpCoef load_coef(const float* c) {
    WarpCoef k;
#pragma unroll
    for (int i = 0; i < 9; ++i) k.r[i] = __ldg(c + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) k.t[i] = __ldg(c + 9 + i);
    return k;
}
// ray = R (x,y,1)^T, evaluated exactly like the reference's matmul over (x, y, 1)
__device__ __forceinline__ void pixel_ray(const WarpCoef& k, float x, float y, float& rx, float& ry, float& rz) {
    rx = k.r[0] * x + k.r[1] * y + k.r[2];
    ry = k.r[3] * x + k.r[4] * y + k.r[5];
    rz = k.r[6] * x + k.r[7] * y + k.r[8];
}
__device__ __forceinline__ void project(const WarpCoef& k, float rx, float ry, float rz, float depth, float& u, float& v) {
    float px = rx * depth + k.t[0];
    float py = ry * depth + k.t[1];
    float pz = rz * depth + k.t[2] + 1e-6f;
    u = px / pz;
    v = py / pz;
}
// same projection with one correctly-rounded reciprocal instead of two divisions (fused cost-volume kernels:
// the difference is <= 1 ulp of the coordinate, ~1e-4 px at w = 1600)
__device__ __forceinline__ void project_fast(const WarpCoef& k, float rx, float ry, float rz, float depth, float& u, float& v) {
    float px = rx * depth + k.t[0];
    float py = ry * depth + k.t[1];
    float inv = __frcp_rn(rz * depth + k.t[2] + 1e-6f);
    u = px * inv;
    v = py * inv;
}

// Bilinear footprint at pixel coords (u,v) with zero padding (grid_sample bilinear/zeros/
// align_corners=True after the reference's normalisation, warping.py:95-101).
struct Taps {
    int x0, y0;
    float w00, w01, w10, w11;  // weights of (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1); 0 when outside
};
__device__ __forceinline__ Taps make_taps(float u, float v, int w, int h) {
    Taps t;
    // NaN / huge coordinates: everything is outside
    if (!(u > -2.f && u < (float)(w + 1) && v > -2.f && v < (float)(h + 1))) {
        t.x0 = 0; t.y0 = 0; t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
        return t;
    }
    float xf = floorf(u), yf = floorf(v);
    float fx = u - xf, fy = v - yf;
    int x0 = (int)xf, y0 = (int)yf;
    bool xa = (x0 >= 0) & (x0 < w), xb = (x0 + 1 >= 0) & (x0 + 1 < w);
    bool ya = (y0 >= 0) & (y0 < h), yb = (y0 + 1 >= 0) & (y0 + 1 < h);
    t.w00 = (xa && ya) ? (1.f - fx) * (1.f - fy) : 0.f;
    t.w01 = (xb && ya) ? fx * (1.f - fy) : 0.f;
    t.w10 = (xa && yb) ? (1.f - fx) * fy : 0.f;
    t.w11 = (xb && yb) ? fx * fy : 0.f;
    t.x0 = x0;
    t.y0 = y0;
    return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
