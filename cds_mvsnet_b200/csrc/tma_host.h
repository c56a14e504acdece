// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled resolved at run time through the CUDA
// runtime, so the library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

void cds_set_error(const char* fmt, ...);

namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// Tensor of `rank` dims (fastest first), no swizzle/interleave, zero fill out of bounds.
// dims[i] elements, strides_bytes[i] for i >= 1 (multiples of 16), box[i] elements (inner box bytes multiple of 16,
// every box dim <= 256).  make_u64 views the data as 8-byte elements (half a 16-byte 8-channel voxel slab) so that a
// box row can be 2 KB = 128 voxels.
inline bool make_map(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { cds_set_error("cuTensorMapEncodeTiled is not available from this driver"); return false; }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 1; i < rank; ++i) gs[i - 1] = strides_bytes[i];
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { cds_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return false; }
    return true;
}

inline bool make_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    return make_map(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box);
}
inline bool make_u64(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    return make_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, base, rank, dims, strides_bytes, box);
}

}  // namespace tma
