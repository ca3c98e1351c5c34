// tma_host.cuh -- host-side tensor-map (TMA descriptor) encoder shared by the tensor-core kernels' launchers.
// cuTensorMapEncodeTiled is fetched through the runtime's driver-entry-point query, so the library links against the CUDA
// runtime only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/monodetr_b200.h"

namespace mdb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}

// dims/strides innermost first; strides in ELEMENTS for dims 1..rank-1 (dim 0 is contiguous).
inline int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
             const uint32_t* box, const uint32_t* estr, bool mn_major = false, bool bf16 = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return MDB_EUNSUPPORTED;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = estr ? estr[i] : 1;
    }
    for (int i = 1; i < rank; ++i) {
        gstr[i - 1] = strides_elems[i] * (bf16 ? 2 : sizeof(float));
        if (gstr[i - 1] % 16) return MDB_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(base) % 16) return MDB_EINVAL;
    CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                     const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : MDB_EINVAL;
}

}  // namespace mdb
