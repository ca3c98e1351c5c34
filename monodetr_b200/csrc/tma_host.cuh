// tma_host.cuh -- host-side tensor-map (TMA descriptor) encoder shared by the tensor-core kernels' launchers.
// cuTensorMapEncodeTiled is fetched through the runtime's driver-entry-point query, so the library links against the CUDA
// runtime only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

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
// Encoded maps are kept in a small per-thread direct-mapped cache keyed by every argument: the eager (non-graph) API re-launches
// the same ~700 (pointer, shape) combinations every step -- torch's caching allocator hands the same addresses back -- and a hit
// costs one hash + one 128-byte compare instead of a driver call.  (A map only describes addresses and strides; it is valid for
// whatever data lives there.)
struct MapKey {
    const void* base;
    uint64_t dims[5], strides[5];
    uint32_t box[5], estr[5];
    int rank, flags;
};
struct MapCacheEntry {
    MapKey key;
    CUtensorMap map;
    bool valid;
};
inline int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box, const uint32_t* estr, bool mn_major = false, bool bf16 = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return MDB_EUNSUPPORTED;
    constexpr int kSlots = 1024;
    static thread_local MapCacheEntry cache[kSlots];
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.base = base; key.rank = rank; key.flags = (mn_major ? 1 : 0) | (bf16 ? 2 : 0);
    uint64_t h = 1469598103934665603ull ^ (uint64_t)(uintptr_t)base;
    for (int i = 0; i < rank; ++i) {
        key.dims[i] = dims[i];
        key.strides[i] = i ? strides_elems[i] : 1;
        key.box[i] = box[i];
        key.estr[i] = estr ? estr[i] : 1;
        h = (h ^ dims[i]) * 1099511628211ull;
        h = (h ^ key.strides[i]) * 1099511628211ull;
        h = (h ^ (((uint64_t)key.box[i] << 32) | key.estr[i])) * 1099511628211ull;
    }
    h = (h ^ (uint64_t)(rank * 4 + key.flags)) * 1099511628211ull;
    MapCacheEntry& e = cache[(h >> 20) % kSlots];
    if (e.valid && memcmp(&e.key, &key, sizeof(key)) == 0) {
        *m = e.map;
        return 0;
    }
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
                     const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MDB_EINVAL;
    e.key = key;
    e.map = *m;
    e.valid = true;
    return 0;
}

}  // namespace mdb
