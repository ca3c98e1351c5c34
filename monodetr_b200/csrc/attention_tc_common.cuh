// attention_tc_common.cuh -- device helpers shared by the tcgen05 attention kernels (attention_tc.cu forward,
// attention_tc_bwd.cu backward): fp32 tile rows -> packed bf16 (hi, lo) operand rows, transposed bf16 tiles, TMEM stores.
#pragma once
#include <stdint.h>

#include "tc_common.cuh"

namespace mdb {

__device__ __forceinline__ float ex2_approx(float x) {      // MUFU.EX2: 2^x, flushes denormals, ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// registers -> TMEM, 16 consecutive 32-bit columns of lane (lane_base + t)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// registers -> TMEM, 8 consecutive 32-bit columns of lane (lane_base + t)
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// fp32 row (128 B, SWIZZLE_128B) -> the same 128 bytes as [hi 32 | lo 32] packed bf16 (same swizzle), scaled by `mul`
__device__ __forceinline__ void split_row_in_place(uint8_t* row_ptr, int row, float mul) {
    uint32_t hw[16], lw[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 r = *reinterpret_cast<const uint4*>(row_ptr + ((c ^ (row & 7)) << 4));
        const float x0 = __uint_as_float(r.x) * mul, x1 = __uint_as_float(r.y) * mul, x2 = __uint_as_float(r.z) * mul,
                    x3 = __uint_as_float(r.w) * mul;
        const uint32_t h01 = pack_bf16x2(x0, x1), h23 = pack_bf16x2(x2, x3);
        hw[2 * c] = h01;
        hw[2 * c + 1] = h23;
        lw[2 * c] = pack_bf16x2(x0 - __uint_as_float(h01 << 16), x1 - __uint_as_float(h01 & 0xFFFF0000u));
        lw[2 * c + 1] = pack_bf16x2(x2 - __uint_as_float(h23 << 16), x3 - __uint_as_float(h23 & 0xFFFF0000u));
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<uint4*>(row_ptr + ((c ^ (row & 7)) << 4)) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
        *reinterpret_cast<uint4*>(row_ptr + (((4 + c) ^ (row & 7)) << 4)) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
    }
}


// One fp32 row (key / query `row` of a tile, 128 B, SWIZZLE_128B) -> column `row` of a K-major TRANSPOSED bf16 tile pair:
// tile[d][row] for d = 0..31, hi at `tt`, lo at `tt + lo_off`.  The transposed tile is [atoms of 64 columns][32 rows d][128 B]
// with SWIZZLE_128B inside each atom.  For a fixed d the 32 lanes of a warp write 32 consecutive columns = 4 swizzled
// 16-byte chunks: conflict-free.
__device__ __forceinline__ void transpose_row_bf16(const uint8_t* row_ptr, int row, uint8_t* tt, int lo_off) {
    uint8_t* base = tt + (row >> 6) * 4096 + (row & 7) * 2;
    const int kc = (row & 63) >> 3;                            // 16-byte chunk (8 columns) inside the atom row
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 r = *reinterpret_cast<const uint4*>(row_ptr + ((c ^ (row & 7)) << 4));
        const float x[4] = {__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int d = c * 4 + e;
            const uint32_t hb = pack_bf16x2(x[e], 0.f) & 0xFFFFu;
            const uint32_t lb = pack_bf16x2(x[e] - __uint_as_float(hb << 16), 0.f) & 0xFFFFu;
            uint8_t* dst = base + d * 128 + ((kc ^ (d & 7)) << 4);
            *reinterpret_cast<uint16_t*>(dst) = (uint16_t)hb;
            *reinterpret_cast<uint16_t*>(dst + lo_off) = (uint16_t)lb;
        }
    }
}

}  // namespace mdb
