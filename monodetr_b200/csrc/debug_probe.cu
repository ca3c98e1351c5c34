// debug_probe.cu -- micro-probes used while tuning (tools/probe_store.py); not on the product path.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
// Every CTA writes `bytes_per_cta` as 128x32-float patches (row pitch `ld` floats) with W warps, the store pattern
// of the conv_gemm epilogue (8 lanes x float4 = one 128-byte row segment, 4 rows per warp instruction).
__global__ void probe_store_kernel(float* out, int ld, int tiles_per_cta, int rows_total) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int r4 = lane >> 3, c4 = lane & 7;
    for (int t = 0; t < tiles_per_cta; ++t) {
        const long long tile = (long long)blockIdx.x * tiles_per_cta + t;      // 128 rows x 128 cols per tile
        const long long row0 = (tile * 128) % rows_total;
        for (int ch = warp; ch < 16; ch += nw) {                                // 16 (quarter, chunk) pairs per tile
            const int q = ch & 3, c0 = (ch >> 2) * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long row = row0 + q * 32 + i * 4 + r4;
                float4 v = make_float4((float)i, (float)t, (float)lane, 1.f);
                *reinterpret_cast<float4*>(out + row * ld + c0 + c4 * 4) = v;
            }
        }
    }
}
}  // namespace

extern "C" int mdb_debug_probe_store(float* out, int ld, int rows_total, int ctas, int warps, int tiles_per_cta, void* stream) {
    probe_store_kernel<<<ctas, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(out, ld, tiles_per_cta, rows_total);
    return (int)cudaGetLastError();
}
