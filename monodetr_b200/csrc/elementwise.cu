// elementwise.cu -- small HBM-bound helper kernels around the tensor-core convolutions:
// weight packing OIHW -> [tap][O][I] (with the FrozenBatchNorm scale folded in, backbone.py:54-64),
// the inverse for weight gradients, and column sums (bias gradients).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/monodetr_b200.h"

namespace {

__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, float* __restrict__ out,
                                   int O, int I, int taps) {
    const long long n = (long long)O * I * taps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT [tap][o][ci] so writes are coalesced
        const int ci = (int)(i % I);
        const long long r = i / I;
        const int o = (int)(r % O);
        const int t = (int)(r / O);
        float v = w[((size_t)o * I + ci) * taps + t];
        if (scale) v *= scale[o];
        out[i] = v;
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int O, int I, int taps,
                                    int accumulate) {
    const long long n = (long long)O * I * taps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT [o][ci][tap]
        const int t = (int)(i % taps);
        const long long r = i / taps;
        const int ci = (int)(r % I);
        const int o = (int)(r / I);
        const float v = dwp[((size_t)t * O + o) * I + ci];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

// out[n] += sum_m x[m][n]; grid.x covers column groups of 32, grid.y row slabs.
__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long M, int N, int rows_per_block) {
    __shared__ float part[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;   // 8 row lanes
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(M, r0 + rows_per_block);
    float acc = 0.f;
    if (col < N)
        for (long long r = r0 + ty; r < r1; r += 8) acc += x[r * N + col];
    part[ty][threadIdx.x & 31] = acc;
    __syncthreads();
    if (ty == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x];
        atomicAdd(out + col, s);
    }
}

}  // namespace

extern "C" {

int mdb_pack_conv_weight_f32(const float* w_oihw, const float* scale, float* w_packed, int O, int I, int taps,
                             void* stream) {
    if (!w_oihw || !w_packed || O <= 0 || I <= 0 || taps <= 0) return MDB_EINVAL;
    const long long n = (long long)O * I * taps;
    const int grid = (int)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
    pack_weight_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w_oihw, scale, w_packed, O, I, taps);
    return (int)cudaGetLastError();
}

int mdb_unpack_conv_wgrad_f32(const float* dw_packed, float* dw_oihw, int O, int I, int taps, int accumulate,
                              void* stream) {
    if (!dw_packed || !dw_oihw || O <= 0 || I <= 0 || taps <= 0) return MDB_EINVAL;
    const long long n = (long long)O * I * taps;
    const int grid = (int)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
    unpack_wgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dw_packed, dw_oihw, O, I, taps, accumulate);
    return (int)cudaGetLastError();
}

int mdb_colsum_f32(const float* x, float* out, long long M, int N, int accumulate, void* stream_) {
    if (!x || !out || M < 0 || N <= 0) return MDB_EINVAL;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * N, stream);
        if (e != cudaSuccess) return (int)e;
    }
    if (M == 0) return 0;
    const int gx = (N + 31) / 32;
    int gy = (int)((M + 511) / 512);
    const int cap = (148 * 8 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    const int rows = (int)((M + gy - 1) / gy);
    colsum_kernel<<<dim3(gx, gy), 256, 0, stream>>>(x, out, M, N, rows);
    return (int)cudaGetLastError();
}

}  // extern "C"
