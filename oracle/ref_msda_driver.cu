// oracle/ref_msda_driver.cu -- TEST / BENCH INFRASTRUCTURE ONLY.
//
// C-ABI driver around the REFERENCE's own CUDA kernels, compiled from the sources where they lie
// (#include of /root/reference/lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh through
// -I; nothing is copied into this repo).  The three headers that file includes but does not use
// (ATen/ATen.h, ATen/cuda/CUDAContext.h, THC/THCAtomics.cuh) are satisfied by empty shims in
// oracle/shim/.  Host chunking by im2col_step follows ms_deform_attn_cuda.cu:50-76, 117-148.
// Output: oracle/_ref/libref_msda.so (git-ignored, travels to the GPU box).  Used as the GPU-side
// oracle and as the "kernel to beat" line of bench.py; never linked into libmonodetr_b200.so.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ms_deform_im2col_cuda.cuh"

namespace {
template <typename T>
int fwd(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc, const T* attn, int B, int S,
        int M, int D, int L, int Lq, int P, T* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int step = B < 64 ? B : 64;
    if (B == 0) return 0;
    if (B % step) return -1;
    for (int n = 0; n < B / step; ++n) {
        ms_deformable_im2col_cuda<T>(stream, value + (size_t)n * step * S * M * D, shapes, lsi,
                                     loc + (size_t)n * step * Lq * M * L * P * 2,
                                     attn + (size_t)n * step * Lq * M * L * P, step, S, M, D, L, Lq, P,
                                     out + (size_t)n * step * Lq * M * D);
    }
    return (int)cudaGetLastError();
}
template <typename T>
int bwd(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc, const T* attn, const T* grad_out,
        int B, int S, int M, int D, int L, int Lq, int P, T* gv, T* gl, T* ga, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    // the reference allocates all three with at::zeros (ms_deform_attn_cuda.cu:121-123)
    cudaMemsetAsync(gv, 0, sizeof(T) * (size_t)B * S * M * D, stream);
    cudaMemsetAsync(gl, 0, sizeof(T) * (size_t)B * Lq * M * L * P * 2, stream);
    cudaMemsetAsync(ga, 0, sizeof(T) * (size_t)B * Lq * M * L * P, stream);
    const int step = B < 64 ? B : 64;
    if (B == 0) return 0;
    if (B % step) return -1;
    for (int n = 0; n < B / step; ++n) {
        ms_deformable_col2im_cuda<T>(stream, grad_out + (size_t)n * step * Lq * M * D,
                                     value + (size_t)n * step * S * M * D, shapes, lsi,
                                     loc + (size_t)n * step * Lq * M * L * P * 2,
                                     attn + (size_t)n * step * Lq * M * L * P, step, S, M, D, L, Lq, P,
                                     gv + (size_t)n * step * S * M * D, gl + (size_t)n * step * Lq * M * L * P * 2,
                                     ga + (size_t)n * step * Lq * M * L * P);
    }
    return (int)cudaGetLastError();
}
}  // namespace

extern "C" {
int ref_msda_forward_f32(const float* v, const int64_t* sh, const int64_t* ls, const float* lo, const float* at,
                         int B, int S, int M, int D, int L, int Lq, int P, float* out, void* st) {
    return fwd<float>(v, sh, ls, lo, at, B, S, M, D, L, Lq, P, out, st);
}
int ref_msda_forward_f64(const double* v, const int64_t* sh, const int64_t* ls, const double* lo, const double* at,
                         int B, int S, int M, int D, int L, int Lq, int P, double* out, void* st) {
    return fwd<double>(v, sh, ls, lo, at, B, S, M, D, L, Lq, P, out, st);
}
int ref_msda_backward_f32(const float* v, const int64_t* sh, const int64_t* ls, const float* lo, const float* at,
                          const float* go, int B, int S, int M, int D, int L, int Lq, int P, float* gv, float* gl,
                          float* ga, void* st) {
    return bwd<float>(v, sh, ls, lo, at, go, B, S, M, D, L, Lq, P, gv, gl, ga, st);
}
int ref_msda_backward_f64(const double* v, const int64_t* sh, const int64_t* ls, const double* lo, const double* at,
                          const double* go, int B, int S, int M, int D, int L, int Lq, int P, double* gv, double* gl,
                          double* ga, void* st) {
    return bwd<double>(v, sh, ls, lo, at, go, B, S, M, D, L, Lq, P, gv, gl, ga, st);
}
}
