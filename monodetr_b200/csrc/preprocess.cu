// preprocess.cu -- input pipeline on the device: the step before the hot path.  The reference's dataset
// (lib/datasets/kitti/kitti_dataset.py:121-163) warps every image on a data-loader worker with PIL
// (`img.transform(resolution, Image.AFFINE, trans_inv, resample=Image.BILINEAR)`), converts to float, normalises and transposes
// to CHW with numpy.  Here the decoded 8-bit images are uploaded as they are (ragged sizes, 3 bytes per pixel) and ONE kernel
// produces the normalised NCHW fp32 batch the backbone reads: optional left-right flip (:140-142), inverse affine map of every
// output pixel centre, PIL's bilinear filter (clamped neighbours, zero fill outside, result truncated to 8 bits), /255,
// (x - mean) / std (:159-161).
// Arithmetic: coordinates and interpolation in fp64 exactly as PIL's C code (Geometry.c) so that the 8-bit value is
// bit-identical; the float conversion in fp32 exactly as numpy's.  HBM-bound: 12 bytes written per output pixel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/monodetr_b200.h"

namespace {

__global__ void __launch_bounds__(256) warp_affine_normalize_kernel(const unsigned char* const* __restrict__ src, const int* __restrict__ src_wh,
                                                                    const long long* __restrict__ src_pitch, const double* __restrict__ trans_inv,
                                                                    const unsigned char* __restrict__ flip, float* __restrict__ out, int Wo,
                                                                    int Ho, float3 mean, float3 stdv) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= Wo) return;
    const int W = src_wh[2 * b], H = src_wh[2 * b + 1];
    const double* a = trans_inv + 6 * b;
    const double xc = x + 0.5, yc = y + 0.5;
    double xin = a[0] * xc + a[1] * yc + a[2];
    double yin = a[3] * xc + a[4] * yc + a[5];
    float v[3] = {0.f, 0.f, 0.f};
    if (!(xin < 0.0 || xin >= (double)W || yin < 0.0 || yin >= (double)H)) {
        xin -= 0.5; yin -= 0.5;
        const int xf = (int)floor(xin), yf = (int)floor(yin);
        const double dx = xin - xf, dy = yin - yf;
        int x0 = min(max(xf, 0), W - 1), x1 = min(max(xf + 1, 0), W - 1);
        const int y0 = min(max(yf, 0), H - 1);
        const bool has_y1 = yf + 1 >= 0 && yf + 1 < H;
        if (flip && flip[b]) { x0 = W - 1 - x0; x1 = W - 1 - x1; }        // sampling the mirrored image
        const unsigned char* r0 = src[b] + (long long)y0 * src_pitch[b];
        const unsigned char* r1 = src[b] + (long long)(yf + 1) * src_pitch[b];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double p00 = r0[x0 * 3 + c], p01 = r0[x1 * 3 + c];
            const double v1 = p00 + (p01 - p00) * dx;
            double v2 = v1;
            if (has_y1) {
                const double p10 = r1[x0 * 3 + c], p11 = r1[x1 * 3 + c];
                v2 = p10 + (p11 - p10) * dx;
            }
            v[c] = (float)(unsigned char)(v1 + (v2 - v1) * dy);
        }
    }
    const size_t plane = (size_t)Ho * Wo;
    float* o = out + (size_t)b * 3 * plane + (size_t)y * Wo + x;
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(v[0], 255.f), mean.x), stdv.x);
    o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(v[1], 255.f), mean.y), stdv.y);
    o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(v[2], 255.f), mean.z), stdv.z);
}

}  // namespace

extern "C" int mdb_warp_affine_normalize_u8(const unsigned char* const* src, const int* src_wh, const long long* src_pitch,
                                            const double* trans_inv, const unsigned char* flip, int B, int out_w, int out_h,
                                            const float* mean3, const float* std3, float* out, void* stream) {
    if (!src || !src_wh || !src_pitch || !trans_inv || !mean3 || !std3 || !out) return MDB_EINVAL;
    if (B <= 0 || out_w <= 0 || out_h <= 0 || out_h > 65535 || B > 65535) return MDB_EINVAL;
    dim3 grid((out_w + 255) / 256, out_h, B);
    warp_affine_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_wh, src_pitch, trans_inv, flip, out, out_w, out_h,
                                                                         make_float3(mean3[0], mean3[1], mean3[2]),
                                                                         make_float3(std3[0], std3[1], std3[2]));
    return (int)cudaGetLastError();
}
