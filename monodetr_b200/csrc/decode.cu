// decode.cu -- inference post-process on the device: the step after MonoDETR.forward in the reference's
// tester (lib/helpers/tester_helper.py:85-100), which moves the raw head outputs to the host and loops over detections in
// Python.  Two kernels, one CTA per image, no host synchronisation:
//   * extract   lib/helpers/decode_helper.py:57-110   sigmoid -> top-k over (query, class) -> gather of the heads -> (B, topk, 37)
//   * decode    lib/helpers/decode_helper.py:8-54     score threshold, 2-d box in pixels, 3-d size / location / heading
//                                                      through the camera matrix -> (B, topk, 14) rows + a count per image
// Latency-bound (B*Q*C = a few hundred candidates per image); fp32.  Parity: tests/test_decode_gpu.py against
// oracle/decode.py and the golden vectors generated from the reference functions.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/monodetr_b200.h"

namespace {

constexpr int kMaxCand = 4096;      // Q * C candidates per image sorted in shared memory (eval: 50 * 3, train: 550 * 3)
constexpr int kDetCols = 37;        // label score xs2d ys2d w h depth heading[24] size3d[3] xs3d ys3d sigma
constexpr int kOutCols = 14;        // cls alpha x0 y0 x1 y1 h w l X Y Z ry score
constexpr int kBins = 12;           // lib/datasets/utils.py:7  num_heading_bin

// Descending by logit (sigmoid is monotonic, so the order is the order of the probabilities the reference sorts);
// equal keys keep the lower flat index first.
__device__ __forceinline__ bool before(float ka, int ia, float kb, int ib) { return ka > kb || (ka == kb && ia < ib); }

__global__ void __launch_bounds__(256) extract_dets_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                                           const float* __restrict__ dim3, const float* __restrict__ depth,
                                                           const float* __restrict__ angle, int Q, int C, int topk, int n2,
                                                           float* __restrict__ dets) {
    extern __shared__ unsigned char smem_raw[];
    float* key = reinterpret_cast<float*>(smem_raw);
    int* idx = reinterpret_cast<int*>(key + n2);
    const int b = blockIdx.x, n = Q * C;
    const float* lg = logits + (size_t)b * n;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        float v = -INFINITY;
        if (i < n) { v = lg[i]; if (v != v) v = -INFINITY; }      // NaN logits sort last
        key[i] = v;
        idx[i] = i < n ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;                  // this run ends "best first"
                    const float ka = key[i], kb = key[p];
                    const int ia = idx[i], ib = idx[p];
                    if (before(kb, ib, ka, ia) == up) { key[i] = kb; key[p] = ka; idx[i] = ib; idx[p] = ia; }
                }
            }
            __syncthreads();
        }
    for (int t = threadIdx.x; t < topk; t += blockDim.x) {
        float* o = dets + ((size_t)b * topk + t) * kDetCols;
        const int flat = idx[t];
        const int q = flat / C, c = flat - q * C;
        const float* bx = boxes + ((size_t)b * Q + q) * 6;
        const float cx = bx[0], cy = bx[1], l = bx[2], r = bx[3], tp = bx[4], bt = bx[5];
        const float x0 = cx - l, y0 = cy - tp, x1 = cx + r, y1 = cy + bt;              // utils/box_ops.py:20-24
        o[0] = (float)c;
        o[1] = 1.f / (1.f + expf(-key[t]));
        o[2] = (x0 + x1) / 2.f; o[3] = (y0 + y1) / 2.f; o[4] = x1 - x0; o[5] = y1 - y0;   // utils/box_ops.py:27-31
        const float* dp = depth + ((size_t)b * Q + q) * 2;
        o[6] = dp[0];
        const float* an = angle + ((size_t)b * Q + q) * (2 * kBins);
#pragma unroll
        for (int k = 0; k < 2 * kBins; ++k) o[7 + k] = an[k];
        const float* d3 = dim3 + ((size_t)b * Q + q) * 3;
        o[31] = d3[0]; o[32] = d3[1]; o[33] = d3[2];
        o[34] = cx; o[35] = cy;
        o[36] = expf(-dp[1]);
    }
}

// One thread per detection row; the scores of a row of `dets` are sorted, so the rows that pass the threshold are a prefix
// and the output needs no compaction: count[b] = length of that prefix, rows beyond it are zero-filled.
__global__ void decode_dets_kernel(const float* __restrict__ dets, const float* __restrict__ img_size, const float* __restrict__ P2,
                                   const float* __restrict__ mean_size, int topk, int C, float threshold, float* __restrict__ out,
                                   int* __restrict__ count) {
    const int b = blockIdx.x;
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const float W = img_size[b * 2 + 0], H = img_size[b * 2 + 1];
    const float* P = P2 + (size_t)b * 12;
    const float fu = P[0], cu = P[2], fv = P[5], cv = P[6];               // kitti_utils.py:150-155
    const float tx = P[3] / (-fu), ty = P[7] / (-fv);
    int mine = 0;
    for (int t = threadIdx.x; t < topk; t += blockDim.x) {
        const float* d = dets + ((size_t)b * topk + t) * kDetCols;
        float* o = out + ((size_t)b * topk + t) * kOutCols;
        const float score = d[1];
        if (!(score >= threshold)) {                                       // decode_helper.py:22  `if score < threshold: continue`
#pragma unroll
            for (int k = 0; k < kOutCols; ++k) o[k] = 0.f;
            continue;
        }
        ++mine;
        int cls = (int)d[0];
        cls = cls < 0 ? 0 : (cls >= C ? C - 1 : cls);
        const float x = d[2] * W, y = d[3] * H, w = d[4] * W, h = d[5] * H;
        const float depth = d[6];
        const float dh = d[31] + mean_size[cls * 3 + 0], dw = d[32] + mean_size[cls * 3 + 1], dl = d[33] + mean_size[cls * 3 + 2];
        const float x3d = d[34] * W, y3d = d[35] * H;
        const float X = ((x3d - cu) * depth) / fu + tx;                    // kitti_utils.py:207-208
        const float Y = ((y3d - cv) * depth) / fv + ty + dh / 2.f;         // decode_helper.py:45
        int best = 0;                                                      // decode_helper.py:174-178, first maximum as np.argmax
        float bv = d[7];
#pragma unroll
        for (int k = 1; k < kBins; ++k) if (d[7 + k] > bv) { bv = d[7 + k]; best = k; }
        const float kPi = 3.14159265358979323846f;
        float alpha = (float)best * (2.f * kPi / (float)kBins) + d[7 + kBins + best];   // lib/datasets/utils.py:19-26
        if (alpha > kPi) alpha -= 2.f * kPi;
        float ry = alpha + atan2f(x - cu, fu);                             // kitti_utils.py:277-282
        if (ry > kPi) ry -= 2.f * kPi;
        if (ry < -kPi) ry += 2.f * kPi;
        o[0] = (float)cls; o[1] = alpha;
        o[2] = x - w / 2.f; o[3] = y - h / 2.f; o[4] = x + w / 2.f; o[5] = y + h / 2.f;
        o[6] = dh; o[7] = dw; o[8] = dl;
        o[9] = X; o[10] = Y; o[11] = depth;
        o[12] = ry;
        o[13] = score * d[36];
    }
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) count[b] = s_count;
}

}  // namespace

extern "C" int mdb_extract_dets_f32(const float* logits, const float* boxes, const float* dim3, const float* depth, const float* angle,
                                    int B, int Q, int C, int topk, float* dets, void* stream) {
    if (!logits || !boxes || !dim3 || !depth || !angle || !dets) return MDB_EINVAL;
    if (B < 0 || Q <= 0 || C <= 0 || topk <= 0 || (long long)Q * C < topk) return MDB_EINVAL;
    if ((long long)Q * C > kMaxCand) return MDB_EUNSUPPORTED;
    if (B == 0) return 0;
    int n2 = 32;
    while (n2 < Q * C) n2 <<= 1;
    extract_dets_kernel<<<B, 256, (size_t)n2 * 8, (cudaStream_t)stream>>>(logits, boxes, dim3, depth, angle, Q, C, topk, n2, dets);
    return (int)cudaGetLastError();
}

extern "C" int mdb_decode_dets_f32(const float* dets, const float* img_size, const float* P2, const float* cls_mean_size, int B,
                                   int topk, int C, float threshold, float* out, int* count, void* stream) {
    if (!dets || !img_size || !P2 || !cls_mean_size || !out || !count) return MDB_EINVAL;
    if (B < 0 || topk <= 0 || C <= 0) return MDB_EINVAL;
    if (B == 0) return 0;
    decode_dets_kernel<<<B, 64, 0, (cudaStream_t)stream>>>(dets, img_size, P2, cls_mean_size, topk, C, threshold, out, count);
    return (int)cudaGetLastError();
}
