/*
 * oracle/msda_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Scalar CPU restatement of the multi-scale deformable attention sampling op
 * (forward gather + backward scatter), used as the checker for the sm_100a kernels in
 * monodetr_b200/csrc/msda.cu.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library.
 *
 * Semantics follow the reference CUDA kernels (citations into /root/reference):
 *   forward  : lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299 (index walk),
 *              :33-84 (4-corner bilinear read, zero outside)
 *   backward : same file :301-403 (per (b,q,m) block), :87-159 (corner scatter + d/dloc, d/dattn)
 *   host     : lib/models/monodetr/ops/src/cuda/ms_deform_attn_cuda.cu:20-153 (zero-initialised outputs)
 * and equal F.grid_sample(bilinear, zeros, align_corners=False) on grid 2*loc-1
 * (lib/models/monodetr/ops/functions/ms_deform_attn_func.py:41-61).
 *
 * Pinned by tests/test_oracle_msda.py against fixtures generated from the reference's own
 * ms_deform_attn_core_pytorch (tools/gen_golden_msda.py -> tests/golden/msda_*.npz).
 *
 * Image coordinate of a sample is defined as  fma(loc, size, -0.5)  in the working precision
 * (the reference's `loc * size - 0.5` is FMA-contracted by nvcc); floor() of it picks the cell.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_MSDA(T, SUF, FMA, FLOOR)                                                          \
void oracle_msda_fwd_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi,            \
                           const T* loc, const T* attn, int B, int S, int M, int D, int L,       \
                           int Lq, int P, T* out)                                                \
{                                                                                                \
    for (int b = 0; b < B; ++b)                                                                  \
    for (int q = 0; q < Lq; ++q)                                                                 \
    for (int m = 0; m < M; ++m) {                                                                \
        const size_t unit = ((size_t)b * Lq + q) * M + m;                                        \
        T* o = out + unit * D;                                                                   \
        for (int c = 0; c < D; ++c) o[c] = (T)0;                                                 \
        for (int l = 0; l < L; ++l) {                                                            \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                        \
            const T* vl = value + ((size_t)b * S + (size_t)lsi[l]) * M * D;                      \
            for (int p = 0; p < P; ++p) {                                                        \
                const size_t pi = (unit * L + l) * P + p;                                        \
                const T a = attn[pi];                                                            \
                const T x = FMA(loc[2 * pi], (T)W, (T)-0.5);                                     \
                const T y = FMA(loc[2 * pi + 1], (T)H, (T)-0.5);                                 \
                if (!(y > (T)-1 && x > (T)-1 && y < (T)H && x < (T)W)) continue;                 \
                const int y0 = (int)FLOOR(y), x0 = (int)FLOOR(x);                                \
                const T ly = y - (T)y0, lx = x - (T)x0, hy = (T)1 - ly, hx = (T)1 - lx;          \
                const T w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;                  \
                for (int c = 0; c < D; ++c) {                                                    \
                    T v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                            \
                    if (y0 >= 0 && x0 >= 0)          v1 = vl[((size_t)(y0 * W + x0) * M + m) * D + c];           \
                    if (y0 >= 0 && x0 + 1 <= W - 1)  v2 = vl[((size_t)(y0 * W + x0 + 1) * M + m) * D + c];       \
                    if (y0 + 1 <= H - 1 && x0 >= 0)  v3 = vl[((size_t)((y0 + 1) * W + x0) * M + m) * D + c];     \
                    if (y0 + 1 <= H - 1 && x0 + 1 <= W - 1)                                      \
                                                     v4 = vl[((size_t)((y0 + 1) * W + x0 + 1) * M + m) * D + c]; \
                    o[c] += a * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);                         \
                }                                                                                \
            }                                                                                    \
        }                                                                                        \
    }                                                                                            \
}                                                                                                \
                                                                                                 \
void oracle_msda_bwd_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi,            \
                           const T* loc, const T* attn, const T* grad_out, int B, int S, int M,  \
                           int D, int L, int Lq, int P, T* grad_value, T* grad_loc, T* grad_attn)\
{                                                                                                \
    memset(grad_value, 0, sizeof(T) * (size_t)B * S * M * D);                                    \
    memset(grad_loc, 0, sizeof(T) * (size_t)B * Lq * M * L * P * 2);                             \
    memset(grad_attn, 0, sizeof(T) * (size_t)B * Lq * M * L * P);                                \
    for (int b = 0; b < B; ++b)                                                                  \
    for (int q = 0; q < Lq; ++q)                                                                 \
    for (int m = 0; m < M; ++m) {                                                                \
        const size_t unit = ((size_t)b * Lq + q) * M + m;                                        \
        const T* g = grad_out + unit * D;                                                        \
        for (int l = 0; l < L; ++l) {                                                            \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                        \
            const size_t lbase = ((size_t)b * S + (size_t)lsi[l]) * M * D;                       \
            const T* vl = value + lbase;                                                         \
            T* gvl = grad_value + lbase;                                                         \
            for (int p = 0; p < P; ++p) {                                                        \
                const size_t pi = (unit * L + l) * P + p;                                        \
                const T a = attn[pi];                                                            \
                const T x = FMA(loc[2 * pi], (T)W, (T)-0.5);                                     \
                const T y = FMA(loc[2 * pi + 1], (T)H, (T)-0.5);                                 \
                if (!(y > (T)-1 && x > (T)-1 && y < (T)H && x < (T)W)) continue;                 \
                const int y0 = (int)FLOOR(y), x0 = (int)FLOOR(x);                                \
                const T ly = y - (T)y0, lx = x - (T)x0, hy = (T)1 - ly, hx = (T)1 - lx;          \
                const T w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;                  \
                const int ok1 = (y0 >= 0 && x0 >= 0), ok2 = (y0 >= 0 && x0 + 1 <= W - 1);        \
                const int ok3 = (y0 + 1 <= H - 1 && x0 >= 0);                                    \
                const int ok4 = (y0 + 1 <= H - 1 && x0 + 1 <= W - 1);                            \
                const size_t o1 = ((size_t)(y0 * W + x0) * M + m) * D;                           \
                const size_t o2 = o1 + (size_t)M * D;                                            \
                const size_t o3 = o1 + (size_t)W * M * D;                                        \
                const size_t o4 = o3 + (size_t)M * D;                                            \
                T ga = 0, gx = 0, gy = 0;                                                        \
                for (int c = 0; c < D; ++c) {                                                    \
                    const T tg = g[c] * a;                                                       \
                    T v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                            \
                    if (ok1) { v1 = vl[o1 + c]; gvl[o1 + c] += w1 * tg; }                        \
                    if (ok2) { v2 = vl[o2 + c]; gvl[o2 + c] += w2 * tg; }                        \
                    if (ok3) { v3 = vl[o3 + c]; gvl[o3 + c] += w3 * tg; }                        \
                    if (ok4) { v4 = vl[o4 + c]; gvl[o4 + c] += w4 * tg; }                        \
                    ga += g[c] * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);                        \
                    gx += tg * (-hy * v1 + hy * v2 - ly * v3 + ly * v4);                         \
                    gy += tg * (-hx * v1 - lx * v2 + hx * v3 + lx * v4);                         \
                }                                                                                \
                grad_attn[pi] = ga;                                                              \
                grad_loc[2 * pi] = (T)W * gx;                                                    \
                grad_loc[2 * pi + 1] = (T)H * gy;                                                \
            }                                                                                    \
        }                                                                                        \
    }                                                                                            \
}

DEFINE_MSDA(float, f32, fmaf, floorf)
DEFINE_MSDA(double, f64, fma, floor)
