// attention.cu -- fused multi-head attention core (softmax(Q K^T / sqrt(d)) V, head dim 32) for sm_100a,
// forward and backward, flash-style: the (Lq x Lk) probability matrix never touches HBM (the reference's
// nn.MultiheadAttention materialises it: 118 MB per image in the depth encoder).
// Replaces F.multi_head_attention_forward's core at depthaware_transformer.py:456-459 (depth cross-attn),
// :496 (group self-attn) and depth_predictor/transformer.py:59 (depth encoder).  The in/out projections
// are separate tensor-core GEMMs (conv_gemm.cu).
//
// Layout: q[b][i][h][32] with token stride ldq floats (so a packed QKV buffer can be passed), same for
// k, v (ldk, ldv), out[b][i][h*32] contiguous-by-token with stride ldo.  key_padding_mask[b][j] (uint8,
// nonzero = ignore) or null.  Dropout on the probabilities uses a counter-based hash RNG keyed by
// (seed, b, h, i, j) so the backward pass regenerates the same mask.
//
// v1 mapping (CUDA cores): 4 threads per query split the keys of each 64-key shared-memory tile,
// each keeps q, an output accumulator and running (max, sum) in registers; partial results are merged
// with shuffles.  Backward = a dQ kernel (same mapping) + a dK/dV kernel (4 threads per key over query tiles).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/monodetr_b200.h"
#include "rng.cuh"

namespace {

constexpr int HD = 32;         // head dim
constexpr int TK = 64;         // tile of keys (fwd, dq) or queries (dkdv) staged in smem
constexpr int ATT_THREADS = 128;
constexpr int ROWS_PER_CTA = ATT_THREADS / 4;   // 32 queries (or keys) per CTA

struct AttnParams {
    const float *q, *k, *v;
    const uint8_t* kpm;        // [B][Lk] or null
    float* out;
    float* lse;                // [B][H][Lq]
    int B, H, Lq, Lk, ldq, ldk, ldv, ldo;
    float scale;               // 1/sqrt(d)
    float drop_p;
    const unsigned long long* seed;   // device pointer (graph-safe), may be null when drop_p == 0
    unsigned long long site;
    // backward
    const float *dout, *o;     // dout / o with stride ldo
    float* delta;              // [B][H][Lq]  (dO . O)
    float *dq, *dk, *dv;       // strides lddq, lddk, lddv
    int lddq, lddk, lddv;
};

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

__device__ __forceinline__ bool keep_elem(const AttnParams& p, unsigned long long seed, int b, int h, int i, int j) {
    const unsigned long long idx = (((unsigned long long)(b * p.H + h) * p.Lq + i) * (unsigned long long)p.Lk + j);
    return mdb::rng_uniform(seed, idx) >= p.drop_p;
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnParams p) {
    __shared__ __align__(16) float sK[TK][HD];
    __shared__ __align__(16) float sV[TK][HD];
    const int b = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 2);
    const int kq = threadIdx.x & 3;
    const bool qok = qi < p.Lq;
    const unsigned long long seed = (p.drop_p > 0.f) ? (*p.seed + p.site * 0x9E3779B97F4A7C15ull) : 0ull;
    const float inv_keep = 1.f / (1.f - p.drop_p);

    float q[HD], acc[HD];
    {
        const float* qp = p.q + ((size_t)b * p.Lq + (qok ? qi : 0)) * p.ldq + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 t = *reinterpret_cast<const float4*>(qp + d);
            q[d] = t.x * p.scale; q[d + 1] = t.y * p.scale; q[d + 2] = t.z * p.scale; q[d + 3] = t.w * p.scale;
        }
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;

    for (int k0 = 0; k0 < p.Lk; k0 += TK) {
        __syncthreads();
        for (int t = threadIdx.x; t < TK * HD / 4; t += ATT_THREADS) {
            const int r = t / (HD / 4), c = (t % (HD / 4)) * 4;
            const int j = k0 + r;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (j < p.Lk) {
                kv = *reinterpret_cast<const float4*>(p.k + ((size_t)b * p.Lk + j) * p.ldk + h * HD + c);
                vv = *reinterpret_cast<const float4*>(p.v + ((size_t)b * p.Lk + j) * p.ldv + h * HD + c);
            }
            *reinterpret_cast<float4*>(&sK[r][c]) = kv;
            *reinterpret_cast<float4*>(&sV[r][c]) = vv;
        }
        __syncthreads();
        const int jmax = min(TK, p.Lk - k0);
        for (int r = kq; r < jmax; r += 4) {
            const int j = k0 + r;
            if (p.kpm && p.kpm[(size_t)b * p.Lk + j]) continue;
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(&sK[r][d]);
                s = fmaf(q[d], kk.x, s); s = fmaf(q[d + 1], kk.y, s); s = fmaf(q[d + 2], kk.z, s); s = fmaf(q[d + 3], kk.w, s);
            }
            const float mn = fmaxf(m, s);
            const float corr = __expf(m - mn);       // exp(-inf) = 0 on the first key
            const float pe = __expf(s - mn);
            l = l * corr + pe;
            float pw = pe;
            if (p.drop_p > 0.f) pw = keep_elem(p, seed, b, h, qi, j) ? pe * inv_keep : 0.f;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 vv = *reinterpret_cast<const float4*>(&sV[r][d]);
                acc[d] = fmaf(acc[d], corr, pw * vv.x);
                acc[d + 1] = fmaf(acc[d + 1], corr, pw * vv.y);
                acc[d + 2] = fmaf(acc[d + 2], corr, pw * vv.z);
                acc[d + 3] = fmaf(acc[d + 3], corr, pw * vv.w);
            }
            m = mn;
        }
    }
    // merge the 4 key-partitions of each query
    float mall = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    mall = fmaxf(mall, __shfl_xor_sync(0xffffffffu, mall, 2));
    const float f = (m == -INFINITY) ? 0.f : __expf(m - mall);
    const float lall = quad_sum(l * f);
    const float inv = lall > 0.f ? 1.f / lall : 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = quad_sum(acc[d] * f) * inv;
    if (qok) {
        float* op = p.out + ((size_t)b * p.Lq + qi) * p.ldo + h * HD;
        // each of the 4 threads writes 8 of the 32 channels
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            if ((d >> 3) == kq) *reinterpret_cast<float4*>(op + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
        }
        if (kq == 0) p.lse[((size_t)b * p.H + h) * p.Lq + qi] = (lall > 0.f) ? (mall + __logf(lall)) : -INFINITY;
    }
}

// delta[b][h][i] = dO_i . O_i
__global__ void attn_delta_kernel(const AttnParams p) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // one thread per (b, i, h)
    const long long n = (long long)p.B * p.Lq * p.H;
    if (idx >= n) return;
    const int h = (int)(idx % p.H);
    const long long bi = idx / p.H;
    const int i = (int)(bi % p.Lq);
    const int b = (int)(bi / p.Lq);
    const float* a = p.dout + (size_t)bi * p.ldo + h * HD;
    const float* o = p.o + (size_t)bi * p.ldo + h * HD;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
        const float4 x = *reinterpret_cast<const float4*>(a + d);
        const float4 y = *reinterpret_cast<const float4*>(o + d);
        s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    p.delta[((size_t)b * p.H + h) * p.Lq + i] = s;
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const AttnParams p) {
    __shared__ __align__(16) float sK[TK][HD];
    __shared__ __align__(16) float sV[TK][HD];
    const int b = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 2);
    const int kq = threadIdx.x & 3;
    const bool qok = qi < p.Lq;
    const unsigned long long seed = (p.drop_p > 0.f) ? (*p.seed + p.site * 0x9E3779B97F4A7C15ull) : 0ull;
    const float inv_keep = 1.f / (1.f - p.drop_p);
    float q[HD], go[HD], dq[HD];
    const size_t row = (size_t)b * p.Lq + (qok ? qi : 0);
    {
        const float* qp = p.q + row * p.ldq + h * HD;
        const float* gp = p.dout + row * p.ldo + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) { q[d] = qp[d] * p.scale; go[d] = gp[d]; dq[d] = 0.f; }
    }
    const size_t st = ((size_t)b * p.H + h) * p.Lq + (qok ? qi : 0);
    const float lse = p.lse[st], delta = p.delta[st];

    for (int k0 = 0; k0 < p.Lk; k0 += TK) {
        __syncthreads();
        for (int t = threadIdx.x; t < TK * HD / 4; t += ATT_THREADS) {
            const int r = t / (HD / 4), c = (t % (HD / 4)) * 4;
            const int j = k0 + r;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (j < p.Lk) {
                kv = *reinterpret_cast<const float4*>(p.k + ((size_t)b * p.Lk + j) * p.ldk + h * HD + c);
                vv = *reinterpret_cast<const float4*>(p.v + ((size_t)b * p.Lk + j) * p.ldv + h * HD + c);
            }
            *reinterpret_cast<float4*>(&sK[r][c]) = kv;
            *reinterpret_cast<float4*>(&sV[r][c]) = vv;
        }
        __syncthreads();
        const int jmax = min(TK, p.Lk - k0);
        for (int r = kq; r < jmax; r += 4) {
            const int j = k0 + r;
            if (p.kpm && p.kpm[(size_t)b * p.Lk + j]) continue;
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(&sK[r][d]);
                const float4 vv = *reinterpret_cast<const float4*>(&sV[r][d]);
                s = fmaf(q[d], kk.x, s); s = fmaf(q[d + 1], kk.y, s); s = fmaf(q[d + 2], kk.z, s); s = fmaf(q[d + 3], kk.w, s);
                dp = fmaf(go[d], vv.x, dp); dp = fmaf(go[d + 1], vv.y, dp); dp = fmaf(go[d + 2], vv.z, dp); dp = fmaf(go[d + 3], vv.w, dp);
            }
            const float pe = (lse == -INFINITY) ? 0.f : __expf(s - lse);
            if (p.drop_p > 0.f) dp = keep_elem(p, seed, b, h, qi, j) ? dp * inv_keep : 0.f;
            const float ds = pe * (dp - delta);
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(&sK[r][d]);
                dq[d] = fmaf(ds, kk.x, dq[d]); dq[d + 1] = fmaf(ds, kk.y, dq[d + 1]);
                dq[d + 2] = fmaf(ds, kk.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kk.w, dq[d + 3]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = quad_sum(dq[d]) * p.scale;
    if (qok) {
        float* op = p.dq + row * p.lddq + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4)
            if ((d >> 3) == kq) *reinterpret_cast<float4*>(op + d) = make_float4(dq[d], dq[d + 1], dq[d + 2], dq[d + 3]);
    }
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const AttnParams p) {
    __shared__ __align__(16) float sQ[TK][HD];
    __shared__ __align__(16) float sG[TK][HD];
    __shared__ float sL[TK], sD[TK];
    const int b = blockIdx.z, h = blockIdx.y;
    const int kj = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 2);
    const int part = threadIdx.x & 3;
    const bool kok = kj < p.Lk;
    const unsigned long long seed = (p.drop_p > 0.f) ? (*p.seed + p.site * 0x9E3779B97F4A7C15ull) : 0ull;
    const float inv_keep = 1.f / (1.f - p.drop_p);
    float kk[HD], vv[HD], dk[HD], dv[HD];
    const size_t row = (size_t)b * p.Lk + (kok ? kj : 0);
    {
        const float* kp = p.k + row * p.ldk + h * HD;
        const float* vp = p.v + row * p.ldv + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) { kk[d] = kp[d]; vv[d] = vp[d]; dk[d] = 0.f; dv[d] = 0.f; }
    }
    const bool masked = !kok || (p.kpm && p.kpm[(size_t)b * p.Lk + kj]);

    for (int q0 = 0; q0 < p.Lq; q0 += TK) {
        __syncthreads();
        for (int t = threadIdx.x; t < TK * HD / 4; t += ATT_THREADS) {
            const int r = t / (HD / 4), c = (t % (HD / 4)) * 4;
            const int i = q0 + r;
            float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), gv = qv;
            if (i < p.Lq) {
                qv = *reinterpret_cast<const float4*>(p.q + ((size_t)b * p.Lq + i) * p.ldq + h * HD + c);
                gv = *reinterpret_cast<const float4*>(p.dout + ((size_t)b * p.Lq + i) * p.ldo + h * HD + c);
            }
            *reinterpret_cast<float4*>(&sQ[r][c]) = qv;
            *reinterpret_cast<float4*>(&sG[r][c]) = gv;
        }
        for (int t = threadIdx.x; t < TK; t += ATT_THREADS) {
            const int i = q0 + t;
            const size_t st = ((size_t)b * p.H + h) * p.Lq + min(i, p.Lq - 1);
            sL[t] = (i < p.Lq) ? p.lse[st] : INFINITY;
            sD[t] = (i < p.Lq) ? p.delta[st] : 0.f;
        }
        __syncthreads();
        if (masked) continue;
        const int imax = min(TK, p.Lq - q0);
        for (int r = part; r < imax; r += 4) {
            const int i = q0 + r;
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 qq = *reinterpret_cast<const float4*>(&sQ[r][d]);
                const float4 gg = *reinterpret_cast<const float4*>(&sG[r][d]);
                s = fmaf(qq.x, kk[d], s); s = fmaf(qq.y, kk[d + 1], s); s = fmaf(qq.z, kk[d + 2], s); s = fmaf(qq.w, kk[d + 3], s);
                dp = fmaf(gg.x, vv[d], dp); dp = fmaf(gg.y, vv[d + 1], dp); dp = fmaf(gg.z, vv[d + 2], dp); dp = fmaf(gg.w, vv[d + 3], dp);
            }
            const float pe = (sL[r] == -INFINITY) ? 0.f : __expf(s * p.scale - sL[r]);
            float pw = pe;
            if (p.drop_p > 0.f) {
                const bool keep = keep_elem(p, seed, b, h, i, kj);
                pw = keep ? pe * inv_keep : 0.f;
                dp = keep ? dp * inv_keep : 0.f;
            }
            const float ds = pe * (dp - sD[r]) * p.scale;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 qq = *reinterpret_cast<const float4*>(&sQ[r][d]);
                const float4 gg = *reinterpret_cast<const float4*>(&sG[r][d]);
                dk[d] = fmaf(ds, qq.x, dk[d]); dk[d + 1] = fmaf(ds, qq.y, dk[d + 1]);
                dk[d + 2] = fmaf(ds, qq.z, dk[d + 2]); dk[d + 3] = fmaf(ds, qq.w, dk[d + 3]);
                dv[d] = fmaf(pw, gg.x, dv[d]); dv[d + 1] = fmaf(pw, gg.y, dv[d + 1]);
                dv[d + 2] = fmaf(pw, gg.z, dv[d + 2]); dv[d + 3] = fmaf(pw, gg.w, dv[d + 3]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) { dk[d] = quad_sum(dk[d]); dv[d] = quad_sum(dv[d]); }
    if (kok) {
        float* dkp = p.dk + row * p.lddk + h * HD;
        float* dvp = p.dv + row * p.lddv + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4)
            if ((d >> 3) == part) {
                *reinterpret_cast<float4*>(dkp + d) = make_float4(dk[d], dk[d + 1], dk[d + 2], dk[d + 3]);
                *reinterpret_cast<float4*>(dvp + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
            }
    }
}

int check(const AttnParams& p) {
    if (p.B <= 0 || p.H <= 0 || p.Lq <= 0 || p.Lk <= 0) return MDB_EINVAL;
    if ((p.ldq | p.ldk | p.ldv | p.ldo) % 4) return MDB_EINVAL;
    if (p.drop_p < 0.f || p.drop_p >= 1.f) return MDB_EINVAL;
    if (p.drop_p > 0.f && !p.seed) return MDB_EINVAL;
    if (p.H > 65535 || p.B > 65535) return MDB_EUNSUPPORTED;
    return 0;
}

}  // namespace

extern "C" {

int mdb_attention_forward_f32(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask,
                              float* out, float* lse, int B, int H, int Lq, int Lk, int head_dim, int ldq, int ldk,
                              int ldv, int ldo, float drop_p, const unsigned long long* seed, unsigned long long site,
                              void* stream) {
    if (head_dim != HD) return MDB_EUNSUPPORTED;
    if (!q || !k || !v || !out || !lse) return MDB_EINVAL;
    AttnParams p{};
    p.q = q; p.k = k; p.v = v; p.kpm = key_padding_mask; p.out = out; p.lse = lse;
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
    p.scale = 1.f / sqrtf((float)HD); p.drop_p = drop_p; p.seed = seed; p.site = site;
    int rc = check(p);
    if (rc) return rc;
    dim3 grid((Lq + ROWS_PER_CTA - 1) / ROWS_PER_CTA, H, B);
    attn_fwd_kernel<<<grid, ATT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return (int)cudaGetLastError();
}

int mdb_attention_backward_f32(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask,
                               const float* out, const float* lse, const float* dout, float* delta_ws, float* dq,
                               float* dk, float* dv, int B, int H, int Lq, int Lk, int head_dim, int ldq, int ldk, int ldv,
                               int ldo, int lddq, int lddk, int lddv, float drop_p, const unsigned long long* seed,
                               unsigned long long site, void* stream_) {
    if (head_dim != HD) return MDB_EUNSUPPORTED;
    if (!q || !k || !v || !out || !lse || !dout || !delta_ws || !dq || !dk || !dv) return MDB_EINVAL;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    AttnParams p{};
    p.q = q; p.k = k; p.v = v; p.kpm = key_padding_mask; p.lse = const_cast<float*>(lse);
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
    p.scale = 1.f / sqrtf((float)HD); p.drop_p = drop_p; p.seed = seed; p.site = site;
    p.dout = dout; p.o = out; p.delta = delta_ws; p.dq = dq; p.dk = dk; p.dv = dv;
    p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    int rc = check(p);
    if (rc) return rc;
    if ((lddq | lddk | lddv) % 4) return MDB_EINVAL;
    const long long n = (long long)B * Lq * H;
    attn_delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p);
    attn_bwd_dq_kernel<<<dim3((Lq + ROWS_PER_CTA - 1) / ROWS_PER_CTA, H, B), ATT_THREADS, 0, stream>>>(p);
    attn_bwd_dkv_kernel<<<dim3((Lk + ROWS_PER_CTA - 1) / ROWS_PER_CTA, H, B), ATT_THREADS, 0, stream>>>(p);
    return (int)cudaGetLastError();
}

}  // extern "C"
