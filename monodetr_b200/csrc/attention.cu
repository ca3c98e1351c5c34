// attention.cu -- fused multi-head attention core (softmax(Q K^T / sqrt(d)) V, head dim 32) for sm_100a,
// forward and backward, flash-style: the (Lq x Lk) probability matrix never touches HBM (the reference's
// nn.MultiheadAttention materialises it: 118 MB per image in the depth encoder).
// Replaces F.multi_head_attention_forward's core at depthaware_transformer.py:456-459 (depth cross-attn),
// :496 (group self-attn) and depth_predictor/transformer.py:59 (depth encoder).  The in/out projections
// are separate tensor-core GEMMs (conv_gemm.cu).
//
// Layout: q[b][i][h][32] with token stride ldq floats (so a packed QKV buffer can be passed), same for
// k, v (ldk, ldv), out[b][i][h*32] with token stride ldo.  key_padding_mask[b][j] (uint8, nonzero = ignore) or
// null.  Dropout on the probabilities uses a counter-based hash RNG keyed by (seed, b, h, i, j) so the backward
// pass regenerates the same mask.
//
// Mapping: register-resident warp-level tensor-core tiles (mma.sync m16n8k8 TF32, fp32 accumulate).  A CTA of
// 4 warps owns 64 rows (queries, or keys in the dK/dV kernel); each warp owns 16 rows and streams 64-row tiles
// of the other operand through padded shared memory (row stride 36 floats: conflict-free fragment loads).
// Scores (forward AND the recomputation in backward) use error-compensated 3xTF32 (hi/lo split in registers) so
// exp() sees fp32-accurate logits; P.V uses the same in the forward pass; the four gradient contractions of the
// backward pass are single-pass TF32 with round-to-nearest operands (measured 8e-4 of max|grad|).  The P (C-fragment) -> A-fragment hand-off needs no shuffles: the k index of the
// second GEMM is permuted (col t <-> key 2t, col t+4 <-> key 2t+1) and V/K/dO/Q rows are fetched in that order.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/monodetr_b200.h"
#include "rng.cuh"

namespace {

constexpr int HD = 32;             // head dim
constexpr int BR = 64;             // rows per CTA (4 warps x 16)
constexpr int BC = 64;             // streamed tile rows
constexpr int LDS = 36;            // padded smem row stride (floats)
constexpr int ATT_THREADS = 128;

struct AttnParams {
    const float *q, *k, *v;
    const uint8_t* kpm;            // [B][Lk] or null
    float* out;
    float* lse;                    // [B][H][Lq]
    int B, H, Lq, Lk, ldq, ldk, ldv, ldo;
    float scale;                   // 1/sqrt(d)
    float drop_p;
    const unsigned long long* seed;
    unsigned long long site;
    const float *dout, *o;
    float* delta;                  // [B][H][Lq]
    float *dq, *dk, *dv;
    int lddq, lddk, lddv;
};

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (16 rows x 32 cols) of a row-major global matrix: rows r0+g, r0+g+8; hi/lo split, optional scaling.
struct AFrag {
    uint32_t hi[4][4];
    uint32_t lo[4][4];
};
__device__ __forceinline__ void load_afrag(AFrag& f, const float* base, int ld, int r0, int nrows, int lane, float mul) {
    const int g = lane >> 2, t = lane & 3;
    const int ra = r0 + g, rb = r0 + g + 8;
    const float* pa = base + (size_t)min(ra, nrows - 1) * ld;
    const float* pb = base + (size_t)min(rb, nrows - 1) * ld;
    const float ma = ra < nrows ? mul : 0.f, mb = rb < nrows ? mul : 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        split_tf32(pa[ks * 8 + t] * ma, f.hi[ks][0], f.lo[ks][0]);
        split_tf32(pb[ks * 8 + t] * mb, f.hi[ks][1], f.lo[ks][1]);
        split_tf32(pa[ks * 8 + t + 4] * ma, f.hi[ks][2], f.lo[ks][2]);
        split_tf32(pb[ks * 8 + t + 4] * mb, f.hi[ks][3], f.lo[ks][3]);
    }
}

// Streamed tiles are stored PRE-SPLIT in shared memory (hi = rn_tf32(x), lo = rn_tf32(x - hi)) by the loader, once
// per element, instead of every warp re-splitting every B fragment it reads (3 ALU ops per register per use).
// C[16 x 64] = A[16 x 32] . S^T, S = smem tile [64][LDS] (rows = the 64 output columns).  NS = 3: error-compensated.
template <int NS>
__device__ __forceinline__ void gemm_nt(float (&c)[8][4], const AFrag& a, const uint32_t (*sh)[LDS], const uint32_t (*sl)[LDS], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const uint32_t b0h = sh[nt * 8 + g][ks * 8 + t], b1h = sh[nt * 8 + g][ks * 8 + t + 4];
            if (NS == 3) {
                const uint32_t b0l = sl[nt * 8 + g][ks * 8 + t], b1l = sl[nt * 8 + g][ks * 8 + t + 4];
                mma8(c[nt], a.lo[ks], b0h, b1h);
                mma8(c[nt], a.hi[ks], b0l, b1l);
            }
            mma8(c[nt], a.hi[ks], b0h, b1h);
        }
    }
}

// acc[16 x 32] += P[16 x 64] . S, P given as C fragments (cols 2t,2t+1 of each 8-wide tile), S = smem [64][LDS].
// k-permutation trick: A col t <-> key 2t, A col t+4 <-> key 2t+1, so A = (c0, c2, c1, c3) and B rows 2t / 2t+1.
template <int NS>
__device__ __forceinline__ void gemm_nn(float (&acc)[4][4], const float (&p)[8][4], const uint32_t (*sh)[LDS], const uint32_t (*sl)[LDS], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
        uint32_t ah[4], al[4];
        split_tf32(p[kt][0], ah[0], al[0]);
        split_tf32(p[kt][2], ah[1], al[1]);
        split_tf32(p[kt][1], ah[2], al[2]);
        split_tf32(p[kt][3], ah[3], al[3]);
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
            const uint32_t b0h = sh[kt * 8 + 2 * t][dn * 8 + g], b1h = sh[kt * 8 + 2 * t + 1][dn * 8 + g];
            if (NS == 3) {
                const uint32_t b0l = sl[kt * 8 + 2 * t][dn * 8 + g], b1l = sl[kt * 8 + 2 * t + 1][dn * 8 + g];
                mma8(acc[dn], al, b0h, b1h);
                mma8(acc[dn], ah, b0l, b1l);
            }
            mma8(acc[dn], ah, b0h, b1h);
        }
    }
}

// cooperative load of a [64][32] tile (rows r0.., zero beyond nrows) into padded smem, split into hi (and lo)
template <bool LO>
__device__ __forceinline__ void load_tile(uint32_t (*sh)[LDS], uint32_t (*sl)[LDS], const float* base, int ld, int r0, int nrows) {
    for (int i = threadIdx.x; i < BC * (HD / 4); i += ATT_THREADS) {
        const int r = i >> 3, c = (i & 7) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < nrows) v = *reinterpret_cast<const float4*>(base + (size_t)(r0 + r) * ld + c);
        uint4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<uint4*>(&sh[r][c]) = h;
        if (LO) *reinterpret_cast<uint4*>(&sl[r][c]) = l;
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

__device__ __forceinline__ float keep_scale(const AttnParams& p, unsigned long long seed, int b, int h, int i, int j, float inv_keep) {
    const unsigned long long idx = (((unsigned long long)(b * p.H + h) * p.Lq + i) * (unsigned long long)p.Lk + j);
    return mdb::rng_uniform(seed, idx) >= p.drop_p ? inv_keep : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnParams p) {
    __shared__ __align__(16) uint32_t sKh[BC][LDS], sKl[BC][LDS];
    __shared__ __align__(16) uint32_t sVh[BC][LDS], sVl[BC][LDS];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * BR + warp * 16;
    const float* qb = p.q + (size_t)b * p.Lq * p.ldq + h * HD;
    const float* kb = p.k + (size_t)b * p.Lk * p.ldk + h * HD;
    const float* vb = p.v + (size_t)b * p.Lk * p.ldv + h * HD;
    const unsigned long long seed = (p.drop_p > 0.f) ? (*p.seed + p.site * 0x9E3779B97F4A7C15ull) : 0ull;
    const float inv_keep = 1.f / (1.f - p.drop_p);
    AFrag qa;
    load_afrag(qa, qb, p.ldq, r0, p.Lq, lane, p.scale);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // rows g and g+8
    const int qi0 = r0 + g, qi1 = r0 + g + 8;

    for (int k0 = 0; k0 < p.Lk; k0 += BC) {
        __syncthreads();
        load_tile<true>(sKh, sKl, kb, p.ldk, k0, p.Lk);
        load_tile<true>(sVh, sVl, vb, p.ldv, k0, p.Lk);
        __syncthreads();
        float s[8][4];
        gemm_nt<3>(s, qa, sKh, sKl, lane);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = k0 + nt * 8 + 2 * t + e;
                const bool dead = (j >= p.Lk) || (p.kpm && p.kpm[(size_t)b * p.Lk + min(j, p.Lk - 1)]);
                if (dead) { s[nt][e] = -INFINITY; s[nt][2 + e] = -INFINITY; }
                mx0 = fmaxf(mx0, s[nt][e]);
                mx1 = fmaxf(mx1, s[nt][2 + e]);
            }
        }
        const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
        const float c0 = (mn0 == -INFINITY) ? 1.f : __expf(m0 - mn0), c1 = (mn1 == -INFINITY) ? 1.f : __expf(m1 - mn1);
        l0 *= c0; l1 *= c1;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) { acc[dn][0] *= c0; acc[dn][1] *= c0; acc[dn][2] *= c1; acc[dn][3] *= c1; }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = k0 + nt * 8 + 2 * t + e;
                float p0 = (mn0 == -INFINITY) ? 0.f : __expf(s[nt][e] - mn0);
                float p1 = (mn1 == -INFINITY) ? 0.f : __expf(s[nt][2 + e] - mn1);
                l0 += p0; l1 += p1;
                if (p.drop_p > 0.f) {
                    p0 *= keep_scale(p, seed, b, h, qi0, j, inv_keep);
                    p1 *= keep_scale(p, seed, b, h, qi1, j, inv_keep);
                }
                s[nt][e] = p0; s[nt][2 + e] = p1;
            }
        }
        m0 = mn0; m1 = mn1;
        gemm_nn<3>(acc, s, sVh, sVl, lane);
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    float* ob = p.out + (size_t)b * p.Lq * p.ldo + h * HD;
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
        if (qi0 < p.Lq) *reinterpret_cast<float2*>(ob + (size_t)qi0 * p.ldo + dn * 8 + 2 * t) = make_float2(acc[dn][0] * i0, acc[dn][1] * i0);
        if (qi1 < p.Lq) *reinterpret_cast<float2*>(ob + (size_t)qi1 * p.ldo + dn * 8 + 2 * t) = make_float2(acc[dn][2] * i1, acc[dn][3] * i1);
    }
    if (t == 0) {
        float* lp = p.lse + ((size_t)b * p.H + h) * p.Lq;
        if (qi0 < p.Lq) lp[qi0] = l0 > 0.f ? m0 + __logf(l0) : -INFINITY;
        if (qi1 < p.Lq) lp[qi1] = l1 > 0.f ? m1 + __logf(l1) : -INFINITY;
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

// dQ: CTA = 64 queries, streams key/value tiles.
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const AttnParams p) {
    __shared__ __align__(16) uint32_t sKh[BC][LDS], sKl[BC][LDS];
    __shared__ __align__(16) uint32_t sVh[BC][LDS];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * BR + warp * 16;
    const float* qb = p.q + (size_t)b * p.Lq * p.ldq + h * HD;
    const float* gb = p.dout + (size_t)b * p.Lq * p.ldo + h * HD;
    const float* kb = p.k + (size_t)b * p.Lk * p.ldk + h * HD;
    const float* vb = p.v + (size_t)b * p.Lk * p.ldv + h * HD;
    const unsigned long long seed = (p.drop_p > 0.f) ? (*p.seed + p.site * 0x9E3779B97F4A7C15ull) : 0ull;
    const float inv_keep = 1.f / (1.f - p.drop_p);
    AFrag qa, ga;
    load_afrag(qa, qb, p.ldq, r0, p.Lq, lane, p.scale);
    load_afrag(ga, gb, p.ldo, r0, p.Lq, lane, 1.f);
    const int qi0 = r0 + g, qi1 = r0 + g + 8;
    const size_t st = ((size_t)b * p.H + h) * p.Lq;
    const float lse0 = qi0 < p.Lq ? p.lse[st + qi0] : INFINITY, lse1 = qi1 < p.Lq ? p.lse[st + qi1] : INFINITY;
    const float dl0 = qi0 < p.Lq ? p.delta[st + qi0] : 0.f, dl1 = qi1 < p.Lq ? p.delta[st + qi1] : 0.f;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    for (int k0 = 0; k0 < p.Lk; k0 += BC) {
        __syncthreads();
        load_tile<true>(sKh, sKl, kb, p.ldk, k0, p.Lk);
        load_tile<false>(sVh, nullptr, vb, p.ldv, k0, p.Lk);
        __syncthreads();
        float s[8][4], dp[8][4];
        gemm_nt<3>(s, qa, sKh, sKl, lane);
        gemm_nt<1>(dp, ga, sVh, nullptr, lane);   // gradients: single-pass TF32 with round-to-nearest operands
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = k0 + nt * 8 + 2 * t + e;
                const bool dead = (j >= p.Lk) || (p.kpm && p.kpm[(size_t)b * p.Lk + min(j, p.Lk - 1)]);
                float p0 = (dead || lse0 == -INFINITY) ? 0.f : __expf(s[nt][e] - lse0);
                float p1 = (dead || lse1 == -INFINITY) ? 0.f : __expf(s[nt][2 + e] - lse1);
                float d0 = dp[nt][e], d1 = dp[nt][2 + e];
                if (p.drop_p > 0.f) {
                    d0 *= keep_scale(p, seed, b, h, qi0, j, inv_keep);
                    d1 *= keep_scale(p, seed, b, h, qi1, j, inv_keep);
                }
                s[nt][e] = p0 * (d0 - dl0);
                s[nt][2 + e] = p1 * (d1 - dl1);
            }
        }
        gemm_nn<1>(acc, s, sKh, nullptr, lane);
    }
    float* ob = p.dq + (size_t)b * p.Lq * p.lddq + h * HD;
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
        if (qi0 < p.Lq) *reinterpret_cast<float2*>(ob + (size_t)qi0 * p.lddq + dn * 8 + 2 * t) = make_float2(acc[dn][0] * p.scale, acc[dn][1] * p.scale);
        if (qi1 < p.Lq) *reinterpret_cast<float2*>(ob + (size_t)qi1 * p.lddq + dn * 8 + 2 * t) = make_float2(acc[dn][2] * p.scale, acc[dn][3] * p.scale);
    }
}

// dK / dV: CTA = 64 keys, streams query tiles (Q, dO, lse, delta).
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const AttnParams p) {
    __shared__ __align__(16) uint32_t sQh[BC][LDS], sQl[BC][LDS];
    __shared__ __align__(16) uint32_t sGh[BC][LDS];
    __shared__ float sL[BC], sD[BC];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * BR + warp * 16;                      // first key row of this warp
    const float* qb = p.q + (size_t)b * p.Lq * p.ldq + h * HD;
    const float* gb = p.dout + (size_t)b * p.Lq * p.ldo + h * HD;
    const float* kb = p.k + (size_t)b * p.Lk * p.ldk + h * HD;
    const float* vb = p.v + (size_t)b * p.Lk * p.ldv + h * HD;
    const unsigned long long seed = (p.drop_p > 0.f) ? (*p.seed + p.site * 0x9E3779B97F4A7C15ull) : 0ull;
    const float inv_keep = 1.f / (1.f - p.drop_p);
    AFrag ka, va;
    load_afrag(ka, kb, p.ldk, r0, p.Lk, lane, p.scale);
    load_afrag(va, vb, p.ldv, r0, p.Lk, lane, 1.f);
    const int kj0 = r0 + g, kj1 = r0 + g + 8;
    const bool dead0 = kj0 >= p.Lk || (p.kpm && p.kpm[(size_t)b * p.Lk + min(kj0, p.Lk - 1)]);
    const bool dead1 = kj1 >= p.Lk || (p.kpm && p.kpm[(size_t)b * p.Lk + min(kj1, p.Lk - 1)]);
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const size_t st = ((size_t)b * p.H + h) * p.Lq;

    for (int q0 = 0; q0 < p.Lq; q0 += BC) {
        __syncthreads();
        load_tile<true>(sQh, sQl, qb, p.ldq, q0, p.Lq);
        load_tile<false>(sGh, nullptr, gb, p.ldo, q0, p.Lq);
        if (threadIdx.x < BC) {
            const int i = q0 + threadIdx.x;
            sL[threadIdx.x] = i < p.Lq ? p.lse[st + i] : INFINITY;     // +inf -> p = 0 for rows past the end
            sD[threadIdx.x] = i < p.Lq ? p.delta[st + i] : 0.f;
        }
        __syncthreads();
        float s[8][4], dp[8][4];
        gemm_nt<3>(s, ka, sQh, sQl, lane);    // S^T[key][query] (already scaled through K)
        gemm_nt<1>(dp, va, sGh, nullptr, lane);   // dP^T[key][query] = V . dO^T
        float pd[8][4];                        // dropped probabilities for dV
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int il = nt * 8 + 2 * t + e, i = q0 + il;
                const float lse = sL[il], dl = sD[il];
                const bool nolse = (lse == -INFINITY);
                float p0 = (dead0 || nolse) ? 0.f : __expf(s[nt][e] - lse);
                float p1 = (dead1 || nolse) ? 0.f : __expf(s[nt][2 + e] - lse);
                float d0 = dp[nt][e], d1 = dp[nt][2 + e];
                float w0 = p0, w1 = p1;
                if (p.drop_p > 0.f) {
                    const float k0s = keep_scale(p, seed, b, h, i, kj0, inv_keep), k1s = keep_scale(p, seed, b, h, i, kj1, inv_keep);
                    w0 *= k0s; w1 *= k1s; d0 *= k0s; d1 *= k1s;
                }
                pd[nt][e] = w0; pd[nt][2 + e] = w1;
                s[nt][e] = p0 * (d0 - dl);
                s[nt][2 + e] = p1 * (d1 - dl);
            }
        }
        gemm_nn<1>(dv, pd, sGh, nullptr, lane);   // dV += P^T_dropped . dO
        gemm_nn<1>(dk, s, sQh, nullptr, lane);    // dK += dS^T . Q
    }
    float* dkb = p.dk + (size_t)b * p.Lk * p.lddk + h * HD;
    float* dvb = p.dv + (size_t)b * p.Lk * p.lddv + h * HD;
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
        if (kj0 < p.Lk) {
            *reinterpret_cast<float2*>(dkb + (size_t)kj0 * p.lddk + dn * 8 + 2 * t) = make_float2(dk[dn][0] * p.scale, dk[dn][1] * p.scale);
            *reinterpret_cast<float2*>(dvb + (size_t)kj0 * p.lddv + dn * 8 + 2 * t) = make_float2(dv[dn][0], dv[dn][1]);
        }
        if (kj1 < p.Lk) {
            *reinterpret_cast<float2*>(dkb + (size_t)kj1 * p.lddk + dn * 8 + 2 * t) = make_float2(dk[dn][2] * p.scale, dk[dn][3] * p.scale);
            *reinterpret_cast<float2*>(dvb + (size_t)kj1 * p.lddv + dn * 8 + 2 * t) = make_float2(dv[dn][2], dv[dn][3]);
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
    dim3 grid((Lq + BR - 1) / BR, H, B);
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
    attn_bwd_dq_kernel<<<dim3((Lq + BR - 1) / BR, H, B), ATT_THREADS, 0, stream>>>(p);
    attn_bwd_dkv_kernel<<<dim3((Lk + BR - 1) / BR, H, B), ATT_THREADS, 0, stream>>>(p);
    return (int)cudaGetLastError();
}

}  // extern "C"
