// attention.cu -- fused multi-head attention core (softmax(Q K^T / sqrt(d)) V, head dim 32) for sm_100a,
// forward and backward, flash-style: the (Lq x Lk) probability matrix never touches HBM (the reference's
// nn.MultiheadAttention materialises it: 118 MB per image in the depth encoder).
// Replaces F.multi_head_attention_forward's core at depthaware_transformer.py:456-459 (depth cross-attn),
// :496 (group self-attn) and depth_predictor/transformer.py:59 (depth encoder).  The in/out projections
// are separate tensor-core GEMMs (conv_gemm.cu).
//
// Layout: q[b][i][h][32] with token stride ldq floats (so a packed QKV buffer can be passed), same for
// k, v (ldk, ldv), out[b][i][h*32] with token stride ldo.  key_padding_mask[b][j] (uint8, nonzero = ignore) or
// null.  Dropout on the probabilities uses a counter-based hash RNG keyed by (seed, site, b, h, i, j) so the backward
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

// Streaming of a [64][32] tile is software-pipelined with cp.async: stage_tile issues this thread's four 16-byte
// asynchronous copies of tile j+1 into a raw staging buffer right before the tensor-core work on tile j; after it,
// unstage_tile reads the same four slots back (own copies only: cp.async.wait_all suffices, no barrier), splits them
// into hi (and lo) and writes the padded tiles.  The HBM/L2 latency hides behind the MMAs instead of stalling all
// four warps at the barrier (long-scoreboard was the top stall of the synchronous version, profiles/r01_attn_fwd_r1.txt)
// and no registers are held across the MMAs (a register-prefetch variant cost 57-80 registers and a CTA per SM).
constexpr int STG = BC * HD;       // floats per staging buffer
__device__ __forceinline__ void stage_tile(float* stg, const float* base, int ld, int r0, int nrows) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * ATT_THREADS;
        const int r = i >> 3, c = (i & 7) * 4;
        const bool ok = r0 + r < nrows;
        const float* src = ok ? base + ((size_t)(r0 + r) * ld + c) : base;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stg) + (uint32_t)i * 16u;
        const int nbytes = ok ? 16 : 0;                              // 0 -> the 16 bytes are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    }
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <bool LO>
__device__ __forceinline__ void unstage_tile(uint32_t (*sh)[LDS], uint32_t (*sl)[LDS], const float* stg) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * ATT_THREADS;
        const int r = i >> 3, c = (i & 7) * 4;
        const float4 v = *reinterpret_cast<const float4*>(stg + i * 4);
        uint4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<uint4*>(&sh[r][c]) = h;
        if (LO) *reinterpret_cast<uint4*>(&sl[r][c]) = l;
    }
}
static_assert(BC * (HD / 4) == 4 * ATT_THREADS, "tile loader mapping");

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Dropout on the probabilities: keep(b, h, i, j) = fmix32-hash of (i * Lk + j) keyed per (seed, site, b, h); the
// same pure function in all three kernels.  Host check: Lq * Lk < 2^32.
struct DropCtx {
    uint32_t key, thr;
    float inv_keep;
    bool on;
};
__device__ __forceinline__ DropCtx make_drop(const AttnParams& p, int b, int h) {
    DropCtx d;
    d.on = p.drop_p > 0.f;
    d.inv_keep = 1.f / (1.f - p.drop_p);
    d.thr = mdb::rng_thr16(p.drop_p);
    d.key = 0u;
    if (d.on) d.key = mdb::rng_key32(*p.seed + p.site * 0x9E3779B97F4A7C15ull, (unsigned long long)(b * p.H + h));
    return d;
}
__device__ __forceinline__ float keep_scale(const DropCtx& d, uint32_t row_base, int j) {   // row_base = i * Lk
    return mdb::rng_keep16(d.key, row_base + (uint32_t)j, d.thr) ? d.inv_keep : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// dead-key flags of the streamed tile (key_padding_mask or past the end), fetched with the tile and read from smem
__device__ __forceinline__ uint32_t fetch_dead(const AttnParams& p, int b, int k0) {
    const int j = k0 + (int)threadIdx.x;
    if (threadIdx.x >= BC) return 0u;
    if (j >= p.Lk) return 1u;
    return (p.kpm && p.kpm[(size_t)b * p.Lk + j]) ? 1u : 0u;
}

__global__ void __launch_bounds__(ATT_THREADS, 4)
attn_fwd_kernel(const AttnParams p) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];             // kFwdSmem bytes: 4 padded tiles, 2 staging buffers, flags
    uint32_t (*sKh)[LDS] = reinterpret_cast<uint32_t (*)[LDS]>(dyn_smem);
    uint32_t (*sKl)[LDS] = sKh + BC;
    uint32_t (*sVh)[LDS] = sKl + BC;
    uint32_t (*sVl)[LDS] = sVh + BC;
    float* gK = reinterpret_cast<float*>(sVl + BC);
    float* gV = gK + STG;
    uint8_t* sDead = reinterpret_cast<uint8_t*>(gV + STG);
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * BR + warp * 16;
    const float* qb = p.q + (size_t)b * p.Lq * p.ldq + h * HD;
    const float* kb = p.k + (size_t)b * p.Lk * p.ldk + h * HD;
    const float* vb = p.v + (size_t)b * p.Lk * p.ldv + h * HD;
    const DropCtx drop = make_drop(p, b, h);
    stage_tile(gK, kb, p.ldk, 0, p.Lk);
    stage_tile(gV, vb, p.ldv, 0, p.Lk);
    stage_commit();
    uint32_t ndead = fetch_dead(p, b, 0);
    AFrag qa;
    load_afrag(qa, qb, p.ldq, r0, p.Lq, lane, p.scale);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // rows g and g+8
    const int qi0 = r0 + g, qi1 = r0 + g + 8;
    const uint32_t rb0 = (uint32_t)qi0 * (uint32_t)p.Lk, rb1 = (uint32_t)qi1 * (uint32_t)p.Lk;

    for (int k0 = 0; k0 < p.Lk; k0 += BC) {
        stage_wait();
        __syncthreads();                       // every warp is done with the previous tiles
        unstage_tile<true>(sKh, sKl, gK);
        unstage_tile<true>(sVh, sVl, gV);
        if (threadIdx.x < BC) sDead[threadIdx.x] = (uint8_t)ndead;
        __syncthreads();
        if (k0 + BC < p.Lk) {                  // next tile's copies fly during this tile's MMAs
            stage_tile(gK, kb, p.ldk, k0 + BC, p.Lk);
            stage_tile(gV, vb, p.ldv, k0 + BC, p.Lk);
            stage_commit();
            ndead = fetch_dead(p, b, k0 + BC);
        }
        float s[8][4];
        gemm_nt<3>(s, qa, sKh, sKl, lane);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uchar2 dd = *reinterpret_cast<const uchar2*>(&sDead[nt * 8 + 2 * t]);
            if (dd.x) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (dd.y) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
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
                if (drop.on) {
                    p0 *= keep_scale(drop, rb0, j);
                    p1 *= keep_scale(drop, rb1, j);
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
__global__ void __launch_bounds__(ATT_THREADS, 3)
attn_bwd_dq_kernel(const AttnParams p) {
    __shared__ __align__(16) uint32_t sKh[BC][LDS], sKl[BC][LDS];
    __shared__ __align__(16) uint32_t sVh[BC][LDS];
    __shared__ __align__(16) float gK[STG], gV[STG];
    __shared__ __align__(8) uint8_t sDead[BC];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * BR + warp * 16;
    const float* qb = p.q + (size_t)b * p.Lq * p.ldq + h * HD;
    const float* gb = p.dout + (size_t)b * p.Lq * p.ldo + h * HD;
    const float* kb = p.k + (size_t)b * p.Lk * p.ldk + h * HD;
    const float* vb = p.v + (size_t)b * p.Lk * p.ldv + h * HD;
    const DropCtx drop = make_drop(p, b, h);
    stage_tile(gK, kb, p.ldk, 0, p.Lk);
    stage_tile(gV, vb, p.ldv, 0, p.Lk);
    stage_commit();
    uint32_t ndead = fetch_dead(p, b, 0);
    AFrag qa, ga;
    load_afrag(qa, qb, p.ldq, r0, p.Lq, lane, p.scale);
    load_afrag(ga, gb, p.ldo, r0, p.Lq, lane, 1.f);
    const int qi0 = r0 + g, qi1 = r0 + g + 8;
    const uint32_t rb0 = (uint32_t)qi0 * (uint32_t)p.Lk, rb1 = (uint32_t)qi1 * (uint32_t)p.Lk;
    const size_t st = ((size_t)b * p.H + h) * p.Lq;
    // lse = -inf (fully masked row) or a row past the end: +inf makes every p = exp(s - lse) exactly 0
    float lse0 = qi0 < p.Lq ? p.lse[st + qi0] : INFINITY, lse1 = qi1 < p.Lq ? p.lse[st + qi1] : INFINITY;
    if (lse0 == -INFINITY) lse0 = INFINITY;
    if (lse1 == -INFINITY) lse1 = INFINITY;
    const float dl0 = qi0 < p.Lq ? p.delta[st + qi0] : 0.f, dl1 = qi1 < p.Lq ? p.delta[st + qi1] : 0.f;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    for (int k0 = 0; k0 < p.Lk; k0 += BC) {
        stage_wait();
        __syncthreads();
        unstage_tile<true>(sKh, sKl, gK);
        unstage_tile<false>(sVh, nullptr, gV);
        if (threadIdx.x < BC) sDead[threadIdx.x] = (uint8_t)ndead;
        __syncthreads();
        if (k0 + BC < p.Lk) {
            stage_tile(gK, kb, p.ldk, k0 + BC, p.Lk);
            stage_tile(gV, vb, p.ldv, k0 + BC, p.Lk);
            stage_commit();
            ndead = fetch_dead(p, b, k0 + BC);
        }
        float s[8][4], dp[8][4];
        gemm_nt<3>(s, qa, sKh, sKl, lane);
        gemm_nt<1>(dp, ga, sVh, nullptr, lane);   // gradients: single-pass TF32 with round-to-nearest operands
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uchar2 dd = *reinterpret_cast<const uchar2*>(&sDead[nt * 8 + 2 * t]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = k0 + nt * 8 + 2 * t + e;
                const bool dead = e ? dd.y : dd.x;
                float p0 = dead ? 0.f : __expf(s[nt][e] - lse0);
                float p1 = dead ? 0.f : __expf(s[nt][2 + e] - lse1);
                float d0 = dp[nt][e], d1 = dp[nt][2 + e];
                if (drop.on) {
                    d0 *= keep_scale(drop, rb0, j);
                    d1 *= keep_scale(drop, rb1, j);
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
__global__ void __launch_bounds__(ATT_THREADS, 3)
attn_bwd_dkv_kernel(const AttnParams p) {
    __shared__ __align__(16) uint32_t sQh[BC][LDS], sQl[BC][LDS];
    __shared__ __align__(16) uint32_t sGh[BC][LDS];
    __shared__ __align__(16) float gQ[STG], gG[STG];
    __shared__ __align__(8) float sL[BC], sD[BC];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = blockIdx.x * BR + warp * 16;                      // first key row of this warp
    const float* qb = p.q + (size_t)b * p.Lq * p.ldq + h * HD;
    const float* gb = p.dout + (size_t)b * p.Lq * p.ldo + h * HD;
    const float* kb = p.k + (size_t)b * p.Lk * p.ldk + h * HD;
    const float* vb = p.v + (size_t)b * p.Lk * p.ldv + h * HD;
    const DropCtx drop = make_drop(p, b, h);
    const size_t st = ((size_t)b * p.H + h) * p.Lq;
    // per-thread row statistics of the streamed query tile (threads 0..63): lse (+inf -> p = 0) and delta
    auto fetch_stats = [&](int q0, float& l, float& d) {
        const int i = q0 + (int)threadIdx.x;
        l = INFINITY; d = 0.f;
        if (threadIdx.x < BC && i < p.Lq) {
            l = p.lse[st + i];
            d = p.delta[st + i];
            if (l == -INFINITY) l = INFINITY;
        }
    };
    float nl, nd;
    stage_tile(gQ, qb, p.ldq, 0, p.Lq);
    stage_tile(gG, gb, p.ldo, 0, p.Lq);
    stage_commit();
    fetch_stats(0, nl, nd);
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

    for (int q0 = 0; q0 < p.Lq; q0 += BC) {
        stage_wait();
        __syncthreads();
        unstage_tile<true>(sQh, sQl, gQ);
        unstage_tile<false>(sGh, nullptr, gG);
        if (threadIdx.x < BC) { sL[threadIdx.x] = nl; sD[threadIdx.x] = nd; }
        __syncthreads();
        if (q0 + BC < p.Lq) {
            stage_tile(gQ, qb, p.ldq, q0 + BC, p.Lq);
            stage_tile(gG, gb, p.ldo, q0 + BC, p.Lq);
            stage_commit();
            fetch_stats(q0 + BC, nl, nd);
        }
        float s[8][4], dp[8][4];
        gemm_nt<3>(s, ka, sQh, sQl, lane);    // S^T[key][query] (already scaled through K)
        gemm_nt<1>(dp, va, sGh, nullptr, lane);   // dP^T[key][query] = V . dO^T
        float pd[8][4];                        // dropped probabilities for dV
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float2 l2 = *reinterpret_cast<const float2*>(&sL[nt * 8 + 2 * t]);
            const float2 d2 = *reinterpret_cast<const float2*>(&sD[nt * 8 + 2 * t]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = q0 + nt * 8 + 2 * t + e;
                const float lse = e ? l2.y : l2.x, dl = e ? d2.y : d2.x;
                float p0 = dead0 ? 0.f : __expf(s[nt][e] - lse);
                float p1 = dead1 ? 0.f : __expf(s[nt][2 + e] - lse);
                float d0 = dp[nt][e], d1 = dp[nt][2 + e];
                float w0 = p0, w1 = p1;
                if (drop.on) {
                    const uint32_t rb = (uint32_t)i * (uint32_t)p.Lk;
                    const float k0s = keep_scale(drop, rb, kj0), k1s = keep_scale(drop, rb, kj1);
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
    if ((unsigned long long)p.Lq * (unsigned long long)p.Lk >= (1ull << 32)) return MDB_EUNSUPPORTED;   // 32-bit mask index
    return 0;
}

}  // namespace

// attention_tc.cu: tcgen05 / TMEM forward for long key sequences (MDB_EUNSUPPORTED = not applicable, use the kernel below)
int mdb_attention_forward_tc(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask, float* out,
                             float* lse, int B, int H, int Lq, int Lk, int ldq, int ldk, int ldv, int ldo, float drop_p,
                             const unsigned long long* seed, unsigned long long site, cudaStream_t stream);

// attention_tc_bwd.cu: tcgen05 / TMEM backward (dQ and dK/dV kernels) for long key sequences
int mdb_attention_backward_tc(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask, const float* lse,
                              const float* dout, const float* delta, float* dq, float* dk, float* dv, int B, int H, int Lq, int Lk,
                              int ldq, int ldk, int ldv, int ldo, int lddq, int lddk, int lddv, float drop_p,
                              const unsigned long long* seed, unsigned long long site, cudaStream_t stream);

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
    rc = mdb_attention_forward_tc(q, k, v, key_padding_mask, out, lse, B, H, Lq, Lk, ldq, ldk, ldv, ldo, drop_p, seed, site,
                                  static_cast<cudaStream_t>(stream));
    if (rc != MDB_EUNSUPPORTED) return rc;                          // ran (0) or failed: either way not the register kernel's call
    dim3 grid((Lq + BR - 1) / BR, H, B);
    constexpr int kFwdSmem = 4 * BC * LDS * 4 + 2 * STG * 4 + BC;
    static bool attr_set[64] = {};                 // the attribute is per (function, device): several devices per process
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev] = true;
    }
    attn_fwd_kernel<<<grid, ATT_THREADS, kFwdSmem, static_cast<cudaStream_t>(stream)>>>(p);
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
    rc = mdb_attention_backward_tc(q, k, v, key_padding_mask, lse, dout, delta_ws, dq, dk, dv, B, H, Lq, Lk, ldq, ldk, ldv, ldo, lddq,
                                   lddk, lddv, drop_p, seed, site, stream);
    if (rc != MDB_EUNSUPPORTED) return rc;
    attn_bwd_dq_kernel<<<dim3((Lq + BR - 1) / BR, H, B), ATT_THREADS, 0, stream>>>(p);
    attn_bwd_dkv_kernel<<<dim3((Lk + BR - 1) / BR, H, B), ATT_THREADS, 0, stream>>>(p);
    return (int)cudaGetLastError();
}

}  // extern "C"
