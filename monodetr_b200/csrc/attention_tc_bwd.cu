// attention_tc_bwd.cu -- tcgen05 / TMEM / TMA backward of the fused multi-head attention core (head dim 32), sm_100a.
// Same math as attn_bwd_dq_kernel / attn_bwd_dkv_kernel in attention.cu (which remain the path for short sequences):
//   P = exp(S - lse),  S = (Q / sqrt d) K^T          dP = dO V^T (x keep / (1 - p) under dropout)
//   dS = P (dP - delta),  delta_i = dO_i . O_i        dQ = dS K / sqrt d    dK = dS^T Q / sqrt d    dV = (P x keep)^T dO
// Every contraction is error-compensated BF16x3 (hi hi + lo hi + hi lo) on the tensor cores with fp32 accumulators in
// TENSOR MEMORY, the same arithmetic as the forward kernel (attention_tc.cu), so forward and backward see the same P.
//
// Two kernels (the two gradients need the score matrix in two orientations: the A operand of a tcgen05 MMA has its M rows
// on the TMEM lanes):
//   attn_bwd_dq_tc_kernel   CTA = 128 queries; per step a tile of 64 keys: S and dP (M = queries), dS -> dQ += dS K.
//   attn_bwd_dkv_tc_kernel  CTA = 128 keys; per step a tile of 64 queries: S^T and dP^T (M = keys), P^T, dS^T ->
//                           dV += P^T dO, dK += dS^T Q.
// Both keep ONE S / dP buffer (TMEM budget 256 columns) and rely on two co-resident CTAs per SM to overlap one CTA's
// tensor-core phase with the other's elementwise phase.  Warp roles per CTA (8 warps): 0 = TMA producer, 1 = MMA issuer,
// 2-3 = operand converters (fp32 tile rows -> bf16 (hi, lo) rows in place, plus the transposed K-major tiles the second
// GEMMs need), 4-7 = elementwise warps (thread = TMEM lane = row).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/monodetr_b200.h"
#include "rng.cuh"
#include "tc_common.cuh"
#include "tma_host.cuh"
#include "attention_tc_common.cuh"

namespace {

using namespace mdb;

constexpr int kRows = 128;                   // rows the CTA owns (queries in dQ, keys in dK/dV)
constexpr int kCols = 64;                    // streamed tile (keys in dQ, queries in dK/dV)
constexpr int kHD = 32;
constexpr int kBigTile = kRows * kHD * 4;    // 16 KiB
constexpr int kSmallTile = kCols * kHD * 4;  // 8 KiB
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct BwdParams {
    const uint8_t* kpm;            // [B][Lk] or null
    const float* lse;              // [B][H][Lq]
    const float* delta;            // [B][H][Lq]
    float *dq, *dk, *dv;
    int B, H, Lq, Lk, lddq, lddk, lddv;
    float scale, drop_p;
    const unsigned long long* seed;
    unsigned long long site;
};

__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }

// ---------------------------------------------------------------------------------------------------------------------
// dQ.  smem: Q [128][hi|lo], dO [128][hi|lo], stages x { K [64][hi|lo], V [64][hi|lo], K^T hi (4 KiB), K^T lo (4 KiB) }
// TMEM (256 columns): S/dS_hi @0 (64), dP @64 (64), dS_lo @128 (32), dQ @160 (32).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDqStages = 3;
constexpr int kDqStageBytes = 3 * kSmallTile;                    // 24 KiB
constexpr int kDqSmem = 2 * kBigTile + kDqStages * kDqStageBytes + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapG,
                      const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV,
                      const __grid_constant__ BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sG = smem + kBigTile;
    uint8_t* sStage = smem + 2 * kBigTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + kDqStages * kDqStageBytes);
    uint64_t* qg_full = bars;                  // TMA -> converters
    uint64_t* qg_ready = bars + 1;             // converters -> MMA (2 arrivals)
    uint64_t* kv_full = bars + 2;              // [stages]
    uint64_t* kv_ready = kv_full + kDqStages;  // [stages] (2 arrivals)
    uint64_t* kv_empty = kv_ready + kDqStages; // [stages] (commit)
    uint64_t* sdp_full = kv_empty + kDqStages; // S and dP of the tile are complete (commit)
    uint64_t* ds_ready = sdp_full + 1;         // dS written (4 arrivals)
    uint64_t* dq_done = ds_ready + 1;          // last dQ MMA complete (commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kRows;
    const int n_kv = (p.Lk + kCols - 1) / kCols;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapG); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV);
        mbar_init(qg_full, 1);
        mbar_init(qg_ready, 2);
        for (int s = 0; s < kDqStages; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_ready[s], 2);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(sdp_full, 1);
        mbar_init(ds_ready, 4);
        mbar_init(dq_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kColS = 0, kColDP = 64, kColDSlo = 128, kColDQ = 160;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(qg_full, 2 * kBigTile);
            tma_load_3d(sQ, &mapQ, qg_full, h * kHD, q0, b);
            tma_load_3d(sG, &mapG, qg_full, h * kHD, q0, b);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kDqStages;
                mbar_wait(&kv_empty[s], ((j / kDqStages) & 1) ^ 1);
                uint8_t* st = sStage + s * kDqStageBytes;
                mbar_arrive_expect_tx(&kv_full[s], 2 * kSmallTile);
                tma_load_3d(st, &mapK, &kv_full[s], h * kHD, j * kCols, b);
                tma_load_3d(st + kSmallTile, &mapV, &kv_full[s], h * kHD, j * kCols, b);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(kRows, kCols);
            constexpr uint32_t idesc_o = make_idesc_bf16(kRows, kHD);
            constexpr uint64_t kDescHi = make_smem_desc(0, 16, 1024, 2) & 0xFFFFFFFF00000000ull;
            constexpr uint32_t kDescLo = (uint32_t)(make_smem_desc(0, 16, 1024, 2) & 0xFFFF0000ull);
            auto desc = [&](uint32_t addr) { return kDescHi | (uint64_t)(kDescLo | ((addr >> 4) & 0x3FFFu)); };
            const uint32_t q_addr = smem_u32(sQ), g_addr = smem_u32(sG);
            mbar_wait(qg_ready, 0);
            tc_fence_after();
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kDqStages;
                mbar_wait(&kv_ready[s], (j / kDqStages) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sStage + s * kDqStageBytes), v_addr = k_addr + kSmallTile, kt_hi = k_addr + 2 * kSmallTile,
                               kt_lo = kt_hi + kSmallTile / 2;
                // S = Q K^T and dP = dO V^T (the previous tile's dS was consumed by MMAs issued before these: in-order pipe)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    umma_bf16(tmem_base + kColS, desc(q_addr + ks * 32), desc(k_addr + ks * 32), idesc_s, ks > 0);
                    umma_bf16(tmem_base + kColS, desc(q_addr + 64 + ks * 32), desc(k_addr + ks * 32), idesc_s, true);
                    umma_bf16(tmem_base + kColS, desc(q_addr + ks * 32), desc(k_addr + 64 + ks * 32), idesc_s, true);
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    umma_bf16(tmem_base + kColDP, desc(g_addr + ks * 32), desc(v_addr + ks * 32), idesc_s, ks > 0);
                    umma_bf16(tmem_base + kColDP, desc(g_addr + 64 + ks * 32), desc(v_addr + ks * 32), idesc_s, true);
                    umma_bf16(tmem_base + kColDP, desc(g_addr + ks * 32), desc(v_addr + 64 + ks * 32), idesc_s, true);
                }
                umma_commit(sdp_full);
                mbar_wait(ds_ready, j & 1);
                tc_fence_after();
                // dQ += dS K: A = dS (tensor memory, 8 columns per 16-key k-step), B = K^T tile [32 d][64 keys]
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_bf16_ta(tmem_base + kColDQ, tmem_base + kColS + ks * 8, desc(kt_hi + ks * 32), idesc_o, (j > 0) || (ks > 0));
                    umma_bf16_ta(tmem_base + kColDQ, tmem_base + kColDSlo + ks * 8, desc(kt_hi + ks * 32), idesc_o, true);
                    umma_bf16_ta(tmem_base + kColDQ, tmem_base + kColS + ks * 8, desc(kt_lo + ks * 32), idesc_o, true);
                }
                umma_commit(&kv_empty[s]);
            }
            umma_commit(dq_done);
        }
    } else if (warp < 4) {
        // converters: 64 threads.  Q (scaled) and dO once, two rows per thread; then K, V (+ K^T) of every stage, one row each
        const int t = (warp - 2) * 32 + lane;                      // 0..63
        const float gscale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
        mbar_wait(qg_full, 0);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int row = t + rr * 64;
            split_row_in_place(sQ + row * 128, row, p.scale);
            split_row_in_place(sG + row * 128, row, gscale);       // dO / (1 - p): dP arrives with the dropout scale folded in
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(qg_ready);
        for (int j = 0; j < n_kv; ++j) {
            const int s = j % kDqStages;
            mbar_wait(&kv_full[s], (j / kDqStages) & 1);
            uint8_t* st = sStage + s * kDqStageBytes;
            transpose_row_bf16(st + t * 128, t, st + 2 * kSmallTile, kSmallTile / 2);      // K^T (before K is rewritten)
            split_row_in_place(st + t * 128, t, 1.f);
            split_row_in_place(st + kSmallTile + t * 128, t, 1.f);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_ready[s]);
        }
    } else {
        // elementwise warps: thread = query row
        const int qtr = warp & 3;
        const int r = qtr * 32 + lane;
        const int qi = q0 + r;
        const uint32_t lane_off = (uint32_t)(qtr * 32) << 16;
        const bool drop_on = p.drop_p > 0.f;
        const uint32_t thr32 = rng_thr16(p.drop_p) << 16;
        uint32_t key = 0u;
        if (drop_on) key = rng_key32(*p.seed + p.site * 0x9E3779B97F4A7C15ull, (unsigned long long)(b * p.H + h));
        const uint32_t row_base = (uint32_t)qi * (uint32_t)p.Lk;
        const uint8_t* kpm = p.kpm ? p.kpm + (size_t)b * p.Lk : nullptr;
        const size_t stat = ((size_t)b * p.H + h) * p.Lq;
        float lse = qi < p.Lq ? p.lse[stat + qi] : INFINITY;
        if (lse == -INFINITY) lse = INFINITY;                      // fully masked row: every p = 0
        const float dl = qi < p.Lq ? p.delta[stat + qi] : 0.f;
        const float mneg = (lse == INFINITY) ? -INFINITY : -lse * kLog2e;
        const bool paired = drop_on && ((row_base & 1u) == 0u);
        for (int j = 0; j < n_kv; ++j) {
            const int k0 = j * kCols;
            mbar_wait(sdp_full, j & 1);
            tc_fence_after();
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(tmem_base + kColS + lane_off, sv);
            tmem_ld_32x32(tmem_base + kColDP + lane_off, dv);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tmem_ld_wait();
                uint32_t dead = 0u;
                const int kb = k0 + c * 32;
                if (kb + 32 > p.Lk) dead = (kb >= p.Lk) ? 0xFFFFFFFFu : (0xFFFFFFFFu << (p.Lk - kb));
                if (kpm && kb < p.Lk) {
                    const int ne = min(32, p.Lk - kb);
                    for (int e = 0; e < ne; ++e) dead |= (kpm[kb + e] ? 1u : 0u) << e;
                }
                const uint32_t jb = row_base + (uint32_t)kb;
                uint32_t hw[16], lw[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), kLog2e, mneg));
                    float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), kLog2e, mneg));
                    if (dead) {
                        if ((dead >> (2 * e)) & 1u) p0 = 0.f;
                        if ((dead >> (2 * e + 1)) & 1u) p1 = 0.f;
                    }
                    float d0 = __uint_as_float(dv[2 * e]), d1 = __uint_as_float(dv[2 * e + 1]);
                    if (drop_on) {      // (dP already carries 1 / (1 - p): dO was scaled when it was split)
                        bool k0b, k1b;
                        if (paired) {
                            const uint32_t hsh = rng_pair32(key, (jb >> 1) + (uint32_t)e);
                            k0b = hsh >= thr32;
                            k1b = rng_pair32_odd(hsh) >= thr32;
                        } else {
                            k0b = rng_keep16(key, jb + 2u * e, thr32 >> 16);
                            k1b = rng_keep16(key, jb + 2u * e + 1u, thr32 >> 16);
                        }
                        d0 = k0b ? d0 : 0.f;
                        d1 = k1b ? d1 : 0.f;
                    }
                    const float s0 = p0 * (d0 - dl), s1 = p1 * (d1 - dl);
                    const uint32_t hh = pack_bf16x2(s0, s1);
                    hw[e] = hh;
                    lw[e] = pack_bf16x2(s0 - __uint_as_float(hh << 16), s1 - __uint_as_float(hh & 0xFFFF0000u));
                }
                if (c == 0) {   // next chunk's loads fly while this chunk's results are stored (sv / dv are dead by now)
                    tmem_ld_32x32(tmem_base + kColS + 32 + lane_off, sv);
                    tmem_ld_32x32(tmem_base + kColDP + 32 + lane_off, dv);
                }
                tmem_st_32x16(tmem_base + kColS + c * 16 + lane_off, hw);
                tmem_st_32x16(tmem_base + kColDSlo + c * 16 + lane_off, lw);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_ready);
        }
        mbar_wait(dq_done, 0);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32(tmem_base + kColDQ + lane_off, o);
        tmem_ld_wait();
        if (qi < p.Lq) {
            float* dst = p.dq + ((size_t)b * p.Lq + qi) * p.lddq + h * kHD;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                *reinterpret_cast<float4*>(dst + 4 * e) = make_float4(__uint_as_float(o[4 * e]) * p.scale, __uint_as_float(o[4 * e + 1]) * p.scale,
                                                                      __uint_as_float(o[4 * e + 2]) * p.scale, __uint_as_float(o[4 * e + 3]) * p.scale);
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dK / dV.  smem: K [128][hi|lo] (scaled), V [128][hi|lo], stages x { Q [64][hi|lo], dO [64][hi|lo], Q^T hi|lo (8 KiB),
// dO^T hi|lo (8 KiB), lse[64], delta[64] }.
// TMEM (256 columns): S^T / P^T_hi @0 (64), dP^T / dS^T_hi @64 (64), P^T_lo @128 (32), dS^T_lo @160 (32), dK @192, dV @224.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kKvStages = 2;
constexpr int kKvStageBytes = 4 * kSmallTile + 1024;             // 33 KiB (statistics in the last KiB)
constexpr int kKvSmem = 2 * kBigTile + kKvStages * kKvStageBytes + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapG,
                       const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV,
                       const __grid_constant__ BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sK = smem;
    uint8_t* sV = smem + kBigTile;
    uint8_t* sStage = smem + 2 * kBigTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + kKvStages * kKvStageBytes);
    uint64_t* kv_full = bars;
    uint64_t* kv_ready = bars + 1;             // (2 arrivals)
    uint64_t* qg_full = bars + 2;              // [stages]
    uint64_t* qg_ready = qg_full + kKvStages;  // [stages] (2 arrivals)
    uint64_t* qg_empty = qg_ready + kKvStages; // [stages] (commit)
    uint64_t* sdp_full = qg_empty + kKvStages;
    uint64_t* ds_ready = sdp_full + 1;         // (4 arrivals)
    uint64_t* acc_done = ds_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, kbase = blockIdx.x * kRows;
    const int n_q = (p.Lq + kCols - 1) / kCols;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapG); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV);
        mbar_init(kv_full, 1);
        mbar_init(kv_ready, 2);
        for (int s = 0; s < kKvStages; ++s) {
            mbar_init(&qg_full[s], 1);
            mbar_init(&qg_ready[s], 2);
            mbar_init(&qg_empty[s], 1);
        }
        mbar_init(sdp_full, 1);
        mbar_init(ds_ready, 4);
        mbar_init(acc_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kColS = 0, kColDP = 64, kColPlo = 128, kColDSlo = 160, kColDK = 192, kColDV = 224;
    const size_t stat = ((size_t)b * p.H + h) * p.Lq;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(kv_full, 2 * kBigTile);
            tma_load_3d(sK, &mapK, kv_full, h * kHD, kbase, b);
            tma_load_3d(sV, &mapV, kv_full, h * kHD, kbase, b);
            for (int j = 0; j < n_q; ++j) {
                const int s = j % kKvStages;
                mbar_wait(&qg_empty[s], ((j / kKvStages) & 1) ^ 1);
                uint8_t* st = sStage + s * kKvStageBytes;
                mbar_arrive_expect_tx(&qg_full[s], 2 * kSmallTile);
                tma_load_3d(st, &mapQ, &qg_full[s], h * kHD, j * kCols, b);
                tma_load_3d(st + kSmallTile, &mapG, &qg_full[s], h * kHD, j * kCols, b);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(kRows, kCols);
            constexpr uint32_t idesc_o = make_idesc_bf16(kRows, kHD);
            constexpr uint64_t kDescHi = make_smem_desc(0, 16, 1024, 2) & 0xFFFFFFFF00000000ull;
            constexpr uint32_t kDescLo = (uint32_t)(make_smem_desc(0, 16, 1024, 2) & 0xFFFF0000ull);
            auto desc = [&](uint32_t addr) { return kDescHi | (uint64_t)(kDescLo | ((addr >> 4) & 0x3FFFu)); };
            const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            mbar_wait(kv_ready, 0);
            tc_fence_after();
            for (int j = 0; j < n_q; ++j) {
                const int s = j % kKvStages;
                mbar_wait(&qg_ready[s], (j / kKvStages) & 1);
                tc_fence_after();
                const uint32_t q_addr = smem_u32(sStage + s * kKvStageBytes), g_addr = q_addr + kSmallTile, qt_hi = q_addr + 2 * kSmallTile,
                               qt_lo = qt_hi + kSmallTile / 2, gt_hi = q_addr + 3 * kSmallTile, gt_lo = gt_hi + kSmallTile / 2;
                // S^T = K Q^T and dP^T = V dO^T (M = keys)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    umma_bf16(tmem_base + kColS, desc(k_addr + ks * 32), desc(q_addr + ks * 32), idesc_s, ks > 0);
                    umma_bf16(tmem_base + kColS, desc(k_addr + 64 + ks * 32), desc(q_addr + ks * 32), idesc_s, true);
                    umma_bf16(tmem_base + kColS, desc(k_addr + ks * 32), desc(q_addr + 64 + ks * 32), idesc_s, true);
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    umma_bf16(tmem_base + kColDP, desc(v_addr + ks * 32), desc(g_addr + ks * 32), idesc_s, ks > 0);
                    umma_bf16(tmem_base + kColDP, desc(v_addr + 64 + ks * 32), desc(g_addr + ks * 32), idesc_s, true);
                    umma_bf16(tmem_base + kColDP, desc(v_addr + ks * 32), desc(g_addr + 64 + ks * 32), idesc_s, true);
                }
                umma_commit(sdp_full);
                mbar_wait(ds_ready, j & 1);
                tc_fence_after();
                // dV += P^T dO (B = dO^T tile [32 d][64 queries]);  dK += dS^T Q (B = Q^T tile)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_bf16_ta(tmem_base + kColDV, tmem_base + kColS + ks * 8, desc(gt_hi + ks * 32), idesc_o, (j > 0) || (ks > 0));
                    umma_bf16_ta(tmem_base + kColDV, tmem_base + kColPlo + ks * 8, desc(gt_hi + ks * 32), idesc_o, true);
                    umma_bf16_ta(tmem_base + kColDV, tmem_base + kColS + ks * 8, desc(gt_lo + ks * 32), idesc_o, true);
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_bf16_ta(tmem_base + kColDK, tmem_base + kColDP + ks * 8, desc(qt_hi + ks * 32), idesc_o, (j > 0) || (ks > 0));
                    umma_bf16_ta(tmem_base + kColDK, tmem_base + kColDSlo + ks * 8, desc(qt_hi + ks * 32), idesc_o, true);
                    umma_bf16_ta(tmem_base + kColDK, tmem_base + kColDP + ks * 8, desc(qt_lo + ks * 32), idesc_o, true);
                }
                umma_commit(&qg_empty[s]);
            }
            umma_commit(acc_done);
        }
    } else if (warp < 4) {
        const int t = (warp - 2) * 32 + lane;                      // 0..63
        const float gscale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
        mbar_wait(kv_full, 0);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int row = t + rr * 64;
            split_row_in_place(sK + row * 128, row, p.scale);
            split_row_in_place(sV + row * 128, row, 1.f);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(kv_ready);
        for (int j = 0; j < n_q; ++j) {
            const int s = j % kKvStages;
            // the statistics slots of this stage are free once the stage is (the producer waited qg_empty before refilling)
            mbar_wait(&qg_full[s], (j / kKvStages) & 1);
            uint8_t* st = sStage + s * kKvStageBytes;
            float* stats = reinterpret_cast<float*>(st + 4 * kSmallTile);
            const int qi = j * kCols + t;
            float l = INFINITY, d = 0.f;
            if (qi < p.Lq) {
                l = p.lse[stat + qi];
                d = p.delta[stat + qi];
                if (l == -INFINITY) l = INFINITY;
            }
            stats[t] = (l == INFINITY) ? -INFINITY : -l * kLog2e;   // exp2 offset of query t
            stats[64 + t] = d;
            transpose_row_bf16(st + t * 128, t, st + 2 * kSmallTile, kSmallTile / 2);                // Q^T
            transpose_row_bf16(st + kSmallTile + t * 128, t, st + 3 * kSmallTile, kSmallTile / 2);   // dO^T (unscaled: dV is scaled at the end)
            split_row_in_place(st + t * 128, t, 1.f);
            split_row_in_place(st + kSmallTile + t * 128, t, gscale);   // dO / (1 - p) -> dP^T carries the dropout scale
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&qg_ready[s]);
        }
    } else {
        // elementwise warps: thread = key row
        const int qtr = warp & 3;
        const int r = qtr * 32 + lane;
        const int kj = kbase + r;
        const uint32_t lane_off = (uint32_t)(qtr * 32) << 16;
        const bool drop_on = p.drop_p > 0.f;
        const float inv_keep = drop_on ? 1.f / (1.f - p.drop_p) : 1.f;
        const uint32_t thr16 = rng_thr16(p.drop_p);
        uint32_t key = 0u;
        if (drop_on) key = rng_key32(*p.seed + p.site * 0x9E3779B97F4A7C15ull, (unsigned long long)(b * p.H + h));
        // keep(query i, key kj) is element i * Lk + kj of the mask: with Lk even, lanes 2m and 2m+1 (keys kj, kj + 1) share the
        // pair word, so every lane hashes only the query columns of its own parity and fetches the others from its neighbour
        const bool share = drop_on && ((p.Lk & 1) == 0) && ((kbase & 1) == 0);
        const bool dead = kj >= p.Lk || (p.kpm && p.kpm[(size_t)b * p.Lk + min(kj, p.Lk - 1)]);
        for (int j = 0; j < n_q; ++j) {
            const int s = j % kKvStages;
            const float* stats = reinterpret_cast<const float*>(sStage + s * kKvStageBytes + 4 * kSmallTile);
            mbar_wait(sdp_full, j & 1);                            // (implies qg_ready[s]: the statistics are in place)
            tc_fence_after();
            mbar_wait(&qg_ready[s], (j / kKvStages) & 1);         // (acquire the converters' statistics writes explicitly)
            uint32_t sv[32], dv[32];
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                tmem_ld_32x32(tmem_base + kColS + c * 32 + lane_off, sv);
                tmem_ld_32x32(tmem_base + kColDP + c * 32 + lane_off, dv);
                tmem_ld_wait();
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                   // 16 query columns at a time: register budget 128
                    uint32_t pw[8], plw[8], dw[8], dlw[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int ee = hf * 8 + e;
                        const int i0 = c * 32 + 2 * ee;             // query column inside the tile
                        const float2 mn = *reinterpret_cast<const float2*>(stats + i0);
                        const float2 dl = *reinterpret_cast<const float2*>(stats + 64 + i0);
                        float p0 = dead ? 0.f : ex2_approx(fmaf(__uint_as_float(sv[2 * ee]), kLog2e, mn.x));
                        float p1 = dead ? 0.f : ex2_approx(fmaf(__uint_as_float(sv[2 * ee + 1]), kLog2e, mn.y));
                        float d0 = __uint_as_float(dv[2 * ee]), d1 = __uint_as_float(dv[2 * ee + 1]);
                        float w0 = p0, w1 = p1;
                        if (drop_on) {      // (1 / (1 - p): dP^T carries it already, dV gets it in the epilogue)
                            const uint32_t qa = (uint32_t)(j * kCols + i0);
                            bool k0b, k1b;
                            if (share) {
                                // this lane hashes column i0 + (lane & 1), the neighbour the other one: word of pair (kj >> 1)
                                const uint32_t mine = rng_pair32(key, ((qa + (uint32_t)(lane & 1)) * (uint32_t)p.Lk + (uint32_t)kj) >> 1);
                                const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                                const uint32_t h0 = (lane & 1) ? other : mine, h1 = (lane & 1) ? mine : other;
                                k0b = ((lane & 1) ? rng_pair32_odd(h0) : h0) >= (thr16 << 16);
                                k1b = ((lane & 1) ? rng_pair32_odd(h1) : h1) >= (thr16 << 16);
                            } else {
                                k0b = rng_keep16(key, qa * (uint32_t)p.Lk + (uint32_t)kj, thr16);
                                k1b = rng_keep16(key, (qa + 1u) * (uint32_t)p.Lk + (uint32_t)kj, thr16);
                            }
                            w0 = k0b ? w0 : 0.f; d0 = k0b ? d0 : 0.f;
                            w1 = k1b ? w1 : 0.f; d1 = k1b ? d1 : 0.f;
                        }
                        const float s0 = p0 * (d0 - dl.x), s1 = p1 * (d1 - dl.y);
                        const uint32_t ph = pack_bf16x2(w0, w1);
                        pw[e] = ph;
                        plw[e] = pack_bf16x2(w0 - __uint_as_float(ph << 16), w1 - __uint_as_float(ph & 0xFFFF0000u));
                        const uint32_t dh = pack_bf16x2(s0, s1);
                        dw[e] = dh;
                        dlw[e] = pack_bf16x2(s0 - __uint_as_float(dh << 16), s1 - __uint_as_float(dh & 0xFFFF0000u));
                    }
                    // packed results of columns [32 c + 16 hf, +16) -> 8 TMEM columns each (in place for the hi parts: those
                    // columns lie below everything this thread has still to read)
                    const uint32_t col = (uint32_t)(c * 16 + hf * 8) + lane_off;
                    tmem_st_32x8(tmem_base + kColS + col, pw);
                    tmem_st_32x8(tmem_base + kColPlo + col, plw);
                    tmem_st_32x8(tmem_base + kColDP + col, dw);
                    tmem_st_32x8(tmem_base + kColDSlo + col, dlw);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_ready);
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        uint32_t ok[32], ov[32];
        tmem_ld_32x32(tmem_base + kColDK + lane_off, ok);
        tmem_ld_32x32(tmem_base + kColDV + lane_off, ov);
        tmem_ld_wait();
        if (kj < p.Lk) {
            float* dkp = p.dk + ((size_t)b * p.Lk + kj) * p.lddk + h * kHD;
            float* dvp = p.dv + ((size_t)b * p.Lk + kj) * p.lddv + h * kHD;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                *reinterpret_cast<float4*>(dkp + 4 * e) = make_float4(__uint_as_float(ok[4 * e]) * p.scale, __uint_as_float(ok[4 * e + 1]) * p.scale,
                                                                      __uint_as_float(ok[4 * e + 2]) * p.scale, __uint_as_float(ok[4 * e + 3]) * p.scale);
                *reinterpret_cast<float4*>(dvp + 4 * e) = make_float4(__uint_as_float(ov[4 * e]) * inv_keep, __uint_as_float(ov[4 * e + 1]) * inv_keep,
                                                                      __uint_as_float(ov[4 * e + 2]) * inv_keep, __uint_as_float(ov[4 * e + 3]) * inv_keep);
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

int make_tok_map(CUtensorMap* m, const float* base, int E, int L, int B, int ld, int rows) {
    uint64_t dims[3] = {(uint64_t)E, (uint64_t)L, (uint64_t)B};
    uint64_t str[3] = {1, (uint64_t)ld, (uint64_t)L * ld};
    uint32_t box[3] = {kHD, (uint32_t)rows, 1};
    return make_map(m, base, 3, dims, str, box, nullptr);
}

}  // namespace

// Launcher used by mdb_attention_backward_f32 (attention.cu) after the delta kernel.  MDB_EUNSUPPORTED = not applicable.
int mdb_attention_backward_tc(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask, const float* lse,
                              const float* dout, const float* delta, float* dq, float* dk, float* dv, int B, int H, int Lq, int Lk,
                              int ldq, int ldk, int ldv, int ldo, int lddq, int lddk, int lddv, float drop_p,
                              const unsigned long long* seed, unsigned long long site, cudaStream_t stream) {
    static const bool disabled = getenv("MDB_ATTN_LEGACY") != nullptr || getenv("MDB_ATTN_BWD_LEGACY") != nullptr;
    if (disabled || Lk < 256) return MDB_EUNSUPPORTED;
    if ((ldq | ldk | ldv | ldo | lddq | lddk | lddv) % 4) return MDB_EUNSUPPORTED;
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15u) return MDB_EUNSUPPORTED;
    const int E = H * kHD;
    BwdParams p;
    memset(&p, 0, sizeof(p));
    p.kpm = key_padding_mask; p.lse = lse; p.delta = delta; p.dq = dq; p.dk = dk; p.dv = dv;
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    p.scale = 1.f / sqrtf((float)kHD); p.drop_p = drop_p; p.seed = seed; p.site = site;
    static bool configured[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKvSmem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    {
        CUtensorMap mq, mg, mk, mv;
        if (make_tok_map(&mq, q, E, Lq, B, ldq, kRows) || make_tok_map(&mg, dout, E, Lq, B, ldo, kRows) ||
            make_tok_map(&mk, k, E, Lk, B, ldk, kCols) || make_tok_map(&mv, v, E, Lk, B, ldv, kCols))
            return MDB_EUNSUPPORTED;
        attn_bwd_dq_tc_kernel<<<dim3((Lq + kRows - 1) / kRows, H, B), kThreads, kDqSmem, stream>>>(mq, mg, mk, mv, p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    {
        CUtensorMap mq, mg, mk, mv;
        if (make_tok_map(&mq, q, E, Lq, B, ldq, kCols) || make_tok_map(&mg, dout, E, Lq, B, ldo, kCols) ||
            make_tok_map(&mk, k, E, Lk, B, ldk, kRows) || make_tok_map(&mv, v, E, Lk, B, ldv, kRows))
            return MDB_EUNSUPPORTED;
        attn_bwd_dkv_tc_kernel<<<dim3((Lk + kRows - 1) / kRows, H, B), kThreads, kKvSmem, stream>>>(mq, mg, mk, mv, p);
        return (int)cudaGetLastError();
    }
}
