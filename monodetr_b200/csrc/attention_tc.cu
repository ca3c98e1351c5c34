// attention_tc.cu -- tcgen05 / TMEM / TMA forward of the fused multi-head attention core (head dim 32), sm_100a.
// Same contract as attn_fwd_kernel in attention.cu (softmax(Q K^T / sqrt(d)) V, key_padding_mask, hash dropout on the
// probabilities, lse for the backward) for the long-key call sites of the model: the depth encoder's self-attention
// (depth_predictor/transformer.py:59, 1920 x 1920 per image and head) and the decoder's depth cross-attention
// (depthaware_transformer.py:456-459, 550 x 1920).
//
// One CTA = 128 queries of one (image, head); it walks the keys in tiles of 128.
//   S = (Q / sqrt(d)) K^T   error-compensated BF16x3 (Q_hi K_hi + Q_lo K_hi + Q_hi K_lo), kind::f16 MMAs M128 N128 K16,
//                           both operands K-major [hi 32 | lo 32] bf16 rows in shared memory (converted in place from the
//                           fp32 tiles TMA delivered), fp32 accumulator in TENSOR MEMORY (two S buffers).
//   softmax                 4 warps, thread = query row = TMEM lane: row max and sum need no shuffles.  Online rescaling
//                           (running max m, running sum l); P is written back to tensor memory as packed bf16 (hi, lo) pairs,
//                           hi IN PLACE over the scores it came from.
//   O += P V                BF16x3 again (P_hi V_hi + P_lo V_hi + P_hi V_lo), A = P from tensor memory, B = V^T as K-major
//                           bf16 tiles the converter warps transpose out of the fp32 V tile; O (128 x 32 fp32) lives in tensor
//                           memory and is rescaled there when the running max moves.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one elected lane), 2-5 = operand converters, 6-9 and 10-13 = two softmax
// SETS.  Set x owns the key tiles j = x (mod 2) with its own S/P buffers, its own O accumulator and its own running
// (m, l): the two sets ping-pong on the tensor pipe without ever synchronising per tile (the exponentials, not the MMAs,
// bound the kernel: 16 MUFU.EX2 per clock per SM, and one warp per scheduler cannot hide their latency); the partial
// results are merged once at the end like a split-KV reduction.  The running reference m only moves when a tile's
// maximum exceeds it by more than 8 (p <= e^8: harmless in fp32 / split bf16), so O is rescaled a couple of times per
// row instead of once per tile; the arithmetic stays exact (l and O are rescaled together).
// Pipelines: 3-stage K/V ring (full / ready / empty mbarriers), per-set s_full / p_ready / pv_done.
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

constexpr int kTile = 128;                   // queries per CTA = keys per step
constexpr int kHD = 32;
constexpr int kTileBytes = kTile * kHD * 4;  // 16 KiB: 128 rows x 128 B
constexpr int kStages = 3;
constexpr int kStageBytes = 3 * kTileBytes;  // K (fp32 -> bf16 [hi|lo] in place) | V fp32 | V^T bf16 hi (8 KiB) + lo (8 KiB)
constexpr int kThreads = 448;
constexpr int kMergeFloats = 36;             // per row: m, l, pad, pad, O[32] (16-byte aligned rows)
constexpr float kLog2e = 1.4426950408889634f;

struct TcAttnParams {
    const uint8_t* kpm;            // [B][Lk] or null
    float* out;
    float* lse;                    // [B][H][Lq]
    int B, H, Lq, Lk, ldo;
    float scale, drop_p;
    const unsigned long long* seed;
    unsigned long long site;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const __grid_constant__ TcAttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                                            // 16 KiB
    uint8_t* sStage = smem + kTileBytes;                           // kStages x 48 KiB
    float* sMerge = reinterpret_cast<float*>(smem + kTileBytes + kStages * kStageBytes);   // [128][kMergeFloats]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTileBytes + kStages * kStageBytes + kTile * kMergeFloats * 4);
    uint64_t* q_full = bars;                   // TMA -> converters
    uint64_t* q_ready = bars + 1;              // converters -> MMA (4 arrivals)
    uint64_t* kv_full = bars + 2;              // [kStages]
    uint64_t* kv_ready = kv_full + kStages;    // [kStages] (4 arrivals)
    uint64_t* kv_empty = kv_ready + kStages;   // [kStages] (tcgen05.commit)
    uint64_t* s_full = kv_empty + kStages;     // [2] (tcgen05.commit)
    uint64_t* p_ready = s_full + 2;            // [2] (4 arrivals)
    uint64_t* pv_done = p_ready + 2;           // [2] (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kTile;
    const int n_kv = (p.Lk + kTile - 1) / kTile;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapK);
        tma_prefetch_desc(&mapV);
        mbar_init(q_full, 1);
        mbar_init(q_ready, 4);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_ready[s], 4);
            mbar_init(&kv_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tensor-memory columns (x = softmax set): S/P_hi[x] at 128 x, P_lo[x] at 256 + 64 x, O[x] at 384 + 32 x
    constexpr uint32_t kColPlo = 256, kColO = 384;

    if (warp == 0) {
        // ================================ TMA producer =============================================
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, kTileBytes);
            tma_load_3d(sQ, &mapQ, q_full, h * kHD, q0, b);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kStages;
                mbar_wait(&kv_empty[s], ((j / kStages) & 1) ^ 1);
                uint8_t* st = sStage + s * kStageBytes;
                mbar_arrive_expect_tx(&kv_full[s], 2 * kTileBytes);
                tma_load_3d(st, &mapK, &kv_full[s], h * kHD, j * kTile, b);
                tma_load_3d(st + kTileBytes, &mapV, &kv_full[s], h * kHD, j * kTile, b);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==============================================
        if (elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(kTile, kTile);
            constexpr uint32_t idesc_o = make_idesc_bf16(kTile, kHD);
            constexpr uint64_t kDescHi = make_smem_desc(0, 16, 1024, 2) & 0xFFFFFFFF00000000ull;     // K-major, SWIZZLE_128B
            constexpr uint32_t kDescLo = (uint32_t)(make_smem_desc(0, 16, 1024, 2) & 0xFFFF0000ull);
            auto desc = [&](uint32_t addr) { return kDescHi | (uint64_t)(kDescLo | ((addr >> 4) & 0x3FFFu)); };
            const uint32_t q_addr = smem_u32(sQ);
            mbar_wait(q_ready, 0);
            tc_fence_after();
            auto issue_s = [&](int j) {
                const int s = j % kStages;
                mbar_wait(&kv_ready[s], (j / kStages) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sStage + s * kStageBytes);
                const uint32_t ts = tmem_base + (uint32_t)(j & 1) * kTile;
                // rows are [hi 32 | lo 32] bf16: k-step ks (16 elements) of hi at byte 32 ks, of lo at 64 + 32 ks
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    umma_bf16(ts, desc(q_addr + ks * 32), desc(k_addr + ks * 32), idesc_s, ks > 0);
                    umma_bf16(ts, desc(q_addr + 64 + ks * 32), desc(k_addr + ks * 32), idesc_s, true);
                    umma_bf16(ts, desc(q_addr + ks * 32), desc(k_addr + 64 + ks * 32), idesc_s, true);
                }
                umma_commit(&s_full[j & 1]);
            };
            issue_s(0);
            if (n_kv > 1) issue_s(1);
            for (int j = 0; j < n_kv; ++j) {
                const int x = j & 1;
                mbar_wait(&p_ready[x], (j >> 1) & 1);
                tc_fence_after();
                const int s = j % kStages;
                const uint32_t vt_hi = smem_u32(sStage + s * kStageBytes + 2 * kTileBytes), vt_lo = vt_hi + kTileBytes / 2;
                const uint32_t tp_hi = tmem_base + (uint32_t)x * kTile, tp_lo = tmem_base + kColPlo + (uint32_t)x * 64;
                const uint32_t to = tmem_base + kColO + (uint32_t)x * kHD;
                // V^T tiles: [2 atoms of 64 keys][32 rows d][128 B]; k-step ks = 16 keys = 32 B inside atom ks / 4;
                // P (A operand, tensor memory): 8 columns per k-step (two bf16 per 32-bit column)
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t off = (uint32_t)(ks >> 2) * 4096u + (uint32_t)(ks & 3) * 32u;
                    umma_bf16_ta(to, tp_hi + ks * 8, desc(vt_hi + off), idesc_o, (j > 1) || (ks > 0));
                    umma_bf16_ta(to, tp_lo + ks * 8, desc(vt_hi + off), idesc_o, true);
                    umma_bf16_ta(to, tp_hi + ks * 8, desc(vt_lo + off), idesc_o, true);
                }
                umma_commit(&kv_empty[s]);                         // K and V of this stage are consumed
                umma_commit(&pv_done[x]);
                if (j + 2 < n_kv) issue_s(j + 2);                  // into the S/P buffer P(j) just left (the tensor pipe is in order)
            }
        }
    } else if (warp < 6) {
        // ================================ operand converters ======================================
        const int row = (warp - 2) * 32 + lane;                   // tile row (query / key) this thread owns
        mbar_wait(q_full, 0);
        split_row_in_place(sQ + row * 128, row, p.scale);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(q_ready);
        for (int j = 0; j < n_kv; ++j) {
            const int s = j % kStages;
            mbar_wait(&kv_full[s], (j / kStages) & 1);
            uint8_t* st = sStage + s * kStageBytes;
            split_row_in_place(st + row * 128, row, 1.f);
            // V row (key `row`): 32 fp32 -> column `row` of the K-major V^T tiles (bf16 hi and lo)
            transpose_row_bf16(st + kTileBytes + row * 128, row, st + 2 * kTileBytes, kTileBytes / 2);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_ready[s]);
        }
    } else {
        // ================================ softmax / epilogue warps ================================
        const int x = (warp - 6) >> 2;                             // softmax set: key tiles j = x (mod 2)
        const int qtr = warp & 3;                                  // TMEM lane quarter this warp may access
        const int r = qtr * 32 + lane;
        const int qi = q0 + r;
        const uint32_t lane_off = (uint32_t)(qtr * 32) << 16;
        const bool drop_on = p.drop_p > 0.f;
        const float inv_keep = 1.f / (1.f - p.drop_p);
        const uint32_t thr32 = rng_thr16(p.drop_p) << 16;
        uint32_t key = 0u;
        if (drop_on) key = rng_key32(*p.seed + p.site * 0x9E3779B97F4A7C15ull, (unsigned long long)(b * p.H + h));
        const uint32_t row_base = (uint32_t)qi * (uint32_t)p.Lk;
        const uint8_t* kpm = p.kpm ? p.kpm + (size_t)b * p.Lk : nullptr;
        const uint32_t ts = tmem_base + (uint32_t)x * kTile + lane_off;
        const uint32_t tpl = tmem_base + kColPlo + (uint32_t)x * 64 + lane_off;
        const uint32_t to = tmem_base + kColO + (uint32_t)x * kHD + lane_off;
        float m_ref = -INFINITY, l_run = 0.f;
        int it = 0;
        for (int j = x; j < n_kv; j += 2, ++it) {
            const int k0 = j * kTile;
            mbar_wait(&s_full[x], it & 1);
            tc_fence_after();
            // dead keys of this tile (past the end / key_padding_mask): one bit per key; all zero on the model's path
            uint32_t dead[4] = {0u, 0u, 0u, 0u};
            if (kpm || k0 + kTile > p.Lk) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t bits = 0u;
                    const int kb = k0 + c * 32;
                    if (kb + 32 > p.Lk) bits = (kb >= p.Lk) ? 0xFFFFFFFFu : (0xFFFFFFFFu << (p.Lk - kb));
                    if (kpm && kb < p.Lk) {
                        if (kb + 32 <= p.Lk && ((reinterpret_cast<uintptr_t>(kpm + kb) & 15u) == 0)) {
                            // 32 mask bytes -> 32 bits: per 4-byte word, non-zero bytes -> bit 7 of each byte -> one nibble
                            const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(kpm + kb));
                            const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(kpm + kb) + 1);
                            const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t nz = ((w[i] | ((w[i] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u) >> 7;
                                bits |= ((nz * 0x10204080u) >> 28) << (4 * i);
                            }
                        } else {
                            const int ne = min(32, p.Lk - kb);
                            for (int e = 0; e < ne; ++e) bits |= (kpm[kb + e] ? 1u : 0u) << e;
                        }
                    }
                    dead[c] = bits;
                }
            }
            // pass 1: row maximum of the tile
            // (tensor-memory loads are software-pipelined in both passes: chunk c + 1 is in flight while chunk c is processed;
            // with the wait right behind every load the ~300-cycle TMEM latency was exposed 8 times per tile)
            float mx = -INFINITY;
            uint32_t v[2][32];
            tmem_ld_32x32(ts, v[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld_wait();
                if (c < 3) tmem_ld_32x32(ts + (c + 1) * 32, v[(c + 1) & 1]);
                else tmem_ld_32x32(ts, v[0]);                              // first chunk of pass 2
                const uint32_t dd = dead[c];
                const uint32_t* vc = v[c & 1];
                if (dd == 0u) {
                    float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        a0 = fmaxf(a0, __uint_as_float(vc[e]));     a1 = fmaxf(a1, __uint_as_float(vc[e + 1]));
                        a2 = fmaxf(a2, __uint_as_float(vc[e + 2])); a3 = fmaxf(a3, __uint_as_float(vc[e + 3]));
                    }
                    mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (!((dd >> e) & 1u)) mx = fmaxf(mx, __uint_as_float(vc[e]));
                }
            }
            // the reference only moves when it must (first live tile) or when the tile tops it by more than 8
            const float m_new = (mx > m_ref + 8.f || m_ref == -INFINITY) ? mx : m_ref;
            const float alpha = (m_new == m_ref || m_ref == -INFINITY) ? 1.f : ex2_approx((m_ref - m_new) * kLog2e);
            const float mneg = (m_new == -INFINITY) ? 0.f : -m_new * kLog2e;
            // pass 2: probabilities -> packed bf16 (hi, lo) back into tensor memory (hi in place over the scores)
            float lsum0 = 0.f, lsum1 = 0.f;
            const bool paired = drop_on && ((row_base & 1u) == 0u);   // (k0 and 32 c are even)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                // (chunk 0 was requested at the end of pass 1 into v[0]; chunk c lives in v[c & 1])
                tmem_ld_wait();
                if (c < 3) tmem_ld_32x32(ts + (c + 1) * 32, v[(c + 1) & 1]);
                const uint32_t* vc = v[c & 1];
                const uint32_t dd = dead[c];
                const uint32_t jb = row_base + (uint32_t)(k0 + c * 32);
                uint32_t hw[16], lw[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(vc[2 * e]), kLog2e, mneg));
                    float p1 = ex2_approx(fmaf(__uint_as_float(vc[2 * e + 1]), kLog2e, mneg));
                    if (dd) {
                        if ((dd >> (2 * e)) & 1u) p0 = 0.f;
                        if ((dd >> (2 * e + 1)) & 1u) p1 = 0.f;
                    }
                    lsum0 += p0;
                    lsum1 += p1;
                    if (drop_on) {      // (the 1 / (1 - p) factor is applied once, to O, in the epilogue)
                        bool k0b, k1b;
                        if (paired) {
                            const uint32_t hsh = rng_pair32(key, (jb >> 1) + (uint32_t)e);
                            k0b = hsh >= thr32;
                            k1b = rng_pair32_odd(hsh) >= thr32;
                        } else {
                            k0b = rng_keep16(key, jb + 2u * e, thr32 >> 16);
                            k1b = rng_keep16(key, jb + 2u * e + 1u, thr32 >> 16);
                        }
                        p0 = k0b ? p0 : 0.f;
                        p1 = k1b ? p1 : 0.f;
                    }
                    const uint32_t hh = pack_bf16x2(p0, p1);
                    hw[e] = hh;
                    lw[e] = pack_bf16x2(p0 - __uint_as_float(hh << 16), p1 - __uint_as_float(hh & 0xFFFF0000u));
                }
                tmem_st_32x16(ts + c * 16, hw);
                tmem_st_32x16(tpl + c * 16, lw);
            }
            l_run = l_run * alpha + (lsum0 + lsum1);
            m_ref = m_new;
            // O[x] accumulated so far (this set's earlier tiles) follows the reference
            if (it > 0) {
                mbar_wait(&pv_done[x], (it - 1) & 1);
                tc_fence_after();
                if (!__all_sync(0xffffffffu, alpha == 1.f)) {
                    uint32_t o[32];
                    tmem_ld_32x32(to, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                    tmem_st_32x32(to, o);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[x]);
        }
        // epilogue: merge the two sets' partial (m, l, O) like a split-KV reduction, then O / l -> out, lse
        uint32_t o[32];
        if (it > 0) {
            mbar_wait(&pv_done[x], (it - 1) & 1);
            tc_fence_after();
            tmem_ld_32x32(to, o);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = 0u;
        }
        float* mrow = sMerge + r * kMergeFloats;
        if (x == 1) {
            mrow[0] = m_ref;
            mrow[1] = l_run;
#pragma unroll
            for (int e = 0; e < 8; ++e) *reinterpret_cast<uint4*>(mrow + 4 + 4 * e) = make_uint4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");             // the 8 softmax warps only
        if (x == 0 && qi < p.Lq) {
            const float m1 = mrow[0], l1 = mrow[1];
            const float m = fmaxf(m_ref, m1);
            const float a0 = (m_ref == -INFINITY) ? 0.f : ex2_approx((m_ref - m) * kLog2e);
            const float a1 = (m1 == -INFINITY) ? 0.f : ex2_approx((m1 - m) * kLog2e);
            const float l = l_run * a0 + l1 * a1;
            const float inv = l > 0.f ? (drop_on ? inv_keep : 1.f) / l : 0.f;
            const float s0 = a0 * inv, s1 = a1 * inv;
            float* dst = p.out + ((size_t)b * p.Lq + qi) * p.ldo + h * kHD;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 o1 = *reinterpret_cast<const float4*>(mrow + 4 + 4 * e);
                *reinterpret_cast<float4*>(dst + 4 * e) =
                    make_float4(__uint_as_float(o[4 * e]) * s0 + o1.x * s1, __uint_as_float(o[4 * e + 1]) * s0 + o1.y * s1,
                                __uint_as_float(o[4 * e + 2]) * s0 + o1.z * s1, __uint_as_float(o[4 * e + 3]) * s0 + o1.w * s1);
            }
            p.lse[((size_t)b * p.H + h) * p.Lq + qi] = l > 0.f ? m + __logf(l) : -INFINITY;
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

// Launcher used by mdb_attention_forward_f32 (attention.cu).  Returns MDB_EUNSUPPORTED when the tensor-core kernel does not
// apply (the caller then runs the legacy kernel); any other non-zero value is an error.
int mdb_attention_forward_tc(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask, float* out,
                             float* lse, int B, int H, int Lq, int Lk, int ldq, int ldk, int ldv, int ldo, float drop_p,
                             const unsigned long long* seed, unsigned long long site, cudaStream_t stream) {
    static const bool disabled = getenv("MDB_ATTN_LEGACY") != nullptr;            // A/B switch (profiling)
    if (disabled || Lk < 256) return MDB_EUNSUPPORTED;                            // short sequences: the register kernel
    if ((ldq | ldk | ldv | ldo) % 4) return MDB_EUNSUPPORTED;
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15u) return MDB_EUNSUPPORTED;
    CUtensorMap mq, mk, mv;
    const int E = H * kHD;
    {
        uint64_t dims[3] = {(uint64_t)E, (uint64_t)Lq, (uint64_t)B};
        uint64_t str[3] = {1, (uint64_t)ldq, (uint64_t)Lq * ldq};
        uint32_t box[3] = {kHD, kTile, 1};
        if (make_map(&mq, q, 3, dims, str, box, nullptr)) return MDB_EUNSUPPORTED;
    }
    {
        uint64_t dims[3] = {(uint64_t)E, (uint64_t)Lk, (uint64_t)B};
        uint64_t str[3] = {1, (uint64_t)ldk, (uint64_t)Lk * ldk};
        uint32_t box[3] = {kHD, kTile, 1};
        if (make_map(&mk, k, 3, dims, str, box, nullptr)) return MDB_EUNSUPPORTED;
        str[1] = (uint64_t)ldv; str[2] = (uint64_t)Lk * ldv;
        if (make_map(&mv, v, 3, dims, str, box, nullptr)) return MDB_EUNSUPPORTED;
    }
    TcAttnParams p;
    memset(&p, 0, sizeof(p));
    p.kpm = key_padding_mask; p.out = out; p.lse = lse;
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.ldo = ldo;
    p.scale = 1.f / sqrtf((float)kHD); p.drop_p = drop_p; p.seed = seed; p.site = site;
    constexpr int smem = kTileBytes + kStages * kStageBytes + kTile * kMergeFloats * 4 + 256 + 1024;
    static bool configured[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    dim3 grid((Lq + kTile - 1) / kTile, H, B);
    attn_fwd_tc_kernel<<<grid, kThreads, smem, stream>>>(mq, mk, mv, p);
    return (int)cudaGetLastError();
}
