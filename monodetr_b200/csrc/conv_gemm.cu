// conv_gemm.cu -- tcgen05 / TMEM / TMA implicit-GEMM kernel family for sm_100a (fp32 storage, TF32
// tensor-core math, fp32 accumulate in TMEM).  One kernel template serves
//   * linear layers and 1x1 convolutions (a GEMM is a 1-tap convolution over a 1-row image),
//   * KxK convolutions, im2col-free: for every filter tap the TMA engine fetches the SHIFTED
//     NHWC activation box (out-of-bounds = zero fill = padding; element strides = conv stride),
//   * their data gradients (same kernel, B operand read MN-major straight from the packed weights),
//   * their weight gradients (both operands MN-major straight from dY / X, split-K with vector reds).
// It replaces the cuDNN / cuBLAS calls behind nn.Conv2d / nn.Linear on the reference path
// (lib/models/monodetr/backbone.py:100-102 torchvision ResNet-50 convs; monodetr.py:83-91 input_proj;
// depth_predictor/depth_predictor.py:29-47; ops/modules/ms_deform_attn.py:138-161 projections;
// depthaware_transformer.py:339-343,467-473 FFN / decoder linears).
//
// Layouts: activations NHWC fp32 [B][H][W][C]; packed weights [tap][Cout][Cin]; output NHWC.
// Tile: 128 output pixels (tw x th pixels of tb consecutive images) x BN output channels, K step = 32 fp32
// (one 128-byte swizzle span).  Persistent CTAs (one per SM) walk the tiles.  Warp roles: warp 0 = TMA producer,
// warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2-5 = 3xTF32 splitter (default precision; they also
// move the A tile into tensor memory), last 4 warps = epilogue (TMEM -> registers -> smem transpose -> fused bias /
// residual / ReLU / ReLU-mask / row-scale -> global).  smem ring of STAGES stages, mbarrier full/empty pairs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../include/monodetr_b200.h"
#include "tc_common.cuh"
#include "tma_host.cuh"

// Role-timeline instrumentation (tools/diag_timeline.py): compiled in only with -DMDB_TIMELINE (MDB_TIMELINE=1 python -m
// monodetr_b200.build); the product library carries neither the clock64 stamps nor their predicates.
#ifdef MDB_TIMELINE
#define MDB_STAMP(cond, slot) do { if (cond) p.dbg[slot] = clock64(); } while (0)
#define MDB_DBG(expr) (expr)
#else
#define MDB_STAMP(cond, slot) do { } while (0)
#define MDB_DBG(expr) false
#endif

namespace {

using namespace mdb;

constexpr int BM = 128;
constexpr int BK = 32;                 // fp32 elements per k-block = 128 bytes
constexpr int kTileABytes = BM * 128;  // 16 KiB
constexpr int kChunkBytes = 32 * 128;  // one MN-major chunk: 32 reduction rows x 128 B
constexpr int kThreadsTC = 192;
constexpr int kMaxTaps = 9;
constexpr int kPatchBytes = 32 * 32 * 4;   // one epilogue warp's 32x32 fp32 transpose patch (XOR-swizzled 16-byte groups, no padding)
constexpr int kEpiWarps = 4;               // epilogue warps of the 3xTF32 kernels (8 = two per TMEM lane quarter was measured slower: 128-register cap -> spills)

struct TcParams {
    // ---- fprop / dgrad: decomposition of the M dimension into th x tw pixel rectangles ----------
    int tiles_x, tiles_y, tw, th, tb;   // an M tile = tw x th pixels of tb consecutive images (tw * th * tb == 128)
    int Ho, Wo;                      // logical output grid covered by tiles (bounds for rows)
    int out_sy, out_sx, out_oy, out_ox, out_H, out_W;   // real output pixel = (y*out_sy+out_oy, ...)
    int in_sy, in_sx;                // A-box origin = (y0*in_sy + dy[tap], x0*in_sx + dx[tap])
    int ntaps, cblocks;
    int tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_w[kMaxTaps];
    // ---- wgrad: reduction over pixel tiles of 32 (rth x rtw), split across blockIdx.z ----------
    int rtiles_x, rtiles_y, rtw, rth, n_img, red_per_split;
    // ---- fprop split-K (few tiles, very long reduction): slice z of the k-blocks writes its partial tile to out + z * slice_stride
    int kb_per_slice, slice_stride;  // 0 = off; a fixed-order reduction kernel sums the slices (deterministic, unlike atomics)
    int total_tiles, n_tiles_n, n_tiles_m;   // persistent tile walk: tile = (z * n_tiles_m + m) * n_tiles_n + n
    int w_sy, w_sx, wg_taps, wg_kw, wg_pad;   // X-box origin = (y0*w_sy + ky - pad, x0*w_sx + kx - pad); all taps in ONE launch
    // ---- epilogue -----------------------------------------------------------------------------
    int Mo_rows;                     // wgrad: number of valid output rows (Cout)
    int No, ldo;
    int relu, atomic_out, round_out;   // round_out: store round-to-nearest TF32 (next consumer is a tensor-core operand)
    const float* bias;               // [No] or null
    const float* residual;           // same indexing as out, or null
    const float* relu_mask;          // same indexing as out: out *= (mask > 0), or null
    const float* rowscale;           // wgrad: [Mo_rows] or null
    float* colsum_out;               // wgrad (TMEM-A kernels): [Mo_rows] += column sums of dy (bias gradient), or null
    float* out;
    long long* dbg;                  // profiling aid (tools/diag_timeline.py): clock64 stamps of CTA 0, or null
};

// Persistent, warp-specialised kernel.  Each CTA walks tiles  tile = blockIdx.x + i * gridDim.x  and keeps
//   warp 0          TMA producer (smem ring of STAGES stages, full/empty mbarriers)
//   warp 1          TMEM allocator + single-thread tcgen05.mma issuer; TWO accumulators in TMEM so that
//   last 4 warps    the epilogue of tile i (tcgen05.ld -> fused bias/residual/ReLU/mask/row-scale -> global)
//                   overlaps the main loop of tile i+1 (tmem_full / tmem_empty mbarriers)
//   warps 2-5       (SPLIT only) splitter warps: error-compensated "3xTF32".  For every landed stage they write
//                   lo = x - trunc_tf32(x) of both tiles next to the raw tiles; the MMA warp then issues
//                   A*B + A_lo*B + A*B_lo (the tensor core truncates the raw fp32 bits itself) = ~fp32 accuracy.
//                   With TA (default) the A tile and its lo part go to TENSOR MEMORY instead (see kStageBytes below).
//   BF (bf16x3)     error-compensated BF16: the B operand arrives PRE-SPLIT from global memory -- packed weights
//                   [tap][n][k-block][hi 32 | lo 32] bf16, one 128-byte swizzle row per (n, k-block), written once per step
//                   by pack_gemm_weights_bf16x3 -- and the splitter warps turn each landed fp32 A row into packed bf16
//                   hi / lo pairs in TENSOR MEMORY.  Per k-block the tensor core issues A_hi*B_hi + A_lo*B_hi + A_hi*B_lo
//                   as kind::f16 MMAs (K = 16): the same three products as 3xTF32 at twice the tensor rate, half the
//                   operand bytes and no shared-memory rewrite of B (the 3xTF32 main loop is shared-memory-bandwidth
//                   bound, profiles/r01_conv_gemm_tmemA_k256_ncu.txt).  Dropped terms are O(2^-17) per product.
//   TS (TMA store)  epilogue variant for outputs without residual / mask / atomics: TMEM -> registers (+ bias, ReLU) ->
//                   the warp's XOR-swizzled smem patch, which IS the SWIZZLE_128B image of a 32-row x 32-column box, ->
//                   one cp.async.bulk.tensor store per chunk (double-buffered patches).  The per-row address arithmetic, the
//                   patch read-back and the 8 predicated STG.128 per chunk of the register path -- which made every GEMM
//                   with K <= 512 epilogue-bound (~1750 cycles per 32x32 chunk, gpurun r2 timelines) -- disappear; edge
//                   clipping is done by the TMA unit.
//                   TS == 2 adds the residual and / or ReLU-mask operands (ResNet conv3 + identity, every masked dgrad): their
//                   32x32 boxes are TMA-LOADED into the warp's patches two chunks ahead (3 buffers, mbarrier per buffer), each
//                   lane combines its row in place and the patch goes out through the same bulk store.
//                   TS == 3: both operands (two patches per buffer).  TS kernels run EW = 8 epilogue warps (two per TMEM lane
//                   quarter, alternating 32-column chunks): the per-chunk latencies (TMEM load, proxy fence, bulk-store issue,
//                   operand loads in flight) are what bounds the epilogue, and they parallelise across warps.
template <int BN, int STAGES, int MODE /*0 fprop/dgrad, 1 wgrad*/, bool B_MN, bool SPLIT, bool TA = false, bool BF = false,
          int TS = 0, int EW = kEpiWarps>
__global__ void __launch_bounds__(SPLIT ? 192 + 32 * EW : 192)
tc_conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapR,
                    const __grid_constant__ CUtensorMap mapM, const __grid_constant__ TcParams p) {
    constexpr bool A_MN = (MODE == 1);
    constexpr int kTileBBytes = BN * 128;
    constexpr int kRawBytes = kTileABytes + kTileBBytes;          // what TMA delivers per stage
    // TA (3xTF32, K-major A): the A tile and its lo part live in TENSOR MEMORY, not smem.  The splitter warps copy each
    // landed A row smem -> registers -> TMEM (tcgen05.st) next to the lo part they compute, and the MMAs take A from TMEM
    // ([a_tmem] operand form).  Per k-block that removes 3 tensor-core reads and 1 write of a 16 KB tile from the shared
    // memory pipe -- the kernel is smem-bandwidth bound (TMA fill + splitter + 3 operand sweeps = 192 KB per k-block at
    // 128 B/clk, profiles/r01_conv_gemm_timeline_k256.txt) -- and frees the smem for one more pipeline stage.
    // TMEM columns: [0, 2*BN) two accumulators, then per stage 32 columns raw A + 32 columns lo A.
    static_assert(!TA || SPLIT, "TMEM-resident A: 3xTF32 kernels only");
    static_assert(!BF || (TA && B_MN == (MODE == 1)), "bf16x3: fprop / dgrad with K-major pre-split weights, or wgrad");
    // bf16x3 wgrad: X arrives fp32 MN-major like dY; the splitter transposes it into a K-major [n][hi 32 | lo 32] bf16 tile
    // (the layout the packed weights have in fprop) behind the raw tiles, so the MMA side is identical to fprop's.
    constexpr bool B_DESC_MN = B_MN && !BF;
    constexpr int kAColsPerStage = BF ? 32 : 64;                  // TMEM columns of one stage's A operand (hi + lo)
    static_assert(!TA || 2 * BN + kAColsPerStage * STAGES <= 512, "TMEM budget");
    constexpr int kStageBytes = BF ? kRawBytes + (MODE == 1 ? kTileBBytes : 0) : (TA ? kTileABytes + 2 * kTileBBytes : (SPLIT ? 2 * kRawBytes : kRawBytes));  // + the lo tiles
    constexpr int kBLoOff = TA ? kTileBBytes : kRawBytes;          // B lo relative to B raw
    constexpr int kTmemCols = TA ? 512 : 2 * BN;
    constexpr int kEpiWarp0 = SPLIT ? 6 : 2;                        // first epilogue warp
    constexpr int EPI = SPLIT ? EW : 4;                             // epilogue warps
    static_assert(!(MODE == 1) || B_MN, "wgrad reads both operands MN-major");

    // NOTE: index the extern array directly.  Rounding the pointer up through uintptr_t made the compiler lose the
    // shared address space and emit GENERIC ld/st for every smem access of the splitter and the epilogue (measured:
    // ~2000 cycles per 32-column epilogue chunk).  The kernel has no static smem, so the dynamic window starts at the
    // CTA's (1024-byte aligned) shared base; the assumption is checked once below.
    static_assert(TS == 0 || MODE == 0, "TMA-store epilogue: fprop / dgrad");
    constexpr int kTsBufs = TS == 2 ? 3 : 2;                                   // patch buffers per epilogue warp
    constexpr int kTsBufBytes = TS == 3 ? 2 * kPatchBytes : kPatchBytes;       // TS == 2: [operand -> output]; 3: [residual -> output | mask]
    constexpr int kTsPatchBytes = TS ? kTsBufs * kTsBufBytes * (SPLIT ? EW : 4) : 0;   // 1024-byte aligned region
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes + kTsPatchBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* split_bar = empty_bar + STAGES;                     // splitter warps -> MMA (SPLIT only)
    uint64_t* tmem_full = split_bar + STAGES;                     // [2]
    uint64_t* tmem_empty = tmem_full + 2;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    [[maybe_unused]] uint64_t* ld_bar = tmem_empty + 3;           // TS == 2: [epilogue warp][buffer] residual / mask loads landed

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- tile decode (identical in every role) -----------------------------------------------------
    struct Tile { int n0, img, y0, x0, m0, red_begin, iters, tap, slice; };
    auto decode = [&](int tix) {
        Tile t;
        t.n0 = (tix % p.n_tiles_n) * BN;
        int r = tix / p.n_tiles_n;
        t.img = t.y0 = t.x0 = t.m0 = t.red_begin = t.tap = t.slice = 0;
        if constexpr (MODE == 0) {
            const int per_img = p.tiles_x * p.tiles_y;
            t.iters = p.ntaps * p.cblocks;
            if (p.kb_per_slice > 0) {
                t.slice = r / p.n_tiles_m;
                r -= t.slice * p.n_tiles_m;
                t.red_begin = t.slice * p.kb_per_slice;
                t.iters = max(0, min(t.iters, t.red_begin + p.kb_per_slice) - t.red_begin);
            }
            const int grp = r / per_img;
            t.img = grp * p.tb;
            r -= grp * per_img;
            t.y0 = (r / p.tiles_x) * p.th;
            t.x0 = (r % p.tiles_x) * p.tw;
        } else {
            t.m0 = (r % p.n_tiles_m) * BM;
            r /= p.n_tiles_m;
            t.tap = r % p.wg_taps;
            const int z = r / p.wg_taps;
            const int total_red = p.n_img * p.rtiles_x * p.rtiles_y;
            t.red_begin = z * p.red_per_split;
            t.iters = max(0, min(total_red, t.red_begin + p.red_per_split) - t.red_begin);
        }
        return t;
    };

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();                     // SWIZZLE_128B tiles need 1024-byte aligned stages
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if constexpr (TS != 0) tma_prefetch_desc(&mapO);
        if constexpr (TS >= 2) {
            tma_prefetch_desc(&mapR);
            tma_prefetch_desc(&mapM);
            for (int i = 0; i < 24; ++i) mbar_init(&ld_bar[i], 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&split_bar[s], 4);                          // one arrival per splitter warp
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], EPI);                       // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch (launch_tc): everything above -- barrier init, tensor-map prefetch, TMEM allocation --
    // touches no global data and may run while the previous kernel of the stream is still finishing; every thread waits here
    // (before any role reads or writes global memory, and before any thread can exit, so that completion of this grid implies
    // completion of its predecessors) until that kernel has completed and its writes are visible.  A no-op for a normal launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ... and let the NEXT kernel of the stream (if it was launched as a programmatic dependent) start its own prologue now:
    // it waits at the same point for this grid to complete.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ================================ TMA producer =============================================
        if (elect_one()) {
            int git = 0;                                          // ring position, continues across tiles
            for (int tix = blockIdx.x; tix < p.total_tiles; tix += gridDim.x) {
                const Tile t = decode(tix);
                for (int it = 0; it < t.iters; ++it, ++git) {
                    const int s = git % STAGES;
                    const uint32_t ph = (git / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    MDB_STAMP(p.dbg && blockIdx.x == 0 && git < 48, 384 + git);
                    uint8_t* a_dst = smem + s * kStageBytes;
                    uint8_t* b_dst = a_dst + kTileABytes;
                    mbar_arrive_expect_tx(&full_bar[s], kRawBytes);
                    if constexpr (MODE == 0) {
                        const int kit = t.red_begin + it;                 // (fprop split-K: this slice's first k-block)
                        const int tap = kit / p.cblocks;
                        const int cb = kit - tap * p.cblocks;
                        tma_load_4d(a_dst, &mapA, &full_bar[s], cb * BK, t.x0 * p.in_sx + p.tap_dx[tap],
                                    t.y0 * p.in_sy + p.tap_dy[tap], t.img);
                        if constexpr (BF) {   // bf16 map: one 128-byte row = [hi 32 | lo 32] of k-block cb
                            tma_load_3d(b_dst, &mapB, &full_bar[s], cb * 64, t.n0, p.tap_w[tap]);
                        } else if constexpr (!B_MN) {
                            tma_load_3d(b_dst, &mapB, &full_bar[s], cb * BK, t.n0, p.tap_w[tap]);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / 32; ++j)
                                tma_load_3d(b_dst + j * kChunkBytes, &mapB, &full_bar[s], t.n0 + j * 32, cb * BK, p.tap_w[tap]);
                        }
                    } else {
                        int r = t.red_begin + it;
                        const int per_img = p.rtiles_x * p.rtiles_y;
                        const int ri = r / per_img;
                        r -= ri * per_img;
                        const int ry = (r / p.rtiles_x) * p.rth;
                        const int rx = (r % p.rtiles_x) * p.rtw;
#pragma unroll
                        for (int j = 0; j < BM / 32; ++j)
                            tma_load_4d(a_dst + j * kChunkBytes, &mapA, &full_bar[s], t.m0 + j * 32, rx, ry, ri);
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j)
                            tma_load_4d(b_dst + j * kChunkBytes, &mapB, &full_bar[s], t.n0 + j * 32,
                                        rx * p.w_sx + (t.tap % p.wg_kw) - p.wg_pad, ry * p.w_sy + (t.tap / p.wg_kw) - p.wg_pad, ri);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==============================================
        if (elect_one()) {
            constexpr uint32_t idesc = BF ? make_idesc_bf16(BM, BN) : make_idesc_tf32(BM, BN, A_MN && !TA, B_MN);   // A from TMEM is always lane = m, column = k
            constexpr uint64_t kDescHiA = (A_MN ? make_smem_desc(0, kChunkBytes, 512, 1) : make_smem_desc(0, 16, 1024, 2)) & 0xFFFFFFFF00000000ull;
            constexpr uint64_t kDescHiB = (B_DESC_MN ? make_smem_desc(0, kChunkBytes, 512, 1) : make_smem_desc(0, 16, 1024, 2)) & 0xFFFFFFFF00000000ull;
            constexpr uint32_t kDescLoA = (uint32_t)((A_MN ? make_smem_desc(0, kChunkBytes, 512, 1) : make_smem_desc(0, 16, 1024, 2)) & 0xFFFF0000ull);
            constexpr uint32_t kDescLoB = (uint32_t)((B_DESC_MN ? make_smem_desc(0, kChunkBytes, 512, 1) : make_smem_desc(0, 16, 1024, 2)) & 0xFFFF0000ull);
            const uint32_t smem_base_u32 = smem_u32(smem);
            int git = 0, lt = 0;                                  // lt counts tiles that have a main loop
            for (int tix = blockIdx.x; tix < p.total_tiles; tix += gridDim.x) {
                const Tile t = decode(tix);
                if (t.iters == 0) continue;
                const int slot = lt & 1;
                [[maybe_unused]] const bool dbg = MDB_DBG(p.dbg && blockIdx.x == 0 && lt < 12);
                MDB_STAMP(dbg, 256 + lt * 4 + 0);
                mbar_wait(&tmem_empty[slot], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                MDB_STAMP(dbg, 256 + lt * 4 + 1);
                const uint32_t tacc = tmem_base + slot * BN;
                for (int it = 0; it < t.iters; ++it, ++git) {
                    const int s = git % STAGES;
                    const uint32_t ph = (git / STAGES) & 1;
                    mbar_wait(SPLIT ? &split_bar[s] : &full_bar[s], ph);
                    tc_fence_after();
                    // Descriptors: every field except the 14-bit start address is a per-kernel constant, so a k-step costs
                    // one add + shift + or per operand (the single issuing thread is otherwise instruction-bound).
                    // K-major (SWIZZLE_128B): 8 fp32 = 32 B further along the 128-B row; 8-row groups 1024 B apart (SBO).
                    // MN-major (128B swizzle, 32-B atoms, 4-row period): 8 reduction rows = 1024 B further per k-step;
                    // 4-row groups 512 B apart (SBO); 32-wide MN chunks 4096 B apart (LBO).
                    const uint32_t a_addr = smem_base_u32 + s * kStageBytes;
                    const uint32_t b_addr = a_addr + kTileABytes + ((BF && MODE == 1) ? kTileBBytes : 0);   // wgrad: the split tile
                    if constexpr (BF) {
                        // bf16x3: k-step = 16 bf16 = 8 TMEM columns of A (two per 32-bit column) / 32 bytes of a B row;
                        // A hi at columns [0,16) of the stage, lo at [16,32); B hi at bytes [0,64) of the row, lo at [64,128).
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint32_t ta = tmem_base + 2 * BN + s * kAColsPerStage + ks * 8;
                            const uint64_t bhi = kDescHiB | (uint64_t)(kDescLoB | (((b_addr + ks * 32) >> 4) & 0x3FFFu));
                            const uint64_t blo = kDescHiB | (uint64_t)(kDescLoB | (((b_addr + 64 + ks * 32) >> 4) & 0x3FFFu));
                            umma_bf16_ta(tacc, ta, bhi, idesc, (it > 0) || (ks > 0));
                            umma_bf16_ta(tacc, ta + 16, bhi, idesc, true);
                            umma_bf16_ta(tacc, ta, blo, idesc, true);
                        }
                    } else {
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint32_t ao = A_MN ? k * 1024 : k * 32, bo = B_MN ? k * 1024 : k * 32;
                        const uint64_t adesc = kDescHiA | (uint64_t)(kDescLoA | (((a_addr + ao) >> 4) & 0x3FFFu));
                        const uint64_t bdesc = kDescHiB | (uint64_t)(kDescLoB | (((b_addr + bo) >> 4) & 0x3FFFu));
                        if constexpr (TA) {
                            const uint32_t ta = tmem_base + 2 * BN + s * 64 + k * 8;      // raw A columns of this k-step; lo at +32
                            const uint64_t blo = kDescHiB | (uint64_t)(kDescLoB | (((b_addr + kBLoOff + bo) >> 4) & 0x3FFFu));
                            umma_tf32_ta(tacc, ta, bdesc, idesc, (it > 0) || (k > 0));
                            umma_tf32_ta(tacc, ta + 32, bdesc, idesc, true);
                            umma_tf32_ta(tacc, ta, blo, idesc, true);
                        } else {
                            umma_tf32(tacc, adesc, bdesc, idesc, (it > 0) || (k > 0));
                            if constexpr (SPLIT) {
                                const uint64_t alo = kDescHiA | (uint64_t)(kDescLoA | (((a_addr + kRawBytes + ao) >> 4) & 0x3FFFu));
                                const uint64_t blo = kDescHiB | (uint64_t)(kDescLoB | (((b_addr + kRawBytes + bo) >> 4) & 0x3FFFu));
                                umma_tf32(tacc, alo, bdesc, idesc, true);
                                umma_tf32(tacc, adesc, blo, idesc, true);
                            }
                        }
                    }
                    }
                    umma_commit(&empty_bar[s]);                   // frees the smem stage when these MMAs retire
                }
                umma_commit(&tmem_full[slot]);                    // accumulator complete
                MDB_STAMP(dbg, 256 + lt * 4 + 2);
                ++lt;
            }
        }
    } else if (SPLIT && warp < kEpiWarp0) {
        // ================================ splitter warps (3xTF32) ===================================
        const int tid = threadIdx.x - 64;                         // 0..127
        int git = 0;
        for (int tix = blockIdx.x; tix < p.total_tiles; tix += gridDim.x) {
            const Tile t = decode(tix);
            [[maybe_unused]] float rsum = 0.f;                    // wgrad bias gradient: this thread's dy channel summed over pixels
            for (int it = 0; it < t.iters; ++it, ++git) {
                const int s = git % STAGES;
                const uint32_t ph = (git / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                if constexpr (BF && MODE == 1) {
                    // wgrad: both operands arrive MN-major -- chunk (warp & 3) holds 32 channels as [32 reduction rows (pixels)]
                    // [128 B], 32-byte atoms XOR-swizzled with the row (SWIZZLE_128B_ATOM_32B: atom ^= row & 3).  This thread
                    // owns dy channel m = its TMEM lane AND x channel n = the same index of the B tile: a warp-wide LDS.32 reads
                    // one full 128-byte row (conflict-free), 32 of them give the thread its channel's 32 pixels = one K-major
                    // row, which is split into packed bf16 (hi, lo) pairs: A -> tensor memory, B -> the K-major smem tile.
                    const int row = (warp & 3) * 32 + lane;
                    const int atom = lane >> 3;
                    uint32_t w[32];
                    {
                        const uint8_t* acol = smem + s * kStageBytes + (warp & 3) * kChunkBytes + (lane & 7) * 4;
                        float q0 = 0.f, q1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float x0 = *reinterpret_cast<const float*>(acol + (2 * j) * 128 + ((atom ^ ((2 * j) & 3)) << 5));
                            const float x1 = *reinterpret_cast<const float*>(acol + (2 * j + 1) * 128 + ((atom ^ ((2 * j + 1) & 3)) << 5));
                            q0 += x0; q1 += x1;              // bias gradient for free (TMA zero-fills pixels / channels past the edge)
                            const uint32_t h = pack_bf16x2(x0, x1);
                            w[j] = h;
                            w[16 + j] = pack_bf16x2(x0 - __uint_as_float(h << 16), x1 - __uint_as_float(h & 0xFFFF0000u));
                        }
                        rsum += q0 + q1;
                    }
                    tmem_st_32x32(tmem_base + 2 * BN + s * kAColsPerStage + ((uint32_t)((warp & 3) * 32) << 16), w);
                    if (row < BN) {
                        const uint8_t* bcol = smem + s * kStageBytes + kTileABytes + (warp & 3) * kChunkBytes + (lane & 7) * 4;
                        uint32_t hw[16], lw[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float x0 = *reinterpret_cast<const float*>(bcol + (2 * j) * 128 + ((atom ^ ((2 * j) & 3)) << 5));
                            const float x1 = *reinterpret_cast<const float*>(bcol + (2 * j + 1) * 128 + ((atom ^ ((2 * j + 1) & 3)) << 5));
                            const uint32_t h = pack_bf16x2(x0, x1);
                            hw[j] = h;
                            lw[j] = pack_bf16x2(x0 - __uint_as_float(h << 16), x1 - __uint_as_float(h & 0xFFFF0000u));
                        }
                        // row n of the K-major tile: 128 bytes = [hi 64 B | lo 64 B], 16-byte chunk c at c ^ (n & 7) (SWIZZLE_128B)
                        uint8_t* brow = smem + s * kStageBytes + kRawBytes + row * 128;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            *reinterpret_cast<uint4*>(brow + ((c ^ (row & 7)) << 4)) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
                            *reinterpret_cast<uint4*>(brow + (((4 + c) ^ (row & 7)) << 4)) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
                        }
                    }
                    tmem_st_wait();
                    tc_fence_before();             // order the TMEM stores before the MMA thread's reads (pairs with its fence::after)
                } else if constexpr (BF) {
                    // This thread owns tile row (warp % 4) * 32 + lane = its TMEM lane: 32 fp32 -> 16 packed bf16x2 hi words
                    // (element 2j in the low half) + 16 lo words, lo = bf16(x - float(hi)) (the subtraction is exact).
                    const int row = (warp & 3) * 32 + lane;
                    const uint8_t* arow = smem + s * kStageBytes + row * 128;
                    uint32_t w[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 r = *reinterpret_cast<const uint4*>(arow + ((c ^ (row & 7)) << 4));
                        const float x0 = __uint_as_float(r.x), x1 = __uint_as_float(r.y), x2 = __uint_as_float(r.z), x3 = __uint_as_float(r.w);
                        const uint32_t h01 = pack_bf16x2(x0, x1), h23 = pack_bf16x2(x2, x3);
                        w[2 * c] = h01;
                        w[2 * c + 1] = h23;
                        w[16 + 2 * c] = pack_bf16x2(x0 - __uint_as_float(h01 << 16), x1 - __uint_as_float(h01 & 0xFFFF0000u));
                        w[16 + 2 * c + 1] = pack_bf16x2(x2 - __uint_as_float(h23 << 16), x3 - __uint_as_float(h23 & 0xFFFF0000u));
                    }
                    tmem_st_32x32(tmem_base + 2 * BN + s * kAColsPerStage + ((uint32_t)((warp & 3) * 32) << 16), w);
                    tmem_st_wait();
                    tc_fence_before();             // order the TMEM stores before the MMA thread's reads (pairs with its fence::after)
                } else if constexpr (TA) {
                    // A: this thread owns tile row (warp % 4) * 32 + lane = its TMEM lane.  K-major SWIZZLE_128B smem: row r
                    // is 128 bytes, 16-byte chunk c sits at c ^ (r & 7) (conflict-free: 8 lanes hit 8 different chunks).
                    const int row = (warp & 3) * 32 + lane;
                    uint32_t hi[32], lo32[32];
                    if constexpr (!A_MN) {
                        const uint8_t* arow = smem + s * kStageBytes + row * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 r = *reinterpret_cast<const uint4*>(arow + ((c ^ (row & 7)) << 4));
                            hi[4 * c] = r.x; hi[4 * c + 1] = r.y; hi[4 * c + 2] = r.z; hi[4 * c + 3] = r.w;
                        }
                    } else {
                        // wgrad: A arrives MN-major -- chunk (warp & 3) holds this warp's 32 m-columns as [32 reduction rows]
                        // [128 B], 32-byte atoms XOR-swizzled with the row (SWIZZLE_128B_ATOM_32B: atom ^= row & 3).  A warp-wide
                        // LDS.32 reads one full 128-byte row (conflict-free); 32 of them transpose the tile into TMEM order.
                        const uint8_t* acol = smem + s * kStageBytes + (warp & 3) * kChunkBytes + (lane & 7) * 4;
                        const int atom = lane >> 3;
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            hi[k] = *reinterpret_cast<const uint32_t*>(acol + k * 128 + ((atom ^ (k & 3)) << 5));
                        // Bias gradient for free: thread = one dy channel, the 32 values = 32 pixels of the reduction tile
                        // (TMA zero-fills pixels / channels past the edge).  Four partial sums keep the add chain short.
                        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                        for (int k = 0; k < 32; k += 4) {
                            q0 += __uint_as_float(hi[k]); q1 += __uint_as_float(hi[k + 1]);
                            q2 += __uint_as_float(hi[k + 2]); q3 += __uint_as_float(hi[k + 3]);
                        }
                        rsum += (q0 + q1) + (q2 + q3);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        lo32[j] = __float_as_uint(__uint_as_float(hi[j]) - __uint_as_float(hi[j] & 0xFFFFE000u));
                    const uint32_t ta = tmem_base + 2 * BN + s * 64 + ((uint32_t)((warp & 3) * 32) << 16);
                    tmem_st_32x32(ta, hi);
                    tmem_st_32x32(ta + 32, lo32);
                    // B: lo part elementwise, as in the smem-only variant
                    const uint4* raw = reinterpret_cast<const uint4*>(smem + s * kStageBytes + kTileABytes);
                    float4* lo = reinterpret_cast<float4*>(smem + s * kStageBytes + kTileABytes + kTileBBytes);
#pragma unroll 4
                    for (int i = tid; i < kTileBBytes / 16; i += 128) {
                        const uint4 r = raw[i];
                        float4 l;
                        l.x = __uint_as_float(r.x) - __uint_as_float(r.x & 0xFFFFE000u);
                        l.y = __uint_as_float(r.y) - __uint_as_float(r.y & 0xFFFFE000u);
                        l.z = __uint_as_float(r.z) - __uint_as_float(r.z & 0xFFFFE000u);
                        l.w = __uint_as_float(r.w) - __uint_as_float(r.w & 0xFFFFE000u);
                        lo[i] = l;
                    }
                    tmem_st_wait();
                    tc_fence_before();             // order the TMEM stores before the MMA thread's reads (pairs with its fence::after)
                } else {
                    const uint4* raw = reinterpret_cast<const uint4*>(smem + s * kStageBytes);
                    float4* lo = reinterpret_cast<float4*>(smem + s * kStageBytes + kRawBytes);
#pragma unroll 4
                    for (int i = tid; i < kRawBytes / 16; i += 128) {
                        const uint4 r = raw[i];
                        float4 l;
                        l.x = __uint_as_float(r.x) - __uint_as_float(r.x & 0xFFFFE000u);
                        l.y = __uint_as_float(r.y) - __uint_as_float(r.y & 0xFFFFE000u);
                        l.z = __uint_as_float(r.z) - __uint_as_float(r.z & 0xFFFFE000u);
                        l.w = __uint_as_float(r.w) - __uint_as_float(r.w & 0xFFFFE000u);
                        lo[i] = l;
                    }
                }
                fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&split_bar[s]);
            }
            if constexpr (TA && A_MN) {
                // every (column tile, tap) re-reads the same dy tiles: only column tile 0 / tap 0 contributes
                const int m = t.m0 + (warp & 3) * 32 + lane;
                if (p.colsum_out && t.n0 == 0 && t.tap == 0 && t.iters > 0 && m < p.Mo_rows) atomicAdd(p.colsum_out + m, rsum);
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ================================ epilogue warps ===========================================
        // TMEM gives thread `lane` the accumulator ROW q*32+lane (32 consecutive columns per tcgen05.ld).  Writing
        // that straight out would touch 32 different cache lines per store instruction, so every 32x32 block is
        // transposed through a private XOR-swizzled smem patch: afterwards 8 lanes x float4 cover one row's 128 bytes and
        // a warp instruction moves 4 full lines -- residual / mask reads use the same coalesced pattern.
        const int q = warp & 3;                // TMEM lane quarter this warp may access
        if constexpr (TS >= 2) {
            // ---- TMA-store epilogue with TMA-loaded residual / mask ---------------------------------------------------
            // TS == 2: exactly one operand, loaded INTO the output patch; TS == 3: residual there, mask in a second patch.
            constexpr int kMaskOff = TS == 3 ? kPatchBytes : 0;
            uint8_t* patches = smem + STAGES * kStageBytes + (warp - kEpiWarp0) * kTsBufs * kTsBufBytes;
            uint64_t* bars = ld_bar + (warp - kEpiWarp0) * kTsBufs;
            const int chunk0 = ((warp - kEpiWarp0) >> 2) * 32;
            constexpr int kChunkStep = 32 * (EPI / 4);
            const int r0 = q * 32;
            const int lx0 = r0 % p.tw, lyt0 = r0 / p.tw, ib0 = lyt0 / p.th, ly0 = lyt0 - ib0 * p.th;
            const uint32_t ld_bytes = (p.residual ? kPatchBytes : 0) + (p.relu_mask ? kPatchBytes : 0);
            // prefetch iterator: (tile, chunk) of the next operand load, kTsBufs - 1 chunks ahead of the chunk being written
            int pf_tix = blockIdx.x, pf_c0 = chunk0, n_issued = 0, n = 0;
            auto pf_norm = [&]() {
                while (pf_tix < p.total_tiles) {
                    if (pf_c0 < BN && (pf_tix % p.n_tiles_n) * BN + pf_c0 < p.No) return true;
                    pf_tix += gridDim.x;
                    pf_c0 = chunk0;
                }
                return false;
            };
            auto pf_issue = [&]() {            // lane 0: residual / mask boxes of chunk #n_issued into buffer n_issued % kTsBufs
                const Tile t = decode(pf_tix);
                uint8_t* buf = patches + (n_issued % kTsBufs) * kTsBufBytes;
                uint64_t* bar = &bars[n_issued % kTsBufs];
                mbar_arrive_expect_tx(bar, ld_bytes);
                if (p.residual) tma_load_4d(buf, &mapR, bar, t.n0 + pf_c0, t.x0 + lx0, t.y0 + ly0, t.img + ib0);
                if (p.relu_mask) tma_load_4d(buf + kMaskOff, &mapM, bar, t.n0 + pf_c0, t.x0 + lx0, t.y0 + ly0, t.img + ib0);
            };
            bool pf_ok = pf_norm();
            for (int i = 0; i < kTsBufs - 1 && pf_ok; ++i) {
                if (lane == 0) pf_issue();
                ++n_issued;
                pf_c0 += kChunkStep;
                pf_ok = pf_norm();
            }
            int lt = 0;
            for (int tix = blockIdx.x; tix < p.total_tiles; tix += gridDim.x) {
                const Tile t = decode(tix);
                const int slot = lt & 1;
                [[maybe_unused]] const bool dbg = MDB_DBG(p.dbg && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0 && lt < 12);
                MDB_STAMP(dbg, lt * 16 + 0);
                if (t.iters > 0) {
                    mbar_wait(&tmem_full[slot], (lt >> 1) & 1);
                    tc_fence_after();
                }
                MDB_STAMP(dbg, lt * 16 + 1);
                const uint32_t taddr_row = tmem_base + slot * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                for (int c0 = chunk0; c0 < BN; c0 += kChunkStep) {
                    if (t.n0 + c0 >= p.No) break;  // uniform across the warp
                    // operands of chunk n + kTsBufs - 1 go into the buffer chunk n - 1 used: its bulk store must be done reading
                    if (pf_ok) {
                        if (lane == 0) {
                            tma_store_wait_read<0>();
                            pf_issue();
                        }
                        ++n_issued;
                        pf_c0 += kChunkStep;
                        pf_ok = pf_norm();
                    }
                    uint8_t* buf = patches + (n % kTsBufs) * kTsBufBytes;
                    uint32_t r[32];
                    if (t.iters > 0) {
                        tmem_ld_32x32(taddr_row + c0, r);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = 0u;
                    }
                    float4 b4[8];
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            b4[j] = (t.n0 + c0 + 4 * j < p.No) ? __ldg(reinterpret_cast<const float4*>(p.bias + t.n0 + c0) + j)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    mbar_wait(&bars[n % kTsBufs], (n / kTsBufs) & 1);     // residual / mask boxes of this chunk have landed
                    if (t.iters > 0) tmem_ld_wait();
                    MDB_STAMP(dbg && c0 < 128, lt * 16 + 2 + (c0 >> 5) * 3);
                    uint8_t* prow = buf + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4* slot4 = reinterpret_cast<float4*>(prow + ((j ^ (lane & 7)) << 4));
                        float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                               __uint_as_float(r[4 * j + 3]));
                        if (p.bias) { v.x += b4[j].x; v.y += b4[j].y; v.z += b4[j].z; v.w += b4[j].w; }
                        if (p.residual) { const float4 e = *slot4; v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w; }
                        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        if (p.relu_mask) {
                            const float4 m4 = *reinterpret_cast<const float4*>(prow + kMaskOff + ((j ^ (lane & 7)) << 4));
                            v.x = m4.x > 0.f ? v.x : 0.f; v.y = m4.y > 0.f ? v.y : 0.f;
                            v.z = m4.z > 0.f ? v.z : 0.f; v.w = m4.w > 0.f ? v.w : 0.f;
                        }
                        *slot4 = v;
                    }
                    fence_proxy_async_smem();      // generic-proxy accesses before the TMA unit's reads (store) and later writes (loads)
                    __syncwarp();
                    MDB_STAMP(dbg && c0 < 128, lt * 16 + 3 + (c0 >> 5) * 3);
                    if (lane == 0) {
                        tma_store_4d(&mapO, buf, t.n0 + c0, t.x0 + lx0, t.y0 + ly0, t.img + ib0);
                        tma_store_commit();
                    }
                    MDB_STAMP(dbg && c0 < 128, lt * 16 + 4 + (c0 >> 5) * 3);
                    ++n;
                }
                if (t.iters > 0) {
                    tc_fence_before();             // TMEM reads of this warp are done: hand the accumulator back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[slot]);
                    ++lt;
                }
            }
            if (lane == 0) tma_store_wait_all();
        } else if constexpr (TS == 1) {
            // ---- TMA-store epilogue ----------------------------------------------------------------------------------
            uint8_t* patches = smem + STAGES * kStageBytes + (warp - kEpiWarp0) * 2 * kPatchBytes;
            const int chunk0 = ((warp - kEpiWarp0) >> 2) * 32;
            // first row of this warp's 32-row slice inside the tile (tw, th, tb are powers of two with product 128)
            const int r0 = q * 32;
            const int lx0 = r0 % p.tw, lyt0 = r0 / p.tw, ib0 = lyt0 / p.th, ly0 = lyt0 - ib0 * p.th;
            int lt = 0, nbuf = 0;
            for (int tix = blockIdx.x; tix < p.total_tiles; tix += gridDim.x) {
                const Tile t = decode(tix);
                const int slot = lt & 1;
                [[maybe_unused]] const bool dbg = MDB_DBG(p.dbg && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0 && lt < 12);
                MDB_STAMP(dbg, lt * 16 + 0);
                if (t.iters > 0) {
                    mbar_wait(&tmem_full[slot], (lt >> 1) & 1);
                    tc_fence_after();
                }
                MDB_STAMP(dbg, lt * 16 + 1);
                const uint32_t taddr_row = tmem_base + slot * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                for (int c0 = chunk0; c0 < BN; c0 += 32 * (EPI / 4)) {
                    if (t.n0 + c0 >= p.No) break;  // uniform across the warp
                    uint8_t* patch = patches + (nbuf & 1) * kPatchBytes;
                    ++nbuf;
                    // the bulk store issued from this buffer two chunks ago must have finished reading it
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                    uint32_t r[32];
                    if (t.iters > 0) {
                        tmem_ld_32x32(taddr_row + c0, r);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = 0u;
                    }
                    float4 b4[8];
                    if (p.bias) {                  // per-column bias: the same 128 bytes for every lane (broadcast loads)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            b4[j] = (t.n0 + c0 + 4 * j < p.No) ? __ldg(reinterpret_cast<const float4*>(p.bias + t.n0 + c0) + j)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (t.iters > 0) tmem_ld_wait();
                    MDB_STAMP(dbg && c0 < 128, lt * 16 + 2 + (c0 >> 5) * 3);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                               __uint_as_float(r[4 * j + 3]));
                        if (p.bias) { v.x += b4[j].x; v.y += b4[j].y; v.z += b4[j].z; v.w += b4[j].w; }
                        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        *reinterpret_cast<float4*>(patch + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
                    }
                    fence_proxy_async_smem();      // generic-proxy writes -> visible to the TMA unit's async-proxy reads
                    __syncwarp();
                    MDB_STAMP(dbg && c0 < 128, lt * 16 + 3 + (c0 >> 5) * 3);
                    if (lane == 0) {
                        tma_store_4d(&mapO, patch, t.n0 + c0, t.x0 + lx0, t.y0 + ly0, t.img + ib0);
                        tma_store_commit();
                    }
                    MDB_STAMP(dbg && c0 < 128, lt * 16 + 4 + (c0 >> 5) * 3);
                }
                if (t.iters > 0) {
                    tc_fence_before();             // TMEM reads of this warp are done: hand the accumulator back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[slot]);
                    ++lt;
                }
            }
            if (lane == 0) tma_store_wait_all();   // every bulk store of this warp has been written out before the CTA exits
        } else {
        float* patch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 512) + (warp - kEpiWarp0) * (32 * 32);
        const int chunk0 = ((warp - kEpiWarp0) >> 2) * 32;       // EPI == 8: warps 4..7 take the odd 32-column chunks
        const int r4 = lane >> 3, c4 = lane & 7;
        int lt = 0;
        for (int tix = blockIdx.x; tix < p.total_tiles; tix += gridDim.x) {
            const Tile t = decode(tix);
            uint32_t roff[8];                  // element offset (fits 32 bit: host check) of (row, column 0), 8 rows per lane
            float rsc[8];
            uint32_t rok = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = q * 32 + i * 4 + r4;
                rsc[i] = 1.f;
                if constexpr (MODE == 0) {
                    const int lyt = row / p.tw, lx = row - lyt * p.tw;
                    const int ib = lyt / p.th, ly = lyt - ib * p.th;       // image within the tile, row within the image
                    const int y = t.y0 + ly, x = t.x0 + lx, img = t.img + ib;
                    if ((y < p.Ho) && (x < p.Wo) && (img < p.n_img)) rok |= 1u << i;
                    const int oy = y * p.out_sy + p.out_oy, ox = x * p.out_sx + p.out_ox;
                    roff[i] = (uint32_t)(((img * p.out_H + oy) * p.out_W + ox) * p.ldo + t.slice * p.slice_stride);
                } else {
                    if ((t.m0 + row) < p.Mo_rows) {
                        rok |= 1u << i;
                        if (p.rowscale) rsc[i] = p.rowscale[t.m0 + row];
                    }
                    roff[i] = (uint32_t)((t.tap * p.Mo_rows + t.m0 + row) * p.ldo);
                }
            }
            const int slot = lt & 1;
            [[maybe_unused]] const bool dbg = MDB_DBG(p.dbg && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0 && lt < 12);
            MDB_STAMP(dbg, lt * 16 + 0);
            if (t.iters > 0) {
                mbar_wait(&tmem_full[slot], (lt >> 1) & 1);
                tc_fence_after();
            }
            MDB_STAMP(dbg, lt * 16 + 1);
            const uint32_t taddr_row = tmem_base + slot * BN + ((uint32_t)(q * 32) << 16);
            // patch element (row, 16-byte group g) lives at row * 32 + ((g ^ (row & 7)) << 2): conflict-free for the row-per-lane
            // writes (8 lanes = 8 groups) and for the 8-lanes-per-row reads.  This lane reads rows i * 4 + r4, group c4.
            const float* prow0 = patch + r4 * 32 + ((c4 ^ r4) << 2);            // even i
            const float* prow1 = patch + r4 * 32 + ((c4 ^ r4 ^ 4) << 2);        // odd i
#pragma unroll 1
            for (int c0 = chunk0; c0 < BN; c0 += 32 * (EPI / 4)) {
                if (t.n0 + c0 >= p.No) break;  // uniform across the warp
                const int n = t.n0 + c0 + c4 * 4;
                const bool full = (t.n0 + c0 + 32 <= p.No) && !(p.ldo & 3);   // uniform: every lane's float4 is in range and 16B aligned
                // Issue the global reads of this chunk (residual / ReLU mask / bias, read-only path) BEFORE waiting on
                // TMEM so their latency overlaps; the two flags are uniform, the bodies are specialised below.
                float4 res[8], msk[8];
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (full) {
                    if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                    if (p.residual) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if ((rok >> i) & 1u) res[i] = __ldg(reinterpret_cast<const float4*>(p.residual + roff[i] + n));
                    }
                    if (p.relu_mask) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if ((rok >> i) & 1u) msk[i] = __ldg(reinterpret_cast<const float4*>(p.relu_mask + roff[i] + n));
                    }
                }
                uint32_t r[32];
                if (t.iters > 0) {
                    tmem_ld_32x32(taddr_row + c0, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                MDB_STAMP(dbg && c0 < 128, lt * 16 + 2 + (c0 >> 5) * 3);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(patch + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                MDB_STAMP(dbg && c0 < 128, lt * 16 + 3 + (c0 >> 5) * 3);
                if (full) {
                    // Straight-line, flag-specialised row loop.  With runtime flags inside it the compiler emitted ~66
                    // dependent instructions per row and, with one epilogue warp per scheduler, each 16 KB chunk took
                    // ~1900 cycles (profiles/r01_conv_gemm_timeline_cta0.txt); here all 8 patch reads are issued first.
                    auto rows = [&](auto RES, auto RELU, auto MASK, auto ROUND, auto ATOMIC) {
                        float4 a[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(((i & 1) ? prow1 : prow0) + i * 4 * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 v = make_float4(fmaf(a[i].x, rsc[i], bias4.x), fmaf(a[i].y, rsc[i], bias4.y),
                                                   fmaf(a[i].z, rsc[i], bias4.z), fmaf(a[i].w, rsc[i], bias4.w));
                            if constexpr (decltype(RES)::value) { v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w; }
                            if constexpr (decltype(RELU)::value) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            if constexpr (decltype(MASK)::value) {
                                v.x = msk[i].x > 0.f ? v.x : 0.f; v.y = msk[i].y > 0.f ? v.y : 0.f;
                                v.z = msk[i].z > 0.f ? v.z : 0.f; v.w = msk[i].w > 0.f ? v.w : 0.f;
                            }
                            if constexpr (decltype(ROUND)::value) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
                            a[i] = v;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if ((rok >> i) & 1u) {
                                float* dst = p.out + roff[i] + n;
                                if constexpr (decltype(ATOMIC)::value) red_add_v4_f32(dst, a[i].x, a[i].y, a[i].z, a[i].w);
                                else *reinterpret_cast<float4*>(dst) = a[i];
                            }
                        }
                    };
                    using T = std::true_type;
                    using F = std::false_type;
                    if (p.atomic_out) rows(F{}, F{}, F{}, F{}, T{});
                    else if (p.round_out) {       // single-pass TF32 mode only
                        if (p.residual) { if (p.relu_mask) rows(T{}, F{}, T{}, T{}, F{}); else if (p.relu) rows(T{}, T{}, F{}, T{}, F{}); else rows(T{}, F{}, F{}, T{}, F{}); }
                        else { if (p.relu_mask) rows(F{}, F{}, T{}, T{}, F{}); else if (p.relu) rows(F{}, T{}, F{}, T{}, F{}); else rows(F{}, F{}, F{}, T{}, F{}); }
                    } else if (p.residual) {
                        if (p.relu_mask) rows(T{}, F{}, T{}, F{}, F{});
                        else if (p.relu) rows(T{}, T{}, F{}, F{}, F{});
                        else rows(T{}, F{}, F{}, F{}, F{});
                    } else {
                        if (p.relu_mask) rows(F{}, F{}, T{}, F{}, F{});
                        else if (p.relu) rows(F{}, T{}, F{}, F{}, F{});
                        else rows(F{}, F{}, F{}, F{}, F{});
                    }
                } else {
                    // ragged right edge (No not a multiple of 32): scalar path, rare (head outputs, padded N)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (!((rok >> i) & 1u)) continue;
#pragma unroll 1
                        for (int e = 0; e < 4 && n + e < p.No; ++e) {
                            const size_t o = (size_t)roff[i] + n + e;
                            float x = ((i & 1) ? prow1 : prow0)[i * 4 * 32 + e] * rsc[i];
                            if (p.bias) x += p.bias[n + e];
                            if (p.residual) x += p.residual[o];
                            if (p.relu) x = fmaxf(x, 0.f);
                            if (p.relu_mask) x = p.relu_mask[o] > 0.f ? x : 0.f;
                            if (p.round_out) x = round_tf32(x);
                            if (p.atomic_out) atomicAdd(p.out + o, x);
                            else p.out[o] = x;
                        }
                    }
                }
                __syncwarp();                  // the patch is rewritten by the next chunk
                MDB_STAMP(dbg && c0 < 128, lt * 16 + 4 + (c0 >> 5) * 3);
            }
            if (t.iters > 0) {
                tc_fence_before();             // TMEM reads of this warp are done: hand the accumulator back
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[slot]);
                ++lt;
            }
        }
        }   // !TS
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
long long* g_dbg = nullptr;   // see mdb_debug_set_timeline
// Arithmetic mode: a process-wide numerical setting (like torch.backends.cuda.matmul.allow_tf32), not per-device state.
// 0 = single-pass TF32 (operands rounded to nearest), 1 = error-compensated 3xTF32, 2 = error-compensated BF16x3 for
// fprop / dgrad (pre-split weights) and wgrad -- the default.
int g_precision = 2;

// Everything that belongs to ONE device lives here, keyed by cudaGetDevice() (several devices per process: nn.DataParallel,
// tools/train_val.py:50-55): SM count, the split-K scratch registered by the caller, and (in launch_tc) the per-kernel
// max-dynamic-smem attribute, which is a per-device property of a function.
constexpr int kMaxDevices = 64;
struct DeviceState {
    int sms = 0;
    float* ws = nullptr;         // split-K scratch: owned by the CALLER (mdb_set_workspace), never freed / reallocated here
    size_t ws_bytes = 0;
};
DeviceState g_dev[kMaxDevices];

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    return dev;
}

int num_sms_tc() {
    DeviceState& d = g_dev[current_device()];
    if (d.sms == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
        d.sms = sms;
    }
    return d.sms;
}

// `grid` carries the logical tile counts (x = column tiles, y = row tiles, z = split-K slices); the kernel is
// launched persistent with min(total_tiles, SMs * resident CTAs) CTAs.
template <int BN, int STAGES, int MODE, bool B_MN, bool SPLIT, bool TA = false, bool BF = false, int TS = 0, int EW = kEpiWarps>
int launch_tc(const CUtensorMap& a, const CUtensorMap& b, TcParams p, dim3 grid, cudaStream_t stream, const CUtensorMap* o = nullptr,
              const CUtensorMap* r = nullptr, const CUtensorMap* m = nullptr) {
    constexpr int stage = BF ? kTileABytes + (MODE == 1 ? 2 : 1) * BN * 128 : (TA ? kTileABytes + 2 * BN * 128 : (SPLIT ? 2 : 1) * (kTileABytes + BN * 128));
    constexpr int smem = STAGES * stage + 1024 /*align slack*/ + 512 /*barriers*/ +
                         (TS == 3 ? 4 : (TS == 2 ? 3 : (TS == 1 ? 2 : 1))) * (SPLIT ? EW : 4) * kPatchBytes;
    static_assert(smem <= 227 * 1024, "dynamic shared memory budget");
    constexpr int threads = SPLIT ? 192 + 32 * EW : 192;
    static bool configured[kMaxDevices] = {};                     // the attribute is per (function, device)
    auto kern = tc_conv_gemm_kernel<BN, STAGES, MODE, B_MN, SPLIT, TA, BF, TS, EW>;
    if (TS != 0 && !o) return MDB_EINVAL;
    const int dev = current_device();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = true;
    }
    p.n_tiles_n = (int)grid.x;
    p.n_tiles_m = (int)grid.y;
    p.total_tiles = (int)(grid.x * grid.y * grid.z);
    const int resident = (!TA && smem <= 112 * 1024 && 2 * BN * 2 <= 512) ? 2 : 1;   // smem and TMEM (2*BN columns per CTA, 512 with TA)
    int ctas = num_sms_tc() * resident;
    if (ctas > p.total_tiles) ctas = p.total_tiles;
    if (ctas < 1) return 0;
    // Launch as a PROGRAMMATIC DEPENDENT of the previous kernel in the stream (CUDA >= 11.8, graph capture >= 12.3; MDB_NO_PDL=1: plain launch): the
    // grid may be scheduled while that kernel drains, runs its prologue and then blocks in griddepcontrol.wait until the
    // predecessor has completed (see the kernel) -- the launch latency and the prologue of the ~470 GEMM launches of a training
    // step leave the critical path.  Other stream dependencies (events, memsets, copies) stay full dependencies.
    static const bool pdl = getenv("MDB_NO_PDL") == nullptr;
    if (pdl) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)ctas, 1, 1);
        cfg.blockDim = dim3((unsigned)threads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, kern, a, b, o ? *o : a, r ? *r : a, m ? *m : a, p);
    }
    kern<<<ctas, threads, smem, stream>>>(a, b, o ? *o : a, r ? *r : a, m ? *m : a, p);
    return (int)cudaGetLastError();
}

// fprop split-K, second pass: y[i] = bias[i % N] + sum_z ws[z][i] in a fixed order.
__global__ void splitk_reduce_kernel(const float4* __restrict__ ws, const float4* __restrict__ bias, float4* __restrict__ y,
                                     long long n4, int N4, int slices, long long stride4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = bias ? bias[i % N4] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < slices; ++z) {
            const float4 v = ws[z * stride4 + i];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        y[i] = a;
    }
}

// Scratch for the split-K partial tiles.  The library never allocates: the caller registers a buffer for the current
// device (mdb_set_workspace; the Python side takes it from torch's caching allocator, which is CUDA-graph and
// multi-stream aware) after asking mdb_conv2d_forward_workspace_bytes.  Too small / missing -> MDB_EWORKSPACE.
int splitk_workspace(size_t bytes, float** out) {
    DeviceState& d = g_dev[current_device()];
    if (bytes > d.ws_bytes || !d.ws) return MDB_EWORKSPACE;
    *out = d.ws;
    return 0;
}

// choose a th x tw rectangle with th*tw == n_pix (128 for M tiles, 32 for wgrad reduction tiles)
void pick_tile(int W, int H, int n_pix, int* tw, int* th) {
    int best_tw = n_pix, best_waste = 1 << 30;
    for (int w = n_pix; w >= 1; w >>= 1) {
        const int h = n_pix / w;
        const int tx = (W + w - 1) / w, ty = (H + h - 1) / h;
        const int waste = tx * w * ty * h - W * H;
        if (waste < best_waste) { best_waste = waste; best_tw = w; }
    }
    *tw = best_tw;
    *th = n_pix / best_tw;
}

// M tile of fprop / dgrad: tw x th pixels of tb consecutive images, tw * th * tb == n_pix (128).  Spanning images keeps
// small feature maps dense (12 x 40 with 8 images: 8x4 pixels x 4 images = 30 full tiles instead of 40 tiles at 75 %).
void pick_tile3(int W, int H, int B, int n_pix, int* tw, int* th, int* tb) {
    long long best = -1;
    *tw = n_pix; *th = 1; *tb = 1;
    for (int b = 1; b <= n_pix; b <<= 1)
        for (int w = n_pix / b; w >= 1; w >>= 1) {
            const int h = n_pix / b / w;
            if (b > 1 && b >= 2 * B) continue;                    // never more padding images than real ones
            const long long cov = (long long)((W + w - 1) / w) * w * ((H + h - 1) / h) * h * ((B + b - 1) / b) * b;
            if (best < 0 || cov < best) { best = cov; *tw = w; *th = h; *tb = b; }   // ties: fewer images, wider rows
        }
}

// Tensor maps of the TMA-store epilogues (TS kernels): out -- and the residual / ReLU-mask operands, which are indexed like
// out -- as (No, W, H, images) with one epilogue warp's 32-row slice of the tw x th x tb tile as the box.  Returns the
// epilogue variant: 0 = register epilogue (strided / ragged / atomic / split-K outputs), 1 = TMA store, 2 = TMA store with
// one TMA-loaded operand (residual or mask), 3 = with both.
int make_epilogue_maps(CUtensorMap* mo, CUtensorMap* mr, CUtensorMap* mm, const TcParams& p) {
    static const bool enabled = getenv("MDB_NO_TMA_STORE") == nullptr;            // A/B switches (profiling)
    static const bool enabled2 = getenv("MDB_NO_TMA_EPILOGUE_LOADS") == nullptr;
    const bool operands = p.residual || p.relu_mask;
    if (!enabled || (operands && !enabled2) || p.atomic_out || p.kb_per_slice > 0 || p.round_out) return 0;
    if (p.out_sx != 1 || p.out_sy != 1 || p.out_ox != 0 || p.out_oy != 0 || (p.ldo & 3) || p.No != p.ldo) return 0;
    if ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.bias) | reinterpret_cast<uintptr_t>(p.residual) |
         reinterpret_cast<uintptr_t>(p.relu_mask)) & 15u)
        return 0;
    const int bw = p.tw < 32 ? p.tw : 32;
    const int bh = p.th < 32 / bw ? p.th : 32 / bw;
    const int bb = 32 / (bw * bh);
    uint64_t dims[4] = {(uint64_t)p.No, (uint64_t)p.out_W, (uint64_t)p.out_H, (uint64_t)p.n_img};
    uint64_t str[4] = {1, (uint64_t)p.ldo, (uint64_t)p.out_W * p.ldo, (uint64_t)p.out_H * p.out_W * p.ldo};
    uint32_t box[4] = {32, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
    if (make_map(mo, p.out, 4, dims, str, box, nullptr) != 0) return 0;
    if (p.residual && make_map(mr, p.residual, 4, dims, str, box, nullptr) != 0) return 0;
    if (p.relu_mask && make_map(mm, p.relu_mask, 4, dims, str, box, nullptr) != 0) return 0;
    return (p.residual && p.relu_mask) ? 3 : (operands ? 2 : 1);
}

// Launch the bf16x3 fprop / dgrad kernel with the epilogue variant the output allows.
int launch_bf(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, dim3 grid, int bn, cudaStream_t stream) {
    CUtensorMap mo, mr, mm;
    const int ts = make_epilogue_maps(&mo, &mr, &mm, p);
    const CUtensorMap* r = p.residual ? &mr : nullptr;
    const CUtensorMap* m = p.relu_mask ? &mm : nullptr;
    static const bool ew8 = getenv("MDB_EPI_WARPS_4") == nullptr;                 // A/B switch (profiling)
    if (ew8) {   // 8 epilogue warps; stages sized so that ring + patches fill the 227 KB
        if (ts == 3) return bn == 64 ? launch_tc<64, 4, 0, false, true, true, true, 3, 8>(ma, mb, p, grid, stream, &mo, r, m)
                                     : launch_tc<128, 3, 0, false, true, true, true, 3, 8>(ma, mb, p, grid, stream, &mo, r, m);
        if (ts == 2) return bn == 64 ? launch_tc<64, 5, 0, false, true, true, true, 2, 8>(ma, mb, p, grid, stream, &mo, r, m)
                                     : launch_tc<128, 4, 0, false, true, true, true, 2, 8>(ma, mb, p, grid, stream, &mo, r, m);
        if (ts == 1) return bn == 64 ? launch_tc<64, 6, 0, false, true, true, true, 1, 8>(ma, mb, p, grid, stream, &mo)
                                     : launch_tc<128, 5, 0, false, true, true, true, 1, 8>(ma, mb, p, grid, stream, &mo);
    } else {
        if (ts == 3) return bn == 64 ? launch_tc<64, 6, 0, false, true, true, true, 3, 4>(ma, mb, p, grid, stream, &mo, r, m)
                                     : launch_tc<128, 5, 0, false, true, true, true, 3, 4>(ma, mb, p, grid, stream, &mo, r, m);
        if (ts == 2) return bn == 64 ? launch_tc<64, 6, 0, false, true, true, true, 2, 4>(ma, mb, p, grid, stream, &mo, r, m)
                                     : launch_tc<128, 5, 0, false, true, true, true, 2, 4>(ma, mb, p, grid, stream, &mo, r, m);
        if (ts == 1) return bn == 64 ? launch_tc<64, 8, 0, false, true, true, true, 1, 4>(ma, mb, p, grid, stream, &mo)
                                     : launch_tc<128, 6, 0, false, true, true, true, 1, 4>(ma, mb, p, grid, stream, &mo);
    }
    return bn == 64 ? launch_tc<64, 8, 0, false, true, true, true>(ma, mb, p, grid, stream)
                    : launch_tc<128, 6, 0, false, true, true, true>(ma, mb, p, grid, stream);
}

struct ConvGeom {
    int B, H, W, Cin, Cout, kh, kw, stride, pad, Ho, Wo;
};

int check_geom(const ConvGeom& g, bool forward = false) {
    if (g.B <= 0 || g.H <= 0 || g.W <= 0 || g.Cin <= 0 || g.Cout <= 0) return MDB_EINVAL;
    if (g.kh != g.kw || (g.kh != 1 && g.kh != 3) || (g.stride != 1 && g.stride != 2)) return MDB_EUNSUPPORTED;
    // TMA needs 16-byte row pitches on every operand it reads: Cin always; Cout only where dy / the output map is an
    // operand (dgrad, wgrad).  The forward epilogue writes ragged Cout with scalar stores.
    if (g.Cin % 4 || (!forward && g.Cout % 4)) return MDB_EUNSUPPORTED;
    return 0;
}

}  // namespace

extern "C" {

#ifdef MDB_TIMELINE
void mdb_debug_set_timeline(long long* device_buf) { g_dbg = device_buf; }
#endif

int mdb_set_precision(int mode) {
    if (mode < 0 || mode > 2) return MDB_EINVAL;
    g_precision = mode;
    return 0;
}
int mdb_get_precision(void) { return g_precision; }

}  // extern "C"

namespace {

// Split-K decision of the forward (shared by the launcher and mdb_conv2d_forward_workspace_bytes): a handful of tiles
// with a very long reduction (the 3x3 stride-2 2048->256 neck convolution: 16-32 tiles x 576 k-blocks kept a fifth of
// the SMs busy for 0.3 ms).  Returns the number of slices (0 = no split-K).
int forward_splitk_slices(int precision, int tiles, int kblocks, int bn, bool plain_epilogue, int Cout, long long out_elems) {
    const int kb_per_slice = 32;
    if (precision == 0 || bn != 128 || tiles * 2 > num_sms_tc() || kblocks < 256 || !plain_epilogue || Cout % 4) return 0;
    const int slices = (kblocks + kb_per_slice - 1) / kb_per_slice;
    return (out_elems * slices < (1ll << 31)) ? slices : 0;
}

// y[B,Ho,Wo,Cout] = act( conv(x[B,H,W,Cin], w) + bias + residual );  w = fp32 packed [kh*kw][Cout][Cin] (bf == false)
// or the pre-split bf16 form [kh*kw][Cout][ceil(Cin/32)][hi 32 | lo 32] (bf == true, precision mode 2).
int conv_forward_impl(const float* x, const void* w_packed, bool bf, const float* bias, const float* residual, float* y,
                      int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int flags,
                      void* stream_, size_t* query_ws = nullptr /* non-null: only report the scratch bytes this call needs */) {
    const int precision = bf ? 2 : (g_precision == 2 ? 1 : g_precision);   // fp32 weights in bf16x3 mode: 3xTF32
    ConvGeom g{B, H, W, Cin, Cout, kh, kw, stride, pad, (H + 2 * pad - kh) / stride + 1, (W + 2 * pad - kw) / stride + 1};
    int rc = check_geom(g, true);
    if (rc) return rc;
    if (!query_ws && (!x || !w_packed || !y)) return MDB_EINVAL;
    if ((unsigned long long)B * g.Ho * g.Wo * Cout >= (1ull << 31)) return MDB_EUNSUPPORTED;   // the epilogue indexes with (signed) 32 bits
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (kh == 1 && stride == 1 && pad == 0) {     // pointwise: the batch of images is one long row of pixels (no tile waste)
        W = B * H * W; H = 1; B = 1;
        g = ConvGeom{B, H, W, Cin, Cout, kh, kw, stride, pad, H, W};
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    pick_tile3(g.Wo, g.Ho, B, BM, &p.tw, &p.th, &p.tb);
    p.tiles_x = (g.Wo + p.tw - 1) / p.tw;
    p.tiles_y = (g.Ho + p.th - 1) / p.th;
    p.n_img = B;
    const int n_groups = (B + p.tb - 1) / p.tb;
    p.Ho = g.Ho; p.Wo = g.Wo;
    p.out_sy = p.out_sx = 1; p.out_oy = p.out_ox = 0; p.out_H = g.Ho; p.out_W = g.Wo;
    p.in_sy = p.in_sx = stride;
    p.ntaps = kh * kw;
    p.cblocks = (Cin + BK - 1) / BK;
    for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) {
            const int t = ky * kw + kx;
            p.tap_dy[t] = ky - pad; p.tap_dx[t] = kx - pad; p.tap_w[t] = t;
        }
    p.wg_taps = 1; p.wg_kw = 1;
    p.No = Cout; p.ldo = Cout; p.relu = flags & 1; p.round_out = (precision == 0) ? ((flags >> 1) & 1) : 0; p.atomic_out = 0;
    p.bias = bias; p.residual = residual; p.relu_mask = nullptr; p.rowscale = nullptr; p.out = y; p.dbg = g_dbg;

    // 128x256 3xTF32 tiles are only used when the TMEM-A variant is switched off (it is faster than them everywhere measured)
    static const bool wide_split = getenv("MDB_NO_TMEM_A") != nullptr && getenv("MDB_NO_WIDE_SPLIT") == nullptr;
    static const bool tmem_a = getenv("MDB_NO_TMEM_A") == nullptr;            // A/B switch (profiling)
    // 3xTF32: a 128x256 tile needs 96 KB per stage -> only 2 stages fit, which cannot hide DRAM latency; it pays off
    // only when the A operand is re-read from L2 (multi-tap convolutions), measured +6 % on 3x3 256->256.
    const bool wide = (precision == 0 || (precision == 1 && wide_split && kh * kw > 1)) && (Cout % 256 == 0) &&
                      ((long long)n_groups * p.tiles_x * p.tiles_y * (Cout / 256) >= 100);
    const int bn = wide ? 256 : (Cout <= 64 ? 64 : 128);
    dim3 grid((Cout + bn - 1) / bn, n_groups * p.tiles_x * p.tiles_y, 1);
    const long long out_elems = (long long)B * g.Ho * g.Wo * Cout;
    const int slices = (tmem_a && (reinterpret_cast<uintptr_t>(bias) & 15u) == 0)
                           ? forward_splitk_slices(precision, (int)(grid.x * grid.y), p.ntaps * p.cblocks, bn, !residual && !p.relu, Cout, out_elems)
                           : 0;
    if (query_ws) {
        *query_ws = sizeof(float) * (size_t)out_elems * slices;
        return 0;
    }

    CUtensorMap ma, mb;
    {   // A: x as (C, W, H, B), box (32, tw*s, th*s, 1), element strides (1, s, s, 1)
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        uint64_t str[4] = {1, (uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin};
        uint32_t box[4] = {BK, (uint32_t)(p.tw * stride), (uint32_t)(p.th * stride), (uint32_t)p.tb};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        if (box[1] > 256 || box[2] > 256) return MDB_EUNSUPPORTED;
        rc = make_map(&ma, x, 4, dims, str, box, es);
        if (rc) return rc;
    }
    if (bf) {   // B: pre-split bf16 weights as (64 * k-blocks, Cout, taps), box (64 = one [hi 32 | lo 32] row, BN, 1)
        const uint64_t kb = (uint64_t)p.cblocks;
        uint64_t dims[3] = {64 * kb, (uint64_t)Cout, (uint64_t)(kh * kw)};
        uint64_t str[3] = {1, 64 * kb, (uint64_t)Cout * 64 * kb};
        uint32_t box[3] = {64, (uint32_t)bn, 1};
        rc = make_map(&mb, w_packed, 3, dims, str, box, nullptr, false, true);
        if (rc) return rc;
    } else {   // B: packed weights as (Cin, Cout, taps), box (32, BN, 1)
        uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(kh * kw)};
        uint64_t str[3] = {1, (uint64_t)Cin, (uint64_t)Cout * Cin};
        uint32_t box[3] = {BK, (uint32_t)bn, 1};
        rc = make_map(&mb, w_packed, 3, dims, str, box, nullptr);
        if (rc) return rc;
    }
    // Split-K (see forward_splitk_slices): slices write partial tiles to the caller's scratch buffer and a second kernel adds
    // them in a fixed order (+ bias): bit-reproducible, unlike atomic accumulation.  The slice length is fixed and only
    // reductions >= 256 k-blocks take this path, so the summation order does not depend on the batch size.
    {
        if (slices > 0) {
            float* ws = nullptr;
            rc = splitk_workspace(sizeof(float) * (size_t)out_elems * slices, &ws);
            if (rc) return rc;
            p.kb_per_slice = 32;
            p.slice_stride = (int)out_elems;
            p.out = ws; p.bias = nullptr;
            grid.z = slices;
            rc = bf ? launch_tc<128, 6, 0, false, true, true, true>(ma, mb, p, grid, stream)
                    : launch_tc<128, 4, 0, false, true, true>(ma, mb, p, grid, stream);
            if (rc) return rc;
            const long long n4 = out_elems / 4;
            const int blocks = (int)((n4 + 255) / 256 > 1184 ? 1184 : (n4 + 255) / 256);
            splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(ws), reinterpret_cast<const float4*>(bias),
                                                             reinterpret_cast<float4*>(y), n4, Cout / 4, slices, n4);
            return (int)cudaGetLastError();
        }
    }
    if (bf) return launch_bf(ma, mb, p, grid, bn, stream);
    if (bn == 64) return (precision == 1) ? (tmem_a ? launch_tc<64, 6, 0, false, true, true>(ma, mb, p, grid, stream)
                                                    : launch_tc<64, 4, 0, false, true>(ma, mb, p, grid, stream))
                                          : launch_tc<64, 6, 0, false, false>(ma, mb, p, grid, stream);
    if (precision == 1 && wide) return launch_tc<256, 2, 0, false, true>(ma, mb, p, grid, stream);
    if (precision == 1) return tmem_a ? launch_tc<128, 4, 0, false, true, true>(ma, mb, p, grid, stream)
                                      : launch_tc<128, 3, 0, false, true>(ma, mb, p, grid, stream);
    if (wide) return launch_tc<256, 4, 0, false, false>(ma, mb, p, grid, stream);
    return launch_tc<128, 5, 0, false, false>(ma, mb, p, grid, stream);
}

// dx[B,H,W,Cin] = (conv_transpose(dy[B,Ho,Wo,Cout], w) + residual) * (relu_mask > 0);  w = fp32 packed [taps][Cout][Cin]
// (read MN-major) or, bf == true, the pre-split TRANSPOSED bf16 form [taps][Cin][ceil(Cout/32)][hi 32 | lo 32] (K-major).
int conv_dgrad_impl(const float* dy, const void* w_packed, bool bf, const float* residual, const float* relu_mask,
                    float* dx, int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                    int flags, void* stream_) {
    const int precision = bf ? 2 : (g_precision == 2 ? 1 : g_precision);
    ConvGeom g{B, H, W, Cin, Cout, kh, kw, stride, pad, (H + 2 * pad - kh) / stride + 1, (W + 2 * pad - kw) / stride + 1};
    int rc = check_geom(g);
    if (rc) return rc;
    if (!dy || !w_packed || !dx) return MDB_EINVAL;
    if ((unsigned long long)B * H * W * Cin >= (1ull << 31)) return MDB_EUNSUPPORTED;           // the epilogue indexes with (signed) 32 bits
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (kh == 1 && stride == 1 && pad == 0) {     // pointwise: one long row of pixels
        W = B * H * W; H = 1; B = 1;
        g = ConvGeom{B, H, W, Cin, Cout, kh, kw, stride, pad, H, W};
    }
    // dx[y,x] = sum_{ky,kx} dy[(y+pad-ky)/s, (x+pad-kx)/s] * W[ky,kx]  where the division is exact.
    // Per output parity class (py,px) (only one class for s == 1) the contributing taps are fixed and the
    // dy access is a unit-stride shifted box: oy = (y + pad - ky)/s = j + (py + pad - ky)/s  for y = s*j + py.
    for (int py = 0; py < stride; ++py)
        for (int px = 0; px < stride; ++px) {
            TcParams p;
            memset(&p, 0, sizeof(p));
            const int Hc = (H - py + stride - 1) / stride, Wc = (W - px + stride - 1) / stride;  // pixels in class
            if (Hc <= 0 || Wc <= 0) continue;
            pick_tile3(Wc, Hc, B, BM, &p.tw, &p.th, &p.tb);
            p.tiles_x = (Wc + p.tw - 1) / p.tw;
            p.tiles_y = (Hc + p.th - 1) / p.th;
            p.n_img = B;
            const int n_groups = (B + p.tb - 1) / p.tb;
            p.Ho = Hc; p.Wo = Wc;
            p.out_sy = p.out_sx = stride; p.out_oy = py; p.out_ox = px; p.out_H = H; p.out_W = W;
            p.in_sy = p.in_sx = 1;
            p.cblocks = (Cout + BK - 1) / BK;
            int nt = 0;
            for (int ky = 0; ky < kh; ++ky) {
                if ((py + pad - ky) % stride) continue;
                for (int kx = 0; kx < kw; ++kx) {
                    if ((px + pad - kx) % stride) continue;
                    // floor division is exact here (numerator divisible by stride; may be negative)
                    p.tap_dy[nt] = (py + pad - ky) / stride;
                    p.tap_dx[nt] = (px + pad - kx) / stride;
                    p.tap_w[nt] = ky * kw + kx;
                    ++nt;
                }
            }
            p.ntaps = nt;
            p.wg_taps = 1; p.wg_kw = 1;
            p.No = Cin; p.ldo = Cin; p.relu = 0; p.atomic_out = 0; p.round_out = (precision == 0) ? ((flags >> 1) & 1) : 0;
            p.bias = nullptr; p.residual = residual; p.relu_mask = relu_mask; p.rowscale = nullptr; p.out = dx;
            CUtensorMap ma, mb;
            {   // A: dy as (Cout, Wo, Ho, B), unit stride
                uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)B};
                uint64_t str[4] = {1, (uint64_t)Cout, (uint64_t)g.Wo * Cout, (uint64_t)g.Ho * g.Wo * Cout};
                uint32_t box[4] = {BK, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
                rc = make_map(&ma, dy, 4, dims, str, box, nullptr);
                if (rc) return rc;
            }
            const int bn = (bf && Cin <= 64) ? 64 : 128;
            if (bf) {   // B (K-major): transposed pre-split weights as (64 * k-blocks, Cin, taps); box (64, BN, 1)
                const uint64_t kb = (uint64_t)p.cblocks;
                uint64_t dims[3] = {64 * kb, (uint64_t)Cin, (uint64_t)(kh * kw)};
                uint64_t str[3] = {1, 64 * kb, (uint64_t)Cin * 64 * kb};
                uint32_t box[3] = {64, (uint32_t)bn, 1};
                rc = make_map(&mb, w_packed, 3, dims, str, box, nullptr, false, true);
                if (rc) return rc;
            } else {   // B (MN-major): packed weights as (Cin, Cout, taps); box = 32 output columns (Cin) x 32 reduction rows (Cout)
                uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(kh * kw)};
                uint64_t str[3] = {1, (uint64_t)Cin, (uint64_t)Cout * Cin};
                uint32_t box[3] = {32, 32, 1};
                rc = make_map(&mb, w_packed, 3, dims, str, box, nullptr, true);
                if (rc) return rc;
            }
            dim3 grid((Cin + bn - 1) / bn, n_groups * p.tiles_x * p.tiles_y, 1);
            if (nt == 0) {   // no tap reaches this parity class (1x1 stride 2): result = (0 + residual) * mask
                p.ntaps = 0;
            }
            static const bool tmem_a = getenv("MDB_NO_TMEM_A") == nullptr;
            if (bf) rc = launch_bf(ma, mb, p, grid, bn, stream);
            else rc = (precision == 1) ? (tmem_a ? launch_tc<128, 4, 0, true, true, true>(ma, mb, p, grid, stream)
                                                 : launch_tc<128, 3, 0, true, true>(ma, mb, p, grid, stream))
                                       : launch_tc<128, 5, 0, true, false>(ma, mb, p, grid, stream);
            if (rc) return rc;
        }
    return 0;
}

}  // namespace

extern "C" {

int mdb_conv2d_forward_f32(const float* x, const float* w_packed, const float* bias, const float* residual, float* y,
                           int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int flags,
                           void* stream) {
    return conv_forward_impl(x, w_packed, false, bias, residual, y, B, H, W, Cin, Cout, kh, kw, stride, pad, flags, stream);
}
int mdb_conv2d_forward_bf16x3(const float* x, const void* w_split, const float* bias, const float* residual, float* y,
                              int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int flags,
                              void* stream) {
    return conv_forward_impl(x, w_split, true, bias, residual, y, B, H, W, Cin, Cout, kh, kw, stride, pad, flags, stream);
}
long long mdb_conv2d_forward_workspace_bytes(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                                             int flags, int has_residual, int split_weights) {
    size_t bytes = 0;
    const float* res = has_residual ? reinterpret_cast<const float*>(16) : nullptr;
    const int rc = conv_forward_impl(nullptr, nullptr, split_weights != 0, nullptr, res, nullptr, B, H, W, Cin, Cout, kh, kw, stride,
                                     pad, flags, nullptr, &bytes);
    return rc ? (long long)rc : (long long)bytes;
}
int mdb_set_workspace(void* buf, unsigned long long bytes) {
    DeviceState& d = g_dev[current_device()];
    d.ws = static_cast<float*>(buf);
    d.ws_bytes = buf ? (size_t)bytes : 0;
    return 0;
}
int mdb_conv2d_dgrad_f32(const float* dy, const float* w_packed, const float* residual, const float* relu_mask,
                         float* dx, int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                         int flags, void* stream) {
    return conv_dgrad_impl(dy, w_packed, false, residual, relu_mask, dx, B, H, W, Cin, Cout, kh, kw, stride, pad, flags, stream);
}
int mdb_conv2d_dgrad_bf16x3(const float* dy, const void* w_split_t, const float* residual, const float* relu_mask,
                            float* dx, int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                            int flags, void* stream) {
    return conv_dgrad_impl(dy, w_split_t, true, residual, relu_mask, dx, B, H, W, Cin, Cout, kh, kw, stride, pad, flags, stream);
}

// dw_packed[tap][Cout][Cin] (+)= rowscale[co] * sum_{b,oy,ox} dy[b,oy,ox,co] * x[b, oy*s+ky-pad, ox*s+kx-pad, ci]
// dw_packed is zero-filled by the call unless accumulate != 0.
int mdb_conv2d_wgrad_f32(const float* dy, const float* x, const float* rowscale, float* dw_packed, int B, int H, int W,
                         int Cin, int Cout, int kh, int kw, int stride, int pad, int accumulate, void* stream_) {
    return mdb_conv2d_wgrad_bias_f32(dy, x, rowscale, dw_packed, nullptr, B, H, W, Cin, Cout, kh, kw, stride, pad, accumulate, stream_);
}

// Same, plus db[Cout] (+)= sum over pixels of dy (the bias gradient) -- a by-product of the 3xTF32 kernel's operand
// staging (the splitter thread that moves dy channel m into tensor memory sums it on the way); the single-pass TF32
// mode has no splitter and runs the stand-alone column-sum kernel instead.
int mdb_conv2d_wgrad_bias_f32(const float* dy, const float* x, const float* rowscale, float* dw_packed, float* db, int B, int H,
                              int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int accumulate, void* stream_) {
    ConvGeom g{B, H, W, Cin, Cout, kh, kw, stride, pad, (H + 2 * pad - kh) / stride + 1, (W + 2 * pad - kw) / stride + 1};
    int rc = check_geom(g);
    if (rc) return rc;
    if (!dy || !x || !dw_packed) return MDB_EINVAL;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (kh == 1 && stride == 1 && pad == 0) {     // pointwise: one long row of pixels (full 32-pixel reduction tiles)
        W = B * H * W; H = 1; B = 1;
        g = ConvGeom{B, H, W, Cin, Cout, kh, kw, stride, pad, H, W};
    }
    const int taps = kh * kw;
    static const bool tmem_a = getenv("MDB_NO_TMEM_A") == nullptr;
    const bool fused_db = db && g_precision != 0 && tmem_a;
    if (!accumulate) {
        // db directly behind dw_packed (as monodetr_b200.tc allocates them): one memset node instead of two
        const bool adjacent = fused_db && db == dw_packed + (size_t)taps * Cout * Cin;
        cudaError_t e = cudaMemsetAsync(dw_packed, 0, sizeof(float) * ((size_t)taps * Cout * Cin + (adjacent ? Cout : 0)), stream);
        if (e == cudaSuccess && fused_db && !adjacent) e = cudaMemsetAsync(db, 0, sizeof(float) * Cout, stream);
        if (e != cudaSuccess) return (int)e;
    }
    if (db && !fused_db) {
        rc = mdb_colsum_f32(dy, db, (long long)B * g.Ho * g.Wo, Cout, accumulate, stream_);
        if (rc) return rc;
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.colsum_out = fused_db ? db : nullptr;
    pick_tile(g.Wo, g.Ho, 32, &p.rtw, &p.rth);
    p.rtiles_x = (g.Wo + p.rtw - 1) / p.rtw;
    p.rtiles_y = (g.Ho + p.rth - 1) / p.rth;
    p.n_img = B;
    const int total_red = B * p.rtiles_x * p.rtiles_y;
    const int bn = (g_precision == 0 && Cin >= 256) ? 256 : 128;
    const int tiles = ((Cout + BM - 1) / BM) * ((Cin + bn - 1) / bn) * taps;
    // split-K over pixel tiles.  Large problems: ~2 waves of CTAs, but at least ~24 reduction steps per CTA so the
    // 128 x BN atomic epilogue stays a small fraction of the work.  Small problems (decoder / head linears, M = 4400 rows
    // = 138 steps): the launch is latency-bound, so spread it over up to one wave with >= 8 steps per CTA (measured
    // 23 us -> ~10 us per launch, ~200 such launches per step).
    const int sms = num_sms_tc();
    int big = (2 * sms + tiles - 1) / tiles, small = sms / tiles;
    if (big > total_red / 24) big = total_red / 24;
    if (small > total_red / 8) small = total_red / 8;
    int splits = big > small ? big : small;
    if (splits < 1 || mdb_get_deterministic()) splits = 1;      // reproducible mode: exactly one accumulation per output element
    p.red_per_split = (total_red + splits - 1) / splits;
    splits = (total_red + p.red_per_split - 1) / p.red_per_split;
    p.w_sy = p.w_sx = stride;
    p.wg_taps = taps; p.wg_kw = kw; p.wg_pad = pad;
    p.Mo_rows = Cout; p.No = Cin; p.ldo = Cin; p.relu = 0; p.atomic_out = 1;
    p.bias = nullptr; p.residual = nullptr; p.relu_mask = nullptr; p.rowscale = rowscale;
    p.out = dw_packed;

    CUtensorMap ma, mb;
    {   // A (MN-major): dy as (Cout, Wo, Ho, B); box = 32 channels x (rtw x rth) reduction pixels
        uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)B};
        uint64_t str[4] = {1, (uint64_t)Cout, (uint64_t)g.Wo * Cout, (uint64_t)g.Ho * g.Wo * Cout};
        uint32_t box[4] = {32, (uint32_t)p.rtw, (uint32_t)p.rth, 1};
        rc = make_map(&ma, dy, 4, dims, str, box, nullptr, true);
        if (rc) return rc;
    }
    {   // B (MN-major): x as (Cin, W, H, B) with element strides s
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        uint64_t str[4] = {1, (uint64_t)Cin, (uint64_t)W * Cin, (uint64_t)H * W * Cin};
        uint32_t box[4] = {32, (uint32_t)(p.rtw * stride), (uint32_t)(p.rth * stride), 1};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        rc = make_map(&mb, x, 4, dims, str, box, es, true);
        if (rc) return rc;
    }
    dim3 grid((Cin + bn - 1) / bn, (Cout + BM - 1) / BM, taps * splits);
    static const bool bf_wgrad = getenv("MDB_NO_BF16_WGRAD") == nullptr;         // A/B switch (profiling)
    if (g_precision == 2 && bf_wgrad) return launch_tc<128, 4, 1, true, true, true, true>(ma, mb, p, grid, stream);
    rc = (g_precision != 0) ? (tmem_a ? launch_tc<128, 4, 1, true, true, true>(ma, mb, p, grid, stream)
                                      : launch_tc<128, 3, 1, true, true>(ma, mb, p, grid, stream))
         : (bn == 256)      ? launch_tc<256, 4, 1, true, false>(ma, mb, p, grid, stream)
                            : launch_tc<128, 5, 1, true, false>(ma, mb, p, grid, stream);
    if (rc) return rc;
    return 0;
}

}  // extern "C"
