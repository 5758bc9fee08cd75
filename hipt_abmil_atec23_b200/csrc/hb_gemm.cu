// hb_gemm.cu — persistent, warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] * W[N,K]^T)
//
//   warp 0     : TMA producer  (A tile 128x64 and W tile per stage, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1     : MMA issuer    (one elected thread, tcgen05.mma kind::f16, K = 16 per instruction, accumulators in TMEM,
//                               double buffered against the epilogue)
//   warp 2     : TMEM allocator
//   warp 3     : residual-tile loader (RESID epilogue: fp32 residual tiles prefetched by TMA into a ring)
//   warps 4-19 : epilogue      (four warpgroups; tcgen05.ld -> fused epilogue -> swizzled smem -> TMA store)
//
// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile; each
// CTA stages its own 128 rows of A and HALF of the W tile, the leader issues the MMAs for both.
//
// This replaces the nn.Linear call sites of the reference's ViT blocks (HIPT_4K/vision_transformer.py:93-95 fc1/fc2,
// :114-116 qkv/proj; HIPT_4K/vision_transformer4k.py:169 phi) and, fed with an im2col'd region, the patch-embed
// convolution (vision_transformer.py:165-169).  nn.Linear.weight is [out,in] = [N,K] K-major, which is exactly the
// UMMA "B K-major" operand, so weights are used as stored (cast to bf16 once at load).
//
// Fused epilogues (HB_EPI_* in include/hipt_b200.h):
//   BIAS_BF16 / BIAS_GELU(_FAST)_BF16   y = act(acc + bias)                                    -> bf16, TMA store
//   BIAS_RESADD_F32                     x += acc + bias                                        -> fp32, TMA reduce-add
//   TOKENS(_GELU)_F32                   token rows with CLS-slot remap + positional table      -> fp32 (+ bf16 copy, row stats)
//   LNFOLD(_GELU)_BF16                  the LayerNorm that precedes the Linear in Block.forward (vision_transformer.py:
//                                       147,151) folded in: A is the UN-normalised residual stream in bf16, W carries
//                                       gamma, and  y = rstd_r * acc - rstd_r * mu_r * c_j + d_j  with per-row
//                                       (mu, rstd) from the row sums the producer epilogue left behind
//   RESID_STATS_F32                     x_new = x_old + acc + bias  (x + drop_path(y), :149,151): residual tile fetched
//                                       by TMA, fp32 x_new and its bf16 copy stored by TMA, per-row (sum, sum of
//                                       squares) of x_new accumulated for the next LayerNorm
#include <stdlib.h>

#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

#ifdef HB_EXP_TRACE
__device__ long long g_trace[4 * 1024];
#define TRACE(role, it, k) do { if (blockIdx.x == 0 && (it) < 120) g_trace[(role) * 1024 + (it) * 8 + (k)] = clock64(); } while (0)
#else
#define TRACE(role, it, k) do { } while (0)
#endif

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;          // 64 bf16 = 128 B = one swizzle span
constexpr int GEMM_EPI_WGS = 4;      // epilogue warpgroups
constexpr int GEMM_EPI_THREADS = 128; // per epilogue warpgroup
constexpr int GEMM_THREADS = 128 + GEMM_EPI_WGS * GEMM_EPI_THREADS;   // 20 warps: producer, mma, tmem-alloc, residual loader, 4 x 4 epilogue
constexpr int STAGE_BYTES_OUT = 128 * 128;   // 128 rows x 128 B staging chunk for the TMA store
constexpr int RESID_RING = 4;        // residual tiles (fp32 128 x 32) prefetched ahead of the epilogue
constexpr int SMEM_LIMIT = 232448;   // 227 KB
constexpr uint32_t BAR_EPI_ALL = 5;  // named barrier over every epilogue thread (1..4: one per warpgroup)

// Epilogue kinds (mirrored in include/hipt_b200.h as HB_EPI_*)
enum : int { EPI_BIAS_BF16 = 0, EPI_BIAS_GELU_BF16 = 1, EPI_BIAS_RESADD_F32 = 2, EPI_TOKENS_F32 = 3, EPI_TOKENS_GELU_F32 = 4,
              EPI_BIAS_GELU_FAST_BF16 = 5, EPI_LNFOLD_BF16 = 6, EPI_LNFOLD_GELU_BF16 = 7, EPI_RESID_STATS_F32 = 8,
              EPI_LNFOLD_GELU2_BF16 = 9, EPI_RESID_BF16 = 10 };

template <int BN, int CG, int EPI, int AST = 0>
struct GemmCfg {
    static constexpr bool RESID = (EPI == EPI_RESID_STATS_F32);
    static constexpr bool RESB = (EPI == EPI_RESID_BF16);
    static constexpr bool TOKENS = (EPI == EPI_TOKENS_F32 || EPI == EPI_TOKENS_GELU_F32);
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_ROWS = BN / CG;                    // W rows staged by this CTA
    static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
    // output side: one 16 KB staging tile per warpgroup; the residual epilogue instead has a ring of fp32 residual
    // tiles (updated in place and stored from there) plus an 8 KB bf16 staging tile per warpgroup; token rows go
    // straight to global memory
    static constexpr int OUT_BYTES = RESID ? (RESID_RING * STAGE_BYTES_OUT + GEMM_EPI_WGS * 8192)
                                           : (TOKENS ? 0 : GEMM_EPI_WGS * STAGE_BYTES_OUT);
    static constexpr int VEC_BYTES = (RESID || TOKENS || EPI == EPI_BIAS_RESADD_F32) ? GEMM_EPI_WGS * 128 * 4 : 16 * 256 * 4;                                        // bias / c slices per warp(group)
    static constexpr int STAT_BYTES = (RESID || TOKENS) ? 2 * GEMM_EPI_WGS * 128 * 8 : 0; // row-stat exchange
    static constexpr int TAIL_BYTES = VEC_BYTES + STAT_BYTES + 640 /* barriers */;
    // AST (A-stationary, K = 384): the whole [128 x 384] A tile stays resident for a group of n-tiles, only W is streamed
    static constexpr int A_RES = AST ? 6 * A_BYTES : 0;
    static constexpr int STAGE_BYTES = AST ? B_BYTES : (A_BYTES + B_BYTES);
    static constexpr int RING_BUDGET = SMEM_LIMIT - 1024 - OUT_BYTES - TAIL_BYTES - A_RES;
    static constexpr int STAGES_FIT = RING_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = 1024 + A_RES + STAGES * STAGE_BYTES + OUT_BYTES + TAIL_BYTES;
    static_assert(STAGES >= 2, "operand ring too shallow");
};

template <int BN, int EPI, int CG, int AST>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_xb,
                 const float* __restrict__ bias, const float* __restrict__ tok_table, float* __restrict__ tok_out,
                 const GemmAux aux, int M, int N, int K, int tokens_per_seq, int n_group) {
    using Cfg = GemmCfg<BN, CG, EPI, AST>;
    static_assert(!AST || CG == 2, "the A-stationary variant runs as CTA pairs");
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool RESB = Cfg::RESB;                    // bf16 residual stream: x = bf16(x + acc + bias), row statistics
    constexpr bool OUT_BF16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_FAST_BF16 ||
                               EPI == EPI_LNFOLD_BF16 || EPI == EPI_LNFOLD_GELU_BF16 || EPI == EPI_LNFOLD_GELU2_BF16 || RESB);
    constexpr bool LNFOLD = (EPI == EPI_LNFOLD_BF16 || EPI == EPI_LNFOLD_GELU_BF16 || EPI == EPI_LNFOLD_GELU2_BF16);
    constexpr bool OUT_TOKENS = Cfg::TOKENS;
    constexpr bool RESID = Cfg::RESID;
    constexpr int CHUNK_COLS = OUT_BF16 ? 64 : 32;      // 128 B of output per row per chunk
    constexpr int NCHUNK = BN / CHUNK_COLS;
    constexpr int NWG = GEMM_EPI_WGS;

    // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked below), and deriving
    // every pointer from the array itself keeps the shared state space (LDS/STS instead of generic accesses)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* smem_a = smem;                                   // AST: resident [6 k-blocks][128 x 128 B]; else the A ring
    uint8_t* smem_b = smem + (AST ? Cfg::A_RES : STAGES * Cfg::A_BYTES);
    uint8_t* smem_o = smem_b + STAGES * Cfg::B_BYTES;
    float* vec_x = reinterpret_cast<float*>(smem_o + Cfg::OUT_BYTES);                 // [NWG][128]
    float* stat_x = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(vec_x) + Cfg::VEC_BYTES);   // [2][NWG][128][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stat_x) + Cfg::STAT_BYTES);
    uint64_t* full_bar = bars;                  // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]  MMA -> TMA
    uint64_t* acc_full = bars + 2 * STAGES;     // [2]       MMA -> epilogue
    uint64_t* acc_empty = bars + 2 * STAGES + 2;// [2]       epilogue -> MMA
    uint64_t* ring_full = bars + 2 * STAGES + 4;             // [RESID_RING] residual tile landed (RESID only)
    uint64_t* ring_empty = ring_full + RESID_RING;           // [RESID_RING] its updated copy has been stored
    uint64_t* res_full = ring_empty + RESID_RING;            // [16] per epilogue warp: bf16 residual chunk landed (RESB)
    uint64_t* a_full = res_full + 16;                        // [6] AST: resident A k-block landed (leader collects both CTAs)
    uint64_t* a_empty = a_full + 6;                          // [6] AST: the last n-tile of the group is done with it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs of the pair)
    const int tile0 = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
    const int tile_step = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;
    constexpr int TILE_M = GEMM_BM * CG;
    if constexpr (CG == 2) cluster_sync_all();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
        if (!OUT_TOKENS) tma_prefetch_desc(&map_out);
        if (RESID || RESB) tma_prefetch_desc(&map_xb);
    }
    if (warp == 1 && lane == 0) {
        // CG = 2: full_bar / acc_empty are used on the leader only and collect arrivals from both CTAs
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], OUT_BF16 ? CG * 8 : CG * NWG * GEMM_EPI_THREADS); }
        for (int i = 0; i < RESID_RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
        for (int i = 0; i < 16; ++i) mbar_init(&res_full[i], 1);
        for (int i = 0; i < 6; ++i) { mbar_init(&a_full[i], CG); mbar_init(&a_empty[i], 1); }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (CG == 2) { tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_tiles = N / BN;
    const int m_tiles = (M + TILE_M - 1) / TILE_M;
    const int total_tiles = m_tiles * n_tiles;
    const int k_blocks = K / GEMM_BK;
    // Tile sequence of this CTA (pair): local tile `it` = (unit ui = it / NG, t = it % NG); unit = tile0 + ui * tile_step
    // covers the n-tiles [g * NG, g * NG + NG) of one m-tile.  NG = 1 (groups = n_tiles) is the plain tile-strided order.
    const int NG = AST ? n_group : 1;
    const int groups = n_tiles / NG;
    const int total_units = m_tiles * groups;
    auto tile_at = [&](uint32_t it_, int& m_idx_, int& n_idx_) -> bool {
        const uint32_t ui = it_ / NG, t = it_ - ui * NG;
        const int unit = tile0 + static_cast<int>(ui) * tile_step;
        if (unit >= total_units) return false;
        m_idx_ = unit / groups;
        n_idx_ = (unit - m_idx_ * groups) * NG + static_cast<int>(t);
        return true;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (AST && lane == 0) {
            // A-stationary: this thread streams W only; the A tiles are loaded by warp 2
            const uint64_t pol_w = policy_evict_last();
            uint32_t stage = 0, phase = 0;
            int m_idx, n_idx;
            for (uint32_t it = 0; tile_at(it, m_idx, n_idx); ++it) {
                const int n0 = n_idx * BN + cta_rank * Cfg::B_ROWS;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    tma_load_2d_2sm_hint(smem_b + stage * Cfg::B_BYTES, &map_w, &full_bar[stage], kb * GEMM_BK, n0, pol_w);
                    if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::B_BYTES);
                    else mbar_arrive_remote(&full_bar[stage], 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (lane == 0) {
            const uint64_t pol_w = policy_evict_last();
            uint32_t stage = 0, phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int m0 = (tile / n_tiles) * TILE_M + cta_rank * GEMM_BM;
                const int n0 = (tile % n_tiles) * BN + cta_rank * Cfg::B_ROWS;
                TRACE(3, (tile - tile0) / tile_step, 0);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (kb == k_blocks - 1) TRACE(3, (tile - tile0) / tile_step, 1);
#ifdef HB_EXP_NOLOAD
                    if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&full_bar[stage], 0); else mbar_arrive(&full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    continue;
#endif
                    if constexpr (CG == 2) {
                        // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the pair
#ifdef HB_EXP_NOW
                        tma_load_2d_2sm(smem_a + stage * Cfg::A_BYTES, &map_a, &full_bar[stage], kb * GEMM_BK, m0);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES));
#elif defined(HB_EXP_NOA)
                        tma_load_2d_2sm_hint(smem_b + stage * Cfg::B_BYTES, &map_w, &full_bar[stage], kb * GEMM_BK, n0, pol_w);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::B_BYTES));
#else
                        tma_load_2d_2sm(smem_a + stage * Cfg::A_BYTES, &map_a, &full_bar[stage], kb * GEMM_BK, m0);
                        tma_load_2d_2sm_hint(smem_b + stage * Cfg::B_BYTES, &map_w, &full_bar[stage], kb * GEMM_BK, n0, pol_w);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
#endif
                        else mbar_arrive_remote(&full_bar[stage], 0);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
                        tma_load_2d(smem_a + stage * Cfg::A_BYTES, &map_a, &full_bar[stage], kb * GEMM_BK, m0);
                        tma_load_2d_hint(smem_b + stage * Cfg::B_BYTES, &map_w, &full_bar[stage], kb * GEMM_BK, n0, pol_w);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the pipeline (warp-uniform control flow and operands, so the descriptors live in
        // uniform registers); one elected lane issues the tcgen05.mma / commit instructions.
        if (AST && cta_rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
            const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
            uint32_t stage = 0, phase = 0;
            int m_idx, n_idx;
            for (uint32_t it = 0; tile_at(it, m_idx, n_idx); ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                const uint32_t ui = it / NG, t = it - ui * NG;
                mbar_wait(&acc_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < 6; ++kb) {
                    if (t == 0) mbar_wait(&a_full[kb], ui & 1);          // first n-tile of the group: A k-block landed
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k128(a_base + kb * Cfg::A_BYTES);
                    const uint64_t db = umma_desc_k128(b_base + stage * Cfg::B_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k)
                            umma_bf16_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_commit_2sm(&empty_bar[stage]);
                        if (t == static_cast<uint32_t>(NG) - 1) umma_commit_2sm(&a_empty[kb]);   // group done with this A k-block
                        if (kb == 5) umma_commit_2sm(&acc_full[as]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (cta_rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
            const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
            uint32_t stage = 0, phase = 0;
            uint32_t it = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                if (lane == 0) TRACE(0, it, 0);
                mbar_wait(&acc_empty[as], aphase ^ 1);
                if (lane == 0) TRACE(0, it, 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k128(a_base + stage * Cfg::A_BYTES);
                    const uint64_t db = umma_desc_k128(b_base + stage * Cfg::B_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            // advance 16 bf16 = 32 B inside the swizzle span: +2 in the (addr >> 4) field
                            if constexpr (CG == 2) umma_bf16_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            else umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        if constexpr (CG == 2) {                   // arrive on the same barrier of BOTH CTAs
                            umma_commit_2sm(&empty_bar[stage]);
                            if (kb == k_blocks - 1) umma_commit_2sm(&acc_full[as]);
                        } else {
                            umma_commit(&empty_bar[stage]);        // smem slot free once these MMAs retire
                            if (kb == k_blocks - 1) umma_commit(&acc_full[as]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0 && kb == 0) TRACE(0, it, 2);
                    if (lane == 0 && kb == k_blocks - 1) TRACE(0, it, 3);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ A-tile loader (A-stationary variant only)
        // One [128 x 384] A tile per unit, k-block by k-block as the previous group's last n-tile releases them, so the
        // reload overlaps that tile's remaining MMAs.
        if constexpr (AST) {
            if (lane == 0) {
                for (uint32_t ui = 0;; ++ui) {
                    const int unit = tile0 + static_cast<int>(ui) * tile_step;
                    if (unit >= total_units) break;
                    const int m0 = (unit / groups) * TILE_M + cta_rank * GEMM_BM;
                    for (int kb = 0; kb < 6; ++kb) {
                        mbar_wait(&a_empty[kb], (ui & 1) ^ 1);
                        tma_load_2d_2sm(smem_a + kb * Cfg::A_BYTES, &map_a, &a_full[kb], kb * GEMM_BK, m0);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&a_full[kb], 2 * Cfg::A_BYTES);
                        else mbar_arrive_remote(&a_full[kb], 0);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ residual loader (RESID only)
        // Streams the fp32 residual tiles of every (tile, chunk) of this CTA, in consumption order, into the ring:
        // it runs up to RESID_RING chunks (one whole 192-wide tile) ahead of the epilogue warpgroups.
        if constexpr (RESID) {
            if (lane == 0) {
                uint32_t q = 0;
                for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                    const int m0 = (tile / n_tiles) * TILE_M + cta_rank * GEMM_BM;
                    const int n0 = (tile % n_tiles) * BN;
                    for (int c = 0; c < NCHUNK; ++c, ++q) {
                        const uint32_t slot = q % RESID_RING, use = q / RESID_RING;
                        mbar_wait(&ring_empty[slot], (use & 1) ^ 1);
                        mbar_arrive_expect_tx(&ring_full[slot], STAGE_BYTES_OUT);
                        tma_load_2d(smem_o + slot * STAGE_BYTES_OUT, &map_out, &ring_full[slot], n0 + c * 32, m0);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue, bf16 output: 16 independent warps
        // The two TMEM accumulator buffers are served by disjoint sets of warps: warpgroups 0-1 drain the even tiles of
        // this CTA, warpgroups 2-3 the odd ones, so that while one set is in its math phase the other waits for its MMAs,
        // loads TMEM or stores - the four warps of an SM sub-partition are never all in the same phase.  Within a set,
        // warp = (TMEM lane quadrant ew, column slot s): 32 accumulator rows x the 64-column chunks c = s' (mod 2),
        // s' = s + k (mod 2) alternating per tile so that odd chunk counts balance.  Nothing synchronises across warps:
        // each has its own 4 KB staging tile, column-vector slices and TMA stores.
        if constexpr (OUT_BF16) {
            const int wgi = (warp - 4) >> 2;
            const int par = wgi >> 1;                                  // accumulator buffer served by this warp
            const int s = wgi & 1;
            const int ew = warp & 3;
            const int sw = lane & 7;
            uint8_t* stage_buf = smem_o + (warp - 4) * 4096;          // [32 rows][128 B], SWIZZLE_128B
            const uint32_t row_addr = smem_u32(stage_buf) + lane * 128;
            float* wvec = vec_x + (warp - 4) * 256;                   // [2 chunks][bias 64 | c 64]
            int m_idx = 0, n_idx = 0;
            bool have_tile = tile_at(par, m_idx, n_idx);                 // this warp's tiles: it = par, par + 2, ...
            float pf_v[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            float2 pf_sq = make_float2(0.f, 0.f);
            auto pf_load = [&](int mi, int ni, uint32_t kk) {          // side inputs of a tile, fetched one tile ahead
                const int sp = (s + kk) & 1;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = sp + 2 * cc;
                    if (c < NCHUNK) {
                        const int col = ni * BN + c * 64 + lane;
                        pf_v[cc][0] = __ldg(bias + col);
                        pf_v[cc][1] = __ldg(bias + col + 32);
                        if constexpr (LNFOLD) {
                            pf_v[cc][2] = __ldg(aux.colvec2 + col);
                            pf_v[cc][3] = __ldg(aux.colvec2 + col + 32);
                        }
                    }
                }
                if constexpr (LNFOLD) {
                    // (sum, sum of squares) of the row: aux.n_part partial sums left by the producing epilogues
                    const int r = mi * TILE_M + cta_rank * GEMM_BM + ew * 32 + lane;
                    pf_sq = make_float2(0.f, 0.f);
                    if (r < M) {
                        const float2* sp2 = reinterpret_cast<const float2*>(aux.row_stats) + r;
#pragma unroll
                        for (int pp = 0; pp < 6; ++pp) {
                            if (pp < aux.n_part) {
                                const float2 t = __ldg(sp2 + static_cast<size_t>(pp) * aux.stats_stride);
                                pf_sq.x += t.x; pf_sq.y += t.y;
                            }
                        }
                    }
                }
            };
            if (have_tile) pf_load(m_idx, n_idx, 0);
            [[maybe_unused]] uint32_t res_use = 0;                     // RESB: residual chunks received so far
            for (uint32_t k = 0; have_tile; ++k) {
                const int m0 = m_idx * TILE_M + cta_rank * GEMM_BM + ew * 32;
                const int n0 = n_idx * BN;
                const uint32_t as = par, aphase = k & 1;
                const int sp = (s + k) & 1;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
                // the side inputs fetched one tile ago go to this warp's shared slices; then fetch the next tile's
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    wvec[cc * 128 + lane] = pf_v[cc][0]; wvec[cc * 128 + 32 + lane] = pf_v[cc][1];
                    if constexpr (LNFOLD) { wvec[cc * 128 + 64 + lane] = pf_v[cc][2]; wvec[cc * 128 + 96 + lane] = pf_v[cc][3]; }
                }
                float rstd = 0.f, nrm = 0.f;
                if constexpr (LNFOLD) {
                    const float mu = pf_sq.x * aux.inv_dim;
                    const float var = fmaxf(pf_sq.y * aux.inv_dim - mu * mu, 0.f);
                    rstd = rsqrtf(var + aux.eps);
                    nrm = -rstd * mu;
                }
                have_tile = tile_at(par + 2 * (k + 1), m_idx, n_idx);   // next tile of this warp (m_idx / n_idx now refer to it)
                if (have_tile) pf_load(m_idx, n_idx, k + 1);
                [[maybe_unused]] const f32x2_t rstd2 = f2_pack(rstd, rstd), nrm2 = f2_pack(nrm, nrm);
                if constexpr (RESB) {
                    // residual chunk of the tile's first column chunk, fetched while the MMAs are still running (the
                    // staging tile doubles as its landing buffer: the previous store must have drained it)
                    if (lane == 0) {
                        tma_store_wait_read<0>();
                        mbar_arrive_expect_tx(&res_full[warp - 4], 4096);
                        tma_load_2d(stage_buf, &map_xb, &res_full[warp - 4], n0 + sp * 64, m0);
                    }
                }

                if (lane == 0 && (warp == 4 || warp == 12)) TRACE(1 + (warp == 12), k, 0);
                mbar_wait(&acc_full[as], aphase);
                if (lane == 0 && (warp == 4 || warp == 12)) TRACE(1 + (warp == 12), k, 1);
                tc_fence_after();
                __syncwarp();                                        // the column-vector slices are visible to all lanes
#ifdef HB_EXP_NOEPI
                tc_fence_before();
                if (lane == 0) { if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&acc_empty[as], 0); else mbar_arrive(&acc_empty[as]); }
                continue;
#endif
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = sp + 2 * cc;
                    if (c >= NCHUNK) break;
                    if (lane == 0 && (warp == 4 || warp == 12)) TRACE(1 + (warp == 12), k, 2 + 3 * cc);
                    const float* wv = wvec + cc * 128;
                    if constexpr (RESB) {
                        // ---- x_new = bf16(x_old + acc + bias) in place in the staging tile, partial row statistics
                        if (cc == 1 && lane == 0) {                  // second chunk: its residual can only land now
                            tma_store_wait_read<0>();
                            mbar_arrive_expect_tx(&res_full[warp - 4], 4096);
                            tma_load_2d(stage_buf, &map_xb, &res_full[warp - 4], n0 + c * 64, m0);
                        }
                        float st_sum = 0.f, st_sq = 0.f;
                        mbar_wait(&res_full[warp - 4], res_use & 1);
                        ++res_use;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t v[32];
                            tmem_ld_32x32(taddr + c * 64 + h * 32, v);
                            tmem_ld_wait();
                            if (h == 1 && c + 2 >= NCHUNK) {         // last TMEM read of the tile: hand the accumulator back
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) {
                                    if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&acc_empty[as], 0);
                                    else mbar_arrive(&acc_empty[as]);
                                }
                            }
                            const float4* b4 = reinterpret_cast<const float4*>(wv + h * 32);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {            // 8 columns = one 16-byte piece of the row
                                const uint32_t addr = row_addr + (((h * 4 + q) ^ sw) << 4);
                                uint4 rr;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w) : "r"(addr));
                                const float4 ba = b4[2 * q], bb = b4[2 * q + 1];
                                float o[8];
                                o[0] = __uint_as_float(rr.x << 16) + __uint_as_float(v[8 * q + 0]) + ba.x;
                                o[1] = __uint_as_float(rr.x & 0xffff0000u) + __uint_as_float(v[8 * q + 1]) + ba.y;
                                o[2] = __uint_as_float(rr.y << 16) + __uint_as_float(v[8 * q + 2]) + ba.z;
                                o[3] = __uint_as_float(rr.y & 0xffff0000u) + __uint_as_float(v[8 * q + 3]) + ba.w;
                                o[4] = __uint_as_float(rr.z << 16) + __uint_as_float(v[8 * q + 4]) + bb.x;
                                o[5] = __uint_as_float(rr.z & 0xffff0000u) + __uint_as_float(v[8 * q + 5]) + bb.y;
                                o[6] = __uint_as_float(rr.w << 16) + __uint_as_float(v[8 * q + 6]) + bb.z;
                                o[7] = __uint_as_float(rr.w & 0xffff0000u) + __uint_as_float(v[8 * q + 7]) + bb.w;
#pragma unroll
                                for (int e = 0; e < 8; ++e) { st_sum += o[e]; st_sq = fmaf(o[e], o[e], st_sq); }
                                sts_u4(addr, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                        pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
                            }
                        }
                        {
                            const int row = m0 + lane;
                            const int slot = (n0 >> 6) + c;
                            if (row < M)
                                *reinterpret_cast<float2*>(aux.stats_out + (static_cast<size_t>(slot) * aux.stats_stride + row) * 2) =
                                    make_float2(st_sum, st_sq);
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&map_out, stage_buf, n0 + c * 64, m0);
                            tma_store_commit();
                        }
                        continue;
                    }
                    uint32_t pk[32];                 // 64 bf16 of this row
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + c * 64 + h * 32, v);
                        tmem_ld_wait();
                        if (h == 1 && c + 2 >= NCHUNK) {             // last TMEM read of the tile: hand the accumulator back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&acc_empty[as], 0);
                                else mbar_arrive(&acc_empty[as]);
                            }
                        }
                        const float4* b4 = reinterpret_cast<const float4*>(wv + h * 32);
                        const float4* c4 = reinterpret_cast<const float4*>(wv + 64 + h * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = b4[j];
                            f32x2_t y0 = f2_pack(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]));
                            f32x2_t y1 = f2_pack(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                            if constexpr (LNFOLD) {      // rstd * acc + (nrm * c + d)
                                const float4 cv = c4[j];
                                y0 = f2_fma(rstd2, y0, f2_fma(nrm2, f2_pack(cv.x, cv.y), f2_pack(b.x, b.y)));
                                y1 = f2_fma(rstd2, y1, f2_fma(nrm2, f2_pack(cv.z, cv.w), f2_pack(b.z, b.w)));
                            } else {
                                y0 = f2_add(y0, f2_pack(b.x, b.y));
                                y1 = f2_add(y1, f2_pack(b.z, b.w));
                            }
                            if constexpr (EPI == EPI_BIAS_GELU_FAST_BF16 || EPI == EPI_LNFOLD_GELU_BF16) {
                                pk[h * 16 + 2 * j] = gelu_fast2_bf16(y0);
                                pk[h * 16 + 2 * j + 1] = gelu_fast2_bf16(y1);
                            } else if constexpr (EPI == EPI_LNFOLD_GELU2_BF16) {
                                pk[h * 16 + 2 * j] = gelu_fast2x2_bf16(y0);
                                pk[h * 16 + 2 * j + 1] = gelu_fast2x2_bf16(y1);
                            } else {
                                float f0, f1, f2, f3;
                                f2_unpack(y0, f0, f1);
                                f2_unpack(y1, f2, f3);
                                if constexpr (EPI == EPI_BIAS_GELU_BF16) {
                                    f0 = gelu_erf(f0); f1 = gelu_erf(f1); f2 = gelu_erf(f2); f3 = gelu_erf(f3);
                                }
                                pk[h * 16 + 2 * j] = pack_bf16x2(f0, f1);
                                pk[h * 16 + 2 * j + 1] = pack_bf16x2(f2, f3);
                            }
                        }
                    }
                    if (lane == 0 && (warp == 4 || warp == 12)) TRACE(1 + (warp == 12), k, 3 + 3 * cc);
                    // the previous TMA store of this warp must have drained the staging tile (it has had the whole
                    // math phase above to do so)
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        sts_u4(row_addr + ((q ^ sw) << 4), make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]));
                    fence_proxy_async_smem();
                    __syncwarp();
#ifndef HB_EXP_NOSTORE
                    if (lane == 0) {
                        tma_store_2d(&map_out, stage_buf, n0 + c * 64, m0);
                        tma_store_commit();
                    }
#endif
                    if (lane == 0 && (warp == 4 || warp == 12)) TRACE(1 + (warp == 12), k, 4 + 3 * cc);
                }
            }
            if (lane == 0) tma_store_wait_all<0>();
        } else {
        // ------------------------------------------------------------------ epilogue, fp32 outputs (4 warpgroups)
        // Each warpgroup covers all 128 accumulator rows (warp & 3 = TMEM lane quadrant, thread = row) and takes the
        // 128-byte-wide column chunks c with c = wg + rot (mod 4); rot advances per tile so that chunk counts that are
        // not a multiple of 4 balance over tiles.
        const int wg = (warp - 4) >> 2;
        const int ew = warp & 3;
        const int t_wg = threadIdx.x & 127;                   // thread index within the warpgroup
        const int row_in_tile = ew * 32 + lane;
        const bool store_leader = ((warp & 3) == 0 && lane == 0);
        float* wg_vec = vec_x + wg * 128;
        const int sw = row_in_tile & 7;
        uint32_t it = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
            const int m0 = (tile / n_tiles) * TILE_M + cta_rank * GEMM_BM;
            const int n_idx = tile % n_tiles;
            const int n0 = n_idx * BN;
            const int row = m0 + row_in_tile;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            // chunk -> warpgroup assignment rotates with the tile count to balance NCHUNK % NWG != 0, except where the
            // warpgroups' partial row statistics are summed afterwards: a rotating grouping would make a row's (sum, sum of
            // squares) depend on WHICH tile of its CTA it was, i.e. on the rows sharing the launch (fp32 is not associative)
            const int c_first = (RESID || OUT_TOKENS) ? wg : ((wg + it * NCHUNK) & (NWG - 1));
            [[maybe_unused]] float st_sum = 0.f, st_sq = 0.f;   // row statistics of what this thread produced in the tile
            const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;

            {
                mbar_wait(&acc_full[as], aphase);
                tc_fence_after();
#ifdef HB_EXP_NOEPI
                tc_fence_before();
                if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&acc_empty[as], 0); else mbar_arrive(&acc_empty[as]);
                continue;
#endif
                if constexpr (OUT_TOKENS) {
                    // fp32 token rows written straight to global with a per-sequence row remap (+1 for the CLS slot)
                    // and a per-(token,col) additive table (positional embedding); optional bf16 copy + row statistics.
                    const int seq = row / tokens_per_seq;
                    const int tok = row - seq * tokens_per_seq;
                    const size_t out_row = static_cast<size_t>(seq) * (tokens_per_seq + 1) + 1 + tok;
#pragma unroll 1
                    for (int c = c_first; c < NCHUNK; c += NWG) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr_row + c * 32, v);
                        tmem_ld_wait();
                        if (row < M) {
                            const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c * 32);
                            const float4* t4 = reinterpret_cast<const float4*>(tok_table + static_cast<size_t>(1 + tok) * N + n0 + c * 32);
                            float4* o4 = reinterpret_cast<float4*>(tok_out + out_row * N + n0 + c * 32);
                            uint4* ob = aux.xb_out ? reinterpret_cast<uint4*>(aux.xb_out + out_row * N + n0 + c * 32) : nullptr;
#pragma unroll
                            for (int j = 0; j < 8; j += 2) {
                                float4 r[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const float4 b = __ldg(b4 + j + u);
                                    const float4 t = __ldg(t4 + j + u);
                                    r[u].x = __uint_as_float(v[4 * (j + u) + 0]) + b.x;
                                    r[u].y = __uint_as_float(v[4 * (j + u) + 1]) + b.y;
                                    r[u].z = __uint_as_float(v[4 * (j + u) + 2]) + b.z;
                                    r[u].w = __uint_as_float(v[4 * (j + u) + 3]) + b.w;
                                    if constexpr (EPI == EPI_TOKENS_GELU_F32) {
                                        r[u].x = gelu_erf(r[u].x); r[u].y = gelu_erf(r[u].y);
                                        r[u].z = gelu_erf(r[u].z); r[u].w = gelu_erf(r[u].w);
                                    }
                                    r[u].x += t.x; r[u].y += t.y; r[u].z += t.z; r[u].w += t.w;
                                    if (tok_out) o4[j + u] = r[u];
                                    st_sum += (r[u].x + r[u].y) + (r[u].z + r[u].w);
                                    st_sq = fmaf(r[u].x, r[u].x, fmaf(r[u].y, r[u].y, fmaf(r[u].z, r[u].z, fmaf(r[u].w, r[u].w, st_sq))));
                                }
                                if (ob) ob[j >> 1] = make_uint4(pack_bf16x2(r[0].x, r[0].y), pack_bf16x2(r[0].z, r[0].w),
                                                                pack_bf16x2(r[1].x, r[1].y), pack_bf16x2(r[1].z, r[1].w));
                            }
                        }
                    }
                } else if constexpr (RESID) {
                    uint8_t* xb_buf = smem_o + RESID_RING * STAGE_BYTES_OUT + wg * 8192;   // [128 rows x 64 B], SWIZZLE_64B
                    uint8_t* xb_row = xb_buf + row_in_tile * 64;
                    const int sw64 = (row_in_tile >> 1) & 3;
#pragma unroll 1
                    for (int c = c_first; c < NCHUNK; c += NWG) {
                        const uint32_t q = it * NCHUNK + c;                      // chunk sequence number of this CTA
                        const uint32_t slot = q % RESID_RING, use = q / RESID_RING;
                        uint8_t* rbuf = smem_o + slot * STAGE_BYTES_OUT;
                        uint8_t* rrow = rbuf + row_in_tile * 128;
                        uint32_t v[32];
                        tmem_ld_32x32(taddr_row + c * 32, v);
                        if (t_wg < 32) wg_vec[t_wg] = __ldg(bias + n0 + c * 32 + t_wg);
                        mbar_wait(&ring_full[slot], use & 1);
                        // bias slice visible; the leader has drained the stores that read the bf16 staging tile
                        named_bar_sync(1 + wg, GEMM_EPI_THREADS);
                        tmem_ld_wait();
                        const float4* b4 = reinterpret_cast<const float4*>(wg_vec);
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bb = b4[j];
                            float4* p = reinterpret_cast<float4*>(rrow + ((j ^ sw) << 4));
                            float4 r = *p;
                            r.x += __uint_as_float(v[4 * j + 0]) + bb.x; r.y += __uint_as_float(v[4 * j + 1]) + bb.y;
                            r.z += __uint_as_float(v[4 * j + 2]) + bb.z; r.w += __uint_as_float(v[4 * j + 3]) + bb.w;
                            *p = r;
                            st_sum += (r.x + r.y) + (r.z + r.w);
                            st_sq = fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, fmaf(r.w, r.w, st_sq))));
                            pk[2 * j] = pack_bf16x2(r.x, r.y);
                            pk[2 * j + 1] = pack_bf16x2(r.z, r.w);
                        }
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq)
                            *reinterpret_cast<uint4*>(xb_row + ((qq ^ sw64) << 4)) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + wg, GEMM_EPI_THREADS);
                        if (store_leader) {
                            tma_store_2d(&map_out, rbuf, n0 + c * 32, m0);
                            tma_store_2d(&map_xb, xb_buf, n0 + c * 32, m0);
                            tma_store_commit();
                            tma_store_wait_read<0>();                            // both tiles read: hand the slot back
                            mbar_arrive(&ring_empty[slot]);
                        }
                    }
                } else {
                    // ---- EPI_BIAS_RESADD_F32: fp32 chunks added into the output by TMA reduce
                    uint8_t* stage_buf = smem_o + wg * STAGE_BYTES_OUT;
                    uint8_t* row_ptr = stage_buf + row_in_tile * 128;
#pragma unroll 1
                    for (int c = c_first; c < NCHUNK; c += NWG) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr_row + c * 32, v);
                        tmem_ld_wait();
                        const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c * 32);
                        float4 r[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(b4 + j);
                            r[j].x = __uint_as_float(v[4 * j + 0]) + b.x; r[j].y = __uint_as_float(v[4 * j + 1]) + b.y;
                            r[j].z = __uint_as_float(v[4 * j + 2]) + b.z; r[j].w = __uint_as_float(v[4 * j + 3]) + b.w;
                        }
                        if (store_leader) tma_store_wait_read<0>();
                        named_bar_sync(1 + wg, GEMM_EPI_THREADS);
#pragma unroll
                        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(row_ptr + ((j ^ sw) << 4)) = r[j];
                        fence_proxy_async_smem();
                        named_bar_sync(1 + wg, GEMM_EPI_THREADS);
                        if (store_leader) {
                            tma_reduce_add_2d(&map_out, stage_buf, n0 + c * 32, m0);
                            tma_store_commit();
                        }
                    }
                }
                // this thread has issued (and waited for) its last TMEM load of the tile: hand the buffer back
                tc_fence_before();
                if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&acc_empty[as], 0);
                else mbar_arrive(&acc_empty[as]);
            }

            if constexpr (RESID || OUT_TOKENS) {
                // row statistics of the produced rows: warpgroups 1-3 hand their partials to warpgroup 0 through shared
                // memory so that exactly one atomicAdd per (row, n-tile) reaches the global array (order-independent)
                if (aux.stats_out != nullptr) {
                    float* sx = stat_x + (it & 1) * (NWG * 256) + row_in_tile * 2;
                    if (wg != 0) { sx[wg * 256] = st_sum; sx[wg * 256 + 1] = st_sq; }
                    named_bar_sync(BAR_EPI_ALL, NWG * GEMM_EPI_THREADS);
                    if (wg == 0 && row < M) {
                        size_t srow = row;
                        if constexpr (OUT_TOKENS) {
                            const int seq = row / tokens_per_seq;
                            srow = static_cast<size_t>(seq) * (tokens_per_seq + 1) + 1 + (row - seq * tokens_per_seq);
                        }
#pragma unroll
                        for (int g = 1; g < NWG; ++g) { st_sum += sx[g * 256]; st_sq += sx[g * 256 + 1]; }
                        atomicAdd(aux.stats_out + srow * 2, st_sum);
                        atomicAdd(aux.stats_out + srow * 2 + 1, st_sq);
                        if (RESID && n_idx == 0 && aux.stats_clear != nullptr)
                            *reinterpret_cast<float2*>(aux.stats_clear + static_cast<size_t>(row) * 2) = make_float2(0.f, 0.f);
                    }
                }
            }
        }
        if (store_leader) tma_store_wait_all<0>();
        }
    }

    tc_fence_before();
    if constexpr (CG == 2) {
        cluster_sync_all();                                    // the peer's remote arrivals have landed
        if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    } else {
        __syncthreads();
        if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
template <int BN, int EPI, int CG, int AST>
static int launch_gemm_cg(const GemmArgs& g, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, CG, EPI, AST>;
    auto kern = gemm_bf16_kernel<BN, EPI, CG, AST>;
    if (set_max_dynamic_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES)) return -1;
    const int m_tiles = (g.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
    const int n_tiles = g.N / BN;
    const int slots = num_sms() / CG;
    // A-stationary grouping: all n-tiles of an m-tile share one A load when there are enough m-tiles to fill the SM pairs
    // evenly; otherwise half of them (twice the units, half the reuse)
    int n_group = 1;
    if (AST) {
        n_group = n_tiles;
        const int waves_full = (m_tiles + slots - 1) / slots;
        if (n_tiles % 2 == 0 && (m_tiles < slots || static_cast<double>(m_tiles) / slots < 0.9 * waves_full)) n_group = n_tiles / 2;
    }
    const int total = AST ? m_tiles * (n_tiles / n_group) : m_tiles * n_tiles;
    const int grid = (total < slots ? total : slots) * CG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (CG == 2) ? 1 : 0;
    HB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, g.map_a, CG == 2 ? g.map_w2 : g.map_w, g.map_out, g.map_xb, g.bias,
                                  g.tok_table, g.tok_out, g.aux, g.M, g.N, g.K, g.tokens_per_seq, n_group));
    count_launch();
    return 0;
}

template <int BN, int EPI>
static int launch_gemm_t(const GemmArgs& g, cudaStream_t stream) {
    if constexpr ((EPI == EPI_LNFOLD_BF16 || EPI == EPI_RESID_BF16) && BN != 256) {
        if (g.cg2 && g.astat) return launch_gemm_cg<BN, EPI, 2, 1>(g, stream);
    }
    if (g.cg2) return launch_gemm_cg<BN, EPI, 2, 0>(g, stream);
    return launch_gemm_cg<BN, EPI, 1, 0>(g, stream);
}

// Instantiated variants: 192-wide tiles for every epilogue (all model widths — 192, 384, 576, 768, 1152 — are multiples of
// 192), 256-wide tiles for the plain / GELU / LayerNorm-folded epilogues of the N >= 1024 GEMMs (fc1).  A 128-wide tile was
// never picked by any model shape and is gone; the residual and token epilogues never ran 256 wide.
template <int BN>
static int launch_gemm_bn(const GemmArgs& g, cudaStream_t stream) {
    switch (g.epi) {
        case EPI_BIAS_BF16: return launch_gemm_t<BN, EPI_BIAS_BF16>(g, stream);
        case EPI_BIAS_GELU_BF16: return launch_gemm_t<BN, EPI_BIAS_GELU_BF16>(g, stream);
        case EPI_BIAS_GELU_FAST_BF16: return launch_gemm_t<BN, EPI_BIAS_GELU_FAST_BF16>(g, stream);
        case EPI_LNFOLD_BF16: return launch_gemm_t<BN, EPI_LNFOLD_BF16>(g, stream);
        case EPI_LNFOLD_GELU_BF16: return launch_gemm_t<BN, EPI_LNFOLD_GELU_BF16>(g, stream);
        case EPI_LNFOLD_GELU2_BF16: return launch_gemm_t<BN, EPI_LNFOLD_GELU2_BF16>(g, stream);
        default: break;
    }
    if constexpr (BN == 192) {
        switch (g.epi) {
            case EPI_BIAS_RESADD_F32: return launch_gemm_t<BN, EPI_BIAS_RESADD_F32>(g, stream);
            case EPI_TOKENS_F32: return launch_gemm_t<BN, EPI_TOKENS_F32>(g, stream);
            case EPI_TOKENS_GELU_F32: return launch_gemm_t<BN, EPI_TOKENS_GELU_F32>(g, stream);
            case EPI_RESID_BF16: return launch_gemm_t<BN, EPI_RESID_BF16>(g, stream);
            case EPI_RESID_STATS_F32: return launch_gemm_t<BN, EPI_RESID_STATS_F32>(g, stream);
            default: break;
        }
    }
    return set_error("hb_gemm: epilogue %d is not available with %d-wide tiles", g.epi, BN);
}

static bool gemm_use_cta_pairs() {      // HB_GEMM_CG=1 forces the single-CTA kernel (debug / comparison)
    static int v = -1;
    if (v < 0) { const char* e = getenv("HB_GEMM_CG"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

int gemm_pick_bn(int N) {
    {   // experiment hook: HB_GEMM_BN forces the tile width when it divides N
        const char* e = getenv("HB_GEMM_BN");
        if (e) { const int v = atoi(e); if ((v == 192 || v == 256) && N % v == 0) return v; }
    }
    if (N % 256 == 0 && N >= 1024) return 256;
    if (N % 192 == 0) return 192;
    return 0;
}

int gemm_prepare(GemmArgs& g, const void* A, const void* W, const float* bias, int epi, void* out, int M, int N, int K,
                 const float* tok_table, int tokens_per_seq, const GemmAux* aux, void* xb_out, size_t out_pitch_bytes) {
    if (M <= 0 || N <= 0 || K <= 0) return set_error("hb_gemm: bad shape M=%d N=%d K=%d", M, N, K);
    if (K % GEMM_BK != 0) return set_error("hb_gemm: K=%d must be a multiple of %d", K, GEMM_BK);
    int bn = gemm_pick_bn(N);
    const bool wide_ok = (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_GELU_FAST_BF16 ||
                          epi == EPI_LNFOLD_BF16 || epi == EPI_LNFOLD_GELU_BF16 || epi == EPI_LNFOLD_GELU2_BF16);
    if (bn == 256 && !wide_ok) bn = (N % 192 == 0) ? 192 : 0;   // residual / token epilogues run 192-wide tiles only
    if (bn == 0) return set_error("hb_gemm: N=%d must be a multiple of 192 (or of 256 from 1024 up for the plain, GELU and LayerNorm-folded epilogues)", N);
    if (bias == nullptr) return set_error("hb_gemm: bias is required");
    g.bn = bn; g.epi = epi; g.M = M; g.N = N; g.K = K; g.bias = bias;
    g.tok_table = tok_table; g.tok_out = nullptr; g.tokens_per_seq = tokens_per_seq;
    GemmAux zero = {};
    g.aux = aux ? *aux : zero;
    if (encode_tmap_2d(&g.map_a, TMAP_BF16, A, M, K, static_cast<uint64_t>(K) * 2, GEMM_BM, GEMM_BK)) return -1;
    if (encode_tmap_2d(&g.map_w, TMAP_BF16, W, N, K, static_cast<uint64_t>(K) * 2, bn, GEMM_BK)) return -1;
    if (encode_tmap_2d(&g.map_w2, TMAP_BF16, W, N, K, static_cast<uint64_t>(K) * 2, bn / 2, GEMM_BK)) return -1;
    g.cg2 = (M > GEMM_BM) && gemm_use_cta_pairs();
    {   // A-stationary variant for the K = 384 GEMMs of the block pipeline (HB_GEMM_ASTAT=0 disables it)
        static int use_astat = -1;
        if (use_astat < 0) { const char* e = getenv("HB_GEMM_ASTAT"); use_astat = (e && e[0] == '0') ? 0 : 1; }
        g.astat = (use_astat && g.cg2 && K == 384 && bn != 256 && (epi == EPI_LNFOLD_BF16 || epi == EPI_RESID_BF16)) ? 1 : 0;
    }
    g.map_xb = g.map_a;                                       // placeholder unless the epilogue uses it
    const bool lnfold = (epi == EPI_LNFOLD_BF16 || epi == EPI_LNFOLD_GELU_BF16 || epi == EPI_LNFOLD_GELU2_BF16);
    const bool out_bf16 = (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_GELU_FAST_BF16 || lnfold ||
                           epi == EPI_RESID_BF16);
    if (out_bf16) {
        if (lnfold && (!g.aux.colvec2 || !g.aux.row_stats))
            return set_error("hb_gemm: the LayerNorm-folded epilogue needs the column vector c and the row statistics");
        if (lnfold && (g.aux.n_part < 1 || g.aux.n_part > 6 || g.aux.stats_stride < M))
            return set_error("hb_gemm: the LayerNorm-folded epilogue needs 1..6 partial-sum planes of at least M rows");
        if (epi == EPI_RESID_BF16) {
            if (!xb_out || !g.aux.stats_out || g.aux.stats_stride < M)
                return set_error("hb_gemm: the bf16 residual epilogue needs the residual source and the statistics planes");
            const uint64_t rp = out_pitch_bytes ? out_pitch_bytes : static_cast<uint64_t>(N) * 2;
            if (encode_tmap_2d(&g.map_xb, TMAP_BF16, xb_out, M, N, rp, 32, 64)) return -1;
        }
        if (encode_tmap_2d(&g.map_out, TMAP_BF16, out, M, N, static_cast<uint64_t>(N) * 2, 32, 64)) return -1;   // one store per epilogue warp: 32 rows x 64 columns
    } else if (epi == EPI_BIAS_RESADD_F32 || epi == EPI_RESID_STATS_F32) {
        // out_pitch_bytes: the fp32 rows may be strided (the CLS rows of a [n_seq, seq_len, N] residual stream)
        const uint64_t pitch = out_pitch_bytes ? out_pitch_bytes : static_cast<uint64_t>(N) * 4;
        if (encode_tmap_2d(&g.map_out, TMAP_F32, out, M, N, pitch, GEMM_BM, 32)) return -1;
        if (epi == EPI_RESID_STATS_F32) {
            if (!xb_out) return set_error("hb_gemm: the residual epilogue needs the bf16 output");
            if (encode_tmap_2d(&g.map_xb, TMAP_BF16, xb_out, M, N, static_cast<uint64_t>(N) * 2, GEMM_BM, 32, 64)) return -1;
        }
    } else if (epi == EPI_TOKENS_F32 || epi == EPI_TOKENS_GELU_F32) {
        if (tok_table == nullptr || tokens_per_seq <= 0) return set_error("hb_gemm: token epilogue needs a table");
        g.map_out = g.map_a;   // unused by the kernel; keep the parameter well-formed
        g.tok_out = static_cast<float*>(out);
        g.aux.xb_out = static_cast<__nv_bfloat16*>(xb_out);
    } else {
        return set_error("hb_gemm: unknown epilogue %d", epi);
    }
    return 0;
}

#ifdef HB_EXP_TRACE
extern "C" int hb_exp_read_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 4 * 1024) == cudaSuccess ? 0 : -1;
}
#endif

int gemm_launch(const GemmArgs& g, cudaStream_t stream) {
    switch (g.bn) {
        case 192: return launch_gemm_bn<192>(g, stream);
        case 256: return launch_gemm_bn<256>(g, stream);
    }
    return set_error("hb_gemm: bad BN %d", g.bn);
}

}  // namespace hb
