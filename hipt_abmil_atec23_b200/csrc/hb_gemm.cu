// hb_gemm.cu — persistent, warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] * W[N,K]^T)
//
//   warp 0     : TMA producer  (A tile 128x64 and W tile per stage, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1     : MMA issuer    (one elected thread, tcgen05.mma kind::f16, K = 16 per instruction, accumulators in TMEM,
//                               double buffered against the epilogue)
//   warp 2     : TMEM allocator
//   warps 4-11 : epilogue      (two warpgroups; tcgen05.ld -> fused epilogue -> swizzled smem -> TMA store)
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

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;          // 64 bf16 = 128 B = one swizzle span
constexpr int GEMM_THREADS = 384;    // 12 warps: producer, mma, tmem-alloc, spare, 2 x 4 epilogue
constexpr int GEMM_EPI_THREADS = 128; // per epilogue warpgroup
constexpr int STAGE_BYTES_OUT = 128 * 128;   // 128 rows x 128 B staging chunk for the TMA store
constexpr int SMEM_LIMIT = 232448;   // 227 KB

// Epilogue kinds (mirrored in include/hipt_b200.h as HB_EPI_*)
enum : int { EPI_BIAS_BF16 = 0, EPI_BIAS_GELU_BF16 = 1, EPI_BIAS_RESADD_F32 = 2, EPI_TOKENS_F32 = 3, EPI_TOKENS_GELU_F32 = 4,
              EPI_BIAS_GELU_FAST_BF16 = 5, EPI_LNFOLD_BF16 = 6, EPI_LNFOLD_GELU_BF16 = 7, EPI_RESID_STATS_F32 = 8 };

template <int BN, int CG, int EPI>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_ROWS = BN / CG;                    // W rows staged by this CTA
    static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
    // per epilogue warpgroup: one 16 KB staging tile; the residual epilogue needs two fp32 residual tiles + 8 KB bf16
    static constexpr int OUT_BYTES_PER_WG = (EPI == EPI_RESID_STATS_F32) ? (2 * STAGE_BYTES_OUT + 8192) : STAGE_BYTES_OUT;
    static constexpr int TAIL_BYTES = 2048 /* row-stat exchange */ + 512 /* barriers */;
    static constexpr int RING_BUDGET = SMEM_LIMIT - 1024 - 2 * OUT_BYTES_PER_WG - TAIL_BYTES;
    static constexpr int STAGES_FIT = RING_BUDGET / (A_BYTES + B_BYTES);
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + 2 * OUT_BYTES_PER_WG + TAIL_BYTES;
    static_assert(STAGES >= 2, "operand ring too shallow");
};

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_xb,
                 const float* __restrict__ bias, const float* __restrict__ tok_table, float* __restrict__ tok_out,
                 const GemmAux aux, int M, int N, int K, int tokens_per_seq) {
    using Cfg = GemmCfg<BN, CG, EPI>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool OUT_BF16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_FAST_BF16 ||
                               EPI == EPI_LNFOLD_BF16 || EPI == EPI_LNFOLD_GELU_BF16);
    constexpr bool LNFOLD = (EPI == EPI_LNFOLD_BF16 || EPI == EPI_LNFOLD_GELU_BF16);
    constexpr bool OUT_TOKENS = (EPI == EPI_TOKENS_F32 || EPI == EPI_TOKENS_GELU_F32);
    constexpr bool RESID = (EPI == EPI_RESID_STATS_F32);
    constexpr int CHUNK_COLS = OUT_BF16 ? 64 : 32;      // 128 B of output per row per chunk

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint8_t* smem_o = smem_b + STAGES * Cfg::B_BYTES;
    float* stat_x = reinterpret_cast<float*>(smem_o + 2 * Cfg::OUT_BYTES_PER_WG);   // [2 parity][128][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stat_x) + 2048);
    uint64_t* full_bar = bars;                  // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]  MMA -> TMA
    uint64_t* acc_full = bars + 2 * STAGES;     // [2]       MMA -> epilogue
    uint64_t* acc_empty = bars + 2 * STAGES + 2;// [2]       epilogue -> MMA
    uint64_t* resid_full = bars + 2 * STAGES + 4;// [2 wg][2] residual tile landed (RESID only)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);

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
        if (RESID) tma_prefetch_desc(&map_xb);
    }
    if (warp == 1 && lane == 0) {
        // CG = 2: full_bar / acc_empty are used on the leader only and collect arrivals from both CTAs
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], CG * 2 * GEMM_EPI_THREADS); }
        for (int i = 0; i < 4; ++i) mbar_init(&resid_full[i], 1);
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

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint64_t pol_w = policy_evict_last();
            uint32_t stage = 0, phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int m0 = (tile / n_tiles) * TILE_M + cta_rank * GEMM_BM;
                const int n0 = (tile % n_tiles) * BN + cta_rank * Cfg::B_ROWS;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (CG == 2) {
                        // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the pair
                        tma_load_2d_2sm(smem_a + stage * Cfg::A_BYTES, &map_a, &full_bar[stage], kb * GEMM_BK, m0);
                        tma_load_2d_2sm_hint(smem_b + stage * Cfg::B_BYTES, &map_w, &full_bar[stage], kb * GEMM_BK, n0, pol_w);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
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
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
            uint32_t stage = 0, phase = 0;
            uint32_t it = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&acc_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k128(smem_u32(smem_a + stage * Cfg::A_BYTES));
                    const uint64_t db = umma_desc_k128(smem_u32(smem_b + stage * Cfg::B_BYTES));
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
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (2 warpgroups)
        // Each warpgroup covers all 128 accumulator rows (warp & 3 = TMEM lane quadrant) and takes every other
        // 128-byte-wide column chunk of the tile, with its own staging buffers, named barrier and TMA thread.
        const int wg = (warp - 4) >> 2;
        const int ew = warp & 3;
        const int row_in_tile = ew * 32 + lane;
        const bool store_leader = ((warp & 3) == 0 && lane == 0);
        uint8_t* wg_buf = smem_o + wg * Cfg::OUT_BYTES_PER_WG;
        const int sw = row_in_tile & 7;
        constexpr int NCHUNK = BN / CHUNK_COLS;
        uint32_t it = 0;
        uint32_t lc = 0;                                      // RESID: chunks processed by this warpgroup so far
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
            const int m0 = (tile / n_tiles) * TILE_M + cta_rank * GEMM_BM;
            const int n_idx = tile % n_tiles;
            const int n0 = n_idx * BN;
            const int row = m0 + row_in_tile;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int c_first = (wg + it * NCHUNK) & 1;       // alternate so odd chunk counts balance across tiles
            float st_sum = 0.f, st_sq = 0.f;                  // row statistics of what this thread produced in the tile

            // per-row LayerNorm factors for the folded epilogues
            float rstd = 0.f, nrm = 0.f;
            if constexpr (LNFOLD) {
                if (row < M) {
                    const float2 sq = __ldg(reinterpret_cast<const float2*>(aux.row_stats) + row);
                    const float mu = sq.x * aux.inv_dim;
                    const float var = fmaxf(sq.y * aux.inv_dim - mu * mu, 0.f);
                    rstd = rsqrtf(var + aux.eps);
                    nrm = -rstd * mu;
                }
            }
            if constexpr (RESID) {
                // fetch the residual tile of this warpgroup's first chunk while the MMAs are still running
                if (store_leader && c_first < NCHUNK) {
                    tma_store_wait_read<0>();
                    const uint32_t b = lc & 1;
                    mbar_arrive_expect_tx(&resid_full[wg * 2 + b], STAGE_BYTES_OUT);
                    tma_load_2d(wg_buf + b * STAGE_BYTES_OUT, &map_out, &resid_full[wg * 2 + b], n0 + c_first * 32, m0);
                }
            }

            mbar_wait(&acc_full[as], aphase);
            tc_fence_after();
            const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;

            if constexpr (OUT_TOKENS) {
                // fp32 token rows written straight to global with a per-sequence row remap (+1 for the CLS slot)
                // and a per-(token,col) additive table (positional embedding); optional bf16 copy + row statistics.
                const int seq = row / tokens_per_seq;
                const int tok = row - seq * tokens_per_seq;
                const size_t out_row = static_cast<size_t>(seq) * (tokens_per_seq + 1) + 1 + tok;
#pragma unroll 1
                for (int c = c_first; c < NCHUNK; c += 2) {
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
                                o4[j + u] = r[u];
                                st_sum += (r[u].x + r[u].y) + (r[u].z + r[u].w);
                                st_sq = fmaf(r[u].x, r[u].x, fmaf(r[u].y, r[u].y, fmaf(r[u].z, r[u].z, fmaf(r[u].w, r[u].w, st_sq))));
                            }
                            if (ob) ob[j >> 1] = make_uint4(pack_bf16x2(r[0].x, r[0].y), pack_bf16x2(r[0].z, r[0].w),
                                                            pack_bf16x2(r[1].x, r[1].y), pack_bf16x2(r[1].z, r[1].w));
                        }
                    }
                }
            } else if constexpr (RESID) {
                uint8_t* xb_buf = wg_buf + 2 * STAGE_BYTES_OUT;                  // [128 rows x 64 B], SWIZZLE_64B
                uint8_t* xb_row = xb_buf + row_in_tile * 64;
                const int sw64 = (row_in_tile >> 1) & 3;
#pragma unroll 1
                for (int c = c_first; c < NCHUNK; c += 2, ++lc) {
                    const uint32_t b = lc & 1;
                    uint8_t* rbuf = wg_buf + b * STAGE_BYTES_OUT;
                    uint8_t* rrow = rbuf + row_in_tile * 128;
                    uint32_t v[32];
                    tmem_ld_32x32(taddr_row + c * 32, v);
                    // every TMA store that read this warpgroup's buffers has drained; then prefetch the next chunk's
                    // residual tile into the other buffer
                    if (store_leader) {
                        tma_store_wait_read<0>();
                        if (c + 2 < NCHUNK) {
                            const uint32_t nb = (lc + 1) & 1;
                            mbar_arrive_expect_tx(&resid_full[wg * 2 + nb], STAGE_BYTES_OUT);
                            tma_load_2d(wg_buf + nb * STAGE_BYTES_OUT, &map_out, &resid_full[wg * 2 + nb], n0 + (c + 2) * 32, m0);
                        }
                    }
                    mbar_wait(&resid_full[wg * 2 + b], (lc >> 1) & 1);
                    tmem_ld_wait();
                    named_bar_sync(1 + wg, GEMM_EPI_THREADS);                    // bf16 staging tile is free for all
                    const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c * 32);
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = __ldg(b4 + j);
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
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4*>(xb_row + ((q ^ sw64) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    fence_proxy_async_smem();
                    named_bar_sync(1 + wg, GEMM_EPI_THREADS);
                    if (store_leader) {
                        tma_store_2d(&map_out, rbuf, n0 + c * 32, m0);
                        tma_store_2d(&map_xb, xb_buf, n0 + c * 32, m0);
                        tma_store_commit();
                    }
                }
            } else {
                uint8_t* stage_buf = wg_buf;
                uint8_t* row_ptr = stage_buf + row_in_tile * 128;
#pragma unroll 1
                for (int c = c_first; c < NCHUNK; c += 2) {
                    if constexpr (OUT_BF16) {
                        uint32_t pk[32];                 // 64 bf16 of this row
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t v[32];
                            tmem_ld_32x32(taddr_row + c * 64 + h * 32, v);
                            tmem_ld_wait();
                            const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c * 64 + h * 32);
                            const float4* c4 = LNFOLD ? reinterpret_cast<const float4*>(aux.colvec2 + n0 + c * 64 + h * 32) : b4;
                            [[maybe_unused]] const f32x2_t rstd2 = f2_pack(rstd, rstd), nrm2 = f2_pack(nrm, nrm);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 b = __ldg(b4 + j);
                                f32x2_t y0 = f2_pack(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]));
                                f32x2_t y1 = f2_pack(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                                if constexpr (LNFOLD) {      // rstd * acc + (nrm * c + d)
                                    const float4 cc = __ldg(c4 + j);
                                    y0 = f2_fma(rstd2, y0, f2_fma(nrm2, f2_pack(cc.x, cc.y), f2_pack(b.x, b.y)));
                                    y1 = f2_fma(rstd2, y1, f2_fma(nrm2, f2_pack(cc.z, cc.w), f2_pack(b.z, b.w)));
                                } else {
                                    y0 = f2_add(y0, f2_pack(b.x, b.y));
                                    y1 = f2_add(y1, f2_pack(b.z, b.w));
                                }
                                if constexpr (EPI == EPI_BIAS_GELU_FAST_BF16 || EPI == EPI_LNFOLD_GELU_BF16) {
                                    pk[h * 16 + 2 * j] = gelu_fast2_bf16(y0);
                                    pk[h * 16 + 2 * j + 1] = gelu_fast2_bf16(y1);
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
                        // the TMA store that last read this warpgroup's staging buffer must have drained it
                        if (store_leader) tma_store_wait_read<0>();
                        named_bar_sync(1 + wg, GEMM_EPI_THREADS);
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<uint4*>(row_ptr + ((q ^ sw) << 4)) =
                                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    } else {
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
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1 + wg, GEMM_EPI_THREADS);
                    if (store_leader) {
                        if constexpr (EPI == EPI_BIAS_RESADD_F32)
                            tma_reduce_add_2d(&map_out, stage_buf, n0 + c * CHUNK_COLS, m0);
                        else
                            tma_store_2d(&map_out, stage_buf, n0 + c * CHUNK_COLS, m0);
                        tma_store_commit();
                    }
                }
            }
            // this thread has issued (and waited for) its last TMEM load of the tile: hand the buffer back
            tc_fence_before();
            if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&acc_empty[as], 0);
            else mbar_arrive(&acc_empty[as]);

            if constexpr (RESID || OUT_TOKENS) {
                // row statistics of the produced rows: warpgroup 1 hands its partial to warpgroup 0 through shared
                // memory so that exactly one atomicAdd per (row, n-tile) reaches the global array (order-independent)
                if (aux.stats_out != nullptr) {
                    float* sx = stat_x + (it & 1) * 256 + row_in_tile * 2;
                    if (wg == 1) { sx[0] = st_sum; sx[1] = st_sq; }
                    named_bar_sync(3, 2 * GEMM_EPI_THREADS);
                    if (wg == 0 && row < M) {
                        size_t srow = row;
                        if constexpr (OUT_TOKENS) {
                            const int seq = row / tokens_per_seq;
                            srow = static_cast<size_t>(seq) * (tokens_per_seq + 1) + 1 + (row - seq * tokens_per_seq);
                        }
                        atomicAdd(aux.stats_out + srow * 2, st_sum + sx[0]);
                        atomicAdd(aux.stats_out + srow * 2 + 1, st_sq + sx[1]);
                        if (RESID && n_idx == 0 && aux.stats_clear != nullptr)
                            *reinterpret_cast<float2*>(aux.stats_clear + static_cast<size_t>(row) * 2) = make_float2(0.f, 0.f);
                    }
                }
            }
        }
        if (store_leader) tma_store_wait_all<0>();
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
template <int BN, int EPI, int CG>
static int launch_gemm_cg(const GemmArgs& g, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, CG, EPI>;
    auto kern = gemm_bf16_kernel<BN, EPI, CG>;
    static bool attr_done = false;     // per instantiation
    if (!attr_done) {
        HB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_done = true;
    }
    const int m_tiles = (g.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
    const int total = m_tiles * (g.N / BN);
    const int slots = num_sms() / CG;
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
                                  g.tok_table, g.tok_out, g.aux, g.M, g.N, g.K, g.tokens_per_seq));
    count_launch();
    return 0;
}

template <int BN, int EPI>
static int launch_gemm_t(const GemmArgs& g, cudaStream_t stream) {
    if (g.cg2) return launch_gemm_cg<BN, EPI, 2>(g, stream);
    return launch_gemm_cg<BN, EPI, 1>(g, stream);
}

template <int BN>
static int launch_gemm_bn(const GemmArgs& g, cudaStream_t stream) {
    switch (g.epi) {
        case EPI_BIAS_BF16: return launch_gemm_t<BN, EPI_BIAS_BF16>(g, stream);
        case EPI_BIAS_GELU_BF16: return launch_gemm_t<BN, EPI_BIAS_GELU_BF16>(g, stream);
        case EPI_BIAS_RESADD_F32: return launch_gemm_t<BN, EPI_BIAS_RESADD_F32>(g, stream);
        case EPI_TOKENS_F32: return launch_gemm_t<BN, EPI_TOKENS_F32>(g, stream);
        case EPI_TOKENS_GELU_F32: return launch_gemm_t<BN, EPI_TOKENS_GELU_F32>(g, stream);
        case EPI_BIAS_GELU_FAST_BF16: return launch_gemm_t<BN, EPI_BIAS_GELU_FAST_BF16>(g, stream);
        case EPI_LNFOLD_BF16: return launch_gemm_t<BN, EPI_LNFOLD_BF16>(g, stream);
        case EPI_LNFOLD_GELU_BF16: return launch_gemm_t<BN, EPI_LNFOLD_GELU_BF16>(g, stream);
        case EPI_RESID_STATS_F32: return launch_gemm_t<BN, EPI_RESID_STATS_F32>(g, stream);
    }
    return set_error("hb_gemm: unknown epilogue %d", g.epi);
}

static bool gemm_use_cta_pairs() {      // HB_GEMM_CG=1 forces the single-CTA kernel (debug / comparison)
    static int v = -1;
    if (v < 0) { const char* e = getenv("HB_GEMM_CG"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

int gemm_pick_bn(int N) {
    if (N % 256 == 0 && N >= 1024) return 256;
    if (N % 192 == 0) return 192;
    if (N % 128 == 0) return 128;
    return 0;
}

int gemm_prepare(GemmArgs& g, const void* A, const void* W, const float* bias, int epi, void* out, int M, int N, int K,
                 const float* tok_table, int tokens_per_seq, const GemmAux* aux, void* xb_out, size_t out_pitch_bytes) {
    if (M <= 0 || N <= 0 || K <= 0) return set_error("hb_gemm: bad shape M=%d N=%d K=%d", M, N, K);
    if (K % GEMM_BK != 0) return set_error("hb_gemm: K=%d must be a multiple of %d", K, GEMM_BK);
    const int bn = gemm_pick_bn(N);
    if (bn == 0) return set_error("hb_gemm: N=%d must be a multiple of 128 or 192", N);
    if (bias == nullptr) return set_error("hb_gemm: bias is required");
    g.bn = bn; g.epi = epi; g.M = M; g.N = N; g.K = K; g.bias = bias;
    g.tok_table = tok_table; g.tok_out = nullptr; g.tokens_per_seq = tokens_per_seq;
    GemmAux zero = {};
    g.aux = aux ? *aux : zero;
    if (encode_tmap_2d(&g.map_a, TMAP_BF16, A, M, K, static_cast<uint64_t>(K) * 2, GEMM_BM, GEMM_BK)) return -1;
    if (encode_tmap_2d(&g.map_w, TMAP_BF16, W, N, K, static_cast<uint64_t>(K) * 2, bn, GEMM_BK)) return -1;
    if (encode_tmap_2d(&g.map_w2, TMAP_BF16, W, N, K, static_cast<uint64_t>(K) * 2, bn / 2, GEMM_BK)) return -1;
    g.cg2 = (M > GEMM_BM) && gemm_use_cta_pairs();
    g.map_xb = g.map_a;                                       // placeholder unless the epilogue uses it
    const bool out_bf16 = (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_GELU_FAST_BF16 ||
                           epi == EPI_LNFOLD_BF16 || epi == EPI_LNFOLD_GELU_BF16);
    if (out_bf16) {
        if ((epi == EPI_LNFOLD_BF16 || epi == EPI_LNFOLD_GELU_BF16) && (!g.aux.colvec2 || !g.aux.row_stats))
            return set_error("hb_gemm: the LayerNorm-folded epilogue needs the column vector c and the row statistics");
        if (encode_tmap_2d(&g.map_out, TMAP_BF16, out, M, N, static_cast<uint64_t>(N) * 2, GEMM_BM, 64)) return -1;
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

int gemm_launch(const GemmArgs& g, cudaStream_t stream) {
    switch (g.bn) {
        case 128: return launch_gemm_bn<128>(g, stream);
        case 192: return launch_gemm_bn<192>(g, stream);
        case 256: return launch_gemm_bn<256>(g, stream);
    }
    return set_error("hb_gemm: bad BN %d", g.bn);
}

}  // namespace hb
