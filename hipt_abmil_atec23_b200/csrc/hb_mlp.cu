// hb_mlp.cu — the whole MLP half of a ViT-256 block as ONE kernel (Block.forward, HIPT_4K/vision_transformer.py:151:
// x = x + drop_path(mlp(norm2(x))) with Mlp.forward :98-104 = fc2(GELU(fc1(.)))), dim 384, hidden 1536:
//
//     xb = bf16( xb + (2 gelu( LN2-folded( xb W1g^T ) )) (W2/2)^T + b2 )        + partial row statistics for norm1
//
// The [rows, 1536] hidden activation never leaves the SM: per 256-row tile of a CTA pair (tcgen05 cta_group::2, each CTA
// owns 128 rows) the bf16 residual tile A = xb[128 x 384] stays resident in shared memory, the hidden dimension is walked
// in 24 chunks of 64 columns, and per chunk
//     S_j  = A W1g[64j..64j+63]^T                 24 MMAs (M256 x N64 x K16) into one of two 64-column TMEM buffers
//     H_j  = bf16(2 gelu(rstd S_j - rstd mu c + d))  epilogue warps: TMEM -> registers -> swizzled shared memory
//     O   += H_j (W2/2)[:, 64j..64j+63]^T           8 MMAs (M256 x N192 x K16, two column halves) into 384 TMEM columns
// so fc1's epilogue is hidden behind twice the MMA work it has in a stand-alone GEMM, and the only HBM traffic is the
// residual tile in and out (bf16) plus the weights (L2-resident).  After the last chunk the tile's epilogue adds the
// residual (read back from the resident A tile), rounds once, leaves the per-64-column (sum, sum of squares) planes the
// LayerNorm folded into the next qkv GEMM needs, and stores the tile from the same shared memory by TMA.
//
//   warp 0      TMA producer: A tile per tile, W1 chunk (this CTA's 32 of the 64 rows) and W2 chunk (its 2 x 96 of the
//               384 rows) through two-stage rings
//   warp 1      MMA issuer (leader CTA of the pair; warp-uniform, one elected lane):  S_j ; O += H_(j-2)  software-pipelined
//   warp 2      TMEM allocator (512 columns: O 0..383, S buffers 384..447 / 448..511)
//   warps 4-19  epilogue: warpgroup g = chunk j mod 4 turns S_j into H_j (thread = row); all 16 warps run the tile epilogue
#include <stdlib.h>

#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

#ifdef HB_EXP_TRACE
__device__ long long g_mlp_trace[8192];
#define MTRACE(idx) do { if (blockIdx.x == 0 && (idx) < 8192) g_mlp_trace[(idx)] = clock64(); } while (0)
extern "C" int hb_exp_read_mlp_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_mlp_trace, sizeof(long long) * 8192) == cudaSuccess ? 0 : -1;
}
#else
#define MTRACE(idx) do { } while (0)
#endif

constexpr int MLP_D = 384;
constexpr int MLP_H = 1536;
constexpr int MLP_CH = 64;                       // hidden columns per chunk
constexpr int MLP_NCH = MLP_H / MLP_CH;          // 24
constexpr int MLP_KB = MLP_D / 64;               // 6 k-blocks of the resident A tile
constexpr int MLP_THREADS = 640;
constexpr int MLP_A_BYTES = MLP_KB * 16384;      // 98304
constexpr int MLP_W1_STAGE = MLP_KB * 4096;      // 24576: [6 k-blocks][32 rows][128 B]
constexpr int MLP_W2_STAGE = 2 * 12288;          // 24576: [2 column halves][96 rows][128 B]
constexpr int MLP_H_BYTES = 16384;               // [128 rows][128 B]
constexpr int MLP_OFF_W1 = MLP_A_BYTES;
constexpr int MLP_OFF_W2 = MLP_OFF_W1 + 2 * MLP_W1_STAGE;
constexpr int MLP_OFF_H = MLP_OFF_W2 + 2 * MLP_W2_STAGE;
constexpr int MLP_OFF_VEC = MLP_OFF_H + 2 * MLP_H_BYTES;      // [4 warpgroups][c 64 | d 64] floats
constexpr int MLP_OFF_BAR = MLP_OFF_VEC + 4 * 512;
constexpr int MLP_SMEM = MLP_OFF_BAR + 256;
static_assert(MLP_SMEM <= 232448, "fused MLP shared memory budget");
constexpr uint32_t MLP_TMEM_S = 384;             // first S column

struct MlpArgs {
    CUtensorMap map_a, map_w1, map_w2, map_out;
    const float* c1;            // [1536] row sums of the rounded gamma-folded fc1 weight
    const float* d1;            // [1536] W1 beta + b1
    const float* b2;            // [384]
    const float* stats_in;      // [6][stride][2] partial (sum, sum of squares) of the rows before norm2
    float* stats_out;           // [6][stride][2] the same of the produced rows (for norm1 of the next block)
    int stats_stride;
    float eps;
    int M;
};

__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_out,
                 const float* __restrict__ c1, const float* __restrict__ d1, const float* __restrict__ b2,
                 const float* __restrict__ stats_in, float* __restrict__ stats_out, int stats_stride, float eps, int M,
                 int reverse) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* sA = smem;
    uint8_t* sW1 = smem + MLP_OFF_W1;
    uint8_t* sW2 = smem + MLP_OFF_W2;
    uint8_t* sH = smem + MLP_OFF_H;
    float* vec = reinterpret_cast<float*>(smem + MLP_OFF_VEC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MLP_OFF_BAR);
    uint64_t* a_full = bars;            // leader: both CTAs' A tiles landed
    uint64_t* a_empty = bars + 1;       // local : the 16 epilogue warps are done with the A tile (and their stores drained)
    uint64_t* w1_full = bars + 2;       // [2] leader
    uint64_t* w1_empty = bars + 4;      // [2] local (MMA commit, multicast)
    uint64_t* w2_full = bars + 6;       // [2] leader
    uint64_t* w2_empty = bars + 8;      // [2] local
    uint64_t* s_full = bars + 10;       // [4] local (MMA commit): S_j is in TMEM; one barrier per consuming warpgroup (j mod 4)
    uint64_t* s_empty = bars + 14;      // [2] leader: 4 warps x 2 CTAs have read S_j
    uint64_t* h_full = bars + 16;       // [2] leader: 4 warps x 2 CTAs have written H_j
    uint64_t* h_empty = bars + 18;      // [2] local (MMA commit): the MMAs reading H_j have retired
    uint64_t* o_full = bars + 20;       // local (MMA commit): the tile's accumulator is complete
    uint64_t* o_empty = bars + 21;      // leader: 16 warps x 2 CTAs have read it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_tiles = (M + 255) / 256;
    cluster_sync_all();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2); tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 2); mbar_init(a_empty, 16);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&w1_full[i], 2); mbar_init(&w1_empty[i], 1);
            mbar_init(&w2_full[i], 2); mbar_init(&w2_empty[i], 1);
            mbar_init(&s_empty[i], 8);
            mbar_init(&h_full[i], 8);  mbar_init(&h_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], 1);
        mbar_init(o_full, 1); mbar_init(o_empty, 32);
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto arrive_leader = [&](uint64_t* bar) {      // barriers that gate the MMA issuer live in the leader CTA
        if (cta_rank != 0) mbar_arrive_remote(bar, 0); else mbar_arrive(bar);
    };

    if (warp == 0) {
        // ---------------------------------------------------------------------------------------------- TMA producer
        if (lane == 0) {
            const uint64_t pol_w = policy_evict_last();
            uint32_t n = 0;                                           // global chunk counter of this CTA pair
            uint32_t ti = 0;
            auto load_w2 = [&](uint32_t nn, int jj) {
                const uint32_t s = nn & 1, u = nn >> 1;
                mbar_wait(&w2_empty[s], (u & 1) ^ 1);
                uint8_t* dst = sW2 + s * MLP_W2_STAGE;
                tma_load_2d_2sm_hint(dst, &map_w2, &w2_full[s], jj * MLP_CH, cta_rank * 96, pol_w);
                tma_load_2d_2sm_hint(dst + 12288, &map_w2, &w2_full[s], jj * MLP_CH, 192 + cta_rank * 96, pol_w);
                if (cta_rank == 0) mbar_arrive_expect_tx(&w2_full[s], 2 * MLP_W2_STAGE);
                else mbar_arrive_remote(&w2_full[s], 0);
            };
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
                const int m0 = (reverse ? n_tiles - 1 - tile : tile) * 256 + cta_rank * 128;
                mbar_wait(a_empty, (ti & 1) ^ 1);
#pragma unroll
                for (int kb = 0; kb < MLP_KB; ++kb) tma_load_2d_2sm(sA + kb * 16384, &map_a, a_full, kb * 64, m0);
                if (cta_rank == 0) mbar_arrive_expect_tx(a_full, 2 * MLP_A_BYTES);
                else mbar_arrive_remote(a_full, 0);
                for (int j = 0; j < MLP_NCH + 2; ++j) {
                    if (j < MLP_NCH) {                                // W1 chunk j
                        const uint32_t nn = n + j, s = nn & 1, u = nn >> 1;
                        mbar_wait(&w1_empty[s], (u & 1) ^ 1);
                        uint8_t* dst = sW1 + s * MLP_W1_STAGE;
#pragma unroll
                        for (int kb = 0; kb < MLP_KB; ++kb)
                            tma_load_2d_2sm_hint(dst + kb * 4096, &map_w1, &w1_full[s], kb * 64, j * MLP_CH + cta_rank * 32, pol_w);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&w1_full[s], 2 * MLP_W1_STAGE);
                        else mbar_arrive_remote(&w1_full[s], 0);
                    }
                    if (j >= 2) load_w2(n + j - 2, j - 2);           // W2 chunk j-2: consumed two chunks behind S_j
                }
                n += MLP_NCH;
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------------------------------------- MMA issuer: S_j
        // Two issuing warps (this one and warp 3) feed the tensor pipe independently, so the S stream runs ahead while
        // the O stream waits for the GELU of its chunk.  Warp-uniform control flow, one elected lane issues.
        if (cta_rank == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(256, MLP_CH);
            const uint32_t a_base = smem_u32(sA), w1_base = smem_u32(sW1);
            uint32_t n = 0, ti = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
                mbar_wait(a_full, ti & 1);
                for (int j = 0; j < MLP_NCH; ++j) {
                    const uint32_t nn = n + j, s = nn & 1, u = nn >> 1;
                    if (lane == 0) MTRACE(ti * 128 + j * 4 + 0);
                    mbar_wait(&w1_full[s], u & 1);
                    mbar_wait(&s_empty[s], (u & 1) ^ 1);
                    if (lane == 0) MTRACE(ti * 128 + j * 4 + 1);
                    tc_fence_after();
                    const uint32_t d_s = tmem_base + MLP_TMEM_S + s * MLP_CH;
                    const uint32_t w1s = w1_base + s * MLP_W1_STAGE;
                    if (elect_one()) {
#pragma unroll
                        for (int kb = 0; kb < MLP_KB; ++kb) {
                            const uint64_t da = umma_desc_k128(a_base + kb * 16384);
                            const uint64_t db = umma_desc_k128(w1s + kb * 4096);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss_2sm(d_s, da + 2 * k, db + 2 * k, idesc_s, (kb | k) != 0);
                        }
                        umma_commit_2sm(&w1_empty[s]);
                        umma_commit_2sm(&s_full[j & 3]);
                    }
                    __syncwarp();
                }
                n += MLP_NCH;
            }
        }
    } else if (warp == 3) {
        // ---------------------------------------------------------------------------------------------- MMA issuer: O += H_j W2_j^T
        if (cta_rank == 0) {
            constexpr uint32_t idesc_o = umma_idesc_bf16(256, 192);
            const uint32_t w2_base = smem_u32(sW2), h_base = smem_u32(sH);
            uint32_t n = 0, ti = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
                mbar_wait(o_empty, (ti & 1) ^ 1);                    // previous tile's accumulator has been read
                for (int jj = 0; jj < MLP_NCH; ++jj) {
                    const uint32_t nn = n + jj, s = nn & 1, u = nn >> 1;
                    if (lane == 0) MTRACE(ti * 128 + jj * 4 + 2);
                    mbar_wait(&w2_full[s], u & 1);
                    mbar_wait(&h_full[s], u & 1);
                    if (lane == 0) MTRACE(ti * 128 + jj * 4 + 3);
                    tc_fence_after();
                    const uint32_t hs = h_base + s * MLP_H_BYTES;
                    const uint32_t w2s = w2_base + s * MLP_W2_STAGE;
                    if (elect_one()) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint64_t da = umma_desc_k128(hs);
                            const uint64_t db = umma_desc_k128(w2s + q * 12288);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss_2sm(tmem_base + q * 192, da + 2 * k, db + 2 * k, idesc_o, (jj | k) != 0);
                        }
                        umma_commit_2sm(&h_empty[s]);
                        umma_commit_2sm(&w2_empty[s]);
                        if (jj == MLP_NCH - 1) umma_commit_2sm(o_full);
                    }
                    __syncwarp();
                }
                n += MLP_NCH;
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------------------------------------- epilogue warps
        const int w16 = warp - 4;
        const int g = w16 >> 2;                                  // warpgroup: chunk j mod 4 / tile-epilogue column slot
        const int ew = warp & 3;                                 // TMEM lane quadrant
        const int t_wg = threadIdx.x & 127;
        const int row_in_tile = ew * 32 + lane;
        const int sw = lane & 7;
        float* wg_vec = vec + g * 128;
        const uint32_t wg_vec_addr = smem_u32(wg_vec);
        const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
        uint32_t n = 0, ti = 0, my_chunks = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++ti) {
            const int m0 = (reverse ? n_tiles - 1 - tile : tile) * 256 + cta_rank * 128;
            const int row = m0 + row_in_tile;
            // LayerNorm factors of this thread's row from the partial sums the producer of xb left behind
            float rstd, nrm;
            {
                float sx = 0.f, sq = 0.f;
                if (row < M) {
                    const float2* sp = reinterpret_cast<const float2*>(stats_in) + row;
#pragma unroll
                    for (int p = 0; p < 6; ++p) {
                        const float2 t = __ldg(sp + static_cast<size_t>(p) * stats_stride);
                        sx += t.x; sq += t.y;
                    }
                }
                const float mu = sx * (1.0f / MLP_D);
                const float var = fmaxf(sq * (1.0f / MLP_D) - mu * mu, 0.f);
                rstd = rsqrtf(var + eps);
                nrm = -rstd * mu;
            }
            const f32x2_t rstd2 = f2_pack(rstd, rstd), nrm2 = f2_pack(nrm, nrm);
            // this warpgroup's slices of (c, d) for its first chunk of the tile
            float vnext = (t_wg < 64) ? __ldg(c1 + g * MLP_CH + t_wg) : __ldg(d1 + g * MLP_CH + t_wg - 64);

            for (int j = g; j < MLP_NCH; j += 4) {
                const uint32_t nn = n + j, s = nn & 1, u = nn >> 1;
                // ---- column vectors of chunk j into the warpgroup's slice (all warps are done with the previous one)
                named_bar_sync(1 + g, 128);
                sts_f1(wg_vec_addr + t_wg * 4, vnext);
                if (j + 4 < MLP_NCH)
                    vnext = (t_wg < 64) ? __ldg(c1 + (j + 4) * MLP_CH + t_wg) : __ldg(d1 + (j + 4) * MLP_CH + t_wg - 64);
                named_bar_sync(1 + g, 128);
                // ---- S_j: TMEM -> registers
                if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + (j >> 2) * 8 + 0);
                mbar_wait(&s_full[g], my_chunks & 1);            // each warpgroup sees every phase of its own barrier
                if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + (j >> 2) * 8 + 1);
                ++my_chunks;
                tc_fence_after();
                uint32_t pk[32];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld_32x32(lane_taddr + MLP_TMEM_S + s * MLP_CH + h * 32, v);
                    tmem_ld_wait();
                    if (h == 1) {                                 // S buffer free for chunk j + 2
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_leader(&s_empty[s]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 cc = lds_f4(wg_vec_addr + (h * 32 + 4 * i) * 4);
                        const float4 dd = lds_f4(wg_vec_addr + (64 + h * 32 + 4 * i) * 4);
                        f32x2_t y0 = f2_pack(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1]));
                        f32x2_t y1 = f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                        y0 = f2_fma(rstd2, y0, f2_fma(nrm2, f2_pack(cc.x, cc.y), f2_pack(dd.x, dd.y)));
                        y1 = f2_fma(rstd2, y1, f2_fma(nrm2, f2_pack(cc.z, cc.w), f2_pack(dd.z, dd.w)));
                        pk[h * 16 + 2 * i] = gelu_fast2x2_bf16(y0);
                        pk[h * 16 + 2 * i + 1] = gelu_fast2x2_bf16(y1);
                    }
                }
                // ---- H_j: registers -> swizzled shared memory (the MMAs that read this buffer two chunks ago have retired)
                if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + (j >> 2) * 8 + 2);
                mbar_wait(&h_empty[s], (u & 1) ^ 1);
                if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + (j >> 2) * 8 + 3);
                const uint32_t h_row = smem_u32(sH + s * MLP_H_BYTES) + row_in_tile * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    sts_u4(h_row + ((q ^ sw) << 4), make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) arrive_leader(&h_full[s]);
                if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + (j >> 2) * 8 + 4);
            }
            if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + 64);

            // ---- tile epilogue: x_new = bf16(x_old + O + b2) in place in the A tile, partial statistics, TMA store
            mbar_wait(o_full, ti & 1);
            if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + 65);
            tc_fence_after();
            for (int c = g; c < MLP_KB; c += 4) {
                const bool last = (c + 4 >= MLP_KB);
                const float4* b4 = reinterpret_cast<const float4*>(b2 + c * 64);        // broadcast loads (L1-resident)
                const uint32_t a_row = smem_u32(sA + c * 16384) + row_in_tile * 128;
                f32x2_t st_sum2 = f2_pack(0.f, 0.f), st_sq2 = st_sum2;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld_32x32(lane_taddr + c * 64 + h * 32, v);
                    tmem_ld_wait();
                    if (h == 1 && last) {                         // accumulator free for the next tile
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_leader(o_empty);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t addr = a_row + (((h * 4 + q) ^ sw) << 4);
                        uint4 rr;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w) : "r"(addr));
                        const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
                        const float4 ba = __ldg(b4 + h * 8 + 2 * q), bb = __ldg(b4 + h * 8 + 2 * q + 1);
                        const f32x2_t bias2[4] = {f2_pack(ba.x, ba.y), f2_pack(ba.z, ba.w), f2_pack(bb.x, bb.y), f2_pack(bb.z, bb.w)};
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {            // packed pairs: (residual + bias) + accumulator, same rounding
                            const f32x2_t res2 = f2_pack(__uint_as_float(rw[e] << 16), __uint_as_float(rw[e] & 0xffff0000u));
                            const f32x2_t acc2 = f2_pack(__uint_as_float(v[8 * q + 2 * e]), __uint_as_float(v[8 * q + 2 * e + 1]));
                            const f32x2_t o2 = f2_add(f2_add(res2, acc2), bias2[e]);
                            st_sum2 = f2_add(st_sum2, o2);
                            st_sq2 = f2_fma(o2, o2, st_sq2);
                            float o0, o1;
                            f2_unpack(o2, o0, o1);
                            pk[e] = pack_bf16x2(o0, o1);
                        }
                        sts_u4(addr, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    }
                }
                if (row < M) {
                    float s0, s1, q0, q1;
                    f2_unpack(st_sum2, s0, s1);
                    f2_unpack(st_sq2, q0, q1);
                    *reinterpret_cast<float2*>(stats_out + (static_cast<size_t>(c) * stats_stride + row) * 2) = make_float2(s0 + s1, q0 + q1);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_out, sA + c * 16384 + ew * 4096, c * 64, m0 + ew * 32);
                    tma_store_commit();
                }
            }
            // the A tile may be overwritten once this warp's stores have read it
            if (lane == 0) { tma_store_wait_read<0>(); mbar_arrive(a_empty); }
            __syncwarp();
            if (warp == 4 && lane == 0) MTRACE(4096 + ti * 128 + 66);
            n += MLP_NCH;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

int mlp_fused_launch(const void* xb_bf16, const void* w1g_bf16, const float* c1, const float* d1, const void* w2h_bf16,
                     const float* b2, const float* stats_in, float* stats_out, int stats_stride, float eps, int M,
                     cudaStream_t stream) {
    if (M <= 0) return 0;
    if (stats_stride < M) return set_error("hb_mlp_fused: statistics plane stride %d < M %d", stats_stride, M);
    CUtensorMap map_a, map_w1, map_w2, map_out;
    if (encode_tmap_2d(&map_a, TMAP_BF16, xb_bf16, M, MLP_D, MLP_D * 2, 128, 64)) return -1;
    if (encode_tmap_2d(&map_w1, TMAP_BF16, w1g_bf16, MLP_H, MLP_D, MLP_D * 2, 32, 64)) return -1;
    if (encode_tmap_2d(&map_w2, TMAP_BF16, w2h_bf16, MLP_D, MLP_H, MLP_H * 2, 96, 64)) return -1;
    if (encode_tmap_2d(&map_out, TMAP_BF16, xb_bf16, M, MLP_D, MLP_D * 2, 32, 64)) return -1;
    if (set_max_dynamic_smem(reinterpret_cast<const void*>(mlp_fused_kernel), MLP_SMEM)) return -1;
    const int n_tiles = (M + 255) / 256;
    const int slots = num_sms() / 2;
    const int grid = (n_tiles < slots ? n_tiles : slots) * 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(MLP_THREADS);
    cfg.dynamicSmemBytes = MLP_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // Tiles are walked from the LAST one down: the proj GEMM before this kernel wrote the residual stream in ascending row order
    // (its last rows are the ones still in L2), and the next qkv GEMM reads it in ascending order, starting with the rows this
    // kernel writes last (HB_MLP_REVERSE=0: ascending, for comparison).
    static int rev = -1;
    if (rev < 0) { const char* e = getenv("HB_MLP_REVERSE"); rev = (e && e[0] == '0') ? 0 : 1; }
    HB_CUDA_OK(cudaLaunchKernelEx(&cfg, mlp_fused_kernel, map_a, map_w1, map_w2, map_out, c1, d1, b2, stats_in, stats_out,
                                  stats_stride, eps, M, rev));
    count_launch();
    return 0;
}

}  // namespace hb
