// hb_embed.cu — unfold + ToTensor/Normalize + patch-embed conv as ONE tcgen05 GEMM that reads the uint8 region from HBM.
//
// Replaces `unfold(2,256,256).unfold(3,256,256)` + `rearrange` (HIPT_4K/hipt_4k.py:64-65), eval_transforms
// (hipt_model_utils.py:113-118, folded into the weights) and PatchEmbed.forward's 16x16/16 conv + flatten + transpose
// (vision_transformer.py:165-170) + the positional add of prepare_tokens (:240-244) for the patch tokens.
//
// GEMM view: M = tokens (row-major 16 x 16 per 256 x 256 patch), N = 384, K = 768 ordered (c, i, j) like
// conv.weight.reshape(384, 768).  A is never materialised in HBM: per 128-token tile (8 token rows x 16 token columns = half a
// patch) and per K-slice of 64 (one channel, 4 pixel rows of every token) TMA brings the raw bytes [8][4][256] into shared
// memory with a 5-D box over (x, row-in-token, token row, channel, image); eight warps turn them into the K-major
// SWIZZLE_128B fp16 A tile (byte b -> 0x6400 | b = 1024 + b exactly, minus 1024: two PRMT + one HSUB2 per four pixels);
// the B operand is fp16(W / std) streamed from L2 by TMA (two 192-row halves per slice); tcgen05 kind::f16 accumulates
// 128 x 384 fp32 in TMEM.  Epilogue: y = acc / 255 + (b - sum W mean / std) + pos[1 + token], rounded once to bf16 into the
// residual stream, plus the per-64-column (sum, sum of squares) planes the LayerNorm folded into the first qkv GEMM needs
// (plain stores, fixed summation order: bit-identical whatever else shares the launch).
//   warp 0      TMA producer (3-stage ring: 8 KB of pixels + 48 KB of weights per stage)
//   warp 1      MMA issuer, TMEM allocator (512 columns, one 384-column accumulator)
//   warps 2-9   uint8 -> fp16 converters
//   warps 10-17 epilogue: thread = (token, half of the 384 columns)
#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

constexpr int EM_THREADS = 576, EM_STAGES = 3;
constexpr int EM_A = 16384, EM_W = 49152, EM_U8 = 8192, EM_STAGE = EM_A + EM_W + EM_U8;      // 73728 = 72 x 1024
constexpr int EM_SMEM = 1024 + EM_STAGES * EM_STAGE + 256;
constexpr int EM_NSL = 12;                                   // K slices of 64: (channel, 4 pixel rows)

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}
// kind::f16 with fp16 A / B (format 0), fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// four pixels -> four exact fp16 values (two packed words)
__device__ __forceinline__ void u8x4_to_f16x4(uint32_t w, uint32_t& lo, uint32_t& hi) {
    uint32_t a = __byte_perm(w, 0x64646464u, 0x4140), b = __byte_perm(w, 0x64646464u, 0x4342);   // 1024 + pixel
    asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(0x64006400u));
    asm("sub.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(b), "r"(0x64006400u));
}

__global__ void __launch_bounds__(EM_THREADS, 1)
embed_u8_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_w,
                const float* __restrict__ bias, float scale, const float* __restrict__ pos, __nv_bfloat16* __restrict__ xb,
                float* __restrict__ stats, int stats_stride, int grid_cols, int ppi, int patch_begin, int n_patches) {
    extern __shared__ uint8_t smem_raw_em[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_em) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + EM_STAGES * EM_STAGE);
    uint64_t* ld_full = bars;             // [3] TMA bytes (pixels + weights)
    uint64_t* a_full = bars + 3;          // [3] 256 converter threads
    uint64_t* st_empty = bars + 6;        // [3] MMA commit
    uint64_t* acc_full = bars + 9;
    uint64_t* acc_empty = bars + 10;      // 256 epilogue threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_img); tma_prefetch_desc(&map_w); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < EM_STAGES; ++i) { mbar_init(&ld_full[i], 1); mbar_init(&a_full[i], 256); mbar_init(&st_empty[i], 1); }
        mbar_init(acc_full, 1); mbar_init(acc_empty, 256);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = 2 * n_patches;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t q = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int g = patch_begin + (tile >> 1), h = tile & 1;
                const int img = g / ppi, local = g - img * ppi;
                const int py = local / grid_cols, px = local - py * grid_cols;
                for (int sl = 0; sl < EM_NSL; ++sl, ++q) {
                    const uint32_t st = q % EM_STAGES, use = q / EM_STAGES;
                    uint8_t* stage = smem + st * EM_STAGE;
                    mbar_wait(&st_empty[st], (use & 1) ^ 1);
                    mbar_arrive_expect_tx(&ld_full[st], EM_U8 + EM_W);
                    tma_load_5d(stage + EM_A + EM_W, &map_img, &ld_full[st], px * 256, (sl & 3) * 4, py * 16 + 8 * h, sl >> 2, img);
                    tma_load_2d(stage + EM_A, &map_w, &ld_full[st], sl * 64, 0);
                    tma_load_2d(stage + EM_A + EM_W / 2, &map_w, &ld_full[st], sl * 64, 192);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_f16(128, 192);
        uint32_t q = 0, t = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            if (t > 0) mbar_wait(acc_empty, (t - 1) & 1);
            tc_fence_after();
            for (int sl = 0; sl < EM_NSL; ++sl, ++q) {
                const uint32_t st = q % EM_STAGES, use = q / EM_STAGES;
                const uint32_t stage = smem_u32(smem + st * EM_STAGE);
                const uint64_t da = umma_desc_k128(stage), db0 = umma_desc_k128(stage + EM_A), db1 = umma_desc_k128(stage + EM_A + EM_W / 2);
                mbar_wait(&a_full[st], use & 1);             // converters arrive after the stage's TMA bytes have landed
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        umma_bf16_ss(tmem_base, da + 2 * kk, db0 + 2 * kk, idesc, (sl | kk) != 0);
                        umma_bf16_ss(tmem_base + 192, da + 2 * kk, db1 + 2 * kk, idesc, (sl | kk) != 0);
                    }
                    umma_commit(&st_empty[st]);
                    if (sl == EM_NSL - 1) umma_commit(acc_full);
                }
                __syncwarp();
            }
        }
    } else if (warp < 10) {
        const int ct = tid - 64;                                 // 0..255
        uint32_t q = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int sl = 0; sl < EM_NSL; ++sl, ++q) {
                const uint32_t st = q % EM_STAGES, use = q / EM_STAGES;
                const uint32_t stage = smem_u32(smem + st * EM_STAGE);
                mbar_wait(&ld_full[st], use & 1);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int item = ct + 256 * u;
                    const int t = item & 127, il = item >> 7;        // token of the tile, pixel row of the slice
                    const uint4 px = lds_u4(stage + EM_A + EM_W + (t >> 4) * 1024 + il * 256 + (t & 15) * 16);
                    uint4 o0, o1;
                    u8x4_to_f16x4(px.x, o0.x, o0.y); u8x4_to_f16x4(px.y, o0.z, o0.w);
                    u8x4_to_f16x4(px.z, o1.x, o1.y); u8x4_to_f16x4(px.w, o1.z, o1.w);
                    const uint32_t row = stage + t * 128;
                    sts_u4(row + (((2 * il) ^ (t & 7)) << 4), o0);
                    sts_u4(row + (((2 * il + 1) ^ (t & 7)) << 4), o1);
                }
                fence_proxy_async_smem();
                mbar_arrive(&a_full[st]);
            }
        }
    } else {
        const int rr = (warp & 3) * 32 + lane;                   // token inside the tile = TMEM lane
        const int ch = (warp - 10) >> 2;                         // column half
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ch * 192;
        const float4* bias4 = reinterpret_cast<const float4*>(bias + ch * 192);
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const int pl = tile >> 1, tok = (tile & 1) * 128 + rr;
            const size_t xrow = static_cast<size_t>(pl) * 257 + 1 + tok;
            const float4* pos4 = reinterpret_cast<const float4*>(pos + static_cast<size_t>(1 + tok) * 384 + ch * 192);
            uint4* out = reinterpret_cast<uint4*>(xb + xrow * 384 + ch * 192);
            mbar_wait(acc_full, t & 1);
            tc_fence_after();
            float s_sum = 0.f, s_sq = 0.f;
#pragma unroll 1
            for (int cb = 0; cb < 6; ++cb) {
                uint32_t v[32];
                tmem_ld_32x32(t_lane + cb * 32, v);
                tmem_ld_wait();
                if (cb == 5) { tc_fence_before(); mbar_arrive(acc_empty); }
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = __ldg(bias4 + cb * 8 + j), p = __ldg(pos4 + cb * 8 + j);
                    const float y0 = fmaf(__uint_as_float(v[4 * j + 0]), scale, b.x + p.x);
                    const float y1 = fmaf(__uint_as_float(v[4 * j + 1]), scale, b.y + p.y);
                    const float y2 = fmaf(__uint_as_float(v[4 * j + 2]), scale, b.z + p.z);
                    const float y3 = fmaf(__uint_as_float(v[4 * j + 3]), scale, b.w + p.w);
                    s_sum += (y0 + y1) + (y2 + y3);
                    s_sq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, s_sq))));
                    pk[2 * j] = pack_bf16x2(y0, y1);
                    pk[2 * j + 1] = pack_bf16x2(y2, y3);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) out[cb * 4 + j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                if (cb & 1) {                                    // one statistics plane per 64 columns
                    const int plane = ch * 3 + (cb >> 1);
                    *reinterpret_cast<float2*>(stats + (static_cast<size_t>(plane) * stats_stride + xrow) * 2) = make_float2(s_sum, s_sq);
                    s_sum = 0.f; s_sq = 0.f;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int embed_u8_launch(const void* image_u8, size_t chan_stride, size_t row_pitch, int grid_cols, int grid_rows,
                    size_t image_stride_bytes, int n_images, int patch_begin, int n_patches, const void* w_f16,
                    const float* bias, float scale, const float* pos_table, void* xb_bf16, float* stats, int stats_stride,
                    cudaStream_t stream) {
    if (n_patches <= 0) return 0;
    CUtensorMap map_img, map_w;
    // (x, row inside the token, token row, channel, image)
    const uint64_t dims[5] = {static_cast<uint64_t>(grid_cols) * 256, 16, static_cast<uint64_t>(grid_rows) * 16, 3,
                              static_cast<uint64_t>(n_images)};
    const uint64_t strides[4] = {row_pitch, 16 * row_pitch, chan_stride, image_stride_bytes ? image_stride_bytes : 3 * chan_stride};
    const uint32_t box[5] = {256, 4, 8, 1, 1};
    if (encode_tmap_u8_nd(&map_img, image_u8, 5, dims, strides, box)) return -1;
    if (encode_tmap_2d(&map_w, TMAP_BF16, w_f16, 384, 768, 768 * 2, 192, 64)) return -1;      // 2-byte elements: fp16 bits
    if (set_max_dynamic_smem(reinterpret_cast<const void*>(embed_u8_kernel), EM_SMEM)) return -1;
    const int n_tiles = 2 * n_patches;
    const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
    embed_u8_kernel<<<grid, EM_THREADS, EM_SMEM, stream>>>(map_img, map_w, bias, scale, pos_table,
                                                           static_cast<__nv_bfloat16*>(xb_bf16), stats, stats_stride, grid_cols,
                                                           grid_cols * grid_rows, patch_begin, n_patches);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hb
