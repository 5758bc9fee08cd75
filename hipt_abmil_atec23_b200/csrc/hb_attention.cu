// hb_attention.cu — fused single-pass softmax(Q K^T * scale) V for short sequences (257 tokens in both ViTs).
//
// Replaces Attention.forward's q@k^T / softmax / attn@v (HIPT_4K/vision_transformer.py:119-128 and
// vision_transformer4k.py:125-136), which materialises a [B, heads, 257, 257] fp32 matrix per block.
// Input is the qkv GEMM output [n_seq*seq_len, 3*heads*hd] bf16 with output channels ordered q | k | v, head-major
// (the reshape(B,N,3,H,hd).permute(2,0,3,1,4) of the reference); output is [n_seq*seq_len, heads*hd] bf16, i.e. the
// (attn @ v).transpose(1,2).reshape(B,N,C) layout the proj Linear consumes.
//
// One CTA per (sequence, head): Q, K, V are staged once in shared memory (cp.async, XOR-swizzled 16 B chunks), each
// warp owns 16-query tiles and runs a flash-style online softmax over 64-key chunks with bf16 tensor-core MMAs
// (fp32 accumulate, fp32 softmax statistics, exp2 with log2(e) folded into the scale).
#include <stdlib.h>

#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

static bool attention_force_legacy() {      // test hook: HB_ATTENTION_LEGACY=1 routes every shape through mma.sync
    static int v = -1;
    if (v < 0) { const char* e = getenv("HB_ATTENTION_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <int HD>
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) {
    if constexpr (HD == 64) return row * 128 + ((chunk ^ (row & 7)) << 4);
    else return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One chunk of NG*16 keys starting at key0: S = Q K^T, online-softmax update, O += P V.
template <int HD, int NG, bool MASK>
__device__ __forceinline__ void attn_chunk(const uint32_t (&qf)[HD / 16][4], uint32_t sK, uint32_t sV, int key0,
                                           int seq_len, float scale_log2, float (&o)[HD / 8][4], float (&m)[2],
                                           float (&l)[2], int lane) {
    constexpr int NT = NG * 2;                   // 8-key n-tiles in this chunk
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
    // S = Q K^T
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int key = key0 + g * 16 + (lane & 7) + ((lane >> 4) << 3);
            const int chunk = kk * 2 + ((lane >> 3) & 1);
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sK + swz_off<HD>(key, chunk), b0, b1, b2, b3);
            mma_bf16_16816(s[2 * g], qf[kk], b0, b1);
            mma_bf16_16816(s[2 * g + 1], qf[kk], b2, b3);
        }
    }
    // scale (+ mask the padded keys)
    const int t2 = (lane & 3) * 2;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = s[j][e] * scale_log2;
            if (MASK) {
                const int key = key0 + j * 8 + t2 + (e & 1);
                if (key >= seq_len) v = -INFINITY;
            }
            s[j][e] = v;
        }
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m[0], mx0), mn1 = fmaxf(m[1], mx1);
    const float a0 = exp2f(m[0] - mn0), a1 = exp2f(m[1] - mn1);
    m[0] = mn0; m[1] = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s[j][0] = exp2f(s[j][0] - mn0); s[j][1] = exp2f(s[j][1] - mn0);
        s[j][2] = exp2f(s[j][2] - mn1); s[j][3] = exp2f(s[j][3] - mn1);
        rs0 += s[j][0] + s[j][1];
        rs1 += s[j][2] + s[j][3];
    }
    l[0] = l[0] * a0 + rs0;
    l[1] = l[1] * a1 + rs1;
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) { o[d][0] *= a0; o[d][1] *= a0; o[d][2] *= a1; o[d][3] *= a1; }
    // O += P V
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * g][0], s[2 * g][1]);
        pa[1] = pack_bf16x2(s[2 * g][2], s[2 * g][3]);
        pa[2] = pack_bf16x2(s[2 * g + 1][0], s[2 * g + 1][1]);
        pa[3] = pack_bf16x2(s[2 * g + 1][2], s[2 * g + 1][3]);
        const int key = key0 + g * 16 + (lane & 15);
#pragma unroll
        for (int dd = 0; dd < HD / 16; ++dd) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(sV + swz_off<HD>(key, dd * 2 + (lane >> 4)), b0, b1, b2, b3);
            mma_bf16_16816(o[2 * dd], pa, b0, b1);
            mma_bf16_16816(o[2 * dd + 1], pa, b2, b3);
        }
    }
}

template <int HD>
__global__ void __launch_bounds__(288, 2) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                        __nv_bfloat16* __restrict__ out, int seq_len, int heads,
                                                        float scale_log2, int cls_only, float* __restrict__ cls_probs) {
    // cls_only: only query row 0 of each sequence is computed and written to a COMPACT [n_seq, D] output — all that
    // the last transformer block needs, because forward() returns x[:, 0] (vision_transformer.py:252-253).
    // cls_probs (cls_only mode, may be NULL): [n_seq, heads, seq_len] fp32, the softmax row of query 0 — what the heatmap
    // code reads from get_last_selfattention as attention[:, :, 0, :] (hipt_4k.py:145-147, 155-157).
    extern __shared__ __align__(128) uint8_t smem_attn[];
    constexpr int CH = HD / 8;                              // 16 B chunks per row
    const int s_pad = (seq_len + 15) & ~15;
    const int D = heads * HD;
    const int seq = blockIdx.x / heads, h = blockIdx.x - seq * heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const uint32_t mat_bytes = s_pad * HD * 2;
    const uint32_t sQ = smem_u32(smem_attn), sK = sQ + mat_bytes, sV = sK + mat_bytes;

    // stage Q, K, V of this (sequence, head)
    const __nv_bfloat16* base = qkv + static_cast<size_t>(seq) * seq_len * 3 * D + h * HD;
    const int per_mat = s_pad * CH;
    for (int idx = threadIdx.x; idx < 3 * per_mat; idx += blockDim.x) {
        const int which = idx / per_mat;
        const int rem = idx - which * per_mat;
        const int r = rem / CH, c = rem - r * CH;
        const uint32_t dst = sQ + which * mat_bytes + swz_off<HD>(r, c);
        if (cls_only && which == 0 && r >= 16) continue;
        if (r < seq_len) {
            cp_async_16(dst, base + static_cast<size_t>(r) * 3 * D + which * D + c * 8);
        } else {
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int q_tiles = cls_only ? 1 : s_pad / 16;
    const int full_chunks = s_pad / 64;
    const int tail_groups = (s_pad - full_chunks * 64) / 16;    // 0..3 groups of 16 keys
    const bool full_has_pad = (full_chunks * 64 > seq_len);     // only when tail_groups == 0 and seq_len % 64 != 0

    for (int qt = warp; qt < q_tiles; qt += n_warps) {
        uint32_t qf[HD / 16][4];
        {
            const int row = qt * 16 + (lane & 15);
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk)
                ldsm_x4(sQ + swz_off<HD>(row, kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
        }
        float o[HD / 8][4];
#pragma unroll
        for (int d = 0; d < HD / 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
        float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

        for (int c = 0; c < full_chunks; ++c) {
            if (full_has_pad && c == full_chunks - 1)
                attn_chunk<HD, 4, true>(qf, sK, sV, c * 64, seq_len, scale_log2, o, m, l, lane);
            else
                attn_chunk<HD, 4, false>(qf, sK, sV, c * 64, seq_len, scale_log2, o, m, l, lane);
        }
        const int k0 = full_chunks * 64;
        if (tail_groups == 1) attn_chunk<HD, 1, true>(qf, sK, sV, k0, seq_len, scale_log2, o, m, l, lane);
        else if (tail_groups == 2) attn_chunk<HD, 2, true>(qf, sK, sV, k0, seq_len, scale_log2, o, m, l, lane);
        else if (tail_groups == 3) attn_chunk<HD, 3, true>(qf, sK, sV, k0, seq_len, scale_log2, o, m, l, lane);

        float l0 = l[0], l1 = l[1];
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        const int r0 = qt * 16 + (lane >> 2), r1 = r0 + 8;
        if (cls_only) {
            if ((lane >> 2) == 0) {
                __nv_bfloat16* orow = out + static_cast<size_t>(seq) * D + h * HD + (lane & 3) * 2;
#pragma unroll
                for (int d = 0; d < HD / 8; ++d)
                    *reinterpret_cast<uint32_t*>(orow + d * 8) = pack_bf16x2(o[d][0] * inv0, o[d][1] * inv0);
            }
            if (cls_probs != nullptr) {
                // second pass over the keys for query 0 only: p_k = exp2(s_k * scale - m) / l with the final (m, l)
                float* prow = cls_probs + (static_cast<size_t>(seq) * heads + h) * seq_len;
                const float mrow = __shfl_sync(0xffffffffu, m[0], 0);
                const int t2 = (lane & 3) * 2;
                for (int key0 = 0; key0 < s_pad; key0 += 16) {
                    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int kk = 0; kk < HD / 16; ++kk) {
                        const int key = key0 + (lane & 7) + ((lane >> 4) << 3);
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4(sK + swz_off<HD>(key, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
                        mma_bf16_16816(c0, qf[kk], b0, b1);
                        mma_bf16_16816(c1, qf[kk], b2, b3);
                    }
                    if ((lane >> 2) == 0) {
                        const float v[4] = {c0[0], c0[1], c1[0], c1[1]};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int key = key0 + (e >> 1) * 8 + t2 + (e & 1);
                            if (key < seq_len) prow[key] = exp2f(v[e] * scale_log2 - mrow) * inv0;
                        }
                    }
                }
            }
            continue;
        }
        __nv_bfloat16* obase = out + static_cast<size_t>(seq) * seq_len * D + h * HD + (lane & 3) * 2;
#pragma unroll
        for (int d = 0; d < HD / 8; ++d) {
            if (r0 < seq_len)
                *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(r0) * D + d * 8) = pack_bf16x2(o[d][0] * inv0, o[d][1] * inv0);
            if (r1 < seq_len)
                *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(r1) * D + d * 8) = pack_bf16x2(o[d][2] * inv1, o[d][3] * inv1);
        }
    }
}

// =====================================================================================================================
// tcgen05 path for the ViT-256 shape (seq_len 257, head_dim 64).
//
// One persistent CTA per SM walks (sequence, head) items.  An item is two 128-query tiles plus the 257th query; the two
// tiles run as two INDEPENDENT pipelines (A: queries 0..127, B: 128..255) so that one pipeline's exp2 phase overlaps the
// other's TMEM loads / waits / epilogue.
//   warp 0        TMA producer: K [272 x 64] and V [272 x 64] (double buffered across items) and one Q tile per pipeline,
//                 SWIZZLE_128B boxes of the qkv matrix (rows past its end read as zero)
//   warps 1, 2    MMA issuers of pipelines A and B: S = Q K^T (M128 x N256, keys 0..255) into the pipeline's 256 TMEM
//                 columns, then O += P V over five key atoms (4 x 64 keys + key 256) with P read K-major from shared
//                 memory and V read MN-major in place; O accumulates into S's first 64 columns, which the softmax has
//                 already consumed when the first atom arrives
//   warp 3        the 257th query row on the mma.sync path against the same K / V tiles
//   warps 4-7     softmax warpgroup of pipeline A, warps 8-11 of pipeline B; thread = query row.  The score against key
//                 256 is a 64-term dot product in the thread; pass 1 reads the row's 256 scores from TMEM (first 128 kept
//                 in registers) for the max, pass 2 produces P = exp2((s - max) * scale * log2 e) atom by atom as bf16 in
//                 the swizzled layout the MMA consumes; the epilogue scales O by 1/sum and stores the tile with TMA.
// Softmax statistics stay fp32; P is rounded to bf16 exactly like the mma.sync kernel.
// =====================================================================================================================
constexpr int ATC_THREADS = 384;
constexpr int ATC_S = 257;
constexpr int ATC_SPAD = 272;
constexpr int ATC_KV_BYTES = ATC_SPAD * 128;            // 34816: [272 keys][64 bf16]
constexpr int ATC_TILE_BYTES = 128 * 128;               // 16384: one 128-row x 128-byte swizzled tile
constexpr int ATC_SMEM = 3 * ATC_KV_BYTES + 6 * ATC_TILE_BYTES + 512 + 1024;
constexpr int ATC_TMEM_COLS = 512;

__device__ __forceinline__ void tmem_ld_x32_ptr(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map16,
                    const __grid_constant__ CUtensorMap map_out, const __nv_bfloat16* __restrict__ qkv,
                    __nv_bfloat16* __restrict__ out, int n_items, int heads, float scale_log2) {
    extern __shared__ uint8_t smem_raw_atc[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_atc) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem;                                   // [34816]            K is dead once Q K^T is done: one buffer
    uint8_t* sV = sK + ATC_KV_BYTES;                      // [2 stages][34816]
    uint8_t* sQ = sV + 2 * ATC_KV_BYTES;                  // [2 pipelines][16384]
    uint8_t* sP = sQ + 2 * ATC_TILE_BYTES;                // [2 pipelines][2 slots][16384]  P atoms; slot 0 also stages O
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * ATC_TILE_BYTES);
    uint64_t* k_full = bars;             // [1]
    uint64_t* k_empty = bars + 1;        // [1]  2 MMA commits + 256 softmax threads (key-256 dot) + tail warp
    uint64_t* v_full = bars + 2;         // [2 stages]
    uint64_t* v_empty = bars + 4;        // [2 stages]  2 MMA commits + tail warp
    uint64_t* q_full = bars + 6;         // [2 pipelines]
    uint64_t* q_empty = bars + 8;        // [2]  MMA commit + 128 softmax threads (they read their Q row for key 256)
    uint64_t* s_full = bars + 10;        // [2]
    uint64_t* o_full = bars + 12;        // [2]
    uint64_t* o_free = bars + 14;        // [2]  128 softmax threads: O (= the S region) may be overwritten
    uint64_t* p_full = bars + 16;        // [2 pipelines][2 slots]  128 softmax threads
    uint64_t* p_empty = bars + 20;       // [2][2]  MMA commit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = heads * 64;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map128); tma_prefetch_desc(&map16); tma_prefetch_desc(&map_out); }
    if (warp == 1 && lane == 0) {
        mbar_init(k_full, 1); mbar_init(k_empty, 2 + 256 + 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 3);
            mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 129);
            mbar_init(&s_full[i], 1);
            mbar_init(&o_full[i], 1);  mbar_init(&o_free[i], 128);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1); }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, ATC_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int my_items = 0;
    if (static_cast<int>(blockIdx.x) < n_items) my_items = (n_items - 1 - blockIdx.x) / gridDim.x + 1;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int seq = item / heads, h = item - seq * heads;
                const int row0 = seq * ATC_S;
                const int st = it & 1;
                mbar_wait(k_empty, (it & 1) ^ 1);
                mbar_arrive_expect_tx(k_full, ATC_KV_BYTES);
                tma_load_2d(sK, &map128, k_full, D + h * 64, row0);
                tma_load_2d(sK + ATC_TILE_BYTES, &map128, k_full, D + h * 64, row0 + 128);
                tma_load_2d(sK + 2 * ATC_TILE_BYTES, &map16, k_full, D + h * 64, row0 + 256);
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    mbar_wait(&q_empty[x], (it & 1) ^ 1);
                    mbar_arrive_expect_tx(&q_full[x], ATC_TILE_BYTES);
                    tma_load_2d(sQ + x * ATC_TILE_BYTES, &map128, &q_full[x], h * 64, row0 + x * 128);
                }
                mbar_wait(&v_empty[st], ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&v_full[st], ATC_KV_BYTES);
                uint8_t* v = sV + st * ATC_KV_BYTES;
                tma_load_2d(v, &map128, &v_full[st], 2 * D + h * 64, row0);
                tma_load_2d(v + ATC_TILE_BYTES, &map128, &v_full[st], 2 * D + h * 64, row0 + 128);
                tma_load_2d(v + 2 * ATC_TILE_BYTES, &map16, &v_full[st], 2 * D + h * 64, row0 + 256);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------------------------------------ MMA issuers
        // The whole warp walks the pipeline (warp-uniform control flow and operands keep the descriptors in uniform
        // registers); one elected lane issues the tcgen05.mma / commit instructions.
        {
            const int x = warp - 1;                                                // pipeline
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
            const uint32_t t_acc = tmem_base + x * 256;
            const uint64_t dq = umma_desc_k128(smem_u32(sQ + x * ATC_TILE_BYTES));
            const uint64_t dk = umma_desc_k128(smem_u32(sK));
            const uint32_t p_base = smem_u32(sP + (x * 2) * ATC_TILE_BYTES);
            uint32_t use0 = 0, use1 = 0;                                           // uses of P slot 0 / 1 so far
            for (int it = 0; it < my_items; ++it) {
                const int st = it & 1;
                mbar_wait(k_full, it & 1);
                mbar_wait(&q_full[x], it & 1);
                if (it > 0) mbar_wait(&o_free[x], (it - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(t_acc, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
                    umma_commit(&s_full[x]);
                    umma_commit(&q_empty[x]);
                    umma_commit(k_empty);
                }
                __syncwarp();
                mbar_wait(&v_full[st], (it >> 1) & 1);
                const uint32_t v_base = smem_u32(sV + st * ATC_KV_BYTES);
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    const int slot = a & 1;
                    const uint32_t use = slot ? use1++ : use0++;
                    mbar_wait(&p_full[x * 2 + slot], use & 1);
                    tc_fence_after();
                    const uint64_t dp = umma_desc_k128(p_base + slot * ATC_TILE_BYTES);
                    constexpr int dummy = 0; (void)dummy;
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (a == 4 && kk > 0) break;                           // tail atom: key 256 only (K = 16)
                            const uint64_t dv = umma_desc_k128(v_base + (a * 64 + kk * 16) * 128);
                            umma_bf16_ss(t_acc, dp + 2 * kk, dv, idesc_pv, (a | kk) != 0);
                        }
                        umma_commit(&p_empty[x * 2 + slot]);
                        if (a == 4) { umma_commit(&o_full[x]); umma_commit(&v_empty[st]); }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------------------------------ query row 256
        // One warp, mma.sync: all 272 scores of the row first (K can then be released), one softmax pass in
        // registers, then P V.  Only fragment row 0 (lanes 0-3) is real; the other rows compute on zeros.
        const int t2 = (lane & 3) * 2;
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int seq = item / heads, h = item - seq * heads;
            const int st = it & 1;
            const size_t row = static_cast<size_t>(seq) * ATC_S + 256;
            uint32_t qf[4][4];
            {
                const uint32_t* qrow = reinterpret_cast<const uint32_t*>(qkv + row * 3 * D + h * 64);
                const bool own = (lane >> 2) == 0;            // A-fragment rows 0 (valid) and 8 (padding)
                const int t = lane & 3;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    qf[kk][0] = own ? __ldg(qrow + kk * 8 + t) : 0u;
                    qf[kk][2] = own ? __ldg(qrow + kk * 8 + 4 + t) : 0u;
                    qf[kk][1] = 0u; qf[kk][3] = 0u;
                }
            }
            mbar_wait(k_full, it & 1);
#ifdef ATC_SKIP_TAIL
            __syncwarp();
            if (lane == 0) { mbar_arrive(k_empty); mbar_wait(&v_full[st], (it >> 1) & 1); mbar_arrive(&v_empty[st]); }
            continue;
#endif
            const uint32_t k_addr = smem_u32(sK);
            float sc[34][2];
#pragma unroll
            for (int g = 0; g < 17; ++g) {
                float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int key = g * 16 + (lane & 7) + ((lane >> 4) << 3);
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(k_addr + swz_off<64>(key, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
                    mma_bf16_16816(c0, qf[kk], b0, b1);
                    mma_bf16_16816(c1, qf[kk], b2, b3);
                }
                sc[2 * g][0] = c0[0]; sc[2 * g][1] = c0[1];
                sc[2 * g + 1][0] = c1[0]; sc[2 * g + 1][1] = c1[1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(k_empty);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 34; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (j * 8 + t2 + e >= ATC_S) sc[j][e] = -INFINITY;
                    mx = fmaxf(mx, sc[j][e]);
                }
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float neg_m = -mx * scale_log2;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 34; ++j) {
                sc[j][0] = ex2_approx(fmaf(sc[j][0], scale_log2, neg_m));
                sc[j][1] = ex2_approx(fmaf(sc[j][1], scale_log2, neg_m));
                sum += sc[j][0] + sc[j][1];
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            mbar_wait(&v_full[st], (it >> 1) & 1);
            const uint32_t v_addr = smem_u32(sV + st * ATC_KV_BYTES);
            float o[8][4];
#pragma unroll
            for (int d = 0; d < 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
#pragma unroll
            for (int g = 0; g < 17; ++g) {
                uint32_t pa[4];
                pa[0] = pack_bf16x2(sc[2 * g][0], sc[2 * g][1]);
                pa[1] = 0u;
                pa[2] = pack_bf16x2(sc[2 * g + 1][0], sc[2 * g + 1][1]);
                pa[3] = 0u;
                const int key = g * 16 + (lane & 15);
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_t(v_addr + swz_off<64>(key, dd * 2 + (lane >> 4)), b0, b1, b2, b3);
                    mma_bf16_16816(o[2 * dd], pa, b0, b1);
                    mma_bf16_16816(o[2 * dd + 1], pa, b2, b3);
                }
            }
            if ((lane >> 2) == 0) {
                const float inv = 1.0f / sum;
                __nv_bfloat16* orow = out + row * D + h * 64 + t2;
#pragma unroll
                for (int d = 0; d < 8; ++d)
                    *reinterpret_cast<uint32_t*>(orow + d * 8) = pack_bf16x2(o[d][0] * inv, o[d][1] * inv);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&v_empty[st]);
        }
    } else {
        // ------------------------------------------------------------------------------------------ softmax warpgroups
        const int x = (warp - 4) >> 2;                        // pipeline = query tile of the item
        const int r = (warp & 3) * 32 + lane;                 // query row inside the tile = TMEM lane
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + x * 256;
        const int sw = r & 7;
        const bool leader = ((warp & 3) == 0 && lane == 0);
        uint8_t* q_row = sQ + x * ATC_TILE_BYTES + r * 128;
        uint8_t* p_row0 = sP + (x * 2) * ATC_TILE_BYTES + r * 128;
        uint8_t* p_row1 = p_row0 + ATC_TILE_BYTES;
        uint32_t use0 = 0, use1 = 0;                          // uses of P slot 0 / 1 so far
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int seq = item / heads, h = item - seq * heads;
            // ---- score against key 256: q_r . k_256 (row 256 of K is row 0 of its own swizzle atom: unswizzled)
            mbar_wait(k_full, it & 1);
            mbar_wait(&q_full[x], it & 1);
            float s256 = 0.f;
            {
                const uint4* k256 = reinterpret_cast<const uint4*>(sK + 2 * ATC_TILE_BYTES);
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 qv = *reinterpret_cast<const uint4*>(q_row + ((c ^ sw) << 4));
                    const uint4 kv = k256[c];
                    acc0 = fmaf(bf16_lo(qv.x), bf16_lo(kv.x), acc0); acc1 = fmaf(bf16_hi(qv.x), bf16_hi(kv.x), acc1);
                    acc0 = fmaf(bf16_lo(qv.y), bf16_lo(kv.y), acc0); acc1 = fmaf(bf16_hi(qv.y), bf16_hi(kv.y), acc1);
                    acc0 = fmaf(bf16_lo(qv.z), bf16_lo(kv.z), acc0); acc1 = fmaf(bf16_hi(qv.z), bf16_hi(kv.z), acc1);
                    acc0 = fmaf(bf16_lo(qv.w), bf16_lo(kv.w), acc0); acc1 = fmaf(bf16_hi(qv.w), bf16_hi(kv.w), acc1);
                }
                s256 = acc0 + acc1;
            }
            mbar_arrive(&q_empty[x]);
            mbar_arrive(k_empty);

            // ---- pass 1: row max over the 256 MMA scores (first 128 stay in registers) and key 256
            mbar_wait(&s_full[x], it & 1);
            tc_fence_after();
            uint32_t sv[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_x32_ptr(t_row + c * 32, sv + c * 32);
            tmem_ld_wait();
            float m0 = s256, m1 = __uint_as_float(sv[1]), m2 = __uint_as_float(sv[2]), m3 = __uint_as_float(sv[3]);
            m0 = fmaxf(m0, __uint_as_float(sv[0]));
#pragma unroll
            for (int j = 4; j < 128; j += 4) {
                m0 = fmaxf(m0, __uint_as_float(sv[j]));     m1 = fmaxf(m1, __uint_as_float(sv[j + 1]));
                m2 = fmaxf(m2, __uint_as_float(sv[j + 2])); m3 = fmaxf(m3, __uint_as_float(sv[j + 3]));
            }
#ifndef ATC_SKIP_RELOAD
#pragma unroll
            for (int c = 4; c < 8; ++c) {
                uint32_t tmp[32];
                tmem_ld_x32_ptr(t_row + c * 32, tmp);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    m0 = fmaxf(m0, __uint_as_float(tmp[j]));     m1 = fmaxf(m1, __uint_as_float(tmp[j + 1]));
                    m2 = fmaxf(m2, __uint_as_float(tmp[j + 2])); m3 = fmaxf(m3, __uint_as_float(tmp[j + 3]));
                }
            }
#endif
            const float neg_m = -fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2;

            // the TMA store of the previous tile must have drained P slot 0 (it doubles as the O staging buffer)
            if (leader) tma_store_wait_read<0>();
            named_bar_sync(4 + x, 128);

            // ---- pass 2: P atoms alternate between the two slots.  exp2 into registers, wait for the slot, publish.
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
#ifndef ATC_SKIP_RELOAD
                if (a == 2) {                                 // keys 128..255 come back from TMEM
                    tmem_ld_x32_ptr(t_row + 128, sv);
                    tmem_ld_x32_ptr(t_row + 160, sv + 32);
                    tmem_ld_x32_ptr(t_row + 192, sv + 64);
                    tmem_ld_x32_ptr(t_row + 224, sv + 96);
                    tmem_ld_wait();
                }
#endif
                uint32_t pk[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int e = (a & 1) * 64 + 2 * j;
                    const float p0 = ex2_approx(fmaf(__uint_as_float(sv[e]), scale_log2, neg_m));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(sv[e + 1]), scale_log2, neg_m));
                    sum0 += p0; sum1 += p1;
                    pk[j] = pack_bf16x2(p0, p1);
                }
                const int slot = a & 1;
                const uint32_t use = slot ? use1++ : use0++;
                mbar_wait(&p_empty[x * 2 + slot], (use & 1) ^ 1);
                uint8_t* p_row = slot ? p_row1 : p_row0;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<uint4*>(p_row + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                fence_proxy_async_smem();
                mbar_arrive(&p_full[x * 2 + slot]);
            }
            {                                                 // tail atom (slot 0): key 256 (+ 15 zero columns)
                const float p0 = ex2_approx(fmaf(s256, scale_log2, neg_m));
                sum0 += p0;
                const uint32_t use = use0++;
                mbar_wait(&p_empty[x * 2], (use & 1) ^ 1);
                *reinterpret_cast<uint4*>(p_row0 + ((0 ^ sw) << 4)) = make_uint4(pack_bf16x2(p0, 0.f), 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(p_row0 + ((1 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                fence_proxy_async_smem();
                mbar_arrive(&p_full[x * 2]);
            }
            const float inv = 1.0f / (sum0 + sum1);

            // ---- epilogue: O (first 64 columns of the S region) -> bf16 -> swizzled staging (P slot 0) -> TMA store
            mbar_wait(&o_full[x], it & 1);
            tc_fence_after();
            tmem_ld_x32_ptr(t_row, sv);
            tmem_ld_x32_ptr(t_row + 32, sv + 32);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_free[x]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint4 w;
                w.x = pack_bf16x2(__uint_as_float(sv[8 * q + 0]) * inv, __uint_as_float(sv[8 * q + 1]) * inv);
                w.y = pack_bf16x2(__uint_as_float(sv[8 * q + 2]) * inv, __uint_as_float(sv[8 * q + 3]) * inv);
                w.z = pack_bf16x2(__uint_as_float(sv[8 * q + 4]) * inv, __uint_as_float(sv[8 * q + 5]) * inv);
                w.w = pack_bf16x2(__uint_as_float(sv[8 * q + 6]) * inv, __uint_as_float(sv[8 * q + 7]) * inv);
                *reinterpret_cast<uint4*>(p_row0 + ((q ^ sw) << 4)) = w;
            }
            fence_proxy_async_smem();
            named_bar_sync(4 + x, 128);
            if (leader) {
                tma_store_2d(&map_out, sP + (x * 2) * ATC_TILE_BYTES, h * 64, seq * ATC_S + x * 128);
                tma_store_commit();
            }
        }
        if (leader) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, ATC_TMEM_COLS);
}

static int attention_tc_launch(const void* qkv_bf16, void* out_bf16, int n_seq, int heads, float scale,
                               cudaStream_t stream) {
    const int D = heads * 64;
    const uint64_t rows = static_cast<uint64_t>(n_seq) * ATC_S;
    CUtensorMap map128, map16, map_out;
    if (encode_tmap_2d(&map128, TMAP_BF16, qkv_bf16, rows, 3 * D, static_cast<uint64_t>(3) * D * 2, 128, 64)) return -1;
    if (encode_tmap_2d(&map16, TMAP_BF16, qkv_bf16, rows, 3 * D, static_cast<uint64_t>(3) * D * 2, 16, 64)) return -1;
    if (encode_tmap_2d(&map_out, TMAP_BF16, out_bf16, rows, D, static_cast<uint64_t>(D) * 2, 128, 64)) return -1;
    static bool attr_done = false;
    if (!attr_done) {
        if (set_max_dynamic_smem(reinterpret_cast<const void*>(attention_tc_kernel), ATC_SMEM)) return -1;
        attr_done = true;
    }
    const int n_items = n_seq * heads;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM, stream>>>(
        map128, map16, map_out, static_cast<const __nv_bfloat16*>(qkv_bf16), static_cast<__nv_bfloat16*>(out_bf16),
        n_items, heads, scale * 1.4426950408889634f);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// CLS-only attention of the last block for head_dim 64 (forward() returns x[:, 0], vision_transformer.py:252-253): one WARP per
// (sequence, head), no shared memory, no mma.sync.  Lane = key for both products: a lane reads whole 128-byte K / V rows (every
// byte of K and V of the launch is read exactly once: the kernel is HBM-bound), scores and softmax live in registers, the 64
// partial outputs per lane are summed across the warp by recursive halving, which leaves dimensions (2 l, 2 l + 1) on lane l.
// The generic mma.sync kernel ran this row on ONE warp of a 128-thread CTA, 17 HMMA chunks in series (264 us per 1,024 sequences
// under ncu against 62 us of HBM time; legacy HMMA issues once per 50-75 cycles on this part).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_cls64_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                              int n_items, int seq_len, int heads, float scale_log2,
                                                              float* __restrict__ cls_probs) {
    const int item = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (item >= n_items) return;
    const int seq = item / heads, h = item - seq * heads;
    const int D = heads * 64;
    const size_t row_stride = static_cast<size_t>(3) * D;     // elements between consecutive tokens
    const __nv_bfloat16* base = qkv + static_cast<size_t>(seq) * seq_len * row_stride + h * 64;
    constexpr int MAXJ = 9;                                    // keys per lane: seq_len <= 288
    float s[MAXJ];
    {
        f32x2_t q2[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {                          // query row 0: the same 128 bytes for every lane
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(base) + c);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) q2[4 * c + e] = f2_pack(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
        }
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int k = lane + 32 * j;
            s[j] = -INFINITY;
            if (k < seq_len) {
                const uint4* krow = reinterpret_cast<const uint4*>(base + static_cast<size_t>(k) * row_stride + D);
                f32x2_t a0 = f2_pack(0.f, 0.f), a1 = a0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 v = __ldg(krow + c);
                    a0 = f2_fma(q2[4 * c + 0], f2_pack(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u)), a0);
                    a1 = f2_fma(q2[4 * c + 1], f2_pack(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u)), a1);
                    a0 = f2_fma(q2[4 * c + 2], f2_pack(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u)), a0);
                    a1 = f2_fma(q2[4 * c + 3], f2_pack(__uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u)), a1);
                }
                float x0, x1;
                f2_unpack(f2_add(a0, a1), x0, x1);
                s[j] = (x0 + x1) * scale_log2;
            }
        }
    }
    float m = s[0];
#pragma unroll
    for (int j = 1; j < MAXJ; ++j) m = fmaxf(m, s[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) { s[j] = exp2f(s[j] - m); l += s[j]; }     // exp2(-inf) = 0 for the keys past the end
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.0f / l;
    if (cls_probs != nullptr) {
        float* prow = cls_probs + static_cast<size_t>(item) * seq_len;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) { const int k = lane + 32 * j; if (k < seq_len) prow[k] = s[j] * inv; }
    }
    f32x2_t o2[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) o2[e] = f2_pack(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        const int k = lane + 32 * j;
        if (k < seq_len) {
            const uint4* vrow = reinterpret_cast<const uint4*>(base + static_cast<size_t>(k) * row_stride + 2 * D);
            const f32x2_t p2 = f2_pack(s[j], s[j]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = __ldg(vrow + c);
                o2[4 * c + 0] = f2_fma(p2, f2_pack(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u)), o2[4 * c + 0]);
                o2[4 * c + 1] = f2_fma(p2, f2_pack(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u)), o2[4 * c + 1]);
                o2[4 * c + 2] = f2_fma(p2, f2_pack(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u)), o2[4 * c + 2]);
                o2[4 * c + 3] = f2_fma(p2, f2_pack(__uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u)), o2[4 * c + 3]);
            }
        }
    }
    // sum over the 32 lanes: a lane hands over the half of its elements its partner keeps; element e survives on lane e
#pragma unroll
    for (int half = 16, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = lane & bit;
#pragma unroll
        for (int e = 0; e < half; ++e) {
            const f32x2_t keep = up ? o2[e + half] : o2[e], send = up ? o2[e] : o2[e + half];
            float s0, s1;
            f2_unpack(send, s0, s1);
            s0 = __shfl_xor_sync(0xffffffffu, s0, bit);
            s1 = __shfl_xor_sync(0xffffffffu, s1, bit);
            o2[e] = f2_add(keep, f2_pack(s0, s1));
        }
    }
    float x0, x1;
    f2_unpack(o2[0], x0, x1);
    reinterpret_cast<uint32_t*>(out + static_cast<size_t>(seq) * D + h * 64)[lane] = pack_bf16x2(x0 * inv, x1 * inv);
}

static bool attention_force_v1() {          // comparison hook: HB_ATTENTION_V1=1 selects the round-1 tcgen05 kernel
    static int v = -1;
    if (v < 0) { const char* e = getenv("HB_ATTENTION_V1"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int attention_launch(const void* qkv_bf16, void* out_bf16, int n_seq, int seq_len, int heads, int head_dim, float scale,
                     cudaStream_t stream, int cls_only, float* cls_probs) {
    if (n_seq <= 0) return 0;
    if (seq_len <= 0 || heads <= 0) return set_error("hb_attention: bad shape");
    if (cls_probs && !cls_only) return set_error("hb_attention: the CLS-row probabilities come from the CLS-only launch");
    if (seq_len == ATC_S && head_dim == 64 && !attention_force_legacy() && !cls_only)
        return attention_force_v1() ? attention_tc_launch(qkv_bf16, out_bf16, n_seq, heads, scale, stream)
                                    : attention_tc2_launch(qkv_bf16, out_bf16, n_seq, heads, scale, stream);
    if (cls_only && head_dim == 64 && seq_len <= 288 && !attention_force_legacy()) {
        const int n_items = n_seq * heads;
        attention_cls64_kernel<<<(n_items + 3) / 4, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(qkv_bf16),
                                                                      static_cast<__nv_bfloat16*>(out_bf16), n_items, seq_len, heads,
                                                                      scale * 1.4426950408889634f, cls_probs);
        count_launch();
        HB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const int s_pad = (seq_len + 15) & ~15;
    const size_t smem = static_cast<size_t>(3) * s_pad * head_dim * 2;
    if (smem > 200 * 1024) return set_error("hb_attention: seq_len %d too long for the single-pass kernel", seq_len);
    const int q_tiles = s_pad / 16;
    int warps = (q_tiles + 1) / 2;
    if (warps > 9) warps = 9;
    if (warps < 1) warps = 1;
    if (cls_only) warps = 4;                         // all four stage K / V, warp 0 computes the single query tile
    const float scale_log2 = scale * 1.4426950408889634f;
    const unsigned grid = static_cast<unsigned>(n_seq) * heads;
    if (head_dim == 64) {
        auto k = attention_kernel<64>;
        if (set_max_dynamic_smem(reinterpret_cast<const void*>(k), 200 * 1024)) return -1;
        k<<<grid, warps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv_bf16),
                                              static_cast<__nv_bfloat16*>(out_bf16), seq_len, heads, scale_log2, cls_only, cls_probs);
    } else if (head_dim == 32) {
        auto k = attention_kernel<32>;
        if (set_max_dynamic_smem(reinterpret_cast<const void*>(k), 200 * 1024)) return -1;
        k<<<grid, warps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv_bf16),
                                              static_cast<__nv_bfloat16*>(out_bf16), seq_len, heads, scale_log2, cls_only, cls_probs);
    } else {
        return set_error("hb_attention: head_dim %d not supported (64 or 32)", head_dim);
    }
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hb
