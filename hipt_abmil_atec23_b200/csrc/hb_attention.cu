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
#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

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
                                                        float scale_log2) {
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
        if (r < seq_len) {
            cp_async_16(dst, base + static_cast<size_t>(r) * 3 * D + which * D + c * 8);
        } else {
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int q_tiles = s_pad / 16;
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

int attention_launch(const void* qkv_bf16, void* out_bf16, int n_seq, int seq_len, int heads, int head_dim, float scale,
                     cudaStream_t stream) {
    if (n_seq <= 0) return 0;
    if (seq_len <= 0 || heads <= 0) return set_error("hb_attention: bad shape");
    const int s_pad = (seq_len + 15) & ~15;
    const size_t smem = static_cast<size_t>(3) * s_pad * head_dim * 2;
    if (smem > 200 * 1024) return set_error("hb_attention: seq_len %d too long for the single-pass kernel", seq_len);
    const int q_tiles = s_pad / 16;
    int warps = (q_tiles + 1) / 2;
    if (warps > 9) warps = 9;
    if (warps < 1) warps = 1;
    const float scale_log2 = scale * 1.4426950408889634f;
    const unsigned grid = static_cast<unsigned>(n_seq) * heads;
    if (head_dim == 64) {
        auto k = attention_kernel<64>;
        HB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        k<<<grid, warps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv_bf16),
                                              static_cast<__nv_bfloat16*>(out_bf16), seq_len, heads, scale_log2);
    } else if (head_dim == 32) {
        auto k = attention_kernel<32>;
        HB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        k<<<grid, warps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv_bf16),
                                              static_cast<__nv_bfloat16*>(out_bf16), seq_len, heads, scale_log2);
    } else {
        return set_error("hb_attention: head_dim %d not supported (64 or 32)", head_dim);
    }
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hb
