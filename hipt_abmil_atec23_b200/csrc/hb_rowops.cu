// hb_rowops.cu — bandwidth-bound row kernels around the GEMMs:
//   * LayerNorm (fp32 in, fp32 statistics, bf16 and/or fp32 out), one warp (dim 384) or half-warp (dim 192) per row,
//     128-bit loads, shuffle reductions.  Replaces nn.LayerNorm(eps=1e-6) at HIPT_4K/vision_transformer.py:138,142,195
//     and vision_transformer4k.py:140-158,183.
//   * im2col of a region into the patch-embed GEMM operand: unfold(2,256,256).unfold(3,256,256) + rearrange
//     (HIPT_4K/hipt_4k.py:64-65) composed with the 16x16/16 convolution's receptive fields
//     (vision_transformer.py:165-169); K order is (c, i, j) = weight.reshape(384, 768).
//   * CLS rows: x[seq, 0, :] = cls_token + pos[0]  (vision_transformer.py:240-244).
#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

// ------------------------------------------------------------------------------------------------ LayerNorm
template <int DIM, bool IN_BF16>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x_any, size_t x_row_stride,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ out_bf16,
                                                        float* __restrict__ out_f32, int rows) {
    constexpr int LANES = DIM / 12;              // lanes per row, 3 float4 per lane
    constexpr int ROWS_PER_WARP = 32 / LANES;
    static_assert(LANES == 32 || LANES == 16, "dim must be 384 or 192");
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES;                // which row of the warp's group
    const int l = lane % LANES;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;

    float4 g[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * LANES + l);
        b[i] = __ldg(reinterpret_cast<const float4*>(beta) + i * LANES + l);
    }
    for (int r0 = warp_global * ROWS_PER_WARP; r0 < rows; r0 += n_warps * ROWS_PER_WARP) {
        const int row = r0 + sub;
        const bool valid = row < rows;
        float4 v[3];
        if constexpr (IN_BF16) {
            const uint2* xr = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x_any) +
                                                             static_cast<size_t>(valid ? row : 0) * x_row_stride);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint2 w = xr[i * LANES + l];
                v[i] = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u),
                                   __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u));
            }
        } else {
            const float4* xr = reinterpret_cast<const float4*>(static_cast<const float*>(x_any) +
                                                               static_cast<size_t>(valid ? row : 0) * x_row_stride);
#pragma unroll
            for (int i = 0; i < 3; ++i) v[i] = xr[i * LANES + l];
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / DIM);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / DIM) + eps);
        if (valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float4 y;
                y.x = v[i].x * rstd * g[i].x + b[i].x;
                y.y = v[i].y * rstd * g[i].y + b[i].y;
                y.z = v[i].z * rstd * g[i].z + b[i].z;
                y.w = v[i].w * rstd * g[i].w + b[i].w;
                const size_t off = static_cast<size_t>(row) * DIM + (i * LANES + l) * 4;
                if (out_bf16) {
                    uint2 pk;
                    pk.x = pack_bf16x2(y.x, y.y);
                    pk.y = pack_bf16x2(y.z, y.w);
                    *reinterpret_cast<uint2*>(out_bf16 + off) = pk;
                }
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = y;
            }
        }
    }
}

int layernorm_launch(const void* x, int x_is_bf16, size_t x_row_stride, const float* gamma, const float* beta, float eps,
                     void* out_bf16, float* out_f32, int rows, int dim, cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (x_row_stride % 4 != 0) return set_error("hb_layernorm: row stride %zu must be a multiple of 4", x_row_stride);
    const int rows_per_block = (dim == 384) ? 8 : 16;
    int blocks = (rows + rows_per_block - 1) / rows_per_block;
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out_bf16);
    if (dim == 384 && x_is_bf16) layernorm_kernel<384, true><<<blocks, 256, 0, stream>>>(x, x_row_stride, gamma, beta, eps, ob, out_f32, rows);
    else if (dim == 384) layernorm_kernel<384, false><<<blocks, 256, 0, stream>>>(x, x_row_stride, gamma, beta, eps, ob, out_f32, rows);
    else if (dim == 192 && x_is_bf16) layernorm_kernel<192, true><<<blocks, 256, 0, stream>>>(x, x_row_stride, gamma, beta, eps, ob, out_f32, rows);
    else if (dim == 192) layernorm_kernel<192, false><<<blocks, 256, 0, stream>>>(x, x_row_stride, gamma, beta, eps, ob, out_f32, rows);
    else
        return set_error("hb_layernorm: dim %d not supported (384 or 192)", dim);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ im2col
// One thread moves one 16-pixel run (fixed c, image row, token column) = 32 B of bf16 output.
template <bool F32>
__global__ void __launch_bounds__(256) im2col_kernel(const uint8_t* __restrict__ img, size_t patch_stride,
                                                     size_t chan_stride, size_t row_pitch, int grid_cols,
                                                     int patch_begin, int n_patches, __nv_bfloat16* __restrict__ a,
                                                     int vec_ok) {
    const size_t total = static_cast<size_t>(n_patches) * 3 * 256 * 16;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int tx = idx & 15;
        const int yrow = (idx >> 4) & 255;          // row inside the 256x256 patch
        const int c = (idx >> 12) % 3;
        const int pl = static_cast<int>(idx / (12288));
        const int p = patch_begin + pl;
        // grid mode (grid_cols > 0): patch p is tile (p / grid_cols, p % grid_cols) of one region image;
        // batch mode (grid_cols == 0): patch p is a separate 256x256 image at p * patch_stride.
        const int p1 = grid_cols > 0 ? p / grid_cols : 0;
        const int p2 = grid_cols > 0 ? p - p1 * grid_cols : 0;
        const int ty = yrow >> 4, i = yrow & 15;
        const size_t src = (grid_cols > 0 ? 0 : static_cast<size_t>(p) * patch_stride) +
                           static_cast<size_t>(c) * chan_stride + static_cast<size_t>(p1 * 256 + yrow) * row_pitch +
                           static_cast<size_t>(p2 * 256 + tx * 16);
        uint32_t pk[8];
        if constexpr (F32) {
            const float* s = reinterpret_cast<const float*>(img) + src;
            if (vec_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 f = *reinterpret_cast<const float4*>(s + 4 * q);
                    pk[2 * q] = pack_bf16x2(f.x, f.y);
                    pk[2 * q + 1] = pack_bf16x2(f.z, f.w);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) pk[q] = pack_bf16x2(s[2 * q], s[2 * q + 1]);
            }
        } else {
            const uint8_t* s = img + src;
            uint32_t w[4];
            if (vec_ok) {
                const uint4 u = *reinterpret_cast<const uint4*>(s);
                w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    w[q] = s[4 * q] | (uint32_t(s[4 * q + 1]) << 8) | (uint32_t(s[4 * q + 2]) << 16) | (uint32_t(s[4 * q + 3]) << 24);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {          // 0..255 is exact in bf16
                pk[2 * q] = pack_bf16x2(static_cast<float>(w[q] & 0xff), static_cast<float>((w[q] >> 8) & 0xff));
                pk[2 * q + 1] = pack_bf16x2(static_cast<float>((w[q] >> 16) & 0xff), static_cast<float>(w[q] >> 24));
            }
        }
        const size_t row = static_cast<size_t>(pl) * 256 + ty * 16 + tx;
        uint4* dst = reinterpret_cast<uint4*>(a + row * 768 + c * 256 + i * 16);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}

int im2col_launch(const void* image, int image_is_f32, size_t patch_stride, size_t chan_stride, size_t row_pitch,
                  int grid_cols, int patch_begin, int n_patches, void* a_bf16, cudaStream_t stream) {
    if (n_patches <= 0) return 0;
    const size_t total = static_cast<size_t>(n_patches) * 12288;
    size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    const size_t esz = image_is_f32 ? 4 : 1;
    const int vec_ok = (reinterpret_cast<uintptr_t>(image) % 16 == 0) && ((chan_stride * esz) % 16 == 0) &&
                       ((row_pitch * esz) % 16 == 0) && ((patch_stride * esz) % 16 == 0);
    if (image_is_f32)
        im2col_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
            static_cast<const uint8_t*>(image), patch_stride, chan_stride, row_pitch, grid_cols, patch_begin, n_patches,
            static_cast<__nv_bfloat16*>(a_bf16), vec_ok);
    else
        im2col_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
            static_cast<const uint8_t*>(image), patch_stride, chan_stride, row_pitch, grid_cols, patch_begin, n_patches,
            static_cast<__nv_bfloat16*>(a_bf16), vec_ok);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ CLS rows
// One warp per sequence: x[seq, 0, :] = cls_token + pos[0], plus the bf16 copy and the row's (sum, sum of squares)
// that the LayerNorm-folded GEMM epilogues consume.
__global__ void cls_rows_kernel(const float* __restrict__ cls_token, const float* __restrict__ pos_table,
                                float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float* __restrict__ stats,
                                int n_seq, int seq_len, int dim) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_seq) return;
    const size_t row = static_cast<size_t>(warp) * seq_len;
    float s = 0.f, q = 0.f;
    for (int d = lane; d < dim; d += 32) {
        const float v = cls_token[d] + pos_table[d];
        if (x) x[row * dim + d] = v;
        if (xb) xb[row * dim + d] = __float2bfloat16(v);
        s += v;
        q = fmaf(v, v, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (stats && lane == 0) { stats[row * 2] = s; stats[row * 2 + 1] = q; }
}

int cls_rows_launch(const float* cls_token, const float* pos_table, float* x, void* xb_bf16, float* stats, int n_seq,
                    int seq_len, int dim, cudaStream_t stream) {
    if (n_seq <= 0) return 0;
    const int blocks = (n_seq * 32 + 255) / 256;
    cls_rows_kernel<<<blocks, 256, 0, stream>>>(cls_token, pos_table, x, static_cast<__nv_bfloat16*>(xb_bf16), stats,
                                                n_seq, seq_len, dim);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hb
