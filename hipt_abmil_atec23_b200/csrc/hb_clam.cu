// hb_clam.cu — CLAM_SB gated-attention MIL pooling over ragged bags, all in fp32.
//
// Reference semantics (models/model_clam.py):
//   h1 = relu(fc(h))                                     :83-85   attention_net[0..1]
//   a = tanh(Wa h1 + ba); b = sigmoid(Wb h1 + bb)         :59-61   Attn_Net_Gated
//   A = Wc (a * b) + bc                                   :62-63   -> A_raw [1, N] after the transpose at :150
//   A = softmax(A, dim=1); M = A @ h1                     :154,180
//   logits = classifiers(M); Y_prob = softmax; Y_hat = top1  :181-183
// Dropout layers are identities at inference (model.eval()).
//
// Kernel 1 (scores): one CTA per 128-instance chunk of one bag, one instance per thread.  The feature tile is staged
// through shared memory with coalesced 128-bit loads (row stride padded to 65 words: conflict-free per-thread rows),
// the first Linear is computed 16 output columns at a time against a k-major weight tile read as broadcast float4,
// h1 stays in shared memory for the gate and for the chunk-local softmax partial (max, sum exp, sum exp*h1).
// Several weight sets ("folds") loop inside the CTA so the tile is fetched from HBM once.
// Kernel 2 (combine): one CTA per (bag, model) merges the chunk partials with the usual max-rescaling and applies the
// bag classifier.
#include <math.h>
#include <string.h>

#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

constexpr int CLAM_CHUNK = 128;     // instances per CTA (32 when L1 > 256 so that h1 still fits in shared memory)
__host__ __device__ inline int clam_chunk_for(int L1) { return L1 <= 256 ? CLAM_CHUNK : 32; }
constexpr int CLAM_KC = 64;         // feature columns staged per step
constexpr int CLAM_OB = 16;         // first-layer output columns per pass
constexpr int CLAM_MAX_MODELS = 8;

struct ClamModel { const float* p[10]; };
struct ClamModels { ClamModel m[CLAM_MAX_MODELS]; };

__device__ __forceinline__ float block_reduce_max_128(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    v = red[0];
    for (int i = 1; i < nw; ++i) v = fmaxf(v, red[i]);
    __syncthreads();
    return v;
}
__device__ __forceinline__ float block_reduce_sum_128(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    v = red[0];
    for (int i = 1; i < nw; ++i) v += red[i];
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(CLAM_CHUNK) clam_scores_kernel(const float* __restrict__ feats,
                                                                  const int32_t* __restrict__ bag_offsets,
                                                                  const __grid_constant__ ClamModels models,
                                                                  int n_models, int n_bags, int total_instances, int L0,
                                                                  int L1, int D, int max_chunks,
                                                                  float* __restrict__ a_raw, float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem_clam[];
    const int CH = blockDim.x;                               // instances per CTA (128 or 32)
    float* sX = smem_clam;                                   // [CH][65]
    float* sW = sX + CH * (CLAM_KC + 1);                     // [64][16]
    float* sE = sW + CLAM_KC * CLAM_OB;                      // [CH]
    float* red = sE + CH;                                    // [4]
    float* sH = red + 4;                                     // [CH][L1+1]
    const int ldh = L1 + 1;

    const int bag = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int start = bag_offsets[bag];
    const int len = bag_offsets[bag + 1] - start;
    const int i0 = chunk * CH;
    if (i0 >= len) return;
    const int n_valid = min(CH, len - i0);
    const bool valid = tid < n_valid;
    const float* xbase = feats + static_cast<size_t>(start + i0) * L0;

    for (int mi = 0; mi < n_models; ++mi) {
        const ClamModel& w = models.m[mi];
        const float* W1 = w.p[0]; const float* b1 = w.p[1];
        const float* Wa = w.p[2]; const float* ba = w.p[3];
        const float* Wb = w.p[4]; const float* bb = w.p[5];
        const float* Wc = w.p[6]; const float* bc = w.p[7];

        // ---- h1 = relu(W1 x + b1), 16 output columns per pass
        for (int ob = 0; ob < L1; ob += CLAM_OB) {
            float acc[CLAM_OB];
#pragma unroll
            for (int j = 0; j < CLAM_OB; ++j) acc[j] = 0.f;
            for (int kc = 0; kc < L0; kc += CLAM_KC) {
                // stage X[:, kc:kc+64] (coalesced float4) and W1[ob:ob+16, kc:kc+64]^T
#pragma unroll 4
                for (int it = 0; it < CLAM_KC / 4; ++it) {
                    const int idx = tid + it * CH;
                    const int r = idx >> 4, c4 = idx & 15;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < n_valid) v = __ldg(reinterpret_cast<const float4*>(xbase + static_cast<size_t>(r) * L0 + kc) + c4);
                    float* d = sX + r * (CLAM_KC + 1) + c4 * 4;
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
                for (int idx = tid; idx < CLAM_KC * CLAM_OB; idx += CH) {
                    const int j = idx / CLAM_KC, k = idx - j * CLAM_KC;      // consecutive threads walk k: coalesced
                    const int col = ob + j;
                    sW[k * CLAM_OB + j] = (col < L1) ? __ldg(W1 + static_cast<size_t>(col) * L0 + kc + k) : 0.f;
                }
                __syncthreads();
                const float* xr = sX + tid * (CLAM_KC + 1);
#pragma unroll 8
                for (int k = 0; k < CLAM_KC; ++k) {
                    const float xv = xr[k];
                    const float4* wr = reinterpret_cast<const float4*>(sW + k * CLAM_OB);
#pragma unroll
                    for (int q = 0; q < CLAM_OB / 4; ++q) {
                        const float4 wv = wr[q];
                        acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < CLAM_OB; ++j) {
                const int col = ob + j;
                if (col < L1) sH[tid * ldh + col] = fmaxf(acc[j] + __ldg(b1 + col), 0.f);
            }
        }
        __syncthreads();

        // ---- gated attention score
        float A = __ldg(bc);
        const float* hr = sH + tid * ldh;
        for (int d = 0; d < D; ++d) {
            float a = __ldg(ba + d), b = __ldg(bb + d);
            const float* wa = Wa + static_cast<size_t>(d) * L1;
            const float* wb = Wb + static_cast<size_t>(d) * L1;
            for (int j = 0; j < L1; ++j) {
                const float hj = hr[j];
                a = fmaf(__ldg(wa + j), hj, a);
                b = fmaf(__ldg(wb + j), hj, b);
            }
            A = fmaf(__ldg(Wc + d), tanhf(a) * (1.0f / (1.0f + expf(-b))), A);
        }
        if (valid) a_raw[static_cast<size_t>(mi) * total_instances + start + i0 + tid] = A;

        // ---- chunk-local softmax partial
        const float mx = block_reduce_max_128(valid ? A : -INFINITY, red);
        const float e = valid ? expf(A - mx) : 0.f;
        sE[tid] = e;
        const float sum = block_reduce_sum_128(e, red);     // contains the __syncthreads that publishes sE
        float* out = partials + (static_cast<size_t>(mi * n_bags + bag) * max_chunks + chunk) * (L1 + 2);
        if (tid == 0) { out[0] = mx; out[1] = sum; }
        for (int j = tid; j < L1; j += CH) {
            float acc = 0.f;
            for (int i = 0; i < n_valid; ++i) acc = fmaf(sE[i], sH[i * ldh + j], acc);
            out[2 + j] = acc;
        }
        __syncthreads();       // sH / sE are rewritten by the next model
    }
}

__global__ void __launch_bounds__(128) clam_combine_kernel(const int32_t* __restrict__ bag_offsets,
                                                           const __grid_constant__ ClamModels models, int n_bags, int L1,
                                                           int C, int max_chunks, const float* __restrict__ partials,
                                                           float* __restrict__ m_out, float* __restrict__ logits,
                                                           float* __restrict__ y_prob, long long* __restrict__ y_hat) {
    extern __shared__ float sM[];                             // [L1] + [C]
    float* sL = sM + L1;
    const int bag = blockIdx.x, mi = blockIdx.y, tid = threadIdx.x;
    const int len = bag_offsets[bag + 1] - bag_offsets[bag];
    const int CH = clam_chunk_for(L1);
    const int n_chunks = (len + CH - 1) / CH;
    const float* base = partials + static_cast<size_t>(mi * n_bags + bag) * max_chunks * (L1 + 2);
    float gmax = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) gmax = fmaxf(gmax, base[static_cast<size_t>(c) * (L1 + 2)]);
    float total = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
        const float* pc = base + static_cast<size_t>(c) * (L1 + 2);
        total += pc[1] * expf(pc[0] - gmax);
    }
    const float inv = (n_chunks > 0) ? 1.0f / total : 0.f;
    for (int j = tid; j < L1; j += blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            const float* pc = base + static_cast<size_t>(c) * (L1 + 2);
            acc = fmaf(pc[2 + j], expf(pc[0] - gmax), acc);
        }
        acc *= inv;
        sM[j] = acc;
        if (m_out) m_out[static_cast<size_t>(mi * n_bags + bag) * L1 + j] = acc;
    }
    __syncthreads();
    const float* Wcls = models.m[mi].p[8];
    const float* bcls = models.m[mi].p[9];
    for (int c = tid; c < C; c += blockDim.x) {
        float acc = __ldg(bcls + c);
        for (int j = 0; j < L1; ++j) acc = fmaf(__ldg(Wcls + static_cast<size_t>(c) * L1 + j), sM[j], acc);
        sL[c] = acc;
        if (logits) logits[static_cast<size_t>(mi * n_bags + bag) * C + c] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        float mx = sL[0];
        int arg = 0;
        for (int c = 1; c < C; ++c) if (sL[c] > mx) { mx = sL[c]; arg = c; }
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(sL[c] - mx);
        if (y_prob) for (int c = 0; c < C; ++c) y_prob[static_cast<size_t>(mi * n_bags + bag) * C + c] = expf(sL[c] - mx) / s;
        if (y_hat) y_hat[mi * n_bags + bag] = arg;
    }
}

size_t clam_workspace_bytes(int max_bag_len, int n_bags, int n_models, int L1) {
    const size_t ch = clam_chunk_for(L1);
    const size_t max_chunks = (static_cast<size_t>(max_bag_len) + ch - 1) / ch;
    return static_cast<size_t>(n_models) * n_bags * (max_chunks ? max_chunks : 1) * (L1 + 2) * sizeof(float);
}

int clam_forward_launch(const float* feats, const int32_t* bag_offsets, int n_bags, int total_instances,
                        int max_bag_len, const void* const* weights_host, int n_models, int L0, int L1, int D, int C,
                        float* a_raw, float* m_out, float* logits, float* y_prob, long long* y_hat, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
    if (n_bags <= 0) return 0;
    if (n_models < 1 || n_models > CLAM_MAX_MODELS) return set_error("hb_clam: n_models must be 1..%d", CLAM_MAX_MODELS);
    if (L0 % CLAM_KC != 0) return set_error("hb_clam: L0=%d must be a multiple of %d", L0, CLAM_KC);
    if (L1 < 1 || D < 1 || C < 1 || C > 64) return set_error("hb_clam: bad dims L1=%d D=%d C=%d", L1, D, C);
    if (!bag_offsets || !weights_host || !a_raw || !workspace) return set_error("hb_clam: null argument");
    if (total_instances > 0 && !feats) return set_error("hb_clam: null features");
    if ((reinterpret_cast<uintptr_t>(feats) & 15) != 0) return set_error("hb_clam: features must be 16 B aligned");
    const size_t need = clam_workspace_bytes(max_bag_len, n_bags, n_models, L1);
    if (workspace_bytes < need) return set_error("hb_clam: workspace %zu < %zu bytes", workspace_bytes, need);
    ClamModels models;
    memset(&models, 0, sizeof(models));
    for (int m = 0; m < n_models; ++m)
        for (int k = 0; k < 10; ++k) {
            models.m[m].p[k] = static_cast<const float*>(weights_host[m * 10 + k]);
            if (!models.m[m].p[k]) return set_error("hb_clam: weight pointer %d of model %d is null", k, m);
        }
    const int CH = clam_chunk_for(L1);
    const int max_chunks = (max_bag_len + CH - 1) / CH;
    float* partials = static_cast<float*>(workspace);
    if (max_chunks > 0) {
        const size_t smem = (static_cast<size_t>(CH) * (CLAM_KC + 1) + CLAM_KC * CLAM_OB + CH + 4 +
                             static_cast<size_t>(CH) * (L1 + 1)) * sizeof(float);
        if (smem > 220 * 1024) return set_error("hb_clam: L1=%d too large for the fused kernel", L1);
        HB_CUDA_OK(cudaFuncSetAttribute(clam_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        dim3 grid(max_chunks, n_bags);
        ProfScope ps(10, stream);
        clam_scores_kernel<<<grid, CH, smem, stream>>>(feats, bag_offsets, models, n_models, n_bags,
                                                                total_instances, L0, L1, D, max_chunks, a_raw, partials);
        count_launch();
    HB_CUDA_OK(cudaGetLastError());
    }
    dim3 grid2(n_bags, n_models);
    ProfScope ps2(11, stream);
    clam_combine_kernel<<<grid2, 128, (L1 + C) * sizeof(float), stream>>>(bag_offsets, models, n_bags, L1, C,
                                                                           max_chunks > 0 ? max_chunks : 1, partials,
                                                                           m_out, logits, y_prob, y_hat);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hb

extern "C" {
size_t hb_clam_workspace_bytes(int max_bag_len, int n_bags, int n_models, int L1) {
    return hb::clam_workspace_bytes(max_bag_len, n_bags, n_models, L1);
}
int hb_clam_sb_forward(const float* feats, const int32_t* bag_offsets, int n_bags, int total_instances,
                       int max_bag_len, const void* const* weights_host, int n_models, int L0, int L1, int D, int C,
                       float* a_raw, float* m_out, float* logits, float* y_prob, int64_t* y_hat, void* workspace,
                       size_t workspace_bytes, void* stream) {
    return hb::clam_forward_launch(feats, bag_offsets, n_bags, total_instances, max_bag_len, weights_host, n_models, L0,
                                   L1, D, C, a_raw, m_out, logits, y_prob, reinterpret_cast<long long*>(y_hat),
                                   workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
}
