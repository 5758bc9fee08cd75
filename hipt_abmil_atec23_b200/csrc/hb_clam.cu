// hb_clam.cu — CLAM_SB gated-attention MIL pooling over ragged bags (forward, training-step backward, Adam), fp32 accuracy.
//
// Reference semantics (models/model_clam.py):
//   h1 = relu(fc(h))                                     :83-85   attention_net[0..1]
//   a = tanh(Wa h1 + ba); b = sigmoid(Wb h1 + bb)         :59-61   Attn_Net_Gated
//   A = Wc (a * b) + bc                                   :62-63   -> A_raw [1, N] after the transpose at :150
//   A = softmax(A, dim=1); M = A @ h1                     :154,180
//   logits = classifiers(M); Y_prob = softmax; Y_hat = top1  :181-183
// Dropout layers are identities at inference (model.eval()).
//
// Forward = work table -> score kernel -> combine:
//   clam_work_table_kernel   ragged bags -> flat (bag, chunk) list, so only CTAs with work are launched
//   clam_scores_tc_kernel    192-d features, L1 = 16 / 32 / 64, n_models * L1 <= 80: the first Linear of every fold as ONE
//                            tcgen05 kind::tf32 GEMM per 128-instance tile and both gate Linears as a second one (A = h1 in
//                            tensor memory), hi / lo operand splitting (fp32-level accuracy), persistent CTAs, TMA ring;
//                            activations, score and chunk softmax / pooling partials in the epilogue warpgroups
//   clam_scores192_kernel    192-d features, any L1 <= 128 (the large heads, tiny inputs, > 5 folds): one CTA per
//                            64-instance chunk, tile staged once with cp.async and reused by every fold, first Linear and gate
//                            as register-tiled SGEMMs on packed f32x2 FMAs (packed along K: operands are natural float4 halves)
//   clam_scores_kernel       any other feature size (e.g. the 1024-d demo checkpoint): one instance per thread
//   clam_combine_kernel      one CTA per (bag, model): merges the chunk partials (max-rescaling), classifier, softmax, argmax
// Algorithmic HBM bytes: 772 per instance forward (features read once for all folds + one score), 1,544 with the
// recomputing backward (clam_bwd_prep_kernel + clam_bwd192_kernel); adam_step_kernel updates every tensor in one launch.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "hb_ptx.cuh"
#include "hb_internal.h"

namespace hb {

constexpr int CLAM_CHUNK = 128;     // generic kernel: instances per CTA (32 when L1 > 256 so that h1 still fits in shared memory)
constexpr int CL_CH = 64;           // 192-d kernel: instances per CTA
constexpr int CL_THREADS = 128;     // 256 when L1 >= 64 (one CTA per SM then: more warps to hide latency)
constexpr int CL_XS = 196;          // feature row stride in floats: 16 B aligned and conflict-free for 8-lane float4 wavefronts
__host__ __device__ inline int clam_kc_for(int L1) { return L1 <= 32 ? 192 : 64; }   // columns of W1 staged per step
// floats of the W1 staging buffer [L1][KC + 4]; the backward reuses it for dz [64][L1 + 4]
__host__ __device__ inline int clam_sw_floats(int L1) {
    const int a = (clam_kc_for(L1) + 4) * L1, b = CL_CH * (L1 + 4);
    return a > b ? a : b;
}
__host__ __device__ inline bool clam_is192(int L0, int L1, int D) {
    return L0 == 192 && L1 <= 128 && (L1 % 8) == 0 && (D % 4) == 0;
}
__host__ __device__ inline int clam_chunk_for(int L0, int L1, int D) {
    return clam_is192(L0, L1, D) ? CL_CH : (L1 <= 256 ? CLAM_CHUNK : 32);
}
constexpr int CLAM_KC = 64;         // feature columns staged per step
constexpr int CLAM_OB = 16;         // first-layer output columns per pass
constexpr int CLAM_MAX_MODELS = 8;

struct ClamModel { const float* p[10]; };
struct ClamModels { ClamModel m[CLAM_MAX_MODELS]; };

// Training-mode dropout (nn.Dropout after the ReLU, model_clam.py:84-85, and inside both gate branches, :50-52).  The keep
// decision of (instance, unit) is a pure function of a 64-bit seed — a counter-based hash, nothing is stored — so the
// recomputing backward sees exactly the forward's masks.  Units: [0, L1) the ReLU outputs, [L1, L1 + D) branch a,
// [L1 + D, L1 + 2 D) branch b.  keep <=> hash >= thresh, thresh = round(p * 2^32) (p = 1 drops everything, like torch).
struct ClamDrop { uint32_t seed_lo, seed_hi, thresh_lo; int all; float scale; };     // scale = 1 / (1 - p); all: p >= 1
__host__ __device__ __forceinline__ uint32_t clam_mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ float clam_keep(const ClamDrop& dr, uint32_t instance, uint32_t unit) {
    if (dr.thresh_lo == 0u && !dr.all) return 1.0f;                              // dropout off
    if (dr.all) return 0.0f;
    const uint32_t h = clam_mix32(clam_mix32(instance * 0x9E3779B1u + dr.seed_lo) ^ (unit * 0x7FEB352Du + dr.seed_hi));
    return h >= dr.thresh_lo ? dr.scale : 0.0f;
}
// one dropout state per model: the paired (multi-trial training) launch gives every trial its own seed
struct ClamDrops { ClamDrop d[8]; };
static ClamDrop clam_drop_make(float p, unsigned long long seed) {
    ClamDrop d = {};
    if (p <= 0.f) return d;
    d.seed_lo = static_cast<uint32_t>(seed); d.seed_hi = static_cast<uint32_t>(seed >> 32);
    if (p >= 1.f) { d.all = 1; return d; }
    double t = static_cast<double>(p) * 4294967296.0 + 0.5;
    if (t < 1.0) t = 1.0;
    if (t > 4294967295.0) t = 4294967295.0;
    d.thresh_lo = static_cast<uint32_t>(t);
    d.scale = 1.0f / (1.0f - p);
    return d;
}

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

// Work table: ragged bags -> flat list of (bag, chunk) items so that the score kernels launch exactly the CTAs that have
// work (a 2-D grid of n_bags x max_chunks is mostly empty CTAs when bag sizes span 50..20,000).  prefix[b] = number of
// chunks before bag b; work[w] = bag << 12-free pair stored as two ints.  One CTA, block-wide scan over the bags.
__global__ void __launch_bounds__(1024) clam_work_table_kernel(const int32_t* __restrict__ bag_offsets, int n_bags, int CH,
                                                               int32_t* __restrict__ prefix, int32_t* __restrict__ work,
                                                               int work_cap) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    __shared__ int s_excl[1024];                              // first work item of each bag of the current block of bags
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();                                  // the score kernel may stage its weights while this table is built
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_bags; b0 += 1024) {
        const int b = b0 + tid;
        const int len = (b < n_bags) ? bag_offsets[b + 1] - bag_offsets[b] : 0;
        const int cnt = (len + CH - 1) / CH;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + inc - cnt;
        s_excl[tid] = excl;
        if (b < n_bags) prefix[b] = excl;
        __syncthreads();
        // the block's items are filled by all threads (a thread per item finds its bag by bisection: one thread per BAG writing
        // its chunks serially costs 157 dependent stores for a 20,000-instance bag and was 11 us of a 150 us forward)
        const int nb = min(1024, n_bags - b0), end = min(carry + warp_tot[31], work_cap);
        for (int w = carry + tid; w < end; w += 1024) {
            int lo = 0, hi = nb;                              // last bag with s_excl <= w (empty bags share their successor's)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_excl[mid] <= w) lo = mid; else hi = mid;
            }
            *reinterpret_cast<int2*>(work + 2 * w) = make_int2(b0 + lo, w - s_excl[lo]);
        }
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) prefix[n_bags] = carry_s;
}

__global__ void __launch_bounds__(CLAM_CHUNK) clam_scores_kernel(const float* __restrict__ feats,
                                                                  const int32_t* __restrict__ bag_offsets,
                                                                  const __grid_constant__ ClamModels models,
                                                                  int n_models, int n_bags, int total_instances, int L0,
                                                                  int L1, int D, const int32_t* __restrict__ prefix,
                                                                  const int32_t* __restrict__ work, int work_cap,
                                                                  float* __restrict__ a_raw, float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem_clam[];
    const int CH = blockDim.x;                               // instances per CTA (128 or 32)
    float* sX = smem_clam;                                   // [CH][65]
    float* sW = sX + CH * (CLAM_KC + 1);                     // [64][16]
    float* sE = sW + CLAM_KC * CLAM_OB;                      // [CH]
    float* red = sE + CH;                                    // [4]
    float* sH = red + 4;                                     // [CH][L1+1]
    const int ldh = L1 + 1;

    const int tid = threadIdx.x;
    const int wi = blockIdx.x;
    if (wi >= prefix[n_bags]) return;
    const int bag = work[2 * wi], chunk = work[2 * wi + 1];
    const int start = bag_offsets[bag];
    const int len = bag_offsets[bag + 1] - start;
    const int i0 = chunk * CH;
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
        float* out = partials + (static_cast<size_t>(mi) * work_cap + wi) * (L1 + 2);
        if (tid == 0) { out[0] = mx; out[1] = sum; }
        for (int j = tid; j < L1; j += CH) {
            float acc = 0.f;
            for (int i = 0; i < n_valid; ++i) acc = fmaf(sE[i], sH[i * ldh + j], acc);
            out[2 + j] = acc;
        }
        __syncthreads();       // sH / sE are rewritten by the next model
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// 192-d kernel
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lds_2f2(uint32_t addr, f32x2_t& a, f32x2_t& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

// h1 tile.  FMAs are packed along K: acc (lo, hi) += (x[k], x[k+1]) * (w[k], w[k+1]), so both operands are natural
// 64-bit pairs of a float4 load (x row of the instance, W1 row of the output column in nn.Linear's own [L1][192]
// layout) and no value is ever duplicated into a pair; h1 = relu(lo + hi + bias).  Thread tile = 2 instances (lane,
// lane + 32) x TN columns; warp w owns column groups w, w + nw, ...; W1 reads are warp-uniform broadcasts.
template <int TN, int L1T>
__device__ __forceinline__ void clam_fc1_192(const float* __restrict__ W1, const float* b1, int L1_rt,
                                             const float* sX, float* sW, float* sH, int ldh, bool two_halves,
                                             const ClamDrop& dr, uint32_t inst_base) {
    const int L1 = L1T ? L1T : L1_rt;                        // L1T != 0: every stride below is a compile-time constant
    constexpr int GPW = (TN == 8) ? 2 : 1;                   // column groups per warp: L1 <= 16 (TN 4, 4 warps), L1 = 32 (TN 8,
                                                             // 4 warps), L1 = 64 / 128 (TN 8, 8 warps)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x, nw = nthreads >> 5;
    const int n_groups = L1 / TN;
    const int KC = clam_kc_for(L1), ldw = KC + 4;
    f32x2_t acc[GPW][2][TN];
#pragma unroll
    for (int g = 0; g < GPW; ++g)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[g][i][c] = f2_pack(0.f, 0.f);
    // 32-instance tiles (backward of the large heads): the second instance of the thread tile repeats the first
    // plain shared-memory loads (not asm): the compiler may hoist and batch them across the unrolled iterations
    const ulonglong2* x0p = reinterpret_cast<const ulonglong2*>(sX + lane * CL_XS);
    const ulonglong2* x1p = two_halves ? reinterpret_cast<const ulonglong2*>(sX + (lane + 32) * CL_XS) : x0p;
    for (int kc = 0; kc < 192; kc += KC) {
        // stage W1[:, kc:kc+KC] as sW[col][k] (row stride KC + 4): straight 16-byte copies
        {
            const int per_row = KC / 4;
            const uint32_t dst0 = smem_u32(sW);
            for (int idx = tid; idx < L1 * per_row; idx += nthreads) {
                const int col = idx / per_row, c4 = idx - col * per_row;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst0 + (col * ldw + c4 * 4) * 4),
                             "l"(W1 + static_cast<size_t>(col) * 192 + kc + c4 * 4) : "memory");
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
            const int grp = warp + nw * g;
            if (grp < n_groups) {
                const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(sW + grp * TN * ldw);
#pragma unroll 4
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    const ulonglong2 xa = x0p[kc / 4 + k4], xb = x1p[kc / 4 + k4];
#pragma unroll
                    for (int c = 0; c < TN; ++c) {
                        const ulonglong2 wv = wp[c * (ldw / 4) + k4];
                        acc[g][0][c] = f2_fma(xa.x, wv.x, acc[g][0][c]);
                        acc[g][1][c] = f2_fma(xb.x, wv.x, acc[g][1][c]);
                        acc[g][0][c] = f2_fma(xa.y, wv.y, acc[g][0][c]);
                        acc[g][1][c] = f2_fma(xb.y, wv.y, acc[g][1][c]);
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int g = 0; g < GPW; ++g) {
        const int grp = warp + nw * g;
        if (grp < n_groups) {
#pragma unroll
            for (int c = 0; c < TN; ++c) {
                const int col = grp * TN + c;
                const float bias = b1[col];                  // shared memory (staged with the fold's other small vectors)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i == 1 && !two_halves) break;
                    float v0, v1;
                    f2_unpack(acc[g][i][c], v0, v1);
                    sH[(lane + 32 * i) * ldh + col] = fmaxf(v0 + v1 + bias, 0.f) * clam_keep(dr, inst_base + lane + 32 * i, col);
                }
            }
        }
    }
}

template <int TN, int L1T>
__global__ void __launch_bounds__(256) clam_scores192_kernel(const float* __restrict__ feats,
                                                                     const int32_t* __restrict__ bag_offsets,
                                                                     const __grid_constant__ ClamModels models,
                                                                     int n_models, int n_bags, int total_instances, int L1_rt,
                                                                     int D_rt, const int32_t* __restrict__ prefix,
                                                                     const int32_t* __restrict__ work, int work_cap,
                                                                     float* __restrict__ a_raw, float* __restrict__ partials,
                                                                     const __grid_constant__ ClamDrops drs, int paired) {
    // paired: model m is evaluated on bag m only (n_bags == n_models) and a_raw is written without the model stride —
    // the multi-trial training step, where every trial has its own weights, bag and dropout seed
    const int L1 = L1T ? L1T : L1_rt, D = L1T ? L1T / 2 : D_rt;    // the HIPT heads have D = L1 / 2 (model_clam.py:81)
    extern __shared__ __align__(16) float smem_clam[];
    const int ldh = L1 + 4;
    float* sX = smem_clam;                                   // [64][196]
    float* sW = sX + CL_CH * CL_XS;                          // [L1][KC + 4]  W1 slice
    float* sH = sW + clam_sw_floats(L1);                     // [64][L1 + 4]
    float* sG = sH + CL_CH * ldh;                            // [2][D][L1]    Wa, Wb
    float* sA = sG + 2 * D * L1;                             // [max(4, D/4)][64]  per-task score partials (reused: chunk sums)
    float* sE = sA + (D / 4 > 4 ? D / 4 : 4) * CL_CH;        // [64]
    float* red = sE + CL_CH;                                 // [8]
    float* sV = red + 8;                                     // b1 [L1] | ba [D] | bb [D] | Wc [D] | bc [1]
    const int nthreads = blockDim.x;

    const int tid = threadIdx.x;
    const int wi = blockIdx.x;
    if (wi >= prefix[n_bags]) return;
    const int bag = work[2 * wi], chunk = work[2 * wi + 1];
    const int start = bag_offsets[bag];
    const int len = bag_offsets[bag + 1] - start;
    const int i0 = chunk * CL_CH;
    const int n_valid = min(CL_CH, len - i0);

    // ---- the feature tile: HBM -> shared memory, once
    {                                                        // cp.async: every 16-byte piece in flight at once, no registers
        const float4* src = reinterpret_cast<const float4*>(feats + static_cast<size_t>(start + i0) * 192);
        const uint32_t dst0 = smem_u32(sX);
        for (int idx = tid; idx < CL_CH * 48; idx += nthreads) {
            const int r = idx / 48, c4 = idx - r * 48;
            const uint32_t dst = dst0 + (r * CL_XS + c4 * 4) * 4;
            const int nbytes = (r < n_valid) ? 16 : 0;       // rows past the bag end are zero-filled
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + (r < n_valid ? idx : 0)), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();

    const int inst = tid & 63, half = tid >> 6;
    const bool valid = half == 0 && inst < n_valid;
    for (int mi = 0; mi < n_models; ++mi) {
        if (paired && mi != bag) continue;
        const ClamModel& w = models.m[mi];
        const ClamDrop& dr = drs.d[paired ? mi : 0];
        // gate weights and every small vector of this fold: issued before the first Linear so that their latency hides
        // under it (the staging of W1 inside clam_fc1_192 waits on the same cp.async group)
        for (int idx = tid; idx < D * L1 / 4; idx += nthreads) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sG + idx * 4)), "l"(w.p[2] + idx * 4) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sG + D * L1 + idx * 4)), "l"(w.p[4] + idx * 4) : "memory");
        }
        for (int idx = tid; idx < L1 + 3 * D + 1; idx += nthreads) {
            const float* src = idx < L1 ? w.p[1] + idx : idx < L1 + D ? w.p[3] + (idx - L1) : idx < L1 + 2 * D ? w.p[5] + (idx - L1 - D)
                             : idx < L1 + 3 * D ? w.p[6] + (idx - L1 - 2 * D) : w.p[7];
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sV + idx)), "l"(src) : "memory");
        }
        clam_fc1_192<TN, L1T>(w.p[0], sV, L1, sX, sW, sH, ldh, true, dr, static_cast<uint32_t>(i0));   // dropout: instance index inside its bag
        __syncthreads();

        // ---- gated attention score, register-tiled like the first Linear: warp task = 4 gate units (their Wa and Wb rows)
        // x the 64 instances (lane, lane + 32), FMAs packed along L1 on natural float4 halves; every task leaves its share
        // of the score, sum_d Wc[d] tanh(a_d) sigmoid(b_d), in sA[task][instance]
        const int n_gt = D / 4;
        {
            const float* ba = sV + L1; const float* bb = ba + D; const float* Wc = bb + D;
            const int warp = tid >> 5, lane = tid & 31, nw = nthreads >> 5;
            const ulonglong2* h0p = reinterpret_cast<const ulonglong2*>(sH + lane * ldh);
            const ulonglong2* h1p = reinterpret_cast<const ulonglong2*>(sH + (lane + 32) * ldh);
            for (int task = warp; task < n_gt; task += nw) {
                f32x2_t acc[2][8];                           // [instance][a0..a3, b0..b3]
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[i][u] = f2_pack(0.f, 0.f);
                const ulonglong2* wap = reinterpret_cast<const ulonglong2*>(sG + (task * 4) * L1);
                const ulonglong2* wbp = reinterpret_cast<const ulonglong2*>(sG + (D + task * 4) * L1);
#pragma unroll 2
                for (int j4 = 0; j4 < L1 / 4; ++j4) {
                    const ulonglong2 ha = h0p[j4], hb = h1p[j4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const ulonglong2 wa = wap[u * (L1 / 4) + j4], wb = wbp[u * (L1 / 4) + j4];
                        acc[0][u] = f2_fma(ha.x, wa.x, acc[0][u]);         acc[1][u] = f2_fma(hb.x, wa.x, acc[1][u]);
                        acc[0][u] = f2_fma(ha.y, wa.y, acc[0][u]);         acc[1][u] = f2_fma(hb.y, wa.y, acc[1][u]);
                        acc[0][4 + u] = f2_fma(ha.x, wb.x, acc[0][4 + u]); acc[1][4 + u] = f2_fma(hb.x, wb.x, acc[1][4 + u]);
                        acc[0][4 + u] = f2_fma(ha.y, wb.y, acc[0][4 + u]); acc[1][4 + u] = f2_fma(hb.y, wb.y, acc[1][4 + u]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float A = 0.f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int d = task * 4 + u;
                        float a0, a1, b0, b1v;
                        f2_unpack(acc[i][u], a0, a1);
                        f2_unpack(acc[i][4 + u], b0, b1v);
                        const uint32_t gi = static_cast<uint32_t>(i0 + lane + 32 * i);
                        const float ka = clam_keep(dr, gi, L1 + d), kb = clam_keep(dr, gi, L1 + D + d);
                        A = fmaf(Wc[d], (ka * tanhf(a0 + a1 + ba[d])) * (kb / (1.0f + expf(-(b0 + b1v + bb[d])))), A);
                    }
                    sA[task * CL_CH + lane + 32 * i] = A;
                }
            }
        }
        __syncthreads();
        float A = sV[L1 + 3 * D];
        for (int q = 0; q < n_gt; ++q) A += sA[q * CL_CH + inst];
        if (valid) a_raw[(paired ? 0 : static_cast<size_t>(mi) * total_instances) + start + i0 + inst] = A;

        // ---- chunk-local softmax partial
        const float mx = block_reduce_max_128(valid ? A : -INFINITY, red);
        const float e = valid ? expf(A - mx) : 0.f;
        if (half == 0) sE[inst] = e;
        const float sum = block_reduce_sum_128(e, red);     // contains the __syncthreads that publishes sE
        float* out = partials + (static_cast<size_t>(mi) * work_cap + wi) * (L1 + 2);
        if (tid == 0) { out[0] = mx; out[1] = sum; }
        {                                                    // sum_i e_i h1[i][:]: instance groups in parallel, then a small reduce
            const int G = nthreads / L1;                     // >= 1 (L1 <= 128 <= nthreads)
            const int ig = tid / L1, j = tid - ig * L1;
            float acc = 0.f;
            if (ig < G)
                for (int i = ig; i < n_valid; i += G) acc = fmaf(sE[i], sH[i * ldh + j], acc);
            sA[tid] = acc;                                   // sA ([4][64] >= nthreads floats) is free again here
            __syncthreads();
            if (tid < L1) {
                float v = 0.f;
                for (int g = 0; g < G; ++g) v += sA[g * L1 + tid];
                out[2 + tid] = v;
            }
        }
        __syncthreads();       // sH / sE / sG / sA are rewritten by the next model
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core scores kernel (192-d features, L1 = 16 / 32 / 64, n_models * L1 <= 80): the first Linear of EVERY fold in one
// tcgen05 GEMM per 128-instance tile AND both gate Linears (a || b stacked along N) as a second, block-diagonal tcgen05 GEMM
// whose A operand is h1 in tensor memory — at fp32-level accuracy by operand splitting.
//
// kind::tf32 reads fp32 words and ignores the low 13 mantissa bits, i.e. it multiplies hi(x) = x & 0xFFFFE000.  With
// lo(x) = x - hi(x) (exact in fp32):   x w = hi(x) hi(w) + hi(x) lo(w) + lo(x) hi(w) + O(2^-21 |x w|)
// First Linear (GEMM 1, per tile: six K-slices of 32 features):
//   terms 1 + 2: A = X tile as loaded by TMA,      B = [W1 ; lo(W1)] stacked along N (the hardware truncates both operands;
//                                                  lo(W1) is precomputed per launch): ONE pass over A, two accumulator halves
//                                                  that the epilogue adds — shared-memory bandwidth (TMA writes, operand reads,
//                                                  LDS) is the busiest unit of this kernel, and A is 3/4 of the operand bytes
//   term 3:      A = lo(X) (tensor memory),        B = W1, accumulated onto the first half
// Gate (GEMM 2, per tile and fold: K = L1, N = 2 D = L1):  the epilogue writes h1 = relu(acc + b1) over the first accumulator
//   half and lo(h1) over the second (tcgen05.st, thread = row = lane), and the MMA warp issues
//   G = h1 [Wa;Wb]^T + h1 lo([Wa;Wb])^T + lo(h1) [Wa;Wb]^T   into a ring of G slots (L1 columns each).
//   Why: with thread = instance every gate weight is warp-uniform, and a uniform shared-memory load still costs one LSU
//   wavefront per 4 bytes — ncu on the previous kernel (profiles/r02o_clam_tc_5fold_full.csv): 383 shared-load wavefronts per
//   warp and fold, the LSU pipe at 72 % of its peak, 272 FFMA per instance and fold.  (The same constants read from the constant
//   bank are slower still: tools/patches/README.md.)  On the tensor core the gate costs 3 L1 / 8 small MMAs per fold and the
//   epilogue keeps only bias, exp / rcp, the score dot product and the pooling partials.
// All folds' W1 are stacked along N (N = n_models * L1), so the 98 KB feature tile is read from HBM once AND multiplied once
// for the whole ensemble.  Persistent CTAs walk the (bag, 128-instance chunk) work table.
//   warp 0      TMA producer: X in 6 K-slices [128 rows x 32 fp32] (SWIZZLE_128B) through a ring of up to 6 stages
//   warp 1      MMA issuer: GEMM 1 of tile t, then the gate of tile t - 1 (its h1 was written while GEMM 1 of tile t ran).
//               The tensor pipe executes in issue order, so GEMM 1 of tile t + 2 may overwrite the accumulator the gate of
//               tile t read its A operand from without any further hand-shake.
//   warps 2-3   idle (they only complete the first warpgroup, which gives its registers away)
//   warps 4-7   lo(X): one thread per row reads the slice from shared memory and writes lo(x) into TENSOR memory (32 columns
//               per stage of a second, shorter ring): term 3 is an A-from-TMEM MMA, so the shared-memory ring holds features only
//   warps 8-11, 12-15  two epilogue warpgroups (even / odd tiles = accumulator 0 / 1), thread = row.
//               Phase A (all folds): TMEM -> +b1, ReLU -> h1 kept in registers for the pooling partials, h1 and lo(h1) -> TMEM.
//               Phase B1 (per fold): G slot -> +bias, tanh * sigmoid on bare ex2 / rcp (operand pre-scaled), score.
//               Phase B2 (all folds together): softmax partials and sum_i e_i h1_i of the warp's 32 rows.
//               Phase B3: the four warps' records merged in shared memory -> one partial record per (tile, fold).
// Registers: 512 threads launch with 128 each; setmaxnreg moves them to 72 (first warpgroup) / 64 (lo) / 184 (epilogue; h1 of up to
// five folds = 80 registers stays live across the phases): 64,512 of the 65,536 the CTA launched with.  (setmaxnreg.inc beyond
// what the CTA's own warpgroups released never completes: the kernel then ends in mbar_wait's trap.)
// ---------------------------------------------------------------------------------------------------------------------
// Ring depth, measured on one hipt_smaller fold (tools/gpu_ring_depth.sh): 4 stages 128.6 us, 5: 119.8, 6: 116.5, 8: 117.3,
// 12: 119.4 — the kernel is not starved for bytes in flight from five stages (80 KB per SM) on, which is what is left beside the
// weights of five folds; more than six stages only deepen the DRAM queues.
#ifndef HB_CLAM_MAX_STAGES
#define HB_CLAM_MAX_STAGES 6
#endif
constexpr int TC_M = 128, TC_KS = 32, TC_NSL = 6, TC_THREADS = 512, TC_SLICE_BYTES = TC_M * 128, TC_MAX_STAGES = HB_CLAM_MAX_STAGES;
constexpr int TC_MAX_GSLOTS = 8;
constexpr float TC_SCALE_A = 2.8853900817779268f, TC_SCALE_B = -1.4426950408889634f;    // 2 log2(e), -log2(e)
// TMEM: two accumulators of acc_stride columns (2 ntot — the X W1 and X lo(W1) halves, later h1 and lo(h1) — rounded up to 32),
// then g_slots gate accumulators of L1 columns, then 32 columns of lo(X) per stage of the lo ring
__host__ __device__ inline int clam_tc_acc_stride(int ntot) { return (2 * ntot + 31) & ~31; }
// gate slots: one ring shared by both epilogue warpgroups, one more slot than a tile has folds (so the gate of tile t + 1 can be
// issued while the last fold of tile t is still being read) if tensor memory has the room beside three lo(X) stages; at least
// n_models (a slot is never reused inside a tile: its reader only starts when the whole tile's gate has been committed)
__host__ __device__ inline int clam_tc_gslots(int n_models, int L1) {
    const int room = (512 - 2 * clam_tc_acc_stride(n_models * L1) - 3 * 32) / L1;
    return n_models + 1 < room ? n_models + 1 : room;
}
__host__ __device__ inline int clam_tc_lo_stages(int n_models, int L1) {
    const int st = (512 - 2 * clam_tc_acc_stride(n_models * L1) - clam_tc_gslots(n_models, L1) * L1) / 32;
    return st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
}
__host__ __device__ inline bool clam_tc_ok(int L0, int L1, int D, int n_models) {
    return L0 == 192 && (L1 == 16 || L1 == 32 || L1 == 64) && D * 2 == L1 && n_models * L1 <= 80;
}
// floats of per-fold epilogue constants: b1 [L1] | ba || bb [L1] | Wc [D] | bc
__host__ __device__ inline int clam_tc_fold_floats(int L1, int D) { return ((2 * L1 + D + 1) + 3) & ~3; }
// bytes of one fold's gate operand [Wa ; Wb] (K-major rows of 128 B, K padded to 32 per slice), hi and lo copies
__host__ __device__ inline int clam_tc_gate_bytes(int L1) { return 2 * ((L1 + 31) / 32) * L1 * 128; }
// bytes of everything except the X ring
__host__ __device__ inline size_t clam_tc_fixed_bytes(int n_models, int L1, int D) {
    const int ntot = n_models * L1;
    return 1024 + 2 * static_cast<size_t>(TC_NSL) * ntot * 128 + static_cast<size_t>(n_models) * clam_tc_gate_bytes(L1) +
           static_cast<size_t>(n_models) * (clam_tc_fold_floats(L1, D) + 8 * (L1 + 2)) * sizeof(float) + 80 * 8;
}
// as many ring stages as fit, up to TC_MAX_STAGES (a stage carries 16 KB of features)
__host__ __device__ inline int clam_tc_stages(int n_models, int L1, int D) {
    const long long room = 232448LL - static_cast<long long>(clam_tc_fixed_bytes(n_models, L1, D));
    const int st = static_cast<int>(room / TC_SLICE_BYTES);
    return st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
}

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
    if constexpr (N == 16) tmem_ld_32x16(taddr, *reinterpret_cast<uint32_t(*)[16]>(v));
    else {
#pragma unroll
        for (int i = 0; i < N; i += 32) tmem_ld_32x32(taddr + i, *reinterpret_cast<uint32_t(*)[32]>(v + i));
    }
}

// Epilogue of clam_scores_tc_kernel for the folds [M0, M1) of every TSTEP-th tile starting at tile t0 (thread = tile row).
// The kernel runs it with all folds and TSTEP = 2: the two epilogue warpgroups alternate tiles.
// History at five folds: while this epilogue was 2,400 straight-line instructions (39 KB; h1 of every fold lives in registers, so
// every loop over folds unrolls) ncu attributed 49 % of the epilogue warps' samples to stall_no_inst; the bare ex2 / rcp gate
// brought it to 1,700 and the kernel from 225 to 189 us, a version with rolled fold loops (h1 re-read from tensor memory)
// measured 208 us (tools/patches/README.md).  In the final kernel the epilogue WAITS for 46 % of its samples: the slice feed
// paces five folds (profiles/r02ad_clam_trace.txt).
struct TcEpi {
    uint64_t *acc_full, *h_full, *g_done, *g_empty;
    uint32_t tmem_base, acc_stride, t_g, sCu, s_merge, bar_id;
    int g_slots, n_work, work_cap, total_instances;
    const int32_t* work;
    const int32_t* bag_offsets;
    float* a_raw;
    float* partials;
};
template <int L1, int F, int M0, int M1, int TSTEP>
__device__ __forceinline__ void clam_tc_epilogue(const TcEpi& E, uint32_t t0, int warp, int lane) {
    constexpr int D = L1 / 2, NF = M1 - M0, ntot = F * L1;
    constexpr int fold_floats = ((2 * L1 + D + 1) + 3) & ~3;
    const int r = (warp & 3) * 32 + lane;                        // tile row = TMEM lane (warp % 4 = lane quadrant)
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t sCu = E.sCu, t_g = E.t_g;
    const int g_slots = E.g_slots;
    uint32_t g_pos = (t0 * F + M0) % g_slots;                    // tile t, fold m uses gate slot (t F + m) mod g_slots
    for (uint32_t t = t0; static_cast<long long>(blockIdx.x) + static_cast<long long>(t) * gridDim.x < E.n_work; t += TSTEP) {
        const int wi = blockIdx.x + t * gridDim.x;
        const uint32_t b = t & 1;                                // accumulator of the tile
        const int bag = E.work[2 * wi], chunk = E.work[2 * wi + 1];
        const int start = E.bag_offsets[bag];
        const int n_valid = min(TC_M, E.bag_offsets[bag + 1] - start - chunk * TC_M);
        const bool valid = r < n_valid;
        const uint32_t t_acc = E.tmem_base + lane_off + b * E.acc_stride;
        float h[NF * L1];
        // ---- phase A: h1 of every fold -> registers and (with its low part) back into tensor memory.  Software-pipelined:
        // the accumulator columns of pass p + 1 are in flight while pass p is computed (inline asm is a compiler barrier,
        // so the overlap has to be spelled out)
        constexpr int CW = L1 < 32 ? L1 : 32;                // columns per pass (register pressure at L1 = 64)
        constexpr int PPF = L1 / CW, NP = NF * PPF;          // passes per fold, passes of this warpgroup per tile
        uint32_t vb[2][CW], wb[2][CW];                       // the X W1 and X lo(W1) halves of the accumulator, double-buffered
        mbar_wait(&E.acc_full[b], (t >> 1) & 1);
        tc_fence_after();
        tmem_ld_cols<CW>(t_acc + M0 * L1, vb[0]);
        tmem_ld_cols<CW>(t_acc + ntot + M0 * L1, wb[0]);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int mi = p / PPF, m = M0 + mi, c0 = (p % PPF) * CW;   // mi: fold index inside this warpgroup's range
            uint32_t* v = vb[p & 1];
            uint32_t* v2 = wb[p & 1];
            tmem_ld_wait();
            if (p + 1 < NP) {
                const int m1 = M0 + (p + 1) / PPF, c1 = ((p + 1) % PPF) * CW;
                tmem_ld_cols<CW>(t_acc + m1 * L1 + c1, vb[(p + 1) & 1]);
                tmem_ld_cols<CW>(t_acc + ntot + m1 * L1 + c1, wb[(p + 1) & 1]);
            }
#pragma unroll
            for (int j = 0; j < CW; j += 4) {
                const float4 b1 = lds_f4(sCu + (m * fold_floats + c0 + j) * 4);
                float* hj = h + mi * L1 + c0 + j;
                hj[0] = fmaxf((__uint_as_float(v[j + 0]) + __uint_as_float(v2[j + 0])) + b1.x, 0.f);
                hj[1] = fmaxf((__uint_as_float(v[j + 1]) + __uint_as_float(v2[j + 1])) + b1.y, 0.f);
                hj[2] = fmaxf((__uint_as_float(v[j + 2]) + __uint_as_float(v2[j + 2])) + b1.z, 0.f);
                hj[3] = fmaxf((__uint_as_float(v[j + 3]) + __uint_as_float(v2[j + 3])) + b1.w, 0.f);
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const float hv = h[mi * L1 + c0 + j];
                v[j] = __float_as_uint(hv);
                v2[j] = __float_as_uint(hv - __uint_as_float(v[j] & 0xFFFFE000u));
            }
#pragma unroll
            for (int j = 0; j < CW; j += 16) {
                tmem_st_32x16(t_acc + m * L1 + c0 + j, v + j);
                tmem_st_32x16(t_acc + ntot + m * L1 + c0 + j, v2 + j);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&E.h_full[b]);
        // ---- phase B1: per fold, gate pre-activations from the slot -> score (the next fold's slot is in flight meanwhile)
        float A[NF];
        {
            constexpr int GB = NF > 1 ? 2 : 1;
            uint32_t gbuf[GB][L1];
            mbar_wait(&E.g_done[b], (t >> 1) & 1);
            tc_fence_after();
            tmem_ld_cols<L1>(t_g + lane_off + g_pos * L1, gbuf[0]);
#pragma unroll
            for (int mi = 0; mi < NF; ++mi) {
                const int m = M0 + mi;
                const uint32_t cm = sCu + m * fold_floats * 4;
                const uint32_t* g = gbuf[mi % GB];
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&E.g_empty[g_pos]);
                if (++g_pos == static_cast<uint32_t>(g_slots)) g_pos = 0;
                if (mi + 1 < NF) tmem_ld_cols<L1>(t_g + lane_off + g_pos * L1, gbuf[(mi + 1) % GB]);
                float a = __uint_as_float(lds_u1(cm + (2 * L1 + D) * 4));          // bc
#pragma unroll
                for (int d = 0; d < D; d += 4) {
                    const float4 ba = lds_f4(cm + (L1 + d) * 4), bb = lds_f4(cm + (L1 + D + d) * 4);
                    const float4 wc = lds_f4(cm + (2 * L1 + d) * 4);
                    const float bav[4] = {ba.x, ba.y, ba.z, ba.w}, bbv[4] = {bb.x, bb.y, bb.z, bb.w};
                    const float wcv[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // tanh(a) sigmoid(b) = (1 - 2 / (1 + e^2a)) / (1 + e^-b) on bare ex2 / rcp (the gate operand and its biases
                        // are pre-scaled by 2 log2 e and -log2 e); inf-safe without clamps: rcp(inf) = 0
                        const float e2a = ex2_approx(__uint_as_float(g[d + i]) + bav[i]);
                        const float enb = ex2_approx(__uint_as_float(g[D + d + i]) + bbv[i]);
                        const float tasb = fmaf(-2.0f, rcp_approx(1.0f + e2a), 1.0f) * rcp_approx(1.0f + enb);
                        a = fmaf(wcv[i], tasb, a);
                    }
                }
                A[mi] = a;
                if (valid) E.a_raw[static_cast<size_t>(m) * E.total_instances + start + chunk * TC_M + r] = a;
            }
        }
        // ---- phase B2: softmax / pooling partials of the WARP's 32 rows for all folds at once.  Every reduction step runs over
        // the folds in its inner loop, so the shuffles of different folds overlap instead of forming one long dependent chain
        // per fold.
        float mx[NF], e[NF], sum[NF];
#pragma unroll
        for (int m = 0; m < NF; ++m) mx[m] = valid ? A[m] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int m = 0; m < NF; ++m) mx[m] = fmaxf(mx[m], __shfl_xor_sync(0xffffffffu, mx[m], o));
        }
#pragma unroll
        for (int m = 0; m < NF; ++m) { e[m] = valid ? ex2_approx((A[m] - mx[m]) * 1.4426950408889634f) : 0.f; sum[m] = e[m]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int m = 0; m < NF; ++m) sum[m] += __shfl_xor_sync(0xffffffffu, sum[m], o);
        }
        // sum_i e_i h1[i][:] over the warp's 32 rows: recursive halving (L1 - 1 shuffles per fold instead of 5 L1): each
        // step a lane hands over the half of its columns its partner keeps; the surviving columns end up on
        // lane-dependent positions: L1 = 16: column lane >> 1, 32: column lane, 64: columns 2 lane, 2 lane + 1
#pragma unroll
        for (int m = 0; m < NF; ++m) {
#pragma unroll
            for (int j = 0; j < L1; ++j) h[m * L1 + j] *= e[m];
        }
#pragma unroll
        for (int half = L1 / 2, bit = 16; half >= 1 && bit >= 1; half >>= 1, bit >>= 1) {
            const bool up = lane & bit;
#pragma unroll
            for (int m = 0; m < NF; ++m) {
#pragma unroll
                for (int j = 0; j < half; ++j) {
                    const float keep = up ? h[m * L1 + j + half] : h[m * L1 + j];
                    const float send = up ? h[m * L1 + j] : h[m * L1 + j + half];
                    h[m * L1 + j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
            }
        }
        // ---- phase B3: the four warps' records of every fold -> ONE record per (tile, fold).  Per-warp records in global
        // memory made clam_combine_kernel walk four times as many (its time is the latency chain over the longest bag's
        // records: 30 us of a 200 us hipt_medium forward); merging here costs two named barriers per TILE (not per fold) and
        // ~30 instructions.
        constexpr int REC = L1 + 2;
        const uint32_t sw = E.s_merge + (warp & 3) * NF * REC * 4;           // this warp's [NF][REC] slot
        named_bar_sync(E.bar_id, 128);                                       // the previous tile's records have been read
#pragma unroll
        for (int m = 0; m < NF; ++m) {
            if (lane == 0) { sts_f1(sw + (m * REC) * 4, mx[m]); sts_f1(sw + (m * REC + 1) * 4, sum[m]); }
            if constexpr (L1 == 16) {                        // 16 columns over 32 lanes: pairs share one
                const float hv = h[m * L1] + __shfl_xor_sync(0xffffffffu, h[m * L1], 1);
                if ((lane & 1) == 0) sts_f1(sw + (m * REC + 2 + (lane >> 1)) * 4, hv);
            } else if constexpr (L1 == 32) {
                sts_f1(sw + (m * REC + 2 + lane) * 4, h[m * L1]);
            } else {
                sts_f1(sw + (m * REC + 2 + 2 * lane) * 4, h[m * L1]);
                sts_f1(sw + (m * REC + 3 + 2 * lane) * 4, h[m * L1 + 1]);
            }
        }
        named_bar_sync(E.bar_id, 128);
        for (int idx = (warp & 3) * 32 + lane; idx < NF * L1; idx += 128) {
            const int m = idx / L1, j = idx - m * L1;
            float mw[4], gm = -INFINITY;
#pragma unroll
            for (int w = 0; w < 4; ++w) { mw[w] = __uint_as_float(lds_u1(E.s_merge + ((w * NF + m) * REC) * 4)); gm = fmaxf(gm, mw[w]); }
            float vec = 0.f, tot = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float wgt = ex2_approx((mw[w] - gm) * 1.4426950408889634f);          // a warp without valid rows: -inf -> 0
                vec = fmaf(wgt, __uint_as_float(lds_u1(E.s_merge + ((w * NF + m) * REC + 2 + j) * 4)), vec);
                tot = fmaf(wgt, __uint_as_float(lds_u1(E.s_merge + ((w * NF + m) * REC + 1) * 4)), tot);
            }
            float* out = E.partials + (static_cast<size_t>(M0 + m) * E.work_cap + wi) * REC;
            if (j == 0) *reinterpret_cast<float2*>(out) = make_float2(gm, tot);
            out[2 + j] = vec;
        }
        g_pos = (g_pos + TSTEP * F - NF) % g_slots;                 // the folds / tiles of the other warpgroup
    }
}

template <int L1, int F>
__global__ void __launch_bounds__(TC_THREADS, 1)
clam_scores_tc_kernel(const __grid_constant__ CUtensorMap map_x, const int32_t* __restrict__ bag_offsets,
                      const __grid_constant__ ClamModels models, int n_bags, int total_instances,
                      const int32_t* __restrict__ prefix, const int32_t* __restrict__ work, int work_cap,
                      float* __restrict__ a_raw, float* __restrict__ partials) {
    constexpr int D = L1 / 2;
    constexpr int GKS = (L1 + 31) / 32;                         // 128-byte K slices of the gate operand
    constexpr int GSL = L1 * 128;                               // bytes of one such slice ([L1 rows][128 B])
    extern __shared__ uint8_t smem_raw_tc[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~uintptr_t(1023));
    constexpr int n_models = F;                                 // the fold count is a template parameter: every loop over
                                                                // folds unrolls, h1 of all folds stays in registers
    constexpr int ntot = n_models * L1;
    const int stages = clam_tc_stages(n_models, L1, D);
    const int lo_stages = clam_tc_lo_stages(n_models, L1);
    const int g_slots = clam_tc_gslots(n_models, L1);
    uint8_t* sXr = smem;                                        // [stages][X 16 KB]
    uint8_t* sWh = sXr + stages * TC_SLICE_BYTES;               // [6 slices][2 ntot rows][128 B]: W1 of all folds, then lo(W1)
    uint8_t* sWl = sWh + ntot * 128;                            // the lo rows of slice 0 (slice stride 2 ntot rows)
    uint8_t* sG = sWh + 2 * TC_NSL * ntot * 128;                // [n_models][hi, lo][GKS][L1 rows][128 B]: [Wa ; Wb]
    float* sC = reinterpret_cast<float*>(sG + n_models * clam_tc_gate_bytes(L1));   // [n_models][fold constants]
    constexpr int fold_floats = ((2 * L1 + D + 1) + 3) & ~3;
    float* sMerge = sC + n_models * fold_floats;                // [2 epilogue warpgroups][4 warps][F][L1 + 2]: partial records of a tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(sMerge + 2 * 4 * n_models * (L1 + 2));
    uint64_t* x_full = bars;              // [12]
    uint64_t* x_empty = bars + 12;        // [12]  MMA commit
    uint64_t* lo_full = bars + 24;        // [12]  128 lo(X) threads
    uint64_t* lo_empty = bars + 36;       // [12]  MMA commit
    uint64_t* acc_full = bars + 48;       // [2]   MMA commit: GEMM 1 of a tile is in the accumulator
    uint64_t* h_full = bars + 50;         // [2]   128 epilogue threads: h1 / lo(h1) are in tensor memory
    uint64_t* g_done = bars + 52;         // [2]   MMA commit: the gate pre-activations of every fold of a tile are in their slots
    uint64_t* g_empty = bars + 54;        // [8]   128 epilogue threads: the slot has been read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 62);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0 && lane == 0) tma_prefetch_desc(&map_x);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < TC_MAX_STAGES; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); mbar_init(&lo_full[i], 128); mbar_init(&lo_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&h_full[i], 128); mbar_init(&g_done[i], 1); }
        for (int i = 0; i < TC_MAX_GSLOTS; ++i) mbar_init(&g_empty[i], 128);
        fence_mbar_init();
    }
    const uint32_t tmem_cols = 512;                             // accumulators + gate slots + the lo(X) ring; one CTA per SM
    if (warp == 1) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
    // W1 (every fold) -> shared memory in the K-major SWIZZLE_128B layout of the B operand, hi as is, lo = w - trunc(w)
    for (int idx = tid; idx < ntot * 192; idx += TC_THREADS) {
        const int n = idx / 192, k = idx - n * 192;
        const int m = n / L1, j = n - m * L1;
        const float w = __ldg(models.m[m].p[0] + j * 192 + k);
        const int sl = k >> 5, kk = k & 31;
        const uint32_t off = sl * 2 * ntot * 128 + n * 128 + ((((kk >> 2) ^ (n & 7)) << 4) | ((kk & 3) << 2));
        *reinterpret_cast<float*>(sWh + off) = w;
        *reinterpret_cast<float*>(sWl + off) = w - __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    }
    // gate operand of every fold: row n = gate unit (a: n < D, b: n >= D), K = the L1 hidden units; same layout, hi and lo
    for (int idx = tid; idx < n_models * L1 * L1; idx += TC_THREADS) {
        const int m = idx / (L1 * L1), rem = idx - m * L1 * L1, n = rem / L1, k = rem - n * L1;
        // pre-scaled so that the epilogue's exponentials are bare ex2: branch a by 2 log2(e) (e^2a), branch b by -log2(e) (e^-b)
        const float w = (n < D) ? __ldg(models.m[m].p[2] + n * L1 + k) * TC_SCALE_A : __ldg(models.m[m].p[4] + (n - D) * L1 + k) * TC_SCALE_B;
        const int sl = k >> 5, kk = k & 31;
        const uint32_t off = (m * 2 * GKS + sl) * GSL + n * 128 + ((((kk >> 2) ^ (n & 7)) << 4) | ((kk & 3) << 2));
        *reinterpret_cast<float*>(sG + off) = w;
        *reinterpret_cast<float*>(sG + off + GKS * GSL) = w - __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    }
    for (int m = 0; m < n_models; ++m) {
        float* c = sC + m * fold_floats;
        const ClamModel& w = models.m[m];
        for (int i = tid; i < L1; i += TC_THREADS) c[i] = __ldg(w.p[1] + i);
        for (int i = tid; i < D; i += TC_THREADS) {
            c[L1 + i] = __ldg(w.p[3] + i) * TC_SCALE_A; c[L1 + D + i] = __ldg(w.p[5] + i) * TC_SCALE_B; c[2 * L1 + i] = __ldg(w.p[6] + i);
        }
        if (tid == 0) c[2 * L1 + D] = __ldg(w.p[7]);
    }
    fence_proxy_async_smem();                                   // generic writes of W -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                                 // everything above overlapped clam_work_table_kernel
    pdl_launch_dependents();                                    // clam_combine_kernel: scheduled as this grid's CTAs retire
    const int n_work = prefix[n_bags];
    const uint32_t acc_stride = clam_tc_acc_stride(ntot);
    const uint32_t t_g = tmem_base + 2 * acc_stride;            // [g_slots][L1 columns]
    const uint32_t t_lo = t_g + g_slots * L1;                   // [lo_stages][32 columns]: lo(X) of a K-slice

    if (warp < 4) {
        setmaxnreg_dec<72>();
        if (warp == 0) {
            // -------------------------------------------------------------------------------------- TMA producer
            if (lane == 0) {
                uint32_t st = 0, ph = 0;                          // ring position and its phase parity (no divisions)
                for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
                    const int bag = work[2 * wi], chunk = work[2 * wi + 1];
                    const int row0 = bag_offsets[bag] + chunk * TC_M;
                    for (int sl = 0; sl < TC_NSL; ++sl) {
                        mbar_wait(&x_empty[st], ph ^ 1);
                        mbar_arrive_expect_tx(&x_full[st], TC_SLICE_BYTES);
                        tma_load_2d(sXr + st * TC_SLICE_BYTES, &map_x, &x_full[st], sl * TC_KS, row0);
                        if (++st == static_cast<uint32_t>(stages)) { st = 0; ph ^= 1; }
                    }
                }
            }
        } else if (warp == 1) {
            // -------------------------------------------------------------------------------------- MMA issuer
            const uint32_t idesc = umma_idesc_tf32(TC_M, ntot);             // lo(X) x W1
            const uint32_t idesc2 = umma_idesc_tf32(TC_M, 2 * ntot);        // X x [W1 ; lo(W1)]
            constexpr uint32_t idesc_g = umma_idesc_tf32(TC_M, L1);         // h1 x [Wa ; Wb] of one fold
            // gate of tile tt (accumulator tt & 1): one slot per fold
            uint32_t g_pos = 0, g_ph = 0;                        // next gate slot, parity of its current use
            bool g_wrapped = false;
            auto issue_gate = [&](uint32_t tt) {
                const uint32_t b = tt & 1;
                mbar_wait(&h_full[b], (tt >> 1) & 1);
                tc_fence_after();
                const uint32_t t_h = tmem_base + b * acc_stride;
                for (int m = 0; m < n_models; ++m) {
                    const uint32_t gs = g_pos;
                    if (g_wrapped) { mbar_wait(&g_empty[gs], g_ph ^ 1); tc_fence_after(); }
                    if (elect_one()) {
                        const uint32_t d_g = t_g + gs * L1;
#pragma unroll
                        for (int k8 = 0; k8 < L1 / 8; ++k8) {
                            const uint64_t dgh = umma_desc_k128(smem_u32(sG + (m * 2 * GKS + (k8 >> 2)) * GSL)) + 2 * (k8 & 3);
                            const uint64_t dgl = umma_desc_k128(smem_u32(sG + (m * 2 * GKS + GKS + (k8 >> 2)) * GSL)) + 2 * (k8 & 3);
                            umma_tf32_ts(d_g, t_h + m * L1 + 8 * k8, dgh, idesc_g, k8 != 0);
                            umma_tf32_ts(d_g, t_h + m * L1 + 8 * k8, dgl, idesc_g, 1);
                            umma_tf32_ts(d_g, t_h + ntot + m * L1 + 8 * k8, dgh, idesc_g, 1);
                        }
                        if (m == n_models - 1) umma_commit(&g_done[b]);
                    }
                    __syncwarp();
                    if (++g_pos == static_cast<uint32_t>(g_slots)) { g_pos = 0; g_ph ^= 1; g_wrapped = true; }
                }
            };
            uint32_t t = 0, st = 0, ph = 0, ls = 0, lph = 0;
            for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x, ++t) {
                const uint32_t b = t & 1;
                const uint32_t d_tmem = tmem_base + b * acc_stride;
                for (int sl = 0; sl < TC_NSL; ++sl) {
                    const uint64_t dx = umma_desc_k128(smem_u32(sXr + st * TC_SLICE_BYTES));
                    const uint64_t dwh = umma_desc_k128(smem_u32(sWh + sl * 2 * ntot * 128));
                    mbar_wait(&x_full[st], ph);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_tf32_ss(d_tmem, dx + 2 * kk, dwh + 2 * kk, idesc2, (sl | kk) != 0);
                    }
                    __syncwarp();
                    mbar_wait(&lo_full[ls], lph);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_tf32_ts(d_tmem, t_lo + ls * 32 + 8 * kk, dwh + 2 * kk, idesc, 1);
                        umma_commit(&x_empty[st]);
                        umma_commit(&lo_empty[ls]);
                        if (sl == TC_NSL - 1) umma_commit(&acc_full[b]);
                    }
                    __syncwarp();
                    if (++st == static_cast<uint32_t>(stages)) { st = 0; ph ^= 1; }
                    if (++ls == static_cast<uint32_t>(lo_stages)) { ls = 0; lph ^= 1; }
                }
                // (issuing this gate earlier — between two K-slices, as soon as h1 is in tensor memory — measured SLOWER:
                // 142 vs 131 us for one fold, 246 vs 243 us for five)
                if (t >= 1) issue_gate(t - 1);
            }
            if (t >= 1) issue_gate(t - 1);
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------------------------------ lo(X)
        setmaxnreg_dec<64>();
        const int r = (warp & 3) * 32 + lane;                    // tile row = TMEM lane (lane quadrant = warp % 4)
        const uint32_t t_my = t_lo + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint32_t st = 0, ph = 0, ls = 0, lph = 0;
        bool lo_wrapped = false;
        for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
            for (int sl = 0; sl < TC_NSL; ++sl) {
                const uint32_t xs = smem_u32(sXr + st * TC_SLICE_BYTES) + r * 128;
                mbar_wait(&x_full[st], ph);
                uint32_t lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {                    // 16-byte piece c = K elements 4 c .. 4 c + 3 of the slice
                    const uint4 v = lds_u4(xs + ((c ^ (r & 7)) << 4));
                    lo[4 * c + 0] = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
                    lo[4 * c + 1] = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
                    lo[4 * c + 2] = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
                    lo[4 * c + 3] = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
                }
                if (lo_wrapped) { mbar_wait(&lo_empty[ls], lph ^ 1); tc_fence_after(); }
                tmem_st_32x16(t_my + ls * 32, lo);
                tmem_st_32x16(t_my + ls * 32 + 16, lo + 16);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&lo_full[ls]);
                if (++st == static_cast<uint32_t>(stages)) { st = 0; ph ^= 1; }
                if (++ls == static_cast<uint32_t>(lo_stages)) { ls = 0; lph ^= 1; lo_wrapped = true; }
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------ epilogue
        setmaxnreg_inc<184>();
        const int eg = (warp - 8) >> 2;                           // epilogue warpgroup
        TcEpi E;
        E.acc_full = acc_full; E.h_full = h_full; E.g_done = g_done; E.g_empty = g_empty;
        E.tmem_base = tmem_base; E.acc_stride = acc_stride; E.t_g = t_g; E.sCu = smem_u32(sC);
        E.g_slots = g_slots; E.n_work = n_work; E.work_cap = work_cap; E.total_instances = total_instances;
        E.work = work; E.bag_offsets = bag_offsets; E.a_raw = a_raw; E.partials = partials;
        E.s_merge = smem_u32(sMerge + eg * 4 * n_models * (L1 + 2)); E.bar_id = 2 + eg;
        // the two warpgroups alternate tiles.  (Both on every tile, each with half of the folds — clam_tc_epilogue<L1, F, 0, FH, 1>
        // and <L1, F, FH, F, 1> — measured slower: 5 folds 240 vs 256 us, hipt_small 2 folds 176 vs 183 us in the same run; and an L2
        // prefetch of the tiles ahead of the shared-memory ring made every configuration 7-60 % slower.)
        clam_tc_epilogue<L1, F, 0, F, 2>(E, eg, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

constexpr int CB_THREADS = 256;
__global__ void __launch_bounds__(CB_THREADS) clam_combine_kernel(const int32_t* __restrict__ bag_offsets,
                                                                  const __grid_constant__ ClamModels models, int n_bags, int L1,
                                                                  int C, int work_cap, int CH, const int32_t* __restrict__ prefix,
                                                                  const float* __restrict__ partials,
                                                                  float* __restrict__ m_out, float* __restrict__ logits,
                                                                  float* __restrict__ y_prob, long long* __restrict__ y_hat,
                                                                  int paired) {
    extern __shared__ float sM[];                             // [L1] + [C] + [2 CB_THREADS] scratch + [8]
    float* sL = sM + L1;
    float* sP = sL + C;                                       // [2 CB_THREADS] per-group partial sums of M
    float* red = sP + 2 * CB_THREADS;
    const int bag = blockIdx.x, mi = paired ? blockIdx.x : blockIdx.y, tid = threadIdx.x;
    const int oi = paired ? bag : mi * n_bags + bag;          // paired (multi-trial): model m pools bag m only, compact outputs
    pdl_wait();                                               // no-op unless launched as a programmatic dependent
    const int len = bag_offsets[bag + 1] - bag_offsets[bag];
    const int n_chunks = (len + CH - 1) / CH;                 // partial records of the bag: one per chunk
    const size_t rec = L1 + 2;
    const float* __restrict__ base = partials + (static_cast<size_t>(mi) * work_cap + prefix[bag]) * rec;
    // records are spread over the threads (a 20,000-instance bag has 628 of them: a serial walk is latency-bound)
    float gm = -INFINITY;
    for (int c = tid; c < n_chunks; c += CB_THREADS) gm = fmaxf(gm, base[c * rec]);
    const float gmax = block_reduce_max_128(gm, red);
    float tl = 0.f;
    for (int c = tid; c < n_chunks; c += CB_THREADS) tl += base[c * rec + 1] * expf(base[c * rec] - gmax);
    const float total = block_reduce_sum_128(tl, red);
    const float inv = (n_chunks > 0) ? 1.0f / total : 0.f;
    if ((L1 & 1) == 0 && L1 <= CB_THREADS) {
        // a thread owns a column PAIR of a group of records; eight records in flight per thread (the walk is latency-bound:
        // a 20,000-instance bag has 628 records, and each is read exactly once)
        const int TPR = L1 / 2, G = CB_THREADS / TPR;         // threads per record, record groups walking the list in parallel
        const int cg = tid / TPR, jp = tid - cg * TPR;
        float acc0 = 0.f, acc1 = 0.f;
        if (cg < G) {
            const float* __restrict__ bj = base + 2 + 2 * jp;
            int c = cg;
            for (; c + 7 * G < n_chunks; c += 8 * G) {
                float2 v[8];
                float mr[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    v[u] = *reinterpret_cast<const float2*>(bj + static_cast<size_t>(c + u * G) * rec);
                    mr[u] = base[static_cast<size_t>(c + u * G) * rec];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float w = expf(mr[u] - gmax);
                    acc0 = fmaf(v[u].x, w, acc0); acc1 = fmaf(v[u].y, w, acc1);
                }
            }
            for (; c < n_chunks; c += G) {
                const float2 v = *reinterpret_cast<const float2*>(bj + static_cast<size_t>(c) * rec);
                const float w = expf(base[static_cast<size_t>(c) * rec] - gmax);
                acc0 = fmaf(v.x, w, acc0); acc1 = fmaf(v.y, w, acc1);
            }
        }
        sP[2 * tid] = acc0; sP[2 * tid + 1] = acc1;           // [group][column]: (cg TPR + jp) 2 = cg L1 + 2 jp
        __syncthreads();
        if (tid < L1) {
            float v = 0.f;
            for (int g = 0; g < G; ++g) v += sP[g * L1 + tid];
            v *= inv;
            sM[tid] = v;
            if (m_out) m_out[static_cast<size_t>(oi) * L1 + tid] = v;
        }
    } else {
        for (int j = tid; j < L1; j += blockDim.x) {
            float acc = 0.f;
            for (int c = 0; c < n_chunks; ++c) acc = fmaf(base[c * rec + 2 + j], expf(base[c * rec] - gmax), acc);
            acc *= inv;
            sM[j] = acc;
            if (m_out) m_out[static_cast<size_t>(oi) * L1 + j] = acc;
        }
    }
    __syncthreads();
    const float* Wcls = models.m[mi].p[8];
    const float* bcls = models.m[mi].p[9];
    // classifier: a warp per class, lanes over the hidden units (one thread per class walking L1 dependent global loads was
    // most of this kernel's time for the wider heads)
    for (int c = tid >> 5; c < C; c += CB_THREADS / 32) {
        float acc = 0.f;
        for (int j = tid & 31; j < L1; j += 32) acc = fmaf(__ldg(Wcls + static_cast<size_t>(c) * L1 + j), sM[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) {
            acc += __ldg(bcls + c);
            sL[c] = acc;
            if (logits) logits[static_cast<size_t>(oi) * C + c] = acc;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float mx = sL[0];
        int arg = 0;
        for (int c = 1; c < C; ++c) if (sL[c] > mx) { mx = sL[c]; arg = c; }
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(sL[c] - mx);
        if (y_prob) for (int c = 0; c < C; ++c) y_prob[static_cast<size_t>(oi) * C + c] = expf(sL[c] - mx) / s;
        if (y_hat) y_hat[oi] = arg;
    }
}

// workspace = [prefix: n_bags + 1 ints][work: 2 * cap ints][partials: n_models * cap * (L1 + 2) floats], cap = bound on
// the number of (bag, chunk) items.  The size query knows only max_bag_len: it uses the smallest chunk of any variant.
static size_t clam_ws_layout(size_t cap, int n_bags, int n_models, int L1, size_t* off_work, size_t* off_part) {
    size_t o = (static_cast<size_t>(n_bags) + 1) * sizeof(int32_t);
    o = (o + 15) & ~static_cast<size_t>(15);
    if (off_work) *off_work = o;
    o += 2 * cap * sizeof(int32_t);
    o = (o + 15) & ~static_cast<size_t>(15);
    if (off_part) *off_part = o;
    return o + static_cast<size_t>(n_models) * cap * (L1 + 2) * sizeof(float);
}
size_t clam_workspace_bytes(int max_bag_len, int n_bags, int n_models, int L1) {
    const size_t max_chunks = (static_cast<size_t>(max_bag_len) + 31) / 32;
    return clam_ws_layout(static_cast<size_t>(n_bags) * (max_chunks ? max_chunks : 1), n_bags, n_models, L1, nullptr, nullptr);
}

int clam_forward_launch(const float* feats, const int32_t* bag_offsets, int n_bags, int total_instances,
                        int max_bag_len, const void* const* weights_host, int n_models, int L0, int L1, int D, int C,
                        float* a_raw, float* m_out, float* logits, float* y_prob, long long* y_hat, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream, float dropout_p, unsigned long long dropout_seed,
                        const unsigned long long* paired_seeds) {
    if (n_bags <= 0) return 0;
    // paired_seeds != NULL: the multi-trial training forward — model m on bag m only, its own dropout seed
    const int paired = paired_seeds != nullptr;
    if (paired && n_bags != n_models) return set_error("hb_clam: the paired launch needs one bag per model");
    ClamDrops drs;
    memset(&drs, 0, sizeof(drs));
    drs.d[0] = clam_drop_make(dropout_p, dropout_seed);
    if (paired) for (int m = 0; m < n_models && m < 8; ++m) drs.d[m] = clam_drop_make(dropout_p, paired_seeds[m]);
    const bool dropping = dropout_p > 0.f;
    if (dropping && !clam_is192(L0, L1, D))
        return set_error("hb_clam: training-mode dropout is implemented for the HIPT heads (192-d features, L1 <= 128)");
    if (n_models < 1 || n_models > CLAM_MAX_MODELS) return set_error("hb_clam: n_models must be 1..%d", CLAM_MAX_MODELS);
    if (L0 % CLAM_KC != 0) return set_error("hb_clam: L0=%d must be a multiple of %d", L0, CLAM_KC);
    if (L1 < 1 || D < 1 || C < 1 || C > 64) return set_error("hb_clam: bad dims L1=%d D=%d C=%d", L1, D, C);
    if (!bag_offsets || !weights_host || !a_raw || !workspace) return set_error("hb_clam: null argument");
    if (total_instances > 0 && !feats) return set_error("hb_clam: null features");
    if ((reinterpret_cast<uintptr_t>(feats) & 15) != 0) return set_error("hb_clam: features must be 16 B aligned");
    const int CH = clam_chunk_for(L0, L1, D);
    const int max_chunks = (max_bag_len + CH - 1) / CH;
    // (bag, chunk) items: every bag has at most one partial chunk
    size_t cap = static_cast<size_t>(total_instances) / CH + n_bags;
    const size_t cap2 = static_cast<size_t>(n_bags) * (max_chunks ? max_chunks : 1);
    if (cap2 < cap) cap = cap2;
    if (cap < 1) cap = 1;
    size_t off_work, off_part;
    const size_t need = clam_ws_layout(cap, n_bags, n_models, L1, &off_work, &off_part);
    if (workspace_bytes < need) return set_error("hb_clam: workspace %zu < %zu bytes", workspace_bytes, need);
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return set_error("hb_clam: workspace must be 16 B aligned");
    int32_t* prefix = static_cast<int32_t*>(workspace);
    int32_t* work = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + off_work);
    const int work_cap = static_cast<int>(cap);
    ClamModels models;
    memset(&models, 0, sizeof(models));
    for (int m = 0; m < n_models; ++m)
        for (int k = 0; k < 10; ++k) {
            models.m[m].p[k] = static_cast<const float*>(weights_host[m * 10 + k]);
            if (!models.m[m].p[k]) return set_error("hb_clam: weight pointer %d of model %d is null", k, m);
        }
    float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + off_part);
    static int tc_env = -1;
    if (tc_env < 0) { const char* e = getenv("HB_CLAM_TC"); tc_env = (e && e[0] == '0') ? 0 : 1; }
    if (paired && !clam_is192(L0, L1, D)) return set_error("hb_clam: the paired launch is implemented for the HIPT heads (192-d features)");
    if (tc_env && !dropping && !paired && max_chunks > 0 && total_instances >= TC_M && clam_tc_ok(L0, L1, D, n_models)) {
        // tensor-core path: 128-instance chunks
        size_t cap_tc = static_cast<size_t>(total_instances) / TC_M + n_bags;
        const size_t cap_tc2 = static_cast<size_t>(n_bags) * ((max_bag_len + TC_M - 1) / TC_M);
        if (cap_tc2 < cap_tc) cap_tc = cap_tc2;
        size_t off_work_tc, off_part_tc;
        const size_t need_tc = clam_ws_layout(cap_tc, n_bags, n_models, L1, &off_work_tc, &off_part_tc);
        if (workspace_bytes < need_tc) return set_error("hb_clam: workspace %zu < %zu bytes", workspace_bytes, need_tc);
        int32_t* work = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + off_work_tc);
        float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + off_part_tc);
        const int work_cap = static_cast<int>(cap_tc);
        {
            ProfScope ps0(15, stream);
            clam_work_table_kernel<<<1, 1024, 0, stream>>>(bag_offsets, n_bags, TC_M, prefix, work, work_cap);
            count_launch();
            HB_CUDA_OK(cudaGetLastError());
        }
        CUtensorMap map_x;
        if (encode_tmap_2d(&map_x, TMAP_F32, feats, static_cast<uint64_t>(total_instances), 192, 192 * 4, TC_M, TC_KS)) return -1;
        const int stages = clam_tc_stages(n_models, L1, D);
        if (stages < 2) return set_error("hb_clam: tensor-core path does not fit shared memory (n_models %d, L1 %d)", n_models, L1);
        const size_t smem = clam_tc_fixed_bytes(n_models, L1, D) + static_cast<size_t>(stages) * TC_SLICE_BYTES;
        decltype(&clam_scores_tc_kernel<16, 1>) kern = nullptr;
        if (L1 == 16) {
            decltype(kern) k16[5] = {clam_scores_tc_kernel<16, 1>, clam_scores_tc_kernel<16, 2>, clam_scores_tc_kernel<16, 3>,
                                     clam_scores_tc_kernel<16, 4>, clam_scores_tc_kernel<16, 5>};
            kern = k16[n_models - 1];
        } else if (L1 == 32) {
            kern = n_models == 1 ? clam_scores_tc_kernel<32, 1> : clam_scores_tc_kernel<32, 2>;
        } else {
            kern = clam_scores_tc_kernel<64, 1>;
        }
        if (clam_tc_lo_stages(n_models, L1) < 2 || clam_tc_gslots(n_models, L1) < n_models) return set_error("hb_clam: tensor-core path does not fit tensor memory");
        if (set_max_dynamic_smem(reinterpret_cast<const void*>(kern), 232448)) return -1;
        int grid = (total_instances / TC_M) + n_bags;
        if (grid > work_cap) grid = work_cap;
        if (grid > num_sms()) grid = num_sms();
        {
            ProfScope ps(10, stream);
            // programmatic dependent launch: the prologue (barriers, tensor-memory allocation, weight staging) overlaps the work
            // table kernel; pdl_wait() in the kernel orders the first read of the table
            cudaLaunchAttribute pdl[1];
            pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            pdl[0].val.programmaticStreamSerializationAllowed = 1;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
            cfg.attrs = pdl; cfg.numAttrs = 1;
            const int32_t* prefix_c = prefix; const int32_t* work_c = work;
            HB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_x, bag_offsets, models, n_bags, total_instances, prefix_c, work_c,
                                          work_cap, a_raw, partials));
            count_launch();
        }
        {
            ProfScope ps2(11, stream);
            cudaLaunchAttribute pdl[1];
            pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            pdl[0].val.programmaticStreamSerializationAllowed = 1;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(n_bags, n_models); cfg.blockDim = dim3(CB_THREADS);
            cfg.dynamicSmemBytes = (L1 + C + 2 * CB_THREADS + 8) * sizeof(float); cfg.stream = stream;
            cfg.attrs = pdl; cfg.numAttrs = 1;
            const int32_t* prefix_c = prefix; const float* partials_c = partials;
            HB_CUDA_OK(cudaLaunchKernelEx(&cfg, clam_combine_kernel, bag_offsets, models, n_bags, L1, C, work_cap, static_cast<int>(TC_M),
                                          prefix_c, partials_c, m_out, logits, y_prob, y_hat, 0));
            count_launch();
        }
        return 0;
    }
    clam_work_table_kernel<<<1, 1024, 0, stream>>>(bag_offsets, n_bags, CH, prefix, work, work_cap);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    if (max_chunks > 0 && clam_is192(L0, L1, D)) {
        const size_t smem = (static_cast<size_t>(CL_CH) * CL_XS + clam_sw_floats(L1) + static_cast<size_t>(CL_CH) * (L1 + 4) +
                             2 * static_cast<size_t>(D) * L1 + ((D / 4 > 4 ? D / 4 : 4) + 1) * CL_CH + 8 + L1 + 3 * D + 4) * sizeof(float);
        const int threads = L1 >= 64 ? 256 : CL_THREADS;
        // the five HIPT heads (model_clam.py:81: 8/4, 16/8, 32/16, 64/32, 128/64) get compile-time strides
        auto kern = (L1 <= 16) ? clam_scores192_kernel<4, 0> : clam_scores192_kernel<8, 0>;
        if (D * 2 == L1) {
            if (L1 == 8) kern = clam_scores192_kernel<4, 8>;
            else if (L1 == 16) kern = clam_scores192_kernel<4, 16>;
            else if (L1 == 32) kern = clam_scores192_kernel<8, 32>;
            else if (L1 == 64) kern = clam_scores192_kernel<8, 64>;
            else if (L1 == 128) kern = clam_scores192_kernel<8, 128>;
        }
        if (set_max_dynamic_smem(reinterpret_cast<const void*>(kern), 220 * 1024)) return -1;
        ProfScope ps(10, stream);
        kern<<<work_cap, threads, smem, stream>>>(feats, bag_offsets, models, n_models, n_bags, total_instances, L1, D,
                                                  prefix, work, work_cap, a_raw, partials, drs, paired);
        count_launch();
        HB_CUDA_OK(cudaGetLastError());
    } else if (max_chunks > 0) {
        const size_t smem = (static_cast<size_t>(CH) * (CLAM_KC + 1) + CLAM_KC * CLAM_OB + CH + 4 +
                             static_cast<size_t>(CH) * (L1 + 1)) * sizeof(float);
        if (smem > 220 * 1024) return set_error("hb_clam: L1=%d too large for the fused kernel", L1);
        if (set_max_dynamic_smem(reinterpret_cast<const void*>(clam_scores_kernel), 220 * 1024)) return -1;
        ProfScope ps(10, stream);
        clam_scores_kernel<<<work_cap, CH, smem, stream>>>(feats, bag_offsets, models, n_models, n_bags, total_instances,
                                                           L0, L1, D, prefix, work, work_cap, a_raw, partials);
        count_launch();
    HB_CUDA_OK(cudaGetLastError());
    }
    dim3 grid2(n_bags, paired ? 1 : n_models);
    ProfScope ps2(11, stream);
    clam_combine_kernel<<<grid2, CB_THREADS, (L1 + C + 2 * CB_THREADS + 8) * sizeof(float), stream>>>(bag_offsets, models, n_bags, L1, C,
                                                                                    work_cap, CH, prefix, partials, m_out,
                                                                                    logits, y_prob, y_hat, paired);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}


// =====================================================================================================================
// Training step around CLAM_SB.forward (utils/core_utils.py:409-423: loss = CE(logits, label); loss.backward();
// optimizer.step()) for one bag and one weight set, 192-d features.
//
// Backward with recomputation (1,544 algorithmic HBM bytes per instance over forward + backward): the forward keeps only
// A_raw [N] and M [L1]; the backward re-reads the feature tile, recomputes h1 and the gate, and forms
//   dM = Wcls^T dlogits (+ dM_ext),  s = dM . M,  alpha_i = softmax(A)_i,  dA_i = alpha_i (dM . h1_i - s) (+ dA_ext_i)
//   dpre_a = dA Wc b (1 - a^2),  dpre_b = dA Wc a b (1 - b),  dh1 = alpha dM + Wa^T dpre_a + Wb^T dpre_b,  dz = dh1 [h1 > 0]
// and the parameter gradients as per-chunk sums added to global memory with atomics (model_clam.py:59-63, 147-183).
// clam_bwd_prep_kernel: one CTA: softmax statistics of A_raw, dM, s, classifier gradients, zeroes the other gradients.
// clam_bwd192_kernel: one CTA per 64-instance chunk.
// =====================================================================================================================
struct ClamGrads { float* p[10]; };

__device__ __forceinline__ void clam_bwd_prep_body(const float* __restrict__ a_raw, int N, const float* __restrict__ M,
                                                   const float* __restrict__ dlogits, const float* __restrict__ dM_ext,
                                                   const float* __restrict__ Wcls, const ClamGrads& g,
                                                   int L1, int D, int C, float* __restrict__ ctx,
                                                   const float* __restrict__ logits, const long long* __restrict__ label,
                                                   float* __restrict__ loss_out) {
    __shared__ float red[8];
    __shared__ float s_dl[64];
    const int tid = threadIdx.x;
    // cross-entropy mode (nn.CrossEntropyLoss of the bag logits against the slide label, core_utils.py:413): dlogits =
    // softmax(logits) - onehot(label), loss = logsumexp(logits) - logits[label]; otherwise dlogits comes from the caller
    if (logits != nullptr) {
        if (tid == 0) {
            float mxl = logits[0];
            for (int c = 1; c < C; ++c) mxl = fmaxf(mxl, logits[c]);
            float se = 0.f;
            for (int c = 0; c < C; ++c) se += expf(logits[c] - mxl);
            const int y = static_cast<int>(*label);
            for (int c = 0; c < C; ++c) s_dl[c] = expf(logits[c] - mxl) / se - (c == y ? 1.f : 0.f);
            if (loss_out) *loss_out = logf(se) + mxl - logits[y];
        }
        __syncthreads();
        dlogits = s_dl;
    }
    float mx = -INFINITY;
    for (int i = tid; i < N; i += 256) mx = fmaxf(mx, a_raw[i]);
    const float gmax = block_reduce_max_128(mx, red);
    float sm = 0.f;
    for (int i = tid; i < N; i += 256) sm += expf(a_raw[i] - gmax);
    const float total = block_reduce_sum_128(sm, red);
    float part = 0.f;
    for (int j = tid; j < L1; j += 256) {
        float dm = dM_ext ? dM_ext[j] : 0.f;
        for (int c = 0; c < C; ++c) dm = fmaf(Wcls[c * L1 + j], dlogits[c], dm);
        ctx[4 + j] = dm;
        part = fmaf(dm, M[j], part);
    }
    const float s = block_reduce_sum_128(part, red);
    if (tid == 0) { ctx[0] = gmax; ctx[1] = (N > 0) ? 1.0f / total : 0.f; ctx[2] = s; }
    for (int idx = tid; idx < C * L1; idx += 256) { const int c = idx / L1; g.p[8][idx] = dlogits[c] * M[idx - c * L1]; }
    for (int c = tid; c < C; c += 256) g.p[9][c] = dlogits[c];
    const int sizes[8] = {L1 * 192, L1, D * L1, D, D * L1, D, D, 1};
    for (int k = 0; k < 8; ++k)
        for (int idx = tid; idx < sizes[k]; idx += 256) g.p[k][idx] = 0.f;
}

// One training trial of the fused multi-trial step (hb_clam_sb_train_step_trials): its bag, weights, gradient buffers,
// forward outputs and dropout state.  Eight of them fit the 4 KB kernel-parameter space.
struct BwdTrial {
    const float* feats; const float* a_raw; const float* M; const float* logits; const long long* label; float* loss; float* ctx;
    int N;
    ClamModel w; ClamGrads g; ClamDrop dr;
};
struct BwdTrials { BwdTrial t[8]; };

__global__ void __launch_bounds__(256) clam_bwd_prep_kernel(const float* __restrict__ a_raw, int N, const float* __restrict__ M,
                                                            const float* __restrict__ dlogits, const float* __restrict__ dM_ext,
                                                            const float* __restrict__ Wcls, const __grid_constant__ ClamGrads g,
                                                            int L1, int D, int C, float* __restrict__ ctx,
                                                            const float* __restrict__ logits, const long long* __restrict__ label,
                                                            float* __restrict__ loss_out) {
    clam_bwd_prep_body(a_raw, N, M, dlogits, dM_ext, Wcls, g, L1, D, C, ctx, logits, label, loss_out);
}
// one CTA per trial: cross-entropy of the trial's logits against its label, softmax statistics, classifier gradients
__global__ void __launch_bounds__(256) clam_bwd_prep_trials_kernel(const __grid_constant__ BwdTrials tr, int L1, int D, int C) {
    const BwdTrial& t = tr.t[blockIdx.x];
    clam_bwd_prep_body(t.a_raw, t.N, t.M, nullptr, nullptr, t.w.p[8], t.g, L1, D, C, t.ctx, t.logits, t.label, t.loss);
}

template <int TN, int L1T>
__device__ __forceinline__ void clam_bwd192_body(const float* __restrict__ feats, int N, const ClamModel& w,
                                                 const float* __restrict__ a_raw, const float* __restrict__ dA_ext,
                                                 const float* __restrict__ ctx, const ClamGrads& g, int L1_rt, int D_rt, int ch,
                                                 const ClamDrop& dr, int chunk) {
    const int L1 = L1T ? L1T : L1_rt, D = L1T ? L1T / 2 : D_rt;
    extern __shared__ __align__(16) float smem_clam[];
    const int ldh = L1 + 4, ldp = 2 * D + 1;
    float* sX = smem_clam;                                   // [ch][196]   ch = 64 instances, 32 for the large heads
    float* sW = sX + ch * CL_XS;                          // [L1][KC + 4]  W1 slice; later dz [64][L1 + 4]
    float* sH = sW + clam_sw_floats(L1);                     // [ch][L1 + 4]
    float* sG = sH + ch * ldh;                            // [2][D][L1]    Wa, Wb
    float* sAB = sG + 2 * D * L1;                            // [64][2D + 1]  a | b
    float* sDP = sAB + ch * ldp;                          // [64][2D + 1]  dpre_a | dpre_b
    float* sDA = sDP + ch * ldp;                          // [64] dA, [64] alpha
    float* sV = sDA + 2 * ch;                             // b1 [L1] | ba [D] | bb [D] | Wc [D] | bc [1] | pad | dM [L1]
    float* sDZ = sW;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int i0 = chunk * ch;
    const int n_valid = min(ch, N - i0);
    const int nsmall = L1 + 3 * D + 1;
    float* sDM = sV + ((nsmall + 3) & ~3);

    {
        const float4* src = reinterpret_cast<const float4*>(feats + static_cast<size_t>(i0) * 192);
        const uint32_t dst0 = smem_u32(sX);
        for (int idx = tid; idx < ch * 48; idx += nthreads) {
            const int r = idx / 48, c4 = idx - r * 48;
            const int nbytes = (r < n_valid) ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (r * CL_XS + c4 * 4) * 4),
                         "l"(src + (r < n_valid ? idx : 0)), "r"(nbytes) : "memory");
        }
        for (int idx = tid; idx < D * L1 / 4; idx += nthreads) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sG + idx * 4)), "l"(w.p[2] + idx * 4) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sG + D * L1 + idx * 4)), "l"(w.p[4] + idx * 4) : "memory");
        }
        for (int idx = tid; idx < nsmall; idx += nthreads) {
            const float* src1 = idx < L1 ? w.p[1] + idx : idx < L1 + D ? w.p[3] + (idx - L1) : idx < L1 + 2 * D ? w.p[5] + (idx - L1 - D)
                              : idx < L1 + 3 * D ? w.p[6] + (idx - L1 - 2 * D) : w.p[7];
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sV + idx)), "l"(src1) : "memory");
        }
        for (int idx = tid; idx < L1; idx += nthreads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sDM + idx)), "l"(ctx + 4 + idx) : "memory");
    }
    clam_fc1_192<TN, L1T>(w.p[0], sV, L1, sX, sW, sH, ldh, ch == 64, dr, static_cast<uint32_t>(i0));   // waits for every cp.async above, ends with a __syncthreads
    const float gmax = ctx[0], inv_total = ctx[1], s_dot = ctx[2];
    const float* ba = sV + L1; const float* bb = ba + D; const float* Wc = bb + D;

    const int parts = nthreads / ch;
    const int inst = tid % ch, part = tid / ch;
    const int dn = D / parts, d0 = part * dn;
    // ---- gate recomputation: a, b of this thread's units; part 0 also forms alpha_i and dA_i
    {
        const uint32_t h_addr = smem_u32(sH + inst * ldh);
        for (int d = d0; d < d0 + dn; ++d) {
            float a0 = ba[d], a1 = 0.f, b0 = bb[d], b1v = 0.f;
            const uint32_t wa_addr = smem_u32(sG + d * L1), wb_addr = smem_u32(sG + (D + d) * L1);
#pragma unroll 4
            for (int j = 0; j < L1; j += 4) {
                const float4 h4 = lds_f4(h_addr + j * 4);
                const float4 wa = lds_f4(wa_addr + j * 4), wb = lds_f4(wb_addr + j * 4);
                a0 = fmaf(wa.x, h4.x, a0); a1 = fmaf(wa.y, h4.y, a1); a0 = fmaf(wa.z, h4.z, a0); a1 = fmaf(wa.w, h4.w, a1);
                b0 = fmaf(wb.x, h4.x, b0); b1v = fmaf(wb.y, h4.y, b1v); b0 = fmaf(wb.z, h4.z, b0); b1v = fmaf(wb.w, h4.w, b1v);
            }
            sAB[inst * ldp + d] = tanhf(a0 + a1);                            // pre-dropout branch outputs
            sAB[inst * ldp + D + d] = 1.0f / (1.0f + expf(-(b0 + b1v)));
        }
        if (part == 0) {
            float dot = 0.f;
            for (int j = 0; j < L1; ++j) dot = fmaf(sDM[j], sH[inst * ldh + j], dot);
            float alpha = 0.f, dA = 0.f;
            if (inst < n_valid) {
                alpha = expf(a_raw[i0 + inst] - gmax) * inv_total;
                dA = alpha * (dot - s_dot) + (dA_ext ? dA_ext[i0 + inst] : 0.f);
            }
            sDA[inst] = dA;
            sDA[ch + inst] = alpha;
        }
    }
    __syncthreads();
    // ---- dpre_a, dpre_b
    {
        const float dA = sDA[inst];
        for (int d = d0; d < d0 + dn; ++d) {
            const float a = sAB[inst * ldp + d], b = sAB[inst * ldp + D + d];
            const float ka = clam_keep(dr, i0 + inst, L1 + d), kb = clam_keep(dr, i0 + inst, L1 + D + d);
            const float t = dA * Wc[d] * ka * kb;                             // A = sum_d Wc_d (ka a)(kb b)
            sDP[inst * ldp + d] = t * b * (1.0f - a * a);
            sDP[inst * ldp + D + d] = t * a * b * (1.0f - b);
        }
    }
    __syncthreads();
    // ---- dz = (alpha dM + Wa^T dpre_a + Wb^T dpre_b) [h1 > 0]     (sDZ aliases the W1 staging buffer: fc1 is done)
    {
        const float alpha = sDA[ch + inst];
        const int jn = L1 / parts, j0 = part * jn;
        for (int j = j0; j < j0 + jn; ++j) {
            float acc = alpha * sDM[j];
            for (int d = 0; d < D; ++d) {
                acc = fmaf(sG[d * L1 + j], sDP[inst * ldp + d], acc);
                acc = fmaf(sG[(D + d) * L1 + j], sDP[inst * ldp + D + d], acc);
            }
            // h1 = keep * relu(z): a dropped unit has h1 = 0 and no gradient, a kept one carries the 1 / (1 - p) scale
            sDZ[inst * ldh + j] = (sH[inst * ldh + j] > 0.f) ? acc * clam_keep(dr, i0 + inst, j) : 0.f;
        }
    }
    __syncthreads();
    // ---- parameter gradients of this chunk -> global (atomics)
    // dW1 [L1][192]: thread tile = 8 output rows x 4 columns, loop over the 64 instances
    for (int task = tid; task < 48 * (L1 / 8); task += nthreads) {
        const int jg = task / 48, kq = task - jg * 48;
        float4 acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t x_addr = smem_u32(sX + kq * 4), z_addr = smem_u32(sDZ + jg * 8);
        for (int i = 0; i < n_valid; ++i) {
            const float4 x4 = lds_f4(x_addr + i * CL_XS * 4);
            const float4 z0 = lds_f4(z_addr + i * ldh * 4), z1 = lds_f4(z_addr + i * ldh * 4 + 16);
            const float z[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc[q].x = fmaf(z[q], x4.x, acc[q].x); acc[q].y = fmaf(z[q], x4.y, acc[q].y);
                acc[q].z = fmaf(z[q], x4.z, acc[q].z); acc[q].w = fmaf(z[q], x4.w, acc[q].w);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float* dst = g.p[0] + static_cast<size_t>(jg * 8 + q) * 192 + kq * 4;
            atomicAdd(dst, acc[q].x); atomicAdd(dst + 1, acc[q].y); atomicAdd(dst + 2, acc[q].z); atomicAdd(dst + 3, acc[q].w);
        }
    }
    // dWa, dWb [D][L1]
    for (int task = tid; task < 2 * D * L1; task += nthreads) {
        const int which = task / (D * L1), rem = task - which * D * L1;
        const int d = rem / L1, j = rem - d * L1;
        float acc = 0.f;
        for (int i = 0; i < n_valid; ++i) acc = fmaf(sDP[i * ldp + which * D + d], sH[i * ldh + j], acc);
        atomicAdd(g.p[which ? 4 : 2] + rem, acc);
    }
    // db1 [L1], dba, dbb, dWc [D], dbc
    for (int task = tid; task < L1 + 3 * D + 1; task += nthreads) {
        float acc = 0.f;
        if (task < L1) {
            for (int i = 0; i < n_valid; ++i) acc += sDZ[i * ldh + task];
            atomicAdd(g.p[1] + task, acc);
        } else if (task < L1 + 2 * D) {
            const int d2 = task - L1;                        // 0..D-1: dba, D..2D-1: dbb
            for (int i = 0; i < n_valid; ++i) acc += sDP[i * ldp + d2];
            atomicAdd(d2 < D ? g.p[3] + d2 : g.p[5] + (d2 - D), acc);
        } else if (task < L1 + 3 * D) {
            const int d = task - L1 - 2 * D;
            for (int i = 0; i < n_valid; ++i)
                acc = fmaf(sDA[i] * clam_keep(dr, i0 + i, L1 + d) * clam_keep(dr, i0 + i, L1 + D + d),
                           sAB[i * ldp + d] * sAB[i * ldp + D + d], acc);
            atomicAdd(g.p[6] + d, acc);
        } else {
            for (int i = 0; i < n_valid; ++i) acc += sDA[i];
            atomicAdd(g.p[7], acc);
        }
    }
}

template <int TN, int L1T>
__global__ void __launch_bounds__(256) clam_bwd192_kernel(const float* __restrict__ feats, int N,
                                                          const __grid_constant__ ClamModel w, const float* __restrict__ a_raw,
                                                          const float* __restrict__ dA_ext, const float* __restrict__ ctx,
                                                          const __grid_constant__ ClamGrads g, int L1_rt, int D_rt, int ch,
                                                          const ClamDrop dr) {
    clam_bwd192_body<TN, L1T>(feats, N, w, a_raw, dA_ext, ctx, g, L1_rt, D_rt, ch, dr, blockIdx.x);
}
// grid (chunks of the longest bag, trials): every trial's recomputing backward in one launch
template <int TN, int L1T>
__global__ void __launch_bounds__(256) clam_bwd192_trials_kernel(const __grid_constant__ BwdTrials tr, int L1_rt, int D_rt, int ch) {
    const BwdTrial& t = tr.t[blockIdx.y];
    if (static_cast<int>(blockIdx.x) * ch >= t.N) return;
    clam_bwd192_body<TN, L1T>(t.feats, t.N, t.w, t.a_raw, nullptr, t.ctx, t.g, L1_rt, D_rt, ch, t.dr, blockIdx.x);
}

int clam_backward_launch(const float* feats, int N, const void* const* weights_host, const float* a_raw, const float* M,
                         const float* dlogits, const float* dM_ext, const float* dA_ext, void* const* grads_host, int L0,
                         int L1, int D, int C, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                         const float* logits = nullptr, const long long* label = nullptr, float* loss_out = nullptr,
                         float dropout_p = 0.f, unsigned long long dropout_seed = 0) {
    const ClamDrop dr = clam_drop_make(dropout_p, dropout_seed);
    if (!clam_is192(L0, L1, D)) return set_error("hb_clam_sb_backward: only 192-d features with L1 <= 128 are supported (L0=%d L1=%d D=%d)", L0, L1, D);
    if (N < 1 || C < 1 || C > 64) return set_error("hb_clam_sb_backward: bad dims N=%d C=%d", N, C);
    if (!feats || !weights_host || !a_raw || !M || !grads_host || !workspace) return set_error("hb_clam_sb_backward: null argument");
    if (!dlogits && !(logits && label)) return set_error("hb_clam_sb_backward: need dlogits, or logits and label");
    if ((reinterpret_cast<uintptr_t>(feats) & 15) != 0) return set_error("hb_clam_sb_backward: features must be 16 B aligned");
    if (workspace_bytes < (4 + static_cast<size_t>(L1)) * sizeof(float)) return set_error("hb_clam_sb_backward: workspace too small");
    ClamModel w;
    ClamGrads g;
    for (int k = 0; k < 10; ++k) {
        w.p[k] = static_cast<const float*>(weights_host[k]);
        g.p[k] = static_cast<float*>(grads_host[k]);
        if (!w.p[k] || !g.p[k]) return set_error("hb_clam_sb_backward: weight / gradient pointer %d is null", k);
    }
    float* ctx = static_cast<float*>(workspace);
    {
        ProfScope ps(12, stream);
        clam_bwd_prep_kernel<<<1, 256, 0, stream>>>(a_raw, N, M, dlogits, dM_ext, w.p[8], g, L1, D, C, ctx, dlogits ? nullptr : logits,
                                                    label, loss_out);
        count_launch();
        HB_CUDA_OK(cudaGetLastError());
    }
    const int ldp = 2 * D + 1;
    const int ch = L1 >= 64 ? 32 : CL_CH;                    // the large heads keep Wa, Wb, a, b, dpre in shared memory: smaller tile
    const int threads = L1 >= 64 ? 256 : CL_THREADS;
    if (D % (threads / ch) != 0 || L1 % (threads / ch) != 0)
        return set_error("hb_clam_sb_backward: D=%d and L1=%d must be multiples of %d", D, L1, threads / ch);
    const size_t smem = (static_cast<size_t>(ch) * CL_XS + clam_sw_floats(L1) + static_cast<size_t>(ch) * (L1 + 4) +
                         2 * static_cast<size_t>(D) * L1 + 2 * static_cast<size_t>(ch) * ldp + 2 * ch +
                         ((L1 + 3 * D + 1 + 3) & ~3) + L1 + 8) * sizeof(float);
    if (smem > 220 * 1024) return set_error("hb_clam_sb_backward: shared memory %zu too large", smem);
    auto kern = (L1 <= 16) ? clam_bwd192_kernel<4, 0> : clam_bwd192_kernel<8, 0>;
    if (D * 2 == L1) {
        if (L1 == 8) kern = clam_bwd192_kernel<4, 8>;
        else if (L1 == 16) kern = clam_bwd192_kernel<4, 16>;
        else if (L1 == 32) kern = clam_bwd192_kernel<8, 32>;
        else if (L1 == 64) kern = clam_bwd192_kernel<8, 64>;
        else if (L1 == 128) kern = clam_bwd192_kernel<8, 128>;
    }
    if (set_max_dynamic_smem(reinterpret_cast<const void*>(kern), 220 * 1024)) return -1;
    ProfScope ps(13, stream);
    kern<<<(N + ch - 1) / ch, threads, smem, stream>>>(feats, N, w, a_raw, dA_ext, ctx, g, L1, D, ch, dr);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-trial training step: T independent trials — the reference packs five Ray Tune trials on
// one GPU as five processes, each running train_loop (utils/core_utils.py:384-426) one bag at a time — advance one step
// each in SIX launches: work table, paired scores (trial t's weights on trial t's bag), paired combine, one prep CTA per
// trial (cross-entropy + its gradient), the recomputing backward over a (chunk, trial) grid, and one Adam launch over all
// 10 T tensors with per-trial learning rate / weight decay / step count.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ADAM_TRIALS = 8;
struct AdamTrialTensors {
    float* p[10 * ADAM_TRIALS]; const float* g[10 * ADAM_TRIALS]; float* m[10 * ADAM_TRIALS]; float* v[10 * ADAM_TRIALS];
    int end[10 * ADAM_TRIALS];
    float lr_bc1[ADAM_TRIALS], bc2_sqrt[ADAM_TRIALS], wd[ADAM_TRIALS];
};
static_assert(sizeof(AdamTrialTensors) <= 4000, "kernel parameter space");

__global__ void __launch_bounds__(256) adam_trials_kernel(const __grid_constant__ AdamTrialTensors t, int n_tensors, int total,
                                                          float b1, float b2, float eps) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int k = 0;
        while (k < n_tensors - 1 && idx >= t.end[k]) ++k;
        const int off = idx - (k ? t.end[k - 1] : 0);
        const int tr = k / 10;
        const float p = t.p[k][off];
        const float g = t.g[k][off] + t.wd[tr] * p;
        const float m = b1 * t.m[k][off] + (1.0f - b1) * g;
        const float v = b2 * t.v[k][off] + (1.0f - b2) * g * g;
        t.m[k][off] = m;
        t.v[k][off] = v;
        t.p[k][off] = p - t.lr_bc1[tr] * m / (sqrtf(v) / t.bc2_sqrt[tr] + eps);
    }
}

int clam_train_trials_launch(const float* feats, const int32_t* bag_offsets_dev, const int32_t* bag_offsets_host, int n_trials,
                             const void* const* weights_host, void* const* grads_host, void* const* exp_avg_host,
                             void* const* exp_avg_sq_host, const long long* labels_dev, const float* lr, const float* weight_decay,
                             const int* step, float beta1, float beta2, float eps, float dropout_p,
                             const unsigned long long* dropout_seeds, float* a_raw, float* m_pooled, float* logits, float* loss,
                             int L0, int L1, int D, int C, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n_trials < 1 || n_trials > ADAM_TRIALS) return set_error("hb_clam_sb_train_step_trials: 1..%d trials per launch", ADAM_TRIALS);
    if (!clam_is192(L0, L1, D)) return set_error("hb_clam_sb_train_step_trials: HIPT heads only (192-d features, L1 <= 128)");
    if (!feats || !bag_offsets_dev || !bag_offsets_host || !weights_host || !grads_host || !exp_avg_host || !exp_avg_sq_host ||
        !labels_dev || !lr || !weight_decay || !step || !dropout_seeds || !a_raw || !m_pooled || !logits || !loss || !workspace)
        return set_error("hb_clam_sb_train_step_trials: null argument");
    int max_len = 0;
    for (int t = 0; t < n_trials; ++t) {
        const int len = bag_offsets_host[t + 1] - bag_offsets_host[t];
        if (len < 1) return set_error("hb_clam_sb_train_step_trials: trial %d has an empty bag", t);
        if (step[t] < 1) return set_error("hb_clam_sb_train_step_trials: step counts from 1");
        max_len = len > max_len ? len : max_len;
    }
    const int total = bag_offsets_host[n_trials];
    // workspace: [forward workspace][n_trials x (4 + L1) floats of backward context]
    const size_t fwd_ws = (clam_workspace_bytes(max_len, n_trials, n_trials, L1) + 15) & ~static_cast<size_t>(15);
    const size_t need = fwd_ws + static_cast<size_t>(n_trials) * (4 + L1) * sizeof(float);
    if (workspace_bytes < need) return set_error("hb_clam_sb_train_step_trials: workspace %zu < %zu bytes", workspace_bytes, need);
    // ---- forward: launches 1-3
    if (clam_forward_launch(feats, bag_offsets_dev, n_trials, total, max_len, weights_host, n_trials, L0, L1, D, C, a_raw, m_pooled,
                            logits, nullptr, nullptr, workspace, fwd_ws, stream, dropout_p, 0ull, dropout_seeds)) return -1;
    // ---- backward: launches 4-5
    BwdTrials tr;
    memset(&tr, 0, sizeof(tr));
    float* ctx0 = reinterpret_cast<float*>(static_cast<char*>(workspace) + fwd_ws);
    for (int t = 0; t < n_trials; ++t) {
        BwdTrial& b = tr.t[t];
        b.feats = feats + static_cast<size_t>(bag_offsets_host[t]) * 192;
        b.a_raw = a_raw + bag_offsets_host[t];
        b.M = m_pooled + static_cast<size_t>(t) * L1;
        b.logits = logits + static_cast<size_t>(t) * C;
        b.label = labels_dev + t;
        b.loss = loss + t;
        b.ctx = ctx0 + static_cast<size_t>(t) * (4 + L1);
        b.N = bag_offsets_host[t + 1] - bag_offsets_host[t];
        b.dr = clam_drop_make(dropout_p, dropout_seeds[t]);
        for (int k = 0; k < 10; ++k) {
            b.w.p[k] = static_cast<const float*>(weights_host[t * 10 + k]);
            b.g.p[k] = static_cast<float*>(grads_host[t * 10 + k]);
            if (!b.w.p[k] || !b.g.p[k]) return set_error("hb_clam_sb_train_step_trials: weight / gradient pointer %d of trial %d is null", k, t);
        }
    }
    {
        ProfScope ps(12, stream);
        clam_bwd_prep_trials_kernel<<<n_trials, 256, 0, stream>>>(tr, L1, D, C);
        count_launch();
        HB_CUDA_OK(cudaGetLastError());
    }
    {
        const int ldp = 2 * D + 1;
        const int ch = L1 >= 64 ? 32 : CL_CH;
        const int threads = L1 >= 64 ? 256 : CL_THREADS;
        if (D % (threads / ch) != 0 || L1 % (threads / ch) != 0)
            return set_error("hb_clam_sb_train_step_trials: D=%d and L1=%d must be multiples of %d", D, L1, threads / ch);
        const size_t smem = (static_cast<size_t>(ch) * CL_XS + clam_sw_floats(L1) + static_cast<size_t>(ch) * (L1 + 4) +
                             2 * static_cast<size_t>(D) * L1 + 2 * static_cast<size_t>(ch) * ldp + 2 * ch +
                             ((L1 + 3 * D + 1 + 3) & ~3) + L1 + 8) * sizeof(float);
        auto kern = (L1 <= 16) ? clam_bwd192_trials_kernel<4, 0> : clam_bwd192_trials_kernel<8, 0>;
        if (D * 2 == L1) {
            if (L1 == 8) kern = clam_bwd192_trials_kernel<4, 8>;
            else if (L1 == 16) kern = clam_bwd192_trials_kernel<4, 16>;
            else if (L1 == 32) kern = clam_bwd192_trials_kernel<8, 32>;
            else if (L1 == 64) kern = clam_bwd192_trials_kernel<8, 64>;
            else if (L1 == 128) kern = clam_bwd192_trials_kernel<8, 128>;
        }
        if (set_max_dynamic_smem(reinterpret_cast<const void*>(kern), 220 * 1024)) return -1;
        ProfScope ps(13, stream);
        kern<<<dim3((max_len + ch - 1) / ch, n_trials), threads, smem, stream>>>(tr, L1, D, ch);
        count_launch();
        HB_CUDA_OK(cudaGetLastError());
    }
    // ---- Adam: launch 6
    AdamTrialTensors at;
    memset(&at, 0, sizeof(at));
    const int numel[10] = {L1 * 192, L1, D * L1, D, D * L1, D, D, 1, C * L1, C};
    int tot = 0;
    for (int t = 0; t < n_trials; ++t) {
        for (int k = 0; k < 10; ++k) {
            const int i = t * 10 + k;
            at.p[i] = static_cast<float*>(const_cast<void*>(weights_host[i]));
            at.g[i] = static_cast<const float*>(grads_host[i]);
            at.m[i] = static_cast<float*>(exp_avg_host[i]);
            at.v[i] = static_cast<float*>(exp_avg_sq_host[i]);
            if (!at.m[i] || !at.v[i]) return set_error("hb_clam_sb_train_step_trials: Adam state pointer %d of trial %d is null", k, t);
            tot += numel[k];
            at.end[i] = tot;
        }
        const float bc1 = 1.0f - powf(beta1, static_cast<float>(step[t]));
        at.lr_bc1[t] = lr[t] / bc1;
        at.bc2_sqrt[t] = sqrtf(1.0f - powf(beta2, static_cast<float>(step[t])));
        at.wd[t] = weight_decay[t];
    }
    int grid = (tot + 255) / 256;
    if (grid > 4 * num_sms()) grid = 4 * num_sms();
    ProfScope ps(14, stream);
    adam_trials_kernel<<<grid, 256, 0, stream>>>(at, n_trials * 10, tot, beta1, beta2, eps);
    count_launch();
    HB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-tensor Adam with L2 weight decay, torch.optim.Adam semantics (utils/utils.py:100-107 get_optim):
//   g += wd p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ADAM_MAX_TENSORS = 16;
struct AdamTensors { float* p[ADAM_MAX_TENSORS]; const float* g[ADAM_MAX_TENSORS]; float* m[ADAM_MAX_TENSORS]; float* v[ADAM_MAX_TENSORS];
                     int end[ADAM_MAX_TENSORS]; };

__global__ void __launch_bounds__(256) adam_step_kernel(const __grid_constant__ AdamTensors t, int n_tensors, int total, float lr,
                                                        float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int k = 0;
        while (k < n_tensors - 1 && idx >= t.end[k]) ++k;
        const int off = idx - (k ? t.end[k - 1] : 0);
        const float p = t.p[k][off];
        const float g = t.g[k][off] + wd * p;
        const float m = b1 * t.m[k][off] + (1.0f - b1) * g;
        const float v = b2 * t.v[k][off] + (1.0f - b2) * g * g;
        t.m[k][off] = m;
        t.v[k][off] = v;
        t.p[k][off] = p - (lr / bc1) * m / (sqrtf(v) / bc2_sqrt + eps);
    }
}

int adam_step_launch(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                     const int* numel, int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay,
                     int step, cudaStream_t stream) {
    if (n_tensors < 1 || n_tensors > ADAM_MAX_TENSORS) return set_error("hb_adam_step: n_tensors must be 1..%d", ADAM_MAX_TENSORS);
    if (step < 1) return set_error("hb_adam_step: step counts from 1");
    AdamTensors t;
    int total = 0;
    for (int k = 0; k < n_tensors; ++k) {
        if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k] || numel[k] < 0) return set_error("hb_adam_step: bad tensor %d", k);
        t.p[k] = static_cast<float*>(params[k]); t.g[k] = static_cast<const float*>(grads[k]);
        t.m[k] = static_cast<float*>(exp_avg[k]); t.v[k] = static_cast<float*>(exp_avg_sq[k]);
        total += numel[k];
        t.end[k] = total;
    }
    if (total == 0) return 0;
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
    int grid = (total + 255) / 256;
    if (grid > 4 * num_sms()) grid = 4 * num_sms();
    ProfScope ps(14, stream);
    adam_step_kernel<<<grid, 256, 0, stream>>>(t, n_tensors, total, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt);
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
                                   workspace, workspace_bytes, static_cast<cudaStream_t>(stream), 0.f, 0ull);
}
int hb_clam_sb_forward_train(const float* feats, const int32_t* bag_offsets, int n_bags, int total_instances,
                             int max_bag_len, const void* const* weights_host, int n_models, int L0, int L1, int D, int C,
                             float* a_raw, float* m_out, float* logits, float* y_prob, int64_t* y_hat, void* workspace,
                             size_t workspace_bytes, float dropout_p, uint64_t dropout_seed, void* stream) {
    if (!(dropout_p >= 0.f && dropout_p <= 1.f)) return hb::set_error("hb_clam_sb_forward_train: dropout_p %f outside [0, 1]", dropout_p);
    return hb::clam_forward_launch(feats, bag_offsets, n_bags, total_instances, max_bag_len, weights_host, n_models, L0,
                                   L1, D, C, a_raw, m_out, logits, y_prob, reinterpret_cast<long long*>(y_hat),
                                   workspace, workspace_bytes, static_cast<cudaStream_t>(stream), dropout_p, dropout_seed);
}
int hb_clam_sb_backward_train(const float* feats, int n_instances, const void* const* weights_host, const float* a_raw,
                              const float* m_pooled, const float* dlogits, const float* dm_ext, const float* da_ext,
                              const float* logits, const int64_t* label, float* loss_out, void* const* grads_host, int L0,
                              int L1, int D, int C, void* workspace, size_t workspace_bytes, float dropout_p,
                              uint64_t dropout_seed, void* stream) {
    if (!(dropout_p >= 0.f && dropout_p <= 1.f)) return hb::set_error("hb_clam_sb_backward_train: dropout_p %f outside [0, 1]", dropout_p);
    return hb::clam_backward_launch(feats, n_instances, weights_host, a_raw, m_pooled, dlogits, dm_ext, da_ext, grads_host, L0,
                                    L1, D, C, workspace, workspace_bytes, static_cast<cudaStream_t>(stream),
                                    dlogits ? nullptr : logits, reinterpret_cast<const long long*>(label), loss_out, dropout_p,
                                    dropout_seed);
}
int hb_clam_dropout_masks(int n_instances, int L1, int D, float dropout_p, uint64_t dropout_seed, float* m1_host,
                          float* ma_host, float* mb_host) {
    if (n_instances < 0 || L1 < 1 || D < 1 || !m1_host || !ma_host || !mb_host) return hb::set_error("hb_clam_dropout_masks: bad argument");
    const hb::ClamDrop dr = hb::clam_drop_make(dropout_p, dropout_seed);
    for (int i = 0; i < n_instances; ++i) {
        for (int j = 0; j < L1; ++j) m1_host[static_cast<size_t>(i) * L1 + j] = hb::clam_keep(dr, i, j);
        for (int d = 0; d < D; ++d) {
            ma_host[static_cast<size_t>(i) * D + d] = hb::clam_keep(dr, i, L1 + d);
            mb_host[static_cast<size_t>(i) * D + d] = hb::clam_keep(dr, i, L1 + D + d);
        }
    }
    return 0;
}
int hb_clam_sb_backward(const float* feats, int n_instances, const void* const* weights_host, const float* a_raw,
                        const float* m_pooled, const float* dlogits, const float* dm_ext, const float* da_ext,
                        void* const* grads_host, int L0, int L1, int D, int C, void* workspace, size_t workspace_bytes,
                        void* stream) {
    return hb::clam_backward_launch(feats, n_instances, weights_host, a_raw, m_pooled, dlogits, dm_ext, da_ext, grads_host,
                                    L0, L1, D, C, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
size_t hb_clam_trials_workspace_bytes(int max_bag_len, int n_trials, int L1) {
    return ((hb::clam_workspace_bytes(max_bag_len, n_trials, n_trials, L1) + 15) & ~static_cast<size_t>(15)) +
           static_cast<size_t>(n_trials) * (4 + L1) * sizeof(float);
}
int hb_clam_sb_train_step_trials(const float* feats, const int32_t* bag_offsets, const int32_t* bag_offsets_host, int n_trials,
                                 const void* const* weights_host, void* const* grads_host, void* const* exp_avg_host,
                                 void* const* exp_avg_sq_host, const int64_t* labels, const float* lr, const float* weight_decay,
                                 const int* step, float beta1, float beta2, float eps, float dropout_p,
                                 const uint64_t* dropout_seeds, float* a_raw, float* m_pooled, float* logits, float* loss, int L0,
                                 int L1, int D, int C, void* workspace, size_t workspace_bytes, void* stream) {
    return hb::clam_train_trials_launch(feats, bag_offsets, bag_offsets_host, n_trials, weights_host, grads_host, exp_avg_host,
                                        exp_avg_sq_host, reinterpret_cast<const long long*>(labels), lr, weight_decay, step, beta1,
                                        beta2, eps, dropout_p, reinterpret_cast<const unsigned long long*>(dropout_seeds), a_raw,
                                        m_pooled, logits, loss, L0, L1, D, C, workspace, workspace_bytes,
                                        static_cast<cudaStream_t>(stream));
}
int hb_adam_step(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                 const int* numel, int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                 void* stream) {
    return hb::adam_step_launch(params, grads, exp_avg, exp_avg_sq, numel, n_tensors, lr, beta1, beta2, eps, weight_decay,
                                step, static_cast<cudaStream_t>(stream));
}
int hb_clam_sb_backward_ce(const float* feats, int n_instances, const void* const* weights_host, const float* a_raw,
                           const float* m_pooled, const float* logits, const int64_t* label, float* loss_out,
                           void* const* grads_host, int L0, int L1, int D, int C, void* workspace, size_t workspace_bytes,
                           void* stream) {
    return hb::clam_backward_launch(feats, n_instances, weights_host, a_raw, m_pooled, nullptr, nullptr, nullptr, grads_host, L0,
                                    L1, D, C, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), logits,
                                    reinterpret_cast<const long long*>(label), loss_out);
}
}
