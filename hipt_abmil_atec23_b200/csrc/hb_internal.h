// hb_internal.h — host-side plumbing shared by the translation units of libhipt_b200.so (not part of the C-ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hb {

int set_error(const char* fmt, ...);          // records the message for hb_last_error(), returns -1
int num_sms();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, so a process that
// drives several GPUs (HIPT_4K(device256=cuda:0, device4k=cuda:1), nn.DataParallel replicas) opts in on each of them
int set_max_dynamic_smem(const void* func, int bytes);

#define HB_CUDA_OK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return ::hb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

void count_launch();                           // every kernel launch of the library bumps hb_launch_count()
struct ProfScope {                             // CUDA-event bracket on the launching stream when profiling is on
    ProfScope(int kind, cudaStream_t st);
    ~ProfScope();
    cudaStream_t st_;
    long long idx_;
    cudaEvent_t a_, b_;
};

enum TmapDtype { TMAP_BF16 = 0, TMAP_F32 = 1, TMAP_U8 = 2 };
// 2-D row-major tensor [rows, cols] with a row pitch in bytes; box [box_rows, box_cols]; SWIZZLE_128B or _64B
// (box_cols * elem_size must equal the swizzle span).  Out-of-bounds elements read as zero and are not written.
int encode_tmap_2d(CUtensorMap* map, TmapDtype dt, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                   uint32_t box_rows, uint32_t box_cols, int swizzle_bytes = 128);

// n-D uint8 tensor (dims fastest first, strides in bytes for dims 1..n-1), no swizzle: dense box in shared memory
int encode_tmap_u8_nd(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);
// fused unfold + normalise + patch embed from uint8 regions (hb_embed.cu)
int embed_u8_launch(const void* image_u8, size_t chan_stride, size_t row_pitch, int grid_cols, int grid_rows,
                    size_t image_stride_bytes, int n_images, int patch_begin, int n_patches, const void* w_f16,
                    const float* bias, float scale, const float* pos_table, void* xb_bf16, float* stats, int stats_stride,
                    cudaStream_t stream);

struct GemmAux {                                 // epilogue side inputs / outputs, passed to the kernel by value
    const float* colvec2;                        // LNFOLD: c[N] = sum_k gamma_k W_jk (bias slot carries d[N])
    const float* row_stats;                      // LNFOLD: [n_part][stats_stride][2] partial (sum, sum of squares) of the rows behind A
    float* stats_out;                            // RESID / TOKENS: (sum, sum of squares) of the produced rows, accumulated
    float* stats_clear;                          // RESID: the other LayerNorm's statistics array, zeroed on the way
    __nv_bfloat16* xb_out;                       // TOKENS: bf16 copy of the produced rows (RESID stores it by TMA)
    float inv_dim, eps;                          // LNFOLD: 1 / normalised width, LayerNorm epsilon
    int n_part;                                  // LNFOLD: partial sums per row (1..6); RESID_BF16 writes one per 64 columns
    int stats_stride;                            // rows between two partial-sum planes
};

struct GemmArgs {
    CUtensorMap map_a, map_w, map_w2, map_out, map_xb;   // map_w2: W with a half-height box for the CTA-pair kernel
    GemmAux aux;
    int cg2;                                     // 1: launch as clusters of 2 CTAs (tcgen05 cta_group::2)
    int astat;                                   // 1: A-stationary variant (K = 384): the A tile is loaded once per group of n-tiles
    const float* bias;
    const float* tok_table;
    float* tok_out;
    int M, N, K, bn, epi, tokens_per_seq;
};
int gemm_pick_bn(int N);
int gemm_prepare(GemmArgs& g, const void* A, const void* W, const float* bias, int epi, void* out, int M, int N, int K,
                 const float* tok_table, int tokens_per_seq, const GemmAux* aux = nullptr, void* xb_out = nullptr,
                 size_t out_pitch_bytes = 0);
// RESID_BF16: `out` is the destination [M, N] bf16 (dense), xb_out the residual source (row pitch out_pitch_bytes, 0 = dense;
// may alias `out`), aux->stats_out the [N/64][stats_stride][2] partial row statistics.
int gemm_launch(const GemmArgs& g, cudaStream_t stream);
// fused fc1 -> GELU -> fc2 -> residual for dim 384 / hidden 1536 (hb_mlp.cu); xb is updated in place
int mlp_fused_launch(const void* xb_bf16, const void* w1g_bf16, const float* c1, const float* d1, const void* w2h_bf16,
                     const float* b2, const float* stats_in, float* stats_out, int stats_stride, float eps, int M,
                     cudaStream_t stream);

// elementwise / row kernels
int layernorm_launch(const void* x, int x_is_bf16, size_t x_row_stride, const float* gamma, const float* beta, float eps,
                     void* out_bf16, float* out_f32, int rows, int dim, cudaStream_t stream);
int attention_launch(const void* qkv_bf16, void* out_bf16, int n_seq, int seq_len, int heads, int head_dim, float scale,
                     cudaStream_t stream, int cls_only = 0, float* cls_probs = nullptr);
// tcgen05 attention for seq_len 257 / head_dim 64 (hb_attention_tc.cu)
int attention_tc2_launch(const void* qkv_bf16, void* out_bf16, int n_seq, int heads, float scale, cudaStream_t stream);
int im2col_launch(const void* image, int image_is_f32, size_t patch_stride, size_t chan_stride, size_t row_pitch,
                  int grid_cols, int patch_begin, int n_patches, void* a_bf16, cudaStream_t stream);
int cls_rows_launch(const float* cls_token, const float* pos_table, float* x, void* xb_bf16, float* stats, int n_seq,
                    int seq_len, int dim, cudaStream_t stream);

int clam_forward_launch(const float* feats, const int32_t* bag_offsets, int n_bags, int total_instances,
                        int max_bag_len, const void* const* weights_host, int n_models, int L0, int L1, int D, int C,
                        float* a_raw, float* m_out, float* logits, float* y_prob, long long* y_hat, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream, float dropout_p = 0.f, unsigned long long dropout_seed = 0,
                        const unsigned long long* paired_seeds = nullptr);
size_t clam_workspace_bytes(int max_bag_len, int n_bags, int n_models, int L1);

}  // namespace hb
