/* hipt_b200.h — C ABI of libhipt_b200.so: the B200 (sm_100a) implementation of the HIPT_4K + CLAM_SB hot path.
 *
 * The reference (scjjb/HIPT_ABMIL_ATEC23) has no FFI: its hot path is reached through Python nn.Module.forward calls
 * that bottom out in ATen/cuBLAS/cuDNN.  Each entry point below therefore cites the reference Python call site it
 * replaces; the Python modules in hipt_abmil_atec23_b200/ (same class names, constructor arguments and state_dict
 * keys as the reference) bind these symbols with ctypes — see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host;
 * `stream` is a cudaStream_t passed as void*; every function returns 0 on success and -1 on failure, after which
 * hb_last_error() describes the failure (thread-local).  Nothing here allocates device memory: callers pass
 * workspaces sized by the *_workspace_bytes() queries.  bf16 buffers are passed as void*.
 */
#ifndef HIPT_B200_H
#define HIPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 4

/* GEMM epilogues */
#define HB_EPI_BIAS_BF16 0        /* out_bf16[M,N]  = A W^T + bias                                  */
#define HB_EPI_BIAS_GELU_BF16 1   /* out_bf16[M,N]  = gelu_erf(A W^T + bias)                        */
#define HB_EPI_BIAS_RESADD_F32 2  /* out_f32[M,N]  += A W^T + bias     (residual stream, in place)  */
#define HB_EPI_TOKENS_F32 3       /* token rows: out_f32[(r/T)*(T+1)+1+r%T, :] = A W^T + bias + table[1+r%T, :] */
#define HB_EPI_TOKENS_GELU_F32 4  /* same with gelu_erf applied before the table add               */
#define HB_EPI_BIAS_GELU_FAST_BF16 5 /* as 1 with the tanh-form GELU fitted to erf (|err| <= 3e-4 |x|), MLP hot path */
#define HB_EPI_LNFOLD_BF16 6         /* hb_gemm_lnfold_bf16 */
#define HB_EPI_LNFOLD_GELU_BF16 7    /* hb_gemm_lnfold_bf16 with gelu */
#define HB_EPI_RESID_STATS_F32 8     /* hb_gemm_resid_stats */
#define HB_EPI_LNFOLD_GELU2_BF16 9   /* hb_gemm_lnfold_bf16 with gelu = 2: TWICE the GELU (0.5 folded into the next Linear) */
#define HB_EPI_RESID_BF16 10         /* hb_gemm_resid_bf16 */

int hb_abi_version(void);
const char* hb_last_error(void);
/* sm_count / compute capability of the current device; fails unless it is sm_100 */
int hb_device_check(int* sm_count, int* cc_major, int* cc_minor);

/* Measurement hooks.  hb_launch_count: kernels launched by this library since load.  hb_prof_enable(1) brackets every
 * kernel launch of the drivers below with CUDA events on the launching stream; hb_prof_read synchronises, sums the
 * elapsed milliseconds and launch counts per kind (HB_PROF_*, + HB_PROF_4K_OFFSET for the ViT-4K plan) and clears. */
#define HB_PROF_IM2COL 0
#define HB_PROF_EMBED_GEMM 1
#define HB_PROF_CLS_ROWS 2
#define HB_PROF_LAYERNORM 3
#define HB_PROF_QKV_GEMM 4
#define HB_PROF_ATTENTION 5
#define HB_PROF_PROJ_GEMM 6
#define HB_PROF_FC1_GEMM 7
#define HB_PROF_FC2_GEMM 8
#define HB_PROF_FINAL_LN 9
#define HB_PROF_CLAM_SCORES 10
#define HB_PROF_CLAM_COMBINE 11
#define HB_PROF_MLP_FUSED 12
#define HB_PROF_4K_OFFSET 16
#define HB_PROF_KINDS 32
long long hb_launch_count(void);
int hb_prof_enable(int on);
int hb_prof_read(double* ms_by_kind, long long* count_by_kind, int n_kinds);

/* ------------------------------------------------------------------------------------------------------------------
 * Single kernels (also the units the parity tests exercise).
 * ---------------------------------------------------------------------------------------------------------------- */

/* nn.Linear (+GELU / +residual): HIPT_4K/vision_transformer.py:93-95,114-116 ; vision_transformer4k.py:169.
 * a_bf16 [M,K] row-major, w_bf16 [N,K] row-major (nn.Linear.weight layout), bias fp32 [N].
 * K % 64 == 0; N % 128 == 0 or N % 192 == 0.  tok_table / tokens_per_seq only for the TOKENS epilogues. */
int hb_gemm_bf16(const void* a_bf16, const void* w_bf16, const float* bias, int epilogue, void* out, int M, int N,
                 int K, const float* tok_table, int tokens_per_seq, void* stream);

/* LayerNorm followed by Linear (norm1 -> attn.qkv, norm2 -> mlp.fc1 [+ GELU]; vision_transformer.py:147,151) as ONE GEMM:
 * xb_bf16 [M,K] is the UN-normalised residual stream in bf16, w_gamma_bf16 [N,K] = bf16(W * gamma), c[j] = sum_k of
 * that rounded weight's row j, d[j] = sum_k beta_k W_jk + bias_j.  row_stats holds n_part (1..6) planes
 * [n_part][stats_stride][2] of partial (sum, sum of squares) of the rows (stats_stride >= M rows per plane); the planes
 * are summed per row.  out[r,j] = act(rstd_r * acc[r,j] - rstd_r * mu_r * c[j] + d[j]); gelu: 0 none, 1 GELU, 2 twice the
 * GELU (the MLP path: the consumer's weights carry the factor 0.5, which is exact in bf16). */
int hb_gemm_lnfold_bf16(const void* xb_bf16, const void* w_gamma_bf16, const float* c, const float* d,
                        const float* row_stats, int n_part, int stats_stride, float eps, int gelu, void* out_bf16, int M,
                        int N, int K, void* stream);

/* Residual update x = x + drop_path(Linear(a)) (vision_transformer.py:149,151) on the bf16 residual stream:
 * out_bf16[M,N] = bf16(float(res_bf16) + a W^T + bias), added in fp32 and rounded once.  res_bf16 rows are
 * res_pitch_bytes apart (0 = dense N*2; the CLS rows of a [n_seq, seq_len, N] stream are a strided source) and may
 * alias out_bf16 (in-place update).  stats_part [N/64][stats_stride][2]: per row and per 64-column group the (sum, sum
 * of squares) of the UNROUNDED values, every plane fully overwritten (no atomics, no zeroing needed). */
int hb_gemm_resid_bf16(const void* a_bf16, const void* w_bf16, const float* bias, const void* res_bf16,
                       size_t res_pitch_bytes, void* out_bf16, float* stats_part, int stats_stride, int M, int N, int K,
                       void* stream);

/* The same residual update on an fp32 stream with a bf16 copy (the first design of the block pipeline; kept as a
 * higher-precision building block): x_f32 [M,N] updated in place, xb_bf16 [M,N] = bf16(x), stats_out [M,2] += (sum, sum
 * of squares) of the new rows (must be zero on entry), stats_clear [M,2] (may be NULL) zeroed. */
int hb_gemm_resid_stats(const void* a_bf16, const void* w_bf16, const float* bias, float* x_f32, void* xb_bf16,
                        float* stats_out, float* stats_clear, int M, int N, int K, void* stream);

/* The MLP half of a ViT-S block as ONE kernel: x = x + fc2(GELU(fc1(norm2(x)))) (vision_transformer.py:151 with Mlp.forward
 * :98-104), dim 384, hidden 1536, on the bf16 residual stream in place; the hidden activation never leaves the SM.
 * w1_gamma_bf16 [1536,384], c1, d1: fc1 with norm2 folded in as for hb_gemm_lnfold_bf16; w2_half_bf16 [384,1536] =
 * 0.5 * fc2.weight (the GELU epilogue emits twice the GELU); stats_in / stats_out: [6][stats_stride][2] partial row
 * statistics planes (read for norm2, written for the next norm1). */
int hb_mlp_fused_bf16(void* xb_bf16, const void* w1_gamma_bf16, const float* c1, const float* d1, const void* w2_half_bf16,
                      const float* b2, const float* stats_in, float* stats_out, int stats_stride, float eps, int M,
                      int dim, int hidden, void* stream);

/* nn.LayerNorm(dim, eps): vision_transformer.py:138,142,195.  x fp32 rows at x_row_stride (elements);
 * writes out_bf16 and/or out_f32 (either may be NULL), densely packed [rows, dim].  dim in {384, 192}. */
int hb_layernorm(const float* x, size_t x_row_stride, const float* gamma, const float* beta, float eps, void* out_bf16,
                 float* out_f32, int rows, int dim, void* stream);
/* the same over bf16 rows (the final norm reads the CLS rows of the bf16 residual stream) */
int hb_layernorm_bf16(const void* x_bf16, size_t x_row_stride, const float* gamma, const float* beta, float eps,
                      void* out_bf16, float* out_f32, int rows, int dim, void* stream);

/* softmax(q k^T * scale) v per (sequence, head): vision_transformer.py:119-128.
 * qkv_bf16 [n_seq*seq_len, 3*heads*head_dim] (q|k|v, head-major); out_bf16 [n_seq*seq_len, heads*head_dim]. */
int hb_attention(const void* qkv_bf16, void* out_bf16, int n_seq, int seq_len, int heads, int head_dim, float scale,
                 void* stream);

/* unfold(2,256,256).unfold(3,256,256) + rearrange (HIPT_4K/hipt_4k.py:64-65) composed with the receptive fields of the
 * 16x16/16 patch-embed conv (vision_transformer.py:165-169): image [3, H, W] (uint8 or fp32; element strides given)
 * -> a_bf16 [n_patches*256, 768], row = patch*256 + ty*16 + tx, col = c*256 + i*16 + j.
 * grid_cols > 0: patch p is tile (p / grid_cols, p % grid_cols) of one region image (patch_stride ignored);
 * grid_cols == 0: patch p is a separate 256x256 image at image + p * patch_stride (a [B,3,256,256] batch). */
int hb_im2col_patches(const void* image, int image_is_f32, size_t patch_stride, size_t chan_stride, size_t row_pitch,
                      int grid_cols, int patch_begin, int n_patches, void* a_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * ViT encoder plans: VisionTransformer.forward (vision_transformer.py:248-253) and VisionTransformer4K.forward
 * (vision_transformer4k.py:241-246).  A plan binds the block weights and a caller-owned workspace and pre-encodes
 * the TMA descriptors of the 4*depth block GEMMs.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct hb_vit_plan hb_vit_plan;

typedef struct hb_vit_config {
    int dim;       /* 384 (ViT-256) / 192 (ViT-4K) */
    int heads;     /* 6 */
    int depth;     /* 12 / 6 */
    int mlp_dim;   /* 1536 / 768 */
    int max_rows;  /* capacity in token rows (n_seq * seq_len) */
    float ln_eps;  /* 1e-6 */
} hb_vit_config;

/* weights[]: [0] cls_token f32[dim], [1] norm.weight, [2] norm.bias, then for block i at 3+10*i (LayerNorms folded, see
 * hb_gemm_lnfold_bf16): qkv w_gamma (bf16 [3dim,dim]), qkv c, qkv d, attn.proj.weight (bf16), attn.proj.bias,
 * fc1 w_gamma (bf16 [mlp,dim]), fc1 c, fc1 d, 0.5 * mlp.fc2.weight (bf16 [dim,mlp]; the plan's fc1 epilogue emits twice the
 * GELU), mlp.fc2.bias.  Non-weight entries fp32. */
size_t hb_vit_workspace_bytes(const hb_vit_config* cfg);
int hb_vit_plan_create(const hb_vit_config* cfg, const void* const* weights_host, int n_weights, void* workspace,
                       size_t workspace_bytes, hb_vit_plan** plan_out);
void hb_vit_plan_destroy(hb_vit_plan* plan);
/* debug / test hooks: run only the first `depth_limit` blocks (<=0: all); fetch a workspace buffer
 * (1 = bf16 residual stream, 2 = qkv bf16, 3 = attention out bf16, 4 = MLP hidden bf16) */
int hb_vit_plan_set_depth_limit(hb_vit_plan* plan, int depth_limit);
int hb_vit_plan_buffer(hb_vit_plan* plan, int which, void** ptr, size_t* bytes);
/* Attention-map export for the hierarchical heatmaps (HIPT_4K/hipt_4k.py:121-164 reads only attention[:, :, 0, 1:] of
 * get_last_selfattention, vision_transformer.py:255-262 / vision_transformer4k.py:248-255): while cls_attn is non-NULL the
 * forward drivers also write the softmax row of the CLS query of the LAST block, [n_seq, heads, seq_len] fp32 (device), from
 * the fused CLS-only attention launch — no [B, heads, 257, 257] matrix, no second model pass.  NULL switches it off. */
int hb_vit_plan_set_cls_attention(hb_vit_plan* plan, float* cls_attn);

/* ViT-256 over n_patches 256x256 patches of one or more region images (HIPT_4K.forward steps 2-3, hipt_4k.py:64-70).
 * grid_cols > 0: the input is n_images-many region images, image_stride BYTES apart, each a grid of
 * patches_per_image = grid_rows * grid_cols tiles; global patch p is tile (p % patches_per_image) of image
 * (p / patches_per_image).  grid_cols == 0: patch p is a separate 256x256 image at image + p * patch_stride elements.
 * Batching two 4096x4096 regions (512 patches, 131,584 token rows = 514 CTA-pair tiles) per call fills the 148 SMs evenly.
 * embed_w_bf16 [dim, 768] / embed_b f32 [dim]: patch_embed.proj with any input normalisation folded in by the caller;
 * pos_table f32 [257, dim]: cls+pos rows after interpolate_pos_encoding (vision_transformer.py:213-233).
 * Outputs the final-LayerNorm CLS rows: cls_f32 [n_patches, dim] and/or cls_bf16 (either may be NULL). */
int hb_vit256_forward(hb_vit_plan* plan, const void* image, int image_is_f32, size_t patch_stride,
                      size_t chan_stride, size_t row_pitch, int grid_cols, int patches_per_image, size_t image_stride_bytes,
                      int patch_begin, int n_patches, const void* embed_w_bf16, const float* embed_b,
                      const float* pos_table, float* cls_f32, void* cls_bf16, void* stream);

/* hb_vit256_forward with the region unfold + ToTensor/Normalize + patch-embed conv fused into ONE tcgen05 GEMM that reads the
 * uint8 regions straight from HBM (hipt_4k.py:64-65 unfold/rearrange, hipt_model_utils.py:113-118 eval_transforms,
 * vision_transformer.py:165-170 PatchEmbed, :240-244 positional add): image_u8 = [n_images][3][grid_rows*256][grid_cols*256]
 * bytes with the given channel / row / image strides (multiples of 16 bytes); patches are numbered image-major, then row-major
 * over the region's 256 x 256 tiles.  embed_w_f16 = fp16 [384][768] of W / std (K order c, i, j), embed_b = b - sum W mean/std,
 * embed_scale = 1/255 (applied to the accumulator).  Everything else as hb_vit256_forward. */
int hb_vit256_forward_u8(hb_vit_plan* plan, const void* image_u8, size_t chan_stride, size_t row_pitch, int grid_cols,
                         int grid_rows, size_t image_stride_bytes, int n_images, int patch_begin, int n_patches,
                         const void* embed_w_f16, const float* embed_b, float embed_scale, const float* pos_table,
                         float* cls_f32, void* cls_bf16, void* stream);

/* ViT-4K over n_regions grids of tokens_per_region ViT-256 CLS tokens (hipt_4k.py:72-75; the reshape/transpose at :73
 * is the identity on token order).  cls256_bf16 [n_regions*tokens_per_region, in_dim]; phi_w_bf16 [dim, in_dim];
 * pos_table f32 [tokens_per_region+1, dim]; out_f32 [n_regions, dim]. */
int hb_vit4k_forward(hb_vit_plan* plan, const void* cls256_bf16, int n_regions, int tokens_per_region, int in_dim,
                     const void* phi_w_bf16, const float* phi_b, const float* pos_table, float* out_f32, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * CLAM_SB gated-attention MIL pooling over ragged bags (models/model_clam.py:59-64 Attn_Net_Gated.forward,
 * :147-191 CLAM_SB.forward), n_models weight sets ("folds") evaluated over the same bags in one pass.
 * feats f32 [total_instances, L0]; bag_offsets int32 [n_bags+1] (device); weights_host[m*10 + k], k =
 * fc.weight[L1,L0], fc.bias, attention_a.weight[D,L1], attention_a.bias, attention_b.weight, attention_b.bias,
 * attention_c.weight[1,D], attention_c.bias[1], classifiers.weight[C,L1], classifiers.bias[C].
 * Outputs (any may be NULL except a_raw): a_raw [n_models, total_instances] (pre-softmax scores), m_out
 * [n_models, n_bags, L1], logits [n_models, n_bags, C], y_prob [n_models, n_bags, C], y_hat int64 [n_models, n_bags].
 * ---------------------------------------------------------------------------------------------------------------- */
size_t hb_clam_workspace_bytes(int max_bag_len, int n_bags, int n_models, int L1);
int hb_clam_sb_forward(const float* feats, const int32_t* bag_offsets, int n_bags, int total_instances,
                       int max_bag_len, const void* const* weights_host, int n_models, int L0, int L1, int D, int C,
                       float* a_raw, float* m_out, float* logits, float* y_prob, int64_t* y_hat, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step around CLAM_SB.forward for one bag (replaces autograd through models/model_clam.py:147-183 inside
 * utils/core_utils.py:409-423 train_loop: loss.backward()).  192-d features, L1 <= 128, fp32.
 * Inputs: the bag, the 10 weight tensors (order of hb_clam_sb_forward), a_raw [N] and m_pooled [L1] as produced by the
 * forward, dlogits [C] = d loss / d logits, optional dm_ext [L1] (gradient arriving at M through results_dict
 * ['features']) and da_ext [N] (gradient arriving at A_raw), or NULL.
 * Output: grads_host[k] = device pointer of the gradient of weight k (same shapes), overwritten.
 * workspace: (4 + L1) floats.  Gradients are summed over 64-instance chunks with atomics (fp32, order not fixed).
 * ------------------------------------------------------------------------------------------------------------------ */
int hb_clam_sb_backward(const float* feats, int n_instances, const void* const* weights_host, const float* a_raw,
                        const float* m_pooled, const float* dlogits, const float* dm_ext, const float* da_ext,
                        void* const* grads_host, int L0, int L1, int D, int C, void* workspace, size_t workspace_bytes,
                        void* stream);

/* hb_clam_sb_backward with the loss of train_loop fused in (utils/core_utils.py:413 loss = loss_fn(logits, label) with
 * nn.CrossEntropyLoss, :423 loss.backward()): dlogits = softmax(logits) - onehot(label) is formed on the device from the forward's
 * logits [C] and the label (int64, device), loss_out (device, may be NULL) = logsumexp(logits) - logits[label]. */
int hb_clam_sb_backward_ce(const float* feats, int n_instances, const void* const* weights_host, const float* a_raw,
                           const float* m_pooled, const float* logits, const int64_t* label, float* loss_out,
                           void* const* grads_host, int L0, int L1, int D, int C, void* workspace, size_t workspace_bytes,
                           void* stream);

/* Training-mode CLAM_SB with ACTIVE dropout (the reference trains its final model at --drop_out 0.85, docs/README.md:186-193):
 * nn.Dropout(p) after the ReLU (models/model_clam.py:84-85) and inside both gate branches of Attn_Net_Gated (:50-52).
 * The keep decision of (instance, unit) is a counter-based hash of dropout_seed — nothing is stored, the recomputing
 * backward regenerates the forward's masks; kept units are scaled by 1 / (1 - p) like torch.  Instance = index inside its
 * bag; units [0,L1) ReLU outputs, [L1,L1+D) branch a, [L1+D,L1+2D) branch b.  HIPT heads only (192-d features, L1 <= 128).
 * hb_clam_sb_forward_train: hb_clam_sb_forward + (dropout_p, dropout_seed).
 * hb_clam_sb_backward_train: hb_clam_sb_backward (dlogits given) or hb_clam_sb_backward_ce (dlogits NULL: logits + label)
 * with the same (dropout_p, dropout_seed) as the forward it differentiates.
 * hb_clam_dropout_masks: the masks themselves ([N,L1], [N,D], [N,D] fp32 HOST arrays holding 0 or 1/(1-p)), for parity
 * tests against the reference module and for the instance-clustering branch (inst_eval reads rows of the dropped h). */
int hb_clam_sb_forward_train(const float* feats, const int32_t* bag_offsets, int n_bags, int total_instances,
                             int max_bag_len, const void* const* weights_host, int n_models, int L0, int L1, int D, int C,
                             float* a_raw, float* m_out, float* logits, float* y_prob, int64_t* y_hat, void* workspace,
                             size_t workspace_bytes, float dropout_p, uint64_t dropout_seed, void* stream);
int hb_clam_sb_backward_train(const float* feats, int n_instances, const void* const* weights_host, const float* a_raw,
                              const float* m_pooled, const float* dlogits, const float* dm_ext, const float* da_ext,
                              const float* logits, const int64_t* label, float* loss_out, void* const* grads_host, int L0,
                              int L1, int D, int C, void* workspace, size_t workspace_bytes, float dropout_p,
                              uint64_t dropout_seed, void* stream);
int hb_clam_dropout_masks(int n_instances, int L1, int D, float dropout_p, uint64_t dropout_seed, float* m1_host,
                          float* ma_host, float* mb_host);

/* Multi-trial training step.  The reference's hyper-parameter search packs several Ray Tune
 * trials on one GPU as separate processes (main.py:40-52), each running train_loop (utils/core_utils.py:384-426) one bag at a
 * time: hundreds of configurations x 5 folds x 3 repeats of a 3.5 k-parameter model, every step pure launch latency.  Here up
 * to 8 independent trials of one head size advance ONE step each — trial t: its own 10 weight tensors, Adam state, bag,
 * label, dropout seed, learning rate, weight decay and step count — in six launches in total: paired forward (work table,
 * scores, combine), cross-entropy + backward (one prep CTA per trial, a (chunk, trial) grid of the recomputing backward),
 * one Adam launch over all 10 * n_trials tensors.  Same arithmetic as hb_clam_sb_forward_train + hb_clam_sb_backward_train
 * (logits + label mode) + hb_adam_step per trial.
 * feats: the trials' bags concatenated [total, 192]; bag_offsets (device) / bag_offsets_host: int32 [n_trials + 1];
 * weights_host / grads_host / exp_avg_host / exp_avg_sq_host: host arrays of 10 * n_trials device pointers, trial-major,
 * tensor order of hb_clam_sb_forward; labels: device int64 [n_trials]; lr / weight_decay / step / dropout_seeds: host arrays
 * [n_trials]; outputs a_raw [total], m_pooled [n_trials, L1], logits [n_trials, C], loss [n_trials] (device). */
size_t hb_clam_trials_workspace_bytes(int max_bag_len, int n_trials, int L1);
int hb_clam_sb_train_step_trials(const float* feats, const int32_t* bag_offsets, const int32_t* bag_offsets_host, int n_trials,
                                 const void* const* weights_host, void* const* grads_host, void* const* exp_avg_host,
                                 void* const* exp_avg_sq_host, const int64_t* labels, const float* lr, const float* weight_decay,
                                 const int* step, float beta1, float beta2, float eps, float dropout_p,
                                 const uint64_t* dropout_seeds, float* a_raw, float* m_pooled, float* logits, float* loss, int L0,
                                 int L1, int D, int C, void* workspace, size_t workspace_bytes, void* stream);

/* Multi-tensor Adam with L2 weight decay in one launch, torch.optim.Adam semantics (utils/utils.py:100-107 get_optim:
 * optim.Adam(..., lr, weight_decay=reg)): g += wd p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 * p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps).  Up to 16 fp32 tensors; step counts from 1. */
int hb_adam_step(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                 const int* numel, int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                 void* stream);

/* Region ingest: JPEG-compressed region tiles -> planar uint8 regions in device memory, the
 * [n, 3, height, width] layout hb_vit256_forward_u8 reads.  Replaces Whole_Slide_Bag_FP.__getitem__
 * (datasets/dataset_h5.py:194-207: OpenSlide read_region -> PIL -> ToTensor/Normalize on the CPU) + collate_features
 * (utils/utils.py:58-61) + the fp32 host -> device copy (HIPT_4K/hipt_4k.py:69) for tiles that are stored as JPEG.
 * The decoder is nvJPEG's batched decode (library code, loaded with dlopen at the first call; it allocates its own scratch
 * device memory — the one exception to "callers pass every buffer").  backend: -1 = GPU-assisted Huffman, falling back to
 * the library default; otherwise an nvjpegBackend_t value.
 * hb_jpeg_decode_tiles: a region is stored the way pyramidal TIFF / SVS files store it — as a grid of independently
 * compressed tile_h x tile_w JPEG tiles (tile = region is the degenerate case).  jpeg_host: n = n_regions * (height / tile_h)
 * * (width / tile_w) pointers to HOST bitstreams, region-major then row-major over the tile grid (keep them alive until the
 * stream has been synchronised); image i lands in its sub-rectangle of the three planes of its region (nvJPEG writes with
 * the region's row pitch).  Large batches of small tiles are what nvJPEG's GPU Huffman stage is built for: n <= max_batch. */
typedef struct hb_jpeg_decoder hb_jpeg_decoder;
int hb_jpeg_decoder_create(hb_jpeg_decoder** out, int max_batch, int backend);
const char* hb_jpeg_decoder_backend(const hb_jpeg_decoder* decoder);
void hb_jpeg_decoder_destroy(hb_jpeg_decoder* decoder);
int hb_jpeg_probe(hb_jpeg_decoder* decoder, const unsigned char* jpeg, size_t length, int* width, int* height,
                  int* components, int* subsampling);
int hb_jpeg_decode_tiles(hb_jpeg_decoder* decoder, const unsigned char* const* jpeg_host, const size_t* lengths, int n,
                         void* regions_u8, int height, int width, int tile_h, int tile_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIPT_B200_H */
