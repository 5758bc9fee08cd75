"""ctypes binding of libhipt_b200.so (the C ABI declared in include/hipt_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.  Tensors cross the boundary as raw
device pointers (`tensor.data_ptr()`) plus sizes; the CUDA stream is torch's current stream.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HB_LIB_PATH") or os.path.join(_PKG, "lib", "libhipt_b200.so")   # override: kernel experiments

HB_EPI_BIAS_BF16 = 0
HB_EPI_BIAS_GELU_BF16 = 1
HB_EPI_BIAS_RESADD_F32 = 2
HB_EPI_TOKENS_F32 = 3
HB_EPI_TOKENS_GELU_F32 = 4
HB_EPI_BIAS_GELU_FAST_BF16 = 5
HB_EPI_LNFOLD_BF16 = 6
HB_EPI_LNFOLD_GELU_BF16 = 7
HB_EPI_RESID_STATS_F32 = 8
HB_EPI_LNFOLD_GELU2_BF16 = 9
HB_EPI_RESID_BF16 = 10


class HbVitConfig(C.Structure):
    _fields_ = [("dim", C.c_int), ("heads", C.c_int), ("depth", C.c_int), ("mlp_dim", C.c_int),
                ("max_rows", C.c_int), ("ln_eps", C.c_float)]


# name -> (restype, argtypes); every symbol of include/hipt_b200.h
SIGNATURES = {
    "hb_abi_version": (C.c_int, []),
    "hb_last_error": (C.c_char_p, []),
    "hb_device_check": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "hb_launch_count": (C.c_longlong, []),
    "hb_prof_enable": (C.c_int, [C.c_int]),
    "hb_prof_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    "hb_gemm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_int, C.c_void_p]),
    "hb_gemm_lnfold_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                      C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hb_gemm_resid_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hb_gemm_resid_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hb_mlp_fused_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hb_layernorm": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_int, C.c_void_p]),
    "hb_layernorm_bf16": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_void_p]),
    "hb_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "hb_im2col_patches": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "hb_vit_workspace_bytes": (C.c_size_t, [C.POINTER(HbVitConfig)]),
    "hb_vit_plan_create": (C.c_int, [C.POINTER(HbVitConfig), C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_void_p)]),
    "hb_vit_plan_destroy": (None, [C.c_void_p]),
    "hb_vit_plan_set_depth_limit": (C.c_int, [C.c_void_p, C.c_int]),
    "hb_vit_plan_set_cls_attention": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_vit_plan_buffer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "hb_vit256_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                    C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "hb_vit256_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "hb_vit4k_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_clam_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "hb_clam_sb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hb_clam_sb_backward": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "hb_clam_sb_backward_ce": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "hb_clam_sb_forward_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_uint64, C.c_void_p]),
    "hb_clam_sb_backward_train": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_float, C.c_uint64,
                                            C.c_void_p]),
    "hb_clam_dropout_masks": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_clam_trials_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hb_clam_sb_train_step_trials": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                               C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_float,
                                               C.c_float, C.c_float, C.c_float, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                               C.c_void_p]),
    "hb_jpeg_decoder_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "hb_jpeg_decoder_backend": (C.c_char_p, [C.c_void_p]),
    "hb_jpeg_decoder_destroy": (None, [C.c_void_p]),
    "hb_jpeg_probe": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hb_jpeg_decode_tiles": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hb_adam_step": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                               C.POINTER(C.c_int), C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                               C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m hipt_abmil_atec23_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libhipt_b200: " + load().hb_last_error().decode("utf-8", "replace"))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: this path has no CPU implementation")


_device_checked = set()


def device_check():
    dev = torch.cuda.current_device()
    if dev not in _device_checked:
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        check(load().hb_device_check(C.byref(sm), C.byref(ma), C.byref(mi)))
        _device_checked.add(dev)


PROF_KINDS = 32
PROF_NAMES = {0: "im2col", 1: "embed_gemm", 2: "cls_rows", 3: "layernorm", 4: "qkv_gemm", 5: "attention",
              6: "proj_gemm", 7: "fc1_gemm", 8: "fc2_gemm", 9: "final_ln", 10: "clam_scores", 11: "clam_combine",
              12: "mlp_fused", 15: "clam_work_table"}


def prof_enable(on=True):
    check(load().hb_prof_enable(int(on)))


def prof_read():
    """{kernel name: (total ms, launches)} since the last read; ViT-4K kinds are prefixed 'vit4k_'."""
    ms = (C.c_double * PROF_KINDS)()
    cnt = (C.c_longlong * PROF_KINDS)()
    check(load().hb_prof_read(ms, cnt, PROF_KINDS))
    out = {}
    for k in range(PROF_KINDS):
        if cnt[k]:
            base = PROF_NAMES.get(k % 16, f"kind{k % 16}")
            out[("vit4k_" if k >= 16 else "") + base] = (ms[k], cnt[k])
    return out


def launch_count():
    return int(load().hb_launch_count())


# ------------------------------------------------------------------------------------------------ thin op wrappers
def gemm_bf16(a, w, bias, epilogue, out=None, tok_table=None, tokens_per_seq=0):
    """out = epilogue(a @ w.T + bias); a [M,K] bf16, w [N,K] bf16, bias [N] fp32."""
    require_cuda(a, "a")
    device_check()
    M, K = a.shape
    N = w.shape[0]
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert a.is_contiguous() and w.is_contiguous() and w.shape[1] == K
    if out is None:
        if epilogue in (HB_EPI_BIAS_BF16, HB_EPI_BIAS_GELU_BF16, HB_EPI_BIAS_GELU_FAST_BF16):
            out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        else:
            raise ValueError("an output tensor is required for this epilogue")
    check(load().hb_gemm_bf16(ptr(a), ptr(w), ptr(bias), epilogue, ptr(out), M, N, K, ptr(tok_table), tokens_per_seq,
                              stream_ptr()))
    return out


def gemm_lnfold_bf16(xb, w_gamma, c, d, row_stats, eps, gelu=0):
    """LayerNorm + Linear (+GELU) as one GEMM on the un-normalised bf16 rows (see hb_gemm_lnfold_bf16); gelu: 0 / 1 / 2
    (2 x GELU).  row_stats: [M, 2] or [n_part, rows >= M, 2] partial (sum, sum of squares) planes."""
    require_cuda(xb, "xb")
    device_check()
    M, K = xb.shape
    N = w_gamma.shape[0]
    if row_stats.dim() == 2:
        row_stats = row_stats.unsqueeze(0)
    assert row_stats.is_contiguous() and row_stats.shape[1] >= M and row_stats.shape[2] == 2
    out = torch.empty((M, N), dtype=torch.bfloat16, device=xb.device)
    check(load().hb_gemm_lnfold_bf16(ptr(xb), ptr(w_gamma), ptr(c), ptr(d), ptr(row_stats), row_stats.shape[0],
                                     row_stats.shape[1], eps, int(gelu), ptr(out), M, N, K, stream_ptr()))
    return out


def gemm_resid_bf16(a, w, bias, res, out=None, res_pitch_bytes=0):
    """out = bf16(res + a @ w.T + bias) (fp32 add, one rounding); res may be out itself (in place, the default) or a
    strided source.  Returns (out, stats_part [N/64, M, 2])."""
    require_cuda(a, "a")
    device_check()
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = res
    stats = torch.empty((N // 64, M, 2), dtype=torch.float32, device=a.device)
    check(load().hb_gemm_resid_bf16(ptr(a), ptr(w), ptr(bias), ptr(res), res_pitch_bytes, ptr(out), ptr(stats), M, M, N, K,
                                    stream_ptr()))
    return out, stats


def mlp_fused_bf16(xb, w1_gamma, c1, d1, w2_half, b2, stats_in):
    """xb (bf16 [M, 384], updated in place) = bf16(xb + (2 gelu(LN-folded fc1)) @ w2_half.T + b2); returns the
    [6, M, 2] partial row statistics of the new rows.  stats_in: [6, rows >= M, 2]."""
    require_cuda(xb, "xb")
    device_check()
    M, D = xb.shape
    H = w1_gamma.shape[0]
    assert stats_in.is_contiguous() and stats_in.shape[0] == 6 and stats_in.shape[1] >= M
    stats_out = torch.empty((6, stats_in.shape[1], 2), dtype=torch.float32, device=xb.device)
    check(load().hb_mlp_fused_bf16(ptr(xb), ptr(w1_gamma), ptr(c1), ptr(d1), ptr(w2_half), ptr(b2), ptr(stats_in),
                                   ptr(stats_out), stats_in.shape[1], 1e-6, M, D, H, stream_ptr()))
    return stats_out


def gemm_resid_stats(a, w, bias, x, xb, stats_out, stats_clear=None):
    """x += a @ w.T + bias in place; xb = bf16(x); stats_out += (row sum, row sum of squares); stats_clear zeroed."""
    require_cuda(a, "a")
    device_check()
    M, K = a.shape
    N = w.shape[0]
    check(load().hb_gemm_resid_stats(ptr(a), ptr(w), ptr(bias), ptr(x), ptr(xb), ptr(stats_out), ptr(stats_clear),
                                     M, N, K, stream_ptr()))
    return x


def layernorm(x, gamma, beta, eps, rows, dim, row_stride=None, want_bf16=True, want_f32=False):
    """nn.LayerNorm over fp32 or bf16 rows (row_stride in elements)."""
    require_cuda(x, "x")
    device_check()
    assert x.dtype in (torch.float32, torch.bfloat16)
    ob = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    of = torch.empty((rows, dim), dtype=torch.float32, device=x.device) if want_f32 else None
    fn = load().hb_layernorm if x.dtype == torch.float32 else load().hb_layernorm_bf16
    check(fn(ptr(x), row_stride if row_stride is not None else dim, ptr(gamma), ptr(beta), eps,
             ptr(ob), ptr(of), rows, dim, stream_ptr()))
    return ob, of


def attention(qkv, n_seq, seq_len, heads, head_dim, scale):
    require_cuda(qkv, "qkv")
    device_check()
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    out = torch.empty((n_seq * seq_len, heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    check(load().hb_attention(ptr(qkv), ptr(out), n_seq, seq_len, heads, head_dim, scale, stream_ptr()))
    return out


def image_layout(image):
    """(patch_stride, chan_stride, row_pitch, grid_cols, n_patches, patches_per_image, image_stride_bytes) of a region
    [3,H,W], a batch of regions [R,3,H,W] (H, W multiples of 256) or a patch batch [B,3,256,256]."""
    assert image.dtype in (torch.uint8, torch.float32) and image.stride(-1) == 1
    if image.dim() == 3:
        assert image.shape[0] == 3 and image.shape[1] % 256 == 0 and image.shape[2] % 256 == 0
        gc = image.shape[2] // 256
        ppi = (image.shape[1] // 256) * gc
        return 0, image.stride(0), image.stride(1), gc, ppi, ppi, 0
    assert image.dim() == 4 and image.shape[1] == 3
    if tuple(image.shape[2:]) == (256, 256):
        return image.stride(0), image.stride(1), image.stride(2), 0, image.shape[0], 1, 0
    assert image.shape[2] % 256 == 0 and image.shape[3] % 256 == 0
    gc = image.shape[3] // 256
    ppi = (image.shape[2] // 256) * gc
    return 0, image.stride(1), image.stride(2), gc, ppi * image.shape[0], ppi, image.stride(0) * image.element_size()


def im2col_patches(image, patch_begin, n_patches):
    """image: region [3, H, W] or patch batch [B, 3, 256, 256]; uint8 or fp32; unit stride along the last dim."""
    require_cuda(image, "image")
    device_check()
    ps, cs, rp, gc, _, _, _ = image_layout(image)
    assert image.dim() == 3 or gc == 0, "im2col_patches takes one region or a patch batch"
    a = torch.empty((n_patches * 256, 768), dtype=torch.bfloat16, device=image.device)
    check(load().hb_im2col_patches(ptr(image), int(image.dtype == torch.float32), ps, cs, rp, gc, patch_begin,
                                   n_patches, ptr(a), stream_ptr()))
    return a
