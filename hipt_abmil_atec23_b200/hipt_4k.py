"""HIPT_4K region encoder on the B200 CUDA path (interface of HIPT_4K/hipt_4k.py:31-76, :308-330 in the reference).

`HIPT_4K(model256_path, model4k_path, device256, device4k).forward(x)` keeps the reference contract — x is a
[1, 3, W, H] fp32 tensor already normalised by `eval_transforms()`, the result is the [1, 192] ViT-4K CLS token on
`device4k` — but nothing is unfolded, copied to the host or bounced between devices (hipt_4k.py:64-74): the 256 x 256
patch gather is address arithmetic inside the patch-embed operand kernel, ViT-256 CLS tokens stay on the GPU as the
bf16 operand of the ViT-4K `phi` GEMM, and the (w,h)-grid reshuffle at :73 is the identity on token order.

Beyond the reference API, `forward_regions_u8` takes raw uint8 regions (what OpenSlide decodes) with the normalisation
folded into the patch-embed weights, and batches ViT-4K over many regions; this is the throughput path bench.py times.
"""
import torch
import torch.nn as nn

from .hipt_model_utils import HIPT_MEAN, HIPT_STD, eval_transforms, get_vit256, get_vit4k, roll_batch2img, tensorbatch2im


def center_crop_offsets(size, crop):
    """torchvision CenterCrop offset: int(round((size - crop) / 2))."""
    return int(round((size - crop) / 2.0))


class HIPT_4K(nn.Module):
    """HIPT model (ViT-256 over [256 x 256] patches, ViT-4K over their CLS-token grid)."""

    def __init__(self, model256_path: str = '../Checkpoints/vit256_small_dino.pth',
                 model4k_path: str = '../Checkpoints/vit4k_xs_dino.pth',
                 device256=torch.device('cuda:0'), device4k=torch.device('cuda:1')):
        super().__init__()
        self.model256 = get_vit256(pretrained_weights=model256_path).to(device256)
        self.model4k = get_vit4k(pretrained_weights=model4k_path).to(device4k)
        self.device256 = device256
        self.device4k = device4k

    @classmethod
    def from_modules(cls, model256, model4k, device256=torch.device('cuda:0'), device4k=None):
        """Build from already-constructed ViTs (random-init parity / benchmark runs without checkpoint files)."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        device4k = device256 if device4k is None else device4k
        for m in (model256, model4k):
            for p in m.parameters():
                p.requires_grad = False
            m.eval()
        self.model256 = model256.to(device256)
        self.model4k = model4k.to(device4k)
        self.device256 = device256
        self.device4k = device4k
        return self

    # ---------------------------------------------------------------------------------------- reference API
    def prepare_img_tensor(self, img: torch.Tensor, patch_size=256):
        """Centre-crop so W and H are multiples of patch_size; returns (img, w_256, h_256) (hipt_4k.py:308-330)."""
        b, c, w, h = img.shape
        cw, ch = w - w % patch_size, h - h % patch_size
        top, left = center_crop_offsets(w, cw), center_crop_offsets(h, ch)
        return img[:, :, top:top + cw, left:left + ch], w // patch_size, h // patch_size

    def _cls256(self, x, mean=None, std=None, want_f32=True):
        """ViT-256 CLS tokens of every 256x256 patch of one region view [3, W, H] on device256."""
        eng = self.model256._engine(x.device)
        if x.stride(-1) != 1:                     # e.g. a channels-last view: the kernels read rows of unit stride
            x = x.contiguous()
        return eng.forward_patches(x, mean=mean, std=std, want_f32=want_f32)

    @torch.no_grad()
    def forward(self, x):
        """x: [1, 3, W, H] normalised fp32 -> [1, 192] fp32 on device4k (hipt_4k.py:48-76)."""
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 3:
            # the reference fails in reshape(w_256, h_256, 384) at :73 for any batch other than 1
            raise RuntimeError(f"HIPT_4K.forward takes one region [1,3,W,H], got {tuple(x.shape)}")
        img, w_256, h_256 = self.prepare_img_tensor(x)
        img = img.to(self.device256, non_blocking=True)[0]
        if img.dtype != torch.float32:
            img = img.float()
        _, cls_bf16 = self._cls256(img, want_f32=False)
        if torch.device(self.device4k) != cls_bf16.device:
            cls_bf16 = cls_bf16.to(self.device4k, non_blocking=True)
        return self.model4k._engine(cls_bf16.device).forward_grid(cls_bf16, 1, w_256, h_256)

    @torch.no_grad()
    def forward_asset_dict(self, x: torch.Tensor):
        """Intermediate representations as numpy arrays (hipt_4k.py:79-118)."""
        img, w_256, h_256 = self.prepare_img_tensor(x)
        img = img.to(self.device256, non_blocking=True)[0].float()
        cls_f32, cls_bf16 = self._cls256(img)
        mean256 = cls_f32.mean(dim=0, keepdim=True)
        cls4k = self.model4k._engine(torch.device(self.device4k)).forward_grid(
            cls_bf16.to(self.device4k), 1, w_256, h_256)
        return {
            'features_cls256': cls_f32.cpu().numpy(),
            'features_mean256': mean256.cpu().numpy(),
            'features_cls4k': cls4k.cpu().numpy(),
            'features_mean256_cls4k': torch.cat([mean256.to(cls4k.device), cls4k], dim=1).cpu().numpy(),
        }

    # ------------------------------------------------------------------------------------ throughput API
    @torch.no_grad()
    def forward_regions_u8(self, regions_u8, mean=HIPT_MEAN, std=HIPT_STD, return_cls256=False):
        """regions_u8: [R, 3, W, H] uint8 on device256 (W, H multiples of 256) -> [R, 192] fp32 on device4k.

        Same arithmetic as `forward(eval_transforms-normalised region)` for each region, with ToTensor + Normalize
        folded into the patch-embed weights; ViT-4K runs once over all R grids."""
        if regions_u8.dtype != torch.uint8 or regions_u8.dim() != 4 or regions_u8.shape[1] != 3:
            raise RuntimeError("forward_regions_u8 takes [R,3,W,H] uint8")
        R, _, W, H = regions_u8.shape
        if W % 256 or H % 256:
            raise RuntimeError("uint8 regions must already be cropped to multiples of 256")
        w_256, h_256 = W // 256, H // 256
        T = w_256 * h_256
        dev = regions_u8.device
        eng = self.model256._engine(dev)
        cls_bf16 = torch.empty((R * T, eng.dim), dtype=torch.bfloat16, device=dev)
        # the whole batch in one call: the engine walks it in launches of its capacity (two 4096x4096 regions each)
        eng.forward_patches(regions_u8, mean=mean, std=std, want_f32=False, out_bf16=cls_bf16)
        if torch.device(self.device4k) != cls_bf16.device:
            cls_bf16 = cls_bf16.to(self.device4k, non_blocking=True)
        out = self.model4k._engine(cls_bf16.device).forward_grid(cls_bf16, R, w_256, h_256)
        return (out, cls_bf16) if return_cls256 else out

    # ------------------------------------------------------------------------------ hierarchical attention maps
    @torch.no_grad()
    def region_cls_attention(self, x):
        """CLS-query attention of the LAST block of both ViTs for one normalised region x [1,3,W,H]:
        (attention_256 [w*h, heads, 256], attention_4k [heads, w*h], w_256, h_256) — exactly the slices
        `get_last_selfattention(...)[:, :, 0, 1:]` that the reference's heatmap code reads (hipt_4k.py:145-147, 155-157),
        emitted by the fused CLS-only attention launch of the forward pass itself: one model pass, no [B,6,257,257] matrix."""
        img, w_256, h_256 = self.prepare_img_tensor(x)
        img = img.to(self.device256, non_blocking=True)[0].float().contiguous()
        eng256 = self.model256._engine(img.device)
        T = w_256 * h_256
        a256 = torch.empty((T, eng256.heads, 257), dtype=torch.float32, device=img.device)
        _, cls_bf16 = eng256.forward_patches(img, want_f32=False, cls_attn=a256)
        if torch.device(self.device4k) != cls_bf16.device:
            cls_bf16 = cls_bf16.to(self.device4k, non_blocking=True)
        eng4k = self.model4k._engine(cls_bf16.device)
        a4k = torch.empty((1, eng4k.heads, T + 1), dtype=torch.float32, device=cls_bf16.device)
        eng4k.forward_grid(cls_bf16, 1, w_256, h_256, cls_attn=a4k)
        return a256[:, :, 1:], a4k[0, :, 1:], w_256, h_256

    @torch.no_grad()
    def _get_region_attention_scores(self, region, scale=1):
        """Interface of hipt_4k.py:121-164: region = PIL image (or HxWx3 uint8 array); returns
        (patches [n, 256/scale, 256/scale, 3] uint8, attention_256 [n, heads, 256/scale, 256/scale],
         attention_4k [heads, W/scale, H/scale]) as numpy arrays."""
        x = eval_transforms()(region).unsqueeze(dim=0)
        a256, a4k, w_256, h_256 = self.region_cls_attention(x)
        nh = a256.shape[1]
        attention_256 = a256.reshape(w_256 * h_256, nh, 16, 16)
        attention_256 = nn.functional.interpolate(attention_256, scale_factor=int(16 / scale), mode="nearest").cpu().numpy()
        attention_4k = a4k.reshape(a4k.shape[0], w_256, h_256)
        attention_4k = nn.functional.interpolate(attention_4k.unsqueeze(0), scale_factor=int(256 / scale), mode="nearest")[0].cpu().numpy()
        batch_256, _, _ = self.prepare_img_tensor(x)
        batch_256 = batch_256.unfold(2, 256, 256).unfold(3, 256, 256)
        batch_256 = batch_256.permute(0, 2, 3, 1, 4, 5).reshape(-1, 3, 256, 256)          # 'b c p1 p2 w h -> (b p1 p2) c w h'
        if scale != 1:
            batch_256 = nn.functional.interpolate(batch_256, scale_factor=(1 / scale), mode="nearest")
        return tensorbatch2im(batch_256), attention_256, attention_4k

    def get_region_attention_heatmaps(self, *a, **k):
        raise NotImplementedError("drawing the blended heatmap images (matplotlib / PIL, hipt_4k.py:167-306) is outside the "
                                  "accelerated hot path; _get_region_attention_scores provides their inputs")
