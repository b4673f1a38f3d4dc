"""The input-pipeline tail in front of the patch embedding: mirror of timm.data.loader.PrefetchLoader
(timm/data/loader.py:80-181) that hands the model the loader's uint8 NCHW batch as it is -- H2D copy on a side stream,
no `.to(float).sub_(mean).div_(std)` pass (:139) -- because libb200timm's patch gather normalises while it reads
(b200_patchify_u8): 1 byte per pixel read instead of 1 read + 4 written + 4 (or 2) read again, and the result is
bit-identical to normalise-in-fp32-then-cast-to-bf16.

    loader = b200timm.data.PrefetchLoader(torch_loader, mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD)
    for images_u8, target in loader:          # uint8 [B, 3, H, W] on the GPU
        logits = model(images_u8)             # b200timm model, or a timm model with the plugin installed

Random erasing works on normalised floats and is not folded in: re_prob > 0 raises B200Error."""
from itertools import repeat
from typing import Optional, Tuple

import torch

from . import ops
from ._lib import B200Error

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)       # timm/data/constants.py:2-3
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _per_channel(x, n):
    """adapt_to_chs (loader.py:66-76): scalar -> n copies; wrong length -> the mean value n times."""
    if not isinstance(x, (tuple, list)):
        return tuple(repeat(x, n))
    if len(x) != n:
        return (sum(x) / len(x),) * n
    return tuple(x)


class PrefetchLoader:
    def __init__(self, loader, mean: Tuple[float, ...] = IMAGENET_DEFAULT_MEAN, std: Tuple[float, ...] = IMAGENET_DEFAULT_STD,
                 channels: int = 3, device: torch.device = torch.device("cuda"), img_dtype: Optional[torch.dtype] = None,
                 fp16: bool = False, re_prob: float = 0., re_mode: str = "const", re_count: int = 1, re_num_splits: int = 0):
        if re_prob > 0.:
            raise B200Error("PrefetchLoader: random erasing is not folded into the patch gather (use timm's loader)")
        self.loader, self.device = loader, torch.device(device)
        self.mean, self.std = _per_channel(mean, channels), _per_channel(std, channels)
        self.img_dtype = torch.uint8        # what the model receives; normalisation happens inside the patch gather
        ops.set_input_norm(self.mean, self.std)

    def __iter__(self):
        if self.device.type != "cuda":
            raise B200Error("PrefetchLoader: libb200timm needs a CUDA device (there is no CPU path)")
        stream = torch.cuda.Stream(device=self.device)
        first = True
        batch = None
        for next_input, next_target in self.loader:
            if next_input.dtype != torch.uint8:
                raise B200Error("PrefetchLoader: expected the uint8 batch of timm's fast_collate (loader.py:28-63)")
            with torch.cuda.stream(stream):
                next_input = next_input.to(device=self.device, non_blocking=True)
                next_target = next_target.to(device=self.device, non_blocking=True)
            if not first:
                yield batch
            first = False
            torch.cuda.current_stream(device=self.device).wait_stream(stream)
            next_input.record_stream(torch.cuda.current_stream(device=self.device))
            next_target.record_stream(torch.cuda.current_stream(device=self.device))
            batch = (next_input, next_target)
        if batch is not None:
            yield batch

    def __len__(self):
        return len(self.loader)

    @property
    def sampler(self):
        return self.loader.sampler

    @property
    def dataset(self):
        return self.loader.dataset

    @property
    def mixup_enabled(self):
        return getattr(self.loader.collate_fn, "mixup_enabled", False)

    @mixup_enabled.setter
    def mixup_enabled(self, x):
        if hasattr(self.loader.collate_fn, "mixup_enabled"):
            self.loader.collate_fn.mixup_enabled = x


class NaFlexPrefetchLoader:
    """Mirror of timm.data.naflex_loader.NaFlexPrefetchLoader (naflex_loader.py:27-193): moves the collated patch dict
    (`patches` uint8 [B, N, P*P*C], `patch_coord`, `patch_valid`) to the GPU on a side stream and hands it to the model AS IS
    -- the `(patches - mean * 255) / (std * 255)` pass (:161) and the fp32 `patches` tensor it materialises are folded into
    the kernel that builds the patch projection's GEMM operand (b200_patches_u8: normalise + bf16 cast + K padding in one
    pass).  Random erasing works on normalised floats and is not folded in: re_prob > 0 raises B200Error."""

    def __init__(self, loader, mean: Tuple[float, ...] = IMAGENET_DEFAULT_MEAN, std: Tuple[float, ...] = IMAGENET_DEFAULT_STD,
                 channels: int = 3, device: torch.device = torch.device("cuda"), img_dtype: Optional[torch.dtype] = None,
                 re_prob: float = 0., re_mode: str = "const", re_count: int = 1, re_num_splits: int = 0,
                 patchify_channels_last: bool = True):
        if re_prob > 0.:
            raise B200Error("NaFlexPrefetchLoader: random erasing is not folded into the patch gather (use timm's loader)")
        if not patchify_channels_last:
            raise B200Error("NaFlexPrefetchLoader: only the P-P-C patch layout (patchify_channels_last=True) is covered")
        self.loader, self.device, self.channels = loader, torch.device(device), channels
        self.mean, self.std = _per_channel(mean, channels), _per_channel(std, channels)
        self.img_dtype = torch.uint8
        ops.set_input_norm(self.mean, self.std)

    def __iter__(self):
        if self.device.type != "cuda":
            raise B200Error("NaFlexPrefetchLoader: libb200timm needs a CUDA device (there is no CPU path)")
        stream = torch.cuda.Stream(device=self.device)
        batch = None
        for next_dict, next_target in self.loader:
            if next_dict["patches"].dtype != torch.uint8:
                raise B200Error("NaFlexPrefetchLoader: expected uint8 patches from the NaFlex collator")
            with torch.cuda.stream(stream):
                next_dict = {k: (v.to(device=self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v)
                             for k, v in next_dict.items()}
                next_target = next_target.to(device=self.device, non_blocking=True)
            if batch is not None:
                yield batch
            cur = torch.cuda.current_stream(device=self.device)
            cur.wait_stream(stream)
            for v in list(next_dict.values()) + [next_target]:
                if isinstance(v, torch.Tensor):
                    v.record_stream(cur)
            batch = (next_dict, next_target)
        if batch is not None:
            yield batch

    def __len__(self):
        return len(self.loader)

    @property
    def sampler(self):
        return self.loader.sampler

    @property
    def dataset(self):
        return self.loader.dataset


set_input_norm = ops.set_input_norm
__all__ = ["PrefetchLoader", "NaFlexPrefetchLoader", "set_input_norm", "IMAGENET_DEFAULT_MEAN", "IMAGENET_DEFAULT_STD"]
