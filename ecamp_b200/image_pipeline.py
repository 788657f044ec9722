"""Image half of the reference loader with the pixel work on the GPU (ECAMP/Pre-training/module/pretrain_datasets.py:27-31,
47-52,113): `pil_loader` -> RandomResizedCrop(448, scale=(0.2, 1.0), BICUBIC) -> RandomHorizontalFlip -> Grayscale(3) ->
ToTensor -> Normalize(0.4721, 0.3037).

What stays on the host: JPEG decoding (to an 8-bit grayscale frame; MIMIC-CXR-JPG is single-channel) and the DRAW of the
random parameters - `draw_params` consumes torch's default generator exactly like torchvision's
`RandomResizedCrop.get_params` followed by `RandomHorizontalFlip` (up to ten (area, log-ratio) attempts, then `randint` for
the corner, then one `rand(1)` for the flip), so a seeded run yields the reference's crops and leaves the generator where the
reference leaves it.  What moves to the GPU: only the crop box of each frame is shipped (packed, pinned), resampled to
448 x 448 with Pillow's own arithmetic (csrc/image_pipeline.cu: bit-identical bytes), flipped, and handed to the step as the
uint8 batch `[B, 448, 448]` that `ECAMP.forward` normalises on the fly (or to `normalize()` for the fp32 `[B, 3, 448, 448]`
tensor of the reference collate).  There is no CPU fallback for the pixel work.
"""
import ctypes
import math

import torch

from . import _lib as L


class GpuImageTransform:
    def __init__(self, size=448, scale=(0.2, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), p_flip=0.5, mean=0.4721, std=0.3037,
                 device="cuda"):
        self.size, self.scale, self.ratio, self.p_flip = int(size), tuple(scale), tuple(ratio), float(p_flip)
        self.mean, self.std = float(mean), float(std)
        self.device = torch.device(device)
        self._stage = None      # pinned staging buffer of the packed crop boxes
        self._ws = None

    # ---- torchvision.transforms.RandomResizedCrop.get_params + RandomHorizontalFlip.forward, same generator draws ----
    def draw_params(self, height, width):
        area = height * width
        log_ratio = torch.log(torch.tensor(self.ratio))
        i = j = h = w = None
        for _ in range(10):
            target_area = area * torch.empty(1).uniform_(self.scale[0], self.scale[1]).item()
            aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
            w_ = int(round(math.sqrt(target_area * aspect_ratio)))
            h_ = int(round(math.sqrt(target_area / aspect_ratio)))
            if 0 < w_ <= width and 0 < h_ <= height:
                i = torch.randint(0, height - h_ + 1, size=(1,)).item()
                j = torch.randint(0, width - w_ + 1, size=(1,)).item()
                h, w = h_, w_
                break
        if h is None:   # fallback to a central crop
            in_ratio = float(width) / float(height)
            if in_ratio < min(self.ratio):
                w = width
                h = int(round(w / min(self.ratio)))
            elif in_ratio > max(self.ratio):
                h = height
                w = int(round(h * max(self.ratio)))
            else:
                w, h = width, height
            i, j = (height - h) // 2, (width - w) // 2
        flip = bool(torch.rand(1) < self.p_flip)
        return i, j, h, w, flip

    def __call__(self, frames, params=None):
        """frames: list of uint8 `[H, W]` tensors (decoded grayscale frames, host memory, sizes may differ); params: list of
        (i, j, h, w, flip) or None to draw them now, frame by frame, in order.  Returns uint8 `[B, size, size]` on the GPU."""
        lib = L.lib()
        B, S = len(frames), self.size
        if params is None:
            params = [self.draw_params(f.shape[0], f.shape[1]) for f in frames]
        total = sum(p[2] * p[3] for p in params)
        if self._stage is None or self._stage.numel() < total:
            self._stage = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        rows = []                      # descriptors are collected as plain ints: one tensor construction per batch
        off = toff = 0
        hmax = smax = 1
        for f, (i, j, h, w, flip) in zip(frames, params):
            if f.dtype != torch.uint8 or f.dim() != 2:
                raise ValueError("GpuImageTransform: frames must be uint8 [H, W] grayscale tensors")
            if not (0 <= i and 0 <= j and 0 < h and 0 < w and i + h <= f.shape[0] and j + w <= f.shape[1]):
                raise ValueError(f"GpuImageTransform: crop box {(i, j, h, w)} outside a {tuple(f.shape)} frame")
            self._stage[off:off + h * w].view(h, w).copy_(f[i:i + h, j:j + w])
            rows.append((off, toff, h | (w << 32), int(flip)))
            off += h * w
            toff += h * S
            hmax = max(hmax, h)
            smax = max(smax, h, w)
        kmax = max(1, lib.ecamp_image_resample_kmax(smax, S))   # the tap count grows with the source size: the largest crop decides
        desc = torch.tensor(rows, dtype=torch.int64).reshape(B, 4)
        crops = self._stage[:total].to(self.device, non_blocking=True)
        desc_d = desc.to(self.device, non_blocking=True)
        need = lib.ecamp_image_resized_crop_ws_bytes(B, S, kmax, ctypes.c_int64(toff))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, S, S, dtype=torch.uint8, device=self.device)
        L.check(lib.ecamp_image_resized_crop(L.ptr(crops), L.ptr(desc_d), B, hmax, S, kmax, L.ptr(self._ws),
                                             ctypes.c_int64(self._ws.numel()), ctypes.c_int64(toff), L.ptr(out), L.cur_stream()),
                "ecamp_image_resized_crop")
        return out

    def normalize(self, gray_u8):
        """Grayscale(3) + ToTensor + Normalize on the GPU: uint8 `[B, S, S]` -> fp32 `[B, 3, S, S]` (the reference collate's
        image tensor, bit-identical with the CPU transform)."""
        B, S = gray_u8.shape[0], gray_u8.shape[-1]
        out = torch.empty(B, 3, S, S, dtype=torch.float32, device=gray_u8.device)
        L.check(L.lib().ecamp_image_u8_normalize(L.ptr(gray_u8), ctypes.c_int64(B), ctypes.c_int64(S * S), ctypes.c_float(self.mean),
                                                 ctypes.c_float(self.std), L.ptr(out), L.cur_stream()), "ecamp_image_u8_normalize")
        return out
