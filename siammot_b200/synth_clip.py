"""Deterministic synthetic video frames (there are no datasets offline).

Frame t is a fixed smooth background plus ``n_obj`` textured rectangles that drift a few pixels
per frame, already in network-input form (3,H,W) float32, roughly zero-mean / unit-range like a
Normalize()'d image (build_augmentation.py:48-50)."""
import math

import torch


def make_clip(n_frames, height, width, n_obj=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    yy = torch.arange(height, dtype=torch.float32)[:, None] / height
    xx = torch.arange(width, dtype=torch.float32)[None, :] / width
    bg = torch.stack([torch.sin(6.0 * xx + 2.0 * c) * torch.cos(5.0 * yy - c) for c in range(3)]) * 0.5
    bg = bg + 0.3 * torch.randn(3, height, width, generator=g)
    cx = torch.rand(n_obj, generator=g) * (width * 0.8) + width * 0.1
    cy = torch.rand(n_obj, generator=g) * (height * 0.6) + height * 0.2
    w = torch.rand(n_obj, generator=g) * (width * 0.12) + width * 0.05
    h = torch.rand(n_obj, generator=g) * (height * 0.35) + height * 0.15
    vx = (torch.rand(n_obj, generator=g) - 0.5) * 8.0
    vy = (torch.rand(n_obj, generator=g) - 0.5) * 4.0
    tex = torch.randn(n_obj, 3, 64, 64, generator=g)
    frames = []
    for t in range(n_frames):
        f = bg.clone()
        for k in range(n_obj):
            x0 = int(max(0, min(width - 2, cx[k] + vx[k] * t - w[k] / 2)))
            y0 = int(max(0, min(height - 2, cy[k] + vy[k] * t - h[k] / 2)))
            x1 = int(max(x0 + 1, min(width, x0 + w[k])))
            y1 = int(max(y0 + 1, min(height, y0 + h[k])))
            patch = torch.nn.functional.interpolate(tex[k][None], size=(y1 - y0, x1 - x0), mode="nearest")[0]
            f[:, y0:y1, x0:x1] = 1.5 * math.copysign(1.0, float(tex[k, 0, 0, 0])) + 0.5 * patch
        frames.append(f)
    return torch.stack(frames)


def make_clip_u8(n_frames, height, width, n_obj=6, seed=0, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """The same clip as decoded video: RGB uint8 (T, H, W, 3) frames whose ToTensor + Normalize(mean, std) image is
    make_clip's frame up to 8-bit quantisation (input of the raw-frame path, siammot_b200/preprocess.py)."""
    f = make_clip(n_frames, height, width, n_obj, seed)
    m = torch.tensor(mean, dtype=torch.float32)[None, :, None, None]
    s = torch.tensor(std, dtype=torch.float32)[None, :, None, None]
    return ((f * s + m).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
