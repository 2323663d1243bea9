"""Synthetic network inputs: the renderer used on the data side.

The reference's dataset renders each sample's input photographs with ``LocalRenderer`` on the CPU, one
``render`` call per view, inside ``__getitem__`` (development/multiImage_pytorch/dataset.py:162-221).  Here
the scenes are sampled exactly as the reference does (same draws from the global CPU generator, same order),
but all ``count`` views are shaded by ONE launch of the render kernel; sensor noise and the clamp to [0,1]
follow.  ``render_inputs`` has the reference method's signature so it can stand in for
``SvbrdfDataset.render_inputs``.
"""
import math

import numpy as np
import torch

from .renderers import render_records
from .utils import generate_normalized_random_direction

MIN_EPS, MAX_EPS = 0.001, 0.02             # dataset.py:164-165
FIXED_LIGHT_DISTANCE = 2.197               # dataset.py:166
FIXED_VIEW_DISTANCE = 2.75                 # dataset.py:167
NOISE_LOG_STD = (math.log(0.005), 0.3)     # dataset.py:214 (np.log(0.005))


def _u(n, lo, hi):
    return torch.empty(n, dtype=torch.float32).uniform_(lo, hi)


def sample_input_scenes(count, use_augmentation=False):
    """``count`` scene records [count,9] for the input photographs, drawn like dataset.py:169-203: the first
    view and light sit (almost) straight above the sample, the others on the hemisphere; with augmentation the
    light intensity, white balance and view distance vary."""
    light = torch.cat([_u(2, -0.75, 0.75), torch.ones(1) * FIXED_LIGHT_DISTANCE], dim=-1).unsqueeze(0)
    if count > 1:
        hemi = generate_normalized_random_direction(count - 1, min_eps=MIN_EPS, max_eps=MAX_EPS) * FIXED_LIGHT_DISTANCE
        light = torch.cat([light, hemi], dim=0)
    colors = torch.tensor([30.0]).unsqueeze(-1)
    if use_augmentation:
        std = torch.exp(torch.empty(1).normal_(mean=-2.0, std=0.5)).numpy()[0]
        colors = torch.abs(torch.empty(count).normal_(mean=20.0, std=float(std))).unsqueeze(-1)
    colors = colors.expand(count, 3)
    if use_augmentation:
        colors = colors * torch.abs(torch.empty(count, 3).normal_(mean=1.0, std=0.03))      # white balance
        view_distance = _u(count, 0.25, 2.75)
    else:
        view_distance = torch.ones(count) * FIXED_VIEW_DISTANCE
    view = torch.cat([_u(2, -0.25, 0.25), view_distance[:1]], dim=-1).unsqueeze(0)
    if count > 1:
        hemi = generate_normalized_random_direction(count - 1, min_eps=MIN_EPS, max_eps=MAX_EPS) * view_distance[1:].unsqueeze(-1)
        view = torch.cat([view, hemi], dim=0)
    return torch.cat((view, light, colors), dim=-1).contiguous()


def render_inputs(svbrdf, count, use_augmentation=False, noise="reference"):
    """[12,H,W] maps -> [count,3,H,W] input images in [0,1] (dataset.py:162-221).

    ``noise``: "reference" draws the Gaussian sensor noise from the global CPU generator exactly like the
    reference (per view: one log-normal std, then 3*H*W normals) so that a seed reproduces the reference's
    images up to the renderer's fp32 tolerance; "device" draws it on the maps' device (fast, different stream);
    ``None`` disables it.  The result lives on ``svbrdf``'s device."""
    if svbrdf.dim() != 3:
        raise ValueError("render_inputs expects one sample [12,H,W], got %s" % (tuple(svbrdf.shape),))
    records = sample_input_scenes(count, use_augmentation)
    images = render_records(svbrdf, records)                         # [count,3,H,W], one kernel launch
    if noise is None:
        return torch.clamp(images, min=0.0, max=1.0)
    h, w = svbrdf.shape[-2:]
    if noise == "reference":
        parts = []
        for _ in range(count):
            std = torch.exp(torch.empty(1).normal_(mean=NOISE_LOG_STD[0], std=NOISE_LOG_STD[1])).numpy()[0]
            parts.append(torch.zeros(1, 3, h, w).normal_(mean=0.0, std=float(std)))
        nz = torch.cat(parts, dim=0).to(images.device, non_blocking=True)
    elif noise == "device":
        std = torch.exp(torch.empty(count).normal_(mean=NOISE_LOG_STD[0], std=NOISE_LOG_STD[1])).to(images.device)
        nz = torch.randn(count, 3, h, w, device=images.device) * std.view(count, 1, 1, 1)
    else:
        raise ValueError("noise must be 'reference', 'device' or None")
    return torch.clamp(images + nz, min=0.0, max=1.0)


__all__ = ["sample_input_scenes", "render_inputs"]
