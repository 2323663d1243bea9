"""Synthetic SVBRDF maps for tests (same distributions as bench.py / SURVEY.md §8d)."""
import torch


def synthetic_maps(batch, size, seed, stress=False, device="cpu"):
    """Unit upper-hemisphere normals (xy ~ N(0,0.3^2), z=1, normalised), diffuse/specular U(0,1),
    roughness U(0.1,1) replicated x3.  ``stress``: independent roughness channels with exact zeros
    and sub-clamp values, non-unit normals, 5 % of them facing away."""
    g = torch.Generator("cpu").manual_seed(seed)
    xy = torch.randn(batch, 2, size, size, generator=g) * 0.3
    n = torch.cat((xy, torch.ones(batch, 1, size, size)), dim=1)
    n = n / n.norm(dim=1, keepdim=True)
    d = torch.rand(batch, 3, size, size, generator=g)
    s = torch.rand(batch, 3, size, size, generator=g)
    if not stress:
        r = (torch.rand(batch, 1, size, size, generator=g) * 0.9 + 0.1).repeat(1, 3, 1, 1)
    else:
        r = torch.rand(batch, 3, size, size, generator=g)
        kill = torch.rand(batch, 3, size, size, generator=g)
        r = torch.where(kill < 0.01, torch.zeros_like(r), r)
        r = torch.where((kill >= 0.01) & (kill < 0.02), r * 1e-3, r)
        scale = 0.5 + torch.rand(batch, 1, size, size, generator=g)
        flip = torch.rand(batch, 1, size, size, generator=g) < 0.05
        n = n * scale
        n[:, 2:3] = torch.where(flip, -n[:, 2:3], n[:, 2:3])
    return torch.cat((n, d, r, s), dim=1).contiguous().to(device)
