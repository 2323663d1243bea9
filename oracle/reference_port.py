"""CPU restatement of the reference's rendering-loss path.  TEST INFRASTRUCTURE ONLY.

This module is the *oracle*: it restates, with eager PyTorch CPU ops and autograd,
what ``mworchel/svbrdf-estimation`` computes on the path

    RenderingLoss.forward  ->  LocalRenderer.render  ->  Cook-Torrance / GGX shading

(reference files, relative to ``development/multiImage_pytorch/``:
``renderers.py:8-104``, ``losses.py:7-63``, ``environment.py:18-55``,
``utils.py:36-58,100-111``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``svbrdf_estimation_b200`` never does and has no CPU fallback.

Parity status: **pinned**.  The reference itself ships no golden vectors for this path
(SURVEY.md §4), so the pins were produced by importing the unmodified reference in the
build container (``tests/golden/make_golden.py``) and are committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them
(bit-exact in fp32 for renders/loss, fp64 to 1e-12).

The arithmetic deliberately keeps the reference's evaluation order (for example
``1/sqrt(d2)**2`` for the falloff and the ill-conditioned GGX denominator) so that the
fp32 result is the reference's fp32 result and the fp64 result is the ground truth the
CUDA kernels are compared with.  It works for any floating dtype: pass fp64 maps for the
ground truth.
"""
import math

import torch

# ----------------------------------------------------------------------------------------
# small vector helpers (renderers.py:8-12)
# ----------------------------------------------------------------------------------------

def channel_dot(a, b):
    """Sum of products over the channel axis (dim -3), kept as a size-1 axis.
    Reference: ``dot_product`` renderers.py:8-9."""
    return torch.sum(a * b, dim=-3, keepdim=True)


def unit(a):
    """``a / |a|`` without an epsilon.  Reference: ``normalize`` renderers.py:11-12."""
    return a / torch.sqrt(channel_dot(a, a))


def positive_mask(x):
    """1 where x > 0 else 0, no gradient.  Reference: ``LocalRenderer.xi`` renderers.py:15-16."""
    return (x > 0.0) * torch.ones_like(x)


# ----------------------------------------------------------------------------------------
# SVBRDF channel contract (utils.py:36-58)
# ----------------------------------------------------------------------------------------

def split_maps(maps):
    """[...,12,H,W] -> normals, diffuse, roughness, specular (3 channels each on dim -3).
    Reference: ``unpack_svbrdf`` utils.py:40-58 (un-encoded branch)."""
    if maps.shape[-3] != 12:
        raise ValueError("expected 12 channels on dim -3, got %d" % maps.shape[-3])
    n, d, r, s = torch.split(maps, 3, dim=-3)
    return n, d, r, s


def join_maps(normals, diffuse, roughness, specular):
    """Inverse of :func:`split_maps`.  Reference: ``pack_svbrdf`` utils.py:36-38."""
    return torch.cat((normals, diffuse, roughness, specular), dim=-3)


# ----------------------------------------------------------------------------------------
# BRDF terms (renderers.py:18-65)
# ----------------------------------------------------------------------------------------

def ggx_distribution(roughness, nh):
    """GGX normal distribution.  Reference: ``compute_microfacet_distribution``
    renderers.py:22-27 (alpha = roughness^2; denominator clamped at 1e-3 before squaring)."""
    alpha = roughness ** 2
    alpha_sq = alpha ** 2
    nh_sq = nh ** 2
    den = torch.clamp(nh_sq * (alpha_sq + (1 - nh_sq) / nh_sq), min=0.001)
    return (alpha_sq * positive_mask(nh)) / (math.pi * den ** 2)


def schlick_fresnel(specular, vh):
    """Schlick approximation with exponent 5.  Reference: ``compute_fresnel`` renderers.py:29-32."""
    return specular + (1.0 - specular) * (1.0 - vh) ** 5


def smith_g1(roughness, xh, xn):
    """One-direction Smith-GGX masking.  Reference: ``compute_g1`` renderers.py:34-38."""
    alpha = roughness ** 2
    alpha_sq = alpha ** 2
    xn_sq = xn ** 2
    return 2 * positive_mask(xh / xn) / (1 + torch.sqrt(1 + alpha_sq * (1.0 - xn_sq) / xn_sq))


def cook_torrance(wi, wo, normals, diffuse, roughness, specular):
    """Diffuse + specular BRDF value.  Reference: ``compute_specular_term`` renderers.py:43-60,
    ``compute_diffuse_term`` :18-20 and ``evaluate_brdf`` :62-65 (k_s := Fresnel term)."""
    half = unit((wi + wo) / 2.0)
    nh = torch.clamp(channel_dot(normals, half), min=0.001)
    vh = torch.clamp(channel_dot(wo, half), min=0.001)
    lh = torch.clamp(channel_dot(wi, half), min=0.001)
    vn = torch.clamp(channel_dot(wo, normals), min=0.001)
    ln = torch.clamp(channel_dot(wi, normals), min=0.001)

    fres = schlick_fresnel(specular, vh)
    geom = smith_g1(roughness, vh, vn) * smith_g1(roughness, lh, ln)
    dist = ggx_distribution(roughness, nh)
    spec_term = fres * geom * dist / (4.0 * vn * ln)
    diff_term = (1.0 - fres) * diffuse / math.pi
    return diff_term + spec_term


# ----------------------------------------------------------------------------------------
# renderer (renderers.py:67-104)
# ----------------------------------------------------------------------------------------

def _as_column(v, like):
    """len-3 list / ndarray / tensor -> [3,1,1] tensor in the dtype/device of ``like``.
    Reference: renderers.py:79,91 (``torch.Tensor(pos).unsqueeze(-1).unsqueeze(-1)``); the
    reference builds an fp32 tensor first, which this keeps so fp64 runs see the same
    fp32-rounded positions."""
    t = torch.as_tensor(v).detach().to(torch.float32).reshape(3, 1, 1)
    return t.to(device=like.device, dtype=like.dtype)


def patch_coords(maps):
    """Surface points of the [-1,1]^2 patch, z = 0: x = linspace(-1,1,W)[col], y = -linspace[row].
    Reference: renderers.py:73-76 (square maps only)."""
    h, w = maps.shape[-2], maps.shape[-1]
    if h != w:
        raise ValueError("the reference renderer only supports square maps (got %dx%d)" % (h, w))
    row = torch.linspace(-1, 1, w, device=maps.device, dtype=maps.dtype)
    xs = row.unsqueeze(0).expand(h, w).unsqueeze(0)
    ys = -1 * xs.transpose(1, 2)
    return torch.cat((xs, ys, torch.zeros_like(xs)), dim=0)


def render(camera_pos, light_pos, light_color, maps):
    """Radiance of the patch under one point light seen from one camera.

    maps [12,H,W] -> [1,3,H,W]; [B,12,H,W] -> [B,3,H,W].  Reference:
    ``LocalRenderer.render`` renderers.py:67-104."""
    coords = patch_coords(maps)
    wo = unit(_as_column(camera_pos, maps) - coords)
    normals, diffuse, roughness, specular = split_maps(maps)
    roughness = torch.clamp(roughness, min=0.001)
    to_light = _as_column(light_pos, maps) - coords
    wi = unit(to_light)
    f = cook_torrance(wi, wo, normals, diffuse, roughness, specular)
    ln = torch.clamp(channel_dot(wi, normals), min=0.0)
    color = _as_column(light_color, maps).unsqueeze(0)
    falloff = 1.0 / torch.sqrt(channel_dot(to_light, to_light)) ** 2
    return (f * (color * falloff)) * ln


# ----------------------------------------------------------------------------------------
# scene sampling (environment.py:18-55, utils.py:100-111) -- uses the global CPU generator
# in the reference's draw order, so torch.manual_seed(s) reproduces the reference's scenes
# ----------------------------------------------------------------------------------------

def cosine_hemisphere_directions(count, min_eps=0.001, max_eps=0.05):
    """Reference: ``generate_normalized_random_direction`` utils.py:100-111."""
    r1 = torch.empty(count, 1, dtype=torch.float32).uniform_(0.0 + min_eps, 1.0 - max_eps)
    r2 = torch.empty(count, 1, dtype=torch.float32).uniform_(0.0, 1.0)
    r = torch.sqrt(r1)
    phi = 2 * math.pi * r2
    return torch.cat([r * torch.cos(phi), r * torch.sin(phi), torch.sqrt(1.0 - r ** 2)], dim=-1)


def sample_random_configs(count):
    """-> (cam[count,3], light[count,3], color[count,3]).  Reference:
    ``generate_random_scenes`` environment.py:18-30 (unit-distance positions, colour 20)."""
    cam = cosine_hemisphere_directions(count, 0.001, 0.1)
    light = cosine_hemisphere_directions(count, 0.001, 0.1)
    return cam, light, torch.full((count, 3), 20.0)


def sample_specular_configs(count):
    """Mirror configurations.  Reference: ``generate_specular_scenes`` environment.py:32-55."""
    view = cosine_hemisphere_directions(count, 0.001, 0.1)
    mirror = view * torch.tensor([-1.0, -1.0, 1.0]).unsqueeze(0)
    dist_view = torch.exp(torch.empty(count, 1, dtype=torch.float32).normal_(mean=0.5, std=0.75))
    dist_light = torch.exp(torch.empty(count, 1, dtype=torch.float32).normal_(mean=0.5, std=0.75))
    shift = torch.cat([torch.empty(count, 2, dtype=torch.float32).uniform_(-1.0, 1.0),
                       torch.zeros((count, 1)) + 0.0001], dim=-1)
    return view * dist_view + shift, mirror * dist_light + shift, torch.full((count, 3), 50.0)


def sample_loss_configs(batch, n_random=3, n_specular=6):
    """Scenes of one ``RenderingLoss.forward`` call, drawn per batch element in the
    reference's order (losses.py:34-35).  -> float32 tensor [batch, n_random+n_specular, 9]
    with rows (cam xyz, light xyz, colour rgb)."""
    out = torch.empty(batch, n_random + n_specular, 9, dtype=torch.float32)
    for b in range(batch):
        rc, rl, rk = sample_random_configs(n_random)
        sc, sl, sk = sample_specular_configs(n_specular)
        out[b, :, 0:3] = torch.cat((rc, sc), dim=0)
        out[b, :, 3:6] = torch.cat((rl, sl), dim=0)
        out[b, :, 6:9] = torch.cat((rk, sk), dim=0)
    return out


# ----------------------------------------------------------------------------------------
# losses (losses.py:7-63)
# ----------------------------------------------------------------------------------------

def render_batch(maps, configs):
    """maps [B,12,H,W], configs [B,N,9] -> [B,N,3,H,W]; element b is rendered under its own
    N configurations, one eager ``render`` call each (losses.py:34-44)."""
    per_sample = []
    for b in range(maps.shape[0]):
        views = [render(configs[b, k, 0:3], configs[b, k, 3:6], configs[b, k, 6:9], maps[b])
                 for k in range(configs.shape[1])]
        per_sample.append(torch.cat(views, dim=0))
    return torch.stack(per_sample, dim=0)


def rendering_loss(input_maps, target_maps, configs):
    """mean |log(R_in + 0.1) - log(R_tgt + 0.1)| over [B,N,3,H,W].  Reference:
    ``RenderingLoss.forward`` losses.py:29-52 with the scenes given explicitly."""
    a = torch.log(render_batch(input_maps, configs) + 0.1)
    b = torch.log(render_batch(target_maps, configs) + 0.1)
    return torch.nn.functional.l1_loss(a, b)


def maps_l1_loss(input_maps, target_maps):
    """Reference: ``SVBRDFL1Loss.forward`` losses.py:7-19 (log with eps 0.01 on diffuse/specular)."""
    n0, d0, r0, s0 = split_maps(input_maps)
    n1, d1, r1, s1 = split_maps(target_maps)
    l1 = torch.nn.functional.l1_loss
    eps = 0.01
    return (l1(n0, n1) + l1(torch.log(d0 + eps), torch.log(d1 + eps))
            + l1(r0, r1) + l1(torch.log(s0 + eps), torch.log(s1 + eps)))


def mixed_loss(input_maps, target_maps, configs, l1_weight=0.1):
    """Reference: ``MixedLoss.forward`` losses.py:62-63."""
    return l1_weight * maps_l1_loss(input_maps, target_maps) + rendering_loss(input_maps, target_maps, configs)


def decode_network_output(encoded):
    """[...,9,H,W] network output in [-1,1] (normal xy, diffuse, roughness, specular) -> [...,12,H,W] maps.
    Reference: ``decode_svbrdf`` utils.py:73-88 (normal = normalize(3x, 3y, 1), roughness repeated x3) followed
    by the [0,1] mapping of diffuse / roughness / specular in ``SingleViewModel.forward`` models.py:340-346
    (``encode_as_unit_interval`` utils.py:92-93)."""
    nxy, diffuse, rough, spec = torch.split(encoded, (2, 3, 1, 3), dim=-3)
    reps = [1] * encoded.dim()
    reps[-3] = 3
    rough = rough.repeat(reps)
    nx, ny = torch.split(nxy.mul(3.0), 1, dim=-3)
    normals = torch.cat([nx, ny, torch.ones_like(nx)], dim=-3)
    normals = torch.div(normals, torch.sqrt(torch.sum(torch.pow(normals, 2.0), dim=-3, keepdim=True)))
    return join_maps(normals, (diffuse + 1) / 2, (rough + 1) / 2, (spec + 1) / 2)


def rendering_loss_and_grad(input_maps, target_maps, configs):
    """Convenience for the parity tests: loss value and d loss / d input via autograd
    (what ``loss.backward()`` gives the reference at main.py:116-117)."""
    x = input_maps.detach().clone().requires_grad_(True)
    loss = rendering_loss(x, target_maps.detach(), configs)
    (g,) = torch.autograd.grad(loss, x)
    return loss.detach(), g
