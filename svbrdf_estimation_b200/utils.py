"""Layout contract and direction sampler of the hot path.

Mirrors the part of the reference's ``utils.py`` the rendering-loss path depends on
(development/multiImage_pytorch/utils.py:36-58 and :100-111); image IO, gamma and cropping
helpers of that file are outside the path and are not provided.
"""
import math

import torch

CHANNELS = 12   # normals(3) diffuse(3) roughness(3) specular(3) on dim -3 (utils.py:36-58)


def pack_svbrdf(normals, diffuse, roughness, specular):
    """Concatenate the four 3-channel maps on dim -3 (utils.py:36-38)."""
    return torch.cat((normals, diffuse, roughness, specular), dim=-3)


def unpack_svbrdf(svbrdf, is_encoded=False):
    """Split a packed SVBRDF on dim -3 (utils.py:40-58).  ``is_encoded`` selects the 9-channel
    network encoding (normal xy, diffuse, 1 roughness, specular)."""
    sizes = (2, 3, 1, 3) if is_encoded else (3, 3, 3, 3)
    if svbrdf.shape[-3] != sum(sizes):
        raise ValueError("expected %d channels on dim -3, got %d" % (sum(sizes), svbrdf.shape[-3]))
    return torch.split(svbrdf, sizes, dim=-3)


def encode_as_unit_interval(tensor):
    """[-1,1] -> [0,1] (utils.py:92-93)."""
    return (tensor + 1) / 2


def decode_from_unit_interval(tensor):
    """[0,1] -> [-1,1] (utils.py:97-98)."""
    return tensor * 2 - 1


def decode_svbrdf(svbrdf):
    """9-channel network encoding -> 12 channels: normal = normalize(3x, 3y, 1), roughness repeated x3
    (utils.py:73-88).  Plain torch ops; the fused loss (``MixedLoss.forward_encoded``) does this in-kernel."""
    nxy, diffuse, roughness, specular = unpack_svbrdf(svbrdf, True)
    reps = [1] * diffuse.dim()
    reps[-3] = 3
    nx, ny = torch.split(nxy.mul(3.0), 1, dim=-3)
    normals = torch.cat([nx, ny, torch.ones_like(nx)], dim=-3)
    normals = normals / torch.sqrt(torch.sum(normals * normals, dim=-3, keepdim=True))
    return pack_svbrdf(normals, diffuse, roughness.repeat(reps), specular)


def decode_network_output(encoded):
    """What ``SingleViewModel.forward`` does after the tanh (models.py:338-346): ``decode_svbrdf`` and the
    [0,1] mapping of diffuse, roughness and specular."""
    n, d, r, s = unpack_svbrdf(decode_svbrdf(encoded))
    return pack_svbrdf(n, encode_as_unit_interval(d), encode_as_unit_interval(r), encode_as_unit_interval(s))


def hemisphere_uniforms_to_directions(r1, r2):
    """Cosine-weighted hemisphere direction from two uniforms (utils.py:104-111); any shape, last
    axis of the result is xyz."""
    r = torch.sqrt(r1)
    phi = 2 * math.pi * r2
    return torch.stack((r * torch.cos(phi), r * torch.sin(phi), torch.sqrt(1.0 - r ** 2)), dim=-1)


def generate_normalized_random_direction(count, min_eps=0.001, max_eps=0.05):
    """[count,3] unit vectors on the upper hemisphere drawn from the global CPU generator with the
    reference's draw order: ``count`` r1 values in [min_eps, 1-max_eps), then ``count`` r2 values in
    [0,1) (utils.py:100-111)."""
    r1 = torch.empty(count, dtype=torch.float32).uniform_(0.0 + min_eps, 1.0 - max_eps)
    r2 = torch.empty(count, dtype=torch.float32).uniform_(0.0, 1.0)
    return hemisphere_uniforms_to_directions(r1, r2)
