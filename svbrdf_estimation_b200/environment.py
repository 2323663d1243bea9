"""Scene records and the light/view configuration samplers of the rendering loss.

Drop-in for the reference's ``environment.py`` (development/multiImage_pytorch/environment.py:4-55):
``Camera``/``Light``/``Scene`` holders and ``generate_random_scenes`` / ``generate_specular_scenes``
consume the global CPU generator in the reference's order, so ``torch.manual_seed(s)`` yields the
same scenes as the reference does.  The CUDA path consumes scenes as packed float32 records
``[.., 9] = (camera xyz, light xyz, light rgb)``; ``sample_loss_configs`` produces a whole
``[B, N, 9]`` block for one loss evaluation without building Python objects.
"""
import math

import numpy as np
import torch

from .utils import generate_normalized_random_direction, hemisphere_uniforms_to_directions

RANDOM_LIGHT_COLOR = 20.0     # environment.py:27
SPECULAR_LIGHT_COLOR = 50.0   # environment.py:52
VIEW_EPS = (0.001, 0.1)       # environment.py:20-21,34
LOG_DISTANCE = (0.5, 0.75)    # environment.py:38-39


class Camera:
    """Pinhole camera looking at the patch centre; only the position matters (environment.py:4-6)."""
    __slots__ = ("pos",)

    def __init__(self, pos):
        self.pos = pos

    def __repr__(self):
        return "Camera(pos=%s)" % (_triple(self.pos),)


class Light:
    """Point light with RGB intensity (environment.py:8-11)."""
    __slots__ = ("pos", "color")

    def __init__(self, pos, color):
        self.pos, self.color = pos, color

    def __repr__(self):
        return "Light(pos=%s, color=%s)" % (_triple(self.pos), _triple(self.color))


class Scene:
    """One light/view configuration (environment.py:13-16)."""
    __slots__ = ("camera", "light")

    def __init__(self, camera, light):
        self.camera, self.light = camera, light

    def __repr__(self):
        return "Scene(%r, %r)" % (self.camera, self.light)


def _triple(v):
    """list / tuple / ndarray / tensor of 3 numbers -> list of 3 python floats rounded to fp32
    (the reference converts with ``torch.Tensor(v)``, i.e. to fp32: renderers.py:79,91,98)."""
    if isinstance(v, torch.Tensor):
        a = v.detach().to(device="cpu", dtype=torch.float32).reshape(-1).numpy()
    else:
        a = np.asarray(v, dtype=np.float32).reshape(-1)
    if a.shape[0] != 3:
        raise ValueError("expected 3 components, got %d" % a.shape[0])
    return [float(a[0]), float(a[1]), float(a[2])]


def scene_record(scene):
    """Scene -> [9] float32 CPU tensor (camera xyz, light xyz, light rgb)."""
    return torch.tensor(_triple(scene.camera.pos) + _triple(scene.light.pos) + _triple(scene.light.color),
                        dtype=torch.float32)


def pack_scenes(scenes):
    """Iterable of Scene -> [N,9] float32 CPU tensor."""
    return torch.stack([scene_record(s) for s in scenes], dim=0)


def unpack_scenes(records):
    """[N,9] tensor -> list of Scene objects holding fp32 CPU tensors (like the reference's samplers)."""
    records = torch.as_tensor(records, dtype=torch.float32).reshape(-1, 9)
    return [Scene(Camera(r[0:3].clone()), Light(r[3:6].clone(), r[6:9].tolist())) for r in records]


# ---- raw draws in the reference's order ------------------------------------------------------

def _draw_random(view_u, light_u):
    """Fills [count,2] buffers with (r1, r2) for the view and then the light directions
    (environment.py:20-21 -> utils.py:101-102: all r1 first, then all r2)."""
    lo, hi = 0.0 + VIEW_EPS[0], 1.0 - VIEW_EPS[1]
    view_u[0].uniform_(lo, hi); view_u[1].uniform_(0.0, 1.0)
    light_u[0].uniform_(lo, hi); light_u[1].uniform_(0.0, 1.0)


def _draw_specular(view_u, log_dist, shift):
    """(r1, r2) of the view direction, log-distances of view and light, xy shift
    (environment.py:34,38-39,44)."""
    view_u[0].uniform_(0.0 + VIEW_EPS[0], 1.0 - VIEW_EPS[1]); view_u[1].uniform_(0.0, 1.0)
    log_dist[0].normal_(mean=LOG_DISTANCE[0], std=LOG_DISTANCE[1])
    log_dist[1].normal_(mean=LOG_DISTANCE[0], std=LOG_DISTANCE[1])
    shift.uniform_(-1.0, 1.0)


def _random_records(view_u, light_u):
    """[...,2,count] uniforms -> [...,count,9] records: unit-distance camera and light, colour 20
    (environment.py:18-30)."""
    cam = hemisphere_uniforms_to_directions(view_u[..., 0, :], view_u[..., 1, :])
    light = hemisphere_uniforms_to_directions(light_u[..., 0, :], light_u[..., 1, :])
    return torch.cat((cam, light, torch.full_like(cam, RANDOM_LIGHT_COLOR)), dim=-1)


def _specular_records(view_u, log_dist, shift):
    """Mirror configurations (environment.py:32-55): light direction = view * (-1,-1,1), independent
    log-normal distances, common xy shift with z = 1e-4, colour 50."""
    view = hemisphere_uniforms_to_directions(view_u[..., 0, :], view_u[..., 1, :])
    mirror = view * torch.tensor([-1.0, -1.0, 1.0])
    dist = torch.exp(log_dist)
    off = torch.cat((shift, torch.zeros_like(shift[..., :1]) + 0.0001), dim=-1)
    cam = view * dist[..., 0, :, None] + off
    light = mirror * dist[..., 1, :, None] + off
    return torch.cat((cam, light, torch.full_like(cam, SPECULAR_LIGHT_COLOR)), dim=-1)


def _affine_like_torch(raw, lo, hi):
    """What ``tensor.uniform_(lo, hi)`` computes from the raw 24-bit uniform ``raw = uniform_(0, 1)``:
    ATen evaluates ``x * (to - from) + from`` with x in double, ``to - from`` in float, and rounds once."""
    lo32, hi32 = np.float32(lo), np.float32(hi)
    return (raw.double() * float(hi32 - lo32) + float(lo32)).float()


def _draws_through_torch(batch, nr, ns, head, tail):
    """The raw draws of one loss evaluation through torch's own generator calls: 2-3 calls per batch element (consecutive
    ``uniform_`` draws are taken as one raw ``uniform_(0, 1)`` block, the two ``normal_`` draws are merged while that
    keeps ATen on its scalar sampling path)."""
    uni = torch.empty(batch, head + tail)
    nrm = torch.empty(batch, 2, ns)
    # uniform blocks in stream order: element 0's directions, then (shift of b + directions of b+1), last shift
    sizes = [head] + [tail + head] * (batch - 1) + [tail]
    ublocks = uni.view(-1).split(sizes)
    if 0 < 2 * ns < 16:
        nblocks = nrm.view(batch, 2 * ns).unbind(0)          # one merged normal_ call per element
    else:
        nblocks = nrm.view(2 * batch, ns).unbind(0)          # two calls per element, as in the reference
    per_elem = len(nblocks) // batch if ns > 0 else 0
    mean, std = LOG_DISTANCE
    if head > 0:
        ublocks[0].uniform_(0.0, 1.0)
    for b in range(batch):
        for j in range(per_elem):
            nblocks[b * per_elem + j].normal_(mean, std)
        if sizes[b + 1] > 0:
            ublocks[b + 1].uniform_(0.0, 1.0)
    return uni, nrm


def _records_native(batch, nr, ns):
    """Scene records through the library's restatement of ATen's CPU generator and of the samplers' float arithmetic
    (csrc/scene_sampler.cpp): the draws of the whole batch come from ONE call on the serialised generator state, which
    is stored back so torch continues where the reference would; sqrt / cos / sin / exp are torch's own, applied once to the
    whole batch, so every value is rounded exactly as in the reference."""
    from . import _cabi
    lib = _cabi.lib()
    nd = 2 * nr + ns
    state = torch.get_rng_state()
    r1, phi = torch.empty(batch, nd), torch.empty(batch, nd)
    logd, shift = torch.empty(batch, 2, ns), torch.empty(batch, ns, 2)
    _cabi.check(lib.svbrdf_b200_reference_scenes_begin(state.data_ptr(), state.numel(), batch, nr, ns, r1.data_ptr(),
                                                       phi.data_ptr(), logd.data_ptr(), shift.data_ptr()))
    torch.set_rng_state(state)
    radius = torch.sqrt(r1)                                                         # utils.py:104
    cos_phi, sin_phi = torch.cos(phi), torch.sin(phi)                               # utils.py:107-108
    z = torch.sqrt(1.0 - radius * radius)                                           # utils.py:109
    dist = torch.exp(logd)                                                          # environment.py:38-39
    out = torch.empty(batch, nr + ns, 9)
    _cabi.check(lib.svbrdf_b200_reference_scenes_finish(batch, nr, ns, radius.data_ptr(), cos_phi.data_ptr(), sin_phi.data_ptr(),
                                                        z.data_ptr(), dist.data_ptr(), shift.data_ptr(), out.data_ptr()))
    return out


def sample_loss_configs(batch, n_random=3, n_specular=6, native_draws=True):
    """Scene records of one ``RenderingLoss.forward`` call: for each batch element, ``n_random``
    random then ``n_specular`` mirror configurations (losses.py:34-35) -> [batch, N, 9] float32 CPU.

    The global CPU generator is consumed in exactly the reference's order (a seed reproduces the
    reference's scenes bit for bit, tests/golden/scenes.npz, and the generator is left in the state the
    reference would leave it in).  ``native_draws``: the library's restatement of ATen's mt19937 / uniform /
    normal sampling and of the samplers' float arithmetic does the whole batch in two C calls around three torch
    ops (~0.1 ms for 64 elements; the reference's own loop takes 23 ms).  Otherwise - and whenever ``normal_``
    would take ATen's vectorised path, n_specular >= 16 - the draws go through torch's generator calls, 2-3 per
    batch element, and the arithmetic through batched torch ops."""
    nr, ns = int(n_random), int(n_specular)
    if batch <= 0 or nr + ns <= 0:
        raise ValueError("batch and the number of configurations must be positive")
    if native_draws and ns < 16:
        return _records_native(batch, nr, ns)
    head, tail = 4 * nr + 2 * ns, 2 * ns            # per element: [view/light r1 r2 | spec r1 r2] ... normals ... [shift]
    uni, nrm = _draws_through_torch(batch, nr, ns, head, tail)
    plan = _sampler_plan(nr, ns)
    # every uniform_(lo, hi) of the reference as ATen computes it from the raw draw: x*(hi-lo)+lo with x in
    # double, (hi-lo) in float, one rounding to float (uniform_real_distribution) - for all columns at once
    u = (uni.double() * plan["scale"] + plan["offset"]).float()
    r = torch.sqrt(u.index_select(1, plan["r1"]))                                      # utils.py:104
    phi = 2 * math.pi * u.index_select(1, plan["r2"])                                  # utils.py:105
    dirs = torch.stack((r * torch.cos(phi), r * torch.sin(phi), torch.sqrt(1.0 - r ** 2)), dim=-1)   # [B, 2nr+ns, 3]
    parts = []
    if nr > 0:
        parts.append(torch.cat((dirs[:, :nr], dirs[:, nr:2 * nr], plan["col_r"].expand(batch, nr, 3)), dim=-1))
    if ns > 0:
        view_s = dirs[:, 2 * nr:]
        off = torch.cat((u[:, head:].view(batch, ns, 2), plan["z"].expand(batch, ns, 1)), dim=-1)   # environment.py:44
        dist = torch.exp(nrm)                                                                       # environment.py:38-39
        cam = view_s * dist[:, 0, :, None] + off                                                    # environment.py:46
        light = (view_s * plan["mirror"]) * dist[:, 1, :, None] + off                               # environment.py:35,47
        parts.append(torch.cat((cam, light, plan["col_s"].expand(batch, ns, 3)), dim=-1))
    return torch.cat(parts, dim=1).contiguous()


_PLANS = {}


def _sampler_plan(nr, ns):
    """Column bookkeeping of one batch element's uniform block, cached per (n_random, n_specular)."""
    key = (nr, ns)
    plan = _PLANS.get(key)
    if plan is None:
        head = 4 * nr + 2 * ns
        lo32, hi32 = np.float32(0.0 + VIEW_EPS[0]), np.float32(1.0 - VIEW_EPS[1])
        scale = torch.ones(head + 2 * ns, dtype=torch.float64)
        offset = torch.zeros(head + 2 * ns, dtype=torch.float64)
        r1_cols = list(range(0, nr)) + list(range(2 * nr, 3 * nr)) + list(range(4 * nr, 4 * nr + ns))
        r2_cols = list(range(nr, 2 * nr)) + list(range(3 * nr, 4 * nr)) + list(range(4 * nr + ns, head))
        scale[r1_cols] = float(hi32 - lo32)
        offset[r1_cols] = float(lo32)
        scale[head:] = float(np.float32(1.0) - np.float32(-1.0))
        offset[head:] = -1.0
        plan = {"scale": scale, "offset": offset,
                "r1": torch.tensor(r1_cols, dtype=torch.long), "r2": torch.tensor(r2_cols, dtype=torch.long),
                "col_r": torch.full((1, 1, 3), RANDOM_LIGHT_COLOR), "col_s": torch.full((1, 1, 3), SPECULAR_LIGHT_COLOR),
                "z": torch.zeros(1, 1, 1) + 0.0001, "mirror": torch.tensor([-1.0, -1.0, 1.0])}
        _PLANS[key] = plan
    return plan


def sample_loss_configs_fast(batch, n_random=3, n_specular=6, generator=None):
    """Same distributions as :func:`sample_loss_configs` with one generator call per quantity for
    the whole batch (not draw-order compatible with the reference)."""
    def uni(shape, lo, hi):
        return torch.empty(shape).uniform_(lo, hi, generator=generator)
    lo, hi = VIEW_EPS[0], 1.0 - VIEW_EPS[1]
    def dirs_u(n):
        return torch.stack((uni((batch, n), lo, hi), uni((batch, n), 0.0, 1.0)), dim=1)
    log_dist = torch.empty(batch, 2, n_specular).normal_(LOG_DISTANCE[0], LOG_DISTANCE[1], generator=generator)
    shift = uni((batch, n_specular, 2), -1.0, 1.0)
    return torch.cat((_random_records(dirs_u(n_random), dirs_u(n_random)),
                      _specular_records(dirs_u(n_specular), log_dist, shift)), dim=1).contiguous()


def sample_loss_configs_native(batch, n_random=3, n_specular=6, seed=0, first_batch_element=0):
    """Scene records from the library's stateless host sampler (``svbrdf_b200_sample_scenes``): the
    scenes of batch element ``e`` depend only on ``(seed, e)``, which makes sharded runs reproducible
    regardless of how the batch is split over ranks.  ~50x faster than the reference-order sampler;
    same distributions, different random stream."""
    from . import _cabi
    out = torch.empty(batch, n_random + n_specular, 9, dtype=torch.float32)
    _cabi.check(_cabi.lib().svbrdf_b200_sample_scenes(int(seed) & (2 ** 64 - 1), int(first_batch_element), int(batch),
                                                      int(n_random), int(n_specular), out.data_ptr()))
    return out


class NativeSceneSampler:
    """Callable ``(batch, n_random, n_specular) -> [B,N,9]`` for ``RenderingLoss(renderer, scene_sampler=...)``:
    fresh scenes on every call (a call counter is mixed into the seed), drawn by the library's host
    sampler.  ``first_batch_element`` is the global index of this rank's first sample, so that sharded
    runs draw exactly the scenes an unsharded run would."""

    def __init__(self, seed=313, first_batch_element=0):
        self.seed, self.first_batch_element, self.calls = int(seed), int(first_batch_element), 0

    def __call__(self, batch, n_random=3, n_specular=6):
        step_seed = (self.seed * 0x9E3779B97F4A7C15 + self.calls * 0xD1342543DE82EF95) & (2 ** 64 - 1)
        self.calls += 1
        return sample_loss_configs_native(batch, n_random, n_specular, step_seed, self.first_batch_element)


def generate_random_scenes(count):
    """``count`` Scenes with independently cosine-sampled view and light directions used as
    positions, light colour 20 (environment.py:18-30)."""
    view_u, light_u = torch.empty(2, count), torch.empty(2, count)
    _draw_random(view_u, light_u)
    return unpack_scenes(_random_records(view_u, light_u))


def generate_specular_scenes(count):
    """``count`` Scenes in mirror configuration, light colour 50 (environment.py:32-55)."""
    view_u, log_dist, shift = torch.empty(2, count), torch.empty(2, count), torch.empty(count, 2)
    _draw_specular(view_u, log_dist, shift)
    return unpack_scenes(_specular_records(view_u, log_dist, shift))


__all__ = ["Camera", "Light", "Scene", "generate_random_scenes", "generate_specular_scenes",
           "generate_normalized_random_direction", "pack_scenes", "unpack_scenes", "scene_record",
           "sample_loss_configs", "sample_loss_configs_fast", "sample_loss_configs_native", "NativeSceneSampler"]
