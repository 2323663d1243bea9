"""In-network renderer behind the reference's plugin interface, running on sm_100a CUDA kernels.

Drop-in for ``renderers.LocalRenderer`` of the reference
(development/multiImage_pytorch/renderers.py:14-104): ``LocalRenderer().render(scene, svbrdf)``
shades a flat material patch under one point light seen from one camera with a Cook-Torrance /
GGX BRDF and is differentiable w.r.t. ``svbrdf``.  The arithmetic happens in
``libsvbrdf_b200.so`` (csrc/kernels.cu, csrc/shading.cuh); this module only validates arguments,
owns the buffers and wires the kernels into autograd.  There is no CPU implementation: tensors
that live on the host are staged through the GPU, and without a CUDA device every call raises.
"""
import torch

from . import _cabi
from .environment import scene_record

_LIN_CACHE = {}


def coordinate_table(width, device):
    """``torch.linspace(-1, 1, W)`` on ``device`` - the patch coordinates of renderers.py:73.
    Computed by torch itself (cached) so pixel positions are bit-identical to the reference's."""
    key = (int(width), device.type, device.index)
    lin = _LIN_CACHE.get(key)
    if lin is None:
        lin = torch.linspace(-1, 1, int(width), device=device, dtype=torch.float32)
        _LIN_CACHE[key] = lin
    return lin


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("svbrdf_estimation_b200 needs a CUDA device: the rendering path has no CPU fallback")


class Origin:
    """Where a caller's tensor lived and in which dtype, so results can follow it (the reference's outputs have the
    device and dtype of its inputs, renderers.py:68)."""
    __slots__ = ("device", "dtype")

    def __init__(self, device, dtype):
        self.device, self.dtype = device, dtype

    def restore(self, t):
        """Result tensor -> the caller's device; float64 callers get float64 back (computed in fp32)."""
        if self.device.type != "cuda":
            t = t.to(self.device)
        return t.to(torch.float64) if self.dtype == torch.float64 else t


def as_device_maps(svbrdf, what="svbrdf"):
    """Validate a packed SVBRDF tensor and return (maps[B,12,H,W] fp32 contiguous on CUDA,
    leading shape, Origin).  The kernels compute in fp32: half-precision maps (a network output under autocast)
    are upcast, float64 maps are computed in fp32 and the results returned as float64."""
    if not isinstance(svbrdf, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % what)
    if svbrdf.dim() < 3 or svbrdf.shape[-3] != 12:
        raise ValueError("%s must have shape [...,12,H,W], got %s" % (what, tuple(svbrdf.shape)))
    if svbrdf.shape[-1] != svbrdf.shape[-2]:
        raise ValueError("only square maps are supported (renderers.py:73-76), got %dx%d"
                         % (svbrdf.shape[-2], svbrdf.shape[-1]))
    if not svbrdf.dtype.is_floating_point:
        raise TypeError("%s must be a floating-point tensor, got %s" % (what, svbrdf.dtype))
    require_cuda()
    origin = Origin(svbrdf.device, svbrdf.dtype)
    lead = tuple(svbrdf.shape[:-3])
    if svbrdf.dtype != torch.float32:
        svbrdf = svbrdf.float()          # differentiable: the gradient comes back in the caller's dtype
    maps = svbrdf if origin.device.type == "cuda" else host_to_device(svbrdf)
    maps = maps.reshape((-1,) + tuple(svbrdf.shape[-3:])).contiguous()
    if maps.shape[0] == 0:
        raise ValueError("%s has an empty batch" % what)
    return maps, lead, origin


def host_to_device(t):
    """Host tensor -> current CUDA device (differentiable).  Small tensors and pinned tensors are one copy; large
    pageable tensors go through ``staging.upload`` (pinned double buffer, CPU copy overlapped with the DMA)."""
    if t.numel() * t.element_size() < (8 << 20) or t.is_pinned() or not t.is_contiguous():
        return t.cuda(non_blocking=t.is_pinned())
    from . import staging
    return staging.upload_with_grad(t) if t.requires_grad else staging.upload(t)


def as_host_records(records, batch=None):
    """-> contiguous float32 CPU tensor [N,9] or [B,N,9]."""
    if isinstance(records, torch.Tensor) and records.device.type == "cpu" and records.dtype == torch.float32 \
            and records.is_contiguous() and not records.requires_grad:
        rec = records
    else:
        rec = torch.as_tensor(records).detach().to(device="cpu", dtype=torch.float32).contiguous()
    if rec.dim() not in (2, 3) or rec.shape[-1] != 9 or rec.shape[-2] == 0:
        raise ValueError("scene records must have shape [N,9] or [B,N,9], got %s" % (tuple(rec.shape),))
    if rec.dim() == 3 and batch is not None and rec.shape[0] != batch:
        raise ValueError("scene records are for %d batch elements, maps have %d" % (rec.shape[0], batch))
    return rec


class _RenderRecords(torch.autograd.Function):
    """images[B,N,3,H,W] = render(maps[B,12,H,W]) under host scene records."""

    @staticmethod
    def forward(ctx, maps, records):
        B, _, H, W = maps.shape
        per_batch = records.dim() == 3
        N = records.shape[-2]
        images = torch.empty((B, N, 3, H, W), device=maps.device, dtype=torch.float32)
        lin = coordinate_table(W, maps.device)
        with torch.cuda.device(maps.device):
            _cabi.check(_cabi.lib().svbrdf_b200_render_forward(
                maps.data_ptr(), B, H, W, records.data_ptr(), N, int(per_batch), lin.data_ptr(),
                images.data_ptr(), torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(maps)
        ctx.records = records
        return images

    @staticmethod
    def backward(ctx, grad_images):
        (maps,) = ctx.saved_tensors
        records = ctx.records
        B, _, H, W = maps.shape
        grad_images = grad_images.contiguous()
        if grad_images.dtype != torch.float32:
            grad_images = grad_images.float()
        grad_maps = torch.empty_like(maps)
        lin = coordinate_table(W, maps.device)
        with torch.cuda.device(maps.device):
            _cabi.check(_cabi.lib().svbrdf_b200_render_backward(
                maps.data_ptr(), B, H, W, records.data_ptr(), records.shape[-2], int(records.dim() == 3),
                lin.data_ptr(), grad_images.data_ptr(), grad_maps.data_ptr(),
                torch.cuda.current_stream().cuda_stream))
        return grad_maps, None


def render_records(svbrdf, records):
    """Batched form of the renderer: ``svbrdf [...,12,H,W]`` under ``records`` ([N,9] shared by the
    whole batch, or [B,N,9] per flattened batch element) -> ``[...,N,3,H,W]`` on ``svbrdf``'s device."""
    maps, lead, origin = as_device_maps(svbrdf)
    rec = as_host_records(records, maps.shape[0])
    images = _RenderRecords.apply(maps, rec)
    images = images.reshape(lead + tuple(images.shape[1:]))
    return origin.restore(images)


class LocalRenderer:
    """Same interface as the reference's ``LocalRenderer`` (renderers.py:14-104, no-argument
    constructor, ``render(scene, svbrdf)``); ``RenderingLoss`` recognises it and uses the fused
    loss kernels instead of calling ``render`` 2*N times per batch element."""

    fused_rendering_loss = True

    def render(self, scene, svbrdf):
        """``svbrdf [12,H,W] -> [1,3,H,W]``; ``[B,12,H,W] -> [B,3,H,W]`` (one scene for the whole
        batch); linear, unclamped radiance; differentiable w.r.t. ``svbrdf`` (renderers.py:67-104)."""
        images = render_records(svbrdf, scene_record(scene).unsqueeze(0))   # [...,1,3,H,W]
        images = images.squeeze(-4)
        return images.unsqueeze(0) if images.dim() == 3 else images


_PATH_TRACER = None


def register_path_tracer(cls):
    """Makes ``RednerRenderer()`` construct ``cls`` - e.g. the reference's own ``renderers.RednerRenderer`` imported
    from its module under another name before ``sys.modules['renderers']`` is swapped."""
    global _PATH_TRACER
    _PATH_TRACER = cls
    return cls


class RednerRenderer:
    """Name kept so that ``from renderers import LocalRenderer, RednerRenderer`` (main.py:12) works against this
    module.  The Redner path tracer (renderers.py:175-270, external ``pyredner``) is outside this library's scope:
    constructing it returns the class given to :func:`register_path_tracer`, or raises with a pointer to the
    reference's module.  Any such renderer plugs into ``RenderingLoss`` through the generic per-scene loop."""

    def __new__(cls, *args, **kwargs):
        if _PATH_TRACER is None:
            raise NotImplementedError(
                "RednerRenderer is not part of svbrdf_estimation_b200 (only the in-network renderer is): import the "
                "reference's development/multiImage_pytorch/renderers.py (needs pyredner) and either use its "
                "RednerRenderer directly or pass it to svbrdf_estimation_b200.renderers.register_path_tracer()")
        return _PATH_TRACER(*args, **kwargs)


__all__ = ["LocalRenderer", "RednerRenderer", "register_path_tracer", "render_records", "coordinate_table"]
