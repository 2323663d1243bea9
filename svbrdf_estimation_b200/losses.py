"""Rendering loss and its callers behind the reference's ``nn.Module`` interface.

Drop-in for ``losses.py`` of the reference (development/multiImage_pytorch/losses.py:7-63).
With the B200 ``LocalRenderer`` plugged in, ``RenderingLoss`` runs ONE fused kernel that shades
input and target maps under all sampled light/view configurations, reduces the log-L1 difference
and writes d loss / d input in the same pass; ``MixedLoss`` folds the map-space L1 terms into that
pass as well.  Any other renderer object (e.g. a path-tracer wrapper with the same
``render(scene, svbrdf)`` method) goes through the generic per-scene loop of the plugin API.
"""
import torch
import torch.nn as nn

from . import _cabi
from . import environment as env
from .renderers import as_device_maps, as_host_records, coordinate_table
from .utils import unpack_svbrdf

EPSILON_RENDER = 0.1   # losses.py:46
EPSILON_L1 = 0.01      # losses.py:13


def _workspace(B, N, H, W, device):
    nbytes = _cabi.lib().svbrdf_b200_workspace_bytes(B, N, H, W)
    return torch.empty((nbytes + 3) // 4, device=device, dtype=torch.float32), nbytes


def _check_pair(input, target):
    if input.shape != target.shape:
        raise ValueError("input and target shapes differ: %s vs %s" % (tuple(input.shape), tuple(target.shape)))
    if input.dim() != 4:
        raise ValueError("expected [B,12,H,W] tensors, got %s" % (tuple(input.shape),))
    if input.device != target.device:
        raise ValueError("input and target are on different devices")


class _NoSwitch:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def _on_device(device):
    """The C ABI works on the CURRENT CUDA device: switch only when the tensors live elsewhere (the context manager
    costs more host time than the launch it guards)."""
    return _NO_SWITCH if device.index == torch.cuda.current_device() else torch.cuda.device(device)


def _launch_loss(a, b, records, l1_weight, out, grad, ws, ws_bytes, lin, encoded=False, accurate=False):
    """One C-ABI loss call on the current stream of ``b``'s device.  ``a`` is the differentiated argument
    ([B,12,H,W] maps, or the [B,9,H,W] encoded network output when ``encoded``), ``grad`` its gradient buffer
    (or None: forward only)."""
    lib = _cabi.lib()
    B, _, H, W = b.shape
    N = records.shape[1]
    gp = grad.data_ptr() if grad is not None else None
    with _on_device(b.device):
        stream = torch.cuda.current_stream().cuda_stream
        if encoded:
            status = lib.svbrdf_b200_mixed_loss_encoded_forward_backward(
                a.data_ptr(), b.data_ptr(), B, H, W, records.data_ptr(), N, float(l1_weight), lin.data_ptr(),
                out.data_ptr(), gp, ws.data_ptr(), ws_bytes, stream)
        elif l1_weight is not None:
            status = lib.svbrdf_b200_mixed_loss_forward_backward(
                a.data_ptr(), b.data_ptr(), B, H, W, records.data_ptr(), N, float(l1_weight), lin.data_ptr(),
                out.data_ptr(), gp, ws.data_ptr(), ws_bytes, stream)
        elif accurate:
            fn = lib.svbrdf_b200_loss_forward_backward_accurate if grad is not None else None
            if fn is not None:
                status = fn(a.data_ptr(), b.data_ptr(), B, H, W, records.data_ptr(), N, lin.data_ptr(), out.data_ptr(),
                            gp, ws.data_ptr(), ws_bytes, stream)
            else:
                status = lib.svbrdf_b200_loss_forward_accurate(
                    a.data_ptr(), b.data_ptr(), B, H, W, records.data_ptr(), N, lin.data_ptr(), out.data_ptr(),
                    ws.data_ptr(), ws_bytes, stream)
        else:
            # RenderingLoss, forward (+ backward when grad is given); writes out[0..2] = loss, loss, 0
            status = lib.svbrdf_b200_loss_layouts(
                a.data_ptr(), _cabi.LAYOUT_MAPS12, b.data_ptr(), _cabi.LAYOUT_MAPS12, B, H, W, records.data_ptr(), N, -1.0,
                lin.data_ptr(), out.data_ptr(), gp, ws.data_ptr(), ws_bytes, stream)
    _cabi.check(status)


class _FusedLoss(torch.autograd.Function):
    """Loss value and d loss / d input for given host scene records.

    ``l1_weight is None`` -> RenderingLoss, otherwise MixedLoss with that weight; ``encoded`` -> ``input`` is
    the 9-channel network output.  ``eager``: the gradient is computed by the same kernel pass as the loss value
    (training: one launch for forward + backward) and handed to autograd on the first ``backward()``.  Without
    ``eager`` (validation) the forward-only kernel runs and nothing is allocated for a gradient; if ``backward()`` is
    called after all - or a second time with ``retain_graph=True`` - the gradient is recomputed then.
    Returns ``(loss, parts)``: ``loss`` is the differentiable 0-dim value, ``parts = [rendering loss, map-L1
    loss]`` is informational and marked non-differentiable."""

    @staticmethod
    def forward(ctx, input, target, records, l1_weight, encoded, accurate=False, eager=True):
        B, _, H, W = target.shape
        dev = target.device
        if accurate and (l1_weight is not None or encoded):
            raise NotImplementedError("accurate=True is available for RenderingLoss (not MixedLoss / encoded input)")
        out = torch.empty(3, device=dev, dtype=torch.float32)
        ws, ws_bytes = _workspace(B, records.shape[1], H, W, dev)
        lin = coordinate_table(W, dev)
        want_in, want_tg = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if want_tg and encoded:
            raise NotImplementedError("the gradient w.r.t. the target is not available for encoded input")
        # the encoded-input kernels only exist in the forward+backward form
        grad_in = torch.empty_like(input) if ((want_in and eager) or encoded) else None
        _launch_loss(input, target, records, l1_weight, out, grad_in, ws, ws_bytes, lin, encoded, accurate)
        grad_tg = None
        if want_tg and eager:
            # the loss is symmetric in its arguments: d/d target = the same kernel with the roles swapped
            grad_tg = torch.empty_like(target)
            _launch_loss(target, input, records, l1_weight, torch.empty_like(out), grad_tg, ws, ws_bytes, lin, False, accurate)
        ctx.grads = (grad_in if want_in else None, grad_tg) if eager else None
        ctx.save_for_backward(input, target)
        ctx.records, ctx.meta = records, (l1_weight, encoded, accurate)
        loss, parts = out[0].reshape(()), out[1:3]        # the finalize kernel writes all three values (map-L1 = 0 for RenderingLoss)
        if accurate:                                      # the accurate entry points write the loss only
            parts = torch.stack((loss.detach(), torch.zeros_like(loss)))
        ctx.mark_non_differentiable(parts)
        return loss, parts

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, _grad_parts):
        want_in, want_tg = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grads, ctx.grads = ctx.grads, None
        if grads is None:
            # validation-mode forward, or a repeated backward (retain_graph=True): evaluate the gradient now
            input, target = ctx.saved_tensors
            l1_weight, encoded, accurate = ctx.meta
            B, _, H, W = target.shape
            ws, ws_bytes = _workspace(B, ctx.records.shape[1], H, W, target.device)
            lin = coordinate_table(W, target.device)
            scratch = torch.empty(3, device=target.device, dtype=torch.float32)
            g_in = g_tg = None
            if want_in:
                g_in = torch.empty_like(input)
                _launch_loss(input, target, ctx.records, l1_weight, scratch, g_in, ws, ws_bytes, lin, encoded, accurate)
            if want_tg:
                g_tg = torch.empty_like(target)
                _launch_loss(target, input, ctx.records, l1_weight, scratch, g_tg, ws, ws_bytes, lin, False, accurate)
            grads = (g_in, g_tg)
        up = grad_loss.reshape(1).to(torch.float32).contiguous()
        lib = _cabi.lib()
        for g in grads:
            if g is not None:
                with _on_device(g.device):
                    # in place; returns on the device when the upstream gradient is 1 (no host sync)
                    _cabi.check(lib.svbrdf_b200_scale_grad(g.data_ptr(), g.numel(), up.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream))
        return grads[0], grads[1], None, None, None, None, None


def mixed_loss_from_encoded(encoded, target, records, l1_weight=0.1, eager=True):
    """``MixedLoss`` evaluated directly on the network's encoded output ``[B,9,H,W]`` (after tanh: normal xy,
    diffuse, roughness, specular in [-1,1]); the decode of models.py:334-346 / utils.py:73-98 and its chain
    rule run inside the loss kernel.  Returns ``(mixed, rendering, map_l1)`` 0-dim tensors;
    only the first is differentiable."""
    if encoded.dim() != 4 or encoded.shape[1] != 9:
        raise ValueError("encoded must be [B,9,H,W], got %s" % (tuple(encoded.shape),))
    if encoded.dtype != torch.float32:
        raise TypeError("encoded must be float32, got %s" % encoded.dtype)
    b, _, origin = as_device_maps(target, "target")
    if tuple(encoded.shape[0:1] + encoded.shape[2:]) != tuple(b.shape[0:1] + b.shape[2:]):
        raise ValueError("encoded %s and target %s do not match" % (tuple(encoded.shape), tuple(target.shape)))
    if encoded.is_cuda and encoded.device != b.device:
        raise ValueError("encoded is on %s, target on %s" % (encoded.device, b.device))
    e = encoded.to(b.device).contiguous()
    rec = as_host_records(records, b.shape[0])
    if rec.dim() != 3:
        raise ValueError("the loss needs per-batch-element scene records [B,N,9]")
    loss, parts = _FusedLoss.apply(e, b, rec, float(l1_weight), True, False, bool(eager))
    return tuple(origin.restore(t) for t in (loss, parts[0], parts[1]))


def _fused_loss(input, target, records, l1_weight, accurate=False, eager=True):
    """-> differentiable 0-dim loss on the input's device (RenderingLoss, or MixedLoss when l1_weight is given)."""
    _check_pair(input, target)
    a, _, origin = as_device_maps(input, "input")
    b, _, _ = as_device_maps(target, "target")
    if a.device != b.device:
        raise ValueError("input and target are on different devices")
    rec = as_host_records(records, a.shape[0])
    if rec.dim() != 3:
        raise ValueError("the loss needs per-batch-element scene records [B,N,9]")
    loss, _ = _FusedLoss.apply(a, b, rec, l1_weight, False, bool(accurate), bool(eager))
    return origin.restore(loss)


def rendering_loss_with_records(input, target, records, accurate=False, eager=True):
    """``RenderingLoss`` for explicit scene records ``[B,N,9]`` (no sampling) - the notebook-style
    fixed-scene loss and the form the parity tests use.  ``accurate=True`` selects the accurate-highlight
    kernels (loss and gradient 20-30x closer to an fp64 evaluation than the reference's own fp32 run,
    about 12 % slower)."""
    return _fused_loss(input, target, records, None, accurate, eager)


class SVBRDFL1Loss(nn.Module):
    """Sum of four mean-L1 terms over the map groups, diffuse and specular compared as
    ``log(x + 0.01)`` (losses.py:7-19).  Stand-alone (unfused) form, plain torch ops on the
    tensors' device; ``MixedLoss`` fuses it into the rendering-loss kernel."""

    def forward(self, input, target):
        n0, d0, r0, s0 = unpack_svbrdf(input)
        n1, d1, r1, s1 = unpack_svbrdf(target)
        l1 = nn.functional.l1_loss
        return (l1(n0, n1) + l1(torch.log(d0 + EPSILON_L1), torch.log(d1 + EPSILON_L1))
                + l1(r0, r1) + l1(torch.log(s0 + EPSILON_L1), torch.log(s1 + EPSILON_L1)))


class RenderingLoss(nn.Module):
    """mean | log(render(input)+0.1) - log(render(target)+0.1) | under freshly sampled light/view
    configurations per batch element (losses.py:21-52).  No parameters, no buffers."""

    def __init__(self, renderer, scene_sampler=None, accurate=False):
        super().__init__()
        self.renderer = renderer
        self.accurate = accurate                # accurate-highlight kernels (see rendering_loss_with_records)
        self.random_configuration_count = 3     # losses.py:26
        self.specular_configuration_count = 6   # losses.py:27
        # None: the reference's sampler (global CPU generator, reference draw order).  Otherwise a
        # callable (batch, n_random, n_specular) -> [B,N,9], e.g. environment.NativeSceneSampler.
        self.scene_sampler = scene_sampler

    def sample_records(self, batch_size):
        """[B,N,9] scene records of one evaluation (fresh scenes per batch element, losses.py:35)."""
        sampler = self.scene_sampler or env.sample_loss_configs
        return sampler(batch_size, self.random_configuration_count, self.specular_configuration_count)

    def forward(self, input, target):
        """In training mode (the default of every ``nn.Module``) the gradient w.r.t. ``input`` comes out of the same
        kernel pass as the value.  After ``.eval()`` only the value is computed (validation, main.py:140); a
        ``backward()`` still works and evaluates the gradient then."""
        if getattr(self.renderer, "fused_rendering_loss", False):
            _check_pair(input, target)
            return _fused_loss(input, target, self.sample_records(input.shape[0]), None, self.accurate, self.training)
        return self._forward_with_plugin(input, target)

    def _forward_with_plugin(self, input, target):
        """Generic renderer plugin: one ``render`` call per (sample, scene, map) like losses.py:34-50."""
        ins, tgs = [], []
        for i in range(input.shape[0]):
            scenes = (env.generate_random_scenes(self.random_configuration_count)
                      + env.generate_specular_scenes(self.specular_configuration_count))
            ins.append(torch.cat([self.renderer.render(s, input[i]) for s in scenes], dim=0))
            tgs.append(torch.cat([self.renderer.render(s, target[i]) for s in scenes], dim=0))
        a = torch.log(torch.stack(ins, dim=0) + EPSILON_RENDER)
        b = torch.log(torch.stack(tgs, dim=0) + EPSILON_RENDER)
        return nn.functional.l1_loss(a, b)


class MixedLoss(nn.Module):
    """``l1_weight * SVBRDFL1Loss + RenderingLoss`` (losses.py:54-63) - what the training script
    builds (main.py:89).  With the fused renderer both terms and their gradient come from one
    kernel pass over the maps."""

    def __init__(self, renderer, l1_weight=0.1, scene_sampler=None):
        super().__init__()
        self.l1_weight = l1_weight
        self.l1_loss = SVBRDFL1Loss()
        self.rendering_loss = RenderingLoss(renderer, scene_sampler)

    def forward(self, input, target):
        rl = self.rendering_loss
        if getattr(rl.renderer, "fused_rendering_loss", False):
            _check_pair(input, target)
            return _fused_loss(input, target, rl.sample_records(input.shape[0]), float(self.l1_weight), False, self.training)
        return self.l1_weight * self.l1_loss(input, target) + rl(input, target)

    def forward_encoded(self, encoded, target):
        """Same loss on the generator's 9-channel output after tanh (what ``SingleViewModel.forward`` feeds
        to ``utils.decode_svbrdf``, models.py:334-338): skips materialising the 12-channel prediction."""
        rl = self.rendering_loss
        if not getattr(rl.renderer, "fused_rendering_loss", False):
            from .utils import decode_network_output
            return self.forward(decode_network_output(encoded), target)
        return mixed_loss_from_encoded(encoded, target, rl.sample_records(target.shape[0]), float(self.l1_weight), self.training)[0]


__all__ = ["SVBRDFL1Loss", "RenderingLoss", "MixedLoss", "rendering_loss_with_records", "mixed_loss_from_encoded"]
