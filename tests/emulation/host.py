"""ctypes wrapper of tests/emulation/emu.cpp (built on demand with g++).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu.cpp")
LIB = os.path.join(HERE, "libemu.so")
HDRS = [os.path.join(HERE, "..", "..", "svbrdf_estimation_b200", "csrc", h) for h in ("shading.cuh", "pixel_ops.cuh")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max([os.path.getmtime(SRC)] + [os.path.getmtime(h) for h in HDRS]):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", LIB, SRC])
        _lib = ctypes.CDLL(LIB)
        _lib.emu_loss_forward_backward.restype = ctypes.c_double
        _lib.emu_loss_forward_backward_accurate.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def lin_table(w):
    import torch
    return torch.linspace(-1, 1, w, dtype=torch.float32).numpy()


def loss_forward_backward(inp, tgt, scenes, lanes=0, accurate=False):
    inp, tgt, scenes = _f32(inp), _f32(tgt), _f32(scenes)
    B, _, H, W = inp.shape
    grad = np.empty_like(inp)
    lin = lin_table(W)
    fn = lib().emu_loss_forward_backward_accurate if accurate else lib().emu_loss_forward_backward
    loss = fn(_p(inp), _p(tgt), B, H, W, _p(scenes), scenes.shape[1], _p(lin), _p(grad), lanes)
    return loss, grad


def render_forward(maps, scenes, lanes=0):
    maps, scenes = _f32(maps), _f32(scenes)
    B, _, H, W = maps.shape
    per_batch = scenes.ndim == 3
    N = scenes.shape[-2]
    out = np.empty((B, N, 3, H, W), dtype=np.float32)
    lin = lin_table(W)
    lib().emu_render_forward(_p(maps), B, H, W, _p(scenes), N, int(per_batch), _p(lin), _p(out), lanes)
    return out


def render_backward(maps, scenes, grad_images, lanes=0):
    maps, scenes, grad_images = _f32(maps), _f32(scenes), _f32(grad_images)
    B, _, H, W = maps.shape
    per_batch = scenes.ndim == 3
    N = scenes.shape[-2]
    out = np.empty_like(maps)
    lin = lin_table(W)
    lib().emu_render_backward(_p(maps), B, H, W, _p(scenes), N, int(per_batch), _p(lin), _p(grad_images), _p(out), lanes)
    return out


def mixed_loss(inp, tgt, scenes, l1_weight=0.1, encoded=False, lanes=0):
    """-> (total, rendering, map_l1), grad (12 or 9 channels)."""
    inp, tgt, scenes = _f32(inp), _f32(tgt), _f32(scenes)
    B, _, H, W = tgt.shape
    grad = np.empty_like(inp)
    lin = lin_table(W)
    out = (ctypes.c_double * 3)()
    lib().emu_loss(_p(inp), _p(tgt), B, H, W, _p(scenes), scenes.shape[1], _p(lin), _p(grad), lanes, 1,
                   ctypes.c_float(l1_weight), int(encoded), out)
    return tuple(out), grad
