"""B200-native rendering-loss path of mworchel/svbrdf-estimation.

Modules mirror the reference's (development/multiImage_pytorch/): ``renderers`` (LocalRenderer),
``losses`` (RenderingLoss, MixedLoss, SVBRDFL1Loss), ``environment`` (Camera/Light/Scene and the
configuration samplers) and ``utils`` (channel layout, direction sampler).  The arithmetic lives in
``libsvbrdf_b200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/svbrdf_b200.h``).
"""
from . import environment, inputs, losses, renderers, utils
from .environment import Camera, Light, Scene, generate_random_scenes, generate_specular_scenes
from .losses import MixedLoss, RenderingLoss, SVBRDFL1Loss, mixed_loss_from_encoded, rendering_loss_with_records
from .renderers import LocalRenderer, render_records

__version__ = "0.1.0"
__all__ = ["environment", "inputs", "losses", "renderers", "utils", "Camera", "Light", "Scene",
           "generate_random_scenes", "generate_specular_scenes", "MixedLoss", "RenderingLoss",
           "SVBRDFL1Loss", "rendering_loss_with_records", "mixed_loss_from_encoded", "LocalRenderer", "render_records"]
