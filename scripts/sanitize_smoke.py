#!/usr/bin/env python
"""Small run of every C-ABI entry point, meant for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import svbrdf_estimation_b200 as S
from svbrdf_estimation_b200 import environment as E
from tests.common import synthetic_maps

for size, n_rand, n_spec, stress in ((24, 3, 6, False), (15, 1, 1, True), (34, 2, 3, True)):
    B = 2
    a = synthetic_maps(B, size, 1, stress=stress).cuda().requires_grad_(True)
    b = synthetic_maps(B, size, 2, stress=stress).cuda()
    rec = E.sample_loss_configs(B, n_rand, n_spec)
    S.rendering_loss_with_records(a, b, rec).backward()
    S.MixedLoss(S.LocalRenderer())(a, b).backward()
    img = S.render_records(a, rec)
    (img * torch.randn_like(img)).sum().backward()
    enc = (torch.rand(B, 9, size, size, device="cuda") * 2 - 1).requires_grad_(True)
    S.mixed_loss_from_encoded(enc, b, rec)[0].backward()
    with torch.no_grad():
        S.rendering_loss_with_records(a, b, rec)
    torch.cuda.synchronize()
    print("ok", size, float(a.grad.abs().sum()), float(enc.grad.abs().sum()))
