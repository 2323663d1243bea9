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

for size, n_rand, n_spec, stress in ((24, 3, 6, False), (15, 1, 1, True), (34, 2, 3, True), (64, 3, 6, False)):   # 64: per-warp row table
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
        S.rendering_loss_with_records(a, b, rec, accurate=True)
    x = a.detach().clone().requires_grad_(True)
    S.rendering_loss_with_records(x, b, rec, accurate=True).backward()
    # 10-channel layout through the C ABI
    from svbrdf_estimation_b200 import _cabi
    lib = _cabi.lib()
    to10 = lambda m: torch.cat((m[:, 0:7], m[:, 9:12]), dim=1).contiguous()
    a10, b10 = to10(a.detach()), to10(b)
    g10, out = torch.empty_like(a10), torch.zeros(3, device="cuda")
    lin = torch.linspace(-1, 1, size, device="cuda")
    nb = lib.svbrdf_b200_workspace_bytes(B, rec.shape[1], size, size)
    ws = torch.empty(nb // 4 + 1, device="cuda")
    _cabi.check(lib.svbrdf_b200_loss_layouts(a10.data_ptr(), 10, b10.data_ptr(), 10, B, size, size, rec.data_ptr(), rec.shape[1], -1.0,
                                             lin.data_ptr(), out.data_ptr(), g10.data_ptr(), ws.data_ptr(), nb, None))
    torch.cuda.synchronize()
    print("ok", size, float(a.grad.abs().sum()), float(enc.grad.abs().sum()))
