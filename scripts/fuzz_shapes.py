#!/usr/bin/env python
"""Randomised shape sweep of the C-ABI entry points against the oracle (fp64, CPU): widths 1..70 (odd and even), batch 1..4,
1..30 records, bench and stress maps, grey and coloured lights.   python scripts/fuzz_shapes.py [cases] [seed] [--emu]
--emu runs the host emulation of the kernel algebra (tests/emulation) instead of the GPU library: no GPU needed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
EMU = "--emu" in sys.argv
if EMU:
    sys.argv.remove("--emu")
    from tests.emulation import host as emu
else:
    import svbrdf_estimation_b200 as S
from oracle import reference_port as O
from tests import parity
from tests.common import synthetic_maps

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = {"loss": 0.0, "grad": 0.0, "render": 0.0, "rgrad": 0.0}
for i in range(cases):
    W = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 48, 63, 64, 64, 65, 70, 128]))
    B, nr, ns = int(rng.integers(1, 5)), int(rng.integers(0, 12)), int(rng.integers(0, 19))
    if nr + ns == 0:
        nr = 1
    stress = bool(rng.integers(0, 2))
    inp, tgt = synthetic_maps(B, W, 100 + i, stress), synthetic_maps(B, W, 200 + i, stress)
    torch.manual_seed(i)
    cfg = O.sample_loss_configs(B, nr, ns)
    if rng.integers(0, 3) == 0:                                   # coloured lights -> the general colour path
        cfg[..., 6:9] *= torch.tensor([1.0, 0.7, 1.3])
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), cfg)
    _, g32 = O.rendering_loss_and_grad(inp, tgt, cfg)              # the reference's own fp32 run: the noise floor of the case
    r64i, r64t = O.render_batch(inp.double(), cfg).numpy(), O.render_batch(tgt.double(), cfg).numpy()
    if EMU:
        loss, g = emu.loss_forward_backward(inp.numpy(), tgt.numpy(), cfg.numpy())
    else:
        x = inp.cuda().requires_grad_(True)
        loss = S.rendering_loss_with_records(x, tgt.cuda(), cfg)
        loss.backward()
        g = x.grad.cpu().numpy()
    keep = parity.unambiguous_pixels(r64i, r64t) & ~parity.clamp_ambiguous_pixels(inp, cfg)
    e_loss = abs(float(loss) - float(l64)) / abs(float(l64))
    assert np.isfinite(g).all()
    e_grad = parity.rel_l2(g * keep, g64.numpy() * keep)
    floor = parity.rel_l2(g32.numpy() * keep, g64.numpy() * keep)
    # Images of a few thousand pixels have little averaging: one highlight pixel whose GGX denominator sits near its clamp
    # (q ~ 1e-3, roughness ~ 0.14) can carry 99 % of the squared error of the normals gradient (measured on the host
    # emulation over 120 random bench-distribution cases: median 5.3e-5, p90 8.4e-5, 4 % above 1e-4, max 2.6e-4; the
    # reference's own fp32 run: 3.4e-5 / 5.2e-5 / 1 % / 1.3e-4; accurate kernels: 1.2e-6 / 2.1e-6 / 0 / 5.3e-6).  The 1e-4
    # bound is asserted at the BASELINE sizes (tests/test_gpu_parity.py); here: REL_L2_STRESS or twice the reference's noise.
    assert e_grad <= max(parity.REL_L2_STRESS, 2.0 * floor), ("gradient", W, B, nr + ns, stress, e_grad, floor)
    r = emu.render_forward(inp.numpy(), cfg.numpy()) if EMU else S.render_records(inp.cuda(), cfg).cpu().numpy()
    e_r = parity.rel_l2(r, r64i)
    w = torch.randn(B, nr + ns, 3, W, W)
    m64 = inp.double().requires_grad_(True)
    (O.render_batch(m64, cfg) * w.double()).sum().backward()
    if EMU:
        rg = emu.render_backward(inp.numpy(), cfg.numpy(), w.numpy())
    else:
        x2 = inp.cuda().requires_grad_(True)
        (S.render_records(x2, cfg) * w.cuda()).sum().backward()
        rg = x2.grad.cpu().numpy()
    keep_r = ~parity.clamp_ambiguous_pixels(inp, cfg)
    e_rg = parity.rel_l2(rg * keep_r, m64.grad.numpy() * keep_r)
    # MixedLoss (losses.py:54-63) on the maps and on the network's 9-channel encoding (models.py:334-346)
    e_mix = e_enc = 0.0
    if i % 2 == 0:
        a64 = inp.double().requires_grad_(True)
        m64l = O.mixed_loss(a64, tgt.double(), cfg, 0.1)
        m64l.backward()
        a32 = inp.clone().requires_grad_(True)
        O.mixed_loss(a32, tgt, cfg, 0.1).backward()
        if EMU:
            (mtot, _, _), gm = emu.mixed_loss(inp.numpy(), tgt.numpy(), cfg.numpy(), 0.1)
        else:
            xm = inp.cuda().requires_grad_(True)
            from svbrdf_estimation_b200.losses import _fused_loss
            ml = _fused_loss(xm, tgt.cuda(), cfg, 0.1)
            ml.backward()
            mtot, gm = float(ml), xm.grad.cpu().numpy()
        assert abs(mtot - float(m64l)) <= 5e-6 * abs(float(m64l)), ("mixed loss", W, B, mtot, float(m64l))
        e_mix = parity.rel_l2(gm * keep, a64.grad.numpy() * keep)
        assert e_mix <= max(3e-4, 2.0 * parity.rel_l2(a32.grad.numpy() * keep, a64.grad.numpy() * keep)), ("mixed gradient", W, B, e_mix)
        enc = torch.rand(B, 9, W, W, generator=torch.Generator().manual_seed(300 + i)) * 2 - 1
        e64 = enc.double().requires_grad_(True)
        dec64 = O.decode_network_output(e64)
        l64e = O.mixed_loss(dec64, tgt.double(), cfg, 0.1)
        l64e.backward()
        e32 = enc.clone().requires_grad_(True)
        O.mixed_loss(O.decode_network_output(e32), tgt, cfg, 0.1).backward()
        keep_e = parity.unambiguous_pixels(O.render_batch(dec64.detach(), cfg).numpy(), r64t) & ~parity.clamp_ambiguous_pixels(dec64.detach(), cfg)
        if EMU:
            (etot, _, _), ge = emu.mixed_loss(enc.numpy(), tgt.numpy(), cfg.numpy(), 0.1, encoded=True)
        else:
            xe = enc.cuda().requires_grad_(True)
            le = S.mixed_loss_from_encoded(xe, tgt.cuda(), cfg, 0.1)[0]
            le.backward()
            etot, ge = float(le), xe.grad.cpu().numpy()
        assert abs(etot - float(l64e)) <= 5e-6 * abs(float(l64e)), ("encoded mixed loss", W, B, etot, float(l64e))
        e_enc = parity.rel_l2(ge * keep_e, e64.grad.numpy() * keep_e)
        assert e_enc <= max(3e-4, 2.0 * parity.rel_l2(e32.grad.numpy() * keep_e, e64.grad.numpy() * keep_e)), ("encoded gradient", W, B, e_enc)
    tag = "W%-3d B%d N%-2d %s" % (W, B, nr + ns, "stress" if stress else "bench ")
    print("%s loss %.1e grad %.1e render %.1e render-grad %.1e mixed-grad %.1e encoded-grad %.1e" % (tag, e_loss, e_grad, e_r, e_rg, e_mix, e_enc), flush=True)
    for k, v in (("loss", e_loss), ("grad", e_grad), ("render", e_r), ("rgrad", e_rg)):
        worst[k] = max(worst[k], v)
print("worst:", worst)
# single-pixel images (W = 1) have no averaging: one ill-conditioned stress pixel is the whole statistic
assert worst["loss"] <= 5e-6 and worst["render"] <= 3e-4 and worst["rgrad"] <= 2e-3, worst       # gradients: checked per case above
