"""CPU check of the kernels' algebra: csrc/shading.cuh compiled for the host (tests/emulation)
against the golden fixtures of the unmodified reference, with the same three-way tolerances the
GPU parity tests use.  Catches mistakes in the rearranged forward math and in the analytic adjoint
without a GPU; the MUFU approximations themselves are only exercised by the -m gpu tests."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests import parity
from tests.emulation import host as emu


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("fixture", ["loss_bench", "loss_stress", "loss_n27", "loss_real"])
def test_loss_and_gradient_three_way(golden, fixture, lanes):
    g = golden(fixture)
    loss, grad = emu.loss_forward_backward(g["input"], g["target"], g["configs"], lanes)
    parity.check_loss(loss, g["loss_f64"])
    for row in parity.check_grad_groups(grad, g["grad_f32"], g["grad_f64"], rel=parity.REL_L2_STRESS if "stress" in fixture else parity.REL_L2):
        print(fixture, row)


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("fixture", ["loss_bench", "loss_stress"])
def test_renders_three_way(golden, fixture, lanes):
    g = golden(fixture)
    got = emu.render_forward(g["input"], g["configs"], lanes)
    print(parity.check_tensor(got, g["renders_f32"], g["renders_f64"], "renders"))
    # the quantity the loss consumes
    dlog = np.abs(np.log(got.astype(np.float64) + 0.1) - np.log(g["renders_f64"] + 0.1)).max()
    assert dlog < 2e-3, dlog
    if fixture == "loss_bench":      # accurate-highlight forward (shading.cuh, ACC): closer to fp64 than the reference's fp32 run
        e64, floor = parity.rel_l2(got, g["renders_f64"]), parity.rel_l2(g["renders_f32"], g["renders_f64"])
        assert e64 <= 0.25 * floor and e64 <= 2e-5, (e64, floor)


def test_render_fixed_scenes(golden):
    g = golden("render_fixed")
    got = emu.render_forward(g["maps"], g["configs"])            # [B,2,3,H,W], one shared scene list
    ref32 = np.moveaxis(g["render4d_f32"], 0, 1)
    ref64 = np.moveaxis(g["render4d_f64"], 0, 1)
    parity.check_tensor(got, ref32, ref64, "render_fixed")


@pytest.mark.parametrize("lanes", [1, 2])
def test_render_backward_matches_oracle_autograd(golden, lanes):
    g = golden("loss_stress")
    maps = torch.from_numpy(g["input"]).double().requires_grad_(True)
    cfg = torch.from_numpy(g["configs"])
    gen = torch.Generator().manual_seed(5)
    w = torch.randn(maps.shape[0], cfg.shape[1], 3, maps.shape[2], maps.shape[3], generator=gen)
    (O.render_batch(maps, cfg) * w.double()).sum().backward()
    maps32 = torch.from_numpy(g["input"]).requires_grad_(True)
    (O.render_batch(maps32, cfg) * w).sum().backward()
    got = emu.render_backward(g["input"], g["configs"], w.numpy(), lanes)
    parity.check_grad_groups(got, maps32.grad.numpy(), maps.grad.numpy(), "render_bwd", rel=parity.REL_L2_RAW_RENDER_GRAD)


@pytest.mark.parametrize("fixture", ["loss_bench", "loss_stress"])
def test_identical_maps_give_zero_loss_and_gradient(golden, fixture):
    g = golden(fixture)
    for lanes in (1, 2):
        loss, grad = emu.loss_forward_backward(g["input"], g["input"], g["configs"], lanes)
        assert loss == 0.0 and not grad.any()


def test_scalar_and_packed_paths_agree_bitwise(golden):
    """The F2 lane type performs the same fp32 operations per lane as the scalar path."""
    g = golden("loss_stress")
    l1, g1 = emu.loss_forward_backward(g["input"], g["target"], g["configs"], 1)
    l2, g2 = emu.loss_forward_backward(g["input"], g["target"], g["configs"], 2)
    np.testing.assert_array_equal(g1, g2)
    assert abs(l1 - l2) <= 1e-8 * abs(l1)   # lane sums are added in a different order


@pytest.mark.parametrize("lanes", [1, 2])
def test_partially_identical_maps_are_exact(golden, lanes):
    """Only diffuse channel 0 differs: the reference renders channels 1 and 2 identically for input and
    target, so they contribute exactly 0 (losses.py:50 sign(0) = 0); the kernels mask those channels."""
    g = golden("loss_bench")
    tgt = g["target"].copy()
    inp = tgt.copy()
    inp[:, 3] = g["input"][:, 3]
    cfg = torch.from_numpy(g["configs"])
    l64, g64 = O.rendering_loss_and_grad(torch.from_numpy(inp).double(), torch.from_numpy(tgt).double(), cfg)
    loss, grad = emu.loss_forward_backward(inp, tgt, g["configs"], lanes)
    parity.check_loss(loss, float(l64))
    for ch in (4, 5, 7, 8, 10, 11):
        assert not grad[:, ch].any() and not g64.numpy()[:, ch].any(), ch
    assert parity.rel_l2(grad, g64.numpy()) <= 1e-4


@pytest.mark.parametrize("lanes", [1, 2])
def test_mixed_loss_matches_reference(golden, lanes):
    g, gm = golden("loss_bench"), golden("mixed")
    torch.manual_seed(int(gm["seed"]))
    cfg = O.sample_loss_configs(g["input"].shape[0])
    (total, _, l1), grad = emu.mixed_loss(g["input"], g["target"], cfg.numpy(), 0.1, False, lanes)
    assert abs(total - float(gm["loss_f32"])) <= 3e-6 * float(gm["loss_f32"])
    assert abs(l1 - float(gm["l1_f32"])) <= 2e-6 * float(gm["l1_f32"])
    for name, s in parity.GROUPS:
        assert parity.rel_l2(grad[:, s], gm["grad_f32"][:, s]) <= 1.5e-4, name


@pytest.mark.parametrize("lanes", [1, 2])
def test_encoded_input_loss_and_gradient(golden, lanes):
    """Model-output epilogue fused into the loss: decode (utils.py:73-98, models.py:340-346) and its chain rule."""
    gd = golden("decode")
    enc = torch.from_numpy(gd["encoded"])                      # [2,9,12,12]
    tgt = torch.from_numpy(golden("loss_bench")["target"][:, :, :12, :12].copy())
    torch.manual_seed(4)
    cfg = O.sample_loss_configs(2)
    e64 = enc.double().requires_grad_(True)
    want = O.mixed_loss(O.decode_network_output(e64), tgt.double(), cfg, 0.1)
    want.backward()
    (total, _, _), grad = emu.mixed_loss(enc.numpy(), tgt.numpy(), cfg.numpy(), 0.1, True, lanes)
    assert abs(total - float(want)) <= 3e-6 * float(want)
    assert grad.shape == (2, 9, 12, 12)
    g64 = e64.grad.numpy()
    for name, s in (("normal_xy", slice(0, 2)), ("diffuse", slice(2, 5)), ("roughness", slice(5, 6)), ("specular", slice(6, 9))):
        assert parity.rel_l2(grad[:, s], g64[:, s]) <= 1.5e-4, (name, parity.rel_l2(grad[:, s], g64[:, s]))


def test_degenerate_maps_stay_finite_and_match(golden):
    """Zero normals, zero / one albedos, zero and >1 roughness, light and camera straight above a pixel."""
    g = golden("loss_bench")
    tgt = g["target"][:, :, :8, :8].copy()
    inp = np.zeros_like(tgt)
    inp[0, 0:3] = 0.0                                   # null normal: every dot product clamps
    inp[0, 3:6], inp[0, 6:9], inp[0, 9:12] = 0.0, 0.0, 1.0
    inp[1, 0:3] = np.array([0.0, 0.0, 1.0], dtype=np.float32).reshape(3, 1, 1)
    inp[1, 3:6], inp[1, 6:9], inp[1, 9:12] = 1.0, 1.5, 0.0   # roughness above 1
    cfg = g["configs"][:, :4].copy()
    cfg[0, 0] = [0.0, 0.0, 1.0, 0.0, 0.0, 1.0, 20.0, 20.0, 20.0]          # camera == light, above the centre
    cfg[1, 1] = [-1.0, 1.0, 0.5, -1.0, 1.0, 0.5, 50.0, 50.0, 50.0]        # both above the top-left corner pixel
    tc = torch.from_numpy(cfg)
    l64, g64 = O.rendering_loss_and_grad(torch.from_numpy(inp).double(), torch.from_numpy(tgt).double(), tc)
    assert torch.isfinite(l64) and torch.isfinite(g64).all()
    for lanes in (1, 2):
        loss, grad = emu.loss_forward_backward(inp, tgt, cfg, lanes)
        assert np.isfinite(loss) and np.isfinite(grad).all()
        assert abs(loss - float(l64)) <= 5e-6 * float(l64)
        assert parity.rel_l2(grad, g64.numpy()) <= 3e-4, parity.rel_l2(grad, g64.numpy())


def test_random_shapes_against_the_oracle():
    """scripts/fuzz_shapes.py on the host emulation: widths 1..70 (odd / even -> both lane types), 1..30 records, bench and
    stress maps, grey and coloured lights; loss, gradient (over pixels without a sign- or clamp-ambiguous term, against the
    reference's own fp32 noise), renders and render gradients vs the fp64 oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_shapes.py"), "12", "7", "--emu"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]


@pytest.mark.parametrize("lanes", [1, 2])
def test_accurate_loss_is_closer_to_fp64_than_the_reference_fp32_run(golden, lanes):
    """svbrdf_b200_loss_forward_backward_accurate (host twin): on unit-normal inputs every gradient group is several times
    closer to the fp64 reference than the reference's own fp32 run."""
    g = golden("loss_bench")
    loss, grad = emu.loss_forward_backward(g["input"], g["target"], g["configs"], lanes, accurate=True)
    parity.check_loss(loss, g["loss_f64"])
    for name, s in parity.GROUPS:
        e64 = parity.rel_l2(grad[:, s], g["grad_f64"][:, s])
        floor = parity.rel_l2(g["grad_f32"][:, s], g["grad_f64"][:, s])
        assert e64 <= max(0.3 * floor, 3e-7), (name, e64, floor)


@pytest.mark.parametrize("lanes", [1, 2])
def test_maps_that_differ_bitwise_but_render_identically_are_exact_zeros(lanes):
    """Two roughness values below the clamp of renderers.py:87 are the same material, and the diffuse albedo does not
    matter where the specular albedo is exactly 1 ((1-F) d = 0, renderers.py:18-20,32): the reference renders such pairs
    bit-identically and their terms contribute exactly 0 (sign(0) = 0, losses.py:50).  The kernels decide identity on the
    parameters the shading consumes, so they do too - instead of handing a full-magnitude +-1/x gradient to rounding noise."""
    from tests.common import synthetic_maps
    tgt = synthetic_maps(2, 16, 3)
    inp = tgt.clone()
    inp[0, 6:9] = 2e-4                      # element 0: roughness below the clamp on both maps, different bits
    tgt[0, 6:9] = 7e-4
    inp[1, 9:12] = 1.0                      # element 1: specular albedo exactly 1 on both maps, different diffuse albedo
    tgt[1, 9:12] = 1.0
    inp[1, 3:6] = torch.rand(3, 16, 16)
    torch.manual_seed(2)
    cfg = O.sample_loss_configs(2)
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), cfg)
    assert float(l64) == 0.0 and not bool(g64.any())                     # the reference: exactly zero
    loss, grad = emu.loss_forward_backward(inp.numpy(), tgt.numpy(), cfg.numpy(), lanes)
    assert loss == 0.0 and not grad.any()
    # a mixed case: only colour channel 1 of element 1 really differs (its specular albedo moves off 1)
    inp[1, 10] = 0.5
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), cfg)
    loss, grad = emu.loss_forward_backward(inp.numpy(), tgt.numpy(), cfg.numpy(), lanes)
    parity.check_loss(loss, float(l64))
    assert not grad[0].any() and not grad[1, [3, 5, 6, 8, 9, 11]].any()
    assert parity.rel_l2(grad, g64.numpy()) <= 1e-4


def test_flipped_term_analysis_recovers_planted_flips():
    """tests/parity.py flipped_term_analysis: plant sign flips on a few sign-ambiguous L1 terms of the fp64 gradient (and on
    no other term) and check that exactly those are found and put back."""
    from tests.common import synthetic_maps
    inp, tgt = synthetic_maps(2, 48, 21), synthetic_maps(2, 48, 22)
    torch.manual_seed(313)
    cfg = O.sample_loss_configs(2)
    x = inp.double().requires_grad_(True)
    l = torch.log(O.render_batch(x, cfg) + 0.1) - torch.log(O.render_batch(tgt.double(), cfg) + 0.1)
    amb = ((l != 0) & (l.abs() < parity.AMBIGUOUS_DLOG)).detach()
    idx = amb.nonzero()
    assert len(idx) >= 6
    planted = idx[:: max(1, len(idx) // 5)][:5]
    sign = torch.sign(l.detach())
    for i in planted:
        sign[tuple(i)] *= -1.0
    (sign * l).mean().backward()                                        # gradient of the loss with those terms flipped
    res = parity.flipped_term_analysis(O, inp, tgt, cfg, x.grad.numpy())
    assert res["flipped"] == len(planted) and res["candidates"] == int(amb.sum())
    assert 0 < res["max_abs_l_flipped"] < parity.AMBIGUOUS_DLOG
    out = parity.check_grad_with_flipped_terms(res, "planted", rel=1e-10)
    assert max(v["raw"] for v in out.values()) > 1e-6                   # the flips were visible before the correction
    # the unmodified fp64 gradient has no flipped term
    res0 = parity.flipped_term_analysis(O, inp, tgt, cfg, res["g64"])
    assert res0["flipped"] == 0
