"""The oracle (oracle/reference_port.py) against the fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O


def t(a, dtype=None):
    x = torch.from_numpy(np.asarray(a))
    return x if dtype is None else x.to(dtype)


def test_direction_sampler_matches_reference(golden):
    g = golden("directions")
    torch.manual_seed(99)
    np.testing.assert_array_equal(O.cosine_hemisphere_directions(5, 0.001, 0.1).numpy(), g["seed99_count5"])


@pytest.mark.parametrize("key,seed,batch,nr,ns", [("seed313_b4_r3_s6", 313, 4, 3, 6), ("seed7_b2_r9_s18", 7, 2, 9, 18)])
def test_scene_sampler_matches_reference_draw_order(golden, key, seed, batch, nr, ns):
    torch.manual_seed(seed)
    got = O.sample_loss_configs(batch, nr, ns).numpy()
    np.testing.assert_array_equal(got, golden("scenes")[key])


@pytest.mark.parametrize("name,dtype,tol", [("f32", torch.float32, 0.0), ("f64", torch.float64, 1e-12)])
def test_render_fixed_scenes(golden, name, dtype, tol):
    g = golden("render_fixed")
    maps, cfg = t(g["maps"], dtype), t(g["configs"])
    for k in range(cfg.shape[0]):
        got = O.render(cfg[k, 0:3], cfg[k, 3:6], cfg[k, 6:9], maps).numpy()
        np.testing.assert_allclose(got, g["render4d_" + name][k], rtol=tol, atol=0)
    got3 = O.render(cfg[0, 0:3], cfg[0, 3:6], cfg[0, 6:9], maps[0]).numpy()
    assert got3.shape == (1, 3) + maps.shape[-2:]
    np.testing.assert_allclose(got3, g["render3d_" + name], rtol=tol, atol=0)


@pytest.mark.parametrize("fixture", ["loss_bench", "loss_stress", "loss_n27", "loss_real"])
@pytest.mark.parametrize("name,dtype", [("f32", torch.float32), ("f64", torch.float64)])
def test_rendering_loss_and_gradient(golden, fixture, name, dtype):
    g = golden(fixture)
    inp, tgt, cfg = t(g["input"], dtype), t(g["target"], dtype), t(g["configs"])
    loss, grad = O.rendering_loss_and_grad(inp, tgt, cfg)
    if dtype == torch.float32:
        # same ops in the same order on the same ATen kernels: bit-exact
        np.testing.assert_array_equal(loss.numpy(), g["loss_f32"])
        np.testing.assert_array_equal(grad.numpy(), g["grad_f32"])
    else:
        np.testing.assert_allclose(loss.numpy(), g["loss_f64"], rtol=1e-13)
        if g["grad_f64"].dtype == np.float64:
            np.testing.assert_allclose(grad.numpy(), g["grad_f64"], rtol=1e-9, atol=1e-15)
        else:                                   # loss_real stores the fp64 gradient rounded to fp32
            np.testing.assert_allclose(grad.numpy(), g["grad_f64"], rtol=2e-7, atol=1e-12)
    if "renders_" + name in g:
        got = O.render_batch(inp, cfg).numpy()
        np.testing.assert_allclose(got, g["renders_" + name], rtol=0 if dtype == torch.float32 else 1e-12, atol=0)


def test_mixed_loss(golden):
    g, gm = golden("loss_bench"), golden("mixed")
    inp, tgt = t(g["input"]), t(g["target"])
    torch.manual_seed(int(gm["seed"]))
    cfg = O.sample_loss_configs(inp.shape[0])
    x = inp.clone().requires_grad_(True)
    val = O.mixed_loss(x, tgt, cfg)
    val.backward()
    np.testing.assert_array_equal(O.maps_l1_loss(inp, tgt).numpy(), gm["l1_f32"])
    np.testing.assert_array_equal(val.detach().numpy(), gm["loss_f32"])
    np.testing.assert_array_equal(x.grad.numpy(), gm["grad_f32"])


def test_pack_unpack_contract():
    # utils.py:186-239 pins the channel order [0:3]=normals,[3:6]=diffuse,[6:9]=roughness,[9:12]=specular
    maps = torch.arange(12.0).reshape(12, 1, 1).expand(12, 2, 2)
    n, d, r, s = O.split_maps(maps)
    assert n[:, 0, 0].tolist() == [0, 1, 2] and d[:, 0, 0].tolist() == [3, 4, 5]
    assert r[:, 0, 0].tolist() == [6, 7, 8] and s[:, 0, 0].tolist() == [9, 10, 11]
    assert torch.equal(O.join_maps(n, d, r, s), maps)
    with pytest.raises(ValueError):
        O.render([0, 0, 1], [0, 0, 1], [1, 1, 1], torch.zeros(12, 4, 6))


def test_decode_network_output(golden):
    g = golden("decode")
    np.testing.assert_array_equal(O.decode_network_output(t(g["encoded"])).numpy(), g["decoded_f32"])
