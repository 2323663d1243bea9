"""Parity of the CUDA path (through the reference-facing Python interface and the C ABI) with the
oracle and the golden fixtures of the unmodified reference.  Needs a GPU: ``pytest -m gpu``."""
import numpy as np
import pytest
import torch

from oracle import reference_port as O
from tests import parity
from tests.common import synthetic_maps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import svbrdf_estimation_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device"
    return pkg


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ours_loss_and_grad(S, inp, tgt, cfg):
    x = cu(inp).requires_grad_(True)
    loss = S.rendering_loss_with_records(x, cu(tgt), torch.from_numpy(cfg))
    loss.backward()
    return float(loss), x.grad.cpu().numpy()


# ---- golden fixtures (outputs of the unmodified reference, fp32 and fp64) ------------------------

@pytest.mark.parametrize("fixture", ["loss_bench", "loss_stress", "loss_n27", "loss_real"])
def test_loss_and_gradient_vs_reference(S, golden, fixture):
    g = golden(fixture)
    loss, grad = ours_loss_and_grad(S, g["input"], g["target"], g["configs"])
    parity.check_loss(loss, g["loss_f64"])
    parity.check_grad_groups(grad, g["grad_f32"], g["grad_f64"], rel=parity.REL_L2_STRESS if "stress" in fixture else parity.REL_L2)


@pytest.mark.parametrize("fixture", ["loss_bench", "loss_stress"])
def test_renders_vs_reference(S, golden, fixture):
    g = golden(fixture)
    got = S.render_records(cu(g["input"]), torch.from_numpy(g["configs"])).cpu().numpy()
    parity.check_tensor(got, g["renders_f32"], g["renders_f64"], "renders", rel=parity.REL_L2_STRESS if "stress" in fixture else parity.REL_L2)
    dlog = np.abs(np.log(got.astype(np.float64) + 0.1) - np.log(g["renders_f64"] + 0.1)).max()
    assert dlog < 2e-3, dlog
    if fixture == "loss_bench":
        # render_forward evaluates the GGX denominator in its accurate-highlight form (shading.cuh, ACC): on unit normals
        # it is several times closer to the fp64 reference than the reference's own fp32 run, and within 1e-4 of that run
        e64, floor = parity.rel_l2(got, g["renders_f64"]), parity.rel_l2(g["renders_f32"], g["renders_f64"])
        assert e64 <= 0.25 * floor and e64 <= 2e-5, (e64, floor)


def test_render_interface_fixed_scenes(S, golden):
    """LocalRenderer.render(scene, svbrdf): 4-D and 3-D inputs, list-valued scene fields
    (renderers.py:284 and final-viz.ipynb cell 11 scenes)."""
    g = golden("render_fixed")
    maps = cu(g["maps"])
    r = S.LocalRenderer()
    for k, cfg in enumerate(g["configs"]):
        scene = S.Scene(S.Camera(cfg[0:3].tolist()), S.Light(cfg[3:6].tolist(), cfg[6:9].tolist()))
        out = r.render(scene, maps)
        assert out.shape == (2, 3, 24, 24) and out.device == maps.device
        parity.check_tensor(out.cpu().numpy(), g["render4d_f32"][k], g["render4d_f64"][k], "render4d[%d]" % k)
    scene = S.Scene(S.Camera(g["configs"][0, 0:3]), S.Light(torch.from_numpy(g["configs"][0, 3:6]), g["configs"][0, 6:9]))
    out3 = r.render(scene, maps[0])
    assert out3.shape == (1, 3, 24, 24)
    parity.check_tensor(out3.cpu().numpy(), g["render3d_f32"], g["render3d_f64"], "render3d")
    out5 = r.render(scene, maps.reshape(1, 2, 12, 24, 24))
    assert out5.shape == (1, 2, 3, 24, 24)
    assert torch.equal(out5[0, 0:1], out3)


def test_mixed_loss_vs_reference(S, golden):
    g, gm = golden("loss_bench"), golden("mixed")
    torch.manual_seed(int(gm["seed"]))
    x = cu(g["input"]).requires_grad_(True)
    val = S.MixedLoss(S.LocalRenderer())(x, cu(g["target"]))
    val.backward()
    assert val.dim() == 0
    assert abs(float(val) - float(gm["loss_f32"])) <= 3e-6 * abs(float(gm["loss_f32"]))
    o32 = gm["grad_f32"]
    torch.manual_seed(int(gm["seed"]))
    cfg = O.sample_loss_configs(g["input"].shape[0])
    x64 = torch.from_numpy(g["input"]).double().requires_grad_(True)
    O.mixed_loss(x64, torch.from_numpy(g["target"]).double(), cfg, 0.1).backward()
    for name, s in parity.GROUPS:
        got = x.grad.cpu().numpy()[:, s]
        assert parity.rel_l2(got, x64.grad.numpy()[:, s]) <= parity.REL_L2, name          # vs the fp64 oracle: 1e-4
        floor = parity.rel_l2(o32[:, s], x64.grad.numpy()[:, s])                          # the reference's own fp32 run
        assert parity.rel_l2(got, o32[:, s]) <= max(parity.REL_L2, 2 * floor), name
    # unfused SVBRDFL1Loss on the device
    l1 = S.SVBRDFL1Loss()(cu(g["input"]), cu(g["target"]))
    assert abs(float(l1) - float(gm["l1_f32"])) <= 1e-5 * float(gm["l1_f32"])


# ---- seeded inputs at sizes the oracle finishes in seconds --------------------------------------

@pytest.mark.parametrize("size,batch,stress", [(64, 3, False), (48, 2, True), (37, 2, False)])
def test_loss_and_gradient_vs_oracle(S, size, batch, stress):
    inp, tgt = synthetic_maps(batch, size, 11, stress), synthetic_maps(batch, size, 12, stress)
    torch.manual_seed(313)
    cfg = O.sample_loss_configs(batch)
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), cfg)
    l32, g32 = O.rendering_loss_and_grad(inp, tgt, cfg)
    loss, grad = ours_loss_and_grad(S, inp.numpy(), tgt.numpy(), cfg.numpy())
    parity.check_loss(loss, float(l64))
    parity.check_grad_groups(grad, g32.numpy(), g64.numpy(), rel=parity.REL_L2_STRESS if stress else parity.REL_L2)


def test_rendering_loss_module_draws_reference_scenes(S):
    """RenderingLoss(LocalRenderer())(input, target) after torch.manual_seed(s) evaluates the same
    scenes as the reference (losses.py:35) -> equals the oracle with oracle-sampled scenes."""
    inp, tgt = synthetic_maps(2, 32, 21), synthetic_maps(2, 32, 22)
    torch.manual_seed(1234)
    cfg = O.sample_loss_configs(2)
    want = float(O.rendering_loss(inp.double(), tgt.double(), cfg))
    mod = S.RenderingLoss(S.LocalRenderer())
    torch.manual_seed(1234)
    got = mod(inp.cuda(), tgt.cuda())
    assert got.dim() == 0 and got.is_cuda
    parity.check_loss(float(got), want)
    # the two configuration counts are honoured at call time (losses.py:26-27)
    mod.random_configuration_count, mod.specular_configuration_count = 9, 18
    torch.manual_seed(77)
    cfg27 = O.sample_loss_configs(2, 9, 18)
    torch.manual_seed(77)
    parity.check_loss(float(mod(inp.cuda(), tgt.cuda())), float(O.rendering_loss(inp.double(), tgt.double(), cfg27)))


def test_render_backward_vs_oracle_autograd(S):
    maps = synthetic_maps(2, 40, 31, stress=True)
    torch.manual_seed(5)
    cfg = O.sample_loss_configs(2)
    w = torch.randn(2, 9, 3, 40, 40)
    m64 = maps.double().requires_grad_(True)
    (O.render_batch(m64, cfg) * w.double()).sum().backward()
    m32 = maps.clone().requires_grad_(True)
    (O.render_batch(m32, cfg) * w).sum().backward()
    x = maps.cuda().requires_grad_(True)
    (S.render_records(x, cfg) * w.cuda()).sum().backward()
    parity.check_grad_groups(x.grad.cpu().numpy(), m32.grad.numpy(), m64.grad.numpy(), "render_bwd", rel=parity.REL_L2_RAW_RENDER_GRAD)


@pytest.mark.parametrize("n_records", [2, 3, 4, 7])
@pytest.mark.parametrize("size", [18, 15])
def test_render_backward_record_ring(S, n_records, size):
    """The upstream gradient streams through a per-thread cp.async ring (3 records in flight): record counts below,
    at and above the ring depth, packed (even width) and one-pixel-per-thread (odd width) kernels."""
    maps = synthetic_maps(2, size, 61 + n_records)
    torch.manual_seed(9)
    cfg = O.sample_loss_configs(2, n_records - 1, 1)
    w = torch.randn(2, n_records, 3, size, size)
    m64 = maps.double().requires_grad_(True)
    (O.render_batch(m64, cfg) * w.double()).sum().backward()
    x = maps.cuda().requires_grad_(True)
    (S.render_records(x, cfg) * w.cuda()).sum().backward()
    for name, s in parity.GROUPS:
        assert parity.rel_l2(x.grad.cpu().numpy()[:, s], m64.grad.numpy()[:, s]) <= parity.REL_L2_RAW_RENDER_GRAD, name


def test_fixed_scene_loss_through_render_autograd(S):
    """Notebook-style loss (website.ipynb cell 15): plain L1 between renders, autograd through render()."""
    maps, tgt = synthetic_maps(1, 32, 41), synthetic_maps(1, 32, 42)
    cfg = torch.tensor([[0.0, -1.0, 2.0, 0.0, 0.0, 2.0, 50.0, 50.0, 50.0]])
    scene = S.Scene(S.Camera([0.0, -1.0, 2.0]), S.Light([0.0, 0.0, 2.0], [50.0, 50.0, 50.0]))
    m64 = maps.double().requires_grad_(True)
    ref = torch.nn.functional.l1_loss(O.render(cfg[0, 0:3], cfg[0, 3:6], cfg[0, 6:9], m64),
                                      O.render(cfg[0, 0:3], cfg[0, 3:6], cfg[0, 6:9], tgt.double()))
    ref.backward()
    x = maps.cuda().requires_grad_(True)
    r = S.LocalRenderer()
    got = torch.nn.functional.l1_loss(r.render(scene, x), r.render(scene, tgt.cuda()))
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-5 * float(ref)
    for name, s in parity.GROUPS:
        assert parity.rel_l2(x.grad.cpu().numpy()[:, s], m64.grad.numpy()[:, s]) <= 2e-4, name


# ---- interface behaviour ---------------------------------------------------------------------------

def test_upstream_gradient_scaling_and_no_grad(S):
    inp, tgt = synthetic_maps(2, 16, 51).cuda(), synthetic_maps(2, 16, 52).cuda()
    cfg = O.sample_loss_configs(2)
    x1 = inp.clone().requires_grad_(True)
    S.rendering_loss_with_records(x1, tgt, cfg).backward()
    x2 = inp.clone().requires_grad_(True)
    (S.rendering_loss_with_records(x2, tgt, cfg) * 2.5).backward()
    torch.testing.assert_close(x2.grad, x1.grad * 2.5, rtol=1e-6, atol=0)
    with torch.no_grad():
        v = S.rendering_loss_with_records(inp, tgt, cfg)
    assert not v.requires_grad
    # forward-only and forward+backward kernels are different instruction streams: equal to rounding
    assert abs(float(v) - float(S.rendering_loss_with_records(x1.detach(), tgt, cfg))) <= 1e-6 * float(v)
    # gradient w.r.t. the target (symmetric loss)
    t = tgt.clone().requires_grad_(True)
    S.rendering_loss_with_records(inp, t, cfg).backward()
    x3 = tgt.clone().requires_grad_(True)
    S.rendering_loss_with_records(x3, inp, cfg).backward()
    torch.testing.assert_close(t.grad, x3.grad, rtol=0, atol=0)


def test_identical_maps_zero_loss_zero_grad(S):
    inp = synthetic_maps(2, 32, 61).cuda().requires_grad_(True)
    loss = S.rendering_loss_with_records(inp, inp.detach().clone(), O.sample_loss_configs(2))
    loss.backward()
    assert float(loss) == 0.0 and not bool(inp.grad.any())


def test_host_tensors_are_staged_through_the_gpu(S):
    maps = synthetic_maps(1, 16, 71)
    scene = S.Scene(S.Camera([0.0, -1.0, 2.0]), S.Light([0.0, 0.0, 2.0], torch.tensor([50.0, 50.0, 50.0])))
    out = S.LocalRenderer().render(scene, maps)          # dataset.py:206-212 calls it like this
    assert out.device.type == "cpu" and out.shape == (1, 3, 16, 16)
    torch.testing.assert_close(out, S.LocalRenderer().render(scene, maps.cuda()).cpu(), rtol=0, atol=0)


def test_argument_errors(S):
    r = S.LocalRenderer()
    scene = S.Scene(S.Camera([0, 0, 1]), S.Light([0, 0, 1], [1, 1, 1]))
    with pytest.raises(ValueError):
        r.render(scene, torch.zeros(12, 8, 6).cuda())          # non-square (renderers.py:73-76)
    with pytest.raises(ValueError):
        r.render(scene, torch.zeros(11, 8, 8).cuda())
    with pytest.raises(TypeError):
        r.render(scene, torch.zeros(12, 8, 8, dtype=torch.int32).cuda())
    with pytest.raises(ValueError):
        S.rendering_loss_with_records(torch.zeros(2, 12, 8, 8).cuda(), torch.zeros(1, 12, 8, 8).cuda(), torch.zeros(2, 9, 9))
    with pytest.raises(ValueError):
        S.rendering_loss_with_records(torch.zeros(2, 12, 8, 8).cuda(), torch.zeros(2, 12, 8, 8).cuda(), torch.zeros(3, 9, 9))


def test_plugin_renderer_falls_back_to_generic_loop(S):
    """A renderer object without the fused marker goes through render() per scene (plugin API)."""
    class Wrapped:
        def __init__(self):
            self.inner, self.calls = S.LocalRenderer(), 0

        def render(self, scene, svbrdf):
            self.calls += 1
            return self.inner.render(scene, svbrdf)

    inp, tgt = synthetic_maps(2, 16, 81).cuda(), synthetic_maps(2, 16, 82).cuda()
    w = Wrapped()
    torch.manual_seed(9)
    a = S.RenderingLoss(w)(inp, tgt)
    torch.manual_seed(9)
    b = S.RenderingLoss(S.LocalRenderer())(inp, tgt)
    assert w.calls == 2 * 2 * 9
    assert abs(float(a) - float(b)) <= 2e-6 * float(b)


def test_ragged_sizes_and_many_records(S):
    """Odd map sizes (tail CTA), 1x1 maps, N that needs several launches (records > parameter block)."""
    for size, batch, n in ((1, 1, 1), (5, 3, 2), (33, 2, 9), (17, 40, 27)):
        inp, tgt = synthetic_maps(batch, size, 91), synthetic_maps(batch, size, 92)
        cfg = O.sample_loss_configs(batch, n // 3 if n >= 3 else 1, n - (n // 3 if n >= 3 else 1)) if n > 1 \
            else torch.tensor([[[0.3, -0.2, 1.5, -0.4, 0.1, 2.0, 30.0, 20.0, 10.0]]])
        cfg = cfg[:, :n].contiguous()
        l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), cfg)
        loss, grad = ours_loss_and_grad(S, inp.numpy(), tgt.numpy(), cfg.numpy())
        assert abs(loss - float(l64)) <= 5e-6 * float(l64), (size, batch, n)
        assert parity.rel_l2(grad, g64.numpy()) <= 2e-4, (size, batch, n)


def test_partially_identical_maps_are_exact(S, golden):
    """Only diffuse channel 0 differs: channels 1 and 2 render identically in the reference and contribute
    exactly 0; the kernels take the channel-wise path and mask them."""
    g = golden("loss_bench")
    tgt = g["target"].copy()
    inp = tgt.copy()
    inp[:, 3] = g["input"][:, 3]
    cfg = torch.from_numpy(g["configs"])
    l64, g64 = O.rendering_loss_and_grad(torch.from_numpy(inp).double(), torch.from_numpy(tgt).double(), cfg)
    loss, grad = ours_loss_and_grad(S, inp, tgt, g["configs"])
    parity.check_loss(loss, float(l64))
    for ch in (4, 5, 7, 8, 10, 11):
        assert not grad[:, ch].any(), ch
    assert parity.rel_l2(grad, g64.numpy()) <= 1e-4


def test_coloured_light_takes_the_general_colour_path(S):
    """Non-grey light colours (r != g != b) select the non-GREY kernels."""
    inp, tgt = synthetic_maps(2, 32, 101), synthetic_maps(2, 32, 102)
    torch.manual_seed(3)
    cfg = O.sample_loss_configs(2)
    cfg[..., 6] *= 0.5
    cfg[..., 8] *= 1.5
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), cfg)
    loss, grad = ours_loss_and_grad(S, inp.numpy(), tgt.numpy(), cfg.numpy())
    parity.check_loss(loss, float(l64))
    assert parity.rel_l2(grad, g64.numpy()) <= 1e-4


def test_encoded_input_mixed_loss(S, golden):
    """MixedLoss.forward_encoded: decode of the network output and its chain rule inside the kernel."""
    gd = golden("decode")
    enc = torch.from_numpy(gd["encoded"])
    tgt = torch.from_numpy(golden("loss_bench")["target"][:, :, :12, :12].copy())
    torch.manual_seed(4)
    cfg = O.sample_loss_configs(2)
    e64 = enc.double().requires_grad_(True)
    want = O.mixed_loss(O.decode_network_output(e64), tgt.double(), cfg, 0.1)
    want.backward()
    x = enc.cuda().requires_grad_(True)
    out = S.mixed_loss_from_encoded(x, tgt.cuda(), cfg, 0.1)
    out[0].backward()
    assert abs(float(out[0]) - float(want)) <= 3e-6 * float(want)
    g64 = e64.grad.numpy()
    for name, s in (("normal_xy", slice(0, 2)), ("diffuse", slice(2, 5)), ("roughness", slice(5, 6)), ("specular", slice(6, 9))):
        assert parity.rel_l2(x.grad.cpu().numpy()[:, s], g64[:, s]) <= parity.REL_L2, name
    # module form draws the reference's scenes and equals the decoded 12-channel path
    torch.manual_seed(9)
    a = S.MixedLoss(S.LocalRenderer()).forward_encoded(enc.cuda(), tgt.cuda())
    torch.manual_seed(9)
    b = S.MixedLoss(S.LocalRenderer())(S.utils.decode_network_output(enc.cuda()), tgt.cuda())
    assert abs(float(a) - float(b)) <= 3e-6 * float(b)


def test_map_optimisation_converges(S):
    """The reference's de-facto integration test of the gradients (website.ipynb cell 15, final-viz.ipynb
    cells 11/14): start from random maps and run Adam through RenderingLoss(LocalRenderer()); the loss
    must fall steadily (the notebook reports 0.049 after 200 steps on its sample)."""
    torch.manual_seed(0)
    target = synthetic_maps(1, 64, 7).cuda()
    est = torch.rand_like(target)
    est[:, 0:3] = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 1.0]).view(1, 3, 1, 1).expand(1, 3, 64, 64).cuda() + 0.01 * torch.randn(1, 3, 64, 64, device="cuda"), dim=1)
    est.requires_grad_(True)
    opt = torch.optim.Adam([est], lr=0.02)
    loss_fn = S.RenderingLoss(S.LocalRenderer())
    first = last = None
    for it in range(150):
        opt.zero_grad()
        loss = loss_fn(est, target)
        loss.backward()
        opt.step()
        with torch.no_grad():
            est[:, 3:].clamp_(0.0, 1.0)
        v = float(loss.detach())
        first = v if first is None else first
        last = v
    assert np.isfinite(last) and last < 0.45 * first, (first, last)


def test_cuda_graph_capture_and_replay(S):
    """All work is stream-ordered and allocation-free inside the C ABI, so a loss evaluation can be captured
    into a CUDA graph (scene records are baked into the kernel parameters at capture time)."""
    from svbrdf_estimation_b200 import _cabi
    lib = _cabi.lib()
    inp, tgt = synthetic_maps(2, 32, 1).cuda(), synthetic_maps(2, 32, 2).cuda()
    rec = O.sample_loss_configs(2)
    lin = torch.linspace(-1, 1, 32, device="cuda")
    nbytes = lib.svbrdf_b200_workspace_bytes(2, 9, 32, 32)
    ws = torch.empty(nbytes // 4 + 1, device="cuda")
    out, grad = torch.zeros(1, device="cuda"), torch.zeros_like(inp)

    def launch():
        _cabi.check(lib.svbrdf_b200_loss_forward_backward(inp.data_ptr(), tgt.data_ptr(), 2, 32, 32, rec.data_ptr(), 9,
                                                          lin.data_ptr(), out.data_ptr(), grad.data_ptr(), ws.data_ptr(),
                                                          nbytes, torch.cuda.current_stream().cuda_stream))
    launch()
    torch.cuda.synchronize()
    want, want_g = float(out), grad.clone()
    graph = torch.cuda.CUDAGraph()
    out.zero_(); grad.zero_()
    with torch.cuda.graph(graph):
        launch()
    out.zero_(); grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert float(out) == want and torch.equal(grad, want_g)


def test_half_precision_inputs_are_upcast_and_second_device(S):
    inp, tgt = synthetic_maps(2, 16, 3).cuda(), synthetic_maps(2, 16, 4).cuda()
    cfg = O.sample_loss_configs(2)
    x16 = inp.to(torch.bfloat16).requires_grad_(True)
    loss = S.rendering_loss_with_records(x16, tgt, cfg)
    loss.backward()
    assert loss.dtype == torch.float32 and x16.grad.dtype == torch.bfloat16
    ref = S.rendering_loss_with_records(x16.detach().float(), tgt, cfg)
    assert abs(float(loss.detach()) - float(ref)) <= 1e-6 * float(ref)   # fwd+bwd vs forward-only kernel: rounding
    if torch.cuda.device_count() > 1:                       # tensors on a device that is not the current one
        a = inp.to("cuda:1").requires_grad_(True)
        l1 = S.rendering_loss_with_records(a, tgt.to("cuda:1"), cfg)
        l1.backward()
        assert l1.device.index == 1 and a.grad.device.index == 1
        assert abs(float(l1.detach()) - float(S.rendering_loss_with_records(inp, tgt, cfg))) <= 1e-7


def test_degenerate_maps_stay_finite_and_match(S, golden):
    """Zero normals, zero / one albedos, zero and >1 roughness, light and camera straight above a pixel."""
    g = golden("loss_bench")
    tgt = g["target"][:, :, :8, :8].copy()
    inp = np.zeros_like(tgt)
    inp[0, 3:6], inp[0, 6:9], inp[0, 9:12] = 0.0, 0.0, 1.0
    inp[1, 0:3] = np.array([0.0, 0.0, 1.0], dtype=np.float32).reshape(3, 1, 1)
    inp[1, 3:6], inp[1, 6:9], inp[1, 9:12] = 1.0, 1.5, 0.0
    cfg = g["configs"][:, :4].copy()
    cfg[0, 0] = [0.0, 0.0, 1.0, 0.0, 0.0, 1.0, 20.0, 20.0, 20.0]
    cfg[1, 1] = [-1.0, 1.0, 0.5, -1.0, 1.0, 0.5, 50.0, 50.0, 50.0]
    l64, g64 = O.rendering_loss_and_grad(torch.from_numpy(inp).double(), torch.from_numpy(tgt).double(), torch.from_numpy(cfg))
    loss, grad = ours_loss_and_grad(S, inp, tgt, cfg)
    assert np.isfinite(loss) and np.isfinite(grad).all()
    assert abs(loss - float(l64)) <= 5e-6 * float(l64)
    assert parity.rel_l2(grad, g64.numpy()) <= 3e-4


def test_dataset_input_synthesis_matches_reference(S, golden):
    """inputs.render_inputs == SvbrdfDataset.render_inputs (dataset.py:162-221) for the same seed."""
    from svbrdf_estimation_b200 import inputs as I
    g = golden("dataset_inputs")
    for tag, aug in (("plain", False), ("aug", True)):
        for dev in ("cuda", "cpu"):
            torch.manual_seed(int(g["seed"]))
            got = I.render_inputs(torch.from_numpy(g["svbrdf"]).to(dev), 3, use_augmentation=aug)
            assert got.device.type == dev and got.shape == (3, 3, 16, 16)
            np.testing.assert_allclose(got.cpu().numpy(), g["inputs_" + tag], rtol=2e-4, atol=2e-5)
    fast = I.render_inputs(torch.from_numpy(g["svbrdf"]).cuda(), 4, use_augmentation=True, noise="device")
    assert fast.shape == (4, 3, 16, 16) and float(fast.min()) >= 0.0 and float(fast.max()) <= 1.0


def _oracle_loss_chunks(inp, tgt, cfg, dev, chunk):
    """fp64 oracle loss of a large batch, evaluated forward-only in chunks on the GPU (equal chunk sizes)."""
    B = inp.shape[0]
    assert B % chunk == 0
    with torch.no_grad():
        vals = [float(O.rendering_loss(inp[i:i + chunk].double().to(dev), tgt[i:i + chunk].double().to(dev), cfg[i:i + chunk]))
                for i in range(0, B, chunk)]
    return sum(vals) / len(vals)


# (name, batch, size, n_random, n_specular, batch elements the fp64 oracle differentiates, forward chunk)
FULL_SIZE_CASES = [
    ("c1_b8_256_n9", 8, 256, 3, 6, 8, 8),            # BASELINE.json configs[0] at its real batch
    ("c2_b64_256_n9", 64, 256, 3, 6, 8, 8),          # configs[1]: the kernel runs all 64 maps, the oracle every 8th
    ("c3_b32_256_n27", 32, 256, 9, 18, 4, 4),        # configs[2]
    ("c4_b16_1024_n9", 16, 1024, 3, 6, 1, 1),        # configs[3]: renderers.py:73-76 coordinates at W = 1024
    ("c5_b32_256_n9", 32, 256, 3, 6, 8, 8),          # configs[4], one rank's slice
]


@pytest.mark.parametrize("case", FULL_SIZE_CASES, ids=[c[0] for c in FULL_SIZE_CASES])
def test_full_resolution_parity_at_the_baseline_configs(S, case):
    """BASELINE.json configs at full size (losses.py:34-50, renderers.py:73-76): the CUDA path runs the whole batch;
    the oracle runs on the GPU in fp64 - the loss over the whole batch (forward only, in chunks), the gradient for a
    strided subset of the batch elements.  Gradients are compared over ALL pixels: the only thing set aside are the
    individual L1 terms the fp32 evaluation took with the other sign, each of which must be sign-ambiguous in fp64
    (0 < |l| < 1e-3) - tests/parity.py flipped_term_analysis."""
    name, batch, size, nr, ns, n_sub, chunk = case
    dev = torch.device("cuda", 0)
    inp, tgt = synthetic_maps(batch, size, 21), synthetic_maps(batch, size, 22)
    torch.manual_seed(313)
    cfg = O.sample_loss_configs(batch, nr, ns)
    x = inp.to(dev).requires_grad_(True)
    loss = S.rendering_loss_with_records(x, tgt.to(dev), cfg)
    loss.backward()
    parity.check_loss(float(loss), _oracle_loss_chunks(inp, tgt, cfg, dev, chunk))
    sub = list(range(0, batch, batch // n_sub))[:n_sub]
    g = x.grad[sub].cpu().numpy()
    res = parity.flipped_term_analysis(O, inp[sub], tgt[sub], cfg[sub], g, scale=batch / float(len(sub)), device=dev)
    out = parity.check_grad_with_flipped_terms(res, name)
    print(name, "candidates %d flipped %d max|l| %.2e" % (res["candidates"], res["flipped"], res["max_abs_l_flipped"]), out)


def test_interface_second_backward_eval_mode_and_dtype_following(S):
    """retain_graph=True allows a second backward (the reference's eager graph does); an .eval()-mode loss computes no
    gradient up front but backward() still works; float64 maps give float64 results (renderers.py:68: dtype follows)."""
    inp, tgt = synthetic_maps(2, 32, 5).cuda(), synthetic_maps(2, 32, 6).cuda()
    cfg = O.sample_loss_configs(2)
    x = inp.clone().requires_grad_(True)
    loss = S.rendering_loss_with_records(x, tgt, cfg)
    loss.backward(retain_graph=True)
    g1 = x.grad.clone()
    loss.backward()                                           # second time: gradient recomputed, accumulated
    torch.testing.assert_close(x.grad, 2 * g1, rtol=1e-6, atol=0)
    with pytest.raises(RuntimeError):
        loss.backward()                                       # graph freed now, like any autograd graph
    with pytest.raises(RuntimeError):                         # once differentiable: no silent zero second derivative
        y = inp.clone().requires_grad_(True)
        torch.autograd.grad(S.rendering_loss_with_records(y, tgt, cfg), y, create_graph=True)[0].sum().backward()
    # validation mode (main.py:140 evaluates the loss without no_grad): no gradient pass, same value, lazy backward
    mod = S.MixedLoss(S.LocalRenderer()).eval()
    x2 = inp.clone().requires_grad_(True)
    torch.manual_seed(4)
    v = mod(x2, tgt)
    torch.manual_seed(4)
    x3 = inp.clone().requires_grad_(True)
    w = S.MixedLoss(S.LocalRenderer())(x3, tgt)
    assert abs(float(v) - float(w)) <= 1e-6 * float(w)
    v.backward(); w.backward()
    torch.testing.assert_close(x2.grad, x3.grad, rtol=0, atol=0)
    # dtype following
    r = S.LocalRenderer()
    scene = S.Scene(S.Camera([0.0, -1.0, 2.0]), S.Light([0.0, 0.0, 2.0], [50.0, 50.0, 50.0]))
    img64 = r.render(scene, inp.double())
    assert img64.dtype == torch.float64
    torch.testing.assert_close(img64, r.render(scene, inp).double(), rtol=0, atol=0)
    x64 = inp.double().requires_grad_(True)
    l64 = S.rendering_loss_with_records(x64, tgt.double(), cfg)
    l64.backward()
    assert l64.dtype == torch.float64 and x64.grad.dtype == torch.float64
    torch.testing.assert_close(x64.grad, g1.double(), rtol=1e-6, atol=0)
    # accurate-highlight kernels without a gradient: the forward-only kernel (no scratch gradient buffer)
    with torch.no_grad():
        va = S.rendering_loss_with_records(inp, tgt, cfg, accurate=True)
    xa = inp.clone().requires_grad_(True)
    vb = S.rendering_loss_with_records(xa, tgt, cfg, accurate=True)
    assert abs(float(va) - float(vb)) <= 1e-6 * float(vb)
    # encoded input on another device than the target is an error, not a cross-device launch
    if torch.cuda.device_count() > 1:
        with pytest.raises(ValueError):
            S.mixed_loss_from_encoded(torch.zeros(2, 9, 32, 32, device="cuda:1"), tgt, cfg)


def test_random_shapes_against_the_oracle_on_gpu(S):
    """scripts/fuzz_shapes.py through the CUDA library (see tests/test_kernel_math_emulation.py for the host twin)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_shapes.py"), "25", "11"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]


def test_accurate_rendering_loss(S, golden):
    """RenderingLoss(..., accurate=True) / rendering_loss_with_records(..., accurate=True): the accurate-highlight kernels.
    Golden fixture: gradients several times closer to fp64 than the reference's own fp32 run; full resolution: 1e-5."""
    g = golden("loss_bench")
    x = cu(g["input"]).requires_grad_(True)
    loss = S.rendering_loss_with_records(x, cu(g["target"]), torch.from_numpy(g["configs"]), accurate=True)
    loss.backward()
    parity.check_loss(float(loss), g["loss_f64"])
    grad = x.grad.cpu().numpy()
    for name, s in parity.GROUPS:
        e64 = parity.rel_l2(grad[:, s], g["grad_f64"][:, s])
        floor = parity.rel_l2(g["grad_f32"][:, s], g["grad_f64"][:, s])
        assert e64 <= max(0.3 * floor, 3e-7), (name, e64, floor)
    # module form, no-grad form, and agreement with the default kernels to the default kernels' accuracy
    mod = S.RenderingLoss(S.LocalRenderer(), accurate=True)
    torch.manual_seed(3)
    a = float(mod(cu(g["input"]), cu(g["target"])))
    torch.manual_seed(3)
    b = float(S.RenderingLoss(S.LocalRenderer())(cu(g["input"]), cu(g["target"])))
    assert abs(a - b) <= 2e-6 * abs(b)
    with pytest.raises(NotImplementedError):
        S.losses._fused_loss(x, cu(g["target"]), torch.from_numpy(g["configs"]), 0.1, True)
