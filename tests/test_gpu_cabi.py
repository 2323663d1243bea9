"""The C ABI called directly (ctypes, raw device pointers) and size-independent properties at the
BASELINE.json sizes.  Needs a GPU: ``pytest -m gpu``."""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests.common import synthetic_maps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from svbrdf_estimation_b200 import _cabi, environment
    assert torch.cuda.is_available()
    return _cabi.lib(), _cabi, environment


def device_loss(lib, _cabi, inp, tgt, rec, grad=True, mixed=None):
    B, _, H, W = inp.shape
    N = rec.shape[1]
    lin = torch.linspace(-1, 1, W, device=inp.device)
    nbytes = lib.svbrdf_b200_workspace_bytes(B, N, H, W)
    ws = torch.empty(nbytes // 4 + 1, device=inp.device)
    out = torch.zeros(3, device=inp.device)
    g = torch.empty_like(inp) if grad else None
    st = torch.cuda.current_stream().cuda_stream
    if mixed is not None:
        _cabi.check(lib.svbrdf_b200_mixed_loss_forward_backward(inp.data_ptr(), tgt.data_ptr(), B, H, W, rec.data_ptr(), N,
                                                                float(mixed), lin.data_ptr(), out.data_ptr(),
                                                                g.data_ptr() if grad else None, ws.data_ptr(), nbytes, st))
    elif grad:
        _cabi.check(lib.svbrdf_b200_loss_forward_backward(inp.data_ptr(), tgt.data_ptr(), B, H, W, rec.data_ptr(), N,
                                                          lin.data_ptr(), out.data_ptr(), g.data_ptr(), ws.data_ptr(), nbytes, st))
    else:
        _cabi.check(lib.svbrdf_b200_loss_forward(inp.data_ptr(), tgt.data_ptr(), B, H, W, rec.data_ptr(), N,
                                                 lin.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, st))
    torch.cuda.synchronize()
    return out.cpu(), g


def test_status_codes_on_device(env):
    lib, _cabi, E = env
    inp = synthetic_maps(1, 8, 1).cuda()
    rec = E.sample_loss_configs(1)
    lin = torch.linspace(-1, 1, 8, device="cuda")
    out = torch.zeros(1, device="cuda")
    ws = torch.empty(64, device="cuda")
    st = lib.svbrdf_b200_loss_forward(inp.data_ptr(), inp.data_ptr(), 1, 8, 8, rec.data_ptr(), 9, lin.data_ptr(),
                                      out.data_ptr(), ws.data_ptr(), 4, None)
    assert st == _cabi.E_INVALID and b"workspace" in lib.svbrdf_b200_last_error()
    st = lib.svbrdf_b200_loss_forward_backward(inp.data_ptr(), inp.data_ptr(), 1, 8, 8, rec.data_ptr(), 9, lin.data_ptr(),
                                               out.data_ptr(), None, ws.data_ptr(), 256, None)
    assert st == _cabi.E_INVALID
    st = lib.svbrdf_b200_scale_grad(None, 10, out.data_ptr(), None)
    assert st == _cabi.E_INVALID
    assert lib.svbrdf_b200_scale_grad(out.data_ptr(), 0, out.data_ptr(), None) == 0


def test_forward_only_forward_backward_and_mixed_agree(env):
    lib, _cabi, E = env
    inp, tgt = synthetic_maps(3, 40, 1).cuda(), synthetic_maps(3, 40, 2).cuda()
    torch.manual_seed(1)
    rec = E.sample_loss_configs(3)
    o_f, _ = device_loss(lib, _cabi, inp, tgt, rec, grad=False)
    o_b, g_b = device_loss(lib, _cabi, inp, tgt, rec, grad=True)
    o_m, g_m = device_loss(lib, _cabi, inp, tgt, rec, grad=True, mixed=0.1)
    o_m0, g_m0 = device_loss(lib, _cabi, inp, tgt, rec, grad=True, mixed=0.0)
    assert abs(float(o_f[0]) - float(o_b[0])) <= 1e-6 * float(o_b[0])
    assert abs(float(o_m[1]) - float(o_b[0])) <= 1e-6 * float(o_b[0])
    assert abs(float(o_m[0]) - (float(o_m[1]) + 0.1 * float(o_m[2]))) <= 1e-6 * float(o_m[0])
    torch.testing.assert_close(g_m0, g_b, rtol=1e-5, atol=1e-12)
    # map-L1 part against plain torch ops on the device
    n0, d0, r0, s0 = inp.split(3, dim=1)
    n1, d1, r1, s1 = tgt.split(3, dim=1)
    l1 = torch.nn.functional.l1_loss
    want = l1(n0, n1) + l1(torch.log(d0 + 0.01), torch.log(d1 + 0.01)) + l1(r0, r1) + l1(torch.log(s0 + 0.01), torch.log(s1 + 0.01))
    assert abs(float(o_m[2]) - float(want)) <= 2e-6 * float(want)


def test_unaligned_pointers_and_odd_width_take_the_scalar_kernels(env):
    """Odd W, or an even W whose planes start at an odd float offset (a contiguous view into a larger buffer: 4-byte but
    not 8-byte aligned), run the one-pixel-per-thread kernels; the result equals the packed kernels' to rounding and the
    oracle's within the parity tolerance.  Pointers that are not even float-aligned are rejected."""
    lib, _cabi, E = env
    import svbrdf_estimation_b200 as S
    from oracle import reference_port as O
    inp, tgt = synthetic_maps(2, 31, 5), synthetic_maps(2, 31, 6)
    torch.manual_seed(2)
    rec = E.sample_loss_configs(2)
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), rec)
    x = inp.cuda().requires_grad_(True)
    loss = S.rendering_loss_with_records(x, tgt.cuda(), rec)
    loss.backward()
    assert abs(float(loss) - float(l64)) <= 2e-6 * float(l64)
    assert float((x.grad.cpu().double() - g64).norm() / g64.norm()) <= 1e-4
    # even width, every tensor shifted by one float inside a larger allocation
    B, size = 2, 32
    inp, tgt = synthetic_maps(B, size, 7), synthetic_maps(B, size, 8)
    n = inp.numel()

    def shifted(t):
        buf = torch.empty(t.numel() + 1, device="cuda")
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.data_ptr() % 8 == 4 and v.is_contiguous()
        return v
    a, b = shifted(inp), shifted(tgt)
    o_al, g_al = device_loss(lib, _cabi, inp.cuda(), tgt.cuda(), rec)
    gbuf = torch.empty(n + 1, device="cuda")
    g_un = gbuf[1:].view(inp.shape)
    lin = shifted(torch.linspace(-1, 1, size))
    nbytes = lib.svbrdf_b200_workspace_bytes(B, 9, size, size)
    ws = torch.empty(nbytes // 4 + 1, device="cuda")
    out = torch.zeros(3, device="cuda")
    _cabi.check(lib.svbrdf_b200_loss_forward_backward(a.data_ptr(), b.data_ptr(), B, size, size, rec.data_ptr(), 9, lin.data_ptr(),
                                                      out.data_ptr(), g_un.data_ptr(), ws.data_ptr(), nbytes, None))
    torch.cuda.synchronize()
    assert abs(float(out[0]) - float(o_al[0])) <= 1e-6 * float(o_al[0])
    torch.testing.assert_close(g_un, g_al, rtol=2e-4, atol=1e-9)          # scalar vs packed lanes: different instruction streams
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), rec)
    assert float((g_un.cpu().double() - g64).norm() / g64.norm()) <= 1e-4
    # renderer: unaligned maps and images
    img_al = S.render_records(inp.cuda(), rec)
    imgbuf = torch.empty(B * 9 * 3 * size * size + 1, device="cuda")
    img_un = imgbuf[1:].view(B, 9, 3, size, size)
    _cabi.check(lib.svbrdf_b200_render_forward(a.data_ptr(), B, size, size, rec.data_ptr(), 9, 1, lin.data_ptr(), img_un.data_ptr(), None))
    torch.cuda.synchronize()
    torch.testing.assert_close(img_un, img_al, rtol=1e-5, atol=1e-7)
    # a pointer that is not float-aligned is an error
    st = lib.svbrdf_b200_loss_forward(a.data_ptr() + 1, b.data_ptr(), B, size, size, rec.data_ptr(), 9, lin.data_ptr(),
                                      out.data_ptr(), ws.data_ptr(), nbytes, None)
    assert st == _cabi.E_INVALID and b"aligned" in lib.svbrdf_b200_last_error()


@pytest.mark.parametrize("workload", ["c2", "c4"])
def test_full_size_properties(env, workload):
    """BASELINE.json sizes (too large for the oracle): determinism, batch-shard additivity, batch
    permutation equivariance, exact zero for identical maps, symmetry of the loss."""
    lib, _cabi, E = env
    B, size, N = (64, 256, 9) if workload == "c2" else (16, 1024, 9)
    inp, tgt = synthetic_maps(B, size, 11).cuda(), synthetic_maps(B, size, 12).cuda()
    torch.manual_seed(5)
    rec = E.sample_loss_configs(B)
    o1, g1 = device_loss(lib, _cabi, inp, tgt, rec)
    o2, g2 = device_loss(lib, _cabi, inp, tgt, rec)
    assert float(o1[0]) == float(o2[0]) and torch.equal(g1, g2)                     # run-to-run deterministic
    assert torch.isfinite(g1).all() and 0.0 < float(o1[0]) < 10.0
    # shards: mean of the shard losses = full loss; shard gradients * (b/B) = full gradient slices
    h = B // 2
    oa, ga = device_loss(lib, _cabi, inp[:h].contiguous(), tgt[:h].contiguous(), rec[:h].contiguous())
    ob, gb = device_loss(lib, _cabi, inp[h:].contiguous(), tgt[h:].contiguous(), rec[h:].contiguous())
    assert abs(0.5 * (float(oa[0]) + float(ob[0])) - float(o1[0])) <= 1e-6 * float(o1[0])
    torch.testing.assert_close(torch.cat((ga, gb)) * 0.5, g1, rtol=1e-5, atol=1e-14)
    # permutation of the batch permutes the gradient
    perm = torch.randperm(B)
    op, gp = device_loss(lib, _cabi, inp[perm].contiguous(), tgt[perm].contiguous(), rec[perm].contiguous())
    assert abs(float(op[0]) - float(o1[0])) <= 1e-6 * float(o1[0])
    assert torch.equal(gp, g1[perm])
    # identical maps: exactly zero
    oz, gz = device_loss(lib, _cabi, inp, inp.clone(), rec)
    assert float(oz[0]) == 0.0 and not bool(gz.any())
    # the loss is symmetric in (input, target)
    os_, _ = device_loss(lib, _cabi, tgt, inp, rec, grad=False)
    assert abs(float(os_[0]) - float(o1[0])) <= 2e-6 * float(o1[0])


def test_render_is_linear_in_light_colour_and_falls_off_with_distance(env):
    import svbrdf_estimation_b200 as S
    maps = synthetic_maps(2, 64, 3).cuda()
    base = torch.tensor([[0.2, -0.3, 1.5, -0.4, 0.5, 2.0, 10.0, 20.0, 30.0]])
    a = S.render_records(maps, base)
    b = S.render_records(maps, base * torch.tensor([1, 1, 1, 1, 1, 1, 2.0, 2.0, 2.0]))
    torch.testing.assert_close(b, 2 * a, rtol=2e-6, atol=0)
    assert (a >= 0).all()


def test_host_entry_matches_device_entry(env):
    lib, _cabi, E = env
    B, size, N = 6, 64, 9
    inp, tgt = synthetic_maps(B, size, 21), synthetic_maps(B, size, 22)
    torch.manual_seed(8)
    rec = E.sample_loss_configs(B)
    o, g = device_loss(lib, _cabi, inp.cuda(), tgt.cuda(), rec)
    ctx = ctypes.c_void_p()
    _cabi.check(lib.svbrdf_b200_ctx_create(ctypes.byref(ctx), 8, 9, size, size))
    try:
        grad = torch.empty_like(inp)
        loss = ctypes.c_float(0)
        # pageable host memory
        _cabi.check(lib.svbrdf_b200_rendering_loss_host(ctx, inp.data_ptr(), tgt.data_ptr(), B, rec.data_ptr(), N,
                                                        ctypes.byref(loss), grad.data_ptr()))
        assert abs(loss.value - float(o[0])) <= 1e-7 * float(o[0])
        assert torch.equal(grad, g.cpu())          # same kernels, same coordinate table (fill_lin == torch.linspace)
        # the context's own pinned buffers, forward only
        n = B * 12 * size * size
        pin = [lib.svbrdf_b200_ctx_pinned(ctx, w) for w in range(3)]
        ctypes.memmove(pin[0], inp.data_ptr(), n * 4)
        ctypes.memmove(pin[1], tgt.data_ptr(), n * 4)
        _cabi.check(lib.svbrdf_b200_rendering_loss_host(ctx, pin[0], pin[1], B, rec.data_ptr(), N, ctypes.byref(loss), None))
        assert abs(loss.value - float(o[0])) <= 1e-6 * float(o[0])
        assert lib.svbrdf_b200_rendering_loss_host(ctx, pin[0], pin[1], 9, rec.data_ptr(), N, ctypes.byref(loss), None) == _cabi.E_INVALID
    finally:
        lib.svbrdf_b200_ctx_destroy(ctx)
    assert lib.svbrdf_b200_rendering_loss_host(None, inp.data_ptr(), tgt.data_ptr(), B, rec.data_ptr(), N,
                                               ctypes.byref(loss), None) == _cabi.E_STATE


def test_work_is_enqueued_on_the_callers_stream(env):
    lib, _cabi, E = env
    import svbrdf_estimation_b200 as S
    inp, tgt = synthetic_maps(4, 64, 31).cuda(), synthetic_maps(4, 64, 32).cuda()
    rec = E.sample_loss_configs(4)
    ref = S.rendering_loss_with_records(inp, tgt, rec)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        got = S.rendering_loss_with_records(inp, tgt, rec)
    s.synchronize()
    assert float(got) == float(ref)


def test_plain_c_program_through_the_host_entry(env, tmp_path):
    """examples/c_abi_demo.c (gcc, include/svbrdf_b200.h + the .so only, no PyTorch in the process) computes the same loss
    and the bit-identical gradient as the Python layer on the inputs it writes out."""
    import shutil
    import subprocess
    import numpy as np
    import svbrdf_estimation_b200 as S
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "svbrdf_estimation_b200")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "c_abi_demo.c"),
                           "-o", exe, "-L" + pkg, "-lsvbrdf_b200", "-Wl,-rpath," + pkg, "-lm"])
    out = subprocess.run([exe, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    c_loss = float(out.stdout.split()[1])
    B, N, H, W = 3, 9, 40, 40
    load = lambda name, shape: torch.from_numpy(np.fromfile(str(tmp_path / name), dtype=np.float32).reshape(shape).copy())
    inp, tgt = load("input.f32", (B, 12, H, W)), load("target.f32", (B, 12, H, W))
    rec, c_grad = load("records.f32", (B, N, 9)), load("grad.f32", (B, 12, H, W))
    x = inp.cuda().requires_grad_(True)
    loss = S.rendering_loss_with_records(x, tgt.cuda(), rec)
    loss.backward()
    assert abs(float(loss) - c_loss) <= 1e-6 * c_loss
    assert torch.equal(x.grad.cpu(), c_grad)
    from oracle import reference_port as O                      # and both agree with the reference's math
    l64, g64 = O.rendering_loss_and_grad(inp.double(), tgt.double(), rec)
    assert abs(c_loss - float(l64)) <= 2e-6 * float(l64)
    assert float((c_grad.double() - g64).norm() / g64.norm()) <= 1e-4


def test_channel_layouts_device_and_host(env):
    """svbrdf_b200_loss_layouts / svbrdf_b200_loss_host: 10-channel maps (roughness stored once) and the 9-channel encoded
    input with a 10-channel target give the 12-channel results - same loss bits, the single roughness gradient is the sum of
    the three - on device pointers and through the host pipeline."""
    lib, _cabi, E = env
    import svbrdf_estimation_b200 as S
    B, size, N = 5, 64, 9
    inp, tgt = synthetic_maps(B, size, 41).cuda(), synthetic_maps(B, size, 42).cuda()
    torch.manual_seed(3)
    rec = E.sample_loss_configs(B)
    to10 = lambda m: torch.cat((m[:, 0:7], m[:, 9:12]), dim=1).contiguous()
    lin = torch.linspace(-1, 1, size, device="cuda")
    nbytes = lib.svbrdf_b200_workspace_bytes(B, N, size, size)
    ws = torch.empty(nbytes // 4 + 1, device="cuda")

    def layouts(a, la, b, lb, l1w, grad):
        out = torch.zeros(3, device="cuda")
        g = torch.empty_like(a) if grad else None
        _cabi.check(lib.svbrdf_b200_loss_layouts(a.data_ptr(), la, b.data_ptr(), lb, B, size, size, rec.data_ptr(), N, l1w,
                                                 lin.data_ptr(), out.data_ptr(), g.data_ptr() if grad else None, ws.data_ptr(), nbytes, None))
        torch.cuda.synchronize()
        return out.cpu(), g
    o12, g12 = device_loss(lib, _cabi, inp, tgt, rec)
    o10, g10 = layouts(to10(inp), 10, to10(tgt), 10, -1.0, True)
    assert float(o10[0]) == float(o12[0])
    assert torch.equal(g10[:, 0:6], g12[:, 0:6]) and torch.equal(g10[:, 7:10], g12[:, 9:12])
    assert torch.equal(g10[:, 6], (g12[:, 6] + g12[:, 7]) + g12[:, 8])
    o10f, _ = layouts(to10(inp), 10, to10(tgt), 10, -1.0, False)
    assert abs(float(o10f[0]) - float(o12[0])) <= 1e-6 * float(o12[0])
    # encoded input: 12- vs 10-channel target
    enc = (torch.rand(B, 9, size, size, device="cuda") * 1.8 - 0.9)
    oe12, ge12 = layouts(enc, 9, tgt, 12, 0.1, True)
    oe10, ge10 = layouts(enc, 9, to10(tgt), 10, 0.1, True)
    assert torch.equal(oe12, oe10) and torch.equal(ge12, ge10)
    ref = S.mixed_loss_from_encoded(enc, tgt, rec, 0.1)
    assert abs(float(oe12[0]) - float(ref[0])) <= 1e-6 * float(ref[0])
    # unsupported combinations are errors
    out = torch.zeros(3, device="cuda")
    bad = lib.svbrdf_b200_loss_layouts(inp.data_ptr(), 12, tgt.data_ptr(), 10, B, size, size, rec.data_ptr(), N, -1.0, lin.data_ptr(),
                                       out.data_ptr(), None, ws.data_ptr(), nbytes, None)
    assert bad == _cabi.E_INVALID and b"layout" in lib.svbrdf_b200_last_error()
    assert lib.svbrdf_b200_loss_layouts(enc.data_ptr(), 9, tgt.data_ptr(), 12, B, size, size, rec.data_ptr(), N, -1.0, lin.data_ptr(),
                                        out.data_ptr(), None, ws.data_ptr(), nbytes, None) == _cabi.E_INVALID
    # host pipeline, pageable host memory
    ctx = ctypes.c_void_p()
    _cabi.check(lib.svbrdf_b200_ctx_create(ctypes.byref(ctx), 8, 9, size, size))
    try:
        h_in, h_tg = to10(inp).cpu(), to10(tgt).cpu()
        h_g = torch.empty_like(h_in)
        res = (ctypes.c_float * 3)()
        _cabi.check(lib.svbrdf_b200_loss_host(ctx, h_in.data_ptr(), 10, h_tg.data_ptr(), 10, B, rec.data_ptr(), N, -1.0, res, h_g.data_ptr()))
        assert res[0] == float(o12[0]) and torch.equal(h_g, g10.cpu())
        h_e, h_ge = enc.cpu(), torch.empty(B, 9, size, size)
        _cabi.check(lib.svbrdf_b200_loss_host(ctx, h_e.data_ptr(), 9, h_tg.data_ptr(), 10, B, rec.data_ptr(), N, 0.1, res, h_ge.data_ptr()))
        assert res[0] == float(oe10[0]) and res[1] == float(oe10[1]) and torch.equal(h_ge, ge10.cpu())
    finally:
        lib.svbrdf_b200_ctx_destroy(ctx)


def test_pageable_host_tensors_through_the_python_api(env):
    """RenderingLoss on large pageable CPU tensors (what a caller without device tensors does): staged through pinned
    double buffers both ways (staging.upload / download); same result as the device path."""
    import svbrdf_estimation_b200 as S
    from svbrdf_estimation_b200 import staging
    lib, _cabi, E = env
    B, size = 6, 256                                     # 18.9 MB per tensor: above the staging threshold, several chunks with a small chunk size
    inp, tgt = synthetic_maps(B, size, 51), synthetic_maps(B, size, 52)
    rec = E.sample_loss_configs(B)
    old = staging.CHUNK_BYTES
    staging.CHUNK_BYTES = 5 << 20
    staging._state.clear()
    try:
        x = inp.clone().requires_grad_(True)
        loss = S.rendering_loss_with_records(x, tgt, rec)
        loss.backward()
        assert loss.device.type == "cpu" and x.grad.device.type == "cpu"
        xd = inp.cuda().requires_grad_(True)
        ld = S.rendering_loss_with_records(xd, tgt.cuda(), rec)
        ld.backward()
        assert float(loss) == float(ld) and torch.equal(x.grad, xd.grad.cpu())
        t = torch.rand(3_000_001)
        assert torch.equal(staging.download(staging.upload(t)), t)
    finally:
        staging.CHUNK_BYTES = old
        staging._state.clear()


def test_bench_line_contract(env):
    """`python bench.py` prints ONE JSON line with the contract's keys; roofline / e2e / train_step_c5 objects are filled from
    live measurements (short run: 5 steps, a global batch of 8 for the configs[4] leg)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline",
                          "--c5-global-batch", "8", "--c5-steps", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["unit"] == "G evals/s" and d["value"] > 50 and d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "c2" and d["gpu_launches"] == 10
    r = d["roofline"]
    assert r["bound"] == "fp32" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.3 < r["frac"] < 1.0
    assert r["traffic"] is None or 0.5 * 604e6 < r["traffic"] < 1.5 * 604e6
    assert d["executed_work"]["sass_model_matches_this_build"] is True and d["executed_work"]["flop_per_eval_executed"] > 100
    e = d["e2e"]
    assert e["value"] > 1 and e["h2d_bytes_per_step"] == 2 * 64 * 10 * 256 * 256 * 4 and e["d2h_bytes_per_step"] == 64 * 10 * 256 * 256 * 4 + 12
    assert e["matches_device_entry"] is True and e["maps12_variant"]["h2d_bytes_per_step"] == 2 * 64 * 12 * 256 * 256 * 4
    assert e["python_api_pageable_tensors"]["value"] > 0.1
    c = d["train_step_c5"]
    assert c["global_batch"] == 8 and c["step_ms"] > 0 and 79 < c["params_M"] < 81 and c["scaling"] == "strong"
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
