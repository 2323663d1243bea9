"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``renderers``, ``losses``, ``environment`` and ``utils`` from
``/root/reference/development/multiImage_pytorch`` (an empty ``pyredner`` module is put into
``sys.modules`` because ``renderers.py:4`` imports it unconditionally; only the out-of-scope
``RednerRenderer`` uses it), evaluates ``LocalRenderer.render`` and ``RenderingLoss`` in fp32
and - through ``torch.set_default_dtype(torch.float64)`` - in fp64 on small seeded inputs and
writes ``*.npz`` files.  Nothing from the reference is copied; only its outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/development/multiImage_pytorch"


def load_reference():
    sys.modules.setdefault("pyredner", types.ModuleType("pyredner"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import environment as ref_env      # noqa: E402
    import losses as ref_losses        # noqa: E402
    import renderers as ref_renderers  # noqa: E402
    import utils as ref_utils          # noqa: E402
    return ref_env, ref_losses, ref_renderers, ref_utils


def synthetic_maps(batch, size, seed, stress=False):
    """Same distributions as SURVEY.md §8(d): unit upper-hemisphere normals, diffuse/specular
    U(0,1), roughness U(0.1,1) replicated x3.  ``stress`` adds the hard cases of §8(c)."""
    g = torch.Generator("cpu").manual_seed(seed)
    xy = torch.randn(batch, 2, size, size, generator=g) * 0.3
    n = torch.cat((xy, torch.ones(batch, 1, size, size)), dim=1)
    n = n / n.norm(dim=1, keepdim=True)
    d = torch.rand(batch, 3, size, size, generator=g)
    s = torch.rand(batch, 3, size, size, generator=g)
    if not stress:
        r = (torch.rand(batch, 1, size, size, generator=g) * 0.9 + 0.1).repeat(1, 3, 1, 1)
    else:
        r = torch.rand(batch, 3, size, size, generator=g)           # independent channels
        kill = torch.rand(batch, 3, size, size, generator=g)
        r = torch.where(kill < 0.01, torch.zeros_like(r), r)        # exact zeros
        r = torch.where((kill >= 0.01) & (kill < 0.02), r * 1e-3, r)  # below the 1e-3 clamp
        scale = 0.5 + torch.rand(batch, 1, size, size, generator=g)  # non-unit normals
        flip = torch.rand(batch, 1, size, size, generator=g) < 0.05  # a few facing away
        n = n * scale
        n[:, 2:3] = torch.where(flip, -n[:, 2:3], n[:, 2:3])
    return torch.cat((n, d, r, s), dim=1).contiguous()


def scenes_from_configs(ref_env, cfg_row):
    """[N,9] fp32 -> list of reference Scene objects with python-float lists (valid under
    either default dtype)."""
    out = []
    for k in range(cfg_row.shape[0]):
        v = [float(x) for x in cfg_row[k]]
        out.append(ref_env.Scene(ref_env.Camera(v[0:3]), ref_env.Light(v[3:6], v[6:9])))
    return out


def sample_configs(ref_env, seed, batch, n_random, n_specular):
    """Reference sampler, reference draw order (losses.py:35), packed to [B,N,9] fp32."""
    torch.manual_seed(seed)
    rows = []
    for _ in range(batch):
        scenes = ref_env.generate_random_scenes(n_random) + ref_env.generate_specular_scenes(n_specular)
        rows.append(torch.stack([torch.cat((torch.as_tensor(s.camera.pos, dtype=torch.float32),
                                            torch.as_tensor(s.light.pos, dtype=torch.float32),
                                            torch.as_tensor(s.light.color, dtype=torch.float32)))
                                 for s in scenes]))
    return torch.stack(rows)


class _FixedScenes:
    """Replaces the two sampler functions of the reference's ``environment`` module so that
    ``RenderingLoss.forward`` consumes pre-recorded configurations."""

    def __init__(self, ref_env, configs, n_random):
        self.ref_env, self.configs, self.n_random, self.b = ref_env, configs, n_random, 0

    def random(self, count):
        assert count == self.n_random
        return scenes_from_configs(self.ref_env, self.configs[self.b, :count])

    def specular(self, count):
        row = self.configs[self.b, self.n_random:]
        assert count == row.shape[0]
        self.b += 1
        return scenes_from_configs(self.ref_env, row)


def reference_loss(ref, dtype, inp, tgt, configs, n_random):
    ref_env, ref_losses, ref_renderers, _ = ref
    torch.set_default_dtype(dtype)
    keep = (ref_env.generate_random_scenes, ref_env.generate_specular_scenes)
    try:
        fixed = _FixedScenes(ref_env, configs, n_random)
        ref_env.generate_random_scenes, ref_env.generate_specular_scenes = fixed.random, fixed.specular
        loss_mod = ref_losses.RenderingLoss(ref_renderers.LocalRenderer())
        loss_mod.random_configuration_count = n_random
        loss_mod.specular_configuration_count = configs.shape[1] - n_random
        x = inp.to(dtype).clone().requires_grad_(True)
        loss = loss_mod(x, tgt.to(dtype))
        loss.backward()
        renderer = ref_renderers.LocalRenderer()
        renders = torch.stack([torch.cat([renderer.render(s, x.detach()[b])
                                          for s in scenes_from_configs(ref_env, configs[b])])
                               for b in range(x.shape[0])])
        return loss.detach().numpy(), x.grad.numpy(), renders.numpy()
    finally:
        ref_env.generate_random_scenes, ref_env.generate_specular_scenes = keep
        torch.set_default_dtype(torch.float32)


def main():
    ref = load_reference()
    ref_env, ref_losses, ref_renderers, ref_utils = ref

    # 1. sampler pins -------------------------------------------------------------------
    scenes = dict(seed313_b4_r3_s6=sample_configs(ref_env, 313, 4, 3, 6).numpy(),
                  seed7_b2_r9_s18=sample_configs(ref_env, 7, 2, 9, 18).numpy())
    # more shapes (odd counts leave a cached Box-Muller sample in the generator; 0 random / 0 specular), each followed by
    # four draws from the global generator: pins the generator state the reference leaves behind
    for seed, batch, nr, ns in ((1, 3, 3, 6), (2, 5, 2, 3), (3, 4, 1, 5), (4, 2, 0, 7), (5, 3, 4, 0), (6, 2, 9, 15), (11, 16, 3, 6)):
        key = "seed%d_b%d_r%d_s%d" % (seed, batch, nr, ns)
        scenes[key] = sample_configs(ref_env, seed, batch, nr, ns).numpy()
        scenes[key + "_next"] = torch.rand(4).numpy()
    np.savez(os.path.join(HERE, "scenes.npz"), **scenes)
    if "--only-scenes" in sys.argv:
        return
    torch.manual_seed(99)
    dirs = ref_utils.generate_normalized_random_direction(5, 0.001, 0.1).numpy()
    np.savez(os.path.join(HERE, "directions.npz"), seed99_count5=dirs)

    # 2. render() under the two fixed scenes the reference's scripts/notebooks use -------
    f32 = lambda *v: [float(np.float32(x)) for x in v]  # noqa: E731  (fp32-representable)
    fixed = np.array([f32(0, -1, 2) + f32(0, 0, 2) + f32(50, 50, 50),          # renderers.py:284
                      f32(0, -1, 1.4) + f32(1, 1, 1.7) + f32(30, 30, 30)],      # final-viz.ipynb cell 11
                     dtype=np.float32)
    maps = synthetic_maps(2, 24, 1001)
    out = {"maps": maps.numpy(), "configs": fixed}
    for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        torch.set_default_dtype(dtype)
        renderer = ref_renderers.LocalRenderer()
        scenes = scenes_from_configs(ref_env, torch.from_numpy(fixed))
        out["render4d_" + name] = torch.stack([renderer.render(s, maps.to(dtype)) for s in scenes]).numpy()  # [2,B,3,H,W]
        out["render3d_" + name] = renderer.render(scenes[0], maps[0].to(dtype)).numpy()                       # [1,3,H,W]
        torch.set_default_dtype(torch.float32)
    np.savez(os.path.join(HERE, "render_fixed.npz"), **out)

    # 3. RenderingLoss fwd+bwd, bench distribution and stress distribution ---------------
    for tag, size, stress, seed_in, seed_tg in (("bench", 24, False, 1002, 2002), ("stress", 16, True, 1003, 2003)):
        inp, tgt = synthetic_maps(2, size, seed_in, stress), synthetic_maps(2, size, seed_tg, stress)
        cfg = sample_configs(ref_env, 313, 2, 3, 6)
        out = {"input": inp.numpy(), "target": tgt.numpy(), "configs": cfg.numpy()}
        for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
            loss, grad, renders = reference_loss(ref, dtype, inp, tgt, cfg, 3)
            out["loss_" + name], out["grad_" + name], out["renders_" + name] = loss, grad, renders
        np.savez(os.path.join(HERE, "loss_%s.npz" % tag), **out)

    # 4. N != 9 (27 configs = 9 random + 18 specular), tiny maps ------------------------
    inp, tgt = synthetic_maps(1, 8, 1004), synthetic_maps(1, 8, 2004)
    cfg = sample_configs(ref_env, 7, 1, 9, 18)
    out = {"input": inp.numpy(), "target": tgt.numpy(), "configs": cfg.numpy()}
    for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        loss, grad, _ = reference_loss(ref, dtype, inp, tgt, cfg, 9)
        out["loss_" + name], out["grad_" + name] = loss, grad
    np.savez(os.path.join(HERE, "loss_n27.npz"), **out)

    # 5. MixedLoss value (losses.py:54-63) on the bench fixture, sampled scenes -----------
    inp, tgt = synthetic_maps(2, 24, 1002), synthetic_maps(2, 24, 2002)
    cfg = sample_configs(ref_env, 313, 2, 3, 6)
    torch.manual_seed(313)
    mixed = ref_losses.MixedLoss(ref_renderers.LocalRenderer())
    x = inp.clone().requires_grad_(True)
    val = mixed(x, tgt)
    val.backward()
    l1 = ref_losses.SVBRDFL1Loss()(inp, tgt)
    np.savez(os.path.join(HERE, "mixed.npz"), loss_f32=val.detach().numpy(), grad_f32=x.grad.numpy(),
             l1_f32=l1.numpy(), seed=np.int64(313))

    # 6. model-output epilogue: decode_svbrdf + [0,1] mapping (utils.py:73-98, models.py:334-346) -------
    g = torch.Generator("cpu").manual_seed(77)
    enc = torch.rand(2, 9, 12, 12, generator=g) * 2 - 1
    dec = ref_utils.decode_svbrdf(enc)
    n, d, r, s = ref_utils.unpack_svbrdf(dec)
    dec = ref_utils.pack_svbrdf(n, ref_utils.encode_as_unit_interval(d), ref_utils.encode_as_unit_interval(r),
                                ref_utils.encode_as_unit_interval(s))
    np.savez(os.path.join(HERE, "decode.npz"), encoded=enc.numpy(), decoded_f32=dec.numpy())

    # 7. dataset input synthesis (dataset.py:162-221): scenes, renders, noise, clamp -----------------------
    for stub in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    import dataset as ref_dataset  # noqa: E402
    captured = []

    class Recorder(ref_renderers.LocalRenderer):
        def render(self, scene, svbrdf):
            captured.append(np.concatenate([np.asarray(scene.camera.pos, dtype=np.float32),
                                            np.asarray(scene.light.pos, dtype=np.float32),
                                            np.asarray(scene.light.color, dtype=np.float32)]))
            return super().render(scene, svbrdf)

    keep_cls = ref_renderers.LocalRenderer
    ref_renderers.LocalRenderer = Recorder
    try:
        maps = synthetic_maps(1, 16, 1005)[0]
        out = {"svbrdf": maps.numpy()}
        for tag, aug in (("plain", False), ("aug", True)):
            captured.clear()
            torch.manual_seed(4242)
            fake = types.SimpleNamespace(use_augmentation=aug)
            imgs = ref_dataset.SvbrdfDataset.render_inputs(fake, maps, 3)
            out["inputs_" + tag] = imgs.numpy()
            out["records_" + tag] = np.stack(captured)
        np.savez(os.path.join(HERE, "dataset_inputs.npz"), seed=np.int64(4242), **out)
    finally:
        ref_renderers.LocalRenderer = keep_cls

    # 8. real materials: 48x48 crops of the reference's toy data (mip/data/{train,test}/*.png), maps read like
    #    dataset.py:105-136 (PNG/255; normals = 2*v - 1, NOT re-normalised; roughness as stored) ---------------
    from PIL import Image

    def read_maps(path, top, left, size=48):
        full = torch.from_numpy(np.asarray(Image.open(path).convert("RGB"), dtype=np.float32) / 255.0).permute(2, 0, 1)
        parts = torch.cat(full.unsqueeze(0).chunk(14, dim=-1), 0)               # 10 inputs + 4 maps, 256 px each
        n, d, r, sp = parts[10] * 2 - 1, parts[11], parts[12], parts[13]
        return torch.cat((n, d, r, sp), dim=0)[:, top:top + size, left:left + size].contiguous()

    data = os.path.join(REF, "data")
    brick = os.path.join(data, "train", "0_10_brick_uneven_stones_0.png")
    parquet = os.path.join(data, "train", "10_1_parquet_floor_0.png")
    foam = os.path.join(data, "test", "11_20_SynteticFoam_0.png")
    inp = torch.stack((read_maps(brick, 100, 60), read_maps(foam, 30, 150)))
    tgt = torch.stack((read_maps(parquet, 64, 64), read_maps(brick, 10, 10)))
    cfg = sample_configs(ref_env, 2024, 2, 3, 6)
    out = {"input": inp.numpy(), "target": tgt.numpy(), "configs": cfg.numpy()}
    for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        loss, grad, renders = reference_loss(ref, dtype, inp, tgt, cfg, 3)
        out["loss_" + name] = loss
        out["grad_" + name] = grad.astype(np.float32)             # stored in fp32: 6e-8 relative, far below the tolerances
    np.savez_compressed(os.path.join(HERE, "loss_real.npz"), **out)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-20s %8.1f KB" % (f, os.path.getsize(os.path.join(HERE, f)) / 1024))


if __name__ == "__main__":
    main()
